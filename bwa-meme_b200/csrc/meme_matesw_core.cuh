// meme_matesw_core.cuh -- mate-rescue Smith-Waterman, the per-pair procedure.
//
// What is computed is exactly what the reference's mem_sam_pe_batch (src/bwamem_pair.cpp:L719-817) leaves in aln[] for a pair:
// kswv::kswv512_u8 / kswv512_16 (src/kswv.cpp:L372-712, L934-1218) in phase 0 -- a full (unbanded) local alignment of the mate
// against the rescue window giving score, te, qe, and the best row maximum at least `score` rows away from te as score2 / te2
// -- then, for pairs that reach min_seed_len * a, the same recurrence over the reversed prefixes up to the first row that
// reaches the score again, giving tb / qb.  The vector kernels run 64 / 32 pairs in lockstep, one pair per SIMD lane; a lane's
// result does not depend on its neighbours (oracle/matesw_oracle.c states it per lane and is pinned against the reference run
// in differently cut batches), so the mapping here is the same one -- ONE THREAD PER PAIR -- with the lane's arrays laid out
// [element][thread]:
//   * the row state (H of the previous row for the diagonal, F of this row) is one 32-bit word per query column in shared
//     memory, the query code of the column (A, C, G, T, N, or the zero-scoring filler the reference pads a query with up to a
//     multiple of 16 / 8 columns, L291-303 / L857-870) a byte beside it;
//   * R = 4 reference rows are swept together: per column one word is loaded and one stored for four cells, E and the
//     left neighbours of the four rows live in registers;
//   * the substitution score is one PRMT on the row's packed score bytes, the cell update is DPX: VIADDMNMX with relu for
//     M = max(min(Hdiag + s, cap), 0) (cap = the unsigned-byte saturation of the 8-bit class, L63-73), E and F; E's loop-carried
//     chain is ONE instruction (max(H - oe, E - e) = max(max(M, F) - oe, E - e) because oe >= e);
//   * the row maximum with the reference's "first maximum wins" rule is a max over (h * 4096 + 4095 - j);
//   * row maxima go to a lane-interleaved int16 array in global memory (written once per row, read back once for score2).
// The 8-bit class differs from the 16-bit class only in cap, in how a suppressed row maximum is encoded (0 / -1) and in the
// overflow rules, so both run through the same code.
//
// Plain C++ below MSW_HD, so that tests/host_sim compiles the same source with g++.
#pragma once
#include <stdint.h>
#include "meme_bsw_is.cuh"            // BSW_ADDMAX, BSW_PRMT

#if defined(__CUDA_ARCH__)
#define MSW_HD __host__ __device__ __forceinline__
#define MSW_ADDMIN_RELU(a, b, c) __viaddmin_s32_relu((a), (b), (c))     /* max(min(a + b, c), 0) */
#else
#define MSW_HD inline
#define MSW_ADDMIN_RELU(a, b, c) ((((a) + (b)) < (c) ? ((a) + (b)) : (c)) > 0 ? (((a) + (b)) < (c) ? ((a) + (b)) : (c)) : 0)
#endif

namespace msw {

enum { XBYTE = 0x10000, XSTOP = 0x20000, XSUBO = 0x40000, XSTART = 0x80000 };      // src/ksw.h:L31-34
enum { MAX_COLS = 512, MAX_ROWS = 32000, ROWS = 4 };

struct Opt { int o_del, e_del, o_ins, e_ins, a, b; };
struct Out { int score, te, qe, score2, te2, tb, qb; };                             // kswr_t, src/ksw.h:L45-50

struct Scratch {
    uint32_t *row;      // word j at row[j * stride], j in [0, cols]
    uint8_t *qrow;      // query code of column j at qrow[j * stride]
    int stride;
    int16_t *rm;        // row maximum i at rm[i * rstride]
    int rstride;
};

MSW_HD int quanta(int len2, bool byte) { const int u = byte ? 16 : 8; return (len2 + u - 1) / u * u; }

// scores of a reference base against query A, C, G, T (bytes 0-3 of lo), N and the filler (bytes 0, 1 of hi)
MSW_HD uint32_t score_bytes(const Opt &o, int tc)
{
    if (tc >= 4) return 0xffffffffu;                                                // w_ambig = -1 (kswv ctor, DEFAULT_AMBIG)
    const uint32_t mm = (uint32_t)(-o.b) & 0xffu;
    return (mm * 0x01010101u) ^ ((mm ^ ((uint32_t)o.a & 0xffu)) << (8 * tc));
}

struct PassOut { int gmax, te, qe, rows; };

// One pass of the kernel over rows [0, len1).  rev_te >= 0 (phase 1): row i reads ref[rev_te - i] while i <= rev_te, and the
// caller has filled qrow with the reversed query prefix.  write_rm: store the row maxima (phase 0).
MSW_HD void pass(const Opt &o, const uint8_t *ref, int len1, int rev_te, int cols, int xtra, bool byte, bool write_rm, const Scratch &S,
                 PassOut &P, uint64_t &cells)
{
    const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins, e_del = o.e_del, e_ins = o.e_ins;
    int shift = 0;
    if (byte) {                                                                     // src/kswv.cpp:L394-402
        int mn = o.a < -o.b ? o.a : -o.b;
        mn = mn < -1 ? mn : -1;
        shift = 256 - (int)(uint8_t)(int8_t)mn;
    }
    const int cap = byte ? 255 - shift : 0x3fffffff;
    const int lim = byte ? 255 : 32767;
    int v = (xtra & XSUBO) ? (xtra & 0xffff) : 0x10000;
    const bool has_minsc = v <= lim; const int minsc = v;
    v = (xtra & XSTOP) ? (xtra & 0xffff) : 0x10000;
    const bool has_endsc = v <= lim; const int endsc = v;
    const int INV = byte ? 0 : -1;
    const int stride = S.stride;
    for (int j = 0; j <= cols; j++) S.row[(size_t)j * stride] = 0;
    int gmax = 0, te = -1, qe = 0, pimax = 0, i = 0;
    bool alive = true, mask = false, minsc_ok = false;
    while (i < len1 && alive) {
        const int nr = len1 - i < ROWS ? len1 - i : ROWS;
        uint32_t plo[ROWS];
        int e[ROWS], hl[ROWS], key[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            int tc = 4;
            if (r < nr) { const int ii = i + r; tc = ref[ii <= rev_te ? rev_te - ii : ii]; }
            plo[r] = score_bytes(o, tc);
            e[r] = 0; hl[r] = 0; key[r] = 0;
        }
        const uint32_t phi = 0x00ffu;
        uint32_t *rp = S.row + stride;                                              // word 1 = column 0
        const uint8_t *qp = S.qrow;
        int dg = 0;                                                                 // H(i-1, j-1): column -1 is 0
        if (nr == ROWS) {
            for (int j = 0; j < cols; j++, rp += stride, qp += stride) {
                const uint32_t wd = *rp;
                const uint32_t sel = (uint32_t)*qp * 0x1111u + 0x8880u;
                int d = dg, f = (int)(wd >> 16);
                dg = (int)(wd & 0xffffu);                                           // H(i-1, j): next column's diagonal
                const int kj = 4095 - j;
                int h = 0;
#pragma unroll
                for (int r = 0; r < ROWS; r++) {
                    const int s = BSW_PRMT(plo[r], phi, sel);
                    const int m = MSW_ADDMIN_RELU(d, s, cap);
                    const int mf = m > f ? m : f;
                    h = mf > e[r] ? mf : e[r];
                    const int kk = h * 4096 + kj;
                    key[r] = kk > key[r] ? kk : key[r];
                    e[r] = BSW_ADDMAX(e[r], -e_ins, BSW_ADDMAX(mf, -oe_ins, 0));
                    f = BSW_ADDMAX(f, -e_del, BSW_ADDMAX(h, -oe_del, 0));          // F of row r + 1 at this column
                    d = hl[r];                                                      // H(r, j-1): the diagonal of row r + 1
                    hl[r] = h;
                }
                *rp = (uint32_t)h | ((uint32_t)f << 16);
            }
        } else {
            for (int j = 0; j < cols; j++, rp += stride, qp += stride) {
                const uint32_t wd = *rp;
                const uint32_t sel = (uint32_t)*qp * 0x1111u + 0x8880u;
                int d = dg, f = (int)(wd >> 16);
                dg = (int)(wd & 0xffffu);
                const int kj = 4095 - j;
                int h = 0;
                for (int r = 0; r < nr; r++) {
                    const int s = BSW_PRMT(plo[r], phi, sel);
                    const int m = MSW_ADDMIN_RELU(d, s, cap);
                    const int mf = m > f ? m : f;
                    h = mf > e[r] ? mf : e[r];
                    const int kk = h * 4096 + kj;
                    key[r] = kk > key[r] ? kk : key[r];
                    e[r] = BSW_ADDMAX(e[r], -e_ins, BSW_ADDMAX(mf, -oe_ins, 0));
                    f = BSW_ADDMAX(f, -e_del, BSW_ADDMAX(h, -oe_del, 0));
                    d = hl[r];
                    hl[r] = h;
                }
                *rp = (uint32_t)h | ((uint32_t)f << 16);
            }
        }
        cells += (uint64_t)nr * (uint64_t)cols;
        // per-row bookkeeping, in row order (Block I / II, src/kswv.cpp:L509-548, L1062-1095)
        for (int r = 0; r < nr; r++) {
            const int imax = key[r] >> 12, iqe = 4095 - (key[r] & 4095);
            if (i > 0) {
                const bool msk = imax > pimax || mask;
                if (write_rm) S.rm[(size_t)(i - 1) * S.rstride] = (int16_t)((!msk && minsc_ok) ? pimax : INV);
                mask = !msk;
            }
            pimax = imax;
            minsc_ok = has_minsc && imax >= minsc;
            if (imax > gmax) { gmax = imax; te = i; qe = iqe; }
            i++;
            bool stop = has_endsc && gmax >= endsc;
            if (byte && gmax + shift >= 255) stop = true;
            if (stop) { alive = false; break; }
        }
    }
    if (i > 0 && write_rm) S.rm[(size_t)(i - 1) * S.rstride] = (int16_t)((!mask && minsc_ok && alive) ? pimax : INV);
    P.gmax = gmax; P.te = te; P.qe = qe; P.rows = i;
}

// the query codes of a pass into the scratch: q[j] (phase 0) or q[last - j] (phase 1: the reversed prefix), then the filler
MSW_HD void fill_query(const uint8_t *q, int len2, int last, int cols, const Scratch &S)
{
    for (int j = 0; j < cols; j++) {
        int c = 5;
        if (j < len2) { c = q[last >= 0 ? last - j : j]; if (c > 4) c = 4; }
        S.qrow[(size_t)j * S.stride] = (uint8_t)c;
    }
}

// mem_sam_pe_batch for one pair.  Limits (checked by the caller): len2 <= MAX_COLS - 16, len1 <= MAX_ROWS, o_del, o_ins >= 0.
MSW_HD void pair(const Opt &o, const uint8_t *ref, int len1, const uint8_t *q, int len2, int xtra, const Scratch &S, Out &A, uint64_t &cells)
{
    const bool byte = (xtra & XBYTE) != 0;                                          // sort_classify, src/bwamem.cpp:L1797-1826
    int shift = 0;
    if (byte) { int mn = o.a < -o.b ? o.a : -o.b; mn = mn < -1 ? mn : -1; shift = 256 - (int)(uint8_t)(int8_t)mn; }
    PassOut P;
    int cols = quanta(len2, byte);
    fill_query(q, len2, -1, cols, S);
    pass(o, ref, len1, -1, cols, xtra, byte, true, S, P, cells);
    A.tb = A.qb = -1; A.score2 = -1; A.te2 = -1;
    A.te = P.te; A.qe = byte ? (P.qe & 0xff) : P.qe;
    A.score = byte ? (P.gmax + shift < 255 ? P.gmax : 255) : P.gmax;
    {   // score2 / te2 (src/kswv.cpp:L611-700, L1143-1206)
        int qmax = o.a > -o.b ? o.a : -o.b;
        qmax = qmax > -1 ? qmax : -1;
        const int val = (P.gmax + qmax - 1) / qmax, low = P.te - val, high = P.te + val;
        int cur = byte ? 0 : -1, te2 = -1;
        for (int k = 0; k < low && k < P.rows; k++) { const int x = S.rm[(size_t)k * S.rstride]; if (x > cur) { cur = x; te2 = k; } }
        for (int k = high + 1 > 0 ? high + 1 : 0; k < P.rows; k++) { const int x = S.rm[(size_t)k * S.rstride]; if (x > cur) { cur = x; te2 = k; } }
        if (byte) { if (A.score != 255) { A.score2 = cur == 0 ? -1 : cur; A.te2 = te2; } }
        else { A.score2 = cur; A.te2 = te2; }
    }
    if ((xtra & XSTART) == 0 || ((xtra & XSUBO) && A.score < (xtra & 0xffff))) return;      // src/bwamem_pair.cpp:L769-776
    const int len2b = A.qe + 1;
    cols = quanta(len2b, byte);
    fill_query(q, len2b, A.qe, cols, S);
    pass(o, ref, len1, A.te, cols, XSTOP | A.score, byte, false, S, P, cells);
    if (A.score == P.gmax) { A.tb = A.te - P.te; A.qb = A.qe - (byte ? (P.qe & 0xff) : P.qe); }   // src/kswv.cpp:L578-581
}

}  // namespace msw
