// meme_bsw_is.cuh -- kernel B, inter-sequence form: one THREAD per extension pair, the score row in shared memory.
//
// What is computed is exactly BandedPairWiseSW::scalarBandedSWA (reference src/bandedSWA.cpp:L116-237), the
// definition the reference's getScores8/16 vectorise with one pair per SIMD lane (L412-4870).  On the GPU the same
// mapping -- one pair per lane of a warp -- is what the extension workload wants: the adaptive band of a seed
// extension is narrow (a few tens of cells per row), so a warp that splits ONE row across its lanes (k_bsw<C> in
// meme_bsw.cu, kept as the int32 overflow class) spends most lanes on cells outside the band and pays a prefix scan,
// three REDUX and several shuffles per row.  Here every lane walks only the live cells [beg, end) of its own pair:
//   * the row (reference eh[]: H(i-1, j-1) and E(i, j) per column) is one 32-bit word per column -- 16 bits H | 16 bits
//     E -- laid out [column][thread] in shared memory, so a warp's 32 loads/stores of one column hit 32 different
//     banks; the query base of the column sits in a byte array of the same layout;
//   * the substitution score is one PRMT: the scores of the row's reference base against A, C, G, T, N are packed as
//     five signed bytes once per row, and the column's query base selects one (with sign replication);
//   * the cell update is DPX: VIADDMNMX for E and F, VIMNMX3 for H; the row maximum with the reference's
//     "last maximum wins" rule is one max over (h * 4096 + j);
//   * pairs are sorted by (query length, reference length) on the device so that the 32 pairs of a warp have similar
//     trip counts, and launched per query-length class so that short rows leave room for more resident warps;
//   * the band retry of the extension driver (w, 2w; reference src/bwamem.cpp:L2968-3161, MAX_BAND_TRY) is decided
//     on the device: a pair whose result fails the acceptance test is recomputed in place by the same thread.
// Scores must fit 16 bits (h0 + qlen * match < 60000) and the query 320 bases; anything else takes the overflow class.
// The per-pair function is __host__ __device__ so that tests/host_sim can run it against the oracle without a GPU.
#pragma once
#include "meme_common.cuh"

#if defined(__CUDA_ARCH__)
#define BSW_HD __host__ __device__ __forceinline__
#define BSW_ADDMAX(a, b, c) __viaddmax_s32((a), (b), (c))          /* max(a + b, c) */
#define BSW_MAX3(a, b, c) __vimax3_s32((a), (b), (c))
// prmt.b32 in its default mode: bit 3 of a selector nibble replicates the sign of the selected byte
// (__byte_perm masks the selector to 3 bits per nibble, so the PTX instruction is written out)
__device__ __forceinline__ int bsw_prmt_dev(unsigned lo, unsigned hi, unsigned sel)
{
    int d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sel));
    return d;
}
#define BSW_PRMT(lo, hi, sel) bsw_prmt_dev((lo), (hi), (sel))
#else
#define BSW_HD inline
#define BSW_ADDMAX(a, b, c) (((a) + (b)) > (c) ? ((a) + (b)) : (c))
#define BSW_MAX3(a, b, c) ((a) > (b) ? ((a) > (c) ? (a) : (c)) : ((b) > (c) ? (b) : (c)))
static inline int bsw_prmt_host(unsigned lo, unsigned hi, unsigned sel)
{   // PTX prmt, default mode: nibble k of sel picks a byte of {hi, lo}; bit 3 of the nibble replicates its sign instead
    unsigned long long v = ((unsigned long long)hi << 32) | lo;
    unsigned r = 0;
    for (int k = 0; k < 4; k++) {
        unsigned nib = (sel >> (4 * k)) & 0xf, byte = (unsigned)(v >> (8 * (nib & 7))) & 0xff;
        if (nib & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * k);
    }
    return (int)r;
}
#define BSW_PRMT(lo, hi, sel) bsw_prmt_host((lo), (hi), (sel))
#endif

struct BswIsConst {
    int o_del, e_del, o_ins, e_ins, zdrop, end_bonus, a, b;
};

#define BSW_IS_MAX_SCORE 60000
#define BSW_IS_MAX_QLEN 320

struct BswIsOut { int score, tle, gtle, qle, gscore, max_off; };

// The reference side of a pair: bytes uploaded by the caller (codes 0..4), or -- when the caller's reference is the
// indexed text that already sits in HBM -- a position in the 2-bit fwd+RC text and a direction (left extensions walk the
// text backwards: src/bwamem.cpp:L2798 copies rseq reversed).  The text carries no N (bntseq randomises them at index
// time, src/bntseq.cpp:L299), so both forms give the same codes.
struct BswTarget {
    const uint8_t *bytes;
    const u64 *pac; u64 pos; int dir;
};
BSW_HD int bsw_target_at(const BswTarget &T, int i)
{
    if (T.bytes) return T.bytes[i];
    const u64 p = T.dir > 0 ? T.pos + (u64)i : T.pos - (u64)i;
    return (int)((T.pac[p >> 5] >> (2 * (p & 31))) & 3);
}

// One pair.  row: the thread's score row, word j at row[j * stride], j in [0, qlen]; qrow: its query bases, byte j at
// qrow[j * stride]; the caller has NOT initialised either.
BSW_HD void bsw_is_one(const BswIsConst &P, int w, const uint8_t *query, int qlen, const BswTarget &target, int tlen, int h0,
                       u32 *row, uint8_t *qrow, const int stride, BswIsOut &out, u64 &cells)
{
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins, e_del = P.e_del, e_ins = P.e_ins;
    // ---- first row (reference L143-145) and the query bases: 16 columns at a time, their loads in flight together ----
    {
        int hv = h0;
        for (int j0 = 0; j0 <= qlen; j0 += 16) {
            uint8_t qb[16];
#pragma unroll
            for (int k = 0; k < 16; k++) qb[k] = j0 + k < qlen ? query[j0 + k] : (uint8_t)4;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int j = j0 + k;
                if (j <= qlen) {
                    int v;
                    if (j == 0) v = h0;
                    else if (j == 1) { hv = h0 > oe_ins ? h0 - oe_ins : 0; v = hv; }
                    else { if (hv > e_ins) hv -= e_ins; else hv = 0; v = hv; }
                    // (the reference stops at the first H <= e_ins and leaves zeros behind: same values)
                    row[j * stride] = (u32)v;
                    qrow[j * stride] = qb[k];
                }
            }
        }
    }
    // ---- band (L147-156) ----
    {
        const int mx = P.a > 0 ? P.a : 0;
        int max_ins = (int)((double)(qlen * mx + P.end_bonus - P.o_ins) / P.e_ins + 1.);
        if (max_ins < 1) max_ins = 1;
        if (w > max_ins) w = max_ins;
        int max_del = (int)((double)(qlen * mx + P.end_bonus - P.o_del) / P.e_del + 1.);
        if (max_del < 1) max_del = 1;
        if (w > max_del) w = max_del;
    }
    int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen;
    int tc_next = tlen > 0 ? bsw_target_at(target, 0) : 0;
    for (int i = 0; i < tlen; i++) {
        const int tc = tc_next;
        if (i + 1 < tlen) tc_next = bsw_target_at(target, i + 1);  // the next row's base arrives while this row is computed
        int f = 0, h1, key = -1;
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        if (beg == 0) { h1 = h0 - (P.o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; } else h1 = 0;
        // scores of this row's reference base against query A, C, G, T (bytes 0-3 of plo) and N (byte 0 of phi);
        // N against anything is -1 (bwa_fill_scmat, reference src/bwa.cpp:L262-271)
        u32 plo = 0xffffffffu;
        const u32 phi = 0xffu;
        if (tc < 4) {
            const u32 mm = (u32)(-P.b) & 0xffu;
            plo = (mm * 0x01010101u) ^ ((mm ^ ((u32)P.a & 0xffu)) << (8 * tc));
        }
        u32 *rp = row + (size_t)beg * stride;
        const uint8_t *qp = qrow + (size_t)beg * stride;
        for (int j = beg; j < end; j++, rp += stride, qp += stride) {
            const u32 wd = *rp;
            const u32 q = *qp;
            const int d = (int)(wd & 0xffffu), e0 = (int)(wd >> 16);
            const int s = BSW_PRMT(plo, phi, q * 0x1111u + 0x8880u);      // byte q, sign-extended
            int M = d + s;
            if (d == 0) M = 0;                                     // L184: M = M ? M + s : 0
            const int h = BSW_MAX3(M, e0, f);
            const int kk = h * 4096 + j;                           // last maximum wins (L188-189)
            key = kk > key ? kk : key;
            const int e1 = BSW_ADDMAX(e0, -e_del, BSW_ADDMAX(M, -oe_del, 0));
            f = BSW_ADDMAX(f, -e_ins, BSW_ADDMAX(M, -oe_ins, 0));
            *rp = (u32)h1 | ((u32)e1 << 16);                       // eh[j] = {H(i, j-1), E(i+1, j)}
            h1 = h;
        }
        cells += (u64)(end > beg ? end - beg : 0);
        {   // eh[end] = {h1, 0} (L199), also when the band is empty; the loop variable ends at max(beg, end)
            row[(size_t)end * stride] = (u32)h1;
            if ((end > beg ? end : beg) == qlen) {                 // L200-203
                if (gscore <= h1) best_ie = i;
                if (gscore < h1) gscore = h1;
            }
        }
        const int m = key < 0 ? 0 : key >> 12, mj = key < 0 ? -1 : key & 0xfff;
        if (m == 0) break;
        if (m > best) {
            best = m; best_i = i; best_j = mj;
            int dd = mj - i; dd = dd < 0 ? -dd : dd;
            max_off = max_off > dd ? max_off : dd;
        } else if (P.zdrop > 0) {
            const int di = i - best_i, dj = mj - best_j;
            if (di > dj) { if (best - m - (di - dj) * e_del > P.zdrop) break; }
            else { if (best - m - (dj - di) * e_ins > P.zdrop) break; }
        }
        // ---- shrink the band to the live cells (L218-221): the usual first probes are loaded together ----
        {
            const u32 wb = beg < end ? row[(size_t)beg * stride] : 1u, we = row[(size_t)end * stride];   // (beg < end here: an empty band ended the loop above)
            int j = beg;
            if (wb == 0 && beg < end) { j++; while (j < end && row[(size_t)j * stride] == 0) j++; }
            beg = j;
            j = end;
            if (we == 0 && end >= beg) { j--; while (j >= beg && row[(size_t)j * stride] == 0) j--; }
            end = j + 2 < qlen ? j + 2 : qlen;
        }
    }
    out.score = best; out.tle = best_i + 1; out.gtle = best_ie + 1; out.qle = best_j + 1;
    out.gscore = gscore; out.max_off = max_off;
}

// One pair with the extension driver's band retries (reference src/bwamem.cpp:L2968-3000): up to `tries` widths
// w, 2w, ...; a result is final when the score equals `prev` (the region's score before this extension), when the best
// cell stayed within 3/4 of the band, or on the last try.  Returns the width that produced the final result.
BSW_HD int bsw_is_retry(const BswIsConst &P, int w, int tries, int prev, const uint8_t *query, int qlen, const BswTarget &target,
                        int tlen, int h0, u32 *row, uint8_t *qrow, const int stride, BswIsOut &out, u64 &cells)
{
    int wt = w;
    for (int t = 0;; t++) {
        wt = w << t;
        bsw_is_one(P, wt, query, qlen, target, tlen, h0, row, qrow, stride, out, cells);
        if (t + 1 >= tries || out.score == prev || out.max_off < (wt >> 1) + (wt >> 2)) break;
    }
    return wt;
}
