// meme_cigar_core.cuh -- the global alignment of one region, as one thread runs it (the
// compute of the SAM stage).
//
// What it computes is the band-doubling loop of mem_reg2aln (src/bwamem.cpp:L2339-2348) around bwa_gen_cigar2
// (src/bwa.cpp:L274-361): the reference segment is taken from the text, query and reference are reversed for regions on the
// reverse strand (so that indels are left-aligned on the forward strand), ksw_global2 (src/ksw.cpp:L560-670: banded
// Needleman-Wunsch with affine gaps and a backtrack matrix) produces score and CIGAR, and NM and the MD string follow from
// a walk over the CIGAR.  Differences in storage, not in results: the reference bases come from the HBM-resident 2-bit text
// by position (bns_get_seq unpacks .pac into a malloc'ed array), there is no query profile (the substitution score is read
// from the 5x5 matrix), the CIGAR is first written, reversed and run-length merged, into the DP row's memory -- which is
// dead once the matrix is filled and always large enough (at most 2*qlen+1 operations) -- and the MD string is written
// straight to its final place once its length is known.
//
// Every scratch array (DP row, backtrack matrix, oriented query) is addressed with a stride: on the device the 32 regions of
// a warp interleave their arrays element by element, so that lanes working on the same row and column touch one contiguous
// line; on the host (tests/host_sim) the stride is 1, or larger to exercise the interleaving.
//
// Plain C++ below MEME_HD, so that tests/host_sim compiles the same source with g++.
#pragma once
#include <stdint.h>
#include "../../include/meme_b200.h"

#ifndef MEME_HD
#ifdef __CUDACC__
#define MEME_HD __host__ __device__ __forceinline__
#else
#define MEME_HD inline
#endif
#endif

namespace cig {

typedef unsigned int gu32;
typedef unsigned long long gu64;

#define CIG_MINUS_INF (-0x40000000)
#define CIG_BATCH 8

struct Seqs {               // one region's two sequences, oriented as bwa_gen_cigar2 aligns them
    const uint8_t *qs;      // query segment, already oriented: base j at qs[j * stride]
    int64_t stride;
    const gu64 *pac;        // 2-bit text, base p at bits 2*(p & 31) of word p >> 5
    int64_t rb;
    int qlen, tlen;
    bool rev;               // rb >= l_pac: both sequences are read backwards (src/bwa.cpp:L288-293)
    MEME_HD int q(int j) const { return qs[(int64_t)j * stride]; }
    MEME_HD int t(int i) const { const gu64 p = (gu64)(rb + (rev ? tlen - 1 - i : i)); return (int)(pac[p >> 5] >> (2 * (p & 31))) & 3; }
};

struct EH { int h, e; };

struct Scr {                // a region's scratch: element k of each array sits at [k * stride]
    EH *eh; uint8_t *z; uint8_t *q; int64_t stride;
    MEME_HD EH &row(int j) const { return eh[(int64_t)j * stride]; }
    MEME_HD uint8_t &zz(int64_t k) const { return z[k * stride]; }
    // the DP row doubles as the (reversed) operation list once the matrix is filled: operation k in half k & 1 of cell k >> 1
    MEME_HD gu32 &op(int k) const { return reinterpret_cast<gu32 *>(&eh[(int64_t)(k >> 1) * stride])[k & 1]; }
};

// the oriented query into the scratch (src/bwa.cpp:L288-291 reverses it in place for reverse-strand regions)
MEME_HD void fill_query(const uint8_t *codes, int64_t q_off, int qlen, bool rev, const Scr &C)
{
    for (int j = 0; j < qlen; ++j) C.q[(int64_t)j * C.stride] = codes[q_off + (rev ? qlen - 1 - j : j)];
}

// the five scores of a reference base against A/C/G/T/N as bytes of one word
MEME_HD gu64 mat_row(const meme_cigar_opt_t &o, int t)
{
    gu64 r = 0;
    for (int k = 0; k < 5; k++) r |= (gu64)(uint8_t)o.mat[t * 5 + k] << (8 * k);
    return r;
}

// the band bwa_gen_cigar2 derives from the caller's w_ (src/bwa.cpp:L305-314)
MEME_HD int band_of(const meme_cigar_opt_t &o, int qlen, int tlen, int w_)
{
    int max_ins = (int)((double)(((qlen + 1) >> 1) * o.mat[0] - o.o_ins) / o.e_ins + 1.);
    int max_del = (int)((double)(((qlen + 1) >> 1) * o.mat[0] - o.o_del) / o.e_del + 1.);
    int max_gap = max_ins > max_del ? max_ins : max_del;
    max_gap = max_gap > 1 ? max_gap : 1;
    const int d = tlen > qlen ? tlen - qlen : qlen - tlen;
    int w = (max_gap + d + 1) >> 1;
    w = w < w_ ? w : w_;
    const int min_w = d + 3;
    return w > min_w ? w : min_w;
}

// append an operation to the (reversed) CIGAR, merging with the previous one (push_cigar, src/ksw.cpp:L545-558)
MEME_HD void push_op(const Scr &C, int &n, int op, int len)
{
    if (n == 0 || op != (int)(C.op(n - 1) & 0xf)) { C.op(n) = (gu32)len << 4 | (gu32)op; n++; }
    else C.op(n - 1) += (gu32)len << 4;
}

// ksw_global2 with backtrack.  Row: qlen+1 cells; matrix: tlen * n_col bytes; on return the row holds the CIGAR in REVERSE
// order, *n_ops its length.
MEME_HD int global_align(const meme_cigar_opt_t &o, const Seqs &S, int w, const Scr &C, int *n_ops, gu64 *cells)
{
    const int qlen = S.qlen, tlen = S.tlen;
    const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins;
    const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    const gu64 r0 = mat_row(o, 0), r1 = mat_row(o, 1), r2 = mat_row(o, 2), r3 = mat_row(o, 3);
    int i, j;
    C.row(0).h = 0; C.row(0).e = CIG_MINUS_INF;
    for (j = 1; j <= qlen && j <= w; ++j) { C.row(j).h = -(o.o_ins + o.e_ins * j); C.row(j).e = CIG_MINUS_INF; }
    for (; j <= qlen; ++j) C.row(j).h = C.row(j).e = CIG_MINUS_INF;
    for (i = 0; i < tlen; ++i) {
        int f = CIG_MINUS_INF, h1, t;
        const int tb = S.t(i);
        const gu64 rowbits = tb == 0 ? r0 : tb == 1 ? r1 : tb == 2 ? r2 : r3;
        const int beg = i > w ? i - w : 0;
        const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
        h1 = beg == 0 ? -(o.o_del + o.e_del * (i + 1)) : CIG_MINUS_INF;
        const int64_t zi = (int64_t)i * n_col - beg;
        *cells += (gu64)(end - beg);
        // the cells of a row depend on each other only through h1 / f, never through the row array (cell j is read, then
        // written, and no other cell of this row touches it): the loads of CIG_BATCH cells are issued together, so that a
        // thread has several memory requests in flight instead of one per cell (ncu: 33 warps stalled on long_scoreboard
        // per issued instruction with one)
        for (j = beg; j < end; j += CIG_BATCH) {
            EH c[CIG_BATCH];
            int qv[CIG_BATCH];
#pragma unroll
            for (int u = 0; u < CIG_BATCH; ++u)
                if (j + u < end) { c[u] = C.row(j + u); qv[u] = S.q(j + u); }
#pragma unroll
            for (int u = 0; u < CIG_BATCH; ++u) {
                if (j + u >= end) break;
                int h, m = c[u].h, e = c[u].e;
                uint8_t d;
                const int hprev = h1;
                m += (int)(int8_t)(rowbits >> (8 * qv[u]));
                d = m >= e ? 0 : 1;
                h = m >= e ? m : e;
                d = h >= f ? d : 2;
                h = h >= f ? h : f;
                h1 = h;
                t = m - oe_del;
                e -= o.e_del;
                d |= e > t ? 1 << 2 : 0;
                e = e > t ? e : t;
                EH &p = C.row(j + u);
                p.h = hprev; p.e = e;
                t = m - oe_ins;
                f -= o.e_ins;
                d |= f > t ? 2 << 4 : 0;
                f = f > t ? f : t;
                C.zz(zi + j + u) = d;
            }
        }
        C.row(end).h = h1; C.row(end).e = CIG_MINUS_INF;
    }
    const int score = C.row(qlen).h;
    // backtrack (L652-667); the row is dead from here on and becomes the operation list
    int n = 0, which = 0, k;
    i = tlen - 1; k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
    while (i >= 0 && k >= 0) {
        which = C.zz((int64_t)i * n_col + (k - (i > w ? i - w : 0))) >> (which << 1) & 3;
        if (which == 0) { push_op(C, n, 0, 1); --i; --k; }
        else if (which == 1) { push_op(C, n, 2, 1); --i; }
        else { push_op(C, n, 1, 1); --k; }
    }
    if (i >= 0) push_op(C, n, 2, i + 1);
    if (k >= 0) push_op(C, n, 1, k + 1);
    *n_ops = n;
    return score;
}

// one call of bwa_gen_cigar2 up to the CIGAR (L283-322)
MEME_HD int gen_cigar(const meme_cigar_opt_t &o, const Seqs &S, int w_, const Scr &C, int *n_ops, gu64 *cells)
{
    if (S.qlen == S.tlen && w_ == 0) {
        const gu64 r0 = mat_row(o, 0), r1 = mat_row(o, 1), r2 = mat_row(o, 2), r3 = mat_row(o, 3);
        int score = 0;
        for (int i = 0; i < S.qlen; ++i) {
            const int tb = S.t(i);
            const gu64 rowbits = tb == 0 ? r0 : tb == 1 ? r1 : tb == 2 ? r2 : r3;
            score += (int)(int8_t)(rowbits >> (8 * S.q(i)));
        }
        C.op(0) = (gu32)S.qlen << 4;
        *n_ops = 1;
        return score;
    }
    return global_align(o, S, band_of(o, S.qlen, S.tlen, w_), C, n_ops, cells);
}

MEME_HD int put_uint(char *dst, int u)          // kputw for u >= 0: decimal digits, returns their number (dst may be null)
{
    char buf[12]; int l = 0;
    do { buf[l++] = (char)('0' + u % 10); u /= 10; } while (u > 0);
    if (dst) for (int x = 0; x < l; x++) dst[x] = buf[l - 1 - x];
    return l;
}

// NM and MD from the CIGAR (L323-353).  The scratch row holds the REVERSED list; md == null: only the length (with the final
// NUL) is wanted.
MEME_HD int nm_md(const Seqs &S, const Scr &C, int n_ops, char *md, int *nm)
{
    const char *int2base = S.rev ? "TGCAN" : "ACGTN";
    int x = 0, y = 0, u = 0, n_mm = 0, n_gap = 0, l = 0;
    for (int k = 0; k < n_ops; ++k) {
        const gu32 c = C.op(n_ops - 1 - k);
        const int op = (int)(c & 0xf), len = (int)(c >> 4);
        if (op == 0) {
            for (int i = 0; i < len; ++i) {
                const int r = S.t(y + i);
                if (S.q(x + i) != r) {
                    l += put_uint(md ? md + l : nullptr, u);
                    if (md) md[l] = int2base[r];
                    ++l; ++n_mm; u = 0;
                } else ++u;
            }
            x += len; y += len;
        } else if (op == 2) {
            if (k > 0 && k < n_ops - 1) {
                l += put_uint(md ? md + l : nullptr, u);
                if (md) md[l] = '^';
                ++l;
                for (int i = 0; i < len; ++i) { if (md) md[l] = int2base[S.t(y + i)]; ++l; }
                u = 0; n_gap += len;
            }
            y += len;
        } else if (op == 1) { x += len; n_gap += len; }
    }
    l += put_uint(md ? md + l : nullptr, u);
    if (md) md[l] = 0;
    ++l;
    *nm = n_mm + n_gap;
    return l;
}

// the loop of mem_reg2aln (L2339-2348): widen the band until the score stops changing, reaches the local score, or the
// band is 4 * opt->w.  Leaves the final reversed CIGAR in the scratch row.
MEME_HD int region_cigar(const meme_cigar_opt_t &o, const Seqs &S, int w2, int truesc, const Scr &C, int *n_ops, gu64 *cells)
{
    int i = 0, score, last_sc = -(1 << 30);
    do {
        w2 = w2 < o.w << 2 ? w2 : o.w << 2;
        score = gen_cigar(o, S, w2, C, n_ops, cells);
        if (score == last_sc || w2 == o.w << 2) break;
        last_sc = score;
        w2 <<= 1;
    } while (++i < 3 && score < truesc - o.a);
    return score;
}

// bytes of backtrack matrix the loop above can need for a task (host and device agree through this one function)
MEME_HD int64_t z_bytes(const meme_cigar_opt_t &o, int qlen, int tlen, int w2)
{
    int64_t need = 0;
    for (int i = 0; i < 3; i++) {
        w2 = w2 < o.w << 2 ? w2 : o.w << 2;
        if (!(qlen == tlen && w2 == 0)) {
            const int w = band_of(o, qlen, tlen, w2);
            const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
            const int64_t b = (int64_t)n_col * tlen;
            need = b > need ? b : need;
        }
        if (w2 == o.w << 2) break;
        w2 <<= 1;
    }
    return need;
}

// band of the first try (0: no matrix), the quantity regions are grouped by before they share a warp
MEME_HD int first_band(const meme_cigar_opt_t &o, int qlen, int tlen, int w2)
{
    w2 = w2 < o.w << 2 ? w2 : o.w << 2;
    return qlen == tlen && w2 == 0 ? 0 : band_of(o, qlen, tlen, w2);
}

}  // namespace cig
