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

struct Seqs {               // one region's two sequences, already oriented as bwa_gen_cigar2 aligns them
    const uint8_t *codes;   // query segment: codes[q_off .. q_off + qlen)
    const gu64 *pac;        // 2-bit text, base p at bits 2*(p & 31) of word p >> 5
    int64_t q_off, rb;
    int qlen, tlen;
    bool rev;               // rb >= l_pac: both sequences are read backwards (src/bwa.cpp:L288-293)
    MEME_HD int q(int j) const { return codes[q_off + (rev ? qlen - 1 - j : j)]; }
    MEME_HD int t(int i) const { const gu64 p = (gu64)(rb + (rev ? tlen - 1 - i : i)); return (int)(pac[p >> 5] >> (2 * (p & 31))) & 3; }
};

struct EH { int h, e; };

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
MEME_HD void push_op(gu32 *ops, int &n, int op, int len)
{
    if (n == 0 || op != (int)(ops[n - 1] & 0xf)) ops[n++] = (gu32)len << 4 | (gu32)op;
    else ops[n - 1] += (gu32)len << 4;
}

// ksw_global2 with backtrack.  eh: qlen+1 cells; z: tlen * n_col bytes; on return ops (aliasing eh) holds the CIGAR in
// REVERSE order, *n_ops its length.
MEME_HD int global_align(const meme_cigar_opt_t &o, const Seqs &S, int w, EH *eh, uint8_t *z, int *n_ops, gu64 *cells)
{
    const int qlen = S.qlen, tlen = S.tlen;
    const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins;
    const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    int i, j;
    eh[0].h = 0; eh[0].e = CIG_MINUS_INF;
    for (j = 1; j <= qlen && j <= w; ++j) { eh[j].h = -(o.o_ins + o.e_ins * j); eh[j].e = CIG_MINUS_INF; }
    for (; j <= qlen; ++j) eh[j].h = eh[j].e = CIG_MINUS_INF;
    for (i = 0; i < tlen; ++i) {
        int f = CIG_MINUS_INF, h1, t;
        const int8_t *row = &o.mat[S.t(i) * 5];
        const int beg = i > w ? i - w : 0;
        const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
        h1 = beg == 0 ? -(o.o_del + o.e_del * (i + 1)) : CIG_MINUS_INF;
        uint8_t *zi = z + (int64_t)i * n_col;
        *cells += (gu64)(end - beg);
        for (j = beg; j < end; ++j) {
            EH *p = &eh[j];
            int h, m = p->h, e = p->e;
            uint8_t d;
            p->h = h1;
            m += row[S.q(j)];
            d = m >= e ? 0 : 1;
            h = m >= e ? m : e;
            d = h >= f ? d : 2;
            h = h >= f ? h : f;
            h1 = h;
            t = m - oe_del;
            e -= o.e_del;
            d |= e > t ? 1 << 2 : 0;
            e = e > t ? e : t;
            p->e = e;
            t = m - oe_ins;
            f -= o.e_ins;
            d |= f > t ? 2 << 4 : 0;
            f = f > t ? f : t;
            zi[j - beg] = d;
        }
        eh[end].h = h1; eh[end].e = CIG_MINUS_INF;
    }
    const int score = eh[qlen].h;
    // backtrack (L652-667); the row is dead from here on and becomes the operation list
    gu32 *ops = reinterpret_cast<gu32 *>(eh);
    int n = 0, which = 0, k;
    i = tlen - 1; k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
    while (i >= 0 && k >= 0) {
        which = z[(int64_t)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
        if (which == 0) { push_op(ops, n, 0, 1); --i; --k; }
        else if (which == 1) { push_op(ops, n, 2, 1); --i; }
        else { push_op(ops, n, 1, 1); --k; }
    }
    if (i >= 0) push_op(ops, n, 2, i + 1);
    if (k >= 0) push_op(ops, n, 1, k + 1);
    *n_ops = n;
    return score;
}

// one call of bwa_gen_cigar2 up to the CIGAR (L283-322); n_ops = -1: rejected (L285, L287)
MEME_HD int gen_cigar(const meme_cigar_opt_t &o, const Seqs &S, int w_, EH *eh, uint8_t *z, int *n_ops, gu64 *cells)
{
    gu32 *ops = reinterpret_cast<gu32 *>(eh);
    if (S.qlen == S.tlen && w_ == 0) {
        int score = 0;
        for (int i = 0; i < S.qlen; ++i) score += o.mat[S.t(i) * 5 + S.q(i)];
        ops[0] = (gu32)S.qlen << 4;
        *n_ops = 1;
        return score;
    }
    return global_align(o, S, band_of(o, S.qlen, S.tlen, w_), eh, z, n_ops, cells);
}

MEME_HD int put_uint(char *dst, int u)          // kputw for u >= 0: decimal digits, returns their number (dst may be null)
{
    char buf[12]; int l = 0;
    do { buf[l++] = (char)('0' + u % 10); u /= 10; } while (u > 0);
    if (dst) for (int x = 0; x < l; x++) dst[x] = buf[l - 1 - x];
    return l;
}

// NM and MD from the CIGAR (L323-353).  ops is the REVERSED list; md == null: only the length (with the final NUL) is wanted.
MEME_HD int nm_md(const Seqs &S, const gu32 *ops, int n_ops, char *md, int *nm)
{
    const char *int2base = S.rev ? "TGCAN" : "ACGTN";
    int x = 0, y = 0, u = 0, n_mm = 0, n_gap = 0, l = 0;
    for (int k = 0; k < n_ops; ++k) {
        const gu32 c = ops[n_ops - 1 - k];
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
// band is 4 * opt->w.  Leaves the final reversed CIGAR in eh.
MEME_HD int region_cigar(const meme_cigar_opt_t &o, const Seqs &S, int w2, int truesc, EH *eh, uint8_t *z, int *n_ops, gu64 *cells)
{
    int i = 0, score, last_sc = -(1 << 30);
    do {
        w2 = w2 < o.w << 2 ? w2 : o.w << 2;
        score = gen_cigar(o, S, w2, eh, z, n_ops, cells);
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

}  // namespace cig
