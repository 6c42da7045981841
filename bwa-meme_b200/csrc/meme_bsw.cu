// meme_bsw.cu -- kernel B: banded affine-gap Smith-Waterman seed extension on sm_100a.
//
// Two kernels: the inter-sequence kernel k_bsw_is (meme_bsw_is.cuh: one thread per pair, row in shared memory) takes
// every pair whose scores fit 13 bits and whose query is at most 320 bases -- all of them on the 150/250 bp read path;
// the warp-per-pair int32 kernel below (k_bsw<C>) and the one-thread generic kernel remain as the overflow classes.
//
// Computes, for every SeqPair, exactly what BandedPairWiseSW::scalarBandedSWA returns
// (reference src/bandedSWA.cpp:L116-237) -- the semantic definition that the reference's
// SSE/AVX2/AVX-512 inter-sequence kernels (getScores8/16, L412-4870) vectorise.
//
// B200 design: the CPU code puts one PAIR per SIMD lane and walks the DP row by row.
// Here one WARP owns a pair and the 32 lanes split the ROW: lane k keeps columns [kC, (k+1)C) of
// the H/E score row in registers (C = ceil((qlen+1)/32), a template parameter), so the row never
// touches memory.  Within a row
//     M(j)   = H(i-1,j-1) ? H(i-1,j-1) + s(i,j) : 0          depends on the previous row only
//     E(i,j) , and F's source term max(M(j)-oe_ins, 0)      likewise
// so the only serial dependency is F(i,j+1) = max(F(i,j)-e_ins, max(M(j)-oe_ins,0)), a max-plus
// prefix scan: each lane scans its C columns, then one 5-step __shfl_up_sync prefix-max over
// G = F_out + e_ins*column stitches the lanes together.  The row maximum (value, LAST column that
// attains it) is one packed shuffle reduction; the adaptive band [beg,end) of the reference
// (L218-221) comes from two ballots.  All per-row control (z-drop, band, gscore) is warp-uniform.
// Slots outside the current band keep their old contents, exactly like the reference's eh[] array.
#include <vector>
#include <thread>
#include <algorithm>
#include <climits>
#include <cub/device/device_radix_sort.cuh>
#include "meme_common.cuh"
#include "meme_bsw_is.cuh"

#define BSW_WARPS 4
#define BSW_NCLASS 10
#define BSW_IS_NCLASS 6
#define BSW_IS_THREADS 128
#define FULL 0xffffffffu
#define NEG_BIG (-(1 << 28))

struct meme_bsw {
    int device;
    cudaStream_t own_stream, stream;
    meme_seqpair_t *d_pairs; size_t cap_pairs;
    uint8_t *d_ref; size_t cap_ref;
    uint8_t *d_qer; size_t cap_qer;
    int32_t *d_scratch; size_t cap_scratch;     // H/E rows of the generic (very long query) path
    int64_t *d_scratch_off; size_t cap_scratch_off;
    u32 *d_order; size_t cap_order;
    int64_t cls_off[BSW_NCLASS + 1];
    // inter-sequence path: device-sorted order of the pairs it takes, ranges per row-size class
    u32 *d_keys, *d_keys2, *d_idx, *d_idx2; size_t cap_sort;
    void *d_sort_tmp; size_t cap_sort_tmp;
    int64_t is_off[BSW_IS_NCLASS + 1]; int64_t n_is, n_ovf;
    int32_t *d_wused; size_t cap_wused;
    int match_a; int attr_set;
    // reference by position: the 2-bit text of an index on the same device, and whether the uploaded batch uses it
    const u64 *text_pac; u64 text_len; int text_mode;
    int64_t *d_ovf_off; size_t cap_ovf_off;
    u64 *d_ctrl; u64 *h_ctrl;
    int64_t n_pairs, n_ref, n_qer;
    u64 launches;
};

struct BswConst {
    int o_del, e_del, o_ins, e_ins, zdrop, end_bonus, a, b, w;
};

__device__ __forceinline__ int score_of(int tc, int qc, int a, int b)
{   // bwa_fill_scmat (reference src/bwa.cpp:L262-271): N against anything is -1
    return (tc > 3 || qc > 3) ? -1 : (tc == qc ? a : -b);
}

__device__ __forceinline__ int clip_band(int w, int qlen, const BswConst &P)
{   // reference L147-156
    int mx = P.a > 0 ? P.a : 0;                 // max of the score matrix (mismatch/N entries are negative)
    int max_ins = (int)((double)(qlen * mx + P.end_bonus - P.o_ins) / P.e_ins + 1.);
    if (max_ins < 1) max_ins = 1;
    if (w > max_ins) w = max_ins;
    int max_del = (int)((double)(qlen * mx + P.end_bonus - P.o_del) / P.e_del + 1.);
    if (max_del < 1) max_del = 1;
    if (w > max_del) w = max_del;
    return w;
}

template <int C>
__device__ void bsw_warp(const BswConst &P, const uint8_t *__restrict__ query, int qlen,
                         const uint8_t *__restrict__ target, int tlen, int h0, meme_seqpair_t *out, u64 &cells)
{
    const int lane = threadIdx.x & 31;
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    const int e_ins = P.e_ins, e_del = P.e_del;
    const int j0 = lane * C;                    // first column (slot) of this lane
    int H[C], E[C], qc[C];
    // row "-1" (reference L143-145): H[0]=h0, H[1]=max(h0-oe_ins,0), then -e_ins while the previous is > e_ins
#pragma unroll
    for (int c = 0; c < C; c++) {
        int j = j0 + c;
        int v = 0;
        if (j == 0) v = h0;
        else if (j <= qlen) {
            // H[j] = h0 - oe_ins - (j-1)*e_ins as long as every earlier H[k] (k>=1) was > e_ins
            int first = h0 > oe_ins ? h0 - oe_ins : 0;
            if (j == 1) v = first;
            else {
                int prev = first - (j - 2) * e_ins;     // H[j-1] if the chain is still alive
                v = (first > 0 && prev > e_ins) ? prev - e_ins : 0;
            }
        }
        H[c] = v;
        E[c] = 0;
        qc[c] = j < qlen ? query[j] : 4;
    }
    const int w = clip_band(P.w, qlen, P);
    int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen;
    int tchunk = 0;
    for (int i = 0; i < tlen; i++) {
        if ((i & 31) == 0) tchunk = (i + lane < tlen) ? target[i + lane] : 4;
        const int tc = __shfl_sync(FULL, tchunk, i & 31);
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        int h1_init = 0;
        if (beg == 0) { h1_init = h0 - (P.o_del + e_del * (i + 1)); if (h1_init < 0) h1_init = 0; }
        int h1_end;                               // H(i, end-1) as the reference leaves it in h1
        int rowkey = -1;                          // (h << 12) | j of the row maximum, last j wins
        if (beg < end) {
            // ---- pass 1: M and the lane-local F chain ----
            int Mv[C], fv[C];
            int lf = NEG_BIG;
#pragma unroll
            for (int c = 0; c < C; c++) {
                int j = j0 + c;
                bool in = j >= beg && j < end;
                int d = H[c];
                int M = d ? d + score_of(tc, qc[c], P.a, P.b) : 0;
                Mv[c] = M;
                fv[c] = lf;
                if (in) lf = __viaddmax_s32(lf, -e_ins, __viaddmax_s32(M, -oe_ins, 0));     // DPX: max(lf - e_ins, max(M - oe_ins, 0))
            }
            // ---- stitch F across lanes: exclusive prefix max of G = F_out + e_ins * (next column) ----
            int endcol = min(end, j0 + C);
            bool active = j0 < end && j0 + C > beg;
            int G = active ? lf + e_ins * endcol : NEG_BIG;
            int pg = G;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(FULL, pg, o);
                if (lane >= o) pg = max(pg, v);
            }
            pg = __shfl_up_sync(FULL, pg, 1);
            if (lane == 0) pg = NEG_BIG;
            // ---- pass 2: H, E, row maximum ----
            int hprev = 0;                         // H(i, j-1) within the lane
            int hlast = 0;
            int newH[C];
            int lkey = -1;
#pragma unroll
            for (int c = 0; c < C; c++) {
                int j = j0 + c;
                bool in = j >= beg && j < end;
                int f;
                if (j == beg) f = 0;
                else { f = max(fv[c], pg - e_ins * j); }
                int e = E[c], M = Mv[c];
                int h = __vimax3_s32_relu(M, e, f);   // DPX; relu: f may be a large negative placeholder outside the chain
                newH[c] = hprev;                   // H(i, j-1); fixed up below for c == 0 and j == beg
                if (in) {
                    E[c] = __viaddmax_s32(e, -e_del, __viaddmax_s32(M, -oe_del, 0));
                    int key = (h << 12) | j;
                    lkey = max(lkey, key);
                    hprev = h;
                    hlast = h;
                } else {
                    hprev = 0;
                }
            }
            // the value to the left of a lane's first column comes from the previous lane's last column
            int left = __shfl_up_sync(FULL, hlast, 1);
            // which lanes own the last computed column of the row / columns beg and end
#pragma unroll
            for (int c = 0; c < C; c++) {
                int j = j0 + c;
                if (j >= beg && j <= end) {
                    int v = (c == 0) ? left : newH[c];
                    if (j == beg) v = h1_init;
                    H[c] = v;
                    if (j == end) E[c] = 0;
                }
            }
            // h1 at loop exit = H(i, end-1): owned by lane (end-1)/C
            {
                int owner = (end - 1) / C;
                h1_end = __shfl_sync(FULL, hlast, owner);
            }
            rowkey = __reduce_max_sync(FULL, lkey);
            cells += (u64)(end - beg);
        } else {
            // empty band: the reference still writes eh[end] = {h1, 0} (L199)
            h1_end = h1_init;
#pragma unroll
            for (int c = 0; c < C; c++)
                if (j0 + c == end) { H[c] = h1_init; E[c] = 0; }
        }
        const int jexit = beg < end ? end : beg;
        if (jexit == qlen) {                      // L200-203
            if (gscore <= h1_end) best_ie = i;
            if (gscore < h1_end) gscore = h1_end;
        }
        const int m = rowkey < 0 ? 0 : rowkey >> 12;
        const int mj = rowkey < 0 ? -1 : rowkey & 0xfff;
        if (m == 0) break;
        if (m > best) {
            best = m; best_i = i; best_j = mj;
            int d = mj - i; d = d < 0 ? -d : d;
            max_off = max(max_off, d);
        } else if (P.zdrop > 0) {
            int di = i - best_i, dj = mj - best_j;
            if (di > dj) { if (best - m - (di - dj) * e_del > P.zdrop) break; }
            else { if (best - m - (dj - di) * e_ins > P.zdrop) break; }
        }
        // ---- shrink the band to the live cells (L218-221) ----
        int lfirst = 1 << 20, llast = -1;
#pragma unroll
        for (int c = 0; c < C; c++) {
            int j = j0 + c;
            if (j >= beg && j <= end && (H[c] | E[c]) != 0) {
                if (j < end) lfirst = min(lfirst, j);
                llast = j;
            }
        }
        lfirst = __reduce_min_sync(FULL, lfirst);
        llast = __reduce_max_sync(FULL, llast);
        int nbeg = lfirst < end ? lfirst : end;
        int jl = llast >= nbeg ? llast : nbeg - 1;
        beg = nbeg;
        end = jl + 2 < qlen ? jl + 2 : qlen;
    }
    if (lane == 0) {
        out->score = best; out->tle = best_i + 1; out->gtle = best_ie + 1; out->qle = best_j + 1;
        out->gscore = gscore; out->max_off = max_off;
    }
}

// Generic path for queries too long for the register kernel: one thread per pair, score row in
// global scratch.  Same recurrence, written for clarity, not speed.
__device__ void bsw_thread(const BswConst &P, const uint8_t *__restrict__ query, int qlen,
                           const uint8_t *__restrict__ target, int tlen, int h0, int32_t *__restrict__ H,
                           int32_t *__restrict__ E, meme_seqpair_t *out, u64 &cells)
{
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    for (int j = 0; j <= qlen + 1; j++) { H[j] = 0; E[j] = 0; }
    H[0] = h0;
    H[1] = h0 > oe_ins ? h0 - oe_ins : 0;
    for (int j = 2; j <= qlen && H[j - 1] > P.e_ins; j++) H[j] = H[j - 1] - P.e_ins;
    const int w = clip_band(P.w, qlen, P);
    int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0, beg = 0, end = qlen;
    for (int i = 0; i < tlen; i++) {
        int tc = target[i], f = 0, m = 0, mj = -1, h1;
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        if (beg == 0) { h1 = h0 - (P.o_del + P.e_del * (i + 1)); if (h1 < 0) h1 = 0; } else h1 = 0;
        int j;
        for (j = beg; j < end; j++) {
            int d = H[j], e = E[j];
            H[j] = h1;
            int M = d ? d + score_of(tc, query[j], P.a, P.b) : 0;
            int h = max(max(M, e), f);
            h1 = h;
            if (h >= m) { m = h; mj = j; }
            int t = M - oe_del; t = t > 0 ? t : 0; e -= P.e_del; E[j] = e > t ? e : t;
            t = M - oe_ins; t = t > 0 ? t : 0; f -= P.e_ins; f = f > t ? f : t;
        }
        cells += (u64)(end > beg ? end - beg : 0);
        H[end] = h1; E[end] = 0;
        if (j == qlen) { if (gscore <= h1) best_ie = i; if (gscore < h1) gscore = h1; }
        if (m == 0) break;
        if (m > best) { best = m; best_i = i; best_j = mj; int d = mj - i; d = d < 0 ? -d : d; max_off = max(max_off, d); }
        else if (P.zdrop > 0) {
            int di = i - best_i, dj = mj - best_j;
            if (di > dj) { if (best - m - (di - dj) * P.e_del > P.zdrop) break; }
            else { if (best - m - (dj - di) * P.e_ins > P.zdrop) break; }
        }
        for (j = beg; j < end && H[j] == 0 && E[j] == 0; j++);
        beg = j;
        for (j = end; j >= beg && H[j] == 0 && E[j] == 0; j--);
        end = j + 2 < qlen ? j + 2 : qlen;
    }
    out->score = best; out->tle = best_i + 1; out->gtle = best_ie + 1; out->qle = best_j + 1;
    out->gscore = gscore; out->max_off = max_off;
}

#define BSW_MAX_REG_QLEN (32 * 12 - 1)

// the extension driver's acceptance test (reference src/bwamem.cpp:L2985-2987): final when the score did not move,
// when the best cell stayed within 3/4 of the band, or on the last try
__device__ __forceinline__ bool bsw_final(int score, int max_off, int prev, int w, int t, int tries)
{
    return t + 1 >= tries || score == prev || max_off < (w >> 1) + (w >> 2);
}

// Overflow class 1: one warp per pair, int32 cells, one instantiation per column count C (registers follow the class).
// `order` lists the pairs of this class.
template <int C>
__global__ void __launch_bounds__(BSW_WARPS * 32)
k_bsw(BswConst P, int tries, int prev_is_h0, meme_seqpair_t *__restrict__ pairs, const u32 *__restrict__ order, int64_t n_work,
      const uint8_t *__restrict__ ref, const uint8_t *__restrict__ qer, int32_t *__restrict__ w_used, u64 *__restrict__ ctrl)
{
    const int lane = threadIdx.x & 31;
    u64 cells = 0;
    for (;;) {
        u64 r = 0;
        if (lane == 0) r = atomicAdd((unsigned long long *)&ctrl[0], 1ull);
        r = __shfl_sync(FULL, r, 0);
        if (r >= (u64)n_work) break;
        meme_seqpair_t *p = pairs + order[r];
        const int h0 = p->h0, prev = prev_is_h0 ? h0 : -1;
        BswConst Q = P;
        for (int t = 0;; t++) {
            Q.w = P.w << t;
            bsw_warp<C>(Q, qer + p->idq, p->len2, ref + p->idr, p->len1, h0, p, cells);
            int fin = 0;
            if (lane == 0) fin = bsw_final(p->score, p->max_off, prev, Q.w, t, tries);
            if (__shfl_sync(FULL, fin, 0)) break;
        }
        if (lane == 0 && w_used) w_used[order[r]] = Q.w;
    }
    if (lane == 0 && cells) atomicAdd((unsigned long long *)&ctrl[1], (unsigned long long)cells);
}

// classes of the warp kernel: C = 1,2,3,4,5,6,8,10,12 and "long" (generic kernel)
static inline int bsw_class_of(int qlen)
{
    if (qlen > BSW_MAX_REG_QLEN) return 9;
    int C = (qlen + 1 + 31) / 32;
    if (C <= 1) return 0;
    if (C <= 6) return C - 1;
    if (C <= 8) return 6;
    if (C <= 10) return 7;
    return 8;
}

// Overflow class 2: queries beyond the register kernel, one thread per pair, rows in global scratch.
__global__ void k_bsw_long(BswConst P, int tries, int prev_is_h0, meme_seqpair_t *__restrict__ pairs, const u32 *__restrict__ order,
                           int64_t n_work, const uint8_t *__restrict__ ref, const uint8_t *__restrict__ qer,
                           const int64_t *__restrict__ off, int32_t *__restrict__ scratch, int32_t *__restrict__ w_used,
                           u64 *__restrict__ ctrl)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_work) return;
    meme_seqpair_t *p = pairs + order[i];
    u64 cells = 0;
    int32_t *H = scratch + off[i], *E = H + p->len2 + 2;
    const int h0 = p->h0, prev = prev_is_h0 ? h0 : -1;
    BswConst Q = P;
    for (int t = 0;; t++) {
        Q.w = P.w << t;
        bsw_thread(Q, qer + p->idq, p->len2, ref + p->idr, p->len1, h0, H, E, p, cells);
        if (bsw_final(p->score, p->max_off, prev, Q.w, t, tries)) break;
    }
    if (w_used) w_used[order[i]] = Q.w;
    atomicAdd((unsigned long long *)&ctrl[1], (unsigned long long)cells);
}

// ---------------------------------------------------------------------------------------------
// the inter-sequence kernel (meme_bsw_is.cuh): one thread per pair, rows [column][thread] in shared memory
// ---------------------------------------------------------------------------------------------
// row words per thread of the five classes (query length + 2 <= words)
static const int BSW_IS_WORDS[BSW_IS_NCLASS] = {34, 66, 98, 130, 162, 322};
static inline int bsw_is_class_of(int qlen)
{
    for (int c = 0; c < BSW_IS_NCLASS; c++) if (qlen + 2 <= BSW_IS_WORDS[c]) return c;
    return -1;
}

__global__ void __launch_bounds__(BSW_IS_THREADS)
k_bsw_is(BswIsConst P, int words, int w, int tries, int prev_is_h0, meme_seqpair_t *__restrict__ pairs, const u32 *__restrict__ order,
         int64_t n_work, const uint8_t *__restrict__ ref, const u64 *__restrict__ text_pac, const uint8_t *__restrict__ qer,
         int32_t *__restrict__ w_used, u64 *__restrict__ ctrl)
{
    extern __shared__ u32 bsw_rows[];
    u32 *row = bsw_rows + threadIdx.x;
    uint8_t *qrow = reinterpret_cast<uint8_t *>(bsw_rows + (size_t)words * BSW_IS_THREADS) + threadIdx.x;
    u64 cells = 0;
    // static striding over the length-sorted order: the 32 pairs of a warp are neighbours in the sort
    for (int64_t k = (int64_t)blockIdx.x * BSW_IS_THREADS + threadIdx.x; k < n_work; k += (int64_t)gridDim.x * BSW_IS_THREADS) {
        const u32 pi = order[k];
        meme_seqpair_t *p = pairs + pi;
        const int h0 = p->h0;
        BswIsOut o;
        BswTarget T;
        if (text_pac) {                         // (id, idr) = position in the 2-bit text, sign of id = direction
            T.bytes = nullptr; T.pac = text_pac; T.dir = p->id < 0 ? -1 : +1;
            T.pos = ((u64)((u32)p->id & 0x7fffffffu) << 32) | (u32)p->idr;
        } else { T.bytes = ref + p->idr; T.pac = nullptr; T.pos = 0; T.dir = 1; }
        const int wu = bsw_is_retry(P, w, tries, prev_is_h0 ? h0 : -1, qer + p->idq, p->len2, T, p->len1, h0, row,
                                    qrow, BSW_IS_THREADS, o, cells);
        p->score = o.score; p->tle = o.tle; p->gtle = o.gtle; p->qle = o.qle; p->gscore = o.gscore; p->max_off = o.max_off;
        if (w_used) w_used[pi] = wu;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cells += __shfl_xor_sync(FULL, cells, o);
    if ((threadIdx.x & 31) == 0 && cells) atomicAdd((unsigned long long *)&ctrl[1], (unsigned long long)cells);
}

// sort keys: pairs of the inter-sequence kernel by (query length, reference length), longest first; pairs of the
// overflow classes get the largest key and are skipped by the inter-sequence launches
__global__ void k_bsw_keys(const meme_seqpair_t *__restrict__ pairs, int64_t n, int match_a, u32 *__restrict__ keys, u32 *__restrict__ idx)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const meme_seqpair_t p = pairs[i];
    const bool ovf = p.len2 > BSW_IS_MAX_QLEN || (long long)p.h0 + (long long)p.len2 * match_a >= BSW_IS_MAX_SCORE || p.h0 < 0;
    u32 l1 = p.len1 > 0xffff ? 0xffffu : (u32)p.len1;
    keys[i] = ovf ? 0xffffffffu : (((u32)p.len2 << 16) | l1);
    idx[i] = (u32)i;
}

// ---------------------------------------------------------------------------------------------
// host API
// ---------------------------------------------------------------------------------------------
extern "C" int meme_bsw_create(int device, meme_bsw_t **out)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        meme_set_error("no CUDA device visible: meme_b200 has no CPU path");
        return MEME_E_NOGPU;
    }
    if (!out) return MEME_E_ARG;
    CUDA_TRY(cudaSetDevice(device));
    meme_bsw *b = (meme_bsw *)calloc(1, sizeof(meme_bsw));
    if (!b) return MEME_E_NOMEM;
    b->device = device;
    CUDA_TRY(cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking));
    b->stream = b->own_stream;
    CUDA_TRY(cudaMalloc((void **)&b->d_ctrl, 8 * 8));
    CUDA_TRY(cudaMallocHost((void **)&b->h_ctrl, 8 * 8));
    b->match_a = 1;
    *out = b;
    return MEME_OK;
}

extern "C" void meme_bsw_free(meme_bsw_t *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    cudaStreamSynchronize(b->stream);
    void *dev[] = {b->d_pairs, b->d_ref, b->d_qer, b->d_scratch, b->d_scratch_off, b->d_ctrl, b->d_order, b->d_keys, b->d_keys2,
                   b->d_idx, b->d_idx2, b->d_sort_tmp, b->d_wused, b->d_ovf_off};
    for (void *p : dev) if (p) cudaFree(p);
    if (b->h_ctrl) cudaFreeHost(b->h_ctrl);
    cudaStreamDestroy(b->own_stream);
    free(b);
}

extern "C" int meme_bsw_set_stream(meme_bsw_t *b, void *cuda_stream)
{
    if (!b) return MEME_E_ARG;
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    b->stream = cuda_stream ? (cudaStream_t)cuda_stream : b->own_stream;
    return MEME_OK;
}

// The match score decides which pairs fit the 13-bit cells of the inter-sequence kernel; it is part of the options of
// a run, but the classes are built at upload: meme_bsw_set_match tells the handle before the upload (default 1).
extern "C" int meme_bsw_set_match(meme_bsw_t *b, int32_t a)
{
    if (!b || a < 0) return MEME_E_ARG;
    b->match_a = a;
    return MEME_OK;
}

// reference by position, overflow classes: their (few) pairs get their reference bytes written out, so that the
// warp-per-pair and generic kernels run unchanged
__global__ void k_bsw_materialise(meme_seqpair_t *__restrict__ pairs, const u32 *__restrict__ order, const int64_t *__restrict__ off,
                                  int64_t n_work, const u64 *__restrict__ pac, uint8_t *__restrict__ ref)
{
    int64_t k = blockIdx.x;
    if (k >= n_work) return;
    meme_seqpair_t *p = pairs + order[k];
    const int dir = p->id < 0 ? -1 : +1;
    const u64 pos = ((u64)((u32)p->id & 0x7fffffffu) << 32) | (u32)p->idr;
    uint8_t *dst = ref + off[k];
    for (int i = threadIdx.x; i < p->len1; i += blockDim.x) {
        const u64 q = dir > 0 ? pos + (u64)i : pos - (u64)i;
        dst[i] = (uint8_t)((pac[q >> 5] >> (2 * (q & 31))) & 3);
    }
    __syncthreads();
    if (threadIdx.x == 0) { p->idr = (int32_t)off[k]; p->id = 0; }
}

// Lets the handle read pair references from the 2-bit text of `ix` (same device) instead of uploaded bytes: pairs of a
// batch uploaded with ref == NULL carry a text position in (id, idr) -- idr = low 32 bits, id = bits 32..62, id < 0 for a
// pair that walks the text backwards (left extensions).  NULL detaches.
extern "C" int meme_bsw_attach_text(meme_bsw_t *b, const meme_index_t *ix)
{
    if (!b) return MEME_E_ARG;
    if (ix && ix->device != b->device) { meme_set_error("index and banded-SW handle live on different devices"); return MEME_E_ARG; }
    b->text_pac = ix ? ix->v.pac : nullptr;
    b->text_len = ix ? ix->v.sa_num : 0;
    return MEME_OK;
}

__global__ void k_bsw_rebase(meme_seqpair_t *__restrict__ pairs, int64_t n, int32_t ref_base, int32_t qer_base)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { pairs[i].idr += ref_base; pairs[i].idq += qer_base; }
}

// Upload of a batch given as `nseg` segments, each with its own pair array and sequence buffers (pair offsets are
// relative to the segment's buffers): what a multi-threaded caller assembles without a gather copy.  The segments are
// laid out back to back on the device and the pair offsets rebased there.
extern "C" int meme_bsw_upload_segments(meme_bsw_t *b, int32_t nseg, const meme_seqpair_t *const *pairs, const int64_t *n_pairs,
                                        const uint8_t *const *ref, const int64_t *n_ref, const uint8_t *const *qer,
                                        const int64_t *n_qer)
{
    if (!b || nseg < 0 || (nseg && (!pairs || !n_pairs || !ref || !n_ref || !qer || !n_qer))) { meme_set_error("bad argument"); return MEME_E_ARG; }
    int64_t tot_p = 0, tot_r = 0, tot_q = 0;
    // reference by position: no reference bytes at all, and a text is attached
    bool text_mode = b->text_pac != nullptr;
    for (int s = 0; s < nseg && text_mode; s++) if (ref[s] || n_ref[s]) text_mode = false;
    for (int s = 0; s < nseg; s++) {
        if (n_pairs[s] < 0 || n_ref[s] < 0 || n_qer[s] < 0 || (n_pairs[s] && (!pairs[s] || (!text_mode && !ref[s]) || !qer[s]))) {
            meme_set_error("bad segment %d", s);
            return MEME_E_ARG;
        }
        tot_p += n_pairs[s]; tot_r += n_ref[s]; tot_q += n_qer[s];
    }
    if (tot_p >= (1ll << 31) || tot_r >= (1ll << 31) || tot_q >= (1ll << 31)) {
        meme_set_error("batch too large for 32-bit pair offsets (%lld pairs, %lld + %lld sequence bytes)", (long long)tot_p, (long long)tot_r, (long long)tot_q);
        return MEME_E_ARG;
    }
    // validation, class counts of the inter-sequence kernel, and the (few) pairs of the overflow classes: one host thread
    // per segment (the segments are the arenas of the caller's threads, a few hundred thousand pairs each)
    int64_t is_cnt[BSW_IS_NCLASS] = {0};
    std::vector<u32> ovf;
    std::vector<int32_t> ovf_len2, ovf_len1;
    {
        struct SegScan { int64_t cnt[BSW_IS_NCLASS] = {0}; std::vector<u32> ovf; std::vector<int32_t> l2, l1; int64_t bad = -1; };
        std::vector<SegScan> scan((size_t)nseg);
        std::vector<int64_t> seg_base((size_t)nseg + 1, 0);
        for (int s = 0; s < nseg; s++) seg_base[(size_t)s + 1] = seg_base[(size_t)s] + n_pairs[s];
        const int match_a = b->match_a;
        const u64 text_len = b->text_len;
        auto scan_seg = [&](int s) {
            SegScan &S = scan[(size_t)s];
            const meme_seqpair_t *ps = pairs[s];
            const int64_t base = seg_base[(size_t)s];
            for (int64_t i = 0; i < n_pairs[s]; i++) {
                const meme_seqpair_t &p = ps[i];
                bool bad = p.len1 < 0 || p.len2 < 0 || p.idq < 0 || (int64_t)p.idq + p.len2 > n_qer[s];
                if (!bad && text_mode) {
                    const u64 pos = ((u64)((u32)p.id & 0x7fffffffu) << 32) | (u32)p.idr;
                    bad = p.len1 > 0 && (p.id < 0 ? (pos >= text_len || pos + 1 < (u64)p.len1) : pos + (u64)p.len1 > text_len);
                } else if (!bad) bad = p.idr < 0 || (int64_t)p.idr + p.len1 > n_ref[s];
                if (bad) { S.bad = i; return; }
                const bool o = p.len2 > BSW_IS_MAX_QLEN || (long long)p.h0 + (long long)p.len2 * match_a >= BSW_IS_MAX_SCORE || p.h0 < 0;
                if (o) { S.ovf.push_back((u32)(base + i)); S.l2.push_back(p.len2); S.l1.push_back(p.len1); }
                else S.cnt[bsw_is_class_of(p.len2)]++;
            }
        };
        if (nseg > 1 && tot_p > 100000) {
            std::vector<std::thread> th;
            for (int s = 0; s < nseg; s++) th.emplace_back(scan_seg, s);
            for (auto &t : th) t.join();
        } else {
            for (int s = 0; s < nseg; s++) scan_seg(s);
        }
        for (int s = 0; s < nseg; s++) {
            const SegScan &S = scan[(size_t)s];
            if (S.bad >= 0) {
                meme_set_error("pair %lld of segment %d points outside the sequence buffers", (long long)S.bad, s);
                return MEME_E_ARG;
            }
            for (int c = 0; c < BSW_IS_NCLASS; c++) is_cnt[c] += S.cnt[c];
            ovf.insert(ovf.end(), S.ovf.begin(), S.ovf.end());
            ovf_len2.insert(ovf_len2.end(), S.l2.begin(), S.l2.end());
            ovf_len1.insert(ovf_len1.end(), S.l1.begin(), S.l1.end());
        }
    }
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t st = b->stream;
    int rc;
    const int64_t n_all = tot_p;
    if ((rc = meme_grow_device((void **)&b->d_pairs, &b->cap_pairs, (size_t)n_all + 1, sizeof(meme_seqpair_t), st))) return rc;
    if ((rc = meme_grow_device((void **)&b->d_ref, &b->cap_ref, (size_t)tot_r + 64, 1, st))) return rc;
    if ((rc = meme_grow_device((void **)&b->d_qer, &b->cap_qer, (size_t)tot_q + 64, 1, st))) return rc;
    {
        int64_t pb = 0, rb = 0, qb = 0;
        for (int s = 0; s < nseg; s++) {
            if (n_pairs[s]) {
                CUDA_TRY(cudaMemcpyAsync(b->d_pairs + pb, pairs[s], (size_t)n_pairs[s] * sizeof(meme_seqpair_t), cudaMemcpyHostToDevice, st));
                if (n_ref[s]) CUDA_TRY(cudaMemcpyAsync(b->d_ref + rb, ref[s], (size_t)n_ref[s], cudaMemcpyHostToDevice, st));
                if (n_qer[s]) CUDA_TRY(cudaMemcpyAsync(b->d_qer + qb, qer[s], (size_t)n_qer[s], cudaMemcpyHostToDevice, st));
                if ((rb && !text_mode) || qb)
                    k_bsw_rebase<<<(unsigned)((n_pairs[s] + 255) / 256), 256, 0, st>>>(b->d_pairs + pb, n_pairs[s], text_mode ? 0 : (int32_t)rb, (int32_t)qb);
            }
            pb += n_pairs[s]; rb += n_ref[s]; qb += n_qer[s];
        }
    }
    b->n_pairs = n_all; b->n_ref = tot_r; b->n_qer = tot_q;
    b->n_ovf = (int64_t)ovf.size(); b->n_is = n_all - b->n_ovf;
    b->text_mode = text_mode ? 1 : 0;
    // ---- inter-sequence order: device radix sort by (query length, reference length), longest first ----
    // sorted descending, the overflow pairs (key 0xffffffff) come first, then the classes from the widest rows down
    b->is_off[BSW_IS_NCLASS] = b->n_ovf;
    for (int c = BSW_IS_NCLASS - 1; c >= 0; c--) b->is_off[c] = b->is_off[c + 1] + is_cnt[c];
    // (is_off[c+1] .. is_off[c]) is the range of class c in the sorted index array
    if (n_all) {
        size_t need = (size_t)n_all + 1;
        if (need > b->cap_sort) {
            void **bufs[] = {(void **)&b->d_keys, (void **)&b->d_keys2, (void **)&b->d_idx, (void **)&b->d_idx2};
            size_t cap = 0;
            for (void **q : bufs) { size_t c0 = b->cap_sort; if ((rc = meme_grow_device(q, &c0, need, 4, st))) return rc; cap = c0; }
            b->cap_sort = cap;
        }
        k_bsw_keys<<<(unsigned)((n_all + 255) / 256), 256, 0, st>>>(b->d_pairs, n_all, b->match_a, b->d_keys, b->d_idx);
        size_t tmp = 0;
        cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, b->d_keys, b->d_keys2, b->d_idx, b->d_idx2, (int)n_all, 0, 32, st);
        if ((rc = meme_grow_device(&b->d_sort_tmp, &b->cap_sort_tmp, tmp + 256, 1, st))) return rc;
        size_t t2 = b->cap_sort_tmp;
        CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(b->d_sort_tmp, t2, b->d_keys, b->d_keys2, b->d_idx, b->d_idx2, (int)n_all,
                                                          0, 32, st));
        b->launches += 2;
    }
    // ---- overflow classes: counting sort of their indices by the warp kernel's column class ----
    {
        int64_t cnt[BSW_NCLASS] = {0};
        for (int32_t l2 : ovf_len2) cnt[bsw_class_of(l2)]++;
        b->cls_off[0] = 0;
        for (int c = 0; c < BSW_NCLASS; c++) b->cls_off[c + 1] = b->cls_off[c] + cnt[c];
        std::vector<u32> order(ovf.size() + 1);
        std::vector<int32_t> olen(ovf.size() + 1), olen1(ovf.size() + 1);
        int64_t cur[BSW_NCLASS];
        for (int c = 0; c < BSW_NCLASS; c++) cur[c] = b->cls_off[c];
        for (size_t k = 0; k < ovf.size(); k++) { int64_t d = cur[bsw_class_of(ovf_len2[k])]++; order[(size_t)d] = ovf[k]; olen[(size_t)d] = ovf_len2[k]; olen1[(size_t)d] = ovf_len1[k]; }
        if ((rc = meme_grow_device((void **)&b->d_order, &b->cap_order, ovf.size() + 1, 4, st))) return rc;
        // rows of the generic kernel (queries beyond the register kernel): 2 * (len2 + 2) int32 per pair
        const int64_t l0 = b->cls_off[BSW_NCLASS - 1], n_long = b->cls_off[BSW_NCLASS] - l0;
        std::vector<int64_t> off((size_t)n_long + 1);
        int64_t acc = 0;
        for (int64_t i = 0; i < n_long; i++) { off[(size_t)i] = acc; acc += 2 * ((int64_t)olen[(size_t)(l0 + i)] + 2); }
        off[(size_t)n_long] = acc;
        if (n_long) {
            if ((rc = meme_grow_device((void **)&b->d_scratch_off, &b->cap_scratch_off, (size_t)n_long + 1, 8, st))) return rc;
            if ((rc = meme_grow_device((void **)&b->d_scratch, &b->cap_scratch, (size_t)acc + 16, 4, st))) return rc;
            CUDA_TRY(cudaMemcpyAsync(b->d_scratch_off, off.data(), (size_t)(n_long + 1) * 8, cudaMemcpyHostToDevice, st));
        }
        std::vector<int64_t> roff;
        if (!ovf.empty()) {
            CUDA_TRY(cudaMemcpyAsync(b->d_order, order.data(), ovf.size() * 4, cudaMemcpyHostToDevice, st));
            if (text_mode) {
                // the overflow kernels read bytes: write these pairs' reference out of the text (class order = `order`)
                roff.resize(ovf.size() + 1);
                int64_t racc = 0;
                for (size_t k = 0; k < ovf.size(); k++) { roff[k] = racc; racc += olen1[k]; }
                roff[ovf.size()] = racc;
                if ((rc = meme_grow_device((void **)&b->d_ref, &b->cap_ref, (size_t)racc + 64, 1, st))) return rc;
                if ((rc = meme_grow_device((void **)&b->d_ovf_off, &b->cap_ovf_off, ovf.size() + 1, 8, st))) return rc;
                CUDA_TRY(cudaMemcpyAsync(b->d_ovf_off, roff.data(), (ovf.size() + 1) * 8, cudaMemcpyHostToDevice, st));
                k_bsw_materialise<<<(unsigned)ovf.size(), 128, 0, st>>>(b->d_pairs, b->d_order, b->d_ovf_off, (int64_t)ovf.size(), b->text_pac, b->d_ref);
                CUDA_TRY(cudaGetLastError());
            }
            CUDA_TRY(cudaStreamSynchronize(st));       // `order`, `off`, `roff` are temporaries
        }
    }
    return MEME_OK;
}

extern "C" int meme_bsw_upload(meme_bsw_t *b, const meme_seqpair_t *pairs, int64_t n_pairs, const uint8_t *ref,
                               int64_t n_ref, const uint8_t *qer, int64_t n_qer)
{
    if (!b || n_pairs < 0 || n_ref < 0 || n_qer < 0 || (n_pairs && (!pairs || !ref || !qer))) {
        meme_set_error("bad argument");
        return MEME_E_ARG;
    }
    return meme_bsw_upload_segments(b, 1, &pairs, &n_pairs, &ref, &n_ref, &qer, &n_qer);
}

// Results of the last run, segment by segment (the six result fields; offsets are returned rebased to the device
// layout, callers keep their own copies of idr/idq), and the final band width of every pair.
extern "C" int meme_bsw_download_segments(meme_bsw_t *b, int32_t nseg, meme_seqpair_t *const *pairs, const int64_t *n_pairs,
                                          int32_t *const *w_used)
{
    if (!b || nseg < 0 || (nseg && (!pairs || !n_pairs))) return MEME_E_ARG;
    CUDA_TRY(cudaSetDevice(b->device));
    int64_t pb = 0;
    for (int s = 0; s < nseg; s++) {
        if (pb + n_pairs[s] > b->n_pairs) { meme_set_error("segments exceed the uploaded batch"); return MEME_E_ARG; }
        if (n_pairs[s]) {
            CUDA_TRY(cudaMemcpyAsync(pairs[s], b->d_pairs + pb, (size_t)n_pairs[s] * sizeof(meme_seqpair_t), cudaMemcpyDeviceToHost, b->stream));
            if (w_used && w_used[s] && b->d_wused)
                CUDA_TRY(cudaMemcpyAsync(w_used[s], b->d_wused + pb, (size_t)n_pairs[s] * 4, cudaMemcpyDeviceToHost, b->stream));
        }
        pb += n_pairs[s];
    }
    CUDA_TRY(cudaMemcpyAsync(b->h_ctrl, b->d_ctrl, 2 * 8, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return MEME_OK;
}

extern "C" int meme_bsw_run_ext(meme_bsw_t *b, int32_t w, int32_t tries, int32_t prev_is_h0, const meme_bsw_opt_t *opt)
{
    if (!b || !opt) { meme_set_error("null argument"); return MEME_E_ARG; }
    if (opt->e_del <= 0 || opt->e_ins <= 0 || w < 0 || tries < 1 || tries > 8) { meme_set_error("bad BSW options"); return MEME_E_ARG; }
    if (opt->a != b->match_a) { meme_set_error("match score %d differs from the one the classes were built for (%d): call meme_bsw_set_match before the upload", opt->a, b->match_a); return MEME_E_ARG; }
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t st = b->stream;
    int64_t n = b->n_pairs;
    CUDA_TRY(cudaMemsetAsync(b->d_ctrl, 0, 8 * 8, st));
    if (n == 0) return MEME_OK;
    int rc;
    if ((rc = meme_grow_device((void **)&b->d_wused, &b->cap_wused, (size_t)n + 1, 4, st))) return rc;
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device));
    // ---- inter-sequence classes ----
    {
        BswIsConst P = {opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, opt->zdrop, opt->end_bonus, opt->a, opt->b};
        if (!b->attr_set) {                     // per device: function attributes belong to the device's context
            CUDA_TRY(cudaFuncSetAttribute(k_bsw_is, cudaFuncAttributeMaxDynamicSharedMemorySize, BSW_IS_WORDS[BSW_IS_NCLASS - 1] * BSW_IS_THREADS * 5));
            b->attr_set = 1;
        }
        for (int c = BSW_IS_NCLASS - 1; c >= 0; c--) {
            const int64_t lo = b->is_off[c + 1], cn = b->is_off[c] - lo;
            if (cn <= 0) continue;
            const size_t sm = (size_t)BSW_IS_WORDS[c] * BSW_IS_THREADS * 5;      // 4-byte H|E word + 1-byte query base per column
            int per_sm = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bsw_is, BSW_IS_THREADS, sm));
            if (per_sm < 1) per_sm = 1;
            int64_t grid = (int64_t)sms * per_sm, need = (cn + BSW_IS_THREADS - 1) / BSW_IS_THREADS;
            if (grid > need) grid = need;
            k_bsw_is<<<(unsigned)grid, BSW_IS_THREADS, sm, st>>>(P, BSW_IS_WORDS[c], w, tries, prev_is_h0, b->d_pairs, b->d_idx2 + lo, cn, b->d_ref,
                                                              b->text_mode ? b->text_pac : nullptr, b->d_qer,
                                                              b->d_wused, b->d_ctrl);
            CUDA_TRY(cudaGetLastError());
            b->launches++;
        }
    }
    if (b->n_ovf == 0) return MEME_OK;
    // ---- overflow classes ----
    BswConst P = {opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, opt->zdrop, opt->end_bonus, opt->a, opt->b, w};
#define LAUNCH_CLASS(ci, CC)                                                                                          \
    do {                                                                                                              \
        int64_t cn = b->cls_off[(ci) + 1] - b->cls_off[ci];                                                           \
        if (cn > 0) {                                                                                                 \
            int per_sm = 0;                                                                                           \
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bsw<CC>, BSW_WARPS * 32, 0));           \
            if (per_sm < 1) per_sm = 1;                                                                               \
            int64_t grid = (int64_t)sms * per_sm, need = (cn + BSW_WARPS - 1) / BSW_WARPS;                            \
            if (grid > need) grid = need;                                                                             \
            CUDA_TRY(cudaMemsetAsync(b->d_ctrl, 0, 8, st));                                                           \
            k_bsw<CC><<<(unsigned)grid, BSW_WARPS * 32, 0, st>>>(P, tries, prev_is_h0, b->d_pairs, b->d_order + b->cls_off[ci], cn, \
                                                                b->d_ref, b->d_qer, b->d_wused, b->d_ctrl);           \
            CUDA_TRY(cudaGetLastError());                                                                             \
            b->launches++;                                                                                            \
        }                                                                                                             \
    } while (0)
    LAUNCH_CLASS(0, 1); LAUNCH_CLASS(1, 2); LAUNCH_CLASS(2, 3); LAUNCH_CLASS(3, 4); LAUNCH_CLASS(4, 5);
    LAUNCH_CLASS(5, 6); LAUNCH_CLASS(6, 8); LAUNCH_CLASS(7, 10); LAUNCH_CLASS(8, 12);
#undef LAUNCH_CLASS
    const int64_t n_long = b->cls_off[BSW_NCLASS] - b->cls_off[BSW_NCLASS - 1];
    if (n_long > 0) {
        k_bsw_long<<<(unsigned)((n_long + 63) / 64), 64, 0, st>>>(P, tries, prev_is_h0, b->d_pairs, b->d_order + b->cls_off[BSW_NCLASS - 1],
                                                               n_long, b->d_ref, b->d_qer, b->d_scratch_off, b->d_scratch, b->d_wused,
                                                               b->d_ctrl);
        CUDA_TRY(cudaGetLastError());
        b->launches++;
    }
    return MEME_OK;
}

extern "C" int meme_bsw_run(meme_bsw_t *b, int32_t w, const meme_bsw_opt_t *opt)
{
    return meme_bsw_run_ext(b, w, 1, 0, opt);
}

extern "C" int meme_bsw_download(meme_bsw_t *b, meme_seqpair_t *pairs)
{
    if (!b || (b->n_pairs && !pairs)) return MEME_E_ARG;
    CUDA_TRY(cudaSetDevice(b->device));
    if (b->n_pairs)
        CUDA_TRY(cudaMemcpyAsync(pairs, b->d_pairs, (size_t)b->n_pairs * sizeof(meme_seqpair_t), cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaMemcpyAsync(b->h_ctrl, b->d_ctrl, 2 * 8, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return MEME_OK;
}

extern "C" int meme_bsw_download_w(meme_bsw_t *b, int32_t *w_used)
{
    if (!b || (b->n_pairs && (!w_used || !b->d_wused))) return MEME_E_ARG;
    CUDA_TRY(cudaSetDevice(b->device));
    if (b->n_pairs) CUDA_TRY(cudaMemcpyAsync(w_used, b->d_wused, (size_t)b->n_pairs * 4, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return MEME_OK;
}

extern "C" int meme_bsw_cells(meme_bsw_t *b, uint64_t *cells)
{
    if (!b || !cells) return MEME_E_ARG;
    CUDA_TRY(cudaMemcpyAsync(b->h_ctrl, b->d_ctrl, 2 * 8, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    *cells = b->h_ctrl[1];
    return MEME_OK;
}

extern "C" int meme_bsw_batch(meme_bsw_t *b, meme_seqpair_t *pairs, int64_t n_pairs, const uint8_t *ref, int64_t n_ref,
                              const uint8_t *qer, int64_t n_qer, int32_t w, const meme_bsw_opt_t *opt)
{
    int rc;
    if (b && opt && (rc = meme_bsw_set_match(b, opt->a))) return rc;
    if ((rc = meme_bsw_upload(b, pairs, n_pairs, ref, n_ref, qer, n_qer))) return rc;
    if ((rc = meme_bsw_run(b, w, opt))) return rc;
    return meme_bsw_download(b, pairs);
}

extern "C" int meme_bsw_batch_ext(meme_bsw_t *b, meme_seqpair_t *pairs, int64_t n_pairs, const uint8_t *ref, int64_t n_ref,
                                  const uint8_t *qer, int64_t n_qer, int32_t w, int32_t tries, int32_t prev_is_h0,
                                  const meme_bsw_opt_t *opt, int32_t *w_used)
{
    int rc;
    if (b && opt && (rc = meme_bsw_set_match(b, opt->a))) return rc;
    if ((rc = meme_bsw_upload(b, pairs, n_pairs, ref, n_ref, qer, n_qer))) return rc;
    if ((rc = meme_bsw_run_ext(b, w, tries, prev_is_h0, opt))) return rc;
    if ((rc = meme_bsw_download(b, pairs))) return rc;
    return w_used ? meme_bsw_download_w(b, w_used) : MEME_OK;
}

// pinned host memory for callers that assemble large batches (the chunk-level extension driver of the glue)
extern "C" int meme_pinned_alloc(size_t bytes, void **out)
{
    if (!out) return MEME_E_ARG;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) { meme_set_error("cudaMallocHost(%zu bytes): %s", bytes, cudaGetErrorString(e)); return MEME_E_NOMEM; }
    return MEME_OK;
}

extern "C" void meme_pinned_free(void *p)
{
    if (p) cudaFreeHost(p);
}
