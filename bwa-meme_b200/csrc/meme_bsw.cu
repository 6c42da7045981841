// meme_bsw.cu -- kernel B: banded affine-gap Smith-Waterman seed extension on sm_100a.
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
#include "meme_common.cuh"

#define BSW_WARPS 4
#define BSW_NCLASS 10
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

// One kernel instantiation per column count C: registers (and therefore resident warps) follow the class
// instead of the worst case.  `order` lists the pairs of this class (built at upload, counting sort by C).
template <int C>
__global__ void __launch_bounds__(BSW_WARPS * 32)
k_bsw(BswConst P, meme_seqpair_t *__restrict__ pairs, const u32 *__restrict__ order, int64_t n_work,
      const uint8_t *__restrict__ ref, const uint8_t *__restrict__ qer, u64 *__restrict__ ctrl)
{
    const int lane = threadIdx.x & 31;
    u64 cells = 0;
    for (;;) {
        u64 r = 0;
        if (lane == 0) r = atomicAdd((unsigned long long *)&ctrl[0], 1ull);
        r = __shfl_sync(FULL, r, 0);
        if (r >= (u64)n_work) break;
        meme_seqpair_t *p = pairs + order[r];
        bsw_warp<C>(P, qer + p->idq, p->len2, ref + p->idr, p->len1, p->h0, p, cells);
    }
    if (lane == 0 && cells) atomicAdd((unsigned long long *)&ctrl[1], (unsigned long long)cells);
}

// classes: C = 1,2,3,4,5,6,8,10,12 and "long" (generic kernel)
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

__global__ void k_bsw_long_offsets(const meme_seqpair_t *__restrict__ pairs, int64_t n_pairs, int64_t *__restrict__ off)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_pairs) off[i] = (i < n_pairs && pairs[i].len2 > BSW_MAX_REG_QLEN) ? 2 * ((int64_t)pairs[i].len2 + 2) : 0;
}

__global__ void k_bsw_long(BswConst P, meme_seqpair_t *__restrict__ pairs, int64_t n_pairs, const uint8_t *__restrict__ ref,
                           const uint8_t *__restrict__ qer, const int64_t *__restrict__ off, int32_t *__restrict__ scratch,
                           u64 *__restrict__ ctrl)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    meme_seqpair_t *p = pairs + i;
    if (p->len2 <= BSW_MAX_REG_QLEN) return;      // (pairs of the register classes are done by k_bsw<C>)
    u64 cells = 0;
    int32_t *H = scratch + off[i], *E = H + p->len2 + 2;
    bsw_thread(P, qer + p->idq, p->len2, ref + p->idr, p->len1, p->h0, H, E, p, cells);
    atomicAdd((unsigned long long *)&ctrl[1], (unsigned long long)cells);
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
    *out = b;
    return MEME_OK;
}

extern "C" void meme_bsw_free(meme_bsw_t *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    cudaStreamSynchronize(b->stream);
    void *dev[] = {b->d_pairs, b->d_ref, b->d_qer, b->d_scratch, b->d_scratch_off, b->d_ctrl, b->d_order};
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

extern "C" int meme_bsw_upload(meme_bsw_t *b, const meme_seqpair_t *pairs, int64_t n_pairs, const uint8_t *ref,
                               int64_t n_ref, const uint8_t *qer, int64_t n_qer)
{
    if (!b || n_pairs < 0 || n_ref < 0 || n_qer < 0 || (n_pairs && (!pairs || !ref || !qer))) {
        meme_set_error("bad argument");
        return MEME_E_ARG;
    }
    for (int64_t i = 0; i < n_pairs; i++) {
        const meme_seqpair_t &p = pairs[i];
        if (p.len1 < 0 || p.len2 < 0 || p.idr < 0 || p.idq < 0 || (int64_t)p.idr + p.len1 > n_ref ||
            (int64_t)p.idq + p.len2 > n_qer) {
            meme_set_error("pair %lld points outside the sequence buffers", (long long)i);
            return MEME_E_ARG;
        }
    }
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t st = b->stream;
    int rc;
    if ((rc = meme_grow_device((void **)&b->d_pairs, &b->cap_pairs, (size_t)n_pairs + 1, sizeof(meme_seqpair_t), st))) return rc;
    if ((rc = meme_grow_device((void **)&b->d_ref, &b->cap_ref, (size_t)n_ref + 64, 1, st))) return rc;
    if ((rc = meme_grow_device((void **)&b->d_qer, &b->cap_qer, (size_t)n_qer + 64, 1, st))) return rc;
    if (n_pairs) {
        CUDA_TRY(cudaMemcpyAsync(b->d_pairs, pairs, (size_t)n_pairs * sizeof(meme_seqpair_t), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(b->d_ref, ref, (size_t)n_ref, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(b->d_qer, qer, (size_t)n_qer, cudaMemcpyHostToDevice, st));
    }
    b->n_pairs = n_pairs; b->n_ref = n_ref; b->n_qer = n_qer;
    // counting sort of the pair indices by register class (columns per lane); within a class longer
    // references first, so that the dynamic scheduler does not end on a long pair
    {
        int64_t cnt[BSW_NCLASS] = {0};
        for (int64_t i = 0; i < n_pairs; i++) cnt[bsw_class_of(pairs[i].len2)]++;
        b->cls_off[0] = 0;
        for (int c = 0; c < BSW_NCLASS; c++) b->cls_off[c + 1] = b->cls_off[c] + cnt[c];
        std::vector<u32> order((size_t)n_pairs + 1);
        int64_t cur[BSW_NCLASS];
        for (int c = 0; c < BSW_NCLASS; c++) cur[c] = b->cls_off[c];
        for (int64_t i = 0; i < n_pairs; i++) order[(size_t)cur[bsw_class_of(pairs[i].len2)]++] = (u32)i;
        if ((rc = meme_grow_device((void **)&b->d_order, &b->cap_order, (size_t)n_pairs + 1, 4, st))) return rc;
        if (n_pairs) {
            CUDA_TRY(cudaMemcpyAsync(b->d_order, order.data(), (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaStreamSynchronize(st));       // `order` is a temporary
        }
    }
    b->h_ctrl[4] = b->cls_off[BSW_NCLASS] > b->cls_off[BSW_NCLASS - 1];   // any pair for the long-query path?
    return MEME_OK;
}

extern "C" int meme_bsw_run(meme_bsw_t *b, int32_t w, const meme_bsw_opt_t *opt)
{
    if (!b || !opt) { meme_set_error("null argument"); return MEME_E_ARG; }
    if (opt->e_del <= 0 || opt->e_ins <= 0 || w < 0) { meme_set_error("bad BSW options"); return MEME_E_ARG; }
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t st = b->stream;
    int64_t n = b->n_pairs;
    CUDA_TRY(cudaMemsetAsync(b->d_ctrl, 0, 8 * 8, st));
    if (n == 0) return MEME_OK;
    BswConst P = {opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, opt->zdrop, opt->end_bonus, opt->a, opt->b, w};
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device));
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
            k_bsw<CC><<<(unsigned)grid, BSW_WARPS * 32, 0, st>>>(P, b->d_pairs, b->d_order + b->cls_off[ci], cn, b->d_ref, \
                                                                b->d_qer, b->d_ctrl);                                 \
            CUDA_TRY(cudaGetLastError());                                                                             \
            b->launches++;                                                                                            \
        }                                                                                                             \
    } while (0)
    LAUNCH_CLASS(0, 1); LAUNCH_CLASS(1, 2); LAUNCH_CLASS(2, 3); LAUNCH_CLASS(3, 4); LAUNCH_CLASS(4, 5);
    LAUNCH_CLASS(5, 6); LAUNCH_CLASS(6, 8); LAUNCH_CLASS(7, 10); LAUNCH_CLASS(8, 12);
#undef LAUNCH_CLASS
    if (b->h_ctrl[4]) {
        int rc;
        if ((rc = meme_grow_device((void **)&b->d_scratch_off, &b->cap_scratch_off, (size_t)n + 1, 8, st))) return rc;
        k_bsw_long_offsets<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(b->d_pairs, n, b->d_scratch_off);
        // exclusive scan on the host side is avoided: offsets are small in number, do it with one thread block
        // (long queries are rare on this path: reads are <= 301 bp)
        std::vector<int64_t> off((size_t)n + 1);
        CUDA_TRY(cudaMemcpyAsync(off.data(), b->d_scratch_off, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        int64_t acc = 0;
        for (int64_t i = 0; i <= n; i++) { int64_t v = off[i]; off[i] = acc; acc += v; }
        CUDA_TRY(cudaMemcpyAsync(b->d_scratch_off, off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
        if ((rc = meme_grow_device((void **)&b->d_scratch, &b->cap_scratch, (size_t)acc + 16, 4, st))) return rc;
        k_bsw_long<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(P, b->d_pairs, n, b->d_ref, b->d_qer, b->d_scratch_off,
                                                          b->d_scratch, b->d_ctrl);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(st));
        b->launches += 2;
    }
    return MEME_OK;
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
    if ((rc = meme_bsw_upload(b, pairs, n_pairs, ref, n_ref, qer, n_qer))) return rc;
    if ((rc = meme_bsw_run(b, w, opt))) return rc;
    return meme_bsw_download(b, pairs);
}
