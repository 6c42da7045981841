// meme_matesw.cu -- mate-rescue Smith-Waterman on the device: replaces mem_sam_pe_batch
// (reference src/bwamem_pair.cpp:L719-817 = kswv::getScores8/16 phases 0 and 1, src/kswv.cpp) for all pairs of a chunk at once.
// The reference runs the step per 512-read block inside worker_sam, 64 / 32 pairs per AVX512 register; here the pairs of a
// whole chunk, assembled by several host threads (one segment each), go through one call.
//
// One thread per pair (meme_matesw_core.cuh).  Pairs are ordered by class (8-bit first, as sort_classify does) and by
// reference length (longest first: the 32 pairs of a warp sweep the same number of rows, and the long tail starts first);
// a launch covers one class of at most 2^17 pairs.  The row state of a thread lives in dynamic shared memory sized by the
// longest query of the launch, laid out [column][thread]; the row maxima of a warp are lane-interleaved in a global scratch
// array carved by the host (it knows every length).
#include <vector>
#include <algorithm>
#include "meme_common.cuh"
#include "meme_matesw_core.cuh"

struct MswTask { int64_t r_off, q_off; int32_t len1, len2, xtra, pad; };

struct MateswWork {
    MswTask *d_tasks; size_t cap_tasks;
    uint32_t *d_perm; size_t cap_perm;
    int64_t *d_rm_off; size_t cap_rm_off;
    uint8_t *d_ref, *d_qer; size_t cap_ref, cap_qer;
    int16_t *d_rm; size_t cap_rm;
    meme_kswr_t *d_out; size_t cap_out;
    u64 *d_ctl;
    MswTask *h_tasks; size_t hcap_tasks;
    uint32_t *h_perm; size_t hcap_perm;
    int64_t *h_rm_off; size_t hcap_rm_off;
    cudaEvent_t ev[2];
    float ms; u64 cells; int ran;
};

void meme_matesw_free(MateswWork *c)
{
    if (!c) return;
    void *dev[] = {c->d_tasks, c->d_perm, c->d_rm_off, c->d_ref, c->d_qer, c->d_rm, c->d_out, c->d_ctl};
    for (void *p : dev) if (p) cudaFree(p);
    void *host[] = {c->h_tasks, c->h_perm, c->h_rm_off};
    for (void *p : host) if (p) cudaFreeHost(p);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    free(c);
}

static int matesw_work(meme_index_t *ix, MateswWork **out)
{
    if (!ix->matesw) {
        MateswWork *c = (MateswWork *)calloc(1, sizeof(MateswWork));
        if (!c) return MEME_E_NOMEM;
        ix->matesw = c;
        CUDA_TRY(cudaSetDevice(ix->device));
        CUDA_TRY(cudaMalloc((void **)&c->d_ctl, 16));
        for (auto &e : c->ev) CUDA_TRY(cudaEventCreate(&e));
    }
    *out = ix->matesw;
    return MEME_OK;
}

#define MSW_TPB 64

__global__ void __launch_bounds__(MSW_TPB) k_matesw(const MswTask *__restrict__ tasks, const uint32_t *__restrict__ perm, int64_t first, int64_t n,
                                                   const uint8_t *__restrict__ ref, const uint8_t *__restrict__ qer, msw::Opt opt, int cols_max,
                                                   int16_t *__restrict__ rm, const int64_t *__restrict__ rm_off, meme_kswr_t *__restrict__ out,
                                                   u64 *__restrict__ ctl)
{
    extern __shared__ uint32_t smem[];
    const int64_t t = (int64_t)blockIdx.x * MSW_TPB + threadIdx.x;
    if (t >= n) return;
    const int64_t tix = first + t;
    const uint32_t id = perm[tix];
    const MswTask T = tasks[id];
    msw::Scratch S;
    S.row = smem + threadIdx.x;
    S.qrow = reinterpret_cast<uint8_t *>(smem + (size_t)(cols_max + 1) * MSW_TPB) + threadIdx.x;
    S.stride = MSW_TPB;
    S.rm = rm + rm_off[t >> 5] + (t & 31);
    S.rstride = 32;
    msw::Out A;
    uint64_t cells = 0;
    msw::pair(opt, ref + T.r_off, T.len1, qer + T.q_off, T.len2, T.xtra, S, A, cells);
    meme_kswr_t r;
    r.score = A.score; r.te = A.te; r.qe = A.qe; r.score2 = A.score2; r.te2 = A.te2; r.tb = A.tb; r.qb = A.qb;
    out[id] = r;
    atomicAdd((unsigned long long *)&ctl[1], (unsigned long long)cells);
}

static size_t msw_smem_bytes(int cols_max) { return (size_t)MSW_TPB * ((size_t)(cols_max + 1) * 4 + (size_t)cols_max + 4); }

extern "C" int meme_matesw_batch(meme_index_t *ix, int32_t nseg, const meme_seqpair_t *const *pairs, const int64_t *n_pairs, const uint8_t *const *ref,
                                 const int64_t *n_ref, const uint8_t *const *qer, const int64_t *n_qer, const meme_matesw_opt_t *opt,
                                 meme_kswr_t *const *aln)
{
    if (!ix || nseg < 0 || !opt || (nseg > 0 && (!pairs || !n_pairs || !ref || !n_ref || !qer || !n_qer || !aln))) { meme_set_error("bad argument"); return MEME_E_ARG; }
    if (opt->e_del <= 0 || opt->e_ins <= 0 || opt->o_del < 0 || opt->o_ins < 0 || opt->a <= 0 || opt->a > 100 || opt->b < 0 || opt->b > 100) {
        meme_set_error("scoring parameters out of range"); return MEME_E_ARG;
    }
    MateswWork *c; int rc;
    if ((rc = matesw_work(ix, &c))) return rc;
    CUDA_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    c->ran = 0; c->ms = 0; c->cells = 0;
    int64_t n = 0, tot_ref = 0, tot_qer = 0;
    for (int s = 0; s < nseg; s++) {
        if (n_pairs[s] < 0 || n_ref[s] < 0 || n_qer[s] < 0 || (n_pairs[s] > 0 && (!pairs[s] || !ref[s] || !qer[s] || !aln[s]))) { meme_set_error("bad segment %d", s); return MEME_E_ARG; }
        n += n_pairs[s]; tot_ref += n_ref[s]; tot_qer += n_qer[s];
    }
    if (n >= (1ll << 31)) { meme_set_error("too many pairs"); return MEME_E_ARG; }
    if (n == 0) { c->ran = 1; return MEME_OK; }
    if ((rc = meme_grow_pinned((void **)&c->h_tasks, &c->hcap_tasks, (size_t)n, sizeof(MswTask)))) return rc;
    if ((rc = meme_grow_pinned((void **)&c->h_perm, &c->hcap_perm, (size_t)n, 4))) return rc;
    if ((rc = meme_grow_pinned((void **)&c->h_rm_off, &c->hcap_rm_off, (size_t)n / 32 + 2, 8))) return rc;
    // ---- tasks in segment order (= output order), limits, class counts ----
    int64_t k0 = 0, rb = 0, qb = 0, n8 = 0;
    int max_len1 = 0;
    for (int s = 0; s < nseg; s++) {
        for (int64_t k = 0; k < n_pairs[s]; k++) {
            const meme_seqpair_t &p = pairs[s][k];
            const bool byte = (p.h0 & msw::XBYTE) != 0;
            if (p.len1 < 0 || p.len2 < 0 || p.idr < 0 || p.idq < 0 || (int64_t)p.idr + p.len1 > n_ref[s] || (int64_t)p.idq + p.len2 > n_qer[s]) {
                meme_set_error("pair %lld of segment %d lies outside its buffers", (long long)k, s); return MEME_E_ARG;
            }
            if (p.len1 > msw::MAX_ROWS || p.len2 > msw::MAX_COLS - 16 || (byte && p.len2 > 255) || (int64_t)p.len2 * opt->a > 32000) {
                meme_set_error("pair %lld of segment %d: lengths %d x %d are beyond the kernel's limits (%d rows, %d columns, 255 in the 8-bit class)",
                               (long long)k, s, p.len1, p.len2, (int)msw::MAX_ROWS, (int)msw::MAX_COLS - 16);
                return MEME_E_ARG;
            }
            MswTask &T = c->h_tasks[k0 + k];
            T.r_off = rb + p.idr; T.q_off = qb + p.idq; T.len1 = p.len1; T.len2 = p.len2; T.xtra = p.h0; T.pad = 0;
            n8 += byte;
            max_len1 = p.len1 > max_len1 ? p.len1 : max_len1;
        }
        k0 += n_pairs[s]; rb += n_ref[s]; qb += n_qer[s];
    }
    // ---- order: class, then reference length descending (counting sort) ----
    {
        std::vector<int64_t> cnt((size_t)2 * (max_len1 + 2), 0);
        auto bucket = [&](const MswTask &T) { return (size_t)((T.xtra & msw::XBYTE) ? 0 : max_len1 + 1) + (size_t)(max_len1 - T.len1); };
        for (int64_t k = 0; k < n; k++) cnt[bucket(c->h_tasks[k]) + 1]++;
        for (size_t b = 1; b < cnt.size(); b++) cnt[b] += cnt[b - 1];
        for (int64_t k = 0; k < n; k++) c->h_perm[cnt[bucket(c->h_tasks[k])]++] = (uint32_t)k;
    }
    // ---- upload ----
    if ((rc = meme_grow_device((void **)&c->d_tasks, &c->cap_tasks, (size_t)n, sizeof(MswTask), st))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_perm, &c->cap_perm, (size_t)n, 4, st))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_ref, &c->cap_ref, (size_t)tot_ref + 64, 1, st))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_qer, &c->cap_qer, (size_t)tot_qer + 64, 1, st))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_out, &c->cap_out, (size_t)n, sizeof(meme_kswr_t), st))) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->d_tasks, c->h_tasks, (size_t)n * sizeof(MswTask), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c->d_perm, c->h_perm, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    rb = qb = 0;
    for (int s = 0; s < nseg; s++) {
        if (n_ref[s]) CUDA_TRY(cudaMemcpyAsync(c->d_ref + rb, ref[s], (size_t)n_ref[s], cudaMemcpyHostToDevice, st));
        if (n_qer[s]) CUDA_TRY(cudaMemcpyAsync(c->d_qer + qb, qer[s], (size_t)n_qer[s], cudaMemcpyHostToDevice, st));
        rb += n_ref[s]; qb += n_qer[s];
    }
    CUDA_TRY(cudaMemsetAsync(c->d_ctl, 0, 16, st));
    msw::Opt O = {opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, opt->a, opt->b};
    const int64_t SLICE = 1 << 17;
    CUDA_TRY(cudaEventRecord(c->ev[0], st));
    for (int cls = 0; cls < 2; cls++) {
        const int64_t c0 = cls == 0 ? 0 : n8, c1 = cls == 0 ? n8 : n;
        for (int64_t s0 = c0; s0 < c1; s0 += SLICE) {
            const int64_t m = c1 - s0 < SLICE ? c1 - s0 : SLICE;
            const int64_t warps = (m + 31) / 32;
            int cols_max = 0;
            int64_t rm_tot = 0;
            for (int64_t w = 0; w < warps; w++) {
                c->h_rm_off[w] = rm_tot;
                rm_tot += ((int64_t)c->h_tasks[c->h_perm[s0 + w * 32]].len1 + 1) * 32;      // longest first within the class
            }
            for (int64_t k = 0; k < m; k++) {
                const MswTask &T = c->h_tasks[c->h_perm[s0 + k]];
                const int q = msw::quanta(T.len2, cls == 0);
                cols_max = q > cols_max ? q : cols_max;
            }
            if ((rc = meme_grow_device((void **)&c->d_rm_off, &c->cap_rm_off, (size_t)warps + 1, 8, st))) return rc;
            if ((rc = meme_grow_device((void **)&c->d_rm, &c->cap_rm, (size_t)rm_tot + 64, 2, st))) return rc;
            CUDA_TRY(cudaMemcpyAsync(c->d_rm_off, c->h_rm_off, (size_t)warps * 8, cudaMemcpyHostToDevice, st));
            const size_t smem = msw_smem_bytes(cols_max);
            CUDA_TRY(cudaFuncSetAttribute(k_matesw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_matesw<<<(unsigned)((m + MSW_TPB - 1) / MSW_TPB), MSW_TPB, smem, st>>>(c->d_tasks, c->d_perm, s0, m, c->d_ref, c->d_qer, O, cols_max, c->d_rm,
                                                                                   c->d_rm_off, c->d_out, c->d_ctl);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaStreamSynchronize(st));       // h_rm_off is reused by the next slice
            ix->launches += 1;
        }
    }
    CUDA_TRY(cudaEventRecord(c->ev[1], st));
    k0 = 0;
    for (int s = 0; s < nseg; s++) {
        if (n_pairs[s]) CUDA_TRY(cudaMemcpyAsync(aln[s], c->d_out + k0, (size_t)n_pairs[s] * sizeof(meme_kswr_t), cudaMemcpyDeviceToHost, st));
        k0 += n_pairs[s];
    }
    u64 ctl[2];
    CUDA_TRY(cudaMemcpyAsync(ctl, c->d_ctl, 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaEventElapsedTime(&c->ms, c->ev[0], c->ev[1]));
    c->cells = ctl[1];
    c->ran = 1;
    return MEME_OK;
}

extern "C" int meme_matesw_last(meme_index_t *ix, float *ms, uint64_t *cells)
{
    if (!ix || !ix->matesw || !ix->matesw->ran) { meme_set_error("no completed meme_matesw_batch"); return MEME_E_ARG; }
    if (ms) *ms = ix->matesw->ms;
    if (cells) *cells = ix->matesw->cells;
    return MEME_OK;
}

// Kernel loaded and the buffers sized for about n_pairs pairs of len1 x len2 ahead of the first timed chunk.
extern "C" int meme_matesw_warmup(meme_index_t *ix, int64_t n_pairs, int len1, int len2)
{
    if (!ix || n_pairs < 0 || len1 < 1 || len2 < 1 || len1 > msw::MAX_ROWS || len2 > msw::MAX_COLS - 16) { meme_set_error("bad argument"); return MEME_E_ARG; }
    MateswWork *c; int rc;
    if ((rc = matesw_work(ix, &c))) return rc;
    CUDA_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    if ((rc = meme_grow_pinned((void **)&c->h_tasks, &c->hcap_tasks, (size_t)n_pairs + 64, sizeof(MswTask)))) return rc;
    if ((rc = meme_grow_pinned((void **)&c->h_perm, &c->hcap_perm, (size_t)n_pairs + 64, 4))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_ref, &c->cap_ref, (size_t)n_pairs * (size_t)len1 + 64, 1, st))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_qer, &c->cap_qer, (size_t)n_pairs * (size_t)len2 + 64, 1, st))) return rc;
    {   // the per-pair arrays and the row-maximum scratch of one launch slice
        const size_t slice = (size_t)(n_pairs < (1 << 17) ? n_pairs : (1 << 17));
        if ((rc = meme_grow_device((void **)&c->d_tasks, &c->cap_tasks, (size_t)n_pairs + 64, sizeof(MswTask), st))) return rc;
        if ((rc = meme_grow_device((void **)&c->d_perm, &c->cap_perm, (size_t)n_pairs + 64, 4, st))) return rc;
        if ((rc = meme_grow_device((void **)&c->d_out, &c->cap_out, (size_t)n_pairs + 64, sizeof(meme_kswr_t), st))) return rc;
        if ((rc = meme_grow_device((void **)&c->d_rm_off, &c->cap_rm_off, slice / 32 + 2, 8, st))) return rc;
        if ((rc = meme_grow_device((void **)&c->d_rm, &c->cap_rm, (slice + 32) * ((size_t)len1 + 1) + 64, 2, st))) return rc;
        if ((rc = meme_grow_pinned((void **)&c->h_rm_off, &c->hcap_rm_off, (size_t)n_pairs / 32 + 2, 8))) return rc;
    }
    const int64_t n = n_pairs < 2048 ? n_pairs : 2048;
    if (n == 0) return MEME_OK;
    std::vector<meme_seqpair_t> pairs((size_t)n);
    std::vector<uint8_t> ref((size_t)n * 64), qer((size_t)n * 32);
    std::vector<meme_kswr_t> aln((size_t)n);
    u64 x = 88172645463325252ull;
    for (auto &b : ref) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; b = (uint8_t)(x & 3); }
    for (auto &b : qer) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; b = (uint8_t)(x & 3); }
    for (int64_t k = 0; k < n; k++) {
        meme_seqpair_t &p = pairs[(size_t)k];
        memset(&p, 0, sizeof p);
        p.idr = (int32_t)(k * 64); p.idq = (int32_t)(k * 32); p.len1 = 64; p.len2 = 32;
        p.h0 = msw::XSUBO | msw::XSTART | ((k & 1) ? msw::XBYTE : 0) | 5;
    }
    const meme_seqpair_t *pp = pairs.data(); const uint8_t *rp = ref.data(), *qp = qer.data(); meme_kswr_t *ap = aln.data();
    const int64_t np = n, nr = (int64_t)ref.size(), nq = (int64_t)qer.size();
    meme_matesw_opt_t o = {6, 1, 6, 1, 1, 4};
    rc = meme_matesw_batch(ix, 1, &pp, &np, &rp, &nr, &qp, &nq, &o, &ap);
    c->ran = 0;
    return rc;
}
