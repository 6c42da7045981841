// meme_chain.cu -- chaining + chain filtering on the device, behind
// meme_chain_run / meme_chain_download / meme_chain_batch_multi of include/meme_b200.h.
//
// Replaces, for a whole batch of reads at once, what mem_kernel1_core_Learned does per read after its ks_introsort:
// mem_chain_Learned (src/bwamem.cpp:L1122-1204) and mem_chain_flt (L599-717).  Input is what meme_seed_run left in HBM
// (SMEM records, sampled hit lists, offsets); output is the filtered chains with their seeds, compact, in read order.  The
// SMEM and hit lists never travel to the host on this path.
//
// One thread per read (meme_chain_core.cuh has the per-read procedure and says why the reference's B-tree and sort are
// reproduced step by step).  A read's scratch -- seed links, chain records, tree nodes -- is carved out of arrays sized by the
// batch's hit count with the offsets the seeding run already computed (a read never has more chains than seeds, and a
// B-tree of order 5 never has more than keys / 4 + 1 nodes), so there is no per-read limit and no allocation in the kernel.
// Two passes: k_chain_build chains and filters and leaves the surviving chain ids in the read's scratch, an exclusive scan
// turns the per-read counts into offsets, k_chain_emit writes the compact records.
#include <cub/cub.cuh>
#include <string>
#include <thread>
#include <vector>
#include "meme_common.cuh"
#include "meme_chain_core.cuh"

using namespace chn;

struct ChainWork {
    meme_contig_t *d_contigs; int n_contigs; int64_t l_pac;
    cu32 *d_next, *d_ql, *d_order, *d_kept; size_t cap_next, cap_ql, cap_order, cap_kept;
    Rec *d_ch; size_t cap_ch;
    Node *d_nodes; size_t cap_nodes;
    int64_t *d_chain_off, *d_seed_off; float *d_frac; size_t cap_chain_off, cap_seed_off, cap_frac;
    meme_chain_t *d_chains; size_t cap_chains;
    meme_chain_seed_t *d_seeds; size_t cap_seeds;
    int *d_err;
    int64_t n_reads, n_chains, n_seeds;
    // pinned results
    int64_t *h_chain_off; float *h_frac; meme_chain_t *h_chains; meme_chain_seed_t *h_seeds;
    size_t hcap_chain_off, hcap_frac, hcap_chains, hcap_seeds;
    cudaEvent_t ev[2];
    int ran;
};

void meme_chain_free(ChainWork *c)
{
    if (!c) return;
    void *dev[] = {c->d_contigs, c->d_next, c->d_ql, c->d_order, c->d_kept, c->d_ch, c->d_nodes, c->d_chain_off, c->d_seed_off,
                   c->d_frac, c->d_chains, c->d_seeds, c->d_err};
    for (void *p : dev) if (p) cudaFree(p);
    void *host[] = {c->h_chain_off, c->h_frac, c->h_chains, c->h_seeds};
    for (void *p : host) if (p) cudaFreeHost(p);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    free(c);
}

static int chain_work(meme_index_t *ix, ChainWork **out)
{
    if (!ix->chain) {
        ChainWork *c = (ChainWork *)calloc(1, sizeof(ChainWork));
        if (!c) return MEME_E_NOMEM;
        ix->chain = c;
        CUDA_TRY(cudaSetDevice(ix->device));
        CUDA_TRY(cudaMalloc((void **)&c->d_err, 8));
        for (auto &e : c->ev) CUDA_TRY(cudaEventCreate(&e));
    }
    *out = ix->chain;
    return MEME_OK;
}

struct ChainArgs {
    const meme_smem_t *smems; const u64 *hits; const int64_t *smem_off, *hit_off, *read_off;
    int64_t n_reads;
    const meme_contig_t *contigs; int n_contigs; int64_t l_pac;
    meme_chain_opt_t opt;
    cu32 *next, *ql, *order, *kept; Rec *ch; Node *nodes;
    int64_t *chain_cnt, *seed_cnt; float *frac;
    int *err;
};

__device__ __forceinline__ void read_slices(const ChainArgs &A, int64_t r, View &v, Scratch &S)
{
    const int64_t s0 = A.smem_off[r], h0 = A.hit_off[r], h1 = A.hit_off[r + 1];
    v.smems = A.smems + s0; v.n_smems = (int)(A.smem_off[r + 1] - s0);
    v.hits = (const cu64 *)(A.hits + h0);
    v.read_len = (int)(A.read_off[r + 1] - A.read_off[r]);
    v.contigs = A.contigs; v.n_contigs = A.n_contigs; v.l_pac = A.l_pac;
    S.next = A.next + h0; S.ql = A.ql + h0; S.ch = A.ch + h0; S.order = A.order + h0; S.kept = A.kept + h0;
    S.nodes = A.nodes + (h0 / 4 + 2 * r);
    S.max_nodes = (int)((h1 - h0) / 4 + 2);
}

__global__ void __launch_bounds__(128) k_chain_build(ChainArgs A)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > A.n_reads) return;
    if (r == A.n_reads) { A.chain_cnt[r] = 0; A.seed_cnt[r] = 0; return; }
    View v; Scratch S;
    read_slices(A, r, v, S);
    int n_out = 0, err = 0; float frac = 0.f;
    const int n_seeds = chain_read(A.opt, v, S, &n_out, &frac, &err);
    if (err) atomicMax(A.err, err);
    A.chain_cnt[r] = n_out; A.seed_cnt[r] = n_seeds; A.frac[r] = frac;
}

__global__ void __launch_bounds__(128) k_chain_emit(ChainArgs A, const int64_t *__restrict__ chain_off, const int64_t *__restrict__ seed_off,
                                                    meme_chain_t *__restrict__ chains, meme_chain_seed_t *__restrict__ seeds)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= A.n_reads) return;
    const int64_t c0 = chain_off[r];
    const int n = (int)(chain_off[r + 1] - c0);
    if (n == 0) return;
    View v; Scratch S;
    read_slices(A, r, v, S);
    int64_t sb = seed_off[r];
    for (int i = 0; i < n; i++) {
        const Rec &c = S.ch[S.order[i]];
        meme_chain_t o;
        o.pos = (int64_t)v.hits[c.first_seed]; o.seed_beg = sb; o.n = c.n; o.rid = c.rid; o.w = c.w; o.first = c.first;
        o.kept = c.kept; o.is_alt = c.is_alt;
        chains[c0 + i] = o;
        cu32 s = c.first_seed;
        for (int j = 0; j < c.n; j++, s = S.next[s]) {
            meme_chain_seed_t q;
            q.rbeg = (int64_t)v.hits[s]; q.qbeg = (int32_t)(S.ql[s] & 0xffff); q.len = (int32_t)(S.ql[s] >> 16);
            seeds[sb + j] = q;
        }
        sb += c.n;
    }
}

extern "C" int meme_chain_set_contigs(meme_index_t *ix, const meme_contig_t *contigs, int32_t n_contigs, int64_t l_pac)
{
    if (!ix || !contigs || n_contigs < 1 || l_pac < 1) { meme_set_error("bad argument"); return MEME_E_ARG; }
    if ((u64)l_pac * 2 > ix->v.sa_num) { meme_set_error("l_pac %lld does not fit an index of %llu suffixes", (long long)l_pac, ix->v.sa_num); return MEME_E_ARG; }
    ChainWork *c; int rc;
    if ((rc = chain_work(ix, &c))) return rc;
    CUDA_TRY(cudaSetDevice(ix->device));
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    if (c->d_contigs) { CUDA_TRY(cudaFree(c->d_contigs)); c->d_contigs = nullptr; }
    CUDA_TRY(cudaMalloc((void **)&c->d_contigs, (size_t)n_contigs * sizeof(meme_contig_t)));
    CUDA_TRY(cudaMemcpy(c->d_contigs, contigs, (size_t)n_contigs * sizeof(meme_contig_t), cudaMemcpyHostToDevice));
    c->n_contigs = n_contigs; c->l_pac = l_pac;
    return MEME_OK;
}

extern "C" int meme_chain_run(meme_index_t *ix, const meme_chain_opt_t *opt)
{
    if (!ix || !opt) { meme_set_error("bad argument"); return MEME_E_ARG; }
    if (!ix->ran || ix->piped) { meme_set_error("meme_chain_run needs the device-resident lists of a meme_seed_run on this handle"); return MEME_E_ARG; }
    ChainWork *c = ix->chain;
    if (!c || !c->d_contigs) { meme_set_error("meme_chain_set_contigs has not been called"); return MEME_E_ARG; }
    if (opt->max_occ <= 0 || ix->last_opt.max_occ != opt->max_occ) {
        meme_set_error("the seeding run sampled hits with max_occ %d, chaining asks for %d (must be equal and > 0)", ix->last_opt.max_occ, opt->max_occ);
        return MEME_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    const int64_t n = ix->n_reads;
    const int64_t nh = n ? (int64_t)ix->h_ctrl[33] : 0;
    int rc;
    c->ran = 0; c->n_reads = n; c->n_chains = 0; c->n_seeds = 0;
    if ((rc = meme_grow_device((void **)&c->d_chain_off, &c->cap_chain_off, (size_t)n + 2, 8, st))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_seed_off, &c->cap_seed_off, (size_t)n + 2, 8, st))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_frac, &c->cap_frac, (size_t)n + 2, 4, st))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_next, &c->cap_next, (size_t)nh + 1, 4, st))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_ql, &c->cap_ql, (size_t)nh + 1, 4, st))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_order, &c->cap_order, (size_t)nh + 1, 4, st))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_kept, &c->cap_kept, (size_t)nh + 1, 4, st))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_ch, &c->cap_ch, (size_t)nh + 1, sizeof(Rec), st))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_nodes, &c->cap_nodes, (size_t)nh / 4 + 2 * (size_t)n + 4, sizeof(Node), st))) return rc;
    CUDA_TRY(cudaEventRecord(c->ev[0], st));
    if (n == 0) { CUDA_TRY(cudaEventRecord(c->ev[1], st)); c->ran = 1; return MEME_OK; }
    ChainArgs A;
    A.smems = ix->d_smems; A.hits = ix->d_hits; A.smem_off = ix->d_smem_off; A.hit_off = ix->d_hit_off; A.read_off = ix->d_read_off;
    A.n_reads = n; A.contigs = c->d_contigs; A.n_contigs = c->n_contigs; A.l_pac = c->l_pac; A.opt = *opt;
    A.next = c->d_next; A.ql = c->d_ql; A.order = c->d_order; A.kept = c->d_kept; A.ch = c->d_ch; A.nodes = c->d_nodes;
    A.chain_cnt = c->d_chain_off; A.seed_cnt = c->d_seed_off; A.frac = c->d_frac; A.err = c->d_err;
    CUDA_TRY(cudaMemsetAsync(c->d_err, 0, 4, st));
    const unsigned blocks = (unsigned)((n + 1 + 127) / 128);
    k_chain_build<<<blocks, 128, 0, st>>>(A);
    CUDA_TRY(cudaGetLastError());
    {
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->d_chain_off, c->d_chain_off, (int)(n + 1), st);
        if ((rc = meme_grow_device(&ix->d_scan_tmp, &ix->cap_scan_tmp, tmp + 256, 1, st))) return rc;
        size_t t2 = ix->cap_scan_tmp;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(ix->d_scan_tmp, t2, c->d_chain_off, c->d_chain_off, (int)(n + 1), st));
        t2 = ix->cap_scan_tmp;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(ix->d_scan_tmp, t2, c->d_seed_off, c->d_seed_off, (int)(n + 1), st));
    }
    int64_t tot[2]; int err = 0;
    CUDA_TRY(cudaMemcpyAsync(&tot[0], c->d_chain_off + n, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&tot[1], c->d_seed_off + n, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&err, c->d_err, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (err) { meme_set_error("chaining kernel: %s", err == 1 ? "per-read scratch exhausted" : "sort stack exhausted"); return MEME_E_OVERFLOW; }
    if ((rc = meme_grow_device((void **)&c->d_chains, &c->cap_chains, (size_t)tot[0] + 1, sizeof(meme_chain_t), st))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_seeds, &c->cap_seeds, (size_t)tot[1] + 1, sizeof(meme_chain_seed_t), st))) return rc;
    k_chain_emit<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(A, c->d_chain_off, c->d_seed_off, c->d_chains, c->d_seeds);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev[1], st));
    ix->launches += 4;
    c->n_chains = tot[0]; c->n_seeds = tot[1];
    c->ran = 1;
    return MEME_OK;
}

extern "C" int meme_chain_last_ms(meme_index_t *ix, float *ms)
{
    if (!ix || !ms || !ix->chain || !ix->chain->ran) { meme_set_error("no completed meme_chain_run"); return MEME_E_ARG; }
    CUDA_TRY(cudaSetDevice(ix->device));
    CUDA_TRY(cudaEventSynchronize(ix->chain->ev[1]));
    CUDA_TRY(cudaEventElapsedTime(ms, ix->chain->ev[0], ix->chain->ev[1]));
    return MEME_OK;
}

// grow a pinned array, keeping its first `keep` elements
static int chain_grow_pinned(void **p, size_t *cap, size_t need, size_t elem, size_t keep)
{
    if (need <= *cap && *p) return MEME_OK;
    size_t ncap = need + need / 4 + 64;
    void *q = nullptr;
    cudaError_t e = cudaMallocHost(&q, ncap * elem);
    if (e != cudaSuccess) { meme_set_error("cudaMallocHost(%zu bytes): %s", ncap * elem, cudaGetErrorString(e)); return MEME_E_NOMEM; }
    if (*p) { if (keep) memcpy(q, *p, keep * elem); cudaFreeHost(*p); }
    *p = q; *cap = ncap;
    return MEME_OK;
}

// the results of the last run appended to the pinned arrays after r0 reads / c0 chains / s0 seeds already there
static int chain_append(meme_index_t *ix, int64_t r0, int64_t c0, int64_t s0)
{
    ChainWork *c = ix->chain;
    cudaStream_t st = ix->stream;
    const int64_t n = c->n_reads, nc = c->n_chains, nsd = c->n_seeds;
    int rc;
    if ((rc = chain_grow_pinned((void **)&c->h_chain_off, &c->hcap_chain_off, (size_t)(r0 + n) + 1, 8, (size_t)r0 + 1))) return rc;
    if ((rc = chain_grow_pinned((void **)&c->h_frac, &c->hcap_frac, (size_t)(r0 + n) + 1, 4, (size_t)r0))) return rc;
    if ((rc = chain_grow_pinned((void **)&c->h_chains, &c->hcap_chains, (size_t)(c0 + nc) + 1, sizeof(meme_chain_t), (size_t)c0))) return rc;
    if ((rc = chain_grow_pinned((void **)&c->h_seeds, &c->hcap_seeds, (size_t)(s0 + nsd) + 1, sizeof(meme_chain_seed_t), (size_t)s0))) return rc;
    if (n == 0) { c->h_chain_off[r0] = c0; return MEME_OK; }
    CUDA_TRY(cudaMemcpyAsync(c->h_chain_off + r0, c->d_chain_off, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(c->h_frac + r0, c->d_frac, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (nc) CUDA_TRY(cudaMemcpyAsync(c->h_chains + c0, c->d_chains, (size_t)nc * sizeof(meme_chain_t), cudaMemcpyDeviceToHost, st));
    if (nsd) CUDA_TRY(cudaMemcpyAsync(c->h_seeds + s0, c->d_seeds, (size_t)nsd * sizeof(meme_chain_seed_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (c0 || s0) {                       // a later sub-batch: rebase onto what is already there
        for (int64_t i = 0; i <= n; i++) c->h_chain_off[r0 + i] += c0;
        for (int64_t i = 0; i < nc; i++) c->h_chains[c0 + i].seed_beg += s0;
    }
    return MEME_OK;
}

static void chain_result(ChainWork *c, int64_t n, int64_t nc, int64_t nsd, meme_chain_result_t *out)
{
    out->n_reads = n; out->n_chains = nc; out->n_seeds = nsd;
    out->chain_off = c->h_chain_off; out->frac_rep = c->h_frac; out->chains = c->h_chains; out->seeds = c->h_seeds;
}

extern "C" int meme_chain_download(meme_index_t *ix, meme_chain_result_t *out)
{
    if (!ix || !out || !ix->chain || !ix->chain->ran) { meme_set_error("no completed meme_chain_run"); return MEME_E_ARG; }
    CUDA_TRY(cudaSetDevice(ix->device));
    int rc = chain_append(ix, 0, 0, 0);
    if (rc) return rc;
    ChainWork *c = ix->chain;
    chain_result(c, c->n_reads, c->n_chains, c->n_seeds, out);
    return MEME_OK;
}

// Chains the lists the last meme_seed_warmup left in HBM once and throws the result away: kernels loaded, scratch, output
// and pinned arrays sized for a chunk of that size.  A one-contig table stands in when none has been set yet.
extern "C" int meme_chain_warmup(meme_index_t *ix)
{
    if (!ix) { meme_set_error("bad argument"); return MEME_E_ARG; }
    if (!ix->ran || ix->piped || ix->last_opt.max_occ <= 0) return MEME_OK;      // nothing resident to warm up on
    int rc;
    ChainWork *c;
    if ((rc = chain_work(ix, &c))) return rc;
    const bool dummy = !c->d_contigs;
    if (dummy) {
        meme_contig_t one = {0, (int32_t)(ix->v.sa_num / 2 > 0x7fffffff ? 0x7fffffff : ix->v.sa_num / 2), 0};
        if ((rc = meme_chain_set_contigs(ix, &one, 1, (int64_t)(ix->v.sa_num / 2)))) return rc;
    }
    meme_chain_opt_t o = {100, 10000, ix->last_opt.max_occ, 19, 0, 1 << 30, 0.5f, 0.5f};
    if ((rc = meme_chain_run(ix, &o))) return rc;
    rc = chain_append(ix, 0, 0, 0);
    if (dummy) { cudaFree(c->d_contigs); c->d_contigs = nullptr; c->n_contigs = 0; }
    return rc;
}

// one device's share of a chunk: sub-batches of MEME_SEED_CHUNK reads, each seeded and chained in HBM; only chains travel
static int chain_batch_one(meme_index_t *ix, const uint8_t *codes, const int64_t *read_off, int64_t n, const meme_seed_opt_t *so,
                           const meme_chain_opt_t *co, meme_chain_result_t *out)
{
    const char *e = getenv("MEME_SEED_CHUNK");
    int64_t chunk = e ? atoll(e) : (1ll << 20);
    if (chunk < 1024) chunk = 1024;
    int rc;
    ChainWork *c;
    if ((rc = chain_work(ix, &c))) return rc;
    int64_t r0 = 0, c0 = 0, s0 = 0;
    std::vector<int64_t> off;
    if (n == 0) {
        if ((rc = chain_grow_pinned((void **)&c->h_chain_off, &c->hcap_chain_off, 1, 8, 0))) return rc;
        c->h_chain_off[0] = 0;
    }
    while (r0 < n) {
        int64_t m = n - r0 <= chunk + chunk / 2 ? n - r0 : chunk;
        const int64_t *o = read_off + r0;
        if (r0) {
            off.resize((size_t)m + 1);
            for (int64_t i = 0; i <= m; i++) off[(size_t)i] = read_off[r0 + i] - read_off[r0];
            o = off.data();
        }
        if ((rc = meme_seed_upload(ix, codes + read_off[r0], o, m))) return rc;
        if ((rc = meme_seed_run(ix, so))) return rc;
        if ((rc = meme_chain_run(ix, co))) return rc;
        if ((rc = chain_append(ix, r0, c0, s0))) return rc;
        r0 += m; c0 += c->n_chains; s0 += c->n_seeds;
    }
    chain_result(c, n, c0, s0, out);
    return MEME_OK;
}

extern "C" int meme_chain_batch_multi(meme_index_t *const *ixs, int ndev, const uint8_t *codes, const int64_t *read_off, int64_t n_reads,
                                      int pair_aligned, const meme_seed_opt_t *seed_opt, const meme_chain_opt_t *chain_opt,
                                      meme_chain_result_t *outs, int64_t *shard_begin)
{
    if (!ixs || ndev < 1 || !read_off || !seed_opt || !chain_opt || !outs || !shard_begin || n_reads < 0 || (n_reads > 0 && !codes)) {
        meme_set_error("bad argument");
        return MEME_E_ARG;
    }
    for (int d = 0; d <= ndev; d++) {
        int64_t b = n_reads * d / ndev;
        if (pair_aligned) b &= ~(int64_t)1;
        shard_begin[d] = d == ndev ? n_reads : b;
    }
    std::vector<int> rc((size_t)ndev, MEME_OK);
    std::vector<std::string> err((size_t)ndev);
    std::vector<std::thread> th;
    for (int d = 0; d < ndev; d++) {
        th.emplace_back([&, d] {
            const int64_t r0 = shard_begin[d], n = shard_begin[d + 1] - r0;
            std::vector<int64_t> o((size_t)n + 1);
            for (int64_t i = 0; i <= n; i++) o[(size_t)i] = read_off[r0 + i] - read_off[r0];
            rc[(size_t)d] = chain_batch_one(ixs[d], codes ? codes + read_off[r0] : nullptr, o.data(), n, seed_opt, chain_opt, &outs[d]);
            if (rc[(size_t)d]) err[(size_t)d] = meme_last_error();
        });
    }
    for (auto &t : th) t.join();
    for (int d = 0; d < ndev; d++)
        if (rc[(size_t)d]) { meme_set_error("shard %d: %s", d, err[(size_t)d].c_str()); return rc[(size_t)d]; }
    return MEME_OK;
}
