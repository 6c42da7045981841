// meme_seed.cu -- kernel A: P-RMI learned-index SMEM seeding on sm_100a, host side and helper kernels.
//
// Replaces, for a whole chunk of reads at once, the body of mem_kernel1_core_Learned up to the SMEM
// sort (reference src/bwamem.cpp:L1249-1397) = Learned_getSMEMsAllPosOneThread (+ reseeding) and
// Learned_bwtSeedStrategyAllPosOneThread{,_mem_tradeoff} (src/LearnedIndex_seeding.cpp:L913-1466),
// built on one search engine that stands for right_smem_search / mem_search / *_tradeoff (L2131-4189).
//
// The seeding kernel itself is in meme_seed_flat.cuh (one task per lane, flat state machine); this file holds
//   * k_pack_reads   2-bit packing of the reads (+ N mask)            (reference src/bwamem.cpp:L1277-1344)
//   * the launch sequence T1 (zig-zag) -> T2 (reseed) -> T3 (third round), pool sort (= ks_introsort, L1397),
//     offset scans and k_gather2 (SMEM records + hit lists, sampled as mem_chain_Learned reads them, L1154-1155)
//   * the C ABI entry points of include/meme_b200.h for kernel A, including the pipelined batch path.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>
#include <utility>
#include "meme_common.cuh"
#include "meme_seed_task.cuh"

namespace tpr {

// ---------------------------------------------------------------------------------------------
// 2-bit packing of the reads: one thread per (read, 32-base word)  (reference src/bwamem.cpp:L1277-1344)
__global__ void k_pack_reads(const uint8_t *__restrict__ codes, const int64_t *__restrict__ read_off, int64_t n_reads,
                             u64 *__restrict__ packed, u32 *__restrict__ nmask)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_reads * TPR_RW) return;
    int64_t r = t / TPR_RW;
    int w = (int)(t % TPR_RW);
    int64_t o = read_off[r];
    int L = (int)(read_off[r + 1] - o);
    u64 v = 0; u32 m = 0;
    for (int k = 0; k < 32; k++) {
        int p = w * 32 + k;
        if (p < L) {
            u32 c = codes[o + p];
            if (c > 3) m |= 1u << k; else v |= (u64)c << (2 * k);
        }
    }
    packed[t] = v;
    nmask[t] = m;
}

__global__ void k_counts2(const u32 *__restrict__ nsmem, const u64 *__restrict__ nhit, int64_t n,
                          int64_t *__restrict__ smem_off, int64_t *__restrict__ hit_off)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    smem_off[i] = i < n ? (int64_t)nsmem[i] : 0;
    hit_off[i] = i < n ? (int64_t)nhit[i] : 0;
}

__device__ __forceinline__ u64 sa_pos2(const IndexView &ix, u64 i)
{
    u64 p = __ldg(ix.pos_lo + i);
    if (ix.pos_wide) p |= (u64)__ldg(ix.pos_hi + i) << 32;
    return p;
}

// read-ordered SMEM records + hit lists from the sorted pool: 8 lanes per read
__global__ void k_gather2(IndexView ix, const u64 *__restrict__ keys, const PoolRec *__restrict__ vals,
                          const int64_t *__restrict__ smem_off, const int64_t *__restrict__ hit_off, int64_t n_reads,
                          int max_occ, meme_smem_t *__restrict__ smems, u64 *__restrict__ hits)
{
    const int gl = threadIdx.x & 7;
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (r >= n_reads) return;
    int64_t s0 = smem_off[r];
    int ns = (int)(smem_off[r + 1] - s0);
    u64 *hdst = hits + hit_off[r];
    int64_t hb = 0;
    for (int k = 0; k < ns; k++) {
        PoolRec a = vals[s0 + k];
        u64 key = keys[s0 + k];
        u64 c = a.cnt, step = 1, take = c;
        if (max_occ > 0) {
            step = c > (u64)max_occ ? c / max_occ : 1;
            take = (c + step - 1) / step;
            if (take > (u64)max_occ) take = (u64)max_occ;
        }
        for (u64 j = gl; j < take; j += 8) hdst[hb + j] = sa_pos2(ix, a.lo + j * step);
        if (gl == 0) {
            meme_smem_t m;
            m.start = (int32_t)((key >> 16) & 0xffff); m.end = (int32_t)(key & 0xffff);
            m.hitbeg = (int32_t)hb; m.hitcount = (int32_t)a.cnt;
            m.cache_refpos = sa_pos2(ix, a.lo);
            smems[s0 + k] = m;
        }
        hb += (int64_t)take;
    }
}

}  // namespace tpr

#include "meme_seed_flat.cuh"
#include "meme_seed_grp.cuh"

using namespace tpr;

template <class K>
static int launch_kernel(meme_index_t *ix, const Args &A, K kernel, size_t dyn_smem = 0)
{
    int sms = 0, per_sm = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ix->device));
    if (dyn_smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TPR_THREADS, dyn_smem));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)sms * per_sm, need = (A.n_tasks + TPR_THREADS - 1) / TPR_THREADS;
    if (grid > need) grid = need;
    CUDA_TRY(cudaMemsetAsync(ix->d_ctrl + C_WORK, 0, 8, ix->stream));
    kernel<<<(unsigned)grid, TPR_THREADS, dyn_smem, ix->stream>>>(ix->v, A);
    CUDA_TRY(cudaGetLastError());
    ix->launches++;
    return MEME_OK;
}

// MEME_SEED_IMPL: "flat" (default) = one task per lane (meme_seed_flat.cuh); "grp" = tasks regrouped by state across
// the CTA (meme_seed_grp.cuh): parity-green and 17-25 active lanes per instruction instead of 6-8, but 2.6x the warp
// instructions and 61 KB of code against a 32 KB instruction cache -- half the speed of the flat kernel (profiles/r2)
static int seed_impl()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("MEME_SEED_IMPL"); v = (e && !strcmp(e, "grp")) ? 1 : 0; }
    return v;
}

template <class K>
static int launch_grp_kernel(meme_index_t *ix, const Args &A, K kernel, size_t dyn_smem)
{
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ix->device));
    static int warps = getenv("MEME_GRP_WARPS") ? atoi(getenv("MEME_GRP_WARPS")) : 12;
    int nw = warps < 1 ? 1 : (warps > 32 ? 32 : warps);
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    int64_t grid = sms, need = (A.n_tasks + 63) / 64;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    CUDA_TRY(cudaMemsetAsync(ix->d_ctrl + C_WORK, 0, 8, ix->stream));
    kernel<<<(unsigned)grid, nw * 32, dyn_smem, ix->stream>>>(ix->v, A);
    CUDA_TRY(cudaGetLastError());
    ix->launches++;
    return MEME_OK;
}

template <int KIND, bool WIDE, bool LONG>
static int launch_grp(meme_index_t *ix, const Args &A, bool stats)
{
    const size_t sm = grp::Lay<WIDE, LONG>::BYTES;
    static_assert(grp::Lay<WIDE, LONG>::BYTES <= 227 * 1024, "task cells exceed the shared memory of an SM");
    return stats ? launch_grp_kernel(ix, A, grp::k_grp<KIND, WIDE, LONG, true>, sm) : launch_grp_kernel(ix, A, grp::k_grp<KIND, WIDE, LONG, false>, sm);
}

template <int KIND>
static int launch_seed(meme_index_t *ix, const Args &A)
{
    if (A.n_tasks <= 0) return MEME_OK;
    // work counters (meme_seed_counters) cost registers and code in the hot loop: off unless asked for
    const char *se = getenv("MEME_SEED_STATS");
    const bool stats = se && atoi(se) != 0;
    if (seed_impl() == 1) {
        const bool lng = ix->max_read_len > 32 * (TPR_RW / 2);
        if (ix->v.pos_wide) return lng ? launch_grp<KIND, true, true>(ix, A, stats) : launch_grp<KIND, true, false>(ix, A, stats);
        return lng ? launch_grp<KIND, false, true>(ix, A, stats) : launch_grp<KIND, false, false>(ix, A, stats);
    }
    const size_t sm = (size_t)TPR_THREADS * flat_sw<KIND>() * sizeof(u32);
    if (ix->v.pos_wide)
        return stats ? launch_kernel(ix, A, k_flat<KIND, true, 2, true, true>, sm) : launch_kernel(ix, A, k_flat<KIND, true, 2, true, false>, sm);
    return stats ? launch_kernel(ix, A, k_flat<KIND, false, FLAT_MINB, true, true>, sm) : launch_kernel(ix, A, k_flat<KIND, false, FLAT_MINB, true, false>, sm);
}

static int seed_run_kernels(meme_index_t *ix, const meme_seed_opt_t *opt)
{
    cudaStream_t st = ix->stream;
    const int64_t n = ix->n_reads;
    int rc;
    // ---- workspace ----
    if ((rc = meme_grow_device((void **)&ix->d_packed, &ix->cap_packed, (size_t)n * TPR_RW + 16, 8, st))) return rc;
    if ((rc = meme_grow_device((void **)&ix->d_nmask, &ix->cap_nmask, (size_t)n * TPR_RW + 16, 4, st))) return rc;
    if ((rc = meme_grow_device((void **)&ix->d_rd_full, &ix->cap_rd_full, (size_t)n + 16, 1, st))) return rc;
    if ((rc = meme_grow_device((void **)&ix->d_rd_maxref, &ix->cap_rd_maxref, (size_t)n + 16, 8, st))) return rc;
    if (!ix->d_pool_keys) {
        size_t want = (size_t)n * 14 + 65536;
        if ((rc = meme_grow_device((void **)&ix->d_pool_keys, &ix->cap_pool_keys, want, 8, st))) return rc;
    }
    if (!ix->d_t2) {
        size_t want = (size_t)n * 3 + 65536;
        if ((rc = meme_grow_device((void **)&ix->d_t2, &ix->cap_t2, want, sizeof(Task2), st))) return rc;
    }
    CUDA_TRY(cudaEventRecord(ix->ev[0], st));
    k_pack_reads<<<(unsigned)((n * TPR_RW + 255) / 256), 256, 0, st>>>(ix->d_codes, ix->d_read_off, n, ix->d_packed, ix->d_nmask);
    ix->launches++;
    u64 n_pool = 0, n_t2 = 0;
    for (int attempt = 0;; attempt++) {
        // pool_vals / sorted copies follow the key capacity
        size_t cap = ix->cap_pool_keys;
        if ((rc = meme_grow_device((void **)&ix->d_pool_vals, &ix->cap_pool_vals, cap, sizeof(PoolRec), st))) return rc;
        if ((rc = meme_grow_device((void **)&ix->d_pool_keys2, &ix->cap_pool_keys2, cap, 8, st))) return rc;
        if ((rc = meme_grow_device((void **)&ix->d_pool_vals2, &ix->cap_pool_vals2, cap, sizeof(PoolRec), st))) return rc;
        CUDA_TRY(cudaMemsetAsync(ix->d_ctrl, 0, C_NCTRL * 8, st));
        CUDA_TRY(cudaMemsetAsync(ix->d_rd_nsmem, 0, (size_t)n * 4, st));
        CUDA_TRY(cudaMemsetAsync(ix->d_rd_nhit, 0, (size_t)n * 8, st));
        Args A;
        A.packed = ix->d_packed; A.nmask = ix->d_nmask; A.read_off = ix->d_read_off;
        A.t2_in = nullptr; A.t2_out = (Task2 *)ix->d_t2; A.t2_cap = ix->cap_t2;
        A.pool_keys = ix->d_pool_keys; A.pool_vals = (PoolRec *)ix->d_pool_vals;
        A.pool_cap = ix->cap_pool_keys < ix->cap_pool_vals ? ix->cap_pool_keys : ix->cap_pool_vals;
        A.rd_nsmem = ix->d_rd_nsmem; A.rd_nhit = ix->d_rd_nhit; A.rd_full = ix->d_rd_full; A.rd_maxref = ix->d_rd_maxref;
        A.ctrl = ix->d_ctrl; A.opt = *opt;
        {
            static int th = getenv("MEME_GRP_TH") ? atoi(getenv("MEME_GRP_TH")) : 24;
            A.grp_th = th;
        }
        A.kind = 1; A.n_tasks = n;
        if ((rc = launch_seed<1>(ix, A))) return rc;
        CUDA_TRY(cudaMemcpyAsync(ix->h_ctrl, ix->d_ctrl, C_NCTRL * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        n_t2 = ix->h_ctrl[C_T2];
        bool ovf = ix->h_ctrl[C_T2_OVF] != 0;
        static const bool dbg = getenv("MEME_SEED_DEBUG") != nullptr;   // per-kernel counters + times on stderr
        auto dump = [&](const char *name, cudaEvent_t e0, cudaEvent_t e1, int64_t tasks) {
            u64 c[C_NCTRL]; float ms = 0;
            cudaMemcpyAsync(c, ix->d_ctrl, sizeof c, cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            cudaEventElapsedTime(&ms, e0, e1);
            fprintf(stderr, "[meme_seed] %s: %lld tasks, %.2f ms (cumulative counters) searches %llu blocks %llu kary %llu ext %llu pac %llu "
                    "round-trips %llu; passes ctl %llu (%.1f lanes) mem %llu (%.1f lanes)\n", name, (long long)tasks, ms, c[C_SEARCH],
                    c[C_BLOCK], c[C_KARY], c[C_EXT], c[C_PAC], c[C_ITER], c[C_PASS_CTL] / 32,
                    c[C_PASS_CTL] ? 32.0 * c[C_LANES_CTL] / c[C_PASS_CTL] : 0.0, c[C_PASS_MEM] / 32,
                    c[C_PASS_MEM] ? 32.0 * c[C_LANES_MEM] / c[C_PASS_MEM] : 0.0);
        };
        cudaEvent_t de[4];
        if (dbg) { for (auto &e : de) cudaEventCreate(&e); cudaEventRecord(de[1], st); dump("T1", ix->ev[0], de[1], n); }
        if (!ovf && opt->rounds >= 2 && n_t2) {
            A.kind = 2; A.n_tasks = (int64_t)n_t2; A.t2_in = (const Task2 *)ix->d_t2; A.t2_out = nullptr; A.t2_cap = 0;
            if ((rc = launch_seed<2>(ix, A))) return rc;
            if (dbg) { cudaEventRecord(de[2], st); dump("T2", de[1], de[2], (int64_t)n_t2); }
        }
        if (!ovf && opt->rounds >= 3 && opt->max_mem_intv > 0) {
            A.kind = 3; A.n_tasks = n; A.t2_in = nullptr;
            if ((rc = launch_seed<3>(ix, A))) return rc;
            if (dbg) { cudaEventRecord(de[3], st); dump("T3", de[2], de[3], n); }
        }
        if (dbg) for (auto &e : de) cudaEventDestroy(e);
        CUDA_TRY(cudaEventRecord(ix->ev[1], st));
        CUDA_TRY(cudaMemcpyAsync(ix->h_ctrl, ix->d_ctrl, C_NCTRL * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        n_pool = ix->h_ctrl[C_POOL];
        if (!ovf && !ix->h_ctrl[C_POOL_OVF]) break;
        if (attempt >= 2) { meme_set_error("seed pools keep overflowing"); return MEME_E_NOMEM; }
        if (ovf && (rc = meme_grow_device((void **)&ix->d_t2, &ix->cap_t2, (size_t)n_t2 * 2 + 65536, sizeof(Task2), st))) return rc;
        if (ix->h_ctrl[C_POOL_OVF] &&
            (rc = meme_grow_device((void **)&ix->d_pool_keys, &ix->cap_pool_keys, (size_t)n_pool * 2 + 65536, 8, st))) return rc;
    }
    // ---- sort by (read, start, end) ----
    CUDA_TRY(cudaEventRecord(ix->ev[2], st));
    {
        int bits = 32;
        while (bits < 64 && (1ull << (bits - 32)) < (u64)n) bits++;
        size_t tmp = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp, ix->d_pool_keys, ix->d_pool_keys2, (PoolRec *)ix->d_pool_vals,
                                        (PoolRec *)ix->d_pool_vals2, (int64_t)n_pool, 0, 64, st);
        if ((rc = meme_grow_device(&ix->d_scan_tmp, &ix->cap_scan_tmp, tmp + 256, 1, st))) return rc;
        size_t t2 = ix->cap_scan_tmp;
        if (n_pool)
            CUDA_TRY(cub::DeviceRadixSort::SortPairs(ix->d_scan_tmp, t2, ix->d_pool_keys, ix->d_pool_keys2, (PoolRec *)ix->d_pool_vals,
                                                     (PoolRec *)ix->d_pool_vals2, (int64_t)n_pool, 0, 64, st));
        (void)bits;
    }
    k_counts2<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(ix->d_rd_nsmem, ix->d_rd_nhit, n, ix->d_smem_off, ix->d_hit_off);
    {
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, ix->d_smem_off, ix->d_smem_off, (int)(n + 1), st);
        if ((rc = meme_grow_device(&ix->d_scan_tmp, &ix->cap_scan_tmp, tmp + 256, 1, st))) return rc;
        size_t t2 = ix->cap_scan_tmp;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(ix->d_scan_tmp, t2, ix->d_smem_off, ix->d_smem_off, (int)(n + 1), st));
        t2 = ix->cap_scan_tmp;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(ix->d_scan_tmp, t2, ix->d_hit_off, ix->d_hit_off, (int)(n + 1), st));
    }
    CUDA_TRY(cudaEventRecord(ix->ev[3], st));
    ix->launches += 4;
    int64_t tot[2];
    CUDA_TRY(cudaMemcpyAsync(&tot[0], ix->d_smem_off + n, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&tot[1], ix->d_hit_off + n, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if ((rc = meme_grow_device((void **)&ix->d_smems, &ix->cap_smems, (size_t)tot[0] + 1, sizeof(meme_smem_t), st))) return rc;
    if ((rc = meme_grow_device((void **)&ix->d_hits, &ix->cap_hits, (size_t)tot[1] + 1, 8, st))) return rc;
    CUDA_TRY(cudaEventRecord(ix->ev[4], st));
    k_gather2<<<(unsigned)((n * 8 + 255) / 256), 256, 0, st>>>(ix->v, ix->d_pool_keys2, (const PoolRec *)ix->d_pool_vals2, ix->d_smem_off,
                                                             ix->d_hit_off, n, opt->max_occ, ix->d_smems, ix->d_hits);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ix->ev[5], st));
    ix->ev_ready = 1;
    ix->launches++;
    ix->h_ctrl[32] = (u64)tot[0];
    ix->h_ctrl[33] = (u64)tot[1];
    // statistics in the slots meme_seed_counters reads
    ix->h_ctrl[8] = ix->h_ctrl[C_SEARCH]; ix->h_ctrl[9] = ix->h_ctrl[C_BLOCK]; ix->h_ctrl[10] = ix->h_ctrl[C_KARY];
    ix->h_ctrl[11] = ix->h_ctrl[C_EXT];
    ix->h_ctrl[2] = 0;                 // no per-read SMEM limit in this path
    ix->h_ctrl[4] = n_t2;              // reseed tasks (reported in the "redo" counter slot)
    ix->h_ctrl[40] = ix->h_ctrl[C_PAC]; ix->h_ctrl[41] = ix->h_ctrl[C_ITER];
    ix->ran = 1;
    return MEME_OK;
}

// ---------------------------------------------------------------------------------------------
// host API
// ---------------------------------------------------------------------------------------------
// input buffers and the per-read arrays of a batch of up to n_reads reads / n_codes bases
static int seed_alloc_inputs(meme_index_t *ix, int64_t n_reads, int64_t n_codes)
{
    int rc;
    cudaStream_t st = ix->stream;
    if ((rc = meme_grow_device((void **)&ix->d_codes, &ix->cap_codes, (size_t)n_codes + 64, 1, st))) return rc;
    size_t old_cap = ix->cap_reads;
    if ((rc = meme_grow_device((void **)&ix->d_read_off, &ix->cap_reads, (size_t)n_reads + 1, 8, st))) return rc;
    if (ix->cap_reads != old_cap || !ix->d_rd_nsmem) {
        size_t c = ix->cap_reads, dummy;
        void **bufs[] = {(void **)&ix->d_rd_nsmem, (void **)&ix->d_rd_nhit, (void **)&ix->d_smem_off, (void **)&ix->d_hit_off};
        size_t el[] = {4, 8, 8, 8};
        for (int k = 0; k < 4; k++) {
            if (*bufs[k]) { CUDA_TRY(cudaFree(*bufs[k])); *bufs[k] = nullptr; }
            dummy = 0;
            if ((rc = meme_grow_device(bufs[k], &dummy, c, el[k], st))) return rc;
        }
    }
    return MEME_OK;
}

extern "C" int meme_seed_upload(meme_index_t *ix, const uint8_t *codes, const int64_t *read_off, int64_t n_reads)
{
    if (!ix || !read_off || n_reads < 0 || (n_reads > 0 && !codes)) { meme_set_error("bad argument"); return MEME_E_ARG; }
    CUDA_TRY(cudaSetDevice(ix->device));
    int64_t n_codes = n_reads ? read_off[n_reads] - read_off[0] : 0;
    if (n_reads && read_off[0] != 0) { meme_set_error("read_off[0] must be 0"); return MEME_E_ARG; }
    int64_t max_len = 0;
    for (int64_t i = 0; i < n_reads; i++) {
        int64_t len = read_off[i + 1] - read_off[i];
        if (len < 0 || len > MEME_MAX_READ_LEN) {
            meme_set_error("read %lld has length %lld (limit %d; the reference asserts <= 301)", (long long)i, (long long)len, MEME_MAX_READ_LEN);
            return MEME_E_ARG;
        }
        if (len > max_len) max_len = len;
    }
    ix->max_read_len = (int)max_len;
    int rc;
    cudaStream_t st = ix->stream;
    if ((rc = seed_alloc_inputs(ix, n_reads, n_codes))) return rc;
    ix->piped = 0;
    if (n_codes) CUDA_TRY(cudaMemcpyAsync(ix->d_codes, codes, (size_t)n_codes, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ix->d_read_off, read_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    ix->n_reads = n_reads;
    ix->n_codes = n_codes;
    ix->ran = 0;
    return MEME_OK;
}

extern "C" int meme_seed_run(meme_index_t *ix, const meme_seed_opt_t *opt)
{
    if (!ix || !opt) { meme_set_error("null argument"); return MEME_E_ARG; }
    if (opt->min_seed_len < 1 || opt->rounds < 1) { meme_set_error("bad seeding options"); return MEME_E_ARG; }
    CUDA_TRY(cudaSetDevice(ix->device));
    ix->last_opt = *opt;
    if (ix->n_reads == 0) { ix->ran = 1; return MEME_OK; }
    return seed_run_kernels(ix, opt);
}

extern "C" int meme_seed_totals(meme_index_t *ix, int64_t *n_smems, int64_t *n_hits)
{
    if (!ix || !ix->ran) { meme_set_error("no completed seeding run"); return MEME_E_ARG; }
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    if (n_smems) *n_smems = ix->n_reads ? (int64_t)ix->h_ctrl[32] : 0;
    if (n_hits) *n_hits = ix->n_reads ? (int64_t)ix->h_ctrl[33] : 0;
    return MEME_OK;
}

// Device-to-device copies of the lists of the last meme_seed_run (any pointer may be NULL): for callers that keep the
// results in HBM, e.g. the NCCL gather of the multi-process bench.  Asynchronous on the handle's stream.
extern "C" int meme_seed_copy_device(meme_index_t *ix, void *d_smems, void *d_hits, void *d_smem_off, void *d_hit_off)
{
    if (!ix || !ix->ran || ix->piped) { meme_set_error("no completed meme_seed_run on this handle"); return MEME_E_ARG; }
    CUDA_TRY(cudaSetDevice(ix->device));
    const int64_t n = ix->n_reads, ns = n ? (int64_t)ix->h_ctrl[32] : 0, nh = n ? (int64_t)ix->h_ctrl[33] : 0;
    cudaStream_t st = ix->stream;
    if (d_smems && ns) CUDA_TRY(cudaMemcpyAsync(d_smems, ix->d_smems, (size_t)ns * sizeof(meme_smem_t), cudaMemcpyDeviceToDevice, st));
    if (d_hits && nh) CUDA_TRY(cudaMemcpyAsync(d_hits, ix->d_hits, (size_t)nh * 8, cudaMemcpyDeviceToDevice, st));
    if (d_smem_off && n) CUDA_TRY(cudaMemcpyAsync(d_smem_off, ix->d_smem_off, (size_t)(n + 1) * 8, cudaMemcpyDeviceToDevice, st));
    if (d_hit_off && n) CUDA_TRY(cudaMemcpyAsync(d_hit_off, ix->d_hit_off, (size_t)(n + 1) * 8, cudaMemcpyDeviceToDevice, st));
    return MEME_OK;
}

extern "C" int meme_seed_last_ms(meme_index_t *ix, float ms[3])
{
    if (!ix || !ms || !ix->ev_ready) { meme_set_error("no timed seeding run"); return MEME_E_ARG; }
    CUDA_TRY(cudaEventSynchronize(ix->ev[5]));
    CUDA_TRY(cudaEventElapsedTime(&ms[0], ix->ev[0], ix->ev[1]));
    CUDA_TRY(cudaEventElapsedTime(&ms[1], ix->ev[2], ix->ev[3]));
    CUDA_TRY(cudaEventElapsedTime(&ms[2], ix->ev[4], ix->ev[5]));
    return MEME_OK;
}

extern "C" int meme_seed_counters(meme_index_t *ix, uint64_t counters[8])
{
    if (!ix || !counters) return MEME_E_ARG;
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    counters[0] = ix->h_ctrl[8]; counters[1] = ix->h_ctrl[9]; counters[2] = ix->h_ctrl[10]; counters[3] = ix->h_ctrl[11];
    counters[4] = ix->h_ctrl[40]; counters[5] = ix->launches; counters[6] = ix->h_ctrl[4]; counters[7] = ix->h_ctrl[41];
    return MEME_OK;
}

extern "C" int meme_seed_download(meme_index_t *ix, meme_seed_result_t *out)
{
    if (!ix || !out || !ix->ran) { meme_set_error("no completed seeding run"); return MEME_E_ARG; }
    CUDA_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    int64_t n = ix->n_reads;
    int64_t ns = n ? (int64_t)ix->h_ctrl[32] : 0, nh = n ? (int64_t)ix->h_ctrl[33] : 0;
    int rc;
    if (ix->piped) {                       // the pipelined batch already assembled the lists on the host
        out->n_reads = ix->piped_reads; out->n_smems = ns; out->n_hits = nh;
        out->smems = ix->h_smems; out->smem_off = ix->h_smem_off; out->hit_off = ix->h_hit_off; out->hits = (const uint64_t *)ix->h_hits;
        return MEME_OK;
    }
    if ((rc = meme_grow_pinned((void **)&ix->h_smems, &ix->hcap_smems, (size_t)ns + 1, sizeof(meme_smem_t)))) return rc;
    if ((rc = meme_grow_pinned((void **)&ix->h_hits, &ix->hcap_hits, (size_t)nh + 1, 8))) return rc;
    size_t c0 = ix->hcap_off;
    if ((rc = meme_grow_pinned((void **)&ix->h_smem_off, &c0, (size_t)n + 1, 8))) return rc;
    c0 = ix->hcap_off;
    if ((rc = meme_grow_pinned((void **)&ix->h_hit_off, &c0, (size_t)n + 1, 8))) return rc;
    ix->hcap_off = c0;
    if (n) {
        if (ns) CUDA_TRY(cudaMemcpyAsync(ix->h_smems, ix->d_smems, (size_t)ns * sizeof(meme_smem_t), cudaMemcpyDeviceToHost, st));
        if (nh) CUDA_TRY(cudaMemcpyAsync(ix->h_hits, ix->d_hits, (size_t)nh * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(ix->h_smem_off, ix->d_smem_off, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(ix->h_hit_off, ix->d_hit_off, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    } else {
        ix->h_smem_off[0] = 0; ix->h_hit_off[0] = 0;
    }
    out->n_reads = n; out->n_smems = ns; out->n_hits = nh;
    out->smems = ix->h_smems; out->smem_off = ix->h_smem_off; out->hit_off = ix->h_hit_off; out->hits = (const uint64_t *)ix->h_hits;
    return MEME_OK;
}

// ---------------------------------------------------------------------------------------------
// Pipelined batch: the chunk (the reference's -K batch, src/fastmap.cpp:L743) is cut into sub-batches of
// MEME_SEED_CHUNK reads; while sub-batch c runs, the lists of c-1 drain to the host and the reads of c+1 arrive,
// on a second stream with a second set of input/output buffers.  Results land in one set of pinned arrays,
// exactly as the one-shot path delivers them.
// ---------------------------------------------------------------------------------------------
__global__ void k_add_base(int64_t *__restrict__ a, int64_t n, int64_t base)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += base;
}

static void seed_swap_sets(meme_index_t *ix)
{
    SeedAlt &a = ix->alt;
    std::swap(ix->d_codes, a.d_codes); std::swap(ix->cap_codes, a.cap_codes);
    std::swap(ix->d_read_off, a.d_read_off); std::swap(ix->d_smem_off, a.d_smem_off); std::swap(ix->d_hit_off, a.d_hit_off);
    std::swap(ix->d_smems, a.d_smems); std::swap(ix->cap_smems, a.cap_smems);
    std::swap(ix->d_hits, a.d_hits); std::swap(ix->cap_hits, a.cap_hits);
}

// grow a pinned array, keeping its first `keep` elements
static int grow_pinned_keep(void **p, size_t *cap, size_t need, size_t elem, size_t keep)
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

static int64_t seed_chunk_reads()
{
    const char *e = getenv("MEME_SEED_CHUNK");
    long long v = e ? atoll(e) : (1ll << 20);
    return v < 1024 ? 1024 : (int64_t)v;
}

static int seed_batch_pipelined(meme_index_t *ix, const uint8_t *codes, const int64_t *read_off, int64_t n,
                                const meme_seed_opt_t *opt, meme_seed_result_t *out, int64_t chunk)
{
    CUDA_TRY(cudaSetDevice(ix->device));
    if (read_off[0] != 0) { meme_set_error("read_off[0] must be 0"); return MEME_E_ARG; }
    int64_t max_len = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t len = read_off[i + 1] - read_off[i];
        if (len < 0 || len > MEME_MAX_READ_LEN) {
            meme_set_error("read %lld has length %lld (limit %d; the reference asserts <= 301)", (long long)i, (long long)len, MEME_MAX_READ_LEN);
            return MEME_E_ARG;
        }
        if (len > max_len) max_len = len;
    }
    ix->max_read_len = (int)max_len;
    const int64_t K = (n + chunk - 1) / chunk;
    int64_t max_codes = 0;
    for (int64_t c = 0; c < K; c++) {
        int64_t r0 = c * chunk, r1 = r0 + chunk < n ? r0 + chunk : n;
        max_codes = std::max(max_codes, read_off[r1] - read_off[r0]);
    }
    int rc;
    cudaStream_t st = ix->stream;
    if ((rc = seed_alloc_inputs(ix, chunk, max_codes))) return rc;
    SeedAlt &A = ix->alt;
    if (A.cap_reads != ix->cap_reads) {                 // second set of per-read offset arrays, same capacity as the first
        CUDA_TRY(cudaDeviceSynchronize());
        void **bufs[] = {(void **)&A.d_read_off, (void **)&A.d_smem_off, (void **)&A.d_hit_off};
        for (void **b : bufs) {
            if (*b) { CUDA_TRY(cudaFree(*b)); *b = nullptr; }
            size_t dummy = 0;
            if ((rc = meme_grow_device(b, &dummy, ix->cap_reads, 8, st))) return rc;
        }
        A.cap_reads = ix->cap_reads;
    }
    if ((rc = meme_grow_device((void **)&A.d_codes, &A.cap_codes, (size_t)max_codes + 64, 1, st))) return rc;
    if (!ix->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ix->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 6; k++) CUDA_TRY(cudaEventCreateWithFlags(&ix->pe[k], cudaEventDisableTiming));
    }
    cudaStream_t cs = ix->copy_stream;
    cudaEvent_t *e_h2d = ix->pe, *e_run = ix->pe + 2, *e_d2h = ix->pe + 4;      // indexed by set (chunk & 1)
    {
        size_t c0 = ix->hcap_off;
        if ((rc = meme_grow_pinned((void **)&ix->h_smem_off, &c0, (size_t)n + 1, 8))) return rc;
        c0 = ix->hcap_off;
        if ((rc = meme_grow_pinned((void **)&ix->h_hit_off, &c0, (size_t)n + 1, 8))) return rc;
        ix->hcap_off = c0;
    }
    struct Part { int64_t r0, n, tot_s, tot_h, base_s, base_h; } part[2] = {};
    int64_t base_s = 0, base_h = 0;
    u64 acc[7] = {0, 0, 0, 0, 0, 0, 0};
    // input of sub-batch c -> the buffers named by the arguments (the set that is NOT installed while c-1 runs)
    auto h2d = [&](int64_t c, uint8_t *d_codes, int64_t *d_read_off) -> int {
        int64_t r0 = c * chunk, r1 = r0 + chunk < n ? r0 + chunk : n, nr = r1 - r0;
        int64_t nc = read_off[r1] - read_off[r0];
        if (nc) CUDA_TRY(cudaMemcpyAsync(d_codes, codes + read_off[r0], (size_t)nc, cudaMemcpyHostToDevice, cs));
        CUDA_TRY(cudaMemcpyAsync(d_read_off, read_off + r0, (size_t)(nr + 1) * 8, cudaMemcpyHostToDevice, cs));
        if (read_off[r0]) k_add_base<<<(unsigned)((nr + 1 + 255) / 256), 256, 0, cs>>>(d_read_off, nr + 1, -read_off[r0]);
        CUDA_TRY(cudaEventRecord(e_h2d[c & 1], cs));
        return MEME_OK;
    };
    // lists of a finished sub-batch -> the pinned arrays, offsets rebased to the whole batch
    auto d2h = [&](const Part &p, int set, const meme_smem_t *d_smems, const u64 *d_hits, int64_t *d_smem_off, int64_t *d_hit_off) -> int {
        int r;
        if ((size_t)(p.base_s + p.tot_s) + 1 > ix->hcap_smems || (size_t)(p.base_h + p.tot_h) + 1 > ix->hcap_hits || !ix->h_smems || !ix->h_hits) {
            CUDA_TRY(cudaStreamSynchronize(cs));        // earlier copies target the old arrays
            // first call: size for the whole batch from the density seen so far
            double scale = (double)n / (double)(p.r0 + p.n) * 1.1;
            size_t want_s = (size_t)((double)(p.base_s + p.tot_s) * scale) + 1024, want_h = (size_t)((double)(p.base_h + p.tot_h) * scale) + 1024;
            if ((r = grow_pinned_keep((void **)&ix->h_smems, &ix->hcap_smems, want_s, sizeof(meme_smem_t), (size_t)p.base_s))) return r;
            if ((r = grow_pinned_keep((void **)&ix->h_hits, &ix->hcap_hits, want_h, 8, (size_t)p.base_h))) return r;
        }
        CUDA_TRY(cudaStreamWaitEvent(cs, e_run[set], 0));
        unsigned g = (unsigned)((p.n + 1 + 255) / 256);
        if (p.base_s) k_add_base<<<g, 256, 0, cs>>>(d_smem_off, p.n + 1, p.base_s);
        if (p.base_h) k_add_base<<<g, 256, 0, cs>>>(d_hit_off, p.n + 1, p.base_h);
        if (p.tot_s) CUDA_TRY(cudaMemcpyAsync(ix->h_smems + p.base_s, d_smems, (size_t)p.tot_s * sizeof(meme_smem_t), cudaMemcpyDeviceToHost, cs));
        if (p.tot_h) CUDA_TRY(cudaMemcpyAsync(ix->h_hits + p.base_h, d_hits, (size_t)p.tot_h * 8, cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaMemcpyAsync(ix->h_smem_off + p.r0, d_smem_off, (size_t)(p.n + 1) * 8, cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaMemcpyAsync(ix->h_hit_off + p.r0, d_hit_off, (size_t)(p.n + 1) * 8, cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaEventRecord(e_d2h[set], cs));
        return MEME_OK;
    };
    CUDA_TRY(cudaStreamSynchronize(st));                // the handle's buffers are idle from here on
    if ((rc = h2d(0, ix->d_codes, ix->d_read_off))) return rc;
    for (int64_t c = 0; c < K; c++) {
        const int set = (int)(c & 1);
        if (c >= 1) {
            seed_swap_sets(ix);                          // installed: set of c; ix->alt: set of c-1 (= set of c+1)
            if ((rc = d2h(part[set ^ 1], set ^ 1, A.d_smems, A.d_hits, A.d_smem_off, A.d_hit_off))) return rc;
        }
        if (c + 1 < K && (rc = h2d(c + 1, A.d_codes, A.d_read_off))) return rc;   // after the d2h wait on e_run[c-1]
        int64_t r0 = c * chunk, r1 = r0 + chunk < n ? r0 + chunk : n;
        CUDA_TRY(cudaStreamWaitEvent(st, e_h2d[set], 0));
        if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(st, e_d2h[set], 0));             // the lists of c-2 have left this set
        ix->n_reads = r1 - r0; ix->n_codes = read_off[r1] - read_off[r0]; ix->ran = 0;
        if ((rc = meme_seed_run(ix, opt))) return rc;
        CUDA_TRY(cudaEventRecord(e_run[set], st));
        Part &p = part[set];
        p.r0 = r0; p.n = r1 - r0; p.tot_s = (int64_t)ix->h_ctrl[32]; p.tot_h = (int64_t)ix->h_ctrl[33];
        p.base_s = base_s; p.base_h = base_h;
        base_s += p.tot_s; base_h += p.tot_h;
        const int slot[7] = {8, 9, 10, 11, 4, 40, 41};
        for (int k = 0; k < 7; k++) acc[k] += ix->h_ctrl[slot[k]];
    }
    {
        const int set = (int)((K - 1) & 1);
        if ((rc = d2h(part[set], set, ix->d_smems, ix->d_hits, ix->d_smem_off, ix->d_hit_off))) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(cs));
    CUDA_TRY(cudaStreamSynchronize(st));
    {
        const int slot[7] = {8, 9, 10, 11, 4, 40, 41};
        for (int k = 0; k < 7; k++) ix->h_ctrl[slot[k]] = acc[k];
    }
    ix->h_ctrl[32] = (u64)base_s; ix->h_ctrl[33] = (u64)base_h;
    ix->piped = 1; ix->piped_reads = n; ix->ran = 1;
    out->n_reads = n; out->n_smems = base_s; out->n_hits = base_h;
    out->smems = ix->h_smems; out->smem_off = ix->h_smem_off; out->hit_off = ix->h_hit_off; out->hits = (const uint64_t *)ix->h_hits;
    return MEME_OK;
}

// ---------------------------------------------------------------------------------------------
// Warm-up: one throw-away batch of reads sampled from the indexed text itself (1 % pseudo-random substitutions), so that
// module loading and every pool the batch path grows on demand (device pools, pinned result arrays) are paid where the
// caller loads the index, not inside its first timed chunk.
// ---------------------------------------------------------------------------------------------
extern "C" int meme_seed_batch(meme_index_t *ix, const uint8_t *codes, const int64_t *read_off, int64_t n_reads,
                               const meme_seed_opt_t *opt, meme_seed_result_t *out);

__global__ void k_warm_reads(IndexView ix, int64_t n_reads, int read_len, uint8_t *__restrict__ codes)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_reads * read_len) return;
    const int64_t r = t / read_len;
    const int k = (int)(t % read_len);
    const u64 span = ix.sa_num > (u64)read_len + 1 ? ix.sa_num - (u64)read_len - 1 : 1;
    const u64 p = ((u64)r * 0x9E3779B97F4A7C15ull >> 11) % span + (u64)k;
    u32 c = (u32)(ix.pac[p >> 5] >> (2 * (p & 31))) & 3;
    u64 h = (u64)t * 0xD6E8FEB86659FD93ull; h ^= h >> 32; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32;
    if (h % 100 == 0) c = (c + 1 + (u32)(h >> 40) % 3) & 3;
    codes[t] = (uint8_t)c;
}

extern "C" int meme_seed_warmup(meme_index_t *ix, int64_t n_reads, int32_t read_len)
{
    if (!ix || n_reads <= 0 || read_len <= 0 || read_len > MEME_MAX_READ_LEN) { meme_set_error("bad argument"); return MEME_E_ARG; }
    CUDA_TRY(cudaSetDevice(ix->device));
    const size_t nb = (size_t)n_reads * (size_t)read_len;
    uint8_t *d = nullptr, *h = nullptr;
    int64_t *off = nullptr;
    CUDA_TRY(cudaMalloc((void **)&d, nb));
    CUDA_TRY(cudaMallocHost((void **)&h, nb));
    CUDA_TRY(cudaMallocHost((void **)&off, (size_t)(n_reads + 1) * 8));
    k_warm_reads<<<(unsigned)((nb + 255) / 256), 256, 0, ix->stream>>>(ix->v, n_reads, read_len, d);
    CUDA_TRY(cudaMemcpyAsync(h, d, nb, cudaMemcpyDeviceToHost, ix->stream));
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    for (int64_t i = 0; i <= n_reads; i++) off[i] = i * read_len;
    meme_seed_opt_t opt = {19, 28, 10, 20, 3, 500};
    meme_seed_result_t res;
    int rc = meme_seed_batch(ix, h, off, n_reads, &opt, &res);
    cudaFree(d); cudaFreeHost(h); cudaFreeHost(off);
    return rc;
}

extern "C" int meme_seed_batch(meme_index_t *ix, const uint8_t *codes, const int64_t *read_off, int64_t n_reads,
                               const meme_seed_opt_t *opt, meme_seed_result_t *out)
{
    int rc;
    if (!ix || !out || !opt || !read_off || n_reads < 0 || (n_reads > 0 && !codes)) { meme_set_error("bad argument"); return MEME_E_ARG; }
    const int64_t chunk = seed_chunk_reads();
    if (n_reads > chunk + chunk / 2) return seed_batch_pipelined(ix, codes, read_off, n_reads, opt, out, chunk);
    if ((rc = meme_seed_upload(ix, codes, read_off, n_reads))) return rc;
    if ((rc = meme_seed_run(ix, opt))) return rc;
    return meme_seed_download(ix, out);
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU: reads are independent units, so a chunk is cut into `ndev` contiguous read ranges (pair-aligned when
// `pair_aligned` != 0: both mates of a pair stay on one device) and every range runs through meme_seed_batch on its own
// device, concurrently, one host thread per device.  The index is replicated (meme_index_load_devices).  There is no
// device-to-device exchange on this path: each device delivers its lists into its own pinned host arrays, and
// shard_begin[d] .. shard_begin[d+1] tells the caller which reads outs[d] describes.
// ---------------------------------------------------------------------------------------------
extern "C" int meme_seed_batch_multi(meme_index_t *const *ixs, int ndev, const uint8_t *codes, const int64_t *read_off, int64_t n_reads,
                                     int pair_aligned, const meme_seed_opt_t *opt, meme_seed_result_t *outs, int64_t *shard_begin)
{
    if (!ixs || ndev < 1 || !read_off || !opt || !outs || !shard_begin || n_reads < 0) { meme_set_error("bad argument"); return MEME_E_ARG; }
    for (int d = 0; d <= ndev; d++) {
        int64_t b = n_reads * d / ndev;
        if (pair_aligned) b &= ~(int64_t)1;
        shard_begin[d] = d == ndev ? n_reads : b;
    }
    std::vector<int> rc((size_t)ndev, MEME_OK);
    std::vector<std::string> err((size_t)ndev);
    std::vector<std::vector<int64_t>> offs((size_t)ndev);
    std::vector<std::thread> th;
    for (int d = 0; d < ndev; d++) {
        th.emplace_back([&, d] {
            const int64_t r0 = shard_begin[d], r1 = shard_begin[d + 1], n = r1 - r0;
            std::vector<int64_t> &o = offs[(size_t)d];
            o.resize((size_t)n + 1);
            for (int64_t i = 0; i <= n; i++) o[(size_t)i] = read_off[r0 + i] - read_off[r0];
            rc[(size_t)d] = meme_seed_batch(ixs[d], codes ? codes + read_off[r0] : nullptr, o.data(), n, opt, &outs[d]);
            if (rc[(size_t)d]) err[(size_t)d] = meme_last_error();
        });
    }
    for (auto &t : th) t.join();
    for (int d = 0; d < ndev; d++)
        if (rc[(size_t)d]) { meme_set_error("shard %d: %s", d, err[(size_t)d].c_str()); return rc[(size_t)d]; }
    return MEME_OK;
}
