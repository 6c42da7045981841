// meme_cigar.cu -- the global alignments of the SAM stage on the device, behind
// meme_cigar_batch of include/meme_b200.h.
//
// Replaces, for all regions of a chunk at once, what mem_reg2aln computes per region when the SAM stage asks for it: the
// band-doubling loop (src/bwamem.cpp:L2339-2348) around bwa_gen_cigar2 (src/bwa.cpp:L274-361) = ksw_global2
// (src/ksw.cpp:L560-670) + NM + MD.  meme_cigar_core.cuh has the per-region procedure.
//
// One thread per region.  Most regions (three quarters on 150 bp reads) need no matrix at all (equal lengths and an inferred
// band of 0: the score is a sum over the diagonal); the others fill a banded matrix of a few thousand cells with a byte of
// backtrack per cell.  Regions are ordered by (band of the first try, reference length) before the alignment kernel runs
// (device radix sort of a 16-bit key), so that the 32 regions of a warp walk the same rows and columns at the same time, and
// a warp's scratch -- DP rows, backtrack matrices, oriented queries -- is interleaved lane by lane: element k of lane l sits
// at [k * 32 + l], so a step of the warp touches one 256-byte line of DP cells and one 32-byte sector of backtrack bytes (the
// first version gave every region its own contiguous arrays: 6 Gcells/s; see profiles/r2/README.md).  Warp scratch is carved
// out of slice-sized arrays by exclusive scans of the per-warp maxima; tasks are processed in slices of 2^18 so that the
// worst case (301 x 301 matrices for every region) stays bounded.  Four kernels per slice: k_cigar_size (needs + sort key),
// k_cigar_group (per-warp maxima), k_cigar_align (alignment, CIGAR length, NM, MD length), k_cigar_emit (compact CIGAR and
// MD strings at the offsets of a scan).
#include <cub/cub.cuh>
#include <vector>
#include <chrono>
#include "meme_common.cuh"
#include "meme_cigar_core.cuh"

using namespace cig;

struct CigarWork {
    meme_cigar_task_t *d_tasks; size_t cap_tasks;
    uint8_t *d_codes; size_t cap_codes;
    int64_t *d_z_off, *d_op_off, *d_md_off; size_t cap_z_off, cap_op_off, cap_md_off;
    uint16_t *d_key, *d_key2; gu32 *d_idx, *d_idx2, *d_slot; size_t cap_key, cap_key2, cap_idx, cap_idx2, cap_slot;
    int64_t *d_g_rows, *d_g_z; size_t cap_g_rows, cap_g_z;
    uint8_t *d_z, *d_q; size_t cap_z, cap_q;
    EH *d_eh; size_t cap_eh;
    meme_cigar_rec_t *d_recs; size_t cap_recs;
    gu32 *d_ops; size_t cap_ops;
    char *d_md; size_t cap_md;
    u64 *d_ctl;                 // [0] error flag, [1] DP cells
    meme_cigar_rec_t *h_recs; gu32 *h_ops; char *h_md; size_t hcap_recs, hcap_ops, hcap_md;
    cudaEvent_t ev[2];
    float ms; u64 cells;
    int ran;
};

void meme_cigar_free(CigarWork *c)
{
    if (!c) return;
    void *dev[] = {c->d_tasks, c->d_codes, c->d_z_off, c->d_op_off, c->d_md_off, c->d_key, c->d_key2, c->d_idx, c->d_idx2, c->d_slot,
                   c->d_g_rows, c->d_g_z, c->d_z, c->d_q, c->d_eh, c->d_recs, c->d_ops, c->d_md, c->d_ctl};
    for (void *p : dev) if (p) cudaFree(p);
    void *host[] = {c->h_recs, c->h_ops, c->h_md};
    for (void *p : host) if (p) cudaFreeHost(p);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    free(c);
}

struct CigarArgs {
    const meme_cigar_task_t *tasks; int64_t n;
    const uint8_t *codes; int64_t n_codes;
    const u64 *pac; u64 sa_num; int64_t l_pac;
    meme_cigar_opt_t opt;
};

__device__ __forceinline__ bool task_valid(const CigarArgs &A, const meme_cigar_task_t &t)
{
    // bwa_gen_cigar2 L285 (empty / bridging the strands) and L287 (bns_get_seq clipped the range); plus what would read
    // outside the arrays
    if (t.qlen <= 0 || t.rb >= t.re || (t.rb < A.l_pac && t.re > A.l_pac)) return false;
    if (t.rb < 0 || t.re > (A.l_pac << 1) || (u64)t.re > A.sa_num) return false;
    if (t.q_off < 0 || t.q_off + t.qlen > A.n_codes || t.qlen > MEME_MAX_READ_LEN || t.re - t.rb > (1 << 20)) return false;
    return true;
}

__global__ void k_cigar_size(CigarArgs A, int64_t *__restrict__ z_need, uint16_t *__restrict__ key, gu32 *__restrict__ idx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    const meme_cigar_task_t t = A.tasks[i];
    int64_t zb = 0;
    int fb = 0, tl = 0;
    if (task_valid(A, t)) {
        tl = (int)(t.re - t.rb);
        zb = z_bytes(A.opt, t.qlen, tl, t.w);
        fb = first_band(A.opt, t.qlen, tl, t.w);
    }
    z_need[i] = zb;
    // widest band first, longest reference first within a band; regions without a matrix last
    const int kb = fb > 255 ? 255 : fb, kt = (tl >> 1) > 255 ? 255 : tl >> 1;
    key[i] = (uint16_t)((255 - kb) << 8 | (255 - kt));
    idx[i] = (gu32)i;
}

// one thread per group of 32 regions (a warp of the alignment kernel): rows of the DP row / query arrays and backtrack bytes
// the group needs = the maxima over its regions
__global__ void k_cigar_group(CigarArgs A, const gu32 *__restrict__ perm, const int64_t *__restrict__ z_need, int64_t n_groups,
                              int64_t *__restrict__ g_rows, int64_t *__restrict__ g_z)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_groups) return;
    int64_t rows = 0, zb = 0;
    if (g < n_groups)
        for (int l = 0; l < 32; l++) {
            const int64_t tix = g * 32 + l;
            if (tix >= A.n) break;
            const int64_t i = perm[tix];
            const meme_cigar_task_t t = A.tasks[i];
            if (!task_valid(A, t)) continue;
            rows = rows > t.qlen + 1 ? rows : t.qlen + 1;
            zb = zb > z_need[i] ? zb : z_need[i];
        }
    g_rows[g] = rows; g_z[g] = zb;
}

struct CigarScratch { EH *eh; uint8_t *z, *q; const int64_t *g_rows, *g_z; };

__device__ __forceinline__ Scr lane_scratch(const CigarScratch &W, int64_t tix)
{
    const int64_t g = tix >> 5, lane = tix & 31;
    Scr C;
    C.eh = W.eh + W.g_rows[g] * 32 + lane;
    C.q = W.q + W.g_rows[g] * 32 + lane;
    C.z = W.z + W.g_z[g] * 32 + lane;
    C.stride = 32;
    return C;
}

__device__ __forceinline__ Seqs task_seqs(const CigarArgs &A, const meme_cigar_task_t &t, const Scr &C)
{
    Seqs S;
    S.qs = C.q; S.stride = C.stride; S.pac = (const gu64 *)A.pac; S.rb = t.rb; S.qlen = t.qlen; S.tlen = (int)(t.re - t.rb);
    S.rev = t.rb >= A.l_pac;
    return S;
}

__global__ void __launch_bounds__(128) k_cigar_align(CigarArgs A, CigarScratch W, const gu32 *__restrict__ perm, gu32 *__restrict__ slot,
                                                     meme_cigar_rec_t *__restrict__ recs, int64_t *__restrict__ op_cnt,
                                                     int64_t *__restrict__ md_cnt, u64 *__restrict__ ctl)
{
    const int64_t tix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tix > A.n) return;
    if (tix == A.n) { op_cnt[A.n] = 0; md_cnt[A.n] = 0; return; }
    const int64_t i = perm[tix];
    slot[i] = (gu32)tix;
    const meme_cigar_task_t t = A.tasks[i];
    meme_cigar_rec_t r;
    r.score = 0; r.nm = -1; r.n_cigar = -1; r.md_len = 0; r.cigar_off = 0; r.md_off = 0;
    if (task_valid(A, t)) {
        const Scr C = lane_scratch(W, tix);
        const Seqs S = task_seqs(A, t, C);
        fill_query(A.codes, t.q_off, t.qlen, S.rev, C);
        int n_ops = 0;
        gu64 cells = 0;
        r.score = region_cigar(A.opt, S, t.w, t.truesc, C, &n_ops, &cells);
        r.n_cigar = n_ops;
        r.md_len = nm_md(S, C, n_ops, nullptr, &r.nm);
        if (cells) atomicAdd((unsigned long long *)&ctl[1], (unsigned long long)cells);
    }
    recs[i] = r;
    op_cnt[i] = r.n_cigar > 0 ? r.n_cigar : 0;
    md_cnt[i] = r.md_len;
}

__global__ void __launch_bounds__(128) k_cigar_emit(CigarArgs A, CigarScratch W, const gu32 *__restrict__ slot,
                                                    const int64_t *__restrict__ op_off, const int64_t *__restrict__ md_off,
                                                    int64_t op_base, int64_t md_base, meme_cigar_rec_t *__restrict__ recs,
                                                    gu32 *__restrict__ ops_out, char *__restrict__ md_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    meme_cigar_rec_t r = recs[i];
    r.cigar_off = op_base + op_off[i];
    r.md_off = md_base + md_off[i];
    recs[i] = r;
    if (r.n_cigar <= 0) return;
    const Scr C = lane_scratch(W, slot[i]);
    gu32 *dst = ops_out + op_off[i];
    for (int k = 0; k < r.n_cigar; k++) dst[k] = C.op(r.n_cigar - 1 - k);
    const Seqs S = task_seqs(A, A.tasks[i], C);
    int nm;
    nm_md(S, C, r.n_cigar, md_out + md_off[i], &nm);
}

static int cigar_work(meme_index_t *ix, CigarWork **out)
{
    if (!ix->cigar) {
        CigarWork *c = (CigarWork *)calloc(1, sizeof(CigarWork));
        if (!c) return MEME_E_NOMEM;
        ix->cigar = c;
        CUDA_TRY(cudaSetDevice(ix->device));
        CUDA_TRY(cudaMalloc((void **)&c->d_ctl, 16));
        for (auto &e : c->ev) CUDA_TRY(cudaEventCreate(&e));
    }
    *out = ix->cigar;
    return MEME_OK;
}

static int cigar_grow_pinned(void **p, size_t *cap, size_t need, size_t elem, size_t keep)
{
    if (need <= *cap && *p) return MEME_OK;
    size_t ncap = need + need / 2 + 1024;
    void *q = nullptr;
    cudaError_t e = cudaMallocHost(&q, ncap * elem);
    if (e != cudaSuccess) { meme_set_error("cudaMallocHost(%zu bytes): %s", ncap * elem, cudaGetErrorString(e)); return MEME_E_NOMEM; }
    if (*p) { if (keep) memcpy(q, *p, keep * elem); cudaFreeHost(*p); }
    *p = q; *cap = ncap;
    return MEME_OK;
}

static int scan_i64(meme_index_t *ix, int64_t *d, int64_t n, cudaStream_t st)
{
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, d, d, (int)n, st);
    int rc;
    if ((rc = meme_grow_device(&ix->d_scan_tmp, &ix->cap_scan_tmp, tmp + 256, 1, st))) return rc;
    size_t t2 = ix->cap_scan_tmp;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(ix->d_scan_tmp, t2, d, d, (int)n, st));
    return MEME_OK;
}

extern "C" int meme_cigar_batch(meme_index_t *ix, const meme_cigar_task_t *tasks, int64_t n_tasks, const uint8_t *codes, int64_t n_codes,
                                int64_t l_pac, const meme_cigar_opt_t *opt, meme_cigar_result_t *out)
{
    if (!ix || !opt || !out || n_tasks < 0 || n_codes < 0 || (n_tasks > 0 && (!tasks || !codes)) || l_pac < 1) { meme_set_error("bad argument"); return MEME_E_ARG; }
    if ((u64)l_pac * 2 > ix->v.sa_num) { meme_set_error("l_pac %lld does not fit an index of %llu suffixes", (long long)l_pac, ix->v.sa_num); return MEME_E_ARG; }
    if (opt->e_del <= 0 || opt->e_ins <= 0 || opt->w < 0) { meme_set_error("bad gap extension penalty / band"); return MEME_E_ARG; }
    CigarWork *c; int rc;
    if ((rc = cigar_work(ix, &c))) return rc;
    CUDA_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    c->ran = 0; c->ms = 0; c->cells = 0;
    const auto Tin = std::chrono::steady_clock::now();
    if ((rc = meme_grow_device((void **)&c->d_codes, &c->cap_codes, (size_t)n_codes + 64, 1, st))) return rc;
    if (n_codes) CUDA_TRY(cudaMemcpyAsync(c->d_codes, codes, (size_t)n_codes, cudaMemcpyHostToDevice, st));
    if ((rc = cigar_grow_pinned((void **)&c->h_recs, &c->hcap_recs, (size_t)n_tasks + 1, sizeof(meme_cigar_rec_t), 0))) return rc;
    CUDA_TRY(cudaMemsetAsync(c->d_ctl, 0, 16, st));
    const int64_t SLICE = 1 << 18;
    int64_t op_base = 0, md_base = 0;
    const bool dbg = getenv("MEME_CIGAR_DEBUG") != nullptr;
    double t_h2d = 0, t_size = 0, t_align = 0, t_emit = 0, t_d2h = 0, t_grow = 0;
    auto now = [&] { if (dbg) cudaStreamSynchronize(st); return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count() * 1e3; };
    auto T0 = now();
    CUDA_TRY(cudaEventRecord(c->ev[0], st));
    for (int64_t s0 = 0; s0 < n_tasks; s0 += SLICE) {
        const int64_t n = n_tasks - s0 < SLICE ? n_tasks - s0 : SLICE;
        const int64_t n_groups = (n + 31) / 32;
        auto Tg = now();
        void **grow[] = {(void **)&c->d_tasks, (void **)&c->d_z_off, (void **)&c->d_op_off, (void **)&c->d_md_off, (void **)&c->d_key,
                         (void **)&c->d_key2, (void **)&c->d_idx, (void **)&c->d_idx2, (void **)&c->d_slot, (void **)&c->d_recs, (void **)&c->d_g_rows,
                         (void **)&c->d_g_z};
        size_t *caps[] = {&c->cap_tasks, &c->cap_z_off, &c->cap_op_off, &c->cap_md_off, &c->cap_key, &c->cap_key2, &c->cap_idx, &c->cap_idx2, &c->cap_slot,
                          &c->cap_recs, &c->cap_g_rows, &c->cap_g_z};
        size_t elem[] = {sizeof(meme_cigar_task_t), 8, 8, 8, 2, 2, 4, 4, 4, sizeof(meme_cigar_rec_t), 8, 8};
        for (int k = 0; k < 12; k++)
            if ((rc = meme_grow_device(grow[k], caps[k], (size_t)n + 2, elem[k], st))) return rc;
        auto Ta = now(); t_grow += secs(Tg, Ta);
        CUDA_TRY(cudaMemcpyAsync(c->d_tasks, tasks + s0, (size_t)n * sizeof(meme_cigar_task_t), cudaMemcpyHostToDevice, st));
        auto Tb = now(); t_h2d += secs(Ta, Tb);
        CigarArgs A;
        A.tasks = c->d_tasks; A.n = n; A.codes = c->d_codes; A.n_codes = n_codes; A.pac = ix->v.pac; A.sa_num = ix->v.sa_num; A.l_pac = l_pac;
        A.opt = *opt;
        const unsigned blocks = (unsigned)((n + 1 + 127) / 128);
        k_cigar_size<<<blocks, 128, 0, st>>>(A, c->d_z_off, c->d_key, c->d_idx);
        CUDA_TRY(cudaGetLastError());
        {
            size_t tmp = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->d_key, c->d_key2, c->d_idx, c->d_idx2, (int)n, 0, 16, st);
            if ((rc = meme_grow_device(&ix->d_scan_tmp, &ix->cap_scan_tmp, tmp + 256, 1, st))) return rc;
            size_t t2 = ix->cap_scan_tmp;
            CUDA_TRY(cub::DeviceRadixSort::SortPairs(ix->d_scan_tmp, t2, c->d_key, c->d_key2, c->d_idx, c->d_idx2, (int)n, 0, 16, st));
        }
        k_cigar_group<<<(unsigned)((n_groups + 1 + 127) / 128), 128, 0, st>>>(A, c->d_idx2, c->d_z_off, n_groups, c->d_g_rows, c->d_g_z);
        CUDA_TRY(cudaGetLastError());
        if ((rc = scan_i64(ix, c->d_g_rows, n_groups + 1, st))) return rc;
        if ((rc = scan_i64(ix, c->d_g_z, n_groups + 1, st))) return rc;
        int64_t tot[2];
        CUDA_TRY(cudaMemcpyAsync(&tot[0], c->d_g_z + n_groups, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(&tot[1], c->d_g_rows + n_groups, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (dbg && ((size_t)tot[0] * 32 + 64 > c->cap_z || (size_t)tot[1] * 32 + 64 > c->cap_eh))
            fprintf(stderr, "[meme_cigar] slice at %lld grows its scratch: backtrack %zu -> %zu bytes, rows %zu -> %zu cells\n", (long long)s0, c->cap_z,
                    (size_t)tot[0] * 32 + 64, c->cap_eh, (size_t)tot[1] * 32 + 64);
        if ((rc = meme_grow_device((void **)&c->d_z, &c->cap_z, (size_t)tot[0] * 32 + 64, 1, st))) return rc;
        if ((rc = meme_grow_device((void **)&c->d_eh, &c->cap_eh, (size_t)tot[1] * 32 + 64, sizeof(EH), st))) return rc;
        if ((rc = meme_grow_device((void **)&c->d_q, &c->cap_q, (size_t)tot[1] * 32 + 64, 1, st))) return rc;
        CigarScratch W;
        W.eh = c->d_eh; W.z = c->d_z; W.q = c->d_q; W.g_rows = c->d_g_rows; W.g_z = c->d_g_z;
        auto Tc = now(); t_size += secs(Tb, Tc);
        k_cigar_align<<<blocks, 128, 0, st>>>(A, W, c->d_idx2, c->d_slot, c->d_recs, c->d_op_off, c->d_md_off, c->d_ctl);
        CUDA_TRY(cudaGetLastError());
        auto Td = now(); t_align += secs(Tc, Td);
        if ((rc = scan_i64(ix, c->d_op_off, n + 1, st))) return rc;
        if ((rc = scan_i64(ix, c->d_md_off, n + 1, st))) return rc;
        CUDA_TRY(cudaMemcpyAsync(&tot[0], c->d_op_off + n, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(&tot[1], c->d_md_off + n, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if ((rc = meme_grow_device((void **)&c->d_ops, &c->cap_ops, (size_t)tot[0] + 8, 4, st))) return rc;
        if ((rc = meme_grow_device((void **)&c->d_md, &c->cap_md, (size_t)tot[1] + 8, 1, st))) return rc;
        k_cigar_emit<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(A, W, c->d_slot, c->d_op_off, c->d_md_off, op_base, md_base, c->d_recs, c->d_ops, c->d_md);
        CUDA_TRY(cudaGetLastError());
        auto Te = now(); t_emit += secs(Td, Te);
        if ((rc = cigar_grow_pinned((void **)&c->h_ops, &c->hcap_ops, (size_t)(op_base + tot[0]) + 1, 4, (size_t)op_base))) return rc;
        if ((rc = cigar_grow_pinned((void **)&c->h_md, &c->hcap_md, (size_t)(md_base + tot[1]) + 1, 1, (size_t)md_base))) return rc;
        CUDA_TRY(cudaMemcpyAsync(c->h_recs + s0, c->d_recs, (size_t)n * sizeof(meme_cigar_rec_t), cudaMemcpyDeviceToHost, st));
        if (tot[0]) CUDA_TRY(cudaMemcpyAsync(c->h_ops + op_base, c->d_ops, (size_t)tot[0] * 4, cudaMemcpyDeviceToHost, st));
        if (tot[1]) CUDA_TRY(cudaMemcpyAsync(c->h_md + md_base, c->d_md, (size_t)tot[1], cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        auto Tf = now(); t_d2h += secs(Te, Tf);
        op_base += tot[0]; md_base += tot[1];
        ix->launches += 4;
    }
    CUDA_TRY(cudaEventRecord(c->ev[1], st));
    u64 ctl[2];
    CUDA_TRY(cudaMemcpyAsync(ctl, c->d_ctl, 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaEventElapsedTime(&c->ms, c->ev[0], c->ev[1]));
    c->cells = ctl[1];
    c->ran = 1;
    if (dbg) fprintf(stderr, "[meme_cigar] %lld tasks: total %.1f ms (codes+setup %.1f), per phase summed over slices: tasks h2d %.1f, size+scan+sort %.1f, align %.1f, "
                     "scan+emit %.1f, d2h %.1f, grow %.1f; %llu cells\n", (long long)n_tasks, secs(Tin, now()), secs(Tin, T0), t_h2d, t_size, t_align, t_emit, t_d2h, t_grow, (unsigned long long)c->cells);
    out->n_tasks = n_tasks; out->n_ops = op_base; out->n_md = md_base;
    out->recs = c->h_recs; out->cigars = c->h_ops; out->md = c->h_md;
    return MEME_OK;
}

// Kernels loaded and every array sized for chunks of about n_regions regions of qlen-base queries before the first timed chunk:
// one full slice of dummy regions runs through every kernel -- a quarter of them with a matrix and a local score they cannot
// reach, so that the band doubles to its widest and the scratch arrays take their worst-case size for this read length -- and
// the pinned result arrays are sized for the whole chunk.  After this a chunk of the expected shape allocates nothing (on some
// boxes a cudaMalloc / cudaFree inside a timed chunk was seen to take 0.7-3.6 s).
extern "C" int meme_cigar_warmup(meme_index_t *ix, int64_t n_regions, int qlen)
{
    if (!ix || n_regions < 0 || qlen < 1 || qlen > 1000) { meme_set_error("bad argument"); return MEME_E_ARG; }
    CigarWork *c; int rc;
    if ((rc = cigar_work(ix, &c))) return rc;
    CUDA_TRY(cudaSetDevice(ix->device));
    const int64_t l_pac = (int64_t)(ix->v.sa_num / 2);
    const int64_t n = n_regions < (1 << 18) ? n_regions : (1 << 18);
    const int ql = l_pac > 8 * (int64_t)qlen ? qlen : 24;
    if (l_pac < 4 * ql || n == 0) return MEME_OK;
    std::vector<meme_cigar_task_t> tasks((size_t)n);
    std::vector<uint8_t> codes((size_t)n * ql);
    u64 x = 88172645463325252ull;
    for (auto &b : codes) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; b = (uint8_t)(x & 3); }
    for (int64_t i = 0; i < n; i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        meme_cigar_task_t &T = tasks[(size_t)i];
        const bool matrix = (i & 3) == 0;
        T.rb = (int64_t)(x % (u64)(l_pac - 2 * ql)); T.re = T.rb + ql + (matrix ? 2 : 0);
        T.q_off = i * ql; T.qlen = ql; T.w = matrix ? 30 : 0; T.truesc = matrix ? ql : 0; T.pad = 0;
    }
    meme_cigar_opt_t o;
    memset(&o, 0, sizeof o);
    for (int a = 0; a < 5; a++) for (int b = 0; b < 5; b++) o.mat[a * 5 + b] = a == 4 || b == 4 ? -1 : a == b ? 1 : -4;
    o.o_del = o.o_ins = 6; o.e_del = o.e_ins = 1; o.w = 100; o.a = 1;
    meme_cigar_result_t r;
    if ((rc = meme_cigar_batch(ix, tasks.data(), n, codes.data(), (int64_t)codes.size(), l_pac, &o, &r))) return rc;
    if ((rc = cigar_grow_pinned((void **)&c->h_recs, &c->hcap_recs, (size_t)n_regions, sizeof(meme_cigar_rec_t), 0))) return rc;
    if ((rc = cigar_grow_pinned((void **)&c->h_ops, &c->hcap_ops, (size_t)n_regions * 2, 4, 0))) return rc;
    if ((rc = cigar_grow_pinned((void **)&c->h_md, &c->hcap_md, (size_t)n_regions * 12, 1, 0))) return rc;
    if ((rc = meme_grow_device((void **)&c->d_codes, &c->cap_codes, (size_t)n_regions * (size_t)qlen / 2 + 64, 1, ix->stream))) return rc;
    if (n == (1 << 18)) {
        // backtrack scratch for a slice in which half of the warps hold full-width matrices (a GB-sized cudaFree + cudaMalloc
        // inside the first timed chunk was measured at 0.26-0.8 s); HBM has the room
        const size_t groups = (size_t)n / 32, full = (size_t)ql * (size_t)(ql + 32);
        if ((rc = meme_grow_device((void **)&c->d_z, &c->cap_z, groups / 2 * 32 * full + 64, 1, ix->stream))) return rc;
        if ((rc = meme_grow_device((void **)&c->d_eh, &c->cap_eh, groups * 32 * (size_t)(ql + 34) + 64, sizeof(EH), ix->stream))) return rc;
        if ((rc = meme_grow_device((void **)&c->d_q, &c->cap_q, groups * 32 * (size_t)(ql + 34) + 64, 1, ix->stream))) return rc;
    }
    c->ran = 0;
    return MEME_OK;
}

extern "C" int meme_cigar_last(meme_index_t *ix, float *ms, uint64_t *cells)
{
    if (!ix || !ix->cigar || !ix->cigar->ran) { meme_set_error("no completed meme_cigar_batch"); return MEME_E_ARG; }
    if (ms) *ms = ix->cigar->ms;
    if (cells) *cells = ix->cigar->cells;
    return MEME_OK;
}
