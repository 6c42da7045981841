// meme_index.cu -- index load + on-GPU runtime build.
//
// Replaces memoryAllocLearned (reference src/fastmap.cpp:L351-638): the reference reads `.pos_packed`,
// expands it in place on the host into 13-byte {pos5,key8} entries and scatters the inverse suffix
// array with OpenMP.  Here the 5-byte entries are streamed to the GPU in chunks and three kernels
// do the expansion into an SoA layout:
//     pos_lo[i] (+pos_hi[i])   suffix position          4 (+1) B   coalesced 128 B per warp probe
//     keys[i]                  first 32 bases, 2 bit    8 B        coalesced 256 B per warp probe
//     isa_lo[p] (+isa_hi[p])   inverse suffix array     4 (+1) B   (mode 3)
//     pac[w]                   fwd+RC text, 32 bases/u64, first base in the LSBs
// Key definition = get_key_of_ref (src/LearnedIndex_seeding.h:L184-201): base r of the suffix in bits
// 2r..2r+1, positions past the text end read as 3 ('T').
#include <stdarg.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <thread>
#include <string>
#include <vector>
#include "meme_common.cuh"

static thread_local char g_err[1024] = "";

void meme_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char *meme_last_error(void) { return g_err; }
extern "C" const char *meme_version(void) { return "meme_b200 0.1 (sm_100a)"; }
extern "C" int meme_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int meme_grow_device(void **p, size_t *cap, size_t need, size_t elem, cudaStream_t s)
{
    if (need <= *cap && *p) return MEME_OK;
    size_t ncap = need + need / 4 + 64;
    if (*p) { CUDA_TRY(cudaStreamSynchronize(s)); CUDA_TRY(cudaFree(*p)); *p = nullptr; *cap = 0; }
    cudaError_t e = cudaMalloc(p, ncap * elem);
    if (e != cudaSuccess) { meme_set_error("cudaMalloc(%zu bytes): %s", ncap * elem, cudaGetErrorString(e)); return MEME_E_NOMEM; }
    *cap = ncap;
    return MEME_OK;
}

int meme_grow_pinned(void **p, size_t *cap, size_t need, size_t elem)
{
    if (need <= *cap && *p) return MEME_OK;
    size_t ncap = need + need / 4 + 64;
    if (*p) { cudaFreeHost(*p); *p = nullptr; *cap = 0; }
    cudaError_t e = cudaMallocHost(p, ncap * elem);
    if (e != cudaSuccess) { meme_set_error("cudaMallocHost(%zu bytes): %s", ncap * elem, cudaGetErrorString(e)); return MEME_E_NOMEM; }
    *cap = ncap;
    return MEME_OK;
}

// ---------------------------------------------------------------------------------------------
// build kernels
// ---------------------------------------------------------------------------------------------

// 32 one-byte codes -> one 64-bit word (first base in the LSBs). `text` is a chunk starting at
// base `base0` (a multiple of 32); words past the text end are zero.
__global__ void k_pack_text(const uint8_t *__restrict__ text, u64 n_chunk, u64 *__restrict__ pac, u64 word0)
{
    u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 b = w * 32;
    if (b >= n_chunk) return;
    u64 v = 0;
    if (b + 32 <= n_chunk) {
        const uint4 *p = reinterpret_cast<const uint4 *>(text + b);   // chunk base is 32-aligned
        uint4 x0 = p[0], x1 = p[1];
        u32 q[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            u32 c = q[k];
            u64 four = (c & 3) | (((c >> 8) & 3) << 2) | (((c >> 16) & 3) << 4) | (((c >> 24) & 3) << 6);
            v |= four << (8 * k);
        }
    } else {
        for (int k = 0; b + k < n_chunk; k++) v |= (u64)(text[b + k] & 3) << (2 * k);
    }
    pac[word0 + w] = v;
}

// 5-byte packed SA entries (uint32 LE = pos>>8, uint8 = pos&0xff; src/Learnedindex.cpp:L469-473)
__global__ void k_unpack_pos(const uint8_t *__restrict__ packed, u64 n_chunk, u64 i0, u32 *__restrict__ pos_lo,
                             uint8_t *__restrict__ pos_hi)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chunk) return;
    const uint8_t *p = packed + i * 5;
    u64 hi = (u64)p[0] | ((u64)p[1] << 8) | ((u64)p[2] << 16) | ((u64)p[3] << 24);
    u64 pos = (hi << 8) | p[4];
    pos_lo[i0 + i] = (u32)pos;
    if (pos_hi) pos_hi[i0 + i] = (uint8_t)(pos >> 32);
}

__device__ __forceinline__ u64 text32(const u64 *__restrict__ pac, u64 p)
{
    u64 w = p >> 5;
    int sh = (int)(p & 31) * 2;
    u64 a = pac[w];
    if (sh == 0) return a;
    return (a >> sh) | (pac[w + 1] << (64 - sh));
}

__global__ void k_build_keys_isa(IndexView v, u64 *__restrict__ keys, u32 *__restrict__ isa_lo,
                                 uint8_t *__restrict__ isa_hi)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.sa_num) return;
    u64 p = v.pos_lo[i];
    if (v.pos_hi) p |= (u64)v.pos_hi[i] << 32;
    if (keys) {
        u64 k = text32(v.pac, p);
        u64 left = v.sa_num - p;                 // bases before the text end
        if (left < 32) k |= ~0ull << (2 * left); // pad with 3 (get_key_of_ref, L187-191)
        keys[i] = k;
    }
    if (isa_lo) {
        isa_lo[p] = (u32)i;
        if (isa_hi) isa_hi[p] = (uint8_t)(i >> 32);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Source {                 // either a file or a host buffer, read in chunks
    FILE *f = nullptr;
    const uint8_t *mem = nullptr;
    u64 size = 0, off = 0;
    size_t read(uint8_t *dst, size_t n)
    {
        if (off + n > size) n = size - off;
        if (f) { if (fread(dst, 1, n, f) != n) return 0; }
        else memcpy(dst, mem + off, n);
        off += n;
        return n;
    }
};

static bool open_file(const std::string &path, Source &s)
{
    s.f = fopen(path.c_str(), "rb");
    if (!s.f) return false;
    fseeko(s.f, 0, SEEK_END);
    s.size = (u64)ftello(s.f);
    fseeko(s.f, 0, SEEK_SET);
    return true;
}

static int upload_rmi(Source &src, RmiRec **d_out, u64 *n_out, cudaStream_t st, u64 *bytes)
{
    u64 n = src.size / 24;
    *n_out = n;
    std::vector<uint8_t> raw(n * 24 + 24);
    if (n && src.read(raw.data(), n * 24) != n * 24) { meme_set_error("short read of RMI parameters"); return MEME_E_IO; }
    std::vector<RmiRec> rec(n ? n : 1);
    for (u64 i = 0; i < n; i++) {
        memcpy(&rec[i].alpha, &raw[i * 24], 8);
        memcpy(&rec[i].beta, &raw[i * 24 + 8], 8);
        memcpy(&rec[i].err, &raw[i * 24 + 16], 8);
        rec[i].pad = 0;
    }
    CUDA_TRY(cudaMalloc((void **)d_out, rec.size() * sizeof(RmiRec)));
    CUDA_TRY(cudaMemcpyAsync(*d_out, rec.data(), rec.size() * sizeof(RmiRec), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *bytes += rec.size() * sizeof(RmiRec);
    return MEME_OK;
}

extern "C" void meme_index_free(meme_index_t *ix);

static int build_index(Source &text, Source &pos, Source &l1, Source &l2, int mode, int device, meme_index_t **out)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        meme_set_error("no CUDA device visible: meme_b200 has no CPU path");
        return MEME_E_NOGPU;
    }
    if (mode < 1 || mode > 3 || !out) { meme_set_error("bad mode/out"); return MEME_E_ARG; }
    if (pos.size != text.size * 5) {
        meme_set_error(".pos_packed holds %llu entries but .0123 holds %llu bases", pos.size / 5, text.size);
        return MEME_E_IO;
    }
    if (l2.size == 0 || (l2.size / 24) & ((l2.size / 24) - 1)) {
        meme_set_error("_L2_PARAMETERS must hold a power-of-two number of 24-byte records");
        return MEME_E_IO;
    }
    CUDA_TRY(cudaSetDevice(device));
    meme_index *ix = (meme_index *)calloc(1, sizeof(meme_index));
    if (!ix) return MEME_E_NOMEM;
    ix->device = device;
    // every early return below releases what was allocated so far (the handle, its arrays, the staging buffers)
    struct Guard {
        meme_index *ix; uint8_t *h_stage = nullptr, *d_stage = nullptr; bool ok = false;
        ~Guard() { if (h_stage) cudaFreeHost(h_stage); if (d_stage) cudaFree(d_stage); if (!ok) meme_index_free(ix); }
    } guard{ix};
    CUDA_TRY(cudaStreamCreateWithFlags(&ix->own_stream, cudaStreamNonBlocking));
    ix->stream = ix->own_stream;
    cudaStream_t st = ix->stream;
    u64 n = text.size;
    IndexView &v = ix->v;
    v.sa_num = n;
    v.mode = mode;
    // MEME_FORCE_WIDE=1 (tests): treat a small index as a >= 2^32-entry one, so that the 5-byte position arrays and
    // the 64-bit-index kernel instantiations (GRCh38 scale) run on fixtures that fit a test
    v.pos_wide = n >= (1ull << 32) || (getenv("MEME_FORCE_WIDE") && atoi(getenv("MEME_FORCE_WIDE")) != 0);
    u64 n_words = (n + 31) / 32 + 2;
    CUDA_TRY(cudaMalloc((void **)&ix->d_pac, n_words * 8));
    CUDA_TRY(cudaMemsetAsync(ix->d_pac, 0, n_words * 8, st));
    CUDA_TRY(cudaMalloc((void **)&ix->d_pos_lo, n * 4));
    ix->device_bytes = n_words * 8 + n * 4;
    if (v.pos_wide) { CUDA_TRY(cudaMalloc((void **)&ix->d_pos_hi, n)); ix->device_bytes += n; }
    if (mode >= 2) { CUDA_TRY(cudaMalloc((void **)&ix->d_keys, n * 8)); ix->device_bytes += n * 8; }
    if (mode == 3) {
        CUDA_TRY(cudaMalloc((void **)&ix->d_isa_lo, n * 4));
        ix->device_bytes += n * 4;
        if (v.pos_wide) { CUDA_TRY(cudaMalloc((void **)&ix->d_isa_hi, n)); ix->device_bytes += n; }
    }
    // chunked staging: pinned host buffer + device buffer
    const size_t CH = (size_t)160 << 20;                    // bytes; multiple of 32 and of 5
    CUDA_TRY(cudaMallocHost((void **)&guard.h_stage, CH));
    CUDA_TRY(cudaMalloc((void **)&guard.d_stage, CH));
    uint8_t *const h_stage = guard.h_stage, *const d_stage = guard.d_stage;
    for (u64 done = 0; done < n;) {                         // text -> pac
        size_t got = text.read(h_stage, CH);
        if (!got) { meme_set_error("short read of .0123"); return MEME_E_IO; }
        CUDA_TRY(cudaMemcpyAsync(d_stage, h_stage, got, cudaMemcpyHostToDevice, st));
        u64 words = (got + 31) / 32;
        k_pack_text<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(d_stage, got, ix->d_pac, done / 32);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(st));
        done += got;
    }
    for (u64 done = 0; done < n;) {                         // pos_packed -> pos
        size_t got = pos.read(h_stage, CH);
        if (!got) { meme_set_error("short read of .pos_packed"); return MEME_E_IO; }
        u64 cnt = got / 5;
        CUDA_TRY(cudaMemcpyAsync(d_stage, h_stage, got, cudaMemcpyHostToDevice, st));
        k_unpack_pos<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(d_stage, cnt, done, ix->d_pos_lo, ix->d_pos_hi);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(st));
        done += cnt;
    }
    v.pac = ix->d_pac; v.pos_lo = ix->d_pos_lo; v.pos_hi = ix->d_pos_hi;
    if (mode >= 2) {
        k_build_keys_isa<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, ix->d_keys, ix->d_isa_lo, ix->d_isa_hi);
        CUDA_TRY(cudaGetLastError());
    }
    v.keys = ix->d_keys; v.isa_lo = ix->d_isa_lo; v.isa_hi = ix->d_isa_hi;
    int rc;
    if ((rc = upload_rmi(l1, &ix->d_l1, &v.n_l1, st, &ix->device_bytes)) != MEME_OK) return rc;
    if ((rc = upload_rmi(l2, &ix->d_l2, &v.n_l2, st, &ix->device_bytes)) != MEME_OK) return rc;
    v.l1 = ix->d_l1; v.l2 = ix->d_l2;
    int bits = 0;
    while ((1ull << bits) < v.n_l2) bits++;
    v.bit_shift = 64 - bits;                                // LearnedIndex_seeding.cpp:L113-119
    CUDA_TRY(cudaMalloc((void **)&ix->d_ctrl, 64 * 8));
    CUDA_TRY(cudaMemsetAsync(ix->d_ctrl, 0, 64 * 8, st));
    CUDA_TRY(cudaMallocHost((void **)&ix->h_ctrl, 64 * 8));
    for (int k = 0; k < 6; k++) CUDA_TRY(cudaEventCreate(&ix->ev[k]));
    CUDA_TRY(cudaStreamSynchronize(st));
    ix->launches = 3;
    guard.ok = true;
    *out = ix;
    return MEME_OK;
}

extern "C" int meme_index_load(const char *prefix, int mode, int device, meme_index_t **out)
{
    if (!prefix || !out) { meme_set_error("null argument"); return MEME_E_ARG; }
    std::string p(prefix);
    Source text, pos, l1, l2;
    int rc = MEME_E_IO;
    if (!open_file(p + ".0123", text)) meme_set_error("cannot open %s.0123", prefix);
    else if (!open_file(p + ".pos_packed", pos)) meme_set_error("cannot open %s.pos_packed", prefix);
    else if (!open_file(p + ".suffixarray_uint64_L1_PARAMETERS", l1)) meme_set_error("cannot open %s.suffixarray_uint64_L1_PARAMETERS", prefix);
    else if (!open_file(p + ".suffixarray_uint64_L2_PARAMETERS", l2)) meme_set_error("cannot open %s.suffixarray_uint64_L2_PARAMETERS", prefix);
    else rc = build_index(text, pos, l1, l2, mode, device, out);
    if (text.f) fclose(text.f);
    if (pos.f) fclose(pos.f);
    if (l1.f) fclose(l1.f);
    if (l2.f) fclose(l2.f);
    return rc;
}

// P-RMI inference for a batch of keys (learned_index_lookup, reference src/LearnedIndex_seeding.cpp:L186-210), with the
// same arithmetic the seeding kernels use on the re-laid-out records: fp64 fma, optional partial layer, clamp.  The
// seeding kernels inline this between their memory requests; this entry point exists so that tests can pin the device
// prediction and error word against the reference formula key by key.
__global__ void k_rmi_lookup(IndexView ix, const u64 *__restrict__ keys, int64_t n, u64 *__restrict__ pred, u64 *__restrict__ err)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 key = keys[i];
    RmiRec r = ix.l2[key >> ix.bit_shift];
    double p = fma(r.beta, __ull2double_rn(key), r.alpha);
    u64 e = r.err;
    if (e >> 63) {
        const u64 st = (e >> 32) & 0x7fffffff;
        const double bound = (double)(e & 0xffffffffull) - 1;
        const u64 sub = p < 0.0 ? 0 : (p > bound ? (u64)bound : (u64)p);
        r = ix.l1[st + sub];
        p = fma(r.beta, __ull2double_rn(key), r.alpha);
        e = r.err;
    }
    const double bound = (double)ix.sa_num - 1.0;
    pred[i] = p < 0.0 ? 0 : (p > bound ? (u64)bound : (u64)p);
    err[i] = e;
}

extern "C" int meme_rmi_lookup(meme_index_t *ix, const uint64_t *keys, int64_t n, uint64_t *pred, uint64_t *err)
{
    if (!ix || n < 0 || (n && (!keys || !pred || !err))) { meme_set_error("bad argument"); return MEME_E_ARG; }
    if (n == 0) return MEME_OK;
    CUDA_TRY(cudaSetDevice(ix->device));
    u64 *d = nullptr;
    CUDA_TRY(cudaMalloc((void **)&d, (size_t)n * 8 * 3));
    CUDA_TRY(cudaMemcpyAsync(d, keys, (size_t)n * 8, cudaMemcpyHostToDevice, ix->stream));
    k_rmi_lookup<<<(unsigned)((n + 255) / 256), 256, 0, ix->stream>>>(ix->v, d, n, d + n, d + 2 * n);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(pred, d + n, (size_t)n * 8, cudaMemcpyDeviceToHost, ix->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(err, d + 2 * n, (size_t)n * 8, cudaMemcpyDeviceToHost, ix->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
    cudaFree(d);
    if (e != cudaSuccess) { meme_set_error("meme_rmi_lookup: %s", cudaGetErrorString(e)); return MEME_E_CUDA; }
    return MEME_OK;
}

// The index replicated on `ndev` devices (reads shard across GPUs, the index does not: SURVEY.md section 8e).  The
// devices are loaded concurrently, one host thread each; out receives ndev handles (all or none).
extern "C" int meme_index_load_devices(const char *prefix, int mode, const int *devices, int ndev, meme_index_t **out)
{
    if (!prefix || !devices || ndev < 1 || !out) { meme_set_error("bad argument"); return MEME_E_ARG; }
    std::vector<int> rc((size_t)ndev, MEME_OK);
    std::vector<std::string> err((size_t)ndev);
    std::vector<std::thread> th;
    for (int d = 0; d < ndev; d++) {
        out[d] = nullptr;
        th.emplace_back([&, d] {
            rc[(size_t)d] = meme_index_load(prefix, mode, devices[d], &out[d]);
            if (rc[(size_t)d]) err[(size_t)d] = meme_last_error();
        });
    }
    for (auto &t : th) t.join();
    for (int d = 0; d < ndev; d++)
        if (rc[(size_t)d]) {
            for (int k = 0; k < ndev; k++) if (out[k]) { meme_index_free(out[k]); out[k] = nullptr; }
            meme_set_error("device %d: %s", devices[d], err[(size_t)d].c_str());
            return rc[(size_t)d];
        }
    return MEME_OK;
}

extern "C" int meme_index_from_host(const uint8_t *text0123, uint64_t sa_num, const uint8_t *pos_packed,
                                    const void *l1_params, uint64_t n_l1, const void *l2_params, uint64_t n_l2,
                                    int mode, int device, meme_index_t **out)
{
    if (!text0123 || !pos_packed || !l1_params || !l2_params || !out) { meme_set_error("null argument"); return MEME_E_ARG; }
    Source text, pos, l1, l2;
    text.mem = text0123; text.size = sa_num;
    pos.mem = pos_packed; pos.size = sa_num * 5;
    l1.mem = (const uint8_t *)l1_params; l1.size = n_l1 * 24;
    l2.mem = (const uint8_t *)l2_params; l2.size = n_l2 * 24;
    return build_index(text, pos, l1, l2, mode, device, out);
}

extern "C" void meme_index_free(meme_index_t *ix)
{
    if (!ix) return;
    cudaSetDevice(ix->device);
    cudaStreamSynchronize(ix->stream);
    if (ix->chain) meme_chain_free(ix->chain);
    if (ix->cigar) meme_cigar_free(ix->cigar);
    if (ix->matesw) meme_matesw_free(ix->matesw);
    void *dev[] = {ix->d_keys, ix->d_pos_lo, ix->d_pos_hi, ix->d_isa_lo, ix->d_isa_hi, ix->d_pac, ix->d_l1, ix->d_l2,
                   ix->d_codes, ix->d_read_off, ix->d_rd_pool_off, ix->d_rd_nsmem, ix->d_rd_nhit, ix->d_smem_off,
                   ix->d_hit_off, ix->d_pool, ix->d_ctrl, ix->d_scan_tmp, ix->d_smems, ix->d_hits, ix->d_redo,
                   ix->d_packed, ix->d_nmask, ix->d_rd_full, ix->d_rd_maxref, ix->d_pool_keys, ix->d_pool_vals, ix->d_pool_keys2,
                   ix->d_pool_vals2, ix->d_t2};
    for (void *p : dev) if (p) cudaFree(p);
    void *alt[] = {ix->alt.d_codes, ix->alt.d_read_off, ix->alt.d_smem_off, ix->alt.d_hit_off, ix->alt.d_smems, ix->alt.d_hits};
    for (void *p : alt) if (p) cudaFree(p);
    if (ix->copy_stream) { cudaStreamDestroy(ix->copy_stream); for (int k = 0; k < 6; k++) cudaEventDestroy(ix->pe[k]); }
    void *host[] = {ix->h_smems, ix->h_hits, ix->h_smem_off, ix->h_hit_off, ix->h_ctrl};
    for (void *p : host) if (p) cudaFreeHost(p);
    for (int k = 0; k < 6; k++) if (ix->ev[k]) cudaEventDestroy(ix->ev[k]);
    if (ix->own_stream) cudaStreamDestroy(ix->own_stream);
    free(ix);
}

extern "C" uint64_t meme_index_sa_num(const meme_index_t *ix) { return ix ? ix->v.sa_num : 0; }
extern "C" uint64_t meme_index_device_bytes(const meme_index_t *ix) { return ix ? ix->device_bytes : 0; }

extern "C" int meme_index_set_stream(meme_index_t *ix, void *cuda_stream)
{
    if (!ix) return MEME_E_ARG;
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    ix->stream = cuda_stream ? (cudaStream_t)cuda_stream : ix->own_stream;
    return MEME_OK;
}

__global__ void k_widen(const u32 *__restrict__ lo, const uint8_t *__restrict__ hi, u64 n, u64 *__restrict__ out)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (u64)lo[i] | (hi ? (u64)hi[i] << 32 : 0);
}

extern "C" int meme_index_download(const meme_index_t *ix, uint64_t *keys, uint64_t *pos, uint64_t *isa)
{
    if (!ix) return MEME_E_ARG;
    CUDA_TRY(cudaSetDevice(ix->device));
    u64 n = ix->v.sa_num;
    if (keys) {
        if (!ix->d_keys) { meme_set_error("index mode 1 holds no inline keys"); return MEME_E_ARG; }
        CUDA_TRY(cudaMemcpy(keys, ix->d_keys, n * 8, cudaMemcpyDeviceToHost));
    }
    u64 *tmp = nullptr;
    if (pos || isa) CUDA_TRY(cudaMalloc((void **)&tmp, n * 8));
    if (pos) {
        k_widen<<<(unsigned)((n + 255) / 256), 256, 0, ix->stream>>>(ix->d_pos_lo, ix->d_pos_hi, n, tmp);
        CUDA_TRY(cudaStreamSynchronize(ix->stream));
        CUDA_TRY(cudaMemcpy(pos, tmp, n * 8, cudaMemcpyDeviceToHost));
    }
    if (isa) {
        if (!ix->d_isa_lo) { cudaFree(tmp); meme_set_error("index holds no inverse SA (mode < 3)"); return MEME_E_ARG; }
        k_widen<<<(unsigned)((n + 255) / 256), 256, 0, ix->stream>>>(ix->d_isa_lo, ix->d_isa_hi, n, tmp);
        CUDA_TRY(cudaStreamSynchronize(ix->stream));
        CUDA_TRY(cudaMemcpy(isa, tmp, n * 8, cudaMemcpyDeviceToHost));
    }
    if (tmp) cudaFree(tmp);
    return MEME_OK;
}
