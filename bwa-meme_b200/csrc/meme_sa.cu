// meme_sa.cu -- the index pipeline on the device: suffix array of the fwd+RC text in the reference's
// order, and the 64-bit keys of its suffixes.  Replaces the sais call and the key loop of the reference's index builder
// (buildSAandLEP / build_learned_index, src/Learnedindex.cpp:L134-555: text = fwd + RC padded with P = max(longest A run,
// longest T run) + 1 `T`s, sais over the padded text, suffixes that start in the padding dropped, L190-242, L469-473; key =
// the suffix's first 32 bases over the padded text, cyclic, L481-500).  The output is what the reference writes to
// <prefix>.pos_packed (5 bytes per entry) and <prefix>.suffixarray_uint64; the file writing stays with the caller.
//
// Algorithm: prefix doubling, B200-sized.  Round 0 sorts all suffixes by their first 27 symbols (a base-5 number in 63 bits:
// A, C, G, T = 1..4, end of string = 0, so that a suffix that is a prefix of another sorts first, as sais orders them); on a
// genome that already leaves only the repeats unresolved.  Every later round sorts ONLY the still ambiguous slots by
// (bucket, rank of the suffix h further on), doubling h: the work shrinks with the repeat content, and everything is a
// device radix sort (cub) between five small kernels.  At chr1 scale (498 M suffixes) the arrays take about 30 GB of the 180 GB.
#include <cub/cub.cuh>
#include <vector>
#include "meme_common.cuh"

namespace {

struct SaBuf {
    std::vector<void *> dev;
    ~SaBuf() { for (void *p : dev) if (p) cudaFree(p); }
    template <class T> int alloc(T **p, size_t n)
    {
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, (n ? n : 1) * sizeof(T));
        if (e != cudaSuccess) { meme_set_error("cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e)); return MEME_E_NOMEM; }
        dev.push_back(q);
        *p = (T *)q;
        return MEME_OK;
    }
};

__device__ __forceinline__ int sa_sym(const uint8_t *__restrict__ text, int64_t n, int64_t N, int64_t i)
{
    return i < n ? (int)text[i] + 1 : (i < N ? 4 : 0);
}

// first 27 symbols of every suffix of the padded text as a base-5 number
__global__ void k_sa_key0(const uint8_t *__restrict__ text, int64_t n, int64_t N, u64 *__restrict__ key, u32 *__restrict__ idx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    u64 k = 0;
#pragma unroll
    for (int r = 0; r < 27; r++) k = k * 5 + (u64)sa_sym(text, n, N, i + r);
    key[i] = k;
    idx[i] = (u32)i;
}

// head[k]: sorted key k differs from its predecessor; mark[k] = the slot of the bucket's head, for the max-scan
template <class K> __global__ void k_sa_heads(const K *__restrict__ key, int64_t m, const u32 *__restrict__ slot, uint8_t *__restrict__ head,
                                              u32 *__restrict__ mark)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const bool h = k == 0 || key[k] != key[k - 1];
    head[k] = h;
    mark[k] = h ? (slot ? slot[k] : (u32)k) : 0u;
}

__global__ void k_sa_rank(const u32 *__restrict__ suf, const u32 *__restrict__ grp, int64_t m, u32 *__restrict__ rank)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) rank[suf[k]] = grp[k];
}

// still ambiguous: the slot is not a bucket of one
__global__ void k_sa_keep(const uint8_t *__restrict__ head, int64_t m, uint8_t *__restrict__ keep)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const bool next_head = k + 1 >= m || head[k + 1];
    keep[k] = !(head[k] && next_head);
}

__global__ void k_sa_comp(const u32 *__restrict__ sa, const u32 *__restrict__ act, int64_t m, const u32 *__restrict__ rank, int64_t N, int64_t h,
                          u64 *__restrict__ comp, u32 *__restrict__ suf)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const u32 s = sa[act[k]];
    const int64_t nb = (int64_t)s + h;
    const u64 sec = nb < N ? (u64)rank[nb] + 1 : 0;
    comp[k] = (u64)rank[s] * (u64)(N + 1) + sec;
    suf[k] = s;
}

__global__ void k_sa_put(const u32 *__restrict__ act, const u32 *__restrict__ suf, int64_t m, u32 *__restrict__ sa)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) sa[act[k]] = suf[k];
}

__global__ void k_sa_iota(u32 *__restrict__ a, int64_t m)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) a[k] = (u32)k;
}

// real suffixes only (those that start inside the text), widened; the key of each: 32 bases over the padded text, cyclic,
// first base in bits 63-62 (src/Learnedindex.cpp:L481-500)
__global__ void k_sa_real(const u32 *__restrict__ sa, int64_t N, int64_t n, uint8_t *__restrict__ flag)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < N) flag[k] = sa[k] < (u32)n;
}

__global__ void k_sa_out(const u32 *__restrict__ sa_real, int64_t n, const uint8_t *__restrict__ text, int64_t N, int64_t *__restrict__ sa_out,
                         u64 *__restrict__ keys)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t s = sa_real[k];
    sa_out[k] = s;
    if (keys) {
        u64 key = 0;
        int64_t p = s;
        for (int r = 0; r < 32; r++) {
            const u64 b = p < n ? (u64)text[p] : 3ull;
            key = (key << 2) | b;
            if (++p == N) p = 0;
        }
        keys[k] = key;
    }
}

struct MaxOp { __device__ __forceinline__ u32 operator()(u32 a, u32 b) const { return a > b ? a : b; } };

int grow_tmp(void **tmp, size_t *cap, size_t need)
{
    if (need <= *cap) return MEME_OK;
    if (*tmp) cudaFree(*tmp);
    *tmp = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc(tmp, need + 1024);
    if (e != cudaSuccess) { meme_set_error("cudaMalloc(%zu bytes): %s", need, cudaGetErrorString(e)); return MEME_E_NOMEM; }
    *cap = need + 1024;
    return MEME_OK;
}

}  // namespace

#define GRID(m) (unsigned)(((m) + 255) / 256), 256

extern "C" int meme_sa_build(int device, const uint8_t *text, int64_t n, int64_t *sa_out, uint64_t *keys_out, int32_t *pad_out, int32_t *rounds_out)
{
    if (!text || !sa_out || n < 1) { meme_set_error("bad argument"); return MEME_E_ARG; }
    // padding: one more than the longest run of A or T (src/Learnedindex.cpp:L190-229)
    int64_t best = 0, run = 0;
    int prev = -1;
    for (int64_t i = 0; i < n; i++) {
        const int c = text[i];
        if (c > 3) { meme_set_error("text code %d at %lld (the indexed text has no N)", c, (long long)i); return MEME_E_ARG; }
        if ((c == 0 || c == 3) && c == prev) run++; else run = (c == 0 || c == 3) ? 1 : 0;
        prev = c;
        if (run > best) best = run;
    }
    const int64_t P = best + 1, N = n + P;
    if (N >= (1ll << 31) - 64) { meme_set_error("text of %lld bases: the 32-bit suffix ids of this builder stop at 2^31", (long long)n); return MEME_E_ARG; }
    if (pad_out) *pad_out = (int32_t)P;
    CUDA_TRY(cudaSetDevice(device));
    cudaStream_t st = 0;
    SaBuf B;
    int rc;
    uint8_t *d_text, *d_head, *d_keep;
    u64 *d_k0, *d_k1;
    u32 *d_v0, *d_v1, *d_sa, *d_rank, *d_act0, *d_act1, *d_mark, *d_grp;
    int64_t *d_cnt;
    if ((rc = B.alloc(&d_text, (size_t)n + 64))) return rc;
    if ((rc = B.alloc(&d_k0, (size_t)N)) || (rc = B.alloc(&d_k1, (size_t)N))) return rc;
    if ((rc = B.alloc(&d_v0, (size_t)N)) || (rc = B.alloc(&d_v1, (size_t)N))) return rc;
    if ((rc = B.alloc(&d_sa, (size_t)N)) || (rc = B.alloc(&d_rank, (size_t)N))) return rc;
    if ((rc = B.alloc(&d_act0, (size_t)N)) || (rc = B.alloc(&d_act1, (size_t)N))) return rc;
    if ((rc = B.alloc(&d_mark, (size_t)N)) || (rc = B.alloc(&d_grp, (size_t)N))) return rc;
    if ((rc = B.alloc(&d_head, (size_t)N + 1)) || (rc = B.alloc(&d_keep, (size_t)N + 1))) return rc;
    if ((rc = B.alloc(&d_cnt, 2))) return rc;
    void *tmp = nullptr; size_t cap = 0;
    struct TmpGuard { void **p; ~TmpGuard() { if (*p) cudaFree(*p); } } tg{&tmp};
    CUDA_TRY(cudaMemcpyAsync(d_text, text, (size_t)n, cudaMemcpyHostToDevice, st));
    // ---- round 0: all suffixes by their first 27 symbols ----
    k_sa_key0<<<GRID(N), 0, st>>>(d_text, n, N, d_k0, d_v0);
    CUDA_TRY(cudaGetLastError());
    {
        cub::DoubleBuffer<u64> dk(d_k0, d_k1);
        cub::DoubleBuffer<u32> dv(d_v0, d_v1);
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int)N, 0, 63, st);
        if ((rc = grow_tmp(&tmp, &cap, need))) return rc;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, need, dk, dv, (int)N, 0, 63, st));
        CUDA_TRY(cudaMemcpyAsync(d_sa, dv.Current(), (size_t)N * 4, cudaMemcpyDeviceToDevice, st));
        k_sa_heads<u64><<<GRID(N), 0, st>>>(dk.Current(), N, nullptr, d_head, d_mark);
        CUDA_TRY(cudaGetLastError());
    }
    int64_t m = N;
    {
        size_t need = 0;
        cub::DeviceScan::InclusiveScan(nullptr, need, d_mark, d_grp, MaxOp(), (int)N, st);
        if ((rc = grow_tmp(&tmp, &cap, need))) return rc;
        CUDA_TRY(cub::DeviceScan::InclusiveScan(tmp, need, d_mark, d_grp, MaxOp(), (int)N, st));
        k_sa_rank<<<GRID(N), 0, st>>>(d_sa, d_grp, N, d_rank);
        k_sa_keep<<<GRID(N), 0, st>>>(d_head, N, d_keep);
        k_sa_iota<<<GRID(N), 0, st>>>(d_act0, N);
        CUDA_TRY(cudaGetLastError());
        need = 0;
        cub::DeviceSelect::Flagged(nullptr, need, d_act0, d_keep, d_act1, d_cnt, (int)N, st);
        if ((rc = grow_tmp(&tmp, &cap, need))) return rc;
        CUDA_TRY(cub::DeviceSelect::Flagged(tmp, need, d_act0, d_keep, d_act1, d_cnt, (int)N, st));
        CUDA_TRY(cudaMemcpyAsync(&m, d_cnt, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    u32 *act = d_act1, *act_other = d_act0;
    int key_bits = 1;
    while ((1ull << key_bits) < (u64)(N + 1)) key_bits++;
    key_bits = 2 * key_bits < 64 ? 2 * key_bits : 64;
    int64_t h = 27;
    int rounds = 0;
    // ---- doubling rounds over the ambiguous slots only ----
    while (m > 0) {
        rounds++;
        if (rounds > 64) { meme_set_error("prefix doubling did not converge"); return MEME_E_CUDA; }
        k_sa_comp<<<GRID(m), 0, st>>>(d_sa, act, m, d_rank, N, h, d_k0, d_v0);
        CUDA_TRY(cudaGetLastError());
        cub::DoubleBuffer<u64> dk(d_k0, d_k1);
        cub::DoubleBuffer<u32> dv(d_v0, d_v1);
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int)m, 0, key_bits, st);
        if ((rc = grow_tmp(&tmp, &cap, need))) return rc;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, need, dk, dv, (int)m, 0, key_bits, st));
        k_sa_put<<<GRID(m), 0, st>>>(act, dv.Current(), m, d_sa);
        k_sa_heads<u64><<<GRID(m), 0, st>>>(dk.Current(), m, act, d_head, d_mark);
        CUDA_TRY(cudaGetLastError());
        need = 0;
        cub::DeviceScan::InclusiveScan(nullptr, need, d_mark, d_grp, MaxOp(), (int)m, st);
        if ((rc = grow_tmp(&tmp, &cap, need))) return rc;
        CUDA_TRY(cub::DeviceScan::InclusiveScan(tmp, need, d_mark, d_grp, MaxOp(), (int)m, st));
        k_sa_rank<<<GRID(m), 0, st>>>(dv.Current(), d_grp, m, d_rank);
        k_sa_keep<<<GRID(m), 0, st>>>(d_head, m, d_keep);
        CUDA_TRY(cudaGetLastError());
        need = 0;
        cub::DeviceSelect::Flagged(nullptr, need, act, d_keep, act_other, d_cnt, (int)m, st);
        if ((rc = grow_tmp(&tmp, &cap, need))) return rc;
        CUDA_TRY(cub::DeviceSelect::Flagged(tmp, need, act, d_keep, act_other, d_cnt, (int)m, st));
        CUDA_TRY(cudaMemcpyAsync(&m, d_cnt, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        u32 *t = act; act = act_other; act_other = t;
        h *= 2;
    }
    if (rounds_out) *rounds_out = rounds;
    // ---- drop the suffixes that start in the padding; widen; keys ----
    {
        k_sa_real<<<GRID(N), 0, st>>>(d_sa, N, n, d_keep);
        CUDA_TRY(cudaGetLastError());
        size_t need = 0;
        cub::DeviceSelect::Flagged(nullptr, need, d_sa, d_keep, d_v0, d_cnt, (int)N, st);
        if ((rc = grow_tmp(&tmp, &cap, need))) return rc;
        CUDA_TRY(cub::DeviceSelect::Flagged(tmp, need, d_sa, d_keep, d_v0, d_cnt, (int)N, st));
        int64_t got = 0;
        CUDA_TRY(cudaMemcpyAsync(&got, d_cnt, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (got != n) { meme_set_error("suffix array has %lld real suffixes, expected %lld", (long long)got, (long long)n); return MEME_E_CUDA; }
        int64_t *d_out = (int64_t *)d_k0;                 // N * 8 bytes, free now
        u64 *d_keys = keys_out ? d_k1 : nullptr;
        k_sa_out<<<GRID(n), 0, st>>>(d_v0, n, d_text, N, d_out, d_keys);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(sa_out, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        if (keys_out) CUDA_TRY(cudaMemcpyAsync(keys_out, d_keys, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return MEME_OK;
}
