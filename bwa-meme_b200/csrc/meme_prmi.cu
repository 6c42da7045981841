// meme_prmi.cu -- P-RMI training on the device: the model files the reference's Rust trainer writes
// for `pwl{BIT},linear,linear_spline` (RMI/rmi_lib/src/train/two_layer.rs; not buildable offline), trained from the sorted
// 64-bit keys of the suffix array.  Alignment results do not depend on the parameter values -- the last-mile search is
// self-correcting (src/LearnedIndex_seeding.cpp:L2277-2357) -- only the lookup cost does, so this is a restatement of the
// published structure (bwa-meme_b200/prmi_train.py documents it and is the CPU twin of this file), not of the Rust numerics:
//   leaf        = key >> (64 - BIT), one line per leaf through the two keys OUTSIDE the leaf (so that unseen keys stay inside)
//   dense leaf  (> 1000 keys): the leaf's line predicts one of cnt/20 sub-models, each a line of the same kind
//   error word  (below << 32) | above over the training keys, +1 slack; dense leaf: 1<<63 | first sub-model << 32 | count
// Everything is one thread per leaf / sub-model / key: segment bounds are binary searches over the sorted keys (the sub-model a
// key falls into is monotone in the key), errors are atomic maxima.  The prediction is evaluated with the same fma the lookup
// uses (learned_index_lookup, L186-210; meme_seed_flat.cuh).
#include <cub/cub.cuh>
#include <vector>
#include "meme_common.cuh"

namespace {

struct Rec { double alpha, beta; u64 err; };        // the 24-byte file record

__device__ __forceinline__ double u2d(u64 k) { return __ull2double_rn(k); }

__device__ __forceinline__ int64_t lower_bound_key(const u64 *__restrict__ keys, int64_t lo, int64_t hi, u64 x)
{
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (keys[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}

// line through (key[first-1], first-1) and (key[last+1], last+1); a leaf at either end of the key space uses 0 / 2^64
__device__ __forceinline__ void fit_line(const u64 *__restrict__ keys, int64_t n, int64_t first, int64_t last, double &alpha, double &beta)
{
    const double k_prev = first > 0 ? u2d(keys[first - 1]) : 0.0;
    const double k_next = last < n - 1 ? u2d(keys[last + 1]) : 18446744073709551616.0;
    const double t_prev = (double)(first - 1), t_next = (double)(last + 1);
    const double dk = k_next - k_prev;
    if (dk > 0) { beta = (t_next - t_prev) / dk; alpha = t_prev - beta * k_prev; }
    else { beta = 0.0; alpha = (t_prev + t_next) * 0.5; }
}

__device__ __forceinline__ int64_t predict(double alpha, double beta, u64 key, int64_t n)
{
    double p = fma(beta, u2d(key), alpha);
    if (!(p > 0.0)) p = 0.0;
    if (p > (double)(n - 1)) p = (double)(n - 1);
    return (int64_t)p;
}

__global__ void k_leaf_fit(const u64 *__restrict__ keys, int64_t n, int bits, int64_t *__restrict__ start, int64_t *__restrict__ nsub,
                           Rec *__restrict__ l2)
{
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nleaf = 1ll << bits;
    if (m >= nleaf) return;
    const int sh = 64 - bits;
    const int64_t s = lower_bound_key(keys, 0, n, (u64)m << sh);
    const int64_t e = m + 1 == nleaf ? n : lower_bound_key(keys, s, n, (u64)(m + 1) << sh);
    start[m] = s;
    if (m + 1 == nleaf) start[nleaf] = n;
    const int64_t cnt = e - s;
    nsub[m] = cnt > 1000 ? cnt / 20 : 0;
    Rec r;
    r.err = 0;
    if (cnt > 0) fit_line(keys, n, s, e - 1, r.alpha, r.beta);
    else { r.alpha = (double)(e < n - 1 ? e : n - 1); r.beta = 0.0; }        // empty leaf: the rank of the next key
    l2[m] = r;
}

// dense leaves: the leaf's line now predicts the sub-model; error word = pointer into the sub-model array
__global__ void k_leaf_dense(const u64 *__restrict__ keys, int bits, const int64_t *__restrict__ start, const int64_t *__restrict__ nsub,
                             const int64_t *__restrict__ sub_start, Rec *__restrict__ l2)
{
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (1ll << bits) || nsub[m] == 0) return;
    const int64_t s = start[m], e = start[m + 1] - 1;
    const double k0 = u2d(keys[s]), k1 = u2d(keys[e]), dk = k1 - k0;
    Rec r;
    r.beta = dk > 0 ? (double)nsub[m] / dk : 0.0;
    r.alpha = -r.beta * k0;
    r.err = (1ull << 63) | ((u64)sub_start[m] << 32) | (u64)nsub[m];
    l2[m] = r;
}

__device__ __forceinline__ int64_t sub_of(const Rec &leaf, int64_t ns, u64 key)
{
    double p = fma(leaf.beta, u2d(key), leaf.alpha);
    if (!(p > 0.0)) p = 0.0;
    int64_t s = (int64_t)p;
    return s > ns - 1 ? ns - 1 : s;
}

// one thread per sub-model: its key range inside the leaf (binary search: the sub-model index is monotone in the key), its line
__global__ void k_sub_fit(const u64 *__restrict__ keys, int64_t n, int bits, const int64_t *__restrict__ start, const int64_t *__restrict__ nsub,
                          const int64_t *__restrict__ sub_start, const int64_t *__restrict__ dense_leaf, int64_t n_dense, int64_t total_sub,
                          const Rec *__restrict__ l2, Rec *__restrict__ l1)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_sub) return;
    // the dense leaf this sub-model belongs to: last dense leaf whose first sub-model is <= g
    int64_t lo = 0, hi = n_dense;
    while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (sub_start[dense_leaf[mid]] <= g) lo = mid; else hi = mid; }
    const int64_t m = dense_leaf[lo];
    const Rec leaf = l2[m];
    const int64_t ns = nsub[m], s = g - sub_start[m], a = start[m], b = start[m + 1];
    int64_t x = a, y = b;                    // first key of the leaf with sub_of >= s
    while (x < y) { const int64_t mid = (x + y) >> 1; if (sub_of(leaf, ns, keys[mid]) < s) x = mid + 1; else y = mid; }
    const int64_t first = x;
    y = b;                                   // first key with sub_of > s
    while (x < y) { const int64_t mid = (x + y) >> 1; if (sub_of(leaf, ns, keys[mid]) <= s) x = mid + 1; else y = mid; }
    const int64_t end = x;
    Rec r;
    r.err = 0;
    if (end > first) fit_line(keys, n, first, end - 1, r.alpha, r.beta);
    else { r.alpha = (double)(end < n - 1 ? end : n - 1); r.beta = 0.0; }   // empty: rank of the next key (of this leaf, or the next leaf's first)
    l1[g] = r;
}

// one thread per key: its model's prediction error, as atomic maxima in the low / high halves of two scratch words per model
__global__ void k_errors(const u64 *__restrict__ keys, int64_t n, int bits, const int64_t *__restrict__ nsub, const int64_t *__restrict__ sub_start,
                         const Rec *__restrict__ l2, const Rec *__restrict__ l1, int *__restrict__ below2, int *__restrict__ above2,
                         int *__restrict__ below1, int *__restrict__ above1)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 key = keys[i];
    const int64_t m = (int64_t)(key >> (64 - bits));
    const Rec leaf = l2[m];
    if (nsub[m] == 0) {
        const int64_t d = predict(leaf.alpha, leaf.beta, key, n) - i;
        if (d > 0) atomicMax(&below2[m], (int)(d > 0x3ffffff0 ? 0x3ffffff0 : d));
        if (d < 0) atomicMax(&above2[m], (int)(-d > 0x3ffffff0 ? 0x3ffffff0 : -d));
    } else {
        const int64_t g = sub_start[m] + sub_of(leaf, nsub[m], key);
        const Rec r = l1[g];
        const int64_t d = predict(r.alpha, r.beta, key, n) - i;
        if (d > 0) atomicMax(&below1[g], (int)(d > 0x3ffffff0 ? 0x3ffffff0 : d));
        if (d < 0) atomicMax(&above1[g], (int)(-d > 0x3ffffff0 ? 0x3ffffff0 : -d));
    }
}

__global__ void k_err_words(Rec *__restrict__ recs, int64_t cnt, const int *__restrict__ below, const int *__restrict__ above, const int64_t *__restrict__ skip_if)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    if (skip_if && skip_if[k] != 0) return;                                   // dense leaf: the word is the sub-model pointer
    recs[k].err = ((u64)(below[k] + 1) << 32) | (u64)(above[k] + 1);          // +1: the window code underflows with a zero window at position 0
}

__global__ void k_dense_flag(const int64_t *__restrict__ nsub, int64_t nleaf, uint8_t *__restrict__ flag, int64_t *__restrict__ iota)
{
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < nleaf) { flag[m] = nsub[m] != 0; iota[m] = m; }
}

struct Bufs {
    std::vector<void *> dev;
    ~Bufs() { for (void *p : dev) if (p) cudaFree(p); }
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

}  // namespace

#define GRID(m) (unsigned)(((m) + 255) / 256), 256

// l2_out: 2^bits records of 24 bytes; l1_out: at most l1_cap records (n / 20 + 1 always suffices), *n_l1 = how many were written
// (at least one: the file is never empty).  stats_out (may be NULL): [0] dense leaves, [1] largest window of a plain leaf.
extern "C" int meme_prmi_train(int device, const uint64_t *keys, int64_t n, int32_t bits, void *l2_out, void *l1_out, int64_t l1_cap, int64_t *n_l1,
                               int64_t *stats_out)
{
    if (!keys || n < 1 || bits < 1 || bits > 30 || !l2_out || !l1_out || !n_l1 || l1_cap < n / 20 + 1) { meme_set_error("bad argument"); return MEME_E_ARG; }
    if (n >= (1ll << 31)) { meme_set_error("too many keys for the 31-bit error fields"); return MEME_E_ARG; }
    CUDA_TRY(cudaSetDevice(device));
    cudaStream_t st = 0;
    const int64_t nleaf = 1ll << bits;
    Bufs B; int rc;
    u64 *d_keys; int64_t *d_start, *d_nsub, *d_sub_start, *d_dense, *d_iota, *d_cnt; Rec *d_l2, *d_l1; int *d_b2, *d_a2, *d_b1, *d_a1; uint8_t *d_flag;
    if ((rc = B.alloc(&d_keys, (size_t)n)) || (rc = B.alloc(&d_start, (size_t)nleaf + 1)) || (rc = B.alloc(&d_nsub, (size_t)nleaf)) ||
        (rc = B.alloc(&d_sub_start, (size_t)nleaf + 1)) || (rc = B.alloc(&d_dense, (size_t)nleaf)) || (rc = B.alloc(&d_iota, (size_t)nleaf)) ||
        (rc = B.alloc(&d_cnt, 2)) || (rc = B.alloc(&d_l2, (size_t)nleaf)) || (rc = B.alloc(&d_b2, (size_t)nleaf)) || (rc = B.alloc(&d_a2, (size_t)nleaf)) ||
        (rc = B.alloc(&d_flag, (size_t)nleaf))) return rc;
    CUDA_TRY(cudaMemcpyAsync(d_keys, keys, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    k_leaf_fit<<<GRID(nleaf), 0, st>>>(d_keys, n, bits, d_start, d_nsub, d_l2);
    CUDA_TRY(cudaGetLastError());
    // first sub-model of every dense leaf; the list of dense leaves
    void *tmp = nullptr; size_t need = 0, need2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, d_nsub, d_sub_start, (int)nleaf, st);
    k_dense_flag<<<GRID(nleaf), 0, st>>>(d_nsub, nleaf, d_flag, d_iota);
    cub::DeviceSelect::Flagged(nullptr, need2, d_iota, d_flag, d_dense, d_cnt, (int)nleaf, st);
    need = need > need2 ? need : need2;
    CUDA_TRY(cudaMalloc(&tmp, need + 256));
    B.dev.push_back(tmp);
    size_t t1 = need + 256;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, t1, d_nsub, d_sub_start, (int)nleaf, st));
    t1 = need + 256;
    CUDA_TRY(cub::DeviceSelect::Flagged(tmp, t1, d_iota, d_flag, d_dense, d_cnt, (int)nleaf, st));
    int64_t n_dense = 0, last_start = 0, last_nsub = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_dense, d_cnt, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&last_start, d_sub_start + nleaf - 1, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&last_nsub, d_nsub + nleaf - 1, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const int64_t total_sub = last_start + last_nsub;
    if (total_sub > l1_cap || total_sub >= (1ll << 31)) { meme_set_error("%lld sub-models do not fit", (long long)total_sub); return MEME_E_ARG; }
    const int64_t nl1 = total_sub > 0 ? total_sub : 1;
    if ((rc = B.alloc(&d_l1, (size_t)nl1)) || (rc = B.alloc(&d_b1, (size_t)nl1)) || (rc = B.alloc(&d_a1, (size_t)nl1))) return rc;
    CUDA_TRY(cudaMemsetAsync(d_l1, 0, (size_t)nl1 * sizeof(Rec), st));
    CUDA_TRY(cudaMemsetAsync(d_b1, 0, (size_t)nl1 * 4, st));
    CUDA_TRY(cudaMemsetAsync(d_a1, 0, (size_t)nl1 * 4, st));
    CUDA_TRY(cudaMemsetAsync(d_b2, 0, (size_t)nleaf * 4, st));
    CUDA_TRY(cudaMemsetAsync(d_a2, 0, (size_t)nleaf * 4, st));
    if (total_sub > 0) {
        k_leaf_dense<<<GRID(nleaf), 0, st>>>(d_keys, bits, d_start, d_nsub, d_sub_start, d_l2);
        k_sub_fit<<<GRID(total_sub), 0, st>>>(d_keys, n, bits, d_start, d_nsub, d_sub_start, d_dense, n_dense, total_sub, d_l2, d_l1);
        CUDA_TRY(cudaGetLastError());
    }
    k_errors<<<GRID(n), 0, st>>>(d_keys, n, bits, d_nsub, d_sub_start, d_l2, d_l1, d_b2, d_a2, d_b1, d_a1);
    k_err_words<<<GRID(nleaf), 0, st>>>(d_l2, nleaf, d_b2, d_a2, d_nsub);
    if (total_sub > 0) k_err_words<<<GRID(total_sub), 0, st>>>(d_l1, total_sub, d_b1, d_a1, nullptr);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(l2_out, d_l2, (size_t)nleaf * sizeof(Rec), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(l1_out, d_l1, (size_t)nl1 * sizeof(Rec), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *n_l1 = nl1;
    if (stats_out) {
        stats_out[0] = n_dense;
        std::vector<int> b((size_t)nleaf), a((size_t)nleaf);
        CUDA_TRY(cudaMemcpy(b.data(), d_b2, (size_t)nleaf * 4, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(a.data(), d_a2, (size_t)nleaf * 4, cudaMemcpyDeviceToHost));
        int64_t mx = 0;
        for (int64_t m = 0; m < nleaf; m++) mx = b[(size_t)m] + a[(size_t)m] > mx ? b[(size_t)m] + a[(size_t)m] : mx;
        stats_out[1] = mx;
    }
    return MEME_OK;
}
