// cuda_shim.h -- TEST INFRASTRUCTURE.  Lets g++ compile bwa-meme_b200/csrc/meme_seed_flat.cuh so that the flat seeding
// kernel's control code runs ONE LANE AT A TIME on the CPU (tests/host_sim/flat_host.cpp, tests/test_flat_host_sim.py).
// A warp of one lane: votes return the lane's predicate, barriers are no-ops, cp.async is memcpy, atomics are plain
// read-modify-writes.  Nothing here is part of the product; the product path is the CUDA build of the same source.
#pragma once
#define MEME_HOST_SIM 1
#include <cuda_runtime.h>      // vector types and the empty host definitions of __device__ / __global__ / __align__
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __noinline__
#define __noinline__ __attribute__((noinline))
#undef __forceinline__
#define __forceinline__ inline
#undef __maxnreg__
#define __maxnreg__(...)

using std::max;
using std::min;

struct SimDim { unsigned x, y, z; };
static const SimDim threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};

template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned long long __brevll(unsigned long long x)
{
    unsigned long long r = 0;
    for (int i = 0; i < 64; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
static inline double __ull2double_rn(unsigned long long x) { return (double)x; }
static inline double __longlong_as_double(long long x) { double d; memcpy(&d, &x, 8); return d; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline unsigned long long atomicExch(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = v; return o; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline int __all_sync(unsigned, int p) { return p; }
static inline int __any_sync(unsigned, int p) { return p; }

static inline unsigned atomicCAS(unsigned *p, unsigned cmp, unsigned v) { unsigned o = *p; if (o == cmp) *p = v; return o; }
static inline unsigned atomicSub(unsigned *p, unsigned v) { unsigned o = *p; *p = o - v; return o; }
static inline int atomicAdd(int *p, int v) { int o = *p; *p = o + v; return o; }
static inline int atomicSub(int *p, int v) { int o = *p; *p = o - v; return o; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }

// the device-only constructs of meme_seed_grp.cuh: a CTA of one warp of one lane
#include <cstdio>
#include <cstdlib>
#define GRP_TRACE(...) do { static int on = getenv("GRP_TRACE") != nullptr; static long cnt = 0; if (on && cnt++ < 400) fprintf(stderr, __VA_ARGS__); } while (0)
#define GRP_SMEM_DECL alignas(16) static unsigned char grp_smem[L::BYTES]
#define GRP_CP4(dst, src) memcpy((void *)(dst), (const void *)(src), 4)
#define GRP_CP8(dst, src) memcpy((void *)(dst), (const void *)(src), 8)
#define GRP_CP16(dst, src) memcpy((void *)(dst), (const void *)(src), 16)
#define GRP_CP_COMMIT() ((void)0)
#define GRP_CP_WAIT(n) ((void)0)
#define GRP_FENCE() ((void)0)
#define GRP_SLEEP(ns) ((void)0)
#define GRP_LANE() 0
#define GRP_WARP() 0
#define GRP_BALLOT(p) ((p) ? 1u : 0u)
#define GRP_SHFL(v, l) (v)
#define GRP_REDUCE_MIN(v) ((int)(v))
#define GRP_SYNCWARP() ((void)0)
#define GRP_SYNCTHREADS() ((void)0)

// the device-only constructs of meme_seed_flat.cuh
#define FLAT_SMEM_DECL static u32 smem[TPR_THREADS * 128]
typedef char *flat_stage_t;
#define FLAT_STAGE_BASE(p) (reinterpret_cast<char *>(p))
#define FLAT_CP4(dst, src) memcpy((dst), (src), 4)
#define FLAT_CP_COMMIT() ((void)0)
#define FLAT_CP_WAIT() ((void)0)
