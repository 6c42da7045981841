// meme_seed_task.cuh -- declarations shared by the thread-per-task seeding kernels (meme_seed_tpr.cu,
// meme_seed_flat.cuh) and by the host-side single-lane simulation of the flat kernel (tests/host_sim/): task
// records, request kinds, control-block slots and the kernel argument block.
#pragma once
#include "meme_common.cuh"

namespace tpr {

#ifndef TPR_THREADS
#define TPR_THREADS 256
#endif
#ifndef FLAT_MINB
#define FLAT_MINB 2                     // resident CTAs per SM the flat kernel is compiled for (registers = 64 K / threads)
#endif
#define TPR_RW 10                       // 64-bit words of a packed read (MEME_MAX_READ_LEN / 32)
#define TPR_SW 31                       // shared 32-bit words per thread: 20 read + 10 N mask + 1 pad (odd stride)

struct PoolRec { u64 lo; u32 cnt; u32 pad; };           // sort value: SA interval
struct __align__(16) Task2 { u32 read, se, cnt, full; u64 lo; u64 pad; };   // reseed task (32 B): SMEM [start|end<<16], hit count, hint

enum { M_NONE = 0, M_BLOCK, M_PAC, M_RMI, M_LD, M_TASK, M_T2REC };
enum { LD_POS = 0, LD_ISA = 1 };
enum { ST_RMI = 0, ST_EXACT = 1, ST_NEAR = 2 };
enum { C_WORK = 0, C_POOL = 1, C_T2 = 2, C_POOL_OVF = 3, C_T2_OVF = 4, C_SEARCH = 8, C_BLOCK = 9, C_KARY = 10, C_EXT = 11,
       C_PAC = 12, C_ITER = 13, C_PASS_CTL = 14, C_PASS_MEM = 15, C_LANES_CTL = 16, C_LANES_MEM = 17, C_NCTRL = 24 };

struct Args {
    const u64 *packed; const u32 *nmask; const int64_t *read_off;
    int64_t n_tasks; int kind;
    const Task2 *t2_in; Task2 *t2_out; u64 t2_cap;
    u64 *pool_keys; PoolRec *pool_vals; u64 pool_cap;
    u32 *rd_nsmem; u64 *rd_nhit; uint8_t *rd_full; u64 *rd_maxref;
    u64 *ctrl;
    meme_seed_opt_t opt;
    int grp_th;                         // regrouped kernel: a pool with at least this many tasks is taken at once
};

__device__ __forceinline__ u64 revcomp32(u64 w)
{   // 32 bases (first base in the LSBs) -> reverse complement, first base in the LSBs
    u64 y = __brevll(w);
    y = ((y & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((y & 0x5555555555555555ull) << 1);
    return ~y;
}

}  // namespace tpr
