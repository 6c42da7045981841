// meme_common.cuh -- shared declarations of the sm_100a library behind include/meme_b200.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/meme_b200.h"

typedef unsigned long long u64;
typedef unsigned int u32;

void meme_set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            meme_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return MEME_E_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

// P-RMI leaf / sub-model record re-laid-out to one 32-byte sector (file records are 24 bytes and
// straddle sectors half of the time; layout is not part of the contract, only results are).
struct __align__(32) RmiRec {
    double alpha, beta;
    u64 err;
    u64 pad;
};

// Device view of the index (all pointers are device memory).
struct IndexView {
    u64 sa_num;
    int mode;              // 1: no inline keys, 2: keys, 3: keys + inverse SA
    int pos_wide;          // sa_num >= 2^32: positions carry a 5th byte in pos_hi / isa_hi
    const u64 *keys;       // [sa_num] first 32 bases of suffix i, base r in bits 2r..2r+1, padded with 3
    const u32 *pos_lo;     // [sa_num]
    const uint8_t *pos_hi; // [sa_num] or null
    const u32 *isa_lo;     // [sa_num] inverse suffix array (mode 3)
    const uint8_t *isa_hi;
    const u64 *pac;        // 2-bit text, fwd+RC, base k of word w at bits 2(k&31); 2 zero words of slack
    const RmiRec *l2;      // 2^bits leaves
    const RmiRec *l1;      // partial third layer
    u64 n_l1, n_l2;
    int bit_shift;         // 64 - bits
};

// second set of input/output buffers of the pipelined batch path (meme_seed.cu: seed_batch_pipelined)
struct SeedAlt {
    uint8_t *d_codes; size_t cap_codes;
    int64_t *d_read_off, *d_smem_off, *d_hit_off; size_t cap_reads;
    meme_smem_t *d_smems; size_t cap_smems;
    u64 *d_hits; size_t cap_hits;
};

struct meme_index {
    int device;
    IndexView v;
    // owned device buffers
    u64 *d_keys; u32 *d_pos_lo; uint8_t *d_pos_hi; u32 *d_isa_lo; uint8_t *d_isa_hi; u64 *d_pac;
    RmiRec *d_l1, *d_l2;
    u64 device_bytes;
    cudaStream_t own_stream, stream;
    // ---- seeding workspace (grown on demand) ----
    uint8_t *d_codes; size_t cap_codes;
    int64_t *d_read_off; size_t cap_reads;
    int64_t n_reads; int64_t n_codes; int max_read_len;
    // per-read: raw pool offset, smem count, hit count (after sampling)
    u64 *d_rd_pool_off; u32 *d_rd_nsmem; u64 *d_rd_nhit; u32 *d_redo;
    int64_t *d_smem_off, *d_hit_off;      // n_reads+1 exclusive scans
    // raw SMEM pool written by the seed kernel: {start,end,cnt,lo}
    void *d_pool; size_t cap_pool;        // in records
    u64 *d_ctrl;                          // [0] work counter, [1] pool bump, [2] overflow flag, [8..] counters
    void *d_scan_tmp; size_t cap_scan_tmp;
    meme_smem_t *d_smems; size_t cap_smems;
    u64 *d_hits; size_t cap_hits;
    // pinned host result buffers
    meme_smem_t *h_smems; size_t hcap_smems;
    u64 *h_hits; size_t hcap_hits;
    int64_t *h_smem_off, *h_hit_off; size_t hcap_off;
    u64 *h_ctrl;
    meme_seed_opt_t last_opt;
    int ran;
    u64 launches;
    // ---- thread-per-task seeding path (meme_seed_tpr.cu) ----
    u64 *d_packed; size_t cap_packed; u32 *d_nmask; size_t cap_nmask;
    uint8_t *d_rd_full; size_t cap_rd_full; u64 *d_rd_maxref; size_t cap_rd_maxref;
    u64 *d_pool_keys; size_t cap_pool_keys; void *d_pool_vals; size_t cap_pool_vals;
    u64 *d_pool_keys2; size_t cap_pool_keys2; void *d_pool_vals2; size_t cap_pool_vals2;
    void *d_t2; size_t cap_t2;
    cudaEvent_t ev[6];                    // k_seed start/stop, offsets start/stop, k_gather start/stop
    int ev_ready;
    // ---- pipelined batch ----
    SeedAlt alt;
    cudaStream_t copy_stream;
    cudaEvent_t pe[6];                    // per buffer set: input arrived [0,1], run finished [2,3], lists drained [4,5]
    int piped; int64_t piped_reads;
    // ---- chaining (meme_chain.cu) ----
    struct ChainWork *chain;
    // ---- global alignment of the SAM stage (meme_cigar.cu) ----
    struct CigarWork *cigar;
    // ---- mate-rescue Smith-Waterman (meme_matesw.cu) ----
    struct MateswWork *matesw;
};
void meme_chain_free(struct ChainWork *c);
void meme_cigar_free(struct CigarWork *c);
void meme_matesw_free(struct MateswWork *c);

int meme_grow_device(void **p, size_t *cap, size_t need, size_t elem, cudaStream_t s);
int meme_grow_pinned(void **p, size_t *cap, size_t need, size_t elem);
