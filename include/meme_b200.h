/* meme_b200.h -- C ABI of the B200-native `bwa-meme mem -7` hot path.
 *
 * Drop-in boundary.  The reference has no plugin mechanism; its hot path is cut at
 * two source-level seams, and every entry point below names the reference interface it replaces
 * (paths relative to the reference tree, kaist-ina/BWA-MEME):
 *
 *   index load + runtime build   memoryAllocLearned()                      src/fastmap.cpp:L351-638
 *                                 learned_index_load()                      src/LearnedIndex_seeding.cpp:L74-122
 *   seeding (kernel A)            mem_kernel1_core_Learned() up to the sort src/bwamem.cpp:L1230-1397
 *                                 = Learned_getSMEMsAllPosOneThread +       src/LearnedIndex_seeding.h:L235-260
 *                                   Learned_bwtSeedStrategyAllPosOneThread{,_mem_tradeoff}
 *   seed extension (kernel B)     BandedPairWiseSW::getScores8/16,          src/bandedSWA.h:L118-135, L257-297
 *                                 scalarBandedSWAWrapper                    src/bandedSWA.cpp:L242-259
 *
 * Plain pointers and sizes only; no C++/torch types.  All functions return 0 on success or a
 * negative MEME_E_* code; nothing calls exit().  A handle is single-owner (one host thread at a
 * time), like the reference's worker_t during step 1 of its pipeline (src/fastmap.cpp:L877-888).
 * Result buffers returned by the library are pinned host memory owned by the handle and stay valid
 * until the next call on the same handle.
 */
#ifndef MEME_B200_H
#define MEME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MEME_OK              0
#define MEME_E_CUDA         -1   /* a CUDA runtime call failed (see meme_last_error) */
#define MEME_E_IO           -2   /* index file missing / short / inconsistent */
#define MEME_E_ARG          -3   /* bad argument (NULL, negative count, read longer than MEME_MAX_READ_LEN) */
#define MEME_E_NOMEM        -4   /* host or device allocation failed */
#define MEME_E_OVERFLOW     -5   /* (reserved; the seeding kernels have no per-read SMEM limit) */
#define MEME_E_NOGPU        -6   /* no CUDA device: this library has no CPU path */

/* The reference asserts len <= ERT_MAX_READ_LEN (301) on the -7 path, src/bwamem.cpp:L1276. */
#define MEME_MAX_READ_LEN          320
#define MEME_MAX_SMEMS_PER_READ    192

typedef struct meme_index meme_index_t;
typedef struct meme_bsw   meme_bsw_t;

/* Same 24-byte layout as the reference's mem_tl (src/LearnedIndex_seeding.h:L121-127). */
typedef struct {
    int32_t  start;         /* SMEM start in the read                                   */
    int32_t  end;           /* SMEM end (exclusive)                                     */
    int32_t  hitbeg;        /* index of the first hit in this read's hit list           */
    int32_t  hitcount;      /* size of the SA interval (before any max_occ sampling)    */
    uint64_t cache_refpos;  /* reference position of the first hit (SA order)           */
} meme_smem_t;

/* mem_opt_t fields the path reads (src/bwamem.cpp:L126-162 defaults in brackets). */
typedef struct {
    int32_t min_seed_len;   /* [19]                                                       */
    int32_t split_len;      /* (int)(min_seed_len*split_factor+.499) [28]                 */
    int32_t split_width;    /* [10]                                                       */
    int32_t max_mem_intv;   /* [20]; 0 disables the third round                           */
    int32_t rounds;         /* 3 = production; 1/2 mirror the seeding harness `steps`     */
    int32_t max_occ;        /* 0: materialise every hit; >0: only the hits mem_chain_Learned
                               samples (k = 0, step, 2*step.. < hitcount, at most max_occ,
                               step = hitcount > max_occ ? hitcount / max_occ : 1,
                               src/bwamem.cpp:L1154-1155)                                 */
} meme_seed_opt_t;

typedef struct {
    int64_t      n_reads;
    int64_t      n_smems;
    int64_t      n_hits;
    const meme_smem_t *smems;     /* n_smems, grouped by read, sorted by (start,end) within a read */
    const int64_t     *smem_off;  /* n_reads+1                                                     */
    const int64_t     *hit_off;   /* n_reads+1; hits of read r = hits[hit_off[r] .. hit_off[r+1])   */
    const uint64_t    *hits;      /* reference positions in [0, 2*l_pac), SA order within an SMEM   */
} meme_seed_result_t;

/* ---- index ------------------------------------------------------------------------------- */

/* Reads <prefix>.0123, <prefix>.pos_packed, <prefix>.suffixarray_uint64_L{1,2}_PARAMETERS, uploads
 * them to `device`, and builds the 64-bit inline keys and (mode 3) the inverse suffix array on the
 * GPU.  mode: 1 = positions only, 2 = + inline keys, 3 = + inverse SA (Makefile:L40-46). */
int meme_index_load(const char *prefix, int mode, int device, meme_index_t **out);

/* The index replicated on ndev devices, loaded concurrently; out receives ndev handles (all or none).  Reads shard
 * across GPUs, the index does not. */
int meme_index_load_devices(const char *prefix, int mode, const int *devices, int ndev, meme_index_t **out);

/* Same, from host buffers (what the files would contain). l1/l2 are 24-byte records. */
int meme_index_from_host(const uint8_t *text0123, uint64_t sa_num, const uint8_t *pos_packed,
                         const void *l1_params, uint64_t n_l1, const void *l2_params, uint64_t n_l2,
                         int mode, int device, meme_index_t **out);

void     meme_index_free(meme_index_t *ix);
uint64_t meme_index_sa_num(const meme_index_t *ix);
uint64_t meme_index_device_bytes(const meme_index_t *ix);
/* Use an existing CUDA stream (a cudaStream_t passed as void*); NULL restores the handle's own. */
int      meme_index_set_stream(meme_index_t *ix, void *cuda_stream);

/* P-RMI inference for a batch of 64-bit keys (first base in the most significant bits), as learned_index_lookup
 * (src/LearnedIndex_seeding.cpp:L186-210) computes it: pred = clamped position estimate, err = the error word of the
 * model that answered (window below = (err>>32)&0x3fffffff, above = err&0x7fffffff, L2145-2146). */
int meme_rmi_lookup(meme_index_t *ix, const uint64_t *keys, int64_t n, uint64_t *pred, uint64_t *err);

/* Copies index arrays back (tests): any pointer may be NULL.  keys/pos/isa have sa_num entries. */
int meme_index_download(const meme_index_t *ix, uint64_t *keys, uint64_t *pos, uint64_t *isa);

/* ---- kernel A: seeding --------------------------------------------------------------------- */

/* reads: base codes 0..3, 4 = N (nst_nt4_table output), concatenated; read r = codes[off[r]..off[r+1]).
 * One call = upload + run + download. */
int meme_seed_batch(meme_index_t *ix, const uint8_t *codes, const int64_t *read_off, int64_t n_reads,
                    const meme_seed_opt_t *opt, meme_seed_result_t *out);

/* Multi-GPU form: the chunk is cut into ndev contiguous read ranges (both mates of a pair on one device when
 * pair_aligned != 0), every range runs on its own device concurrently; outs[d] describes reads
 * shard_begin[d] .. shard_begin[d+1] (shard_begin has ndev+1 entries).  No device-to-device exchange: each device
 * delivers into its own pinned host arrays. */
int meme_seed_batch_multi(meme_index_t *const *ixs, int ndev, const uint8_t *codes, const int64_t *read_off, int64_t n_reads,
                          int pair_aligned, const meme_seed_opt_t *opt, meme_seed_result_t *outs, int64_t *shard_begin);

/* One throw-away batch of n_reads reads of read_len bases sampled from the indexed text: loads the kernels and sizes
 * every pool of the batch path (device pools, pinned result arrays) for chunks of that size, so that a caller's first
 * real chunk does not pay for it.  Call once after meme_index_load with the expected chunk size. */
int meme_seed_warmup(meme_index_t *ix, int64_t n_reads, int32_t read_len);

/* The same in three steps, so that a caller (and bench.py) can keep inputs resident in HBM. */
int meme_seed_upload(meme_index_t *ix, const uint8_t *codes, const int64_t *read_off, int64_t n_reads);
int meme_seed_run(meme_index_t *ix, const meme_seed_opt_t *opt);          /* synchronises between the task kinds; the final gather is asynchronous */
int meme_seed_download(meme_index_t *ix, meme_seed_result_t *out);        /* synchronises               */
/* Device-to-device copies of the lists of the last meme_seed_run (any pointer may be NULL; sizes from
 * meme_seed_totals): for callers that keep the results in HBM (the NCCL gather of the multi-process bench). */
int meme_seed_copy_device(meme_index_t *ix, void *d_smems, void *d_hits, void *d_smem_off, void *d_hit_off);
/* Totals of the last run without copying the lists (synchronises). */
int meme_seed_totals(meme_index_t *ix, int64_t *n_smems, int64_t *n_hits);
/* Work counters of the last run: [0] searches, [1] contiguous SA block probes, [2] k-ary narrowing rounds,
 * [3] extension rounds, [4] packed-reference rounds, [5] kernel launches since the handle was created, [6] reseed tasks,
 * [7] memory round trips.  [0..4] and [7] are counted only when MEME_SEED_STATS=1 (they cost registers in the hot loop). */
int meme_seed_counters(meme_index_t *ix, uint64_t counters[8]);
/* CUDA-event durations (ms) of the last meme_seed_run on its stream: [0] the seeding kernel,
 * [1] offset scans, [2] the hit-materialisation kernel. */
int meme_seed_last_ms(meme_index_t *ix, float ms[3]);

/* ---- chaining: seeds -> chains -> filtered chains, on the device ---------------- */

/* What the path reads of bntann1_t (src/bntseq.h:L41-49). */
typedef struct {
    int64_t offset;
    int32_t len;
    int32_t is_alt;
} meme_contig_t;

/* mem_opt_t fields read by mem_chain_Learned / test_and_merge / mem_chain_flt (defaults of src/bwamem.cpp:L126-162). */
typedef struct {
    int32_t w;                  /* [100] band width: seeds join a chain within w of its diagonal   */
    int32_t max_chain_gap;      /* [10000]                                                         */
    int32_t max_occ;            /* [500]; must equal the max_occ the seeding run sampled with       */
    int32_t min_seed_len;       /* [19]                                                            */
    int32_t min_chain_weight;   /* [0]                                                             */
    int32_t max_chain_extend;   /* [1<<30]                                                         */
    float   mask_level;         /* [0.50]                                                          */
    float   drop_ratio;         /* [0.50]                                                          */
} meme_chain_opt_t;

/* A seed of a chain: the fields of mem_seed_t that chaining sets (src/bwamem.h:L124-127); score == len. */
typedef struct {
    int64_t rbeg;
    int32_t qbeg, len;
} meme_chain_seed_t;

/* A chain as mem_chain_flt leaves it (mem_chain_t, src/bwamem.h:L132-139). */
typedef struct {
    int64_t pos;                /* rbeg of the first seed (the B-tree key)                         */
    int64_t seed_beg;           /* its seeds: seeds[seed_beg .. seed_beg + n)                      */
    int32_t n, rid;
    int32_t w, first;           /* weight; first shadowed chain (index in the sorted list) or -1   */
    int32_t kept, is_alt;
} meme_chain_t;

typedef struct {
    int64_t n_reads, n_chains, n_seeds;
    const int64_t           *chain_off;   /* n_reads+1: chains of read r = chains[chain_off[r] .. chain_off[r+1]), in the
                                             order mem_kernel1_core_Learned leaves them in chain_ar[r]                 */
    const float             *frac_rep;    /* n_reads                                                               */
    const meme_chain_t      *chains;
    const meme_chain_seed_t *seeds;
} meme_chain_result_t;

/* The contig table (bns->anns, bns->l_pac): seeds bridging two contigs or the forward/reverse boundary are dropped
 * (bns_intv2rid, src/bntseq.cpp:L408-416), chains never cross contigs, ALT contigs are treated as the filter treats them. */
int meme_chain_set_contigs(meme_index_t *ix, const meme_contig_t *contigs, int32_t n_contigs, int64_t l_pac);
/* mem_chain_Learned + mem_chain_flt (src/bwamem.cpp:L1122-1204, L599-717; mem_flt_chained_seeds, L565-597, does nothing
 * for reads shorter than about 700 bases and is outside this path's 320-base limit) for every read of the last
 * meme_seed_run, from its device-resident SMEM and hit lists.  That run must have sampled hits with the same max_occ. */
int meme_chain_run(meme_index_t *ix, const meme_chain_opt_t *opt);
int meme_chain_download(meme_index_t *ix, meme_chain_result_t *out);        /* synchronises; pinned, owned by the handle */
/* CUDA-event duration (ms) of the last meme_chain_run. */
int meme_chain_last_ms(meme_index_t *ix, float *ms);
/* After meme_seed_warmup: chains that batch once, so that the first real chunk does not pay for module load and pool sizing. */
int meme_chain_warmup(meme_index_t *ix);
/* Seeding + chaining of a chunk across ndev devices (same sharding as meme_seed_batch_multi); only chains and their seeds
 * travel to the host -- the SMEM and hit lists stay in HBM. */
int meme_chain_batch_multi(meme_index_t *const *ixs, int ndev, const uint8_t *codes, const int64_t *read_off, int64_t n_reads,
                           int pair_aligned, const meme_seed_opt_t *seed_opt, const meme_chain_opt_t *chain_opt,
                           meme_chain_result_t *outs, int64_t *shard_begin);

/* ---- global alignment of the SAM stage: score, CIGAR, NM, MD of every region -------- */

/* One region: what mem_reg2aln hands to bwa_gen_cigar2 (src/bwamem.cpp:L2326-2343). */
typedef struct {
    int64_t rb, re;         /* the region on the fwd+RC text                                            */
    int64_t q_off;          /* first base of the query segment in `codes` (read start + qb)             */
    int32_t qlen;           /* qe - qb                                                                  */
    int32_t w;              /* band to start with: max of the two infer_bw values, limited as L2338 does */
    int32_t truesc;         /* the region's local score (the loop stops once the global score is within a of it) */
    int32_t pad;
} meme_cigar_task_t;

typedef struct {
    int8_t  mat[25];        /* opt->mat (bwa_fill_scmat)                                                */
    int8_t  pad[3];
    int32_t o_del, e_del, o_ins, e_ins;
    int32_t w;              /* opt->w: the loop never goes beyond a band of 4w                          */
    int32_t a;              /* opt->a                                                                   */
} meme_cigar_opt_t;

typedef struct {
    int32_t score;          /* global alignment score of the final try                                  */
    int32_t nm;             /* NM                                                                       */
    int32_t n_cigar;        /* number of CIGAR operations; -1: rejected as bwa_gen_cigar2 rejects it (src/bwa.cpp:L285-287) */
    int32_t md_len;         /* bytes of the MD string including its NUL                                 */
    int64_t cigar_off;      /* cigars[cigar_off .. + n_cigar): len << 4 | op, op 0 = M, 1 = I, 2 = D     */
    int64_t md_off;         /* md + md_off: the MD string                                               */
} meme_cigar_rec_t;

typedef struct {
    int64_t n_tasks, n_ops, n_md;
    const meme_cigar_rec_t *recs;
    const uint32_t         *cigars;
    const char             *md;
} meme_cigar_result_t;

/* For every task: the band-doubling loop of mem_reg2aln (src/bwamem.cpp:L2339-2348) around bwa_gen_cigar2
 * (src/bwa.cpp:L274-361) = ksw_global2 (src/ksw.cpp:L560-670) + NM/MD.  The reference bases are read from the indexed text
 * in HBM (l_pac = bns->l_pac), the query bases from `codes` (0..3, 4 = N).  Results are pinned host arrays owned by the
 * handle, valid until the next call. */
int meme_cigar_batch(meme_index_t *ix, const meme_cigar_task_t *tasks, int64_t n_tasks, const uint8_t *codes, int64_t n_codes,
                     int64_t l_pac, const meme_cigar_opt_t *opt, meme_cigar_result_t *out);
/* Kernels loaded and every array sized for about n_regions regions of qlen-base reads per call, ahead of the first timed
 * chunk: one slice of dummy regions with worst-case bands runs through every kernel, the backtrack scratch is sized for a
 * slice in which half of the warps hold full-width matrices (3.6 GB at 150 bases), the pinned result arrays for the whole
 * chunk.  (The reference has no counterpart: its first chunk costs what the others cost.) */
int meme_cigar_warmup(meme_index_t *ix, int64_t n_regions, int qlen);
/* CUDA-event duration (ms) of the kernels of the last meme_cigar_batch, and the DP cells they evaluated. */
int meme_cigar_last(meme_index_t *ix, float *ms, uint64_t *cells);

/* ---- kernel B: banded Smith-Waterman seed extension ------------------------------------------ */

/* Byte-identical to the reference's SeqPair (src/bandedSWA.h:L90-99). */
typedef struct {
    int32_t idr, idq, id;
    int32_t len1, len2;     /* len1: reference length, len2: query length */
    int32_t h0;
    int32_t seqid, regid;
    int32_t score, tle, gtle, qle;
    int32_t gscore, max_off;
} meme_seqpair_t;

/* Constructor arguments of BandedPairWiseSW (src/bandedSWA.h:L118-123). */
typedef struct {
    int32_t o_del, e_del, o_ins, e_ins, zdrop, end_bonus;
    int32_t a, b;           /* match score, mismatch penalty (positive); N scores -1 (bwa_fill_scmat) */
} meme_bsw_opt_t;

int  meme_bsw_create(int device, meme_bsw_t **out);
void meme_bsw_free(meme_bsw_t *b);
int  meme_bsw_set_stream(meme_bsw_t *b, void *cuda_stream);

/* Fills score, tle, gtle, qle, gscore, max_off of every pair, exactly as
 * scalarBandedSWA(len2, qer+idq, len1, ref+idr, w, h0, ...) does.  Sequences are codes 0..4.
 * No pre-sorting or SIMD-width padding is required (cf. src/bwamem.cpp:L2529-2568). */
int meme_bsw_batch(meme_bsw_t *b, meme_seqpair_t *pairs, int64_t n_pairs,
                   const uint8_t *ref, int64_t n_ref, const uint8_t *qer, int64_t n_qer,
                   int32_t w, const meme_bsw_opt_t *opt);

/* The same with the band retries of the extension driver decided on the device (replaces the MAX_BAND_TRY loops
 * around getScores8/16/scalar in mem_chain2aln_across_reads_V2, src/bwamem.cpp:L2968-3376): every pair is run with
 * w, 2w, ... (at most `tries` widths) until its result is final -- score == prev, or max_off < 3w/4, or the last try --
 * where prev is -1 (left extensions: the region has no score yet) or the pair's h0 (right extensions, prev_is_h0 = 1:
 * the region's score after the left extension).  w_used (may be NULL) receives the width of the final try. */
int meme_bsw_batch_ext(meme_bsw_t *b, meme_seqpair_t *pairs, int64_t n_pairs,
                       const uint8_t *ref, int64_t n_ref, const uint8_t *qer, int64_t n_qer,
                       int32_t w, int32_t tries, int32_t prev_is_h0, const meme_bsw_opt_t *opt, int32_t *w_used);

/* Three-step variant (inputs stay in HBM between run calls).  The match score decides which pairs fit the 13-bit
 * cells of the main kernel, and the classes are built at upload: call meme_bsw_set_match first when it is not 1. */
int meme_bsw_set_match(meme_bsw_t *b, int32_t a);
int meme_bsw_upload(meme_bsw_t *b, const meme_seqpair_t *pairs, int64_t n_pairs,
                    const uint8_t *ref, int64_t n_ref, const uint8_t *qer, int64_t n_qer);
int meme_bsw_run(meme_bsw_t *b, int32_t w, const meme_bsw_opt_t *opt);             /* asynchronous on the stream */
int meme_bsw_run_ext(meme_bsw_t *b, int32_t w, int32_t tries, int32_t prev_is_h0, const meme_bsw_opt_t *opt);
int meme_bsw_download(meme_bsw_t *b, meme_seqpair_t *pairs);                       /* synchronises */
int meme_bsw_download_w(meme_bsw_t *b, int32_t *w_used);
/* A batch assembled by several host threads: nseg segments, each with its own pair array and sequence buffers
 * (idr/idq relative to the segment's buffers).  Replaces the per-thread seqPairArray / seqBuf* of the reference's
 * mem_cache (src/bwamem.cpp:L2579-2590) when the whole chunk is extended in one call. */
int meme_bsw_upload_segments(meme_bsw_t *b, int32_t nseg, const meme_seqpair_t *const *pairs, const int64_t *n_pairs,
                             const uint8_t *const *ref, const int64_t *n_ref, const uint8_t *const *qer, const int64_t *n_qer);
int meme_bsw_download_segments(meme_bsw_t *b, int32_t nseg, meme_seqpair_t *const *pairs, const int64_t *n_pairs,
                               int32_t *const *w_used);
/* Reference by position: when the reference side of the pairs IS the indexed text that already sits in HBM (the
 * extension driver reads it from ref_string = the .0123 file, src/bwamem.cpp:L2670-2681), nothing of it needs uploading.
 * After meme_bsw_attach_text(b, ix) (same device), a batch uploaded with ref == NULL / n_ref == 0 in every segment is
 * read as: (id, idr) of a pair = position in the fwd+RC text (idr = low 32 bits, id = bits 32..62), id < 0 = the pair
 * walks the text backwards from that position (left extensions).  ix == NULL detaches. */
int meme_bsw_attach_text(meme_bsw_t *b, const meme_index_t *ix);
/* DP cells evaluated by the last run (band-limited, retries included), for GCUPS accounting. */
int meme_bsw_cells(meme_bsw_t *b, uint64_t *cells);

/* ---- mate rescue: batched local Smith-Waterman --------------------- */

/* Byte-identical to the reference's kswr_t (src/ksw.h:L45-50). */
typedef struct {
    int32_t score, te, qe, score2, te2, tb, qb;
} meme_kswr_t;

/* Constructor arguments of kswv (src/kswv.h, src/bwamem_pair.cpp:L750): gap penalties, match score, mismatch penalty (positive). */
typedef struct {
    int32_t o_del, e_del, o_ins, e_ins, a, b;
} meme_matesw_opt_t;

/* Replaces sort_classify + mem_sam_pe_batch (src/bwamem.cpp:L1797-1826, src/bwamem_pair.cpp:L719-817: kswv::getScores8/16
 * phase 0, selection of the pairs that reach their minimum score, reversal, phase 1) for a batch assembled by several host
 * threads: nseg segments, each with its own SeqPair array and sequence buffers as mem_matesw_batch_pre fills them
 * (src/bwamem_pair.cpp:L1060-1222: idr / idq relative to the segment's buffers, len1 = window length, len2 = mate length,
 * h0 = the KSW_X* flags | minimum score; sequences are codes 0..4 and are NOT modified).  aln[s][k] receives the kswr_t of
 * pairs[s][k], i.e. what the reference leaves in aln[pairs[k].regid].  No sorting, classing or SIMD-width padding is
 * required from the caller.  Limits: mate length <= 496 (255 in the 8-bit class, h0 & KSW_XBYTE), window <= 32000. */
int meme_matesw_batch(meme_index_t *ix, int32_t nseg, const meme_seqpair_t *const *pairs, const int64_t *n_pairs,
                      const uint8_t *const *ref, const int64_t *n_ref, const uint8_t *const *qer, const int64_t *n_qer,
                      const meme_matesw_opt_t *opt, meme_kswr_t *const *aln);
/* CUDA-event duration (ms) of the kernels of the last meme_matesw_batch, and the DP cells they evaluated. */
int meme_matesw_last(meme_index_t *ix, float *ms, uint64_t *cells);
/* Kernel loaded and buffers sized for about n_pairs pairs of len1 x len2 ahead of the first timed chunk. */
int meme_matesw_warmup(meme_index_t *ix, int64_t n_pairs, int len1, int len2);

/* Pinned host memory for callers that assemble large batches (the chunk-level extension driver of the glue). */
int  meme_pinned_alloc(size_t bytes, void **out);
void meme_pinned_free(void *p);

/* ---- index pipeline ---------------------------------------------- */

/* Suffix array of the fwd+RC text (codes 0..3, n bases) in the reference's order, and optionally the 64-bit key of every
 * suffix: replaces the sais call and the key loop of the reference's index builder (buildSAandLEP / build_learned_index,
 * src/Learnedindex.cpp:L134-555).  sa_out[n] is what the reference packs into <prefix>.pos_packed, keys_out[n] (may be NULL)
 * what it writes to <prefix>.suffixarray_uint64 after the count; *pad_out = the number of `T`s the reference pads the text with
 * (L190-229), *rounds_out = doubling rounds after the first sort.  Host pointers; prefix doubling with device radix sorts.
 * Limit: n + pad < 2^31 (chr1 scale; 32-bit suffix ids). */
int meme_sa_build(int device, const uint8_t *text, int64_t n, int64_t *sa_out, uint64_t *keys_out, int32_t *pad_out, int32_t *rounds_out);

/* P-RMI training on the device: the `pwl{BIT},linear,linear_spline` model files of the reference's Rust trainer
 * (RMI/rmi_lib/src/train/two_layer.rs; build_rmis_dna.sh) from the n sorted keys of meme_sa_build.  l2_out: 2^bits records of
 * 24 bytes {f64 alpha, f64 beta, u64 err} (<prefix>.suffixarray_uint64_L2_PARAMETERS), l1_out: the sub-models of dense leaves
 * (_L1_PARAMETERS; l1_cap >= n / 20 + 1 records, *n_l1 = records written, at least one); _L0_PARAMETERS is the single word `bits`.
 * Format-compatible with the reference loader (learned_index_load, src/LearnedIndex_seeding.cpp:L74-122); the values are this
 * repo's own (the last-mile search is self-correcting, so alignment results do not depend on them).  stats_out (may be NULL):
 * [0] dense leaves, [1] widest error window of a plain leaf. */
int meme_prmi_train(int device, const uint64_t *keys, int64_t n, int32_t bits, void *l2_out, void *l1_out, int64_t l1_cap, int64_t *n_l1,
                    int64_t *stats_out);

/* ---- misc ------------------------------------------------------------------------------------ */
const char *meme_last_error(void);     /* thread-local text of the last failure */
int         meme_device_count(void);
const char *meme_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MEME_B200_H */
