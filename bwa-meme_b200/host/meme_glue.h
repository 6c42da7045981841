// meme_glue.h -- host-side C++ mirror of the reference's hot-path interface over the C ABI.
//
// These are the four calls a maintainer of kaist-ina/BWA-MEME adds to route `bwa-meme mem -7` through
// libmeme_b200.so (INTEGRATION.md shows the diff; integration/apply_patch.py applies it to a scratch copy):
//
//   meme_glue_set_prefix   in memoryAllocLearned()            src/fastmap.cpp:L351   (index prefix)
//   meme_glue_seed_chunk   in mem_process_seqs(), before kt_for(worker_bwt)  src/bwamem.cpp:L1941
//   meme_glue_take         in mem_kernel1_core_Learned(), around the seeding calls  src/bwamem.cpp:L1366-1397
//   meme_glue_take_chains  first line of that function's per-read loop: with MEME_B200_CHAIN=1 the read's filtered chains
//                          come from the GPU and the loop body (seeding, mem_chain_Learned, mem_chain_flt) is skipped  L1249
//   meme_glue_dump_chains  after that loop (test instrumentation, MEME_B200_DUMP_CHAINS=path)          L1410
//   meme_sam_prepass       in mem_process_seqs(), before kt_for(worker_sam) = meme_cigar_chunk + meme_matesw_chunk:
//   meme_cigar_chunk       with MEME_B200_SAM=1 the global alignment (score,
//                          CIGAR, NM, MD) of every region of the chunk is computed on the GPU   src/bwamem.cpp:L1962
//   meme_glue_take_cigar   in mem_reg2aln(), in front of the band-doubling loop around bwa_gen_cigar2: a region the pre-pass
//                          has seen takes its result from there                                  src/bwamem.cpp:L2339
//   meme_matesw_chunk      paired-end chunks: with MEME_B200_MATE=1 the mate-rescue
//                          Smith-Waterman pairs of the whole chunk (collected with the reference's own mem_sam_pe_batch_pre) are
//                          aligned on the GPU                                                   src/bwamem.cpp:L1962
//   meme_glue_take_matesw  in worker_sam(), in front of mem_sam_pe_batch: a block's kswr_t results come from the pre-pass
//                                                                                              src/bwamem.cpp:L1877
//   meme_glue_bsw          first line of BandedPairWiseSW::getScores8/16 and scalarBandedSWAWrapper
//                                                             src/bandedSWA.cpp:L242, L412, L1117, L1970, L2664, L3469, L4198
//                          (per-call seam; reached only by callers other than the chunk-level driver below)
//   meme_extend_chunk      in mem_process_seqs(), instead of kt_for(worker_aln)     src/bwamem.cpp:L1945
//                          (defined in meme_extend.inl, which the patched bwamem.cpp includes at its end)
//
// Everything is off unless the environment says MEME_B200=1 (seeding) / MEME_B200_CHAIN=1 (+ chaining) / MEME_B200_BSW=1 (extension), so the
// patched binary is byte-for-byte the reference's behaviour by default.
#pragma once
#include <stdint.h>
#include "bwamem.h"
#include "ksw.h"
#include "bandedSWA.h"
#include "LearnedIndex_seeding.h"

void meme_glue_set_prefix(const char *idx_prefix);
void meme_glue_seed_chunk(const mem_opt_t *opt, const bntseq_t *bns, bseq1_t *seqs, int n, int use_learned);
bool meme_glue_take(const bseq1_t *s, mem_tlv *smems, u64v *hits);
bool meme_glue_take_chains(bseq1_t *s, int l, mem_chain_v *chn);
void meme_glue_dump_chains(const bseq1_t *seq, int nseq, const mem_chain_v *chain_ar);
bool meme_glue_take_cigar(const char *query, const mem_alnreg_t *ar, uint32_t **cigar, int *n_cigar, int *NM);
bool meme_glue_take_matesw(int pair_begin, int pair_end, int64_t pcnt, kswr_t *aln);
bool meme_glue_bsw(SeqPair *pairs, uint8_t *seq_buf_ref, uint8_t *seq_buf_qer, int n_pairs, int w, int o_del, int e_del,
                   int o_ins, int e_ins, int zdrop, int end_bonus, int a, int b);

// used by meme_extend.inl
#include <vector>
#include "../../include/meme_b200.h"
bool meme_glue_cigar_enabled();
int64_t meme_glue_code_off(int read);
meme_cigar_task_t *meme_glue_cigar_tasks(int64_t n);      // the chunk's task array, page-locked, owned by the glue
void meme_glue_cigar_run(std::vector<int32_t> &reg_w, std::vector<int64_t> &task_off, const bseq1_t *seqs,
                         int n_reads, const mem_opt_t *opt, int64_t l_pac);
bool meme_glue_matesw_enabled();
bool meme_glue_matesw_have();
const std::vector<int64_t> &meme_glue_matesw_off();
void meme_glue_matesw_skip();
void meme_glue_matesw_run(int nseg, SeqPair **pairs, int64_t *n_pairs, uint8_t **ref, int64_t *n_ref, uint8_t **qer, int64_t *n_qer, const mem_opt_t *opt,
                          std::vector<int64_t> &pair_off, double collect_ms);
bool meme_glue_bsw_enabled();
bool meme_glue_bsw_text_mode();
void meme_glue_pinned_reserve(int threads, long reads_per_chunk);
void meme_glue_bsw_warmup(long reads);
void meme_glue_pinned_reserve_matesw(int threads, long reads_per_chunk);
void meme_glue_matesw_arena_sizes(int threads, long reads_per_chunk, size_t sizes[3]);
void *meme_glue_pinned_alloc(size_t bytes);
void meme_glue_pinned_free(void *p);
void meme_glue_bsw_segments(int nseg, SeqPair **pairs, int64_t *n_pairs, uint8_t **ref, int64_t *n_ref, uint8_t **qer, int64_t *n_qer,
                            int w, int tries, int prev_is_h0, int o_del, int e_del, int o_ins, int e_ins, int zdrop, int end_bonus,
                            int a, int b, int32_t **w_used);
// the same in two steps: start returns at once (the batch is driven by one host thread per device), wait returns when the results
// are in the pair arrays; one job at a time
void *meme_glue_bsw_segments_start(int nseg, SeqPair **pairs, int64_t *n_pairs, uint8_t **ref, int64_t *n_ref, uint8_t **qer, int64_t *n_qer,
                                   int w, int tries, int prev_is_h0, int o_del, int e_del, int o_ins, int e_ins, int zdrop, int end_bonus,
                                   int a, int b, int32_t **w_used);
void meme_glue_bsw_segments_wait(void *job);
