#!/usr/bin/env python
"""Write patched copies of three reference files (never touching the reference tree).

usage: apply_patch.py <reference root> <out dir>

The patch is the one shown in INTEGRATION.md: one include + four call sites of bwa-meme_b200/host/meme_glue.h.
It is applied by anchor text (not by line number) and fails loudly if an anchor is missing or ambiguous.
"""
import os
import re
import sys

HOOK_BSW = ("    if (meme_glue_bsw({arr}, seqBufRef, seqBufQer, numPairs, w, o_del, e_del, o_ins, e_ins, zdrop, end_bonus, "
            "w_match, -w_mismatch)) return;\n")


def patch_bwamem(src: str) -> str:
    out, n_seed, n_close, n_chunk, n_inc, n_aln, n_take, n_dump, in_k1 = [], 0, 0, 0, 0, 0, 0, 0, False
    n_sam, n_cig, n_cig_end, in_r2a = 0, 0, 0, False
    n_msw = n_msw_pre = 0
    for line in src.splitlines(keepends=True):
        if line.startswith("mem_aln_t mem_reg2aln("):
            in_r2a = True
        if in_r2a and line.strip() == "i = 0; a.cigar = 0;" and not n_cig:
            out.append("    if (meme_glue_take_cigar(query_, ar, &a.cigar, &a.n_cigar, &NM)) goto meme_cigar_done;   // meme_b200: GPU alignment of this region, if any\n")
            n_cig += 1
        if in_r2a and line.strip() == "} while (++i < 3 && score < ar->truesc - opt->a);" and not n_cig_end:
            out.append(line)
            out.append("meme_cigar_done:;   // meme_b200\n")
            n_cig_end += 1
            in_r2a = False
            continue
        if re.match(r"\s*kt_for\(worker_sam, &w,\s+n_\);", line):
            out.append("    meme_sam_prepass(w, n_);   // meme_b200: global alignments of all regions + mate-rescue Smith-Waterman of the whole chunk on the GPU\n")
            n_sam += 1
        if line.startswith("static void worker_sam("):
            out.append("bool meme_matesw_restore(worker_t *w, int pb, int pe, int tid, int64_t *pcnt);   // meme_b200 (meme_extend.inl)\n")
        if line.strip() == "int32_t gcnt = 0;" and not n_msw_pre:
            out.append(line)
            out.append("        if (!meme_matesw_restore(w, start >> 1, end >> 1, tid, &pcnt))   // meme_b200: the block's rescue pairs were collected by the chunk-level pre-pass\n")
            n_msw_pre += 1
            continue
        if line.strip() == "mem_sam_pe_batch(w->opt, &w->mmc, pcnt, pcnt8, aln, maxRefLen, maxQerLen, tid);" and not n_msw:
            out.append("        if (!meme_glue_take_matesw(start >> 1, end >> 1, pcnt, aln))   // meme_b200: the block's mate-rescue results from the GPU, if any\n")
            n_msw += 1
        if line.startswith("int mem_kernel1_core_Learned("):
            in_k1 = True
        if in_k1 and line.strip() == "smems->n = 0;" and not n_take:
            out.append("        if (meme_glue_take_chains(&seq_[l], l, &chain_ar[l])) continue;   // meme_b200: GPU chains for this read, if any\n")
            n_take += 1
        if in_k1 and "tprof[LEARNED_SEED_CHAIN][tid] += __rdtsc() - tim;" in line:
            out.append("    meme_glue_dump_chains(seq_, nseq, chain_ar);   // meme_b200: test instrumentation (MEME_B200_DUMP_CHAINS)\n")
            n_dump += 1
            in_k1 = False
        if line.startswith('#include "bwamem.h"') and not n_inc:
            out.append(line)
            out.append('#include "meme_glue.h"\n')
            n_inc += 1
            continue
        if "Learned_getSMEMsAllPosOneThread(&iaux, &raux, smems, hits, hasN, split_len,opt->split_width);" in line and not line.lstrip().startswith("//"):
            out.append("        if (!meme_glue_take(&seq_[l], smems, hits)) {   // meme_b200: GPU seeds for this read, if any\n")
            out.append(line)
            n_seed += 1
            continue
        if "ks_introsort(mem_smem_sort_lt_learned, smems->n, smems->a);" in line:
            out.append("        }   // meme_b200\n")
            out.append(line)
            n_close += 1
            continue
        if re.match(r"\s*kt_for\(worker_bwt, &w, n_\);", line):
            out.append("    meme_glue_seed_chunk(opt, w.fmi->idx->bns, seqs, n_, w.useLearned);   // meme_b200: whole-chunk GPU seeding (+ chaining)\n")
            out.append(line)
            n_chunk += 1
            continue
        if re.match(r"\s*kt_for\(worker_aln, &w, n_\);", line):
            out.append("    if (!meme_extend_chunk(w, n_))   // meme_b200: whole-chunk GPU seed extension\n")
            out.append(line)
            n_aln += 1
            continue
        if line.startswith("void mem_process_seqs(") and not n_aln:
            out.append("bool meme_extend_chunk(worker_t &w, int n);   // meme_b200 (meme_extend.inl)\n")
            out.append("void meme_sam_prepass(worker_t &w, int n);    // meme_b200 (meme_extend.inl)\n")
        out.append(line)
    assert (n_inc, n_seed, n_close, n_chunk, n_aln, n_take, n_dump) == (1, 1, 1, 1, 1, 1, 1), (n_inc, n_seed, n_close, n_chunk, n_aln, n_take, n_dump)
    assert (n_sam, n_cig, n_cig_end, n_msw, n_msw_pre) == (1, 1, 1, 1, 1), (n_sam, n_cig, n_cig_end, n_msw, n_msw_pre)
    return "".join(out) + '\n#include "meme_extend.inl"   // meme_b200: chunk-level seed extension\n'


def patch_banded(src: str) -> str:
    lines = src.splitlines(keepends=True)
    out, i, n = [], 0, 0
    while i < len(lines):
        line = lines[i]
        m = re.match(r"void BandedPairWiseSW::(getScores8|getScores16|scalarBandedSWAWrapper)\(SeqPair \*(\w+),", line)
        out.append(line)
        i += 1
        if m:
            arr = m.group(2)
            # copy the rest of the signature up to and including the line that opens the body
            while True:
                l2 = lines[i]
                out.append(l2)
                i += 1
                if "{" in l2:
                    break
            out.append(HOOK_BSW.format(arr=arr))
            n += 1
    assert n == 7, n
    return '#include "meme_glue.h"\n' + "".join(out)


def patch_fastmap(src: str) -> str:
    out, n, armed = [], 0, False
    for line in src.splitlines(keepends=True):
        out.append(line)
        if line.startswith("void memoryAllocLearned("):
            armed = True
        if armed and line.strip() == "{":
            out.append("    meme_glue_set_prefix(idx_prefix);   // meme_b200\n")
            armed = False
            n += 1
    assert n == 1, n
    return '#include "meme_glue.h"\n' + "".join(out)


def main():
    ref, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    for name, fn in (("bwamem.cpp", patch_bwamem), ("bandedSWA.cpp", patch_banded), ("fastmap.cpp", patch_fastmap)):
        src = open(os.path.join(ref, "src", name), newline="", encoding="latin-1").read()
        open(os.path.join(out, name), "w", newline="", encoding="latin-1").write(fn(src))
    print("patched bwamem.cpp, bandedSWA.cpp, fastmap.cpp ->", out)


if __name__ == "__main__":
    main()
