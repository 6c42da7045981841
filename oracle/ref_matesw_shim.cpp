// oracle/ref_matesw_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Calls the UNMODIFIED reference's mate-rescue batch step -- sort_classify (src/bwamem.cpp:L1797-1826) followed by
// mem_sam_pe_batch (src/bwamem_pair.cpp:L719-817: kswv::getScores8/16 phase 0, selection, reversal, phase 1; AVX512 builds
// only) -- on a file of SW pairs and writes the kswr_t of every pair, so that oracle/matesw_oracle.c and the CUDA kernel
// can be pinned against outputs of the reference itself.
//
// usage: matesw_shim <in.bin> <out.bin>
//        matesw_shim <in.bin> - bench <threads> <reps>    timing of the same step, blocks of pairs per host thread
//
// in.bin : int64 n, n_ref, n_qer; int32 opt[6] = {o_del, e_del, o_ins, e_ins, a, b}; int32 pad[2];
//          n records {int64 r_off, q_off; int32 len1, len2, xtra, pad}; n_ref reference bytes; n_qer query bytes (codes 0..4)
// out.bin: n * 7 int32 = {score, te, qe, score2, te2, tb, qb}
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <chrono>
#include <omp.h>
#include "bwamem.h"
#include "kswv.h"
#include "macro.h"

uint64_t proc_freq, tprof[LIM_R][LIM_C], prof[LIM_R];

int64_t sort_classify(mem_cache *mmc, int64_t pcnt, int tid);      // src/bwamem.cpp:L1797 (no header declares it)

struct Rec { int64_t r_off, q_off; int32_t len1, len2, xtra, pad; };

// one block of pairs through the reference, as worker_sam runs it for a block of reads (src/bwamem.cpp:L1870-1877)
static void run_block(const mem_opt_t *opt, const Rec *recs, int64_t n, const uint8_t *ref, const uint8_t *qer, kswr_t *out, int tid, mem_cache *mmc,
                      std::vector<uint8_t> &bref, std::vector<uint8_t> &bqer)
{
    SeqPair *arr = mmc->seqPairArrayLeft128[tid];
    int64_t ro = 0, qo = 0;
    int32_t maxRefLen = 0, maxQerLen = 0;
    for (int64_t k = 0; k < n; k++) {
        const Rec &r = recs[k];
        SeqPair sp;
        memset(&sp, 0xff, sizeof sp);
        sp.idr = (int32_t)ro; sp.idq = (int32_t)qo; sp.len1 = r.len1; sp.len2 = r.len2; sp.h0 = r.xtra; sp.regid = (int32_t)k;
        memcpy(bref.data() + ro, ref + r.r_off, (size_t)r.len1);       // mem_sam_pe_batch reverses sequences in place
        memcpy(bqer.data() + qo, qer + r.q_off, (size_t)r.len2);
        ro += r.len1; qo += r.len2;
        maxRefLen = std::max(maxRefLen, r.len1); maxQerLen = std::max(maxQerLen, r.len2);
        arr[k] = sp;
    }
    mmc->seqBufLeftRef[tid * CACHE_LINE] = bref.data();
    mmc->seqBufLeftQer[tid * CACHE_LINE] = bqer.data();
    int64_t pcnt = n;
    int64_t pcnt8 = sort_classify(mmc, pcnt, tid);
    kswr_t *aln = (kswr_t *)_mm_malloc((size_t)(pcnt + SIMD_WIDTH8) * sizeof(kswr_t), 64);
    memset(aln, 0x5a, (size_t)(pcnt + SIMD_WIDTH8) * sizeof(kswr_t));
    mem_sam_pe_batch(opt, mmc, pcnt, pcnt8, aln, maxRefLen, maxQerLen, tid);
    memcpy(out, aln, (size_t)n * sizeof(kswr_t));
    _mm_free(aln);
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <in.bin> <out.bin> | <in.bin> - bench <threads> <reps>\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 3; }
    int64_t hdr[3]; int32_t par[8];
    if (fread(hdr, 8, 3, f) != 3 || fread(par, 4, 8, f) != 8) return 4;
    const int64_t n = hdr[0];
    std::vector<Rec> recs((size_t)n);
    if (n && (int64_t)fread(recs.data(), sizeof(Rec), (size_t)n, f) != n) return 4;
    std::vector<uint8_t> ref((size_t)hdr[1] + 64), qer((size_t)hdr[2] + 64);
    if ((int64_t)fread(ref.data(), 1, (size_t)hdr[1], f) != hdr[1] || (int64_t)fread(qer.data(), 1, (size_t)hdr[2], f) != hdr[2]) return 4;
    fclose(f);
    mem_opt_t *opt = mem_opt_init();
    opt->o_del = par[0]; opt->e_del = par[1]; opt->o_ins = par[2]; opt->e_ins = par[3]; opt->a = par[4]; opt->b = par[5];
    bwa_fill_scmat(opt->a, opt->b, opt->mat);
    const bool bench = argc >= 6 && !strcmp(argv[3], "bench");
    const int threads = bench ? atoi(argv[4]) : 1, reps = bench ? atoi(argv[5]) : 1;
    int64_t BLK = bench ? 512 : (n > 0 ? n : 1);                 // bench: about what a 512-read block of worker_sam collects
    if (getenv("MATESW_BLK")) BLK = std::max(1L, atol(getenv("MATESW_BLK")));
    static mem_cache mmc;
    std::vector<std::vector<uint8_t>> bref((size_t)threads), bqer((size_t)threads);
    int64_t mr = 0, mq = 0;
    for (int64_t b0 = 0; b0 < n; b0 += BLK) {
        int64_t r = 0, q = 0;
        for (int64_t k = b0; k < std::min(n, b0 + BLK); k++) { r += recs[(size_t)k].len1; q += recs[(size_t)k].len2; }
        mr = std::max(mr, r); mq = std::max(mq, q);
    }
    for (int t = 0; t < threads; t++) {
        mmc.seqPairArrayLeft128[t] = (SeqPair *)malloc((size_t)(BLK + MAX_LINE_LEN + 2 * SIMD_WIDTH8) * sizeof(SeqPair));
        mmc.seqPairArrayRight128[t] = (SeqPair *)malloc((size_t)(BLK + MAX_LINE_LEN + 2 * SIMD_WIDTH8) * sizeof(SeqPair));
        bref[(size_t)t].resize((size_t)mr + 4096); bqer[(size_t)t].resize((size_t)mq + 4096);
    }
    std::vector<kswr_t> out((size_t)n + 1);
    double best = 1e30, sum = 0;
    for (int rep = 0; rep < reps; rep++) {
        auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
        for (int64_t b0 = 0; b0 < n; b0 += BLK) {
            const int tid = omp_get_thread_num();
            run_block(opt, recs.data() + b0, std::min(BLK, n - b0), ref.data(), qer.data(), out.data() + b0, tid, &mmc, bref[(size_t)tid], bqer[(size_t)tid]);
        }
        const double ms = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e3;
        best = std::min(best, ms); sum += ms;
    }
    if (bench) {
        double cells = 0;
        for (auto &r : recs) cells += (double)r.len1 * r.len2;
        printf("{\"pairs\": %ld, \"threads\": %d, \"reps\": %d, \"ms_mean\": %.3f, \"ms_best\": %.3f, \"cells\": %.0f, \"gcups\": %.3f}\n", (long)n, threads, reps,
               sum / reps, best, cells, cells / (sum / reps) / 1e6);
        return 0;
    }
    FILE *g = fopen(argv[2], "wb");
    if (!g) { perror(argv[2]); return 5; }
    fwrite(out.data(), sizeof(kswr_t), (size_t)n, g);
    fclose(g);
    return 0;
}
