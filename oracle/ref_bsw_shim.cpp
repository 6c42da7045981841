// oracle/ref_bsw_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin driver around the UNMODIFIED reference class BandedPairWiseSW
// (/root/reference/src/bandedSWA.h:L118-135, bandedSWA.cpp:L116-259) so that the
// oracle restatement (oracle/bsw_oracle.c) and the CUDA kernel can be pinned
// against outputs of the reference itself.
//
// usage: bsw_shim <in.bin> <out.bin> <scalar|vec16|vec8>
//        bsw_shim <pairs.dump> - bench <threads> <reps>      timing of the reference's getScores8/16 on a pair dump
//                                                            (MEME_B200_DUMP_PAIRS format, bwa-meme_b200/host/meme_glue.cpp)
//
// in.bin  : int32 header[12] = {npairs, w, o_del, e_del, o_ins, e_ins, zdrop, end_bonus, a, b, nref, nqer}
//           then npairs * 14 int32 (the reference's SeqPair layout, bandedSWA.h:L90-99)
//           then nref bytes (reference codes 0..4), nqer bytes (query codes 0..4)
// out.bin : npairs * 6 int32 = {score, tle, gtle, qle, gscore, max_off}
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <string>
#include <algorithm>
#include <chrono>
#include <omp.h>
#include "bandedSWA.h"
#include "bwa.h"

// the reference's translation units expect these globals (main.cpp:L42-46)
uint64_t proc_freq, tprof[LIM_R][LIM_C], prof[LIM_R];

// The extension driver's use of the vector kernels, timed: pairs classed 8-bit / 16-bit as in src/bwamem.cpp:L2804-2812,
// each class sorted by length (sortPairsLen), blocks of pairs handed to `threads` host threads (one BandedPairWiseSW
// each, as every kt_for worker has), first with w, then the pairs that are not final again with 2w (L2968-3161).
static int bench_dump(const char *path, int threads, int reps)
{
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); return 1; }
    int64_t hdr[3]; int32_t par[12];
    if (fread(hdr, 8, 3, f) != 3 || fread(par, 4, 12, f) != 12) return 1;
    const int64_t np = hdr[0];
    std::vector<SeqPair> all((size_t)np);
    if ((int64_t)fread(all.data(), sizeof(SeqPair), (size_t)np, f) != np) return 1;
    std::vector<uint8_t> ref((size_t)hdr[1] + (1 << 16)), qer((size_t)hdr[2] + (1 << 16));
    if ((int64_t)fread(ref.data(), 1, (size_t)hdr[1], f) != hdr[1] || (int64_t)fread(qer.data(), 1, (size_t)hdr[2], f) != hdr[2]) return 1;
    fclose(f);
    const int w = par[0], tries = par[1], prev_is_h0 = par[2], a = par[9], b = par[10];
    int8_t mat[25];
    bwa_fill_scmat(a, b, mat);
    std::vector<SeqPair> c8, c16;
    int64_t n_scalar = 0;
    for (auto &p : all) {
        int minval = p.h0 + std::min(p.len1, p.len2) * a;
        if (p.len1 < MAX_SEQ_LEN8 && p.len2 < MAX_SEQ_LEN8 && minval < MAX_SEQ_LEN8) c8.push_back(p);
        else if (p.len1 < MAX_SEQ_LEN16 && p.len2 < MAX_SEQ_LEN16 && minval < MAX_SEQ_LEN16) c16.push_back(p);
        else n_scalar++;
    }
    auto by_len = [](const SeqPair &x, const SeqPair &y) { return x.len1 < y.len1; };
    std::stable_sort(c8.begin(), c8.end(), by_len);
    std::stable_sort(c16.begin(), c16.end(), by_len);
    const int BLK = 4096;
    std::vector<BandedPairWiseSW *> bsw((size_t)threads);
    for (int t = 0; t < threads; t++)
        bsw[(size_t)t] = new BandedPairWiseSW(par[3], par[4], par[5], par[6], par[7], par[8], mat, (int8_t)a, (int8_t)b, 1);
    double best = 1e30, sum = 0;
    int64_t n_retry = 0;
    for (int r = 0; r < reps + 1; r++) {
        std::vector<SeqPair> w8 = c8, w16 = c16, r8, r16;
        w8.resize(c8.size() + 128); w16.resize(c16.size() + 128);
        auto t0 = std::chrono::steady_clock::now();
        for (int t = 0; t < tries; t++) {
            const int wt = w << t;
            const int64_t n8 = (int64_t)w8.size() - 128, n16 = (int64_t)w16.size() - 128;
            const int64_t b8 = (n8 + BLK - 1) / BLK, b16 = (n16 + BLK - 1) / BLK;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
            for (int64_t k = 0; k < b8 + b16; k++) {
                BandedPairWiseSW *B = bsw[(size_t)omp_get_thread_num()];
                // every block gets its own padded copy: the wrappers write SIMD-width padding entries behind the block
                std::vector<SeqPair> blk((size_t)BLK + 128);
                if (k < b16) {
                    int64_t lo = k * BLK, n = std::min<int64_t>(BLK, n16 - lo);
                    std::copy(w16.begin() + lo, w16.begin() + lo + n, blk.begin());
                    B->getScores16(blk.data(), ref.data(), qer.data(), (int)n, 1, wt);
                    std::copy(blk.begin(), blk.begin() + n, w16.begin() + lo);
                } else {
                    int64_t lo = (k - b16) * BLK, n = std::min<int64_t>(BLK, n8 - lo);
                    std::copy(w8.begin() + lo, w8.begin() + lo + n, blk.begin());
                    B->getScores8(blk.data(), ref.data(), qer.data(), (int)n, 1, wt);
                    std::copy(blk.begin(), blk.begin() + n, w8.begin() + lo);
                }
            }
            if (t + 1 == tries) break;
            // the pairs whose result is not final go again with twice the band
            r8.clear(); r16.clear();
            for (int64_t i = 0; i < n8; i++) { const SeqPair &p = w8[(size_t)i]; if (!(p.score == (prev_is_h0 ? p.h0 : -1) || p.max_off < (wt >> 1) + (wt >> 2))) r8.push_back(p); }
            for (int64_t i = 0; i < n16; i++) { const SeqPair &p = w16[(size_t)i]; if (!(p.score == (prev_is_h0 ? p.h0 : -1) || p.max_off < (wt >> 1) + (wt >> 2))) r16.push_back(p); }
            n_retry = (int64_t)(r8.size() + r16.size());
            r8.resize(r8.size() + 128); r16.resize(r16.size() + 128);
            w8.swap(r8); w16.swap(r16);
        }
        double ms = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e3;
        if (r > 0) { best = std::min(best, ms); sum += ms; }          // the first pass warms up
    }
    printf("{\"pairs\": %lld, \"pairs_8bit\": %zu, \"pairs_16bit\": %zu, \"pairs_scalar_skipped\": %lld, \"retried\": %lld, \"threads\": %d, "
           "\"reps\": %d, \"ms_mean\": %.3f, \"ms_best\": %.3f, \"pairs_per_s\": %.1f}\n", (long long)np, c8.size(), c16.size(),
           (long long)n_scalar, (long long)n_retry, threads, reps, sum / reps, best, (double)(c8.size() + c16.size()) / (sum / reps / 1e3));
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 6 && std::string(argv[3]) == "bench") return bench_dump(argv[1], atoi(argv[4]), atoi(argv[5]));
    if (argc < 4) { fprintf(stderr, "usage: %s in.bin out.bin scalar|vec16|vec8\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    int32_t hdr[12];
    if (fread(hdr, 4, 12, f) != 12) return 1;
    int np = hdr[0], w = hdr[1];
    std::vector<SeqPair> pairs(np + 64);
    static_assert(sizeof(SeqPair) == 56, "SeqPair layout");
    if ((int)fread(pairs.data(), sizeof(SeqPair), np, f) != np) return 1;
    // generous tail padding: the vector kernels read SIMD-width blocks past the end
    std::vector<uint8_t> ref((size_t)hdr[10] + (1 << 16)), qer((size_t)hdr[11] + (1 << 16));
    if ((int)fread(ref.data(), 1, hdr[10], f) != hdr[10]) return 1;
    if ((int)fread(qer.data(), 1, hdr[11], f) != hdr[11]) return 1;
    fclose(f);

    int8_t mat[25];
    bwa_fill_scmat(hdr[8], hdr[9], mat);
    BandedPairWiseSW bsw(hdr[2], hdr[3], hdr[4], hdr[5], hdr[6], hdr[7], mat, (int8_t)hdr[8], (int8_t)hdr[9], 1);
    std::string mode = argv[3];
    if (mode == "scalar") {
        bsw.scalarBandedSWAWrapper(pairs.data(), ref.data(), qer.data(), np, 1, w);
    }
#if __AVX2__ || __AVX512BW__
    else if (mode == "vec16") {
        bsw.getScores16(pairs.data(), ref.data(), qer.data(), np, 1, w);
    } else if (mode == "vec8") {
        bsw.getScores8(pairs.data(), ref.data(), qer.data(), np, 1, w);
    }
#endif
    else { fprintf(stderr, "unknown mode %s\n", argv[3]); return 2; }

    f = fopen(argv[2], "wb");
    if (!f) { perror(argv[2]); return 1; }
    for (int i = 0; i < np; i++) {
        int32_t o[6] = {pairs[i].score, pairs[i].tle, pairs[i].gtle, pairs[i].qle, pairs[i].gscore, pairs[i].max_off};
        fwrite(o, 4, 6, f);
    }
    fclose(f);
    return 0;
}
