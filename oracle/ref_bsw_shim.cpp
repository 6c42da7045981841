// oracle/ref_bsw_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin driver around the UNMODIFIED reference class BandedPairWiseSW
// (/root/reference/src/bandedSWA.h:L118-135, bandedSWA.cpp:L116-259) so that the
// oracle restatement (oracle/bsw_oracle.c) and the CUDA kernel can be pinned
// against outputs of the reference itself.
//
// usage: bsw_shim <in.bin> <out.bin> <scalar|vec16|vec8>
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
#include "bandedSWA.h"
#include "bwa.h"

// the reference's translation units expect these globals (main.cpp:L42-46)
uint64_t proc_freq, tprof[LIM_R][LIM_C], prof[LIM_R];

int main(int argc, char **argv)
{
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
