// matesw_host.cpp -- TEST INFRASTRUCTURE.  The per-pair mate-rescue Smith-Waterman of the product
// (bwa-meme_b200/csrc/meme_matesw_core.cuh, __host__ __device__) compiled with g++ and run pair by pair on the CPU with the scratch
// discipline of the kernel (arrays laid out as lane `lane` of `stride` interleaved pairs), so that it can be checked against the
// oracle and the reference's outputs without a GPU.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../bwa-meme_b200/csrc/meme_matesw_core.cuh"

struct Rec { int64_t r_off, q_off; int32_t len1, len2, xtra, pad; };

extern "C" int matesw_sim(const Rec *recs, int64_t n, const uint8_t *ref, const uint8_t *qer, const msw::Opt *opt, msw::Out *out, uint64_t *cells_out,
                          int stride, int lane)
{
    uint64_t cells = 0;
    for (int64_t k = 0; k < n; k++) {
        const Rec &r = recs[k];
        const bool byte = (r.xtra & msw::XBYTE) != 0;
        if (r.len2 > msw::MAX_COLS - 16 || r.len1 > msw::MAX_ROWS || (byte && r.len2 > 255)) return -1;
        const int cols = msw::quanta(r.len2, byte);
        std::vector<uint32_t> row(((size_t)cols + 1) * (size_t)stride, 0xdeadbeefu);
        std::vector<uint8_t> qrow(((size_t)cols + 1) * (size_t)stride, 0xee);
        std::vector<int16_t> rm(((size_t)r.len1 + 1) * (size_t)stride, (int16_t)0x7777);
        msw::Scratch S;
        S.row = row.data() + lane; S.qrow = qrow.data() + lane; S.stride = stride; S.rm = rm.data() + lane; S.rstride = stride;
        msw::pair(*opt, ref + r.r_off, r.len1, qer + r.q_off, r.len2, r.xtra, S, out[k], cells);
        for (size_t x = 0; x < row.size(); x++) if ((int)(x % (size_t)stride) != lane && row[x] != 0xdeadbeefu) return -2;      // other lanes untouched
        for (size_t x = 0; x < rm.size(); x++) if ((int)(x % (size_t)stride) != lane && rm[x] != (int16_t)0x7777) return -2;
    }
    if (cells_out) *cells_out = cells;
    return 0;
}
