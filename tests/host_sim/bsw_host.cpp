// bsw_host.cpp -- TEST INFRASTRUCTURE.  The inter-sequence banded-SW pair function of the product
// (bwa-meme_b200/csrc/meme_bsw_is.cuh, __host__ __device__) compiled with g++ and run pair by pair on the CPU, so that
// its recurrence, band handling and retry rule can be checked against the oracle without a GPU.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../bwa-meme_b200/csrc/meme_bsw_is.cuh"

extern "C" int bsw_is_sim(meme_seqpair_t *pairs, int64_t n, const uint8_t *ref, const uint8_t *qer, int w, int tries, int prev_is_h0,
                          const meme_bsw_opt_t *opt, int32_t *w_used, uint64_t *cells_out)
{
    BswIsConst P = {opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, opt->zdrop, opt->end_bonus, opt->a, opt->b};
    std::vector<u32> row;
    std::vector<uint8_t> qrow;
    u64 cells = 0;
    for (int64_t i = 0; i < n; i++) {
        meme_seqpair_t &p = pairs[i];
        if (p.len2 > BSW_IS_MAX_QLEN || p.h0 + p.len2 * (opt->a > 0 ? opt->a : 0) >= BSW_IS_MAX_SCORE) return -1;
        row.assign((size_t)p.len2 + 2, 0xdeadbeefu);
        qrow.assign((size_t)p.len2 + 2, 0xee);
        BswIsOut o;
        BswTarget T = {ref + p.idr, nullptr, 0, 1};
        int wu = bsw_is_retry(P, w, tries, prev_is_h0 ? p.h0 : -1, qer + p.idq, p.len2, T, p.len1, p.h0, row.data(), qrow.data(), 1, o, cells);
        p.score = o.score; p.tle = o.tle; p.gtle = o.gtle; p.qle = o.qle; p.gscore = o.gscore; p.max_off = o.max_off;
        if (w_used) w_used[i] = wu;
    }
    if (cells_out) *cells_out = cells;
    return 0;
}

// The same with the reference side given as positions in a 2-bit text (the "reference by position" form of the product):
// text = codes 0..3; pos[i] = first text position of pair i, dir[i] = +1 / -1.
extern "C" int bsw_is_sim_text(meme_seqpair_t *pairs, int64_t n, const uint8_t *text, int64_t n_text, const int64_t *pos, const int32_t *dir,
                               const uint8_t *qer, int w, int tries, int prev_is_h0, const meme_bsw_opt_t *opt, int32_t *w_used)
{
    BswIsConst P = {opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, opt->zdrop, opt->end_bonus, opt->a, opt->b};
    std::vector<u64> pac((size_t)(n_text + 31) / 32 + 2, 0);
    for (int64_t i = 0; i < n_text; i++) pac[(size_t)(i >> 5)] |= (u64)(text[i] & 3) << (2 * (i & 31));
    std::vector<u32> row;
    std::vector<uint8_t> qrow;
    u64 cells = 0;
    for (int64_t i = 0; i < n; i++) {
        meme_seqpair_t &p = pairs[i];
        row.assign((size_t)p.len2 + 2, 0xdeadbeefu);
        qrow.assign((size_t)p.len2 + 2, 0xee);
        BswIsOut o;
        BswTarget T = {nullptr, pac.data(), (u64)pos[i], dir[i]};
        int wu = bsw_is_retry(P, w, tries, prev_is_h0 ? p.h0 : -1, qer + p.idq, p.len2, T, p.len1, p.h0, row.data(), qrow.data(), 1, o, cells);
        p.score = o.score; p.tle = o.tle; p.gtle = o.gtle; p.qle = o.qle; p.gscore = o.gscore; p.max_off = o.max_off;
        if (w_used) w_used[i] = wu;
    }
    return 0;
}
