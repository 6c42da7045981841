// flat_host.cpp -- TEST INFRASTRUCTURE.  Single-lane CPU run of the flat seeding kernel (the CUDA source itself,
// compiled through cuda_shim.h) over an index given as the reference's file contents.  Mirrors what meme_index.cu
// (array build) and tpr_seed_run / k_gather2 (launch order, sort, hit lists) do on the device, in plain C++.
#include "cuda_shim.h"
#include <numeric>
#include <vector>
#include "../../bwa-meme_b200/csrc/meme_seed_flat.cuh"
#include "../../bwa-meme_b200/csrc/meme_seed_grp.cuh"

using namespace tpr;

namespace {
struct Result {
    std::vector<meme_smem_t> smems;
    std::vector<int64_t> smem_off, hit_off;
    std::vector<uint64_t> hits;
} g_res;

int g_impl = 0;       // 0: flat kernel, 1: regrouped kernel (meme_seed_grp.cuh)

template <int KIND, bool WIDE>
void run_kind(const IndexView &ix, Args &A, bool lng)
{
    if (g_impl == 0) k_flat<KIND, WIDE, 2, true, true>(ix, A);
    else if (lng) grp::k_grp<KIND, WIDE, true, true>(ix, A);
    else grp::k_grp<KIND, WIDE, false, true>(ix, A);
}

template <bool WIDE>
void run_kinds(const IndexView &ix, Args &A, const meme_seed_opt_t &opt, int64_t n, std::vector<Task2> &t2, u64 *ctrl, bool lng)
{
    A.kind = 1; A.n_tasks = n; ctrl[C_WORK] = 0;
    run_kind<1, WIDE>(ix, A, lng);
    const u64 n_t2 = ctrl[C_T2];
    if (opt.rounds >= 2 && n_t2) {
        A.kind = 2; A.n_tasks = (int64_t)n_t2; A.t2_in = t2.data(); A.t2_out = nullptr; A.t2_cap = 0; ctrl[C_WORK] = 0;
        run_kind<2, WIDE>(ix, A, lng);
    }
    if (opt.rounds >= 3 && opt.max_mem_intv > 0) {
        A.kind = 3; A.n_tasks = n; A.t2_in = nullptr; ctrl[C_WORK] = 0;
        run_kind<3, WIDE>(ix, A, lng);
    }
}
}  // namespace

extern "C" void flat_sim_set_impl(int impl) { g_impl = impl; }

extern "C" int flat_sim_seed(const uint8_t *text, uint64_t sa_num, const uint8_t *pos_packed, const void *l1_raw, uint64_t n_l1,
                             const void *l2_raw, uint64_t n_l2, int mode, int wide, const uint8_t *codes, const int64_t *read_off,
                             int64_t n_reads, const meme_seed_opt_t *opt, meme_seed_result_t *out, uint64_t counters[8])
{
    const u64 n = sa_num;
    // ---- index arrays (meme_index.cu: k_pack_text, k_unpack_pos, k_build_keys_isa, upload_rmi) ----
    std::vector<u64> pac((n + 31) / 32 + 2, 0);
    for (u64 i = 0; i < n; i++) pac[i >> 5] |= (u64)(text[i] & 3) << (2 * (i & 31));
    std::vector<u32> pos_lo(n), isa_lo(mode == 3 ? n : 0);
    std::vector<uint8_t> pos_hi(wide ? n + 8 : 0), isa_hi(wide && mode == 3 ? n + 8 : 0);
    std::vector<u64> keys(mode >= 2 ? n : 0);
    for (u64 i = 0; i < n; i++) {
        const uint8_t *p = pos_packed + i * 5;
        u64 hi = (u64)p[0] | ((u64)p[1] << 8) | ((u64)p[2] << 16) | ((u64)p[3] << 24);
        u64 pos = (hi << 8) | p[4];
        pos_lo[i] = (u32)pos;
        if (wide) pos_hi[i] = (uint8_t)(pos >> 32);
        if (mode >= 2) {
            u64 w = pos >> 5; int sh = (int)(pos & 31) * 2;
            u64 k = sh ? (pac[w] >> sh) | (pac[w + 1] << (64 - sh)) : pac[w];
            u64 left = n - pos;
            if (left < 32) k |= ~0ull << (2 * left);
            keys[i] = k;
        }
        if (mode == 3) { isa_lo[pos] = (u32)i; if (wide) isa_hi[pos] = (uint8_t)(i >> 32); }
    }
    auto rmi = [](const void *raw, u64 cnt) {
        std::vector<RmiRec> r(cnt ? cnt : 1);
        const uint8_t *b = (const uint8_t *)raw;
        for (u64 i = 0; i < cnt; i++) { memcpy(&r[i].alpha, b + i * 24, 8); memcpy(&r[i].beta, b + i * 24 + 8, 8); memcpy(&r[i].err, b + i * 24 + 16, 8); r[i].pad = 0; }
        return r;
    };
    std::vector<RmiRec> l1 = rmi(l1_raw, n_l1), l2 = rmi(l2_raw, n_l2);
    IndexView ix;
    memset(&ix, 0, sizeof ix);
    ix.sa_num = n; ix.mode = mode; ix.pos_wide = wide;
    ix.keys = mode >= 2 ? keys.data() : nullptr; ix.pos_lo = pos_lo.data(); ix.pos_hi = wide ? pos_hi.data() : nullptr;
    ix.isa_lo = mode == 3 ? isa_lo.data() : nullptr; ix.isa_hi = (wide && mode == 3) ? isa_hi.data() : nullptr;
    ix.pac = pac.data(); ix.l1 = l1.data(); ix.l2 = l2.data(); ix.n_l1 = n_l1; ix.n_l2 = n_l2;
    int bits = 0;
    while ((1ull << bits) < n_l2) bits++;
    ix.bit_shift = 64 - bits;
    // ---- reads (k_pack_reads) ----
    const int64_t nr = n_reads;
    std::vector<u64> packed((size_t)nr * TPR_RW + 16, 0);
    std::vector<u32> nmask((size_t)nr * TPR_RW + 16, 0);
    int max_len = 0;
    for (int64_t r = 0; r < nr; r++) {
        int64_t o = read_off[r]; int L = (int)(read_off[r + 1] - o);
        if (L < 0 || L > MEME_MAX_READ_LEN) return MEME_E_ARG;
        if (L > max_len) max_len = L;
        for (int p = 0; p < L; p++) {
            u32 c = codes[o + p];
            if (c > 3) nmask[r * TPR_RW + (p >> 5)] |= 1u << (p & 31);
            else packed[r * TPR_RW + (p >> 5)] |= (u64)c << (2 * (p & 31));
        }
    }
    // ---- pools and the three task kinds (tpr_seed_run) ----
    const size_t pool_cap = (size_t)nr * 64 + 65536, t2_cap = (size_t)nr * 16 + 65536;
    std::vector<u64> pool_keys(pool_cap, ~0ull);
    std::vector<PoolRec> pool_vals(pool_cap);
    std::vector<Task2> t2(t2_cap);
    std::vector<u32> rd_nsmem(nr + 1, 0);
    std::vector<u64> rd_nhit(nr + 1, 0), rd_maxref(nr + 1, 0);
    std::vector<uint8_t> rd_full(nr + 1, 0);
    u64 ctrl[64] = {0};
    Args A;
    memset(&A, 0, sizeof A);
    A.packed = packed.data(); A.nmask = nmask.data(); A.read_off = read_off;
    A.t2_in = nullptr; A.t2_out = t2.data(); A.t2_cap = t2_cap;
    A.pool_keys = pool_keys.data(); A.pool_vals = pool_vals.data(); A.pool_cap = pool_cap;
    A.rd_nsmem = rd_nsmem.data(); A.rd_nhit = rd_nhit.data(); A.rd_full = rd_full.data(); A.rd_maxref = rd_maxref.data();
    A.ctrl = ctrl; A.opt = *opt; A.grp_th = 24;
    const bool lng = max_len > 32 * (TPR_RW / 2);
    if (nr > 0) { if (wide) run_kinds<true>(ix, A, *opt, nr, t2, ctrl, lng); else run_kinds<false>(ix, A, *opt, nr, t2, ctrl, lng); }
    if (ctrl[C_POOL_OVF] || ctrl[C_T2_OVF]) return MEME_E_OVERFLOW;
    if (counters) { counters[0] = ctrl[C_SEARCH]; counters[1] = ctrl[C_BLOCK]; counters[2] = ctrl[C_KARY]; counters[3] = ctrl[C_EXT];
                    counters[4] = ctrl[C_PAC]; counters[5] = ctrl[C_ITER]; counters[6] = ctrl[C_T2]; counters[7] = ctrl[C_POOL]; }
    // ---- sort by (read, start, end), then SMEM records and hit lists (k_gather2) ----
    const u64 n_pool = std::min<u64>(ctrl[C_POOL], pool_cap);
    std::vector<u32> order;
    for (u64 k = 0; k < n_pool; k++) if (pool_keys[k] != ~0ull) order.push_back((u32)k);
    std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return pool_keys[a] < pool_keys[b]; });
    Result &R = g_res;
    R.smems.clear(); R.hits.clear(); R.smem_off.assign(nr + 1, 0); R.hit_off.assign(nr + 1, 0);
    auto sa_pos = [&](u64 i) { return (u64)pos_lo[i] | (wide ? (u64)pos_hi[i] << 32 : 0); };
    size_t k = 0;
    for (int64_t r = 0; r < nr; r++) {
        R.smem_off[r] = (int64_t)R.smems.size(); R.hit_off[r] = (int64_t)R.hits.size();
        int64_t hb = 0;
        for (; k < order.size() && (int64_t)(pool_keys[order[k]] >> 32) == r; k++) {
            u64 key = pool_keys[order[k]]; const PoolRec &a = pool_vals[order[k]];
            u64 c = a.cnt, step = 1, take = c;
            if (opt->max_occ > 0) {
                step = c > (u64)opt->max_occ ? c / opt->max_occ : 1;
                take = (c + step - 1) / step;
                if (take > (u64)opt->max_occ) take = (u64)opt->max_occ;
            }
            for (u64 j = 0; j < take; j++) R.hits.push_back(sa_pos(a.lo + j * step));
            meme_smem_t m;
            m.start = (int32_t)((key >> 16) & 0xffff); m.end = (int32_t)(key & 0xffff);
            m.hitbeg = (int32_t)hb; m.hitcount = (int32_t)a.cnt; m.cache_refpos = sa_pos(a.lo);
            R.smems.push_back(m);
            hb += (int64_t)take;
        }
    }
    R.smem_off[nr] = (int64_t)R.smems.size(); R.hit_off[nr] = (int64_t)R.hits.size();
    out->n_reads = nr; out->n_smems = (int64_t)R.smems.size(); out->n_hits = (int64_t)R.hits.size();
    out->smems = R.smems.data(); out->smem_off = R.smem_off.data(); out->hit_off = R.hit_off.data(); out->hits = R.hits.data();
    return MEME_OK;
}
