// meme_seed_grp.cuh -- kernel A, regrouped: the seeding state machine with CTA-wide regrouping of tasks by state.
//
// Same tasks, same search engine and same results as meme_seed_flat.cuh (T1 zig-zag pass of a read, T2 one reseed,
// T3 third round of a read; reference src/LearnedIndex_seeding.cpp:L913-972, L1691-2126, L974-1466 and the searches
// L2131-4189), but WHO executes WHAT is different.  In the flat kernel a lane owns a task and walks its own path
// through the state machine: 4-8 of 32 lanes are active per instruction (ncu, profiles/r1).  Here
//   * the state of every task lives in SHARED memory: one CTA per SM keeps K x 32 task slots (K = 12..16) as
//     struct-of-arrays cells [field][slot]; slot (k, l) belongs to "lane class" l, i.e. it is only ever touched by
//     lane l of some warp, so every field access of a warp is bank-conflict-free by construction;
//   * every slot sits in exactly one POOL (a byte per slot): the control-entry states of the state machine
//     (P_S_RMI1, P_S_LOC, P_E, ...), "needs a block probe" (Q_BLOCK), "needs packed-reference words" (Q_PAC),
//     "empty" (Q_EMPTY), or busy.  A warp repeatedly picks the fullest pool, claims one slot of that pool per lane
//     (a per-pool token makes the claim exclusive: one CAS per warp, no per-lane atomics) and runs ONE step for all
//     claimed tasks with the warp converged on the same entry state;
//   * a step ends by moving each task to its next pool.  Memory requests are steps of their own: the Q_BLOCK /
//     Q_PAC passes issue the cp.async copies of 32 tasks at once (one 64-byte-aligned block of 8 keys + 32 bytes of
//     positions = six 16-byte copies when the block is aligned, 8/4-byte copies otherwise), the batch goes into the
//     warp's FIFO of pending batches, the warp does other passes meanwhile, and retires the batch later
//     (cp.async.wait_group N) with a fully converged compare loop.
// The control code of a step is the flat kernel's state machine verbatim (goto-based); its variables are references
// to the task's cells, so the compiler loads/stores exactly what a path touches.
// tests/host_sim compiles this file with g++ (one lane, one warp) to check the logic without a GPU.
#pragma once
#include "meme_seed_task.cuh"

#ifndef MEME_HOST_SIM
#define GRP_SMEM_DECL extern __shared__ __align__(16) unsigned char grp_smem[]
#define GRP_CP4(dst, src) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((u32)__cvta_generic_to_shared(dst)), "l"(src) : "memory")
#define GRP_CP8(dst, src) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((u32)__cvta_generic_to_shared(dst)), "l"(src) : "memory")
#define GRP_CP16(dst, src) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((u32)__cvta_generic_to_shared(dst)), "l"(src) : "memory")
#define GRP_CP_COMMIT() asm volatile("cp.async.commit_group;" ::: "memory")
#define GRP_CP_WAIT(n) asm volatile("cp.async.wait_group %0;" ::"n"(n) : "memory")
#define GRP_FENCE() __threadfence_block()
#define GRP_SLEEP(ns) __nanosleep(ns)
#define GRP_LANE() (threadIdx.x & 31)
#define GRP_WARP() (threadIdx.x >> 5)
#define GRP_BALLOT(p) __ballot_sync(0xffffffffu, (p))
#define GRP_SHFL(v, l) __shfl_sync(0xffffffffu, (v), (l))
#define GRP_REDUCE_MIN(v) __reduce_min_sync(0xffffffffu, (unsigned)(v))
#define GRP_SYNCWARP() __syncwarp()
#define GRP_SYNCTHREADS() __syncthreads()
#endif

#ifndef GRP_TRACE
#define GRP_TRACE(...) ((void)0)
#endif

namespace grp {
using namespace tpr;

// A slot's state byte: a control label (< 0x40, the block its task resumes at: pool PL_CTL), or one of the codes below.
enum { Q_BLOCK = 0x40, Q_PAC = 0x41, Q_EMPTY = 0x42, SL_BUSY = 0xfe, SL_DEAD = 0xff };
// the four pools a warp can claim from
enum { PL_CTL = 0, PL_BLOCK = 1, PL_PAC = 2, PL_EMPTY = 3 };
enum { R_INIT_L = 0, R_INIT_R, R_LV_L, R_LV_R };
enum { RMI_L2 = 0, RMI_L1 = 1 };
enum { B_BLOCK = 0, B_PAC = 1, B_SIMPLE = 2 };
#define GRP_NONE 0xffffu
#if defined(GRP_DEPTH) && GRP_DEPTH > 4
#error "GRP_DEPTH > 4: extend the cp.async.wait_group cases"
#endif
#ifndef GRP_DEPTH
#define GRP_DEPTH 3                       // pending batches per warp
#endif

template <bool WIDE> struct IdxT { typedef u32 type; typedef int stype; };
template <> struct IdxT<true> { typedef u64 type; typedef long long stype; };

// idx_t cells
enum { I_s_hint, I_center, I_lo_b, I_hi_b, I_k_stride, I_k_pa, I_base, I_lo, I_hi, I_e_from, I_e_cur, I_e_edge, I_max_refpos,
       I_cache_refpos, I_r3_last_lo, I_m_base, I_rad_lo, I_rad_hi, I_lv_limit, I_e_s, I_e_limit, I_r3_last_cnt, I_m_stride, N_IDX };
// u32 cells
enum { U_rid, U_min_intv, U_s_need, U_q0lo, U_q0hi, U_pphi_lo, U_pphi_hi, N_U32 };
// 16-bit cells
enum { S_state, S_mreq, S_L_, S_pivot, S_sp, S_next, S_cache_pivot, S_cache_pivot_end, S_q_strand, S_q_from, S_q_V, S_s_pivot, S_it,
       S_nbl, S_nbr, S_lev, S_e_L, S_e_nb, S_e_guard, S_as_pv, S_m_cap, S_min_seed_len, S_s_msl, S_d_ret, S_nvalid, S_lv_ret,
       S_e_dir, S_e_cnt, S_e_ret, S_e_up, S_e_sh, S_as_ret, S_as_right, S_as_emit, S_s_emit, S_s_stepwise, S_s_isa, S_k_contig,
       S_s_mode, S_full_match, S_use_cached, S_m_count, S_m_ldk, S_r_less, S_undet, S_r_count, S_rl_off, S_bl_off, S_pac_e,
       S_pac_k0, S_anyn, S_m_fast, N_S16 };

template <bool WIDE, bool LONG>
struct Lay {
    typedef typename IdxT<WIDE>::type idx_t;
    static constexpr int RWT = LONG ? TPR_RW : TPR_RW / 2;      // 64-bit words of a packed read kept per task
    static constexpr int K = WIDE ? (LONG ? 12 : 13) : (LONG ? 15 : 16);
    static constexpr int NT = K * 32;
    static constexpr int NQ = WIDE ? 8 : 6;                     // 16-byte staging cells: 4 of keys, 2 of positions (+2 of high bytes)
    static constexpr int N_RD = 3 * RWT;                        // u32 cells: 2*RWT read halves + RWT N-mask words
    static constexpr size_t OFF_STG = 0;
    static constexpr size_t OFF_IDX = OFF_STG + (size_t)NQ * NT * 16;
    static constexpr size_t OFF_U32 = OFF_IDX + (size_t)N_IDX * NT * sizeof(idx_t);
    static constexpr size_t OFF_RD = OFF_U32 + (size_t)N_U32 * NT * 4;
    static constexpr size_t OFF_S16 = OFF_RD + (size_t)N_RD * NT * 4;
    static constexpr size_t OFF_LCP = OFF_S16 + (size_t)N_S16 * NT * 2;
    static constexpr size_t OFF_SST = OFF_LCP + (size_t)16 * NT * 2;    // state bytes [32 lane classes][16 slots]
    static constexpr size_t OFF_TOK = OFF_SST + 512;                    // u32 [0..3] pool tokens, [4..7] pool task counts, [8] live slots
    static constexpr size_t BYTES = OFF_TOK + 128;
};

struct OutState { u64 pslot, pend, tslot, tend; };
struct Stats { u32 search, block, kary, ext, pac, iter, pass_ctl, pass_mem, lanes_ctl, lanes_mem; };

template <int KIND>
__device__ __noinline__ void emit_record(const Args &A, OutState &o, u32 rid, int start, int end, u64 lo_, u64 cnt,
                                         bool full_match, u64 max_refpos)
{
    if (o.pslot == o.pend) { o.pslot = atomicAdd((unsigned long long *)&A.ctrl[C_POOL], 16ull); o.pend = o.pslot + 16; }
    u64 k = o.pslot++;
    u32 c32 = cnt > 0xffffffffull ? 0xffffffffu : (u32)cnt;
    if (k < A.pool_cap) {
        A.pool_keys[k] = ((u64)rid << 32) | ((u64)start << 16) | (u64)end;
        PoolRec v; v.lo = lo_; v.cnt = c32; v.pad = 0;
        A.pool_vals[k] = v;
    } else atomicExch((unsigned long long *)&A.ctrl[C_POOL_OVF], 1ull);
    u32 c = c32;
    if (A.opt.max_occ > 0) {
        u32 mo = (u32)A.opt.max_occ;
        u32 step = c > mo ? c / mo : 1;
        u32 t = (c - 1) / step + 1;
        c = c == 0 ? 0 : (t < mo ? t : mo);
    }
    atomicAdd(&A.rd_nsmem[rid], 1u);
    atomicAdd((unsigned long long *)&A.rd_nhit[rid], (unsigned long long)c);
    // T1: long and rare SMEMs become reseed tasks (reference L923-947)
    if (KIND == 1 && A.opt.rounds >= 2 && end - start >= A.opt.split_len && c32 <= (u32)A.opt.split_width) {
        if (o.tslot == o.tend) { o.tslot = atomicAdd((unsigned long long *)&A.ctrl[C_T2], 8ull); o.tend = o.tslot + 8; }
        u64 t = o.tslot++;
        if (t < A.t2_cap) {
            Task2 x; x.read = rid; x.se = (u32)start | ((u32)end << 16); x.cnt = c32;
            x.full = full_match ? 1u : 0u; x.lo = full_match ? max_refpos : lo_; x.pad = 0;
            A.t2_out[t] = x;
        } else atomicExch((unsigned long long *)&A.ctrl[C_T2_OVF], 1ull);
    }
}

template <int KIND>
__device__ __noinline__ void finish_thread(const Args &A, OutState &o, const Stats &st)
{
    // reserved but unused slots must not look like records
    for (; o.pslot < o.pend; o.pslot++) if (o.pslot < A.pool_cap) A.pool_keys[o.pslot] = ~0ull;
    if (KIND == 1)
        for (; o.tslot < o.tend; o.tslot++)
            if (o.tslot < A.t2_cap) { Task2 x; x.read = 0xffffffffu; x.se = 0; x.cnt = 0; x.full = 0; x.lo = 0; x.pad = 0; A.t2_out[o.tslot] = x; }
    if (st.search | st.block | st.kary | st.ext | st.pac | st.iter) {
        atomicAdd((unsigned long long *)&A.ctrl[C_SEARCH], (unsigned long long)st.search);
        atomicAdd((unsigned long long *)&A.ctrl[C_BLOCK], (unsigned long long)st.block);
        atomicAdd((unsigned long long *)&A.ctrl[C_KARY], (unsigned long long)st.kary);
        atomicAdd((unsigned long long *)&A.ctrl[C_EXT], (unsigned long long)st.ext);
        atomicAdd((unsigned long long *)&A.ctrl[C_PAC], (unsigned long long)st.pac);
        atomicAdd((unsigned long long *)&A.ctrl[C_ITER], (unsigned long long)st.iter);
    }
    if (st.pass_ctl | st.pass_mem) {
        atomicAdd((unsigned long long *)&A.ctrl[C_PASS_CTL], (unsigned long long)st.pass_ctl);
        atomicAdd((unsigned long long *)&A.ctrl[C_PASS_MEM], (unsigned long long)st.pass_mem);
        atomicAdd((unsigned long long *)&A.ctrl[C_LANES_CTL], (unsigned long long)st.lanes_ctl);
        atomicAdd((unsigned long long *)&A.ctrl[C_LANES_MEM], (unsigned long long)st.lanes_mem);
    }
}

// ---- cell accessors: a task's cell of field F lives at region + (F * NT + sid) * sizeof(cell) ----
#define GRP_DI(name) idx_t &name = *reinterpret_cast<idx_t *>(smb + L::OFF_IDX + ((size_t)I_##name * L::NT + sid) * sizeof(idx_t))
#define GRP_DU(name) u32 &name = *reinterpret_cast<u32 *>(smb + L::OFF_U32 + ((size_t)U_##name * L::NT + sid) * 4)
#define GRP_DS(name) short &name = *reinterpret_cast<short *>(smb + L::OFF_S16 + ((size_t)S_##name * L::NT + sid) * 2)
#define GRP_RD(c) (*reinterpret_cast<const u32 *>(smb + L::OFF_RD + ((size_t)(c) * L::NT + sid) * 4))
#define GRP_LCP(j) (*reinterpret_cast<unsigned short *>(smb + L::OFF_LCP + ((size_t)(j) * L::NT + sid) * 2))
#define GRP_STGQ(q) (reinterpret_cast<uint4 *>(smb + L::OFF_STG + ((size_t)(q) * L::NT + sid) * 16))

// 32 bases of the forward read from `start` (any integer: bases outside the kept words are 0)
template <class L>
__device__ __forceinline__ u64 fwwin(const unsigned char *smb, u32 sid, int start)
{
    if (start < -32 || start >= 32 * L::RWT) return 0;
    const int w = start >> 5, sh = (start & 31) * 2;
    u64 a = 0, b = 0;
    if (w >= 0) a = (u64)GRP_RD(2 * w) | ((u64)GRP_RD(2 * w + 1) << 32);
    if (w + 1 < L::RWT) b = (u64)GRP_RD(2 * w + 2) | ((u64)GRP_RD(2 * w + 3) << 32);
    return sh == 0 ? a : ((a >> sh) | (b << (64 - sh)));
}

// RMI key of a query window (Tokenization, reference L795-901): first base in the most significant bits, 3-padded
__device__ __forceinline__ u64 rmi_key_of(u64 q0, int q_V)
{
    u64 k = q0;
    if (q_V < 32) k |= ~0ull << (2 * q_V);
    u64 y = __brevll(k);
    return ((y & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((y & 0x5555555555555555ull) << 1);
}

// =====================================================================================================================
// control pass: the state machine as a converged interpreter over its basic blocks
// =====================================================================================================================
// Every lane holds the label of the block its task executes next.  The warp repeatedly takes the SMALLEST label any
// lane holds (one REDUX) and the lanes at that label run the block together; a block ends by naming the next label.
// Labels are numbered along the flow of a search (entry blocks first, the shared downstream blocks -- located, level
// handling, driver continuations, start selection, search begin -- later), so lanes that split at a data-dependent
// branch wait at the merge points until the stragglers arrive.  A lane stops at L_STOP: its task has a memory request
// pending or has ended.  (In the flat kernel the same code is one goto-based function per lane; with the gotos a warp
// never reconverges after the first data-dependent branch: 3-6 active lanes per instruction, profiles/r2.)
enum {
    L_TASK_LOADED = 0, L_S_ISA, L_S_RMI1, L_S_RMI2, L_S_CENTER, L_S_LOC, L_LOC_INIT, L_LOC_NEXT, L_P_E, L_E_DONE, L_LOCATED,
    L_INIT_RIGHT, L_LV_RIGHT, L_LV_DONE, L_LEVELED, L_LEVEL_LOOP, L_LV_BEGIN, L_D3_LEVELS, L_SEARCH_DONE, L_D1_M, L_D2_C,
    L_C_D1_F, L_C_D1_L, L_C_D1_R, L_C_D2_F, L_C_D2_L, L_C_D2_R, L_C_D3_S, L_D2_LOOP, L_D1_LOOP, L_D1_ZZ, L_D2_FIRST, L_D3_LOOP,
    L_AUTO_START, L_S_BEGIN, L_E_BEGIN, L_E_PROBE, L_C_TASK, L_STOP
};
static_assert(L_STOP < 0x40, "labels double as state bytes below the pool codes");

template <int KIND, bool WIDE, bool LONG, bool STATS>
__device__ __forceinline__ void ctl_pass(const IndexView &ix, const Args &A, unsigned char *smb, const u32 sid, const bool have,
                                         OutState &out, Stats &st, bool &ended)
{
    typedef Lay<WIDE, LONG> L;
    typedef typename IdxT<WIDE>::type idx_t;
    typedef typename IdxT<WIDE>::stype sidx_t;
    const idx_t n = (idx_t)ix.sa_num;
    const int opt_msl = A.opt.min_seed_len;
    // ---- the task's cells ----
    GRP_DI(s_hint); GRP_DI(center); GRP_DI(lo_b); GRP_DI(hi_b); GRP_DI(k_stride); GRP_DI(k_pa); GRP_DI(base); GRP_DI(lo); GRP_DI(hi);
    GRP_DI(e_from); GRP_DI(e_cur); GRP_DI(e_edge); GRP_DI(max_refpos); GRP_DI(cache_refpos); GRP_DI(r3_last_lo); GRP_DI(m_base);
    GRP_DI(rad_lo); GRP_DI(rad_hi); GRP_DI(lv_limit); GRP_DI(e_s); GRP_DI(e_limit); GRP_DI(r3_last_cnt);
    sidx_t &m_stride = *reinterpret_cast<sidx_t *>(smb + L::OFF_IDX + ((size_t)I_m_stride * L::NT + sid) * sizeof(idx_t));
    GRP_DU(rid); GRP_DU(min_intv); GRP_DU(s_need); GRP_DU(q0lo); GRP_DU(q0hi);
    GRP_DS(state); GRP_DS(mreq); GRP_DS(L_); GRP_DS(pivot); GRP_DS(sp); GRP_DS(next); GRP_DS(cache_pivot); GRP_DS(cache_pivot_end);
    GRP_DS(q_strand); GRP_DS(q_from); GRP_DS(q_V); GRP_DS(s_pivot); GRP_DS(it); GRP_DS(nbl); GRP_DS(nbr); GRP_DS(lev); GRP_DS(e_L);
    GRP_DS(e_nb); GRP_DS(e_guard); GRP_DS(as_pv); GRP_DS(m_cap); GRP_DS(min_seed_len); GRP_DS(s_msl); GRP_DS(d_ret); GRP_DS(nvalid);
    GRP_DS(lv_ret); GRP_DS(e_dir); GRP_DS(e_cnt); GRP_DS(e_ret); GRP_DS(e_up); GRP_DS(e_sh); GRP_DS(as_ret); GRP_DS(as_right);
    GRP_DS(as_emit); GRP_DS(s_emit); GRP_DS(s_stepwise); GRP_DS(s_isa); GRP_DS(k_contig); GRP_DS(s_mode); GRP_DS(full_match);
    GRP_DS(use_cached); GRP_DS(m_count); GRP_DS(m_ldk); GRP_DS(r_less); GRP_DS(r_count); GRP_DS(rl_off); GRP_DS(bl_off);
    GRP_DS(anyn); GRP_DS(m_fast);
    const idx_t IDX_MAX = (idx_t)~(idx_t)0;

    // ---- helpers ----
    auto qword = [&](int k) -> u64 {
        const int f = q_from + 32 * k;
        const u64 v = fwwin<L>(smb, sid, q_strand == 0 ? f : L_ - 32 - f);
        return q_strand == 0 ? v : revcomp32(v);
    };
    auto nmw = [&](int w) -> u32 { return GRP_RD(2 * L::RWT + w); };
    auto is_n = [&](int p) -> bool { return anyn && ((nmw(p >> 5) >> (p & 31)) & 1); };
    auto next_n = [&](int from) -> int {
        if (!anyn) return L_;
        int w = from >> 5;
        u32 m = nmw(w) >> (from & 31);
        if (m) { int p = from + __ffs(m) - 1; return p < L_ ? p : L_; }
        for (w++; w * 32 < L_; w++) { u32 x = nmw(w); if (x) { int p = w * 32 + __ffs(x) - 1; return p < L_ ? p : L_; } }
        return L_;
    };
    auto prev_n = [&](int p) -> int {
        if (!anyn) return -1;
        int w = p >> 5;
        u32 m = nmw(w) & (0xffffffffu >> (31 - (p & 31)));
        if (m) return w * 32 + 31 - __clz(m);
        for (w--; w >= 0; w--) { u32 x = nmw(w); if (x) return w * 32 + 31 - __clz(x); }
        return -1;
    };
#define RL(j) ((int)GRP_LCP(rl_off + (j)))
#define BL(j) ((int)GRP_LCP(bl_off + (j)))
    auto lead_less = [&]() -> int { int c = __ffs(~(u32)(unsigned short)r_less) - 1; return c > r_count ? (int)r_count : c; };
    auto req_block = [&](idx_t b, sidx_t stride, int count, int cap) {
        m_base = b; m_stride = stride; m_count = (short)count; m_cap = (short)cap; r_count = (short)count;
        m_fast = (stride == 1 && count == 8 && (b & 7) == 0) ? 1 : 0;
        if (count > 0) mreq = M_BLOCK; else { mreq = M_NONE; r_less = 0; }
    };
    // a contiguous block of (up to) 8 entries that must cover [want_lo, want_hi]: the 64-byte-aligned block when it does
    // (one DRAM burst of keys + one sector of positions), else the block starting at `b`
    auto req_contig = [&](idx_t b, int cap, idx_t want_lo, idx_t want_hi) {
        if (n >= 8 && b > n - 8) b = n - 8;
        const idx_t al = want_lo & ~(idx_t)7;
        if (al + 8 <= n && want_hi < al + 8) b = al;
        int c = n - b < 8 ? (int)(n - b) : 8;
        base = b; nvalid = (short)c;
        req_block(b, 1, c, cap);
        if (STATS) st.block++;
    };
    auto push = [&](int start, int end, idx_t lo_, u64 cnt) {
        emit_record<KIND>(A, out, rid, start, end, (u64)lo_, cnt, full_match != 0, (u64)max_refpos);
    };
    auto req_rmi = [&](int table, u64 index) { mreq = M_RMI; m_ldk = (short)table; m_base = (idx_t)index; };
    auto req_ld = [&](int which, idx_t index) { mreq = M_LD; m_ldk = (short)which; m_base = index; };
    // results of the simple loads sit in the staging cells
    auto ld_val = [&]() -> idx_t {
        const uint4 v = *GRP_STGQ(0);
        if (!WIDE) return (idx_t)v.x;
        return (idx_t)((u64)v.x | ((u64)((v.y >> (8 * ((u32)m_base & 3))) & 0xff) << 32));
    };

    ended = false;
    int lab = L_STOP;
    if (have) { lab = state; mreq = M_NONE; }
// leave the block for label X / stop with a memory request that resumes at label X / after a request that may be empty
#define GO(X) do { lab = (X); goto next_block; } while (0)
#define STOP_AT(X) do { state = (short)(X); lab = L_STOP; goto next_block; } while (0)
#define REQ_OR_GO(X) do { if (mreq != M_NONE) STOP_AT(X); GO(X); } while (0)
    for (;;) {
        const int cur = GRP_REDUCE_MIN(lab);
        if (cur >= L_STOP) break;
        if (lab == cur) {
            switch (cur) {
            case L_TASK_LOADED:
                min_seed_len = (short)opt_msl;
                if (L_ <= 0) GO(L_C_TASK);
                if (KIND == 1) {
                    full_match = 0; use_cached = 0; max_refpos = 0; min_intv = 1;
                    pivot = 0;
                    GO(L_D1_LOOP);
                }
                if (KIND == 2) {
                    if (use_cached) {                            // cache_refpos = SA[lo] (reference L934-939)
                        req_ld(LD_POS, cache_refpos);
                        STOP_AT(L_D2_C);
                    }
                    GO(L_D2_FIRST);
                }
                // KIND == 3
                use_cached = 0; min_intv = (u32)A.opt.max_mem_intv; min_seed_len = (short)(opt_msl + 1);
                full_match = (ix.mode == 3 && A.rd_full[rid] != 0) ? 1 : 0;
                max_refpos = full_match ? (idx_t)A.rd_maxref[rid] : 0;
                pivot = 0;
                GO(L_D3_LOOP);

            // ------------------------------------------------ T1: zig-zag (reference L913-972, L1691-1894)
            case L_D1_LOOP:
                if (KIND != 1) GO(L_C_TASK);
                while (pivot < L_ && is_n(pivot)) pivot = (L_ - pivot < min_seed_len) ? L_ : (short)(pivot + 1);
                if (pivot >= L_) {
                    if (A.rd_full) { A.rd_full[rid] = full_match ? 1 : 0; A.rd_maxref[rid] = max_refpos; }
                    GO(L_C_TASK);
                }
                if (pivot != 0 && !is_n(pivot - 1)) { sp = pivot; GO(L_D1_ZZ); }
                // first search of a segment: always through the RMI
                s_pivot = pivot; q_strand = 0; s_emit = 1; s_need = 1; s_msl = min_seed_len; s_mode = ST_RMI; s_isa = 0;
                s_stepwise = 0; d_ret = L_C_D1_F;
                GO(L_S_BEGIN);
            case L_C_D1_F:
                if (KIND != 1) GO(L_C_TASK);
                if (ix.mode == 3 && lev == L_) {                 // whole read matched (L1870-1875)
                    req_ld(LD_POS, lo);
                    STOP_AT(L_D1_M);
                }
                pivot += lev;
                GO(L_D1_LOOP);
            case L_D1_M:
                if (KIND != 1) GO(L_C_TASK);
                full_match = 1; max_refpos = ld_val();
                pivot += lev;
                GO(L_D1_LOOP);
            case L_D1_ZZ:
                if (KIND != 1) GO(L_C_TASK);
                for (;;) {
                    if (sp >= L_) { pivot = L_; GO(L_D1_LOOP); }
                    if (!is_n(sp)) break;
                    if (L_ - sp < min_seed_len) { pivot = L_; sp = L_; }
                    else { sp += 1; pivot += 1; }
                }
                as_pv = pivot; as_right = 0; as_emit = 0; as_ret = L_C_D1_L;
                GO(L_AUTO_START);
            case L_C_D1_L:
                if (KIND != 1) GO(L_C_TASK);
                pivot = pivot - lev + 1;
                if (L_ - pivot < min_seed_len) { pivot = L_; GO(L_D1_LOOP); }
                as_pv = pivot; as_right = 1; as_emit = 1; as_ret = L_C_D1_R;
                GO(L_AUTO_START);
            case L_C_D1_R:
                if (KIND != 1) GO(L_C_TASK);
                if (pivot + lev <= sp) { pivot = L_; GO(L_D1_LOOP); }   // progress guard (see meme_seed_flat.cuh)
                sp = pivot + lev;
                pivot = sp;
                GO(L_D1_ZZ);

            // ------------------------------------------------ T2: reseeding (reference L1897-2126)
            case L_D2_C:
                if (KIND != 2) GO(L_C_TASK);
                cache_refpos = ld_val();
                GO(L_D2_FIRST);
            case L_D2_FIRST:
                as_pv = pivot; as_right = 1; as_emit = 0; as_ret = L_C_D2_F;
                GO(L_AUTO_START);
            case L_C_D2_F:
                if (KIND != 2) GO(L_C_TASK);
                next = pivot + lev; sp = pivot;
                GO(L_D2_LOOP);
            case L_D2_LOOP:
                if (sp >= next) GO(L_C_TASK);
                as_pv = pivot; as_right = 0; as_emit = 0; as_ret = L_C_D2_L;
                GO(L_AUTO_START);
            case L_C_D2_L:
                if (KIND != 2) GO(L_C_TASK);
                pivot = pivot - lev + 1;
                if (next - pivot < min_seed_len) GO(L_C_TASK);
                as_pv = pivot; as_right = 1; as_emit = 1; as_ret = L_C_D2_R;
                GO(L_AUTO_START);
            case L_C_D2_R:
                if (KIND != 2) GO(L_C_TASK);
                if (pivot + lev <= sp) GO(L_C_TASK);                    // progress guard
                sp = pivot + lev;
                pivot = sp;
                GO(L_D2_LOOP);

            // ------------------------------------------------ T3: third round (reference L974-1466)
            case L_D3_LOOP:
                if (KIND != 3) GO(L_C_TASK);
                for (;;) {
                    if (pivot >= L_ - min_seed_len + 1) break;
                    if (is_n(pivot)) { pivot++; continue; }
                    int V = next_n(pivot) - pivot;
                    if (V < min_seed_len) { pivot += V; continue; }
                    break;
                }
                if (pivot >= L_ - min_seed_len + 1) GO(L_C_TASK);
                s_pivot = pivot; q_strand = 0; s_emit = 0; s_need = min_intv; s_msl = min_seed_len; s_stepwise = 1;
                d_ret = L_C_D3_S;
                if (full_match) { s_mode = ST_EXACT; s_hint = max_refpos + (idx_t)pivot; s_isa = 1; }
                else { s_mode = ST_RMI; s_isa = 0; }
                GO(L_S_BEGIN);
            case L_C_D3_S:
                if (KIND != 3) GO(L_C_TASK);
                if (lev < min_seed_len) { pivot += min_seed_len; GO(L_D3_LOOP); }
                r3_last_lo = lo; r3_last_cnt = 0; it = 0;
                GO(L_D3_LEVELS);
            case L_D3_LEVELS:
                if (KIND != 3) GO(L_C_TASK);
                {
                    u64 cnt = (u64)(hi - lo) + 1;
                    if (cnt >= (u64)min_intv) {                                  // L1243-1251
                        if (r3_last_cnt > 0) push(pivot, pivot + max((int)lev + 1, (int)min_seed_len), r3_last_lo, (u64)r3_last_cnt);
                        pivot += lev + 1;
                        GO(L_D3_LOOP);
                    }
                    if (max(nbl, nbr) < min_seed_len || ++it > 400) {            // L1252-1258
                        push(pivot, pivot + min_seed_len, lo, cnt);
                        pivot += min_seed_len;
                        GO(L_D3_LOOP);
                    }
                    r3_last_lo = lo; r3_last_cnt = (idx_t)cnt;
                }
                lv_limit = (idx_t)min_intv; lv_ret = 1;
                GO(L_LV_BEGIN);

            // ------------------------------------------------ start selection (reference L1741-1816 / L1926-2040)
            case L_AUTO_START:
                s_pivot = as_pv; q_strand = as_right ? 0 : 1; s_emit = as_emit; s_need = min_intv; s_msl = min_seed_len;
                s_stepwise = 0; d_ret = as_ret; s_mode = ST_RMI; s_isa = 0;
                if (KIND != 1 && ix.mode == 3) {
                    if (full_match) {
                        s_mode = ST_EXACT; s_isa = 1;
                        s_hint = as_right ? max_refpos + (idx_t)as_pv : n - max_refpos - (idx_t)as_pv - 1;
                    } else if (KIND == 2 && use_cached) {
                        if (as_right && cache_pivot_end >= as_pv + 18 && as_pv >= cache_pivot) {
                            s_mode = ST_NEAR; s_isa = 1; s_hint = cache_refpos + (idx_t)(as_pv - cache_pivot);
                        } else if (!as_right && as_pv < cache_pivot_end && cache_pivot + 18 <= as_pv) {
                            s_mode = ST_NEAR; s_isa = 1; s_hint = n + (idx_t)cache_pivot - cache_refpos - (idx_t)as_pv - 1;
                        }
                    }
                }
                GO(L_S_BEGIN);
            // ------------------------------------------------ search (right_smem_search / mem_search / *_tradeoff)
            case L_S_BEGIN:
                if (STATS) st.search++;
                if (q_strand == 0) { q_from = s_pivot; q_V = (short)(next_n(s_pivot) - s_pivot); }
                else { q_from = (short)(L_ - 1 - s_pivot); q_V = (short)(s_pivot - prev_n(s_pivot)); }
                {
                    const u64 q0 = qword(0);
                    q0lo = (u32)q0; q0hi = (u32)(q0 >> 32);
                    if (KIND != 1 && s_isa) { req_ld(LD_ISA, s_hint); STOP_AT(L_S_ISA); }
                    const u64 key = rmi_key_of(q0, q_V);                 // RMI key = Tokenization (L795-901)
                    req_rmi(RMI_L2, key >> ix.bit_shift);
                }
                STOP_AT(L_S_RMI1);
            case L_S_ISA:
                if (KIND == 1) GO(L_C_TASK);
                s_hint = ld_val();
                if (s_mode == ST_EXACT) {
                    req_contig(s_hint > 3 ? s_hint - 3 : 0, q_V, s_hint > 0 ? s_hint - 1 : 0, s_hint + 1 < n ? s_hint + 1 : n - 1);
                    STOP_AT(L_LOCATED);
                }
                center = s_hint; rad_lo = rad_hi = 2;
                GO(L_LOC_INIT);
            case L_S_RMI1:
            case L_S_RMI2:
                {
                    const uint4 ra = *GRP_STGQ(0), rb = *GRP_STGQ(1);
                    const double rmi_a = __longlong_as_double((long long)((u64)ra.x | ((u64)ra.y << 32)));
                    const double rmi_b = __longlong_as_double((long long)((u64)ra.z | ((u64)ra.w << 32)));
                    const u64 rmi_e = (u64)rb.x | ((u64)rb.y << 32);
                    const u64 key = rmi_key_of((u64)q0lo | ((u64)q0hi << 32), q_V);
                    const double p = fma(rmi_b, __ull2double_rn(key), rmi_a);
                    if (cur == L_S_RMI1 && (rmi_e >> 63)) {
                        u64 s0 = (rmi_e >> 32) & 0x7fffffff;
                        double bound = (double)(rmi_e & 0xffffffffull) - 1;
                        u64 sub = p < 0.0 ? 0 : (p > bound ? (u64)bound : (u64)p);
                        req_rmi(RMI_L1, s0 + sub);
                        STOP_AT(L_S_RMI2);
                    }
                    double bound = (double)ix.sa_num - 1.0;
                    center = p < 0.0 ? 0 : (p > bound ? (idx_t)bound : (idx_t)p);
                    rad_lo = (idx_t)((rmi_e >> 32) & 0x3fffffff);
                    rad_hi = (idx_t)(rmi_e & 0x7fffffff);
                }
                GO(L_LOC_INIT);
            case L_LOC_INIT:
                lo_b = 0; hi_b = n; it = 0;
                // prediction errors concentrate near zero: try the block around the prediction first
                if ((u64)rad_lo + (u64)rad_hi + 2 > 8 && (u64)rad_lo + (u64)rad_hi <= 40) {
                    req_contig(center > 4 ? center - 4 : 0, q_V, center > 1 ? center - 1 : 0, center + 2 < n ? center + 2 : n - 1);
                    STOP_AT(L_S_CENTER);
                }
                GO(L_LOC_NEXT);
            case L_S_CENTER:
                {
                    int c = lead_less();
                    bool left_ok = c > 0 || base == 0, right_ok = c < nvalid || base + (idx_t)nvalid == n;
                    if (left_ok && right_ok) GO(L_LOCATED);
                    if (!left_ok) hi_b = base; else lo_b = base + (idx_t)nvalid;
                }
                GO(L_LOC_NEXT);
            case L_LOC_NEXT:
                {
                    u64 c64 = center;
                    u64 pa64 = c64 > (u64)rad_lo ? c64 - (u64)rad_lo : 0;
                    u64 pb64 = c64 + (u64)rad_hi + 1;
                    idx_t pa = pa64 < (u64)lo_b ? lo_b : (pa64 > (u64)hi_b ? hi_b : (idx_t)pa64);
                    idx_t pb = pb64 < (u64)lo_b ? lo_b : (pb64 > (u64)hi_b ? hi_b : (idx_t)pb64);
                    idx_t span = pb - pa;
                    k_contig = span <= 6;
                    if (k_contig) {
                        idx_t slack = (6 - span) / 2;
                        req_contig(pa > 1 + slack ? pa - 1 - slack : 0, q_V, pa > 0 ? pa - 1 : 0, pb < n ? pb : n - 1);
                    } else {
                        k_stride = (span + 7) / 8; k_pa = pa;
                        idx_t first = pa + k_stride - 1;
                        int cnt = 0;
                        if (first < pb) { idx_t c = (pb - 1 - first) / k_stride + 1; cnt = c < 8 ? (int)c : 8; }
                        req_block(first, (sidx_t)k_stride, cnt, q_V);
                        if (STATS) st.kary++;
                    }
                }
                REQ_OR_GO(L_S_LOC);
            case L_S_LOC:
                if (k_contig) {
                    int c = lead_less();
                    bool left_ok = c > 0 || base == 0, right_ok = c < nvalid || base + (idx_t)nvalid == n;
                    if (left_ok && right_ok) GO(L_LOCATED);
                    if (!left_ok) hi_b = base; else lo_b = base + (idx_t)nvalid;
                } else {
                    int c = lead_less();
                    if (c > 0) { idx_t nl = k_pa + (idx_t)c * k_stride; if (nl > lo_b) lo_b = nl; }
                    if (c < r_count) { idx_t nh = k_pa + (idx_t)(c + 1) * k_stride - 1; if (nh < hi_b) hi_b = nh; }
                }
                rad_lo = rad_lo > (IDX_MAX >> 4) ? IDX_MAX >> 1 : rad_lo * 8 + 8;
                rad_hi = rad_hi > (IDX_MAX >> 4) ? IDX_MAX >> 1 : rad_hi * 8 + 8;
                if (++it < 64) GO(L_LOC_NEXT);
                GO(L_LOCATED);                                   // not reached for a sorted suffix array: use the last block
            case L_LOCATED:
                // ---- first level = maximal lcp in the block, [lo,hi] = its interval ----
                bl_off = rl_off; rl_off ^= 8;                   // keep this block's LCPs; later probes use the other buffer
                {
                    int m = -1, first = 0;
#pragma unroll
                    for (int j = 0; j < 8; j++) { int l = j < nvalid ? BL(j) : -1; if (l > m) { m = l; first = j; } }
                    lev = (short)m;
                    int last = first;
                    while (last + 1 < nvalid && BL(last + 1) >= lev) last++;
                    lo = base + (idx_t)first; hi = base + (idx_t)last;
                    const bool exact = s_emit && lev >= s_msl;
                    if (!exact && (u64)(hi - lo) + 1 >= (u64)s_need) { nbl = nbr = lev; GO(L_LEVELED); }
                    lv_limit = exact ? IDX_MAX : (idx_t)s_need;
                    nbl = first > 0 ? (short)BL(first - 1) : (base == 0 ? 0 : -1);
                    nbr = last < nvalid - 1 ? (short)BL(last + 1) : (base + (idx_t)nvalid == n ? 0 : -1);
                }
                if (nbl < 0) { e_dir = -1; e_from = lo; e_L = lev; e_limit = lv_limit; e_ret = R_INIT_L; GO(L_E_BEGIN); }
                GO(L_INIT_RIGHT);
            case L_INIT_RIGHT:
                if (nbr < 0) { e_dir = +1; e_from = hi; e_L = lev; e_limit = lv_limit; e_ret = R_INIT_R; GO(L_E_BEGIN); }
                GO(L_LEVELED);
            case L_LEVELED:
                if (KIND == 3 && s_stepwise) GO(L_SEARCH_DONE);
                it = 0;
                GO(L_LEVEL_LOOP);
            case L_LEVEL_LOOP:
                if (KIND != 1 && (u64)(hi - lo) + 1 < (u64)s_need && it++ < 400) {
                    lv_limit = (s_emit && max(nbl, nbr) >= s_msl) ? IDX_MAX : (idx_t)s_need;
                    lv_ret = 0;
                    GO(L_LV_BEGIN);
                }
                if (KIND != 3 && s_emit && lev >= s_msl) push(s_pivot, s_pivot + lev, lo, (u64)(hi - lo) + 1);
                GO(L_SEARCH_DONE);
            case L_SEARCH_DONE:
                GO(d_ret);                                       // the driver continuation of this search

            // ------------------------------------------------ one level drop (reference L2569-2574)
            case L_LV_BEGIN:
                if (KIND == 1) GO(L_C_TASK);
                lev = nbl > nbr ? nbl : nbr;
                if (lev <= 0) { lev = 0; lo = 0; hi = n - 1; nbl = nbr = 0; GO(L_LV_DONE); }
                if (nbl >= lev) {
                    if (lo > base && lo <= base + (idx_t)nvalid) {            // lo-1 lies in the located block
                        int z = (int)(lo - base) - 1;
                        while (z >= 0 && BL(z) >= lev) z--;
                        if (z < 0) { lo = base; nbl = base == 0 ? 0 : -1; }
                        else { lo = base + (idx_t)(z + 1); nbl = (short)BL(z); }
                    } else nbl = lo == 0 ? 0 : -1;
                    if (nbl < 0) { e_dir = -1; e_from = lo; e_L = lev; e_limit = lv_limit; e_ret = R_LV_L; GO(L_E_BEGIN); }
                }
                GO(L_LV_RIGHT);
            case L_LV_RIGHT:
                if (nbr >= lev) {
                    if (hi + 1 >= base && hi + 1 < base + (idx_t)nvalid) {    // hi+1 lies in the located block
                        int z = (int)(hi + 1 - base);
                        while (z < nvalid && BL(z) >= lev) z++;
                        if (z >= nvalid) { hi = base + (idx_t)nvalid - 1; nbr = base + (idx_t)nvalid == n ? 0 : -1; }
                        else { hi = base + (idx_t)z - 1; nbr = (short)BL(z); }
                    } else nbr = hi == n - 1 ? 0 : -1;
                    if (nbr < 0) { e_dir = +1; e_from = hi; e_L = lev; e_limit = lv_limit; e_ret = R_LV_R; GO(L_E_BEGIN); }
                }
                GO(L_LV_DONE);
            case L_LV_DONE:
                if (lv_ret == 0) GO(L_LEVEL_LOOP);
                GO(L_D3_LEVELS);

            // ------------------------------------------------ extension: 8-ary gallop, then 8-ary narrowing
            case L_E_BEGIN:
                e_cur = e_from; e_s = 1; e_sh = 0; e_up = 1; e_guard = 0;
                GO(L_E_PROBE);
            case L_E_PROBE:
                if (++e_guard > 400) { e_edge = e_cur; e_nb = 0; GO(L_E_DONE); }
                if (e_dir > 0) {
                    u64 room = (u64)(n - 1 - e_cur) >> e_sh;
                    e_cnt = room < 8 ? (short)room : 8;
                    req_block(e_cur + e_s, (sidx_t)e_s, e_cnt, e_L);
                } else {
                    u64 room = (u64)e_cur >> e_sh;
                    e_cnt = room < 8 ? (short)room : 8;
                    req_block(e_cur - e_s, -(sidx_t)e_s, e_cnt, e_L);
                }
                if (STATS) st.ext++;
                REQ_OR_GO(L_P_E);
            case L_P_E:
                {
                    int c = 0;
                    while (c < e_cnt && RL(c) >= e_L) c++;
                    e_cur = e_dir > 0 ? e_cur + (idx_t)c * e_s : e_cur - (idx_t)c * e_s;
                    if (c == 8) {
                        idx_t gone = e_dir > 0 ? e_cur - e_from : e_from - e_cur;
                        if (gone >= e_limit) { e_edge = e_cur; e_nb = e_L; GO(L_E_DONE); }
                        if (e_up && e_s <= (n >> 3)) { e_s *= 8; e_sh += 3; }
                        GO(L_E_PROBE);
                    }
                    if (e_s == 1) { e_nb = c < e_cnt ? (short)RL(c) : 0; e_edge = e_cur; GO(L_E_DONE); }
                    e_up = 0;
                    e_s >>= 3; e_sh -= 3;
                }
                GO(L_E_PROBE);
            case L_E_DONE:
                if (e_ret == R_INIT_L) { lo = e_edge; nbl = e_nb; GO(L_INIT_RIGHT); }
                if (e_ret == R_INIT_R) { hi = e_edge; nbr = e_nb; GO(L_LEVELED); }
                if (e_ret == R_LV_L) { lo = e_edge; nbl = e_nb; GO(L_LV_RIGHT); }
                hi = e_edge; nbr = e_nb;
                GO(L_LV_DONE);
            default:                                             // L_C_TASK: the task has ended
                ended = true;
                mreq = M_NONE;
                lab = L_STOP;
                break;
            }
        }
    next_block:;
    }
#undef GO
#undef STOP_AT
#undef REQ_OR_GO
#undef RL
#undef BL
}

// =====================================================================================================================
// issue + compare steps
// =====================================================================================================================
// the copies of one block probe: 8 keys (two per 16-byte cell) + 8 positions (+ the words holding the high bytes)
template <bool WIDE, bool LONG>
__device__ __forceinline__ void issue_block(const IndexView &ix, unsigned char *smb, const u32 sid)
{
    typedef Lay<WIDE, LONG> L;
    typedef typename IdxT<WIDE>::type idx_t;
    typedef typename IdxT<WIDE>::stype sidx_t;
    GRP_DI(m_base);
    const sidx_t m_stride = *reinterpret_cast<sidx_t *>(smb + L::OFF_IDX + ((size_t)I_m_stride * L::NT + sid) * sizeof(idx_t));
    GRP_DS(m_count); GRP_DS(m_fast);
    const bool has_keys = ix.keys != nullptr;
    if (m_fast) {
        // aligned block: 64 bytes of keys + 32 bytes of positions (+ 8 high bytes), 16-byte copies
        const char *kp = reinterpret_cast<const char *>(ix.keys + m_base);
        const char *pq = reinterpret_cast<const char *>(ix.pos_lo + m_base);
        if (has_keys) {
#pragma unroll
            for (int q = 0; q < 4; q++) GRP_CP16(GRP_STGQ(q), kp + 16 * q);
        }
        GRP_CP16(GRP_STGQ(4), pq);
        GRP_CP16(GRP_STGQ(5), pq + 16);
        if (WIDE) GRP_CP8(GRP_STGQ(6), reinterpret_cast<const char *>(ix.pos_hi + m_base));
    } else {
        idx_t idx = m_base;
        char *kd = reinterpret_cast<char *>(GRP_STGQ(0));
        char *pd = reinterpret_cast<char *>(GRP_STGQ(4));
        char *hd = reinterpret_cast<char *>(GRP_STGQ(L::NQ - 2));
#pragma unroll 1
        for (int j = 0; j < m_count; j++) {
            // entry j: key -> cell j/2, half j&1; position -> cell 4 + j/4, word j&3
            if (has_keys) GRP_CP8(kd + (size_t)(j >> 1) * (L::NT * 16) + 8 * (j & 1), ix.keys + idx);
            GRP_CP4(pd + (size_t)(j >> 2) * (L::NT * 16) + 4 * (j & 3), ix.pos_lo + idx);
            if (WIDE) GRP_CP4(hd + (size_t)(j >> 2) * (L::NT * 16) + 4 * (j & 3), ix.pos_hi + (idx & ~(idx_t)3));
            idx += (idx_t)m_stride;
        }
    }
}

// position of entry j of the staged block
template <bool WIDE, bool LONG>
__device__ __forceinline__ u64 staged_pos(const IndexView &ix, unsigned char *smb, const u32 sid, int j)
{
    typedef Lay<WIDE, LONG> L;
    typedef typename IdxT<WIDE>::type idx_t;
    typedef typename IdxT<WIDE>::stype sidx_t;
    const u32 *pw = reinterpret_cast<const u32 *>(GRP_STGQ(4 + (j >> 2)));
    u64 p = pw[j & 3];
    if (WIDE) {
        GRP_DI(m_base); GRP_DS(m_fast);
        if (m_fast) {
            const u32 *hw = reinterpret_cast<const u32 *>(GRP_STGQ(6));
            p |= (u64)((hw[j >> 2] >> (8 * (j & 3))) & 0xff) << 32;
        } else {
            const sidx_t m_stride = *reinterpret_cast<sidx_t *>(smb + L::OFF_IDX + ((size_t)I_m_stride * L::NT + sid) * sizeof(idx_t));
            const idx_t idx = m_base + (idx_t)((sidx_t)j * m_stride);
            const u32 *hw = reinterpret_cast<const u32 *>(GRP_STGQ(L::NQ - 2 + (j >> 2)));
            p |= (u64)((hw[j & 3] >> (8 * ((u32)idx & 3))) & 0xff) << 32;
        }
    }
    return p;
}

// compare the staged block with the query: LCP of every entry (capped), "suffix < query" bits, entries whose first 32
// bases match and need the packed reference (compare_read_and_ref_binary, reference L226-601)
template <bool WIDE, bool LONG>
__device__ __forceinline__ void compare_block(const IndexView &ix, unsigned char *smb, const u32 sid)
{
    typedef Lay<WIDE, LONG> L;
    GRP_DS(m_count); GRP_DS(m_cap); GRP_DS(rl_off); GRP_DS(r_less); GRP_DS(undet); GRP_DS(pac_e); GRP_DS(pac_k0); GRP_DS(mreq);
    GRP_DU(q0lo); GRP_DU(q0hi);
    const bool has_keys = ix.keys != nullptr;
    u32 less = 0, und = 0;
    const u32 qlo = q0lo, qhi = q0hi;
    const int cnt = m_count, cap0 = m_cap, ro = rl_off;
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
        if (2 * jj < cnt) {
            const uint4 kq = *GRP_STGQ(jj);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int j = 2 * jj + h;
                if (j < cnt) {
                    const u32 klo = h ? kq.z : kq.x, khi = h ? kq.w : kq.y;
                    const u64 p = staged_pos<WIDE, LONG>(ix, smb, sid, j);
                    const u64 ref_len = ix.sa_num - p;
                    const int cap = ref_len < (u64)cap0 ? (int)ref_len : cap0;
                    int l = 0;
                    if (!has_keys) und |= 1u << j;
                    else {
                        // first differing base in 32-bit halves (16 bases each), branch-free; xx == 0: all 32 bases equal
                        const u32 xl = klo ^ qlo, xh = khi ^ qhi;
                        const bool hi_half = xl == 0;
                        const u32 xx = hi_half ? xh : xl, rr = hi_half ? khi : klo, qq = hi_half ? qhi : qlo;
                        const int sh = (__ffs((int)xx) - 1) & 30;
                        const int b = xx ? (sh >> 1) + (hi_half ? 16 : 0) : 32;
                        l = min(b, cap);
                        if (xx != 0 && b < cap && ((rr >> sh) & 3) < ((qq >> sh) & 3)) less |= 1u << j;
                        if (xx == 0 && cap > 32) und |= 1u << j;
                    }
                    GRP_LCP(ro + j) = (unsigned short)l;
                }
            }
        }
    }
    r_less = (short)less; undet = (short)und;
    if (und) { pac_e = (short)(__ffs(und) - 1); pac_k0 = has_keys ? 1 : 0; mreq = M_PAC; }
    else mreq = M_NONE;
}

// up to five 64-bit words of packed reference for entry pac_e, from word pac_k0 of its suffix (only the words the
// capped compare can reach: the text has two words of slack, not five)
template <bool WIDE, bool LONG>
__device__ __forceinline__ void issue_pac(const IndexView &ix, unsigned char *smb, const u32 sid)
{
    typedef Lay<WIDE, LONG> L;
    GRP_DS(pac_e); GRP_DS(pac_k0); GRP_DS(m_cap);
    const u64 p = staged_pos<WIDE, LONG>(ix, smb, sid, pac_e);
    const u64 ref_len = ix.sa_num - p;
    const int cap = ref_len < (u64)m_cap ? (int)ref_len : (int)m_cap;
    const u64 *wp = ix.pac + (p >> 5) + pac_k0;
    char *d = reinterpret_cast<char *>(GRP_STGQ(0));
#pragma unroll
    for (int j = 0; j < 5; j++)
        if (j == 0 || 32 * (pac_k0 + j - 1) < cap) GRP_CP8(d + (size_t)(j >> 1) * (L::NT * 16) + 8 * (j & 1), wp + j);
}

template <bool WIDE, bool LONG, bool STATS>
__device__ __forceinline__ void compare_pac(const IndexView &ix, unsigned char *smb, const u32 sid, Stats &st)
{
    typedef Lay<WIDE, LONG> L;
    GRP_DS(m_cap); GRP_DS(rl_off); GRP_DS(r_less); GRP_DS(undet); GRP_DS(pac_e); GRP_DS(pac_k0); GRP_DS(mreq);
    GRP_DS(q_strand); GRP_DS(q_from); GRP_DS(L_);
    GRP_DU(q0lo); GRP_DU(q0hi);
    if (STATS) st.pac++;
    const bool has_keys = ix.keys != nullptr;
    const u64 p = staged_pos<WIDE, LONG>(ix, smb, sid, pac_e);
    const u64 ref_len = ix.sa_num - p;
    const int cap = ref_len < (u64)m_cap ? (int)ref_len : (int)m_cap;
    const int sh = (int)(p & 31) * 2;
    int res = -1; bool less = false;
    const uint4 a = *GRP_STGQ(0), b = *GRP_STGQ(1), c = *GRP_STGQ(2);
    u64 w0 = (u64)a.x | ((u64)a.y << 32), w1 = (u64)a.z | ((u64)a.w << 32), w2 = (u64)b.x | ((u64)b.y << 32),
        w3 = (u64)b.z | ((u64)b.w << 32), w4 = (u64)c.x | ((u64)c.y << 32);
    const u64 q0 = (u64)q0lo | ((u64)q0hi << 32);
#pragma unroll 1
    for (int j = 0; j < 4 && res < 0; j++) {
        int k = pac_k0 + j;
        if (32 * k >= cap) res = cap;
        else {
            u64 rr = sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0;
            u64 qq;
            if (k == 0) qq = q0;
            else {
                const int f = q_from + 32 * k;
                const u64 v = fwwin<L>(smb, sid, q_strand == 0 ? f : L_ - 32 - f);
                qq = q_strand == 0 ? v : revcomp32(v);
            }
            u64 x = rr ^ qq;
            if (x) {
                int bb = (__ffsll((long long)x) - 1) >> 1;
                int l = 32 * k + bb;
                if (l >= cap) res = cap;
                else { res = l; less = ((rr >> (2 * bb)) & 3) < ((qq >> (2 * bb)) & 3); }
            }
        }
        w0 = w1; w1 = w2; w2 = w3; w3 = w4;
    }
    if (res < 0) pac_k0 += 4;                      // four more words of the same entry
    else {
        GRP_LCP(rl_off + pac_e) = (unsigned short)res;
        u32 und = (u32)(unsigned short)undet & ~(1u << pac_e);
        if (less) r_less = (short)((u32)(unsigned short)r_less | (1u << pac_e));
        undet = (short)und;
        if (und) { pac_e = (short)(__ffs(und) - 1); pac_k0 = has_keys ? 1 : 0; }
        else mreq = M_NONE;
    }
}

// the simple loads: one P-RMI record, one position / inverse-SA entry
template <bool WIDE, bool LONG>
__device__ __forceinline__ void issue_simple(const IndexView &ix, unsigned char *smb, const u32 sid)
{
    typedef Lay<WIDE, LONG> L;
    typedef typename IdxT<WIDE>::type idx_t;
    GRP_DI(m_base); GRP_DS(mreq); GRP_DS(m_ldk);
    if (mreq == M_RMI) {
        const RmiRec *r = (m_ldk == RMI_L2 ? ix.l2 : ix.l1) + m_base;
        GRP_CP16(GRP_STGQ(0), reinterpret_cast<const char *>(r));
        GRP_CP16(GRP_STGQ(1), reinterpret_cast<const char *>(r) + 16);
    } else {                                        // M_LD
        const u32 *lo = m_ldk == LD_POS ? ix.pos_lo : ix.isa_lo;
        char *d = reinterpret_cast<char *>(GRP_STGQ(0));
        GRP_CP4(d, lo + m_base);
        if (WIDE) {
            const uint8_t *hi = m_ldk == LD_POS ? ix.pos_hi : ix.isa_hi;
            GRP_CP4(d + 4, hi + (m_base & ~(idx_t)3));
        }
    }
}

// the packed read of a task: RWT 64-bit words + RWT N-mask words into the read cells, the two offsets into staging
template <bool WIDE, bool LONG>
__device__ __forceinline__ void issue_task(const Args &A, unsigned char *smb, const u32 sid, u32 rid)
{
    typedef Lay<WIDE, LONG> L;
    const u64 *src = A.packed + (u64)rid * TPR_RW;
    const u32 *nsrc = A.nmask + (u64)rid * TPR_RW;
    char *rd = reinterpret_cast<char *>(smb + L::OFF_RD + (size_t)sid * 4);
#pragma unroll
    for (int w = 0; w < L::RWT; w++) {
        GRP_CP4(rd + (size_t)(2 * w) * (L::NT * 4), reinterpret_cast<const u32 *>(src + w));
        GRP_CP4(rd + (size_t)(2 * w + 1) * (L::NT * 4), reinterpret_cast<const u32 *>(src + w) + 1);
        GRP_CP4(rd + (size_t)(2 * L::RWT + w) * (L::NT * 4), nsrc + w);
    }
    char *d = reinterpret_cast<char *>(GRP_STGQ(0));
    GRP_CP8(d, A.read_off + rid);
    GRP_CP8(d + 8, A.read_off + rid + 1);
}

// =====================================================================================================================
// the kernel
// =====================================================================================================================
__device__ __forceinline__ u32 zero_bytes(u32 x) { return (x - 0x01010101u) & ~x & 0x80808080u; }
__device__ __forceinline__ uint4 ld_state16(const volatile unsigned char *p)
{   // the 16 state bytes of a lane class, read afresh (other warps move slots between pools)
    const volatile u32 *w = reinterpret_cast<const volatile u32 *>(p);
    uint4 v; v.x = w[0]; v.y = w[1]; v.z = w[2]; v.w = w[3];
    return v;
}

template <int KIND, bool WIDE, bool LONG, bool STATS>
__global__ void __launch_bounds__(1024, 1) k_grp(IndexView ix, Args A)
{
    typedef Lay<WIDE, LONG> L;
    typedef typename IdxT<WIDE>::type idx_t;
    GRP_SMEM_DECL;
    unsigned char *const smb = grp_smem;
    const int lane = GRP_LANE();
    // [0..3] tokens of the four pools, [4..7] task counts of the pools, [8] live slots
    volatile u32 *const tok = reinterpret_cast<volatile u32 *>(smb + L::OFF_TOK);
    volatile int *const cnt = reinterpret_cast<volatile int *>(smb + L::OFF_TOK) + 4;
    volatile unsigned char *const sst = smb + L::OFF_SST + lane * 16;      // this lane class's state bytes

    // ---- init: every slot empty (slots beyond K are dead), every token free ----
    const u32 lanes_here = blockDim.x < 32u ? blockDim.x : 32u;
    for (u32 i = threadIdx.x; i < 512; i += blockDim.x) smb[L::OFF_SST + i] = (i & 15) < (u32)L::K ? (unsigned char)Q_EMPTY : (unsigned char)SL_DEAD;
    for (u32 i = threadIdx.x; i < 32; i += blockDim.x) tok[i] = (i == 8 || i == 4 + PL_EMPTY) ? (u32)L::K * lanes_here : 0u;
    GRP_SYNCTHREADS();

    OutState out; out.pslot = out.pend = out.tslot = out.tend = 0;
    Stats st = {};
    u32 pend[GRP_DEPTH]; int pkind[GRP_DEPTH];
#pragma unroll
    for (int d = 0; d < GRP_DEPTH; d++) { pend[d] = GRP_NONE; pkind[d] = B_SIMPLE; }
    int npend = 0;

    // move tasks between pools: the state byte of each moved slot, then one counter update per destination pool
    auto publish = [&](const u32 sid, const int newbyte) {
        GRP_FENCE();
        if (newbyte >= 0) sst[sid >> 5] = (unsigned char)newbyte;
        const int np = newbyte < 0 ? -1 : (newbyte < 0x40 ? (int)PL_CTL : (newbyte == SL_DEAD ? -1 : newbyte - 0x40 + 1));
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const u32 m = GRP_BALLOT(np == p);
            if (m && lane == 0) atomicAdd(const_cast<int *>(&cnt[p]), __popc(m));
        }
    };

    for (;;) {
        // ================================ pick a pool: the fullest one ================================
        int pool = -1, pool_n = 0;
        if (npend < GRP_DEPTH) {
            const int c0 = cnt[PL_CTL], c1 = cnt[PL_BLOCK], c2 = cnt[PL_PAC], c3 = cnt[PL_EMPTY];
            pool = PL_CTL; pool_n = c0;
            if (c1 > pool_n) { pool = PL_BLOCK; pool_n = c1; }
            if (c2 > pool_n) { pool = PL_PAC; pool_n = c2; }
            if (c3 > pool_n) { pool = PL_EMPTY; pool_n = c3; }
            // counters are read by every lane on its own: agree on lane 0's view
            pool = GRP_SHFL(pool, 0); pool_n = GRP_SHFL(pool_n, 0);
            if (pool_n <= 0) pool = -1;
        }
        const bool take = pool >= 0 && (pool_n >= A.grp_th || npend == 0);
        GRP_TRACE("pass: pool %d n %d npend %d take %d live %u\n", pool, pool_n, npend, (int)take, tok[8]);
        if (!take) {
            if (npend == 0) {                       // nothing ready, nothing pending
                if (tok[8] == 0) break;             // every slot is dead: done
                GRP_SLEEP(100);
                continue;
            }
            // ================================ retire the oldest pending batch ================================
            const u32 sid = pend[0]; const int kind = pkind[0];
#pragma unroll
            for (int d = 0; d + 1 < GRP_DEPTH; d++) { pend[d] = pend[d + 1]; pkind[d] = pkind[d + 1]; }
            pend[GRP_DEPTH - 1] = GRP_NONE;
            npend--;
            if (npend == 0) GRP_CP_WAIT(0); else if (npend == 1) GRP_CP_WAIT(1); else if (npend == 2) GRP_CP_WAIT(2); else GRP_CP_WAIT(3);
            u32 requeue = GRP_NONE; int newbyte = -1;
            if (sid != GRP_NONE) {
                GRP_DS(mreq); GRP_DS(state);
                GRP_TRACE("  retire sid %u kind %d mreq %d state %d\n", sid, kind, (int)mreq, (int)state);
                if (kind == B_BLOCK) {
                    compare_block<WIDE, LONG>(ix, smb, sid);
                    newbyte = mreq == M_PAC ? (int)Q_PAC : (int)state;
                } else if (kind == B_PAC) {
                    compare_pac<WIDE, LONG, STATS>(ix, smb, sid, st);
                    newbyte = mreq == M_PAC ? (int)Q_PAC : (int)state;
                } else if (mreq == M_T2REC) {
                    // reseed task record (emitted by T1): read, SMEM [start,end), hit count, start hint
                    const uint4 a = *GRP_STGQ(0), b = *GRP_STGQ(1);
                    GRP_DU(rid); GRP_DU(min_intv); GRP_DS(pivot); GRP_DS(full_match); GRP_DS(use_cached); GRP_DS(cache_pivot);
                    GRP_DS(cache_pivot_end); GRP_DI(max_refpos); GRP_DI(cache_refpos);
                    rid = a.x;
                    const u32 se = a.y, c = a.z, full = a.w;
                    const u64 lo64 = (u64)b.x | ((u64)b.y << 32);
                    const int qb = se & 0xffff, qe = se >> 16;
                    pivot = (short)((qb + qe) >> 1);
                    min_intv = c + 1;
                    full_match = full != 0;
                    max_refpos = full ? (idx_t)lo64 : 0;
                    use_cached = (!full && ix.mode == 3 && c >= 1 && c < 10) ? 1 : 0;
                    cache_pivot = (short)qb; cache_pivot_end = (short)qe; cache_refpos = (idx_t)lo64;
                    if (a.x == 0xffffffffu) newbyte = Q_EMPTY;       // unused task slot
                    else { mreq = M_TASK; issue_task<WIDE, LONG>(A, smb, sid, a.x); requeue = sid; }
                } else if (mreq == M_TASK) {
                    const uint4 o = *GRP_STGQ(0);
                    GRP_DS(L_); GRP_DS(anyn); GRP_DS(rl_off); GRP_DS(bl_off);
                    const long long o0 = (long long)((u64)o.x | ((u64)o.y << 32)), o1 = (long long)((u64)o.z | ((u64)o.w << 32));
                    L_ = (short)(o1 - o0);
                    u32 any = 0;
#pragma unroll
                    for (int w = 0; w < L::RWT; w++) any |= GRP_RD(2 * L::RWT + w);
                    anyn = any != 0;
                    rl_off = 0; bl_off = 8;
                    state = L_TASK_LOADED;
                    newbyte = L_TASK_LOADED;
                } else {
                    newbyte = (int)state;           // M_RMI / M_LD: the result waits in the staging cells
                }
                if (STATS) { st.iter++; }
            }
            if (STATS) { st.pass_mem++; st.lanes_mem += sid != GRP_NONE; }
            const bool any_rq = GRP_BALLOT(requeue != GRP_NONE) != 0;
            if (any_rq) {
                GRP_CP_COMMIT();
#pragma unroll
                for (int d = 0; d < GRP_DEPTH; d++) if (d == npend) { pend[d] = requeue; pkind[d] = B_SIMPLE; }
                npend++;
            }
            publish(sid, newbyte);
            continue;
        }
        // ================================ claim one slot of the pool per lane ================================
        u32 got = 0;
        if (lane == 0) got = atomicCAS(const_cast<u32 *>(&tok[pool]), 0u, 1u) == 0u;
        got = GRP_SHFL(got, 0);
        if (!got) continue;                          // another warp is claiming from this pool right now
        GRP_FENCE();
        u32 sid = GRP_NONE;
        {
            uint4 v = ld_state16(sst);
            u32 z0, z1, z2, z3;
            if (pool == PL_CTL) {                    // any control label: the two top bits of the byte are clear
                z0 = zero_bytes(v.x & 0xc0c0c0c0u); z1 = zero_bytes(v.y & 0xc0c0c0c0u);
                z2 = zero_bytes(v.z & 0xc0c0c0c0u); z3 = zero_bytes(v.w & 0xc0c0c0c0u);
            } else {
                const u32 pat = (u32)(0x40 + pool - 1) * 0x01010101u;
                z0 = zero_bytes(v.x ^ pat); z1 = zero_bytes(v.y ^ pat); z2 = zero_bytes(v.z ^ pat); z3 = zero_bytes(v.w ^ pat);
            }
            int k = -1;
            if (z0) k = (__ffs(z0) - 1) >> 3;
            else if (z1) k = 4 + ((__ffs(z1) - 1) >> 3);
            else if (z2) k = 8 + ((__ffs(z2) - 1) >> 3);
            else if (z3) k = 12 + ((__ffs(z3) - 1) >> 3);
            if (k >= 0) { sst[k] = (unsigned char)SL_BUSY; sid = (u32)k * 32 + lane; }
        }
        {
            const u32 m = GRP_BALLOT(sid != GRP_NONE);
            if (lane == 0 && m) atomicSub(const_cast<int *>(&cnt[pool]), __popc(m));
        }
        GRP_FENCE();
        GRP_SYNCWARP();
        if (lane == 0) tok[pool] = 0u;
        // ================================ run the step ================================
        int newbyte = -1;                            // pool the task moves to now (else it waits in this warp's FIFO)
        bool queued = false; int bkind = B_SIMPLE;
        const bool have = sid != GRP_NONE;
        const u32 sid_ = have ? sid : (u32)lane;     // lanes without a task compute addresses only
        if (pool == PL_EMPTY) {
            // ---- start tasks: one global counter bump for the whole warp ----
            const u32 m = GRP_BALLOT(have);
            u64 t0 = 0;
            if (lane == 0 && m) t0 = atomicAdd((unsigned long long *)&A.ctrl[C_WORK], (unsigned long long)__popc(m));
            t0 = GRP_SHFL(t0, 0);
            if (have) {
                const u64 task = t0 + (u64)__popc(m & ((1u << lane) - 1));
                if (task >= (u64)A.n_tasks) {
                    newbyte = SL_DEAD;
                    atomicSub(const_cast<u32 *>(&tok[8]), 1u);
                } else {
                    GRP_DS(mreq);
                    if (KIND == 2) {
                        mreq = M_T2REC;
                        const char *t2 = reinterpret_cast<const char *>(A.t2_in + task);
                        GRP_CP16(GRP_STGQ(0), t2);
                        GRP_CP16(GRP_STGQ(1), t2 + 16);
                    } else {
                        GRP_DU(rid);
                        rid = (u32)task;
                        mreq = M_TASK;
                        issue_task<WIDE, LONG>(A, smb, sid, (u32)task);
                    }
                    queued = true;
                }
            }
        } else if (pool == PL_BLOCK) {
            if (have) { issue_block<WIDE, LONG>(ix, smb, sid); queued = true; }
            bkind = B_BLOCK;
            if (STATS) { st.pass_mem++; st.lanes_mem += have; }
        } else if (pool == PL_PAC) {
            if (have) { issue_pac<WIDE, LONG>(ix, smb, sid); queued = true; }
            bkind = B_PAC;
            if (STATS) { st.pass_mem++; st.lanes_mem += have; }
        } else {
            bool ended;
            ctl_pass<KIND, WIDE, LONG, STATS>(ix, A, smb, sid_, have, out, st, ended);
            if (have) {
                const u32 sid = sid_;
                GRP_DS(mreq);
                if (ended) newbyte = Q_EMPTY;
                else if (mreq == M_BLOCK) newbyte = Q_BLOCK;
                else { issue_simple<WIDE, LONG>(ix, smb, sid); queued = true; }
            }
            if (STATS) { st.pass_ctl++; st.lanes_ctl += have; }
        }
        GRP_SYNCWARP();
        if (GRP_BALLOT(queued) != 0) {
            GRP_CP_COMMIT();
#pragma unroll
            for (int d = 0; d < GRP_DEPTH; d++) if (d == npend) { pend[d] = queued ? sid : GRP_NONE; pkind[d] = bkind; }
            npend++;
        }
        publish(sid, newbyte);
    }
    finish_thread<KIND>(A, out, st);
}

}  // namespace grp
