// meme_seed_flat.cuh -- the thread-per-task seeding kernel as ONE flat state machine (the default kernel A).
//
// One lane owns one task (T1: zig-zag pass of a read, T2: one reseed, T3: third round of a read) and alternates
// between a control phase -- its own path through `switch (state)` until it needs memory -- and a memory phase the
// whole warp runs together.  What shapes the code (profiles/r1/experiments):
//   * the warp is reconverged explicitly at the head of every iteration, between the load and the compare phase
//     and at every pass through the switch; the goto-based control flow has no reconvergence point of its own and
//     left warps in independently looping fragments (3 active lanes per instruction even in the loads);
//   * a block probe (8 suffix-array entries) is staged with cp.async into per-lane shared words and compared by a
//     rolled, branch-free loop; LCPs stay in shared memory;
//   * code size is budgeted against the 32 KB instruction cache: every piece of control code exists once
//     (continuation codes instead of call sites), cold paths are __noinline__, the work counters are compiled out
//     unless STATS, loops over the 8 entries are rolled;
//   * suffix-array indices and positions are 32-bit unless the index is wide (sa_num >= 2^32), a template
//     parameter, which halves the integer work of the control code.
// Included by meme_seed.cu (launch sequence, sort, gather) and by tests/host_sim (single-lane CPU run of this source).
#pragma once
#include "meme_seed_task.cuh"

// The few device-only constructs go through these macros so that tests/host_sim can compile this file with g++ and
// run the same control code one lane at a time against the oracle (no GPU needed to check a logic change).
#ifndef MEME_HOST_SIM
#define FLAT_SMEM_DECL extern __shared__ u32 smem[]
typedef u32 flat_stage_t;                                   // shared-window byte address
#define FLAT_STAGE_BASE(p) ((u32)__cvta_generic_to_shared(p))
#define FLAT_CP4(dst, src) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory")
#define FLAT_CP_COMMIT() asm volatile("cp.async.commit_group;" ::: "memory")
#define FLAT_CP_WAIT() asm volatile("cp.async.wait_group 0;" ::: "memory")
#endif

namespace tpr {

enum {
    P_TASK = 0, P_TASK_LOADED, P_S_ISA, P_S_EXACT, P_S_RMI1, P_S_RMI2, P_S_CENTER, P_S_LOC, P_E,
    P_D1_L, P_D1_R, P_D1_F, P_D1_M, P_D2_C, P_D2_F, P_D2_L, P_D2_R, P_D3_S, P_D3_LV, P_AUTO, P_SBEGIN, P_LOCATED
};
enum { R_INIT_L = 0, R_INIT_R, R_LV_L, R_LV_R };

// per-thread shared words: [0,2) zero pair (read word -1), [2,22) packed read words 0..9, [22,24) zero pair
// (read word 10), [24,34) N-mask words, the OR of the N masks, then FLAT_STG staging words (stride 75: odd, conflict-free)
#define FLAT_SW 35
#define FLAT_RD 2
#define FLAT_NM 24
#define FLAT_ANYN 34                      // OR of the N-mask words (the pad word)

// Tuning switches (profiles/r1/experiments has the measurements behind the defaults).
// Control rounds per iteration: a lane that needs more than this many passes through the switch before its next
// memory request sits the memory phase out and joins the first round of the next iteration, where more lanes share
// its state (the late rounds run with 1-3 lanes).
#ifndef FLAT_MAXROUNDS
#define FLAT_MAXROUNDS 1000
#endif
#ifndef FLAT_VOTE
#define FLAT_VOTE 1                       // 1: lanes re-enter the control switch together (vote-controlled loop)
#endif
// Shared control code (start selection, search begin, located) can be entered as a state of its own, so that lanes
// arriving from different states run it together.  Measured: more active lanes per instruction, but ~300 more
// instructions of code and more passes through the switch; no gain for any kind once the kernels fit the instruction
// cache (profiles/r1/experiments), so it is off.
#ifndef FLAT_MERGE
#define FLAT_MERGE(KIND) 0
#endif
#define FLAT_GOTO(label, st) do { if (FLAT_MERGE(KIND)) { state = st; goto next_state; } else goto label; } while (0)
// Block probes are staged in shared memory: cp.async copies the 8 keys (two halves) and positions straight into the
// thread's staging words (a rolled loop with every load in flight and no register staging), the rolled compare loop
// reads them back and leaves the 8 LCPs in one of two LCP buffers (the block a search is located in keeps its LCPs
// while extension probes use the other buffer).  Against fully unrolled register code this took ~600 instructions
// out of the hot loop, which is what lets the level/extension code of T2 and T3 stay inside the 32 KB instruction
// cache ("no instruction" stalls 3.8 -> 0.6 per issue, ncu; profiles/r1/experiments).
#ifndef FLAT_BLK_UNROLL
#define FLAT_BLK_UNROLL 1
#endif
#define FLAT_STG 40                       // key low halves [0,8), key high halves [8,16), positions [16,24), LCPs [24,32) [32,40)
#define FLAT_LCP0 24
#define FLAT_LCP1 32
template <int KIND> __host__ __device__ constexpr int flat_sw() { return FLAT_SW + FLAT_STG; }

template <bool WIDE> struct IdxT { typedef u32 type; };
template <> struct IdxT<true> { typedef u64 type; };

// Cold paths are kept out of the hot loop's instruction footprint (and their state out of registers):
// emitting an SMEM record happens ~8 times per read, loading a task ~2 times, against ~135 round trips.
struct OutState { u64 pslot, pend, tslot, tend, n_hit; u32 n_emit; };

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
    o.n_emit++;
    u32 c = c32;
    if (A.opt.max_occ > 0) {
        u32 mo = (u32)A.opt.max_occ;
        u32 step = c > mo ? c / mo : 1;
        u32 t = (c - 1) / step + 1;
        c = c == 0 ? 0 : (t < mo ? t : mo);
    }
    o.n_hit += c;
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

__device__ __noinline__ void flush_counts(const Args &A, OutState &o, u32 rid)
{
    if (o.n_emit) {
        atomicAdd(&A.rd_nsmem[rid], o.n_emit);
        atomicAdd((unsigned long long *)&A.rd_nhit[rid], (unsigned long long)o.n_hit);
        o.n_emit = 0; o.n_hit = 0;
    }
}

__device__ __noinline__ void load_task_words(const Args &A, u32 rid, u32 *sm)
{
    const u64 *src = A.packed + (u64)rid * TPR_RW;
    const u32 *nsrc = A.nmask + (u64)rid * TPR_RW;
    u32 any = 0;
#pragma unroll
    for (int w = 0; w < TPR_RW; w++) {
        u64 v = __ldg(src + w);
        sm[FLAT_RD + 2 * w] = (u32)v; sm[FLAT_RD + 2 * w + 1] = (u32)(v >> 32);
        u32 m = __ldg(nsrc + w);
        sm[FLAT_NM + w] = m;
        any |= m;
    }
    sm[FLAT_ANYN] = any;                   // most reads have no N: next_n / prev_n answer without scanning
}

template <int KIND>
__device__ __noinline__ void finish_thread(const Args &A, OutState &o, u32 st_search, u32 st_block, u32 st_kary, u32 st_ext,
                                           u32 st_pac, u32 st_iter)
{
    // reserved but unused slots must not look like records
    for (; o.pslot < o.pend; o.pslot++) if (o.pslot < A.pool_cap) A.pool_keys[o.pslot] = ~0ull;
    if (KIND == 1)
        for (; o.tslot < o.tend; o.tslot++)
            if (o.tslot < A.t2_cap) { Task2 x; x.read = 0xffffffffu; x.se = 0; x.cnt = 0; x.full = 0; x.lo = 0; x.pad = 0; A.t2_out[o.tslot] = x; }
    atomicAdd((unsigned long long *)&A.ctrl[C_SEARCH], (unsigned long long)st_search);
    atomicAdd((unsigned long long *)&A.ctrl[C_BLOCK], (unsigned long long)st_block);
    atomicAdd((unsigned long long *)&A.ctrl[C_KARY], (unsigned long long)st_kary);
    atomicAdd((unsigned long long *)&A.ctrl[C_EXT], (unsigned long long)st_ext);
    atomicAdd((unsigned long long *)&A.ctrl[C_PAC], (unsigned long long)st_pac);
    atomicAdd((unsigned long long *)&A.ctrl[C_ITER], (unsigned long long)st_iter);
}

template <int KIND, bool WIDE, int MINB, bool SYNC, bool STATS>
#ifdef FLAT_MAXNREG                                    // experiment: register cap independent of the CTA shape
__global__ void __maxnreg__(FLAT_MAXNREG) k_flat(IndexView ix, Args A)
#else
__global__ void __launch_bounds__(TPR_THREADS, MINB) k_flat(IndexView ix, Args A)
#endif
{
    typedef typename IdxT<WIDE>::type idx_t;
    constexpr int BU = FLAT_BLK_UNROLL;
    constexpr int SW = flat_sw<KIND>();
    FLAT_SMEM_DECL;
    u32 *const sm = &smem[threadIdx.x * SW];
    u32 *const stg = sm + FLAT_SW;
    sm[0] = sm[1] = sm[FLAT_RD + 2 * TPR_RW] = sm[FLAT_RD + 2 * TPR_RW + 1] = 0;
    const idx_t n = (idx_t)ix.sa_num;
    const bool has_keys = ix.keys != nullptr;
    const int opt_msl = A.opt.min_seed_len;

    // ---- task ----
    u32 rid = 0; int L = 0; u64 task = 0;
    int state = P_TASK;
    // ---- request + probe results ----
    int mreq = M_NONE, m_count = 0, m_cap = 0, m_ldk = 0;
    idx_t m_base = 0; long long m_stride = 1; const RmiRec *m_rec = nullptr;
    u32 r_less = 0, undet = 0; int r_count = 0;
    int rl_off = FLAT_LCP0, bl_off = FLAT_LCP1;          // LCP buffers of the latest probe / of the located block
    u64 pphi = 0; int pac_e = 0, pac_k0 = 0;
    u64 ld_val = 0; double rmi_a = 0, rmi_b = 0; u64 rmi_e = 0;
    // ---- query ----
    int q_strand = 0, q_from = 0, q_V = 0; u64 q0 = 0;
    // ---- search ----
    bool s_emit = false, s_stepwise = false, s_isa = false, k_contig = false;
    int s_pivot = 0, s_need = 1, s_msl = 19, s_mode = ST_RMI, d_ret = 0, it = 0;
    idx_t s_hint = 0, center = 0, lo_b = 0, hi_b = 0, k_stride = 0, k_pa = 0, base = 0;
    u64 rad_lo = 0, rad_hi = 0;
    int nvalid = 0;
    idx_t lo = 0, hi = 0; int nbl = 0, nbr = 0, lev = 0;
    u64 lv_limit = 0; int lv_ret = 0;
    // ---- extension ----
    int e_dir = 0, e_L = 0, e_cnt = 0, e_nb = 0, e_guard = 0, e_ret = 0; bool e_up = true;
    idx_t e_from = 0, e_cur = 0, e_edge = 0; u64 e_s = 1, e_limit = 0; int e_sh = 0;
    // ---- driver ----
    int pivot = 0, sp = 0, next = 0, min_intv = 1, min_seed_len = opt_msl, cache_pivot = 0, cache_pivot_end = 0;
    bool full_match = false, use_cached = false;
    idx_t max_refpos = 0, cache_refpos = 0, r3_last_lo = 0; u64 r3_last_cnt = 0;
    // auto-start scratch
    int as_pv = 0, as_ret = 0; bool as_right = false, as_emit = false;
    // ---- output ----
    OutState out; out.pslot = out.pend = out.tslot = out.tend = out.n_hit = 0; out.n_emit = 0;
    u32 st_search = 0, st_block = 0, st_kary = 0, st_ext = 0, st_pac = 0, st_iter = 0;

    // ------------------------------------------------------------------ helpers (lambdas: inlined once per use)
    // 32 bases of the forward read from `start` (any integer: bases outside the read are 0); read words -1 and
    // TPR_RW are zero pads, so the window needs no per-word bounds checks
    auto fwwin = [&](int start) -> u64 {
        if (start < -32 || start >= 32 * TPR_RW) return 0;
        const int w = start >> 5, sh = (start & 31) * 2;
        const u32 *p = sm + FLAT_RD + 2 * w;
        const u64 a = (u64)p[0] | ((u64)p[1] << 32), b = (u64)p[2] | ((u64)p[3] << 32);
        return sh == 0 ? a : ((a >> sh) | (b << (64 - sh)));
    };
    auto qword = [&](int k) -> u64 {
        const int f = q_from + 32 * k;
        const u64 v = fwwin(q_strand == 0 ? f : L - 32 - f);
        return q_strand == 0 ? v : revcomp32(v);
    };
    auto is_n = [&](int p) -> bool { return (sm[FLAT_NM + (p >> 5)] >> (p & 31)) & 1; };
    auto next_n = [&](int from) -> int {
        if (!sm[FLAT_ANYN]) return L;
        int w = from >> 5;
        u32 m = sm[FLAT_NM + w] >> (from & 31);
        if (m) { int p = from + __ffs(m) - 1; return p < L ? p : L; }
        for (w++; w * 32 < L; w++) { u32 x = sm[FLAT_NM + w]; if (x) { int p = w * 32 + __ffs(x) - 1; return p < L ? p : L; } }
        return L;
    };
    auto prev_n = [&](int p) -> int {
        if (!sm[FLAT_ANYN]) return -1;
        int w = p >> 5;
        u32 m = sm[FLAT_NM + w] & (0xffffffffu >> (31 - (p & 31)));
        if (m) return w * 32 + 31 - __clz(m);
        for (w--; w >= 0; w--) { u32 x = sm[FLAT_NM + w]; if (x) return w * 32 + 31 - __clz(x); }
        return -1;
    };
#define RL(j) ((int)stg[rl_off + (j)])
#define BL(j) ((int)stg[bl_off + (j)])
    auto lead_less = [&]() -> int { int c = __ffs(~r_less) - 1; return c > r_count ? r_count : c; };
    auto req_block = [&](idx_t b, long long stride, int count, int cap) {
        m_base = b; m_stride = stride; m_count = count; m_cap = cap; r_count = count;
        if (count > 0) mreq = M_BLOCK; else { mreq = M_NONE; r_less = 0; }
    };
    auto req_contig = [&](idx_t b, int cap) {
        if (n >= 8 && b > n - 8) b = n - 8;
        int c = n - b < 8 ? (int)(n - b) : 8;
        base = b; nvalid = c;
        req_block(b, 1, c, cap);
        if (STATS) st_block++;
    };
    auto push = [&](int start, int end, idx_t lo_, u64 cnt) {
        emit_record<KIND>(A, out, rid, start, end, (u64)lo_, cnt, full_match, (u64)max_refpos);
    };

    // SYNC: the warp reconverges at the head of every iteration (lanes that ran out of tasks idle in the loop
    // until the whole warp is done), so that the memory and compare phases run with every lane of a request
    // kind together.  Without it the irreducible control flow below leaves the warp split in independently
    // looping fragments (ncu: 3 active lanes per instruction even in the block-probe loads, profiles/r1).
    bool done = false;
    for (;;) {
        if (SYNC) {
            __syncwarp();
            if (__all_sync(0xffffffffu, done)) break;
        }
#if FLAT_VOTE
        for (int round = 0; round < FLAT_MAXROUNDS && __any_sync(0xffffffffu, !done && mreq == M_NONE); round++) if (!done && mreq == M_NONE) {
#else
        while (!done && mreq == M_NONE) {
#endif
            // =========================== control: one flat state machine ===========================
            switch (state) {
            case P_TASK:
            c_task:                                         // task ends jump here: no extra pass through the switch
                flush_counts(A, out, rid);
                task = atomicAdd((unsigned long long *)&A.ctrl[C_WORK], 1ull);
                if (task >= (u64)A.n_tasks) { done = true; break; }
                if (KIND == 2) { mreq = M_T2REC; }
                else { rid = (u32)task; mreq = M_TASK; }
                state = P_TASK_LOADED;
                break;

            case P_TASK_LOADED:
                if (KIND == 2 && rid == 0xffffffffu) { state = P_TASK; break; }          // unused task slot
                L = (int)(A.read_off[rid + 1] - A.read_off[rid]);
                min_seed_len = opt_msl;
                if (L <= 0) { state = P_TASK; break; }
                if (KIND == 1) {
                    full_match = false; use_cached = false; max_refpos = 0; min_intv = 1;
                    pivot = 0;
                    goto d1_loop;
                }
                if (KIND == 2) {
                    if (use_cached) {                        // cache_refpos = SA[lo] (reference L934-939)
                        mreq = M_LD; m_ldk = LD_POS; m_base = cache_refpos;
                        state = P_D2_C;
                        break;
                    }
                    goto d2_first;
                }
                // KIND == 3
                use_cached = false; min_intv = A.opt.max_mem_intv; min_seed_len = opt_msl + 1;
                full_match = ix.mode == 3 && A.rd_full[rid] != 0;
                max_refpos = full_match ? (idx_t)A.rd_maxref[rid] : 0;
                pivot = 0;
                goto d3_loop;

            // ------------------------------------------------ T1: zig-zag (reference L913-972, L1691-1894)
            d1_loop:
                if (KIND != 1) { state = P_TASK; break; }
                while (pivot < L && is_n(pivot)) pivot = (L - pivot < min_seed_len) ? L : pivot + 1;
                if (pivot >= L) {
                    if (A.rd_full) { A.rd_full[rid] = full_match ? 1 : 0; A.rd_maxref[rid] = max_refpos; }
                    goto c_task;
                }
                if (pivot != 0 && !is_n(pivot - 1)) { sp = pivot; goto d1_zz; }
                // first search of a segment: always through the RMI
                s_pivot = pivot; q_strand = 0; s_emit = true; s_need = 1; s_msl = min_seed_len; s_mode = ST_RMI; s_isa = false;
                s_stepwise = false; d_ret = P_D1_F;
                FLAT_GOTO(s_begin, P_SBEGIN);
            case P_D1_F:
            c_d1_f:
                if (KIND != 1) { state = P_TASK; break; }
                if (ix.mode == 3 && lev == L) {              // whole read matched (L1870-1875)
                    mreq = M_LD; m_ldk = LD_POS; m_base = lo;
                    state = P_D1_M;
                    break;
                }
                pivot += lev;
                goto d1_loop;
            case P_D1_M:
                if (KIND != 1) { state = P_TASK; break; }
                full_match = true; max_refpos = (idx_t)ld_val;
                pivot += lev;
                goto d1_loop;
            d1_zz:
                if (KIND != 1) { state = P_TASK; break; }
                if (sp >= L) { pivot = L; goto d1_loop; }
                if (is_n(sp)) {
                    if (L - sp < min_seed_len) { pivot = L; sp = L; }
                    else { sp += 1; pivot += 1; }
                    goto d1_zz;
                }
                as_pv = pivot; as_right = false; as_emit = false; as_ret = P_D1_L;
                FLAT_GOTO(auto_start, P_AUTO);
            case P_D1_L:
            c_d1_l:
                if (KIND != 1) { state = P_TASK; break; }
                pivot = pivot - lev + 1;
                if (L - pivot < min_seed_len) { pivot = L; goto d1_loop; }
                as_pv = pivot; as_right = true; as_emit = true; as_ret = P_D1_R;
                FLAT_GOTO(auto_start, P_AUTO);
            case P_D1_R:
            c_d1_r:
                if (KIND != 1) { state = P_TASK; break; }
                // progress guard (not in the reference, which loops forever here when the text ends with its longest
                // T run and a match runs into the text end: DESIGN.md section 2): the zig-zag must move forward
                if (pivot + lev <= sp) { pivot = L; goto d1_loop; }
                sp = pivot + lev;
                pivot = sp;
                goto d1_zz;

            // ------------------------------------------------ T2: reseeding (reference L1897-2126)
            case P_D2_C:
                if (KIND != 2) { state = P_TASK; break; }
                cache_refpos = (idx_t)ld_val;
            d2_first:
                as_pv = pivot; as_right = true; as_emit = false; as_ret = P_D2_F;
                FLAT_GOTO(auto_start, P_AUTO);
            case P_D2_F:
            c_d2_f:
                if (KIND != 2) { state = P_TASK; break; }
                next = pivot + lev; sp = pivot;
            d2_loop:
                if (sp >= next) goto c_task;
                as_pv = pivot; as_right = false; as_emit = false; as_ret = P_D2_L;
                FLAT_GOTO(auto_start, P_AUTO);
            case P_D2_L:
            c_d2_l:
                if (KIND != 2) { state = P_TASK; break; }
                pivot = pivot - lev + 1;
                if (next - pivot < min_seed_len) goto c_task;
                as_pv = pivot; as_right = true; as_emit = true; as_ret = P_D2_R;
                FLAT_GOTO(auto_start, P_AUTO);
            case P_D2_R:
            c_d2_r:
                if (KIND != 2) { state = P_TASK; break; }
                if (pivot + lev <= sp) goto c_task;             // progress guard, as in the zig-zag of T1
                sp = pivot + lev;
                pivot = sp;
                goto d2_loop;

            // ------------------------------------------------ T3: third round (reference L974-1466)
            d3_loop:
                if (KIND != 3) { state = P_TASK; break; }
                for (;;) {
                    if (pivot >= L - min_seed_len + 1) break;
                    if (is_n(pivot)) { pivot++; continue; }
                    int V = next_n(pivot) - pivot;
                    if (V < min_seed_len) { pivot += V; continue; }
                    break;
                }
                if (pivot >= L - min_seed_len + 1) goto c_task;
                s_pivot = pivot; q_strand = 0; s_emit = false; s_need = min_intv; s_msl = min_seed_len; s_stepwise = true;
                d_ret = P_D3_S;
                if (full_match) { s_mode = ST_EXACT; s_hint = max_refpos + pivot; s_isa = true; }
                else { s_mode = ST_RMI; s_isa = false; }
                FLAT_GOTO(s_begin, P_SBEGIN);
            case P_D3_S:
            c_d3_s:
                if (KIND != 3) { state = P_TASK; break; }
                if (lev < min_seed_len) { pivot += min_seed_len; goto d3_loop; }
                r3_last_lo = lo; r3_last_cnt = 0; it = 0;
            d3_levels:
                if (KIND != 3) { state = P_TASK; break; }
                {
                    u64 cnt = (u64)(hi - lo) + 1;
                    if (cnt >= (u64)min_intv) {                                  // L1243-1251
                        if (r3_last_cnt > 0) push(pivot, pivot + max(lev + 1, min_seed_len), r3_last_lo, r3_last_cnt);
                        pivot += lev + 1;
                        goto d3_loop;
                    }
                    if (max(nbl, nbr) < min_seed_len || ++it > 400) {            // L1252-1258
                        push(pivot, pivot + min_seed_len, lo, cnt);
                        pivot += min_seed_len;
                        goto d3_loop;
                    }
                    r3_last_lo = lo; r3_last_cnt = cnt;
                }
                lv_limit = (u64)min_intv; lv_ret = 1;
                goto lv_begin;
            case P_D3_LV:                                   // (only reached through lv_done)
                goto d3_levels;

            // ------------------------------------------------ start selection (reference L1741-1816 / L1926-2040)
            case P_AUTO:
            auto_start:
                s_pivot = as_pv; q_strand = as_right ? 0 : 1; s_emit = as_emit; s_need = min_intv; s_msl = min_seed_len;
                s_stepwise = false; d_ret = as_ret; s_mode = ST_RMI; s_isa = false;
                if (KIND != 1 && ix.mode == 3) {
                    if (full_match) {
                        s_mode = ST_EXACT; s_isa = true;
                        s_hint = as_right ? max_refpos + as_pv : n - max_refpos - as_pv - 1;
                    } else if (KIND == 2 && use_cached) {
                        if (as_right && cache_pivot_end >= as_pv + 18 && as_pv >= cache_pivot) {
                            s_mode = ST_NEAR; s_isa = true; s_hint = cache_refpos + as_pv - cache_pivot;
                        } else if (!as_right && as_pv < cache_pivot_end && cache_pivot + 18 <= as_pv) {
                            s_mode = ST_NEAR; s_isa = true; s_hint = n + cache_pivot - cache_refpos - as_pv - 1;
                        }
                    }
                }
                FLAT_GOTO(s_begin, P_SBEGIN);
            // ------------------------------------------------ search (right_smem_search / mem_search / *_tradeoff)
            case P_SBEGIN:
            s_begin:
                if (STATS) st_search++;
                if (q_strand == 0) { q_from = s_pivot; q_V = next_n(s_pivot) - s_pivot; }
                else { q_from = L - 1 - s_pivot; q_V = s_pivot - prev_n(s_pivot); }
                q0 = qword(0);
                if (KIND != 1 && s_isa) { mreq = M_LD; m_ldk = LD_ISA; m_base = s_hint; state = P_S_ISA; break; }
                goto s_rmi;
            case P_S_ISA:
                if (KIND == 1) { state = P_TASK; break; }
                s_hint = (idx_t)ld_val;
                if (s_mode == ST_EXACT) { req_contig(s_hint > 3 ? s_hint - 3 : 0, q_V); state = P_S_EXACT; break; }
                center = s_hint; rad_lo = rad_hi = 2;
                goto loc_init;
            case P_S_EXACT:
                FLAT_GOTO(located, P_LOCATED);
            s_rmi:
                {
                    u64 k = q0;                             // RMI key = Tokenization (L795-901)
                    if (q_V < 32) k |= ~0ull << (2 * q_V);
                    u64 y = __brevll(k);
                    ld_val = ((y & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((y & 0x5555555555555555ull) << 1);
                    mreq = M_RMI; m_rec = ix.l2 + (ld_val >> ix.bit_shift);
                    state = P_S_RMI1;
                }
                break;
            case P_S_RMI1:
                if (rmi_e >> 63) {
                    double p = fma(rmi_b, __ull2double_rn(ld_val), rmi_a);
                    u64 st = (rmi_e >> 32) & 0x7fffffff;
                    double bound = (double)(rmi_e & 0xffffffffull) - 1;
                    u64 sub = p < 0.0 ? 0 : (p > bound ? (u64)bound : (u64)p);
                    mreq = M_RMI; m_rec = ix.l1 + st + sub;
                    state = P_S_RMI2;
                    break;
                }
                // fall through
            case P_S_RMI2:
                {
                    double p = fma(rmi_b, __ull2double_rn(ld_val), rmi_a);
                    double bound = (double)ix.sa_num - 1.0;
                    center = p < 0.0 ? 0 : (p > bound ? (idx_t)bound : (idx_t)p);
                    rad_lo = (rmi_e >> 32) & 0x3fffffff;
                    rad_hi = rmi_e & 0x7fffffff;
                }
            loc_init:
                lo_b = 0; hi_b = n; it = 0;
                // prediction errors concentrate near zero: try the block centred on the prediction first
                if (rad_lo + rad_hi + 2 > 8 && rad_lo + rad_hi <= 40) {
                    req_contig(center > 4 ? center - 4 : 0, q_V);
                    state = P_S_CENTER;
                    break;
                }
                goto loc_next;
            case P_S_CENTER:
                {
                    int c = lead_less();
                    bool left_ok = c > 0 || base == 0, right_ok = c < nvalid || base + nvalid == n;
                    if (left_ok && right_ok) FLAT_GOTO(located, P_LOCATED);
                    if (!left_ok) hi_b = base; else lo_b = base + nvalid;
                }
            loc_next:
                {
                    u64 c64 = center;
                    u64 pa64 = c64 > rad_lo ? c64 - rad_lo : 0;
                    u64 pb64 = c64 + rad_hi + 1;
                    idx_t pa = pa64 < (u64)lo_b ? lo_b : (pa64 > (u64)hi_b ? hi_b : (idx_t)pa64);
                    idx_t pb = pb64 < (u64)lo_b ? lo_b : (pb64 > (u64)hi_b ? hi_b : (idx_t)pb64);
                    idx_t span = pb - pa;
                    k_contig = span <= 6;
                    if (k_contig) {
                        idx_t slack = (6 - span) / 2;
                        req_contig(pa > 1 + slack ? pa - 1 - slack : 0, q_V);
                    } else {
                        k_stride = (span + 7) / 8; k_pa = pa;
                        idx_t first = pa + k_stride - 1;
                        int cnt = 0;
                        if (first < pb) { idx_t c = (pb - 1 - first) / k_stride + 1; cnt = c < 8 ? (int)c : 8; }
                        req_block(first, (long long)k_stride, cnt, q_V);
                        if (STATS) st_kary++;
                    }
                    state = P_S_LOC;
                }
                break;
            case P_S_LOC:
                if (k_contig) {
                    int c = lead_less();
                    bool left_ok = c > 0 || base == 0, right_ok = c < nvalid || base + nvalid == n;
                    if (left_ok && right_ok) FLAT_GOTO(located, P_LOCATED);
                    if (!left_ok) hi_b = base; else lo_b = base + nvalid;
                } else {
                    int c = lead_less();
                    if (c > 0) { idx_t nl = k_pa + (idx_t)c * k_stride; if (nl > lo_b) lo_b = nl; }
                    if (c < r_count) { idx_t nh = k_pa + (idx_t)(c + 1) * k_stride - 1; if (nh < hi_b) hi_b = nh; }
                }
                rad_lo = rad_lo * 8 + 8;
                rad_hi = rad_hi * 8 + 8;
                if (++it < 64) goto loc_next;
                // not reached for a sorted suffix array: use the last block
            case P_LOCATED:
            located:
                // ---- first level = maximal lcp in the block, [lo,hi] = its interval ----
                bl_off = rl_off; rl_off ^= FLAT_LCP0 ^ FLAT_LCP1;       // keep this block's LCPs; later probes use the other buffer
                {
                    int m = -1, first = 0;
#pragma unroll
                    for (int j = 0; j < 8; j++) { int l = j < nvalid ? BL(j) : -1; if (l > m) { m = l; first = j; } }
                    lev = m;
                    int last = first;
                    while (last + 1 < nvalid && BL(last + 1) >= lev) last++;
                    lo = base + first; hi = base + last;
                    const bool exact = s_emit && lev >= s_msl;
                    if (!exact && (u64)(hi - lo) + 1 >= (u64)s_need) { nbl = nbr = lev; goto leveled; }
                    lv_limit = exact ? ~0ull : (u64)s_need;
                    nbl = first > 0 ? BL(first - 1) : (base == 0 ? 0 : -1);
                    nbr = last < nvalid - 1 ? BL(last + 1) : (base + nvalid == n ? 0 : -1);
                }
                if (nbl < 0) { e_dir = -1; e_from = lo; e_L = lev; e_limit = lv_limit; e_ret = R_INIT_L; goto e_begin; }
            init_right:
                if (nbr < 0) { e_dir = +1; e_from = hi; e_L = lev; e_limit = lv_limit; e_ret = R_INIT_R; goto e_begin; }
            leveled:
                if (KIND == 3 && s_stepwise) goto search_done;
                it = 0;
            level_loop:
                if (KIND != 1 && (u64)(hi - lo) + 1 < (u64)s_need && it++ < 400) {
                    lv_limit = (s_emit && max(nbl, nbr) >= s_msl) ? ~0ull : (u64)s_need;
                    lv_ret = 0;
                    goto lv_begin;
                }
                if (KIND != 3 && s_emit && lev >= s_msl) push(s_pivot, s_pivot + lev, lo, (u64)(hi - lo) + 1);
            search_done:
                // straight to the driver continuation of this kind (no extra pass through the state switch)
                if (KIND == 1) { if (d_ret == P_D1_F) goto c_d1_f; if (d_ret == P_D1_L) goto c_d1_l; goto c_d1_r; }
                if (KIND == 2) { if (d_ret == P_D2_F) goto c_d2_f; if (d_ret == P_D2_L) goto c_d2_l; goto c_d2_r; }
                goto c_d3_s;

            // ------------------------------------------------ one level drop (reference L2569-2574)
            lv_begin:
                if (KIND == 1) { state = P_TASK; break; }
                lev = nbl > nbr ? nbl : nbr;
                if (lev <= 0) { lev = 0; lo = 0; hi = n - 1; nbl = nbr = 0; goto lv_done; }
                if (nbl >= lev) {
                    if (lo > base && lo <= base + (idx_t)nvalid) {            // lo-1 lies in the located block
                        int z = (int)(lo - base) - 1;
                        while (z >= 0 && BL(z) >= lev) z--;
                        if (z < 0) { lo = base; nbl = base == 0 ? 0 : -1; }
                        else { lo = base + z + 1; nbl = BL(z); }
                    } else nbl = lo == 0 ? 0 : -1;
                    if (nbl < 0) { e_dir = -1; e_from = lo; e_L = lev; e_limit = lv_limit; e_ret = R_LV_L; goto e_begin; }
                }
            lv_right:
                if (nbr >= lev) {
                    if (hi + 1 >= base && hi + 1 < base + (idx_t)nvalid) {    // hi+1 lies in the located block
                        int z = (int)(hi + 1 - base);
                        while (z < nvalid && BL(z) >= lev) z++;
                        if (z >= nvalid) { hi = base + nvalid - 1; nbr = base + nvalid == n ? 0 : -1; }
                        else { hi = base + z - 1; nbr = BL(z); }
                    } else nbr = hi == n - 1 ? 0 : -1;
                    if (nbr < 0) { e_dir = +1; e_from = hi; e_L = lev; e_limit = lv_limit; e_ret = R_LV_R; goto e_begin; }
                }
            lv_done:
                if (lv_ret == 0) goto level_loop;
                goto d3_levels;

            // ------------------------------------------------ extension: 8-ary gallop, then 8-ary narrowing
            e_begin:
                e_cur = e_from; e_s = 1; e_sh = 0; e_up = true; e_guard = 0;
            e_probe:
                if (++e_guard > 400) { e_edge = e_cur; e_nb = 0; goto e_done; }
                if (e_dir > 0) {
                    u64 room = (u64)(n - 1 - e_cur) >> e_sh;
                    e_cnt = room < 8 ? (int)room : 8;
                    req_block(e_cur + (idx_t)e_s, (long long)e_s, e_cnt, e_L);
                } else {
                    u64 room = (u64)e_cur >> e_sh;
                    e_cnt = room < 8 ? (int)room : 8;
                    req_block(e_cur - (idx_t)e_s, -(long long)e_s, e_cnt, e_L);
                }
                if (STATS) st_ext++;
                state = P_E;
                break;
            case P_E:
                {
                    int c = 0;
                    while (c < e_cnt && RL(c) >= e_L) c++;
                    e_cur = e_dir > 0 ? e_cur + (idx_t)((u64)c * e_s) : e_cur - (idx_t)((u64)c * e_s);
                    if (c == 8) {
                        u64 gone = e_dir > 0 ? (u64)(e_cur - e_from) : (u64)(e_from - e_cur);
                        if (gone >= e_limit) { e_edge = e_cur; e_nb = e_L; goto e_done; }
                        if (e_up && e_s < (1ull << 40)) { e_s *= 8; e_sh += 3; }
                        goto e_probe;
                    }
                    if (e_s == 1) { e_nb = c < e_cnt ? RL(c) : 0; e_edge = e_cur; goto e_done; }
                    e_up = false;
                    e_s >>= 3; e_sh -= 3;
                    goto e_probe;
                }
            e_done:
                switch (e_ret) {
                case R_INIT_L: lo = e_edge; nbl = e_nb; goto init_right;
                case R_INIT_R: hi = e_edge; nbr = e_nb; goto leveled;
                case R_LV_L: lo = e_edge; nbl = e_nb; goto lv_right;
                default: hi = e_edge; nbr = e_nb; goto lv_done;
                }
            }
        next_state:;
        }
        if (!SYNC && done) break;
        if (SYNC) __syncwarp();
        // =========================== memory phase: every lane issues the loads of its own request ===========================
        u64 buf[8];
        const int req = done ? (int)M_NONE : mreq;
        if (STATS && req != M_NONE) st_iter++;
        if (req == M_BLOCK) {
            // asynchronous copies straight into the staging words: a rolled loop with every load in flight
            flat_stage_t dst = FLAT_STAGE_BASE(stg);
            const char *kp = reinterpret_cast<const char *>(ix.keys + m_base);
            const char *pq = reinterpret_cast<const char *>(ix.pos_lo + m_base);
#pragma unroll 1
            for (int j = 0; j < m_count; j++) {
                if (has_keys) {
                    FLAT_CP4(dst, kp);
                    FLAT_CP4(dst + 32, kp + 4);
                }
                FLAT_CP4(dst + 64, pq);
                dst += 4; kp += m_stride * 8; pq += m_stride * 4;
            }
            FLAT_CP_COMMIT();
            if (WIDE) {
                u64 h = 0; idx_t idx = m_base;
#pragma unroll
                for (int j = 0; j < 8; j++) { if (j < m_count) h |= (u64)__ldg(ix.pos_hi + idx) << (8 * j); idx += (idx_t)m_stride; }
                pphi = h;
            }
        } else if (req == M_PAC) {
            const u32 v = stg[16 + pac_e];
            u64 p = (u64)v | (WIDE ? ((pphi >> (8 * pac_e)) & 0xff) << 32 : 0);
            u64 ref_len = ix.sa_num - p;
            int cap = ref_len < (u64)m_cap ? (int)ref_len : m_cap;
            const u64 *wp = ix.pac + (p >> 5) + pac_k0;
#pragma unroll
            for (int j = 0; j < 5; j++) buf[j] = (j == 0 || 32 * (pac_k0 + j - 1) < cap) ? __ldg(wp + j) : 0;
        } else if (req == M_RMI) {
            const ulonglong2 *r2 = reinterpret_cast<const ulonglong2 *>(m_rec);
            ulonglong2 a = __ldg(r2), b = __ldg(r2 + 1);
            buf[0] = a.x; buf[1] = a.y; buf[2] = b.x;
        } else if (req == M_LD) {
            if (m_ldk == LD_POS) {
                buf[0] = __ldg(ix.pos_lo + m_base);
                if (WIDE) buf[0] |= (u64)__ldg(ix.pos_hi + m_base) << 32;
            } else {
                buf[0] = __ldg(ix.isa_lo + m_base);
                if (WIDE) buf[0] |= (u64)__ldg(ix.isa_hi + m_base) << 32;
            }
        } else if (req == M_T2REC) {
            const ulonglong2 *t2 = reinterpret_cast<const ulonglong2 *>(A.t2_in + task);
            ulonglong2 a = __ldg(t2), b = __ldg(t2 + 1);
            buf[0] = a.x; buf[1] = a.y; buf[2] = b.x;
        } else if (req == M_TASK) {
            load_task_words(A, rid, sm);
        }
        if (SYNC) __syncwarp();
        // =========================== compare phase ===========================
        if (req == M_BLOCK) {
            u32 less = 0, und = 0;
            const u32 qlo = (u32)q0, qhi = (u32)(q0 >> 32);
            u32 *const rl = stg + rl_off;
            FLAT_CP_WAIT();
#pragma unroll BU
            for (int j = 0; j < m_count; j++) {
                const u32 klo = stg[j], khi = stg[8 + j], pj = stg[16 + j];
                int cap;
                if (WIDE) {
                    u64 p = (u64)pj | (((pphi >> (8 * j)) & 0xff) << 32);
                    u64 ref_len = ix.sa_num - p;
                    cap = ref_len < (u64)m_cap ? (int)ref_len : m_cap;
                } else {
                    u32 ref_len = (u32)n - pj;
                    cap = (int)min(ref_len, (u32)m_cap);
                }
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
                rl[j] = (u32)l;
            }
            r_less = less; undet = und;
            if (und) { pac_e = __ffs(und) - 1; pac_k0 = has_keys ? 1 : 0; mreq = M_PAC; }
            else mreq = M_NONE;
        } else if (req == M_PAC) {
            if (STATS) st_pac++;
            const u32 v = stg[16 + pac_e];
            u64 p = (u64)v | (WIDE ? ((pphi >> (8 * pac_e)) & 0xff) << 32 : 0);
            u64 ref_len = ix.sa_num - p;
            const int cap = ref_len < (u64)m_cap ? (int)ref_len : m_cap;
            const int sh = (int)(p & 31) * 2;
            int res = -1; bool less = false;
            u64 w0 = buf[0], w1 = buf[1], w2 = buf[2], w3 = buf[3], w4 = buf[4];
#pragma unroll 1
            for (int j = 0; j < 4 && res < 0; j++) {
                int k = pac_k0 + j;
                if (32 * k >= cap) res = cap;
                else {
                    u64 rr = sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0;
                    u64 qq = k == 0 ? q0 : qword(k);
                    u64 x = rr ^ qq;
                    if (x) {
                        int b = (__ffsll((long long)x) - 1) >> 1;
                        int l = 32 * k + b;
                        if (l >= cap) res = cap;
                        else { res = l; less = ((rr >> (2 * b)) & 3) < ((qq >> (2 * b)) & 3); }
                    }
                }
                w0 = w1; w1 = w2; w2 = w3; w3 = w4;
            }
            if (res < 0) pac_k0 += 4;                      // four more words of the same entry
            else {
                stg[rl_off + pac_e] = (u32)res;
                if (less) r_less |= 1u << pac_e;
                undet &= ~(1u << pac_e);
                if (undet) { pac_e = __ffs(undet) - 1; pac_k0 = has_keys ? 1 : 0; }
                else mreq = M_NONE;
            }
        } else if (req == M_RMI) {
            rmi_a = __longlong_as_double((long long)buf[0]);
            rmi_b = __longlong_as_double((long long)buf[1]);
            rmi_e = buf[2];
            mreq = M_NONE;
        } else if (req == M_LD) {
            ld_val = buf[0];
            mreq = M_NONE;
        } else if (req == M_T2REC) {
            rid = (u32)buf[0];
            u32 se = (u32)(buf[0] >> 32), cnt = (u32)buf[1], full = (u32)(buf[1] >> 32);
            int qb = se & 0xffff, qe = se >> 16;
            pivot = (qb + qe) >> 1;
            min_intv = (int)cnt + 1;
            full_match = full != 0;
            max_refpos = full ? (idx_t)buf[2] : 0;
            use_cached = !full && ix.mode == 3 && cnt >= 1 && cnt < 10;
            cache_pivot = qb; cache_pivot_end = qe; cache_refpos = (idx_t)buf[2];
            mreq = rid == 0xffffffffu ? M_NONE : M_TASK;
        } else if (req == M_TASK) {
            mreq = M_NONE;
        }
    }
    finish_thread<KIND>(A, out, st_search, st_block, st_kary, st_ext, st_pac, st_iter);
#undef RL
#undef BL
}

}  // namespace tpr
