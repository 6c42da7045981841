// meme_seed_tpr.cu -- kernel A, thread-per-task generation: a warp-synchronous state machine.
//
// Why (measured, profiles/r1): seeding is a chain of ~190 dependent DRAM round trips per read.  With one
// warp (or an 8-lane group) per read an SM keeps only 24-96 chains in flight and sits at ~4 % of DRAM
// bandwidth, and divergent sub-warp groups do not overlap their stalls.  Here every LANE owns a task and
// the kernel body is one loop whose memory phase all 32 lanes execute together:
//
//     for (;;) {  control (divergent, ALU only): run the task's coroutines until they need memory
//                 memory  (convergent)         : each lane issues the loads of ITS OWN request
//                 compare (convergent)         : 2-bit compares; results go back to the coroutines   }
//
// so a warp keeps 32 independent requests in flight.  A request is one of
//     BLOCK  8 suffix-array entries: inline 64-bit keys + positions (contiguous, or strided for 8-ary steps)
//     PAC    5 consecutive 64-bit words of the 2-bit reference, for one entry whose 32-base key matched
//     RMI    one 32-byte P-RMI leaf / sub-model record (fp64 fma = learned_index_lookup, reference L186-210)
//     LD     one inverse-SA entry or one SA position
//     TASK   the task's packed read (2-bit words + N mask) into shared memory
// The search logic is that of the cooperative kernel (meme_seed.cu) with "lane j of the group" replaced by
// "entry j of my block".  Results are canonical (they depend on the suffix array only), hence identical.
//
// The per-read program of the reference is cut into three task kinds that share the search engine:
//     T1  zig-zag SMEM pass of one read      Learned_getSMEMsAllPos/OnePos_step1     (reference L913-972, L1691-1894)
//     T2  reseeding of one long, rare SMEM   Learned_getSMEMsOnePosOneThread         (L1897-2126)
//     T3  third round of one read            Learned_bwtSeedStrategyAllPosOneThread* (L974-1466)
// T2 tasks are emitted by T1 as it finds qualifying SMEMs; T3 only takes T1's "whole read matched" flag as a
// start hint.  Tasks append raw records {read,start,end | SA lo,count} to a pool that is radix-sorted by
// (read, start, end) = the reference's ks_introsort(mem_smem_sort_lt_learned) (src/bwamem.cpp:L1397).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "meme_common.cuh"

#include "meme_seed_task.cuh"

namespace tpr {

#define CO_AWAIT(pcvar, label) do { pcvar = label; return false; case label:; } while (0)

struct Thread {
    const IndexView &ix;
    const Args &A;
    u32 *sm;
    // task
    u32 rid; int L;
    // memory request
    int mreq; u64 m_base; long long m_stride; int m_count, m_cap, m_ldk; const RmiRec *m_rec;
    // probe results
    u64 r_lo, r_hi; u32 r_less, undet; int r_count;
    u32 pp[8]; u64 pphi; int pac_e, pac_k0;
    u64 ld_val; double rmi_a, rmi_b; u64 rmi_e;
    // query
    int q_strand, q_from, q_V; u64 q0;
    // search coroutine
    int spc; bool s_active, s_emit, s_stepwise, s_isa;
    int s_pivot, s_need, s_msl, s_mode; u64 s_hint;
    u64 center, rad_lo, rad_hi, lo_b, hi_b, k_stride, k_pa; int it; bool k_contig;
    u64 base; int nvalid; u64 b_lo, b_hi;              // the located block: 8 x u16 lcps
    u64 lo, hi; int nbl, nbr, lev;
    // level coroutine
    int lvpc; u64 lv_limit;
    // extend coroutine
    int epc, e_dir, e_L, e_cnt, e_nb, e_guard; u64 e_from, e_cur, e_s, e_limit, e_edge; bool e_up;
    // driver
    int dpc, pivot, sp, next;
    bool full_match, use_cached; u64 max_refpos, cache_refpos; int cache_pivot, cache_pivot_end, min_intv, min_seed_len;
    u64 r3_last_lo, r3_last_cnt;
    // output
    u64 pslot, pend, tslot, tend; u32 n_emit; u64 n_hit;
    u32 st_search, st_block, st_kary, st_ext, st_pac, st_iter;

    __device__ Thread(const IndexView &ix_, const Args &A_, u32 *sm_) : ix(ix_), A(A_), sm(sm_) {}

    // ---------------------------------------------------------------- packed read
    __device__ __forceinline__ u64 rdw(int w) const
    {
        if (w < 0 || w >= TPR_RW) return 0;
        return (u64)sm[2 * w] | ((u64)sm[2 * w + 1] << 32);
    }
    __device__ __forceinline__ u64 fwwin(int start) const
    {   // 32 forward bases from `start` (zeros outside the read)
        if (start < 0) return start <= -32 ? 0 : rdw(0) << (2 * (-start));
        int w = start >> 5, sh = (start & 31) * 2;
        u64 a = rdw(w);
        if (sh == 0) return a;
        return (a >> sh) | (rdw(w + 1) << (64 - sh));
    }
    __device__ __forceinline__ u64 qword(int k) const
    {   // 32 query bases from q_from + 32k; the reverse strand is derived from the forward words:
        // rc[f..f+32) = revcomp(fw[L-1-f-31 .. L-1-f])
        int f = q_from + 32 * k;
        if (q_strand == 0) return fwwin(f);
        return revcomp32(fwwin(L - 32 - f));
    }
    __device__ __forceinline__ bool is_n(int p) const { return (sm[20 + (p >> 5)] >> (p & 31)) & 1; }
    __device__ int next_n(int from) const
    {
        int w = from >> 5;
        u32 m = sm[20 + w] >> (from & 31);
        if (m) { int p = from + __ffs(m) - 1; return p < L ? p : L; }
        for (w++; w * 32 < L; w++) {
            u32 x = sm[20 + w];
            if (x) { int p = w * 32 + __ffs(x) - 1; return p < L ? p : L; }
        }
        return L;
    }
    __device__ int prev_n(int p) const
    {
        int w = p >> 5;
        u32 m = sm[20 + w] & (0xffffffffu >> (31 - (p & 31)));
        if (m) return w * 32 + 31 - __clz(m);
        for (w--; w >= 0; w--) {
            u32 x = sm[20 + w];
            if (x) return w * 32 + 31 - __clz(x);
        }
        return -1;
    }

    // ---------------------------------------------------------------- results
    __device__ __forceinline__ static int get16(u64 a, u64 b, int j) { return (int)(((j < 4 ? a : b) >> (16 * (j & 3))) & 0xffff); }
    __device__ __forceinline__ int rl(int j) const { return get16(r_lo, r_hi, j); }      // last probe
    __device__ __forceinline__ int bl(int j) const { return get16(b_lo, b_hi, j); }      // located block
    __device__ __forceinline__ void set_rl(int j, int v)
    {
        u64 m = 0xffffull << (16 * (j & 3)), x = (u64)(v & 0xffff) << (16 * (j & 3));
        if (j < 4) r_lo = (r_lo & ~m) | x; else r_hi = (r_hi & ~m) | x;
    }
    __device__ __forceinline__ u64 get_pos(int j) const
    {
        u32 v = pp[0];
#pragma unroll
        for (int k = 1; k < 8; k++) v = j == k ? pp[k] : v;
        return (u64)v | (((pphi >> (8 * j)) & 0xff) << 32);
    }
    __device__ __forceinline__ int lead_less() const
    {
        int c = __ffs(~r_less) - 1;
        return c > r_count ? r_count : c;
    }
    __device__ __forceinline__ void req_block(u64 b, long long stride, int count, int cap)
    {
        m_base = b; m_stride = stride; m_count = count; m_cap = cap; r_count = count;
        if (count > 0) mreq = M_BLOCK; else { mreq = M_NONE; r_lo = r_hi = 0; r_less = 0; }
    }
    __device__ __forceinline__ void req_contig(u64 b, int cap)
    {
        u64 n = ix.sa_num;
        if (n >= 8 && b > n - 8) b = n - 8;
        int c = n - b < 8 ? (int)(n - b) : 8;
        base = b; nvalid = c;
        req_block(b, 1, c, cap);
        st_block++;
    }
    __device__ __forceinline__ void req_ld(int kind, u64 idx) { mreq = M_LD; m_ldk = kind; m_base = idx; }

    // ---------------------------------------------------------------- output
    __device__ void push(int start, int end, u64 lo_, u64 cnt)
    {
        if (pslot == pend) { pslot = atomicAdd((unsigned long long *)&A.ctrl[C_POOL], 16ull); pend = pslot + 16; }
        u64 k = pslot++;
        u32 c32 = cnt > 0xffffffffull ? 0xffffffffu : (u32)cnt;
        if (k < A.pool_cap) {
            A.pool_keys[k] = ((u64)rid << 32) | ((u64)start << 16) | (u64)end;
            PoolRec v; v.lo = lo_; v.cnt = c32; v.pad = 0;
            A.pool_vals[k] = v;
        } else atomicExch((unsigned long long *)&A.ctrl[C_POOL_OVF], 1ull);
        n_emit++;
        u64 c = c32;
        if (A.opt.max_occ > 0) {
            u64 step = c > (u64)A.opt.max_occ ? c / A.opt.max_occ : 1;
            u64 t = (c + step - 1) / step;
            c = t < (u64)A.opt.max_occ ? t : (u64)A.opt.max_occ;
        }
        n_hit += c;
        // T1: long and rare SMEMs become reseed tasks (reference L923-947)
        if (A.kind == 1 && A.opt.rounds >= 2 && end - start >= A.opt.split_len && c32 <= (u32)A.opt.split_width) {
            if (tslot == tend) { tslot = atomicAdd((unsigned long long *)&A.ctrl[C_T2], 8ull); tend = tslot + 8; }
            u64 t = tslot++;
            if (t < A.t2_cap) {
                Task2 x; x.read = rid; x.se = (u32)start | ((u32)end << 16); x.cnt = c32;
                x.full = full_match ? 1u : 0u; x.lo = full_match ? max_refpos : lo_; x.pad = 0;
                A.t2_out[t] = x;
            } else atomicExch((unsigned long long *)&A.ctrl[C_T2_OVF], 1ull);
        }
    }
    __device__ void flush_slots()
    {   // reserved but unused slots must not look like records
        for (; pslot < pend; pslot++) if (pslot < A.pool_cap) A.pool_keys[pslot] = ~0ull;
        for (; tslot < tend; tslot++)
            if (tslot < A.t2_cap) { Task2 x; x.read = 0xffffffffu; x.se = 0; x.cnt = 0; x.full = 0; x.lo = 0; x.pad = 0; A.t2_out[tslot] = x; }
    }

    // ---------------------------------------------------------------- extend coroutine
    // furthest index in direction e_dir reachable from e_from through entries with lcp >= e_L (8-ary gallop,
    // then 8-ary narrowing); e_nb = lcp of the first entry beyond (0 at the array ends).  Stops early once
    // e_limit entries beyond e_from qualify.
    __device__ __forceinline__ void E_start(int dir, u64 from, int Lreq, u64 limit)
    {
        epc = 0; e_dir = dir; e_from = from; e_L = Lreq; e_limit = limit;
    }
    __device__ __forceinline__ bool E_step()
    {
        const u64 n = ix.sa_num;
        switch (epc) {
        case 0:
            e_cur = e_from; e_s = 1; e_up = true; e_guard = 0;
            for (;;) {
                if (++e_guard > 400) { e_edge = e_cur; e_nb = 0; return true; }
                if (e_dir > 0) {
                    u64 room = (n - 1 - e_cur) / e_s;
                    e_cnt = room < 8 ? (int)room : 8;
                    req_block(e_cur + e_s, (long long)e_s, e_cnt, e_L);
                } else {
                    u64 room = e_cur / e_s;
                    e_cnt = room < 8 ? (int)room : 8;
                    req_block(e_cur - e_s, -(long long)e_s, e_cnt, e_L);
                }
                st_ext++;
                CO_AWAIT(epc, 1);
                {
                    int c = 0;
                    while (c < e_cnt && rl(c) >= e_L) c++;
                    e_cur = e_dir > 0 ? e_cur + (u64)c * e_s : e_cur - (u64)c * e_s;
                    if (c == 8) {
                        u64 gone = e_dir > 0 ? e_cur - e_from : e_from - e_cur;
                        if (gone >= e_limit) { e_edge = e_cur; e_nb = e_L; return true; }
                        if (e_up && e_s < (1ull << 40)) e_s *= 8;
                        continue;
                    }
                    if (e_s == 1) { e_nb = c < e_cnt ? rl(c) : 0; e_edge = e_cur; return true; }
                    e_up = false;
                    e_s /= 8;
                }
            }
        }
        return true;
    }

    // ---------------------------------------------------------------- one level drop (reference L2569-2574)
    __device__ __forceinline__ bool LV_step()
    {
        const u64 n = ix.sa_num;
        switch (lvpc) {
        case 0:
            lev = nbl > nbr ? nbl : nbr;
            if (lev <= 0) { lev = 0; lo = 0; hi = n - 1; nbl = nbr = 0; return true; }
            if (nbl >= lev) {
                if (lo > base && lo <= base + (u64)nvalid) {              // lo-1 lies in the located block
                    int z = (int)(lo - base) - 1;
                    while (z >= 0 && bl(z) >= lev) z--;
                    if (z < 0) { lo = base; nbl = base == 0 ? 0 : -1; }
                    else { lo = base + z + 1; nbl = bl(z); }
                } else nbl = lo == 0 ? 0 : -1;
                if (nbl < 0) {
                    E_start(-1, lo, lev, lv_limit);
                    while (!E_step()) CO_AWAIT(lvpc, 1);
                    lo = e_edge; nbl = e_nb;
                }
            }
            if (nbr >= lev) {
                if (hi + 1 >= base && hi + 1 < base + (u64)nvalid) {      // hi+1 lies in the located block
                    int z = (int)(hi + 1 - base);
                    while (z < nvalid && bl(z) >= lev) z++;
                    if (z >= nvalid) { hi = base + nvalid - 1; nbr = base + nvalid == n ? 0 : -1; }
                    else { hi = base + z - 1; nbr = bl(z); }
                } else nbr = hi == n - 1 ? 0 : -1;
                if (nbr < 0) {
                    E_start(+1, hi, lev, lv_limit);
                    while (!E_step()) CO_AWAIT(lvpc, 2);
                    hi = e_edge; nbr = e_nb;
                }
            }
        }
        return true;
    }

    // ---------------------------------------------------------------- search coroutine
    // right_smem_search / mem_search / *_tradeoff: the level of the query with >= s_need occurrences;
    // emits the SMEM when asked and long enough.  s_stepwise: stop after the first level (third round).
    __device__ __forceinline__ void start_search(int pv, bool right, bool emit, int need, int msl, int mode, u64 hint, bool isa)
    {
        s_pivot = pv; s_emit = emit; s_need = need; s_msl = msl; s_mode = mode; s_hint = hint; s_isa = isa; s_stepwise = false;
        q_strand = right ? 0 : 1;
        q_from = right ? pv : L - 1 - pv;
        q_V = right ? next_n(pv) - pv : pv - prev_n(pv);
        spc = 0; s_active = true;
    }
    __device__ void start_search_auto(int pv, bool right, bool emit)
    {   // start selection of reference L1741-1816 / L1926-2040
        if (ix.mode == 3) {
            if (full_match) {
                u64 rp = right ? max_refpos + pv : ix.sa_num - max_refpos - pv - 1;
                start_search(pv, right, emit, min_intv, min_seed_len, ST_EXACT, rp, true);
                return;
            }
            if (use_cached) {
                if (right && cache_pivot_end >= pv + 18 && pv >= cache_pivot) {
                    start_search(pv, true, emit, min_intv, min_seed_len, ST_NEAR, cache_refpos + pv - cache_pivot, true);
                    return;
                }
                if (!right && pv < cache_pivot_end && cache_pivot + 18 <= pv) {
                    start_search(pv, false, emit, min_intv, min_seed_len, ST_NEAR,
                                 ix.sa_num + cache_pivot - cache_refpos - pv - 1, true);
                    return;
                }
            }
        }
        start_search(pv, right, emit, min_intv, min_seed_len, ST_RMI, 0, false);
    }

    __device__ __forceinline__ bool S_step()
    {
        const u64 n = ix.sa_num;
        switch (spc) {
        case 0:
            st_search++;
            q0 = qword(0);
            if (s_isa) {                                  // the hint is a reference position: inverse SA first
                req_ld(LD_ISA, s_hint);
                CO_AWAIT(spc, 1);
                s_hint = ld_val;
            }
            if (s_mode == ST_EXACT) {
                req_contig(s_hint > 3 ? s_hint - 3 : 0, q_V);
                CO_AWAIT(spc, 2);
                goto located;
            }
            if (s_mode == ST_RMI) {
                {
                    u64 k = q0;                           // RMI key = Tokenization (L795-901): first base in the MSBs, pad 3
                    if (q_V < 32) k |= ~0ull << (2 * q_V);
                    u64 y = __brevll(k);
                    ld_val = ((y & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((y & 0x5555555555555555ull) << 1);
                    mreq = M_RMI; m_rec = ix.l2 + (ld_val >> ix.bit_shift);
                }
                CO_AWAIT(spc, 3);
                if (rmi_e >> 63) {
                    {
                        double p = fma(rmi_b, __ull2double_rn(ld_val), rmi_a);
                        u64 st = (rmi_e >> 32) & 0x7fffffff;
                        double bound = (double)(rmi_e & 0xffffffffull) - 1;
                        u64 sub = p < 0.0 ? 0 : (p > bound ? (u64)bound : (u64)p);
                        mreq = M_RMI; m_rec = ix.l1 + st + sub;
                    }
                    CO_AWAIT(spc, 4);
                }
                {
                    double p = fma(rmi_b, __ull2double_rn(ld_val), rmi_a);
                    double bound = (double)n - 1.0;
                    center = p < 0.0 ? 0 : (p > bound ? (u64)bound : (u64)p);
                    rad_lo = (rmi_e >> 32) & 0x3fffffff;
                    rad_hi = rmi_e & 0x7fffffff;
                }
            } else { center = s_hint; rad_lo = rad_hi = 2; }
            lo_b = 0; hi_b = n;
            // prediction errors concentrate near zero: try the block centred on the prediction first
            if (rad_lo + rad_hi + 2 > 8 && rad_lo + rad_hi <= 40) {
                req_contig(center > 4 ? center - 4 : 0, q_V);
                CO_AWAIT(spc, 5);
                {
                    int c = lead_less();
                    bool left_ok = c > 0 || base == 0, right_ok = c < nvalid || base + nvalid == n;
                    if (left_ok && right_ok) goto located;
                    if (!left_ok) hi_b = base; else lo_b = base + nvalid;
                }
            }
            for (it = 0; it < 64; it++) {
                {
                    u64 pa = center > rad_lo ? center - rad_lo : 0;
                    pa = pa < lo_b ? lo_b : (pa > hi_b ? hi_b : pa);
                    u64 pb = center + rad_hi + 1;
                    pb = pb < lo_b ? lo_b : (pb > hi_b ? hi_b : pb);
                    u64 span = pb - pa;
                    k_contig = span <= 6;
                    if (k_contig) {
                        u64 slack = (6 - span) / 2;
                        req_contig(pa > 1 + slack ? pa - 1 - slack : 0, q_V);
                    } else {
                        k_stride = (span + 7) / 8; k_pa = pa;
                        u64 first = pa + k_stride - 1;
                        int cnt = 0;
                        if (first < pb) { u64 c = (pb - 1 - first) / k_stride + 1; cnt = c < 8 ? (int)c : 8; }
                        req_block(first, (long long)k_stride, cnt, q_V);
                        st_kary++;
                    }
                }
                CO_AWAIT(spc, 6);
                if (k_contig) {
                    int c = lead_less();
                    bool left_ok = c > 0 || base == 0, right_ok = c < nvalid || base + nvalid == n;
                    if (left_ok && right_ok) goto located;
                    if (!left_ok) hi_b = base; else lo_b = base + nvalid;
                } else {
                    int c = lead_less();
                    if (c > 0) { u64 nl = k_pa + (u64)c * k_stride; if (nl > lo_b) lo_b = nl; }
                    if (c < r_count) { u64 nh = k_pa + (u64)(c + 1) * k_stride - 1; if (nh < hi_b) hi_b = nh; }
                }
                rad_lo = rad_lo * 8 + 8;
                rad_hi = rad_hi * 8 + 8;
            }
        located:
            // ---- first level = maximal lcp in the block, [lo,hi] = its interval ----
            b_lo = r_lo; b_hi = r_hi;
            {
                int m = -1, first = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) { int l = j < nvalid ? bl(j) : -1; if (l > m) { m = l; first = j; } }
                lev = m;
                int last = first;
                while (last + 1 < nvalid && bl(last + 1) >= lev) last++;
                lo = base + first; hi = base + last;
                const bool exact = s_emit && lev >= s_msl;
                if (!exact && hi - lo + 1 >= (u64)s_need) { nbl = nbr = lev; goto leveled; }
                lv_limit = exact ? ~0ull : (u64)s_need;
                nbl = first > 0 ? bl(first - 1) : (base == 0 ? 0 : -1);
                nbr = last < nvalid - 1 ? bl(last + 1) : (base + nvalid == n ? 0 : -1);
            }
            if (nbl < 0) {
                E_start(-1, lo, lev, lv_limit);
                while (!E_step()) CO_AWAIT(spc, 7);
                lo = e_edge; nbl = e_nb;
            }
            if (nbr < 0) {
                E_start(+1, hi, lev, lv_limit);
                while (!E_step()) CO_AWAIT(spc, 8);
                hi = e_edge; nbr = e_nb;
            }
        leveled:
            if (s_stepwise) return true;
            for (it = 0; hi - lo + 1 < (u64)s_need && it < 400; it++) {
                lv_limit = (s_emit && max(nbl, nbr) >= s_msl) ? ~0ull : (u64)s_need;
                lvpc = 0;
                while (!LV_step()) CO_AWAIT(spc, 9);
            }
            if (s_emit && lev >= s_msl) push(s_pivot, s_pivot + lev, lo, hi - lo + 1);
        }
        return true;
    }

    // ---------------------------------------------------------------- drivers
    // T1: zig-zag pass (reference L913-972 outer loop + L1691-1894)
    __device__ __forceinline__ bool D1_step()
    {
        switch (dpc) {
        case 0:
            pivot = 0;
            while (pivot < L) {
                if (is_n(pivot)) { pivot = (L - pivot < min_seed_len) ? L : pivot + 1; continue; }
                if (pivot != 0 && !is_n(pivot - 1)) {
                    sp = pivot;
                    while (sp < L) {
                        if (is_n(sp)) {
                            if (L - sp < min_seed_len) { pivot = L; sp = L; }
                            else { sp += 1; pivot += 1; }
                            continue;
                        }
                        start_search_auto(pivot, false, false);
                        CO_AWAIT(dpc, 1);
                        pivot = pivot - lev + 1;
                        if (L - pivot < min_seed_len) break;
                        start_search_auto(pivot, true, true);
                        CO_AWAIT(dpc, 2);
                        sp = pivot + lev;
                        pivot = sp;
                    }
                    pivot = L;
                } else {
                    start_search(pivot, true, true, 1, min_seed_len, ST_RMI, 0, false);
                    CO_AWAIT(dpc, 3);
                    if (ix.mode == 3 && lev == L) {
                        // whole read matched (L1870-1875): later searches start from the inverse SA
                        req_ld(LD_POS, lo);
                        CO_AWAIT(dpc, 4);
                        full_match = true; max_refpos = ld_val;
                    }
                    pivot += lev;
                }
            }
            if (A.rd_full) { A.rd_full[rid] = full_match ? 1 : 0; A.rd_maxref[rid] = max_refpos; }
        }
        return true;
    }

    // T2: reseeding of one SMEM (reference L1897-2126); the task fetch preloaded pivot/min_intv/cache fields
    __device__ __forceinline__ bool D2_step()
    {
        switch (dpc) {
        case 0:
            if (use_cached) {                           // cache_refpos = SA[lo] (L934-939)
                req_ld(LD_POS, cache_refpos);
                CO_AWAIT(dpc, 1);
                cache_refpos = ld_val;
            }
            start_search_auto(pivot, true, false);
            CO_AWAIT(dpc, 2);
            next = pivot + lev; sp = pivot;
            while (sp < next) {
                start_search_auto(pivot, false, false);
                CO_AWAIT(dpc, 3);
                pivot = pivot - lev + 1;
                if (next - pivot < min_seed_len) break;
                start_search_auto(pivot, true, true);
                CO_AWAIT(dpc, 4);
                sp = pivot + lev;
                pivot = sp;
            }
        }
        return true;
    }

    // T3: third round (reference L974-1466)
    __device__ __forceinline__ bool D3_step()
    {
        const int msl = A.opt.min_seed_len + 1, mi = A.opt.max_mem_intv;
        switch (dpc) {
        case 0:
            pivot = 0;
            while (pivot < L - msl + 1) {
                if (is_n(pivot)) { pivot++; continue; }
                {
                    int V = next_n(pivot) - pivot;
                    if (V < msl) { pivot += V; continue; }
                }
                if (ix.mode == 3 && full_match) start_search(pivot, true, false, mi, msl, ST_EXACT, max_refpos + pivot, true);
                else start_search(pivot, true, false, mi, msl, ST_RMI, 0, false);
                s_stepwise = true;
                CO_AWAIT(dpc, 1);
                if (lev < msl) { pivot += msl; continue; }
                r3_last_lo = lo; r3_last_cnt = 0;
                for (it = 0; it < 400; it++) {
                    {
                        u64 cnt = hi - lo + 1;
                        if (cnt >= (u64)mi) {                                     // L1243-1251
                            if (r3_last_cnt > 0) push(pivot, pivot + max(lev + 1, msl), r3_last_lo, r3_last_cnt);
                            pivot += lev + 1;
                            break;
                        }
                        if (max(nbl, nbr) < msl) {                                // L1252-1258
                            push(pivot, pivot + msl, lo, cnt);
                            pivot += msl;
                            break;
                        }
                        r3_last_lo = lo; r3_last_cnt = cnt;
                    }
                    lv_limit = (u64)mi;
                    lvpc = 0;
                    while (!LV_step()) CO_AWAIT(dpc, 2);
                }
            }
        }
        return true;
    }
};
#undef CO_AWAIT

// ---------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(TPR_THREADS, 2) k_tpr(IndexView ix, Args A)
{
    __shared__ u32 smem[TPR_THREADS * TPR_SW];
    Thread T(ix, A, &smem[threadIdx.x * TPR_SW]);
    T.mreq = M_NONE; T.s_active = false;
    T.pslot = T.pend = T.tslot = T.tend = 0;
    T.st_search = T.st_block = T.st_kary = T.st_ext = T.st_pac = T.st_iter = 0;
    T.r_lo = T.r_hi = 0; T.r_less = 0; T.undet = 0; T.r_count = 0; T.pphi = 0;
    T.full_match = false; T.use_cached = false; T.max_refpos = 0; T.cache_refpos = 0;
    T.rid = 0; T.L = 0; T.n_emit = 0; T.n_hit = 0;
    int tstate = 0;                                  // 0: need a task, 1: record/read loading, 2: running
    u64 task = 0;
    const u64 n = ix.sa_num;
    bool alive = true;
    while (alive) {
        // ================= control: run coroutines until a memory request is pending =================
        while (T.mreq == M_NONE) {
            if (tstate == 0) {
                task = atomicAdd((unsigned long long *)&A.ctrl[C_WORK], 1ull);
                if (task >= (u64)A.n_tasks) { alive = false; break; }
                if (KIND == 2) { T.mreq = M_T2REC; tstate = 1; break; }
                T.rid = (u32)task;
                T.mreq = M_TASK; tstate = 1;
                break;
            }
            if (tstate == 1) {
                // read is in shared memory: set the task up
                if (KIND == 2 && T.rid == 0xffffffffu) { tstate = 0; continue; }     // unused task slot
                T.L = (int)(A.read_off[T.rid + 1] - A.read_off[T.rid]);
                T.n_emit = 0; T.n_hit = 0;
                T.min_seed_len = A.opt.min_seed_len;
                T.dpc = 0; T.s_active = false;
                if (KIND == 1) { T.full_match = false; T.use_cached = false; T.max_refpos = 0; T.min_intv = 1; }
                if (KIND == 3) {
                    T.use_cached = false; T.min_intv = A.opt.max_mem_intv; T.min_seed_len = A.opt.min_seed_len + 1;
                    T.full_match = ix.mode == 3 && A.rd_full[T.rid] != 0;
                    T.max_refpos = T.full_match ? A.rd_maxref[T.rid] : 0;
                }
                tstate = 2;
                if (T.L <= 0) { tstate = 0; continue; }
            }
            bool done;
            if (T.s_active) {
                done = false;
                if (T.S_step()) T.s_active = false;
            } else {
                done = KIND == 1 ? T.D1_step() : (KIND == 2 ? T.D2_step() : T.D3_step());
            }
            if (done) {
                if (T.n_emit) {
                    atomicAdd(&A.rd_nsmem[T.rid], T.n_emit);
                    atomicAdd((unsigned long long *)&A.rd_nhit[T.rid], (unsigned long long)T.n_hit);
                }
                tstate = 0;
            }
        }
        if (!alive) break;
        T.st_iter++;
        // ================= memory phase: every lane issues the loads of its own request =================
        u64 buf[8];
        const int req = T.mreq;
        if (req == M_BLOCK) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (j < T.m_count) {
                    u64 idx = T.m_base + (u64)((long long)j * T.m_stride);
                    buf[j] = ix.keys ? __ldg(ix.keys + idx) : 0;
                    T.pp[j] = __ldg(ix.pos_lo + idx);
                }
            }
            if (ix.pos_wide) {
                u64 h = 0;
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (j < T.m_count) h |= (u64)__ldg(ix.pos_hi + T.m_base + (u64)((long long)j * T.m_stride)) << (8 * j);
                T.pphi = h;
            }
        } else if (req == M_PAC) {
            u64 p = T.get_pos(T.pac_e);
            u64 ref_len = n - p;
            int cap = ref_len < (u64)T.m_cap ? (int)ref_len : T.m_cap;
            const u64 *wp = ix.pac + (p >> 5) + T.pac_k0;
#pragma unroll
            for (int j = 0; j < 5; j++) buf[j] = (j == 0 || 32 * (T.pac_k0 + j - 1) < cap) ? __ldg(wp + j) : 0;
        } else if (req == M_RMI) {
            const ulonglong2 *r2 = reinterpret_cast<const ulonglong2 *>(T.m_rec);
            ulonglong2 a = __ldg(r2), b = __ldg(r2 + 1);
            buf[0] = a.x; buf[1] = a.y; buf[2] = b.x;
        } else if (req == M_LD) {
            if (T.m_ldk == LD_POS) {
                buf[0] = __ldg(ix.pos_lo + T.m_base);
                if (ix.pos_wide) buf[0] |= (u64)__ldg(ix.pos_hi + T.m_base) << 32;
            } else {
                buf[0] = __ldg(ix.isa_lo + T.m_base);
                if (ix.pos_wide) buf[0] |= (u64)__ldg(ix.isa_hi + T.m_base) << 32;
            }
        } else if (req == M_T2REC) {
            const ulonglong2 *t2 = reinterpret_cast<const ulonglong2 *>(A.t2_in + task);
            ulonglong2 a = __ldg(t2), b = __ldg(t2 + 1);
            buf[0] = a.x; buf[1] = a.y; buf[2] = b.x;
        } else if (req == M_TASK) {
            const u64 *src = A.packed + (u64)T.rid * TPR_RW;
            const u32 *nsrc = A.nmask + (u64)T.rid * TPR_RW;
#pragma unroll
            for (int w = 0; w < TPR_RW; w++) {
                u64 v = __ldg(src + w);
                T.sm[2 * w] = (u32)v; T.sm[2 * w + 1] = (u32)(v >> 32);
                T.sm[20 + w] = __ldg(nsrc + w);
            }
        }
        // ================= compare phase =================
        if (req == M_BLOCK) {
            u64 lo16 = 0, hi16 = 0; u32 less = 0, und = 0;
            const u64 q = T.q0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (j < T.m_count) {
                    u64 p = (u64)T.pp[j] | (((T.pphi >> (8 * j)) & 0xff) << 32);
                    u64 ref_len = n - p;
                    int cap = ref_len < (u64)T.m_cap ? (int)ref_len : T.m_cap;
                    int l;
                    if (!ix.keys) { und |= 1u << j; l = 0; }
                    else {
                        u64 x = buf[j] ^ q;
                        if (x) {
                            int b = (__ffsll((long long)x) - 1) >> 1;
                            if (b >= cap) l = cap;
                            else { l = b; if (((buf[j] >> (2 * b)) & 3) < ((q >> (2 * b)) & 3)) less |= 1u << j; }
                        } else if (32 >= cap) l = cap;
                        else { und |= 1u << j; l = 32; }
                    }
                    if (j < 4) lo16 |= (u64)l << (16 * j); else hi16 |= (u64)l << (16 * (j - 4));
                }
            }
            T.r_lo = lo16; T.r_hi = hi16; T.r_less = less; T.undet = und;
            if (und) { T.pac_e = __ffs(und) - 1; T.pac_k0 = ix.keys ? 1 : 0; T.mreq = M_PAC; }
            else T.mreq = M_NONE;
        } else if (req == M_PAC) {
            T.st_pac++;
            const int e = T.pac_e;
            u64 p = T.get_pos(e);
            u64 ref_len = n - p;
            const int cap = ref_len < (u64)T.m_cap ? (int)ref_len : T.m_cap;
            const int sh = (int)(p & 31) * 2;
            int res = -1; bool less = false;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (res < 0) {
                    int k = T.pac_k0 + j;
                    if (32 * k >= cap) res = cap;
                    else {
                        u64 rr = sh ? (buf[j] >> sh) | (buf[j + 1] << (64 - sh)) : buf[j];
                        u64 qq = k == 0 ? T.q0 : T.qword(k);
                        u64 x = rr ^ qq;
                        if (x) {
                            int b = (__ffsll((long long)x) - 1) >> 1;
                            int l = 32 * k + b;
                            if (l >= cap) res = cap;
                            else { res = l; less = ((rr >> (2 * b)) & 3) < ((qq >> (2 * b)) & 3); }
                        }
                    }
                }
            }
            if (res < 0) T.pac_k0 += 4;                   // four more words of the same entry
            else {
                T.set_rl(e, res);
                if (less) T.r_less |= 1u << e;
                T.undet &= ~(1u << e);
                if (T.undet) { T.pac_e = __ffs(T.undet) - 1; T.pac_k0 = ix.keys ? 1 : 0; }
                else T.mreq = M_NONE;
            }
        } else if (req == M_RMI) {
            T.rmi_a = __longlong_as_double((long long)buf[0]);
            T.rmi_b = __longlong_as_double((long long)buf[1]);
            T.rmi_e = buf[2];
            T.mreq = M_NONE;
        } else if (req == M_LD) {
            T.ld_val = buf[0];
            T.mreq = M_NONE;
        } else if (req == M_T2REC) {
            // Task2 {read, se, cnt, full, lo}
            T.rid = (u32)buf[0];
            u32 se = (u32)(buf[0] >> 32), cnt = (u32)buf[1], full = (u32)(buf[1] >> 32);
            int qb = se & 0xffff, qe = se >> 16;
            T.pivot = (qb + qe) >> 1;
            T.min_intv = (int)cnt + 1;
            T.full_match = full != 0;
            T.max_refpos = full ? buf[2] : 0;
            T.use_cached = !full && ix.mode == 3 && cnt >= 1 && cnt < 10;
            T.cache_pivot = qb; T.cache_pivot_end = qe; T.cache_refpos = buf[2];
            if (T.rid == 0xffffffffu) { T.mreq = M_NONE; }   // unused slot: tstate 1 will skip it
            else T.mreq = M_TASK;
        } else if (req == M_TASK) {
            T.mreq = M_NONE;
        }
    }
    T.flush_slots();
    atomicAdd((unsigned long long *)&A.ctrl[C_SEARCH], (unsigned long long)T.st_search);
    atomicAdd((unsigned long long *)&A.ctrl[C_BLOCK], (unsigned long long)T.st_block);
    atomicAdd((unsigned long long *)&A.ctrl[C_KARY], (unsigned long long)T.st_kary);
    atomicAdd((unsigned long long *)&A.ctrl[C_EXT], (unsigned long long)T.st_ext);
    atomicAdd((unsigned long long *)&A.ctrl[C_PAC], (unsigned long long)T.st_pac);
    atomicAdd((unsigned long long *)&A.ctrl[C_ITER], (unsigned long long)T.st_iter);
}

// ---------------------------------------------------------------------------------------------
// 2-bit packing of the reads: one thread per (read, 32-base word)  (reference src/bwamem.cpp:L1277-1344)
__global__ void k_pack_reads(const uint8_t *__restrict__ codes, const int64_t *__restrict__ read_off, int64_t n_reads,
                             u64 *__restrict__ packed, u32 *__restrict__ nmask)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_reads * TPR_RW) return;
    int64_t r = t / TPR_RW;
    int w = (int)(t % TPR_RW);
    int64_t o = read_off[r];
    int L = (int)(read_off[r + 1] - o);
    u64 v = 0; u32 m = 0;
    for (int k = 0; k < 32; k++) {
        int p = w * 32 + k;
        if (p < L) {
            u32 c = codes[o + p];
            if (c > 3) m |= 1u << k; else v |= (u64)c << (2 * k);
        }
    }
    packed[t] = v;
    nmask[t] = m;
}

__global__ void k_counts2(const u32 *__restrict__ nsmem, const u64 *__restrict__ nhit, int64_t n,
                          int64_t *__restrict__ smem_off, int64_t *__restrict__ hit_off)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    smem_off[i] = i < n ? (int64_t)nsmem[i] : 0;
    hit_off[i] = i < n ? (int64_t)nhit[i] : 0;
}

__device__ __forceinline__ u64 sa_pos2(const IndexView &ix, u64 i)
{
    u64 p = __ldg(ix.pos_lo + i);
    if (ix.pos_wide) p |= (u64)__ldg(ix.pos_hi + i) << 32;
    return p;
}

// read-ordered SMEM records + hit lists from the sorted pool: 8 lanes per read
__global__ void k_gather2(IndexView ix, const u64 *__restrict__ keys, const PoolRec *__restrict__ vals,
                          const int64_t *__restrict__ smem_off, const int64_t *__restrict__ hit_off, int64_t n_reads,
                          int max_occ, meme_smem_t *__restrict__ smems, u64 *__restrict__ hits)
{
    const int gl = threadIdx.x & 7;
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (r >= n_reads) return;
    int64_t s0 = smem_off[r];
    int ns = (int)(smem_off[r + 1] - s0);
    u64 *hdst = hits + hit_off[r];
    int64_t hb = 0;
    for (int k = 0; k < ns; k++) {
        PoolRec a = vals[s0 + k];
        u64 key = keys[s0 + k];
        u64 c = a.cnt, step = 1, take = c;
        if (max_occ > 0) {
            step = c > (u64)max_occ ? c / max_occ : 1;
            take = (c + step - 1) / step;
            if (take > (u64)max_occ) take = (u64)max_occ;
        }
        for (u64 j = gl; j < take; j += 8) hdst[hb + j] = sa_pos2(ix, a.lo + j * step);
        if (gl == 0) {
            meme_smem_t m;
            m.start = (int32_t)((key >> 16) & 0xffff); m.end = (int32_t)(key & 0xffff);
            m.hitbeg = (int32_t)hb; m.hitcount = (int32_t)a.cnt;
            m.cache_refpos = sa_pos2(ix, a.lo);
            smems[s0 + k] = m;
        }
        hb += (int64_t)take;
    }
}

}  // namespace tpr

#include "meme_seed_flat.cuh"

// ---------------------------------------------------------------------------------------------
// host pipeline
// ---------------------------------------------------------------------------------------------
using namespace tpr;

// MEME_SEED_IMPL: "flat" (default) = flat state machine, "tpr" = coroutine formulation, "coop" = group per read
static int tpr_variant()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("MEME_SEED_IMPL"); v = (e && !strcmp(e, "tpr")) ? 1 : 0; }
    return v;
}

template <class K>
static int launch_kernel(meme_index_t *ix, const Args &A, K kernel, size_t dyn_smem = 0)
{
    int sms = 0, per_sm = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ix->device));
    if (dyn_smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TPR_THREADS, dyn_smem));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)sms * per_sm, need = (A.n_tasks + TPR_THREADS - 1) / TPR_THREADS;
    if (grid > need) grid = need;
    CUDA_TRY(cudaMemsetAsync(ix->d_ctrl + C_WORK, 0, 8, ix->stream));
    kernel<<<(unsigned)grid, TPR_THREADS, dyn_smem, ix->stream>>>(ix->v, A);
    CUDA_TRY(cudaGetLastError());
    ix->launches++;
    return MEME_OK;
}

template <int KIND>
static int launch_tpr(meme_index_t *ix, const Args &A)
{
    if (A.n_tasks <= 0) return MEME_OK;
    if (tpr_variant() == 1) return launch_kernel(ix, A, k_tpr<KIND>);
    static int sync = getenv("MEME_SEED_SYNC") ? atoi(getenv("MEME_SEED_SYNC")) : 1;   // warp reconvergence per iteration
    // work counters (meme_seed_counters [0..3]) cost registers and code in the hot loop: off unless asked for
    const char *se = getenv("MEME_SEED_STATS");
    const bool stats = se && atoi(se) != 0;
    const size_t sm = (size_t)TPR_THREADS * flat_sw<KIND>() * sizeof(u32);
    if (ix->v.pos_wide)
        return stats ? launch_kernel(ix, A, k_flat<KIND, true, 2, true, true>, sm) : launch_kernel(ix, A, k_flat<KIND, true, 2, true, false>, sm);
    if (!sync) return launch_kernel(ix, A, k_flat<KIND, false, 2, false, false>, sm);
    static int minb = getenv("MEME_SEED_MINB") ? atoi(getenv("MEME_SEED_MINB")) : 2;   // resident CTAs per SM asked of the compiler
    if (minb == 3 && !stats) return launch_kernel(ix, A, k_flat<KIND, false, 3, true, false>, sm);
    return stats ? launch_kernel(ix, A, k_flat<KIND, false, FLAT_MINB, true, true>, sm) : launch_kernel(ix, A, k_flat<KIND, false, FLAT_MINB, true, false>, sm);
}

int tpr_seed_run(meme_index_t *ix, const meme_seed_opt_t *opt)
{
    cudaStream_t st = ix->stream;
    const int64_t n = ix->n_reads;
    int rc;
    // ---- workspace ----
    if ((rc = meme_grow_device((void **)&ix->d_packed, &ix->cap_packed, (size_t)n * TPR_RW + 16, 8, st))) return rc;
    if ((rc = meme_grow_device((void **)&ix->d_nmask, &ix->cap_nmask, (size_t)n * TPR_RW + 16, 4, st))) return rc;
    if ((rc = meme_grow_device((void **)&ix->d_rd_full, &ix->cap_rd_full, (size_t)n + 16, 1, st))) return rc;
    if ((rc = meme_grow_device((void **)&ix->d_rd_maxref, &ix->cap_rd_maxref, (size_t)n + 16, 8, st))) return rc;
    if (!ix->d_pool_keys) {
        size_t want = (size_t)n * 14 + 65536;
        if ((rc = meme_grow_device((void **)&ix->d_pool_keys, &ix->cap_pool_keys, want, 8, st))) return rc;
    }
    if (!ix->d_t2) {
        size_t want = (size_t)n * 3 + 65536;
        if ((rc = meme_grow_device((void **)&ix->d_t2, &ix->cap_t2, want, sizeof(Task2), st))) return rc;
    }
    CUDA_TRY(cudaEventRecord(ix->ev[0], st));
    k_pack_reads<<<(unsigned)((n * TPR_RW + 255) / 256), 256, 0, st>>>(ix->d_codes, ix->d_read_off, n, ix->d_packed, ix->d_nmask);
    ix->launches++;
    u64 n_pool = 0, n_t2 = 0;
    for (int attempt = 0;; attempt++) {
        // pool_vals / sorted copies follow the key capacity
        size_t cap = ix->cap_pool_keys;
        if ((rc = meme_grow_device((void **)&ix->d_pool_vals, &ix->cap_pool_vals, cap, sizeof(PoolRec), st))) return rc;
        if ((rc = meme_grow_device((void **)&ix->d_pool_keys2, &ix->cap_pool_keys2, cap, 8, st))) return rc;
        if ((rc = meme_grow_device((void **)&ix->d_pool_vals2, &ix->cap_pool_vals2, cap, sizeof(PoolRec), st))) return rc;
        CUDA_TRY(cudaMemsetAsync(ix->d_ctrl, 0, 16 * 8, st));
        CUDA_TRY(cudaMemsetAsync(ix->d_rd_nsmem, 0, (size_t)n * 4, st));
        CUDA_TRY(cudaMemsetAsync(ix->d_rd_nhit, 0, (size_t)n * 8, st));
        Args A;
        A.packed = ix->d_packed; A.nmask = ix->d_nmask; A.read_off = ix->d_read_off;
        A.t2_in = nullptr; A.t2_out = (Task2 *)ix->d_t2; A.t2_cap = ix->cap_t2;
        A.pool_keys = ix->d_pool_keys; A.pool_vals = (PoolRec *)ix->d_pool_vals;
        A.pool_cap = ix->cap_pool_keys < ix->cap_pool_vals ? ix->cap_pool_keys : ix->cap_pool_vals;
        A.rd_nsmem = ix->d_rd_nsmem; A.rd_nhit = ix->d_rd_nhit; A.rd_full = ix->d_rd_full; A.rd_maxref = ix->d_rd_maxref;
        A.ctrl = ix->d_ctrl; A.opt = *opt;
        A.kind = 1; A.n_tasks = n;
        if ((rc = launch_tpr<1>(ix, A))) return rc;
        CUDA_TRY(cudaMemcpyAsync(ix->h_ctrl, ix->d_ctrl, 16 * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        n_t2 = ix->h_ctrl[C_T2];
        bool ovf = ix->h_ctrl[C_T2_OVF] != 0;
        static const bool dbg = getenv("MEME_SEED_DEBUG") != nullptr;   // per-kernel counters + times on stderr
        auto dump = [&](const char *name, cudaEvent_t e0, cudaEvent_t e1, int64_t tasks) {
            u64 c[16]; float ms = 0;
            cudaMemcpyAsync(c, ix->d_ctrl, sizeof c, cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            cudaEventElapsedTime(&ms, e0, e1);
            fprintf(stderr, "[meme_seed] %s: %lld tasks, %.2f ms (cumulative counters) searches %llu blocks %llu kary %llu ext %llu pac %llu "
                    "round-trips %llu\n", name, (long long)tasks, ms, c[C_SEARCH], c[C_BLOCK], c[C_KARY], c[C_EXT], c[C_PAC], c[C_ITER]);
        };
        cudaEvent_t de[4];
        if (dbg) { for (auto &e : de) cudaEventCreate(&e); cudaEventRecord(de[1], st); dump("T1", ix->ev[0], de[1], n); }
        if (!ovf && opt->rounds >= 2 && n_t2) {
            A.kind = 2; A.n_tasks = (int64_t)n_t2; A.t2_in = (const Task2 *)ix->d_t2; A.t2_out = nullptr; A.t2_cap = 0;
            if ((rc = launch_tpr<2>(ix, A))) return rc;
            if (dbg) { cudaEventRecord(de[2], st); dump("T2", de[1], de[2], (int64_t)n_t2); }
        }
        if (!ovf && opt->rounds >= 3 && opt->max_mem_intv > 0) {
            A.kind = 3; A.n_tasks = n; A.t2_in = nullptr;
            if ((rc = launch_tpr<3>(ix, A))) return rc;
            if (dbg) { cudaEventRecord(de[3], st); dump("T3", de[2], de[3], n); }
        }
        if (dbg) for (auto &e : de) cudaEventDestroy(e);
        CUDA_TRY(cudaEventRecord(ix->ev[1], st));
        CUDA_TRY(cudaMemcpyAsync(ix->h_ctrl, ix->d_ctrl, 16 * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        n_pool = ix->h_ctrl[C_POOL];
        if (!ovf && !ix->h_ctrl[C_POOL_OVF]) break;
        if (attempt >= 2) { meme_set_error("seed pools keep overflowing"); return MEME_E_NOMEM; }
        if (ovf && (rc = meme_grow_device((void **)&ix->d_t2, &ix->cap_t2, (size_t)n_t2 * 2 + 65536, sizeof(Task2), st))) return rc;
        if (ix->h_ctrl[C_POOL_OVF] &&
            (rc = meme_grow_device((void **)&ix->d_pool_keys, &ix->cap_pool_keys, (size_t)n_pool * 2 + 65536, 8, st))) return rc;
    }
    // ---- sort by (read, start, end) ----
    CUDA_TRY(cudaEventRecord(ix->ev[2], st));
    {
        int bits = 32;
        while (bits < 64 && (1ull << (bits - 32)) < (u64)n) bits++;
        size_t tmp = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp, ix->d_pool_keys, ix->d_pool_keys2, (PoolRec *)ix->d_pool_vals,
                                        (PoolRec *)ix->d_pool_vals2, (int64_t)n_pool, 0, 64, st);
        if ((rc = meme_grow_device(&ix->d_scan_tmp, &ix->cap_scan_tmp, tmp + 256, 1, st))) return rc;
        size_t t2 = ix->cap_scan_tmp;
        if (n_pool)
            CUDA_TRY(cub::DeviceRadixSort::SortPairs(ix->d_scan_tmp, t2, ix->d_pool_keys, ix->d_pool_keys2, (PoolRec *)ix->d_pool_vals,
                                                     (PoolRec *)ix->d_pool_vals2, (int64_t)n_pool, 0, 64, st));
        (void)bits;
    }
    k_counts2<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(ix->d_rd_nsmem, ix->d_rd_nhit, n, ix->d_smem_off, ix->d_hit_off);
    {
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, ix->d_smem_off, ix->d_smem_off, (int)(n + 1), st);
        if ((rc = meme_grow_device(&ix->d_scan_tmp, &ix->cap_scan_tmp, tmp + 256, 1, st))) return rc;
        size_t t2 = ix->cap_scan_tmp;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(ix->d_scan_tmp, t2, ix->d_smem_off, ix->d_smem_off, (int)(n + 1), st));
        t2 = ix->cap_scan_tmp;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(ix->d_scan_tmp, t2, ix->d_hit_off, ix->d_hit_off, (int)(n + 1), st));
    }
    CUDA_TRY(cudaEventRecord(ix->ev[3], st));
    ix->launches += 4;
    int64_t tot[2];
    CUDA_TRY(cudaMemcpyAsync(&tot[0], ix->d_smem_off + n, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&tot[1], ix->d_hit_off + n, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if ((rc = meme_grow_device((void **)&ix->d_smems, &ix->cap_smems, (size_t)tot[0] + 1, sizeof(meme_smem_t), st))) return rc;
    if ((rc = meme_grow_device((void **)&ix->d_hits, &ix->cap_hits, (size_t)tot[1] + 1, 8, st))) return rc;
    CUDA_TRY(cudaEventRecord(ix->ev[4], st));
    k_gather2<<<(unsigned)((n * 8 + 255) / 256), 256, 0, st>>>(ix->v, ix->d_pool_keys2, (const PoolRec *)ix->d_pool_vals2, ix->d_smem_off,
                                                             ix->d_hit_off, n, opt->max_occ, ix->d_smems, ix->d_hits);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ix->ev[5], st));
    ix->ev_ready = 1;
    ix->launches++;
    ix->h_ctrl[32] = (u64)tot[0];
    ix->h_ctrl[33] = (u64)tot[1];
    // statistics in the slots meme_seed_counters reads
    ix->h_ctrl[8] = ix->h_ctrl[C_SEARCH]; ix->h_ctrl[9] = ix->h_ctrl[C_BLOCK]; ix->h_ctrl[10] = ix->h_ctrl[C_KARY];
    ix->h_ctrl[11] = ix->h_ctrl[C_EXT];
    ix->h_ctrl[2] = 0;                 // no per-read SMEM limit in this path
    ix->h_ctrl[4] = n_t2;              // reseed tasks (reported in the "redo" counter slot)
    ix->h_ctrl[40] = ix->h_ctrl[C_PAC]; ix->h_ctrl[41] = ix->h_ctrl[C_ITER];
    ix->ran = 1;
    return MEME_OK;
}
