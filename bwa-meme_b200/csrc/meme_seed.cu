// meme_seed.cu -- kernel A: P-RMI learned-index SMEM seeding on sm_100a.
//
// Replaces, for a whole chunk of reads at once, the body of mem_kernel1_core_Learned up to the SMEM
// sort (reference src/bwamem.cpp:L1249-1397) = Learned_getSMEMsAllPosOneThread (+ reseeding) and
// Learned_bwtSeedStrategyAllPosOneThread{,_mem_tradeoff} (src/LearnedIndex_seeding.cpp:L913-1466),
// built on one search primitive that stands for right_smem_search / mem_search / *_tradeoff
// (L2131-4189).
//
// B200 design:
//   * one warp per read, reads handed out by an atomic counter (persistent grid = SMs x resident CTAs);
//     the per-read control flow (zig-zag, reseed, third round) is warp-uniform, so there is no
//     divergence between lanes -- lanes only differ in WHICH suffix-array entry they probe;
//   * the reference's last-mile search is a chain of ~6 dependent single-entry probes
//     (binary search in the RMI error window, +-1 linear fix-up, neighbour probes).  Here the 32
//     lanes load a CONTIGUOUS block of 32 SA entries around the prediction in one coalesced
//     request (256 B of keys + 128 B of positions) and a __ballot_sync over "suffix < read"
//     finds the insertion point; the same block already holds the neighbours' LCPs, so the SA
//     interval, the level drops for min_intv > 1 and the third-round "shortest seed" logic are
//     evaluated from registers with shuffles.  Windows wider than 32 entries are narrowed 32-ary
//     (one round trip per factor of 32) and large intervals are bounded by a 32-ary gallop;
//   * 2-bit compare: XOR of the inline 64-bit key with the read's 32-base word + __ffsll; only
//     lanes whose first 32 bases match touch the packed reference (funnel-shifted 64-bit words);
//   * RMI leaf evaluation is the reference's fp64 fma on a 32-byte re-laid-out record (one sector);
//   * results are canonical (they depend on the suffix array only), which is what makes this
//     cooperative formulation bit-exact with the reference's sequential one.
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <utility>
#include "meme_common.cuh"

#define BLOCK_THREADS 256
#define RW 12                           // 64-bit words per packed strand (MEME_MAX_READ_LEN/32 + slack)
#define NMW (MEME_MAX_READ_LEN / 32)    // 32-bit N-mask words per strand
#define MAXS MEME_MAX_SMEMS_PER_READ
#define FULL 0xffffffffu

struct RawSmem {            // 16 bytes
    u32 se;                 // start | end << 16
    u32 cnt;                // SA interval size
    u64 lo;                 // first SA index of the interval
};

template <int MAXS_T>
struct GroupSm {
    u64 fw[RW], rc[RW];
    u32 nfw[NMW], nrc[NMW];
    RawSmem sm[MAXS_T];
};

enum { CT_WORK = 0, CT_POOL = 1, CT_OVF = 2, CT_POOL_OVF = 3, CT_REDO = 4, CT_SEARCH = 8, CT_BLOCK = 9, CT_KARY = 10, CT_EXT = 11 };
enum { START_RMI = 0, START_ISA_EXACT = 1, START_ISA_NEAR = 2 };

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 interleave32(u32 lo_bits, u32 hi_bits)
{   // bit k of lo_bits -> bit 2k, bit k of hi_bits -> bit 2k+1
    u64 a = lo_bits, b = hi_bits;
    a = (a | (a << 16)) & 0x0000FFFF0000FFFFull; b = (b | (b << 16)) & 0x0000FFFF0000FFFFull;
    a = (a | (a << 8)) & 0x00FF00FF00FF00FFull;  b = (b | (b << 8)) & 0x00FF00FF00FF00FFull;
    a = (a | (a << 4)) & 0x0F0F0F0F0F0F0F0Full;  b = (b | (b << 4)) & 0x0F0F0F0F0F0F0F0Full;
    a = (a | (a << 2)) & 0x3333333333333333ull;  b = (b | (b << 2)) & 0x3333333333333333ull;
    a = (a | (a << 1)) & 0x5555555555555555ull;  b = (b | (b << 1)) & 0x5555555555555555ull;
    return a | (b << 1);
}

__device__ __forceinline__ u64 text32(const u64 *__restrict__ pac, u64 p)
{
    u64 w = p >> 5;
    int sh = (int)(p & 31) * 2;
    u64 a = __ldg(pac + w);
    if (sh == 0) return a;
    return (a >> sh) | (__ldg(pac + w + 1) << (64 - sh));
}

__device__ __forceinline__ u64 qword(const u64 *qw, int from)
{   // 32 read bases starting at `from` (shared memory, zero past the end)
    int w = from >> 5, sh = (from & 31) * 2;
    if (w >= RW - 1) return w < RW ? qw[w] >> sh : 0;
    u64 a = qw[w];
    if (sh == 0) return a;
    return (a >> sh) | (qw[w + 1] << (64 - sh));
}

__device__ __forceinline__ int next_n(const u32 *nm, int from, int L)
{   // first N at or after `from` (or L)
    int w = from >> 5;
    u32 m = nm[w] >> (from & 31);
    if (m) { int p = from + __ffs(m) - 1; return p < L ? p : L; }
    for (w++; w * 32 < L; w++)
        if (nm[w]) { int p = w * 32 + __ffs(nm[w]) - 1; return p < L ? p : L; }
    return L;
}

__device__ __forceinline__ u64 sa_pos(const IndexView &ix, u64 i)
{
    u64 p = __ldg(ix.pos_lo + i);
    if (ix.pos_wide) p |= (u64)__ldg(ix.pos_hi + i) << 32;
    return p;
}

__device__ __forceinline__ u64 isa_at(const IndexView &ix, u64 p)
{
    u64 i = __ldg(ix.isa_lo + p);
    if (ix.pos_wide) i |= (u64)__ldg(ix.isa_hi + p) << 32;
    return i;
}

// lcp(i) = min(LCP(T[SA[i]..], Q), V, sa_num - SA[i]) and "suffix < read" -- the contract of
// compare_read_and_ref_binary* (reference L226-601); exact matches and text-end suffixes are "not less".
// The inline key and the position are fetched together; once the first 32 bases match, up to four
// further 32-base words of the packed reference are requested at once (one DRAM round trip).
__device__ __noinline__ int lcp_at(const IndexView &ix, const u64 *qw, int from, int V, u64 idx, bool &less)
{
    u64 p = sa_pos(ix, idx);
    u64 r = ix.keys ? __ldg(ix.keys + idx) : 0;
    u64 ref_len = ix.sa_num - p;
    int cap = ref_len < (u64)V ? (int)ref_len : V;
    if (!ix.keys) r = text32(ix.pac, p);
    less = false;
    {
        u64 q = qword(qw, from);
        u64 x = r ^ q;
        if (x) {
            int b = (__ffsll((long long)x) - 1) >> 1;
            if (b >= cap) return cap;
            less = ((r >> (2 * b)) & 3) < ((q >> (2 * b)) & 3);
            return b;
        }
        if (32 >= cap) return cap;
    }
    const u64 *wp = ix.pac + (p >> 5);
    const int sh = (int)(p & 31) * 2;
    for (int k0 = 1;; k0 += 4) {
        u64 a[5];
#pragma unroll
        for (int j = 0; j < 5; j++) a[j] = (j == 0 || 32 * (k0 + j - 1) < cap) ? __ldg(wp + k0 + j) : 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int k = k0 + j;
            if (32 * k >= cap) return cap;
            u64 rr = sh ? (a[j] >> sh) | (a[j + 1] << (64 - sh)) : a[j];
            u64 q = qword(qw, from + 32 * k);
            u64 x = rr ^ q;
            if (x) {
                int b = (__ffsll((long long)x) - 1) >> 1;
                int l = 32 * k + b;
                if (l >= cap) return cap;
                less = ((rr >> (2 * b)) & 3) < ((q >> (2 * b)) & 3);
                return l;
            }
        }
    }
}

// P-RMI inference = learned_index_lookup (reference L186-210): fp64 fma on the leaf record,
// optional partial third layer, clamp to [0, sa_num-1].
__device__ __forceinline__ u64 fclamp(double v, double bound)
{
    if (v < 0.0) return 0;
    return v > bound ? (u64)bound : (u64)v;
}

__device__ __noinline__ u64 rmi_lookup(const IndexView &ix, u64 key, u64 &cerr, u64 &err)
{
    const RmiRec *rec = ix.l2 + (key >> ix.bit_shift);
    double a = __ldg(&rec->alpha), b = __ldg(&rec->beta);
    u64 e = __ldg(&rec->err);
    double kd = __ull2double_rn(key);
    double p = fma(b, kd, a);
    if (e >> 63) {
        u64 st = (e >> 32) & 0x7fffffff;
        double num = (double)(e & 0xffffffffull);
        rec = ix.l1 + st + fclamp(p, num - 1);
        a = __ldg(&rec->alpha); b = __ldg(&rec->beta); e = __ldg(&rec->err);
        p = fma(b, kd, a);
    }
    cerr = (e >> 32) & 0x3fffffff;
    err = e & 0x7fffffff;
    return fclamp(p, (double)ix.sa_num - 1.0);
}

// RMI key = Tokenization (reference L795-901): 32 bases, first base in the MSBs, positions at or
// after min(V,32) padded with 3.
__device__ __forceinline__ u64 rmi_key(u64 q0, int V)
{
    if (V < 32) q0 |= ~0ull << (2 * V);
    u64 y = __brevll(q0);
    return ((y & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((y & 0x5555555555555555ull) << 1);
}

// ---------------------------------------------------------------------------------------------
// per-group search machinery: G lanes (8, 16 or 32) cooperate on one read
// ---------------------------------------------------------------------------------------------
template <int G, int MAXS_T>
struct Searcher {
    static constexpr u32 GM = G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1);
    const IndexView &ix;
    GroupSm<MAXS_T> *ws;
    int gl;                 // lane within the group
    int gshift;             // first warp lane of the group
    u32 gmask;              // warp mask of the group
    int L;                  // read length
    int n_sm;               // SMEMs collected so far
    int budget;
    bool overflow;
    u32 st_search, st_block, st_kary, st_ext;
    const u64 *qw; int qfrom;           // query of the search in flight
    u64 base; int nvalid; int lcp;      // block state; lcp: this lane's value (-1 if lane invalid)
    u32 lessmask;
    u64 lo, hi; int nbl, nbr; int lev;  // level state

    __device__ Searcher(const IndexView &ix_, GroupSm<MAXS_T> *ws_, int lane) : ix(ix_), ws(ws_)
    {
        gl = lane % G;
        gshift = lane - gl;
        gmask = GM << gshift;
    }
    __device__ __forceinline__ u32 gballot(bool p) const { return (__ballot_sync(gmask, p) >> gshift) & GM; }
    template <typename T> __device__ __forceinline__ T gshfl(T v, int src) const { return __shfl_sync(gmask, v, src, G); }
    __device__ __forceinline__ static int lead(u32 m) { return m == GM ? G : __ffs(~m) - 1; }   // leading ones

    __device__ __noinline__ void probe_block(u64 b, int V)
    {   // contiguous block [b, b+G): full lcp per lane
        u64 n = ix.sa_num;
        base = b;
        u64 idx = b + gl;
        bool valid = idx < n;
        bool less = false;
        lcp = valid ? lcp_at(ix, qw, qfrom, V, idx, less) : -1;
        nvalid = __popc(gballot(valid));
        lessmask = gballot(valid && less);
        st_block++;
    }

    // Locate a block that contains a suffix with the maximal lcp. center/rad = hint window.
    __device__ __noinline__ void locate(int V, u64 center, u64 rad_lo, u64 rad_hi)
    {
        const u64 n = ix.sa_num;
        u64 lo_b = 0, hi_b = n;                 // insertion point t in [lo_b, hi_b]
        // the prediction error is concentrated near zero: if the window is not much wider than a
        // block, first try the block centred on the prediction itself
        if (rad_lo + rad_hi + 2 > (u64)G && rad_lo + rad_hi <= (u64)(3 * G)) {
            u64 b = center > (u64)(G / 2) ? center - G / 2 : 0;
            if (n >= (u64)G && b > n - G) b = n - G;
            probe_block(b, V);
            int c = lead(lessmask);
            bool left_ok = c > 0 || b == 0;
            bool right_ok = c < nvalid || b + nvalid == n;
            if (left_ok && right_ok) return;
            if (!left_ok) hi_b = b; else lo_b = b + nvalid;
        }
        for (int it = 0; it < 48; it++) {
            u64 pa = center > rad_lo ? center - rad_lo : 0;
            pa = pa < lo_b ? lo_b : (pa > hi_b ? hi_b : pa);
            u64 pb = center + rad_hi + 1;
            pb = pb < lo_b ? lo_b : (pb > hi_b ? hi_b : pb);
            u64 span = pb - pa;
            if (span <= (u64)(G - 2)) {
                // contiguous: entries pa-1 .. pb (span + 2 of them), centred in the block
                u64 slack = (G - 2 - span) / 2;
                u64 b = pa > 1 + slack ? pa - 1 - slack : 0;
                if (n >= (u64)G && b > n - G) b = n - G;
                probe_block(b, V);
                int c = lead(lessmask);
                bool left_ok = c > 0 || b == 0;
                bool right_ok = c < nvalid || b + nvalid == n;
                if (left_ok && right_ok) return;
                if (!left_ok) hi_b = b; else lo_b = b + nvalid;
            } else {
                u64 stride = (span + G - 1) / G;
                u64 idx = pa + (u64)(gl + 1) * stride - 1;
                bool valid = idx < pb;
                bool less = false;
                if (valid) lcp_at(ix, qw, qfrom, V, idx, less);
                int nv = __popc(gballot(valid));
                int c = lead(gballot(valid && less));
                if (c > nv) c = nv;
                if (c > 0) { u64 nl = pa + (u64)c * stride; if (nl > lo_b) lo_b = nl; }
                if (c < nv) { u64 nh = pa + (u64)(c + 1) * stride - 1; if (nh < hi_b) hi_b = nh; }
                st_kary++;
            }
            // widen the tentative window (the hint may be wrong: the reference's linear fix-up,
            // L2297-2357, makes the same guarantee)
            rad_lo = rad_lo * G + G;
            rad_hi = rad_hi * G + G;
        }
        // not reached for a sorted suffix array; keep the last block
    }

    // Furthest index in direction dir (+1/-1) reachable from `from` (known to satisfy lcp >= Lreq)
    // through entries that all satisfy lcp >= Lreq; nb = lcp of the first entry that does not (0 at
    // the array ends).  G-ary gallop then G-ary narrowing.  Stops early once `limit` entries
    // beyond `from` are known to qualify (the exact edge is then not needed).
    __device__ __noinline__ void extend(int dir, u64 from, int Lreq, u64 limit, u64 &edge, int &nb)
    {
        const u64 n = ix.sa_num;
        u64 cur = from, s = 1;
        bool up = true;
        for (int guard = 0;; guard++) {
            if (guard > 400) { overflow = true; edge = cur; nb = 0; return; }   // cannot happen on a sorted SA
            u64 off = (u64)(gl + 1) * s;
            bool valid;
            u64 idx;
            if (dir > 0) { idx = cur + off; valid = idx < n && idx > cur; }
            else { valid = cur >= off; idx = cur - off; }
            bool less;
            int l = valid ? lcp_at(ix, qw, qfrom, Lreq, idx, less) : 0;
            int c = lead(gballot(valid && l >= Lreq));
            st_ext++;
            cur = dir > 0 ? cur + (u64)c * s : cur - (u64)c * s;
            if (c == G) {
                u64 gone = dir > 0 ? cur - from : from - cur;
                if (gone >= limit) { edge = cur; nb = Lreq; return; }
                if (up && s < (1ull << 40)) s *= G;
                continue;
            }
            if (s == 1) { nb = gshfl(l, c); edge = cur; return; }
            up = false;
            s /= G;
        }
    }

    // Start the level walk from the located block: level = maximal lcp, [lo,hi] = its interval.
    // The interval is made exact (extended past the block) only if it will be emitted
    // (exact_from <= level) or if it is still smaller than `need`.
    __device__ __noinline__ void level_init(u64 need, int exact_from)
    {
        const u64 n = ix.sa_num;
        int m = lcp;
#pragma unroll
        for (int o = G / 2; o; o >>= 1) m = max(m, __shfl_xor_sync(gmask, m, o, G));
        lev = m;
        u32 mask = gballot(lcp >= lev && lcp >= 0);
        int anchor = __ffs(mask) - 1;
        // contiguous run of qualifying lanes that contains the anchor
        u32 above = ~mask & ~((2u << anchor) - 1);       // zeros above the anchor (bits >= G count as zeros of the run)
        int last = above ? __ffs(above) - 2 : 31;
        if (last > nvalid - 1) last = nvalid - 1;
        int first = anchor;
        lo = base + first; hi = base + last;
        const bool exact = lev >= exact_from;
        if (!exact && hi - lo + 1 >= need) { nbl = nbr = lev; return; }
        const u64 limit = exact ? ~0ull : need;
        if (first > 0) nbl = gshfl(lcp, first - 1);
        else if (base == 0) nbl = 0;
        else extend(-1, lo, lev, limit, lo, nbl);
        if (last < nvalid - 1) nbr = gshfl(lcp, last + 1);
        else if (base + nvalid == n) nbr = 0;
        else extend(+1, hi, lev, limit, hi, nbr);
    }

    // Drop to the next level (the larger neighbour lcp) and regrow the interval (reference L2569-2574).
    __device__ __noinline__ void level_next(u64 limit)
    {
        const u64 n = ix.sa_num;
        lev = max(nbl, nbr);
        if (lev <= 0) { lev = 0; lo = 0; hi = n - 1; nbl = nbr = 0; return; }
        u32 mask = gballot(lcp >= lev && lcp >= 0);
        if (nbl >= lev) {
            if (lo > base && lo <= base + (u64)nvalid) {          // lo-1 lies in the block
                int p = (int)(lo - base);
                u32 x = ~mask & (p >= 32 ? FULL : ((1u << p) - 1));
                if (x == 0) {
                    lo = base;
                    if (base == 0) nbl = 0; else extend(-1, lo, lev, limit, lo, nbl);
                } else {
                    int z = 31 - __clz(x);
                    lo = base + z + 1;
                    nbl = gshfl(lcp, z);
                }
            } else if (lo == 0) nbl = 0;
            else extend(-1, lo, lev, limit, lo, nbl);
        }
        if (nbr >= lev) {
            if (hi + 1 >= base && hi + 1 < base + (u64)nvalid) {  // hi+1 lies in the block
                int p = (int)(hi + 1 - base);
                u32 x = ~mask & ~((1u << p) - 1) & (nvalid >= 32 ? FULL : ((1u << nvalid) - 1));
                if (x == 0) {
                    hi = base + nvalid - 1;
                    if (base + nvalid == n) nbr = 0; else extend(+1, hi, lev, limit, hi, nbr);
                } else {
                    int z = __ffs(x) - 1;
                    hi = base + z - 1;
                    nbr = gshfl(lcp, z);
                }
            } else if (hi == n - 1) nbr = 0;
            else extend(+1, hi, lev, limit, hi, nbr);
        }
    }

    __device__ void begin(int strand, int from, int &V)
    {
        qw = strand ? ws->rc : ws->fw;
        qfrom = from;
        V = next_n(strand ? ws->nrc : ws->nfw, from, L) - from;
        st_search++;
    }

    __device__ __noinline__ void locate_by(int V, int start_mode, u64 hint)
    {
        if (start_mode == START_ISA_EXACT) {
            u64 b = hint > (u64)(G / 2 - 1) ? hint - (G / 2 - 1) : 0;
            if (ix.sa_num >= (u64)G && b > ix.sa_num - G) b = ix.sa_num - G;
            probe_block(b, V);
        } else if (start_mode == START_ISA_NEAR) {
            locate(V, hint, G / 4, G / 4);
        } else {
            u64 cerr, err;
            u64 est = rmi_lookup(ix, rmi_key(qword(qw, qfrom), V), cerr, err);
            locate(V, est, cerr, err);
        }
    }

    __device__ void push(int start, int end, u64 lo_, u64 cnt)
    {
        if (n_sm >= MAXS_T) { overflow = true; return; }
        if (gl == 0) {
            RawSmem r;
            r.se = (u32)start | ((u32)end << 16);
            r.cnt = cnt > 0xffffffffull ? 0xffffffffu : (u32)cnt;
            r.lo = lo_;
            ws->sm[n_sm] = r;
        }
        n_sm++;
    }

    // right_smem_search / mem_search and their _tradeoff twins: level of the query with at least
    // min_intv occurrences; emits the SMEM when asked and long enough.
    __device__ __noinline__ int search(int pivot, bool right, bool emit, int min_intv, int min_seed_len, int start_mode, u64 hint,
                          u64 *pos_out)
    {
        int V;
        begin(right ? 0 : 1, right ? pivot : L - 1 - pivot, V);
        locate_by(V, start_mode, hint);
        u64 need = (u64)min_intv;
        if (++budget > 20000) overflow = true;                        // runaway guard (never hit on valid input)
        level_init(need, emit ? min_seed_len : 0x7fffffff);
        // an interval is exact as long as every extension ran to its end; extensions are only cut
        // short (limit) when the level is too short to be emitted
        while (hi - lo + 1 < need && !overflow) {
            bool exact_needed = emit && max(nbl, nbr) >= min_seed_len;
            level_next(exact_needed ? ~0ull : need);
        }
        if (emit && lev >= min_seed_len) push(pivot, pivot + lev, lo, hi - lo + 1);
        if (pos_out) *pos_out = lo;
        return lev;
    }
};

// ---------------------------------------------------------------------------------------------
// the seeding kernel
// ---------------------------------------------------------------------------------------------
struct ReadState {
    bool full_match; u64 max_refpos;
    bool use_cached; int cache_pivot, cache_pivot_end; u64 cache_refpos;
    int min_intv, min_seed_len;
};

template <class S_t>
__device__ __noinline__ int do_search(S_t &S, ReadState &R, int pivot, bool right, bool emit)
{   // start selection of reference L1741-1816 / L1926-2040
    const IndexView &ix = S.ix;
    if (ix.mode == 3) {
        if (R.full_match) {
            u64 rp = right ? R.max_refpos + pivot : ix.sa_num - R.max_refpos - pivot - 1;
            return S.search(pivot, right, emit, R.min_intv, R.min_seed_len, START_ISA_EXACT, isa_at(ix, rp), nullptr);
        }
        if (R.use_cached) {
            if (right && R.cache_pivot_end >= pivot + 18 && pivot >= R.cache_pivot)
                return S.search(pivot, true, emit, R.min_intv, R.min_seed_len, START_ISA_NEAR,
                                isa_at(ix, R.cache_refpos + pivot - R.cache_pivot), nullptr);
            if (!right && pivot < R.cache_pivot_end && R.cache_pivot + 18 <= pivot)
                return S.search(pivot, false, emit, R.min_intv, R.min_seed_len, START_ISA_NEAR,
                                isa_at(ix, ix.sa_num + R.cache_pivot - R.cache_refpos - pivot - 1), nullptr);
        }
    }
    return S.search(pivot, right, emit, R.min_intv, R.min_seed_len, START_RMI, 0, nullptr);
}

__device__ __forceinline__ bool is_n(const u32 *nm, int p) { return (nm[p >> 5] >> (p & 31)) & 1; }

struct SeedArgs {
    const uint8_t *codes; const int64_t *read_off; int64_t n_reads;
    const u32 *worklist; int64_t n_work;        // null: work items are reads 0..n_reads-1
    u32 *redo;                                  // null in the fallback launch
    RawSmem *pool; u64 pool_cap;
    u64 *rd_pool_off; u32 *rd_nsmem; u64 *rd_nhit; u64 *ctrl;
    meme_seed_opt_t opt;
};

template <int G, int MAXS_T, int MINB>
__global__ void __launch_bounds__(BLOCK_THREADS, MINB) k_seed(IndexView ix, SeedArgs A)
{
    constexpr int GROUPS = BLOCK_THREADS / G;
    __shared__ GroupSm<MAXS_T> gsm[GROUPS];
    const int lane = threadIdx.x & 31;
    typedef Searcher<G, MAXS_T> S_t;
    GroupSm<MAXS_T> *ws = &gsm[threadIdx.x / G];
    S_t S(ix, ws, lane);
    const int gl = S.gl;
    const u32 gmask = S.gmask;
    const meme_seed_opt_t &opt = A.opt;
    S.st_search = S.st_block = S.st_kary = S.st_ext = 0;
    for (;;) {
        u64 wi = 0;
        if (gl == 0) wi = atomicAdd((unsigned long long *)&A.ctrl[CT_WORK], 1ull);
        wi = S.gshfl(wi, 0);
        if (wi >= (u64)A.n_work) break;
        const u64 r = A.worklist ? A.worklist[wi] : wi;
        const int64_t o0 = A.read_off[r];
        const int L = (int)(A.read_off[r + 1] - o0);
        const uint8_t *seq = A.codes + o0;
        __syncwarp(gmask);
        // ---- pack the read and its reverse complement (reference src/bwamem.cpp:L1277-1344) ----
        const int nw = (L + 31) >> 5;
        for (int w = 0; w < RW; w++) {
            u32 f0 = 0, f1 = 0, r0 = 0, r1 = 0, mf = 0, mr = 0;
            if (w < nw) {
#pragma unroll
                for (int sub = 0; sub < 32 / G; sub++) {
                    int p = w * 32 + sub * G + gl;
                    u32 cf = p < L ? seq[p] : 0, cr = p < L ? seq[L - 1 - p] : 0;
                    bool nf = cf > 3, nr = cr > 3;
                    cf = nf ? 0 : cf;
                    cr = (nr || p >= L) ? 0 : 3 - cr;
                    f0 |= S.gballot(cf & 1) << (sub * G); f1 |= S.gballot(cf & 2) << (sub * G);
                    r0 |= S.gballot(cr & 1) << (sub * G); r1 |= S.gballot(cr & 2) << (sub * G);
                    mf |= S.gballot(nf) << (sub * G); mr |= S.gballot(nr) << (sub * G);
                }
            }
            if (gl == 0) {
                ws->fw[w] = interleave32(f0, f1);
                ws->rc[w] = interleave32(r0, r1);
                if (w < NMW) { ws->nfw[w] = mf; ws->nrc[w] = mr; }
            }
        }
        __syncwarp(gmask);
        S.L = L; S.n_sm = 0; S.overflow = false; S.budget = 0;
        ReadState R;
        R.full_match = false; R.use_cached = false; R.max_refpos = 0;
        R.min_intv = 1; R.min_seed_len = opt.min_seed_len;
        const u32 *nf = ws->nfw;
        // ---- rounds 1+2: Learned_getSMEMsAllPosOneThread (reference L913-972) ----
        int pivot = 0;
        while (pivot < L && !S.overflow) {
            int before = S.n_sm;
            // --- Learned_getSMEMsOnePosOneThread_step1 (L1691-1894) ---
            if (is_n(nf, pivot)) {
                pivot = (L - pivot < R.min_seed_len) ? L : pivot + 1;
            } else if (pivot != 0 && !is_n(nf, pivot - 1)) {
                int sp = pivot;
                while (sp < L && !S.overflow) {
                    if (is_n(nf, sp)) {
                        if (L - sp < R.min_seed_len) { pivot = L; sp = L; }
                        else { sp += 1; pivot += 1; }
                        continue;
                    }
                    int l = do_search(S, R, pivot, false, false);
                    pivot = pivot - l + 1;
                    if (L - pivot < R.min_seed_len) break;
                    int rr = do_search(S, R, pivot, true, true);
                    sp = pivot + rr;
                    pivot = sp;
                }
                pivot = L;
            } else {
                u64 pos;
                int len = S.search(pivot, true, true, 1, R.min_seed_len, START_RMI, 0, &pos);
                if (ix.mode == 3 && len == L) { R.full_match = true; R.max_refpos = sa_pos(ix, pos); }
                pivot += len;
            }
            int after = S.n_sm;
            if (opt.rounds < 2) continue;
            // --- reseeding: Learned_getSMEMsOnePosOneThread (L1897-2126) for long, rare SMEMs ---
            __syncwarp(gmask);
            for (int k = before; k < after && k < MAXS_T && !S.overflow; k++) {
                RawSmem m = ws->sm[k];
                int qb = m.se & 0xffff, qe = m.se >> 16;
                if (qe - qb < opt.split_len || m.cnt > (u32)opt.split_width) continue;
                R.min_intv = (int)m.cnt + 1;
                R.use_cached = false;
                if (ix.mode == 3 && m.cnt >= 1 && m.cnt < 10) {
                    R.use_cached = true; R.cache_pivot = qb; R.cache_pivot_end = qe;
                    R.cache_refpos = sa_pos(ix, m.lo);
                }
                int p2 = (qb + qe) >> 1;
                // pivot sits inside an SMEM: not N, predecessor not N
                int len = do_search(S, R, p2, true, false);
                int next = p2 + len, sp = p2;
                while (sp < next && !S.overflow) {
                    int l = do_search(S, R, p2, false, false);
                    p2 = p2 - l + 1;
                    if (next - p2 < R.min_seed_len) break;
                    int rr = do_search(S, R, p2, true, true);
                    sp = p2 + rr;
                    p2 = sp;
                    __syncwarp(gmask);
                }
                R.use_cached = false;
                R.min_intv = 1;
            }
        }
        // ---- round 3: Learned_bwtSeedStrategyAllPosOneThread{,_mem_tradeoff} (L974-1466) ----
        if (opt.rounds >= 3 && opt.max_mem_intv > 0) {
            const int msl = opt.min_seed_len + 1, mi = opt.max_mem_intv;
            pivot = 0;
            while (pivot < L - msl + 1 && !S.overflow) {
                if (is_n(nf, pivot)) { pivot++; continue; }
                int V;
                S.begin(0, pivot, V);
                if (++S.budget > 20000) S.overflow = true;
                if (V < msl) { S.st_search--; pivot += V; continue; }
                if (R.full_match) S.locate_by(V, START_ISA_EXACT, isa_at(ix, R.max_refpos + pivot));
                else S.locate_by(V, START_RMI, 0);
                S.level_init((u64)mi, 0x7fffffff);
                if (S.lev < msl) { pivot += msl; continue; }
                u64 last_lo = S.lo, last_cnt = 0, cnt;
                int emit_len;
                for (int guard = 0;; guard++) {
                    if (guard > 400) { S.overflow = true; cnt = 0; emit_len = msl; break; }
                    cnt = S.hi - S.lo + 1;
                    if (cnt >= (u64)mi) { cnt = last_cnt; emit_len = S.lev + 1; break; }   // L1243-1251
                    if (max(S.nbl, S.nbr) < msl) { last_lo = S.lo; emit_len = msl; break; } // L1252-1258
                    last_lo = S.lo; last_cnt = cnt;
                    S.level_next((u64)mi);
                }
                if (cnt > 0 && cnt < (u64)mi) S.push(pivot, pivot + max(emit_len, msl), last_lo, cnt);
                pivot += emit_len;
            }
        }
        // ---- sort by (start, end, emission order) = ks_introsort(mem_smem_sort_lt_learned), bwamem.cpp:L1397 ----
        __syncwarp(gmask);
        if (S.overflow && A.redo) {
            // too many SMEMs for the group's shared-memory list: hand the read to the wide fallback launch
            if (gl == 0) {
                u64 k = atomicAdd((unsigned long long *)&A.ctrl[CT_REDO], 1ull);
                A.redo[k] = (u32)r;
                A.rd_pool_off[r] = 0; A.rd_nsmem[r] = 0; A.rd_nhit[r] = 0;
            }
            continue;
        }
        const int ns = S.n_sm;
        u64 poff = 0;
        if (gl == 0) poff = atomicAdd((unsigned long long *)&A.ctrl[CT_POOL], (unsigned long long)ns);
        poff = S.gshfl(poff, 0);
        bool fits = poff + ns <= A.pool_cap;
        u64 nh = 0;
        for (int i = gl; i < ns; i += G) {
            RawSmem a = ws->sm[i];
            u32 ka = ((a.se & 0xffff) << 16) | (a.se >> 16);
            int rank = 0;
            for (int j = 0; j < ns; j++) {
                u32 se = ws->sm[j].se;
                u32 kb = ((se & 0xffff) << 16) | (se >> 16);
                rank += (kb < ka) || (kb == ka && j < i);
            }
            if (fits) A.pool[poff + rank] = a;
            u64 c = a.cnt;
            if (opt.max_occ > 0) {
                u64 step = c > (u64)opt.max_occ ? c / opt.max_occ : 1;
                u64 k = (c + step - 1) / step;
                c = k < (u64)opt.max_occ ? k : (u64)opt.max_occ;
            }
            nh += c;
        }
#pragma unroll
        for (int o = G / 2; o; o >>= 1) nh += __shfl_xor_sync(gmask, nh, o, G);
        if (gl == 0) {
            A.rd_pool_off[r] = poff;
            A.rd_nsmem[r] = fits ? ns : 0;
            A.rd_nhit[r] = fits ? nh : 0;
            if (!fits) atomicExch((unsigned long long *)&A.ctrl[CT_POOL_OVF], 1ull);
            if (S.overflow) atomicAdd((unsigned long long *)&A.ctrl[CT_OVF], 1ull);
        }
    }
    // statistics
    if (gl == 0) {
        atomicAdd((unsigned long long *)&A.ctrl[CT_SEARCH], (unsigned long long)S.st_search);
        atomicAdd((unsigned long long *)&A.ctrl[CT_BLOCK], (unsigned long long)S.st_block);
        atomicAdd((unsigned long long *)&A.ctrl[CT_KARY], (unsigned long long)S.st_kary);
        atomicAdd((unsigned long long *)&A.ctrl[CT_EXT], (unsigned long long)S.st_ext);
    }
}

// per-read counts -> int64 arrays of n+1 entries (last = 0) for the exclusive scans
__global__ void k_counts(const u32 *__restrict__ nsmem, const u64 *__restrict__ nhit, int64_t n,
                         int64_t *__restrict__ smem_off, int64_t *__restrict__ hit_off)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    smem_off[i] = i < n ? (int64_t)nsmem[i] : 0;
    hit_off[i] = i < n ? (int64_t)nhit[i] : 0;
}

// Materialise the read-ordered SMEM records and hit lists: hits = SA[lo + k*step] (reference L2639-2657,
// sampled as mem_chain_Learned consumes them when max_occ > 0, src/bwamem.cpp:L1154-1155).
__global__ void k_gather(IndexView ix, const RawSmem *__restrict__ pool, const u64 *__restrict__ rd_pool_off,
                         const int64_t *__restrict__ smem_off, const int64_t *__restrict__ hit_off, int64_t n_reads,
                         int max_occ, meme_smem_t *__restrict__ smems, u64 *__restrict__ hits)
{
    const int lane = threadIdx.x & 31;
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_reads) return;
    int64_t s0 = smem_off[r];
    int ns = (int)(smem_off[r + 1] - s0);
    const RawSmem *src = pool + rd_pool_off[r];
    u64 *hdst = hits + hit_off[r];
    int64_t hb = 0;
    for (int k = 0; k < ns; k++) {
        RawSmem a = src[k];
        u64 c = a.cnt, step = 1, take = c;
        if (max_occ > 0) {
            step = c > (u64)max_occ ? c / max_occ : 1;
            take = (c + step - 1) / step;
            if (take > (u64)max_occ) take = (u64)max_occ;
        }
        for (u64 j = lane; j < take; j += 32) hdst[hb + j] = sa_pos(ix, a.lo + j * step);
        if (lane == 0) {
            meme_smem_t m;
            m.start = a.se & 0xffff; m.end = a.se >> 16;
            m.hitbeg = (int32_t)hb; m.hitcount = (int32_t)a.cnt;
            m.cache_refpos = sa_pos(ix, a.lo);
            smems[s0 + k] = m;
        }
        hb += (int64_t)take;
    }
}

// ---------------------------------------------------------------------------------------------
// host API
// ---------------------------------------------------------------------------------------------
// input buffers and the per-read arrays of a batch of up to n_reads reads / n_codes bases
static int seed_alloc_inputs(meme_index_t *ix, int64_t n_reads, int64_t n_codes)
{
    int rc;
    cudaStream_t st = ix->stream;
    if ((rc = meme_grow_device((void **)&ix->d_codes, &ix->cap_codes, (size_t)n_codes + 64, 1, st))) return rc;
    size_t old_cap = ix->cap_reads;
    if ((rc = meme_grow_device((void **)&ix->d_read_off, &ix->cap_reads, (size_t)n_reads + 1, 8, st))) return rc;
    if (ix->cap_reads != old_cap || !ix->d_rd_pool_off) {
        size_t c = ix->cap_reads, dummy;
        void **bufs[] = {(void **)&ix->d_rd_pool_off, (void **)&ix->d_rd_nsmem, (void **)&ix->d_rd_nhit,
                         (void **)&ix->d_smem_off, (void **)&ix->d_hit_off, (void **)&ix->d_redo};
        size_t el[] = {8, 4, 8, 8, 8, 4};
        for (int k = 0; k < 6; k++) {
            if (*bufs[k]) { CUDA_TRY(cudaFree(*bufs[k])); *bufs[k] = nullptr; }
            dummy = 0;
            if ((rc = meme_grow_device(bufs[k], &dummy, c, el[k], st))) return rc;
        }
    }
    return MEME_OK;
}

extern "C" int meme_seed_upload(meme_index_t *ix, const uint8_t *codes, const int64_t *read_off, int64_t n_reads)
{
    if (!ix || !read_off || n_reads < 0 || (n_reads > 0 && !codes)) { meme_set_error("bad argument"); return MEME_E_ARG; }
    CUDA_TRY(cudaSetDevice(ix->device));
    int64_t n_codes = n_reads ? read_off[n_reads] - read_off[0] : 0;
    if (n_reads && read_off[0] != 0) { meme_set_error("read_off[0] must be 0"); return MEME_E_ARG; }
    for (int64_t i = 0; i < n_reads; i++) {
        int64_t len = read_off[i + 1] - read_off[i];
        if (len < 0 || len > MEME_MAX_READ_LEN) {
            meme_set_error("read %lld has length %lld (limit %d; the reference asserts <= 301)", (long long)i, (long long)len, MEME_MAX_READ_LEN);
            return MEME_E_ARG;
        }
    }
    int rc;
    cudaStream_t st = ix->stream;
    if ((rc = seed_alloc_inputs(ix, n_reads, n_codes))) return rc;
    ix->piped = 0;
    if (n_codes) CUDA_TRY(cudaMemcpyAsync(ix->d_codes, codes, (size_t)n_codes, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ix->d_read_off, read_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    ix->n_reads = n_reads;
    ix->n_codes = n_codes;
    ix->ran = 0;
    return MEME_OK;
}

template <int G, int MAXS_T, int MINB>
static int launch_one(meme_index_t *ix, const SeedArgs &A)
{
    int dev_sms = 0, per_sm = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, ix->device));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_seed<G, MAXS_T, MINB>, BLOCK_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    const int groups = BLOCK_THREADS / G;
    int64_t need_blocks = (A.n_work + groups - 1) / groups;
    int64_t grid = (int64_t)dev_sms * per_sm;           // persistent: one wave of resident CTAs
    if (grid > need_blocks) grid = need_blocks;
    if (grid < 1) grid = 1;
    k_seed<G, MAXS_T, MINB><<<(unsigned)grid, BLOCK_THREADS, 0, ix->stream>>>(ix->v, A);
    CUDA_TRY(cudaGetLastError());
    ix->launches++;
    return MEME_OK;
}

// lanes per read of the main launch: 8, 16 or 32 (MEME_SEED_GROUP overrides)
static int seed_group()
{
    static int g = 0;
    if (!g) {
        const char *e = getenv("MEME_SEED_GROUP");
        g = e ? atoi(e) : 8;
        if (g != 8 && g != 16 && g != 32) g = 8;
    }
    return g;
}

static int seed_launch(meme_index_t *ix, const meme_seed_opt_t *opt)
{
    cudaStream_t st = ix->stream;
    SeedArgs A;
    A.codes = ix->d_codes; A.read_off = ix->d_read_off; A.n_reads = ix->n_reads;
    A.worklist = nullptr; A.n_work = ix->n_reads; A.redo = ix->d_redo;
    A.pool = (RawSmem *)ix->d_pool; A.pool_cap = ix->cap_pool;
    A.rd_pool_off = ix->d_rd_pool_off; A.rd_nsmem = ix->d_rd_nsmem; A.rd_nhit = ix->d_rd_nhit; A.ctrl = ix->d_ctrl;
    A.opt = *opt;
    CUDA_TRY(cudaMemsetAsync(ix->d_ctrl, 0, 16 * 8, st));
    CUDA_TRY(cudaEventRecord(ix->ev[0], st));
    int rc;
    static int minb = getenv("MEME_SEED_MINB") ? atoi(getenv("MEME_SEED_MINB")) : 0;
    switch (seed_group() * 10 + minb) {
    case 320: case 323: rc = launch_one<32, 64, 3>(ix, A); break;
    case 324: rc = launch_one<32, 64, 4>(ix, A); break;
    case 160: case 163: rc = launch_one<16, 64, 3>(ix, A); break;
    case 164: rc = launch_one<16, 64, 4>(ix, A); break;
    case 162: rc = launch_one<16, 64, 2>(ix, A); break;
    case 83: rc = launch_one<8, 48, 3>(ix, A); break;
    case 84: rc = launch_one<8, 48, 4>(ix, A); break;
    default: rc = launch_one<8, 48, 2>(ix, A); break;
    }
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ix->ev[1], st));
    CUDA_TRY(cudaMemcpyAsync(ix->h_ctrl, ix->d_ctrl, 16 * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    u64 n_redo = ix->h_ctrl[CT_REDO];
    if (n_redo) {
        // reads whose SMEM list outgrew the group's shared-memory slot: one warp per read, long list
        A.worklist = ix->d_redo; A.n_work = (int64_t)n_redo; A.redo = nullptr;
        CUDA_TRY(cudaMemsetAsync(ix->d_ctrl + CT_WORK, 0, 8, st));
        if ((rc = launch_one<32, MAXS, 1>(ix, A))) return rc;
        CUDA_TRY(cudaEventRecord(ix->ev[1], st));
        CUDA_TRY(cudaMemcpyAsync(ix->h_ctrl, ix->d_ctrl, 16 * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        ix->h_ctrl[CT_REDO] = n_redo;
    }
    return MEME_OK;
}

extern "C" int meme_seed_run(meme_index_t *ix, const meme_seed_opt_t *opt)
{
    if (!ix || !opt) { meme_set_error("null argument"); return MEME_E_ARG; }
    if (opt->min_seed_len < 1 || opt->rounds < 1) { meme_set_error("bad seeding options"); return MEME_E_ARG; }
    CUDA_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    int64_t n = ix->n_reads;
    int rc;
    ix->last_opt = *opt;
    if (n == 0) { ix->ran = 1; return MEME_OK; }
    {   // MEME_SEED_IMPL=coop selects the cooperative (group-per-read) kernel; default is the thread-per-task
        // flat state machine (meme_seed_flat.cuh); "tpr" is its coroutine formulation
        static int impl = -1;
        if (impl < 0) { const char *e = getenv("MEME_SEED_IMPL"); impl = (e && !strcmp(e, "coop")) ? 0 : 1; }
        if (impl == 1) return tpr_seed_run(ix, opt);
    }
    if (!ix->d_pool) {
        size_t want = (size_t)n * 24 + 4096;
        if ((rc = meme_grow_device(&ix->d_pool, &ix->cap_pool, want, sizeof(RawSmem), st))) return rc;
    }
    for (int attempt = 0;; attempt++) {
        if ((rc = seed_launch(ix, opt))) return rc;
        if (ix->h_ctrl[CT_OVF]) {
            meme_set_error("%llu read(s) produced more than %d SMEMs", (unsigned long long)ix->h_ctrl[CT_OVF], MAXS);
            return MEME_E_OVERFLOW;
        }
        // the pool is bump-allocated; if it overflowed, grow it to the exact need and rerun
        if (!ix->h_ctrl[CT_POOL_OVF]) break;
        if (attempt >= 2) { meme_set_error("SMEM pool overflow persists"); return MEME_E_NOMEM; }
        if ((rc = meme_grow_device(&ix->d_pool, &ix->cap_pool, (size_t)ix->h_ctrl[CT_POOL] + 1024, sizeof(RawSmem), st))) return rc;
    }
    // offsets
    CUDA_TRY(cudaEventRecord(ix->ev[2], st));
    k_counts<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(ix->d_rd_nsmem, ix->d_rd_nhit, n, ix->d_smem_off, ix->d_hit_off);
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, ix->d_smem_off, ix->d_smem_off, (int)(n + 1), st);
    if ((rc = meme_grow_device(&ix->d_scan_tmp, &ix->cap_scan_tmp, tmp + 256, 1, st))) return rc;
    size_t t2 = ix->cap_scan_tmp;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(ix->d_scan_tmp, t2, ix->d_smem_off, ix->d_smem_off, (int)(n + 1), st));
    t2 = ix->cap_scan_tmp;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(ix->d_scan_tmp, t2, ix->d_hit_off, ix->d_hit_off, (int)(n + 1), st));
    CUDA_TRY(cudaEventRecord(ix->ev[3], st));
    ix->launches += 3;
    int64_t tot[2];
    CUDA_TRY(cudaMemcpyAsync(&tot[0], ix->d_smem_off + n, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&tot[1], ix->d_hit_off + n, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if ((rc = meme_grow_device((void **)&ix->d_smems, &ix->cap_smems, (size_t)tot[0] + 1, sizeof(meme_smem_t), st))) return rc;
    if ((rc = meme_grow_device((void **)&ix->d_hits, &ix->cap_hits, (size_t)tot[1] + 1, 8, st))) return rc;
    int64_t threads = n * 32;
    CUDA_TRY(cudaEventRecord(ix->ev[4], st));
    k_gather<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(ix->v, (const RawSmem *)ix->d_pool, ix->d_rd_pool_off,
                                                              ix->d_smem_off, ix->d_hit_off, n, opt->max_occ,
                                                              ix->d_smems, ix->d_hits);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ix->ev[5], st));
    ix->ev_ready = 1;
    ix->launches++;
    ix->h_ctrl[32] = (u64)tot[0];
    ix->h_ctrl[33] = (u64)tot[1];
    ix->ran = 1;
    return MEME_OK;
}

extern "C" int meme_seed_totals(meme_index_t *ix, int64_t *n_smems, int64_t *n_hits)
{
    if (!ix || !ix->ran) { meme_set_error("no completed seeding run"); return MEME_E_ARG; }
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    if (n_smems) *n_smems = ix->n_reads ? (int64_t)ix->h_ctrl[32] : 0;
    if (n_hits) *n_hits = ix->n_reads ? (int64_t)ix->h_ctrl[33] : 0;
    return MEME_OK;
}

extern "C" int meme_seed_last_ms(meme_index_t *ix, float ms[3])
{
    if (!ix || !ms || !ix->ev_ready) { meme_set_error("no timed seeding run"); return MEME_E_ARG; }
    CUDA_TRY(cudaEventSynchronize(ix->ev[5]));
    CUDA_TRY(cudaEventElapsedTime(&ms[0], ix->ev[0], ix->ev[1]));
    CUDA_TRY(cudaEventElapsedTime(&ms[1], ix->ev[2], ix->ev[3]));
    CUDA_TRY(cudaEventElapsedTime(&ms[2], ix->ev[4], ix->ev[5]));
    return MEME_OK;
}

extern "C" int meme_seed_counters(meme_index_t *ix, uint64_t counters[8])
{
    if (!ix || !counters) return MEME_E_ARG;
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    counters[0] = ix->h_ctrl[CT_SEARCH]; counters[1] = ix->h_ctrl[CT_BLOCK]; counters[2] = ix->h_ctrl[CT_KARY];
    counters[3] = ix->h_ctrl[CT_EXT]; counters[4] = ix->h_ctrl[CT_OVF]; counters[5] = ix->launches;
    counters[6] = ix->h_ctrl[CT_REDO]; counters[7] = (u64)seed_group();
    if (ix->h_ctrl[41]) { counters[4] = ix->h_ctrl[40]; counters[7] = ix->h_ctrl[41]; }   // flat/tpr: PAC rounds, total round trips
    return MEME_OK;
}

extern "C" int meme_seed_download(meme_index_t *ix, meme_seed_result_t *out)
{
    if (!ix || !out || !ix->ran) { meme_set_error("no completed seeding run"); return MEME_E_ARG; }
    CUDA_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    int64_t n = ix->n_reads;
    int64_t ns = n ? (int64_t)ix->h_ctrl[32] : 0, nh = n ? (int64_t)ix->h_ctrl[33] : 0;
    int rc;
    if (ix->piped) {                       // the pipelined batch already assembled the lists on the host
        out->n_reads = ix->piped_reads; out->n_smems = ns; out->n_hits = nh;
        out->smems = ix->h_smems; out->smem_off = ix->h_smem_off; out->hit_off = ix->h_hit_off; out->hits = (const uint64_t *)ix->h_hits;
        return MEME_OK;
    }
    if ((rc = meme_grow_pinned((void **)&ix->h_smems, &ix->hcap_smems, (size_t)ns + 1, sizeof(meme_smem_t)))) return rc;
    if ((rc = meme_grow_pinned((void **)&ix->h_hits, &ix->hcap_hits, (size_t)nh + 1, 8))) return rc;
    size_t c0 = ix->hcap_off;
    if ((rc = meme_grow_pinned((void **)&ix->h_smem_off, &c0, (size_t)n + 1, 8))) return rc;
    c0 = ix->hcap_off;
    if ((rc = meme_grow_pinned((void **)&ix->h_hit_off, &c0, (size_t)n + 1, 8))) return rc;
    ix->hcap_off = c0;
    if (n) {
        if (ns) CUDA_TRY(cudaMemcpyAsync(ix->h_smems, ix->d_smems, (size_t)ns * sizeof(meme_smem_t), cudaMemcpyDeviceToHost, st));
        if (nh) CUDA_TRY(cudaMemcpyAsync(ix->h_hits, ix->d_hits, (size_t)nh * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(ix->h_smem_off, ix->d_smem_off, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(ix->h_hit_off, ix->d_hit_off, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    } else {
        ix->h_smem_off[0] = 0; ix->h_hit_off[0] = 0;
    }
    out->n_reads = n; out->n_smems = ns; out->n_hits = nh;
    out->smems = ix->h_smems; out->smem_off = ix->h_smem_off; out->hit_off = ix->h_hit_off; out->hits = (const uint64_t *)ix->h_hits;
    return MEME_OK;
}

// ---------------------------------------------------------------------------------------------
// Pipelined batch: the chunk (the reference's -K batch, src/fastmap.cpp:L743) is cut into sub-batches of
// MEME_SEED_CHUNK reads; while sub-batch c runs, the lists of c-1 drain to the host and the reads of c+1 arrive,
// on a second stream with a second set of input/output buffers.  Results land in one set of pinned arrays,
// exactly as the one-shot path delivers them.
// ---------------------------------------------------------------------------------------------
__global__ void k_add_base(int64_t *__restrict__ a, int64_t n, int64_t base)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += base;
}

static void seed_swap_sets(meme_index_t *ix)
{
    SeedAlt &a = ix->alt;
    std::swap(ix->d_codes, a.d_codes); std::swap(ix->cap_codes, a.cap_codes);
    std::swap(ix->d_read_off, a.d_read_off); std::swap(ix->d_smem_off, a.d_smem_off); std::swap(ix->d_hit_off, a.d_hit_off);
    std::swap(ix->d_smems, a.d_smems); std::swap(ix->cap_smems, a.cap_smems);
    std::swap(ix->d_hits, a.d_hits); std::swap(ix->cap_hits, a.cap_hits);
}

// grow a pinned array, keeping its first `keep` elements
static int grow_pinned_keep(void **p, size_t *cap, size_t need, size_t elem, size_t keep)
{
    if (need <= *cap && *p) return MEME_OK;
    size_t ncap = need + need / 4 + 64;
    void *q = nullptr;
    cudaError_t e = cudaMallocHost(&q, ncap * elem);
    if (e != cudaSuccess) { meme_set_error("cudaMallocHost(%zu bytes): %s", ncap * elem, cudaGetErrorString(e)); return MEME_E_NOMEM; }
    if (*p) { if (keep) memcpy(q, *p, keep * elem); cudaFreeHost(*p); }
    *p = q; *cap = ncap;
    return MEME_OK;
}

static int64_t seed_chunk_reads()
{
    const char *e = getenv("MEME_SEED_CHUNK");
    long long v = e ? atoll(e) : (1ll << 20);
    return v < 1024 ? 1024 : (int64_t)v;
}

static int seed_batch_pipelined(meme_index_t *ix, const uint8_t *codes, const int64_t *read_off, int64_t n,
                                const meme_seed_opt_t *opt, meme_seed_result_t *out, int64_t chunk)
{
    CUDA_TRY(cudaSetDevice(ix->device));
    if (read_off[0] != 0) { meme_set_error("read_off[0] must be 0"); return MEME_E_ARG; }
    for (int64_t i = 0; i < n; i++) {
        int64_t len = read_off[i + 1] - read_off[i];
        if (len < 0 || len > MEME_MAX_READ_LEN) {
            meme_set_error("read %lld has length %lld (limit %d; the reference asserts <= 301)", (long long)i, (long long)len, MEME_MAX_READ_LEN);
            return MEME_E_ARG;
        }
    }
    const int64_t K = (n + chunk - 1) / chunk;
    int64_t max_codes = 0;
    for (int64_t c = 0; c < K; c++) {
        int64_t r0 = c * chunk, r1 = r0 + chunk < n ? r0 + chunk : n;
        max_codes = std::max(max_codes, read_off[r1] - read_off[r0]);
    }
    int rc;
    cudaStream_t st = ix->stream;
    if ((rc = seed_alloc_inputs(ix, chunk, max_codes))) return rc;
    SeedAlt &A = ix->alt;
    if (A.cap_reads != ix->cap_reads) {                 // second set of per-read offset arrays, same capacity as the first
        CUDA_TRY(cudaDeviceSynchronize());
        void **bufs[] = {(void **)&A.d_read_off, (void **)&A.d_smem_off, (void **)&A.d_hit_off};
        for (void **b : bufs) {
            if (*b) { CUDA_TRY(cudaFree(*b)); *b = nullptr; }
            size_t dummy = 0;
            if ((rc = meme_grow_device(b, &dummy, ix->cap_reads, 8, st))) return rc;
        }
        A.cap_reads = ix->cap_reads;
    }
    if ((rc = meme_grow_device((void **)&A.d_codes, &A.cap_codes, (size_t)max_codes + 64, 1, st))) return rc;
    if (!ix->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ix->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 6; k++) CUDA_TRY(cudaEventCreateWithFlags(&ix->pe[k], cudaEventDisableTiming));
    }
    cudaStream_t cs = ix->copy_stream;
    cudaEvent_t *e_h2d = ix->pe, *e_run = ix->pe + 2, *e_d2h = ix->pe + 4;      // indexed by set (chunk & 1)
    {
        size_t c0 = ix->hcap_off;
        if ((rc = meme_grow_pinned((void **)&ix->h_smem_off, &c0, (size_t)n + 1, 8))) return rc;
        c0 = ix->hcap_off;
        if ((rc = meme_grow_pinned((void **)&ix->h_hit_off, &c0, (size_t)n + 1, 8))) return rc;
        ix->hcap_off = c0;
    }
    struct Part { int64_t r0, n, tot_s, tot_h, base_s, base_h; } part[2] = {};
    int64_t base_s = 0, base_h = 0;
    u64 acc[7] = {0, 0, 0, 0, 0, 0, 0};
    // input of sub-batch c -> the buffers named by the arguments (the set that is NOT installed while c-1 runs)
    auto h2d = [&](int64_t c, uint8_t *d_codes, int64_t *d_read_off) -> int {
        int64_t r0 = c * chunk, r1 = r0 + chunk < n ? r0 + chunk : n, nr = r1 - r0;
        int64_t nc = read_off[r1] - read_off[r0];
        if (nc) CUDA_TRY(cudaMemcpyAsync(d_codes, codes + read_off[r0], (size_t)nc, cudaMemcpyHostToDevice, cs));
        CUDA_TRY(cudaMemcpyAsync(d_read_off, read_off + r0, (size_t)(nr + 1) * 8, cudaMemcpyHostToDevice, cs));
        if (read_off[r0]) k_add_base<<<(unsigned)((nr + 1 + 255) / 256), 256, 0, cs>>>(d_read_off, nr + 1, -read_off[r0]);
        CUDA_TRY(cudaEventRecord(e_h2d[c & 1], cs));
        return MEME_OK;
    };
    // lists of a finished sub-batch -> the pinned arrays, offsets rebased to the whole batch
    auto d2h = [&](const Part &p, int set, const meme_smem_t *d_smems, const u64 *d_hits, int64_t *d_smem_off, int64_t *d_hit_off) -> int {
        int r;
        if ((size_t)(p.base_s + p.tot_s) + 1 > ix->hcap_smems || (size_t)(p.base_h + p.tot_h) + 1 > ix->hcap_hits || !ix->h_smems || !ix->h_hits) {
            CUDA_TRY(cudaStreamSynchronize(cs));        // earlier copies target the old arrays
            // first call: size for the whole batch from the density seen so far
            double scale = (double)n / (double)(p.r0 + p.n) * 1.1;
            size_t want_s = (size_t)((double)(p.base_s + p.tot_s) * scale) + 1024, want_h = (size_t)((double)(p.base_h + p.tot_h) * scale) + 1024;
            if ((r = grow_pinned_keep((void **)&ix->h_smems, &ix->hcap_smems, want_s, sizeof(meme_smem_t), (size_t)p.base_s))) return r;
            if ((r = grow_pinned_keep((void **)&ix->h_hits, &ix->hcap_hits, want_h, 8, (size_t)p.base_h))) return r;
        }
        CUDA_TRY(cudaStreamWaitEvent(cs, e_run[set], 0));
        unsigned g = (unsigned)((p.n + 1 + 255) / 256);
        if (p.base_s) k_add_base<<<g, 256, 0, cs>>>(d_smem_off, p.n + 1, p.base_s);
        if (p.base_h) k_add_base<<<g, 256, 0, cs>>>(d_hit_off, p.n + 1, p.base_h);
        if (p.tot_s) CUDA_TRY(cudaMemcpyAsync(ix->h_smems + p.base_s, d_smems, (size_t)p.tot_s * sizeof(meme_smem_t), cudaMemcpyDeviceToHost, cs));
        if (p.tot_h) CUDA_TRY(cudaMemcpyAsync(ix->h_hits + p.base_h, d_hits, (size_t)p.tot_h * 8, cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaMemcpyAsync(ix->h_smem_off + p.r0, d_smem_off, (size_t)(p.n + 1) * 8, cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaMemcpyAsync(ix->h_hit_off + p.r0, d_hit_off, (size_t)(p.n + 1) * 8, cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaEventRecord(e_d2h[set], cs));
        return MEME_OK;
    };
    CUDA_TRY(cudaStreamSynchronize(st));                // the handle's buffers are idle from here on
    if ((rc = h2d(0, ix->d_codes, ix->d_read_off))) return rc;
    for (int64_t c = 0; c < K; c++) {
        const int set = (int)(c & 1);
        if (c >= 1) {
            seed_swap_sets(ix);                          // installed: set of c; ix->alt: set of c-1 (= set of c+1)
            if ((rc = d2h(part[set ^ 1], set ^ 1, A.d_smems, A.d_hits, A.d_smem_off, A.d_hit_off))) return rc;
        }
        if (c + 1 < K && (rc = h2d(c + 1, A.d_codes, A.d_read_off))) return rc;   // after the d2h wait on e_run[c-1]
        int64_t r0 = c * chunk, r1 = r0 + chunk < n ? r0 + chunk : n;
        CUDA_TRY(cudaStreamWaitEvent(st, e_h2d[set], 0));
        if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(st, e_d2h[set], 0));             // the lists of c-2 have left this set
        ix->n_reads = r1 - r0; ix->n_codes = read_off[r1] - read_off[r0]; ix->ran = 0;
        if ((rc = meme_seed_run(ix, opt))) return rc;
        CUDA_TRY(cudaEventRecord(e_run[set], st));
        Part &p = part[set];
        p.r0 = r0; p.n = r1 - r0; p.tot_s = (int64_t)ix->h_ctrl[32]; p.tot_h = (int64_t)ix->h_ctrl[33];
        p.base_s = base_s; p.base_h = base_h;
        base_s += p.tot_s; base_h += p.tot_h;
        const int slot[7] = {8, 9, 10, 11, 4, 40, 41};
        for (int k = 0; k < 7; k++) acc[k] += ix->h_ctrl[slot[k]];
    }
    {
        const int set = (int)((K - 1) & 1);
        if ((rc = d2h(part[set], set, ix->d_smems, ix->d_hits, ix->d_smem_off, ix->d_hit_off))) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(cs));
    CUDA_TRY(cudaStreamSynchronize(st));
    {
        const int slot[7] = {8, 9, 10, 11, 4, 40, 41};
        for (int k = 0; k < 7; k++) ix->h_ctrl[slot[k]] = acc[k];
    }
    ix->h_ctrl[32] = (u64)base_s; ix->h_ctrl[33] = (u64)base_h;
    ix->piped = 1; ix->piped_reads = n; ix->ran = 1;
    out->n_reads = n; out->n_smems = base_s; out->n_hits = base_h;
    out->smems = ix->h_smems; out->smem_off = ix->h_smem_off; out->hit_off = ix->h_hit_off; out->hits = (const uint64_t *)ix->h_hits;
    return MEME_OK;
}

extern "C" int meme_seed_batch(meme_index_t *ix, const uint8_t *codes, const int64_t *read_off, int64_t n_reads,
                               const meme_seed_opt_t *opt, meme_seed_result_t *out)
{
    int rc;
    if (!ix || !out || !opt || !read_off || n_reads < 0 || (n_reads > 0 && !codes)) { meme_set_error("bad argument"); return MEME_E_ARG; }
    const int64_t chunk = seed_chunk_reads();
    if (n_reads > chunk + chunk / 2) return seed_batch_pipelined(ix, codes, read_off, n_reads, opt, out, chunk);
    if ((rc = meme_seed_upload(ix, codes, read_off, n_reads))) return rc;
    if ((rc = meme_seed_run(ix, opt))) return rc;
    return meme_seed_download(ix, out);
}
