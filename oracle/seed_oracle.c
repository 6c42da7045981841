/* oracle/seed_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's P-RMI learned-index SMEM seeding
 * (kernel A of SURVEY.md §8a).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this; the product path never does.
 *
 * Parity pinning: tests/test_oracle_seed.py compares this file's output with SMEM
 * dumps of the UNMODIFIED reference (oracle/_ref/learned_seeding_mode3.*, i.e.
 * test/Learned_seeding_big_read.cpp run with steps=4) -- committed as golden fixtures
 * under tests/golden/ and regenerated live whenever oracle/_ref exists.
 *
 * What is restated, and where it lives in the reference (paths under /root/reference):
 *   rmi_lookup()        src/LearnedIndex_seeding.cpp:L186-210   learned_index_lookup
 *   make_key()          src/LearnedIndex_seeding.cpp:L795-901   Tokenization (portable version)
 *   cmp_suffix()        src/LearnedIndex_seeding.cpp:L226-601   compare_read_and_ref_binary{,_left}*
 *   locate_window()     src/LearnedIndex_seeding.cpp:L2262-2357 bounded binary search + linear fix-up
 *   locate_from_isa()   src/LearnedIndex_seeding.cpp:L3236-3330 exponential search around an ISA hint
 *   level_search()      src/LearnedIndex_seeding.cpp:L2359-2662 (right_smem_search), L2667-3204
 *                       (mem_search), L3208-3702, L3704-4189 (the *_tradeoff twins)
 *   seed_step1()        src/LearnedIndex_seeding.cpp:L1691-1894 Learned_getSMEMsOnePosOneThread_step1
 *   seed_onepos()       src/LearnedIndex_seeding.cpp:L1897-2126 Learned_getSMEMsOnePosOneThread
 *   seed_allpos()       src/LearnedIndex_seeding.cpp:L913-972   Learned_getSMEMsAllPosOneThread
 *   seed_round3()       src/LearnedIndex_seeding.cpp:L974-1283, L1284-1466 Learned_bwtSeedStrategyAllPosOneThread*
 *   oracle_seed_reads() src/bwamem.cpp:L1249-1397               mem_kernel1_core_Learned up to the sort
 *
 * The reference holds four near-identical copies of the search (right/left x learned/ISA start);
 * their outputs are canonical (they depend only on the suffix array, not on where the search
 * starts), so they are stated once here.  The last-mile probe sequence follows the reference
 * (window binary search, +-1 linear fix-up); the SA-interval growth follows the reference's plain
 * linear variant (EXPONENTIAL_SMEMSEARCH==0 branch, L2570-2631), which yields the same interval
 * as the default exponential variant.
 *
 * Work counters are accumulated per call so that bench.py can state the
 * algorithmic bytes per read:  Lk, Lp, P, W, I, H, S.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>

typedef struct {
    uint64_t sa_num;
    uint8_t *text;      /* .0123: fwd + RC, one code (0..3) per byte           */
    uint64_t *sa;       /* suffix array positions, unpacked from .pos_packed   */
    uint64_t *isa;      /* inverse suffix array (mode 3) or NULL               */
    uint8_t *l1, *l2;   /* 24-byte records {f64 alpha, f64 beta, u64 err}      */
    uint64_t n_l1, n_l2;
    int bit_shift;
    int mode;           /* 1, 2 or 3                          */
} oracle_index_t;

typedef struct {
    int32_t start, end;     /* [start,end) in the read                        */
    int32_t hitbeg;         /* index into this read's hit list                */
    int32_t hitcount;
    uint64_t sa_lo;         /* first SA index of the interval (extra info)     */
} oracle_smem_t;            /* same 24-byte shape as the reference's mem_tl    */

typedef struct {
    int min_seed_len;       /* opt->min_seed_len   (19)  */
    int split_len;          /* (int)(min_seed_len*split_factor+.499) (28) */
    int split_width;        /* opt->split_width    (10)  */
    int max_mem_intv;       /* opt->max_mem_intv   (20)  */
    int rounds;             /* 1: step1 only, 2: + reseed, 3: + third round (harness `steps`) */
} oracle_seed_opt_t;

enum { C_LK, C_LP, C_P, C_W, C_I, C_H, C_S, C_SEARCH, C_N };

typedef struct {
    int64_t n_smems, n_hits;
    oracle_smem_t *smems;
    int64_t *smem_off;      /* nreads+1 */
    int64_t *hit_off;       /* nreads+1 */
    uint64_t *hits;
    uint64_t counters[C_N];
} oracle_seed_result_t;

/* ------------------------------------------------------------------------------------------ */
static void *xread(const char *path, size_t *size)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void *p = malloc(n > 0 ? n : 1);
    if (p && fread(p, 1, n, f) != (size_t)n) { free(p); p = NULL; }
    fclose(f);
    *size = n;
    return p;
}

void oracle_index_free(oracle_index_t *x)
{
    if (!x) return;
    free(x->text); free(x->sa); free(x->isa); free(x->l1); free(x->l2); free(x);
}

/* Loads what memoryAllocLearned (src/fastmap.cpp:L351-638) loads: .0123, .pos_packed,
 * _L1/_L2_PARAMETERS; builds the inverse SA for mode 3 (fastmap.cpp:L586-589). */
oracle_index_t *oracle_index_load(const char *prefix, int mode)
{
    char path[4096];
    size_t sz;
    oracle_index_t *x = (oracle_index_t *)calloc(1, sizeof(*x));
    x->mode = mode;
    snprintf(path, sizeof path, "%s.0123", prefix);
    x->text = (uint8_t *)xread(path, &sz);
    if (!x->text) goto fail;
    x->sa_num = sz;
    snprintf(path, sizeof path, "%s.pos_packed", prefix);
    uint8_t *pp = (uint8_t *)xread(path, &sz);
    if (!pp || sz != x->sa_num * 5) { free(pp); goto fail; }
    x->sa = (uint64_t *)malloc(x->sa_num * 8);
    for (uint64_t i = 0; i < x->sa_num; i++) {
        uint32_t hi;
        memcpy(&hi, pp + i * 5, 4);
        x->sa[i] = ((uint64_t)hi << 8) | pp[i * 5 + 4];
    }
    free(pp);
    if (mode == 3) {
        x->isa = (uint64_t *)malloc(x->sa_num * 8);
        for (uint64_t i = 0; i < x->sa_num; i++) x->isa[x->sa[i]] = i;
    }
    snprintf(path, sizeof path, "%s.suffixarray_uint64_L1_PARAMETERS", prefix);
    x->l1 = (uint8_t *)xread(path, &sz);
    if (!x->l1) goto fail;
    x->n_l1 = sz / 24;
    snprintf(path, sizeof path, "%s.suffixarray_uint64_L2_PARAMETERS", prefix);
    x->l2 = (uint8_t *)xread(path, &sz);
    if (!x->l2) goto fail;
    x->n_l2 = sz / 24;
    x->bit_shift = 64 - (__builtin_ffsll((long long)x->n_l2) - 1); /* LearnedIndex_seeding.cpp:L113-119 */
    return x;
fail:
    oracle_index_free(x);
    return NULL;
}

uint64_t oracle_index_sa_num(const oracle_index_t *x) { return x->sa_num; }

static uint64_t rmi_lookup_raw(const oracle_index_t *x, uint64_t key, uint64_t *err);

/* learned_index_lookup for a batch of keys (tests pin the device inference against it) */
void oracle_rmi_lookup(const oracle_index_t *x, const uint64_t *keys, int64_t n, uint64_t *pred, uint64_t *err)
{
    for (int64_t i = 0; i < n; i++) pred[i] = rmi_lookup_raw(x, keys[i], &err[i]);
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const oracle_index_t *ix;
    const uint8_t *fw;      /* read codes 0..4                       */
    uint8_t *rc;            /* reverse complement, 4 stays 4         */
    int l_seq;
    int min_seed_len;
    int min_intv;
    int full_match;         /* raux->max_l_seq == l_seq  (L1870-1875) */
    uint64_t max_refpos;
    /* reseed cache (L934-939) */
    int use_cached, cache_pivot, cache_pivot_end;
    uint64_t cache_refpos;
    /* output */
    oracle_smem_t *sm; int n_sm, m_sm;
    uint64_t *hits; int64_t n_hits, m_hits;
    uint64_t *cnt;
} rd_t;

static inline size_t fclamp(double v, double bound)
{
    if (v < 0.0) return 0;
    return v > bound ? (size_t)bound : (size_t)v;
}

/* learned_index_lookup, L186-210 */
static uint64_t rmi_lookup_raw(const oracle_index_t *x, uint64_t key, uint64_t *err)
{
    size_t m = (size_t)(key >> x->bit_shift);
    double a, b;
    memcpy(&a, x->l2 + m * 24, 8);
    memcpy(&b, x->l2 + m * 24 + 8, 8);
    memcpy(err, x->l2 + m * 24 + 16, 8);
    double p = fma(b, (double)key, a);
    if (*err >> 63) {
        size_t st = (size_t)((*err >> 32) & 0x7fffffff);
        double num = (double)(*err & 0xffffffffu);
        m = st + fclamp(p, num - 1);
        memcpy(&a, x->l1 + m * 24, 8);
        memcpy(&b, x->l1 + m * 24 + 8, 8);
        memcpy(err, x->l1 + m * 24 + 16, 8);
        p = fma(b, (double)key, a);
    }
    return fclamp(p, (double)x->sa_num - 1.0);
}

static uint64_t rmi_lookup(rd_t *r, uint64_t key, uint64_t *err)
{
    const oracle_index_t *x = r->ix;
    size_t m = (size_t)(key >> x->bit_shift);
    double a, b;
    memcpy(&a, x->l2 + m * 24, 8);
    memcpy(&b, x->l2 + m * 24 + 8, 8);
    memcpy(err, x->l2 + m * 24 + 16, 8);
    double p = fma(b, (double)key, a);
    r->cnt[C_LK]++;
    if (*err >> 63) {
        size_t st = (size_t)((*err >> 32) & 0x7fffffff);
        double num = (double)(*err & 0xffffffffu);
        m = st + fclamp(p, num - 1);
        memcpy(&a, x->l1 + m * 24, 8);
        memcpy(&b, x->l1 + m * 24 + 8, 8);
        memcpy(err, x->l1 + m * 24 + 16, 8);
        p = fma(b, (double)key, a);
        r->cnt[C_LP]++;
    }
    return fclamp(p, (double)x->sa_num - 1.0);
}

/* Tokenization (portable), L795-901: 32-base key, first base in the MSBs, padded with 3;
 * *amb = first N at/after the start (or l_seq). q points at the strand buffer start. */
static uint64_t make_key(const uint8_t *q, int from, int l_seq, int *amb)
{
    uint64_t key = 0;
    int r, len = l_seq - from;
    *amb = l_seq;
    for (r = 0; r < len && r < 32; r++) {
        if (q[from + r] >= 4) { *amb = from + r; break; }
        key = (key << 2) | q[from + r];
    }
    for (; r < 32; r++) key = (key << 2) | 3;
    if (*amb == l_seq)
        for (r = 32; r < len; r++)
            if (q[from + r] >= 4) { *amb = from + r; break; }
    return key;
}

/* compare_read_and_ref_binary*, L226-601.  q = first base of the query on its strand.
 * Returns lcp(i) = min(LCP, valid_len, sa_num - SA[i]); *less = "reference suffix < read"
 * (the function's return value), *exact = exact_match_flag. */
static inline int cmp_suffix(rd_t *r, uint64_t i, const uint8_t *q, int valid_len, int *less, int *exact)
{
    const oracle_index_t *x = r->ix;
    uint64_t pos = x->sa[i];
    uint64_t ref_len = x->sa_num - pos;
    int read_len = ref_len < (uint64_t)valid_len ? (int)ref_len : valid_len;
    const uint8_t *t = x->text + pos;
    int k = 0;
    while (k < read_len && t[k] == q[k]) k++;
    r->cnt[C_P]++;
    {   /* pac words touched beyond the inline 32-base key (mode 2/3) or all words (mode 1) */
        int examined = k < read_len ? k + 1 : read_len;
        int base = x->mode == 1 ? 0 : 32;
        if (examined > base) r->cnt[C_W] += (uint64_t)((examined - base + 31) / 32);
    }
    *exact = 0;
    if (k < read_len) { *less = t[k] < q[k]; return k; }
    if ((uint64_t)read_len < ref_len) { *exact = 1; *less = 1; return read_len; }
    *less = 0;          /* text end: the padded-T suffix counts as larger (L269-273) */
    return (int)ref_len;
}

/* Bounded binary search in the RMI window + linear fix-up, L2262-2357.
 * Returns a position with the maximal lcp; *mlen = that lcp. */
static uint64_t locate_window(rd_t *r, const uint8_t *q, int V, uint64_t est, uint64_t enc_err, int *mlen)
{
    uint64_t n_sa = r->ix->sa_num;
    uint64_t cerr = (enc_err >> 32) & 0x3fffffff, err = enc_err & 0x7fffffff;
    uint64_t ub = est + err >= n_sa - 1 ? n_sa - 2 : est + err;
    uint64_t lb = est > cerr ? est - cerr : 1;
    if (lb >= ub) lb = ub > 0 ? ub - 1 : 0;   /* (the reference underflows here when ub == 0: degenerate parameter files) */
    uint64_t n = ub - lb + 1, mid = est, half;
    int less, exact = 0, ml = 0, last;
    while ((half = n >> 1)) {
        mid = lb + half;
        ml = cmp_suffix(r, mid, q, V, &less, &exact);
        if (less) lb = mid;
        if (exact) break;
        n -= half;
    }
    uint64_t it;
    if (exact) { *mlen = ml; return lb; }
    if (mid != lb) {
        last = ml; it = lb;
        for (;;) {
            ml = cmp_suffix(r, it, q, V, &less, &exact);
            if (less || exact) break;
            if (it == 0) break;
            it--; last = ml;
        }
        if (last > ml) { it++; ml = last; }
    } else {
        last = ml; it = lb + 1 >= n_sa ? n_sa - 1 : lb + 1;
        for (;;) {
            ml = cmp_suffix(r, it, q, V, &less, &exact);
            if (!less || exact) break;
            if (it == n_sa - 1) break;
            it++; last = ml;
        }
        if (last > ml) { it--; ml = last; }
    }
    *mlen = ml;
    return it;
}

/* Exponential search around an inverse-SA hint (start 2, x5), then the same binary search and
 * fix-up: mem_search_tradeoff L3236-3330 / right_smem_search_tradeoff L3733-3830. */
static uint64_t locate_from_isa(rd_t *r, const uint8_t *q, int V, uint64_t est, int *mlen)
{
    uint64_t n_sa = r->ix->sa_num, mv = 2, it = est, lb, ub, n, mid, half;
    int less, exact = 0, ml, last;
    if (it < 2) it = 2;
    if (it > n_sa - 1 - 2) it = n_sa - 1 - 2;
    ml = cmp_suffix(r, it, q, V, &less, &exact);
    if (exact) { *mlen = ml; return it; }
    if (!less) {
        it -= mv;
        for (;;) {
            ml = cmp_suffix(r, it, q, V, &less, &exact);
            if (less || exact) break;
            if (it == 0) break;
            mv *= 5;
            if (it > mv) it -= mv; else { mv = it; it = 0; }
        }
        if (exact) { *mlen = ml; return it; }
        ub = it + mv >= n_sa - 1 ? n_sa - 2 : it + mv;
        lb = it > 0 ? it : 1;
    } else {
        it += mv;
        for (;;) {
            ml = cmp_suffix(r, it, q, V, &less, &exact);
            if (!less || exact) break;
            if (it == n_sa - 1) break;
            mv *= 5;
            if (n_sa - 1 > it + mv) it += mv; else { mv = n_sa - it - 1; it = n_sa - 1; }
        }
        if (exact) { *mlen = ml; return it; }
        ub = it >= n_sa - 1 ? n_sa - 2 : it;
        lb = it > mv ? it - mv : 1;
    }
    if (lb > ub) lb = ub;
    n = ub - lb + 1; mid = it;
    while ((half = n >> 1)) {
        mid = lb + half;
        ml = cmp_suffix(r, mid, q, V, &less, &exact);
        if (less) lb = mid;
        if (exact) break;
        n -= half;
    }
    if (exact) { *mlen = ml; return lb; }
    if (mid != lb) {
        last = ml; it = lb;
        for (;;) {
            ml = cmp_suffix(r, it, q, V, &less, &exact);
            if (less || exact) break;
            if (it == 0) break;
            it--; last = ml;
        }
        if (last > ml) { it++; ml = last; }
    } else {
        last = ml; it = lb + 1 >= n_sa ? n_sa - 1 : lb + 1;
        for (;;) {
            ml = cmp_suffix(r, it, q, V, &less, &exact);
            if (!less || exact) break;
            if (it == n_sa - 1) break;
            it++; last = ml;
        }
        if (last > ml) { it--; ml = last; }
    }
    *mlen = ml;
    return it;
}

static void push_smem(rd_t *r, int start, int end, uint64_t lo, uint64_t cnt)
{
    if (r->n_sm == r->m_sm) {
        r->m_sm = r->m_sm ? r->m_sm * 2 : 64;
        r->sm = (oracle_smem_t *)realloc(r->sm, (size_t)r->m_sm * sizeof(oracle_smem_t));
    }
    if (r->n_hits + (int64_t)cnt > r->m_hits) {
        while (r->n_hits + (int64_t)cnt > r->m_hits) r->m_hits = r->m_hits ? r->m_hits * 2 : 1024;
        r->hits = (uint64_t *)realloc(r->hits, (size_t)r->m_hits * 8);
    }
    oracle_smem_t *m = &r->sm[r->n_sm++];
    m->start = start; m->end = end; m->hitbeg = (int32_t)r->n_hits; m->hitcount = (int32_t)cnt; m->sa_lo = lo;
    for (uint64_t i = 0; i < cnt; i++) r->hits[r->n_hits++] = r->ix->sa[lo + i];
    r->cnt[C_S]++;
    r->cnt[C_H] += cnt;
}

/* From a best position `it` with lcp `ml`: grow the SA interval of suffixes sharing `ml` bases; while
 * it holds fewer than min_intv suffixes, drop ml to the larger neighbour lcp and regrow
 * (L2359-2631).  Returns the level; [*lo, *lo + *cnt) is its interval. */
static int grow_level(rd_t *r, const uint8_t *q, uint64_t it, int ml, uint64_t *lo_out, uint64_t *cnt_out)
{
    uint64_t n_sa = r->ix->sa_num, lo = it, hi = it;
    int less, exact;
    for (;;) {
        int lm = 0, um = 0;
        while (lo > 0) {
            lm = cmp_suffix(r, lo - 1, q, ml, &less, &exact);
            if (lm < ml) break;
            lo--; lm = 0;
        }
        if (lo == 0) lm = 0;
        while (hi < n_sa - 1) {
            um = cmp_suffix(r, hi + 1, q, ml, &less, &exact);
            if (um < ml) break;
            hi++; um = 0;
        }
        if (hi == n_sa - 1) um = 0;
        if (hi - lo + 1 >= (uint64_t)r->min_intv) break;
        ml = lm > um ? lm : um;
    }
    *lo_out = lo; *cnt_out = hi - lo + 1;
    return ml;
}

enum { START_RMI, START_ISA_EXACT, START_ISA_NEAR };

/* One search = right_smem_search / mem_search / their _tradeoff twins.
 * right: query = fw[pivot..]; left: query = rc[l_seq-1-pivot..].  Returns the level. */
static int level_search(rd_t *r, int pivot, int right, int emit, int start_mode, uint64_t isa_pos, uint64_t *pos_out)
{
    const uint8_t *buf = right ? r->fw : r->rc;
    int from = right ? pivot : r->l_seq - 1 - pivot;
    int amb, ml;
    uint64_t key = make_key(buf, from, r->l_seq, &amb);
    int V = amb - from;
    const uint8_t *q = buf + from;
    uint64_t it;
    r->cnt[C_SEARCH]++;
    if (start_mode == START_ISA_EXACT) {          /* no_search: L3241-3243 */
        it = isa_pos; ml = V;
    } else if (start_mode == START_ISA_NEAR) {
        it = locate_from_isa(r, q, V, isa_pos, &ml);
    } else {
        uint64_t enc, est = rmi_lookup(r, key, &enc);
        it = locate_window(r, q, V, est, enc, &ml);
    }
    uint64_t lo = it, cnt = 1;
    /* for min_intv==1 the interval is only needed when an SMEM is emitted, but the level is the same */
    if (r->min_intv > 1 || (emit && ml >= r->min_seed_len))
        ml = grow_level(r, q, it, ml, &lo, &cnt);
    if (emit && ml >= r->min_seed_len) push_smem(r, pivot, pivot + ml, lo, cnt);
    if (pos_out) *pos_out = lo;
    return ml;
}

static inline uint64_t isa_at(rd_t *r, uint64_t refpos)
{
    r->cnt[C_I]++;
    return r->ix->isa[refpos];
}

/* picks the start of a search the way L1741-1771 / L1780-1816 do */
static int do_search(rd_t *r, int pivot, int right, int emit)
{
    uint64_t n_sa = r->ix->sa_num;
    if (r->ix->mode == 3) {
        if (r->full_match) {
            uint64_t rp = right ? r->max_refpos + pivot : n_sa - r->max_refpos - pivot - 1;
            return level_search(r, pivot, right, emit, START_ISA_EXACT, isa_at(r, rp), NULL);
        }
        if (r->use_cached) {
            if (right && r->cache_pivot_end >= pivot + 18 && pivot >= r->cache_pivot)
                return level_search(r, pivot, 1, emit, START_ISA_NEAR,
                                    isa_at(r, r->cache_refpos + pivot - r->cache_pivot), NULL);
            if (!right && pivot < r->cache_pivot_end && r->cache_pivot + 18 <= pivot)
                return level_search(r, pivot, 0, emit, START_ISA_NEAR,
                                    isa_at(r, n_sa + r->cache_pivot - r->cache_refpos - pivot - 1), NULL);
        }
    }
    return level_search(r, pivot, right, emit, START_RMI, 0, NULL);
}

/* Learned_getSMEMsOnePosOneThread_step1, L1691-1894.  Returns the next pivot. */
static int seed_step1(rd_t *r, int pivot)
{
    int L = r->l_seq;
    if (r->fw[pivot] >= 4) return L - pivot < r->min_seed_len ? L : pivot + 1;
    if (pivot != 0 && r->fw[pivot - 1] < 4) {
        int sp = pivot;
        while (sp < L) {
            if (r->fw[sp] >= 4) {
                if (L - sp < r->min_seed_len) { pivot = L; sp = L; }
                else { sp += 1; pivot += 1; }
                continue;
            }
            int l = do_search(r, pivot, 0, 0);
            pivot = pivot - l + 1;
            if (L - pivot < r->min_seed_len) break;
            int rr = do_search(r, pivot, 1, 1);
            /* PROGRESS GUARD (not in the reference): when the text ends with its longest T run, a match into the text
             * end can make the right search shorter than the left one and the reference loops here forever
             * (tests/test_text_end_guard.py); the zig-zag of this segment ends instead. */
            if (pivot + rr <= sp) break;
            sp = pivot + rr;
            pivot = sp;
        }
        return L;
    }
    /* first search of a segment: always through the RMI (L1851-1857) */
    uint64_t pos;
    int len = level_search(r, pivot, 1, 1, START_RMI, 0, &pos);
    if (r->ix->mode == 3 && len == L) {     /* whole read matched: L1870-1875 */
        r->full_match = 1;
        r->max_refpos = r->ix->sa[pos];
    }
    return pivot + len;
}

/* Learned_getSMEMsOnePosOneThread (reseed), L1897-2126 */
static void seed_onepos(rd_t *r, int pivot)
{
    int L = r->l_seq;
    if (r->fw[pivot] >= 4) return;
    if (pivot != 0 && r->fw[pivot - 1] < 4) {
        int len = do_search(r, pivot, 1, 0);
        int next = pivot + len, sp = pivot;
        while (sp < next) {
            int l = do_search(r, pivot, 0, 0);
            pivot = pivot - l + 1;
            if (next - pivot < r->min_seed_len) break;
            int rr = do_search(r, pivot, 1, 1);
            if (pivot + rr <= sp) break;            /* progress guard, as in seed_step1 */
            sp = pivot + rr;
            pivot = sp;
        }
        return;
    }
    (void)L;
    level_search(r, pivot, 1, 1, START_RMI, 0, NULL);
}

/* Learned_getSMEMsAllPosOneThread, L913-972 (+ step1only L904-911 when rounds==1) */
static void seed_allpos(rd_t *r, const oracle_seed_opt_t *o)
{
    int pivot = 0;
    while (pivot < r->l_seq) {
        int before = r->n_sm;
        pivot = seed_step1(r, pivot);
        int after = r->n_sm;
        if (o->rounds < 2) continue;
        for (int k = before; k < after; k++) {
            int qb = r->sm[k].start, qe = r->sm[k].end, hc = r->sm[k].hitcount;
            if (qe - qb < o->split_len || hc > o->split_width) continue;
            int keep = r->min_intv;
            r->min_intv = hc + 1;
            r->use_cached = 0;
            if (r->ix->mode == 3 && hc >= 1 && hc < 10) {
                r->use_cached = 1;
                r->cache_pivot = qb; r->cache_pivot_end = qe;
                r->cache_refpos = r->ix->sa[r->sm[k].sa_lo];
            }
            seed_onepos(r, (qb + qe) >> 1);
            r->use_cached = 0;
            r->min_intv = keep;
        }
    }
}

/* Learned_bwtSeedStrategyAllPosOneThread{,_mem_tradeoff}, L974-1283 / L1284-1466 */
static void seed_round3(rd_t *r)
{
    int L = r->l_seq, msl = r->min_seed_len, pivot = 0;
    uint64_t n_sa = r->ix->sa_num;
    while (pivot < L - msl + 1) {
        if (r->fw[pivot] >= 4) { pivot++; continue; }
        int amb, ml, less, exact;
        uint64_t key = make_key(r->fw, pivot, L, &amb);
        int V = amb - pivot;
        if (V < msl) { pivot += V; continue; }
        const uint8_t *q = r->fw + pivot;
        uint64_t it;
        r->cnt[C_SEARCH]++;
        if (r->full_match) { it = isa_at(r, r->max_refpos + pivot); ml = V; }
        else { uint64_t enc, est = rmi_lookup(r, key, &enc); it = locate_window(r, q, V, est, enc, &ml); }
        if (ml < msl) { pivot += msl; continue; }
        uint64_t lo = it, hi = it, last_lo = it, last_cnt = 0, cnt;
        int emit_len;
        for (;;) {
            int lm = 0, um = 0;
            /* grow at this level; counting may stop early once min_intv is reached (L1224-1236) */
            while (lo > 0 && hi - lo + 1 < (uint64_t)r->min_intv) {
                lm = cmp_suffix(r, lo - 1, q, ml, &less, &exact);
                if (lm < ml) break;
                lo--; lm = 0;
            }
            if (lo == 0) lm = 0;
            while (hi < n_sa - 1 && hi - lo + 1 < (uint64_t)r->min_intv) {
                um = cmp_suffix(r, hi + 1, q, ml, &less, &exact);
                if (um < ml) break;
                hi++; um = 0;
            }
            if (hi == n_sa - 1) um = 0;
            cnt = hi - lo + 1;
            if (cnt >= (uint64_t)r->min_intv) {            /* L1243-1251 */
                cnt = last_cnt; lo = last_lo; emit_len = ml + 1;
                break;
            }
            if ((lm > um ? lm : um) < msl) {               /* L1252-1258 */
                emit_len = msl;
                break;
            }
            last_cnt = cnt; last_lo = lo;
            ml = lm > um ? lm : um;
        }
        if (cnt > 0 && cnt < (uint64_t)r->min_intv) {      /* L1265-1277 */
            if (emit_len < msl) emit_len = msl;
            push_smem(r, pivot, pivot + emit_len, lo, cnt);
        }
        pivot += emit_len;
    }
}

static int smem_cmp(const void *a, const void *b)
{
    const oracle_smem_t *x = (const oracle_smem_t *)a, *y = (const oracle_smem_t *)b;
    if (x->start != y->start) return x->start < y->start ? -1 : 1;
    if (x->end != y->end) return x->end < y->end ? -1 : 1;
    return x->hitbeg < y->hitbeg ? -1 : (x->hitbeg > y->hitbeg);   /* stable */
}

void oracle_seed_result_free(oracle_seed_result_t *res)
{
    if (!res) return;
    free(res->smems); free(res->smem_off); free(res->hit_off); free(res->hits); free(res);
}

/* mem_kernel1_core_Learned, src/bwamem.cpp:L1249-1397, for a batch of reads given as codes 0..4 */
oracle_seed_result_t *oracle_seed_reads(const oracle_index_t *ix, const uint8_t *codes, const int64_t *off,
                                        int nreads, const oracle_seed_opt_t *opt)
{
    oracle_seed_result_t *res = (oracle_seed_result_t *)calloc(1, sizeof(*res));
    res->smem_off = (int64_t *)calloc((size_t)nreads + 1, 8);
    res->hit_off = (int64_t *)calloc((size_t)nreads + 1, 8);
    int64_t m_sm = 0, m_hits = 0;
    rd_t r;
    memset(&r, 0, sizeof r);
    r.ix = ix;
    r.cnt = res->counters;
    uint8_t *rc = NULL; int m_rc = 0;
    for (int i = 0; i < nreads; i++) {
        const uint8_t *seq = codes + off[i];
        int len = (int)(off[i + 1] - off[i]);
        if (len > m_rc) { m_rc = len * 2; rc = (uint8_t *)realloc(rc, (size_t)m_rc); }
        for (int k = 0; k < len; k++) rc[len - 1 - k] = seq[k] < 4 ? 3 - seq[k] : 4;
        r.fw = seq; r.rc = rc; r.l_seq = len;
        r.n_sm = 0; r.n_hits = 0;
        r.full_match = 0; r.use_cached = 0;
        r.min_seed_len = opt->min_seed_len;
        r.min_intv = 1;
        if (len > 0) {
            seed_allpos(&r, opt);
            if (opt->rounds >= 3 && opt->max_mem_intv > 0) {
                r.min_intv = opt->max_mem_intv;
                r.min_seed_len = opt->min_seed_len + 1;
                seed_round3(&r);
            }
        }
        qsort(r.sm, (size_t)r.n_sm, sizeof(oracle_smem_t), smem_cmp);
        /* append, rewriting hitbeg so that each read's hit list is in sorted-SMEM order */
        if (res->n_smems + r.n_sm > m_sm) {
            m_sm = (res->n_smems + r.n_sm) * 2 + 1024;
            res->smems = (oracle_smem_t *)realloc(res->smems, (size_t)m_sm * sizeof(oracle_smem_t));
        }
        if (res->n_hits + r.n_hits > m_hits) {
            m_hits = (res->n_hits + r.n_hits) * 2 + 4096;
            res->hits = (uint64_t *)realloc(res->hits, (size_t)m_hits * 8);
        }
        int64_t hb = 0;
        for (int k = 0; k < r.n_sm; k++) {
            oracle_smem_t m = r.sm[k];
            memcpy(res->hits + res->n_hits + hb, r.hits + m.hitbeg, (size_t)m.hitcount * 8);
            m.hitbeg = (int32_t)hb;
            hb += m.hitcount;
            res->smems[res->n_smems + k] = m;
        }
        res->n_smems += r.n_sm;
        res->n_hits += hb;
        res->smem_off[i + 1] = res->n_smems;
        res->hit_off[i + 1] = res->n_hits;
    }
    free(rc); free(r.sm); free(r.hits);
    return res;
}
