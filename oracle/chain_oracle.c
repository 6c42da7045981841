/* oracle/chain_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's chaining stage of `bwa-meme mem -7`: what
 * mem_kernel1_core_Learned does with a read's sorted SMEMs and hit list after seeding.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the product path never does.
 *
 * Parity pinning: tests/test_oracle_chain.py compares this file's output with chain dumps of the reference's own
 * mem_chain_Learned + mem_chain_flt (the patched aligner of INTEGRATION.md run with every hook OFF executes the
 * unmodified reference code and, with MEME_B200_DUMP_CHAINS=path, prints chain_ar after the chaining stage) --
 * committed as tests/golden/chains_*.txt and regenerated live whenever oracle/_ref exists.
 *
 * What is restated, and where it lives in the reference (paths under /root/reference):
 *   pos2rid(), intv2rid()   src/bntseq.cpp:L392-416      bns_pos2rid, bns_intv2rid (bns_depos: src/bntseq.h:L89-92)
 *   bt_*()                  src/kbtree.h:L58-238, L356-377  kb_init, __kb_getp_aux, kb_intervalp, __kb_split,
 *                                                        __kb_putp_aux, kb_putp, __kb_traverse for KBTREE_INIT(chn, mem_chain_t,
 *                                                        chain_cmp) (src/bwamem.cpp:L43-44)
 *   merge_seed()            src/bwamem.cpp:L448-493      test_and_merge
 *   chain_read()            src/bwamem.cpp:L1122-1204    mem_chain_Learned
 *   weight()                src/bwamem.cpp:L525-545      mem_chain_weight
 *   sort_chains()           src/ksort.h:L155-236         __ks_insertsort, ks_combsort, ks_introsort for KSORT_INIT(mem_flt,
 *                                                        mem_chain_t, flt_lt) (src/bwamem.cpp:L80-81)
 *   filter_chains()         src/bwamem.cpp:L599-717      mem_chain_flt (one read = one range)
 * mem_flt_chained_seeds (L565-597) skips reads for which MEM_MINSC_COEF*log(l) > MEM_SEEDSW_COEF*l, i.e. every read
 * shorter than about 700 bases; the -7 path asserts reads of at most 301 bases, so it is not restated.
 *
 * Records are moved BY VALUE through tree nodes and the sort, as the reference moves mem_chain_t, and a chain's seeds
 * sit in a growing array: deliberately the reference's storage, not the index-based layout of the CUDA kernel.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

typedef struct { int32_t start, end, hitbeg, hitcount; uint64_t sa_lo; } smem_t;     /* mem_tl, LearnedIndex_seeding.h:L121-127 */
typedef struct { int64_t offset; int32_t len, is_alt; } contig_t;
typedef struct { int32_t w, max_chain_gap, max_occ, min_seed_len, min_chain_weight, max_chain_extend; float mask_level, drop_ratio; } chain_opt_t;
typedef struct { int64_t rbeg; int32_t qbeg, len; } seed_t;
typedef struct { int64_t pos, seed_beg; int32_t n, rid, w, first, kept, is_alt; } chain_out_t;

typedef struct {
    int64_t pos;
    seed_t *seeds;
    int n, m, rid, first;
    int w, kept, is_alt;
} chain_t;

typedef struct {
    const contig_t *contigs; int n_contigs; int64_t l_pac;
} bns_t;

static int pos2rid(const bns_t *b, int64_t pos_f)
{
    int left, mid, right;
    if (pos_f >= b->l_pac) return -1;
    left = 0; mid = 0; right = b->n_contigs;
    while (left < right) {
        mid = (left + right) >> 1;
        if (pos_f >= b->contigs[mid].offset) {
            if (mid == b->n_contigs - 1) break;
            if (pos_f < b->contigs[mid + 1].offset) break;
            left = mid + 1;
        } else right = mid;
    }
    return mid;
}

static int64_t depos(const bns_t *b, int64_t pos) { return pos >= b->l_pac ? (b->l_pac << 1) - 1 - pos : pos; }

static int intv2rid(const bns_t *b, int64_t rb, int64_t re)
{
    int rid_b, rid_e;
    if (rb < b->l_pac && re > b->l_pac) return -2;
    rid_b = pos2rid(b, depos(b, rb));
    rid_e = rb < re ? pos2rid(b, depos(b, re - 1)) : rid_b;
    return rid_b == rid_e ? rid_b : -1;
}

/* ---- the B-tree: nodes of at most 2t-1 chain records, t = 5 for a 48-byte record and 520-byte nodes ---- */
#define BT_T ((int)((((512 + 8) - 4 - sizeof(void *)) / (sizeof(void *) + 48) + 1) >> 1))
#define BT_MAX 16

typedef struct bt_node {
    int n, internal;
    chain_t key[BT_MAX];
    struct bt_node *ptr[BT_MAX + 1];
} bt_node;

typedef struct { bt_node *root; int n_keys, t; } btree;

static int cmp_pos(int64_t a, int64_t b) { return (b < a) - (a < b); }

/* position of the first key not below k; *r tells how k compares with the key the caller is left at */
static int bt_find(const bt_node *x, int64_t k, int *r)
{
    int tr, *rr = r ? r : &tr, begin = 0, end = x->n;
    if (x->n == 0) return -1;
    while (begin < end) {
        int mid = (begin + end) >> 1;
        if (cmp_pos(x->key[mid].pos, k) < 0) begin = mid + 1; else end = mid;
    }
    if (begin == x->n) { *rr = 1; return x->n - 1; }
    if ((*rr = cmp_pos(k, x->key[begin].pos)) < 0) --begin;
    return begin;
}

static chain_t *bt_lower(btree *b, int64_t k)
{
    chain_t *lower = 0;
    bt_node *x = b->root;
    while (x) {
        int r = 0, i = bt_find(x, k, &r);
        if (i >= 0 && r == 0) return &x->key[i];
        if (i >= 0) lower = &x->key[i];
        if (!x->internal) return lower;
        x = x->ptr[i + 1];
    }
    return lower;
}

static void bt_split(btree *b, bt_node *x, int i, bt_node *y)
{
    bt_node *z = (bt_node *)calloc(1, sizeof(bt_node));
    const int t = b->t;
    z->internal = y->internal;
    z->n = t - 1;
    memcpy(z->key, y->key + t, sizeof(chain_t) * (t - 1));
    if (y->internal) memcpy(z->ptr, y->ptr + t, sizeof(bt_node *) * t);
    y->n = t - 1;
    memmove(x->ptr + i + 2, x->ptr + i + 1, sizeof(bt_node *) * (x->n - i));
    x->ptr[i + 1] = z;
    memmove(x->key + i + 1, x->key + i, sizeof(chain_t) * (x->n - i));
    x->key[i] = y->key[t - 1];
    ++x->n;
}

static void bt_put_nonfull(btree *b, bt_node *x, const chain_t *k)
{
    int i;
    if (!x->internal) {
        i = bt_find(x, k->pos, 0);
        if (i != x->n - 1) memmove(x->key + i + 2, x->key + i + 1, (x->n - i - 1) * sizeof(chain_t));
        x->key[i + 1] = *k;
        ++x->n;
    } else {
        i = bt_find(x, k->pos, 0) + 1;
        if (x->ptr[i]->n == 2 * b->t - 1) {
            bt_split(b, x, i, x->ptr[i]);
            if (cmp_pos(k->pos, x->key[i].pos) > 0) ++i;
        }
        bt_put_nonfull(b, x->ptr[i], k);
    }
}

static void bt_put(btree *b, const chain_t *k)
{
    bt_node *r = b->root, *s;
    ++b->n_keys;
    if (r->n == 2 * b->t - 1) {
        s = (bt_node *)calloc(1, sizeof(bt_node));
        b->root = s; s->internal = 1; s->n = 0;
        s->ptr[0] = r;
        bt_split(b, s, 0, r);
        r = s;
    }
    bt_put_nonfull(b, r, k);
}

static void bt_walk(const bt_node *x, chain_t *out, int *n)
{
    int i;
    for (i = 0; i <= x->n; i++) {
        if (x->internal) bt_walk(x->ptr[i], out, n);
        if (i < x->n) out[(*n)++] = x->key[i];
    }
}

static void bt_free(bt_node *x)
{
    int i;
    if (!x) return;
    if (x->internal) for (i = 0; i <= x->n; i++) bt_free(x->ptr[i]);
    free(x);
}

/* test_and_merge: 1 = the seed is accounted for by chain c (contained, or appended) */
static int merge_seed(const chain_opt_t *o, int64_t l_pac, chain_t *c, const seed_t *p, int seed_rid)
{
    const seed_t *last = &c->seeds[c->n - 1];
    int64_t qend = last->qbeg + last->len, rend = last->rbeg + last->len, x, y;
    if (seed_rid != c->rid) return 0;
    if (p->qbeg >= c->seeds[0].qbeg && p->qbeg + p->len <= qend && p->rbeg >= c->seeds[0].rbeg && p->rbeg + p->len <= rend)
        return 1;
    if ((last->rbeg < l_pac || c->seeds[0].rbeg < l_pac) && p->rbeg >= l_pac) return 0;
    x = p->qbeg - last->qbeg;
    y = p->rbeg - last->rbeg;
    if (y >= 0 && x - y <= o->w && y - x <= o->w && x - last->len < o->max_chain_gap && y - last->len < o->max_chain_gap) {
        if (c->n == c->m) {
            c->m <<= 1;
            c->seeds = (seed_t *)realloc(c->seeds, c->m * sizeof(seed_t));
        }
        c->seeds[c->n++] = *p;
        return 1;
    }
    return 0;
}

static int weight(const chain_t *c)
{
    int64_t end;
    int j, w = 0, tmp;
    for (j = 0, end = 0; j < c->n; ++j) {
        const seed_t *s = &c->seeds[j];
        if (s->qbeg >= end) w += s->len;
        else if (s->qbeg + s->len > end) w += s->qbeg + s->len - end;
        end = end > s->qbeg + s->len ? end : s->qbeg + s->len;
    }
    tmp = w; w = 0;
    for (j = 0, end = 0; j < c->n; ++j) {
        const seed_t *s = &c->seeds[j];
        if (s->rbeg >= end) w += s->len;
        else if (s->rbeg + s->len > end) w += s->rbeg + s->len - end;
        end = end > s->rbeg + s->len ? end : s->rbeg + s->len;
    }
    w = w < tmp ? w : tmp;
    return w < 1 << 30 ? w : (1 << 30) - 1;
}

/* ---- the sort of mem_chain_flt: heavier first, klib's introsort moved record by record ---- */
#define HEAVIER(a, b) ((a).w > (b).w)

static void insertion(chain_t *s, chain_t *t)
{
    chain_t *i, *j, tmp;
    for (i = s + 1; i < t; ++i)
        for (j = i; j > s && HEAVIER(*j, *(j - 1)); --j) { tmp = *j; *j = *(j - 1); *(j - 1) = tmp; }
}

static void comb(size_t n, chain_t a[])
{
    const double shrink_factor = 1.2473309501039786540366528676643;
    int do_swap;
    size_t gap = n;
    chain_t tmp, *i, *j;
    do {
        if (gap > 2) {
            gap = (size_t)(gap / shrink_factor);
            if (gap == 9 || gap == 10) gap = 11;
        }
        do_swap = 0;
        for (i = a; i < a + n - gap; ++i) {
            j = i + gap;
            if (HEAVIER(*j, *i)) { tmp = *i; *i = *j; *j = tmp; do_swap = 1; }
        }
    } while (do_swap || gap > 2);
    if (gap != 1) insertion(a, a + n);
}

typedef struct { chain_t *left, *right; int depth; } sort_frame;

static void sort_chains(size_t n, chain_t a[])
{
    int d;
    sort_frame *top, *stack;
    chain_t rp, swap_tmp;
    chain_t *s, *t, *i, *j, *k;
    if (n < 1) return;
    else if (n == 2) {
        if (HEAVIER(a[1], a[0])) { swap_tmp = a[0]; a[0] = a[1]; a[1] = swap_tmp; }
        return;
    }
    for (d = 2; 1ul << d < n; ++d);
    stack = (sort_frame *)malloc(sizeof(sort_frame) * ((sizeof(size_t) * d) + 2));
    top = stack; s = a; t = a + (n - 1); d <<= 1;
    while (1) {
        if (s < t) {
            if (--d == 0) {
                comb(t - s + 1, s);
                t = s;
                continue;
            }
            i = s; j = t; k = i + ((j - i) >> 1) + 1;
            if (HEAVIER(*k, *i)) {
                if (HEAVIER(*k, *j)) k = j;
            } else k = HEAVIER(*j, *i) ? i : j;
            rp = *k;
            if (k != t) { swap_tmp = *k; *k = *t; *t = swap_tmp; }
            for (;;) {
                do ++i; while (HEAVIER(*i, rp));
                do --j; while (i <= j && HEAVIER(rp, *j));
                if (j <= i) break;
                swap_tmp = *i; *i = *j; *j = swap_tmp;
            }
            swap_tmp = *i; *i = *t; *t = swap_tmp;
            if (i - s > t - i) {
                if (i - s > 16) { top->left = s; top->right = i - 1; top->depth = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { top->left = i + 1; top->right = t; top->depth = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == stack) {
                free(stack);
                insertion(a, a + n);
                return;
            } else { --top; s = top->left; t = top->right; d = top->depth; }
        }
    }
}

#define C_BEG(c) ((c).seeds[0].qbeg)
#define C_END(c) ((c).seeds[(c).n - 1].qbeg + (c).seeds[(c).n - 1].len)

/* mem_chain_flt for the chains of one read; dropped chains lose their seed arrays; returns the number kept */
static int filter_chains(const chain_opt_t *o, int n_chn, chain_t *a)
{
    int i, k, n_kept = 0, n_dead = 0, *kept;
    seed_t **dead;
    if (n_chn == 0) return 0;
    dead = (seed_t **)malloc(sizeof(seed_t *) * n_chn);
    for (i = k = 0; i < n_chn; ++i) {
        chain_t *c = &a[i];
        c->first = -1; c->kept = 0;
        c->w = weight(c);
        if (c->w < o->min_chain_weight) dead[n_dead++] = c->seeds;
        else a[k++] = *c;
    }
    /* When every chain is too light the reference carries on with one range [0,1) over the stale first record
     * (L624-641; its seed array may already be freed there).  The oracle keeps that record's seeds alive so that the
     * outcome is defined: one chain, the first in tree order. */
    for (i = k ? 0 : 1; i < n_dead; ++i) free(dead[i]);
    free(dead);
    n_chn = k ? k : 1;
    sort_chains(n_chn, a);
    kept = (int *)malloc(sizeof(int) * n_chn);
    a[0].kept = 3;
    kept[n_kept++] = 0;
    for (i = 1; i < n_chn; ++i) {
        int large_ovlp = 0;
        for (k = 0; k < n_kept; ++k) {
            int j = kept[k];
            int b_max = C_BEG(a[j]) > C_BEG(a[i]) ? C_BEG(a[j]) : C_BEG(a[i]);
            int e_min = C_END(a[j]) < C_END(a[i]) ? C_END(a[j]) : C_END(a[i]);
            if (e_min > b_max && (!a[j].is_alt || a[i].is_alt)) {
                int li = C_END(a[i]) - C_BEG(a[i]);
                int lj = C_END(a[j]) - C_BEG(a[j]);
                int min_l = li < lj ? li : lj;
                if (e_min - b_max >= min_l * o->mask_level && min_l < o->max_chain_gap) {
                    large_ovlp = 1;
                    if (a[j].first < 0) a[j].first = i;
                    if (a[i].w < a[j].w * o->drop_ratio && a[j].w - a[i].w >= o->min_seed_len << 1) break;
                }
            }
        }
        if (k == n_kept) {
            kept[n_kept++] = i;
            a[i].kept = large_ovlp ? 2 : 3;
        }
    }
    for (i = 0; i < n_kept; ++i) {
        chain_t *c = &a[kept[i]];
        if (c->first >= 0) a[c->first].kept = 1;
    }
    free(kept);
    for (i = k = 0; i < n_chn; ++i) {
        if (a[i].kept == 0 || a[i].kept == 3) continue;
        if (++k >= o->max_chain_extend) break;
    }
    for (; i < n_chn; ++i)
        if (a[i].kept < 3) a[i].kept = 0;
    for (i = k = 0; i < n_chn; ++i) {
        chain_t *c = &a[i];
        if (c->kept == 0) { free(c->seeds); c->seeds = 0; }
        else a[k++] = a[i];
    }
    return k;
}

/* mem_chain_Learned for one read: smems sorted by (start, end), hits = the read's FULL hit list (SA order per SMEM) */
static int chain_read(const chain_opt_t *o, const bns_t *bns, int len, const smem_t *smems, int n_smems, const uint64_t *hits,
                      chain_t **out, float *frac_rep)
{
    int i, b = 0, e = 0, l_rep = 0, n = 0;
    btree tree;
    chain_t *a;
    *out = 0; *frac_rep = 0;
    if (len < o->min_seed_len) return 0;
    tree.t = BT_T; tree.n_keys = 0;
    tree.root = (bt_node *)calloc(1, sizeof(bt_node));
    for (i = 0; i < n_smems; ++i) {
        const smem_t *p = &smems[i];
        int sb = p->start, se = p->end;
        if (p->hitcount <= o->max_occ) continue;
        if (sb > e) l_rep += e - b, b = sb, e = se;
        else e = e > se ? e : se;
    }
    l_rep += e - b;
    for (i = 0; i < n_smems; ++i) {
        const smem_t *p = &smems[i];
        int step, count;
        int64_t k;
        step = p->hitcount > o->max_occ ? p->hitcount / o->max_occ : 1;
        for (k = count = 0; k < p->hitcount && count < o->max_occ; k += step, ++count) {
            chain_t tmp, *lower;
            seed_t s;
            int rid, to_add = 0;
            s.rbeg = tmp.pos = (int64_t)hits[p->hitbeg + k];
            s.qbeg = p->start;
            s.len = p->end - p->start;
            rid = intv2rid(bns, s.rbeg, s.rbeg + s.len);
            if (rid < 0) continue;
            if (tree.n_keys) {
                lower = bt_lower(&tree, tmp.pos);
                if (!lower || !merge_seed(o, bns->l_pac, lower, &s, rid)) to_add = 1;
            } else to_add = 1;
            if (to_add) {
                tmp.n = 1; tmp.m = 4;
                tmp.seeds = (seed_t *)calloc(tmp.m, sizeof(seed_t));
                tmp.seeds[0] = s;
                tmp.rid = rid;
                tmp.is_alt = !!bns->contigs[rid].is_alt;
                tmp.first = -1; tmp.w = 0; tmp.kept = 0;
                bt_put(&tree, &tmp);
            }
        }
    }
    a = (chain_t *)malloc(sizeof(chain_t) * (tree.n_keys + 1));
    bt_walk(tree.root, a, &n);
    bt_free(tree.root);
    *frac_rep = (float)l_rep / len;
    *out = a;
    return n;
}

typedef struct {
    int64_t n_reads, n_chains, n_seeds;
    int64_t *chain_off; float *frac_rep; chain_out_t *chains; seed_t *seeds;
} chain_result_t;

/* All reads of a batch.  smems/hits are grouped by read through smem_off/hit_off; hitbeg is relative to the read's list. */
chain_result_t *oracle_chain_reads(const smem_t *smems, const int64_t *smem_off, const uint64_t *hits, const int64_t *hit_off,
                                   const int32_t *read_len, int64_t n_reads, const contig_t *contigs, int32_t n_contigs,
                                   int64_t l_pac, const chain_opt_t *o)
{
    chain_result_t *R = (chain_result_t *)calloc(1, sizeof(chain_result_t));
    bns_t bns;
    size_t cap_c = 1024, cap_s = 4096;
    int64_t r;
    bns.contigs = contigs; bns.n_contigs = n_contigs; bns.l_pac = l_pac;
    R->n_reads = n_reads;
    R->chain_off = (int64_t *)calloc((size_t)n_reads + 1, 8);
    R->frac_rep = (float *)calloc((size_t)n_reads + 1, 4);
    R->chains = (chain_out_t *)malloc(cap_c * sizeof(chain_out_t));
    R->seeds = (seed_t *)malloc(cap_s * sizeof(seed_t));
    for (r = 0; r < n_reads; r++) {
        chain_t *a = 0;
        int i, n = chain_read(o, &bns, read_len[r], smems + smem_off[r], (int)(smem_off[r + 1] - smem_off[r]), hits + hit_off[r], &a,
                              &R->frac_rep[r]);
        n = filter_chains(o, n, a);
        R->chain_off[r] = R->n_chains;
        for (i = 0; i < n; i++) {
            chain_out_t c;
            if ((size_t)R->n_chains + 1 > cap_c) { cap_c *= 2; R->chains = (chain_out_t *)realloc(R->chains, cap_c * sizeof(chain_out_t)); }
            while ((size_t)R->n_seeds + a[i].n > cap_s) { cap_s *= 2; R->seeds = (seed_t *)realloc(R->seeds, cap_s * sizeof(seed_t)); }
            c.pos = a[i].pos; c.seed_beg = R->n_seeds; c.n = a[i].n; c.rid = a[i].rid; c.w = a[i].w; c.first = a[i].first;
            c.kept = a[i].kept; c.is_alt = a[i].is_alt;
            memcpy(R->seeds + R->n_seeds, a[i].seeds, sizeof(seed_t) * a[i].n);
            R->chains[R->n_chains++] = c;
            R->n_seeds += a[i].n;
            free(a[i].seeds);
        }
        free(a);
    }
    R->chain_off[n_reads] = R->n_chains;
    return R;
}

void oracle_chain_result_free(chain_result_t *R)
{
    if (!R) return;
    free(R->chain_off); free(R->frac_rep); free(R->chains); free(R->seeds); free(R);
}
