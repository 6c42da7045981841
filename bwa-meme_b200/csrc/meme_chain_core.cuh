// meme_chain_core.cuh -- chaining of one read's seeds, as one thread runs it.
//
// What it computes is the reference's mem_chain_Learned (src/bwamem.cpp:L1122-1204: seeds in SMEM order, each merged into the
// closest chain below it or opened as a new chain, chains kept in a B-tree keyed by reference position) followed by
// mem_chain_flt (src/bwamem.cpp:L599-717: chain weights, sort by weight, overlap filter).  The outcome depends on details that
// look incidental -- which of several chains with EQUAL position kb_intervalp returns (src/kbtree.h:L158-176: the first equal
// key of the highest node on the search path), where kb_putp places a duplicate (L204-221: right after the first equal key
// of the leaf), and how the unstable ks_introsort (src/ksort.h:L185-236) orders chains of equal weight -- so the tree (order
// t = 5, as kb_init(chn, KB_DEFAULT_SIZE + 8) computes for the 48-byte mem_chain_t) and the sort are reproduced operation by
// operation.  What is NOT reproduced is their storage: nodes hold 4-byte chain ids instead of 48-byte records, a chain's
// seeds are a linked list through one per-read array (the reference grows a seed array per chain), the sort moves ids.
//
// Plain C++ without CUDA types below MEME_HD, so that tests/host_sim compiles the same source with g++.
#pragma once
#include <stdint.h>
#include "../../include/meme_b200.h"

#ifdef __CUDACC__
#define MEME_HD __host__ __device__ __forceinline__
#else
#define MEME_HD inline
#endif

namespace chn {

typedef unsigned int cu32;
typedef unsigned long long cu64;

enum { KB_T = 5, KB_MAXKEYS = 2 * KB_T - 1 };
#define CHN_NIL 0xffffffffu

struct Rec {            // one chain while it is being built / filtered
    cu32 first_seed, last_seed;   // seed = index into the read's (sampled) hit list
    int n, rid;
    int w, first;
    int kept, is_alt;
};

struct Node {           // kbtree node: keys are chain ids, ptr are node ids
    int n, is_internal;
    cu32 key[KB_MAXKEYS];
    cu32 ptr[KB_MAXKEYS + 1];
};

struct Scratch {        // per-read slices (see meme_chain.cu for the sizing)
    cu32 *next;         // [seeds] next seed of the same chain
    cu32 *ql;           // [seeds] qbeg | len << 16
    Rec *ch;            // [seeds] chains (never more than seeds)
    cu32 *order;        // [seeds] chain ids: tree order, then sorted / filtered order
    cu32 *kept;         // [seeds] positions of the chains kept by the overlap filter
    Node *nodes;        // [seeds / 4 + 2]
    int max_nodes;
};

struct View {           // the read's seeding results and what chaining needs of the index annotation
    const meme_smem_t *smems; int n_smems;
    const cu64 *hits;           // sampled: SMEM i owns hits[hitbeg .. hitbeg + take(i))
    int read_len;
    const meme_contig_t *contigs; int n_contigs;
    int64_t l_pac;
};

// bns_pos2rid (src/bntseq.cpp:L392-406), including what it returns when the loop runs out without bracketing
MEME_HD int pos2rid(const View &v, int64_t pos_f)
{
    if (pos_f >= v.l_pac) return -1;
    int left = 0, mid = 0, right = v.n_contigs;
    while (left < right) {
        mid = (left + right) >> 1;
        if (pos_f >= v.contigs[mid].offset) {
            if (mid == v.n_contigs - 1) break;
            if (pos_f < v.contigs[mid + 1].offset) break;
            left = mid + 1;
        } else right = mid;
    }
    return mid;
}

// bns_intv2rid (src/bntseq.cpp:L408-416)
MEME_HD int intv2rid(const View &v, int64_t rb, int64_t re)
{
    if (rb < v.l_pac && re > v.l_pac) return -2;
    const int64_t two = v.l_pac << 1;
    int rid_b = pos2rid(v, rb >= v.l_pac ? two - 1 - rb : rb);
    int rid_e = rb < re ? pos2rid(v, re - 1 >= v.l_pac ? two - 1 - (re - 1) : re - 1) : rid_b;
    return rid_b == rid_e ? rid_b : -1;
}

struct Tree {
    Node *nodes; int n_nodes, max_nodes, n_keys; cu32 root;
    const cu64 *hits; const Rec *ch;
    bool overflow;
    MEME_HD int64_t pos_of(cu32 chain) const { return (int64_t)hits[ch[chain].first_seed]; }
    // __kb_getp_aux (src/kbtree.h:L124-138): first key >= k; r < 0 / == 0 / > 0 as k compares with the key looked at
    MEME_HD int getp_aux(const Node &x, int64_t k, int *r) const
    {
        int begin = 0, end = x.n;
        if (x.n == 0) return -1;
        while (begin < end) {
            int mid = (begin + end) >> 1;
            if (pos_of(x.key[mid]) < k) begin = mid + 1; else end = mid;
        }
        if (begin == x.n) { *r = 1; return x.n - 1; }
        int64_t p = pos_of(x.key[begin]);
        *r = (p < k) - (k < p);
        if (*r < 0) --begin;
        return begin;
    }
    // kb_intervalp (L158-176), lower bound only: the chain at or below position k, CHN_NIL if none
    MEME_HD cu32 lower(int64_t k) const
    {
        cu32 low = CHN_NIL, x = root;
        for (;;) {
            const Node &nd = nodes[x];
            int r = 0, i = getp_aux(nd, k, &r);
            if (i >= 0 && r == 0) return nd.key[i];
            if (i >= 0) low = nd.key[i];
            if (!nd.is_internal) return low;
            x = nd.ptr[i + 1];
        }
    }
    MEME_HD cu32 new_node(int internal)
    {
        if (n_nodes >= max_nodes) { overflow = true; return 0; }
        Node &z = nodes[n_nodes];
        z.n = 0; z.is_internal = internal;
        return (cu32)n_nodes++;
    }
    // __kb_split (L183-199): child y = ptr[i] of x is full; its upper t-1 keys move to a new right sibling
    MEME_HD void split(cu32 xi, int i, cu32 yi)
    {
        cu32 zi = new_node(nodes[yi].is_internal);
        if (overflow) return;
        Node &x = nodes[xi], &y = nodes[yi], &z = nodes[zi];
        z.n = KB_T - 1;
        for (int k = 0; k < KB_T - 1; k++) z.key[k] = y.key[KB_T + k];
        if (y.is_internal) for (int k = 0; k < KB_T; k++) z.ptr[k] = y.ptr[KB_T + k];
        y.n = KB_T - 1;
        for (int k = x.n; k > i; k--) x.ptr[k + 1] = x.ptr[k];
        x.ptr[i + 1] = zi;
        for (int k = x.n; k > i; k--) x.key[k] = x.key[k - 1];
        x.key[i] = y.key[KB_T - 1];
        ++x.n;
    }
    // kb_putp + __kb_putp_aux (L200-238), the recursion unrolled into a descent
    MEME_HD void put(cu32 chain)
    {
        const int64_t k = pos_of(chain);
        ++n_keys;
        cu32 r = root;
        if (nodes[r].n == KB_MAXKEYS) {
            cu32 s = new_node(1);
            if (overflow) return;
            root = s;
            nodes[s].ptr[0] = r;
            split(s, 0, r);
            if (overflow) return;
            r = s;
        }
        for (;;) {
            Node &x = nodes[r];
            int dummy = 0;
            if (!x.is_internal) {
                int i = getp_aux(x, k, &dummy);
                for (int m = x.n - 1; m > i; m--) x.key[m + 1] = x.key[m];
                x.key[i + 1] = chain;
                ++x.n;
                return;
            }
            int i = getp_aux(x, k, &dummy) + 1;
            if (nodes[x.ptr[i]].n == KB_MAXKEYS) {
                split(r, i, x.ptr[i]);
                if (overflow) return;
                int64_t p = pos_of(nodes[r].key[i]);
                if (k > p) ++i;
            }
            r = nodes[r].ptr[i];
        }
    }
    // __kb_traverse (L356-377): keys in order
    MEME_HD int traverse(cu32 *out) const
    {
        // explicit stack of (node, next index); the tree of a read is at most 9 levels deep (5^8 > 96 k chains)
        cu32 sx[12]; int si[12]; int sp = 0, n_out = 0;
        sx[0] = root; si[0] = 0;
        for (;;) {
            const Node &nd = nodes[sx[sp]];
            if (nd.is_internal && si[sp] <= nd.n) {            // next subtree of this node
                if (sp == 11) return -1;
                sx[sp + 1] = nd.ptr[si[sp]]; si[sp + 1] = 0; ++sp;
                continue;
            }
            if (!nd.is_internal) for (int k = 0; k < nd.n; k++) out[n_out++] = nd.key[k];
            if (sp == 0) return n_out;
            --sp;                                              // back in the parent: the key after that subtree
            const Node &pa = nodes[sx[sp]];
            if (si[sp] < pa.n) out[n_out++] = pa.key[si[sp]];
            ++si[sp];
        }
    }
};

// mem_chain_weight (src/bwamem.cpp:L525-545)
MEME_HD int chain_weight(const Rec &c, const Scratch &S, const cu64 *hits)
{
    int64_t end = 0;
    int w = 0, tmp;
    cu32 s = c.first_seed;
    for (int j = 0; j < c.n; j++, s = S.next[s]) {
        const int qb = (int)(S.ql[s] & 0xffff), len = (int)(S.ql[s] >> 16);
        if (qb >= end) w += len;
        else if (qb + len > end) w += qb + len - (int)end;
        end = end > qb + len ? end : qb + len;
    }
    tmp = w; w = 0; end = 0;
    s = c.first_seed;
    for (int j = 0; j < c.n; j++, s = S.next[s]) {
        const int64_t rb = (int64_t)hits[s]; const int len = (int)(S.ql[s] >> 16);
        if (rb >= end) w += len;
        else if (rb + len > end) w += (int)(rb + len - end);
        end = end > rb + len ? end : rb + len;
    }
    w = w < tmp ? w : tmp;
    return w < 1 << 30 ? w : (1 << 30) - 1;
}

// flt_lt (src/bwamem.cpp:L80): heavier chains first
#define CHN_LT(a, b) (S.ch[a].w > S.ch[b].w)

MEME_HD void insertsort(cu32 *s, cu32 *t, const Scratch &S)
{
    for (cu32 *i = s + 1; i < t; ++i)
        for (cu32 *j = i; j > s && CHN_LT(*j, *(j - 1)); --j) { cu32 tmp = *j; *j = *(j - 1); *(j - 1) = tmp; }
}

// ks_combsort (src/ksort.h:L163-184)
MEME_HD void combsort(size_t n, cu32 *a, const Scratch &S)
{
    const double shrink_factor = 1.2473309501039786540366528676643;
    int do_swap;
    size_t gap = n;
    do {
        if (gap > 2) {
            gap = (size_t)(gap / shrink_factor);
            if (gap == 9 || gap == 10) gap = 11;
        }
        do_swap = 0;
        for (cu32 *i = a; i < a + n - gap; ++i) {
            cu32 *j = i + gap;
            if (CHN_LT(*j, *i)) { cu32 tmp = *i; *i = *j; *j = tmp; do_swap = 1; }
        }
    } while (do_swap || gap > 2);
    if (gap != 1) insertsort(a, a + n, S);
}

// ks_introsort (src/ksort.h:L185-236).  Returns false if the explicit stack would overflow (never: the larger part is
// pushed, the smaller one continued, so the depth is bounded by log2 n).
MEME_HD bool introsort(size_t n, cu32 *a, const Scratch &S)
{
    if (n < 1) return true;
    if (n == 2) {
        if (CHN_LT(a[1], a[0])) { cu32 tmp = a[0]; a[0] = a[1]; a[1] = tmp; }
        return true;
    }
    int d;
    for (d = 2; 1ul << d < n; ++d);
    struct { cu32 *left, *right; int depth; } stack[48];
    int top = 0;
    cu32 *s = a, *t = a + (n - 1), *i, *j, *k;
    d <<= 1;
    for (;;) {
        if (s < t) {
            if (--d == 0) {
                combsort((size_t)(t - s) + 1, s, S);
                t = s;
                continue;
            }
            i = s; j = t; k = i + ((j - i) >> 1) + 1;
            if (CHN_LT(*k, *i)) {
                if (CHN_LT(*k, *j)) k = j;
            } else k = CHN_LT(*j, *i) ? i : j;
            const cu32 rp = *k;
            if (k != t) { cu32 tmp = *k; *k = *t; *t = tmp; }
            for (;;) {
                do ++i; while (CHN_LT(*i, rp));
                do --j; while (i <= j && CHN_LT(rp, *j));
                if (j <= i) break;
                cu32 tmp = *i; *i = *j; *j = tmp;
            }
            { cu32 tmp = *i; *i = *t; *t = tmp; }
            if (i - s > t - i) {
                if (i - s > 16) { if (top == 48) return false; stack[top].left = s; stack[top].right = i - 1; stack[top].depth = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { if (top == 48) return false; stack[top].left = i + 1; stack[top].right = t; stack[top].depth = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == 0) { insertsort(a, a + n, S); return true; }
            --top; s = stack[top].left; t = stack[top].right; d = stack[top].depth;
        }
    }
}

// One read.  Leaves the ids of the surviving chains in S.order[0 .. *n_out) (final order) and returns the number of
// seeds they hold; *frac_rep as mem_chain_Learned sets it.  err: 1 = scratch overflow (sizing bug), 2 = sort stack.
MEME_HD int chain_read(const meme_chain_opt_t &o, const View &v, const Scratch &S, int *n_out, float *frac_rep, int *err)
{
    *n_out = 0; *frac_rep = 0.f;
    if (v.read_len < o.min_seed_len) return 0;                 // L1134
    // ---- fraction of the read covered by repetitive SMEMs (L1136-1143) ----
    int b = 0, e = 0, l_rep = 0;
    for (int i = 0; i < v.n_smems; i++) {
        const meme_smem_t &p = v.smems[i];
        if (p.hitcount <= o.max_occ) continue;
        if (p.start > e) { l_rep += e - b; b = p.start; e = p.end; }
        else e = e > p.end ? e : p.end;
    }
    l_rep += e - b;
    *frac_rep = (float)l_rep / v.read_len;
    // ---- seeds into chains (L1145-1191) ----
    Tree T;
    T.nodes = S.nodes; T.max_nodes = S.max_nodes; T.n_nodes = 0; T.n_keys = 0; T.hits = v.hits; T.ch = S.ch; T.overflow = false;
    T.root = T.new_node(0);
    int n_chains = 0;
    for (int i = 0; i < v.n_smems && !T.overflow; i++) {
        const meme_smem_t &p = v.smems[i];
        const int slen = p.end - p.start;
        const int64_t step = p.hitcount > o.max_occ ? p.hitcount / o.max_occ : 1;
        int count = 0;
        for (int64_t k = 0; k < p.hitcount && count < o.max_occ; k += step, ++count) {
            const cu32 sd = (cu32)(p.hitbeg + count);
            const int64_t rbeg = (int64_t)v.hits[sd];
            S.ql[sd] = (cu32)p.start | (cu32)slen << 16;
            S.next[sd] = CHN_NIL;
            const int rid = intv2rid(v, rbeg, rbeg + slen);
            if (rid < 0) continue;
            bool to_add = true;
            if (T.n_keys) {
                const cu32 lo = T.lower(rbeg);
                if (lo != CHN_NIL) {
                    // test_and_merge (L448-493)
                    Rec &c = S.ch[lo];
                    const cu32 ls = c.last_seed, fs = c.first_seed;
                    const int l_qb = (int)(S.ql[ls] & 0xffff), l_len = (int)(S.ql[ls] >> 16);
                    const int64_t l_rb = (int64_t)v.hits[ls], f_rb = (int64_t)v.hits[fs];
                    const int f_qb = (int)(S.ql[fs] & 0xffff);
                    const int64_t qend = l_qb + l_len, rend = l_rb + l_len;
                    if (rid != c.rid) to_add = true;
                    else if (p.start >= f_qb && p.start + slen <= qend && rbeg >= f_rb && rbeg + slen <= rend) to_add = false;   // contained
                    else if ((l_rb < v.l_pac || f_rb < v.l_pac) && rbeg >= v.l_pac) to_add = true;                             // other strand
                    else {
                        const int64_t x = p.start - l_qb, y = rbeg - l_rb;
                        if (y >= 0 && x - y <= o.w && y - x <= o.w && x - l_len < o.max_chain_gap && y - l_len < o.max_chain_gap) {
                            S.next[ls] = sd; c.last_seed = sd; c.n++;
                            to_add = false;
                        }
                    }
                }
            }
            if (to_add) {
                Rec &c = S.ch[n_chains];
                c.first_seed = c.last_seed = sd; c.n = 1; c.rid = rid; c.is_alt = v.contigs[rid].is_alt != 0;
                c.w = 0; c.first = -1; c.kept = 0;
                T.put((cu32)n_chains);
                n_chains++;
                if (T.overflow) break;
            }
        }
    }
    if (T.overflow) { *err = 1; return 0; }
    int n = T.traverse(S.order);
    if (n != n_chains) { *err = 1; return 0; }
    if (n == 0) return 0;
    // ---- mem_chain_flt (L599-717) ----
    int k = 0;
    for (int i = 0; i < n; i++) {
        Rec &c = S.ch[S.order[i]];
        c.first = -1; c.kept = 0;
        c.w = chain_weight(c, S, v.hits);
        if (c.w >= o.min_chain_weight) S.order[k++] = S.order[i];
    }
    // (if every chain is lighter than min_chain_weight the reference goes on with the stale first record, L624-641: one
    // "chain" survives.  Its seed array may already be freed there; here it is the first chain in tree order, intact.)
    n = k ? k : 1;
    if (!introsort((size_t)n, S.order, S)) { *err = 2; return 0; }
    cu32 *a = S.order;
    int n_kept = 0;
    S.ch[a[0]].kept = 3;
    S.kept[n_kept++] = 0;
#define CHN_BEG(c) ((int)(S.ql[(c).first_seed] & 0xffff))
#define CHN_END(c) ((int)(S.ql[(c).last_seed] & 0xffff) + (int)(S.ql[(c).last_seed] >> 16))
    for (int i = 1; i < n; ++i) {
        Rec &ci = S.ch[a[i]];
        int large_ovlp = 0;
        const int bi = CHN_BEG(ci), ei = CHN_END(ci);
        for (k = 0; k < n_kept; ++k) {
            const int j = (int)S.kept[k];
            Rec &cj = S.ch[a[j]];
            const int bj = CHN_BEG(cj), ej = CHN_END(cj);
            const int b_max = bj > bi ? bj : bi, e_min = ej < ei ? ej : ei;
            if (e_min > b_max && (!cj.is_alt || ci.is_alt)) {
                const int li = ei - bi, lj = ej - bj, min_l = li < lj ? li : lj;
                if ((float)(e_min - b_max) >= (float)min_l * o.mask_level && min_l < o.max_chain_gap) {
                    large_ovlp = 1;
                    if (cj.first < 0) cj.first = i;
                    if ((float)ci.w < (float)cj.w * o.drop_ratio && cj.w - ci.w >= o.min_seed_len << 1) break;
                }
            }
        }
        if (k == n_kept) {
            S.kept[n_kept++] = (cu32)i;
            ci.kept = large_ovlp ? 2 : 3;
        }
    }
    for (int i = 0; i < n_kept; ++i) {
        const Rec &c = S.ch[a[S.kept[i]]];
        if (c.first >= 0) S.ch[a[c.first]].kept = 1;
    }
    int i;
    for (i = k = 0; i < n; ++i) {
        const int kp = S.ch[a[i]].kept;
        if (kp == 0 || kp == 3) continue;
        if (++k >= o.max_chain_extend) break;
    }
    for (; i < n; ++i)
        if (S.ch[a[i]].kept < 3) S.ch[a[i]].kept = 0;
    int n_seeds = 0;
    for (i = k = 0; i < n; ++i)
        if (S.ch[a[i]].kept != 0) { n_seeds += S.ch[a[i]].n; a[k++] = a[i]; }
    *n_out = k;
    return n_seeds;
}

}  // namespace chn
