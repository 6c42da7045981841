/* oracle/cigar_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the compute of the reference's SAM stage: what mem_reg2aln works out
 * for a region before the record is formatted -- global alignment score, CIGAR, NM and the MD string.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the product path never does.
 *
 * Parity pinning: tests/test_oracle_cigar.py compares this file's output with the reference's own mem_reg2aln
 * (oracle/ref_cigar_shim.cpp links the unmodified reference objects and calls it case by case; its outputs are committed
 * as tests/golden/cigar_*.json and regenerated live whenever oracle/_ref exists).
 *
 * What is restated, and where it lives in the reference (paths under /root/reference):
 *   infer_w()        src/bwamem.cpp:L2393-2400, L2334-2338   infer_bw and how mem_reg2aln combines and limits the two values
 *   nw_banded()      src/ksw.cpp:L545-670                    push_cigar, ksw_global2
 *   gen_cigar()      src/bwa.cpp:L274-361                    bwa_gen_cigar2 (band, strand reversal, NM, MD)
 *   region_cigar()   src/bwamem.cpp:L2339-2348               the band-doubling loop of mem_reg2aln
 * The reference segment comes from the .0123 text (forward strand followed by its reverse complement, one code per byte),
 * which is what bns_get_seq unpacks from .pac for the same coordinates.
 *
 * Storage follows the reference (unpacked sequences copied and reversed in place, a query profile, a CIGAR that is built
 * backwards and then reversed, a kstring-like MD), deliberately not the layout of the CUDA kernel.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

typedef struct { int64_t rb, re, q_off; int32_t qlen, w, truesc, pad; } task_t;
typedef struct { int8_t mat[25]; int8_t pad[3]; int32_t o_del, e_del, o_ins, e_ins, w, a; } opt_t;
typedef struct { int32_t score, nm, n_cigar, md_len; int64_t cigar_off, md_off; } rec_t;

#define NEG_INF (-0x40000000)

int oracle_infer_w(int l_query_seg, int l_ref_seg, int truesc, int a, int o_del, int e_del, int o_ins, int e_ins, int opt_w, int reg_w)
{
    int k, w2 = 0;
    for (k = 0; k < 2; k++) {
        int q = k ? o_ins : o_del, r = k ? e_ins : e_del, l1 = l_query_seg, l2 = l_ref_seg, w;
        if (l1 == l2 && l1 * a - truesc < (q + r - a) << 1) w = 0;
        else {
            w = ((double)((l1 < l2 ? l1 : l2) * a - truesc - q) / r + 2.);
            if (w < abs(l1 - l2)) w = abs(l1 - l2);
        }
        w2 = w2 > w ? w2 : w;
    }
    if (w2 > opt_w) w2 = w2 < reg_w ? w2 : reg_w;
    return w2;
}

typedef struct { int h, e; } cell_t;

static uint32_t *cigar_push(int *n, int *m, uint32_t *cigar, int op, int len)
{
    if (*n == 0 || op != (int)(cigar[*n - 1] & 0xf)) {
        if (*n == *m) { *m = *m ? *m << 1 : 4; cigar = (uint32_t *)realloc(cigar, (size_t)*m << 2); }
        cigar[(*n)++] = (uint32_t)len << 4 | op;
    } else cigar[*n - 1] += (uint32_t)len << 4;
    return cigar;
}

static int nw_banded(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins,
                     int e_ins, int w, int *n_cigar_, uint32_t **cigar_)
{
    const int m = 5;
    int i, j, k, oe_del = o_del + e_del, oe_ins = o_ins + e_ins, score, n_col;
    cell_t *eh;
    int8_t *qp;
    uint8_t *z;
    n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    z = (uint8_t *)malloc((size_t)n_col * tlen + 1);
    qp = (int8_t *)malloc((size_t)qlen * m);
    eh = (cell_t *)calloc(qlen + 1, sizeof(cell_t));
    for (k = i = 0; k < m; ++k) {
        const int8_t *p = &mat[k * m];
        for (j = 0; j < qlen; ++j) qp[i++] = p[query[j]];
    }
    eh[0].h = 0; eh[0].e = NEG_INF;
    for (j = 1; j <= qlen && j <= w; ++j) eh[j].h = -(o_ins + e_ins * j), eh[j].e = NEG_INF;
    for (; j <= qlen; ++j) eh[j].h = eh[j].e = NEG_INF;
    for (i = 0; i < tlen; ++i) {
        int f = NEG_INF, h1, beg, end, t;
        int8_t *q = &qp[target[i] * qlen];
        uint8_t *zi = &z[(long)i * n_col];
        beg = i > w ? i - w : 0;
        end = i + w + 1 < qlen ? i + w + 1 : qlen;
        h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : NEG_INF;
        for (j = beg; j < end; ++j) {
            cell_t *p = &eh[j];
            int h, mm = p->h, e = p->e;
            uint8_t d;
            p->h = h1;
            mm += q[j];
            d = mm >= e ? 0 : 1;
            h = mm >= e ? mm : e;
            d = h >= f ? d : 2;
            h = h >= f ? h : f;
            h1 = h;
            t = mm - oe_del;
            e -= e_del;
            d |= e > t ? 1 << 2 : 0;
            e = e > t ? e : t;
            p->e = e;
            t = mm - oe_ins;
            f -= e_ins;
            d |= f > t ? 2 << 4 : 0;
            f = f > t ? f : t;
            zi[j - beg] = d;
        }
        eh[end].h = h1; eh[end].e = NEG_INF;
    }
    score = eh[qlen].h;
    {
        int n_cigar = 0, m_cigar = 0, which = 0;
        uint32_t *cigar = 0, tmp;
        i = tlen - 1; k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
        while (i >= 0 && k >= 0) {
            which = z[(long)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
            if (which == 0) cigar = cigar_push(&n_cigar, &m_cigar, cigar, 0, 1), --i, --k;
            else if (which == 1) cigar = cigar_push(&n_cigar, &m_cigar, cigar, 2, 1), --i;
            else cigar = cigar_push(&n_cigar, &m_cigar, cigar, 1, 1), --k;
        }
        if (i >= 0) cigar = cigar_push(&n_cigar, &m_cigar, cigar, 2, i + 1);
        if (k >= 0) cigar = cigar_push(&n_cigar, &m_cigar, cigar, 1, k + 1);
        for (i = 0; i < n_cigar >> 1; ++i) tmp = cigar[i], cigar[i] = cigar[n_cigar - 1 - i], cigar[n_cigar - 1 - i] = tmp;
        *n_cigar_ = n_cigar; *cigar_ = cigar;
    }
    free(eh); free(qp); free(z);
    return score;
}

typedef struct { char *s; size_t l, m; } str_t;
static void str_putc(str_t *s, int c) { if (s->l + 2 > s->m) { s->m = s->m ? s->m << 1 : 32; s->s = (char *)realloc(s->s, s->m); } s->s[s->l++] = (char)c; s->s[s->l] = 0; }
static void str_putw(str_t *s, int u) { char b[16]; int l = 0, x; do { b[l++] = '0' + u % 10; u /= 10; } while (u > 0); for (x = l - 1; x >= 0; --x) str_putc(s, b[x]); }

/* One call of bwa_gen_cigar2; returns 0 if rejected.  md receives a malloc'ed string. */
static int gen_cigar(const opt_t *o, int w_, int64_t l_pac, const uint8_t *text, int l_query, const uint8_t *query_in, int64_t rb, int64_t re,
                     int *score, int *n_cigar, uint32_t **cigar_out, int *NM, char **md)
{
    uint8_t *query, *rseq, tmp;
    uint32_t *cigar = 0;
    int64_t rlen;
    int i;
    *n_cigar = 0; *NM = -1; *cigar_out = 0; *md = 0;
    if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return 0;
    if (rb < 0 || re > l_pac << 1) return 0;             /* bns_get_seq clips the range: re - rb != rlen */
    rlen = re - rb;
    rseq = (uint8_t *)malloc(rlen);
    memcpy(rseq, text + rb, rlen);
    query = (uint8_t *)malloc(l_query);
    memcpy(query, query_in, l_query);
    if (rb >= l_pac) {
        for (i = 0; i < l_query >> 1; ++i) tmp = query[i], query[i] = query[l_query - 1 - i], query[l_query - 1 - i] = tmp;
        for (i = 0; i < rlen >> 1; ++i) tmp = rseq[i], rseq[i] = rseq[rlen - 1 - i], rseq[rlen - 1 - i] = tmp;
    }
    if (l_query == re - rb && w_ == 0) {
        cigar = (uint32_t *)malloc(4);
        cigar[0] = (uint32_t)l_query << 4 | 0;
        *n_cigar = 1;
        for (i = 0, *score = 0; i < l_query; ++i) *score += o->mat[rseq[i] * 5 + query[i]];
    } else {
        int w, max_gap, max_ins, max_del, min_w;
        max_ins = (int)((double)(((l_query + 1) >> 1) * o->mat[0] - o->o_ins) / o->e_ins + 1.);
        max_del = (int)((double)(((l_query + 1) >> 1) * o->mat[0] - o->o_del) / o->e_del + 1.);
        max_gap = max_ins > max_del ? max_ins : max_del;
        max_gap = max_gap > 1 ? max_gap : 1;
        w = (max_gap + abs((int)(rlen - l_query)) + 1) >> 1;
        w = w < w_ ? w : w_;
        min_w = abs((int)(rlen - l_query)) + 3;
        w = w > min_w ? w : min_w;
        *score = nw_banded(l_query, query, (int)rlen, rseq, o->mat, o->o_del, o->e_del, o->o_ins, o->e_ins, w, n_cigar, &cigar);
    }
    {
        int k, x, y, u, n_mm = 0, n_gap = 0;
        str_t str = {0, 0, 0};
        const char *int2base = rb < l_pac ? "ACGTN" : "TGCAN";
        for (k = 0, x = y = u = 0; k < *n_cigar; ++k) {
            int op = cigar[k] & 0xf, len = cigar[k] >> 4;
            if (op == 0) {
                for (i = 0; i < len; ++i) {
                    if (query[x + i] != rseq[y + i]) { str_putw(&str, u); str_putc(&str, int2base[rseq[y + i]]); ++n_mm; u = 0; }
                    else ++u;
                }
                x += len; y += len;
            } else if (op == 2) {
                if (k > 0 && k < *n_cigar - 1) {
                    str_putw(&str, u); str_putc(&str, '^');
                    for (i = 0; i < len; ++i) str_putc(&str, int2base[rseq[y + i]]);
                    u = 0; n_gap += len;
                }
                y += len;
            } else if (op == 1) x += len, n_gap += len;
        }
        str_putw(&str, u);
        *NM = n_mm + n_gap;
        *md = str.s;
    }
    free(query); free(rseq);
    *cigar_out = cigar;
    return 1;
}

typedef struct {
    int64_t n_tasks, n_ops, n_md;
    rec_t *recs; uint32_t *cigars; char *md;
} result_t;

/* The loop of mem_reg2aln for every task.  text = .0123 (2 * l_pac codes), codes = the reads' base codes. */
result_t *oracle_cigar_regions(const uint8_t *text, int64_t l_pac, const uint8_t *codes, const task_t *tasks, int64_t n, const opt_t *o)
{
    result_t *R = (result_t *)calloc(1, sizeof(result_t));
    size_t cap_ops = 1024, cap_md = 4096;
    int64_t t;
    R->n_tasks = n;
    R->recs = (rec_t *)calloc((size_t)n + 1, sizeof(rec_t));
    R->cigars = (uint32_t *)malloc(cap_ops * 4);
    R->md = (char *)malloc(cap_md);
    for (t = 0; t < n; t++) {
        const task_t *k = &tasks[t];
        int i = 0, w2 = k->w, score = 0, last_sc = -(1 << 30), n_cigar = 0, NM = -1, ok = 0;
        uint32_t *cigar = 0;
        char *md = 0;
        rec_t *r = &R->recs[t];
        do {
            free(cigar); free(md);
            w2 = w2 < o->w << 2 ? w2 : o->w << 2;
            ok = gen_cigar(o, w2, l_pac, text, k->qlen, codes + k->q_off, k->rb, k->re, &score, &n_cigar, &cigar, &NM, &md);
            if (!ok) break;
            if (score == last_sc || w2 == o->w << 2) break;
            last_sc = score;
            w2 <<= 1;
        } while (++i < 3 && score < k->truesc - o->a);
        r->cigar_off = R->n_ops; r->md_off = R->n_md;
        if (!ok) { r->score = 0; r->nm = -1; r->n_cigar = -1; r->md_len = 0; continue; }
        r->score = score; r->nm = NM; r->n_cigar = n_cigar; r->md_len = (int)strlen(md) + 1;
        while (R->n_ops + n_cigar > (int64_t)cap_ops) { cap_ops *= 2; R->cigars = (uint32_t *)realloc(R->cigars, cap_ops * 4); }
        while (R->n_md + r->md_len > (int64_t)cap_md) { cap_md *= 2; R->md = (char *)realloc(R->md, cap_md); }
        memcpy(R->cigars + R->n_ops, cigar, (size_t)n_cigar * 4);
        memcpy(R->md + R->n_md, md, r->md_len);
        R->n_ops += n_cigar; R->n_md += r->md_len;
        free(cigar); free(md);
    }
    return R;
}

void oracle_cigar_result_free(result_t *R)
{
    if (!R) return;
    free(R->recs); free(R->cigars); free(R->md); free(R);
}
