/* oracle/matesw_oracle.c -- TEST INFRASTRUCTURE ONLY: never linked, imported or executed by the product path.
 *
 * CPU restatement of the reference's mate-rescue Smith-Waterman batch step, one pair at a time:
 *   mem_sam_pe_batch          src/bwamem_pair.cpp:L719-817   phase 0 for every pair, selection, reversal, phase 1
 *   kswv::kswvBatchWrapper8   src/kswv.cpp:L177-368          how a lane's two sequences are coded and padded (8-bit class)
 *   kswv::kswv512_u8          src/kswv.cpp:L372-712          one lane of the 64-lane unsigned 8-bit kernel
 *   kswv::kswvBatchWrapper16  src/kswv.cpp:L726-930          the same for the 16-bit class
 *   kswv::kswv512_16          src/kswv.cpp:L934-1218         one lane of the 32-lane signed 16-bit kernel
 * The vector kernels run 64 / 32 pairs in lockstep; what one lane computes does not depend on its neighbours (rows past a
 * lane's own reference are padding whose scores only decay; the batch-wide loop bounds maxl / minh / limit are re-masked per
 * lane), which tests/test_oracle_matesw.py checks by running the reference on the same pairs in differently cut batches.
 * Pinned against outputs of the reference itself (oracle/ref_matesw_shim.cpp -> tests/golden/matesw_ref.bin, live re-check
 * where /root/reference is mounted). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define XBYTE  0x10000
#define XSTOP  0x20000
#define XSUBO  0x40000
#define XSTART 0x80000

typedef struct { int32_t score, te, qe, score2, te2, tb, qb; } kswr;          /* kswr_t, src/ksw.h:L45-50 */
typedef struct { int32_t o_del, e_del, o_ins, e_ins, a, b; } mopt;            /* b: mismatch penalty, positive */
typedef struct { int64_t r_off, q_off; int32_t len1, len2, xtra, pad; } mrec;

typedef struct { int score, te, qe, score2, te2; } pass_t;

static int addu8(int a, int b) { int s = a + b; return s > 255 ? 255 : s; }   /* _mm512_adds_epu8 */
static int subu8(int a, int b) { int s = a - b; return s < 0 ? 0 : s; }       /* _mm512_subs_epu8 */
static int max2(int a, int b) { return a > b ? a : b; }

/* One lane of kswv512_u8 (byte = 1) or kswv512_16 (byte = 0), phase 0 or 1.  r / q: raw codes 0..4.  rowmax: len1 + 1 cells. */
static void lane(const uint8_t *r, int len1, const uint8_t *q, int len2, int xtra, const mopt *o, int byte, int phase, pass_t *out,
                 int *rowmax)
{
    const int w_match = o->a, w_mismatch = -o->b, w_ambig = -1;                /* kswv ctor L114-131, DEFAULT_AMBIG */
    const int unit = byte ? 16 : 8;
    const int cols = (len2 + unit - 1) / unit * unit;                          /* "quanta": L291-293 / L857-858 */
    const int INV = byte ? 0 : -1;                                             /* what a suppressed row maximum is stored as */
    int shift = 0, tab[32];
    memset(tab, 0, sizeof tab);
    if (byte) {                                                                /* L394-416 */
        int mn = w_match < w_mismatch ? w_match : w_mismatch;
        mn = mn < w_ambig ? mn : w_ambig;
        shift = 256 - (int)(uint8_t)(int8_t)mn;
        tab[0] = w_match;
        tab[1] = tab[2] = tab[3] = w_mismatch;
        for (int k = 4; k <= 12; k++) tab[k] = w_ambig;
        for (int k = 0; k < 16; k++) tab[k] = (int)(uint8_t)(int8_t)(tab[k] + shift);
    } else {                                                                   /* L968-972 */
        tab[0] = w_match;
        tab[1] = tab[2] = tab[3] = w_mismatch;
        for (int k = 12; k <= 19; k++) tab[k] = w_ambig;
        tab[31] = w_ambig;
    }
    const int lim = byte ? 255 : 32767;
    int v = (xtra & XSUBO) ? (xtra & 0xffff) : 0x10000;                        /* L425-441 / L980-997 */
    const int has_minsc = v <= lim, minsc = v;
    v = (xtra & XSTOP) ? (xtra & 0xffff) : 0x10000;
    const int has_endsc = v <= lim, endsc = v;
    const int oe_del = o->o_del + o->e_del, oe_ins = o->o_ins + o->e_ins;
    int *H0 = (int *)calloc((size_t)cols + 2, sizeof(int)), *H1 = (int *)calloc((size_t)cols + 2, sizeof(int));
    int *F = (int *)calloc((size_t)cols + 2, sizeof(int));
    int gmax = 0, te = -1, qe = 0, alive = 1, mask = 0, minsc_ok = 0, pimax = 0, i;
    for (i = 0; i < len1; i++) {
        int e = 0, imax = 0, iqe = -1;
        const int s1 = byte ? (r[i] == 4 ? 4 : r[i]) : (r[i] == 4 ? 15 : r[i]);          /* AMBR / AMBR16 */
        for (int j = 0; j < cols; j++) {
            const int s2 = j < len2 ? (byte ? (q[j] == 4 ? 8 : q[j]) : (q[j] == 4 ? 16 : q[j])) : (byte ? 5 : 26);   /* AMBQ, DUMMY5 / AMBQ16, DUMMY3 */
            const int x = s1 ^ s2, h00 = H0[j], f11 = F[j + 1];
            int m, h, f21;
            if (byte) {                                                        /* MAIN_SAM_CODE8_OPT, L63-85 */
                int sbt = tab[x & 15];
                if (s2 == 5) sbt = shift;
                m = subu8(addu8(h00, sbt), shift);
                h = max2(max2(m, e), f11);
                if (h > imax) { imax = h; iqe = j & 0xff; }
                e = max2(subu8(h, oe_ins), subu8(e, o->e_ins));
                f21 = max2(subu8(h, oe_del), subu8(f11, o->e_del));
            } else {                                                           /* MAIN_SAM_CODE16_OPT, L87-108 */
                m = h00 + tab[x & 31];
                h = max2(max2(max2(m, e), f11), 0);
                if (h > imax) { imax = h; iqe = j; }
                e = max2(h - oe_ins, e - o->e_ins);
                f21 = max2(h - oe_del, f11 - o->e_del);
            }
            H1[j + 1] = h;
            F[j + 1] = f21;
        }
        if (i > 0) {                                                           /* Block I, L510-522 / L1063-1078 */
            const int msk = imax > pimax || mask;
            rowmax[i - 1] = (!msk && minsc_ok && alive) ? pimax : INV;
            mask = !msk;
        }
        pimax = imax;
        minsc_ok = has_minsc && imax >= minsc;
        if (alive && imax > gmax) { gmax = imax; te = i; qe = iqe; }          /* Block II */
        int stop = has_endsc && gmax >= endsc;
        if (byte && addu8(gmax, shift) >= 255) stop = 1;
        if (stop) { alive = 0; i++; break; }
        int *S = H1; H1 = H0; H0 = S;
    }
    if (i > 0) rowmax[i - 1] = (!mask && minsc_ok && alive) ? pimax : INV;     /* L551-554 / L1099-1102 */
    const int rows = i;
    free(H0); free(H1); free(F);
    out->te = te; out->qe = qe; out->score2 = -1; out->te2 = -1;
    if (phase) { out->score = gmax; return; }                                  /* L576-581: raw score, te, qe */
    int live = 1;
    if (byte) { out->score = gmax + shift < 255 ? gmax : 255; live = out->score != 255; } else out->score = gmax;
    /* score2 / te2: the best row maximum at least `val` rows away from te (L611-700 / L1143-1206) */
    const int qmax = max2(max2(w_match, w_mismatch), w_ambig);
    const int val = (gmax + qmax - 1) / qmax, low = te - val, high = te + val;
    int cur = byte ? 0 : -1, te2 = -1;
    for (int k = 0; k < low && k < rows; k++) if (rowmax[k] > cur) { cur = rowmax[k]; te2 = k; }
    for (int k = high + 1 > 0 ? high + 1 : 0; k < rows && k < len1; k++) if (rowmax[k] > cur) { cur = rowmax[k]; te2 = k; }
    if (byte) { if (live) { out->score2 = cur == 0 ? -1 : cur; out->te2 = te2; } }
    else { out->score2 = cur; out->te2 = te2; }
}

static void rev(uint8_t *s, int l) { for (int i = 0; i < l >> 1; i++) { uint8_t t = s[i]; s[i] = s[l - 1 - i]; s[l - 1 - i] = t; } }

/* mem_sam_pe_batch for one pair */
void oracle_matesw_pair(const uint8_t *ref, int len1, const uint8_t *qer, int len2, int xtra, const mopt *o, kswr *a)
{
    const int byte = (xtra & XBYTE) != 0;                                      /* sort_classify, src/bwamem.cpp:L1797-1826 */
    int *rowmax = (int *)malloc(((size_t)len1 + 2) * sizeof(int));
    pass_t p;
    a->tb = a->qb = -1;                                                        /* L745-748 */
    lane(ref, len1, qer, len2, xtra, o, byte, 0, &p, rowmax);
    a->score = p.score; a->te = p.te; a->qe = p.qe; a->score2 = p.score2; a->te2 = p.te2;
    if ((xtra & XSTART) != 0 && !((xtra & XSUBO) && a->score < (xtra & 0xffff))) {      /* L769-784 */
        uint8_t *r2 = (uint8_t *)malloc((size_t)len1 + 1), *q2 = (uint8_t *)malloc((size_t)len2 + 1);
        memcpy(r2, ref, (size_t)len1); memcpy(q2, qer, (size_t)len2);
        rev(q2, a->qe + 1); rev(r2, a->te + 1);
        lane(r2, len1, q2, a->qe + 1, XSTOP | a->score, o, byte, 1, &p, rowmax);
        if (a->score == p.score) { a->tb = a->te - p.te; a->qb = a->qe - p.qe; }        /* L578-581 */
        free(r2); free(q2);
    }
    free(rowmax);
}

void oracle_matesw(const mrec *recs, int64_t n, const uint8_t *ref, const uint8_t *qer, const mopt *o, kswr *out)
{
    for (int64_t k = 0; k < n; k++)
        oracle_matesw_pair(ref + recs[k].r_off, recs[k].len1, qer + recs[k].q_off, recs[k].len2, recs[k].xtra, o, &out[k]);
}
