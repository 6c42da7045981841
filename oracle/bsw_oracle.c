/* oracle/bsw_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's banded affine-gap *extension* alignment
 * (kernel B of SURVEY.md §8a, row B1): BandedPairWiseSW::scalarBandedSWA,
 * /root/reference/src/bandedSWA.cpp:L116-237, and its batch wrapper L242-259.
 * The vector kernels (getScores8/16, L412-4870) and ksw_extend2 (src/ksw.cpp:L434-535)
 * compute the same numbers for the default gap-extension penalties.
 *
 * Parity pinning: tests/test_oracle_bsw.py runs the UNMODIFIED reference class through
 * oracle/ref_bsw_shim.cpp (scalar and 16-bit vector entry points) on seeded pair sets and
 * compares all six outputs; the resulting vectors are committed under tests/golden/.
 *
 * Semantics worth knowing (all reproduced below, each cites the reference line):
 *   - the score row carries H(i-1,j-1) in slot j and E(i,j); slots outside the current band
 *     keep whatever an earlier row left there (L176-197) -- "stale" reads are part of the contract;
 *   - M = H(i-1,j-1) ? H(i-1,j-1)+s : 0  (L184): a zero cell cannot be extended diagonally;
 *   - E/F open from M, not from H (L190-197);
 *   - band [beg,end) is clipped to |i-j|<=w, then shrunk to the non-zero cells of the row
 *     just computed (L218-221); w itself is first clipped by max_ins/max_del (L147-156);
 *   - row loop stops at an all-zero row (L206) or on z-drop (L210-216);
 *   - gscore/max_ie are updated only by rows that reach the query end (L202-205).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t idr, idq, id;
    int32_t len1, len2;     /* len1 = reference (target) length, len2 = query length */
    int32_t h0;
    int32_t seqid, regid;
    int32_t score, tle, gtle, qle;
    int32_t gscore, max_off;
} oracle_seqpair_t;         /* byte-identical to the reference's SeqPair, bandedSWA.h:L90-99 */

typedef struct {
    int32_t o_del, e_del, o_ins, e_ins, zdrop, end_bonus;
    int32_t a, b;           /* match score, mismatch penalty (positive) */
} oracle_bsw_opt_t;

/* bwa_fill_scmat, src/bwa.cpp:L262-271: 5x5, N scores -1 against everything */
static void fill_mat(int a, int b, int8_t mat[25])
{
    int k = 0;
    for (int i = 0; i < 4; i++) {
        for (int j = 0; j < 4; j++) mat[k++] = (int8_t)(i == j ? a : -b);
        mat[k++] = -1;
    }
    for (int j = 0; j < 5; j++) mat[k++] = -1;
}

uint64_t oracle_bsw_cells;  /* cells evaluated inside the band (for GCUPS accounting) */

int oracle_bsw_one(const oracle_bsw_opt_t *o, int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                   int w, int h0, int *qle, int *tle, int *gtle, int *gscore_out, int *max_off_out)
{
    int8_t mat[25];
    fill_mat(o->a, o->b, mat);
    const int oe_del = o->o_del + o->e_del, oe_ins = o->o_ins + o->e_ins;
    int32_t *H = (int32_t *)calloc((size_t)qlen + 2, sizeof(int32_t));   /* H(i-1, j-1) at slot j */
    int32_t *E = (int32_t *)calloc((size_t)qlen + 2, sizeof(int32_t));   /* E(i, j) at slot j      */

    /* row "-1": L143-145 */
    H[0] = h0;
    H[1] = h0 > oe_ins ? h0 - oe_ins : 0;
    for (int j = 2; j <= qlen && H[j - 1] > o->e_ins; j++) H[j] = H[j - 1] - o->e_ins;

    /* clip the band by the longest gap the scores could pay for: L147-156 */
    int mx = 0;
    for (int i = 0; i < 25; i++) if (mat[i] > mx) mx = mat[i];
    int max_ins = (int)((double)(qlen * mx + o->end_bonus - o->o_ins) / o->e_ins + 1.);
    if (max_ins < 1) max_ins = 1;
    if (w > max_ins) w = max_ins;
    int max_del = (int)((double)(qlen * mx + o->end_bonus - o->o_del) / o->e_del + 1.);
    if (max_del < 1) max_del = 1;
    if (w > max_del) w = max_del;

    int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen;
    for (int i = 0; i < tlen; i++) {
        const int8_t *srow = &mat[target[i] * 5];
        int f = 0, rowmax = 0, rowmax_j = -1, hleft;
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        if (beg == 0) {                                 /* first column: L163-166 */
            hleft = h0 - (o->o_del + o->e_del * (i + 1));
            if (hleft < 0) hleft = 0;
        } else hleft = 0;
        int j;
        for (j = beg; j < end; j++) {
            int diag = H[j], e = E[j];
            H[j] = hleft;                               /* becomes H(i, j-1) for the next row */
            int M = diag ? diag + srow[query[j]] : 0;
            int h = M > e ? M : e;
            if (f > h) h = f;
            hleft = h;
            if (h >= rowmax) { rowmax = h; rowmax_j = j; }  /* `mj = m > h ? mj : j`: the LAST maximum wins, L188-189 */
            int t = M - oe_del; if (t < 0) t = 0;
            e -= o->e_del; if (t > e) e = t;
            E[j] = e;
            t = M - oe_ins; if (t < 0) t = 0;
            f -= o->e_ins; if (t > f) f = t;
            oracle_bsw_cells++;
        }
        H[end] = hleft; E[end] = 0;                     /* L199 */
        if (j == qlen) {                                /* L200-203 */
            if (gscore <= hleft) best_ie = i;           /* max_ie = gscore > h1 ? max_ie : i */
            if (gscore < hleft) gscore = hleft;
        }
        if (rowmax == 0) break;
        if (rowmax > best) {
            best = rowmax; best_i = i; best_j = rowmax_j;
            int d = rowmax_j - i; if (d < 0) d = -d;
            if (d > max_off) max_off = d;
        } else if (o->zdrop > 0) {                      /* L210-216 */
            int di = i - best_i, dj = rowmax_j - best_j;
            if (di > dj) { if (best - rowmax - (di - dj) * o->e_del > o->zdrop) break; }
            else         { if (best - rowmax - (dj - di) * o->e_ins > o->zdrop) break; }
        }
        /* shrink the band to the live cells: L218-221 */
        for (j = beg; j < end && H[j] == 0 && E[j] == 0; j++);
        beg = j;
        for (j = end; j >= beg && H[j] == 0 && E[j] == 0; j--);
        end = j + 2 < qlen ? j + 2 : qlen;
    }
    free(H); free(E);
    if (qle) *qle = best_j + 1;
    if (tle) *tle = best_i + 1;
    if (gtle) *gtle = best_ie + 1;
    if (gscore_out) *gscore_out = gscore;
    if (max_off_out) *max_off_out = max_off;
    return best;
}

/* scalarBandedSWAWrapper, bandedSWA.cpp:L242-259 */
void oracle_bsw_batch(const oracle_bsw_opt_t *o, oracle_seqpair_t *pairs, const uint8_t *ref, const uint8_t *qer,
                      int npairs, int w)
{
    for (int i = 0; i < npairs; i++) {
        oracle_seqpair_t *p = &pairs[i];
        p->score = oracle_bsw_one(o, p->len2, qer + p->idq, p->len1, ref + p->idr, w, p->h0,
                                  &p->qle, &p->tle, &p->gtle, &p->gscore, &p->max_off);
    }
}
