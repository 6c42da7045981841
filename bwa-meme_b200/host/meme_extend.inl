// meme_extend.inl -- chunk-level seed extension: the host side of kernel B for a whole chunk of reads.
//
// Included at the end of the (patched) reference src/bwamem.cpp, so that it can call that file's own helpers
// (cal_max_gap, mem_sort_dedup_patch[_mate_sort], the mem_* types).  It stands where the reference runs
//     kt_for(worker_aln) -> mem_kernel2_core -> mem_chain2aln_across_reads_V2        src/bwamem.cpp:L1945, L1633-1723, L2573-3490
// once per 512-read block and host thread, and does the same work for the WHOLE chunk in three steps:
//   1. every host thread walks its share of the reads and writes the left / right extension pairs of all their seeds
//      into its own pinned arena (alignment-region bookkeeping as in L2612-2934: reference window per chain, seeds in
//      score order, reversed copies for the left side);
//      (with the index resident on the GPU the reference side of a pair is just a position in that text: no bytes copied
//      or uploaded);
//   2. ONE device call per side (meme_glue_bsw_segments -> meme_bsw_upload_segments / meme_bsw_run_ext): the band
//      retries (MAX_BAND_TRY, w then 2w, L2968-3376) are decided by the kernel, the right side starts from the scores
//      the left side produced (h0 = region score, L3168-3173);
//   3. results are folded into the regions (clip decision with pen_clip5/3, seed coverage), contained seeds are purged
//      (L3391-3485), and the tail of mem_kernel2_core runs per read (chains freed, empty regions dropped,
//      mem_sort_dedup_patch, ALT marking, L1666-1719).
// The 8-bit / 16-bit / scalar classing of the reference (L2804-2812) is a CPU SIMD artefact and does not exist here.
// Off unless MEME_B200_BSW=1; SAM output is identical either way (tests/test_e2e_sam_gpu.py).
#include <algorithm>
#include <vector>
#include <omp.h>
#include <thread>

namespace meme_ext {

// pairs and sequences of one side, written by one host thread; pinned, kept across chunks
struct Arena {
    SeqPair *pairs = nullptr; int64_t n = 0, cap = 0;
    uint8_t *ref = nullptr; int64_t n_ref = 0, cap_ref = 0;
    uint8_t *qer = nullptr; int64_t n_qer = 0, cap_qer = 0;
    int32_t *w_used = nullptr; int64_t cap_w = 0;

    template <class T> static void grow(T *&p, int64_t &cap, int64_t keep, int64_t need)
    {
        if (need <= cap) return;
        int64_t ncap = std::max<int64_t>(need + need / 2, 1 << 16);
        T *q = (T *)meme_glue_pinned_alloc((size_t)ncap * sizeof(T));
        if (p) { if (keep) memcpy(q, p, (size_t)keep * sizeof(T)); meme_glue_pinned_free(p); }
        p = q; cap = ncap;
    }
    template <class T> static void reserve(T *&p, int64_t &cap, int64_t count)
    {   // exact size (no growth margin): matches a pre-pinned block of the glue's cache
        if (p || count <= cap) return;
        p = (T *)meme_glue_pinned_alloc((size_t)count * sizeof(T));
        cap = count;
    }
    void reset() { n = n_ref = n_qer = 0; }
    SeqPair *new_pair() { grow(pairs, cap, n, n + 1); return &pairs[n++]; }
    uint8_t *take_ref(int64_t len) { grow(ref, cap_ref, n_ref, n_ref + len + 1); uint8_t *p = ref + n_ref; n_ref += len; return p; }
    uint8_t *take_qer(int64_t len) { grow(qer, cap_qer, n_qer, n_qer + len + 1); uint8_t *p = qer + n_qer; n_qer += len; return p; }
};

struct ThreadState {
    Arena left[2], right[2];            // [half of the chunk]: the second half's pairs are built while the first half's run on the GPU
    std::vector<uint32_t> srt;          // seeds of every chain in ascending score order, read after read
    std::vector<uint64_t> key;
};

static std::vector<ThreadState> g_ts;
static std::vector<int64_t> g_srt_off;  // per read: offset of its first chain in its thread's srt
static std::vector<int> g_blk_owner;    // per block of reads: the host thread that built its pairs (owner of its srt slice)

static inline void seed_coverage(mem_alnreg_t *a)
{   // seeds of the chain fully inside the region (L3003-3013)
    if (a->rb != H0_ && a->qb != H0_ && a->qe != H0_ && a->re != H0_) {
        int cov = 0;
        for (int i = 0; i < a->c->n; ++i) {
            const mem_seed_t *t = &a->c->seeds[i];
            if (t->qbeg >= a->qb && t->qbeg + t->len <= a->qe && t->rbeg >= a->rb && t->rbeg + t->len <= a->re) cov += t->len;
        }
        a->seedcov = cov;
    }
}

// step 1 for one read: regions and extension pairs of all seeds of all chains
static void collect_read(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, uint8_t *ref_string, const bseq1_t *seq,
                         int l, mem_chain_v *chn, mem_alnreg_v *av, ThreadState &T, const int half, const bool by_position)
{
    Arena &AL = T.left[half], &AR = T.right[half];
    const int64_t l_pac = bns->l_pac;
    const uint8_t *query = (const uint8_t *)seq->seq;
    const int l_query = seq->l_seq;
    kv_init(*av);
    size_t total = 0;
    for (size_t j = 0; j < chn->n; j++) total += chn->a[j].n;
    av->m = total;
    av->a = (mem_alnreg_t *)calloc(av->m, sizeof(mem_alnreg_t));
    for (size_t j = 0; j < chn->n; j++) {
        mem_chain_t *c = &chn->a[j];
        if (c->n == 0) continue;
        // the window of the reference any extension of this chain can reach (L2649-2668)
        int64_t rmax0 = l_pac << 1, rmax1 = 0;
        for (int i = 0; i < c->n; ++i) {
            const mem_seed_t *t = &c->seeds[i];
            int64_t b = t->rbeg - (t->qbeg + cal_max_gap(opt, t->qbeg));
            int64_t e = t->rbeg + t->len + ((l_query - t->qbeg - t->len) + cal_max_gap(opt, l_query - t->qbeg - t->len));
            if (b < rmax0) rmax0 = b;
            if (e > rmax1) rmax1 = e;
        }
        if (rmax0 < 0) rmax0 = 0;
        if (rmax1 > l_pac << 1) rmax1 = l_pac << 1;
        if (rmax0 < l_pac && l_pac < rmax1) {          // never across the forward / reverse boundary
            if (c->seeds[0].rbeg < l_pac) rmax1 = l_pac; else rmax0 = l_pac;
        }
        int rid = 0;
        const uint8_t *rseq = bns_fetch_seq_v2(bns, pac, &rmax0, c->seeds[0].rbeg, &rmax1, &rid, ref_string, nullptr);
        assert(c->rid == rid);
        // seeds by score, best first (keys are unique: any sort gives the reference's order, L2693-2696)
        T.key.resize((size_t)c->n);
        for (int i = 0; i < c->n; ++i) T.key[(size_t)i] = (uint64_t)c->seeds[i].score << 32 | (uint32_t)i;
        std::sort(T.key.begin(), T.key.end());
        for (int i = 0; i < c->n; ++i) T.srt.push_back((uint32_t)T.key[(size_t)i]);
        for (int k = c->n - 1; k >= 0; k--) {
            mem_seed_t *s = &c->seeds[(uint32_t)T.key[(size_t)k]];
            mem_alnreg_t *a = &av->a[av->n++];
            memset(a, 0, sizeof(mem_alnreg_t));
            s->aln = (int)av->n - 1;
            a->w = opt->w;
            a->score = a->truesc = -1;
            a->rid = c->rid;
            a->frac_rep = c->frac_rep;
            a->seedlen0 = s->len;
            a->c = c;
            a->rb = a->qb = a->re = a->qe = H0_;
            if (s->qbeg) {                              // left extension: reversed query prefix against reversed reference
                SeqPair *sp = AL.new_pair();
                memset(sp, 0, sizeof(SeqPair));
                const int64_t tl = s->rbeg - rmax0;
                sp->h0 = s->len * opt->a;
                sp->seqid = l; sp->regid = (int)av->n - 1;
                sp->len2 = s->qbeg; sp->len1 = (int32_t)tl;
                sp->idq = (int32_t)AL.n_qer;
                uint8_t *qs = AL.take_qer(s->qbeg);
                for (int i = 0; i < s->qbeg; ++i) qs[i] = query[s->qbeg - 1 - i];
                if (by_position) {                      // the reversed reference = the text read backwards from rbeg - 1
                    const uint64_t p = tl > 0 ? (uint64_t)(s->rbeg - 1) : 0;
                    sp->idr = (int32_t)(uint32_t)p; sp->id = (int32_t)((uint32_t)(p >> 32) | 0x80000000u);
                } else {
                    sp->idr = (int32_t)AL.n_ref;
                    uint8_t *rs = AL.take_ref(tl);
                    for (int64_t i = 0; i < tl; ++i) rs[i] = rseq[tl - 1 - i];
                }
                a->qb = s->qbeg; a->rb = s->rbeg;
            } else {
                a->score = a->truesc = s->len * opt->a; a->qb = 0; a->rb = s->rbeg;
            }
            if (s->qbeg + s->len != l_query) {          // right extension
                const int64_t qe = s->qbeg + s->len, re = s->rbeg + s->len - rmax0;
                assert(re >= 0);
                SeqPair *sp = AR.new_pair();
                memset(sp, 0, sizeof(SeqPair));
                sp->h0 = H0_;                           // known after the left side
                sp->seqid = l; sp->regid = (int)av->n - 1;
                sp->len2 = (int32_t)(l_query - qe); sp->len1 = (int32_t)(rmax1 - rmax0 - re);
                sp->idq = (int32_t)AR.n_qer;
                memcpy(AR.take_qer(sp->len2), query + qe, (size_t)sp->len2);
                if (by_position) {
                    const uint64_t p = (uint64_t)(rmax0 + re);
                    sp->idr = (int32_t)(uint32_t)p; sp->id = (int32_t)(uint32_t)(p >> 32);
                } else {
                    sp->idr = (int32_t)AR.n_ref;
                    memcpy(AR.take_ref(sp->len1), rseq + re, (size_t)sp->len1);
                }
                a->qe = (int)qe; a->re = rmax0 + re;
            } else {
                a->qe = l_query; a->re = s->rbeg + s->len;
                if (a->rb != H0_ && a->qb != H0_) {     // (L2920-2930)
                    int cov = 0;
                    for (int i = 0; i < c->n; ++i) {
                        const mem_seed_t *t = &c->seeds[i];
                        if (t->qbeg >= a->qb && t->qbeg + t->len <= a->qe && t->rbeg >= a->rb && t->rbeg + t->len <= a->re) cov += t->len;
                    }
                    a->seedcov = cov;
                }
            }
        }
    }
}

// step 3a for one read: seeds (almost) contained in an alignment already made do not keep theirs (L3391-3485)
static void purge_contained(const mem_opt_t *opt, const bseq1_t *seq, mem_chain_v *chn, mem_alnreg_v *av, uint32_t *srtg)
{
    const int l_query = seq->l_seq;
    int lim = 0, s_start = 0;
    for (size_t j = 0; j < chn->n; j++) {
        mem_chain_t *c = &chn->a[j];
        uint32_t *srt2 = srtg + s_start;
        s_start += c->n;
        for (int k = c->n - 1; k >= 0; k--) {
            mem_seed_t *s = &c->seeds[srt2[k]];
            size_t i = 0;
            int v = 0;
            for (i = 0; i < av->n && v < lim; ++i) {    // has an extension covering this seed been made before?
                mem_alnreg_t *p = &av->a[i];
                if (p->qb == -1 && p->qe == -1) continue;
                int64_t rd;
                int qd, w, max_gap;
                if (s->rbeg < p->rb || s->rbeg + s->len > p->re || s->qbeg < p->qb || s->qbeg + s->len > p->qe) { v++; continue; }
                if (s->len - p->seedlen0 > .1 * l_query) { v++; continue; }
                qd = s->qbeg - p->qb; rd = s->rbeg - p->rb;            // the region ahead of the seed
                max_gap = cal_max_gap(opt, qd < rd ? qd : rd);
                w = max_gap < p->w ? max_gap : p->w;
                if (qd - rd < w && rd - qd < w) break;
                qd = p->qe - (s->qbeg + s->len); rd = p->re - (s->rbeg + s->len);   // and the region behind it
                max_gap = cal_max_gap(opt, qd < rd ? qd : rd);
                w = max_gap < p->w ? max_gap : p->w;
                if (qd - rd < w && rd - qd < w) break;
                v++;
            }
            if (v < lim) {                              // contained: keep it only if an overlapping seed says otherwise
                for (v = k + 1; v < c->n; ++v) {
                    if (srt2[v] == UINT_MAX) continue;
                    const mem_seed_t *t = &c->seeds[srt2[v]];
                    if (t->len < s->len * .95) continue;
                    if (s->qbeg <= t->qbeg && s->qbeg + s->len - t->qbeg >= s->len >> 2 && t->qbeg - s->qbeg != t->rbeg - s->rbeg) break;
                    if (t->qbeg <= s->qbeg && t->qbeg + t->len - s->qbeg >= s->len >> 2 && s->qbeg - t->qbeg != s->rbeg - t->rbeg) break;
                }
                if (v == c->n) {
                    mem_alnreg_t *ar = &av->a[s->aln];
                    ar->qb = ar->qe = -1;
                    srt2[k] = UINT_MAX;
                    continue;
                }
            }
            lim++;
        }
    }
}

// step 3b for one read: the tail of mem_kernel2_core (L1666-1719)
static void finish_read(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, bseq1_t *seq, mem_chain_v *chn, mem_alnreg_v *av)
{
    for (size_t i = 0; i < chn->n; ++i)
        if (chn->a[i].m > SEEDS_PER_CHAIN) free(chn->a[i].seeds);
    free(chn->a);
    mem_alnreg_t *a = av->a;
    size_t m = 0;
#if MATE_SORT
    av->useMateSort = true;
#endif
    for (size_t i = 0; i < av->n; ++i)
        if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
    av->n = m;
#if MATE_SORT
    av->n = mem_sort_dedup_patch_mate_sort(opt, bns, pac, (uint8_t *)seq->seq, av->n, av->a, &av->useMateSort);
#else
    av->n = mem_sort_dedup_patch(opt, bns, pac, (uint8_t *)seq->seq, av->n, av->a);
#endif
    for (size_t i = 0; i < av->n; ++i) {
        mem_alnreg_t *p = &av->a[i];
        if (p->rid >= 0 && bns->anns[p->rid].is_alt) p->is_alt = 1;
    }
}

}  // namespace meme_ext

// Replaces kt_for(worker_aln, &w, n) in mem_process_seqs for the whole chunk.  Returns false when the GPU extension is
// off (the caller then runs the reference's own loop).
bool meme_extend_chunk(worker_t &w, int n)
{
    using namespace meme_ext;
    if (!meme_glue_bsw_enabled() || n <= 0) return false;
    const mem_opt_t *opt = w.opt;
    const bntseq_t *bns = w.fmi->idx->bns;
    const uint8_t *pac = w.fmi->idx->pac;
    const int T = std::max(1, (int)w.nthreads);
    const bool by_position = meme_glue_bsw_text_mode();     // the reference side of the pairs is read from the resident text
    if ((int)g_ts.size() < T) g_ts.resize((size_t)T);
    g_srt_off.resize((size_t)n + 1);
    double t0 = omp_get_wtime();
    // Blocks of reads are handed out dynamically (a host thread that shares its core with the aligner's reader / writer threads
    // takes fewer), one arena set per host thread.  The chunk is cut into two halves of blocks: while one half's pairs run on
    // the GPU (one call in flight at a time, driven by the glue's per-device threads), the host builds, folds or finishes the
    // other half (MEME_B200_EXT_HALVES=0: one piece, as before).
    const int BLK = 1024, n_blk = (n + BLK - 1) / BLK;
    const int n_half = (n_blk >= 8 && !(getenv("MEME_B200_EXT_HALVES") && atoi(getenv("MEME_B200_EXT_HALVES")) == 0) &&
                        !getenv("MEME_B200_DUMP_PAIRS")) ? 2 : 1;              // (the pair dump wants one left and one right batch)
    const int blk_cut[3] = {0, n_half == 2 ? n_blk / 2 : n_blk, n_blk};
    g_blk_owner.resize((size_t)n_blk);
    double t_pairs = 0, t_left = 0, t_fold = 0, t_right = 0, t_finish = 0;
    int64_t tot_left = 0, tot_right = 0;
#pragma omp parallel num_threads(T)
    {
        ThreadState &S = g_ts[(size_t)omp_get_thread_num()];
        S.srt.clear();
        for (int h = 0; h < 2; h++) { S.left[h].reset(); S.right[h].reset(); }
        // first chunk: the arenas take their typical share (up to 6 pairs per read, 200 + 100 sequence bytes per pair), from
        // the blocks the glue page-locked at index load when it could
        const int64_t est = (int64_t)(n / T + 1) * 6 / n_half;
        for (int h = 0; h < n_half; h++)
            for (Arena *A : {&S.left[h], &S.right[h]}) {
                Arena::reserve(A->pairs, A->cap, est); if (!by_position) Arena::reserve(A->ref, A->cap_ref, est * 200); Arena::reserve(A->qer, A->cap_qer, est * 100);
                Arena::reserve(A->w_used, A->cap_w, est);
            }
    }
    // ---- 1. regions + pairs of one half ----
    auto build_pairs = [&](int h) {
        const double ta = omp_get_wtime();
#pragma omp parallel num_threads(T)
        {
            const int t = omp_get_thread_num();
            ThreadState &S = g_ts[(size_t)t];
#pragma omp for schedule(dynamic, 1)
            for (int b = blk_cut[h]; b < blk_cut[h + 1]; b++) {
                g_blk_owner[(size_t)b] = t;
                const int r0 = b * BLK, r1 = std::min(n, r0 + BLK);
                for (int l = r0; l < r1; l++) {
                    g_srt_off[(size_t)l] = (int64_t)S.srt.size();
                    collect_read(opt, bns, pac, w.ref_string, &w.seqs[l], l, &w.chain_ar[l], &w.regs[l], S, h, by_position);
                }
            }
        }
        t_pairs += omp_get_wtime() - ta;
    };
    // ---- 2. one side of one half on the device: start (returns at once), wait ----
    std::vector<SeqPair *> pp((size_t)T);
    std::vector<int64_t> np((size_t)T), nr((size_t)T), nq((size_t)T);
    std::vector<uint8_t *> pr((size_t)T), pq((size_t)T);
    std::vector<int32_t *> pw((size_t)T);
    void *job = nullptr;
    double t_job = 0;
    auto start_side = [&](int h, bool right) {
        int64_t tot = 0;
        for (int t = 0; t < T; t++) {
            Arena &A = right ? g_ts[(size_t)t].right[h] : g_ts[(size_t)t].left[h];
            Arena::grow(A.w_used, A.cap_w, 0, A.n + 1);
            pp[(size_t)t] = A.pairs; np[(size_t)t] = A.n; pr[(size_t)t] = by_position ? nullptr : A.ref; nr[(size_t)t] = A.n_ref; pq[(size_t)t] = A.qer;
            nq[(size_t)t] = A.n_qer; pw[(size_t)t] = A.w_used; tot += A.n;
        }
        (right ? tot_right : tot_left) += tot;
        t_job = omp_get_wtime();
        job = tot ? meme_glue_bsw_segments_start(T, pp.data(), np.data(), pr.data(), nr.data(), pq.data(), nq.data(), opt->w, MAX_BAND_TRY, right ? 1 : 0, opt->o_del,
                                                 opt->e_del, opt->o_ins, opt->e_ins, opt->zdrop, right ? opt->pen_clip3 : opt->pen_clip5, opt->a, opt->b, pw.data())
                  : nullptr;
    };
    auto wait_side = [&](bool right) {
        const double ta = omp_get_wtime();
        if (job) meme_glue_bsw_segments_wait(job);
        job = nullptr;
        (right ? t_right : t_left) += omp_get_wtime() - ta;          // what the host waited, not what the device took
    };
    // ---- 2a'. left side folded into the regions (L2968-3161); the right side starts from the region's score so far (L3168-3173) ----
    auto fold_left = [&](int h) {
        const double ta = omp_get_wtime();
#pragma omp parallel num_threads(T)
        {
            ThreadState &S = g_ts[(size_t)omp_get_thread_num()];
            Arena &L = S.left[h], &R = S.right[h];
            for (int64_t k = 0; k < L.n; k++) {
                const SeqPair *sp = &L.pairs[k];
                mem_alnreg_t *a = &w.regs[sp->seqid].a[sp->regid];
                a->score = sp->score;
                if (sp->gscore <= 0 || sp->gscore <= a->score - opt->pen_clip5) { a->qb -= sp->qle; a->rb -= sp->tle; a->truesc = a->score; }
                else { a->qb = 0; a->rb -= sp->gtle; a->truesc = sp->gscore; }
                a->w = max_(a->w, L.w_used[k]);
                seed_coverage(a);
            }
            for (int64_t k = 0; k < R.n; k++) {
                SeqPair *sp = &R.pairs[k];
                sp->h0 = w.regs[sp->seqid].a[sp->regid].score;
            }
        }
        t_fold += omp_get_wtime() - ta;
    };
    // ---- 3. fold the right side, purge contained seeds, finish every read of the half ----
    auto finish_half = [&](int h) {
        const double ta = omp_get_wtime();
#pragma omp parallel num_threads(T)
        {
            ThreadState &S = g_ts[(size_t)omp_get_thread_num()];
            Arena &R = S.right[h];
            for (int64_t k = 0; k < R.n; k++) {
                const SeqPair *sp = &R.pairs[k];
                mem_alnreg_t *a = &w.regs[sp->seqid].a[sp->regid];
                a->score = sp->score;
                if (sp->gscore <= 0 || sp->gscore <= a->score - opt->pen_clip3) { a->qe += sp->qle; a->re += sp->tle; a->truesc += a->score - sp->h0; }
                else { a->qe = w.seqs[sp->seqid].l_seq; a->re += sp->gtle; a->truesc += sp->gscore - sp->h0; }
                a->w = max_(a->w, R.w_used[k]);
                seed_coverage(a);
            }
            // every region of the half has both sides folded before any of its reads is finished (a block may be finished by
            // another thread than the one that built its pairs)
#pragma omp barrier
#pragma omp for schedule(dynamic, 1)
            for (int b = blk_cut[h]; b < blk_cut[h + 1]; b++) {
                uint32_t *srt = g_ts[(size_t)g_blk_owner[(size_t)b]].srt.data();
                const int r0 = b * BLK, r1 = std::min(n, r0 + BLK);
                for (int l = r0; l < r1; l++) {
                    purge_contained(opt, &w.seqs[l], &w.chain_ar[l], &w.regs[l], srt + g_srt_off[(size_t)l]);
                    finish_read(opt, bns, pac, &w.seqs[l], &w.chain_ar[l], &w.regs[l]);
                }
            }
        }
        t_finish += omp_get_wtime() - ta;
    };
    if (n_half == 1) {
        build_pairs(0);
        start_side(0, false); wait_side(false);
        fold_left(0);
        start_side(0, true); wait_side(true);
        finish_half(0);
    } else {
        build_pairs(0);
        start_side(0, false); build_pairs(1); wait_side(false);        // left(A)  || pairs(B)
        start_side(1, false); fold_left(0);   wait_side(false);        // left(B)  || fold(A)
        start_side(0, true);  fold_left(1);   wait_side(true);         // right(A) || fold(B)
        start_side(1, true);  finish_half(0); wait_side(true);         // right(B) || finish(A)
        finish_half(1);
    }
    const double t1 = t0 + t_pairs, t2 = t1 + t_left, t3 = t2 + t_fold, t4 = t3 + t_right, t5 = t4 + t_finish;
    (void)t_job;
    fprintf(stderr, "[meme_b200] extended %d reads on the GPU (%d piece%s): %ld left + %ld right pairs; pairs %.1f ms, left %.1f ms, fold %.1f ms, "
            "right %.1f ms, finish %.1f ms (left / right = what the host waited for the device)\n", n, n_half, n_half > 1 ? "s, overlapped" : "", (long)tot_left, (long)tot_right, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3,
            (t4 - t3) * 1e3, (t5 - t4) * 1e3);
    return true;
}


// Called in front of kt_for(worker_sam) in mem_process_seqs: the band-doubling loop of mem_reg2aln around bwa_gen_cigar2
// (L2339-2348) for EVERY region of the chunk in one device call per GPU.  worker_sam then finds the results through
// meme_glue_take_cigar; which regions it actually prints (primary, supplementary, XA) is its own business, so all are done.
void meme_cigar_chunk(worker_t &w, int n)
{
    if (!meme_glue_cigar_enabled() || n <= 0) return;
    const mem_opt_t *opt = w.opt;
    const int64_t l_pac = w.fmi->idx->bns->l_pac;
    static std::vector<int64_t> task_off;          // one chunk at a time (mem_process_seqs); buffers reused across chunks
    static std::vector<int32_t> reg_w;
    task_off.resize((size_t)n + 1);
    task_off[0] = 0;
    for (int i = 0; i < n; i++) {
        int c = 0;
        const mem_alnreg_v &v = w.regs[i];
        for (size_t k = 0; k < v.n; k++) if (v.a[k].rb >= 0 && v.a[k].re >= 0) c++;
        task_off[(size_t)i + 1] = task_off[(size_t)i] + c;
    }
    meme_cigar_task_t *tasks = meme_glue_cigar_tasks(task_off[(size_t)n]);
    reg_w.resize((size_t)task_off[(size_t)n]);
#pragma omp parallel for schedule(static) num_threads(std::max(1, (int)w.nthreads))
    for (int i = 0; i < n; i++) {
        const mem_alnreg_v &v = w.regs[i];
        int64_t t = task_off[(size_t)i];
        for (size_t k = 0; k < v.n; k++) {
            const mem_alnreg_t *ar = &v.a[k];
            if (ar->rb < 0 || ar->re < 0) continue;
            const int qb = ar->qb, qe = ar->qe;
            const int64_t rb = ar->rb, re = ar->re;
            int tmp = infer_bw(qe - qb, re - rb, ar->truesc, opt->a, opt->o_del, opt->e_del);        // L2334-2338
            int w2 = infer_bw(qe - qb, re - rb, ar->truesc, opt->a, opt->o_ins, opt->e_ins);
            w2 = w2 > tmp ? w2 : tmp;
            if (w2 > opt->w) w2 = w2 < ar->w ? w2 : ar->w;
            meme_cigar_task_t &T = tasks[(size_t)t];
            T.rb = rb; T.re = re; T.q_off = meme_glue_code_off(i) + qb; T.qlen = qe - qb; T.w = w2; T.truesc = ar->truesc; T.pad = 0;
            reg_w[(size_t)t] = ar->w;
            t++;
        }
    }
    meme_glue_cigar_run(reg_w, task_off, w.seqs, n, opt, l_pac);
}


// Called in front of kt_for(worker_sam) for paired-end chunks: the mate-rescue Smith-Waterman of the WHOLE chunk in one device
// call (MEME_B200_MATE=1).  The reference collects the rescue pairs of a 512-read block with mem_sam_pe_batch_pre, aligns them
// with kswv (mem_sam_pe_batch) and consumes the results in mem_sam_pe_batch_post, block by block inside worker_sam
// (src/bwamem.cpp:L1851-1901).  mem_sam_pe_batch_pre only reads the regions, so it is run here first for every read pair -- the
// reference's own function, nothing restated -- and the pairs it produces are copied into per-thread pinned arenas; worker_sam
// then runs unchanged up to the point where it would call mem_sam_pe_batch and takes the block's results from here instead
// (meme_glue_take_matesw: aln[regid] of a block = the results of its read pairs in the order they were produced).
static std::vector<meme_ext::Arena> g_msw_ar;
static std::vector<int64_t> g_msw_pair_off;
// what mem_sam_pe_batch_pre leaves in `gar` (pair index or -1 per region and orientation) for every read pair, so that worker_sam
// does not have to run it a second time: per host thread the values, per read pair where they are
static std::vector<std::vector<int32_t>> g_msw_gar;
static std::vector<int64_t> g_msw_gar_off;      // per read pair: offset in its thread's values; [np] unused
static std::vector<int32_t> g_msw_gar_cnt;
static double g_msw_collect_ms = 0;

static bool meme_matesw_wanted(worker_t &w, int n)
{
    return meme_glue_matesw_enabled() && (w.opt->flag & MEM_F_PE) && !(w.opt->flag & MEM_F_NO_RESCUE) && n >= 2;
}

// host part: the chunk's rescue pairs into the per-thread arenas.  false = nothing to run on the device for this chunk.
static bool meme_matesw_collect(worker_t &w, int n)
{
    meme_glue_matesw_skip();       // nothing from an earlier chunk
    if (!meme_matesw_wanted(w, n)) return false;
    std::vector<meme_ext::Arena> &ar = g_msw_ar;
    std::vector<int64_t> &pair_off = g_msw_pair_off;
    const int T = std::max(1, (int)w.nthreads);
    const int np = n >> 1;
    if ((int)ar.size() < T) ar.resize((size_t)T);
    if ((int)g_msw_gar.size() < T) g_msw_gar.resize((size_t)T);
    g_msw_gar_off.resize((size_t)np + 1);
    g_msw_gar_cnt.resize((size_t)np + 1);
    {   // first chunk: the blocks the glue pinned when the index was loaded (best fit from its cache)
        size_t sz[3];
        meme_glue_matesw_arena_sizes(T, n, sz);
        for (int t = 0; t < T; t++) {
            meme_ext::Arena &A = ar[(size_t)t];
            meme_ext::Arena::reserve(A.pairs, A.cap, (int64_t)(sz[0] / sizeof(SeqPair)));
            meme_ext::Arena::reserve(A.ref, A.cap_ref, (int64_t)sz[1]);
            meme_ext::Arena::reserve(A.qer, A.cap_qer, (int64_t)sz[2]);
        }
    }
    pair_off.assign((size_t)np + 1, 0);
    auto t0 = omp_get_wtime();
    bool too_long = false;
#pragma omp parallel num_threads(T)
    {
        const int tid = omp_get_thread_num();
        const int p0 = (int)((int64_t)np * tid / T), p1 = (int)((int64_t)np * (tid + 1) / T);
        meme_ext::Arena &A = ar[(size_t)tid];
        A.reset();
        std::vector<int32_t> &G = g_msw_gar[(size_t)tid];
        G.clear();
        for (int p = p0; p < p1; p++) {
            int64_t pcnt = 0;
            int32_t gcnt = 0, maxRefLen = 0, maxQerLen = 0;
            mem_sam_pe_batch_pre(w.opt, w.fmi->idx->bns, w.fmi->idx->pac, w.pes, (uint64_t)((w.n_processed >> 1) + p), &w.seqs[2 * p], &w.regs[2 * p],
                                 &w.mmc, pcnt, gcnt, maxRefLen, maxQerLen, tid);
            pair_off[(size_t)p + 1] = pcnt;
            const int32_t *gar = (const int32_t *)w.mmc.seqPairArrayAux[tid];
            g_msw_gar_off[(size_t)p] = (int64_t)G.size();
            g_msw_gar_cnt[(size_t)p] = gcnt;
            G.insert(G.end(), gar, gar + gcnt);
            const SeqPair *sp = w.mmc.seqPairArrayLeft128[tid];
            const uint8_t *rbuf = w.mmc.seqBufLeftRef[tid * CACHE_LINE], *qbuf = w.mmc.seqBufLeftQer[tid * CACHE_LINE];
            for (int64_t k = 0; k < pcnt; k++) {
                SeqPair q = sp[k];
                if (q.len1 > 32000 || q.len2 > 496 || ((q.h0 & KSW_XBYTE) && q.len2 > 255)) too_long = true;
                uint8_t *r = A.take_ref(q.len1), *s = A.take_qer(q.len2);
                memcpy(r, rbuf + q.idr, (size_t)q.len1);
                memcpy(s, qbuf + q.idq, (size_t)q.len2);
                q.idr = (int32_t)(r - A.ref); q.idq = (int32_t)(s - A.qer);
                *A.new_pair() = q;
            }
        }
    }
    if (too_long) { meme_glue_matesw_skip(); return false; }        // beyond the kernel's limits: this chunk takes the reference's own code
    for (int p = 0; p < np; p++) pair_off[(size_t)p + 1] += pair_off[(size_t)p];
    g_msw_collect_ms = (omp_get_wtime() - t0) * 1e3;
    return true;
}

// device part: one call per GPU on the collected segments
static void meme_matesw_run(worker_t &w)
{
    std::vector<meme_ext::Arena> &ar = g_msw_ar;
    const int T = std::max(1, (int)w.nthreads);
    std::vector<SeqPair *> pp((size_t)T); std::vector<uint8_t *> rp((size_t)T), qp((size_t)T);
    std::vector<int64_t> pn((size_t)T), rn((size_t)T), qn((size_t)T);
    for (int t = 0; t < T; t++) { meme_ext::Arena &A = ar[(size_t)t]; pp[t] = A.pairs; pn[t] = A.n; rp[t] = A.ref; rn[t] = A.n_ref; qp[t] = A.qer; qn[t] = A.n_qer; }
    meme_glue_matesw_run(T, pp.data(), pn.data(), rp.data(), rn.data(), qp.data(), qn.data(), w.opt, g_msw_pair_off, g_msw_collect_ms);
}

// worker_sam, in front of its own mem_sam_pe_batch_pre loop over the read pairs [pb, pe) of a block: what that loop would leave
// behind -- the block's `gar` array (pair indexes relative to the block) and pair count -- from the chunk-level pre-pass.
// false: nothing collected for this chunk, the caller runs its loop.
bool meme_matesw_restore(worker_t *w, int pb, int pe, int tid, int64_t *pcnt)
{
    if (!meme_glue_matesw_have() || !getenv("MEME_B200_MATE_RESTORE")) return false;
    const std::vector<int64_t> &off = meme_glue_matesw_off();          // per read pair: index of its first result (the glue owns it now)
    const int np = (int)off.size() - 1;
    const int T = std::max(1, (int)w->nthreads);
    if (np < 1 || pe > np || (int)g_msw_gar.size() < T) return false;
    int32_t *gar = (int32_t *)w->mmc.seqPairArrayAux[tid];
    const int64_t base = off[(size_t)pb];
    int64_t g = 0;
    for (int p = pb; p < pe; p++) {
        int t = (int)(((int64_t)p * T + T - 1) / np);                      // owner of read pair p in the static partition of the pre-pass
        if (t >= T) t = T - 1;
        while (t > 0 && (int)((int64_t)np * t / T) > p) t--;
        while (t + 1 < T && (int)((int64_t)np * (t + 1) / T) <= p) t++;
        const int32_t *src = g_msw_gar[(size_t)t].data() + g_msw_gar_off[(size_t)p];
        const int64_t shift = off[(size_t)p] - base;
        for (int k = 0; k < g_msw_gar_cnt[(size_t)p]; k++) gar[g++] = src[k] < 0 ? -1 : (int32_t)(src[k] + shift);
    }
    *pcnt = off[(size_t)pe] - base;
    return true;
}

void meme_matesw_chunk(worker_t &w, int n)
{
    if (meme_matesw_collect(w, n)) meme_matesw_run(w);
}

// Both pre-passes of the SAM stage.  The global alignments keep the GPU busy for most of their time while the host waits, and
// collecting the rescue pairs is host work only, so with both stages on the alignments are driven from a second thread while the
// workers collect (MEME_B200_NO_OVERLAP=1: one after the other).  Measured on 1 M pairs, alternating runs on one box: 221-224
// against 203-209 Mbases/s (profiles/r2/README.md; a first attempt lost to a cold-start stall that was the warm-up's fault).
void meme_sam_prepass(worker_t &w, int n)
{
    if (!meme_glue_cigar_enabled() || !meme_matesw_wanted(w, n) || getenv("MEME_B200_NO_OVERLAP")) {
        meme_cigar_chunk(w, n);
        meme_matesw_chunk(w, n);
        return;
    }
    std::thread cig([&] { meme_cigar_chunk(w, n); });
    const bool have = meme_matesw_collect(w, n);
    cig.join();
    if (have) meme_matesw_run(w);
}
