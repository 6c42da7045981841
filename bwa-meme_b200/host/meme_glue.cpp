// meme_glue.cpp -- see meme_glue.h.  Compiled together with the (patched) reference sources; binds
// libmeme_b200.so with dlopen so that the aligner has no link-time CUDA dependency.
#include <dlfcn.h>
#include <malloc.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <algorithm>
#include <vector>
#include <mutex>
#include <atomic>
#include <thread>
#include <chrono>
#include "meme_glue.h"
#include "../../include/meme_b200.h"

namespace {

struct Api {
    void *so = nullptr;
    decltype(&meme_index_load) index_load = nullptr;
    decltype(&meme_seed_batch) seed_batch = nullptr;
    decltype(&meme_seed_warmup) seed_warmup = nullptr;
    decltype(&meme_index_load_devices) index_load_devices = nullptr;
    decltype(&meme_seed_batch_multi) seed_batch_multi = nullptr;
    decltype(&meme_bsw_create) bsw_create = nullptr;
    decltype(&meme_bsw_batch) bsw_batch = nullptr;
    decltype(&meme_last_error) last_error = nullptr;
    decltype(&meme_bsw_upload_segments) bsw_upload_segments = nullptr;
    decltype(&meme_bsw_download_segments) bsw_download_segments = nullptr;
    decltype(&meme_bsw_run_ext) bsw_run_ext = nullptr;
    decltype(&meme_bsw_set_match) bsw_set_match = nullptr;
    decltype(&meme_bsw_attach_text) bsw_attach_text = nullptr;
    decltype(&meme_pinned_alloc) pinned_alloc = nullptr;
    decltype(&meme_pinned_free) pinned_free = nullptr;
    decltype(&meme_chain_set_contigs) chain_set_contigs = nullptr;
    decltype(&meme_chain_batch_multi) chain_batch_multi = nullptr;
    decltype(&meme_chain_warmup) chain_warmup = nullptr;
    decltype(&meme_cigar_batch) cigar_batch = nullptr;
    decltype(&meme_cigar_warmup) cigar_warmup = nullptr;
    decltype(&meme_matesw_batch) matesw_batch = nullptr;
    decltype(&meme_matesw_warmup) matesw_warmup = nullptr;
    bool ok = false;
};

Api &api()
{
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *path = getenv("MEME_B200_LIB");
        a.so = dlopen(path ? path : "libmeme_b200.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.so) { fprintf(stderr, "[meme_b200] cannot load the CUDA library: %s\n", dlerror()); exit(EXIT_FAILURE); }
        a.index_load = (decltype(a.index_load))dlsym(a.so, "meme_index_load");
        a.seed_batch = (decltype(a.seed_batch))dlsym(a.so, "meme_seed_batch");
        a.seed_warmup = (decltype(a.seed_warmup))dlsym(a.so, "meme_seed_warmup");
        a.index_load_devices = (decltype(a.index_load_devices))dlsym(a.so, "meme_index_load_devices");
        a.seed_batch_multi = (decltype(a.seed_batch_multi))dlsym(a.so, "meme_seed_batch_multi");
        a.bsw_create = (decltype(a.bsw_create))dlsym(a.so, "meme_bsw_create");
        a.bsw_batch = (decltype(a.bsw_batch))dlsym(a.so, "meme_bsw_batch");
        a.last_error = (decltype(a.last_error))dlsym(a.so, "meme_last_error");
        a.bsw_upload_segments = (decltype(a.bsw_upload_segments))dlsym(a.so, "meme_bsw_upload_segments");
        a.bsw_download_segments = (decltype(a.bsw_download_segments))dlsym(a.so, "meme_bsw_download_segments");
        a.bsw_run_ext = (decltype(a.bsw_run_ext))dlsym(a.so, "meme_bsw_run_ext");
        a.bsw_set_match = (decltype(a.bsw_set_match))dlsym(a.so, "meme_bsw_set_match");
        a.bsw_attach_text = (decltype(a.bsw_attach_text))dlsym(a.so, "meme_bsw_attach_text");
        a.pinned_alloc = (decltype(a.pinned_alloc))dlsym(a.so, "meme_pinned_alloc");
        a.pinned_free = (decltype(a.pinned_free))dlsym(a.so, "meme_pinned_free");
        a.chain_set_contigs = (decltype(a.chain_set_contigs))dlsym(a.so, "meme_chain_set_contigs");
        a.chain_batch_multi = (decltype(a.chain_batch_multi))dlsym(a.so, "meme_chain_batch_multi");
        a.chain_warmup = (decltype(a.chain_warmup))dlsym(a.so, "meme_chain_warmup");
        a.cigar_batch = (decltype(a.cigar_batch))dlsym(a.so, "meme_cigar_batch");
        a.cigar_warmup = (decltype(a.cigar_warmup))dlsym(a.so, "meme_cigar_warmup");
        a.matesw_batch = (decltype(a.matesw_batch))dlsym(a.so, "meme_matesw_batch");
        a.matesw_warmup = (decltype(a.matesw_warmup))dlsym(a.so, "meme_matesw_warmup");
        a.ok = a.matesw_batch && a.matesw_warmup && a.cigar_warmup && a.cigar_batch && a.chain_warmup && a.chain_set_contigs && a.chain_batch_multi && a.index_load && a.seed_batch && a.bsw_create && a.bsw_batch && a.last_error && a.bsw_upload_segments &&
               a.bsw_download_segments && a.bsw_run_ext && a.bsw_set_match && a.bsw_attach_text && a.pinned_alloc && a.pinned_free && a.index_load_devices &&
               a.seed_batch_multi;
        if (!a.ok) { fprintf(stderr, "[meme_b200] libmeme_b200.so lacks required symbols\n"); exit(EXIT_FAILURE); }
    });
    return a;
}

bool env_on(const char *name)
{
    const char *e = getenv(name);
    return e && e[0] == '1';
}

std::string g_prefix;
// one index handle per GPU (MEME_B200_DEVICES=0,1,..; default: MEME_B200_DEVICE or 0): reads shard, the index is replicated
std::vector<int> g_devices;
std::vector<meme_index_t *> g_index;
std::vector<meme_seed_result_t> g_shard;       // results of the chunk being processed, one per device
std::vector<int64_t> g_shard_begin;
// results of the chunk being processed (valid until the next meme_glue_seed_chunk)
const bseq1_t *g_base = nullptr;
int g_n = 0;
bool g_have = false;
int g_max_occ = 0;
// MEME_B200_CHAIN=1: chains instead of SMEM/hit lists (meme_chain_batch_multi); the seeds of the chunk's chains live in one
// arena of mem_seed_t that the chain records handed to the aligner point into (valid until the next chunk)
std::vector<meme_chain_result_t> g_cshard;
std::vector<int64_t> g_cseed_base;
bool g_have_chains = false;
bool g_contigs_set = false;
mem_seed_t *g_seed_arena = nullptr;
size_t g_seed_arena_cap = 0;
double g_seed_s = 0;
long g_seed_reads = 0;
// MEME_B200_SAM=1: the global alignments of the SAM stage, computed for every region of the chunk before worker_sam runs
// base codes of the chunk's reads (filled by meme_glue_seed_chunk), in page-locked memory: the seeding and the SAM-stage calls
// both upload them, and a pageable source halves the copy rate
struct PinnedBytes {
    uint8_t *p = nullptr; size_t n = 0, cap = 0;
    uint8_t *data() { return p; }
    size_t size() const { return n; }
    void resize(size_t need);
};
PinnedBytes g_codes;
std::vector<int64_t> g_code_off;
bool g_have_cigar = false;
PinnedBytes g_ctask_buf;                        // the chunk's meme_cigar_task_t array (page-locked: it is uploaded as it is)
meme_cigar_task_t *g_ctasks = nullptr;
int64_t g_n_ctasks = 0;
std::vector<int32_t> g_creg_w;
std::vector<int64_t> g_ctask_off;               // per read: its tasks
std::vector<meme_cigar_result_t> g_cres;        // per device
std::vector<int64_t> g_cres_begin;              // first task of each device's share
// seqs[i].seq -> i: open-addressing table of read indexes (-1 = empty), filled by all threads
std::vector<std::atomic<int32_t>> g_cslot;
std::vector<const char *> g_cseq;
size_t g_cmask = 0;
inline size_t ptr_hash(const char *p) { uint64_t x = (uint64_t)(uintptr_t)p; x ^= x >> 29; x *= 0x9E3779B97F4A7C15ull; return (size_t)(x >> 20); }
std::atomic<long> g_cig_hit{0}, g_cig_miss{0};   // lookups of mem_reg2aln answered from the GPU results / left to its own code

[[noreturn]] void die(const char *what, int rc)
{
    fprintf(stderr, "[meme_b200] %s failed (%d): %s\n", what, rc, api().last_error());
    exit(EXIT_FAILURE);      // the reference's own error protocol
}

void PinnedBytes::resize(size_t need)
{
    if (need > cap) {
        void *q = nullptr;
        const size_t ncap = need + need / 8 + 4096;
        int rc = api().pinned_alloc(ncap, &q);
        if (rc) die("meme_pinned_alloc", rc);
        if (p) api().pinned_free(p);              // contents are rewritten by every chunk
        p = (uint8_t *)q; cap = ncap;
    }
    n = need;
}

}  // namespace

static void ensure_index()
{
    if (!g_index.empty()) return;
    Api &a = api();
    if (!getenv("MEME_B200_NO_MALLOPT")) {
        // Every chunk allocates and frees several hundred MB of small blocks (regions, chains, SAM strings) from 16 threads while
        // the aligner's reader and writer threads do the same; with glibc's defaults the arenas are trimmed and re-grown every
        // chunk, and the page faults and mmap calls of all threads serialise on the address-space lock (the extension driver's
        // pair construction ran 4x slower whenever the reader thread was parsing).  The reference's own build links mimalloc for
        // this; the recipe here cannot (oracle/Makefile.ref), so the glue at least tells glibc to keep what it has.
        mallopt(M_TRIM_THRESHOLD, 0x7fffffff);
        mallopt(M_MMAP_THRESHOLD, 32 << 20);
        mallopt(M_TOP_PAD, 64 << 20);
    }
    int mode = getenv("MEME_B200_MODE") ? atoi(getenv("MEME_B200_MODE")) : MODE;
    if (const char *ds = getenv("MEME_B200_DEVICES")) {
        for (const char *p = ds; *p;) { g_devices.push_back(atoi(p)); while (*p && *p != ',') p++; if (*p) p++; }
    }
    if (g_devices.empty()) g_devices.push_back(getenv("MEME_B200_DEVICE") ? atoi(getenv("MEME_B200_DEVICE")) : 0);
    const int nd = (int)g_devices.size();
    g_index.assign((size_t)nd, nullptr);
    g_shard.resize((size_t)nd);
    g_shard_begin.assign((size_t)nd + 1, 0);
    auto t0 = std::chrono::steady_clock::now();
    int rc = a.index_load_devices(g_prefix.c_str(), mode, g_devices.data(), nd, g_index.data());
    if (rc) die("meme_index_load_devices", rc);
    fprintf(stderr, "[meme_b200] index %s resident on %d GPU(s) (mode %d) in %.2f s\n", g_prefix.c_str(), nd, mode,
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    // kernels loaded and pools sized for a typical chunk before the first timed chunk (-K 100000000 at 150 bp ~ 667 k reads)
    long warm = getenv("MEME_B200_WARM_READS") ? atol(getenv("MEME_B200_WARM_READS")) : 700000;
    if (warm > 0) {
        auto w0 = std::chrono::steady_clock::now();
        for (int d = 0; d < nd; d++) {
            rc = a.seed_warmup(g_index[(size_t)d], warm / nd + 1, 150);
            if (rc) die("meme_seed_warmup", rc);
            if (env_on("MEME_B200_CHAIN") && (rc = a.chain_warmup(g_index[(size_t)d]))) die("meme_chain_warmup", rc);
            if (env_on("MEME_B200_MATE") && (rc = a.matesw_warmup(g_index[(size_t)d], warm / nd + 1, 640, 150))) die("meme_matesw_warmup", rc);
            if (env_on("MEME_B200_SAM") && (rc = a.cigar_warmup(g_index[(size_t)d], warm * 3 / nd + 1, 150))) die("meme_cigar_warmup", rc);
        }
        // the chunk-sized host buffers of the glue, touched once so that the first chunk does not pay their page faults
        g_codes.resize((size_t)warm * 152 + 1);
        memset(g_codes.data(), 0, g_codes.size());
        if (env_on("MEME_B200_SAM")) g_ctask_buf.resize((size_t)warm * 3 * sizeof(meme_cigar_task_t));
        g_code_off.resize((size_t)warm + 1);
        memset(g_code_off.data(), 0, g_code_off.size() * sizeof(g_code_off[0]));
        fprintf(stderr, "[meme_b200] warm-up batch of %ld reads in %.2f s\n", warm,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count());
        if (env_on("MEME_B200_BSW")) {
            int threads = getenv("MEME_B200_THREADS") ? atoi(getenv("MEME_B200_THREADS")) : (int)std::thread::hardware_concurrency();
            cpu_set_t cs;
            if (!getenv("MEME_B200_THREADS") && sched_getaffinity(0, sizeof cs, &cs) == 0) threads = CPU_COUNT(&cs);
            meme_glue_pinned_reserve(threads, warm);
            meme_glue_bsw_warmup(warm);
        }
        if (env_on("MEME_B200_MATE")) {
            int threads = getenv("MEME_B200_THREADS") ? atoi(getenv("MEME_B200_THREADS")) : (int)std::thread::hardware_concurrency();
            cpu_set_t cs;
            if (!getenv("MEME_B200_THREADS") && sched_getaffinity(0, sizeof cs, &cs) == 0) threads = CPU_COUNT(&cs);
            meme_glue_pinned_reserve_matesw(threads, warm);
        }
    }
}

// Called from memoryAllocLearned(): the index goes to HBM where the reference loads its own copy, i.e. outside
// the per-chunk timers of mem_process_seqs ("Processed N reads in ... real sec", src/bwamem.cpp:L1968).
void meme_glue_set_prefix(const char *idx_prefix)
{
    g_prefix = idx_prefix ? idx_prefix : "";
    if (env_on("MEME_B200")) ensure_index();
}

// One GPU seeding call for the whole chunk (the reference seeds 512 reads at a time per thread).
void meme_glue_seed_chunk(const mem_opt_t *opt, const bntseq_t *bns, bseq1_t *seqs, int n, int use_learned)
{
    g_have = false;
    g_have_chains = false;
    if (!use_learned || !env_on("MEME_B200") || n <= 0) return;
    Api &a = api();
    ensure_index();
    auto t0 = std::chrono::steady_clock::now();
    // base codes: nst_nt4_table of the ASCII read; worker_bwt converts seqs[i].seq in place later
    // (src/bwamem.cpp:L1277-1280) and later stages rely on that, so the reads are not touched here
    PinnedBytes &codes = g_codes;
    std::vector<int64_t> &off = g_code_off;
    g_have_cigar = false;
    off.resize((size_t)n + 1);
    size_t tot = 0;
    for (int i = 0; i < n; i++) { off[i] = (int64_t)tot; tot += (size_t)seqs[i].l_seq; }
    off[n] = (int64_t)tot;
    codes.resize(tot + 1);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const unsigned char *s = (const unsigned char *)seqs[i].seq;
        uint8_t *d = codes.data() + off[i];
        for (int k = 0; k < seqs[i].l_seq; k++) d[k] = s[k] < 4 ? s[k] : nst_nt4_table[s[k]];
    }
    auto t1 = std::chrono::steady_clock::now();
    meme_seed_opt_t so;
    so.min_seed_len = opt->min_seed_len;
    so.split_len = (int)(opt->min_seed_len * opt->split_factor + .499);
    so.split_width = opt->split_width;
    so.max_mem_intv = (int32_t)opt->max_mem_intv;
    so.rounds = 3;
    // only the hits mem_chain_Learned reads are materialised (k = 0, step, 2*step, ... at most max_occ of them,
    // src/bwamem.cpp:L1154-1155): a repeat SMEM with 1e5 occurrences costs max_occ positions, not 1e5
    so.max_occ = opt->max_occ > 0 ? opt->max_occ : 0;
    g_max_occ = so.max_occ;
    if (env_on("MEME_B200_CHAIN") && so.max_occ > 0 && bns) {
        // seeding AND chaining on the GPU: mem_chain_Learned + mem_chain_flt run on the device-resident lists, only the
        // filtered chains and their seeds come back (src/bwamem.cpp:L1398-1407 for every read of the chunk)
        if (!g_contigs_set) {
            std::vector<meme_contig_t> ct((size_t)bns->n_seqs);
            for (int i = 0; i < bns->n_seqs; i++) { ct[(size_t)i].offset = bns->anns[i].offset; ct[(size_t)i].len = bns->anns[i].len; ct[(size_t)i].is_alt = bns->anns[i].is_alt; }
            for (auto ix : g_index) {
                int rc = a.chain_set_contigs(ix, ct.data(), bns->n_seqs, bns->l_pac);
                if (rc) die("meme_chain_set_contigs", rc);
            }
            g_contigs_set = true;
        }
        meme_chain_opt_t co;
        co.w = opt->w; co.max_chain_gap = opt->max_chain_gap; co.max_occ = opt->max_occ; co.min_seed_len = opt->min_seed_len;
        co.min_chain_weight = opt->min_chain_weight; co.max_chain_extend = opt->max_chain_extend;
        co.mask_level = opt->mask_level; co.drop_ratio = opt->drop_ratio;
        g_cshard.resize(g_index.size());
        int rc = a.chain_batch_multi(g_index.data(), (int)g_index.size(), codes.data(), off.data(), n, (opt->flag & MEM_F_PE) ? 1 : 0, &so, &co,
                                     g_cshard.data(), g_shard_begin.data());
        if (rc) die("meme_chain_batch_multi", rc);
        long tot_chains = 0, tot_seeds = 0;
        g_cseed_base.assign(g_cshard.size() + 1, 0);
        for (size_t d = 0; d < g_cshard.size(); d++) {
            tot_chains += (long)g_cshard[d].n_chains; tot_seeds += (long)g_cshard[d].n_seeds;
            g_cseed_base[d + 1] = g_cseed_base[d] + g_cshard[d].n_seeds;
        }
        if ((size_t)tot_seeds + 1 > g_seed_arena_cap) {
            free(g_seed_arena);
            g_seed_arena_cap = (size_t)tot_seeds + (size_t)tot_seeds / 4 + 1024;
            g_seed_arena = (mem_seed_t *)malloc(g_seed_arena_cap * sizeof(mem_seed_t));
            if (!g_seed_arena) { fprintf(stderr, "[meme_b200] out of memory (seed arena)\n"); exit(EXIT_FAILURE); }
        }
        g_base = seqs; g_n = n; g_have_chains = true;
        g_seed_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        g_seed_reads += n;
        auto t2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[meme_b200] seeded and chained %d reads on the GPU: %ld chains, %ld seeds (cumulative %.3f s for %ld reads; this chunk: "
                "codes %.1f ms, meme_chain_batch %.1f ms)\n", n, tot_chains, tot_seeds, g_seed_s, g_seed_reads,
                std::chrono::duration<double>(t1 - t0).count() * 1e3, std::chrono::duration<double>(t2 - t1).count() * 1e3);
        return;
    }
    // pairs stay together: with paired-end input consecutive reads are mates (the chunk holds whole pairs)
    int rc = a.seed_batch_multi(g_index.data(), (int)g_index.size(), codes.data(), off.data(), n, (opt->flag & MEM_F_PE) ? 1 : 0, &so,
                                g_shard.data(), g_shard_begin.data());
    if (rc) die("meme_seed_batch_multi", rc);
    long tot_smems = 0, tot_hits = 0;
    for (auto &r : g_shard) { tot_smems += (long)r.n_smems; tot_hits += (long)r.n_hits; }
    g_base = seqs; g_n = n; g_have = true;
    g_seed_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    g_seed_reads += n;
    auto t2 = std::chrono::steady_clock::now();
    fprintf(stderr, "[meme_b200] seeded %d reads on the GPU: %ld SMEMs, %ld hits (cumulative %.3f s for %ld reads; this chunk: "
            "codes %.1f ms, meme_seed_batch %.1f ms)\n", n, tot_smems, tot_hits, g_seed_s, g_seed_reads,
            std::chrono::duration<double>(t1 - t0).count() * 1e3, std::chrono::duration<double>(t2 - t1).count() * 1e3);
}

// Fill the per-thread kvecs exactly as the reference's seeding + ks_introsort would leave them.
bool meme_glue_take(const bseq1_t *s, mem_tlv *smems, u64v *hits)
{
    if (!g_have) return false;
    long idx = (long)(s - g_base);
    if (idx < 0 || idx >= g_n) return false;
    size_t sh = 0;
    while (sh + 1 < g_shard.size() && idx >= g_shard_begin[sh + 1]) sh++;
    const meme_seed_result_t &g_res = g_shard[sh];
    idx -= (long)g_shard_begin[sh];
    int64_t s0 = g_res.smem_off[idx], s1 = g_res.smem_off[idx + 1];
    int64_t h0 = g_res.hit_off[idx], h1 = g_res.hit_off[idx + 1];
    size_t ns = (size_t)(s1 - s0), nh = (size_t)(h1 - h0);
    if (smems->m < ns) kv_resize(mem_tl, *smems, ns + 16);
    static_assert(sizeof(mem_tl) == sizeof(meme_smem_t), "mem_tl layout");
    memcpy(smems->a, g_res.smems + s0, ns * sizeof(mem_tl));
    smems->n = ns;
    // hits arrive compact (the sampled ones only); mem_chain_Learned indexes hits[hitbeg + k*step], so SMEMs with
    // more than max_occ occurrences get their samples scattered to the strided slots of a full-size list
    size_t full = 0; bool sampled = false;
    for (size_t i = 0; i < ns; i++) {
        size_t c = (size_t)smems->a[i].hitcount;
        full += c;
        if (g_max_occ > 0 && c > (size_t)g_max_occ) sampled = true;
    }
    if (!sampled) {
        if (hits->m < nh) kv_resize(uint64_t, *hits, nh + 16);
        memcpy(hits->a, g_res.hits + h0, nh * sizeof(uint64_t));
        hits->n = nh;
        return true;
    }
    if (hits->m < full) kv_resize(uint64_t, *hits, full + 16);
    const uint64_t *src = g_res.hits + h0;
    size_t beg = 0;
    for (size_t i = 0; i < ns; i++) {
        mem_tl &m = smems->a[i];
        size_t c = (size_t)m.hitcount, mo = (size_t)g_max_occ;
        size_t step = c > mo ? c / mo : 1, take = (c + step - 1) / step;
        if (take > mo) take = mo;
        const uint64_t *q = src + m.hitbeg;
        for (size_t j = 0; j < take; j++) hits->a[beg + j * step] = q[j];
        m.hitbeg = (int)beg;
        beg += c;
    }
    hits->n = full;
    return true;
}

// The read's chains exactly as mem_chain_Learned + mem_chain_flt would leave them in chain_ar[l] (src/bwamem.cpp:L1398-1407);
// true = done, the caller skips its own seeding and chaining of this read.
bool meme_glue_take_chains(bseq1_t *s, int l, mem_chain_v *chn)
{
    if (!g_have_chains) return false;
    long idx = (long)(s - g_base);
    if (idx < 0 || idx >= g_n) return false;
    size_t sh = 0;
    while (sh + 1 < g_cshard.size() && idx >= g_shard_begin[sh + 1]) sh++;
    const meme_chain_result_t &R = g_cshard[sh];
    idx -= (long)g_shard_begin[sh];
    // base codes in place, which the caller's loop does before seeding and later stages rely on (L1277-1280)
    for (int i = 0; i < s->l_seq; ++i) s->seq[i] = s->seq[i] < 4 ? s->seq[i] : nst_nt4_table[(int)s->seq[i]];
    const int64_t c0 = R.chain_off[idx], n = R.chain_off[idx + 1] - c0;
    chn->n = (size_t)n; chn->m = (size_t)n;
    chn->a = n ? (mem_chain_t *)malloc((size_t)n * sizeof(mem_chain_t)) : nullptr;
    if (n && !chn->a) { fprintf(stderr, "[meme_b200] out of memory (chains)\n"); exit(EXIT_FAILURE); }
    const float frac = R.frac_rep[idx];
    for (int64_t i = 0; i < n; i++) {
        const meme_chain_t &g = R.chains[c0 + i];
        mem_chain_t &c = chn->a[i];
        memset(&c, 0, sizeof c);
        c.seqid = l; c.n = g.n; c.m = SEEDS_PER_CHAIN; c.first = g.first; c.rid = g.rid;
        c.w = (uint32_t)g.w; c.kept = (uint32_t)g.kept; c.is_alt = (uint32_t)g.is_alt;
        c.frac_rep = frac; c.pos = g.pos;
        mem_seed_t *dst = g_seed_arena + g_cseed_base[sh] + g.seed_beg;
        const meme_chain_seed_t *src = R.seeds + g.seed_beg;
        for (int j = 0; j < g.n; j++) {
            dst[j].rbeg = src[j].rbeg; dst[j].qbeg = src[j].qbeg; dst[j].len = src[j].len; dst[j].score = src[j].len;
            dst[j].done = 0; dst[j].aln = 0;
        }
        c.seeds = dst;
    }
    return true;
}

// MEME_B200_DUMP_CHAINS=path: one line per read with chain_ar as the chaining stage leaves it -- whichever code produced it
// (with every hook off: the reference's own mem_chain_Learned + mem_chain_flt).  Test instrumentation.
void meme_glue_dump_chains(const bseq1_t *seq, int nseq, const mem_chain_v *chain_ar)
{
    static const char *path = getenv("MEME_B200_DUMP_CHAINS");
    if (!path) return;
    static std::mutex mu;
    static FILE *f = nullptr;
    std::string buf;
    char tmp[256];
    for (int l = 0; l < nseq; l++) {
        const mem_chain_v &v = chain_ar[l];
        snprintf(tmp, sizeof tmp, "%s %zu", seq[l].name, v.n);
        buf += tmp;
        for (size_t i = 0; i < v.n; i++) {
            const mem_chain_t &c = v.a[i];
            snprintf(tmp, sizeof tmp, " [pos=%ld rid=%d n=%d w=%d kept=%d alt=%d first=%d frac=%.6f", (long)c.pos, c.rid, c.n, (int)c.w, (int)c.kept,
                     (int)c.is_alt, c.first, c.frac_rep);
            buf += tmp;
            for (int j = 0; j < c.n; j++) {
                snprintf(tmp, sizeof tmp, " %d,%d,%ld", c.seeds[j].qbeg, c.seeds[j].len, (long)c.seeds[j].rbeg);
                buf += tmp;
            }
            buf += "]";
        }
        buf += "\n";
    }
    std::lock_guard<std::mutex> lk(mu);
    if (!f) f = fopen(path, "w");
    if (!f) { fprintf(stderr, "[meme_b200] cannot write %s\n", path); exit(EXIT_FAILURE); }
    fwrite(buf.data(), 1, buf.size(), f);
    fflush(f);
}

// ---- SAM stage: CIGAR / NM / MD of every region of the chunk from the GPU (MEME_B200_SAM=1) ----
bool meme_glue_cigar_enabled()
{
    return env_on("MEME_B200_SAM") && (g_have || g_have_chains) && !g_index.empty();
}

int64_t meme_glue_code_off(int read) { return g_code_off[(size_t)read]; }

// tasks: every region of the chunk (task_off[i] .. task_off[i+1] are read i's); reg_w: the regions' ar->w.  One
// meme_cigar_batch per device on contiguous shares; the results stay in the handles' pinned arrays until the next chunk.
meme_cigar_task_t *meme_glue_cigar_tasks(int64_t n)
{
    g_ctask_buf.resize((size_t)(n > 0 ? n : 1) * sizeof(meme_cigar_task_t));
    g_ctasks = (meme_cigar_task_t *)g_ctask_buf.data();
    g_n_ctasks = n;
    return g_ctasks;
}

void meme_glue_cigar_run(std::vector<int32_t> &reg_w, std::vector<int64_t> &task_off, const bseq1_t *seqs,
                         int n_reads, const mem_opt_t *opt, int64_t l_pac)
{
    Api &a = api();
    auto t0 = std::chrono::steady_clock::now();
    g_creg_w.swap(reg_w); g_ctask_off.swap(task_off);
    const int64_t n = g_n_ctasks;
    const int nd = (int)g_index.size();
    meme_cigar_opt_t co;
    memset(&co, 0, sizeof co);
    for (int k = 0; k < 25; k++) co.mat[k] = opt->mat[k];
    co.o_del = opt->o_del; co.e_del = opt->e_del; co.o_ins = opt->o_ins; co.e_ins = opt->e_ins; co.w = opt->w; co.a = opt->a;
    g_cres.resize((size_t)nd);
    g_cres_begin.assign((size_t)nd + 1, 0);
    for (int d = 0; d <= nd; d++) g_cres_begin[(size_t)d] = n * d / nd;
    std::vector<int> rc((size_t)nd, 0);
    std::vector<std::string> err((size_t)nd);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++)
        th.emplace_back([&, d] {
            const int64_t b = g_cres_begin[(size_t)d], e = g_cres_begin[(size_t)d + 1];
            rc[(size_t)d] = a.cigar_batch(g_index[(size_t)d], g_ctasks + b, e - b, g_codes.data(), (int64_t)g_code_off[(size_t)n_reads], l_pac, &co,
                                          &g_cres[(size_t)d]);
            if (rc[(size_t)d]) err[(size_t)d] = a.last_error();
        });
    for (auto &t : th) t.join();
    const double t_call = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e3;
    for (int d = 0; d < nd; d++)
        if (rc[(size_t)d]) { fprintf(stderr, "[meme_b200] meme_cigar_batch failed (%d): %s\n", rc[(size_t)d], err[(size_t)d].c_str()); exit(EXIT_FAILURE); }
    size_t slots = 1024;
    while (slots < (size_t)n_reads * 3) slots <<= 1;
    if (g_cslot.size() != slots) { std::vector<std::atomic<int32_t>> t(slots); g_cslot.swap(t); }
    g_cmask = slots - 1;
    g_cseq.resize((size_t)n_reads);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)slots; k++) g_cslot[(size_t)k].store(-1, std::memory_order_relaxed);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_reads; i++) {
        g_cseq[(size_t)i] = seqs[i].seq;
        for (size_t h = ptr_hash(seqs[i].seq) & g_cmask;; h = (h + 1) & g_cmask) {
            int32_t empty = -1;
            if (g_cslot[h].compare_exchange_strong(empty, i, std::memory_order_relaxed)) break;
        }
    }
    g_have_cigar = true;
    static bool at_exit = false;
    if (!at_exit) {
        at_exit = true;
        atexit([] { fprintf(stderr, "[meme_b200] SAM stage: %ld alignments taken from the GPU results, %ld computed by mem_reg2aln itself "
                            "(regions created after the pre-pass: mate rescue)\n", g_cig_hit.load(), g_cig_miss.load()); });
    }
    long ops = 0;
    for (auto &r : g_cres) ops += (long)r.n_ops;
    fprintf(stderr, "[meme_b200] global alignments of %ld regions on the GPU: %ld CIGAR operations, %.1f ms (device calls %.1f ms)\n", (long)n, ops,
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e3, t_call);
}

// mem_reg2aln asks for one region's alignment: true = *cigar (malloc'ed: n_cigar operations followed by the MD string, as
// bwa_gen_cigar2 returns it), *n_cigar and *NM are set and the caller skips its own loop.  Regions the pre-pass has not seen
// (created later by mate rescue) return false and take the caller's own code.
bool meme_glue_take_cigar(const char *query, const mem_alnreg_t *ar, uint32_t **cigar, int *n_cigar, int *NM)
{
    if (!g_have_cigar) return false;
    int i = -1;
    for (size_t h = ptr_hash(query) & g_cmask;; h = (h + 1) & g_cmask) {
        const int32_t v = g_cslot[h].load(std::memory_order_relaxed);
        if (v < 0) break;
        if (g_cseq[(size_t)v] == query) { i = v; break; }
    }
    if (i < 0) { g_cig_miss++; return false; }
    for (int64_t t = g_ctask_off[(size_t)i]; t < g_ctask_off[(size_t)i + 1]; t++) {
        const meme_cigar_task_t &k = g_ctasks[(size_t)t];
        if (k.rb != ar->rb || k.re != ar->re || k.truesc != ar->truesc || g_creg_w[(size_t)t] != ar->w) continue;
        if (k.q_off != g_code_off[(size_t)i] + ar->qb || k.qlen != ar->qe - ar->qb) continue;
        size_t d = 0;
        while (d + 1 < g_cres.size() && t >= g_cres_begin[d + 1]) d++;
        const meme_cigar_result_t &R = g_cres[d];
        const meme_cigar_rec_t &r = R.recs[t - g_cres_begin[d]];
        if (r.n_cigar < 0) { g_cig_miss++; return false; }
        g_cig_hit++;
        uint32_t *c = (uint32_t *)malloc((size_t)r.n_cigar * 4 + (size_t)r.md_len);
        if (!c) { fprintf(stderr, "[meme_b200] out of memory (cigar)\n"); exit(EXIT_FAILURE); }
        memcpy(c, R.cigars + r.cigar_off, (size_t)r.n_cigar * 4);
        memcpy(c + r.n_cigar, R.md + r.md_off, (size_t)r.md_len);
        *cigar = c; *n_cigar = r.n_cigar; *NM = r.nm;
        return true;
    }
    g_cig_miss++;
    return false;
}

// ---- mate rescue: the Smith-Waterman pairs of a paired-end chunk on the GPU (MEME_B200_MATE=1) ----
namespace {
std::vector<meme_kswr_t> g_msw_res;              // results of all pairs of the chunk, in the order the pre-pass produced them
std::vector<int64_t> g_msw_off;                  // per read pair: index of its first result
bool g_have_msw = false;
std::atomic<long> g_msw_taken{0}, g_msw_blocks{0};
static_assert(sizeof(meme_kswr_t) == sizeof(kswr_t) && offsetof(meme_kswr_t, qb) == offsetof(kswr_t, qb) && offsetof(meme_kswr_t, te2) == offsetof(kswr_t, te2),
              "meme_kswr_t must mirror kswr_t (src/ksw.h)");
}

bool meme_glue_matesw_enabled() { return env_on("MEME_B200_MATE") && !g_index.empty(); }
void meme_glue_matesw_skip() { g_have_msw = false; }
bool meme_glue_matesw_have() { return g_have_msw; }
const std::vector<int64_t> &meme_glue_matesw_off() { return g_msw_off; }

// nseg segments written by the host threads (contiguous shares of the read pairs, in order) -> one device batch per GPU
void meme_glue_matesw_run(int nseg, SeqPair **pairs, int64_t *n_pairs, uint8_t **ref, int64_t *n_ref, uint8_t **qer, int64_t *n_qer, const mem_opt_t *opt,
                          std::vector<int64_t> &pair_off, double collect_ms)
{
    Api &a = api();
    auto t0 = std::chrono::steady_clock::now();
    g_have_msw = false;
    g_msw_off.swap(pair_off);
    int64_t n = 0;
    std::vector<int64_t> base((size_t)nseg + 1, 0);
    for (int s = 0; s < nseg; s++) { base[(size_t)s] = n; n += n_pairs[s]; }
    base[(size_t)nseg] = n;
    if (g_msw_off.empty() || g_msw_off.back() != n) { fprintf(stderr, "[meme_b200] mate rescue: %ld pairs collected, %ld counted\n", (long)n, g_msw_off.empty() ? -1L : (long)g_msw_off.back()); exit(EXIT_FAILURE); }
    g_msw_res.resize((size_t)n + 1);
    meme_matesw_opt_t mo = {opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, opt->a, opt->b};
    const int nd = (int)g_index.size();
    std::vector<int> rc((size_t)nd, 0);
    std::vector<std::string> err((size_t)nd);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++)
        th.emplace_back([&, d] {
            const int s0 = nseg * d / nd, s1 = nseg * (d + 1) / nd;        // whole segments per device
            if (s1 <= s0) return;
            std::vector<meme_kswr_t *> out((size_t)(s1 - s0));
            for (int s = s0; s < s1; s++) out[(size_t)(s - s0)] = g_msw_res.data() + base[(size_t)s];
            rc[(size_t)d] = a.matesw_batch(g_index[(size_t)d], s1 - s0, (const meme_seqpair_t *const *)(pairs + s0), n_pairs + s0, ref + s0, n_ref + s0, qer + s0, n_qer + s0,
                                           &mo, out.data());
            if (rc[(size_t)d]) err[(size_t)d] = a.last_error();
        });
    for (auto &t : th) t.join();
    for (int d = 0; d < nd; d++)
        if (rc[(size_t)d]) { fprintf(stderr, "[meme_b200] meme_matesw_batch failed (%d): %s\n", rc[(size_t)d], err[(size_t)d].c_str()); exit(EXIT_FAILURE); }
    g_have_msw = true;
    static bool at_exit = false;
    if (!at_exit) {
        at_exit = true;
        atexit([] { fprintf(stderr, "[meme_b200] mate rescue: %ld Smith-Waterman results in %ld blocks of worker_sam taken from the GPU\n", g_msw_taken.load(), g_msw_blocks.load()); });
    }
    int64_t nr = 0;
    for (int s = 0; s < nseg; s++) nr += n_ref[s];
    fprintf(stderr, "[meme_b200] mate rescue: %ld Smith-Waterman pairs (%.0f MB of windows) on the GPU: collect %.1f ms, align %.1f ms\n", (long)n, nr / 1e6, collect_ms,
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e3);
}

// worker_sam, where it would call mem_sam_pe_batch for the read pairs [pair_begin, pair_end) of a block: aln[regid] for the
// block's pcnt pairs.  false = nothing from the GPU for this chunk (the caller runs the reference's own batch step).
bool meme_glue_take_matesw(int pair_begin, int pair_end, int64_t pcnt, kswr_t *aln)
{
    if (!g_have_msw) return false;
    const int64_t b = g_msw_off[(size_t)pair_begin], e = g_msw_off[(size_t)pair_end];
    if (e - b != pcnt) { fprintf(stderr, "[meme_b200] mate rescue: block of read pairs %d..%d has %ld pairs, the pre-pass saw %ld\n", pair_begin, pair_end, (long)pcnt, (long)(e - b)); exit(EXIT_FAILURE); }
    if (pcnt) memcpy(aln, g_msw_res.data() + b, (size_t)pcnt * sizeof(kswr_t));
    g_msw_taken += (long)pcnt; g_msw_blocks++;
    return true;
}

// Seed extension through the GPU: same contract as getScores8/16 / scalarBandedSWAWrapper (fills the six
// result fields of every pair in place).  One handle per host thread (kt_for worker).
bool meme_glue_bsw(SeqPair *pairs, uint8_t *seq_buf_ref, uint8_t *seq_buf_qer, int n_pairs, int w, int o_del, int e_del,
                   int o_ins, int e_ins, int zdrop, int end_bonus, int a, int b)
{
    if (!env_on("MEME_B200_BSW")) return false;
    if (n_pairs <= 0) return true;
    Api &A = api();
    static thread_local meme_bsw_t *h = nullptr;
    static thread_local std::vector<meme_seqpair_t> tmp;
    if (!h) {
        int dev = g_devices.empty() ? (getenv("MEME_B200_DEVICE") ? atoi(getenv("MEME_B200_DEVICE")) : 0) : g_devices[0];
        int rc = A.bsw_create(dev, &h);
        if (rc) die("meme_bsw_create", rc);
    }
    static_assert(sizeof(SeqPair) == sizeof(meme_seqpair_t), "SeqPair layout");
    // the pairs address slices of per-thread sequence buffers: rebase to the touched range
    int64_t r0 = INT64_MAX, r1 = 0, q0 = INT64_MAX, q1 = 0;
    for (int i = 0; i < n_pairs; i++) {
        const SeqPair &p = pairs[i];
        if (p.idr < r0) r0 = p.idr;
        if ((int64_t)p.idr + p.len1 > r1) r1 = (int64_t)p.idr + p.len1;
        if (p.idq < q0) q0 = p.idq;
        if ((int64_t)p.idq + p.len2 > q1) q1 = (int64_t)p.idq + p.len2;
    }
    tmp.resize((size_t)n_pairs);
    memcpy(tmp.data(), pairs, (size_t)n_pairs * sizeof(SeqPair));
    for (int i = 0; i < n_pairs; i++) { tmp[i].idr -= (int32_t)r0; tmp[i].idq -= (int32_t)q0; }
    meme_bsw_opt_t o = {o_del, e_del, o_ins, e_ins, zdrop, end_bonus, a, b};
    int rc = A.bsw_batch(h, tmp.data(), n_pairs, seq_buf_ref + r0, r1 - r0, seq_buf_qer + q0, q1 - q0, w, &o);
    if (rc) die("meme_bsw_batch", rc);
    for (int i = 0; i < n_pairs; i++) {
        pairs[i].score = tmp[i].score; pairs[i].tle = tmp[i].tle; pairs[i].gtle = tmp[i].gtle;
        pairs[i].qle = tmp[i].qle; pairs[i].gscore = tmp[i].gscore; pairs[i].max_off = tmp[i].max_off;
    }
    return true;
}

// ---- chunk-level seed extension (meme_extend.inl, included by the patched bwamem.cpp) ----
bool meme_glue_bsw_enabled() { return env_on("MEME_B200_BSW"); }

// Reference by position: with the index resident on the GPUs (MEME_B200=1) the extension pairs name their reference side
// as a position in that text and nothing of it is uploaded.  Off when pairs are dumped (the dump carries bytes) or when
// MEME_B200_BSW_BYTES=1 asks for the byte form.
bool meme_glue_bsw_text_mode()
{
    return !g_index.empty() && !getenv("MEME_B200_DUMP_PAIRS") && !env_on("MEME_B200_BSW_BYTES");
}

// Pinned blocks reserved where the index is loaded (outside the aligner's per-chunk timers): page-locking the ~2 GB of
// pair arenas a chunk needs costs about a second, which the first chunk would otherwise pay.
namespace {
struct PinnedBlock { void *p; size_t bytes; };
std::vector<PinnedBlock> g_pin_cache;
std::mutex g_pin_mu;
}

void meme_glue_pinned_reserve(int threads, long reads_per_chunk)
{
    if (threads < 1 || reads_per_chunk < 1) return;
    auto t0 = std::chrono::steady_clock::now();
    // the same sizes meme_extend.inl asks for on its first chunk (6 pairs per read of head-room, 56 + 200 + 100 + 4 bytes each;
    // per host thread two sides x two halves of the chunk)
    const size_t est = (size_t)(reads_per_chunk / threads + 1) * 6 / 2;
    const size_t sizes[4] = {est * sizeof(SeqPair), est * 200, est * 100, est * 4};
    size_t total = 0;
    for (int t = 0; t < threads * 4; t++)
        for (size_t b : sizes) {
            if (b == est * 200 && meme_glue_bsw_text_mode()) continue;     // no reference bytes: pairs name text positions
            void *p = nullptr;
            if (api().pinned_alloc(b, &p)) return;          // not fatal: the arenas fall back to allocating on demand
            std::lock_guard<std::mutex> lk(g_pin_mu);
            g_pin_cache.push_back({p, b});
            total += b;
        }
    fprintf(stderr, "[meme_b200] %.2f GB of pinned pair arenas reserved for %d threads in %.2f s\n", total / 1e9, threads,
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
}

// per-thread arenas of the mate-rescue pre-pass (meme_matesw_chunk asks for the same sizes on its first chunk)
void meme_glue_matesw_arena_sizes(int threads, long reads_per_chunk, size_t sizes[3])
{
    const size_t est = (size_t)(reads_per_chunk / (threads > 0 ? threads : 1) + 1);      // about two rescue pairs per read pair
    sizes[0] = est * sizeof(SeqPair); sizes[1] = est * 700; sizes[2] = est * 160;
}

void meme_glue_pinned_reserve_matesw(int threads, long reads_per_chunk)
{
    if (threads < 1 || reads_per_chunk < 1) return;
    auto t0 = std::chrono::steady_clock::now();
    size_t sizes[3], total = 0;
    meme_glue_matesw_arena_sizes(threads, reads_per_chunk, sizes);
    for (int t = 0; t < threads; t++)
        for (size_t b : sizes) {
            void *p = nullptr;
            if (api().pinned_alloc(b, &p)) return;
            std::lock_guard<std::mutex> lk(g_pin_mu);
            g_pin_cache.push_back({p, b});
            total += b;
        }
    fprintf(stderr, "[meme_b200] %.2f GB of pinned mate-rescue arenas reserved for %d threads in %.2f s\n", total / 1e9, threads,
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
}

void *meme_glue_pinned_alloc(size_t bytes)
{
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        size_t best = (size_t)-1;
        for (size_t i = 0; i < g_pin_cache.size(); i++)
            if (g_pin_cache[i].bytes >= bytes && (best == (size_t)-1 || g_pin_cache[i].bytes < g_pin_cache[best].bytes)) best = i;   // best fit
        if (best != (size_t)-1) { void *p = g_pin_cache[best].p; g_pin_cache.erase(g_pin_cache.begin() + (long)best); return p; }
    }
    void *p = nullptr;
    int rc = api().pinned_alloc(bytes, &p);
    if (rc) die("meme_pinned_alloc", rc);
    return p;
}

void meme_glue_pinned_free(void *p) { api().pinned_free(p); }

// One side of the extension of a whole chunk: segments written by the host threads -> one device batch with the band
// retries decided by the kernel -> results back into the same pair arrays (and the band width that was final).
// one kernel-B handle per device, created on first use (or by the warm-up at index load)
static std::vector<meme_bsw_t *> &bsw_handles()
{
    static std::vector<meme_bsw_t *> hs;
    if (hs.empty()) {
        Api &A = api();
        if (g_devices.empty()) g_devices.push_back(getenv("MEME_B200_DEVICE") ? atoi(getenv("MEME_B200_DEVICE")) : 0);
        for (int dev : g_devices) {
            meme_bsw_t *h = nullptr;
            int rc = A.bsw_create(dev, &h);
            if (rc) die("meme_bsw_create", rc);
            if (meme_glue_bsw_text_mode()) {
                rc = A.bsw_attach_text(h, g_index[hs.size()]);
                if (rc) die("meme_bsw_attach_text", rc);
            }
            hs.push_back(h);
        }
    }
    return hs;
}

// Kernel B loaded and its device buffers sized for a chunk before the first timed chunk: one throw-away batch of short dummy
// pairs per device, about as many as a chunk of `reads` reads produces on its larger (right) side.
void meme_glue_bsw_warmup(long reads)
{
    if (reads < 1) return;
    auto t0 = std::chrono::steady_clock::now();
    Api &A = api();
    std::vector<meme_bsw_t *> &hs = bsw_handles();
    const int nd = (int)hs.size();
    const int64_t n = (int64_t)reads * 9 / 2 / nd + 1024;           // ~4.1 right-side pairs per read, with head-room
    const int len = 80;
    std::vector<meme_seqpair_t> pairs((size_t)n);
    std::vector<uint8_t> qer((size_t)n * len, 0), ref;
    const bool text = meme_glue_bsw_text_mode();
    if (!text) ref.assign((size_t)n * len, 1);
    memset(pairs.data(), 0, pairs.size() * sizeof(meme_seqpair_t));
    for (int64_t k = 0; k < n; k++) {
        meme_seqpair_t &p = pairs[(size_t)k];
        p.idq = (int32_t)(k * len); p.len1 = len; p.len2 = len; p.h0 = 20; p.regid = (int32_t)k;
        p.idr = text ? (int32_t)(k & 255) : (int32_t)(k * len); p.id = 0;
    }
    meme_bsw_opt_t o = {6, 1, 6, 1, 100, 5, 1, 4};
    for (int d = 0; d < nd; d++) {
        const meme_seqpair_t *pp = pairs.data(); meme_seqpair_t *pw = pairs.data();
        const int64_t np = n, nr = text ? 0 : (int64_t)ref.size(), nq = (int64_t)qer.size();
        const uint8_t *pr = text ? nullptr : ref.data(), *pq = qer.data();
        std::vector<int32_t> wu((size_t)n);
        int32_t *pwu = wu.data();
        int rc = A.bsw_set_match(hs[(size_t)d], 1);
        if (!rc) rc = A.bsw_upload_segments(hs[(size_t)d], 1, &pp, &np, &pr, &nr, &pq, &nq);
        if (!rc) rc = A.bsw_run_ext(hs[(size_t)d], 100, 2, 1, &o);
        if (!rc) rc = A.bsw_download_segments(hs[(size_t)d], 1, &pw, &np, &pwu);
        if (rc) die("kernel-B warm-up", rc);
    }
    fprintf(stderr, "[meme_b200] kernel-B warm-up batch of %ld pairs per device in %.2f s\n", (long)n,
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
}

namespace {
struct BswJob {          // one extension batch in flight: a driving host thread per device, and copies of the caller's segment arrays
    std::vector<std::thread> th;
    std::vector<int> rcs;
    std::vector<std::string> errs;
    std::vector<SeqPair *> pairs; std::vector<int64_t> n_pairs, n_ref, n_qer; std::vector<uint8_t *> ref, qer; std::vector<int32_t *> w_used;
};
}

void *meme_glue_bsw_segments_start(int nseg, SeqPair **pairs, int64_t *n_pairs, uint8_t **ref, int64_t *n_ref, uint8_t **qer, int64_t *n_qer,
                            int w, int tries, int prev_is_h0, int o_del, int e_del, int o_ins, int e_ins, int zdrop, int end_bonus,
                            int a, int b, int32_t **w_used)
{
    Api &A = api();
    std::vector<meme_bsw_t *> &hs = bsw_handles();
    static_assert(sizeof(SeqPair) == sizeof(meme_seqpair_t), "SeqPair layout");
    meme_bsw_opt_t o = {o_del, e_del, o_ins, e_ins, zdrop, end_bonus, a, b};
    // MEME_B200_DUMP_PAIRS=<path>: the extension pairs of the first chunk (left side, then right side with its real h0)
    // are written out once, for the kernel-B bench and its CPU baseline (bench.py, profiles/bench_bsw.py)
    static int dumped = 0;
    if (const char *dp = getenv("MEME_B200_DUMP_PAIRS")) {
        if (dumped < 2) {
            std::string path = std::string(dp) + (prev_is_h0 ? ".right" : ".left");
            FILE *f = fopen(path.c_str(), "wb");
            if (f) {
                int64_t tp = 0, tr = 0, tq = 0;
                for (int s = 0; s < nseg; s++) { tp += n_pairs[s]; tr += n_ref[s]; tq += n_qer[s]; }
                int64_t hdr[3] = {tp, tr, tq};
                int32_t par[12] = {w, tries, prev_is_h0, o_del, e_del, o_ins, e_ins, zdrop, end_bonus, a, b, 0};
                fwrite(hdr, 8, 3, f); fwrite(par, 4, 12, f);
                int64_t rb = 0, qb = 0;
                for (int s = 0; s < nseg; s++) {
                    for (int64_t i = 0; i < n_pairs[s]; i++) {
                        SeqPair p = pairs[s][i];
                        p.idr += (int32_t)rb; p.idq += (int32_t)qb;
                        fwrite(&p, sizeof p, 1, f);
                    }
                    rb += n_ref[s]; qb += n_qer[s];
                }
                for (int s = 0; s < nseg; s++) fwrite(ref[s], 1, (size_t)n_ref[s], f);
                for (int s = 0; s < nseg; s++) fwrite(qer[s], 1, (size_t)n_qer[s], f);
                fclose(f);
            }
            dumped++;
        }
    }
    // the segments (one per host thread) are dealt round-robin to the devices; every device runs its share concurrently, driven
    // by one host thread each; the caller may do other work until it waits for the job
    const int nd = (int)hs.size();
    BswJob *J = new BswJob;
    J->rcs.assign((size_t)nd, 0); J->errs.resize((size_t)nd);
    J->pairs.assign(pairs, pairs + nseg); J->n_pairs.assign(n_pairs, n_pairs + nseg); J->ref.assign(ref, ref + nseg); J->n_ref.assign(n_ref, n_ref + nseg);
    J->qer.assign(qer, qer + nseg); J->n_qer.assign(n_qer, n_qer + nseg); J->w_used.assign(w_used, w_used + nseg);
    std::vector<meme_bsw_t *> *hp = &hs;
    for (int d = 0; d < nd; d++) {
        J->th.emplace_back([J, hp, d, nd, nseg, a, w, tries, prev_is_h0, o] {
            Api &A = api();
            std::vector<const meme_seqpair_t *> sp; std::vector<meme_seqpair_t *> spw; std::vector<int64_t> np, nr, nq;
            std::vector<const uint8_t *> pr, pq; std::vector<int32_t *> pw;
            for (int s = d; s < nseg; s += nd) {
                sp.push_back((const meme_seqpair_t *)J->pairs[(size_t)s]); spw.push_back((meme_seqpair_t *)J->pairs[(size_t)s]); np.push_back(J->n_pairs[(size_t)s]);
                pr.push_back(J->ref[(size_t)s]); nr.push_back(J->n_ref[(size_t)s]); pq.push_back(J->qer[(size_t)s]); nq.push_back(J->n_qer[(size_t)s]);
                pw.push_back(J->w_used[(size_t)s]);
            }
            const int k = (int)sp.size();
            meme_bsw_t *h = (*hp)[(size_t)d];
            int rc = A.bsw_set_match(h, a);
            if (!rc) rc = A.bsw_upload_segments(h, k, sp.data(), np.data(), pr.data(), nr.data(), pq.data(), nq.data());
            if (!rc) rc = A.bsw_run_ext(h, w, tries, prev_is_h0, &o);
            if (!rc) rc = A.bsw_download_segments(h, k, spw.data(), np.data(), pw.data());
            J->rcs[(size_t)d] = rc;
            if (rc) J->errs[(size_t)d] = A.last_error();
        });
    }
    return J;
}

void meme_glue_bsw_segments_wait(void *job)
{
    BswJob *J = (BswJob *)job;
    for (auto &t : J->th) t.join();
    for (size_t d = 0; d < J->rcs.size(); d++)
        if (J->rcs[d]) { fprintf(stderr, "[meme_b200] seed extension on GPU %d failed (%d): %s\n", g_devices[d], J->rcs[d], J->errs[d].c_str()); exit(EXIT_FAILURE); }
    delete J;
}

void meme_glue_bsw_segments(int nseg, SeqPair **pairs, int64_t *n_pairs, uint8_t **ref, int64_t *n_ref, uint8_t **qer, int64_t *n_qer,
                            int w, int tries, int prev_is_h0, int o_del, int e_del, int o_ins, int e_ins, int zdrop, int end_bonus,
                            int a, int b, int32_t **w_used)
{
    meme_glue_bsw_segments_wait(meme_glue_bsw_segments_start(nseg, pairs, n_pairs, ref, n_ref, qer, n_qer, w, tries, prev_is_h0, o_del, e_del, o_ins, e_ins,
                                                             zdrop, end_bonus, a, b, w_used));
}
