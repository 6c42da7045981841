// chain_host.cpp -- TEST INFRASTRUCTURE.  The per-read chaining procedure of the product (bwa-meme_b200/csrc/meme_chain_core.cuh,
// __host__ __device__) compiled with g++ and run read by read on the CPU with the scratch layout of the kernels
// (meme_chain.cu: read_slices, k_chain_build, k_chain_emit), so that the B-tree, the sort and the filter can be checked against
// the oracle and the reference dumps without a GPU.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../bwa-meme_b200/csrc/meme_chain_core.cuh"

using namespace chn;

// Inputs as the device sees them: SMEM records with hitbeg into the read's SAMPLED hit list.  Outputs sized by the caller:
// chain_off[n+1], frac[n], chains[<= total hits], seeds[<= total hits].  Returns 0, or the kernel's error code.
extern "C" int chain_sim(const meme_smem_t *smems, const int64_t *smem_off, const uint64_t *hits, const int64_t *hit_off, const int32_t *read_len,
                         int64_t n_reads, const meme_contig_t *contigs, int32_t n_contigs, int64_t l_pac, const meme_chain_opt_t *opt,
                         int64_t *chain_off, float *frac, meme_chain_t *chains, meme_chain_seed_t *seeds)
{
    const int64_t nh = hit_off[n_reads];
    std::vector<cu32> next((size_t)nh + 1), ql((size_t)nh + 1), order((size_t)nh + 1), kept((size_t)nh + 1);
    std::vector<Rec> ch((size_t)nh + 1);
    std::vector<Node> nodes((size_t)nh / 4 + 2 * (size_t)n_reads + 4);
    int64_t nc = 0, ns = 0;
    for (int64_t r = 0; r < n_reads; r++) {
        const int64_t s0 = smem_off[r], h0 = hit_off[r], h1 = hit_off[r + 1];
        View v; Scratch S;
        v.smems = smems + s0; v.n_smems = (int)(smem_off[r + 1] - s0);
        v.hits = (const cu64 *)(hits + h0);
        v.read_len = read_len[r];
        v.contigs = contigs; v.n_contigs = n_contigs; v.l_pac = l_pac;
        S.next = next.data() + h0; S.ql = ql.data() + h0; S.ch = ch.data() + h0; S.order = order.data() + h0; S.kept = kept.data() + h0;
        S.nodes = nodes.data() + (h0 / 4 + 2 * r);
        S.max_nodes = (int)((h1 - h0) / 4 + 2);
        int n_out = 0, err = 0;
        chain_read(*opt, v, S, &n_out, &frac[r], &err);
        if (err) return err;
        chain_off[r] = nc;
        for (int i = 0; i < n_out; i++) {
            const Rec &c = S.ch[S.order[i]];
            meme_chain_t o;
            o.pos = (int64_t)v.hits[c.first_seed]; o.seed_beg = ns; o.n = c.n; o.rid = c.rid; o.w = c.w; o.first = c.first;
            o.kept = c.kept; o.is_alt = c.is_alt;
            chains[nc++] = o;
            cu32 s = c.first_seed;
            for (int j = 0; j < c.n; j++, s = S.next[s]) {
                seeds[ns].rbeg = (int64_t)v.hits[s]; seeds[ns].qbeg = (int32_t)(S.ql[s] & 0xffff); seeds[ns].len = (int32_t)(S.ql[s] >> 16);
                ns++;
            }
        }
    }
    chain_off[n_reads] = nc;
    return 0;
}
