// cigar_host.cpp -- TEST INFRASTRUCTURE.  The per-region global alignment of the product (bwa-meme_b200/csrc/meme_cigar_core.cuh,
// __host__ __device__) compiled with g++ and run region by region on the CPU with the scratch discipline of the kernels
// (meme_cigar.cu: DP row reused as the reversed operation list, backtrack matrix sized by z_bytes, two passes over the CIGAR for
// NM / MD length and the MD text), so that it can be checked against the oracle and the reference's outputs without a GPU.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../bwa-meme_b200/csrc/meme_cigar_core.cuh"

using namespace cig;

// text: the .0123 codes (2 * l_pac).  Outputs sized by the caller: recs[n], cigars / md large enough (ops <= 2*qlen+1 per task,
// md <= 4 * tlen + 16).  stride / lane: the scratch arrays of a region are laid out as lane `lane` of `stride` interleaved
// regions (32 on the device, 1 = contiguous).  Returns 0, or -1 if a region needed more backtrack bytes than z_bytes() promised.
extern "C" int cigar_sim(const uint8_t *text, int64_t l_pac, const uint8_t *codes, int64_t n_codes, const meme_cigar_task_t *tasks, int64_t n,
                         const meme_cigar_opt_t *opt, meme_cigar_rec_t *recs, uint32_t *cigars, char *md, uint64_t *cells_out, int stride, int lane)
{
    std::vector<gu64> pac((size_t)(2 * l_pac + 31) / 32 + 2, 0);
    for (int64_t i = 0; i < 2 * l_pac; i++) pac[(size_t)(i >> 5)] |= (gu64)(text[i] & 3) << (2 * (i & 31));
    int64_t no = 0, nm_off = 0;
    gu64 cells = 0;
    for (int64_t i = 0; i < n; i++) {
        const meme_cigar_task_t &t = tasks[i];
        meme_cigar_rec_t r;
        r.score = 0; r.nm = -1; r.n_cigar = -1; r.md_len = 0; r.cigar_off = no; r.md_off = nm_off;
        const bool ok = !(t.qlen <= 0 || t.rb >= t.re || (t.rb < l_pac && t.re > l_pac)) && t.rb >= 0 && t.re <= 2 * l_pac &&
                        t.q_off >= 0 && t.q_off + t.qlen <= n_codes;
        if (ok) {
            const int tlen = (int)(t.re - t.rb);
            const int64_t zb = z_bytes(*opt, t.qlen, tlen, t.w);
            std::vector<EH> eh(((size_t)t.qlen + 1) * (size_t)stride);
            std::vector<uint8_t> q(((size_t)t.qlen + 1) * (size_t)stride, 0xee);
            std::vector<uint8_t> z(((size_t)zb + 1) * (size_t)stride, 0xee);     // other lanes' bytes must stay untouched
            Scr C;
            C.eh = eh.data() + lane; C.q = q.data() + lane; C.z = z.data() + lane; C.stride = stride;
            Seqs S;
            S.qs = C.q; S.stride = stride; S.pac = pac.data(); S.rb = t.rb; S.qlen = t.qlen; S.tlen = tlen; S.rev = t.rb >= l_pac;
            fill_query(codes, t.q_off, t.qlen, S.rev, C);
            int n_ops = 0;
            r.score = region_cigar(*opt, S, t.w, t.truesc, C, &n_ops, &cells);
            for (size_t k = 0; k < z.size(); k++)
                if ((int)(k % (size_t)stride) != lane || k / (size_t)stride >= (size_t)zb) { if (z[k] != 0xee) return -1; }
            r.n_cigar = n_ops;
            r.md_len = nm_md(S, C, n_ops, nullptr, &r.nm);
            for (int k = 0; k < n_ops; k++) cigars[no + k] = C.op(n_ops - 1 - k);
            int nm2;
            if (nm_md(S, C, n_ops, md + nm_off, &nm2) != r.md_len || nm2 != r.nm) return -2;
            no += n_ops; nm_off += r.md_len;
        }
        recs[i] = r;
    }
    if (cells_out) *cells_out = cells;
    return 0;
}
