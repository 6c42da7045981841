// oracle/ref_cigar_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Calls the UNMODIFIED reference's mem_reg2aln (src/bwamem.cpp:L2314-2387; linked from the reference's own objects by
// oracle/Makefile.ref) on a file of regions and prints, one JSON line per region, what it worked out: CIGAR (with the
// clipping and the squeezed leading/trailing deletion mem_reg2aln applies), NM, MD, position and strand.  Used by
// tests/golden/make_golden.py to generate tests/golden/cigar_*.json and by tests/test_oracle_cigar.py to re-check them live.
//
// usage: cigar_shim <index prefix> <cases.bin>
//   cases.bin: int64 n_cases, int64 n_codes, uint8 codes[n_codes] (reads, codes 0..4), then n_cases records
//              { int64 rb, re, read_off; int32 l_query, qb, qe, truesc, reg_w, pad }
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include "bwamem.h"
#include "bntseq.h"
#include "macro.h"

// the reference's profiling counters live in its main.cpp (src/main.cpp:L42), which is not linked here
uint64_t proc_freq, tprof[LIM_R][LIM_C], prof[LIM_R];

struct Case { int64_t rb, re, read_off; int32_t l_query, qb, qe, truesc, reg_w, pad; };

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <prefix> <cases.bin>\n", argv[0]); return 2; }
    bntseq_t *bns = bns_restore(argv[1]);
    if (!bns) return 3;
    std::vector<uint8_t> pac((size_t)bns->l_pac / 4 + 1, 0);
    if (fread(pac.data(), 1, pac.size(), bns->fp_pac) == 0) return 4;
    FILE *f = fopen(argv[2], "rb");
    if (!f) return 5;
    int64_t n = 0, n_codes = 0;
    if (fread(&n, 8, 1, f) != 1 || fread(&n_codes, 8, 1, f) != 1) return 6;
    std::vector<uint8_t> codes((size_t)n_codes);
    if (n_codes && fread(codes.data(), 1, (size_t)n_codes, f) != (size_t)n_codes) return 6;
    std::vector<Case> cases((size_t)n);
    if (n && fread(cases.data(), sizeof(Case), (size_t)n, f) != (size_t)n) return 6;
    fclose(f);
    mem_opt_t *opt = mem_opt_init();
    for (int64_t i = 0; i < n; i++) {
        const Case &c = cases[(size_t)i];
        mem_alnreg_t ar;
        memset(&ar, 0, sizeof ar);
        ar.rb = c.rb; ar.re = c.re; ar.qb = c.qb; ar.qe = c.qe; ar.truesc = c.truesc; ar.score = c.truesc; ar.w = c.reg_w;
        ar.secondary = -1; ar.rid = bns_intv2rid(bns, c.rb, c.re);
        mem_aln_t a = mem_reg2aln(opt, bns, pac.data(), c.l_query, (const char *)codes.data() + c.read_off, &ar);
        printf("{\"n_cigar\": %d, \"cigar\": [", a.n_cigar);
        for (int k = 0; k < a.n_cigar; k++) printf("%s%u", k ? ", " : "", a.cigar[k]);
        printf("], \"nm\": %d, \"md\": \"%s\", \"pos\": %ld, \"rev\": %d, \"rid\": %d}\n", a.NM, (char *)(a.cigar + a.n_cigar), (long)a.pos, (int)a.is_rev, a.rid);
        free(a.cigar);
    }
    free(opt);
    return 0;
}
