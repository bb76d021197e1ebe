// synth_lib.cpp -- C API over synth_core.h (libppp_synth.so): generates the packed ppp_read
// records of a synthetic data set directly (no SAM/BAM round trip), for bench.py and tests.
// The records equal what the front end (frontend.h) extracts from the file ppp_synth writes
// for the same (genome, reads, seed); tests/test_frontend.py checks that.
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "ppp_types.h"
#include "synth_core.h"

using namespace ppp_synth;

struct ppp_synth_model {
    Model m;
};

static inline bool packRead(const Model& m, const Read& r, uint32_t minLen, uint32_t maxLen, int mode, uint64_t gLo, uint64_t gHiPlus, uint64_t gHiMinus, ppp_read& out) {
    if (r.unmapped) return false;
    uint32_t aln = r.alnLen();
    if (aln < minLen || aln > maxLen) return false;
    uint32_t nh = (mode == PPP_MULTIHITS_UNIQUE) ? 1u : r.nh;
    if (mode == PPP_MULTIHITS_DISCARD && nh != 1) return false;
    uint32_t key = r.minus ? (uint32_t)r.beginPos + aln : (uint32_t)r.beginPos;
    uint64_t g = m.cum[r.contig] + key;
    if (g < gLo || g >= (r.minus ? gHiMinus : gHiPlus)) return false;
    out.key = key;
    out.meta = (nh << PPP_READ_NH_SHIFT) | (r.a10 ? PPP_READ_A10 : 0u) | (r.minus ? PPP_READ_MINUS : 0u);
    return true;
}

extern "C" {

PPP_API ppp_synth_model* ppp_synth_new(const char* genome, int n_contigs, const uint64_t* lens, uint64_t seed, uint64_t n_reads, int nh1) {
    std::vector<std::string> names;
    std::vector<uint64_t> l;
    if (genome && genome[0]) {
        if (!preset(genome, names, l)) return nullptr;
    } else {
        for (int i = 0; i < n_contigs; ++i) { names.push_back("ctg" + std::to_string(i + 1)); l.push_back(lens[i]); }
    }
    ppp_synth_model* s = new ppp_synth_model;
    s->m.init(names, l, seed, n_reads);
    s->m.nh1 = nh1 != 0;
    return s;
}
PPP_API void ppp_synth_free(ppp_synth_model* s) { delete s; }
PPP_API int ppp_synth_n_contigs(const ppp_synth_model* s) { return (int)s->m.lens.size(); }
PPP_API uint64_t ppp_synth_contig_length(const ppp_synth_model* s, int c) { return s->m.lens[c]; }
PPP_API const char* ppp_synth_contig_name(const ppp_synth_model* s, int c) { return s->m.names[c].c_str(); }
PPP_API uint64_t ppp_synth_n_intervals(const ppp_synth_model* s) { return s->m.intervals.size(); }
/* Transposon-like intervals (0-based, end exclusive), 3 x uint32 each: contig, start, end. */
PPP_API void ppp_synth_intervals(const ppp_synth_model* s, uint32_t* out) {
    for (size_t k = 0; k < s->m.intervals.size(); ++k) {
        out[3 * k] = (uint32_t)s->m.intervals[k].contig;
        out[3 * k + 1] = s->m.intervals[k].start;
        out[3 * k + 2] = s->m.intervals[k].start + s->m.intervals[k].len;
    }
}

/* Packed reads of read indices [i0, i1) whose key lies in the global (concatenated, unpadded) coordinate
 * range [g_lo, g_hi_plus) for + reads / [g_lo, g_hi_minus) for - reads, per contig and in index order.
 * Pass out == NULL to only count (counts[c] = reads of contig c); otherwise out[c] must hold counts[c] records. */
PPP_API int ppp_synth_generate(const ppp_synth_model* s, uint64_t i0, uint64_t i1, uint64_t g_lo, uint64_t g_hi_plus, uint64_t g_hi_minus, uint32_t min_len,
                               uint32_t max_len, int mode, int threads, uint64_t* counts, ppp_read** out) {
    const Model& m = s->m;
    size_t nc = m.lens.size();
    unsigned nt = threads > 0 ? (unsigned)threads : std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::vector<uint64_t> > cnt(nt, std::vector<uint64_t>(nc, 0));
    auto range = [&](unsigned t, uint64_t& lo, uint64_t& hi) { lo = i0 + (i1 - i0) * t / nt; hi = i0 + (i1 - i0) * (t + 1) / nt; };
    {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; ++t)
            pool.emplace_back([&, t]() {
                uint64_t lo, hi;
                range(t, lo, hi);
                Read r;
                ppp_read p;
                for (uint64_t i = lo; i < hi; ++i) {
                    m.read(i, r);
                    if (packRead(m, r, min_len, max_len, mode, g_lo, g_hi_plus, g_hi_minus, p)) cnt[t][r.contig]++;
                }
            });
        for (auto& th : pool) th.join();
    }
    for (size_t c = 0; c < nc; ++c) { counts[c] = 0; for (unsigned t = 0; t < nt; ++t) counts[c] += cnt[t][c]; }
    if (!out) return 0;
    std::vector<std::vector<uint64_t> > off(nt, std::vector<uint64_t>(nc, 0));
    for (size_t c = 0; c < nc; ++c) { uint64_t a = 0; for (unsigned t = 0; t < nt; ++t) { off[t][c] = a; a += cnt[t][c]; } }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([&, t]() {
            uint64_t lo, hi;
            range(t, lo, hi);
            Read r;
            ppp_read p;
            std::vector<uint64_t> w = off[t];
            for (uint64_t i = lo; i < hi; ++i) {
                m.read(i, r);
                if (packRead(m, r, min_len, max_len, mode, g_lo, g_hi_plus, g_hi_minus, p)) out[r.contig][w[r.contig]++] = p;
            }
        });
    for (auto& th : pool) th.join();
    return 0;
}

}  // extern "C"
