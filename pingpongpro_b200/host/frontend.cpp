// frontend.cpp -- whole-file decode into packed per-contig reads + the C API of ppp_host.h.
#include "frontend.h"

#include <fstream>
#include <map>

#include "annotation.h"
#include "fast_inflate.h"
#include "ppp_host.h"

namespace ppp {

void extract_record_fn(void* user, const AlnView& v, std::vector<uint8_t>& out) {
    ExtractedRead e;
    if (!extract_read(v, *static_cast<const ExtractOptions*>(user), e.read, e.contig, e.weight)) return;
    const uint8_t* p = reinterpret_cast<const uint8_t*>(&e);
    out.insert(out.end(), p, p + sizeof e);
}

int decode_file(const char* path, const ExtractOptions& o, DecodedFile& out, int threads) {
    ExtractOptions opts = o;  // read by the reader's worker threads: declared before the reader, so that it outlives them on every return path
    AlnReader rd;
    if (!rd.open(path)) return 1;
    out.names = rd.names();
    out.lengths = rd.lengths();
    out.reads.assign(out.names.size(), std::vector<ppp_read>());
    if (threads > 1) {
        rd.startBatches(threads, extract_record_fn, &opts);
        const uint8_t* bytes;
        size_t nBytes;
        uint64_t n;
        int err;
        while (rd.nextBatch(bytes, nBytes, n, err)) {
            out.nRecords += n;
            const ExtractedRead* e = reinterpret_cast<const ExtractedRead*>(bytes);
            for (size_t k = 0, m = nBytes / sizeof(ExtractedRead); k < m; ++k) {
                if ((size_t)e[k].contig >= out.reads.size()) {  // contig that is missing from the header
                    out.names = rd.names();
                    out.lengths = rd.lengths();
                    out.reads.resize(out.names.size());
                    if ((size_t)e[k].contig >= out.reads.size()) return 2;
                }
                out.reads[e[k].contig].push_back(e[k].read);
                out.totalWeight += (double)e[k].weight;  // :488, file order
                ++out.nKept;
            }
            if (err) return 2;
        }
        out.nRejectedGuesses = rd.rejectedBoundaryGuesses();
        return rd.good() ? 0 : 2;
    }
    AlnView v;
    while (!rd.atEnd()) {
        if (rd.next(v) != 0) return 2;
        ++out.nRecords;
        ppp_read r;
        int32_t contig;
        float w;
        if (!extract_read(v, o, r, contig, w)) continue;
        if ((size_t)contig >= out.reads.size()) {  // contig that is missing from the header
            out.names = rd.names();
            out.lengths = rd.lengths();
            out.reads.resize(out.names.size());
            if ((size_t)contig >= out.reads.size()) return 2;
        }
        out.reads[contig].push_back(r);
        out.totalWeight += (double)w;  // :488, file order
        ++out.nKept;
    }
    if (!rd.good()) return 2;
    return 0;
}

}  // namespace ppp

struct ppp_fe {
    ppp::DecodedFile f;
};

extern "C" {

int ppp_fe_decode(const char* path, uint32_t min_len, uint32_t max_len, int mode, ppp_fe** out) {
    return ppp_fe_decode_mt(path, min_len, max_len, mode, 1, out);
}

int ppp_fe_decode_mt(const char* path, uint32_t min_len, uint32_t max_len, int mode, int threads, ppp_fe** out) {
    *out = nullptr;
    ppp_fe* fe = new ppp_fe;
    ppp::ExtractOptions o;
    o.minLen = min_len;
    o.maxLen = max_len;
    o.mode = mode;
    int rc = ppp::decode_file(path, o, fe->f, threads);
    if (rc != 0) { delete fe; return rc; }
    *out = fe;
    return 0;
}
int ppp_fast_inflate(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) { return ppp::fast_inflate(in, in_len, out, out_len) ? 1 : 0; }
void ppp_fe_free(ppp_fe* fe) { delete fe; }
int ppp_fe_n_contigs(const ppp_fe* fe) { return (int)fe->f.names.size(); }
const char* ppp_fe_contig_name(const ppp_fe* fe, int c) { return fe->f.names[c].c_str(); }
uint64_t ppp_fe_contig_length(const ppp_fe* fe, int c) { return fe->f.lengths[c]; }
size_t ppp_fe_n_reads(const ppp_fe* fe, int c) { return fe->f.reads[c].size(); }
const ppp_read* ppp_fe_reads(const ppp_fe* fe, int c) { return fe->f.reads[c].data(); }
double ppp_fe_total_weight(const ppp_fe* fe) { return fe->f.totalWeight; }
uint64_t ppp_fe_n_records(const ppp_fe* fe) { return fe->f.nRecords; }
uint64_t ppp_fe_n_kept(const ppp_fe* fe) { return fe->f.nKept; }
uint64_t ppp_fe_rejected_guesses(const ppp_fe* fe) { return fe->f.nRejectedGuesses; }

struct ppp_annot_impl {
    std::vector<std::string> names;
    std::vector<ppp_annot_interval> iv;
    std::vector<std::string> ids;
};

int ppp_annot_read(const char* const* paths, int n_paths, const char* const* contig_names, int n_names, ppp_annot** out) {
    *out = nullptr;
    ppp_annot_impl* a = new ppp_annot_impl;
    for (int i = 0; i < n_names; ++i) a->names.push_back(contig_names[i]);
    ppp::TransposonsPerGenome tp;
    for (int i = 0; i < n_paths; ++i) {
        std::ifstream fs(paths[i]);
        if (fs.fail()) { delete a; return 1; }
        ppp::readTransposons(fs, ppp::formatFromPath(paths[i]), tp, a->names);
    }
    for (const auto& kv : tp)
        for (const ppp::Transposon& t : kv.second) {
            a->iv.push_back(ppp_annot_interval{kv.first, t.strand, t.start, t.end});
            a->ids.push_back(t.identifier);
        }
    *out = reinterpret_cast<ppp_annot*>(a);
    return 0;
}
int ppp_annot_predict(const uint32_t* contigs, const uint32_t* positions, size_t n, const char* const* contig_names, int n_names, uint32_t range,
                      ppp_annot** out) {
    *out = nullptr;
    ppp_annot_impl* a = new ppp_annot_impl;
    for (int i = 0; i < n_names; ++i) a->names.push_back(contig_names[i]);
    std::map<unsigned, std::vector<unsigned> > perContig;  // loci arrive in (contig, position) order
    for (size_t i = 0; i < n; ++i) {
        if (contigs[i] >= (uint32_t)n_names) { delete a; return 1; }
        perContig[contigs[i]].push_back(positions[i]);
    }
    ppp::TransposonsPerGenome tp;
    ppp::predictTransposons(perContig, a->names, range, tp);
    for (const auto& kv : tp)
        for (const ppp::Transposon& t : kv.second) {
            a->iv.push_back(ppp_annot_interval{kv.first, t.strand, t.start, t.end});
            a->ids.push_back(t.identifier);
        }
    *out = reinterpret_cast<ppp_annot*>(a);
    return 0;
}
void ppp_annot_free(ppp_annot* a) { delete reinterpret_cast<ppp_annot_impl*>(a); }
size_t ppp_annot_count(const ppp_annot* a) { return reinterpret_cast<const ppp_annot_impl*>(a)->iv.size(); }
void ppp_annot_intervals(const ppp_annot* a, ppp_annot_interval* out) {
    const ppp_annot_impl* p = reinterpret_cast<const ppp_annot_impl*>(a);
    for (size_t i = 0; i < p->iv.size(); ++i) out[i] = p->iv[i];
}
const char* ppp_annot_identifier(const ppp_annot* a, size_t i) { return reinterpret_cast<const ppp_annot_impl*>(a)->ids[i].c_str(); }
int ppp_annot_n_names(const ppp_annot* a) { return (int)reinterpret_cast<const ppp_annot_impl*>(a)->names.size(); }
const char* ppp_annot_name(const ppp_annot* a, int i) { return reinterpret_cast<const ppp_annot_impl*>(a)->names[i].c_str(); }

}  // extern "C"
