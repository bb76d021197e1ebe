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

namespace {

// One pass over a file.  `keep`: the reads are stored per contig (decode_file); otherwise only names and extents are
// collected (scan_extents).  extents[c] = 1 + the largest 5'-end key among the kept reads of contig c.
int walk_file(const char* path, const ExtractOptions& o, DecodedFile& out, std::vector<uint64_t>& extents, int threads, bool keep) {
    ExtractOptions opts = o;  // read by the reader's worker threads: declared before the reader, so that it outlives them on every return path
    AlnReader rd;
    if (!rd.open(path)) return 1;
    out.names = rd.names();
    out.lengths = rd.lengths();
    out.reads.assign(out.names.size(), std::vector<ppp_read>());
    extents.assign(out.names.size(), 0);
    auto take = [&](int32_t contig, const ppp_read& r, float w) -> bool {
        if ((size_t)contig >= out.reads.size()) {  // contig that is missing from the header: the name store grew
            out.names = rd.names();
            out.lengths = rd.lengths();
            out.reads.resize(out.names.size());
            extents.resize(out.names.size(), 0);
            if ((size_t)contig >= out.reads.size()) return false;
        }
        if (keep) out.reads[contig].push_back(r);
        if ((uint64_t)r.key + 1 > extents[contig]) extents[contig] = (uint64_t)r.key + 1;
        out.totalWeight += (double)w;  // :488, file order
        ++out.nKept;
        return true;
    };
    int rc = 0;
    if (threads > 1) {
        rd.startBatches(threads, extract_record_fn, &opts);
        const uint8_t* bytes;
        size_t nBytes;
        uint64_t n;
        int err;
        while (rd.nextBatch(bytes, nBytes, n, err)) {
            out.nRecords += n;
            const ExtractedRead* e = reinterpret_cast<const ExtractedRead*>(bytes);
            for (size_t k = 0, m = nBytes / sizeof(ExtractedRead); k < m; ++k)
                if (!take(e[k].contig, e[k].read, e[k].weight)) return 2;
            if (err) return 2;
        }
        out.nRejectedGuesses = rd.rejectedBoundaryGuesses();
        rc = rd.good() ? 0 : 2;
    } else {
        AlnView v;
        while (!rd.atEnd()) {
            if (rd.next(v) != 0) return 2;
            ++out.nRecords;
            ppp_read r;
            int32_t contig;
            float w;
            if (!extract_read(v, o, r, contig, w)) continue;
            if (!take(contig, r, w)) return 2;
        }
        if (!rd.good()) rc = 2;
    }
    // names that only filtered records carried (the name store grows for every record, like SeqAn's)
    out.names = rd.names();
    out.lengths = rd.lengths();
    out.reads.resize(out.names.size());
    extents.resize(out.names.size(), 0);
    return rc;
}

}  // namespace

int decode_file(const char* path, const ExtractOptions& o, DecodedFile& out, int threads) {
    std::vector<uint64_t> extents;
    const int rc = walk_file(path, o, out, extents, threads, true);
    // a contig without an @SQ length (headerless SAM, or a name the header lacks): the reference works without lengths
    // (std::map per contig); the dense stack arrays get the extent of the reads instead
    for (size_t c = 0; c < out.lengths.size() && c < extents.size(); ++c)
        if (out.lengths[c] == 0) out.lengths[c] = extents[c];
    return rc;
}

int scan_extents(const char* path, const ExtractOptions& o, std::vector<std::string>& names, std::vector<uint64_t>& extents, int threads) {
    DecodedFile f;
    const int rc = walk_file(path, o, f, extents, threads, false);
    names = f.names;
    for (size_t c = 0; c < extents.size() && c < f.lengths.size(); ++c)
        if (f.lengths[c] > extents[c]) extents[c] = f.lengths[c];  // (a length the header does give wins)
    return rc;
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
