// frontend.h -- per-record feature extraction: the host half of countReadsInBamFile
// (pingpongpro.cpp:415-484), producing the packed ppp_read upload records.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "aln_reader.h"
#include "ppp_types.h"

namespace ppp {

struct ExtractOptions {
    uint32_t minLen = 24;  // -l (:261-263)
    uint32_t maxLen = 32;  // -L (:264-266)
    int mode = PPP_MULTIHITS_WEIGHTED;
};

// Read weight, pingpongpro.cpp:430-456.
static inline float read_weight(int mode, uint32_t nh) {
    if (mode == PPP_MULTIHITS_UNIQUE) return 1.0f;
    if (mode == PPP_MULTIHITS_WEIGHTED) return (float)(1.0 / (double)nh);
    return nh == 1 ? 1.0f : 0.0f;
}

// Restates pingpongpro.cpp:415-484 for one record.  Returns true when the record contributes to a
// stack; then `out` is its packed form, `contig` its rID and `weight` the value added at :487-488.
static inline bool extract_read(const AlnView& v, const ExtractOptions& o, ppp_read& out, int32_t& contig, float& weight) {
    if (v.beginPos == 2147483647 || v.beginPos == -1) return false;  // :415 (INVALID_POS / unmapped)
    uint64_t alnLen = 0;                                             // :418-423
    for (uint32_t i = 0; i < v.nCigar; ++i) {
        char op = cigar_op_char(v.cigar[i]);
        if (op == 'M' || op == 'N' || op == 'D' || op == '=' || op == 'X') alnLen += cigar_op_count(v.cigar[i]);
    }
    if (alnLen < o.minLen || alnLen > o.maxLen) return false;        // :426
    uint32_t nh = 1;                                                 // :442-444 (tags are not read in unique mode, :432)
    if (o.mode != PPP_MULTIHITS_UNIQUE && v.hasNH) nh = v.nh;
    // NH:i:0 would weigh +inf in -m weighted (:449; the reference then converts inf to unsigned at :508, undefined behaviour):
    // such a record is malformed and is skipped.  NH beyond 2^30 - 1 does not fit the packed record: host and device use the
    // same clamped value (a weight below 1e-9 either way).  Both are documented divergences.
    if (o.mode == PPP_MULTIHITS_WEIGHTED && nh == 0) return false;
    if (nh > PPP_READ_NH_MAX) nh = PPP_READ_NH_MAX;
    weight = read_weight(o.mode, nh);
    if (weight <= 0) return false;                                   // :458
    if (v.rID < 0) return false;  // no contig: the reference would index its name store out of bounds (documented divergence)
    uint32_t meta = nh << PPP_READ_NH_SHIFT;
    if (v.flag & 0x10u) {                                            // :461
        out.key = (uint32_t)((uint64_t)(uint32_t)v.beginPos + alnLen);  // :464
        uint64_t clip = 0;                                           // :467-469
        if (v.nCigar > 1 && cigar_op_char(v.cigar[v.nCigar - 1]) == 'S') clip = cigar_op_count(v.cigar[v.nCigar - 1]);
        uint64_t idx = (uint64_t)v.lSeq - clip - 1 - 9;              // :470 (size_t arithmetic)
        if (idx < v.lSeq) { char b = v.base((uint32_t)idx); if (b == 'T' || b == 't') meta |= PPP_READ_A10; }
        meta |= PPP_READ_MINUS;
    } else {
        out.key = (uint32_t)v.beginPos;                              // :476
        uint64_t clip = 0;                                           // :479-481
        if (v.nCigar > 0 && cigar_op_char(v.cigar[0]) == 'S') clip = cigar_op_count(v.cigar[0]);
        uint64_t idx = clip + 9;                                     // :482
        if (idx < v.lSeq) { char b = v.base((uint32_t)idx); if (b == 'A' || b == 'a') meta |= PPP_READ_A10; }
    }
    out.meta = meta;
    contig = v.rID;
    return true;
}

// What the worker threads keep of a record in batch mode (AlnReader::startBatches): 16 bytes.
struct ExtractedRead {
    int32_t contig;
    ppp_read read;
    float weight;  // the value added at :487-488
};
// AlnReader::RecordFn over extract_read; `user` is the ExtractOptions.
void extract_record_fn(void* user, const AlnView& v, std::vector<uint8_t>& out);

// A fully decoded input file: reads per contig in file order.
struct DecodedFile {
    std::vector<std::string> names;
    std::vector<uint64_t> lengths;
    std::vector<std::vector<ppp_read> > reads;
    double totalWeight = 0;  // :488
    uint64_t nRecords = 0, nKept = 0;
    uint64_t nRejectedGuesses = 0;  // AlnReader::rejectedBoundaryGuesses()
};

// 0 = ok, 1 = cannot open, 2 = malformed record.
// threads <= 1: the sequential record walk; otherwise inflate / line parsing, record decoding and extraction all run on
// `threads` workers and this thread only appends the kept reads in file order.
// A contig whose length the header does not give (headerless SAM; a name the @SQ lines lack) gets the extent of its kept
// reads (largest 5'-end key + 1) as its length.
int decode_file(const char* path, const ExtractOptions& o, DecodedFile& out, int threads = 1);

// The same pass without keeping the reads: the name store after the whole file (names the header lacks appended in order
// of appearance, like SeqAn's) and per contig 1 + the largest 5'-end key of the kept reads.  The command line runs it
// over headerless SAM input before it sizes the stack arrays.  Return codes of decode_file.
int scan_extents(const char* path, const ExtractOptions& o, std::vector<std::string>& names, std::vector<uint64_t>& extents, int threads = 1);

}  // namespace ppp
