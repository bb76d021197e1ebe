// aln_reader.h -- SAM / BGZF-BAM record walker, optionally with worker threads (own code, zlib only).
//
// This is the ONE alignment decoder of the project.  It replaces the SeqAn 1.4
// BamStream the reference uses at pingpongpro.cpp:402-415 and :1482-1520 and is
// shared by
//   * the new host front end (pingpongpro_b200/host/*.cpp), and
//   * the SeqAn compatibility shim that lets the UNMODIFIED reference compile
//     as the parity oracle (oracle/seqan_shim/seqan/bam_io.h),
// so a decode difference can never masquerade as a kernel bug.
//
// Decode semantics kept: SAM POS -> beginPos = POS-1
// (POS 0 -> -1 = unmapped); rID = index of RNAME in @SQ order ('*' -> -1);
// SEQ text kept with its case for SAM, upper-case letters for BAM 4-bit codes;
// NH is read when its tag type is an integer type (SAM 'i'; BAM c C s S i I).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

namespace ppp {

// One CIGAR element in BAM encoding: (count << 4) | opcode, opcode indexes "MIDNSHP=X".
static inline char cigar_op_char(uint32_t e) { return "MIDNSHP=X???????"[e & 15u]; }
static inline uint32_t cigar_op_count(uint32_t e) { return e >> 4; }

// A decoded record.  Pointers stay valid until the next call of AlnReader::next().
struct AlnView {
    int32_t rID = -1;
    int32_t beginPos = -1;       // 0-based; -1 when unmapped (SAM POS 0 / BAM pos -1)
    uint32_t flag = 0;
    const uint32_t* cigar = nullptr;  // BAM-encoded elements
    uint32_t nCigar = 0;
    uint32_t lSeq = 0;           // number of bases (0 when SEQ is '*')
    const char* seqText = nullptr;    // SAM: lSeq chars.  nullptr for BAM.
    const uint8_t* seq4 = nullptr;    // BAM: 4-bit packed.  nullptr for SAM.
    bool hasNH = false;          // an integer-typed NH tag exists
    uint32_t nh = 1;             // its value (as unsigned, like extractTagValue(unsigned&))

    inline char base(uint32_t i) const {
        if (seqText) return seqText[i];
        return "=ACMGRSVTWYHKDBN"[(seq4[i >> 1] >> ((~i & 1u) << 2)) & 15u];
    }
};

// SAM: a small direct-mapped cache in front of the RNAME -> rID hash map (one per parsing thread).  Sorted input repeats one
// name for millions of lines, unsorted input cycles through a handful; either way the std::string construction and the
// hashing of the general map (60 ns per line) are avoided.  Keyed by the last (up to) eight bytes of the name and its length.
struct SamNameCache {
    struct Entry {
        uint64_t tail = 0;
        uint32_t len = 0;
        int32_t id = -1;
        std::string name;  // compared in full only for names longer than eight bytes
    };
    static const int kSlots = 64;
    Entry slot[kSlots];
    void clear() { for (Entry& e : slot) e = Entry(); }
};

class AlnReader {
public:
    AlnReader();
    ~AlnReader();
    AlnReader(const AlnReader&) = delete;
    AlnReader& operator=(const AlnReader&) = delete;

    // Opens a file; format is detected by the gzip/BGZF magic (else SAM text).
    // Reads the header.  Returns false when the file cannot be opened or the
    // header is malformed.
    bool open(const char* path);
    void close();
    bool good() const { return good_; }
    bool isBam() const { return isBam_; }

    // Contig names / lengths in @SQ (SAM) or reference-table (BAM) order.
    const std::vector<std::string>& names() const { return names_; }
    const std::vector<uint64_t>& lengths() const { return lengths_; }

    // True when no further record exists.
    bool atEnd();
    // Decodes the next record.  Returns 0 on success, non-zero on a malformed record
    // ("Failed to read record", pingpongpro.cpp:411).
    int next(AlnView& out);

    // Records decoded so far.
    uint64_t recordCount() const { return nRecords_; }
    // Batch mode, BAM: runs of blocks whose guessed first record boundary the in-order check rejected (they were decoded
    // again from the true boundary).  0 on every real file seen so far; the tests provoke it.
    uint64_t rejectedBoundaryGuesses() const;

    // BAM only: inflate BGZF blocks on `n` worker threads (read-ahead, in order) while next() keeps walking the records
    // on the calling thread.  Call after open(); n <= 1 keeps the sequential path.
    void setThreads(int n);

    // Batch mode: the per-record work runs on the worker threads as well.  `fn` is called for every record (on any
    // thread, concurrently; it appends what it wants to keep to `out`), and nextBatch() hands the bytes produced for
    // consecutive records back in file order.  BAM: the worker that inflates a run of blocks also decodes the records
    // inside it, from a record boundary it guesses; the calling thread, which sees the runs in order, verifies every
    // guess against the true boundary and decodes the few records that straddle two runs.  SAM: the workers parse whole
    // lines.  Use either atEnd()/next() or startBatches()/nextBatch() on one file, not both.
    typedef void (*RecordFn)(void* user, const AlnView& v, std::vector<uint8_t>& out);
    void startBatches(int threads, RecordFn fn, void* user);
    // False at the end of the input.  out[0, outLen) stays valid until the next call.  error != 0: a malformed record
    // follows the records of this batch ("Failed to read record"); the call after that returns false.
    bool nextBatch(const uint8_t*& out, size_t& outLen, uint64_t& nRecords, int& error);

private:
    bool fill();                  // refill raw buffer (decompressing BGZF blocks for BAM)
    bool ensure(size_t n);        // make >= n bytes available at cur_ (BAM)
    bool readBgzfBlock();
    bool parseSamHeaderLine(const char* line, size_t len);
    int nextSam(AlnView& out);
    int nextBam(AlnView& out);
    bool getLine(const char*& line, size_t& len);
    int32_t internExtraName(const std::string& name);

    FILE* fp_ = nullptr;
    bool good_ = false;
    bool isBam_ = false;
    bool eofRaw_ = false;
    std::vector<std::string> names_;
    std::vector<uint64_t> lengths_;
    // uncompressed byte window [cur_, end_) inside buf_
    std::vector<uint8_t> buf_;
    size_t cur_ = 0, end_ = 0;
    std::vector<uint8_t> zin_;    // compressed block scratch
    std::vector<uint32_t> cigarScratch_;
    uint64_t nRecords_ = 0;
    // name -> rID lookup for SAM
    struct NameMap;
    NameMap* nameMap_ = nullptr;
    struct Pipeline;
    Pipeline* pipe_ = nullptr;    // parallel BGZF inflate (setThreads)
    struct SamPipeline;
    SamPipeline* samPipe_ = nullptr;  // parallel SAM line parsing (startBatches)
    struct BamBatcher;
    BamBatcher* bamBatch_ = nullptr;  // in-order verification of the workers' record decoding (startBatches)
    RecordFn batchFn_ = nullptr;
    void* batchUser_ = nullptr;
    bool batchFailed_ = false;
    std::unordered_map<std::string, int32_t> extraNames_;  // SAM contigs that are missing from the header, in order of appearance
    SamNameCache nameCache_;      // SAM: RNAME -> rID of the records seen lately
};

}  // namespace ppp
