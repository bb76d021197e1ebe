// aln_reader.cpp -- see aln_reader.h.  Own implementation over the system zlib.
#include "aln_reader.h"

#include "fast_inflate.h"

#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <sys/mman.h>
#include <sys/stat.h>
#include <zlib.h>

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>

namespace ppp {

struct AlnReader::NameMap {
    std::unordered_map<std::string, int32_t> m;
};

static const size_t kRawChunk = 4u << 20;

// ---------------------------------------------------------------------------
// BAM record decoding (shared by the sequential walk and the batch workers)
// ---------------------------------------------------------------------------
static bool skipBamTag(const uint8_t* p, size_t n, size_t& used, char type);

static size_t bamTypeSize(char t) {
    switch (t) {
        case 'A': case 'c': case 'C': return 1;
        case 's': case 'S': return 2;
        case 'i': case 'I': case 'f': return 4;
        default: return 0;
    }
}

static bool skipBamTag(const uint8_t* p, size_t n, size_t& used, char type) {
    size_t sz = bamTypeSize(type);
    if (sz) { if (n < sz) return false; used = sz; return true; }
    if (type == 'Z' || type == 'H') {
        const void* z = memchr(p, 0, n);
        if (!z) return false;
        used = (size_t)((const uint8_t*)z - p) + 1;
        return true;
    }
    if (type == 'B') {
        if (n < 5) return false;
        size_t es = bamTypeSize((char)p[0]);
        uint32_t cnt;
        memcpy(&cnt, p + 1, 4);
        if (!es || n < 5 + es * (size_t)cnt) return false;
        used = 5 + es * (size_t)cnt;
        return true;
    }
    return false;
}

// `p` points behind the block_size word of one record of `blockSize` bytes.  False: malformed.
static bool parseBamRecord(const uint8_t* p, uint32_t blockSize, std::vector<uint32_t>& cigarScratch, AlnView& out) {
    if (blockSize < 32) return false;
    int32_t refID, pos;
    memcpy(&refID, p, 4);
    memcpy(&pos, p + 4, 4);
    uint32_t lReadName = p[8];
    uint16_t nCigar, flag;
    memcpy(&nCigar, p + 12, 2);
    memcpy(&flag, p + 14, 2);
    uint32_t lSeq;
    memcpy(&lSeq, p + 16, 4);
    size_t off = 32 + lReadName;
    size_t need = off + 4 * (size_t)nCigar + ((size_t)lSeq + 1) / 2 + lSeq;
    if (need > blockSize) return false;
    out.rID = refID;
    out.beginPos = pos;
    out.flag = flag;
    cigarScratch.resize(nCigar);
    if (nCigar) memcpy(cigarScratch.data(), p + off, 4 * (size_t)nCigar);  // unaligned source
    out.cigar = cigarScratch.data();
    out.nCigar = nCigar;
    off += 4 * (size_t)nCigar;
    out.seqText = nullptr;
    out.seq4 = p + off;
    out.lSeq = lSeq;
    off += ((size_t)lSeq + 1) / 2 + lSeq;
    // tags
    out.hasNH = false;
    out.nh = 1;
    const uint8_t* t = p + off;
    size_t n = blockSize - off;
    while (n >= 3) {
        char type = (char)t[2];
        size_t used = 0;
        if (!skipBamTag(t + 3, n - 3, used, type)) break;
        if (t[0] == 'N' && t[1] == 'H') {
            const uint8_t* v = t + 3;
            switch (type) {
                case 'c': out.hasNH = true; out.nh = (uint32_t)(int32_t)(int8_t)v[0]; break;
                case 'C': out.hasNH = true; out.nh = v[0]; break;
                case 's': { int16_t x; memcpy(&x, v, 2); out.hasNH = true; out.nh = (uint32_t)(int32_t)x; break; }
                case 'S': { uint16_t x; memcpy(&x, v, 2); out.hasNH = true; out.nh = x; break; }
                case 'i': { int32_t x; memcpy(&x, v, 4); out.hasNH = true; out.nh = (uint32_t)x; break; }
                case 'I': { uint32_t x; memcpy(&x, v, 4); out.hasNH = true; out.nh = x; break; }
                default: break;
            }
            break;
        }
        t += 3 + used;
        n -= 3 + used;
    }
    return true;
}

// ---------------------------------------------------------------------------
// parallel BGZF inflate: one reader thread cuts the file into jobs of whole blocks, a pool inflates them, the
// consumer takes the jobs back in file order.  BGZF blocks are independent deflate streams, so this is exact.
// ---------------------------------------------------------------------------
struct AlnReader::Pipeline {
    struct Block { size_t off, clen; uint32_t isize; size_t outOff; };
    struct Job {
        // plain arrays, not vectors: resizing a vector zero-fills it, and clearing ~5 bytes of output per byte of input
        // on the ONE reader thread was the slowest stage of the whole BAM path
        std::unique_ptr<uint8_t[]> raw, out;
        size_t rawCap = 0, outCap = 0, outLen = 0;
        std::vector<Block> blocks;
        bool done = false, bad = false;
        // batch mode: what the inflating worker decoded itself -- the records in out[specStart, tailOff), their kept
        // bytes in recOut[kHeadroom, ...)
        std::vector<uint8_t> recOut;
        uint64_t nParsed = 0;
        size_t specStart = 0, tailOff = 0;
        bool parsed = false;    // a start was guessed and the records behind it were decoded
        bool parseBad = false;  // ... up to a malformed one at tailOff
        void needRaw(size_t n) { if (rawCap < n) { raw.reset(new uint8_t[n]); rawCap = n; } }
        void needOut(size_t n) { if (outCap < n) { out.reset(new uint8_t[n]); outCap = n; } }
    };
    FILE* fp;
    // batch mode (AlnReader::startBatches): the worker that inflated a job also decodes its records, block by block,
    // while they are in its cache -- see findRecordStart and AlnReader::BamBatcher
    static constexpr size_t kHeadroom = 1024;  // bytes left free in front of Job::recOut for the consumer's stitched records
    AlnReader::RecordFn fn = nullptr;
    void* user = nullptr;
    int32_t nRef = 0;
    bool noGuess = false;  // test knob (PPP_BAM_NO_GUESS): every job goes through the consumer's own decoding
    std::mutex mu;
    std::condition_variable cvWork, cvDone, cvSpace;
    std::deque<std::shared_ptr<Job> > todo, ordered;  // ordered: every job in file order until consumed
    bool eof = false, stop = false, readError = false;
    size_t maxOutstanding;
    std::thread reader;
    std::vector<std::thread> workers;
    std::shared_ptr<Job> current;  // keeps the chunk handed to the consumer alive
    // finished jobs are reused (their buffers keep their capacity): fresh multi-megabyte allocations mean mmap / page
    // faults / munmap, which serialise the worker threads on the address space
    std::vector<std::shared_ptr<Job> > pool;
    std::shared_ptr<Job> obtain() {
        std::lock_guard<std::mutex> l(mu);
        if (pool.empty()) return std::make_shared<Job>();
        std::shared_ptr<Job> j = pool.back();
        pool.pop_back();
        return j;
    }
    void recycle(std::shared_ptr<Job>& j) {
        if (!j) return;
        j->outLen = 0;
        j->blocks.clear();
        j->done = j->bad = false;
        j->recOut.clear();
        j->nParsed = 0;
        j->specStart = j->tailOff = 0;
        j->parsed = j->parseBad = false;
        std::lock_guard<std::mutex> l(mu);
        if (pool.size() < 256) pool.push_back(std::move(j));
        j.reset();
    }

    Pipeline(FILE* f, int threads, AlnReader::RecordFn fn_ = nullptr, void* user_ = nullptr, int32_t nRef_ = 0)
        : fp(f), fn(fn_), user(user_), nRef(nRef_), noGuess(getenv("PPP_BAM_NO_GUESS") != nullptr), maxOutstanding((size_t)threads * 3 + 2) {
        reader = std::thread([this] { readLoop(); });
        for (int i = 0; i < threads; ++i) workers.emplace_back([this] { workLoop(); });
    }
    ~Pipeline() {
        {
            std::lock_guard<std::mutex> l(mu);
            stop = true;
        }
        cvWork.notify_all();
        cvSpace.notify_all();
        cvDone.notify_all();
        reader.join();
        for (auto& w : workers) w.join();
    }
    // One large read per job (straight into the job's buffer), block headers parsed in place; a block that the read cut
    // in two is carried over to the front of the next job.
    void readLoop() {
        const size_t kJobBytes = 1u << 20;
        std::vector<uint8_t> carry;
        for (;;) {
            std::shared_ptr<Job> job = obtain();
            const size_t have = carry.size();
            job->needRaw(have + kJobBytes + 8);  // (+8: the decoder's bit reader may look a few bytes past a block)
            if (have) memcpy(job->raw.get(), carry.data(), have);
            const size_t got = fread(job->raw.get() + have, 1, kJobBytes, fp);
            const size_t total = have + got;
            const uint8_t* r = job->raw.get();
            size_t off = 0, outTotal = 0;
            bool bad = false;
            while (total - off >= 18) {
                if (r[off] != 31 || r[off + 1] != 139 || r[off + 2] != 8 || !(r[off + 3] & 4)) { bad = true; break; }
                const uint32_t xlen = r[off + 10] | (r[off + 11] << 8);
                if (total - off < 12 + (size_t)xlen) break;  // header incomplete
                int bsize = -1;
                for (uint32_t i = 0; i + 4 <= xlen;) {
                    const uint8_t* x = r + off + 12 + i;
                    const uint32_t slen = x[2] | (x[3] << 8);
                    if (x[0] == 'B' && x[1] == 'C' && slen == 2 && i + 6 <= xlen) bsize = x[4] | (x[5] << 8);
                    i += 4 + slen;
                }
                if (bsize < 0 || (size_t)bsize + 1 < 12 + (size_t)xlen + 8) { bad = true; break; }
                const size_t blockLen = (size_t)bsize + 1;
                if (total - off < blockLen) break;  // block incomplete
                uint32_t isize;
                memcpy(&isize, r + off + blockLen - 4, 4);
                job->blocks.push_back(Block{off + 12 + xlen, blockLen - xlen - 12 - 8, isize, outTotal});
                outTotal += isize;
                off += blockLen;
            }
            const bool fileEnd = got == 0;
            if (!bad) {
                carry.assign(r + off, r + total);
                if (fileEnd && !carry.empty()) bad = true;  // the file ends inside a block
                if (!fileEnd && off == 0 && total >= kJobBytes + 65536 + 18) bad = true;  // no block fits: not BGZF
            }
            job->needOut(outTotal + 16);
            job->outLen = outTotal;
            job->bad = bad;
            {
                std::unique_lock<std::mutex> l(mu);
                cvSpace.wait(l, [this] { return stop || ordered.size() < maxOutstanding; });
                if (stop) return;
                if (!job->blocks.empty() || bad) {
                    ordered.push_back(job);
                    if (bad) { job->done = true; cvDone.notify_all(); }
                    else { todo.push_back(job); cvWork.notify_one(); }
                }
                if (fileEnd || bad) { eof = true; cvDone.notify_all(); cvWork.notify_all(); return; }
            }
        }
    }
    // ---- batch mode: record decoding inside the inflate worker ----
    struct RecordWalk {
        enum State { kSearching, kWalking, kGaveUp } state = kSearching;
        size_t start = 0, pos = 0;
        std::vector<uint32_t> cigarScratch;
        AlnView view;
    };
    // Does a BAM record that is consistent with itself and with the header start at data[q]?  (block_size, refID,
    // next_refID within the reference table, the variable-length fields inside the block, a NUL behind the read name.)
    // `known` = bytes available.  1 = yes, 0 = no, -1 = the fixed part is not fully available.
    int plausibleRecord(const uint8_t* data, size_t q, size_t known) const {
        if (known - q < 36) return -1;
        uint32_t bs, lSeq;
        int32_t refID, pos, nextRef;
        uint16_t nCigar;
        memcpy(&bs, data + q, 4);
        memcpy(&refID, data + q + 4, 4);
        memcpy(&pos, data + q + 8, 4);
        const uint32_t lName = data[q + 12];
        memcpy(&nCigar, data + q + 16, 2);
        memcpy(&lSeq, data + q + 20, 4);
        memcpy(&nextRef, data + q + 24, 4);
        if (bs < 32 || bs > (1u << 24) || refID < -1 || refID >= nRef || nextRef < -1 || nextRef >= nRef || pos < -1 || lName == 0 || lSeq > (1u << 24)) return 0;
        if (32 + (size_t)lName + 4 * (size_t)nCigar + ((size_t)lSeq + 1) / 2 + lSeq > bs) return 0;
        const size_t nul = q + 4 + 32 + lName - 1;
        if (nul < known && data[nul] != 0) return 0;
        return 1;
    }
    // A guess at the first record boundary in a job's output: the first offset (among the first 64 KB) from which eight
    // consecutive plausible records follow, or as many as the known bytes hold (at least three, unless the job ends).
    // >= 0: offset; -1: more bytes are needed to decide; -2: none.
    // The guess is only ever USED after the in-order consumer has verified it against the true boundary.
    long findRecordStart(const uint8_t* data, size_t known, size_t total) const {
        const size_t limit = std::min<size_t>(known, 65536);
        for (size_t p = 0; p < limit; ++p) {
            size_t q = p;
            int count = 0;
            for (;;) {
                int r = plausibleRecord(data, q, known);
                if (r == 0) { count = -1; break; }
                if (r < 0) break;  // ran out of known bytes
                uint32_t bs;
                memcpy(&bs, data + q, 4);
                q += 4 + (size_t)bs;
                if (++count == 8) break;
                if (q >= known) break;
            }
            if (count < 0) continue;
            if (count >= 8 || count >= 3 || (known == total && count >= 1)) return (long)p;
            if (known < total) return -1;  // plausible so far, too short to tell
        }
        if (known < 65536 && known < total) return -1;
        return -2;
    }
    // Decodes the records that are complete within out[0, known).
    void decodeAvailable(Job& job, RecordWalk& w, size_t known) {
        const uint8_t* data = job.out.get();
        if (w.state == RecordWalk::kSearching) {
            long s = findRecordStart(data, known, job.outLen);
            if (s == -1) return;
            if (s < 0) { w.state = RecordWalk::kGaveUp; return; }
            w.state = RecordWalk::kWalking;
            w.start = w.pos = (size_t)s;
        }
        if (w.state != RecordWalk::kWalking || job.parseBad) return;
        while (known - w.pos >= 4) {
            uint32_t bs;
            memcpy(&bs, data + w.pos, 4);
            if (known - w.pos - 4 < (size_t)bs) break;
            if (!parseBamRecord(data + w.pos + 4, bs, w.cigarScratch, w.view)) { job.parseBad = true; break; }
            fn(user, w.view, job.recOut);
            ++job.nParsed;
            w.pos += 4 + (size_t)bs;
        }
    }
    // Next job in file order (inflated, and in batch mode decoded); null at the end of the file or after requestStop.
    // The caller hands it back with recycle().
    std::shared_ptr<Job> nextJob() {
        std::unique_lock<std::mutex> l(mu);
        cvDone.wait(l, [this] { return stop || (!ordered.empty() && ordered.front()->done) || (ordered.empty() && eof); });
        if (stop || ordered.empty()) return nullptr;
        std::shared_ptr<Job> j = ordered.front();
        ordered.pop_front();
        cvSpace.notify_one();
        return j;
    }
    void workLoop() {
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> l(mu);
                cvWork.wait(l, [this] { return stop || !todo.empty() || eof; });
                if (stop) return;
                if (todo.empty()) { if (eof) return; continue; }
                job = todo.front();
                todo.pop_front();
            }
            bool bad = false;
            z_stream zs;
            bool zsLive = false;
            RecordWalk walk;
            if (fn && !noGuess) job->recOut.resize(kHeadroom);
            for (const Block& b : job->blocks) {
                if (bad) break;
                if (b.isize == 0) continue;
                // own decoder first; zlib takes over (and judges the block) whenever that one declines
                if (!fast_inflate(job->raw.get() + b.off, b.clen, job->out.get() + b.outOff, b.isize)) {
                    if (!zsLive) {
                        memset(&zs, 0, sizeof zs);
                        if (inflateInit2(&zs, -15) != Z_OK) { bad = true; break; }
                        zsLive = true;
                    }
                    inflateReset(&zs);
                    zs.next_in = job->raw.get() + b.off;
                    zs.avail_in = (uInt)b.clen;
                    zs.next_out = job->out.get() + b.outOff;
                    zs.avail_out = b.isize;
                    int rc = inflate(&zs, Z_FINISH);
                    if (rc != Z_STREAM_END || zs.avail_out != 0) { bad = true; break; }
                }
                if (fn && !noGuess) decodeAvailable(*job, walk, b.outOff + b.isize);  // the block just written is still in this core's cache
            }
            if (zsLive) inflateEnd(&zs);
            if (fn && !noGuess && !bad && walk.state == RecordWalk::kWalking) {
                job->parsed = true;
                job->specStart = walk.start;
                job->tailOff = walk.pos;
            }
            {
                std::lock_guard<std::mutex> l(mu);
                job->bad = job->bad || bad;
                job->done = true;
            }
            cvDone.notify_all();
        }
    }
    // Next inflated chunk in file order; false at end of file (or after requestStop).  `bad` is set on a malformed
    // block.  `keep` (optional) receives shared ownership of the chunk; otherwise it lives until the next call.
    bool next(const uint8_t*& data, size_t& len, bool& bad, std::shared_ptr<Job>* keep = nullptr) {
        std::unique_lock<std::mutex> l(mu);
        cvDone.wait(l, [this] { return stop || (!ordered.empty() && ordered.front()->done) || (ordered.empty() && eof); });
        if (stop || ordered.empty()) return false;
        if (current) {  // record mode: the chunk handed out last time is no longer referenced
            current->outLen = 0;
            current->blocks.clear();
            current->done = current->bad = false;
            if (pool.size() < 256) pool.push_back(current);
        }
        current = ordered.front();
        ordered.pop_front();
        cvSpace.notify_one();
        bad = current->bad;
        data = current->out.get();
        len = current->outLen;
        if (keep) { *keep = current; current.reset(); }
        return true;
    }
    void requestStop() {
        {
            std::lock_guard<std::mutex> l(mu);
            stop = true;
        }
        cvWork.notify_all();
        cvSpace.notify_all();
        cvDone.notify_all();
    }
};

// ---------------------------------------------------------------------------
// SAM line parser and its worker-thread pipeline
// ---------------------------------------------------------------------------
// Offsets of the tabs of s[0, n) in pos[0, cap); returns how many were stored (cap: there may be more).
static const int kMaxTabs = 48;
static inline int findTabs(const char* s, size_t n, uint32_t* pos, int cap) {
    int k = 0;
    size_t i = 0;
#if defined(__SSE2__)
    const __m128i tab = _mm_set1_epi8('\t');
    for (; i + 16 <= n; i += 16) {
        unsigned m = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i)), tab));
        while (m) {
            if (k == cap) return k;
            pos[k++] = (uint32_t)i + (uint32_t)__builtin_ctz(m);
            m &= m - 1;
        }
    }
#endif
    for (; i < n; ++i)
        if (s[i] == '\t') {
            if (k == cap) return k;
            pos[k++] = (uint32_t)i;
        }
    return k;
}

static inline bool parseU32(const char* s, size_t n, uint32_t& v) {
    if (n == 0) return false;
    uint64_t x = 0;
    for (size_t i = 0; i < n; ++i) {
        if (s[i] < '0' || s[i] > '9') return false;
        x = x * 10 + (uint64_t)(s[i] - '0');
        if (x > 0xffffffffull) return false;
    }
    v = (uint32_t)x;
    return true;
}

static inline int cigarOpCode(char c) {
    switch (c) {
        case 'M': return 0; case 'I': return 1; case 'D': return 2; case 'N': return 3; case 'S': return 4;
        case 'H': return 5; case 'P': return 6; case '=': return 7; case 'X': return 8; default: return -1;
    }
}

// One alignment line (without its terminator).  `names` is only read.  Returns 0 = ok, 1 = malformed, 2 = RNAME is not in
// `names` (everything else is filled in; the caller decides whether the name store grows).  CIGAR elements are appended
// to `cigars` from out.cigarOff on; SEQ is referenced by its offset in the line.
struct SamFields {
    int32_t rID, beginPos;
    uint32_t flag, cigarOff, nCigar, seqOff, lSeq, nh, nameOff, nameLen;
    bool hasNH;
};
static int parseSamLine(const char* line, size_t len, const std::unordered_map<std::string, int32_t>& names, SamNameCache& cache,
                        std::vector<uint32_t>& cigars, SamFields& out) {
    // every tab of the line in one pass (a memchr call per field costs more than the whole rest of the parser)
    uint32_t tabs[kMaxTabs];
    const int nTabs = findTabs(line, len, tabs, kMaxTabs);
    if (nTabs < 10) return 1;  // fewer than the 11 mandatory fields
    const char* f[11];
    size_t fl[11];
    for (int k = 0; k < 11; ++k) {
        const size_t b = k ? (size_t)tabs[k - 1] + 1 : 0, e = k < nTabs ? (size_t)tabs[k] : len;
        f[k] = line + b;
        fl[k] = e - b;
    }
    uint32_t flag, pos;
    if (!parseU32(f[1], fl[1], flag) || !parseU32(f[3], fl[3], pos)) return 1;
    out.flag = flag;
    out.beginPos = (int32_t)pos - 1;
    out.nameOff = (uint32_t)(f[2] - line);
    out.nameLen = (uint32_t)fl[2];
    int rc = 0;
    if (fl[2] == 1 && f[2][0] == '*') {
        out.rID = -1;
    } else {
        // the last (up to) eight bytes of the name; RNAME is the third field, so bytes in front of it exist -- eight of
        // them unless the line is degenerate, which takes the general map
        const size_t nameEnd = (size_t)(f[2] - line) + fl[2];
        SamNameCache::Entry* slot = nullptr;
        uint64_t tail = 0;
        if (nameEnd >= 8 && fl[2] != 0) {
            memcpy(&tail, line + nameEnd - 8, 8);
            if (fl[2] < 8) tail >>= 8 * (8 - fl[2]);  // (little endian: the name's bytes are the high ones)
            slot = &cache.slot[((tail ^ fl[2]) * 0x9E3779B97F4A7C15ull) >> 58];
        }
        if (slot && slot->id >= 0 && slot->tail == tail && slot->len == fl[2] && (fl[2] <= 8 || memcmp(slot->name.data(), f[2], fl[2]) == 0)) {
            out.rID = slot->id;
        } else {
            auto it = names.find(std::string(f[2], fl[2]));
            if (it == names.end()) { out.rID = -2; rc = 2; }
            else {
                out.rID = it->second;
                if (slot) { slot->tail = tail; slot->len = (uint32_t)fl[2]; slot->id = it->second; slot->name.assign(f[2], fl[2]); }
            }
        }
    }
    // CIGAR
    out.cigarOff = (uint32_t)cigars.size();
    if (!(fl[5] == 1 && f[5][0] == '*')) {
        uint64_t cnt = 0;
        bool haveDigit = false;
        for (size_t k = 0; k < fl[5]; ++k) {
            char c = f[5][k];
            if (c >= '0' && c <= '9') { cnt = cnt * 10 + (uint64_t)(c - '0'); haveDigit = true; if (cnt >= (1ull << 28)) return 1; }
            else {
                int op = cigarOpCode(c);
                if (op < 0 || !haveDigit) return 1;
                cigars.push_back((uint32_t)(cnt << 4) | (uint32_t)op);
                cnt = 0;
                haveDigit = false;
            }
        }
        if (haveDigit) return 1;
    }
    out.nCigar = (uint32_t)cigars.size() - out.cigarOff;
    // SEQ
    out.seqOff = (uint32_t)(f[9] - line);
    out.lSeq = (fl[9] == 1 && f[9][0] == '*') ? 0u : (uint32_t)fl[9];
    // optional fields: the first one that reads NH:i:<something> decides (:440-444)
    out.hasNH = false;
    out.nh = 1;
    auto tryField = [&](size_t b, size_t e) {  // true: this was the NH field
        if (e - b < 6 || memcmp(line + b, "NH:i:", 5) != 0) return false;
        const char* sv = line + b + 5;
        size_t n = e - b - 5;
        bool neg = sv[0] == '-';
        uint32_t v;
        if (parseU32(sv + (neg ? 1 : 0), n - (neg ? 1 : 0), v)) {
            out.hasNH = true;
            out.nh = neg ? (uint32_t)(-(int64_t)v) : v;
        }
        return true;
    };
    int k = 11;
    for (; k <= nTabs; ++k) {  // field k lies between tab k-1 and tab k (or the end of the line)
        const size_t b = (size_t)tabs[k - 1] + 1, e = k < nTabs ? (size_t)tabs[k] : len;
        if (b >= len || tryField(b, e)) return rc;
    }
    if (nTabs == kMaxTabs) {  // more tabs than the table holds: the rest the slow way
        size_t i = (size_t)tabs[kMaxTabs - 1] + 1;
        while (i < len) {
            const char* t = (const char*)memchr(line + i, '\t', len - i);
            size_t j = t ? (size_t)(t - line) : len;
            if (tryField(i, j)) break;
            i = j + 1;
        }
    }
    return rc;
}

// ---------------------------------------------------------------------------
// SAM batch mode: one reader thread cuts the text into jobs of whole lines, a pool parses them (name lookups against
// the header's contig table, which no longer changes) and applies the record function, the consumer takes the jobs'
// outputs back in file order.  A job with an RNAME that is not in the header is redone by the consumer, which
// extends the name store exactly as the sequential reader does.
// ---------------------------------------------------------------------------
static inline void fillView(AlnView& out, const SamFields& f, const char* line, const uint32_t* cigars) {
    out.rID = f.rID;
    out.beginPos = f.beginPos;
    out.flag = f.flag;
    out.cigar = cigars + f.cigarOff;
    out.nCigar = f.nCigar;
    out.seq4 = nullptr;
    out.seqText = line + f.seqOff;
    out.lSeq = f.lSeq;
    out.hasNH = f.hasNH;
    out.nh = f.nh;
}

struct AlnReader::SamPipeline {
    typedef AlnReader::RecordFn RecordFn;
    struct Rec { SamFields f; uint32_t lineOff, lineLen; };
    struct Job {
        std::unique_ptr<char[]> text;  // (not a vector: no zero fill of the read buffer)
        const char* data = nullptr;    // the job's lines: text.get(), or a window of the memory-mapped file
        size_t textLen = 0, textCap = 0;
        std::vector<Rec> recs;
        std::vector<uint32_t> cigars;
        std::vector<uint8_t> out;  // batch mode: what the record function kept
        bool bad = false;          // the line after recs.back() is malformed
        bool unknownName = false;  // batch mode: a contig that is missing from the header -- the consumer redoes this job
        bool done = false;
    };
    FILE* fp;
    RecordFn fn = nullptr;
    void* user = nullptr;
    const std::unordered_map<std::string, int32_t>& names;
    std::vector<char> carry;  // bytes already read by the header parser, then the partial last line of each block
    // A regular file is memory-mapped and the jobs are windows of the mapping: the one serial stage then touches a
    // page per job instead of copying the whole file through fread (2.1 GB of SAM text = 0.5 s on one core, more than
    // the parsing takes on sixteen); the page faults of a cached file spread over the workers.
    const char* map = nullptr;
    size_t mapLen = 0, mapPos = 0;
    std::mutex mu;
    std::condition_variable cvWork, cvDone, cvSpace;
    std::deque<std::shared_ptr<Job> > todo, ordered;
    bool eof = false, stop = false;
    size_t maxOutstanding;
    std::thread reader;
    std::vector<std::thread> workers;
    // consumed jobs are reused: their buffers keep their capacity, so the steady state allocates nothing (fresh
    // multi-megabyte allocations mean mmap / page faults / munmap, which serialise the threads on the address space)
    std::vector<std::shared_ptr<Job> > pool;
    std::shared_ptr<Job> held;  // batch mode: the job whose output the caller is looking at
    std::shared_ptr<Job> obtain() {
        std::lock_guard<std::mutex> l(mu);
        if (pool.empty()) return std::make_shared<Job>();
        std::shared_ptr<Job> j = pool.back();
        pool.pop_back();
        return j;
    }
    void recycle(std::shared_ptr<Job>& j) {
        if (!j) return;
        j->recs.clear();
        j->cigars.clear();
        j->out.clear();
        j->bad = j->unknownName = j->done = false;
        j->textLen = 0;
        std::lock_guard<std::mutex> l(mu);
        pool.push_back(std::move(j));
        j.reset();
    }

    SamPipeline(FILE* f, int threads, const uint8_t* initial, size_t nInitial, const std::unordered_map<std::string, int32_t>& nm, RecordFn fn_ = nullptr,
                void* user_ = nullptr)
        : fp(f), fn(fn_), user(user_), names(nm), carry((const char*)initial, (const char*)initial + nInitial), maxOutstanding((size_t)threads * 3 + 2) {
        struct stat st;
        const long at = ftell(fp);  // (the header parser has read `nInitial` bytes beyond what it consumed)
        if (!getenv("PPP_NO_MMAP") && at >= 0 && (size_t)at >= nInitial && fstat(fileno(fp), &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0 &&
            (uint64_t)at <= (uint64_t)st.st_size) {
            void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fileno(fp), 0);
            if (m != MAP_FAILED) {
                madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
                map = (const char*)m;
                mapLen = (size_t)st.st_size;
                mapPos = (size_t)at - nInitial;
                carry.clear();
            }
        }
        reader = std::thread([this] { if (map) mapLoop(); else readLoop(); });
        for (int i = 0; i < threads; ++i) workers.emplace_back([this] { workLoop(); });
    }
    ~SamPipeline() {
        {
            std::lock_guard<std::mutex> l(mu);
            stop = true;
        }
        cvWork.notify_all();
        cvSpace.notify_all();
        cvDone.notify_all();
        reader.join();
        for (auto& w : workers) w.join();
        if (map) munmap(const_cast<char*>(map), mapLen);
    }
    // Hands a job with `textLen` bytes of whole lines to the workers; false when the pipeline is being torn down.
    bool submit(std::shared_ptr<Job>& job) {
        std::unique_lock<std::mutex> l(mu);
        cvSpace.wait(l, [this] { return stop || ordered.size() < maxOutstanding; });
        if (stop) return false;
        if (job->textLen) {
            ordered.push_back(job);
            todo.push_back(job);
            cvWork.notify_one();
        }
        return true;
    }
    void finishInput() {
        std::lock_guard<std::mutex> l(mu);
        eof = true;
        cvDone.notify_all();
        cvWork.notify_all();
    }
    // Memory-mapped input: jobs of ~2 MB, each ending behind a line feed (the last one at the end of the file).
    void mapLoop() {
        const size_t kJobBytes = 2u << 20;
        size_t pos = mapPos;
        while (pos < mapLen) {
            size_t end = std::min(pos + kJobBytes, mapLen);
            if (end < mapLen) {
                const void* nl = memrchr(map + pos, '\n', end - pos);
                if (nl) end = (size_t)((const char*)nl - map) + 1;
                else {  // a line longer than a job: up to its end
                    const void* fwd = memchr(map + end, '\n', mapLen - end);
                    end = fwd ? (size_t)((const char*)fwd - map) + 1 : mapLen;
                }
            }
            std::shared_ptr<Job> job = obtain();
            job->data = map + pos;
            job->textLen = end - pos;
            if (!submit(job)) return;
            pos = end;
        }
        finishInput();
    }
    void readLoop() {
        const size_t kJobBytes = 2u << 20;
        bool fileEnd = false;
        while (!fileEnd) {
            std::shared_ptr<Job> job = obtain();
            const size_t have = carry.size();
            if (job->textCap < have + kJobBytes) {
                job->textCap = have + kJobBytes + (64u << 10);
                job->text.reset(new char[job->textCap]);
            }
            if (have) memcpy(job->text.get(), carry.data(), have);
            size_t got = fread(job->text.get() + have, 1, kJobBytes, fp);
            size_t total = have + got;
            if (got == 0) {
                fileEnd = true;  // the rest (a last line without '\n') is the final job
                carry.clear();
            } else {
                size_t cut = total;
                while (cut > 0 && job->text[cut - 1] != '\n') --cut;
                if (cut == 0) { carry.assign(job->text.get(), job->text.get() + total); recycle(job); continue; }  // no complete line yet
                carry.assign(job->text.get() + cut, job->text.get() + total);
                total = cut;
            }
            job->data = job->text.get();
            job->textLen = total;
            if (!submit(job)) return;
        }
        finishInput();
    }
    void workLoop() {
        SamNameCache cache;
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> l(mu);
                cvWork.wait(l, [this] { return stop || !todo.empty() || eof; });
                if (stop) return;
                if (todo.empty()) { if (eof) return; continue; }
                job = todo.front();
                todo.pop_front();
            }
            const char* t = job->data;
            const size_t n = job->textLen;
            job->recs.reserve(n / 96);
            size_t p = 0;
            while (p < n) {
                if (t[p] == '\n' || t[p] == '\r') { ++p; continue; }  // blank lines are not records
                const char* nl = (const char*)memchr(t + p, '\n', n - p);
                size_t e = nl ? (size_t)(nl - t) : n, len = e - p;
                if (len && t[p + len - 1] == '\r') --len;
                Rec r;
                r.lineOff = (uint32_t)p;
                r.lineLen = (uint32_t)len;
                int rc = parseSamLine(t + p, len, names, cache, job->cigars, r.f);
                if (rc == 1) { job->bad = true; break; }
                job->recs.push_back(r);
                if (rc == 2) job->unknownName = true;
                p = e + 1;
            }
            if (fn && !job->unknownName) {
                AlnView v;
                for (const Rec& r : job->recs) {
                    fillView(v, r.f, t + r.lineOff, job->cigars.data());
                    fn(user, v, job->out);
                }
            }
            {
                std::lock_guard<std::mutex> l(mu);
                job->done = true;
            }
            cvDone.notify_all();
        }
    }
    // Next parsed job in file order; null at end of file.
    std::shared_ptr<Job> next() {
        std::unique_lock<std::mutex> l(mu);
        cvDone.wait(l, [this] { return (!ordered.empty() && ordered.front()->done) || (ordered.empty() && eof); });
        if (ordered.empty()) return nullptr;
        std::shared_ptr<Job> j = ordered.front();
        ordered.pop_front();
        cvSpace.notify_one();
        return j;
    }
};

// ---------------------------------------------------------------------------
// BAM batch mode, consumer side.  The worker that inflated a job has already decoded the records that lie wholly
// inside it, starting at a boundary it GUESSED (Pipeline::findRecordStart); this stage, which sees the jobs in file
// order, knows the true boundary -- it owns the bytes of the record that the previous job left unfinished -- and
//   * accepts the guess when those bytes plus the job's bytes in front of the guess are a whole number of records
//     (walking from a true boundary and landing on the guess proves that the guess is one), decoding just these few
//     stitched records itself, or
//   * otherwise decodes the whole job again from the true boundary (never observed on real files; the tests force it).
// Either way the records and their order are exactly those of the sequential walk.
// ---------------------------------------------------------------------------
struct AlnReader::BamBatcher {
    typedef AlnReader::RecordFn RecordFn;
    Pipeline* src;
    RecordFn fn;
    void* user;
    std::vector<uint8_t> carry;  // starts at a true record boundary: bytes the header parser had read ahead, then unfinished records
    std::shared_ptr<Pipeline::Job> held;  // the job whose output the caller is looking at
    std::vector<uint8_t> stitchedOut, redoOut;
    std::vector<uint32_t> cigarScratch;
    bool finished = false;
    uint64_t nGuessesRejected = 0;

    BamBatcher(Pipeline* p, RecordFn f, void* u, const uint8_t* initial, size_t nInitial) : src(p), fn(f), user(u), carry(initial, initial + nInitial) {}

    // Decodes the whole records at the front of `carry` into `out` and removes them.  False: a malformed record.
    bool drainCarry(std::vector<uint8_t>& out, uint64_t& nRecords) {
        AlnView v;
        size_t q = 0;
        bool ok = true;
        while (carry.size() - q >= 4) {
            uint32_t bs;
            memcpy(&bs, carry.data() + q, 4);
            if (carry.size() - q - 4 < (size_t)bs) break;
            if (!parseBamRecord(carry.data() + q + 4, bs, cigarScratch, v)) { ok = false; break; }
            fn(user, v, out);
            ++nRecords;
            q += 4 + (size_t)bs;
        }
        carry.erase(carry.begin(), carry.begin() + q);
        return ok;
    }
    // True when carry + data[0, guess) is a whole number of records (then `guess` is a record boundary).
    bool guessHolds(const uint8_t* data, size_t guess) const {
        // the records in question: all of carry, continued by the first `guess` bytes of the job
        const size_t total = carry.size() + guess;
        auto byteAt = [&](size_t i) { return i < carry.size() ? carry[i] : data[i - carry.size()]; };
        size_t q = 0;
        while (q < total) {
            if (total - q < 4) return false;
            const uint32_t bs = (uint32_t)byteAt(q) | ((uint32_t)byteAt(q + 1) << 8) | ((uint32_t)byteAt(q + 2) << 16) | ((uint32_t)byteAt(q + 3) << 24);
            if (total - q - 4 < (size_t)bs) return false;
            q += 4 + (size_t)bs;
        }
        return true;
    }
    // Next job's records in file order.  False at the end of the input.
    bool next(const uint8_t*& out, size_t& outLen, uint64_t& nRecords, bool& bad) {
        out = nullptr;
        outLen = 0;
        nRecords = 0;
        bad = false;
        if (held) { src->recycle(held); }
        if (finished) return false;
        std::shared_ptr<Pipeline::Job> j = src->nextJob();
        if (!j) {
            finished = true;
            if (carry.empty()) return false;
            bad = true;  // a truncated record at the end is malformed
            return true;
        }
        held = j;
        if (j->bad) { bad = true; finished = true; return true; }
        const uint8_t* data = j->out.get();
        const size_t len = j->outLen;
        if (j->parsed && guessHolds(data, j->specStart)) {
            stitchedOut.clear();
            if (j->specStart || !carry.empty()) {
                carry.insert(carry.end(), data, data + j->specStart);
                if (!drainCarry(stitchedOut, nRecords)) { bad = true; finished = true; }
            }
            if (!bad) {
                // the worker left room in front of its output for the few stitched records
                std::vector<uint8_t>& ro = j->recOut;
                size_t begin = Pipeline::kHeadroom;
                if (stitchedOut.size() <= begin) {
                    begin -= stitchedOut.size();
                    if (!stitchedOut.empty()) memcpy(ro.data() + begin, stitchedOut.data(), stitchedOut.size());
                } else {
                    ro.insert(ro.begin() + (ptrdiff_t)begin, stitchedOut.begin(), stitchedOut.end());
                }
                out = ro.data() + begin;
                outLen = ro.size() - begin;
                nRecords += j->nParsed;
                carry.assign(data + j->tailOff, data + len);
                if (j->parseBad) { bad = true; finished = true; }
                return true;
            }
            out = stitchedOut.data();
            outLen = stitchedOut.size();
            return true;
        }
        // no usable guess: everything from the true boundary, here
        if (j->parsed) ++nGuessesRejected;
        redoOut.clear();
        carry.insert(carry.end(), data, data + len);
        if (!drainCarry(redoOut, nRecords)) { bad = true; finished = true; }
        out = redoOut.data();
        outLen = redoOut.size();
        return true;
    }
};

AlnReader::AlnReader() {}
AlnReader::~AlnReader() { close(); }

void AlnReader::setThreads(int n) {
    if (!isBam_ || !good_ || pipe_ || n <= 1) return;
    pipe_ = new Pipeline(fp_, n);  // continues at the current file offset (after the blocks already consumed)
}

void AlnReader::close() {
    delete bamBatch_;  // (before its source)
    bamBatch_ = nullptr;
    batchFn_ = nullptr;
    batchFailed_ = false;
    delete pipe_;
    pipe_ = nullptr;
    delete samPipe_;
    samPipe_ = nullptr;
    extraNames_.clear();
    nameCache_.clear();
    if (fp_) fclose(fp_);
    fp_ = nullptr;
    delete nameMap_;
    nameMap_ = nullptr;
    good_ = false;
}

// ---------------------------------------------------------------------------
// raw / BGZF input
// ---------------------------------------------------------------------------
static inline void compact(std::vector<uint8_t>& buf, size_t& cur, size_t& end) {
    if (cur == 0) return;
    if (end > cur) memmove(buf.data(), buf.data() + cur, end - cur);
    end -= cur;
    cur = 0;
}

bool AlnReader::readBgzfBlock() {
    uint8_t hdr[12];
    size_t got = fread(hdr, 1, 12, fp_);
    if (got == 0) { eofRaw_ = true; return false; }
    if (got != 12 || hdr[0] != 31 || hdr[1] != 139 || hdr[2] != 8 || !(hdr[3] & 4)) { good_ = false; eofRaw_ = true; return false; }
    uint32_t xlen = hdr[10] | (hdr[11] << 8);
    uint8_t extra[65536];
    if (fread(extra, 1, xlen, fp_) != xlen) { good_ = false; eofRaw_ = true; return false; }
    int bsize = -1;
    for (uint32_t i = 0; i + 4 <= xlen;) {
        uint32_t slen = extra[i + 2] | (extra[i + 3] << 8);
        if (extra[i] == 'B' && extra[i + 1] == 'C' && slen == 2) bsize = extra[i + 4] | (extra[i + 5] << 8);
        i += 4 + slen;
    }
    if (bsize < 0) { good_ = false; eofRaw_ = true; return false; }
    size_t clen = (size_t)bsize + 1 - xlen - 12 - 8;
    zin_.resize(clen + 8);
    if (fread(zin_.data(), 1, clen + 8, fp_) != clen + 8) { good_ = false; eofRaw_ = true; return false; }
    uint32_t isize;
    memcpy(&isize, zin_.data() + clen + 4, 4);
    if (isize == 0) return true;  // empty block (e.g. the EOF marker)
    if (buf_.size() < end_ + isize) buf_.resize(end_ + isize + kRawChunk);
    if (fast_inflate(zin_.data(), clen, buf_.data() + end_, isize)) {
        end_ += isize;
        return true;
    }
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) { good_ = false; return false; }
    zs.next_in = zin_.data();
    zs.avail_in = (uInt)clen;
    zs.next_out = buf_.data() + end_;
    zs.avail_out = isize;
    int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (rc != Z_STREAM_END || zs.avail_out != 0) { good_ = false; eofRaw_ = true; return false; }
    end_ += isize;
    return true;
}

bool AlnReader::fill() {
    if (eofRaw_) return false;
    compact(buf_, cur_, end_);
    if (isBam_ && pipe_) {
        const uint8_t* data;
        size_t len;
        bool bad = false;
        size_t before = end_;
        while (end_ == before) {
            if (!pipe_->next(data, len, bad)) { eofRaw_ = true; return false; }
            if (bad) { good_ = false; eofRaw_ = true; return false; }
            if (buf_.size() < end_ + len) buf_.resize(end_ + len + kRawChunk);
            memcpy(buf_.data() + end_, data, len);
            end_ += len;
        }
        return true;
    }
    if (isBam_) {
        size_t before = end_;
        while (!eofRaw_ && end_ == before) readBgzfBlock();
        return end_ > before;
    }
    if (buf_.size() < end_ + kRawChunk) buf_.resize(end_ + kRawChunk);
    size_t got = fread(buf_.data() + end_, 1, kRawChunk, fp_);
    if (got == 0) { eofRaw_ = true; return false; }
    end_ += got;
    return true;
}

bool AlnReader::ensure(size_t n) {
    while (end_ - cur_ < n)
        if (!fill()) return false;
    return true;
}

// Returns one text line without its terminator ('\n' or "\r\n").
bool AlnReader::getLine(const char*& line, size_t& len) {
    size_t scanFrom = cur_;
    for (;;) {
        const uint8_t* p = (const uint8_t*)memchr(buf_.data() + scanFrom, '\n', end_ - scanFrom);
        if (p) {
            line = (const char*)buf_.data() + cur_;
            len = (size_t)(p - (buf_.data() + cur_));
            cur_ += len + 1;
            if (len && line[len - 1] == '\r') --len;
            return true;
        }
        size_t have = end_ - cur_;
        if (!fill()) {
            if (have == 0) return false;
            line = (const char*)buf_.data() + cur_;  // last line without '\n'
            len = have;
            cur_ = end_;
            if (len && line[len - 1] == '\r') --len;
            return true;
        }
        scanFrom = cur_ + have;  // fill() compacted: cur_ == 0
    }
}

// ---------------------------------------------------------------------------
// open / header
// ---------------------------------------------------------------------------
bool AlnReader::parseSamHeaderLine(const char* line, size_t len) {
    if (len < 3 || line[1] != 'S' || line[2] != 'Q') return true;  // only @SQ matters
    std::string name;
    uint64_t ln = 0;
    size_t i = 3;
    while (i < len) {
        if (line[i] == '\t') { ++i; continue; }
        size_t j = i;
        while (j < len && line[j] != '\t') ++j;
        if (j - i >= 3 && line[i + 2] == ':') {
            if (line[i] == 'S' && line[i + 1] == 'N') name.assign(line + i + 3, j - i - 3);
            else if (line[i] == 'L' && line[i + 1] == 'N') ln = strtoull(std::string(line + i + 3, j - i - 3).c_str(), nullptr, 10);
        }
        i = j;
    }
    nameMap_->m.emplace(name, (int32_t)names_.size());
    names_.push_back(name);
    lengths_.push_back(ln);
    return true;
}

bool AlnReader::open(const char* path) {
    close();
    fp_ = fopen(path, "rb");
    if (!fp_) return false;
    setvbuf(fp_, nullptr, _IOFBF, 1 << 20);
    nameMap_ = new NameMap;
    names_.clear();
    lengths_.clear();
    cur_ = end_ = 0;
    eofRaw_ = false;
    nRecords_ = 0;
    good_ = true;
    int c0 = fgetc(fp_), c1 = fgetc(fp_);
    isBam_ = (c0 == 31 && c1 == 139);
    rewind(fp_);
    if (isBam_) {
        if (!ensure(12) || memcmp(buf_.data() + cur_, "BAM\1", 4) != 0) { good_ = false; return false; }
        uint32_t lText;
        memcpy(&lText, buf_.data() + cur_ + 4, 4);
        cur_ += 8;
        if (!ensure((size_t)lText + 4)) { good_ = false; return false; }
        cur_ += lText;  // the reference table below is authoritative for names and lengths
        uint32_t nRef;
        memcpy(&nRef, buf_.data() + cur_, 4);
        cur_ += 4;
        for (uint32_t r = 0; r < nRef; ++r) {
            if (!ensure(4)) { good_ = false; return false; }
            uint32_t lName;
            memcpy(&lName, buf_.data() + cur_, 4);
            cur_ += 4;
            if (!ensure((size_t)lName + 4)) { good_ = false; return false; }
            std::string name((const char*)buf_.data() + cur_, lName ? lName - 1 : 0);
            cur_ += lName;
            uint32_t lRef;
            memcpy(&lRef, buf_.data() + cur_, 4);
            cur_ += 4;
            names_.push_back(name);
            lengths_.push_back(lRef);
        }
        return good_;
    }
    // SAM: consume '@' lines
    for (;;) {
        if (end_ == cur_ && !fill()) break;
        if (buf_[cur_] != '@') break;
        const char* line;
        size_t len;
        if (!getLine(line, len)) break;
        parseSamHeaderLine(line, len);
    }
    return good_;
}

bool AlnReader::atEnd() {
    if (!good_ && !fp_) return true;
    for (;;) {
        if (!isBam_)  // SAM: blank lines (e.g. a trailing newline) are not records
            while (cur_ < end_ && (buf_[cur_] == '\n' || buf_[cur_] == '\r')) ++cur_;
        if (end_ > cur_) return false;
        if (!fill()) return true;
    }
}

// ---------------------------------------------------------------------------
// SAM record
// ---------------------------------------------------------------------------
int AlnReader::nextSam(AlnView& out) {
    SamFields fld;
    const char* line;
    size_t len;
    if (atEnd() || !getLine(line, len) || len == 0) return 1;
    cigarScratch_.clear();
    if (parseSamLine(line, len, nameMap_->m, nameCache_, cigarScratch_, fld) == 1) return 1;
    if (fld.rID == -2) fld.rID = internExtraName(std::string(line + fld.nameOff, fld.nameLen));
    fillView(out, fld, line, cigarScratch_.data());
    ++nRecords_;
    return 0;
}

// A contig that is missing from the header: the name store grows in order of appearance.
int32_t AlnReader::internExtraName(const std::string& name) {
    auto it = extraNames_.find(name);
    if (it != extraNames_.end()) return it->second;
    int32_t id = (int32_t)names_.size();
    extraNames_.emplace(name, id);
    names_.push_back(name);
    lengths_.push_back(0);
    return id;
}

// ---------------------------------------------------------------------------
// BAM record
// ---------------------------------------------------------------------------
int AlnReader::nextBam(AlnView& out) {
    if (!ensure(4)) return 1;
    uint32_t blockSize;
    memcpy(&blockSize, buf_.data() + cur_, 4);
    if (blockSize < 32 || !ensure((size_t)blockSize + 4)) return 1;
    const uint8_t* p = buf_.data() + cur_ + 4;
    cur_ += (size_t)blockSize + 4;
    if (!parseBamRecord(p, blockSize, cigarScratch_, out)) return 1;
    ++nRecords_;
    return 0;
}

// ---------------------------------------------------------------------------
// batch mode
// ---------------------------------------------------------------------------
void AlnReader::startBatches(int threads, RecordFn fn, void* user) {
    if (!good_ || pipe_ || samPipe_ || bamBatch_ || !fn) return;
    if (threads < 1) threads = 1;
    batchFn_ = fn;
    batchUser_ = user;
    if (isBam_) {
        pipe_ = new Pipeline(fp_, threads, fn, user, (int32_t)names_.size());
        bamBatch_ = new BamBatcher(pipe_, fn, user, buf_.data() + cur_, end_ - cur_);
    } else {
        samPipe_ = new SamPipeline(fp_, threads, buf_.data() + cur_, end_ - cur_, nameMap_->m, fn, user);
    }
    cur_ = end_ = 0;
}

bool AlnReader::nextBatch(const uint8_t*& out, size_t& outLen, uint64_t& nRecords, int& error) {
    out = nullptr;
    outLen = 0;
    nRecords = 0;
    error = 0;
    if (batchFailed_ || !batchFn_) return false;
    if (bamBatch_) {
        bool bad = false;
        if (!bamBatch_->next(out, outLen, nRecords, bad)) return false;
        nRecords_ += nRecords;
        if (bad) { error = 1; batchFailed_ = true; }
        return true;
    }
    if (samPipe_->held) samPipe_->recycle(samPipe_->held);  // the caller is done with the previous job's output
    std::shared_ptr<SamPipeline::Job> j = samPipe_->next();
    if (!j) return false;
    if (j->unknownName) {  // a contig that is missing from the header: this job is redone here, where the name store may grow
        AlnView v;
        for (const SamPipeline::Rec& r : j->recs) {
            SamFields f = r.f;
            const char* line = j->data + r.lineOff;
            if (f.rID == -2) f.rID = internExtraName(std::string(line + f.nameOff, f.nameLen));
            fillView(v, f, line, j->cigars.data());
            batchFn_(batchUser_, v, j->out);
        }
    }
    out = j->out.data();
    outLen = j->out.size();
    nRecords = j->recs.size();
    nRecords_ += nRecords;
    if (j->bad) { error = 1; batchFailed_ = true; }
    samPipe_->held = j;
    return true;
}

uint64_t AlnReader::rejectedBoundaryGuesses() const { return bamBatch_ ? bamBatch_->nGuessesRejected : 0; }

int AlnReader::next(AlnView& out) {
    if (!good_) return 1;
    return isBam_ ? nextBam(out) : nextSam(out);
}

}  // namespace ppp
