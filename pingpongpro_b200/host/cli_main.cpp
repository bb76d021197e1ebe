// cli_main.cpp -- the drop-in `pingpongpro` command: same flags, inputs, messages, exit codes and output files as
// the reference's main() (pingpongpro.cpp:1461-1638).  Alignment decode and feature extraction run here on the
// host (frontend.h) and feed pinned, double-buffered uploads; stacks, overlap scan, FDRs and transposon scores
// are computed by the CUDA layer through the C ABI of ppp_b200.h.  There is no CPU fallback.
#include <sys/stat.h>
#include <sys/wait.h>
#include <fcntl.h>
#include <signal.h>
#include <unistd.h>

#include <cerrno>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "aln_reader.h"
#include "annotation.h"
#include "cli_args.h"
#include "frontend.h"
#include "ppp_b200.h"
#include "writers.h"

using namespace ppp;

namespace {

// stopwatch of pingpongpro.cpp:364-390: alternate calls start / stop, 1-second resolution, stderr, only with -v.
unsigned stopwatch(const std::string& operation, unsigned verbosity) {
    static time_t start = 0;
    unsigned elapsed = 0;
    if (start != 0) { elapsed = (unsigned)(time(nullptr) - start); start = 0; }
    else start = time(nullptr);
    if (verbosity >= 3) {
        if (!operation.empty()) std::cerr << operation << " ... ";
        else std::cerr << "done (" << elapsed << " seconds)" << std::endl;
    }
    std::cerr.flush();
    return elapsed;
}
unsigned stopwatch(unsigned verbosity) { return stopwatch("", verbosity); }

// PPP_TIMING=1: wall-clock marks on stderr (not part of the reference's interface)
void mark(const char* what) {
    static const bool on = getenv("PPP_TIMING") != nullptr;
    static const auto t0 = std::chrono::steady_clock::now();
    if (on) fprintf(stderr, "[ppp %8.3f s] %s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), what);
}

int gpuError(const char* what) {
    std::cerr << "pingpongpro: " << what << ": " << ppp_last_error() << std::endl;
    return 1;
}

// Pinned double buffer that streams one contig's reads to the GPU in file order.
struct Uploader {
    static const size_t kBatch = (size_t)1 << 20;
    ppp_ctx* ctx = nullptr;
    // two pinned buffers in the 5-byte upload form of ppp_reads_submit_packed: u32 keys + u8 meta (strand, A10, NH code)
    uint32_t* pinnedKeys[2] = {nullptr, nullptr};
    uint8_t* pinnedMeta[2] = {nullptr, nullptr};
    int which = 0;
    std::vector<std::vector<ppp_read> > pending;  // per contig
    uint64_t submitted = 0;

    void prepare(size_t nContigs) { pending.assign(nContigs, std::vector<ppp_read>()); }
    // Called once the context exists (it is created on a helper thread while decoding has already started).
    bool attach(ppp_ctx* c) {
        ctx = c;
        for (int k = 0; k < 2; ++k) {
            if (ppp_pinned_alloc(kBatch * sizeof(uint32_t), (void**)&pinnedKeys[k]) != PPP_OK) return false;
            if (ppp_pinned_alloc(kBatch, (void**)&pinnedMeta[k]) != PPP_OK) return false;
        }
        return true;
    }
    bool flush(size_t contig) {
        std::vector<ppp_read>& v = pending[contig];
        for (size_t o = 0; o < v.size();) {
            // pack up to kBatch reads; the multi-hit counts travel as 6-bit codes into a table of the (at most 64) distinct
            // values of the piece -- a 65th value closes the piece early
            uint32_t table[PPP_PACKED_NH_CODES];
            int nCodes = 0;
            int8_t codeOfSmall[256];
            memset(codeOfSmall, -1, sizeof codeOfSmall);
            uint32_t* keys = pinnedKeys[which];
            uint8_t* meta = pinnedMeta[which];
            const size_t end = std::min(v.size(), o + kBatch);
            size_t j = o;
            for (; j < end; ++j) {
                const uint32_t nh = v[j].meta >> PPP_READ_NH_SHIFT;
                int code = nh < 256 ? codeOfSmall[nh] : -1;
                if (code < 0) {
                    for (int t = 0; t < nCodes && code < 0; ++t)
                        if (table[t] == nh) code = t;
                    if (code < 0) {
                        if (nCodes == PPP_PACKED_NH_CODES) break;
                        code = nCodes;
                        table[nCodes++] = nh;
                    }
                    if (nh < 256) codeOfSmall[nh] = (int8_t)code;
                }
                keys[j - o] = v[j].key;
                meta[j - o] = (uint8_t)((v[j].meta & 3u) | ((uint32_t)code << 2));
            }
            for (int t = nCodes; t < PPP_PACKED_NH_CODES; ++t) table[t] = 1;
            if (ppp_reads_submit_packed(ctx, (int32_t)contig, keys, meta, table, j - o) != PPP_OK) return false;
            submitted += j - o;
            which ^= 1;  // the other buffer may be refilled as soon as this submit has returned
            o = j;
        }
        v.clear();
        return true;
    }
    // appends n consecutive reads of one contig (file order)
    bool addRun(size_t contig, const ExtractedRead* e, size_t n) {
        std::vector<ppp_read>& v = pending[contig];
        const size_t old = v.size();
        v.resize(old + n);
        for (size_t i = 0; i < n; ++i) v[old + i] = e[i].read;
        return v.size() < kBatch || !ctx || flush(contig);  // without a context yet, reads simply accumulate
    }
    // the same for a range shard: keeps the reads whose key lies in [lo, hi) of their contig (the context's own range plus
    // its halo -- a superset is fine, the context ignores what it does not own)
    bool addRunRouted(size_t contig, const ExtractedRead* e, size_t n, uint32_t lo, uint32_t hi) {
        std::vector<ppp_read>& v = pending[contig];
        for (size_t i = 0; i < n; ++i)
            if (e[i].read.key >= lo && e[i].read.key < hi) v.push_back(e[i].read);
        return v.size() < kBatch || !ctx || flush(contig);
    }
    bool flushAll() {
        for (size_t c = 0; c < pending.size(); ++c)
            if (!pending[c].empty() && !flush(c)) return false;
        return true;
    }
    ~Uploader() {
        for (int k = 0; k < 2; ++k) {
            if (pinnedKeys[k]) ppp_pinned_free(pinnedKeys[k]);
            if (pinnedMeta[k]) ppp_pinned_free(pinnedMeta[k]);
        }
    }
};

// One position range of the genome on one GPU (--gpus N; a single shard without a range otherwise).
struct Shard {
    ppp_ctx* ctx = nullptr;
    Uploader up;
    uint64_t rangeBegin = 0, rangeEnd = 0;
    std::vector<uint32_t> lo, hi;  // per contig: keys worth sending to this shard (own range + minus-strand halo)
    std::string error;             // ppp_last_error() is per thread: kept by the thread that saw it
};

// Keys of every contig that a shard accepts (ppp_config.range_begin / range_end, ppp_b200.h): its own range, up to
// 64 keys beyond the contig when it holds the contig's last base, plus max_overlap keys of halo.
void shardKeyRanges(Shard& sh, const std::vector<uint64_t>& lengths, int maxOverlap) {
    sh.lo.assign(lengths.size(), 0);
    sh.hi.assign(lengths.size(), 0);
    uint64_t cum = 0;
    for (size_t c = 0; c < lengths.size(); ++c) {
        const uint64_t L = lengths[c];
        const uint64_t lo = sh.rangeBegin > cum ? sh.rangeBegin - cum : 0;
        const bool holdsEnd = sh.rangeEnd >= cum + L;  // ... and with it every key beyond the contig (the context reports those)
        const uint64_t hi = holdsEnd ? 0xffffffffull : (sh.rangeEnd > cum ? sh.rangeEnd - cum : 0);
        if (L != 0 && lo < L && hi > lo) {
            sh.lo[c] = (uint32_t)lo;
            sh.hi[c] = (uint32_t)std::min<uint64_t>(hi + (uint64_t)maxOverlap + 1, 0xffffffffull);
        }
        cum += L;
    }
}

int shardError(const std::vector<Shard>& shards) {
    for (size_t k = 0; k < shards.size(); ++k)
        if (!shards[k].error.empty()) std::cerr << "pingpongpro: " << shards[k].error << (shards.size() > 1 ? " (shard " + std::to_string(k) + ")" : std::string()) << std::endl;
    return 1;
}

// ppp_height_hist + ppp_scan of every shard; with several shards one thread each, because the calls meet inside the
// group's all-reduces (height histogram, grouped counts).
bool histAndScan(std::vector<Shard>& shards, ppp_group* group) {
    auto one = [&](Shard& sh) {
        if (ppp_height_hist(sh.ctx, nullptr, nullptr) != PPP_OK) { sh.error = std::string("height histogram failed: ") + ppp_last_error(); return false; }
        if (ppp_scan(sh.ctx, nullptr, nullptr) != PPP_OK) { sh.error = std::string("overlap scan failed: ") + ppp_last_error(); return false; }
        return true;
    };
    if (shards.size() == 1) return one(shards[0]);
    std::vector<std::thread> th;
    std::vector<char> ok(shards.size(), 0);
    for (size_t k = 0; k < shards.size(); ++k)
        th.emplace_back([&, k] {
            ok[k] = one(shards[k]) ? 1 : 0;
            if (!ok[k]) ppp_group_abort(group);  // the other ranks must not wait for this one
        });
    for (std::thread& t : th) t.join();
    for (char c : ok)
        if (!c) return false;
    return true;
}

// ppp_score of every shard (no collective inside); the loci of the shards concatenate in rank order = (contig, position).
bool scoreAll(std::vector<Shard>& shards, uint32_t minStackHeight, uint32_t* nBins, float* cells, std::vector<ppp_locus>& joined, const ppp_locus*& loci, size_t& nLoci) {
    joined.clear();
    for (size_t k = 0; k < shards.size(); ++k) {
        const ppp_locus* l = nullptr;
        size_t n = 0;
        if (ppp_score(shards[k].ctx, minStackHeight, k == 0 ? nBins : nullptr, nullptr, k == 0 ? cells : nullptr, nullptr, &l, &n) != PPP_OK) {
            shards[k].error = std::string("scoring failed: ") + ppp_last_error();
            return false;
        }
        if (shards.size() == 1) { loci = l; nLoci = n; return true; }
        joined.insert(joined.end(), l, l + n);
    }
    loci = joined.data();
    nLoci = joined.size();
    return true;
}

std::vector<ScoredTransposon> scoreTransposons(std::vector<Shard>& shards, const TransposonsPerGenome& tp, bool& ok) {
    std::vector<ScoredTransposon> rows;
    std::vector<ppp_interval> iv;
    for (const auto& kv : tp)
        for (const Transposon& t : kv.second) {
            ScoredTransposon s;
            s.t = t;
            s.contig = kv.first;
            memset(&s.r, 0, sizeof s.r);
            rows.push_back(s);
            iv.push_back(ppp_interval{kv.first, t.strand, t.start, t.end});
        }
    std::vector<ppp_tp_result> res(iv.size());
    std::vector<uint32_t> order(iv.size());
    if (shards.size() == 1) ok = iv.empty() || ppp_transposons(shards[0].ctx, iv.data(), iv.size(), res.data(), order.data()) == PPP_OK;
    else {
        std::vector<ppp_ctx*> ctxs;
        for (Shard& sh : shards) ctxs.push_back(sh.ctx);
        ok = iv.empty() || ppp_transposons_sharded(ctxs.data(), (int)ctxs.size(), iv.data(), iv.size(), res.data(), order.data()) == PPP_OK;
    }
    if (ok)
        for (size_t k = 0; k < rows.size(); ++k) rows[order[k]].r = res[k];  // results come back in (contig, start, end) order = our order
    return rows;
}

void plotTransposons(const std::vector<ScoredTransposon>& rows, const std::string& fileName, const AppOptions& o, int W) {
    std::vector<std::string> titles;
    std::vector<std::vector<float> > hist;
    for (const ScoredTransposon& s : rows) {
        std::ostringstream ss;
        ss << "z-scores of transposon " << s.t.identifier << std::endl << "(p-value for overlap of " << 10 << " nt = " << s.r.p_value << ")";  // :1450-1452
        titles.push_back(ss.str());
        hist.push_back(std::vector<float>(s.r.histogram, s.r.histogram + W));
    }
    plotHistograms(fileName, titles, hist, o.minOverlap, o.maxOverlap, 10);
}

}  // namespace

// The whole program; returns the exit status (the CUDA runtime's own process-exit teardown is skipped by the caller).
static int run(int argc, char const** argv) {
    AppOptions options;
    if (parseCommandLine(options, argc, argv) != PARSE_OK) return 1;  // :1465-1466 (help and version too)
    mark("start");

    std::vector<std::string> names;   // bamNameStore
    std::vector<uint64_t> lengths;
    // --gpus N: the genome is cut into N position ranges, one context (GPU) each; SURVEY §8(e)
    const int nShards = options.gpus > 1 ? options.gpus : 1;
    std::vector<Shard> shards(nShards);
    ppp_group* group = nullptr;
    bool haveContexts = false;
    double totalReadCount = 0;
    const int W = options.maxOverlap - options.minOverlap + 1, ppIdx = 10 - options.minOverlap;

    if (options.verbosity >= 3) std::cerr << "Counting reads in SAM/BAM files" << std::endl;
    for (const std::string& path : options.inputFiles) {
        stopwatch(std::string("  ") + path, options.verbosity);
        ExtractOptions eo;  // read by the reader's worker threads: declared before the reader, so that it outlives them on every return path
        eo.minLen = options.minAlignmentLength;
        eo.maxLen = options.maxAlignmentLength;
        eo.mode = options.countMultiHits;
        AlnReader reader;
        if (!reader.open(path.c_str())) {
            std::cerr << "Failed to open input file: " << path << std::endl;  // :1485
            return 1;
        }
        if (!haveContexts) {
            names = reader.names();
            lengths = reader.lengths();
            if (lengths.empty() && !reader.isBam()) {
                // SAM text without @SQ lines: the reference (SeqAn) grows its name store while it reads and never needs a
                // length (:1493-1517).  The dense stack arrays do, so such input is read twice: a first pass over every input
                // file collects the names (in order of appearance) and the extent of the kept reads per contig.
                const int scanThreads = options.threads > 0 ? options.threads : (int)std::thread::hardware_concurrency();
                for (const std::string& p : options.inputFiles) {
                    std::vector<std::string> nm;
                    std::vector<uint64_t> ext;
                    const int rc = scan_extents(p.c_str(), eo, nm, ext, scanThreads > 2 ? scanThreads - 2 : 1);
                    if (rc == 1) {
                        std::cerr << "Failed to open input file: " << p << std::endl;  // :1485
                        return 1;
                    }
                    if (rc != 0) {
                        std::cerr << "Failed to read record" << std::endl;  // :411
                        return 1;
                    }
                    if (&p == &options.inputFiles.front()) {
                        names = nm;
                        lengths = ext;
                    } else if (nm != names) {
                        std::cerr << "@SQ header lines of '" << p << "' differ from those of previous input files" << std::endl;  // :1514
                        return 1;
                    } else {
                        for (size_t c = 0; c < lengths.size(); ++c) lengths[c] = std::max(lengths[c], ext[c]);
                    }
                }
                mark("headerless input scanned");
            }
            if (lengths.empty()) lengths.push_back(0);
            mark("header read");
            // CUDA start-up (cuInit + context creation, 0.4-0.8 s on the measured boxes) is a fixed cost.  Letting the
            // decode workers run ahead meanwhile (on a helper thread, or simply by starting them first) was measured
            // twice and does not pay: with all cores busy the driver's start-up takes 1.1-1.4 s instead of 0.5 s.
            uint64_t total = 0;
            for (uint64_t L : lengths) total += L;
            int nDevices = 1;
            if (nShards > 1) {
                if (ppp_device_count(&nDevices) != PPP_OK || nDevices < 1) return gpuError("cannot initialise the CUDA context");
                if (ppp_group_create(nShards, &group) != PPP_OK) return gpuError("cannot create the context group");
            }
            for (int k = 0; k < nShards; ++k) {
                Shard& sh = shards[k];
                sh.up.prepare(lengths.size());
                ppp_config cfg;
                memset(&cfg, 0, sizeof cfg);
                cfg.device = nShards > 1 ? (options.device + k) % nDevices : options.device;  // more shards than GPUs: they share
                cfg.n_contigs = (uint32_t)lengths.size();
                cfg.contig_lengths = lengths.data();
                cfg.min_overlap = options.minOverlap;
                cfg.max_overlap = options.maxOverlap;
                cfg.pingpong_overlap = 10;
                cfg.multi_hits = options.countMultiHits;
                if (nShards > 1) {
                    sh.rangeBegin = total * (uint64_t)k / (uint64_t)nShards;
                    sh.rangeEnd = total * (uint64_t)(k + 1) / (uint64_t)nShards;
                    if (sh.rangeEnd == sh.rangeBegin) {
                        std::cerr << "pingpongpro: --gpus " << nShards << " leaves a GPU without a single position of this genome" << std::endl;
                        return 1;
                    }
                    cfg.range_begin = sh.rangeBegin;
                    cfg.range_end = sh.rangeEnd;
                    shardKeyRanges(sh, lengths, options.maxOverlap);
                }
                if (ppp_init(&cfg, &sh.ctx) != PPP_OK) return gpuError("cannot initialise the CUDA context");
                if (group && ppp_group_attach(sh.ctx, group, k) != PPP_OK) return gpuError("cannot attach the context to its group");
                if (!sh.up.attach(sh.ctx)) return gpuError("cannot allocate pinned upload buffers");
            }
            haveContexts = true;
            mark("ppp_init done");
        }
        bool foreignContig = false;
        // inflate / line parsing, record decoding and extraction run on worker threads; this thread appends the kept
        // reads in file order (the order of the float additions of -m weighted, :487) and feeds the uploads
        int nt = options.threads > 0 ? options.threads : (int)std::thread::hardware_concurrency();
        reader.startBatches(nt > 2 ? nt - 2 : 1, extract_record_fn, &eo);
        const uint8_t* bytes;
        size_t nBytes;
        uint64_t nRecords;
        int err;
        while (reader.nextBatch(bytes, nBytes, nRecords, err)) {
            const ExtractedRead* e = reinterpret_cast<const ExtractedRead*>(bytes);
            for (size_t k = 0, m = nBytes / sizeof(ExtractedRead); k < m;) {  // runs of one contig are appended in bulk
                const int32_t contig = e[k].contig;
                size_t j = k;
                for (; j < m && e[j].contig == contig; ++j) totalReadCount += (double)e[j].weight;  // :488, file order
                if ((size_t)contig >= lengths.size()) foreignContig = true;  // header differs: reported below
                else if (nShards == 1) {
                    if (!shards[0].up.addRun((size_t)contig, e + k, j - k)) return gpuError("upload failed");
                } else {
                    for (Shard& sh : shards)
                        if (sh.hi[contig] && !sh.up.addRunRouted((size_t)contig, e + k, j - k, sh.lo[contig], sh.hi[contig])) return gpuError("upload failed");
                }
                k = j;
            }
            if (err) {
                std::cerr << "Failed to read record" << std::endl;  // :411
                return 1;
            }
        }
        if (!reader.good()) {
            std::cerr << "Failed to read record" << std::endl;
            return 1;
        }
        // A contig that the header lacks although it lists others: the reference (SeqAn) appends such names while reading and
        // works without lengths; the dense stack arrays need every length before the first record.  (Input without any @SQ
        // line is scanned first, see above; a partial header is a documented divergence)
        if (foreignContig) {
            std::cerr << "'" << path << "': a record names a contig that has no @SQ header line (contig lengths must be known before the first record)" << std::endl;
            return 1;
        }
        // @SQ lines of all inputs must be identical (:1494-1517)
        if (reader.names() != names) {
            std::cerr << "@SQ header lines of '" << path << "' differ from those of previous input files" << std::endl;  // :1514
            return 1;
        }
        reader.close();
        stopwatch(options.verbosity);
    }
    mark("decode done");
    for (Shard& sh : shards)
        if (!sh.up.flushAll()) return gpuError("upload failed");
    for (Shard& sh : shards)
        if (ppp_stacks_finish(sh.ctx) != PPP_OK) return gpuError("stack build failed");
    mark("stacks finished");

    TransposonsPerGenome transposons;
    if (!options.transposonFiles.empty()) {
        if (options.verbosity >= 3) std::cerr << "Loading transposon coordinates" << std::endl;
        for (const std::string& path : options.transposonFiles) {
            stopwatch(std::string("  ") + path, options.verbosity);
            std::ifstream fs(path.c_str());
            if (fs.fail()) {
                std::cerr << "Failed to open transposon file \"" << path << "\"." << std::endl;  // :1539
                return 1;
            }
            readTransposons(fs, formatFromPath(path), transposons, names);
            stopwatch(options.verbosity);
        }
    }

    if (!options.output.empty()) {  // :1563-1576
        mkdir(options.output.c_str(), 0777);
        if (chdir(options.output.c_str()) != 0) {
            std::cerr << "Failed to open output directory: " << options.output;  // :1573
            return 1;
        }
    }

    stopwatch("Binning stacks", options.verbosity);
    if (!histAndScan(shards, group)) return shardError(shards);
    stopwatch(options.verbosity);

    stopwatch("Calculating FDR for putative ping-pong signatures", options.verbosity);
    uint32_t nBins = 0;
    std::vector<float> cells((size_t)W * PPP_HEIGHT_SCORE_BINS * 4);
    const ppp_locus* loci = nullptr;
    size_t nLoci = 0;
    std::vector<ppp_locus> joined;  // several shards: their loci, concatenated
    if (!scoreAll(shards, options.minStackHeight, &nBins, options.plot ? cells.data() : nullptr, joined, loci, nLoci)) return shardError(shards);
    stopwatch(options.verbosity);

    if (options.plot) {  // generateGroupedStackCountsPlot, :944-981
        stopwatch("Rendering plots for z-scores of ping-pong signatures", options.verbosity);
        std::vector<std::string> titles;
        std::vector<std::vector<float> > hist;
        for (int i = (int)nBins - 1; i >= 0; --i)
            for (unsigned j = 0; j < 2; ++j)
                for (unsigned k = 0; k < 2; ++k) {
                    std::vector<float> h(W);
                    for (int o = 0; o < W; ++o) h[o] = cells[(((size_t)o * nBins + i) * 2 + j) * 2 + k];
                    std::ostringstream ss;
                    ss << "z-scores of signatures with the following properties:" << std::endl;
                    ss << "stack height score of " << (nBins - 1 - i) << std::endl;
                    if (j != 0) ss << "no ";
                    ss << "adenine at position 10" << std::endl;
                    ss << "stack height " << ((k == 0) ? "above" : "below") << " the local coverage";
                    titles.push_back(ss.str());
                    hist.push_back(h);
                }
        plotHistograms("ping-pong_signatures_z-scores", titles, hist, options.minOverlap, options.maxOverlap, 10);
        stopwatch(options.verbosity);
    }

    mark("scan + score done");
    stopwatch("Writing ping-pong signatures to file", options.verbosity);
    writeSignatures(loci, nLoci, names, options.browserTracks);
    stopwatch(options.verbosity);

    if (!options.transposonFiles.empty()) {
        stopwatch("Checking input transposons for ping-pong activity", options.verbosity);
        bool ok = true;
        std::vector<ScoredTransposon> rows = scoreTransposons(shards, transposons, ok);
        if (!ok) return gpuError("transposon scoring failed");
        stopwatch(options.verbosity);
        stopwatch("Writing input transposons to file", options.verbosity);
        writeTransposons(rows, names, options.browserTracks, "transposons", totalReadCount, ppIdx);
        stopwatch(options.verbosity);
        if (options.plot) {
            stopwatch("Rendering plots for z-scores of input transposons", options.verbosity);
            plotTransposons(rows, "transposons_z-scores", options, W);
            stopwatch(options.verbosity);
        }
    }

    if (options.predictTransposonsRange > 0) {
        stopwatch("Predicting transposons based on ping-pong activity", options.verbosity);
        // prediction walks ALL overlap-10 signatures, whatever -s says (:1327)
        if (options.minStackHeight != 0 && !scoreAll(shards, 0, nullptr, nullptr, joined, loci, nLoci)) return shardError(shards);
        std::map<unsigned, std::vector<unsigned> > positions;
        for (size_t i = 0; i < nLoci; ++i) positions[loci[i].contig].push_back(loci[i].position);
        // findSuppressedTransposons has indexed the signature map with every annotated contig (:1203-1206), which
        // leaves empty lists behind; the prediction loop stops at the first of them (:1327)
        if (!options.transposonFiles.empty())
            for (const auto& kv : transposons) positions[kv.first];
        TransposonsPerGenome putative;
        predictTransposons(positions, names, options.predictTransposonsRange, putative);
        bool ok = true;
        std::vector<ScoredTransposon> rows = scoreTransposons(shards, putative, ok);
        if (!ok) return gpuError("transposon scoring failed");
        stopwatch(options.verbosity);
        stopwatch("Writing predicted transposons to file", options.verbosity);
        writeTransposons(rows, names, options.browserTracks, "predicted_transposons", totalReadCount, ppIdx);
        stopwatch(options.verbosity);
        if (options.plot) {
            stopwatch("Rendering plots for z-scores of predicted transposons", options.verbosity);
            plotTransposons(rows, "predicted_transposons_z-scores", options, W);
            stopwatch(options.verbosity);
        }
    }
    mark("outputs written");
    return 0;
}

// Releasing a CUDA context (here: > 1 GB of device memory and the pinned buffers) costs the driver 0.3-0.4 s when the
// process exits -- after every output byte is on disk.  The work therefore runs in a child process, which hands its
// exit status to this tiny parent the moment it is done; the parent returns to the caller at once and the driver's
// teardown of the child proceeds in the background.  (The fork happens before anything touches CUDA.)
// That is OPT-IN (PPP_DETACH_TEARDOWN=1): the child still holds its device memory for a moment after the caller was told
// the run has finished, and a signal aimed at the parent does not reach it (signals are forwarded below, the race with
// a GPU job started right afterwards remains).  By default everything runs in one process.
static volatile pid_t g_child = 0;
static void forwardSignal(int sig) {
    if (g_child > 0) kill(g_child, sig);
}

int main(int argc, char const** argv) {
    int fds[2];
    if (!getenv("PPP_DETACH_TEARDOWN") || getenv("PPP_NO_FORK") || pipe2(fds, O_CLOEXEC) != 0) {
        int rc = run(argc, argv);
        fflush(nullptr);
        _exit(rc);  // every output stream has been closed; skip the CUDA runtime's atexit teardown
    }
    pid_t pid = fork();
    if (pid < 0) {
        close(fds[0]);
        close(fds[1]);
        int rc = run(argc, argv);
        fflush(nullptr);
        _exit(rc);
    }
    if (pid == 0) {
        close(fds[0]);
        int rc = run(argc, argv);
        fflush(nullptr);
        std::cout.flush();
        std::cerr.flush();
        unsigned char status = (unsigned char)rc;
        ssize_t w = write(fds[1], &status, 1);
        (void)w;
        close(fds[1]);
        close(0);
        close(1);  // a caller that captures stdout / stderr must not wait for the background teardown
        close(2);
        _exit(rc);
    }
    close(fds[1]);
    g_child = pid;
    signal(SIGTERM, forwardSignal);
    signal(SIGINT, forwardSignal);
    unsigned char status = 1;
    ssize_t n;
    do n = read(fds[0], &status, 1); while (n < 0 && errno == EINTR);
    if (n == 1) _exit(status);
    int st = 0;  // the child died without reporting: pass on how
    if (waitpid(pid, &st, 0) == pid && WIFEXITED(st)) _exit(WEXITSTATUS(st));
    _exit(WIFSIGNALED(st) ? 128 + WTERMSIG(st) : 1);
}
