// api.cu -- the C ABI of include/ppp_b200.h: context, memory, stream orchestration and the small pieces of
// host arithmetic that must use the HOST libm to stay bit-identical to the reference (log10f-based bin
// thresholds, exp-based Gaussian weights).  No CPU fallback exists anywhere in this file: every entry point
// that computes launches CUDA kernels and fails with PPP_ERR_CUDA when that is impossible.
#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "host_tables.h"
#include "local_group.h"
#include "log10f_emul.h"
#include "fadd_chain.h"
#include "nccl_loader.h"
#include "ppp_b200.h"
#include "transposons.cuh"

using namespace pppk;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess) return fail(PPP_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

constexpr uint32_t kKeySlack = 64;          // keys up to contig length + 63 are accepted (a - key may equal the length, :464)
constexpr size_t kChunk = (size_t)1 << 23;  // records per upload / sort chunk (64 MB)
constexpr int kFoldStreams = 4;
constexpr uint32_t kTallCap = 1u << 12;      // tall stacks (rounded height >= kScanLowBins) per rank in the compact exchange: first size ...
constexpr uint32_t kTallCapMax = 1u << 20;   // ... and the largest it grows to when a pass finds more
constexpr int kStages = 4;                  // depth of the upload -> sort -> fold pipeline (buffers are allocated on first use)
constexpr int kTile = kScanTile;

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    bool used = false;
    const char* what = nullptr;  // label in the PPP_K1_TIMELINE listing
};

// PPP_K1_TIMELINE=1: start and end of every copy, sort and fold of the stack builder on the device, relative to the first
// one, printed by ppp_stacks_finish (how well the three streams overlap)
bool k1Timeline() {
    static const bool on = getenv("PPP_K1_TIMELINE") != nullptr;
    return on;
}

struct PinnedScalars {
    // mirror of the tail of PassState (one copy): totals[2], maxCount, maxHeight, ticket, overflow, pad
    unsigned long long totals[2];
    unsigned long long maxCount;
    uint32_t maxHeight, ticket, tallOverflow, pad;
    // mirror of ScoreState
    unsigned long long nLoci;
    uint32_t scoreTicket, scorePad;
    uint32_t nBins;
    uint32_t rerun;                    // multi-GPU: ranks whose signature store overflowed in this pass (all-reduced)
    uint32_t stackTotals[2];           // stacks per strand (ppp_stacks_finish)
    unsigned long long counters[2];
};
static_assert(offsetof(PassState, maxHeight) - offsetof(PassState, totals) == 24, "PassState tail layout");

}  // namespace

struct ppp_ctx {
    int device = 0;
    int minOv = 3, maxOv = 23, ppOv = 10, W = 21, mode = 0, repr = kReprFloat;
    std::vector<uint64_t> contigLen;
    std::vector<Slot> slotOfContig;  // lenMinus == 0: contig not in this context
    std::vector<Slot> ownedSlots;    // sorted by off
    uint64_t G = 0, ownedPositions = 0;
    uint32_t nTiles = 0, haloLo = 0, haloHi = 0;

    cudaStream_t stream = nullptr, copyStream = nullptr, sortStream = nullptr;
    cudaStream_t foldStream[kFoldStreams] = {};  // -m weighted: folds of different contigs may overlap (see buildChunk)
    uint32_t* H = nullptr;
    uint32_t* occ = nullptr;  // occupancy bitmap of H (2G bits + padding)
    size_t occWords = 0;
    Slot* dSlots = nullptr;
    unsigned long long* dCounters = nullptr;  // [0] foreign, [1] range errors

    // upload double buffers
    // upload ring: kStages device buffers (a long fold must not hold up the copies and sorts behind it), two pinned
    // bounce buffers for callers whose memory is pageable
    ppp_read* dStage[kStages] = {};
    uint8_t* dPacked[kStages] = {};          // ppp_reads_submit_packed: keys (4 B x kChunk), meta (1 B x kChunk), NH table
    uint32_t* hNhTable = nullptr;            // (pinned) one NH table per stage
    ppp_read* hStage[2] = {nullptr, nullptr};
    cudaEvent_t evCopied[kStages] = {}, evConsumed[kStages] = {};
    bool slotUsed[kStages] = {};
    unsigned submitCount = 0;
    uint64_t readsSubmitted = 0;
    std::vector<Timer> buildTimers;
    float buildMs = 0.f;
    // -m weighted: the sorts of the next batches (sortStream) overlap the fold of batch k (stream); one scratch set per stage
    SortScratch sort[kStages] = {};
    cudaEvent_t evSorted[kStages] = {}, evFolded[kStages] = {}, evReset = nullptr, evFork = nullptr, evJoin = nullptr;
    bool sortUsed[kStages] = {};
    unsigned batchCount = 0;

    // compact stack store (compact.cu), rebuilt by ppp_stacks_finish
    uint32_t* dCmp = nullptr;            // plus-strand words, then minus-strand words
    size_t cmpCapacity = 0;
    uint32_t *dTileCntP = nullptr, *dTileCntM = nullptr, *dTileOffP = nullptr, *dTileOffM = nullptr;  // [nTiles + 1] each
    uint64_t nStacksP = 0, nStacksM = 0;

    // scan state
    PassState* dState = nullptr;             // + the look-back states of the tiles + the grouped counts (one allocation, one memset per pass)
    size_t passZeroBytes = 0;
    unsigned long long* dTileState = nullptr;
    uint32_t* dFreqTail = nullptr;
    uint32_t prevMaxHeight = 0;
    // signature records (structure of arrays, (position, overlap) order) and the overlap-`pingpong` stream
    uint32_t* dRecPos = nullptr;
    float *dRecRp = nullptr, *dRecRm = nullptr;
    uint8_t* dRecMisc = nullptr;
    uint16_t* dRecBin = nullptr;
    PairRec* dPP = nullptr;
    LocusOut* dLoci = nullptr;
    ScoreState* dScoreState = nullptr;       // + the look-back states of k_score's tiles
    unsigned long long* dScoreTiles = nullptr;
    size_t scoreTileWords = 0;
    uint64_t pairCapacity = 0, ppCapacity = 0, cfgPairCapacity = 0;
    // array-of-structures view of the records + per-record FDRs, materialised on demand (pairs_download, transposons)
    PairRec* dPairs = nullptr;
    float* dPairFdr = nullptr;
    uint64_t viewCapacity = 0;
    bool viewValid = false, viewHasFdr = false;
    uint32_t* dScanScratch = nullptr;
    size_t scanScratchWords = 0;
    uint32_t* dGrouped = nullptr;            // [W * 1000 * 4] + 1 word: ranks that must re-run the pass
    float *dThresholds = nullptr, *hThresholds = nullptr;
    float *dCells = nullptr, *dFdr = nullptr;
    uint16_t* dBinMap = nullptr;
    uint16_t* dGroupStart = nullptr;
    uint32_t* dNBins = nullptr;
    uint32_t* dBinOcc = nullptr;
    double *dXs = nullptr, *dWs = nullptr;
    int nSteps = 0;
    PinnedScalars* hs = nullptr;
    std::vector<uint64_t> hFreq;
    ppp_locus* hLoci = nullptr;
    size_t hLociCap = 0;
    uint64_t nPairs = 0, nPP = 0;
    uint32_t maxHeight = 0;
    float binScale = 0.f;  // 999 / maxHeightScore: only seeds the device's bin proposal
    bool finished = false, haveHist = false, haveScan = false, haveScore = false;
    bool localOverflow = false;  // this context's signature store was too small in the last scan pass
    // pass bookkeeping (see "the scan + score pass" below)
    bool stageScan = false, stageBin = false, stageScore = false, settled = true;
    uint32_t scoreMinHeight = 0;
    cudaEvent_t evLut = nullptr;           // the pass scalars (counts, maximum frequency) are in host memory
    bool deviceThresholds = true;          // bin thresholds from log10f_emul.h on the device, verified against the host libm
    float* hDevThresholds = nullptr;       // (pinned) the device's table, for that verification
    float hostTableFreq = -1.0f;           // maximum frequency hThresholds was computed for in this pass
    BinParams* dBinParams = nullptr;
    BinParams* hBinParams = nullptr;       // (pinned)

    ppp_allreduce_fn allreduce = nullptr;
    void* allreduceUser = nullptr;
    void* ncclComm = nullptr;  // communicator of the built-in collective (ppp_nccl_attach)
    struct GroupMember* groupMember = nullptr;  // membership of an in-process group (ppp_group_attach)
    int rank = 0, world = 0;   // known (ppp_set_rank / ppp_nccl_attach): tall stacks are exchanged as compact lists
    uint32_t* dGather = nullptr;  // [world][kTallHeader + tallCap], see launch_tall_apply
    uint32_t tallCap = 0;
    bool tallCapFixed = false;  // PPP_TALL_CAP: the list does not grow, an overflow goes to the dense tables
    float thresholdFreq = 0.f;  // the maximum frequency hThresholds was built for (0: none)
    bool denseTail = false;    // a list overflowed once: this context reduces the dense tables from then on

    Timer tScan, tLut, tBin, tScore, tTransposon, tEx1, tEx2, tPass;
    int launches = 0;
};

namespace {

int timerStart(ppp_ctx* c, Timer& t, cudaStream_t s = nullptr) {
    if (!t.a) { CK(cudaEventCreate(&t.a)); CK(cudaEventCreate(&t.b)); }
    CK(cudaEventRecord(t.a, s ? s : c->stream));
    return PPP_OK;
}
int timerStop(ppp_ctx* c, Timer& t, cudaStream_t s = nullptr) {
    CK(cudaEventRecord(t.b, s ? s : c->stream));
    t.used = true;
    return PPP_OK;
}
float timerMs(Timer& t) {
    float ms = 0.f;
    if (t.used && cudaEventElapsedTime(&ms, t.a, t.b) != cudaSuccess) ms = 0.f;
    return ms;
}

int buildLayout(ppp_ctx* c, const ppp_config* cfg) {
    uint64_t total = 0;
    for (uint32_t i = 0; i < cfg->n_contigs; ++i) total += cfg->contig_lengths[i];
    uint64_t rb = cfg->range_begin, re = cfg->range_end;
    if (rb == 0 && re == 0) re = total;
    if (rb > re) return fail(PPP_ERR_ARG, "range_begin > range_end");
    if (re > total) re = total;
    c->slotOfContig.assign(cfg->n_contigs, Slot{0, 0, 0, 0, 0, 0});
    uint64_t off = 0, cum = 0;
    c->ownedPositions = 0;
    c->haloLo = c->haloHi = 0;
    for (uint32_t i = 0; i < cfg->n_contigs; ++i) {
        uint64_t L = cfg->contig_lengths[i];
        uint64_t keySpace = L + kKeySlack;
        if (keySpace + c->maxOv + 8 >= 0xffffffffull) return fail(PPP_ERR_ARG, "contig %u too long", i);
        uint64_t lo = rb > cum ? rb - cum : 0;
        uint64_t hi;
        bool holdsEnd = re >= cum + L;
        if (holdsEnd) hi = keySpace; else hi = re > cum ? re - cum : 0;
        if (L == 0 || lo >= L || hi <= lo) { cum += L; c->slotOfContig[i].keyLimit = (uint32_t)keySpace; c->slotOfContig[i].contig = i; continue; }
        Slot s;
        s.off = (uint32_t)off;
        s.first = (uint32_t)lo;
        s.lenPlus = (uint32_t)(hi - lo);
        uint64_t lm = holdsEnd ? hi - lo : std::min<uint64_t>(hi - lo + c->maxOv, keySpace - lo);
        s.lenMinus = (uint32_t)lm;
        s.keyLimit = (uint32_t)keySpace;
        s.contig = i;
        if (s.lenMinus > s.lenPlus) { c->haloLo = s.off + s.lenPlus; c->haloHi = s.off + s.lenMinus; }
        uint64_t words = std::max<uint64_t>(s.lenMinus, (uint64_t)s.lenPlus + c->maxOv) + 1;
        words = (words + 3) & ~3ull;
        c->slotOfContig[i] = s;
        c->ownedSlots.push_back(s);
        c->ownedPositions += std::min<uint64_t>(hi, L) - lo;
        off += words;
        if (off >= 0xfffff000ull) return fail(PPP_ERR_ARG, "more than 2^32 positions in one context: shard the genome (range_begin/range_end)");
        cum += L;
    }
    c->G = ((off + kTile - 1) / kTile) * kTile;
    if (c->G == 0) c->G = kTile;
    c->nTiles = (uint32_t)(c->G / kTile);
    return PPP_OK;
}

int ensurePairCapacity(ppp_ctx* c, uint64_t want) {
    if (want <= c->pairCapacity && c->dRecPos) return PPP_OK;
    if (want >= kMaxPairCapacity) return fail(PPP_ERR_NOMEM, "signature store would exceed 2^32 records");
    void* old[] = {c->dRecPos, c->dRecRp, c->dRecRm, c->dRecMisc, c->dRecBin, c->dPP, c->dLoci, c->dScoreState};
    for (void* q : old)
        if (q) cudaFree(q);
    c->dRecPos = nullptr;
    c->dRecRp = c->dRecRm = nullptr;
    c->dRecMisc = nullptr;
    c->dRecBin = nullptr;
    c->dPP = nullptr;
    c->dLoci = nullptr;
    c->dScoreState = nullptr;
    c->dScoreTiles = nullptr;
    c->pairCapacity = c->ppCapacity = 0;
    // every record may be an overlap-`pingpong` record (a library made of nothing but 10-nt pairs), up to the field width of
    // the look-back state
    const uint64_t pp = std::min<uint64_t>(want, kMaxLociCapacity);
    CK(cudaMalloc(&c->dRecPos, want * sizeof(uint32_t)));
    CK(cudaMalloc(&c->dRecRp, want * sizeof(float)));
    CK(cudaMalloc(&c->dRecRm, want * sizeof(float)));
    CK(cudaMalloc(&c->dRecMisc, want + 16));
    CK(cudaMalloc(&c->dRecBin, want * sizeof(uint16_t) + 16));
    CK(cudaMalloc(&c->dPP, pp * sizeof(PairRec)));
    CK(cudaMalloc(&c->dLoci, pp * sizeof(LocusOut)));
    c->scoreTileWords = score_tiles(pp);
    CK(cudaMalloc(&c->dScoreState, sizeof(ScoreState) + c->scoreTileWords * sizeof(unsigned long long)));
    c->dScoreTiles = reinterpret_cast<unsigned long long*>(c->dScoreState + 1);
    c->pairCapacity = want;
    c->ppCapacity = pp;
    return PPP_OK;
}

// Array-of-structures view of the records (+ their FDRs once the table exists) for the consumers outside the scan +
// score pass: ppp_pairs_download and the transposon kernels.
int ensurePairView(ppp_ctx* c) {
    const bool wantFdr = c->haveScore;
    if (c->viewValid && (c->viewHasFdr || !wantFdr)) return PPP_OK;
    if (c->viewCapacity < c->nPairs || !c->dPairs) {
        if (c->dPairs) { cudaFree(c->dPairs); cudaFree(c->dPairFdr); }
        c->dPairs = nullptr;
        c->dPairFdr = nullptr;
        c->viewCapacity = 0;
        const uint64_t cap = std::max<uint64_t>(c->nPairs, 1);
        CK(cudaMalloc(&c->dPairs, cap * sizeof(PairRec)));
        CK(cudaMalloc(&c->dPairFdr, cap * sizeof(float)));
        c->viewCapacity = cap;
    }
    if (!wantFdr) CK(cudaMemsetAsync(c->dPairFdr, 0, std::max<uint64_t>(c->nPairs, 1) * sizeof(float), c->stream));
    CK(launch_pairs_view(c->dRecPos, c->dRecRp, c->dRecRm, c->dRecMisc, c->dRecBin, (size_t)c->nPairs, wantFdr ? c->dFdr : nullptr, c->dBinMap, c->W, c->dPairs,
                         c->dPairFdr, c->stream, &c->launches));
    c->viewValid = true;
    c->viewHasFdr = wantFdr;
    return PPP_OK;
}

int ensureSortScratch(ppp_ctx* c, int set) {
    SortScratch& s = c->sort[set];
    if (s.keysA) return PPP_OK;
    s.capacity = kChunk;
    CK(cudaMalloc(&s.keysA, kChunk * 4));
    CK(cudaMalloc(&s.keysB, kChunk * 4));
    CK(cudaMalloc(&s.valsA, kChunk * 4));
    CK(cudaMalloc(&s.valsB, kChunk * 4));
    s.blockHistWords = sort_block_hist_words(kChunk);
    CK(cudaMalloc(&s.blockHist, s.blockHistWords * 4));
    s.scanScratchWords = exclusive_scan_scratch_words(s.blockHistWords);
    CK(cudaMalloc(&s.scanScratch, s.scanScratchWords * 4));
    CK(cudaEventCreateWithFlags(&c->evSorted[set], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->evFolded[set], cudaEventDisableTiming));
    return PPP_OK;
}

// Stack-builder kernels for n device-resident records of one contig (n <= kChunk).  `ready` (may be null): the
// records are valid once it has fired; `consumed` (may be null) is recorded when the kernels no longer read them.
int buildChunk(ppp_ctx* c, const Slot& slot, const ppp_read* dReads, size_t n, cudaEvent_t ready, cudaEvent_t consumed) {
    Timer t;
    int rc;
    if (c->repr == kReprInt) {
        if (ready) CK(cudaStreamWaitEvent(c->stream, ready, 0));
        if ((rc = timerStart(c, t))) return rc;
        CK(launch_stack_build_int(dReads, n, c->H, c->G, c->occ, slot, c->mode, c->dCounters, c->stream, &c->launches));
        if ((rc = timerStop(c, t))) return rc;
        if (consumed) CK(cudaEventRecord(consumed, c->stream));
        c->buildTimers.push_back(t);
        return PPP_OK;
    }
    // The additions of one position must happen in file order: the folds of one contig stay on one stream, in
    // submission order.  Contigs share no positions, so their folds may overlap -- the tail of a fold is a single
    // warp working through the tallest stack.  The sort has no ordering constraint at all and runs ahead on its
    // own stream.
    cudaStream_t fs = c->foldStream[(unsigned)slot.contig % kFoldStreams];
    if (n <= (size_t)kSmallBatch) {  // one launch instead of ~25 (many small contigs, tails of files)
        if (ready) CK(cudaStreamWaitEvent(fs, ready, 0));
        if ((rc = timerStart(c, t, fs))) return rc;
        CK(launch_small_batch(dReads, n, c->H, c->G, c->occ, slot, c->dCounters, fs, &c->launches));
        if ((rc = timerStop(c, t, fs))) return rc;
        c->buildTimers.push_back(t);
        if (consumed) CK(cudaEventRecord(consumed, fs));
        return PPP_OK;
    }
    const int set = (int)(c->batchCount % kStages);
    if ((rc = ensureSortScratch(c, set))) return rc;
    if (ready) CK(cudaStreamWaitEvent(c->sortStream, ready, 0));
    if (c->sortUsed[set]) CK(cudaStreamWaitEvent(c->sortStream, c->evFolded[set], 0));  // scratch set free again
    SortedRun run;
    if ((rc = timerStart(c, t, c->sortStream))) return rc;
    CK(launch_stack_sort(dReads, n, c->G, slot, c->dCounters, c->sort[set], c->sortStream, &c->launches, &run));
    if ((rc = timerStop(c, t, c->sortStream))) return rc;
    t.what = "sort";
    c->buildTimers.push_back(t);
    if (consumed) CK(cudaEventRecord(consumed, c->sortStream));
    CK(cudaEventRecord(c->evSorted[set], c->sortStream));
    CK(cudaStreamWaitEvent(fs, c->evSorted[set], 0));
    Timer tf;
    if ((rc = timerStart(c, tf, fs))) return rc;
    CK(launch_stack_fold(run, c->H, c->G, c->occ, c->sort[set].capacity, fs, &c->launches));
    if ((rc = timerStop(c, tf, fs))) return rc;
    tf.what = "fold";
    c->buildTimers.push_back(tf);
    CK(cudaEventRecord(c->evFolded[set], fs));
    c->sortUsed[set] = true;
    c->batchCount++;
    return PPP_OK;
}

int foldBuildTimers(ppp_ctx* c) {
    if (k1Timeline() && !c->buildTimers.empty()) {
        cudaEvent_t t0 = c->buildTimers[0].a;
        for (Timer& t : c->buildTimers) {
            float a = 0.f, b = 0.f;
            if (!t.used || cudaEventElapsedTime(&a, t0, t.a) != cudaSuccess || cudaEventElapsedTime(&b, t0, t.b) != cudaSuccess) continue;
            fprintf(stderr, "[ppp_k1] %-6s %9.3f .. %9.3f ms  (%.3f)\n", t.what ? t.what : "?", a, b, b - a);
        }
        cudaGetLastError();
    }
    for (Timer& t : c->buildTimers) {
        if (!(t.what && t.what[0] == 'c')) c->buildMs += timerMs(t);  // (the timeline's copy spans are not kernels)
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    c->buildTimers.clear();
    return PPP_OK;
}

int hookReduce(ppp_ctx* c, void* buf, size_t count, int dtype, int op) {
    if (!c->allreduce || count == 0) return PPP_OK;
    int rc = c->allreduce(c->allreduceUser, buf, count, dtype, op, (void*)c->stream);
    if (rc != 0) return fail(PPP_ERR_CUDA, "all-reduce hook failed with code %d", rc);
    return PPP_OK;
}

// PPP_TRACE=1: host wall time between the phases of one call, to stderr (where a multi-GPU step waits); `at` is the time
// since the first traced call of the process, so the gaps BETWEEN calls (the caller's own time) can be read off as well
struct Trace {
    bool on;
    const char* call;
    std::chrono::steady_clock::time_point t;
    static std::chrono::steady_clock::time_point epoch() {
        static const std::chrono::steady_clock::time_point e = std::chrono::steady_clock::now();
        return e;
    }
    explicit Trace(const char* name) : on(getenv("PPP_TRACE") != nullptr), call(name), t(std::chrono::steady_clock::now()) {
        if (on) (void)epoch();
    }
    void mark(const char* what) {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[ppp_trace %s] %-28s +%.3f ms  (at %.3f)\n", call, what, std::chrono::duration<double, std::milli>(now - t).count(),
                std::chrono::duration<double, std::milli>(now - epoch()).count());
        t = now;
    }
};

void invalidateResults(ppp_ctx* c) {
    c->haveHist = c->haveScan = c->haveScore = c->viewValid = false;
    c->stageScan = c->stageBin = c->stageScore = false;
    c->settled = true;
}

// ---- built-in collective: NCCL, bound at run time (nccl_loader.cpp) --------------------------------
int ncclFail(const char* what, int rc) { return fail(PPP_ERR_CUDA, "%s: %s", what, nccl().getErrorString(rc)); }

// ppp_allreduce_fn over the context's communicator; dtype / op codes of ppp_b200.h -> ncclUint32 (3) / ncclUint64 (5), ncclSum (0) / ncclMax (2)
int ncclHook(void* user, void* buf, size_t count, int dtype, int op, void* stream) {
    ppp_ctx* c = static_cast<ppp_ctx*>(user);
    return nccl().allReduce(buf, buf, count, dtype == 1 ? 5 : 3, op == 1 ? 2 : 0, c->ncclComm, static_cast<cudaStream_t>(stream));
}

}  // namespace

// ---- built-in collective for contexts of ONE process: all-reduce through host memory (local_group.cpp) ----
struct ppp_group {
    LocalGroup g;
    explicit ppp_group(int n) : g(n) {}
};
struct GroupMember {
    ppp_group* group;
    int rank;
};

namespace {

int groupHook(void* user, void* buf, size_t count, int dtype, int op, void* stream) {
    GroupMember* m = static_cast<GroupMember*>(user);
    return m->group->g.allReduce(m->rank, buf, count, dtype, op, static_cast<cudaStream_t>(stream));
}

}  // namespace

extern "C" {

const char* ppp_last_error(void) { return g_err.c_str(); }

int ppp_device_count(int* n) {
    int k = 0;
    cudaError_t e = cudaGetDeviceCount(&k);
    if (e != cudaSuccess) { *n = 0; return fail(PPP_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *n = k;
    return PPP_OK;
}

int ppp_warmup(int32_t device) {
    CK(cudaSetDevice(device));
    CK(cudaFree(nullptr));
    return PPP_OK;
}

int ppp_pinned_alloc(size_t bytes, void** ptr) {
    CK(cudaMallocHost(ptr, bytes));
    return PPP_OK;
}
int ppp_pinned_free(void* ptr) {
    CK(cudaFreeHost(ptr));
    return PPP_OK;
}

static void initMark(const char* what) {  // PPP_TIMING=1: where the start-up time goes
    static const bool on = getenv("PPP_TIMING") != nullptr;
    static const auto t0 = std::chrono::steady_clock::now();
    if (on) fprintf(stderr, "[ppp_init %7.3f s] %s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), what);
}

int ppp_init(const ppp_config* cfg, ppp_ctx** out) {
    if (!cfg || !out) return fail(PPP_ERR_ARG, "null argument");
    initMark("enter");
    *out = nullptr;
    if (cfg->n_contigs == 0 || !cfg->contig_lengths) return fail(PPP_ERR_ARG, "no contigs");
    if (cfg->min_overlap < 0 || cfg->max_overlap < cfg->min_overlap || cfg->max_overlap > 32 || cfg->pingpong_overlap < cfg->min_overlap ||
        cfg->pingpong_overlap > cfg->max_overlap || cfg->max_overlap - cfg->min_overlap + 1 > PPP_MAX_OVERLAPS || cfg->max_overlap - cfg->min_overlap + 1 < 3)
        return fail(PPP_ERR_ARG, "overlaps must satisfy 0 <= min <= pingpong <= max <= 32 with at least 3 overlaps");
    if (cfg->multi_hits < 0 || cfg->multi_hits > 2) return fail(PPP_ERR_ARG, "multi_hits must be PPP_MULTIHITS_*");
    int nDev = 0;
    cudaError_t e = cudaGetDeviceCount(&nDev);
    if (e != cudaSuccess || nDev == 0)
        return fail(PPP_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= nDev) return fail(PPP_ERR_ARG, "device %d out of range (%d devices)", cfg->device, nDev);
    initMark("cudaGetDeviceCount");
    CK(cudaSetDevice(cfg->device));
    CK(cudaFree(nullptr));
    initMark("context created");
    ppp_ctx* c = new ppp_ctx;
    c->device = cfg->device;
    c->minOv = cfg->min_overlap;
    c->maxOv = cfg->max_overlap;
    c->ppOv = cfg->pingpong_overlap;
    c->W = c->maxOv - c->minOv + 1;
    c->mode = cfg->multi_hits;
    c->repr = cfg->multi_hits == PPP_MULTIHITS_WEIGHTED ? kReprFloat : kReprInt;
    c->cfgPairCapacity = cfg->pair_capacity;
    c->contigLen.assign(cfg->contig_lengths, cfg->contig_lengths + cfg->n_contigs);
    int rc = buildLayout(c, cfg);
    if (rc) { delete c; return rc; }
#define CKI(expr)                                                                                                     \
    do {                                                                                                              \
        cudaError_t e__ = (expr);                                                                                     \
        if (e__ != cudaSuccess) { int r__ = fail(PPP_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); ppp_destroy(c); return r__; } \
    } while (0)
    CKI(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CKI(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
    CKI(cudaStreamCreateWithFlags(&c->sortStream, cudaStreamNonBlocking));
    for (int k = 0; k < kFoldStreams; ++k) CKI(cudaStreamCreateWithFlags(&c->foldStream[k], cudaStreamNonBlocking));
    CKI(cudaEventCreateWithFlags(&c->evReset, cudaEventDisableTiming));
    CKI(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
    CKI(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
    size_t hWords = 2 * c->G + 64;
    CKI(cudaMalloc(&c->H, hWords * sizeof(uint32_t)));
    CKI(cudaMemsetAsync(c->H, 0, hWords * sizeof(uint32_t), c->stream));
    c->occWords = (size_t)(2 * c->G / 32) + 64;
    CKI(cudaMalloc(&c->occ, c->occWords * sizeof(uint32_t)));
    CKI(cudaMemsetAsync(c->occ, 0, c->occWords * sizeof(uint32_t), c->stream));
    initMark("stack arrays allocated");
    if (!c->ownedSlots.empty()) {
        CKI(cudaMalloc(&c->dSlots, c->ownedSlots.size() * sizeof(Slot)));
        CKI(cudaMemcpyAsync(c->dSlots, c->ownedSlots.data(), c->ownedSlots.size() * sizeof(Slot), cudaMemcpyHostToDevice, c->stream));
    }
    CKI(cudaMalloc(&c->dCounters, 2 * sizeof(unsigned long long)));
    CKI(cudaMemsetAsync(c->dCounters, 0, 2 * sizeof(unsigned long long), c->stream));
    for (int k = 0; k < kStages; ++k) {
        CKI(cudaEventCreateWithFlags(&c->evCopied[k], cudaEventDisableTiming));
        CKI(cudaEventCreateWithFlags(&c->evConsumed[k], cudaEventDisableTiming));
    }
    {   // everything a pass zeroes at its start, in one allocation: pass scalars, look-back states, grouped counts (+ rerun word)
        const size_t groupedBytes = (((size_t)c->W * kBins * 4 + 1) * sizeof(uint32_t) + 15) & ~(size_t)15;
        c->passZeroBytes = sizeof(PassState) + (size_t)c->nTiles * sizeof(unsigned long long) + groupedBytes;
        CKI(cudaMalloc(&c->dState, c->passZeroBytes));
        c->dTileState = reinterpret_cast<unsigned long long*>(c->dState + 1);
        c->dGrouped = reinterpret_cast<uint32_t*>(c->dTileState + c->nTiles);
    }
    CKI(cudaMalloc(&c->dFreqTail, (size_t)kFreqCapacity * sizeof(uint32_t)));
    CKI(cudaMemsetAsync(c->dFreqTail, 0, (size_t)kFreqCapacity * sizeof(uint32_t), c->stream));
    CKI(cudaMalloc(&c->dTileCntP, ((size_t)c->nTiles + 1) * 4));
    CKI(cudaMalloc(&c->dTileCntM, ((size_t)c->nTiles + 1) * 4));
    CKI(cudaMalloc(&c->dTileOffP, ((size_t)c->nTiles + 1) * 4));
    CKI(cudaMalloc(&c->dTileOffM, ((size_t)c->nTiles + 1) * 4));
    c->scanScratchWords = exclusive_scan_scratch_words((size_t)c->nTiles + 1);
    CKI(cudaMalloc(&c->dScanScratch, c->scanScratchWords * 4));
    size_t tableWords = (size_t)c->W * kBins * 4;
    CKI(cudaMalloc(&c->dCells, tableWords * 4));
    CKI(cudaMalloc(&c->dFdr, tableWords * 4));
    CKI(cudaMalloc(&c->dBinMap, kBins * sizeof(uint16_t)));
    CKI(cudaMalloc(&c->dGroupStart, (kBins + 1) * sizeof(uint16_t)));
    CKI(cudaMalloc(&c->dNBins, sizeof(uint32_t)));
    CKI(cudaMalloc(&c->dBinOcc, collapse_occ_words() * sizeof(uint32_t)));
    CKI(cudaMalloc(&c->dThresholds, kBins * sizeof(float)));
    initMark("device tables allocated");
    CKI(cudaMallocHost(&c->hThresholds, kBins * sizeof(float)));
    CKI(cudaMallocHost(&c->hDevThresholds, kBins * sizeof(float)));
    CKI(cudaMallocHost(&c->hBinParams, sizeof(BinParams)));
    CKI(cudaMalloc(&c->dBinParams, sizeof(BinParams)));
    CKI(cudaMemsetAsync(c->dBinParams, 0, sizeof(BinParams), c->stream));
    CKI(cudaEventCreateWithFlags(&c->evLut, cudaEventDisableTiming));
    if (getenv("PPP_HOST_THRESHOLDS")) c->deviceThresholds = false;  // test knob: the fall-back of a host whose libm differs
    CKI(cudaMallocHost(&c->hs, sizeof(PinnedScalars)));
    memset(c->hs, 0, sizeof(PinnedScalars));
    std::vector<double> xs, ws;
    gaussianTable(xs, ws);
    c->nSteps = (int)xs.size();
    CKI(cudaMalloc(&c->dXs, xs.size() * sizeof(double)));
    CKI(cudaMalloc(&c->dWs, ws.size() * sizeof(double)));
    CKI(cudaMemcpyAsync(c->dXs, xs.data(), xs.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CKI(cudaMemcpyAsync(c->dWs, ws.data(), ws.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    initMark("pinned + tables");
    CKI(cudaStreamSynchronize(c->stream));
    initMark("synchronised");
#undef CKI
    *out = c;
    return PPP_OK;
}

void ppp_destroy(ppp_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copyStream) cudaStreamSynchronize(c->copyStream);
    if (c->sortStream) cudaStreamSynchronize(c->sortStream);
    for (int k = 0; k < kFoldStreams; ++k)
        if (c->foldStream[k]) cudaStreamSynchronize(c->foldStream[k]);
    foldBuildTimers(c);
    if (c->ncclComm) nccl().commDestroy(c->ncclComm);
    delete c->groupMember;
    if (c->dGather) cudaFree(c->dGather);
    void* dev[] = {c->H, c->occ, c->dSlots, c->dCounters, c->dCmp, c->dTileCntP, c->dTileCntM, c->dTileOffP, c->dTileOffM, c->dState, c->dFreqTail,
                   c->dRecPos, c->dRecRp, c->dRecRm, c->dRecMisc, c->dRecBin, c->dPP, c->dLoci, c->dScoreState, c->dPairs, c->dPairFdr, c->dScanScratch,
                   c->dThresholds, c->dCells, c->dFdr, c->dBinMap, c->dNBins, c->dXs, c->dWs, c->dGroupStart, c->dBinParams, c->dBinOcc};
    for (void* p : dev)
        if (p) cudaFree(p);
    for (int k = 0; k < kStages; ++k) {
        void* stage[] = {c->dStage[k], c->dPacked[k], c->sort[k].keysA, c->sort[k].keysB, c->sort[k].valsA, c->sort[k].valsB, c->sort[k].blockHist, c->sort[k].scanScratch};
        for (void* p : stage)
            if (p) cudaFree(p);
    }
    void* pinned[] = {c->hStage[0], c->hStage[1], c->hNhTable, c->hThresholds, c->hDevThresholds, c->hBinParams, c->hs, c->hLoci};
    for (void* p : pinned)
        if (p) cudaFreeHost(p);
    for (int k = 0; k < kStages; ++k) {
        if (c->evCopied[k]) cudaEventDestroy(c->evCopied[k]);
        if (c->evConsumed[k]) cudaEventDestroy(c->evConsumed[k]);
        if (c->evSorted[k]) cudaEventDestroy(c->evSorted[k]);
        if (c->evFolded[k]) cudaEventDestroy(c->evFolded[k]);
    }
    for (Timer* t : {&c->tScan, &c->tLut, &c->tBin, &c->tScore, &c->tTransposon})
        if (t->a) { cudaEventDestroy(t->a); cudaEventDestroy(t->b); }
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copyStream) cudaStreamDestroy(c->copyStream);
    if (c->sortStream) cudaStreamDestroy(c->sortStream);
    for (int k = 0; k < kFoldStreams; ++k)
        if (c->foldStream[k]) cudaStreamDestroy(c->foldStream[k]);
    if (c->evLut) cudaEventDestroy(c->evLut);
    if (c->evReset) cudaEventDestroy(c->evReset);
    if (c->evFork) cudaEventDestroy(c->evFork);
    if (c->evJoin) cudaEventDestroy(c->evJoin);
    delete c;
}

int ppp_reset(ppp_ctx* c) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->copyStream));
    for (int k = 0; k < kFoldStreams; ++k) CK(cudaStreamSynchronize(c->foldStream[k]));  // folds still in flight write H
    if (c->finished && ((uint64_t)c->nStacksP + c->nStacksM) * 8 < c->G) {
        // a sparse store (the usual case: a few per cent of the positions hold a stack): only the occupied sectors
        CK(launch_sparse_clear(c->H, c->occ, (size_t)(2 * c->G / 32), c->stream, &c->launches));
    } else {
        CK(cudaMemsetAsync(c->H, 0, (2 * c->G + 64) * sizeof(uint32_t), c->stream));
        CK(cudaMemsetAsync(c->occ, 0, c->occWords * sizeof(uint32_t), c->stream));
    }
    CK(cudaMemsetAsync(c->dCounters, 0, 2 * sizeof(unsigned long long), c->stream));
    CK(cudaEventRecord(c->evReset, c->stream));  // the sort stream's kernels count into dCounters
    CK(cudaStreamWaitEvent(c->sortStream, c->evReset, 0));
    for (int k = 0; k < kFoldStreams; ++k) CK(cudaStreamWaitEvent(c->foldStream[k], c->evReset, 0));
    c->readsSubmitted = 0;
    c->finished = false;
    invalidateResults(c);
    return PPP_OK;
}

int ppp_set_allreduce(ppp_ctx* c, ppp_allreduce_fn fn, void* user) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    c->allreduce = fn;
    c->allreduceUser = user;
    return PPP_OK;
}

int ppp_set_rank(ppp_ctx* c, int rank, int world) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    if (world < 1 || rank < 0 || rank >= world) return fail(PPP_ERR_ARG, "rank %d of %d", rank, world);
    if (c->dGather) return fail(PPP_ERR_STATE, "the rank cannot change after the first scan");
    c->rank = rank;
    c->world = world;
    c->tallCap = kTallCap;
    if (const char* e = getenv("PPP_TALL_CAP")) { c->tallCap = (uint32_t)std::max(1L, atol(e)); c->tallCapFixed = true; }  // test knob: forces the dense fall-back
    if (const char* e = getenv("PPP_TALL_CAP_START")) c->tallCap = (uint32_t)std::max(1L, atol(e));  // test knob: a list that has to grow
    c->tallCap = (c->tallCap + 1u) & ~1u;  // slots stay 8-byte aligned (64-bit low bins inside)
    return PPP_OK;
}

int ppp_nccl_unique_id(void* id) {
    if (!id) return fail(PPP_ERR_ARG, "null argument");
    if (!nccl().ok) return fail(PPP_ERR_STATE, "%s", nccl().why.c_str());
    int rc = nccl().getUniqueId(static_cast<NcclId*>(id));
    return rc ? ncclFail("ncclGetUniqueId", rc) : PPP_OK;
}

int ppp_nccl_attach(ppp_ctx* c, const void* id, int rank, int world) {
    if (!c || !id) return fail(PPP_ERR_ARG, "null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(PPP_ERR_ARG, "rank %d of %d", rank, world);
    if (c->ncclComm) return fail(PPP_ERR_STATE, "a communicator is already attached");
    if (!nccl().ok) return fail(PPP_ERR_STATE, "%s", nccl().why.c_str());
    CK(cudaSetDevice(c->device));
    NcclId nid;
    memcpy(&nid, id, sizeof nid);
    int rc = nccl().commInitRank(&c->ncclComm, world, nid, rank);
    if (rc) { c->ncclComm = nullptr; return ncclFail("ncclCommInitRank", rc); }
    c->allreduce = ncclHook;
    c->allreduceUser = c;
    return ppp_set_rank(c, rank, world);
}

int ppp_group_create(int n_ranks, ppp_group** out) {
    if (!out || n_ranks < 1) return fail(PPP_ERR_ARG, "ppp_group_create: %d ranks", n_ranks);
    *out = new ppp_group(n_ranks);
    return PPP_OK;
}

int ppp_group_attach(ppp_ctx* c, ppp_group* g, int rank) {
    if (!c || !g) return fail(PPP_ERR_ARG, "null argument");
    if (rank < 0 || rank >= g->g.size()) return fail(PPP_ERR_ARG, "rank %d of %d", rank, g->g.size());
    if (c->groupMember || c->ncclComm) return fail(PPP_ERR_STATE, "a collective is already attached");
    c->groupMember = new GroupMember{g, rank};
    c->allreduce = groupHook;
    c->allreduceUser = c->groupMember;
    return ppp_set_rank(c, rank, g->g.size());
}

void ppp_group_abort(ppp_group* g) {
    if (g) g->g.abort();
}

void ppp_group_destroy(ppp_group* g) { delete g; }

int ppp_reads_submit_device(ppp_ctx* c, int32_t contig, const ppp_read* dReads, size_t n) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    if (contig < 0 || (size_t)contig >= c->slotOfContig.size()) return fail(PPP_ERR_ARG, "contig %d out of range", contig);
    CK(cudaSetDevice(c->device));
    invalidateResults(c);
    c->finished = false;
    const Slot& slot = c->slotOfContig[contig];
    for (size_t o = 0; o < n; o += kChunk) {
        size_t k = std::min(kChunk, n - o);
        int rc = buildChunk(c, slot, dReads + o, k, nullptr, nullptr);
        if (rc) return rc;
    }
    c->readsSubmitted += n;
    return PPP_OK;
}

int ppp_reads_submit(ppp_ctx* c, int32_t contig, const ppp_read* reads, size_t n, uint64_t /*first_record_index*/) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    if (contig < 0 || (size_t)contig >= c->slotOfContig.size()) return fail(PPP_ERR_ARG, "contig %d out of range", contig);
    if (n && !reads) return fail(PPP_ERR_ARG, "null reads");
    CK(cudaSetDevice(c->device));
    invalidateResults(c);
    c->finished = false;
    const Slot& slot = c->slotOfContig[contig];
    cudaPointerAttributes attr;
    bool pinned = false;
    if (n && cudaPointerGetAttributes(&attr, reads) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    for (size_t o = 0; o < n; o += kChunk) {
        size_t k = std::min(kChunk, n - o);
        const int s = (int)(c->submitCount % kStages), prev = (int)((c->submitCount + kStages - 1) % kStages);
        // the double-buffer contract: the copy of the previous submission has left its host buffer (copies run in
        // order on one stream, so every earlier one has, too)
        if (c->slotUsed[prev]) CK(cudaEventSynchronize(c->evCopied[prev]));
        if (!c->dStage[s]) CK(cudaMalloc(&c->dStage[s], kChunk * sizeof(ppp_read)));
        const ppp_read* src = reads + o;
        if (!pinned) {
            const int hs = (int)(c->submitCount & 1u);  // last used two submissions ago: that copy is complete (see above)
            if (!c->hStage[hs]) CK(cudaMallocHost(&c->hStage[hs], kChunk * sizeof(ppp_read)));
            memcpy(c->hStage[hs], src, k * sizeof(ppp_read));
            src = c->hStage[hs];
        }
        if (c->slotUsed[s]) CK(cudaStreamWaitEvent(c->copyStream, c->evConsumed[s], 0));  // kernels done with dStage[s]
        CK(cudaMemcpyAsync(c->dStage[s], src, k * sizeof(ppp_read), cudaMemcpyHostToDevice, c->copyStream));
        CK(cudaEventRecord(c->evCopied[s], c->copyStream));
        int rc = buildChunk(c, slot, c->dStage[s], k, c->evCopied[s], c->evConsumed[s]);
        if (rc) return rc;
        c->slotUsed[s] = true;
        c->submitCount++;
    }
    c->readsSubmitted += n;
    return PPP_OK;
}

int ppp_reads_submit_packed(ppp_ctx* c, int32_t contig, const uint32_t* keys, const uint8_t* meta, const uint32_t* nhTable, size_t n) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    if (contig < 0 || (size_t)contig >= c->slotOfContig.size()) return fail(PPP_ERR_ARG, "contig %d out of range", contig);
    if (n && (!keys || !meta || !nhTable)) return fail(PPP_ERR_ARG, "null reads");
    CK(cudaSetDevice(c->device));
    invalidateResults(c);
    c->finished = false;
    const Slot& slot = c->slotOfContig[contig];
    if (!c->hNhTable) CK(cudaMallocHost(&c->hNhTable, (size_t)kStages * PPP_PACKED_NH_CODES * sizeof(uint32_t)));
    for (size_t o = 0; o < n; o += kChunk) {
        size_t k = std::min(kChunk, n - o);
        const int s = (int)(c->submitCount % kStages), prev = (int)((c->submitCount + kStages - 1) % kStages);
        if (c->slotUsed[prev]) CK(cudaEventSynchronize(c->evCopied[prev]));  // the double-buffer contract of ppp_reads_submit
        if (c->slotUsed[s]) CK(cudaEventSynchronize(c->evCopied[s]));        // this stage's NH table slot is free again
        if (!c->dStage[s]) CK(cudaMalloc(&c->dStage[s], kChunk * sizeof(ppp_read)));
        if (!c->dPacked[s]) CK(cudaMalloc(&c->dPacked[s], kChunk * 5 + PPP_PACKED_NH_CODES * sizeof(uint32_t)));
        uint32_t* dKeys = reinterpret_cast<uint32_t*>(c->dPacked[s]);
        uint32_t* dTable = reinterpret_cast<uint32_t*>(c->dPacked[s] + kChunk * 4);
        uint8_t* dMeta = c->dPacked[s] + kChunk * 4 + PPP_PACKED_NH_CODES * sizeof(uint32_t);
        uint32_t* hTable = c->hNhTable + (size_t)s * PPP_PACKED_NH_CODES;
        memcpy(hTable, nhTable, PPP_PACKED_NH_CODES * sizeof(uint32_t));
        if (c->slotUsed[s]) CK(cudaStreamWaitEvent(c->copyStream, c->evConsumed[s], 0));  // kernels done with dStage[s]
        Timer tc;
        if (k1Timeline()) { int trc = timerStart(c, tc, c->copyStream); if (trc) return trc; }
        CK(cudaMemcpyAsync(dKeys, keys + o, k * sizeof(uint32_t), cudaMemcpyHostToDevice, c->copyStream));
        CK(cudaMemcpyAsync(dMeta, meta + o, k, cudaMemcpyHostToDevice, c->copyStream));
        CK(cudaMemcpyAsync(dTable, hTable, PPP_PACKED_NH_CODES * sizeof(uint32_t), cudaMemcpyHostToDevice, c->copyStream));
        CK(launch_unpack_reads(dKeys, dMeta, dTable, k, c->dStage[s], c->copyStream, &c->launches));
        if (k1Timeline()) { int trc = timerStop(c, tc, c->copyStream); if (trc) return trc; tc.what = "copy"; c->buildTimers.push_back(tc); }
        CK(cudaEventRecord(c->evCopied[s], c->copyStream));
        int rc = buildChunk(c, slot, c->dStage[s], k, c->evCopied[s], c->evConsumed[s]);
        if (rc) return rc;
        c->slotUsed[s] = true;
        c->submitCount++;
    }
    c->readsSubmitted += n;
    return PPP_OK;
}

int ppp_stacks_finish(ppp_ctx* c) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->copyStream));
    CK(cudaStreamSynchronize(c->sortStream));
    for (int k = 0; k < kFoldStreams; ++k) CK(cudaStreamSynchronize(c->foldStream[k]));
    // the stacks are final: write them down as the compact, position-ordered store the scan reads (compact.cu)
    Timer t;
    int rc;
    if ((rc = timerStart(c, t))) return rc;
    CK(launch_tile_popc(c->occ, c->G, c->nTiles, c->dTileCntP, c->dTileCntM, c->stream, &c->launches));
    CK(exclusive_scan_u32(c->dTileCntP, c->dTileOffP, (size_t)c->nTiles + 1, c->dScanScratch, c->scanScratchWords, c->stream, &c->launches));
    CK(exclusive_scan_u32(c->dTileCntM, c->dTileOffM, (size_t)c->nTiles + 1, c->dScanScratch, c->scanScratchWords, c->stream, &c->launches));
    CK(cudaMemcpyAsync(&c->hs->stackTotals[0], c->dTileOffP + c->nTiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(&c->hs->stackTotals[1], c->dTileOffM + c->nTiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(c->hs->counters, c->dCounters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->nStacksP = c->hs->stackTotals[0];
    c->nStacksM = c->hs->stackTotals[1];
    const size_t need = (size_t)c->nStacksP + (size_t)c->nStacksM + 64;
    if (need > c->cmpCapacity) {
        if (c->dCmp) cudaFree(c->dCmp);
        c->dCmp = nullptr;
        c->cmpCapacity = 0;
        const size_t cap = need + need / 8;
        CK(cudaMalloc(&c->dCmp, cap * sizeof(uint32_t)));
        c->cmpCapacity = cap;
    }
    CK(launch_compact_write(c->H, c->occ, c->G, c->nTiles, c->dTileOffP, c->dTileOffM, c->dCmp, c->dCmp + c->nStacksP, c->stream, &c->launches));
    if ((rc = timerStop(c, t))) return rc;
    t.what = "finish";
    c->buildTimers.push_back(t);
    CK(cudaStreamSynchronize(c->stream));
    foldBuildTimers(c);
    c->finished = true;
    if (c->hs->counters[1]) return fail(PPP_ERR_RANGE, "%llu read(s) lie beyond the end of their contig", c->hs->counters[1]);
    return PPP_OK;
}

int ppp_reads_foreign(ppp_ctx* c, uint64_t* n) {
    if (!c || !n) return fail(PPP_ERR_ARG, "null argument");
    if (!c->finished) return fail(PPP_ERR_STATE, "call ppp_stacks_finish first");
    *n = c->hs->counters[0];
    return PPP_OK;
}

int ppp_stacks_download(ppp_ctx* c, int32_t contig, int32_t strand, uint64_t begin, uint64_t count, float* reads, uint8_t* a10) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    if (contig < 0 || (size_t)contig >= c->slotOfContig.size() || strand < 0 || strand > 1) return fail(PPP_ERR_ARG, "bad contig/strand");
    CK(cudaSetDevice(c->device));
    const Slot& s = c->slotOfContig[contig];
    uint64_t len = strand ? s.lenMinus : s.lenPlus;
    if (begin < s.first || begin - s.first + count > len) return fail(PPP_ERR_ARG, "key range outside this context");
    std::vector<uint32_t> w(count);
    const uint32_t* src = c->H + (strand ? c->G : 0) + s.off + (begin - s.first);
    CK(cudaMemcpyAsync(w.data(), src, count * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (uint64_t i = 0; i < count; ++i) {
        uint32_t v = w[i] & kValueMask;
        float f;
        if (c->repr == kReprInt) f = (float)(v < kSaturate ? v : kSaturate);
        else memcpy(&f, &v, 4);
        if (reads) reads[i] = f;
        if (a10) a10[i] = (uint8_t)(w[i] >> 31);
    }
    return PPP_OK;
}

}  // extern "C"

namespace {

// ---- the scan + score pass ---------------------------------------------------------------------------------------
// A pass has three stages -- scan (k_scan, histogram exchange, frequency LUT, bin thresholds), bin (k_bin, exchange of the
// grouped counts) and score (collapse, FDR table, k_score) -- which ppp_height_hist / ppp_scan / ppp_score ENQUEUE.  The host
// looks at the pass (`settle`) only when a call has to return something: a caller that asks for nothing before ppp_score
// gets the whole pass as one stream of kernels with a single synchronisation at its end.  What the host has to decide --
// did the signature store overflow (on any rank), did the tall-stack lists overflow, does the device's threshold table
// equal the one the host libm gives -- is decided then, and when the answer is "no good" the cause is removed and every
// stage that had been enqueued is enqueued again.  All of these decisions are functions of all-reduced data (or of this
// host's libm), so the ranks of a sharded run take them together.

int enqueueScanStage(ppp_ctx* c) {
    int rc;
    Trace tr("scan stage");
    tr.mark("begin");
    if ((rc = timerStart(c, c->tPass))) return rc;
    c->tPass.used = false;
    CK(cudaMemsetAsync(c->dState, 0, c->passZeroBytes, c->stream));  // pass scalars, look-back states, grouped counts
    if (c->prevMaxHeight >= kScanLowBins)
        CK(cudaMemsetAsync(c->dFreqTail + kScanLowBins, 0, ((size_t)c->prevMaxHeight - kScanLowBins + 1) * sizeof(uint32_t), c->stream));
    ScanParams p;
    p.occ = c->occ;
    p.cmpP = c->dCmp;
    p.cmpM = c->dCmp + c->nStacksP;
    p.tileOffP = c->dTileOffP;
    p.tileOffM = c->dTileOffM;
    p.G = c->G;
    p.nTiles = c->nTiles;
    p.minOv = c->minOv;
    p.W = c->W;
    p.ppIdx = c->ppOv - c->minOv;
    p.haloLo = c->haloLo;
    p.haloHi = c->haloHi;
    p.state = c->dState;
    p.tileState = c->dTileState;
    p.freqLow = c->dState->freqLow;
    p.freqTail = c->dFreqTail;
    p.maxHeight = &c->dState->maxHeight;
    p.recPos = c->dRecPos;
    p.recRp = c->dRecRp;
    p.recRm = c->dRecRm;
    p.recMisc = c->dRecMisc;
    p.capacity = c->pairCapacity;
    p.ppRecs = c->dPP;
    p.ppCapacity = c->ppCapacity;
    // multi-GPU with known rank: heights beyond the shared-memory bins go to this rank's slot of the gather buffer
    const bool gatherTall = c->allreduce && c->world > 1 && !c->denseTail;
    const size_t slotWords = (size_t)kTallHeader + c->tallCap;
    p.tallList = nullptr;
    p.tallCount = nullptr;
    p.tallCap = 0;
    p.maxTail = &c->dState->maxTail;
    if (gatherTall) {
        if (!c->dGather) CK(cudaMalloc(&c->dGather, (size_t)c->world * slotWords * sizeof(uint32_t)));
        CK(cudaMemsetAsync(c->dGather, 0, (size_t)c->world * slotWords * sizeof(uint32_t), c->stream));
        uint32_t* slot = c->dGather + (size_t)c->rank * slotWords;
        p.tallCount = slot;
        p.maxHeight = slot + 1;
        p.freqLow = reinterpret_cast<unsigned long long*>(slot + kTallLow);
        p.tallList = slot + kTallHeader;
        p.tallCap = c->tallCap;
    }
    if ((rc = timerStart(c, c->tScan))) return rc;
    CK(launch_scan(p, c->repr, c->stream, &c->launches));
    if ((rc = timerStop(c, c->tScan))) return rc;
    if (c->allreduce && (rc = timerStart(c, c->tEx1))) return rc;
    if (gatherTall) {
        // one all-reduce carries every rank's low bins, tall stacks and maximum height; no host round trip in between
        if ((rc = hookReduce(c, c->dGather, (size_t)c->world * slotWords, 0, 0))) return rc;
        CK(launch_tall_apply(c->dGather, c->world, c->tallCap, c->dState->freqLow, c->dFreqTail, &c->dState->maxHeight, &c->dState->overflow, &c->dState->maxTail,
                             c->stream, &c->launches));
    } else if (c->allreduce) {
        // sizes of the table reductions depend on the global maximum height: one extra round trip
        if ((rc = hookReduce(c, &c->dState->maxHeight, 1, 0, 1))) return rc;
        CK(cudaMemcpyAsync(&c->hs->maxHeight, &c->dState->maxHeight, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        tr.mark("scan + max-height reduce");
        if ((rc = hookReduce(c, c->dState->freqLow, kScanLowBins, 1, 0))) return rc;
        if (c->hs->maxHeight >= kScanLowBins) {
            if ((rc = hookReduce(c, c->dFreqTail + kScanLowBins, (size_t)c->hs->maxHeight - kScanLowBins + 1, 0, 0))) return rc;
            CK(launch_tail_max(c->dFreqTail, &c->dState->maxHeight, &c->dState->maxTail, c->stream, &c->launches));
        }
    }
    if (c->allreduce && (rc = timerStop(c, c->tEx1))) return rc;
    // the largest height frequency and, from it, the bin thresholds (one small kernel; no frequency table is materialised:
    // the binning kernels read the counts themselves)
    if ((rc = timerStart(c, c->tLut))) return rc;
    CK(launch_bin_thresholds(c->dState, c->dThresholds, c->dBinParams, c->stream, &c->launches));
    if ((rc = timerStop(c, c->tLut))) return rc;
    CK(cudaMemcpyAsync(c->hs->totals, c->dState->totals, 40, cudaMemcpyDeviceToHost, c->stream));  // totals[2], maxCount, maxHeight, ticket, overflow, maxTail
    CK(cudaEventRecord(c->evLut, c->stream));
    if (c->deviceThresholds) CK(cudaMemcpyAsync(c->hDevThresholds, c->dThresholds, kBins * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    c->hostTableFreq = -1.0f;
    c->stageScan = true;
    c->stageBin = c->stageScore = false;
    c->settled = false;
    tr.mark("enqueued");
    return PPP_OK;
}

// The threshold table of the host libm for the pass's maximum frequency (needs the scalars of the scan stage: waits for them).
int hostThresholds(ppp_ctx* c) {
    CK(cudaEventSynchronize(c->evLut));
    const float maxFreq = (float)(c->hs->maxCount < kSaturate ? (uint32_t)c->hs->maxCount : kSaturate);  // :547-550
    if (c->hostTableFreq == maxFreq) return PPP_OK;
    for (int b = 0; b < kBins; ++b) c->hThresholds[b] = std::numeric_limits<float>::infinity();
    if (maxFreq > 1.0f) {
        computeThresholds(maxFreq, c->hThresholds);
        if (!validateThresholds(maxFreq, c->hThresholds)) return fail(PPP_ERR_STATE, "host log10f is not monotone around a bin threshold");
    }
    c->hostTableFreq = maxFreq;
    return PPP_OK;
}

int enqueueBinStage(ppp_ctx* c) {
    int rc;
    const size_t tableWords = (size_t)c->W * kBins * 4;
    if (!c->deviceThresholds) {  // this host's libm is not the one log10f_emul.h restates: the table comes from the host, in mid-pass
        if ((rc = hostThresholds(c))) return rc;
        const float maxFreq = c->hostTableFreq;
        c->hBinParams->maxFreq = maxFreq;
        c->hBinParams->binScale = maxFreq > 1.0f ? 999.0f / log10f(maxFreq * maxFreq) : 0.0f;
        CK(cudaMemcpyAsync(c->dThresholds, c->hThresholds, kBins * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->dBinParams, c->hBinParams, sizeof(BinParams), cudaMemcpyHostToDevice, c->stream));
    }
    if ((rc = timerStart(c, c->tBin))) return rc;
    if (c->stageBin) CK(cudaMemsetAsync(c->dGrouped, 0, (tableWords + 1) * 4, c->stream));  // (a second ppp_scan on one scan pass: the scan stage zeroed them for the first)
    // (the record count is on the device; the grid is sized for a full store, surplus blocks leave at once)
    CK(launch_bin(c->dRecRp, c->dRecRm, c->dRecMisc, c->dRecBin, c->dState->totals, c->pairCapacity, c->pairCapacity, FreqTables{c->dState->freqLow, c->dFreqTail}, c->dThresholds, c->dBinParams, c->W,
                  c->dGrouped, c->stream, &c->launches));
    if (c->allreduce) {
        // a shard whose signature store overflowed cannot re-run alone (the pass contains collectives): the number of
        // such ranks travels with the grouped counts, and when it is not zero EVERY rank repeats the pass (settle)
        CK(launch_overflow_flag(c->dState->totals, c->pairCapacity, c->ppCapacity, c->dGrouped + tableWords, c->stream, &c->launches));
    }
    if ((rc = timerStop(c, c->tBin))) return rc;
    if (c->allreduce && (rc = timerStart(c, c->tEx2))) return rc;
    if ((rc = hookReduce(c, c->dGrouped, tableWords + 1, 0, 0))) return rc;
    if (c->allreduce && (rc = timerStop(c, c->tEx2))) return rc;
    if (c->allreduce) CK(cudaMemcpyAsync(&c->hs->rerun, c->dGrouped + tableWords, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    c->stageBin = true;
    c->stageScore = false;
    c->settled = false;
    return PPP_OK;
}

int enqueueScoreStage(ppp_ctx* c, uint32_t minStackHeight) {
    int rc;
    const int ppIdx = c->ppOv - c->minOv;
    if ((rc = timerStart(c, c->tScore))) return rc;
    if (minStackHeight > 0) CK(cudaMemsetAsync(c->dScoreState, 0, sizeof(ScoreState) + c->scoreTileWords * sizeof(unsigned long long), c->stream));
    CK(launch_collapse(c->dGrouped, c->W, c->dCells, c->dBinMap, c->dNBins, c->dGroupStart, c->dBinOcc, c->stream, &c->launches));
    CK(launch_fdr_table(c->dCells, c->dNBins, kBins, c->W, ppIdx, c->dXs, c->dWs, c->nSteps, c->dFdr, c->stream, &c->launches));
    CK(launch_score(c->dPP, c->dState->totals, c->ppCapacity, c->ppCapacity, c->dScoreState, c->dScoreTiles, FreqTables{c->dState->freqLow, c->dFreqTail}, c->dThresholds, c->dBinParams, c->dFdr,
                    c->dBinMap, c->W, ppIdx, minStackHeight, c->dSlots, (uint32_t)c->ownedSlots.size(), c->dLoci, c->stream, &c->launches));
    if ((rc = timerStop(c, c->tScore))) return rc;
    if (minStackHeight > 0) CK(cudaMemcpyAsync(&c->hs->nLoci, c->dScoreState, sizeof(ScoreState), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(&c->hs->nBins, c->dNBins, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    if (c->tPass.a && (rc = timerStop(c, c->tPass))) return rc;
    c->scoreMinHeight = minStackHeight;
    c->stageScore = true;
    c->settled = false;
    return PPP_OK;
}

// Waits for everything enqueued so far, validates it and repairs + repeats the pass when it has to (see above).
int settle(ppp_ctx* c) {
    if (c->settled || !c->stageScan) return PPP_OK;
    Trace tr("settle");
    int rc;
    for (int attempt = 0;; ++attempt) {
        if ((rc = hostThresholds(c))) return rc;  // with the host libm, while the device works on the later stages
        tr.mark("host thresholds");
        CK(cudaStreamSynchronize(c->stream));
        tr.mark("stream");
        c->prevMaxHeight = std::max(c->prevMaxHeight, c->hs->maxHeight);
        bool again = false;
        if (c->hs->tallOverflow) {
            // some rank's list of tall stacks did not fit (hs->tallOverflow = the longest list; the same on every rank: read from
            // the reduced buffer): longer lists, unless dense height tables up to the tallest stack are the smaller exchange
            const uint64_t need = (uint64_t)c->hs->tallOverflow + c->hs->tallOverflow / 4 + 64;
            const uint64_t dense = (uint64_t)c->hs->maxHeight + 1;
            if (c->tallCapFixed || need > kTallCapMax || need * (uint64_t)c->world > dense) {
                c->denseTail = true;
            } else {
                c->tallCap = (uint32_t)((need + 63) & ~(uint64_t)63);
                if (getenv("PPP_TRACE")) fprintf(stderr, "[ppp_trace settle] tall-stack lists grow to %u entries per rank\n", c->tallCap);
                if (c->dGather) cudaFree(c->dGather);
                c->dGather = nullptr;
            }
            again = true;
        }
        c->localOverflow = c->hs->totals[0] > c->pairCapacity || c->hs->totals[1] > c->ppCapacity;
        if (c->allreduce) {
            if (c->stageBin && c->hs->rerun) again = true;  // some rank overflowed: every rank repeats, the overflowed ones with a larger store
        } else if (c->localOverflow) {
            again = true;
        }
        if (c->deviceThresholds && memcmp(c->hDevThresholds, c->hThresholds, (kBins - 1) * sizeof(float)) != 0) {
            c->deviceThresholds = false;  // another libm than the one restated in log10f_emul.h: host tables from now on
            again = true;
        }
        if (!again) break;
        if (attempt >= 3) return fail(PPP_ERR_NOMEM, "signature store overflow");
        if (c->localOverflow) {
            const uint64_t want = std::max<uint64_t>(c->hs->totals[0], c->hs->totals[1]);
            if ((rc = ensurePairCapacity(c, want + want / 8 + 1024))) return rc;
        }
        const bool bin = c->stageBin, score = c->stageScore;
        if ((rc = enqueueScanStage(c))) return rc;
        if (bin && (rc = enqueueBinStage(c))) return rc;
        if (score && (rc = enqueueScoreStage(c, c->scoreMinHeight))) return rc;
    }
    c->maxHeight = c->hs->maxHeight;
    c->nPairs = std::min<uint64_t>(c->hs->totals[0], c->pairCapacity);
    c->nPP = std::min<uint64_t>(c->hs->totals[1], c->ppCapacity);
    c->thresholdFreq = c->hostTableFreq > 1.0f ? c->hostTableFreq : 0.0f;
    c->settled = !c->localOverflow;  // (a shard that overflowed before its bin stage is settled by ppp_scan)
    if (!(c->hostTableFreq > 1.0f) && c->nPairs > 0 && !c->allreduce)
        return fail(PPP_ERR_DEGENERATE, "every rounded stack height occurs once: the reference's height score is 0/0 (pingpongpro.cpp:551/:593)");
    return PPP_OK;
}

}  // namespace

extern "C" {

int ppp_height_hist(ppp_ctx* c, const uint64_t** freq, uint32_t* maxHeight) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    if (!c->finished) return fail(PPP_ERR_STATE, "call ppp_stacks_finish before ppp_height_hist");
    CK(cudaSetDevice(c->device));
    invalidateResults(c);
    // signatures are ~0.1-0.3 per read on real densities; a scan that overflows the store reports the exact count and the
    // pass is repeated with a larger one (a shard sizes the store for one signature per read)
    uint64_t want = c->cfgPairCapacity ? c->cfgPairCapacity : std::max<uint64_t>((uint64_t)1 << 20, c->allreduce ? c->readsSubmitted : c->readsSubmitted / 3);
    if (want < c->pairCapacity) want = c->pairCapacity;
    int rc = ensurePairCapacity(c, want);
    if (rc) return rc;
    if ((rc = enqueueScanStage(c))) return rc;
    c->haveHist = true;
    if (!freq && !maxHeight) return PPP_OK;  // nothing to return: the host looks at the pass later
    if ((rc = settle(c))) { c->haveHist = false; return rc; }
    if (freq) {
        uint32_t mh = c->maxHeight;
        c->hFreq.assign((size_t)mh + 1, 0);
        std::vector<unsigned long long> low(kScanLowBins);
        CK(cudaMemcpyAsync(low.data(), c->dState->freqLow, kScanLowBins * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        std::vector<uint32_t> tail;
        if (mh >= kScanLowBins) {
            tail.resize((size_t)mh - kScanLowBins + 1);
            CK(cudaMemcpyAsync(tail.data(), c->dFreqTail + kScanLowBins, tail.size() * 4, cudaMemcpyDeviceToHost, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
        for (uint32_t h = 0; h <= mh; ++h) c->hFreq[h] = h < kScanLowBins ? low[h] : tail[h - kScanLowBins];
        *freq = c->hFreq.data();
    }
    if (maxHeight) *maxHeight = c->maxHeight;
    return PPP_OK;
}

int ppp_scan(ppp_ctx* c, uint64_t* grouped, uint64_t* nPairs) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    if (!c->haveHist) return fail(PPP_ERR_STATE, "call ppp_height_hist before ppp_scan");
    CK(cudaSetDevice(c->device));
    int rc;
    const size_t tableWords = (size_t)c->W * kBins * 4;
    Trace tr("scan");
    if ((rc = enqueueBinStage(c))) return rc;
    tr.mark("bin stage enqueued");
    c->haveScan = true;
    c->viewValid = false;
    if (!grouped && !nPairs) return PPP_OK;
    if ((rc = settle(c))) { c->haveScan = false; return rc; }
    if (grouped) {
        std::vector<uint32_t> g(tableWords);
        CK(cudaMemcpyAsync(g.data(), c->dGrouped, tableWords * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int o = 0; o < c->W; ++o)  // device tables are bin-major, the ABI is [overlap][bin][bias][local]
            for (int b = 0; b < kBins; ++b)
                for (int k = 0; k < 4; ++k) grouped[((size_t)o * kBins + b) * 4 + k] = g[((size_t)b * c->W + o) * 4 + k];
    }
    if (nPairs) *nPairs = c->nPairs;
    return PPP_OK;
}

int ppp_score(ppp_ctx* c, uint32_t minStackHeight, uint32_t* nBins, uint16_t* binMap, float* cells, float* fdrTable, const ppp_locus** loci, size_t* nLoci) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    if (!c->haveScan) return fail(PPP_ERR_STATE, "call ppp_scan before ppp_score");
    CK(cudaSetDevice(c->device));
    int rc;
    Trace tr("score");
    if ((rc = enqueueScoreStage(c, minStackHeight))) return rc;
    tr.mark("score stage enqueued");
    if ((rc = settle(c))) return rc;
    tr.mark("pass settled");
    size_t n = minStackHeight > 0 ? (size_t)c->hs->nLoci : (size_t)c->nPP;  // -s 0: every overlap-`pingpong` record is a locus
    uint32_t C = c->hs->nBins;
    if (loci) {
        if (n > c->hLociCap) {
            if (c->hLoci) cudaFreeHost(c->hLoci);
            c->hLoci = nullptr;
            c->hLociCap = 0;
            CK(cudaMallocHost(&c->hLoci, (n + n / 4 + 1024) * sizeof(ppp_locus)));
            c->hLociCap = n + n / 4 + 1024;
        }
        static_assert(sizeof(ppp_locus) == sizeof(LocusOut), "locus layout");
        if (n) CK(cudaMemcpyAsync(c->hLoci, c->dLoci, n * sizeof(ppp_locus), cudaMemcpyDeviceToHost, c->stream));
    }
    size_t tableWords = (size_t)c->W * kBins * 4;
    std::vector<float> tmpCells, tmpFdr;
    if (cells) { tmpCells.resize(tableWords); CK(cudaMemcpyAsync(tmpCells.data(), c->dCells, tableWords * 4, cudaMemcpyDeviceToHost, c->stream)); }
    if (fdrTable) { tmpFdr.resize(tableWords); CK(cudaMemcpyAsync(tmpFdr.data(), c->dFdr, tableWords * 4, cudaMemcpyDeviceToHost, c->stream)); }
    if (binMap) CK(cudaMemcpyAsync(binMap, c->dBinMap, kBins * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream));
    if (loci || cells || fdrTable || binMap) CK(cudaStreamSynchronize(c->stream));
    tr.mark("loci to host");
    // device tables are bin-major; the ABI returns them as [overlap][C][bias][local]
    for (int o = 0; o < c->W; ++o)
        for (uint32_t b = 0; b < C; ++b)
            for (int k = 0; k < 4; ++k) {
                if (cells) cells[((size_t)o * C + b) * 4 + k] = tmpCells[((size_t)b * c->W + o) * 4 + k];
                if (fdrTable) fdrTable[((size_t)o * C + b) * 4 + k] = tmpFdr[((size_t)b * c->W + o) * 4 + k];
            }
    if (nBins) *nBins = C;
    if (loci) *loci = c->hLoci;
    if (nLoci) *nLoci = n;
    c->haveScore = true;
    return PPP_OK;
}

int ppp_pass_info(ppp_ctx* c, uint32_t* maxHeight, uint64_t* nPairs, uint64_t* nLoci, uint64_t* nStacks) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    if (!c->haveHist) return fail(PPP_ERR_STATE, "call ppp_height_hist before ppp_pass_info");
    CK(cudaSetDevice(c->device));
    int rc = settle(c);
    if (rc) return rc;
    if (maxHeight) *maxHeight = c->maxHeight;
    if (nPairs) *nPairs = c->nPairs;
    if (nLoci) *nLoci = c->nPP;
    if (nStacks) *nStacks = c->nStacksP + c->nStacksM;
    return PPP_OK;
}

int ppp_pairs_download(ppp_ctx* c, ppp_pair* out, size_t capacity, size_t* n) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    if (!c->haveScan) return fail(PPP_ERR_STATE, "call ppp_scan before ppp_pairs_download");
    CK(cudaSetDevice(c->device));
    { int rcSettle = settle(c); if (rcSettle) return rcSettle; }
    if (n) *n = (size_t)c->nPairs;
    if (!out) return PPP_OK;
    if (capacity < c->nPairs) return fail(PPP_ERR_ARG, "capacity %zu < %llu signatures", capacity, (unsigned long long)c->nPairs);
    int rcView = ensurePairView(c);
    if (rcView) return rcView;
    std::vector<PairRec> recs((size_t)c->nPairs);
    std::vector<float> fdr((size_t)c->nPairs, 0.f);
    std::vector<uint16_t> binMap(kBins, 0);
    if (c->nPairs) CK(cudaMemcpyAsync(recs.data(), c->dPairs, recs.size() * sizeof(PairRec), cudaMemcpyDeviceToHost, c->stream));
    if (c->haveScore) {
        if (c->nPairs) CK(cudaMemcpyAsync(fdr.data(), c->dPairFdr, fdr.size() * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(binMap.data(), c->dBinMap, kBins * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    size_t slot = 0;
    for (size_t i = 0; i < recs.size(); ++i) {
        const PairRec& r = recs[i];
        while (slot + 1 < c->ownedSlots.size() && c->ownedSlots[slot + 1].off <= r.lpos) ++slot;  // records are position-ordered
        const Slot& s = c->ownedSlots[slot];
        ppp_pair& p = out[i];
        p.contig = s.contig;
        p.position = r.lpos - s.off + s.first;
        p.overlap = (r.misc & 31u) + (uint32_t)c->minOv;
        p.height_score_bin = (r.misc >> 7) & 1023u;
        p.collapsed_bin = c->haveScore ? binMap[p.height_score_bin] : 0;
        p.local_bin = (r.misc >> 5) & 1u;
        p.bias_bin = (r.misc >> 6) & 1u;
        p.reads_plus = r.rp;
        p.reads_minus = r.rm;
        p.fdr = fdr[i];
    }
    return PPP_OK;
}

int ppp_transposons(ppp_ctx* c, const ppp_interval* intervals, size_t n, ppp_tp_result* results, uint32_t* order) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    if (!c->haveScore) return fail(PPP_ERR_STATE, "call ppp_score before ppp_transposons");
    if (n && (!intervals || !results)) return fail(PPP_ERR_ARG, "null argument");
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = ensurePairView(c))) return rc;
    if ((rc = timerStart(c, c->tTransposon))) return rc;
    TransposonJob job;
    job.pairs = c->dPairs;
    job.pairFdr = c->dPairFdr;
    job.nPairs = c->nPairs;
    job.W = c->W;
    job.ppIdx = c->ppOv - c->minOv;
    job.slots = &c->slotOfContig;
    job.stream = c->stream;
    job.launches = &c->launches;
    std::string err;
    // one process per GPU: the shards score the transposons together (those that straddle a range boundary included)
    const bool collective = c->allreduce && c->world > 1;
    cudaError_t e = collective ? run_transposons_collective(job, c->rank, c->world, [c](void* buf, size_t words) { return hookReduce(c, buf, words, 0, 0); },
                                                            intervals, n, results, order, err)
                               : run_transposons(job, intervals, n, results, order, err);
    if (e != cudaSuccess) return fail(PPP_ERR_CUDA, "transposons: %s %s", cudaGetErrorString(e), err.c_str());
    if (!err.empty()) return fail(PPP_ERR_ARG, "%s", err.c_str());
    if ((rc = timerStop(c, c->tTransposon))) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return PPP_OK;
}

int ppp_transposons_sharded(ppp_ctx* const* ctxs, int n_ctx, const ppp_interval* intervals, size_t n, ppp_tp_result* results, uint32_t* order) {
    if (!ctxs || n_ctx <= 0) return fail(PPP_ERR_ARG, "no contexts");
    if (n && (!intervals || !results)) return fail(PPP_ERR_ARG, "null argument");
    std::vector<TransposonJob> jobs(n_ctx);
    std::vector<int> devices(n_ctx);
    bool havePrev = false;
    uint64_t prevKey = 0;
    for (int k = 0; k < n_ctx; ++k) {
        ppp_ctx* c = ctxs[k];
        if (!c) return fail(PPP_ERR_ARG, "null context");
        if (!c->haveScore) return fail(PPP_ERR_STATE, "call ppp_score on every context before ppp_transposons_sharded");
        CK(cudaSetDevice(c->device));
        int rcView = ensurePairView(c);
        if (rcView) return rcView;
        if (c->contigLen != ctxs[0]->contigLen || c->W != ctxs[0]->W || c->minOv != ctxs[0]->minOv || c->ppOv != ctxs[0]->ppOv)
            return fail(PPP_ERR_ARG, "contexts %d and 0 do not describe the same genome and overlap range", k);
        if (!c->ownedSlots.empty()) {  // ranges must ascend: the sums are chained in position order
            const Slot& f = c->ownedSlots.front();
            uint64_t key = ((uint64_t)f.contig << 32) | f.first;
            if (havePrev && key <= prevKey) return fail(PPP_ERR_ARG, "contexts must be passed in ascending range order");
            const Slot& l = c->ownedSlots.back();
            prevKey = ((uint64_t)l.contig << 32) | l.first;
            havePrev = true;
        }
        TransposonJob& job = jobs[k];
        job.pairs = c->dPairs;
        job.pairFdr = c->dPairFdr;
        job.nPairs = c->nPairs;
        job.W = c->W;
        job.ppIdx = c->ppOv - c->minOv;
        job.slots = &c->slotOfContig;
        job.stream = c->stream;
        job.launches = &c->launches;
        devices[k] = c->device;
    }
    std::string err;
    cudaError_t e = n_ctx == 1 ? (cudaSetDevice(devices[0]), run_transposons(jobs[0], intervals, n, results, order, err))
                               : run_transposons_sharded(jobs.data(), devices.data(), n_ctx, intervals, n, results, order, err);
    if (e != cudaSuccess) return fail(PPP_ERR_CUDA, "transposons: %s %s", cudaGetErrorString(e), err.c_str());
    if (!err.empty()) return fail(PPP_ERR_ARG, "%s", err.c_str());
    return PPP_OK;
}

int ppp_emul_bin_thresholds(float max_freq, float* thresholds) {
    if (!thresholds || !(max_freq > 1.0f)) return fail(PPP_ERR_ARG, "ppp_emul_bin_thresholds: max_freq must be > 1");
    for (int b = 1; b < kBins; ++b) thresholds[b - 1] = emul_bin_threshold(b, max_freq);
    return PPP_OK;
}

int ppp_emul_log10f(const float* x, size_t n, float* out) {
    if (n && (!x || !out)) return fail(PPP_ERR_ARG, "null argument");
    for (size_t i = 0; i < n; ++i) out[i] = emul_log10f(x[i]);
    return PPP_OK;
}

int ppp_emul_fadd_chain(float h0, const float* w, size_t n, uint32_t piece, float* out, uint64_t* realAdditions) {
    if ((n && !w) || !out || piece == 0) return fail(PPP_ERR_ARG, "bad argument");
    uint32_t bits;
    memcpy(&bits, &h0, 4);
    uint64_t real = 0;
    for (size_t b = 0; b < n; b += piece) {
        const size_t e = std::min(n, b + (size_t)piece);
        AddMap m = {0u, 0u};
        for (size_t i = b; i < e; ++i) {
            uint32_t wb;
            memcpy(&wb, &w[i], 4);
            m = addmap_compose(m, addmap_of(bits >> 23, wb));
        }
        if (!addmap_crosses(bits, m)) { bits = addmap_apply(bits, m); continue; }
        float h;
        memcpy(&h, &bits, 4);
        for (size_t i = b; i < e; ++i) { volatile float t = h + w[i]; h = t; }
        memcpy(&bits, &h, 4);
        real += e - b;
    }
    memcpy(out, &bits, 4);
    if (realAdditions) *realAdditions = real;
    return PPP_OK;
}

int ppp_host_bin_thresholds(float max_freq, float* thresholds) {
    if (!thresholds || !(max_freq > 1.0f)) return fail(PPP_ERR_ARG, "ppp_host_bin_thresholds: max_freq must be > 1");
    computeThresholds(max_freq, thresholds);
    if (!validateThresholds(max_freq, thresholds)) return fail(PPP_ERR_STATE, "host log10f is not monotone around a bin threshold");
    return PPP_OK;
}

int ppp_timings_get(ppp_ctx* c, ppp_timings* out) {
    if (!c || !out) return fail(PPP_ERR_ARG, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    foldBuildTimers(c);
    memset(out, 0, sizeof *out);
    out->stack_build_ms = c->buildMs;
    out->scan_ms = timerMs(c->tScan);
    out->lut_ms = timerMs(c->tLut);
    out->bin_ms = timerMs(c->tBin);
    out->score_ms = timerMs(c->tScore);
    out->transposon_ms = timerMs(c->tTransposon);
    out->launches = (uint32_t)c->launches;
    out->exchange_ms = timerMs(c->tEx1) + timerMs(c->tEx2);
    out->pass_ms = timerMs(c->tPass);
    return PPP_OK;
}

int ppp_timings_reset(ppp_ctx* c) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    foldBuildTimers(c);
    c->buildMs = 0.f;
    c->launches = 0;
    for (Timer* t : {&c->tScan, &c->tLut, &c->tBin, &c->tScore, &c->tTransposon, &c->tEx1, &c->tEx2, &c->tPass}) t->used = false;
    return PPP_OK;
}

int ppp_layout(ppp_ctx* c, uint64_t* wordsPerStrand, uint64_t* ownedPositions) {
    if (!c) return fail(PPP_ERR_ARG, "null context");
    if (wordsPerStrand) *wordsPerStrand = c->G;
    if (ownedPositions) *ownedPositions = c->ownedPositions;
    return PPP_OK;
}

}  // extern "C"
