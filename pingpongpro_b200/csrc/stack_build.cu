// stack_build.cu -- K1, the per-strand 5'-end stack builder.
//
// Replaces the accumulation of countReadsInBamFile, pingpongpro.cpp:458-488:
//     stack[strand][contig][key].reads += weight;   .AAtPosition10 |= a10
// over the dense word arrays described in common.cuh.
//
// Integer path (-m unique, -m discard: every weight is 1, so the float sum of :487 is the exact
// count clipped at 2^24): one warp-aggregated red.add per distinct word per warp (reads of a hot stack
// that sit in the same warp are merged with match.any before they reach L2) and a red.or for the A10 bit.
//
// Weighted path (-m weighted): `reads += 1.0/NH` is a float accumulation in FILE ORDER (:487), which no
// unordered scatter can reproduce bit-exactly.  Each batch is therefore stably radix-sorted by word index
// (own LSD sort, 8-bit digits: block histograms -> device scan -> stable scatter) and every run of equal
// keys is folded sequentially with __fadd_rn, continuing from the word's current value so that
// successive batches extend the same left-to-right sum: runs of up to 32 records by one thread each
// (k_fold_runs), longer ones by a producer/consumer warp pair (k_fold_long), runs of 8192 records and more by a whole
// block that composes the additions as integer maps per binade instead of executing them one by one (k_fold_huge,
// fadd_chain.h), batches of up to 4096 records
// in a single launch (k_small_batch).  The launchers are split into a sort half and a fold half so that the
// caller (api.cu: buildChunk) can run the sort of the next batches beside the folds of earlier ones.
//
// Both paths also maintain the OCCUPANCY BITMAP (1 bit per word of H, same index space: word i of H <-> bit i):
// the thread that turns a word from 0 to non-zero sets its bit.  The scan (scan.cu) reads the bitmaps instead of
// streaming the ~96 % empty arrays.
//
// Bound: L2 atomic throughput / 32-byte sector read-modify-write traffic, not streaming bandwidth.
// Algorithmic bytes: 8 N (records) + 8 N (word read+write) + 8 G (zero fill at init/reset).
#include "common.cuh"
#include "fadd_chain.h"

namespace pppk {

namespace {

constexpr int kBuildThreads = 256;

// counters[0] = foreign reads (another shard's), counters[1] = reads beyond their contig (range error)
template <bool WEIGHTED>
__device__ __forceinline__ bool locate(const ppp_read r, const Slot& slot, int mode, uint32_t& pos, unsigned long long* counters) {
    uint32_t strand = r.meta & PPP_READ_MINUS;
    uint32_t nh = r.meta >> PPP_READ_NH_SHIFT;
    if (!WEIGHTED && mode == PPP_MULTIHITS_DISCARD && nh != 1) return false;  // :454 + :458
    if (r.key >= slot.keyLimit) { atomicAdd(&counters[1], 1ull); return false; }
    uint32_t rel = r.key - slot.first;  // wraps to a huge value when key < first
    uint32_t limit = strand ? slot.lenMinus : slot.lenPlus;
    if (rel >= limit) { atomicAdd(&counters[0], 1ull); return false; }
    pos = slot.off + rel;
    return true;
}

__global__ void __launch_bounds__(kBuildThreads) k_stack_build_int(const ppp_read* __restrict__ reads, size_t n, uint32_t* __restrict__ H, uint64_t G,
                                                                    uint32_t* __restrict__ occ, Slot slot, int mode, unsigned long long* counters) {
    const unsigned lane = threadIdx.x & 31;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t nRound = (n + 31) & ~(size_t)31;  // whole warps stay converged for the warp primitives
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nRound; i += stride) {
        bool valid = i < n;
        ppp_read r = {0, 0};
        if (valid) {
            uint2 raw = __ldg(reinterpret_cast<const uint2*>(reads) + i);
            r.key = raw.x;
            r.meta = raw.y;
        }
        uint32_t pos = 0;
        if (valid) valid = locate<false>(r, slot, mode, pos, counters);
        unsigned long long widx = (r.meta & PPP_READ_MINUS) ? G + pos : (unsigned long long)pos;
        unsigned active = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            unsigned peers = __match_any_sync(active, widx);       // lanes that hit the same word
            unsigned a10 = __ballot_sync(active, (r.meta & PPP_READ_A10) != 0);
            if (lane == (unsigned)(__ffs(peers) - 1)) {
                uint32_t old = atomicAdd(&H[widx], (uint32_t)__popc(peers));  // reads += 1 per read (:487)
                if ((old & kValueMask) == 0u) atomicOr(&occ[widx >> 5], 1u << (widx & 31u));  // first read of this stack: mark it occupied
                if (a10 & peers) atomicOr(&H[widx], kA10Bit);       // AAtPosition10 = true (:471/:483)
            }
        }
    }
}

// ------------------------------------------------------------------ weighted path: keys
__global__ void __launch_bounds__(kBuildThreads) k_sort_keys(const ppp_read* __restrict__ reads, size_t n, Slot slot, uint32_t* __restrict__ keys,
                                                              uint32_t* __restrict__ vals, unsigned long long* counters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint2 raw = __ldg(reinterpret_cast<const uint2*>(reads) + i);
    ppp_read r = {raw.x, raw.y};
    uint32_t pos;
    bool ok = locate<true>(r, slot, PPP_MULTIHITS_WEIGHTED, pos, counters);
    keys[i] = ok ? pos : kInvalidPos;  // invalid records sort to the very end
    vals[i] = r.meta;
}

// ------------------------------------------------------------------ stable LSD radix sort
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortPerWarp = 512;                      // consecutive keys owned by one warp
constexpr int kSortTile = kSortWarps * kSortPerWarp;  // 4096 keys per block
constexpr int kSortRounds = kSortPerWarp / 32;

__global__ void __launch_bounds__(kSortThreads) k_radix_hist(const uint32_t* __restrict__ keys, size_t n, int shift, uint32_t* __restrict__ blockHist,
                                                              uint32_t nBlocks) {
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    size_t base = (size_t)blockIdx.x * kSortTile;
    for (int k = threadIdx.x; k < kSortTile; k += kSortThreads) {
        size_t i = base + k;
        if (i < n) atomicAdd(&hist[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    blockHist[(size_t)threadIdx.x * nBlocks + blockIdx.x] = hist[threadIdx.x];  // [digit][block]
}

// offsets = exclusive scan of blockHist ([digit][block] flattened): the first output slot of (digit, block).
// The tile is first put in digit order in shared memory (stable: warp, round, lane = input order), then written out, so
// that consecutive threads store to consecutive addresses wherever the tile holds several keys of one digit -- a direct
// scatter costs one 32-byte sector write per 4-byte key and is bound by exactly that.
__global__ void __launch_bounds__(kSortThreads, 3) k_radix_scatter(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                                                 uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut, size_t n, int shift,
                                                                 const uint32_t* __restrict__ offsets, uint32_t nBlocks) {
    __shared__ uint32_t cnt[kSortWarps][256];
    __shared__ uint32_t sKey[kSortTile], sVal[kSortTile];
    __shared__ uint32_t sBase[256];      // destination of the tile's first key of a digit, minus that key's place in the tile
    __shared__ uint32_t sWarpSum[kSortWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < kSortWarps * 256; k += kSortThreads) (&cnt[0][0])[k] = 0;
    __syncthreads();
    const size_t tileBase = (size_t)blockIdx.x * kSortTile;
    const size_t wbase = tileBase + (size_t)warp * kSortPerWarp;
    uint32_t key[kSortRounds], val[kSortRounds];
    uint32_t rank[kSortRounds / 2];  // place among the warp's earlier keys of the same digit, two per register
    // phase 1: per-warp digit counts and ranks (each warp owns a contiguous run of the tile, so order = warp, round, lane)
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        size_t i = wbase + r * 32 + lane;
        bool ok = i < n;
        key[r] = ok ? keysIn[i] : 0xffffffffu;
        val[r] = ok ? valsIn[i] : 0;
    }
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const bool ok = wbase + r * 32 + lane < n;
        const unsigned act = __ballot_sync(0xffffffffu, ok);
        uint32_t rk = 0;
        if (ok) {
            const uint32_t d = (key[r] >> shift) & 255u;
            const unsigned peers = __match_any_sync(act, d);
            rk = cnt[warp][d] + __popc(peers & ((1u << lane) - 1u));
            __syncwarp(act);
            if (lane == __ffs(peers) - 1) cnt[warp][d] += __popc(peers);
        }
        __syncwarp();
        if (r & 1) rank[r / 2] |= rk << 16; else rank[r / 2] = rk;
    }
    __syncthreads();
    // phase 2: place of (warp, digit) inside the digit-ordered tile, and where the tile's keys of a digit go
    {
        const uint32_t d = threadIdx.x;
        uint32_t c[kSortWarps], total = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) { c[w] = cnt[w][d]; total += c[w]; }
        uint32_t inc = total;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, k);
            if (lane >= k) inc += t;
        }
        if (lane == 31) sWarpSum[warp] = inc;
        __syncthreads();
        uint32_t before = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w)
            if (w < warp) before += sWarpSum[w];
        uint32_t run = before + inc - total;  // keys of smaller digits in this tile
        sBase[d] = offsets[(size_t)d * nBlocks + blockIdx.x] - run;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) { cnt[w][d] = run; run += c[w]; }
    }
    __syncthreads();
    // phase 3: the tile in digit order
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        if (wbase + r * 32 + lane < n) {
            const uint32_t d = (key[r] >> shift) & 255u;
            const uint32_t j = cnt[warp][d] + ((r & 1) ? rank[r / 2] >> 16 : rank[r / 2] & 0xffffu);
            sKey[j] = key[r];
            sVal[j] = val[r];
        }
    }
    __syncthreads();
    // phase 4: out, in tile order
    const uint32_t count = (uint32_t)(n - tileBase < (size_t)kSortTile ? n - tileBase : (size_t)kSortTile);
#pragma unroll 4
    for (uint32_t j = threadIdx.x; j < count; j += kSortThreads) {
        const uint32_t k = sKey[j];
        const uint32_t dst = sBase[(k >> shift) & 255u] + j;
        keysOut[dst] = k;
        valsOut[dst] = sVal[j];
    }
}

// ------------------------------------------------------------------ ordered fold
// Runs of equal keys are in file order (stable sort); plus- and minus-strand records of the same position are
// folded into their own words.  Short runs: one thread per run (the thread at the run's first element).  Runs
// longer than kShortRun (piRNA stacks are heavy tailed: a single position can hold > 10^6 reads) are deferred to
// k_fold_long, where a warp finds the end of the run, streams its records with a deep software pipeline, evaluates
// the weights 32 at a time and only the __fadd_rn chain itself stays serial (adding +0.0f for the other strand's
// records is exact).
constexpr int kShortRun = 32;
constexpr int kFoldBlock = 256;  // records per pipeline stage of k_fold_long
constexpr int kWeightLut = 64;
constexpr int kHugeRun = 8192;        // runs at least this long leave the serial chain of k_fold_long for k_fold_huge
constexpr int kHugeThreads = 512;
constexpr int kHugePerLane = 32;      // consecutive records per lane and trip of k_fold_huge
constexpr int kHugeTrip = kHugeThreads * kHugePerLane;

__device__ __forceinline__ float read_weight_weighted(uint32_t meta) {
    // readWeight = 1.0 / multiHits in double, narrowed to float (:449).  For multiHits < 2^24 the correctly rounded
    // single-precision reciprocal is the same number: 1/n stays at least 2^-24/n (relative) away from every
    // rounding tie of the float format, far more than the 2^-53 the intermediate double could move it
    // (checked exhaustively in tests/test_frontend.py).
    const uint32_t nh = meta >> PPP_READ_NH_SHIFT;
    if (nh < (1u << 24)) return __frcp_rn((float)nh);
    return (float)__ddiv_rn(1.0, (double)nh);
}

__global__ void __launch_bounds__(kBuildThreads) k_fold_runs(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, size_t n,
                                                              uint32_t* __restrict__ H, uint64_t G, uint32_t* __restrict__ occ,
                                                              unsigned long long* __restrict__ longRuns, unsigned int* __restrict__ nLong, unsigned int longCap) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t pos = keys[i];
    if (pos == kInvalidPos) return;
    if (i > 0 && keys[i - 1] == pos) return;  // not a run head
    size_t end = i + 1;
    while (end < n && end - i <= (size_t)kShortRun && keys[end] == pos) ++end;
    if (end - i > (size_t)kShortRun) {  // long run: hand over to a warp
        unsigned int slot = atomicAdd(nLong, 1u);
        if (slot < longCap) { longRuns[slot] = (unsigned long long)i; return; }
        while (end < n && keys[end] == pos) ++end;  // list full (cannot happen with longCap = n / kShortRun): fold here
    }
    uint32_t wp = H[pos], wm = H[G + pos];
    float hp = __uint_as_float(wp & kValueMask), hm = __uint_as_float(wm & kValueMask);
    uint32_t ap = wp & kA10Bit, am = wm & kA10Bit;
    bool anyP = false, anyM = false;
    for (size_t j = i; j < end; ++j) {
        uint32_t meta = vals[j];
        float w = read_weight_weighted(meta);
        if (!(w > 0.0f)) continue;                     // :458
        if (meta & PPP_READ_MINUS) {
            hm = __fadd_rn(hm, w);                     // :487
            if (meta & PPP_READ_A10) am = kA10Bit;     // :471
            anyM = true;
        } else {
            hp = __fadd_rn(hp, w);
            if (meta & PPP_READ_A10) ap = kA10Bit;     // :483
            anyP = true;
        }
    }
    if (anyP) {
        H[pos] = (__float_as_uint(hp) & kValueMask) | ap;
        if (wp == 0u) atomicOr(&occ[pos >> 5], 1u << (pos & 31u));  // new stack: mark it occupied
    }
    if (anyM) {
        H[G + pos] = (__float_as_uint(hm) & kValueMask) | am;
        if (wm == 0u) atomicOr(&occ[(G + pos) >> 5], 1u << (pos & 31u));
    }
}

// End of the run of `pos` that starts at keys[i] (keys are sorted, keys[i + kShortRun] == pos is known): the whole warp
// gallops (lane l probes i + 32 * 2^l) and then narrows the bracket 32-fold per round of independent loads.
__device__ __forceinline__ size_t warp_run_end(const uint32_t* __restrict__ keys, size_t i, size_t n, uint32_t pos, unsigned lane) {
    size_t p = i + ((size_t)kShortRun << lane);
    unsigned hit = __ballot_sync(0xffffffffu, lane < 26 && p < n && keys[p] == pos);  // a prefix of the lanes
    int k = __popc(hit);                                                              // >= 1 (lane 0 probes i + 32)
    size_t lo = i + ((size_t)kShortRun << (k - 1));                                   // keys[lo] == pos
    size_t hi = i + ((size_t)kShortRun << k);                                         // keys[hi] != pos (or hi >= n)
    if (hi > n) hi = n;
    while (hi - lo > 1) {
        const size_t step = (hi - lo + 31) / 32;
        p = lo + (size_t)(lane + 1) * step;
        hit = __ballot_sync(0xffffffffu, p < hi && keys[p] == pos);
        k = __popc(hit);
        const size_t nlo = lo + (size_t)k * step, nhi = lo + (size_t)(k + 1) * step;
        lo = nlo;
        if (nhi < hi) hi = nhi;
    }
    return hi;
}

// Barrier among the 64 threads of one warp pair (named barriers 1..; 0 is __syncthreads).
__device__ __forceinline__ void pair_barrier(unsigned pair) { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); }

// One WARP PAIR per long run.  Nothing hides latency for a run but the threads that own it, so the work is split by
// kind: the producer warp streams the records (256 per stage, the next stage's loads in flight), turns NH into
// weights 32 at a time through a shared-memory table and stages them per strand in shared memory; the consumer warp
// reads a staged block back with broadcast 128-bit loads and does nothing but the dependent additions (:487, file
// order) -- 4.5 cycles per read (tools/ub_chain.cu) instead of 6.8 when one warp did both.  Two staging buffers; one
// pair barrier per block.
__global__ void __launch_bounds__(kBuildThreads) k_fold_long(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, size_t n,
                                                              uint32_t* __restrict__ H, uint64_t G, uint32_t* __restrict__ occ,
                                                              const unsigned long long* __restrict__ longRuns, const unsigned int* __restrict__ nLong,
                                                              unsigned int longCap, unsigned long long* __restrict__ hugeRuns, unsigned int* __restrict__ nHuge) {
    constexpr int kPairs = kBuildThreads / 64;
    constexpr int kDepth = kFoldBlock / 32;
    __shared__ __align__(16) float stage[kPairs][2][2][kFoldBlock];  // [pair][buffer][strand][record]
    __shared__ unsigned sMask[kPairs][2][2];                        // [pair][buffer][strand]: lanes that staged a record of that strand
    __shared__ unsigned long long sEnd[kPairs];
    __shared__ unsigned sFlags[kPairs][2];                          // per strand: OR of the metas (A10), bit 31 = any record
    __shared__ float weightOf[kWeightLut];  // weights of the common multiplicities (the reciprocal is ~1/4 of the instructions otherwise)
    if (threadIdx.x < kWeightLut) weightOf[threadIdx.x] = read_weight_weighted((uint32_t)threadIdx.x << PPP_READ_NH_SHIFT);
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const unsigned pair = threadIdx.x >> 6;
    const bool producer = ((threadIdx.x >> 5) & 1u) == 0;
    const unsigned pairsPerGrid = gridDim.x * kPairs;
    unsigned int count = *nLong;
    if (count > longCap) count = longCap;
    for (unsigned int run = blockIdx.x * kPairs + pair; run < count; run += pairsPerGrid) {
        const size_t i = (size_t)longRuns[run];
        const uint32_t pos = keys[i];
        if (producer) {
            const size_t end = warp_run_end(keys, i, n, pos, lane);
            const bool huge = end - i >= (size_t)kHugeRun;  // a chain this long is folded by a whole block (k_fold_huge)
            if (lane == 0) sEnd[pair] = huge ? ~0ull : end;
            if (huge) {
                if (lane == 0) {
                    const unsigned int slot = atomicAdd(nHuge, 1u);
                    hugeRuns[2 * (size_t)slot] = (unsigned long long)i;
                    hugeRuns[2 * (size_t)slot + 1] = (unsigned long long)end;
                }
                pair_barrier(pair);  // the consumer reads the mark ...
                pair_barrier(pair);  // ... before the next run overwrites it
                continue;
            }
            const size_t nBlocks = (end - i + kFoldBlock - 1) / kFoldBlock;
            uint32_t vq[kDepth];
            auto fetch = [&](size_t blk) {
                const size_t base = i + blk * kFoldBlock;
                const uint32_t* src = vals + base + lane;
                const unsigned left = (unsigned)min((size_t)kFoldBlock, end - base);
#pragma unroll
                for (int u = 0; u < kDepth; ++u) vq[u] = u * 32 + lane < left ? src[u * 32] : 0u;
            };
            unsigned accP = 0, accM = 0;
            // weights of block `blk` (held in vq) -> shared memory, branch free (the eight table look-ups are in flight
            // together).  Every weight of this mode is positive (1 / NH, +inf for NH:i:0), so `weight > 0` of :458 is moot.
            auto stageBlock = [&](size_t blk) {
                float* sp = stage[pair][blk & 1][0];
                float* sm = stage[pair][blk & 1][1];
                const unsigned left = (unsigned)min((size_t)kFoldBlock, end - (i + blk * kFoldBlock));
                uint32_t vc[kDepth];
#pragma unroll
                for (int u = 0; u < kDepth; ++u) vc[u] = vq[u];
                if (blk + 1 < nBlocks) fetch(blk + 1);  // in flight while this block is staged and the previous one is added
                unsigned laneP = 0, laneM = 0, big = 0;
#pragma unroll
                for (int u = 0; u < kDepth; ++u) {
                    const bool in = u * 32 + lane < left;
                    const uint32_t nh = vc[u] >> PPP_READ_NH_SHIFT;
                    const float wt = weightOf[min(nh, (uint32_t)kWeightLut - 1u)];
                    big |= nh;
                    const bool minus = (vc[u] & PPP_READ_MINUS) != 0;
                    sp[u * 32 + lane] = in && !minus ? wt : 0.0f;  // adding +0.0f is exact
                    sm[u * 32 + lane] = in && minus ? wt : 0.0f;
                    laneP |= in && !minus ? vc[u] | 0x80000000u : 0u;  // A10 flag: :483
                    laneM |= in && minus ? vc[u] | 0x80000000u : 0u;   // :471
                }
                if (__any_sync(0xffffffffu, big >= (uint32_t)kWeightLut)) {  // a multiplicity beyond the table: evaluate those records
#pragma unroll  // (registers cannot be indexed dynamically)
                    for (int u = 0; u < kDepth; ++u) {
                        if ((vc[u] >> PPP_READ_NH_SHIFT) < (uint32_t)kWeightLut || !(u * 32 + lane < left)) continue;
                        const float wt = read_weight_weighted(vc[u]);
                        if (vc[u] & PPP_READ_MINUS) sm[u * 32 + lane] = wt; else sp[u * 32 + lane] = wt;
                    }
                }
                accP |= laneP;
                accM |= laneM;
                const unsigned mp = __ballot_sync(0xffffffffu, laneP != 0), mm = __ballot_sync(0xffffffffu, laneM != 0);
                if (lane == 0) { sMask[pair][blk & 1][0] = mp; sMask[pair][blk & 1][1] = mm; }
            };
            fetch(0);
            stageBlock(0);
            pair_barrier(pair);  // block 0 (and the run's end) visible to the consumer
            for (size_t blk = 0; blk < nBlocks; ++blk) {
                if (blk + 1 < nBlocks) stageBlock(blk + 1);  // the other buffer: its previous block was consumed before the last barrier
                if (blk + 1 == nBlocks) {
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) { accP |= __shfl_xor_sync(0xffffffffu, accP, d); accM |= __shfl_xor_sync(0xffffffffu, accM, d); }
                    if (lane == 0) { sFlags[pair][0] = accP; sFlags[pair][1] = accM; }
                }
                pair_barrier(pair);  // block blk added, block blk + 1 staged
            }
            pair_barrier(pair);      // the consumer has read the flags of this run
        } else {
            const uint32_t wp = H[pos], wm = H[G + pos];
            float hp = __uint_as_float(wp & kValueMask), hm = __uint_as_float(wm & kValueMask);
            pair_barrier(pair);
            const size_t end = (size_t)sEnd[pair];
            if (end == ~(size_t)0) { pair_barrier(pair); continue; }  // handed over to k_fold_huge
            const size_t nBlocks = (end - i + kFoldBlock - 1) / kFoldBlock;
            for (size_t blk = 0; blk < nBlocks; ++blk) {
                const float4* qp = reinterpret_cast<const float4*>(stage[pair][blk & 1][0]);
                const float4* qm = reinterpret_cast<const float4*>(stage[pair][blk & 1][1]);
                const unsigned maskP = sMask[pair][blk & 1][0], maskM = sMask[pair][blk & 1][1];
                const unsigned left = (unsigned)min((size_t)kFoldBlock, end - (i + blk * kFoldBlock));
                if (left == (unsigned)kFoldBlock && (maskP == 0 || maskM == 0)) {  // a full block of one strand (the common case by far)
                    const bool isMinus = maskP == 0;
                    const float4* q = isMinus ? qm : qp;
                    float h = isMinus ? hm : hp;
#pragma unroll
                    for (int t = 0; t < kFoldBlock / 4; ++t) {
                        const float4 v = q[t];
                        h = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(h, v.x), v.y), v.z), v.w);
                    }
                    if (isMinus) hm = h; else hp = h;
                } else {  // both strands at this position (two independent chains), or the run ends inside this block
                    const int quads = (int)((left + 3) / 4);  // records past the end of the run were staged as zeros
#pragma unroll 4
                    for (int t = 0; t < quads; ++t) {
                        const float4 v = qp[t], x = qm[t];
                        hp = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(hp, v.x), v.y), v.z), v.w);
                        hm = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(hm, x.x), x.y), x.z), x.w);
                    }
                }
                pair_barrier(pair);
            }
            const unsigned flagsP = sFlags[pair][0], flagsM = sFlags[pair][1];
            if (lane == 0) {
                if (flagsP) {
                    H[pos] = (__float_as_uint(hp) & kValueMask) | (wp & kA10Bit) | ((flagsP & PPP_READ_A10) ? kA10Bit : 0u);
                    if (wp == 0u) atomicOr(&occ[pos >> 5], 1u << (pos & 31u));
                }
                if (flagsM) {
                    H[G + pos] = (__float_as_uint(hm) & kValueMask) | (wm & kA10Bit) | ((flagsM & PPP_READ_A10) ? kA10Bit : 0u);
                    if (wm == 0u) atomicOr(&occ[(G + pos) >> 5], 1u << (pos & 31u));
                }
            }
            pair_barrier(pair);
        }
    }
}

// One BLOCK per run of at least kHugeRun records.  The additions of :487 stay in file order, but they are not executed one
// after the other: inside the binade of the current height every addition is an integer map on the bit pattern of the
// height (fadd_chain.h), and maps compose.  Per trip of kHugeTrip records every lane composes the maps of its 32
// consecutive records (looked up per NH code in a table built for the current binade), a warp scan and a short serial
// walk over the warp totals give the height after the trip.  When the height would leave its binade in a trip, the lane
// in which that happens adds its 32 records for real and the next trip starts behind them, in the new binade (a height
// changes binade some 40 times in its life).  Both strands' chains of the position are carried side by side.
__global__ void __launch_bounds__(kHugeThreads) k_fold_huge(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t* __restrict__ H, uint64_t G,
                                                             uint32_t* __restrict__ occ, const unsigned long long* __restrict__ hugeRuns,
                                                             const unsigned int* __restrict__ nHuge) {
    constexpr int kWarps = kHugeThreads / 32;
    __shared__ float weightOf[kWeightLut];
    __shared__ AddMap sTab[2][kWeightLut];   // [strand][NH code]: the addition of that weight in the binade sTabExp[strand]
    __shared__ uint32_t sTabExp[2];
    __shared__ AddMap sTotal[2][kWarps];     // [strand][warp]: the warp's 1024 records composed
    __shared__ uint32_t sStart[2][kWarps];   // bit pattern of the height before the warp's records
    __shared__ uint32_t sBits[2];            // the heights (plus, minus) before the current trip
    __shared__ unsigned sFlags[2];
    __shared__ int sCross;                   // warp in which a height leaves its binade during this trip, or -1
    __shared__ unsigned long long sBase;
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < kWeightLut) weightOf[tid] = read_weight_weighted((uint32_t)tid << PPP_READ_NH_SHIFT);
    const unsigned int count = *nHuge;
    for (unsigned int run = blockIdx.x; run < count; run += gridDim.x) {
        const size_t i = (size_t)hugeRuns[2 * (size_t)run], end = (size_t)hugeRuns[2 * (size_t)run + 1];
        const uint32_t pos = keys[i];
        uint32_t wordP = 0, wordM = 0;
        unsigned accP = 0, accM = 0;
        __syncthreads();  // the previous run's shared state has been read
        if (tid == 0) {
            wordP = H[pos];
            wordM = H[G + pos];
            float hp = __uint_as_float(wordP & kValueMask), hm = __uint_as_float(wordM & kValueMask);
            size_t b = i;
            for (; (b & 3) != 0 && b < end; ++b) {  // up to three records in front of the first 16-byte boundary
                const uint32_t meta = vals[b];
                const float w = read_weight_weighted(meta);
                if (meta & PPP_READ_MINUS) { hm = __fadd_rn(hm, w); accM |= meta | 0x80000000u; }
                else { hp = __fadd_rn(hp, w); accP |= meta | 0x80000000u; }
            }
            sBits[0] = __float_as_uint(hp);
            sBits[1] = __float_as_uint(hm);
            sBase = b;
            sTabExp[0] = sTabExp[1] = ~0u;
            sFlags[0] = sFlags[1] = 0u;
        }
        __syncthreads();
        for (;;) {
            const size_t base = (size_t)sBase;
            if (base >= end) break;
            const uint32_t bitsP = sBits[0], bitsM = sBits[1];
            const uint32_t expP = bitsP >> 23, expM = bitsM >> 23;
            if (sTabExp[0] != expP || sTabExp[1] != expM) {  // (block-uniform) a new binade: the maps of the common weights
                __syncthreads();
                if (tid < kWeightLut) sTab[0][tid] = addmap_of(expP, __float_as_uint(weightOf[tid]));
                else if (tid < 2 * kWeightLut) sTab[1][tid - kWeightLut] = addmap_of(expM, __float_as_uint(weightOf[tid - kWeightLut]));
                if (tid == 0) { sTabExp[0] = expP; sTabExp[1] = expM; }
                __syncthreads();
            }
            // ---- this lane's 32 consecutive records -> one map per strand
            const size_t idx0 = base + (size_t)tid * kHugePerLane;
            AddMap mp = {0u, 0u}, mm = {0u, 0u};
            auto take = [&](uint32_t meta) {
                const uint32_t nh = meta >> PPP_READ_NH_SHIFT;
                const bool minus = (meta & PPP_READ_MINUS) != 0;
                AddMap m;
                if (nh < (uint32_t)kWeightLut) m = sTab[minus ? 1 : 0][nh];
                else m = addmap_of(minus ? expM : expP, __float_as_uint(read_weight_weighted(meta)));
                // (a run is nearly always of one strand: the branch is warp-uniform then)
                if (minus) { mm = addmap_compose(mm, m); accM |= meta | 0x80000000u; }   // A10 flag: :471
                else { mp = addmap_compose(mp, m); accP |= meta | 0x80000000u; }        // :483
            };
            if (idx0 + kHugePerLane <= end) {
                const uint4* src = reinterpret_cast<const uint4*>(vals + idx0);
                uint4 v[kHugePerLane / 4];
#pragma unroll
                for (int u = 0; u < kHugePerLane / 4; ++u) v[u] = __ldg(src + u);
#pragma unroll
                for (int u = 0; u < kHugePerLane / 4; ++u) { take(v[u].x); take(v[u].y); take(v[u].z); take(v[u].w); }
            } else {
                for (size_t j = idx0; j < end; ++j) take(vals[j]);
            }
            // ---- inclusive scan over the warp, warp totals, serial walk over the warps
            AddMap ip = mp, im = mm;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                AddMap qp, qm;
                qp.a0 = __shfl_up_sync(0xffffffffu, ip.a0, d); qp.a1 = __shfl_up_sync(0xffffffffu, ip.a1, d);
                qm.a0 = __shfl_up_sync(0xffffffffu, im.a0, d); qm.a1 = __shfl_up_sync(0xffffffffu, im.a1, d);
                if (lane >= (unsigned)d) { ip = addmap_compose(qp, ip); im = addmap_compose(qm, im); }
            }
            if (lane == 31) { sTotal[0][warp] = ip; sTotal[1][warp] = im; }
            __syncthreads();
            if (tid == 0) {
                uint32_t bp = bitsP, bm = bitsM;
                int cross = -1;
                for (int w = 0; w < kWarps; ++w) {
                    sStart[0][w] = bp;
                    sStart[1][w] = bm;
                    if (addmap_crosses(bp, sTotal[0][w]) || addmap_crosses(bm, sTotal[1][w])) { cross = w; break; }
                    bp = addmap_apply(bp, sTotal[0][w]);
                    bm = addmap_apply(bm, sTotal[1][w]);
                }
                sCross = cross;
                if (cross < 0) { sBits[0] = bp; sBits[1] = bm; sBase = base + kHugeTrip; }
            }
            __syncthreads();
            const int cross = sCross;
            if (cross >= 0 && (int)warp == cross) {
                const uint32_t bp = sStart[0][warp], bm = sStart[1][warp];
                const unsigned first = __ballot_sync(0xffffffffu, addmap_crosses(bp, ip) || addmap_crosses(bm, im));  // (monotone: a suffix of the lanes)
                const int lc = __ffs(first) - 1;
                // the maps of the lanes in front of lane lc are valid: its starting heights
                AddMap ep, em;
                ep.a0 = __shfl_up_sync(0xffffffffu, ip.a0, 1); ep.a1 = __shfl_up_sync(0xffffffffu, ip.a1, 1);
                em.a0 = __shfl_up_sync(0xffffffffu, im.a0, 1); em.a1 = __shfl_up_sync(0xffffffffu, im.a1, 1);
                if ((int)lane == lc) {
                    float hp = __uint_as_float(lane ? addmap_apply(bp, ep) : bp), hm = __uint_as_float(lane ? addmap_apply(bm, em) : bm);
                    const size_t stop = idx0 + kHugePerLane < end ? idx0 + kHugePerLane : end;
                    for (size_t j = idx0; j < stop; ++j) {  // real additions across the binade boundary
                        const uint32_t meta = vals[j];
                        const float w = read_weight_weighted(meta);
                        if (meta & PPP_READ_MINUS) hm = __fadd_rn(hm, w); else hp = __fadd_rn(hp, w);
                    }
                    sBits[0] = __float_as_uint(hp);
                    sBits[1] = __float_as_uint(hm);
                    sBase = idx0 + kHugePerLane;
                }
            }
            __syncthreads();
        }
        // ---- the run's A10 flags, then the two words
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { accP |= __shfl_xor_sync(0xffffffffu, accP, d); accM |= __shfl_xor_sync(0xffffffffu, accM, d); }
        if (lane == 0) { if (accP) atomicOr(&sFlags[0], accP); if (accM) atomicOr(&sFlags[1], accM); }
        __syncthreads();
        if (tid == 0) {
            const unsigned flagsP = sFlags[0], flagsM = sFlags[1];
            if (flagsP) {
                H[pos] = (sBits[0] & kValueMask) | (wordP & kA10Bit) | ((flagsP & PPP_READ_A10) ? kA10Bit : 0u);
                if (wordP == 0u) atomicOr(&occ[pos >> 5], 1u << (pos & 31u));
            }
            if (flagsM) {
                H[G + pos] = (sBits[1] & kValueMask) | (wordM & kA10Bit) | ((flagsM & PPP_READ_A10) ? kA10Bit : 0u);
                if (wordM == 0u) atomicOr(&occ[(G + pos) >> 5], 1u << (pos & 31u));
            }
        }
    }
}

// ------------------------------------------------------------------ small batches (-m weighted)
// A batch of at most kSmallBatch records (a small contig of a fragmented assembly, the last reads of a file) in ONE
// launch of one block instead of the ~25 launches of the radix sort + folds: the records' word indices are sorted in
// shared memory together with their index in the batch (bitonic network over 64-bit composites, hence stable), then
// one thread per run of equal words folds it in file order, continuing from the word's current value like the big path.
__global__ void __launch_bounds__(kBuildThreads) k_small_batch(const ppp_read* __restrict__ reads, uint32_t n, Slot slot, uint32_t* __restrict__ H, uint64_t G,
                                                                uint32_t* __restrict__ occ, unsigned long long* counters) {
    __shared__ unsigned long long item[kSmallBatch];  // (word index of H, both strands) << 12 | index in the batch
    constexpr unsigned long long kNone = ~0ull;
    uint32_t m = 1;
    while (m < n) m <<= 1;  // bitonic networks want a power of two; the padding sorts to the end
    for (uint32_t i = threadIdx.x; i < m; i += kBuildThreads) {
        unsigned long long v = kNone;
        if (i < n) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reads) + i);
            const ppp_read r = {raw.x, raw.y};
            uint32_t pos;
            if (locate<true>(r, slot, PPP_MULTIHITS_WEIGHTED, pos, counters))
                v = ((((r.meta & PPP_READ_MINUS) ? G : 0ull) + pos) << 12) | i;
        }
        item[i] = v;
    }
    __syncthreads();
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < m / 2; t += kBuildThreads) {
                const uint32_t a = ((t & ~(j - 1)) << 1) | (t & (j - 1)), b = a | j;  // the pair of this compare-exchange
                const bool up = (a & k) == 0;
                const unsigned long long x = item[a], y = item[b];
                if ((x > y) == up) { item[a] = y; item[b] = x; }
            }
            __syncthreads();
        }
    for (uint32_t j = threadIdx.x; j < n; j += kBuildThreads) {
        const unsigned long long v = item[j];
        if (v == kNone) continue;
        const unsigned long long word = v >> 12;
        if (j > 0 && (item[j - 1] >> 12) == word) continue;  // not a run head
        const uint32_t w0 = H[word];
        float h = __uint_as_float(w0 & kValueMask);
        uint32_t a10 = w0 & kA10Bit;
        for (uint32_t e = j; e < n && (item[e] >> 12) == word; ++e) {
            const uint32_t meta = __ldg(&reads[item[e] & 4095u].meta);
            h = __fadd_rn(h, read_weight_weighted(meta));  // :487 (every weight of this mode is positive, :458)
            if (meta & PPP_READ_A10) a10 = kA10Bit;         // :471 / :483
        }
        H[word] = (__float_as_uint(h) & kValueMask) | a10;
        if (w0 == 0u) atomicOr(&occ[word >> 5], 1u << (word & 31u));
    }
}

// 5-byte upload form -> ppp_read records (ppp_reads_submit_packed): meta bits 2..7 index the call's NH table
__global__ void __launch_bounds__(kBuildThreads) k_unpack_reads(const uint32_t* __restrict__ keys, const uint8_t* __restrict__ meta, const uint32_t* __restrict__ nhTable,
                                                                 size_t n, ppp_read* __restrict__ out) {
    __shared__ uint32_t sNh[PPP_PACKED_NH_CODES];
    if (threadIdx.x < PPP_PACKED_NH_CODES) sNh[threadIdx.x] = nhTable[threadIdx.x];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t m = meta[i];
        uint2 r;
        r.x = keys[i];
        r.y = (sNh[m >> 2] << PPP_READ_NH_SHIFT) | (m & 3u);
        reinterpret_cast<uint2*>(out)[i] = r;
    }
}

inline unsigned grid_for(size_t n, int threads, unsigned cap) {
    size_t b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    return (unsigned)(b < cap ? b : cap);
}

}  // namespace

cudaError_t launch_stack_build_int(const ppp_read* reads, size_t n, uint32_t* H, uint64_t G, uint32_t* occ, Slot slot, int mode,
                                   unsigned long long* counters, cudaStream_t s, int* launches) {
    if (n == 0) return cudaSuccess;
    k_stack_build_int<<<grid_for(n, kBuildThreads, 148 * 16), kBuildThreads, 0, s>>>(reads, n, H, G, occ, slot, mode, counters);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_unpack_reads(const uint32_t* keys, const uint8_t* meta, const uint32_t* nhTable, size_t n, ppp_read* out, cudaStream_t s, int* launches) {
    if (n == 0) return cudaSuccess;
    k_unpack_reads<<<grid_for(n, kBuildThreads * 4, 148 * 8), kBuildThreads, 0, s>>>(keys, meta, nhTable, n, out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_small_batch(const ppp_read* reads, size_t n, uint32_t* H, uint64_t G, uint32_t* occ, Slot slot, unsigned long long* counters,
                               cudaStream_t s, int* launches) {
    if (n == 0) return cudaSuccess;
    if (n > (size_t)kSmallBatch) return cudaErrorInvalidValue;
    k_small_batch<<<1, kBuildThreads, 0, s>>>(reads, (uint32_t)n, slot, H, G, occ, counters);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

size_t sort_block_hist_words(size_t n) { return ((n + kSortTile - 1) / kSortTile) * 256; }

cudaError_t launch_stack_sort(const ppp_read* reads, size_t n, uint64_t G, Slot slot, unsigned long long* counters, const SortScratch& scr,
                              cudaStream_t s, int* launches, SortedRun* out) {
    out->n = 0;
    if (n == 0) return cudaSuccess;
    if (n > scr.capacity) return cudaErrorInvalidValue;
    unsigned blocks = (unsigned)((n + kBuildThreads - 1) / kBuildThreads);
    k_sort_keys<<<blocks, kBuildThreads, 0, s>>>(reads, n, slot, scr.keysA, scr.valsA, counters);
    if (launches) *launches += 1;
    // key bits that can differ: positions < G, plus the all-ones invalid key
    int bits = 1;
    while (bits < 32 && (G >> bits) != 0) ++bits;
    uint32_t nBlocks = (uint32_t)((n + kSortTile - 1) / kSortTile);
    uint32_t *kin = scr.keysA, *vin = scr.valsA, *kout = scr.keysB, *vout = scr.valsB;
    for (int shift = 0; shift < 32; shift += 8) {
        // passes above the highest position bit only matter for the invalid key (all ones): its digits are 255
        // everywhere, so it stays behind every valid key as long as the top pass is executed.
        if (shift >= bits && shift + 8 < 32) continue;
        k_radix_hist<<<nBlocks, kSortThreads, 0, s>>>(kin, n, shift, scr.blockHist, nBlocks);
        cudaError_t e = exclusive_scan_u32(scr.blockHist, scr.blockHist, (size_t)nBlocks * 256, scr.scanScratch, scr.scanScratchWords, s, launches);
        if (e != cudaSuccess) return e;
        k_radix_scatter<<<nBlocks, kSortThreads, 0, s>>>(kin, vin, kout, vout, n, shift, scr.blockHist, nBlocks);
        if (launches) *launches += 2;
        uint32_t* t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    out->keys = kin;
    out->vals = vin;
    out->spareA = kout;
    out->spareB = vout;
    out->n = n;
    return cudaGetLastError();
}

cudaError_t launch_stack_fold(const SortedRun& r, uint32_t* H, uint64_t G, uint32_t* occ, size_t capacity, cudaStream_t s, int* launches) {
    if (r.n == 0) return cudaSuccess;
    unsigned blocks = (unsigned)((r.n + kBuildThreads - 1) / kBuildThreads);
    // the sort's output buffers of the last pass are free again: spareA holds the long-run list, spareB[0] its length
    unsigned long long* longRuns = reinterpret_cast<unsigned long long*>(r.spareA);
    unsigned int* nLong = reinterpret_cast<unsigned int*>(r.spareB);
    unsigned int longCap = (unsigned int)(capacity / 2);
    // spareB[1] = number of huge runs, their (first, end) pairs behind it (at most n / kHugeRun of them)
    unsigned int* nHuge = nLong + 1;
    unsigned long long* hugeRuns = reinterpret_cast<unsigned long long*>(r.spareB + 2);
    cudaError_t e = cudaMemsetAsync(nLong, 0, 2 * sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
    k_fold_runs<<<blocks, kBuildThreads, 0, s>>>(r.keys, r.vals, r.n, H, G, occ, longRuns, nLong, longCap);
    k_fold_long<<<148 * 2, kBuildThreads, 0, s>>>(r.keys, r.vals, r.n, H, G, occ, longRuns, nLong, longCap, hugeRuns, nHuge);
    k_fold_huge<<<148 * 2, kHugeThreads, 0, s>>>(r.keys, r.vals, H, G, occ, hugeRuns, nHuge);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

}  // namespace pppk
