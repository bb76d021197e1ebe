// scan.cu -- the pass over the stack store (K2 + K3 of SURVEY.md §8(a)).
//
// The stack store the pass reads is COMPACT (compact.cu builds it when the stack builder finishes): per strand the
// non-empty stack words in position order, an occupancy bitmap (1 bit per position) and the index of every tile's
// first stack.  Like the reference's std::map walk the work is proportional to the number of stacks, but every byte
// of it is a coalesced stream: G/4 bytes of bitmap + 4 bytes per stack + the signature records.
//
// k_scan (fused): ONE pass that
//   (K2) builds the height -> frequency histogram of mapHeightsToScores (pingpongpro.cpp:502-509): the compact words of
//        a tile are read densely, 256 consecutive words per step, whatever their positions;
//   (K3) finds, for every plus stack, the minus stacks at overlaps min..max (:559-577): the window test is one funnel
//        shift over two minus-strand bitmap words in shared memory, the rank of a position inside the tile (prefix
//        popcount) is the index of its word in the compact array, and the present stacks of a window are CONSECUTIVE
//        compact words; then the vicinity mean / max (:570-578), the local-height bin (:597-599), the base-bias bin
//        (:602) and one record per (plus stack, minus stack, overlap) (:608).
// Records leave the kernel in global (position, overlap) order, straight into their final place: tiles are claimed in
// order from a ticket counter, publish how many records they hold and find their first index by decoupled look-back
// (Merrill & Garland) over the tiles before them.  The overlap-`pingpong` records are written a second time into a
// stream of their own (16 B each) -- the scoring kernel (score.cu) reads nothing else.
// Everything of countStacksByGroup that needs the finished histogram (the height-score bin :589-594 and the grouped
// counter :605) is deferred to k_bin, which reads 9 bytes and writes 2 bytes per record.
//
// Bound: HBM streaming.  Algorithmic bytes per launch: G/4 (bitmaps) + 4 S (stack words) + 13 P (records) + 16 P10
// (overlap-`pingpong` stream) + 8 per tile (compact offsets); k_bin 11 P.  (SURVEY §8(d)'s dense accounting of the same
// work is 16 G + 36 P.)
#include <algorithm>

#include "common.cuh"
#include "log10f_emul.h"

namespace pppk {

namespace {

constexpr int kThreads = 128;                  // small blocks, many per SM: a tile is a chain of latencies (ticket, bitmap load, look-back), hidden by other tiles
constexpr int kWarps = kThreads / 32;
constexpr int kWords = 8;                      // consecutive 32-position occupancy words owned by one thread, per strand
constexpr int kTile = kScanTile;               // positions per tile
constexpr int kTileWords = kTile / 32;         // occupancy words per tile and strand
static_assert(kTile == kThreads * kWords * 32, "tile geometry");
static_assert(kTile <= (1 << 15) && PPP_MAX_OVERLAPS <= 32, "prefix arrays are 16 bit; a window fits the 64 positions after a word's first");

__device__ __forceinline__ unsigned long long warp_inclusive64(unsigned long long v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) { return *reinterpret_cast<const volatile unsigned long long*>(p); }
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) { *reinterpret_cast<volatile unsigned long long*>(p) = v; }

// Decoupled look-back, executed by one whole warp: the packed sum of the aggregates of all tiles before `tile`.  A
// tile's state is ONE 64-bit word (flag + value), so a plain volatile load sees a consistent pair and no fence is
// needed.  Tiles are claimed in ticket order, hence every earlier tile belongs to a block that is already running --
// and a block publishes a tile's aggregate within microseconds of claiming it.  A step covers 128 tiles (four
// independent loads per lane): the resident blocks tend to run in step, so a tile may have to walk over a whole wave.
__device__ __forceinline__ unsigned long long lookback_exclusive(const unsigned long long* __restrict__ state, uint32_t tile, int lane) {
    unsigned long long sum = 0;
    int64_t j = (int64_t)tile - 1;  // the nearest tile of the current window
    while (j >= 0) {
        unsigned long long v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t mine = j - 4 * lane - k;
            v[k] = mine >= 0 ? ld_state(state + mine) : kLbPrefix;  // before the first tile: an inclusive prefix of zero
        }
        unsigned long long part = 0;
        bool closed = false;  // this lane met a tile that knows its inclusive prefix
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t mine = j - 4 * lane - k;
            while ((v[k] >> 62) == 0ull) v[k] = ld_state(state + mine);
            if (!closed) part += v[k] & kLbValueMask;
            closed = closed || (v[k] >> 62) == 2ull;
        }
        const unsigned closedLanes = __ballot_sync(0xffffffffu, closed);
        const int stop = closedLanes ? __ffs(closedLanes) - 1 : 32;
        if (lane > stop) part = 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
        sum += part;
        if (closedLanes) break;
        j -= 128;
    }
    return sum;
}

// a stack taller than the shared-memory bins (rare): out of line, so that it costs the streaming loop no registers
__device__ __noinline__ void count_tall(uint32_t* freqTail, uint32_t* tallList, unsigned int* tallCount, uint32_t tallCap, uint32_t* maxTail, uint32_t r) {
    if (r >= kFreqCapacity) r = kFreqCapacity - 1;  // (+inf of a malformed NH:i:0 record: never index past the table)
    if (tallList) {
        unsigned int i = atomicAdd(tallCount, 1u);
        if (i < tallCap) tallList[i] = r;
    } else {
        atomicMax(maxTail, atomicAdd(&freqTail[r], 1u) + 1u);
    }
}

// 4-byte asynchronous copy global -> shared (LDGSTS): the tile's occupancy words travel without occupying registers
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// per-thread counts of a tile in one 64-bit word, so that ONE warp scan carries all four:
//   bits 0..18 records, 19..32 ping-pong records, 33..46 plus stacks, 47..60 minus stacks
// (a warp owns 32 * kWords * 32 positions: totals <= 2^18 records, 2^13 of the others)
constexpr int kPkPp = 19, kPkPlus = 33, kPkMinus = 47;
static_assert(32 * kWords * 32 * PPP_MAX_OVERLAPS < (1 << kPkPp) && 32 * kWords * 32 < (1 << (kPkPlus - kPkPp)), "packed warp-scan fields");
__device__ __forceinline__ uint32_t pk_rec(unsigned long long v) { return (uint32_t)v & 0x7ffffu; }
__device__ __forceinline__ uint32_t pk_pp(unsigned long long v) { return (uint32_t)(v >> kPkPp) & 0x3fffu; }
__device__ __forceinline__ uint32_t pk_plus(unsigned long long v) { return (uint32_t)(v >> kPkPlus) & 0x3fffu; }
__device__ __forceinline__ uint32_t pk_minus(unsigned long long v) { return (uint32_t)(v >> kPkMinus) & 0x3fffu; }

// the minus-strand occupancy bits at distance sh.. of a position whose word and the following one are (w0, w1), sh < 64
__device__ __forceinline__ uint32_t window_bits(uint32_t w0, uint32_t w1, uint32_t sh, uint32_t wMask) {
    return (sh < 32u ? __funnelshift_r(w0, w1, sh) : (w1 >> (sh - 32u))) & wMask;
}

template <int REPR>
__global__ void __launch_bounds__(kThreads, 10) k_scan(ScanParams P) {
    __shared__ uint32_t sMask[kTileWords + 2];     // minus-strand occupancy words of the tile + the word after the tile + a zero
    __shared__ uint32_t sPlus[kTileWords];         // plus strand
    __shared__ uint32_t sOff[4];                   // compact index of the tile's first / past-the-last plus stack, first / past-the-last minus stack
    // per occupancy word, what the tile holds BEFORE the word's positions:
    __shared__ uint32_t sRecOff[kTileWords + 1];   // records (last entry: of the whole tile)
    __shared__ uint16_t sPpOff[kTileWords];        // ping-pong records
    __shared__ uint16_t sPrefP[kTileWords];        // plus stacks
    __shared__ uint16_t sPrefM[kTileWords + 2];    // minus stacks (entry kTileWords: the tile's total = rank of the halo word)
    __shared__ unsigned long long sWarp[kWarps];   // per warp: packed totals
    __shared__ unsigned long long sAgg, sBase;     // packed (kLb* layout) records of the tile / before the tile
    __shared__ uint32_t sHist[kScanLowBins];
    __shared__ uint32_t sTile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < (int)kScanLowBins; k += kThreads) sHist[k] = 0;
    if (tid == 0) sMask[kTileWords + 1] = 0;

    const uint32_t ONE = REPR == kReprInt ? 1u : 0x3f800000u;
    const uint32_t wMask = P.W >= 32 ? 0xffffffffu : ((1u << P.W) - 1u);
    const float fW = (float)P.W;
    const uint32_t myWord = (uint32_t)tid * kWords;  // this thread owns kWords consecutive occupancy words of a tile
    uint32_t ones = 0, localMax = 0, small = 0;  // small: four 8-bit counters for rounded heights 0..3
    unsigned long long found = 0;                // (thread 0) packed records this block found
    auto flush_small = [&]() {
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if ((small >> (8 * r)) & 0xffu) atomicAdd(&sHist[r], (small >> (8 * r)) & 0xffu);
        small = 0;
    };
    auto classify = [&](uint32_t w) {  // K2: freq[(unsigned)(0.5 + reads)] += 1 (:508)
        if ((w & kValueMask) == ONE) { ++ones; return; }
        const uint32_t r = word_round_height<REPR>(w);
        if (r < 4u) { small += 1u << (8 * r); if ((small & 0x80808080u) != 0u) flush_small(); }  // heights 0..3: byte counters in a register
        else if (r < kScanLowBins) atomicAdd(&sHist[r], 1u);
        else count_tall(P.freqTail, P.tallList, P.tallCount, P.tallCap, P.maxTail, r);
        localMax = max(localMax, r);
    };

    // A tile is claimed only when the block is ready to count and publish it at once: whoever holds an earlier ticket
    // publishes its record counts within microseconds, whatever its record phase takes afterwards.  (Claiming ahead, to
    // prefetch the occupancy words, made every block wait for the slowest record phase of the previous round.)
    for (;;) {
        __syncthreads();  // the previous tile's shared state is no longer read
        if (tid == 0) sTile = atomicAdd(&P.state->ticket, 1u);
        __syncthreads();
        const uint32_t tile = sTile;
        if (tile >= P.nTiles) break;
        const uint32_t tileBase = tile * (uint32_t)kTile;
        {   // the tile's occupancy words and compact offsets, global -> shared, asynchronously
            const size_t w = (size_t)tile * kTileWords;
#pragma unroll
            for (int k = 0; k < kWords; ++k) {
                cp_async4(&sPlus[k * kThreads + tid], P.occ + w + k * kThreads + tid);
                cp_async4(&sMask[k * kThreads + tid], P.occ + P.G / 32 + w + k * kThreads + tid);
            }
            if (tid == 0) cp_async4(&sMask[kTileWords], P.occ + P.G / 32 + w + kTileWords);
            if (tid >= 32 && tid < 36) cp_async4(&sOff[tid - 32], ((tid & 2) ? P.tileOffM : P.tileOffP) + tile + (tid & 1));
            cp_async_commit();
            cp_async_wait();
        }
        __syncthreads();  // (A) the tile's occupancy words are complete
        {   // whoever claims the tile one grid ahead finds its occupancy words in L2 (one 128-byte line per thread)
            const uint32_t ahead = tile + gridDim.x;
            if (ahead < P.nTiles && tid < 2 * (kTileWords / 32)) {
                const uint32_t* line = P.occ + (tid < kTileWords / 32 ? 0 : P.G / 32) + (size_t)ahead * kTileWords + (size_t)(tid % (kTileWords / 32)) * 32;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
            }
        }

        // ---- counting phase (bitmaps only): how many records the tile holds, published at once, so that the blocks
        // holding later tiles never wait for this block's record phase
        // (per-word counts wait in the shared-memory arrays that will hold their prefixes: no register arrays)
        uint32_t sumRec = 0, sumPp = 0, sumPlus = 0, sumMinus = 0;
        {
            uint32_t mw = sMask[myWord];
#pragma unroll 2
            for (int k = 0; k < kWords; ++k) {
                const uint32_t pw = sPlus[myWord + k], w0 = mw, w1 = sMask[myWord + k + 1];
                mw = w1;
                uint32_t cnt = 0, npp = 0;
                for (uint32_t nz = pw; nz; nz &= nz - 1) {  // a plus stack has one record per minus stack at overlaps min..max (:559-577)
                    const uint32_t bits = window_bits(w0, w1, (uint32_t)(__ffs(nz) - 1) + (uint32_t)P.minOv, wMask);
                    cnt += __popc(bits);
                    npp += (bits >> P.ppIdx) & 1u;
                }
                sRecOff[myWord + k] = cnt;
                sPpOff[myWord + k] = (uint16_t)npp;
                sumRec += cnt;
                sumPp += npp;
                sumPlus += (uint32_t)__popc(pw);
                sumMinus += (uint32_t)__popc(w0);
            }
        }
        // one scan over the warp for all four counts; records are ordered (word, position, overlap) == (position, overlap)
        unsigned long long ex;
        {
            const unsigned long long pack = (unsigned long long)sumRec | ((unsigned long long)sumPp << kPkPp) | ((unsigned long long)sumPlus << kPkPlus) |
                                            ((unsigned long long)sumMinus << kPkMinus);
            const unsigned long long inc = warp_inclusive64(pack, lane);
            if (lane == 31) sWarp[warp] = inc;
            ex = inc - pack;
        }
        __syncthreads();  // (B)
        {   // complete the counts before this thread's words over the tile, word by word
            uint32_t rec = pk_rec(ex), pp = pk_pp(ex), plus = pk_plus(ex), minus = pk_minus(ex), totRec = 0, totPp = 0, totMinus = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const unsigned long long t = sWarp[w];
                if (w < warp) { rec += pk_rec(t); pp += pk_pp(t); plus += pk_plus(t); minus += pk_minus(t); }
                totRec += pk_rec(t);
                totPp += pk_pp(t);
                totMinus += pk_minus(t);
            }
#pragma unroll 2
            for (int k = 0; k < kWords; ++k) {
                const uint32_t cnt = sRecOff[myWord + k], npp = sPpOff[myWord + k];
                sRecOff[myWord + k] = rec;
                sPpOff[myWord + k] = (uint16_t)pp;
                sPrefP[myWord + k] = (uint16_t)plus;
                sPrefM[myWord + k] = (uint16_t)minus;
                rec += cnt;
                pp += npp;
                plus += (uint32_t)__popc(sPlus[myWord + k]);
                minus += (uint32_t)__popc(sMask[myWord + k]);
            }
            if (tid == 0) {  // the tile's aggregate, published before any of its records is written
                const unsigned long long agg = (unsigned long long)totRec | ((unsigned long long)totPp << kLbRecBits);
                sAgg = agg;
                sRecOff[kTileWords] = totRec;
                sPrefM[kTileWords] = (uint16_t)totMinus;
                sPrefM[kTileWords + 1] = (uint16_t)totMinus;
                st_state(P.tileState + tile, kLbAggregate | agg);
            }
        }
        __syncthreads();  // (C) the aggregate and the per-word counts are visible

        // ---- K2: height histogram, dense over the tile's compact words (halo copies of a neighbouring shard are not counted)
        {
            const uint32_t offP = sOff[0], endP = sOff[1], offM = sOff[2], endM = sOff[3];
            for (uint32_t i = offP + tid; i < endP; i += kThreads * 2) {
                const uint32_t w0 = __ldg(P.cmpP + i);
                const uint32_t w1 = i + kThreads < endP ? __ldg(P.cmpP + i + kThreads) : 0u;
                classify(w0);
                if (w1) classify(w1);
            }
            uint32_t skipLo = 0xffffffffu, skipHi = 0xffffffffu;  // ranks (within the tile) of the halo copies
            if (tileBase < P.haloHi && tileBase + kTile > P.haloLo) {
                auto rankOf = [&](uint32_t q) {  // minus stacks of the tile at tile positions < q
                    if (q >= (uint32_t)kTile) return (uint32_t)sPrefM[kTileWords];
                    return (uint32_t)sPrefM[q >> 5] + (uint32_t)__popc(sMask[q >> 5] & ((1u << (q & 31)) - 1u));
                };
                skipLo = rankOf(P.haloLo > tileBase ? P.haloLo - tileBase : 0u);
                skipHi = rankOf(P.haloHi - tileBase);
            }
            for (uint32_t i = offM + tid; i < endM; i += kThreads * 2) {
                const uint32_t r0 = i - offM, r1 = r0 + kThreads;
                const uint32_t w0 = (r0 >= skipLo && r0 < skipHi) ? 0u : __ldg(P.cmpM + i);
                const uint32_t w1 = (i + kThreads < endM && !(r1 >= skipLo && r1 < skipHi)) ? __ldg(P.cmpM + i + kThreads) : 0u;
                if (w0) classify(w0);
                if (w1) classify(w1);
            }
        }

        // where the tile's records go: look back over the tiles before it (by now they have published)
        if (warp == 0) {
            const unsigned long long before = lookback_exclusive(P.tileState, tile, lane);
            if (lane == 0) {
                st_state(P.tileState + tile, kLbPrefix | ((before + sAgg) & kLbValueMask));
                found += sAgg;
                sBase = before;
            }
        }
        const uint32_t total = sRecOff[kTileWords];
        if (!total) continue;  // (block-uniform)
        __syncthreads();  // sBase is visible

        // ---- record phase: ONE THREAD PER RECORD (dense, whatever the stacks' window sizes).  Record r of the tile belongs
        // to the last occupancy word whose record offset is <= r (binary search over the per-word offsets), inside the word
        // to the plus stack whose windows cover it, and to the k-th present minus stack of that window.
        const unsigned long long recBase = sBase & kLbRecMask, ppBase = sBase >> kLbRecBits;
        const uint32_t offP = sOff[0], offM = sOff[2];
        for (uint32_t r = tid; r < total; r += kThreads) {
            uint32_t word = 0;
#pragma unroll
            for (int step = kTileWords / 2; step > 0; step >>= 1)
                if (sRecOff[word + step] <= r) word += step;
            const uint32_t pw = sPlus[word], w0 = sMask[word], w1 = sMask[word + 1];
            uint32_t left = r - sRecOff[word], b = 0, bits = 0;
            for (uint32_t nz = pw; nz; nz &= nz - 1) {  // (a word with records has a plus stack that covers `left`)
                b = (uint32_t)(__ffs(nz) - 1);
                bits = window_bits(w0, w1, b + (uint32_t)P.minOv, wMask);
                const uint32_t n = (uint32_t)__popc(bits);
                if (left < n) break;
                left -= n;
            }
            const int k = (int)left;                              // the record is the k-th present stack of the window
            const uint32_t tp = word * 32u + b;
            const uint32_t wp = __ldg(P.cmpP + offP + sPrefP[word] + __popc(pw & ((1u << b) - 1u)));
            // the present stacks of the window are consecutive compact words, starting at the rank of its first position
            const uint32_t q = tp + (uint32_t)P.minOv, qw = q >> 5;
            const uint32_t* win = P.cmpM + offM + sPrefM[qw] + __popc(sMask[qw] & ((1u << (q & 31u)) - 1u));
            const int n = __popc(bits);
            // vicinity sum / max over the PRESENT stacks only, in overlap order: an absent stack would add
            // +0.0f, which is an exact no-op (:564-577)
            float sum = 0.0f, mx = 0.0f;
            uint32_t wm = 0, o = 0, rest = bits;
            for (int j = 0; j < n; ++j, rest &= rest - 1) {
                const uint32_t w = __ldg(win + j);
                const float mq = word_height<REPR>(w);
                sum = __fadd_rn(sum, mq);
                mx = fmaxf(mx, mq);
                if (j == k) { wm = w; o = (uint32_t)(__ffs(rest) - 1); }  // ... at overlap index o
            }
            const float mean = __fdiv_rn(sum, fW);                    // :578
            const float hp = word_height<REPR>(wp), mo = word_height<REPR>(wm);
            const float local = __fdiv_rn(__fsub_rn(mo, __fsub_rn(mean, __fdiv_rn(mo, fW))), mx);  // :597
            const uint32_t lbin = local < 0.2f ? 1u : 0u;               // :599 ((double)local < 0.2: no float lies in [0.2, 0.2f))
            const uint32_t bbin = ((wp | wm) & kA10Bit) ? 0u : 1u;     // :602
            const uint32_t misc = o | (lbin << 5) | (bbin << 6);
            const unsigned long long idx = recBase + r;
            if (idx < P.capacity) {                                    // :608
                P.recPos[idx] = tileBase + tp;
                P.recRp[idx] = hp;
                P.recRm[idx] = mo;
                P.recMisc[idx] = (uint8_t)misc;
            }
            if ((int)o == P.ppIdx) {
                // its index in the ping-pong stream: ping-pong records before the word + those of the word's earlier positions
                const uint32_t ppMask = pw & window_bits(w0, w1, (uint32_t)(P.minOv + P.ppIdx), 0xffffffffu);
                const unsigned long long j = ppBase + sPpOff[word] + __popc(ppMask & ((1u << b) - 1u));
                if (j < P.ppCapacity) {
                    PairRec rr;
                    rr.lpos = tileBase + tp;
                    rr.rp = hp;
                    rr.rm = mo;
                    rr.misc = misc;
                    *reinterpret_cast<uint4*>(&P.ppRecs[j]) = *reinterpret_cast<uint4*>(&rr);
                }
            }
        }
    }

    // flush the privatised histogram and this block's record counts
    flush_small();
    if (ones) { atomicAdd(&sHist[1], ones); localMax = max(localMax, 1u); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) localMax = max(localMax, __shfl_xor_sync(0xffffffffu, localMax, d));
    if (lane == 0 && localMax) atomicMax(P.maxHeight, localMax);
    if (tid == 0 && found) { atomicAdd(&P.state->totals[0], found & kLbRecMask); atomicAdd(&P.state->totals[1], found >> kLbRecBits); }
    __syncthreads();
    for (int k = tid; k < (int)kScanLowBins; k += kThreads) {
        uint32_t c = sHist[k];
        if (c) atomicAdd(&P.freqLow[k], (unsigned long long)c);
    }
}

// see launch_tall_apply in common.cuh
__global__ void __launch_bounds__(256) k_tall_apply(const uint32_t* __restrict__ gather, int world, uint32_t cap, unsigned long long* __restrict__ freqLow,
                                                    uint32_t* __restrict__ freqTail, uint32_t* __restrict__ maxHeight, uint32_t* __restrict__ overflow,
                                                    uint32_t* __restrict__ maxTail) {
    const size_t slotWords = (size_t)kTallHeader + cap;
    uint32_t best = 0;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < kScanLowBins; k += gridDim.x * blockDim.x) {
        unsigned long long sum = 0;
        for (int r = 0; r < world; ++r) sum += reinterpret_cast<const unsigned long long*>(gather + (size_t)r * slotWords + kTallLow)[k];
        freqLow[k] = sum;
    }
    for (int r = 0; r < world; ++r) {
        const uint32_t* slot = gather + (size_t)r * slotWords;
        uint32_t n = slot[0];
        best = max(best, slot[1]);
        if (n > cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(overflow, n); n = cap; }  // (the longest list: the host sizes the next attempt from it)
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
            atomicMax(maxTail, atomicAdd(&freqTail[slot[kTallHeader + i]], 1u) + 1u);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *maxHeight = best;
}

// The largest count of freqTail[kScanLowBins .. maxHeight] -- only where the tail counts were all-reduced as dense tables
// (the counting kernels track it themselves otherwise).
__global__ void __launch_bounds__(256) k_tail_max(const uint32_t* __restrict__ freqTail, const uint32_t* __restrict__ maxHeight, uint32_t* __restrict__ maxTail) {
    const uint32_t maxH = *maxHeight;
    uint32_t best = 0;
    for (uint64_t h = (uint64_t)kScanLowBins + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h <= maxH; h += (uint64_t)gridDim.x * blockDim.x) best = max(best, freqTail[h]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, d));
    if ((threadIdx.x & 31) == 0 && best) atomicMax(maxTail, best);
}

// The bin thresholds of k_bin / k_score from the (global) maximum height frequency, on the device (log10f_emul.h): one
// thread per threshold.  The host recomputes the table with its libm while the device works on, and compares (api.cu).
__global__ void __launch_bounds__(32) k_bin_thresholds(PassState* __restrict__ state, float* __restrict__ thresholds, BinParams* __restrict__ params) {
    // the largest height frequency (:547-550): every block finds it for itself (512 low bins + the tracked maximum of the tail)
    unsigned long long c = state->maxTail;
    {
        unsigned long long v[kScanLowBins / 32];
#pragma unroll
        for (int k = 0; k < (int)kScanLowBins / 32; ++k) v[k] = state->freqLow[k * 32 + threadIdx.x];  // (all loads in flight together)
#pragma unroll
        for (int k = 0; k < (int)kScanLowBins / 32; ++k) c = v[k] > c ? v[k] : c;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, c, d); c = o > c ? o : c; }
    if (blockIdx.x == 0 && threadIdx.x == 0) state->maxCount = c;
    const float maxFreq = (float)(c < kSaturate ? (uint32_t)c : kSaturate);  // the float counter of :508, :547-550
    const int b = blockIdx.x * 32 + threadIdx.x;  // threshold of bin b (1..999) lives at [b - 1]; one warp per SM: the search is a dependent chain
    if (b == 0) {
        params->maxFreq = maxFreq;
        params->binScale = maxFreq > 1.0f ? __fdiv_rn(999.0f, emul_log10f(__fmul_rn(maxFreq, maxFreq))) : 0.0f;
        thresholds[kBins - 1] = __uint_as_float(0x7f800000u);
    } else if (b < kBins) {
        thresholds[b - 1] = maxFreq > 1.0f ? emul_bin_threshold(b, maxFreq) : __uint_as_float(0x7f800000u);
    }
}

// K3 part 2: height-score bin (:589-594) + grouped counter (:605) of every signature record.
// The bin of a height-score product is the number of host-built thresholds <= hs (see api.cu).  A device log10
// approximation proposes the bin, two short loops against the exact thresholds correct it, so the result is the
// exact step function whatever the approximation error.  Pairs of stacks with rounded heights below 4 (the bulk:
// singletons) take their bin from a 16-entry table evaluated once per block.
// The grouped counters are concentrated on few cells: those of the (height 1, height 1) bin are counted in per-warp
// private shared-memory tables (index = the record's 7 misc bits), all others in a direct-mapped shared-memory cache of
// (cell, count) slots; a record whose slot is taken by another cell goes straight to the global (L2-resident) table.
constexpr int kBinThreads = 512;
constexpr int kBinWarps = kBinThreads / 32;
constexpr int kBinCache = 4096;
constexpr uint32_t kBinEmpty = 0xffffffffu;
__global__ void __launch_bounds__(kBinThreads) k_bin(const float* __restrict__ recRp, const float* __restrict__ recRm, const uint8_t* __restrict__ recMisc,
                                                     uint16_t* __restrict__ recBin, const unsigned long long* __restrict__ totals, uint64_t capacity,
                                                     FreqTables freq, const float* __restrict__ thresholds, const BinParams* __restrict__ params, int W,
                                                     uint32_t* __restrict__ grouped) {
    __shared__ float sThr[kBins];
    __shared__ uint32_t sKey[kBinCache], sVal[kBinCache];
    __shared__ uint32_t sOnes[kBinWarps][128];
    __shared__ uint16_t sSmallBin[16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < kBins - 1; k += kBinThreads) sThr[k] = thresholds[k];
    for (int k = tid; k < kBinCache; k += kBinThreads) { sKey[k] = kBinEmpty; sVal[k] = 0; }
    for (int k = tid; k < kBinWarps * 128; k += kBinThreads) (&sOnes[0][0])[k] = 0;
    const float binScale = params->binScale;
    __syncthreads();
    auto bin_of = [&](float hs) {
        int b = (int)(__log10f(hs) * binScale + 0.5f);             // proposal only
        b = min(max(b, 0), kBins - 1);
        while (b > 0 && sThr[b - 1] > hs) --b;                    // exact: bin = #thresholds <= hs
        while (b < kBins - 1 && sThr[b] <= hs) ++b;
        return b;
    };
    if (tid < 16) sSmallBin[tid] = (uint16_t)bin_of(__fmul_rn(freq_of(freq, tid >> 2), freq_of(freq, tid & 3)));  // (a height that never occurs has frequency 0: never looked up)
    __syncthreads();
    const uint32_t onesBin = sSmallBin[5];  // rounded heights (1, 1)
    unsigned long long n = totals[0];
    if (n > capacity) n = capacity;
    constexpr int kUnroll = 4;  // records in flight per thread: the loop is bound by the latency of its three loads
    for (size_t i0 = (size_t)blockIdx.x * kBinThreads * kUnroll + tid; i0 < n; i0 += (size_t)gridDim.x * kBinThreads * kUnroll) {
        float rp[kUnroll], rm[kUnroll];
        uint32_t mi[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const size_t i = i0 + (size_t)u * kBinThreads;
            const bool in = i < n;
            rp[u] = in ? recRp[i] : 0.0f;
            rm[u] = in ? recRm[i] : 0.0f;
            mi[u] = in ? (uint32_t)recMisc[i] : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (mi[u] == 0xffffffffu) continue;
            const uint32_t misc = mi[u];
            const uint32_t hp = round_height(rp[u]), hm = round_height(rm[u]);  // heightScoreMap[0.5 + reads], :582 / :589
            uint32_t b;
            if ((hp | hm) < 4u) b = sSmallBin[hp * 4 + hm];
            else b = (uint32_t)bin_of(__fmul_rn(freq_of(freq, hp), freq_of(freq, hm)));
            recBin[i0 + (size_t)u * kBinThreads] = (uint16_t)b;
            if (b == onesBin) { atomicAdd(&sOnes[warp][misc & 127u], 1u); continue; }
            const uint32_t cell = (((b * (uint32_t)W + (misc & 31u)) * 2u + ((misc >> 6) & 1u)) * 2u) + ((misc >> 5) & 1u);  // bin-major table, :605
            const uint32_t h = (cell * 2654435761u) >> 20;
            uint32_t key = sKey[h];
            if (key == kBinEmpty) key = atomicCAS(&sKey[h], kBinEmpty, cell);
            if (key == kBinEmpty || key == cell) atomicAdd(&sVal[h], 1u);
            else atomicAdd(&grouped[cell], 1u);
        }
    }
    __syncthreads();
    for (int k = tid; k < kBinCache; k += kBinThreads)
        if (sKey[k] != kBinEmpty && sVal[k]) atomicAdd(&grouped[sKey[k]], sVal[k]);
    if (tid < 128) {
        uint32_t c = 0;
        for (int w = 0; w < kBinWarps; ++w) c += sOnes[w][tid];
        if (c) {
            const uint32_t misc = (uint32_t)tid;
            atomicAdd(&grouped[(((onesBin * (uint32_t)W + (misc & 31u)) * 2u + ((misc >> 6) & 1u)) * 2u) + ((misc >> 5) & 1u)], c);
        }
    }
    (void)lane;
}

// Multi-GPU: this rank's "my signature store was too small" as a 0/1 word behind the grouped counts (summed by their all-reduce)
__global__ void k_overflow_flag(const unsigned long long* __restrict__ totals, uint64_t capacity, uint64_t ppCapacity, uint32_t* __restrict__ flag) {
    *flag = (totals[0] > capacity || totals[1] > ppCapacity) ? 1u : 0u;
}

// Test / writer / K5 view of the records: array-of-structures PairRec with the bin in misc bits 7..16 and the FDR of every
// record (:750-753), materialised on demand (ppp_pairs_download, ppp_transposons) -- never inside a scan + score pass.
__global__ void __launch_bounds__(256) k_pairs_view(const uint32_t* __restrict__ recPos, const float* __restrict__ recRp, const float* __restrict__ recRm,
                                                    const uint8_t* __restrict__ recMisc, const uint16_t* __restrict__ recBin, size_t n,
                                                    const float* __restrict__ fdrTable, const uint16_t* __restrict__ binMap, int W, PairRec* __restrict__ pairs,
                                                    float* __restrict__ pairFdr) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t misc = recMisc[i], bin = recBin[i];
    PairRec r;
    r.lpos = recPos[i];
    r.rp = recRp[i];
    r.rm = recRm[i];
    r.misc = misc | (bin << 7);
    *reinterpret_cast<uint4*>(&pairs[i]) = *reinterpret_cast<uint4*>(&r);
    if (fdrTable) pairFdr[i] = fdrTable[((size_t)binMap[bin] * W + (misc & 31u)) * 4 + ((misc >> 6) & 1u) * 2 + ((misc >> 5) & 1u)];  // :682, :753
}

}  // namespace

cudaError_t launch_scan(const ScanParams& p, int repr, cudaStream_t s, int* launches) {
    if (p.nTiles == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int perSm = 10;
    if (repr == kReprInt) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_scan<kReprInt>, kThreads, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_scan<kReprFloat>, kThreads, 0);
    if (perSm < 1) perSm = 1;
    unsigned grid = (unsigned)sms * (unsigned)perSm;
    if (grid > p.nTiles) grid = p.nTiles;
    if (repr == kReprInt) k_scan<kReprInt><<<grid, kThreads, 0, s>>>(p);
    else k_scan<kReprFloat><<<grid, kThreads, 0, s>>>(p);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_tall_apply(const uint32_t* gather, int world, uint32_t cap, unsigned long long* freqLow, uint32_t* freqTail, uint32_t* maxHeight,
                              uint32_t* overflow, uint32_t* maxTail, cudaStream_t s, int* launches) {
    k_tall_apply<<<148, 256, 0, s>>>(gather, world, cap, freqLow, freqTail, maxHeight, overflow, maxTail);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_tail_max(const uint32_t* freqTail, const uint32_t* maxHeight, uint32_t* maxTail, cudaStream_t s, int* launches) {
    k_tail_max<<<148 * 4, 256, 0, s>>>(freqTail, maxHeight, maxTail);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_bin_thresholds(PassState* state, float* thresholds, BinParams* params, cudaStream_t s, int* launches) {
    k_bin_thresholds<<<(kBins + 31) / 32, 32, 0, s>>>(state, thresholds, params);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_bin(const float* recRp, const float* recRm, const uint8_t* recMisc, uint16_t* recBin, const unsigned long long* totals, uint64_t capacity,
                       uint64_t expected, FreqTables freq, const float* thresholds, const BinParams* params, int W, uint32_t* grouped, cudaStream_t s, int* launches) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t want = (expected + kBinThreads * 16 - 1) / (kBinThreads * 16);  // >= 16 records per thread: the tables are flushed once per block
    unsigned blocks = (unsigned)std::min<uint64_t>(std::max<uint64_t>(want, 1), (uint64_t)sms * 4);
    k_bin<<<blocks, kBinThreads, 0, s>>>(recRp, recRm, recMisc, recBin, totals, capacity, freq, thresholds, params, W, grouped);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_overflow_flag(const unsigned long long* totals, uint64_t capacity, uint64_t ppCapacity, uint32_t* flag, cudaStream_t s, int* launches) {
    k_overflow_flag<<<1, 1, 0, s>>>(totals, capacity, ppCapacity, flag);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_pairs_view(const uint32_t* recPos, const float* recRp, const float* recRm, const uint8_t* recMisc, const uint16_t* recBin, size_t n,
                              const float* fdrTable, const uint16_t* binMap, int W, PairRec* pairs, float* pairFdr, cudaStream_t s, int* launches) {
    if (n == 0) return cudaSuccess;
    k_pairs_view<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(recPos, recRp, recRm, recMisc, recBin, n, fdrTable, binMap, W, pairs, pairFdr);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace pppk
