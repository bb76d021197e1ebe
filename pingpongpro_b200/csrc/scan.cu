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

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kPerThread = 32;                 // consecutive positions owned by one thread = one occupancy word per strand
constexpr int kTile = kScanTile;               // 8192 positions per tile
constexpr int kCandCap = 512;                  // candidates staged per round (a tile rarely has more)
static_assert(kTile == kThreads * kPerThread, "tile geometry");
constexpr int kBatch = 4;  // tiles claimed per ticket (one atomic on a single address costs ~1 ns: a human-sized genome has 378 k tiles)
static_assert((kBatch & (kBatch - 1)) == 0 && kBatch <= 32, "the batch aggregate is a butterfly sum over kBatch lanes");
static_assert(kTile <= (1 << 14) && kTile * PPP_MAX_OVERLAPS <= (1 << 20), "block-scan fields: 20 bits of records, 16 of candidates, 16 of ping-pong records");

__device__ __forceinline__ uint32_t warp_inclusive(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
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

// Decoupled look-back, executed by one whole warp: the packed sum of the aggregates of all batches before `batch`.  A
// batch's state is ONE 64-bit word (flag + value), so a plain volatile load sees a consistent pair and no fence is
// needed.  Batches are claimed in ticket order, hence every earlier batch belongs to a block that is already running.
// The resident blocks tend to run in step (they all publish, then all look back), so a batch may have to walk over every
// batch of its wave before it meets a finished prefix: a step covers 128 batches (four independent loads per lane).
__device__ __forceinline__ unsigned long long lookback_exclusive(const unsigned long long* __restrict__ state, uint32_t batch, int lane) {
    unsigned long long sum = 0;
    int64_t j = (int64_t)batch - 1;  // the nearest batch of the current window
    while (j >= 0) {
        unsigned long long v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t mine = j - 4 * lane - k;
            v[k] = mine >= 0 ? ld_state(state + mine) : kLbPrefix;  // before the first batch: an inclusive prefix of zero
        }
        unsigned long long part = 0;
        bool closed = false;  // this lane met a batch that knows its inclusive prefix
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
__device__ __noinline__ void count_tall(uint32_t* freqTail, uint32_t* tallList, unsigned int* tallCount, uint32_t tallCap, uint32_t r) {
    if (r >= kFreqCapacity) r = kFreqCapacity - 1;  // (+inf of a malformed NH:i:0 record: never index past the table)
    if (tallList) {
        unsigned int i = atomicAdd(tallCount, 1u);
        if (i < tallCap) tallList[i] = r;
    } else {
        atomicAdd(&freqTail[r], 1u);
    }
}

// 4-byte asynchronous copy global -> shared (LDGSTS): the next tile's occupancy words travel without occupying registers
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int REPR>
__global__ void __launch_bounds__(kThreads, 5) k_scan(ScanParams P) {
    // occupancy words of the batch's tiles
    __shared__ uint32_t sMask[kBatch][kThreads + 2];  // minus strand + the word after the tile + a zero
    __shared__ uint32_t sPlus[kBatch][kThreads];      // plus strand
    __shared__ uint32_t sOff[kBatch][4];              // compact index of the tile's first / past-the-last plus stack, first / past-the-last minus stack
    // per thread and tile, between the counting phase and the record phase
    __shared__ uint32_t sCand[kBatch][kThreads];         // which of the thread's plus stacks are candidates
    __shared__ uint32_t sExA[kBatch][kThreads];          // within the warp: plus stacks | minus stacks << 16 before the thread's positions
    __shared__ uint32_t sExB[kBatch][kThreads];          // within the warp: records | candidates << 18 before them
    __shared__ uint16_t sExPp[kBatch][kThreads];         // within the warp: ping-pong records before them
    __shared__ uint16_t sPrefM[kBatch][kThreads + 2];    // minus stacks of the tile before each word of sMask
    __shared__ uint32_t sWarpA[kBatch][kWarps];              // per warp: plus stacks | minus stacks << 16
    __shared__ unsigned long long sWarpB[kBatch][kWarps];    // per warp: records | candidates << 20 | ping-pong records << 36
    __shared__ unsigned long long sAgg[kBatch + 1], sBase[kBatch];  // packed (kLb* layout) records of the tile (last: of the batch) / before the tile
    __shared__ uint32_t sHist[kScanLowBins];
    __shared__ uint32_t sNextBatch;
    __shared__ uint32_t sCandBits[kCandCap], sCandOff[kCandCap];  // per candidate: present overlaps / first record (tile-relative)
    __shared__ uint16_t sCandTp[kCandCap], sCandRank[kCandCap], sCandPp[kCandCap];  // tile position / rank among the tile's plus stacks / ping-pong record

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < (int)kScanLowBins; k += kThreads) sHist[k] = 0;
    if (tid < kBatch) sMask[tid][kThreads + 1] = 0;

    const uint32_t ONE = REPR == kReprInt ? 1u : 0x3f800000u;
    const uint32_t wMask = P.W >= 32 ? 0xffffffffu : ((1u << P.W) - 1u);
    const float fW = (float)P.W;
    const uint32_t myPos = (uint32_t)tid * kPerThread;  // this thread owns 32 consecutive positions of a tile
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
        else count_tall(P.freqTail, P.tallList, P.tallCount, P.tallCap, r);
        localMax = max(localMax, r);
    };
    // the occupancy words and compact offsets of a batch of tiles, global -> shared, asynchronously
    auto fetch = [&](uint32_t first) {
        const uint32_t n = min(P.batchTiles, P.nTiles - first);
        for (uint32_t s = 0; s < n; ++s) {
            const size_t w = (size_t)(first + s) * kThreads;
            cp_async4(&sPlus[s][tid], P.occ + w + tid);
            cp_async4(&sMask[s][tid], P.occ + P.G / 32 + w + tid);
            if (tid == 0) cp_async4(&sMask[s][kThreads], P.occ + P.G / 32 + w + kThreads);
            if (tid >= 32 && tid < 36) cp_async4(&sOff[s][tid - 32], ((tid & 2) ? P.tileOffM : P.tileOffP) + first + s + (tid & 1));
        }
        cp_async_commit();
    };

    // A batch is claimed only when the block is ready to count and publish it at once: whoever holds an earlier ticket
    // publishes its record counts within microseconds, whatever its record phase takes afterwards.  (Claiming ahead, to
    // prefetch the occupancy words, made every block wait for the slowest record phase of the previous round.)
    for (;;) {
        __syncthreads();  // the previous batch's shared state is no longer read
        if (tid == 0) sNextBatch = atomicAdd(&P.state->ticket, P.batchTiles);
        __syncthreads();
        const uint32_t batch = sNextBatch;
        if (batch >= P.nTiles) break;
        const uint32_t nb = min(P.batchTiles, P.nTiles - batch);
        fetch(batch);
        cp_async_wait();
        __syncthreads();  // (A) the batch's occupancy words are complete

        // ---- counting phase (bitmaps only): how many records every tile of the batch holds, published at once, so that
        // the blocks holding later tiles never wait for this block's record phase
        for (uint32_t s = 0; s < nb; ++s) {
            const uint32_t nzP = sPlus[s][tid], nzM = sMask[s][tid];
            uint32_t cnt = 0, ncand = 0, npp = 0, cand = 0;
            uint32_t nz = nzP;
            while (nz) {  // candidates = plus stacks with at least one minus stack at overlaps min..max
                const int b = __ffs(nz) - 1;
                nz &= nz - 1;
                const uint32_t bit = myPos + b + P.minOv;
                const uint32_t bits = __funnelshift_r(sMask[s][bit >> 5], sMask[s][(bit >> 5) + 1], bit & 31) & wMask;
                if (bits) { cnt += __popc(bits); ++ncand; cand |= 1u << b; npp += (bits >> P.ppIdx) & 1u; }
            }
            // scans over the warp: stacks per strand (their ranks index the compact arrays), and (records, candidates,
            // ping-pong records), which are ordered (thread, position[, overlap]) == (position[, overlap])
            const uint32_t packA = (uint32_t)__popc(nzP) | ((uint32_t)__popc(nzM) << 16);
            const unsigned long long packB = (unsigned long long)cnt | ((unsigned long long)ncand << 20) | ((unsigned long long)npp << 36);
            const uint32_t incA = warp_inclusive(packA, lane);
            const unsigned long long incB = warp_inclusive64(packB, lane);
            if (lane == 31) { sWarpA[s][warp] = incA; sWarpB[s][warp] = incB; }
            const unsigned long long exB = incB - packB;
            sCand[s][tid] = cand;
            sExA[s][tid] = incA - packA;
            sExB[s][tid] = (uint32_t)(exB & 0xfffffu) | ((uint32_t)((exB >> 20) & 0xffffu) << 18);  // (within a warp: records < 2^15, candidates < 2^10)
            sExPp[s][tid] = (uint16_t)(exB >> 36);
        }
        __syncthreads();  // (B)
        const uint32_t bi = batch / P.batchTiles;  // (tickets are multiples of batchTiles)
        if (tid < 32) {  // the batch's aggregate, published before any of its records is written
            unsigned long long agg = 0;
            if (tid < (int)nb) {
                unsigned long long t = 0;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) t += sWarpB[tid][w];
                agg = (t & 0xfffffull) | ((t >> 36) << kLbRecBits);
                sAgg[tid] = agg;
            }
#pragma unroll
            for (int d = 1; d < kBatch; d <<= 1) agg += __shfl_xor_sync(0xffffffffu, agg, d);
            if (tid == 0) { sAgg[kBatch] = agg; st_state(P.tileState + bi, kLbAggregate | agg); }
        }
        // complete the ranks of the minus strand (window starts, halo exclusion)
        for (uint32_t s = 0; s < nb; ++s) {
            uint32_t before = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w)
                if (w < warp) before += sWarpA[s][w];
            const uint32_t exM = (before + sExA[s][tid]) >> 16;
            sPrefM[s][tid] = (uint16_t)exM;
            if (tid == kThreads - 1) sPrefM[s][kThreads] = (uint16_t)(exM + __popc(sMask[s][tid]));
        }
        __syncthreads();  // (C) aggregates and sPrefM are visible

        // ---- K2: height histogram, dense over the tiles' compact words (halo copies of a neighbouring shard are not counted)
        for (uint32_t s = 0; s < nb; ++s) {
            const uint32_t tileBase = (batch + s) * (uint32_t)kTile;
            const uint32_t offP = sOff[s][0], endP = sOff[s][1], offM = sOff[s][2], endM = sOff[s][3];
            for (uint32_t i = offP + tid; i < endP; i += kThreads * 2) {
                const uint32_t w0 = __ldg(P.cmpP + i);
                const uint32_t w1 = i + kThreads < endP ? __ldg(P.cmpP + i + kThreads) : 0u;
                classify(w0);
                if (w1) classify(w1);
            }
            uint32_t skipLo = 0xffffffffu, skipHi = 0xffffffffu;  // ranks (within the tile) of the halo copies
            if (tileBase < P.haloHi && tileBase + kTile > P.haloLo) {
                auto rankOf = [&](uint32_t q) {  // minus stacks of the tile at tile positions < q
                    if (q >= (uint32_t)kTile) return (uint32_t)sPrefM[s][kThreads];
                    return (uint32_t)sPrefM[s][q >> 5] + (uint32_t)__popc(sMask[s][q >> 5] & ((1u << (q & 31)) - 1u));
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

        // where the batch's records go: look back over the batches before it (by now they have published)
        if (warp == 0) {
            unsigned long long before = lookback_exclusive(P.tileState, bi, lane);
            if (lane == 0) {
                st_state(P.tileState + bi, kLbPrefix | ((before + sAgg[kBatch]) & kLbValueMask));
                found += sAgg[kBatch];
                for (uint32_t s = 0; s < nb; ++s) {
                    sBase[s] = before;
                    before += sAgg[s];
                }
            }
        }

        // ---- record phase, tile by tile
        for (uint32_t s = 0; s < nb; ++s) {
            unsigned long long totalB = 0, exB = 0;
            uint32_t beforeA = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const unsigned long long b = sWarpB[s][w];
                if (w < warp) { exB += b; beforeA += sWarpA[s][w]; }
                totalB += b;
            }
            const uint32_t total = (uint32_t)(totalB & 0xfffffu), candTotal = (uint32_t)((totalB >> 20) & 0xffffu);
            if (!total) continue;  // (block-uniform)
            const uint32_t tileBase = (batch + s) * (uint32_t)kTile;
            const uint32_t nzP = sPlus[s][tid], cand = sCand[s][tid];
            const uint32_t rankP = ((beforeA + sExA[s][tid]) & 0xffffu);  // plus stacks of the tile before this thread's positions
            const uint32_t firstRec = (uint32_t)(exB & 0xfffffu) + (sExB[s][tid] & 0x3ffffu);
            const uint32_t firstCand = (uint32_t)((exB >> 20) & 0xffffu) + (sExB[s][tid] >> 18);
            const uint32_t firstPp = (uint32_t)(exB >> 36) + sExPp[s][tid];
            const uint32_t rounds = (candTotal + kCandCap - 1) / kCandCap;
            for (uint32_t round = 0; round < rounds; ++round) {
                const uint32_t cbase = round * kCandCap;
                __syncthreads();  // the previous list is no longer read (first round: sBase of this batch is visible, too)
                {
                    uint32_t off = firstRec, rank = firstCand, pp = firstPp, cm = cand;
                    while (cm) {
                        const int b = __ffs(cm) - 1;
                        cm &= cm - 1;
                        const uint32_t tp = myPos + b, bit = tp + P.minOv;
                        const uint32_t bits = __funnelshift_r(sMask[s][bit >> 5], sMask[s][(bit >> 5) + 1], bit & 31) & wMask;
                        const uint32_t slot = rank - cbase;  // wraps to a huge value for earlier rounds
                        if (slot < (uint32_t)kCandCap) {
                            sCandTp[slot] = (uint16_t)tp;
                            sCandBits[slot] = bits;
                            sCandOff[slot] = off;
                            sCandRank[slot] = (uint16_t)(rankP + __popc(nzP & ((1u << b) - 1u)));
                            sCandPp[slot] = (uint16_t)pp;
                        }
                        off += __popc(bits);
                        pp += (bits >> P.ppIdx) & 1u;
                        ++rank;
                    }
                }
                __syncthreads();  // the list is complete
                const unsigned long long recBase = sBase[s] & kLbRecMask, ppBase = sBase[s] >> kLbRecBits;
                const uint32_t nIn = min((uint32_t)kCandCap, candTotal - cbase);
                for (uint32_t c = tid; c < nIn; c += kThreads) {  // dense: one thread per candidate
                    const uint32_t tp = sCandTp[c];
                    uint32_t bits = sCandBits[c];
                    const uint32_t wp = __ldg(P.cmpP + sOff[s][0] + sCandRank[c]);
                    const uint32_t bit = tp + P.minOv, word = bit >> 5;
                    // the present stacks of the window are consecutive compact words, starting at the rank of its first position
                    const uint32_t* win = P.cmpM + sOff[s][2] + sPrefM[s][word] + __popc(sMask[s][word] & ((1u << (bit & 31)) - 1u));
                    const int n = __popc(bits);
                    // vicinity sum / max over the PRESENT stacks only, in overlap order: an absent stack would add
                    // +0.0f, which is an exact no-op (:564-577)
                    float sum = 0.0f, mx = 0.0f;
                    for (int k = 0; k < n; ++k) {
                        const float mq = word_height<REPR>(__ldg(win + k));
                        sum = __fadd_rn(sum, mq);
                        mx = fmaxf(mx, mq);
                    }
                    const float mean = __fdiv_rn(sum, fW);                    // :578
                    const float hp = word_height<REPR>(wp);
                    unsigned long long idx = recBase + sCandOff[c];
                    for (int k = 0; k < n; ++k, ++idx) {                      // :584-586
                        const int o = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const uint32_t wm = __ldg(win + k);  // second touch: served by L1
                        const float mo = word_height<REPR>(wm);
                        const float local = __fdiv_rn(__fsub_rn(mo, __fsub_rn(mean, __fdiv_rn(mo, fW))), mx);  // :597
                        const uint32_t lbin = local < 0.2f ? 1u : 0u;               // :599 ((double)local < 0.2: no float lies in [0.2, 0.2f))
                        const uint32_t bbin = ((wp | wm) & kA10Bit) ? 0u : 1u;     // :602
                        const uint32_t misc = (uint32_t)o | (lbin << 5) | (bbin << 6);
                        if (idx < P.capacity) {                                    // :608
                            P.recPos[idx] = tileBase + tp;
                            P.recRp[idx] = hp;
                            P.recRm[idx] = mo;
                            P.recMisc[idx] = (uint8_t)misc;
                        }
                        if (o == P.ppIdx) {
                            const unsigned long long j = ppBase + sCandPp[c];
                            if (j < P.ppCapacity) {
                                PairRec r;
                                r.lpos = tileBase + tp;
                                r.rp = hp;
                                r.rm = mo;
                                r.misc = misc;
                                *reinterpret_cast<uint4*>(&P.ppRecs[j]) = *reinterpret_cast<uint4*>(&r);
                            }
                        }
                    }
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
                                                    uint32_t* __restrict__ freqTail, uint32_t* __restrict__ maxHeight, uint32_t* __restrict__ overflow) {
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
        if (n > cap) { if (blockIdx.x == 0 && threadIdx.x == 0) *overflow = 1u; n = cap; }
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&freqTail[slot[kTallHeader + i]], 1u);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *maxHeight = best;
}

// freq LUT: lut[h] = float(min(count[h], 2^24)) -- the value the reference's float counter holds (:508) --
// and the largest count (for maxHeightScore, :547-551).  Grid-stride up to the device-resident max height.
__global__ void __launch_bounds__(256) k_freq_lut(const unsigned long long* __restrict__ freqLow, const uint32_t* __restrict__ freqTail,
                                                  const uint32_t* __restrict__ maxHeight, float* __restrict__ lut, unsigned long long* maxCount) {
    const uint32_t maxH = *maxHeight;
    unsigned long long best = 0;
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h <= maxH; h += (uint64_t)gridDim.x * blockDim.x) {
        unsigned long long c = h < kScanLowBins ? freqLow[h] : (unsigned long long)freqTail[h];
        lut[h] = (float)(c < kSaturate ? (uint32_t)c : kSaturate);
        best = c > best ? c : best;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
        best = o > best ? o : best;
    }
    if ((threadIdx.x & 31) == 0 && best) atomicMax(maxCount, best);
}

// The bin thresholds of k_bin / k_score from the (global) maximum height frequency, on the device (log10f_emul.h): one
// thread per threshold.  The host recomputes the table with its libm while the device works on, and compares (api.cu).
__global__ void __launch_bounds__(32) k_bin_thresholds(const unsigned long long* __restrict__ maxCount, float* __restrict__ thresholds, BinParams* __restrict__ params) {
    const unsigned long long c = *maxCount;
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
                                                     const float* __restrict__ lut, const float* __restrict__ thresholds, const BinParams* __restrict__ params, int W,
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
    if (tid < 16) sSmallBin[tid] = (uint16_t)bin_of(__fmul_rn(lut[tid >> 2], lut[tid & 3]));  // (a height that never occurs has lut 0: never looked up)
    __syncthreads();
    const uint32_t onesBin = sSmallBin[5];  // rounded heights (1, 1)
    unsigned long long n = totals[0];
    if (n > capacity) n = capacity;
    for (size_t i = (size_t)blockIdx.x * kBinThreads + tid; i < n; i += (size_t)gridDim.x * kBinThreads) {
        const float rp = recRp[i], rm = recRm[i];
        const uint32_t misc = recMisc[i];
        const uint32_t hp = round_height(rp), hm = round_height(rm);  // heightScoreMap[0.5 + reads], :582 / :589
        uint32_t b;
        if ((hp | hm) < 4u) b = sSmallBin[hp * 4 + hm];
        else b = (uint32_t)bin_of(__fmul_rn(lut[hp], lut[hm]));
        recBin[i] = (uint16_t)b;
        if (b == onesBin) { atomicAdd(&sOnes[warp][misc & 127u], 1u); continue; }
        const uint32_t cell = (((b * (uint32_t)W + (misc & 31u)) * 2u + ((misc >> 6) & 1u)) * 2u) + ((misc >> 5) & 1u);  // bin-major table, :605
        const uint32_t h = (cell * 2654435761u) >> 20;
        uint32_t key = sKey[h];
        if (key == kBinEmpty) key = atomicCAS(&sKey[h], kBinEmpty, cell);
        if (key == kBinEmpty || key == cell) atomicAdd(&sVal[h], 1u);
        else atomicAdd(&grouped[cell], 1u);
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
    int perSm = 5;
    if (repr == kReprInt) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_scan<kReprInt>, kThreads, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_scan<kReprFloat>, kThreads, 0);
    if (perSm < 1) perSm = 1;
    unsigned grid = (unsigned)sms * (unsigned)perSm;
    const unsigned batches = (p.nTiles + p.batchTiles - 1) / p.batchTiles;
    if (grid > batches) grid = batches;
    if (repr == kReprInt) k_scan<kReprInt><<<grid, kThreads, 0, s>>>(p);
    else k_scan<kReprFloat><<<grid, kThreads, 0, s>>>(p);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// tiles claimed per ticket: as many as keep every block busy for a few tickets, at most kBatch
uint32_t scan_batch_tiles(uint32_t nTiles) {
    uint32_t b = nTiles / (148u * 5u * 4u);
    return b < 1u ? 1u : (b > (uint32_t)kBatch ? (uint32_t)kBatch : b);
}

cudaError_t launch_tall_apply(const uint32_t* gather, int world, uint32_t cap, unsigned long long* freqLow, uint32_t* freqTail, uint32_t* maxHeight,
                              uint32_t* overflow, cudaStream_t s, int* launches) {
    k_tall_apply<<<148, 256, 0, s>>>(gather, world, cap, freqLow, freqTail, maxHeight, overflow);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_freq_lut(const unsigned long long* freqLow, const uint32_t* freqTail, const uint32_t* maxHeight, float* lut,
                            unsigned long long* maxCount, cudaStream_t s, int* launches) {
    k_freq_lut<<<148 * 4, 256, 0, s>>>(freqLow, freqTail, maxHeight, lut, maxCount);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_bin_thresholds(const unsigned long long* maxCount, float* thresholds, BinParams* params, cudaStream_t s, int* launches) {
    k_bin_thresholds<<<(kBins + 31) / 32, 32, 0, s>>>(maxCount, thresholds, params);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_bin(const float* recRp, const float* recRm, const uint8_t* recMisc, uint16_t* recBin, const unsigned long long* totals, uint64_t capacity,
                       uint64_t expected, const float* lut, const float* thresholds, const BinParams* params, int W, uint32_t* grouped, cudaStream_t s, int* launches) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t want = (expected + kBinThreads * 8 - 1) / (kBinThreads * 8);  // >= 8 records per thread: the tables are flushed once per block
    unsigned blocks = (unsigned)std::min<uint64_t>(std::max<uint64_t>(want, 1), (uint64_t)sms * 4);
    k_bin<<<blocks, kBinThreads, 0, s>>>(recRp, recRm, recMisc, recBin, totals, capacity, lut, thresholds, params, W, grouped);
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
