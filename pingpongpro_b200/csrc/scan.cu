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

namespace pppk {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kPerThread = 32;                 // consecutive positions owned by one thread = one occupancy word per strand
constexpr int kTile = kScanTile;               // 8192 positions per tile
constexpr int kCandCap = 512;                  // candidates staged per round (a tile rarely has more)
static_assert(kTile == kThreads * kPerThread, "tile geometry");
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

// Decoupled look-back, executed by one whole warp: the packed sum of the aggregates of all tiles before `tile`.  A tile's
// state is ONE 64-bit word (flag + value), so a plain volatile load sees a consistent pair and no fence is needed.
// Tiles are claimed in ticket order, hence every earlier tile belongs to a block that is already running.
__device__ __forceinline__ unsigned long long lookback_exclusive(const unsigned long long* __restrict__ state, uint32_t tile, int lane) {
    unsigned long long sum = 0;
    int64_t j = (int64_t)tile - 1;  // the nearest tile of the current window
    while (j >= 0) {
        const int64_t mine = j - lane;
        unsigned long long v = kLbPrefix;  // before the first tile: an inclusive prefix of zero
        if (mine >= 0) {
            v = ld_state(state + mine);
            while ((v >> 62) == 0ull) v = ld_state(state + mine);
        }
        const unsigned prefixLanes = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
        const int stop = prefixLanes ? __ffs(prefixLanes) - 1 : 32;  // the nearest tile that already knows its inclusive prefix
        unsigned long long part = lane <= stop ? (v & kLbValueMask) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
        sum += part;
        if (prefixLanes) break;
        j -= 32;
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
    __shared__ uint32_t sMask[2][kThreads + 2];   // minus-strand occupancy words of the tile + the word after it + a zero
    __shared__ uint32_t sPlus[2][kThreads];       // plus-strand occupancy words of the tile
    __shared__ uint32_t sOff[2][4];               // compact index of the tile's first / past-the-last plus stack, first / past-the-last minus stack
    __shared__ uint16_t sPrefM[kThreads + 2];     // minus stacks of the tile before each word of sMask
    __shared__ uint32_t sHist[kScanLowBins];
    __shared__ uint32_t sWarpA[kWarps];               // per warp: plus stacks | minus stacks << 16
    __shared__ unsigned long long sWarpB[kWarps];     // per warp: records | candidates << 20 | ping-pong records << 36
    __shared__ unsigned long long sBase;              // packed number of records before this tile (kLb* layout)
    __shared__ unsigned long long sCarry;             // (thread 0) the same for the next tile of the batch
    __shared__ unsigned long long sFound[2];          // (thread 0) records / ping-pong records this block found
    __shared__ uint32_t sNextBatch;
    __shared__ uint32_t sCandBits[kCandCap], sCandOff[kCandCap];  // per candidate: present overlaps / first record (tile-relative)
    __shared__ uint16_t sCandTp[kCandCap], sCandRank[kCandCap], sCandPp[kCandCap];  // tile position / rank among the tile's plus stacks / ping-pong record

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < (int)kScanLowBins; k += kThreads) sHist[k] = 0;
    if (tid < 2) sMask[tid][kThreads + 1] = 0;
    if (tid == 0) {
        sNextBatch = atomicAdd(&P.state->ticket, P.batchTiles);
        sCarry = 0;
        sFound[0] = sFound[1] = 0;
    }
    __syncthreads();

    const uint32_t ONE = REPR == kReprInt ? 1u : 0x3f800000u;
    const uint32_t wMask = P.W >= 32 ? 0xffffffffu : ((1u << P.W) - 1u);
    const float fW = (float)P.W;
    const uint32_t myPos = (uint32_t)tid * kPerThread;  // this thread owns 32 consecutive positions of the tile
    uint32_t ones = 0, localMax = 0, small = 0;  // small: four 8-bit counters for rounded heights 0..3
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
    // the occupancy words and compact offsets of a tile, global -> shared, asynchronously
    auto fetch = [&](uint32_t t, int into) {
        const size_t w = (size_t)t * kThreads;
        cp_async4(&sPlus[into][tid], P.occ + w + tid);
        cp_async4(&sMask[into][tid], P.occ + P.G / 32 + w + tid);
        if (tid == 0) cp_async4(&sMask[into][kThreads], P.occ + P.G / 32 + w + kThreads);
        if (tid >= 32 && tid < 36) cp_async4(&sOff[into][tid - 32], ((tid & 2) ? P.tileOffM : P.tileOffP) + t + (tid & 1));
        cp_async_commit();
    };
    uint32_t batch = sNextBatch;
    int buf = 0;
    if (batch < P.nTiles) fetch(batch, buf);

    while (batch < P.nTiles) {
        const uint32_t batchEnd = min(batch + P.batchTiles, P.nTiles);
        for (uint32_t tile = batch; tile < batchEnd; ++tile, buf ^= 1) {
            const uint32_t tileBase = tile * (uint32_t)kTile;
            cp_async_wait();
            __syncthreads();  // (A) the tile's occupancy words are complete
            const uint32_t nzP = sPlus[buf][tid], nzM = sMask[buf][tid];

            // K3, bitmaps only: candidates = plus stacks with at least one minus stack at overlaps min..max
            uint32_t cnt = 0, ncand = 0, npp = 0, cand = 0;
            {
                uint32_t nz = nzP;
                while (nz) {
                    const int b = __ffs(nz) - 1;
                    nz &= nz - 1;
                    const uint32_t bit = myPos + b + P.minOv;
                    const uint32_t bits = __funnelshift_r(sMask[buf][bit >> 5], sMask[buf][(bit >> 5) + 1], bit & 31) & wMask;
                    if (bits) { cnt += __popc(bits); ++ncand; cand |= 1u << b; npp += (bits >> P.ppIdx) & 1u; }
                }
            }
            // block scans: stacks per strand (their ranks index the compact arrays), and (records, candidates, ping-pong
            // records), which are ordered (thread, position[, overlap]) == (position[, overlap])
            const uint32_t packA = (uint32_t)__popc(nzP) | ((uint32_t)__popc(nzM) << 16);
            const unsigned long long packB = (unsigned long long)cnt | ((unsigned long long)ncand << 20) | ((unsigned long long)npp << 36);
            uint32_t exA = warp_inclusive(packA, lane);
            unsigned long long exB = warp_inclusive64(packB, lane);
            if (lane == 31) { sWarpA[warp] = exA; sWarpB[warp] = exB; }
            exA -= packA;
            exB -= packB;
            __syncthreads();  // (B)
            unsigned long long totalB = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const unsigned long long b = sWarpB[w];
                if (w < warp) { exA += sWarpA[w]; exB += b; }
                totalB += b;
            }
            // exA: stacks of the tile before this thread's positions; exB: records / candidates / ping-pong records before them
            const uint32_t total = (uint32_t)(totalB & 0xfffffu), candTotal = (uint32_t)((totalB >> 20) & 0xffffu);
            sPrefM[tid] = (uint16_t)(exA >> 16);
            if (tid == kThreads - 1) sPrefM[kThreads] = (uint16_t)((exA >> 16) + __popc(nzM));
            const unsigned long long aggregate = (unsigned long long)total | ((totalB >> 36) << kLbRecBits);
            const bool knowsPrefix = tile == 0 || tile != batch;  // the first tile of all, or a tile whose predecessor this block scanned itself
            if (tid == 0) {
                sFound[0] += total;
                sFound[1] += totalB >> 36;
                st_state(P.tileState + tile, knowsPrefix ? (kLbPrefix | ((sCarry + aggregate) & kLbValueMask)) : (kLbAggregate | aggregate));
                if (tile + 1 == batchEnd) sNextBatch = atomicAdd(&P.state->ticket, P.batchTiles);
            }

            // K2: height histogram, dense over the tile's compact words (halo copies of a neighbouring shard are not counted)
            {
                const uint32_t offP = sOff[buf][0], endP = sOff[buf][1], offM = sOff[buf][2], endM = sOff[buf][3];
                for (uint32_t i = offP + tid; i < endP; i += kThreads * 2) {
                    const uint32_t w0 = __ldg(P.cmpP + i);
                    const uint32_t w1 = i + kThreads < endP ? __ldg(P.cmpP + i + kThreads) : 0u;
                    classify(w0);
                    if (w1) classify(w1);
                }
                uint32_t skipLo = 0xffffffffu, skipHi = 0xffffffffu;  // ranks (within the tile) of the halo copies
                if (tileBase < P.haloHi && tileBase + kTile > P.haloLo) {  // block-uniform, at most a few tiles per shard
                    __syncthreads();  // sPrefM of this tile
                    auto rankOf = [&](uint32_t q) {  // minus stacks of the tile at tile positions < q
                        if (q >= (uint32_t)kTile) return (uint32_t)sPrefM[kThreads];
                        return (uint32_t)sPrefM[q >> 5] + (uint32_t)__popc(sMask[buf][q >> 5] & ((1u << (q & 31)) - 1u));
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

            // where the tile's records go: look back over the tiles before it (warp 0; the other warps stage meanwhile)
            if (warp == 0) {
                unsigned long long before;
                if (knowsPrefix) {
                    before = sCarry;
                } else {
                    before = lookback_exclusive(P.tileState, tile, lane);
                    if (lane == 0) st_state(P.tileState + tile, kLbPrefix | ((before + aggregate) & kLbValueMask));
                }
                if (lane == 0) { sBase = before; sCarry = before + aggregate; }
            }

            const uint32_t rounds = total ? (candTotal + kCandCap - 1) / kCandCap : 1u;  // (block-uniform)
            for (uint32_t round = 0; round < rounds; ++round) {
                const uint32_t cbase = round * kCandCap;
                if (round) __syncthreads();  // the previous round's list is no longer read
                {
                    uint32_t off = (uint32_t)(exB & 0xfffffu), rank = (uint32_t)((exB >> 20) & 0xffffu), pp = (uint32_t)(exB >> 36), cm = cand;
                    while (cm) {
                        const int b = __ffs(cm) - 1;
                        cm &= cm - 1;
                        const uint32_t tp = myPos + b, bit = tp + P.minOv;
                        const uint32_t bits = __funnelshift_r(sMask[buf][bit >> 5], sMask[buf][(bit >> 5) + 1], bit & 31) & wMask;
                        const uint32_t slot = rank - cbase;  // wraps to a huge value for earlier rounds
                        if (slot < (uint32_t)kCandCap) {
                            sCandTp[slot] = (uint16_t)tp;
                            sCandBits[slot] = bits;
                            sCandOff[slot] = off;
                            sCandRank[slot] = (uint16_t)((exA & 0xffffu) + __popc(nzP & ((1u << b) - 1u)));
                            sCandPp[slot] = (uint16_t)pp;
                        }
                        off += __popc(bits);
                        pp += (bits >> P.ppIdx) & 1u;
                        ++rank;
                    }
                }
                __syncthreads();  // (C) the list, sBase, sPrefM and sNextBatch are visible
                if (round == 0) {  // the next tile's occupancy words travel while the records are written
                    const uint32_t nextTile = tile + 1 < batchEnd ? tile + 1 : sNextBatch;
                    if (nextTile < P.nTiles) fetch(nextTile, buf ^ 1);
                }
                if (!total) break;
                const unsigned long long recBase = sBase & kLbRecMask, ppBase = sBase >> kLbRecBits;
                const uint32_t nIn = min((uint32_t)kCandCap, candTotal - cbase);
                for (uint32_t c = tid; c < nIn; c += kThreads) {  // dense: one thread per candidate
                    const uint32_t tp = sCandTp[c];
                    uint32_t bits = sCandBits[c];
                    const uint32_t wp = __ldg(P.cmpP + sOff[buf][0] + sCandRank[c]);
                    const uint32_t bit = tp + P.minOv, word = bit >> 5;
                    // the present stacks of the window are consecutive compact words, starting at the rank of its first position
                    const uint32_t* win = P.cmpM + sOff[buf][2] + sPrefM[word] + __popc(sMask[buf][word] & ((1u << (bit & 31)) - 1u));
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
                        const uint32_t lbin = ((double)local < 0.2) ? 1u : 0u;     // :599
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
        batch = sNextBatch;  // (published by the last tile's barrier C; rewritten only after the next batch's barrier B)
    }

    // flush the privatised histogram and this block's record counts
    flush_small();
    if (ones) { atomicAdd(&sHist[1], ones); localMax = max(localMax, 1u); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) localMax = max(localMax, __shfl_xor_sync(0xffffffffu, localMax, d));
    if (lane == 0 && localMax) atomicMax(P.maxHeight, localMax);
    if (tid == 0 && (sFound[0] | sFound[1])) { atomicAdd(&P.state->totals[0], sFound[0]); atomicAdd(&P.state->totals[1], sFound[1]); }
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
                                                     const float* __restrict__ lut, const float* __restrict__ thresholds, float binScale, int W,
                                                     uint32_t* __restrict__ grouped) {
    __shared__ float sThr[kBins];
    __shared__ uint32_t sKey[kBinCache], sVal[kBinCache];
    __shared__ uint32_t sOnes[kBinWarps][128];
    __shared__ uint16_t sSmallBin[16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < kBins - 1; k += kBinThreads) sThr[k] = thresholds[k];
    for (int k = tid; k < kBinCache; k += kBinThreads) { sKey[k] = kBinEmpty; sVal[k] = 0; }
    for (int k = tid; k < kBinWarps * 128; k += kBinThreads) (&sOnes[0][0])[k] = 0;
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

// tiles claimed per ticket: one atomic on a single address costs ~1 ns, so large genomes take several tiles at a time
uint32_t scan_batch_tiles(uint32_t nTiles) {
    uint32_t b = nTiles / (148u * 6u * 8u);
    return b < 1u ? 1u : (b > 8u ? 8u : b);
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

cudaError_t launch_bin(const float* recRp, const float* recRm, const uint8_t* recMisc, uint16_t* recBin, const unsigned long long* totals, uint64_t capacity,
                       uint64_t expected, const float* lut, const float* thresholds, float binScale, int W, uint32_t* grouped, cudaStream_t s, int* launches) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t want = (expected + kBinThreads * 8 - 1) / (kBinThreads * 8);  // >= 8 records per thread: the tables are flushed once per block
    unsigned blocks = (unsigned)std::min<uint64_t>(std::max<uint64_t>(want, 1), (uint64_t)sms * 4);
    k_bin<<<blocks, kBinThreads, 0, s>>>(recRp, recRm, recMisc, recBin, totals, capacity, lut, thresholds, binScale, W, grouped);
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
