// scan.cu -- the streaming pass over the dense stack arrays (K2 + K3 of SURVEY.md §8(a)).
//
// k_scan (fused, occupancy driven): ONE pass that
//   (K2) builds the height -> frequency histogram of mapHeightsToScores (pingpongpro.cpp:502-509), and
//   (K3) finds, for every plus stack, the minus stacks at overlaps min..max (:559-577), computes the vicinity
//        mean / max (:570-578), the local-height bin (:597-599) and the base-bias bin (:602) and appends
//        one record per (plus stack, minus stack, overlap) (:608).
// Everything of countStacksByGroup that needs the finished histogram (the height-score bin :589-594 and the
// grouped counter :605) is deferred to k_bin_group, which touches only the (few) signature records.
//
// The dense stack arrays are ~96 % zeros.  A streaming version of this kernel (8 bytes per position, 256-bit loads,
// shared-memory staging; 0.38 ms on configs[1] = 50 % of the HBM roofline, 70-80 % on the sparser mouse / human
// sized genomes -- profiles/README.md) spent ~1500 instructions per 64 words mostly discovering that.  This version
// reads the OCCUPANCY BITMAPS that the stack builder maintains (1 bit per word: G/4 bytes instead of 8 G) and gathers
// only the words that exist, so the work is proportional to the number of stacks, like the reference's std::map walk
// -- but over 8192 positions per CTA iteration in parallel:
//   * a thread owns 32 consecutive positions = one bitmap word per strand (prefetched one tile ahead);
//   * histogram: loop over the set bits, gather the word (one 32-byte sector), classify (height 1 in a register,
//     heights 0..3 in packed byte counters, bins < 512 in shared memory, the tall tail by red.add to a global table);
//   * the window test of a plus stack is one 64-bit funnel shift over two minus-strand bitmap words in shared memory;
//   * all per-signature work is dense: block scan of (records, candidates), candidate list in shared memory, one
//     thread per CANDIDATE gathers the present window words, evaluates the float expressions in the reference's
//     order and writes its records.  The records of a tile are written contiguously (one atomicAdd on a global
//     cursor) in (position, overlap) order; k_bin_group later copies the runs into global position order.
//
// Bound: DRAM sector gathers (32 B per touched sector) + latency.  Algorithmic bytes in SURVEY §8(d)'s dense
// accounting stay 8 * G + 16 * P per launch; the traffic actually moved is G/4 + 32 B per occupied sector + 16 * P.
#include "common.cuh"

namespace pppk {

namespace {

constexpr int kThreads = 256;
constexpr int kPerThread = 32;                 // consecutive positions owned by one thread = one occupancy word per strand
constexpr int kTile = kScanTile;               // 8192 positions per CTA iteration
constexpr int kCandCap = 512;                  // candidates staged per round (a tile rarely has more)
static_assert(kTile == kThreads * kPerThread, "tile geometry");

__device__ __forceinline__ uint32_t warp_inclusive(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// a stack taller than the shared-memory bins (rare): out of line, so that it costs the streaming loop no registers
__device__ __noinline__ void count_tall(uint32_t* freqTail, uint32_t* tallList, unsigned int* tallCount, uint32_t tallCap, uint32_t r) {
    if (tallList) {
        unsigned int i = atomicAdd(tallCount, 1u);
        if (i < tallCap) tallList[i] = r;
    } else {
        atomicAdd(&freqTail[r], 1u);
    }
}

template <int REPR>
__global__ void __launch_bounds__(kThreads, 8) k_scan(ScanParams P) {
    __shared__ uint32_t sMask[2][kThreads + 2];  // minus-strand occupancy words of the tile + the word after it
    __shared__ uint32_t sHist[kScanLowBins];
    __shared__ uint32_t sWarpTot[kThreads / 32];
    __shared__ unsigned long long sBase;
    __shared__ uint32_t sCandBits[kCandCap], sCandOff[kCandCap];  // per candidate: present overlaps / first record (tile-relative)
    __shared__ uint16_t sCandTp[kCandCap];                        // its tile position

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < (int)kScanLowBins; k += kThreads) sHist[k] = 0;
    if (tid < 2) sMask[tid][kThreads + 1] = 0;
    __syncthreads();

    const uint32_t ONE = REPR == kReprInt ? 1u : 0x3f800000u;
    const uint32_t wMask = P.W >= 32 ? 0xffffffffu : ((1u << P.W) - 1u);
    const float fW = (float)P.W;
    const uint32_t myPos = (uint32_t)tid * kPerThread;  // this thread owns 32 consecutive positions of the tile
    const uint32_t* __restrict__ occP = P.occ;
    const uint32_t* __restrict__ occM = P.occ + P.G / 32;
    const uint32_t* __restrict__ Hp = P.H;
    const uint32_t* __restrict__ Hm = P.H + P.G;
    uint32_t ones = 0, localMax = 0, small = 0;  // small: four 8-bit counters for rounded heights 0..3
    auto flush_small = [&]() {
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if ((small >> (8 * r)) & 0xffu) atomicAdd(&sHist[r], (small >> (8 * r)) & 0xffu);
        small = 0;
    };

    // the occupancy words of the next tile are fetched while the current one is processed
    uint32_t nextP = 0, nextM = 0, nextHalo = 0;
    auto fetch = [&](uint32_t t) {
        const size_t w = (size_t)t * kThreads;
        nextP = __ldg(occP + w + tid);
        nextM = __ldg(occM + w + tid);
        if (tid == 0) nextHalo = __ldg(occM + w + kThreads);
    };
    if (blockIdx.x < P.nTiles) fetch(blockIdx.x);

    int buf = 0;
    for (uint32_t tile = blockIdx.x; tile < P.nTiles; tile += gridDim.x, buf ^= 1) {
        const uint32_t tileBase = tile * (uint32_t)kTile;
        const uint32_t nzP = nextP, nzM = nextM;
        sMask[buf][tid] = nzM;
        if (tid == 0) sMask[buf][kThreads] = nextHalo;
        if (tile + gridDim.x < P.nTiles) fetch(tile + gridDim.x);

        // K2: height histogram over the stacks that exist (set bits), gathered from the dense arrays.
        {
            uint32_t mm = nzM;
            if (tileBase < P.haloHi && tileBase + kTile > P.haloLo) {  // halo copies of a neighbouring shard are not counted
                uint32_t first = tileBase + myPos;
                uint32_t lo = P.haloLo > first ? P.haloLo - first : 0u, hi = P.haloHi > first ? P.haloHi - first : 0u;
                uint32_t below = lo >= 32 ? 0xffffffffu : ((1u << lo) - 1u), upto = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
                mm &= below | ~upto;
            }
            // up to four gathers (two per strand) are in flight per trip before any of them is classified
            uint32_t nzA = nzP, nzB = mm;
            while (nzA | nzB) {
                uint32_t wv[4] = {0u, 0u, 0u, 0u};
                if (nzA) { wv[0] = __ldg(Hp + tileBase + myPos + (uint32_t)(__ffs(nzA) - 1)); nzA &= nzA - 1; }
                if (nzB) { wv[1] = __ldg(Hm + tileBase + myPos + (uint32_t)(__ffs(nzB) - 1)); nzB &= nzB - 1; }
                if (nzA) { wv[2] = __ldg(Hp + tileBase + myPos + (uint32_t)(__ffs(nzA) - 1)); nzA &= nzA - 1; }
                if (nzB) { wv[3] = __ldg(Hm + tileBase + myPos + (uint32_t)(__ffs(nzB) - 1)); nzB &= nzB - 1; }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!wv[k]) continue;
                    if ((wv[k] & kValueMask) == ONE) { ++ones; continue; }
                    uint32_t r = word_round_height<REPR>(wv[k]);
                    if (r < 4u) { small += 1u << (8 * r); if ((small & 0x80808080u) != 0u) flush_small(); }  // heights 0..3: byte counters in a register
                    else if (r < kScanLowBins) atomicAdd(&sHist[r], 1u);
                    else count_tall(P.freqTail, P.tallList, P.tallCount, P.tallCap, r);
                    localMax = max(localMax, r);
                }
            }
        }
        __syncthreads();  // (A) the tile's minus occupancy words are complete

        // K3: candidates = plus stacks with at least one minus stack at overlaps min..max (bitmaps only)
        uint32_t cnt = 0, ncand = 0, cand = 0;
        {
            uint32_t nz = nzP;
            while (nz) {
                int b = __ffs(nz) - 1;
                nz &= nz - 1;
                uint32_t bit = myPos + b + P.minOv;
                uint32_t bits = __funnelshift_r(sMask[buf][bit >> 5], sMask[buf][(bit >> 5) + 1], bit & 31) & wMask;
                if (bits) { cnt += __popc(bits); ++ncand; cand |= 1u << b; }
            }
        }
        // records and candidates are ordered (thread, position[, overlap]) == (position[, overlap]);
        // one warp scan carries both counters (records in the low 20 bits, candidates above)
        const uint32_t packed = cnt | (ncand << 20);
        const uint32_t inc = warp_inclusive(packed, lane);
        if (lane == 31) sWarpTot[warp] = inc;
        __syncthreads();  // (B)
        uint32_t beforeP = 0, totalP = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) {
            uint32_t t = sWarpTot[w];
            if (w < warp) beforeP += t;
            totalP += t;
        }
        const uint32_t total = totalP & 0xfffffu, candTotal = totalP >> 20;
        if (tid == 0) {
            unsigned long long base = total ? atomicAdd(P.cursor, (unsigned long long)total) : 0ull;
            sBase = base;
            P.tileCnt[tile] = total;
            P.tileBase[tile] = (uint32_t)base;
        }
        if (total) {  // block-uniform: the rare path, processed densely (one thread per candidate)
            const uint32_t mine = beforeP + inc - packed;
            const uint32_t firstRec = mine & 0xfffffu, firstCand = mine >> 20;  // tile-relative ranks of this thread's first record / candidate
            for (uint32_t cbase = 0; cbase < candTotal; cbase += kCandCap) {
                if (cbase) __syncthreads();  // the previous round's list is no longer read
                uint32_t off = firstRec, rank = firstCand, cm = cand;
                while (cm) {
                    int b = __ffs(cm) - 1;
                    cm &= cm - 1;
                    uint32_t tp = myPos + b, bit = tp + P.minOv;
                    uint32_t bits = __funnelshift_r(sMask[buf][bit >> 5], sMask[buf][(bit >> 5) + 1], bit & 31) & wMask;
                    uint32_t slot = rank - cbase;  // wraps to a huge value for earlier rounds
                    if (slot < (uint32_t)kCandCap) { sCandTp[slot] = (uint16_t)tp; sCandBits[slot] = bits; sCandOff[slot] = off; }
                    off += __popc(bits);
                    ++rank;
                }
                __syncthreads();  // (C)
                const uint32_t nIn = min((uint32_t)kCandCap, candTotal - cbase);
                for (uint32_t c = tid; c < nIn; c += kThreads) {
                    const uint32_t tp = sCandTp[c];
                    uint32_t bits = sCandBits[c];
                    const uint32_t wp = __ldg(Hp + tileBase + tp);
                    const uint32_t* win = Hm + tileBase + tp + P.minOv;  // first window word
                    // vicinity sum / max over the PRESENT stacks only, in overlap order: an absent stack would add
                    // +0.0f, which is an exact no-op (:564-577)
                    float sum = 0.0f, mx = 0.0f;
                    for (uint32_t rest = bits; rest; rest &= rest - 1) {
                        float mq = word_height<REPR>(__ldg(win + (__ffs(rest) - 1)));
                        sum = __fadd_rn(sum, mq);
                        mx = fmaxf(mx, mq);
                    }
                    const float mean = __fdiv_rn(sum, fW);                    // :578
                    const float hp = word_height<REPR>(wp);
                    unsigned long long idx = sBase + sCandOff[c];
                    while (bits) {                                            // :584-586
                        const int o = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const uint32_t wm = __ldg(win + o);  // second touch: served by L1
                        const float mo = word_height<REPR>(wm);
                        const float local = __fdiv_rn(__fsub_rn(mo, __fsub_rn(mean, __fdiv_rn(mo, fW))), mx);  // :597
                        const uint32_t lbin = ((double)local < 0.2) ? 1u : 0u;     // :599
                        const uint32_t bbin = ((wp | wm) & kA10Bit) ? 0u : 1u;     // :602
                        if (idx < P.capacity) {
                            PairRec r;
                            r.lpos = tileBase + tp;
                            r.rp = hp;
                            r.rm = mo;
                            r.misc = (uint32_t)o | (lbin << 5) | (bbin << 6);
                            *reinterpret_cast<uint4*>(&P.pairsRaw[idx]) = *reinterpret_cast<uint4*>(&r);
                        }
                        ++idx;
                    }
                }
            }
        }
    }

    // flush the privatised histogram
    flush_small();
    if (ones) { atomicAdd(&sHist[1], ones); localMax = max(localMax, 1u); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) localMax = max(localMax, __shfl_xor_sync(0xffffffffu, localMax, d));
    if (lane == 0 && localMax) atomicMax(P.maxHeight, localMax);
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

// K3 part 2: height-score bin (:589-594) + grouped counter (:605) of every signature, and the copy of the
// per-tile runs into global (position, overlap) order.  Persistent blocks; each WARP takes whole tiles, so a
// record's tile (and with it its ordered destination) is known without a search.
// The bin of a height-score product is the number of host-built thresholds <= hs (see api.cu).  A device log10
// approximation proposes the bin, two short loops against the exact thresholds correct it, so the result is the
// exact step function whatever the approximation error.
// The grouped counters are concentrated on a few hundred hot cells (singleton x singleton pairs all land in the
// top bin), so updates are merged per warp (match.any), then per block in a shared-memory hash table that lives
// for the whole kernel; only the table's entries reach the global (L2) counters.
constexpr int kBgThreads = 512;
constexpr int kBgWarps = kBgThreads / 32;
constexpr int kBgHash = 4096;
constexpr int kBgPerLane = 2;               // records per lane in flight
constexpr int kBgItem = 32 * kBgPerLane;    // records per work item
constexpr uint32_t kBgEmpty = 0xffffffffu;
__global__ void __launch_bounds__(kBgThreads) k_bin_group(const PairRec* __restrict__ raw, PairRec* __restrict__ ordered, const uint32_t* __restrict__ tileCnt,
                                                          const uint32_t* __restrict__ tileBase, const uint32_t* __restrict__ tileDst, uint32_t nTiles,
                                                          const float* __restrict__ lut, const float* __restrict__ thresholds, float binScale, int W,
                                                          uint32_t* __restrict__ grouped, uint32_t* __restrict__ tileCursor) {
    __shared__ float sThr[kBins];
    __shared__ uint32_t sKey[kBgHash], sVal[kBgHash];
    __shared__ uint32_t sFirst, sItemStart[33 + 16], sCnt[32], sSrc[32], sDst[32];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int k = tid; k < kBins - 1; k += kBgThreads) sThr[k] = thresholds[k];
    for (int k = tid; k < kBgHash; k += kBgThreads) { sKey[k] = kBgEmpty; sVal[k] = 0; }
    __syncthreads();
    auto bin_of = [&](float hs) {
        int b = (int)(__log10f(hs) * binScale + 0.5f);             // proposal only
        b = min(max(b, 0), kBins - 1);
        while (b > 0 && sThr[b - 1] > hs) --b;                    // exact: bin = #thresholds <= hs
        while (b < kBins - 1 && sThr[b] <= hs) ++b;
        return b;
    };
    auto count_cell = [&](bool ok, uint32_t cell) {                // :605, one update per distinct cell per warp
        const unsigned act = __ballot_sync(0xffffffffu, ok);
        if (!ok) return;
        const unsigned peers = __match_any_sync(act, cell);
        if (lane != __ffs(peers) - 1) return;
        const uint32_t n = (uint32_t)__popc(peers);
        uint32_t h = (cell * 2654435761u) >> 20;
        for (int probe = 0; probe < 32; ++probe) {
            uint32_t old = atomicCAS(&sKey[h], kBgEmpty, cell);
            if (old == kBgEmpty || old == cell) { atomicAdd(&sVal[h], n); return; }
            h = (h + 1) & (kBgHash - 1);
        }
        atomicAdd(&grouped[cell], n);
    };
    // Work is handed out dynamically, 32 tiles per block at a time (one atomic per batch: a single global counter
    // serialises at ~1 ns per atomic).  Signature-dense tiles hold thousands of records and every 64-record step is
    // latency bound, so the block's warps share the 64-record items of ALL tiles of the batch instead of owning tiles.
    const int warp = tid >> 5;
    for (;;) {
        if (tid == 0) sFirst = atomicAdd(tileCursor, 32u);
        __syncthreads();
        const uint32_t first = sFirst;
        if (first >= nTiles) break;
        if (warp == 0) {
            const uint32_t tile = first + lane;
            const uint32_t cnt = tile < nTiles ? tileCnt[tile] : 0u;
            sCnt[lane] = cnt;
            sSrc[lane] = tile < nTiles ? tileBase[tile] : 0u;
            sDst[lane] = tile < nTiles ? tileDst[tile] : 0u;
            const uint32_t items = (cnt + (uint32_t)kBgItem - 1u) / (uint32_t)kBgItem;
            const uint32_t inc = warp_inclusive(items, lane);
            sItemStart[lane] = inc - items;
            if (lane == 31) sItemStart[32] = inc;
        }
        __syncthreads();
        const uint32_t totalItems = sItemStart[32];
        for (uint32_t item = warp; item < totalItems; item += kBgWarps) {
            int t = 0;  // tile slot with sItemStart[t] <= item < sItemStart[t + 1] (warp-uniform)
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
                if (sItemStart[t + step] <= item) t += step;
            const uint32_t cnt = sCnt[t], base = (item - sItemStart[t]) * (uint32_t)kBgItem;
            const size_t src = sSrc[t], dst = sDst[t];
            PairRec r[kBgPerLane];
            bool ok[kBgPerLane];
            float fp[kBgPerLane], fm[kBgPerLane];
#pragma unroll
            for (int u = 0; u < kBgPerLane; ++u) {  // kBgPerLane independent record loads in flight per lane
                const uint32_t i = base + u * 32 + lane;
                ok[u] = i < cnt;
                r[u] = PairRec{0, 1.f, 1.f, 0};
                if (ok[u]) *reinterpret_cast<uint4*>(&r[u]) = *reinterpret_cast<const uint4*>(&raw[src + i]);
            }
#pragma unroll
            for (int u = 0; u < kBgPerLane; ++u) {  // heightScoreMap[0.5 + reads], :582 / :589
                fp[u] = lut[round_height(r[u].rp)];
                fm[u] = lut[round_height(r[u].rm)];
            }
#pragma unroll
            for (int u = 0; u < kBgPerLane; ++u) {
                uint32_t cell = 0;
                if (ok[u]) {
                    int b = bin_of(__fmul_rn(fp[u], fm[u]));
                    r[u].misc |= (uint32_t)b << 7;
                    cell = ((((uint32_t)b * (uint32_t)W + (r[u].misc & 31u)) * 2u + ((r[u].misc >> 6) & 1u)) * 2u) + ((r[u].misc >> 5) & 1u);  // bin-major table
                    *reinterpret_cast<uint4*>(&ordered[dst + base + u * 32 + lane]) = *reinterpret_cast<uint4*>(&r[u]);
                }
                count_cell(ok[u], cell);
            }
        }
        __syncthreads();  // the batch descriptors are overwritten next
    }
    __syncthreads();
    for (int k = tid; k < kBgHash; k += kBgThreads)
        if (sKey[k] != kBgEmpty) atomicAdd(&grouped[sKey[k]], sVal[k]);
}

}  // namespace

cudaError_t launch_scan(const ScanParams& p, int repr, cudaStream_t s, int* launches) {
    if (p.nTiles == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int perSm = 8;
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

cudaError_t launch_bin_group(const PairRec* raw, PairRec* ordered, const uint32_t* tileCnt, const uint32_t* tileBase, const uint32_t* tileDst, uint32_t nTiles,
                             const float* lut, const float* thresholds, float binScale, int W, uint32_t* grouped, uint32_t* tileCursor, cudaStream_t s,
                             int* launches) {
    if (nTiles == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned blocks = (nTiles + kBgWarps - 1) / kBgWarps;
    if (blocks > (unsigned)sms * 4u) blocks = (unsigned)sms * 4u;  // 4 x 512 threads = full occupancy; the kernel is load-latency bound
    cudaError_t e = cudaMemsetAsync(tileCursor, 0, sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    k_bin_group<<<blocks, kBgThreads, 0, s>>>(raw, ordered, tileCnt, tileBase, tileDst, nTiles, lut, thresholds, binScale, W, grouped, tileCursor);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace pppk
