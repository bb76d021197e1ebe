// scan.cu -- the streaming pass over the dense stack arrays (K2 + K3 of SURVEY.md §8(a)).
//
// k_scan (fused): ONE pass over H+ and H- (8 bytes per genome position) that
//   (K2) builds the height -> frequency histogram of mapHeightsToScores (pingpongpro.cpp:502-509), and
//   (K3) finds, for every plus stack, the minus stacks at overlaps min..max (:559-577), computes the vicinity
//        mean / max (:570-578), the local-height bin (:597-599) and the base-bias bin (:602) and appends
//        one record per (plus stack, minus stack, overlap) (:608).
// Everything of countStacksByGroup that needs the finished histogram (the height-score bin :589-594 and the
// grouped counter :605) is deferred to k_bin_group, which touches only the (few) signature records.
// Fusing the two passes halves the DRAM traffic of "K2 + K3" from 16 to 8 bytes per position.
//
// Tiling: a CTA of 128 threads walks tiles of 4096 positions; each thread owns 32 CONSECUTIVE positions, i.e.
// exactly one 32-bit occupancy word per strand.  They are fetched as four 256-bit streaming loads per strand
// (LDG.E.256, L1::no_allocate, one full 32-byte sector each) through a two-deep register pipeline that also
// prefetches the next tile, and staged in shared memory with the 1-bit/word occupancy mask of the minus strand;
// the window test of a plus stack is then one 64-bit funnel shift over two mask words instead of `W` loads.
// The stack arrays are sparse, but in a 32-wide warp "some lane has a stack here" is almost always true, so
// per-position branches buy nothing.  Instead all per-stack work loops over the SET BITS of the thread's own
// occupancy words (trip count = the busiest lane's stack count; 32 positions per thread amortise it), and all
// per-signature work is made dense: candidates are ranked with a block scan, listed in shared memory, and one
// thread per CANDIDATE evaluates the window and writes its records.  The records of a tile are written
// contiguously (one atomicAdd on a global cursor) in (position, overlap) order and the tile's run is published in
// tileBase/tileCnt; k_bin_group later copies the runs into global position order.
// Histogram: bins < 512 are privatised per CTA in shared memory (height 1, by far the most common, is
// counted in a register), the rare tall stacks go to a global table with red.add.
//
// Bound: HBM bandwidth.  Algorithmic bytes per launch: 8 * G + 16 * P  (G padded positions, P signatures).
#include "common.cuh"

namespace pppk {

namespace {

constexpr int kThreads = 128;
constexpr int kPerThread = 32;                 // consecutive positions owned by one thread = one occupancy word
constexpr int kTile = kScanTile;               // 4096 positions per CTA iteration
constexpr int kHalo = 32;
constexpr int kMaskWords = kTile / 32 + 2;
constexpr int kCandCap = 224;                  // candidates staged per round
static_assert(kTile == kThreads * kPerThread, "tile geometry");

__device__ __forceinline__ void ld_stream256(uint32_t (&r)[8], const uint32_t* p) {  // LDG.E.256, 32-byte aligned
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}

__device__ __forceinline__ uint32_t warp_inclusive(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Shared-memory index of tile position p.  A thread owns one 32-word row; without a swizzle every 128-bit store of a
// warp would hit the same 4 banks (32-way conflict).  The 4-word groups of a row are XOR-permuted with the row
// number, so 8 consecutive rows cover all 32 banks; positions stay addressable one by one (the window reads cross rows).
__device__ __forceinline__ uint32_t sw(uint32_t p) { return p ^ (((p >> 5) & 7u) << 2); }

__device__ __forceinline__ uint32_t nz_byte(const uint32_t (&w)[8]) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) m |= (w[j] != 0u ? 1u : 0u) << j;
    return m;
}

template <int REPR>
__global__ void __launch_bounds__(kThreads, 6) k_scan(ScanParams P) {
    __shared__ __align__(32) uint32_t sP[kTile];          // plus strand of the tile
    __shared__ __align__(32) uint32_t sM[kTile + kHalo];  // minus strand + right halo
    __shared__ uint32_t sMask[kMaskWords];                // 1 bit per minus word: stack present
    __shared__ uint32_t sHist[kScanLowBins];
    __shared__ uint32_t sWarpTot[kThreads / 32];
    __shared__ unsigned long long sBase;
    __shared__ uint32_t sCandBits[kCandCap], sCandOff[kCandCap];  // per candidate: present overlaps / first record (tile-relative)
    __shared__ uint16_t sCandTp[kCandCap];                        // its tile position

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < (int)kScanLowBins; k += kThreads) sHist[k] = 0;
    if (tid == 0) sMask[kMaskWords - 1] = 0;
    __syncthreads();

    const uint32_t ONE = REPR == kReprInt ? 1u : 0x3f800000u;
    const uint32_t wMask = P.W >= 32 ? 0xffffffffu : ((1u << P.W) - 1u);
    const float fW = (float)P.W;
    const uint32_t myPos = (uint32_t)tid * kPerThread;  // this thread owns 32 consecutive positions of the tile
    const uint32_t rowSw = (uint32_t)tid & 7u;          // group permutation of this thread's row (see sw())
    uint32_t ones = 0, localMax = 0, small = 0;  // small: four 8-bit counters for rounded heights 0..3
    auto flush_small = [&]() {
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if ((small >> (8 * r)) & 0xffu) atomicAdd(&sHist[r], (small >> (8 * r)) & 0xffu);
        small = 0;
    };

    // register pipeline: the 32 positions are fetched as 4 sub-steps of 8 words per strand (LDG.E.256); the loads
    // of sub-step s+1 (or of the next tile's sub-step 0) are in flight while sub-step s is staged.
    uint32_t pa[8], ma[8], pb[8], mb[8], hw[8];
    auto issue = [&](uint32_t (&p)[8], uint32_t (&m)[8], uint32_t t, int sub) {
        const uint32_t* src = P.H + (size_t)t * kTile + myPos + sub * 8;
        ld_stream256(p, src);
        ld_stream256(m, src + P.G);
    };
    if (blockIdx.x < P.nTiles) issue(pa, ma, blockIdx.x, 0);

    for (uint32_t tile = blockIdx.x; tile < P.nTiles; tile += gridDim.x) {
        const uint32_t tileBase = tile * (uint32_t)kTile;
        if (tid < kHalo / 8) ld_stream256(hw, P.H + P.G + tileBase + kTile + tid * 8);
        uint32_t nzP = 0, nzM = 0;
        auto stage = [&](const uint32_t (&p)[8], const uint32_t (&m)[8], int sub) {
            const uint32_t g0 = myPos + (((uint32_t)(sub * 2) ^ rowSw) << 2), g1 = myPos + (((uint32_t)(sub * 2 + 1) ^ rowSw) << 2);
            *reinterpret_cast<uint4*>(&sP[g0]) = make_uint4(p[0], p[1], p[2], p[3]);
            *reinterpret_cast<uint4*>(&sP[g1]) = make_uint4(p[4], p[5], p[6], p[7]);
            *reinterpret_cast<uint4*>(&sM[g0]) = make_uint4(m[0], m[1], m[2], m[3]);
            *reinterpret_cast<uint4*>(&sM[g1]) = make_uint4(m[4], m[5], m[6], m[7]);
            nzP |= nz_byte(p) << (sub * 8);
            nzM |= nz_byte(m) << (sub * 8);
        };
        issue(pb, mb, tile, 1);
        stage(pa, ma, 0);
        issue(pa, ma, tile, 2);
        stage(pb, mb, 1);
        issue(pb, mb, tile, 3);
        stage(pa, ma, 2);
        if (tile + gridDim.x < P.nTiles) issue(pa, ma, tile + gridDim.x, 0);  // next tile, in flight during the phases below
        stage(pb, mb, 3);
        sMask[tid] = nzM;
        if (warp == 0) {
            uint32_t v = 0;
            if (lane < kHalo / 8) {
                *reinterpret_cast<uint4*>(&sM[kTile + lane * 8]) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(&sM[kTile + lane * 8 + 4]) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                v = nz_byte(hw) << (8 * lane);
            }
            v |= __shfl_xor_sync(0xffffffffu, v, 1);
            v |= __shfl_xor_sync(0xffffffffu, v, 2);
            if (lane == 0) sMask[kTile / 32] = v;
        }

        // K2: height histogram.  Loop over the set bits of this thread's own occupancy words (iterations = the
        // busiest lane's stack count, not 64), re-reading the word it staged itself.
        {
            uint32_t mm = nzM;
            if (tileBase < P.haloHi && tileBase + kTile > P.haloLo) {  // halo copies of a neighbouring shard are not counted
                uint32_t first = tileBase + myPos;
                uint32_t lo = P.haloLo > first ? P.haloLo - first : 0u, hi = P.haloHi > first ? P.haloHi - first : 0u;
                uint32_t below = lo >= 32 ? 0xffffffffu : ((1u << lo) - 1u), upto = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
                mm &= below | ~upto;
            }
            // both strands advance together (two independent load -> classify chains per trip)
            uint32_t nzA = nzP, nzB = mm;
            while (nzA | nzB) {
#pragma unroll
                for (int strand = 0; strand < 2; ++strand) {
                    uint32_t& nz = strand ? nzB : nzA;
                    if (!nz) continue;
                    const uint32_t* src = strand ? sM : sP;
                    int b = __ffs(nz) - 1;
                    nz &= nz - 1;
                    uint32_t wv = src[myPos + ((uint32_t)b ^ (rowSw << 2))];
                    if ((wv & kValueMask) == ONE) { ++ones; continue; }
                    uint32_t r = word_round_height<REPR>(wv);
                    if (r < 4u) { small += 1u << (8 * r); if ((small & 0x80808080u) != 0u) flush_small(); }  // heights 0..3: byte counters in a register
                    else if (r < kScanLowBins) atomicAdd(&sHist[r], 1u);
                    else atomicAdd(&P.freqTail[r], 1u);
                    localMax = max(localMax, r);
                }
            }
        }
        __syncthreads();  // (A) both strands + mask are complete

        // K3: candidates = plus stacks with at least one minus stack at overlaps min..max
        uint32_t cnt = 0, ncand = 0, cand = 0;
        {
            uint32_t nz = nzP;
            while (nz) {
                int b = __ffs(nz) - 1;
                nz &= nz - 1;
                uint32_t bit = myPos + b + P.minOv;
                uint32_t bits = __funnelshift_r(sMask[bit >> 5], sMask[(bit >> 5) + 1], bit & 31) & wMask;
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
                    uint32_t bits = __funnelshift_r(sMask[bit >> 5], sMask[(bit >> 5) + 1], bit & 31) & wMask;
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
                    const uint32_t wp = sP[sw(tp)];
                    const uint32_t w0 = tp + P.minOv;  // first window position (tile-relative)
                    // vicinity sum / max over the PRESENT stacks only, in overlap order: an absent stack would add
                    // +0.0f, which is an exact no-op (:564-577)
                    float sum = 0.0f, mx = 0.0f;
                    for (uint32_t rest = bits; rest; rest &= rest - 1) {
                        float mq = word_height<REPR>(sM[sw(w0 + (uint32_t)(__ffs(rest) - 1))]);
                        sum = __fadd_rn(sum, mq);
                        mx = fmaxf(mx, mq);
                    }
                    const float mean = __fdiv_rn(sum, fW);                    // :578
                    const float hp = word_height<REPR>(wp);
                    unsigned long long idx = sBase + sCandOff[c];
                    while (bits) {                                            // :584-586
                        const int o = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const uint32_t wm = sM[sw(w0 + o)];
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
        __syncthreads();  // (E) the staged tile is no longer read
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
            const uint32_t items = (cnt + 63u) >> 6;
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
            const uint32_t cnt = sCnt[t], base = (item - sItemStart[t]) << 6;
            const size_t src = sSrc[t], dst = sDst[t];
            const uint32_t i0 = base + lane, i1 = base + 32 + lane;  // two records per lane in flight
            const bool ok0 = i0 < cnt, ok1 = i1 < cnt;
            PairRec r0 = {0, 1.f, 1.f, 0}, r1 = {0, 1.f, 1.f, 0};
            if (ok0) *reinterpret_cast<uint4*>(&r0) = *reinterpret_cast<const uint4*>(&raw[src + i0]);
            if (ok1) *reinterpret_cast<uint4*>(&r1) = *reinterpret_cast<const uint4*>(&raw[src + i1]);
            const float fp0 = lut[round_height(r0.rp)], fm0 = lut[round_height(r0.rm)];  // heightScoreMap[0.5 + reads], :582 / :589
            const float fp1 = lut[round_height(r1.rp)], fm1 = lut[round_height(r1.rm)];
            uint32_t cell0 = 0, cell1 = 0;
            if (ok0) {
                int b = bin_of(__fmul_rn(fp0, fm0));
                r0.misc |= (uint32_t)b << 7;
                cell0 = ((((uint32_t)b * (uint32_t)W + (r0.misc & 31u)) * 2u + ((r0.misc >> 6) & 1u)) * 2u) + ((r0.misc >> 5) & 1u);  // bin-major table
                *reinterpret_cast<uint4*>(&ordered[dst + i0]) = *reinterpret_cast<uint4*>(&r0);
            }
            if (ok1) {
                int b = bin_of(__fmul_rn(fp1, fm1));
                r1.misc |= (uint32_t)b << 7;
                cell1 = ((((uint32_t)b * (uint32_t)W + (r1.misc & 31u)) * 2u + ((r1.misc >> 6) & 1u)) * 2u) + ((r1.misc >> 5) & 1u);
                *reinterpret_cast<uint4*>(&ordered[dst + i1]) = *reinterpret_cast<uint4*>(&r1);
            }
            count_cell(ok0, cell0);
            count_cell(ok1, cell1);
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
    int perSm = 6;
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
