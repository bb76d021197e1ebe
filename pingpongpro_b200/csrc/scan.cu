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
// Tiling: a CTA of 256 threads walks tiles of 2048 positions.  Each thread owns 8 consecutive positions and
// loads them with ONE streaming 256-bit load per strand (LDG.E.256, L1::no_allocate): a warp-wide load covers
// 1 KB of contiguous memory.  Both strands of the tile plus a 32-word right halo of the minus strand are staged
// in shared memory together with a 1-bit/word occupancy mask of the minus strand (built with warp shuffles), so
// the window test of a plus stack is one 64-bit funnel shift instead of `W` loads.
// The kernel is instruction-issue sensitive (the stack arrays are sparse, but in a 32-wide warp "some lane has a
// stack" is almost always true), so all per-stack work loops over the SET BITS of the thread's own occupancy
// mask, and all per-signature work is made dense: candidates write (position, overlap) entries in record
// order into a shared list (block scan of the counts), then one thread per RECORD evaluates the window.
// Shared memory is double buffered: two barriers per tile (+1 when the tile has signatures).  The records of a
// tile are written contiguously (one atomicAdd on a global cursor) in (position, overlap) order and the tile's
// run is published in tileBase/tileCnt; k_bin_group later copies the runs into global position order.
// Histogram: bins < 2048 are privatised per CTA in shared memory (height 1, by far the most common, is
// counted in a register), the rare tall stacks go to a global table with red.add.
//
// Bound: HBM bandwidth.  Algorithmic bytes per launch: 8 * G + 16 * P  (G padded positions, P signatures).
#include "common.cuh"

namespace pppk {

namespace {

constexpr int kThreads = 256;
constexpr int kTile = 2048;
constexpr int kHalo = 32;
constexpr int kMaskWords = kTile / 32 + 2;
constexpr int kCandCap = 512;   // candidates staged per round (a tile rarely has more)

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

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

__device__ __forceinline__ uint32_t nz_nibble(const uint4& v) {
    return (v.x != 0u ? 1u : 0u) | (v.y != 0u ? 2u : 0u) | (v.z != 0u ? 4u : 0u) | (v.w != 0u ? 8u : 0u);
}

template <int REPR>
__global__ void __launch_bounds__(kThreads, 4) k_scan(ScanParams P) {
    __shared__ __align__(32) uint32_t sP[2][kTile];          // plus strand of the tile
    __shared__ __align__(32) uint32_t sM[2][kTile + kHalo];  // minus strand + right halo
    __shared__ uint32_t sMask[2][kMaskWords];                // 1 bit per minus word: stack present
    __shared__ uint32_t sHist[kLowBins];
    __shared__ uint32_t sWarpTot[kThreads / 32];
    __shared__ unsigned long long sBase;
    __shared__ uint32_t sCandBits[kCandCap];                 // per candidate (plus stack with partners): present overlaps
    __shared__ uint16_t sCandTp[kCandCap], sCandOff[kCandCap];  // its tile position / first record of the tile it writes

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < (int)kLowBins; k += kThreads) sHist[k] = 0;
    if (tid < 2) sMask[tid][kMaskWords - 1] = 0;
    __syncthreads();

    const uint32_t ONE = REPR == kReprInt ? 1u : 0x3f800000u;
    const uint32_t wMask = P.W >= 32 ? 0xffffffffu : ((1u << P.W) - 1u);
    const float fW = (float)P.W;
    const uint32_t myPos = (uint32_t)tid * 8u;  // this thread owns 8 consecutive positions of the tile
    uint32_t ones = 0, localMax = 0;

    // software pipeline: the loads of tile t+1 are issued as soon as tile t has been staged, so they are in
    // flight while tile t is processed
    uint32_t pw[8], mw[8];
    uint4 h = make_uint4(0, 0, 0, 0);
    auto issue_loads = [&](uint32_t t) {
        const uint32_t tb = t * (uint32_t)kTile;
        ld_stream256(pw, P.H + tb + myPos);
        ld_stream256(mw, P.H + P.G + tb + myPos);
        if (tid < kHalo / 4) h = ld_stream(reinterpret_cast<const uint4*>(P.H + P.G + tb + kTile) + tid);
    };
    if (blockIdx.x < P.nTiles) issue_loads(blockIdx.x);

    int buf = 0;
    for (uint32_t tile = blockIdx.x; tile < P.nTiles; tile += gridDim.x, buf ^= 1) {
        const uint32_t tileBase = tile * (uint32_t)kTile;

        // stage both strands (the rare paths below index them dynamically) + the minus occupancy mask
        *reinterpret_cast<uint4*>(&sP[buf][myPos]) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
        *reinterpret_cast<uint4*>(&sP[buf][myPos + 4]) = make_uint4(pw[4], pw[5], pw[6], pw[7]);
        *reinterpret_cast<uint4*>(&sM[buf][myPos]) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
        *reinterpret_cast<uint4*>(&sM[buf][myPos + 4]) = make_uint4(mw[4], mw[5], mw[6], mw[7]);
        uint32_t nzP = 0, nzM = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            nzP |= (pw[j] != 0u ? 1u : 0u) << j;
            nzM |= (mw[j] != 0u ? 1u : 0u) << j;
        }
        {
            uint32_t v = nzM << (8 * (lane & 3));
            v |= __shfl_xor_sync(0xffffffffu, v, 1);
            v |= __shfl_xor_sync(0xffffffffu, v, 2);
            if ((lane & 3) == 0) sMask[buf][warp * 8 + (lane >> 2)] = v;
        }
        if (warp == 0) {
            if (lane < kHalo / 4) *reinterpret_cast<uint4*>(&sM[buf][kTile + lane * 4]) = h;
            uint32_t v = lane < kHalo / 4 ? nz_nibble(h) << (4 * lane) : 0u;
            v |= __shfl_xor_sync(0xffffffffu, v, 1);
            v |= __shfl_xor_sync(0xffffffffu, v, 2);
            v |= __shfl_xor_sync(0xffffffffu, v, 4);
            if (lane == 0) sMask[buf][kTile / 32] = v;
        }
        if (tile + gridDim.x < P.nTiles) issue_loads(tile + gridDim.x);  // prefetch (pw / mw / h are dead from here on)

        // K2: height histogram.  Loop over the set bits of this thread's own occupancy mask (iterations = the
        // busiest lane's stack count, not 16), re-reading the word it staged itself.
        {
            uint32_t nz = nzP | (nzM << 8);
            if (tileBase < P.haloHi && tileBase + kTile > P.haloLo) {  // halo copies of a neighbouring shard are not counted
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t lp = tileBase + myPos + j;
                    if (lp >= P.haloLo && lp < P.haloHi) nz &= ~(256u << j);
                }
            }
            while (nz) {
                int b = __ffs(nz) - 1;
                nz &= nz - 1;
                uint32_t wv = (b & 8) ? sM[buf][myPos + (b & 7)] : sP[buf][myPos + b];
                if ((wv & kValueMask) == ONE) { ++ones; continue; }
                uint32_t r = word_round_height<REPR>(wv);
                if (r < kLowBins) atomicAdd(&sHist[r], 1u);
                else atomicAdd(&P.freqTail[r], 1u);
                localMax = max(localMax, r);
            }
        }
        __syncthreads();  // (A) both strands + mask of this buffer are complete

        // K3: candidates = plus stacks with at least one minus stack at overlaps min..max
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
        // one warp scan carries both counters (records < 2^14 per warp in the low half, candidates above)
        const uint32_t packed = cnt | (ncand << 16);
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
        const uint32_t total = totalP & 0xffffu, candTotal = totalP >> 16;
        if (tid == 0) {
            unsigned long long base = total ? atomicAdd(P.cursor, (unsigned long long)total) : 0ull;
            sBase = base;
            P.tileCnt[tile] = total;
            P.tileBase[tile] = (uint32_t)base;
        }
        if (total) {  // block-uniform: the rare path, processed densely (one thread per candidate)
            const uint32_t mine = beforeP + inc - packed;
            const uint32_t firstRec = mine & 0xffffu, firstCand = mine >> 16;  // tile-relative ranks of this thread's first record / candidate
            for (uint32_t cbase = 0; cbase < candTotal; cbase += kCandCap) {
                if (cbase) __syncthreads();  // the previous round's list is no longer read
                uint32_t off = firstRec, rank = firstCand, cm = cand;
                while (cm) {
                    int b = __ffs(cm) - 1;
                    cm &= cm - 1;
                    uint32_t tp = myPos + b, bit = tp + P.minOv;
                    uint32_t bits = __funnelshift_r(sMask[buf][bit >> 5], sMask[buf][(bit >> 5) + 1], bit & 31) & wMask;
                    uint32_t slot = rank - cbase;  // wraps to a huge value for earlier rounds
                    if (slot < (uint32_t)kCandCap) { sCandTp[slot] = (uint16_t)tp; sCandBits[slot] = bits; sCandOff[slot] = (uint16_t)off; }
                    off += __popc(bits);
                    ++rank;
                }
                __syncthreads();  // (C)
                const uint32_t nIn = min((uint32_t)kCandCap, candTotal - cbase);
                for (uint32_t c = tid; c < nIn; c += kThreads) {
                    const uint32_t tp = sCandTp[c];
                    uint32_t bits = sCandBits[c];
                    const uint32_t wp = sP[buf][tp];
                    const uint32_t* win = &sM[buf][tp + P.minOv];
                    float sum = 0.0f, mx = 0.0f;
                    for (int q = 0; q < P.W; ++q) {                          // :564-577 (absent stacks add +0)
                        float mq = word_height<REPR>(win[q]);
                        sum = __fadd_rn(sum, mq);
                        mx = fmaxf(mx, mq);
                    }
                    const float mean = __fdiv_rn(sum, fW);                    // :578
                    const float hp = word_height<REPR>(wp);
                    unsigned long long idx = sBase + sCandOff[c];
                    while (bits) {                                            // :584-586
                        const int o = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const uint32_t wm = win[o];
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
    if (ones) { atomicAdd(&sHist[1], ones); localMax = max(localMax, 1u); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) localMax = max(localMax, __shfl_xor_sync(0xffffffffu, localMax, d));
    if (lane == 0 && localMax) atomicMax(P.maxHeight, localMax);
    __syncthreads();
    for (int k = tid; k < (int)kLowBins; k += kThreads) {
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
        unsigned long long c = h < kLowBins ? freqLow[h] : (unsigned long long)freqTail[h];
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
// per-tile runs into global (position, overlap) order.  A block handles 256 consecutive tiles.
constexpr int kBgThreads = 256;
__global__ void __launch_bounds__(kBgThreads) k_bin_group(const PairRec* __restrict__ raw, PairRec* __restrict__ ordered, const uint32_t* __restrict__ tileCnt,
                                                          const uint32_t* __restrict__ tileBase, const uint32_t* __restrict__ tileDst, uint32_t nTiles,
                                                          const float* __restrict__ lut, const float* __restrict__ thresholds, uint32_t* __restrict__ grouped) {
    __shared__ float sThr[kBins];
    __shared__ uint32_t sStart[kBgThreads + 1];  // exclusive prefix of the block's tile counts
    __shared__ uint32_t sSrc[kBgThreads];
    __shared__ uint32_t sWarp[kBgThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < kBins - 1; k += kBgThreads) sThr[k] = thresholds[k];
    uint32_t tile = blockIdx.x * kBgThreads + tid;
    uint32_t c = tile < nTiles ? tileCnt[tile] : 0u;
    sSrc[tid] = tile < nTiles ? tileBase[tile] : 0u;
    uint32_t inc = warp_inclusive(c, lane);
    if (lane == 31) sWarp[warp] = inc;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kBgThreads / 32; ++w) {
        uint32_t t = sWarp[w];
        if (w < warp) before += t;
        total += t;
    }
    sStart[tid] = before + inc - c;
    if (tid == 0) sStart[kBgThreads] = total;
    __syncthreads();
    if (total == 0) return;
    const uint32_t dstBase = tileDst[blockIdx.x * kBgThreads];  // ordered index of the block's first record
    const uint32_t rounds = (total + kBgThreads - 1) / kBgThreads;
    for (uint32_t it = 0; it < rounds; ++it) {
        uint32_t item = it * kBgThreads + tid;
        bool ok = item < total;
        PairRec r;
        uint32_t cell = 0;
        if (ok) {
            int lo = 0, hi = kBgThreads;  // last tile slot whose start <= item
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (sStart[mid] <= item) lo = mid; else hi = mid;
            }
            *reinterpret_cast<uint4*>(&r) = *reinterpret_cast<const uint4*>(&raw[(size_t)sSrc[lo] + (item - sStart[lo])]);
            float fp = lut[round_height(r.rp)];                       // heightScoreMap[0.5 + reads], :582
            float fm = lut[round_height(r.rm)];                       // :589
            float hs = __fmul_rn(fp, fm);
            int b = 0, e = kBins - 1;                                 // bin = #thresholds <= hs (host-built, see api.cu)
            while (b < e) {
                int mid = (b + e) >> 1;
                if (sThr[mid] <= hs) b = mid + 1; else e = mid;
            }
            r.misc |= (uint32_t)b << 7;
            cell = (((r.misc & 31u) * kBins + (uint32_t)b) * 2u + ((r.misc >> 6) & 1u)) * 2u + ((r.misc >> 5) & 1u);
            *reinterpret_cast<uint4*>(&ordered[(size_t)dstBase + item]) = *reinterpret_cast<uint4*>(&r);
        }
        unsigned act = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            unsigned peers = __match_any_sync(act, cell);
            if (lane == __ffs(peers) - 1) atomicAdd(&grouped[cell], (uint32_t)__popc(peers));  // :605
        }
    }
}

}  // namespace

cudaError_t launch_scan(const ScanParams& p, int repr, cudaStream_t s, int* launches) {
    if (p.nTiles == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int perSm = 4;
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
                             const float* lut, const float* thresholds, uint32_t* grouped, cudaStream_t s, int* launches) {
    if (nTiles == 0) return cudaSuccess;
    unsigned blocks = (nTiles + kBgThreads - 1) / kBgThreads;
    k_bin_group<<<blocks, kBgThreads, 0, s>>>(raw, ordered, tileCnt, tileBase, tileDst, nTiles, lut, thresholds, grouped);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace pppk
