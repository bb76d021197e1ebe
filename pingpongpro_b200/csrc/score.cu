// score.cu -- K4: collapseBins (pingpongpro.cpp:631-686), calculateFDRs (:695-754) and the selection of
// the reported loci (:901-903), all on the device so that scan -> score needs no host round trip.
//
//   k_collapse_*  the serial merge decision of :670 runs on per-bin occupancy bit masks (one warp, prefix-OR);
//                 the float additions of :658 (counters clipped at 2^24) are exact integer sums below 2^24.
//   k_fdr_table   one warp per (collapsed bin, bias, local, overlap): background mean / sigma (:705-718)
//                 and the 1001-step Gaussian integral of :724-734: terms evaluated in parallel with IEEE double
//                 operations (__dadd_rn/__dmul_rn/__ddiv_rn, no FMA), summed by one lane in the reference's
//                 order; x and the weights 0.01/sqrt(2 pi) * exp(-x^2/2) are tabulated by the host with the
//                 host libm.
//   k_score_*     per-signature FDR lookup (:750-753) fused with the -s filter (:903) and an
//                 order-preserving compaction (count -> scan -> write) of the overlap-10 loci.
// All of it is latency-bound bookkeeping on kilobytes .. a few megabytes.
#include <algorithm>

#include "common.cuh"

namespace pppk {

namespace {

// Device tables are BIN-MAJOR: cell(bin, overlap, bias, local) = ((bin * W + overlap) * 2 + bias) * 2 + local, so
// that the W*4 cells of one height-score bin are contiguous (the ABI transposes to [overlap][bin][bias][local]).
constexpr int kMaxCellWords = (PPP_MAX_OVERLAPS * 4 + 31) / 32;  // occupancy bits of one bin's cells

// collapseBins, pingpongpro.cpp:631-686.  A merged bin is closed at the first source bin after which none of its
// W*4 cells is <= 0 (:661, :670).  Counts are >= 0 and float additions of non-negative numbers never return to 0,
// so that decision depends only on WHICH cells of each source bin are non-zero:
//   k_collapse_occ   one occupancy bit per (bin, cell);
//   k_collapse_map   one warp replays the serial walk of :641-673 on those bits, 32 bins at a time with a prefix-OR
//                    over the lanes -> oldBinCollapsedBinMap, group starts, C;
//   k_collapse_cells one warp per cell of a merged bin adds min(count, 2^24) (the reference's float counter) of its source
//                    bins: exactly when the sum stays below 2^24, else with the literal in-order float additions of :658.
// one occupancy bit per (bin, cell): one WARP per bin (many blocks: the table is 336 KB of L2-resident counters, and one
// block reading all of it was 30 us of load latency)
__global__ void __launch_bounds__(256) k_collapse_occ(const uint32_t* __restrict__ grouped, int W, uint32_t* __restrict__ binOcc /*[kBins][kMaxCellWords]*/) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= kBins) return;
    const int cellsPerBin = W * 4;
#pragma unroll
    for (int w = 0; w < kMaxCellWords; ++w) {
        const int c = w * 32 + lane;
        const uint32_t v = c < cellsPerBin ? grouped[(size_t)b * cellsPerBin + c] : 0u;
        const unsigned m = __ballot_sync(0xffffffffu, v != 0u);
        if (lane == 0) binOcc[b * kMaxCellWords + w] = m;
    }
}

__global__ void __launch_bounds__(256) k_collapse_map(const uint32_t* __restrict__ binOcc, int W, uint16_t* __restrict__ binMap,
                                                      uint16_t* __restrict__ groupStart, uint32_t* __restrict__ nBins) {
    __shared__ uint32_t sOcc[kBins * kMaxCellWords];  // (one round of parallel loads instead of one L2 latency per 32 bins)
    const int cellsPerBin = W * 4, tid = threadIdx.x;
    {
        constexpr int kRounds = (kBins * kMaxCellWords + 255) / 256;
        uint32_t v[kRounds];
#pragma unroll
        for (int r = 0; r < kRounds; ++r) { const int k = r * 256 + tid; v[r] = k < kBins * kMaxCellWords ? binOcc[k] : 0u; }  // all loads in flight together
#pragma unroll
        for (int r = 0; r < kRounds; ++r) { const int k = r * 256 + tid; if (k < kBins * kMaxCellWords) sOcc[k] = v[r]; }
    }
    __syncthreads();
    if (tid >= 32) return;
    const int lane = tid;
    uint32_t full[kMaxCellWords], carry[kMaxCellWords];
#pragma unroll
    for (int w = 0; w < kMaxCellWords; ++w) {
        int bitsHere = cellsPerBin - 32 * w;
        full[w] = bitsHere >= 32 ? 0xffffffffu : (bitsHere > 0 ? ((1u << bitsHere) - 1u) : 0u);
        carry[w] = 0;
    }
    uint32_t collapsed = 0;
    int lastClose = -1;
    if (lane == 0) groupStart[0] = 0;
    for (int chunk = 0; chunk * 32 < kBins; ++chunk) {
        const int b = chunk * 32 + lane;
        const bool valid = b < kBins;
        uint32_t m[kMaxCellWords];
#pragma unroll
        for (int w = 0; w < kMaxCellWords; ++w) m[w] = valid ? sOcc[b * kMaxCellWords + w] : 0u;
        {   // most chunks cannot close a merged bin (all 32 bins together still leave a cell empty): no prefix scan for those
            bool canClose = true;
#pragma unroll
            for (int w = 0; w < kMaxCellWords; ++w) {
                const uint32_t all = __reduce_or_sync(0xffffffffu, m[w]) | carry[w];
                canClose = canClose && all == full[w];
            }
            if (!canClose) {
                if (valid) binMap[b] = (uint16_t)collapsed;  // :666
#pragma unroll
                for (int w = 0; w < kMaxCellWords; ++w) carry[w] |= __reduce_or_sync(0xffffffffu, m[w]);
                continue;
            }
        }
        int start = 0;
        while (start < 32) {
            uint32_t pre[kMaxCellWords];
#pragma unroll
            for (int w = 0; w < kMaxCellWords; ++w) pre[w] = lane >= start ? m[w] : 0u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                for (int w = 0; w < kMaxCellWords; ++w) {
                    uint32_t t = __shfl_up_sync(0xffffffffu, pre[w], d);
                    if (lane >= d) pre[w] |= t;
                }
            }
            bool complete = valid && lane >= start;
#pragma unroll
            for (int w = 0; w < kMaxCellWords; ++w) { pre[w] |= carry[w]; complete &= pre[w] == full[w]; }  // no cell <= 0 (:661)
            const unsigned closes = __ballot_sync(0xffffffffu, complete);
            const int fc = closes ? __ffs(closes) - 1 : 31;
            if (valid && lane >= start && lane <= fc) binMap[b] = (uint16_t)collapsed;  // :666
            if (!closes) {
#pragma unroll
                for (int w = 0; w < kMaxCellWords; ++w) carry[w] = __shfl_sync(0xffffffffu, pre[w], 31);
                break;
            }
            ++collapsed;                                                                 // :670-672
            lastClose = chunk * 32 + fc;
            if (lane == 0 && lastClose + 1 < kBins) groupStart[collapsed] = (uint16_t)(lastClose + 1);
#pragma unroll
            for (int w = 0; w < kMaxCellWords; ++w) carry[w] = 0;
            start = fc + 1;
        }
    }
    if (lastClose != kBins - 1) ++collapsed;  // the loop of :641 ends with the last bin, whatever is still empty
    if (lane == 0) { groupStart[collapsed] = (uint16_t)kBins; *nBins = collapsed; }
}

// The cells of the merged bins (:658: float additions of the source bins' counters, clipped at 2^24 like the reference's
// float counters, in bin order): one WARP per (merged bin, cell).  The lanes add the source bins exactly in 64-bit
// integers; a sum below 2^24 IS the result of the in-order float additions (no rounding can have happened), a larger one
// is recomputed by one lane with the literal sequence of float additions.
__global__ void __launch_bounds__(256) k_collapse_cells(const uint32_t* __restrict__ grouped, int W, const uint16_t* __restrict__ groupStart,
                                                        const uint32_t* __restrict__ nBins, float* __restrict__ cells) {
    const int cellsPerBin = W * 4, lane = threadIdx.x & 31;
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= (int)*nBins * cellsPerBin) return;
    const int g = k / cellsPerBin, c = k - g * cellsPerBin;
    const int first = groupStart[g], last = groupStart[g + 1];
    unsigned long long total = 0;
    for (int b0 = first + lane; b0 < last; b0 += 32 * 8) {  // eight loads in flight per lane
        uint32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int b = b0 + 32 * u; v[u] = b < last ? grouped[(size_t)b * cellsPerBin + c] : 0u; }
#pragma unroll
        for (int u = 0; u < 8; ++u) total += v[u] < kSaturate ? v[u] : kSaturate;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
    if (lane) return;
    float a;
    if (total < kSaturate) {
        a = (float)(uint32_t)total;
    } else {
        a = 0.f;
        for (int b = first; b < last; ++b) {
            const uint32_t v = grouped[(size_t)b * cellsPerBin + c];
            a = __fadd_rn(a, (float)(v < kSaturate ? v : kSaturate));
        }
    }
    cells[k] = a;
}

// calculateFDRs, pingpongpro.cpp:700-747.  One warp per (merged bin, bias, local, overlap).  The 1001 integrand terms of
// :726-733 are independent: the lanes evaluate them (IEEE double ops, no FMA; x and w(x) from the host libm) into
// shared memory, then ONE lane adds them in the reference's order, so the sum is the same sequence of fp64 additions.
constexpr int kFdrWarps = 4;
constexpr int kFdrMaxSteps = 1008;
__global__ void __launch_bounds__(kFdrWarps * 32) k_fdr_table(const float* __restrict__ cells, const uint32_t* __restrict__ nBins, int W, int ppIdx,
                                                               const double* __restrict__ xs, const double* __restrict__ ws, int nSteps, float* __restrict__ fdr) {
    __shared__ double sTerm[kFdrWarps][kFdrMaxSteps];
    __shared__ double sX[kFdrMaxSteps], sWt[kFdrMaxSteps];  // the integration grid and its Gaussian weights (host libm), once per block
    {
        constexpr int kRounds = (kFdrMaxSteps + kFdrWarps * 32 - 1) / (kFdrWarps * 32);
        double vx[kRounds], vw[kRounds];
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {  // all loads in flight together
            const int i = r * kFdrWarps * 32 + threadIdx.x;
            vx[r] = i < nSteps ? xs[i] : 0.0;
            vw[r] = i < nSteps ? ws[i] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            const int i = r * kFdrWarps * 32 + threadIdx.x;
            if (i < nSteps) { sX[i] = vx[r]; sWt[i] = vw[r]; }
        }
    }
    __syncthreads();
    const uint32_t C = *nBins;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t item = blockIdx.x * kFdrWarps + warp; item < C * 4u * (uint32_t)W; item += gridDim.x * kFdrWarps) {
    const uint32_t ov = item % (uint32_t)W, cell = item / (uint32_t)W;  // cell = bin * 4 + bias * 2 + local
    const uint32_t bin = cell >> 2, k = cell & 3u;
    const float mine = lane < W ? cells[((size_t)bin * W + lane) * 4 + k] : 0.f;  // lane o holds c_o: one round of loads
    auto at = [&](int o) { return __shfl_sync(0xffffffffu, mine, o); };
    float mean = 0.f;
    for (int o = 0; o < W; ++o)
        if (o != ppIdx) mean = __fadd_rn(mean, at(o));                       // :705-708
    mean = __fdiv_rn(mean, (float)(W - 1));                                  // :709
    float sd = 0.f;
    for (int o = 0; o < W; ++o)
        if (o != ppIdx) {
            double d = (double)__fsub_rn(at(o), mean);
            sd = __double2float_rn(__dadd_rn((double)sd, __dmul_rn(d, d)));  // :715 (float += double)
        }
    sd = __double2float_rn(sqrt(__dmul_rn(__ddiv_rn(1.0, (double)(W - 2)), (double)sd)));  // :716 (sqrt is IEEE-exact)
    if ((double)sd <= 1E-10) sd = (float)1E-10;                              // :717-718
    const float c = at((int)ov);
    const double thr = (double)__fdiv_rn(__fsub_rn(c, mean), sd);            // :726
    const double dsd = (double)sd, dmean = (double)mean, dc = (double)c;
    for (int i = lane; i < nSteps; i += 32) {                                // :724
        const double x = sX[i], w = sWt[i];
        sTerm[warp][i] = (x >= thr) ? w                                                           // :728
                                    : __ddiv_rn(__dmul_rn(w, __dadd_rn(__dmul_rn(x, dsd), dmean)), dc);  // :732
    }
    __syncwarp();
    if (lane == 0) {
        double acc = 0.0;
        int i = 0;
        for (; i + 8 <= nSteps; i += 8) {  // eight loads in flight, the additions in the reference's order
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = sTerm[warp][i + u];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, t[u]);
        }
        for (; i < nSteps; ++i) acc = __dadd_rn(acc, sTerm[warp][i]);
        if (acc < 0.0) acc = 0.0; else if (acc > 1.0) acc = 1.0;             // :736-743 (NaN stays NaN)
        fdr[((size_t)bin * W + ov) * 4 + k] = __double2float_rn(acc);        // :745
    }
    __syncwarp();  // the terms are rewritten by the warp's next item
    }
}

// K4, per-locus half: the reported loci = the overlap-`pingpong` records (the stream k_scan wrote in position order) that
// pass the -s filter (:903), each with the FDR of its (collapsed bin, bias, local) cell (:750-753), compacted in order in
// ONE pass: blocks claim tiles of 1024 records from a ticket counter and find their first output slot by decoupled
// look-back (see scan.cu).  The height-score bin of a record is re-evaluated here (same thresholds, same step function as
// k_bin) instead of being fetched from the all-overlap record arrays.
constexpr int kScoreThreads = 256;
constexpr int kScoreRounds = 4;
constexpr int kScoreTile = kScoreThreads * kScoreRounds;
constexpr unsigned long long kFlagAggregate = 1ull << 62, kFlagPrefix = 2ull << 62, kFlagValue = (1ull << 62) - 1;

__global__ void __launch_bounds__(kScoreThreads) k_score(const PairRec* __restrict__ pp, const unsigned long long* __restrict__ totals, uint64_t ppCapacity,
                                                         ScoreState* __restrict__ st, unsigned long long* __restrict__ tileState, FreqTables freq,
                                                         const float* __restrict__ thresholds, const BinParams* __restrict__ params, const float* __restrict__ fdrTable,
                                                         const uint16_t* __restrict__ binMap, int W, int ppIdx, float minHeight, const Slot* __restrict__ slots,
                                                         uint32_t nSlots, LocusOut* __restrict__ loci) {
    __shared__ float sThr[kBins];
    __shared__ uint16_t sSmallBin[16];
    __shared__ uint32_t sCnt[kScoreRounds * (kScoreThreads / 32) + 1];
    __shared__ uint32_t sTile;
    __shared__ unsigned long long sBase;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        float v[(kBins + kScoreThreads - 1) / kScoreThreads];
#pragma unroll
        for (int r = 0; r < (kBins + kScoreThreads - 1) / kScoreThreads; ++r) { const int k = r * kScoreThreads + tid; v[r] = k < kBins - 1 ? thresholds[k] : 0.0f; }
#pragma unroll
        for (int r = 0; r < (kBins + kScoreThreads - 1) / kScoreThreads; ++r) { const int k = r * kScoreThreads + tid; if (k < kBins - 1) sThr[k] = v[r]; }
    }
    const float binScale = params->binScale;
    __syncthreads();
    auto bin_of = [&](float hs) {
        int b = (int)(__log10f(hs) * binScale + 0.5f);             // proposal only
        b = min(max(b, 0), kBins - 1);
        while (b > 0 && sThr[b - 1] > hs) --b;                    // exact: bin = #thresholds <= hs
        while (b < kBins - 1 && sThr[b] <= hs) ++b;
        return b;
    };
    if (tid < 16) sSmallBin[tid] = (uint16_t)bin_of(__fmul_rn(freq_of(freq, tid >> 2), freq_of(freq, tid & 3)));
    unsigned long long n = totals[1];
    if (n > ppCapacity) n = ppCapacity;
    const uint32_t nTiles = (uint32_t)((n + kScoreTile - 1) / kScoreTile);
    const bool keepAll = !(minHeight > 0.0f);  // -s 0 (the default): every record is reported (heights are > 0), record i is locus i
    for (uint32_t staticTile = blockIdx.x;; staticTile += gridDim.x) {
        __syncthreads();  // sTile / sCnt / sBase of the previous tile are no longer read; sSmallBin is complete
        if (!keepAll) {
            if (tid == 0) sTile = atomicAdd(&st->ticket, 1u);
            __syncthreads();
        }
        const uint32_t tile = keepAll ? staticTile : sTile;
        if (tile >= nTiles) break;
        const size_t first = (size_t)tile * kScoreTile;
        PairRec r[kScoreRounds];
        bool keep[kScoreRounds];
        unsigned ballot[kScoreRounds];
#pragma unroll
        for (int k = 0; k < kScoreRounds; ++k) {  // record order inside the tile = (round, thread)
            const size_t i = first + (size_t)k * kScoreThreads + tid;
            keep[k] = false;
            if (i < n) {
                *reinterpret_cast<uint4*>(&r[k]) = *reinterpret_cast<const uint4*>(&pp[i]);
                keep[k] = r[k].rp >= minHeight && r[k].rm >= minHeight;  // :903
            }
            ballot[k] = __ballot_sync(0xffffffffu, keep[k]);
            if (lane == 0) sCnt[k * (kScoreThreads / 32) + warp] = keepAll ? (uint32_t)(k * kScoreThreads + warp * 32) : (uint32_t)__popc(ballot[k]);
        }
        if (keepAll) {
            if (tid == 0) sBase = first;
            __syncthreads();
        } else {
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the 32 (round, warp) counts, then the look-back over the tiles before this one
            const uint32_t c = sCnt[lane];
            uint32_t inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
            sCnt[lane] = inc - c;
            unsigned long long before = 0;
            if (tile == 0) {
                if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(tileState) = kFlagPrefix | total;
            } else {
                if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(tileState + tile) = kFlagAggregate | total;
                int64_t j = (int64_t)tile - 1;
                while (j >= 0) {
                    const int64_t mine = j - lane;
                    unsigned long long v = kFlagPrefix;
                    if (mine >= 0) {
                        v = *reinterpret_cast<const volatile unsigned long long*>(tileState + mine);
                        while ((v >> 62) == 0ull) v = *reinterpret_cast<const volatile unsigned long long*>(tileState + mine);
                    }
                    const unsigned prefixLanes = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
                    const int stop = prefixLanes ? __ffs(prefixLanes) - 1 : 32;
                    unsigned long long part = lane <= stop ? (v & kFlagValue) : 0ull;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                    before += part;
                    if (prefixLanes) break;
                    j -= 32;
                }
                if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(tileState + tile) = kFlagPrefix | (before + total);
            }
            if (lane == 0) {
                sBase = before;
                if (tile + 1 == nTiles) st->nLoci = before + total;
            }
        }
        __syncthreads();
        }
        const unsigned long long base = sBase;
#pragma unroll
        for (int k = 0; k < kScoreRounds; ++k) {
            if (!keep[k]) continue;
            const uint32_t hp = round_height(r[k].rp), hm = round_height(r[k].rm);  // heightScoreMap[0.5 + reads], :582 / :589
            const uint32_t bin = (hp | hm) < 4u ? sSmallBin[hp * 4 + hm] : (uint32_t)bin_of(__fmul_rn(freq_of(freq, hp), freq_of(freq, hm)));
            const uint32_t cb = binMap[bin];                                                      // :682
            LocusOut o;
            o.fdr = fdrTable[((size_t)cb * W + ppIdx) * 4 + ((r[k].misc >> 6) & 1u) * 2 + ((r[k].misc >> 5) & 1u)];  // :753
            uint32_t lo = 0, hi = nSlots;  // slot that contains lpos
            while (hi - lo > 1) {
                uint32_t mid = (lo + hi) >> 1;
                if (slots[mid].off <= r[k].lpos) lo = mid; else hi = mid;
            }
            o.contig = slots[lo].contig;
            o.position = r[k].lpos - slots[lo].off + slots[lo].first;
            o.rp = r[k].rp;
            o.rm = r[k].rm;
            loci[base + sCnt[k * (kScoreThreads / 32) + warp] + (keepAll ? (uint32_t)lane : (uint32_t)__popc(ballot[k] & ((1u << lane) - 1u)))] = o;
        }
    }
}

}  // namespace

cudaError_t launch_collapse(const uint32_t* grouped, int W, float* cells, uint16_t* binMap, uint32_t* nBins, uint16_t* groupStart, uint32_t* binOcc,
                            cudaStream_t s, int* launches) {
    const int entries = kBins * W * 4;
    k_collapse_occ<<<(kBins + 7) / 8, 256, 0, s>>>(grouped, W, binOcc);
    k_collapse_map<<<1, 256, 0, s>>>(binOcc, W, binMap, groupStart, nBins);
    k_collapse_cells<<<(entries + 7) / 8, 256, 0, s>>>(grouped, W, groupStart, nBins, cells);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

size_t collapse_occ_words() { return (size_t)kBins * kMaxCellWords; }

cudaError_t launch_fdr_table(const float* cells, const uint32_t* nBins, uint32_t maxBins, int W, int ppIdx, const double* xs, const double* ws, int nSteps,
                             float* fdr, cudaStream_t s, int* launches) {
    if (nSteps > kFdrMaxSteps) return cudaErrorInvalidValue;
    (void)maxBins;  // the kernel reads the real bin count; one warp per cell, the warps of 4 blocks per SM share the cells
    k_fdr_table<<<148 * 4, kFdrWarps * 32, 0, s>>>(cells, nBins, W, ppIdx, xs, ws, nSteps, fdr);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_score(const PairRec* pp, const unsigned long long* totals, uint64_t ppCapacity, uint64_t expected, ScoreState* state,
                         unsigned long long* tileState, FreqTables freq,
                         const float* thresholds, const BinParams* params, const float* fdrTable, const uint16_t* binMap, int W, int ppIdx, uint32_t minHeight,
                         const Slot* slots, uint32_t nSlots, LocusOut* loci, cudaStream_t s, int* launches) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t tiles = (expected + kScoreTile - 1) / kScoreTile;
    unsigned blocks = (unsigned)std::min<uint64_t>(std::max<uint64_t>(tiles, 1), (uint64_t)sms * 6);
    float mh = (float)minHeight;  // `reads >= minStackHeight` compares float with unsigned converted to float (:903)
    k_score<<<blocks, kScoreThreads, 0, s>>>(pp, totals, ppCapacity, state, tileState, freq, thresholds, params, fdrTable, binMap, W, ppIdx, mh, slots, nSlots, loci);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

size_t score_tiles(uint64_t records) { return (size_t)((records + kScoreTile - 1) / kScoreTile) + 1; }

}  // namespace pppk
