// score.cu -- K4: collapseBins (pingpongpro.cpp:631-686), calculateFDRs (:695-754) and the selection of
// the reported loci (:901-903), all on the device so that scan -> score needs no host round trip.
//
//   k_collapse    one CTA; the 1000 x (W x 2 x 2) table is staged through shared memory in chunks and one
//                 warp walks the bins sequentially (the merge decision of :670 is inherently serial),
//                 using the same float additions as :658 on counters clipped at 2^24.
//   k_fdr_table   one thread per (collapsed bin, bias, local, overlap): background mean / sigma (:705-718)
//                 and the 1001-step Gaussian integral of :724-734 as the same sequence of IEEE double
//                 operations (__dadd_rn/__dmul_rn/__ddiv_rn, no FMA); x and the weights
//                 0.01/sqrt(2 pi) * exp(-x^2/2) are tabulated by the host with the host libm.
//   k_score_*     per-signature FDR lookup (:750-753) fused with the -s filter (:903) and an
//                 order-preserving compaction (count -> scan -> write) of the overlap-10 loci.
// All of it is latency-bound bookkeeping on kilobytes .. a few megabytes.
#include "common.cuh"

namespace pppk {

namespace {

constexpr int kCollapseThreads = 256;
constexpr int kChunkBins = 100;

__global__ void __launch_bounds__(kCollapseThreads) k_collapse(const uint32_t* __restrict__ grouped, int W, float* __restrict__ cells,
                                                                uint16_t* __restrict__ binMap, uint32_t* __restrict__ nBins) {
    extern __shared__ float sChunk[];  // [kChunkBins][cellsPerBin]
    const int cellsPerBin = W * 4;
    const int perLane = (cellsPerBin + 31) / 32;  // <= 4
    const int tid = threadIdx.x, lane = tid & 31;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t collapsed = 0;
    bool open = false;  // a collapsed bin is being accumulated (its merge loop has not ended)
    for (int chunk = 0; chunk < kBins; chunk += kChunkBins) {
        __syncthreads();
        for (int k = tid; k < kChunkBins * cellsPerBin; k += kCollapseThreads) {
            int b = k / cellsPerBin, c = k - b * cellsPerBin;  // c = overlap * 4 + bias * 2 + local
            uint32_t v = grouped[((size_t)(c >> 2) * kBins + (chunk + b)) * 4 + (c & 3)];
            sChunk[k] = (float)(v < kSaturate ? v : kSaturate);  // the reference's float counter (:605)
        }
        __syncthreads();
        if (tid < 32) {
            for (int b = 0; b < kChunkBins; ++b) {
                if (!open) {  // :644-647
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[q] = 0.f;
                    open = true;
                }
                bool empty = false;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int c = lane + 32 * q;
                    if (q < perLane && c < cellsPerBin) {
                        acc[q] = __fadd_rn(acc[q], sChunk[b * cellsPerBin + c]);  // :658
                        empty |= acc[q] <= 0.f;                                   // :661
                    }
                }
                if (lane == 0) binMap[chunk + b] = (uint16_t)collapsed;            // :666
                bool anyEmpty = __any_sync(0xffffffffu, empty);
                bool last = (chunk + b == kBins - 1);
                if (!anyEmpty || last) {                                           // :670
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        int c = lane + 32 * q;
                        if (q < perLane && c < cellsPerBin) cells[((size_t)(c >> 2) * kBins + collapsed) * 4 + (c & 3)] = acc[q];
                    }
                    ++collapsed;
                    open = false;
                }
            }
        }
    }
    if (tid == 0) *nBins = collapsed;
}

__global__ void __launch_bounds__(128) k_fdr_table(const float* __restrict__ cells, const uint32_t* __restrict__ nBins, int W, int ppIdx,
                                                   const double* __restrict__ xs, const double* __restrict__ ws, int nSteps, float* __restrict__ fdr) {
    const uint32_t C = *nBins;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= C * 4u * (uint32_t)W) return;
    const uint32_t ov = t % (uint32_t)W, cell = t / (uint32_t)W;  // cell = bin * 4 + bias * 2 + local
    const uint32_t bin = cell >> 2, k = cell & 3u;
    auto at = [&](int o) { return cells[((size_t)o * kBins + bin) * 4 + k]; };
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
    double acc = 0.0;
    for (int i = 0; i < nSteps; ++i) {                                       // :724
        double x = xs[i], w = ws[i];
        if (x >= thr) acc = __dadd_rn(acc, w);                               // :728
        else acc = __dadd_rn(acc, __ddiv_rn(__dmul_rn(w, __dadd_rn(__dmul_rn(x, dsd), dmean)), dc));  // :732
    }
    if (acc < 0.0) acc = 0.0; else if (acc > 1.0) acc = 1.0;                 // :736-743 (NaN stays NaN)
    fdr[((size_t)ov * kBins + bin) * 4 + k] = __double2float_rn(acc);        // :745
}

constexpr int kScoreThreads = 256;

__device__ __forceinline__ bool reported(const PairRec& r, int ppIdx, float minHeight) {
    return (int)(r.misc & 31u) == ppIdx && r.rp >= minHeight && r.rm >= minHeight;  // :903
}

__global__ void __launch_bounds__(kScoreThreads) k_score_count(const PairRec* __restrict__ pairs, size_t n, int ppIdx, float minHeight, uint32_t* __restrict__ blockCnt) {
    size_t i = (size_t)blockIdx.x * kScoreThreads + threadIdx.x;
    bool keep = false;
    if (i < n) {
        PairRec r;
        *reinterpret_cast<uint4*>(&r) = *reinterpret_cast<const uint4*>(&pairs[i]);
        keep = reported(r, ppIdx, minHeight);
    }
    int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) blockCnt[blockIdx.x] = (uint32_t)c;
}

__global__ void __launch_bounds__(kScoreThreads) k_score_write(const PairRec* __restrict__ pairs, size_t n, const float* __restrict__ fdrTable,
                                                               const uint16_t* __restrict__ binMap, int ppIdx, float minHeight, const Slot* __restrict__ slots,
                                                               uint32_t nSlots, const uint32_t* __restrict__ blockOff, uint32_t nBlocks, float* __restrict__ pairFdr,
                                                               LocusOut* __restrict__ loci, unsigned long long* nLoci) {
    __shared__ uint32_t sWarp[kScoreThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    size_t i = (size_t)blockIdx.x * kScoreThreads + threadIdx.x;
    bool keep = false;
    PairRec r;
    float fdr = 0.f;
    if (i < n) {
        *reinterpret_cast<uint4*>(&r) = *reinterpret_cast<const uint4*>(&pairs[i]);
        uint32_t ov = r.misc & 31u, bin = binMap[(r.misc >> 7) & 1023u];                      // :682
        fdr = fdrTable[((size_t)ov * kBins + bin) * 4 + ((r.misc >> 6) & 1u) * 2 + ((r.misc >> 5) & 1u)];  // :753
        pairFdr[i] = fdr;
        keep = reported(r, ppIdx, minHeight);
    }
    unsigned ballot = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) sWarp[warp] = __popc(ballot);
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < warp; ++w) before += sWarp[w];
    if (keep) {
        uint32_t dst = blockOff[blockIdx.x] + before + __popc(ballot & ((1u << lane) - 1u));
        uint32_t lo = 0, hi = nSlots;  // slot that contains lpos
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (slots[mid].off <= r.lpos) lo = mid; else hi = mid;
        }
        LocusOut o;
        o.contig = slots[lo].contig;
        o.position = r.lpos - slots[lo].off + slots[lo].first;
        o.fdr = fdr;
        o.rp = r.rp;
        o.rm = r.rm;
        loci[dst] = o;
    }
    if (blockIdx.x == nBlocks - 1 && threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < kScoreThreads / 32; ++w) tot += sWarp[w];
        *nLoci = (unsigned long long)blockOff[blockIdx.x] + tot;
    }
}

}  // namespace

cudaError_t launch_collapse(const uint32_t* grouped, int W, float* cells, uint16_t* binMap, uint32_t* nBins, cudaStream_t s, int* launches) {
    size_t smem = (size_t)kChunkBins * W * 4 * sizeof(float);
    k_collapse<<<1, kCollapseThreads, smem, s>>>(grouped, W, cells, binMap, nBins);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_fdr_table(const float* cells, const uint32_t* nBins, int W, int ppIdx, const double* xs, const double* ws, int nSteps, float* fdr,
                             cudaStream_t s, int* launches) {
    unsigned threads = (unsigned)kBins * 4u * (unsigned)W;  // upper bound; the kernel reads the real bin count
    k_fdr_table<<<(threads + 127) / 128, 128, 0, s>>>(cells, nBins, W, ppIdx, xs, ws, nSteps, fdr);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_score_pairs(const PairRec* pairs, size_t nPairs, const float* fdrTable, const uint16_t* binMap, int ppIdx, uint32_t minHeight,
                               const Slot* slots, uint32_t nSlots, float* pairFdr, LocusOut* loci, unsigned long long* nLoci, uint32_t* blockCnt,
                               uint32_t* scanScratch, size_t scanScratchWords, cudaStream_t s, int* launches) {
    if (nPairs == 0) return cudaMemsetAsync(nLoci, 0, sizeof(unsigned long long), s);
    unsigned blocks = (unsigned)((nPairs + kScoreThreads - 1) / kScoreThreads);
    float mh = (float)minHeight;  // `reads >= minStackHeight` compares float with unsigned converted to float (:903)
    k_score_count<<<blocks, kScoreThreads, 0, s>>>(pairs, nPairs, ppIdx, mh, blockCnt);
    cudaError_t e = exclusive_scan_u32(blockCnt, blockCnt, blocks, scanScratch, scanScratchWords, s, launches);
    if (e != cudaSuccess) return e;
    k_score_write<<<blocks, kScoreThreads, 0, s>>>(pairs, nPairs, fdrTable, binMap, ppIdx, mh, slots, nSlots, blockCnt, blocks, pairFdr, loci, nLoci);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

}  // namespace pppk
