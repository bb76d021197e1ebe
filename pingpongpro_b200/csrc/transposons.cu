// transposons.cu -- K5: findSuppressedTransposons (pingpongpro.cpp:1195-1274); q-values (:1276-1312) on the host.
//
// The reference walks, per contig and overlap, one std::list of signatures with a persistent iterator while
// it visits the contig's transposons in (start, end) order, adding (r+ + r-) * (1 - fdr) of every signature
// inside [start, end] in position order (float, :1227).  Here:
//   1. the signatures (held in (position, overlap) order) are stably partitioned by overlap, which yields the
//      reference's list order (overlap, contig, position);
//   2. one thread per (transposon, overlap) finds its signature range by binary search;
//   3. one thread per (contig, overlap) replays the iterator's only history dependence: after a transposon has
//      consumed a list to its end, the reference dereferences list.end() (:1215, :1222), which with libstdc++
//      reads the list's SIZE as a position; a later transposon whose start is >= that size finds nothing
//      (SURVEY App. D #4).  That flag is reproduced so that overlapping annotation gives what the reference gives;
//   4. one thread per (transposon, overlap) folds its range sequentially with __fadd_rn (order matters in fp32);
//   5. one thread per transposon computes mean / sigma / z (:1244-1266) and the literal p-value loop (:1268-1269)
//      in fp64 (device exp; the result agrees with the host libm to ~1e-16 relative, tolerance 1e-9).
// Latency-bound: kilobytes to megabytes.
#include "transposons.cuh"

#include <algorithm>
#include <cmath>

namespace pppk {

namespace {

struct TpDev {
    uint32_t start, end;    // contig coordinates, end inclusive
    uint32_t off, first;    // slot of the contig
    uint32_t cLo, cHi;      // lpos range of the contig's plus positions (cLo == cHi: contig not held)
};

constexpr int kPartThreads = 256;

__global__ void __launch_bounds__(kPartThreads) k_ov_count(const PairRec* __restrict__ pairs, size_t n, uint32_t* __restrict__ blockHist, uint32_t nBlocks) {
    __shared__ uint32_t h[32];
    if (threadIdx.x < 32) h[threadIdx.x] = 0;
    __syncthreads();
    size_t i = (size_t)blockIdx.x * kPartThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[pairs[i].misc & 31u], 1u);
    __syncthreads();
    if (threadIdx.x < 32) blockHist[(size_t)threadIdx.x * nBlocks + blockIdx.x] = h[threadIdx.x];
}

// Stable scatter by overlap index; also converts every signature to what the fold needs.
__global__ void __launch_bounds__(kPartThreads) k_ov_scatter(const PairRec* __restrict__ pairs, const float* __restrict__ fdr, size_t n,
                                                              const uint32_t* __restrict__ offsets, uint32_t nBlocks, uint32_t* __restrict__ sPos,
                                                              float* __restrict__ sScore, float* __restrict__ sRp, float* __restrict__ sRm) {
    __shared__ uint32_t cnt[kPartThreads / 32][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    cnt[warp][lane] = 0;
    __syncthreads();
    size_t i = (size_t)blockIdx.x * kPartThreads + threadIdx.x;
    bool ok = i < n;
    PairRec r = {0, 0.f, 0.f, 0};
    if (ok) *reinterpret_cast<uint4*>(&r) = *reinterpret_cast<const uint4*>(&pairs[i]);
    uint32_t o = r.misc & 31u;
    unsigned act = __ballot_sync(0xffffffffu, ok);
    unsigned peers = 0;
    if (ok) {
        peers = __match_any_sync(act, o);
        if (lane == __ffs(peers) - 1) cnt[warp][o] = __popc(peers);
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // first slot of (warp, overlap) inside the block's output ranges
        uint32_t run = offsets[(size_t)threadIdx.x * nBlocks + blockIdx.x];
        for (int w = 0; w < kPartThreads / 32; ++w) {
            uint32_t c = cnt[w][threadIdx.x];
            cnt[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
    if (ok) {
        uint32_t dst = cnt[warp][o] + __popc(peers & ((1u << lane) - 1u));
        sPos[dst] = r.lpos;
        sScore[dst] = __fmul_rn(__fadd_rn(r.rp, r.rm), __fsub_rn(1.0f, fdr[i]));  // (r+ + r-) * (1 - fdr), :1227
        sRp[dst] = r.rp;
        sRm[dst] = r.rm;
    }
}

__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* a, uint32_t lo, uint32_t hi, uint64_t key) {  // first i with a[i] >= key
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if ((uint64_t)a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// bounds[(t*W+o)*4 + {0,1,2,3}] = list begin L0, list end L1, range begin lo, range end hi
__global__ void k_tp_bounds(const TpDev* __restrict__ tp, uint32_t nTp, int W, const uint32_t* __restrict__ offsets, uint32_t nBlocks, uint32_t nPairs,
                            const uint32_t* __restrict__ sPos, uint32_t* __restrict__ bounds) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nTp * (uint32_t)W) return;
    uint32_t t = g / (uint32_t)W, o = g % (uint32_t)W;
    TpDev d = tp[t];
    uint32_t segLo = offsets[(size_t)o * nBlocks];
    uint32_t segHi = (o + 1 < 32) ? offsets[(size_t)(o + 1) * nBlocks] : nPairs;
    uint32_t L0 = lower_bound_u32(sPos, segLo, segHi, d.cLo), L1 = lower_bound_u32(sPos, segLo, segHi, d.cHi);
    // position P of a signature = lpos - off + first;  P >= start  <=>  lpos >= start - first + off (when start >= first)
    uint32_t lo = d.start >= d.first ? lower_bound_u32(sPos, L0, L1, (uint64_t)d.start - d.first + d.off) : L0;
    uint32_t hi = d.end >= d.first ? lower_bound_u32(sPos, lo, L1, (uint64_t)d.end - d.first + d.off + 1) : lo;
    uint32_t* b = bounds + (size_t)g * 4;
    b[0] = L0; b[1] = L1; b[2] = lo; b[3] = hi;
}

// The iterator's history: one thread per (contig group, overlap), sequential over the group's transposons.
__global__ void k_tp_skip(const TpDev* __restrict__ tp, const uint32_t* __restrict__ groupStart, uint32_t nGroups, int W, const uint32_t* __restrict__ bounds,
                          uint8_t* __restrict__ skip) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nGroups * (uint32_t)W) return;
    uint32_t grp = g / (uint32_t)W, o = g % (uint32_t)W;
    bool atEnd = false;  // iterator == list.end() with a non-empty list
    for (uint32_t t = groupStart[grp]; t < groupStart[grp + 1]; ++t) {
        const uint32_t* b = bounds + ((size_t)t * W + o) * 4;
        uint32_t len = b[1] - b[0];
        // :1215: while (it != begin && it->position > start) --it;  at end() "position" is the list size
        bool s = atEnd && len > 0 && !(len > tp[t].start);
        skip[(size_t)t * W + o] = s ? 1 : 0;
        atEnd = s ? true : (len > 0 && b[3] == b[1]);
    }
}

__global__ void k_tp_sums(uint32_t nTp, int W, int ppIdx, const uint32_t* __restrict__ bounds, const uint8_t* __restrict__ skip, const float* __restrict__ sScore,
                          const float* __restrict__ sRp, const float* __restrict__ sRm, float* __restrict__ hist, float* __restrict__ rPlus,
                          float* __restrict__ rMinus) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nTp * (uint32_t)W) return;
    uint32_t t = g / (uint32_t)W, o = g % (uint32_t)W;
    const uint32_t* b = bounds + (size_t)g * 4;
    float sum = 0.f, rp = 0.f, rm = 0.f;
    if (!skip[g]) {
        for (uint32_t i = b[2]; i < b[3]; ++i) {
            sum = __fadd_rn(sum, sScore[i]);                                      // :1227
            if ((int)o == ppIdx) { rp = __fadd_rn(rp, sRp[i]); rm = __fadd_rn(rm, sRm[i]); }  // :1230-1234
        }
    }
    hist[(size_t)t * 32 + o] = sum;                                               // :1240
    if ((int)o == ppIdx) { rPlus[t] = rp; rMinus[t] = rm; }
}

// Range sharding: the signatures of one (transposon, overlap) may be spread over several contexts (position ranges in
// ascending order).  The sums are sequential single-precision additions in position order, so the contexts fold one
// after the other, each continuing from the values the previous one left in hist / rPlus / rMinus (zero for the first).
__global__ void k_tp_sums_chain(uint32_t nTp, int W, int ppIdx, const uint32_t* __restrict__ bounds, const uint8_t* __restrict__ skip, const float* __restrict__ sScore,
                                const float* __restrict__ sRp, const float* __restrict__ sRm, float* __restrict__ hist, float* __restrict__ rPlus,
                                float* __restrict__ rMinus) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nTp * (uint32_t)W) return;
    uint32_t t = g / (uint32_t)W, o = g % (uint32_t)W;
    const uint32_t* b = bounds + (size_t)g * 4;
    if (skip[g] || b[2] >= b[3]) return;  // nothing of this list lives here: the running values pass through
    const bool pp = (int)o == ppIdx;
    float sum = hist[(size_t)t * 32 + o], rp = 0.f, rm = 0.f;
    if (pp) { rp = rPlus[t]; rm = rMinus[t]; }
    for (uint32_t i = b[2]; i < b[3]; ++i) {
        sum = __fadd_rn(sum, sScore[i]);                                          // :1227
        if (pp) { rp = __fadd_rn(rp, sRp[i]); rm = __fadd_rn(rm, sRm[i]); }       // :1230-1234
    }
    hist[(size_t)t * 32 + o] = sum;                                               // :1240
    if (pp) { rPlus[t] = rp; rMinus[t] = rm; }
}

__global__ void k_tp_stats(uint32_t nTp, int W, int ppIdx, double coeff, const float* __restrict__ hist, double* __restrict__ zOut, double* __restrict__ pOut) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nTp) return;
    const float* h = hist + (size_t)t * 32;
    float mean = 0.f;
    for (int o = 0; o < W; ++o)
        if (o != ppIdx) mean = __fadd_rn(mean, h[o]);                             // :1244-1247
    mean = __fdiv_rn(mean, (float)(W - 1));                                       // :1248
    if (mean == 0.f && h[ppIdx] == 0.f) {                                         // :1250-1253
        zOut[t] = __longlong_as_double(0x7ff8000000000000ll);
        pOut[t] = 1.0;
        return;
    }
    float sd = 0.f;
    for (int o = 0; o < W; ++o)
        if (o != ppIdx) {
            double d = (double)__fsub_rn(h[o], mean);
            sd = __double2float_rn(__dadd_rn((double)sd, __dmul_rn(d, d)));       // :1260
        }
    sd = __double2float_rn(sqrt(__dmul_rn(__ddiv_rn(1.0, (double)(W - 2)), (double)sd)));  // :1261
    if ((double)sd <= 1E-10) sd = (float)1E-10;                                   // :1262-1263
    double z = (double)__fdiv_rn(__fsub_rn(h[ppIdx], mean), sd);                  // :1266
    double p = 0.0;
    // |z| >= 1e13: x += 0.01 stops advancing and the reference never returns; every term is exp(-huge) = 0 anyway.
    // z > 40 (or z < -45): all terms underflow to exactly 0.
    if (fabs(z) < 1e13 && z <= 40.0 && z >= -45.0) {
        const double upper = __dadd_rn(z, 5.0);
        for (double x = z; x <= upper; x = __dadd_rn(x, 0.01))                    // :1268
            p = __dadd_rn(p, __dmul_rn(coeff, exp(__dmul_rn(__dmul_rn(-0.5, x), x))));  // :1269
    }
    zOut[t] = z;
    pOut[t] = p;
}

struct DevBuf {
    void* p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    ~DevBuf() { if (p) cudaFree(p); }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

}  // namespace

#define TCK(expr)                          \
    do {                                   \
        cudaError_t e__ = (expr);          \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

namespace {

// reference order: contig (std::map key), then start, end (:206-212), ties in input order (stable list::sort, :1174);
// groupStart = first index of every contig group (+ n)
void sortIntervals(const ppp_interval* intervals, size_t n, std::vector<uint32_t>& perm, std::vector<uint32_t>& groupStart) {
    perm.resize(n);
    for (size_t i = 0; i < n; ++i) perm[i] = (uint32_t)i;
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) {
        const ppp_interval &x = intervals[a], &y = intervals[b];
        if (x.contig != y.contig) return x.contig < y.contig;
        if (x.start != y.start) return x.start < y.start;
        return x.end < y.end;
    });
    groupStart.clear();
    for (size_t k = 0; k < n; ++k)
        if (k == 0 || intervals[perm[k - 1]].contig != intervals[perm[k]].contig) groupStart.push_back((uint32_t)k);
    groupStart.push_back((uint32_t)n);
}

// The sorted intervals in the coordinates of one context's slots.
void localIntervals(const std::vector<Slot>& slots, const ppp_interval* intervals, const std::vector<uint32_t>& perm, std::vector<TpDev>& tp) {
    const size_t n = perm.size();
    tp.resize(n);
    for (size_t k = 0; k < n; ++k) {
        const ppp_interval& iv = intervals[perm[k]];
        TpDev d;
        d.start = iv.start;
        d.end = iv.end;
        d.off = d.first = d.cLo = d.cHi = 0;
        if (iv.contig < slots.size()) {  // contigs unknown to the alignment header have no signatures (:1101-1106)
            const Slot& s = slots[iv.contig];
            if (s.lenMinus) { d.off = s.off; d.first = s.first; d.cLo = s.off; d.cHi = s.off + s.lenPlus; }
        }
        tp[k] = d;
    }
}

// Host tail: copies the sums and statistics into results[] (sorted order) and computes the q-values (:1276-1312).
void finishResults(size_t n, int ppIdx, const std::vector<float>& hist, const std::vector<float>& rp, const std::vector<float>& rm, const std::vector<double>& z,
                   const std::vector<double>& p, const std::vector<uint32_t>& perm, ppp_tp_result* results, uint32_t* order) {
    for (size_t k = 0; k < n; ++k) {
        ppp_tp_result& r = results[k];
        for (int o = 0; o < 32; ++o) r.histogram[o] = hist[k * 32 + o];
        r.reads_plus = rp[k];
        r.reads_minus = rm[k];
        r.z = z[k];
        r.p = p[k];
        r.p_value = (float)p[k];  // :1271
        r.q_value = 1.0f;         // :187
        if (order) order[k] = perm[k];
    }
    // q-values, :1276-1312: stable sort by p ascending (:1281), walked from the largest p with i = 1, 2, ...
    std::vector<uint32_t> byP(n);
    for (size_t k = 0; k < n; ++k) byP[k] = (uint32_t)k;
    std::stable_sort(byP.begin(), byP.end(), [&](uint32_t a, uint32_t b) { return results[a].p_value < results[b].p_value; });
    unsigned i = 1;
    float prevP = 1, prevQ = 1;
    for (size_t k = n; k-- > 0;) {
        ppp_tp_result& r = results[byP[k]];
        if (r.histogram[ppIdx] / (r.reads_plus + r.reads_minus) < 0.2) r.q_value = 1;  // :1289-1291
        else if (r.p_value == prevP) r.q_value = prevQ;                                // :1293-1298
        else {
            r.q_value = r.p_value * (float)n / (float)i;                               // :1302
            if (r.q_value > 1) r.q_value = 1;
            prevP = r.p_value;
            prevQ = r.q_value;
        }
        i++;
    }
}

}  // namespace

cudaError_t run_transposons(const TransposonJob& job, const ppp_interval* intervals, size_t n, ppp_tp_result* results, uint32_t* order, std::string& err) {
    if (n == 0) return cudaSuccess;
    if (n >= 0x7fffffffull / 32) { err = "too many intervals"; return cudaSuccess; }
    const int W = job.W;
    std::vector<uint32_t> perm, groupStart;
    sortIntervals(intervals, n, perm, groupStart);
    std::vector<TpDev> tp;
    localIntervals(*job.slots, intervals, perm, tp);
    const uint32_t nGroups = (uint32_t)groupStart.size() - 1;
    const uint32_t nPairs = (uint32_t)job.nPairs;
    const uint32_t nBlocks = std::max<uint32_t>(1, (nPairs + kPartThreads - 1) / kPartThreads);
    const size_t histWords = (size_t)32 * nBlocks;

    DevBuf bHist, bScratch, bPos, bScore, bRp, bRm, bTp, bGroups, bBounds, bSkip, bOutHist, bOutRp, bOutRm, bZ, bP;
    TCK(bHist.alloc(histWords * 4));
    size_t scratchWords = exclusive_scan_scratch_words(histWords);
    TCK(bScratch.alloc(scratchWords * 4));
    TCK(bPos.alloc((size_t)nPairs * 4));
    TCK(bScore.alloc((size_t)nPairs * 4));
    TCK(bRp.alloc((size_t)nPairs * 4));
    TCK(bRm.alloc((size_t)nPairs * 4));
    TCK(bTp.alloc(n * sizeof(TpDev)));
    TCK(bGroups.alloc(groupStart.size() * 4));
    TCK(bBounds.alloc(n * W * 4 * sizeof(uint32_t)));
    TCK(bSkip.alloc(n * W));
    TCK(bOutHist.alloc(n * 32 * sizeof(float)));
    TCK(bOutRp.alloc(n * sizeof(float)));
    TCK(bOutRm.alloc(n * sizeof(float)));
    TCK(bZ.alloc(n * sizeof(double)));
    TCK(bP.alloc(n * sizeof(double)));
    cudaStream_t s = job.stream;
    TCK(cudaMemcpyAsync(bTp.p, tp.data(), n * sizeof(TpDev), cudaMemcpyHostToDevice, s));
    TCK(cudaMemcpyAsync(bGroups.p, groupStart.data(), groupStart.size() * 4, cudaMemcpyHostToDevice, s));
    TCK(cudaMemsetAsync(bHist.p, 0, histWords * 4, s));
    TCK(cudaMemsetAsync(bOutHist.p, 0, n * 32 * sizeof(float), s));

    if (nPairs) {
        k_ov_count<<<nBlocks, kPartThreads, 0, s>>>(job.pairs, nPairs, bHist.as<uint32_t>(), nBlocks);
        *job.launches += 1;
    }
    TCK(exclusive_scan_u32(bHist.as<uint32_t>(), bHist.as<uint32_t>(), histWords, bScratch.as<uint32_t>(), scratchWords, s, job.launches));
    if (nPairs) {
        k_ov_scatter<<<nBlocks, kPartThreads, 0, s>>>(job.pairs, job.pairFdr, nPairs, bHist.as<uint32_t>(), nBlocks, bPos.as<uint32_t>(), bScore.as<float>(),
                                                      bRp.as<float>(), bRm.as<float>());
        *job.launches += 1;
    }
    const uint32_t nTp = (uint32_t)n;
    const unsigned th = 128;
    k_tp_bounds<<<(nTp * W + th - 1) / th, th, 0, s>>>(bTp.as<TpDev>(), nTp, W, bHist.as<uint32_t>(), nBlocks, nPairs, bPos.as<uint32_t>(), bBounds.as<uint32_t>());
    k_tp_skip<<<(nGroups * W + th - 1) / th, th, 0, s>>>(bTp.as<TpDev>(), bGroups.as<uint32_t>(), nGroups, W, bBounds.as<uint32_t>(), bSkip.as<uint8_t>());
    k_tp_sums<<<(nTp * W + th - 1) / th, th, 0, s>>>(nTp, W, job.ppIdx, bBounds.as<uint32_t>(), bSkip.as<uint8_t>(), bScore.as<float>(), bRp.as<float>(),
                                                     bRm.as<float>(), bOutHist.as<float>(), bOutRp.as<float>(), bOutRm.as<float>());
    const double pi = 3.14159265358979323846;  // :50
    const double coeff = 0.01 * 1 / sqrt(2 * pi);
    k_tp_stats<<<(nTp + th - 1) / th, th, 0, s>>>(nTp, W, job.ppIdx, coeff, bOutHist.as<float>(), bZ.as<double>(), bP.as<double>());
    *job.launches += 4;
    TCK(cudaGetLastError());

    std::vector<float> hist(n * 32), rp(n), rm(n);
    std::vector<double> z(n), p(n);
    TCK(cudaMemcpyAsync(hist.data(), bOutHist.p, n * 32 * sizeof(float), cudaMemcpyDeviceToHost, s));
    TCK(cudaMemcpyAsync(rp.data(), bOutRp.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    TCK(cudaMemcpyAsync(rm.data(), bOutRm.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    TCK(cudaMemcpyAsync(z.data(), bZ.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    TCK(cudaMemcpyAsync(p.data(), bP.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    TCK(cudaStreamSynchronize(s));
    finishResults(n, job.ppIdx, hist, rp, rm, z, p, perm, results, order);
    return cudaSuccess;
}

namespace {

// What one context keeps on its device between the phases of the sharded run.
struct ShardState {
    DevBuf bHist, bScratch, bPos, bScore, bRp, bRm, bTp, bBounds, bSkip, bOutHist, bOutRp, bOutRm;
    std::vector<uint32_t> bounds;  // host copy of k_tp_bounds' output
};

}  // namespace

// The same over the contexts of a range-sharded genome (jobs[] in ascending range order, one host thread, contexts on
// one or several devices).  A transposon that straddles a range boundary -- or lies in another range altogether -- is
// handled exactly like in the single-context run:
//   1. every context partitions ITS signatures by overlap and finds, per (transposon, overlap), its local list and range;
//   2. the iterator history of step 3 above depends only on the length of the whole list (sum of the local lengths) and
//      on whether a transposon consumed it to the end (true iff it did so in every context): those two facts are
//      combined on the host and replayed by k_tp_skip on the first context's device;
//   3. the sums are chained: context after context continues the sequential single-precision additions from the
//      values of its predecessor (k_tp_sums_chain), so the additions happen in global position order;
//   4. mean / sigma / z / p on the last context's device, q-values on the host.
cudaError_t run_transposons_sharded(const TransposonJob* jobs, const int* devices, int nShards, const ppp_interval* intervals, size_t n, ppp_tp_result* results,
                                    uint32_t* order, std::string& err) {
    if (nShards <= 0) { err = "no contexts"; return cudaSuccess; }
    if (n == 0) return cudaSuccess;
    if (n >= 0x7fffffffull / 32) { err = "too many intervals"; return cudaSuccess; }
    const int W = jobs[0].W, ppIdx = jobs[0].ppIdx;
    for (int k = 1; k < nShards; ++k)
        if (jobs[k].W != W || jobs[k].ppIdx != ppIdx) { err = "contexts differ in their overlap range"; return cudaSuccess; }
    std::vector<uint32_t> perm, groupStart;
    sortIntervals(intervals, n, perm, groupStart);
    const uint32_t nGroups = (uint32_t)groupStart.size() - 1;
    const uint32_t nTp = (uint32_t)n;
    const size_t nCells = n * (size_t)W;
    const unsigned th = 128;
    std::vector<ShardState> st(nShards);
    std::vector<TpDev> tp;

    // 1. local partitions and bounds
    for (int k = 0; k < nShards; ++k) {
        const TransposonJob& job = jobs[k];
        ShardState& S = st[k];
        TCK(cudaSetDevice(devices[k]));
        localIntervals(*job.slots, intervals, perm, tp);
        const uint32_t nPairs = (uint32_t)job.nPairs;
        const uint32_t nBlocks = std::max<uint32_t>(1, (nPairs + kPartThreads - 1) / kPartThreads);
        const size_t histWords = (size_t)32 * nBlocks;
        const size_t scratchWords = exclusive_scan_scratch_words(histWords);
        TCK(S.bHist.alloc(histWords * 4));
        TCK(S.bScratch.alloc(scratchWords * 4));
        TCK(S.bPos.alloc((size_t)nPairs * 4));
        TCK(S.bScore.alloc((size_t)nPairs * 4));
        TCK(S.bRp.alloc((size_t)nPairs * 4));
        TCK(S.bRm.alloc((size_t)nPairs * 4));
        TCK(S.bTp.alloc(n * sizeof(TpDev)));
        TCK(S.bBounds.alloc(nCells * 4 * sizeof(uint32_t)));
        TCK(S.bSkip.alloc(nCells));
        TCK(S.bOutHist.alloc(n * 32 * sizeof(float)));
        TCK(S.bOutRp.alloc(n * sizeof(float)));
        TCK(S.bOutRm.alloc(n * sizeof(float)));
        cudaStream_t s = job.stream;
        TCK(cudaMemcpyAsync(S.bTp.p, tp.data(), n * sizeof(TpDev), cudaMemcpyHostToDevice, s));
        TCK(cudaMemsetAsync(S.bHist.p, 0, histWords * 4, s));
        if (nPairs) {
            k_ov_count<<<nBlocks, kPartThreads, 0, s>>>(job.pairs, nPairs, S.bHist.as<uint32_t>(), nBlocks);
            *job.launches += 1;
        }
        TCK(exclusive_scan_u32(S.bHist.as<uint32_t>(), S.bHist.as<uint32_t>(), histWords, S.bScratch.as<uint32_t>(), scratchWords, s, job.launches));
        if (nPairs) {
            k_ov_scatter<<<nBlocks, kPartThreads, 0, s>>>(job.pairs, job.pairFdr, nPairs, S.bHist.as<uint32_t>(), nBlocks, S.bPos.as<uint32_t>(), S.bScore.as<float>(),
                                                          S.bRp.as<float>(), S.bRm.as<float>());
            *job.launches += 1;
        }
        k_tp_bounds<<<(nTp * W + th - 1) / th, th, 0, s>>>(S.bTp.as<TpDev>(), nTp, W, S.bHist.as<uint32_t>(), nBlocks, nPairs, S.bPos.as<uint32_t>(), S.bBounds.as<uint32_t>());
        *job.launches += 1;
        TCK(cudaGetLastError());
        S.bounds.resize(nCells * 4);
        TCK(cudaMemcpyAsync(S.bounds.data(), S.bBounds.p, nCells * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));  // (pageable: the copy is synchronous for the host)
        TCK(cudaStreamSynchronize(s));  // tp is refilled for the next context
    }

    // 2. whole-list facts -> iterator history, replayed on the first context's device
    std::vector<uint32_t> whole(nCells * 4);
    for (size_t g = 0; g < nCells; ++g) {
        uint64_t len = 0;
        bool consumed = true;
        for (int k = 0; k < nShards; ++k) {
            const uint32_t* b = st[k].bounds.data() + g * 4;
            len += b[1] - b[0];
            consumed = consumed && b[3] == b[1];
        }
        if (len > 0xfffffffeull) { err = "more than 2^32 signatures in one list"; return cudaSuccess; }
        uint32_t* w = whole.data() + g * 4;
        w[0] = 0;
        w[1] = (uint32_t)len;
        w[2] = 0;
        w[3] = consumed ? (uint32_t)len : 0xffffffffu;  // k_tp_skip only asks whether [3] == [1]
    }
    std::vector<uint8_t> skip(nCells);
    {
        TCK(cudaSetDevice(devices[0]));
        cudaStream_t s = jobs[0].stream;
        DevBuf bWhole, bGroups;
        TCK(bWhole.alloc(nCells * 4 * sizeof(uint32_t)));
        TCK(bGroups.alloc(groupStart.size() * 4));
        TCK(cudaMemcpyAsync(bWhole.p, whole.data(), nCells * 4 * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        TCK(cudaMemcpyAsync(bGroups.p, groupStart.data(), groupStart.size() * 4, cudaMemcpyHostToDevice, s));
        k_tp_skip<<<(nGroups * W + th - 1) / th, th, 0, s>>>(st[0].bTp.as<TpDev>(), bGroups.as<uint32_t>(), nGroups, W, bWhole.as<uint32_t>(), st[0].bSkip.as<uint8_t>());
        *jobs[0].launches += 1;
        TCK(cudaGetLastError());
        TCK(cudaMemcpyAsync(skip.data(), st[0].bSkip.p, nCells, cudaMemcpyDeviceToHost, s));
        TCK(cudaStreamSynchronize(s));
    }

    // 3. chained sums (running values travel through the host; kilobytes to a few megabytes)
    std::vector<float> hist(n * 32, 0.f), rp(n, 0.f), rm(n, 0.f);
    for (int k = 0; k < nShards; ++k) {
        ShardState& S = st[k];
        TCK(cudaSetDevice(devices[k]));
        cudaStream_t s = jobs[k].stream;
        if (k > 0) TCK(cudaMemcpyAsync(S.bSkip.p, skip.data(), nCells, cudaMemcpyHostToDevice, s));
        TCK(cudaMemcpyAsync(S.bOutHist.p, hist.data(), n * 32 * sizeof(float), cudaMemcpyHostToDevice, s));
        TCK(cudaMemcpyAsync(S.bOutRp.p, rp.data(), n * sizeof(float), cudaMemcpyHostToDevice, s));
        TCK(cudaMemcpyAsync(S.bOutRm.p, rm.data(), n * sizeof(float), cudaMemcpyHostToDevice, s));
        k_tp_sums_chain<<<(nTp * W + th - 1) / th, th, 0, s>>>(nTp, W, ppIdx, S.bBounds.as<uint32_t>(), S.bSkip.as<uint8_t>(), S.bScore.as<float>(), S.bRp.as<float>(),
                                                               S.bRm.as<float>(), S.bOutHist.as<float>(), S.bOutRp.as<float>(), S.bOutRm.as<float>());
        *jobs[k].launches += 1;
        TCK(cudaGetLastError());
        if (k + 1 < nShards) {
            TCK(cudaMemcpyAsync(hist.data(), S.bOutHist.p, n * 32 * sizeof(float), cudaMemcpyDeviceToHost, s));
            TCK(cudaMemcpyAsync(rp.data(), S.bOutRp.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
            TCK(cudaMemcpyAsync(rm.data(), S.bOutRm.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
            TCK(cudaStreamSynchronize(s));
        }
    }

    // 4. statistics where the final sums are
    {
        ShardState& S = st[nShards - 1];
        TCK(cudaSetDevice(devices[nShards - 1]));
        cudaStream_t s = jobs[nShards - 1].stream;
        DevBuf bZ, bP;
        TCK(bZ.alloc(n * sizeof(double)));
        TCK(bP.alloc(n * sizeof(double)));
        const double pi = 3.14159265358979323846;  // :50
        const double coeff = 0.01 * 1 / sqrt(2 * pi);
        k_tp_stats<<<(nTp + th - 1) / th, th, 0, s>>>(nTp, W, ppIdx, coeff, S.bOutHist.as<float>(), bZ.as<double>(), bP.as<double>());
        *jobs[nShards - 1].launches += 1;
        TCK(cudaGetLastError());
        std::vector<double> z(n), p(n);
        TCK(cudaMemcpyAsync(hist.data(), S.bOutHist.p, n * 32 * sizeof(float), cudaMemcpyDeviceToHost, s));
        TCK(cudaMemcpyAsync(rp.data(), S.bOutRp.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
        TCK(cudaMemcpyAsync(rm.data(), S.bOutRm.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
        TCK(cudaMemcpyAsync(z.data(), bZ.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
        TCK(cudaMemcpyAsync(p.data(), bP.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
        TCK(cudaStreamSynchronize(s));
        finishResults(n, ppIdx, hist, rp, rm, z, p, perm, results, order);
    }
    return cudaSuccess;
}

namespace {

__global__ void k_tp_list_facts(uint32_t nCells, const uint32_t* __restrict__ bounds, uint32_t* __restrict__ facts) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nCells) return;
    const uint32_t* b = bounds + (size_t)g * 4;
    facts[2 * g] = b[1] - b[0];                 // signatures of this (transposon, overlap)'s list held here
    facts[2 * g + 1] = b[3] == b[1] ? 1u : 0u;  // the transposon's range reaches the end of the local list
}

// bounds of the WHOLE list for k_tp_skip: [0, total length, 0, consumed on every rank ? length : none]
__global__ void k_tp_whole_bounds(uint32_t nCells, const uint32_t* __restrict__ facts, uint32_t world, uint32_t* __restrict__ whole) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nCells) return;
    uint32_t* w = whole + (size_t)g * 4;
    w[0] = 0;
    w[1] = facts[2 * g];
    w[2] = 0;
    w[3] = facts[2 * g + 1] == world ? facts[2 * g] : 0xffffffffu;  // k_tp_skip only asks whether [3] == [1]
}

}  // namespace

cudaError_t run_transposons_collective(const TransposonJob& job, int rank, int world, const std::function<int(void*, size_t)>& reduce, const ppp_interval* intervals,
                                       size_t n, ppp_tp_result* results, uint32_t* order, std::string& err) {
    if (n == 0) return cudaSuccess;
    if (n >= 0x7fffffffull / 64) { err = "too many intervals"; return cudaSuccess; }
    const int W = job.W;
    std::vector<uint32_t> perm, groupStart;
    sortIntervals(intervals, n, perm, groupStart);
    std::vector<TpDev> tp;
    localIntervals(*job.slots, intervals, perm, tp);
    const uint32_t nGroups = (uint32_t)groupStart.size() - 1;
    const uint32_t nTp = (uint32_t)n;
    const uint32_t nCells = nTp * (uint32_t)W;
    const uint32_t nPairs = (uint32_t)job.nPairs;
    const uint32_t nBlocks = std::max<uint32_t>(1, (nPairs + kPartThreads - 1) / kPartThreads);
    const size_t histWords = (size_t)32 * nBlocks;
    const size_t scratchWords = exclusive_scan_scratch_words(histWords);
    const unsigned th = 128;
    cudaStream_t s = job.stream;

    DevBuf bHist, bScratch, bPos, bScore, bRp, bRm, bTp, bGroups, bBounds, bWhole, bFacts, bSkip, bRun, bZ, bP;
    TCK(bHist.alloc(histWords * 4));
    TCK(bScratch.alloc(scratchWords * 4));
    TCK(bPos.alloc((size_t)nPairs * 4));
    TCK(bScore.alloc((size_t)nPairs * 4));
    TCK(bRp.alloc((size_t)nPairs * 4));
    TCK(bRm.alloc((size_t)nPairs * 4));
    TCK(bTp.alloc(n * sizeof(TpDev)));
    TCK(bGroups.alloc(groupStart.size() * 4));
    TCK(bBounds.alloc((size_t)nCells * 4 * sizeof(uint32_t)));
    TCK(bWhole.alloc((size_t)nCells * 4 * sizeof(uint32_t)));
    TCK(bFacts.alloc((size_t)nCells * 2 * sizeof(uint32_t)));
    TCK(bSkip.alloc(nCells));
    const size_t runWords = n * 34;  // running sums: hist[n][32], R+[n], R-[n] (floats, exchanged as 32-bit patterns)
    TCK(bRun.alloc(runWords * sizeof(float)));
    TCK(bZ.alloc(n * sizeof(double)));
    TCK(bP.alloc(n * sizeof(double)));
    float* runHist = bRun.as<float>();
    float* runRp = runHist + n * 32;
    float* runRm = runRp + n;

    // 1. this rank's partition by overlap and its local bounds per (transposon, overlap)
    TCK(cudaMemcpyAsync(bTp.p, tp.data(), n * sizeof(TpDev), cudaMemcpyHostToDevice, s));
    TCK(cudaMemcpyAsync(bGroups.p, groupStart.data(), groupStart.size() * 4, cudaMemcpyHostToDevice, s));
    TCK(cudaMemsetAsync(bHist.p, 0, histWords * 4, s));
    if (nPairs) {
        k_ov_count<<<nBlocks, kPartThreads, 0, s>>>(job.pairs, nPairs, bHist.as<uint32_t>(), nBlocks);
        *job.launches += 1;
    }
    TCK(exclusive_scan_u32(bHist.as<uint32_t>(), bHist.as<uint32_t>(), histWords, bScratch.as<uint32_t>(), scratchWords, s, job.launches));
    if (nPairs) {
        k_ov_scatter<<<nBlocks, kPartThreads, 0, s>>>(job.pairs, job.pairFdr, nPairs, bHist.as<uint32_t>(), nBlocks, bPos.as<uint32_t>(), bScore.as<float>(),
                                                      bRp.as<float>(), bRm.as<float>());
        *job.launches += 1;
    }
    k_tp_bounds<<<(nCells + th - 1) / th, th, 0, s>>>(bTp.as<TpDev>(), nTp, W, bHist.as<uint32_t>(), nBlocks, nPairs, bPos.as<uint32_t>(), bBounds.as<uint32_t>());
    // 2. whole-list facts (summed over the ranks) -> the iterator history, replayed identically on every rank
    k_tp_list_facts<<<(nCells + th - 1) / th, th, 0, s>>>(nCells, bBounds.as<uint32_t>(), bFacts.as<uint32_t>());
    *job.launches += 2;
    TCK(cudaGetLastError());
    if (reduce(bFacts.p, (size_t)nCells * 2) != 0) { err = "all-reduce failed"; return cudaSuccess; }
    k_tp_whole_bounds<<<(nCells + th - 1) / th, th, 0, s>>>(nCells, bFacts.as<uint32_t>(), (uint32_t)world, bWhole.as<uint32_t>());
    k_tp_skip<<<(nGroups * W + th - 1) / th, th, 0, s>>>(bTp.as<TpDev>(), bGroups.as<uint32_t>(), nGroups, W, bWhole.as<uint32_t>(), bSkip.as<uint8_t>());
    *job.launches += 2;
    TCK(cudaGetLastError());
    // 3. the sums, rank after rank in position order
    TCK(cudaMemsetAsync(bRun.p, 0, runWords * sizeof(float), s));
    for (int k = 0; k < world; ++k) {
        if (k == rank) {
            k_tp_sums_chain<<<(nCells + th - 1) / th, th, 0, s>>>(nTp, W, job.ppIdx, bBounds.as<uint32_t>(), bSkip.as<uint8_t>(), bScore.as<float>(), bRp.as<float>(),
                                                                  bRm.as<float>(), runHist, runRp, runRm);
            *job.launches += 1;
            TCK(cudaGetLastError());
        } else {
            TCK(cudaMemsetAsync(bRun.p, 0, runWords * sizeof(float), s));  // (+0.0f is the all-zero pattern)
        }
        if (reduce(bRun.p, runWords) != 0) { err = "all-reduce failed"; return cudaSuccess; }
    }
    // 4. statistics, replicated
    const double pi = 3.14159265358979323846;  // :50
    const double coeff = 0.01 * 1 / sqrt(2 * pi);
    k_tp_stats<<<(nTp + th - 1) / th, th, 0, s>>>(nTp, W, job.ppIdx, coeff, runHist, bZ.as<double>(), bP.as<double>());
    *job.launches += 1;
    TCK(cudaGetLastError());
    std::vector<float> hist(n * 32), rp(n), rm(n);
    std::vector<double> z(n), p(n);
    TCK(cudaMemcpyAsync(hist.data(), runHist, n * 32 * sizeof(float), cudaMemcpyDeviceToHost, s));
    TCK(cudaMemcpyAsync(rp.data(), runRp, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    TCK(cudaMemcpyAsync(rm.data(), runRm, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    TCK(cudaMemcpyAsync(z.data(), bZ.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    TCK(cudaMemcpyAsync(p.data(), bP.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    TCK(cudaStreamSynchronize(s));
    finishResults(n, job.ppIdx, hist, rp, rm, z, p, perm, results, order);
    return cudaSuccess;
}

}  // namespace pppk
