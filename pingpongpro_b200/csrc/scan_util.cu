// scan_util.cu -- device-wide exclusive prefix sum over uint32 (three small launches, no host sync).
// Used for: ordered destinations of the per-tile signature runs, radix-sort digit offsets, and the
// order-preserving compaction of the reported loci.
#include "common.cuh"

namespace pppk {

namespace {
constexpr int kPrefixThreads = 1024;
constexpr int kPrefixItems = 4;
constexpr int kPrefixTile = kPrefixThreads * kPrefixItems;

__device__ __forceinline__ uint32_t warp_inclusive(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Exclusive scan of one value per thread across the block; returns the exclusive prefix, *total = block sum.
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t* total, uint32_t* sh /*[33]*/) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = warp_inclusive(v, lane);
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0;
        uint32_t wi = warp_inclusive(w, lane);
        sh[lane] = wi - w;
        if (lane == 31) sh[32] = wi;
    }
    __syncthreads();
    uint32_t r = sh[warp] + inc - v;
    *total = sh[32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kPrefixThreads) k_prefix_reduce(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ blockSums) {
    __shared__ uint32_t sh[33];
    size_t base = (size_t)blockIdx.x * kPrefixTile + (size_t)threadIdx.x * kPrefixItems;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kPrefixItems; ++k)
        if (base + k < n) s += in[base + k];
    uint32_t total;
    block_exclusive(s, &total, sh);
    if (threadIdx.x == 0) blockSums[blockIdx.x] = total;
}

// In-place exclusive scan of a short array by one block (loops with a carry).
__global__ void __launch_bounds__(kPrefixThreads) k_prefix_single(uint32_t* data, size_t n) {
    __shared__ uint32_t sh[33];
    uint32_t carry = 0;
    for (size_t base = 0; base < n; base += kPrefixThreads) {
        size_t i = base + threadIdx.x;
        uint32_t v = i < n ? data[i] : 0;
        uint32_t total;
        uint32_t ex = block_exclusive(v, &total, sh);
        if (i < n) data[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(kPrefixThreads) k_prefix_apply(const uint32_t* in, uint32_t* out, size_t n, const uint32_t* __restrict__ blockSums) {
    __shared__ uint32_t sh[33];
    size_t base = (size_t)blockIdx.x * kPrefixTile + (size_t)threadIdx.x * kPrefixItems;
    uint32_t v[kPrefixItems];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kPrefixItems; ++k) {
        v[k] = base + k < n ? in[base + k] : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t ex = block_exclusive(s, &total, sh) + blockSums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kPrefixItems; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
}
}  // namespace

size_t exclusive_scan_scratch_words(size_t n) { return (n + kPrefixTile - 1) / kPrefixTile + 1; }

// out[i] = sum(in[0..i)).  in == out is allowed.  scratch: exclusive_scan_scratch_words(n) words.
cudaError_t exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* scratch, size_t scratchWords, cudaStream_t s, int* launches) {
    if (n == 0) return cudaSuccess;
    size_t nBlocks = (n + kPrefixTile - 1) / kPrefixTile;
    if (scratchWords < nBlocks) return cudaErrorInvalidValue;
    k_prefix_reduce<<<(unsigned)nBlocks, kPrefixThreads, 0, s>>>(in, n, scratch);
    k_prefix_single<<<1, kPrefixThreads, 0, s>>>(scratch, nBlocks);
    k_prefix_apply<<<(unsigned)nBlocks, kPrefixThreads, 0, s>>>(in, out, n, scratch);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

}  // namespace pppk
