// compact.cu -- the last step of the stack builder (ppp_stacks_finish): the dense accumulation arrays become the
// COMPACT stack store the scan reads.
//
// The stack builder (stack_build.cu) accumulates into dense per-position word arrays, because reads arrive in file
// order, batch after batch, file after file (TReadStacksPerGenome of pingpongpro.cpp:103-105 is a std::map that is
// inserted into at random).  Once the input has ended the stacks are what the reference's map holds: a sorted set.  This
// file writes them down as such -- per strand the non-empty words in position order, next to the occupancy bitmap the
// builder maintained -- so that every later pass streams 4 bytes per stack instead of gathering one 32-byte sector per
// stack from arrays that are ~96 % zeros:
//   k_tile_popc      stacks per tile of 32768 positions and strand (reads the bitmaps: G/4 bytes)
//   exclusive scan   -> index of every tile's first stack (tileOff, nTiles + 1 entries per strand)
//   k_compact_write  rank of every stack inside its tile (block scan of popcounts), gather of the words, dense write
// Bound: the sector gathers of k_compact_write (32 B per occupied sector), once per input instead of once per scan.
#include "common.cuh"

namespace pppk {

namespace {

constexpr int kThreads = 256;
constexpr int kTile = kScanTile;
constexpr int kTileWords = kTile / 32;             // occupancy words per tile and strand
constexpr int kWords = kTileWords / kThreads;      // consecutive words owned by one thread of k_compact_write
static_assert(kTile == kThreads * kWords * 32 && kTileWords % 32 == 0, "tile geometry");

__global__ void __launch_bounds__(kThreads) k_tile_popc(const uint32_t* __restrict__ occ, uint64_t G, uint32_t nTiles, uint32_t* __restrict__ cntP,
                                                        uint32_t* __restrict__ cntM) {
    const int lane = threadIdx.x & 31;
    const uint32_t warpsPerGrid = gridDim.x * (kThreads / 32);
    const uint32_t* __restrict__ occM = occ + G / 32;
    for (uint32_t tile = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); tile < nTiles; tile += warpsPerGrid) {
        const size_t w = (size_t)tile * kTileWords;
        uint32_t p = 0, m = 0;
#pragma unroll 8
        for (int k = 0; k < kTileWords / 32; ++k) {
            p += __popc(__ldg(occ + w + k * 32 + lane));
            m += __popc(__ldg(occM + w + k * 32 + lane));
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { p += __shfl_xor_sync(0xffffffffu, p, d); m += __shfl_xor_sync(0xffffffffu, m, d); }
        if (lane == 0) { cntP[tile] = p; cntM[tile] = m; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { cntP[nTiles] = 0; cntM[nTiles] = 0; }  // the scan's last output is the total
}

__global__ void __launch_bounds__(kThreads) k_compact_write(const uint32_t* __restrict__ H, const uint32_t* __restrict__ occ, uint64_t G, uint32_t nTiles,
                                                            const uint32_t* __restrict__ offP, const uint32_t* __restrict__ offM, uint32_t* __restrict__ cmpP,
                                                            uint32_t* __restrict__ cmpM) {
    __shared__ uint32_t sWarp[kThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* __restrict__ occM = occ + G / 32;
    const uint32_t* __restrict__ Hp = H;
    const uint32_t* __restrict__ Hm = H + G;
    for (uint32_t tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        const size_t w = (size_t)tile * kTileWords + (size_t)tid * kWords;  // this thread's kWords consecutive words
        uint32_t nzP[kWords], nzM[kWords], pack = 0;
#pragma unroll
        for (int k = 0; k < kWords; ++k) {
            nzP[k] = __ldg(occ + w + k);
            nzM[k] = __ldg(occM + w + k);
            pack += (uint32_t)__popc(nzP[k]) | ((uint32_t)__popc(nzM[k]) << 16);
        }
        uint32_t inc = pack;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) sWarp[warp] = inc;
        __syncthreads();
        uint32_t before = 0;
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k)
            if (k < warp) before += sWarp[k];
        const uint32_t ex = before + inc - pack;  // plus stacks | minus stacks << 16 of the tile before this thread's words (a tile holds at most 2^15 of each)
        uint32_t* __restrict__ dp = cmpP + offP[tile] + (ex & 0xffffu);
        uint32_t* __restrict__ dm = cmpM + offM[tile] + (ex >> 16);
#pragma unroll
        for (int k = 0; k < kWords; ++k) {
            const uint32_t base = tile * (uint32_t)kTile + ((uint32_t)tid * kWords + k) * 32u;
            // up to four gathers (two per strand) in flight per trip
            uint32_t a = nzP[k], b = nzM[k];
            while (a | b) {
                uint32_t v[4] = {0u, 0u, 0u, 0u};
                if (a) { v[0] = __ldg(Hp + base + (uint32_t)(__ffs(a) - 1)); a &= a - 1; }
                if (b) { v[1] = __ldg(Hm + base + (uint32_t)(__ffs(b) - 1)); b &= b - 1; }
                if (a) { v[2] = __ldg(Hp + base + (uint32_t)(__ffs(a) - 1)); a &= a - 1; }
                if (b) { v[3] = __ldg(Hm + base + (uint32_t)(__ffs(b) - 1)); b &= b - 1; }
                if (v[0]) *dp++ = v[0];
                if (v[2]) *dp++ = v[2];
                if (v[1]) *dm++ = v[1];
                if (v[3]) *dm++ = v[3];
            }
        }
        __syncthreads();  // sWarp is rewritten by the next tile
    }
}

// ppp_reset of a sparse store: a word of H is non-zero only where its occupancy bit is set, so the bitmaps say which
// 32-byte sectors (8 positions = one bitmap byte) have to be zeroed -- instead of a memset over both dense arrays.
__global__ void __launch_bounds__(kThreads) k_sparse_clear(uint32_t* __restrict__ H, uint32_t* __restrict__ occ, size_t occWords) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t w = (size_t)blockIdx.x * kThreads + threadIdx.x; w < occWords; w += stride) {
        const uint32_t m = occ[w];
        if (m == 0u) continue;
        occ[w] = 0u;
        uint4* dst = reinterpret_cast<uint4*>(H + w * 32);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if ((m >> (8 * b)) & 0xffu) { dst[2 * b] = z; dst[2 * b + 1] = z; }
    }
}

}  // namespace

cudaError_t launch_sparse_clear(uint32_t* H, uint32_t* occ, size_t occWords, cudaStream_t s, int* launches) {
    if (occWords == 0) return cudaSuccess;
    k_sparse_clear<<<148 * 16, kThreads, 0, s>>>(H, occ, occWords);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_tile_popc(const uint32_t* occ, uint64_t G, uint32_t nTiles, uint32_t* cntP, uint32_t* cntM, cudaStream_t s, int* launches) {
    unsigned blocks = (nTiles + 7) / 8;
    if (blocks > 148u * 8u) blocks = 148u * 8u;
    if (blocks < 1) blocks = 1;
    k_tile_popc<<<blocks, kThreads, 0, s>>>(occ, G, nTiles, cntP, cntM);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_compact_write(const uint32_t* H, const uint32_t* occ, uint64_t G, uint32_t nTiles, const uint32_t* offP, const uint32_t* offM,
                                 uint32_t* cmpP, uint32_t* cmpM, cudaStream_t s, int* launches) {
    if (nTiles == 0) return cudaSuccess;
    unsigned blocks = nTiles < 148u * 8u ? nTiles : 148u * 8u;
    k_compact_write<<<blocks, kThreads, 0, s>>>(H, occ, G, nTiles, offP, offM, cmpP, cmpM);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace pppk
