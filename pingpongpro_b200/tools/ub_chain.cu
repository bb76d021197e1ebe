// Microbenchmark behind the long-run fold of stack_build.cu: cycles per dependent __fadd_rn when the addend comes from a
// register, from broadcast 128-bit shared-memory loads (one or two chains) or from warp shuffles.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ub_chain ub_chain.cu
// B200: 4.04 / 4.47 / 5.08 / 6.61 cycles per addition (one warp), 4.05 / 4.71 / 6.65 / 9.33 (eight warps).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_chain(float* out, const float* in, long long* cyc, int mode) {
    __shared__ __align__(16) float st[256];
    st[threadIdx.x % 256] = in[threadIdx.x % 256];
    __syncthreads();
    float h = in[300];
    float g = in[301];
    long long t0 = clock64();
    if (mode == 0) {
        float c = in[0];
#pragma unroll 1
        for (int r = 0; r < 64; ++r) {
#pragma unroll
            for (int i = 0; i < 256; ++i) h = __fadd_rn(h, c);
        }
    } else if (mode == 1) {
        const float4* q = reinterpret_cast<const float4*>(st);
#pragma unroll 1
        for (int r = 0; r < 64; ++r) {
#pragma unroll
            for (int t = 0; t < 64; ++t) { float4 v = q[t]; h = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(h, v.x), v.y), v.z), v.w); }
        }
    } else if (mode == 2) {  // two independent chains
        const float4* q = reinterpret_cast<const float4*>(st);
#pragma unroll 1
        for (int r = 0; r < 64; ++r) {
#pragma unroll
            for (int t = 0; t < 64; ++t) { float4 v = q[t]; h = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(h, v.x), v.y), v.z), v.w); g = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(g, v.w), v.z), v.y), v.x); }
        }
    } else if (mode == 3) {  // shuffle-fed
        float w = in[threadIdx.x & 31];
#pragma unroll 1
        for (int r = 0; r < 512; ++r) {
#pragma unroll
            for (int l = 0; l < 32; ++l) h = __fadd_rn(h, __shfl_sync(0xffffffffu, w, l));
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = h + g; cyc[0] = t1 - t0; }
}
int main() {
    float *in, *out; long long* cyc;
    cudaMalloc(&in, 4096); cudaMalloc(&out, 16); cudaMalloc(&cyc, 8);
    float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = 1.0f / (1 + i % 7);
    cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
    for (int threads : {32, 256}) for (int mode = 0; mode < 4; ++mode) {
        long long c = 0;
        for (int rep = 0; rep < 2; ++rep) { k_chain<<<1, threads>>>(out, in, cyc, mode); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); }
        printf("threads %d mode %d: %.2f cycles per addition\n", threads, mode, (double)c / 16384.0);
    }
    return 0;
}
