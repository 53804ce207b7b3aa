// Latency / issue rate of legacy mma.sync (HMMA.16816 bf16 -> fp32) and of the IEEE division / square root sequences on
// sm_100a, measured with %clock64 inside one CTA.  Informs ll2_kernel.cu (tensor-core GEMV at batch 1).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_lat mma_lat.cu && ./mma_lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c) :: "memory"); return c; }

template <int NACC>
__global__ void k_mma(long long* out, float* sink, int iters) {
    float acc[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    uint32_t a = 0x3f803f80u + threadIdx.x, b = 0x3f803f80u;
    __syncthreads();
    const long long t0 = clk();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) mma(acc[i], a, a, a, a, b, b);
    }
    const long long t1 = clk();
    float s = 0.f;
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
}
__global__ void k_div(long long* out, float* sink, int iters, float x0) {
    float x = x0 + threadIdx.x;
    __syncthreads();
    const long long t0 = clk();
    for (int it = 0; it < iters; ++it) {
        const float mean = __fdiv_rn(x, 768.0f);
        x = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, 1e-5f))) + 3.0f;
    }
    const long long t1 = clk();
    sink[threadIdx.x] = x;
    if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
}
__global__ void k_lut(long long* out, float* sink, const unsigned short* lut, int iters) {
    uint32_t i = threadIdx.x * 37u;
    __syncthreads();
    const long long t0 = clk();
    for (int it = 0; it < iters; ++it) i = (uint32_t)__ldg(lut + (i & 0xffffu)) + it * 977u;   // dependent global loads
    const long long t1 = clk();
    sink[threadIdx.x] = (float)i;
    if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
}

int main() {
    long long* d_out; float* d_sink; unsigned short* d_lut;
    cudaMalloc(&d_out, 64 * sizeof(long long)); cudaMalloc(&d_sink, 1 << 20); cudaMalloc(&d_lut, 131072); cudaMemset(d_lut, 1, 131072);
    long long h[64];
    const int iters = 1000;
    for (int warps : {1, 4, 7, 8}) {
        k_mma<1><<<1, warps * 32>>>(d_out, d_sink, iters); cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("HMMA.16816 dependent chain, %d warps: %.1f cycles per mma\n", warps, (double)h[0] / iters);
        k_mma<2><<<1, warps * 32>>>(d_out, d_sink, iters); cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("HMMA.16816 2 accumulators,   %d warps: %.1f cycles per mma\n", warps, (double)h[0] / iters / 2);
        k_mma<4><<<1, warps * 32>>>(d_out, d_sink, iters); cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("HMMA.16816 4 accumulators,   %d warps: %.1f cycles per mma\n", warps, (double)h[0] / iters / 4);
    }
    k_div<<<1, 32>>>(d_out, d_sink, iters, 500.f); cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("fdiv_rn + fsqrt_rn + fdiv_rn chain (RMSNorm scale): %.1f cycles\n", (double)h[0] / iters);
    k_lut<<<1, 32>>>(d_out, d_sink, d_lut, iters); cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("dependent __ldg (L1/L2 hit): %.1f cycles per load\n", (double)h[0] / iters);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
