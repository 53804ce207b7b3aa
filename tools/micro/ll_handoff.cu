// Micro-benchmark of the LL hand-off: what does an all-to-all exchange of a small vector through L2 cost
// on this GPU, as a function of vector size, replicas and polling threads?  (Build: nvcc -arch=sm_100a.)
//   mode 0: ping-pong between CTA 0 and CTA 1 (one word each way)           -> one-way latency
//   mode 1: every CTA publishes W words (R replicas), every CTA polls all N*W words with T threads,
//           then a __syncthreads; repeated ITERS times                       -> per-phase floor
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint2 ld_relaxed_v2(const void* p) {
    uint2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const void* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_v2(void* p, uint32_t a, uint32_t b) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}

__global__ void pingpong(unsigned long long* buf, int iters, long long* cycles) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    if (me > 1) return;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (me == 0) {
            st_relaxed_v2(buf, 7u, (uint32_t)i);
            while (ld_relaxed_v2(buf + 64).y != (uint32_t)i) {}
        } else {
            while (ld_relaxed_v2(buf).y != (uint32_t)i) {}
            st_relaxed_v2(buf + 64, 9u, (uint32_t)i);
        }
    }
    if (me == 0) *cycles = clock64() - t0;
}

// vector of n_words words per phase, replicated R times; words dealt to CTAs in contiguous blocks
__global__ void __launch_bounds__(512, 1)
alltoall(unsigned long long* buf, int n_words, int R, int T, int iters, int work_cycles, long long* cycles) {
    const int cta = blockIdx.x, n = gridDim.x, tid = threadIdx.x;
    const int w0 = (int)((long long)n_words * cta / n), w1 = (int)((long long)n_words * (cta + 1) / n);
    const int rep = cta % R;
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 1; it <= iters; ++it) {
        unsigned long long* region = buf + (size_t)(it & 3) * n_words * R;  // 4 rotating regions
        // "compute"
        if (work_cycles > 0) { long long s = clock64(); while (clock64() - s < work_cycles) {} }
        // publish my words: one thread per (word, replica)
        for (int i = tid; i < (w1 - w0) * R; i += blockDim.x) {
            const int w = w0 + i / R, r = i % R;
            st_relaxed_v2(region + (size_t)r * n_words + w, (uint32_t)w, (uint32_t)it);
        }
        // poll all words (pairs of words per thread), T polling threads
        const unsigned long long* src = region + (size_t)rep * n_words;
        for (int i = tid; i < n_words / 2 && tid < T; i += T) {
            uint4 v;
            do { v = ld_relaxed_v4(src + 2 * i); } while (v.y != (uint32_t)it || v.w != (uint32_t)it);
            acc += v.x + v.z;
        }
        __syncthreads();
    }
    if (cta == 0 && tid == 0) *cycles = clock64() - t0;
    if (acc == 0xdeadbeef) buf[0] = acc;
}

int main(int argc, char** argv) {
    int dev = 0;
    cudaSetDevice(dev);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    const int n_sms = prop.multiProcessorCount;
    unsigned long long* buf;
    long long* cyc;
    cudaMalloc(&buf, 64 << 20);
    cudaMemset(buf, 0, 64 << 20);
    cudaMallocManaged(&cyc, 8);
    const int iters = 2000;
    pingpong<<<2, 32>>>(buf, iters, cyc);
    cudaDeviceSynchronize();
    printf("SMs %d, clock %.0f MHz\n", n_sms, clk_khz / 1e3);
    printf("ping-pong: %.0f cycles per round trip -> %.0f cycles (%.3f us) one way\n", (double)*cyc / iters, (double)*cyc / iters / 2,
           (double)*cyc / iters / 2 / (clk_khz / 1e3));
    const int words[] = {384, 1536};
    const int reps[] = {1, 2, 4, 8};
    const int threads[] = {96, 192, 512};
    for (int work : {0, 600}) {
        for (int w : words) {
            for (int R : reps) {
                for (int T : threads) {
                    if (T * 2 > w * 2 && T > w / 2) continue;
                    cudaMemset(buf, 0, 64 << 20);
                    void* args[] = {&buf, (void*)&w, (void*)&R, (void*)&T, (void*)&iters, (void*)&work, &cyc};
                    cudaError_t e = cudaLaunchCooperativeKernel((void*)alltoall, dim3(n_sms), dim3(512), args, 0, 0);
                    cudaDeviceSynchronize();
                    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("launch failed\n"); return 1; }
                    printf("all-to-all work=%4d cyc words=%5d rep=%d poll_threads=%3d : %7.0f cycles/phase (%.3f us)\n", work, w, R, T,
                           (double)*cyc / iters, (double)*cyc / iters / (clk_khz / 1e3));
                }
            }
        }
    }
    return 0;
}
