// Micro-benchmark: does it matter HOW the consumers poll?  Flat all-to-all (every CTA publishes its words, every CTA
// polls all words).  mode 0: ld.relaxed.gpu.v4 (two words per load);  mode 1: atom.or.b64 with 0 (executes at the home
// L2 slice, one word per atomic);  mode 2: ld.relaxed.sys.v4;  mode 3: ld.volatile.v4;  mode 4: ld.global.cv.v4.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void st_relaxed_v2(void* p, uint32_t a, uint32_t b) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
template <int MODE>
__device__ __forceinline__ uint4 poll_pair(const unsigned long long* p) {
    uint4 v;
    if (MODE == 0) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    else if (MODE == 2) asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    else if (MODE == 3) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    else if (MODE == 4) asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    else {
        unsigned long long a, b;
        asm volatile("atom.relaxed.gpu.global.or.b64 %0, [%1], 0;" : "=l"(a) : "l"(p) : "memory");
        asm volatile("atom.relaxed.gpu.global.or.b64 %0, [%1], 0;" : "=l"(b) : "l"(p + 1) : "memory");
        v.x = (uint32_t)a; v.y = (uint32_t)(a >> 32); v.z = (uint32_t)b; v.w = (uint32_t)(b >> 32);
    }
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(384, 1)
alltoall(unsigned long long* buf, int n_words, int R, int iters, int work_cycles, int delay_ns, long long* cycles) {
    const int cta = blockIdx.x, n = gridDim.x, tid = threadIdx.x;
    const int w0 = (int)((long long)n_words * cta / n), w1 = (int)((long long)n_words * (cta + 1) / n);
    const int rep = cta % R;
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 1; it <= iters; ++it) {
        unsigned long long* region = buf + (size_t)(it & 3) * n_words * R;
        if (work_cycles > 0) { long long s = clock64(); while (clock64() - s < work_cycles) {} }
        for (int i = tid; i < (w1 - w0) * R; i += blockDim.x) {
            const int w = w0 + i / R, r = i % R;
            st_relaxed_v2(region + (size_t)r * n_words + w, (uint32_t)w, (uint32_t)it);
        }
        if (delay_ns) __nanosleep(delay_ns);
        const unsigned long long* src = region + (size_t)rep * n_words;
        for (int i = tid; i < n_words / 2; i += blockDim.x) {
            uint4 v;
            do { v = poll_pair<MODE>(src + 2 * i); } while (v.y != (uint32_t)it || v.w != (uint32_t)it);
            acc += v.x + v.z;
        }
        __syncthreads();
    }
    if (cta == 0 && tid == 0) *cycles = clock64() - t0;
    if (acc == 0xdeadbeef) buf[0] = acc;
}

template <int MODE>
static void run(const char* name, unsigned long long* buf, long long* cyc, int n_sms, double mhz) {
    const int iters = 2000;
    for (int w : {384, 1536}) for (int R : {1, 8}) for (int delay : {0, 256}) {
        int work = 600;
        cudaMemset(buf, 0, 64 << 20);
        void* args[] = {&buf, (void*)&w, (void*)&R, (void*)&iters, (void*)&work, (void*)&delay, &cyc};
        cudaError_t e = cudaLaunchCooperativeKernel((void*)alltoall<MODE>, dim3(n_sms), dim3(384), args, 0, 0);
        cudaError_t e2 = cudaDeviceSynchronize();
        if (e != cudaSuccess || e2 != cudaSuccess) { printf("launch failed\n"); return; }
        printf("%-16s words=%5d rep=%d delay=%3d : %7.0f cycles/phase (%.3f us)\n", name, w, R, delay, (double)*cyc / iters, (double)*cyc / iters / mhz);
    }
}

int main() {
    cudaSetDevice(0);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    unsigned long long* buf;
    long long* cyc;
    cudaMalloc(&buf, 64 << 20);
    cudaMallocManaged(&cyc, 8);
    run<0>("ld.relaxed.gpu", buf, cyc, prop.multiProcessorCount, clk_khz / 1e3);
    run<1>("atom.or 0", buf, cyc, prop.multiProcessorCount, clk_khz / 1e3);
    run<2>("ld.relaxed.sys", buf, cyc, prop.multiProcessorCount, clk_khz / 1e3);
    run<3>("ld.volatile", buf, cyc, prop.multiProcessorCount, clk_khz / 1e3);
    run<4>("ld.cv", buf, cyc, prop.multiProcessorCount, clk_khz / 1e3);
    return 0;
}
