// Micro-benchmark: hierarchical all-to-all.  Clusters of C CTAs; every CTA publishes its words to L2 (LL words,
// R replicas), polls only 1/C of the vector from L2 and forwards what it received to all C members of its cluster
// through distributed shared memory (8-byte st.shared::cluster of {payload, epoch}); consumers spin on their OWN shared
// memory.  L2 poll traffic drops by C.  Compare with ll_handoff.cu (flat: every CTA polls everything).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint4 ld_relaxed_v4(const void* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_v2(void* p, uint32_t a, uint32_t b) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_volatile_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

template <int C>
__global__ void __launch_bounds__(384, 1)
alltoall_cluster(unsigned long long* buf, int n_words, int R, int iters, int work_cycles, int delay_ns, long long* cycles) {
    extern __shared__ __align__(16) unsigned long long sbuf[];  // [2][n_words] words {payload, epoch}
    cg::cluster_group cluster = cg::this_cluster();
    const int cta = blockIdx.x, n = gridDim.x, tid = threadIdx.x;
    const int crank = (int)cluster.block_rank(), cid = cta / C;
    const int w0 = (int)((long long)n_words * cta / n), w1 = (int)((long long)n_words * (cta + 1) / n);
    const int rep = cid % R;
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(sbuf);
    for (int i = tid; i < 2 * n_words; i += blockDim.x) sbuf[i] = 0ull;
    cluster.sync();
    long long t0 = clock64();
    uint32_t acc = 0;
    // my share of the poll: pairs of words [p0, p1)
    const int pairs = n_words / 2;
    const int p0 = pairs * crank / C, p1 = pairs * (crank + 1) / C;
    for (int it = 1; it <= iters; ++it) {
        unsigned long long* region = buf + (size_t)(it & 3) * n_words * R;
        const uint32_t sb = s0 + (uint32_t)(it & 1) * n_words * 8u;
        if (work_cycles > 0) { long long s = clock64(); while (clock64() - s < work_cycles) {} }
        for (int i = tid; i < (w1 - w0) * R; i += blockDim.x) {
            const int w = w0 + i / R, r = i % R;
            st_relaxed_v2(region + (size_t)r * n_words + w, (uint32_t)w, (uint32_t)it);
        }
        if (delay_ns) __nanosleep(delay_ns);
        // poll my share from L2 and forward it to every member of the cluster
        const unsigned long long* src = region + (size_t)rep * n_words;
        for (int i = p0 + tid; i < p1; i += blockDim.x) {
            uint4 v;
            do { v = ld_relaxed_v4(src + 2 * i); } while (v.y != (uint32_t)it || v.w != (uint32_t)it);
#pragma unroll
            for (int r = 0; r < C; ++r) st_cluster_v4(mapa(sb + (uint32_t)i * 16u, r), v);
        }
        // wait until the whole vector sits in my shared memory
        for (int i = tid; i < pairs; i += blockDim.x) {
            uint4 v;
            do { v = ld_shared_volatile_v4(sb + (uint32_t)i * 16u); } while (v.y != (uint32_t)it || v.w != (uint32_t)it);
            acc += v.x + v.z;
        }
        __syncthreads();
    }
    if (cta == 0 && tid == 0) *cycles = clock64() - t0;
    if (acc == 0xdeadbeef) buf[0] = acc;
    cluster.sync();
}

template <int C>
static void run(unsigned long long* buf, long long* cyc, int n_sms, double mhz) {
    const int iters = 2000;
    int max_clusters = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = 2 * 1536 * 8;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    if (C > 8) cudaFuncSetAttribute(alltoall_cluster<C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cfg.gridDim = dim3(C);
    cudaOccupancyMaxActiveClusters(&max_clusters, alltoall_cluster<C>, &cfg);
    int n_ctas = max_clusters * C;
    if (n_ctas > n_sms) n_ctas = n_sms / C * C;
    printf("cluster size %d: %d clusters co-resident -> %d CTAs\n", C, max_clusters, n_ctas);
    if (n_ctas == 0) return;
    cfg.gridDim = dim3(n_ctas);
    for (int work : {0, 600}) for (int w : {384, 1536}) for (int R : {1, 4}) for (int delay : {0, 256}) {
        cudaMemset(buf, 0, 64 << 20);
        cudaError_t e = cudaLaunchKernelEx(&cfg, alltoall_cluster<C>, buf, w, R, iters, work, delay, cyc);
        cudaError_t e2 = cudaDeviceSynchronize();
        if (e != cudaSuccess || e2 != cudaSuccess) { printf("  launch failed: %s %s\n", cudaGetErrorString(e), cudaGetErrorString(e2)); return; }
        printf("  C=%d work=%4d words=%5d rep=%d delay=%3d : %7.0f cycles/phase (%.3f us)\n", C, work, w, R, delay, (double)*cyc / iters,
               (double)*cyc / iters / mhz);
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
    run<1>(buf, cyc, prop.multiProcessorCount, clk_khz / 1e3);
    run<2>(buf, cyc, prop.multiProcessorCount, clk_khz / 1e3);
    run<4>(buf, cyc, prop.multiProcessorCount, clk_khz / 1e3);
    run<8>(buf, cyc, prop.multiProcessorCount, clk_khz / 1e3);
    run<16>(buf, cyc, prop.multiProcessorCount, clk_khz / 1e3);
    return 0;
}
