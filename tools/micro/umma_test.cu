// Stand-alone check of csrc/umma.cuh on the GPU: C[128][2n] = A[128][K] * B[2n][K]^T as two consecutive tiles of one
// CTA (ring / barrier parities carry over), bf16 inputs, fp32 accumulators, against a CPU fp64 reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/micro/umma_test tools/micro/umma_test.cu
//   tools/micro/umma_test            # all shapes; exit code 0 iff every tile matches
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../smoltts_b200/csrc/tmap_host.h"
#include "umma_legacy.cuh"  // product umma.cuh + the retired tile forms

using namespace smol;

constexpr int NT = 512;

__global__ void __launch_bounds__(NT, 1)
umma_test_kernel(const uint16_t* A, const uint16_t* B, float* C, int K, int n_blk, int m_valid, int variant, int xf, int paired, int tiles,
                 const CUtensorMap* maps) {
    extern __shared__ __align__(1024) unsigned char ring_raw[];
    unsigned char* ring = reinterpret_cast<unsigned char*>(((uintptr_t)ring_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) umma::Bars bars;
    umma::Pipe pipe;
    umma::setup(&bars, pipe, variant == 2 ? (uint32_t)umma::kAcc : 1u, variant == 2 ? 2u : 1u);
    for (int tt = 0; tt < tiles; ++tt) {
        const int t = tt & 1;
        auto row_a = [&](int r) { return r < m_valid ? A + (size_t)r * K : (const uint16_t*)nullptr; };
        auto row_b = [&](int j) { return B + (size_t)(t * n_blk + j) * K; };
        auto xform = [&](int r, int k0, uint4& v) {  // doubles every element (exact in bf16)
            uint16_t* e = reinterpret_cast<uint16_t*>(&v);
            for (int i = 0; i < 8; ++i) {
                float f = __uint_as_float(((uint32_t)e[i]) << 16) * 2.0f;
                e[i] = (uint16_t)(__float_as_uint(f) >> 16);
            }
            (void)r; (void)k0;
        };
        if (variant >= 2) {  // TMA staging (what the decode kernel uses): maps[0] = A (rows past m_valid are outside the map: zero)
            umma::BSrc b0, b1;
            // maps[1]: weight boxes of n_blk rows, maps[2]: of n_blk / 2 rows (the paired halves)
            b0.tm = maps + (paired ? 2 : 1); b0.row0 = t * n_blk; b0.n = paired ? n_blk / 2 : n_blk; b0.box = b0.n;
            b1.tm = maps + 2; b1.row0 = t * n_blk + n_blk / 2; b1.n = paired ? n_blk / 2 : 0; b1.box = n_blk / 2;
            if (variant == 3) umma::tile_mma_tma_exp<NT, false, decltype(xform), 1>(ring, &bars, pipe, K, maps, 0, b0, b1, xform);
            else if (variant == 4) umma::tile_mma_tma_exp<NT, false, decltype(xform), 2>(ring, &bars, pipe, K, maps, 0, b0, b1, xform);
            else if (variant == 5) umma::tile_mma_tma_exp<NT, false, decltype(xform), 4>(ring, &bars, pipe, K, maps, 0, b0, b1, xform);
            else if (variant == 6) umma::tile_mma_tma_exp<NT, false, decltype(xform), 8>(ring, &bars, pipe, K, maps, 0, b0, b1, xform);
            else if (xf) umma::tile_mma_tma_exp<NT, true, decltype(xform), 0>(ring, &bars, pipe, K, maps, 0, b0, b1, xform);
            else umma::tile_mma_tma<NT>(ring, &bars, pipe, K, maps, 0, b0, b1);
        } else if (variant == 1) {  // cp.async staging (kept for the A/B timing)
            if (xf) umma::tile_mma_cpasync<NT, true>(ring, &bars, pipe, K, n_blk, row_a, row_b, xform);
            else umma::tile_mma_cpasync<NT, false>(ring, &bars, pipe, K, n_blk, row_a, row_b, xform);
        } else {
            if (xf) umma::tile_mma<NT, true>(ring, &bars, pipe, K, n_blk, row_a, row_b, xform);
            else umma::tile_mma<NT, false>(ring, &bars, pipe, K, n_blk, row_a, row_b, xform);
        }
        if (paired) {
            umma::tile_epilogue_paired(&bars, n_blk / 2, [&](int row, int c0, const float (&a)[8], const float (&b)[8]) {
                for (int i = 0; i < 8; ++i) {
                    C[(size_t)row * 2 * n_blk + t * n_blk + c0 + i] = a[i];
                    C[(size_t)row * 2 * n_blk + t * n_blk + n_blk / 2 + c0 + i] = b[i];
                }
            }, variant == 2 ? n_blk : 0);
        } else {
            umma::tile_epilogue(&bars, n_blk, [&](int row, int c0, const float (&v)[8]) {
                for (int i = 0; i < 8; ++i) C[(size_t)row * 2 * n_blk + t * n_blk + c0 + i] = v[i];
            }, variant == 2 ? n_blk : 0);
        }
    }
    umma::teardown(&bars);
}

static uint16_t f2bf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7FFF + ((u >> 16) & 1);
    return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static double run(int K, int n, int m_valid, int variant, int xf, int paired, int ctas, int tiles = 2, float* us_per_tile = nullptr) {
    std::vector<uint16_t> hA((size_t)128 * K), hB((size_t)2 * n * K);
    srand(K * 131 + n);
    for (auto& v : hA) v = f2bf((rand() % 2001 - 1000) / 1000.0f);
    for (auto& v : hB) v = f2bf((rand() % 2001 - 1000) / 1000.0f);
    uint16_t *dA, *dB;
    float* dC;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dC, (size_t)128 * 2 * n * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xFF, (size_t)128 * 2 * n * 4);
    const int smem = umma::kRingBytes + 1024;
    cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    CUtensorMap hmaps[3];
    CUtensorMap* dmaps;
    if (!smol::make_tensor_map_2d(&hmaps[0], dA, m_valid, K, K, 128) || !smol::make_tensor_map_2d(&hmaps[1], dB, 2 * n, K, K, n) ||
        !smol::make_tensor_map_2d(&hmaps[2], dB, 2 * n, K, K, n >= 32 ? n / 2 : n)) {
        printf("tensor map encode failed\n");
        exit(4);
    }
    cudaMalloc(&dmaps, sizeof(hmaps));
    cudaMemcpy(dmaps, hmaps, sizeof(hmaps), cudaMemcpyHostToDevice);
    umma_test_kernel<<<ctas, NT, smem>>>(dA, dB, dC, K, n, m_valid, variant, xf, paired, tiles, dmaps);
    cudaError_t e = cudaDeviceSynchronize();
    if (us_per_tile && e == cudaSuccess) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int r = 0; r < 5; ++r) umma_test_kernel<<<ctas, NT, smem>>>(dA, dB, dC, K, n, m_valid, variant, xf, paired, tiles, dmaps);
        cudaEventRecord(e1);
        e = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        *us_per_tile = ms * 1e3f / 5 / tiles;
    }
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(3); }
    std::vector<float> hC((size_t)128 * 2 * n);
    cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int r = 0; r < m_valid; ++r)
        for (int j = 0; j < 2 * n; ++j) {
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += (double)bf2f(hA[(size_t)r * K + k]) * bf2f(hB[(size_t)j * K + k]);
            if (xf) acc *= 2;
            double d = fabs(acc - hC[(size_t)r * 2 * n + j]);
            if (!(d <= worst)) worst = d;  // NaN-safe
        }
    for (int r = m_valid; r < 128; ++r)
        for (int j = 0; j < 2 * n; ++j)
            if (hC[(size_t)r * 2 * n + j] != 0.0f) worst = fmax(worst, 1e9);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dmaps);
    return worst;
}

int main(int argc, char** argv) {
    int bad = 0;
    if (argc > 1 && argv[1][0] == 'a') {  // experiment: A operand staged in tensor memory (tcgen05.cp + TS MMA)
        for (int n : {16, 32, 96}) {
            double w = run(768, n, 128, 5, 0, 0, 1);
            float us = 0;
            run(3072, n, 128, 5, 0, 0, 1, 40, &us);
            printf("A in TMEM: n=%2d max |err| = %.3g (K=768); K=3072: %.2f us per tile (%.0f ns per stage)\n", n, w, us, us * 1e3 / 48);
            run(3072, n, 128, 6, 0, 0, 1, 40, &us);
            printf("M = 64 instruction shape (timing only): n=%2d K=3072: %.2f us per tile (%.0f ns per stage)\n", n, us, us * 1e3 / 48);
            run(3072, n, 128, 2, 0, 0, 1, 40, &us);
            printf("shipped form (four issuing threads, four accumulators): n=%2d K=3072: %.2f us per tile (%.0f ns per stage)\n", n, us, us * 1e3 / 48);
        }
        return 0;
    }
    if (argc > 1) {  // timing of the TMA form only (ring depth experiments: -DUMMA_STAGES=n)
        for (int variant : {2, 3, 4})
            for (int n : {16, 32, 96})
                for (int ctas : {1, 148}) {
                    float us = 0;
                    run(3072, n, 128, variant, 0, 0, ctas, 40, &us);
                    printf("%s stages=%d K=3072 n=%2d ctas=%3d: %.2f us per tile (%.0f ns per stage)\n",
                           variant == 2 ? "TMA+MMA " : variant == 3 ? "MMA only" : "TMA only", umma::kStages, n, ctas, us, us * 1e3 / 48);
                }
        return 0;
    }
    const int Ks[] = {64, 320, 768, 3072};
    const int Ns[] = {16, 32, 64, 96};
    {   // the LBO / SBO convention of umma.cuh (the swapped one addresses shared memory out of range and faults)
        double w = run(768, 32, 128, 0, 0, 0, 1);
        if (w < 1e-2) w = run(768, 32, 128, 1, 0, 0, 1);  // and the cp.async form
        printf("no-swizzle forms: max |err| = %.3g\n", w);
        if (w < 1e-2) w = run(768, 32, 128, 2, 0, 0, 1);  // TMA + 128-byte swizzle
        printf("stride convention: max |err| = %.3g (K=768 n=32)\n", w);
        if (!(w < 1e-2)) { printf("FAIL: stride convention is wrong\n"); return 1; }
    }
    for (int K : Ks)
        for (int n : Ns)
            for (int variant = 0; variant < 3; ++variant) {
                const int m_valid = variant == 1 ? 77 : 128, xf = variant == 1, paired = variant == 2;
                if (paired && n < 32) continue;  // halves are whole 16-row weight boxes
                double w = run(K, n, m_valid, 2, xf, paired, variant == 2 ? 148 : 1);
                if (w < 1e-3 * sqrt((double)K) * (xf ? 2 : 1)) w = run(K, n, m_valid, 0, xf, paired, 1);
                const double tol = 1e-3 * sqrt((double)K) * (xf ? 2 : 1);
                printf("K=%4d n=%3d rows=%3d xform=%d paired=%d: max |err| = %.3g %s\n", K, n, m_valid, xf, paired, w, w < tol ? "ok" : "FAIL");
                if (!(w < tol)) bad = 1;
            }
    // time per tile (all CTAs compute the same tile: the activation rows are shared the way the units of a phase share them)
    for (int variant = 0; variant < 3; ++variant)
        for (int K : {768, 3072})
            for (int ctas : {1, 48, 148})
                for (int xf = 0; xf < 2; ++xf) {
                    float us = 0;
                    run(K, 32, 128, variant, xf, 0, ctas, 40, &us);
                    printf("%s staging K=%4d n=32 xform=%d ctas=%3d: %.2f us per tile (%.0f ns per 64-element stage)\n",
                           variant == 2 ? "TMA sw128" : variant ? "cp.async " : "registers", K, xf, ctas, us, us * 1e3 / (K / 64));
                }
    printf(bad ? "FAIL\n" : "PASS\n");
    return bad;
}
