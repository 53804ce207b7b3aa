// Measured-and-retired predecessors of the shipped tcgen05 tile (smoltts_b200/csrc/umma.cuh: tile_mma_tma), kept for the
// A/B micro-benchmark tools/micro/umma_test.cu (profiles/r1c_umma_microbench.txt).  NOT part of the product library.
//   * tile_mma_cpasync / tile_mma: operands staged with per-thread 16-byte copies (cp.async, or registers) into the
//     canonical NO-swizzle layout of the UMMA descriptor,
//         byte offset of element (row r, k) inside a stage = (k / 8) * LBO + (r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2
//     with an optional in-place transform, fence.proxy.async and a block barrier per stage: 0.8-1.2 us per stage.
//   * tile_mma_tma_exp: the TMA form with its experiment switches (SKIP: 1 = no copies, 2 = no MMAs, 4 = A operand through
//     tensor memory, 8 = one issuing thread; XFORM = rewrite the activation pieces in shared memory before the MMA).
#pragma once

#include "../../smoltts_b200/csrc/umma.cuh"

namespace smol {
namespace umma {

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Copies of ring chunk `kc` of the tile into stage `s`: A = 128 rows x 64 k, B = n_blk rows x 64 k.
// Piece p of a warp's 32: 4 consecutive 16-byte pieces of 8 consecutive rows (64 contiguous global bytes per
// row; the 8 rows of one K chunk are 128 contiguous shared-memory bytes: conflict-free stores).
template <int NT, class RowA, class RowB>
__device__ __forceinline__ void issue_chunk(unsigned char* ring, int s, int kc, int n_blk, RowA row_a, RowB row_b) {
    const uint32_t a0 = s2u(ring + (size_t)s * kStageBytes), b0 = a0 + kStageA;
    const int tid = threadIdx.x;
#pragma unroll
    for (int j = 0; j < (kM * 8 + NT - 1) / NT; ++j) {
        const int p = tid + NT * j;
        if (p < kM * 8) {
            const int lane = p & 31, w = p >> 5;
            const int c8 = ((w & 1) << 2) | (lane & 3), row = ((w >> 1) << 3) | (lane >> 2);
            const uint16_t* src = row_a(row);
            const uint32_t dst = a0 + c8 * (kM * 16) + row * 16;
            // rows past the batch: zero fill (src-size 0; the address only has to be a valid global one)
            if (src) cp_async16(dst, src + kc * kBK + c8 * 8, 16u);
            else cp_async16(dst, row_b(0), 0u);
        }
    }
#pragma unroll
    for (int j = 0; j < (kMaxN * 8 + NT - 1) / NT; ++j) {
        const int p = tid + NT * j;
        if (p < n_blk * 8) {
            const int lane = p & 31, w = p >> 5;
            const int c8 = ((w & 1) << 2) | (lane & 3), row = ((w >> 1) << 3) | (lane >> 2);
            const uint16_t* src = row_b(row);
            const uint32_t dst = b0 + c8 * (n_blk * 16) + row * 16;
            cp_async16(dst, src + kc * kBK + c8 * 8, 16u);
        }
    }
}

// acc[128][n_blk] (TMEM) = A[128][K] * B[n_blk][K]^T.  All NT threads of the CTA call this (uniform arguments).
// row_a(r) -> pointer to the K bf16 of tile row r (nullptr: zero row); row_b(j) -> weight row j;
// xform(r, k0, uint4&) rewrites the 8 elements [k0, k0 + 8) of row r in place before the MMA sees them (or no-op).
// On return the accumulator is complete and visible to tcgen05.ld of every thread.
template <int NT, bool XFORM, class RowA, class RowB, class Xform>
__device__ __forceinline__ void tile_mma_cpasync(unsigned char* ring, Bars* bars, Pipe& pipe, int K, int n_blk, RowA row_a, RowB row_b,
                                                 Xform xform, bool swap_strides = false) {
    const int nk = K / kBK;
    const int tid = threadIdx.x;
    const uint32_t idesc = instr_desc(n_blk);
    const uint32_t tmem = bars->tmem_base;
    const uint32_t g0 = pipe.chunk;
    auto acquire_and_issue = [&](int kc) {
        if (kc < nk) {
            const uint32_t g = g0 + (uint32_t)kc;
            const int s = (int)(g % kStages);
            if (g >= (uint32_t)kStages) bar_wait(&bars->free_[s], (g / kStages - 1u) & 1u);
            issue_chunk<NT>(ring, s, kc, n_blk, row_a, row_b);
        }
        cp_async_commit();
    };
    for (int kc = 0; kc < kAhead; ++kc) acquire_and_issue(kc);
    for (int kc = 0; kc < nk; ++kc) {
        const uint32_t g = g0 + (uint32_t)kc;
        const int s = (int)(g % kStages);
        unsigned char* stage = ring + (size_t)s * kStageBytes;
        cp_async_wait<kAhead - 1>();
        if (XFORM) {
#pragma unroll
            for (int j = 0; j < (kM * 8 + NT - 1) / NT; ++j) {
                const int p = tid + NT * j;
                if (p < kM * 8) {
                    const int lane = p & 31, w = p >> 5;
                    const int c8 = ((w & 1) << 2) | (lane & 3), row = ((w >> 1) << 3) | (lane >> 2);
                    uint4* q = reinterpret_cast<uint4*>(stage + c8 * (kM * 16) + row * 16);
                    uint4 v = *q;
                    xform(row, kc * kBK + c8 * 8, v);
                    *q = v;
                }
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            fence_after_sync();
            const uint32_t a0 = s2u(stage), b0 = a0 + kStageA;
            const uint32_t lbo_a = kM * 16, lbo_b = (uint32_t)n_blk * 16, sbo = 128;
#pragma unroll
            for (int j = 0; j < kBK / 16; ++j) {
                const uint64_t ad = swap_strides ? smem_desc(a0 + 2 * j * lbo_a, sbo, lbo_a) : smem_desc(a0 + 2 * j * lbo_a, lbo_a, sbo);
                const uint64_t bd = swap_strides ? smem_desc(b0 + 2 * j * lbo_b, sbo, lbo_b) : smem_desc(b0 + 2 * j * lbo_b, lbo_b, sbo);
                mma_bf16(tmem, ad, bd, idesc, (kc | j) ? 1u : 0u);
            }
            commit(&bars->free_[s]);
            if (kc == nk - 1) commit(&bars->done);
        }
        __syncwarp();
        acquire_and_issue(kc + kAhead);
    }
    cp_async_wait<0>();
    pipe.chunk = g0 + (uint32_t)nk;
    bar_wait(&bars->done, pipe.tile & 1u);
    pipe.tile += 1;
    fence_after_sync();
}

// Register-staged form of the same tile (the one the decode kernel uses): every thread keeps its pieces of the next
// kPre stages in registers (ld.global.cg issued kPre stages ahead), applies the transform there, stores them to the
// stage with st.shared and fences.  Unlike the cp.async form no asynchronous shared-memory write is in flight when
// fence.proxy.async executes -- measured on B200: the fence waits for the thread's outstanding cp.async groups, which
// serialises that pipeline at one L2 round trip per stage.
constexpr int kPre = 2;
template <int NT, bool XFORM, class RowA, class RowB, class Xform>
__device__ __forceinline__ void tile_mma(unsigned char* ring, Bars* bars, Pipe& pipe, int K, int n_blk, RowA row_a, RowB row_b,
                                         Xform xform) {
    constexpr int PA = (kM * 8 + NT - 1) / NT, PB = (kMaxN * 8 + NT - 1) / NT;
    const int nk = K / kBK;
    const int tid = threadIdx.x;
    const uint32_t idesc = instr_desc(n_blk);
    const uint32_t tmem = bars->tmem_base;
    const uint32_t g0 = pipe.chunk;
    // piece i of this thread: rows / K chunks are a function of (tid, i) only, so one pointer per piece is all the state
    const uint16_t* src_a[PA];
    const uint16_t* src_b[PB];
    auto piece = [&](int i, int& c8, int& row) {
        const int p = tid + NT * i, lane = p & 31, w = p >> 5;
        c8 = ((w & 1) << 2) | (lane & 3);
        row = ((w >> 1) << 3) | (lane >> 2);
    };
#pragma unroll
    for (int i = 0; i < PA; ++i) {
        int c8, row;
        piece(i, c8, row);
        const uint16_t* r = tid + NT * i < kM * 8 ? row_a(row) : nullptr;
        src_a[i] = r ? r + c8 * 8 : nullptr;
    }
#pragma unroll
    for (int i = 0; i < PB; ++i) {
        int c8, row;
        piece(i, c8, row);
        src_b[i] = tid + NT * i < n_blk * 8 ? row_b(row) + c8 * 8 : nullptr;
    }
    uint4 ra[kPre][PA], rb[kPre][PB];
    auto load = [&](uint4 (&va)[PA], uint4 (&vb)[PB], int kc) {
        if (kc >= nk) return;
#pragma unroll
        for (int i = 0; i < PA; ++i)
            va[i] = src_a[i] ? __ldcg(reinterpret_cast<const uint4*>(src_a[i] + kc * kBK)) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int i = 0; i < PB; ++i)
            if (src_b[i]) vb[i] = __ldcg(reinterpret_cast<const uint4*>(src_b[i] + kc * kBK));
    };
#pragma unroll
    for (int j = 0; j < kPre; ++j) load(ra[j], rb[j], j);
    for (int kc0 = 0; kc0 < nk; kc0 += kPre) {
#pragma unroll
        for (int j = 0; j < kPre; ++j) {
            const int kc = kc0 + j;
            if (kc < nk) {
                const uint32_t g = g0 + (uint32_t)kc;
                const int s = (int)(g % kStages);
                unsigned char* stage = ring + (size_t)s * kStageBytes;
                if (g >= (uint32_t)kStages) bar_wait(&bars->free_[s], (g / kStages - 1u) & 1u);
#pragma unroll
                for (int i = 0; i < PA; ++i) {
                    if (tid + NT * i < kM * 8) {
                        int c8, row;
                        piece(i, c8, row);
                        uint4 v = ra[j][i];
                        if (XFORM) xform(row, kc * kBK + c8 * 8, v);
                        *reinterpret_cast<uint4*>(stage + c8 * (kM * 16) + row * 16) = v;
                    }
                }
#pragma unroll
                for (int i = 0; i < PB; ++i) {
                    if (src_b[i]) {
                        int c8, row;
                        piece(i, c8, row);
                        *reinterpret_cast<uint4*>(stage + kStageA + c8 * (n_blk * 16) + row * 16) = rb[j][i];
                    }
                }
                load(ra[j], rb[j], kc + kPre);
                fence_async_smem();
                __syncthreads();
                if (tid == 0) {
                    fence_after_sync();
                    const uint32_t a0 = s2u(stage), b0 = a0 + kStageA;
                    const uint32_t lbo_a = kM * 16, lbo_b = (uint32_t)n_blk * 16, sbo = 128;
#pragma unroll
                    for (int q = 0; q < kBK / 16; ++q)
                        mma_bf16(tmem, smem_desc(a0 + 2 * q * lbo_a, lbo_a, sbo), smem_desc(b0 + 2 * q * lbo_b, lbo_b, sbo), idesc,
                                 (kc | q) ? 1u : 0u);
                    commit(&bars->free_[s]);
                    if (kc == nk - 1) commit(&bars->done);
                }
                __syncwarp();
            }
        }
    }
    pipe.chunk = g0 + (uint32_t)nk;
    bar_wait(&bars->done, pipe.tile & 1u);
    pipe.tile += 1;
    fence_after_sync();
}

// acc[128][b0.n + b1.n] (TMEM) = A[rows a_row0.. +128][K] * [B0 rows | B1 rows]^T.  All NT threads call this with uniform
// arguments; `ring` is 1024-byte aligned.  xform(r, k0, uint4&) rewrites elements [k0, k0 + 8) of tile row r (XFORM only).
// SKIP (micro-benchmark only): 1 = no copies (the producer only arrives on `full`), 2 = no MMAs (plain arrives on `free`).
template <int NT, bool XFORM, class Xform, int SKIP = 0>
__device__ __forceinline__ void tile_mma_tma_exp(unsigned char* ring, Bars* bars, Pipe& pipe, int K, const void* tm_a, int a_row0,
                                             BSrc b0, BSrc b1, Xform xform, int k0 = 0, int a_rows = kM,
                                             int weight_policy = kWeightsDefault) {
    const int nk = K / kBK;  // K elements starting at column k0 of both operands (split-K units)
    const int tid = threadIdx.x;
    const int n_blk = b0.n + b1.n;
    const uint32_t idesc = instr_desc(n_blk);
    const uint32_t tmem = bars->tmem_base;
    const uint32_t g0 = pipe.chunk;
    // a_rows: rows of the activation box behind tm_a (fewer than 128 when the batch is small: the MMA still spans 128
    // rows, the rest of the stage holds stale rows whose accumulator lanes nobody reads)
    const uint32_t stage_tx = (uint32_t)((a_rows + n_blk) * kBK * 2);
    const uint64_t w_policy = weight_policy == kWeightsKeep ? l2_policy_evict_last() : l2_policy_evict_first();
    // Two producing threads (a TMA instruction costs its issuing thread ~60 ns: measured 285 / 341 / 655 ns per stage with
    // 2 / 3 / 7 copies): `part` 0 = the activation box, 1 = the weight boxes, 2 = both (micro-benchmark variants).  Each
    // part arrives on `full` with its own byte count.
    auto produce = [&](int kc, int part) {
        const uint32_t g = g0 + (uint32_t)kc;
        const int s = (int)(g % kStages);
        if (g >= (uint32_t)kStages) bar_wait(&bars->free_[s], (g / kStages - 1u) & 1u);
        const uint32_t a = s2u(ring + (size_t)s * kStageBytes), b = a + kStageA;
        if (SKIP == 1) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(&bars->full_[s])) : "memory");
            return;
        }
        if (part != 1) {
            bar_expect_tx(&bars->full_[s], part == 2 ? stage_tx : (uint32_t)(a_rows * kBK * 2));
            tma_load_2d(a, tm_a, k0 + kc * kBK, a_row0, &bars->full_[s]);
        }
        if (part != 0) {
            if (part == 1) bar_expect_tx(&bars->full_[s], (uint32_t)(n_blk * kBK * 2));
            if (weight_policy == kWeightsDefault) {
                for (int j = 0; j < b0.n; j += b0.box) tma_load_2d(b + j * (kBK * 2), b0.tm, k0 + kc * kBK, b0.row0 + j, &bars->full_[s]);
                for (int j = 0; j < b1.n; j += b1.box)
                    tma_load_2d(b + (b0.n + j) * (kBK * 2), b1.tm, k0 + kc * kBK, b1.row0 + j, &bars->full_[s]);
            } else {
                for (int j = 0; j < b0.n; j += b0.box)
                    tma_load_2d_hint(b + j * (kBK * 2), b0.tm, k0 + kc * kBK, b0.row0 + j, &bars->full_[s], w_policy);
                for (int j = 0; j < b1.n; j += b1.box)
                    tma_load_2d_hint(b + (b0.n + j) * (kBK * 2), b1.tm, k0 + kc * kBK, b1.row0 + j, &bars->full_[s], w_policy);
            }
        }
    };
    auto issue_mma = [&](int kc) {  // one thread
        const uint32_t g = g0 + (uint32_t)kc;
        const int s = (int)(g % kStages);
        const uint32_t a = s2u(ring + (size_t)s * kStageBytes), b = a + kStageA;
        if (SKIP == 2) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(&bars->free_[s])) : "memory");
            if (kc == nk - 1) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(&bars->done)) : "memory");
            return;
        }
        fence_after_sync();
        if (SKIP == 4) {  // experiment: A operand copied shared -> tensor memory (tcgen05.cp), MMA reads it from there
#pragma unroll
            for (int q = 0; q < kBK / 16; ++q) {
                const uint32_t ta = tmem + 128u + 8u * (uint32_t)q;
                asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(ta), "l"(smem_desc_sw128(a + q * 32)) : "memory");
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                    ::"r"(tmem), "r"(ta), "l"(smem_desc_sw128(b + q * 32)), "r"(idesc), "r"((kc | q) ? 1u : 0u) : "memory");
            }
        } else {
#pragma unroll
        for (int q = 0; q < kBK / 16; ++q)
            mma_bf16(tmem + (uint32_t)((q % kAcc) * n_blk), smem_desc_sw128(a + q * 32), smem_desc_sw128(b + q * 32), idesc,
                     (kc > 0 || q >= kAcc) ? 1u : 0u);
        }
        // (the barriers of the shipped form expect kAcc arrivals: one issuing thread commits kAcc times)
#pragma unroll
        for (int i = 0; i < (SKIP == 0 ? kAcc : 1); ++i) {
            commit(&bars->free_[s]);
            if (kc == nk - 1) commit(&bars->done);
        }
    };
    if (!XFORM) {
        // two working threads; everybody else sleeps in the block barrier below instead of polling `done` (512 polling
        // threads take issue slots from the producer's and the MMA thread's dependent chains)
        if (tid == 32) {
            for (int kc = 0; kc < nk; ++kc) produce(kc, SKIP == 0 ? 0 : 2);
        } else if (SKIP == 0 && tid == 96) {
            for (int kc = 0; kc < nk; ++kc) produce(kc, 1);
        } else if (SKIP == 0 && (tid & 63) == 0 && tid < 64 * kAcc) {
            // kAcc issuing threads (warps 0, 2, 4, 6), one K step of every stage and one accumulator each; every one of them
            // commits its own MMAs, so `free` and `done` are initialised with kAcc arrivals (setup(..., kAcc))
            const int q = tid >> 6;
            for (int kc = 0; kc < nk; ++kc) {
                const uint32_t g = g0 + (uint32_t)kc;
                const int s = (int)(g % kStages);
                bar_wait(&bars->full_[s], (g / kStages) & 1u);
                const uint32_t a = s2u(ring + (size_t)s * kStageBytes), b = a + kStageA;
                fence_after_sync();
#pragma unroll
                for (int qq = 0; qq < kBK / 16; qq += kAcc)
                    mma_bf16(tmem + (uint32_t)(q * n_blk), smem_desc_sw128(a + (qq + q) * 32), smem_desc_sw128(b + (qq + q) * 32), idesc,
                             (kc > 0 || qq > 0) ? 1u : 0u);
                commit(&bars->free_[s]);
                if (kc == nk - 1) commit(&bars->done);
            }
            if (tid == 0) { bar_wait(&bars->done, pipe.tile & 1u); fence_before_sync(); }
        } else if (SKIP != 0 && tid == 0) {  // micro-benchmark variants: one issuing thread
            for (int kc = 0; kc < nk; ++kc) {
                const uint32_t g = g0 + (uint32_t)kc;
                bar_wait(&bars->full_[g % kStages], (g / kStages) & 1u);
                issue_mma(kc);
            }
            bar_wait(&bars->done, pipe.tile & 1u);
            fence_before_sync();
        }
        __syncwarp();
        __syncthreads();
        pipe.chunk = g0 + (uint32_t)nk;
        pipe.tile += 1;
        fence_after_sync();
        return;
    } else {
        if (tid == 32)
            for (int kc = 0; kc < kStages - 1 && kc < nk; ++kc) { produce(kc, 0); produce(kc, 1); }
        __syncwarp();
        for (int kc = 0; kc < nk; ++kc) {
            const uint32_t g = g0 + (uint32_t)kc;
            const int s = (int)(g % kStages);
            unsigned char* stage = ring + (size_t)s * kStageBytes;
            bar_wait(&bars->full_[s], (g / kStages) & 1u);
#pragma unroll
            for (int i = 0; i < (kM * 8 + NT - 1) / NT; ++i) {
                const int p = tid + NT * i;
                if (p < kM * 8) {
                    const int row = p >> 3, phys = p & 7, c8 = phys ^ (row & 7);
                    uint4* q = reinterpret_cast<uint4*>(stage + row * 128 + phys * 16);
                    uint4 v = *q;
                    xform(row, kc * kBK + c8 * 8, v);
                    *q = v;
                }
            }
            fence_async_smem();
            __syncthreads();
            if (tid == 0) issue_mma(kc);
            if (tid == 32 && kc + kStages - 1 < nk) { produce(kc + kStages - 1, 0); produce(kc + kStages - 1, 1); }
            __syncwarp();
        }
    }
    pipe.chunk = g0 + (uint32_t)nk;
    bar_wait(&bars->done, pipe.tile & 1u);
    pipe.tile += 1;
    fence_after_sync();
}

}  // namespace umma
}  // namespace smol
