// tcgen05 / TMEM building blocks for the batched decode path (sm_100a only).
//
// One output tile = 128 sequences (TMEM lanes) x n_blk weight rows (TMEM columns, fp32), accumulated by
// tcgen05.mma.cta_group::1.kind::f16 over K in ring stages of 64 elements.  Both operands are K-major
// (activations [row][K], weights [out][K] -- the checkpoint's own layout).
//
// THE SHIPPED FORM is tile_mma_tma (further down): TMA tensor loads into 128-byte-swizzled stages, two producing
// threads, four MMA-issuing threads with an accumulator each.  The two forms at the top of the file are its measured
// predecessors, kept for the A/B in tools/micro/umma_test (profiles/r1c_umma_microbench.txt): they stage the operands
// with per-thread 16-byte copies (cp.async, or registers) into the canonical NO-swizzle layout of the UMMA descriptor,
//
//     byte offset of element (row r, k) inside a stage = (k / 8) * LBO + (r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2
//
// (8-row x 16-byte core matrices, SBO = 128 B between 8-row groups, LBO = rows * 16 B between the 16-byte K chunks),
// with an optional in-place transform of the thread's own pieces, fence.proxy.async and a block barrier per stage,
// and cost 0.8-1.2 us per stage against 0.29-0.5 us.  All forms read the accumulator back with tcgen05.ld.32x32b
// (thread = one sequence, 8 consecutive columns per load).
#pragma once

#include <stdint.h>

namespace smol {
namespace umma {

constexpr int kM = 128;                      // sequences per tile (TMEM lanes)
constexpr int kBK = 64;                      // K elements per ring stage (4 MMAs of K = 16)
constexpr int kMaxN = 96;                    // widest tile (w1|w3: 48 + 48 rows)
#ifndef UMMA_STAGES
#define UMMA_STAGES 5
#endif
constexpr int kStages = UMMA_STAGES;
constexpr int kAhead = 3;                    // stages in flight ahead of the MMA
constexpr int kStageA = kM * kBK * 2;        // 16 KB
constexpr int kStageB = kMaxN * kBK * 2;     // 12 KB
constexpr int kStageBytes = kStageA + kStageB;
constexpr int kRingBytes = kStages * kStageBytes;  // 140 KB
constexpr int kAcc = 4;                       // accumulators per tile (TMA form) = issuing threads: K step q of every stage is issued by
                                             // thread 64 * q into accumulator q (disjoint TMEM column ranges, summed by the epilogue).
                                             // Measured (tools/micro/umma_test a): ONE thread issuing the four tcgen05.mma of a stage
                                             // needs ~360 ns per stage whatever M (64 / 128), N (16..32) or the accumulator dependencies
                                             // are -- the issue path of the thread, not the tensor pipe; four issuing threads: 305 ns
                                             // (then bounded by the TMA stream, 285 ns alone).
constexpr int kTmemCols = 512;               // allocation (power of two >= kAcc * kMaxN)

struct Bars {
    uint64_t free_[kStages];  // stage consumed by its MMAs (tcgen05.commit)
    uint64_t full_[kStages];  // TMA form: the stage's bytes have landed (complete_tx)
    uint64_t done;            // accumulator of the tile complete
    uint32_t tmem_base;
    uint32_t pad;
};

// Running counters (uniform across the CTA): ring stages used and tiles finished since the barriers were initialised.
struct Pipe {
    uint32_t chunk;
    uint32_t tile;
};

__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2u(bar)), "r"(count) : "memory");
}
// Bounded wait: a barrier that never completes is a protocol bug -- trap instead of hanging the GPU.
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = s2u(bar);
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(dst_smem)), "r"((uint32_t)kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t base) {  // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"((uint32_t)kTmemCols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, no swizzle (layout type 0), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// D fp32, A/B bf16, both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t instr_desc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2u(bar)) : "memory");
}
// 8 consecutive fp32 columns of this thread's lane (lane quadrant = warp % 4).
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Called once per kernel by all threads (before any tile): barriers + TMEM allocation by warp 0.
__device__ __forceinline__ void setup(Bars* bars, Pipe& pipe, uint32_t issuers = 1u, uint32_t producers = 1u) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { bar_init(&bars->free_[s], issuers); bar_init(&bars->full_[s], producers); }
        bar_init(&bars->done, issuers);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if ((threadIdx.x >> 5) == 0) tmem_alloc(&bars->tmem_base);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    pipe.chunk = 0;
    pipe.tile = 0;
}
__device__ __forceinline__ void teardown(Bars* bars) {
    fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_free(bars->tmem_base);
}

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

// ------------------------------------------------------------------------------------------------------------------
// TMA form (the one the decode kernel uses).  Both operands are row-major [rows][K] bf16 buffers described by 2-D tensor
// maps with a 64-element x R-row box and the 128-byte swizzle, so a stage is rows of 128 bytes whose 16-byte chunks are
// XOR-ed with (row % 8) -- the canonical SWIZZLE_128B K-major layout of the UMMA descriptor (8-row groups 1024 B apart,
// K advanced by adding 32 B to the start address).  One thread issues the copies of a stage (activation box of 128 rows,
// weight boxes of 16 rows), the bytes complete on the stage's `full` mbarrier, one thread issues the MMAs; nothing else
// touches the operands unless a transform (RMSNorm) has to rewrite the activation pieces in shared memory.
// Measured on B200 (tools/micro/umma_test): per-thread 16-byte loads (cp.async or registers) cost 0.7-1.0 us per stage --
// 8 cache lines per warp instruction through the L1 pipe -- which is what this form removes.
// ------------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(s2u(bar)), "r"(c0), "r"(c1) : "memory");
}
// The same with an L2 eviction policy (createpolicy): weights that are read once per frame leave L2 first, weights that
// every depth step re-reads stay.
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const void* tmap, int c0, int c1, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"(tmap), "r"(s2u(bar)), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// K-major, 128-byte swizzle (layout type 2), 8-row groups 1024 B apart, descriptor version 1.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// One weight operand: rows [row0, row0 + n) of tensor map `tm`, copied in boxes of `box` rows (n a multiple of box).
struct BSrc {
    const void* tm;   // map whose box is `box` rows high
    int row0, n, box;
};
enum WeightPolicy { kWeightsDefault = 0, kWeightsStream = 1, kWeightsKeep = 2 };

// acc[128][b0.n + b1.n] (TMEM) = A[rows a_row0.. +128][K] * [B0 rows | B1 rows]^T.  All NT threads call this with uniform
// arguments; `ring` is 1024-byte aligned.  xform(r, k0, uint4&) rewrites elements [k0, k0 + 8) of tile row r (XFORM only).
// SKIP (micro-benchmark only): 1 = no copies (the producer only arrives on `full`), 2 = no MMAs (plain arrives on `free`).
template <int NT, bool XFORM, class Xform, int SKIP = 0>
__device__ __forceinline__ void tile_mma_tma(unsigned char* ring, Bars* bars, Pipe& pipe, int K, const void* tm_a, int a_row0,
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

// Hands 8 consecutive accumulator columns of one sequence to epi(row, col0, v[8]); with `paired` the thread also gets the
// columns half a tile further (w1 | w3 halves): epi2(row, col0, gate[8], up[8]).  Ends with the fences + block barrier
// that let the next tile overwrite the accumulator.
// acc_stride > 0: the tile was accumulated in kAcc accumulators acc_stride columns apart (TMA form).  All loads of a call
// are issued before the one tcgen05.wait::ld.
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_sum(const uint32_t (&r)[kAcc][8], int n_acc, float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[0][i]);
#pragma unroll
    for (int a = 1; a < kAcc; ++a)
        if (a < n_acc) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __fadd_rn(v[i], __uint_as_float(r[a][i]));
        }
}
__device__ __forceinline__ void tmem_ld8_sum(uint32_t taddr, int acc_stride, float (&v)[8]) {
    uint32_t r[kAcc][8];
    const int n_acc = acc_stride > 0 ? kAcc : 1;
#pragma unroll
    for (int a = 0; a < kAcc; ++a)
        if (a < n_acc) tmem_ld8_issue(taddr + (uint32_t)(a * acc_stride), r[a]);
    tmem_ld_wait();
    tmem_sum(r, n_acc, v);
}
__device__ __forceinline__ void tmem_ld8_sum2(uint32_t ta, uint32_t tb, int acc_stride, float (&va)[8], float (&vb)[8]) {
    uint32_t ra[kAcc][8], rb[kAcc][8];
    const int n_acc = acc_stride > 0 ? kAcc : 1;
#pragma unroll
    for (int a = 0; a < kAcc; ++a)
        if (a < n_acc) {
            tmem_ld8_issue(ta + (uint32_t)(a * acc_stride), ra[a]);
            tmem_ld8_issue(tb + (uint32_t)(a * acc_stride), rb[a]);
        }
    tmem_ld_wait();
    tmem_sum(ra, n_acc, va);
    tmem_sum(rb, n_acc, vb);
}

template <class Epi>
__device__ __forceinline__ void tile_epilogue(Bars* bars, int n_cols, Epi epi, int acc_stride = 0) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = ((w & 3) << 5) | lane;
    const uint32_t base = bars->tmem_base + ((uint32_t)((w & 3) << 5) << 16);
    const int n_warps = blockDim.x >> 5;
    for (int cg = w >> 2; cg * 8 < n_cols; cg += n_warps >> 2) {
        float v[8];
        tmem_ld8_sum(base + (uint32_t)(cg * 8), acc_stride, v);
        epi(row, cg * 8, v);
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
}
template <class Epi2>
__device__ __forceinline__ void tile_epilogue_paired(Bars* bars, int half_cols, Epi2 epi2, int acc_stride = 0) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = ((w & 3) << 5) | lane;
    const uint32_t base = bars->tmem_base + ((uint32_t)((w & 3) << 5) << 16);
    const int n_warps = blockDim.x >> 5;
    for (int cg = w >> 2; cg * 8 < half_cols; cg += n_warps >> 2) {
        float a[8], b[8];
        tmem_ld8_sum2(base + (uint32_t)(cg * 8), base + (uint32_t)(half_cols + cg * 8), acc_stride, a, b);
        epi2(row, cg * 8, a, b);
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
}

}  // namespace umma
}  // namespace smol
