// tcgen05 / TMEM building blocks for the batched decode path (sm_100a only).
//
// One output tile = 128 sequences (TMEM lanes) x n_blk weight rows (TMEM columns, fp32), accumulated by
// tcgen05.mma.cta_group::1.kind::f16 over K in ring stages of 64 elements.  Both operands are K-major
// (activations [row][K], weights [out][K] -- the checkpoint's own layout).
//
// The tile is tile_mma_tma: TMA tensor loads into 128-byte-swizzled stages, two producing threads, four MMA-issuing
// threads with an accumulator each; the accumulator is read back with tcgen05.ld.32x32b (thread = one sequence, 8
// consecutive columns per load).  Its measured-and-retired predecessors (per-thread cp.async / register staging into
// the no-swizzle layout, the A operand through tensor memory) live in tools/micro/umma_legacy.cuh with the
// micro-benchmark that compares them (profiles/r1c_umma_microbench.txt); they are not compiled into the library.
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

// ------------------------------------------------------------------------------------------------------------------
// Both operands are row-major [rows][K] bf16 buffers described by 2-D tensor maps with a 64-element x R-row box and the
// 128-byte swizzle, so a stage is rows of 128 bytes whose 16-byte chunks are XOR-ed with (row % 8) -- the canonical
// SWIZZLE_128B K-major layout of the UMMA descriptor (8-row groups 1024 B apart, K advanced by adding 32 B to the start
// address).  The bytes complete on the stage's `full` mbarrier; nothing but TMA and the tensor core touches the operands.
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

// The two TMA-issuing threads of a tile (warps 1 and 3).  decode_kernel.cu / tc_phases.cuh order the activations that other
// CTAs wrote with st.global against these threads' async-proxy reads (fence.proxy.async after every grid barrier): the
// fence predicates use the same constants.
constexpr int kTmaProducerA = 32;   // activation boxes
constexpr int kTmaProducerB = 96;   // weight boxes

// acc[128][b0.n + b1.n] (TMEM) = A[rows a_row0.. +128][K] * [B0 rows | B1 rows]^T.  All NT threads call this with uniform
// arguments; `ring` is 1024-byte aligned.  K elements starting at column k0 of both operands (split-K units).
template <int NT>
__device__ __forceinline__ void tile_mma_tma(unsigned char* ring, Bars* bars, Pipe& pipe, int K, const void* tm_a, int a_row0,
                                             BSrc b0, BSrc b1, int k0 = 0, int a_rows = kM, int weight_policy = kWeightsDefault) {
    const int nk = K / kBK;
    const int tid = threadIdx.x;
    const int n_blk = b0.n + b1.n;
    const uint32_t idesc = instr_desc(n_blk);
    const uint32_t tmem = bars->tmem_base;
    const uint32_t g0 = pipe.chunk;
    // a_rows: rows of the activation box behind tm_a (fewer than 128 when the batch is small: the MMA still spans 128
    // rows, the rest of the stage holds stale rows whose accumulator lanes nobody reads)
    const uint64_t w_policy = weight_policy == kWeightsKeep ? l2_policy_evict_last() : l2_policy_evict_first();
    // Two producing threads (a TMA instruction costs its issuing thread ~60 ns: measured 285 / 341 / 655 ns per stage with
    // 2 / 3 / 7 copies): `part` 0 = the activation box, 1 = the weight boxes.  Each part arrives on `full` with its own
    // byte count.
    auto produce = [&](int kc, int part) {
        const uint32_t g = g0 + (uint32_t)kc;
        const int s = (int)(g % kStages);
        if (g >= (uint32_t)kStages) bar_wait(&bars->free_[s], (g / kStages - 1u) & 1u);
        const uint32_t a = s2u(ring + (size_t)s * kStageBytes), b = a + kStageA;
        if (part == 0) {
            bar_expect_tx(&bars->full_[s], (uint32_t)(a_rows * kBK * 2));
            tma_load_2d(a, tm_a, k0 + kc * kBK, a_row0, &bars->full_[s]);
            return;
        }
        bar_expect_tx(&bars->full_[s], (uint32_t)(n_blk * kBK * 2));
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
    };
    // six working threads; everybody else sleeps in the block barrier below instead of polling `done` (512 polling
    // threads take issue slots from the producers' and the MMA threads' dependent chains)
    if (tid == kTmaProducerA) {
        for (int kc = 0; kc < nk; ++kc) produce(kc, 0);
    } else if (tid == kTmaProducerB) {
        for (int kc = 0; kc < nk; ++kc) produce(kc, 1);
    } else if ((tid & 63) == 0 && tid < 64 * kAcc) {
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
    }
    __syncwarp();
    __syncthreads();
    pipe.chunk = g0 + (uint32_t)nk;
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
