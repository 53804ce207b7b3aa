// The DualAR decode step as a DATA-FLOW persistent kernel (sm_100a) -- the bs=1 latency path.
//
// Same phase program, same arithmetic and the same rounding points as decode_kernel.cu (results are
// bit-identical, tests/test_gpu_dataflow.py), but no grid barrier anywhere.  At batch 1 a frame is a
// chain of ~200 dependent matrix-vector products; what bounds it is the hand-off between them and the
// instruction stream, not bytes.  So:
//
//  * every phase PUBLISHES its output as 8-byte words {two bf16 | epoch} ("LL" protocol: the flag
//    travels in the same store as the data).  A consumer polls exactly the words it needs with
//    ld.relaxed.gpu and starts the moment they carry the epoch of (frame, phase): one L2 store + one L2
//    load per hand-off, no fence, no atomic, no barrier, no skew wait.  Each vector is written to
//    kLLRep replicas so that the 148 polling CTAs spread over more L2 slices.
//  * weights never wait for activations: the last warp of every CTA is a TMA producer that streams the
//    CTA's rows of ALL coming phases (cp.async.bulk -> shared-memory ring, full/empty mbarriers) as
//    far ahead as the ring allows (~1.5 layers), so HBM runs continuously under the compute chain.
//  * sampled ids are published once per CTA; each CTA tracks seq_len / step / finished itself, so
//    one launch runs any number of frames with no host round trip and no global state on the path.
//  * the whole per-frame loop is ONE small body (a few thousand instructions, batch 1 only): every phase
//    of a frame runs exactly once, so code that does not stay in the instruction cache costs an L2 round
//    trip per 128 bytes of SASS -- an earlier, fully templated version of this kernel spent most of its
//    time fetching instructions.  Cold paths (embedding, frame boundary, non-greedy sampling) are out of
//    line.
//
// Reference map (P = modeling/model/rq_transformer.py, M = mlx lm/rq_transformer.py, G = mlx
// lm/generate.py): embed P:205-221; RMSNorm P:601-613; QKV/RoPE/attention P:535-570,616-640;
// FeedForward P:573-582; block P:492-501; slow step P:223-260 / M:173-192; depth loop P:409-448,
// M:194-220, G:110-141; sampling G:88-99,118-132; frame assembly + stop rule G:143-171.

#define SMOL_BLOCK_SYNC() asm volatile("bar.sync 1, 352;" ::: "memory")  // the 11 consumer warps

#include "common.cuh"
#include "dev_model.h"
#include "sampler.cuh"

namespace smol {

namespace ll {

constexpr int kStages = 16;   // mbarrier pairs of the weight ring (stages in flight)
constexpr int kLLDepth = 8;   // depth positions the in-register depth attention carries
constexpr int kDescWords = 16;  // 64-byte per-CTA phase descriptor (see build_desc)

// A word that never arrives is a protocol bug, not a wait: trap after ~1 s of spinning instead of hanging the GPU.
// (A back-off between retries was measured too: it only adds latency.)
#define LL_SPIN_GUARD(n) do { if (++(n) > (1u << 22)) __trap(); } while (0)

// ---- small helpers -------------------------------------------------------------------------------
constexpr int kCons = kLLWarps * 32;  // consumer threads
__device__ __forceinline__ void csync() { SMOL_BLOCK_SYNC(); }  // the consumer warps
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int ldcg_i32(const void* p) { return __ldcg(reinterpret_cast<const int*>(p)); }
__device__ __forceinline__ uint4 ldcg_v4(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    f[0] = bf_lo(v.x); f[1] = bf_hi(v.x); f[2] = bf_lo(v.y); f[3] = bf_hi(v.y);
    f[4] = bf_lo(v.z); f[5] = bf_hi(v.z); f[6] = bf_lo(v.w); f[7] = bf_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 v;
    v.x = pack_bf16(f[0], f[1]); v.y = pack_bf16(f[2], f[3]); v.z = pack_bf16(f[4], f[5]); v.w = pack_bf16(f[6], f[7]);
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// shared-memory accesses by 32-bit shared address (keeps the compiler from falling back to generic LD/ST)
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) { return __uint_as_float(lds_u32(a)); }
__device__ __forceinline__ void sts_v4(uint32_t a, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { sts_u32(a, __float_as_uint(v)); }

// ---- LL words ------------------------------------------------------------------------------------
// Polling loads: ld.relaxed.gpu (a strong load that always goes to L2).  ld.global.cv looks twice as fast in an isolated
// exchange micro-benchmark (tools/micro/ll_atom.cu) but makes no difference inside this kernel (the L1 next to 227 KB of
// shared memory is too small to matter); atomics as polls are slower.  Each 8-byte word {payload, epoch} is one aligned
// access, so a reader sees a word entirely old or entirely new.
__device__ __forceinline__ uint4 ld_relaxed_v4(const void* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint2 ld_relaxed_v2(const void* p) {
    uint2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_v2(void* p, uint32_t a, uint32_t b) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
// one word, polled
__device__ __forceinline__ uint32_t ll_get(const unsigned long long* w, uint32_t epoch) {
    uint2 v = ld_relaxed_v2(w);
    uint32_t spins = 0;
    while (v.y != epoch) { LL_SPIN_GUARD(spins); v = ld_relaxed_v2(w); }
    return v.x;
}
// 8 consecutive elements = 4 words = two 16-byte loads, polled together; returns the 8 bf16
__device__ __forceinline__ uint4 ll_get8(const unsigned long long* w, uint32_t epoch) {
    uint4 a = ld_relaxed_v4(w), b = ld_relaxed_v4(w + 2);
    uint32_t spins = 0;
    while (a.y != epoch || a.w != epoch || b.y != epoch || b.w != epoch) {
        LL_SPIN_GUARD(spins);
        a = ld_relaxed_v4(w);
        b = ld_relaxed_v4(w + 2);
    }
    return make_uint4(a.x, a.z, b.x, b.z);
}
// publish one word to `nrep` replicas `len` words apart
__device__ __forceinline__ void ll_put(unsigned long long* w, int len, int nrep, uint32_t payload, uint32_t epoch) {
#pragma unroll 1
    for (int r = 0; r < nrep; ++r) st_relaxed_v2(w + (size_t)r * len, payload, epoch);
}

// ---- mbarrier + TMA bulk copy ----------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// relaxed: the arriving warp's reads of the stage have been consumed (data dependence) and nothing it wrote
// needs to be ordered before the producer's refill -- a releasing arrive would wait for its global stores
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    for (;;) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        LL_SPIN_GUARD(spins);
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}

// ---- per-CTA shared state ------------------------------------------------------------------------
__shared__ __align__(8) uint64_t s_full[kStages];
__shared__ __align__(8) uint64_t s_empty[kStages];
__shared__ SampleScratch s_sc;
__shared__ float s_part[kWarps];
__shared__ float s_kpart[64];          // partial sums of the K slices of long rows (w2)
__shared__ uint32_t s_cand[kWarps];    // greedy candidates of the warps (bf16 logit << 16 | 4095 - index... see cand_pack)
__shared__ int s_pos, s_step, s_fin;   // cached positions (seq_len) / frames emitted / stop flag, tracked locally
__shared__ int s_tok[kMaxRows];        // pending input column
__shared__ int s_nw[kMaxRows];         // ids sampled in the current frame
__shared__ int s_spf, s_wodd;          // ring stages per iteration / (weight phases per iteration) & 1
__shared__ uint32_t s_epoch0;          // phases executed by earlier launches
constexpr int kBtabCache = 256;
__shared__ int s_btab[kBtabCache];     // the sequence's block table (constant during a launch): saves an L2 round trip per KV access

struct SmemPlan {
    int xs_bytes;       // one activation buffer
    int res_bytes;      // one residual buffer
    int scratch_bytes;
    int desc_bytes;     // per-CTA phase descriptors
    int fq_bytes;       // depth transformer: current query row
    int fkv_bytes;      // depth transformer: K/V of the frame's depth steps [n_flayer][kLLDepth][2 * n_kv * 64] bf16
    int ring_bytes;
};

// ---- weight plan ---------------------------------------------------------------------------------
// A phase's GEMV is U units; a unit is one row PAIR (two adjacent outputs = one LL word); the gated MLP
// unit is two adjacent rows of w1 plus the same two rows of w3 (two ring stages).  CTA c owns units
// [U*c/n, U*(c+1)/n): one contiguous row block per matrix -> one bulk copy.
struct Plan {
    const uint16_t* w0;
    const uint16_t* w1;
    int K, U;
};

__device__ __forceinline__ Plan phase_plan(const DevModel& M, const Phase& ph) {
    Plan pl;
    pl.w0 = nullptr; pl.w1 = nullptr; pl.K = 0; pl.U = 0;
    const bool fast = ph.fast != 0;
    const int D = fast ? M.fdim : M.dim, F = fast ? M.finter : M.inter;
    const DevLayer& L = fast ? M.fast_layers[ph.layer] : M.layers[ph.layer];
    switch (ph.kind) {
        case PH_QKV: {
            const int Hq = fast ? M.fn_head : M.n_head, Hkv = fast ? M.fn_kv : M.n_kv;
            pl.w0 = L.wqkv; pl.K = D; pl.U = (Hq + 2 * Hkv) * kHeadDim / 2;
        } break;
        case PH_WO: pl.w0 = L.wo; pl.K = D; pl.U = D / 2; break;
        case PH_W13: pl.w0 = L.w1; pl.w1 = L.w3; pl.K = D; pl.U = F / 2; break;
        case PH_W2: pl.w0 = L.w2; pl.K = F; pl.U = D / 2; break;
        case PH_HEAD: {
            const int N = fast ? M.codebook_size : M.vocab;
            pl.w0 = fast ? M.fast_output + (M.depthwise_output ? (size_t)ph.depth_pos * N * D : 0) : M.head;
            pl.K = D; pl.U = N / 2;
        } break;
        default: break;
    }
    return pl;
}
__device__ __forceinline__ int unit_begin(int U, int cta, int n) { return (U * cta) / n; }
// ring allocator replayed identically by the producer and by the consumers
__device__ __forceinline__ uint32_t ring_alloc(uint32_t& head, uint32_t bytes, uint32_t ring_bytes) {
    if (head + bytes > ring_bytes) head = 0;
    const uint32_t start = head;
    head += bytes;
    return start;
}

// ---- TMA producer (one lane of the last warp) -------------------------------------------------------------
__device__ __noinline__ void producer(const DevModel& M, const CallArgs& A, uint32_t dsc0, uint32_t ring, uint32_t ring_bytes) {
    const int cta = blockIdx.x, n_ctas = gridDim.x;
    uint64_t pol_stream, pol_keep;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    uint32_t held[kStages];  // ring bytes (gap included) held by each in-flight stage
    uint32_t wseq = 0, rel = 0, head = 0, used = 0;
    const uint32_t full0 = smem_u32(s_full), empty0 = smem_u32(s_empty);
    for (int it = 0; it < A.n_iter; ++it) {
        for (int p = A.phase_begin; p < A.phase_end; ++p) {
            const Phase ph = unpack_phase(M.prog[p]);
            if (ph.kind == PH_ATTN || ph.kind == PH_SAMPLE) continue;
            const Plan pl = phase_plan(M, ph);
            const int u0 = unit_begin(pl.U, cta, n_ctas);
            const uint32_t bytes = (uint32_t)(unit_begin(pl.U, cta + 1, n_ctas) - u0) * 4u * (uint32_t)pl.K;
            // the slow transformer's weights are read once per frame: evict first; the depth
            // transformer's are re-read by every depth step: keep them in L2
            const uint64_t pol = (ph.fast && ph.kind != PH_HEAD) ? pol_keep : pol_stream;
            for (int part = 0; part < (pl.w1 ? 2 : 1); ++part) {
                const uint32_t start = lds_u32(dsc0 + (uint32_t)p * (kDescWords * 4u) + 56u + 4u * part);
                const uint32_t gap = (start < head) ? (ring_bytes - head) : (start - head);  // wrap (or restart of the layout)
                // wait until the ring has room and the stage's barrier pair is free
                while (used + gap + bytes > ring_bytes || wseq - rel >= (uint32_t)kStages) {
                    mbar_wait(empty0 + 8u * (rel % kStages), (rel / kStages) & 1u);
                    used -= held[rel % kStages];
                    ++rel;
                }
                head = start + bytes;
                held[wseq % kStages] = gap + bytes;
                used += gap + bytes;
                const uint32_t bar = full0 + 8u * (wseq % kStages);
                mbar_expect_tx(bar, bytes);
                if (bytes) bulk_g2s(ring + start, (part ? pl.w1 : pl.w0) + (size_t)2 * u0 * pl.K, bytes, bar, pol);
                ++wseq;
            }
        }
    }
}

// ---- cold paths --------------------------------------------------------------------------------------
__device__ __forceinline__ int input_token(const DevModel& M, const CallArgs& A, int iter, int r) {
    if (A.mode == 1) {
        int t = iter + A.iter_base;
        const int len = ldcg_i32(A.prompt_len);
        if (t > len - 1) t = len - 1;
        return ldcg_i32(A.prompt + (size_t)r * A.s_max + t);
    }
    return s_tok[r];
}
// BaseTransformer.embed (P:205-221) for one 8-element chunk (once per frame)
__device__ __noinline__ uint4 embed_chunk(const DevModel& M, const CallArgs& A, int iter, int ch) {
    const int D = M.dim;
    const int t0 = input_token(M, A, iter, 0);
    bool use_vq;
    if (M.mlx_embed_mask) use_vq = (t0 >= M.semantic_start && t0 <= M.semantic_end);
    else use_vq = input_token(M, A, iter, 1) != 0;
    float f[8];
    unpack8(ldcg_v4(M.embeddings + (size_t)t0 * D + ch * 8), f);
    if (use_vq) {
        float s[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = 0.f;
        for (int r0 = 1; r0 < M.n_rows; r0 += 8) {  // eight row loads in flight, summed in row order
            uint4 rows[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + i;
                if (r < M.n_rows) {
                    const int row = input_token(M, A, iter, r) + (M.dup0 ? (r - 1) : r) * M.codebook_size;
                    rows[i] = ldcg_v4(M.codebook_embeddings + (size_t)row * D + ch * 8);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (r0 + i < M.n_rows) {
                    float g[8];
                    unpack8(rows[i], g);
#pragma unroll
                    for (int e = 0; e < 8; ++e) s[e] = __fadd_rn(s[e], g[e]);
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = bf16_round(__fadd_rn(f[e], bf16_round(s[e])));
    }
    return pack8(f);
}

__device__ __forceinline__ int sample_phase_index(const DevModel& M, int r) {  // program index of row r's SAMPLE phase
    const int n_slow = 5 * M.n_layer;
    if (r == 0) return n_slow + 1;
    return n_slow + 2 + (r - 1) * (4 * M.n_flayer + 2) + 4 * M.n_flayer + 1;
}
// ids published by the sampler CTA: word of row r for this CTA, epoch of that row's SAMPLE phase
__device__ __forceinline__ int token_word_get(const DevModel& M, int r, uint32_t epoch) {
    return (int)ll_get(M.ll_tok + (size_t)r * kLLMaxCtas + blockIdx.x, epoch);
}
// Greedy rows (temperature 0) and forced ids need no sampler CTA: every CTA works the id out itself (argmax over
// the published logits / the forced id) -- no SAMPLE hand-off on the critical path.
__device__ __forceinline__ bool row_is_local(const DevModel& M, const CallArgs& A, int r) {
    if (A.repeat & 2) return false;  // experiment switch: always go through the sampler CTA
    return M.force != nullptr || (r == 0 ? A.s.temp : A.s.fast_temp) == 0.0f;
}
// Greedy candidate = (logit, index) packed so that unsigned max picks the larger logit and, among equal logits, the
// smaller index (torch / mx argmax tie rule): an order-preserving map of the bf16 logit in the high 16 bits, the
// inverted index in the low 16.  Logits are bf16-rounded, so nothing is lost.
__device__ __forceinline__ uint32_t cand_pack(float v, int idx) {
    uint32_t b = __float_as_uint(v) >> 16;
    b = (b & 0x8000u) ? (~b & 0xffffu) : (b | 0x8000u);
    return (b << 16) | (uint32_t)(0xffff - idx);
}
__device__ __forceinline__ int cand_index(uint32_t c) { return 0xffff - (int)(c & 0xffffu); }
// frame boundary: adopt the ids sampled in the previous frame (G:143-166).  e_prev0 = epoch of phase 0 of
// the previous frame.
__device__ __noinline__ void frame_boundary(const DevModel& M, const CallArgs& A, uint32_t e_prev0) {
    const int tid = threadIdx.x;
    if (tid < M.n_rows && !row_is_local(M, A, tid)) s_nw[tid] = token_word_get(M, tid, e_prev0 + (uint32_t)sample_phase_index(M, tid));
    csync();
    if (tid == 0 && s_fin == 0) {
        for (int r = 0; r < M.n_rows; ++r) s_tok[r] = s_nw[r];
        s_pos += 1;
        s_step += 1;
        if (A.b.finished != nullptr && A.s.audio_only && !A.s.ignore_stop && s_nw[0] == M.im_end) s_fin = 1;
    }
    __threadfence();  // acquire side of the once-per-frame fence: KV appended last frame is visible
    csync();
}

__device__ __forceinline__ const unsigned long long* ll_src(const DevModel& M, int p, int rep) {
    return M.ll + M.ll_off[p] + (size_t)rep * M.ll_len[p];
}
__device__ __forceinline__ unsigned long long* ll_dst(const DevModel& M, int p) { return M.ll + M.ll_off[p]; }

// ---- attention of the depth transformer (<= 8 positions) -----------------------------------------------
// Recomputed by every CTA into xs (the input of the wo GEMV).  Step A: all consumer threads poll this depth
// step's q|k|v words once and park them in shared memory (q row; K/V appended to the CTA's own copy of the
// frame's depth K/V) -- the history of earlier steps is already there, so a phase reads 5 KB from L2 instead
// of re-reading every earlier step.  Step B: one warp per head; arithmetic identical to decode_kernel.cu:
// fast_attention_rows (4 lanes per position, two-pass softmax, probabilities rounded to bf16 for PV).
__device__ __forceinline__ void fast_attention(const DevModel& M, int layer, int depth_pos, const unsigned long long* qkv,
                                               uint32_t e_qkv, uint32_t fq, uint32_t fkv, uint32_t xs) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Hq = M.fn_head, Hkv = M.fn_kv, G = Hq / Hkv;
    const int q_rows = Hq * kHeadDim, kvw = 2 * Hkv * kHeadDim;  // elements
    const uint32_t slot = fkv + (uint32_t)((layer * kLLDepth + depth_pos) * kvw) * 2u;
    for (int i = tid; i < (q_rows + kvw) / 4; i += kCons) {  // 4 elements = 2 words = one 16-byte load
        uint4 v = ld_relaxed_v4(qkv + 2 * i);
        uint32_t spins = 0;
        while (v.y != e_qkv || v.w != e_qkv) { LL_SPIN_GUARD(spins); v = ld_relaxed_v4(qkv + 2 * i); }
        const int e = 4 * i;
        const uint32_t dst = e < q_rows ? fq + (uint32_t)e * 2u : slot + (uint32_t)(e - q_rows) * 2u;
        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(dst), "r"(v.x), "r"(v.z) : "memory");
    }
    csync();
    const int j = lane >> 2, part = lane & 3;
    const bool ok = j <= depth_pos;
    const uint32_t lay = fkv + (uint32_t)(layer * kLLDepth * kvw) * 2u;
#pragma unroll 1
    for (int hq = warp; hq < Hq; hq += kLLWarps) {
        const int kvh = hq / G;
        float s = 0.f;
        if (ok) {
            const uint32_t qa = fq + (uint32_t)(hq * kHeadDim + part * 16) * 2u;
            const uint32_t ka = lay + (uint32_t)(j * kvw + kvh * kHeadDim + part * 16) * 2u;
            float qf[8], kf[8];
            unpack8(lds_v4(qa), qf); unpack8(lds_v4(ka), kf);
#pragma unroll
            for (int e = 0; e < 8; ++e) s = fmaf(qf[e], kf[e], s);
            unpack8(lds_v4(qa + 16u), qf); unpack8(lds_v4(ka + 16u), kf);
#pragma unroll
            for (int e = 0; e < 8; ++e) s = fmaf(qf[e], kf[e], s);
        }
        uint32_t vw[kLLDepth];
#pragma unroll
        for (int jj = 0; jj < kLLDepth; ++jj)
            vw[jj] = jj <= depth_pos ? lds_u32(lay + (uint32_t)(jj * kvw + Hkv * kHeadDim + kvh * kHeadDim + 2 * lane) * 2u) : 0u;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const float sc = ok ? s * 0.125f : -INFINITY;
        float m = sc;
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
        const float pe = ok ? expf(sc - m) : 0.f;
        float l = pe + 0.f;
        l += __shfl_xor_sync(0xffffffffu, l, 4);
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 16);
        const float pb = bf16_round(pe);
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int jj = 0; jj < kLLDepth; ++jj) {
            const float pj = __shfl_sync(0xffffffffu, pb, jj * 4);
            o0 = fmaf(pj, bf_lo(vw[jj]), o0);
            o1 = fmaf(pj, bf_hi(vw[jj]), o1);
        }
        const float inv = 1.0f / l;
        sts_u32(xs + (uint32_t)(hq * kHeadDim + 2 * lane) * 2u, pack_bf16(bf16_round(o0 * inv), bf16_round(o1 * inv)));
    }
}

// ---- split-KV decode attention over the paged cache (slow layers) ----------------------------------
// Unit = (query head, split): the same splits, the same position -> (warp, lane group) map, the same merge
// order and arithmetic as decode_kernel.cu: phase_attn_g, so results are bit-identical -- but one head per
// unit instead of a GQA group (K/V come from L2 either way; 3x the CTAs work in parallel and the code is a
// third of the size).  q and the newest position's K/V come from the QKV phase's published words, older
// positions from the paged pool; the result leaves as LL words: directly (one split) or as (m, l, o[64])
// partials that the head's combiner CTA merges in split order.
__device__ __forceinline__ int attn_splits(const DevModel& M, int n_ctas) {
    (void)M; (void)n_ctas;
    return kMaxSplits;  // splits depend on the sequence's length only (decode_kernel.cu: attn_splits)
}
__device__ __forceinline__ unsigned long long* partial_words(const DevModel& M, int layer, int hq, int s) {
    return M.ll_partial + (((size_t)(layer & 1) * M.n_head + hq) * kMaxSplits + s) * kPartialStride;
}

__device__ __noinline__ void phase_attn(const DevModel& M, const CallArgs& A, int layer, int p, uint32_t epoch, uint32_t scratch) {
    const int cta = blockIdx.x, n_ctas = gridDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rep = cta % kLLRep;
    const int Hq = M.n_head, Hkv = M.n_kv, ps = M.page_size, G = Hq / Hkv;
    const int q_words = Hq * kHeadDim / 2, k_end_words = (Hq + Hkv) * kHeadDim / 2;
    const int dch = lane & 7, psub = lane >> 3;
    const int s_max = attn_splits(M, n_ctas);
    const int pos_new = s_pos;
    int Lb = pos_new + 1;
    const int cap = A.b.max_pages * ps;
    if (Lb > cap) Lb = cap;
    int ns = (Lb + kSplitMin - 1) / kSplitMin;
    if (ns > s_max) ns = s_max;
    if (ns < 1) ns = 1;
    const int chunk = (Lb + ns - 1) / ns;
    const int n_units = Hq * ns;
    const int len_out = M.ll_len[p];
    const unsigned long long* qkv = ll_src(M, p - 1, rep);
    const size_t head_stride = (size_t)ps * kHeadDim;

    for (int u = cta; u < n_units; u += n_ctas) {
        const int hq = u / ns, s = u - hq * ns, kvh = hq / G;
        const int p0 = s * chunk, p1 = min(Lb, p0 + chunk);
        float qf[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) qf[e] = 0.f;
        bool have_q = false;
        // the position -> (warp slot, lane group) map has 16 warp slots (decode_kernel.cu); 11 consumer warps carry them
#pragma unroll 1
        for (int vw = warp; vw < kWarps; vw += kLLWarps) {
        float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        for (int pb = p0 + vw * 4; pb < p1; pb += kWarps * 4) {
            const int pp = pb + psub;
            const bool valid = pp < p1;
            const bool newest = valid && (pp == pos_new);
            uint4 kk = make_uint4(0u, 0u, 0u, 0u), vv = kk;
            if (valid && !newest) {  // cached position: issue the pool loads before waiting for q
                const int pg = pp / ps;
                const int page = pg < kBtabCache ? s_btab[pg] : ldcg_i32(A.b.block_table + pg);
                const uint16_t* kp = M.kv_pool + ((((size_t)page * M.n_layer + layer) * 2) * Hkv + kvh) * head_stride
                                     + (size_t)(pp % ps) * kHeadDim + dch * 8;
                kk = ldcg_v4(kp);
                vv = ldcg_v4(kp + (size_t)Hkv * head_stride);
            }
            {
                // the polls of this step -- q (first step only) and, for the lane group that owns the newest position, its
                // K/V from the QKV phase's words -- are in flight together
                const unsigned long long* qp = qkv + hq * 32 + dch * 4;
                const unsigned long long* kp2 = qkv + q_words + kvh * 32 + dch * 4;
                const unsigned long long* vp2 = qkv + k_end_words + kvh * 32 + dch * 4;
                const uint32_t e = epoch - 1;
                const bool need_q = !have_q;
                uint4 q0 = make_uint4(0u, e, 0u, e), q1 = q0, k0 = q0, k1 = q0, v0 = q0, v1 = q0;
                if (need_q) { q0 = ld_relaxed_v4(qp); q1 = ld_relaxed_v4(qp + 2); }
                if (newest) { k0 = ld_relaxed_v4(kp2); k1 = ld_relaxed_v4(kp2 + 2); v0 = ld_relaxed_v4(vp2); v1 = ld_relaxed_v4(vp2 + 2); }
                uint32_t spins = 0;
                for (;;) {
                    bool ready = true;
                    if (q0.y != e || q0.w != e) { ready = false; q0 = ld_relaxed_v4(qp); }
                    if (q1.y != e || q1.w != e) { ready = false; q1 = ld_relaxed_v4(qp + 2); }
                    if (k0.y != e || k0.w != e) { ready = false; k0 = ld_relaxed_v4(kp2); }
                    if (k1.y != e || k1.w != e) { ready = false; k1 = ld_relaxed_v4(kp2 + 2); }
                    if (v0.y != e || v0.w != e) { ready = false; v0 = ld_relaxed_v4(vp2); }
                    if (v1.y != e || v1.w != e) { ready = false; v1 = ld_relaxed_v4(vp2 + 2); }
                    if (ready) break;
                    LL_SPIN_GUARD(spins);
                }
                if (need_q) {
                    unpack8(make_uint4(q0.x, q0.z, q1.x, q1.z), qf);
#pragma unroll
                    for (int i = 0; i < 8; ++i) qf[i] *= 0.125f;  // 1/sqrt(64), exact
                    have_q = true;
                }
                if (newest) { kk = make_uint4(k0.x, k0.z, k1.x, k1.z); vv = make_uint4(v0.x, v0.z, v1.x, v1.z); }
            }
            float kf[8], vf[8];
            unpack8(kk, kf);
            unpack8(vv, vf);
            float sc = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) sc = fmaf(qf[e], kf[e], sc);
            sc += __shfl_xor_sync(0xffffffffu, sc, 1);
            sc += __shfl_xor_sync(0xffffffffu, sc, 2);
            sc += __shfl_xor_sync(0xffffffffu, sc, 4);
            if (valid) {
                const float mn = fmaxf(m, sc);
                const float corr = (m == -INFINITY) ? 0.f : expf(m - mn);
                const float pe = expf(sc - mn);
                l = fmaf(l, corr, pe);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(acc[e], corr, pe * vf[e]);
                m = mn;
            }
        }
        // merge the four position groups of the warp
#pragma unroll 1
        for (int o = 8; o <= 16; o <<= 1) {
            const float mo = __shfl_xor_sync(0xffffffffu, m, o);
            const float lo = __shfl_xor_sync(0xffffffffu, l, o);
            const float mn = fmaxf(m, mo);
            const float c1 = (m == -INFINITY) ? 0.f : expf(m - mn);
            const float c2 = (mo == -INFINITY) ? 0.f : expf(mo - mn);
            l = fmaf(l, c1, __fmul_rn(lo, c2));  // explicit: both kernels must contract identically
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float ao = __shfl_xor_sync(0xffffffffu, acc[e], o);
                acc[e] = fmaf(acc[e], c1, __fmul_rn(ao, c2));
            }
            m = mn;
        }
        if (psub == 0) {
            const uint32_t dst = scratch + (uint32_t)(vw * kPartialStride) * 4u;
            if (dch == 0) { sts_f32(dst, m); sts_f32(dst + 4u, l); }
#pragma unroll
            for (int e = 0; e < 8; ++e) sts_f32(dst + (uint32_t)(2 + dch * 8 + e) * 4u, acc[e]);
        }
        }  // warp slots
        csync();
        // merge the warps (fixed order)
        if (tid < kHeadDim) {
            const int d = tid;
            float Mx = -INFINITY, Ls = 0.f, Os = 0.f;
#pragma unroll 1
            for (int wi = 0; wi < kWarps; ++wi) Mx = fmaxf(Mx, lds_f32(scratch + (uint32_t)(wi * kPartialStride) * 4u));
#pragma unroll 1
            for (int wi = 0; wi < kWarps; ++wi) {
                const uint32_t src = scratch + (uint32_t)(wi * kPartialStride) * 4u;
                const float mw = lds_f32(src);
                if (mw == -INFINITY) continue;
                const float sc = expf(mw - Mx);
                Ls = fmaf(lds_f32(src + 4u), sc, Ls);
                Os = fmaf(lds_f32(src + (uint32_t)(2 + d) * 4u), sc, Os);
            }
            if (ns == 1) {
                const float o = bf16_round(Os / Ls);
                const float o_next = __shfl_down_sync(0xffffffffu, o, 1);
                if (!(d & 1)) ll_put(ll_dst(M, p) + (hq * kHeadDim + d) / 2, len_out, kLLRep, pack_bf16(o, o_next), epoch);
            } else {
                unsigned long long* dst = partial_words(M, layer, hq, s);
                if (d == 0) {
                    st_relaxed_v2(dst, __float_as_uint(Mx), epoch);
                    st_relaxed_v2(dst + 1, __float_as_uint(Ls), epoch);
                }
                st_relaxed_v2(dst + 2 + d, __float_as_uint(Os), epoch);
            }
        }
        csync();
    }
    if (ns == 1) return;
    // combine the splits of one head in split order: CTAs after the attention units take one head each.
    // All of the head's partial words are polled in parallel into shared memory first (one L2 round trip).
    for (int hq = 0; hq < Hq; ++hq) {
        if ((n_units + hq) % n_ctas != cta) continue;
        {   // all of the thread's words in flight before the first epoch is looked at (32 * 66 words / 352 threads = 6)
            constexpr int kPer = (kMaxSplits * kPartialStride + kCons - 1) / kCons;
            uint2 v[kPer];
            const int n_words = ns * kPartialStride;
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int i = tid + k * kCons;
                v[k] = make_uint2(0u, epoch);
                if (i < n_words) v[k] = ld_relaxed_v2(partial_words(M, layer, hq, i / kPartialStride) + i % kPartialStride);
            }
            uint32_t spins = 0;
            for (;;) {
                bool ready = true;
#pragma unroll
                for (int k = 0; k < kPer; ++k) {
                    if (v[k].y != epoch) {
                        const int i = tid + k * kCons;
                        ready = false;
                        v[k] = ld_relaxed_v2(partial_words(M, layer, hq, i / kPartialStride) + i % kPartialStride);
                    }
                }
                if (ready) break;
                LL_SPIN_GUARD(spins);
            }
#pragma unroll
            for (int k = 0; k < kPer; ++k)
                if (tid + k * kCons < n_words) sts_u32(scratch + (uint32_t)(tid + k * kCons) * 4u, v[k].x);
        }
        csync();
        if (tid < kHeadDim) {
            const int d = tid;
            float Mg = -INFINITY, Lg = 0.f, Og = 0.f;
#pragma unroll 1
            for (int si = 0; si < ns; ++si) Mg = fmaxf(Mg, lds_f32(scratch + (uint32_t)(si * kPartialStride) * 4u));
#pragma unroll 1
            for (int si = 0; si < ns; ++si) {
                const uint32_t src = scratch + (uint32_t)(si * kPartialStride) * 4u;
                const float ms = lds_f32(src);
                if (ms == -INFINITY) continue;
                const float sc = expf(ms - Mg);
                Lg = fmaf(lds_f32(src + 4u), sc, Lg);
                Og = fmaf(lds_f32(src + (uint32_t)(2 + d) * 4u), sc, Og);
            }
            const float o = bf16_round(Og / Lg);
            const float o_next = __shfl_down_sync(0xffffffffu, o, 1);
            if (!(d & 1)) ll_put(ll_dst(M, p) + (hq * kHeadDim + d) / 2, len_out, kLLRep, pack_bf16(o, o_next), epoch);
        }
        csync();
    }
}

// ---- sampling (G:88-99 slow, G:118-132 depth) + frame assembly and the stop rule (G:143-166) ------------
__device__ __noinline__ void phase_sample(const DevModel& M, const CallArgs& A, int is_fast, int depth_pos, int p, uint32_t epoch,
                                          float* lg) {
    const int tid = threadIdx.x, n_ctas = gridDim.x;
    const bool fast = is_fast != 0;
    const int N = fast ? M.codebook_size : M.vocab;
    const int r = fast ? 1 + depth_pos : 0;
    const int R = M.n_rows;
    const bool local = row_is_local(M, A, r);
    if (!local && blockIdx.x != 0) return;  // sampled rows: CTA 0 is the sampler, the others pick the id up when they need it
    int tok;
    if (M.force != nullptr) {
        tok = ldcg_i32(M.force + r);
    } else if (local) {
        // greedy: every CTA published its best (logit, index) with the HEAD phase's epoch; warp 0 reduces them
        if (tid < 32) {
            const unsigned long long* src = M.ll_cand + ((size_t)r * kLLRep + (blockIdx.x % kLLRep)) * kLLMaxCtas;
            // all of the lane's words in flight before the first epoch is looked at (kLLMaxCtas / 32 = 8 per lane)
            uint2 v[kLLMaxCtas / 32];
#pragma unroll
            for (int k = 0; k < kLLMaxCtas / 32; ++k) {
                v[k] = make_uint2(0u, epoch - 1);
                if (tid + 32 * k < n_ctas) v[k] = ld_relaxed_v2(src + tid + 32 * k);
            }
            uint32_t spins = 0;
            for (;;) {
                bool ready = true;
#pragma unroll
                for (int k = 0; k < kLLMaxCtas / 32; ++k)
                    if (v[k].y != epoch - 1) { ready = false; v[k] = ld_relaxed_v2(src + tid + 32 * k); }
                if (ready) break;
                LL_SPIN_GUARD(spins);
            }
            uint32_t best = 0u;
#pragma unroll
            for (int k = 0; k < kLLMaxCtas / 32; ++k) best = v[k].x > best ? v[k].x : best;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const uint32_t ob = __shfl_xor_sync(0xffffffffu, best, o);
                best = ob > best ? ob : best;
            }
            if (tid == 0) s_cand[0] = best;
        }
        csync();
        tok = cand_index(s_cand[0]);
    } else {
        const unsigned long long* src = ll_src(M, p - 1, 0);  // the HEAD phase's logits (replica 0)
        for (int i = tid; i < N / 4; i += kCons) {
            uint4 v = ld_relaxed_v4(src + 2 * i);
            uint32_t spins = 0;
            while (v.y != epoch - 1 || v.w != epoch - 1) { LL_SPIN_GUARD(spins); v = ld_relaxed_v4(src + 2 * i); }
            *reinterpret_cast<float4*>(lg + 4 * i) = make_float4(bf_lo(v.x), bf_hi(v.x), bf_lo(v.z), bf_hi(v.z));
        }
        csync();
        const float temp = fast ? A.s.fast_temp : A.s.temp;
        if (temp == 0.0f) {
            tok = block_argmax<kCons>(lg, N, s_sc);
        } else {
            const uint32_t seq_id = A.b.seq_id ? (uint32_t)ldcg_i32(A.b.seq_id) : 0u;
            tok = sample_row<kCons>(lg, N, temp, fast ? 0 : A.s.top_k, fast ? 1.0f : A.s.top_p, A.s.min_p, A.s.seed,
                                    (uint32_t)s_step, seq_id, (uint32_t)r, s_sc);
        }
    }
    const bool last = fast && depth_pos == M.depth - 1;
    if (local) {
        if (tid == 0) s_nw[r] = tok;
    } else {
        if (last) __threadfence();  // frame boundary: keep the release chain cumulative
        if (tid < n_ctas) st_relaxed_v2(M.ll_tok + (size_t)r * kLLMaxCtas + tid, (uint32_t)tok, epoch);
    }
    if (blockIdx.x == 0 && tid == 0) {
        M.frame_tokens[r] = tok;
        if (last && s_fin == 0) {
            // frame assembly for the host and the next launch; every CTA applies the same update locally
            const int st = s_step;
            int slow = 0;
            for (int rr = 0; rr < R; ++rr) {
                int v;
                if (rr == r) v = tok;
                else if (rr == 0 && !row_is_local(M, A, 0)) v = token_word_get(M, 0, epoch - (uint32_t)(p - sample_phase_index(M, 0)));
                else v = s_nw[rr];
                if (rr == 0) slow = v;
                A.b.tokens[rr] = v;
                if (A.b.out_codes != nullptr && st < A.b.max_frames) A.b.out_codes[(size_t)st * R + rr] = v;
            }
            if (A.b.step) A.b.step[0] = st + 1;
            if (M.frame_ns != nullptr && st < M.frame_ns_cap) M.frame_ns[st] = globaltimer_ns();
            A.b.seq_len[0] = s_pos + 1;
            if (A.b.finished != nullptr && A.s.audio_only && !A.s.ignore_stop && slow == M.im_end) A.b.finished[0] = 1;
        }
    }
    csync();
}

// ---- weight rows x one activation row, all in shared memory ----------------------------------------------
// acc_r = sum_k W_r[k] * x[k]: lane-strided 8-element chunks, two accumulator chains per row (elements
// 0-3 / 4-7 of a chunk), butterfly warp sum -- the summation order of decode_kernel.cu: gemv_units.
template <int R>
__device__ __forceinline__ void gemv_rows(const unsigned char* sm, const uint32_t (&w)[R], uint32_t xs, int c0, int nchunks, int lane, float (&r)[R]) {
    // plain shared-memory loads through the kernel's own array (the compiler may batch them); kT chunks (lane, lane+32, ..)
    // per trip so that their loads are in flight before the first FMA -- chunks are still accumulated in ascending order
    constexpr int kT = R <= 2 ? 3 : 1;
    float a0[R], a1[R];
#pragma unroll
    for (int i = 0; i < R; ++i) { a0[i] = 0.f; a1[i] = 0.f; }
#pragma unroll 1
    for (int ch = c0 + lane; ch < nchunks; ch += 32 * kT) {
        uint4 xv[kT], wv[R][kT];
#pragma unroll
        for (int t = 0; t < kT; ++t) {
            const int cc = min(ch + 32 * t, nchunks - 1);  // clamp: a chunk past the end is loaded but not used
            xv[t] = *reinterpret_cast<const uint4*>(sm + xs + (uint32_t)cc * 16u);
#pragma unroll
            for (int i = 0; i < R; ++i) wv[i][t] = *reinterpret_cast<const uint4*>(sm + w[i] + (uint32_t)cc * 16u);
        }
#pragma unroll
        for (int t = 0; t < kT; ++t) {
            if (ch + 32 * t < nchunks) {
                float x[8];
                unpack8(xv[t], x);
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    float f[8];
                    unpack8(wv[i][t], f);
                    a0[i] = fmaf(f[0], x[0], a0[i]); a1[i] = fmaf(f[4], x[4], a1[i]);
                    a0[i] = fmaf(f[1], x[1], a0[i]); a1[i] = fmaf(f[5], x[5], a1[i]);
                    a0[i] = fmaf(f[2], x[2], a0[i]); a1[i] = fmaf(f[6], x[6], a1[i]);
                    a0[i] = fmaf(f[3], x[3], a0[i]); a1[i] = fmaf(f[7], x[7], a1[i]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) r[i] = a0[i] + a1[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < R; ++i) r[i] += __shfl_xor_sync(0xffffffffu, r[i], o);
    }
}

// ---- per-CTA phase descriptors ------------------------------------------------------------------------
// The schedule of a frame is static: everything a phase needs that does not depend on data (pointers into the
// LL regions for this CTA's replica, the CTA's unit range, sizes, flags) is worked out ONCE per launch into
// shared memory.  The per-phase critical path is one warp's dependent instruction chain; pointer arithmetic,
// integer divisions and constant-bank loads do not belong on it.
//   word  0-1  src      LL source region (this CTA's replica) | fast_embeddings row base
//         2-3  normw    RMSNorm weight
//         4-5  out      first output word of this CTA
//         6-7  aux      QKV: RoPE table (row base) | HEAD: fp32 logits dump
//         8    e_back   source epoch = epoch - e_back
//         9    bytes    bytes of one ring stage
//        10    len_out | nch << 16
//        11    u0 | nu << 16
//        12    kind | fast << 4 | src_kind << 8 | nrep << 12 (6 bits) | n_parts << 18 | res_sel << 20 | layer << 24
//        13    depth_pos | stage ordinal << 8 | activation buffer << 24
//        14    ring offset of the stage (gated MLP: of the w1 rows)      15    ring offset of the w3 rows
// The ring is laid out identically in every iteration (the allocator restarts at offset 0), so that the only
// loop-carried state of a consumer thread is (iteration, phase): everything else is read from here.

__device__ __forceinline__ void build_desc(const DevModel& M, const CallArgs& A, int p, uint32_t dsc) {
    const int cta = blockIdx.x, n_ctas = gridDim.x, rep = cta % kLLRep;
    const Phase ph = unpack_phase(M.prog[p]);
    const bool fast = ph.fast != 0;
    const int kind = ph.kind;
    const int n_slow = 5 * M.n_layer;
    uint32_t w[kDescWords];
#pragma unroll
    for (int i = 0; i < kDescWords; ++i) w[i] = 0u;
    int src_kind = 3, nrep = kLLRep, n_parts = 1, nu = 0, u0 = 0, nch = 0;
    if (kind != PH_ATTN && kind != PH_SAMPLE) {
        const Plan pl = phase_plan(M, ph);
        const int D = fast ? M.fdim : M.dim;
        u0 = unit_begin(pl.U, cta, n_ctas);
        nu = unit_begin(pl.U, cta + 1, n_ctas) - u0;
        nch = pl.K >> 3;
        n_parts = pl.w1 ? 2 : 1;
        int p_src = p - 1;
        unsigned long long src = 0ull;
        if (kind == PH_QKV && ph.layer == 0) {
            if (!fast) src_kind = 0;
            else if (ph.depth_pos == 0) p_src = n_slow - 1;  // the slow transformer's pre-norm hidden state (P:259, M:191)
            else {
                src_kind = 2;  // embedding of the previous depth code (G:136-140)
                const int off = M.depthwise_wte ? (M.dup0 ? ph.depth_pos - 1 : ph.depth_pos) * M.codebook_size : 0;
                src = (unsigned long long)(M.fast_embeddings + (size_t)off * D);
            }
        }
        if (src_kind == 3) src = (unsigned long long)ll_src(M, p_src, rep);
        const DevLayer& L = fast ? M.fast_layers[ph.layer] : M.layers[ph.layer];
        const uint16_t* nw = kind == PH_QKV ? L.attention_norm : kind == PH_W13 ? L.ffn_norm : kind == PH_HEAD ? (fast ? M.fast_norm : M.norm) : nullptr;
        const unsigned long long out = (unsigned long long)(ll_dst(M, p) + u0);
        unsigned long long aux = 0ull;
        if (kind == PH_QKV) aux = (unsigned long long)(fast ? M.fast_rope + (size_t)ph.depth_pos * kHeadDim : M.rope);
        if (kind == PH_HEAD) {
            aux = (unsigned long long)(fast ? M.depth_logits + (size_t)ph.depth_pos * M.codebook_size : M.token_logits);
            nrep = 1;  // logits words are read by the sampler CTA only (greedy rows use the candidate words)
        }
        w[0] = (uint32_t)src; w[1] = (uint32_t)(src >> 32);
        w[2] = (uint32_t)(unsigned long long)nw; w[3] = (uint32_t)((unsigned long long)nw >> 32);
        w[4] = (uint32_t)out; w[5] = (uint32_t)(out >> 32);
        w[6] = (uint32_t)aux; w[7] = (uint32_t)(aux >> 32);
        w[8] = (uint32_t)(p - p_src);
        w[9] = (uint32_t)nu * 4u * (uint32_t)pl.K;
    }
    w[10] = (uint32_t)M.ll_len[p] | ((uint32_t)nch << 16);
    w[11] = (uint32_t)u0 | ((uint32_t)nu << 16);
    const int res_sel = (kind == PH_QKV || kind == PH_WO) ? 0 : 1;
    w[12] = (uint32_t)kind | ((uint32_t)(fast ? 1 : 0) << 4) | ((uint32_t)src_kind << 8) | ((uint32_t)nrep << 12) |
            ((uint32_t)n_parts << 18) | ((uint32_t)res_sel << 20) | ((uint32_t)ph.layer << 24);
    w[13] = (uint32_t)ph.depth_pos;
#pragma unroll
    for (int i = 0; i < kDescWords / 4; ++i) sts_v4(dsc + 16u * i, make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]));
    (void)A;
}

// RMSNorm (P:601-613) of one row by ONE warp, lane l owning the chunks l, l+32, l+64 (raw[i], already loaded):
// same arithmetic and summation order as decode_kernel.cu: norm_rows (chunk sums, butterfly per 32 chunks,
// then the partial sums of the <= 3 chunk groups in order).  Writes the raw row to `res` (if nonzero) and the
// normalised row to xs.
__device__ __forceinline__ void norm_store(const uint4 (&raw)[3], const uint4 (&wv)[3], int nch, int lane, float eps,
                                           uint32_t xs, uint32_t res) {
    float ss[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float x[8];
        unpack8(raw[i], x);
        ss[i] = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) ss[i] = fmaf(x[e], x[e], ss[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < 3; ++i) ss[i] += __shfl_xor_sync(0xffffffffu, ss[i], o);
    }
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (32 * i < nch) t += ss[i];
    const float mean = __fdiv_rn(t, (float)(nch * 8));
    const float rr = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, eps)));
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (lane + 32 * i < nch) {
            if (res) sts_v4(res + (uint32_t)(lane + 32 * i) * 16u, raw[i]);
            float x[8], wf[8], o[8];
            unpack8(raw[i], x);
            unpack8(wv[i], wf);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = bf16_round(__fmul_rn(bf16_round(__fmul_rn(x[e], rr)), wf[e]));
            sts_v4(xs + (uint32_t)(lane + 32 * i) * 16u, pack8(o));
        }
    }
}

// The two prologues whose input is not an LL region (once per frame / once per depth step, warp 0 only):
// the token embedding (slow layer 0, P:205-221) and the embedding of the previous depth code (G:136-140).
__device__ __noinline__ void prologue_embed(const DevModel& M, const CallArgs& A, int it, int p, uint32_t epoch, int src_kind,
                                            int depth_pos, const uint16_t* emb_base, const uint16_t* normw, int nch,
                                            uint32_t xs, uint32_t res) {
    const int lane = threadIdx.x & 31;
    uint4 raw[3], wv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        raw[i] = make_uint4(0u, 0u, 0u, 0u); wv[i] = raw[i];
        if (lane + 32 * i < nch) wv[i] = __ldg(reinterpret_cast<const uint4*>(normw + (lane + 32 * i) * 8));
    }
    if (src_kind == 2) {
        const int r = depth_pos;  // row of depth code depth_pos-1
        int code;
        if (row_is_local(M, A, r)) code = s_nw[r];
        else {
            code = token_word_get(M, r, epoch - (uint32_t)(p - sample_phase_index(M, r)));
            if (lane == 0) s_nw[r] = code;
        }
        const uint16_t* row = emb_base + (size_t)code * (nch * 8);
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (lane + 32 * i < nch) raw[i] = ldcg_v4(row + (lane + 32 * i) * 8);
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (lane + 32 * i < nch) raw[i] = embed_chunk(M, A, it, lane + 32 * i);
    }
    norm_store(raw, wv, nch, lane, M.eps, xs, res);
}

// KV append (K:12-22) of one bf16 pair into the paged pool (slow QKV phases, lane 0 of the k / v units)
__device__ __noinline__ void kv_append(const DevModel& M, const CallArgs& A, int it, int layer, int n0, uint32_t word) {
    const int Hq = M.n_head, Hkv = M.n_kv, ps = M.page_size;
    const int q_rows = Hq * kHeadDim, k_end = (Hq + Hkv) * kHeadDim;
    const int pos = s_pos;
    if (pos >= A.b.max_pages * ps) return;
    const bool active = (A.mode == 1) ? (it + A.iter_base < ldcg_i32(A.prompt_len) - 1) : (s_fin == 0);
    if (!active) return;
    const int pg = pos / ps;
    const int page = pg < kBtabCache ? s_btab[pg] : ldcg_i32(A.b.block_table + pg);
    const int is_v = n0 >= k_end ? 1 : 0;
    const int n1 = n0 - (is_v ? k_end : q_rows);
    const int kvh = n1 / kHeadDim, d = n1 & (kHeadDim - 1);
    uint16_t* dst = M.kv_pool + ((((size_t)page * M.n_layer + layer) * 2 + is_v) * Hkv + kvh) * ((size_t)ps * kHeadDim)
                    + (size_t)(pos % ps) * kHeadDim + d;
    *reinterpret_cast<uint32_t*>(dst) = word;
}

template <bool kTrace>
__global__ void __launch_bounds__(kLLThreads, 1)
smol_ll_kernel(const __grid_constant__ DevModel M, const __grid_constant__ CallArgs A, const __grid_constant__ SmemPlan SP) {
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sm0 = smem_u32(smem_dyn);
    const uint32_t full0 = smem_u32(s_full), empty0 = smem_u32(s_empty);
    const int per_iter = A.phase_end;
#define LL_XS0 (sm0)
#define LL_RES0 (sm0 + 2u * SP.xs_bytes)
#define LL_SCRATCH (sm0 + 2u * SP.xs_bytes + 2u * SP.res_bytes)
#define LL_DSC0 (LL_SCRATCH + SP.scratch_bytes)
#define LL_FQ (LL_DSC0 + SP.desc_bytes)
#define LL_FKV (LL_FQ + SP.fq_bytes)
#define LL_RING (LL_FKV + SP.fkv_bytes)

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(full0 + 8u * i, 1u); mbar_init(empty0 + 8u * i, (uint32_t)kLLWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_pos = ldcg_i32(A.b.seq_len);
        s_step = A.b.step ? ldcg_i32(A.b.step) : 0;
        s_fin = (A.mode == 0 && A.b.finished != nullptr) ? (int)__ldcg(A.b.finished) : 0;
    }
    if (A.mode == 0 && tid < M.n_rows) { s_tok[tid] = ldcg_i32(A.b.tokens + tid); s_nw[tid] = 0; }
    for (int p = tid; p < per_iter; p += kLLThreads) build_desc(M, A, p, LL_DSC0 + (uint32_t)p * (kDescWords * 4u));
    for (int i = tid; i < kBtabCache && i < A.b.max_pages; i += kLLThreads) s_btab[i] = ldcg_i32(A.b.block_table + i);
    __syncthreads();
    if (tid == 0) {  // sequential pass: ring layout of one iteration, stage ordinals, activation buffer parity
        uint32_t head = 0, stage = 0, wph = 0;
        for (int p = 0; p < per_iter; ++p) {
            const uint32_t dsc = LL_DSC0 + (uint32_t)p * (kDescWords * 4u);
            const uint32_t flags = lds_u32(dsc + 48u), bytes = lds_u32(dsc + 36u);
            const uint32_t kind = flags & 15u, n_parts = (flags >> 18) & 3u;
            if (kind == PH_ATTN || kind == PH_SAMPLE) continue;
            sts_u32(dsc + 52u, lds_u32(dsc + 52u) | (stage << 8) | ((wph & 1u) << 24));
            sts_u32(dsc + 56u, ring_alloc(head, bytes, (uint32_t)SP.ring_bytes));
            if (n_parts == 2) sts_u32(dsc + 60u, ring_alloc(head, bytes, (uint32_t)SP.ring_bytes));
            stage += n_parts;
            wph += 1;
        }
        s_spf = (int)stage;
        s_wodd = (int)(wph & 1u);
        s_epoch0 = (uint32_t)ldcg_i32(M.ll_epoch);
    }
    __syncthreads();

    if (warp == kLLWarps) {  // the TMA producer warp
        if (lane == 0) producer(M, A, LL_DSC0, LL_RING, (uint32_t)SP.ring_bytes);
        return;
    }

    // cycle trace of the launch's last iteration (kTrace): CTA 0 and CTA n/2; thread 0 stamps 0 start, 1 input words
    // arrived, 7 input row staged, 2 past the block barrier, 6 end; lane 0 of warp 1 (first GEMV unit of the CTA) stamps
    // 3 GEMV start, 4 GEMV done, 5 result published
    unsigned long long* tr = nullptr;
    if (kTrace && M.prof != nullptr && (tid == 0 || tid == 32) && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2))
        tr = M.prof + 2 * kMaxProg + 64 + (blockIdx.x == 0 ? 0 : 1) * kMaxProg * 8;
    // skew trace (kTrace): EVERY CTA, thread 0, %globaltimer when its input words have arrived and at the end of the phase
    unsigned long long* trg = nullptr;
    if (kTrace && M.prof != nullptr && tid == 0) trg = M.prof + 2 * kMaxProg + 64 + 2 * kMaxProg * 8 + (size_t)blockIdx.x * (kMaxProg * 2);
#define LL_SKEW(i) do { if (kTrace && trg) trg[p * 2 + (i)] = globaltimer_ns(); } while (0)
#define LL_TRACE(i) do { if (kTrace && tr && tid == 0) { unsigned long long c_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_) :: "memory"); tr[p * 8 + (i)] = c_; } } while (0)
#define LL_TRACE_AUX(i) do { if (kTrace && tr && tid != 0) { unsigned long long c_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_) :: "memory"); tr[p * 8 + (i)] = c_; } } while (0)
    // the only loop-carried state of a consumer thread is (it, p): everything else comes from the descriptors
    for (int it = 0; it < A.n_iter; ++it) {
#pragma unroll 1
        for (int p = 0; p < per_iter; ++p) {
            const uint32_t dsc = LL_DSC0 + (uint32_t)p * (kDescWords * 4u);
            const uint32_t epoch = s_epoch0 + (uint32_t)(it * per_iter + p) + 1u;
            const uint4 d3 = lds_v4(dsc + 48u);
            const int kind = (int)(d3.x & 15u);
            const bool fast = (d3.x >> 4) & 1u;
            const int layer = (int)(d3.x >> 24), depth_pos = (int)(d3.y & 255u);
            LL_TRACE(0);
            if (kind == PH_ATTN) {
                phase_attn(M, A, layer, p, epoch, LL_SCRATCH);
                LL_TRACE(6);
                continue;
            }
            if (kind == PH_SAMPLE) {
                phase_sample(M, A, fast ? 1 : 0, depth_pos, p, epoch,
                             reinterpret_cast<float*>(smem_dyn + 2 * SP.xs_bytes + 2 * SP.res_bytes));
                LL_TRACE(6);
                continue;
            }
            // ================= a phase that streams weights =================
            const uint4 d2 = lds_v4(dsc + 32u);
            const uint32_t bytes = d2.y;
            const int nch = (int)(d2.z >> 16);
            const int n_parts = (int)((d3.x >> 18) & 3u);
            const uint32_t xcur = ((d3.y >> 24) ^ (uint32_t)(it & s_wodd)) & 1u;
            const uint32_t xs = LL_XS0 + xcur * (uint32_t)SP.xs_bytes;
            const uint32_t res = LL_RES0 + ((d3.x >> 20) & 1u) * (uint32_t)SP.res_bytes;
            const uint32_t wbase = LL_RING + d3.z, wbase2 = LL_RING + d3.w;
            const uint32_t wseq = (uint32_t)(it * s_spf) + ((d3.y >> 8) & 0xffffu);
            (void)bytes;

            // ---- prologue: the phase's input row -> xs -----------------------------------------------------
            {
                const uint4 d0 = lds_v4(dsc);
                const unsigned long long* src = reinterpret_cast<const unsigned long long*>((unsigned long long)d0.x | ((unsigned long long)d0.y << 32));
                const uint16_t* normw = reinterpret_cast<const uint16_t*>((unsigned long long)d0.z | ((unsigned long long)d0.w << 32));
                const uint32_t e_src = epoch - d2.x;
                const int src_kind = (int)((d3.x >> 8) & 15u);
                if (src_kind != 3) {
                    if (src_kind == 0 && A.mode == 0 && it > 0) frame_boundary(M, A, epoch - (uint32_t)per_iter);
                    if (warp == 0)
                        prologue_embed(M, A, it, p, epoch, src_kind, depth_pos, reinterpret_cast<const uint16_t*>(src), normw, nch, xs, res);
                    LL_TRACE(1);
                } else if (kind == PH_WO && fast) {
                    fast_attention(M, layer, depth_pos, src, e_src, LL_FQ, LL_FKV, xs);
                    LL_TRACE(1);
                } else if (kind == PH_WO || kind == PH_W2) {
                    // up to two chunks per thread, all loads in flight before the first epoch is looked at; a retry
                    // re-reads only the words that have not arrived
                    const int c1 = tid + kCons;
                    if (tid < nch) {
                        const bool two = c1 < nch;
                        const unsigned long long* s0 = src + tid * 4;
                        const unsigned long long* s1 = src + (two ? c1 : tid) * 4;
                        uint4 a0 = ld_relaxed_v4(s0), b0 = ld_relaxed_v4(s0 + 2), a1 = a0, b1 = b0;
                        if (two) { a1 = ld_relaxed_v4(s1); b1 = ld_relaxed_v4(s1 + 2); }
                        uint32_t spins = 0;
                        for (;;) {
                            const bool r0 = a0.y == e_src && a0.w == e_src, r1 = b0.y == e_src && b0.w == e_src;
                            const bool r2 = a1.y == e_src && a1.w == e_src, r3 = b1.y == e_src && b1.w == e_src;
                            if (r0 && r1 && r2 && r3) break;
                            LL_SPIN_GUARD(spins);
                            if (!r0) a0 = ld_relaxed_v4(s0);
                            if (!r1) b0 = ld_relaxed_v4(s0 + 2);
                            if (!r2) a1 = ld_relaxed_v4(s1);
                            if (!r3) b1 = ld_relaxed_v4(s1 + 2);
                        }
                        sts_v4(xs + (uint32_t)tid * 16u, make_uint4(a0.x, a0.z, b0.x, b0.z));
                        if (two) sts_v4(xs + (uint32_t)c1 * 16u, make_uint4(a1.x, a1.z, b1.x, b1.z));
                    }
                    LL_TRACE(1);
                } else if (warp == 0) {
                    uint4 raw[3], wv[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        raw[i] = make_uint4(0u, 0u, 0u, 0u); wv[i] = raw[i];
                        if (lane + 32 * i < nch) wv[i] = __ldg(reinterpret_cast<const uint4*>(normw + (lane + 32 * i) * 8));
                    }
                    // every load is issued before the first epoch is looked at; a retry re-reads only what is missing
                    uint4 a[3], b[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        a[i] = make_uint4(0u, e_src, 0u, e_src); b[i] = a[i];
                        if (lane + 32 * i < nch) {
                            a[i] = ld_relaxed_v4(src + (lane + 32 * i) * 4);
                            b[i] = ld_relaxed_v4(src + (lane + 32 * i) * 4 + 2);
                        }
                    }
                    uint32_t spins = 0;
                    for (;;) {
                        bool ready = true;
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            if (a[i].y != e_src || a[i].w != e_src) { ready = false; a[i] = ld_relaxed_v4(src + (lane + 32 * i) * 4); }
                            if (b[i].y != e_src || b[i].w != e_src) { ready = false; b[i] = ld_relaxed_v4(src + (lane + 32 * i) * 4 + 2); }
                        }
                        if (ready) break;
                        LL_SPIN_GUARD(spins);
                    }
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        if (lane + 32 * i < nch) raw[i] = make_uint4(a[i].x, a[i].z, b[i].x, b[i].z);
                    LL_TRACE(1);
                    LL_SKEW(0);
                    norm_store(raw, wv, nch, lane, M.eps, xs, kind != PH_HEAD ? res : 0u);
                }
            }
            // Everything the GEMV of this warp's first unit needs that does not depend on the staged row is worked out
            // BEFORE the block barrier (the warps that do not stage the row are idle here): descriptor fields, the unit
            // index, the RoPE pair / the residual pair (written two barriers ago).
            const uint4 d1 = lds_v4(dsc + 16u);
            unsigned long long* out = reinterpret_cast<unsigned long long*>((unsigned long long)d1.x | ((unsigned long long)d1.y << 32));
            const unsigned long long aux = (unsigned long long)d1.z | ((unsigned long long)d1.w << 32);
            const int len_out = (int)(d2.z & 0xffffu);
            const int u0 = (int)(d2.w & 0xffffu), nu = (int)(d2.w >> 16);
            const int nrep = (int)((d3.x >> 12) & 63u);
            const uint32_t row_bytes = (uint32_t)nch * 16u;
            const int Hq = fast ? M.fn_head : M.n_head, Hkv = fast ? M.fn_kv : M.n_kv;
            const int q_rows = Hq * kHeadDim, k_end = (Hq + Hkv) * kHeadDim;
            // units go to warps 1, 2, .., 10, 0: the norm warp, which reaches the barrier last, gets a unit only when
            // every other warp has one  (option "ll_flags" bit 3 switches the rotation off)
            const int jw = (A.repeat & 8) ? warp : (warp + kLLWarps - 1) % kLLWarps;
            uint32_t pre_first = 0u;
            if (jw < nu) {
                const int n0 = 2 * (u0 + jw);
                if (kind == PH_QKV) {
                    if (n0 < k_end) {
                        const uint16_t* table = reinterpret_cast<const uint16_t*>(aux) + (fast ? 0 : (size_t)min(s_pos, M.max_seq_len - 1) * kHeadDim);
                        pre_first = __ldg(reinterpret_cast<const uint32_t*>(table + (n0 & (kHeadDim - 1))));
                    }
                } else if (kind == PH_WO || kind == PH_W2) {
                    pre_first = lds_u32(res + (uint32_t)n0 * 2u);
                }
            }
            // the last consumer warp (never on the critical path of a prologue) makes sure the weights have landed;
            // the block barrier below passes that on to everyone
            if (warp == kLLWarps - 1) {
                for (int i = 0; i < n_parts; ++i) mbar_wait(full0 + 8u * ((wseq + i) % kStages), ((wseq + i) / kStages) & 1u);
            }
            LL_TRACE(7);
            csync();
            LL_TRACE(2);

            uint32_t warp_cand = 0u;  // HEAD phases: this warp's best (logit, index)
            {
                if (fast && kind == PH_HEAD && depth_pos == M.depth - 1) __threadfence();  // release side of the once-per-frame fence

                // long rows (w2): the K slices of 128 chunks of a unit go to different warps, partial sums meet in
                // shared memory and are added in slice order (the order decode_kernel.cu uses inside one warp)
                const int n_slices = (nch + 127) >> 7;
                const bool ksplit = !(A.repeat & 1);
                if (ksplit && n_slices > 1) {
#pragma unroll 1
                    for (int t = warp; t < nu * n_slices; t += kLLWarps) {
                        const int j = t / n_slices, sl = t - j * n_slices;
                        const uint32_t w[2] = {wbase - sm0 + (uint32_t)(2 * j) * row_bytes, wbase - sm0 + (uint32_t)(2 * j + 1) * row_bytes};
                        float r[2];
                        gemv_rows<2>(smem_dyn, w, xs - sm0, sl * 128, min(nch, sl * 128 + 128), lane, r);
                        if (lane == 0) { s_kpart[2 * t] = r[0]; s_kpart[2 * t + 1] = r[1]; }
                    }
                    csync();
                }
#pragma unroll 1
                for (int j = jw; j < nu; j += kLLWarps) {
                    const int n0 = 2 * (u0 + j);
                    // operands of the epilogue that do not depend on the GEMV: issue their loads first
                    uint32_t pre = pre_first;
                    if (j == jw) {
                        // loaded before the barrier
                    } else if (kind == PH_QKV) {
                        if (n0 < k_end) {
                            const uint16_t* table = reinterpret_cast<const uint16_t*>(aux) + (fast ? 0 : (size_t)min(s_pos, M.max_seq_len - 1) * kHeadDim);
                            pre = __ldg(reinterpret_cast<const uint32_t*>(table + (n0 & (kHeadDim - 1))));
                        }
                    } else if (kind == PH_WO || kind == PH_W2) {
                        pre = lds_u32(res + (uint32_t)n0 * 2u);
                    }
                    float a0, a1, g0 = 0.f, g1 = 0.f;
                    if (j == jw) LL_TRACE_AUX(3);
                    if (kind == PH_W13) {
                        const uint32_t w[4] = {wbase - sm0 + (uint32_t)(2 * j) * row_bytes, wbase - sm0 + (uint32_t)(2 * j + 1) * row_bytes,
                                               wbase2 - sm0 + (uint32_t)(2 * j) * row_bytes, wbase2 - sm0 + (uint32_t)(2 * j + 1) * row_bytes};
                        float r[4];
                        gemv_rows<4>(smem_dyn, w, xs - sm0, 0, nch, lane, r);
                        a0 = __fadd_rn(0.f, r[0]); a1 = __fadd_rn(0.f, r[1]); g0 = __fadd_rn(0.f, r[2]); g1 = __fadd_rn(0.f, r[3]);
                    } else if (ksplit && n_slices > 1) {
                        a0 = 0.f; a1 = 0.f;
                        for (int sl = 0; sl < n_slices; ++sl) {
                            a0 = __fadd_rn(a0, s_kpart[2 * (j * n_slices + sl)]);
                            a1 = __fadd_rn(a1, s_kpart[2 * (j * n_slices + sl) + 1]);
                        }
                    } else if (n_slices > 1) {
                        const uint32_t w[2] = {wbase - sm0 + (uint32_t)(2 * j) * row_bytes, wbase - sm0 + (uint32_t)(2 * j + 1) * row_bytes};
                        a0 = 0.f; a1 = 0.f;
#pragma unroll 1
                        for (int sl = 0; sl < n_slices; ++sl) {
                            float r[2];
                            gemv_rows<2>(smem_dyn, w, xs - sm0, sl * 128, min(nch, sl * 128 + 128), lane, r);
                            a0 = __fadd_rn(a0, r[0]); a1 = __fadd_rn(a1, r[1]);
                        }
                    } else {
                        const uint32_t w[2] = {wbase - sm0 + (uint32_t)(2 * j) * row_bytes, wbase - sm0 + (uint32_t)(2 * j + 1) * row_bytes};
                        float r[2];
                        gemv_rows<2>(smem_dyn, w, xs - sm0, 0, nch, lane, r);
                        a0 = __fadd_rn(0.f, r[0]); a1 = __fadd_rn(0.f, r[1]);
                    }
                    if (j == jw) LL_TRACE_AUX(4);
                    uint32_t word = 0u;
                    if (lane != 0) {
                        // lanes 1.. only help publishing
                    } else if (kind == PH_QKV) {
                        float v0 = bf16_round(a0), v1 = bf16_round(a1);
                        if (n0 < k_end) {  // q and k rows: interleaved-pair RoPE with the bf16 table (P:616-640)
                            const float co = bf_lo(pre), si = bf_hi(pre);
                            const float r0 = bf16_round(__fsub_rn(__fmul_rn(v0, co), __fmul_rn(v1, si)));
                            const float r1 = bf16_round(__fadd_rn(__fmul_rn(v1, co), __fmul_rn(v0, si)));
                            v0 = r0; v1 = r1;
                        }
                        word = pack_bf16(v0, v1);
                    } else if (kind == PH_W13) {
                        const float x0 = bf16_round(a0), x1 = bf16_round(a1);
                        const float s0 = bf16_round(__fdiv_rn(x0, __fadd_rn(1.0f, expf(-x0))));  // F.silu in fp32, bf16 out
                        const float s1 = bf16_round(__fdiv_rn(x1, __fadd_rn(1.0f, expf(-x1))));
                        word = pack_bf16(__fmul_rn(s0, bf16_round(g0)), __fmul_rn(s1, bf16_round(g1)));
                    } else if (kind == PH_HEAD) {
                        const float l0 = bf16_round(a0), l1 = bf16_round(a1);
                        *reinterpret_cast<float2*>(reinterpret_cast<float*>(aux) + n0) = make_float2(l0, l1);
                        word = pack_bf16(l0, l1);
                        const uint32_t c0 = cand_pack(l0, n0), c1 = cand_pack(l1, n0 + 1);
                        warp_cand = max(warp_cand, max(c0, c1));
                    } else {
                        // wo: h = x + wo(attn)  (P:499)      w2: x' = h + w2(act)  (P:500)
                        word = pack_bf16(__fadd_rn(bf_lo(pre), bf16_round(a0)), __fadd_rn(bf_hi(pre), bf16_round(a1)));
                    }
                    // one replica per lane: the copies leave in one store instruction
                    word = __shfl_sync(0xffffffffu, word, 0);
                    if (lane < nrep) st_relaxed_v2(out + j + (size_t)lane * len_out, word, epoch);
                    if (j == jw) LL_TRACE_AUX(5);
                    if (lane == 0 && kind == PH_QKV && !fast && n0 >= q_rows) kv_append(M, A, it, layer, n0, word);
                }
            }
            if (kind == PH_HEAD && M.force == nullptr && !(A.repeat & 2) && (fast ? A.s.fast_temp : A.s.temp) == 0.0f) {
                // greedy row: the CTA's best (logit, index) goes out as one word per replica
                if (lane == 0) s_cand[warp] = warp_cand;
                csync();
                if (warp == 0) {
                    uint32_t best = lane < kLLWarps ? s_cand[lane] : 0u;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const uint32_t ob = __shfl_xor_sync(0xffffffffu, best, o);
                        best = ob > best ? ob : best;
                    }
                    const int r = fast ? 1 + depth_pos : 0;
                    if (lane < kLLRep) st_relaxed_v2(M.ll_cand + ((size_t)r * kLLRep + lane) * kLLMaxCtas + blockIdx.x, best, epoch);
                }
            }
            // this warp is done with the stage(s): hand the ring space back to the producer
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(empty0 + 8u * (wseq % kStages));
                if (n_parts == 2) mbar_arrive(empty0 + 8u * ((wseq + 1) % kStages));
            }
            // A CTA that is done early would poll for words that cannot be there yet, and that slows the very stores it is
            // waiting for -- also when only ONE warp polls (skipping either measure before the norm phases costs +15 %).
            // So: a block barrier (idle warps cannot run ahead and poll) and a hold-off of a fraction of the hand-off
            // latency before the next phase's first poll (-20 % and -14 % frame time).  Option "ll_flags" (CallArgs.repeat): bit 2
            // drops the barrier, bits 8.. override the hold-off in units of 64 ns.
            if (!(A.repeat & 4)) csync();
            if ((A.repeat >> 8) != 255) __nanosleep((A.repeat >> 8) ? (unsigned)(A.repeat >> 8) * 64u : 256u);
            LL_TRACE(6);
            LL_SKEW(1);
        }
        if (A.mode == 1) {  // prefill: a sequence still inside its prompt advances one position
            csync();
            if (tid == 0 && it + A.iter_base < ldcg_i32(A.prompt_len) - 1) s_pos += 1;
            csync();
        }
    }
    if (blockIdx.x == 0) {
        if (tid == 0) *M.ll_epoch = s_epoch0 + (uint32_t)A.n_iter * (uint32_t)per_iter;
        if (A.mode == 1) {
            if (tid == 0) A.b.seq_len[0] = s_pos;
            if (A.finalize && tid < M.n_rows) {
                // leave the last prompt column as the pending input of the first decode frame (G:66-73)
                const int len = ldcg_i32(A.prompt_len);
                A.b.tokens[tid] = ldcg_i32(A.prompt + (size_t)tid * A.s_max + (len - 1));
            }
        }
    }
}

}  // namespace ll

// ---- host-side launch helpers (called from capi.cu) ------------------------------------------------
// Shared-memory plan of the data-flow kernel (batch 1); returns the dynamic bytes (0 = this model does not fit).
size_t ll_smem_plan(const DevModel& M, int bt, int n_ctas, int* xs_bytes, int* res_bytes, int* scratch_bytes, int* ring_bytes) {
    if (bt != 1) return 0;
    if (M.depth > ll::kLLDepth) return 0;
    if (M.vocab > ll::kCons * kSampleMaxPerThread || M.codebook_size > ll::kCons * kSampleMaxPerThread) return 0;
    int kmax = M.dim;
    if (M.inter > kmax) kmax = M.inter;
    if (M.fdim > kmax) kmax = M.fdim;
    if (M.finter > kmax) kmax = M.finter;
    if (kmax > 2 * ll::kCons * 8) return 0;  // at most two 8-element chunks per consumer thread
    const int dmax = M.dim > M.fdim ? M.dim : M.fdim;
    auto up = [](size_t v) { return (int)((v + 127) / 128 * 128); };
    *xs_bytes = up((size_t)kmax * 2);
    *res_bytes = up((size_t)dmax * 2);
    const int nl = M.vocab > M.codebook_size ? M.vocab : M.codebook_size;
    size_t sc = (size_t)kWarps * kPartialStride * sizeof(float);
    if ((size_t)nl * sizeof(float) > sc) sc = (size_t)nl * sizeof(float);
    *scratch_bytes = up(sc);
    const size_t desc_bytes = (size_t)kMaxProg / 2 * ll::kDescWords * 4;  // frames of up to 256 phases
    if (phases_per_frame(M.n_layer, M.n_flayer, M.depth) > kMaxProg / 2 || M.dim > 768 || M.fdim > 768) return 0;
    const size_t fq_bytes = (size_t)up((size_t)M.fdim * 2);
    const size_t fkv_bytes = (size_t)up((size_t)M.n_flayer * ll::kLLDepth * 2 * M.fn_kv * kHeadDim * 2);
    const size_t fixed = 2 * (size_t)*xs_bytes + 2 * (size_t)*res_bytes + (size_t)*scratch_bytes + desc_bytes + fq_bytes + fkv_bytes;
    // static shared (barriers, sampler scratch, state) stays below 2 KB.  (Leaving 35 KB more to the L1 changes nothing:
    // measured 649 vs 653 us/frame with a 190 KB cap.)
    const size_t budget = 227 * 1024 - 2048;
    // heaviest stage of one CTA
    auto per = [&](long units, long unit_bytes) { return (size_t)((units + n_ctas - 1) / n_ctas) * (size_t)unit_bytes; };
    size_t need = 0;
    auto upd = [&](size_t v) { if (v > need) need = v; };
    upd(per((M.n_head + 2 * M.n_kv) * kHeadDim / 2, 4L * M.dim));
    upd(per(M.dim / 2, 4L * M.dim));
    upd(per(M.inter / 2, 4L * M.dim));  // one part (w1 or w3 rows)
    upd(per(M.dim / 2, 4L * M.inter));
    upd(per(M.vocab / 2, 4L * M.dim));
    upd(per((M.fn_head + 2 * M.fn_kv) * kHeadDim / 2, 4L * M.fdim));
    upd(per(M.fdim / 2, 4L * M.fdim));
    upd(per(M.finter / 2, 4L * M.fdim));
    upd(per(M.fdim / 2, 4L * M.finter));
    upd(per(M.codebook_size / 2, 4L * M.fdim));
    {   // partial sums of the K slices of long rows live in a 64-float shared array
        const int f = M.inter > M.finter ? M.inter : M.finter, d = M.dim > M.fdim ? M.dim : M.fdim;
        if (((d / 2 + n_ctas - 1) / n_ctas) * ((f + 1023) / 1024) * 2 > 64) return 0;
    }
    // the ring allocator is a pure function of the stage sizes (producer and consumers replay it
    // independently); it cannot stall forever as long as two of the largest stages fit
    if (fixed + 2 * need > budget) return 0;
    {   // the gated MLP needs BOTH of its stages resident at once, possibly behind a wrap gap shorter than one stage
        const size_t part = per(M.inter / 2, 4L * M.dim) > per(M.finter / 2, 4L * M.fdim) ? per(M.inter / 2, 4L * M.dim) : per(M.finter / 2, 4L * M.fdim);
        if (fixed + 3 * part > budget) return 0;
    }
    *ring_bytes = (int)((budget - fixed) / 128 * 128);
    return fixed + (size_t)*ring_bytes;
}

static size_t g_ll_smem_configured = 0;
cudaError_t ll_configure(int bt, size_t smem) {
    (void)bt;
    if (smem <= g_ll_smem_configured) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(ll::smol_ll_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ll::smol_ll_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) g_ll_smem_configured = smem;
    return e;
}
cudaError_t ll_max_ctas(int bt, size_t smem, int* per_sm) {
    (void)bt;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, ll::smol_ll_kernel<true>, kLLThreads, smem);
}
cudaError_t ll_launch(const DevModel& M, const CallArgs& A, int bt, int n_ctas, size_t smem, int xs_bytes, int res_bytes,
                      int scratch_bytes, int ring_bytes, cudaStream_t stream) {
    (void)bt;
    ll::SmemPlan sp;
    sp.xs_bytes = xs_bytes; sp.res_bytes = res_bytes; sp.scratch_bytes = scratch_bytes; sp.ring_bytes = ring_bytes;
    sp.desc_bytes = kMaxProg / 2 * ll::kDescWords * 4;
    sp.fq_bytes = (int)(((size_t)M.fdim * 2 + 127) / 128 * 128);
    sp.fkv_bytes = (int)(((size_t)M.n_flayer * ll::kLLDepth * 2 * M.fn_kv * kHeadDim * 2 + 127) / 128 * 128);
    void* args[3] = {(void*)&M, (void*)&A, (void*)&sp};
    // cooperative launch only for the co-residency guarantee: CTAs spin on each other's words
    const void* fn = M.prof != nullptr ? (const void*)ll::smol_ll_kernel<true> : (const void*)ll::smol_ll_kernel<false>;
    return cudaLaunchCooperativeKernel(fn, dim3(n_ctas), dim3(kLLThreads), args, smem, stream);
}

}  // namespace smol
