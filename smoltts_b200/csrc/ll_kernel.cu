// The DualAR decode step as a DATA-FLOW persistent kernel (sm_100a) -- the latency path (batch <= 8).
//
// Same phase program, same arithmetic and the same rounding points as decode_kernel.cu (results are
// bit-identical, tests/test_gpu_decode.py), but no grid barrier anywhere.  At batch 1 a frame is a chain
// of ~200 dependent matrix-vector products; what bounds it is the hand-off between them, not bytes.  So:
//
//  * every phase PUBLISHES its output as 8-byte words {two bf16 | epoch} ("LL" protocol: the flag
//    travels in the same store as the data).  A consumer polls exactly the words it needs with
//    ld.relaxed.gpu and starts the moment they carry the epoch of (frame, phase): one L2 store + one L2
//    load per hand-off, no fence, no atomic, no barrier, no skew wait.  Each vector is written to
//    kLLRep replicas so that the 148 polling CTAs spread over more L2 slices.
//  * weights never wait for activations: warp 16 of every CTA is a TMA producer that streams the
//    CTA's rows of ALL coming phases (cp.async.bulk -> shared-memory ring, full/empty mbarriers) as
//    far ahead as the ring allows (~1.5 layers), so HBM runs continuously under the compute chain.
//  * sampled ids are published once per CTA; each CTA tracks seq_len / step / finished itself, so
//    one launch runs any number of frames with no host round trip and no global state on the path.
//
// Reference map (P = modeling/model/rq_transformer.py, M = mlx lm/rq_transformer.py, G = mlx
// lm/generate.py): embed P:205-221; RMSNorm P:601-613; QKV/RoPE/attention P:535-570,616-640;
// FeedForward P:573-582; block P:492-501; slow step P:223-260 / M:173-192; depth loop P:409-448,
// M:194-220, G:110-141; sampling G:88-99,118-132; frame assembly + stop rule G:143-171.

#define SMOL_BLOCK_SYNC() asm volatile("bar.sync 1, 512;" ::: "memory")

#include "common.cuh"
#include "dev_model.h"
#include "sampler.cuh"

namespace smol {

namespace ll {

constexpr int kStages = 16;  // mbarrier pairs of the weight ring (phases in flight)

// ---- small helpers -------------------------------------------------------------------------------
__device__ __forceinline__ void csync() { SMOL_BLOCK_SYNC(); }  // the 16 consumer warps
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int ldcg_i32(const void* p) { return __ldcg(reinterpret_cast<const int*>(p)); }
__device__ __forceinline__ uint4 ldcg_v4(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint16_t f_to_bf(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float bf_to_f(uint16_t v) { return __uint_as_float(((uint32_t)v) << 16); }
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

// ---- LL words ------------------------------------------------------------------------------------
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
    uint2 v;
    do { v = ld_relaxed_v2(w); } while (v.y != epoch);
    return v.x;
}
// N consecutive 16-byte pairs of words (4 bf16 each), polled together
template <int N>
__device__ __forceinline__ void ll_get4(const unsigned long long* w, uint32_t epoch, uint32_t (&out)[2 * N]) {
    const uint4* p = reinterpret_cast<const uint4*>(w);
    bool ok;
    do {
        ok = true;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint4 v = ld_relaxed_v4(p + i);
            out[2 * i] = v.x; out[2 * i + 1] = v.z;
            ok = ok && (v.y == epoch) && (v.w == epoch);
        }
    } while (!ok);
}
// 8 consecutive elements (4 words) -> fp32
__device__ __forceinline__ void ll_get8f(const unsigned long long* w, uint32_t epoch, float (&f)[8]) {
    uint32_t u[4];
    ll_get4<2>(w, epoch, u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = bf_lo(u[i]); f[2 * i + 1] = bf_hi(u[i]); }
}
// publish one word to `nrep` replicas `len` words apart
__device__ __forceinline__ void ll_put(unsigned long long* w, int len, int nrep, uint32_t payload, uint32_t epoch) {
    for (int r = 0; r < nrep; ++r) st_relaxed_v2(w + (size_t)r * len, payload, epoch);
}

// ---- mbarrier + TMA bulk copy ----------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "LL_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LL_WAIT_DONE;\n\t"
        "bra LL_WAIT_LOOP;\n\t"
        "LL_WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// ---- per-CTA shared state ------------------------------------------------------------------------
struct Shared {
    uint64_t full[kStages];
    uint64_t empty[kStages];
    SampleScratch sc;
    float part[kWarps];
    int pos[kLLMaxBatch];      // cached positions (seq_len) of each sequence, tracked locally
    int step[kLLMaxBatch];
    int fin[kLLMaxBatch];
    int tok[kLLMaxBatch][kMaxRows];  // pending input column
    int nw[kLLMaxBatch][kMaxRows];   // ids sampled in the current frame
};

struct Ctx {
    int cta, n_ctas, warp, lane, tid, rep;
    Shared* sh;
    uint16_t* xs[2];       // activation rows of the batch tile, bf16, double-buffered
    uint16_t* res[2];      // residual rows: [0] input of the attention half, [1] input of the MLP half
    float* scratch;        // attention merge / sampler logits
    unsigned char* ring;   // weight ring
    uint32_t ring_bytes;
    int xcur;              // xs buffer of the current use
    int iter;
};

// ---- weight plan ---------------------------------------------------------------------------------
// A phase's GEMV is U units; a unit is one row PAIR (two adjacent outputs = one LL word); the gated MLP
// unit is two adjacent rows of w1 plus the same two rows of w3.  CTA c owns units [U*c/n, U*(c+1)/n):
// one contiguous row block per matrix -> one bulk copy.
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
__device__ __forceinline__ int unit_begin(int U, int cta, int n) { return (int)(((long long)U * cta) / n); }
// bytes of ONE part of the phase (the gated MLP has two parts: this CTA's w1 rows, then its w3 rows; each
// part is its own ring stage so that no stage exceeds half the ring)
__device__ __forceinline__ uint32_t part_bytes(const Plan& pl, int cta, int n) {
    const int nu = unit_begin(pl.U, cta + 1, n) - unit_begin(pl.U, cta, n);
    return (uint32_t)nu * 2u * (uint32_t)pl.K * 2u;
}
// ring allocator replayed identically by the producer and by the consumers
__device__ __forceinline__ uint32_t ring_alloc(uint32_t& head, uint32_t bytes, uint32_t ring_bytes, uint32_t& gap) {
    gap = 0;
    if (head + bytes > ring_bytes) { gap = ring_bytes - head; head = 0; }
    const uint32_t start = head;
    head += bytes;
    return start;
}

// ---- TMA producer (one lane of warp 16) -------------------------------------------------------------
__device__ __noinline__ void producer(const DevModel& M, const CallArgs& A, const Ctx& c) {
    Shared& S = *c.sh;
    uint64_t pol_stream, pol_keep;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    uint32_t held[kStages];  // ring bytes (gap included) held by each in-flight phase
    uint32_t wseq = 0, rel = 0, head = 0, used = 0;
    for (int it = 0; it < A.n_iter; ++it) {
        for (int p = A.phase_begin; p < A.phase_end; ++p) {
            const Phase ph = unpack_phase(M.prog[p]);
            if (ph.kind == PH_ATTN || ph.kind == PH_SAMPLE) continue;
            const Plan pl = phase_plan(M, ph);
            const uint32_t bytes = part_bytes(pl, c.cta, c.n_ctas);
            const int u0 = unit_begin(pl.U, c.cta, c.n_ctas);
            // the slow transformer's weights are read once per frame: evict first; the depth
            // transformer's are re-read by every depth step: keep them in L2
            const uint64_t pol = (ph.fast && ph.kind != PH_HEAD) ? pol_keep : pol_stream;
            for (int part = 0; part < (pl.w1 ? 2 : 1); ++part) {
                uint32_t gap = (head + bytes > c.ring_bytes) ? (c.ring_bytes - head) : 0u;
                // wait until the ring has room and the stage's barrier pair is free
                while (used + gap + bytes > c.ring_bytes || wseq - rel >= (uint32_t)kStages) {
                    mbar_wait(&S.empty[rel % kStages], (rel / kStages) & 1u);
                    used -= held[rel % kStages];
                    ++rel;
                }
                const uint32_t start = ring_alloc(head, bytes, c.ring_bytes, gap);
                held[wseq % kStages] = gap + bytes;
                used += gap + bytes;
                uint64_t* bar = &S.full[wseq % kStages];
                mbar_expect_tx(bar, bytes);
                if (bytes) bulk_g2s(c.ring + start, (part ? pl.w1 : pl.w0) + (size_t)2 * u0 * pl.K, bytes, bar, pol);
                ++wseq;
            }
        }
    }
}

// ---- inputs ----------------------------------------------------------------------------------------
__device__ __forceinline__ bool seq_active(const DevModel& M, const CallArgs& A, const Ctx& c, int b) {
    if (A.mode == 1) return c.iter + A.iter_base < ldcg_i32(A.prompt_len + b) - 1;
    return c.sh->fin[b] == 0;
}
__device__ __forceinline__ int input_token(const DevModel& M, const CallArgs& A, const Ctx& c, int b, int r) {
    if (A.mode == 1) {
        int t = c.iter + A.iter_base;
        const int len = ldcg_i32(A.prompt_len + b);
        if (t > len - 1) t = len - 1;
        return ldcg_i32(A.prompt + ((size_t)b * M.n_rows + r) * A.s_max + t);
    }
    return c.sh->tok[b][r];
}
// BaseTransformer.embed (P:205-221) for one 8-element chunk
__device__ __forceinline__ void embed_chunk(const DevModel& M, const CallArgs& A, const Ctx& c, int b, int ch, float (&f)[8]) {
    const int D = M.dim;
    const int t0 = input_token(M, A, c, b, 0);
    bool use_vq;
    if (M.mlx_embed_mask) use_vq = (t0 >= M.semantic_start && t0 <= M.semantic_end);
    else use_vq = input_token(M, A, c, b, 1) != 0;
    const uint4 row0 = ldcg_v4(M.embeddings + (size_t)t0 * D + ch * 8);
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
    if (use_vq) {
        for (int r0 = 1; r0 < M.n_rows; r0 += 8) {  // eight row loads in flight, summed in row order
            uint4 rows[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + i;
                if (r < M.n_rows) {
                    const int code = input_token(M, A, c, b, r);
                    const int row = code + (M.dup0 ? (r - 1) : r) * M.codebook_size;
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
    }
    unpack8(row0, f);
    if (use_vq) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = bf16_round(__fadd_rn(f[e], bf16_round(s[e])));
    }
}

// region of phase p, sequence b, this CTA's replica (read side) / replica 0 (write side)
__device__ __forceinline__ const unsigned long long* ll_src(const DevModel& M, const Ctx& c, int p, int b) {
    return M.ll + M.ll_off[p] + (size_t)(b * kLLRep + c.rep) * M.ll_len[p];
}
__device__ __forceinline__ unsigned long long* ll_dst(const DevModel& M, int p, int b) {
    return M.ll + M.ll_off[p] + (size_t)(b * kLLRep) * M.ll_len[p];
}

// RMSNorm (P:601-613) of nb rows delivered chunk-wise by load_chunk(b, ch, f) into xs (bf16) and, raw,
// into the residual buffer.  Same arithmetic as decode_kernel.cu: one thread per 8-element chunk.
template <class LoadChunk>
__device__ __forceinline__ void norm_rows(const Ctx& c, int K, int nb, LoadChunk load_chunk, const uint16_t* w, float eps,
                                          uint16_t* xs, uint16_t* res) {
    const int nch = K >> 3;
    const int wpr = (nch + 31) >> 5;
    const int rows_per_pass = kWarps / wpr;
    const int wr = c.warp / wpr, wi = c.warp - wr * wpr;
    const int ch = wi * 32 + c.lane;
    for (int bp = 0; bp < nb; bp += rows_per_pass) {
        const int b = bp + wr;
        const bool act = (wr < rows_per_pass) && (b < nb) && (ch < nch);
        float x[8];
        float ss = 0.f;
        uint4 wv = make_uint4(0u, 0u, 0u, 0u);
        if (act) {
            wv = __ldg(reinterpret_cast<const uint4*>(w + ch * 8));
            load_chunk(b, ch, x);
#pragma unroll
            for (int e = 0; e < 8; ++e) ss = fmaf(x[e], x[e], ss);
        }
        ss = warp_sum(ss);
        if (c.lane == 0) c.sh->part[c.warp] = ss;
        csync();
        if (act) {
            float t = 0.f;
            for (int i = 0; i < wpr; ++i) t += c.sh->part[wr * wpr + i];
            const float mean = __fdiv_rn(t, (float)K);
            const float r = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, eps)));
            if (res != nullptr) *reinterpret_cast<uint4*>(res + (size_t)b * K + ch * 8) = pack8(x);
            float wf[8], o[8];
            unpack8(wv, wf);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = bf16_round(__fmul_rn(bf16_round(__fmul_rn(x[e], r)), wf[e]));
            *reinterpret_cast<uint4*>(xs + (size_t)b * K + ch * 8) = pack8(o);
        }
        csync();
    }
}

// nb rows of K elements straight from an LL region into xs
__device__ __forceinline__ void fill_rows(const DevModel& M, const Ctx& c, int K, int nb, int p_src, uint32_t epoch, uint16_t* xs) {
    const int nch = K >> 3;
    for (int idx = c.tid; idx < nb * nch; idx += kThreads) {
        const int b = idx / nch, ch = idx - b * nch;
        uint32_t u[4];
        ll_get4<2>(ll_src(M, c, p_src, b) + ch * 4, epoch, u);
        *reinterpret_cast<uint4*>(xs + (size_t)b * K + ch * 8) = make_uint4(u[0], u[1], u[2], u[3]);
    }
    csync();
}

// Attention of the depth transformer (<= kMaxDepth positions), recomputed by every CTA straight from the
// published q/k/v words of this and the earlier depth steps into xs (the input of the wo GEMV).  One warp
// per (sequence, head); arithmetic identical to decode_kernel.cu: fast_attention_rows.
__device__ __forceinline__ void fast_attention_rows(const DevModel& M, const Ctx& c, int layer, int depth_pos, int p_now,
                                                    uint32_t epoch_now, int nb, uint16_t* xs) {
    const int Hq = M.fn_head, Hkv = M.fn_kv, G = Hq / Hkv, D = M.fdim;
    const int q_words = Hq * kHeadDim / 2, k_end_words = (Hq + Hkv) * kHeadDim / 2;
    const int per = 4 * M.n_flayer + 2;
    const int jl = c.lane >> 2, part = c.lane & 3;
    const int p_qkv = p_now - 1;                   // this step's QKV phase
    const uint32_t e_qkv = epoch_now - 1;
    for (int pair = c.warp; pair < nb * Hq; pair += kWarps) {
        const int b = pair / Hq, hq = pair - b * Hq, kvh = hq / G;
        const bool deep = depth_pos >= 8;
        // issue everything, then wait: q (8 words), k of my position(s) (8 words each), v of every position (1 word each)
        uint32_t qw[8], kw[2][8], vw[kMaxDepth];
        bool ok[2];
        ok[0] = jl <= depth_pos; ok[1] = deep && (jl + 8 <= depth_pos);
        ll_get4<4>(ll_src(M, c, p_qkv, b) + hq * 32 + part * 8, e_qkv, qw);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            if (ok[h2]) {
                const int j = jl + 8 * h2;
                const int pj = p_qkv - (depth_pos - j) * per;
                ll_get4<4>(ll_src(M, c, pj, b) + q_words + kvh * 32 + part * 8, e_qkv - (uint32_t)((depth_pos - j) * per), kw[h2]);
            }
        }
#pragma unroll
        for (int j = 0; j < kMaxDepth; ++j) {
            vw[j] = 0u;
            if (j <= depth_pos && (j < 8 || deep)) {
                const int pj = p_qkv - (depth_pos - j) * per;
                vw[j] = ll_get(ll_src(M, c, pj, b) + k_end_words + kvh * 32 + c.lane, e_qkv - (uint32_t)((depth_pos - j) * per));
            }
        }
        float qf[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) { qf[2 * i] = bf_lo(qw[i]); qf[2 * i + 1] = bf_hi(qw[i]); }
        float sc[2];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            sc[h2] = -INFINITY;
            if (h2 == 0 || deep) {
                float s = 0.f;
                if (ok[h2]) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        s = fmaf(qf[2 * i], bf_lo(kw[h2][i]), s);
                        s = fmaf(qf[2 * i + 1], bf_hi(kw[h2][i]), s);
                    }
                }
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (ok[h2]) sc[h2] = s * 0.125f;
            }
        }
        float m = fmaxf(sc[0], sc[1]);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
        const float pe0 = ok[0] ? expf(sc[0] - m) : 0.f;
        const float pe1 = ok[1] ? expf(sc[1] - m) : 0.f;
        float l = pe0 + pe1;
        l += __shfl_xor_sync(0xffffffffu, l, 4);
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 16);
        const float pb0 = bf16_round(pe0), pb1 = bf16_round(pe1);
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float pj = __shfl_sync(0xffffffffu, pb0, j * 4);
            o0 = fmaf(pj, bf_lo(vw[j]), o0);
            o1 = fmaf(pj, bf_hi(vw[j]), o1);
        }
        if (deep) {
#pragma unroll
            for (int j = 8; j < kMaxDepth; ++j) {
                const float pj = __shfl_sync(0xffffffffu, pb1, (j & 7) * 4);
                if (j <= depth_pos) {
                    o0 = fmaf(pj, bf_lo(vw[j]), o0);
                    o1 = fmaf(pj, bf_hi(vw[j]), o1);
                }
            }
        }
        const float inv = 1.0f / l;
        *reinterpret_cast<uint32_t*>(xs + (size_t)b * D + hq * kHeadDim + 2 * c.lane) = pack_bf16(bf16_round(o0 * inv), bf16_round(o1 * inv));
    }
}

// ---- split-KV decode attention over the paged cache (slow layers) ----------------------------------
// Same units, same merge order and the same arithmetic as decode_kernel.cu: phase_attn_g; q and the newest
// position's K/V come from the QKV phase's published words, older positions from the paged pool, and the
// result leaves as LL words: directly (one split) or as (m, l, o[64]) partials that the head's combiner
// CTA merges in split order.
__device__ __forceinline__ int attn_splits(const DevModel& M, int batch, int n_ctas) {
    const int pairs = batch * M.n_kv;
    int s = (2 * n_ctas + pairs - 1) / pairs;
    if (s < 1) s = 1;
    if (s > kMaxSplits) s = kMaxSplits;
    return s;
}
__device__ __forceinline__ int seq_splits(const DevModel& M, const CallArgs& A, int pos, int s_max, int& Lb) {
    Lb = pos + 1;
    const int cap = A.b.max_pages * M.page_size;
    if (Lb > cap) Lb = cap;
    int ns = (Lb + kSplitMin - 1) / kSplitMin;
    if (ns > s_max) ns = s_max;
    if (ns < 1) ns = 1;
    return ns;
}
__device__ __forceinline__ unsigned long long* partial_words(const DevModel& M, int layer, int b, int hq, int s) {
    return M.ll_partial + ((((size_t)(layer & 1) * M.ll_batch + b) * M.n_head + hq) * kMaxSplits + s) * kPartialStride;
}

__device__ __noinline__ void phase_attn(const DevModel& M, const CallArgs& A, Ctx& c, const Phase& ph, int p, uint32_t epoch) {
    constexpr int GM = kMaxGroup;
    const int Hq = M.n_head, Hkv = M.n_kv, ps = M.page_size, G = Hq / Hkv;
    const int q_words = Hq * kHeadDim / 2, k_end_words = (Hq + Hkv) * kHeadDim / 2;
    const int dch = c.lane & 7, psub = c.lane >> 3;
    float* red = c.scratch;  // [kWarps][G][kPartialStride]
    const int s_max = attn_splits(M, A.batch, c.n_ctas);
    const int n_units = A.batch * Hkv * s_max;
    const int len_out = M.ll_len[p];
    for (int u = c.cta; u < n_units; u += c.n_ctas) {
        const int s = u % s_max, kvh = (u / s_max) % Hkv, b = u / (s_max * Hkv);
        int Lb;
        const int pos_new = c.sh->pos[b];
        const int ns = seq_splits(M, A, pos_new, s_max, Lb);
        if (s >= ns) continue;
        const int chunk = (Lb + ns - 1) / ns;
        const int p0 = s * chunk, p1 = min(Lb, p0 + chunk);
        const unsigned long long* qkv = ll_src(M, c, p - 1, b);
        const int32_t* bt = A.b.block_table + (size_t)b * A.b.max_pages;
        const size_t head_stride = (size_t)ps * kHeadDim;

        float m[GM], l[GM], acc[GM][8];
#pragma unroll
        for (int g = 0; g < GM; ++g) {
            m[g] = -INFINITY; l[g] = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[g][e] = 0.f;
        }
        float qf[GM][8];
        bool have_q = false;
        for (int pb = p0 + c.warp * 4; pb < p1; pb += kWarps * 4) {
            const int pp = pb + psub;
            const bool valid = pp < p1;
            float kf[8], vf[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { kf[e] = 0.f; vf[e] = 0.f; }
            uint4 kk = make_uint4(0u, 0u, 0u, 0u), vv = kk;
            const bool newest = valid && (pp == pos_new);
            if (valid && !newest) {  // cached position: issue the pool loads before waiting for q
                const int page = ldcg_i32(bt + pp / ps);
                const uint16_t* kp = M.kv_pool + ((((size_t)page * M.n_layer + ph.layer) * 2) * Hkv + kvh) * head_stride
                                     + (size_t)(pp % ps) * kHeadDim + dch * 8;
                kk = ldcg_v4(kp);
                vv = ldcg_v4(kp + (size_t)Hkv * head_stride);
            }
            if (!have_q) {
#pragma unroll
                for (int g = 0; g < GM; ++g) {
                    if (g >= G) continue;
                    ll_get8f(qkv + (kvh * G + g) * 32 + dch * 4, epoch - 1, qf[g]);
#pragma unroll
                    for (int e = 0; e < 8; ++e) qf[g][e] *= 0.125f;  // 1/sqrt(64), exact
                }
                have_q = true;
            }
            if (newest) {
                ll_get8f(qkv + q_words + kvh * 32 + dch * 4, epoch - 1, kf);
                ll_get8f(qkv + k_end_words + kvh * 32 + dch * 4, epoch - 1, vf);
            } else {
                unpack8(kk, kf);
                unpack8(vv, vf);
            }
#pragma unroll
            for (int g = 0; g < GM; ++g) {
                if (g >= G) continue;
                float sc = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) sc = fmaf(qf[g][e], kf[e], sc);
                sc += __shfl_xor_sync(0xffffffffu, sc, 1);
                sc += __shfl_xor_sync(0xffffffffu, sc, 2);
                sc += __shfl_xor_sync(0xffffffffu, sc, 4);
                if (valid) {
                    const float mn = fmaxf(m[g], sc);
                    const float corr = (m[g] == -INFINITY) ? 0.f : expf(m[g] - mn);
                    const float pe = expf(sc - mn);
                    l[g] = fmaf(l[g], corr, pe);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[g][e] = fmaf(acc[g][e], corr, pe * vf[e]);
                    m[g] = mn;
                }
            }
        }
        // merge the four position groups of the warp
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
#pragma unroll
            for (int g = 0; g < GM; ++g) {
                if (g >= G) continue;
                const float mo = __shfl_xor_sync(0xffffffffu, m[g], o);
                const float lo = __shfl_xor_sync(0xffffffffu, l[g], o);
                const float mn = fmaxf(m[g], mo);
                const float c1 = (m[g] == -INFINITY) ? 0.f : expf(m[g] - mn);
                const float c2 = (mo == -INFINITY) ? 0.f : expf(mo - mn);
                l[g] = l[g] * c1 + lo * c2;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float ao = __shfl_xor_sync(0xffffffffu, acc[g][e], o);
                    acc[g][e] = acc[g][e] * c1 + ao * c2;
                }
                m[g] = mn;
            }
        }
        if (psub == 0) {
#pragma unroll
            for (int g = 0; g < GM; ++g) {
                if (g >= G) continue;
                float* dst = red + (c.warp * G + g) * kPartialStride;
                if (dch == 0) { dst[0] = m[g]; dst[1] = l[g]; }
#pragma unroll
                for (int e = 0; e < 8; ++e) dst[2 + dch * 8 + e] = acc[g][e];
            }
        }
        csync();
        // merge the warps (fixed order)
        float Mx = -INFINITY, Ls = 0.f, Os = 0.f;
        const int g = c.tid / kHeadDim, d = c.tid & (kHeadDim - 1);
        const bool owner = c.tid < G * kHeadDim;
        if (owner) {
            for (int wi = 0; wi < kWarps; ++wi) Mx = fmaxf(Mx, red[(wi * G + g) * kPartialStride]);
            for (int wi = 0; wi < kWarps; ++wi) {
                const float* src = red + (wi * G + g) * kPartialStride;
                if (src[0] == -INFINITY) continue;
                const float sc = expf(src[0] - Mx);
                Ls = fmaf(src[1], sc, Ls);
                Os = fmaf(src[2 + d], sc, Os);
            }
        }
        const int hq = kvh * G + g;
        if (ns == 1) {
            const float o = owner ? bf16_round(Os / Ls) : 0.f;
            const float o_next = __shfl_down_sync(0xffffffffu, o, 1);
            if (owner && !(d & 1))
                ll_put(ll_dst(M, p, b) + (hq * kHeadDim + d) / 2, len_out, kLLRep, pack_bf16(o, o_next), epoch);
        } else if (owner) {
            unsigned long long* dst = partial_words(M, ph.layer, b, hq, s);
            if (d == 0) {
                st_relaxed_v2(dst, __float_as_uint(Mx), epoch);
                st_relaxed_v2(dst + 1, __float_as_uint(Ls), epoch);
            }
            st_relaxed_v2(dst + 2 + d, __float_as_uint(Os), epoch);
        }
        csync();
    }
    // combine the splits of one (sequence, head): one CTA per head, chosen after the attention units
    for (int t = 0; t < A.batch * Hq; ++t) {
        if ((n_units + t) % c.n_ctas != c.cta) continue;
        const int b = t / Hq, hq = t - b * Hq;
        int Lb;
        const int ns = seq_splits(M, A, c.sh->pos[b], s_max, Lb);
        if (ns == 1) continue;
        if (c.tid < kHeadDim) {
            const int d = c.tid;
            float ms[kMaxSplits], lsum[kMaxSplits], os[kMaxSplits];
            bool ok;
            do {  // all splits polled together
                ok = true;
#pragma unroll 4
                for (int si = 0; si < ns; ++si) {
                    const unsigned long long* src = partial_words(M, ph.layer, b, hq, si);
                    const uint4 ml = ld_relaxed_v4(src);
                    const uint2 ov = ld_relaxed_v2(src + 2 + d);
                    ok = ok && ml.y == epoch && ml.w == epoch && ov.y == epoch;
                    ms[si] = __uint_as_float(ml.x); lsum[si] = __uint_as_float(ml.z); os[si] = __uint_as_float(ov.x);
                }
            } while (!ok);
            float Mg = -INFINITY;
            for (int si = 0; si < ns; ++si) Mg = fmaxf(Mg, ms[si]);
            float Lg = 0.f, Og = 0.f;
            for (int si = 0; si < ns; ++si) {
                if (ms[si] == -INFINITY) continue;
                const float sc = expf(ms[si] - Mg);
                Lg = fmaf(lsum[si], sc, Lg);
                Og = fmaf(os[si], sc, Og);
            }
            const float o = bf16_round(Og / Lg);
            const float o_next = __shfl_down_sync(0xffffffffu, o, 1);
            if (!(d & 1)) ll_put(ll_dst(M, p, b) + (hq * kHeadDim + d) / 2, len_out, kLLRep, pack_bf16(o, o_next), epoch);
        }
    }
}

// ---- GEMV over one unit (row pair), weights and activations in shared memory -------------------------
// acc[r][b] = sum_k W_r[k] * xs[b][k]: lane-strided 8-element chunks, two accumulator chains per row
// (elements 0-3 / 4-7 of a chunk), butterfly warp sum -- the summation order of decode_kernel.cu.
template <int BT, int R>
__device__ __forceinline__ void gemv_unit(const Ctx& c, const uint4* const (&wrow)[R], const uint16_t* xs, int K, float (&mine)[R]) {
    const int nchunks = K >> 3;
    float acc0[R][BT], acc1[R][BT];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int b = 0; b < BT; ++b) { acc0[r][b] = 0.f; acc1[r][b] = 0.f; }
#pragma unroll(BT <= 2 ? 3 : 1)
    for (int ch = c.lane; ch < nchunks; ch += 32) {
        float wf[R][8];
#pragma unroll
        for (int r = 0; r < R; ++r) unpack8(wrow[r][ch], wf[r]);
#pragma unroll
        for (int b = 0; b < BT; ++b) {
            float x[8];
            unpack8(*reinterpret_cast<const uint4*>(xs + (size_t)b * K + ch * 8), x);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float a0 = acc0[r][b], a1 = acc1[r][b];
                a0 = fmaf(wf[r][0], x[0], a0); a1 = fmaf(wf[r][4], x[4], a1);
                a0 = fmaf(wf[r][1], x[1], a0); a1 = fmaf(wf[r][5], x[5], a1);
                a0 = fmaf(wf[r][2], x[2], a0); a1 = fmaf(wf[r][6], x[6], a1);
                a0 = fmaf(wf[r][3], x[3], a0); a1 = fmaf(wf[r][7], x[7], a1);
                acc0[r][b] = a0; acc1[r][b] = a1;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        mine[r] = 0.f;
#pragma unroll
        for (int b = 0; b < BT; ++b) {
            const float s = warp_sum(acc0[r][b] + acc1[r][b]);
            if (c.lane == b) mine[r] = s;
        }
    }
}

// Resolve ids published by the sampler CTAs: word (b, r) of this CTA, epoch of that row's SAMPLE phase.
__device__ __forceinline__ int token_word_get(const DevModel& M, const Ctx& c, int b, int r, uint32_t epoch) {
    return (int)ll_get(M.ll_tok + ((size_t)b * M.n_rows + r) * kLLMaxCtas + c.cta, epoch);
}
__device__ __forceinline__ int sample_phase_index(const DevModel& M, int r) {  // program index of row r's SAMPLE phase
    const int n_slow = 5 * M.n_layer;
    if (r == 0) return n_slow + 1;
    return n_slow + 2 + (r - 1) * (4 * M.n_flayer + 2) + 4 * M.n_flayer + 1;
}

// ---- every phase that streams weights ---------------------------------------------------------------
template <int BT>
__device__ void phase_gemv(const DevModel& M, const CallArgs& A, Ctx& c, const Phase& ph, int p, uint32_t epoch,
                           uint32_t& wseq, uint32_t& ring_head, int per_iter) {
    Shared& S = *c.sh;
    const bool fast = ph.fast != 0;
    const int kind = ph.kind;
    const int D = fast ? M.fdim : M.dim, F = fast ? M.finter : M.inter;
    const int Hq = fast ? M.fn_head : M.n_head, Hkv = fast ? M.fn_kv : M.n_kv;
    const DevLayer& L = fast ? M.fast_layers[ph.layer] : M.layers[ph.layer];
    const int q_rows = Hq * kHeadDim, k_end = (Hq + Hkv) * kHeadDim;
    const int n_head_rows = fast ? M.codebook_size : M.vocab;
    const int K = (kind == PH_W2) ? F : D;
    const bool normed = kind == PH_QKV || kind == PH_W13 || kind == PH_HEAD;
    const uint16_t* norm_w = kind == PH_QKV ? L.attention_norm : kind == PH_W13 ? L.ffn_norm : (fast ? M.fast_norm : M.norm);
    const uint16_t* table = fast ? M.fast_rope : M.rope;
    const int nb = A.batch;
    const int n_slow = 5 * M.n_layer;
    const bool prof = (M.prof != nullptr) && c.cta == 0 && c.tid == 0;
    unsigned long long* seg = M.prof + 2 * kMaxProg + (size_t)(kind + (fast ? 8 : 0)) * 4;
    unsigned long long ts = 0;
    if (prof) ts = globaltimer_ns();

    c.xcur ^= 1;
    uint16_t* xs = c.xs[c.xcur];

    // ---- prologue: the phase's input rows -> xs -------------------------------------------------------
    if (normed) {
        // 0 token embedding (P:205-221), 2 embedding of the previous depth code (G:136-140), 3 an LL region
        int src_kind = 3, p_src = p - 1;
        uint32_t e_src = epoch - 1;
        uint16_t* res = kind == PH_QKV ? c.res[0] : kind == PH_W13 ? c.res[1] : nullptr;
        if (kind == PH_QKV && ph.layer == 0) {
            if (!fast) {
                src_kind = 0;
                if (A.mode == 0 && c.iter > 0) {
                    // frame boundary: adopt the ids sampled in the previous frame (G:143-166)
                    if (c.tid < nb * M.n_rows) {
                        const int b = c.tid / M.n_rows, r = c.tid - b * M.n_rows;
                        S.nw[b][r] = token_word_get(M, c, b, r, epoch - (uint32_t)p - (uint32_t)per_iter + (uint32_t)sample_phase_index(M, r));
                    }
                    csync();
                    if (c.tid < nb) {
                        const int b = c.tid;
                        if (S.fin[b] == 0) {
                            for (int r = 0; r < M.n_rows; ++r) S.tok[b][r] = S.nw[b][r];
                            S.pos[b] += 1;
                            S.step[b] += 1;
                            if (A.b.finished != nullptr && A.s.audio_only && !A.s.ignore_stop && S.nw[b][0] == M.im_end) S.fin[b] = 1;
                        }
                    }
                    csync();
                    __threadfence();  // acquire side of the once-per-frame fence: KV written last frame is visible
                }
            } else if (ph.depth_pos == 0) {
                p_src = n_slow - 1;  // the slow transformer's pre-norm hidden state (P:259, M:191)
                e_src = epoch - (uint32_t)(p - p_src);
            } else {
                src_kind = 2;
                if (c.tid < nb) {
                    const int r = ph.depth_pos;  // row of depth code depth_pos-1
                    S.nw[c.tid][r] = token_word_get(M, c, c.tid, r, epoch - (uint32_t)(p - sample_phase_index(M, r)));
                }
                csync();
            }
        }
        norm_rows(c, D, nb, [&](int b, int ch, float (&f)[8]) {
            if (src_kind == 0) { embed_chunk(M, A, c, b, ch, f); return; }
            if (src_kind == 2) {
                const int code = S.nw[b][ph.depth_pos];
                const int off = M.depthwise_wte ? (M.dup0 ? ph.depth_pos - 1 : ph.depth_pos) * M.codebook_size : 0;
                unpack8(ldcg_v4(M.fast_embeddings + (size_t)(code + off) * D + ch * 8), f);
                return;
            }
            ll_get8f(ll_src(M, c, p_src, b) + ch * 4, e_src, f);
        }, norm_w, M.eps, xs, res);
    } else if (kind == PH_WO && fast) {
        fast_attention_rows(M, c, ph.layer, ph.depth_pos, p, epoch, nb, xs);
        csync();
    } else {
        fill_rows(M, c, K, nb, p - 1, epoch - 1, xs);
    }
    if (prof) { const unsigned long long t = globaltimer_ns(); seg[0] += t - ts; ts = t; }

    // ---- this CTA's weight rows ------------------------------------------------------------------------
    const Plan pl = phase_plan(M, ph);
    const int u0 = unit_begin(pl.U, c.cta, c.n_ctas), u1 = unit_begin(pl.U, c.cta + 1, c.n_ctas);
    const uint32_t bytes = part_bytes(pl, c.cta, c.n_ctas);
    const int n_parts = pl.w1 ? 2 : 1;
    uint32_t gap;
    const unsigned char* wbase = c.ring + ring_alloc(ring_head, bytes, c.ring_bytes, gap);
    const unsigned char* wbase2 = wbase;
    if (n_parts == 2) wbase2 = c.ring + ring_alloc(ring_head, bytes, c.ring_bytes, gap);
    if (c.warp < u1 - u0) {
        for (int i = 0; i < n_parts; ++i) mbar_wait(&S.full[(wseq + i) % kStages], ((wseq + i) / kStages) & 1u);
    }
    if (prof) { const unsigned long long t = globaltimer_ns(); seg[1] += t - ts; ts = t; }

    const size_t row_bytes = (size_t)K * 2;
    unsigned long long* out0 = ll_dst(M, p, 0);
    const int len_out = M.ll_len[p];
    const bool last_depth_head = fast && kind == PH_HEAD && ph.depth_pos == M.depth - 1;
    if (last_depth_head) __threadfence();  // release side of the once-per-frame fence (KV appended this frame)

    for (int j = c.warp; j < u1 - u0; j += kWarps) {
        const int u = u0 + j;
        const int n0 = 2 * u;
        if (kind == PH_W13) {
            const uint4* const wrow[4] = {
                reinterpret_cast<const uint4*>(wbase + (size_t)(2 * j) * row_bytes),
                reinterpret_cast<const uint4*>(wbase + (size_t)(2 * j + 1) * row_bytes),
                reinterpret_cast<const uint4*>(wbase2 + (size_t)(2 * j) * row_bytes),
                reinterpret_cast<const uint4*>(wbase2 + (size_t)(2 * j + 1) * row_bytes)};
            float acc[4];
            gemv_unit<BT, 4>(c, wrow, xs, K, acc);
            if (c.lane < nb) {
                float o[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float a = bf16_round(acc[i]), g = bf16_round(acc[2 + i]);
                    const float sg = bf16_round(__fdiv_rn(a, __fadd_rn(1.0f, expf(-a))));  // F.silu in fp32, bf16 out
                    o[i] = __fmul_rn(sg, g);
                }
                ll_put(out0 + (size_t)c.lane * kLLRep * len_out + u, len_out, kLLRep, pack_bf16(o[0], o[1]), epoch);
            }
            continue;
        }
        const uint4* const wrow[2] = {reinterpret_cast<const uint4*>(wbase + (size_t)(2 * j) * row_bytes),
                                      reinterpret_cast<const uint4*>(wbase + (size_t)(2 * j + 1) * row_bytes)};
        float acc[2];
        gemv_unit<BT, 2>(c, wrow, xs, K, acc);
        if (c.lane >= nb) continue;
        const int b = c.lane;
        unsigned long long* out = out0 + (size_t)b * kLLRep * len_out + u;
        if (kind == PH_QKV) {
            const int pos = fast ? ph.depth_pos : S.pos[b];
            float v0 = bf16_round(acc[0]), v1 = bf16_round(acc[1]);
            if (n0 < k_end) {  // q and k rows: interleaved-pair RoPE with the bf16 table (P:616-640)
                const int jj = (n0 & (kHeadDim - 1)) >> 1;
                const uint32_t cs = __ldg(reinterpret_cast<const uint32_t*>(table + ((size_t)pos * (kHeadDim / 2) + jj) * 2));
                const float co = bf_lo(cs), si = bf_hi(cs);
                const float r0 = bf16_round(__fsub_rn(__fmul_rn(v0, co), __fmul_rn(v1, si)));
                const float r1 = bf16_round(__fadd_rn(__fmul_rn(v1, co), __fmul_rn(v0, si)));
                v0 = r0; v1 = r1;
            }
            const uint32_t packed = pack_bf16(v0, v1);
            ll_put(out, len_out, kLLRep, packed, epoch);
            if (!fast && n0 >= q_rows && seq_active(M, A, c, b)) {  // KV append (K:12-22) into the paged pool
                const int is_v = n0 >= k_end ? 1 : 0;
                const int n1 = n0 - (is_v ? k_end : q_rows);
                const int kvh = n1 / kHeadDim, d = n1 & (kHeadDim - 1);
                const int ps = M.page_size;
                if (pos < A.b.max_pages * ps) {
                    const int page = ldcg_i32(A.b.block_table + (size_t)b * A.b.max_pages + pos / ps);
                    uint16_t* dst = M.kv_pool + ((((size_t)page * M.n_layer + ph.layer) * 2 + is_v) * Hkv + kvh) * ((size_t)ps * kHeadDim)
                                    + (size_t)(pos % ps) * kHeadDim + d;
                    *reinterpret_cast<uint32_t*>(dst) = packed;
                }
            }
        } else if (kind == PH_HEAD) {
            const float l0 = bf16_round(acc[0]), l1 = bf16_round(acc[1]);
            float* dbg = fast ? M.depth_logits + ((size_t)b * M.depth + ph.depth_pos) * n_head_rows : M.token_logits + (size_t)b * n_head_rows;
            *reinterpret_cast<float2*>(dbg + n0) = make_float2(l0, l1);
            ll_put(out, len_out, 1, pack_bf16(l0, l1), epoch);
        } else {
            // wo: h = x + wo(attn)  (P:499)      w2: x' = h + w2(act)  (P:500)
            const uint16_t* res = (kind == PH_WO ? c.res[0] : c.res[1]) + (size_t)b * D;
            const uint32_t rr = *reinterpret_cast<const uint32_t*>(res + n0);
            const float o0 = __fadd_rn(bf_lo(rr), bf16_round(acc[0]));
            const float o1 = __fadd_rn(bf_hi(rr), bf16_round(acc[1]));
            ll_put(out, len_out, kLLRep, pack_bf16(o0, o1), epoch);
        }
    }
    // this warp is done with the stage: hand the ring space back to the producer
    __syncwarp();
    if (c.lane == 0) {
        for (int i = 0; i < n_parts; ++i) mbar_arrive(&S.empty[(wseq + i) % kStages]);
    }
    wseq += n_parts;
    if (prof) { const unsigned long long t = globaltimer_ns(); seg[2] += t - ts; ts = t; }
}

// ---- sampling (G:88-99 slow, G:118-132 depth) + frame assembly and the stop rule (G:143-166) ------------
__device__ __noinline__ void phase_sample(const DevModel& M, const CallArgs& A, Ctx& c, const Phase& ph, int p, uint32_t epoch) {
    Shared& S = *c.sh;
    const bool fast = ph.fast != 0;
    const int N = fast ? M.codebook_size : M.vocab;
    const int r = fast ? 1 + ph.depth_pos : 0;
    const int R = M.n_rows;
    float* lg = c.scratch;
    for (int b = c.cta; b < A.batch; b += c.n_ctas) {
        const unsigned long long* src = M.ll + M.ll_off[p - 1] + (size_t)(b * kLLRep) * M.ll_len[p - 1];  // replica 0
        for (int i = c.tid; i < N / 4; i += kThreads) {
            uint32_t u[2];
            ll_get4<1>(src + 2 * i, epoch - 1, u);
            *reinterpret_cast<float4*>(lg + 4 * i) = make_float4(bf_lo(u[0]), bf_hi(u[0]), bf_lo(u[1]), bf_hi(u[1]));
        }
        csync();
        const float temp = fast ? A.s.fast_temp : A.s.temp;
        const uint32_t seq_id = A.b.seq_id ? (uint32_t)ldcg_i32(A.b.seq_id + b) : (uint32_t)b;
        int tok = sample_row(lg, N, temp, fast ? 0 : A.s.top_k, fast ? 1.0f : A.s.top_p, A.s.min_p, A.s.seed,
                             (uint32_t)S.step[b], seq_id, (uint32_t)r, S.sc);
        if (M.force != nullptr) tok = ldcg_i32(M.force + (size_t)b * R + r);
        if (fast && ph.depth_pos == M.depth - 1) __threadfence();  // frame boundary: keep the release chain cumulative
        if (c.tid < c.n_ctas) st_relaxed_v2(M.ll_tok + ((size_t)b * R + r) * kLLMaxCtas + c.tid, (uint32_t)tok, epoch);
        if (c.tid == 0) {
            M.frame_tokens[(size_t)b * R + r] = tok;
            if (fast && ph.depth_pos == M.depth - 1 && S.fin[b] == 0) {
                // frame assembly for the host and the next launch; every CTA applies the same update locally
                const int st = S.step[b];
                int slow = 0;
                for (int rr = 0; rr < R; ++rr) {
                    int v;
                    if (rr == r) v = tok;
                    else if (rr == 0) v = token_word_get(M, c, b, 0, epoch - (uint32_t)(p - sample_phase_index(M, 0)));
                    else v = S.nw[b][rr];
                    if (rr == 0) slow = v;
                    A.b.tokens[(size_t)b * R + rr] = v;
                    if (A.b.out_codes != nullptr && st < A.b.max_frames)
                        A.b.out_codes[((size_t)b * A.b.max_frames + st) * R + rr] = v;
                }
                if (A.b.step) A.b.step[b] = st + 1;
                A.b.seq_len[b] = S.pos[b] + 1;
                if (A.b.finished != nullptr && A.s.audio_only && !A.s.ignore_stop && slow == M.im_end) A.b.finished[b] = 1;
            }
        }
        csync();
    }
}

struct SmemPlan {
    int xs_bytes;       // one activation buffer
    int res_bytes;      // one residual buffer
    int scratch_bytes;
    int ring_bytes;
};

template <int BT>
__global__ void __launch_bounds__(kLLThreads, 1)
smol_ll_kernel(const __grid_constant__ DevModel M, const __grid_constant__ CallArgs A, const __grid_constant__ SmemPlan SP) {
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ Shared S;
    Ctx c;
    c.cta = blockIdx.x; c.n_ctas = gridDim.x; c.tid = threadIdx.x;
    c.warp = threadIdx.x >> 5; c.lane = threadIdx.x & 31;
    c.rep = c.cta % kLLRep;
    c.sh = &S;
    unsigned char* sp = smem_dyn;
    c.xs[0] = reinterpret_cast<uint16_t*>(sp); sp += SP.xs_bytes;
    c.xs[1] = reinterpret_cast<uint16_t*>(sp); sp += SP.xs_bytes;
    c.res[0] = reinterpret_cast<uint16_t*>(sp); sp += SP.res_bytes;
    c.res[1] = reinterpret_cast<uint16_t*>(sp); sp += SP.res_bytes;
    c.scratch = reinterpret_cast<float*>(sp); sp += SP.scratch_bytes;
    c.ring = sp;
    c.ring_bytes = (uint32_t)SP.ring_bytes;
    c.xcur = 0;
    c.iter = 0;

    if (c.tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&S.full[i], 1u); mbar_init(&S.empty[i], (uint32_t)kWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (c.tid < A.batch) {
        S.pos[c.tid] = ldcg_i32(A.b.seq_len + c.tid);
        S.step[c.tid] = A.b.step ? ldcg_i32(A.b.step + c.tid) : 0;
        S.fin[c.tid] = (A.mode == 0 && A.b.finished != nullptr) ? (int)__ldcg(A.b.finished + c.tid) : 0;
    }
    if (A.mode == 0 && c.tid < A.batch * M.n_rows) {
        const int b = c.tid / M.n_rows, r = c.tid - b * M.n_rows;
        S.tok[b][r] = ldcg_i32(A.b.tokens + c.tid);
        S.nw[b][r] = 0;
    }
    __syncthreads();

    if (c.warp == kWarps) {  // the TMA producer warp
        if (c.lane == 0) producer(M, A, c);
        return;
    }

    const uint32_t epoch0 = (uint32_t)ldcg_i32(M.ll_epoch);
    const int per_iter = A.phase_end - A.phase_begin;
    const bool prof = (M.prof != nullptr) && c.cta == 0 && c.tid == 0;
    uint32_t wseq = 0, ring_head = 0;
    unsigned long long t0 = 0;
    for (int it = 0; it < A.n_iter; ++it) {
        c.iter = it;
        for (int p = A.phase_begin; p < A.phase_end; ++p) {
            const Phase ph = unpack_phase(M.prog[p]);
            const uint32_t epoch = epoch0 + (uint32_t)it * (uint32_t)per_iter + (uint32_t)p + 1u;
            if (prof) t0 = globaltimer_ns();
            if (ph.kind == PH_ATTN) phase_attn(M, A, c, ph, p, epoch);
            else if (ph.kind == PH_SAMPLE) phase_sample(M, A, c, ph, p, epoch);
            else phase_gemv<BT>(M, A, c, ph, p, epoch, wseq, ring_head, per_iter);
            if (prof) M.prof[2 * p] += globaltimer_ns() - t0;
        }
        if (A.mode == 1) {  // prefill: sequences still inside their prompt advance one position
            csync();
            if (c.tid < A.batch && seq_active(M, A, c, c.tid)) S.pos[c.tid] += 1;
            csync();
        }
    }
    if (c.cta == 0) {
        if (c.tid == 0) *M.ll_epoch = epoch0 + (uint32_t)A.n_iter * (uint32_t)per_iter;
        if (A.mode == 1) {
            if (c.tid < A.batch) A.b.seq_len[c.tid] = S.pos[c.tid];
            if (A.finalize) {
                // leave the last prompt column as the pending input of the first decode frame (G:66-73)
                for (int i = c.tid; i < A.batch * M.n_rows; i += kThreads) {
                    const int b = i / M.n_rows, r = i - b * M.n_rows;
                    const int len = ldcg_i32(A.prompt_len + b);
                    A.b.tokens[i] = ldcg_i32(A.prompt + ((size_t)b * M.n_rows + r) * A.s_max + (len - 1));
                }
            }
        }
    }
}

}  // namespace ll

// ---- host-side launch helpers (called from capi.cu) ------------------------------------------------
static const void* ll_fn(int bt) {
    switch (bt) {
        case 1: return (const void*)ll::smol_ll_kernel<1>;
        case 2: return (const void*)ll::smol_ll_kernel<2>;
        case 4: return (const void*)ll::smol_ll_kernel<4>;
        default: return (const void*)ll::smol_ll_kernel<8>;
    }
}
static int ll_tile_index(int bt) { return bt <= 1 ? 0 : bt <= 2 ? 1 : bt <= 4 ? 2 : 3; }

// Shared-memory plan of the data-flow kernel for batch tile bt; returns the dynamic bytes (0 = does not fit).
size_t ll_smem_plan(const DevModel& M, int bt, int n_ctas, int* xs_bytes, int* res_bytes, int* scratch_bytes, int* ring_bytes) {
    int kmax = M.dim;
    if (M.inter > kmax) kmax = M.inter;
    if (M.fdim > kmax) kmax = M.fdim;
    if (M.finter > kmax) kmax = M.finter;
    const int dmax = M.dim > M.fdim ? M.dim : M.fdim;
    auto up = [](size_t v) { return (int)((v + 127) / 128 * 128); };
    *xs_bytes = up((size_t)bt * kmax * 2);
    *res_bytes = up((size_t)bt * dmax * 2);
    const int nl = M.vocab > M.codebook_size ? M.vocab : M.codebook_size;
    size_t sc = (size_t)kWarps * kMaxGroup * kPartialStride * sizeof(float);
    if ((size_t)nl * sizeof(float) > sc) sc = (size_t)nl * sizeof(float);
    *scratch_bytes = up(sc);
    const size_t fixed = 2 * (size_t)*xs_bytes + 2 * (size_t)*res_bytes + (size_t)*scratch_bytes;
    const size_t budget = 227 * 1024 - 3072;  // static shared (barriers, sampler scratch, state) stays below 3 KB
    // heaviest phase of one CTA
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
    // the ring allocator is a pure function of the stage sizes (producer and consumers replay it
    // independently); it cannot stall forever as long as two of the largest stages fit
    if (fixed + 2 * need > budget) return 0;
    *ring_bytes = (int)((budget - fixed) / 128 * 128);
    return fixed + (size_t)*ring_bytes;
}

static size_t g_ll_smem_configured[4] = {0, 0, 0, 0};
cudaError_t ll_configure(int bt, size_t smem) {
    size_t& cur = g_ll_smem_configured[ll_tile_index(bt)];
    if (smem <= cur) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(ll_fn(bt), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) cur = smem;
    return e;
}
cudaError_t ll_max_ctas(int bt, size_t smem, int* per_sm) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, ll_fn(bt), kLLThreads, smem);
}
cudaError_t ll_launch(const DevModel& M, const CallArgs& A, int bt, int n_ctas, size_t smem, int xs_bytes, int res_bytes,
                      int scratch_bytes, int ring_bytes, cudaStream_t stream) {
    ll::SmemPlan sp;
    sp.xs_bytes = xs_bytes; sp.res_bytes = res_bytes; sp.scratch_bytes = scratch_bytes; sp.ring_bytes = ring_bytes;
    void* args[3] = {(void*)&M, (void*)&A, (void*)&sp};
    // cooperative launch only for the co-residency guarantee: CTAs spin on each other's words
    return cudaLaunchCooperativeKernel(ll_fn(bt), dim3(n_ctas), dim3(kLLThreads), args, smem, stream);
}

}  // namespace smol
