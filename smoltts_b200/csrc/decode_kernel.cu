// The DualAR decode step as ONE persistent kernel (sm_100a).
//
// One CTA per SM (148 on B200), 16 warps each.  A frame is a fixed program of phases
// (dev_model.h: decode_phase); every CTA walks the same program and the CTAs meet at a
// release/acquire grid barrier between phases, so one cooperative launch runs any number of
// frames (or prompt positions) without returning to the host:
//
//   slow layer l :  QKV   [embed ->] RMSNorm -> wqkv GEMV -> RoPE -> paged KV append      (P:205-221,535-552,601-640; K:12-22)
//                   ATTN  split-KV decode attention over the paged cache, GQA             (P:556-566)
//                   WO    wo GEMV + residual                                               (P:570,499)
//                   W13   RMSNorm -> w1/w3 GEMV -> silu(a)*b                                (P:581-582)
//                   W2    w2 GEMV + residual                                               (P:582,500)
//   HEAD / SAMPLE   final norm -> tied LM head -> argmax | top-k/top-p/temperature          (P:250-255; G:88-99)
//   depth step i :  per fast layer QKV, WO (attention over <= depth on-chip positions fused in),
//                   W13, W2; then HEAD (fast_norm + W_i) and SAMPLE                       (P:409-448,585-598; M:194-220; G:110-141)
//
// (P = modeling/model/rq_transformer.py, M = mlx lm/rq_transformer.py, G = mlx lm/generate.py,
//  K = mlx lm/cache.py of the reference.)
//
// All matrix work at this batch size is weight streaming: each warp owns whole weight rows, issues
// all of a row's 16-byte loads up front (ld.global.nc, no L1 allocation) and reduces with
// shuffles; activations of up to 8 sequences sit in shared memory as fp32.  Rounding points follow
// the reference's eager bf16 forward exactly (bf16 after every Linear, each RMSNorm stage, RoPE,
// SDPA, silu, the product and each residual add; fp32 inside), so only summation order differs.
//
// The same kernel launched non-cooperatively runs exactly one phase (phase_end == phase_begin+1):
// that is the "one kernel per op" mode used for CUDA-graph capture, per-phase profiling and tests.

#include "common.cuh"
#include "dev_model.h"
#include "sampler.cuh"

namespace smol {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldcg_v4(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint32_t ldcg_u32(const void* p) { return __ldcg(reinterpret_cast<const unsigned int*>(p)); }
__device__ __forceinline__ int ldcg_i32(const void* p) { return __ldcg(reinterpret_cast<const int*>(p)); }
__device__ __forceinline__ float ldcg_f32(const void* p) { return __ldcg(reinterpret_cast<const float*>(p)); }
__device__ __forceinline__ float bf_to_f(uint16_t v) { return __uint_as_float(((uint32_t)v) << 16); }
__device__ __forceinline__ uint16_t f_to_bf(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ uint16_t ldcg_u16(const uint16_t* p) {
    return __ldcg(reinterpret_cast<const unsigned short*>(p));
}

// Position of element k of a K-vector inside the shared-memory activation row.  Elements are kept
// as two half-rows (first / second float4 of every 8-element chunk) so that the 16-byte reads of
// 32 consecutive lanes are contiguous and bank-conflict free.
__device__ __forceinline__ int xs_index(int k, int K) {
    const int c = k >> 3, e = k & 7;
    return (e < 4) ? (c * 4 + e) : ((K >> 1) + c * 4 + (e - 4));
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    f[0] = bf_lo(v.x); f[1] = bf_hi(v.x); f[2] = bf_lo(v.y); f[3] = bf_hi(v.y);
    f[4] = bf_lo(v.z); f[5] = bf_hi(v.z); f[6] = bf_lo(v.w); f[7] = bf_hi(v.w);
}

__device__ __forceinline__ void store_chunk(float* row, int K, int c, const float (&f)[8]) {
    *reinterpret_cast<float4*>(row + c * 4) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(row + (K >> 1) + c * 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// Release/acquire grid barrier.  All CTAs are co-resident (cooperative launch).  `target` is the
// cumulative arrival count this barrier completes at; the counter only ever grows (wrap-safe compare).
__device__ __forceinline__ void grid_barrier(uint32_t* ctr, uint32_t& target, uint32_t n_ctas) {
    target += n_ctas;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
        uint32_t v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while ((int32_t)(v - target) < 0);
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct Ctx {
    int cta, n_ctas, warp, lane, tid;
    float* xs;          // dynamic shared memory
    int iter;           // frame (decode) or prompt position (prefill) inside this launch
};

__shared__ SampleScratch g_sc;
__shared__ float g_red[kWarps * kBatchTile];
__shared__ float g_rstd[kBatchTile];
__shared__ int g_pos[kBatchTile];
__shared__ int g_flag;

// ------------------------------------------------------------------------------------------------
// activation staging
// ------------------------------------------------------------------------------------------------

// xs[b][:] = rows[b][0:K] (bf16 in global, read through L2), b < nb.
template <class RowPtr>
__device__ __forceinline__ void fill_rows(const Ctx& c, int K, int nb, RowPtr rowptr) {
    const int nch = K >> 3;
    for (int idx = c.tid; idx < nb * nch; idx += kThreads) {
        const int b = idx / nch, ch = idx - b * nch;
        const uint16_t* src = rowptr(b);
        float f[8];
        unpack8(ldcg_v4(src + ch * 8), f);
        store_chunk(c.xs + (size_t)b * K, K, ch, f);
    }
}

// In-place RMSNorm of xs rows (reference RMSNorm.forward :607-613): fp32 normalise, round to
// bf16, multiply by the bf16 weight, round again.  Every CTA does this redundantly for the
// sequences of the current batch tile.
__device__ __forceinline__ void rmsnorm_rows(const Ctx& c, int K, int nb, const uint16_t* w, float eps) {
    float ss[kBatchTile];
#pragma unroll
    for (int b = 0; b < kBatchTile; ++b) ss[b] = 0.f;
    for (int i = c.tid; i < K; i += kThreads) {
#pragma unroll
        for (int b = 0; b < kBatchTile; ++b)
            if (b < nb) { const float v = c.xs[(size_t)b * K + i]; ss[b] = fmaf(v, v, ss[b]); }
    }
#pragma unroll
    for (int b = 0; b < kBatchTile; ++b) ss[b] = warp_sum(ss[b]);
    if (c.lane == 0) {
#pragma unroll
        for (int b = 0; b < kBatchTile; ++b) g_red[c.warp * kBatchTile + b] = ss[b];
    }
    __syncthreads();
    if (c.tid < nb) {
        float t = 0.f;
        for (int wi = 0; wi < kWarps; ++wi) t += g_red[wi * kBatchTile + c.tid];
        const float mean = __fdiv_rn(t, (float)K);
        g_rstd[c.tid] = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, eps)));
    }
    __syncthreads();
    const int nch = K >> 3;
    for (int idx = c.tid; idx < nch * 2; idx += kThreads) {
        // idx walks the physical layout: first half-row then second half-row, 4 elements at a time
        const int half = idx / nch, ch = idx - half * nch;
        const int k0 = ch * 8 + half * 4;
        const uint2 wv = __ldg(reinterpret_cast<const uint2*>(w + k0));
        const float w0 = bf_lo(wv.x), w1 = bf_hi(wv.x), w2 = bf_lo(wv.y), w3 = bf_hi(wv.y);
#pragma unroll
        for (int b = 0; b < kBatchTile; ++b) {
            if (b < nb) {
                float4* p = reinterpret_cast<float4*>(c.xs + (size_t)b * K + half * (K >> 1) + ch * 4);
                float4 v = *p;
                const float r = g_rstd[b];
                v.x = bf16_round(__fmul_rn(bf16_round(__fmul_rn(v.x, r)), w0));
                v.y = bf16_round(__fmul_rn(bf16_round(__fmul_rn(v.y, r)), w1));
                v.z = bf16_round(__fmul_rn(bf16_round(__fmul_rn(v.z, r)), w2));
                v.w = bf16_round(__fmul_rn(bf16_round(__fmul_rn(v.w, r)), w3));
                *p = v;
            }
        }
    }
    __syncthreads();
}

// Copies the raw (pre-norm) xs rows to a global bf16 buffer [.. + b][K] (only CTA 0 calls it).
__device__ __forceinline__ void spill_rows(const Ctx& c, int K, int nb, uint16_t* dst, int b0) {
    for (int idx = c.tid; idx < nb * (K >> 1); idx += kThreads) {
        const int b = idx / (K >> 1), k = (idx - b * (K >> 1)) * 2;
        const float lo = c.xs[(size_t)b * K + xs_index(k, K)];
        const float hi = c.xs[(size_t)b * K + xs_index(k + 1, K)];
        *reinterpret_cast<uint32_t*>(dst + (size_t)(b0 + b) * K + k) = pack_bf16(lo, hi);
    }
}

// ------------------------------------------------------------------------------------------------
// weight-streaming GEMV over the units owned by this CTA
// ------------------------------------------------------------------------------------------------
// unit u -> R weight rows of `nchunks` 16-byte chunks.  Units are dealt round-robin to CTAs, and a
// CTA's units round-robin to its warps.  acc[r][b] = sum_k W_r[k] * xs[b][k], reduced over the warp.
template <int BT, int NCH, int R, class RowFn, class Epi>
__device__ __forceinline__ void gemv_units(const Ctx& c, int n_units, int K, RowFn rows, Epi epi) {
    const int nchunks = K >> 3;
    const int half = K >> 1;
    for (int u = c.cta + c.warp * c.n_ctas; u < n_units; u += kWarps * c.n_ctas) {
        const uint16_t* rp[R];
#pragma unroll
        for (int r = 0; r < R; ++r) rp[r] = rows(u, r);
        uint4 wv[R][NCH];
#pragma unroll
        for (int it = 0; it < NCH; ++it) {
            const int ch = c.lane + 32 * it;
#pragma unroll
            for (int r = 0; r < R; ++r)
                wv[r][it] = (ch < nchunks) ? ldg_stream(rp[r] + ch * 8) : make_uint4(0u, 0u, 0u, 0u);
        }
        float acc[R][BT];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int b = 0; b < BT; ++b) acc[r][b] = 0.f;
#pragma unroll
        for (int it = 0; it < NCH; ++it) {
            const int ch = c.lane + 32 * it;
            if (ch < nchunks) {
                float wf[R][8];
#pragma unroll
                for (int r = 0; r < R; ++r) unpack8(wv[r][it], wf[r]);
#pragma unroll
                for (int b = 0; b < BT; ++b) {
                    const float* xr = c.xs + (size_t)b * K + ch * 4;
                    const float4 x0 = *reinterpret_cast<const float4*>(xr);
                    const float4 x1 = *reinterpret_cast<const float4*>(xr + half);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float a = acc[r][b];
                        a = fmaf(wf[r][0], x0.x, a); a = fmaf(wf[r][1], x0.y, a);
                        a = fmaf(wf[r][2], x0.z, a); a = fmaf(wf[r][3], x0.w, a);
                        a = fmaf(wf[r][4], x1.x, a); a = fmaf(wf[r][5], x1.y, a);
                        a = fmaf(wf[r][6], x1.z, a); a = fmaf(wf[r][7], x1.w, a);
                        acc[r][b] = a;
                    }
                }
            }
        }
        float mine[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            mine[r] = 0.f;
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                const float s = warp_sum(acc[r][b]);
                if (c.lane == b) mine[r] = s;
            }
        }
        epi(u, mine);  // lane b holds the sums of sequence b of the tile
    }
}

template <int R, class RowFn, class Epi>
__device__ __forceinline__ void gemv_dispatch(const Ctx& c, int nb, int n_units, int K, RowFn rows, Epi epi) {
    const int need = ((K >> 3) + 31) / 32;
#define SMOL_GEMV_BT(NCH)                                                                 \
    do {                                                                                  \
        if (nb == 1) gemv_units<1, NCH, R>(c, n_units, K, rows, epi);                     \
        else if (nb == 2) gemv_units<2, NCH, R>(c, n_units, K, rows, epi);                \
        else if (nb <= 4) gemv_units<4, NCH, R>(c, n_units, K, rows, epi);                \
        else gemv_units<8, NCH, R>(c, n_units, K, rows, epi);                             \
    } while (0)
    if (R == 2) {
        if (need <= 1) SMOL_GEMV_BT(1);
        else if (need <= 3) SMOL_GEMV_BT(3);
        else SMOL_GEMV_BT(6);
    } else {
        if (need <= 1) SMOL_GEMV_BT(1);
        else if (need <= 3) SMOL_GEMV_BT(3);
        else if (need <= 6) SMOL_GEMV_BT(6);
        else SMOL_GEMV_BT(12);
    }
#undef SMOL_GEMV_BT
}

// ------------------------------------------------------------------------------------------------
// phases
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool seq_active(const DevModel& M, const CallArgs& A, const Ctx& c, int bg) {
    if (A.mode == 1) return c.iter + A.iter_base < ldcg_i32(A.prompt_len + bg) - 1;
    return A.b.finished == nullptr || __ldcg(A.b.finished + bg) == 0;
}

// id of grid row r of sequence bg for the slow embedding
__device__ __forceinline__ int input_token(const DevModel& M, const CallArgs& A, const Ctx& c, int bg, int r) {
    if (A.mode == 1) {
        int t = c.iter + A.iter_base;
        const int len = ldcg_i32(A.prompt_len + bg);
        if (t > len - 1) t = len - 1;
        return ldcg_i32(A.prompt + ((size_t)bg * M.n_rows + r) * A.s_max + t);
    }
    return ldcg_i32(A.b.tokens + (size_t)bg * M.n_rows + r);
}

// BaseTransformer.embed (P:205-221): text row + sum of the codebook rows, zeroed by the
// PyTorch rule (row-1 code == 0) or the MLX rule (row-0 id outside the semantic range, M:162-169).
__device__ __forceinline__ void embed_rows(const DevModel& M, const CallArgs& A, const Ctx& c, int b0, int nb) {
    const int D = M.dim, nch = D >> 3;
    for (int idx = c.tid; idx < nb * nch; idx += kThreads) {
        const int b = idx / nch, ch = idx - b * nch, bg = b0 + b;
        const int t0 = input_token(M, A, c, bg, 0);
        float f[8];
        unpack8(ldcg_v4(M.embeddings + (size_t)t0 * D + ch * 8), f);
        bool use_vq;
        if (M.mlx_embed_mask) use_vq = (t0 >= M.semantic_start && t0 <= M.semantic_end);
        else use_vq = input_token(M, A, c, bg, 1) != 0;
        if (use_vq) {
            float s[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) s[e] = 0.f;
            for (int r = 1; r < M.n_rows; ++r) {
                const int code = input_token(M, A, c, bg, r);
                const int row = code + (M.dup0 ? (r - 1) : r) * M.codebook_size;
                float g[8];
                unpack8(ldcg_v4(M.codebook_embeddings + (size_t)row * D + ch * 8), g);
#pragma unroll
                for (int e = 0; e < 8; ++e) s[e] = __fadd_rn(s[e], g[e]);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = bf16_round(__fadd_rn(f[e], bf16_round(s[e])));
        }
        store_chunk(c.xs + (size_t)b * D, D, ch, f);
    }
}

// Attention of the fast transformer for the current batch tile, computed redundantly by every CTA
// straight into xs (the input of the wo GEMV): <= depth cached positions per sequence.
__device__ __forceinline__ void fast_attention_rows(const DevModel& M, const Ctx& c, int layer, int depth_pos,
                                                    int b0, int nb) {
    const int Hq = M.fn_head, Hkv = M.fn_kv, G = Hq / Hkv, D = M.fdim;
    const int kvw = Hkv * kHeadDim;
    for (int pair = c.warp; pair < nb * Hq; pair += kWarps) {
        const int b = pair / Hq, hq = pair - b * Hq, bg = b0 + b, kvh = hq / G;
        const uint32_t qp = ldcg_u32(M.q + (size_t)bg * Hq * kHeadDim + hq * kHeadDim + 2 * c.lane);
        const float q0 = bf_lo(qp), q1 = bf_hi(qp);
        const uint16_t* kb = M.fkv + ((size_t)(bg * M.n_flayer + layer) * 2) * M.depth * kvw + kvh * kHeadDim + 2 * c.lane;
        const uint16_t* vb = kb + (size_t)M.depth * kvw;
        // <= kMaxDepth positions: scores first, then one softmax with the global max.  The probabilities
        // are rounded to bf16 before the PV product and the row sum keeps the unrounded values, which is
        // what the reference's fused SDPA does for bf16 inputs (flash kernels, CPU and CUDA alike).
        float sj[kMaxDepth];
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < kMaxDepth; ++j) {
            sj[j] = -INFINITY;
            if (j <= depth_pos) {
                const uint32_t kp = ldcg_u32(kb + (size_t)j * kvw);
                float s = fmaf(q0, bf_lo(kp), q1 * bf_hi(kp));
                s = warp_sum(s) * 0.125f;
                sj[j] = s;
                m = fmaxf(m, s);
            }
        }
        float l = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxDepth; ++j) {
            if (j <= depth_pos) {
                const uint32_t vp = ldcg_u32(vb + (size_t)j * kvw);
                const float pe = expf(sj[j] - m);
                l += pe;
                const float pb = bf16_round(pe);
                o0 = fmaf(pb, bf_lo(vp), o0);
                o1 = fmaf(pb, bf_hi(vp), o1);
            }
        }
        const float inv = 1.0f / l;
        float* row = c.xs + (size_t)b * D;
        const int k = hq * kHeadDim + 2 * c.lane;
        row[xs_index(k, D)] = bf16_round(o0 * inv);
        row[xs_index(k + 1, D)] = bf16_round(o1 * inv);
    }
}

__device__ void phase_qkv(const DevModel& M, const CallArgs& A, const Ctx& c, const Phase& ph) {
    const bool fast = ph.fast != 0;
    const int D = fast ? M.fdim : M.dim;
    const int Hq = fast ? M.fn_head : M.n_head, Hkv = fast ? M.fn_kv : M.n_kv;
    const DevLayer& L = fast ? M.fast_layers[ph.layer] : M.layers[ph.layer];
    const int q_rows = Hq * kHeadDim, k_end = (Hq + Hkv) * kHeadDim;
    const int n_units = (Hq + 2 * Hkv) * kHeadDim / 2;
    const uint16_t* table = fast ? M.fast_rope : M.rope;
    for (int b0 = 0; b0 < A.batch; b0 += kBatchTile) {
        const int nb = min(kBatchTile, A.batch - b0);
        if (!fast && ph.layer == 0) {
            embed_rows(M, A, c, b0, nb);
            __syncthreads();
            if (c.cta == 0) spill_rows(c, D, nb, M.x, b0);
        } else if (fast && ph.layer == 0 && !A.fast_from_xf) {
            if (ph.depth_pos == 0) {
                fill_rows(c, D, nb, [&](int b) { return M.x + (size_t)(b0 + b) * D; });
            } else {
                fill_rows(c, D, nb, [&](int b) {
                    const int code = ldcg_i32(M.frame_tokens + (size_t)(b0 + b) * M.n_rows + ph.depth_pos);
                    const int off = M.depthwise_wte ? (M.dup0 ? ph.depth_pos - 1 : ph.depth_pos) * M.codebook_size : 0;
                    return M.fast_embeddings + (size_t)(code + off) * D;
                });
            }
            __syncthreads();
            if (c.cta == 0) spill_rows(c, D, nb, M.xf, b0);
        } else {
            const uint16_t* src = fast ? M.xf : M.x;
            fill_rows(c, D, nb, [&](int b) { return src + (size_t)(b0 + b) * D; });
        }
        if (c.tid < nb) g_pos[c.tid] = fast ? ph.depth_pos : ldcg_i32(A.b.seq_len + b0 + c.tid);
        __syncthreads();
        rmsnorm_rows(c, D, nb, L.attention_norm, M.eps);

        auto rows = [&](int u, int r) { return L.wqkv + (size_t)(2 * u + r) * D; };
        auto epi = [&](int u, const float (&acc)[2]) {
            if (c.lane >= nb) return;
            const int bg = b0 + c.lane, n0 = 2 * u, pos = g_pos[c.lane];
            float v0 = bf16_round(acc[0]), v1 = bf16_round(acc[1]);
            if (n0 < k_end) {  // q and k rows: interleaved-pair RoPE with the bf16 table (P:616-640)
                const int j = (n0 & (kHeadDim - 1)) >> 1;
                const uint32_t cs = __ldg(reinterpret_cast<const uint32_t*>(table + ((size_t)pos * (kHeadDim / 2) + j) * 2));
                const float co = bf_lo(cs), si = bf_hi(cs);
                const float r0 = bf16_round(__fsub_rn(__fmul_rn(v0, co), __fmul_rn(v1, si)));
                const float r1 = bf16_round(__fadd_rn(__fmul_rn(v1, co), __fmul_rn(v0, si)));
                v0 = r0; v1 = r1;
            }
            const uint32_t packed = pack_bf16(v0, v1);
            if (n0 < q_rows) {
                *reinterpret_cast<uint32_t*>(M.q + (size_t)bg * q_rows + n0) = packed;
                return;
            }
            const int is_v = n0 >= k_end ? 1 : 0;
            const int n1 = n0 - (is_v ? k_end : q_rows);
            const int kvh = n1 / kHeadDim, d = n1 & (kHeadDim - 1);
            if (fast) {
                uint16_t* dst = M.fkv + (((size_t)(bg * M.n_flayer + ph.layer) * 2 + is_v) * M.depth + ph.depth_pos) * (Hkv * kHeadDim)
                                + kvh * kHeadDim + d;
                *reinterpret_cast<uint32_t*>(dst) = packed;
            } else {
                if (!seq_active(M, A, c, bg)) return;
                const int ps = M.page_size;
                if (pos >= A.b.max_pages * ps) return;
                const int page = ldcg_i32(A.b.block_table + (size_t)bg * A.b.max_pages + pos / ps);
                uint16_t* dst = M.kv_pool + ((((size_t)page * M.n_layer + ph.layer) * 2 + is_v) * Hkv + kvh) * ((size_t)ps * kHeadDim)
                                + (size_t)(pos % ps) * kHeadDim + d;
                *reinterpret_cast<uint32_t*>(dst) = packed;
            }
        };
        gemv_dispatch<2>(c, nb, n_units, D, rows, epi);
        __syncthreads();
    }
}

// Split-KV decode attention over the paged cache.  Unit = (sequence, kv head, split); the G query
// heads of a kv head share every K/V read.  Each 8-lane group owns one cached position per step
// and keeps its own online-softmax state; states are merged across groups, warps and (through a
// last-arriver fix-up in global memory) splits, always in a fixed order.
template <int G>
__device__ void phase_attn_g(const DevModel& M, const CallArgs& A, const Ctx& c, const Phase& ph, int s_max) {
    const int Hq = M.n_head, Hkv = M.n_kv, ps = M.page_size;
    const int dch = c.lane & 7, psub = c.lane >> 3;
    float* red = c.xs;  // [kWarps][G][kPartialStride]
    const int n_units = A.batch * Hkv * s_max;
    for (int u = c.cta; u < n_units; u += c.n_ctas) {
        const int s = u % s_max, kvh = (u / s_max) % Hkv, b = u / (s_max * Hkv);
        int Lb = ldcg_i32(A.b.seq_len + b) + 1;
        const int cap = A.b.max_pages * ps;
        if (Lb > cap) Lb = cap;
        int ns = (Lb + kSplitMin - 1) / kSplitMin;
        if (ns > s_max) ns = s_max;
        if (ns < 1) ns = 1;
        if (s >= ns) continue;
        const int chunk = (Lb + ns - 1) / ns;
        const int p0 = s * chunk, p1 = min(Lb, p0 + chunk);

        float qf[G][8];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            unpack8(ldcg_v4(M.q + (size_t)b * Hq * kHeadDim + (kvh * G + g) * kHeadDim + dch * 8), qf[g]);
#pragma unroll
            for (int e = 0; e < 8; ++e) qf[g][e] *= 0.125f;  // 1/sqrt(64), exact
        }
        float m[G], l[G], acc[G][8];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            m[g] = -INFINITY; l[g] = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[g][e] = 0.f;
        }
        const int32_t* bt = A.b.block_table + (size_t)b * A.b.max_pages;
        const size_t head_stride = (size_t)ps * kHeadDim;
        for (int pb = p0 + c.warp * 4; pb < p1; pb += kWarps * 4) {
            const int p = pb + psub;
            const bool valid = p < p1;
            uint4 kk = make_uint4(0u, 0u, 0u, 0u), vv = kk;
            if (valid) {
                const int page = ldcg_i32(bt + p / ps);
                const uint16_t* kp = M.kv_pool + ((((size_t)page * M.n_layer + ph.layer) * 2) * Hkv + kvh) * head_stride
                                     + (size_t)(p % ps) * kHeadDim + dch * 8;
                kk = ldcg_v4(kp);
                vv = ldcg_v4(kp + (size_t)Hkv * head_stride);
            }
            float kf[8], vf[8];
            unpack8(kk, kf);
            unpack8(vv, vf);
#pragma unroll
            for (int g = 0; g < G; ++g) {
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
            for (int g = 0; g < G; ++g) {
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
            for (int g = 0; g < G; ++g) {
                float* dst = red + (c.warp * G + g) * kPartialStride;
                if (dch == 0) { dst[0] = m[g]; dst[1] = l[g]; }
#pragma unroll
                for (int e = 0; e < 8; ++e) dst[2 + dch * 8 + e] = acc[g][e];
            }
        }
        __syncthreads();
        // merge the warps (fixed order)
        float Mx = -INFINITY, Ls = 0.f, Os = 0.f;
        const int g = c.tid / kHeadDim, d = c.tid & (kHeadDim - 1);
        if (c.tid < G * kHeadDim) {
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
            if (c.tid < G * kHeadDim) M.attn[(size_t)b * Hq * kHeadDim + hq * kHeadDim + d] = f_to_bf(Os / Ls);
        } else {
            if (c.tid < G * kHeadDim) {
                float* dst = M.partial + (((size_t)b * Hq + hq) * kMaxSplits + s) * kPartialStride;
                if (d == 0) { dst[0] = Mx; dst[1] = Ls; }
                dst[2 + d] = Os;
            }
            __threadfence();
            __syncthreads();
            if (c.tid == 0) {
                const uint32_t old = atomicAdd(M.split_count + (size_t)b * Hkv + kvh, 1u);
                g_flag = (old == (uint32_t)(ns - 1)) ? 1 : 0;
            }
            __syncthreads();
            if (g_flag) {  // last split of this (sequence, kv head): combine all splits in split order
                __threadfence();
                if (c.tid < G * kHeadDim) {
                    const float* base = M.partial + ((size_t)b * Hq + hq) * kMaxSplits * kPartialStride;
                    float Mg = -INFINITY;
                    for (int si = 0; si < ns; ++si) Mg = fmaxf(Mg, ldcg_f32(base + si * kPartialStride));
                    float Lg = 0.f, Og = 0.f;
                    for (int si = 0; si < ns; ++si) {
                        const float* src = base + si * kPartialStride;
                        const float ms = ldcg_f32(src);
                        if (ms == -INFINITY) continue;
                        const float sc = expf(ms - Mg);
                        Lg = fmaf(ldcg_f32(src + 1), sc, Lg);
                        Og = fmaf(ldcg_f32(src + 2 + d), sc, Og);
                    }
                    M.attn[(size_t)b * Hq * kHeadDim + hq * kHeadDim + d] = f_to_bf(Og / Lg);
                }
                if (c.tid == 0) M.split_count[(size_t)b * Hkv + kvh] = 0u;
            }
        }
        __syncthreads();
    }
}

__device__ int attn_splits(const DevModel& M, int batch, int n_ctas) {
    const int pairs = batch * M.n_kv;
    int s = (2 * n_ctas + pairs - 1) / pairs;
    if (s < 1) s = 1;
    if (s > kMaxSplits) s = kMaxSplits;
    return s;
}

__device__ void phase_attn(const DevModel& M, const CallArgs& A, const Ctx& c, const Phase& ph) {
    const int s_max = attn_splits(M, A.batch, c.n_ctas);
    switch (M.n_head / M.n_kv) {
        case 1: phase_attn_g<1>(M, A, c, ph, s_max); break;
        case 2: phase_attn_g<2>(M, A, c, ph, s_max); break;
        case 3: phase_attn_g<3>(M, A, c, ph, s_max); break;
        default: phase_attn_g<4>(M, A, c, ph, s_max); break;
    }
}

__device__ void phase_wo(const DevModel& M, const CallArgs& A, const Ctx& c, const Phase& ph) {
    const bool fast = ph.fast != 0;
    const int D = fast ? M.fdim : M.dim;
    const DevLayer& L = fast ? M.fast_layers[ph.layer] : M.layers[ph.layer];
    const uint16_t* res = fast ? M.xf : M.x;
    for (int b0 = 0; b0 < A.batch; b0 += kBatchTile) {
        const int nb = min(kBatchTile, A.batch - b0);
        if (fast) fast_attention_rows(M, c, ph.layer, ph.depth_pos, b0, nb);
        else fill_rows(c, D, nb, [&](int b) { return M.attn + (size_t)(b0 + b) * D; });
        __syncthreads();
        auto rows = [&](int u, int) { return L.wo + (size_t)u * D; };
        auto epi = [&](int u, const float (&acc)[1]) {
            if (c.lane >= nb) return;
            const size_t o = (size_t)(b0 + c.lane) * D + u;
            M.h[o] = f_to_bf(__fadd_rn(bf_to_f(ldcg_u16(res + o)), bf16_round(acc[0])));
        };
        gemv_dispatch<1>(c, nb, D, D, rows, epi);
        __syncthreads();
    }
}

__device__ void phase_w13(const DevModel& M, const CallArgs& A, const Ctx& c, const Phase& ph) {
    const bool fast = ph.fast != 0;
    const int D = fast ? M.fdim : M.dim, F = fast ? M.finter : M.inter;
    const DevLayer& L = fast ? M.fast_layers[ph.layer] : M.layers[ph.layer];
    for (int b0 = 0; b0 < A.batch; b0 += kBatchTile) {
        const int nb = min(kBatchTile, A.batch - b0);
        fill_rows(c, D, nb, [&](int b) { return M.h + (size_t)(b0 + b) * D; });
        __syncthreads();
        rmsnorm_rows(c, D, nb, L.ffn_norm, M.eps);
        auto rows = [&](int u, int r) { return (r == 0 ? L.w1 : L.w3) + (size_t)u * D; };
        auto epi = [&](int u, const float (&acc)[2]) {
            if (c.lane >= nb) return;
            const float a = bf16_round(acc[0]), g = bf16_round(acc[1]);
            const float s = bf16_round(__fdiv_rn(a, __fadd_rn(1.0f, expf(-a))));  // F.silu in fp32, bf16 out
            M.act[(size_t)(b0 + c.lane) * F + u] = f_to_bf(__fmul_rn(s, g));
        };
        gemv_dispatch<2>(c, nb, F, D, rows, epi);
        __syncthreads();
    }
}

__device__ void phase_w2(const DevModel& M, const CallArgs& A, const Ctx& c, const Phase& ph) {
    const bool fast = ph.fast != 0;
    const int D = fast ? M.fdim : M.dim, F = fast ? M.finter : M.inter;
    const DevLayer& L = fast ? M.fast_layers[ph.layer] : M.layers[ph.layer];
    uint16_t* dst = fast ? M.xf : M.x;
    for (int b0 = 0; b0 < A.batch; b0 += kBatchTile) {
        const int nb = min(kBatchTile, A.batch - b0);
        fill_rows(c, F, nb, [&](int b) { return M.act + (size_t)(b0 + b) * F; });
        __syncthreads();
        auto rows = [&](int u, int) { return L.w2 + (size_t)u * F; };
        auto epi = [&](int u, const float (&acc)[1]) {
            if (c.lane >= nb) return;
            const size_t o = (size_t)(b0 + c.lane) * D + u;
            dst[o] = f_to_bf(__fadd_rn(bf_to_f(ldcg_u16(M.h + o)), bf16_round(acc[0])));
        };
        gemv_dispatch<1>(c, nb, D, F, rows, epi);
        __syncthreads();
    }
}

__device__ void phase_head(const DevModel& M, const CallArgs& A, const Ctx& c, const Phase& ph) {
    const bool fast = ph.fast != 0;
    const int D = fast ? M.fdim : M.dim;
    const int N = fast ? M.codebook_size : M.vocab;
    const uint16_t* W = fast ? M.fast_output + (M.depthwise_output ? (size_t)ph.depth_pos * N * D : 0) : M.head;
    const uint16_t* src = fast ? M.xf : M.x;
    for (int b0 = 0; b0 < A.batch; b0 += kBatchTile) {
        const int nb = min(kBatchTile, A.batch - b0);
        fill_rows(c, D, nb, [&](int b) { return src + (size_t)(b0 + b) * D; });
        __syncthreads();
        rmsnorm_rows(c, D, nb, fast ? M.fast_norm : M.norm, M.eps);
        auto rows = [&](int u, int) { return W + (size_t)u * D; };
        auto epi = [&](int u, const float (&acc)[1]) {
            if (c.lane >= nb) return;
            const int bg = b0 + c.lane;
            float* out = fast ? M.depth_logits + ((size_t)bg * M.depth + ph.depth_pos) * N : M.token_logits + (size_t)bg * N;
            out[u] = bf16_round(acc[0]);
        };
        gemv_dispatch<1>(c, nb, N, D, rows, epi);
        __syncthreads();
    }
}

// Sampling of one id per sequence (G:88-99 slow, G:118-132 depth) and, after the last depth code,
// frame assembly and the stop rule (G:143-166).
__device__ void phase_sample(const DevModel& M, const CallArgs& A, const Ctx& c, const Phase& ph) {
    const bool fast = ph.fast != 0;
    const int N = fast ? M.codebook_size : M.vocab;
    const int r = fast ? 1 + ph.depth_pos : 0;
    const int R = M.n_rows;
    for (int b = c.cta; b < A.batch; b += c.n_ctas) {
        const float* src = fast ? M.depth_logits + ((size_t)b * M.depth + ph.depth_pos) * N : M.token_logits + (size_t)b * N;
        for (int i = c.tid; i < N; i += kThreads) c.xs[i] = ldcg_f32(src + i);
        __syncthreads();
        const uint32_t step = A.b.step ? (uint32_t)ldcg_i32(A.b.step + b) : 0u;
        const uint32_t seq_id = A.b.seq_id ? (uint32_t)ldcg_i32(A.b.seq_id + b) : (uint32_t)b;
        const float temp = fast ? A.s.fast_temp : A.s.temp;
        int tok = sample_row(c.xs, N, temp, fast ? 0 : A.s.top_k, fast ? 1.0f : A.s.top_p, A.s.min_p, A.s.seed,
                             step, seq_id, (uint32_t)r, g_sc);
        if (M.force != nullptr) tok = ldcg_i32(M.force + (size_t)b * R + r);
        if (c.tid == 0) {
            M.frame_tokens[(size_t)b * R + r] = tok;
            if (fast && ph.depth_pos == M.depth - 1) {
                const bool live = A.b.finished == nullptr || __ldcg(A.b.finished + b) == 0;
                if (live) {
                    const int st = A.b.step ? ldcg_i32(A.b.step + b) : 0;
                    int slow = 0;
                    for (int rr = 0; rr < R; ++rr) {
                        const int v = (rr == r) ? tok : ldcg_i32(M.frame_tokens + (size_t)b * R + rr);
                        if (rr == 0) slow = v;
                        A.b.tokens[(size_t)b * R + rr] = v;
                        if (A.b.out_codes != nullptr && st < A.b.max_frames)
                            A.b.out_codes[((size_t)b * A.b.max_frames + st) * R + rr] = v;
                    }
                    if (A.b.step) A.b.step[b] = st + 1;
                    A.b.seq_len[b] = ldcg_i32(A.b.seq_len + b) + 1;
                    if (A.b.finished != nullptr && A.s.audio_only && !A.s.ignore_stop && slow == M.im_end)
                        A.b.finished[b] = 1;
                }
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void run_phase(const DevModel& M, const CallArgs& A, const Ctx& c, const Phase& ph) {
    switch (ph.kind) {
        case PH_QKV: phase_qkv(M, A, c, ph); break;
        case PH_ATTN: phase_attn(M, A, c, ph); break;
        case PH_WO: phase_wo(M, A, c, ph); break;
        case PH_W13: phase_w13(M, A, c, ph); break;
        case PH_W2: phase_w2(M, A, c, ph); break;
        case PH_HEAD: phase_head(M, A, c, ph); break;
        default: phase_sample(M, A, c, ph); break;
    }
}

__global__ void __launch_bounds__(kThreads, 1)
smol_decode_kernel(const __grid_constant__ DevModel M, const __grid_constant__ CallArgs A) {
    extern __shared__ __align__(16) float smem_dyn[];
    Ctx c;
    c.cta = blockIdx.x; c.n_ctas = gridDim.x; c.tid = threadIdx.x;
    c.warp = threadIdx.x >> 5; c.lane = threadIdx.x & 31;
    c.xs = smem_dyn;
    c.iter = 0;

    uint32_t target = 0;
    if (A.cooperative) target = ldcg_u32(M.barrier + 1);
    const int per_iter = (A.mode == 1) ? phases_per_prefill_step(M.n_layer)
                                       : phases_per_frame(M.n_layer, M.n_flayer, M.depth);
    const bool prof_cta = (M.prof != nullptr) && c.cta == 0;  // CTA-uniform
    const bool prof = prof_cta && c.tid == 0;
    unsigned long long t0 = 0, t1 = 0;
    for (int it = 0; it < A.n_iter; ++it) {
        c.iter = it;
        for (int p = A.phase_begin; p < A.phase_end; ++p) {
            const Phase ph = decode_phase(p, M.n_layer, M.n_flayer);
            if (prof) t0 = globaltimer_ns();
            run_phase(M, A, c, ph);
            if (prof_cta) {
                __syncthreads();
                if (prof) t1 = globaltimer_ns();
            }
            if (c.cta == 0 && p == per_iter - 1 && A.mode == 1) {
                // prefill bookkeeping: sequences still inside their prompt advance one position
                for (int b = c.tid; b < A.batch; b += kThreads)
                    if (seq_active(M, A, c, b)) A.b.seq_len[b] = ldcg_i32(A.b.seq_len + b) + 1;
            }
            if (c.cta == 0 && A.mode == 0 && A.advance && p == A.phase_end - 1) {
                for (int b = c.tid; b < A.batch; b += kThreads) A.b.seq_len[b] = ldcg_i32(A.b.seq_len + b) + 1;
            }
            const bool last = (it == A.n_iter - 1) && (p == A.phase_end - 1);
            if (A.cooperative && !last) grid_barrier(M.barrier, target, (uint32_t)c.n_ctas);
            if (prof) {  // [2p] CTA 0's own time in the phase, [2p+1] its wait at the barrier that follows
                const unsigned long long t2 = globaltimer_ns();
                M.prof[2 * p] += t1 - t0;
                M.prof[2 * p + 1] += t2 - t1;
            }
        }
    }
    if (A.mode == 1 && A.finalize && c.cta == 0) {
        // leave the last prompt column as the pending input of the first decode frame (G:66-73)
        for (int i = c.tid; i < A.batch * M.n_rows; i += kThreads) {
            const int b = i / M.n_rows, r = i - b * M.n_rows;
            const int len = ldcg_i32(A.prompt_len + b);
            A.b.tokens[i] = ldcg_i32(A.prompt + ((size_t)b * M.n_rows + r) * A.s_max + (len - 1));
        }
    }
    if (A.cooperative && c.cta == 0 && c.tid == 0) M.barrier[1] = target;
}

// Stand-alone sampler on caller-provided logits [B][n] (smol_sample): one CTA per sequence.
__global__ void __launch_bounds__(kThreads, 1)
smol_sample_kernel(const float* logits, int n, int batch, SmolSampling s, int stream_id, const int32_t* seq_id,
                   const int32_t* step, int32_t* out) {
    extern __shared__ __align__(16) float smem_dyn[];
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        for (int i = threadIdx.x; i < n; i += kThreads) smem_dyn[i] = logits[(size_t)b * n + i];
        __syncthreads();
        const bool fast = stream_id != 0;
        const int tok = sample_row(smem_dyn, n, fast ? s.fast_temp : s.temp, fast ? 0 : s.top_k, fast ? 1.0f : s.top_p,
                                   s.min_p, s.seed, step ? (uint32_t)step[b] : 0u, seq_id ? (uint32_t)seq_id[b] : (uint32_t)b,
                                   (uint32_t)stream_id, g_sc);
        if (threadIdx.x == 0) out[b] = tok;
        __syncthreads();
    }
}

// smol_fast_embed: remember depth code `depth_pos` of every sequence; the next depth step embeds it.
__global__ void smol_store_codes_kernel(int32_t* frame_tokens, const int32_t* codes, int batch, int n_rows, int row) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) frame_tokens[(size_t)b * n_rows + row] = codes[b];
}

cudaError_t sample_launch(const float* logits, int n, int batch, const SmolSampling& s, int stream_id,
                          const int32_t* seq_id, const int32_t* step, int32_t* out, cudaStream_t stream) {
    const size_t smem = (size_t)n * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(smol_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return e;
    smol_sample_kernel<<<batch < 1024 ? batch : 1024, kThreads, smem, stream>>>(logits, n, batch, s, stream_id, seq_id, step, out);
    return cudaGetLastError();
}

cudaError_t store_codes_launch(int32_t* frame_tokens, const int32_t* codes, int batch, int n_rows, int row,
                               cudaStream_t stream) {
    smol_store_codes_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(frame_tokens, codes, batch, n_rows, row);
    return cudaGetLastError();
}

// ---- host-side launch helpers (called from capi.cu) ----------------------------------------------
size_t decode_smem_bytes(const DevModel& M) {
    int kmax = M.dim;
    if (M.inter > kmax) kmax = M.inter;
    if (M.fdim > kmax) kmax = M.fdim;
    if (M.finter > kmax) kmax = M.finter;
    size_t a = (size_t)kBatchTile * kmax * sizeof(float);
    size_t b = (size_t)kWarps * kMaxGroup * kPartialStride * sizeof(float);
    int nl = M.vocab > M.codebook_size ? M.vocab : M.codebook_size;
    size_t cbytes = (size_t)nl * sizeof(float);
    size_t m = a > b ? a : b;
    return m > cbytes ? m : cbytes;
}

// The attribute is per function, not per model: it only ever grows (several models may coexist).
static size_t g_smem_configured = 0;
cudaError_t decode_configure(size_t smem) {
    if (smem <= g_smem_configured) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(smol_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) g_smem_configured = smem;
    return e;
}

cudaError_t decode_max_ctas(size_t smem, int* per_sm) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, smol_decode_kernel, kThreads, smem);
}

cudaError_t decode_launch(const DevModel& M, const CallArgs& A, int n_ctas, size_t smem, cudaStream_t stream) {
    if (A.cooperative) {
        void* args[2] = {(void*)&M, (void*)&A};
        return cudaLaunchCooperativeKernel((const void*)smol_decode_kernel, dim3(n_ctas), dim3(kThreads), args, smem, stream);
    }
    smol_decode_kernel<<<n_ctas, kThreads, smem, stream>>>(M, A);
    return cudaGetLastError();
}

}  // namespace smol
