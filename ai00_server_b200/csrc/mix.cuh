// Residual update + LayerNorm + token shift + static lerps, one 256-thread group per token.
//
// Replaces web-rwkv's `layer_norm`, `token_shift` and `add` WGSL dispatches on the path
// `Runtime::infer` (reference run.rs:1143).  Fused so the
// residual stream makes one HBM/L2 round trip per half-layer:
//     x_out = x_in + gate (.) sum_p part[p]            (projection results of the previous
//                                                       half-layer, one partial per TP rank)
//     xx    = LN(x_out)                                 eps 1e-5
//     prev  = first token of its slot ? shift_state[slot] : LN(x_out[t-1])   (recomputed)
//     sx    = prev - xx
//     mix_j = f16(xx + sx * mu_j)  -> A16 operand buffers of the following projections
// The shift state of the PREVIOUS LN stage is committed here (dst <- last token's xx), which
// keeps every state write strictly after all reads of the old value with no extra launch.
//
// Code shape: these rows run once per launch / per phase of the whole-step kernel, i.e. with a
// cold instruction cache (measured: straight-line unrolled code costs microseconds of
// instruction fetch), so every pass is a small rolled loop over a shared-memory copy of the row.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int LN_THREADS = 256;
constexpr int LN_MAXC = 8192;       // largest row the stand-alone kernels buffer
constexpr float LN_EPS = 1e-5f;

struct LnMixParams {
    const float* x_in;      // [T, C]
    float* x_out;           // [T, C] (may alias x_in only when n_parts == 0)
    int C;
    MetaView meta;
    int n_parts;
    const float* parts[8];  // [T, C] each
    int n_gate;             // 0 = no gate; else column-blocked gate, one block per TP rank
    int gate_cl;            // columns per gate block (C / n_gate)
    const float* gates[8];  // [T, gate_cl] each
    const float* ln_w;
    const float* ln_b;
    const float* shift_state;   // [S, C] rows of this layer / kind
    int n_mix;
    const float* mu[6];
    __half* mix_out[6];
    int kq_tile;            // C / 32
    float* xx_out;          // [T, C]
    float* sx_out;          // [T, C] or null
    float* commit_dst;      // [S, C] or null
    const float* commit_src;    // [T, C]
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// residual source shared by the LN stages: x_in + gate (.) sum_p parts[p]
struct ResidualSrc {     // held by value: pointers into kernel-parameter space would turn into slow generic loads
    const float* x_in;
    int C, n_parts, n_gate, gate_cl;
    const float* parts[8];
    const float* gates[8];
};
template <typename P>
__device__ __forceinline__ ResidualSrc make_residual_src(const P& p) {
    ResidualSrc r;
    r.x_in = p.x_in; r.C = p.C; r.n_parts = p.n_parts; r.n_gate = p.n_gate; r.gate_cl = p.gate_cl;
#pragma unroll
    for (int i = 0; i < 8; ++i) { r.parts[i] = p.parts[i]; r.gates[i] = p.gates[i]; }
    return r;
}

// updated residual at (t, c..c+3)
__device__ __noinline__ float4 residual4(const ResidualSrc& r, const int t, const int c) {
    float4 a = ld4(r.x_in + (size_t)t * r.C + c);
    if (r.n_parts > 0) {
        float4 s = ld4(r.parts[0] + (size_t)t * r.C + c);
#pragma unroll 1
        for (int q = 1; q < r.n_parts; ++q) {      // fixed rank order: deterministic, identical on all ranks
            const float4 b = ld4(r.parts[q] + (size_t)t * r.C + c);
            s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
        }
        if (r.n_gate > 0) {
            const int gb = c / r.gate_cl;
            const float4 gt = ld4(r.gates[gb] + (size_t)t * r.gate_cl + (c - gb * r.gate_cl));
            s.x *= gt.x; s.y *= gt.y; s.z *= gt.z; s.w *= gt.w;
        }
        a.x += s.x; a.y += s.y; a.z += s.z; a.w += s.w;
    }
    return a;
}

// mean / rstd of row t of the updated residual; optionally keeps the row in `row` and writes x_out
template <bool MEGA>
__device__ __forceinline__ void ln_stats(const ResidualSrc& r, const int t, float* row, float* x_out, float* red, float& mean,
                                         float& rstd) {
    const int C = r.C;
    float s = 0.f;
#pragma unroll 1
    for (int c = 4 * threadIdx.x; c < C; c += 4 * LN_THREADS) {
        const float4 a = residual4(r, t, c);
        if (row) *reinterpret_cast<float4*>(row + c) = a;
        if (x_out) *reinterpret_cast<float4*>(x_out + (size_t)t * C + c) = a;
        s += (a.x + a.y) + (a.z + a.w);
    }
    mean = block_sum<MEGA>(s, red) / (float)C;
    float s2 = 0.f;
#pragma unroll 1
    for (int c = 4 * threadIdx.x; c < C; c += 4 * LN_THREADS) {
        const float4 a = row ? *reinterpret_cast<const float4*>(row + c) : residual4(r, t, c);
        const float dx = a.x - mean, dy = a.y - mean, dz = a.z - mean, dw = a.w - mean;
        s2 += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float var = block_sum<MEGA>(s2, red) / (float)C;
    rstd = 1.0f / sqrtf(var + LN_EPS);
}

__device__ __forceinline__ float4 ln_apply(const float4 a, const float mean, const float rstd, const float4 w, const float4 b) {
    float4 o;
    o.x = (a.x - mean) * rstd * w.x + b.x;
    o.y = (a.y - mean) * rstd * w.y + b.y;
    o.z = (a.z - mean) * rstd * w.z + b.z;
    o.w = (a.w - mean) * rstd * w.w + b.w;
    return o;
}

// `row`: shared-memory buffer of >= C floats
template <bool MEGA>
__device__ __forceinline__ void ln_mix_row(const LnMixParams& p, const int t, float* row, float* red) {
    const int C = p.C;
    const int slot = p.meta.tok_slot()[t];
    const int prev_t = p.meta.tok_prev()[t];
    const bool last = p.meta.tok_last()[t] != 0;
    const ResidualSrc r = make_residual_src(p);
    float mean, rstd, pmean = 0.f, prstd = 0.f;
    ln_stats<MEGA>(r, t, row, (p.x_out != p.x_in || p.n_parts > 0) ? p.x_out : nullptr, red, mean, rstd);
    if (prev_t >= 0) ln_stats<MEGA>(r, prev_t, nullptr, nullptr, red, pmean, prstd);
#pragma unroll 1
    for (int c = 4 * threadIdx.x; c < C; c += 4 * LN_THREADS) {
        const float4 w = ld4(p.ln_w + c), b = ld4(p.ln_b + c);
        const float4 xx = ln_apply(*reinterpret_cast<const float4*>(row + c), mean, rstd, w, b);
        const float4 pv = (prev_t >= 0) ? ln_apply(residual4(r, prev_t, c), pmean, prstd, w, b)
                                        : ld4(p.shift_state + (size_t)slot * C + c);
        float4 sx;
        sx.x = pv.x - xx.x; sx.y = pv.y - xx.y; sx.z = pv.z - xx.z; sx.w = pv.w - xx.w;
        *reinterpret_cast<float4*>(p.xx_out + (size_t)t * C + c) = xx;
        if (p.sx_out) *reinterpret_cast<float4*>(p.sx_out + (size_t)t * C + c) = sx;
#pragma unroll 1
        for (int m = 0; m < p.n_mix; ++m) {
            const float4 mu = ld4(p.mu[m] + c);
            uint2 o;
            o.x = pack_h2(xx.x + sx.x * mu.x, xx.y + sx.y * mu.y);
            o.y = pack_h2(xx.z + sx.z * mu.z, xx.w + sx.w * mu.w);
            *reinterpret_cast<uint2*>(p.mix_out[m] + a16_index(t, c, p.kq_tile)) = o;
        }
        if (last && p.commit_dst)
            *reinterpret_cast<float4*>(p.commit_dst + (size_t)slot * C + c) = ld4(p.commit_src + (size_t)t * C + c);
    }
}

__global__ void __launch_bounds__(LN_THREADS) ln_mix_kernel(const __grid_constant__ LnMixParams p) {
    extern __shared__ __align__(16) float ln_row[];
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x;
    if (t >= p.meta.T()) return;
    ln_mix_row<false>(p, t, ln_row, red);
}

// ---------------------------------------------------------------------------------------
// Embedding gather + LN0 (reference: web-rwkv embeds on the CPU and runs ln0 on the device,
// SURVEY.md §2.2 K1; here the f16 table lives in HBM and both are one kernel).
// ---------------------------------------------------------------------------------------
struct EmbedParams {
    const __half* emb;     // [V, C]
    int C, V;
    MetaView meta;
    const float* ln_w;
    const float* ln_b;
    float* x_out;          // [T, C]
};

template <bool MEGA>
__device__ __forceinline__ void embed_row(const EmbedParams& p, const int t, float* row, float* red) {
    const int C = p.C;
    int tok = p.meta.tok()[t];
    tok = min(max(tok, 0), p.V - 1);
    float s = 0.f;
#pragma unroll 1
    for (int c = 4 * threadIdx.x; c < C; c += 4 * LN_THREADS) {
        const uint2 raw = *reinterpret_cast<const uint2*>(p.emb + (size_t)tok * C + c);
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
        *reinterpret_cast<float4*>(row + c) = make_float4(lo.x, lo.y, hi.x, hi.y);
        s += (lo.x + lo.y) + (hi.x + hi.y);
    }
    const float mean = block_sum<MEGA>(s, red) / (float)C;
    float s2 = 0.f;
#pragma unroll 1
    for (int c = 4 * threadIdx.x; c < C; c += 4 * LN_THREADS) {
        const float4 a = *reinterpret_cast<const float4*>(row + c);
        const float dx = a.x - mean, dy = a.y - mean, dz = a.z - mean, dw = a.w - mean;
        s2 += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float var = block_sum<MEGA>(s2, red) / (float)C;
    const float rstd = 1.0f / sqrtf(var + LN_EPS);
#pragma unroll 1
    for (int c = 4 * threadIdx.x; c < C; c += 4 * LN_THREADS)
        *reinterpret_cast<float4*>(p.x_out + (size_t)t * C + c) =
            ln_apply(*reinterpret_cast<const float4*>(row + c), mean, rstd, ld4(p.ln_w + c), ld4(p.ln_b + c));
}

__global__ void __launch_bounds__(LN_THREADS) embed_ln0_kernel(const __grid_constant__ EmbedParams p) {
    extern __shared__ __align__(16) float ln_row[];
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x;
    if (t >= p.meta.T()) return;
    embed_row<false>(p, t, ln_row, red);
}

// ---------------------------------------------------------------------------------------
// Final residual update + ln_out for the rows that need logits (RnnOption::Last -> last token
// of the slot, Full -> every token; reference run.rs:812-822, 710-724), gathered into the A16
// operand of the head GEMM.  Also commits the last layer's channel-mix shift state.
// ---------------------------------------------------------------------------------------
struct LnOutParams {
    const float* x_in;      // [T, C]
    int C;
    MetaView meta;
    int n_parts;
    const float* parts[8];
    int n_gate;
    int gate_cl;
    const float* gates[8];
    const float* ln_w;
    const float* ln_b;
    __half* head_in;        // A16 [R rows, C]
    int kq_tile;
    float* commit_dst;
    const float* commit_src;
    float* hidden_out;      // optional [T, C]: updated residual (debug / states endpoint)
};

template <bool MEGA>
__device__ __forceinline__ void ln_out_row(const LnOutParams& p, const int t, float* rowbuf, float* red) {
    const int C = p.C;
    const int slot = p.meta.tok_slot()[t];
    const bool last = p.meta.tok_last()[t] != 0;
    const int row = p.meta.tok_outrow()[t];
    if (last && p.commit_dst) {
#pragma unroll 1
        for (int c = 4 * threadIdx.x; c < C; c += 4 * LN_THREADS)
            *reinterpret_cast<float4*>(p.commit_dst + (size_t)slot * C + c) = ld4(p.commit_src + (size_t)t * C + c);
    }
    if (row < 0 && !p.hidden_out) return;
    const ResidualSrc r = make_residual_src(p);
    float mean, rstd;
    ln_stats<MEGA>(r, t, rowbuf, p.hidden_out, red, mean, rstd);
    if (row < 0) return;
#pragma unroll 1
    for (int c = 4 * threadIdx.x; c < C; c += 4 * LN_THREADS) {
        const float4 y = ln_apply(*reinterpret_cast<const float4*>(rowbuf + c), mean, rstd, ld4(p.ln_w + c), ld4(p.ln_b + c));
        uint2 o;
        o.x = pack_h2(y.x, y.y);
        o.y = pack_h2(y.z, y.w);
        *reinterpret_cast<uint2*>(p.head_in + a16_index(row, c, p.kq_tile)) = o;
    }
}

__global__ void __launch_bounds__(LN_THREADS) ln_out_kernel(const __grid_constant__ LnOutParams p) {
    extern __shared__ __align__(16) float ln_row[];
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x;
    if (t >= p.meta.T()) return;
    ln_out_row<false>(p, t, ln_row, red);
}

}  // namespace b200
