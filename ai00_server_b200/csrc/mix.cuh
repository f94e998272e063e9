// Residual update + LayerNorm + token shift + static lerps, one CTA per token.
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
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int LN_THREADS = 256;
constexpr int LN_MAXV = 8;      // float4 per thread: C <= 8192
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

// loads row `t` of the updated residual into v[], returns nothing; nv = float4 count for this thread
__device__ __forceinline__ void ln_load_row(const LnMixParams& p, int t, float4 (&v)[LN_MAXV]) {
    const int C = p.C;
#pragma unroll
    for (int j = 0; j < LN_MAXV; ++j) {
        const int c = 4 * (threadIdx.x + LN_THREADS * j);
        if (c < C) {
            float4 a = ld4(p.x_in + (size_t)t * C + c);
            if (p.n_parts > 0) {
                float4 s = ld4(p.parts[0] + (size_t)t * C + c);
                for (int q = 1; q < p.n_parts; ++q) {      // fixed rank order: deterministic, identical on all ranks
                    const float4 b = ld4(p.parts[q] + (size_t)t * C + c);
                    s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
                }
                if (p.n_gate > 0) {
                    const int gb = c / p.gate_cl;
                    const float4 gt = ld4(p.gates[gb] + (size_t)t * p.gate_cl + (c - gb * p.gate_cl));
                    s.x *= gt.x; s.y *= gt.y; s.z *= gt.z; s.w *= gt.w;
                }
                a.x += s.x; a.y += s.y; a.z += s.z; a.w += s.w;
            }
            v[j] = a;
        } else {
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// in-place LayerNorm of v[] (two-pass, f32)
__device__ __forceinline__ void ln_normalize(int C, const float* w, const float* b, float4 (&v)[LN_MAXV], float* red) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXV; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const float mean = block_sum(s, red) / (float)C;
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXV; ++j) {
        const int c = 4 * (threadIdx.x + LN_THREADS * j);
        if (c < C) {
            const float dx = v[j].x - mean, dy = v[j].y - mean, dz = v[j].z - mean, dw = v[j].w - mean;
            s2 += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
    }
    const float var = block_sum(s2, red) / (float)C;
    const float rstd = 1.0f / sqrtf(var + LN_EPS);
#pragma unroll
    for (int j = 0; j < LN_MAXV; ++j) {
        const int c = 4 * (threadIdx.x + LN_THREADS * j);
        if (c < C) {
            const float4 ww = ld4(w + c), bb = ld4(b + c);
            v[j].x = (v[j].x - mean) * rstd * ww.x + bb.x;
            v[j].y = (v[j].y - mean) * rstd * ww.y + bb.y;
            v[j].z = (v[j].z - mean) * rstd * ww.z + bb.z;
            v[j].w = (v[j].w - mean) * rstd * ww.w + bb.w;
        }
    }
}

__global__ void __launch_bounds__(LN_THREADS) ln_mix_kernel(const __grid_constant__ LnMixParams p) {
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x;
    if (t >= p.meta.T()) return;
    const int C = p.C;
    const int slot = p.meta.tok_slot()[t];
    const int prev_t = p.meta.tok_prev()[t];

    float4 v[LN_MAXV], pv[LN_MAXV];
    ln_load_row(p, t, v);
    if (p.x_out != p.x_in || p.n_parts > 0) {
#pragma unroll
        for (int j = 0; j < LN_MAXV; ++j) {
            const int c = 4 * (threadIdx.x + LN_THREADS * j);
            if (c < C) *reinterpret_cast<float4*>(p.x_out + (size_t)t * C + c) = v[j];
        }
    }
    ln_normalize(C, p.ln_w, p.ln_b, v, red);

    if (prev_t >= 0) {
        ln_load_row(p, prev_t, pv);
        ln_normalize(C, p.ln_w, p.ln_b, pv, red);
    } else {
#pragma unroll
        for (int j = 0; j < LN_MAXV; ++j) {
            const int c = 4 * (threadIdx.x + LN_THREADS * j);
            pv[j] = (c < C) ? ld4(p.shift_state + (size_t)slot * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }

    const bool last = p.meta.tok_last()[t] != 0;
#pragma unroll
    for (int j = 0; j < LN_MAXV; ++j) {
        const int c = 4 * (threadIdx.x + LN_THREADS * j);
        if (c >= C) continue;
        const float4 xx = v[j];
        float4 sx;
        sx.x = pv[j].x - xx.x; sx.y = pv[j].y - xx.y; sx.z = pv[j].z - xx.z; sx.w = pv[j].w - xx.w;
        *reinterpret_cast<float4*>(p.xx_out + (size_t)t * C + c) = xx;
        if (p.sx_out) *reinterpret_cast<float4*>(p.sx_out + (size_t)t * C + c) = sx;
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

__global__ void __launch_bounds__(LN_THREADS) embed_ln0_kernel(const __grid_constant__ EmbedParams p) {
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x;
    if (t >= p.meta.T()) return;
    const int C = p.C;
    int tok = p.meta.tok()[t];
    tok = min(max(tok, 0), p.V - 1);
    float4 v[LN_MAXV];
#pragma unroll
    for (int j = 0; j < LN_MAXV; ++j) {
        const int c = 4 * (threadIdx.x + LN_THREADS * j);
        if (c < C) {
            const uint2 raw = *reinterpret_cast<const uint2*>(p.emb + (size_t)tok * C + c);
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
            const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
            v[j] = make_float4(lo.x, lo.y, hi.x, hi.y);
        } else {
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    ln_normalize(C, p.ln_w, p.ln_b, v, red);
#pragma unroll
    for (int j = 0; j < LN_MAXV; ++j) {
        const int c = 4 * (threadIdx.x + LN_THREADS * j);
        if (c < C) *reinterpret_cast<float4*>(p.x_out + (size_t)t * C + c) = v[j];
    }
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

__global__ void __launch_bounds__(LN_THREADS) ln_out_kernel(const __grid_constant__ LnOutParams p) {
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    const int t = blockIdx.x;
    if (t >= p.meta.T()) return;
    const int C = p.C;
    const int slot = p.meta.tok_slot()[t];
    const bool last = p.meta.tok_last()[t] != 0;
    const int row = p.meta.tok_outrow()[t];
    if (last && p.commit_dst) {
#pragma unroll
        for (int j = 0; j < LN_MAXV; ++j) {
            const int c = 4 * (threadIdx.x + LN_THREADS * j);
            if (c < C)
                *reinterpret_cast<float4*>(p.commit_dst + (size_t)slot * C + c) = ld4(p.commit_src + (size_t)t * C + c);
        }
    }
    if (row < 0 && !p.hidden_out) return;
    LnMixParams q;     // reuse the row loader
    q.x_in = p.x_in; q.C = C; q.n_parts = p.n_parts; q.n_gate = p.n_gate; q.gate_cl = p.gate_cl;
#pragma unroll
    for (int i = 0; i < 8; ++i) { q.parts[i] = p.parts[i]; q.gates[i] = p.gates[i]; }
    float4 v[LN_MAXV];
    ln_load_row(q, t, v);
    if (p.hidden_out) {
#pragma unroll
        for (int j = 0; j < LN_MAXV; ++j) {
            const int c = 4 * (threadIdx.x + LN_THREADS * j);
            if (c < C) *reinterpret_cast<float4*>(p.hidden_out + (size_t)t * C + c) = v[j];
        }
    }
    if (row < 0) return;
    ln_normalize(C, p.ln_w, p.ln_b, v, red);
#pragma unroll
    for (int j = 0; j < LN_MAXV; ++j) {
        const int c = 4 * (threadIdx.x + LN_THREADS * j);
        if (c < C) {
            uint2 o;
            o.x = pack_h2(v[j].x, v[j].y);
            o.y = pack_h2(v[j].z, v[j].w);
            *reinterpret_cast<uint2*>(p.head_in + a16_index(row, c, p.kq_tile)) = o;
        }
    }
}

}  // namespace b200
