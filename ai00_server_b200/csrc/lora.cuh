// CUDA-core phases of the whole-step kernel for the tiny LoRA projections of RWKV-6's data-dependent
// token shift (reference path: web-rwkv's `token_shift` + small matmul dispatches under
// `Runtime::infer`, run.rs:1143; math SURVEY.md App. A):
//   small-N:  m = tanh(W1 @ xxx)              W1 [5*Dm, C]   -> 320 outputs that each need a full-C dot
//   small-K:  x_j = xx + sx * (mu_j + W2_j @ m_j)   W2 [5, C, Dm]  -> 20480 outputs with 64-long dots
// Measured on B200, routing these 2.6 MB matrices through the stream-K tensor-core GEMM costs
// ~15-19 us per phase (cross-CTA fix-up of K-split tiles, latency chains); here no output is
// shared between CTAs, so there is no reduction across CTAs and each phase is a couple of L2
// round trips.  Weights are read in their original row-major [out, in] layout.
#pragma once
#include "gemm.cuh"

namespace b200 {

struct SmallNParams {        // out[tok][n] = act( sum_k W[n][k] * A[tok][k] ),  N small, K = C
    const __half* W;         // [N][K] row-major
    int N, K;
    const __half* A;         // A16 [T][K]
    int a_kq;
    __half* out;             // A16, column groups of `grp`
    int grp, grp_stride, out_kq;
    int act;
    const int* nrows;
};

struct SmallKSeg {           // out[tok][n] = epilogue( sum_k W[n][k] * A[tok][k] ),  K small
    const __half* W;         // [N][K] row-major
    int N, K;
    const __half* A;         // A16 [T][Kpad]
    int a_kq;
    int out_mode, act;       // OutMode / Act of gemm.cuh
    const float* bias;
    void* out;
    int ldo;
    const float* aux0;       // OUT_LERP_A16: xx
    const float* aux1;       //               sx
    const float* aux2;       //               mu
    int ld_aux;
};
struct SmallKParams {
    int nseg;
    SmallKSeg seg[5];
    const int* nrows;
};

constexpr int LORA_MAX_ROWS_PER_CTA = 4;

// 256 consumer threads; `scratch`: >= 8*16*LORA_MAX_ROWS_PER_CTA floats of shared memory
template <bool MEGA>
__device__ __forceinline__ void smalln_phase(const SmallNParams& p, const int cta, const int G, float* scratch) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N, K8 = p.K >> 3, a_k8 = p.a_kq * 4;
    const int nrows = min(*p.nrows, 16);
    if (cta >= N) return;
    // this thread's two 16-byte k chunks of the activations (all 16 tokens), kept across the CTA's rows
    const int per_thread = (K8 + CONSUMER_THREADS - 1) / CONSUMER_THREADS;     // 2 at C = 4096
    for (int ch0 = 0; ch0 < per_thread; ch0 += 2) {
        uint4 xa[2][16];
        int chunk[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            chunk[u] = (ch0 + u) * CONSUMER_THREADS + tid;
            if (ch0 + u < per_thread && chunk[u] < K8) {
                const uint4* src = reinterpret_cast<const uint4*>(p.A + (size_t)chunk[u] * 128);
#pragma unroll
                for (int t = 0; t < 16; ++t) xa[u][t] = src[t];
            } else {
                chunk[u] = -1;
#pragma unroll
                for (int t = 0; t < 16; ++t) xa[u][t] = make_uint4(0, 0, 0, 0);
            }
        }
        (void)a_k8;
        int ri = 0;
#pragma unroll 1
        for (int n = cta; n < N; n += G, ++ri) {
            float acc[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) acc[t] = 0.f;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (chunk[u] < 0) continue;
                const uint4 wv = *reinterpret_cast<const uint4*>(p.W + (size_t)n * p.K + (size_t)chunk[u] * 8);
                const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const __half2* xh = reinterpret_cast<const __half2*>(&xa[u][t]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 wf = __half22float2(wh[e]), xf = __half22float2(xh[e]);
                        acc[t] = fmaf(wf.x, xf.x, acc[t]);
                        acc[t] = fmaf(wf.y, xf.y, acc[t]);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < 16; ++t) acc[t] = warp_sum(acc[t]);
            if (lane == 0) {
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    float* dst = scratch + (ri * 8 + warp) * 16 + t;
                    if (ch0 == 0) *dst = acc[t];
                    else *dst += acc[t];
                }
            }
        }
    }
    cta_sync<MEGA>();
    // finish: thread (row ri, token t) sums the 8 warps
    const int rows_here = (N - cta + G - 1) / G;
    if (tid < rows_here * 16) {
        const int ri = tid >> 4, t = tid & 15;
        const int n = cta + ri * G;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += scratch[(ri * 8 + w) * 16 + t];
        if (t < nrows) {
            const float y = apply_act(s, p.act);
            const int gi = (p.grp > 0) ? n / p.grp : 0;
            const int nn = n - gi * (p.grp > 0 ? p.grp : 0);
            p.out[(size_t)gi * p.grp_stride + a16_index(t, nn, p.out_kq)] = f2h_sat(y);
        }
    }
}

// `in_s`: >= 16 * Kmax floats of shared memory ([k][16 tokens])
template <bool MEGA>
__device__ __forceinline__ void smallk_phase(const SmallKParams& p, const int cta, const int G, float* in_s) {
    const int tid = threadIdx.x;
    const int nrows = min(*p.nrows, 16);
    int total = 0;
    for (int s = 0; s < p.nseg; ++s) total += p.seg[s].N;
    const int lo = (int)((long long)cta * total / G), hi = (int)((long long)(cta + 1) * total / G);
    int seg_base = 0;
#pragma unroll 1
    for (int s = 0; s < p.nseg; ++s) {
        const SmallKSeg sg = p.seg[s];             // registers
        const int r0 = max(lo, seg_base), r1 = min(hi, seg_base + sg.N);
        seg_base += sg.N;
        if (r0 >= r1) continue;
        const int K = sg.K;
        cta_sync<MEGA>();                          // in_s reuse across segments
        // activations of this segment -> shared, f32 [k][16]
        for (int i = tid; i < (K >> 3) * 16; i += CONSUMER_THREADS) {
            const int k8 = i >> 4, t = i & 15;
            const uint4 raw = *reinterpret_cast<const uint4*>(sg.A + ((size_t)k8 * 16 + t) * 8);
            const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h[e]);
                in_s[(k8 * 8 + 2 * e) * 16 + t] = f.x;
                in_s[(k8 * 8 + 2 * e + 1) * 16 + t] = f.y;
            }
        }
        cta_sync<MEGA>();
#pragma unroll 1
        for (int row = r0 + tid; row < r1; row += CONSUMER_THREADS) {
            const int n = row - (seg_base - sg.N);
            float acc[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) acc[t] = 0.f;
            // epilogue operands requested up front: their latency hides behind the dot products
            float xx[16], sx[16];
            float mu = 0.f, bias = 0.f;
            if (sg.out_mode == OUT_LERP_A16) {
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    xx[t] = (t < nrows) ? sg.aux0[(size_t)t * sg.ld_aux + n] : 0.f;
                    sx[t] = (t < nrows) ? sg.aux1[(size_t)t * sg.ld_aux + n] : 0.f;
                }
                mu = sg.aux2[n];
            }
            if (sg.bias) bias = sg.bias[n];
            const uint4* wrow = reinterpret_cast<const uint4*>(sg.W + (size_t)n * K);
#pragma unroll 1
            for (int k8 = 0; k8 < (K >> 3); k8 += 8) {
                uint4 wv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) wv[u] = (k8 + u < (K >> 3)) ? wrow[k8 + u] : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (k8 + u >= (K >> 3)) break;
                    const __half* wh = reinterpret_cast<const __half*>(&wv[u]);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float wf = __half2float(wh[e]);
                        const float4* xr = reinterpret_cast<const float4*>(in_s + ((k8 + u) * 8 + e) * 16);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 x4 = xr[q];
                            acc[4 * q] = fmaf(wf, x4.x, acc[4 * q]);
                            acc[4 * q + 1] = fmaf(wf, x4.y, acc[4 * q + 1]);
                            acc[4 * q + 2] = fmaf(wf, x4.z, acc[4 * q + 2]);
                            acc[4 * q + 3] = fmaf(wf, x4.w, acc[4 * q + 3]);
                        }
                    }
                }
            }
            if (sg.out_mode == OUT_F32) {
                float* o = reinterpret_cast<float*>(sg.out) + n;
#pragma unroll
                for (int t = 0; t < 16; ++t)
                    if (t < nrows) o[(size_t)t * sg.ldo] = apply_act(acc[t] + bias, sg.act);
            } else {
                __half* base = reinterpret_cast<__half*>(sg.out);
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    if (t >= nrows) break;
                    float y = apply_act(acc[t] + bias, sg.act);
                    if (sg.out_mode == OUT_LERP_A16) y = xx[t] + sx[t] * (mu + y);
                    base[a16_index(t, n, sg.ldo)] = f2h_sat(y);
                }
            }
        }
    }
}

// stand-alone launches of the same phases (per-op kernel chain)
__global__ void __launch_bounds__(CONSUMER_THREADS) smalln_kernel(const __grid_constant__ SmallNParams p) {
    __shared__ float scratch[LORA_MAX_ROWS_PER_CTA * 8 * 16];
    pdl_launch_dependents();
    pdl_wait();
    smalln_phase<false>(p, blockIdx.x, gridDim.x, scratch);
}
__global__ void __launch_bounds__(CONSUMER_THREADS) smallk_kernel(const __grid_constant__ SmallKParams p) {
    extern __shared__ __align__(16) float lora_in_s[];
    pdl_launch_dependents();
    pdl_wait();
    smallk_phase<false>(p, blockIdx.x, gridDim.x, lora_in_s);
}

}  // namespace b200
