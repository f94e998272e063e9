// WKV recurrences: v5 / v6 multi-head state update and v7 delta rule, fused with the per-head
// GroupNorm (eps 64e-5), the v7 bonus term and the output gate.
//
// Replaces web-rwkv's `time_mix_v5` / `time_mix_v6` / `time_mix_v7` + `group_norm` WGSL
// dispatches under `Runtime::infer` (reference run.rs:1143; SURVEY.md §2.2 K6/K6'/K7, math in
// App. A / App. B).
//
// One 256-thread group per (head, slot).  The 64x64 f32 head state (16 KB) lives in HBM as
// M[value][key] for every version (v6's S[key][value] is stored transposed; the API layout is
// restored by the state import/export kernels), so that
//   * each thread owns a 4(value) x 4(key) patch: 4 coalesced 16-byte accesses, rows moved as full
//     256-byte runs;
//   * every reduction of the recurrence runs over the KEY index = across the 16 lanes of a
//     half-warp -> pure shuffles, no shared-memory round trip:
//       v5/v6: out[v] = sum_k r[k] * (u[k] k[k] v[v] + M[v][k]);  M[v][k] = k[k] v[v] + w[k] M[v][k]
//       v7:    sa[v]  = sum_k M[v][k] * (-kk[k]);
//              M[v][k] = M[v][k] w[k] + sa[v] (kk[k] a[k]) + v[v] k[k];   out[v] = sum_k M[v][k] r[k]
//   * state is read once and written once per step, the recurrence loops over the slot's tokens
//     with the state in registers (prefill chunks).
// Output: f16( GroupNorm(out) [+ bonus] * gate ) written straight into the A16 operand of the
// output projection.
//
// Two callers share `wkv_slot`: the stand-alone kernel below (one CTA per (head, slot), state
// through coalesced vector loads) and the persistent whole-step kernel (mega.cuh), where the
// state tiles arrive through the bulk-TMA stage ring and the v6 decay LoRA is evaluated in place.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int WKV_THREADS = 256;
constexpr int WKV_N = 64;            // head size (all supported RWKV v5/v6/v7 models)
constexpr float GN_EPS = 64e-5f;

struct WkvParams {
    int version;            // 5, 6, 7
    int ld;                 // row stride of r/k/v/g/w/a/nu (floats) = local channels
    MetaView meta;
    float* state;           // [S][H][64][64] this layer, M[value][key]
    int H;                  // local heads
    const float* r;
    const float* k;
    const float* v;
    const float* g;         // v5/v6: silu(gate proj); v7: gate LoRA output
    const float* w;         // [T, ld] decay in (0,1) (v6/v7); null for v5
    const float* w_static;  // [ld] v5 decay exp(-exp(time_decay))
    const float* u;         // [ld] v5/v6 time_first
    const float* lnx_w;
    const float* lnx_b;
    // v7
    const float* a;         // [T, ld] in-context learning rate
    const float* nu;        // [T, ld] value-residual gate (layers > 0)
    float* v_first;         // [T, ld] layer 0 writes, later layers read
    int layer0;
    const float* k_k;
    const float* k_a;
    const float* r_k;
    __half* out;            // A16 [T, ld]
    int kq_tile;
    // v6, whole-step kernel only: decay LoRA stage 2 evaluated inside the WKV phase
    const __half* wd2t;     // [H][Dd][64] f16: time_decay_w2 rows of each head, k-major
    const float* decay_bias;    // [ld] time_decay
    const __half* d1;       // A16 [T, Dd]: tanh(time_decay_w1 @ xw)
    int d1_kq;
    int Dd;
    unsigned long long* trace;  // profiling aid (null in production)
};

struct WkvShared {
    float r[WKV_N], k[WKV_N], v[WKV_N], w[WKV_N], b[WKV_N], o[WKV_N];
    float red[4];
};

// Runs the recurrence for `nt` tokens starting at token index t0 on head h with the state patch
// m[4] (rows 4*ig+e, cols 4*j4..) in registers.  `w_local`: optional shared-memory decay rows
// [token][64] (whole-step kernel, v6) indexed from local token `lt0`.
// `pre`: optional shared-memory copy of the head's per-token vectors, [array][token][64] with arrays
// r, k, v, g (, w, a, nu for v7) and `pre_stride` floats between arrays (whole-step kernel: gathered
// once per WKV unit so the per-token loop never waits on L2).
template <int VER, bool MEGA>
__device__ __forceinline__ void wkv_slot(const WkvParams& p, const int h, const int t0, const int nt, float4 (&m)[4],
                                         WkvShared& sm, const float* w_local, const int lt0, const float* pre = nullptr,
                                         const int pre_stride = 0) {
    const int tid = threadIdx.x;
    const int ig = tid >> 4;           // value rows 4*ig .. 4*ig+3
    const int j4 = tid & 15;           // key cols  4*j4 .. 4*j4+3
    const int ch = h * WKV_N;          // channel base of this head
    float u4[4] = {0.f, 0.f, 0.f, 0.f};
    if (VER != 7) {
#pragma unroll
        for (int f = 0; f < 4; ++f) u4[f] = p.u[ch + j4 * 4 + f];
    }
    for (int tt = 0; tt < nt; ++tt) {
        const int t = t0 + tt;
        const size_t row = (size_t)t * p.ld + ch;
        // ---- per-token head vectors -> shared ----
        if (tid < WKV_N) {
            const int c = tid;
            const int pi = (lt0 + tt) * WKV_N + c;
            float r = pre ? pre[pi] : p.r[row + c];
            float k = pre ? pre[pre_stride + pi] : p.k[row + c];
            float v = pre ? pre[2 * pre_stride + pi] : p.v[row + c];
            float w;
            if (VER == 5) w = p.w_static[ch + c];
            else if (w_local) w = w_local[pi];
            else w = pre ? pre[4 * pre_stride + pi] : p.w[row + c];
            if (VER == 7) {
                const float a = pre ? pre[5 * pre_stride + pi] : p.a[row + c];
                float kk = k * p.k_k[ch + c];
                // l2 norm over the head: two warps
                float ss = warp_sum(kk * kk);
                if ((tid & 31) == 0) sm.red[tid >> 5] = ss;
                asm volatile("bar.sync 2, 64;" ::: "memory");
                ss = sm.red[0] + sm.red[1];
                kk = kk / fmaxf(sqrtf(ss), 1e-12f);
                k = k * (1.f + (a - 1.f) * p.k_a[ch + c]);
                if (p.layer0) p.v_first[row + c] = v;
                else v = v + (p.v_first[row + c] - v) * (pre ? pre[6 * pre_stride + pi] : p.nu[row + c]);
                float bonus = warp_sum(r * k * p.r_k[ch + c]);
                if ((tid & 31) == 0) sm.red[2 + (tid >> 5)] = bonus;
                sm.b[c] = kk * a;       // kk (.) a
                sm.o[c] = -kk;          // reuse o as -kk until the output phase
            }
            sm.r[c] = r; sm.k[c] = k; sm.v[c] = v; sm.w[c] = w;
        }
        cta_sync<MEGA>();

        float rr[4], kk_[4], ww[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) { rr[f] = sm.r[j4 * 4 + f]; kk_[f] = sm.k[j4 * 4 + f]; ww[f] = sm.w[j4 * 4 + f]; }
        float o[4];
        if (VER != 7) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float vv = sm.v[ig * 4 + e];
                float* me = reinterpret_cast<float*>(&m[e]);
                float acc = 0.f;
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    const float kv = kk_[f] * vv;
                    acc += rr[f] * (u4[f] * kv + me[f]);
                    me[f] = kv + ww[f] * me[f];
                }
                o[e] = acc;
            }
        } else {
            float nk[4], ka[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) { nk[f] = sm.o[j4 * 4 + f]; ka[f] = sm.b[j4 * 4 + f]; }
            float sa[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float* me = reinterpret_cast<const float*>(&m[e]);
                sa[e] = (me[0] * nk[0] + me[1] * nk[1]) + (me[2] * nk[2] + me[3] * nk[3]);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1)
#pragma unroll
                for (int e = 0; e < 4; ++e) sa[e] += __shfl_xor_sync(0xffffffffu, sa[e], off);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float vv = sm.v[ig * 4 + e];
                float* me = reinterpret_cast<float*>(&m[e]);
                float acc = 0.f;
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    me[f] = me[f] * ww[f] + sa[e] * ka[f] + vv * kk_[f];
                    acc += me[f] * rr[f];
                }
                o[e] = acc;
            }
        }
#pragma unroll
        for (int off = 8; off > 0; off >>= 1)
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] += __shfl_xor_sync(0xffffffffu, o[e], off);
        cta_sync<MEGA>();              // all reads of sm.o (-kk) done before it is overwritten
        if (j4 == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) sm.o[ig * 4 + e] = o[e];
        }
        cta_sync<MEGA>();

        // ---- GroupNorm over the head + gate, warp 0: 2 channels per lane ----
        if (tid < 32) {
            const float x0 = sm.o[tid], x1 = sm.o[tid + 32];
            const float mean = warp_sum(x0 + x1) * (1.f / WKV_N);
            const float d0 = x0 - mean, d1 = x1 - mean;
            const float var = warp_sum(d0 * d0 + d1 * d1) * (1.f / WKV_N);
            const float rstd = 1.0f / sqrtf(var + GN_EPS);
            float y0 = d0 * rstd * p.lnx_w[ch + tid] + p.lnx_b[ch + tid];
            float y1 = d1 * rstd * p.lnx_w[ch + tid + 32] + p.lnx_b[ch + tid + 32];
            if (VER == 7) {
                const float bonus = sm.red[2] + sm.red[3];
                y0 += bonus * sm.v[tid];
                y1 += bonus * sm.v[tid + 32];
            }
            const int gi = (lt0 + tt) * WKV_N + tid;
            y0 *= pre ? pre[3 * pre_stride + gi] : p.g[row + tid];
            y1 *= pre ? pre[3 * pre_stride + gi + 32] : p.g[row + tid + 32];
            p.out[a16_index(t, ch + tid, p.kq_tile)] = f2h_sat(y0);
            p.out[a16_index(t, ch + tid + 32, p.kq_tile)] = f2h_sat(y1);
        }
        cta_sync<MEGA>();              // shared vectors are rewritten by the next token
    }
}

// dynamic shared memory of the stand-alone kernel when the v6 decay LoRA stage 2 is folded in:
// [Dd][64] halves (k-major slice of time_decay_w2) | [max tokens][64] floats (decays) | [Dd] halves | [4][64] floats
__host__ __device__ inline size_t wkv_fold_smem_bytes(int Dd, int max_tokens) {
    return (size_t)WKV_N * Dd * 2 + (size_t)max_tokens * WKV_N * 4 + (size_t)Dd * 2 + 4 * WKV_N * 4 + 64;
}

template <int VER>
__global__ void __launch_bounds__(WKV_THREADS) wkv_kernel(const __grid_constant__ WkvParams p, const int max_tokens) {
    __shared__ WkvShared sm;
    extern __shared__ __align__(16) uint8_t wkv_dyn[];
    trace_stamp(p.trace, 0);
    pdl_launch_dependents();
    const int si = blockIdx.y;
    const int h = blockIdx.x;
    const int tid = threadIdx.x;
    const int ig = tid >> 4, j4 = tid & 15;
    const bool fold = (VER == 6) && p.wd2t != nullptr;
    // the decay-LoRA slice is a weight: fetch it before waiting on the producer kernel
    __half* wt = reinterpret_cast<__half*>(wkv_dyn);
    if (fold) {
        const uint4* src = reinterpret_cast<const uint4*>(p.wd2t + (size_t)h * WKV_N * p.Dd);
        uint4* dst = reinterpret_cast<uint4*>(wt);
        for (int i = tid; i < WKV_N * p.Dd / 8; i += WKV_THREADS) dst[i] = src[i];
    }
    pdl_wait();
    trace_stamp(p.trace, 1);
    if (si >= p.meta.nslots()) return;
    const int slot = p.meta.slot_id()[si];
    const int t0 = p.meta.slot_start()[si];
    const int nt = p.meta.slot_count()[si];

    float* M = p.state + ((size_t)slot * p.H + h) * (WKV_N * WKV_N);
    float4 m[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) m[e] = __ldcs(reinterpret_cast<const float4*>(M + (ig * 4 + e) * WKV_N + j4 * 4));

    const float* w_local = nullptr;
    if (fold) {
        // w[t][c] = exp(-exp(time_decay[c] + sum_k Wd2[c][k] * tanh(Wd1 xw)[t][k]))   
        const int Dd = p.Dd;
        float* wl = reinterpret_cast<float*>(wkv_dyn + (size_t)WKV_N * Dd * 2);
        __half* ds = reinterpret_cast<__half*>(wl + (size_t)max_tokens * WKV_N);
        float* part = reinterpret_cast<float*>(wkv_dyn + (size_t)WKV_N * Dd * 2 + (size_t)max_tokens * WKV_N * 4 + (((size_t)Dd * 2 + 15) & ~(size_t)15));
        const int c = tid & (WKV_N - 1), qk = tid >> 6;
        const int kq0 = qk * (Dd >> 2), kq1 = kq0 + (Dd >> 2);
        const float bias = p.decay_bias[h * WKV_N + c];
        for (int tt = 0; tt < nt; ++tt) {
            __syncthreads();
            for (int k = tid; k < Dd; k += WKV_THREADS) ds[k] = p.d1[a16_index(t0 + tt, k, p.d1_kq)];
            __syncthreads();
            float acc = 0.f;
            for (int k = kq0; k < kq1; ++k) acc = fmaf(__half2float(wt[k * WKV_N + c]), __half2float(ds[k]), acc);
            part[qk * WKV_N + c] = acc;
            __syncthreads();
            if (tid < WKV_N) {
                const float s = (part[c] + part[WKV_N + c]) + (part[2 * WKV_N + c] + part[3 * WKV_N + c]);
                wl[tt * WKV_N + c] = expf(-expf(bias + s));
            }
        }
        __syncthreads();
        w_local = wl;
    }
    wkv_slot<VER, false>(p, h, t0, nt, m, sm, w_local, 0);
#pragma unroll
    for (int e = 0; e < 4; ++e) __stcs(reinterpret_cast<float4*>(M + (ig * 4 + e) * WKV_N + j4 * 4), m[e]);
    trace_stamp(p.trace, 7);
}

}  // namespace b200
