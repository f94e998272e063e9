// Device-side helpers shared by all kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// ---------------------------------------------------------------------------------------
// Per-step metadata, one flat int array in HBM (rewritten by a single H2D copy per step so
// the captured CUDA graph never changes).  Layout: header, then arrays sized by the
// engine's token_chunk (maxT) and max_batch (maxS).
// ---------------------------------------------------------------------------------------
struct MetaView {
    const int* base;
    int maxT, maxS;
    __host__ __device__ int T() const { return base[0]; }        // tokens this step
    __host__ __device__ int nslots() const { return base[1]; }   // active slots this step
    __host__ __device__ int R() const { return base[2]; }        // logits rows this step
    __host__ __device__ const int* tok() const { return base + 8; }                       // [maxT] token ids
    __host__ __device__ const int* tok_slot() const { return base + 8 + maxT; }           // [maxT] state slot of token
    __host__ __device__ const int* tok_prev() const { return base + 8 + 2 * maxT; }       // [maxT] t-1 or -1 (take shift state)
    __host__ __device__ const int* tok_last() const { return base + 8 + 3 * maxT; }       // [maxT] 1 if last token of its slot in this step
    __host__ __device__ const int* out_tok() const { return base + 8 + 4 * maxT; }        // [maxT] logits row r -> token index
    __host__ __device__ const int* tok_outrow() const { return base + 8 + 5 * maxT; }     // [maxT] token -> logits row or -1
    __host__ __device__ const int* slot_id() const { return base + 8 + 6 * maxT; }        // [maxS] active slot -> state slot
    __host__ __device__ const int* slot_start() const { return base + 8 + 6 * maxT + maxS; }
    __host__ __device__ const int* slot_count() const { return base + 8 + 6 * maxT + 2 * maxS; }
    __host__ __device__ static size_t ints(int maxT, int maxS) { return 8 + 6 * (size_t)maxT + 3 * (size_t)maxS; }
};

// ---------------------------------------------------------------------------------------
// "A16" activation layout: the f16 operand of every projection, stored so that the
// k-range a GEMM stage needs is one contiguous 2 KB run per 16-token tile:
//   [m_tile][k32 block][16 rows][32 halves]
// ---------------------------------------------------------------------------------------
__host__ __device__ inline size_t a16_index(int m, int k, int kq_per_tile) {
    return (((size_t)(m >> 4) * kq_per_tile + (k >> 5)) * 16 + (m & 15)) * 32 + (k & 31);
}

__device__ __forceinline__ __half f2h_sat(float v) {
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    return __float2half_rn(v);
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __halves2half2(f2h_sat(a), f2h_sat(b));
    return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------------------------------
// mbarrier + 1-D bulk TMA (cp.async.bulk -> SASS UBLKCP)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity)
        : "memory");
}
// L2 policy for streamed-once weights
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// D[16x8] += A[16x16] * B[16x8], f16 operands, f32 accumulate (legacy tensor path: HMMA)
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Programmatic dependent launch: everything before this call may overlap the tail of the
// preceding kernel in the stream/graph (weights are immutable, so weight prefetch may).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum; `red` is >= 32 floats of shared memory; result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();   // protect `red` from the previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (lane < nw) ? red[lane] : 0.f;
    t = warp_sum(t);
    return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (lane < nw) ? red[lane] : -INFINITY;
    t = warp_max(t);
    return t;
}

// ---------------------------------------------------------------------------------------
// activations (exact-ish: expf/tanhf, not the fast intrinsics, to stay inside the 1e-3 budget)
// ---------------------------------------------------------------------------------------
enum Act : int { ACT_NONE = 0, ACT_TANH, ACT_SIGMOID, ACT_SILU, ACT_RELU2, ACT_EXPNEGEXP, ACT_V7DECAY };

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case ACT_TANH: return tanhf(v);
        case ACT_SIGMOID: return sigmoidf_(v);
        case ACT_SILU: return v * sigmoidf_(v);
        case ACT_RELU2: { float r = fmaxf(v, 0.f); return r * r; }
        case ACT_EXPNEGEXP: return expf(-expf(v));                       // v6 decay
        case ACT_V7DECAY: return expf(-0.606531f * sigmoidf_(v));        // v7 decay, exp(-0.5)
        default: return v;
    }
}

}  // namespace b200
