// Skinny projection GEMM for decode:  Y[T, N] = X[T, K] (f16) * W[N, K]^T (f16), f32 accumulate.
//
// Replaces web-rwkv's WGSL `matmul_vec_fp16` / `matmul_mat_fp16` dispatches that the reference
// reaches through `Runtime::infer` (reference crates/ai00-core/src/run.rs:1143): the R/K/V/G/O,
// channel-mix, LoRA and head projections.
//
// HBM-bound by construction (T <= 64 rows per pass, each weight byte is used once per step), so
// the design is about keeping ~200 KB of weight bytes in flight per SM and never stalling the
// stream:
//   * weights are re-tiled ONCE at load into 16 KB stage blocks (128 output rows x 64 k) already
//     in the order the tensor-core B fragments are read, so a pipeline stage is ONE contiguous
//     1-D bulk TMA copy (cp.async.bulk -> UBLKCP) and shared-memory reads are conflict-free
//     128-bit loads with no ldmatrix/swizzle;
//   * activations use the A16 layout (common.cuh) so a stage's X slice is one 2 KB bulk copy;
//   * one producer lane drives a 9-12 stage mbarrier ring, 8 consumer warps issue
//     mma.sync.m16n8k16 (tokens are the M=16 operand; the k index inside each 32-wide block is
//     permuted identically for A and B so both are plain 16-byte row chunks);
//   * work is split stream-K style: the launch's stage blocks (all segments, all tiles) form one
//     linear sequence cut into gridDim.x equal contiguous ranges, so every SM streams the same
//     number of bytes whatever the matrix shapes;
//   * tiles cut across CTAs are reduced deterministically: each contributor writes its partial
//     tile to an L2-resident workspace, bumps a per-tile counter, and the LAST arriver sums the
//     partials in fixed slot order and runs the fused epilogue (no spinning, no float atomics).
//   * a launch carries up to 8 "segments" (independent matrices, own input, K and epilogue) so
//     R/K/V/G + decay-LoRA, or the five ddlerp LoRAs, go out as ONE kernel.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int GEMM_BN = 128;                 // output rows (weight rows) per tile
constexpr int GEMM_BK = 64;                  // k per stage block
constexpr int GEMM_WBYTES = GEMM_BN * GEMM_BK * 2;   // 16 KB
constexpr int GEMM_ABYTES = 16 * GEMM_BK * 2;        // 2 KB per 16-token tile
constexpr int GEMM_CONSUMER_WARPS = 8;
constexpr int GEMM_THREADS = (GEMM_CONSUMER_WARPS + 1) * 32;
constexpr int GEMM_MAX_SEG = 8;
constexpr int GEMM_SMEM_BUDGET = 221184;     // 216 KB for stages

enum OutMode : int {
    OUT_F32 = 0,        // out[m*ldo + n] = y                              (f32 row-major)
    OUT_A16 = 1,        // f16 A16 layout (input of the next projection), optional column groups
    OUT_LERP_A16 = 2,   // v6 ddlerp: f16( xx + sx * (mu[n] + y) ) in A16 layout
};

struct GemmSeg {
    const __half* A;      // A16 activations for this segment
    int KB;               // 64-wide k blocks (K padded up)
    int tiles;            // 128-row output tiles (N padded up)
    int N;                // valid output columns
    int blk_begin;        // first linear stage block of this segment
    int tile_begin;       // first global tile index of this segment
    int out_mode;
    int act;
    const float* bias;    // optional [N], added before the activation
    void* out;
    int ldo;              // OUT_F32: row stride (floats); A16 modes: k32 blocks per m-tile of the destination
    int grp;              // OUT_A16: columns per destination matrix (0 = single matrix)
    int grp_stride;       // OUT_A16: halves between destination matrices
    const float* aux0;    // OUT_LERP_A16: xx [T, ld_aux]
    const float* aux1;    // OUT_LERP_A16: sx [T, ld_aux]
    const float* aux2;    // OUT_LERP_A16: mu [N]
    int ld_aux;
};

struct GemmParams {
    const uint8_t* W;       // packed stage blocks, linear order
    int nseg;
    int total_blocks;
    int max_contrib;        // workspace slots per tile
    float* ws;              // [total_tiles][max_contrib][MT*16][128] partial tiles
    unsigned* counters;     // [total_tiles], zero between launches
    const int* nrows;       // device: valid token rows
    GemmSeg seg[GEMM_MAX_SEG];
};

template <int MT>
struct GemmCfg {
    static constexpr int STAGE_BYTES = GEMM_WBYTES + MT * GEMM_ABYTES;
    static constexpr int NSTAGE = (GEMM_SMEM_BUDGET / STAGE_BYTES) > 12 ? 12 : (GEMM_SMEM_BUDGET / STAGE_BYTES);
    static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 2 * NSTAGE * 8 + 64;
};

__device__ __forceinline__ int gemm_find_seg(const GemmParams& p, int b) {
    int s = 0;
#pragma unroll 1
    while (s + 1 < p.nseg && b >= p.seg[s + 1].blk_begin) ++s;
    return s;
}

// one (row, two adjacent columns) result -> fused epilogue
__device__ __forceinline__ void gemm_store2(const GemmSeg& sg, int m, int n, float v0, float v1) {
    if (n >= sg.N) return;
    const bool has1 = (n + 1) < sg.N;
    if (sg.bias) {
        v0 += sg.bias[n];
        if (has1) v1 += sg.bias[n + 1];
    }
    v0 = apply_act(v0, sg.act);
    v1 = apply_act(v1, sg.act);
    if (sg.out_mode == OUT_F32) {
        float* o = reinterpret_cast<float*>(sg.out) + (size_t)m * sg.ldo + n;
        if (has1 && ((reinterpret_cast<uintptr_t>(o) & 7) == 0)) {
            *reinterpret_cast<float2*>(o) = make_float2(v0, v1);
        } else {
            o[0] = v0;
            if (has1) o[1] = v1;
        }
        return;
    }
    if (sg.out_mode == OUT_LERP_A16) {
        const size_t a = (size_t)m * sg.ld_aux + n;
        v0 = sg.aux0[a] + sg.aux1[a] * (sg.aux2[n] + v0);
        if (has1) v1 = sg.aux0[a + 1] + sg.aux1[a + 1] * (sg.aux2[n + 1] + v1);
    }
    __half* base = reinterpret_cast<__half*>(sg.out);
    int nn = n;
    if (sg.grp > 0) {
        const int gi = n / sg.grp;
        base += (size_t)gi * sg.grp_stride;
        nn = n - gi * sg.grp;
    }
    __half* o = base + a16_index(m, nn, sg.ldo);
    if (has1) {
        *reinterpret_cast<uint32_t*>(o) = pack_h2(v0, v1);    // n even -> 4-byte aligned, same 32-block
    } else {
        o[0] = f2h_sat(v0);
    }
}

template <int MT>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_kernel(const __grid_constant__ GemmParams p) {
    using Cfg = GemmCfg<MT>;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int s_last;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full_bar = smem_base + Cfg::NSTAGE * Cfg::STAGE_BYTES;
    const uint32_t empty_bar = full_bar + Cfg::NSTAGE * 8;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long TB = p.total_blocks;
    const int G = gridDim.x, cta = blockIdx.x;
    const int b0 = (int)((long long)cta * TB / G);
    const int b1 = (int)((long long)(cta + 1) * TB / G);

    if (tid == 0) {
        for (int s = 0; s < Cfg::NSTAGE; ++s) {
            mbar_init(full_bar + s * 8, 1);
            mbar_init(empty_bar + s * 8, GEMM_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == GEMM_CONSUMER_WARPS) {
        // ===================== producer: one lane streams stage blocks =====================
        if (lane == 0) {
            const uint64_t pol_w = l2_policy_evict_first();
            const uint64_t pol_a = l2_policy_evict_last();
            // Weights never change: the first ring-full of weight blocks is requested BEFORE
            // waiting on the preceding kernel (programmatic dependent launch), so HBM keeps
            // streaming across the kernel boundary.  Activations are fetched after the wait.
            const int npre = min(b1 - b0, Cfg::NSTAGE);
            for (int i = 0; i < npre; ++i) {
                const uint32_t st = smem_base + i * Cfg::STAGE_BYTES;
                mbar_expect_tx(full_bar + i * 8, Cfg::STAGE_BYTES);
                bulk_g2s_hint(st, p.W + (size_t)(b0 + i) * GEMM_WBYTES, GEMM_WBYTES, full_bar + i * 8, pol_w);
            }
            pdl_wait();
            int seg = gemm_find_seg(p, b0);
            for (int b = b0, it = 0; b < b1; ++b, ++it) {
                while (seg + 1 < p.nseg && b >= p.seg[seg + 1].blk_begin) ++seg;
                const GemmSeg& sg = p.seg[seg];
                const int kb = (b - sg.blk_begin) % sg.KB;
                const int stage = it % Cfg::NSTAGE;
                const uint32_t st = smem_base + stage * Cfg::STAGE_BYTES;
                const uint32_t fb = full_bar + stage * 8;
                if (it >= Cfg::NSTAGE) {
                    mbar_wait(empty_bar + stage * 8, ((it / Cfg::NSTAGE) - 1) & 1);
                    mbar_expect_tx(fb, Cfg::STAGE_BYTES);
                    bulk_g2s_hint(st, p.W + (size_t)b * GEMM_WBYTES, GEMM_WBYTES, fb, pol_w);
                }
                const int kq_tile = sg.KB * 2;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
                    bulk_g2s_hint(st + GEMM_WBYTES + mt * GEMM_ABYTES,
                                  sg.A + ((size_t)mt * kq_tile + 2 * kb) * 512, GEMM_ABYTES, fb, pol_a);
            }
        }
        return;
    }

    // ============================== consumers: 8 warps =====================================
    pdl_wait();
    const int nrows = *p.nrows;
    const int g = lane >> 2, q = lane & 3;
    float acc[MT][2][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][j][e] = 0.f;

    int seg = gemm_find_seg(p, b0);
    for (int b = b0, it = 0; b < b1; ++b, ++it) {
        while (seg + 1 < p.nseg && b >= p.seg[seg + 1].blk_begin) ++seg;
        const GemmSeg& sg = p.seg[seg];
        const int rel = b - sg.blk_begin;
        const int tile_local = rel / sg.KB;
        const int kb = rel - tile_local * sg.KB;
        const int stage = it % Cfg::NSTAGE;
        const uint32_t st = smem_base + stage * Cfg::STAGE_BYTES;
        mbar_wait(full_bar + stage * 8, (it / Cfg::NSTAGE) & 1);
#pragma unroll
        for (int kq = 0; kq < 2; ++kq) {
            uint4 wv[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) wv[j] = lds128(st + (((warp * 2 + j) * 2 + kq) * 512) + lane * 16);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const uint32_t ab = st + GEMM_WBYTES + mt * GEMM_ABYTES + kq * 1024 + lane * 16;
                const uint4 alo = lds128(ab);          // token row g,   k chunk q
                const uint4 ahi = lds128(ab + 512);    // token row g+8, k chunk q
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    mma_16816(acc[mt][j], alo.x, ahi.x, alo.y, ahi.y, wv[j].x, wv[j].y);
                    mma_16816(acc[mt][j], alo.z, ahi.z, alo.w, ahi.w, wv[j].z, wv[j].w);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar + stage * 8);

        if (kb == sg.KB - 1 || b == b1 - 1) {
            // ---- this CTA's share of the tile is complete ----
            const long long tb0 = (long long)sg.blk_begin + (long long)tile_local * sg.KB;
            const int c_first = (int)(((tb0 + 1) * G - 1) / TB);
            const int c_last = (int)(((tb0 + sg.KB) * G - 1) / TB);
            const int ncontrib = c_last - c_first + 1;
            const int gtile = sg.tile_begin + tile_local;
            bool do_epilogue = true;
            if (ncontrib > 1) {
                float* wsl = p.ws + ((size_t)gtile * p.max_contrib + (cta - c_first)) * (MT * 16 * GEMM_BN);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int col = (warp * 2 + j) * 8 + 2 * q;
                        *reinterpret_cast<float2*>(wsl + (mt * 16 + g) * GEMM_BN + col) =
                            make_float2(acc[mt][j][0], acc[mt][j][1]);
                        *reinterpret_cast<float2*>(wsl + (mt * 16 + g + 8) * GEMM_BN + col) =
                            make_float2(acc[mt][j][2], acc[mt][j][3]);
                    }
                __threadfence();
                named_bar_sync(1, GEMM_CONSUMER_WARPS * 32);
                if (tid == 0) {
                    const unsigned old = atomicAdd(p.counters + gtile, 1u);
                    s_last = (old == (unsigned)(ncontrib - 1));
                    if (s_last) p.counters[gtile] = 0;     // ready for the next launch
                }
                named_bar_sync(1, GEMM_CONSUMER_WARPS * 32);
                do_epilogue = (s_last != 0);
                if (do_epilogue) {
                    __threadfence();
                    const float* ws0 = p.ws + (size_t)gtile * p.max_contrib * (MT * 16 * GEMM_BN);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int col = (warp * 2 + j) * 8 + 2 * q;
                            float2 lo = make_float2(0.f, 0.f), hi = make_float2(0.f, 0.f);
                            for (int s = 0; s < ncontrib; ++s) {   // fixed order -> deterministic
                                const float* w_ = ws0 + (size_t)s * (MT * 16 * GEMM_BN);
                                const float2 a = __ldcg(reinterpret_cast<const float2*>(w_ + (mt * 16 + g) * GEMM_BN + col));
                                const float2 c = __ldcg(reinterpret_cast<const float2*>(w_ + (mt * 16 + g + 8) * GEMM_BN + col));
                                lo.x += a.x; lo.y += a.y; hi.x += c.x; hi.y += c.y;
                            }
                            acc[mt][j][0] = lo.x; acc[mt][j][1] = lo.y;
                            acc[mt][j][2] = hi.x; acc[mt][j][3] = hi.y;
                        }
                }
                // s_last is rewritten only after the next pair of barriers: safe to fall through
            }
            if (do_epilogue) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int n = tile_local * GEMM_BN + (warp * 2 + j) * 8 + 2 * q;
                        const int m0 = mt * 16 + g;
                        if (m0 < nrows) gemm_store2(sg, m0, n, acc[mt][j][0], acc[mt][j][1]);
                        if (m0 + 8 < nrows) gemm_store2(sg, m0 + 8, n, acc[mt][j][2], acc[mt][j][3]);
                    }
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[mt][j][e] = 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------
// One-time weight re-tiling:  W[N, K] row-major f16  ->  stage blocks
//   block(tile, kb) = [nb 16][kq 2][row 8][k 32] halves, zero padded.
// Supports a row-parallel / column-parallel shard: source sub-matrix rows [n0, n0+N), cols [k0, k0+K)
// of a matrix with row stride ld.
// ---------------------------------------------------------------------------------------
__global__ void repack_weight_kernel(const __half* __restrict__ src, int ld, int n0, int k0, int N, int K,
                                     int tiles, int KB, uint4* __restrict__ dst) {
    // one thread per 16-byte chunk (8 halves)
    const size_t nchunk = (size_t)tiles * KB * (GEMM_WBYTES / 16);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nchunk; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int c8 = r % 4; r /= 4;        // 8-half chunk within the 32-k row
        const int row = r % 8; r /= 8;
        const int kq = r % 2; r /= 2;
        const int nb = r % 16; r /= 16;
        const int kb = r % KB; r /= KB;
        const int tile = (int)r;
        const int n = tile * GEMM_BN + nb * 8 + row;
        const int k = kb * GEMM_BK + kq * 32 + c8 * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (n < N) {
            const __half* s = src + (size_t)(n0 + n) * ld + k0 + k;
            if (k + 8 <= K && ((reinterpret_cast<uintptr_t>(s) & 15) == 0)) {
                v = *reinterpret_cast<const uint4*>(s);
            } else {
                __half tmp[8];
                for (int e = 0; e < 8; ++e) tmp[e] = (k + e < K) ? s[e] : __float2half(0.f);
                v = *reinterpret_cast<uint4*>(tmp);
            }
        }
        dst[i] = v;
    }
}

}  // namespace b200
