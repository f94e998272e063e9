// Persistent whole-step decode kernel ("megakernel"): ONE cooperative launch runs every layer of
// the model for a decode step (<= 16 tokens), one CTA per SM.
//
// Why: at batch 16 a 7B decode step is ~330 dependent launches of 5-35 us each; measured on B200
// every kernel boundary costs ~6 us of drained HBM pipeline (profiles/), i.e. more time than the
// 14.7 GB weight stream itself.  Here the model is a device-side "program" of phases
// (LN/mix rows, projection GEMMs, WKV) separated by grid barriers, and the weight stream never
// stops:
//   * warp 8 of every CTA is a producer that walks the whole program ahead of the consumers and
//     keeps the 5-stage / 180 KB shared-memory ring full with 1-D bulk TMA copies of the CTA's
//     weight stage blocks (immutable, so they may be fetched arbitrarily early), WKV head states
//     and decay-LoRA slices; grid barriers and the small phases are hidden behind the ring;
//   * only the 4 KB activation slice of a GEMM stage depends on the previous phase: a second
//     producer lane (warp 10) issues those copies once the consumers have passed the phase's barrier;
//   * warp 9 issues the tcgen05.mma stream for every GEMM phase (accumulators double-buffered in
//     TMEM), warps 0-3 drain them (gemm_epilogue_role) and, with warps 4-7, run the LN/mix rows and
//     the WKV units; all of it reuses the device functions of the stand-alone kernels, so both
//     paths compute the same results;
//   * the v6 decay LoRA stage 2 (w = exp(-exp(time_decay + Wd2 tanh(.)))) is evaluated inside the
//     WKV phase from a TMA-staged k-major slice of time_decay_w2, removing one phase per layer.
// Replaces the whole per-step dispatch chain web-rwkv records for `Runtime::infer`
// (reference crates/ai00-core/src/run.rs:1143).
#pragma once
#include "gemm.cuh"
#include "lora.cuh"
#include "misc.cuh"
#include "mix.cuh"
#include "wkv.cuh"

namespace b200 {

constexpr int MEGA_STAGE_BYTES = GEMM_WBYTES + GEMM_ABYTES;     // 36 KB: weight block + 1 token tile
constexpr int MEGA_NSTAGE = 5;
constexpr int MEGA_CONSUMER_WARPS = CONSUMER_THREADS / 32;      // warps 0-7 (0-3 double as GEMM epilogue warps)
constexpr int MEGA_PRODUCER_WARP = MEGA_CONSUMER_WARPS;         // warp 8: TMA producer of weight blocks / states / LoRA slices
constexpr int MEGA_MMA_WARP = MEGA_CONSUMER_WARPS + 1;          // warp 9: tcgen05.mma issuer, owns the TMEM allocation
constexpr int MEGA_APROD_WARP = MEGA_CONSUMER_WARPS + 2;        // warp 10: TMA producer of the activation slices
constexpr int MEGA_THREADS = (MEGA_CONSUMER_WARPS + 3) * 32;
constexpr int MEGA_TMEM_COLS = 32;                              // 2 x (128 lanes x 16 token columns) accumulators
constexpr int MEGA_SMEM_BYTES = MEGA_NSTAGE * MEGA_STAGE_BYTES + (2 * MEGA_NSTAGE + 4) * 8 + 16 + 64;
constexpr int MEGA_MAX_DD = 128;                                // decay LoRA rank limit: one k-major slice per ring stage
constexpr int MEGA_MAX_TOK = 16;
constexpr int MEGA_MAX_GROUP = 8;                               // slots per WKV unit

enum PhaseType : int { PH_EMBED = 0, PH_LN = 1, PH_GEMM = 2, PH_WKV = 3, PH_LNOUT = 4, PH_SMALLN = 5, PH_SMALLK = 6, PH_TPBAR = 7 };
struct Phase {
    int type, idx;
};
struct GemmLaunchDev {
    GemmParams p;
    int ncta;          // CTAs that take part in this launch
    int pad;
};
struct MegaParams {
    const Phase* phases;
    int nphase;
    int version;
    const EmbedParams* embed;
    const LnMixParams* ln;
    const GemmLaunchDev* gemm;
    const WkvParams* wkv;
    const LnOutParams* lnout;
    const SmallNParams* smalln;
    const SmallKParams* smallk;
    TpBar tp;          // tensor-parallel rendezvous (world > 1)
    unsigned* gbar;    // [0] arrivals, [1] generation (sense-reversing grid barrier, never reset)
    MetaView meta;
    unsigned long long* trace;   // optional [4 CTAs][nphase][12]: t(phase entered), t(work done), t(barrier passed), 8 in-phase stamps
};

// ring stages one CTA consumes in a WKV phase (the MMA warp and the activation producer skip over them)
__device__ __forceinline__ int wkv_phase_stages(const int cta, const int G, const int H, const int nslots, const int gs, const bool v6) {
    const int units = H * ((nslots + gs - 1) / gs);
    int n = 0;
    for (int u = cta; u < units; u += G) {
        const int s0 = (u / H) * gs;
        n += (v6 ? 1 : 0) + min(gs, nslots - s0);
    }
    return n;
}

// ---------------------------------------------------------------------------------------
// producers (one thread each).  A single thread can issue a bulk copy only every ~0.3 us
// (measured, profiles/r01_stream_microbench.md), so the two kinds of payload get their own lane:
//   W lane: walks the whole program ahead of the consumers.  Weight blocks, head states and LoRA
//           slices are immutable / untouched during the step, so each is requested as soon as a
//           ring slot is free; it arms the slot's full barrier and publishes its sequence number.
//   A lane: the 4 KB activation slice of a GEMM stage is produced by the previous phase: issued
//           once the consumers have passed that phase's grid barrier (sh.s_started) and the W lane
//           has armed the slot (sh.s_wseq).
// ---------------------------------------------------------------------------------------
__device__ void mega_w_producer(const MegaParams& mp, const int cta, const int G, const uint32_t smem_base,
                                const uint32_t full_bar, const uint32_t empty_bar, volatile unsigned* s_wseq) {
    const uint64_t pol_w = l2_policy_evict_first();
    const uint64_t pol_a = l2_policy_evict_last();
    const int nslots = mp.meta.nslots();
    const int gs = max(1, min(MEGA_MAX_GROUP, mp.meta.base[3]));
    unsigned seq = 0;
    int slot = 0;
    uint32_t parity = 1;              // parity of the previous use of the slot's empty barrier
    auto acquire = [&]() -> uint32_t {
        if (seq >= (unsigned)MEGA_NSTAGE) mbar_wait(empty_bar + slot * 8, parity, 21);
        return smem_base + slot * MEGA_STAGE_BYTES;
    };
    auto commit = [&]() {
        ++seq;
        __threadfence_block();
        *s_wseq = seq;
        if (++slot == MEGA_NSTAGE) { slot = 0; parity ^= 1u; }
    };
    for (int ph = 0; ph < mp.nphase; ++ph) {
        const Phase P = mp.phases[ph];
        if (P.type == PH_GEMM) {
            const GemmLaunchDev& g = mp.gemm[P.idx];
            if (cta >= g.ncta) continue;
            const long long TB = g.p.total_blocks;
            const int b0 = (int)((long long)cta * TB / g.ncta);
            const int b1 = (int)((long long)(cta + 1) * TB / g.ncta);
            const uint8_t* wsrc = g.p.W + (size_t)b0 * GEMM_WBYTES;
            for (int b = b0; b < b1; ++b) {
                const uint32_t st = acquire();
                const uint32_t fb = full_bar + slot * 8;
                mbar_expect_tx(fb, MEGA_STAGE_BYTES);
                bulk_g2s_hint(st, wsrc, GEMM_WBYTES, fb, pol_w);
                wsrc += GEMM_WBYTES;
                commit();
            }
        } else if (P.type == PH_WKV) {
            const WkvParams& w = mp.wkv[P.idx];
            const int units = w.H * ((nslots + gs - 1) / gs);
            for (int u = cta; u < units; u += G) {
                const int h = u % w.H;
                const int s0 = (u / w.H) * gs;
                const int ns = min(gs, nslots - s0);
                if (mp.version == 6) {
                    const uint32_t st = acquire();
                    const uint32_t fb = full_bar + slot * 8;
                    const uint32_t bytes = (uint32_t)(WKV_N * w.Dd * 2);
                    mbar_expect_tx(fb, bytes);
                    bulk_g2s_hint(st, w.wd2t + (size_t)h * WKV_N * w.Dd, bytes, fb, pol_a);
                    commit();
                }
                for (int sl = 0; sl < ns; ++sl) {
                    const int slot_id = mp.meta.slot_id()[s0 + sl];
                    const uint32_t st = acquire();
                    const uint32_t fb = full_bar + slot * 8;
                    mbar_expect_tx(fb, WKV_N * WKV_N * 4);
                    bulk_g2s_hint(st, w.state + ((size_t)slot_id * w.H + h) * (WKV_N * WKV_N), WKV_N * WKV_N * 4, fb, pol_w);
                    commit();
                }
            }
        }
    }
}

__device__ void mega_a_producer(const MegaParams& mp, const int cta, const int G, const uint32_t smem_base,
                                const uint32_t full_bar, volatile int* s_started, volatile unsigned* s_wseq) {
    const uint64_t pol_a = l2_policy_evict_last();
    const int nslots = mp.meta.nslots();
    const int gs = max(1, min(MEGA_MAX_GROUP, mp.meta.base[3]));
    unsigned seq = 0;
    int slot = 0;
    for (int ph = 0; ph < mp.nphase; ++ph) {
        const Phase P = mp.phases[ph];
        if (P.type == PH_GEMM) {
            const GemmLaunchDev& g = mp.gemm[P.idx];
            if (cta >= g.ncta) continue;
            const long long TB = g.p.total_blocks;
            const int b0 = (int)((long long)cta * TB / g.ncta);
            const int b1 = (int)((long long)(cta + 1) * TB / g.ncta);
            if (b0 >= b1) continue;
            int seg = gemm_find_seg(g.p, b0);
            const __half* A = g.p.seg[seg].A;
            int KB = g.p.seg[seg].KB;
            int kb = (b0 - g.p.seg[seg].blk_begin) % KB;
            int left = g.p.seg[seg].blk_begin + g.p.seg[seg].tiles * KB - b0;
            { SpinGuard sg_; while (*s_started < ph) { __nanosleep(40); sg_.poll(WD_A_STARTED, (unsigned)ph, (unsigned)*s_started, seq); } }   // the phase that writes these activations is done
            fence_proxy_async();
            for (int b = b0; b < b1; ++b) {
                { SpinGuard sg_; while (*s_wseq <= seq) { __nanosleep(20); sg_.poll(WD_A_WSEQ, (unsigned)ph, *s_wseq, seq); } }   // slot armed by the W lane
                bulk_g2s_hint(smem_base + slot * MEGA_STAGE_BYTES + GEMM_WBYTES, A + (size_t)(GEMM_K8 * kb) * 128, GEMM_ABYTES,
                              full_bar + slot * 8, pol_a);
                ++seq;
                if (++slot == MEGA_NSTAGE) slot = 0;
                if (++kb == KB) kb = 0;
                if (--left == 0 && b + 1 < b1) {
                    ++seg;
                    A = g.p.seg[seg].A;
                    KB = g.p.seg[seg].KB;
                    kb = 0;
                    left = g.p.seg[seg].tiles * KB;
                }
            }
        } else if (P.type == PH_WKV) {
            const int n = wkv_phase_stages(cta, G, mp.wkv[P.idx].H, nslots, gs, mp.version == 6);
            seq += n;
            slot = (slot + n) % MEGA_NSTAGE;
        }
    }
}

// ---------------------------------------------------------------------------------------
// grid barrier (all CTAs co-resident: cooperative launch, one CTA per SM).  Sense-reversing:
// gbar[0] counts arrivals, gbar[1] is the generation; release on arrival, acquire on departure.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mega_grid_barrier(unsigned* gbar, const int G, volatile int* s_started, const int next_phase) {
    fence_proxy_async();                          // generic-proxy writes -> visible to later bulk (async-proxy) reads
    named_bar_sync(1, CONSUMER_THREADS);          // CTA-scope: every consumer's writes happen-before thread 0's release
    if (threadIdx.x == 0) {
        const unsigned gen = ld_relaxed_gpu(gbar + 1);      // cannot advance before this CTA arrives
        const unsigned old = atom_add_acq_rel_gpu(gbar, 1u);
        if (old == (unsigned)(G - 1)) {
            st_relaxed_gpu(gbar, 0u);
            st_release_gpu(gbar + 1, gen + 1);
        } else {
            SpinGuard sg_;
            while (ld_acquire_gpu(gbar + 1) == gen) sg_.poll(WD_GRIDBAR, (unsigned)next_phase, gen, old);
        }
        __threadfence();                          // acquire for the whole CTA; drops this SM's stale L1 lines
        *s_started = next_phase;
    }
    named_bar_sync(1, CONSUMER_THREADS);
}

// consumer-side release of a ring slot (WKV stages): every barrier of the ring has ONE arrival
// (GEMM stages are released by tcgen05.commit), so sync the consumers and let thread 0 arrive
__device__ __forceinline__ void ring_release(const uint32_t empty_bar, RingPos& rp) {
    named_bar_sync(1, CONSUMER_THREADS);
    if (threadIdx.x == 0) mbar_arrive(empty_bar + rp.stage * 8);
    rp.advance<MEGA_NSTAGE>(1);
}

constexpr int MEGA_MAX_C = 4096;                                // widest model of this path (LN rows live in registers)
constexpr int MEGA_PRE_ARRAYS = 7;                              // r, k, v, g, w, a, nu
struct MegaShared {
    WkvShared wkv;
    float pre[MEGA_PRE_ARRAYS * MEGA_MAX_TOK * WKV_N];   // per-token head vectors of the current WKV unit
    float red[32];
    float w_s[MEGA_MAX_TOK * WKV_N];              // WKV phases (v6): decays of the current unit [token][64]
    __half d_s[MEGA_MAX_TOK * MEGA_MAX_DD];       // tanh(Wd1 xw) rows of the unit's tokens
    int tok_s[MEGA_MAX_TOK];
    int s_last;
    int s_started;
    unsigned s_wseq;
    unsigned long long stamps[8];
};

template <int VER>
__device__ void mega_wkv_phase(const MegaParams& mp, const WkvParams& w, const int cta, const int G, const uint8_t* smem_gen,
                               const uint32_t smem_base, const uint32_t full_bar, const uint32_t empty_bar, RingPos& rp,
                               MegaShared& sh) {
    const int tid = threadIdx.x;
    const int ig = tid >> 4, j4 = tid & 15;
    const MetaView& mv = mp.meta;
    const int nslots = mv.nslots();
    const int gs = max(1, min(MEGA_MAX_GROUP, mv.base[3]));
    const int units = w.H * ((nslots + gs - 1) / gs);
    auto stamp = [&](int i) { if (mp.trace && tid == 0) sh.stamps[i] = globaltimer_ns(); };
    stamp(0);
    for (int u = cta; u < units; u += G) {
        const int h = u % w.H;
        const int s0 = (u / w.H) * gs;
        const int ns = min(gs, nslots - s0);
        // ---- tokens of this unit, then ONE batched gather of everything the recurrence needs ----
        if (tid < MEGA_MAX_TOK) {
            int tok = -1, acc = 0;
            for (int sl = 0; sl < ns; ++sl) {
                const int t0 = mv.slot_start()[s0 + sl], nt = mv.slot_count()[s0 + sl];
                if (tid >= acc && tid < acc + nt) tok = t0 + (tid - acc);
                acc += nt;
            }
            sh.tok_s[tid] = tok;
        }
        cta_sync<true>();
        int ntok = 0;
        for (int sl = 0; sl < ns; ++sl) ntok += mv.slot_count()[s0 + sl];
        ntok = min(ntok, MEGA_MAX_TOK);
        {
            constexpr int NA = (VER == 7) ? 7 : 4;
            const float* srcs[7] = {w.r, w.k, w.v, w.g, w.w, w.a, w.nu};
            const int per = ntok * (WKV_N / 4);
#pragma unroll
            for (int arr = 0; arr < NA; ++arr) {
                if (VER == 7 && arr == 6 && w.layer0) continue;
                for (int i = tid; i < per; i += CONSUMER_THREADS) {
                    const int lt = i / (WKV_N / 4), c4 = (i - lt * (WKV_N / 4)) * 4;
                    const float4 val = *reinterpret_cast<const float4*>(srcs[arr] + (size_t)sh.tok_s[lt] * w.ld + h * WKV_N + c4);
                    *reinterpret_cast<float4*>(&sh.pre[(arr * MEGA_MAX_TOK + lt) * WKV_N + c4]) = val;
                }
            }
        }
        if (VER == 6) {
            // ---- decay LoRA stage 2 for every token of this unit ----
            const int Dd = w.Dd;
            for (int i = tid; i < ntok * (Dd >> 1); i += CONSUMER_THREADS) {
                const int lt = i / (Dd >> 1), k = (i - lt * (Dd >> 1)) * 2;
                const uint32_t v2 = *reinterpret_cast<const uint32_t*>(w.d1 + a16_index(sh.tok_s[lt], k, w.d1_kq));
                *reinterpret_cast<uint32_t*>(&sh.d_s[lt * MEGA_MAX_DD + k]) = v2;
            }
            stamp(1);
            mbar_wait(full_bar + rp.stage * 8, rp.phase, 22);
            cta_sync<true>();
            stamp(2);
            const __half* wt = reinterpret_cast<const __half*>(smem_gen + rp.stage * MEGA_STAGE_BYTES);
            const int c = tid & (WKV_N - 1);
            for (int lt = tid >> 6; lt < ntok; lt += CONSUMER_THREADS / WKV_N) {
                float acc = 0.f;
                const __half* d = &sh.d_s[lt * MEGA_MAX_DD];
#pragma unroll 8
                for (int k = 0; k < Dd; ++k) acc = fmaf(__half2float(wt[k * WKV_N + c]), __half2float(d[k]), acc);
                sh.w_s[lt * WKV_N + c] = expf(-expf(w.decay_bias[h * WKV_N + c] + acc));
            }
            ring_release(empty_bar, rp);       // (its barrier also publishes w_s)
            stamp(3);
        }
        if (VER != 6) cta_sync<true>();        // publish sh.pre (v6: done by the LoRA stage's barriers)
        int lt0 = 0;
        for (int sl = 0; sl < ns; ++sl) {
            const int si = s0 + sl;
            const int slot = mv.slot_id()[si];
            const int t0 = mv.slot_start()[si], nt = mv.slot_count()[si];
            mbar_wait(full_bar + rp.stage * 8, rp.phase, 23);
            if (sl == 0) stamp(4);
            const uint32_t st = smem_base + rp.stage * MEGA_STAGE_BYTES;
            float4 m[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint4 raw = lds128(st + ((ig * 4 + e) * WKV_N + j4 * 4) * 4);
                m[e] = make_float4(__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z), __uint_as_float(raw.w));
            }
            ring_release(empty_bar, rp);            // the patch is in registers: free the slot early
            wkv_slot<VER, true>(w, h, t0, nt, reinterpret_cast<float (&)[4][4]>(m), sh.wkv, VER == 6 ? sh.w_s : nullptr, lt0, sh.pre, MEGA_MAX_TOK * WKV_N);
            float* M = w.state + ((size_t)slot * w.H + h) * (WKV_N * WKV_N);
#pragma unroll
            for (int e = 0; e < 4; ++e) __stcs(reinterpret_cast<float4*>(M + (ig * 4 + e) * WKV_N + j4 * 4), m[e]);
            if (sl == 0) stamp(5);
            if (sl == ns - 1) stamp(6);
            lt0 += nt;
        }
    }
}

template <int VER>
__global__ void __launch_bounds__(MEGA_THREADS, 1) mega_step_kernel(const __grid_constant__ MegaParams mp) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ MegaShared sh;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full_bar = smem_base + MEGA_NSTAGE * MEGA_STAGE_BYTES;
    const uint32_t empty_bar = full_bar + MEGA_NSTAGE * 8;
    const uint32_t tfull_bar = empty_bar + MEGA_NSTAGE * 8;
    const uint32_t tempty_bar = tfull_bar + 2 * 8;
    const uint32_t tmem_slot = tempty_bar + 2 * 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x;

    if (tid == 0) {
        for (int s = 0; s < MEGA_NSTAGE; ++s) {
            mbar_init(full_bar + s * 8, 1);
            mbar_init(empty_bar + s * 8, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar + s * 8, 1);
            mbar_init(tempty_bar + s * 8, GEMM_EPI_WARPS);
        }
        mbar_fence_init();
        sh.s_started = 0;
        sh.s_last = 0;
        sh.s_wseq = 0;
    }
    if (warp == MEGA_MMA_WARP) tc_alloc(tmem_slot, MEGA_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

    const int nslots = mp.meta.nslots();
    const int gs = max(1, min(MEGA_MAX_GROUP, mp.meta.base[3]));

    if (warp == MEGA_PRODUCER_WARP) {
        if (lane == 0) mega_w_producer(mp, cta, G, smem_base, full_bar, empty_bar, &sh.s_wseq);
    } else if (warp == MEGA_APROD_WARP) {
        if (lane == 0) mega_a_producer(mp, cta, G, smem_base, full_bar, &sh.s_started, &sh.s_wseq);
    } else if (warp == MEGA_MMA_WARP) {
        if (lane == 0) {
            RingPos rp{0, 0u};
            unsigned segcount = 0;
            for (int ph = 0; ph < mp.nphase; ++ph) {
                const Phase P = mp.phases[ph];
                if (P.type == PH_GEMM) {
                    const GemmLaunchDev& g = mp.gemm[P.idx];
                    if (cta >= g.ncta) continue;
                    const long long TB = g.p.total_blocks;
                    const int b0 = (int)((long long)cta * TB / g.ncta);
                    const int b1 = (int)((long long)(cta + 1) * TB / g.ncta);
                    if (b0 < b1) {
                        // This lane skips the stages of WKV phases, so it could get more than one use ahead
                        // of a ring slot, where a parity wait can no longer tell the uses apart: start a
                        // GEMM phase only once the consumers have entered it (all earlier uses retired).
                        {
                            volatile int* st_ = &sh.s_started;
                            SpinGuard sg_;
                            while (*st_ < ph) { __nanosleep(40); sg_.poll(WD_A_STARTED, (unsigned)ph, (unsigned)*st_, 9999u); }
                        }
                        gemm_mma_role<1, MEGA_NSTAGE, MEGA_STAGE_BYTES>(g.p, b0, b1, smem_base, full_bar, empty_bar, tfull_bar,
                                                                        tempty_bar, tmem_base, rp, segcount);
                    }
                } else if (P.type == PH_WKV) {
                    rp.advance<MEGA_NSTAGE>(wkv_phase_stages(cta, G, mp.wkv[P.idx].H, nslots, gs, mp.version == 6));
                }
            }
        }
    } else {
        // ================================ consumers: warps 0-7 ================================
        RingPos rp{0, 0u};
        unsigned segcount = 0;
        const int T = mp.meta.T();
        for (int ph = 0; ph < mp.nphase; ++ph) {
            const Phase P = mp.phases[ph];
            switch (P.type) {
                case PH_EMBED:
                    if (cta < T) embed_row<true>(*mp.embed, cta, sh.red);
                    break;
                case PH_LN:
                    if (cta < T) ln_mix_row<true>(mp.ln[P.idx], cta, sh.red, mp.trace ? sh.stamps : nullptr);
                    break;
                case PH_LNOUT:
                    if (cta < T) ln_out_row<true>(*mp.lnout, cta, sh.red);
                    break;
                case PH_SMALLN:
                    smalln_phase<true>(mp.smalln[P.idx], cta, G, sh.pre);
                    break;
                case PH_SMALLK:
                    smallk_phase<true>(mp.smallk[P.idx], cta, G, sh.pre);
                    break;
                case PH_TPBAR:          // the preceding grid barrier ordered every CTA's partial sums
                    if (cta == 0 && warp == 0) tp_barrier(mp.tp, lane);
                    break;
                case PH_GEMM: {
                    const GemmLaunchDev& g = mp.gemm[P.idx];
                    if (cta < g.ncta) {
                        const long long TB = g.p.total_blocks;
                        const int b0 = (int)((long long)cta * TB / g.ncta);
                        const int b1 = (int)((long long)(cta + 1) * TB / g.ncta);
                        if (b0 < b1) {
                            if (warp < GEMM_EPI_WARPS)
                                gemm_epilogue_role<1>(g.p, cta, g.ncta, b0, b1, tfull_bar, tempty_bar, tmem_base, segcount,
                                                      *g.p.nrows, &sh.s_last);
                            rp.advance<MEGA_NSTAGE>(b1 - b0);
                        }
                    }
                    break;
                }
                case PH_WKV: {
                    const WkvParams& w = mp.wkv[P.idx];
                    mega_wkv_phase<VER>(mp, w, cta, G, smem, smem_base, full_bar, empty_bar, rp, sh);
                    break;
                }
                default: break;
            }
            const int tsel = (cta == 0) ? 0 : (cta == 15 ? 1 : (cta == 74 ? 2 : (cta == G - 1 ? 3 : -1)));
            if (mp.trace && tsel >= 0 && tid == 0) {
                unsigned long long* tr = mp.trace + ((size_t)tsel * mp.nphase + ph) * 12;
                tr[0] = globaltimer_ns();
                for (int i = 0; i < 8; ++i) { tr[4 + i] = sh.stamps[i]; sh.stamps[i] = 0; }
            }
            if (ph + 1 < mp.nphase) mega_grid_barrier(mp.gbar, G, &sh.s_started, ph + 1);
            if (mp.trace && tsel >= 0 && tid == 0) mp.trace[((size_t)tsel * mp.nphase + ph) * 12 + 1] = globaltimer_ns();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MEGA_MMA_WARP) tc_dealloc(tmem_base, MEGA_TMEM_COLS);
}

}  // namespace b200
