// Persistent whole-step decode kernel ("megakernel"): ONE cooperative launch runs every layer of
// the model for a decode step (<= 16 tokens), one CTA per SM.
//
// Why: at batch 16 a 7B decode step is ~330 dependent launches of 5-35 us each; measured on B200
// every kernel boundary costs ~6 us of drained HBM pipeline (profiles/), i.e. more time than the
// 14.7 GB weight stream itself.  Here the model is a device-side "program" of phases
// (LN/mix rows, projection GEMMs, WKV) separated by grid barriers, and the weight stream never
// stops:
//   * warp 8 of every CTA is a producer that walks the whole program ahead of the consumers and
//     keeps the 11-stage / 200 KB shared-memory ring full with 1-D bulk TMA copies of the CTA's
//     weight stage blocks (immutable, so they may be fetched arbitrarily early), WKV head states
//     and decay-LoRA slices; grid barriers and the small phases are hidden behind the ring;
//   * only the 2 KB activation slice of a GEMM stage depends on the previous phase: a second
//     producer cursor issues those copies once the consumers have passed the phase's barrier;
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
#include "mix.cuh"
#include "wkv.cuh"

namespace b200 {

constexpr int MEGA_STAGE_BYTES = GEMM_WBYTES + GEMM_ABYTES;     // 18 KB: weight block + 1 token tile
constexpr int MEGA_NSTAGE = 11;
constexpr int MEGA_CONSUMER_WARPS = CONSUMER_THREADS / 32;      // warps 0-7 (0-3 double as GEMM epilogue warps)
constexpr int MEGA_PRODUCER_WARP = MEGA_CONSUMER_WARPS;         // warp 8: TMA producer
constexpr int MEGA_MMA_WARP = MEGA_CONSUMER_WARPS + 1;          // warp 9: tcgen05.mma issuer, owns the TMEM allocation
constexpr int MEGA_THREADS = (MEGA_CONSUMER_WARPS + 2) * 32;
constexpr int MEGA_TMEM_COLS = 32;                              // 2 x (128 lanes x 16 token columns) accumulators
constexpr int MEGA_SMEM_BYTES = MEGA_NSTAGE * MEGA_STAGE_BYTES + (2 * MEGA_NSTAGE + 4) * 8 + 16 + 64;
constexpr int MEGA_MAX_DD = 128;                                // decay LoRA rank limit: one k-major slice per ring stage
constexpr int MEGA_MAX_TOK = 16;
constexpr int MEGA_MAX_GROUP = 8;                               // slots per WKV unit

enum PhaseType : int { PH_EMBED = 0, PH_LN = 1, PH_GEMM = 2, PH_WKV = 3, PH_LNOUT = 4 };
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
    unsigned* gbar;    // [0] arrivals, [1] generation (sense-reversing grid barrier, never reset)
    MetaView meta;
    unsigned long long* trace;   // optional [4 CTAs][nphase][4]: t(work done), t(barrier passed), cycles blocked on the ring, cycles in phase
};

// ---------------------------------------------------------------------------------------
// producer (one thread per CTA).  It walks the program ahead of the consumers.  Main payloads
// (weight blocks, head states, LoRA slices) are immutable or not touched during the step, so
// they are issued as soon as a ring slot is free.  The 2 KB activation slice of a GEMM stage is
// produced by the previous phase: its source/phase are parked per ring slot and issued by a
// second cursor once the consumers have passed that phase's grid barrier (sh.s_started).
// ---------------------------------------------------------------------------------------
struct ProducerState {
    uint32_t smem_base, full_bar, empty_bar;
    volatile int* s_started;
    const __half* a_src[MEGA_NSTAGE];   // null: stage has no activation slice
    int a_ph[MEGA_NSTAGE];
    unsigned wseq, aseq;
    int wslot, aslot;
    uint32_t wparity;
    int seen_started;
    uint64_t pol_w, pol_a;
};

__device__ __forceinline__ void producer_drain_a(ProducerState& ps) {
    while (ps.aseq < ps.wseq) {
        const __half* src = ps.a_src[ps.aslot];
        if (src) {
            if (ps.a_ph[ps.aslot] > ps.seen_started) {
                const int st = *ps.s_started;
                if (st <= ps.seen_started) return;          // phase not started yet
                ps.seen_started = st;
                fence_proxy_async();                         // once per observed phase
                if (ps.a_ph[ps.aslot] > st) return;
            }
            bulk_g2s_hint(ps.smem_base + ps.aslot * MEGA_STAGE_BYTES + GEMM_WBYTES, src, GEMM_ABYTES,
                          ps.full_bar + ps.aslot * 8, ps.pol_a);
        }
        ++ps.aseq;
        if (++ps.aslot == MEGA_NSTAGE) ps.aslot = 0;
    }
}

// waits for the next ring slot, keeping the activation cursor moving; returns the slot's smem address
__device__ __forceinline__ uint32_t producer_acquire(ProducerState& ps) {
    if (ps.wseq >= (unsigned)MEGA_NSTAGE) {
        while (!mbar_test(ps.empty_bar + ps.wslot * 8, ps.wparity)) producer_drain_a(ps);
    }
    return ps.smem_base + ps.wslot * MEGA_STAGE_BYTES;
}
__device__ __forceinline__ void producer_commit(ProducerState& ps, const __half* a_src, int ph) {
    ps.a_src[ps.wslot] = a_src;
    ps.a_ph[ps.wslot] = ph;
    ++ps.wseq;
    if (++ps.wslot == MEGA_NSTAGE) { ps.wslot = 0; ps.wparity ^= 1; }
    producer_drain_a(ps);
}

__device__ void mega_producer(const MegaParams& mp, const int cta, const int G, const uint32_t smem_base,
                              const uint32_t full_bar, const uint32_t empty_bar, volatile int* s_started) {
    ProducerState ps;
    ps.smem_base = smem_base; ps.full_bar = full_bar; ps.empty_bar = empty_bar; ps.s_started = s_started;
    ps.wseq = ps.aseq = 0; ps.wslot = ps.aslot = 0; ps.wparity = 1; ps.seen_started = 0;
    ps.pol_w = l2_policy_evict_first();
    ps.pol_a = l2_policy_evict_last();
    const int nslots = mp.meta.nslots();
    const int gs = max(1, min(MEGA_MAX_GROUP, mp.meta.base[3]));
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
            const uint8_t* wsrc = g.p.W + (size_t)b0 * GEMM_WBYTES;
            for (int b = b0; b < b1; ++b) {
                const uint32_t st = producer_acquire(ps);
                const uint32_t fb = full_bar + ps.wslot * 8;
                mbar_expect_tx(fb, MEGA_STAGE_BYTES);
                bulk_g2s_hint(st, wsrc, GEMM_WBYTES, fb, ps.pol_w);
                wsrc += GEMM_WBYTES;
                producer_commit(ps, A + (size_t)(8 * kb) * 128, ph);
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
            const WkvParams& w = mp.wkv[P.idx];
            const int units = w.H * ((nslots + gs - 1) / gs);
            for (int u = cta; u < units; u += G) {
                const int h = u % w.H;
                const int s0 = (u / w.H) * gs;
                const int ns = min(gs, nslots - s0);
                if (mp.version == 6) {
                    const uint32_t st = producer_acquire(ps);
                    const uint32_t fb = full_bar + ps.wslot * 8;
                    const uint32_t bytes = (uint32_t)(WKV_N * w.Dd * 2);
                    mbar_expect_tx(fb, bytes);
                    bulk_g2s_hint(st, w.wd2t + (size_t)h * WKV_N * w.Dd, bytes, fb, ps.pol_a);
                    producer_commit(ps, nullptr, ph);
                }
                for (int sl = 0; sl < ns; ++sl) {
                    const int slot = mp.meta.slot_id()[s0 + sl];
                    const uint32_t st = producer_acquire(ps);
                    const uint32_t fb = full_bar + ps.wslot * 8;
                    mbar_expect_tx(fb, WKV_N * WKV_N * 4);
                    bulk_g2s_hint(st, w.state + ((size_t)slot * w.H + h) * (WKV_N * WKV_N), WKV_N * WKV_N * 4, fb, ps.pol_w);
                    producer_commit(ps, nullptr, ph);
                }
            }
        }
    }
    while (ps.aseq < ps.wseq) producer_drain_a(ps);
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
            while (ld_acquire_gpu(gbar + 1) == gen) { }
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

// ring stages one CTA consumes in a WKV phase (the MMA warp skips over them)
__device__ __forceinline__ int wkv_phase_stages(const int cta, const int G, const int H, const int nslots, const int gs, const bool v6) {
    const int units = H * ((nslots + gs - 1) / gs);
    int n = 0;
    for (int u = cta; u < units; u += G) {
        const int s0 = (u / H) * gs;
        n += (v6 ? 1 : 0) + min(gs, nslots - s0);
    }
    return n;
}

constexpr int MEGA_MAX_C = 4096;                                // LN row buffer of this path
struct MegaShared {
    WkvShared wkv;
    float red[32];
    union {
        float row[MEGA_MAX_C];                        // LN / embed phases: the token's residual row
        struct {
            float w_s[MEGA_MAX_TOK * WKV_N];          // WKV phases (v6): decays of the current unit [token][64]
            __half d_s[MEGA_MAX_TOK * MEGA_MAX_DD];   // tanh(Wd1 xw) rows of the unit's tokens
        };
    };
    int tok_s[MEGA_MAX_TOK];
    int s_last;
    int s_started;
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
    for (int u = cta; u < units; u += G) {
        const int h = u % w.H;
        const int s0 = (u / w.H) * gs;
        const int ns = min(gs, nslots - s0);
        if (VER == 6) {
            // ---- decay LoRA stage 2 for every token of this unit ----
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
            const int Dd = w.Dd;
            for (int i = tid; i < ntok * (Dd >> 1); i += CONSUMER_THREADS) {
                const int lt = i / (Dd >> 1), k = (i - lt * (Dd >> 1)) * 2;
                const uint32_t v2 = *reinterpret_cast<const uint32_t*>(w.d1 + a16_index(sh.tok_s[lt], k, w.d1_kq));
                *reinterpret_cast<uint32_t*>(&sh.d_s[lt * MEGA_MAX_DD + k]) = v2;
            }
            mbar_wait(full_bar + rp.stage * 8, rp.phase);
            cta_sync<true>();
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
        }
        int lt0 = 0;
        for (int sl = 0; sl < ns; ++sl) {
            const int si = s0 + sl;
            const int slot = mv.slot_id()[si];
            const int t0 = mv.slot_start()[si], nt = mv.slot_count()[si];
            mbar_wait(full_bar + rp.stage * 8, rp.phase);
            const uint32_t st = smem_base + rp.stage * MEGA_STAGE_BYTES;
            float4 m[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint4 raw = lds128(st + ((ig * 4 + e) * WKV_N + j4 * 4) * 4);
                m[e] = make_float4(__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z), __uint_as_float(raw.w));
            }
            ring_release(empty_bar, rp);            // the patch is in registers: free the slot early
            wkv_slot<VER, true>(w, h, t0, nt, m, sh.wkv, VER == 6 ? sh.w_s : nullptr, lt0);
            float* M = w.state + ((size_t)slot * w.H + h) * (WKV_N * WKV_N);
#pragma unroll
            for (int e = 0; e < 4; ++e) __stcs(reinterpret_cast<float4*>(M + (ig * 4 + e) * WKV_N + j4 * 4), m[e]);
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
    }
    if (warp == MEGA_MMA_WARP) tc_alloc(tmem_slot, MEGA_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

    const int nslots = mp.meta.nslots();
    const int gs = max(1, min(MEGA_MAX_GROUP, mp.meta.base[3]));

    if (warp == MEGA_PRODUCER_WARP) {
        if (lane == 0) mega_producer(mp, cta, G, smem_base, full_bar, empty_bar, &sh.s_started);
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
                    if (b0 < b1)
                        gemm_mma_role<1, MEGA_NSTAGE, MEGA_STAGE_BYTES>(g.p, b0, b1, smem_base, full_bar, empty_bar, tfull_bar,
                                                                        tempty_bar, tmem_base, rp, segcount);
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
                    if (cta < T) embed_row<true>(*mp.embed, cta, sh.row, sh.red);
                    break;
                case PH_LN:
                    if (cta < T) ln_mix_row<true>(mp.ln[P.idx], cta, sh.row, sh.red);
                    break;
                case PH_LNOUT:
                    if (cta < T) ln_out_row<true>(*mp.lnout, cta, sh.row, sh.red);
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
                unsigned long long t;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
                mp.trace[((size_t)tsel * mp.nphase + ph) * 4] = t;
            }
            if (ph + 1 < mp.nphase) mega_grid_barrier(mp.gbar, G, &sh.s_started, ph + 1);
            if (mp.trace && tsel >= 0 && tid == 0) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
                mp.trace[((size_t)tsel * mp.nphase + ph) * 4 + 1] = t;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MEGA_MMA_WARP) tc_dealloc(tmem_base, MEGA_TMEM_COLS);
}

}  // namespace b200
