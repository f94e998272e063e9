/* OpenMP fill for the synthetic `.st` image (bench/test tooling, not product code).
 * Bit-identical to ai00_server_b200/synth.py::fill_numpy: murmur3 finaliser over the
 * element index keyed by a per-tensor seed, mapped to uniform [lo, hi) in f32 (separate
 * multiply and add: build with -ffp-contract=off), rounded to f16 nearest-even. */
#include <stdint.h>

void synth_fill_f16(void* dst_, uint64_t n, uint32_t seed, float lo, float hi) {
    _Float16* dst = (_Float16*)dst_;
    const float span = hi - lo;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        uint32_t h = (uint32_t)(uint64_t)i * 0x9E3779B1u + seed;
        h ^= h >> 16;
        h *= 0x85EBCA6Bu;
        h ^= h >> 13;
        h *= 0xC2B2AE35u;
        h ^= h >> 16;
        float u = (float)(h >> 8) * (1.0f / 16777216.0f);
        float t = u * span;
        dst[i] = (_Float16)(t + lo);
    }
}
