"""CPU oracle: NumPy restatement of web-rwkv's weight-only quantisation (Int8 and NF4).

TEST INFRASTRUCTURE ONLY (same rule as rwkv_numpy.py: tests/, __graft_entry__.smoke() and bench.py's CPU legs only).

PARITY UNPINNED.  The reference selects these formats with `quant` / `quant_type` (reference crates/ai00-core/src/lib.rs:211-215,
465, 484, 694-704; reload.rs:23-27: the first `quant` layers are quantised, `Quant::{None, Int8, NF4, SF4}`), the arithmetic
lives in the un-vendored crate `web-rwkv = 0.10.18` (`tensor/ops.rs` + `shaders/quant_mat_int8.wgsl`, `quant_mat_nf4.wgsl`,
`matmul_vec_int8.wgsl`, `matmul_vec_nf4.wgsl`), and the reference ships no vectors for it.  Restated from the published
algorithm [UPSTREAM-RECALL]:

* Int8: the row-major [out, in] f16 matrix is cut into blocks of 128 consecutive elements (`INT8_BLOCK_SIZE`); a block keeps
  (min, max) as two f16 and every element as `pack4x8unorm((w - min) / (max - min))`, i.e. round(clamp(x, 0, 1) * 255);
  the product uses  w' = min + (q / 255) * (max - min).
* NF4: blocks of 64 consecutive elements (`NF4_BLOCK_SIZE`) keep absmax as f16; every element is the index of the NF4 level
  (QLoRA's 16 quantiles of N(0,1), table below) nearest to w / absmax, eight 4-bit indices per u32, element i in bits
  [4i, 4i+4); the product uses  w' = level[q] * absmax.
* Which matrices: the eight projection matrices of a quantised layer (att receptance / key / value / gate / output, ffn key /
  value / receptance); embeddings, head, LoRA / adapter matrices and every vector stay f16.

The engine contract (what `dequant_*(..., contract="engine")` returns and the GPU must reproduce bit for bit): the tensor
cores take f16 operands, so the dequantised weight is rounded to f16 -- Int8: f16(q * s + min) with s = f16((max - min) / 255)
as ONE fused multiply-add (HFMA2), NF4: f16(f16(level[q]) * absmax).  `contract="f32"` is the reference's f32 arithmetic;
tests bound the distance between the two (below the f16 rounding of the weight itself).
"""
from __future__ import annotations

import numpy as np

INT8_BLOCK = 128
NF4_BLOCK = 64
QUANT_NONE, QUANT_INT8, QUANT_NF4 = 0, 1, 2

# bitsandbytes / QLoRA NormalFloat4 levels (normalised N(0,1) quantiles; tests/test_quant_cpu.py re-derives them from the published construction)
NF4_LEVELS = np.array([
    -1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635, -0.18477343022823334,
    -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725, 0.24611230194568634, 0.33791524171829224,
    0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0], dtype=np.float32)

QUANT_MATRICES = ("att.receptance.weight", "att.key.weight", "att.value.weight", "att.gate.weight", "att.output.weight",
                  "ffn.key.weight", "ffn.value.weight", "ffn.receptance.weight")


def quant_int8(w16: np.ndarray):
    """[out, in] f16 -> (codes u8 [out, in], min f16 [out, in/128], max f16 [out, in/128])."""
    w16 = np.asarray(w16, np.float16)
    n, k = w16.shape
    assert k % INT8_BLOCK == 0, "Int8 blocks are 128 consecutive input elements"
    b = w16.astype(np.float32).reshape(n, k // INT8_BLOCK, INT8_BLOCK)
    mn, mx = b.min(axis=2), b.max(axis=2)
    rng = (mx - mn).astype(np.float32)
    safe = np.where(rng > 0, rng, np.float32(1))
    x = ((b - mn[..., None]).astype(np.float32) / safe[..., None]).astype(np.float32)
    x = np.clip(x, np.float32(0), np.float32(1))
    q = np.floor((x * np.float32(255)).astype(np.float32) + np.float32(0.5)).astype(np.uint8)
    q[rng <= 0] = 0
    return q.reshape(n, k), mn.astype(np.float16), mx.astype(np.float16)


def int8_scale(mn16: np.ndarray, mx16: np.ndarray) -> np.ndarray:
    """s = f16((max - min) / 255), evaluated in f32."""
    return ((mx16.astype(np.float32) - mn16.astype(np.float32)) / np.float32(255)).astype(np.float16)


def dequant_int8(q: np.ndarray, mn16: np.ndarray, mx16: np.ndarray, contract: str = "engine") -> np.ndarray:
    n, k = q.shape
    qb = q.reshape(n, k // INT8_BLOCK, INT8_BLOCK)
    if contract == "engine":
        s = int8_scale(mn16, mx16).astype(np.float64)[..., None]
        w = (qb.astype(np.float64) * s + mn16.astype(np.float64)[..., None]).astype(np.float16)   # exact product, one rounding
        return w.reshape(n, k)
    mn, mx = mn16.astype(np.float32)[..., None], mx16.astype(np.float32)[..., None]
    return (mn + (qb.astype(np.float32) / np.float32(255)) * (mx - mn)).astype(np.float32).reshape(n, k)


def quant_nf4(w16: np.ndarray):
    """[out, in] f16 -> (codes u8 in 0..15 [out, in], absmax f16 [out, in/64])."""
    w16 = np.asarray(w16, np.float16)
    n, k = w16.shape
    assert k % NF4_BLOCK == 0, "NF4 blocks are 64 consecutive input elements"
    b = w16.astype(np.float32).reshape(n, k // NF4_BLOCK, NF4_BLOCK)
    am = np.abs(b).max(axis=2)
    safe = np.where(am > 0, am, np.float32(1))
    x = (b / safe[..., None]).astype(np.float32)
    # nearest level, the lower one on a tie (what a first-minimum scan over |x - level[i]|, i = 0..15, in f32 returns): locate x
    # between the level midpoints, then settle against both neighbours with the f32 distances themselves
    mids = ((NF4_LEVELS[:-1].astype(np.float64) + NF4_LEVELS[1:].astype(np.float64)) / 2).astype(np.float32)
    q = np.searchsorted(mids, x.reshape(-1)).reshape(x.shape).astype(np.int8)
    dist = lambda c: np.abs(x - NF4_LEVELS[c])
    lo = np.maximum(q - 1, 0)
    q = np.where(dist(lo) <= dist(q), lo, q)
    hi = np.minimum(q + 1, 15)
    q = np.where(dist(hi) < dist(q), hi, q).astype(np.uint8)
    q[am <= 0] = 7                                   # level 0.0
    return q.reshape(n, k), am.astype(np.float16)


def dequant_nf4(q: np.ndarray, am16: np.ndarray, contract: str = "engine") -> np.ndarray:
    n, k = q.shape
    qb = q.reshape(n, k // NF4_BLOCK, NF4_BLOCK)
    if contract == "engine":
        lv = NF4_LEVELS.astype(np.float16).astype(np.float64)[qb]
        return (lv * am16.astype(np.float64)[..., None]).astype(np.float16).reshape(n, k)
    return (NF4_LEVELS[qb] * am16.astype(np.float32)[..., None]).astype(np.float32).reshape(n, k)


def pack_nf4(q: np.ndarray) -> np.ndarray:
    """codes [out, in] -> u32 [out, in/8], element i of a group of eight in bits [4i, 4i+4)."""
    n, k = q.shape
    g = q.reshape(n, k // 8, 8).astype(np.uint32)
    out = np.zeros((n, k // 8), np.uint32)
    for i in range(8):
        out |= g[:, :, i] << np.uint32(4 * i)
    return out


def quantize_model(weights: dict[str, np.ndarray], layers: int, qtype: int, contract: str = "engine") -> dict[str, np.ndarray]:
    """The weights the forward pass of a model loaded with `quant = layers`, `quant_type = qtype` multiplies with
    (reference lib.rs:465: `(0..quant).map(|layer| (layer, quant_type))`)."""
    out = dict(weights)
    if qtype == QUANT_NONE:
        return out
    for l in range(layers):
        for m in QUANT_MATRICES:
            name = f"blocks.{l}.{m}"
            if name not in weights:
                continue                             # v7 has no gate / ffn.receptance matrices
            if qtype == QUANT_INT8:
                out[name] = dequant_int8(*quant_int8(weights[name]), contract=contract)
            elif qtype == QUANT_NF4:
                out[name] = dequant_nf4(*quant_nf4(weights[name]), contract=contract)
            else:
                raise ValueError("unsupported quant type (SF4 is not restated)")
    return out


def quant_weight_bytes(n: int, k: int, qtype: int) -> int:
    """Bytes one pass over an [n, k] matrix streams."""
    if qtype == QUANT_INT8:
        return n * k + (n * k // INT8_BLOCK) * 4
    if qtype == QUANT_NF4:
        return n * k // 2 + (n * k // NF4_BLOCK) * 2
    return n * k * 2
