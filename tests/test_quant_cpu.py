"""CPU tests of the quantisation oracle (oracle/quant_numpy.py): internal consistency, the error bounds of the two formats, and
the one quantitative anchor the reference repository offers for them -- the documented VRAM of quantised models."""
import numpy as np

from ai00_server_b200 import synth
from oracle import quant_numpy as Q


def _w(n=96, k=512, seed=0, scale=0.05):
    return (np.random.default_rng(seed).standard_normal((n, k)) * scale).astype(np.float16)


def test_int8_codes_and_reconstruction_bound():
    w = _w()
    q, mn, mx = Q.quant_int8(w)
    assert q.dtype == np.uint8 and q.shape == w.shape and mn.shape == (96, 4)
    b = w.astype(np.float32).reshape(96, 4, 128)
    assert (mn.astype(np.float32) == b.min(2)).all() and (mx.astype(np.float32) == b.max(2)).all()
    # every block uses the whole code range
    qb = q.reshape(96, 4, 128)
    assert (qb.min(2) == 0).all() and (qb.max(2) == 255).all()
    step = (mx.astype(np.float32) - mn.astype(np.float32)) / 255
    for contract in ("f32", "engine"):
        d = np.abs(Q.dequant_int8(q, mn, mx, contract).astype(np.float32).reshape(96, 4, 128) - b)
        # half a step, plus the f16 rounding of scale and result in the engine contract
        slack = 0.0 if contract == "f32" else 1.5e-3 * np.abs(b).max()
        assert (d <= 0.5 * step[..., None] * 1.001 + 1e-7 + slack).all()
    # the two contracts differ by f16 rounding only
    e = Q.dequant_int8(q, mn, mx, "engine").astype(np.float32)
    f = Q.dequant_int8(q, mn, mx, "f32")
    assert np.abs(e - f).max() <= 2.0 ** -10 * np.abs(f).max()


def test_int8_degenerate_blocks():
    w = np.zeros((2, 128), np.float16)
    w[1, :] = np.float16(0.25)
    q, mn, mx = Q.quant_int8(w)
    assert (q == 0).all()
    assert (Q.dequant_int8(q, mn, mx).astype(np.float32) == w.astype(np.float32)).all()


def test_nf4_codes_are_nearest_levels():
    w = _w(seed=3)
    q, am = Q.quant_nf4(w)
    assert q.max() <= 15 and am.shape == (96, 8)
    b = w.astype(np.float32).reshape(96, 8, 64)
    assert (am.astype(np.float32) == np.abs(b).max(2)).all()
    x = b / am.astype(np.float32)[..., None]
    chosen = np.abs(x - Q.NF4_LEVELS[q.reshape(96, 8, 64)])
    best = np.abs(x[..., None] - Q.NF4_LEVELS).min(-1)
    assert (chosen == best).all()
    # identical to the first-minimum scan over the sixteen distances, also on the decision boundaries
    t = w.copy()
    t[0, :15] = ((Q.NF4_LEVELS[:-1] + Q.NF4_LEVELS[1:]) / 2).astype(np.float16)
    t[0, 15] = 1
    t[1, :16] = Q.NF4_LEVELS.astype(np.float16)
    qt, amt = Q.quant_nf4(t)
    xt = t.astype(np.float32).reshape(96, 8, 64) / amt.astype(np.float32)[..., None]
    assert (qt.reshape(96, 8, 64) == np.abs(xt[..., None] - Q.NF4_LEVELS).astype(np.float32).argmin(-1)).all()
    # the element that defines absmax is reproduced exactly (levels -1 and +1)
    rec = Q.dequant_nf4(q, am, "f32").reshape(96, 8, 64)
    idx = np.abs(b).argmax(2)
    assert np.allclose(np.take_along_axis(rec, idx[..., None], 2), np.take_along_axis(b, idx[..., None], 2))
    # zero block -> level 7 (0.0)
    z = np.zeros((1, 64), np.float16)
    qz, amz = Q.quant_nf4(z)
    assert (qz == 7).all() and (Q.dequant_nf4(qz, amz) == 0).all()
    # packing: element i of a group of eight in bits [4i, 4i+4)
    pk = Q.pack_nf4(q)
    assert pk.shape == (96, 64) and int(pk[0, 0]) & 15 == int(q[0, 0]) and (int(pk[0, 0]) >> 28) == int(q[0, 7])


def test_nf4_level_table_is_the_normalfloat_quantile_table():
    lv = Q.NF4_LEVELS
    assert lv.shape == (16,) and lv[0] == -1 and lv[7] == 0 and lv[15] == 1 and (np.diff(lv) > 0).all()
    # asymmetric by construction: 8 levels on the positive side, 7 on the negative
    assert (lv > 0).sum() == 8 and (lv < 0).sum() == 7
    # the published construction (QLoRA, Dettmers et al. 2023, App. E / bitsandbytes `create_normal_map(offset=0.9677083)`):
    # 8 quantiles of N(0, 1) on the positive side, 7 on the negative, an exact zero, normalised to [-1, 1]
    from scipy.stats import norm
    offset = 0.9677083
    pos = norm.ppf(np.linspace(offset, 0.5, 9)[:-1])
    neg = -norm.ppf(np.linspace(offset, 0.5, 8)[:-1])
    v = np.sort(np.concatenate([pos, [0.0], neg]))
    v /= v.max()
    assert np.abs(v - lv).max() < 1e-6


def test_quantize_model_touches_only_projection_matrices_of_the_first_layers():
    from oracle import rwkv_numpy as O
    w = O.parse_st(synth.make_st(synth.PRESETS["tiny6"], 0))
    wq = Q.quantize_model(w, 1, Q.QUANT_INT8)
    changed = sorted(k for k in w if not np.array_equal(np.asarray(w[k], np.float32), np.asarray(wq[k], np.float32)))
    assert changed == sorted(f"blocks.0.{m}" for m in Q.QUANT_MATRICES)
    assert Q.quantize_model(w, 2, Q.QUANT_NONE)["blocks.0.att.key.weight"] is w["blocks.0.att.key.weight"]


def test_bytes_per_weight_reproduce_the_documented_vram():
    """docs/doc-guide/quick-start.md:16-32 of the reference: RWKV-6 7B takes 14.4 GB in fp16, 8.2 GB in Int8, 5.2 GB in NF4.  With
    every layer quantised and embeddings + head in f16 the formats restated here give those footprints (the documented figures are
    measured VRAM, activations and state included, hence the tolerance; the smaller models' rows of that table do not scale the
    same way -- 3B: 6.5 / 4.4 / 2.6 -- so only the 7B row is used as an anchor)."""
    s = synth.PRESETS["v6-7b"]
    C, F, V, L = s.C, s.F, s.V, s.L
    mats = [(C, C)] * 5 + [(F, C), (C, F), (C, C)]
    fixed = 2 * V * C * 2
    for qt, want in zip((Q.QUANT_NONE, Q.QUANT_INT8, Q.QUANT_NF4), (14.4, 8.2, 5.2)):
        total = fixed + L * sum(Q.quant_weight_bytes(n, k, qt) for n, k in mats)
        assert abs(total / 1e9 - want) / want < 0.12, (qt, total / 1e9, want)
