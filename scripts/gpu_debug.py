"""Localise a parity failure: one-layer models, every intermediate buffer vs the oracle trace.
Run on the GPU box: python scripts/gpu_debug.py [preset ...]"""
import dataclasses
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai00_server_b200 import capi, runtime, synth  # noqa: E402
from oracle import rwkv_numpy as O  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run(preset, L=1, toks=(5, 9)):
    shp = dataclasses.replace(synth.PRESETS[preset], L=L)
    st = synth.make_st(shp, 0)
    m = runtime.Model(st, max_batch=2, token_chunk_size=16)
    orc = O.Oracle(O.parse_st(st), "f16")
    orc.trace = {}
    state = orc.state_init()
    state[:] = np.random.default_rng(1).standard_normal(state.shape).astype(np.float32) * 0.5
    m.state.load(state, 1)
    back = m.state.back(1)
    print(f"== {preset} L={L}: load/back identity:", np.array_equal(back, state))
    want, wst = orc.run(list(toks), state, full=True)
    got = m.infer_raw([1], [len(toks)], list(toks), [capi.OPTION_FULL])[0]
    T = len(toks)
    l = L - 1
    names6 = {"xx1": "xx1", "sx1": "sx1", "r": "r", "k": "k", "v": "v", "g": "g", "w": "w", "part_att": "part_att",
              "xx2": "xx2", "rr": "rr", "part_ffn": "part_ffn", "a_out": "wkv_out", "a_kk": "kk"}
    ver = m.info["version"]
    if ver == 6:
        names6.update({"a_x1": "xk", "a_x2": "xv", "a_x3": "xr", "a_x4": "xg"})
    if ver == 7:
        names6.pop("sx1"); names6.pop("rr")
        names6.update({"a": "a"})
    for buf, tr in names6.items():
        key = f"{l}.{tr}"
        if key not in orc.trace:
            continue
        try:
            d = m.debug_read(buf, rows=T)[T - 1]
        except capi.B200Error as e:
            print(f"   {buf:10s} <err {e}>")
            continue
        w_ = orc.trace[key]
        print(f"   {buf:10s} vs {tr:10s} rel={rel(d[:w_.size], w_):.3e}  |ref|max={np.abs(w_).max():.3f}")
    print("   logits rel", rel(got, want), "argmax", got.argmax(1), want.argmax(1))
    print("   state rel", rel(m.state.back(1), wst))
    m.close()


if __name__ == "__main__":
    presets = sys.argv[1:] or ["tiny6", "tiny5", "tiny7"]
    for p in presets:
        run(p, 1)
        run(p, 2, toks=(5, 9, 11))
