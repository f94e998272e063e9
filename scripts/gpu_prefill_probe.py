"""Per-launch windows of steps with 16 / 32 / 64 / 128 tokens (one token per slot): how the projection kernels scale with the
number of 16-token operand tiles (MT = 1, 2, 4, 8).  usage: gpu_prefill_probe.py [preset]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai00_server_b200 import runtime, synth

preset = sys.argv[1] if len(sys.argv) > 1 else "v6-3b"
st = synth.make_st(preset, 0)
m = runtime.Model(st, max_batch=128, token_chunk_size=128)
rng = np.random.default_rng(0)
for s in range(128):
    m.state.load(m.state.init(), s) if s == 0 else None
for n in (16, 32, 64, 128):
    slots = list(range(n))
    toks = rng.integers(1, 60000, n).astype(np.uint32)
    wins, step_us = m.profile_insitu(slots, toks, reps=3)
    cls = {}
    for w in wins:
        ty = w["type"]
        name = f"gemm_{ty - 1000000}MiB" if ty >= 1000000 else {0: "ln_mix", 2: "wkv", 6: "front_half"}.get(ty, str(ty))
        a = cls.setdefault(name, [0.0, 0])
        a[0] += w["end_us"] - w["start_us"]; a[1] += 1
    tot = sum(v[0] for v in cls.values())
    print(f"--- {preset}: {n} tokens per step: step {step_us:.0f} us, windows {tot:.0f} us")
    for k, v in sorted(cls.items()):
        print(f"    {k:14s} n={v[1]:3d} avg {v[0] / v[1]:8.2f} us  sum {v[0]:9.1f}")
m.close()
