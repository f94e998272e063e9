"""Generate the committed golden fixtures under tests/golden/.

TEST INFRASTRUCTURE.  Run here (CPU container) with `python -m oracle.make_golden`.

Two kinds of fixture:

1. `wkv6_fla.npz`, `wkv7_fla.npz` — INDEPENDENT pins: inputs and outputs of
   flash-linear-attention's pure-torch naive recurrences
   (`fla.ops.rwkv6.recurrent_naive.naive_recurrent_rwkv6`,
   `fla.ops.generalized_delta_rule.dplr.naive.dplr_recurrence`), which are installed in
   this image but cannot travel to the GPU box as a dependency of the tests.  The
   oracle's WKV recurrences must reproduce them (tests/test_oracle.py) and so must the
   CUDA kernels (tests/test_gpu_kernels.py).
2. `model_<preset>.npz` — regression goldens of the oracle itself on the synthetic
   tiny models (logits for a fixed token run, both activation contracts, and a state
   checksum).  They pin the oracle against accidental edits and give the C restatement
   (oracle/rwkv_ref.c) and the CUDA engine a fixture that does not need the oracle at
   run time.  They are NOT an independent check of the math — the reference holds no
   golden vectors for this path (parity unpinned).
"""
from __future__ import annotations

import os
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

GOLDEN_TOKENS = [1, 5, 9, 33, 2, 7, 300, 41, 41, 8, 0, 17]


def gen_wkv6(rng):
    import torch
    from fla.ops.rwkv6.recurrent_naive import naive_recurrent_rwkv6
    T, H, N = 6, 3, 64
    r = rng.standard_normal((T, H, N)).astype(np.float32)
    k = rng.standard_normal((T, H, N)).astype(np.float32)
    v = rng.standard_normal((T, H, N)).astype(np.float32)
    wraw = rng.uniform(-6, 0, (T, H, N)).astype(np.float32)          # w = exp(-exp(wraw))
    u = rng.standard_normal((H, N)).astype(np.float32) * 0.3
    S0 = rng.standard_normal((H, N, N)).astype(np.float32)
    tt = lambda a: torch.from_numpy(a).permute(1, 0, 2)[None].contiguous()   # [1,H,T,N]
    o, ht = naive_recurrent_rwkv6(tt(r), tt(k), tt(v), -torch.exp(tt(wraw)), torch.from_numpy(u),
                                  scale=1.0, initial_state=torch.from_numpy(S0)[None], output_final_state=True)
    np.savez_compressed(os.path.join(GOLD, "wkv6_fla.npz"), r=r, k=k, v=v, wraw=wraw, u=u, S0=S0,
                        out=o[0].permute(1, 0, 2).numpy(), S=ht[0].numpy())


def gen_wkv7(rng):
    import torch
    from fla.ops.generalized_delta_rule.dplr.naive import dplr_recurrence
    T, H, N = 6, 3, 64
    r = rng.standard_normal((T, H, N)).astype(np.float32)
    k = rng.standard_normal((T, H, N)).astype(np.float32)
    v = rng.standard_normal((T, H, N)).astype(np.float32)
    w = np.exp(-0.606531 / (1 + np.exp(-rng.standard_normal((T, H, N))))).astype(np.float32)
    kk = rng.standard_normal((T, H, N)).astype(np.float32)
    kk /= np.linalg.norm(kk, axis=-1, keepdims=True)
    a = (1 / (1 + np.exp(-rng.standard_normal((T, H, N))))).astype(np.float32)
    S0 = rng.standard_normal((H, N, N)).astype(np.float32)            # [H, value, key]
    tt = lambda x: torch.from_numpy(x).permute(1, 0, 2)[None].contiguous()
    # fla state is [K, V] = transpose of the official [V, K]; it scales q by N^-1/2.
    o, ht = dplr_recurrence(tt(r) * (N ** 0.5), tt(k), tt(v), tt(-kk), tt(kk * a), torch.log(tt(w)),
                            initial_state=torch.from_numpy(S0.transpose(0, 2, 1).copy())[None])
    np.savez_compressed(os.path.join(GOLD, "wkv7_fla.npz"), r=r, k=k, v=v, w=w, kk=kk, a=a, S0=S0,
                        out=o[0].permute(1, 0, 2).numpy(), S=ht[0].permute(0, 2, 1).numpy())


def gen_models():
    from ai00_server_b200 import synth
    from oracle import rwkv_numpy as O
    for preset in ("tiny5", "tiny6", "tiny7"):
        w = O.parse_st(synth.make_st(preset, seed=0, force_numpy=True))
        rec = {"tokens": np.asarray(GOLDEN_TOKENS, np.int64)}
        for act in ("f16", "f32"):
            orc = O.Oracle(w, act)
            logits, st = orc.run(GOLDEN_TOKENS, orc.state_init(), full=True)
            rec[f"logits_{act}"] = logits
            rec[f"state_sum_{act}"] = np.array([st.astype(np.float64).sum(), np.abs(st).astype(np.float64).sum()])
            if act == "f16":
                rec["state_f16"] = st
        np.savez_compressed(os.path.join(GOLD, f"model_{preset}.npz"), **rec)


def main():
    warnings.filterwarnings("ignore")
    os.makedirs(GOLD, exist_ok=True)
    rng = np.random.default_rng(20260923)
    gen_wkv6(rng)
    gen_wkv7(rng)
    gen_models()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
