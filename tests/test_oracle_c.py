"""The C/OpenMP restatement (oracle/rwkv_ref.c, the cpu_baseline / reference arm) agrees with the
NumPy oracle and the committed goldens.  Two f16-contract implementations differ by f16 rounding
flips of individual operands (~1e-4 of the logits range), well inside the 1e-3 budget."""
import os

import numpy as np
import pytest

from ai00_server_b200 import synth
from oracle import ref_c
from oracle import rwkv_numpy as O

if not os.path.exists(ref_c.LIB_PATH):
    from ai00_server_b200 import build
    build.build_oracle()


@pytest.mark.parametrize("preset,act", [("tiny6", "f16"), ("tiny6", "f32"), ("tiny5", "f16"), ("small6", "f16"), ("tiny7", "f16"),
                                        ("tiny7", "f32")])
def test_c_oracle_matches_numpy_oracle(preset, act):
    w = O.parse_st(synth.make_st(preset, 0))
    rc, orc = ref_c.RefC(w, act), O.Oracle(w, act)
    B, T = 3, 4
    rng = np.random.default_rng(2)
    toks = rng.integers(0, orc.info.num_vocab, size=(T, B))
    st = rc.state_init(B)
    sts = [orc.state_init() for _ in range(B)]
    tol = 1e-3 if act == "f16" else 2e-5
    for t in range(T):
        lg = rc.decode_step(toks[t], st)
        for b in range(B):
            want, sts[b] = orc.run([int(toks[t, b])], sts[b])
            assert np.abs(lg[b] - want[0]).max() <= tol * np.abs(want).max()
            assert lg[b].argmax() == want.argmax()
    for b in range(B):
        assert np.abs(st[b] - sts[b]).max() <= 10 * tol * max(1.0, np.abs(sts[b]).max())


@pytest.mark.parametrize("preset", ["tiny5", "tiny6", "tiny7"])
def test_c_oracle_matches_goldens(golden_dir, preset):
    g = np.load(os.path.join(golden_dir, f"model_{preset}.npz"))
    w = O.parse_st(synth.make_st(preset, 0))
    rc = ref_c.RefC(w, "f16")
    st = rc.state_init(1)
    rows = [rc.decode_step([int(t)], st)[0] for t in g["tokens"]]
    got = np.stack(rows)
    assert np.abs(got - g["logits_f16"]).max() <= 1e-3 * np.abs(g["logits_f16"]).max()
    assert (got.argmax(1) == g["logits_f16"].argmax(1)).all()


@pytest.mark.parametrize("preset,fixture", [("tiny6", "model6_fla.npz"), ("tiny7", "model7_fla.npz")])
def test_c_oracle_matches_fla_causal_lm_fixture(golden_dir, preset, fixture):
    """The C restatement (the CPU arm the bench times and the checker of the full-size GPU tests) against the logits of
    flash-linear-attention's whole models on the same weights (oracle/make_golden_fla_layers.py)."""
    g = np.load(os.path.join(golden_dir, fixture))
    w = O.parse_st(synth.make_st(preset, 0))
    rc = ref_c.RefC(w, "f32")
    st = rc.state_init(1)
    got = np.stack([rc.decode_step([int(t)], st)[0] for t in g["tokens"]])
    assert np.abs(got - g["logits"]).max() <= 2e-5 * np.abs(g["logits"]).max()
    assert (got.argmax(1) == g["logits"].argmax(1)).all()
