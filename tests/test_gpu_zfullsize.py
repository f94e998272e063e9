"""BASELINE.json's north-star shape (RWKV-6 7B, fp16) on the GPU: size-independent properties, plus the distance to the
C oracle next to the noise floor of the f16-operand contract at this depth (two correct implementations are
~4e-3 apart at 32 layers, so the 1e-3 bound of the small-model tests cannot be asked here).  Default per-op path."""
import os

import numpy as np
import pytest

from ai00_server_b200 import capi, runtime, synth
from oracle import ref_c
from oracle import rwkv_numpy as O

pytestmark = pytest.mark.gpu
PRESET = "v6-7b"

if not os.path.exists(ref_c.LIB_PATH):
    from ai00_server_b200 import build
    build.build_oracle()


@pytest.fixture(scope="module")
def big():
    st = synth.make_st(PRESET, 0)
    os.environ["B200RWKV_MEGA"] = "0"
    try:
        m = runtime.Model(st, max_batch=4, token_chunk_size=32)
    finally:
        os.environ.pop("B200RWKV_MEGA", None)
    yield m, st
    m.close()


def test_full_size_decode_is_deterministic_and_batch_invariant(big):
    m, _ = big
    rng = np.random.default_rng(17)
    zero = m.state.init()
    toks = rng.integers(1, 60000, size=(3, 4))

    def run(slots):
        for s in slots:
            m.state.load(zero, s)
        out = None
        for i in range(3):
            out = m.infer_raw(slots, [1] * len(slots), toks[i, slots].tolist(), [capi.OPTION_LAST] * len(slots))
        return {s: out[j].copy() for j, s in enumerate(slots)}

    a = run([0, 1, 2, 3])
    b = run([0, 1, 2, 3])
    for s in range(4):
        assert np.array_equal(a[s], b[s])                  # same inputs, same bits
        assert np.isfinite(a[s]).all()
    c = run([2])                                            # slot 2 alone: same step shape, other slots absent
    assert np.array_equal(c[2], a[2])


def test_full_size_state_roundtrip(big):
    m, _ = big
    rng = np.random.default_rng(18)
    m.state.load(m.state.init(), 1)
    m.infer_raw([1], [5], rng.integers(1, 60000, size=5).tolist(), [capi.OPTION_LAST])
    st = m.state.back(1)
    assert np.isfinite(st).all() and np.abs(st).max() > 0
    m.state.load(st, 3)
    assert np.array_equal(m.state.back(3), st)              # load o back = id
    t = int(rng.integers(1, 60000))
    x = m.infer_raw([1], [1], [t], [capi.OPTION_LAST])[0]
    y = m.infer_raw([3], [1], [t], [capi.OPTION_LAST])[0]
    assert np.array_equal(x, y)                             # a restored state continues identically


def test_full_size_distance_to_the_oracle(big):
    m, st = big
    w = O.parse_st(st)
    c16, c32 = ref_c.RefC(w, "f16"), ref_c.RefC(w, "f32")
    s16, s32 = c16.state_init(2), c32.state_init(2)
    for s in range(2):
        m.state.load(m.state.init(), s)
    rng = np.random.default_rng(19)
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    for i in range(2):
        toks = rng.integers(1, 60000, size=2)
        rows = np.concatenate(m.infer_raw([0, 1], [1, 1], toks.tolist(), [capi.OPTION_LAST] * 2))
        a, b = c16.decode_step(toks, s16), c32.decode_step(toks, s32)
        floor = rel(a, b)                                   # oracle vs oracle: what f16 operand rounding alone does at this depth
        print(f"step {i}: engine vs C f16-contract {rel(rows, a):.2e}, vs C f32-contract {rel(rows, b):.2e}, oracle-vs-oracle {floor:.2e}")
        assert rel(rows, a) <= max(3e-2, 5 * floor) and rel(rows, b) <= max(3e-2, 5 * floor)


def test_front_half_kernel_with_the_7b_lora_rank_matches_the_oracle():
    """The RWKV-6 front-half kernel is templated on the ddlerp LoRA rank: 32 (every CI preset) and 64 (only the 7B shape).
    A 4-layer model with rank 64 puts the 7B instantiation under the 1e-3 bound of the small-model tests."""
    import dataclasses
    shp = dataclasses.replace(synth.PRESETS["small6"], Dm=64, Dd=128)
    st = synth.make_st(shp, 0)
    os.environ["B200RWKV_MEGA"] = "0"
    try:
        m = runtime.Model(st, max_batch=4, token_chunk_size=32)
    finally:
        os.environ.pop("B200RWKV_MEGA", None)
    try:
        orc = O.Oracle(O.parse_st(st), "f16")
        rng = np.random.default_rng(23)
        counts = [1, 3, 1]
        sts = [orc.state_init() for _ in counts]
        for s in range(3):
            m.state.load(m.state.init(), s)
        for _ in range(4):
            toks = [rng.integers(1, 2000, size=n).tolist() for n in counts]
            rows = m.infer_raw([0, 1, 2], counts, sum(toks, []), [capi.OPTION_LAST] * 3)
            for s in range(3):
                want, sts[s] = orc.run(toks[s], sts[s])
                err = float(np.abs(rows[s] - want).max() / np.abs(want).max())
                assert err <= 1e-3 and rows[s].argmax() == want.argmax(), (s, err)
    finally:
        m.close()
