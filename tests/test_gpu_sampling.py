"""GPU front half of sampling (b200rwkv_sample_topk, csrc/sample.cuh) through the C ABI against the CPU restatement of
run.rs:664-697 + sampler/nucleus.rs (oracle/sampling_numpy.py): identical candidate ids (bit-exact, ties by token id),
probabilities to f32 rounding, identical sampled tokens over a generation loop, per-slot rows surviving later steps of other
slots, and the error behaviour."""
import numpy as np
import pytest

from ai00_server_b200 import capi, runtime, synth
from oracle import sampling_numpy as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    st = synth.make_st("small6", 0)            # V = 2048: one segment
    m = runtime.Model(st, max_batch=4, token_chunk_size=32)
    yield m
    m.close()


@pytest.fixture(scope="module")
def eng_wide():
    import dataclasses
    st = synth.make_st(dataclasses.replace(synth.PRESETS["tiny6"], V=65536), 0)     # the World vocabulary: 32 segments
    m = runtime.Model(st, max_batch=4, token_chunk_size=32)
    yield m
    m.close()


def logits_of(m, slots, toks):
    """host copy of the rows (separate call on a restored state would change nothing: same step, same bits)"""
    return m.infer_raw(slots, [1] * len(slots), toks, [capi.OPTION_LAST] * len(slots))


@pytest.mark.parametrize("which", ["eng", "eng_wide"])
def test_topk_matches_the_full_vocabulary_sort(which, request):
    m = request.getfixturevalue(which)
    V = m.info["num_vocab"]
    rng = np.random.default_rng(1)
    slots = [0, 1, 2]
    for s in slots:
        m.state.load(m.state.init(), s)
    rows = logits_of(m, slots, [5, 6, 7])
    rows = [r[0].copy() for r in rows]
    pen = [{int(t): float(v) for t, v in zip(rng.integers(0, V, 50), rng.random(50))}, {}, {3: 0.3}]
    bias = [{}, {int(rows[1].argmax()): -50.0, 11: 4.0}, {}]
    allow = rng.random((3, V)) < 0.5
    allow[0] = True
    for top_k in (1, 7, 128):
        ids, probs = m.sample_topk(slots, penalties=pen, bias=bias, allow=allow, top_k=top_k)
        for i in range(3):
            wi, wp = S.sorted_candidates(rows[i], pen[i], allow[i], bias[i], top_k=top_k)
            assert np.array_equal(ids[i], wi), (which, top_k, i)
            np.testing.assert_allclose(probs[i], wp, rtol=2e-5, atol=1e-12)
    ids, probs = m.sample_topk(slots, top_k=1)                     # greedy, no adjustments
    assert ids[:, 0].tolist() == [int(r.argmax()) for r in rows]


def test_rows_belong_to_slots_not_to_steps(eng):
    """Slot 0's row must survive later steps that only contain other slots (the reference samples every slot in its own
    task while the infer loop goes on)."""
    m = eng
    for s in range(3):
        m.state.load(m.state.init(), s)
    r0 = m.infer_raw([0, 1], [2, 1], [4, 9, 8], [capi.OPTION_LAST] * 2)[0][0].copy()
    m.infer_raw([1, 2], [1, 3], [3, 5, 6, 7], [capi.OPTION_LAST] * 2, keep_on_device=True)
    ids, _ = m.sample_topk([0], top_k=5)
    wi, _ = S.sorted_candidates(r0, top_k=5)
    assert np.array_equal(ids[0], wi)


def test_generation_loop_equals_the_reference_flow(eng):
    """Decode 24 tokens on two slots: (a) logits to the host + full-vocabulary CPU sampler (what the reference does),
    (b) logits kept in HBM + GPU front half + candidate sampler.  Same random draws -> the same tokens."""
    m = eng
    rng = np.random.default_rng(2)
    draws = rng.random((24, 2))
    prompt = [[3, 4, 5], [9]]
    outs = []
    for mode in ("host", "gpu"):
        for s in range(2):
            m.state.load(m.state.init(), s)
        full = [S.NucleusSampler(top_p=0.8, temperature=1.1) for _ in range(2)]
        cand = [runtime.NucleusSampler(top_p=0.8, temperature=1.1) for _ in range(2)]
        for s in range(2):
            full[s].init(prompt[s]); cand[s].init(prompt[s])
        toks = [p[:] for p in prompt]
        feed = [p[:] for p in prompt]
        for step in range(24):
            if mode == "host":
                rows = m.infer_raw([0, 1], [len(f) for f in feed], sum(feed, []), [capi.OPTION_LAST] * 2)
                nxt = [S.sample_token(rows[s][0], full[s], rand=draws[step, s]) for s in range(2)]
            else:
                m.infer_raw([0, 1], [len(f) for f in feed], sum(feed, []), [capi.OPTION_LAST] * 2, keep_on_device=True)
                ids, probs = m.sample_topk([0, 1], penalties=[c.penalties for c in cand], top_k=128)
                nxt = [cand[s].sample_candidates(ids[s], probs[s], draws[step, s]) for s in range(2)]
            for s in range(2):
                toks[s].append(nxt[s])
            feed = [[t] for t in nxt]
        outs.append(toks)
    assert outs[0] == outs[1]


def test_sample_topk_errors(eng):
    m = eng
    m.state.load(m.state.init(), 3)
    with pytest.raises(capi.B200Error) as ei:
        m.sample_topk([99], top_k=4)
    assert ei.value.code == capi.ERR_STATE
    with pytest.raises(capi.B200Error) as ei:
        m.sample_topk([3], top_k=4)                  # slot 3 never produced a row
    assert ei.value.code == capi.ERR_STATE
    with pytest.raises(capi.B200Error) as ei:
        m.sample_topk([0], top_k=129)
    assert ei.value.code == capi.ERR_INVALID
    with pytest.raises(capi.B200Error) as ei:
        m.sample_topk([0, 0], top_k=4)
    assert ei.value.code == capi.ERR_INVALID
