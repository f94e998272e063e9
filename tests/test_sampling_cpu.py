"""Host back half of the sampler (runtime.NucleusSampler over <= 128 candidates) against the full-vocabulary restatement of the
reference's NucleusSampler (oracle/sampling_numpy.py; crates/ai00-core/src/sampler/nucleus.rs:50-123): same tokens, same
penalty state, for random logits, with and without penalties / bias / grammar masks.  No GPU: the candidates here come from the
oracle's own sort, so this pins the split of the algorithm, not the kernels (tests/test_gpu_sampling.py does that)."""
import numpy as np
import pytest

from ai00_server_b200 import runtime
from oracle import sampling_numpy as S


@pytest.mark.parametrize("params", [dict(), dict(top_p=0.9, top_k=40, temperature=1.3), dict(top_p=0.0, top_k=1),
                                    dict(top_p=1.0, top_k=128, temperature=0.7, presence_penalty=0.5, frequency_penalty=0.1)])
def test_candidate_sampler_equals_full_vocabulary_sampler(params):
    rng = np.random.default_rng(4)
    V = 4096
    full, cand = S.NucleusSampler(**params), runtime.NucleusSampler(**params)
    prompt = rng.integers(0, V, size=20).tolist()
    full.init(prompt); cand.init(prompt)
    assert full.penalties == cand.penalties
    bias = {5: 2.0, 77: -3.5}
    for step in range(40):
        logits = (rng.standard_normal(V) * 3).astype(np.float32)
        allow = None if step % 3 else rng.random(V) < 0.6
        rand = float(rng.random())
        want = S.sample_token(logits, full, allow=allow, bias=bias, rand=rand)
        ids, probs = S.sorted_candidates(logits, cand.penalties, allow, bias, top_k=128)
        got = cand.sample_candidates(ids, probs, rand)
        assert got == want, step
        assert full.penalties == cand.penalties


def test_greedy_is_the_argmax_of_the_adjusted_row():
    rng = np.random.default_rng(5)
    logits = rng.standard_normal(1000).astype(np.float32)
    s = S.NucleusSampler(top_k=1, top_p=0.0, presence_penalty=0.0, frequency_penalty=0.0)
    assert S.sample_token(logits, s, rand=0.999) == int(logits.argmax())
    ids, probs = S.sorted_candidates(logits, top_k=4)
    assert ids[0] == logits.argmax() and np.all(np.diff(probs) <= 0)
    assert abs(float(S.softmax_row(logits).sum()) - 1.0) < 1e-5


def test_ties_are_ordered_by_token_id():
    logits = np.zeros(64, np.float32)
    logits[[9, 3, 40]] = 5.0
    ids, probs = S.sorted_candidates(logits, top_k=5)
    assert ids.tolist()[:3] == [3, 9, 40] and ids.tolist()[3:] == [0, 1]
    assert probs[0] == probs[1] == probs[2]
