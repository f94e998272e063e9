"""CPU restatement of the reference's per-token sampling path.  TEST INFRASTRUCTURE ONLY (see oracle/rwkv_numpy.py):
only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this; the product path never does.

Follows, line by line:
  crates/ai00-core/src/run.rs:664-697          `sample()`: copy the logits row, Sampler::transform, Formatter::transform,
                                               bias add, softmax, Sampler::sample
  crates/ai00-core/src/sampler/nucleus.rs:50-59   NucleusSampler::init (penalties from the prompt)
  crates/ai00-core/src/sampler/nucleus.rs:61-67   transform: output[token] -= penalty
  crates/ai00-core/src/sampler/nucleus.rs:69-123  sample: sort the WHOLE vocabulary by probability, rev, take(top_k), top_p scan,
                                               pow(1/temperature), renormalise, cumulative, draw, penalty update
  crates/ai00-core/src/sampler/bnf.rs:37-40    Formatter::transform = kbnf mask_logits: disallowed tokens -> -inf

All arithmetic in f32 like the Rust code.  Ties: the reference sorts with voracious_sort (unstable), so the order of equal
probabilities is unspecified there; this restatement orders by (adjusted logit descending, token id ascending), which is one
of the reference's possible outcomes (the probability is a non-decreasing function of the logit) and is what the GPU front
half (csrc/sample.cuh) defines.  Pinning: the reference has no tests or vectors for this path, so this file is
pinned only by construction -- parity unpinned, like the model oracle.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def softmax_row(x: np.ndarray) -> np.ndarray:
    """exp(x - max) / sum, f32 (web-rwkv `softmax`, reference run.rs:1179)."""
    x = np.asarray(x, f32)
    m = x.max()
    if not np.isfinite(m):
        return np.zeros_like(x)
    e = np.exp(x - m, dtype=f32)
    return (e / e.sum(dtype=f32)).astype(f32)


def adjusted_logits(logits, penalties: dict | None, allow, bias: dict | None) -> np.ndarray:
    """run.rs:671-682: sampler.transform, formatter.transform, bias."""
    data = np.array(logits, f32, copy=True)
    for t, p in (penalties or {}).items():
        data[int(t)] = f32(data[int(t)] - f32(p))
    if allow is not None:
        data[~np.asarray(allow, bool)] = -np.inf
    for t, b in (bias or {}).items():
        data[int(t)] = f32(data[int(t)] + f32(b))
    return data


def sorted_candidates(logits, penalties=None, allow=None, bias=None, top_k: int = 128):
    """(ids, probs) of the `top_k` best tokens, best first: what nucleus.rs:70-80 keeps of its full sort."""
    data = adjusted_logits(logits, penalties, allow, bias)
    probs = softmax_row(data)
    order = np.lexsort((np.arange(data.size), -data.astype(np.float64)))      # logit descending, id ascending
    ids = order[:top_k].astype(np.uint32)
    return ids, probs[ids]


class NucleusSampler:
    """nucleus.rs:13-123 over the full vocabulary, state included."""

    def __init__(self, top_p=0.5, top_k=128, temperature=1.0, presence_penalty=0.3, frequency_penalty=0.3,
                 penalty_decay=0.99654026):
        self.top_p, self.top_k, self.temperature = f32(top_p), int(top_k), f32(temperature)
        self.presence_penalty, self.frequency_penalty, self.penalty_decay = f32(presence_penalty), f32(frequency_penalty), f32(penalty_decay)
        self.penalties: dict[int, np.float32] = {}

    def init(self, model_tokens):
        for index, token in enumerate(reversed(list(model_tokens))):
            pen = self.penalties.pop(int(token), self.presence_penalty)
            pen = f32(pen + self.frequency_penalty * f32(np.power(self.penalty_decay, f32(index))))
            self.penalties[int(token)] = pen

    def transform(self, output: np.ndarray) -> None:
        for t, p in self.penalties.items():
            output[t] = f32(output[t] - p)

    def sample(self, probs: np.ndarray, rand: float, order_key: np.ndarray | None = None) -> int:
        """`order_key`: what to sort by (the adjusted logits; defaults to the probabilities themselves)."""
        key = np.asarray(probs if order_key is None else order_key, np.float64)
        order = np.lexsort((np.arange(key.size), -key))[: self.top_k]
        kept, cum = [], f32(0.0)
        for i in order:
            if cum > self.top_p:
                break
            x = f32(probs[i])
            cum = f32(cum + x)
            kept.append((int(i), f32(np.power(x, f32(1.0) / self.temperature))))
        total = f32(0.0)
        for _, x in kept:
            total = f32(total + x)
        token, cum = kept[0][0], f32(0.0)
        for i, x in kept:
            cum = f32(cum + f32(x / total))
            if f32(rand) <= cum:
                token = i
                break
        for t in self.penalties:
            self.penalties[t] = f32(self.penalties[t] * self.penalty_decay)
        self.penalties[token] = f32(self.penalties[token] + self.frequency_penalty) if token in self.penalties else self.presence_penalty
        return token


def sample_token(logits, sampler: NucleusSampler, allow=None, bias=None, rand: float = 0.5) -> int:
    """run.rs:664-697 for one slot."""
    data = adjusted_logits(logits, sampler.penalties, allow, bias)
    return sampler.sample(softmax_row(data), rand, order_key=data)
