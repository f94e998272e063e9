"""Operator-level known-answer tests: the CUDA WKV kernels, driven alone through the C ABI (b200rwkv_op_wkv), against the
committed flash-linear-attention fixtures -- inputs and outputs of fla's pure-torch naive recurrences
(oracle/make_golden.py: naive_recurrent_rwkv6, dplr_recurrence), the one pin of this path that does not come from this
repository's own restatement.  The kernel fuses the per-head GroupNorm behind the recurrence, so the recurrence output is
compared after the same normalisation (eps 64e-5, weight 1, bias 0, gate 1) at the f16 resolution the kernel writes; the
state it leaves is compared in f32."""
import os

import numpy as np
import pytest

from ai00_server_b200 import capi

pytestmark = pytest.mark.gpu
GN_EPS = 64e-5


def group_norm(x):          # [T, H, N] over N
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + GN_EPS)


@pytest.mark.parametrize("split", [(6,), (3, 3), (1, 4, 1)])
def test_wkv6_kernel_reproduces_the_fla_fixture(golden_dir, split):
    g = np.load(os.path.join(golden_dir, "wkv6_fla.npz"))
    w = np.exp(-np.exp(g["wraw"].astype(np.float64))).astype(np.float32)
    M = np.ascontiguousarray(g["S0"].transpose(0, 2, 1))            # fixture S[key][value] -> device M[value][key]
    outs, t0 = [], 0
    for n in split:                                                  # one launch per run of tokens, state carried
        sl = slice(t0, t0 + n)
        o, M = capi.op_wkv(6, g["r"][sl], g["k"][sl], g["v"][sl], w[sl], M, u=g["u"])
        outs.append(o); t0 += n
    got = np.concatenate(outs)
    want = group_norm(g["out"])
    assert np.abs(got - want).max() <= 2e-3                          # f16 output of O(1) values
    assert np.abs(M.transpose(0, 2, 1) - g["S"]).max() / np.abs(g["S"]).max() <= 1e-5


def test_wkv5_kernel_with_static_decay(golden_dir):
    """RWKV-5 is the same recurrence with a per-channel static decay: the fixture's first-token decay for all tokens, checked
    against a float64 evaluation of the recurrence."""
    g = np.load(os.path.join(golden_dir, "wkv6_fla.npz"))
    w = np.exp(-np.exp(g["wraw"][0].astype(np.float64)))             # [H, N]
    r, k, v, u = (g[n].astype(np.float64) for n in ("r", "k", "v", "u"))
    S = g["S0"].astype(np.float64).copy()                            # [H, i key, j value]
    want = np.zeros_like(r)
    for t in range(r.shape[0]):
        kv = k[t][:, :, None] * v[t][:, None, :]
        want[t] = np.einsum("hi,hij->hj", r[t], u[:, :, None] * kv + S)
        S = kv + w[:, :, None] * S
    got, M = capi.op_wkv(5, g["r"], g["k"], g["v"], w.astype(np.float32), np.ascontiguousarray(g["S0"].transpose(0, 2, 1)), u=g["u"])
    assert np.abs(got - group_norm(want)).max() <= 2e-3
    assert np.abs(M.transpose(0, 2, 1) - S).max() / np.abs(S).max() <= 1e-5


def test_wkv7_kernel_reproduces_the_fla_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "wkv7_fla.npz"))
    T, H, N = g["r"].shape
    M = g["S0"].copy()                                               # fixture and device both [H, value, key]
    outs = []
    zeros = np.zeros((H, N), np.float32)
    for t in range(T):
        # the fixture supplies kk directly; the kernel derives it as normalize_head(k * k_k): choose k_k = kk / k for this token
        kk_over_k = (g["kk"][t] / g["k"][t]).astype(np.float32)
        o, M = capi.op_wkv(7, g["r"][t:t + 1], g["k"][t:t + 1], g["v"][t:t + 1], g["w"][t:t + 1], M, a=g["a"][t:t + 1],
                           k_k=kk_over_k, k_a=zeros, r_k=zeros)
        outs.append(o)
    got = np.concatenate(outs)
    assert np.abs(got - group_norm(g["out"])).max() <= 2e-3
    assert np.abs(M - g["S"]).max() / np.abs(g["S"]).max() <= 2e-5
