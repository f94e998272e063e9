"""ctypes wrapper of the C/OpenMP oracle (oracle/rwkv_ref.c).  TEST INFRASTRUCTURE ONLY — see the
header of rwkv_ref.c; built by ai00_server_b200.build.build_oracle() into oracle/liboracle_ref.so."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import rwkv_numpy as O

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_ref.so")
_P = C.c_void_p

_LAYER_FIELDS = ["ln1_w", "ln1_b", "ln2_w", "ln2_b", "mix_x", "mix_w", "mix_k", "mix_v", "mix_r", "mix_g", "mix_w1", "mix_w2",
                 "decay", "decay_w1", "decay_w2", "first", "wr", "wk", "wv", "wg", "wo", "lnx_w", "lnx_b", "fmix_k", "fmix_r",
                 "fk", "fr", "fv",
                 "x_r", "x_w", "x_k", "x_v", "x_a", "x_g", "w0", "w1", "w2", "a0", "a1", "a2", "v0", "v1", "v2", "g1", "g2",
                 "k_k", "k_a", "r_k", "fx_k"]


class RefLayer(C.Structure):
    _fields_ = [(n, _P) for n in _LAYER_FIELDS]


class RefModel(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("version", "L", "C", "F", "V", "H", "N", "Dm", "Dd", "act_f16", "Dw", "Da", "Dv", "Dg")] + \
               [(n, _P) for n in ("emb", "ln0_w", "ln0_b", "lnout_w", "lnout_b", "head", "layers")]


def _lib():
    lib = C.CDLL(LIB_PATH)
    lib.ref_decode_step.argtypes = [C.POINTER(RefModel), C.c_int32, _P, _P, _P]
    lib.ref_decode_step.restype = C.c_int32
    lib.ref_num_threads.restype = C.c_int32
    lib.ref_set_num_threads.argtypes = [C.c_int32]
    lib.ref_set_num_threads.restype = None
    return lib


class RefC:
    """v5/v6/v7 decode step on B slots; state is [B, L, N+2, C] f32 (web-rwkv layout per slot)."""

    def __init__(self, weights: dict[str, np.ndarray], act: str = "f16"):
        self.lib = _lib()
        self.w = weights
        i = O.model_info(weights)
        if i.version not in (5, 6, 7):
            raise ValueError("unsupported model version")
        self.info = i

        def p(name):
            a = weights[name]
            assert a.dtype == np.float16 and a.flags["C_CONTIGUOUS"]
            return a.ctypes.data

        self._layers = (RefLayer * i.num_layer)()
        v6, v7 = i.version == 6, i.version == 7
        dims = {"Dw": 0, "Da": 0, "Dv": 0, "Dg": 0}
        for l in range(i.num_layer):
            b, a, f = f"blocks.{l}.", f"blocks.{l}.att.", f"blocks.{l}.ffn."
            ly = self._layers[l]
            ly.ln1_w, ly.ln1_b, ly.ln2_w, ly.ln2_b = p(b + "ln1.weight"), p(b + "ln1.bias"), p(b + "ln2.weight"), p(b + "ln2.bias")
            ly.wr, ly.wk, ly.wv, ly.wo = (p(a + n + ".weight") for n in ("receptance", "key", "value", "output"))
            ly.lnx_w, ly.lnx_b = p(a + "ln_x.weight"), p(a + "ln_x.bias")
            ly.fk, ly.fv = p(f + "key.weight"), p(f + "value.weight")
            if v7:
                for n in ("x_r", "x_w", "x_k", "x_v", "x_a", "x_g", "w0", "w1", "w2", "a0", "a1", "a2", "g1", "g2", "k_k", "k_a", "r_k"):
                    setattr(ly, n, p(a + n))
                if l > 0:
                    ly.v0, ly.v1, ly.v2 = p(a + "v0"), p(a + "v1"), p(a + "v2")
                    dims["Dv"] = weights[a + "v1"].shape[0]
                ly.fx_k = p(f + "x_k")
                dims["Dw"], dims["Da"], dims["Dg"] = (weights[a + n].shape[0] for n in ("w1", "a1", "g1"))
                continue
            for n in ("k", "v", "r", "g"):
                setattr(ly, "mix_" + n, p(a + "time_mix_" + n))
            if v6:
                ly.mix_x, ly.mix_w = p(a + "time_mix_x"), p(a + "time_mix_w")
                ly.mix_w1, ly.mix_w2 = p(a + "time_mix_w1"), p(a + "time_mix_w2")
                ly.decay_w1, ly.decay_w2 = p(a + "time_decay_w1"), p(a + "time_decay_w2")
            ly.decay, ly.first = p(a + "time_decay"), p(a + "time_first")
            ly.wg = p(a + "gate.weight")
            ly.fmix_k, ly.fmix_r = p(f + "time_mix_k"), p(f + "time_mix_r")
            ly.fr = p(f + "receptance.weight")
        m = RefModel()
        m.version, m.L, m.C, m.F, m.V, m.H, m.N = i.version, i.num_layer, i.num_emb, i.num_hidden, i.num_vocab, i.num_head, i.head_size
        m.Dm, m.Dd, m.act_f16 = i.time_mix_adapter, i.time_decay_adapter, int(act == "f16")
        m.Dw, m.Da, m.Dv, m.Dg = dims["Dw"], dims["Da"], dims["Dv"], dims["Dg"]
        m.emb, m.head = p("emb.weight"), p("head.weight")
        m.ln0_w, m.ln0_b = p("blocks.0.ln0.weight"), p("blocks.0.ln0.bias")
        m.lnout_w, m.lnout_b = p("ln_out.weight"), p("ln_out.bias")
        m.layers = C.cast(self._layers, _P)
        self.model = m

    def num_threads(self) -> int:
        return int(self.lib.ref_num_threads())

    def set_num_threads(self, n: int) -> None:
        self.lib.ref_set_num_threads(int(n))

    def state_init(self, B: int) -> np.ndarray:
        i = self.info
        return np.zeros((B, i.num_layer, i.head_size + 2, i.num_emb), np.float32)

    def decode_step(self, tokens, state: np.ndarray) -> np.ndarray:
        tok = np.ascontiguousarray(tokens, dtype=np.int32)
        B = tok.size
        assert state.dtype == np.float32 and state.flags["C_CONTIGUOUS"] and state.shape[0] == B
        logits = np.empty((B, self.info.num_vocab), np.float32)
        rc = self.lib.ref_decode_step(C.byref(self.model), B, tok.ctypes.data, state.ctypes.data, logits.ctypes.data)
        assert rc == 0
        return logits
