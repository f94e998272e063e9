"""Layer-level INDEPENDENT pins of the RWKV-6 and RWKV-7 oracles: flash-linear-attention's own `RWKV6Attention.forward` /
`RWKV6FeedForward.forward` / `RWKV7Attention.forward` / `RWKV7FeedForward.forward` (fla 0.5.1, installed in this image) evaluated on
the synthetic tiny models' weights.

TEST INFRASTRUCTURE.  Run here (CPU container): `python -m oracle.make_golden_fla_layers` -> tests/golden/layer6_fla.npz, layer7_fla.npz.

fla's layers are written for GPUs: three leaf operations inside them are Triton kernels.  They are replaced by their definitional
equivalents so that the layer's OWN composition code -- token shift, the data-dependent lerp with its two LoRA stages and the
(r, w, k, v, g) ordering, the decay transform w -> -exp(w), the bonus, GroupNorm x gate, the output projection, the channel mix --
runs unmodified on the CPU:
  * `token_shift(x)`                     -> fla's own `token_shift_ref` (`ZeroPad2d((0, 0, 1, -1))(x) - x`)
  * `fused_recurrent_rwkv6(r, k, v, w, u)` -> fla's pure-torch `naive_recurrent_rwkv6` (the op tests/golden/wkv6_fla.npz already pins)
  * `fla.modules.GroupNorm.forward`      -> `torch.nn.functional.group_norm` with the module's own groups / weight / bias / eps
The weight mapping is the one of fla's `utils/convert_from_rwkv6.py` applied to the `.st` names (reference
assets/scripts/convert_safetensors.py:22-101): BlinkDL orders the five lerp branches (w, k, v, r, g), fla (r, w, k, v, g).
`norm_eps = 64e-5` is passed explicitly to the RWKV-6 layer (fla's default 1e-5 is not the official head_size_divisor = 8 value), so
that fixture does not pin the RWKV-6 GroupNorm epsilon; everything else in the two sub-layers it does.
RWKV-7 (`fuse_norm=False`: torch GroupNorm and `F.normalize`, no Triton): `fused_addcmul_rwkv7`, `fused_k_rwkv7`,
`gate_output_correction` -> fla's own `torch_addcmul_rwkv7`, `k_update_ref`, `gate_output_correction_ref`; the recurrence -> fla's
pure-torch `dplr_recurrence` with a = -kk, b = kk * a exactly as the layer's chunk branch calls it (already pinned by
tests/golden/wkv7_fla.npz).  There the GroupNorm epsilon is fla's own `head_dim * norm_eps` = 64e-5, and the decay constant its own
-0.6065306597126334: both pinned.
"""
from __future__ import annotations

import os
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
TOKENS = [1, 5, 9, 33, 2, 7, 300, 41, 41, 8]
PRESET, SEED = "tiny6", 0


def gen_v6():
    warnings.filterwarnings("ignore")
    import torch
    import torch.nn.functional as F
    from fla.layers import rwkv6 as fl
    from fla.models.rwkv6 import modeling_rwkv6 as fm
    from fla.ops.rwkv6.recurrent_naive import naive_recurrent_rwkv6

    from ai00_server_b200 import synth
    from oracle import rwkv_numpy as O

    from fla.modules.token_shift import token_shift_ref

    def token_shift_cpu(x, cu_seqlens=None, *a, **k):
        assert cu_seqlens is None
        return token_shift_ref(x)

    def recurrent_cpu(r, k, v, w, u, scale=1.0, initial_state=None, output_final_state=False, cu_seqlens=None):
        hf = lambda t: t.transpose(1, 2).contiguous()                    # [B, T, H, D] -> [B, H, T, D]
        o, ht = naive_recurrent_rwkv6(hf(r), hf(k), hf(v), hf(w), u, scale=scale, initial_state=initial_state,
                                      output_final_state=output_final_state)
        return o.transpose(1, 2).contiguous(), ht

    fl.token_shift = token_shift_cpu
    fm.token_shift = token_shift_cpu
    fl.fused_recurrent_rwkv6 = recurrent_cpu

    shp = synth.PRESETS[PRESET]
    w = O.parse_st(synth.make_st(PRESET, seed=SEED, force_numpy=True))
    C, H, Dm, Dd, Fh = shp.C, shp.H, shp.Dm, shp.Dd, shp.F
    t32 = lambda a: torch.from_numpy(np.asarray(a, np.float32))

    # the oracle (f32 contract) supplies the sub-layers' inputs: LayerNorm'ed hidden rows of every token
    orc = O.Oracle(w, "f32")
    orc.trace = {}
    state = orc.state_init()
    rows = {k: [] for k in ("xx1", "part_att", "xx2", "rr", "part_ffn")}
    L = shp.L
    per_layer = [{k: [] for k in rows} for _ in range(L)]
    for tok in TOKENS:
        orc._token(int(tok), state)
        for l in range(L):
            for k in rows:
                per_layer[l][k].append(orc.trace[f"{l}.{k}"].copy())

    rec = {"tokens": np.asarray(TOKENS, np.int64), "layers": np.asarray(L)}
    for l in range(L):
        a, f = f"blocks.{l}.att.", f"blocks.{l}.ffn."
        att = fl.RWKV6Attention(mode="chunk", hidden_size=C, expand_k=1.0, expand_v=1.0, num_heads=H, proj_low_rank_dim=Dm,
                                gate_low_rank_dim=Dd, norm_eps=64e-5, layer_idx=l)
        order = [3, 0, 1, 2, 4]                                          # fla (r, w, k, v, g) <- BlinkDL (w, k, v, r, g)
        names = ["time_mix_w", "time_mix_k", "time_mix_v", "time_mix_r", "time_mix_g"]
        w1 = np.asarray(w[a + "time_mix_w1"], np.float32).reshape(5, Dm, C)[order].reshape(5 * Dm, C)      # .st: [5*Dm, C]
        w2 = np.asarray(w[a + "time_mix_w2"], np.float32)[order]                                          # .st: [5, C, Dm]
        sd = {
            "x_proj.0.mu": t32(w[a + "time_mix_x"]).reshape(C),
            "x_proj.0.linear.weight": t32(w1),
            "x_proj.2.weight": t32(np.transpose(w2, (1, 0, 2)).reshape(C, 5 * Dm)),                         # [C, (n r)]
            "x_bias": t32(np.stack([np.asarray(w[a + names[i]], np.float32).reshape(C) for i in order])),
            "r_proj.linear.weight": t32(w[a + "receptance.weight"]),
            "k_proj.linear.weight": t32(w[a + "key.weight"]),
            "v_proj.linear.weight": t32(w[a + "value.weight"]),
            "g_proj.linear.weight": t32(w[a + "gate.weight"]),
            "w_proj.linear.lora.0.weight": t32(w[a + "time_decay_w1"]),                                   # .st: [Dd, C]
            "w_proj.linear.lora.2.weight": t32(w[a + "time_decay_w2"]),                                   # .st: [C, Dd]
            "w_proj.linear.lora.2.bias": t32(w[a + "time_decay"]).reshape(C),
            "bonus": t32(w[a + "time_first"]).reshape(H, C // H),
            "g_norm.weight": t32(w[a + "ln_x.weight"]).reshape(C),
            "g_norm.bias": t32(w[a + "ln_x.bias"]).reshape(C),
            "o_proj.weight": t32(w[a + "output.weight"]),
        }
        missing, unexpected = att.load_state_dict(sd, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        gn = att.g_norm
        att.g_norm.forward = lambda x, gn=gn: F.group_norm(x.reshape(-1, x.shape[-1]), gn.num_groups, gn.weight, gn.bias,
                                                           gn.eps).reshape(x.shape)
        att.gate_fn = F.silu                                             # 'swish'
        ffn = fm.RWKV6FeedForward(hidden_size=C, intermediate_size=Fh, layer_idx=l)
        fsd = {
            "key.mu": t32(w[f + "time_mix_k"]).reshape(C), "key.linear.weight": t32(w[f + "key.weight"]),
            "receptance.mu": t32(w[f + "time_mix_r"]).reshape(C), "receptance.linear.weight": t32(w[f + "receptance.weight"]),
            "value.weight": t32(w[f + "value.weight"]),
        }
        missing, unexpected = ffn.load_state_dict(fsd, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        ffn.act_fn = lambda x: torch.relu(x) ** 2                        # 'sqrelu'
        xx1 = np.stack(per_layer[l]["xx1"]).astype(np.float32)
        xx2 = np.stack(per_layer[l]["xx2"]).astype(np.float32)
        with torch.no_grad():
            o_att = att(t32(xx1)[None])[0][0].numpy()
            o_ffn = ffn(t32(xx2)[None])[0][0].numpy()
        rec[f"att_in_{l}"], rec[f"att_out_{l}"] = xx1, o_att
        rec[f"ffn_in_{l}"], rec[f"ffn_out_{l}"] = xx2, o_ffn
        want_att = np.stack(per_layer[l]["part_att"])
        want_ffn = np.stack(per_layer[l]["rr"]) * np.stack(per_layer[l]["part_ffn"])
        ea = float(np.abs(o_att - want_att).max() / np.abs(want_att).max())
        ef = float(np.abs(o_ffn - want_ffn).max() / np.abs(want_ffn).max())
        print(f"layer {l}: fla time-mix vs oracle {ea:.2e}, fla channel-mix vs oracle {ef:.2e}")
    np.savez_compressed(os.path.join(GOLD, "layer6_fla.npz"), **rec)
    print("wrote", os.path.join(GOLD, "layer6_fla.npz"))


def gen_v7():
    warnings.filterwarnings("ignore")
    import torch
    from fla.layers import rwkv7 as fl
    from fla.models.rwkv7 import modeling_rwkv7 as fm
    from fla.modules.token_shift import token_shift_ref
    from fla.ops.generalized_delta_rule.dplr.naive import dplr_recurrence
    from fla.ops.rwkv7.fused_addcmul import torch_addcmul_rwkv7
    from fla.ops.rwkv7.fused_k_update import k_update_ref
    from fla.ops.rwkv7.gate_output_correction import gate_output_correction_ref

    from ai00_server_b200 import synth
    from oracle import rwkv_numpy as O

    def token_shift_cpu(x, cu_seqlens=None, cache=None, output_cache=False, **k):
        assert cu_seqlens is None and cache is None
        d = token_shift_ref(x)
        return (d, x[:, -1]) if output_cache else d

    def recurrent_cpu(r, w, k, v, kk, a, scale=1.0, initial_state=None, output_final_state=False, cu_seqlens=None):
        hf = lambda t: t.transpose(1, 2).contiguous()                    # [B, T, H, D] -> [B, H, T, D]
        n = r.shape[-1]
        # the layer's chunk branch: chunk_rwkv7(r, w, k, v, a = -kk, b = kk * a, scale = 1); dplr_recurrence scales q by n^-1/2
        o, ht = dplr_recurrence(hf(r) * (n ** 0.5), hf(k), hf(v), hf(-kk), hf(kk * a), hf(w), initial_state=initial_state)
        return o.transpose(1, 2).contiguous(), ht

    fl.token_shift = token_shift_cpu
    fm.token_shift = token_shift_cpu
    fl.fused_addcmul_rwkv7 = torch_addcmul_rwkv7
    fl.fused_k_rwkv7 = k_update_ref
    fl.gate_output_correction = gate_output_correction_ref
    fl.fused_mul_recurrent_rwkv7 = recurrent_cpu

    preset = "tiny7"
    shp = synth.PRESETS[preset]
    w = O.parse_st(synth.make_st(preset, seed=SEED, force_numpy=True))
    C, H, L, Fh = shp.C, shp.H, shp.L, shp.F
    t32 = lambda a: torch.from_numpy(np.asarray(a, np.float32))
    orc = O.Oracle(w, "f32")
    orc.trace = {}
    state = orc.state_init()
    keys = ("xx1", "part_att", "xx2", "part_ffn")
    per_layer = [{k: [] for k in keys} for _ in range(L)]
    for tok in TOKENS:
        orc._token(int(tok), state)
        for l in range(L):
            for k in keys:
                per_layer[l][k].append(orc.trace[f"{l}.{k}"].copy())
    rec = {"tokens": np.asarray(TOKENS, np.int64), "layers": np.asarray(L)}
    v_first = None
    for l in range(L):
        a, f = f"blocks.{l}.att.", f"blocks.{l}.ffn."
        att = fl.RWKV7Attention(mode="chunk", hidden_size=C, head_dim=C // H, decay_low_rank_dim=shp.Dd, gate_low_rank_dim=shp.Dg,
                                a_low_rank_dim=shp.Da, v_low_rank_dim=shp.Dv, layer_idx=l, fuse_norm=False, num_hidden_layers=L)
        vec = lambda n: t32(w[a + n]).reshape(-1)
        sd = {f"x_{c}": t32(w[a + f"x_{c}"]).reshape(1, 1, C) for c in "rwkvag"}
        sd.update({
            "k_k": vec("k_k"), "k_a": vec("k_a"), "r_k": t32(w[a + "r_k"]).reshape(H, C // H),
            "r_proj.weight": t32(w[a + "receptance.weight"]), "k_proj.weight": t32(w[a + "key.weight"]),
            "v_proj.weight": t32(w[a + "value.weight"]), "o_proj.weight": t32(w[a + "output.weight"]),
            "w_lora.lora.0.weight": t32(w[a + "w1"]), "w_lora.lora.2.weight": t32(w[a + "w2"]), "w_lora.lora.2.bias": vec("w0"),
            "a_lora.lora.0.weight": t32(w[a + "a1"]), "a_lora.lora.2.weight": t32(w[a + "a2"]), "a_lora.lora.2.bias": vec("a0"),
            "g_lora.lora.0.weight": t32(w[a + "g1"]), "g_lora.lora.2.weight": t32(w[a + "g2"]),
            "g_norm.weight": vec("ln_x.weight"), "g_norm.bias": vec("ln_x.bias"),
        })
        if l > 0:
            sd.update({"v_lora.lora.0.weight": t32(w[a + "v1"]), "v_lora.lora.2.weight": t32(w[a + "v2"]), "v_lora.lora.2.bias": vec("v0")})
        missing, unexpected = att.load_state_dict(sd, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        att.eval()
        ffn = fm.RWKV7FeedForward(hidden_size=C, intermediate_size=Fh, layer_idx=l, num_hidden_layers=L)
        missing, unexpected = ffn.load_state_dict({"x_k": t32(w[f + "x_k"]).reshape(C), "key.weight": t32(w[f + "key.weight"]),
                                                   "value.weight": t32(w[f + "value.weight"])}, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        ffn.act_fn = lambda x: torch.relu(x) ** 2                        # 'sqrelu'
        xx1 = np.stack(per_layer[l]["xx1"]).astype(np.float32)
        xx2 = np.stack(per_layer[l]["xx2"]).astype(np.float32)
        with torch.no_grad():
            o_att, _, _, v_first = att(t32(xx1)[None], v_first=v_first)
            o_att = o_att[0].numpy()
            o_ffn = ffn(t32(xx2)[None])[0][0].numpy()
        rec[f"att_in_{l}"], rec[f"att_out_{l}"] = xx1, o_att
        rec[f"ffn_in_{l}"], rec[f"ffn_out_{l}"] = xx2, o_ffn
        want_att, want_ffn = np.stack(per_layer[l]["part_att"]), np.stack(per_layer[l]["part_ffn"])
        ea = float(np.abs(o_att - want_att).max() / np.abs(want_att).max())
        ef = float(np.abs(o_ffn - want_ffn).max() / np.abs(want_ffn).max())
        print(f"v7 layer {l}: fla time-mix vs oracle {ea:.2e}, fla channel-mix vs oracle {ef:.2e}")
    np.savez_compressed(os.path.join(GOLD, "layer7_fla.npz"), **rec)
    print("wrote", os.path.join(GOLD, "layer7_fla.npz"))


def gen_models():
    """Whole-model pins: fla's `RWKV6ForCausalLM` / `RWKV7ForCausalLM` (`fuse_norm=False`: torch LayerNorm / GroupNorm) on the tiny
    models' weights, same leaf replacements as above -> tests/golden/model6_fla.npz, model7_fla.npz (logits of every token)."""
    warnings.filterwarnings("ignore")
    import torch
    import torch.nn.functional as F
    from fla.layers import rwkv6 as fl6
    from fla.layers import rwkv7 as fl7
    from fla.models.rwkv6 import RWKV6Config, RWKV6ForCausalLM
    from fla.models.rwkv6 import modeling_rwkv6 as fm6
    from fla.models.rwkv7 import RWKV7Config, RWKV7ForCausalLM
    from fla.models.rwkv7 import modeling_rwkv7 as fm7
    from fla.modules.token_shift import token_shift_ref
    from fla.ops.generalized_delta_rule.dplr.naive import dplr_recurrence
    from fla.ops.rwkv6.recurrent_naive import naive_recurrent_rwkv6
    from fla.ops.rwkv7.fused_addcmul import torch_addcmul_rwkv7
    from fla.ops.rwkv7.fused_k_update import k_update_ref
    from fla.ops.rwkv7.gate_output_correction import gate_output_correction_ref

    from ai00_server_b200 import synth
    from oracle import rwkv_numpy as O

    def token_shift_cpu(x, cu_seqlens=None, cache=None, output_cache=False, **k):
        assert cu_seqlens is None and cache is None
        d = token_shift_ref(x)
        return (d, x[:, -1]) if output_cache else d

    hf = lambda t: t.transpose(1, 2).contiguous()

    def rec6(r, k, v, w, u, scale=1.0, initial_state=None, output_final_state=False, cu_seqlens=None):
        o, ht = naive_recurrent_rwkv6(hf(r), hf(k), hf(v), hf(w), u, scale=scale, initial_state=initial_state,
                                      output_final_state=output_final_state)
        return o.transpose(1, 2).contiguous(), ht

    def rec7(r, w, k, v, kk, a, scale=1.0, initial_state=None, output_final_state=False, cu_seqlens=None):
        o, ht = dplr_recurrence(hf(r) * (r.shape[-1] ** 0.5), hf(k), hf(v), hf(-kk), hf(kk * a), hf(w), initial_state=initial_state)
        return o.transpose(1, 2).contiguous(), ht

    for mod in (fl6, fm6, fl7, fm7):
        mod.token_shift = token_shift_cpu
    fl6.fused_recurrent_rwkv6 = rec6
    fl7.fused_addcmul_rwkv7, fl7.fused_k_rwkv7, fl7.gate_output_correction, fl7.fused_mul_recurrent_rwkv7 = \
        torch_addcmul_rwkv7, k_update_ref, gate_output_correction_ref, rec7
    t32 = lambda a: torch.from_numpy(np.asarray(a, np.float32))
    sqrelu = lambda x: torch.relu(x) ** 2

    for ver, preset in ((6, "tiny6"), (7, "tiny7")):
        shp = synth.PRESETS[preset]
        w = O.parse_st(synth.make_st(preset, seed=SEED, force_numpy=True))
        C, H, L, Fh, V = shp.C, shp.H, shp.L, shp.F, shp.V
        if ver == 6:
            cfg = RWKV6Config(hidden_size=C, expand_k=1.0, expand_v=1.0, intermediate_size=Fh, num_hidden_layers=L, num_heads=H,
                              proj_low_rank_dim=shp.Dm, gate_low_rank_dim=shp.Dd, norm_eps=1e-5, fuse_norm=False, fuse_cross_entropy=False,
                              vocab_size=V, use_cache=False)
            model = RWKV6ForCausalLM(cfg)
        else:
            cfg = RWKV7Config(hidden_size=C, intermediate_size=Fh, num_hidden_layers=L, head_dim=C // H, decay_low_rank_dim=shp.Dd,
                              gate_low_rank_dim=shp.Dg, a_low_rank_dim=shp.Da, v_low_rank_dim=shp.Dv, norm_eps=1e-5, fuse_norm=False,
                              fuse_cross_entropy=False, vocab_size=V, use_cache=False)
            model = RWKV7ForCausalLM(cfg)
        sd = {"model.embeddings.weight": t32(w["emb.weight"]), "lm_head.weight": t32(w["head.weight"]),
              "model.norm.weight": t32(w["ln_out.weight"]).reshape(C), "model.norm.bias": t32(w["ln_out.bias"]).reshape(C),
              "model.layers.0.pre_norm.weight": t32(w["blocks.0.ln0.weight"]).reshape(C),
              "model.layers.0.pre_norm.bias": t32(w["blocks.0.ln0.bias"]).reshape(C)}
        for l in range(L):
            b, a, f, m = f"blocks.{l}.", f"blocks.{l}.att.", f"blocks.{l}.ffn.", f"model.layers.{l}."
            vec = lambda n: t32(w[n]).reshape(-1)
            sd.update({m + "attn_norm.weight": vec(b + "ln1.weight"), m + "attn_norm.bias": vec(b + "ln1.bias"),
                       m + "ffn_norm.weight": vec(b + "ln2.weight"), m + "ffn_norm.bias": vec(b + "ln2.bias"),
                       m + "attn.g_norm.weight": vec(a + "ln_x.weight"), m + "attn.g_norm.bias": vec(a + "ln_x.bias"),
                       m + "attn.o_proj.weight": t32(w[a + "output.weight"]),
                       m + "ffn.key.weight" if ver == 7 else m + "ffn.key.linear.weight": t32(w[f + "key.weight"]),
                       m + "ffn.value.weight": t32(w[f + "value.weight"])})
            if ver == 6:
                order = [3, 0, 1, 2, 4]
                names = ["time_mix_w", "time_mix_k", "time_mix_v", "time_mix_r", "time_mix_g"]
                w1 = np.asarray(w[a + "time_mix_w1"], np.float32).reshape(5, shp.Dm, C)[order].reshape(5 * shp.Dm, C)
                w2 = np.asarray(w[a + "time_mix_w2"], np.float32)[order]
                sd.update({m + "attn.x_proj.0.mu": vec(a + "time_mix_x"), m + "attn.x_proj.0.linear.weight": t32(w1),
                           m + "attn.x_proj.2.weight": t32(np.transpose(w2, (1, 0, 2)).reshape(C, 5 * shp.Dm)),
                           m + "attn.x_bias": t32(np.stack([np.asarray(w[a + names[i]], np.float32).reshape(C) for i in order])),
                           m + "attn.r_proj.linear.weight": t32(w[a + "receptance.weight"]), m + "attn.k_proj.linear.weight": t32(w[a + "key.weight"]),
                           m + "attn.v_proj.linear.weight": t32(w[a + "value.weight"]), m + "attn.g_proj.linear.weight": t32(w[a + "gate.weight"]),
                           m + "attn.w_proj.linear.lora.0.weight": t32(w[a + "time_decay_w1"]),
                           m + "attn.w_proj.linear.lora.2.weight": t32(w[a + "time_decay_w2"]),
                           m + "attn.w_proj.linear.lora.2.bias": vec(a + "time_decay"), m + "attn.bonus": t32(w[a + "time_first"]).reshape(H, C // H),
                           m + "ffn.key.mu": vec(f + "time_mix_k"), m + "ffn.receptance.mu": vec(f + "time_mix_r"),
                           m + "ffn.receptance.linear.weight": t32(w[f + "receptance.weight"])})
            else:
                sd.update({m + f"attn.x_{c}": t32(w[a + f"x_{c}"]).reshape(1, 1, C) for c in "rwkvag"})
                sd.update({m + "attn.k_k": vec(a + "k_k"), m + "attn.k_a": vec(a + "k_a"), m + "attn.r_k": t32(w[a + "r_k"]).reshape(H, C // H),
                           m + "attn.r_proj.weight": t32(w[a + "receptance.weight"]), m + "attn.k_proj.weight": t32(w[a + "key.weight"]),
                           m + "attn.v_proj.weight": t32(w[a + "value.weight"]),
                           m + "attn.w_lora.lora.0.weight": t32(w[a + "w1"]), m + "attn.w_lora.lora.2.weight": t32(w[a + "w2"]),
                           m + "attn.w_lora.lora.2.bias": vec(a + "w0"),
                           m + "attn.a_lora.lora.0.weight": t32(w[a + "a1"]), m + "attn.a_lora.lora.2.weight": t32(w[a + "a2"]),
                           m + "attn.a_lora.lora.2.bias": vec(a + "a0"),
                           m + "attn.g_lora.lora.0.weight": t32(w[a + "g1"]), m + "attn.g_lora.lora.2.weight": t32(w[a + "g2"]),
                           m + "ffn.x_k": vec(f + "x_k")})
                if l > 0:
                    sd.update({m + "attn.v_lora.lora.0.weight": t32(w[a + "v1"]), m + "attn.v_lora.lora.2.weight": t32(w[a + "v2"]),
                               m + "attn.v_lora.lora.2.bias": vec(a + "v0")})
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        model.eval()
        for blk in model.model.layers:
            blk.ffn.act_fn = sqrelu
            if ver == 6:
                gn = blk.attn.g_norm
                eps6 = 64e-5                      # official head_size_divisor = 8: not fla's default, see the module docstring
                blk.attn.g_norm.forward = lambda x, gn=gn: F.group_norm(x.reshape(-1, x.shape[-1]), gn.num_groups, gn.weight, gn.bias,
                                                                        eps6).reshape(x.shape)
                blk.attn.gate_fn = F.silu
        with torch.no_grad():
            out = model(input_ids=torch.tensor([TOKENS]), use_cache=True)
        logits = out.logits[0].numpy()
        orc = O.Oracle(w, "f32")
        want, want_state = orc.run(TOKENS, orc.state_init(), full=True)
        err = float(np.abs(logits - want).max() / np.abs(want).max())
        print(f"v{ver} {preset}: fla ForCausalLM logits vs oracle {err:.2e}, argmax equal {bool((logits.argmax(1) == want.argmax(1)).all())}")
        # the recurrent state fla's cache holds after the run, per layer: token-shift rows of both sub-layers and the WKV state as
        # S[head][key][value] (fla keeps [K, V] for both versions; the oracle's RWKV-7 state is [value][key])
        N = C // H
        rec = {"tokens": np.asarray(TOKENS, np.int64), "logits": logits.astype(np.float32)}
        worst = 0.0
        for l in range(L):
            st = out.past_key_values[l]
            S = st["recurrent_state"][0].numpy().astype(np.float32)                      # [H, K, V]
            rec[f"att_shift_{l}"] = st["conv_state"].reshape(-1).numpy().astype(np.float32)
            rec[f"ffn_shift_{l}"] = st["ffn_state"].reshape(-1).numpy().astype(np.float32)
            rec[f"wkv_kv_{l}"] = S
            o_s = want_state[l, 1:1 + N].reshape(N, H, N).transpose(1, 0, 2)             # oracle rows: [head][i][j]
            o_kv = o_s if ver == 6 else o_s.transpose(0, 2, 1)
            worst = max(worst, float(np.abs(S - o_kv).max() / np.abs(o_kv).max()),
                        float(np.abs(rec[f"att_shift_{l}"] - want_state[l, 0]).max()), float(np.abs(rec[f"ffn_shift_{l}"] - want_state[l, N + 1]).max()))
        print(f"v{ver} {preset}: fla cache (shift rows, WKV state) vs oracle state, worst {worst:.2e}")
        np.savez_compressed(os.path.join(GOLD, f"model{ver}_fla.npz"), **rec)


if __name__ == "__main__":
    gen_v6()
    gen_v7()
    gen_models()
