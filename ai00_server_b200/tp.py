"""Tensor-parallel plumbing over torch.distributed (one process per GPU).

The engines exchange nothing but a 128-byte opaque handle per rank (a CUDA IPC handle of the
rank's comm block); this module all-gathers those blobs over whatever process group is up
(NCCL on the GPU box, gloo in the CPU tests) and hands them, rank-ordered, to
b200rwkv_tp_connect.  Every data-path exchange afterwards (partial-sum reads, gate blocks, logits
shards, rendezvous flags) happens inside the CUDA kernels over NVLink peer memory.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def shard_plan(info: dict, world: int, rank: int) -> dict:
    """Which slice of the model a rank owns (mirrors b200rwkv_engine::build):
    heads / R,K,V,G rows / ffn.key rows / head rows are column-parallel, att.output and ffn.value
    are row-parallel (partials summed by the next LN stage), everything small is replicated."""
    H, N, C, F, V = info["num_head"], info["head_size"], info["num_emb"], info["num_hidden"], info["num_vocab"]
    if H % world or F % (8 * world) or V % world:
        raise ValueError("heads / hidden / vocab do not shard evenly over the tensor-parallel world")
    Hl, Cl, Fl, Vl = H // world, C // world, F // world, V // world
    return {"heads": (rank * Hl, (rank + 1) * Hl), "channels": (rank * Cl, (rank + 1) * Cl),
            "hidden": (rank * Fl, (rank + 1) * Fl), "vocab": (rank * Vl, (rank + 1) * Vl),
            "partials_per_ln": world * max(s for s in (4, 3, 2, 1) if s <= 8 // world)}


def export_handle(model) -> np.ndarray:
    buf = np.zeros(capi.TP_HANDLE_BYTES, np.uint8)
    capi.check(capi.lib().b200rwkv_tp_export(model._h, capi.ptr(buf)), model._h)
    return buf


def gather_handles(local: np.ndarray, group=None) -> np.ndarray:
    """All-gather the per-rank handle blobs; returns [world, TP_HANDLE_BYTES] uint8, rank-ordered."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.from_numpy(local.copy()).to(dev)
    outs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine, group=group)
    return np.stack([o.cpu().numpy() for o in outs], 0)


def connect(model, group=None) -> None:
    """Wire a rank's engine to its peers (call on every rank)."""
    allh = np.ascontiguousarray(gather_handles(export_handle(model), group))
    capi.check(capi.lib().b200rwkv_tp_connect(model._h, capi.ptr(allh)), model._h)


def connect_local(models) -> None:
    """All ranks in this process (tests: several ranks on one GPU)."""
    arr = (C.c_void_p * len(models))(*[m._h for m in models])
    capi.check(capi.lib().b200rwkv_tp_connect_local(arr, len(models)))
