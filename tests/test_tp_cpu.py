"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes exchange the opaque
handle blobs exactly as bench.py does on the GPU box; the shard plan covers the model exactly."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from ai00_server_b200 import capi, synth, tp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_plan_partitions_the_model():
    info = capi.info_from_st(synth.make_st("small6", 0))
    for world in (1, 2, 4):
        plans = [tp.shard_plan(info, world, r) for r in range(world)]
        for key, total in (("heads", info["num_head"]), ("channels", info["num_emb"]), ("hidden", info["num_hidden"]),
                           ("vocab", info["num_vocab"])):
            edges = [p[key] for p in plans]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        assert plans[0]["partials_per_ln"] <= 8
    info7b = {"num_head": 64, "head_size": 64, "num_emb": 4096, "num_hidden": 14336, "num_vocab": 65536}
    assert tp.shard_plan(info7b, 8, 7)["heads"] == (56, 64)
    with pytest.raises(ValueError):
        tp.shard_plan(info, 3, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_handle_all_gather_over_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np
        import torch.distributed as dist
        from ai00_server_b200 import capi, tp
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        blob = np.full(capi.TP_HANDLE_BYTES, rank + 1, np.uint8)
        blob[0] = 200 + rank
        allh = tp.gather_handles(blob)
        assert allh.shape == (world, capi.TP_HANDLE_BYTES)
        for r in range(world):
            assert allh[r, 0] == 200 + r and (allh[r, 1:] == r + 1).all()
        dist.barrier()
        dist.destroy_process_group()
        sys.stdout.write("rank %d ok" % rank + chr(10))      # one write per rank: the two ranks share the pipe
        sys.stdout.flush()
    """))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout
