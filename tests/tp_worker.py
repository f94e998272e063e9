"""torchrun worker: tensor-parallel engine (one process per GPU) vs the single-GPU engine.
Launched by tests/test_gpu_tp_multiproc.py and scripts; prints TP_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ai00_server_b200 import capi, runtime, synth, tp  # noqa: E402


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    exact = os.environ.get("B200RWKV_TEST_EXACT") == "1"          # precision 1 (f32 activations) on both sides
    for preset in sys.argv[1:] or ["small6", "small7", "small5"]:          # 8 heads each: every world size up to 8 divides them
        st = synth.make_st(preset, 0)
        m = runtime.Model(st, max_batch=4, token_chunk_size=32, device=local, rank=rank, world=world, exact=exact)
        tp.connect(m)
        rng = np.random.default_rng(3)
        runs = [rng.integers(1, 500, size=n).tolist() for n in (5, 1, 7)]
        args = ([0, 1, 2], [len(r) for r in runs], [t for r in runs for t in r],
                [capi.OPTION_FULL, capi.OPTION_LAST, capi.OPTION_LAST])
        for s in range(3):
            m.state.load(m.state.init(), s)
        got = np.concatenate(m.infer_raw(*args))
        for _ in range(3):                                     # decode steps on top
            got2 = np.concatenate(m.infer_raw([0, 1, 2], [1, 1, 1], [9, 8, 7], [capi.OPTION_LAST] * 3))
        if rank == 0:
            single = runtime.Model(st, max_batch=4, token_chunk_size=32, device=local, exact=exact)
            for s in range(3):
                single.state.load(single.state.init(), s)
            want = np.concatenate(single.infer_raw(*args))
            for _ in range(3):
                want2 = np.concatenate(single.infer_raw([0, 1, 2], [1, 1, 1], [9, 8, 7], [capi.OPTION_LAST] * 3))
            e1, e2 = rel_err(got, want), rel_err(got2, want2)
            tol = 1e-4 if exact else 1e-3          # f32 activations: only the f32 summation order differs between the shardings
            ok = e1 <= tol and e2 <= tol and (got.argmax(1) == want.argmax(1)).all() and (got2.argmax(1) == want2.argmax(1)).all()
            print(f"{preset}: world={world} exact={exact} prefill rel={e1:.2e} decode rel={e2:.2e} argmax_ok={ok}", flush=True)
            assert ok
            single.close()
        dist.barrier()
        m.close()
    if rank == 0:
        print("TP_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
