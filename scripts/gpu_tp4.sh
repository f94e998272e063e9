#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
echo "== tp worker (4 ranks)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 tests/tp_worker.py > gpurun_out/tp4_worker.log 2>&1; echo "rc=$?"; grep -E "world=|TP_OK|Error|watchdog" gpurun_out/tp4_worker.log | tail -12
echo "== bench N=4"; B200RWKV_BENCH_CPU_STEPS=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 64 --warmup 4 > gpurun_out/bench_n4.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/bench_n4.log | cut -c1-900
