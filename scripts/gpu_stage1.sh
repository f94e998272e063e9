#!/bin/bash
# First GPU bring-up: intermediates vs oracle, parity tests, small + full bench.  Logs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
echo "== debug" ; timeout 600 python scripts/gpu_debug.py > gpurun_out/debug.log 2>&1; echo "debug rc=$?"
tail -n 60 gpurun_out/debug.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 30 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 5 gpurun_out/smoke.log
echo "== bench 1b6"; B200RWKV_BENCH_PRESET=v6-1b6 B200RWKV_BENCH_CPU_STEPS=2 timeout 600 python bench.py --steps 32 --warmup 4 > gpurun_out/bench_1b6.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/bench_1b6.log
echo "== bench 7b"; timeout 1200 python bench.py --steps 64 --warmup 4 > gpurun_out/bench_7b.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/bench_7b.log
