#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_gpu.log
echo "== bench default"; timeout 900 python bench.py --steps 128 --warmup 8 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_r01.json; tail -n 3 gpurun_out/bench_r01.err
export B200RWKV_BENCH_PROMPT=0 B200RWKV_BENCH_CPU_STEPS=0 B200RWKV_GRAPH=0
K='regex:gemm_kernel|wkv_kernel|ln_mix|ln_out|embed_ln0'
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 969 -c 646 --csv \
   --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full gemm"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 700 -c 8 \
   -o gpurun_out/r01_gemm_full -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
