#!/bin/bash
# Round-end validation + evidence: GPU tests, smoke, default bench (both arms), ncu launch list and one full-set layer.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-200
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench (default flags)"; timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -c 2500 gpurun_out/bench_n1.json
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; tail -c 900 gpurun_out/bench_ref.json
export B200RWKV_BENCH_PROMPT=0 B200RWKV_BENCH_CPU_STEPS=0
K='regex:gemm_kernel|wkv_kernel|ln_mix|ln_out|embed_ln0|pre6_kernel'
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1400 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
echo "rc=$?"; tail -n 1 gpurun_out/ncu_list.log | cut -c1-200; wc -l gpurun_out/launches.csv
echo "== ncu full: one layer (prefetch chain off so DRAM traffic is attributed to the launch that uses it)"
B200RWKV_PREFETCH_BLOCKS=0 timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 463 -c 7 \
   -o gpurun_out/prof_layer -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -n 1 gpurun_out/ncu_full.log | cut -c1-200; ls -la gpurun_out/prof_layer.ncu-rep
