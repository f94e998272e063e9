#!/bin/bash
# Profiling pass: launch list (serialised per-kernel durations), ncu full set on the GEMM, PDL/graph A-B.
mkdir -p gpurun_out
export B200RWKV_BENCH_PROMPT=0 B200RWKV_BENCH_CPU_STEPS=0
K='regex:gemm_kernel|wkv_kernel|ln_mix|ln_out|embed_ln0'
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 969 -c 646 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
echo "rc=$?"; tail -n 2 gpurun_out/ncu_list.log | cut -c1-300
echo "== ncu full (one layer of GEMM launches)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 700 -c 8 \
   -o gpurun_out/prof_gemm -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -n 2 gpurun_out/ncu_full.log | cut -c1-300
for v in "1 1" "1 0" "0 1" "0 0"; do
  set -- $v
  echo "== bench graph=$1 pdl=$2"
  B200RWKV_GRAPH=$1 B200RWKV_PDL=$2 timeout 600 python bench.py --steps 64 --warmup 4 > gpurun_out/bench_g$1_p$2.log 2>&1
  python - <<PY
import json
l=open("gpurun_out/bench_g$1_p$2.log").read().strip().splitlines()[-1]
try:
    d=json.loads(l); print("ms/step", d["ms_per_step"], "tok/s", d["value"], "e2e", d["e2e"]["value"])
except Exception as e: print("ERR", l[:300])
PY
done
