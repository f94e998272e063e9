#!/bin/bash
# correctness re-check + timing attribution by skipping kernel classes (results of skipped runs are garbage; timing only)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
export B200RWKV_BENCH_PROMPT=0 B200RWKV_BENCH_CPU_STEPS=0
for sk in 0 1 2 4 7 8 23 15; do
  B200RWKV_SKIP=$sk timeout 600 python bench.py --steps 64 --warmup 4 > gpurun_out/bench_skip$sk.log 2>&1
  python - <<PY
import json
l=open("gpurun_out/bench_skip$sk.log").read().strip().splitlines()[-1]
try:
    d=json.loads(l); r=d["roofline"]; print("skip=$sk ms/step %.3f tok/s %.0f e2e_ms %.3f | prof gemm %.3f wkv %.3f ln %.3f"%(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], r["class_ms_per_step"]["gemm"], r["class_ms_per_step"]["wkv"], r["class_ms_per_step"]["ln_mix"]))
except Exception as e: print("ERR", l[-400:])
PY
done
