#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/pytest_gpu.log
export B200RWKV_BENCH_CPU_STEPS=0
for mg in 1 0; do
  echo "== bench 7b mega=$mg"; B200RWKV_MEGA=$mg timeout 600 python bench.py --steps 64 --warmup 4 > gpurun_out/bench_mega$mg.log 2>&1; echo "rc=$?"
  python - <<PY
import json
l=open("gpurun_out/bench_mega$mg.log").read().strip().splitlines()[-1]
try:
    d=json.loads(l); r=d["roofline"]; print("mega=$mg ms/step %.3f tok/s %.0f e2e_ms %.3f step_frac %.3f launches %d"%(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], r["step_frac"], d["gpu_launches"]))
except Exception as e: print("ERR", l[-600:])
PY
done
timeout 300 python scripts/gpu_trace.py 2>&1 | tail -34

