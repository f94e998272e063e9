#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_gpu.log
echo "== bench default"; B200RWKV_BENCH_CPU_STEPS=0 timeout 900 python bench.py --steps 64 --warmup 4 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_a.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("ms/step %.3f tok/s %.0f e2e_ms %.3f step_frac %.3f launches/step %d | prof gemm %.2f wkv %.2f ln %.2f"%(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], r["step_frac"], d["gpu_launches"]/d["steps"], r["class_ms_per_step"]["gemm"], r["class_ms_per_step"]["wkv"], r["class_ms_per_step"]["ln_mix"]))
except Exception as e: print("ERR", e, open("gpurun_out/bench_a.err").read()[-800:])
PY
