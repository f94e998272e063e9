#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_gpu.log | cut -c1-250
export B200RWKV_BENCH_CPU_STEPS=0
timeout 600 python bench.py --steps 64 --warmup 4 > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_cur.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("ms/step %.3f tok/s %.0f e2e %.0f step_frac %.3f launches/step %d"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["step_frac"], d["gpu_launches"]/d["steps"]))
except Exception as e: print("ERR", e, open("gpurun_out/bench_cur.err").read()[-800:])
PY
echo "== step trace"; timeout 600 python scripts/gpu_steptrace.py > gpurun_out/steptrace.log 2>&1; echo "rc=$?"; tail -n 46 gpurun_out/steptrace.log | grep -v "gemm32MB\|gemm112MB" | cut -c1-250
