#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/pytest_gpu.log | cut -c1-250
export B200RWKV_BENCH_CPU_STEPS=0
for v in "1 1" "0 1" "0 0"; do
  set -- $v
  B200RWKV_FUSED_PRE=$1 B200RWKV_LN_CLUSTER=$2 timeout 600 python bench.py --steps 64 --warmup 4 > gpurun_out/bench_f$1_l$2.json 2> gpurun_out/bench_f$1_l$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_f$1_l$2.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("fused=$1 ln_cluster=$2 ms/step %.3f tok/s %.0f e2e %.0f step_frac %.3f launches/step %d | prof gemm %.2f wkv %.2f ln %.2f other %.2f"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["step_frac"], d["gpu_launches"]/d["steps"], r["class_ms_per_step"]["gemm"], r["class_ms_per_step"]["wkv"], r["class_ms_per_step"]["ln_mix"], r["class_ms_per_step"]["other"]))
except Exception as e: print("ERR", e, open("gpurun_out/bench_f$1_l$2.err").read()[-800:])
PY
done
