#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_gpu.log | cut -c1-250
export B200RWKV_BENCH_CPU_STEPS=0
for v in "0 0" "0 10" "2 10" "2 0" "0 16" "2 16"; do
  set -- $v
  B200RWKV_GEMM_RING=$1 B200RWKV_PREFETCH_BLOCKS=$2 timeout 600 python bench.py --steps 64 --warmup 4 > gpurun_out/bench_r$1_p$2.json 2> gpurun_out/bench_r$1_p$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r$1_p$2.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("ring=$1 prefetch=$2 ms/step %.3f tok/s %.0f e2e %.0f step_frac %.3f launches/step %d"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["step_frac"], d["gpu_launches"]/d["steps"]))
except Exception as e: print("ERR", e, open("gpurun_out/bench_r$1_p$2.err").read()[-800:])
PY
done
echo "== step trace (ring 2, prefetch 10)"; B200RWKV_GEMM_RING=2 B200RWKV_PREFETCH_BLOCKS=10 timeout 600 python scripts/gpu_steptrace.py > gpurun_out/steptrace.log 2>&1; echo "rc=$?"; tail -n 26 gpurun_out/steptrace.log | cut -c1-200
