#!/bin/bash
mkdir -p gpurun_out
export B200RWKV_GEMM_RING=2 B200RWKV_PREFETCH_BLOCKS=16
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_gpu.log | cut -c1-250
export B200RWKV_BENCH_CPU_STEPS=0
for v in "new" "old"; do
  if [ $v = old ]; then export B200RWKV_OLD_GRID=1; fi
  timeout 600 python bench.py --steps 64 --warmup 4 > gpurun_out/bench_g$v.json 2> gpurun_out/bench_g$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_g$v.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("grid=$v ms/step %.3f tok/s %.0f e2e %.0f step_frac %.3f launches/step %d roofline %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["step_frac"], d["gpu_launches"]/d["steps"], {k:r[k] for k in ("achieved","frac")}))
except Exception as e: print("ERR", e, open("gpurun_out/bench_g$v.err").read()[-800:])
PY
done
unset B200RWKV_OLD_GRID
echo "== step trace"; timeout 600 python scripts/gpu_steptrace.py > gpurun_out/steptrace.log 2>&1; echo "rc=$?"; tail -n 42 gpurun_out/steptrace.log | cut -c1-250
