#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/pytest_gpu.log | cut -c1-250
export B200RWKV_BENCH_CPU_STEPS=0
run() {
  timeout 600 python bench.py --steps 64 --warmup 4 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$1.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$1: ms/step %.3f tok/s %.0f e2e %.0f step_frac %.3f"%(d["ms_per_step"], d["value"], d["e2e"]["value"], r["step_frac"]))
except Exception as e: print("ERR $1", e, open("gpurun_out/bench_$1.err").read()[-800:])
PY
}
run base
B200RWKV_KR_GRID=128 run kr128
B200RWKV_PREFETCH_BLOCKS=28 run pf28
B200RWKV_PREFETCH_BLOCKS=0 run pf0
B200RWKV_GEMM_RING=0 run ring0
