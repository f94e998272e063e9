#!/bin/bash
# Round 2, second GPU call (1 GPU): sampling front half, RWKV-7 with the 2.9B LoRA ranks (default + exact), L2 prefetch probe,
# in-situ roofline on the three bench configs, step trace of the 3B / batch-1 and 2.9B / batch-8 shapes.
mkdir -p gpurun_out
O=gpurun_out
run() { local name=$1 t=$2; shift 2; local t0=$(date +%s); timeout "$t" "$@" > "$O/$name.log" 2>&1; echo "== $name rc=$? ($(( $(date +%s) - t0 )) s)"; }
run pytest_sampling 600 python -m pytest tests/test_gpu_sampling.py -m gpu -q
tail -n 15 $O/pytest_sampling.log | cut -c1-300
run pytest_v7_ranks 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "2b9_lora_ranks"
grep -E "worst rel|passed|failed|Error|assert " $O/pytest_v7_ranks.log | tail -n 20 | cut -c1-300
run prefetch_probe 600 python scripts/gpu_prefetch_probe.py
cat $O/prefetch_probe.log | cut -c1-200
export B200RWKV_BENCH_CPU_STEPS=0
run bench_7b 600 python bench.py --steps 64 --warmup 4
python - <<'PY'
import json
for n in ("bench_7b",):
    try:
        d = json.loads(open(f"gpurun_out/{n}.log").read().strip().splitlines()[-1]); r = d["roofline"]
        print(n, "ms/step", d["ms_per_step"], "frac", r["frac"], "step_frac", r["step_frac"], "class_us", r["class_us_per_step"], "between", r["between_windows_us"], "p10/50/90", r["step_ms_p10_p50_p90"])
        print(json.dumps(r["per_launch_class"]))
    except Exception as ex:
        print(n, "no line", ex)
PY
for cfg in "v6-3b 1" "v7-2b9 8"; do
  set -- $cfg
  B200RWKV_BENCH_PRESET=$1 B200RWKV_BENCH_BATCH=$2 run steptrace_$1 600 python scripts/gpu_steptrace.py
  head -n 20 $O/steptrace_$1.log | cut -c1-160
  grep -E "marginal" $O/steptrace_$1.log | cut -c1-160
done
