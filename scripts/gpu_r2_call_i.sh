#!/bin/bash
# Round 2 (1 GPU): decode epilogue restored, multi-tile epilogue with compile-time activation: fast suite, scaling probe,
# bench default (with the precision-1 line) and prefill.
mkdir -p gpurun_out
O=gpurun_out
run() { local name=$1 t=$2; shift 2; local t0=$(date +%s); timeout "$t" "$@" > "$O/$name.log" 2>&1; echo "== $name rc=$? ($(( $(date +%s) - t0 )) s)"; }
run pytest_gpu 1200 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_zfullsize.py
tail -n 6 $O/pytest_gpu.log | cut -c1-300
run prefill_probe 600 python scripts/gpu_prefill_probe.py v6-3b
grep -E "tokens per step|gemm_56MiB|gemm_0MiB|gemm_50MiB" $O/prefill_probe.log | cut -c1-160
run bench_n1 600 python bench.py --steps 64 --warmup 4
run bench_prefill 900 python bench.py --mode prefill --steps 1
python - <<'PY'
import json
for n in ("bench_n1", "bench_prefill"):
    try:
        d = [json.loads(l) for l in open(f"gpurun_out/{n}.log") if l.startswith("{")][-1]; r = d["roofline"]
        print(n, "| value %.1f ms/step %.4f e2e %.1f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"]), r.get("class_us_per_step"), d.get("precision1"), d.get("pass_breakdown"), r.get("tensor_tflops_achieved"))
        if "per_launch_class" in r: print("   ", {k: (v["launches"], round(v["avg_us"], 2)) for k, v in r["per_launch_class"].items()})
    except Exception as ex:
        print(n, "no line", ex)
PY
