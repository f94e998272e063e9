#!/bin/bash
# Round 2 (1 GPU): the th-row A16 layout / one-MMA-per-k-step change: fast GPU suite, scaling of the projections with the
# token tile count, bench default / exact / prefill, then the compute-sanitizer passes.
mkdir -p gpurun_out
O=gpurun_out
run() { local name=$1 t=$2; shift 2; local t0=$(date +%s); timeout "$t" "$@" > "$O/$name.log" 2>&1; echo "== $name rc=$? ($(( $(date +%s) - t0 )) s)"; }
run pytest_gpu 1200 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_zfullsize.py
tail -n 12 $O/pytest_gpu.log | cut -c1-300
run prefill_probe 600 python scripts/gpu_prefill_probe.py v6-3b
cat $O/prefill_probe.log | cut -c1-160
export B200RWKV_BENCH_CPU_STEPS=0
run bench_n1 600 python bench.py --steps 64 --warmup 4
run bench_exact 600 python bench.py --exact --steps 64 --warmup 4
unset B200RWKV_BENCH_CPU_STEPS
run bench_prefill 900 python bench.py --mode prefill --steps 1
python - <<'PY'
import json
for n in ("bench_n1", "bench_exact", "bench_prefill"):
    try:
        d = [json.loads(l) for l in open(f"gpurun_out/{n}.log") if l.startswith("{")][-1]; r = d["roofline"]
        print(n, "| value %.1f ms/step %.4f e2e %.1f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"]), r.get("class_us_per_step"), d.get("pass_breakdown"), d.get("parity_check"), r.get("tensor_tflops_achieved"))
    except Exception as ex:
        print(n, "no line", ex)
PY
run full_parity_7b 1200 python -m pytest tests/test_gpu_zfullsize.py -m gpu -q -s -k "parity and v6-7b"
grep -E "passed|failed|Error|error|assert" $O/full_parity_7b.log | tail -n 4 | cut -c1-300
bash scripts/gpu_sanitize.sh
