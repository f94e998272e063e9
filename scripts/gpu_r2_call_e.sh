#!/bin/bash
# Round 2, validation call (1 GPU) of the hardened library: whole GPU suite (full-size parity separately per shape), bench
# lines of the four configs + exact mode, step traces after the split-K change.
mkdir -p gpurun_out
O=gpurun_out
run() { local name=$1 t=$2; shift 2; local t0=$(date +%s); timeout "$t" "$@" > "$O/$name.log" 2>&1; echo "== $name rc=$? ($(( $(date +%s) - t0 )) s)"; }
run pytest_gpu 1200 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_zfullsize.py
tail -n 12 $O/pytest_gpu.log | cut -c1-300
rm -f $O/parity_fullsize.jsonl
run full_props 900 python -m pytest tests/test_gpu_zfullsize.py -m gpu -q -k "deterministic or roundtrip or front_half"
tail -n 3 $O/full_props.log | cut -c1-300
for c in v6-3b v7-2b9 v6-7b; do
  run full_parity_$c 1200 python -m pytest tests/test_gpu_zfullsize.py -m gpu -q -s -k "parity and $c"
  grep -E "passed|failed|Error|error|assert" $O/full_parity_$c.log | tail -n 6 | cut -c1-400
done
cat $O/parity_fullsize.jsonl 2>/dev/null | cut -c1-700
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
tail -n 3 $O/smoke.log
run bench_n1 900 python bench.py --steps 64 --warmup 4
run bench_cfg2_v6_3b_b1 600 python bench.py --preset v6-3b --batch 1 --steps 64 --warmup 4
run bench_cfg4_v7_2b9_b8 600 python bench.py --preset v7-2b9 --batch 8 --steps 64 --warmup 4
B200RWKV_BENCH_CPU_STEPS=0 run bench_exact 600 python bench.py --exact --steps 64 --warmup 4
run bench_prefill 900 python bench.py --mode prefill --steps 1
python - <<'PY'
import json
for n in ("bench_n1", "bench_cfg2_v6_3b_b1", "bench_cfg4_v7_2b9_b8", "bench_exact", "bench_prefill"):
    try:
        d = [json.loads(l) for l in open(f"gpurun_out/{n}.log") if l.startswith("{")][-1]; r = d["roofline"]
        print(n, "|", d["metric"], "| value %.1f ms/step %.4f e2e %.1f frac %.3f step_frac %s launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], r.get("step_frac"), d.get("gpu_launches")), d.get("cpu_baseline"), d.get("parity_check"), r.get("class_us_per_step"), r.get("between_windows_us"), r.get("tensor_tflops_achieved"))
    except Exception as ex:
        print(n, "no line", ex)
PY
