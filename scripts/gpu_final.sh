#!/bin/bash
# Round-end validation + evidence (1 GPU): GPU tests, smoke, bench lines (headline 7B/16, cfg 2, cfg 4, exact mode, prefill),
# reference arm, ncu launch list of the bench command, one --set full capture of a layer and of the head launch.
mkdir -p gpurun_out
O=gpurun_out
run() { local name=$1 t=$2; shift 2; local t0=$(date +%s); timeout "$t" "$@" > "$O/$name.log" 2>&1; echo "== $name rc=$? ($(( $(date +%s) - t0 )) s)"; }
run pytest_gpu 1500 python -m pytest tests -m gpu -q
tail -n 5 $O/pytest_gpu.log | cut -c1-300
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
tail -n 3 $O/smoke.log
run bench_n1 900 python bench.py
tail -n 1 $O/bench_n1.log | cut -c1-2500
run bench_cfg2_v6_3b_b1 600 python bench.py --preset v6-3b --batch 1
run bench_cfg4_v7_2b9_b8 600 python bench.py --preset v7-2b9 --batch 8
B200RWKV_BENCH_CPU_STEPS=0 run bench_exact 600 python bench.py --exact --steps 64 --warmup 4
run bench_prefill 900 python bench.py --mode prefill --steps 1
python - <<'PY'
import json
for n in ("bench_cfg2_v6_3b_b1", "bench_cfg4_v7_2b9_b8", "bench_exact", "bench_prefill"):
    try:
        d = [json.loads(l) for l in open(f"gpurun_out/{n}.log") if l.startswith("{")][-1]; r = d["roofline"]
        print(n, d["metric"], "value %.1f ms/step %.4f e2e %.1f frac %.3f step_frac %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], r.get("step_frac")), d.get("cpu_baseline"), d.get("parity_check"))
    except Exception as ex:
        print(n, "no line", ex)
PY
run bench_ref 900 python bench.py --impl reference --steps 20 --warmup 5
tail -n 1 $O/bench_ref.log | cut -c1-900
export B200RWKV_BENCH_PROMPT=0 B200RWKV_BENCH_CPU_STEPS=0
K='regex:gemm_kernel|wkv_kernel|ln_mix|ln_out|embed_ln0|pre6_kernel|keep_rows'
echo "== ncu launch list (same command as the bench)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1600 --csv \
   --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 > $O/ncu_list.log 2>&1
echo "rc=$?"; wc -l $O/launches.csv
echo "== ncu full: one layer + the head (prefetch chain off needs the debug build; default build: traffic includes the prefetch)"
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 513 -c 7 \
   -o $O/prof_layer -f python bench.py --steps 2 --warmup 3 > $O/ncu_full.log 2>&1
echo "rc=$?"; ls -la $O/prof_layer.ncu-rep
timeout 900 ncu --set full --clock-control none -k "regex:gemm_kernel" -s 386 -c 1 \
   -o $O/prof_head -f python bench.py --steps 2 --warmup 3 > $O/ncu_head.log 2>&1
echo "rc=$?"; ls -la $O/prof_head.ncu-rep
