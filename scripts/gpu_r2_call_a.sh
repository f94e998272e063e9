#!/bin/bash
# Round 2, first GPU call (1 GPU): the whole GPU suite on the default path, the kernel variants written at the end of
# round 1 (first time on hardware), parity at the BASELINE shapes, A/B of the variants on the bench, cfg 2 / cfg 4 lines.
# Every step has its own timeout and log under gpurun_out/; a step that kills its CUDA context does not take the rest down.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L; echo "host cpus: $(nproc)"
run() { # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "$O/$name.log" 2>&1
  echo "== $name rc=$? ($(( $(date +%s) - t0 )) s)"
}
run pytest_default 900 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_zfullsize.py
tail -n 3 $O/pytest_default.log | cut -c1-300
B200RWKV_TEST_EXPERIMENTAL=1 run pytest_experimental 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k experimental
tail -n 12 $O/pytest_experimental.log | cut -c1-300
rm -f $O/parity_fullsize.jsonl
run full_props 900 python -m pytest tests/test_gpu_zfullsize.py -m gpu -q -k "deterministic or roundtrip or front_half"
tail -n 3 $O/full_props.log | cut -c1-300
for c in v6-3b v7-2b9 v6-7b; do
  run full_parity_$c 1200 python -m pytest tests/test_gpu_zfullsize.py -m gpu -q -s -k "parity and $c"
  grep -E "passed|failed|Error|error|assert" $O/full_parity_$c.log | tail -n 6 | cut -c1-400
done
cat $O/parity_fullsize.jsonl 2>/dev/null | cut -c1-600
export B200RWKV_BENCH_CPU_STEPS=0
ab() { # name env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 64 --warmup 4 > "$O/ab_$name.json" 2> "$O/ab_$name.err"
  echo "== ab $name rc=$? $(python - "$O/ab_$name.json" <<'EOF'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("ms/step %.4f value %.1f e2e %.1f launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"] // d["steps"]))
except Exception as ex:
    print("no line:", ex)
EOF
)"
}
ab default B200RWKV_X=0
ab finisher_kr128 B200RWKV_FINISHER=1 B200RWKV_KR_GRID=128
ab wkv_stream2 B200RWKV_WKV_STREAM=2
ab wkv_stream4 B200RWKV_WKV_STREAM=4
ab default_again B200RWKV_X=0
env timeout 600 python bench.py --steps 64 --warmup 4 --exact > $O/ab_exact.json 2> $O/ab_exact.err; echo "== ab exact rc=$?"; tail -c 600 $O/ab_exact.json | cut -c1-600
unset B200RWKV_BENCH_CPU_STEPS
run bench_cfg2_v6_3b_b1 600 python bench.py --preset v6-3b --batch 1 --steps 64 --warmup 4
tail -n 1 $O/bench_cfg2_v6_3b_b1.log | cut -c1-1500
run bench_cfg4_v7_2b9_b8 600 python bench.py --preset v7-2b9 --batch 8 --steps 64 --warmup 4
tail -n 1 $O/bench_cfg4_v7_2b9_b8.log | cut -c1-1500
run steptrace_default 600 python scripts/gpu_steptrace.py
head -n 24 $O/steptrace_default.log | cut -c1-200
