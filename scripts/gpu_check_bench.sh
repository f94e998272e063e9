#!/bin/bash
mkdir -p gpurun_out
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_ref.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','steps')}, d['cpu_baseline']['cores'])"
echo "== bench (short)"; timeout 600 python bench.py --steps 32 --warmup 4 > gpurun_out/bench_n1_short.json 2> gpurun_out/bench_n1_short.err; echo "rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_n1_short.json').read().strip().splitlines()[-1]); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['traffic'], r['algorithmic_bytes_per_step_gemm'], d['cpu_baseline'])"
