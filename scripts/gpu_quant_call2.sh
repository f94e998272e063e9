#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_quant.py -q -m gpu 2>&1 | tail -30 > gpurun_out/quant_pytest.log
tail -6 gpurun_out/quant_pytest.log
timeout 300 python scripts/gpu_quant_probe.py > gpurun_out/quant_probe_ts.log 2>&1
grep -v "^  gemm   \(16\|8\) MiB" gpurun_out/quant_probe_ts.log | cut -c1-330 | tail -60
for q in int8 nf4; do
  timeout 240 python bench.py --quant $q --steps 64 --warmup 4 > gpurun_out/bench_quant_$q.log 2>&1
  grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_quant_$q.log | head -1
done
