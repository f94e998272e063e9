#!/bin/bash
# Quantised projections on hardware: parity tests, then the 7B / batch-16 decode bench with Int8 and NF4 weights, then the
# prefill bench (for its parity spot check against the oracle floor).
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_quant.py -q -m gpu 2>&1 | tail -40 > gpurun_out/quant_pytest.log
tail -5 gpurun_out/quant_pytest.log
for q in int8 nf4; do
  timeout 240 python bench.py --quant $q --steps 64 --warmup 4 > gpurun_out/bench_quant_$q.log 2>&1
  tail -c 600 gpurun_out/bench_quant_$q.log; echo
done
timeout 300 python bench.py --mode prefill --steps 1 --warmup 3 > gpurun_out/bench_prefill2.log 2>&1
grep -o '"parity_check".*' gpurun_out/bench_prefill2.log | tail -1
