#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_gpu.log | cut -c1-250
echo "== step trace"; timeout 600 python scripts/gpu_steptrace.py > gpurun_out/steptrace.log 2>&1; echo "rc=$?"; tail -n 40 gpurun_out/steptrace.log | cut -c1-200
