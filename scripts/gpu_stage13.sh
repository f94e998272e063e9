#!/bin/bash
mkdir -p gpurun_out
echo "== step trace (ring 2, prefetch 16)"; B200RWKV_GEMM_RING=2 B200RWKV_PREFETCH_BLOCKS=16 timeout 600 python scripts/gpu_steptrace.py > gpurun_out/steptrace.log 2>&1; echo "rc=$?"; tail -n 44 gpurun_out/steptrace.log | cut -c1-330
