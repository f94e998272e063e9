#!/bin/bash
# compute-sanitizer passes over the CI-sized GPU tests (memcheck, racecheck, synccheck); summaries under gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
T='tests/test_gpu_parity.py::test_logits_match_oracle tests/test_gpu_parity.py::test_short_ragged_steps_use_the_cluster_kernels tests/test_gpu_parity.py::test_f32_activation_mode_tracks_the_f32_oracle tests/test_gpu_sampling.py::test_topk_matches_the_full_vocabulary_sort tests/test_gpu_kernels.py'
for tool in memcheck racecheck synccheck; do
  t0=$(date +%s)
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 python -m pytest $T -m gpu -q -x -k "not small6 and not eng_wide" > $O/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$? ($(( $(date +%s) - t0 )) s)"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/sanitizer_$tool.log | tail -n 4
done
