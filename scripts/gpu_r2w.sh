#!/bin/bash
# Round-2 GPU visit W: compute-sanitizer over the kernels added in round 2 (k_sim_wide: CTA per tree, k_reroot_bulk: bulk-copy
# re-root) -- memcheck, synccheck (mbarrier / barrier misuse), racecheck (shared-memory hazards), initcheck.
TAG=${1:-r2w}
O=gpurun_out; mkdir -p $O; rm -f $O/${TAG}_sanitizer_summary.log
SEL='(test_cta_per_tree_c_loop_vs_oracle and (c4 or othello_weighted or very_deep or wide_F300 or go_muzero)) or test_cta_per_tree_deep_weighted or test_cta_per_tree_rebuilds or test_reroot_odd_row_sizes or (test_api_fused_vs_oracle and (c4 or go_muzero)) or (test_deep_paths_vs_oracle and deep_wide_go)'
for tool in memcheck synccheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 30 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > $O/${TAG}_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $O/${TAG}_sanitizer_summary.log
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $O/${TAG}_sanitizer_$tool.log | tail -3 | tee -a $O/${TAG}_sanitizer_summary.log
done
