#!/bin/bash
# compute-sanitizer over a slice of the GPU parity tests: memcheck (out-of-bounds / misaligned), racecheck (shared-memory
# hazards in k_reroot_all / k_sim's staged best-table), initcheck (reads of uninitialised device memory).
O=gpurun_out; mkdir -p $O; rm -f $O/sanitizer_summary.log
SEL='(test_api_fused_vs_oracle and (c4 or othello_weighted or go_muzero or very_deep)) or test_reroot_odd_row_sizes or (test_buffer_updates_match_reference_fixture and replay_small) or (test_programmatic_launch_vs_oracle and c4) or (test_deep_paths_vs_oracle and (deep_wide_go or deep_narrow) and True) or test_two_player_game_step_matches_reference_fixture or test_user_eval_and_env_functions or test_sample_matches_reference_fixture'
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_replay.py tests/test_gpu_two_player.py tests/test_gpu_user_fns.py -x -q -m gpu -k "$SEL" > $O/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $O/sanitizer_summary.log
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $O/sanitizer_$tool.log | tail -3 | tee -a $O/sanitizer_summary.log
done
