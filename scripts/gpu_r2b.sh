#!/bin/bash
# Round-2 GPU visit B: CTA-per-tree kernel (k_sim_wide) + bulk-copy re-root parity, A/B numbers, hand-over experiment.
# usage: gpurun --timeout 2400 -- 'bash scripts/gpu_r2b.sh r2b'
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 120 ./scripts/microbench_handshake > $O/${TAG}_handshake.log 2>&1; echo "rc=$?" >> $O/${TAG}_handshake.log; cat $O/${TAG}_handshake.log
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -15 $O/${TAG}_pytest_gpu.log
# re-root A/B: bulk copies (default) vs the LDGSTS gather
for impl in bulk ldgsts; do
  TZ_REROOT_IMPL=$impl timeout 300 python bench.py --skip-cpu --skip-e2e --steps 12 2>$O/${TAG}_ab_${impl}.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']['reroot']; print('reroot $impl', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms;', r['avg_launch_us'], 'us', r['frac'])" | tee -a $O/${TAG}_reroot_ab.log
done
# round-1 library on the same box (regression check of the headline)
if [ -d scratch_r1 ]; then (cd scratch_r1 && timeout 300 python bench.py --skip-cpu --skip-e2e --skip-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('round-1 library:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms')") | tee $O/${TAG}_r1_same_box.log; fi
timeout 300 python bench.py --skip-cpu --skip-e2e --skip-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('this library:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms')" | tee -a $O/${TAG}_r1_same_box.log
# warps per tree on the wide shapes
for w in 1 2 4 8; do
  timeout 600 python bench.py --workload cfg4 --sim-warps $w --skip-cpu --skip-e2e --steps 3 2>$O/${TAG}_go_w$w.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('go_9x9 sim_warps=$w', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us', r['per_simulation_us'])" | tee -a $O/${TAG}_warps.log
  timeout 600 python bench.py --workload cfg3 --sim-warps $w --skip-cpu --skip-e2e --steps 6 2>$O/${TAG}_oth_w$w.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('othello-weighted sim_warps=$w', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us', r['per_simulation_us'])" | tee -a $O/${TAG}_warps.log
done
ls -la $O | tail -12
