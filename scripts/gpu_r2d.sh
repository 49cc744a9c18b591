#!/bin/bash
# Round-2 GPU visit D: split build; k_sim_wide with <= 72 registers (one wave), staged best-table, L2 row prefetch.
TAG=${1:-r2d}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -8 $O/${TAG}_pytest_gpu.log
for w in 4 8; do
  timeout 600 python bench.py --workload cfg4 --sim-warps $w --skip-cpu --skip-e2e --steps 3 2>$O/${TAG}_go_w$w.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('go_9x9 sim_warps=$w', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us', r['per_simulation_us'], 'reroot', r['reroot']['avg_launch_us'], r['reroot']['frac'])" | tee -a $O/${TAG}_warps.log
done
for w in 2 4; do
  timeout 600 python bench.py --workload cfg3 --sim-warps $w --skip-cpu --skip-e2e --steps 6 2>$O/${TAG}_oth_w$w.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('othello-weighted sim_warps=$w', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us', r['per_simulation_us'])" | tee -a $O/${TAG}_warps.log
done
timeout 300 python bench.py --skip-cpu --skip-e2e --steps 12 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg2:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms', d['roofline']['per_simulation_us'])" | tee -a $O/${TAG}_warps.log
timeout 300 python scripts/phase_r2.py wide go_9x9 1024 800 1600 4 > $O/${TAG}_phase_wide_go.log 2>&1; cat $O/${TAG}_phase_wide_go.log
timeout 300 python scripts/phase_r2.py wide othello 512 200 400 2 weighted > $O/${TAG}_phase_wide_othello.log 2>&1; cat $O/${TAG}_phase_wide_othello.log
timeout 300 python scripts/phase_r2.py reroot go_9x9 1024 800 1600 4 > $O/${TAG}_phase_reroot_go.log 2>&1; tail -14 $O/${TAG}_phase_reroot_go.log
