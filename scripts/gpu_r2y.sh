#!/bin/bash
# Round-2 GPU visit Y: straight-line weighted_value (div_core + branch-free tz_expf): parity + othello / 2048 / connect_four benches.
TAG=${1:-r2y}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -4 $O/${TAG}_pytest_gpu.log
for wl in cfg3 cfg3 cfg2; do
  timeout 600 python bench.py --workload $wl --skip-cpu --skip-e2e --steps 6 2>$O/${TAG}_$wl.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$wl', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us frac', r['frac'], '; reroot', r['reroot']['avg_launch_us'], 'us frac', r['reroot']['frac'])" | tee -a $O/${TAG}_bench.log
done
timeout 300 python scripts/phase_r2.py wide othello 512 200 400 2 weighted > $O/${TAG}_phase_wide_othello.log 2>&1; tail -16 $O/${TAG}_phase_wide_othello.log
