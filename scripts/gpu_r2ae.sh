#!/bin/bash
# Round-2 GPU visit AE: re-root front in one round trip (root edge row + parents requested with the action).
TAG=${1:-r2ae}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -4 $O/${TAG}_pytest_gpu.log
run() {
  timeout 600 python bench.py --workload $1 --skip-cpu --skip-e2e --steps 8 $2 2>$O/${TAG}_$1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us frac', r['frac'], '; reroot', r['reroot']['avg_launch_us'], 'us frac', r['reroot']['frac'])" | tee -a $O/${TAG}_bench.log
}
run cfg2; run cfg3; run cfg4; run cfg5; run cfg2
timeout 300 python scripts/phase_r2.py reroot connect_four 1024 128 256 1 > $O/${TAG}_phase_reroot_c4.log 2>&1; tail -12 $O/${TAG}_phase_reroot_c4.log
