#!/bin/bash
# Round-2 GPU visit F: re-root staging policy (few large CTAs for wide rows), byte tables as words; weighted register cap lifted.
TAG=${1:-r2f}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -8 $O/${TAG}_pytest_gpu.log
for wl in cfg2 cfg4 cfg3 cfg5; do
  timeout 600 python bench.py --workload $wl --skip-cpu --skip-e2e --steps 4 2>$O/${TAG}_$wl.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$wl', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us frac', r['frac'], '; reroot', r['reroot']['avg_launch_us'], 'us frac', r['reroot']['frac'])" | tee -a $O/${TAG}_bench.log
done
timeout 300 python scripts/phase_r2.py reroot go_9x9 1024 800 1600 4 > $O/${TAG}_phase_reroot_go.log 2>&1; tail -14 $O/${TAG}_phase_reroot_go.log
timeout 300 python scripts/phase_r2.py reroot connect_four 1024 128 256 1 > $O/${TAG}_phase_reroot_c4.log 2>&1; tail -14 $O/${TAG}_phase_reroot_c4.log
