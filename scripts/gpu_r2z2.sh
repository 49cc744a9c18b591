#!/bin/bash
TAG=${1:-r2z}
O=gpurun_out
mkdir -p $O
for wl in cfg4 cfg4 cfg5; do
  timeout 600 python bench.py --workload $wl --skip-cpu --skip-e2e --steps 6 2>$O/${TAG}_$wl.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$wl', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us frac', r['frac'], '; reroot', r['reroot']['avg_launch_us'], 'us frac', r['reroot']['frac'])" | tee -a $O/${TAG}_bench.log
done
