#!/bin/bash
# Round-2 GPU visit AA: narrow selector (one lane per path level) straight-line over its FM register slots.
TAG=${1:-r2aa}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -4 $O/${TAG}_pytest_gpu.log
for wl in cfg2 cfg5 cfg2 cfg5; do
  timeout 600 python bench.py --workload $wl --skip-cpu --skip-e2e --steps 8 2>$O/${TAG}_$wl.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$wl', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us frac', r['frac'], '; reroot', r['reroot']['avg_launch_us'], 'us frac', r['reroot']['frac'], 'ordinary', (d.get('ordinary_launches') or {}).get('value'))" | tee -a $O/${TAG}_bench.log
done
