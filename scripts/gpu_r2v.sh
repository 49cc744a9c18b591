#!/bin/bash
# Round-2 GPU visit V: two path levels per scoring warp and round in k_sim_wide (TZ_WIDE_DEBUG bit 2), A/B on the go_9x9 shape.
TAG=${1:-r2v}
O=gpurun_out
mkdir -p $O
TZ_WIDE_DEBUG=4 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cta_per_tree or go_9x9" > $O/${TAG}_pytest_u2.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_u2.log; tail -4 $O/${TAG}_pytest_u2.log
for rep in 1 2; do
for dbg in 0 4; do
  TZ_WIDE_DEBUG=$dbg timeout 600 python bench.py --workload cfg4 --skip-cpu --skip-e2e --steps 4 2>$O/${TAG}_cfg4_$dbg.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('cfg4 dbg=$dbg', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us frac', r['frac'])" | tee -a $O/${TAG}_ab.log
done
done
