#!/bin/bash
# Round-2 GPU visit P: four lanes per path level in k_sim's narrow selector (quad_select).
TAG=${1:-r2p}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -12 $O/${TAG}_pytest_gpu.log
for rep in 1 2; do
  timeout 300 python bench.py --skip-cpu --skip-e2e --steps 24 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('this library:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; ordinary', d['ordinary_launches']['value']/1e6, d['roofline']['per_simulation_us'], d['roofline']['avg_launch_us'])" | tee -a $O/${TAG}_variants.log
if [ -d scratch_r1 ]; then (cd scratch_r1 && timeout 300 python bench.py --skip-cpu --skip-e2e --skip-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('round-1 library:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms')") | tee -a $O/${TAG}_variants.log; fi
done
for wl in cfg5 cfg1; do
  timeout 300 python bench.py --workload $wl --skip-cpu --skip-e2e --skip-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$wl:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms')" | tee -a $O/${TAG}_variants.log
done
