#!/bin/bash
# Round-2 GPU visit N: leaf stand-in with its round-1 parameter list; this library vs the round-1 library, alternating.
TAG=${1:-r2n}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -4 $O/${TAG}_pytest_gpu.log
for rep in 1 2 3; do
  timeout 300 python bench.py --skip-cpu --skip-e2e --skip-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('this library:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; ordinary', d['ordinary_launches']['value']/1e6)" | tee -a $O/${TAG}_variants.log
if [ -d scratch_r1 ]; then (cd scratch_r1 && timeout 300 python bench.py --skip-cpu --skip-e2e --skip-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('round-1 library:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms')") | tee -a $O/${TAG}_variants.log; fi
done
timeout 300 python bench.py --skip-cpu --skip-e2e --steps 12 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['roofline']['per_simulation_us'], d['roofline']['avg_launch_us'])" | tee -a $O/${TAG}_variants.log
