#!/bin/bash
# Round-2 GPU visit AM: k_sim_wide split around griddepcontrol.wait (round trips 1 and 2 before it) -- parity, launch modes.
TAG=${1:-r2am}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -3 $O/${TAG}_pytest_gpu.log
run() {
  timeout 600 python bench.py --workload $1 --skip-cpu --skip-e2e --skip-roofline --steps 6 $2 2>$O/${TAG}_$1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 [$2]', round(d['value']/1e6,2), 'M sims/s', round(d['ms_per_step'],3), 'ms')" | tee -a $O/${TAG}_modes.log
}
for wl in cfg3 cfg4; do
  run $wl "--no-pdl"
  run $wl "--pdl --pdl-bits 1 --leaf-pdl 0"
  run $wl "--pdl --pdl-bits 1 --leaf-pdl 1"
  run $wl "--pdl --pdl-bits 3 --leaf-pdl 1"
done
