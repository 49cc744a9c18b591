#!/bin/bash
# Round-2 GPU visit AK: which launch mode per shape?  (search launched programmatically or not) x (leaf stand-in cooperating or not)
TAG=${1:-r2ak}
O=gpurun_out
mkdir -p $O
run() {
  timeout 600 python bench.py --workload $1 --skip-cpu --skip-e2e --skip-roofline --steps 6 $2 2>$O/${TAG}_$1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 [$2]', round(d['value']/1e6,2), 'M sims/s', round(d['ms_per_step'],3), 'ms')" | tee -a $O/${TAG}_modes.log
}
for wl in cfg2 cfg3 cfg4 cfg5; do
  run $wl "--no-pdl"
  run $wl "--pdl --pdl-bits 1 --leaf-pdl 0"
  run $wl "--pdl --pdl-bits 1 --leaf-pdl 1"
  run $wl "--pdl --pdl-bits 3 --leaf-pdl 1"
done
