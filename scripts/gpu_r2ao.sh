#!/bin/bash
# Round-2 GPU visit AO: same-box A/B of the k_sim_wide split around griddepcontrol.wait (libtz_b200_split.so) against the
# committed library, alternating.
TAG=${1:-r2ao}
O=gpurun_out
mkdir -p $O
run() {
  TZ_B200_LIB=$3 timeout 600 python bench.py --workload $1 --skip-cpu --skip-e2e --skip-roofline --steps 6 $2 2>$O/${TAG}_$1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 [$2] $3', round(d['value']/1e6,2), 'M sims/s', round(d['ms_per_step'],3), 'ms')" | tee -a $O/${TAG}_ab.log
}
for rep in 1 2; do
for wl in cfg4 cfg3; do
  run $wl "--no-pdl" libtz_b200.so
  run $wl "--no-pdl" libtz_b200_split.so
  run $wl "--pdl --pdl-bits 3 --leaf-pdl 1" libtz_b200.so
  run $wl "--pdl --pdl-bits 3 --leaf-pdl 1" libtz_b200_split.so
done
done
