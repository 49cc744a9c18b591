#!/bin/bash
# Round-2 GPU visit AU: L2 eviction policies in k_sim_wide as a compile-time variant (libtz_b200_l2hint.so, -DTZ_WIDE_L2HINT)
# against the default library on the same box.
TAG=${1:-r2au}
O=gpurun_out
mkdir -p $O
TZ_B200_LIB=libtz_b200_l2hint.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cta_per_tree or go_9x9 or othello" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -2 $O/${TAG}_pytest.log
run() {
  TZ_B200_LIB=$2 timeout 600 python bench.py --workload $1 --skip-cpu --skip-e2e --skip-roofline --steps 6 2>$O/${TAG}_$1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 $2', round(d['value']/1e6,2), 'M sims/s', round(d['ms_per_step'],3), 'ms')" | tee -a $O/${TAG}_ab.log
}
for rep in 1 2; do run cfg4 libtz_b200.so; run cfg4 libtz_b200_l2hint.so; done
run cfg3 libtz_b200.so; run cfg3 libtz_b200_l2hint.so
