#!/bin/bash
# Round-2 GPU visit AS: L2 eviction policies in k_sim_wide (TZ_WIDE_DEBUG bit 4: path rows evict-last, embedding rows evict-first).
TAG=${1:-r2as}
O=gpurun_out
mkdir -p $O
TZ_WIDE_DEBUG=16 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cta_per_tree or go_9x9 or othello" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
run() {
  TZ_WIDE_DEBUG=$2 timeout 600 python bench.py --workload $1 --skip-cpu --skip-e2e --steps 6 2>$O/${TAG}_$1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 dbg=$2', round(d['value']/1e6,2), 'M sims/s', round(d['ms_per_step'],3), 'ms; launch', round(r['avg_launch_us'],2), 'us')" | tee -a $O/${TAG}_ab.log
}
for rep in 1 2; do run cfg4 0; run cfg4 16; run cfg3 0; run cfg3 16; done
