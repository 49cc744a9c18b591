#!/bin/bash
# Round-2 GPU visit AG: narrow selector scoring its register slots in pairs.
TAG=${1:-r2ag}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_user_fns.py -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -3 $O/${TAG}_pytest_gpu.log
run() {
  timeout 600 python bench.py --workload $1 --skip-cpu --skip-e2e --steps 8 $2 2>$O/${TAG}_$1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us frac', r['frac'], '; reroot', r['reroot']['avg_launch_us'], 'us frac', r['reroot']['frac'])" | tee -a $O/${TAG}_bench.log
}
run cfg2; run cfg5; run cfg2; run cfg5
