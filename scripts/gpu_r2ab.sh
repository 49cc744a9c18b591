#!/bin/bash
# Round-2 GPU visit AB: signed reduction keys + acq_rel fence in the weighted hand-over; ordinary launches for the othello / 2048
# shapes; go_9x9 with and without programmatic launches.
TAG=${1:-r2ab}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -4 $O/${TAG}_pytest_gpu.log
run() {
  timeout 600 python bench.py --workload $1 --skip-cpu --skip-e2e --steps 6 $2 2>$O/${TAG}_$1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us frac', r['frac'], '; reroot', r['reroot']['avg_launch_us'], 'us frac', r['reroot']['frac'], 'pdl', d['config']['programmatic_dependent_launch'], 'ordinary', (d.get('ordinary_launches') or {}).get('value'))" | tee -a $O/${TAG}_bench.log
}
run cfg2; run cfg3; run cfg3 --pdl; run cfg4; run cfg4 --pdl; run cfg5; run cfg5 --pdl; run cfg2
