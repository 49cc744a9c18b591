#!/bin/bash
# Round-2 GPU visit AC: best-table staged by one bulk copy in k_sim_wide.
TAG=${1:-r2ac}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -4 $O/${TAG}_pytest_gpu.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_cta_per_tree_c_loop_vs_oracle and (c4 or go_muzero or very_deep)" > $O/${TAG}_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -2 $O/${TAG}_synccheck.log
run() {
  timeout 600 python bench.py --workload $1 --skip-cpu --skip-e2e --steps 6 $2 2>$O/${TAG}_$1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 $2', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us frac', r['frac'], '; reroot', r['reroot']['avg_launch_us'], 'us frac', r['reroot']['frac'], 'pdl', d['config']['programmatic_dependent_launch'])" | tee -a $O/${TAG}_bench.log
}
run cfg4; run cfg3; run cfg4; run cfg3
