#!/bin/bash
# Round-2 GPU visit E: A/B of the k_sim_wide options (L2 prefetch, staged best-table), per-level cost by path length.
TAG=${1:-r2e}
O=gpurun_out
mkdir -p $O
for dbg in 0 1 2 3; do
  TZ_WIDE_DEBUG=$dbg timeout 600 python bench.py --workload cfg4 --sim-warps 4 --skip-cpu --skip-e2e --steps 3 2>$O/${TAG}_go_d$dbg.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('go_9x9 W=4 debug=$dbg', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us')" | tee -a $O/${TAG}_ab.log
  TZ_WIDE_DEBUG=$dbg timeout 600 python bench.py --workload cfg3 --sim-warps 2 --skip-cpu --skip-e2e --steps 6 2>$O/${TAG}_oth_d$dbg.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('othello-weighted W=2 debug=$dbg', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us')" | tee -a $O/${TAG}_ab.log
  TZ_WIDE_DEBUG=$dbg timeout 600 python bench.py --workload cfg3 --sim-warps 4 --skip-cpu --skip-e2e --steps 6 2>$O/${TAG}_oth4_d$dbg.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('othello-weighted W=4 debug=$dbg', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us')" | tee -a $O/${TAG}_ab.log
done
timeout 300 python scripts/phase_r2.py wide go_9x9 1024 800 1600 4 > $O/${TAG}_phase_wide_go.log 2>&1; cat $O/${TAG}_phase_wide_go.log
timeout 300 python scripts/phase_r2.py wide othello 512 200 400 2 weighted > $O/${TAG}_phase_wide_othello.log 2>&1; cat $O/${TAG}_phase_wide_othello.log
