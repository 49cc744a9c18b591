#!/bin/bash
# Round-2 GPU visit AI: the connect_four shape at 1 K .. 64 K trees per GPU (where does the per-simulation launch stop being
# latency-bound?), final kernels.
TAG=${1:-r2ai}
O=gpurun_out
mkdir -p $O
for envs in 1024 4096 16384 65536; do
  for mode in "" "--no-pdl"; do
  timeout 600 python bench.py --workload cfg2 --envs $envs --skip-cpu --skip-e2e --steps 4 --warmup 3 $mode 2>$O/${TAG}_$envs.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('envs $envs $mode:', round(d['value']/1e6,1), 'M sims/s', round(d['ms_per_step'],3), 'ms; k_sim', round(r['avg_launch_us'],2), 'us/launch', round(r['achieved'],0), 'GB/s frac', round(r['frac'],3), '; reroot', round(r['reroot']['avg_launch_us'],1), 'us frac', round(r['reroot']['frac'],3), '; per-sim', {k: round(v,2) for k,v in r['per_simulation_us'].items()})" | tee -a $O/${TAG}_sweep.log
  done
done
