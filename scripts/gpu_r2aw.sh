#!/bin/bash
# Round-2 GPU visit AW: register-capped k_sim instantiation for batches of more than 2 x 9 x SMs trees -- parity, and the
# connect_four shape at 4 K .. 64 K trees with the cap forced off / on (TZ_SIM_OCC).
TAG=${1:-r2aw}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "large_batch or full_size or test_programmatic_launch_vs_oracle" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -2 $O/${TAG}_pytest.log
for envs in 4096 16384 65536; do
  for occ in 0 1; do
  TZ_SIM_OCC=$occ timeout 600 python bench.py --workload cfg2 --envs $envs --no-pdl --skip-cpu --skip-e2e --steps 4 --warmup 3 2>$O/${TAG}_$envs.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('envs $envs capped=$occ:', round(d['value']/1e6,1), 'M sims/s', round(d['ms_per_step'],3), 'ms; k_sim', round(r['avg_launch_us'],2), 'us/launch frac', round(r['frac'],3))" | tee -a $O/${TAG}_sweep.log
  done
done
TZ_SIM_OCC=1 timeout 300 python bench.py --workload cfg2 --skip-cpu --skip-e2e --skip-roofline --steps 6 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('envs 1024 capped=1 (forced):', round(d['value']/1e6,1))" | tee -a $O/${TAG}_sweep.log
