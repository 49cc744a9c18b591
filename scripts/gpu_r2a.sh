#!/bin/bash
# Round-2 GPU visit A: parity tests of the new ABI (registry, timeline, masked reset, bench shapes), the hand-over
# experiment, the restructured bench (headline + named configs at single-GPU shares), in-step L2 metrics of k_sim.
# usage: gpurun --timeout 2400 -- 'bash scripts/gpu_r2a.sh r2a'
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -5 $O/${TAG}_pytest_gpu.log
timeout 120 ./scripts/microbench_handshake > $O/${TAG}_handshake.log 2>&1; echo "rc=$?" >> $O/${TAG}_handshake.log; cat $O/${TAG}_handshake.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; cat $O/${TAG}_bench_cfg2.json; tail -5 $O/${TAG}_bench_cfg2.err
timeout 900 python bench.py --named cfg3,cfg4,cfg5 --skip-cpu --steps 8 > $O/${TAG}_bench_named.json 2> $O/${TAG}_bench_named.err; cat $O/${TAG}_bench_named.json; tail -5 $O/${TAG}_bench_named.err
# in-step (caches NOT flushed between kernels, single pass) L2 / DRAM counters of the per-simulation kernel
timeout 600 ncu --metrics lts__t_sector_hit_rate.pct,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --cache-control none --clock-control none -k regex:k_sim -s 700 -c 128 --csv --log-file $O/${TAG}_ksim_instep_l2.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline --no-graph > $O/${TAG}_ncu_instep_run.log 2>&1
tail -3 $O/${TAG}_ksim_instep_l2.csv
ls -la $O | tail -20
