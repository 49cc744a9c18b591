#!/bin/bash
# One GPU box visit: parity tests, both bench arms, ncu launch list + one full capture of k_sim, scratch timings.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh r1e'
TAG=${1:-r1x}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -3 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; cat $O/${TAG}_bench_cfg2.json; tail -3 $O/${TAG}_bench_cfg2.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 3 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; cat $O/${TAG}_bench_ref.json
if [ "$2" != "short" ]; then
timeout 600 python scripts/quick_bench.py > $O/${TAG}_quick.log 2>&1; cat $O/${TAG}_quick.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 560 --csv --log-file $O/${TAG}_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline > $O/${TAG}_ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sim -s 600 -c 3 -o $O/${TAG}_ksim \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline > $O/${TAG}_ncu_full_run.log 2>&1
fi
ls -la $O
