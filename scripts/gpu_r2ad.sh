#!/bin/bash
# Round-2 GPU visit AD: final-code evidence -- tests, smoke, bench (both arms, named shapes), ncu launch list of a configs[1]
# step, ncu --set full captures of k_sim (configs[1]), k_sim_wide plain (go_9x9) and weighted (othello), phase stamps.
TAG=${1:-r2ad}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -4 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; python -c "
import json; d=json.load(open('$O/${TAG}_bench_cfg2.json')); print('cfg2', d['value']/1e6, d['ms_per_step'], 'e2e', d['e2e']['value']/1e6, 'lockstep', d['e2e']['lockstep']['value']/1e6, 'roof', d['roofline']['frac'], 'reroot', d['roofline']['reroot']['frac'], 'cpu', d['cpu_baseline']['value']/1e6, d['cpu_baseline']['numpy_port']['value'], 'ordinary', d['ordinary_launches']['value']/1e6)"; tail -3 $O/${TAG}_bench_cfg2.err
timeout 900 python bench.py --named cfg3,cfg4,cfg5 --skip-cpu --steps 8 > $O/${TAG}_bench_named.json 2> $O/${TAG}_bench_named.err; python -c "
import json; d=json.load(open('$O/${TAG}_bench_named.json'))
for n in d['config']['named']: print(n['name'], n['value']/1e6, 'M sims/s', n['ms_per_step'], 'ms e2e', n['e2e']['value']/1e6, 'lock', n['e2e']['lockstep']['value']/1e6, 'roof', n['roofline']['frac'], 'reroot', n['roofline']['reroot']['frac'])"; tail -3 $O/${TAG}_bench_named.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 3 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; cut -c1-200 $O/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 560 --csv --log-file $O/${TAG}_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline > $O/${TAG}_ncu_launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sim -s 600 -c 3 -o $O/${TAG}_ksim \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline > $O/${TAG}_ncu_ksim_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sim_wide -s 3000 -c 3 -o $O/${TAG}_kwide_go \
    python bench.py --workload cfg4 --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline > $O/${TAG}_ncu_kwide_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sim_wide -s 900 -c 3 -o $O/${TAG}_kwide_othello \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline > $O/${TAG}_ncu_kwide_othello_run.log 2>&1
timeout 200 python scripts/phase_r2.py wide othello 512 200 400 2 weighted > $O/${TAG}_phase_wide_othello.log 2>&1; tail -14 $O/${TAG}_phase_wide_othello.log
timeout 300 python scripts/phase_r2.py wide go_9x9 1024 800 1600 4 > $O/${TAG}_phase_wide_go.log 2>&1; tail -30 $O/${TAG}_phase_wide_go.log
ls -la $O/${TAG}_*.ncu-rep
