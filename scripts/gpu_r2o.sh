#!/bin/bash
# Round-2 GPU visit O: final-code evidence -- bench (both arms), ncu launch list of a configs[1] step, ncu --set full captures of
# k_sim, k_sim_wide (go_9x9 shape) and k_reroot_bulk, instruction counts of the two re-root gathers.
TAG=${1:-r2o}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -3 $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; python -c "
import json; d=json.load(open('$O/${TAG}_bench_cfg2.json')); print('cfg2', d['value']/1e6, d['ms_per_step'], 'e2e', d['e2e']['value']/1e6, 'lockstep', d['e2e']['lockstep']['value']/1e6, 'roof', d['roofline']['frac'], 'reroot', d['roofline']['reroot']['frac'], d['roofline']['reroot']['avg_launch_us'])"; tail -3 $O/${TAG}_bench_cfg2.err
timeout 600 python bench.py --workload cfg5 --skip-cpu --steps 8 > $O/${TAG}_bench_cfg5.json 2> $O/${TAG}_bench_cfg5.err; python -c "
import json; d=json.load(open('$O/${TAG}_bench_cfg5.json')); print('cfg5', d['value']/1e6, d['ms_per_step'], 'e2e', d['e2e']['value']/1e6, 'lockstep', d['e2e']['lockstep']['value']/1e6)"; tail -3 $O/${TAG}_bench_cfg5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 560 --csv --log-file $O/${TAG}_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline > $O/${TAG}_ncu_launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sim -s 600 -c 3 -o $O/${TAG}_ksim \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline > $O/${TAG}_ncu_ksim_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reroot -s 4 -c 2 -o $O/${TAG}_reroot_cfg2 \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline > $O/${TAG}_ncu_reroot_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sim_wide -s 3000 -c 3 -o $O/${TAG}_kwide_go \
    python bench.py --workload cfg4 --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline > $O/${TAG}_ncu_kwide_run.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_reroot -s 3 -c 1 -o $O/${TAG}_reroot_go \
    python bench.py --workload cfg4 --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline > $O/${TAG}_ncu_reroot_go_run.log 2>&1
for impl in bulk ldgsts; do
TZ_REROOT_IMPL=$impl timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_reroot -s 4 -c 2 --csv --log-file $O/${TAG}_reroot_inst_$impl.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline > /dev/null 2>&1
grep -E "inst_executed|time_duration|dram__bytes" $O/${TAG}_reroot_inst_$impl.csv | awk -F, '{print $5, $(NF-2), $(NF-1), $NF}' | tail -8
done
ls -la $O/*.ncu-rep
