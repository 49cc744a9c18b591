#!/bin/bash
# Round-2 GPU visit AR: in-step L2 / DRAM counters (caches NOT flushed between kernels, single pass) of the final per-simulation
# kernels: k_sim on configs[1], k_sim_wide on the go_9x9 and othello shapes.
TAG=${1:-r2ar}
O=gpurun_out
mkdir -p $O
M=lts__t_sector_hit_rate.pct,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --cache-control none --clock-control none -k regex:k_sim -s 700 -c 128 --csv --log-file $O/${TAG}_ksim_instep_l2.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline --no-graph > $O/${TAG}_run1.log 2>&1
timeout 600 ncu --metrics $M --cache-control none --clock-control none -k regex:k_sim_wide -s 3300 -c 200 --csv --log-file $O/${TAG}_kwide_go_instep_l2.csv \
    python bench.py --workload cfg4 --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline --no-graph > $O/${TAG}_run2.log 2>&1
timeout 600 ncu --metrics $M --cache-control none --clock-control none -k regex:k_sim_wide -s 900 -c 200 --csv --log-file $O/${TAG}_kwide_othello_instep_l2.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline --no-graph > $O/${TAG}_run3.log 2>&1
wc -l $O/${TAG}_*_instep_l2.csv
