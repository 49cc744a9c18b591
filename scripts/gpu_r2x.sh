#!/bin/bash
# Round-2 GPU visit X: ncu --set full with source of the weighted CTA-per-tree kernel on the othello shape (where does a level of
# the backup chain spend its ~1.1 us?)
TAG=${1:-r2x}
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sim_wide -s 900 -c 2 -o $O/${TAG}_kwide_othello \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-roofline > $O/${TAG}_ncu_run.log 2>&1
tail -3 $O/${TAG}_ncu_run.log
ls -la $O/${TAG}_*.ncu-rep
