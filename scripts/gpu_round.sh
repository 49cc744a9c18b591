set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r1d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1d_pytest_gpu.log
tail -3 gpurun_out/r1d_pytest_gpu.log
timeout 600 python scripts/quick_bench.py > gpurun_out/r1d_quick.log 2>&1
cat gpurun_out/r1d_quick.log
timeout 600 python scripts/phase_clocks.py > gpurun_out/r1d_phase.log 2>&1
tail -40 gpurun_out/r1d_phase.log
