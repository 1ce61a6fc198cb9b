set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench_sweep.py --cbs-only --max-batch 1 --mini 2>&1 | tail -6 | cut -c1-250
