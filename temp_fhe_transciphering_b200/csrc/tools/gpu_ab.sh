# development helper: run under gpurun from the repo root; writes logs to gpurun_out/
set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
