# development helper: run under gpurun from the repo root; writes logs to gpurun_out/
set -x
mkdir -p gpurun_out
for v in 4; do
  CBS_BR_VARIANT=$v timeout -s KILL 120 python temp_fhe_transciphering_b200/csrc/tools/brbench.py 592 || exit 1
  CBS_BR_VARIANT=$v timeout -s KILL 120 python temp_fhe_transciphering_b200/csrc/tools/brbench.py 1024
  CBS_BR_VARIANT=$v timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "blind_rotate or circuit_bootstrap or two_blocks" 2>&1 | tail -2
done
CBS_BR_VARIANT=4 timeout -s KILL 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline | cut -c1-400
CBS_BR_VARIANT=3 timeout -s KILL 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline | cut -c1-400
