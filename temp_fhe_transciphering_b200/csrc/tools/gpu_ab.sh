set -x
for tv in 3 4; do
  CBS_TRACE_VARIANT=$tv timeout -s KILL 120 python temp_fhe_transciphering_b200/csrc/tools/brbench.py 1024
  CBS_TRACE_VARIANT=$tv timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "trace or circuit_bootstrap or two_blocks" 2>&1 | tail -2
done
