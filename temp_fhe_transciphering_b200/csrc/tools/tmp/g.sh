set -x
timeout -s KILL 60 python temp_fhe_transciphering_b200/csrc/tools/brbench.py 256 || exit 1
timeout -s KILL 60 python temp_fhe_transciphering_b200/csrc/tools/brbench.py 8
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "blind_rotate or circuit_bootstrap or two_blocks" 2>&1 | tail -2
