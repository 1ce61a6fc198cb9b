set -x
cd /root/repo; mkdir -p gpurun_out
./temp_fhe_transciphering_b200/bin/bench_fft > gpurun_out/fftbench_r01b.log 2>&1
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "blind_rotate or circuit_bootstrap" > gpurun_out/par_br.log 2>&1; tail -3 gpurun_out/par_br.log
for v in 2 3; do CBS_BR_VARIANT=$v python temp_fhe_transciphering_b200/csrc/tools/brbench.py 1024; CBS_BR_VARIANT=$v python temp_fhe_transciphering_b200/csrc/tools/brbench.py 592; done > gpurun_out/brbench_r01b.log 2>&1
CBS_BR_PROF=1 CBS_BR_VARIANT=3 python temp_fhe_transciphering_b200/csrc/tools/brbench.py 592 >> gpurun_out/brbench_r01b.log 2>&1
cat gpurun_out/fftbench_r01b.log gpurun_out/brbench_r01b.log
