# development helper: run under gpurun from the repo root; writes logs to gpurun_out/
set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout -s KILL 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
timeout -s KILL 900 python bench_sweep.py --out gpurun_out/sweep_r01b.jsonl > gpurun_out/sweep.log 2>&1; tail -3 gpurun_out/sweep.log | cut -c1-200
timeout -s KILL 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
timeout -s KILL 120 python temp_fhe_transciphering_b200/csrc/tools/brbench.py 1024 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 2 -c 1 -f -o gpurun_out/br_x2_r01 python temp_fhe_transciphering_b200/csrc/tools/brbench.py 1024 > gpurun_out/ncu2.log 2>&1
timeout -s KILL 120 python temp_fhe_transciphering_b200/csrc/tools/brbench.py 256 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_ll -s 2 -c 1 -f -o gpurun_out/br_ll_r01 python temp_fhe_transciphering_b200/csrc/tools/brbench.py 256 > gpurun_out/ncu3.log 2>&1
tail -n 2 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
