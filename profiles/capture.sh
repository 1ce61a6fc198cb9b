#!/bin/bash
# How the numbers and ncu evidence under profiles/ are captured.  Run on a B200 box from the repo root, e.g.
#   gpurun --timeout 2700 -- 'bash profiles/capture.sh'
# Raw outputs land in gpurun_out/ (scratch); profiles/summarize.py turns them into the tracked CSV summaries.
# Every ncu run is preceded by the same command without ncu (B200_PROFILING.md), and nothing is timed under ncu.
set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
timeout -s KILL 900 python bench_sweep.py --out gpurun_out/sweep.jsonl > gpurun_out/sweep.log 2>&1
# launch list of one timed step (launches 155..266 of the process = the timed device-resident step)
timeout -s KILL 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
# full captures of the two blind-rotation kernels
timeout -s KILL 120 python temp_fhe_transciphering_b200/csrc/tools/brbench.py 1024 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 2 -c 1 -f -o gpurun_out/br_throughput \
    python temp_fhe_transciphering_b200/csrc/tools/brbench.py 1024 > gpurun_out/ncu2.log 2>&1
timeout -s KILL 120 python temp_fhe_transciphering_b200/csrc/tools/brbench.py 256 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_ll -s 2 -c 1 -f -o gpurun_out/br_team \
    python temp_fhe_transciphering_b200/csrc/tools/brbench.py 256 > gpurun_out/ncu3.log 2>&1
# then, here (no GPU needed):
#   python profiles/summarize.py launches gpurun_out/launches.csv 155 266 profiles/<round>_launch_summary.csv "<comment>"
#   python profiles/summarize.py full gpurun_out/br_throughput.ncu-rep profiles/<round>_blind_rotate_ncu_full.csv "<comment>"
