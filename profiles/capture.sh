#!/bin/bash
# How the numbers and ncu evidence under profiles/ are captured (round 2).  Run on a B200 box from the repo root:
#   gpurun --timeout 2400 -- 'bash profiles/capture.sh'
# Raw outputs land in gpurun_out/ (scratch); profiles/summarize.py turns them into the tracked CSV summaries.
# Every ncu run is preceded by the same command without ncu (B200_PROFILING.md), and nothing is timed under ncu.
set -x
mkdir -p gpurun_out
B=temp_fhe_transciphering_b200/csrc/tools/brbench.py
timeout -s KILL 600 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; cat gpurun_out/r02_bench_default.json
timeout -s KILL 900 python bench_sweep.py --out gpurun_out/r02_sweep_1gpu.jsonl > gpurun_out/r02_sweep_1gpu.log 2>&1
# launch list of one timed step
timeout -s KILL 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
# full captures: the throughput blind rotation on the 512-ciphertext lane shape the step launches, the team kernel, the trace
timeout -s KILL 120 python $B 512 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_blind_rotate_v4 -s 2 -c 1 -f -o gpurun_out/r02_br_v4_512 \
    python $B 512 > gpurun_out/ncu2.log 2>&1
timeout -s KILL 120 python $B 256 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_blind_rotate_ll -s 2 -c 1 -f -o gpurun_out/r02_br_team \
    python $B 256 > gpurun_out/ncu3.log 2>&1
timeout -s KILL 120 python $B 512 > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_v3 -s 2 -c 1 -f -o gpurun_out/r02_trace \
    python $B 512 > gpurun_out/ncu4.log 2>&1
# then, here (no GPU needed):
#   python profiles/summarize.py launches gpurun_out/r02_launches.csv FIRST LAST profiles/r02_launch_summary.csv "<comment>"
#   python profiles/summarize.py full gpurun_out/r02_br_v4_512.ncu-rep profiles/r02_blind_rotate_ncu_full.csv "<comment>"
#   python profiles/summarize.py full gpurun_out/r02_trace.ncu-rep profiles/r02_trace_ncu_full.csv "<comment>"
#   cuobjdump -sass temp_fhe_transciphering_b200/libcbs_b200.so | grep -oE '\b(UBLKCP|SYNCS[A-Z.0-9]*|DFMA|DADD|DMUL|SHFL\.IDX|UTMALDG|UTC[A-Z]*MMA|UTCCP[A-Z.]*|LDTM[.x0-9]*|STTM[.x0-9]*|HMMA|DMMA)\b' | sort | uniq -c > profiles/r02_sass_grep.txt
