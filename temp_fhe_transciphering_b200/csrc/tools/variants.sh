#!/bin/bash
# development tool: parity tests + timing of the blind-rotation variants selected by CBS_BR_V (run on the GPU box)
out=gpurun_out/r02_brbench_v5.log
: > $out
for v in ${VARIANTS:-0 1 2 3 4 5}; do
  echo "== CBS_BR_V=$v" >> $out
  CBS_BR_V=$v timeout -s KILL 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "blind_rotate" 2>&1 | tail -3 >> $out
  for b in ${BATCHES:-592 4096}; do
    CBS_BR_V=$v timeout -s KILL 60 python temp_fhe_transciphering_b200/csrc/tools/brbench.py $b 2>&1 | tail -1 >> $out
  done
done
cat $out
