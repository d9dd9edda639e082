#!/bin/bash
# usage: tools/dp_ab.sh <tag> ...   - times the 10k-layer shortest path with build variants
mkdir -p gpurun_out; : > gpurun_out/dp_ab.jsonl
for t in "$@"; do
  if [ "$t" = default ]; then unset ACRO_B200_LIB; else export ACRO_B200_LIB=$PWD/acrobotics_b200/libacro_b200_$t.so; fi
  echo "{\"tag\": \"$t\"}" >> gpurun_out/dp_ab.jsonl
  python tools/kernel_bench.py --n 4000000 --reps 5 --only planner >> gpurun_out/dp_ab.jsonl 2>> gpurun_out/dp_ab.err
done
cut -c1-200 gpurun_out/dp_ab.jsonl
