#!/bin/bash
# usage: tools/k2_ab.sh <out.jsonl> <tag> [<tag> ...]   ("default" = libacro_b200.so)
# K2_N="100000000 12500000" selects the batch sizes
out=$1; shift
: > "$out"
for n in ${K2_N:-100000000}; do
for t in "$@"; do
  if [ "$t" = default ]; then unset ACRO_B200_LIB; else export ACRO_B200_LIB=$PWD/acrobotics_b200/libacro_b200_$t.so; fi
  python tools/k2_quick.py --tag "$t" --n $n >> "$out" 2>> "${out%.jsonl}.err"
done
done
cat "$out"
