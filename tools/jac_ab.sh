mkdir -p gpurun_out
: > gpurun_out/jac_ab.jsonl
for t in default jacA jacB jacF jacC jacD jacE; do
  if [ "$t" = default ]; then unset ACRO_B200_LIB; else export ACRO_B200_LIB=$PWD/acrobotics_b200/libacro_b200_$t.so; fi
  echo "{\"tag\": \"$t\"}" >> gpurun_out/jac_ab.jsonl
  python tools/kernel_bench.py --n 4000000 --reps 7 --only jac >> gpurun_out/jac_ab.jsonl 2>> gpurun_out/jac_ab.err
done
export ACRO_B200_LIB=$PWD/acrobotics_b200/libacro_b200_jacC.so
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "jac" 2>&1 | tail -3
cat gpurun_out/jac_ab.jsonl
