for v in "" _c5 _c6 _d12 _d28; do
  for g in 64; do
    echo "== lib$v grid $g"
    ACRO_B200_LIB=$PWD/acrobotics_b200/libacro_b200$v.so ACRO_K2_GRID=$g python tools/kernel_bench.py --only k2 --n 16000000 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    if d['kernel'] in ('fk_cc_f64[scene16]','fk_cc_f64[self_only]'): print(d['kernel'], d['ms_best'], '%.3g'%d['units_per_s'])
"
  done
done
for g in 32 48 96 128; do
  echo "== lib grid $g"
  ACRO_K2_GRID=$g python tools/kernel_bench.py --only k2 --n 16000000 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    if d['kernel'] in ('fk_cc_f64[scene16]',): print(d['kernel'], d['ms_best'], '%.3g'%d['units_per_s'])
"
done
