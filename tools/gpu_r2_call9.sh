#!/bin/bash
# round 2, GPU call 9: spatial (Morton) pair order + CSR of the surviving quartets -- parity suite, A/B against the r01 order
# (FOCK_PAIR_ORDER=diag, same build), near-field lane statistics through ncu source counters, bench.
mkdir -p gpurun_out
O=gpurun_out
echo "== gpu tests (default build)"; timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_call9.log 2>&1; echo "rc=$?"; tail -5 $O/tests_call9.log
: > $O/variants.log
for ord in morton diag; do
  for cfg in "alkane40 6-31G 0 1" "w16 6-31G 0 0" "w16 6-31G 1 1" "w32 cc-pVDZ 0 1" "w32 6-31G** 0 1" "benzene 6-31G 0 0"; do
    FOCK_PAIR_ORDER=$ord TAG=$ord timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/variants.log
  done
done
cat $O/variants.log
for cfg in "w32 cc-pVDZ 0 1:w32cc" "alkane40 6-31G 0 1:alkane40"; do
  a=${cfg%%:*}; n=${cfg##*:}
  REPS=0 timeout 900 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -k regex:eri_class --csv --log-file $O/launches9_$n.csv python tools/time_config.py $a > $O/launches9_$n.log 2>&1; echo "$n rc=$?"
done
echo "== bench"; timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d['parity'])
for c in d['configs']: print(c['config']['workload'][:28], c['config']['storage'], c['config']['mode'], round(c['ms_per_step'],3), round(c['e2e']['ms_per_step'],3), round(c['roofline_eri']['frac'],3), c['parity']['g_vs_oracle_max_abs_diff'])
PY
