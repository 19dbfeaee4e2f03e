#!/bin/bash
# round 2, GPU call 7: tuned defaults (early contraction for pp / d bras, r01 order for ss / ps bras, 80-register light classes,
# three resident blocks for the L >= 4 d classes): parity suite, register-budget A/B, executed-flop recount, ncu --set full, bench.
mkdir -p gpurun_out
O=gpurun_out
echo "== gpu tests (default build)"; timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_call7.log 2>&1; echo "rc=$?"; tail -3 $O/tests_call7.log
bash tools/link_variants.sh > $O/link.log 2>&1
: > $O/variants.log
for lib in default $(ls build/var/libfock_*.so 2>/dev/null); do
  if [ "$lib" = default ]; then unset FOCK_B200_LIB; else export FOCK_B200_LIB=$PWD/$lib; fi
  for cfg in "alkane40 6-31G 0 1" "w16 6-31G 0 0" "w32 cc-pVDZ 0 1" "w32 6-31G** 0 1"; do
    TAG=$(basename $lib) timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/variants.log
  done
done
unset FOCK_B200_LIB
cat $O/variants.log
FL=gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum
echo "== executed flops, w32cc"; REPS=0 timeout 900 ncu --metrics $FL --clock-control none -k regex:eri_class -c 1200 --csv --log-file $O/flops_w32cc.csv python tools/time_config.py w32 cc-pVDZ 0 1 > $O/flops_w32cc.log 2>&1; echo "rc=$?"
echo "== executed flops, alkane40"; REPS=0 timeout 600 ncu --metrics $FL --clock-control none -k regex:eri_class -c 200 --csv --log-file $O/flops_alkane40.csv python tools/time_config.py alkane40 6-31G 0 1 > $O/flops_alkane40.log 2>&1; echo "rc=$?"
echo "== executed flops, w32 6-31G**"; REPS=0 timeout 900 ncu --metrics $FL --clock-control none -k regex:eri_class -c 1200 --csv --log-file $O/flops_w32.csv python tools/time_config.py w32 '6-31G**' 0 1 > $O/flops_w32.log 2>&1; echo "rc=$?"
NB="--kernel-name-base mangled"
cap() {  # name, mangled-name regex
  REPS=0 timeout 600 ncu --set full --clock-control none --import-source on $NB -k regex:"$2" -s 1 -c 1 -f -o $O/ncu7_$1 python tools/time_config.py w32 cc-pVDZ 0 1 > $O/ncu7_$1.log 2>&1; echo "$1 rc=$?"
  ncu -i $O/ncu7_$1.ncu-rep --page raw --csv > $O/ncu7_$1.raw.csv 2>/dev/null
  ncu -i $O/ncu7_$1.ncu-rep --page source --csv --print-source sass > $O/ncu7_$1.sass.csv 2>/dev/null
  rm -f $O/ncu7_$1.ncu-rep
}
cap psss 'eri_class_kernelILi1ELi0ELi0ELi0ELi0ELi0ELi3ELi0ELi1ENS_15BlockDirectSink'
cap ppps 'eri_class_kernelILi1ELi1ELi1ELi0ELi0ELi0ELi9ELi0ELi3ENS_15BlockDirectSink'
cap dpps 'eri_class_kernelILi2ELi1ELi1ELi0ELi2ELi0ELi18ELi0ELi1ENS_15BlockDirectSink'
echo "== bench"; timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d['parity'])
for c in d['configs']: print(c['config']['workload'][:28], c['config']['storage'], c['config']['mode'], round(c['ms_per_step'],3), round(c['e2e']['ms_per_step'],3), round(c['roofline_eri']['frac'],3), c['parity']['g_vs_oracle_max_abs_diff'])
PY
