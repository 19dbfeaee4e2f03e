#!/bin/bash
# round 2, final evidence of the committed state on one B200: GPU suite, smoke, bench (+ reference arm), executed-flop recount,
# ncu --set full of the dominant kernels, launch list of the bench command.
mkdir -p gpurun_out
O=gpurun_out
echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -x -q > $O/final_tests.log 2>&1; echo "rc=$?" >> $O/final_tests.log; tail -4 $O/final_tests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "rc=$?" >> $O/final_smoke.log; tail -2 $O/final_smoke.log
echo "== bench"; timeout 900 python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], d['roofline'].get('frac_executed'), d['parity'], d['clocks'])
for c in d['configs']: print(c['config']['workload'][:28], c['config']['storage'], c['config']['mode'], round(c['ms_per_step'],3), round(c['e2e']['ms_per_step'],3), round(c['roofline_eri']['frac'],3), c['parity']['g_vs_oracle_max_abs_diff'], c.get('scf_shaped'))
print(d['cpu_baseline']['value'], d['cpu_baseline'].get('s_per_step'))
PY
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/final_bench_reference.json 2> $O/final_bench_reference.err; echo "rc=$?"; head -c 300 $O/final_bench_reference.json; echo
FL=gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum
echo "== executed flops, w32cc"; REPS=0 timeout 900 ncu --metrics $FL --clock-control none -k regex:eri_class -c 1200 --csv --log-file $O/flops_w32cc.csv python tools/time_config.py w32 cc-pVDZ 0 1 > $O/flops_w32cc.log 2>&1; echo "rc=$?"
echo "== executed flops, alkane40"; REPS=0 timeout 600 ncu --metrics $FL --clock-control none -k regex:eri_class -c 200 --csv --log-file $O/flops_alkane40.csv python tools/time_config.py alkane40 6-31G 0 1 > $O/flops_alkane40.log 2>&1; echo "rc=$?"
echo "== executed flops, w32 6-31G**"; REPS=0 timeout 900 ncu --metrics $FL --clock-control none -k regex:eri_class -c 1200 --csv --log-file $O/flops_w32.csv python tools/time_config.py w32 '6-31G**' 0 1 > $O/flops_w32.log 2>&1; echo "rc=$?"
NB="--kernel-name-base mangled"
cap() {  # name, molecule args, mangled-name regex
  REPS=0 timeout 600 ncu --set full --clock-control none --import-source on $NB -k regex:"$3" -s 1 -c 1 -f -o $O/ncuF_$1 python tools/time_config.py $2 > $O/ncuF_$1.log 2>&1; echo "$1 rc=$?"
  ncu -i $O/ncuF_$1.ncu-rep --page raw --csv > $O/ncuF_$1.raw.csv 2>/dev/null
  ncu -i $O/ncuF_$1.ncu-rep --page source --csv --print-source sass > $O/ncuF_$1.sass.csv 2>/dev/null
  rm -f $O/ncuF_$1.ncu-rep
}
cap w32cc_psss "w32 cc-pVDZ 0 1" 'eri_class_kernelILi1ELi0ELi0ELi0ELi0ELi0ELi3ELi0ELi1ENS_15BlockDirectSink'
cap w32cc_ppps "w32 cc-pVDZ 0 1" 'eri_class_kernelILi1ELi1ELi1ELi0ELi0ELi0ELi9ELi0ELi3ENS_15BlockDirectSink'
cap w32cc_dpps "w32 cc-pVDZ 0 1" 'eri_class_kernelILi2ELi1ELi1ELi0ELi2ELi0ELi18ELi0ELi1ENS_15BlockDirectSink'
cap alkane40_psss "alkane40 6-31G 0 1" 'eri_class_kernelILi1ELi0ELi0ELi0ELi0ELi0ELi3ELi0ELi1ENS_15BlockDirectSink'
cap w16_symv "w16 6-31G 0 0" 'symv_packed_kernel'
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/final_bench_under_ncu.log 2>&1; echo "rc=$?"
ls -la $O | tail -30
