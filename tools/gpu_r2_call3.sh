#!/bin/bash
# round 2, GPU call 3: ncu evidence for the headline workload (config 5, cc-pVDZ-shaped): executed-flop counts of every class kernel of
# one G build, --set full of the dominant s/p kernel and of two d-class slice kernels; A/B of the d-class register bound.
mkdir -p gpurun_out
O=gpurun_out
FL=gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum
echo "== executed flops, w32cc"; REPS=0 timeout 900 ncu --metrics $FL --clock-control none -k regex:eri_class -c 1200 --csv --log-file $O/flops_w32cc.csv python tools/time_config.py w32 cc-pVDZ 0 1 > $O/flops_w32cc.log 2>&1; echo "rc=$?"
echo "== executed flops, alkane40"; REPS=0 timeout 600 ncu --metrics $FL --clock-control none -k regex:eri_class -c 200 --csv --log-file $O/flops_alkane40.csv python tools/time_config.py alkane40 6-31G 0 1 > $O/flops_alkane40.log 2>&1; echo "rc=$?"
echo "== ncu full (ps|ss) w32cc"; REPS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"eri_class_kernel<1, 0, 0, 0" -s 1 -c 1 -f -o $O/ncu_w32cc_psss python tools/time_config.py w32 cc-pVDZ 0 1 > $O/ncu_w32cc_psss.log 2>&1; echo "rc=$?"
echo "== ncu full (dp|ps) slice w32cc"; REPS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"eri_class_kernel<2, 1, 1, 0, 2, 0, 18, 0, 1" -s 1 -c 1 -f -o $O/ncu_w32cc_dpps python tools/time_config.py w32 cc-pVDZ 0 1 > $O/ncu_w32cc_dpps.log 2>&1; echo "rc=$?"
echo "== ncu full (dp|pp) slice w32cc"; REPS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"eri_class_kernel<2, 1, 1, 1, 2, 0, 18, 0, 1" -s 1 -c 1 -f -o $O/ncu_w32cc_dppp python tools/time_config.py w32 cc-pVDZ 0 1 > $O/ncu_w32cc_dppp.log 2>&1; echo "rc=$?"
echo "== ncu full (pp|ps) w32cc"; REPS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"eri_class_kernel<1, 1, 1, 0" -s 1 -c 1 -f -o $O/ncu_w32cc_ppps python tools/time_config.py w32 cc-pVDZ 0 1 > $O/ncu_w32cc_ppps.log 2>&1; echo "rc=$?"
ls -la $O/*.ncu-rep
echo "== variants"; bash tools/link_variants.sh > $O/link.log 2>&1
: > $O/variants.log
for lib in default $(ls build/var/libfock_*.so 2>/dev/null); do
  if [ "$lib" = default ]; then unset FOCK_B200_LIB; else export FOCK_B200_LIB=$PWD/$lib; fi
  for cfg in "alkane40 6-31G 0 1" "w32 cc-pVDZ 0 1" "w32 6-31G** 0 1"; do
    TAG=$(basename $lib) timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/variants.log
  done
done
unset FOCK_B200_LIB
cat $O/variants.log
echo "== gpu tests (new default build)"; timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_full.log 2>&1; echo "rc=$?" >> $O/tests_full.log; tail -5 $O/tests_full.log
