#!/bin/bash
# round 2, GPU call 5: ncu --set full (with source) of the headline workload's dominant kernels, one launch each.
mkdir -p gpurun_out
O=gpurun_out
NB="--kernel-name-base mangled"
cap() {  # name, mangled-name regex
  REPS=0 timeout 600 ncu --set full --clock-control none --import-source on $NB -k regex:"$2" -s 1 -c 1 -f -o $O/ncu_w32cc_$1 python tools/time_config.py w32 cc-pVDZ 0 1 > $O/ncu_w32cc_$1.log 2>&1; echo "$1 rc=$?"
}
cap psss 'eri_class_kernelILi1ELi0ELi0ELi0ELi0ELi0ELi3ELi0ELi1ENS_15BlockDirectSink'
cap ssss 'eri_class_kernelILi0ELi0ELi0ELi0ELi0ELi0ELi1ELi0ELi1ENS_15BlockDirectSink'
cap ppps 'eri_class_kernelILi1ELi1ELi1ELi0ELi0ELi0ELi9ELi0ELi3ENS_15BlockDirectSink'
cap dpps 'eri_class_kernelILi2ELi1ELi1ELi0ELi2ELi0ELi18ELi0ELi1ENS_15BlockDirectSink'
cap dppp 'eri_class_kernelILi2ELi1ELi1ELi1ELi2ELi0ELi18ELi0ELi1ENS_15BlockDirectSink'
ls -la $O/*.ncu-rep
for n in psss ssss ppps dpps dppp; do
  ncu -i $O/ncu_w32cc_$n.ncu-rep --page raw --csv > $O/ncu_w32cc_$n.raw.csv 2>/dev/null
  ncu -i $O/ncu_w32cc_$n.ncu-rep --page source --csv --print-source sass > $O/ncu_w32cc_$n.sass.csv 2>/dev/null
done
rm -f $O/*.ncu-rep   # the raw / source pages above are what gets summarised (gpurun_out/ is capped at 64 MiB)
ls -la $O/*.csv | head -20
echo "== in-core A/B: dynamic draw / job order"; bash tools/link_variants.sh > $O/link.log 2>&1
: > $O/variants.log
for lib in default build/var/libfock_static.so; do
  if [ "$lib" = default ]; then unset FOCK_B200_LIB; else export FOCK_B200_LIB=$PWD/$lib; fi
  for cfg in "w16 6-31G 0 0" "benzene 6-31G 0 0" "w16 6-31G 0 1" "alkane40 6-31G 0 1"; do
    TAG=$(basename $lib) timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/variants.log
    TAG=$(basename $lib)-nolpt FOCK_NO_LPT=1 timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/variants.log
  done
done
unset FOCK_B200_LIB
cat $O/variants.log
echo "== bench (default) on this box"; timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; tail -c 400 $O/bench_default.json
