#!/bin/bash
# round 2, GPU call 10: why the spatial order does not pay -- ncu --set full of the (ps|ss) launch of C40H82 in both orders
mkdir -p gpurun_out
O=gpurun_out
NB="--kernel-name-base mangled"
for ord in morton diag; do
  FOCK_PAIR_ORDER=$ord REPS=0 timeout 600 ncu --set full --clock-control none --import-source on $NB -k regex:'eri_class_kernelILi1ELi0ELi0ELi0ELi0ELi0ELi3ELi0ELi1ENS_15BlockDirectSink' -s 1 -c 1 -f -o $O/ncu10_$ord python tools/time_config.py alkane40 6-31G 0 1 > $O/ncu10_$ord.log 2>&1; echo "$ord rc=$?"
  ncu -i $O/ncu10_$ord.ncu-rep --page raw --csv > $O/ncu10_$ord.raw.csv 2>/dev/null
  ncu -i $O/ncu10_$ord.ncu-rep --page source --csv --print-source sass > $O/ncu10_$ord.sass.csv 2>/dev/null
  rm -f $O/ncu10_$ord.ncu-rep
done
