#!/bin/bash
# round 2, GPU call 23: per-source-line profile of a d-class slice ((dp|ps), config 5) and of (ps|ss) C40H82
mkdir -p gpurun_out
O=gpurun_out
NB="--kernel-name-base mangled"
cap() {
  REPS=0 timeout 600 ncu --set full --clock-control none --import-source on $NB -k regex:"$3" -s 1 -c 1 -f -o $O/ncuS_$1 python tools/time_config.py $2 > $O/ncuS_$1.log 2>&1; echo "$1 rc=$?"
  ncu -i $O/ncuS_$1.ncu-rep --page source --csv --print-source cuda > $O/ncuS_$1.cuda.csv 2>/dev/null
  rm -f $O/ncuS_$1.ncu-rep
}
cap w32_dpps "w32 6-31G** 0 1" 'eri_class_kernelILi2ELi1ELi1ELi0ELi2ELi0ELi18ELi0ELi1ENS_15BlockDirectSink'
cap alkane40_psss "alkane40 6-31G 0 1" 'eri_class_kernelILi1ELi0ELi0ELi0ELi0ELi0ELi3ELi0ELi1ENS_15BlockDirectSink'
ls -la $O/ncuS_*
