#!/bin/bash
# ncu --set full (with source) of the class kernels of one integral-direct G build (default workload)
mkdir -p gpurun_out
O=gpurun_out
WL=${1:-alkane40}; BASIS=${2:-6-31G}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eri_class_kernel -s 9 -c 9 -f -o $O/ncu_class_$WL python tools/time_config.py $WL $BASIS 0 1 > $O/ncu_class_$WL.log 2>&1; echo "rc=$?"
tail -3 $O/ncu_class_$WL.log
ls -la $O/*.ncu-rep
