#!/bin/bash
# A/B timing of experimental library builds (build/var/libfock_*.so, made by tools/build_variants.sh) against the in-tree one;
# CHECK_LIB=build/var/libfock_<name>.so additionally runs the golden + parity GPU tests against that build.
# usage (on the GPU box):  gpurun -- 'CHECK_LIB=build/var/libfock_nq.so bash tools/gpu_variants.sh'
mkdir -p gpurun_out
O=gpurun_out/variants.log
: > $O
for lib in default $(ls build/var/libfock_*.so 2>/dev/null); do
  if [ "$lib" = default ]; then unset FOCK_B200_LIB; else export FOCK_B200_LIB=$PWD/$lib; fi
  for cfg in "alkane40 6-31G 0 1" "w16 6-31G 0 0" "w16 6-31G 1 1" "w32 cc-pVDZ 0 1"; do
    TAG=$(basename $lib) timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O   # min / median of 6 G builds inside
  done
done
cat $O
if [ -n "$CHECK_LIB" ]; then
  FOCK_B200_LIB=$PWD/$CHECK_LIB timeout 600 python -m pytest tests/test_golden.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
fi
