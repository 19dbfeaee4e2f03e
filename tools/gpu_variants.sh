#!/bin/bash
# A/B timing of experimental library builds (build/var/libfock_*.so) against the in-tree one.
mkdir -p gpurun_out
O=gpurun_out/variants.log
: > $O
for lib in default $(ls build/var/libfock_*.so 2>/dev/null); do
  if [ "$lib" = default ]; then unset FOCK_B200_LIB; else export FOCK_B200_LIB=$PWD/$lib; fi
  for cfg in "alkane40 6-31G 0 1" "w16 6-31G 0 0" "w32 6-31G** 0 1" "w32 cc-pVDZ 0 1"; do
    TAG=$(basename $lib) timeout 120 python tools/time_config.py $cfg 2>&1 | head -1 >> $O
  done
done
cat $O
