#!/bin/bash
# round 2, GPU call 22: unroll factor of the bra-primitive loop of the light classes (A/B)
mkdir -p gpurun_out
O=gpurun_out
bash tools/link_variants.sh > $O/link.log 2>&1
: > $O/variants.log
for lib in default $(ls build/var/libfock_*.so 2>/dev/null); do
  if [ "$lib" = default ]; then unset FOCK_B200_LIB; else export FOCK_B200_LIB=$PWD/$lib; fi
  for cfg in "alkane40 6-31G 0 1" "w32 cc-pVDZ 0 1" "w16 6-31G 1 1"; do
    TAG=$(basename $lib) timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/variants.log
  done
done
unset FOCK_B200_LIB
cat $O/variants.log
