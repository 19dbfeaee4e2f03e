#!/bin/bash
# round 2, GPU call 29: (pp|ps) as three one-component slices at 128 registers (4 resident blocks) against one 27-accumulator launch -- A/B + parity
mkdir -p gpurun_out
O=gpurun_out
bash tools/link_variants.sh > $O/link.log 2>&1
: > $O/variants.log
for lib in default $(ls build/var/libfock_*.so 2>/dev/null); do
  if [ "$lib" = default ]; then unset FOCK_B200_LIB; else export FOCK_B200_LIB=$PWD/$lib; fi
  for cfg in "alkane40 6-31G 0 1" "w32 cc-pVDZ 0 1" "w32 6-31G** 0 1" "w16 6-31G 0 0"; do
    TAG=$(basename $lib) timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/variants.log
  done
done
cat $O/variants.log
FOCK_B200_LIB=$PWD/build/var/libfock_ppps1.so timeout 600 python -m pytest tests/test_gpu_golden_large.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
