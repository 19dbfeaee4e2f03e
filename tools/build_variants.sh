#!/bin/bash
# Build experimental variants of chosen translation units into build/var/<name>/<tu>.o; tools/link_variants.sh turns every
# such directory into build/var/libfock_<name>.so (the other objects come from the last full build).
# usage: tools/build_variants.sh "name:-DFLAG=... [-DFLAG2]" ...                (class_launch.cu = the s/p class kernels)
#        TUS="dclass_dp_block dclass_ds_block" tools/build_variants.sh "h3:-DFOCK_MINB_HEAVY=3"
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TUS=${TUS:-class_launch}
cd $ROOT/fock_rs_b200/csrc
for v in "$@"; do
  n=${v%%:*}; f=${v#*:}
  mkdir -p $ROOT/build/var/$n
  for tu in $TUS; do
    ( /usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC -Xptxas -v $f -c $tu.cu -o $ROOT/build/var/$n/$tu.o 2> $ROOT/build/var/$n/$tu.ptxas.log ) &
  done
done
wait
for v in "$@"; do
  n=${v%%:*}; echo "== $n"; ls -la $ROOT/build/var/$n/*.o | awk '{print $5, $9}'
  grep -h "eri_class_kernelILi1ELi0ELi0ELi0ELi0E.*BlockDirect\|eri_class_kernelILi0ELi0ELi0ELi0ELi0E.*BlockDirect\|eri_class_kernelILi1ELi0ELi1ELi0ELi0E.*BlockDirect\|eri_class_kernelILi1ELi1ELi0ELi0ELi0E.*BlockDirect" -A3 $ROOT/build/var/$n/*.ptxas.log 2>/dev/null | grep -E "registers|spill"
done
