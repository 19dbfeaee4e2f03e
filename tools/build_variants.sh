#!/bin/bash
# Build experimental variants of class_launch.cu (the s/p class kernels) into build/var/libfock_<name>.so.
# usage: tools/build_variants.sh "name:-DFLAG=..." ...   (the other objects come from the last full build)
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/build/var
cd $ROOT/fock_rs_b200/csrc
for v in "$@"; do
  n=${v%%:*}; f=${v#*:}
  ( /usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC -Xptxas -v $f -c class_launch.cu -o $ROOT/build/var/class_launch_$n.o 2> $ROOT/build/var/$n.ptxas.log
    objs=$(ls $ROOT/build/obj/*.o | grep -v class_launch.o)
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o $ROOT/build/var/libfock_$n.so $objs $ROOT/build/var/class_launch_$n.o -lcudart ) > /dev/null 2>&1 &
done
wait
for v in "$@"; do
  n=${v%%:*}; echo "== $n"
  grep -h "eri_class_kernelILi1ELi0ELi0ELi0ELi0E.*BlockDirect\|eri_class_kernelILi0ELi0ELi0ELi0ELi0E.*BlockDirect\|eri_class_kernelILi1ELi0ELi1ELi0ELi0E.*BlockDirect\|eri_class_kernelILi1ELi1ELi0ELi0ELi0E.*BlockDirect" -A3 $ROOT/build/var/$n.ptxas.log | grep -E "registers|spill"
done
ls $ROOT/build/var/*.so
