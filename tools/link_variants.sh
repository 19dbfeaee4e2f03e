#!/bin/bash
# Link build/var/libfock_<name>.so for every variant directory build/var/<name>/ (objects made by tools/build_variants.sh): the
# variant's objects replace the same-named ones of the last full build (build/obj).  Run on the GPU box: shipping objects
# is several times smaller than shipping every variant's 88 MB library.
ROOT=$(cd "$(dirname "$0")/.." && pwd)
for d in $ROOT/build/var/*/; do
  [ -d "$d" ] || continue
  n=$(basename $d)
  objs=""
  for o in $ROOT/build/obj/*.o; do
    b=$(basename $o)
    if [ -f "$d/$b" ]; then objs="$objs $d/$b"; else objs="$objs $o"; fi
  done
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o $ROOT/build/var/libfock_$n.so $objs -lcudart -ldl || echo "link $n failed"
done
ls -la $ROOT/build/var/*.so
