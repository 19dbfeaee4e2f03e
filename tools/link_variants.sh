#!/bin/bash
# Link build/var/libfock_<name>.so from build/var/class_launch_<name>.o + the other objects of the last full build
# (run on the GPU box: shipping the objects once is 5x smaller than shipping every variant's 88 MB library).
ROOT=$(cd "$(dirname "$0")/.." && pwd)
objs=$(ls $ROOT/build/obj/*.o | grep -v class_launch.o)
for o in $ROOT/build/var/class_launch_*.o; do
  n=$(basename $o .o); n=${n#class_launch_}
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o $ROOT/build/var/libfock_$n.so $objs $o -lcudart -ldl || echo "link $n failed"
done
ls -la $ROOT/build/var/*.so
