#!/bin/bash
# round 2, GPU call 1: new parity tests (full-size goldens, modes, C++ host), A/B of the kernel variants, sanitizer.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/smi.txt 2>&1
echo "== new tests"; timeout 1200 python -m pytest tests/test_gpu_golden_large.py tests/test_gpu_modes.py tests/test_hostcpp.py -m gpu -q -s --durations=10 > $O/tests_new.log 2>&1; echo "rc=$?" >> $O/tests_new.log; tail -25 $O/tests_new.log
echo "== variants"; bash tools/link_variants.sh > $O/link.log 2>&1; CHECK_LIB=build/var/libfock_all.so bash tools/gpu_variants.sh > $O/variants_run.log 2>&1; tail -50 $O/variants_run.log
echo "== sanitizer"
for tool in memcheck racecheck; do
  for c in sp d; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py $c > $O/sanitizer_${tool}_$c.log 2>&1; echo "$tool $c rc=$?"; tail -4 $O/sanitizer_${tool}_$c.log
  done
done
