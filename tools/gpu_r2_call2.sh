#!/bin/bash
# round 2, GPU call 2: parity suites after the plan-table fix, A/B of the tile scheduler variants, near-field statistics,
# sanitizer, the new bench line (both arms), launch list of one bench step.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/smi.txt 2>&1
echo "== new tests"; timeout 1500 python -m pytest tests/test_gpu_golden_large.py tests/test_gpu_modes.py tests/test_hostcpp.py tests/test_golden.py -m gpu -q -s --durations=10 > $O/tests_new.log 2>&1; echo "rc=$?" >> $O/tests_new.log; grep "max|dG|" $O/tests_new.log; tail -12 $O/tests_new.log
echo "== variants"; bash tools/link_variants.sh > $O/link.log 2>&1
: > $O/variants.log
for lib in default build/var/libfock_dyn.so build/var/libfock_nott.so; do
  if [ "$lib" = default ]; then unset FOCK_B200_LIB; else export FOCK_B200_LIB=$PWD/$lib; fi
  for cfg in "alkane40 6-31G 0 1" "w16 6-31G 1 1" "w32 cc-pVDZ 0 1" "w32 6-31G** 0 1"; do
    TAG=$(basename $lib) timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/variants.log
    TAG=$(basename $lib)-nolpt FOCK_NO_LPT=1 timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/variants.log
  done
done
unset FOCK_B200_LIB
cat $O/variants.log
echo "== near stats"
for cfg in "alkane40 6-31G 0 1" "w16 6-31G 1 1" "w32 cc-pVDZ 0 1"; do
  echo "-- $cfg" >> $O/near_stats.log
  REPS=0 FOCK_B200_LIB=$PWD/build/var/libfock_stats.so timeout 300 python tools/time_config.py $cfg 2>&1 | grep NEAR_STATS | tail -6 >> $O/near_stats.log
done
cat $O/near_stats.log
echo "== parity of the dyn variant"; FOCK_B200_LIB=$PWD/build/var/libfock_dyn.so timeout 900 python -m pytest tests/test_golden.py tests/test_gpu_parity.py tests/test_gpu_golden_large.py -m gpu -x -q 2>&1 | tail -3
echo "== sanitizer"
for c in sp d; do
  timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_case.py $c > $O/sanitizer_memcheck_$c.log 2>&1; echo "memcheck $c rc=$?"; tail -3 $O/sanitizer_memcheck_$c.log
done
echo "== bench"; timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; tail -c 1500 $O/bench_default.json; tail -5 $O/bench_default.err
echo "== bench reference arm (short)"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?"; tail -c 1200 $O/bench_reference.json
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > $O/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_golden_large.py --deselect tests/test_gpu_modes.py --durations=8 > $O/tests_full.log 2>&1; echo "rc=$?" >> $O/tests_full.log; tail -14 $O/tests_full.log
