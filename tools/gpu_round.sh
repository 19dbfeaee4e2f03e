#!/bin/bash
# One gpurun call: new-feature GPU tests, default bench, config-5 timing, launch list, full GPU suite, knob sweep.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
echo "== new tests" ; timeout 600 python -m pytest tests/test_orbital.py tests/test_ccpvdz.py tests/test_hostcpp.py -m gpu -q > $O/tests_new.log 2>&1; echo "rc=$?" >> $O/tests_new.log; tail -15 $O/tests_new.log
echo "== bench" ; timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; tail -c 600 $O/bench_default.json
echo "== config5 cc" ; TAG=cc timeout 300 python tools/time_config.py w32 cc-pVDZ 0 1 > $O/time_w32cc.log 2>&1; tail -3 $O/time_w32cc.log
echo "== launch list" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu full voxel" ; timeout 300 ncu --set full --clock-control none --import-source on -k regex:orbital_voxel -c 1 -f -o $O/ncu_voxel python tools/voxel_once.py > $O/ncu_voxel.log 2>&1; echo "rc=$?"
echo "== full gpu suite" ; timeout 900 python -m pytest tests -m gpu -x -q --ignore=tests/test_orbital.py --ignore=tests/test_ccpvdz.py --ignore=tests/test_hostcpp.py --durations=8 > $O/tests_full.log 2>&1; echo "rc=$?" >> $O/tests_full.log; tail -14 $O/tests_full.log
echo "== knob sweep"
for bps in 4 6 12 16; do TAG=bps$bps FOCK_BLOCKS_PER_SM=$bps timeout 120 python tools/time_config.py alkane40 6-31G 0 1 2>&1 | head -1; done > $O/knobs.log 2>&1
for s in 1 2; do TAG=streams$s FOCK_STREAMS=$s timeout 120 python tools/time_config.py alkane40 6-31G 0 1 2>&1 | head -1; done >> $O/knobs.log 2>&1
cat $O/knobs.log
