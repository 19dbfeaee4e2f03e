#!/bin/bash
# Final validation of the committed state on one B200: full GPU suite, smoke, default bench, reference arm, launch list.
mkdir -p gpurun_out
O=gpurun_out
echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -x -q > $O/final_tests.log 2>&1; echo "rc=$?" >> $O/final_tests.log; tail -4 $O/final_tests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "rc=$?" >> $O/final_smoke.log; tail -2 $O/final_smoke.log
echo "== bench"; timeout 600 python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "rc=$?"; head -c 400 $O/final_bench.json; echo
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/final_bench_reference.json 2> $O/final_bench_reference.err; echo "rc=$?"; head -c 300 $O/final_bench_reference.json; echo
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/final_bench_under_ncu.log 2>&1; echo "rc=$?"
