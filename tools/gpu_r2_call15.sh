#!/bin/bash
# round 2, GPU call 15: page-locked host matrices in fock_build_g (ABI v5), new tests, bench with the reference arm
mkdir -p gpurun_out
O=gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_call15.log 2>&1; echo "rc=$?"; tail -3 $O/tests_call15.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline'].get('frac_executed'), d['parity'])
for c in d['configs']: print(c['config']['workload'][:28], c['config']['storage'], c['config']['mode'], round(c['ms_per_step'],3), round(c['e2e']['ms_per_step'],3), round(c['roofline_eri']['frac'],3), c['parity']['g_vs_oracle_max_abs_diff'])
print(d['cpu_baseline'])
PY
