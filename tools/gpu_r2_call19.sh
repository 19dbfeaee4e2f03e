#!/bin/bash
# round 2, GPU call 19: independent DFMA probe, bench with the unscreened like-for-like configuration, compiled C++ host test
mkdir -p gpurun_out
O=gpurun_out
build/dfma_probe > $O/dfma_probe.json; cat $O/dfma_probe.json
echo "== hostcpp"; timeout 600 python -m pytest tests/test_hostcpp.py -m gpu -q 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline'].get('frac_executed'))
for c in d['configs']: print(c['config']['workload'][:28], c['config']['storage'], c['config']['mode'], c['config']['screening'][:12], round(c['ms_per_step'],3), round(c['e2e']['ms_per_step'],3), round(c['roofline_eri']['frac'],3), c['parity']['g_vs_oracle_max_abs_diff'], c['stats']['class_shell_quartets_evaluated'], c['stats']['shell_quartets_unique'])
PY
