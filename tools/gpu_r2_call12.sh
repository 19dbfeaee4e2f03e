#!/bin/bash
# round 2, GPU call 12: shared-memory staging of the bra records (light classes) -- parity suite, A/B, bench
mkdir -p gpurun_out
O=gpurun_out
echo "== gpu tests (default build)"; timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_call12.log 2>&1; echo "rc=$?"; tail -3 $O/tests_call12.log
bash tools/link_variants.sh > $O/link.log 2>&1
: > $O/variants.log
for lib in default $(ls build/var/libfock_*.so 2>/dev/null); do
  if [ "$lib" = default ]; then unset FOCK_B200_LIB; else export FOCK_B200_LIB=$PWD/$lib; fi
  for cfg in "alkane40 6-31G 0 1" "w16 6-31G 0 0" "w32 cc-pVDZ 0 1" "w32 6-31G** 0 1"; do
    TAG=$(basename $lib) timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/variants.log
  done
done
unset FOCK_B200_LIB
cat $O/variants.log
echo "== bench"; timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d['parity'])
for c in d['configs']: print(c['config']['workload'][:28], c['config']['storage'], c['config']['mode'], round(c['ms_per_step'],3), round(c['e2e']['ms_per_step'],3), round(c['roofline_eri']['frac'],3), c['parity']['g_vs_oracle_max_abs_diff'])
PY
