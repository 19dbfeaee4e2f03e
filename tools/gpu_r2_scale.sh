#!/bin/bash
# bench at N ranks only (run with gpurun --gpus N): one point of the 1/2/4/8 scaling curve
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('n_gpus','ms_per_step')}, 'e2e', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), d['parity']['g_vs_oracle_max_abs_diff'])
for c in d['configs']: print('   ', c['config']['workload'][:28], c['config']['storage'], c['config']['mode'], c['config']['screening'][:3], round(c['ms_per_step'],3), round(c['e2e']['ms_per_step'],3), c['parity']['g_vs_oracle_max_abs_diff'], (c.get('scf_shaped') or {}).get('ms_per_g_build'))
PY
