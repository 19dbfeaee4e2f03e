#!/bin/bash
# N-GPU evidence (run with gpurun --gpus N): N-rank check against the oracle / golden matrices, bench at N ranks and at 1
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/smi_n$N.txt 2>&1
echo "== multi_gpu_check $N ranks"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > $O/multi_check_$N.log 2>&1; echo "rc=$?"; grep -v "^W\|^\[W\|^$" $O/multi_check_$N.log | tail -30
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "rc=$?"
echo "== bench N=1 (same box)"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n1_samebox.json 2> $O/bench_n1_samebox.err; echo "rc=$?"
python - <<PY
import json
for f in ('gpurun_out/bench_n$N.json','gpurun_out/bench_n1_samebox.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, {k:d[k] for k in ('n_gpus','ms_per_step')}, 'e2e', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), d['parity']['g_vs_oracle_max_abs_diff'])
    for c in d['configs']: print('   ', c['config']['workload'][:28], c['config']['storage'], c['config']['mode'], round(c['ms_per_step'],3), round(c['e2e']['ms_per_step'],3), c['parity']['g_vs_oracle_max_abs_diff'])
PY
