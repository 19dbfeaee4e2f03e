#!/bin/bash
# round 2, GPU call 4 (2 GPUs): the collective behind the C ABI from compiled host code, the 2-rank check against the oracle,
# bench at N=2 (both arms), then single-GPU ncu --set full captures of the headline workload's kernels.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $O/smi.txt 2>&1
echo "== hostcpp multi 2"; timeout 600 python -m pytest tests/test_hostcpp.py -m gpu -q -s > $O/hostcpp.log 2>&1; echo "rc=$?"; tail -4 $O/hostcpp.log
build/hostcpp_main multi 2 2>&1 | tail -2
echo "== multi_gpu_check 2 ranks"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > $O/multi_check_2.log 2>&1; echo "rc=$?"; grep -v "^W\|^\[W" $O/multi_check_2.log | tail -45
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d['parity'])
for c in d['configs']: print(c['config']['workload'][:28], c['config']['storage'], c['config']['mode'], round(c['ms_per_step'],3), round(c['e2e']['ms_per_step'],3), c['parity']['g_vs_oracle_max_abs_diff'])
PY
echo "== bench N=1 (for the same box)"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline'].get('frac_executed'), d['parity'])
for c in d['configs']: print(c['config']['workload'][:28], c['config']['storage'], c['config']['mode'], round(c['ms_per_step'],3), round(c['e2e']['ms_per_step'],3), c['parity']['g_vs_oracle_max_abs_diff'])
PY
echo "== reference arm under torchrun N=2 (threads must not collapse to 1)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --ref-budget 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['cpu_baseline']['cores'], d['ms_per_step'], d['sample_fraction'])"
NB="--kernel-name-base demangled"
echo "== ncu full (ps|ss) w32cc"; REPS=0 timeout 600 ncu --set full --clock-control none --import-source on $NB -k regex:"eri_class_kernel<1, 0, 0, 0, " -s 1 -c 1 -f -o $O/ncu_w32cc_psss python tools/time_config.py w32 cc-pVDZ 0 1 > $O/ncu_w32cc_psss.log 2>&1; echo "rc=$?"
echo "== ncu full (dp|ps) slice w32cc"; REPS=0 timeout 600 ncu --set full --clock-control none --import-source on $NB -k regex:"eri_class_kernel<2, 1, 1, 0, 2, 0, 18, 0, 1" -s 1 -c 1 -f -o $O/ncu_w32cc_dpps python tools/time_config.py w32 cc-pVDZ 0 1 > $O/ncu_w32cc_dpps.log 2>&1; echo "rc=$?"
echo "== ncu full (dp|pp) slice w32cc"; REPS=0 timeout 600 ncu --set full --clock-control none --import-source on $NB -k regex:"eri_class_kernel<2, 1, 1, 1, 2, 0, 18, 0, 1" -s 1 -c 1 -f -o $O/ncu_w32cc_dppp python tools/time_config.py w32 cc-pVDZ 0 1 > $O/ncu_w32cc_dppp.log 2>&1; echo "rc=$?"
echo "== ncu full (pp|ps) w32cc"; REPS=0 timeout 600 ncu --set full --clock-control none --import-source on $NB -k regex:"eri_class_kernel<1, 1, 1, 0, " -s 1 -c 1 -f -o $O/ncu_w32cc_ppps python tools/time_config.py w32 cc-pVDZ 0 1 > $O/ncu_w32cc_ppps.log 2>&1; echo "rc=$?"
ls -la $O/*.ncu-rep
