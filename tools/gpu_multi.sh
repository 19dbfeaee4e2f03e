#!/bin/bash
# N-GPU check (run with gpurun --gpus N): sharded builds vs the oracle, then the bench line at N ranks.
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > $O/multi_check.log 2>&1; echo "check rc=$?"; tail -4 $O/multi_check.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --no-cpu-baseline > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?"; tail -c 1500 $O/bench_n$N.json | head -c 1500
