#!/bin/bash
# Scaling points on one multi-GPU box (run with gpurun --gpus 8): bench line at 8 and 4 ranks, headline workload only.
mkdir -p gpurun_out
O=gpurun_out
for N in 8 4; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --no-cpu-baseline --no-other-configs > $O/scale_n$N.json 2> $O/scale_n$N.err; echo "N=$N rc=$?"; head -c 330 $O/scale_n$N.json; echo
done
