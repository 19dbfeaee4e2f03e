#!/bin/bash
# round 2, GPU call 13: one rank's share of an 8-GPU (and 4-, 2-GPU) run emulated on one GPU: what the tails / fixed costs take
mkdir -p gpurun_out
O=gpurun_out
: > $O/rank_emu.log
for cfg in "w32 cc-pVDZ 0 1" "alkane40 6-31G 0 1" "w32 6-31G** 0 1"; do
  timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/rank_emu.log
  for w in 2 4 8; do
    for r in 0 $((w-1)); do
      RANK_EMU=$r WORLD_EMU=$w timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/rank_emu.log
    done
  done
done
cat $O/rank_emu.log
echo "== per-launch list, rank 3 of 8, w32cc"
RANK_EMU=3 WORLD_EMU=8 REPS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches13_w32cc_r3of8.csv python tools/time_config.py w32 cc-pVDZ 0 1 > $O/launches13.log 2>&1; echo "rc=$?"
