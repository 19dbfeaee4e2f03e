#!/bin/bash
# round 2, GPU call 14: slice launches dealt over 16 streams, per-rank K-split limit -- parity, 1/8-share emulation A/B
mkdir -p gpurun_out
O=gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_call14.log 2>&1; echo "rc=$?"; tail -3 $O/tests_call14.log
: > $O/rank_emu.log
for var in "default" "nospread:FOCK_NO_SLICE_SPREAD=1 FOCK_STREAMS=4" "s8:FOCK_STREAMS=8"; do
  n=${var%%:*}; e=${var#*:}; [ "$n" = "$var" ] && e=""
  for cfg in "w32 cc-pVDZ 0 1" "alkane40 6-31G 0 1" "w32 6-31G** 0 1"; do
    env $e TAG=$n timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/rank_emu.log
    env $e TAG=$n RANK_EMU=3 WORLD_EMU=8 timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/rank_emu.log
  done
done
cat $O/rank_emu.log
