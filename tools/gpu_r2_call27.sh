#!/bin/bash
# round 2, GPU call 27: fewer blocks (fewer Boys-table copies) for launches with few tiles -- A/B at full size and on a 1/8 share
mkdir -p gpurun_out
O=gpurun_out
: > $O/variants.log
for tpw in 1 2 4 8; do
  for cfg in "w32 cc-pVDZ 0 1" "w32 6-31G** 0 1" "alkane40 6-31G 0 1"; do
    FOCK_TILES_PER_WARP=$tpw TAG=tpw$tpw timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/variants.log
    FOCK_TILES_PER_WARP=$tpw TAG=tpw$tpw RANK_EMU=3 WORLD_EMU=8 timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/variants.log
  done
done
cat $O/variants.log
