#!/bin/bash
# round 2, GPU call 20: register-resident in-core kernels for the L >= 6 class pairs (dclass_hit_*.cu) -- parity suite, in-core d-shell timing
mkdir -p gpurun_out
O=gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_call20.log 2>&1; echo "rc=$?"; tail -3 $O/tests_call20.log
for cfg in "w4 6-31G** 0 0" "w4 6-31G** 1 0" "w16 6-31G** 0 0"; do timeout 300 python tools/time_config.py $cfg 2>&1 | head -2; done
echo "== same with the warp-per-quartet fallback for d shells (kernel=2 is not the same thing; FOCK_NO_HI_INCORE=1 restores r01 dispatch)"
for cfg in "w4 6-31G** 0 0" "w16 6-31G** 0 0"; do FOCK_NO_HI_INCORE=1 timeout 300 python tools/time_config.py $cfg 2>&1 | head -2; done
