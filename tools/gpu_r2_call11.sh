#!/bin/bash
# round 2, GPU call 11: (magnitude bucket, Morton) pair order -- parity suite, bucket-width sweep against the r01 order (same build)
mkdir -p gpurun_out
O=gpurun_out
echo "== gpu tests (default build)"; timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_call11.log 2>&1; echo "rc=$?"; tail -3 $O/tests_call11.log
: > $O/variants.log
for ord in diag b0 b2 b4 b6 b8; do
  if [ $ord = diag ]; then export FOCK_PAIR_ORDER=diag; else export FOCK_PAIR_ORDER=spatial FOCK_ORDER_BUCKET_BITS=${ord#b}; fi
  for cfg in "alkane40 6-31G 0 1" "w16 6-31G 0 0" "w32 cc-pVDZ 0 1" "w32 6-31G** 0 1"; do
    TAG=$ord timeout 300 python tools/time_config.py $cfg 2>&1 | head -1 >> $O/variants.log
  done
done
unset FOCK_PAIR_ORDER FOCK_ORDER_BUCKET_BITS
cat $O/variants.log
