#!/bin/bash
set +e
export CGB_SPIN_TIMEOUT_MS=5000
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python profiles/ab_iter.py --sizes 40000,40000:2,20000,40000:8,40000:4 --iters 200 --reps 4 \
   --set schedule=1,balance=0 --set schedule=1,balance=1 --set schedule=1,balance=2 --set schedule=0 --set schedule=1,balance=2,gemv_variant=9 --set schedule=1,balance=1,gemv_variant=9 \
   --out $OUT/ab_balance.jsonl > $OUT/ab_balance.log 2>&1
timeout 200 python profiles/stress_persist.py --reps 3 --sizes 148,255,1001 > $OUT/stress2.log 2>&1; echo "exit $?" >> $OUT/stress2.log
