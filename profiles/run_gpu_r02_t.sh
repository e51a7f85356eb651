#!/bin/bash
# Round 2, call T (1 GPU): poll back-off cap sweep (persistent schedule, loopback shards).
set +e
export CGB_SPIN_TIMEOUT_MS=5000
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python profiles/ab_iter.py --sizes 40000:8,40000:4,40000 --iters 200 --reps 5 \
  --set poll_ns=320 --set poll_ns=160 --set poll_ns=80 --set poll_ns=40 --set poll_ns=20 --set poll_ns=640 --set poll_ns=320 \
  --out $OUT/ab_pollns.jsonl > $OUT/ab_pollns.log 2>&1; echo "ab exit $?" >> $OUT/ab_pollns.log
echo done > $OUT/done.txt
