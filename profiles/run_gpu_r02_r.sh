#!/bin/bash
# Round 2, call R (1 GPU): ramp prefetch (L2 prefetch of the steps after the on-chip ones, issued when p arrives) sweep.
set +e
export CGB_SPIN_TIMEOUT_MS=5000
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python profiles/ab_iter.py --sizes 40000:8,70000:8,40000:4,40000 --iters 200 --reps 5 \
  --set l2_ramp=0 --set l2_ramp=1 --set l2_ramp=2 --set l2_ramp=3 --set l2_ramp=4 --set l2_ramp=6 --set l2_ramp=0 \
  --out $OUT/ab_l2ramp.jsonl > $OUT/ab_l2ramp.log 2>&1; echo "ab exit $?" >> $OUT/ab_l2ramp.log
echo done > $OUT/done.txt
