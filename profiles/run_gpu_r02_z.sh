#!/bin/bash
# Round 2, call Z (1 GPU): evict_last hint on the boundary + ramp prefetch, alternating with no hint, all shard shapes.
set +e
export CGB_SPIN_TIMEOUT_MS=5000
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python profiles/ab_iter.py --sizes 40000:8,70000:8,40000:4,40000:2,40000 --iters 200 --reps 4 \
  --set l2_pf_policy=0 --set l2_pf_policy=1 --set l2_pf_policy=0 --set l2_pf_policy=1 --set l2_pf_policy=1,l2_ramp=4 --set l2_pf_policy=1,l2_ramp=0 \
  --out $OUT/ab_l2policy2.jsonl > $OUT/ab_l2policy2.log 2>&1; echo "ab exit $?" >> $OUT/ab_l2policy2.log
echo done > $OUT/done.txt
