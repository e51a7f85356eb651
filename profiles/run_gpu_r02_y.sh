#!/bin/bash
# Round 2, call Y (1 GPU): L2 policy of the boundary prefetch (none / evict_last / evict_first) x depth.
set +e
export CGB_SPIN_TIMEOUT_MS=5000
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python profiles/ab_iter.py --sizes 40000:8,40000:4 --iters 200 --reps 4 \
  --set l2_pf_policy=0,l2_prefetch=4 --set l2_pf_policy=1,l2_prefetch=4 --set l2_pf_policy=1,l2_prefetch=6 --set l2_pf_policy=1,l2_prefetch=8 \
  --set l2_pf_policy=2,l2_prefetch=4 --set l2_pf_policy=2,l2_prefetch=8 --set l2_pf_policy=0,l2_prefetch=4 \
  --out $OUT/ab_l2policy.jsonl > $OUT/ab_l2policy.log 2>&1; echo "ab exit $?" >> $OUT/ab_l2policy.log
echo done > $OUT/done.txt
