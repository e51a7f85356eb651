#!/bin/bash
# Round 2, call Q (1 GPU): full suite after the gemv.cu body refactor; boundary L2 prefetch depth sweep (persistent schedule).
set +e
export CGB_SPIN_TIMEOUT_MS=5000
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 600 python profiles/ab_iter.py --sizes 70000:8,40000:8,40000:4,40000 --iters 200 --reps 5 \
  --set l2_prefetch=4 --set l2_prefetch=0 --set l2_prefetch=6 --set l2_prefetch=8 --set l2_prefetch=12 --set l2_prefetch=16 --set l2_prefetch=4 \
  --out $OUT/ab_l2depth.jsonl > $OUT/ab_l2depth.log 2>&1; echo "ab exit $?" >> $OUT/ab_l2depth.log
echo done > $OUT/done.txt
