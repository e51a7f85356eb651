#!/bin/bash
# Round 2, call I (1 GPU): does the L2 prefetch of the next mat-vec's tiles hide the vector phases?
set +e
export CGB_SPIN_TIMEOUT_MS=5000
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "persistent or interleave or generated_bitwise" > $OUT/pytest_persist.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_persist.log
timeout 300 python profiles/ab_iter.py --sizes 40000:8,10000,40000,40000:4 \
   --set schedule=1,l2_prefetch=0 --set schedule=1,l2_prefetch=1 --set schedule=1,l2_prefetch=2 --set schedule=1,l2_prefetch=3 --set schedule=1,l2_prefetch=4 --set schedule=1,l2_prefetch=8 \
   --set schedule=1,l2_prefetch=16 \
   --out $OUT/ab_l2pf.jsonl > $OUT/ab_l2pf.log 2>&1
echo done > $OUT/done.txt
