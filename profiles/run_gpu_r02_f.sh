#!/bin/bash
# Round 2, call F (1 GPU): chunk-based p'Ap + adaptive row balancing -- full GPU suite, then timing.
set +e
export CGB_SPIN_TIMEOUT_MS=5000
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/trace_iter.jsonl
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
true \
   --set schedule=0 --set schedule=1,balance=0 --set schedule=1,balance=1 --set schedule=1,balance=1,gemv_variant=9 --set schedule=1,balance=1,gemv_variant=10 \
   --out $OUT/ab_loopback.jsonl > $OUT/ab_loopback.log 2>&1
timeout 300 python profiles/trace_iter.py --case 40000:8 --case 40000:1 \
    --set schedule=1,balance=1 --set schedule=1,balance=0 --npz $OUT/trace_npz > $OUT/trace_f.log 2>&1
echo done > $OUT/done.txt
