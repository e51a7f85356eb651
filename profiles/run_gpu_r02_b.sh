#!/bin/bash
# Round 2, call B (1 GPU): first run of the persistent schedule -- parity, then timing.
set +e
export CGB_SPIN_TIMEOUT_MS=3000
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/trace_iter.jsonl
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "persistent or interleave or generated_bitwise" > $OUT/pytest_persist.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_persist.log
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 300 python profiles/ab_iter.py --sizes 40000,14142,10000 --set schedule=0 --set schedule=1 --out $OUT/ab_sched.jsonl > $OUT/ab_sched.log 2>&1
timeout 300 python profiles/ab_iter.py --sizes 40000:8,40000:4,40000:2,56568:8 \
   --set schedule=0 --set schedule=1 --set schedule=1,gemv_variant=10 --set schedule=1,gemv_variant=9 --set schedule=0,gemv_variant=10 \
   --out $OUT/ab_loopback.jsonl > $OUT/ab_loopback.log 2>&1
echo done > $OUT/done.txt
