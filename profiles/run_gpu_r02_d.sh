#!/bin/bash
# Round 2, call D (1 GPU): persistent kernel -- parity, timing per shape, timeline.
set +e
export CGB_SPIN_TIMEOUT_MS=3000
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/trace_iter.jsonl
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "persistent or interleave or generated_bitwise or nonzero_x0 or mtx_bitwise or early_stop" > $OUT/pytest_persist.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_persist.log
timeout 300 python profiles/ab_iter.py --sizes 40000:8,40000:4,40000,56568:8,10000 \
   --set schedule=0 --set schedule=1 --set schedule=1,gemv_variant=9 --set schedule=1,gemv_variant=10 --set schedule=1,gemv_variant=2 \
   --set schedule=1,gemv_variant=0,l2_prefetch=0 --set schedule=1,l2_prefetch=8 --set schedule=1,l2_prefetch=4,l2_prefetch_mode=1  --set schedule=1,l2_prefetch=8,l2_prefetch_mode=1 \
   --out $OUT/ab_loopback.jsonl > $OUT/ab_loopback.log 2>&1
timeout 300 python profiles/trace_iter.py --case 40000:8 --case 40000:1 \
    --set schedule=1 --set schedule=1,gemv_variant=9 --npz $OUT/trace_npz > $OUT/trace_d.log 2>&1
echo done > $OUT/done.txt
