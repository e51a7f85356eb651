#!/bin/bash
# Round 2, call C (1 GPU): timeline of the persistent kernel on the 8-way shard (loopback).
set +e
export CGB_SPIN_TIMEOUT_MS=3000
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/trace_iter.jsonl
timeout 300 python profiles/trace_iter.py --case 40000:8 --case 40000:1 \
    --set schedule=1,l2_prefetch=4 --set schedule=1,l2_prefetch=0 --set schedule=1,l2_prefetch=8 --set schedule=1,l2_prefetch=16 \
    --set schedule=1,l2_prefetch=4,gemv_variant=9 --set schedule=1,l2_prefetch=8,gemv_variant=9 --set schedule=1,l2_prefetch=4,gemv_variant=10 \
    --npz $OUT/trace_npz > $OUT/trace_c.log 2>&1
echo done > $OUT/done.txt
