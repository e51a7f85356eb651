#!/bin/bash
# Round 2, call S (1 GPU): tensor-map A/B on the 5000-row shard and N=10000 (the shapes VERDICT r01 names); final driver-style bench line.
set +e
export CGB_SPIN_TIMEOUT_MS=5000
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python profiles/tensormap_ab.py --shapes 40000:8,10000:1,40000:1 --reps 50 > $OUT/tensormap_ab2.jsonl 2> $OUT/tensormap_ab2.err; echo "ab exit $?" >> $OUT/tensormap_ab2.err
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_g1_final.json 2> $OUT/bench_g1_final.err; echo "bench exit $?" >> $OUT/bench_g1_final.err
echo done > $OUT/done.txt
