#!/bin/bash
# Round 2, call L (1 GPU): DRAM traffic of the persistent kernel for the other shard shapes of configs[2] / configs[3]
# (ncu on ONE GPU: rank 0's shard in loopback), for bench.py's roofline.traffic.
set +e
export CGB_SPIN_TIMEOUT_MS=20000
OUT=gpurun_out
mkdir -p $OUT
for CASE in 40000:2 40000:4 28284:2 56568:8 20000:1; do
  TAG=$(echo $CASE | tr ':' 'w')
  SH="python profiles/ab_iter.py --sizes $CASE --set schedule=1 --iters 20 --reps 1 --out $OUT/ab_ncu.jsonl"
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:cg_persist -s 1 -c 1 --csv --log-file $OUT/traffic_$TAG.csv $SH > $OUT/ncu_traffic_$TAG.log 2>&1
done
echo done > $OUT/done.txt
