#!/bin/bash
# Round 2, call K (4 GPUs): bench lines at 4 GPUs exactly as the driver launches them (CPU baseline on).
set +e
G=${1:-4}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29621 bench.py --gpus $G --steps 5 --warmup 3 --cpu-iters 20 > $OUT/bench_g$G.json 2> $OUT/bench_g$G.err; echo "exit $?" >> $OUT/bench_g$G.err
timeout 300 $TR --master-port 29622 bench.py --gpus $G --steps 5 --warmup 3 --no-cpu-baseline --workload weak > $OUT/bench_g${G}_weak.json 2> $OUT/bench_g${G}_weak.err; echo "exit $?" >> $OUT/bench_g${G}_weak.err
timeout 300 $TR --master-port 29623 bench.py --impl reference --gpus $G --steps 5 --warmup 3 > $OUT/bench_ref_g$G.json 2> $OUT/bench_ref_g$G.err; echo "exit $?" >> $OUT/bench_ref_g$G.err
echo done > $OUT/done_g$G.txt
