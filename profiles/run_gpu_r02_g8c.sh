#!/bin/bash
# Round 2, closing 8-GPU call (final binary): multi-GPU parity tests at 8 ranks, then the scaling run on ONE box
# (N = 8, 4, 2, 1 back to back, launched as the driver launches them).
set +e
OUT=gpurun_out
mkdir -p $OUT
export CGB_SPIN_TIMEOUT_MS=8000
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q -k "not nccl" > $OUT/pytest_gpu_g8.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_g8.log
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "nccl" > $OUT/pytest_gpu_g8_nccl.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_g8_nccl.log
unset CGB_SPIN_TIMEOUT_MS
port=29650
for G in 8 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $port"
  extra="--no-cpu-baseline"; [ $G = 8 ] && extra=""
  timeout 400 $TR bench.py --gpus $G --steps 20 --warmup 5 $extra > $OUT/bench_scale_g$G.json 2> $OUT/bench_scale_g$G.err; echo "exit $?" >> $OUT/bench_scale_g$G.err
  port=$((port+1))
done
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_scale_g1.json 2> $OUT/bench_scale_g1.err; echo "exit $?" >> $OUT/bench_scale_g1.err
echo done > $OUT/done_g8.txt
