#!/bin/bash
# Round 2, final 8-GPU call: multi-GPU parity tests at 8 ranks + the bench lines exactly as the driver launches them.
set +e
G=8
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29631 bench.py --gpus $G --steps 20 --warmup 5 > $OUT/bench_g$G.json 2> $OUT/bench_g$G.err; echo "exit $?" >> $OUT/bench_g$G.err
timeout 300 $TR --master-port 29632 bench.py --gpus $G --steps 5 --warmup 3 --no-cpu-baseline --workload weak > $OUT/bench_g${G}_weak.json 2> $OUT/bench_g${G}_weak.err; echo "exit $?" >> $OUT/bench_g${G}_weak.err
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q -k "not nccl" > $OUT/pytest_gpu_g$G.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_g$G.log
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "nccl" > $OUT/pytest_gpu_g${G}_nccl.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_g${G}_nccl.log
echo done > $OUT/done_g$G.txt
