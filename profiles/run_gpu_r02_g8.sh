#!/bin/bash
# Round 2, 8-GPU call: multi-GPU parity tests at 8 ranks, bench lines (both schedules, weak), real-peer timeline.
#   gpurun --gpus 8 --timeout 1500 -- 'bash profiles/run_gpu_r02_g8.sh'
set +e
G=8
export CGB_SPIN_TIMEOUT_MS=15000
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/trace_iter.jsonl
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $OUT/smi_g$G.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 bench.py --gpus $G --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_g$G.json 2> $OUT/bench_g$G.err; echo "exit $?" >> $OUT/bench_g$G.err
timeout 300 $TR --master-port 29612 bench.py --gpus $G --steps 5 --warmup 3 --no-cpu-baseline --schedule 0 > $OUT/bench_g${G}_graph.json 2> $OUT/bench_g${G}_graph.err; echo "exit $?" >> $OUT/bench_g${G}_graph.err
timeout 300 $TR --master-port 29613 bench.py --gpus $G --steps 5 --warmup 3 --no-cpu-baseline --workload weak > $OUT/bench_g${G}_weak.json 2> $OUT/bench_g${G}_weak.err; echo "exit $?" >> $OUT/bench_g${G}_weak.err
timeout 200 $TR --master-port 29614 profiles/trace_iter.py --real --case 40000:8 --set schedule=1 --set schedule=0 --npz $OUT/trace_npz_g8 --out $OUT/trace_g8.jsonl > $OUT/trace_g8.log 2>&1
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q -k "not nccl" > $OUT/pytest_gpu_g$G.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_g$G.log
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "nccl" > $OUT/pytest_gpu_g${G}_nccl.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_g${G}_nccl.log
echo done > $OUT/done_g$G.txt
