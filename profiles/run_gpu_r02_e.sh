#!/bin/bash
# Round 2, call E (G GPUs): the whole GPU test suite incl. the multi-GPU tests, then bench lines.
#   gpurun --gpus G --timeout 1500 -- 'bash profiles/run_gpu_r02_e.sh G'
set +e
G=${1:-2}
export CGB_SPIN_TIMEOUT_MS=10000
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $OUT/smi_g$G.txt
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_g$G.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_g$G.log
timeout 300 python bench.py --steps 3 --warmup 3 --cpu-iters 20 > $OUT/bench_g1.json 2> $OUT/bench_g1.err; echo "exit $?" >> $OUT/bench_g1.err
if [ $G -gt 1 ]; then
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $G --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_g$G.json 2> $OUT/bench_g$G.err; echo "exit $?" >> $OUT/bench_g$G.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus $G --steps 3 --warmup 3 --no-cpu-baseline --workload weak > $OUT/bench_g${G}_weak.json 2> $OUT/bench_g${G}_weak.err; echo "exit $?" >> $OUT/bench_g${G}_weak.err
fi
echo done > $OUT/done_g$G.txt
