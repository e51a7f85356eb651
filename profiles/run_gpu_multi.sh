#!/bin/bash
# Multi-GPU call:  gpurun --gpus G --timeout 1200 -- 'bash profiles/run_gpu_multi.sh G [tests|multitests|notests] [first_g]'
set +e
G=${1:-2}
MODE=${2:-notests}
FIRST=${3:-1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $OUT/smi_g$G.txt
nvidia-smi topo -m >> $OUT/smi_g$G.txt 2>&1
if [ "$MODE" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_g$G.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_g$G.log
elif [ "$MODE" = "multitests" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q > $OUT/pytest_gpu_g$G.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_g$G.log
fi
run_bench() { # gpus exchange tag extra...
  local g=$1 ex=$2 tag=$3; shift 3
  if [ "$g" = "1" ]; then
    timeout 600 python bench.py --no-cpu-baseline "$@" > $OUT/bench_${tag}.json 2> $OUT/bench_${tag}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 \
      --master-port 29611 bench.py --gpus $g --exchange $ex "$@" > $OUT/bench_${tag}.json 2> $OUT/bench_${tag}.err
  fi
  echo "exit $?" >> $OUT/bench_${tag}.err
}
g=$FIRST
while [ $g -le $G ]; do
  if [ $g -eq 1 ]; then
    run_bench 1 1 g1
  else
    run_bench $g 1 g${g}_fused
    run_bench $g 0 g${g}_nccl
    run_bench $g 1 g${g}_fused_weak --workload weak
  fi
  g=$((g*2))
done
# the reference-style command line on all GPUs (results row "n,psize,seconds")
CG=conjugate-gradient_b200/host/cgsolver
( CGB_GPUS=$G CGB_JSON=$OUT/cli_g$G.jsonl timeout 300 $CG 40000 $OUT/results_g$G.txt 200
  CGB_GPUS=$G CGB_JSON=$OUT/cli_g$G.jsonl timeout 300 $CG 20000 $OUT/results_g$G.txt ) > $OUT/cli_g$G.log 2>&1
echo done > $OUT/done_g$G.txt
