#!/bin/bash
# Round 2, call A (1 GPU): parity tests, then the iteration timeline (profiles/trace_iter.py)
# of the 1-GPU system and of ONE rank's shard of the 8-way split (loopback), and a variant sweep.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/run_gpu_r02_a.sh'
set +e
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/trace_iter.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 300 python profiles/trace_iter.py --case 40000:8 --case 40000:1 --case 14142:1 --case 40000:4 \
    --set l2_prefetch=4 --set l2_prefetch=0 --set l2_prefetch=8 --npz $OUT/trace_npz > $OUT/trace_a.log 2>&1
timeout 300 python profiles/trace_iter.py --case 40000:8 \
    --set gemv_variant=6 --set gemv_variant=4 --set gemv_variant=2 --set gemv_variant=10 --set gemv_variant=5 \
    --npz $OUT/trace_npz > $OUT/trace_b.log 2>&1
timeout 600 python profiles/tune_gemv.py --out $OUT/tune_gemv.jsonl \
    --cases 40000:1,40000:2,40000:4,40000:8,20000:1,28284:2,56568:8,10000:1,14142:1,4096:1 > $OUT/tune.log 2>&1
echo done > $OUT/done.txt
