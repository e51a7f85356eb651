#!/bin/bash
# Round 2, call W (2 GPUs): multi-GPU + CLI tests after the loopback clamp (incl. the ragged-rows CLI case); the small all-kernels pass.
set +e
export CGB_SPIN_TIMEOUT_MS=5000
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q > $OUT/pytest_gpu_g2.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_g2.log
timeout 100 python profiles/sanitize_small.py > $OUT/sanitize_plain.log 2>&1; echo "plain exit $?" >> $OUT/sanitize_plain.log
echo done > $OUT/done.txt
