#!/bin/bash
# Round 2, call V (1 GPU): full suite with the loopback clamp + the small all-kernels pass.
set +e
export CGB_SPIN_TIMEOUT_MS=5000
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python profiles/sanitize_small.py > $OUT/sanitize_plain.log 2>&1; echo "plain exit $?" >> $OUT/sanitize_plain.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
echo done > $OUT/done.txt
