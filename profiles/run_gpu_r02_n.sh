#!/bin/bash
# Round 2, call N (1 GPU): stress of the persistent kernel on tiny / ragged systems, then the full suite.
set +e
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python profiles/stress_persist.py --reps 10 > $OUT/stress.log 2>&1; echo "exit $?" >> $OUT/stress.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
