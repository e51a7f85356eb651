#!/bin/bash
# Round 2, call U (1 GPU): compute-sanitizer memcheck over one small pass of every kernel.
set +e
export CGB_SPIN_TIMEOUT_MS=60000
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python profiles/sanitize_small.py > $OUT/sanitize_plain.log 2>&1; echo "plain exit $?" >> $OUT/sanitize_plain.log
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 1 python profiles/sanitize_small.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck exit $?" >> $OUT/sanitize_memcheck.log
echo done > $OUT/done.txt
