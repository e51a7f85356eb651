#!/bin/bash
# Round 2, call X (1 GPU): last look at the final build -- smoke() and bench.py with no flags.
set +e
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
timeout 400 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench exit $?" >> $OUT/bench_default.err
echo done > $OUT/done.txt
