#!/bin/bash
# One gpurun call of round 1: parity tests, smoke, the bench line, CLI runs, ncu evidence.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/run_gpu_r01.sh'
set +e
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/smi.txt
nproc >> $OUT/smi.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/smi.txt; free -g | head -2 >> $OUT/smi.txt

timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log

timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench exit $?" >> $OUT/bench_n1.err
timeout 600 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err

# the host program, both command lines (results rows + DEBUG lines)
CG=conjugate-gradient_b200/host/cgsolver
python -c "import sys; sys.path.insert(0,'oracle'); import oracle; oracle.write_lap2d_5pt_mtx('$OUT/lap2D_5pt_n100.mtx', 100)"
( CGB_JSON=$OUT/cli.jsonl timeout 300 $CG 20000 $OUT/results_form1.txt
  CGB_JSON=$OUT/cli.jsonl timeout 300 $CG 40000 $OUT/results_form1.txt 200
  CGB_JSON=$OUT/cli.jsonl timeout 300 $CG $OUT/lap2D_5pt_n100.mtx 256 512 true $OUT/results_form2.txt
  CGB_JSON=$OUT/cli.jsonl timeout 300 $CG $OUT/lap2D_5pt_n100.mtx 128, 16, false $OUT/results_form2.txt
) > $OUT/cli.log 2>&1

# ncu: launch list of a short bench run, then one full capture of the mat-vec kernel
CMD="python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu-baseline"
$CMD > $OUT/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
$CMD > $OUT/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_tma -s 30 -c 3 \
    -f -o $OUT/prof_gemv $CMD > $OUT/ncu_full.log 2>&1
echo done > $OUT/done.txt
