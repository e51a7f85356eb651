#!/bin/bash
# BASELINE.json configs[4]: lap2D_5pt_n100.mtx NUM_THREADS/BLOCK_WIDTH sweep on one B200 -- the
# reference CUDA solver rebuilt for sm_100a (oracle/_ref/cgsolver_cuda_ref, unmodified sources)
# beside the product's cgsolver, same command line, same results-file format.
#   gpurun --timeout 1500 -- 'bash profiles/sweep_config5.sh'
set +e
OUT=gpurun_out/config5
mkdir -p $OUT
MTX=$OUT/lap2D_5pt_n100.mtx
python -c "import sys; sys.path.insert(0,'oracle'); import oracle; oracle.write_lap2d_5pt_mtx('$MTX', 100)"
REF=oracle/_ref/cgsolver_cuda_ref
OURS=conjugate-gradient_b200/host/cgsolver
rm -f $OUT/ref_*.txt $OUT/ours_*.txt $OUT/*.log
run() { # binary tag T BW flag
  timeout 120 $1 $MTX $3 $4 $5 $OUT/$2.txt >> $OUT/$2.log 2>&1 || echo "$3,$4,timeout_or_error" >> $OUT/$2.txt
}
# code/CUDA/cg.run:20-23 -- BLOCK_WIDTH = N, row kernel (false) and column kernel (true)
for T in 2 8 32 128 512 1024; do
  run $REF ref_naive $T 10000 false;  run $OURS ours_naive $T 10000 false
  run $REF ref_naive_t $T 10000 true; run $OURS ours_naive_t $T 10000 true
done
# code/CUDA/cg.run:26-30 -- column kernel sweep (BLOCK_WIDTH = 1 left out: minutes per run)
for T in 32 64 128 256 512 1024; do
  for BW in 4 16 1024 4096; do
    run $REF ref_t $T $BW true; run $OURS ours_t $T $BW true
  done
done
grep -h "STEP" $OUT/ref_t.log | sort | uniq -c > $OUT/ref_step_lines.txt
grep -h "STEP" $OUT/ours_t.log | sort | uniq -c > $OUT/ours_step_lines.txt
echo done > $OUT/done.txt
