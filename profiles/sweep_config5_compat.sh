#!/bin/bash
# Config 5 with CGB_KERNEL=compat: the reference's own launch topologies re-created in
# csrc/compat.cu (NUM_THREADS / BLOCK_WIDTH literal, deterministic reduction) on one B200.
#   gpurun --timeout 900 -- 'bash profiles/sweep_config5_compat.sh'
set +e
OUT=gpurun_out/config5
mkdir -p $OUT
MTX=$OUT/lap2D_5pt_n100.mtx
python -c "import sys; sys.path.insert(0,'oracle'); import oracle; oracle.write_lap2d_5pt_mtx('$MTX', 100)"
OURS=conjugate-gradient_b200/host/cgsolver
rm -f $OUT/compat_*.txt $OUT/compat_*.log
run() { # tag T BW flag
  CGB_KERNEL=compat timeout 60 $OURS $MTX $2 $3 $4 $OUT/$1.txt >> $OUT/$1.log 2>&1 || echo "$2,$3,timeout_or_error" >> $OUT/$1.txt
}
for T in 2 8 32 128 512 1024; do
  run compat_naive $T 10000 false
  run compat_naive_t $T 10000 true
done
for T in 32 64 128 256 512 1024; do
  for BW in 1 4 16 1024 4096; do
    run compat_t $T $BW true
  done
done
grep -h "STEP" $OUT/compat_t.log | sort | uniq -c > $OUT/compat_step_lines.txt
echo done > $OUT/done_compat.txt
