#!/bin/bash
# Round 2, call P (1 GPU): tensor-map (UTMALDG.2D) vs 1-D bulk (UBLKCP) producer A/B, parity of the new variants, ncu pair.
set +e
export CGB_SPIN_TIMEOUT_MS=5000
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variant" > $OUT/pytest_tm2d.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_tm2d.log
timeout 600 python profiles/tensormap_ab.py > $OUT/tensormap_ab.jsonl 2> $OUT/tensormap_ab.err; echo "ab exit $?" >> $OUT/tensormap_ab.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_t -o $OUT/tensormap_pair python profiles/tensormap_ab.py --once --shapes 69856:1 > $OUT/ncu_tm2d.log 2>&1; echo "ncu exit $?" >> $OUT/ncu_tm2d.log
echo done > $OUT/done.txt
