#!/bin/bash
# Round 2, call J (1 GPU): ncu evidence (launch list of a bench.py run, --set full of the persistent kernel on
# the full matrix and on the 8-way shard), weak-scaling G=1 line, config-5 CLI sweep with a clock record.
set +e
export CGB_SPIN_TIMEOUT_MS=20000
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu-baseline --no-autotune"
$CMD > $OUT/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
$CMD > $OUT/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cg_persist -s 3 -c 1 \
    -f -o $OUT/prof_persist $CMD > $OUT/ncu_full.log 2>&1
SH="python profiles/ab_iter.py --sizes 40000:8 --set schedule=1 --iters 20 --reps 1 --out $OUT/ab_ncu.jsonl"
$SH > $OUT/plain3.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cg_persist -s 1 -c 1 \
    -f -o $OUT/prof_persist_shard $SH > $OUT/ncu_shard.log 2>&1
timeout 300 python bench.py --workload weak --steps 5 --warmup 3 --cpu-iters 40 > $OUT/bench_g1_weak.json 2> $OUT/bench_g1_weak.err
# config 5 with a clock record
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 500 > $OUT/config5_clocks.csv 2>/dev/null &
SMI=$!
bash profiles/sweep_config5.sh
kill $SMI
echo done > $OUT/done.txt
