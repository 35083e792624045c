#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, ncu launch list.  Usage: gpurun -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/smi.txt; nproc >> $OUT/smi.txt; free -g >> $OUT/smi.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $OUT/gpu_tests.log
tail -5 $OUT/gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --clips 16 --e2e-steps 1 > $OUT/ncu_bench.log 2>&1
tail -2 $OUT/ncu_bench.log | cut -c1-300
