#!/bin/bash
# A/B timing of developer builds on the GPU box (one gpurun call):
#   here:    python -m respmon_b200.build --variant w24 "-DPU_MAX_WARPS=24"      (repeat per variant; the .so files travel)
#   gpurun:  gpurun --timeout 600 -- 'bash tools/variants.sh tag w24 alu1 ...'
# For every variant: the calibration parity tests, then per-kernel device times of 10 steps at the bench shapes.
TAG=${1:-variants}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for V in base "$@"; do
  if [ "$V" = base ]; then unset RESPMON_B200_LIB; else export RESPMON_B200_LIB=$PWD/respmon_b200/_variants/librespmon_b200.$V.so; fi
  echo "== $V" | tee -a $OUT/summary.txt
  if [ -n "$VARIANT_TESTS" ]; then timeout 200 python -m pytest tests -m gpu -q -x -k "calibrate or pipeline" 2>&1 | tail -1 | tee -a $OUT/summary.txt; fi
  timeout 100 python tools/bench_stage.py 64 10 0 > $OUT/stage_$V.log 2>&1
  head -8 $OUT/stage_$V.log | tee -a $OUT/summary.txt
  timeout 100 python tools/bench_stage.py 64 10 1 > $OUT/stage_defer_$V.log 2>&1     # overlapped steps (two engines, deferred join)
  echo "deferred: $(head -1 $OUT/stage_defer_$V.log)" | tee -a $OUT/summary.txt
done
