#!/bin/bash
# A/B timing of developer builds on the GPU box (one gpurun call):
#   here:    python -m respmon_b200.build --variant u1 "-DPF_UNROLL=1"      (repeat per variant; the .so files travel)
#   gpurun:  gpurun --timeout 600 -- 'bash tools/variants.sh tag "python tools/bench_pyramid.py 640 480 64" u1 dyn ...'
TAG=$1; CMD=$2; shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for V in base "$@"; do
  if [ "$V" = base ]; then unset RESPMON_B200_LIB; else export RESPMON_B200_LIB=$PWD/respmon_b200/_variants/librespmon_b200.$V.so; fi
  echo "== $V" | tee -a $OUT/summary.txt
  timeout 200 $CMD 2>&1 | grep -vE "^\s*$|Warning" | head -${HEAD:-12} | tee -a $OUT/summary.txt
done
