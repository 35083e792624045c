#!/bin/bash
# gpurun -- 'bash tools/gpu_tests_only.sh tag "pytest -k expr" [ncu kernel regex]'
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$2" ]; then K=(-k "$2"); else K=(); fi
timeout 1200 python -m pytest tests -m gpu -q "${K[@]}" 2>&1 | tail -60 > $OUT/gpu_tests.log
cat $OUT/gpu_tests.log | tail -50
if [ -n "$3" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches.csv \
      python tools/prof_step.py 64 2 > $OUT/launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 -f -o $OUT/$3 \
      python tools/prof_step.py 64 2 > $OUT/ncu_$3.log 2>&1
  tail -3 $OUT/ncu_$3.log
fi
