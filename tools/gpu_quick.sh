#!/bin/bash
# Quick GPU pass: parity tests + bench (no CPU baseline).  Usage: gpurun -- 'bash tools/gpu_quick.sh tag [pytest -k expr]'
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$2" ]; then K=(-k "$2"); else K=(); fi
timeout 900 python -m pytest tests -m gpu -q "${K[@]}" 2>&1 | tail -40 > $OUT/gpu_tests.log
tail -15 $OUT/gpu_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err
tail -5 $OUT/bench.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench.json"))
    print("value %.0f frames/s  ms/step %.2f  e2e %.0f  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
    print("roofline", d["roofline"])
    print("clocks", d["clocks"], "results", d["results"])
    for k in d["kernels"][:12]: print("  %-30s %8.3f ms %5.1f%%" % (k["name"], k["ms_per_step"], 100*k["share"]))
except Exception as e: print("bench parse failed", e)
PY
