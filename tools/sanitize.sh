#!/bin/bash
# compute-sanitizer over the whole path on a small clip (memcheck, racecheck, synccheck).  gpurun -- 'bash tools/sanitize.sh'
OUT=gpurun_out/sanitize
mkdir -p $OUT
cat > /tmp/san_driver.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.engine import Engine, results_to_numpy
eng = Engine(0)
specs = [synth.clip_spec(s, 160, 120, 192) for s in (3, 4)]
dq8 = np.stack([synth.displacement_q8(s) for s in specs])
clips = eng.synth_clips(specs, dq8)
rec = eng.run_batch(clips, 10.0, cal_first=1, cal_len=64)
print(results_to_numpy(rec))
PY
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_driver.py > $OUT/$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/$tool.log | tail -1)"
done
