#!/bin/bash
# compute-sanitizer over the whole path on small clips (memcheck, racecheck, synccheck, initcheck): a one-strip shape
# (160x120) and a three-strip shape (640x480: interior + edge variants of the fused TMA pyramid kernel, named barriers).
#   gpurun --timeout 2400 -- 'bash tools/sanitize.sh r02_sanitize'
OUT=gpurun_out/${1:-sanitize}
mkdir -p $OUT
cat > /tmp/san_driver.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.engine import Engine, results_to_numpy
eng = Engine(0)
for (w, h, t, cal, seeds) in ((160, 120, 192, 64, (3, 4)), (640, 480, 96, 32, (0, 2))):
    specs = [synth.clip_spec(s, w, h, t) for s in seeds]
    dq8 = np.stack([synth.displacement_q8(s) for s in specs])
    clips = eng.synth_clips(specs, dq8)
    rec = eng.run_batch(clips, 10.0, cal_first=1, cal_len=cal)
    print(w, h, results_to_numpy(rec))
PY
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_driver.py > $OUT/$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/$tool.log | tail -1)"
done
