#!/bin/bash
# compute-sanitizer over the whole path on small clips (memcheck, racecheck, synccheck, initcheck): a one-strip shape
# (160x120), a three-strip shape (640x480: named slot barriers, static two-stage ring), a 1080p frame window (level 4 in the
# record, dynamic ring, odd block count), the mapped batch path (crop kernel reading pinned host memory) and a live cohort
# with 'average' extraction.
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
# the 1080p class: level 4 of the fused pyramid kernel in the record (pyramid_g4), 2 stages x 18 warps, odd block count
spec = [synth.clip_spec(6, 1920, 1080, 12)]
big = eng.synth_clips(spec, np.stack([synth.displacement_q8(s) for s in spec]))
for g4 in (0, 1):
    eng.set_option("pyramid_g4", g4)
    lap = eng.pyramid_build_clips(big, 1, 8)
    print("1080p g4", g4, float(lap.abs().sum()))
eng.set_option("pyramid_g4", 0)
del big, lap
# mapped path: the crop kernel reads ROI rows out of pinned host memory; live cohort with 'average' extraction
from respmon_b200.batch import BatchMonitor
from respmon_b200.live import LiveCohort
specs = [synth.clip_spec(s, 160, 120, 192) for s in (3, 4, 5)]
host = torch.from_numpy(np.stack([synth.make_clip(s) for s in specs])).pin_memory()
mon = BatchMonitor(0, chunk_clips=2)
print("mapped", mon.run(host, 10.0, cal_first=1, cal_len=64)["bpm"], mon.run(host, 10.0, cal_first=1, cal_len=64)["bpm"], mon.reruns)
live = LiveCohort(2, 160, 120, 10.0, cal_len=64, method="average")
for lo in range(0, 192, 24):
    out = live.push(host[:2, lo:lo + 24])
print("live average", out["state"], out["bpm"])
PY
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_driver.py > $OUT/$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/$tool.log | tail -1)"
done
