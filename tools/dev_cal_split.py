"""Developer tool (untimed experiment): Engine.run_batch(cal_parts=k) -- the calibration of k groups of clips on side
streams next to each other -- against the default, at the bench shapes: step time and identical records.
    python tools/dev_cal_split.py [n_clips] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.engine import Engine, results_to_numpy

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
eng = Engine(0)
specs = [synth.clip_spec(i, 640, 480, 256) for i in range(n_clips)]
dq8 = np.stack([synth.displacement_q8(s) for s in specs])
clips = eng.synth_clips(specs, dq8)
base = None
for parts in (1, 2, 4):
    rec = eng.run_batch(clips, 10.0, cal_parts=parts)
    for _ in range(2):
        eng.run_batch(clips, 10.0, cal_parts=parts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.run_batch(clips, 10.0, cal_parts=parts)
    e1.record()
    torch.cuda.synchronize()
    cal = []
    for _ in range(steps):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        eng.locate(clips, 10.0, 1, 128, parts=parts)
        c1.record()
        torch.cuda.synchronize()
        cal.append(c0.elapsed_time(c1))
    r = results_to_numpy(rec)
    cur = r.tobytes()
    same = "-" if base is None else (cur == base)
    base = base or cur
    print("cal_parts %d: step %.3f ms, calibration alone %.3f ms, records identical to default: %s" % (
        parts, e0.elapsed_time(e1) / steps, float(np.median(cal)), same))
