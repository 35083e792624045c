import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.engine import Engine
eng = Engine(0)
specs = [synth.clip_spec(i, 640, 480, 256) for i in range(64)]
dq8 = np.stack([synth.displacement_q8(s) for s in specs])
clips = eng.synth_clips(specs, dq8)
rec, taps = eng.run_batch(clips, 10.0, keep=True)
roi = taps["roi"].cpu().numpy(); npts = taps["npts"].cpu().numpy()
area = roi[:, 2] * roi[:, 3]
print("npts  :", sorted(npts.tolist()))
print("area  :", sorted(area.tolist()))
print("w     :", sorted(roi[:,2].tolist())); print("h     :", sorted(roi[:,3].tolist()))
# time LK per clip alone
for i in np.argsort(-npts)[:4].tolist() + np.argsort(-area)[:3].tolist():
    c = clips[i:i+1]; r = taps["roi"][i:i+1]
    for _ in range(2): eng.measure_flow(c, r, 130, 126)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.measure_flow(c, r, 130, 126); b.record(); torch.cuda.synchronize()
    print("clip %2d npts %3d roi %s  measure_flow %.2f ms" % (i, npts[i], roi[i].tolist(), a.elapsed_time(b)))
