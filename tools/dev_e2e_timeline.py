"""Developer tool: host-side timeline of BatchMonitor.submit()/collect() at the bench shape (where does the H2D queue drain?).
    python tools/dev_e2e_timeline.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.batch import BatchMonitor
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mon = BatchMonitor(0, chunk_clips=int(os.environ.get("RM_CHUNK", "32")), mapped=os.environ.get("RM_MAPPED", "1") == "1")
eng = mon.engine
specs = [synth.clip_spec(i, 640, 480, 256) for i in range(64)]
dq8 = np.stack([synth.displacement_q8(s) for s in specs])
clips = eng.synth_clips(specs, dq8)
host = torch.empty(clips.shape, dtype=torch.uint8).pin_memory()
host.copy_(clips)
torch.cuda.synchronize()
mon.run(host, 10.0)
mon.run(host, 10.0)
torch.cuda.synchronize()
mon.trace = []
t0 = time.perf_counter()
prev = None
for k in range(steps):
    mon.trace.append(("submit", k, time.perf_counter()))
    ticket = mon.submit(host, 10.0)
    mon.trace.append(("submitted", k, time.perf_counter()))
    if prev is not None:
        mon.collect(prev)
        mon.trace.append(("collected", k - 1, time.perf_counter()))
    prev = ticket
mon.collect(prev)
t1 = time.perf_counter()
print("%.2f ms per step, %.0f frames/s" % ((t1 - t0) / steps * 1e3, 64 * 256 * steps / (t1 - t0)))
last = t0
for label, i, t in mon.trace:
    if t - t0 > (t1 - t0) * 0.4 and t - t0 < (t1 - t0) * 0.8:
        print("%9.2f ms  (+%6.2f)  %-16s %d" % ((t - t0) * 1e3, (t - last) * 1e3, label, i))
    last = t
