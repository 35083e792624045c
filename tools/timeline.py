"""Developer tool: timeline of one run_batch step (kernel start/end from the library's own CUDA events).
    python tools/timeline.py [n_clips]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.engine import Engine
n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = Engine(0)
specs = [synth.clip_spec(i, 640, 480, 256) for i in range(n_clips)]
dq8 = np.stack([synth.displacement_q8(s) for s in specs])
clips = eng.synth_clips(specs, dq8)
for _ in range(3):
    eng.run_batch(clips, 10.0)
torch.cuda.synchronize()
eng.profile(True)
t0 = time.perf_counter()
eng.run_batch(clips, 10.0)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host: enqueue %.2f ms, until idle %.2f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
tl = eng.profile_timeline()
prev_end = 0.0
for name, a, b in tl:
    print("%-30s start %7.3f  end %7.3f  dur %6.3f  %s" % (name, a, b, b - a, "gap %.3f" % (a - prev_end) if a - prev_end > 0.02 else ""))
    prev_end = max(prev_end, b)
