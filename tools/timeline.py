"""Developer tool: start / end of every kernel launch of one run_batch step at the bench shape (CUDA events recorded by the
library on the launching streams, rm_profile_*), to see what overlaps what.   python tools/timeline.py [n_clips] [chunks]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.engine import Engine
n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = Engine(0)
if len(sys.argv) > 2:
    eng.set_option("measure_chunks", int(sys.argv[2]))
specs = [synth.clip_spec(i, 640, 480, 256) for i in range(n_clips)]
clips = eng.synth_clips(specs, np.stack([synth.displacement_q8(s) for s in specs]))
for _ in range(3):
    eng.run_batch(clips, 10.0)
torch.cuda.synchronize()
eng.profile(True)
eng.run_batch(clips, 10.0)
torch.cuda.synchronize()
for name, a, b in sorted(eng.profile_timeline(), key=lambda r: r[1]):
    print("%8.3f -> %8.3f  (%6.3f ms)  %s" % (a, b, b - a, name))
