"""Developer tool: n steps of run_batch at the bench shape (for ncu).  python tools/prof_step.py [n_clips] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.engine import Engine
n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = Engine(0)
specs = [synth.clip_spec(i, 640, 480, 256) for i in range(n_clips)]
dq8 = np.stack([synth.displacement_q8(s) for s in specs])
clips = eng.synth_clips(specs, dq8)
for _ in range(steps):
    eng.run_batch(clips, 10.0)
torch.cuda.synchronize()
