"""Developer tool: per-kernel device times of one run_batch step at the bench shapes (not the contract; see bench.py).
    python tools/bench_stage.py [n_clips] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.engine import Engine, results_to_numpy

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
eng = Engine(0)
if os.environ.get("RM_CHUNKS"):
    eng.set_option("measure_chunks", int(os.environ["RM_CHUNKS"]))
specs = [synth.clip_spec(i, 640, 480, 256) for i in range(n_clips)]
dq8 = np.stack([synth.displacement_q8(s) for s in specs])
clips = eng.synth_clips(specs, dq8)
for _ in range(3):
    rec = eng.run_batch(clips, 10.0)
torch.cuda.synchronize()
eng.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    rec = eng.run_batch(clips, 10.0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
prof = eng.profile_report()
r = results_to_numpy(rec)
print("step %.3f ms  -> %.0f frames/s   ok %d/%d   bpm checksum %.6f" % (ms, n_clips * 256 / ms * 1e3, int((r["status"] == 0).sum()), n_clips,
      float(np.nansum(r["bpm"]))))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print("  %-32s %8.3f ms/step  x%d" % (k, v[0] / steps, v[1] // steps))
