"""Developer tool: per-kernel device times of run_batch steps at the bench shapes (not the contract; see bench.py).
    python tools/bench_stage.py [n_clips] [steps] [cal_parts]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.engine import Engine, results_to_numpy

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
parts = int(sys.argv[3]) if len(sys.argv) > 3 else 1
engs = [Engine(0)]
for e in engs:
    if os.environ.get("RM_CHUNKS"):
        e.set_option("measure_chunks", int(os.environ["RM_CHUNKS"]))
    for opt in ("pyramid_mode", "pyramid_cfg", "temporal_sparse"):
        if os.environ.get("RM_" + opt.upper()):
            e.set_option(opt, int(os.environ["RM_" + opt.upper()]))
eng = engs[0]
specs = [synth.clip_spec(i, 640, 480, 256) for i in range(n_clips)]
dq8 = np.stack([synth.displacement_q8(s) for s in specs])
clips = eng.synth_clips(specs, dq8)
recs = [torch.empty((n_clips, 32), dtype=torch.uint8, device="cuda") for _ in engs]

def run(n):
    for k in range(n):
        engs[k % len(engs)].run_batch(clips, 10.0, out=recs[k % len(engs)], cal_parts=parts)
    return recs[(n - 1) % len(engs)]

run(3)
torch.cuda.synchronize()
for e in engs:
    e.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
rec = run(steps)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
prof = {}
for e in engs:
    for k, v in e.profile_report().items():
        a = prof.setdefault(k, [0.0, 0]); a[0] += v[0]; a[1] += v[1]
r = results_to_numpy(rec)
print("step %.3f ms  -> %.0f frames/s   ok %d/%d   bpm checksum %.6f" % (ms, n_clips * 256 / ms * 1e3, int((r["status"] == 0).sum()), n_clips,
      float(np.nansum(r["bpm"]))))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print("  %-32s %8.3f ms/step  x%d" % (k, v[0] / steps, v[1] // steps))
