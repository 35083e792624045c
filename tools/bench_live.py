"""Developer tool: live-stream capacity of one GPU -- a cohort of cameras pushed k frames at a time (respmon_b200/live.py).
    python tools/bench_live.py [n_cameras] [k]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.engine import Engine
from respmon_b200.live import LiveCohort

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
k = int(sys.argv[2]) if len(sys.argv) > 2 else 8
W, H, T = 640, 480, 130 + 140 + 40 * max(8, k)
eng = Engine(0)
base = 16
specs = [synth.clip_spec(i, W, H, T) for i in range(base)]
dq8 = np.stack([synth.displacement_q8(s) for s in specs])
clips16 = eng.synth_clips(specs, dq8)                       # 16 distinct cameras, tiled to n
live = LiveCohort(n, W, H, 10.0, device=0, ring_len=33)
def block(lo, hi):
    return clips16[:, lo:hi].repeat((n + base - 1) // base, 1, 1, 1)[:n].contiguous()
pos = 0
while live.state != "measure" or live.n_measured < 140:     # steady state: full 128-sample windows
    live.push(block(pos, pos + 26)); pos += 26
torch.cuda.synchronize()
host = block(pos, pos + k).cpu().pin_memory()
for mode in ("device-resident frames", "host frames (H2D inside)"):
    ts = []
    for it in range(12):
        blk = block(pos, pos + k) if mode.startswith("device") else host
        torch.cuda.synchronize(); t0 = time.perf_counter()
        live.push(blk)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        pos += k
    ms = 1e3 * float(np.median(ts[2:]))
    print("%-28s %d cameras x %d frames per push: %.2f ms  -> %.0f frames/s, %.0f cameras at 10 frames/s"
          % (mode, n, k, ms, n * k / ms * 1e3, n * k / ms * 1e3 / 10))
print("ok cameras:", int((live.latest()["status"] == 0).sum()))
