"""Developer tool: the uint8 pyramid stage (frames in -> packed Laplacian records out) per mode at the bench shape.
    python tools/bench_pyramid.py [W H n_clips]
"split" = level 3 through HBM + pyramid_tail_kernel (the fallback), "fused c" = the one-pass TMA kernel in ring /
occupancy configuration c (0: chosen by frame width, 1: 4 stages x 18 warps, 2: 2 x 18, 3: 2 x 24).
Prints the stage time from CUDA events, the SURVEY 8(d) algorithmic bytes (W*H*1 + record*8 per frame) over it, and
whether the records equal mode 0's bit for bit."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.engine import Engine

W = int(sys.argv[1]) if len(sys.argv) > 1 else 640
H = int(sys.argv[2]) if len(sys.argv) > 2 else 480
n_clips = int(sys.argv[3]) if len(sys.argv) > 3 else 64
T, first, length = 256, 1, 128
eng = Engine(0)
specs = [synth.clip_spec(i, W, H, T) for i in range(n_clips)]
dq8 = np.stack([synth.displacement_q8(s) for s in specs])
clips = eng.synth_clips(specs, dq8)
rec_len = eng.record_len(W, H)
peak = 6451.5
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
base = None
for mode in [int(m) for m in os.environ.get('RM_MODES', '-1,0,1,2,3').split(',')]:
    eng.set_option("pyramid_mode", 0 if mode < 0 else 1)
    eng.set_option("pyramid_cfg", max(mode, 0))
    eng.set_option("pyramid_g4", int(os.environ.get("RM_G4", "0")))
    out = eng.pyramid_build_clips(clips, first, length)
    for _ in range(3):
        eng.pyramid_build_clips(clips, first, length)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.pyramid_build_clips(clips, first, length)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    nbytes = n_clips * length * (W * H + rec_len * 8)
    same = "-" if base is None else bool(torch.equal(out, base))
    base = out.clone() if base is None else base
    print("%dx%d x%d clips, %s: stage %.3f ms (min %.3f)  %.0f GB/s = %.3f of %.1f   identical to split: %s" % (
        W, H, n_clips, "split" if mode < 0 else "fused %d" % mode, ms, min(ts), nbytes / ms / 1e6, nbytes / ms / 1e6 / peak, peak, same))
