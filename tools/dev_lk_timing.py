"""Developer tool: per-phase cycle counts of the LK tracker (needs a build with RM_NVCC_EXTRA=-DLK_TIMING)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200 import synth, _cabi
from respmon_b200.engine import Engine
eng = Engine(0)
lib = C.CDLL(_cabi.LIB_PATH)
specs = [synth.clip_spec(i, 640, 480, 256) for i in range(64)]
dq8 = np.stack([synth.displacement_q8(s) for s in specs])
clips = eng.synth_clips(specs, dq8)
eng.set_option("measure_chunks", 1)
rec, taps = eng.run_batch(clips, 10.0, keep=True)
torch.cuda.synchronize()
out = (C.c_longlong * (64 * 8 * 8))()
lib.rm_debug_lk_timing(out, 1)
rec = eng.run_batch(clips, 10.0)
torch.cuda.synchronize()
lib.rm_debug_lk_timing(out, 0)
a = np.array(out[:]).reshape(64, 8, 8)
tot = a[:, :, :4].sum(-1)
order = np.dstack(np.unravel_index(np.argsort(-tot, axis=None), tot.shape))[0][:10]
npts = taps["npts"].cpu().numpy()
print("clip blk  npts_clip  roi    total_cycles  wait build track book   (cycles/frame)")
for c, b in order:
    r = a[c, b]; nf = max(r[4], 1)
    print("%3d %2d  %3d  %3dx%-3d  %9d   %6d %6d %6d %6d" % (c, b, npts[c], r[6], r[7], tot[c, b], r[0] / nf, r[1] / nf, r[2] / nf, r[3] / nf))
print("active blocks", int((tot > 0).sum()), "median total", np.median(tot[tot > 0]))

sub = a[63, 7, :5].astype(float) / max(a[22, 0, 4], 1)
print("build sub-phases of clip 22 blk 0 (cycles/frame): L0 fill %.0f, L1 down %.0f, L1 border %.0f, L2 down %.0f, L2 border %.0f" % tuple(sub))
