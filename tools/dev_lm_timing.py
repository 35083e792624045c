"""Developer tool: per-section cycle counts of the LM fit (needs a build with RM_NVCC_EXTRA=-DLM_TIMING)."""
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
rec, taps = eng.run_batch(clips, 10.0, keep=True)
torch.cuda.synchronize()
out = (C.c_ulonglong * 8)()
lib.rm_debug_lm_timing(out, 1)
n_sub = int(sys.argv[1]) if len(sys.argv) > 1 else 64
d = taps["data"][:n_sub].contiguous()
eng.signal_bpm(d, 10.0)
torch.cuda.synchronize()
lib.rm_debug_lm_timing(out, 0)
v = list(out)
names = ["fdjac(3 resid)", "qrfac", "qtf+R+gnorm", "lmpar", "trial resid+ratio", "loop top"]
outer, inner = v[7], v[6]
print("outer iterations", outer, "inner (lmpar calls)", inner, "clips", n_sub)
tot = sum(v[:6])
for n, c in zip(names, v[:6]):
    print("  %-20s %6.1f%%  %8.0f cycles per outer iteration" % (n, 100.0 * c / tot, c / max(outer, 1)))
print("  total %.0f cycles per outer iteration (groups share warps: wall cycles include the other groups' turns)" % (tot / max(outer, 1)))
