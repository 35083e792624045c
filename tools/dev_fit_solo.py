"""Developer tool (untimed experiment): the Gaussian-fit options against the default, at the bench shapes --
"fit_bail_nfev" (the first pass gives up on a fit after N evaluations, a second pass runs those fits one per warp),
"fit_blocks_per_sm" (fewer resident fit blocks) and "fit_sync" (warp-synchronous first pass).
    python tools/dev_fit_solo.py [n_clips] [steps]
Prints the step time per setting and checks that the result records and the per-frame BPM history do not change."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.engine import Engine, results_to_numpy

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
eng = Engine(0)
specs = [synth.clip_spec(i, 640, 480, 256) for i in range(n_clips)]
dq8 = np.stack([synth.displacement_q8(s) for s in specs])
clips = eng.synth_clips(specs, dq8)
base = None
for bail, per_sm, sync in ((0, 0, 0), (0, 1, 0), (40, 0, 0), (60, 0, 0), (100, 0, 0), (200, 0, 0), (100, 1, 0),
                           (0, 0, 1), (60, 0, 1), (100, 0, 1), (200, 0, 1), (100, 1, 1)):
    eng.set_option("fit_bail_nfev", bail)
    eng.set_option("fit_blocks_per_sm", per_sm)
    eng.set_option("fit_sync", sync)
    rec, taps = eng.run_batch(clips, 10.0, keep=True)
    for _ in range(2):
        eng.run_batch(clips, 10.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.run_batch(clips, 10.0)
    e1.record()
    torch.cuda.synchronize()
    r = results_to_numpy(rec)
    bpm = taps["bpm"].cpu().numpy()
    cur = (r["bpm"].copy(), r["status"].copy(), r["n_peaks"].copy(), bpm.copy())
    same = "-"
    if base is None:
        base = cur
    else:
        same = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(base, cur))
    print("fit_bail_nfev %3d, fit_blocks_per_sm %d, fit_sync %d: step %.3f ms   records and BPM history identical to default: %s" % (
        bail, per_sm, sync, e0.elapsed_time(e1) / steps, same))
