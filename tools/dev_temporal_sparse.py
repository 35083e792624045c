"""Developer tool (untimed experiment): option "temporal_sparse" (band-pass through the kept bins only) against the
default FFT kernel: per-kernel time at the bench shape, output difference, bit-identity with the any-T direct kernel, and
whether a whole batch still yields the same records.
    python tools/dev_temporal_sparse.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.engine import Engine, results_to_numpy

eng = Engine(0)
g = torch.Generator(device="cuda").manual_seed(0)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (n, T, P, fps) in [(64, 128, 1600, 10.0), (8, 100, 1600, 10.0), (8, 128, 331, 30.0), (4, 256, 400, 30.0)]:
    lap = torch.randn((n, T, P), dtype=torch.float64, device="cuda", generator=g)
    eng.set_option("temporal_sparse", 0)
    ref = eng.temporal_bandpass(lap, fps)
    t_ref = timed(lambda: eng.temporal_bandpass(lap, fps))
    eng.set_option("temporal_sparse", 1)
    got = eng.temporal_bandpass(lap, fps)
    t_new = timed(lambda: eng.temporal_bandpass(lap, fps))
    d = (got - ref).abs().max().item()
    print("n=%d T=%d P=%d fps=%g: default %.3f ms, sparse %.3f ms, max |diff| %.3e (scale %.3e), bit-identical %s" % (
        n, T, P, fps, t_ref, t_new, d, ref.abs().max().item(), bool(torch.equal(got, ref))))

specs = [synth.clip_spec(i, 640, 480, 256) for i in range(64)]
dq8 = np.stack([synth.displacement_q8(s) for s in specs])
clips = eng.synth_clips(specs, dq8)
out = []
for on in (0, 1):
    eng.set_option("temporal_sparse", on)
    rec = eng.run_batch(clips, 10.0)
    torch.cuda.synchronize()
    out.append(results_to_numpy(rec).copy())
    print("temporal_sparse=%d: step %.3f ms" % (on, timed(lambda: eng.run_batch(clips, 10.0), 5)))
same_roi = all(np.array_equal(out[0][k], out[1][k]) for k in ("x", "y", "w", "h", "status"))
print("ROI / status identical:", same_roi, " max |dBPM|:", float(np.nanmax(np.abs(out[0]["bpm"] - out[1]["bpm"]))))
