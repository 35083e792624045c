"""Developer tool: the heat-map bytes (base.py:562-564: time mean, min-max normalise, *255 truncated to uint8) that differ
from the unmodified reference's on every golden fixture -- how many, where, by how much, and how far the pre-truncation
value is from an integer -- for the CPU oracle and for the CUDA path (FFT and sparse band-pass).  The truncation is
discontinuous: a value like 37.9999999999 vs 38.0000000001 lands on different bytes, so "differs by 1 on a handful of
pixels" is the expected signature of float64 rounding in an equivalent evaluation order, and none may cross the threshold.
    python tools/heat_flips.py            (GPU box: oracle + CUDA;  without a GPU: oracle only)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cpu_path as P
from respmon_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def describe(tag, heat, ref):
    d = heat.astype(int) - ref.astype(int)
    ys, xs = np.nonzero(d)
    cross = int(np.count_nonzero((heat > P.THRESHOLD) != (ref > P.THRESHOLD)))
    where = ", ".join("(%d,%d): %d->%d" % (x, y, ref[y, x], heat[y, x]) for y, x in list(zip(ys, xs))[:8])
    print("  %-28s %d of %d bytes differ (max |d| %d, %d across the threshold)  %s" % (
        tag, len(ys), heat.size, int(np.abs(d).max()) if len(ys) else 0, cross, where))


def main():
    try:
        import torch
        gpu = torch.cuda.is_available()
    except Exception:
        gpu = False
    eng = None
    if gpu:
        from respmon_b200.engine import Engine
        eng = Engine(0)
    for name in ("vga_s0", "vga_s2", "qvga_s1", "odd_s3", "qvga_long_s4"):
        fix = np.load(os.path.join(GOLD, name + ".npz"))
        W, H, T, seed = (int(v) for v in fix["spec"][:4])
        clip = synth.make_clip(synth.clip_spec(seed, W, H, T))
        ref = fix["heat_u8"]
        print("%s (%dx%d), ROI %s" % (name, W, H, tuple(int(v) for v in fix["roi"])))
        taps = {}
        P.locate(P.u8_to_unit(clip[1:129]), float(fix["fps"]), taps=taps)
        describe("CPU oracle", taps["heat_u8"], ref)
        if eng is not None:
            import torch
            d = torch.from_numpy(clip[None, 1:129]).cuda()
            for sparse in (1, 0):
                eng.set_option("temporal_sparse", sparse)
                heat, _ = eng.calibrate_heatmaps(d, float(fix["fps"]))
                describe("CUDA, %s band-pass" % ("sparse" if sparse else "FFT"), heat.cpu().numpy()[0], ref)
            eng.set_option("temporal_sparse", 1)


if __name__ == "__main__":
    main()
