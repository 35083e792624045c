"""Developer tool: whole-clip parity of the CUDA path against the CPU oracle (oracle/cpu_path.run_clip, TEST INFRASTRUCTURE
used as the checker) on many synthetic clips per resolution -- ROI, per-clip status, the motion signal of the last window,
peak indices, the BPM history and the final BPM.     python tools/parity_sweep.py [clips per resolution]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiprocessing as mp
import numpy as np

T, FPS = 256, 10.0


def oracle(args):
    seed, w, h = args
    import cv2
    cv2.setNumThreads(1)
    from oracle import cpu_path as P
    from respmon_b200 import synth
    res = P.run_clip(synth.make_clip(synth.clip_spec(seed, w, h, T)), fps=FPS)
    return seed, res["roi"], res["bpm"], np.array(res["data"]), list(res["peaks"]), np.array(res["freq"])


def main():
    import torch
    from respmon_b200 import synth
    from respmon_b200.engine import Engine, results_to_numpy
    n_per = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    eng = Engine(0)
    workers = max(1, min(32, len(os.sched_getaffinity(0))))
    for (w, h, n) in ((640, 480, n_per), (320, 240, n_per), (1280, 720, max(4, n_per // 4)), (1920, 1080, max(2, n_per // 8))):
        seeds = [5000 + 17 * k for k in range(n)]
        specs = [synth.clip_spec(s, w, h, T) for s in seeds]
        clips = eng.synth_clips(specs, np.stack([synth.displacement_q8(s) for s in specs]))
        rec, taps = eng.run_batch(clips, FPS, keep=True)
        got = results_to_numpy(rec)
        data = taps["data"].cpu().numpy()
        bpm_hist = taps["bpm"].cpu().numpy()
        peaks = taps["peaks"].cpu().numpy()
        npk = taps["npeaks"].cpu().numpy()
        t0 = time.time()
        with mp.get_context("spawn").Pool(workers) as pool:
            ref = pool.map(oracle, [(s, w, h) for s in seeds], chunksize=1)
        roi_same = bpm_max = rms_max = peaks_same = hist_max = n_ok = 0
        for i, (seed, roi, bpm, d, pk, freq) in enumerate(ref):
            g = got[i]
            ok = roi is not None
            roi_same += (ok and int(g["status"]) in (0, 3, 4) and (int(g["x"]), int(g["y"]), int(g["w"]), int(g["h"])) == tuple(roi)) \
                or (not ok and int(g["status"]) == 1)
            if not ok:
                continue
            n_ok += 1
            if len(d):
                rms_max = max(rms_max, float(np.sqrt(np.nanmean((data[i][:len(d)] - d) ** 2))))
            if bpm is not None and not np.isnan(g["bpm"]):
                bpm_max = max(bpm_max, abs(float(g["bpm"]) - bpm))
            peaks_same += list(peaks[i][:npk[i]]) == [int(p) for p in pk]
            h_gpu = bpm_hist[i][~np.isnan(bpm_hist[i])]
            if len(h_gpu) == len(freq) and len(freq):
                hist_max = max(hist_max, float(np.abs(h_gpu - freq).max()))
            elif len(h_gpu) != len(freq):
                hist_max = float("inf")
        print("%4dx%-4d %3d clips: ROI identical %d/%d, clips with ROI %d, peak lists identical %d/%d, max |dBPM| final %.2e, "
              "history %.2e, max signal RMS %.2e   (oracle %.0f s on %d processes)" % (
                  w, h, n, roi_same, n, n_ok, peaks_same, n_ok, bpm_max, hist_max, rms_max, time.time() - t0, workers), flush=True)


if __name__ == "__main__":
    main()
