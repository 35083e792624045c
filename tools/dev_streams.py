"""Developer tool: K steps of run_batch at the bench shape, alternating over S streams (one rm_handle per stream), against
one stream.  python tools/dev_streams.py [n_clips] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.engine import Engine, RESULT_DTYPE
n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
eng0 = Engine(0)
specs = [synth.clip_spec(i, 640, 480, 256) for i in range(n_clips)]
dq8 = np.stack([synth.displacement_q8(s) for s in specs])
clips = eng0.synth_clips(specs, dq8)
base = None
for S in (1, 2, 3, 4):
    engs = [eng0] + [Engine(0) for _ in range(S - 1)]
    streams = [torch.cuda.Stream() for _ in range(S)]
    recs = [torch.empty((n_clips, RESULT_DTYPE.itemsize), dtype=torch.uint8, device="cuda") for _ in range(S)]
    def run(n):
        for k in range(n):
            with torch.cuda.stream(streams[k % S]):
                engs[k % S].run_batch(clips, 10.0, out=recs[k % S])
    run(2 * S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record()
    for s in streams:
        s.wait_stream(cur)
    run(steps)
    for s in streams:
        cur.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    r = recs[0].cpu().numpy()
    same = "-" if base is None else bool(np.array_equal(r, base))
    base = r if base is None else base
    print("%d clips per step, %d stream(s): %.3f ms per step, %.3f M frames/s, records identical: %s" % (
        n_clips, S, ms, n_clips * 256 / ms / 1e3, same), flush=True)
