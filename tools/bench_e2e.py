"""Developer tool: end-to-end (host memory) throughput of BatchMonitor.run for a few pipeline settings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200 import synth
from respmon_b200.batch import BatchMonitor
from respmon_b200.engine import Engine

n_clips = 64
eng = Engine(0)
specs = [synth.clip_spec(i, 640, 480, 256) for i in range(n_clips)]
dq8 = np.stack([synth.displacement_q8(s) for s in specs])
clips = eng.synth_clips(specs, dq8)
host = torch.empty(clips.shape, dtype=torch.uint8).pin_memory()
host.copy_(clips); torch.cuda.synchronize()
del clips
ref = None
for chunk, streams, mch in [(32, 2, 4), (32, 2, 1), (32, 3, 2), (16, 2, 4), (64, 2, 4)]:
    mon = BatchMonitor(0, chunk_clips=chunk, measure_streams=streams)
    for e in mon._measure_engines:
        e.set_option("measure_chunks", mch)
    out = mon.run(host, 10.0)
    torch.cuda.synchronize()
    steps = 12
    t0 = time.perf_counter()
    prev = None
    for _ in range(steps):                      # pipelined like bench.py's e2e leg
        ticket = mon.submit(host, 10.0)
        if prev is not None:
            out = mon.collect(prev)
        prev = ticket
    out = mon.collect(prev)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    if ref is None: ref = out
    same = all(np.array_equal(out[f], ref[f], equal_nan=True) for f in out.dtype.names)
    print("chunk %2d streams %d measure_chunks %d : %.1f ms/step  %.0f frames/s  same=%s" % (chunk, streams, mch, dt * 1e3, n_clips * 256 / dt, same), flush=True)
    del mon
