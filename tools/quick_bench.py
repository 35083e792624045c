"""Per-kernel timing at BASELINE config-2 shapes (developer tool; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from respmon_b200.engine import Engine
from respmon_b200 import synth

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), sum(ts)/len(ts)

eng=Engine(0)
n_clips=int(sys.argv[1]) if len(sys.argv)>1 else 32
T=256; W,H=640,480
specs=[synth.clip_spec(s,W,H,T) for s in range(n_clips)]
dq8=np.stack([synth.displacement_q8(s) for s in specs])
clips=eng.synth_clips(specs,dq8); torch.cuda.synchronize()
nf=n_clips*T
for name,x in (("u8",clips),("f32",None)):
    if x is None:
        x=(clips[:n_clips//4].float()*(1/255)).contiguous(); 
    nfx=x.shape[0]*T
    mn,av=timeit(lambda: eng.pyramid_build(x))
    bytes_=nfx*(W*H*x.element_size()+1600*8)
    print(f"pyramid_build {name}: {mn:.3f} ms min / {av:.3f} avg  -> {nfx/mn*1e3/1e6:.3f} Mframes/s  {bytes_/mn/1e6:.1f} GB/s algorithmic (front+tail)")
lap=eng.pyramid_build(clips)
mn,av=timeit(lambda: eng.temporal_bandpass(lap,10.0))
print(f"temporal: {mn:.3f} ms -> {nf/mn*1e3/1e6:.2f} Mframes/s")
bp=eng.temporal_bandpass(lap,10.0)
mn,av=timeit(lambda: eng.heatmap(bp,W,H),n=3,warm=1)
print(f"heatmap: {mn:.3f} ms -> {nf/mn*1e3/1e6:.3f} Mframes/s")
mn,av=timeit(lambda: eng.calibrate_heatmaps(clips,10.0),n=3,warm=1)
print(f"calibrate total: {mn:.3f} ms -> {nf/mn*1e3/1e6:.3f} Mframes/s")
