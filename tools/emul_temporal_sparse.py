"""Thread-by-thread Python transliteration of temporal_sparse_kernel (respmon_b200/csrc/temporal.cu): same tables, same
tile / thread index arithmetic, compared with the oracle band-pass.  CPU only; checks indexing, not the CUDA build."""
import sys
sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import cpu_path as P
def emulate(x, fps, fmin, fmax, amp):
    T, Pn = x.shape
    lo, hi = P.temporal_bounds(T, fps, fmin, fmax)
    kept = []
    for j in range(T):
        k_ = True
        if hi > 0 and j >= hi and j < T - hi: k_ = False
        if lo != 0 and (j < lo or j >= T - lo): k_ = False
        if k_: kept.append(j)
    K = len(kept); assert 1 <= K <= 32 and T <= 256
    Tp = T + 2
    F = np.zeros(K*Tp); G = np.zeros(K*Tp)
    for e in range(K*T):
        a = e // T; t = e - a*T
        j = kept[a]; k = (j+1) >> 1; imag = (j != 0) and not (j & 1)
        m1 = (k*t) % T; m2 = (j*t) % T
        F[a*Tp+t] = -np.sin(2*np.pi*m1/T) if imag else np.cos(2*np.pi*m1/T)
        G[a*Tp+t] = np.cos(2*np.pi*m2/T)
    out = np.full_like(x, np.nan)
    tiles = (Pn + 31)//32
    inv_T = 1.0/T
    for tile in range(tiles):
        c0 = tile*32
        X = np.zeros(T*32)
        for i in range(T*32):
            t = i >> 5; c = i & 31
            X[i] = x[t, c0+c] if c0 + c < Pn else 0.0
        Q = np.full(32*32, np.nan)
        for tid in range(128):
            cg = tid & 7; rg = tid >> 3
            v0 = rg < K; v1 = rg + 16 < K
            f0 = (rg if v0 else 0)*Tp; f1 = ((rg+16) if v1 else 0)*Tp
            a0 = np.zeros(4); a1 = np.zeros(4)
            for t in range(T):
                xs = X[t*32 + 4*cg: t*32 + 4*cg + 4]
                a0 += F[f0+t]*xs; a1 += F[f1+t]*xs
            if v0: Q[rg*32+4*cg: rg*32+4*cg+4] = a0
            if v1: Q[(rg+16)*32+4*cg:(rg+16)*32+4*cg+4] = a1
        for tbase in range(0, T, 128):
            for tid in range(128):
                cg = tid & 7; rg = tid >> 3
                acc = np.zeros((8,4))
                for a in range(K):
                    q = Q[a*32+4*cg: a*32+4*cg+4]
                    for i in range(8):
                        t = tbase + rg + 16*i
                        gv = G[a*Tp + t] if t < T else 0.0
                        acc[i] += q*gv
                for i in range(8):
                    t = tbase + rg + 16*i
                    if t < T:
                        for c in range(4):
                            if c0 + 4*cg + c < Pn: out[t, c0+4*cg+c] = acc[i][c]*inv_T*amp
    return out, K
rng = np.random.default_rng(1)
for T, Pn, fps in [(128, 50, 10.0), (100, 33, 10.0), (64, 70, 7.5), (256, 40, 30.0)]:
    x = rng.standard_normal((T, Pn))
    try:
        got, K = emulate(x, fps, 0.1, 1.0, 500.0)
    except AssertionError:
        print(T, Pn, fps, "falls back (K or T out of range)"); continue
    ref = P.temporal_filter(x, fps, 0.1, 1.0, 500.0)
    print(T, Pn, fps, "K", K, "nan left:", int(np.isnan(got).sum()), "max diff", np.abs(got-ref).max())
