"""Many independent clips through calibrate + measure: the batch form of RespiratoryMonitor.run() (base.py:409-513).

`BatchMonitor.run(clips, fps)` takes HOST clips (n, T, H, W) uint8, streams them to the GPU in chunks on a copy
stream while the previous chunk is being processed (two device buffers), and returns one 32-byte result record per
clip (`engine.RESULT_DTYPE`).  Under torch.distributed every rank processes its own shard of clips; the only
collective is one all-gather of the result records at the end.
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import RESULT_DTYPE, Engine


def shard_range(n_clips: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of clips owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balance_clips(shapes, world: int) -> list[list[int]]:
    """Mixed-resolution batches (BASELINE config 5): greedy longest-first assignment of clips to ranks by T*H*W so
    every GPU streams about the same number of pixels.  shapes: [(T,H,W)] -> per-rank lists of clip indices (sorted)."""
    cost = [int(t) * int(h) * int(w) for t, h, w in shapes]
    order = sorted(range(len(shapes)), key=lambda i: (-cost[i], i))
    load = [0] * world
    owners = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owners[r].append(i)
        load[r] += cost[i]
    return [sorted(o) for o in owners]


def gather_records(local: np.ndarray, counts: list[int] | None = None) -> np.ndarray:
    """All-gather of per-clip result records over the default process group (no-op without one).

    `local` is a RESULT_DTYPE array; ranks may hold different numbers of clips (`counts`, else equal)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if counts is None:
        counts = [len(local)] * world
    cap = max(counts)
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros((cap, RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    if len(local):
        buf[:len(local)] = torch.from_numpy(local.view(np.uint8).reshape(len(local), -1).copy()).to(dev)
    out = torch.empty((world * cap, RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape(world, cap, RESULT_DTYPE.itemsize)
    return np.concatenate([out[r, :counts[r]].reshape(-1).view(RESULT_DTYPE) for r in range(world)])


class BatchMonitor:
    """Calibrate + measure for a batch of whole clips held in host memory."""

    def __init__(self, device: int | None = None, chunk_clips: int = 8, method: str = "flow", **hyper):
        self.engine = Engine(device, **hyper)
        self.chunk_clips = int(chunk_clips)
        self.method = method
        self._bufs = [None, None]
        self._copy_stream = torch.cuda.Stream(self.engine.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _buffer(self, slot, shape):
        b = self._bufs[slot]
        if b is None or b.shape[1:] != shape[1:] or b.shape[0] < shape[0]:
            b = torch.empty(shape, dtype=torch.uint8, device=self.engine.device)
            self._bufs[slot] = b
        return b[:shape[0]]

    def run(self, clips, fps: float, cal_first: int = 1, cal_len: int = 128) -> np.ndarray:
        """clips: (n,T,H,W) uint8 numpy array or (preferably pinned) CPU tensor -> RESULT_DTYPE array of n records."""
        host = torch.from_numpy(clips) if isinstance(clips, np.ndarray) else clips
        assert host.dtype == torch.uint8 and host.dim() == 4 and not host.is_cuda
        n = host.shape[0]
        eng = self.engine
        compute = torch.cuda.current_stream(eng.device)
        chunks = [(lo, min(n, lo + self.chunk_clips)) for lo in range(0, n, self.chunk_clips)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]
        records = torch.empty((n, RESULT_DTYPE.itemsize), dtype=torch.uint8, device=eng.device)

        def upload(i):
            lo, hi = chunks[i]
            slot = i & 1
            dst = self._buffer(slot, (hi - lo,) + tuple(host.shape[1:]))
            with torch.cuda.stream(self._copy_stream):
                if i >= 2:
                    self._copy_stream.wait_event(freed[slot])     # the chunk that used this buffer has been processed
                dst.copy_(host[lo:hi], non_blocking=True)
                ready[slot].record(self._copy_stream)
            self.h2d_bytes += dst.numel()
            return dst

        pending = upload(0) if chunks else None
        for i, (lo, hi) in enumerate(chunks):
            cur = pending
            if i + 1 < len(chunks):
                pending = upload(i + 1)                            # overlaps the processing of chunk i
            compute.wait_event(ready[i & 1])
            eng.run_batch(cur, fps, cal_first=cal_first, cal_len=cal_len, method=self.method, out=records[lo:hi])
            freed[i & 1].record(compute)
        out = records.cpu().numpy().view(RESULT_DTYPE).reshape(-1)   # the device->host read of the step's result
        self.d2h_bytes += records.numel()
        return out

    def run_mixed(self, clips: list, fps: float, cal_first: int = 1, cal_len: int = 128) -> np.ndarray:
        """Ragged batch (BASELINE config 5): clips is a list of (T,H,W) uint8 arrays of different sizes.  Clips are
        grouped into resolution classes (one kernel configuration and one set of level sizes per class) and every
        class goes through `run`; records come back in the order the clips were given."""
        classes = {}
        for i, c in enumerate(clips):
            classes.setdefault(tuple(c.shape), []).append(i)
        out = np.zeros(len(clips), RESULT_DTYPE)
        for shape, idx in sorted(classes.items()):
            if isinstance(clips[idx[0]], np.ndarray):
                stack = np.stack([clips[i] for i in idx])
            else:
                stack = torch.stack([clips[i] for i in idx])
            out[idx] = self.run(stack, fps, cal_first=cal_first, cal_len=cal_len)
        return out
