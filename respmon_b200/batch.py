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
    """Calibrate + measure for a batch of whole clips held in host memory.

    Only the bytes the path reads cross PCIe: the calibration window of every clip (frames cal_first ..
    cal_first+cal_len-1, base.py:429-434) goes up first; once locate() has produced the ROI, the measure frames go up
    as ROI crops -- the reference itself only ever looks at `frame[y:y+h, x:x+w]` of those frames (base.py:471).
    Uploads run on a copy stream and overlap the kernels of the previous chunk (two buffers of each kind).

    Clips in PINNED host memory take the mapped path (`mapped=True`): the crop kernel reads the ROI's rows of the measure
    frames straight out of the host buffer (pinned memory is device-addressable), so the ROI never visits the host, no
    staging copy is made and submit() does not wait for anything -- the host -> device queue never drains between chunks
    or batches.  The crop tensors are sized by `roi_cap` (w, h), which grows to the largest ROI seen; a batch in which
    an ROI exceeded it is noticed in collect() and its chunk re-run through the staged path."""

    def __init__(self, device: int | None = None, chunk_clips: int = 32, method: str = "flow", crop_upload: bool = True,
                 measure_streams: int = 2, measure_chunks: int = 4, mapped: bool = True, roi_cap: tuple = (64, 64), **hyper):
        self.engine = Engine(device, **hyper)
        self._measure_engines = [Engine(self.engine.device_index, **hyper) for _ in range(max(1, measure_streams))]
        for e in self._measure_engines:
            e.set_option("measure_chunks", measure_chunks)
        self._measure_streams = [torch.cuda.Stream(self.engine.device) for _ in self._measure_engines]
        self._crop_stream = torch.cuda.Stream(self.engine.device)
        self.chunk_clips = int(chunk_clips)
        self.method = method
        self.crop_upload = bool(crop_upload)
        self.mapped = bool(mapped)
        self.roi_cap = (int(roi_cap[0]), int(roi_cap[1]))
        self.reruns = 0                  # chunks re-run because an ROI exceeded roi_cap
        self._bufs = {}
        self._pinned = {}
        self._ev = None
        self._copy_stream = torch.cuda.Stream(self.engine.device)
        self._read_stream = torch.cuda.Stream(self.engine.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _buffer(self, key, shape):
        """Device uint8 buffer of at least prod(shape) bytes, viewed as `shape`."""
        n = int(np.prod(shape))
        b = self._bufs.get(key)
        if b is None or b.numel() < n:
            if b is not None:
                torch.cuda.synchronize(self.engine.device)   # an earlier batch may still be using the buffer being replaced
            b = torch.empty(max(n, 1), dtype=torch.uint8, device=self.engine.device)
            self._bufs[key] = b
        return b[:n].view(shape)

    def _pinned_buffer(self, key, shape, dtype=torch.uint8):
        n = int(np.prod(shape))
        b = self._pinned.get(key)
        if b is None or b.numel() < n or b.dtype != dtype:
            if b is not None:
                torch.cuda.synchronize(self.engine.device)   # a copy out of the buffer being replaced may be in flight
            b = torch.empty(max(n, 1), dtype=dtype).pin_memory()
            self._pinned[key] = b
        return b[:n].view(shape)

    def run(self, clips, fps: float, cal_first: int = 1, cal_len: int = 128) -> np.ndarray:
        """clips: (n,T,H,W) uint8 numpy array or (preferably pinned) CPU tensor -> RESULT_DTYPE array of n records.
        submit() followed by collect(); use the two directly to overlap consecutive batches."""
        return self.collect(self.submit(clips, fps, cal_first, cal_len))

    def _slot_events(self):
        if self._ev is None:
            n_ms = len(self._measure_streams)
            mk = lambda k: [torch.cuda.Event() for _ in range(k)]          # noqa: E731
            self._ev = dict(cal_ready=mk(2), cal_freed=mk(2), roi_done=mk(2), crop_ready=mk(n_ms), stage_freed=mk(n_ms),
                            crop_dev_freed=mk(n_ms))
            self._used = dict(cal=[False, False], ms=[False] * n_ms)
            self._seq = 0                                                  # chunks submitted so far (slot rotation)
        return self._ev

    def _upload_cal(self, host, chunk, slot, cal_first, cal_len, H, W):
        """Calibration windows of the clips [lo, hi) -> device buffer `slot` (of two), on the copy stream."""
        lo, hi = chunk
        E, used, copy = self._slot_events(), self._used, self._copy_stream
        dst = self._buffer(("cal", slot), (hi - lo, cal_len, H, W))
        with torch.cuda.stream(copy):
            if used["cal"][slot]:
                copy.wait_event(E["cal_freed"][slot])             # locate() of the chunk that last used this buffer is done
            for c in range(lo, hi):                               # one contiguous block per clip
                dst[c - lo].copy_(host[c][cal_first:cal_first + cal_len], non_blocking=True)
            E["cal_ready"][slot].record(copy)
        used["cal"][slot] = True
        self.h2d_bytes += dst.numel()
        return dst

    def submit(self, clips, fps: float, cal_first: int = 1, cal_len: int = 128):
        """Enqueue a batch and return a ticket for collect().  Returns as soon as the last chunk's measure stage has been
        enqueued, so the upload of the next submit() overlaps the tail of this one (buffers and streams are handed from
        batch to batch through events; the results of every batch are complete when its collect() returns).

        Per chunk: calibration window H2D (copy stream) -> locate() (calibrate stream) -> ROI to the host -> ROI crops
        of the measure frames H2D (second copy stream) -> LK measure + BPM on one of `measure_streams` streams.  The
        measure kernels are latency bound (frames are sequential), so consecutive chunks run them
        concurrently on different streams, each through its own rm_handle (handles own their scratch)."""
        if isinstance(clips, (list, tuple)):
            # clips held one by one (e.g. the members of one resolution class of a ragged batch): no stacked copy is made,
            # every clip's calibration window and ROI crops go up from where the clip lies (pinned or pageable)
            host = [torch.from_numpy(c) if isinstance(c, np.ndarray) else c for c in clips]
            assert all(c.dtype == torch.uint8 and c.dim() == 3 and not c.is_cuda and c.is_contiguous() and
                       c.shape == host[0].shape for c in host)
            n, (T, H, W) = len(host), tuple(host[0].shape)
            host_np = [c.numpy() for c in host]
        else:
            host = torch.from_numpy(clips) if isinstance(clips, np.ndarray) else clips
            assert host.dtype == torch.uint8 and host.dim() == 4 and not host.is_cuda and host.is_contiguous()
            n, T, H, W = host.shape
            host_np = host.numpy()
        measure_first = cal_first + cal_len + 1
        n_meas = T - measure_first
        assert cal_first >= 0 and n_meas >= 1
        if not self.crop_upload:
            if isinstance(host, list):
                host = torch.stack(host)
            return dict(done=self._run_full_frames(host, fps, cal_first, cal_len))
        if self.mapped and (all(c.is_pinned() for c in host) if isinstance(host, list) else host.is_pinned()):
            return self._submit_mapped(host, n, T, H, W, fps, cal_first, cal_len)
        eng = self.engine
        dev = eng.device
        main = torch.cuda.current_stream(dev)
        copy, copy2 = self._copy_stream, self._crop_stream
        n_ms = len(self._measure_streams)
        chunks = self._chunk_schedule(n)
        E = self._slot_events()
        used = self._used
        records = torch.empty((n, RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
        keep = [host]                                                  # tensors shared across streams stay alive
        seq0 = self._seq

        def upload_cal(i):
            return self._upload_cal(host, chunks[i], (seq0 + i) & 1, cal_first, cal_len, H, W)

        pending = upload_cal(0) if chunks else None
        for i, (lo, hi) in enumerate(chunks):
            slot, ms = (seq0 + i) & 1, (seq0 + i) % n_ms
            cal = pending
            if i + 1 < len(chunks):
                pending = upload_cal(i + 1)                       # overlaps everything below
            m = hi - lo
            main.wait_event(E["cal_ready"][slot])
            roi, status, _ = eng.locate(cal, fps, 0, cal_len)
            E["cal_freed"][slot].record(main)
            roi_host = self._pinned_buffer(("roi", slot), (m, 4), torch.int32)
            st_host = self._pinned_buffer(("st", slot), (m,), torch.int32)
            roi_host.copy_(roi, non_blocking=True)
            st_host.copy_(status, non_blocking=True)
            E["roi_done"][slot].record(main)
            self.d2h_bytes += roi_host.numel() * 4 + st_host.numel() * 4
            E["roi_done"][slot].synchronize()                     # the ROI decides which bytes go up next
            r = roi_host.numpy()
            ok = st_host.numpy() == 0
            mw = int(max(1, r[ok, 2].max())) if ok.any() else 1
            mh = int(max(1, r[ok, 3].max())) if ok.any() else 1
            stage = self._pinned_buffer(("crop", ms), (m, n_meas, mh, mw))
            if used["ms"][ms]:
                E["stage_freed"][ms].synchronize()                # the H2D that last used this staging area has left it
            sv = stage.numpy()
            for c in range(m):
                if ok[c]:
                    x, y, w, h = (int(v) for v in r[c])
                    sv[c, :, :h, :w] = host_np[lo + c][measure_first:, y:y + h, x:x + w]   # base.py:471
            crops = self._buffer(("cropdev", ms), (m, n_meas, mh, mw))
            with torch.cuda.stream(copy2):
                if used["ms"][ms]:
                    copy2.wait_event(E["crop_dev_freed"][ms])     # the measure stage that last read this buffer is done
                crops.copy_(stage, non_blocking=True)
                E["crop_ready"][ms].record(copy2)
                E["stage_freed"][ms].record(copy2)
            used["ms"][ms] = True
            self.h2d_bytes += crops.numel()
            mstream, meng = self._measure_streams[ms], self._measure_engines[ms]
            with torch.cuda.stream(mstream):
                mstream.wait_event(E["roi_done"][slot])
                mstream.wait_event(E["crop_ready"][ms])
                roi0 = roi.clone()
                roi0[:, :2] = 0                                   # the crop's own origin
                st = status.clone()
                if self.method == "flow":
                    sig = meng.measure_signal(crops, roi0, 0, n_meas, fps, status=st, max_roi=(mw, mh))
                    data = sig["data"]
                else:
                    data = meng.measure_average(crops, roi0, 0, n_meas)
                    sig = meng.signal_bpm(data, fps, status=st)
                meng.pack_results(sig["bpm"], roi, st, sig["npeaks"], out=records[lo:hi])
                E["crop_dev_freed"][ms].record(mstream)
            keep.append((roi, status, roi0, st, data, sig))
        self._seq = seq0 + len(chunks)
        done = []
        for mstream in self._measure_streams:
            e = torch.cuda.Event()
            e.record(mstream)
            done.append(e)
        return dict(records=records, done_events=done, keep=keep)

    def collect(self, ticket) -> np.ndarray:
        """Wait for a submitted batch and read its records back (the device->host read of the step's result)."""
        if "done" in ticket:
            return ticket["done"]
        if "mapped" in ticket:
            return self._collect_mapped(ticket)
        main = torch.cuda.current_stream(self.engine.device)
        for e in ticket["done_events"]:
            main.wait_event(e)
        out = ticket["records"].cpu().numpy().view(RESULT_DTYPE).reshape(-1)
        self.d2h_bytes += ticket["records"].numel()
        ticket["keep"] = None
        return out

    # ------------------------------------------------------------------ mapped path (clips in pinned host memory)
    def _submit_mapped(self, host, n, T, H, W, fps, cal_first, cal_len):
        """Per chunk: calibration window H2D (copy stream) -> locate() (caller's stream) -> on one of the measure streams:
        crop kernel reading the ROI's rows of the measure frames from the mapped host buffer, LK measure + BPM, records.
        Nothing here waits on the host: buffers pass from chunk to chunk and batch to batch through events."""
        from . import _cabi
        from .engine import RM_U8, _ptr
        import ctypes as C
        eng = self.engine
        dev = eng.device
        main = torch.cuda.current_stream(dev)
        n_ms = len(self._measure_streams)
        measure_first = cal_first + cal_len + 1
        n_meas = T - measure_first
        chunks = self._chunk_schedule(n)
        E = self._slot_events()
        records = torch.empty((n, RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
        keep = [host]
        seq0 = self._seq
        cap_w, cap_h = min(W, self.roi_cap[0]), min(H, self.roi_cap[1])
        as_list = isinstance(host, list)

        def upload_cal(i):
            return self._upload_cal(host, chunks[i], (seq0 + i) & 1, cal_first, cal_len, H, W)

        pending = upload_cal(0) if chunks else None
        for i, (lo, hi) in enumerate(chunks):
            slot, ms = (seq0 + i) & 1, (seq0 + i) % n_ms
            cal = pending
            if i + 1 < len(chunks):
                pending = upload_cal(i + 1)
            m = hi - lo
            main.wait_event(E["cal_ready"][slot])
            roi, status, _ = eng.locate(cal, fps, 0, cal_len)
            E["cal_freed"][slot].record(main)
            E["roi_done"][slot].record(main)
            mstream, meng = self._measure_streams[ms], self._measure_engines[ms]
            with torch.cuda.stream(mstream):
                mstream.wait_event(E["roi_done"][slot])
                if as_list:
                    descs = (_cabi.RmClipDesc * m)()
                    for k in range(m):
                        descs[k] = _cabi.RmClipDesc(host[lo + k].data_ptr(), W, H, T, W)   # base = NULL: absolute addresses
                    d_descs = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(dev, non_blocking=True)
                    crops = torch.zeros((m, n_meas, cap_h, cap_w), dtype=torch.uint8, device=dev)
                    meng._call("rm_crop_frames_ragged", C.c_void_p(0), RM_U8, _ptr(d_descs), m, _ptr(roi), measure_first,
                               n_meas, _ptr(crops), cap_w, cap_h, meng._stream())
                    keep.append(d_descs)
                else:
                    crops = meng.crop_frames(host[lo:hi], roi, measure_first, n_meas, out_size=(cap_w, cap_h))
                roi0 = roi.clone()
                roi0[:, :2] = 0                                   # the crop's own origin
                st = status.clone()
                if self.method == "flow":
                    sig = meng.measure_signal(crops, roi0, 0, n_meas, fps, status=st, max_roi=(cap_w, cap_h))
                    data = sig["data"]
                else:
                    data = meng.measure_average(crops, roi0, 0, n_meas)
                    sig = meng.signal_bpm(data, fps, status=st)
                meng.pack_results(sig["bpm"], roi, st, sig["npeaks"], out=records[lo:hi])
            keep.append((roi, status, roi0, st, data, sig, crops))
        self._seq = seq0 + len(chunks)
        done = []
        for mstream in self._measure_streams:
            e = torch.cuda.Event()
            e.record(mstream)
            done.append(e)
        return dict(records=records, done_events=done, keep=keep, chunks=chunks, cap=(cap_w, cap_h),
                    mapped=dict(host=host, fps=fps, cal_first=cal_first, cal_len=cal_len, n_meas=n_meas, W=W))

    def _collect_mapped(self, ticket) -> np.ndarray:
        # read back on a stream of its own: the caller's stream may already hold the next batch's locate() calls, which
        # wait for uploads that have not happened yet
        rd = self._read_stream
        rec = ticket["records"]
        host_rec = torch.empty(rec.shape, dtype=torch.uint8).pin_memory() if rec.numel() else torch.empty(rec.shape, dtype=torch.uint8)
        with torch.cuda.stream(rd):
            for e in ticket["done_events"]:
                rd.wait_event(e)
            host_rec.copy_(rec, non_blocking=True)
        rd.synchronize()
        out = host_rec.numpy().view(RESULT_DTYPE).reshape(-1).copy()
        self.d2h_bytes += rec.numel()
        a = ticket["mapped"]
        cap_w, cap_h = ticket["cap"]
        found = out["status"] != 1                                            # every clip but RM_CLIP_NO_ROI has a box
        # bytes the crop kernel read over PCIe: the ROI's rows of the measure frames as aligned 32-bit words
        span = ((out["x"] & 3) + out["w"] + 3) // 4 * 4 if a["W"] % 4 == 0 else out["w"]
        fits = (out["w"] <= cap_w) & (out["h"] <= cap_h)
        self.h2d_bytes += int((span * out["h"] * a["n_meas"])[found & fits].sum())
        over = found & ~fits
        if over.any():
            # an ROI larger than the crop tensors: remember the size for the batches to come and redo the chunks concerned
            # through the staged path, which sizes its crops after reading the ROIs
            self.roi_cap = (max(self.roi_cap[0], (int(out["w"][found].max()) + 15) // 16 * 16),
                            max(self.roi_cap[1], (int(out["h"][found].max()) + 15) // 16 * 16))
            host = a["host"]
            mapped, self.mapped = self.mapped, False
            try:
                for lo, hi in ticket["chunks"]:
                    if over[lo:hi].any():
                        self.reruns += 1
                        out[lo:hi] = self.collect(self.submit(host[lo:hi], a["fps"], a["cal_first"], a["cal_len"]))
            finally:
                self.mapped = mapped
        ticket["keep"] = None
        return out

    def _chunk_schedule(self, n):
        """Uniform chunks of chunk_clips clips.  (A schedule that tapers towards the end -- to shorten the tail no upload
        hides -- measured no better on B200: every extra chunk costs a host round trip for its ROI.)"""
        return [(lo, min(n, lo + self.chunk_clips)) for lo in range(0, n, self.chunk_clips)]

    def _run_full_frames(self, host, fps, cal_first, cal_len):
        """Whole clips go up (2x the bytes of `run`); kept for comparison and for ROIs too large to be worth cropping."""
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
            dst = self._buffer(("full", slot), (hi - lo,) + tuple(host.shape[1:]))
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
        grouped into resolution classes (one kernel configuration, one set of level sizes and one TMA tensor map per
        class); the clips of a class are NOT copied together on the host -- each goes up from where it lies -- and all
        classes are submitted before the first is collected, so the uploads and the latency-bound stages of one class
        overlap the kernels of another.  Records come back in the order the clips were given."""
        classes = {}
        for i, c in enumerate(clips):
            classes.setdefault(tuple(c.shape), []).append(i)
        out = np.zeros(len(clips), RESULT_DTYPE)
        tickets = []
        for shape, idx in sorted(classes.items(), key=lambda kv: -int(np.prod(kv[0])) * len(kv[1])):   # heaviest first
            tickets.append((idx, self.submit([clips[i] for i in idx], fps, cal_first=cal_first, cal_len=cal_len)))
        for idx, ticket in tickets:
            out[idx] = self.collect(ticket)
        return out
