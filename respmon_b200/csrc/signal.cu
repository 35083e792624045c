// measure() for every frame of every clip (base.py:340-352 with find_peaks base.py:312-338) and result packing.
// signal_filter_peaks_kernel: one thread per (clip, frame) filters the 128-sample rolling window (Butterworth filtfilt)
// and picks the peaks (peakutils.indexes), queueing one Gaussian fit per peak.  signal_fit_kernel: a persistent grid
// pulls fits from the queue through an atomic cursor; SIG_FIT_G lanes share one Levenberg-Marquardt fit (MINPACK lmdif
// as SciPy's curve_fit runs it, lm_group.cuh).  signal_bpm_kernel: 60 / mean interval of the accepted peaks.
// The arithmetic is in signal_core.h (shared with the host tests).  Compiled with -fmad=false.
#include "common.cuh"
#include "signal_core.h"
#include "lm_group.cuh"

struct SignalParams {
  const double* data;     // (n_clips, n_frames)
  const int32_t* status;  // (n_clips) or null
  int n_clips, n_frames;
  int f0, f1;             // windows (frames) handled by this launch
  int win_f0, win_frames; // scratch rows are indexed by (clip, f - win_f0) with win_frames rows per clip
  int last_f;             // the frame whose window is exported through filtered / peaks / npeaks
  double dt;              // 1 / fps
  double b[SC_MAX_ORDER + 1], a[SC_MAX_ORDER + 1];
  int nc;                 // filter_order + 1
  int width;              // floor(fps / freq_max)
  int buf_len, init_len;
  double thres, cutoff;
  double* tvals;          // (n_frames) running sum of dt (base.py:481-484)
  double* bpm;            // (n_clips, n_frames)
  double* filtered;       // (n_clips, buf_len) last window (nullable)
  int32_t* peaks;         // (n_clips, buf_len) last window, -1 terminated (nullable)
  int32_t* npeaks;        // (n_clips) (nullable)
};

__global__ void tvals_kernel(double* tvals, int n, double dt) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < n; ++i) {
      if (i > 0) t = t + dt;   // self.t.append(self.t[-1] + 1/fps)
      tvals[i] = t;
    }
  }
}

// Stage A, one thread per (clip, frame) window: zero-phase Butterworth (filtfilt) and peakutils.indexes; every
// candidate peak that find_peaks would hand to gaussian_fit (base.py:318-327) becomes one item of the fit queue.
#define SIG_MAX_CAND 64   // peaks per window (min_dist >= 1 on <= 128 samples)

struct SignalScratch {
  double* filt;            // (n_windows, buf_len) filtered windows
  unsigned char* cand;     // (n_windows, SIG_MAX_CAND) candidate indices
  unsigned char* acc;      // (n_windows, SIG_MAX_CAND) 1 = accepted by the Gaussian gate
  int* ncand;              // (n_windows) candidates, or -1 when measure() does not run / raises for this window
  unsigned* queue;         // fit work items: window * SIG_MAX_CAND + k
  unsigned* queue_n;
  unsigned* cursor;        // next queue item to fit
};

__global__ void __launch_bounds__(64) signal_filter_peaks_kernel(const SignalParams p, const SignalScratch s) {
  const int clip = blockIdx.y;
  const int f = p.f0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= p.f1) return;
  const long long win = (long long)clip * p.win_frames + (f - p.win_f0);
  const bool clip_ok = !p.status || p.status[clip] == RM_CLIP_OK || p.status[clip] == RM_CLIP_TRACK_LOST;
  const int n = f + 1 < p.buf_len ? f + 1 : p.buf_len;
  const double* data = p.data + (long long)clip * p.n_frames + (f + 1 - n);
  bool run = clip_ok && (f + 1 > p.init_len);       // len(self.data) > measure_initialization_length (base.py:489)
  if (run)
    for (int i = 0; i < n; ++i) run &= (data[i] == data[i]);   // a NaN sample means tracking was lost
  int ncand = -1;
  if (run) {
    double filtered[SC_MAX_WIN];
    double ext[SC_MAX_WIN + 6 * (SC_MAX_ORDER + 1)], dy[SC_MAX_WIN];
    unsigned char mark[SC_MAX_WIN];
    int cand[SC_MAX_WIN];
    if (sc_filtfilt(p.b, p.a, p.nc, data, n, filtered, ext) == 0) {
      double* fo = s.filt + win * p.buf_len;
      for (int i = 0; i < n; ++i) fo[i] = filtered[i];
      ncand = sc_peak_indexes(filtered, n, p.thres, p.width, cand, dy, ext, mark);
      if (ncand > SIG_MAX_CAND) ncand = SIG_MAX_CAND;
      for (int k = 0; k < ncand; ++k) {
        const int idx = cand[k];
        s.cand[win * SIG_MAX_CAND + k] = (unsigned char)idx;
        s.acc[win * SIG_MAX_CAND + k] = 0;
        int w = p.width;                             // base.py:319-323
        if (idx - p.width < 0) w = idx;
        if (idx + w > n) w = n - idx;
        if (2 * w >= 3)                              // fewer points: gaussian_fit raises, the peak is dropped (base.py:336)
          s.queue[atomicAdd(s.queue_n, 1u)] = (unsigned)(win * SIG_MAX_CAND + k);
      }
    }
  }
  s.ncand[win] = ncand;
}

// Stage B, one group of G lanes per queued candidate: the Gaussian gate, peakutils.gaussian_fit = curve_fit =
// MINPACK lmdif (lm_group.cuh).  The fit is a long dependent float64 chain, so the stage is latency bound.  G = 32:
// a warp works on one fit, so its lanes never diverge (with several fits per warp every group's instruction stream
// is issued separately and the warp runs the SUM of its groups' iterations).  The grid is persistent: warps pull
// fits from the queue through an atomic cursor until it is empty, which balances the 10x spread in LM iterations.
#ifndef SIG_FIT_THREADS
#define SIG_FIT_THREADS 128
#endif
#ifndef SIG_FIT_G
#define SIG_FIT_G 4
#endif
#ifndef SIG_FIT_MINB
#define SIG_FIT_MINB 1
#endif
template <int G>
__global__ void __launch_bounds__(SIG_FIT_THREADS, SIG_FIT_MINB) signal_fit_kernel(const SignalParams p, const SignalScratch s,
                                                                                    int m_cap) {
  extern __shared__ __align__(16) double fit_smem[];
  const int lane = threadIdx.x & 31;
  LmGroup g;
  g.sub = lane & (G - 1);
  g.mask = (G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u)) << (lane & ~(G - 1));
  const int group_in_block = (int)(threadIdx.x / G);
  const unsigned total = *s.queue_n;
  const unsigned* queue = s.queue;
  unsigned* cursor = s.cursor;
  double* xs = fit_smem + (size_t)group_in_block * 7 * m_cap;
  double* ys = xs + m_cap;
  double* fvec = ys + m_cap;
  double* wa4 = fvec + m_cap;
  double* fjac = wa4 + m_cap;
  for (;;) {
    unsigned item = 0;
    if (g.sub == 0) item = atomicAdd(cursor, 1u);
    item = __shfl_sync(g.mask, item, lane & ~(G - 1));
    if (item >= total) return;                        // whole groups leave together
    const unsigned q = queue[item];
    const long long win = q / SIG_MAX_CAND;
    const int k = q % SIG_MAX_CAND;
    const int f = (int)(win % p.win_frames) + p.win_f0;
    const int n = f + 1 < p.buf_len ? f + 1 : p.buf_len;
    const int idx = s.cand[win * SIG_MAX_CAND + k];
    int w = p.width;                                   // base.py:319-323
    if (idx - p.width < 0) w = idx;
    if (idx + w > n) w = n - idx;
    int m = 2 * w;
    if (m > m_cap) m = m_cap;
    const double* t = p.tvals + (f + 1 - n) + (idx - w);
    const double* y = s.filt + win * p.buf_len + (idx - w);
    double mx = -INFINITY;
    __syncwarp(g.mask);                                // the previous fit's reads of xs/ys are done
    for (int i = g.sub; i < m; i += G) {
      xs[i] = t[i];
      ys[i] = y[i];
      mx = fmax(mx, y[i]);
    }
    mx = lmg_max<G>(g, mx);
    __syncwarp(g.mask);
    double par[SC_NP] = {mx, xs[0], (xs[1] - xs[0]) * 5.0};   // peakutils.gaussian_fit initial guess
    const int info = lmg_lmdif_gauss<G>(g, m, xs, ys, par, fvec, wa4, fjac);
    if (g.sub == 0) s.acc[win * SIG_MAX_CAND + k] = (info >= 1 && info <= 4 && par[2] < p.cutoff) ? 1 : 0;   // base.py:334, 336
  }
}

// Stage C, one thread per window: BPM = 60 / mean interval of the accepted peaks (base.py:347-352).
__global__ void signal_bpm_kernel(const SignalParams p, const SignalScratch s) {
  const int clip = blockIdx.y;
  const int f = p.f0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= p.f1) return;
  const long long win = (long long)clip * p.win_frames + (f - p.win_f0);
  const int n = f + 1 < p.buf_len ? f + 1 : p.buf_len;
  const double* t = p.tvals + (f + 1 - n);
  const int ncand = s.ncand[win];
  const bool last = (f == p.last_f);
  int nacc = 0, prev = -1;
  double sum = 0.0;
  for (int k = 0; k < ncand; ++k) {
    if (!s.acc[win * SIG_MAX_CAND + k]) continue;
    const int idx = s.cand[win * SIG_MAX_CAND + k];
    if (nacc > 0) sum += t[idx] - t[prev];
    prev = idx;
    if (last && p.peaks) p.peaks[(long long)clip * p.buf_len + nacc] = idx;
    ++nacc;
  }
  p.bpm[(long long)clip * p.n_frames + f] = nacc >= 2 ? 60.0 / (sum / (double)(nacc - 1)) : NAN;
  if (last) {
    if (p.npeaks) p.npeaks[clip] = nacc;
    if (p.filtered)
      for (int i = 0; i < p.buf_len; ++i)
        p.filtered[(long long)clip * p.buf_len + i] = (ncand >= 0 && i < n) ? s.filt[win * p.buf_len + i] : NAN;
    if (p.peaks)
      for (int i = nacc; i < p.buf_len; ++i) p.peaks[(long long)clip * p.buf_len + i] = -1;
  }
}

__global__ void pack_results_kernel(const double* __restrict__ bpm, const int32_t* __restrict__ roi,
                                    const int32_t* __restrict__ status, const int32_t* __restrict__ npeaks, int n_clips,
                                    int n_frames, int n_scan, rm_result* __restrict__ out) {
  const int clip = blockIdx.x * blockDim.x + threadIdx.x;
  if (clip >= n_clips) return;
  rm_result r;
  r.bpm = NAN;
  for (int f = n_scan - 1; f >= 0; --f) {        // freq[-1]: the most recent frame that appended a BPM
    const double v = bpm[(long long)clip * n_frames + f];
    if (v == v) { r.bpm = v; break; }
  }
  r.x = roi[clip * 4]; r.y = roi[clip * 4 + 1]; r.w = roi[clip * 4 + 2]; r.h = roi[clip * 4 + 3];
  r.status = status[clip];
  if (r.status == RM_CLIP_OK && !(r.bpm == r.bpm)) r.status = RM_CLIP_NO_PEAKS;
  r.n_peaks = npeaks ? npeaks[clip] : 0;
  out[clip] = r;
}

struct SignalJob {
  SignalParams p;
  SignalScratch sc;        // queue / queue_n / cursor are those of chunk 0; chunk c uses its own slice
  int m_cap, n_chunks, grid_cap;
  size_t fit_smem, max_items_per_frame;
};

// Validation, filter design, time axis and scratch for measure() over (n_clips, n_frames) windows that will be
// processed in up to n_chunks frame ranges (possibly concurrently with the producer of `data`).  Allocates; launches
// only tvals_kernel (on st).
int32_t rmi_signal_setup(rm_handle* h, const double* data, int32_t n_clips, int32_t n_frames, double fps, double* bpm_out,
                         double* filtered_out, int32_t* peaks_out, int32_t* npeaks_out, const int32_t* status,
                         int n_chunks, cudaStream_t st, int win_f0, int win_frames) {
  if (win_frames <= 0) { win_f0 = 0; win_frames = n_frames; }   // whole clips: one scratch row per frame
  RM_CHECK_ARG(h, h && data && bpm_out && n_clips >= 0 && n_frames >= 1 && fps > 0, "null pointer or bad size");
  const int order = h->p.filter_order;
  if (order < 1 || order > SC_MAX_ORDER || h->p.measure_buffer_len > SC_MAX_WIN || h->p.measure_buffer_len < 2)
    return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: needs filter_order <= 7 and measure_buffer_len <= 128", __func__);
  if (!h->sig_job) h->sig_job = calloc(1, sizeof(SignalJob));
  if (!h->sig_job) return rm_fail(h, RM_ERR_INVALID, "%s: out of host memory", __func__);
  SignalJob* job = reinterpret_cast<SignalJob*>(h->sig_job);
  SignalParams& p = job->p;
  memset(&p, 0, sizeof(p));
  // butter_lowpass(cutoff = freq_max*0.5, fs = fps): normal_cutoff = cutoff / (0.5*fs)  (transforms.py:59-60, base.py:342)
  const double wn = (h->p.freq_max * 0.5) / (0.5 * fps);
  if (!(wn > 0.0 && wn < 1.0)) return rm_fail(h, RM_ERR_INVALID, "%s: cutoff outside (0, Nyquist)", __func__);
  rm_butter_lowpass(order, wn, p.b, p.a);
  p.nc = order + 1;
  p.width = (int)floor(fps / h->p.freq_max);       // base.py:441
  if (2 * p.width > SC_MAX_FIT) return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: fps/freq_max > 32 not supported", __func__);
  p.data = data; p.status = status; p.n_clips = n_clips; p.n_frames = n_frames;
  p.f0 = 0; p.f1 = n_frames;
  p.win_f0 = win_f0; p.win_frames = win_frames;
  p.last_f = win_f0 + win_frames - 1;
  p.dt = 1.0 / fps;
  p.buf_len = h->p.measure_buffer_len; p.init_len = h->p.measure_init_len;
  p.thres = h->p.peak_threshold; p.cutoff = h->p.gaussian_cutoff;
  p.bpm = bpm_out; p.filtered = filtered_out; p.peaks = peaks_out; p.npeaks = npeaks_out;
  job->n_chunks = n_chunks < 1 ? 1 : (n_chunks > RM_MAX_CHUNKS ? RM_MAX_CHUNKS : n_chunks);
  if (n_clips == 0) return RM_OK;
  DeviceGuard dg(h->device);
  if (h->tvals_cap < n_frames) {
    if (h->d_tvals) cudaFree(h->d_tvals);
    h->d_tvals = nullptr;
    h->tvals_cap = 0;
    RM_CUDA(h, cudaMalloc((void**)&h->d_tvals, (size_t)n_frames * 8));
    h->tvals_cap = n_frames;
  }
  p.tvals = h->d_tvals;
  // scratch owned by the handle, grown on demand
  const size_t n_win = (size_t)n_clips * win_frames;
  const size_t need = n_win * p.buf_len * 8 + n_win * SIG_MAX_CAND * 2 + n_win * 4 + n_win * SIG_MAX_CAND * 4 + 1024;
  if (h->sig_scratch_bytes < need) {
    if (h->d_sig_scratch) cudaFree(h->d_sig_scratch);
    h->d_sig_scratch = nullptr;
    h->sig_scratch_bytes = 0;
    RM_CUDA(h, cudaMalloc(&h->d_sig_scratch, need));
    h->sig_scratch_bytes = need;
  }
  SignalScratch& sc = job->sc;
  unsigned char* base = reinterpret_cast<unsigned char*>(h->d_sig_scratch);
  sc.filt = reinterpret_cast<double*>(base);                base += n_win * p.buf_len * 8;
  sc.queue = reinterpret_cast<unsigned*>(base);             base += n_win * SIG_MAX_CAND * 4;
  sc.ncand = reinterpret_cast<int*>(base);                  base += n_win * 4;
  sc.queue_n = reinterpret_cast<unsigned*>(base);           base += 256;   // 4 counters per chunk (RM_MAX_CHUNKS <= 16)
  sc.cand = base;                                           base += n_win * SIG_MAX_CAND;
  sc.acc = base;
  sc.cursor = sc.queue_n + 1;
  // every window holds at most (buf_len / width + 1) candidates that survive min_dist = width
  job->m_cap = 2 * p.width < SC_MAX_FIT ? 2 * p.width : SC_MAX_FIT;
  if (job->m_cap < 4) job->m_cap = 4;
  const int groups = SIG_FIT_THREADS / SIG_FIT_G;
  job->fit_smem = (size_t)groups * 7 * job->m_cap * sizeof(double);
  const size_t per_win = (size_t)(p.buf_len / (p.width > 0 ? p.width : 1)) + 2;
  job->max_items_per_frame = (size_t)n_clips * (per_win < SIG_MAX_CAND ? per_win : SIG_MAX_CAND);
  int per_sm = 0;
  RM_CUDA(h, cudaFuncSetAttribute(signal_fit_kernel<SIG_FIT_G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)job->fit_smem));
  RM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, signal_fit_kernel<SIG_FIT_G>, SIG_FIT_THREADS,
                                                           job->fit_smem));
  job->grid_cap = h->sm_count * (per_sm > 0 ? per_sm : 1);   // persistent grid: what can be resident
  RM_CUDA(h, cudaMemsetAsync(sc.queue_n, 0, 256, st));
  RM_PROF(h, st, "tvals_kernel");
  tvals_kernel<<<1, 32, 0, st>>>(p.tvals, n_frames, p.dt);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

// measure() for the windows ending at frames [f0, f1) (chunk index `chunk` selects the fit queue) on stream st.
int32_t rmi_signal_range(rm_handle* h, int f0, int f1, int chunk, cudaStream_t st, cudaStream_t st_fit,
                         cudaEvent_t ev_filtered) {
  SignalJob* job = reinterpret_cast<SignalJob*>(h->sig_job);
  if (!job || chunk < 0 || chunk >= RM_MAX_CHUNKS) return rm_fail(h, RM_ERR_INVALID, "%s: no signal job", __func__);
  if (job->p.n_clips == 0 || f1 <= f0) return RM_OK;
  SignalParams p = job->p;
  SignalScratch sc = job->sc;
  p.f0 = f0; p.f1 = f1;
  sc.queue = job->sc.queue + (size_t)p.n_clips * (f0 - p.win_f0) * SIG_MAX_CAND;   // a slice no other chunk's windows reach
  sc.queue_n = job->sc.queue_n + 4 * chunk;
  sc.cursor = sc.queue_n + 1;
  dim3 grid(div_up(f1 - f0, 64), p.n_clips);
  RM_PROF(h, st, "signal_filter_peaks_kernel");
  signal_filter_peaks_kernel<<<grid, 64, 0, st>>>(p, sc);
  RM_LAUNCH_CHECK(h);
  if (st_fit != st) {   // the fits (and the BPM that folds them) continue on their own stream
    RM_CUDA(h, cudaEventRecord(ev_filtered, st));
    RM_CUDA(h, cudaStreamWaitEvent(st_fit, ev_filtered, 0));
  }
  const int groups = SIG_FIT_THREADS / SIG_FIT_G;
  long long grid_fit = div_up((long long)(job->max_items_per_frame * (size_t)(f1 - f0)), groups);
  if (grid_fit > job->grid_cap) grid_fit = job->grid_cap;
  RM_PROF(h, st_fit, "signal_fit_kernel");
  signal_fit_kernel<SIG_FIT_G><<<(int)grid_fit, SIG_FIT_THREADS, job->fit_smem, st_fit>>>(p, sc, job->m_cap);
  RM_LAUNCH_CHECK(h);
  RM_PROF(h, st_fit, "signal_bpm_kernel");
  signal_bpm_kernel<<<grid, 64, 0, st_fit>>>(p, sc);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

extern "C" int32_t rm_signal_bpm(rm_handle* h, const double* data, int32_t n_clips, int32_t n_frames, double fps,
                                 double* bpm_out, double* filtered_out, int32_t* peaks_out, int32_t* npeaks_out,
                                 const int32_t* status, void* stream) {
  if (!h) return RM_ERR_INVALID;
  int32_t rc = rmi_signal_setup(h, data, n_clips, n_frames, fps, bpm_out, filtered_out, peaks_out, npeaks_out, status, 1,
                        (cudaStream_t)stream, 0, 0);
  if (rc != RM_OK || n_clips == 0) return rc;
  DeviceGuard dg(h->device);
  return rmi_signal_range(h, 0, n_frames, 0, (cudaStream_t)stream, (cudaStream_t)stream, nullptr);
}

extern "C" int32_t rm_pack_results(rm_handle* h, const double* bpm, const int32_t* roi, const int32_t* status,
                                   const int32_t* npeaks, int32_t n_clips, int32_t n_frames, rm_result* out, void* stream) {
  RM_CHECK_ARG(h, h && bpm && roi && status && out && n_clips >= 0 && n_frames >= 1, "null pointer or bad size");
  if (n_clips == 0) return RM_OK;
  DeviceGuard dg(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  RM_PROF(h, st, "pack_results_kernel");
  pack_results_kernel<<<div_up(n_clips, 128), 128, 0, st>>>(bpm, roi, status, npeaks, n_clips, n_frames, n_frames, out);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

// Live streams: the same records from (n_clips, cap) histories of which the first n_valid frames have been measured.
extern "C" int32_t rm_pack_results_stream(rm_handle* h, const double* bpm, const int32_t* roi, const int32_t* status,
                                          const int32_t* npeaks, int32_t n_clips, int32_t cap, int32_t n_valid,
                                          rm_result* out, void* stream) {
  RM_CHECK_ARG(h, h && bpm && roi && status && out && n_clips >= 0 && cap >= 1 && n_valid >= 0 && n_valid <= cap,
               "null pointer or bad size");
  if (n_clips == 0) return RM_OK;
  DeviceGuard dg(h->device);
  RM_PROF(h, (cudaStream_t)stream, "pack_results_kernel");
  pack_results_kernel<<<div_up(n_clips, 128), 128, 0, (cudaStream_t)stream>>>(bpm, roi, status, npeaks, n_clips, cap,
                                                                              n_valid, out);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}
