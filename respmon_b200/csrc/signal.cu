// measure() for every frame of every clip (base.py:340-352 with find_peaks base.py:312-338) and result packing.
// One thread per (clip, frame): the 128-sample rolling window is filtered (Butterworth filtfilt), peaks are picked
// (peakutils.indexes), each is gated by a Levenberg-Marquardt Gaussian fit (MINPACK lmdif as SciPy's curve_fit
// runs it) and the BPM is 60 / mean peak interval.  The arithmetic is in signal_core.h (shared with the host tests).
// Compiled with -fmad=false.
#include "common.cuh"
#include "signal_core.h"

struct SignalParams {
  const double* data;     // (n_clips, n_frames)
  const int32_t* status;  // (n_clips) or null
  int n_clips, n_frames;
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

__global__ void __launch_bounds__(64) signal_bpm_kernel(const SignalParams p) {
  const int clip = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= p.n_frames) return;
  double* bpm_out = p.bpm + (long long)clip * p.n_frames + f;
  const bool last = (f == p.n_frames - 1);
  const bool clip_ok = !p.status || p.status[clip] == RM_CLIP_OK || p.status[clip] == RM_CLIP_TRACK_LOST;
  int n = f + 1 < p.buf_len ? f + 1 : p.buf_len;
  const double* data = p.data + (long long)clip * p.n_frames + (f + 1 - n);
  const double* t = p.tvals + (f + 1 - n);
  bool run = clip_ok && (f + 1 > p.init_len);       // len(self.data) > measure_initialization_length (base.py:489)
  if (run)
    for (int i = 0; i < n; ++i) run &= (data[i] == data[i]);   // a NaN sample means tracking was lost
  int nacc = 0;
  double bpm = NAN;
  double filtered[SC_MAX_WIN];
  int peaks[SC_MAX_WIN];
  if (run) {
    ScScratch scratch;
    nacc = sc_measure_window(data, t, n, p.b, p.a, p.nc, p.width, p.thres, p.cutoff, filtered, peaks, &bpm, &scratch);
    if (nacc < 0) { nacc = 0; run = false; }
  }
  *bpm_out = bpm;
  if (last) {
    if (p.npeaks) p.npeaks[clip] = nacc;
    if (p.filtered)
      for (int i = 0; i < p.buf_len; ++i) p.filtered[(long long)clip * p.buf_len + i] = (run && i < n) ? filtered[i] : NAN;
    if (p.peaks)
      for (int i = 0; i < p.buf_len; ++i) p.peaks[(long long)clip * p.buf_len + i] = (i < nacc) ? peaks[i] : -1;
  }
}

__global__ void pack_results_kernel(const double* __restrict__ bpm, const int32_t* __restrict__ roi,
                                    const int32_t* __restrict__ status, const int32_t* __restrict__ npeaks, int n_clips,
                                    int n_frames, rm_result* __restrict__ out) {
  const int clip = blockIdx.x * blockDim.x + threadIdx.x;
  if (clip >= n_clips) return;
  rm_result r;
  r.bpm = NAN;
  for (int f = n_frames - 1; f >= 0; --f) {      // freq[-1]: the most recent frame that appended a BPM
    const double v = bpm[(long long)clip * n_frames + f];
    if (v == v) { r.bpm = v; break; }
  }
  r.x = roi[clip * 4]; r.y = roi[clip * 4 + 1]; r.w = roi[clip * 4 + 2]; r.h = roi[clip * 4 + 3];
  r.status = status[clip];
  if (r.status == RM_CLIP_OK && !(r.bpm == r.bpm)) r.status = RM_CLIP_NO_PEAKS;
  r.n_peaks = npeaks ? npeaks[clip] : 0;
  out[clip] = r;
}

extern "C" int32_t rm_signal_bpm(rm_handle* h, const double* data, int32_t n_clips, int32_t n_frames, double fps,
                                 double* bpm_out, double* filtered_out, int32_t* peaks_out, int32_t* npeaks_out,
                                 const int32_t* status, void* stream) {
  RM_CHECK_ARG(h, h && data && bpm_out && n_clips >= 0 && n_frames >= 1 && fps > 0, "null pointer or bad size");
  const int order = h->p.filter_order;
  if (order < 1 || order > SC_MAX_ORDER || h->p.measure_buffer_len > SC_MAX_WIN || h->p.measure_buffer_len < 2)
    return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: needs filter_order <= 7 and measure_buffer_len <= 128", __func__);
  if (n_clips == 0) return RM_OK;
  SignalParams p;
  memset(&p, 0, sizeof(p));
  // butter_lowpass(cutoff = freq_max*0.5, fs = fps): normal_cutoff = cutoff / (0.5*fs)  (transforms.py:59-60, base.py:342)
  const double wn = (h->p.freq_max * 0.5) / (0.5 * fps);
  if (!(wn > 0.0 && wn < 1.0)) return rm_fail(h, RM_ERR_INVALID, "%s: cutoff outside (0, Nyquist)", __func__);
  rm_butter_lowpass(order, wn, p.b, p.a);
  p.nc = order + 1;
  p.width = (int)floor(fps / h->p.freq_max);       // base.py:441
  if (2 * p.width > SC_MAX_FIT) return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: fps/freq_max > 32 not supported", __func__);
  p.data = data; p.status = status; p.n_clips = n_clips; p.n_frames = n_frames;
  p.dt = 1.0 / fps;
  p.buf_len = h->p.measure_buffer_len; p.init_len = h->p.measure_init_len;
  p.thres = h->p.peak_threshold; p.cutoff = h->p.gaussian_cutoff;
  p.bpm = bpm_out; p.filtered = filtered_out; p.peaks = peaks_out; p.npeaks = npeaks_out;
  DeviceGuard dg(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->tvals_cap < n_frames) {
    if (h->d_tvals) cudaFree(h->d_tvals);
    h->d_tvals = nullptr;
    h->tvals_cap = 0;
    RM_CUDA(h, cudaMalloc((void**)&h->d_tvals, (size_t)n_frames * 8));
    h->tvals_cap = n_frames;
  }
  p.tvals = h->d_tvals;
  RM_PROF(h, st, "tvals_kernel");
  tvals_kernel<<<1, 32, 0, st>>>(p.tvals, n_frames, p.dt);
  RM_LAUNCH_CHECK(h);
  dim3 grid(div_up(n_frames, 64), n_clips);
  RM_PROF(h, st, "signal_bpm_kernel");
  signal_bpm_kernel<<<grid, 64, 0, st>>>(p);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

extern "C" int32_t rm_pack_results(rm_handle* h, const double* bpm, const int32_t* roi, const int32_t* status,
                                   const int32_t* npeaks, int32_t n_clips, int32_t n_frames, rm_result* out, void* stream) {
  RM_CHECK_ARG(h, h && bpm && roi && status && out && n_clips >= 0 && n_frames >= 1, "null pointer or bad size");
  if (n_clips == 0) return RM_OK;
  DeviceGuard dg(h->device);
  RM_PROF(h, (cudaStream_t)stream, "pack_results_kernel");
  pack_results_kernel<<<div_up(n_clips, 128), 128, 0, (cudaStream_t)stream>>>(bpm, roi, status, npeaks, n_clips, n_frames, out);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}
