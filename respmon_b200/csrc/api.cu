// Handle lifetime, host-side helpers and the synthetic clip generator.
#include <math.h>
#include <stdlib.h>

#include <complex>

#include "common.cuh"

extern "C" int32_t rm_version(void) { return RM_VERSION; }

extern "C" int32_t rm_default_params(rm_params* p) {
  if (!p) return RM_ERR_INVALID;
  p->pyramid_levels = 9;
  p->skip_levels_at_top = 4;
  p->freq_min = 0.1;
  p->freq_max = 1.0;
  p->amplification = 500.0;
  p->temporal_threshold = 0.7;
  p->threshold = 20;
  p->max_corners = 100;
  p->quality_level = 0.3;
  p->min_distance = 7;
  p->block_size = 7;
  p->lk_win = 15;
  p->lk_max_level = 2;
  p->lk_max_iter = 10;
  p->lk_eps = 0.03;
  p->lk_min_eig = 1e-4;
  p->gaussian_cutoff = 10.0;
  p->filter_order = 3;
  p->measure_buffer_len = 128;
  p->measure_init_len = 12;
  p->peak_threshold = 0.3;
  return RM_OK;
}

extern "C" int32_t rm_lossy_u8_lut(uint8_t* lut) {
  if (!lut) return RM_ERR_INVALID;
  // uint8 -> gray * (1/255) (transforms.py:22) -> * 255 truncated into uint8 (transforms.py:28)
  volatile double inv = 1.0 / 255;
  for (int k = 0; k < 256; ++k) {
    volatile double f = (double)k * inv;
    volatile double g = f * 255;
    lut[k] = (uint8_t)(int)g;
  }
  return RM_OK;
}

// scipy.signal.butter(order, wn, 'low', analog=False): Butterworth prototype poles, lp2lp with the pre-warped
// frequency, bilinear transform at fs = 2, zpk -> tf (transforms.py:58-63).
extern "C" int32_t rm_butter_lowpass(int32_t order, double wn, double* b, double* a) {
  if (order < 1 || order > 7 || !(wn > 0.0 && wn < 1.0) || !b || !a) return RM_ERR_INVALID;
  typedef std::complex<double> cd;
  const double fs = 2.0;
  const double warped = 2.0 * fs * tan(M_PI * wn / fs);
  cd pz[8];
  cd prod_den(1.0, 0.0);
  for (int i = 0; i < order; ++i) {
    int m = -order + 1 + 2 * i;
    cd pole = -std::exp(cd(0.0, M_PI * m / (2.0 * order))) * warped;
    pz[i] = (2.0 * fs + pole) / (2.0 * fs - pole);
    prod_den *= (2.0 * fs - pole);
  }
  const double gain = pow(warped, order) * (cd(1.0, 0.0) / prod_den).real();
  cd pa[9];   // poly(pz)
  pa[0] = cd(1.0, 0.0);
  for (int i = 1; i <= order; ++i) pa[i] = cd(0.0, 0.0);
  for (int i = 0; i < order; ++i)
    for (int k = i + 1; k >= 1; --k) pa[k] -= pz[i] * pa[k - 1];
  double pb[9];   // poly of `order` zeros at -1: binomial coefficients
  pb[0] = 1.0;
  for (int i = 1; i <= order; ++i) pb[i] = 0.0;
  for (int i = 0; i < order; ++i)
    for (int k = i + 1; k >= 1; --k) pb[k] += pb[k - 1];
  for (int i = 0; i <= order; ++i) {
    a[i] = pa[i].real();
    b[i] = gain * pb[i];
  }
  return RM_OK;
}

extern "C" int32_t rm_level_sizes(int32_t W, int32_t H, int32_t n_levels, int32_t* wh_out) {
  if (W < 1 || H < 1 || n_levels < 1 || n_levels > RM_MAX_LEVELS || !wh_out) return RM_ERR_INVALID;
  LevelGeom g = make_geom(W, H, n_levels);
  for (int l = 0; l < n_levels; ++l) {
    wh_out[2 * l] = g.w[l];
    wh_out[2 * l + 1] = g.h[l];
  }
  return RM_OK;
}

extern "C" int32_t rm_create(const rm_params* params, int32_t device, rm_handle** out) {
  if (!out) return RM_ERR_INVALID;
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) return RM_ERR_CUDA;
  rm_handle* h = (rm_handle*)calloc(1, sizeof(rm_handle));
  if (!h) return RM_ERR_INVALID;
  if (params) h->p = *params;
  else rm_default_params(&h->p);
  h->device = device;
  DeviceGuard dg(device);
  cudaDeviceProp prop;
  if (!dg.ok || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    free(h);
    return RM_ERR_CUDA;
  }
  if (prop.major < 10) {   // sm_100a cubins only: refuse anything else loudly
    free(h);
    return RM_ERR_UNSUPPORTED;
  }
  h->sm_count = prop.multiProcessorCount;
  h->smem_optin = (int)prop.sharedMemPerBlockOptin;
  rm_lossy_u8_lut(h->lut);
  if (cudaMalloc((void**)&h->d_lut, 256) != cudaSuccess ||
      cudaMemcpy(h->d_lut, h->lut, 256, cudaMemcpyHostToDevice) != cudaSuccess) {
    free(h);
    return RM_ERR_CUDA;
  }
  h->measure_chunks = 4;
  h->pyramid_mode = 1;
  h->temporal_sparse = 1;
  bool ok = cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < RM_MAX_CHUNKS; ++i) {
    ok = cudaEventCreateWithFlags(&h->ev_chunk[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_filt[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&h->fit_stream[i], cudaStreamNonBlocking) == cudaSuccess;
  }
  if (!ok) {
    cudaFree(h->d_lut);
    free(h);
    return RM_ERR_CUDA;
  }
  h->err[0] = 0;
  *out = h;
  return RM_OK;
}

extern "C" int32_t rm_destroy(rm_handle* h) {
  if (!h) return RM_OK;
  {
    DeviceGuard dg(h->device);
    if (h->d_lut) cudaFree(h->d_lut);
    if (h->d_tvals) cudaFree(h->d_tvals);
    if (h->d_sig_scratch) cudaFree(h->d_sig_scratch);
    free(h->sig_job);
    if (h->d_lk_pts) cudaFree(h->d_lk_pts);
    if (h->d_lk_idx) cudaFree(h->d_lk_idx);
    if (h->d_lk_n) cudaFree(h->d_lk_n);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    for (int i = 0; i < RM_MAX_CHUNKS; ++i) {
      if (h->ev_chunk[i]) cudaEventDestroy(h->ev_chunk[i]);
      if (h->ev_filt[i]) cudaEventDestroy(h->ev_filt[i]);
      if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
      if (h->fit_stream[i]) cudaStreamDestroy(h->fit_stream[i]);
    }
    for (int i = 0; i < h->prof_cap; ++i) {
      cudaEventDestroy(h->prof_slots[i].a);
      cudaEventDestroy(h->prof_slots[i].b);
    }
  }
  free(h->prof_slots);
  free(h->prof_table);
  free(h);
  return RM_OK;
}

extern "C" int32_t rm_set_option(rm_handle* h, const char* name, int64_t value) {
  if (!h || !name) return RM_ERR_INVALID;
  if (strcmp(name, "force_global_lk") == 0) { h->force_global_lk = value != 0; return RM_OK; }
  if (strcmp(name, "force_generic_front") == 0) { h->force_generic_front = value != 0; return RM_OK; }
  if (strcmp(name, "pyramid_mode") == 0) { h->pyramid_mode = value != 0; return RM_OK; }
  if (strcmp(name, "pyramid_g4") == 0) { h->pyramid_g4 = value < 0 || value > 2 ? 0 : (int)value; return RM_OK; }
  if (strcmp(name, "pyramid_cfg") == 0) { h->pyramid_cfg = value < 0 || value > 3 ? 0 : (int)value; return RM_OK; }
  if (strcmp(name, "no_minmax_seed") == 0) { h->no_minmax_seed = value != 0; return RM_OK; }
  if (strcmp(name, "temporal_sparse") == 0) { h->temporal_sparse = value != 0; return RM_OK; }
  if (strcmp(name, "measure_chunks") == 0) {
    if (value < 1 || value > RM_MAX_CHUNKS) return rm_fail(h, RM_ERR_INVALID, "%s: measure_chunks must be 1..16", __func__);
    h->measure_chunks = (int)value;
    return RM_OK;
  }
  return rm_fail(h, RM_ERR_INVALID, "%s: unknown option", __func__);
}

// ---------------------------------------------------------------------------------------------------- profiling
extern "C" int32_t rm_profile_enable(rm_handle* h, int32_t on) {
  if (!h) return RM_ERR_INVALID;
  h->prof_on = on ? 1 : 0;
  return RM_OK;
}
extern "C" int32_t rm_profile_reset(rm_handle* h) {
  if (!h) return RM_ERR_INVALID;
  h->prof_n = 0;
  h->prof_open = 0;
  h->prof_entries = 0;
  return RM_OK;
}
// Waits for the recorded launches, folds their durations into the per-kernel table; returns the table size (or < 0).
// Timeline view of the recorded launches (before rm_profile_collect folds them): start / end of launch i in ms relative
// to the start of launch 0.  Returns the number of recorded launches when i < 0.
extern "C" int32_t rm_profile_slot(rm_handle* h, int32_t i, const char** name, double* start_ms, double* end_ms) {
  if (!h) return RM_ERR_INVALID;
  if (i < 0) return h->prof_n;
  if (i >= h->prof_n || !name || !start_ms || !end_ms) return RM_ERR_INVALID;
  DeviceGuard dg(h->device);
  rm_prof_slot* s0 = &h->prof_slots[0];
  rm_prof_slot* s = &h->prof_slots[i];
  float a = 0.f, b = 0.f;
  RM_CUDA(h, cudaEventSynchronize(s->b));
  RM_CUDA(h, cudaEventElapsedTime(&a, s0->a, s->a));
  RM_CUDA(h, cudaEventElapsedTime(&b, s0->a, s->b));
  *name = s->name;
  *start_ms = a;
  *end_ms = b;
  return RM_OK;
}

extern "C" int32_t rm_profile_collect(rm_handle* h) {
  if (!h) return RM_ERR_INVALID;
  DeviceGuard dg(h->device);
  if (!h->prof_table) h->prof_table = (rm_prof_entry*)calloc(RM_PROF_MAX_ENTRIES, sizeof(rm_prof_entry));
  if (!h->prof_table) return RM_ERR_INVALID;
  for (int i = 0; i < h->prof_n; ++i) {
    rm_prof_slot* s = &h->prof_slots[i];
    float ms = 0.f;
    RM_CUDA(h, cudaEventSynchronize(s->b));
    RM_CUDA(h, cudaEventElapsedTime(&ms, s->a, s->b));
    int e = 0;
    for (; e < h->prof_entries; ++e)
      if (h->prof_table[e].name == s->name || strcmp(h->prof_table[e].name, s->name) == 0) break;
    if (e == h->prof_entries) {
      if (e >= RM_PROF_MAX_ENTRIES) continue;
      h->prof_table[e].name = s->name;
      h->prof_table[e].total_ms = 0.0;
      h->prof_table[e].launches = 0;
      h->prof_entries++;
    }
    h->prof_table[e].total_ms += ms;
    h->prof_table[e].launches++;
  }
  h->prof_n = 0;
  return h->prof_entries;
}
extern "C" int32_t rm_profile_entry(rm_handle* h, int32_t i, const char** name, double* total_ms, int64_t* launches) {
  if (!h || i < 0 || i >= h->prof_entries || !name || !total_ms || !launches) return RM_ERR_INVALID;
  *name = h->prof_table[i].name;
  *total_ms = h->prof_table[i].total_ms;
  *launches = h->prof_table[i].launches;
  return RM_OK;
}

extern "C" const char* rm_last_error(rm_handle* h) { return h ? h->err : "null handle"; }
extern "C" int64_t rm_launch_count(rm_handle* h) { return h ? h->launches : 0; }

// ---------------------------------------------------------------------------------------------------- synthetic clips
// SURVEY.md App. D / respmon_b200/synth.py -- integer only, bit-identical to the numpy generator.
__device__ __forceinline__ unsigned h32(int ix, int iy, unsigned s) {
  unsigned v = (unsigned)ix * 0x9E3779B1u + (unsigned)iy * 0x85EBCA77u + s * 0xC2B2AE3Du;
  v ^= v >> 16;
  v *= 0x7FEB352Du;
  v ^= v >> 15;
  v *= 0x846CA68Bu;
  v ^= v >> 16;
  return v >> 24;
}
__device__ __forceinline__ int vnoise_q8(long long xq, long long yq, int e, unsigned s) {
  const int sh = 8 + e;
  int ix = (int)(xq >> sh), iy = (int)(yq >> sh);
  int fx = (int)((xq & ((1ll << sh) - 1)) >> e), fy = (int)((yq & ((1ll << sh) - 1)) >> e);
  int v00 = h32(ix, iy, s), v10 = h32(ix + 1, iy, s), v01 = h32(ix, iy + 1, s), v11 = h32(ix + 1, iy + 1, s);
  int top = v00 * (256 - fx) + v10 * fx;
  int bot = v01 * (256 - fx) + v11 * fx;
  return (top * (256 - fy) + bot * fy) >> 8;
}

__global__ void synth_clips_kernel(const rm_clip_spec* __restrict__ specs, const int32_t* __restrict__ dq8,
                                   uint8_t* __restrict__ out) {
  const rm_clip_spec sp = specs[blockIdx.z];
  const int T = sp.n_frames, W = sp.width, H = sp.height;
  const long long hw = (long long)W * H;
  uint8_t* clip = out + (long long)blockIdx.z * T * hw;
  const int32_t* d = dq8 + (long long)blockIdx.z * T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)(i / W);
    const bool in_patch = x >= sp.x0 && x < sp.x0 + sp.w0 && y >= sp.y0 && y < sp.y0 + sp.h0;
    if (!in_patch) {
      const long long xq = (long long)x << 8, yq = (long long)y << 8;
      uint8_t bg = (uint8_t)((vnoise_q8(xq, yq, 4, (unsigned)sp.seed) + vnoise_q8(xq, yq, 3, (unsigned)sp.seed + 1)) >> 9);
      for (int t = 0; t < T; ++t) clip[(long long)t * hw + i] = bg;
    } else {
      const long long xq = (long long)x << 8;
      for (int t = 0; t < T; ++t) {
        const long long yq = ((long long)y << 8) + d[t] + (64 << 8);
        clip[(long long)t * hw + i] =
            (uint8_t)((vnoise_q8(xq, yq, 3, (unsigned)sp.seed + 7) + vnoise_q8(xq, yq, 2, (unsigned)sp.seed + 8)) >> 9);
      }
    }
  }
}

extern "C" int32_t rm_synth_clips(rm_handle* h, const rm_clip_spec* specs, const int32_t* dq8, int32_t n_clips,
                                  uint8_t* out, void* stream) {
  RM_CHECK_ARG(h, h && specs && dq8 && out && n_clips >= 0, "null pointer or bad size");
  if (n_clips == 0) return RM_OK;
  DeviceGuard dg(h->device);
  dim3 grid(h->sm_count * 2 / (n_clips < 8 ? n_clips : 8) + 1, 1, n_clips);
  RM_PROF(h, (cudaStream_t)stream, "synth_clips_kernel");
  synth_clips_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(specs, dq8, out);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

// ---------------------------------------------------------------------------------------------------- frame ingest
// cv2.cvtColor(frame, COLOR_BGR2GRAY) on 8-bit frames (next_frame, base.py:230): OpenCV's 15-bit fixed point,
// gray = (3735 B + 19235 G + 9798 R + 2^14) >> 15 (exhaustively identical to cv2 over all 2^24 colours).
__global__ void bgr_to_gray_kernel(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint8_t* px = bgr + 3 * i;
    gray[i] = (uint8_t)((3735u * px[0] + 19235u * px[1] + 9798u * px[2] + (1u << 14)) >> 15);
  }
}

extern "C" int32_t rm_bgr_to_gray(rm_handle* h, const uint8_t* bgr, uint8_t* gray_out, int64_t n_pixels, void* stream) {
  RM_CHECK_ARG(h, h && bgr && gray_out && n_pixels >= 0, "null pointer or bad size");
  if (n_pixels == 0) return RM_OK;
  DeviceGuard dg(h->device);
  const long long want = (n_pixels + 255) / 256;
  const int blocks = (int)(want < (long long)h->sm_count * 8 ? want : (long long)h->sm_count * 8);
  RM_PROF(h, (cudaStream_t)stream, "bgr_to_gray_kernel");
  bgr_to_gray_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(bgr, gray_out, n_pixels);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}
