// Shared declarations for the respmon_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/respmon_b200.h"

#define RM_MAX_LEVELS 16
#define RM_MAX_CHUNKS 16   // frame chunks of the measure pipeline

struct rm_handle {
  rm_params p;
  int device;
  int sm_count;
  int smem_optin;       // max dynamic shared memory per block (opt-in)
  long long launches;
  char err[512];
  uint8_t lut[256];     // lossy u8 round trip (transforms.py:20-29)
  uint8_t* d_lut;       // device copy
  double* d_tvals;      // running time axis of the measure buffers (base.py:481-484), grown on demand
  int tvals_cap;
  // per-kernel device timing (the reference's tools.Benchmarker tags, tools.py:60-82, at kernel granularity)
  void* d_sig_scratch;  // measure() scratch (filtered windows, candidate peaks, fit queue), grown on demand
  size_t sig_scratch_bytes;
  // measure pipeline (rm_measure_signal): LK walks the frames in chunks on the caller's stream while the signal
  // stage of the finished chunks runs on aux_stream
  cudaStream_t aux_stream;                   // PCA + filtfilt/peaks of the chunks, in frame order
  cudaStream_t fit_stream[RM_MAX_CHUNKS];    // one per chunk: the Gaussian-fit gates of different chunks overlap
  cudaEvent_t ev_fork, ev_join, ev_chunk[RM_MAX_CHUNKS], ev_filt[RM_MAX_CHUNKS], ev_done[RM_MAX_CHUNKS];
  int measure_chunks;       // option "measure_chunks"
  int temporal_sparse;      // option "temporal_sparse": 1 (default) = band-pass through the kept bins where that is cheaper
  float* d_lk_pts;          // (cap_clips, 128, 2) points carried from one LK chunk to the next
  int* d_lk_idx;            // (cap_clips, 128) original corner index of every carried point
  int* d_lk_n;              // (cap_clips, 16) points each block of a clip still tracks
  int lk_state_cap;
  void* sig_job;            // host-side SignalJob of the measure pipeline (signal.cu)
  int no_minmax_seed;       // tests: pass 1 of the heat map without the seed kernel (no pruning at the start)
  int pyramid_mode;         // option "pyramid_mode": 1 = fused TMA kernel where the frames allow it (default), 0 = always the
                            // fallback (level 3 through HBM + pyramid_tail_kernel) -- tests compare the two bit for bit
  int pyramid_g4;           // option "pyramid_g4": level-4 image of the fused kernel: 0 = where more frame slots fit (default),
                            // 1 = shared memory, 2 = in the record
  int pyramid_cfg;          // option "pyramid_cfg": fused kernel's (ring stages, warps per CTA): 0 = by frame width (default),
                            // 1 = (4, 18), 2 = (3, 21), 3 = (2, 24)
  int force_global_lk;
  int force_generic_front;  // tests: float64 pyramid front even for uint8 frames the integer front supports  // tests: take the global-memory LK path even when the ROI fits shared memory
  int prof_on;
  int prof_n, prof_cap, prof_open;
  struct rm_prof_slot* prof_slots;
  int prof_entries;
  struct rm_prof_entry* prof_table;
};
struct rm_prof_slot {
  const char* name;
  cudaEvent_t a, b;
  cudaStream_t st;
};
struct rm_prof_entry {
  const char* name;
  double total_ms;
  long long launches;
};
#define RM_PROF_MAX_SLOTS 16384
#define RM_PROF_MAX_ENTRIES 64
// open a timing slot on `st` for the launch that follows; RM_LAUNCH_CHECK closes it
static inline void rm_prof_begin(rm_handle* h, cudaStream_t st, const char* name) {
  if (!h->prof_on || h->prof_n >= RM_PROF_MAX_SLOTS) return;
  if (!h->prof_slots) h->prof_slots = (rm_prof_slot*)calloc(RM_PROF_MAX_SLOTS, sizeof(rm_prof_slot));
  if (!h->prof_slots) return;
  rm_prof_slot* s = &h->prof_slots[h->prof_n];
  if (h->prof_n >= h->prof_cap) {
    if (cudaEventCreate(&s->a) != cudaSuccess || cudaEventCreate(&s->b) != cudaSuccess) return;
    h->prof_cap = h->prof_n + 1;
  }
  s->name = name;
  s->st = st;
  cudaEventRecord(s->a, st);
  h->prof_open = 1;
}
static inline void rm_prof_end(rm_handle* h) {
  if (!h->prof_open) return;
  rm_prof_slot* s = &h->prof_slots[h->prof_n];
  cudaEventRecord(s->b, s->st);
  h->prof_n++;
  h->prof_open = 0;
}
#define RM_PROF(h, st, name) rm_prof_begin((h), (cudaStream_t)(st), (name))

// ---- error plumbing --------------------------------------------------------------------------------------------
static inline int32_t rm_fail(rm_handle* h, int32_t code, const char* fmt, const char* a = "", long long b = 0,
                              long long c = 0) {
  if (h) snprintf(h->err, sizeof(h->err), fmt, a, b, c);
  return code;
}
#define RM_CHECK_ARG(h, cond, msg)                                              \
  do {                                                                          \
    if (!(cond)) return rm_fail((h), RM_ERR_INVALID, "%s: invalid argument: " msg, __func__); \
  } while (0)
#define RM_CUDA(h, call)                                                                               \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) {                                                                           \
      if (h) snprintf((h)->err, sizeof((h)->err), "%s: %s -> %s", __func__, #call, cudaGetErrorString(e_)); \
      return RM_ERR_CUDA;                                                                              \
    }                                                                                                  \
  } while (0)
#define RM_LAUNCH_CHECK(h)                                                                              \
  do {                                                                                                  \
    (h)->launches++;                                                                                    \
    rm_prof_end(h);                                                                                     \
    cudaError_t e_ = cudaGetLastError();                                                                \
    if (e_ != cudaSuccess) {                                                                            \
      snprintf((h)->err, sizeof((h)->err), "%s: kernel launch failed: %s", __func__, cudaGetErrorString(e_)); \
      return RM_ERR_CUDA;                                                                               \
    }                                                                                                   \
  } while (0)

struct DeviceGuard {
  int prev;
  bool ok;
  explicit DeviceGuard(int dev) : prev(-1), ok(true) {
    if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
    if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ---- geometry -------------------------------------------------------------------------------------------------------
struct LevelGeom {
  int n_levels;
  int w[RM_MAX_LEVELS], h[RM_MAX_LEVELS];
};
static inline LevelGeom make_geom(int W, int H, int n_levels) {
  LevelGeom g;
  g.n_levels = n_levels;
  g.w[0] = W;
  g.h[0] = H;
  for (int l = 1; l < n_levels; ++l) {
    g.w[l] = (g.w[l - 1] + 1) / 2;
    g.h[l] = (g.h[l - 1] + 1) / 2;
  }
  return g;
}
// offsets of levels skip..n_levels-2 inside the packed per-frame Laplacian record
struct RecordGeom {
  int first, last;           // levels first..last inclusive
  int w[RM_MAX_LEVELS], h[RM_MAX_LEVELS];
  int off[RM_MAX_LEVELS];    // offset (doubles) of level l inside the record, indexed by level
  int len;                   // total doubles per frame
};
static inline RecordGeom make_record(const LevelGeom& g, int skip) {
  RecordGeom r;
  r.first = skip;
  r.last = g.n_levels - 2;
  r.len = 0;
  for (int l = 0; l < RM_MAX_LEVELS; ++l) {
    r.w[l] = l < g.n_levels ? g.w[l] : 0;
    r.h[l] = l < g.n_levels ? g.h[l] : 0;
    r.off[l] = 0;
  }
  for (int l = r.first; l <= r.last; ++l) {
    r.off[l] = r.len;
    r.len += g.w[l] * g.h[l];
  }
  return r;
}

// ---- device helpers ------------------------------------------------------------------------------------------------
#include "pyr_core.h"   // reflect101, tap5, up_taps / up_combine: shared with the host build of the tests

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- internal entry points shared between translation units (not part of the C ABI) ---------------------------------
int32_t rmi_signal_setup(rm_handle* h, const double* data, int32_t n_clips, int32_t n_frames, double fps, double* bpm_out,
                         double* filtered_out, int32_t* peaks_out, int32_t* npeaks_out, const int32_t* status,
                         int n_chunks, cudaStream_t st, int win_f0, int win_frames);
int32_t rmi_signal_range(rm_handle* h, int f0, int f1, int chunk, cudaStream_t st, cudaStream_t st_fit,
                         cudaEvent_t ev_filtered);
