// Shared declarations for the respmon_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/respmon_b200.h"

#define RM_MAX_LEVELS 16

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
};

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
// OpenCV BORDER_REFLECT_101 for an index at most one reflection away.
__host__ __device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// The 5-tap [1 4 6 4 1] combination, one rounding per fused step: (a+e) + 4(b+d) + 6c.
__device__ __forceinline__ double tap5(double a, double b, double c, double d, double e) {
  return fma(6.0, c, fma(4.0, b + d, a + e));
}

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
