// Calibration heat map: collapse of the band-passed pyramid, intensity clip, time average, normalise to uint8.
//
// Reference: collapse_laplacian_video_pyramid (pyramid.py:51-69) called from transforms.py:182 on a pyramid whose
// levels 0..skip-1 and top are all zero, then transforms.py:184-192 (global min/max over (T,H,W), top = max -
// (max-min)*threshold, values >= top replaced by min) and base.py:562-564 (mean over T, min-max normalise, *255
// truncated to uint8).
//
// Nothing of size (T,H,W) is ever stored: the collapsed full-resolution value of every pixel-frame is evaluated twice
// (pass 1: min/max, pass 2: clipped mean).  Per frame the coarse part (levels top-1..skip, 1600 values at VGA) is
// collapsed once by collapse_head_kernel into A_skip (40x30); the two passes then upsample A_skip by `skip` pyrUp
// steps on the fly: levels skip-1..2 as small shared-memory patches per 64x64 output tile, the last two steps
// (16x the pixels) entirely in registers, 4x4 outputs per thread.  All 1/64 factors are exact powers of two and are
// folded into one final scale.  These passes are FP64-ALU bound (about 11 flop per pixel-frame), not HBM bound.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------- pyrUp taps
struct Tap3 {
  int i0, i1, i2;   // source indices (border rules applied)
  int odd;          // 1: 4*(s[i0]+s[i1]) ; 0: s[i0] + 6 s[i1] + s[i2]
};
__host__ __device__ __forceinline__ Tap3 tap3(int o, int n) {
  Tap3 t;
  int i = o >> 1;
  int nx = (i + 1 < n) ? i + 1 : n - 1;
  t.odd = o & 1;
  if (t.odd) { t.i0 = i; t.i1 = nx; t.i2 = nx; }
  else { t.i0 = reflect101(i - 1, n); t.i1 = i; t.i2 = nx; }
  return t;
}
__device__ __forceinline__ double up3(int odd, double a, double b, double c) {
  return odd ? 4.0 * (a + b) : fma(6.0, b, a + c);
}

// ---------------------------------------------------------------------------------------------------- collapse head
struct HeadParams {
  const double* bp;     // (n_frames, record_len)
  double* a_out;        // (n_frames, h[first]*w[first])
  long long n_frames;
  int first, last;      // record levels
  int w[RM_MAX_LEVELS], h[RM_MAX_LEVELS], off[RM_MAX_LEVELS];
  int record_len;
};

__global__ void __launch_bounds__(256) collapse_head_kernel(const HeadParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* a = reinterpret_cast<double*>(smem_raw);
  for (long long f = blockIdx.x; f < p.n_frames; f += gridDim.x) {
    const double* src = p.bp + f * p.record_len;
    for (int i = threadIdx.x; i < p.record_len; i += blockDim.x) a[i] = src[i];
    __syncthreads();
    // img = pyrUp(img) + level, from the coarsest band-passed level down (pyramid.py:53-55); the top level is zeros
    for (int l = p.last - 1; l >= p.first; --l) {
      const double* s = a + p.off[l + 1];
      double* d = a + p.off[l];
      const int sw = p.w[l + 1], sh = p.h[l + 1], dw = p.w[l], dh = p.h[l];
      for (int i = threadIdx.x; i < dw * dh; i += blockDim.x) {
        int x = i % dw, y = i / dw;
        Tap3 tx = tap3(x, sw), ty = tap3(y, sh);
        const double* r0 = s + ty.i0 * sw;
        const double* r1 = s + ty.i1 * sw;
        const double* r2 = s + ty.i2 * sw;
        double h0 = up3(tx.odd, r0[tx.i0], r0[tx.i1], r0[tx.i2]);
        double h1 = up3(tx.odd, r1[tx.i0], r1[tx.i1], r1[tx.i2]);
        double h2 = up3(tx.odd, r2[tx.i0], r2[tx.i1], r2[tx.i2]);
        d[i] = up3(ty.odd, h0, h1, h2) * (1.0 / 64.0) + d[i];
      }
      __syncthreads();
    }
    const int n0 = p.w[p.first] * p.h[p.first];
    double* dst = p.a_out + f * n0;
    for (int i = threadIdx.x; i < n0; i += blockDim.x) dst[i] = a[p.off[p.first] + i];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------- tile passes
#define HM_TILE 64          // output tile (level 0) is 64x64, 16x16 threads, 4x4 px per thread
#define HM_MAX_STAGES 4     // shared-memory levels skip-1 .. 2
#define HM_FR 2             // frames per staging round

struct TileParams {
  const double* a_in;      // (n_clips, T, h[s]*w[s])
  int n_clips, T;
  int s;                   // number of pyrUp steps (= skip)
  int w[RM_MAX_LEVELS], h[RM_MAX_LEVELS];
  int tiles_x, tiles_y;
  double scale;            // 2^(-6 s)
  // pass 1 out / pass 2 in
  unsigned long long* minmax_keys;   // (n_clips, 4) order-preserving keys: raw min, raw max, avg min, avg max
  double threshold;        // temporal_threshold (pass 2)
  double* avg_out;         // (n_clips, H, W) pass 2
};

__device__ __forceinline__ unsigned long long f64_key(double v) {   // monotone map double -> uint64
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double key_f64(unsigned long long k) {
  unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}

// ranges of source indices a destination range [o0,o1] (inclusive) touches along one axis of pyrUp
__host__ __device__ __forceinline__ void src_range(int o0, int o1, int n, int* lo, int* hi) {
  int a = (o0 >> 1) - 1;
  if (a < 0) a = 0;
  int b = (o1 >> 1) + 1;
  if (b > n - 1) b = n - 1;
  if (n > 1 && b < 1) b = 1;   // reflect-101 of index -1 is index 1
  *lo = a;
  *hi = b;
}

struct PatchGeom {   // per tile: inclusive index ranges at levels 2..s, patch pitches and shared-memory offsets
  int x0[RM_MAX_LEVELS], x1[RM_MAX_LEVELS], y0[RM_MAX_LEVELS], y1[RM_MAX_LEVELS];
  int off[RM_MAX_LEVELS];   // doubles, per frame slot
  int per_frame;            // doubles per frame slot
};
__host__ __device__ inline PatchGeom make_patch(const TileParams& p, int tx, int ty) {
  PatchGeom g;
  // level-0 tile
  int X0 = tx * HM_TILE, Y0 = ty * HM_TILE;
  int X1 = min(p.w[0], X0 + HM_TILE) - 1, Y1 = min(p.h[0], Y0 + HM_TILE) - 1;
  int lx0 = X0, lx1 = X1, ly0 = Y0, ly1 = Y1;
  int acc = 0;
  for (int l = 1; l <= p.s; ++l) {
    int a, b, c, d;
    src_range(lx0, lx1, p.w[l], &a, &b);
    src_range(ly0, ly1, p.h[l], &c, &d);
    g.x0[l] = a; g.x1[l] = b; g.y0[l] = c; g.y1[l] = d;
    lx0 = a; lx1 = b; ly0 = c; ly1 = d;
    g.off[l] = 0;
    if (l >= 2 && l < p.s) {
      g.off[l] = acc;
      acc += (b - a + 1) * (d - c + 1);
    }
  }
  g.per_frame = acc;
  return g;
}

// One axis of the register stage.  A thread owns level-0 outputs 4i..4i+3; they read four level-1 "slots"
//   m0 = L1[2i-1] (index -1 -> 1), m1 = L1[2i], m2 = L1[min(2i+1, n1-1)], m3 = L1[min(2i+2, n1-1)]
// and the slots read four level-2 values v0..v3 at indices reflect101(i-1), i, min(i+1,n2-1), min(i+2,n2-1):
//   m0 = 4(v0+v1)   m1 = v0+6v1+v2   m2 = 4(v1+v2)   m3 = v1+6v2+v3       (pyrUp even/odd taps, App. A.2)
//   out0 = m0+6m1+m2   out1 = 4(m1+m2)   out2 = m1+6m2+m3   out3 = 4(m2+m3)
// At the far border the clamped slot is a copy of its neighbour: c2 (m2 := m1), c3 (0: as computed, 1: m3 := m2,
// 2: m3 := m1).  Everything is statically indexed, so it all lives in registers.
struct AxisGeom {
  int v[4];   // level-2 indices (absolute)
  int c2, c3;
};
__device__ __forceinline__ AxisGeom axis_geom(int i, int n1, int n2) {
  AxisGeom a;
  a.v[0] = reflect101(i - 1, n2);
  a.v[1] = min(i, n2 - 1);
  a.v[2] = min(i + 1, n2 - 1);
  a.v[3] = min(i + 2, n2 - 1);
  a.c2 = (2 * i + 1 > n1 - 1);
  a.c3 = (2 * i + 2 <= n1 - 1) ? 0 : ((n1 - 1 == 2 * i + 1) ? 1 : 2);
  return a;
}
__device__ __forceinline__ void slots4(double v0, double v1, double v2, double v3, int c2, int c3, double m[4]) {
  m[0] = 4.0 * (v0 + v1);
  m[1] = fma(6.0, v1, v0 + v2);
  double m2 = 4.0 * (v1 + v2);
  m[2] = c2 ? m[1] : m2;
  double m3 = fma(6.0, v2, v1 + v3);
  m[3] = (c3 == 0) ? m3 : (c3 == 1 ? m[2] : m[1]);
}
__device__ __forceinline__ void outs4(const double m[4], double o[4]) {
  o[0] = fma(6.0, m[1], m[0] + m[2]);
  o[1] = 4.0 * (m[1] + m[2]);
  o[2] = fma(6.0, m[2], m[1] + m[3]);
  o[3] = 4.0 * (m[2] + m[3]);
}

template <int PASS>
__global__ void __launch_bounds__(256) upsample_pass_kernel(const TileParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  __shared__ double red_a[8], red_b[8];
  const int tile = blockIdx.x;
  const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
  const int clip = blockIdx.y;
  const PatchGeom pg = make_patch(p, tx, ty);
  const int s = p.s;
  const int tid = threadIdx.x;
  const int bx = tid & 15, by = tid >> 4;
  const int X0 = tx * HM_TILE + 4 * bx, Y0 = ty * HM_TILE + 4 * by;   // my 4x4 output block
  const bool active = X0 < p.w[0] && Y0 < p.h[0];

  // level-2 source of the register stage: the last shared patch (s > 2) or A_s in global memory (s == 2)
  const int l2_pitch = (s > 2) ? (pg.x1[2] - pg.x0[2] + 1) : p.w[2];
  const int l2_x0 = (s > 2) ? pg.x0[2] : 0, l2_y0 = (s > 2) ? pg.y0[2] : 0;
  AxisGeom gx = axis_geom(active ? (X0 >> 2) : 0, p.w[1], p.w[2]);
  AxisGeom gy = axis_geom(active ? (Y0 >> 2) : 0, p.h[1], p.h[2]);
  int offs[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) offs[r][c] = (gy.v[r] - l2_y0) * l2_pitch + (gx.v[c] - l2_x0);

  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  double vmin = INFINITY, vmax = -INFINITY;
  double top_s = 0.0, repl_s = 0.0;
  if (PASS == 2) {
    double lo = key_f64(p.minmax_keys[clip * 4 + 0]), hi = key_f64(p.minmax_keys[clip * 4 + 1]);
    // transforms.py:185-189 on the scaled values; the comparison runs in the unscaled domain (scale is 2^-k: exact)
    double top = hi - (hi - lo) * p.threshold;
    top_s = top / p.scale;
    repl_s = lo / p.scale;
  }
  bool okx[4], oky[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    okx[k] = X0 + k < p.w[0];
    oky[k] = Y0 + k < p.h[0];
  }

  const int n_src = p.w[s] * p.h[s];
  const double* a_clip = p.a_in + (long long)clip * p.T * n_src;

  for (int t0 = 0; t0 < p.T; t0 += HM_FR) {
    const int nf = min(HM_FR, p.T - t0);
    // ---- shared-memory stages: level s (global) -> s-1 -> ... -> 2 (values unscaled: x64 per step) ------------------
    for (int l = s - 1; l >= 2; --l) {
      const int pw = pg.x1[l] - pg.x0[l] + 1, ph = pg.y1[l] - pg.y0[l] + 1;
      const int sw = p.w[l + 1], sh = p.h[l + 1];
      const bool from_global = (l + 1 == s);
      const int spitch = from_global ? sw : (pg.x1[l + 1] - pg.x0[l + 1] + 1);
      const int sx0 = from_global ? 0 : pg.x0[l + 1], sy0 = from_global ? 0 : pg.y0[l + 1];
      for (int i = tid; i < nf * pw * ph; i += blockDim.x) {
        int f = i / (pw * ph), r = i - f * pw * ph;
        int y = r / pw, x = r - y * pw;
        const double* src = from_global ? a_clip + (long long)(t0 + f) * n_src
                                        : sm + (size_t)f * pg.per_frame + pg.off[l + 1];
        Tap3 ax = tap3(pg.x0[l] + x, sw), ay = tap3(pg.y0[l] + y, sh);
        const double* r0 = src + (ay.i0 - sy0) * spitch - sx0;
        const double* r1 = src + (ay.i1 - sy0) * spitch - sx0;
        const double* r2 = src + (ay.i2 - sy0) * spitch - sx0;
        double h0 = up3(ax.odd, r0[ax.i0], r0[ax.i1], r0[ax.i2]);
        double h1 = up3(ax.odd, r1[ax.i0], r1[ax.i1], r1[ax.i2]);
        double h2 = up3(ax.odd, r2[ax.i0], r2[ax.i1], r2[ax.i2]);
        sm[(size_t)f * pg.per_frame + pg.off[l] + r] = up3(ay.odd, h0, h1, h2);
      }
      __syncthreads();
    }
    // ---- register stage: level 2 -> 1 -> 0, 4x4 outputs per thread ---------------------------------------------------
    if (active) {
      for (int f = 0; f < nf; ++f) {
        const double* l2 = (s > 2) ? sm + (size_t)f * pg.per_frame + pg.off[2] : a_clip + (long long)(t0 + f) * n_src;
        double hx[4][4];   // [level-2 row][level-1 x slot]
#pragma unroll
        for (int r = 0; r < 4; ++r)
          slots4(l2[offs[r][0]], l2[offs[r][1]], l2[offs[r][2]], l2[offs[r][3]], gx.c2, gx.c3, hx[r]);
        double l1[4][4];   // [level-1 y slot][level-1 x slot]
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double m[4];
          slots4(hx[0][c], hx[1][c], hx[2][c], hx[3][c], gy.c2, gy.c3, m);
#pragma unroll
          for (int r = 0; r < 4; ++r) l1[r][c] = m[r];
        }
        double ox[4][4];   // [level-1 y slot][level-0 x]
#pragma unroll
        for (int r = 0; r < 4; ++r) outs4(l1[r], ox[r]);
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
          double m[4] = {ox[0][kx], ox[1][kx], ox[2][kx], ox[3][kx]};
          double o[4];
          outs4(m, o);
#pragma unroll
          for (int ky = 0; ky < 4; ++ky) {
            const double v = o[ky];
            if (PASS == 1) {
              if (okx[kx] && oky[ky]) {
                vmin = fmin(vmin, v);
                vmax = fmax(vmax, v);
              }
            } else {
              acc[ky][kx] += (v >= top_s) ? repl_s : v;
            }
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue ---------------------------------------------------------------------------------------------------------
  if (PASS == 2) {
    vmin = INFINITY;
    vmax = -INFINITY;
    if (active) {
      double* dst = p.avg_out + (long long)clip * p.w[0] * p.h[0];
#pragma unroll
      for (int ky = 0; ky < 4; ++ky)
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
          if (okx[kx] && oky[ky]) {
            double avg = (acc[ky][kx] * p.scale) / (double)p.T;   // np.average over T (base.py:562)
            dst[(long long)(Y0 + ky) * p.w[0] + X0 + kx] = avg;
            vmin = fmin(vmin, avg);
            vmax = fmax(vmax, avg);
          }
        }
    }
  } else {
    vmin *= p.scale;   // exact: power of two
    vmax *= p.scale;
  }
  // block reduce, one atomic pair per block
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  if (lane == 0) { red_a[warp] = vmin; red_b[warp] = vmax; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) { vmin = fmin(vmin, red_a[w]); vmax = fmax(vmax, red_b[w]); }
    const int base = clip * 4 + (PASS == 1 ? 0 : 2);
    if (vmin <= vmax) {
      atomicMin(&p.minmax_keys[base + 0], f64_key(vmin));
      atomicMax(&p.minmax_keys[base + 1], f64_key(vmax));
    }
  }
}

__global__ void minmax_init_kernel(unsigned long long* keys, int n_clips) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_clips * 4) keys[i] = (i & 1) ? 0ull : ~0ull;   // min slots start at +max key, max slots at 0
}

// base.py:563-564: (avg - min) / (max - min) * 255, truncated to uint8 (NaN from a flat map becomes 0)
__global__ void heat_normalise_kernel(const double* __restrict__ avg, const unsigned long long* __restrict__ keys,
                                      uint8_t* __restrict__ heat, double* __restrict__ minmax_out, int n_clips,
                                      long long hw) {
  const int clip = blockIdx.y;
  const double mn = key_f64(keys[clip * 4 + 2]), mx = key_f64(keys[clip * 4 + 3]);
  const double range = mx - mn;
  const double* a = avg + clip * hw;
  uint8_t* o = heat + clip * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    double v = ((a[i] - mn) / range) * 255.0;
    o[i] = (v == v) ? (uint8_t)(int)v : (uint8_t)0;
  }
  if (minmax_out && blockIdx.x == 0 && threadIdx.x < 4)
    minmax_out[clip * 4 + threadIdx.x] = key_f64(keys[clip * 4 + threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------------- generic volume ops
// (stand-alone eulerian_magnification_bandpass API: transforms.py:184-192 and base.py:562 on materialised arrays)
__global__ void volume_minmax_kernel(const double* __restrict__ x, long long n, unsigned long long* keys) {
  double vmin = INFINITY, vmax = -INFINITY;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = x[i];
    vmin = fmin(vmin, v);
    vmax = fmax(vmax, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  if ((threadIdx.x & 31) == 0 && vmin <= vmax) {
    atomicMin(&keys[0], f64_key(vmin));
    atomicMax(&keys[1], f64_key(vmax));
  }
}
__global__ void volume_clip_kernel(const double* __restrict__ x, double* __restrict__ y, long long n,
                                   const unsigned long long* keys, double threshold) {
  const double lo = key_f64(keys[0]), hi = key_f64(keys[1]);
  const double top = hi - (hi - lo) * threshold;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = x[i];
    y[i] = (v >= top) ? lo : v;
  }
}
__global__ void volume_mean0_kernel(const double* __restrict__ x, double* __restrict__ out, int T, long long hw) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int t = 0; t < T; ++t) acc += x[(long long)t * hw + i];
    out[i] = acc / (double)T;
  }
}
__global__ void keys_to_f64_kernel(const unsigned long long* keys, double* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = key_f64(keys[i]);
}

extern "C" int32_t rm_volume_clip_mean(rm_handle* h, const double* raw, double* clipped_out, double* avg_out,
                                       double* minmax_out, int32_t T, int64_t hw, double threshold, void* workspace,
                                       void* stream) {
  RM_CHECK_ARG(h, h && raw && workspace && T >= 1 && hw >= 1, "null pointer or bad size");
  DeviceGuard dg(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(workspace);   // 4 keys
  RM_PROF(h, st, "minmax_init_kernel");
  minmax_init_kernel<<<1, 32, 0, st>>>(keys, 1);
  RM_LAUNCH_CHECK(h);
  long long n = (long long)T * hw;
  int grid = (int)((n + 255) / 256 < (long long)h->sm_count * 8 ? (n + 255) / 256 : (long long)h->sm_count * 8);
  RM_PROF(h, st, "volume_minmax_kernel");
  volume_minmax_kernel<<<grid, 256, 0, st>>>(raw, n, keys);
  RM_LAUNCH_CHECK(h);
  const double* mean_src = raw;
  if (clipped_out) {
    RM_PROF(h, st, "volume_clip_kernel");
    volume_clip_kernel<<<grid, 256, 0, st>>>(raw, clipped_out, n, keys, threshold);
    RM_LAUNCH_CHECK(h);
    mean_src = clipped_out;
  }
  if (avg_out) {
    int g2 = (int)((hw + 255) / 256);
    RM_PROF(h, st, "volume_mean0_kernel");
    volume_mean0_kernel<<<g2, 256, 0, st>>>(mean_src, avg_out, T, hw);
    RM_LAUNCH_CHECK(h);
  }
  if (minmax_out) {
    RM_PROF(h, st, "keys_to_f64_kernel");
    keys_to_f64_kernel<<<1, 32, 0, st>>>(keys, minmax_out, 2);
    RM_LAUNCH_CHECK(h);
  }
  return RM_OK;
}

// ---------------------------------------------------------------------------------------------------- host side
extern "C" int32_t rm_heatmap_workspace_bytes(rm_handle* h, int32_t W, int32_t H, int32_t n_clips, int32_t T, size_t* out) {
  RM_CHECK_ARG(h, h && out && W >= 1 && H >= 1 && n_clips >= 0 && T >= 1, "null pointer or bad size");
  LevelGeom g = make_geom(W, H, h->p.pyramid_levels);
  int s = h->p.skip_levels_at_top;
  size_t a = (size_t)n_clips * T * g.w[s] * g.h[s] * 8;       // A_skip
  size_t avg = (size_t)n_clips * W * H * 8;                    // time average
  size_t keys = (size_t)n_clips * 4 * 8;
  *out = a + avg + keys + 3 * 256;
  return RM_OK;
}

extern "C" int32_t rm_heatmap(rm_handle* h, const double* bp, int32_t n_clips, int32_t T, int32_t W, int32_t H,
                              uint8_t* heat_out, double* minmax_out, void* workspace, size_t workspace_bytes, void* stream) {
  RM_CHECK_ARG(h, h && bp && heat_out && n_clips >= 0 && T >= 1 && W >= 1 && H >= 1, "null pointer or bad size");
  const int L = h->p.pyramid_levels, s = h->p.skip_levels_at_top;
  if (s < 2 || s - 2 > HM_MAX_STAGES || L - 1 <= s || L > RM_MAX_LEVELS)
    return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: fused heat map needs 2 <= skip <= 6 and skip < levels-1", __func__);
  if (n_clips == 0) return RM_OK;
  size_t need = 0;
  rm_heatmap_workspace_bytes(h, W, H, n_clips, T, &need);
  if (!workspace || workspace_bytes < need)
    return rm_fail(h, RM_ERR_WORKSPACE, "%s: workspace too small (%lld needed, %lld given)", __func__, (long long)need,
                   (long long)workspace_bytes);
  DeviceGuard dg(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  LevelGeom g = make_geom(W, H, L);
  RecordGeom rec = make_record(g, s);
  uintptr_t base = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
  double* a_skip = reinterpret_cast<double*>(base);
  base += ((size_t)n_clips * T * g.w[s] * g.h[s] * 8 + 255) & ~(size_t)255;
  double* avg = reinterpret_cast<double*>(base);
  base += ((size_t)n_clips * W * H * 8 + 255) & ~(size_t)255;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(base);

  HeadParams hp;
  memset(&hp, 0, sizeof(hp));
  hp.bp = bp;
  hp.a_out = a_skip;
  hp.n_frames = (long long)n_clips * T;
  hp.first = rec.first;
  hp.last = rec.last;
  for (int l = 0; l < L; ++l) {
    hp.w[l] = g.w[l];
    hp.h[l] = g.h[l];
    hp.off[l] = rec.off[l];
  }
  hp.record_len = rec.len;
  int head_smem = rec.len * 8;
  if (head_smem > h->smem_optin) return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: record too large for shared memory", __func__);
  RM_CUDA(h, cudaFuncSetAttribute(collapse_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, head_smem));
  long long hgrid = hp.n_frames < (long long)h->sm_count * 8 ? hp.n_frames : (long long)h->sm_count * 8;
  RM_PROF(h, st, "collapse_head_kernel");
  collapse_head_kernel<<<(unsigned)hgrid, 256, head_smem, st>>>(hp);
  RM_LAUNCH_CHECK(h);

  RM_PROF(h, st, "minmax_init_kernel");
  minmax_init_kernel<<<div_up(n_clips * 4, 128), 128, 0, st>>>(keys, n_clips);
  RM_LAUNCH_CHECK(h);

  TileParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.a_in = a_skip;
  tp.n_clips = n_clips;
  tp.T = T;
  tp.s = s;
  for (int l = 0; l <= s; ++l) {
    tp.w[l] = g.w[l];
    tp.h[l] = g.h[l];
  }
  tp.tiles_x = (W + HM_TILE - 1) / HM_TILE;
  tp.tiles_y = (H + HM_TILE - 1) / HM_TILE;
  tp.scale = 1.0;
  for (int l = 0; l < s; ++l) tp.scale *= 1.0 / 64.0;
  tp.minmax_keys = keys;
  tp.threshold = h->p.temporal_threshold;
  tp.avg_out = avg;
  int max_per_frame = 0;
  for (int ty = 0; ty < tp.tiles_y; ++ty)
    for (int tx = 0; tx < tp.tiles_x; ++tx) {
      PatchGeom pg = make_patch(tp, tx, ty);
      if (pg.per_frame > max_per_frame) max_per_frame = pg.per_frame;
    }
  size_t smem = (size_t)max_per_frame * HM_FR * 8 + 16;
  if ((int)smem > h->smem_optin) return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: patch too large for shared memory", __func__);
  RM_CUDA(h, cudaFuncSetAttribute(upsample_pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RM_CUDA(h, cudaFuncSetAttribute(upsample_pass_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(tp.tiles_x * tp.tiles_y, n_clips);
  RM_PROF(h, st, "upsample_pass_kernel<1>");
  upsample_pass_kernel<1><<<grid, 256, smem, st>>>(tp);
  RM_LAUNCH_CHECK(h);
  RM_PROF(h, st, "upsample_pass_kernel<2>");
  upsample_pass_kernel<2><<<grid, 256, smem, st>>>(tp);
  RM_LAUNCH_CHECK(h);
  long long hw = (long long)W * H;
  dim3 ngrid((unsigned)((hw + 255) / 256 < 1024 ? (hw + 255) / 256 : 1024), n_clips);
  RM_PROF(h, st, "heat_normalise_kernel");
  heat_normalise_kernel<<<ngrid, 256, 0, st>>>(avg, keys, heat_out, minmax_out, n_clips, hw);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}
