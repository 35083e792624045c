// Laplacian/Gaussian pyramid kernels.
//
// Reference: pyramid.py:9-48 (create_gaussian_image_pyramid / create_laplacian_image_pyramid /
// create_laplacian_video_pyramid), i.e. chains of cv2.pyrDown / cv2.pyrUp on float64 images.
//
//  * pyr_down_f64_kernel / pyr_up_f64_kernel : one level, any size -- the stand-alone pyramid API.
//  * pyramid_front_kernel<T>                  : the hot kernel.  Streams a frame once from HBM and reduces it by
//    `n_steps` pyrDown levels (640x480 -> 40x30 for skip=4) without ever materialising levels 0..skip-1:
//    bands of rows are staged in shared memory (cp.async, 3 stages), and thread groups -- one per level, one
//    thread per output column, the 5-row vertical window kept in registers -- form a systolic pipeline with one
//    __syncthreads per band.  HBM traffic = the frame (read once) + the level-`skip` image (written once).
//  * pyramid_tail_kernel                      : per frame, G_skip -> G_skip+1..G_top and the Laplacian levels
//    skip..top-1 that transforms.py:156-170 actually filters, packed into one record per frame.
#include "common.cuh"
#include "pyramid_u8.cuh"

// ---------------------------------------------------------------------------------------------------- single level
__global__ void pyr_down_f64_kernel(const double* __restrict__ src, double* __restrict__ dst, long long n_img, int sw,
                                    int sh, int dw, int dh) {
  long long total = n_img * dw * dh;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int x = (int)(idx % dw);
    int y = (int)((idx / dw) % dh);
    long long img = idx / ((long long)dw * dh);
    const double* s = src + img * sw * sh;
    int xs[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) xs[k] = reflect101(2 * x + k - 2, sw);
    double r[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const double* row = s + (long long)reflect101(2 * y + k - 2, sh) * sw;
      r[k] = tap5(row[xs[0]], row[xs[1]], row[xs[2]], row[xs[3]], row[xs[4]]);
    }
    dst[idx] = tap5(r[0], r[1], r[2], r[3], r[4]) * (1.0 / 256.0);
  }
}

__global__ void pyr_up_f64_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                  const double* __restrict__ other, int mode, long long n_img, int sw, int sh, int dw,
                                  int dh) {
  long long total = n_img * dw * dh;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int x = (int)(idx % dw);
    int y = (int)((idx / dw) % dh);
    long long img = idx / ((long long)dw * dh);
    const double* s = src + img * sw * sh;
    UpTaps tx = up_taps(x, sw), ty = up_taps(y, sh);
    const double* r0 = s + (long long)ty.i0 * sw;
    const double* r1 = s + (long long)ty.i1 * sw;
    const double* r2 = s + (long long)ty.i2 * sw;
    double h0 = up_combine(tx, r0[tx.i0], r0[tx.i1], r0[tx.i2]);
    double h1 = up_combine(tx, r1[tx.i0], r1[tx.i1], r1[tx.i2]);
    double h2 = up_combine(tx, r2[tx.i0], r2[tx.i1], r2[tx.i2]);
    double up = up_combine(ty, h0, h1, h2) * (1.0 / 64.0);
    double o = (mode == 0) ? up : (mode == 1 ? other[idx] - up : up + other[idx]);
    dst[idx] = o;
  }
}

template <typename T>
__global__ void to_f64_kernel(const T* __restrict__ src, double* __restrict__ dst, long long n, double scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (double)src[i] * scale;
}

// ---------------------------------------------------------------------------------------------------- fused front
#define FRONT_MAX_STEPS 6
#define FRONT_STAGES 3

struct FrontParams {
  const void* frames;
  double* g_out;           // (n_frames, h[n_steps], w[n_steps]) float64
  long long n_frames;
  long long frame_elems;   // W*H
  // frame f of the batch is source frame (f / seg_len) * seg_stride + seg_first + f % seg_len: a window of every clip
  long long seg_len, seg_stride, seg_first;
  int n_steps;             // number of pyrDown steps (= skip_levels_at_top)
  int band_rows;           // level-0 rows per tick
  int n_bands;             // ceil(H / band_rows)
  int lw[FRONT_MAX_STEPS + 1], lh[FRONT_MAX_STEPS + 1];   // level sizes 0..n_steps
  int max_final_cols;      // strip width in final-level columns
  int vec_ok;              // 16-byte cp.async path usable (alignment)
  double out_scale;        // 2^(-8 n_steps) [* 1/255 for u8]
};

__device__ __forceinline__ int rows_emitted(int R, int Hl) {  // output rows complete once R input rows are in
  if (R <= 0) return 0;
  return (R == Hl) ? (Hl + 1) / 2 : (R - 1) / 2;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <typename T>
__device__ __forceinline__ double load_px(const unsigned char* row, int c) {
  return (double)reinterpret_cast<const T*>(row)[c];
}

// Strip geometry, identical on host and device.
struct StripGeom {
  int lo[FRONT_MAX_STEPS + 1], hi[FRONT_MAX_STEPS + 1];   // column range [lo,hi) needed at each level
  int goff[FRONT_MAX_STEPS], gn[FRONT_MAX_STEPS];         // thread offset / count of group g (computes level g+1)
  int n_threads;
  int ring_cap[FRONT_MAX_STEPS + 1];                      // rows in the shared ring of level g (1..n_steps-1)
  int ring_off[FRONT_MAX_STEPS + 1];                      // byte offset of ring g in dynamic smem
  int band_x0;                                            // first level-0 column staged (aligned down)
  int band_pitch;                                         // bytes per staged row
  int band_bytes;                                         // bytes per stage
  int smem_bytes;
};
__host__ __device__ inline StripGeom make_strip(const FrontParams& p, int strip, int elem_size) {
  StripGeom s;
  int ns = p.n_steps;
  int c0 = strip * p.max_final_cols;
  int c1 = min(p.lw[ns], c0 + p.max_final_cols);
  s.lo[ns] = c0;
  s.hi[ns] = c1;
  for (int g = ns - 1; g >= 0; --g) {
    s.lo[g] = max(0, 2 * s.lo[g + 1] - 2);
    s.hi[g] = min(p.lw[g], 2 * (s.hi[g + 1] - 1) + 3);
  }
  int off = 0;
  for (int g = 0; g < ns; ++g) {
    s.goff[g] = off;
    s.gn[g] = s.hi[g + 1] - s.lo[g + 1];
    off += (s.gn[g] + 31) & ~31;
  }
  s.n_threads = off;
  int align_elems = 16 / elem_size;
  s.band_x0 = (s.lo[0] / align_elems) * align_elems;
  s.band_pitch = (((s.hi[0] - s.band_x0) * elem_size + 15) / 16) * 16;
  s.band_bytes = s.band_pitch * p.band_rows;
  int bytes = s.band_bytes * FRONT_STAGES;
  int rows = p.band_rows;
  for (int g = 1; g < ns; ++g) {
    rows = (rows + 1) / 2;
    s.ring_cap[g] = 2 * (rows + 2);
    s.ring_off[g] = bytes;
    bytes += s.ring_cap[g] * (s.hi[g] - s.lo[g]) * 8;
  }
  s.smem_bytes = bytes;
  return s;
}

template <typename T>
__global__ void __launch_bounds__(1024, 1) pyramid_front_kernel(const FrontParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const StripGeom sg = make_strip(p, blockIdx.x, (int)sizeof(T));
  const int ns = p.n_steps;
  const int tid = threadIdx.x;

  // which group am I in, which column do I own
  int g = -1, tg = 0;
  for (int k = 0; k < ns; ++k)
    if (tid >= sg.goff[k] && tid < sg.goff[k] + sg.gn[k]) {
      g = k;
      tg = tid - sg.goff[k];
    }
  const int x_out = (g >= 0) ? sg.lo[g + 1] + tg : 0;                 // column at level g+1
  int cin[5] = {0, 0, 0, 0, 0};                                       // input columns, relative to the staged row
  if (g >= 0) {
    const int base = (g == 0) ? sg.band_x0 : sg.lo[g];
#pragma unroll
    for (int k = 0; k < 5; ++k) cin[k] = reflect101(2 * x_out + k - 2, p.lw[g]) - base;
  }
  const int Hin = (g >= 0) ? p.lh[g] : 1;
  const int in_pitch = (g > 0) ? (sg.hi[g] - sg.lo[g]) : 0;           // doubles per ring row (g>0)
  const double* ring_in = (g > 0) ? reinterpret_cast<const double*>(smem + sg.ring_off[g]) : nullptr;
  const int in_cap = (g > 0) ? sg.ring_cap[g] : 1;
  const bool last_group = (g == ns - 1);
  double* ring_out = (g >= 0 && !last_group) ? reinterpret_cast<double*>(smem + sg.ring_off[g + 1]) : nullptr;
  const int out_pitch = (g >= 0 && !last_group) ? (sg.hi[g + 1] - sg.lo[g + 1]) : 0;
  const int out_cap = (g >= 0 && !last_group) ? sg.ring_cap[g + 1] : 1;
  const int Hout = (Hin + 1) / 2;

  // frames handled by this CTA: blockIdx.y, blockIdx.y + gridDim.y, ...
  const long long n_local = (p.n_frames - blockIdx.y + gridDim.y - 1) / gridDim.y;
  const long long n_band_ticks = n_local * p.n_bands;
  const long long n_ticks = n_band_ticks + (ns - 1);

  // cooperative staging of band tick v into stage v % FRONT_STAGES
  const int chunks_per_row = sg.band_pitch / 16;
  const int row_valid_elems = min(p.lw[0], sg.band_x0 + sg.band_pitch / (int)sizeof(T)) - sg.band_x0;
  auto stage_band = [&](long long v) {
    if (v < n_band_ticks) {
      long long fl = v / p.n_bands;
      int j = (int)(v % p.n_bands);
      long long frame = blockIdx.y + fl * gridDim.y;
      const long long sframe = (frame / p.seg_len) * p.seg_stride + p.seg_first + frame % p.seg_len;
      const T* fsrc = reinterpret_cast<const T*>(p.frames) + sframe * p.frame_elems;
      unsigned char* dst = smem + (size_t)(v % FRONT_STAGES) * sg.band_bytes;
      int r_begin = j * p.band_rows;
      int n_rows = min(p.band_rows, p.lh[0] - r_begin);
      if (p.vec_ok) {
        int total = n_rows * chunks_per_row;
        for (int c = tid; c < total; c += blockDim.x) {
          int r = c / chunks_per_row, cc = c - r * chunks_per_row;
          int e0 = cc * (16 / (int)sizeof(T));
          int valid = min(16, max(0, (row_valid_elems - e0) * (int)sizeof(T)));
          const T* src = fsrc + (long long)(r_begin + r) * p.lw[0] + sg.band_x0 + e0;
          cp_async16(dst + r * sg.band_pitch + cc * 16, valid > 0 ? (const void*)src : (const void*)fsrc, valid);
        }
      } else {
        int total = n_rows * row_valid_elems;
        for (int c = tid; c < total; c += blockDim.x) {
          int r = c / row_valid_elems, e = c - r * row_valid_elems;
          reinterpret_cast<T*>(dst + r * sg.band_pitch)[e] = fsrc[(long long)(r_begin + r) * p.lw[0] + sg.band_x0 + e];
        }
      }
    }
    cp_async_commit();
  };

  stage_band(0);
  stage_band(1);
  cp_async_wait<1>();
  __syncthreads();

  double hr0 = 0, hr1 = 0, hr2 = 0, hr3 = 0, hr4 = 0;   // horizontally filtered rows r-4..r of my column

  for (long long tick = 0; tick < n_ticks; ++tick) {
    stage_band(tick + 2);

    const long long v = tick - g;    // the band tick my group works on now
    if (g >= 0 && v >= 0 && v < n_band_ticks) {
      const long long fl = v / p.n_bands;
      const int j = (int)(v % p.n_bands);
      // rows of my input level that became available with band j of this frame
      int Rprev = (j == 0) ? 0 : min(p.lh[0], j * p.band_rows);
      int Rcur = min(p.lh[0], (j + 1) * p.band_rows);
      for (int k = 0; k < g; ++k) {
        Rprev = rows_emitted(Rprev, p.lh[k]);
        Rcur = rows_emitted(Rcur, p.lh[k]);
      }
      const unsigned char* band = smem + (size_t)(v % FRONT_STAGES) * sg.band_bytes;
      const long long in_base = fl * Hin;      // running row counters keep ring slots of consecutive frames apart
      const long long out_base = fl * Hout;
      const long long frame = blockIdx.y + fl * gridDim.y;
      double* gdst = p.g_out + frame * ((long long)p.lw[ns] * p.lh[ns]);

      for (int r = Rprev; r < Rcur; ++r) {
        double a, b, c, d, e;
        if (g == 0) {
          const unsigned char* row = band + (r - j * p.band_rows) * sg.band_pitch;
          a = load_px<T>(row, cin[0]); b = load_px<T>(row, cin[1]); c = load_px<T>(row, cin[2]);
          d = load_px<T>(row, cin[3]); e = load_px<T>(row, cin[4]);
        } else {
          const double* row = ring_in + (size_t)((in_base + r) % in_cap) * in_pitch;
          a = row[cin[0]]; b = row[cin[1]]; c = row[cin[2]]; d = row[cin[3]]; e = row[cin[4]];
        }
        hr0 = hr1; hr1 = hr2; hr2 = hr3; hr3 = hr4;
        hr4 = tap5(a, b, c, d, e);

        // vertical 5-tap: emit every output row whose window is complete (reflect-101 at both ends)
        double o0 = 0, o1 = 0;
        int y0 = -1, y1 = -1;
        if (r >= 2 && !(r & 1)) {
          y0 = (r - 2) >> 1;
          o0 = (y0 == 0) ? tap5(hr4, hr3, hr2, hr3, hr4) : tap5(hr0, hr1, hr2, hr3, hr4);
        }
        if (r == Hin - 1) {
          if (Hin == 1) { y1 = 0; o1 = tap5(hr4, hr4, hr4, hr4, hr4); }
          else if (Hin == 2) { y1 = 0; o1 = tap5(hr3, hr4, hr3, hr4, hr3); }
          else if (Hin & 1) { y1 = r >> 1; o1 = tap5(hr2, hr3, hr4, hr3, hr2); }
          else { y1 = (r - 1) >> 1; o1 = tap5(hr1, hr2, hr3, hr4, hr3); }
        }
        if (last_group) {
          if (y0 >= 0) gdst[(long long)y0 * p.lw[ns] + x_out] = o0 * p.out_scale;
          if (y1 >= 0) gdst[(long long)y1 * p.lw[ns] + x_out] = o1 * p.out_scale;
        } else {
          if (y0 >= 0) ring_out[(size_t)((out_base + y0) % out_cap) * out_pitch + tg] = o0;
          if (y1 >= 0) ring_out[(size_t)((out_base + y1) % out_cap) * out_pitch + tg] = o1;
        }
      }
    }
    cp_async_wait<1>();   // band tick+1 has landed (tick+2 may still be in flight)
    __syncthreads();
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------------- tail
struct TailParams {
  const double* g_in;     // (n_frames, h[first], w[first]), or null when the integer front ran:
  const uint32_t* g3_in;  // (n_frames, h[first-1], w[first-1]) exact integers, level `first` = pyrDown * g_scale
  double g_scale;
  double* lap_out;        // (n_frames, record_len)
  long long n_frames;
  int first, top;         // G levels first..top are built; Laplacian levels first..top-1 are written
  int w[RM_MAX_LEVELS], h[RM_MAX_LEVELS], rec_off[RM_MAX_LEVELS];
  int record_len;
};

__global__ void __launch_bounds__(256) pyramid_tail_kernel(const TailParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* g = reinterpret_cast<double*>(smem_raw);
  int goff[RM_MAX_LEVELS];
  int acc = 0;
  for (int l = p.first; l <= p.top; ++l) {
    goff[l] = acc;
    acc += p.w[l] * p.h[l];
  }
  for (long long f = blockIdx.x; f < p.n_frames; f += gridDim.x) {
    const int n0 = p.w[p.first] * p.h[p.first];
    if (p.g3_in) {
      // integer level first-1 -> level first: the sums stay exact in float64 (< 2^53), one scale at the end
      const int sw = p.w[p.first - 1], sh = p.h[p.first - 1], dw = p.w[p.first], dh = p.h[p.first];
      uint32_t* s = reinterpret_cast<uint32_t*>(g + acc);     // staged as integers: half the shared memory of float64
      const uint32_t* src = p.g3_in + f * (long long)(sw * sh);
      // 16-byte loads, all of a thread's loads in flight before the first store (the staging loop held 38 % of this
      // kernel's stall samples, ncu r01k: 0.300 -> 0.232 ms per 8192 VGA frames, r02a)
      if (((sw * sh) & 3) == 0 && (acc & 1) == 0) {      // 16-byte aligned on both sides
        const uint4* src4 = reinterpret_cast<const uint4*>(src);
        uint4* s4 = reinterpret_cast<uint4*>(s);
        const int n4 = (sw * sh) >> 2;
#pragma unroll 1
        for (int i0 = threadIdx.x; i0 < n4; i0 += 8 * blockDim.x) {
          uint4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (i0 + u * (int)blockDim.x < n4) v[u] = src4[i0 + u * blockDim.x];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (i0 + u * (int)blockDim.x < n4) s4[i0 + u * blockDim.x] = v[u];
        }
      } else
      for (int i = threadIdx.x; i < sw * sh; i += blockDim.x) s[i] = src[i];
      __syncthreads();
      for (int i = threadIdx.x; i < dw * dh; i += blockDim.x) {
        int x = i % dw, y = i / dw;
        double r[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const uint32_t* row = s + reflect101(2 * y + k - 2, sh) * sw;
          r[k] = tap5((double)row[reflect101(2 * x - 2, sw)], (double)row[reflect101(2 * x - 1, sw)], (double)row[2 * x],
                      (double)row[reflect101(2 * x + 1, sw)], (double)row[reflect101(2 * x + 2, sw)]);
        }
        g[i] = tap5(r[0], r[1], r[2], r[3], r[4]) * p.g_scale;
      }
    } else {
      const double* src = p.g_in + f * n0;
      for (int i = threadIdx.x; i < n0; i += blockDim.x) g[i] = src[i];
    }
    __syncthreads();
    for (int l = p.first; l < p.top; ++l) {   // G_{l+1} = pyrDown(G_l)
      const double* s = g + goff[l];
      double* d = g + goff[l + 1];
      const int sw = p.w[l], sh = p.h[l], dw = p.w[l + 1], dh = p.h[l + 1];
      for (int i = threadIdx.x; i < dw * dh; i += blockDim.x) {
        int x = i % dw, y = i / dw;
        double r[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const double* row = s + reflect101(2 * y + k - 2, sh) * sw;
          r[k] = tap5(row[reflect101(2 * x - 2, sw)], row[reflect101(2 * x - 1, sw)], row[2 * x],
                      row[reflect101(2 * x + 1, sw)], row[reflect101(2 * x + 2, sw)]);
        }
        d[i] = tap5(r[0], r[1], r[2], r[3], r[4]) * (1.0 / 256.0);
      }
      __syncthreads();
    }
    double* out = p.lap_out + f * p.record_len;
    for (int l = p.first; l < p.top; ++l) {   // L_l = G_l - pyrUp(G_{l+1})   (pyramid.py:24-26)
      const double* s = g + goff[l + 1];
      const double* cur = g + goff[l];
      const int sw = p.w[l + 1], sh = p.h[l + 1], dw = p.w[l], dh = p.h[l];
      for (int i = threadIdx.x; i < dw * dh; i += blockDim.x) {
        int x = i % dw, y = i / dw;
        UpTaps tx = up_taps(x, sw), ty = up_taps(y, sh);
        const double* r0 = s + ty.i0 * sw;
        const double* r1 = s + ty.i1 * sw;
        const double* r2 = s + ty.i2 * sw;
        double h0 = up_combine(tx, r0[tx.i0], r0[tx.i1], r0[tx.i2]);
        double h1 = up_combine(tx, r1[tx.i0], r1[tx.i1], r1[tx.i2]);
        double h2 = up_combine(tx, r2[tx.i0], r2[tx.i1], r2[tx.i2]);
        out[p.rec_off[l] + i] = cur[i] - up_combine(ty, h0, h1, h2) * (1.0 / 64.0);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------- host side
static inline int grid_for(long long total, int block, int sm_count) {
  long long blocks = (total + block - 1) / block;
  long long cap = (long long)sm_count * 16;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

extern "C" int32_t rm_to_f64(rm_handle* h, const void* src, int32_t dtype, double* dst, int64_t n, void* stream) {
  RM_CHECK_ARG(h, h && src && dst && n >= 0, "null pointer or negative size");
  if (n == 0) return RM_OK;
  DeviceGuard dg(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(n, 256, h->sm_count);
  if (dtype == RM_U8) { RM_PROF(h, st, "to_f64_kernel<uint8_t>"); to_f64_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)src, dst, n, 1.0 / 255); }
  else if (dtype == RM_F32) { RM_PROF(h, st, "to_f64_kernel<float>"); to_f64_kernel<float><<<grid, 256, 0, st>>>((const float*)src, dst, n, 1.0); }
  else if (dtype == RM_F64) { RM_PROF(h, st, "to_f64_kernel<double>"); to_f64_kernel<double><<<grid, 256, 0, st>>>((const double*)src, dst, n, 1.0); }
  else return rm_fail(h, RM_ERR_INVALID, "%s: unknown dtype", __func__);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

// float_to_uint8 (transforms.py:26-29): img * 255 stored into a uint8 array, i.e. truncated toward zero.
__global__ void f64_to_u8_kernel(const double* __restrict__ src, uint8_t* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (uint8_t)(int)(src[i] * 255.0);
}
extern "C" int32_t rm_f64_to_u8(rm_handle* h, const double* src, uint8_t* dst, int64_t n, void* stream) {
  RM_CHECK_ARG(h, h && src && dst && n >= 0, "null pointer or negative size");
  if (n == 0) return RM_OK;
  DeviceGuard dg(h->device);
  RM_PROF(h, (cudaStream_t)stream, "f64_to_u8_kernel");
  f64_to_u8_kernel<<<grid_for(n, 256, h->sm_count), 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

extern "C" int32_t rm_pyr_down_f64(rm_handle* h, const double* src, double* dst, int64_t n_img, int32_t sw, int32_t sh,
                                   void* stream) {
  RM_CHECK_ARG(h, h && src && dst && n_img >= 0 && sw >= 1 && sh >= 1, "null pointer or bad size");
  if (n_img == 0) return RM_OK;
  DeviceGuard dg(h->device);
  int dw = (sw + 1) / 2, dh = (sh + 1) / 2;
  int grid = grid_for(n_img * dw * dh, 256, h->sm_count);
  RM_PROF(h, (cudaStream_t)stream, "pyr_down_f64_kernel");
  pyr_down_f64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, n_img, sw, sh, dw, dh);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

extern "C" int32_t rm_pyr_up_f64(rm_handle* h, const double* src, double* dst, const double* other, int32_t mode,
                                 int64_t n_img, int32_t sw, int32_t sh, int32_t dw, int32_t dh, void* stream) {
  RM_CHECK_ARG(h, h && src && dst && n_img >= 0 && sw >= 1 && sh >= 1, "null pointer or bad size");
  RM_CHECK_ARG(h, mode >= 0 && mode <= 2 && (mode == 0 || other), "mode needs `other`");
  // cv2.pyrUp accepts dstsize with |dst - 2 src| == dst % 2 per axis
  RM_CHECK_ARG(h, (dw == 2 * sw || (dw == 2 * sw - 1 && (dw & 1))) && (dh == 2 * sh || (dh == 2 * sh - 1 && (dh & 1))),
               "dstsize must be 2n or 2n-1 (odd)");
  if (n_img == 0) return RM_OK;
  DeviceGuard dg(h->device);
  int grid = grid_for(n_img * dw * dh, 256, h->sm_count);
  RM_PROF(h, (cudaStream_t)stream, "pyr_up_f64_kernel");
  pyr_up_f64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, other, mode, n_img, sw, sh, dw, dh);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

extern "C" int32_t rm_lap_record_len(rm_handle* h, int32_t W, int32_t H, int64_t* out) {
  RM_CHECK_ARG(h, h && out && W >= 1 && H >= 1, "null pointer or bad size");
  LevelGeom g = make_geom(W, H, h->p.pyramid_levels);
  *out = make_record(g, h->p.skip_levels_at_top).len;
  return RM_OK;
}

extern "C" int32_t rm_pyramid_workspace_bytes(rm_handle* h, int32_t W, int32_t H, int64_t n_frames, size_t* out) {
  RM_CHECK_ARG(h, h && out && W >= 1 && H >= 1 && n_frames >= 0, "null pointer or bad size");
  LevelGeom g = make_geom(W, H, h->p.pyramid_levels);
  int s = h->p.skip_levels_at_top;
  size_t a = (size_t)n_frames * g.w[s] * g.h[s] * sizeof(double);                    // level `skip`, float64
  size_t b = s >= 1 ? (size_t)n_frames * g.w[s - 1] * g.h[s - 1] * sizeof(uint32_t) : 0;   // level skip-1, integers
  *out = (a > b ? a : b) + 256;
  return RM_OK;
}

static int pick_band_rows(int W, int elem) {
  int row_bytes = W * elem;
  int rows = 8;
  while (rows > 2 && rows * row_bytes * FRONT_STAGES > 96 * 1024) rows >>= 1;
  return rows;
}

template <typename T>
static int32_t launch_front(rm_handle* h, FrontParams& p, int n_strips, cudaStream_t st) {
  int max_threads = 0, max_smem = 0;
  for (int s = 0; s < n_strips; ++s) {
    StripGeom sg = make_strip(p, s, (int)sizeof(T));
    if (sg.n_threads > max_threads) max_threads = sg.n_threads;
    if (sg.smem_bytes > max_smem) max_smem = sg.smem_bytes;
  }
  if (max_threads > 1024 || max_smem > h->smem_optin)
    return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: strip needs %lld threads / %lld B shared memory", __func__, max_threads,
                   max_smem);
  RM_CUDA(h, cudaFuncSetAttribute(pyramid_front_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  int occ = 1;
  RM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pyramid_front_kernel<T>, max_threads, max_smem));
  if (occ < 1) occ = 1;
  long long ctas = (long long)h->sm_count * occ;
  long long gy = ctas / n_strips;
  if (gy < 1) gy = 1;
  if (gy > p.n_frames) gy = p.n_frames;
  if (gy > 65535) gy = 65535;
  dim3 grid(n_strips, (unsigned)gy);
  RM_PROF(h, st, sizeof(T) == 1 ? "pyramid_front_kernel<u8>" : (sizeof(T) == 4 ? "pyramid_front_kernel<f32>" : "pyramid_front_kernel<f64>"));
  pyramid_front_kernel<T><<<grid, max_threads, max_smem, st>>>(p);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

static int32_t pyramid_build_impl(rm_handle* h, const void* frames, int32_t dtype, int64_t n_frames, int64_t seg_len,
                                  int64_t seg_stride, int64_t seg_first, int32_t W, int32_t H, double* lap_out,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  RM_CHECK_ARG(h, h && frames && lap_out && W >= 1 && H >= 1 && n_frames >= 0, "null pointer or bad size");
  RM_CHECK_ARG(h, dtype == RM_U8 || dtype == RM_F32 || dtype == RM_F64 || dtype == RM_BGR8, "unknown dtype");
  const int L = h->p.pyramid_levels, s = h->p.skip_levels_at_top;
  if (s < 1 || s > FRONT_MAX_STEPS || L - 1 <= s || L > RM_MAX_LEVELS)
    return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: fused path needs 1 <= skip <= 6 and skip < levels-1", __func__);
  if (n_frames == 0) return RM_OK;
  size_t need = 0;
  rm_pyramid_workspace_bytes(h, W, H, n_frames, &need);
  if (!workspace || workspace_bytes < need)
    return rm_fail(h, RM_ERR_WORKSPACE, "%s: workspace too small (%lld needed, %lld given)", __func__, (long long)need,
                   (long long)workspace_bytes);
  DeviceGuard dg(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  LevelGeom g = make_geom(W, H, L);
  RecordGeom rec = make_record(g, s);
  double* g_skip = reinterpret_cast<double*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);

  const bool bgr = dtype == RM_BGR8;
  if (bgr && !pu_supported(frames, W, H, s))
    return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: BGR frames need W, H multiples of 8 and H >= 24 (else rm_bgr_to_gray first)", __func__);
  const bool integer_front = bgr || (dtype == RM_U8 && !h->force_generic_front && pu_supported(frames, W, H, s));
  if (integer_front) {
    if (!bgr && pu_best_mode(h, frames, W, H))      // one kernel: the record is written, there is no tail launch
      return pu_launch_fused(h, (const uint8_t*)frames, lap_out, n_frames, seg_len, seg_stride, seg_first, W, H, st);
    int32_t rc = pu_launch_front(h, (const uint8_t*)frames, bgr ? 1 : 0, reinterpret_cast<uint32_t*>(g_skip), n_frames,
                                 seg_len, seg_stride, seg_first, W, H, st);
    if (rc != RM_OK) return rc;
  } else {
    const int elem = dtype == RM_U8 ? 1 : (dtype == RM_F32 ? 4 : 8);
    FrontParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.frames = frames;
    fp.g_out = g_skip;
    fp.n_frames = n_frames;
    fp.frame_elems = (long long)W * H;
    fp.seg_len = seg_len;
    fp.seg_stride = seg_stride;
    fp.seg_first = seg_first;
    fp.n_steps = s;
    fp.band_rows = pick_band_rows(W < 704 ? W : 704, elem);
    fp.n_bands = (H + fp.band_rows - 1) / fp.band_rows;
    for (int l = 0; l <= s; ++l) {
      fp.lw[l] = g.w[l];
      fp.lh[l] = g.h[l];
    }
    fp.max_final_cols = 40;
    fp.vec_ok = (((uintptr_t)frames % 16) == 0 && ((long long)W * elem) % 16 == 0 && ((long long)W * H * elem) % 16 == 0);
    fp.out_scale = 1.0;
    for (int l = 0; l < s; ++l) fp.out_scale *= 1.0 / 256.0;
    if (dtype == RM_U8) fp.out_scale *= 1.0 / 255;
    int n_strips = (g.w[s] + fp.max_final_cols - 1) / fp.max_final_cols;
    int32_t rc;
    if (dtype == RM_U8) rc = launch_front<uint8_t>(h, fp, n_strips, st);
    else if (dtype == RM_F32) rc = launch_front<float>(h, fp, n_strips, st);
    else rc = launch_front<double>(h, fp, n_strips, st);
    if (rc != RM_OK) return rc;

  }
  TailParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.g_in = integer_front ? nullptr : g_skip;
  tp.g3_in = integer_front ? reinterpret_cast<const uint32_t*>(g_skip) : nullptr;
  tp.g_scale = 1.0;
  for (int l = 0; l < s; ++l) tp.g_scale *= 1.0 / 256.0;
  tp.g_scale *= 1.0 / 255;
  tp.lap_out = lap_out;
  tp.n_frames = n_frames;
  tp.first = s;
  tp.top = L - 1;
  int tail_elems = 0;
  for (int l = 0; l < L; ++l) {
    tp.w[l] = g.w[l];
    tp.h[l] = g.h[l];
    tp.rec_off[l] = rec.off[l];
    if (l >= s) tail_elems += g.w[l] * g.h[l];
  }
  tp.record_len = rec.len;
  int tail_smem = tail_elems * (int)sizeof(double);
  if (integer_front) tail_smem += g.w[s - 1] * g.h[s - 1] * (int)sizeof(uint32_t);
  if (tail_smem > h->smem_optin)
    return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: level %lld image too large for the tail kernel (%lld B)", __func__, s,
                   tail_smem);
  RM_CUDA(h, cudaFuncSetAttribute(pyramid_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tail_smem));
  long long tgrid = n_frames < (long long)h->sm_count * 8 ? n_frames : (long long)h->sm_count * 8;
  RM_PROF(h, st, "pyramid_tail_kernel");
  pyramid_tail_kernel<<<(unsigned)tgrid, 256, tail_smem, st>>>(tp);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

extern "C" int32_t rm_pyramid_build(rm_handle* h, const void* frames, int32_t dtype, int64_t n_frames, int32_t W,
                                    int32_t H, double* lap_out, void* workspace, size_t workspace_bytes, void* stream) {
  return pyramid_build_impl(h, frames, dtype, n_frames, n_frames > 0 ? n_frames : 1, 0, 0, W, H, lap_out, workspace,
                            workspace_bytes, stream);
}

extern "C" int32_t rm_pyramid_build_clips(rm_handle* h, const void* frames, int32_t dtype, int32_t n_clips, int32_t T,
                                          int32_t first_frame, int32_t n_frames, int32_t W, int32_t H, double* lap_out,
                                          void* workspace, size_t workspace_bytes, void* stream) {
  RM_CHECK_ARG(h, h && n_clips >= 0 && T >= 1 && first_frame >= 0 && n_frames >= 1 && first_frame + n_frames <= T,
               "frame window outside the clip");
  return pyramid_build_impl(h, frames, dtype, (int64_t)n_clips * n_frames, n_frames, T, first_frame, W, H, lap_out,
                            workspace, workspace_bytes, stream);
}
