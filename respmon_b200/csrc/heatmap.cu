// Calibration heat map: collapse of the band-passed pyramid, intensity clip, time average, normalise to uint8.
//
// Reference: collapse_laplacian_video_pyramid (pyramid.py:51-69) called from transforms.py:182 on a pyramid whose
// levels 0..skip-1 and top are all zero, then transforms.py:184-192 (global min/max over (T,H,W), top = max -
// (max-min)*threshold, values >= top replaced by min) and base.py:562-564 (mean over T, min-max normalise, *255
// truncated to uint8).
//
// Nothing of size (T,H,W) is ever stored: the collapsed full-resolution value of a pixel-frame is evaluated on the fly,
// up to twice (pass 1: min/max, pass 2: clipped mean).  Per frame the coarse part (levels top-1..skip, 1600 values at
// VGA) is collapsed once by collapse_head_kernel into A_skip (40x30) and upsampled to level 3 (level 2 when pruning is
// off) by up_level_kernel.  The two tile passes own 64x32 output tiles: per evaluated (tile, frame) the level-3 patch is
// staged in shared memory by a cp.async ring, expanded to the tile's level-2 patch, and the last two pyrUp steps (16x
// the pixels) run entirely in registers, 4x4 outputs per thread.  pyrUp is a convex combination, so the level-`skip`
// values under a tile bound everything the tile can produce: tile-frames that cannot change the clip's min/max (pass 1)
// or whose values are all clipped (pass 2) are not evaluated at all -- exactly, not approximately (tile_bounds_kernel,
// minmax_seed_kernel).  All 1/64 factors are exact powers of two and are folded into one final scale.  Where evaluated
// the passes are FP64-ALU bound (about 11 flop per pixel-frame), not HBM bound.
#include "common.cuh"
#include "heat_core.h"

// ---------------------------------------------------------------------------------------------------- collapse head
// Per frame: collapse the band-passed levels last..first (img = pyrUp(img) + level, pyramid.py:53-55) into A_first.
// The levels below `first` are all-zero in the band-passed pyramid, so the rest of the collapse is pure pyrUp:
// up_level_kernel takes A_first down to A_2 (UNSCALED: x64 per step, the powers of two are folded into the final
// scale), the image the two tile passes start from -- 160x120 at 640x480: 153 KB per frame, written once, read twice.
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
    for (int l = p.last - 1; l >= p.first; --l) {
      const double* s = a + p.off[l + 1];
      double* d = a + p.off[l];
      const int sw = p.w[l + 1], sh = p.h[l + 1], dw = p.w[l], dh = p.h[l];
      for (int i = threadIdx.x; i < dw * dh; i += blockDim.x) {
        const int y = i / dw, x = i - y * dw;
        d[i] = up_at(s, sw, sh, x, y) * (1.0 / 64.0) + d[i];
      }
      __syncthreads();
    }
    const int n0 = p.w[p.first] * p.h[p.first];
    double* dst = p.a_out + f * n0;
    for (int i = threadIdx.x; i < n0; i += blockDim.x) dst[i] = a[p.off[p.first] + i];
    __syncthreads();
  }
}

// One unscaled pyrUp step on a stack of images: (n_img, sh, sw) -> (n_img, dh, dw).  One thread per source pixel (x, y):
// it loads the 3x3 source neighbourhood once and produces the 2x2 outputs (2x..2x+1, 2y..2y+1) -- the horizontal taps
// of the three rows first, then the vertical ones, in the operation order of up_at (the kernel is a stream: 1 read,
// 4 writes per thread).
__global__ void __launch_bounds__(256) up_level_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                                       long long n_img, int sw, int sh, int dw, int dh) {
  const long long per_img = (long long)sw * sh;
  const long long total = n_img * per_img;
  const bool small = total < (1ll << 31);      // 32-bit index arithmetic (a 64-bit division costs about a hundred instructions)
#pragma unroll 1
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long img = small ? (long long)((unsigned)idx / (unsigned)per_img) : idx / per_img;
    const int r = (int)(idx - img * per_img);
    const int y = r / sw, x = r - y * sw;
    const double* s = src + img * per_img;
    double* d = dst + img * (long long)dw * dh;
    const int xm = reflect101(x - 1, sw), xp = x + 1 < sw ? x + 1 : sw - 1;
    const int ym = reflect101(y - 1, sh), yp = y + 1 < sh ? y + 1 : sh - 1;
    const double* r0 = s + (long long)ym * sw;
    const double* r1 = s + (long long)y * sw;
    const double* r2 = s + (long long)yp * sw;
    const double a0 = r0[xm], a1 = r0[x], a2 = r0[xp];
    const double b0 = r1[xm], b1 = r1[x], b2 = r1[xp];
    const double c0 = r2[xm], c1 = r2[x], c2 = r2[xp];
    // even output column 2x: s[x-1] + 6 s[x] + s[x+1]; odd column 2x+1: 4 (s[x] + s[x+1])      (up3)
    const double ea = fma(6.0, a1, a0 + a2), eb = fma(6.0, b1, b0 + b2), ec = fma(6.0, c1, c0 + c2);
    const double oa = 4.0 * (a1 + a2), ob = 4.0 * (b1 + b2), oc = 4.0 * (c1 + c2);
    const int X = 2 * x, Y = 2 * y;
    // even output row 2y: rows y-1, y, y+1; odd row 2y+1: rows y, y+1
    if (Y < dh) {
      if (X < dw) d[(long long)Y * dw + X] = fma(6.0, eb, ea + ec);
      if (X + 1 < dw) d[(long long)Y * dw + X + 1] = fma(6.0, ob, oa + oc);
    }
    if (Y + 1 < dh) {
      if (X < dw) d[(long long)(Y + 1) * dw + X] = 4.0 * (eb + ec);
      if (X + 1 < dw) d[(long long)(Y + 1) * dw + X + 1] = 4.0 * (ob + oc);
    }
  }
}

// ---------------------------------------------------------------------------------------------------- tile passes
// A thread owns a 4x4 block of level-0 pixels of its tile for all T frames of its clip: it reads the 4x4 level-2 values
// around the block from the tile's staged patch and runs the last two pyrUp steps in registers.
#define HM_TW 64            // output tile: 64 x 32 level-0 pixels, 16 x 8 threads
#define HM_TH 32

struct TileParams {
  const double* a2;        // (n_clips, T, h[2]*w[2]) level-2 images, unscaled by 64^(s-2)
  int n_clips, T;
  int w[3], h[3];          // sizes of levels 0, 1, 2
  int tiles_x, tiles_y;
  double scale;            // 2^(-6 s)
  unsigned long long* minmax_keys;   // (n_clips, 4) order-preserving keys: raw min, raw max, avg min, avg max
  double threshold;        // temporal_threshold (pass 2)
  double* avg_out;         // (n_clips, H, W) pass 2
  const double2* bounds;   // (n_clips, tiles, T) smallest / largest level-`skip` value each tile-frame depends on, or null
  const double* a3;        // lazy level 2: (n_clips, T, h3 * w3) level-3 images (unscaled), or null -> a2 is materialised
  int w3, h3;
  const double* a_top;     // (n_clips, T, h_top * w_top) level `skip` images (tile_bounds_kernel)
  int top_w, top_h, n_up;  // their size; number of pyrUp steps from there to level 0
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

#define HM_STAGES 4
#define HM_PW 20            // pitch (doubles) of a staged level-2 patch: 64/4 + 3 columns, padded
#define HM_PH 11            // 32/4 + 3 rows
#define HM_COPIES 2         // ceil(19 * 11 / 128) cp.async per thread per frame
template <int PASS, bool EDGE>
#define HM_P3 96            // level-3 patch of a tile: at most 12 x 8 values
__device__ __forceinline__ void upsample_pass_body(const TileParams& p, double* red_a, double* red_b, double* stage,
                                                   double* stage3, int* frame_list) {
  const int tile = blockIdx.x;
  const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
  const int clip = blockIdx.y;
  const int tid = threadIdx.x;
  const int bx = tid & 15, by = tid >> 4;
  const int X0 = tx * HM_TW + 4 * bx, Y0 = ty * HM_TH + 4 * by;   // my 4x4 output block
  const bool active = X0 < p.w[0] && Y0 < p.h[0];
  const AxisGeom gx = axis_geom(active ? (X0 >> 2) : 0, p.w[1], p.w[2]);
  const AxisGeom gy = axis_geom(active ? (Y0 >> 2) : 0, p.h[1], p.h[2]);
  int offs[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) offs[r][c] = gy.v[r] * p.w[2] + gx.v[c];
  bool okx[4], oky[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    okx[k] = !EDGE || X0 + k < p.w[0];
    oky[k] = !EDGE || Y0 + k < p.h[0];
  }
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  double vmin = INFINITY, vmax = -INFINITY;
  double top_s = 0.0, repl_s = 0.0;
  // Pass 1 looks only at the frames that can still change the clip's min / max.  pyrUp is a convex combination, so every
  // level-0 value of this tile in frame t lies between the smallest and the largest level-`skip` value of the tile's
  // neighbourhood (tile_bounds_kernel); if that interval -- widened by 1e-12 relative, far above the few ulps the
  // evaluation can add -- is inside the extremes already known when the block starts (the seed kernel's and earlier
  // blocks'), the frame is skipped: nothing is staged or evaluated for it.  The result is bit-identical.
  int n_list = p.T;
  if (PASS == 1 && p.bounds) {
    __shared__ int s_count;
    if (tid == 0) s_count = 0;
    __syncthreads();
    const unsigned long long kmin = p.minmax_keys[clip * 4 + 0], kmax = p.minmax_keys[clip * 4 + 1];
    const bool known = kmin <= kmax;   // the keys are at their initial values otherwise (min key > max key)
    const double gmin = known ? key_f64(kmin) : 0.0, gmax = known ? key_f64(kmax) : 0.0;
    const double2* b = p.bounds + ((long long)clip * gridDim.x + tile) * p.T;
    for (int t = tid; t < p.T; t += blockDim.x) {
      const double2 lh = b[t];
      const double mrg = 1e-12 * fmax(fabs(lh.x), fabs(lh.y));
      const bool skip = known && lh.y + mrg <= gmax && lh.x - mrg >= gmin;
      if (!skip) frame_list[atomicAdd(&s_count, 1)] = t;
    }
    __syncthreads();
    n_list = s_count;
  }
  if (PASS == 2) {
    const double lo = key_f64(p.minmax_keys[clip * 4 + 0]), hi = key_f64(p.minmax_keys[clip * 4 + 1]);
    // transforms.py:185-189 on the scaled values; the comparison runs in the unscaled domain (scale is 2^-k: exact)
    const double top = hi - (hi - lo) * p.threshold;
    top_s = top / p.scale;
    repl_s = lo / p.scale;
    if (p.bounds) {
      // Pass 2 replaces every value >= top by the minimum (transforms.py:190-192).  Where the convexity bound says that
      // ALL level-0 values of this tile in frame t are >= top (with the same rounding margin), the frame contributes the
      // minimum to every pixel whatever the values are: it is neither staged nor evaluated, its additions still happen,
      // in frame order.  That is the fate of everything that does not move: top sits 30 % above the most negative value.
      const double2* b = p.bounds + ((long long)clip * gridDim.x + tile) * p.T;
      // the ordered list by ballot + prefix counts instead of one thread walking the flags (that walk and the wait for it
      // were 12 % of the passes' stall samples, ncu r01n; pass 2 0.374 -> 0.326 ms, r02a).  Same list.
      __shared__ int s_cnt[4];
      int n = 0;
      for (int t0 = 0; t0 < p.T; t0 += 128) {                 // blockDim.x == 128: warp w holds frames t0 + 32 w ..
        const int t = t0 + tid;
        bool f = false;
        if (t < p.T) {
          const double2 lh = b[t];
          const double mrg = 1e-12 * fmax(fabs(lh.x), fabs(lh.y));
          f = !(lh.x - mrg >= top);
        }
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if ((tid & 31) == 0) s_cnt[tid >> 5] = __popc(m);
        __syncthreads();
        int off = n;
        for (int w = 0; w < (tid >> 5); ++w) off += s_cnt[w];
        if (f) frame_list[off + __popc(m & ((1u << (tid & 31)) - 1u))] = t;
        n += s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
        __syncthreads();
      }
      n_list = n;
    }
  }
  const long long n2 = (long long)p.w[2] * p.h[2];
  const double* a_clip = p.a2 + (long long)clip * p.T * n2;

  // The tile's level-2 neighbourhood (at most 19 x 11 values per frame) is staged in shared memory by cp.async, HM_STAGES
  // frames deep, so the loads of the register stage are shared-memory hits and the global latency hides behind the
  // float64 work of the frames in between.  One __syncthreads per frame.
  const int xlo = max(0, tx * (HM_TW / 4) - 1), xhi = min(p.w[2] - 1, tx * (HM_TW / 4) + HM_TW / 4 + 1);
  const int ylo = max(0, ty * (HM_TH / 4) - 1), yhi = min(p.h[2] - 1, ty * (HM_TH / 4) + HM_TH / 4 + 1);
  const int pwid = xhi - xlo + 1, n_patch = pwid * (yhi - ylo + 1);
  int src_off[HM_COPIES], dst_off[HM_COPIES];
#pragma unroll
  for (int k = 0; k < HM_COPIES; ++k) {
    const int e = tid + k * 128;
    const int r = e / pwid, c = e - r * pwid;
    src_off[k] = e < n_patch ? (ylo + r) * p.w[2] + xlo + c : -1;
    dst_off[k] = r * HM_PW + c;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) offs[r][c] = (gy.v[r] - ylo) * HM_PW + (gx.v[c] - xlo);
  const unsigned stage_base = (unsigned)__cvta_generic_to_shared(stage);
  // Lazy level 2 (p.a3 != null): only the tile-frames that are evaluated at all need their level-2 patch, so level 2 is
  // not materialised; the ring stages the level-3 patch (at most 12 x 8 values) and the block expands it into the patch
  // buffer with up_level_kernel's arithmetic -- one more barrier per evaluated frame, one 1.3 GB array less.
  const bool lazy = p.a3 != nullptr;
  const int x3lo = lazy ? max(0, (xlo >> 1) - 1) : 0, x3hi = lazy ? min(p.w3 - 1, (xhi >> 1) + 1) : 0;
  const int y3lo = lazy ? max(0, (ylo >> 1) - 1) : 0, y3hi = lazy ? min(p.h3 - 1, (yhi >> 1) + 1) : 0;
  const int pw3 = x3hi - x3lo + 1, n3 = pw3 * (y3hi - y3lo + 1);
  const long long n3img = (long long)p.w3 * p.h3;
  const int src3 = (lazy && tid < n3) ? (y3lo + tid / pw3) * p.w3 + x3lo + tid % pw3 : -1;
  const unsigned stage3_base = (unsigned)__cvta_generic_to_shared(stage3);
  auto issue = [&](int k) {
    if (k < n_list) {
      const int t = p.bounds ? frame_list[k] : k;
      if (lazy) {
        if (src3 >= 0) {
          const double* l3 = p.a3 + ((long long)clip * p.T + t) * n3img;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(stage3_base + (unsigned)(((k % HM_STAGES) * HM_P3 + tid) * 8)),
                       "l"(l3 + src3) : "memory");
        }
      } else {
        const double* l2 = a_clip + t * n2;
        const unsigned dst = stage_base + (unsigned)((k % HM_STAGES) * HM_PW * HM_PH * 8);
#pragma unroll
        for (int c = 0; c < HM_COPIES; ++c)
          if (src_off[c] >= 0)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst + dst_off[c] * 8), "l"(l2 + src_off[c])
                         : "memory");
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
#pragma unroll
  for (int t = 0; t < HM_STAGES - 1; ++t) issue(t);

  {
    // pass 1 walks the frame list; pass 2 walks every frame in order and consumes the list as it goes (k = position of
    // the next staged frame): a frame that is not on the list only adds the replacement value
    int k = 0;
    const int n_iter = (PASS == 2 && p.bounds) ? p.T : n_list;
    int it0 = 0;
    if (PASS == 2 && p.bounds) {
      // Until the first frame that is evaluated every accumulator of the tile receives the same additions
      // (0 + repl + repl + ...): one chain instead of sixteen.  Most tiles evaluate no frame at all.
      it0 = n_list > 0 ? frame_list[0] : p.T;                // block-uniform
      double pre = 0.0;
#pragma unroll 4
      for (int it = 0; it < it0; ++it) pre += repl_s;
#pragma unroll
      for (int ky = 0; ky < 4; ++ky)
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) acc[ky][kx] = pre;
    }
    for (int it = it0; it < n_iter; ++it) {
      if (PASS == 2 && p.bounds) {
        if (k >= n_list || frame_list[k] != it) {          // block-uniform
          if (active) {
#pragma unroll
            for (int ky = 0; ky < 4; ++ky)
#pragma unroll
              for (int kx = 0; kx < 4; ++kx) acc[ky][kx] += repl_s;
          }
          continue;
        }
      }
      const int t = k++;               // position in the staging ring
      {
      asm volatile("cp.async.wait_group %0;\n" ::"n"(HM_STAGES - 2) : "memory");
      __syncthreads();                 // frame t has landed for everyone; everyone is done with frame t-1's stage
      issue(t + HM_STAGES - 1);        // refills the stage frame t-1 used
      if (lazy) {                      // expand the level-3 patch of this frame into the level-2 patch buffer (slot 0)
        const double* s3 = stage3 + (t % HM_STAGES) * HM_P3;
#pragma unroll
        for (int c = 0; c < HM_COPIES; ++c) {
          const int e = tid + c * 128;
          if (e < n_patch) {
            const int r = e / pwid, cc = e - r * pwid;
            stage[r * HM_PW + cc] = a2_value(s3, pw3, x3lo, y3lo, p.w3, p.h3, xlo + cc, ylo + r);
          }
        }
        __syncthreads();               // the patch is complete (the barrier above keeps the previous frame's readers out)
      }
      }
      if (!active) continue;
      const double* l2 = lazy ? stage : stage + (t % HM_STAGES) * HM_PW * HM_PH;
      double v[4][4], o[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) v[r][c] = l2[offs[r][c]];
      block4x4<EDGE>(v, gx, gy, o);
      if (PASS == 1) {
        // Candidate filter on the high words: a new maximum above a positive running maximum is a positive double whose
        // high word is >= the running one as a signed int; a new minimum below a negative running minimum is a negative
        // double whose high word is >= the running one as an unsigned int.  The exact float64 update runs only then.
        int mh = INT_MIN;
        unsigned nh = 0u;
#pragma unroll
        for (int ky = 0; ky < 4; ++ky)
#pragma unroll
          for (int kx = 0; kx < 4; ++kx) {
            if (okx[kx] && oky[ky]) {
              const int hw = __double2hiint(o[ky][kx]);
              mh = max(mh, hw);
              nh = max(nh, (unsigned)hw);
            }
          }
        const int thr_max = vmax > 0.0 ? __double2hiint(vmax) : INT_MIN;
        const unsigned thr_min = vmin < 0.0 ? (unsigned)__double2hiint(vmin) : 0u;
        if (mh >= thr_max || nh >= thr_min) {
#pragma unroll
          for (int ky = 0; ky < 4; ++ky)
#pragma unroll
            for (int kx = 0; kx < 4; ++kx)
              if (okx[kx] && oky[ky]) {
                vmin = fmin(vmin, o[ky][kx]);
                vmax = fmax(vmax, o[ky][kx]);
              }
        }
      } else {
#pragma unroll
        for (int ky = 0; ky < 4; ++ky)
#pragma unroll
          for (int kx = 0; kx < 4; ++kx) acc[ky][kx] += (o[ky][kx] >= top_s) ? repl_s : o[ky][kx];
      }
    }
  }

  // ---- epilogue ---------------------------------------------------------------------------------------------------------
  if (PASS == 2) {
    vmin = INFINITY;
    vmax = -INFINITY;
    if (active) {
      double* dst = p.avg_out + (long long)clip * p.w[0] * p.h[0];
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        double row[4];
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
          row[kx] = (acc[ky][kx] * p.scale) / (double)p.T;   // np.average over T (base.py:562)
          if (okx[kx] && oky[ky]) {
            vmin = fmin(vmin, row[kx]);
            vmax = fmax(vmax, row[kx]);
          }
        }
        if (!EDGE) {
          double2* d2 = reinterpret_cast<double2*>(dst + (long long)(Y0 + ky) * p.w[0] + X0);
          d2[0] = make_double2(row[0], row[1]);
          d2[1] = make_double2(row[2], row[3]);
        } else {
#pragma unroll
          for (int kx = 0; kx < 4; ++kx)
            if (okx[kx] && oky[ky]) dst[(long long)(Y0 + ky) * p.w[0] + X0 + kx] = row[kx];
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
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { vmin = fmin(vmin, red_a[w]); vmax = fmax(vmax, red_b[w]); }
    const int base = clip * 4 + (PASS == 1 ? 0 : 2);
    if (vmin <= vmax) {
      atomicMin(&p.minmax_keys[base + 0], f64_key(vmin));
      atomicMax(&p.minmax_keys[base + 1], f64_key(vmax));
    }
  }
}

// (lo, hi) of the level-`skip` values that the level-0 pixels of tile `tile` depend on in frame t.  The footprint is the
// tile's pixel range pushed up one level at a time: x at level l needs floor(x/2) - 1 .. floor(x/2) + 1 at level l+1 (the
// border rules -- reflect-101 below, clamp above -- only ever pick indices inside that range once it is clamped to the
// image).  One thread per (frame, tile, clip).
__global__ void __launch_bounds__(256) tile_bounds_kernel(const TileParams p, double2* __restrict__ out) {
  extern __shared__ double tb_img[];   // the level-`skip` image of this (frame, clip)
  const int tiles = p.tiles_x * p.tiles_y;
  const int clip = blockIdx.y, t = blockIdx.x;
  const int n_top = p.top_w * p.top_h;
  const double* a = p.a_top + ((long long)clip * p.T + t) * n_top;
  for (int i = threadIdx.x; i < n_top; i += blockDim.x) tb_img[i] = a[i];
  __syncthreads();
  for (int tile = threadIdx.x; tile < tiles; tile += blockDim.x) {
    const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
    int x0 = tx * HM_TW, x1 = min(p.w[0], x0 + HM_TW) - 1, y0 = ty * HM_TH, y1 = min(p.h[0], y0 + HM_TH) - 1;
    for (int l = 0; l < p.n_up; ++l) {
      x0 = (x0 >> 1) - 1; x1 = (x1 >> 1) + 1;
      y0 = (y0 >> 1) - 1; y1 = (y1 >> 1) + 1;
    }
    x0 = max(x0, 0); y0 = max(y0, 0);
    x1 = min(x1, p.top_w - 1); y1 = min(y1, p.top_h - 1);
    double lo = INFINITY, hi = -INFINITY;
    for (int y = y0; y <= y1; ++y)
      for (int x = x0; x <= x1; ++x) {
        const double v = tb_img[y * p.top_w + x];
        lo = fmin(lo, v);
        hi = fmax(hi, v);
      }
    out[((long long)clip * tiles + tile) * p.T + t] = make_double2(lo, hi);
  }
}

// Seeds the clip's min/max keys before pass 1: per (clip, sampled frame) the positions of the level-2 maximum and minimum
// are located and the 4x4 level-0 blocks above them are evaluated exactly as pass 1 would.  Any value so obtained is a
// value pass 1 would also have produced, so the keys are valid lower bounds of the extremes; pass 1 then skips every
// (tile, frame) that cannot beat them.  One block per (sampled frame, clip).
__global__ void __launch_bounds__(256) minmax_seed_kernel(const TileParams p, int frame_stride) {
  const int clip = blockIdx.y, t = blockIdx.x * frame_stride, tid = threadIdx.x;
  if (t >= p.T) return;
  // the extremes are located on level 2 when it is materialised, else on level 3 (lazy level 2): any block evaluated
  // exactly seeds valid bounds, the choice only affects how tight they are
  const bool lazy = p.a3 != nullptr;
  const int sw = lazy ? p.w3 : p.w[2], sh = lazy ? p.h3 : p.h[2];
  const int n_src = sw * sh;
  const double* src = (lazy ? p.a3 : p.a2) + ((long long)clip * p.T + t) * n_src;
  double vmx = -INFINITY, vmn = INFINITY;
  int imx = 0, imn = 0;
  for (int i = tid; i < n_src; i += 256) {
    const double v = src[i];
    if (v > vmx) { vmx = v; imx = i; }
    if (v < vmn) { vmn = v; imn = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, vmx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, imx, o);
    if (ov > vmx) { vmx = ov; imx = oi; }
    const double uv = __shfl_xor_sync(0xffffffffu, vmn, o);
    const int ui = __shfl_xor_sync(0xffffffffu, imn, o);
    if (uv < vmn) { vmn = uv; imn = ui; }
  }
  __shared__ double s_mx[8], s_mn[8];
  __shared__ int s_imx[8], s_imn[8];
  if ((tid & 31) == 0) { s_mx[tid >> 5] = vmx; s_imx[tid >> 5] = imx; s_mn[tid >> 5] = vmn; s_imn[tid >> 5] = imn; }
  __syncthreads();
  if (tid < 2) {   // thread 0: the block above the maximum, thread 1: above the minimum
    double best = tid == 0 ? -INFINITY : INFINITY;
    int idx = 0;
    for (int w = 0; w < 8; ++w) {
      const double v = tid == 0 ? s_mx[w] : s_mn[w];
      if (tid == 0 ? v > best : v < best) { best = v; idx = tid == 0 ? s_imx[w] : s_imn[w]; }
    }
    int i2x = idx % sw, i2y = idx / sw;
    if (lazy) {    // level-3 position -> the level-2 sample on top of it
      i2x = min(2 * i2x, p.w[2] - 1);
      i2y = min(2 * i2y, p.h[2] - 1);
    }
    const int X0 = 4 * i2x, Y0 = 4 * i2y;
    const AxisGeom gx = axis_geom(i2x, p.w[1], p.w[2]), gy = axis_geom(i2y, p.h[1], p.h[2]);
    double v[4][4], o[4][4];
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c)
        v[r][c] = lazy ? a2_value(src, p.w3, 0, 0, p.w3, p.h3, gx.v[c], gy.v[r]) : src[gy.v[r] * p.w[2] + gx.v[c]];
    block4x4<true>(v, gx, gy, o);
    double vmin = INFINITY, vmax = -INFINITY;
    for (int ky = 0; ky < 4; ++ky)
      for (int kx = 0; kx < 4; ++kx)
        if (X0 + kx < p.w[0] && Y0 + ky < p.h[0]) { vmin = fmin(vmin, o[ky][kx]); vmax = fmax(vmax, o[ky][kx]); }
    if (vmin <= vmax) {
      atomicMin(&p.minmax_keys[clip * 4 + 0], f64_key(vmin * p.scale));
      atomicMax(&p.minmax_keys[clip * 4 + 1], f64_key(vmax * p.scale));
    }
  }
}

// 128 registers (no spills) for a fourth resident block: pass 2 0.410 -> 0.363 ms per 64-clip step (r02a)
#define HM_LAUNCH_BOUNDS __launch_bounds__(128, 4)
template <int PASS>
__global__ void HM_LAUNCH_BOUNDS upsample_pass_kernel(const TileParams p) {
  __shared__ double red_a[4], red_b[4];
  __shared__ __align__(16) double stage[HM_STAGES * HM_PW * HM_PH];
  const int tx = blockIdx.x % p.tiles_x, ty = blockIdx.x / p.tiles_x;
  // tiles that touch the right / bottom image border (or an unaligned row) take the guarded variant
  const bool edge = (tx + 1) * HM_TW >= p.w[0] || (ty + 1) * HM_TH >= p.h[0] || (p.w[0] & 1);
  __shared__ __align__(16) double stage3[HM_STAGES * HM_P3];
  extern __shared__ int frame_list_smem[];   // pass 1: T ints
  if (edge) upsample_pass_body<PASS, true>(p, red_a, red_b, stage, stage3, frame_list_smem);
  else upsample_pass_body<PASS, false>(p, red_a, red_b, stage, stage3, frame_list_smem);
}

__global__ void minmax_init_kernel(unsigned long long* keys, int n_clips) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_clips * 4) keys[i] = (i & 1) ? 0ull : ~0ull;   // min slots start at +max key, max slots at 0
}

// base.py:563-564: (avg - min) / (max - min) * 255, truncated to uint8 (NaN from a flat map becomes 0)
__device__ __forceinline__ uint8_t heat_u8(double a, double mn, double range) {
  const double v = ((a - mn) / range) * 255.0;       // base.py:563-564: min-max normalise, *255, truncate
  return (v == v) ? (uint8_t)(int)v : (uint8_t)0;
}
// VEC: hw is a multiple of 4, so every clip plane starts 32-byte aligned in `avg` and 4-byte aligned in `heat`: a thread
// converts four pixels per iteration (two 16-byte loads in flight, one 4-byte store).
template <bool VEC>
__global__ void __launch_bounds__(256) heat_normalise_kernel(const double* __restrict__ avg,
                                                             const unsigned long long* __restrict__ keys,
                                                             uint8_t* __restrict__ heat, double* __restrict__ minmax_out,
                                                             int n_clips, long long hw) {
  const int clip = blockIdx.y;
  const double mn = key_f64(keys[clip * 4 + 2]), mx = key_f64(keys[clip * 4 + 3]);
  const double range = mx - mn;
  const double* a = avg + clip * hw;
  uint8_t* o = heat + clip * hw;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (VEC) {
    const double2* a2 = reinterpret_cast<const double2*>(a);
    uchar4* o4 = reinterpret_cast<uchar4*>(o);
    // (not unrolled: a thread has one trip, and the trip count an unrolled grid-stride loop needs is a 64-bit division
    // -- 370 of the 500 instructions a thread executed, ncu r04b)
#pragma unroll 1
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < hw / 4; q += stride) {
      const double2 lo = a2[2 * q], hi = a2[2 * q + 1];
      o4[q] = make_uchar4(heat_u8(lo.x, mn, range), heat_u8(lo.y, mn, range), heat_u8(hi.x, mn, range),
                          heat_u8(hi.y, mn, range));
    }
  } else {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw; i += stride) o[i] = heat_u8(a[i], mn, range);
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
// Pruning (and with it the lazy level 2) needs the frame list to fit in shared memory and level `s` in one block.
static bool heatmap_prunes(const rm_handle* h, const LevelGeom& g, int T, int s) {
  return !h->no_minmax_seed && s < g.n_levels && (size_t)T * 5 + 32 <= 32768 && g.w[s] * g.h[s] * 8 <= h->smem_optin;
}
// Lowest level that is materialised: 3 when the passes expand level 2 themselves, else 2.
static int heatmap_lowest_level(const rm_handle* h, const LevelGeom& g, int T, int s) {
  return (heatmap_prunes(h, g, T, s) && s >= 3) ? 3 : 2;
}

extern "C" int32_t rm_heatmap_workspace_bytes(rm_handle* h, int32_t W, int32_t H, int32_t n_clips, int32_t T, size_t* out) {
  RM_CHECK_ARG(h, h && out && W >= 1 && H >= 1 && n_clips >= 0 && T >= 1, "null pointer or bad size");
  LevelGeom g = make_geom(W, H, h->p.pyramid_levels);
  const int s = h->p.skip_levels_at_top;
  size_t a = 0;                                                // A_s .. A_lowest (depends on the "no_minmax_seed" option)
  for (int l = heatmap_lowest_level(h, g, T, s); l <= s && l < g.n_levels; ++l)
    a += (((size_t)n_clips * T * g.w[l] * g.h[l] * 8) + 255) & ~(size_t)255;
  size_t avg = (size_t)n_clips * W * H * 8;                    // time average
  size_t keys = ((size_t)n_clips * 4 * 8 + 255) & ~(size_t)255;
  const size_t tiles = (size_t)((W + HM_TW - 1) / HM_TW) * ((H + HM_TH - 1) / HM_TH);
  size_t bounds = (size_t)n_clips * tiles * T * sizeof(double2);      // pass-1 pruning bounds
  *out = a + avg + keys + bounds + 5 * 256;
  return RM_OK;
}

extern "C" int32_t rm_heatmap(rm_handle* h, const double* bp, int32_t n_clips, int32_t T, int32_t W, int32_t H,
                              uint8_t* heat_out, double* minmax_out, void* workspace, size_t workspace_bytes, void* stream) {
  RM_CHECK_ARG(h, h && bp && heat_out && n_clips >= 0 && T >= 1 && W >= 1 && H >= 1, "null pointer or bad size");
  const int L = h->p.pyramid_levels, s = h->p.skip_levels_at_top;
  if (s < 2 || L - 1 <= s || L > RM_MAX_LEVELS)
    return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: fused heat map needs skip >= 2 and skip < levels-1", __func__);
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
  double* a_lvl[RM_MAX_LEVELS] = {nullptr};
  const bool prune = heatmap_prunes(h, g, T, s);
  const int lowest = heatmap_lowest_level(h, g, T, s);
  const bool lazy_l2 = lowest == 3;     // level 2 is expanded per evaluated tile-frame from level 3
  for (int l = s; l >= lowest; --l) {
    a_lvl[l] = reinterpret_cast<double*>(base);
    base += ((size_t)n_clips * T * g.w[l] * g.h[l] * 8 + 255) & ~(size_t)255;
  }
  double* avg = reinterpret_cast<double*>(base);
  base += ((size_t)n_clips * W * H * 8 + 255) & ~(size_t)255;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(base);
  base += ((size_t)n_clips * 4 * 8 + 255) & ~(size_t)255;
  double2* bounds = reinterpret_cast<double2*>(base);
  const long long n_frames = (long long)n_clips * T;

  HeadParams hp;
  memset(&hp, 0, sizeof(hp));
  hp.bp = bp;
  hp.a_out = a_lvl[s];
  hp.n_frames = n_frames;
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
  long long hgrid = n_frames < (long long)h->sm_count * 8 ? n_frames : (long long)h->sm_count * 8;
  RM_PROF(h, st, "collapse_head_kernel");
  collapse_head_kernel<<<(unsigned)hgrid, 256, head_smem, st>>>(hp);
  RM_LAUNCH_CHECK(h);
  for (int l = s - 1; l >= lowest; --l) {   // A_{l+1} -> A_l, unscaled
    const long long total = n_frames * g.w[l + 1] * g.h[l + 1];
    const long long blocks = (total + 255) / 256;
    const long long cap = (long long)h->sm_count * 32;
    RM_PROF(h, st, "up_level_kernel");
    up_level_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(a_lvl[l + 1], a_lvl[l], n_frames, g.w[l + 1],
                                                                           g.h[l + 1], g.w[l], g.h[l]);
    RM_LAUNCH_CHECK(h);
  }

  RM_PROF(h, st, "minmax_init_kernel");
  minmax_init_kernel<<<div_up(n_clips * 4, 128), 128, 0, st>>>(keys, n_clips);
  RM_LAUNCH_CHECK(h);

  TileParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.a2 = a_lvl[2];
  tp.n_clips = n_clips;
  tp.T = T;
  for (int l = 0; l <= 2; ++l) {
    tp.w[l] = g.w[l];
    tp.h[l] = g.h[l];
  }
  tp.tiles_x = (W + HM_TW - 1) / HM_TW;
  tp.tiles_y = (H + HM_TH - 1) / HM_TH;
  tp.scale = 1.0;
  for (int l = 0; l < s; ++l) tp.scale *= 1.0 / 64.0;
  tp.minmax_keys = keys;
  tp.threshold = h->p.temporal_threshold;
  tp.avg_out = avg;
  dim3 grid(tp.tiles_x * tp.tiles_y, n_clips);
  tp.a_top = a_lvl[s];
  tp.top_w = g.w[s];
  tp.top_h = g.h[s];
  tp.n_up = s;
  if (lazy_l2) {
    tp.a3 = a_lvl[3];
    tp.w3 = g.w[3];
    tp.h3 = g.h[3];
  }
  if (prune) {
    const int tb_smem = g.w[s] * g.h[s] * 8;
    RM_CUDA(h, cudaFuncSetAttribute(tile_bounds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tb_smem));
    RM_PROF(h, st, "tile_bounds_kernel");
    tile_bounds_kernel<<<dim3(T, n_clips), 256, tb_smem, st>>>(tp, bounds);
    RM_LAUNCH_CHECK(h);
    tp.bounds = bounds;
    const int stride = T >= 8 ? 2 : 1;
    RM_PROF(h, st, "minmax_seed_kernel");
    minmax_seed_kernel<<<dim3((T + stride - 1) / stride, n_clips), 256, 0, st>>>(tp, stride);
    RM_LAUNCH_CHECK(h);
  }
  RM_PROF(h, st, "upsample_pass_kernel<1>");
  const size_t list_smem = tp.bounds ? (size_t)T * 4 + (size_t)((T + 3) / 4) * 4 + 16 : 0;   // frame list, flags, count
  upsample_pass_kernel<1><<<grid, 128, list_smem, st>>>(tp);
  RM_LAUNCH_CHECK(h);
  RM_PROF(h, st, "upsample_pass_kernel<2>");
  upsample_pass_kernel<2><<<grid, 128, list_smem, st>>>(tp);
  RM_LAUNCH_CHECK(h);
  long long hw = (long long)W * H;
  const bool vec = hw % 4 == 0 && ((uintptr_t)heat_out & 3) == 0;
  const long long items = vec ? hw / 4 : hw;
  dim3 ngrid((unsigned)((items + 255) / 256 < 1024 ? (items + 255) / 256 : 1024), n_clips);
  RM_PROF(h, st, "heat_normalise_kernel");
  if (vec)
    heat_normalise_kernel<true><<<ngrid, 256, 0, st>>>(avg, keys, heat_out, minmax_out, n_clips, hw);
  else
    heat_normalise_kernel<false><<<ngrid, 256, 0, st>>>(avg, keys, heat_out, minmax_out, n_clips, hw);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}
