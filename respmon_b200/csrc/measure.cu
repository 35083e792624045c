// Motion measurement over a whole clip: extract_motion(), 'flow' and 'average' branches (base.py:354-407).
//
//   float_to_uint8(crop)              transforms.py:26-29 after uint8_to_float -> a 256-entry LUT on the uint8 frame
//   cv2.goodFeaturesToTrack           base.py:365-366  -> gftt_cov / gftt_eig / gftt_select kernels
//   cv2.calcOpticalFlowPyrLK          base.py:371-372  -> lk_track_smem_kernel: the clip's LUT-mapped ROI crops and their
//                                                         uint8 pyramids live in shared memory (prefetched a frame
//                                                         ahead), one warp per corner, the corners of a clip split over
//                                                         blocks of at most LK_PPB, frames walked in resumable chunks;
//                                                         lk_pyr_kernel + lk_track_kernel (global-memory pyramids) for
//                                                         ROIs too large to stage
//   mean(old - new), PCA projection   base.py:388-405  -> motion_reduce_kernel (ordered float32 mean over the surviving
//                                                         points) + motion_pca_kernel
//   rm_measure_signal                                  -> the tracker chunks on the caller's stream, PCA / filtfilt /
//                                                         peaks / LM gate / BPM of finished chunks on the handle's own
//                                                         streams (signal.cu)
//
// Third-party semantics follow SURVEY.md App. A.5 / A.6 and oracle/np_kernels.py (validated against cv2 there).
// The per-frame working set is a few KB per clip: these kernels are latency bound; throughput comes from the batch.
// Compile with -fmad=false: float expressions mirror OpenCV's non-fused arithmetic.
#include "common.cuh"
#include "signal_core.h"

#define LK_MAX_PTS 128
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
#define LK_WARPS 8
#define LK_MAX_LEVELS 4

__device__ __forceinline__ int reflect101_multi(int p, int n) {   // cv::borderInterpolate(BORDER_REFLECT_101)
  if (n == 1) return 0;
  while ((unsigned)p >= (unsigned)n) p = (p < 0) ? -p : 2 * n - 2 - p;
  return p;
}
__device__ __forceinline__ unsigned f32_key(float v) {
  unsigned b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_f32(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct MeasureParams {
  const uint8_t* frames;   // (n_clips, T, H, W)
  const int32_t* roi;      // (n_clips, 4) x,y,w,h
  const uint8_t* lut;
  int n_clips, T, W, H, first_frame, n_frames;
  int maxw, maxh;          // workspace was sized for ROIs up to maxw x maxh
  // hyper-parameters
  int max_corners, min_distance, block_size;
  float quality;
  int win, max_level, max_iter;
  double eps2;             // lk_eps^2 (double, as cv::TermCriteria keeps it)
  float min_eig;
  int buf_len;             // measure_buffer_len
  // workspace
  float* cov;              // (n_clips, 3, maxw*maxh)
  float* eig;              // (n_clips, maxw*maxh)
  unsigned* eigmax;        // (n_clips) key of the maximum
  unsigned long long* cand;// (n_clips, maxw*maxh)
  uint8_t* pyr[LK_MAX_LEVELS];      // levels 1..: (n_clips, n_frames, lvl_elems[l])
  long long lvl_elems[LK_MAX_LEVELS];
  // outputs
  float* motion;           // (n_clips, n_frames, 2)
  double* data;            // (n_clips, n_frames)
  int32_t* npts;           // (n_clips)
  int32_t* status;         // (n_clips)
  float* pts_dbg;          // (n_clips, n_frames, LK_MAX_PTS, 2) or null: per-frame tracked points (tests)
  // frame chunk [f0, f1) of this launch and the tracker state carried between chunks (shared-memory tracker only)
  int f0, f1;
  float* st_pts;           // (n_clips, LK_MAX_PTS, 2)   points of block b start at slot b * ppb
  int32_t* st_idx;         // (n_clips, LK_MAX_PTS)      their indices in the first frame's corner list
  int32_t* st_n;           // (n_clips, LK_MAX_BLOCKS)   points block b still tracks
  // shared-memory tracker: the corners of a clip are split over `bpc` blocks of at most `ppb` points each (one warp per
  // point, so a clip with many corners is not the batch's critical path); every block writes old-new of its surviving
  // points per frame, indexed by the point's original index, and motion_reduce_kernel folds them in that order
  int bpc, ppb;
  float2* delta;           // (n_clips, delta_frames, LK_MAX_PTS) for frames delta_f0 .., NaN = the point is gone
  int delta_f0, delta_frames;
  int ring;                // > 0: `frames` is a ring of that many frames per clip, frame f lives in slot f % ring
};
#define LK_MAX_BLOCKS 16
#ifndef LK_PPB
#define LK_PPB 16
#endif

__device__ __forceinline__ bool roi_ok_geom(const MeasureParams& p, int clip, int& x, int& y, int& w, int& h) {
  x = p.roi[clip * 4 + 0]; y = p.roi[clip * 4 + 1]; w = p.roi[clip * 4 + 2]; h = p.roi[clip * 4 + 3];
  return w >= 1 && h >= 1 && x >= 0 && y >= 0 && x + w <= p.W && y + h <= p.H && w <= p.maxw && h <= p.maxh;
}
__device__ __forceinline__ bool roi_ok(const MeasureParams& p, int clip, int& x, int& y, int& w, int& h) {
  x = p.roi[clip * 4 + 0]; y = p.roi[clip * 4 + 1]; w = p.roi[clip * 4 + 2]; h = p.roi[clip * 4 + 3];
  return p.status[clip] == RM_CLIP_OK && w >= 1 && h >= 1 && x >= 0 && y >= 0 && x + w <= p.W && y + h <= p.H &&
         w <= p.maxw && h <= p.maxh;
}

// ------------------------------------------------------------------------------------------------ Shi-Tomasi
// Sobel-3 derivatives scaled by 1/(4*block*255), products (cornerEigenValsVecs).  One thread per ROI pixel.
__global__ void gftt_cov_kernel(const MeasureParams p) {
  const int clip = blockIdx.y;
  int rx, ry, rw, rh;
  if (!roi_ok(p, clip, rx, ry, rw, rh)) return;
  const uint8_t* img = p.frames + ((long long)clip * p.T + p.first_frame) * p.W * p.H + (long long)ry * p.W + rx;
  const float k1 = (float)(1.0 / (4.0 * p.block_size * 255.0));
  const float k0 = 2.0f * k1;
  const long long plane = (long long)p.maxw * p.maxh;
  float* cov = p.cov + (long long)clip * 3 * plane;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rw * rh; i += gridDim.x * blockDim.x) {
    const int x = i % rw, y = i / rw;
    const int xm = reflect101_multi(x - 1, rw), xp = reflect101_multi(x + 1, rw);
    const int ym = reflect101_multi(y - 1, rh), yp = reflect101_multi(y + 1, rh);
    float s[3][3];
    const int ys[3] = {ym, y, yp}, xs[3] = {xm, x, xp};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) s[a][b] = (float)p.lut[img[(long long)ys[a] * p.W + xs[b]]];
    // Dx: row pass [-1 0 1] (exact), column pass [1 2 1]*scale = fma(up+down, k1, mid*k0)
    const float r0 = s[0][2] - s[0][0], r1 = s[1][2] - s[1][0], r2 = s[2][2] - s[2][0];
    const float dx = fmaf(r0 + r2, k1, r1 * k0);
    // Dy: row pass [1 2 1]*scale = fma(k1, right, fma(k0, mid, k1*left)), column pass [-1 0 1]
    const float t0 = fmaf(k1, s[0][2], fmaf(k0, s[0][1], k1 * s[0][0]));
    const float t2 = fmaf(k1, s[2][2], fmaf(k0, s[2][1], k1 * s[2][0]));
    const float dy = t2 - t0;
    cov[i] = dx * dx;
    cov[plane + i] = dx * dy;
    cov[2 * plane + i] = dy * dy;
  }
}

// Un-normalised block x block box sums (float64 accumulation like cv::boxFilter), min eigenvalue, running maximum.
__global__ void gftt_eig_kernel(const MeasureParams p) {
  const int clip = blockIdx.y;
  int rx, ry, rw, rh;
  if (!roi_ok(p, clip, rx, ry, rw, rh)) return;
  const long long plane = (long long)p.maxw * p.maxh;
  const float* cov = p.cov + (long long)clip * 3 * plane;
  float* eig = p.eig + (long long)clip * plane;
  const int r = p.block_size / 2;
  unsigned best = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rw * rh; i += gridDim.x * blockDim.x) {
    const int x = i % rw, y = i / rw;
    double sxx = 0.0, sxy = 0.0, syy = 0.0;
    for (int dy = -r; dy < p.block_size - r; ++dy) {
      const int yy = reflect101_multi(y + dy, rh);
      double rxx = 0.0, rxy = 0.0, ryy = 0.0;
      for (int dx = -r; dx < p.block_size - r; ++dx) {
        const int j = yy * rw + reflect101_multi(x + dx, rw);
        rxx += (double)cov[j];
        rxy += (double)cov[plane + j];
        ryy += (double)cov[2 * plane + j];
      }
      sxx += rxx; sxy += rxy; syy += ryy;
    }
    const float a = (float)sxx * 0.5f, b = (float)sxy, c = (float)syy * 0.5f;
    const float d = a - c;
    const float v = (a + c) - sqrtf(d * d + b * b);
    eig[i] = v;
    const unsigned k = f32_key(v);
    best = k > best ? k : best;
  }
  for (int o = 16; o > 0; o >>= 1) {
    unsigned other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if ((threadIdx.x & 31) == 0 && best) atomicMax(&p.eigmax[clip], best);
}

// threshold at quality*max, 3x3 local maxima on interior pixels, then greedy selection in descending order of
// (value, address) with the minimum-distance rule.  One block per clip.
__global__ void __launch_bounds__(256) gftt_select_kernel(const MeasureParams p, float* pts_out) {
  const int clip = blockIdx.x;
  int rx, ry, rw, rh;
  __shared__ int s_ncand, s_naccept;
  __shared__ unsigned long long s_red[8];
  __shared__ unsigned long long s_best;
  if (threadIdx.x == 0) { s_ncand = 0; s_naccept = 0; }
  __syncthreads();
  if (!roi_ok(p, clip, rx, ry, rw, rh)) {
    if (threadIdx.x == 0) p.npts[clip] = 0;
    return;
  }
  const long long plane = (long long)p.maxw * p.maxh;
  const float* eig = p.eig + (long long)clip * plane;
  unsigned long long* cand = p.cand + (long long)clip * plane;
  const float maxv = key_f32(p.eigmax[clip]);
  const float thr = (float)((double)maxv * (double)p.quality);   // cv::threshold receives a double, compares floats
  for (int i = threadIdx.x; i < rw * rh; i += blockDim.x) {
    const int x = i % rw, y = i / rw;
    if (x < 1 || y < 1 || x >= rw - 1 || y >= rh - 1) continue;
    const float v = eig[i];
    if (!(v > thr) || v == 0.0f) continue;
    bool is_max = true;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) is_max &= (v >= eig[(y + dy) * rw + x + dx]);
    if (is_max) {
      const int slot = atomicAdd(&s_ncand, 1);
      cand[slot] = ((unsigned long long)f32_key(v) << 32) | (unsigned)i;
    }
  }
  __syncthreads();
  const int ncand = s_ncand;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int md2 = p.min_distance * p.min_distance;
  float* out = pts_out + (long long)clip * LK_MAX_PTS * 2;
  const int limit = p.max_corners < LK_MAX_PTS ? p.max_corners : LK_MAX_PTS;
  for (;;) {
    unsigned long long best = 0ull;
    for (int i = threadIdx.x; i < ncand; i += blockDim.x) best = cand[i] > best ? cand[i] : best;
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) s_red[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) best = s_red[w] > best ? s_red[w] : best;
      s_best = best;
      if (best) {
        const int addr = (int)(best & 0xffffffffu);
        out[2 * s_naccept] = (float)(addr % rw);
        out[2 * s_naccept + 1] = (float)(addr / rw);
        s_naccept++;
      }
    }
    __syncthreads();
    const unsigned long long b = s_best;
    if (!b || s_naccept >= limit) break;
    const int addr = (int)(b & 0xffffffffu);
    const int bx = addr % rw, by = addr / rw;
    for (int i = threadIdx.x; i < ncand; i += blockDim.x) {
      const unsigned long long c = cand[i];
      if (!c) continue;
      const int a2 = (int)(c & 0xffffffffu);
      const int dx = a2 % rw - bx, dy = a2 / rw - by;
      if (c == b || (p.min_distance >= 1 && dx * dx + dy * dy < md2)) cand[i] = 0ull;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    p.npts[clip] = s_naccept;
    if (s_naccept == 0) p.status[clip] = RM_CLIP_NO_CORNERS;
  }
}

// ------------------------------------------------------------------------------------------------ LK pyramids
__device__ __forceinline__ int lk_num_levels(int w, int h, int win, int max_level, int* lw, int* lh) {
  int n = 1;
  lw[0] = w; lh[0] = h;
  while (n <= max_level && n < LK_MAX_LEVELS) {
    const int nw = (lw[n - 1] + 1) / 2, nh = (lh[n - 1] + 1) / 2;
    if (nw <= win || nh <= win) break;
    lw[n] = nw; lh[n] = nh;
    ++n;
  }
  return n;
}

// cv2.pyrDown on uint8 (exact integers, (sum + 128) >> 8, REFLECT_101): level `lvl` of every measure frame.
__global__ void lk_pyr_kernel(const MeasureParams p, int lvl) {
  const int clip = blockIdx.z, f = blockIdx.y;
  int rx, ry, rw, rh;
  if (!roi_ok(p, clip, rx, ry, rw, rh)) return;
  int lw[LK_MAX_LEVELS], lh[LK_MAX_LEVELS];
  const int nlev = lk_num_levels(rw, rh, p.win, p.max_level, lw, lh);
  if (lvl >= nlev) return;
  const int sw = lw[lvl - 1], sh = lh[lvl - 1], dw = lw[lvl], dh = lh[lvl];
  const uint8_t* src;
  int spitch;
  const bool use_lut = (lvl == 1);
  if (lvl == 1) {
    src = p.frames + ((long long)clip * p.T + p.first_frame + f) * p.W * p.H + (long long)ry * p.W + rx;
    spitch = p.W;
  } else {
    src = p.pyr[lvl - 1] + ((long long)clip * p.n_frames + f) * p.lvl_elems[lvl - 1];
    spitch = sw;
  }
  uint8_t* dst = p.pyr[lvl] + ((long long)clip * p.n_frames + f) * p.lvl_elems[lvl];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dw * dh; i += gridDim.x * blockDim.x) {
    const int x = i % dw, y = i / dw;
    int xs[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) xs[k] = reflect101_multi(2 * x + k - 2, sw);
    int r[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const uint8_t* row = src + (long long)reflect101_multi(2 * y + k - 2, sh) * spitch;
      int v[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) v[q] = use_lut ? p.lut[row[xs[q]]] : row[xs[q]];
      r[k] = v[0] + v[4] + 4 * (v[1] + v[3]) + 6 * v[2];
    }
    dst[i] = (uint8_t)((r[0] + r[4] + 4 * (r[1] + r[3]) + 6 * r[2] + 128) >> 8);
  }
}

// ------------------------------------------------------------------------------------------------ LK tracking
struct LkImg {            // image in global memory, reflect-101 resolved per access (fallback path for huge ROIs)
  const uint8_t* p;
  int w, h, pitch;
  bool lut;
};
__device__ __forceinline__ int lk_px(const LkImg& im, const uint8_t* lut, int y, int x) {
  const uint8_t v = im.p[(long long)reflect101_multi(y, im.h) * im.pitch + reflect101_multi(x, im.w)];
  return im.lut ? lut[v] : v;
}
#define LK_PAD_EXTRA 1    // images staged in shared memory carry a reflect-101 border of win + LK_PAD_EXTRA pixels

__device__ __forceinline__ long long warp_sum_ll(long long v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }

// ---- window sums in cv::LKTrackerInvoker's float32 order ------------------------------------------------------------
// OpenCV accumulates the 2x2 gradient matrix and the mismatch vector in float32 (modules/video/src/lkpyramid.cpp, SSE2
// universal intrinsics; restated and pinned bit for bit against cv2 in oracle/np_kernels.py cv_window_sum /
// cv_mismatch_sums): per window row the first nsimd = 8*(win/8) columns go through a 4-lane accumulator (lane j takes
// columns j and j+4 of a group of 8, rows in order), the other columns through one scalar accumulator (row-major);
// total = scalar + ((q0 + q2) + (q1 + q3)).  The rounding of those partial sums decides, now and then, whether an
// iteration's exit test fires, and a carried point then differs by ~0.01 px for the rest of the clip -- so the order
// is part of "results identical to the reference".
// Window layout over the warp: lane 2y holds columns 0..7 of window row y, lane 2y+1 columns 8..15 (zero past the
// window; adding 0.0f changes nothing).  When the sum of |terms| over the window is below 2^24 every partial sum is an
// exactly representable integer and the float32 result IS the exact integer total: the callers test that first
// (lk_sums_exact) and only walk the rows in order when a partial sum can actually round.
__device__ __forceinline__ bool lk_sums_exact(unsigned lane_abs) {
  return __reduce_add_sync(0xffffffffu, min(lane_abs, 1u << 24)) < (1u << 24);
}
struct LkAcc {     // one sum's accumulators
  float q0, q1, q2, q3, sc;
  __device__ __forceinline__ void clear() { q0 = q1 = q2 = q3 = sc = 0.f; }
  // eight consecutive columns of one row, either as a SIMD group or as scalar columns
  __device__ __forceinline__ void simd(const float v[8]) {
    q0 = (q0 + v[0]) + v[4]; q1 = (q1 + v[1]) + v[5]; q2 = (q2 + v[2]) + v[6]; q3 = (q3 + v[3]) + v[7];
  }
  __device__ __forceinline__ void scalar(const float v[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) sc += v[k];
  }
  __device__ __forceinline__ float total(bool any_simd) const { return any_simd ? sc + ((q0 + q2) + (q1 + q3)) : sc; }
};
// Second tier (the window total can round but no single accumulator can): the five accumulators of a sum -- q0..q3 and the
// scalar one -- as exact integers, each below 2^24, combined in float32 the way OpenCV combines them.  On the bench clips the
// total of Ix^2 + Iy^2 passes 2^24 in 44 % of the windows, an accumulator only in 10 % (oracle statistics).
// Returns false when an accumulator can round: the caller then walks the rows in order (lk_cv_gradient_sums).
__device__ __forceinline__ bool lk_gradient_sums_by_chain(const int Ixv[8], const int Iyv[8], int win, int lane, float& A11,
                                                          float& A12, float& A22) {
  const int nsimd = (win >> 3) << 3;
  const bool simd_lane = (lane & 1) ? nsimd == 16 : nsimd >= 8;     // this lane's 8 columns are a SIMD group
  unsigned q11[4] = {0, 0, 0, 0}, q22[4] = {0, 0, 0, 0}, s11 = 0, s22 = 0;
  int q12[4] = {0, 0, 0, 0}, s12 = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int xx = Ixv[k] * Ixv[k], xy = Ixv[k] * Iyv[k], yy = Iyv[k] * Iyv[k];
    q11[k & 3] += xx; q12[k & 3] += xy; q22[k & 3] += yy;
    s11 += xx; s12 += xy; s22 += yy;
  }
  const unsigned cap = 1u << 24;
  unsigned t11[5], t22[5];
  int t12[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {       // fifteen independent warp reductions (no serial dependence between them)
    const unsigned v11 = j < 4 ? (simd_lane ? q11[j & 3] : 0u) : (simd_lane ? 0u : s11);
    const unsigned v22 = j < 4 ? (simd_lane ? q22[j & 3] : 0u) : (simd_lane ? 0u : s22);
    const int v12 = j < 4 ? (simd_lane ? q12[j & 3] : 0) : (simd_lane ? 0 : s12);
    t11[j] = __reduce_add_sync(0xffffffffu, min(v11, cap));
    t22[j] = __reduce_add_sync(0xffffffffu, min(v22, cap));
    t12[j] = __reduce_add_sync(0xffffffffu, v12);      // |v12| <= (v11 + v22) / 2 per lane: exact whenever the two above are
  }
  bool ok = true;
  float f11[5], f12[5], f22[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    ok = ok && t11[j] < cap && t22[j] < cap;             // then |sum of Ix Iy| <= (t11 + t22) / 2 < 2^24 as well
    f11[j] = (float)t11[j]; f22[j] = (float)t22[j]; f12[j] = (float)t12[j];
  }
  if (!ok) return false;
  A11 = nsimd ? f11[4] + ((f11[0] + f11[2]) + (f11[1] + f11[3])) : f11[4];
  A12 = nsimd ? f12[4] + ((f12[0] + f12[2]) + (f12[1] + f12[3])) : f12[4];
  A22 = nsimd ? f22[4] + ((f22[0] + f22[2]) + (f22[1] + f22[3])) : f22[4];
  return true;
}
// sums of Ix*Ix, Ix*Iy, Iy*Iy (products < 2^24: exact in float32) in OpenCV's order; every lane gets the result
__device__ __noinline__ void lk_cv_gradient_sums_impl(const int* Ixv, const int* Iyv, int win, float* out) {
  float A11, A12, A22;
  const int nsimd = (win >> 3) << 3;
  LkAcc a11, a12, a22;
  a11.clear(); a12.clear(); a22.clear();
  for (int y = 0; y < win; ++y) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float xx[8], xy[8], yy[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int ix = __shfl_sync(0xffffffffu, Ixv[k], 2 * y + half), iy = __shfl_sync(0xffffffffu, Iyv[k], 2 * y + half);
        xx[k] = (float)(ix * ix); xy[k] = (float)(ix * iy); yy[k] = (float)(iy * iy);
      }
      if (nsimd >= 8 * (half + 1)) { a11.simd(xx); a12.simd(xy); a22.simd(yy); }
      else { a11.scalar(xx); a12.scalar(xy); a22.scalar(yy); }
    }
  }
  A11 = a11.total(nsimd > 0); A12 = a12.total(nsimd > 0); A22 = a22.total(nsimd > 0);
  out[0] = A11; out[1] = A12; out[2] = A22;
}
// the rarely taken call works on copies in local memory: the callers' arrays stay in registers
__device__ __forceinline__ void lk_cv_gradient_sums(const int Ixv[8], const int Iyv[8], int win, float& A11, float& A12,
                                                    float& A22) {
  int tx[8], ty[8];
  float out[3];
#pragma unroll
  for (int k = 0; k < 8; ++k) { tx[k] = Ixv[k]; ty[k] = Iyv[k]; }
  lk_cv_gradient_sums_impl(tx, ty, win, out);
  A11 = out[0]; A12 = out[1]; A22 = out[2];
}
// sums of diff*Ix, diff*Iy in OpenCV's order: in a SIMD group the products of columns (j, j+4) are added as integers
// (v_dotprod) before the conversion to float32; scalar columns convert every product.  dx[k] = diff_k * Ix_k etc.
__device__ __noinline__ void lk_cv_mismatch_sums_impl(const int* dx, const int* dy, int win, float* out) {
  float b1, b2;
  const int nsimd = (win >> 3) << 3;
  LkAcc ax, ay;
  ax.clear(); ay.clear();
  for (int y = 0; y < win; ++y) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      int vx[8], vy[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        vx[k] = __shfl_sync(0xffffffffu, dx[k], 2 * y + half);
        vy[k] = __shfl_sync(0xffffffffu, dy[k], 2 * y + half);
      }
      if (nsimd >= 8 * (half + 1)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float px = (float)(vx[j] + vx[j + 4]), py = (float)(vy[j] + vy[j + 4]);
          if (j == 0) { ax.q0 += px; ay.q0 += py; }
          if (j == 1) { ax.q1 += px; ay.q1 += py; }
          if (j == 2) { ax.q2 += px; ay.q2 += py; }
          if (j == 3) { ax.q3 += px; ay.q3 += py; }
        }
      } else {
        float fx[8], fy[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { fx[k] = (float)vx[k]; fy[k] = (float)vy[k]; }
        ax.scalar(fx); ay.scalar(fy);
      }
    }
  }
  b1 = ax.total(nsimd > 0); b2 = ay.total(nsimd > 0);
  out[0] = b1; out[1] = b2;
}
__device__ __forceinline__ void lk_cv_mismatch_sums(const int dx[8], const int dy[8], int win, float& b1, float& b2) {
  int tx[8], ty[8];
  float out[2];
#pragma unroll
  for (int k = 0; k < 8; ++k) { tx[k] = dx[k]; ty[k] = dy[k]; }
  lk_cv_mismatch_sums_impl(tx, ty, win, out);
  b1 = out[0]; b2 = out[1];
}

// One warp tracks one point through all pyramid levels (cv::LKTrackerInvoker).  patch / deriv are per-warp shared
// scratch: (win+3)^2 ints and (win+1)^2 short2.  Returns status (1 = tracked).
template <typename Img>
__device__ int lk_track_point(const MeasureParams& p, const Img* prev, const Img* next, int nlev,
                              const uint8_t* lut, float px, float py, float* out_x, float* out_y, short* patch,
                              short2* deriv, int lane) {
  const int win = p.win;
  const int pw = win + 3, dwid = win + 1;
  const float half = (float)(win - 1) * 0.5f;
  const float FLT_SCALE = 1.0f / (float)(1 << 20);
  float nx = 0.f, ny = 0.f;
  int status = 1;
  for (int level = nlev - 1; level >= 0; --level) {
    const float inv = (float)(1.0 / (double)(1 << level));
    float ppx = px * inv, ppy = py * inv;
    if (level == nlev - 1) { nx = ppx; ny = ppy; }
    else { nx = nx * 2.f; ny = ny * 2.f; }
    const Img& I = prev[level];
    const Img& J = next[level];
    ppx -= half; ppy -= half;
    const int ix = (int)floorf(ppx), iy = (int)floorf(ppy);
    if (ix < -win || ix >= I.w || iy < -win || iy >= I.h) {
      if (level == 0) status = 0;
      continue;
    }
    float a = ppx - (float)ix, b = ppy - (float)iy;
    int iw00 = __float2int_rn((1.f - a) * (1.f - b) * 16384.f);
    int iw01 = __float2int_rn(a * (1.f - b) * 16384.f);
    int iw10 = __float2int_rn((1.f - a) * b * 16384.f);
    int iw11 = 16384 - iw00 - iw01 - iw10;
    // stage the (win+3)^2 neighbourhood of the previous image (reflect-101 padding like cv::copyMakeBorder)
    __syncwarp();
    for (int i = lane; i < pw * pw; i += 32) {
      const int r = i / pw, c = i - r * pw;
      patch[i] = (short)lk_px(I, lut, iy - 1 + r, ix - 1 + c);
    }
    __syncwarp();
    // Scharr derivatives (calcScharrDeriv) on the (win+1)^2 positions the window touches; zero outside the image
    for (int i = lane; i < dwid * dwid; i += 32) {
      const int y = i / dwid, x = i - y * dwid;
      short2 d = make_short2(0, 0);
      if (iy + y >= 0 && iy + y < I.h && ix + x >= 0 && ix + x < I.w) {
        const short* r0 = patch + y * pw + x;          // rows y-1, y, y+1 of the window position -> patch rows y..y+2
        const short* r1 = r0 + pw;
        const short* r2 = r1 + pw;
        const int t0l = (r0[0] + r2[0]) * 3 + r1[0] * 10, t0r = (r0[2] + r2[2]) * 3 + r1[2] * 10;
        const int t1l = r2[0] - r0[0], t1c = r2[1] - r0[1], t1r = r2[2] - r0[2];
        d.x = (short)(t0r - t0l);
        d.y = (short)((t1r + t1l) * 3 + t1c * 10);
      }
      deriv[i] = d;
    }
    __syncwarp();
    int Iw[8], Ixv[8], Iyv[8];
    long long sA11 = 0, sA12 = 0, sA22 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int wy = lane >> 1, wx = (lane & 1) * 8 + k;   // lane 2y: columns 0..7 of window row y, lane 2y+1: 8..15
      Iw[k] = 0; Ixv[k] = 0; Iyv[k] = 0;
      if (wy < win && wx < win) {
        const short* pr = patch + (wy + 1) * pw + wx + 1;
        Iw[k] = descale(pr[0] * iw00 + pr[1] * iw01 + pr[pw] * iw10 + pr[pw + 1] * iw11, 9);
        const short2 d00 = deriv[wy * dwid + wx], d01 = deriv[wy * dwid + wx + 1];
        const short2 d10 = deriv[(wy + 1) * dwid + wx], d11 = deriv[(wy + 1) * dwid + wx + 1];
        Ixv[k] = descale(d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11, 14);
        Iyv[k] = descale(d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11, 14);
        sA11 += (long long)Ixv[k] * Ixv[k];
        sA12 += (long long)Ixv[k] * Iyv[k];
        sA22 += (long long)Iyv[k] * Iyv[k];
      }
    }
    float A11, A12, A22;
    if (lk_sums_exact((unsigned)(sA11 + sA22))) {      // no partial sum can round: the exact totals are OpenCV's floats
      A11 = (float)warp_sum_ll(sA11); A12 = (float)warp_sum_ll(sA12); A22 = (float)warp_sum_ll(sA22);
    } else if (!lk_gradient_sums_by_chain(Ixv, Iyv, win, lane, A11, A12, A22)) {
      lk_cv_gradient_sums(Ixv, Iyv, win, A11, A12, A22);
    }
    A11 *= FLT_SCALE; A12 *= FLT_SCALE; A22 *= FLT_SCALE;
    float D = A11 * A22 - A12 * A12;
    const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
    if (minEig < p.min_eig || D < 1.1920929e-07f) {
      if (level == 0) status = 0;
      continue;
    }
    D = 1.f / D;
    float qx = nx - half, qy = ny - half;
    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < p.max_iter; ++j) {
      const int jx = (int)floorf(qx), jy = (int)floorf(qy);
      if (jx < -win || jx >= J.w || jy < -win || jy >= J.h) {
        if (level == 0) status = 0;
        break;
      }
      a = qx - (float)jx; b = qy - (float)jy;
      iw00 = __float2int_rn((1.f - a) * (1.f - b) * 16384.f);
      iw01 = __float2int_rn(a * (1.f - b) * 16384.f);
      iw10 = __float2int_rn((1.f - a) * b * 16384.f);
      iw11 = 16384 - iw00 - iw01 - iw10;
      long long sb1 = 0, sb2 = 0;
      int dxv[8], dyv[8];
      unsigned babs = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int wy = lane >> 1, wx = (lane & 1) * 8 + k;
        dxv[k] = 0; dyv[k] = 0;
        if (wy < win && wx < win) {
          const int j00 = lk_px(J, lut, jy + wy, jx + wx), j01 = lk_px(J, lut, jy + wy, jx + wx + 1);
          const int j10 = lk_px(J, lut, jy + wy + 1, jx + wx), j11 = lk_px(J, lut, jy + wy + 1, jx + wx + 1);
          const int diff = descale(j00 * iw00 + j01 * iw01 + j10 * iw10 + j11 * iw11, 9) - Iw[k];
          dxv[k] = diff * Ixv[k]; dyv[k] = diff * Iyv[k];
          sb1 += dxv[k]; sb2 += dyv[k];
          babs += (unsigned)abs(dxv[k]) + (unsigned)abs(dyv[k]);
        }
      }
      float b1, b2;
      if (lk_sums_exact(babs)) { b1 = (float)warp_sum_ll(sb1); b2 = (float)warp_sum_ll(sb2); }
      else lk_cv_mismatch_sums(dxv, dyv, win, b1, b2);
      b1 *= FLT_SCALE; b2 *= FLT_SCALE;
      const float dx = (A12 * b2 - A22 * b1) * D, dy = (A12 * b1 - A11 * b2) * D;
      qx += dx; qy += dy;
      nx = qx + half; ny = qy + half;
      if ((double)dx * (double)dx + (double)dy * (double)dy <= p.eps2) break;
      if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
        nx -= dx * 0.5f; ny -= dy * 0.5f;
        break;
      }
      pdx = dx; pdy = dy;
    }
  }
  *out_x = nx; *out_y = ny;
  return status;
}

__global__ void __launch_bounds__(LK_WARPS * 32) lk_track_kernel(const MeasureParams p, const float* pts0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int clip = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float s_pts[LK_MAX_PTS][2], s_new[LK_MAX_PTS][2];
  __shared__ int s_st[LK_MAX_PTS];
  __shared__ int s_n, s_lost;
  __shared__ uint8_t s_lut[256];
  const int pw = p.win + 3, dwid = p.win + 1;
  short* patch = reinterpret_cast<short*>(smem_raw) + (size_t)warp * (((pw * pw + 1) & ~1) + 2 * dwid * dwid);
  short2* deriv = reinterpret_cast<short2*>(patch + ((pw * pw + 1) & ~1));
  int rx, ry, rw, rh;
  const bool ok = roi_ok(p, clip, rx, ry, rw, rh);
  float* motion = p.motion + (long long)clip * p.n_frames * 2;
  if (!ok || p.npts[clip] <= 0) {
    for (int f = threadIdx.x; f < p.n_frames; f += blockDim.x) { motion[2 * f] = NAN; motion[2 * f + 1] = NAN; }
    return;
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = p.lut[i];
  if (threadIdx.x == 0) { s_n = p.npts[clip]; s_lost = 0; }
  for (int i = threadIdx.x; i < p.npts[clip]; i += blockDim.x) {
    s_pts[i][0] = pts0[((long long)clip * LK_MAX_PTS + i) * 2];
    s_pts[i][1] = pts0[((long long)clip * LK_MAX_PTS + i) * 2 + 1];
  }
  if (threadIdx.x == 0) { motion[0] = 0.f; motion[1] = 0.f; }
  __syncthreads();
  int lw[LK_MAX_LEVELS], lh[LK_MAX_LEVELS];
  const int nlev = lk_num_levels(rw, rh, p.win, p.max_level, lw, lh);
  for (int f = 1; f < p.n_frames; ++f) {
    LkImg prev[LK_MAX_LEVELS], next[LK_MAX_LEVELS];
    for (int l = 0; l < nlev; ++l) {
      if (l == 0) {
        const uint8_t* base = p.frames + ((long long)clip * p.T + p.first_frame) * p.W * p.H + (long long)ry * p.W + rx;
        prev[0] = {base + (long long)(f - 1) * p.W * p.H, rw, rh, p.W, true};
        next[0] = {base + (long long)f * p.W * p.H, rw, rh, p.W, true};
      } else {
        const uint8_t* base = p.pyr[l] + (long long)clip * p.n_frames * p.lvl_elems[l];
        prev[l] = {base + (long long)(f - 1) * p.lvl_elems[l], lw[l], lh[l], lw[l], false};
        next[l] = {base + (long long)f * p.lvl_elems[l], lw[l], lh[l], lw[l], false};
      }
    }
    const int n = s_n;
    for (int i = warp; i < n; i += LK_WARPS) {
      float ox, oy;
      const int st = lk_track_point(p, prev, next, nlev, s_lut, s_pts[i][0], s_pts[i][1], &ox, &oy, patch, deriv, lane);
      if (lane == 0) { s_new[i][0] = ox; s_new[i][1] = oy; s_st[i] = st; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      // good_new = p1[st == 1], good_old = pts[st == 1]; mean(good_old - good_new, axis=0) in float32 (base.py:377-389)
      int m = 0;
      float sx = 0.f, sy = 0.f;
      for (int i = 0; i < n; ++i) {
        if (s_st[i]) {
          sx += s_pts[i][0] - s_new[i][0];
          sy += s_pts[i][1] - s_new[i][1];
          s_pts[m][0] = s_new[i][0];
          s_pts[m][1] = s_new[i][1];
          ++m;
        }
      }
      s_n = m;
      if (m == 0) {
        s_lost = 1;
      } else {
        motion[2 * f] = sx / (float)m;
        motion[2 * f + 1] = sy / (float)m;
      }
    }
    __syncthreads();
    if (p.pts_dbg) {
      float* dbg = p.pts_dbg + ((long long)clip * p.n_frames + f) * LK_MAX_PTS * 2;
      for (int i = threadIdx.x; i < LK_MAX_PTS; i += blockDim.x) {
        dbg[2 * i] = i < s_n ? s_pts[i][0] : NAN;
        dbg[2 * i + 1] = i < s_n ? s_pts[i][1] : NAN;
      }
    }
    if (s_lost) {   // tracking lost: extract_motion returns nan from here on (base.py:385-386)
      for (int g = f + threadIdx.x; g < p.n_frames; g += blockDim.x) { motion[2 * g] = NAN; motion[2 * g + 1] = NAN; }
      if (threadIdx.x == 0) p.status[clip] = RM_CLIP_TRACK_LOST;
      return;
    }
  }
}

// ------------------------------------------------------------------------------------------------ LK tracking, shared memory
// The production path: the ROI crops of consecutive frames live in shared memory with the LUT applied and a
// reflect-101 border of win+1 pixels materialised, so the iteration reads pixels without any border arithmetic and
// never touches global memory.  The crop of frame f+1 is copied in with cp.async while frame f is being tracked; the
// uint8 pyramid levels (cv2.buildOpticalFlowPyramid) are rebuilt in shared memory per frame.  Everything is addressed
// by offsets into the one dynamic shared array (keeps the accesses LDS/STS), loops are division free, and the window
// sums are reduced with REDUX (two 32-bit reductions per exact 64-bit sum).
#ifndef LKS_WARPS
#define LKS_WARPS 16
#endif
extern __shared__ __align__(16) unsigned char lks_smem[];

struct LkSmemLayout {
  int nlev;
  int lw[LK_MAX_LEVELS], lh[LK_MAX_LEVELS];
  int pitch[LK_MAX_LEVELS], off[LK_MAX_LEVELS];   // padded pitch / byte offset of padded level l inside one pyramid
  int pad;
  int pyr_bytes;                                  // one padded pyramid
  int raw_pitch, raw_bytes;                       // crop as copied from the frame (columns aligned down/up to 4)
  int sums_off;                                   // per-warp scratch of lks_gradient_sums_ordered (LKS_SUMS_BYTES each)
  int total;
};
#define LKS_SUMS_BYTES (3 * 16 * 16 * 4)          // three matrices x 16 window rows x 16 columns, float32
__host__ __device__ inline LkSmemLayout lk_smem_layout(int rw, int rh, int win, int max_level, int warps) {
  LkSmemLayout L;
  L.pad = win + LK_PAD_EXTRA;
  L.nlev = 1;
  L.lw[0] = rw; L.lh[0] = rh;
  while (L.nlev <= max_level && L.nlev < LK_MAX_LEVELS) {
    const int nw = (L.lw[L.nlev - 1] + 1) / 2, nh = (L.lh[L.nlev - 1] + 1) / 2;
    if (nw <= win || nh <= win) break;
    L.lw[L.nlev] = nw; L.lh[L.nlev] = nh;
    ++L.nlev;
  }
  int off = 0;
  for (int l = 0; l < L.nlev; ++l) {
    L.pitch[l] = (L.lw[l] + 2 * L.pad + 3) & ~3;   // word-aligned rows: the pyramid build moves 4 pixels at a time
    L.off[l] = off;
    off += (L.pitch[l] * (L.lh[l] + 2 * L.pad) + 15) & ~15;
  }
  L.pyr_bytes = off;
  L.raw_pitch = (rw + 3 + 3) & ~3;
  L.raw_bytes = (L.raw_pitch * rh + 15) & ~15;
  L.sums_off = 2 * L.pyr_bytes + 2 * L.raw_bytes;
  L.total = L.sums_off + warps * LKS_SUMS_BYTES;
  return L;
}

__device__ __forceinline__ void lk_cp_async4(unsigned smem_addr, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
// exact sum over the warp of per-lane partials |v| < 2^30: two 32-bit REDUX instead of five 64-bit shuffle steps
__device__ __forceinline__ long long warp_sum_split(int v) {
  const int lo = v & 0x3fff, hi = v >> 14;
  return ((long long)__reduce_add_sync(0xffffffffu, hi) << 14) + (long long)__reduce_add_sync(0xffffffffu, lo);
}

struct LkSLevel {   // one padded level in shared memory: `org` is the offset of pixel (0,0)
  int org, w, h, pitch;
};

// The gradient sums in OpenCV's order (see lk_cv_gradient_sums) when a partial sum can round -- 44 % of the windows on
// the bench clips -- without moving every product through a shuffle: the lanes park their float products in the warp's
// scratch (row y of the window = 16 consecutive floats per matrix), then fifteen lanes walk the fifteen accumulators
// (q0..q3 of the three matrices in lanes 0..11, the scalar accumulators in lanes 12..14) over the rows in order.
// About 210 warp instructions per window against 1950 through shuffles (ncu r02j: the shuffle form doubled the kernel).
__device__ __forceinline__ void lks_gradient_sums_ordered(float* S, const int Ixv[8], const int Iyv[8], int win, int lane,
                                                          float& A11, float& A12, float& A22) {
  const int nsimd = (win >> 3) << 3;
  __syncwarp();
  {
    float4* d = reinterpret_cast<float4*>(S + (lane >> 1) * 16 + (lane & 1) * 8);    // lane 2y + half: row y, columns 8 half ..
    float xx[8], xy[8], yy[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      xx[k] = (float)(Ixv[k] * Ixv[k]); xy[k] = (float)(Ixv[k] * Iyv[k]); yy[k] = (float)(Iyv[k] * Iyv[k]);   // < 2^24: exact
    }
    d[0] = make_float4(xx[0], xx[1], xx[2], xx[3]); d[1] = make_float4(xx[4], xx[5], xx[6], xx[7]);
    d[64] = make_float4(xy[0], xy[1], xy[2], xy[3]); d[65] = make_float4(xy[4], xy[5], xy[6], xy[7]);
    d[128] = make_float4(yy[0], yy[1], yy[2], yy[3]); d[129] = make_float4(yy[4], yy[5], yy[6], yy[7]);
  }
  __syncwarp();
  float acc = 0.f;
  if (lane < 12) {                                  // q_j of matrix lane / 4: columns j, j+4 of every SIMD group, rows in order
    const float* r = S + (lane >> 2) * 256 + (lane & 3);
    if (nsimd >= 8)
      for (int y = 0; y < win; ++y, r += 16) {
        acc = (acc + r[0]) + r[4];
        if (nsimd == 16) acc = (acc + r[8]) + r[12];
      }
  } else if (lane < 15) {                           // the scalar accumulator of matrix lane - 12: the other columns, row-major
    const float* r = S + (lane - 12) * 256;
    for (int y = 0; y < win; ++y, r += 16) {
      if (nsimd == 0) {
#pragma unroll
        for (int x = 0; x < 8; ++x) acc += r[x];
      }
      if (nsimd < 16) {
        const float4 a = *reinterpret_cast<const float4*>(r + 8), b = *reinterpret_cast<const float4*>(r + 12);
        acc += a.x; acc += a.y; acc += a.z; acc += a.w; acc += b.x; acc += b.y; acc += b.z; acc += b.w;
      }
    }
  }
  __syncwarp();
  float t[3];
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    const float q0 = __shfl_sync(0xffffffffu, acc, 4 * m), q1 = __shfl_sync(0xffffffffu, acc, 4 * m + 1);
    const float q2 = __shfl_sync(0xffffffffu, acc, 4 * m + 2), q3 = __shfl_sync(0xffffffffu, acc, 4 * m + 3);
    const float sc = __shfl_sync(0xffffffffu, acc, 12 + m);
    t[m] = nsimd ? sc + ((q0 + q2) + (q1 + q3)) : sc;
  }
  A11 = t[0]; A12 = t[1]; A22 = t[2];
}

// One warp tracks one point through the levels (cv::LKTrackerInvoker), images in shared memory.
// wq[k] = offset (wy * pitch-independent pair) of the lane's k-th window pixel: wy = wq >> 8, wx = wq & 255.
__device__ int lks_track_point(const MeasureParams& p, const LkSLevel* prev, const LkSLevel* next, int nlev, float px,
                               float py, float* out_x, float* out_y, const int wq[8], float* sums_scratch,
                               int lane) {
  const int win = p.win;
  const float half = (float)(win - 1) * 0.5f;
  const float FLT_SCALE = 1.0f / (float)(1 << 20);
  float nx = 0.f, ny = 0.f;
  int status = 1;
  for (int level = nlev - 1; level >= 0; --level) {
    const float inv = (float)(1.0 / (double)(1 << level));
    float ppx = px * inv, ppy = py * inv;
    if (level == nlev - 1) { nx = ppx; ny = ppy; }
    else { nx = nx * 2.f; ny = ny * 2.f; }
    const LkSLevel I = prev[level];
    const LkSLevel J = next[level];
    ppx -= half; ppy -= half;
    const int ix = (int)floorf(ppx), iy = (int)floorf(ppy);
    if (ix < -win || ix >= I.w || iy < -win || iy >= I.h) {
      if (level == 0) status = 0;
      continue;
    }
    float a = ppx - (float)ix, b = ppy - (float)iy;
    int iw00 = __float2int_rn((1.f - a) * (1.f - b) * 16384.f);
    int iw01 = __float2int_rn(a * (1.f - b) * 16384.f);
    int iw10 = __float2int_rn((1.f - a) * b * 16384.f);
    int iw11 = 16384 - iw00 - iw01 - iw10;
    // The lane's 8 window pixels sit side by side in window row wy: everything they need of the previous image -- the
    // intensities for the bilinear sample and the 3x3 neighbourhoods of the Scharr derivatives (calcScharrDeriv) at
    // rows wy, wy+1 and columns wx0..wx0+8 -- is a 4 x 11 byte block, walked column by column in registers.  The
    // derivative image is zero outside the image (cv::copyMakeBorder CONSTANT), the intensity image is reflect-padded.
    int Iw[8], Ixv[8], Iyv[8];
    int sA11 = 0, sA12 = 0, sA22 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { Iw[k] = 0; Ixv[k] = 0; Iyv[k] = 0; }
    if (wq[0] >= 0) {
      const int wy = wq[0] >> 8, wx0 = wq[0] & 255;
      const unsigned char* src = lks_smem + I.org + (iy + wy - 1) * I.pitch + (ix + wx0 - 1);
      const bool row_ok0 = iy + wy >= 0 && iy + wy < I.h, row_ok1 = iy + wy + 1 >= 0 && iy + wy + 1 < I.h;
      int c0[4], c1[4], c2[4];   // three neighbouring columns of the block, rows wy-1 .. wy+2
#pragma unroll
      for (int r = 0; r < 4; ++r) { c0[r] = src[r * I.pitch]; c1[r] = src[r * I.pitch + 1]; }
      int pdx0 = 0, pdy0 = 0, pdx1 = 0, pdy1 = 0;   // derivatives of the previous column at rows wy, wy+1
#pragma unroll
      for (int cc = 0; cc <= 8; ++cc) {
#pragma unroll
        for (int r = 0; r < 4; ++r) c2[r] = src[r * I.pitch + cc + 2];
        const bool col_ok = ix + wx0 + cc >= 0 && ix + wx0 + cc < I.w;
        // Scharr at (wy, wx0+cc) from rows 0..2 and at (wy+1, wx0+cc) from rows 1..3
        int dx0 = (short)(((c2[0] + c2[2]) * 3 + c2[1] * 10) - ((c0[0] + c0[2]) * 3 + c0[1] * 10));
        int dy0 = (short)(((c2[2] - c2[0]) + (c0[2] - c0[0])) * 3 + (c1[2] - c1[0]) * 10);
        int dx1 = (short)(((c2[1] + c2[3]) * 3 + c2[2] * 10) - ((c0[1] + c0[3]) * 3 + c0[2] * 10));
        int dy1 = (short)(((c2[3] - c2[1]) + (c0[3] - c0[1])) * 3 + (c1[3] - c1[1]) * 10);
        if (!(col_ok && row_ok0)) { dx0 = 0; dy0 = 0; }
        if (!(col_ok && row_ok1)) { dx1 = 0; dy1 = 0; }
        if (cc >= 1) {
          const int k = cc - 1;            // pixel k: columns k (previous) and k+1 (this one)
          if (wx0 + k < win) {
            // at this point c0 = column k, c1 = column k+1 of the intensities (block columns k+1, k+2 before the shift)
            Iw[k] = descale(c0[1] * iw00 + c1[1] * iw01 + c0[2] * iw10 + c1[2] * iw11, 9);
            Ixv[k] = descale(pdx0 * iw00 + dx0 * iw01 + pdx1 * iw10 + dx1 * iw11, 14);
            Iyv[k] = descale(pdy0 * iw00 + dy0 * iw01 + pdy1 * iw10 + dy1 * iw11, 14);
            sA11 += Ixv[k] * Ixv[k];      // |Ix| <= 4080: eight products stay far below 2^30
            sA12 += Ixv[k] * Iyv[k];
            sA22 += Iyv[k] * Iyv[k];
          }
        }
        pdx0 = dx0; pdy0 = dy0; pdx1 = dx1; pdy1 = dy1;
#pragma unroll
        for (int r = 0; r < 4; ++r) { c0[r] = c1[r]; c1[r] = c2[r]; }
      }
    }
    float A11, A12, A22;
    // On textured ROIs the window total of Ix^2 + Iy^2 is practically never below 2^24 (bench clips: 0.5 % of the windows,
    // counted with an instrumented build, r02x), so the first tier is skipped here: 58 % of the windows are settled by the per-accumulator
    // sums, 41 % walk the rows in order.
    (void)sA11; (void)sA12; (void)sA22;
    if (!lk_gradient_sums_by_chain(Ixv, Iyv, win, lane, A11, A12, A22)) {
      lks_gradient_sums_ordered(sums_scratch, Ixv, Iyv, win, lane, A11, A12, A22);
    }
    A11 *= FLT_SCALE; A12 *= FLT_SCALE; A22 *= FLT_SCALE;
    float D = A11 * A22 - A12 * A12;
    const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
    if (minEig < p.min_eig || D < 1.1920929e-07f) {
      if (level == 0) status = 0;
      continue;
    }
    D = 1.f / D;
    unsigned aI[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) aI[k] = (unsigned)(abs(Ixv[k]) + abs(Iyv[k]));
    float qx = nx - half, qy = ny - half;
    float pdx = 0.f, pdy = 0.f;
    // the lane's window pixels are 8 consecutive columns of one row (wq[0] is the first): 2 x 9 bytes of J per iteration
    const int jrow = (wq[0] >= 0 ? (wq[0] >> 8) : 0) * J.pitch + (wq[0] >= 0 ? (wq[0] & 255) : 0);
    for (int j = 0; j < p.max_iter; ++j) {
      const int jx = (int)floorf(qx), jy = (int)floorf(qy);
      if (jx < -win || jx >= J.w || jy < -win || jy >= J.h) {
        if (level == 0) status = 0;
        break;
      }
      a = qx - (float)jx; b = qy - (float)jy;
      iw00 = __float2int_rn((1.f - a) * (1.f - b) * 16384.f);
      iw01 = __float2int_rn(a * (1.f - b) * 16384.f);
      iw10 = __float2int_rn((1.f - a) * b * 16384.f);
      iw11 = 16384 - iw00 - iw01 - iw10;
      const unsigned char* r0 = lks_smem + J.org + jy * J.pitch + jx + jrow;
      const unsigned char* r1 = r0 + J.pitch;
      int sb1 = 0, sb2 = 0;
      int diffv[8];
      unsigned babs = 0;             // sum of |diff| (|Ix| + |Iy|) >= sum of |diff Ix| + |diff Iy|
      int t0 = r0[0], t1 = r1[0];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int u0 = r0[k + 1], u1 = r1[k + 1];
        // pixels past the end of the window carry Ix = Iy = 0: whatever they read contributes nothing
        const int diff = descale(t0 * iw00 + u0 * iw01 + t1 * iw10 + u1 * iw11, 9) - Iw[k];
        diffv[k] = diff;
        sb1 += diff * Ixv[k];       // |diff| <= 8160, |Ix| <= 4080: eight products stay below 2^30
        sb2 += diff * Iyv[k];
        babs += (unsigned)abs(diff) * aI[k];
        t0 = u0; t1 = u1;
      }
      float b1, b2;
      if (lk_sums_exact(babs)) { b1 = (float)warp_sum_split(sb1); b2 = (float)warp_sum_split(sb2); }
      else {                         // a partial sum can round (never seen on the bench clips): rows in order
        int dxv[8], dyv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { dxv[k] = diffv[k] * Ixv[k]; dyv[k] = diffv[k] * Iyv[k]; }
        lk_cv_mismatch_sums(dxv, dyv, win, b1, b2);
      }
      b1 *= FLT_SCALE; b2 *= FLT_SCALE;
      const float dx = (A12 * b2 - A22 * b1) * D, dy = (A12 * b1 - A11 * b2) * D;
      qx += dx; qy += dy;
      nx = qx + half; ny = qy + half;
      if ((double)dx * (double)dx + (double)dy * (double)dy <= p.eps2) break;
      if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
        nx -= dx * 0.5f; ny -= dy * 0.5f;
        break;
      }
      pdx = dx; pdy = dy;
    }
  }
  *out_x = nx; *out_y = ny;
  return status;
}

__global__ void __launch_bounds__(LKS_WARPS * 32, 512 / (LKS_WARPS * 32)) lk_track_smem_kernel(const MeasureParams p, const float* pts0,
                                                                       int max_total) {
  const int clip = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int blk = blockIdx.y;                  // which slice of the clip's corners this block tracks
  __shared__ float s_pts[LK_MAX_PTS][2], s_new[LK_MAX_PTS][2];
  __shared__ int s_st[LK_MAX_PTS], s_idx[LK_MAX_PTS];
  __shared__ int s_n;
  __shared__ uint8_t s_lut[256];
  int rx, ry, rw, rh;
  const bool ok = roi_ok(p, clip, rx, ry, rw, rh);
  const bool resume = p.f0 > 0;               // a later chunk: points and count come from the previous launch
  int n_start = 0;
  if (resume) n_start = p.st_n[clip * LK_MAX_BLOCKS + blk];
  else if (ok && p.npts[clip] > 0) { n_start = p.npts[clip] - blk * p.ppb; if (n_start > p.ppb) n_start = p.ppb; }
  if (n_start <= 0) {   // no corners in this slice, a clip without ROI / corners (motion_reduce_kernel writes its NaNs),
                        // or every point of the slice lost in an earlier chunk
    if (!resume && p.st_n && tid == 0) p.st_n[clip * LK_MAX_BLOCKS + blk] = 0;
    return;
  }
  const LkSmemLayout L = lk_smem_layout(rw, rh, p.win, p.max_level, LKS_WARPS);
  if (L.total > max_total) return;   // cannot happen: the host sized shared memory for the largest ROI
  const int raw_base = 2 * L.pyr_bytes;

  for (int i = tid; i < 256; i += blockDim.x) s_lut[i] = p.lut[i];
  if (tid == 0) s_n = n_start;
  {
    const float* src = resume ? p.st_pts : pts0;
    for (int i = tid; i < n_start; i += blockDim.x) {
      const long long slot = (long long)clip * LK_MAX_PTS + blk * p.ppb + i;
      s_pts[i][0] = src[slot * 2];
      s_pts[i][1] = src[slot * 2 + 1];
      s_idx[i] = resume ? p.st_idx[slot] : blk * p.ppb + i;
    }
  }
  int wq[8];   // the lane's window pixels (wy << 8 | wx), -1 past the end of the window: row lane/2, columns 8*(lane&1)..+7
  {
    const int wy = lane >> 1, wx0 = (lane & 1) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) wq[k] = (wy < p.win && wx0 + k < p.win) ? (wy << 8) | (wx0 + k) : -1;
  }

  const long long frame_elems = (long long)p.W * p.H;
  const uint8_t* clip_base = p.frames + ((long long)clip * p.T + p.first_frame) * frame_elems;
  const int x_al = rx & ~3;
  const int raw_cols = ((rx + rw + 3) & ~3) - x_al;
  const bool aligned = (((unsigned long long)p.frames & 3) == 0) && (p.W % 4 == 0) && (frame_elems % 4 == 0);
  const int xoff = rx - x_al;
  const unsigned smem_base = (unsigned)__cvta_generic_to_shared(lks_smem);

  // the thread's first element and per-step increments of a block-strided 2-D loop over `cols` columns
  auto stride2d = [&](int cols, int& r0, int& c0, int& dr, int& dc) {
    r0 = tid / cols; c0 = tid - r0 * cols;
    dr = (int)blockDim.x / cols; dc = (int)blockDim.x - dr * cols;
  };

  auto stage_raw = [&](int f, int slot) {
    const uint8_t* src = clip_base + (long long)(p.ring > 0 ? f % p.ring : f) * frame_elems + (long long)ry * p.W + x_al;
    const int dst = raw_base + slot * L.raw_bytes;
    if (aligned) {
      const int chunks = raw_cols >> 2;
      int r, c, dr, dc;
      stride2d(chunks, r, c, dr, dc);
      for (; r < rh; r += dr) {
        lk_cp_async4(smem_base + dst + r * L.raw_pitch + c * 4, src + (long long)r * p.W + c * 4);
        c += dc;
        if (c >= chunks) { c -= chunks; ++r; }
      }
    } else {
      int r, c, dr, dc;
      stride2d(rw, r, c, dr, dc);
      for (; r < rh; r += dr) {
        lks_smem[dst + r * L.raw_pitch + xoff + c] = src[(long long)r * p.W + xoff + c];
        c += dc;
        if (c >= rw) { c -= rw; ++r; }
      }
    }
    cp_async_commit();
  };
  // raw crop -> padded uint8 pyramid (LUT, reflect-101 border; pyrDown levels with their own borders)
  // Linear block-strided loops (index -> row/column by one multiply-high: exact for n * d < 2^32) that the compiler can
  // unroll, so several independent load -> LUT -> store chains are in flight per thread; the stage is latency bound.
  auto magic_of = [](int d) { return (unsigned)((0x100000000ull + (unsigned)d - 1u) / (unsigned)d); };
  auto refl1 = [](int v, int n) { return v < 0 ? -v : (v >= n ? 2 * n - 2 - v : v); };   // one reflection (n > |overhang|)
  const int nthr = LKS_WARPS * 32;
  auto build = [&](int raw_slot, int pyr_slot) {
    const int src = raw_base + raw_slot * L.raw_bytes + xoff;
    const int base = pyr_slot * L.pyr_bytes;
    const bool words = (L.pad & 3) == 0;           // interior columns start on a word boundary
    if (words && rw > L.pad + 4 && rh > L.pad) {
      // level 0, border included, one 32-bit word (4 pixels) per step: interior words are two aligned loads of the raw
      // crop funnel-shifted to its byte offset, four LUT look-ups and one store; words that touch the left / right border
      // take their pixels one by one through the reflection
      const int pitch = L.pitch[0], ph = rh + 2 * L.pad, nw = pitch >> 2;
      const unsigned mg = magic_of(nw);
      const int total = ph * nw;
      unsigned* dst = reinterpret_cast<unsigned*>(lks_smem + base + L.off[0]);
      const int raw0 = raw_base + raw_slot * L.raw_bytes;   // word-aligned start of the raw rows
#pragma unroll 4
      for (int i = tid; i < total; i += nthr) {
        const int py = (int)__umulhi((unsigned)i, mg), j = i - py * nw;
        const int y = refl1(py - L.pad, rh);
        const int x0 = 4 * j - L.pad;
        unsigned v;
        if (x0 >= 0 && x0 + 3 < rw) {
          const int o = y * L.raw_pitch + xoff + x0;
          const unsigned* wp = reinterpret_cast<const unsigned*>(lks_smem + raw0 + (o & ~3));
          v = __funnelshift_r(wp[0], wp[1], (o & 3) * 8);
        } else {
          const unsigned char* row = lks_smem + raw0 + y * L.raw_pitch + xoff;
          v = 0;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            int x = refl1(x0 + b, rw);
            x = x < 0 ? 0 : (x >= rw ? rw - 1 : x);            // slack columns past the padded width
            v |= (unsigned)row[x] << (8 * b);
          }
        }
        dst[py * nw + j] = (unsigned)s_lut[v & 255] | ((unsigned)s_lut[(v >> 8) & 255] << 8) |
                           ((unsigned)s_lut[(v >> 16) & 255] << 16) | ((unsigned)s_lut[v >> 24] << 24);
      }
    } else {
      const int pitch = L.pitch[0], ph = rh + 2 * L.pad, pwid = rw + 2 * L.pad;
      const unsigned mg = magic_of(pwid);
      const int total = ph * pwid;
      unsigned char* dst = lks_smem + base + L.off[0];
      const unsigned char* raw = lks_smem + src;
      if (rw > L.pad && rh > L.pad) {
#pragma unroll 4
        for (int i = tid; i < total; i += nthr) {
          const int py = (int)__umulhi((unsigned)i, mg), pxx = i - py * pwid;
          const int y = refl1(py - L.pad, rh), x = refl1(pxx - L.pad, rw);
          dst[py * pitch + pxx] = s_lut[raw[y * L.raw_pitch + x]];
        }
      } else {
        for (int i = tid; i < total; i += nthr) {
          const int py = (int)__umulhi((unsigned)i, mg), pxx = i - py * pwid;
          const int y = reflect101_multi(py - L.pad, rh), x = reflect101_multi(pxx - L.pad, rw);
          dst[py * pitch + pxx] = s_lut[raw[y * L.raw_pitch + x]];
        }
      }
    }
    __syncthreads();
    for (int l = 1; l < L.nlev; ++l) {
      const int spitch = L.pitch[l - 1], dpitch = L.pitch[l];
      const int s0 = base + L.off[l - 1] + L.pad * spitch + L.pad;
      const int d0 = base + L.off[l] + L.pad * dpitch + L.pad;
      const int dw = L.lw[l], dh = L.lh[l];
      if (words) {
        // cv::pyrDown (uint8), four outputs per step: per source row four aligned words hold the 11 bytes the four
        // 5-tap windows span (they start two bytes into the first word); the taps 1 4 6 4 are one dp4a each
        const int qw = (dw + 3) >> 2;
        const unsigned mg = magic_of(qw);
        const int total = qw * dh;
#pragma unroll 1
        for (int i = tid; i < total; i += nthr) {
          const int y = (int)__umulhi((unsigned)i, mg), x = 4 * (i - y * qw);
          const unsigned* c = reinterpret_cast<const unsigned*>(lks_smem + s0 + (2 * y - 2) * spitch + 2 * x - 4);
          int r[4][5];
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const unsigned* row = c + k * (spitch >> 2);
            const unsigned w0 = row[0], w1 = row[1], w2 = row[2], w3 = row[3];
            r[0][k] = (int)__dp4a(__funnelshift_r(w0, w1, 16), 0x04060401u, (w1 >> 16) & 255u);
            r[1][k] = (int)__dp4a(w1, 0x04060401u, w2 & 255u);
            r[2][k] = (int)__dp4a(__funnelshift_r(w1, w2, 16), 0x04060401u, (w2 >> 16) & 255u);
            r[3][k] = (int)__dp4a(w2, 0x04060401u, w3 & 255u);
          }
          unsigned out = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            out |= (unsigned)((r[q][0] + r[q][4] + 4 * (r[q][1] + r[q][3]) + 6 * r[q][2] + 128) >> 8) << (8 * q);
          *reinterpret_cast<unsigned*>(lks_smem + d0 + y * dpitch + x) = out;   // columns past dw land in the border,
        }                                                                         // which the next pass overwrites
      } else {
        const unsigned mg = magic_of(dw);
        const int total = dw * dh;
#pragma unroll 2
        for (int i = tid; i < total; i += nthr) {
          const int y = (int)__umulhi((unsigned)i, mg), x = i - y * dw;
          const unsigned char* c = lks_smem + s0 + (2 * y - 2) * spitch + 2 * x - 2;
          int r[5];
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const unsigned char* row = c + k * spitch;
            r[k] = row[0] + row[4] + 4 * (row[1] + row[3]) + 6 * row[2];
          }
          lks_smem[d0 + y * dpitch + x] = (unsigned char)((r[0] + r[4] + 4 * (r[1] + r[3]) + 6 * r[2] + 128) >> 8);
        }
      }
      __syncthreads();
      {
        // border pixels only: `pad` full rows above and below, 2 * pad columns beside every image row
        const int pad = L.pad, pwid = dw + 2 * pad;
        const int n_tb = 2 * pad * pwid, total = n_tb + dh * 2 * pad;
        const unsigned mg = magic_of(pwid), mg2 = magic_of(2 * pad);
        const bool simple = dw > pad && dh > pad;
#pragma unroll 4
        for (int i = tid; i < total; i += nthr) {
          int yy, xx;
          if (i < n_tb) {
            const int r = (int)__umulhi((unsigned)i, mg);
            xx = i - r * pwid - pad;
            yy = r < pad ? r - pad : dh + (r - pad);
          } else {
            const int j = i - n_tb, r = (int)__umulhi((unsigned)j, mg2), k = j - r * 2 * pad;
            yy = r;
            xx = k < pad ? k - pad : dw + (k - pad);
          }
          const int sy = simple ? refl1(yy, dh) : reflect101_multi(yy, dh);
          const int sx = simple ? refl1(xx, dw) : reflect101_multi(xx, dw);
          lks_smem[d0 + yy * dpitch + xx] = lks_smem[d0 + sy * dpitch + sx];
        }
      }
      __syncthreads();
    }
  };
  auto levels = [&](int pyr_slot, LkSLevel* out) {
#pragma unroll
    for (int l = 0; l < LK_MAX_LEVELS; ++l)
      if (l < L.nlev)
        out[l] = {pyr_slot * L.pyr_bytes + L.off[l] + L.pad * L.pitch[l] + L.pad, L.lw[l], L.lh[l], L.pitch[l]};
  };

  // the frame before the chunk is the tracker's "previous image": frame 0 for the first chunk
  const int f_first = resume ? p.f0 : 1;
  stage_raw(f_first - 1, (f_first - 1) & 1);
  cp_async_wait<0>();
  __syncthreads();
  build((f_first - 1) & 1, (f_first - 1) & 1);
  if (f_first < p.f1) stage_raw(f_first, f_first & 1);
  for (int f = f_first; f < p.f1; ++f) {
    cp_async_wait<0>();
    __syncthreads();
    build(f & 1, f & 1);
    if (f + 1 < p.f1) stage_raw(f + 1, (f + 1) & 1);   // lands while this frame is tracked
    LkSLevel prev[LK_MAX_LEVELS], next[LK_MAX_LEVELS];
    levels((f - 1) & 1, prev);
    levels(f & 1, next);
    const int n = s_n;
    for (int i = warp; i < n; i += LKS_WARPS) {
      float ox, oy;
      const int st = lks_track_point(p, prev, next, L.nlev, s_pts[i][0], s_pts[i][1], &ox, &oy, wq,
                                     reinterpret_cast<float*>(lks_smem + L.sums_off) + warp * (LKS_SUMS_BYTES / 4), lane);
      if (lane == 0) { s_new[i][0] = ox; s_new[i][1] = oy; s_st[i] = st; }
    }
    __syncthreads();
    if (tid == 0) {
      // good_new = p1[st == 1], good_old = pts[st == 1] (base.py:377-382): survivors keep their relative order; their
      // old - new goes out under the point's original index for the ordered float32 mean (base.py:388)
      float2* out = p.delta + ((long long)clip * p.delta_frames + (f - p.delta_f0)) * LK_MAX_PTS;
      int m = 0;
      for (int i = 0; i < n; ++i) {
        if (s_st[i]) {
          out[s_idx[i]] = make_float2(s_pts[i][0] - s_new[i][0], s_pts[i][1] - s_new[i][1]);
          s_pts[m][0] = s_new[i][0];
          s_pts[m][1] = s_new[i][1];
          s_idx[m] = s_idx[i];
          ++m;
        }
      }
      s_n = m;
    }
    __syncthreads();
    if (p.pts_dbg) {   // single-block mode only (bpc == 1): the compacted list is the reference's motion_key_points
      float* dbg = p.pts_dbg + ((long long)clip * p.n_frames + f) * LK_MAX_PTS * 2;
      for (int i = tid; i < LK_MAX_PTS; i += blockDim.x) {
        dbg[2 * i] = i < s_n ? s_pts[i][0] : NAN;
        dbg[2 * i + 1] = i < s_n ? s_pts[i][1] : NAN;
      }
    }
    if (s_n == 0) {    // every corner of this slice is gone; nothing left to track here
      if (tid == 0 && p.st_n) p.st_n[clip * LK_MAX_BLOCKS + blk] = 0;
      cp_async_wait<0>();
      return;
    }
  }
  __syncthreads();
  if (p.st_n) {   // hand the tracker state to the next chunk
    for (int i = tid; i < s_n; i += blockDim.x) {
      const long long slot = (long long)clip * LK_MAX_PTS + blk * p.ppb + i;
      p.st_pts[slot * 2] = s_pts[i][0];
      p.st_pts[slot * 2 + 1] = s_pts[i][1];
      p.st_idx[slot] = s_idx[i];
    }
    if (tid == 0) p.st_n[clip * LK_MAX_BLOCKS + blk] = s_n;
  }
}

// mean(good_old - good_new, axis=0) (base.py:388) for frames [f0, f1): float32, the surviving points in their original
// order (numpy reduces axis 0 of an (N,2) array row by row).  A frame without survivors is where extract_motion starts
// returning nan (base.py:385-386): the clip is marked TRACK_LOST; points never come back, so later frames are NaN too.
__global__ void motion_reduce_kernel(const MeasureParams p) {
  const int clip = blockIdx.y;
  int rx, ry, rw, rh;
  const int st = p.status[clip];
  const bool bad = st == RM_CLIP_NO_ROI || st == RM_CLIP_NO_CORNERS || p.npts[clip] <= 0 ||
                   !(roi_ok_geom(p, clip, rx, ry, rw, rh));
  float* motion = p.motion + (long long)clip * p.n_frames * 2;
  const int n0 = p.npts[clip];
  for (int f = p.f0 + blockIdx.x * blockDim.x + threadIdx.x; f < p.f1; f += gridDim.x * blockDim.x) {
    float mx = NAN, my = NAN;
    if (!bad) {
      if (f == 0) { mx = 0.f; my = 0.f; }
      else {
        const float2* d = p.delta + ((long long)clip * p.delta_frames + (f - p.delta_f0)) * LK_MAX_PTS;
        int m = 0;
        float sx = 0.f, sy = 0.f;
        for (int i = 0; i < n0; ++i) {
          const float2 v = d[i];
          if (v.x == v.x) { sx += v.x; sy += v.y; ++m; }
        }
        if (m > 0) { mx = sx / (float)m; my = sy / (float)m; }
        else p.status[clip] = RM_CLIP_TRACK_LOST;
      }
    }
    motion[2 * f] = mx;
    motion[2 * f + 1] = my;
  }
}

// data[f] (base.py:396-407): 0.0 for the first two frames, then the PCA projection of the rolling motion history.
__global__ void motion_pca_kernel(const MeasureParams p) {
  const int clip = blockIdx.y;
  const float* motion = p.motion + (long long)clip * p.n_frames * 2;
  double* data = p.data + (long long)clip * p.n_frames;
  for (int f = p.f0 + blockIdx.x * blockDim.x + threadIdx.x; f < p.f1; f += gridDim.x * blockDim.x) {
    double v;
    if (p.status[clip] == RM_CLIP_NO_ROI || p.status[clip] == RM_CLIP_NO_CORNERS) v = NAN;
    else if (f < 2) v = 0.0;
    else if (isnan(motion[2 * f])) v = NAN;
    else {
      // motion_data holds m_1..m_f, rolled to the last buf_len entries (base.py:473-475)
      const int n = f < p.buf_len ? f : p.buf_len;
      v = sc_pca_project_last(motion + 2 * (f - n + 1), n);
    }
    data[f] = v;
  }
}

// extract_motion 'average' (base.py:355-358): np.average of the float64 crop = pairwise sum of gray*(1/255) / count.
__global__ void measure_average_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ roi, int T, int W,
                                       int H, int first_frame, int n_frames, double* __restrict__ data) {
  const int clip = blockIdx.y, f = blockIdx.x;
  const int rx = roi[clip * 4], ry = roi[clip * 4 + 1], rw = roi[clip * 4 + 2], rh = roi[clip * 4 + 3];
  __shared__ double red[8];
  const uint8_t* img = frames + ((long long)clip * T + first_frame + f) * W * H + (long long)ry * W + rx;
  double acc = 0.0;
  const double inv = 1.0 / 255;
  for (int i = threadIdx.x; i < rw * rh; i += blockDim.x) acc += (double)img[(long long)(i / rw) * W + (i % rw)] * inv;
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) acc += red[w];
    data[(long long)clip * n_frames + f] = acc / (double)(rw * rh);
  }
}

// ------------------------------------------------------------------------------------------------ host side
struct MeasureLayout {
  size_t cov, eig, eigmax, cand, pts0, delta, pyr[LK_MAX_LEVELS], total;
  long long lvl_elems[LK_MAX_LEVELS];
};
static MeasureLayout measure_layout(const rm_handle* h, int maxw, int maxh, int n_clips, int n_frames) {
  MeasureLayout L;
  memset(&L, 0, sizeof(L));
  size_t off = 256;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t plane = (size_t)maxw * maxh;
  L.cov = take((size_t)n_clips * 3 * plane * 4);
  L.eig = take((size_t)n_clips * plane * 4);
  L.eigmax = take((size_t)n_clips * 4);
  L.cand = take((size_t)n_clips * plane * 8);
  L.pts0 = take((size_t)n_clips * LK_MAX_PTS * 2 * 4);
  L.delta = take((size_t)n_clips * n_frames * LK_MAX_PTS * sizeof(float2));
  int w = maxw, hh = maxh;
  for (int l = 1; l < LK_MAX_LEVELS && l <= h->p.lk_max_level; ++l) {
    w = (w + 1) / 2; hh = (hh + 1) / 2;
    L.lvl_elems[l] = (long long)w * hh;
    L.pyr[l] = take((size_t)n_clips * n_frames * w * hh);
  }
  L.total = off;
  return L;
}

extern "C" int32_t rm_measure_workspace_bytes(rm_handle* h, int32_t max_roi_w, int32_t max_roi_h, int32_t n_clips,
                                              int32_t n_frames, size_t* out) {
  RM_CHECK_ARG(h, h && out && max_roi_w >= 1 && max_roi_h >= 1 && n_clips >= 0 && n_frames >= 0, "bad size");
  *out = measure_layout(h, max_roi_w, max_roi_h, n_clips, n_frames).total;
  return RM_OK;
}

struct MeasureJob {
  MeasureParams p;
  MeasureLayout L;
  LkSmemLayout SL;
  float* pts0;
  bool smem_path;
};

static int32_t measure_setup(rm_handle* h, const uint8_t* frames, int32_t n_clips, int32_t T, int32_t W, int32_t H,
                             const int32_t* roi, int32_t max_roi_w, int32_t max_roi_h, int32_t first_frame,
                             int32_t n_frames, double* data_out, float* motion_out, int32_t* npts_out,
                             int32_t* status_io, float* pts_dbg, void* workspace, size_t workspace_bytes,
                             MeasureJob* job, int ring = 0, int layout_frames = -1) {
  // ring > 0 (streaming): `frames` holds the last `ring` = T frames of every clip, n_frames is the capacity of the
  // per-frame output arrays and layout_frames the number of frames one call processes (sizes the workspace)
  if (layout_frames < 0) layout_frames = n_frames;
  RM_CHECK_ARG(h, h && frames && roi && data_out && motion_out && npts_out && status_io, "null pointer");
  RM_CHECK_ARG(h, n_clips >= 0 && T >= 1 && W >= 1 && H >= 1 && first_frame >= 0 && n_frames >= 1 &&
                      (ring > 0 || first_frame + n_frames <= T), "frame range outside the clip");
  RM_CHECK_ARG(h, max_roi_w >= 1 && max_roi_h >= 1 && max_roi_w <= W && max_roi_h <= H, "bad max ROI size");
  if (h->p.max_corners > LK_MAX_PTS || h->p.max_corners < 1 || h->p.lk_win < 3 || h->p.lk_win > 31 ||
      h->p.lk_win * h->p.lk_win > 256 || h->p.lk_max_level < 0 || h->p.lk_max_level >= LK_MAX_LEVELS)
    return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: needs max_corners <= 128, 3 <= lk_win <= 16, lk_max_level <= 3", __func__);
  MeasureLayout L = measure_layout(h, max_roi_w, max_roi_h, n_clips, layout_frames);
  if (n_clips > 0 && (!workspace || workspace_bytes < L.total))
    return rm_fail(h, RM_ERR_WORKSPACE, "%s: workspace too small (%lld needed, %lld given)", __func__, (long long)L.total,
                   (long long)workspace_bytes);
  unsigned char* ws = reinterpret_cast<unsigned char*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  MeasureParams& p = job->p;
  memset(&p, 0, sizeof(p));
  p.frames = frames; p.roi = roi; p.lut = h->d_lut;
  p.n_clips = n_clips; p.T = T; p.W = W; p.H = H; p.first_frame = first_frame; p.n_frames = n_frames;
  p.maxw = max_roi_w; p.maxh = max_roi_h;
  p.max_corners = h->p.max_corners; p.min_distance = h->p.min_distance; p.block_size = h->p.block_size;
  p.quality = (float)h->p.quality_level;
  p.win = h->p.lk_win; p.max_level = h->p.lk_max_level;
  p.max_iter = h->p.lk_max_iter < 0 ? 0 : (h->p.lk_max_iter > 100 ? 100 : h->p.lk_max_iter);
  double eps = h->p.lk_eps < 0 ? 0 : (h->p.lk_eps > 10 ? 10 : h->p.lk_eps);
  p.eps2 = eps * eps;
  p.min_eig = (float)h->p.lk_min_eig;
  p.buf_len = h->p.measure_buffer_len;
  p.cov = reinterpret_cast<float*>(ws + L.cov);
  p.eig = reinterpret_cast<float*>(ws + L.eig);
  p.eigmax = reinterpret_cast<unsigned*>(ws + L.eigmax);
  p.cand = reinterpret_cast<unsigned long long*>(ws + L.cand);
  job->pts0 = reinterpret_cast<float*>(ws + L.pts0);
  for (int l = 1; l < LK_MAX_LEVELS; ++l) {
    p.pyr[l] = L.lvl_elems[l] ? ws + L.pyr[l] : nullptr;
    p.lvl_elems[l] = L.lvl_elems[l];
  }
  p.motion = motion_out; p.data = data_out; p.npts = npts_out; p.status = status_io; p.pts_dbg = pts_dbg;
  p.f0 = 0; p.f1 = n_frames;
  p.delta = reinterpret_cast<float2*>(ws + L.delta);
  p.delta_f0 = 0; p.delta_frames = layout_frames;
  p.ring = ring;
  if (pts_dbg) { p.bpc = 1; p.ppb = LK_MAX_PTS; }          // the diagnostic dump wants one compacted list per clip
  else {
    p.ppb = LK_PPB;                                        // at most one corner per warp, half of the warps per block
    p.bpc = (h->p.max_corners + p.ppb - 1) / p.ppb;
    if (p.bpc > LK_MAX_BLOCKS) { p.bpc = LK_MAX_BLOCKS; p.ppb = (h->p.max_corners + p.bpc - 1) / p.bpc; }
  }
  job->L = L;
  job->SL = lk_smem_layout(max_roi_w, max_roi_h, p.win, p.max_level, LKS_WARPS);
  job->smem_path = job->SL.total + 4096 <= h->smem_optin && !h->force_global_lk;
  return RM_OK;
}

// Shi-Tomasi corners of the first measure frame (base.py:365-366)
static int32_t measure_gftt(rm_handle* h, MeasureJob* job, cudaStream_t st) {
  const MeasureParams& p = job->p;
  RM_CUDA(h, cudaMemsetAsync(p.eigmax, 0, (size_t)p.n_clips * 4, st));
  const int roi_px = p.maxw * p.maxh;
  dim3 g1(div_up(roi_px, 256) < 64 ? div_up(roi_px, 256) : 64, p.n_clips);
  RM_PROF(h, st, "gftt_cov_kernel");
  gftt_cov_kernel<<<g1, 256, 0, st>>>(p);
  RM_LAUNCH_CHECK(h);
  RM_PROF(h, st, "gftt_eig_kernel");
  gftt_eig_kernel<<<g1, 256, 0, st>>>(p);
  RM_LAUNCH_CHECK(h);
  RM_PROF(h, st, "gftt_select_kernel");
  gftt_select_kernel<<<p.n_clips, 256, 0, st>>>(p, job->pts0);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

// LK over frames [f0, f1).  The shared-memory tracker can stop and resume at any frame (state in h->d_lk_*); the
// global-memory fallback for very large ROIs runs the whole range at once (call it with f0 == 0 only).
static int32_t measure_lk(rm_handle* h, MeasureJob* job, int f0, int f1, bool carry_state, cudaStream_t st) {
  MeasureParams p = job->p;
  p.f0 = f0; p.f1 = f1;
  if (job->smem_path) {
    p.st_pts = carry_state ? h->d_lk_pts : nullptr;
    p.st_idx = carry_state ? h->d_lk_idx : nullptr;
    p.st_n = carry_state ? h->d_lk_n : nullptr;
    if (f0 == p.delta_f0)
      RM_CUDA(h, cudaMemsetAsync(p.delta, 0xFF, (size_t)p.n_clips * p.delta_frames * LK_MAX_PTS * sizeof(float2), st));
    // production path: crops and their pyramids live in shared memory
    RM_CUDA(h, cudaFuncSetAttribute(lk_track_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, job->SL.total));
    RM_PROF(h, st, "lk_track_smem_kernel");
    lk_track_smem_kernel<<<dim3(p.n_clips, p.bpc), LKS_WARPS * 32, job->SL.total, st>>>(p, job->pts0, job->SL.total);
    RM_LAUNCH_CHECK(h);
  } else {
    if (f0 != 0) return RM_OK;          // done by the first call
    p.f1 = p.n_frames;
    // ROI too large for shared memory: pyramids of every frame in the workspace, pixels fetched from global memory
    for (int l = 1; l < LK_MAX_LEVELS && l <= h->p.lk_max_level; ++l) {
      dim3 g2(div_up(job->L.lvl_elems[l], 256) < 32 ? div_up(job->L.lvl_elems[l], 256) : 32, p.n_frames, p.n_clips);
      RM_PROF(h, st, "lk_pyr_kernel");
      lk_pyr_kernel<<<g2, 256, 0, st>>>(p, l);
      RM_LAUNCH_CHECK(h);
    }
    const int pw = p.win + 3, dwid = p.win + 1;
    const size_t per_warp = (size_t)(((pw * pw + 1) & ~1) + 2 * dwid * dwid) * sizeof(short);
    RM_PROF(h, st, "lk_track_kernel");
    lk_track_kernel<<<p.n_clips, LK_WARPS * 32, per_warp * LK_WARPS, st>>>(p, job->pts0);
    RM_LAUNCH_CHECK(h);
  }
  return RM_OK;
}

static int32_t measure_pca(rm_handle* h, MeasureJob* job, int f0, int f1, cudaStream_t st) {
  MeasureParams p = job->p;
  p.f0 = f0; p.f1 = f1;
  if (job->smem_path) {   // the blocks of a clip left per-point displacements: fold them into motion[f0..f1)
    dim3 g0(div_up(f1 - f0, 64), p.n_clips);
    RM_PROF(h, st, "motion_reduce_kernel");
    motion_reduce_kernel<<<g0, 64, 0, st>>>(p);
    RM_LAUNCH_CHECK(h);
  }
  dim3 g3(div_up(f1 - f0, 128), p.n_clips);
  RM_PROF(h, st, "motion_pca_kernel");
  motion_pca_kernel<<<g3, 128, 0, st>>>(p);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

static int32_t measure_flow_impl(rm_handle* h, const uint8_t* frames, int32_t n_clips, int32_t T, int32_t W, int32_t H,
                                 const int32_t* roi, int32_t max_roi_w, int32_t max_roi_h, int32_t first_frame,
                                 int32_t n_frames, double* data_out, float* motion_out, int32_t* npts_out,
                                 int32_t* status_io, float* pts_dbg, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  MeasureJob job;
  int32_t rc = measure_setup(h, frames, n_clips, T, W, H, roi, max_roi_w, max_roi_h, first_frame, n_frames, data_out,
                             motion_out, npts_out, status_io, pts_dbg, workspace, workspace_bytes, &job);
  if (rc != RM_OK || n_clips == 0) return rc;
  DeviceGuard dg(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = measure_gftt(h, &job, st)) != RM_OK) return rc;
  if ((rc = measure_lk(h, &job, 0, n_frames, false, st)) != RM_OK) return rc;
  return measure_pca(h, &job, 0, n_frames, st);
}

// extract_motion + measure() for whole clips as one pipeline (base.py:464-495): rm_measure_flow followed by
// rm_signal_bpm, except that the tracker walks the frames in `measure_chunks` chunks on the caller's stream and the
// signal stage of every finished chunk runs on streams owned by the handle underneath the next chunk's tracking: PCA,
// filtfilt and peak picking in frame order on one stream, the Gaussian-fit gate + BPM of each chunk on its own stream
// (a handful of fits run ten times longer than the rest; their tails must not queue behind each other).  Both stages are latency bound (sequential frames; sequential LM
// iterations), so overlapping them is what shortens the step.  Results are identical to the two separate calls.
extern "C" int32_t rm_measure_signal(rm_handle* h, const uint8_t* frames, int32_t n_clips, int32_t T, int32_t W, int32_t H,
                                     const int32_t* roi, int32_t max_roi_w, int32_t max_roi_h, int32_t first_frame,
                                     int32_t n_frames, double fps, double* data_out, float* motion_out,
                                     int32_t* npts_out, int32_t* status_io, double* bpm_out, double* filtered_out,
                                     int32_t* peaks_out, int32_t* npeaks_out, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  RM_CHECK_ARG(h, h && bpm_out && fps > 0, "null pointer or bad fps");
  MeasureJob job;
  int32_t rc = measure_setup(h, frames, n_clips, T, W, H, roi, max_roi_w, max_roi_h, first_frame, n_frames, data_out,
                             motion_out, npts_out, status_io, nullptr, workspace, workspace_bytes, &job);
  if (rc != RM_OK || n_clips == 0) return rc;
  DeviceGuard dg(h->device);
  cudaStream_t sa = (cudaStream_t)stream, sb = h->aux_stream;
  int n_chunks = job.smem_path ? h->measure_chunks : 1;
  if (n_chunks > n_frames) n_chunks = n_frames;
  if (n_chunks > RM_MAX_CHUNKS) n_chunks = RM_MAX_CHUNKS;
  if (h->lk_state_cap < n_clips) {
    if (h->d_lk_pts) cudaFree(h->d_lk_pts);
    if (h->d_lk_idx) cudaFree(h->d_lk_idx);
    if (h->d_lk_n) cudaFree(h->d_lk_n);
    h->d_lk_pts = nullptr; h->d_lk_idx = nullptr; h->d_lk_n = nullptr; h->lk_state_cap = 0;
    RM_CUDA(h, cudaMalloc((void**)&h->d_lk_pts, (size_t)n_clips * LK_MAX_PTS * 2 * sizeof(float)));
    RM_CUDA(h, cudaMalloc((void**)&h->d_lk_idx, (size_t)n_clips * LK_MAX_PTS * sizeof(int)));
    RM_CUDA(h, cudaMalloc((void**)&h->d_lk_n, (size_t)n_clips * LK_MAX_BLOCKS * sizeof(int)));
    h->lk_state_cap = n_clips;
  }
  // everything that allocates happens before the first launch
  if ((rc = rmi_signal_setup(h, data_out, n_clips, n_frames, fps, bpm_out, filtered_out, peaks_out, npeaks_out, status_io,
                             n_chunks, sa, 0, 0)) != RM_OK)
    return rc;
  if ((rc = measure_gftt(h, &job, sa)) != RM_OK) return rc;
  RM_CUDA(h, cudaEventRecord(h->ev_fork, sa));
  RM_CUDA(h, cudaStreamWaitEvent(sb, h->ev_fork, 0));
  // Chunk boundaries: a short first chunk (the first measure() windows -- the shortest data, the most degenerate and
  // therefore longest Gaussian fits -- start as early as possible), the rest split evenly.
  int bounds[RM_MAX_CHUNKS + 1];
  bounds[0] = 0;
  if (n_chunks == 1) bounds[1] = n_frames;
  else {
    int first = h->p.measure_init_len + 8;
    if (first > n_frames / n_chunks) first = n_frames / n_chunks;
    if (first < 1) first = 1;
    const int mid_chunks = n_chunks - 1, mid_frames = n_frames - first;
    bounds[1] = first;
    for (int c = 2; c <= 1 + mid_chunks; ++c) bounds[c] = first + (int)((long long)mid_frames * (c - 1) / mid_chunks);
  }
  for (int c = 0; c < n_chunks; ++c) {
    const int f0 = bounds[c], f1 = bounds[c + 1];
    if ((rc = measure_lk(h, &job, f0, f1, n_chunks > 1, sa)) != RM_OK) return rc;
    RM_CUDA(h, cudaEventRecord(h->ev_chunk[c], sa));
    RM_CUDA(h, cudaStreamWaitEvent(sb, h->ev_chunk[c], 0));
    if ((rc = measure_pca(h, &job, f0, f1, sb)) != RM_OK) return rc;
    if ((rc = rmi_signal_range(h, f0, f1, c, sb, h->fit_stream[c], h->ev_filt[c])) != RM_OK) return rc;
    RM_CUDA(h, cudaEventRecord(h->ev_done[c], h->fit_stream[c]));
  }
  for (int c = 0; c < n_chunks; ++c) RM_CUDA(h, cudaStreamWaitEvent(sa, h->ev_done[c], 0));   // join
  return RM_OK;
}

// The measure branch for LIVE streams (base.py:464-495 frame by frame): a cohort of n_clips cameras whose measure states
// started on the same frame.  `frames` is a ring of ring_len ROI crops per camera (crop of absolute measure frame f in
// slot f % ring_len, written by rm_crop_to_ring); every call tracks the new frames [f_begin, f_end) from the tracker
// state the handle carries from the previous call (f_begin = 0: corners are detected on frame 0), appends their motion,
// data and BPM at their absolute positions in the (n_clips, cap) arrays and leaves filtered / peaks of the window that
// ends at f_end - 1.  Results equal those of rm_measure_signal on the whole clip, whatever the block sizes (tested).
// One handle per cohort; ring_len > f_end - f_begin; the shared-memory tracker only (ROI up to about 100 x 100).
extern "C" int32_t rm_measure_signal_stream(rm_handle* h, const uint8_t* frames, int32_t n_clips, int32_t ring_len,
                                            int32_t W, int32_t H, const int32_t* roi, int32_t max_roi_w,
                                            int32_t max_roi_h, int32_t cap, int32_t f_begin, int32_t f_end, double fps,
                                            double* data_out, float* motion_out, int32_t* npts_out, int32_t* status_io,
                                            double* bpm_out, double* filtered_out, int32_t* peaks_out,
                                            int32_t* npeaks_out, void* workspace, size_t workspace_bytes, void* stream) {
  RM_CHECK_ARG(h, h && bpm_out && fps > 0, "null pointer or bad fps");
  RM_CHECK_ARG(h, f_begin >= 0 && f_end > f_begin && f_end <= cap && f_end - f_begin < ring_len, "bad frame block");
  MeasureJob job;
  const int k = f_end - f_begin;
  int32_t rc = measure_setup(h, frames, n_clips, ring_len, W, H, roi, max_roi_w, max_roi_h, 0, cap, data_out, motion_out,
                             npts_out, status_io, nullptr, workspace, workspace_bytes, &job, ring_len, k);
  if (rc != RM_OK || n_clips == 0) return rc;
  if (!job.smem_path) return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: ROI too large for the shared-memory tracker", __func__);
  DeviceGuard dg(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->lk_state_cap < n_clips) {
    if (f_begin != 0) return rm_fail(h, RM_ERR_INVALID, "%s: no tracker state for this cohort (start with f_begin = 0)", __func__);
    if (h->d_lk_pts) cudaFree(h->d_lk_pts);
    if (h->d_lk_idx) cudaFree(h->d_lk_idx);
    if (h->d_lk_n) cudaFree(h->d_lk_n);
    h->d_lk_pts = nullptr; h->d_lk_idx = nullptr; h->d_lk_n = nullptr; h->lk_state_cap = 0;
    RM_CUDA(h, cudaMalloc((void**)&h->d_lk_pts, (size_t)n_clips * LK_MAX_PTS * 2 * sizeof(float)));
    RM_CUDA(h, cudaMalloc((void**)&h->d_lk_idx, (size_t)n_clips * LK_MAX_PTS * sizeof(int)));
    RM_CUDA(h, cudaMalloc((void**)&h->d_lk_n, (size_t)n_clips * LK_MAX_BLOCKS * sizeof(int)));
    h->lk_state_cap = n_clips;
  }
  if ((rc = rmi_signal_setup(h, data_out, n_clips, cap, fps, bpm_out, filtered_out, peaks_out, npeaks_out, status_io, 1, st,
                             f_begin, k)) != RM_OK)
    return rc;
  job.p.delta_f0 = f_begin;
  job.p.delta_frames = k;
  if (f_begin == 0 && (rc = measure_gftt(h, &job, st)) != RM_OK) return rc;
  if ((rc = measure_lk(h, &job, f_begin, f_end, true, st)) != RM_OK) return rc;
  if ((rc = measure_pca(h, &job, f_begin, f_end, st)) != RM_OK) return rc;
  return rmi_signal_range(h, f_begin, f_end, 0, st, st, nullptr);
}

// 'average' extraction for live cohorts: the mean of the float crop (base.py:355-358) of the new frames [f_begin, f_end)
// read from the crop ring, written at their absolute positions of the (n_clips, cap) history -- the same sums in the same
// order as measure_average_kernel on the whole clip.
__global__ void measure_average_ring_kernel(const uint8_t* __restrict__ ring, const int32_t* __restrict__ roi, int ring_len,
                                            int W, int H, int cap, int f_begin, double* __restrict__ data) {
  const int clip = blockIdx.y, f = f_begin + blockIdx.x;
  const int rw = roi[clip * 4 + 2], rh = roi[clip * 4 + 3];
  __shared__ double red[8];
  const uint8_t* img = ring + ((long long)clip * ring_len + f % ring_len) * W * H;
  double acc = 0.0;
  const double inv = 1.0 / 255;
  for (int i = threadIdx.x; i < rw * rh; i += blockDim.x) acc += (double)img[(long long)(i / rw) * W + (i % rw)] * inv;
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) acc += red[w];
    data[(long long)clip * cap + f] = acc / (double)(rw * rh);
  }
}

// rm_measure_signal_stream for motion_extraction_method = 'average': no tracker, no state between calls.
extern "C" int32_t rm_measure_average_stream(rm_handle* h, const uint8_t* ring, int32_t n_clips, int32_t ring_len, int32_t W,
                                             int32_t H, const int32_t* roi, int32_t cap, int32_t f_begin, int32_t f_end,
                                             double fps, double* data_out, int32_t* status_io, double* bpm_out,
                                             double* filtered_out, int32_t* peaks_out, int32_t* npeaks_out, void* stream) {
  RM_CHECK_ARG(h, h && ring && roi && data_out && bpm_out && fps > 0, "null pointer or bad fps");
  RM_CHECK_ARG(h, f_begin >= 0 && f_end > f_begin && f_end <= cap && f_end - f_begin < ring_len, "bad frame block");
  if (n_clips == 0) return RM_OK;
  DeviceGuard dg(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  int32_t rc;
  if ((rc = rmi_signal_setup(h, data_out, n_clips, cap, fps, bpm_out, filtered_out, peaks_out, npeaks_out, status_io, 1, st,
                             f_begin, f_end - f_begin)) != RM_OK)
    return rc;
  dim3 grid(f_end - f_begin, n_clips);
  RM_PROF(h, st, "measure_average_ring_kernel");
  measure_average_ring_kernel<<<grid, 256, 0, st>>>(ring, roi, ring_len, W, H, cap, f_begin, data_out);
  RM_LAUNCH_CHECK(h);
  return rmi_signal_range(h, f_begin, f_end, 0, st, st, nullptr);
}

// Copies the ROI crop of k new frames of every camera into its crop ring: frames (n_clips, k, H, W), roi (n_clips, 4) in
// frame coordinates, ring (n_clips, ring_len, ring_h, ring_w); frame j of the block goes to slot (f_first + j) % ring_len.
__global__ void crop_to_ring_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ roi,
                                    uint8_t* __restrict__ ring, int k, int W, int H, int ring_len, int ring_w, int ring_h,
                                    int f_first) {
  const int clip = blockIdx.y, j = blockIdx.x;
  const int x = roi[clip * 4], y = roi[clip * 4 + 1], w = roi[clip * 4 + 2], hh = roi[clip * 4 + 3];
  if (w < 1 || hh < 1 || x < 0 || y < 0 || x + w > W || y + hh > H || w > ring_w || hh > ring_h) return;
  const uint8_t* src = frames + ((long long)clip * k + j) * W * H + (long long)y * W + x;
  uint8_t* dst = ring + ((long long)clip * ring_len + (f_first + j) % ring_len) * ring_w * ring_h;
  for (int i = threadIdx.x; i < w * hh; i += blockDim.x) {
    const int r = i / w, c = i - r * w;
    dst[r * ring_w + c] = src[(long long)r * W + c];
  }
}
// ROI crops of a run of frames as one contiguous gray tensor: frames (n_clips, T, H, W[, 3]) of dtype RM_U8 / RM_BGR8,
// roi (n_clips, 4) -> out (n_clips, n_frames, out_h, out_w) holding frame[y:y+h, x:x+w] (base.py:471) top-left aligned.
// For BGR frames cv2.cvtColor's fixed point (next_frame, base.py:230) is applied to the ROI's pixels only -- the measure
// stage never needs the rest of the frame in gray.
// Gray rows are read one row per warp as aligned 32-bit words (one request of up to 128 contiguous bytes per row): the
// frames may lie in pinned HOST memory mapped into the device's address space -- then the crop is the upload, and only the
// ROI's rows cross PCIe -- where byte-sized requests would cost a bus transaction each.  `end` = one past the clip's last
// byte: a row's last word is assembled from bytes if it would reach past it.
__device__ __forceinline__ void crop_rows_gray(const uint8_t* __restrict__ src, long long row_stride, const uint8_t* end,
                                               uint8_t* __restrict__ dst, int w, int hh, int out_w) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < hh; r += nwarps) {
    const uint8_t* row = src + (long long)r * row_stride;
    const int mis = (int)((uintptr_t)row & 3);
    const uint8_t* w0 = row - mis;
    uint8_t* drow = dst + r * out_w;
    const int nwords = (mis + w + 3) >> 2;
    for (int k = lane; k < nwords; k += 32) {
      const uint8_t* wp = w0 + 4 * k;
      unsigned v;
      if (wp + 4 <= end) v = *reinterpret_cast<const unsigned*>(wp);
      else {
        v = 0;
        for (int b = 0; b < 4; ++b)
          if (wp + b < end) v |= (unsigned)wp[b] << (8 * b);
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int c = 4 * k - mis + b;
        if (c >= 0 && c < w) drow[c] = (uint8_t)(v >> (8 * b));
      }
    }
  }
}

template <bool BGR>
__global__ void crop_frames_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ roi,
                                   uint8_t* __restrict__ out, int T, int first, int n_frames, int W, int H, int out_w,
                                   int out_h) {
  const int clip = blockIdx.y, j = blockIdx.x;
  const int x = roi[clip * 4], y = roi[clip * 4 + 1], w = roi[clip * 4 + 2], hh = roi[clip * 4 + 3];
  if (w < 1 || hh < 1 || x < 0 || y < 0 || x + w > W || y + hh > H || w > out_w || hh > out_h) return;
  const int PX = BGR ? 3 : 1;
  const uint8_t* src = frames + (((long long)clip * T + first + j) * W * H + (long long)y * W + x) * PX;
  uint8_t* dst = out + ((long long)clip * n_frames + j) * out_w * out_h;
  if (!BGR) {
    crop_rows_gray(src, W, frames + (long long)(clip + 1) * T * W * H, dst, w, hh, out_w);
    return;
  }
  for (int i = threadIdx.x; i < w * hh; i += blockDim.x) {
    const int r = i / w, c = i - r * w;
    const uint8_t* px = src + ((long long)r * W + c) * PX;
    dst[r * out_w + c] = (uint8_t)((3735u * px[0] + 19235u * px[1] + 9798u * px[2] + (1u << 14)) >> 15);
  }
}
extern "C" int32_t rm_crop_frames(rm_handle* h, const void* frames, int32_t dtype, int32_t n_clips, int32_t T, int32_t W,
                                  int32_t H, const int32_t* roi, int32_t first_frame, int32_t n_frames, uint8_t* out,
                                  int32_t out_w, int32_t out_h, void* stream) {
  RM_CHECK_ARG(h, h && frames && roi && out && n_clips >= 0 && first_frame >= 0 && n_frames >= 1 &&
                      first_frame + n_frames <= T && out_w >= 1 && out_h >= 1, "bad argument");
  RM_CHECK_ARG(h, dtype == RM_U8 || dtype == RM_BGR8, "frames must be RM_U8 or RM_BGR8");
  if (n_clips == 0) return RM_OK;
  DeviceGuard dg(h->device);
  RM_PROF(h, (cudaStream_t)stream, "crop_frames_kernel");
  if (dtype == RM_BGR8)
    crop_frames_kernel<true><<<dim3(n_frames, n_clips), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)frames, roi, out, T,
                                                                                        first_frame, n_frames, W, H, out_w, out_h);
  else
    crop_frames_kernel<false><<<dim3(n_frames, n_clips), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)frames, roi, out, T,
                                                                                         first_frame, n_frames, W, H, out_w, out_h);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

// The same for a RAGGED batch (BASELINE config 5: clips of different resolutions in one batch): every clip is described by
// an rm_clip_desc -- where its first frame lies, its frame size, its length, its row pitch -- and one launch crops the ROIs
// of all clips, whatever their resolution class, into one (n_clips, n_frames, out_h, out_w) tensor.  From there on the
// measure stage is a single batch: one tracker / PCA / filter / fit pipeline over all classes.
template <bool BGR>
__global__ void crop_frames_ragged_kernel(const uint8_t* __restrict__ base, const rm_clip_desc* __restrict__ descs,
                                          const int32_t* __restrict__ roi, uint8_t* __restrict__ out, int first,
                                          int n_frames, int out_w, int out_h) {
  const int clip = blockIdx.y, j = blockIdx.x;
  const rm_clip_desc d = descs[clip];
  const int x = roi[clip * 4], y = roi[clip * 4 + 1], w = roi[clip * 4 + 2], hh = roi[clip * 4 + 3];
  if (w < 1 || hh < 1 || x < 0 || y < 0 || x + w > d.W || y + hh > d.H || w > out_w || hh > out_h || first + j >= d.T) return;
  const int PX = BGR ? 3 : 1;
  const uint8_t* src = base + d.frame_offset + ((long long)(first + j) * d.H + y) * d.row_stride + (long long)x * PX;
  uint8_t* dst = out + ((long long)clip * n_frames + j) * out_w * out_h;
  if (!BGR) {
    crop_rows_gray(src, d.row_stride, base + d.frame_offset + (long long)d.T * d.H * d.row_stride, dst, w, hh, out_w);
    return;
  }
  for (int i = threadIdx.x; i < w * hh; i += blockDim.x) {
    const int r = i / w, c = i - r * w;
    const uint8_t* px = src + (long long)r * d.row_stride + c * PX;
    dst[r * out_w + c] = (uint8_t)((3735u * px[0] + 19235u * px[1] + 9798u * px[2] + (1u << 14)) >> 15);
  }
}
extern "C" int32_t rm_crop_frames_ragged(rm_handle* h, const void* base, int32_t dtype, const rm_clip_desc* descs,
                                         int32_t n_clips, const int32_t* roi, int32_t first_frame, int32_t n_frames,
                                         uint8_t* out, int32_t out_w, int32_t out_h, void* stream) {
  RM_CHECK_ARG(h, h && descs && roi && out && n_clips >= 0 && first_frame >= 0 && n_frames >= 1 && out_w >= 1 && out_h >= 1,
               "bad argument");
  RM_CHECK_ARG(h, dtype == RM_U8 || dtype == RM_BGR8, "frames must be RM_U8 or RM_BGR8");
  if (n_clips == 0) return RM_OK;
  DeviceGuard dg(h->device);
  RM_PROF(h, (cudaStream_t)stream, "crop_frames_ragged_kernel");
  if (dtype == RM_BGR8)
    crop_frames_ragged_kernel<true><<<dim3(n_frames, n_clips), 256, 0, (cudaStream_t)stream>>>(
        (const uint8_t*)base, descs, roi, out, first_frame, n_frames, out_w, out_h);
  else
    crop_frames_ragged_kernel<false><<<dim3(n_frames, n_clips), 256, 0, (cudaStream_t)stream>>>(
        (const uint8_t*)base, descs, roi, out, first_frame, n_frames, out_w, out_h);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

extern "C" int32_t rm_crop_to_ring(rm_handle* h, const uint8_t* frames, int32_t n_clips, int32_t k, int32_t W, int32_t H,
                                   const int32_t* roi, uint8_t* ring, int32_t ring_len, int32_t ring_w, int32_t ring_h,
                                   int32_t f_first, void* stream) {
  RM_CHECK_ARG(h, h && frames && roi && ring && n_clips >= 0 && k >= 1 && k <= ring_len && f_first >= 0, "bad argument");
  if (n_clips == 0) return RM_OK;
  DeviceGuard dg(h->device);
  RM_PROF(h, (cudaStream_t)stream, "crop_to_ring_kernel");
  crop_to_ring_kernel<<<dim3(k, n_clips), 256, 0, (cudaStream_t)stream>>>(frames, roi, ring, k, W, H, ring_len, ring_w,
                                                                           ring_h, f_first);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

extern "C" int32_t rm_measure_flow(rm_handle* h, const uint8_t* frames, int32_t n_clips, int32_t T, int32_t W, int32_t H,
                                   const int32_t* roi, int32_t max_roi_w, int32_t max_roi_h, int32_t first_frame,
                                   int32_t n_frames, double* data_out, float* motion_out, int32_t* npts_out,
                                   int32_t* status_io, void* workspace, size_t workspace_bytes, void* stream) {
  return measure_flow_impl(h, frames, n_clips, T, W, H, roi, max_roi_w, max_roi_h, first_frame, n_frames, data_out,
                           motion_out, npts_out, status_io, nullptr, workspace, workspace_bytes, stream);
}

// Same, additionally dumping the tracked points after every frame: pts_out (n_clips, n_frames, 128, 2) float32, NaN padded.
extern "C" int32_t rm_measure_flow_debug(rm_handle* h, const uint8_t* frames, int32_t n_clips, int32_t T, int32_t W,
                                         int32_t H, const int32_t* roi, int32_t max_roi_w, int32_t max_roi_h,
                                         int32_t first_frame, int32_t n_frames, double* data_out, float* motion_out,
                                         int32_t* npts_out, int32_t* status_io, float* pts_out, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  RM_CHECK_ARG(h, pts_out != nullptr, "null pts_out");
  return measure_flow_impl(h, frames, n_clips, T, W, H, roi, max_roi_w, max_roi_h, first_frame, n_frames, data_out,
                           motion_out, npts_out, status_io, pts_out, workspace, workspace_bytes, stream);
}

extern "C" int32_t rm_measure_average(rm_handle* h, const uint8_t* frames, int32_t n_clips, int32_t T, int32_t W,
                                      int32_t H, const int32_t* roi, int32_t first_frame, int32_t n_frames,
                                      double* data_out, void* stream) {
  RM_CHECK_ARG(h, h && frames && roi && data_out, "null pointer");
  RM_CHECK_ARG(h, n_clips >= 0 && first_frame >= 0 && n_frames >= 1 && first_frame + n_frames <= T, "frame range");
  if (n_clips == 0) return RM_OK;
  DeviceGuard dg(h->device);
  dim3 grid(n_frames, n_clips);
  RM_PROF(h, (cudaStream_t)stream, "measure_average_kernel");
  measure_average_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frames, roi, T, W, H, first_frame, n_frames, data_out);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}
