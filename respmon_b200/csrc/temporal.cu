// Temporal band-pass of every pyramid column (transforms.py:82-102).
//
// The reference does  p = fftpack.rfft(x)  (packed real layout), zeroes p[hi:T-hi] and -- if lo != 0 -- p[:lo] and
// p[T-lo:], then takes  real(fftpack.ifft(p))  of that *real packed array* and multiplies by the amplification,
// i.e.  r[t] = amp/T * sum_j p[j] cos(2 pi j t / T)  (SURVEY.md App. A.4).  Both transforms are computed here as a
// length-T complex FFT with zero imaginary input: one warp per column, radix-2 decimation-in-time in shared memory
// for power-of-two T, and an O(T*K) direct evaluation of the kept bins for any other T.
// Layout: (n_clips, T, P) float64, P = packed record length; 8 adjacent columns per block so that every global
// access touches whole 32-byte sectors.
#include "common.cuh"

#define TB_WARPS 8

struct TemporalParams {
  const double* in;
  double* out;
  long long n_clips;
  long long P;     // columns per clip
  int T;
  int logT;        // log2(T) or -1
  int lo, hi;      // bound_low / bound_high (transforms.py:89-90)
  double inv_T;    // 1/T
  double amp;
};

__device__ __forceinline__ bool bin_kept(int j, int T, int lo, int hi) {
  // numpy slice semantics of transforms.py:91-94: p[hi:-hi] (empty when hi == 0), p[:lo], p[-lo:]
  if (hi > 0 && j >= hi && j < T - hi) return false;
  if (lo != 0 && (j < lo || j >= T - lo)) return false;
  return true;
}

__device__ __forceinline__ void warp_fft_pow2(double2* buf, const double2* tw, int T, int logT, int lane) {
  // in-place radix-2 DIT on bit-reversed input.  tw holds the twiddles stage by stage, stage s (half = 2^(s-1)) at
  // tw[half - 1 + k] = exp(-2 pi i k 2^(logT - s) / T), k < half: contiguous per stage, so a warp's reads do not pile up
  // on one bank the way a single strided table does
  for (int s = 1; s <= logT; ++s) {
    const int half = 1 << (s - 1);
    for (int b = lane; b < (T >> 1); b += 32) {
      int grp = b >> (s - 1), k = b & (half - 1);
      int i0 = (grp << s) + k, i1 = i0 + half;
      double2 w = tw[half - 1 + k];
      double2 u = buf[i0], v = buf[i1];
      double2 t = make_double2(v.x * w.x - v.y * w.y, v.x * w.y + v.y * w.x);
      buf[i0] = make_double2(u.x + t.x, u.y + t.y);
      buf[i1] = make_double2(u.x - t.x, u.y - t.y);
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(TB_WARPS * 32) temporal_pow2_kernel(const TemporalParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = p.T;
  const int TS = T + 1;   // column stride (elements): neighbouring columns start 16 B apart -> different banks
  double2* tw = reinterpret_cast<double2*>(smem_raw);                       // T (T - 1 used): twiddles per stage
  double2* bufs = tw + T;                                                   // TB_WARPS * TS
  double* qs = reinterpret_cast<double*>(bufs + (size_t)TB_WARPS * TS);     // TB_WARPS * T
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < T - 1; e += blockDim.x) {
    // entry e = half - 1 + k of stage s: the value exp(-2 pi i (k * T / 2^s) / T) of the single-table layout
    const int s_ = 32 - __clz(e + 1);            // stage (1-based): half = 2^(s_-1) <= e + 1 < 2^s_
    const int half = 1 << (s_ - 1), k = e + 1 - half;
    const int kk = k * (T >> s_);
    double sn, cs;
    sincospi(2.0 * (double)kk / (double)T, &sn, &cs);
    tw[e] = make_double2(cs, -sn);
  }
  const long long groups_per_clip = (p.P + TB_WARPS - 1) / TB_WARPS;
  const long long n_groups = p.n_clips * groups_per_clip;
  const int rev_shift = 32 - p.logT;
  double2* buf = bufs + (size_t)warp * TS;
  double* q = qs + (size_t)warp * T;
  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const long long clip = grp / groups_per_clip;
    const long long c0 = (grp % groups_per_clip) * TB_WARPS;
    const double* src = p.in + clip * T * p.P;
    double* dst = p.out + clip * T * p.P;
    __syncthreads();   // previous group's buffers are free, twiddles are written
    // stage 8 columns x T samples, bit-reversed along T:   thread -> (t = i / 8, col = i % 8)
    for (int i = threadIdx.x; i < T * TB_WARPS; i += blockDim.x) {
      int t = i / TB_WARPS, c = i % TB_WARPS;
      double v = (c0 + c < p.P) ? src[(long long)t * p.P + c0 + c] : 0.0;
      int tr = (p.logT == 0) ? 0 : (int)(__brev((unsigned)t) >> rev_shift);
      bufs[(size_t)c * TS + tr] = make_double2(v, 0.0);
    }
    __syncthreads();
    if (c0 + warp < p.P) {
      warp_fft_pow2(buf, tw, T, p.logT, lane);
      // packed real spectrum (scipy.fftpack.rfft layout), masked
      for (int j = lane; j < T; j += 32) {
        double v;
        if (j == 0) v = buf[0].x;
        else if (j == T - 1 && !(T & 1)) v = buf[T >> 1].x;
        else v = (j & 1) ? buf[(j + 1) >> 1].x : buf[j >> 1].y;
        q[j] = bin_kept(j, T, p.lo, p.hi) ? v : 0.0;
      }
      __syncwarp();
      for (int j = lane; j < T; j += 32) {
        int jr = (p.logT == 0) ? 0 : (int)(__brev((unsigned)j) >> rev_shift);
        buf[jr] = make_double2(q[j], 0.0);
      }
      __syncwarp();
      warp_fft_pow2(buf, tw, T, p.logT, lane);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < T * TB_WARPS; i += blockDim.x) {
      int t = i / TB_WARPS, c = i % TB_WARPS;
      if (c0 + c < p.P) dst[(long long)t * p.P + c0 + c] = bufs[(size_t)c * TS + t].x * p.inv_T * p.amp;
    }
  }
}

// Any T: direct evaluation.  p[j] only for kept j, then r[t] = amp/T * sum_j p[j] cos(2 pi j t / T).
__global__ void __launch_bounds__(TB_WARPS * 32) temporal_direct_kernel(const TemporalParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = p.T;
  double* ct = reinterpret_cast<double*>(smem_raw);   // cos(2 pi m / T)
  double* st = ct + T;                                 // sin(2 pi m / T)
  double* xs = st + T;                                 // TB_WARPS * T
  double* qs = xs + (size_t)TB_WARPS * T;              // TB_WARPS * T
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m = threadIdx.x; m < T; m += blockDim.x) sincospi(2.0 * (double)m / (double)T, &st[m], &ct[m]);
  const long long groups_per_clip = (p.P + TB_WARPS - 1) / TB_WARPS;
  const long long n_groups = p.n_clips * groups_per_clip;
  double* x = xs + (size_t)warp * T;
  double* q = qs + (size_t)warp * T;
  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const long long clip = grp / groups_per_clip;
    const long long c0 = (grp % groups_per_clip) * TB_WARPS;
    const double* src = p.in + clip * T * p.P;
    double* dst = p.out + clip * T * p.P;
    __syncthreads();
    for (int i = threadIdx.x; i < T * TB_WARPS; i += blockDim.x) {
      int t = i / TB_WARPS, c = i % TB_WARPS;
      xs[(size_t)c * T + t] = (c0 + c < p.P) ? src[(long long)t * p.P + c0 + c] : 0.0;
    }
    __syncthreads();
    if (c0 + warp < p.P) {
      for (int j = lane; j < T; j += 32) {
        double acc = 0.0;
        if (bin_kept(j, T, p.lo, p.hi)) {
          // packed index j -> harmonic k and real/imag part
          int k = (j + 1) >> 1;
          bool imag = (j != 0) && !(j & 1);   // even j > 0 holds Im X_k; for even T, j = T-1 is odd -> Re X_{T/2}
          int m = 0;
          for (int t = 0; t < T; ++t) {
            acc += imag ? -x[t] * st[m] : x[t] * ct[m];
            m += k;
            if (m >= T) m -= T;
          }
        }
        q[j] = acc;
      }
      __syncwarp();
      for (int t = lane; t < T; t += 32) {
        double acc = 0.0;
        int m = 0;
        for (int j = 0; j < T; ++j) {
          acc += q[j] * ct[m];
          m += t;
          if (m >= T) m -= T;
        }
        x[t] = acc * p.inv_T * p.amp;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < T * TB_WARPS; i += blockDim.x) {
      int t = i / TB_WARPS, c = i % TB_WARPS;
      if (c0 + c < p.P) dst[(long long)t * p.P + c0 + c] = xs[(size_t)c * T + t];
    }
  }
}

// Sparse form (the default where it applies; option "temporal_sparse" = 0 switches it off): the mask keeps K of the T
// packed bins (24 of 128 at 10 fps), so
//   q[a] = sum_t F[a][t] x[t]        F[a][t] = cos or -sin of harmonic k(j_a) at t        (only the kept bins of the rfft)
//   r[t] = amp/T sum_a G[a][t] q[a]  G[a][t] = cos(2 pi j_a t / T)                         (the real part of the ifft)
// is 2 K T multiply-adds per column instead of two length-T FFTs, with no bit reversal and no per-stage barriers.  Same
// sums in the same order as temporal_direct_kernel (up to the DC offset taken off below); the FFT kernel differs from
// both by rounding.
// A block owns a tile of TS_COLS columns: the tile is staged in shared memory, a thread accumulates 2 bins x 4 columns in
// stage 1 and 8 time samples x 4 columns in stage 2; the coefficient tables are built once per (persistent) block.
#define TS_COLS 32
#define TS_THREADS 128
#define TS_MAXK 32
struct SparseParams {
  TemporalParams b;
  int K;
  int kept[TS_MAXK];   // kept packed indices, increasing
};

__global__ void __launch_bounds__(TS_THREADS) temporal_sparse_kernel(const SparseParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = p.b.T, K = p.K, Tp = T + 2;       // padded table rows: neighbouring bins start 4 banks apart
  double* F = reinterpret_cast<double*>(smem_raw);   // K * Tp
  double* G = F + (size_t)K * Tp;                    // K * Tp
  double* X = G + (size_t)K * Tp;                    // T * TS_COLS
  double* Q = X + (size_t)T * TS_COLS;               // TS_MAXK * TS_COLS
  const int tid = threadIdx.x;
  for (int e = tid; e < K * T; e += TS_THREADS) {
    const int a = e / T, t = e - a * T;
    const int j = p.kept[a];
    const int k = (j + 1) >> 1;
    const bool imag = (j != 0) && !(j & 1);
    const int m1 = (int)(((long long)k * t) % T), m2 = (int)(((long long)j * t) % T);
    double sn, cs, sn2, cs2;
    sincospi(2.0 * (double)m1 / (double)T, &sn, &cs);
    sincospi(2.0 * (double)m2 / (double)T, &sn2, &cs2);
    F[a * Tp + t] = imag ? -sn : cs;
    G[a * Tp + t] = cs2;
  }
  const long long tiles_per_clip = (p.b.P + TS_COLS - 1) / TS_COLS;
  const long long n_tiles = p.b.n_clips * tiles_per_clip;
  const int cg = tid & 7, rg = tid >> 3;          // 4 columns 4cg..4cg+3; bin / time-sample group 0..15
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long clip = tile / tiles_per_clip;
    const long long c0 = (tile % tiles_per_clip) * TS_COLS;
    const double* src = p.b.in + clip * T * p.b.P;
    double* dst = p.b.out + clip * T * p.b.P;
    __syncthreads();                               // tables written; the previous tile's X and Q are free
    // When the mask drops the DC bin the filter does not see a constant offset: take the column's first sample off
    // before the sums.  A column that does not change in time (every static pixel) then gives exact zeros, as the
    // FFT's butterflies do -- locate() depends on it: a static scene must have max == min (base.py:563, :569-570),
    // not the rounding residue of sum_t cos(.) x.
    const bool drop_dc = p.kept[0] != 0;
    const int cc = tid & 31;                         // TS_THREADS is a multiple of 32: a thread stages one column
    const double x_first = (drop_dc && c0 + cc < p.b.P) ? src[c0 + cc] : 0.0;
    for (int i = tid; i < T * TS_COLS; i += TS_THREADS) {
      const int t = i >> 5, c = i & 31;
      X[i] = (c0 + c < p.b.P) ? src[(long long)t * p.b.P + c0 + c] - x_first : 0.0;
    }
    __syncthreads();
    {   // stage 1: bins rg and rg + 16
      const bool v0 = rg < K, v1 = rg + 16 < K;
      const double* f0 = F + (size_t)(v0 ? rg : 0) * Tp;
      const double* f1 = F + (size_t)(v1 ? rg + 16 : 0) * Tp;
      double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
      const double2* x2 = reinterpret_cast<const double2*>(X + 4 * cg);
#pragma unroll 4
      for (int t = 0; t < T; ++t) {
        const double2 xa = x2[t * (TS_COLS / 2)], xb = x2[t * (TS_COLS / 2) + 1];
        const double c0f = f0[t], c1f = f1[t];
        a0[0] += c0f * xa.x; a0[1] += c0f * xa.y; a0[2] += c0f * xb.x; a0[3] += c0f * xb.y;
        a1[0] += c1f * xa.x; a1[1] += c1f * xa.y; a1[2] += c1f * xb.x; a1[3] += c1f * xb.y;
      }
      if (v0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) Q[rg * TS_COLS + 4 * cg + c] = a0[c];
      }
      if (v1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) Q[(rg + 16) * TS_COLS + 4 * cg + c] = a1[c];
      }
    }
    __syncthreads();
    for (int tbase = 0; tbase < T; tbase += 128) {   // stage 2: time samples tbase + rg + 16 i
      double acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.0;
      const double2* q2 = reinterpret_cast<const double2*>(Q + 4 * cg);
      for (int a = 0; a < K; ++a) {
        const double2 qa = q2[a * (TS_COLS / 2)], qb = q2[a * (TS_COLS / 2) + 1];
        const double* g = G + (size_t)a * Tp + tbase + rg;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const double gv = (tbase + rg + 16 * i < T) ? g[16 * i] : 0.0;
          acc[i][0] += qa.x * gv; acc[i][1] += qa.y * gv; acc[i][2] += qb.x * gv; acc[i][3] += qb.y * gv;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int t = tbase + rg + 16 * i;
        if (t < T) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c0 + 4 * cg + c < p.b.P) dst[(long long)t * p.b.P + c0 + 4 * cg + c] = acc[i][c] * p.b.inv_T * p.b.amp;
        }
      }
    }
  }
}

// fftfreq-based bounds, restating transforms.py:88-90 / scipy.fftpack.fftfreq: f[j] = k_j * (1 / (T * d)), d = 1/fps.
extern "C" int32_t rm_temporal_bounds(int32_t T, double fps, double freq_min, double freq_max, int32_t* lo, int32_t* hi) {
  if (T < 1 || !(fps > 0) || !lo || !hi) return RM_ERR_INVALID;
  const double d = 1.0 / fps;
  const double val = 1.0 / ((double)T * d);
  const int npos = (T - 1) / 2 + 1;
  int best_lo = 0, best_hi = 0;
  double dlo = 0, dhi = 0;
  for (int j = 0; j < T; ++j) {
    int k = (j < npos) ? j : j - T;
    double f = (double)k * val;
    double a = fabs(f - freq_min), b = fabs(f - freq_max);
    if (j == 0 || a < dlo) { dlo = a; best_lo = j; }
    if (j == 0 || b < dhi) { dhi = b; best_hi = j; }
  }
  *lo = best_lo;
  *hi = best_hi;
  return RM_OK;
}

extern "C" int32_t rm_temporal_bandpass(rm_handle* h, const double* lap, double* bp_out, int32_t n_clips, int32_t T,
                                        int64_t record_len, double fps, void* stream) {
  RM_CHECK_ARG(h, h && lap && bp_out && n_clips >= 0 && T >= 1 && record_len >= 1 && fps > 0, "null pointer or bad size");
  if (n_clips == 0) return RM_OK;
  DeviceGuard dg(h->device);
  TemporalParams p;
  p.in = lap;
  p.out = bp_out;
  p.n_clips = n_clips;
  p.P = record_len;
  p.T = T;
  p.logT = -1;
  for (int b = 0; b < 31; ++b)
    if ((1 << b) == T) p.logT = b;
  rm_temporal_bounds(T, fps, h->p.freq_min, h->p.freq_max, &p.lo, &p.hi);
  p.inv_T = 1.0 / (double)T;
  p.amp = h->p.amplification;
  const long long n_groups = (long long)n_clips * ((record_len + TB_WARPS - 1) / TB_WARPS);
  cudaStream_t st = (cudaStream_t)stream;
  // 64 clips x 1600 columns, T = 128, K = 24: 0.220 ms against the FFT kernel's 0.374 (r02a); any T that is not a power
  // of two: against temporal_direct_kernel (T = 100: 0.036 ms against 0.196), with which it is bit-identical
  if (h->temporal_sparse && T <= 256) {
    SparseParams sp;
    sp.b = p;
    sp.K = 0;
    bool fits = true;
    for (int j = 0; j < T; ++j) {
      bool kept = true;     // bin_kept(), host side
      if (p.hi > 0 && j >= p.hi && j < T - p.hi) kept = false;
      if (p.lo != 0 && (j < p.lo || j >= T - p.lo)) kept = false;
      if (!kept) continue;
      if (sp.K == TS_MAXK) { fits = false; break; }
      sp.kept[sp.K++] = j;
    }
    const size_t smem = ((size_t)2 * sp.K * (T + 2) + (size_t)T * TS_COLS + (size_t)TS_MAXK * TS_COLS) * sizeof(double);
    if (fits && sp.K >= 1 && (p.logT < 0 || 4 * sp.K <= T) && (int)smem <= h->smem_optin) {
      RM_CUDA(h, cudaFuncSetAttribute(temporal_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int occ = 1;
      RM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, temporal_sparse_kernel, TS_THREADS, smem));
      const long long n_tiles = (long long)n_clips * ((record_len + TS_COLS - 1) / TS_COLS);
      long long grid = (long long)h->sm_count * (occ < 1 ? 1 : occ);
      if (grid > n_tiles) grid = n_tiles;
      RM_PROF(h, st, "temporal_sparse_kernel");
      temporal_sparse_kernel<<<(unsigned)grid, TS_THREADS, smem, st>>>(sp);
      RM_LAUNCH_CHECK(h);
      return RM_OK;
    }
  }
  if (p.logT >= 0 && T <= 2048) {
    size_t smem = (size_t)T * 16 + (size_t)TB_WARPS * (T + 1) * 16 + (size_t)TB_WARPS * T * 8;
    if ((int)smem > h->smem_optin) return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: T too large for shared memory", __func__);
    RM_CUDA(h, cudaFuncSetAttribute(temporal_pow2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    RM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, temporal_pow2_kernel, TB_WARPS * 32, smem));
    long long grid = (long long)h->sm_count * (occ < 1 ? 1 : occ);
    if (grid > n_groups) grid = n_groups;
    RM_PROF(h, st, "temporal_pow2_kernel");
    temporal_pow2_kernel<<<(unsigned)grid, TB_WARPS * 32, smem, st>>>(p);
  } else {
    size_t smem = (size_t)2 * T * 8 + (size_t)2 * TB_WARPS * T * 8;
    if ((int)smem > h->smem_optin) return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: T too large for shared memory", __func__);
    RM_CUDA(h, cudaFuncSetAttribute(temporal_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    RM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, temporal_direct_kernel, TB_WARPS * 32, smem));
    long long grid = (long long)h->sm_count * (occ < 1 ? 1 : occ);
    if (grid > n_groups) grid = n_groups;
    RM_PROF(h, st, "temporal_direct_kernel");
    temporal_direct_kernel<<<(unsigned)grid, TB_WARPS * 32, smem, st>>>(p);
  }
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}
