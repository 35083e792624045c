// Integer front of the Gaussian pyramid for uint8 frames: levels 0 -> 1 -> 2 -> 3 in one pass over the frame.
//
// Reference: the first three cv2.pyrDown calls of create_gaussian_image_pyramid (pyramid.py:9-17) on frames that are
// gray/255 (transforms.py:20-23).  A uint8 frame makes every Gaussian level an exact integer over a power of two:
//   level 1 <= 255*2^8 (u16), level 2 <= 255*2^16, level 3 <= 255*2^24 (u32),
// so the levels are carried as integers -- no rounding at all until the single scale by 2^-32/255 at level 4 -- and
// the 5-tap rows become dp4a / dp2a dot products.
//
// Work decomposition: one WARP streams one vertical strip of one frame top to bottom, 8 pixels per lane per row
// (a 256-pixel-wide strip), with no block-level synchronisation at all.
//   * rows are copied 8 at a time into a warp-private shared-memory ring with cp.async (4 stages, 3 in flight), every
//     lane copies and later reads back only its own 8 bytes;
//   * neighbouring pixels come from the neighbouring lanes by shuffle; two lanes on each interior side of a strip are
//     halo (recomputed by the neighbouring strip), image borders are reflect-101 by byte permutes in the edge lanes;
//   * the vertical 5-tap windows of the three levels are rolling registers; a block of 8 input rows yields 4 level-1
//     rows, 2 level-2 rows and 1 level-3 row, which the payload lanes store (4 B each, coalesced);
//   * the top border is handled by streaming 16 mirrored rows first (reflect-101 about index 0 commutes with the
//     symmetric kernel and the 2:1 decimation); the bottom border (even sizes do not commute) by an explicit flush.
// Fused tail (modes 1 and 2): the level-3 rows go to a per-slot image in shared memory instead of HBM; when the strips of
// a frame are through, the same warps build levels 4..8 and the Laplacians 4..7 (float64, the arithmetic of
// pyramid_tail_kernel operation by operation) in the slot's idle ring and write the packed record.  HBM traffic per
// frame is then SURVEY 8(d)'s figure: W*H bytes read once + the record (1600 doubles at VGA) written once.
// Mode 2 stages the rows with TMA: one cp.async.bulk.tensor.3d (256 bytes x 8 rows, one frame) per stage issued by
// lane 0 and an mbarrier per stage, instead of eight 8-byte cp.async per lane; out-of-frame columns are zero-filled
// by the tensor map.  Mode 0 (level 3 to HBM, pyramid_tail_kernel in pyramid.cu finishes) remains for frames whose
// level images do not fit the shared memory of a slot.
#include <cuda.h>
#include "common.cuh"
#include "pyramid_u8.cuh"

#define PU_STAGES 4
#define PU_ROWS 8
// Developer switch (untimed experiment, default off): the x1 taps of the 5-tap rows as a mask on the ALU pipe instead of a
// dot product with the coefficient vector (…, 0, 1) on the multiplier pipe.  Bit 0: level-1 rows (dp2a), bit 1: level-0
// rows (dp4a).  Integer arithmetic either way, identical results.
#ifndef PU_ALU_TAPS
#define PU_ALU_TAPS 0
#endif
#ifndef PU_MAX_WARPS
#define PU_MAX_WARPS 16     // warps per CTA; each warp owns PU_STAGES * 2 KB of shared memory
#endif
#ifndef PU_BOUND_WARPS
#define PU_BOUND_WARPS PU_MAX_WARPS   // warps the register budget is computed for (launch bounds only): a larger value
#endif                                // leaves registers for another kernel's blocks on the SM (overlapped steps)
#ifndef PU_CTAS_PER_SM
#define PU_CTAS_PER_SM 1    // resident CTAs per SM the grid is sized for (smaller CTAs leave room for the measure stage of
#endif                      // the previous batch when steps overlap: rm_join / "defer_join")
#define PU_STAGE_BYTES (PU_ROWS * 256)

__device__ __forceinline__ void pu_cp_async8(void* smem_dst, const void* gsrc, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void pu_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void pu_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void pu_mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pu_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pu_mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// one box of the (W, H, frames) uint8 tensor: 256 columns x PU_ROWS rows of frame z, top-left (x, y); bytes outside
// the frame arrive as zeros
__device__ __forceinline__ void pu_tma_load(unsigned dst, const CUtensorMap* map, int x, int y, int z, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
      : "memory");
}

// (a + e) + 4 (b + d) + 6 c; also valid lane-wise on two packed 16-bit values when nothing overflows 16 bits
__device__ __forceinline__ unsigned v5(unsigned a, unsigned b, unsigned c, unsigned d, unsigned e) {
  return ((b + d + c) << 2) + (a + e + (c + c));   // adds and one shift-add: keeps the multiplier pipe for the dot products
}

struct PuState {
  unsigned a[4], b[4];     // level 0->1: horizontally filtered rows r-4..r-1, packed column pairs (k0,k1) and (k2,k3)
  unsigned g0[3], g1[3];   // level 1->2: horizontally filtered level-1 rows, the lane's two level-2 columns
  unsigned c[3];           // level 2->3: horizontally filtered level-2 rows, the lane's level-3 column
};

// horizontal 5-tap at level 1 -> the lane's two level-2 columns.  P = level-1 columns (4L, 4L+1), Q = (4L+2, 4L+3).
template <bool LEFT, bool RIGHT>
__device__ __forceinline__ void pu_h1(unsigned P, unsigned Q, int lane, int last_lane, unsigned& g0, unsigned& g1) {
  unsigned Ql = __shfl_up_sync(0xffffffffu, Q, 1);
  unsigned Pr = __shfl_down_sync(0xffffffffu, P, 1);
  if (LEFT && lane == 0) Ql = __byte_perm(P, Q, 0x3254);   // columns -2, -1 are columns 2, 1
  if (RIGHT && lane == last_lane) Pr = Q;                  // column W1 is column W1-2
#if PU_ALU_TAPS & 1
  g0 = __dp2a_lo(Ql, 0x0401u, __dp2a_lo(P, 0x0406u, Q & 0xffffu));
  g1 = __dp2a_lo(P, 0x0401u, __dp2a_lo(Q, 0x0406u, Pr & 0xffffu));
#else
  g0 = __dp2a_lo(Ql, 0x0401u, __dp2a_lo(P, 0x0406u, __dp2a_lo(Q, 0x0001u, 0u)));
  g1 = __dp2a_lo(P, 0x0401u, __dp2a_lo(Q, 0x0406u, __dp2a_lo(Pr, 0x0001u, 0u)));
#endif
}
// horizontal 5-tap at level 2 -> the lane's level-3 column.  q0, q1 = level-2 columns (2L, 2L+1).
template <bool LEFT, bool RIGHT>
__device__ __forceinline__ unsigned pu_h2(unsigned q0, unsigned q1, int lane, int last_lane) {
  unsigned ql0 = __shfl_up_sync(0xffffffffu, q0, 1);
  unsigned ql1 = __shfl_up_sync(0xffffffffu, q1, 1);
  unsigned qr0 = __shfl_down_sync(0xffffffffu, q0, 1);
  if (LEFT && lane == 0) { ql0 = qr0; ql1 = q1; }
  if (RIGHT && lane == last_lane) qr0 = q0;
  return v5(ql0, ql1, q0, q1, qr0);
}

template <bool LEFT, bool RIGHT>
__device__ __forceinline__ void pu_block(PuState& s, const uint2 w[PU_ROWS], int lane, int last_lane, unsigned& out3) {
  unsigned ha[PU_ROWS], hb[PU_ROWS];
#pragma unroll
  for (int i = 0; i < PU_ROWS; ++i) {
    const unsigned w0 = w[i].x, w1 = w[i].y;
    unsigned wl = __shfl_up_sync(0xffffffffu, w1, 1);
    unsigned wr = __shfl_down_sync(0xffffffffu, w0, 1);
    if (LEFT && lane == 0) wl = __byte_perm(w0, w1, 0x1234);          // pixels -2, -1 are pixels 2, 1
    if (RIGHT && lane == last_lane) wr = __byte_perm(w1, 0u, 0x0002); // pixel W is pixel W-2
    const unsigned k0 = __dp4a(wl, 0x04010000u, __dp4a(w0, 0x00010406u, 0u));
#if PU_ALU_TAPS & 2
    const unsigned k1 = __dp4a(w0, 0x04060401u, w1 & 0xffu);
    const unsigned k2 = __dp4a(w0, 0x04010000u, __dp4a(w1, 0x00010406u, 0u));
    const unsigned k3 = __dp4a(w1, 0x04060401u, wr & 0xffu);
#else
    const unsigned k1 = __dp4a(w0, 0x04060401u, __dp4a(w1, 0x00000001u, 0u));
    const unsigned k2 = __dp4a(w0, 0x04010000u, __dp4a(w1, 0x00010406u, 0u));
    const unsigned k3 = __dp4a(w1, 0x04060401u, __dp4a(wr, 0x00000001u, 0u));
#endif
    ha[i] = k0 + (k1 << 16);
    hb[i] = k2 + (k3 << 16);
  }
  // level-1 rows 4b-1 .. 4b+2 (packed pairs)
  unsigned P[4], Q[4];
  P[0] = v5(s.a[0], s.a[1], s.a[2], s.a[3], ha[0]);
  P[1] = v5(s.a[2], s.a[3], ha[0], ha[1], ha[2]);
  P[2] = v5(ha[0], ha[1], ha[2], ha[3], ha[4]);
  P[3] = v5(ha[2], ha[3], ha[4], ha[5], ha[6]);
  Q[0] = v5(s.b[0], s.b[1], s.b[2], s.b[3], hb[0]);
  Q[1] = v5(s.b[2], s.b[3], hb[0], hb[1], hb[2]);
  Q[2] = v5(hb[0], hb[1], hb[2], hb[3], hb[4]);
  Q[3] = v5(hb[2], hb[3], hb[4], hb[5], hb[6]);
#pragma unroll
  for (int i = 0; i < 4; ++i) { s.a[i] = ha[4 + i]; s.b[i] = hb[4 + i]; }
  unsigned n0[4], n1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) pu_h1<LEFT, RIGHT>(P[i], Q[i], lane, last_lane, n0[i], n1[i]);
  // level-2 rows 2b-1, 2b
  unsigned q0[2], q1[2];
  q0[0] = v5(s.g0[0], s.g0[1], s.g0[2], n0[0], n0[1]);
  q0[1] = v5(s.g0[2], n0[0], n0[1], n0[2], n0[3]);
  q1[0] = v5(s.g1[0], s.g1[1], s.g1[2], n1[0], n1[1]);
  q1[1] = v5(s.g1[2], n1[0], n1[1], n1[2], n1[3]);
#pragma unroll
  for (int i = 0; i < 3; ++i) { s.g0[i] = n0[1 + i]; s.g1[i] = n1[1 + i]; }
  const unsigned m0 = pu_h2<LEFT, RIGHT>(q0[0], q1[0], lane, last_lane);
  const unsigned m1 = pu_h2<LEFT, RIGHT>(q0[1], q1[1], lane, last_lane);
  // level-3 row b-1
  out3 = v5(s.c[0], s.c[1], s.c[2], m0, m1);
  s.c[0] = s.c[2]; s.c[1] = m0; s.c[2] = m1;
}

// bottom border of even-sized levels: the last output row of each level uses the window (n-4, n-3, n-2, n-1, n-2)
template <bool LEFT, bool RIGHT>
__device__ __forceinline__ unsigned pu_flush(const PuState& s, int lane, int last_lane) {
  const unsigned P = v5(s.a[0], s.a[1], s.a[2], s.a[3], s.a[2]);
  const unsigned Q = v5(s.b[0], s.b[1], s.b[2], s.b[3], s.b[2]);
  unsigned n0, n1;
  pu_h1<LEFT, RIGHT>(P, Q, lane, last_lane, n0, n1);
  const unsigned q0 = v5(s.g0[0], s.g0[1], s.g0[2], n0, s.g0[2]);
  const unsigned q1 = v5(s.g1[0], s.g1[1], s.g1[2], n1, s.g1[2]);
  const unsigned m = pu_h2<LEFT, RIGHT>(q0, q1, lane, last_lane);
  return v5(s.c[0], s.c[1], s.c[2], m, s.c[2]);
}

// MODE 0: level-3 rows to HBM (g3);  MODE 1: to the slot's shared image (l3), rows staged by cp.async;  MODE 2: same, rows
// staged by TMA.  `phase` (MODE 2) holds the parity the warp waits for next on each of its PU_STAGES mbarriers.
template <int WT, bool LEFT, bool RIGHT, int MODE>
__device__ __forceinline__ void pu_run_frame(const PuParams& p, const CUtensorMap* tmap, const uint8_t* __restrict__ fsrc,
                                             int sframe, uint32_t* __restrict__ g3, unsigned char* ring, unsigned mbar,
                                             unsigned& phase, int lane, int col, int store_lo, int store_hi,
                                             int last_lane) {
  const int W = WT ? WT : p.W;        // a compile-time width turns the row offsets of the copies into immediates
  const bool in_img = col >= 0 && col < p.W3 && lane <= store_hi + PU_HALO_LANES;
  const uint8_t* lsrc = fsrc + (in_img ? 8 * col : 0);
  const int src_bytes = in_img ? 8 : 0;
  const int nblk = p.H >> 3;
  unsigned char* my = ring + lane * 8;
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
  const int x0 = 8 * (col - lane);    // first column of the warp's 256-byte window (>= 0; the right end may hang over)
  auto issue = [&](int b) {
    if (MODE == 2) {
      if (b < nblk) {
        __syncwarp();                                              // every lane is done with the stage's previous rows
        if (lane == 0) {
          const int stage = (b + 2) & (PU_STAGES - 1);
          pu_mbar_expect_tx(mbar + 8 * stage, PU_STAGE_BYTES);
          // rows above the frame are the mirrored rows 16..9 / 8..1: fetched in frame order, read back bottom-up
          pu_tma_load(ring_s + stage * PU_STAGE_BYTES, tmap, x0, b >= 0 ? 8 * b : -8 * b - 7, sframe, mbar + 8 * stage);
        }
      }
      return;
    }
    if (b < nblk) {
      unsigned char* dst = my + ((b + 2) & (PU_STAGES - 1)) * PU_STAGE_BYTES;
      if (b >= 0) {
        const uint8_t* blk = lsrc + (long long)(8 * b) * W;
#pragma unroll
        for (int i = 0; i < PU_ROWS; ++i) pu_cp_async8(dst + i * 256, blk + i * W, src_bytes);
      } else {
#pragma unroll
        for (int i = 0; i < PU_ROWS; ++i)                      // mirrored rows above the frame
          pu_cp_async8(dst + i * 256, lsrc + (long long)(-(8 * b + i)) * W, src_bytes);
      }
    }
    pu_commit();
  };
  PuState s;
#pragma unroll
  for (int i = 0; i < 4; ++i) { s.a[i] = 0; s.b[i] = 0; }
#pragma unroll
  for (int i = 0; i < 3; ++i) { s.g0[i] = 0; s.g1[i] = 0; s.c[i] = 0; }
  issue(-2);
  issue(-1);
  issue(0);
  const bool storing = lane >= store_lo && lane <= store_hi;
  uint32_t* out = g3 + col;
  auto step = [&](int b, bool mirrored) {
    issue(b + 3);
    const int stage = (b + 2) & (PU_STAGES - 1);
    if (MODE == 2) {
      pu_mbar_wait(mbar + 8 * stage, (phase >> stage) & 1u);
      phase ^= 1u << stage;
    } else {
      pu_wait<3>();                                            // block b has landed (3 younger groups may be in flight)
    }
    const unsigned char* src = my + stage * PU_STAGE_BYTES;
    uint2 w[PU_ROWS];
#pragma unroll
    for (int i = 0; i < PU_ROWS; ++i)
      w[i] = *reinterpret_cast<const uint2*>(src + ((MODE == 2 && mirrored) ? PU_ROWS - 1 - i : i) * 256);
    unsigned o;
    pu_block<LEFT, RIGHT>(s, w, lane, last_lane, o);
    if (b >= 1 && storing) out[(long long)(b - 1) * p.W3] = o;
  };
  step(-2, true);
  step(-1, true);
#pragma unroll 2
  for (int b = 0; b < nblk; ++b) step(b, false);
  const unsigned o = pu_flush<LEFT, RIGHT>(s, lane, last_lane);
  if (storing) out[(long long)(nblk - 1) * p.W3] = o;
  if (MODE != 2) pu_wait<0>();
}

// ---------------------------------------------------------------------------------------------------- fused tail
__device__ __forceinline__ void pu_slot_sync(int slot, int nt) {
  if (nt == 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;\n" ::"r"(slot + 1), "r"(nt) : "memory");
}
__device__ __forceinline__ double* pu_level(const PuParams& p, int l, unsigned char* ring_slot, unsigned char* extra_slot) {
  const int o = p.lvl_off[l];
  return reinterpret_cast<double*>(o >= 0 ? ring_slot + o : extra_slot + (-o - 1));
}
// G_{l+1} = pyrDown(G_l) on float64 (pyramid.py:13-15), the arithmetic of pyramid_tail_kernel
__device__ __forceinline__ void pu_down(const double* __restrict__ s, double* __restrict__ d, int sw, int sh, int dw, int dh,
                                        int t, int nt) {
  for (int i = t; i < dw * dh; i += nt) {
    const int x = i % dw, y = i / dw;
    const int x0 = reflect101(2 * x - 2, sw), x1 = reflect101(2 * x - 1, sw), x3 = reflect101(2 * x + 1, sw),
              x4 = reflect101(2 * x + 2, sw);
    double r[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const double* row = s + reflect101(2 * y + k - 2, sh) * sw;
      r[k] = tap5(row[x0], row[x1], row[2 * x], row[x3], row[x4]);
    }
    d[i] = tap5(r[0], r[1], r[2], r[3], r[4]) * (1.0 / 256.0);
  }
}
// L_l = G_l - pyrUp(G_{l+1})   (pyramid.py:24-26)
__device__ __forceinline__ void pu_lap(const double* __restrict__ cur, const double* __restrict__ s, double* __restrict__ out,
                                       int sw, int sh, int dw, int dh, int t, int nt) {
  for (int i = t; i < dw * dh; i += nt) {
    const int x = i % dw, y = i / dw;
    const UpTaps tx = up_taps(x, sw), ty = up_taps(y, sh);
    const double* r0 = s + ty.i0 * sw;
    const double* r1 = s + ty.i1 * sw;
    const double* r2 = s + ty.i2 * sw;
    const double h0 = up_combine(tx, r0[tx.i0], r0[tx.i1], r0[tx.i2]);
    const double h1 = up_combine(tx, r1[tx.i0], r1[tx.i1], r1[tx.i2]);
    const double h2 = up_combine(tx, r2[tx.i0], r2[tx.i1], r2[tx.i2]);
    out[i] = cur[i] - up_combine(ty, h0, h1, h2) * (1.0 / 64.0);
  }
}
// Levels first..top and the Laplacian record of one frame, by the nt threads of the frame's slot (t = 0..nt-1).
// l3: the frame's level-3 image (exact integers) in shared memory; the slot's ring is idle and holds the levels.
__device__ __noinline__ void pu_tail(const PuParams& p, const uint32_t* __restrict__ l3, unsigned char* ring_slot,
                                     unsigned char* extra_slot, double* __restrict__ rec, int slot, int t, int nt) {
  const int f = p.first, top = p.top;
  {   // integer level first-1 -> level first: the 5x5 sums stay exact in float64 (< 2^53), one scale at the end
    const int sw = p.W3, sh = p.H3, dw = p.w[f], dh = p.h[f];
    double* g = pu_level(p, f, ring_slot, extra_slot);
    for (int i = t; i < dw * dh; i += nt) {
      const int x = i % dw, y = i / dw;
      const int x0 = reflect101(2 * x - 2, sw), x1 = reflect101(2 * x - 1, sw), x3 = reflect101(2 * x + 1, sw),
                x4 = reflect101(2 * x + 2, sw);
      double r[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const uint32_t* row = l3 + reflect101(2 * y + k - 2, sh) * sw;
        r[k] = tap5((double)row[x0], (double)row[x1], (double)row[2 * x], (double)row[x3], (double)row[x4]);
      }
      g[i] = tap5(r[0], r[1], r[2], r[3], r[4]) * p.g_scale;
    }
  }
  pu_slot_sync(slot, nt);
  if (f + 1 <= top)
    pu_down(pu_level(p, f, ring_slot, extra_slot), pu_level(p, f + 1, ring_slot, extra_slot), p.w[f], p.h[f], p.w[f + 1],
            p.h[f + 1], t, nt);
  pu_slot_sync(slot, nt);
  // the small levels are a chain of tiny images: one warp walks it while the others write the largest Laplacian
  if (t < 32) {
    for (int l = f + 1; l < top; ++l) {
      pu_down(pu_level(p, l, ring_slot, extra_slot), pu_level(p, l + 1, ring_slot, extra_slot), p.w[l], p.h[l], p.w[l + 1],
              p.h[l + 1], t, 32);
      __syncwarp();
    }
  }
  if (nt == 32 || t >= 32) {
    const int t2 = nt == 32 ? t : t - 32, nt2 = nt == 32 ? 32 : nt - 32;
    pu_lap(pu_level(p, f, ring_slot, extra_slot), pu_level(p, f + 1, ring_slot, extra_slot), rec + p.rec_off[f], p.w[f + 1],
           p.h[f + 1], p.w[f], p.h[f], t2, nt2);
  }
  pu_slot_sync(slot, nt);
  for (int l = f + 1; l < top; ++l)
    pu_lap(pu_level(p, l, ring_slot, extra_slot), pu_level(p, l + 1, ring_slot, extra_slot), rec + p.rec_off[l], p.w[l + 1],
           p.h[l + 1], p.w[l], p.h[l], t, nt);
}

template <int WT, int MODE>
__global__ void __launch_bounds__(PU_BOUND_WARPS * 32, PU_CTAS_PER_SM)
    pyramid_front_u8_kernel(const __grid_constant__ PuParams p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp / p.n_strips, strip = warp - slot * p.n_strips;
  unsigned char* ring = smem + (size_t)warp * PU_STAGES * PU_STAGE_BYTES;
  const bool left = strip == 0, right = strip == p.n_strips - 1;
  const int c0 = strip * p.cols_per_strip;
  const int c1 = min(p.W3, c0 + p.cols_per_strip);
  const int lane_off = left ? 0 : PU_HALO_LANES;
  const int col = c0 - lane_off + lane;
  const int store_lo = lane_off, store_hi = lane_off + (c1 - c0) - 1;
  const long long g3_elems = (long long)p.W3 * p.H3;
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(smem + p.mbar_base) + warp * PU_STAGES * 8;
  unsigned phase = 0;
  if (MODE == 2) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < PU_STAGES; ++i) pu_mbar_init(mbar + 8 * i, 1);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
  }
  uint32_t* l3 = reinterpret_cast<uint32_t*>(smem + p.l3_base + (size_t)slot * p.l3_stride);
  unsigned char* ring_slot = smem + (size_t)slot * p.n_strips * PU_STAGES * PU_STAGE_BYTES;
  unsigned char* extra_slot = smem + p.extra_base + (size_t)slot * p.extra_stride;
  for (long long frame = (long long)blockIdx.x * p.frames_per_cta + slot; frame < p.n_frames;
       frame += (long long)gridDim.x * p.frames_per_cta) {
#ifdef PU_IDX32
    // Developer switch (untimed experiment): the 64-bit division costs 8 % of the kernel's stall samples (ncu r01m) for one
    // use per frame; frame counts and segment lengths fit 32 bits (checked by pu_launch)
    const unsigned fr = (unsigned)frame, sl = (unsigned)p.seg_len;
    const long long sframe = (long long)(fr / sl) * p.seg_stride + p.seg_first + (long long)(fr % sl);
#else
    const long long sframe = (frame / p.seg_len) * p.seg_stride + p.seg_first + frame % p.seg_len;
#endif
    const uint8_t* fsrc = p.frames + sframe * p.frame_elems;
    uint32_t* g3 = MODE == 0 ? p.g3 + frame * g3_elems : l3;
    // The strips of a frame share their halo columns.  Left alone, the warps of a slot drift apart over the frames (edge
    // strips are cheaper) until a halo sector read by one warp has left L2 before its neighbour asks for it -- 23 % extra
    // DRAM reads at 8192 frames per launch (ncu, profiles/r01g).  A named barrier per slot re-aligns them every frame
    // (fused tail: it also keeps the next frame's rows out of the ring until every warp has left the tail).
    if (p.n_strips > 1) asm volatile("bar.sync %0, %1;\n" ::"r"(slot + 1), "r"(p.n_strips * 32) : "memory");
    else if (MODE != 0) __syncwarp();
    if (left && right) pu_run_frame<WT, true, true, MODE>(p, &tmap, fsrc, (int)sframe, g3, ring, mbar, phase, lane, col, store_lo, store_hi, store_hi);
    else if (left) pu_run_frame<WT, true, false, MODE>(p, &tmap, fsrc, (int)sframe, g3, ring, mbar, phase, lane, col, store_lo, store_hi, store_hi);
    else if (right) pu_run_frame<WT, false, true, MODE>(p, &tmap, fsrc, (int)sframe, g3, ring, mbar, phase, lane, col, store_lo, store_hi, store_hi);
    else pu_run_frame<WT, false, false, MODE>(p, &tmap, fsrc, (int)sframe, g3, ring, mbar, phase, lane, col, store_lo, store_hi, store_hi);
    if (MODE != 0) {
      const int nt = p.n_strips * 32;
      pu_slot_sync(slot, nt);                                  // the level-3 image is complete, the ring is idle
      pu_tail(p, l3, ring_slot, extra_slot, p.lap_out + frame * p.record_len, slot, strip * 32 + lane, nt);
      // the levels were written to the ring through the generic proxy; the next frame's rows arrive through the async one
      if (MODE == 2) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
  }
}

// The integer path needs: uint8 frames, four front levels, W and H multiples of 8 (all three decimated sizes even),
// at least 17 rows (the mirrored lead-in) and 8-byte aligned rows.
bool pu_supported(const void* frames, int W, int H, int skip) {
  return skip == 4 && W % 8 == 0 && H % 8 == 0 && H >= 24 && W >= 16 && ((uintptr_t)frames % 8) == 0;
}

struct PuPlan {
  PuParams p;
  int warps, smem;
};
// geometry and shared-memory layout of a launch; false if the fused layout does not fit the SM
static bool pu_plan(rm_handle* h, int mode, int W, int H, PuPlan& pl) {
  PuParams& p = pl.p;
  memset(&p, 0, sizeof(p));
  p.frame_elems = (long long)W * H;
  p.W = W; p.H = H; p.W3 = W / 8; p.H3 = H / 8;
  const int cap = 32 - 2 * PU_HALO_LANES;                       // payload lanes of an interior strip
  p.n_strips = (p.W3 + cap - 1) / cap;
  p.cols_per_strip = (p.W3 + p.n_strips - 1) / p.n_strips;
  if (p.n_strips > 16) return false;
  const int L = h->p.pyramid_levels, s = h->p.skip_levels_at_top;
  LevelGeom g = make_geom(W, H, L);
  RecordGeom rec = make_record(g, s);
  p.first = s; p.top = L - 1; p.record_len = rec.len;
  p.g_scale = 1.0 / 255;
  for (int l = 0; l < s; ++l) p.g_scale *= 1.0 / 256.0;
  for (int l = 0; l < L; ++l) { p.w[l] = g.w[l]; p.h[l] = g.h[l]; p.rec_off[l] = rec.off[l]; }
  const int ring_slot = p.n_strips * PU_STAGES * PU_STAGE_BYTES;
  int extra = 0, ring_used = 0;
  if (mode != 0) {
    for (int l = s; l < L; ++l) {                               // the levels live in the slot's idle ring where they fit
      const int bytes = g.w[l] * g.h[l] * 8;
      if (ring_used + bytes <= ring_slot) { p.lvl_off[l] = ring_used; ring_used += bytes; }
      else { p.lvl_off[l] = -(extra + 1); extra += bytes; }
    }
  }
  const int l3_bytes = mode != 0 ? ((p.W3 * p.H3 * 4 + 15) & ~15) : 0;
  extra = (extra + 15) & ~15;
  int fpc = PU_MAX_WARPS / p.n_strips;
  if (fpc < 1) return false;
  if (mode != 0) {
    const int per_slot = ring_slot + l3_bytes + extra + (mode == 2 ? p.n_strips * PU_STAGES * 8 : 0);
    const int fit = (h->smem_optin - 128) / per_slot;
    if (fit < 1) return false;
    if (fpc > fit) fpc = fit;
    if (fpc > 15) fpc = 15;                                     // named barriers 1..15
  }
  p.frames_per_cta = fpc;
  pl.warps = fpc * p.n_strips;
  p.ring_bytes = pl.warps * PU_STAGES * PU_STAGE_BYTES;
  p.l3_base = p.ring_bytes; p.l3_stride = l3_bytes;
  p.extra_base = p.l3_base + fpc * l3_bytes; p.extra_stride = extra;
  p.mbar_base = p.extra_base + fpc * extra;
  pl.smem = p.mbar_base + (mode == 2 ? pl.warps * PU_STAGES * 8 : 0);
  return pl.smem <= h->smem_optin;
}

typedef CUresult (*pu_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static pu_encode_fn pu_encoder() {
  static pu_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (pu_encode_fn)sym;
  }
  return fn;
}

// The best mode this build supports for these frames: 2 (TMA) needs 16-byte aligned rows and frames, 1 needs the fused
// layout to fit, 0 otherwise.  Option "pyramid_mode" (0/1/2) caps it (A/B timing, tests of every path).
int pu_best_mode(rm_handle* h, const void* frames, int W, int H, int n_levels, int skip) {
  (void)n_levels; (void)skip;
  PuPlan pl;
  int mode = h->pyramid_mode;
  if (mode == 2 && !(W % 16 == 0 && ((uintptr_t)frames % 16) == 0 && pu_encoder() && pu_plan(h, 2, W, H, pl))) mode = 1;
  if (mode == 1 && !pu_plan(h, 1, W, H, pl)) mode = 0;
  return mode;
}

template <int MODE>
static int32_t pu_launch_mode(rm_handle* h, const PuPlan& pl, const CUtensorMap& map, long long ctas, cudaStream_t st) {
  const int W = pl.p.W;
  void (*kern)(const PuParams, const CUtensorMap) =
      W == 640 ? pyramid_front_u8_kernel<640, MODE>
      : W == 1280 ? pyramid_front_u8_kernel<1280, MODE>
      : W == 1920 ? pyramid_front_u8_kernel<1920, MODE>
      : W == 320 ? pyramid_front_u8_kernel<320, MODE> : pyramid_front_u8_kernel<0, MODE>;
  RM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem));
  RM_PROF(h, st, MODE == 0 ? "pyramid_front_u8_kernel" : (MODE == 1 ? "pyramid_u8_fused_kernel" : "pyramid_u8_fused_tma_kernel"));
  kern<<<(unsigned)ctas, pl.warps * 32, pl.smem, st>>>(pl.p, map);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

int32_t pu_launch(rm_handle* h, int mode, const uint8_t* frames, uint32_t* g3, double* lap_out, long long n_frames,
                  long long seg_len, long long seg_stride, long long seg_first, int W, int H, cudaStream_t st) {
  PuPlan pl;
  if (!pu_plan(h, mode, W, H, pl))
    return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: frame too wide / level images too large for mode %lld", __func__, mode);
  PuParams& p = pl.p;
  p.frames = frames; p.g3 = g3; p.lap_out = lap_out; p.n_frames = n_frames;
  p.seg_len = seg_len; p.seg_stride = seg_stride; p.seg_first = seg_first;
#ifdef PU_IDX32
  if (n_frames >= (1ll << 31) || seg_len >= (1ll << 31) || seg_len < 1)
    return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: more than 2^31 frames", __func__);
#endif
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (mode == 2) {
    // every source frame the launch can touch, as one (W, H, frames) uint8 tensor; box = 256 columns x PU_ROWS rows
    const long long last = n_frames - 1;
    const long long n_src = (last / seg_len) * seg_stride + seg_first + last % seg_len + 1;
    if (n_src >= (1ll << 31)) return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: more than 2^31 source frames", __func__);
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_src};
    const cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
    const cuuint32_t box[3] = {256, PU_ROWS, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    pu_encode_fn enc = pu_encoder();
    if (!enc) return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: cuTensorMapEncodeTiled not available", __func__);
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)frames, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return rm_fail(h, RM_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%lld)", __func__, (long long)r);
  }
  long long ctas = (n_frames + p.frames_per_cta - 1) / p.frames_per_cta;
  if (ctas > (long long)h->sm_count * PU_CTAS_PER_SM) ctas = (long long)h->sm_count * PU_CTAS_PER_SM;
  if (mode == 0) return pu_launch_mode<0>(h, pl, map, ctas, st);
  if (mode == 1) return pu_launch_mode<1>(h, pl, map, ctas, st);
  return pu_launch_mode<2>(h, pl, map, ctas, st);
}
