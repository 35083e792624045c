// The pyramid stage for uint8 frames: frame in -> packed Laplacian record (levels 4..7) out, one kernel, one pass.
//
// Reference: create_laplacian_image_pyramid (pyramid.py:9-28) on frames that are gray/255 (transforms.py:20-23), of which
// transforms.py:156-170 only ever reads the Laplacian levels skip..levels-2.  A uint8 frame makes every Gaussian level
// an exact integer over a power of two:
//   level 1 <= 255*2^8 (u16), level 2 <= 255*2^16, level 3 <= 255*2^24 (u32), level 4 <= 255*2^32 (u64),
// so levels 0..4 are carried as integers -- no rounding at all until the single scale by 2^-32/255 at level 4 -- and
// the 5-tap rows of the first levels are dp4a / dp2a dot products.
//
// Work decomposition: one WARP streams one vertical strip of one frame top to bottom, 8 pixels per lane per row (a
// 256-pixel window); the warps of a frame form a *slot*, a CTA holds several slots.
//   * rows arrive 8 at a time in a warp-private shared-memory ring.  pyramid_u8_fused_kernel: one TMA box per stage
//     (cp.async.bulk.tensor.3d: 256 bytes x 8 rows of one frame, zero-filled outside the frame) issued by lane 0 and
//     an mbarrier per stage; every lane reads back only its own 8 bytes per row.  pyramid_front_u8_kernel (fallback,
//     rows that are not 16-byte multiples): eight 8-byte cp.async per lane per stage;
//   * neighbouring pixels come from the neighbouring lanes by shuffle; lanes on the interior sides of a strip are
//     halo (recomputed by the neighbouring strip), image borders are reflect-101 by byte permutes in the edge lanes;
//   * the vertical 5-tap windows of levels 1..3 are rolling registers: a block of 8 input rows yields 4 level-1 rows,
//     2 level-2 rows and 1 level-3 row per lane;
//   * fused kernel: the level-3 row is filtered horizontally across the lanes (64-bit integers) and every second
//     block the even-column lanes emit a level-4 value (exact integer * 2^-32/255, the reference's float64 value to
//     within its own rounding) into the slot's level-4 image in shared memory.  When the strips of a frame are
//     through, the slot's warps build levels 5..8 (pyramid.py:13-15) and the Laplacians 4..7 (pyramid.py:24-26) in
//     float64 with the arithmetic of pyramid_tail_kernel operation by operation, and write the record;
//   * the top border is handled by streaming 16 mirrored rows first (reflect-101 about index 0 commutes with the
//     symmetric kernel and the 2:1 decimation); the bottom border (even sizes do not commute) by an explicit flush;
//     level 4 reflects its own top and bottom rows explicitly.
// HBM traffic per frame (fused): W*H bytes read once (+ halo columns, L2 hits) and the record written once -- the
// algorithmic bytes of SURVEY 8(d).  The fallback writes level 3 (W*H/16 bytes) and pyramid_tail_kernel finishes.
#include <cuda.h>
#include "common.cuh"
#include "pyramid_u8.cuh"

#define PU_ROWS 8
#define PU_STAGE_BYTES (PU_ROWS * 256)
#define PU_FRONT_STAGES 4      // fallback kernel: cp.async ring depth
#define PU_BGR_STAGES 3        // ... for BGR frames (a stage is three times as large)
#define PU_FRONT_WARPS 24      // fallback kernel: warps per CTA (85 registers; 0.626 -> 0.569 ms per 8192 VGA frames, r02a)

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
__device__ __forceinline__ unsigned long long v5q(unsigned long long a, unsigned long long b, unsigned long long c,
                                                  unsigned long long d, unsigned long long e) {
  return ((b + d + c) << 2) + (a + e + (c + c));
}

struct PuState {
  unsigned a[4], b[4];     // level 0->1: horizontally filtered rows r-4..r-1, packed column pairs (k0,k1) and (k2,k3)
  unsigned g0[3], g1[3];   // level 1->2: horizontally filtered level-1 rows, the lane's two level-2 columns
  unsigned c[3];           // level 2->3: horizontally filtered level-2 rows, the lane's level-3 column
};

// horizontal 5-tap at level 1 -> the lane's two level-2 columns.  P = level-1 columns (4L, 4L+1), Q = (4L+2, 4L+3).
template <bool LEFT, bool RIGHT>
__device__ __forceinline__ void pu_h1(unsigned P, unsigned Q, int lane, int first_lane, int last_lane, unsigned& g0, unsigned& g1) {
  unsigned Ql = __shfl_up_sync(0xffffffffu, Q, 1);
  unsigned Pr = __shfl_down_sync(0xffffffffu, P, 1);
  if (LEFT && lane == first_lane) Ql = __byte_perm(P, Q, 0x3254);   // columns -2, -1 are columns 2, 1
  if (RIGHT && lane == last_lane) Pr = Q;                  // column W1 is column W1-2
  g0 = __dp2a_lo(Ql, 0x0401u, __dp2a_lo(P, 0x0406u, __dp2a_lo(Q, 0x0001u, 0u)));
  g1 = __dp2a_lo(P, 0x0401u, __dp2a_lo(Q, 0x0406u, __dp2a_lo(Pr, 0x0001u, 0u)));
}
// horizontal 5-tap at level 2 -> the lane's level-3 column.  q0, q1 = level-2 columns (2L, 2L+1).
template <bool LEFT, bool RIGHT>
__device__ __forceinline__ unsigned pu_h2(unsigned q0, unsigned q1, int lane, int first_lane, int last_lane) {
  unsigned ql0 = __shfl_up_sync(0xffffffffu, q0, 1);
  unsigned ql1 = __shfl_up_sync(0xffffffffu, q1, 1);
  unsigned qr0 = __shfl_down_sync(0xffffffffu, q0, 1);
  if (LEFT && lane == first_lane) { ql0 = qr0; ql1 = q1; }
  if (RIGHT && lane == last_lane) qr0 = q0;
  return v5(ql0, ql1, q0, q1, qr0);
}

template <bool LEFT, bool RIGHT>
__device__ __forceinline__ void pu_block(PuState& s, const uint2 w[PU_ROWS], int lane, int first_lane, int last_lane, unsigned& out3) {
  unsigned ha[PU_ROWS], hb[PU_ROWS];
#pragma unroll
  for (int i = 0; i < PU_ROWS; ++i) {
    const unsigned w0 = w[i].x, w1 = w[i].y;
    unsigned wl = __shfl_up_sync(0xffffffffu, w1, 1);
    unsigned wr = __shfl_down_sync(0xffffffffu, w0, 1);
    if (LEFT && lane == first_lane) wl = __byte_perm(w0, w1, 0x1234);          // pixels -2, -1 are pixels 2, 1
    if (RIGHT && lane == last_lane) wr = __byte_perm(w1, 0u, 0x0002); // pixel W is pixel W-2
    const unsigned k0 = __dp4a(wl, 0x04010000u, __dp4a(w0, 0x00010406u, 0u));
    const unsigned k1 = __dp4a(w0, 0x04060401u, __dp4a(w1, 0x00000001u, 0u));
    const unsigned k2 = __dp4a(w0, 0x04010000u, __dp4a(w1, 0x00010406u, 0u));
    const unsigned k3 = __dp4a(w1, 0x04060401u, __dp4a(wr, 0x00000001u, 0u));
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
  for (int i = 0; i < 4; ++i) pu_h1<LEFT, RIGHT>(P[i], Q[i], lane, first_lane, last_lane, n0[i], n1[i]);
  // level-2 rows 2b-1, 2b
  unsigned q0[2], q1[2];
  q0[0] = v5(s.g0[0], s.g0[1], s.g0[2], n0[0], n0[1]);
  q0[1] = v5(s.g0[2], n0[0], n0[1], n0[2], n0[3]);
  q1[0] = v5(s.g1[0], s.g1[1], s.g1[2], n1[0], n1[1]);
  q1[1] = v5(s.g1[2], n1[0], n1[1], n1[2], n1[3]);
#pragma unroll
  for (int i = 0; i < 3; ++i) { s.g0[i] = n0[1 + i]; s.g1[i] = n1[1 + i]; }
  const unsigned m0 = pu_h2<LEFT, RIGHT>(q0[0], q1[0], lane, first_lane, last_lane);
  const unsigned m1 = pu_h2<LEFT, RIGHT>(q0[1], q1[1], lane, first_lane, last_lane);
  // level-3 row b-1
  out3 = v5(s.c[0], s.c[1], s.c[2], m0, m1);
  s.c[0] = s.c[2]; s.c[1] = m0; s.c[2] = m1;
}

// bottom border of even-sized levels: the last output row of each level uses the window (n-4, n-3, n-2, n-1, n-2)
template <bool LEFT, bool RIGHT>
__device__ __forceinline__ unsigned pu_flush(const PuState& s, int lane, int first_lane, int last_lane) {
  const unsigned P = v5(s.a[0], s.a[1], s.a[2], s.a[3], s.a[2]);
  const unsigned Q = v5(s.b[0], s.b[1], s.b[2], s.b[3], s.b[2]);
  unsigned n0, n1;
  pu_h1<LEFT, RIGHT>(P, Q, lane, first_lane, last_lane, n0, n1);
  const unsigned q0 = v5(s.g0[0], s.g0[1], s.g0[2], n0, s.g0[2]);
  const unsigned q1 = v5(s.g1[0], s.g1[1], s.g1[2], n1, s.g1[2]);
  const unsigned m = pu_h2<LEFT, RIGHT>(q0, q1, lane, first_lane, last_lane);
  return v5(s.c[0], s.c[1], s.c[2], m, s.c[2]);
}

__device__ __forceinline__ void pu_clear(PuState& s) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { s.a[i] = 0; s.b[i] = 0; }
#pragma unroll
  for (int i = 0; i < 3; ++i) { s.g0[i] = 0; s.g1[i] = 0; s.c[i] = 0; }
}

// ==================================================================================================== fallback front
// Level 3 (exact integers) to HBM; pyramid_tail_kernel (pyramid.cu) finishes.  Rows by cp.async, any W % 8 == 0.
// BGR = true: the frames are 8-bit BGR as cv2.VideoCapture delivers them (next_frame, base.py:227-231) and
// cv2.cvtColor(frame, COLOR_BGR2GRAY) is folded into the load: a lane copies the 24 bytes of its 8 pixels per row and
// converts them in registers with OpenCV's 15-bit fixed point, gray = (3735 B + 19235 G + 9798 R + 2^14) >> 15 (the
// weights split into byte halves so that two dp4a per pixel do the products) -- the frame is read once, 3 bytes per
// pixel, instead of converted in a pass of its own (3 read + 1 written) and read again.
__device__ __forceinline__ unsigned pu_gray_of(unsigned px) {      // px = (B, G, R, anything)
  return (__dp4a(px, 0x00462397u, 16384u) + (__dp4a(px, 0x00264b0eu, 0u) << 8)) >> 15;   // lo (151, 35, 70), hi (14, 75, 38)
}
__device__ __forceinline__ uint2 pu_bgr24_to_gray8(uint2 a, uint2 b, uint2 c) {
  // 24 bytes r0..r5 = 8 pixels of 3 bytes; pixel k starts at byte 3k
  const unsigned r0 = a.x, r1 = a.y, r2 = b.x, r3 = b.y, r4 = c.x, r5 = c.y;
  const unsigned g0 = pu_gray_of(r0), g1 = pu_gray_of(__byte_perm(r0, r1, 0x0543)), g2 = pu_gray_of(__byte_perm(r1, r2, 0x0432)),
                 g3 = pu_gray_of(r2 >> 8);
  const unsigned g4 = pu_gray_of(r3), g5 = pu_gray_of(__byte_perm(r3, r4, 0x0543)), g6 = pu_gray_of(__byte_perm(r4, r5, 0x0432)),
                 g7 = pu_gray_of(r5 >> 8);
  return make_uint2(g0 | (g1 << 8) | (g2 << 16) | (g3 << 24), g4 | (g5 << 8) | (g6 << 16) | (g7 << 24));
}

template <int WT, bool LEFT, bool RIGHT, bool BGR>
__device__ __forceinline__ void pu_front_frame(const PuParams& p, const uint8_t* __restrict__ fsrc, uint32_t* __restrict__ g3,
                                               unsigned char* ring, int lane, int col, int store_lo, int store_hi) {
  constexpr int PX = BGR ? 3 : 1;                       // bytes per pixel in the frame
  constexpr int STAGES = BGR ? PU_BGR_STAGES : PU_FRONT_STAGES;
  constexpr int ROWB = 256 * PX, STAGEB = PU_ROWS * ROWB;
  const int W = WT ? WT : p.W;        // a compile-time width turns the row offsets of the copies into immediates
  const int last_lane = store_hi;
  const bool in_img = col >= 0 && col < p.W3 && lane <= store_hi + PU_HALO_LANES;
  const uint8_t* lsrc = fsrc + (in_img ? 8 * PX * col : 0);
  const int src_bytes = in_img ? 8 : 0;
  const int nblk = p.H >> 3;
  unsigned char* my = ring + lane * 8 * PX;
  auto issue = [&](int b) {
    if (b < nblk) {
      unsigned char* dst = my + ((b + 2) % STAGES) * STAGEB;
#pragma unroll
      for (int i = 0; i < PU_ROWS; ++i) {
        // rows above the frame (b < 0) are its mirrored rows
        const uint8_t* row = lsrc + (long long)(b >= 0 ? 8 * b + i : -(8 * b + i)) * (W * PX);
#pragma unroll
        for (int c = 0; c < PX; ++c) pu_cp_async8(dst + i * ROWB + 8 * c, row + 8 * c, src_bytes);
      }
    }
    pu_commit();
  };
  PuState s;
  pu_clear(s);
#pragma unroll
  for (int b = -2; b < STAGES - 3; ++b) issue(b);             // STAGES - 1 blocks in flight
  const bool storing = lane >= store_lo && lane <= store_hi;
  uint32_t* out = g3 + col;
#pragma unroll 2
  for (int b = -2; b < nblk; ++b) {
    issue(b + STAGES - 1);
    pu_wait<STAGES - 1>();                                     // block b has landed (STAGES - 1 younger groups may be in flight)
    const unsigned char* src = my + ((b + 2) % STAGES) * STAGEB;
    uint2 w[PU_ROWS];
#pragma unroll
    for (int i = 0; i < PU_ROWS; ++i) {
      const uint2* q = reinterpret_cast<const uint2*>(src + i * ROWB);
      w[i] = BGR ? pu_bgr24_to_gray8(q[0], q[1], q[2]) : q[0];
    }
    unsigned o;
    pu_block<LEFT, RIGHT>(s, w, lane, 0, last_lane, o);
    if (b >= 1 && storing) out[(long long)(b - 1) * p.W3] = o;
  }
  const unsigned o = pu_flush<LEFT, RIGHT>(s, lane, 0, last_lane);
  if (storing) out[(long long)(nblk - 1) * p.W3] = o;
  pu_wait<0>();
}

// frame f of the batch is source frame (f / seg_len) * seg_stride + seg_first + f % seg_len; 32-bit division (the 64-bit
// one cost 8 % of the front kernel's stall samples for one use per frame, ncu r01m; the launch checks the ranges)
__device__ __forceinline__ long long pu_source_frame(long long frame, long long seg_len, long long seg_stride,
                                                     long long seg_first) {
  const unsigned fr = (unsigned)frame, sl = (unsigned)seg_len;
  return (long long)(fr / sl) * seg_stride + seg_first + (long long)(fr % sl);
}

template <int WT, bool BGR>
__global__ void __launch_bounds__(PU_FRONT_WARPS * 32, 1) pyramid_front_u8_kernel(const PuParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp / p.n_strips, strip = warp - slot * p.n_strips;
  unsigned char* ring = smem + (size_t)warp * (BGR ? PU_BGR_STAGES * 3 : PU_FRONT_STAGES) * PU_STAGE_BYTES;
  const bool left = strip == 0, right = strip == p.n_strips - 1;
  const int c0 = strip * p.cols_per_strip;
  const int c1 = min(p.W3, c0 + p.cols_per_strip);
  const int lane_off = left ? 0 : PU_HALO_LANES;
  const int col = c0 - lane_off + lane;
  const int store_lo = lane_off, store_hi = lane_off + (c1 - c0) - 1;
  const long long g3_elems = (long long)p.W3 * p.H3;
  for (long long frame = (long long)blockIdx.x * p.frames_per_cta + slot; frame < p.n_frames;
       frame += (long long)gridDim.x * p.frames_per_cta) {
    const long long sframe = pu_source_frame(frame, p.seg_len, p.seg_stride, p.seg_first);
    const uint8_t* fsrc = p.frames + sframe * p.frame_elems;
    uint32_t* g3 = p.g3 + frame * g3_elems;
    // The strips of a frame share their halo columns.  Left alone, the warps of a slot drift apart over the frames (edge
    // strips are cheaper) until a halo sector read by one warp has left L2 before its neighbour asks for it -- 23 % extra
    // DRAM reads at 8192 frames per launch (ncu, profiles/r01g).  A named barrier per slot re-aligns them every frame.
    if (p.n_strips > 1) asm volatile("bar.sync %0, %1;\n" ::"r"(slot + 1), "r"(p.n_strips * 32) : "memory");
    if (left && right) pu_front_frame<WT, true, true, BGR>(p, fsrc, g3, ring, lane, col, store_lo, store_hi);
    else if (left) pu_front_frame<WT, true, false, BGR>(p, fsrc, g3, ring, lane, col, store_lo, store_hi);
    else if (right) pu_front_frame<WT, false, true, BGR>(p, fsrc, g3, ring, lane, col, store_lo, store_hi);
    else pu_front_frame<WT, false, false, BGR>(p, fsrc, g3, ring, lane, col, store_lo, store_hi);
  }
}

// ==================================================================================================== fused kernel
// Level-4 columns are the unit of a strip: strip s produces columns [k0, k1); lane L holds level-3 column base + L
// (base even, so the 256-byte window starts on a 16-byte boundary).  Level 3 is valid in lanes 2..30 of an interior
// window (lanes 0.. of the left strip, ..last_lane of the right one), a level-4 column with centre lane c needs lanes
// c-2..c+2: an interior strip yields 13 columns, the edge strips 15 / 14.
__device__ __forceinline__ void pf_slot_sync(int slot, int nt) {
  if (nt == 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;\n" ::"r"(slot + 1), "r"(nt) : "memory");
}
__device__ __forceinline__ double* pf_level(const PfParams& p, unsigned char* lvl_slot, int l) {
  return reinterpret_cast<double*>(lvl_slot + p.lvl_off[l]);
}
// i / w for i < 2^16, 2 <= w < 2^16 with m = ceil(2^32 / w); m = 0 stands for w = 1
__device__ __forceinline__ int pf_div(int i, unsigned m) { return m ? (int)__umulhi((unsigned)i, m) : i; }

// G_{l+1} = pyrDown(G_l) on float64 (pyramid.py:13-15), the arithmetic of pyramid_tail_kernel
// SRC_SMEM = false: s is global memory this kernel has just written (g4_global) -- not __restrict__, so that the loads
// cannot become non-coherent ones
template <bool SRC_SMEM>
__device__ __forceinline__ void pf_down_t(const double* s, double* __restrict__ d, int sw, int sh, int dw, int dh,
                                        unsigned mdw, int t, int nt) {
  for (int i = t; i < dw * dh; i += nt) {
    const int y = pf_div(i, mdw), x = i - y * dw;
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
__device__ __forceinline__ void pf_down(const double* __restrict__ s, double* __restrict__ d, int sw, int sh, int dw, int dh,
                                        unsigned mdw, int t, int nt) {
  pf_down_t<true>(s, d, sw, sh, dw, dh, mdw, t, nt);
}
// L_l = G_l - pyrUp(G_{l+1}) (pyramid.py:24-26), one thread per source pixel = 2x2 outputs from its 3x3 neighbourhood;
// per output the operations of up_taps / up_combine (pyr_core.h) in their order:
//   even index 2i: s[refl(i-1)] + s[min(i+1,n-1)], then fma(6, s[i], .);   odd index 2i+1: 4 * (s[i] + s[min(i+1,n-1)])
// IN_PLACE: cur is out (global memory: every thread reads the four values of `cur` it replaces)
template <bool IN_PLACE>
__device__ __forceinline__ void pf_lap_t(const double* cur, const double* __restrict__ s, double* out,
                                         int sw, int sh, int dw, int dh, unsigned msw, int t, int nt) {
  for (int i = t; i < sw * sh; i += nt) {
    const int bj = pf_div(i, msw), bi = i - bj * sw;
    const int cm = reflect101(bi - 1, sw), cp = bi + 1 < sw - 1 ? bi + 1 : sw - 1;
    const int rm = reflect101(bj - 1, sh), rp = bj + 1 < sh - 1 ? bj + 1 : sh - 1;
    double he[3], ho[3];
    const int rows[3] = {rm, bj, rp};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double* row = s + rows[k] * sw;
      const double a = row[cm], b = row[bi], c = row[cp];
      he[k] = fma(6.0, b, a + c);
      ho[k] = 4.0 * (b + c);
    }
    const int x = 2 * bi, y = 2 * bj;
    const bool x1 = x + 1 < dw, y1 = y + 1 < dh;
    const double* c0 = cur + y * dw + x;
    double* o0 = out + y * dw + x;
    o0[0] = c0[0] - fma(6.0, he[1], he[0] + he[2]) * (1.0 / 64.0);
    if (x1) o0[1] = c0[1] - fma(6.0, ho[1], ho[0] + ho[2]) * (1.0 / 64.0);
    if (y1) {
      o0[dw] = c0[dw] - (4.0 * (he[1] + he[2])) * (1.0 / 64.0);
      if (x1) o0[dw + 1] = c0[dw + 1] - (4.0 * (ho[1] + ho[2])) * (1.0 / 64.0);
    }
  }
}

__device__ __forceinline__ void pf_lap(const double* __restrict__ cur, const double* __restrict__ s, double* __restrict__ out,
                                       int sw, int sh, int dw, int dh, unsigned msw, int t, int nt) {
  pf_lap_t<false>(cur, s, out, sw, sh, dw, dh, msw, t, nt);
}
// Levels first+1..top and the Laplacian record of one frame from the slot's level-`first` image, by the nt threads of
// the slot (t = 0..nt-1).
// The level-`first` image the stream wrote is in the slot's shared memory, or (G4G, PfParams::g4_global) in place in the
// record, where its Laplacian replaces it.
template <bool G4G>
__device__ __noinline__ void pf_tail(const PfParams& p, unsigned char* lvl_slot, double* rec, int slot, int t, int nt) {
  const int f = p.first, top = p.top;
  const double* g_first = G4G ? rec + p.rec_off[f] : pf_level(p, lvl_slot, f);
  if (G4G) pf_down_t<false>(g_first, pf_level(p, lvl_slot, f + 1), p.w[f], p.h[f], p.w[f + 1], p.h[f + 1], p.magic[f + 1], t, nt);
  else pf_down(g_first, pf_level(p, lvl_slot, f + 1), p.w[f], p.h[f], p.w[f + 1], p.h[f + 1], p.magic[f + 1], t, nt);
  pf_slot_sync(slot, nt);
  // the small levels are a chain of tiny images: one warp walks it while the others write the largest Laplacian
  if (t < 32) {
    for (int l = f + 1; l < top; ++l) {
      pf_down(pf_level(p, lvl_slot, l), pf_level(p, lvl_slot, l + 1), p.w[l], p.h[l], p.w[l + 1], p.h[l + 1], p.magic[l + 1], t, 32);
      __syncwarp();
    }
  }
  if (nt == 32 || t >= 32) {
    const int t2 = nt == 32 ? t : t - 32, nt2 = nt == 32 ? 32 : nt - 32;
    if (G4G)
      pf_lap_t<true>(g_first, pf_level(p, lvl_slot, f + 1), rec + p.rec_off[f], p.w[f + 1], p.h[f + 1], p.w[f], p.h[f],
                     p.magic[f + 1], t2, nt2);
    else
      pf_lap(g_first, pf_level(p, lvl_slot, f + 1), rec + p.rec_off[f], p.w[f + 1], p.h[f + 1], p.w[f], p.h[f],
             p.magic[f + 1], t2, nt2);
  }
  pf_slot_sync(slot, nt);
  for (int l = f + 1; l < top; ++l)
    pf_lap(pf_level(p, lvl_slot, l), pf_level(p, lvl_slot, l + 1), rec + p.rec_off[l], p.w[l + 1], p.h[l + 1], p.w[l], p.h[l],
           p.magic[l + 1], t, nt);
}

// ---- per-lane constants of a strip's warp
// Image borders are reflect-101.  In the fused kernel they cost no instructions in the stream: the lane that holds the
// image's first / last column gets its own *filter weights* (the mirrored taps are folded onto the columns they mirror and
// the neighbour's contribution is weighted zero) and its own shuffle source lanes, prepared once per kernel.
struct PfLane {
  int s_m2, s_m1, s_p1, s_p2;   // level 3: source lanes of columns col-2, col-1, col+1, col+2
  int k;                        // level-4 column the lane emits, -1: none
  // level 0 (8 pixels per lane, w0 = pixels 0..3, w1 = 4..7, wl / wr = the neighbours' w1 / w0):
  //   k0 = dp4a(wl, l0_wl) + dp4a(w0, l0_w0)        interior: (.,.,1,4) (6,4,1,.)     first lane: 0, (6,8,2,.)
  //   k3 = dp4a(w1, l0_w1) + (wr & l0_wr)           interior: (1,4,6,4), byte mask    last lane: (1,4,7,4), 0
  unsigned l0_wl, l0_w0, l0_w1, l0_wr;
  // level 1 (P = columns (0,1), Q = (2,3) as 16-bit pairs, Ql / Pr the neighbours' Q / P), byte pairs for dp2a lo | hi:
  //   g0 = dp2a_lo(Ql, h1_a) + dp2a_hi(P, h1_a) + dp2a_lo(Q, h1_b)
  //   g1 = dp2a_hi(P, h1_c) + dp2a_hi(Q, h1_b) + dp2a_lo(Pr, h1_c)
  unsigned h1_a, h1_b, h1_c;
  int h2_l0, h2_l1, h2_r0;      // level 2: source lanes of columns 2L-2, 2L-1 (q0 / q1 of the left lane) and 2L+2
};

__device__ __forceinline__ void pf_lane_setup(PfLane& ln, int lane, int col, int base, int W3, int first_lane, int last_lane,
                                              int k0, int k1) {
  ln.s_m2 = (reflect101(col - 2, W3) - base) & 31; ln.s_m1 = (reflect101(col - 1, W3) - base) & 31;
  ln.s_p1 = (reflect101(col + 1, W3) - base) & 31; ln.s_p2 = (reflect101(col + 2, W3) - base) & 31;
  ln.k = (!(col & 1) && (col >> 1) >= k0 && (col >> 1) < k1) ? (col >> 1) : -1;
  const bool first = lane == first_lane, last = lane == last_lane;
  ln.l0_wl = first ? 0u : 0x04010000u;  ln.l0_w0 = first ? 0x00020806u : 0x00010406u;
  ln.l0_w1 = last ? 0x04070401u : 0x04060401u;  ln.l0_wr = last ? 0u : 0xffu;
  ln.h1_a = first ? 0x08060000u : 0x04060401u;
  ln.h1_b = (last ? 0x04070000u : 0x04060000u) | (first ? 0x0002u : 0x0001u);
  ln.h1_c = last ? 0x04010000u : 0x04010001u;
  ln.h2_l0 = first ? lane + 1 : (lane ? lane - 1 : 0);
  ln.h2_l1 = first ? lane : (lane ? lane - 1 : 0);
  ln.h2_r0 = last ? lane : (lane < 31 ? lane + 1 : 31);
}

// pu_block for the fused kernel: the same sums, image borders by the lane's weights (EDGE) or none at all
template <bool EDGE, bool MASK_TAP>
__device__ __forceinline__ void pf_block(PuState& s, const uint2 w[PU_ROWS], const PfLane& ln, unsigned& out3) {
  unsigned ha[PU_ROWS], hb[PU_ROWS];
#pragma unroll
  for (int i = 0; i < PU_ROWS; ++i) {
    const unsigned w0 = w[i].x, w1 = w[i].y;
    const unsigned wl = __shfl_up_sync(0xffffffffu, w1, 1);
    const unsigned wr = __shfl_down_sync(0xffffffffu, w0, 1);
    // MASK_TAP: a single tap of weight 1 is a byte mask (ALU pipe) feeding the accumulator, not a dot product of its own
    // -- the integer-multiply pipe is the busier one: VGA 0.580 -> 0.559 ms, 1080p 0.590 -> 0.579 (two-stage rings); the
    // four-stage configuration at 96 registers (720p) is 2 % faster with the dot product (r03r)
    const unsigned k0 = __dp4a(wl, EDGE ? ln.l0_wl : 0x04010000u, __dp4a(w0, EDGE ? ln.l0_w0 : 0x00010406u, 0u));
    const unsigned k2 = __dp4a(w0, 0x04010000u, __dp4a(w1, 0x00010406u, 0u));
    unsigned k1, k3;
    if (MASK_TAP) {
      k1 = __dp4a(w0, 0x04060401u, w1 & 0xffu);
      k3 = __dp4a(w1, EDGE ? ln.l0_w1 : 0x04060401u, wr & (EDGE ? ln.l0_wr : 0xffu));
    } else {
      k1 = __dp4a(w0, 0x04060401u, __dp4a(w1, 0x00000001u, 0u));
      k3 = __dp4a(w1, EDGE ? ln.l0_w1 : 0x04060401u, __dp4a(wr, EDGE ? (ln.l0_wr & 1u) : 1u, 0u));
    }
    ha[i] = k0 + (k1 << 16);       // (a byte permute instead of the shift-add, i.e. ALU instead of FMA pipe: +-1 %, r03c)
    hb[i] = k2 + (k3 << 16);
  }
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
  for (int i = 0; i < 4; ++i) {
    const unsigned Ql = __shfl_up_sync(0xffffffffu, Q[i], 1);
    const unsigned Pr = __shfl_down_sync(0xffffffffu, P[i], 1);
    const unsigned a = EDGE ? ln.h1_a : 0x04060401u, b = EDGE ? ln.h1_b : 0x04060001u, c = EDGE ? ln.h1_c : 0x04010001u;
    n0[i] = __dp2a_lo(Ql, a, __dp2a_hi(P[i], a, __dp2a_lo(Q[i], b, 0u)));
    n1[i] = __dp2a_hi(P[i], c, __dp2a_hi(Q[i], b, __dp2a_lo(Pr, c, 0u)));
  }
  unsigned q0[2], q1[2];
  q0[0] = v5(s.g0[0], s.g0[1], s.g0[2], n0[0], n0[1]);
  q0[1] = v5(s.g0[2], n0[0], n0[1], n0[2], n0[3]);
  q1[0] = v5(s.g1[0], s.g1[1], s.g1[2], n1[0], n1[1]);
  q1[1] = v5(s.g1[2], n1[0], n1[1], n1[2], n1[3]);
#pragma unroll
  for (int i = 0; i < 3; ++i) { s.g0[i] = n0[1 + i]; s.g1[i] = n1[1 + i]; }
  unsigned m[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    unsigned ql0, ql1, qr0;
    if (EDGE) {
      ql0 = __shfl_sync(0xffffffffu, q0[i], ln.h2_l0); ql1 = __shfl_sync(0xffffffffu, q1[i], ln.h2_l1);
      qr0 = __shfl_sync(0xffffffffu, q0[i], ln.h2_r0);
    } else {
      ql0 = __shfl_up_sync(0xffffffffu, q0[i], 1); ql1 = __shfl_up_sync(0xffffffffu, q1[i], 1);
      qr0 = __shfl_down_sync(0xffffffffu, q0[i], 1);
    }
    m[i] = v5(ql0, ql1, q0[i], q1[i], qr0);
  }
  out3 = v5(s.c[0], s.c[1], s.c[2], m[0], m[1]);
  s.c[0] = s.c[2]; s.c[1] = m[0]; s.c[2] = m[1];
}

// bottom border of even-sized levels (pu_flush with the lane's weights)
template <bool EDGE>
__device__ __forceinline__ unsigned pf_flush(const PuState& s, const PfLane& ln) {
  const unsigned P = v5(s.a[0], s.a[1], s.a[2], s.a[3], s.a[2]);
  const unsigned Q = v5(s.b[0], s.b[1], s.b[2], s.b[3], s.b[2]);
  const unsigned Ql = __shfl_up_sync(0xffffffffu, Q, 1);
  const unsigned Pr = __shfl_down_sync(0xffffffffu, P, 1);
  const unsigned a = EDGE ? ln.h1_a : 0x04060401u, b = EDGE ? ln.h1_b : 0x04060001u, c = EDGE ? ln.h1_c : 0x04010001u;
  const unsigned n0 = __dp2a_lo(Ql, a, __dp2a_hi(P, a, __dp2a_lo(Q, b, 0u)));
  const unsigned n1 = __dp2a_hi(P, c, __dp2a_hi(Q, b, __dp2a_lo(Pr, c, 0u)));
  const unsigned q0 = v5(s.g0[0], s.g0[1], s.g0[2], n0, s.g0[2]);
  const unsigned q1 = v5(s.g1[0], s.g1[1], s.g1[2], n1, s.g1[2]);
  const unsigned ql0 = __shfl_sync(0xffffffffu, q0, ln.h2_l0), ql1 = __shfl_sync(0xffffffffu, q1, ln.h2_l1);
  const unsigned qr0 = __shfl_sync(0xffffffffu, q0, ln.h2_r0);
  const unsigned m = v5(ql0, ql1, q0, q1, qr0);
  return v5(s.c[0], s.c[1], s.c[2], m, s.c[2]);
}

// ---- a warp's TMA ring: S stages of 8 rows x 256 bytes, one mbarrier per stage.  Everything here is warp-uniform.
struct PfRing {
  unsigned ring_s, mbar;   // shared-space addresses of stage 0 and of its mbarrier
  unsigned lds;            // shared-space address of the lane's 8 bytes in row 0 of stage 0
  unsigned soff;           // byte offset of the stage the next block is consumed from (stage * PU_STAGE_BYTES)
  unsigned par;            // parity awaited on the stages of the current round of the ring
  int fy, fz;              // next fetch: the box whose first row is fy of source frame fz (-1: nothing left)
  int x0;                  // first pixel column of the strip's window
};
__device__ __forceinline__ bool pf_elect() {
  unsigned ok;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok));
  return ok != 0;
}
// the box at rows fy.. of frame fz into the stage at byte offset soff; all lanes are done with what the stage held
__device__ __forceinline__ void pf_issue(const PfRing& r, const CUtensorMap* tmap, unsigned soff) {
  __syncwarp();
  if (pf_elect()) {
    const unsigned bar = r.mbar + (soff >> 8);          // 8 bytes of barrier per 2048 bytes of stage
    pu_mbar_expect_tx(bar, PU_STAGE_BYTES);
    pu_tma_load(r.ring_s + soff, tmap, r.x0, r.fy, r.fz, bar);
  }
}
// Fetch order of a frame: the mirrored lead-in rows 9..16 and 1..8 (read back bottom-up: rows 16..9 are rows -16..-9 of the
// reflect-101 extension, 8..1 rows -8..-1), then rows 0.., 8.., ..., H-8..; then the next frame of the slot.
__device__ __forceinline__ void pf_fetch_any(PfRing& r, const CUtensorMap* tmap, unsigned soff, int H, int next_sframe) {
  if (r.fz >= 0) pf_issue(r, tmap, soff);
  r.fy = r.fy == 9 ? 1 : (r.fy == 1 ? 0 : r.fy + 8);
  if (r.fy == H) { r.fy = 9; r.fz = next_sframe; }
}
__device__ __forceinline__ uint2 pf_lds8(unsigned addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

// level 4 from the level-3 rows of the lane's column.  A level-3 row is filtered across the lanes in 64-bit integers and
// continues in float64: every value is an integer below 2^41, so all sums are exact (= the 64-bit integer ones) and the
// only rounding is the final scale by 2^-32 / 255.  Output row R = h[2R-2] + 4 h[2R-1] + 6 h[2R] + 4 h[2R+1] + h[2R+2] is
// built in accumulators as the rows arrive: `cur` collects output r/2 (r even), `nxt` the one after.
struct PfL4 {
  double cur, nxt;
};
__device__ __forceinline__ double pf_l3_hfilter(unsigned o, const PfLane& ln) {
  const unsigned a = __shfl_sync(0xffffffffu, o, ln.s_m2), b = __shfl_sync(0xffffffffu, o, ln.s_m1);
  const unsigned d = __shfl_sync(0xffffffffu, o, ln.s_p1), e = __shfl_sync(0xffffffffu, o, ln.s_p2);
  const unsigned long long hq = (unsigned long long)a + e + 4ull * ((unsigned long long)b + d) + 6ull * o;
  return (double)(long long)hq;
}
// rows r >= 3 (no border left): odd rows add 4 h to both outputs; even rows complete output r/2 - 1
template <bool ODD>
__device__ __forceinline__ void pf_l4_row(PfL4& a, double h, int r, double* __restrict__ g4k, int W4, double g_scale) {
  if (ODD) {
    a.cur = fma(4.0, h, a.cur);
    a.nxt = fma(4.0, h, a.nxt);
  } else {
    if (g4k) g4k[((r - 2) >> 1) * W4] = (a.cur + h) * g_scale;
    a.cur = fma(6.0, h, a.nxt);
    a.nxt = h;
  }
}
// any row, including the top border (rows -2, -1 are rows 2, 1)
__device__ __forceinline__ void pf_l4_row_any(PfL4& a, double h, int r, double* __restrict__ g4k, int W4, double g_scale) {
  if (r == 0) { a.cur = 6.0 * h; a.nxt = h; }
  else if (r == 1) { a.cur = fma(8.0, h, a.cur); a.nxt = fma(4.0, h, a.nxt); }
  else if (r == 2) {
    if (g4k) g4k[0] = fma(2.0, h, a.cur) * g_scale;
    a.cur = fma(6.0, h, a.nxt);
    a.nxt = h;
  } else if (r & 1) pf_l4_row<true>(a, h, r, g4k, W4, g_scale);
  else pf_l4_row<false>(a, h, r, g4k, W4, g_scale);
}

// One frame of one strip: blocks -2 .. nblk-1 of 8 rows.  The steady state (`hot`: blocks 4 .. nblk-S, in pairs) has no
// special case left in it -- no mirrored rows, no level-4 border, fetches well inside the frame, and with STATIC (two stages, an
// even number of blocks per frame) the ring stage of every block is a compile-time constant; the few blocks at either end of
// the frame run one general (`cold`) copy of the body.
template <int S, bool STATIC>
__device__ __forceinline__ void pf_run_frame(const PfParams& p, const CUtensorMap* tmap, PfRing& rg, int next_sframe,
                                             double* __restrict__ g4k, const PfLane& ln) {
  const int nblk = p.H >> 3, H = p.H;
  PuState s;
  pu_clear(s);
  PfL4 l4;
  l4.cur = 0; l4.nxt = 0;
  const int W4 = p.w[p.first];
  const double g_scale = p.g_scale;
  auto cold = [&](int b) {
    const unsigned prev = rg.soff == 0 ? (S - 1) * PU_STAGE_BYTES : rg.soff - PU_STAGE_BYTES;
    pf_fetch_any(rg, tmap, prev, H, next_sframe);    // the stage consumed by the previous block takes block b+S-1
    pu_mbar_wait(rg.mbar + (rg.soff >> 8), rg.par);
    const bool mirrored = b < 0;
    const unsigned src = rg.lds + rg.soff + (mirrored ? (PU_ROWS - 1) * 256 : 0);
    const int rstep = mirrored ? -256 : 256;
    uint2 w[PU_ROWS];
#pragma unroll
    for (int i = 0; i < PU_ROWS; ++i) w[i] = pf_lds8(src + i * rstep);
    rg.soff += PU_STAGE_BYTES;
    if (rg.soff == S * PU_STAGE_BYTES) { rg.soff = 0; rg.par ^= 1u; }
    unsigned o;
    pf_block<true, S == 2>(s, w, ln, o);
    if (b >= 1) pf_l4_row_any(l4, pf_l3_hfilter(o, ln), b - 1, g4k, W4, g_scale);
  };
  // hot pairs start at an even block (row b-1 odd, then row b even); the last one fetches block nblk-2 at most, so that
  // the step past the frame's last box (next frame, mirrored rows) is the general body's
  const int hot_end = nblk - S - ((nblk - S) & 1);
  const int lead_end = nblk < 4 ? nblk : 4;
  int b = -2;
#pragma unroll 1
  for (; b < lead_end; ++b) cold(b);
  if (hot_end > 4) {
#pragma unroll 1
    for (; b < hot_end; b += 2) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        // STATIC: block b+j is consumed from stage j and its predecessor's stage, 1 - j, takes block b+j+1
        const unsigned soff = STATIC ? j * PU_STAGE_BYTES : rg.soff;
        const unsigned prev = STATIC ? (1 - j) * PU_STAGE_BYTES : (soff == 0 ? (S - 1) * PU_STAGE_BYTES : soff - PU_STAGE_BYTES);
        pf_issue(rg, tmap, prev);
        rg.fy += 8;
        pu_mbar_wait(rg.mbar + (soff >> 8), rg.par);
        const unsigned src = rg.lds + soff;
        uint2 w[PU_ROWS];
#pragma unroll
        for (int i = 0; i < PU_ROWS; ++i) w[i] = pf_lds8(src + i * 256);
        if (STATIC) {
          if (j == 1) rg.par ^= 1u;
        } else {
          rg.soff += PU_STAGE_BYTES;
          if (rg.soff == S * PU_STAGE_BYTES) { rg.soff = 0; rg.par ^= 1u; }
        }
        unsigned o;
        pf_block<true, S == 2>(s, w, ln, o);
        const double h = pf_l3_hfilter(o, ln);
        if (j == 0) pf_l4_row<true>(l4, h, b - 1, g4k, W4, g_scale);
        else pf_l4_row<false>(l4, h, b, g4k, W4, g_scale);
      }
    }
  }
#pragma unroll 1
  for (; b < nblk; ++b) cold(b);
  // The last level-3 row (r = H3 - 1) comes out of the flush; it also closes level 4, whose last output row needs the rows
  // below the image: H3 odd (r even):  rows r+1, r+2 are rows r-1, r-2:  out = 6 h + 2 (h[r-2] + 4 h[r-1]) = 6 h + 2 nxt;
  //                  H3 even (r odd):  row r+1 is row r-1:               out = cur + 4 h + h[r-1]          = cur + 4 h + nxt.
  const int r = nblk - 1;
  const double h = pf_l3_hfilter(pf_flush<true>(s, ln), ln);
  if (g4k) {
    if (r & 1) g4k[(r >> 1) * W4] = (fma(4.0, h, l4.cur) + l4.nxt) * g_scale;
    else {
      g4k[((r - 2) >> 1) * W4] = (r == 2 ? fma(2.0, h, l4.cur) : l4.cur + h) * g_scale;
      g4k[(r >> 1) * W4] = fma(6.0, h, l4.nxt + l4.nxt) * g_scale;
    }
  }
}

// the frames of one slot's strip warp, one after the other
template <int S, bool G4G>
__device__ __forceinline__ void pf_frames(const PfParams& p, const CUtensorMap& tmap, PfRing& rg, const PfLane& ln,
                                          unsigned char* lvl_slot, int slot, int strip, int lane) {
  const int nt = p.n_strips * 32, t = strip * 32 + lane;
  const long long stride = (long long)gridDim.x * p.frames_per_cta;
  long long frame = (long long)blockIdx.x * p.frames_per_cta + slot;
  rg.fy = 9;
  rg.fz = frame < p.n_frames ? (int)pu_source_frame(frame, p.seg_len, p.seg_stride, p.seg_first) : -1;
  {   // the first S-1 blocks of the slot's first frame
    const int next0 = frame + stride < p.n_frames ? (int)pu_source_frame(frame + stride, p.seg_len, p.seg_stride, p.seg_first) : -1;
#pragma unroll
    for (int i = 0; i < S - 1; ++i) pf_fetch_any(rg, &tmap, i * PU_STAGE_BYTES, p.H, next0);
  }
  // two stages and an even number of 8-row blocks per frame (+ the 2 lead-in blocks): every frame starts on stage 0
  const bool static_ring = S == 2 && !((p.H >> 3) & 1);
  double* g4_smem = pf_level(p, lvl_slot, p.first);
  for (; frame < p.n_frames; frame += stride) {
    const int next_sframe = frame + stride < p.n_frames ? (int)pu_source_frame(frame + stride, p.seg_len, p.seg_stride, p.seg_first) : -1;
    double* rec = p.lap_out + frame * p.record_len;
    double* g4 = G4G ? rec + p.rec_off[p.first] : g4_smem;
    double* g4k = ln.k >= 0 ? g4 + ln.k : nullptr;               // the lane's level-4 column, if it emits one
    if (S == 2 && static_ring) pf_run_frame<S, S == 2>(p, &tmap, rg, next_sframe, g4k, ln);
    else pf_run_frame<S, false>(p, &tmap, rg, next_sframe, g4k, ln);
    // The barrier that completes the level-4 image also re-aligns the strips of the frame (they share halo columns:
    // left alone they drift apart until a halo sector has left L2 before the neighbour asks for it, profiles/r01g).
    pf_slot_sync(slot, nt);
    pf_tail<G4G>(p, lvl_slot, rec, slot, t, nt);
    pf_slot_sync(slot, nt);                                    // the level images are free for the next frame
  }
}

// MAXW warps per CTA; registers are allocated to warps four at a time: 65536 / (32 * MAXW rounded up to 4), rounded down
// to the allocation unit of 8 per thread -- 96 for 18 warps, 80 for 21 and 24
template <int S, int MAXW>
__global__ void __maxnreg__((65536 / (32 * ((MAXW + 3) & ~3))) & ~7)
    pyramid_u8_fused_kernel(const __grid_constant__ PfParams p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem[];
  // the warp index through a shuffle: the compiler then knows that everything derived from it (strip, slot, ring and
  // barrier addresses, frame numbers, TMA coordinates) is warp-uniform
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int slot = warp / p.n_strips, strip = warp - slot * p.n_strips;
  const bool left = strip == 0, right = strip == p.n_strips - 1;
  const int base = p.strip_base[strip];
  PfLane ln;
  // lanes holding the image's first / last level-3 column (-1: the strip lacks that border)
  pf_lane_setup(ln, lane, base + lane, base, p.W3, left ? 0 : -1, right ? p.W3 - 1 - base : -1, p.strip_k0[strip],
                p.strip_k1[strip]);
  const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem);
  PfRing rg;
  rg.ring_s = smem_s + warp * S * PU_STAGE_BYTES;
  rg.lds = rg.ring_s + lane * 8;
  rg.mbar = smem_s + p.mbar_base + warp * S * 8;
  rg.soff = 0; rg.par = 0;
  rg.x0 = 8 * base;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < S; ++i) pu_mbar_init(rg.mbar + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  unsigned char* lvl_slot = smem + p.lvl_base + (size_t)slot * p.lvl_stride;
  // Level 4 of a frame goes to the slot's shared memory, or -- wide frames, whose level-4 image (65 KB at 1080p) would leave
  // room for one frame slot per SM -- straight into the frame's record, where the tail turns it into its Laplacian in
  // place.  Two instantiations, so that every access keeps its address space (a launch runs one of them).
  if (p.g4_global) pf_frames<S, true>(p, tmap, rg, ln, lvl_slot, slot, strip, lane);
  else pf_frames<S, false>(p, tmap, rg, ln, lvl_slot, slot, strip, lane);
}

// ==================================================================================================== host side
// The integer paths need: uint8 frames, four front levels, W and H multiples of 8 (all three decimated sizes even),
// at least 17 rows (the mirrored lead-in) and 8-byte aligned rows.
bool pu_supported(const void* frames, int W, int H, int skip) {
  return skip == 4 && W % 8 == 0 && H % 8 == 0 && H >= 24 && W >= 16 && ((uintptr_t)frames % 8) == 0;
}

struct PfPlan {
  PfParams p;
  int warps, smem, stages, maxw;
};
// Strip geometry, shared-memory layout and kernel configuration of a fused launch; false if it does not fit.
static bool pf_plan(rm_handle* h, int W, int H, PfPlan& pl) {
  PfParams& p = pl.p;
  memset(&p, 0, sizeof(p));
  const int L = h->p.pyramid_levels, s = h->p.skip_levels_at_top;
  if (W % 16 || L > RM_MAX_LEVELS || s != 4 || L - 1 <= s) return false;
  LevelGeom g = make_geom(W, H, L);
  RecordGeom rec = make_record(g, s);
  p.W = W; p.H = H; p.W3 = W / 8; p.H3 = H / 8;
  p.first = s; p.top = L - 1; p.record_len = rec.len;
  p.g_scale = 1.0 / 255;
  for (int l = 0; l < s; ++l) p.g_scale *= 1.0 / 256.0;
  for (int l = 0; l < L; ++l) {
    p.w[l] = g.w[l]; p.h[l] = g.h[l]; p.rec_off[l] = rec.off[l];
    p.magic[l] = g.w[l] > 1 ? (unsigned)((0x100000000ull + g.w[l] - 1) / g.w[l]) : 0u;
    if (l >= s && (long long)g.w[l] * g.h[l] >= 65536) return false;      // pf_div
  }
  // level images of a slot: levels first..top, or first+1..top with level `first` in the record (g4_global)
  auto layout = [&](bool g4_global) {
    int bytes = 0;
    for (int l = s; l < L; ++l) {
      if (l == s && g4_global) { p.lvl_off[l] = 0; continue; }
      p.lvl_off[l] = bytes; bytes += g.w[l] * g.h[l] * 8;
    }
    return (bytes + 15) & ~15;
  };
  // strips in level-4 columns
  const int K = g.w[s];
  int n = 0, k0 = 0;
  while (k0 < K) {
    if (n == 16) return false;
    const int base = n == 0 ? 0 : 2 * k0 - 4;
    int k1;
    if (base + 31 >= p.W3 - 1) k1 = K;                          // the window reaches the right border: the rest
    else k1 = (base + 28) / 2 + 1;                              // centre lane <= 28
    if (k1 > K) k1 = K;
    p.strip_base[n] = base; p.strip_k0[n] = k0; p.strip_k1[n] = k1;
    k0 = k1; ++n;
  }
  p.n_strips = n;
  // ring depth / warps per CTA: as many frame slots as shared memory, the register file and 15 named barriers allow
  const int cfg[3][2] = {{4, 18}, {2, 18}, {2, 24}};            // (stages, max warps)
  auto slots = [&](int S, int maxw, bool g4_global) {
    int want = maxw / n;
    if (want > 15) want = 15;
    const int per_slot = n * S * PU_STAGE_BYTES + layout(g4_global) + n * S * 8;
    const int fit = (h->smem_optin - 256) / per_slot;
    return want < fit ? want : fit;
  };
  // measured (profiles/r03_pyramid_configs.txt): frames of up to three strips (VGA: 3 x 8 slots) want the most warps, 2 stages
  // x 24 warps at 80 registers (0.558 against 0.625 ms per 8192 VGA frames with 2 x 18 at 96 and 0.676 with 4 x 18); wider
  // frames, whose level images leave room for two or three slots, want the 96 registers, and the shallow ring leaves them
  // the shared memory for a third slot (720p: 0.465 against 0.485 ms per 2048 frames with 4 x 18) or, with level 4 in the
  // record, for a second one and the L1 cache the record is then read through (1080p: 0.589 against 0.609 with 2 x 24 and
  // 0.714 with 4 x 18 per 1024 frames).
  int pick = h->pyramid_cfg >= 1 && h->pyramid_cfg <= 3 ? h->pyramid_cfg - 1 : (n <= 3 ? 2 : 1);
  for (int tries = 0; tries < 3; ++tries, pick = (pick + 1) % 3) {
    const int S = cfg[pick][0], maxw = cfg[pick][1];
    const int in_smem = slots(S, maxw, false), in_rec = slots(S, maxw, true);
    p.g4_global = h->pyramid_g4 == 2 || (h->pyramid_g4 == 0 && in_smem < 2 && in_rec > in_smem);
    const int fpc = p.g4_global ? in_rec : in_smem;
    const int lvl_bytes = layout(p.g4_global != 0);
    if (fpc < 1) continue;
    p.frames_per_cta = fpc;
    pl.warps = fpc * n; pl.stages = S; pl.maxw = maxw;
    p.lvl_base = pl.warps * S * PU_STAGE_BYTES; p.lvl_stride = lvl_bytes;
    p.mbar_base = p.lvl_base + fpc * lvl_bytes;
    pl.smem = p.mbar_base + pl.warps * S * 8;
    return true;
  }
  return false;
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

// 1: the fused TMA kernel can take these frames (rows and frames 16-byte multiples, the level images fit a slot);
// 0: fallback (level 3 through HBM + pyramid_tail_kernel).  Option "pyramid_mode" = 0 forces the fallback.
int pu_best_mode(rm_handle* h, const void* frames, int W, int H) {
  PfPlan pl;
  if (h->pyramid_mode == 0 || ((uintptr_t)frames % 16) || !pu_encoder() || !pf_plan(h, W, H, pl)) return 0;
  return 1;
}

template <int S, int MAXW>
static int32_t pf_launch_cfg(rm_handle* h, const PfPlan& pl, const CUtensorMap& map, long long ctas, cudaStream_t st) {
  RM_CUDA(h, cudaFuncSetAttribute(pyramid_u8_fused_kernel<S, MAXW>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem));
  RM_PROF(h, st, "pyramid_u8_fused_kernel");
  pyramid_u8_fused_kernel<S, MAXW><<<(unsigned)ctas, pl.warps * 32, pl.smem, st>>>(pl.p, map);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}

int32_t pu_launch_fused(rm_handle* h, const uint8_t* frames, double* lap_out, long long n_frames, long long seg_len,
                        long long seg_stride, long long seg_first, int W, int H, cudaStream_t st) {
  PfPlan pl;
  if (!pf_plan(h, W, H, pl)) return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: frames do not fit the fused pyramid kernel", __func__);
  PfParams& p = pl.p;
  p.lap_out = lap_out; p.n_frames = n_frames;
  p.seg_len = seg_len; p.seg_stride = seg_stride; p.seg_first = seg_first;
  // every source frame the launch can touch, as one (W, H, frames) uint8 tensor; box = 256 columns x PU_ROWS rows
  const long long last = n_frames - 1;
  const long long n_src = (last / seg_len) * seg_stride + seg_first + last % seg_len + 1;
  if (n_frames >= (1ll << 31) || seg_len >= (1ll << 31) || seg_len < 1 || n_src >= (1ll << 31))
    return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: more than 2^31 frames", __func__);
  CUtensorMap map;
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
  long long ctas = (n_frames + p.frames_per_cta - 1) / p.frames_per_cta;
  if (ctas > h->sm_count) ctas = h->sm_count;
  if (pl.stages == 4) return pf_launch_cfg<4, 18>(h, pl, map, ctas, st);
  if (pl.maxw == 18) return pf_launch_cfg<2, 18>(h, pl, map, ctas, st);
  return pf_launch_cfg<2, 24>(h, pl, map, ctas, st);
}

int32_t pu_launch_front(rm_handle* h, const uint8_t* frames, int bgr, uint32_t* g3, long long n_frames, long long seg_len,
                        long long seg_stride, long long seg_first, int W, int H, cudaStream_t st) {
  PuParams p;
  memset(&p, 0, sizeof(p));
  p.frames = frames; p.g3 = g3; p.n_frames = n_frames; p.frame_elems = (long long)W * H * (bgr ? 3 : 1);
  p.seg_len = seg_len; p.seg_stride = seg_stride; p.seg_first = seg_first;
  p.W = W; p.H = H; p.W3 = W / 8; p.H3 = H / 8;
  const int cap = 32 - 2 * PU_HALO_LANES;                       // payload lanes of an interior strip
  p.n_strips = (p.W3 + cap - 1) / cap;
  p.cols_per_strip = (p.W3 + p.n_strips - 1) / p.n_strips;
  if (p.n_strips > 16) return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: frame wider than 3584 pixels", __func__);
  if (n_frames >= (1ll << 31) || seg_len >= (1ll << 31) || seg_len < 1)
    return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: more than 2^31 frames", __func__);
  const int ring_per_warp = (bgr ? PU_BGR_STAGES * 3 : PU_FRONT_STAGES) * PU_STAGE_BYTES;
  int max_warps = PU_FRONT_WARPS;
  if (max_warps * ring_per_warp > h->smem_optin) max_warps = h->smem_optin / ring_per_warp;
  p.frames_per_cta = max_warps / p.n_strips;
  if (p.frames_per_cta > 15) p.frames_per_cta = 15;             // named barriers 1..15
  if (p.frames_per_cta < 1) return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: frame too wide for the BGR ring", __func__);
  const int warps = p.frames_per_cta * p.n_strips;
  const int smem = warps * ring_per_warp;
  void (*kern)(const PuParams);
  if (bgr)
    kern = W == 640 ? pyramid_front_u8_kernel<640, true> : W == 1280 ? pyramid_front_u8_kernel<1280, true>
           : W == 1920 ? pyramid_front_u8_kernel<1920, true> : pyramid_front_u8_kernel<0, true>;
  else
    kern = W == 640 ? pyramid_front_u8_kernel<640, false> : W == 1280 ? pyramid_front_u8_kernel<1280, false>
           : W == 1920 ? pyramid_front_u8_kernel<1920, false>
           : W == 320 ? pyramid_front_u8_kernel<320, false> : pyramid_front_u8_kernel<0, false>;
  RM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long ctas = (n_frames + p.frames_per_cta - 1) / p.frames_per_cta;
  if (ctas > h->sm_count) ctas = h->sm_count;
  RM_PROF(h, st, bgr ? "pyramid_front_bgr_kernel" : "pyramid_front_u8_kernel");
  kern<<<(unsigned)ctas, warps * 32, smem, st>>>(p);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}
