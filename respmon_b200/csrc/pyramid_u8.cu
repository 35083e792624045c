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
__device__ __forceinline__ void pf_down(const double* __restrict__ s, double* __restrict__ d, int sw, int sh, int dw, int dh,
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
// L_l = G_l - pyrUp(G_{l+1}) (pyramid.py:24-26), one thread per source pixel = 2x2 outputs from its 3x3 neighbourhood;
// per output the operations of up_taps / up_combine (pyr_core.h) in their order:
//   even index 2i: s[refl(i-1)] + s[min(i+1,n-1)], then fma(6, s[i], .);   odd index 2i+1: 4 * (s[i] + s[min(i+1,n-1)])
__device__ __forceinline__ void pf_lap(const double* __restrict__ cur, const double* __restrict__ s, double* __restrict__ out,
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

// Levels first+1..top and the Laplacian record of one frame from the slot's level-`first` image, by the nt threads of
// the slot (t = 0..nt-1).
__device__ __noinline__ void pf_tail(const PfParams& p, unsigned char* lvl_slot, double* __restrict__ rec, int slot, int t,
                                     int nt) {
  const int f = p.first, top = p.top;
  pf_down(pf_level(p, lvl_slot, f), pf_level(p, lvl_slot, f + 1), p.w[f], p.h[f], p.w[f + 1], p.h[f + 1], p.magic[f + 1], t, nt);
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
    pf_lap(pf_level(p, lvl_slot, f), pf_level(p, lvl_slot, f + 1), rec + p.rec_off[f], p.w[f + 1], p.h[f + 1], p.w[f], p.h[f],
           p.magic[f + 1], t2, nt2);
  }
  pf_slot_sync(slot, nt);
  for (int l = f + 1; l < top; ++l)
    pf_lap(pf_level(p, lvl_slot, l), pf_level(p, lvl_slot, l + 1), rec + p.rec_off[l], p.w[l + 1], p.h[l + 1], p.w[l], p.h[l],
           p.magic[l + 1], t, nt);
}

struct PfLane {       // what a lane does with its level-3 / level-4 column
  int s_m2, s_m1, s_p1, s_p2;   // source lanes of the level-3 columns col-2, col-1, col+1, col+2 (reflect-101 at the image border)
  int k;                        // level-4 column the lane emits, -1: none
};

// S-stage TMA ring: `stage` is the ring slot of the next block to consume, `phase` the parity awaited per slot.
// Two instantiations: interior strips (no border code at all) and edge strips, whose image-border fix-ups in pu_block
// are predicated on the lane (first_lane / last_lane = -1 where the strip lacks that border) instead of compiled per
// kind of edge.  Three or four variants per CTA (left, interior, right, both) times an unrolled loop plus two peeled
// copies for the mirrored blocks did not fit the 32 KB instruction cache level: 1.5 no_instruction stalls per issue,
// 0.774 -> 0.708 ms per 8192 VGA frames with a single variant (r02g, r02h); the interior variant is what 720p / 1080p
// frames (4 of 6, 7 of 9 strips) mostly run.
template <int S, bool EDGE>
__device__ __forceinline__ void pf_run_frame(const PfParams& p, const CUtensorMap* tmap,
                                             int sframe, int next_sframe, unsigned ring_s, const unsigned char* my,
                                             unsigned mbar, unsigned& phase, int& stage, double* __restrict__ g4,
                                             const PfLane& ln, int lane, int x0, int first_lane, int last_lane) {
  const int nblk = p.H >> 3;
  // block b (-2 .. nblk-1) of this frame, then blocks -2 .. of the next one: the ring never drains between frames
  auto issue = [&](int b, int st) {
    int z = sframe;
    if (b >= nblk) { b -= nblk + 2; z = next_sframe; }
    if (z < 0) return;
    __syncwarp();                                              // every lane is done with the stage's previous rows
    if (lane == 0) {
      pu_mbar_expect_tx(mbar + 8 * st, PU_STAGE_BYTES);
      // rows above the frame are the mirrored rows 16..9 / 8..1: fetched in frame order, read back bottom-up
      pu_tma_load(ring_s + st * PU_STAGE_BYTES, tmap, x0, b >= 0 ? 8 * b : -8 * b - 7, z, mbar + 8 * st);
    }
  };
  PuState s;
  pu_clear(s);
  unsigned long long hz0 = 0, hz1 = 0, hz2 = 0, hz3 = 0;       // horizontally filtered level-3 rows r-4 .. r-1
  const int W4 = p.w[p.first];
  // one level-3 row (r = 0 .. H3-1) of the lane's column: filter across the lanes, every second row emit level 4
  auto level3_row = [&](unsigned o, int r) {
    const unsigned a = __shfl_sync(0xffffffffu, o, ln.s_m2), b = __shfl_sync(0xffffffffu, o, ln.s_m1);
    const unsigned d = __shfl_sync(0xffffffffu, o, ln.s_p1), e = __shfl_sync(0xffffffffu, o, ln.s_p2);
    const unsigned long long hz = (unsigned long long)a + e + 4ull * ((unsigned long long)b + d) + 6ull * o;
    if (!(r & 1) && r >= 2) {
      const unsigned long long v = r == 2 ? v5q(hz, hz3, hz2, hz3, hz) : v5q(hz0, hz1, hz2, hz3, hz);
      if (ln.k >= 0) g4[((r - 2) >> 1) * W4 + ln.k] = (double)v * p.g_scale;
    }
    hz0 = hz1; hz1 = hz2; hz2 = hz3; hz3 = hz;
  };
  auto step = [&](int b, bool mirrored) {
    // the stage consumed S-1 steps ago (block b-1) is free again: fetch block b+S-1 into it
    issue(b + S - 1, stage == 0 ? S - 1 : stage - 1);
    pu_mbar_wait(mbar + 8 * stage, (phase >> stage) & 1u);
    phase ^= 1u << stage;
    const unsigned char* src = my + stage * PU_STAGE_BYTES;
    stage = stage + 1 == S ? 0 : stage + 1;
    uint2 w[PU_ROWS];
#pragma unroll
    for (int i = 0; i < PU_ROWS; ++i) w[i] = *reinterpret_cast<const uint2*>(src + (mirrored ? PU_ROWS - 1 - i : i) * 256);
    unsigned o;
    pu_block<EDGE, EDGE>(s, w, lane, first_lane, last_lane, o);
    if (b >= 1) level3_row(o, b - 1);
  };
#pragma unroll 1
  for (int b = -2; b < 0; ++b) step(b, true);                  // one cold copy of the body for the two mirrored blocks
#pragma unroll 2
  for (int b = 0; b < nblk; ++b) step(b, false);
  const int H3 = nblk;
  level3_row(pu_flush<EDGE, EDGE>(s, lane, first_lane, last_lane), H3 - 1);
  // bottom border of level 4: hz3 = row H3-1, hz2 = H3-2, ...
  const unsigned long long v = (H3 & 1) ? v5q(hz1, hz2, hz3, hz2, hz1) : v5q(hz0, hz1, hz2, hz3, hz2);
  if (ln.k >= 0) g4[((H3 - 1) >> 1) * W4 + ln.k] = (double)v * p.g_scale;
}

// MAXW warps per CTA; registers are allocated to warps four at a time: 65536 / (32 * MAXW rounded up to 4), rounded down
// to the allocation unit of 8 per thread -- 96 for 18 warps, 80 for 21 and 24
template <int S, int MAXW>
__global__ void __maxnreg__((65536 / (32 * ((MAXW + 3) & ~3))) & ~7)
    pyramid_u8_fused_kernel(const __grid_constant__ PfParams p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem[];
  // the warp index through a shuffle: the compiler then knows that everything derived from it (strip, slot, ring and
  // barrier addresses, frame numbers, TMA coordinates) is warp-uniform and keeps it in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int slot = warp / p.n_strips, strip = warp - slot * p.n_strips;
  const bool left = strip == 0, right = strip == p.n_strips - 1;
  const int base = p.strip_base[strip];
  const int col = base + lane;
  const int first_lane = left ? 0 : -1;                         // lanes holding the image's first / last level-3 column
  const int last_lane = right ? p.W3 - 1 - base : -1;
  PfLane ln;
  ln.s_m2 = (reflect101(col - 2, p.W3) - base) & 31; ln.s_m1 = (reflect101(col - 1, p.W3) - base) & 31;
  ln.s_p1 = (reflect101(col + 1, p.W3) - base) & 31; ln.s_p2 = (reflect101(col + 2, p.W3) - base) & 31;
  ln.k = (!(col & 1) && (col >> 1) >= p.strip_k0[strip] && (col >> 1) < p.strip_k1[strip]) ? (col >> 1) : -1;
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(smem) + warp * S * PU_STAGE_BYTES;
  const unsigned char* my = smem + (size_t)warp * S * PU_STAGE_BYTES + lane * 8;
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(smem + p.mbar_base) + warp * S * 8;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < S; ++i) pu_mbar_init(mbar + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  unsigned char* lvl_slot = smem + p.lvl_base + (size_t)slot * p.lvl_stride;
  double* g4 = pf_level(p, lvl_slot, p.first);
  const int nt = p.n_strips * 32, t = strip * 32 + lane;
  const long long stride = (long long)gridDim.x * p.frames_per_cta;
  long long frame = (long long)blockIdx.x * p.frames_per_cta + slot;
  unsigned phase = 0;
  int stage = 0;
  if (frame < p.n_frames) {   // the first S-1 blocks of the slot's first frame
    const int z = (int)pu_source_frame(frame, p.seg_len, p.seg_stride, p.seg_first);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < S - 1; ++i) {
        const int b = i - 2;
        pu_mbar_expect_tx(mbar + 8 * i, PU_STAGE_BYTES);
        pu_tma_load(ring_s + i * PU_STAGE_BYTES, &tmap, 8 * base, b >= 0 ? 8 * b : -8 * b - 7, z, mbar + 8 * i);
      }
    }
  }
  for (; frame < p.n_frames; frame += stride) {
    const int sframe = (int)pu_source_frame(frame, p.seg_len, p.seg_stride, p.seg_first);
    const int next_sframe = frame + stride < p.n_frames ? (int)pu_source_frame(frame + stride, p.seg_len, p.seg_stride, p.seg_first) : -1;
    if (left || right || p.one_variant)
      pf_run_frame<S, true>(p, &tmap, sframe, next_sframe, ring_s, my, mbar, phase, stage, g4, ln, lane, 8 * base, first_lane,
                            last_lane);
    else
      pf_run_frame<S, false>(p, &tmap, sframe, next_sframe, ring_s, my, mbar, phase, stage, g4, ln, lane, 8 * base, first_lane,
                             last_lane);
    // The barrier that completes the level-4 image also re-aligns the strips of the frame (they share halo columns:
    // left alone they drift apart until a halo sector has left L2 before the neighbour asks for it, profiles/r01g).
    pf_slot_sync(slot, nt);
    pf_tail(p, lvl_slot, p.lap_out + frame * p.record_len, slot, t, nt);
    pf_slot_sync(slot, nt);                                    // the level images are free for the next frame
  }
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
  int lvl_bytes = 0;
  for (int l = 0; l < L; ++l) {
    p.w[l] = g.w[l]; p.h[l] = g.h[l]; p.rec_off[l] = rec.off[l];
    p.magic[l] = g.w[l] > 1 ? (unsigned)((0x100000000ull + g.w[l] - 1) / g.w[l]) : 0u;
    if (l >= s) { p.lvl_off[l] = lvl_bytes; lvl_bytes += g.w[l] * g.h[l] * 8; }
    if (l >= s && (long long)g.w[l] * g.h[l] >= 65536) return false;      // pf_div
  }
  lvl_bytes = (lvl_bytes + 15) & ~15;
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
  // Few strips per frame are mostly edge strips: let the interior ones run the edge code too (one hot loop in the
  // instruction cache instead of two): 0.734 -> 0.708 ms per 8192 VGA frames; wide frames keep the lean interior loop.
  p.one_variant = h->pyramid_variants == 1 || (h->pyramid_variants == 0 && n <= 3);
  // ring depth / warps per CTA: as many frame slots as shared memory, the register file and 15 named barriers allow
  const int cfg[3][2] = {{4, 18}, {3, 21}, {2, 24}};            // (stages, max warps)
  // measured (r02l): frames of up to three strips (VGA: 3 x 8 slots) want the most warps, 2 stages x 24 warps -- 0.667 ms
  // per 8192 VGA frames against 0.749 with 4 x 18; wider frames, whose level images leave room for few slots anyway,
  // want the deep ring and the 96 registers: 720p 0.527 against 0.605 ms per 2048 frames, 1080p 0.783 against 0.880
  int pick = h->pyramid_cfg >= 1 && h->pyramid_cfg <= 3 ? h->pyramid_cfg - 1 : (n <= 3 ? 2 : 0);
  for (int tries = 0; tries < 3; ++tries, pick = (pick + 1) % 3) {
    const int S = cfg[pick][0], maxw = cfg[pick][1];
    const int per_slot = n * S * PU_STAGE_BYTES + lvl_bytes + n * S * 8;
    int fpc = maxw / n;
    const int fit = (h->smem_optin - 256) / per_slot;
    if (fpc > fit) fpc = fit;
    if (fpc > 15) fpc = 15;
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
  if (pl.stages == 3) return pf_launch_cfg<3, 21>(h, pl, map, ctas, st);
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
