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
// HBM traffic per frame: W*H bytes read once (+ the halo columns, normally L2 hits) and W*H/16 bytes written.
// The remaining levels (3 -> 4 -> ... and the Laplacians) are the tail kernel's (pyramid.cu).
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

template <int WT, bool LEFT, bool RIGHT>
__device__ __forceinline__ void pu_run_frame(const PuParams& p, const uint8_t* __restrict__ fsrc, uint32_t* __restrict__ g3,
                                             unsigned char* ring, int lane, int col, int store_lo, int store_hi,
                                             int last_lane) {
  const int W = WT ? WT : p.W;        // a compile-time width turns the row offsets of the copies into immediates
  const bool in_img = col >= 0 && col < p.W3 && lane <= store_hi + PU_HALO_LANES;
  const uint8_t* lsrc = fsrc + (in_img ? 8 * col : 0);
  const int src_bytes = in_img ? 8 : 0;
  const int nblk = p.H >> 3;
  unsigned char* my = ring + lane * 8;
  auto issue = [&](int b) {
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
#pragma unroll 2
  for (int b = -2; b < nblk; ++b) {
    issue(b + 3);
    pu_wait<3>();                                              // block b has landed (3 younger groups may be in flight)
    const unsigned char* src = my + ((b + 2) & (PU_STAGES - 1)) * PU_STAGE_BYTES;
    uint2 w[PU_ROWS];
#pragma unroll
    for (int i = 0; i < PU_ROWS; ++i) w[i] = *reinterpret_cast<const uint2*>(src + i * 256);
    unsigned o;
    pu_block<LEFT, RIGHT>(s, w, lane, last_lane, o);
    if (b >= 1 && storing) out[(long long)(b - 1) * p.W3] = o;
  }
  const unsigned o = pu_flush<LEFT, RIGHT>(s, lane, last_lane);
  if (storing) out[(long long)(nblk - 1) * p.W3] = o;
  pu_wait<0>();
}

template <int WT>
__global__ void __launch_bounds__(PU_BOUND_WARPS * 32, PU_CTAS_PER_SM) pyramid_front_u8_kernel(const PuParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
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
    uint32_t* g3 = p.g3 + frame * g3_elems;
    // The strips of a frame share their halo columns.  Left alone, the warps of a slot drift apart over the frames (edge
    // strips are cheaper) until a halo sector read by one warp has left L2 before its neighbour asks for it -- 23 % extra
    // DRAM reads at 8192 frames per launch (ncu, profiles/r01g).  A named barrier per slot re-aligns them every frame.
    if (p.n_strips > 1) asm volatile("bar.sync %0, %1;\n" ::"r"(slot + 1), "r"(p.n_strips * 32) : "memory");
    if (left && right) pu_run_frame<WT, true, true>(p, fsrc, g3, ring, lane, col, store_lo, store_hi, store_hi);
    else if (left) pu_run_frame<WT, true, false>(p, fsrc, g3, ring, lane, col, store_lo, store_hi, store_hi);
    else if (right) pu_run_frame<WT, false, true>(p, fsrc, g3, ring, lane, col, store_lo, store_hi, store_hi);
    else pu_run_frame<WT, false, false>(p, fsrc, g3, ring, lane, col, store_lo, store_hi, store_hi);
  }
}

// The integer path needs: uint8 frames, four front levels, W and H multiples of 8 (all three decimated sizes even),
// at least 17 rows (the mirrored lead-in) and 8-byte aligned rows.
bool pu_supported(const void* frames, int W, int H, int skip) {
  return skip == 4 && W % 8 == 0 && H % 8 == 0 && H >= 24 && W >= 16 && ((uintptr_t)frames % 8) == 0;
}

int32_t pu_launch(rm_handle* h, const uint8_t* frames, uint32_t* g3, long long n_frames, long long seg_len,
                  long long seg_stride, long long seg_first, int W, int H, cudaStream_t st) {
  PuParams p;
  memset(&p, 0, sizeof(p));
  p.frames = frames; p.g3 = g3; p.n_frames = n_frames; p.frame_elems = (long long)W * H;
  p.seg_len = seg_len; p.seg_stride = seg_stride; p.seg_first = seg_first;
  p.W = W; p.H = H; p.W3 = W / 8; p.H3 = H / 8;
  const int cap = 32 - 2 * PU_HALO_LANES;                       // payload lanes of an interior strip
  p.n_strips = (p.W3 + cap - 1) / cap;
  p.cols_per_strip = (p.W3 + p.n_strips - 1) / p.n_strips;
  if (p.n_strips > 16) return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: frame wider than 3584 pixels", __func__);
#ifdef PU_IDX32
  if (n_frames >= (1ll << 31) || seg_len >= (1ll << 31) || seg_len < 1)
    return rm_fail(h, RM_ERR_UNSUPPORTED, "%s: more than 2^31 frames", __func__);
#endif
  p.frames_per_cta = PU_MAX_WARPS / p.n_strips;
  const int warps = p.frames_per_cta * p.n_strips;
  const int smem = warps * PU_STAGES * PU_STAGE_BYTES;
  void (*kern)(const PuParams) = W == 640 ? pyramid_front_u8_kernel<640>
                                 : W == 1280 ? pyramid_front_u8_kernel<1280>
                                 : W == 1920 ? pyramid_front_u8_kernel<1920>
                                 : W == 320 ? pyramid_front_u8_kernel<320> : pyramid_front_u8_kernel<0>;
  RM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long ctas = (n_frames + p.frames_per_cta - 1) / p.frames_per_cta;
  if (ctas > (long long)h->sm_count * PU_CTAS_PER_SM) ctas = (long long)h->sm_count * PU_CTAS_PER_SM;
  RM_PROF(h, st, "pyramid_front_u8_kernel");
  kern<<<(unsigned)ctas, warps * 32, smem, st>>>(p);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}
