// Integer uint8 front of the pyramid (pyramid_u8.cu): launch interface used by pyramid.cu.
#pragma once
#include "common.cuh"

#define PU_HALO_LANES 2

struct PuParams {
  const uint8_t* frames;
  uint32_t* g3;            // (n_frames, H/8, W/8) Gaussian level 3 as exact integers: value * 2^24 * 255 (split path only)
  long long n_frames;
  long long frame_elems;   // W*H
  // frame f of the batch is source frame (f / seg_len) * seg_stride + seg_first + f % seg_len
  long long seg_len, seg_stride, seg_first;
  int W, H, W3, H3;
  int n_strips;            // vertical strips per frame, one warp each
  int cols_per_strip;      // level-3 columns stored by a strip
  int frames_per_cta;
  // fused tail: level 3 stays in shared memory, the warps of a frame slot finish the pyramid (pyramid.py:13-15 for
  // levels first..top, pyramid.py:24-26 for the Laplacians first..top-1) and write the packed record themselves
  double* lap_out;         // (n_frames, record_len)
  double g_scale;          // level `first` = exact integer 5x5 sum over level 3 * g_scale  (2^-32 / 255)
  int first, top;
  int w[RM_MAX_LEVELS], h[RM_MAX_LEVELS], rec_off[RM_MAX_LEVELS];
  int lvl_off[RM_MAX_LEVELS];   // byte offset of Gaussian level l: >= 0 inside the slot's ring, < 0: -(off+1) inside its extra region
  int record_len;
  int ring_bytes;          // all rings (warps * PU_STAGES * PU_STAGE_BYTES)
  int l3_base, l3_stride;  // byte offset of slot 0's level-3 image in shared memory, bytes per slot
  int extra_base, extra_stride;
  int mbar_base;           // byte offset of the mbarriers (TMA path): PU_STAGES per warp
};

// mode: 0 = level 3 to HBM, pyramid_tail_kernel finishes (any size the integer front supports);
//       1 = fused tail, rows staged by cp.async;  2 = fused tail, rows staged by TMA (cp.async.bulk.tensor + mbarrier)
bool pu_supported(const void* frames, int W, int H, int skip);
int pu_best_mode(rm_handle* h, const void* frames, int W, int H, int n_levels, int skip);
int32_t pu_launch(rm_handle* h, int mode, const uint8_t* frames, uint32_t* g3, double* lap_out, long long n_frames,
                  long long seg_len, long long seg_stride, long long seg_first, int W, int H, cudaStream_t st);
