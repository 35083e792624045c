// uint8 pyramid stage (pyramid_u8.cu): launch interface used by pyramid.cu.
#pragma once
#include "common.cuh"

#define PU_HALO_LANES 2

struct PuParams {          // fallback front: level 3 to HBM
  const uint8_t* frames;
  uint32_t* g3;            // (n_frames, H/8, W/8) Gaussian level 3 as exact integers: value * 2^24 * 255
  long long n_frames;
  long long frame_elems;   // W*H
  // frame f of the batch is source frame (f / seg_len) * seg_stride + seg_first + f % seg_len
  long long seg_len, seg_stride, seg_first;
  int W, H, W3, H3;
  int n_strips;            // vertical strips per frame, one warp each
  int cols_per_strip;      // level-3 columns stored by a strip
  int frames_per_cta;
};

struct PfParams {          // fused kernel: frames -> packed Laplacian records
  double* lap_out;         // (n_frames, record_len)
  double g_scale;          // level `first` = exact integer * g_scale  (2^-32 / 255)
  long long n_frames, seg_len, seg_stride, seg_first;
  int W, H, W3, H3;
  int n_strips, frames_per_cta;
  int strip_base[16];      // level-3 column held by lane 0 of the strip's warp (even)
  int strip_k0[16], strip_k1[16];   // level-`first` columns [k0, k1) the strip emits
  int first, top;          // Gaussian levels first..top are built; Laplacian levels first..top-1 are written
  int g4_global;           // level `first` is written into the record (and turned into its Laplacian there), not into shared memory
  int w[RM_MAX_LEVELS], h[RM_MAX_LEVELS], rec_off[RM_MAX_LEVELS];
  int lvl_off[RM_MAX_LEVELS];       // byte offset of Gaussian level l inside a slot's level images
  unsigned magic[RM_MAX_LEVELS];    // ceil(2^32 / w[l])
  int record_len;
  int lvl_base, lvl_stride;         // byte offset of slot 0's level images in shared memory, bytes per slot
  int mbar_base;                    // byte offset of the mbarriers: stages per warp
};

bool pu_supported(const void* frames, int W, int H, int skip);
// 1: pu_launch_fused can take these frames; 0: pu_launch_front + pyramid_tail_kernel
int pu_best_mode(rm_handle* h, const void* frames, int W, int H);
int32_t pu_launch_fused(rm_handle* h, const uint8_t* frames, double* lap_out, long long n_frames, long long seg_len,
                        long long seg_stride, long long seg_first, int W, int H, cudaStream_t st);
int32_t pu_launch_front(rm_handle* h, const uint8_t* frames, int bgr, uint32_t* g3, long long n_frames, long long seg_len,
                        long long seg_stride, long long seg_first, int W, int H, cudaStream_t st);
