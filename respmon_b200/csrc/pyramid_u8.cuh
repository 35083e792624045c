// Integer uint8 front of the pyramid (pyramid_u8.cu): launch interface used by pyramid.cu.
#pragma once
#include "common.cuh"

#define PU_HALO_LANES 2

struct PuParams {
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

bool pu_supported(const void* frames, int W, int H, int skip);
int32_t pu_launch(rm_handle* h, const uint8_t* frames, uint32_t* g3, long long n_frames, long long seg_len,
                  long long seg_stride, long long seg_first, int W, int H, cudaStream_t st);
