// ROI selection on the device: the tail of locate() (base.py:566-575).
//
//   cv2.threshold(avg, threshold, 255, THRESH_BINARY)      -> fg = heat > threshold
//   cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)   -> 8-connected components by union-find (the root of a
//                                                             component is its first pixel in raster order, which is
//                                                             where Suzuki-Abe starts its outer border), then one
//                                                             thread per root follows the outer border (roi_core.h)
//   max(contours, key=cv2.contourArea)                     -> atomicMax of (2*area, start index) keys per clip
//   cv2.boundingRect                                       -> every traced border leaves its box in a per-clip list;
//                                                             the winner's is looked up
//
// A component nested in a hole of another one is not an external contour, but its polygon lies strictly inside the
// enclosing one, so it can never be the maximum: treating every component as a candidate gives the same answer.
// The working set is one uint8 heat map and one int32 label plane per clip; the kernels are latency bound and run
// once per calibration, concurrently over all clips of the batch.
#include "common.cuh"
#include "roi_core.h"

struct RoiParams {
  const uint8_t* heat;      // (n_clips, H, W)
  int32_t* labels;          // (n_clips, H, W)
  unsigned long long* best; // (n_clips)
  int n_clips, W, H, threshold;
  int32_t* roi;             // (n_clips, 4)
  int32_t* status;          // (n_clips)
  // every traced component leaves (key, packed bounding box) in its clip's list so that the winner's box is looked up
  // instead of being traced a second time; a clip with more than ROI_LIST_CAP components falls back to the second trace
  unsigned long long* list; // (n_clips, ROI_LIST_CAP, 2)
  int* list_n;              // (n_clips)
};
#define ROI_LIST_CAP 2048

__device__ __forceinline__ int uf_find(const int32_t* L, int i) {
  const volatile int32_t* V = L;   // parents only ever decrease; a stale read is still an ancestor
  int p = V[i];
  while (p != i) {
    i = p;
    p = V[i];
  }
  return i;
}
__device__ __forceinline__ void uf_union(int32_t* L, int a, int b) {
  for (;;) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a == b) return;
    if (a > b) { int t = a; a = b; b = t; }
    const int old = atomicMin(&L[b], a);   // hang the larger root under the smaller one
    if (old == b) return;
    b = old;
  }
}

__global__ void roi_init_kernel(const RoiParams p) {
  const long long hw = (long long)p.W * p.H;
  const uint8_t* heat = p.heat + blockIdx.y * hw;
  int32_t* L = p.labels + blockIdx.y * hw;
  if ((hw & 3) == 0 && (((uintptr_t)p.heat | (uintptr_t)p.labels) & 15) == 0) {
    // four pixels per thread: one 4-byte load, one 16-byte store
    const uchar4* h4 = reinterpret_cast<const uchar4*>(heat);
    int4* L4 = reinterpret_cast<int4*>(L);
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < hw / 4; q += (long long)gridDim.x * blockDim.x) {
      const uchar4 v = h4[q];
      const int i = (int)(4 * q);
      L4[q] = make_int4(v.x > p.threshold ? i : -1, v.y > p.threshold ? i + 1 : -1, v.z > p.threshold ? i + 2 : -1,
                        v.w > p.threshold ? i + 3 : -1);
    }
  } else {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x)
      L[i] = heat[i] > p.threshold ? (int)i : -1;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { p.best[blockIdx.y] = 0ull; p.list_n[blockIdx.y] = 0; }
}

__global__ void roi_merge_kernel(const RoiParams p) {
  const long long hw = (long long)p.W * p.H;
  int32_t* L = p.labels + blockIdx.y * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    if (L[i] < 0) continue;
    const int x = (int)(i % p.W), y = (int)(i / p.W);
    // 8-connectivity with the redundant unions left out: a foreground pixel above is adjacent to the left, upper-left and
    // upper-right neighbours, which join it through their own unions; likewise the left neighbour covers the upper-left
    const long long up = i - p.W;
    const bool f_left = x > 0 && L[i - 1] >= 0;
    const bool f_up = y > 0 && L[up] >= 0;
    if (f_up) {
      uf_union(L, (int)i, (int)up);
    } else {
      if (f_left) uf_union(L, (int)i, (int)i - 1);
      else if (y > 0 && x > 0 && L[up - 1] >= 0) uf_union(L, (int)i, (int)up - 1);
      if (y > 0 && x + 1 < p.W && L[up + 1] >= 0) uf_union(L, (int)i, (int)up + 1);
    }
  }
}

struct HeatFg {
  const uint8_t* heat;
  int W, H, thr;
  __device__ __forceinline__ bool operator()(int x, int y) const {
    return x >= 0 && y >= 0 && x < W && y < H && heat[(long long)y * W + x] > thr;
  }
};

__global__ void roi_trace_kernel(const RoiParams p) {
  const long long hw = (long long)p.W * p.H;
  const int32_t* L = p.labels + blockIdx.y * hw;
  const HeatFg fg{p.heat + blockIdx.y * hw, p.W, p.H, p.threshold};
  const int max_steps = (int)(4 * hw + 8 < 0x7fffffff ? 4 * hw + 8 : 0x7fffffff);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    if (L[i] != (int)i) continue;   // not the first pixel of a component
    const RoiTrace t = roi_trace_outer((int)(i % p.W), (int)(i / p.W), fg, max_steps);
    const unsigned long long key = roi_key(t.area2, (int)i);
    atomicMax(&p.best[blockIdx.y], key);
    const int slot = atomicAdd(&p.list_n[blockIdx.y], 1);
    if (slot < ROI_LIST_CAP && p.W < 65536 && p.H < 65536) {
      unsigned long long* e = p.list + ((long long)blockIdx.y * ROI_LIST_CAP + slot) * 2;
      e[0] = key;
      e[1] = (unsigned long long)t.x0 | ((unsigned long long)t.y0 << 16) | ((unsigned long long)t.x1 << 32) |
             ((unsigned long long)t.y1 << 48);
    }
  }
}

__global__ void roi_finish_kernel(const RoiParams p) {
  const int clip = blockIdx.x * blockDim.x + threadIdx.x;
  if (clip >= p.n_clips) return;
  const long long hw = (long long)p.W * p.H;
  const unsigned long long k = p.best[clip];
  int32_t* out = p.roi + clip * 4;
  if (!k) {   // no contour: locate() returns None (base.py:569-570)
    out[0] = out[1] = out[2] = out[3] = 0;
    if (p.status) p.status[clip] = RM_CLIP_NO_ROI;
    return;
  }
  if (p.list_n[clip] <= ROI_LIST_CAP && p.W < 65536 && p.H < 65536) {   // the winner's box is in the list
    const unsigned long long* e = p.list + (long long)clip * ROI_LIST_CAP * 2;
    for (int i = 0; i < p.list_n[clip]; ++i)
      if (e[2 * i] == k) {
        const unsigned long long b = e[2 * i + 1];
        const int x0 = (int)(b & 0xffff), y0 = (int)((b >> 16) & 0xffff), x1 = (int)((b >> 32) & 0xffff), y1 = (int)(b >> 48);
        out[0] = x0; out[1] = y0; out[2] = x1 - x0 + 1; out[3] = y1 - y0 + 1;
        if (p.status) p.status[clip] = RM_CLIP_OK;
        return;
      }
  }
  const int start = (int)(k & 0xffffffffu) - 1;
  const HeatFg fg{p.heat + clip * hw, p.W, p.H, p.threshold};
  const int max_steps = (int)(4 * hw + 8 < 0x7fffffff ? 4 * hw + 8 : 0x7fffffff);
  const RoiTrace t = roi_trace_outer(start % p.W, start / p.W, fg, max_steps);
  out[0] = t.x0; out[1] = t.y0; out[2] = t.x1 - t.x0 + 1; out[3] = t.y1 - t.y0 + 1;
  if (p.status) p.status[clip] = RM_CLIP_OK;
}

extern "C" int32_t rm_roi_workspace_bytes(rm_handle* h, int32_t W, int32_t H, int32_t n_clips, size_t* out) {
  RM_CHECK_ARG(h, h && out && W >= 1 && H >= 1 && n_clips >= 0, "null pointer or bad size");
  *out = (size_t)n_clips * W * H * 4 + (size_t)n_clips * 8 + (size_t)n_clips * ROI_LIST_CAP * 16 + (size_t)n_clips * 4 +
         4 * 256;
  return RM_OK;
}

extern "C" int32_t rm_roi_select(rm_handle* h, const uint8_t* heat, int32_t n_clips, int32_t W, int32_t H,
                                 int32_t threshold, int32_t* roi_out, int32_t* status_out, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  RM_CHECK_ARG(h, h && heat && roi_out && n_clips >= 0 && W >= 1 && H >= 1, "null pointer or bad size");
  RM_CHECK_ARG(h, (long long)W * H < (1ll << 30), "image too large");
  if (n_clips == 0) return RM_OK;
  size_t need = 0;
  rm_roi_workspace_bytes(h, W, H, n_clips, &need);
  if (!workspace || workspace_bytes < need)
    return rm_fail(h, RM_ERR_WORKSPACE, "%s: workspace too small (%lld needed, %lld given)", __func__, (long long)need,
                   (long long)workspace_bytes);
  DeviceGuard dg(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  RoiParams p;
  uintptr_t base = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
  p.labels = reinterpret_cast<int32_t*>(base);
  base += ((size_t)n_clips * W * H * 4 + 255) & ~(size_t)255;
  p.best = reinterpret_cast<unsigned long long*>(base);
  base += ((size_t)n_clips * 8 + 255) & ~(size_t)255;
  p.list = reinterpret_cast<unsigned long long*>(base);
  base += ((size_t)n_clips * ROI_LIST_CAP * 16 + 255) & ~(size_t)255;
  p.list_n = reinterpret_cast<int*>(base);
  p.heat = heat; p.n_clips = n_clips; p.W = W; p.H = H; p.threshold = threshold;
  p.roi = roi_out; p.status = status_out;
  const long long hw = (long long)W * H;
  dim3 grid((unsigned)(div_up(hw, 256) < 2 * h->sm_count ? div_up(hw, 256) : 2 * h->sm_count), n_clips);
  RM_PROF(h, st, "roi_init_kernel");
  roi_init_kernel<<<grid, 256, 0, st>>>(p);
  RM_LAUNCH_CHECK(h);
  RM_PROF(h, st, "roi_merge_kernel");
  roi_merge_kernel<<<grid, 256, 0, st>>>(p);
  RM_LAUNCH_CHECK(h);
  RM_PROF(h, st, "roi_trace_kernel");
  roi_trace_kernel<<<grid, 256, 0, st>>>(p);
  RM_LAUNCH_CHECK(h);
  RM_PROF(h, st, "roi_finish_kernel");
  roi_finish_kernel<<<div_up(n_clips, 64), 64, 0, st>>>(p);
  RM_LAUNCH_CHECK(h);
  return RM_OK;
}
