// TEST INFRASTRUCTURE: compiles respmon_b200/csrc/roi_core.h (the header the CUDA ROI kernel uses) for the host so the
// border following / selection rule can be compared with cv2 on a machine without a GPU.  Never loaded by the product.
#include <vector>

#include "../../respmon_b200/csrc/roi_core.h"

extern "C" int host_select_roi(const unsigned char* bin, int W, int H, int* xywh, long long* area2_out) {
  auto fg = [&](int x, int y) { return x >= 0 && y >= 0 && x < W && y < H && bin[y * W + x] != 0; };
  std::vector<char> seen((size_t)W * H, 0);
  std::vector<int> stack;
  unsigned long long best = 0;
  RoiTrace best_t{};
  for (int i = 0; i < W * H; ++i) {
    if (!bin[i] || seen[i]) continue;
    // i is the first pixel of a new 8-connected component in raster order; flood-fill marks the rest
    stack.push_back(i);
    seen[i] = 1;
    while (!stack.empty()) {
      int p = stack.back();
      stack.pop_back();
      int px = p % W, py = p / W;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          int qx = px + dx, qy = py + dy;
          if (fg(qx, qy) && !seen[qy * W + qx]) {
            seen[qy * W + qx] = 1;
            stack.push_back(qy * W + qx);
          }
        }
    }
    RoiTrace t = roi_trace_outer(i % W, i / W, fg, 4 * W * H + 8);
    unsigned long long k = roi_key(t.area2, i);
    if (k > best) { best = k; best_t = t; }
  }
  if (!best) return 0;
  xywh[0] = best_t.x0; xywh[1] = best_t.y0; xywh[2] = best_t.x1 - best_t.x0 + 1; xywh[3] = best_t.y1 - best_t.y0 + 1;
  *area2_out = best_t.area2;
  return 1;
}
