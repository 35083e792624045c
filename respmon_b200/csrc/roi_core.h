// Outer-border following of one 8-connected component of a binary image, written once for device and host.
//
// Replaces cv2.findContours(RETR_EXTERNAL) + cv2.contourArea + cv2.boundingRect as locate() uses them
// (base.py:568-575): Suzuki-Abe border following started at the component's first pixel in raster order (whose left
// neighbour is background by construction), the polygon through the visited pixel centres, its shoelace area and
// its bounding box.  CHAIN_APPROX_SIMPLE only drops collinear points, which leaves both results unchanged.
// The same header is compiled by g++ into a test-only library (tests/hostsim) and checked against cv2 there.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define ROI_HD __host__ __device__ __forceinline__
#else
#define ROI_HD inline
#endif

struct RoiTrace {
  long long area2;        // |sum(prev.x*cur.y - prev.y*cur.x)| = 2 * cv2.contourArea
  int x0, y0, x1, y1;     // inclusive bounding box of the border (== that of the component)
  int steps;
};

// `fg(x, y)` must return 0 outside the image.  Direction codes follow OpenCV: 0 E, 1 NE, 2 N, 3 NW, 4 W, 5 SW, 6 S, 7 SE.
template <typename Fg>
ROI_HD RoiTrace roi_trace_outer(int sx, int sy, Fg fg, int max_steps) {
  const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1};
  const int dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
  RoiTrace r;
  r.area2 = 0;
  r.x0 = r.x1 = sx;
  r.y0 = r.y1 = sy;
  r.steps = 0;
  // The 8 neighbours of the current pixel are probed together (independent loads, one round trip) into a bit mask, bit k
  // = direction k; the direction searches below then run on the mask.
  auto mask8 = [&](int x, int y) {
    unsigned m = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 8; ++k) m |= (fg(x + dx[k], y + dy[k]) ? 1u : 0u) << k;
    return m;
  };
  // first non-zero neighbour clockwise from W
  int s = 4;
  int x1 = sx, y1 = sy;
  bool found = false;
  unsigned m = mask8(sx, sy);
  for (int k = 0; k < 7; ++k) {
    s = (s - 1) & 7;
    if ((m >> s) & 1u) { x1 = sx + dx[s]; y1 = sy + dy[s]; found = true; break; }
  }
  if (!found) return r;   // isolated pixel: area 0, bbox 1x1
  int x3 = sx, y3 = sy;
  long long acc = 0;
  int px = 0, py = 0, fx = 0, fy = 0;   // previous / first emitted points
  bool have = false;
  for (;;) {
    int x4, y4;
    for (;;) {   // next border pixel counter-clockwise
      s = (s + 1) & 7;
      if ((m >> s) & 1u) break;
    }
    x4 = x3 + dx[s];
    y4 = y3 + dy[s];
    if (have) acc += (long long)px * y3 - (long long)py * x3;
    else { fx = x3; fy = y3; have = true; }
    px = x3; py = y3;
    r.x0 = x3 < r.x0 ? x3 : r.x0; r.x1 = x3 > r.x1 ? x3 : r.x1;
    r.y0 = y3 < r.y0 ? y3 : r.y0; r.y1 = y3 > r.y1 ? y3 : r.y1;
    ++r.steps;
    if ((x4 == sx && y4 == sy && x3 == x1 && y3 == y1) || r.steps >= max_steps) break;
    x3 = x4; y3 = y4;
    s = (s + 4) & 7;
    m = mask8(x3, y3);
  }
  acc += (long long)px * fy - (long long)py * fx;   // close the polygon
  r.area2 = acc < 0 ? -acc : acc;
  return r;
}

// Selection key: largest area first; among equal areas the component whose first pixel comes LAST in raster order
// (cv2 lists contours in reverse discovery order and Python's max() keeps the first maximum, base.py:571).
ROI_HD unsigned long long roi_key(long long area2, int start_index) {
  return ((unsigned long long)area2 << 32) | (unsigned)(start_index + 1);
}
