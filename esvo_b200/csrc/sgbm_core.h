// esvo_b200 product code -- semi-global block matching of the initialisation (integer, bit-exact).
//
// Replaces the cv::StereoSGBM::compute call of esvo_Mapping::InitializationAtTime
// (esvo_core/src/esvo_Mapping.cpp:101-108 create(0, 48, 11, 8*11*11, 32*11*11, -1, 0, 11); :444 compute) --
// OpenCV's MODE_SGBM: Birchfield-Tomasi cost on the clipped x-derivative and on the intensities, 11x11 box
// aggregation with replicated borders, five 16-bit path costs per pixel (left, top-left, top, top-right, right),
// first-minimum disparity, uniqueness test, sub-pixel parabola, left-right check, 3x3 median.
//
// Every stage is written as a per-work-item function (HD = host + device): sgbm.cu runs them as one CUDA thread per
// image row / pixel / aggregation path; tests/sgbm_host_check.cpp runs the very same functions in plain loops, so the
// arithmetic is verified against cv2 (through oracle/sgbm.py) even where no GPU is available.  Integer work: the bar
// is bit-exact.  One thread per path is slow by GPU standards (a few hundred threads) but this is a one-off at start-up
// on a 346x260 image; the hot path does not contain it.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define SGBM_HD __host__ __device__ __forceinline__
#else
#define SGBM_HD inline
#endif

namespace esvo_sgbm {

constexpr int kMaxD = 128;        // numDisparities <= 128
constexpr int kMaxCost = 32767;   // CostType = short
constexpr int kDispShift = 4, kDispScale = 16;

struct Dims {
  int W, H;            // image
  int D;               // numDisparities (minDisparity is 0)
  int W1, minX1;       // matched columns: x in [minX1, W), W1 = W - minX1, minX1 = D
  int SW2, SH2;        // half window
  int P1, P2, uniq, d12, ftzero;
};

SGBM_HD int imin(int a, int b) { return a < b ? a : b; }
SGBM_HD int imax(int a, int b) { return a > b ? a : b; }
SGBM_HD int clip_tab(int k, int ftzero) { return imin(imax(k, -ftzero), ftzero) + ftzero; }

// ---- stage 1 (per image row y, per image): the two "channels" BT compares -- clipped x-derivative and intensity --
// with OpenCV's border columns (= tab[0]), plus the min / max of each value with its two half-pixel interpolations.
// out: 6 rows of W int16: a0, lo0, hi0 (derivative), a1, lo1, hi1 (intensity).
SGBM_HD void prep_row(const uint8_t* img, int pitch, const Dims& dm, int y, int16_t* out) {
  const int W = dm.W, H = dm.H;
  const uint8_t* r = img + (size_t)y * pitch;
  const uint8_t* n = y > 0 ? r - pitch : r;
  const uint8_t* s = y < H - 1 ? r + pitch : r;
  int16_t* a0 = out; int16_t* a1 = out + 3 * W;
  const int t0 = clip_tab(0, dm.ftzero);
  for (int x = 0; x < W; ++x) {
    if (x == 0 || x == W - 1) { a0[x] = (int16_t)t0; a1[x] = (int16_t)t0; continue; }
    const int g = ((int)r[x + 1] - (int)r[x - 1]) * 2 + (int)n[x + 1] - (int)n[x - 1] + (int)s[x + 1] - (int)s[x - 1];
    a0[x] = (int16_t)clip_tab(g, dm.ftzero);
    a1[x] = (int16_t)r[x];
  }
  for (int c = 0; c < 2; ++c) {
    const int16_t* a = out + 3 * W * c;
    int16_t* lo = out + 3 * W * c + W; int16_t* hi = out + 3 * W * c + 2 * W;
    for (int x = 0; x < W; ++x) {
      const int v = a[x];
      const int vl = x > 0 ? (v + a[x - 1]) / 2 : v;
      const int vr = x < W - 1 ? (v + a[x + 1]) / 2 : v;
      lo[x] = (int16_t)imin(imin(vl, vr), v);
      hi[x] = (int16_t)imax(imax(vl, vr), v);
    }
  }
}

// ---- stage 2 (per row y, matched column x1): Birchfield-Tomasi pixel cost for every disparity (calcPixelCostBT)
SGBM_HD void pixel_cost(const int16_t* pl, const int16_t* pr, const Dims& dm, int x1, int16_t* cost /*D*/) {
  const int W = dm.W, x = x1 + dm.minX1;
  for (int d = 0; d < dm.D; ++d) {
    int sum = 0;
    for (int c = 0; c < 2; ++c) {
      const int16_t* a = pl + 3 * W * c; const int16_t* b = pr + 3 * W * c;
      const int u = a[x], u0 = a[W + x], u1 = a[2 * W + x];
      const int xr = x - d;
      const int v = b[xr], v0 = b[W + xr], v1 = b[2 * W + xr];
      const int c0 = imax(imax(0, u - v1), v0 - u);
      const int c1 = imax(imax(0, v - u1), u0 - v);
      sum += imin(c0, c1) >> (c == 0 ? 0 : 2);
    }
    cost[d] = (int16_t)sum;
  }
}

// ---- stage 3 / 4 (per pixel): horizontal, then vertical box sum with replicated borders (hsumAdd / C of the reference)
SGBM_HD void box_h(const int16_t* pix_row /*W1*D*/, const Dims& dm, int x1, int16_t* out /*D*/) {
  for (int d = 0; d < dm.D; ++d) {
    int s = 0;
    for (int k = -dm.SW2; k <= dm.SW2; ++k) s += pix_row[(size_t)imin(imax(x1 + k, 0), dm.W1 - 1) * dm.D + d];
    out[d] = (int16_t)s;
  }
}
SGBM_HD void box_v(const int16_t* hsum /*H*W1*D*/, const Dims& dm, int y, int x1, int16_t* out /*D*/) {
  for (int d = 0; d < dm.D; ++d) {
    int s = 0;
    for (int k = -dm.SH2; k <= dm.SH2; ++k) s += hsum[((size_t)imin(imax(y + k, 0), dm.H - 1) * dm.W1 + x1) * dm.D + d];
    out[d] = (int16_t)s;
  }
}

// ---- stage 5 (per path): L_r(p, d) = C(p, d) + min(L_r(p-r, d), L_r(p-r, d-1) + P1, L_r(p-r, d+1) + P1, min_k L_r(p-r, k) + P2)
//                                   - min_k L_r(p-r, k),   L_r = 0 outside the matched area, stored as 16-bit
// walked from (x, y) in steps of (dx, dy) until it leaves the matched area.
SGBM_HD void walk_path(const int16_t* C, int16_t* L, const Dims& dm, int x, int y, int dx, int dy) {
  int16_t prev[kMaxD + 2];
  const int D = dm.D;
  for (int d = 0; d < D + 2; ++d) prev[d] = 0;
  int prev_min = 0;
  while (x >= 0 && x < dm.W1 && y >= 0 && y < dm.H) {
    const size_t o = ((size_t)y * dm.W1 + x) * D;
    const int delta = prev_min + dm.P2;
    prev[0] = (int16_t)kMaxCost; prev[D + 1] = (int16_t)kMaxCost;
    int mn = kMaxCost;
    int16_t cur[kMaxD];
    for (int d = 0; d < D; ++d) {
      const int l = (int)C[o + d] + imin(imin((int)prev[d + 1], (int)prev[d] + dm.P1), imin((int)prev[d + 2] + dm.P1, delta)) - prev_min;
      cur[d] = (int16_t)l;
      mn = imin(mn, l);
    }
    for (int d = 0; d < D; ++d) { L[o + d] = cur[d]; prev[d + 1] = cur[d]; }
    prev_min = (int16_t)mn;
    x += dx; y += dy;
  }
}

SGBM_HD int sat16(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

// ---- stage 6 (per image row): total cost, first-minimum disparity, uniqueness test, sub-pixel refinement and the
// left-right check, sequentially from right to left exactly like the reference (the order decides ties in disp2).
// L0..L3 = left, top-left, top, top-right; Lr = right.  scratch: 2*W ints.  out: W int16 (before the median).
SGBM_HD void select_row(const int16_t* L0, const int16_t* L1, const int16_t* L2, const int16_t* L3, const int16_t* Lr, const Dims& dm,
                        int y, int32_t* scratch, int16_t* out) {
  const int W = dm.W, D = dm.D, INVALID = -kDispScale;
  int32_t* disp2 = scratch; int32_t* disp2cost = scratch + W;
  for (int x = 0; x < W; ++x) { out[x] = (int16_t)INVALID; disp2[x] = INVALID; disp2cost[x] = kMaxCost; }
  for (int x = dm.W1 - 1; x >= 0; --x) {
    const size_t o = ((size_t)y * dm.W1 + x) * D;
    int16_t Sp[kMaxD];
    int minS = kMaxCost, best = -1;
    for (int d = 0; d < D; ++d) {
      const int s4 = sat16((int)L0[o + d] + (int)L1[o + d] + (int)L2[o + d] + (int)L3[o + d]);
      const int sv = sat16(s4 + (int)Lr[o + d]);
      Sp[d] = (int16_t)sv;
      if (sv < minS) { minS = sv; best = d; }
    }
    int d = 0;
    for (; d < D; ++d)
      if ((int)Sp[d] * (100 - dm.uniq) < minS * 100 && (best - d > 1 || d - best > 1)) break;
    if (d < D) continue;
    d = best;
    const int x2 = x + dm.minX1 - d;
    if (disp2cost[x2] > minS) { disp2cost[x2] = minS; disp2[x2] = d; }
    if (0 < d && d < D - 1) {
      const int denom2 = imax((int)Sp[d - 1] + (int)Sp[d + 1] - 2 * (int)Sp[d], 1);
      d = d * kDispScale + (((int)Sp[d - 1] - (int)Sp[d + 1]) * kDispScale + denom2) / (denom2 * 2);
    } else d *= kDispScale;
    out[x + dm.minX1] = (int16_t)d;
  }
  for (int x = dm.minX1; x < W; ++x) {
    const int d1 = out[x];
    if (d1 == INVALID) continue;
    const int dl = d1 >> kDispShift, dh = (d1 + kDispScale - 1) >> kDispShift;
    const int xl = x - dl, xh = x - dh;
    const int a = disp2[xl >= 0 && xl < W ? xl : 0], b = disp2[xh >= 0 && xh < W ? xh : 0];
    if (0 <= xl && xl < W && a >= 0 && (a - dl > dm.d12 || dl - a > dm.d12) &&
        0 <= xh && xh < W && b >= 0 && (b - dh > dm.d12 || dh - b > dm.d12))
      out[x] = (int16_t)INVALID;
  }
}

// ---- stage 7 (per pixel): cv::medianBlur 3x3 on CV_16S, BORDER_REPLICATE
SGBM_HD int16_t median3(const int16_t* img, const Dims& dm, int y, int x) {
  int v[9], n = 0;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) v[n++] = img[(size_t)imin(imax(y + dy, 0), dm.H - 1) * dm.W + imin(imax(x + dx, 0), dm.W - 1)];
  for (int i = 1; i < 9; ++i) { int k = v[i], j = i - 1; while (j >= 0 && v[j] > k) { v[j + 1] = v[j]; --j; } v[j + 1] = k; }
  return (int16_t)v[4];
}

// parameter normalisation of StereoSGBM::compute (stereosgbm.cpp: computeDisparitySGBM preamble)
SGBM_HD Dims make_dims(int W, int H, int num_disparities, int block_size, int P1, int P2, int disp12_max_diff, int pre_filter_cap,
                       int uniqueness_ratio) {
  Dims dm;
  dm.W = W; dm.H = H; dm.D = num_disparities; dm.minX1 = num_disparities; dm.W1 = W - num_disparities;
  const int sad = block_size > 0 ? block_size : 5;
  dm.SW2 = dm.SH2 = sad / 2;
  dm.P1 = P1 > 0 ? P1 : 2;
  dm.P2 = imax(P2 > 0 ? P2 : 5, dm.P1 + 1);
  dm.uniq = uniqueness_ratio >= 0 ? uniqueness_ratio : 10;
  dm.d12 = disp12_max_diff > 0 ? disp12_max_diff : 1;
  dm.ftzero = imax(pre_filter_cap, 15) | 1;
  return dm;
}

}  // namespace esvo_sgbm
