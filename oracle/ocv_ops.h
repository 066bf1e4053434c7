// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into or called from the product path.
//
// CPU restatement of the OpenCV operations the reference hot path calls (OpenCV itself is an
// un-vendored third-party dependency: dependencies.yaml / README.md:24-26 pin "3.2 or 4.x").
// Each function names the reference call site it serves and the OpenCV semantics it restates;
// tests/test_oracle_image_ops.py pins every one of them against Python cv2 4.13 in this image.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace oracle {

// cv::Mat::convertTo(CV_8U) from f64 = saturate_cast<uchar>(cvRound(v)); cvRound is
// round-half-to-even (lrint).  Call site: TimeSurface.cpp:127.
static inline uint8_t cvt_u8(double v) {
  long r = std::lrint(v);  // default rounding mode = RNE
  return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

static inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * (len - 1) - p;
  }
  return p;
}

// cv::medianBlur(u8, ksize=3): BORDER_REPLICATE.  Call site: TimeSurface.cpp:131.
// Generic odd ksize supported (the cfg always yields 3).
static inline void median_blur_u8(const uint8_t* src, uint8_t* dst, int W, int H, int ksize) {
  const int r = ksize / 2;
  std::vector<uint8_t> win((size_t)ksize * ksize);
  std::vector<uint8_t> out((size_t)W * H);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int n = 0;
      for (int dy = -r; dy <= r; ++dy)
        for (int dx = -r; dx <= r; ++dx) {
          int yy = std::min(std::max(y + dy, 0), H - 1);
          int xx = std::min(std::max(x + dx, 0), W - 1);
          win[n++] = src[(size_t)yy * W + xx];
        }
      std::nth_element(win.begin(), win.begin() + n / 2, win.begin() + n);
      out[(size_t)y * W + x] = win[n / 2];
    }
  std::memcpy(dst, out.data(), out.size());
}

// cv::remap(u8, INTER_LINEAR, CV_32FC1 maps, BORDER_CONSTANT(0)).  Call site: TimeSurface.cpp:149.
// OpenCV quantises the map to 1/32 px (INTER_BITS=5): sx = cvRound(mapx*32); ix = sx>>5;
// fx = sx&31; weights from a 32x32 table scaled by 2^15 (exact here: 32*(32-fx)*(32-fy) etc.),
// dst = (sum w*src + 2^14) >> 15, with out-of-image taps replaced by the border value 0.
// The integer coordinate is saturated to short like OpenCV does.
static inline int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }
static inline void remap_bilinear_u8(const uint8_t* src, uint8_t* dst, int W, int H,
                                     const float* mapx, const float* mapy) {
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      size_t i = (size_t)y * W + x;
      int sx = (int)std::lrintf(mapx[i] * 32.0f);
      int sy = (int)std::lrintf(mapy[i] * 32.0f);
      int fx = sx & 31, fy = sy & 31;
      int ix = sat_short(sx >> 5), iy = sat_short(sy >> 5);
      int w00 = 32 * (32 - fx) * (32 - fy), w01 = 32 * fx * (32 - fy);
      int w10 = 32 * (32 - fx) * fy, w11 = 32 * fx * fy;
      auto px = [&](int xx, int yy) -> int {
        return (xx >= 0 && xx < W && yy >= 0 && yy < H) ? src[(size_t)yy * W + xx] : 0;
      };
      int v = w00 * px(ix, iy) + w01 * px(ix + 1, iy) + w10 * px(ix, iy + 1) + w11 * px(ix + 1, iy + 1);
      dst[i] = (uint8_t)((v + (1 << 14)) >> 15);
    }
}

// cv::remap(f32, INTER_LINEAR, BORDER_CONSTANT(0)) as used for the undistort-rectify validity
// mask (CameraSystem.cpp:63-72): same 1/32 coordinate quantisation, float weights from the
// 32-entry linear table.
static inline void remap_bilinear_f32(const float* src, float* dst, int W, int H, const float* mapx,
                                      const float* mapy) {
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      size_t i = (size_t)y * W + x;
      int sx = (int)std::lrintf(mapx[i] * 32.0f);
      int sy = (int)std::lrintf(mapy[i] * 32.0f);
      int fx = sx & 31, fy = sy & 31;
      int ix = sat_short(sx >> 5), iy = sat_short(sy >> 5);
      float ax = fx * (1.0f / 32), ay = fy * (1.0f / 32);
      float w00 = (1 - ax) * (1 - ay), w01 = ax * (1 - ay), w10 = (1 - ax) * ay, w11 = ax * ay;
      auto px = [&](int xx, int yy) -> float {
        return (xx >= 0 && xx < W && yy >= 0 && yy < H) ? src[(size_t)yy * W + xx] : 0.0f;
      };
      dst[i] = px(ix, iy) * w00 + px(ix + 1, iy) * w01 + px(ix, iy + 1) * w10 + px(ix + 1, iy + 1) * w11;
    }
}

// cv::GaussianBlur(u8, Size(5,5), sigma=0): for ksize<=7 and sigma<=0 OpenCV uses the fixed
// kernel [1 4 6 4 1]/16 per axis; the u8 path runs in 8.8 fixed point: horizontal pass exact,
// vertical pass sum/256 rounded half-up.  BORDER_REFLECT_101.
// Call sites: TimeSurfaceObservation.h:110-113 (SmoothTimeSurface), :125 (negative TS).
// kernel sizes 3 ([1 2 1]/4) and 5 supported; the cfgs only use 5.
static inline void gaussian_blur_u8(const uint8_t* src, uint8_t* dst, int W, int H, int ksize) {
  static const int k3[3] = {1, 2, 1}, k5[5] = {1, 4, 6, 4, 1};
  const int* k = ksize == 3 ? k3 : k5;
  const int r = ksize / 2;
  const int norm = ksize == 3 ? 4 : 16;  // per axis
  std::vector<int> tmp((size_t)W * H);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int s = 0;
      for (int d = -r; d <= r; ++d) s += k[d + r] * src[(size_t)y * W + reflect101(x + d, W)];
      tmp[(size_t)y * W + x] = s;  // scaled by norm
    }
  const int tot = norm * norm;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int s = 0;
      for (int d = -r; d <= r; ++d) s += k[d + r] * tmp[(size_t)reflect101(y + d, H) * W + x];
      dst[(size_t)y * W + x] = (uint8_t)((s + tot / 2) / tot);
    }
}

// cv::Sobel(src f64, CV_64F, dx, dy) ksize 3, scale 1, BORDER_REFLECT_101 (unnormalised).
// Call sites: TimeSurfaceObservation.h:142-143.  dx: [-1 0 1] (x) [1 2 1]^T (y).
static inline void sobel3_f64(const double* src, double* ddx, double* ddy, int W, int H) {
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
      int ym = reflect101(y - 1, H), yp = reflect101(y + 1, H);
      auto S = [&](int xx, int yy) { return src[(size_t)yy * W + xx]; };
      if (ddx)
        ddx[(size_t)y * W + x] =
            (S(xp, ym) - S(xm, ym)) + 2.0 * (S(xp, y) - S(xm, y)) + (S(xp, yp) - S(xm, yp));
      if (ddy)
        ddy[(size_t)y * W + x] =
            (S(xm, yp) - S(xm, ym)) + 2.0 * (S(x, yp) - S(x, ym)) + (S(xp, yp) - S(xp, ym));
    }
}

// ---- 3x3 helpers ----
static inline void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += A[i * 3 + k] * B[k * 3 + j];
      C[i * 3 + j] = s;
    }
}
static inline bool mat3_inv(const double* A, double* B) {
  double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  if (det == 0) return false;
  double id = 1.0 / det;
  B[0] = c00 * id; B[1] = (A[2] * A[7] - A[1] * A[8]) * id; B[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  B[3] = c01 * id; B[4] = (A[0] * A[8] - A[2] * A[6]) * id; B[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  B[6] = c02 * id; B[7] = (A[1] * A[6] - A[0] * A[7]) * id; B[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  return true;
}

// cv::initUndistortRectifyMap(K, D(4), R, P, size, CV_32FC1) (plumb_bob) and
// cv::fisheye::initUndistortRectifyMap (equidistant).  Call sites: TimeSurface.cpp:341-353,
// CameraSystem.cpp:62,82.  For each rectified pixel (u,v): [x y w]^T = (P33*R)^-1 [u v 1]^T,
// normalise, distort, project with K; maps are stored as float.
static inline void init_undistort_rectify_map(const double* K, const double* D, const double* R,
                                              const double* P, int W, int H, bool equidistant,
                                              float* map1, float* map2) {
  double P33[9] = {P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10]};
  double PR[9], iR[9] = {0};
  mat3_mul(P33, R, PR);
  mat3_inv(PR, iR);
  const double fx = K[0], fy = K[4], u0 = K[2], v0 = K[5];
  const double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3];
  for (int i = 0; i < H; ++i) {
    double _x = i * iR[1] + iR[2], _y = i * iR[4] + iR[5], _w = i * iR[7] + iR[8];
    for (int j = 0; j < W; ++j, _x += iR[0], _y += iR[3], _w += iR[6]) {
      double u, v;
      if (!equidistant) {
        double w = 1.0 / _w, x = _x * w, y = _y * w;
        double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
        double kr = 1 + ((0 * r2 + k2) * r2 + k1) * r2;  // k3..k6 = 0 (D has 4 coeffs)
        double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2);
        double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy;
        u = fx * xd + u0;
        v = fy * yd + v0;
      } else {
        // fisheye: theta_d = theta (1 + k1 t^2 + k2 t^4 + k3 t^6 + k4 t^8)
        if (_w <= 0) { u = (_x > 0) ? -INFINITY : INFINITY; v = (_y > 0) ? -INFINITY : INFINITY; }
        else {
          double x = _x / _w, y = _y / _w;
          double r = std::sqrt(x * x + y * y);
          double theta = std::atan(r);
          double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
          double theta_d = theta * (1 + D[0] * t2 + D[1] * t4 + D[2] * t6 + D[3] * t8);
          double scale = (r == 0) ? 1.0 : theta_d / r;
          u = fx * x * scale + u0;
          v = fy * y * scale + v0;
        }
      }
      map1[(size_t)i * W + j] = (float)u;
      map2[(size_t)i * W + j] = (float)v;
    }
  }
}

// cv::undistortPoints(raw Point2f, K, D, R, P) (plumb_bob, 5 fixed-point iterations =
// OpenCV's default TermCriteria(COUNT,5,0.01)) and cv::fisheye::undistortPoints.
// Call sites: TimeSurface.cpp:374-386, CameraSystem.cpp:59,79.  Output is Point2f (float),
// which the reference then widens to double.
static inline void undistort_points(const double* K, const double* D, const double* R,
                                    const double* P, int W, int H, bool equidistant, double* lut_xy) {
  double P33[9] = {P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10]};
  double RR[9];
  mat3_mul(P33, R, RR);
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  const double ifx = 1.0 / fx, ify = 1.0 / fy;
  for (int yy = 0; yy < H; ++yy)
    for (int xx = 0; xx < W; ++xx) {
      double x, y;
      if (!equidistant) {
        double x0 = x = ((double)(float)xx - cx) * ifx;
        double y0 = y = ((double)(float)yy - cy) * ify;
        const double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3];
        for (int it = 0; it < 5; ++it) {
          double r2 = x * x + y * y;
          double icdist = 1.0 / (1 + ((0 * r2 + k2) * r2 + k1) * r2);
          if (icdist < 0) { x = x0; y = y0; break; }
          double deltaX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
          double deltaY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
          x = (x0 - deltaX) * icdist;
          y = (y0 - deltaY) * icdist;
        }
      } else {
        // cv::fisheye::undistortPoints: Newton iterations on theta (<=10, eps 1e-8)
        double pwx = ((double)(float)xx - cx) / fx, pwy = ((double)(float)yy - cy) / fy;
        double theta_d = std::sqrt(pwx * pwx + pwy * pwy);
        theta_d = std::min(std::max(-M_PI / 2., theta_d), M_PI / 2.);
        double scale = 0.0;
        bool converged = false;
        double theta = theta_d;
        if (std::fabs(theta_d) > 1e-8) {
          for (int j = 0; j < 10; ++j) {
            double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
            double k0t2 = D[0] * t2, k1t4 = D[1] * t4, k2t6 = D[2] * t6, k3t8 = D[3] * t8;
            double fix = (theta * (1 + k0t2 + k1t4 + k2t6 + k3t8) - theta_d) /
                         (1 + 3 * k0t2 + 5 * k1t4 + 7 * k2t6 + 9 * k3t8);
            theta = theta - fix;
            if (std::fabs(fix) < 1e-8) { converged = true; break; }
          }
          scale = std::tan(theta) / theta_d;
        } else {
          converged = true;
        }
        bool flipped = (theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0);
        if (converged && !flipped) { x = pwx * scale; y = pwy * scale; }
        else { x = -1000000.0; y = -1000000.0; }
      }
      double X = RR[0] * x + RR[1] * y + RR[2];
      double Y = RR[3] * x + RR[4] * y + RR[5];
      double Wd = 1.0 / (RR[6] * x + RR[7] * y + RR[8]);
      size_t i = (size_t)yy * W + xx;
      lut_xy[2 * i + 0] = (double)(float)(X * Wd);
      lut_xy[2 * i + 1] = (double)(float)(Y * Wd);
    }
}

// UndistortRectify_mask_ (CameraSystem.cpp:63-72 / :83-92): remap(ones f32) -> threshold
// (>0.999 plumb_bob, >0.1 equidistant) -> {0,255}.
static inline void undistort_rectify_mask(const float* map1, const float* map2, int W, int H,
                                          bool equidistant, uint8_t* mask) {
  std::vector<float> ones((size_t)W * H, 1.0f), out((size_t)W * H);
  remap_bilinear_f32(ones.data(), out.data(), W, H, map1, map2);
  const float thr = equidistant ? 0.1f : 0.999f;
  for (size_t i = 0; i < out.size(); ++i) mask[i] = out[i] > thr ? 255 : 0;
}

}  // namespace oracle
