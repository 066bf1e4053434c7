// esvo_b200 product code -- one-time host setup (not a kernel):
//   * rectification tables of each camera: the maps of cv::initUndistortRectifyMap, the
//     raw->rectified LUT of cv::undistortPoints and the undistort-rectify validity mask,
//     as the reference computes them in TimeSurface::cameraInfoCallback
//     (esvo_time_surface/src/TimeSurface.cpp:313-401) and
//     PerspectiveCamera::preComputeRectifiedCoordinate (esvo_core/src/container/CameraSystem.cpp:37-112);
//   * stereo baseline (CameraSystem.cpp:161-166) and parameter defaults.
// Integrators who need bit-identical tables to their OpenCV build can override them with
// esvo_set_rectify_tables().
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace esvo {
namespace {

struct M3 {
  double a[9];
  double operator()(int r, int c) const { return a[r * 3 + c]; }
};
M3 mul3(const M3& A, const M3& B) {
  M3 C;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) C.a[r * 3 + c] = A(r, 0) * B(0, c) + A(r, 1) * B(1, c) + A(r, 2) * B(2, c);
  return C;
}
M3 inv3(const M3& A) {
  M3 B;
  const double* m = A.a;
  double c0 = m[4] * m[8] - m[5] * m[7], c1 = m[5] * m[6] - m[3] * m[8], c2 = m[3] * m[7] - m[4] * m[6];
  double idet = 1.0 / (m[0] * c0 + m[1] * c1 + m[2] * c2);
  B.a[0] = c0 * idet; B.a[3] = c1 * idet; B.a[6] = c2 * idet;
  B.a[1] = (m[2] * m[7] - m[1] * m[8]) * idet; B.a[4] = (m[0] * m[8] - m[2] * m[6]) * idet; B.a[7] = (m[1] * m[6] - m[0] * m[7]) * idet;
  B.a[2] = (m[1] * m[5] - m[2] * m[4]) * idet; B.a[5] = (m[2] * m[3] - m[0] * m[5]) * idet; B.a[8] = (m[0] * m[4] - m[1] * m[3]) * idet;
  return B;
}
M3 proj33(const double* P) { return M3{{P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10]}}; }
M3 as3(const double* p) { M3 m; std::memcpy(m.a, p, sizeof(m.a)); return m; }

// normalised camera ray -> distorted pixel
struct Distorter {
  const HostCamera& c;
  void operator()(double x, double y, double w, double& u, double& v) const {
    const double fx = c.K[0], fy = c.K[4], u0 = c.K[2], v0 = c.K[5];
    if (!c.equidistant) {
      double iw = 1.0 / w, xn = x * iw, yn = y * iw;
      double xx = xn * xn, yy = yn * yn, r2 = xx + yy, xy2 = 2 * xn * yn;
      double radial = 1 + ((0 * r2 + c.D[1]) * r2 + c.D[0]) * r2;
      u = fx * (xn * radial + c.D[2] * xy2 + c.D[3] * (r2 + 2 * xx)) + u0;
      v = fy * (yn * radial + c.D[2] * (r2 + 2 * yy) + c.D[3] * xy2) + v0;
    } else {
      if (w <= 0) { u = x > 0 ? -INFINITY : INFINITY; v = y > 0 ? -INFINITY : INFINITY; return; }
      double xn = x / w, yn = y / w, r = std::sqrt(xn * xn + yn * yn), th = std::atan(r);
      double t2 = th * th, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
      double thd = th * (1 + c.D[0] * t2 + c.D[1] * t4 + c.D[2] * t6 + c.D[3] * t8);
      double sc = (r == 0) ? 1.0 : thd / r;
      u = fx * xn * sc + u0; v = fy * yn * sc + v0;
    }
  }
};

// distorted pixel -> normalised undistorted ray (OpenCV's fixed-point / Newton inversions)
void undistort_ray(const HostCamera& c, double px, double py, double& x, double& y) {
  const double fx = c.K[0], fy = c.K[4], cx = c.K[2], cy = c.K[5];
  if (!c.equidistant) {
    const double x0 = (px - cx) * (1.0 / fx), y0 = (py - cy) * (1.0 / fy);
    x = x0; y = y0;
    for (int it = 0; it < 5; ++it) {  // TermCriteria(COUNT, 5, 0.01)
      double r2 = x * x + y * y;
      double icd = 1.0 / (1 + ((0 * r2 + c.D[1]) * r2 + c.D[0]) * r2);
      if (icd < 0) { x = x0; y = y0; break; }
      double dx = 2 * c.D[2] * x * y + c.D[3] * (r2 + 2 * x * x);
      double dy = c.D[2] * (r2 + 2 * y * y) + 2 * c.D[3] * x * y;
      x = (x0 - dx) * icd; y = (y0 - dy) * icd;
    }
  } else {
    double wx = (px - cx) / fx, wy = (py - cy) / fy;
    double thd = std::min(std::max(-M_PI / 2., std::sqrt(wx * wx + wy * wy)), M_PI / 2.);
    double th = thd, scale = 0.0;
    bool conv = false;
    if (std::fabs(thd) > 1e-8) {
      for (int j = 0; j < 10; ++j) {
        double t2 = th * th, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
        double a = c.D[0] * t2, b = c.D[1] * t4, cc = c.D[2] * t6, d = c.D[3] * t8;
        double fix = (th * (1 + a + b + cc + d) - thd) / (1 + 3 * a + 5 * b + 7 * cc + 9 * d);
        th -= fix;
        if (std::fabs(fix) < 1e-8) { conv = true; break; }
      }
      scale = std::tan(th) / thd;
    } else conv = true;
    bool flipped = (thd < 0 && th > 0) || (thd > 0 && th < 0);
    if (conv && !flipped) { x = wx * scale; y = wy * scale; }
    else { x = y = -1000000.0; }
  }
}

inline int sat16(int v) { return std::max(-32768, std::min(32767, v)); }

}  // namespace

void host_camera_init(HostCamera& cam, const esvo_calib& c) {
  cam.W = c.width; cam.H = c.height; cam.equidistant = c.distortion_model == ESVO_DIST_EQUIDISTANT;
  std::memcpy(cam.K, c.K, sizeof(cam.K)); std::memcpy(cam.D, c.D, sizeof(cam.D));
  std::memcpy(cam.R, c.R, sizeof(cam.R)); std::memcpy(cam.P, c.P, sizeof(cam.P));
  const int W = cam.W, H = cam.H;
  const size_t n = (size_t)W * H;
  cam.map1.resize(n); cam.map2.resize(n); cam.lut.resize(2 * n); cam.mask.resize(n);
  const M3 PR = mul3(proj33(cam.P), as3(cam.R));
  const M3 iPR = inv3(PR);
  Distorter dist{cam};
  // rectified pixel -> raw pixel (maps, stored as float like CV_32FC1)
  for (int v = 0; v < H; ++v) {
    double x = v * iPR(0, 1) + iPR(0, 2), y = v * iPR(1, 1) + iPR(1, 2), w = v * iPR(2, 1) + iPR(2, 2);
    for (int u = 0; u < W; ++u, x += iPR(0, 0), y += iPR(1, 0), w += iPR(2, 0)) {
      double ru, rv;
      dist(x, y, w, ru, rv);
      cam.map1[(size_t)v * W + u] = (float)ru;
      cam.map2[(size_t)v * W + u] = (float)rv;
    }
  }
  // raw pixel -> rectified coordinates (Point2f precision, widened to double)
  for (int v = 0; v < H; ++v)
    for (int u = 0; u < W; ++u) {
      double x, y;
      undistort_ray(cam, (double)(float)u, (double)(float)v, x, y);
      double X = PR(0, 0) * x + PR(0, 1) * y + PR(0, 2), Y = PR(1, 0) * x + PR(1, 1) * y + PR(1, 2);
      double iw = 1.0 / (PR(2, 0) * x + PR(2, 1) * y + PR(2, 2));
      cam.lut[2 * ((size_t)v * W + u)] = (double)(float)(X * iw);
      cam.lut[2 * ((size_t)v * W + u) + 1] = (double)(float)(Y * iw);
    }
  // validity mask: bilinear remap of an all-ones float image (border 0) thresholded
  const float thr = cam.equidistant ? 0.1f : 0.999f;
  for (size_t i = 0; i < n; ++i) {
    int sx = (int)std::lrintf(cam.map1[i] * 32.0f), sy = (int)std::lrintf(cam.map2[i] * 32.0f);
    float ax = (sx & 31) * (1.0f / 32), ay = (sy & 31) * (1.0f / 32);
    int ix = sat16(sx >> 5), iy = sat16(sy >> 5);
    auto in = [&](int xx, int yy) { return (xx >= 0 && xx < W && yy >= 0 && yy < H) ? 1.0f : 0.0f; };
    float val = in(ix, iy) * ((1 - ax) * (1 - ay)) + in(ix + 1, iy) * (ax * (1 - ay)) +
                in(ix, iy + 1) * ((1 - ax) * ay) + in(ix + 1, iy + 1) * (ax * ay);
    cam.mask[i] = val > thr ? 255 : 0;
  }
}

double host_baseline(const HostCamera& right) {
  M3 Pi = inv3(proj33(right.P));
  double t[3];
  for (int i = 0; i < 3; ++i) t[i] = Pi(i, 0) * right.P[3] + Pi(i, 1) * right.P[7] + Pi(i, 2) * right.P[11];
  return std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
}

void host_default_params(esvo_params* p) {
  std::memset(p, 0, sizeof(*p));
  // esvo_time_surface/src/TimeSurface.cpp:23-30
  p->decay_ms = 30; p->ignore_polarity = 1; p->median_blur_kernel_size = 1; p->max_event_queue_len = 20;
  p->time_surface_mode = ESVO_TS_BACKWARD;
  // esvo_core/src/esvo_Mapping.cpp:36-94
  p->patch_size_x = 25; p->patch_size_y = 25; p->bm_min_disparity = 3; p->bm_max_disparity = 40; p->bm_step = 1;
  p->bm_updown = 0; p->smooth_time_surface = 0; p->bm_zncc_threshold = 0.1;
  p->lsnorm = ESVO_LSNORM_TDIST; p->max_iteration = 10; p->td_nu = 0; p->td_scale = 0;
  p->invdepth_min_range = 0.16; p->invdepth_max_range = 2.0; p->residual_vis_threshold = 15;
  p->stdvar_vis_threshold = 0.005; p->age_vis_threshold = 0; p->fusion_radius = 0;
  p->fusion_strategy = ESVO_FUSION_CONST_FRAMES; p->max_num_fusion_frames = 10; p->max_num_fusion_points = 2000;
  p->regularization = 0; p->reg_radius = 5; p->reg_min_neighbours = 8; p->reg_min_close_neighbours = 8;
  // esvo_core/src/esvo_Tracking.cpp:24-36
  p->trk_patch_size_x = 25; p->trk_patch_size_y = 25; p->trk_kernel_size = 15; p->trk_lsnorm = ESVO_TRK_LSNORM_L2;
  p->trk_huber_threshold = 10.0; p->trk_max_registration_points = 500; p->trk_batch_size = 200;
  p->trk_max_iteration = 10; p->trk_min_num_events = 1000;
  p->num_thread_mapping = 4;  // utils.h:36
}

}  // namespace esvo
