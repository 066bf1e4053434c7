// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into or called from the product path.
//
// Shared state of the CPU restatement: camera system, parameters, small math helpers.
// Follows esvo_core/src/container/CameraSystem.cpp and the esvo_Mapping constructor.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../include/esvo_b200.h"
#include "ocv_ops.h"

namespace oracle {

// ros::Time::toSec() / ros::Duration::toSec(): (double)sec + 1e-9*(double)nsec with
// nsec normalised into [0,1e9)  (SURVEY A.3).
static inline double ns_to_sec(int64_t ns) {
  int64_t sec = ns / 1000000000LL, nsec = ns % 1000000000LL;
  if (nsec < 0) { nsec += 1000000000LL; sec -= 1; }
  return (double)sec + 1e-9 * (double)nsec;
}

struct Mat4 {
  double m[16];
  double& operator()(int r, int c) { return m[r * 4 + c]; }
  double operator()(int r, int c) const { return m[r * 4 + c]; }
  static Mat4 identity() {
    Mat4 I; std::memset(I.m, 0, sizeof(I.m));
    I.m[0] = I.m[5] = I.m[10] = I.m[15] = 1; return I;
  }
  static Mat4 from(const double* p) { Mat4 M; std::memcpy(M.m, p, sizeof(M.m)); return M; }
};
static inline Mat4 mul(const Mat4& A, const Mat4& B) {
  Mat4 C;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += A(i, k) * B(k, j);
      C(i, j) = s;
    }
  return C;
}
// rigid inverse: kindr QuatTransformation::inverse() = (R^T, -R^T t)
static inline Mat4 rigid_inverse(const Mat4& T) {
  Mat4 I = Mat4::identity();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) I(i, j) = T(j, i);
  for (int i = 0; i < 3; ++i) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += I(i, k) * T(k, 3);
    I(i, 3) = -s;
  }
  return I;
}
// general 4x4 inverse by cofactors (Eigen's fixed-size 4x4 inverse is cofactor based too);
// used by cam2World exactly like CameraSystem.cpp:120-139.
static inline bool inverse4(const double* m, double* inv) {
  double t[16];
  t[0] = m[5]*m[10]*m[15] - m[5]*m[11]*m[14] - m[9]*m[6]*m[15] + m[9]*m[7]*m[14] + m[13]*m[6]*m[11] - m[13]*m[7]*m[10];
  t[4] = -m[4]*m[10]*m[15] + m[4]*m[11]*m[14] + m[8]*m[6]*m[15] - m[8]*m[7]*m[14] - m[12]*m[6]*m[11] + m[12]*m[7]*m[10];
  t[8] = m[4]*m[9]*m[15] - m[4]*m[11]*m[13] - m[8]*m[5]*m[15] + m[8]*m[7]*m[13] + m[12]*m[5]*m[11] - m[12]*m[7]*m[9];
  t[12] = -m[4]*m[9]*m[14] + m[4]*m[10]*m[13] + m[8]*m[5]*m[14] - m[8]*m[6]*m[13] - m[12]*m[5]*m[10] + m[12]*m[6]*m[9];
  t[1] = -m[1]*m[10]*m[15] + m[1]*m[11]*m[14] + m[9]*m[2]*m[15] - m[9]*m[3]*m[14] - m[13]*m[2]*m[11] + m[13]*m[3]*m[10];
  t[5] = m[0]*m[10]*m[15] - m[0]*m[11]*m[14] - m[8]*m[2]*m[15] + m[8]*m[3]*m[14] + m[12]*m[2]*m[11] - m[12]*m[3]*m[10];
  t[9] = -m[0]*m[9]*m[15] + m[0]*m[11]*m[13] + m[8]*m[1]*m[15] - m[8]*m[3]*m[13] - m[12]*m[1]*m[11] + m[12]*m[3]*m[9];
  t[13] = m[0]*m[9]*m[14] - m[0]*m[10]*m[13] - m[8]*m[1]*m[14] + m[8]*m[2]*m[13] + m[12]*m[1]*m[10] - m[12]*m[2]*m[9];
  t[2] = m[1]*m[6]*m[15] - m[1]*m[7]*m[14] - m[5]*m[2]*m[15] + m[5]*m[3]*m[14] + m[13]*m[2]*m[7] - m[13]*m[3]*m[6];
  t[6] = -m[0]*m[6]*m[15] + m[0]*m[7]*m[14] + m[4]*m[2]*m[15] - m[4]*m[3]*m[14] - m[12]*m[2]*m[7] + m[12]*m[3]*m[6];
  t[10] = m[0]*m[5]*m[15] - m[0]*m[7]*m[13] - m[4]*m[1]*m[15] + m[4]*m[3]*m[13] + m[12]*m[1]*m[7] - m[12]*m[3]*m[5];
  t[14] = -m[0]*m[5]*m[14] + m[0]*m[6]*m[13] + m[4]*m[1]*m[14] - m[4]*m[2]*m[13] - m[12]*m[1]*m[6] + m[12]*m[2]*m[5];
  t[3] = -m[1]*m[6]*m[11] + m[1]*m[7]*m[10] + m[5]*m[2]*m[11] - m[5]*m[3]*m[10] - m[9]*m[2]*m[7] + m[9]*m[3]*m[6];
  t[7] = m[0]*m[6]*m[11] - m[0]*m[7]*m[10] - m[4]*m[2]*m[11] + m[4]*m[3]*m[10] + m[8]*m[2]*m[7] - m[8]*m[3]*m[6];
  t[11] = -m[0]*m[5]*m[11] + m[0]*m[7]*m[9] + m[4]*m[1]*m[11] - m[4]*m[3]*m[9] - m[8]*m[1]*m[7] + m[8]*m[3]*m[5];
  t[15] = m[0]*m[5]*m[10] - m[0]*m[6]*m[9] - m[4]*m[1]*m[10] + m[4]*m[2]*m[9] + m[8]*m[1]*m[6] - m[8]*m[2]*m[5];
  double det = m[0]*t[0] + m[1]*t[4] + m[2]*t[8] + m[3]*t[12];
  if (det == 0) return false;
  det = 1.0 / det;
  for (int i = 0; i < 16; ++i) inv[i] = t[i] * det;
  return true;
}

struct Camera {
  int W = 0, H = 0;
  bool equidistant = false;
  double K[9], D[4], R[9], P[12];
  std::vector<float> map1, map2;      // initUndistortRectifyMap (CV_32FC1)
  std::vector<double> lut;            // precomputed_rectified_points_ (x,y per raw pixel)
  std::vector<uint8_t> mask;          // UndistortRectify_mask_ {0,255}

  void init(const esvo_calib& c) {
    W = c.width; H = c.height; equidistant = c.distortion_model == ESVO_DIST_EQUIDISTANT;
    std::memcpy(K, c.K, sizeof(K)); std::memcpy(D, c.D, sizeof(D));
    std::memcpy(R, c.R, sizeof(R)); std::memcpy(P, c.P, sizeof(P));
    size_t n = (size_t)W * H;
    map1.resize(n); map2.resize(n); lut.resize(2 * n); mask.resize(n);
    init_undistort_rectify_map(K, D, R, P, W, H, equidistant, map1.data(), map2.data());
    undistort_points(K, D, R, P, W, H, equidistant, lut.data());
    undistort_rectify_mask(map1.data(), map2.data(), W, H, equidistant, mask.data());
  }
  // PerspectiveCamera::cam2World (CameraSystem.cpp:120-139)
  void cam2World(const double x[2], double invDepth, double p[3]) const {
    double z = 1.0 / invDepth;
    double Pt[16], Pi[16];
    std::memcpy(Pt, P, 12 * sizeof(double));
    Pt[12] = 0; Pt[13] = 0; Pt[14] = 0; Pt[15] = z;
    inverse4(Pt, Pi);
    double xs[4] = {x[0], x[1], 1, 1}, ps[4];
    for (int i = 0; i < 4; ++i) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += (z * Pi[i * 4 + k]) * xs[k];
      ps[i] = s;
    }
    p[0] = ps[0] / ps[3]; p[1] = ps[1] / ps[3]; p[2] = ps[2] / ps[3];
  }
  // PerspectiveCamera::world2Cam (CameraSystem.cpp:141-148)
  void world2Cam(const double p[3], double x[2]) const {
    double h[3];
    for (int i = 0; i < 3; ++i) h[i] = P[i * 4 + 0] * p[0] + P[i * 4 + 1] * p[1] + P[i * 4 + 2] * p[2] + P[i * 4 + 3];
    x[0] = h[0] / h[2]; x[1] = h[1] / h[2];
  }
};

struct CameraSystem {
  Camera left, right;
  double baseline = 0;
  void init(const esvo_calib& l, const esvo_calib& r) {
    left.init(l); right.init(r);
    // CameraSystem::computeBaseline (CameraSystem.cpp:161-166)
    double P33[9] = {right.P[0], right.P[1], right.P[2], right.P[4], right.P[5], right.P[6],
                     right.P[8], right.P[9], right.P[10]};
    double Pi[9];
    mat3_inv(P33, Pi);
    double t[3];
    for (int i = 0; i < 3; ++i) t[i] = Pi[i * 3 + 0] * right.P[3] + Pi[i * 3 + 1] * right.P[7] + Pi[i * 3 + 2] * right.P[11];
    baseline = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  }
};

// Disparity clip of the esvo_Mapping constructor (esvo_Mapping.cpp:110-116).
static inline void clip_disparity(const CameraSystem& cs, const esvo_params& p, size_t& dmin, size_t& dmax) {
  double f = (cs.left.P[0] + cs.left.P[5]) / 2;
  double b = cs.baseline;
  size_t minD = std::max(size_t(std::floor(f * b * p.invdepth_min_range)), (size_t)0);
  size_t maxD = size_t(std::ceil(f * b * p.invdepth_max_range));
  dmin = std::max(minD, (size_t)p.bm_min_disparity);
  dmax = std::min(maxD, (size_t)p.bm_max_disparity);
}

static inline void default_params(esvo_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->decay_ms = 30; p->ignore_polarity = 1; p->median_blur_kernel_size = 1;
  p->max_event_queue_len = 20; p->time_surface_mode = ESVO_TS_BACKWARD;
  p->patch_size_x = 25; p->patch_size_y = 25; p->bm_min_disparity = 3; p->bm_max_disparity = 40;
  p->bm_step = 1; p->bm_updown = 0; p->smooth_time_surface = 0; p->bm_zncc_threshold = 0.1;
  p->lsnorm = ESVO_LSNORM_TDIST; p->max_iteration = 10; p->td_nu = 0; p->td_scale = 0;
  p->invdepth_min_range = 0.16; p->invdepth_max_range = 2.0; p->residual_vis_threshold = 15;
  p->stdvar_vis_threshold = 0.005; p->age_vis_threshold = 0; p->fusion_radius = 0;
  p->fusion_strategy = ESVO_FUSION_CONST_FRAMES; p->max_num_fusion_frames = 10;
  p->max_num_fusion_points = 2000; p->regularization = 0; p->reg_radius = 5;
  p->reg_min_neighbours = 8; p->reg_min_close_neighbours = 8;
  p->trk_patch_size_x = 25; p->trk_patch_size_y = 25; p->trk_kernel_size = 15;
  p->trk_lsnorm = ESVO_TRK_LSNORM_L2; p->trk_huber_threshold = 10.0;
  p->trk_max_registration_points = 500; p->trk_batch_size = 200; p->trk_max_iteration = 10;
  p->trk_min_num_events = 1000; p->num_thread_mapping = 4;
}

// POD DepthPoint with the reference's update rules (DepthPoint.cpp).
struct DepthPoint {
  int64_t row = 0, col = 0;
  double x[2] = {0.5, 0.5};
  double invDepth = -1.0, scale2 = 0, nu = 0, variance = 0, residual = 0;
  int64_t age = 0;
  double p_cam[3] = {0, 0, 0};
  Mat4 T_world_cam = Mat4::identity();
  DepthPoint() {}
  DepthPoint(int64_t r, int64_t c) : row(r), col(c) { x[0] = c + 0.5; x[1] = r + 0.5; }
  bool valid() const { return invDepth > -1e-6; }
  bool valid(double var_thr, double age_thr, double rho_max, double rho_min) const {
    return invDepth > -1e-6 && (double)age >= age_thr && variance <= var_thr && invDepth <= rho_max &&
           invDepth >= rho_min;
  }
  // DepthPoint::update (DepthPoint.cpp:145-164)
  void update(double rho, double var) {
    if (invDepth > -1e-6) {
      double temp = invDepth;
      invDepth = (variance * rho + var * temp) / (variance + var);
      temp = variance;
      variance = (temp * var) / (temp + var);
    } else { invDepth = rho; variance = var; }
    if (variance < 1e-6) variance = 1e-6;  // boundVariance
  }
  // DepthPoint::update_studentT (DepthPoint.cpp:166-188)
  void update_studentT(double rho, double s2, double var, double nu_in) {
    if (invDepth > -1e-6) {
      double nu_u = std::min(nu_in, nu);
      double rho_u = (s2 * invDepth + scale2 * rho) / (scale2 + s2);
      double d = invDepth - rho;
      double s2_u = (nu_u + (d * d) / (scale2 + s2)) / (nu_u + 1) * (scale2 * s2) / (scale2 + s2);
      invDepth = rho_u; scale2 = s2_u; nu = nu_u + 1;
      variance = nu / (nu - 2) * scale2;
      age++;
    } else { invDepth = rho; scale2 = s2; variance = var; nu = nu_in; }
  }
  // DepthPoint::copy -- everything but the location
  void copy_from(const DepthPoint& o) {
    invDepth = o.invDepth; variance = o.variance; scale2 = o.scale2; nu = o.nu;
    x[0] = o.x[0]; x[1] = o.x[1]; std::memcpy(p_cam, o.p_cam, sizeof(p_cam));
    T_world_cam = o.T_world_cam; residual = o.residual; age = o.age;
  }
};

static inline void to_pod(const DepthPoint& d, esvo_depth_point* o) {
  o->row = (int32_t)d.row; o->col = (int32_t)d.col; o->x[0] = d.x[0]; o->x[1] = d.x[1];
  o->inv_depth = d.invDepth; o->scale2 = d.scale2; o->nu = d.nu; o->variance = d.variance;
  o->residual = d.residual; o->age = d.age; std::memcpy(o->p_cam, d.p_cam, sizeof(d.p_cam));
  std::memcpy(o->T_world_cam, d.T_world_cam.m, sizeof(d.T_world_cam.m));
}
static inline DepthPoint from_pod(const esvo_depth_point& o) {
  DepthPoint d; d.row = o.row; d.col = o.col; d.x[0] = o.x[0]; d.x[1] = o.x[1];
  d.invDepth = o.inv_depth; d.scale2 = o.scale2; d.nu = o.nu; d.variance = o.variance;
  d.residual = o.residual; d.age = o.age; std::memcpy(d.p_cam, o.p_cam, sizeof(d.p_cam));
  d.T_world_cam = Mat4::from(o.T_world_cam);
  return d;
}

}  // namespace oracle
