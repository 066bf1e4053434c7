// TEST INFRASTRUCTURE ONLY (see oracle/o_core.h) -- CPU restatement of the event-to-event matcher the reference keeps
// for its comparison modes: esvo_core::core::EventMatcher (esvo_core/src/core/EventMatcher.cpp:51-346, the method of
// Ieng et al. 2018) as esvo_MVStereo drives it (esvo_core/src/esvo_MVStereo.cpp:257-266 modes PURE_EVENT_MATCHING /
// EM_PLUS_ESTIMATION, eventSlicingForEM :1008-1040) and the EventMatchPair -> DepthPoint conversion vEMP2vDP (:1072-1097).
// Parity unpinned against the reference binary (it cannot be built here); pinned by an independent numpy re-derivation
// (tests/indep_numpy.py::event_match, tests/test_indep_pins.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "o_mapping.h"

namespace oracle {

// ros::Time(double) (ros/time.h fromSec): sec = floor(t), nsec = round((t - sec) * 1e9), normalised.
static inline int64_t sec_to_ns(double t) {
  int64_t sec = (int64_t)std::floor(t);
  int64_t nsec = (int64_t)std::llround((t - (double)sec) * 1e9);
  sec += nsec / 1000000000LL; nsec %= 1000000000LL;
  return sec * 1000000000LL + nsec;
}

struct EventMatcher {
  const CameraSystem* cs = nullptr;
  const TsObs* obs = nullptr;
  int NT = 1;                                   // NUM_THREAD_ (output order only)
  double time_thr = 10e-5, epi_thr = 0.5, ncc_thr = 0.1;
  int wx = 25, wy = 5;
  uint64_t n_evals = 0;                         // zncc_cost calls

  // zncc_cost (:253-274): mean-free patches scaled by 1/(norm + 1e-6); patches traversed in storage order
  static double zncc_cost(const double* l, const double* r, size_t n) {
    double sl = 0, sr = 0;
    for (size_t i = 0; i < n; ++i) { sl += l[i]; sr += r[i]; }
    const double ml = sl / (double)n, mr = sr / (double)n;
    double ql = 0, qr = 0;
    for (size_t i = 0; i < n; ++i) { const double a = l[i] - ml, b = r[i] - mr; ql += a * a; qr += b * b; }
    const double nl = std::sqrt(ql) + 1e-6, nr = std::sqrt(qr) + 1e-6;
    double s = 0;
    for (size_t i = 0; i < n; ++i) s += ((l[i] - ml) / nl) * ((r[i] - mr) / nr);
    return 0.5 * (1 - s);
  }
  // warping2 (:276-306)
  bool warping2(const double x[2], double invDepth, const double* T_left_rv /*3x4 of a 4x4*/, double x1[2], double x2[2]) const {
    double p_rv[3], pl[3];
    cs->left.cam2World(x, invDepth, p_rv);
    for (int i = 0; i < 3; ++i)
      pl[i] = T_left_rv[i * 4 + 0] * p_rv[0] + T_left_rv[i * 4 + 1] * p_rv[1] + T_left_rv[i * 4 + 2] * p_rv[2] + T_left_rv[i * 4 + 3];
    cs->left.world2Cam(pl, x1);      // P(:, :3) p + P(:, 3), divided by the third component
    cs->right.world2Cam(pl, x2);
    const int width = cs->left.W, height = cs->left.H;
    if (x1[0] < (wx - 1) / 2 || x1[0] > width - (wx - 1) / 2 || x1[1] < (wy - 1) / 2 || x1[1] > height - (wy - 1) / 2) return false;
    if (x2[0] < (wx - 1) / 2 || x2[0] > width - (wx - 1) / 2 || x2[1] < (wy - 1) / 2 || x2[1] > height - (wy - 1) / 2) return false;
    return true;
  }
  // match_an_event (:60-163).  Right-camera candidates: time-ordered arrays.
  bool match_an_event(uint16_t ex, uint16_t ey, int64_t et, uint8_t epol, const Mat4& T_world_rv, const uint16_t* rx, const uint16_t* ry,
                      const int64_t* rt, const uint8_t* rp, size_t nr, Seed& em) {
    // --- temporal check: ros::Time(ts.toSec() -/+ thr/2), candidates = [lower_bound(low), lower_bound(up)) on toSec() doubles
    const double te = ns_to_sec(et);
    const double lowS = ns_to_sec(sec_to_ns(te - time_thr / 2)), upS = ns_to_sec(sec_to_ns(te + time_thr / 2));
    auto lower = [&](double t) { size_t lo = 0, hi = nr; while (lo < hi) { size_t mid = (lo + hi) / 2; if (ns_to_sec(rt[mid]) < t) lo = mid + 1; else hi = mid; } return lo; };
    const size_t cb = lower(lowS), ce = lower(upS);
    std::vector<size_t> time_ok;
    for (size_t j = cb; j < ce; ++j) {
      const double tj = ns_to_sec(rt[j]);
      if (tj >= lowS && tj <= upS && ((epol != 0) == (rp[j] != 0))) time_ok.push_back(j);
    }
    if (time_ok.empty()) return false;
    // --- epipolar check
    const Camera& L = cs->left; const Camera& R = cs->right;
    const size_t li = (size_t)ey * L.W + ex;
    const double xl[2] = {L.lut[2 * li], L.lut[2 * li + 1]};
    std::vector<size_t> epi_ok;
    for (size_t j : time_ok) {
      const size_t ri = (size_t)ry[j] * R.W + rx[j];
      const double xr0 = R.lut[2 * ri], xr1 = R.lut[2 * ri + 1];
      if (std::fabs(xl[1] - xr1) <= epi_thr && xr0 < xl[0]) epi_ok.push_back(j);
    }
    if (epi_ok.empty()) return false;
    // --- motion check (ZNCC of the two warped time-surface patches)
    const double b = cs->baseline, f = L.P[0];
    double min_cost = 1.0; size_t best = 0; double best_depth = 0;
    const Mat4 T_left_rv = mul(rigid_inverse(obs->tr), T_world_rv);
    const int N = wx * wy;
    std::vector<double> pl(N), pr(N);
    for (size_t q = 0; q < epi_ok.size(); ++q) {
      const size_t j = epi_ok[q];
      const size_t ri = (size_t)ry[j] * R.W + rx[j];
      const double disparity = xl[0] - R.lut[2 * ri];
      const double depth = b * f / disparity;
      double x1[2], x2[2];
      if (!warping2(xl, 1.0 / depth, T_left_rv.m, x1, x2)) continue;
      if (!(patchInterpolation(obs->TS_left.data(), obs->W, obs->H, x1, wx, wy, pl.data()) &&
            patchInterpolation(obs->TS_right.data(), obs->W, obs->H, x2, wx, wy, pr.data()))) continue;
      const double cost = zncc_cost(pl.data(), pr.data(), (size_t)N);
      ++n_evals;
      if (cost < min_cost) { min_cost = cost; best = q; best_depth = depth; }
    }
    if (min_cost > ncc_thr) return false;
    const size_t jb = epi_ok[best];
    const size_t rb = (size_t)ry[jb] * R.W + rx[jb];
    em.x_left_raw[0] = ex; em.x_left_raw[1] = ey;
    em.x_left[0] = xl[0]; em.x_left[1] = xl[1];
    em.x_right[0] = R.lut[2 * rb]; em.x_right[1] = R.lut[2 * rb + 1];
    em.t_ns = et; em.trans = T_world_rv;
    em.invDepth = 1.0 / best_depth; em.cost = min_cost; em.disp = em.x_left[0] - em.x_right[0];
    return true;
  }
  // match_all_HyperThread (:185-231) + match (:233-251): NT interleaved jobs over the events of all slices (contiguous from the
  // first slice's begin), per-thread results concatenated.  slice_counts = EventSlice::numEvents_, slice_poses = transf_.
  void match_all(const uint16_t* lx, const uint16_t* ly, const int64_t* lt, const uint8_t* lp, size_t nl, const int32_t* slice_counts,
                 const double* slice_poses, size_t n_slices, const uint16_t* rx, const uint16_t* ry, const int64_t* rt, const uint8_t* rp,
                 size_t nr, std::vector<Seed>& vEMP) {
    vEMP.clear(); n_evals = 0;
    std::vector<int32_t> slice_of;
    for (size_t s = 0; s < n_slices; ++s) slice_of.insert(slice_of.end(), (size_t)std::max(slice_counts[s], 0), (int32_t)s);
    const size_t total = std::min(slice_of.size(), nl);
    for (int tid = 0; tid < NT; ++tid)
      for (size_t i = (size_t)tid; i < total; i += (size_t)NT) {
        Seed em;
        if (match_an_event(lx[i], ly[i], lt[i], lp[i], Mat4::from(slice_poses + 16 * (size_t)slice_of[i]), rx, ry, rt, rp, nr, em)) vEMP.push_back(em);
      }
  }
};

// esvo_MVStereo::vEMP2vDP (esvo_MVStereo.cpp:1072-1097)
static inline void vEMP2vDP(const CameraSystem& cs, const std::vector<Seed>& vEMP, double age_vis_threshold, std::vector<DepthPoint>& vdp) {
  vdp.clear(); vdp.reserve(vEMP.size());
  for (const Seed& e : vEMP) {
    DepthPoint dp((int64_t)std::floor(e.x_left[1]), (int64_t)std::floor(e.x_left[0]));   // DepthPoint(row, col)
    dp.x[0] = e.x_left[0]; dp.x[1] = e.x_left[1];                                         // update_x
    cs.left.cam2World(e.x_left, e.invDepth, dp.p_cam);                                    // update_p_cam
    dp.update(e.invDepth, 0.0);                                                           // var_pseudo = 0 -> boundVariance 1e-6
    dp.residual = e.cost;
    dp.age = (int64_t)age_vis_threshold;
    dp.T_world_cam = e.trans;                                                             // updatePose
    vdp.push_back(dp);
  }
}

}  // namespace oracle
