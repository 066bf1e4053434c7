// esvo_b200 -- C++ host shim over the C ABI (include/esvo_b200.h).
//
// ROS-free mirror of the reference's class surface for the hot path, so that code written against
// esvo_core::core::{EventBM, DepthProblemSolver, DepthFusion, RegProblemSolverLM} and
// esvo_time_surface::TimeSurface keeps its shape: same class and method names, same argument meaning,
// same "bool return for data-dependent failure, no exceptions" convention
// (reference: esvo_core/include/esvo_core/core/*.h, esvo_time_surface/include/esvo_time_surface/TimeSurface.h).
// ROS / Eigen / OpenCV types are replaced by plain structs: ros::Time -> int64 ns, Transformation -> row-major
// double[16], cv::Mat mono8 -> uint8_t*, dvs_msgs::Event -> esvo::Event.  Header-only; link with libesvo_b200.so.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../esvo_b200.h"

namespace esvo {

struct Event { uint16_t x, y; int64_t ts; bool polarity; };            // dvs_msgs::Event
using Pose = std::array<double, 16>;                                    // Transformation (T_world_cam, row-major)
using StampTransformationMap = std::vector<std::pair<int64_t, Pose>>;   // sorted by stamp
using EventMatchPair = esvo_seed;                                       // core::EventMatchPair
using DepthPoint = esvo_depth_point;                                    // container::DepthPoint

// container::CameraSystem + the parameter yaml of one node: owns the esvo_ctx of one event stream / GPU.
class CameraSystem {
 public:
  using Ptr = std::shared_ptr<CameraSystem>;
  CameraSystem(const esvo_calib& left, const esvo_calib& right, const esvo_params& params, int device = 0) {
    int st = 0;
    ctx_ = esvo_create(device, &left, &right, &params, &st);
    if (!ctx_) throw std::runtime_error("esvo_create failed with status " + std::to_string(st) + " (no CPU fallback)");
    params_ = params; width_ = left.width; height_ = left.height;
    std::copy(left.P, left.P + 12, P_left_);
    double d[4]; esvo_get_derived(ctx_, d); baseline_ = d[0];
  }
  ~CameraSystem() { esvo_destroy(ctx_); }
  CameraSystem(const CameraSystem&) = delete;
  CameraSystem& operator=(const CameraSystem&) = delete;
  esvo_ctx* ctx() const { return ctx_; }
  const esvo_params& params() const { return params_; }
  int width_ = 0, height_ = 0;
  double baseline_ = 0;
  double P_left_[12] = {0};   // cam_left_ptr_->P_
 private:
  esvo_ctx* ctx_ = nullptr;
  esvo_params params_;
};

// container::TimeSurfaceObservation (stereo mono8 pair + pose); images are borrowed, not copied.
struct TimeSurfaceObservation {
  const uint8_t* left = nullptr;    // nullptr = "the image esvo_ts_build produced on the device"
  const uint8_t* right = nullptr;
  Pose tr_{};
  size_t id_ = 0;
  void setTransformation(const Pose& tr) { tr_ = tr; }
};
using StampedTimeSurfaceObs = std::pair<int64_t, TimeSurfaceObservation>;

}  // namespace esvo

namespace esvo_time_surface {
// esvo_time_surface::TimeSurface (TimeSurface.h:98-170) without the ROS node handle: the three callbacks
// become plain methods, the published image is returned to the caller.
class TimeSurface {
 public:
  TimeSurface(esvo::CameraSystem::Ptr cs, int cam) : cs_(std::move(cs)), cam_(cam) {}
  // eventsCallback (TimeSurface.cpp:403-425)
  bool eventsCallback(const std::vector<esvo::Event>& events) {
    x_.resize(events.size()); y_.resize(events.size()); t_.resize(events.size()); p_.resize(events.size());
    for (size_t i = 0; i < events.size(); ++i) { x_[i] = events[i].x; y_[i] = events[i].y; t_[i] = events[i].ts; p_[i] = events[i].polarity; }
    return esvo_ts_push_events(cs_->ctx(), cam_, x_.data(), y_.data(), t_.data(), p_.data(), events.size()) == ESVO_OK;
  }
  // syncCallback -> createTimeSurfaceAtTime (TimeSurface.cpp:52-152,293-311); time_surface_out: H*W mono8 or nullptr
  bool createTimeSurfaceAtTime(int64_t external_sync_time_ns, uint8_t* time_surface_out) {
    return esvo_ts_build(cs_->ctx(), cam_, external_sync_time_ns, nullptr, time_surface_out) == ESVO_OK;
  }
  void clearEventQueue() { esvo_ts_reset(cs_->ctx(), cam_); }
  // The reference tolerates stamps that arrive out of order (its insertion sort then re-queues events_.back(),
  // TimeSurface.cpp:410-422).  Off by default here (drivers deliver ordered packets; saves two kernels per push);
  // switch it on for sources that do not guarantee ordering -- results are then identical to the reference's.
  bool setUnorderedInput(bool enable) { return esvo_ts_set_unordered_input(cs_->ctx(), cam_, enable ? 1 : 0) == ESVO_OK; }
 private:
  esvo::CameraSystem::Ptr cs_;
  int cam_;
  std::vector<uint16_t> x_, y_;
  std::vector<int64_t> t_;
  std::vector<uint8_t> p_;
};
}  // namespace esvo_time_surface

namespace esvo_core {
namespace core {
using esvo::CameraSystem;
using esvo::DepthPoint;
using esvo::EventMatchPair;

// core::EventBM (EventBM.h:27-60).  Patch size / disparity range / step / threshold come from the params the
// CameraSystem was created with (the reference passes them to resetParameters from the same yaml).
class EventBM {
 public:
  explicit EventBM(CameraSystem::Ptr camSysPtr) : camSysPtr_(std::move(camSysPtr)) {}
  void createMatchProblem(esvo::StampedTimeSurfaceObs* pStampedTsObs, esvo::StampTransformationMap* pSt_map,
                          std::vector<esvo::Event*>* pvEventsPtr) {
    pStampedTsObs_ = pStampedTsObs; pSt_map_ = pSt_map;
    ex_.clear(); ey_.clear(); et_.clear();
    for (auto* e : *pvEventsPtr) { ex_.push_back(e->x); ey_.push_back(e->y); et_.push_back(e->ts); }
  }
  void match_all_HyperThread(std::vector<EventMatchPair>& vEMP) {
    vEMP.clear();
    if (!pStampedTsObs_ || !pSt_map_) return;
    const auto& obs = pStampedTsObs_->second;
    if (esvo_set_ts_pair(camSysPtr_->ctx(), obs.left, obs.right, obs.tr_.data()) != ESVO_OK) return;
    std::vector<int64_t> pt; std::vector<double> poses;
    for (auto& st : *pSt_map_) { pt.push_back(st.first); poses.insert(poses.end(), st.second.begin(), st.second.end()); }
    vEMP.resize(ex_.size());
    size_t n = vEMP.size();
    if (esvo_bm_match(camSysPtr_->ctx(), ex_.data(), ey_.data(), et_.data(), ex_.size(), pt.data(), poses.data(), pt.size(),
                      vEMP.data(), &n, &n_evals_) != ESVO_OK) n = 0;
    vEMP.resize(n);
  }
  void match_all_SingleThread(std::vector<EventMatchPair>& vEMP) { match_all_HyperThread(vEMP); }
  uint64_t n_evals_ = 0;
 private:
  CameraSystem::Ptr camSysPtr_;
  esvo::StampedTimeSurfaceObs* pStampedTsObs_ = nullptr;
  esvo::StampTransformationMap* pSt_map_ = nullptr;
  std::vector<uint16_t> ex_, ey_;
  std::vector<int64_t> et_;
};

enum DepthProblemType { ANALYTICAL, NUMERICAL };
// core::DepthProblemSolver (DepthProblemSolver.h:35-60)
class DepthProblemSolver {
 public:
  DepthProblemSolver(CameraSystem::Ptr& camSysPtr, DepthProblemType dpType = NUMERICAL) : camSysPtr_(camSysPtr), dpType_(dpType) {}
  void solve(std::vector<EventMatchPair>* pvEMP, esvo::StampedTimeSurfaceObs* pStampedTsObs, std::vector<DepthPoint>& vdp) {
    vdp.clear();
    if (dpType_ != NUMERICAL || !pvEMP || pvEMP->empty()) return;   // the reference exit(-1)s on ANALYTICAL
    const auto& obs = pStampedTsObs->second;
    if (esvo_set_ts_pair(camSysPtr_->ctx(), obs.left, obs.right, obs.tr_.data()) != ESVO_OK) return;
    vdp.resize(pvEMP->size());
    size_t n = vdp.size();
    if (esvo_depth_solve(camSysPtr_->ctx(), pvEMP->data(), pvEMP->size(), vdp.data(), &n, &n_evals_) != ESVO_OK) n = 0;
    vdp.resize(n);
  }
  void pointCulling(std::vector<DepthPoint>& vdp, double std_variance_threshold, double cost_threshold,
                    double invDepth_min_range, double invDepth_max_range) {
    size_t n = vdp.size();
    if (esvo_depth_cull(camSysPtr_->ctx(), vdp.data(), &n, std_variance_threshold, cost_threshold, invDepth_min_range,
                        invDepth_max_range) == ESVO_OK)
      vdp.resize(n);
  }
  DepthProblemType getProblemType() { return dpType_; }
  uint64_t n_evals_ = 0;
 private:
  CameraSystem::Ptr camSysPtr_;
  DepthProblemType dpType_;
};

// container::DepthFrame: the fused map lives on the device inside the ctx; this handle carries its pose.
struct DepthFrame {
  using Ptr = std::shared_ptr<DepthFrame>;
  esvo::Pose T_world_frame_{};
  size_t id_ = 0;
  bool fresh_ = true;   // a newly constructed DepthFrame starts from an empty map (esvo_Mapping.cpp:268-272)
  void setTransformation(const esvo::Pose& T) { T_world_frame_ = T; }
  void setId(size_t id) { id_ = id; }
};

// core::DepthFusion (DepthFusion.h:21-59)
class DepthFusion {
 public:
  explicit DepthFusion(CameraSystem::Ptr& camSysPtr) : camSysPtr_(camSysPtr) {}
  int update(std::vector<DepthPoint>& dp_obs, DepthFrame::Ptr& df, int fusion_radius) {
    int nf = 0;
    if (esvo_fuse(camSysPtr_->ctx(), dp_obs.data(), dp_obs.size(), df->T_world_frame_.data(), fusion_radius, df->fresh_ ? 1 : 0,
                  &nf) != ESVO_OK)
      return 0;
    df->fresh_ = false;
    return nf;
  }
  // naive_propagation (DepthFusion.cpp:232-288): nearest wins, no fusion -- the accumulation of the MVStereo comparison modes
  void naive_propagation(std::vector<DepthPoint>& dp_obs, DepthFrame::Ptr& df) {
    if (esvo_naive_propagate(camSysPtr_->ctx(), dp_obs.data(), dp_obs.size(), df->T_world_frame_.data(), df->fresh_ ? 1 : 0) == ESVO_OK)
      df->fresh_ = false;
  }
  // DepthMap::clean (SmartGrid.h:222-243) and the element list, exposed here because the map is ctx-resident
  void clean(double var_threshold, double age_threshold, double range_max, double range_min) {
    esvo_map_clean(camSysPtr_->ctx(), var_threshold, age_threshold, range_max, range_min);
  }
  void getElements(std::vector<DepthPoint>& out) {
    out.resize((size_t)camSysPtr_->width_ * camSysPtr_->height_);
    size_t n = out.size();
    if (esvo_map_download(camSysPtr_->ctx(), out.data(), &n) != ESVO_OK) n = 0;
    out.resize(n);
  }
 private:
  CameraSystem::Ptr camSysPtr_;
};

// core::DepthRegularization (DepthRegularization.h:20)
class DepthRegularization {
 public:
  explicit DepthRegularization(CameraSystem::Ptr& camSysPtr) : camSysPtr_(camSysPtr) {}
  void apply() { esvo_map_regularize(camSysPtr_->ctx()); }
 private:
  CameraSystem::Ptr camSysPtr_;
};

// core::EventSlice (EventMatcher.h:17-28): a run of left events that shares one virtual-view pose
struct EventSlice {
  explicit EventSlice(double SLICE_THICKNESS = 2 * 1e-3) : SLICE_THICKNESS_(SLICE_THICKNESS) {}
  size_t numEvents_ = 0;
  double SLICE_THICKNESS_;
  int64_t t_median_ = 0;
  esvo::Pose transf_{};
  std::vector<esvo::Event*>::iterator it_begin_, it_end_;
};

// core::EventMatcher (EventMatcher.h:30-100): the event-to-event matcher of [26] the reference keeps for comparison.
class EventMatcher {
 public:
  EventMatcher(CameraSystem::Ptr camSysPtr, size_t numThread = 1, double Time_THRESHOLD = 10e-5, double EPIPOLAR_THRESHOLD = 0.5,
               double TS_NCC_THRESHOLD = 0.1, size_t patch_size_X = 25, size_t patch_size_Y = 5, size_t patch_intensity_threshold = 125,
               double patch_valid_ratio = 0.1)
      : camSysPtr_(std::move(camSysPtr)), NUM_THREAD_(numThread) {
    resetParameters(Time_THRESHOLD, EPIPOLAR_THRESHOLD, TS_NCC_THRESHOLD, patch_size_X, patch_size_Y, patch_intensity_threshold, patch_valid_ratio);
  }
  void resetParameters(double Time_THRESHOLD, double EPIPOLAR_THRESHOLD, double TS_NCC_THRESHOLD, size_t patch_size_X, size_t patch_size_Y,
                       size_t patch_intensity_threshold, double patch_valid_ratio) {
    prm_.time_threshold_s = Time_THRESHOLD; prm_.epipolar_threshold = EPIPOLAR_THRESHOLD; prm_.ts_ncc_threshold = TS_NCC_THRESHOLD;
    prm_.patch_size_x = (int32_t)patch_size_X; prm_.patch_size_y = (int32_t)patch_size_Y; prm_.num_thread = (int32_t)NUM_THREAD_; prm_._pad = 0;
    patch_intensity_threshold_ = patch_intensity_threshold; patch_valid_ratio_ = patch_valid_ratio;   // unused by the reference as well
  }
  void createMatchProblem(esvo::StampedTimeSurfaceObs* pTS_obs, std::vector<EventSlice>* vEventSlice_ptr, std::vector<esvo::Event*>* vEventPtr_cand) {
    pTS_obs_ = pTS_obs; pvEventSlice_ = vEventSlice_ptr; pvCandEventPtr_ = vEventPtr_cand;
  }
  void match_all_HyperThread(std::vector<EventMatchPair>& vEMP) {
    vEMP.clear();
    if (!pTS_obs_ || !pvEventSlice_ || !pvCandEventPtr_ || pvEventSlice_->empty()) return;
    const auto& obs = pTS_obs_->second;
    if (esvo_set_ts_pair(camSysPtr_->ctx(), obs.left, obs.right, obs.tr_.data()) != ESVO_OK) return;
    // the events of all slices are contiguous from the first slice's begin (match(), EventMatcher.cpp:233-251)
    std::vector<int32_t> counts; std::vector<double> poses; size_t total = 0;
    for (auto& es : *pvEventSlice_) { counts.push_back((int32_t)es.numEvents_); poses.insert(poses.end(), es.transf_.begin(), es.transf_.end()); total += es.numEvents_; }
    std::vector<uint16_t> lx, ly, rx, ry; std::vector<int64_t> lt, rt; std::vector<uint8_t> lp, rp;
    auto it = (*pvEventSlice_)[0].it_begin_;
    for (size_t i = 0; i < total; ++i, ++it) { lx.push_back((*it)->x); ly.push_back((*it)->y); lt.push_back((*it)->ts); lp.push_back((*it)->polarity); }
    for (auto* e : *pvCandEventPtr_) { rx.push_back(e->x); ry.push_back(e->y); rt.push_back(e->ts); rp.push_back(e->polarity); }
    vEMP.resize(total);
    size_t n = total;
    if (esvo_em_match(camSysPtr_->ctx(), &prm_, lx.data(), ly.data(), lt.data(), lp.data(), total, counts.data(), poses.data(), counts.size(),
                      rx.data(), ry.data(), rt.data(), rp.data(), rx.size(), vEMP.data(), &n, &n_evals_) != ESVO_OK)
      n = 0;
    vEMP.resize(n);
  }
  // plain event order.  (The reference's single-thread walk additionally stops one event short of every slice's end,
  // EventMatcher.cpp:165-183; esvo_MVStereo only calls the HyperThread form, so that quirk is not reproduced.)
  void match_all_SingleThread(std::vector<EventMatchPair>& vEMP) {
    const int32_t nt = prm_.num_thread; prm_.num_thread = 1; match_all_HyperThread(vEMP); prm_.num_thread = nt;
  }
  uint64_t n_evals_ = 0;
 private:
  CameraSystem::Ptr camSysPtr_;
  size_t NUM_THREAD_;
  esvo_em_params prm_{};
  size_t patch_intensity_threshold_ = 125; double patch_valid_ratio_ = 0.1;
  esvo::StampedTimeSurfaceObs* pTS_obs_ = nullptr;
  std::vector<EventSlice>* pvEventSlice_ = nullptr;
  std::vector<esvo::Event*>* pvCandEventPtr_ = nullptr;
};

// core::RefFrame / CurFrame (RegProblemLM.h:58-74)
struct RefFrame { int64_t t_ = 0; std::vector<float> vPointXYZ_; /* n*3, world frame; permuted in place */ esvo::Pose tr_{}; };
struct CurFrame { int64_t t_ = 0; const uint8_t* ts_left = nullptr; esvo::Pose tr_{}; size_t numEventsSinceLastObs_ = 0; };
struct LM_statics { size_t nPoints_ = 0, nfev_ = 0, nIter_ = 0; };
enum RegProblemType { REG_NUMERICAL, REG_ANALYTICAL };

// core::RegProblemSolverLM (RegProblemSolverLM.h:37-50)
class RegProblemSolverLM {
 public:
  RegProblemSolverLM(CameraSystem::Ptr& camSysPtr, RegProblemType rpType = REG_ANALYTICAL) : camSysPtr_(camSysPtr), rpType_(rpType) {}
  bool resetRegProblem(RefFrame* ref, CurFrame* cur) {
    cur_ = cur;
    int rc = esvo_track_reset(camSysPtr_->ctx(), ref->vPointXYZ_.data(), ref->vPointXYZ_.size() / 3, ref->tr_.data(),
                              cur->tr_.data(), cur->ts_left);
    lmStatics_ = LM_statics();
    return rc == ESVO_OK;   // 1 = not enough points in the local map -> the system re-initialises
  }
  bool solve_numerical() { return solve(0); }
  bool solve_analytical() { return solve(1); }
  LM_statics lmStatics_;
 private:
  bool solve(int analytical) {
    esvo_lm_stats st{};
    if (!cur_ || esvo_track_solve(camSysPtr_->ctx(), analytical, cur_->tr_.data(), &st) != ESVO_OK) return false;
    lmStatics_.nPoints_ = (size_t)st.n_points; lmStatics_.nfev_ = (size_t)st.nfev; lmStatics_.nIter_ = (size_t)st.n_iter;
    return true;
  }
  CameraSystem::Ptr camSysPtr_;
  RegProblemType rpType_;
  CurFrame* cur_ = nullptr;
};

}  // namespace core

// The event front-end of esvo_Mapping::dataTransferring (esvo_Mapping.cpp:536-603), host logic feeding the hot path:
//  * selectCloseEvents: the newest-first, at most PROCESS_EVENT_NUM left events with stamps in
//    [t_end - 10*BM_half_slice_thickness, t_end)  (:562-575; events_left_ is time-ordered);
//  * samplePoseStamps: the virtual-view stamps t_begin, t_begin + 0.05*BM_half_slice_thickness, ... <= t_end at
//    which the node looks up tf poses to fill st_map_ (:585-599).
// Both use the reference's ros::Time <-> double conversions (toSec() compares, Time(double) construction).
namespace frontend {
inline double toSec(int64_t ns) { int64_t s = ns / 1000000000LL, n = ns % 1000000000LL; if (n < 0) { n += 1000000000LL; --s; } return (double)s + 1e-9 * (double)n; }
inline int64_t fromSec(double t) {   // ros::Time(double): sec = floor(t), nsec = round((t - sec) * 1e9), carry on overflow
  int64_t sec = (int64_t)std::floor(t);
  int64_t nsec = (int64_t)std::llround((t - (double)sec) * 1e9);
  sec += nsec / 1000000000LL; nsec %= 1000000000LL;
  return sec * 1000000000LL + nsec;
}
inline void selectCloseEvents(std::vector<esvo::Event>& events_left /* time-ordered */, int64_t t_end_ns, double BM_half_slice_thickness,
                              size_t PROCESS_EVENT_NUM, std::vector<esvo::Event*>& vCloseEventsPtr_left) {
  vCloseEventsPtr_left.clear();
  if (events_left.empty()) return;
  const double t_end = toSec(t_end_ns);
  const int64_t t_begin_ns = fromSec(std::max(0.0, t_end - 10 * BM_half_slice_thickness));
  auto lower = [&](int64_t t) {   // tools::EventBuffer_lower_bound (utils.h:50-55): compares toSec() DOUBLES (stamps that share a double are equal)
    const double ts = toSec(t);
    size_t lo = 0, hi = events_left.size();
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (toSec(events_left[mid].ts) < ts) lo = mid + 1; else hi = mid; }
    return lo;
  };
  size_t ev_end = lower(t_end_ns);
  const size_t ev_begin = lower(t_begin_ns);
  // The walk starts AT lower_bound(t_end) and stops before lower_bound(t_begin) (:570-574).  When no event is at/after
  // t_end (the observation stamp is newer than the whole buffer -- the common case online) the reference's first push is
  // deque::end()._M_cur, one slot past the newest event (UB: an unconstructed event); it then still walks back over the
  // newest valid events.  We skip only that slot but keep its place in the PROCESS_EVENT_NUM budget.
  size_t budget = PROCESS_EVENT_NUM;
  if (ev_end == events_left.size() && ev_end != ev_begin && budget > 0) { --ev_end; --budget; }
  while (ev_end != ev_begin && vCloseEventsPtr_left.size() < budget) {
    vCloseEventsPtr_left.push_back(&events_left[ev_end]);
    --ev_end;
  }
}
//  * selectSGMEvents: the INITIALIZATION branch of dataTransferring (:538-552) -- window 2*BM_half_slice_thickness and
//    `<=` instead of `<` on the count (one event more than PROCESS_EVENT_NUM).
inline void selectSGMEvents(std::vector<esvo::Event>& events_left, int64_t t_end_ns, double BM_half_slice_thickness, size_t PROCESS_EVENT_NUM,
                            std::vector<esvo::Event*>& vEventsPtr_left_SGM) {
  vEventsPtr_left_SGM.clear();
  if (events_left.empty()) return;
  const int64_t t_begin_ns = fromSec(std::max(0.0, toSec(t_end_ns) - 2 * BM_half_slice_thickness));
  auto lower = [&](int64_t t) {
    const double ts = toSec(t);
    size_t lo = 0, hi = events_left.size();
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (toSec(events_left[mid].ts) < ts) lo = mid + 1; else hi = mid; }
    return lo;
  };
  size_t ev_end = lower(t_end_ns);
  const size_t ev_begin = lower(t_begin_ns);
  size_t budget = PROCESS_EVENT_NUM + 1;      // `<=` on the count: one event more than PROCESS_EVENT_NUM
  if (ev_end == events_left.size() && ev_end != ev_begin) { --ev_end; --budget; }   // skip the one-past-the-end slot (see selectCloseEvents)
  while (ev_end != ev_begin && vEventsPtr_left_SGM.size() < budget) { vEventsPtr_left_SGM.push_back(&events_left[ev_end]); --ev_end; }
}
inline std::vector<int64_t> samplePoseStamps(int64_t t_end_ns, double BM_half_slice_thickness) {
  std::vector<int64_t> out;
  const double t_end = toSec(t_end_ns);
  int64_t t_tmp = fromSec(std::max(0.0, t_end - 10 * BM_half_slice_thickness));
  while (toSec(t_tmp) <= t_end) { out.push_back(t_tmp); t_tmp = fromSec(toSec(t_tmp) + 0.05 * BM_half_slice_thickness); }
  return out;
}
//  * createDenoisingMask / extractDenoisedEvents (esvo_Mapping.cpp:1046-1072, bDenoising_): binary event map of ALL
//    window events -> cv::medianBlur 3x3 (on a 0/255 image: a pixel survives iff at least 5 of its 3x3 neighbours,
//    borders replicated, carry an event) -> keep the close events that sit on surviving pixels, at most maxNum.
inline void createDenoisingMask(const std::vector<esvo::Event*>& vAllEventsPtr, std::vector<uint8_t>& mask, size_t row, size_t col) {
  std::vector<uint8_t> eventMap(row * col, 0);                       // Visualization::plot_eventMap (Visualization.cpp:96-104)
  for (const esvo::Event* e : vAllEventsPtr) if (e->x < col && e->y < row) eventMap[(size_t)e->y * col + e->x] = 255;
  mask.assign(row * col, 0);
  for (size_t y = 0; y < row; ++y)
    for (size_t x = 0; x < col; ++x) {
      int on = 0;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const size_t yy = (size_t)std::min<long>(std::max<long>((long)y + dy, 0), (long)row - 1);
          const size_t xx = (size_t)std::min<long>(std::max<long>((long)x + dx, 0), (long)col - 1);
          on += eventMap[yy * col + xx] != 0;
        }
      mask[y * col + x] = on >= 5 ? 255 : 0;
    }
}
inline void extractDenoisedEvents(const std::vector<esvo::Event*>& vCloseEventsPtr, std::vector<esvo::Event*>& vEdgeEventsPtr,
                                  const std::vector<uint8_t>& mask, size_t col, size_t maxNum) {
  vEdgeEventsPtr.reserve(vCloseEventsPtr.size());
  for (esvo::Event* e : vCloseEventsPtr) {
    if (vEdgeEventsPtr.size() >= maxNum) break;
    if (mask[(size_t)e->y * col + e->x] == 255) vEdgeEventsPtr.push_back(e);
  }
}
//  * the local-map hand-off of publishPointCloud (esvo_Mapping.cpp:909-953): p_world = R p_cam + t as pcl::PointXYZ
//    (f32 x,y,z) in DepthMap iteration order; `near` additionally keeps ||p_cam|| < visualize_range (:930-931).
//    The result is what RefFrame::vPointXYZ_ / esvo_track_reset consume on the tracking side (esvo_Tracking.cpp:202-234).
inline void packPointCloud(const std::vector<esvo::DepthPoint>& elems, const esvo::Pose& T_world_result, std::vector<float>& xyz,
                           std::vector<float>* xyz_near = nullptr, double visualize_range = 0.0) {
  xyz.clear(); xyz.reserve(3 * elems.size());
  if (xyz_near) xyz_near->clear();
  const double* T = T_world_result.data();
  for (const esvo::DepthPoint& d : elems) {
    const double* p = d.p_cam;
    const double w[3] = {T[0] * p[0] + T[1] * p[1] + T[2] * p[2] + T[3], T[4] * p[0] + T[5] * p[1] + T[6] * p[2] + T[7],
                         T[8] * p[0] + T[9] * p[1] + T[10] * p[2] + T[11]};
    for (int k = 0; k < 3; ++k) xyz.push_back((float)w[k]);
    if (xyz_near && std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) < visualize_range)
      for (int k = 0; k < 3; ++k) xyz_near->push_back((float)w[k]);
  }
}
//  * eventSlicingForEM (esvo_MVStereo.cpp:1008-1040): cuts the time-ordered left events of [t_lowBound, t_upBound) into
//    floor((t_up - t_low) / EM_Slice_Thickness) slices; a slice runs from its first event to lower_bound(first stamp +
//    thickness) INCLUSIVE, the next one starts behind it; its pose is the trajectory at the median event's stamp.
template <class PoseAt>
inline void eventSlicingForEM(std::vector<esvo::Event*>& vEventsPtr_left, int64_t t_lowBound_ns, int64_t t_upBound_ns, double EM_Slice_Thickness,
                              PoseAt&& getPoseAt, std::vector<core::EventSlice>& eventSlices) {
  eventSlices.clear();
  if (vEventsPtr_left.empty()) return;
  const size_t numSlice = (size_t)std::floor((toSec(t_upBound_ns) - toSec(t_lowBound_ns)) / EM_Slice_Thickness);
  auto it_tmp = vEventsPtr_left.begin();
  for (size_t i = 0; i < numSlice; ++i) {
    core::EventSlice es(EM_Slice_Thickness);
    es.it_begin_ = it_tmp;
    const double t_end = toSec(fromSec(toSec((*it_tmp)->ts) + es.SLICE_THICKNESS_));
    es.it_end_ = std::lower_bound(vEventsPtr_left.begin(), vEventsPtr_left.end(), t_end,
                                  [](const esvo::Event* e, double t) { return toSec(e->ts) < t; });   // tools::EventVecPtr_lower_bound (utils.h:43-48)
    if (es.it_end_ == vEventsPtr_left.end()) --es.it_end_;
    es.numEvents_ = (size_t)std::distance(es.it_begin_, es.it_end_) + 1;
    auto it_median = es.it_begin_;
    std::advance(it_median, es.numEvents_ / 2);
    es.t_median_ = (*it_median)->ts;
    getPoseAt(es.t_median_, es.transf_);
    eventSlices.push_back(es);
    it_tmp = es.it_end_;
    ++it_tmp;
    if (it_tmp == vEventsPtr_left.end()) break;
  }
}
}  // namespace frontend

// esvo_core::esvo_Mapping::MappingAtTime (esvo_Mapping.cpp:261-399) as one call with device-resident hand-off.
class esvo_Mapping {
 public:
  explicit esvo_Mapping(esvo::CameraSystem::Ptr cs) : cs_(std::move(cs)) {}
  struct Counters { uint64_t n_events, n_seeds, n_solved, n_culled, n_fusions, bm_evals, lm_evals, map_size; };
  bool MappingAtTime(const esvo::StampedTimeSurfaceObs& TS_obs, const std::vector<esvo::Event*>& vCloseEventsPtr_left,
                     const esvo::StampTransformationMap& st_map, Counters* out = nullptr) {
    const auto& obs = TS_obs.second;
    if (esvo_set_ts_pair(cs_->ctx(), obs.left, obs.right, obs.tr_.data()) != ESVO_OK) return false;
    std::vector<uint16_t> ex, ey; std::vector<int64_t> et, pt; std::vector<double> poses;
    for (auto* e : vCloseEventsPtr_left) { ex.push_back(e->x); ey.push_back(e->y); et.push_back(e->ts); }
    for (auto& st : st_map) { pt.push_back(st.first); poses.insert(poses.end(), st.second.begin(), st.second.end()); }
    uint64_t c[8];
    if (esvo_mapping_at_time(cs_->ctx(), ex.data(), ey.data(), et.data(), ex.size(), pt.data(), poses.data(), pt.size(), c) != ESVO_OK)
      return false;
    if (out) *out = Counters{c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]};
    return true;
  }
  // InitializationAtTime (esvo_Mapping.cpp:433-492).  The SGM stays where it is -- the node's
  //   sgbm_->compute(TS_left, TS_right, dispMap)       (cv::StereoSGBM(0,48,11,968,3872,-1,0,11), esvo_Mapping.cpp:101-108,444)
  // -- and its CV_16S result is passed in; everything downstream (edge mask AND disparity, DepthPoint creation,
  // naive_propagation into the new DepthFrame, first window vector) runs on the device.  Returns false like the
  // reference when fewer than INIT_SGM_DP_NUM_Threshold points survive.
  bool InitializationAtTime(const esvo::StampedTimeSurfaceObs& TS_obs, const int16_t* dispMap16,
                            const std::vector<esvo::Event*>& vEventsPtr_left_SGM, size_t INIT_SGM_DP_NUM_Threshold, size_t* n_points = nullptr) {
    std::vector<uint16_t> ex, ey;
    for (auto* e : vEventsPtr_left_SGM) { ex.push_back(e->x); ey.push_back(e->y); }
    size_t n = 0; int accepted = 0;
    if (esvo_init_from_disparity(cs_->ctx(), dispMap16, ex.data(), ey.data(), ex.size(), TS_obs.second.tr_.data(), INIT_SGM_DP_NUM_Threshold,
                                 &n, &accepted) != ESVO_OK)
      return false;
    if (n_points) *n_points = n;
    return accepted != 0;
  }
  // Same, with the SGM itself on the device as well (esvo_sgbm_compute: bit-exact restatement of cv::StereoSGBM MODE_SGBM with
  // the reference's parameters, esvo_Mapping.cpp:101-108) -- no OpenCV needed on the node side.
  bool InitializationAtTime(const esvo::StampedTimeSurfaceObs& TS_obs, const std::vector<esvo::Event*>& vEventsPtr_left_SGM,
                            size_t INIT_SGM_DP_NUM_Threshold, size_t* n_points = nullptr) {
    const int num_disparities = 16 * 3, block_size = 11;                       // esvo_Mapping.cpp:101-105
    std::vector<int16_t> disp((size_t)cs_->width_ * cs_->height_);
    if (esvo_sgbm_compute(cs_->ctx(), TS_obs.second.left, TS_obs.second.right, num_disparities, block_size, 8 * block_size * block_size,
                          32 * block_size * block_size, -1, 0, 11, disp.data()) != ESVO_OK)
      return false;
    return InitializationAtTime(TS_obs, disp.data(), vEventsPtr_left_SGM, INIT_SGM_DP_NUM_Threshold, n_points);
  }
  void reset() { esvo_mapping_reset(cs_->ctx()); }
 private:
  esvo::CameraSystem::Ptr cs_;
};

// esvo_core::esvo_MVStereo::MappingAtTime (esvo_MVStereo.cpp:239-520): the known-pose multi-view-stereo node the reference
// uses to compare five depth estimators on the same fusion back-end.  The node's dataTransferring (:544-668) fills the
// public input members below (front-end helpers: frontend::selectCloseEvents / samplePoseStamps / eventSlicingForEM);
// MappingAtTime then runs the selected mode on the device through the same ABI entry points the mapper uses.
class esvo_MVStereo {
 public:
  enum eMVStereoMode { PURE_EVENT_MATCHING, PURE_BLOCK_MATCHING, EM_PLUS_ESTIMATION, BM_PLUS_ESTIMATION, PURE_SEMI_GLOBAL_MATCHING };   // esvo_MVStereo.h:38-45
  using PoseProvider = std::function<bool(int64_t, esvo::Pose&)>;
  esvo_MVStereo(esvo::CameraSystem::Ptr cs, eMVStereoMode msm, size_t NUM_THREAD_MAPPING = 4)
      : msm_(msm), em_(cs, NUM_THREAD_MAPPING), ebm_(cs), dpSolver_(cs), dFusor_(cs), dRegularizor_(cs), cs_(cs) {
    const esvo_params& p = cs_->params();
    em_.resetParameters(EM_Time_THRESHOLD_, EM_EPIPOLAR_THRESHOLD_, EM_TS_NCC_THRESHOLD_, (size_t)p.patch_size_x, (size_t)p.patch_size_y, 125, 0.1);   // :85-91
    maxNumFusionFrames_ = (size_t)p.max_num_fusion_frames; maxNumFusionPoints_ = (size_t)p.max_num_fusion_points;
  }
  void setPoseProvider(PoseProvider f) { getPoseAt_ = std::move(f); }
  void resetEMParameters(double thickness, double time_thr, double epi_thr, double ncc_thr) {
    EM_Slice_Thickness_ = thickness; EM_Time_THRESHOLD_ = time_thr; EM_EPIPOLAR_THRESHOLD_ = epi_thr; EM_TS_NCC_THRESHOLD_ = ncc_thr;
    const esvo_params& p = cs_->params();
    em_.resetParameters(time_thr, epi_thr, ncc_thr, (size_t)p.patch_size_x, (size_t)p.patch_size_y, 125, 0.1);
  }
  // ---- inputs, as dataTransferring leaves them (:544-668) ----
  esvo::StampedTimeSurfaceObs TS_obs_;
  std::vector<esvo::Event*> vEventsPtr_left_, vEventsPtr_right_;        // EM: all events of [t_lowBound_, t_upBound_)
  int64_t t_lowBound_ = 0, t_upBound_ = 0;
  std::vector<esvo::Event*> vCloseEventsPtr_left_;                       // BM: newest-first window events (already denoised / truncated)
  esvo::StampTransformationMap st_map_;
  std::vector<esvo::Event*> vEventsPtr_left_SGM_;                        // SGM: events that draw the edge mask
  // ---- results ----
  std::vector<esvo::EventMatchPair> vEMP_;
  std::deque<std::vector<esvo::DepthPoint>> dqvDepthPoints_;
  size_t numFusionCount_ = 0;

  bool MappingAtTime() {
    const esvo_params& p = cs_->params();
    core::DepthFrame::Ptr depthFramePtr_new = std::make_shared<core::DepthFrame>();             // :246-251
    depthFramePtr_new->setId(TS_obs_.second.id_);
    depthFramePtr_new->setTransformation(TS_obs_.second.tr_);
    depthFramePtr_ = depthFramePtr_new;
    vEMP_.clear();
    if (msm_ == PURE_EVENT_MATCHING || msm_ == EM_PLUS_ESTIMATION) {                            // :257-306
      std::vector<core::EventSlice> eventSlices;
      frontend::eventSlicingForEM(vEventsPtr_left_, t_lowBound_, t_upBound_, EM_Slice_Thickness_,
                                  [&](int64_t t, esvo::Pose& T) { return getPoseAt_ ? getPoseAt_(t, T) : false; }, eventSlices);
      em_.createMatchProblem(&TS_obs_, &eventSlices, &vEventsPtr_right_);
      em_.match_all_HyperThread(vEMP_);
      if (vEMP_.empty()) return false;
      if (msm_ == PURE_EVENT_MATCHING) { std::vector<esvo::DepthPoint> vdp; vEMP2vDP(vEMP_, vdp); return accumulateNaive(vdp); }
    }
    if (msm_ == PURE_SEMI_GLOBAL_MATCHING) {                                                    // :311-378
      std::vector<int16_t> dispMap((size_t)cs_->width_ * cs_->height_);
      if (esvo_set_ts_pair(cs_->ctx(), TS_obs_.second.left, TS_obs_.second.right, TS_obs_.second.tr_.data()) != ESVO_OK) return false;
      if (esvo_sgbm_compute(cs_->ctx(), nullptr, nullptr, (int)num_disparities_, (int)block_size_, (int)P1_, (int)P2_, -1, 0, 11, dispMap.data()) != ESVO_OK)
        return false;
      std::vector<esvo::DepthPoint> vdp_sgm;
      sgmPoints(dispMap, vdp_sgm);
      return accumulateNaive(vdp_sgm);
    }
    if (msm_ == PURE_BLOCK_MATCHING || msm_ == BM_PLUS_ESTIMATION) {                            // :383-431
      ebm_.createMatchProblem(&TS_obs_, &st_map_, &vCloseEventsPtr_left_);
      ebm_.match_all_HyperThread(vEMP_);
      if (msm_ == PURE_BLOCK_MATCHING) { std::vector<esvo::DepthPoint> vdp; vEMP2vDP(vEMP_, vdp); return accumulateNaive(vdp); }
    }
    // EM_PLUS_ESTIMATION / BM_PLUS_ESTIMATION: nonlinear optimisation, culling, fusion, clean, regularisation (:437-507)
    std::vector<esvo::DepthPoint> vdp;
    dpSolver_.solve(&vEMP_, &TS_obs_, vdp);
    dpSolver_.pointCulling(vdp, p.stdvar_vis_threshold, p.residual_vis_threshold * p.residual_vis_threshold * p.patch_size_x * p.patch_size_y,
                           p.invdepth_min_range, p.invdepth_max_range);
    dqvDepthPoints_.push_back(vdp);
    if (p.fusion_strategy == ESVO_FUSION_CONST_POINTS) {
      auto total = [&] { size_t n = 0; for (auto& v : dqvDepthPoints_) n += v.size(); return n; };
      while ((double)total() > 1.5 * (double)maxNumFusionPoints_) dqvDepthPoints_.pop_front();
    } else {
      while (dqvDepthPoints_.size() > maxNumFusionFrames_) dqvDepthPoints_.pop_front();
    }
    numFusionCount_ = 0;
    for (auto it = dqvDepthPoints_.rbegin(); it != dqvDepthPoints_.rend(); ++it) numFusionCount_ += (size_t)dFusor_.update(*it, depthFramePtr_, p.fusion_radius);
    dFusor_.clean(p.stdvar_vis_threshold * p.stdvar_vis_threshold, p.age_vis_threshold, p.invdepth_max_range, p.invdepth_min_range);
    if (p.regularization) dRegularizor_.apply();
    return true;
  }
  // vEMP2vDP (:1072-1097)
  void vEMP2vDP(std::vector<esvo::EventMatchPair>& vEMP, std::vector<esvo::DepthPoint>& vdp) {
    vdp.resize(vEMP.size());
    if (esvo_seeds_to_points(cs_->ctx(), vEMP.data(), vEMP.size(), vdp.data()) != ESVO_OK) vdp.clear();
  }
  void getElements(std::vector<esvo::DepthPoint>& out) { dFusor_.getElements(out); }
  void reset() { dqvDepthPoints_.clear(); }
  eMVStereoMode msm_;
  double EM_Slice_Thickness_ = 1e-3, EM_Time_THRESHOLD_ = 5e-5, EM_EPIPOLAR_THRESHOLD_ = 0.5, EM_TS_NCC_THRESHOLD_ = 0.1;   // :81-84
  size_t num_disparities_ = 16 * 3, block_size_ = 11, P1_ = 8 * 11 * 11, P2_ = 32 * 11 * 11;                                // :100-107
  size_t maxNumFusionFrames_ = 20, maxNumFusionPoints_ = 2000;
 private:
  // accumulation of modes 0, 1, 4 (:280-286,355-361,423-427): window of maxNumFusionFrames_ vectors, newest first, nearest wins
  bool accumulateNaive(std::vector<esvo::DepthPoint>& vdp) {
    dqvDepthPoints_.push_back(vdp);
    while (dqvDepthPoints_.size() > maxNumFusionFrames_) dqvDepthPoints_.pop_front();
    for (auto it = dqvDepthPoints_.rbegin(); it != dqvDepthPoints_.rend(); ++it) dFusor_.naive_propagation(*it, depthFramePtr_);
    return true;
  }
  // the PURE_SEMI_GLOBAL_MATCHING branch between the SGM call and the accumulation (:320-352) with createEdgeMask
  // (undistorted events, radius 0, :1130-1175) inlined: one Gaussian point per event whose rectified pixel lies right of
  // numDisparities and carries a non-negative disparity
  void sgmPoints(const std::vector<int16_t>& dispMap, std::vector<esvo::DepthPoint>& vdp_sgm) {
    const int W = cs_->width_, H = cs_->height_;
    if (lut_.empty()) { lut_.resize((size_t)2 * W * H); esvo_get_rectify_tables(cs_->ctx(), 0, nullptr, nullptr, lut_.data(), nullptr); }
    double d[4]; esvo_get_derived(cs_->ctx(), d);
    std::vector<esvo::EventMatchPair> pseudo;
    for (auto* e : vEventsPtr_left_SGM_) {
      const double* coor = &lut_[2 * ((size_t)e->y * W + e->x)];
      const int x = (int)std::floor(coor[0]), y = (int)std::floor(coor[1]);
      if (x < 0 || x >= W || y < 0 || y >= H) continue;
      if ((size_t)x < num_disparities_) continue;
      const double disp = dispMap[(size_t)y * W + x] / 16.0;
      if (disp < 0) continue;
      esvo::EventMatchPair s{};
      s.x_left[0] = x * 1.0; s.x_left[1] = y * 1.0;
      s.inv_depth = disp / (cs_->P_left_[0] * d[0]);
      s.cost = 0.0;
      std::copy(TS_obs_.second.tr_.begin(), TS_obs_.second.tr_.end(), s.T_world_virtual);
      pseudo.push_back(s);
    }
    vEMP2vDP(pseudo, vdp_sgm);
    for (size_t i = 0; i < vdp_sgm.size(); ++i) {       // DepthPoint dp(x, y): the reference passes (x, y) as (row, col) (:336)
      vdp_sgm[i].row = (int32_t)pseudo[i].x_left[0]; vdp_sgm[i].col = (int32_t)pseudo[i].x_left[1];
    }
  }
  core::EventMatcher em_;
  core::EventBM ebm_;
  core::DepthProblemSolver dpSolver_;
  core::DepthFusion dFusor_;
  core::DepthRegularization dRegularizor_;
  esvo::CameraSystem::Ptr cs_;
  core::DepthFrame::Ptr depthFramePtr_;
  PoseProvider getPoseAt_;
  std::vector<double> lut_;
};

// esvo_core::esvo_Tracking (esvo_Tracking.h:51-53; esvo_Tracking.cpp:79-265 TrackingLoop / refDataTransferring /
// curDataTransferring, :279-377 callbacks) without ROS.  The subscriptions become plain methods that fill the same buffers
// (refPCMap_, TS_history_, events_left_); the tf look-up becomes a pose provider; one pass of TrackingLoop's body is
// TrackingLoopOnce().  The registration itself runs on the device through RegProblemSolverLM (esvo_track_reset/solve).
class esvo_Tracking {
 public:
  enum TrackingStatus { IDLE, WORKING };
  using PoseProvider = std::function<bool(int64_t /*stamp ns*/, esvo::Pose& /*T_world_dvs*/)>;   // getPoseAt (:359-377)
  esvo_Tracking(esvo::CameraSystem::Ptr cs, core::RegProblemType rpType = core::REG_ANALYTICAL, size_t TS_HISTORY_LENGTH = 100,
                size_t REF_HISTORY_LENGTH = 5)
      : rpSolver_(cs, rpType), cs_(cs), rpType_(rpType), TS_HISTORY_LENGTH_(TS_HISTORY_LENGTH), REF_HISTORY_LENGTH_(REF_HISTORY_LENGTH) {
    T_world_cur_ = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  }
  void setPoseProvider(PoseProvider f) { getPoseAt_ = std::move(f); }
  // ---- callbacks (:279-357) ----
  // refMapCallback: the local map as the mapping node publishes it (pcl::PointXYZ: f32 x,y,z in the WORLD frame, i.e. the
  // output of frontend::packPointCloud); at most REF_HISTORY_LENGTH maps are kept.
  void refMapCallback(int64_t stamp_ns, const std::vector<float>& xyz) {
    refPCMap_[stamp_ns] = std::make_shared<std::vector<float>>(xyz);
    while (refPCMap_.size() > REF_HISTORY_LENGTH_) refPCMap_.erase(refPCMap_.begin());
  }
  // timeSurfaceCallback: mono8 left image (H*W); the tracker only reads the left one (:331-349).
  void timeSurfaceCallback(int64_t stamp_ns, const uint8_t* time_surface_left) {
    const size_t n = (size_t)cs_->width_ * cs_->height_;
    TS_history_[stamp_ns] = TsEntry{std::vector<uint8_t>(time_surface_left, time_surface_left + n), TS_id_++};
    while (TS_history_.size() > TS_HISTORY_LENGTH_) TS_history_.erase(TS_history_.begin());
  }
  // eventsCallback: insertion-sorted event buffer, capped at 5e6 (:294-321)
  void eventsCallback(const std::vector<esvo::Event>& events) {
    for (const esvo::Event& e : events) {
      events_left_.push_back(e);
      long i = (long)events_left_.size() - 2;
      while (i >= 0 && events_left_[(size_t)i].ts > e.ts) { events_left_[(size_t)i + 1] = events_left_[(size_t)i]; --i; }
      events_left_[(size_t)(i + 1)] = e;
    }
    static constexpr size_t MAX_EVENT_QUEUE_LENGTH = 5000000;
    if (events_left_.size() > MAX_EVENT_QUEUE_LENGTH)
      events_left_.erase(events_left_.begin(), events_left_.begin() + (long)(events_left_.size() - MAX_EVENT_QUEUE_LENGTH));
  }
  // ---- refDataTransferring (:177-211) ----
  bool refDataTransferring() {
    ref_.t_ = refPCMap_.rbegin()->first;
    if (ESVO_System_Status_ == "INITIALIZATION" && ets_ == IDLE) ref_.tr_ = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if (ESVO_System_Status_ == "WORKING" || (ESVO_System_Status_ == "INITIALIZATION" && ets_ == WORKING)) {
      if (!getPoseAt_ || !getPoseAt_(ref_.t_, ref_.tr_)) return false;   // the reference exit(-1)s here ("logic error")
    }
    ref_.vPointXYZ_ = *refPCMap_.rbegin()->second;    // the reference keeps pointers into the cloud; the device path permutes a copy
    return true;
  }
  // ---- curDataTransferring (:213-245) ----
  bool curDataTransferring() {
    const size_t ev_last = lower_bound_ev(cur_.t_);
    auto TS_it = TS_history_.rbegin();
    if (cur_.t_ == TS_it->first) return false;        // TS_history may not have been updated yet
    cur_.t_ = TS_it->first;
    cur_.ts_left = TS_it->second.left.data();
    if (ESVO_System_Status_ == "INITIALIZATION" && ets_ == IDLE) cur_.tr_ = ref_.tr_;
    if (ESVO_System_Status_ == "WORKING" || (ESVO_System_Status_ == "INITIALIZATION" && ets_ == WORKING)) cur_.tr_ = T_world_cur_;
    const size_t ev_cur = lower_bound_ev(cur_.t_);
    cur_.numEventsSinceLastObs_ = ev_cur - ev_last + 1;
    return true;
  }
  // ---- one pass of TrackingLoop's body (:84-171).  Returns true when a pose was produced (then also in T_world_cur_). ----
  bool TrackingLoopOnce() {
    if (refPCMap_.size() < 1 || TS_history_.size() < 1) return false;                       // keep idling
    if (ESVO_System_Status_ == "INITIALIZATION" && ets_ == WORKING) { reset(); return false; }
    if (ESVO_System_Status_ == "TERMINATE") return false;
    if (frontend::toSec(ref_.t_) < frontend::toSec(refPCMap_.rbegin()->first))            // new reference map arrived
      if (!refDataTransferring()) return false;
    if (frontend::toSec(cur_.t_) < frontend::toSec(TS_history_.rbegin()->first)) {          // new observation arrived
      if (frontend::toSec(ref_.t_) >= frontend::toSec(TS_history_.rbegin()->first)) return false;   // obs must come after the ref (reference: exit(-1))
      if (!curDataTransferring()) return false;
    } else return false;
    if (rpSolver_.resetRegProblem(&ref_, &cur_)) {
      if (ets_ == IDLE) ets_ = WORKING;
      if (ESVO_System_Status_ != "WORKING") ESVO_System_Status_ = "WORKING";
      const bool ok = rpType_ == core::REG_NUMERICAL ? rpSolver_.solve_numerical() : rpSolver_.solve_analytical();
      if (!ok) return false;
      T_world_cur_ = cur_.tr_;
      lTimestamp_.push_back(cur_.t_); lPose_.push_back(cur_.tr_);                             // publishPose / saveTrajectory payload
      return true;
    }
    ESVO_System_Status_ = "INITIALIZATION";
    ets_ = IDLE;
    return false;
  }
  void reset() { ets_ = IDLE; TS_id_ = 0; TS_history_.clear(); refPCMap_.clear(); events_left_.clear(); }   // (:247-255)

  std::string ESVO_System_Status_ = "INITIALIZATION";   // the /ESVO_SYSTEM_STATUS parameter
  TrackingStatus ets_ = IDLE;
  core::RefFrame ref_;
  core::CurFrame cur_;
  esvo::Pose T_world_cur_;
  std::vector<int64_t> lTimestamp_;
  std::vector<esvo::Pose> lPose_;
  core::RegProblemSolverLM rpSolver_;
 private:
  size_t lower_bound_ev(int64_t t) const {   // tools::EventBuffer_lower_bound (utils.h:50-55): toSec() doubles
    const double ts = frontend::toSec(t);
    size_t lo = 0, hi = events_left_.size();
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (frontend::toSec(events_left_[mid].ts) < ts) lo = mid + 1; else hi = mid; }
    return lo;
  }
  struct TsEntry { std::vector<uint8_t> left; size_t id; };
  esvo::CameraSystem::Ptr cs_;
  core::RegProblemType rpType_;
  size_t TS_HISTORY_LENGTH_, REF_HISTORY_LENGTH_, TS_id_ = 0;
  std::map<int64_t, std::shared_ptr<std::vector<float>>> refPCMap_;
  std::map<int64_t, TsEntry> TS_history_;
  std::deque<esvo::Event> events_left_;
  PoseProvider getPoseAt_;
};
}  // namespace esvo_core
