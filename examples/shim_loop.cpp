// Mapping -> local-map hand-off -> Tracking through the reference-shaped C++ shim (include/esvo_b200/esvo_core.hpp),
// driven by a binary scenario file so that a test can feed the SAME inputs to the CPU oracle and compare
// (tests/test_cpp_shim.py::test_shim_mapping_to_tracking_matches_oracle).  Sequence, as the two reference nodes run it:
//   mapping node : dataTransferring (frontend::selectCloseEvents / samplePoseStamps) -> esvo_Mapping::MappingAtTime
//                  -> publishPointCloud (frontend::packPointCloud)
//   tracking node: refMapCallback / timeSurfaceCallback / eventsCallback -> esvo_Tracking::TrackingLoopOnce
// usage: shim_loop <scenario.bin> <result.bin>
// scenario: i32 W,H,rig(0 hkust); i64 t_obs, t_cur; f64 T_world_left[16]; f64 bm_half_slice, i32 process_event_num;
//           i32 n_events; events (u16 x, u16 y, i64 t, u8 pol) SoA; u8 ts_left[W*H], ts_right[W*H], ts_cur_left[W*H];
//           i32 n_traj; traj stamps i64[n_traj]; traj poses f64[n_traj*16]   (pose look-up = nearest stamp at or after t)
#include <cstdio>
#include <cstring>
#include <fstream>

#include "esvo_b200/esvo_core.hpp"

template <class T> static bool rd(std::ifstream& f, T* p, size_t n) { f.read((char*)p, (std::streamsize)(n * sizeof(T))); return (bool)f; }

int main(int argc, char** argv) {
  if (argc < 3) { std::printf("usage: shim_loop <scenario.bin> <result.bin>\n"); return 64; }
  std::ifstream f(argv[1], std::ios::binary);
  int32_t W, H, rig; int64_t t_obs, t_cur; double T[16], half_slice; int32_t process_event_num, n_ev;
  if (!rd(f, &W, 1) || !rd(f, &H, 1) || !rd(f, &rig, 1) || !rd(f, &t_obs, 1) || !rd(f, &t_cur, 1) || !rd(f, T, 16) || !rd(f, &half_slice, 1) ||
      !rd(f, &process_event_num, 1) || !rd(f, &n_ev, 1)) return 65;
  std::vector<uint16_t> ex(n_ev), ey(n_ev); std::vector<int64_t> et(n_ev); std::vector<uint8_t> ep(n_ev);
  rd(f, ex.data(), n_ev); rd(f, ey.data(), n_ev); rd(f, et.data(), n_ev); rd(f, ep.data(), n_ev);
  std::vector<uint8_t> tl((size_t)W * H), tr((size_t)W * H), tc((size_t)W * H);
  rd(f, tl.data(), tl.size()); rd(f, tr.data(), tr.size()); rd(f, tc.data(), tc.size());
  int32_t n_traj; rd(f, &n_traj, 1);
  std::vector<int64_t> tt(n_traj); std::vector<double> tp((size_t)n_traj * 16);
  rd(f, tt.data(), n_traj);
  if (!rd(f, tp.data(), tp.size())) return 65;

  // calib/hkust + cfg/{mapping,tracking}/*_hkust.yaml (the values tests/configs.py uses)
  esvo_calib l{}, r{};
  l.width = r.width = W; l.height = r.height = H;
  const double Kl[9] = {263.796, 0, 176.994, 0, 263.738, 124.373, 0, 0, 1}, Dl[4] = {-0.386589, 0.157241, 0.000322143, 6.13759e-06};
  const double Rl[9] = {0.999809, 0.0161928, 0.0109163, -0.0162088, 0.999868, 0.0013701, -0.0108927, -0.00154678, 0.999939};
  const double Pl[12] = {189.705, 0, 165.382, 0, 0, 189.705, 121.295, 0, 0, 0, 1, 0};
  const double Kr[9] = {263.485, 0, 162.942, 0, 263.276, 118.029, -0.0151344, 0.00133093, 0.999885}, Dr[4] = {-0.383425, 0.152823, -0.000257745, 0.000268432};
  const double Rr[9] = {0.9993960957463914, 0.0034732142808621717, -0.03457427641222047, -0.0035085878889783376, 0.9999933816804096,
                        -0.0009625000798637905, 0.03457070461958685, 0.0010832257094615543, 0.9994016675011942};
  double Pr[12]; std::memcpy(Pr, Pl, sizeof(Pl)); Pr[3] = -13.8634;
  std::memcpy(l.K, Kl, 72); std::memcpy(l.D, Dl, 32); std::memcpy(l.R, Rl, 72); std::memcpy(l.P, Pl, 96);
  std::memcpy(r.K, Kr, 72); std::memcpy(r.D, Dr, 32); std::memcpy(r.R, Rr, 72); std::memcpy(r.P, Pr, 96);
  esvo_params p; esvo_default_params(&p);
  p.patch_size_x = 15; p.patch_size_y = 7; p.bm_step = 1; p.bm_zncc_threshold = 0.1; p.lsnorm = ESVO_LSNORM_TDIST; p.max_iteration = 10;
  p.age_vis_threshold = 1; p.fusion_strategy = ESVO_FUSION_CONST_FRAMES; p.num_thread_mapping = 4;
  p.trk_patch_size_x = p.trk_patch_size_y = 1; p.trk_kernel_size = 5; p.trk_lsnorm = ESVO_TRK_LSNORM_HUBER; p.trk_huber_threshold = 50;
  p.trk_max_registration_points = 2000; p.trk_max_iteration = 10; p.trk_min_num_events = 1000; p.trk_batch_size = 500;
  p.invdepth_min_range = 0.25; p.invdepth_max_range = 2.0; p.residual_vis_threshold = 20; p.stdvar_vis_threshold = 0.15;
  p.fusion_radius = 0; p.max_num_fusion_frames = 20; p.max_num_fusion_points = 4000; p.smooth_time_surface = 0; p.regularization = 1;
  p.reg_radius = 5; p.reg_min_neighbours = 8; p.reg_min_close_neighbours = 8; p.td_nu = 2.1897; p.td_scale = 16.6397;
  p.bm_min_disparity = 1; p.bm_max_disparity = 40;
  esvo::CameraSystem::Ptr cs;
  try { cs = std::make_shared<esvo::CameraSystem>(l, r, p, 0); }
  catch (const std::exception& e) { std::printf("%s\n", e.what()); return 2; }

  auto poseAt = [&](int64_t t, esvo::Pose& out) {   // stand-in for tf: first trajectory sample at or after t
    for (int i = 0; i < n_traj; ++i) if (tt[(size_t)i] >= t) { std::memcpy(out.data(), &tp[(size_t)i * 16], 128); return true; }
    return false;
  };
  // ---------------- mapping node ----------------
  std::vector<esvo::Event> events_left(n_ev);
  for (int i = 0; i < n_ev; ++i) events_left[(size_t)i] = {ex[(size_t)i], ey[(size_t)i], et[(size_t)i], ep[(size_t)i] != 0};
  std::vector<esvo::Event*> vCloseEventsPtr_left;
  esvo_core::frontend::selectCloseEvents(events_left, t_obs, half_slice, (size_t)process_event_num, vCloseEventsPtr_left);
  esvo::StampTransformationMap st_map;
  for (int64_t ts : esvo_core::frontend::samplePoseStamps(t_obs, half_slice)) { esvo::Pose q; if (poseAt(ts, q)) st_map.push_back({ts, q}); }
  esvo::StampedTimeSurfaceObs obs; obs.first = t_obs; obs.second.left = tl.data(); obs.second.right = tr.data();
  std::memcpy(obs.second.tr_.data(), T, 128);
  esvo_core::esvo_Mapping mapping(cs);
  esvo_core::esvo_Mapping::Counters ctr{};
  if (!mapping.MappingAtTime(obs, vCloseEventsPtr_left, st_map, &ctr)) { std::printf("MappingAtTime failed: %s\n", esvo_last_error(cs->ctx())); return 3; }
  std::vector<esvo::DepthPoint> elems;
  esvo_core::core::DepthFusion fusor(cs);
  fusor.getElements(elems);
  std::vector<float> cloud;
  esvo_core::frontend::packPointCloud(elems, obs.second.tr_, cloud);
  // ---------------- tracking node ----------------
  esvo_core::esvo_Tracking tracking(cs, esvo_core::core::REG_ANALYTICAL);
  tracking.setPoseProvider(poseAt);
  tracking.eventsCallback(events_left);
  tracking.refMapCallback(t_obs, cloud);
  tracking.timeSurfaceCallback(t_cur, tc.data());
  esvo_track_srand(cs->ctx(), 1);
  const bool idle_first = tracking.TrackingLoopOnce();            // INITIALIZATION + IDLE: ref pose = identity, cur pose = ref pose
  // a node that is already WORKING (the usual state): reference pose from the trajectory, prior = last tracked pose
  esvo_core::esvo_Tracking tracking2(cs, esvo_core::core::REG_ANALYTICAL);
  tracking2.setPoseProvider(poseAt);
  tracking2.ESVO_System_Status_ = "WORKING"; tracking2.ets_ = esvo_core::esvo_Tracking::WORKING;
  std::memcpy(tracking2.T_world_cur_.data(), T, 128);
  tracking2.eventsCallback(events_left);
  tracking2.refMapCallback(t_obs, cloud);
  tracking2.timeSurfaceCallback(t_cur, tc.data());
  esvo_track_srand(cs->ctx(), 1);
  const bool ok2 = tracking2.TrackingLoopOnce();

  std::ofstream o(argv[2], std::ios::binary);
  const int32_t hdr[8] = {(int32_t)vCloseEventsPtr_left.size(), (int32_t)st_map.size(), (int32_t)elems.size(), (int32_t)(cloud.size() / 3),
                          idle_first ? 1 : 0, ok2 ? 1 : 0, (int32_t)tracking2.cur_.numEventsSinceLastObs_, (int32_t)tracking2.rpSolver_.lmStatics_.nIter_};
  o.write((const char*)hdr, sizeof(hdr));
  o.write((const char*)&ctr, sizeof(ctr));
  o.write((const char*)tracking.T_world_cur_.data(), 128);
  o.write((const char*)tracking2.T_world_cur_.data(), 128);
  for (auto* e : vCloseEventsPtr_left) { const int64_t idx = e - events_left.data(); o.write((const char*)&idx, 8); }
  for (auto& st : st_map) o.write((const char*)&st.first, 8);
  o.write((const char*)cloud.data(), (std::streamsize)(cloud.size() * 4));
  std::printf("shim loop: %zu close events, %zu poses, map %zu, tracked(idle)=%d tracked(working)=%d nIter=%zu\n", vCloseEventsPtr_left.size(),
              st_map.size(), elems.size(), (int)idle_first, (int)ok2, tracking2.rpSolver_.lmStatics_.nIter_);
  return 0;
}
