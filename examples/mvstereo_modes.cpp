// The five MVStereoMode sequences of esvo_MVStereo::MappingAtTime (esvo_core/src/esvo_MVStereo.cpp:239-520) through the
// reference-shaped C++ shim, driven by a binary scenario file so that a test can run the same inputs through the CPU oracle
// (tests/test_cpp_shim.py::test_shim_mvstereo_modes_match_oracle).
// usage: mvstereo_modes <scenario.bin> <result.bin>
// scenario: i32 W,H; i64 t_obs, t_low, t_up; f64 T_world_left[16]; u8 ts_left[W*H], ts_right[W*H];
//           i32 n_left; left events SoA (u16 x, u16 y, i64 t, u8 pol); i32 n_right; right events SoA;
//           i32 n_close; close events SoA (u16 x, u16 y, i64 t);  i32 n_st; st stamps i64[n_st]; st poses f64[n_st*16];
//           i32 n_traj; traj stamps i64[n_traj]; traj poses f64[n_traj*16]   (pose look-up = the sample with exactly that stamp)
// result: per mode 0..4: i32 n_match, i32 n_map, then n_map x (i32 row, i32 col, f64 inv_depth, f64 variance, f64 residual, i64 age)
#include <cstdio>
#include <cstring>
#include <fstream>

#include "esvo_b200/esvo_core.hpp"

template <class T> static bool rd(std::ifstream& f, T* p, size_t n) { f.read((char*)p, (std::streamsize)(n * sizeof(T))); return (bool)f; }

int main(int argc, char** argv) {
  if (argc < 3) { std::printf("usage: mvstereo_modes <scenario.bin> <result.bin>\n"); return 64; }
  std::ifstream f(argv[1], std::ios::binary);
  int32_t W, H; int64_t t_obs, t_low, t_up; double T[16];
  if (!rd(f, &W, 1) || !rd(f, &H, 1) || !rd(f, &t_obs, 1) || !rd(f, &t_low, 1) || !rd(f, &t_up, 1) || !rd(f, T, 16)) return 65;
  std::vector<uint8_t> tl((size_t)W * H), tr((size_t)W * H);
  rd(f, tl.data(), tl.size()); rd(f, tr.data(), tr.size());
  auto read_events = [&](std::vector<esvo::Event>& ev, bool with_pol) {
    int32_t n; rd(f, &n, 1);
    std::vector<uint16_t> x(n), y(n); std::vector<int64_t> t(n); std::vector<uint8_t> p(n, 1);
    rd(f, x.data(), n); rd(f, y.data(), n); rd(f, t.data(), n); if (with_pol) rd(f, p.data(), n);
    ev.resize(n);
    for (int i = 0; i < n; ++i) ev[(size_t)i] = {x[(size_t)i], y[(size_t)i], t[(size_t)i], p[(size_t)i] != 0};
  };
  std::vector<esvo::Event> left, right, close;
  read_events(left, true); read_events(right, true); read_events(close, false);
  int32_t n_st; rd(f, &n_st, 1);
  std::vector<int64_t> st_t(n_st); std::vector<double> st_p((size_t)n_st * 16);
  rd(f, st_t.data(), n_st); rd(f, st_p.data(), st_p.size());
  int32_t n_traj; rd(f, &n_traj, 1);
  std::vector<int64_t> tt(n_traj); std::vector<double> tp((size_t)n_traj * 16);
  rd(f, tt.data(), n_traj);
  if (!rd(f, tp.data(), tp.size())) return 65;

  // calib/hkust + cfg/mvstereo-style parameters (the values tests/configs.py uses for the hkust rig)
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
  p.invdepth_min_range = 0.25; p.invdepth_max_range = 2.0; p.residual_vis_threshold = 20; p.stdvar_vis_threshold = 0.15;
  p.fusion_radius = 0; p.max_num_fusion_frames = 20; p.max_num_fusion_points = 4000; p.smooth_time_surface = 0; p.regularization = 1;
  p.reg_radius = 5; p.reg_min_neighbours = 8; p.reg_min_close_neighbours = 8; p.td_nu = 2.1897; p.td_scale = 16.6397;
  p.bm_min_disparity = 1; p.bm_max_disparity = 40;
  esvo::CameraSystem::Ptr cs;
  try { cs = std::make_shared<esvo::CameraSystem>(l, r, p, 0); }
  catch (const std::exception& e) { std::printf("%s\n", e.what()); return 2; }

  auto poseAt = [&](int64_t t, esvo::Pose& out) {
    for (int i = 0; i < n_traj; ++i) if (tt[(size_t)i] == t) { std::memcpy(out.data(), &tp[(size_t)i * 16], 128); return true; }
    return false;
  };
  std::ofstream o(argv[2], std::ios::binary);
  for (int mode = 0; mode < 5; ++mode) {
    esvo_core::esvo_MVStereo node(cs, (esvo_core::esvo_MVStereo::eMVStereoMode)mode, 4);
    node.setPoseProvider(poseAt);
    node.resetEMParameters(1e-3, 5e-4, 1.0, 0.1);                       // cfg/mvstereo/mvstereo_rpg.yaml:25-28
    node.TS_obs_.first = t_obs; node.TS_obs_.second.left = tl.data(); node.TS_obs_.second.right = tr.data();
    std::memcpy(node.TS_obs_.second.tr_.data(), T, 128);
    node.t_lowBound_ = t_low; node.t_upBound_ = t_up;
    for (auto& e : left) node.vEventsPtr_left_.push_back(&e);
    for (auto& e : right) node.vEventsPtr_right_.push_back(&e);
    for (auto& e : close) { node.vCloseEventsPtr_left_.push_back(&e); node.vEventsPtr_left_SGM_.push_back(&e); }
    for (int i = 0; i < n_st; ++i) { esvo::Pose q; std::memcpy(q.data(), &st_p[(size_t)i * 16], 128); node.st_map_.push_back({st_t[(size_t)i], q}); }
    const bool ok = node.MappingAtTime();
    std::vector<esvo::DepthPoint> elems;
    if (ok) node.getElements(elems);
    const int32_t hdr[2] = {(int32_t)node.vEMP_.size(), (int32_t)elems.size()};
    o.write((const char*)hdr, sizeof(hdr));
    for (auto& d : elems) {
      o.write((const char*)&d.row, 4); o.write((const char*)&d.col, 4); o.write((const char*)&d.inv_depth, 8);
      o.write((const char*)&d.variance, 8); o.write((const char*)&d.residual, 8); o.write((const char*)&d.age, 8);
    }
    std::printf("mode %d: ok=%d matches %zu, map %zu\n", mode, (int)ok, node.vEMP_.size(), elems.size());
  }
  return 0;
}
