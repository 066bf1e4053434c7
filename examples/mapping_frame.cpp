// Minimal C++ use of the reference-shaped shim (include/esvo_b200/esvo_core.hpp): one time-surface pair,
// block matching, LM refinement, culling, fusion.  Build:
//   g++ -std=c++17 -Iinclude examples/mapping_frame.cpp -Lesvo_b200/_build -lesvo_b200 -Wl,-rpath,$PWD/esvo_b200/_build -o mapping_frame
#include <cstdio>
#include <cstring>

#include "esvo_b200/esvo_core.hpp"

int main() {
  esvo_calib l{}, r{};
  l.width = r.width = 346; l.height = r.height = 260;
  const double K[9] = {263.796, 0, 176.994, 0, 263.738, 124.373, 0, 0, 1}, I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  const double Pl[12] = {189.705, 0, 165.382, 0, 0, 189.705, 121.295, 0, 0, 0, 1, 0};
  double Pr[12]; std::memcpy(Pr, Pl, sizeof(Pl)); Pr[3] = -13.8634;
  std::memcpy(l.K, K, sizeof(K)); std::memcpy(r.K, K, sizeof(K)); std::memcpy(l.R, I3, sizeof(I3)); std::memcpy(r.R, I3, sizeof(I3));
  std::memcpy(l.P, Pl, sizeof(Pl)); std::memcpy(r.P, Pr, sizeof(Pr));
  esvo_params p; esvo_default_params(&p);
  p.patch_size_x = 15; p.patch_size_y = 7; p.td_nu = 2.1897; p.td_scale = 16.6397; p.invdepth_min_range = 0.25; p.invdepth_max_range = 2;
  p.bm_min_disparity = 1; p.residual_vis_threshold = 20; p.stdvar_vis_threshold = 0.15;
  esvo::CameraSystem::Ptr cs;
  try { cs = std::make_shared<esvo::CameraSystem>(l, r, p, 0); }
  catch (const std::exception& e) { std::printf("%s\n", e.what()); return 2; }
  // synthetic fronto-parallel pair, disparity 9
  std::vector<uint8_t> tl(346 * 260, 0), tr(346 * 260, 0);
  for (int y = 40; y < 220; ++y) for (int x = 60; x < 300; ++x) tl[y * 346 + x] = (uint8_t)(30 + (x * 37 + y * 91) % 200);
  for (int y = 0; y < 260; ++y) for (int x = 0; x + 9 < 346; ++x) tr[y * 346 + x] = tl[y * 346 + x + 9];
  esvo::StampedTimeSurfaceObs obs; obs.first = 1000000000; obs.second.left = tl.data(); obs.second.right = tr.data();
  obs.second.tr_ = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  std::vector<esvo::Event> ev; std::vector<esvo::Event*> evp;
  for (int i = 0; i < 500; ++i) ev.push_back({(uint16_t)(100 + i % 150), (uint16_t)(60 + (i * 7) % 140), 1000000000, true});
  for (auto& e : ev) evp.push_back(&e);
  esvo::StampTransformationMap st_map{{1000000000, obs.second.tr_}};
  esvo_core::core::EventBM ebm(cs);
  std::vector<esvo::EventMatchPair> vEMP;
  ebm.createMatchProblem(&obs, &st_map, &evp);
  ebm.match_all_HyperThread(vEMP);
  esvo_core::core::DepthProblemSolver solver(cs);
  std::vector<esvo::DepthPoint> vdp;
  solver.solve(&vEMP, &obs, vdp);
  solver.pointCulling(vdp, p.stdvar_vis_threshold, p.residual_vis_threshold * p.residual_vis_threshold * 105, 0.25, 2.0);
  auto df = std::make_shared<esvo_core::core::DepthFrame>(); df->setTransformation(obs.second.tr_);
  esvo_core::core::DepthFusion fusor(cs);
  int nf = fusor.update(vdp, df, 0);
  std::printf("BM %zu seeds (%llu evals), LM %zu points (%llu evals), %d fusions; rho[0]=%.6f (expect %.6f)\n", vEMP.size(),
              (unsigned long long)ebm.n_evals_, vdp.size(), (unsigned long long)solver.n_evals_, nf, vdp.empty() ? 0.0 : vdp[0].inv_depth,
              9 / 13.8634);
  return vdp.empty() ? 1 : 0;
}
