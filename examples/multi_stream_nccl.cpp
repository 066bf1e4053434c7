// N independent stereo event streams, one per GPU, in ONE C++ host process (one thread per GPU), results gathered with a
// single ncclAllGather of the fixed stream record (include/esvo_b200/multi_stream.hpp) -- SURVEY.md 8e.  No data-path
// collective: every stream runs esvo_Mapping::MappingAtTime on its own esvo_ctx.
// usage: multi_stream_nccl <scenario.bin> [n_gpus = all] [frames = 3]      (scenario format: examples/shim_loop.cpp)
#include <cstdio>
#include <cstring>
#include <fstream>
#include <thread>

#include "esvo_b200/multi_stream.hpp"

template <class T> static bool rd(std::ifstream& f, T* p, size_t n) { f.read((char*)p, (std::streamsize)(n * sizeof(T))); return (bool)f; }

int main(int argc, char** argv) {
  if (argc < 2) { std::printf("usage: multi_stream_nccl <scenario.bin> [n_gpus] [frames]\n"); return 64; }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { std::printf("no CUDA device (no CPU fallback)\n"); return 2; }
  const int world = argc > 2 ? std::min(std::atoi(argv[2]), n_dev) : n_dev;
  const int frames = argc > 3 ? std::atoi(argv[3]) : 3;
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

  std::vector<int> devs(world);
  for (int i = 0; i < world; ++i) devs[(size_t)i] = i;
  std::vector<ncclComm_t> comms(world);
  if (ncclCommInitAll(comms.data(), world, devs.data()) != ncclSuccess) { std::printf("ncclCommInitAll failed\n"); return 3; }

  std::vector<std::vector<double>> gathered(world);
  std::vector<int> status(world, 0);
  auto run_stream = [&](int rank) {
    try {
      auto cs = std::make_shared<esvo::CameraSystem>(l, r, p, rank);
      // every stream takes its own slice of the event buffer (rank-dependent budget) so that the records differ
      std::vector<esvo::Event> events_left(n_ev);
      for (int i = 0; i < n_ev; ++i) events_left[(size_t)i] = {ex[(size_t)i], ey[(size_t)i], et[(size_t)i], ep[(size_t)i] != 0};
      std::vector<esvo::Event*> close;
      esvo_core::frontend::selectCloseEvents(events_left, t_obs, half_slice, (size_t)process_event_num - 200 * (size_t)rank, close);
      esvo::StampTransformationMap st_map;
      for (int64_t ts : esvo_core::frontend::samplePoseStamps(t_obs, half_slice))
        for (int i = 0; i < n_traj; ++i) if (tt[(size_t)i] >= ts) { esvo::Pose q; std::memcpy(q.data(), &tp[(size_t)i * 16], 128); st_map.push_back({ts, q}); break; }
      esvo::StampedTimeSurfaceObs obs; obs.first = t_obs; obs.second.left = tl.data(); obs.second.right = tr.data();
      std::memcpy(obs.second.tr_.data(), T, 128);
      esvo_core::esvo_Mapping mapping(cs);
      esvo_core::esvo_Mapping::Counters ctr{};
      for (int k = 0; k < frames; ++k)
        if (!mapping.MappingAtTime(obs, close, st_map, &ctr)) { status[(size_t)rank] = 4; return; }
      std::vector<esvo::DepthPoint> elems;
      esvo_core::core::DepthFusion fusor(cs);
      fusor.getElements(elems);
      double rec[esvo::kStreamRecordFields];
      esvo::makeStreamRecord(rank, frames, ctr, esvo::mapChecksum(elems), rec);
      if (!esvo::gatherStreamRecords(comms[(size_t)rank], world, rank, rec, gathered[(size_t)rank])) status[(size_t)rank] = 5;
    } catch (const std::exception& e) { std::printf("rank %d: %s\n", rank, e.what()); status[(size_t)rank] = 2; }
  };
  std::vector<std::thread> th;
  for (int rank = 0; rank < world; ++rank) th.emplace_back(run_stream, rank);
  for (auto& t : th) t.join();
  for (int rank = 0; rank < world; ++rank) ncclCommDestroy(comms[(size_t)rank]);
  for (int rank = 0; rank < world; ++rank) if (status[(size_t)rank]) { std::printf("rank %d failed with %d\n", rank, status[(size_t)rank]); return status[(size_t)rank]; }
  // every rank holds the same table; print rank 0's and check that the others agree
  for (int rank = 1; rank < world; ++rank) if (gathered[(size_t)rank] != gathered[0]) { std::printf("rank %d received a different table\n", rank); return 6; }
  std::printf("streams %d frames %d\n", world, frames);
  for (int s = 0; s < world; ++s) {
    std::printf("record");
    for (int q = 0; q < esvo::kStreamRecordFields; ++q) std::printf(" %.17g", gathered[0][(size_t)s * esvo::kStreamRecordFields + q]);
    std::printf("\n");
  }
  return 0;
}
