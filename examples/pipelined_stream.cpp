// Pipelined use of the C ABI (include/esvo_b200.h): S frames in flight, every call only enqueues work, results are
// collected S-1 frames later.  This is how bench.py reaches 0.32 ms per mapping frame; results are identical to strictly
// sequential operation (tests/test_gpu_full_size.py::test_pipeline_depth16_wraps_around).  Build:
//   g++ -std=c++17 -Iinclude examples/pipelined_stream.cpp -Lesvo_b200/_build -lesvo_b200 -Wl,-rpath,$PWD/esvo_b200/_build -o pipelined_stream
// A real node fills FrameInputs from its event queues and tf (esvo_core::frontend::selectCloseEvents / samplePoseStamps in
// include/esvo_b200/esvo_core.hpp restate that selection); here they are synthetic.
#include <cstdio>
#include <cstring>
#include <deque>
#include <vector>

#include "esvo_b200.h"

struct FrameInputs {                       // what one call of esvo_Mapping::dataTransferring hands over
  std::vector<uint16_t> ex[2], ey[2]; std::vector<int64_t> et[2]; std::vector<uint8_t> ep[2];   // new events of both cameras
  int64_t t_ts_ns;                                                                                 // time-surface stamp
  double T_world_left[16];
  std::vector<uint16_t> sx, sy; std::vector<int64_t> st;                                           // vCloseEventsPtr_left_
  std::vector<int64_t> pose_t; std::vector<double> poses;                                          // st_map_
};

static FrameInputs synthetic_frame(int k) {
  FrameInputs f;
  const int64_t t0 = 1000000000LL + (int64_t)k * 10000000LL;      // 10 ms per frame
  for (int cam = 0; cam < 2; ++cam)
    for (int i = 0; i < 20000; ++i) {
      const int x = 30 + (i * 7 + k * 3) % 280, y = 20 + (i * 13) % 220;
      f.ex[cam].push_back((uint16_t)(x - (cam ? 9 : 0))); f.ey[cam].push_back((uint16_t)y);
      f.et[cam].push_back(t0 + (int64_t)i * 499); f.ep[cam].push_back(1);
    }
  f.t_ts_ns = t0 + 10000000LL - 1;
  const double I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  std::memcpy(f.T_world_left, I, sizeof(I));
  for (int i = 0; i < 2000; ++i) { f.sx.push_back(f.ex[0][19999 - i]); f.sy.push_back(f.ey[0][19999 - i]); f.st.push_back(f.et[0][19999 - i]); }
  f.pose_t.push_back(t0); f.poses.insert(f.poses.end(), I, I + 16);
  return f;
}

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
  p.bm_min_disparity = 1; p.residual_vis_threshold = 20; p.stdvar_vis_threshold = 0.15; p.max_num_fusion_frames = 5;
  int status = 0;
  esvo_ctx* ctx = esvo_create(0, &l, &r, &p, &status);
  if (!ctx) { std::printf("esvo_create failed (status %d): a CUDA device is required, there is no CPU fallback\n", status); return 2; }
  const int S = 8, FRAMES = 24;
  if (esvo_set_pipeline_depth(ctx, S) != ESVO_OK) return 3;
  std::deque<int64_t> tickets;
  std::deque<FrameInputs> in_flight;             // host buffers must outlive the asynchronous copies made from them
  std::vector<esvo_depth_point> map((size_t)346 * 260);
  auto collect = [&]() {
    size_t n = map.size(); uint64_t c[8];
    if (esvo_results_end(ctx, tickets.front(), map.data(), &n, c) != ESVO_OK) { std::printf("results_end: %s\n", esvo_last_error(ctx)); return false; }
    std::printf("frame %lld: %llu seeds, %llu points after culling, %llu fusions, map %zu\n", (long long)tickets.front(),
                (unsigned long long)c[1], (unsigned long long)c[3], (unsigned long long)c[4], n);
    tickets.pop_front(); in_flight.pop_front();
    return true;
  };
  for (int k = 0; k < FRAMES; ++k) {
    if ((int)tickets.size() >= S - 1 && !collect()) return 4;
    in_flight.push_back(synthetic_frame(k));
    FrameInputs& f = in_flight.back();
    for (int cam = 0; cam < 2; ++cam) {
      if (esvo_stage_ts_events(ctx, cam, f.ex[cam].data(), f.ey[cam].data(), f.et[cam].data(), f.ep[cam].data(), f.ex[cam].size()) != ESVO_OK ||
          esvo_run_ts_build(ctx, cam, f.t_ts_ns) != ESVO_OK) { std::printf("ts: %s\n", esvo_last_error(ctx)); return 5; }
    }
    if (esvo_set_ts_pair_dev(ctx, f.T_world_left) != ESVO_OK ||                       // zero-copy hand-off of the freshly built pair
        esvo_stage_mapping_inputs(ctx, f.sx.data(), f.sy.data(), f.st.data(), f.sx.size(), f.pose_t.data(), f.poses.data(), f.pose_t.size()) != ESVO_OK ||
        esvo_run_mapping(ctx) != ESVO_OK) { std::printf("mapping: %s\n", esvo_last_error(ctx)); return 6; }
    int64_t t; if (esvo_results_begin(ctx, &t) != ESVO_OK) return 7;
    tickets.push_back(t);
  }
  while (!tickets.empty()) if (!collect()) return 4;
  esvo_destroy(ctx);
  return 0;
}
