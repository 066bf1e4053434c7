// Dumps what the host-side event front-end (esvo_core::frontend, include/esvo_b200/esvo_core.hpp) selects for a given event
// buffer, so that a test can compare it with an independent re-derivation (tests/indep_numpy.py) -- no GPU, no library needed.
// usage: frontend_dump <in.bin> <out.bin>
// in : i32 n; i64 t[n] (time-ordered stamps, ns); i64 t_end; f64 half_slice; i32 process_event_num; i64 t_low; i64 t_up; f64 em_thickness
// out: i32 n_close; i64 idx[n_close]; i32 n_sgm; i64 idx[n_sgm]; i32 n_stamps; i64 stamps[n_stamps];
//      i32 n_slices; (i32 count, i64 median stamp)[n_slices]
#include <cstdio>
#include <fstream>

#include "esvo_b200/esvo_core.hpp"

int main(int argc, char** argv) {
  if (argc < 3) { std::printf("usage: frontend_dump <in.bin> <out.bin>\n"); return 64; }
  std::ifstream f(argv[1], std::ios::binary);
  int32_t n = 0;
  f.read((char*)&n, 4);
  std::vector<int64_t> t((size_t)n);
  f.read((char*)t.data(), (std::streamsize)n * 8);
  int64_t t_end, t_low, t_up; double half_slice, thickness; int32_t pen;
  f.read((char*)&t_end, 8); f.read((char*)&half_slice, 8); f.read((char*)&pen, 4); f.read((char*)&t_low, 8); f.read((char*)&t_up, 8);
  f.read((char*)&thickness, 8);
  if (!f) return 65;
  std::vector<esvo::Event> ev((size_t)n);
  for (int i = 0; i < n; ++i) ev[(size_t)i] = {(uint16_t)(i % 346), (uint16_t)(i % 260), t[(size_t)i], true};
  using namespace esvo_core::frontend;
  std::ofstream o(argv[2], std::ios::binary);
  auto dump = [&](const std::vector<esvo::Event*>& sel) {
    const int32_t m = (int32_t)sel.size(); o.write((const char*)&m, 4);
    for (auto* e : sel) { const int64_t idx = e - ev.data(); o.write((const char*)&idx, 8); }
  };
  std::vector<esvo::Event*> sel;
  selectCloseEvents(ev, t_end, half_slice, (size_t)pen, sel); dump(sel);
  selectSGMEvents(ev, t_end, half_slice, (size_t)pen, sel); dump(sel);
  const auto st = samplePoseStamps(t_end, half_slice);
  const int32_t ns = (int32_t)st.size(); o.write((const char*)&ns, 4); o.write((const char*)st.data(), (std::streamsize)ns * 8);
  // eventSlicingForEM over the events of [t_low, t_up) as esvo_MVStereo::dataTransferring gathers them (:587-603)
  std::vector<esvo::Event*> window;
  for (auto& e : ev) if (e.ts >= t_low && e.ts < t_up) window.push_back(&e);
  if (!window.empty()) window.pop_back();                                   // ev_left_upBound--: the last event before t_up stays out
  std::vector<esvo_core::core::EventSlice> sl;
  eventSlicingForEM(window, t_low, t_up, thickness, [](int64_t, esvo::Pose& T) { T.fill(0); return true; }, sl);
  const int32_t nsl = (int32_t)sl.size(); o.write((const char*)&nsl, 4);
  for (auto& s : sl) { const int32_t c = (int32_t)s.numEvents_; o.write((const char*)&c, 4); o.write((const char*)&s.t_median_, 8); }
  return 0;
}
