// Self-check of the host-side event front-end helpers (esvo_core::frontend) -- no GPU needed.
#include <cstdio>
#include "esvo_b200/esvo_core.hpp"
int main() {
  using namespace esvo_core::frontend;
  std::vector<esvo::Event> ev;
  for (int i = 0; i < 20000; ++i) ev.push_back({(uint16_t)(i % 346), (uint16_t)(i % 260), 1000000000LL + (int64_t)i * 1000, true});  // 1 us apart
  std::vector<esvo::Event*> sel;
  const int64_t t_end = 1000000000LL + 15000 * 1000;       // an event sits exactly at t_end
  selectCloseEvents(ev, t_end, 0.001, 1000, sel);
  int bad = 0;
  if (sel.size() != 1000) bad |= 1;
  if (!sel.empty() && sel.front()->ts != t_end) bad |= 2;                       // starts at lower_bound(t_end)
  for (size_t i = 1; i < sel.size(); ++i) if (sel[i]->ts >= sel[i - 1]->ts) bad |= 4;   // newest first
  selectCloseEvents(ev, t_end, 0.00001, 1000, sel);                             // window of 100 us -> 100 events
  if (sel.size() != 100) bad |= 8;
  // observation stamp newer than every buffered event: the newest events are still selected (one budget slot goes to
  // the reference's one-past-the-end read)
  {
    std::vector<esvo::Event*> s2;
    const int64_t t_new = ev.back().ts + 5000;
    selectCloseEvents(ev, t_new, 0.001, 1000, s2);
    if (s2.size() != 999 || s2.front() != &ev.back()) bad |= 4096;
    selectCloseEvents(ev, t_new, 0.0000001, 1000, s2);            // 1 us window: lower_bound(t_begin) == end as well -> nothing
    if (!s2.empty()) bad |= 8192;
    selectSGMEvents(ev, t_new, 0.001, 1000, s2);                  // 2 ms window = 2000 events available, budget 1001 - 1
    if (s2.size() != 1000 || s2.front() != &ev.back()) bad |= 16384;
  }
  auto st = samplePoseStamps(t_end, 0.001);
  if (st.size() != 201) bad |= 16;                                             // 10 ms window / 50 us
  if (!st.empty() && (st.front() != t_end - 10000000 || st.back() > t_end)) bad |= 32;
  if (fromSec(toSec(1234567891234567890LL)) / 1000 != 1234567891234567890LL / 1000) bad |= 64;
  // denoising: a 3x3 block of events survives entirely only in its centre; an isolated event never does
  {
    std::vector<esvo::Event> ev2; std::vector<esvo::Event*> all, close, kept;
    for (int dy = 0; dy < 3; ++dy) for (int dx = 0; dx < 3; ++dx) ev2.push_back({(uint16_t)(50 + dx), (uint16_t)(40 + dy), 1, true});
    ev2.push_back({200, 100, 2, true});                       // isolated
    for (int dx = 0; dx < 3; ++dx) for (int dy = 0; dy < 2; ++dy) ev2.push_back({(uint16_t)dx, (uint16_t)dy, 3, true});  // 3x2 block at the corner (borders replicate)
    for (auto& e : ev2) { all.push_back(&e); close.push_back(&e); }
    std::vector<uint8_t> mask;
    createDenoisingMask(all, mask, 260, 346);
    if (mask[41 * 346 + 51] != 255 || mask[100 * 346 + 200] != 0) bad |= 128;
    if (mask[0] != 255) bad |= 256;                            // corner pixel: replicated border makes 6 of 9 neighbours set
    int on = 0; for (uint8_t m : mask) on += m == 255;
    extractDenoisedEvents(close, kept, mask, 346, 4);
    if (kept.size() != 4) bad |= 512;                          // maxNum honoured, arrival order kept
    for (auto* e : kept) if (mask[(size_t)e->y * 346 + e->x] != 255) bad |= 1024;
    std::vector<esvo::DepthPoint> el(2); el[0].p_cam[0] = 1; el[0].p_cam[1] = 2; el[0].p_cam[2] = 3; el[1].p_cam[0] = 0; el[1].p_cam[1] = 0; el[1].p_cam[2] = 10;
    esvo::Pose T = {0, -1, 0, 5, 1, 0, 0, 6, 0, 0, 1, 7, 0, 0, 0, 1};
    std::vector<float> xyz, nearp;
    packPointCloud(el, T, xyz, &nearp, 5.0);
    if (xyz.size() != 6 || xyz[0] != 3.f || xyz[1] != 7.f || xyz[2] != 10.f || nearp.size() != 3) bad |= 2048;
    (void)on;
  }
  // eventSlicingForEM: 1 us apart events, 1 ms slices over [ev[0], ev[0] + 5.5 ms): 5 slices of 1001 events (the bound event
  // belongs to the slice), contiguous, median stamp = the 500th event of the slice
  {
    std::vector<esvo::Event*> ptrs; for (auto& e : ev) ptrs.push_back(&e);
    std::vector<esvo_core::core::EventSlice> sl;
    int calls = 0;
    eventSlicingForEM(ptrs, ev[0].ts, ev[0].ts + 5500000, 1e-3, [&](int64_t, esvo::Pose& T) { ++calls; T.fill(0); return true; }, sl);
    if (sl.size() != 5 || calls != 5) bad |= 32768;
    size_t at = 0;
    for (auto& e : sl) {
      if (e.numEvents_ != 1001 || (size_t)(e.it_begin_ - ptrs.begin()) != at || e.t_median_ != ev[at + 500].ts) bad |= 65536;
      at += e.numEvents_;
    }
  }
  std::printf("frontend check %s (flags %d): %zu events, %zu stamps\n", bad ? "FAILED" : "ok", bad, sel.size(), st.size());
  return bad;
}
