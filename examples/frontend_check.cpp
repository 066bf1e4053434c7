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
  auto st = samplePoseStamps(t_end, 0.001);
  if (st.size() != 201) bad |= 16;                                             // 10 ms window / 50 us
  if (!st.empty() && (st.front() != t_end - 10000000 || st.back() > t_end)) bad |= 32;
  if (fromSec(toSec(1234567891234567890LL)) / 1000 != 1234567891234567890LL / 1000) bad |= 64;
  std::printf("frontend check %s (flags %d): %zu events, %zu stamps\n", bad ? "FAILED" : "ok", bad, sel.size(), st.size());
  return bad;
}
