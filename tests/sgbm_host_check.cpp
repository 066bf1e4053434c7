// Runs the per-work-item SGBM functions of esvo_b200/csrc/sgbm_core.h (the ones sgbm.cu launches as CUDA threads) in
// plain loops on the host, so that their integer arithmetic can be pinned against cv2 / oracle/sgbm.py without a GPU.
// usage: sgbm_host_check <in.bin> <out.bin>
//   in : int32 W, H, numDisparities, blockSize, P1, P2, disp12MaxDiff, preFilterCap, uniquenessRatio; then left, right (H*W u8 each)
//   out: H*W int16 (CV_16S disparity * 16)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../esvo_b200/csrc/sgbm_core.h"

using namespace esvo_sgbm;

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  int32_t h[9];
  if (std::fread(h, 4, 9, f) != 9) return 4;
  const int W = h[0], H = h[1];
  std::vector<uint8_t> L((size_t)W * H), R((size_t)W * H);
  if (std::fread(L.data(), 1, L.size(), f) != L.size() || std::fread(R.data(), 1, R.size(), f) != R.size()) return 5;
  std::fclose(f);
  const Dims dm = make_dims(W, H, h[2], h[3], h[4], h[5], h[6], h[7], h[8]);
  if (dm.W1 <= 0 || dm.D > kMaxD) return 6;
  const size_t vol = (size_t)H * dm.W1 * dm.D;
  std::vector<int16_t> pl((size_t)H * 6 * W), pr((size_t)H * 6 * W), pix(vol), hs(vol), C(vol), L0(vol), L1(vol), L2(vol), L3(vol), Lr(vol);
  for (int y = 0; y < H; ++y) { prep_row(L.data(), W, dm, y, &pl[(size_t)y * 6 * W]); prep_row(R.data(), W, dm, y, &pr[(size_t)y * 6 * W]); }
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < dm.W1; ++x) pixel_cost(&pl[(size_t)y * 6 * W], &pr[(size_t)y * 6 * W], dm, x, &pix[((size_t)y * dm.W1 + x) * dm.D]);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < dm.W1; ++x) box_h(&pix[(size_t)y * dm.W1 * dm.D], dm, x, &hs[((size_t)y * dm.W1 + x) * dm.D]);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < dm.W1; ++x) box_v(hs.data(), dm, y, x, &C[((size_t)y * dm.W1 + x) * dm.D]);
  // the five path families (same start points as the kernels in sgbm.cu)
  for (int y = 0; y < H; ++y) { walk_path(C.data(), L0.data(), dm, 0, y, 1, 0); walk_path(C.data(), Lr.data(), dm, dm.W1 - 1, y, -1, 0); }
  for (int x = 0; x < dm.W1; ++x) { walk_path(C.data(), L2.data(), dm, x, 0, 0, 1); walk_path(C.data(), L1.data(), dm, x, 0, 1, 1); walk_path(C.data(), L3.data(), dm, x, 0, -1, 1); }
  for (int y = 1; y < H; ++y) { walk_path(C.data(), L1.data(), dm, 0, y, 1, 1); walk_path(C.data(), L3.data(), dm, dm.W1 - 1, y, -1, 1); }
  std::vector<int16_t> raw((size_t)W * H), out((size_t)W * H);
  std::vector<int32_t> scratch(2 * (size_t)W);
  for (int y = 0; y < H; ++y) select_row(L0.data(), L1.data(), L2.data(), L3.data(), Lr.data(), dm, y, scratch.data(), &raw[(size_t)y * W]);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) out[(size_t)y * W + x] = median3(raw.data(), dm, y, x);
  f = std::fopen(argv[2], "wb");
  if (!f) return 7;
  std::fwrite(out.data(), 2, out.size(), f);
  std::fclose(f);
  return 0;
}
