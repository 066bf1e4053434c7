// Minimal TMA 2-D u8 box load test (debug aid for bm_tma_kernel): ./tma_min <mode 0 param desc | 1 global desc> <var> <bw 48|64>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
template <int BW, int MODE, int VAR>
__global__ void k(const __grid_constant__ CUtensorMap tm, const CUtensorMap* tm_g, int x0, int y0, uint8_t* out) {
  __shared__ __align__(128) uint8_t s[8 * BW];
  __shared__ __align__(8) unsigned long long bar;
  const unsigned b = (unsigned)__cvta_generic_to_shared(&bar), d = (unsigned)__cvta_generic_to_shared(s);
  const unsigned long long desc = MODE == 0 ? reinterpret_cast<unsigned long long>(&tm) : reinterpret_cast<unsigned long long>(tm_g);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    if (VAR == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (VAR == 1) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((unsigned)(BW * 7)) : "memory");
    if (VAR == 4)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                   ::"r"(d), "l"(desc), "r"(b), "r"(x0), "r"(y0), "l"(0x1000000000000000ULL) : "memory");
    else if (VAR == 5)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(d), "l"(desc), "r"(b), "r"(x0), "r"(y0) : "memory");
    else if (VAR != 3)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(d), "l"(desc), "r"(b), "r"(x0), "r"(y0) : "memory");
    else   // no TMA at all: complete the transaction by hand (tests the mbarrier half alone)
      asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(b), "r"((unsigned)(BW * 7)) : "memory");
  }
  __syncwarp();
  unsigned done = 0, par = 0;
  while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(done) : "r"(b), "r"(par) : "memory");
  for (int i = threadIdx.x; i < BW * 7; i += 32) out[i] = s[i];
}
typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int BW> int run(int mode, int var) {
  const int W = 346, H = 260, P = 352;
  std::vector<uint8_t> img(P * H);
  for (int y = 0; y < H; ++y) for (int x = 0; x < P; ++x) img[y * P + x] = (uint8_t)((x * 7 + y * 13) & 255);
  uint8_t *d_img, *d_out; cudaMalloc(&d_img, P * H); cudaMalloc(&d_out, BW * 8);
  cudaMemcpy(d_img, img.data(), P * H, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  enc_t enc = (enc_t)fn;
  alignas(64) CUtensorMap tm;
  cuuint64_t gd[2] = {(cuuint64_t)(getenv("TMA_W352") ? P : W), (cuuint64_t)H}, gs[1] = {(cuuint64_t)P};
  { const unsigned long long* q = (const unsigned long long*)&tm; (void)q; }
  cuuint32_t box[2] = {BW, 7}, es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_img, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc %d (bw %d mode %d var %d) desc words:", (int)r, BW, mode, var);
  for (int i = 0; i < 8; ++i) printf(" %016llx", ((const unsigned long long*)&tm)[i]);
  printf("\n");
  CUtensorMap* tm_g; cudaMalloc(&tm_g, sizeof(tm)); cudaMemcpy(tm_g, &tm, sizeof(tm), cudaMemcpyHostToDevice);
  for (int x0 : {100, -5, 320}) {
    cudaMemset(d_out, 0xee, BW * 8);
    if (mode == 0 && var == 0) k<BW, 0, 0><<<1, 32>>>(tm, tm_g, x0, 50, d_out);
    if (mode == 1 && var == 0) k<BW, 1, 0><<<1, 32>>>(tm, tm_g, x0, 50, d_out);
    if (mode == 0 && var == 1) k<BW, 0, 1><<<1, 32>>>(tm, tm_g, x0, 50, d_out);
    if (mode == 1 && var == 1) k<BW, 1, 1><<<1, 32>>>(tm, tm_g, x0, 50, d_out);
    if (mode == 0 && var == 2) k<BW, 0, 2><<<1, 32>>>(tm, tm_g, x0, 50, d_out);
    if (mode == 0 && var == 3) k<BW, 0, 3><<<1, 32>>>(tm, tm_g, x0, 50, d_out);
    if (mode == 0 && var == 4) k<BW, 0, 4><<<1, 32>>>(tm, tm_g, x0, 50, d_out);
    if (mode == 0 && var == 5) k<BW, 0, 5><<<1, 32>>>(tm, tm_g, x0, 50, d_out);
    if (mode == 0 && var == 6) k<BW, 0, 0><<<1, 32>>>(tm, tm_g, 96, 48, d_out);   // 16-byte aligned coordinates
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint8_t> o(BW * 7); cudaMemcpy(o.data(), d_out, BW * 7, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < 7; ++y) for (int x = 0; x < BW; ++x) { int gx = x0 + x; uint8_t want = (gx >= 0 && gx < W) ? img[(50 + y) * P + gx] : 0; bad += o[y * BW + x] != want; }
    printf("  x0 %d: %s, mismatches %d\n", x0, cudaGetErrorString(e), bad);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0, var = argc > 2 ? atoi(argv[2]) : 0, bw = argc > 3 ? atoi(argv[3]) : 48;
  return bw == 64 ? run<64>(mode, var) : run<48>(mode, var);
}
