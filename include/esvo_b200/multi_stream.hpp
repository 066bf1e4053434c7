// esvo_b200 -- multi-GPU plumbing for the C++ host: independent stereo event streams are sharded one per
// GPU, there is no collective inside the hot path; the only communication is one ncclAllGather of a fixed record per stream
// per reporting interval {stream_id, frames, n_seeds, n_solved, n_culled, n_fusions, bm_evals, lm_evals, map_size, checksum}
// -- the C++ counterpart of esvo_b200/dist.py (same record layout, same order-sensitive map checksum).
// Header-only; needs <nccl.h> and the CUDA runtime in addition to libesvo_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <vector>

#include "esvo_core.hpp"

namespace esvo {

constexpr int kStreamRecordFields = 10;

// sum_i (i+1) * (row_i*4096 + col_i + rho_i) over the map in element order (esvo_b200/dist.py::map_checksum)
inline double mapChecksum(const std::vector<DepthPoint>& elems) {
  double s = 0;
  for (size_t i = 0; i < elems.size(); ++i) s += (double)(i + 1) * (elems[i].row * 4096.0 + elems[i].col + elems[i].inv_depth);
  return s;
}
inline void makeStreamRecord(int stream_id, int frames, const esvo_core::esvo_Mapping::Counters& c, double checksum, double rec[kStreamRecordFields]) {
  const double v[kStreamRecordFields] = {(double)stream_id, (double)frames, (double)c.n_seeds, (double)c.n_solved, (double)c.n_culled, (double)c.n_fusions,
                                         (double)c.bm_evals, (double)c.lm_evals, (double)c.map_size, checksum};
  for (int i = 0; i < kStreamRecordFields; ++i) rec[i] = v[i];
}
// One all-gather of the stream records over the communicator; every rank receives world x kStreamRecordFields doubles.
// Call from the thread / process that owns `device`; false on a CUDA or NCCL error.
inline bool gatherStreamRecords(ncclComm_t comm, int world, int device, const double rec[kStreamRecordFields], std::vector<double>& all) {
  if (cudaSetDevice(device) != cudaSuccess) return false;
  cudaStream_t st = nullptr; double *d_send = nullptr, *d_recv = nullptr;
  bool ok = cudaStreamCreate(&st) == cudaSuccess && cudaMalloc(&d_send, kStreamRecordFields * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&d_recv, (size_t)world * kStreamRecordFields * sizeof(double)) == cudaSuccess;
  ok = ok && cudaMemcpyAsync(d_send, rec, kStreamRecordFields * sizeof(double), cudaMemcpyHostToDevice, st) == cudaSuccess;
  ok = ok && ncclAllGather(d_send, d_recv, kStreamRecordFields, ncclDouble, comm, st) == ncclSuccess;
  all.assign((size_t)world * kStreamRecordFields, 0.0);
  ok = ok && cudaMemcpyAsync(all.data(), d_recv, all.size() * sizeof(double), cudaMemcpyDeviceToHost, st) == cudaSuccess;
  ok = ok && cudaStreamSynchronize(st) == cudaSuccess;
  if (d_send) cudaFree(d_send);
  if (d_recv) cudaFree(d_recv);
  if (st) cudaStreamDestroy(st);
  return ok;
}

}  // namespace esvo
