// wfm_calib.cu — in-run calibrations the bench line quotes its rooflines against
// ("calibrate with an FMA microbenchmark in the same run"; VERDICT r1 item 5:
// "in-run pinned cudaMemcpy D2H microbenchmark").  Diagnostics only: nothing on the sampling
// path calls these.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include "../../include/wfm_b200.h"

namespace {

// kChains independent DFMA chains per thread (dependent-issue latency hidden by the chains and
// by the resident warps): what the FP64 pipe issues when nothing else competes for it.
template <int kChains>
__global__ void __launch_bounds__(256) dfma_kernel(double* __restrict__ sink, int iters, double a, double b) {
  double v[kChains];
#pragma unroll
  for (int k = 0; k < kChains; ++k) v[k] = (double)(threadIdx.x + k) * 1e-3;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int k = 0; k < kChains; ++k) v[k] = fma(v[k], a, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < kChains; ++k) s += v[k];
  if (s == 123.456) sink[0] = s;  // never true: keeps the chains alive
}

}  // namespace

extern "C" {

// out[0] = fp64 FMA instructions per second per THREAD-lane summed over the device (i.e. DFMA/s; x2 = FLOP/s),
// out[1] = the kernel's duration in ms.  Best of `reps` launches, CUDA events on `stream`.
int wfm_calibrate_fp64(double* out, int32_t reps, void* stream) {
  if (!out) return WFM_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return WFM_ECUDA;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return WFM_ECUDA;
  double* sink = nullptr;
  if (cudaMalloc(&sink, 64) != cudaSuccess) return WFM_ENOMEM;
  constexpr int kChains = 8;
  const int iters = 2048, blocks = sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < std::max(reps, 1) + 1; ++r) {
    cudaEventRecord(e0, st);
    dfma_kernel<kChains><<<blocks, threads, 0, st>>>(sink, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0) best = std::min(best, ms);  // launch 0 warms up
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  if (cudaGetLastError() != cudaSuccess || best > 1e29f) return WFM_ECUDA;
  const double n = (double)blocks * threads * (double)iters * 8.0 * kChains;
  out[0] = n / (best * 1e-3);
  out[1] = best;
  return WFM_OK;
}

// Pinned-memory copy ceiling of this process's device: out[0] = GB/s of `reps` back-to-back
// cudaMemcpyAsync of `bytes` between a pinned host buffer and device memory (dir 0: device->host,
// 1: host->device), best single copy; out[1] = GB/s over all reps (sustained).  Own buffers.
int wfm_calibrate_copy(int64_t bytes, int32_t dir, int32_t reps, double* out) {
  if (!out || bytes <= 0) return WFM_EINVAL;
  void *h = nullptr, *d = nullptr;
  if (cudaMalloc(&d, (size_t)bytes) != cudaSuccess) return WFM_ENOMEM;
  if (cudaHostAlloc(&h, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) {
    cudaFree(d);
    return WFM_ENOMEM;
  }
  memset(h, 0, (size_t)bytes);
  cudaMemset(d, 0, (size_t)bytes);
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f, total = 0.f;
  const int n = std::max(reps, 1);
  for (int r = 0; r < n + 1; ++r) {
    cudaEventRecord(e0, st);
    if (dir == 0) cudaMemcpyAsync(h, d, (size_t)bytes, cudaMemcpyDeviceToHost, st);
    else cudaMemcpyAsync(d, h, (size_t)bytes, cudaMemcpyHostToDevice, st);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0) {
      best = std::min(best, ms);
      total += ms;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaStreamDestroy(st);
  cudaFreeHost(h);
  cudaFree(d);
  if (cudaGetLastError() != cudaSuccess || best > 1e29f) return WFM_ECUDA;
  out[0] = (double)bytes / (best * 1e-3) / 1e9;
  out[1] = (double)bytes * n / (total * 1e-3) / 1e9;
  return WFM_OK;
}

}  // extern "C"
