// Memory-access pattern of the K3 column passes without the transform: how long do the loads and stores alone take?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/fft_pattern_probe tools/probes/fft_pattern_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
template <int LOGC>
__global__ void __launch_bounds__(512) pass_a(const double* __restrict__ x, double2* __restrict__ s, int N1, int N2, int64_t n) {
  const int C = 1 << LOGC, c0 = blockIdx.x << LOGC;
  const double* pa = x + 2 * (int64_t)blockIdx.y * n;
  const double* pb = pa + n;
  double2* out = s + (int64_t)blockIdx.y * n;
  const int nb = N1 / 5;
  for (int jj = threadIdx.x; jj < nb * C; jj += blockDim.x) {
    const int c = jj & (C - 1), j = jj >> LOGC;
    double2 v[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      const int64_t idx = (int64_t)(j + r * nb) * N2 + c0 + c;
      v[r] = make_double2(pa[idx], pb[idx]);
    }
#pragma unroll
    for (int r = 0; r < 5; ++r) out[(int64_t)(j + r * nb) * N2 + c0 + c] = v[r];
  }
}
template <int LOGC>
__global__ void __launch_bounds__(512) pass_c(const double2* __restrict__ s, double* __restrict__ y, int N1, int N2, int64_t n) {
  const int C = 1 << LOGC, c0 = blockIdx.x << LOGC;
  double* pa = y + 2 * (int64_t)blockIdx.y * n;
  double* pb = pa + n;
  const double2* in = s + (int64_t)blockIdx.y * n;
  const int nb = N1 / 5;
  for (int jj = threadIdx.x; jj < nb * C; jj += blockDim.x) {
    const int c = jj & (C - 1), j = jj >> LOGC;
    double2 v[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) v[r] = in[(int64_t)(j + r * nb) * N2 + c0 + c];
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      const int64_t idx = (int64_t)(j + r * nb) * N2 + c0 + c;
      pa[idx] = v[r].x;
      pb[idx] = v[r].y;
    }
  }
}
__global__ void rows_copy(double2* __restrict__ s, int64_t total) {  // the row pass: contiguous read-modify-write
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    double2 v = s[i];
    v.x += 1.0;
    s[i] = v;
  }
}
template <class F>
static float time_it(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < 10; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / 10;
}
int main() {
  const int N1 = 625, N2 = 640, pairs = 128;
  const int64_t n = (int64_t)N1 * N2;
  double *x, *y; double2* s;
  cudaMalloc(&x, sizeof(double) * n * 2 * pairs); cudaMalloc(&y, sizeof(double) * n * 2 * pairs); cudaMalloc(&s, sizeof(double2) * n * pairs);
  cudaMemset(x, 0, sizeof(double) * n * 2 * pairs); cudaMemset(s, 0, sizeof(double2) * n * pairs);
  printf("pass A pattern  C=4: %.3f ms   C=8: %.3f ms   C=16: %.3f ms\n",
         time_it([&] { pass_a<2><<<dim3(N2 / 4, pairs), 512>>>(x, s, N1, N2, n); }),
         time_it([&] { pass_a<3><<<dim3(N2 / 8, pairs), 512>>>(x, s, N1, N2, n); }),
         time_it([&] { pass_a<4><<<dim3(N2 / 16, pairs), 512>>>(x, s, N1, N2, n); }));
  printf("pass C pattern  C=4: %.3f ms   C=8: %.3f ms   C=16: %.3f ms\n",
         time_it([&] { pass_c<2><<<dim3(N2 / 4, pairs), 512>>>(s, y, N1, N2, n); }),
         time_it([&] { pass_c<3><<<dim3(N2 / 8, pairs), 512>>>(s, y, N1, N2, n); }),
         time_it([&] { pass_c<4><<<dim3(N2 / 16, pairs), 512>>>(s, y, N1, N2, n); }));
  printf("row pass pattern (contiguous r/w of the scratch): %.3f ms\n", time_it([&] { rows_copy<<<148 * 8, 512>>>(s, n * pairs); }));
  printf("(bytes per pass: %.2f GB)\n", (double)(n * pairs * 32) / 1e9);
  return 0;
}
