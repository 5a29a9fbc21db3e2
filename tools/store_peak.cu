// store_peak.cu — what a pure-store kernel reaches on this GPU (calibrates K1's roofline):
//   (a) st.global.v2.f64 from registers, grid-stride
//   (b) cp.async.bulk shared->global from a CONSTANT shared buffer, one issuing lane per warp
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_peak store_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void st_kernel(double2* out, size_t n2, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const double2 val = make_double2(v, v);
  for (; i < n2; i += stride) out[i] = val;
}

// streaming (evict-first) variant
__global__ void st_cs_kernel(double2* out, size_t n2, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n2; i += stride) __stcs(out + i, make_double2(v, v));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool kHint>
__global__ void bulk_kernel(unsigned char* out, size_t bytes, uint32_t chunk) {
  extern __shared__ __align__(128) unsigned char buf[];
  for (uint32_t i = threadIdx.x * 16; i < chunk; i += blockDim.x * 16) *reinterpret_cast<uint4*>(buf + i) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const size_t n_warps = (size_t)gridDim.x * (blockDim.x / 32);
  const size_t n_chunks = bytes / chunk;
  if (lane == 0) {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    for (size_t c = warp; c < n_chunks; c += n_warps) {
      if (kHint)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(out + c * chunk), "r"(smem_u32(buf)), "r"(chunk), "l"(policy) : "memory");
      else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * chunk), "r"(smem_u32(buf)), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      // bound the number of outstanding groups per thread
      asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <typename F>
float time_it(F f, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main() {
  const size_t bytes = (size_t)4 << 30;
  unsigned char* out;
  CK(cudaMalloc(&out, bytes));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("SMs %d, buffer %.2f GB\n", sms, bytes / 1e9);
  {
    float ms = time_it([&] { CK(cudaMemsetAsync(out, 0, bytes)); }, 10);
    printf("cudaMemset                         %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
  }
  for (int per_sm : {2, 4, 8}) for (int threads : {256, 512}) {
    float ms = time_it([&] { st_kernel<<<sms * per_sm, threads>>>((double2*)out, bytes / 16, 0.0); }, 10);
    printf("st.v2.f64   grid %2dxSM x %3d       %8.3f ms  %8.1f GB/s\n", per_sm, threads, ms, bytes / ms / 1e6);
  }
  {
    float ms = time_it([&] { st_cs_kernel<<<sms * 8, 256>>>((double2*)out, bytes / 16, 0.0); }, 10);
    printf("st.cs.v2.f64 grid 8xSM x 256       %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
  }
  {
    float ms = time_it([&] { st_kernel<<<(unsigned)(bytes / 16 / 256), 256>>>((double2*)out, bytes / 16, 0.0); }, 10);
    printf("st.v2.f64   one elem per thread    %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
  }
  for (uint32_t chunk : {1024u, 2048u, 4096u, 8192u, 16384u, 32768u}) for (int warps : {1, 4, 8}) for (int per_sm : {1, 2}) {
    CK(cudaFuncSetAttribute(bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    float ms = time_it([&] { bulk_kernel<true><<<sms * per_sm, warps * 32, chunk>>>(out, bytes, chunk); }, 10);
    float ms2 = time_it([&] { bulk_kernel<false><<<sms * per_sm, warps * 32, chunk>>>(out, bytes, chunk); }, 10);
    printf("bulk s2g chunk %5u B, %d warps x %d CTA/SM: evict_first %8.3f ms %8.1f GB/s | default %8.3f ms %8.1f GB/s\n", chunk, warps, per_sm, ms,
           bytes / ms / 1e6, ms2, bytes / ms2 / 1e6);
  }
  return 0;
}
