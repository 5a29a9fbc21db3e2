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

// NON-persistent bulk stores: every warp stores `per_warp` consecutive chunks and exits
// (grid = n_chunks / (warps * per_warp)): does a huge short-lived grid reach what memset reaches?
__global__ void bulk_once_kernel(unsigned char* out, size_t bytes, uint32_t chunk, int per_warp) {
  extern __shared__ __align__(128) unsigned char buf[];
  for (uint32_t i = threadIdx.x * 16; i < chunk; i += blockDim.x * 16) *reinterpret_cast<uint4*>(buf + i) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const size_t n_chunks = bytes / chunk;
  if (lane == 0) {
    for (int k = 0; k < per_warp; ++k) {
      const size_t c = warp * per_warp + k;
      if (c >= n_chunks) break;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * chunk), "r"(smem_u32(buf)), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}
// persistent, but every CTA owns a CONTIGUOUS range of the buffer (blocked instead of round-robin)
__global__ void bulk_blocked_kernel(unsigned char* out, size_t bytes, uint32_t chunk) {
  extern __shared__ __align__(128) unsigned char buf[];
  for (uint32_t i = threadIdx.x * 16; i < chunk; i += blockDim.x * 16) *reinterpret_cast<uint4*>(buf + i) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int lane = threadIdx.x & 31, wpc = blockDim.x / 32;
  const size_t n_chunks = bytes / chunk;
  const size_t per_cta = (n_chunks + gridDim.x - 1) / gridDim.x;
  const size_t lo = blockIdx.x * per_cta, hi = lo + per_cta < n_chunks ? lo + per_cta : n_chunks;
  if (lane == 0) {
    for (size_t c = lo + (threadIdx.x >> 5); c < hi; c += wpc) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * chunk), "r"(smem_u32(buf)), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
// persistent grid, chunks handed out DYNAMICALLY (atomic counter, one fetch per chunk per warp)
__global__ void bulk_dynamic_kernel(unsigned char* out, size_t bytes, uint32_t chunk, unsigned long long* counter, int batch) {
  extern __shared__ __align__(128) unsigned char buf[];
  for (uint32_t i = threadIdx.x * 16; i < chunk; i += blockDim.x * 16) *reinterpret_cast<uint4*>(buf + i) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t n_chunks = bytes / chunk;
  if (lane == 0) {
    for (;;) {
      const size_t c0 = atomicAdd(counter, (unsigned long long)batch);
      if (c0 >= n_chunks) break;
      for (size_t c = c0; c < c0 + batch && c < n_chunks; ++c) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * chunk), "r"(smem_u32(buf)), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
// registers, non-persistent: every thread stores `per_thread` 16-byte elements (strided by the CTA) and exits
__global__ void st_few_kernel(double2* out, size_t n2, double v, int per_thread) {
  const size_t base = (size_t)blockIdx.x * blockDim.x * per_thread + threadIdx.x;
  const double2 val = make_double2(v, v);
  for (int k = 0; k < per_thread; ++k) {
    const size_t i = base + (size_t)k * blockDim.x;
    if (i < n2) out[i] = val;
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
  for (int per_thread : {2, 4, 16, 64}) {
    const size_t n2 = bytes / 16;
    float ms = time_it([&] { st_few_kernel<<<(unsigned)((n2 + 256ull * per_thread - 1) / (256ull * per_thread)), 256>>>((double2*)out, n2, 0.0, per_thread); }, 10);
    printf("st.v2.f64   %2d elems per thread, non-persistent  %8.3f ms  %8.1f GB/s\n", per_thread, ms, bytes / ms / 1e6);
  }
  CK(cudaFuncSetAttribute(bulk_once_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(bulk_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  for (uint32_t chunk : {4096u, 10240u, 16384u}) for (int warps : {4, 8}) for (int per_warp : {1, 2, 4, 16}) {
    const size_t n_chunks = bytes / chunk;
    const unsigned grid = (unsigned)((n_chunks + (size_t)warps * per_warp - 1) / ((size_t)warps * per_warp));
    float ms = time_it([&] { bulk_once_kernel<<<grid, warps * 32, chunk>>>(out, bytes, chunk, per_warp); }, 10);
    printf("bulk s2g NON-persistent chunk %5u B, %d warps/CTA, %2d chunks/warp (grid %7u)  %8.3f ms  %8.1f GB/s\n", chunk, warps, per_warp, grid, ms, bytes / ms / 1e6);
  }
  for (uint32_t chunk : {4096u, 10240u}) for (int per_sm : {1, 2}) {
    float ms = time_it([&] { bulk_blocked_kernel<<<sms * per_sm, 256, chunk>>>(out, bytes, chunk); }, 10);
    printf("bulk s2g persistent BLOCKED ranges chunk %5u B, 8 warps x %d CTA/SM  %8.3f ms  %8.1f GB/s\n", chunk, per_sm, ms, bytes / ms / 1e6);
  }
  {
    unsigned long long* counter;
    CK(cudaMalloc(&counter, 8));
    CK(cudaFuncSetAttribute(bulk_dynamic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (uint32_t chunk : {10240u}) for (int per_sm : {2}) for (int batch : {1, 4, 8, 16, 32}) {
      float ms = time_it([&] { CK(cudaMemsetAsync(counter, 0, 8)); bulk_dynamic_kernel<<<sms * per_sm, 256, chunk>>>(out, bytes, chunk, counter, batch); }, 10);
      printf("bulk s2g persistent DYNAMIC (atomic counter, %2d chunks per fetch) chunk %5u B, 8 warps x %d CTA/SM  %8.3f ms  %8.1f GB/s\n", batch, chunk, per_sm, ms, bytes / ms / 1e6);
    }
  }
  for (uint32_t chunk : {8192u}) for (int warps : {8}) for (int per_sm : {1, 2}) {
    CK(cudaFuncSetAttribute(bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    float ms = time_it([&] { bulk_kernel<true><<<sms * per_sm, warps * 32, chunk>>>(out, bytes, chunk); }, 10);
    float ms2 = time_it([&] { bulk_kernel<false><<<sms * per_sm, warps * 32, chunk>>>(out, bytes, chunk); }, 10);
    printf("bulk s2g chunk %5u B, %d warps x %d CTA/SM: evict_first %8.3f ms %8.1f GB/s | default %8.3f ms %8.1f GB/s\n", chunk, warps, per_sm, ms,
           bytes / ms / 1e6, ms2, bytes / ms2 / 1e6);
  }
  return 0;
}
