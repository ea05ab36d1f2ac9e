// Micro-benchmark: throughput of the ways a CTA can add an fp32 tile into global memory on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bw red_bw.cu && ./red_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: cp.reduce.async.bulk add.f32 (chunk bytes from smem), mode 1: cp.async.bulk store
template <int MODE>
__global__ void k_bulk(float* dst, size_t region_floats, int chunk_bytes, int iters, int shared_region) {
  extern __shared__ __align__(128) uint8_t sm[];
  for (int i = threadIdx.x; i < chunk_bytes / 4; i += blockDim.x) ((float*)sm)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    size_t chunk_f = chunk_bytes / 4;
    size_t nchunks = region_floats / chunk_f;
    size_t c = shared_region ? (size_t)blockIdx.x * 7 : (size_t)blockIdx.x * (nchunks / gridDim.x);
    for (int it = 0; it < iters; ++it) {
      float* p = dst + (c % nchunks) * chunk_f;
      if (MODE == 0)
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(p), "r"(smem_u32(sm)), "r"(chunk_bytes) : "memory");
      else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p), "r"(smem_u32(sm)), "r"(chunk_bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if ((it & 3) == 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
      c += shared_region ? 1 : 1;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// per-thread red.global.add: mode 0 = v4 with 1 KB row stride per lane (current kernel's pattern),
// mode 1 = v4, lanes contiguous (512 B per warp instr), mode 2 = scalar f32 coalesced
template <int MODE>
__global__ void k_red(float* dst, size_t region_floats, int iters) {
  // each CTA handles tiles of 128 rows x 256 floats (128 KB); 128 threads
  size_t tile_f = 128 * 256;
  size_t ntiles = region_floats / tile_f;
  size_t t = (size_t)blockIdx.x * (ntiles / gridDim.x);
  for (int it = 0; it < iters; ++it) {
    float* base = dst + (t % ntiles) * tile_f;
    if (MODE == 0) {
      float* p = base + (size_t)threadIdx.x * 256;
#pragma unroll 8
      for (int c = 0; c < 256; c += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p + c), "f"(1.0f) : "memory");
    } else if (MODE == 1) {
#pragma unroll 8
      for (int c = 0; c < 64; ++c)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(base + ((size_t)c * 128 + threadIdx.x) * 4), "f"(1.0f) : "memory");
    } else {
#pragma unroll 8
      for (int c = 0; c < 256; ++c)
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(base + (size_t)c * 128 + threadIdx.x), "f"(1.0f) : "memory");
    }
    ++t;
  }
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  size_t region = 64ull << 20;  // floats = 256 MB
  float* d; CK(cudaMalloc(&d, region * 4)); CK(cudaMemset(d, 0, region * 4));
  int sms = 148;
  for (int chunk : {4096, 16384, 32768, 65536}) {
    int iters = 2000;
    CK(cudaFuncSetAttribute(k_bulk<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(k_bulk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int shared : {0, 1}) {
      float ms0 = time_ms([&] { k_bulk<0><<<sms, 128, chunk>>>(d, shared ? (8u << 20) : region, chunk, iters, shared); });
      float ms1 = time_ms([&] { k_bulk<1><<<sms, 128, chunk>>>(d, shared ? (8u << 20) : region, chunk, iters, shared); });
      double bytes = (double)sms * iters * chunk;
      printf("bulk chunk=%6d B %s: reduce.add.f32 %7.1f GB/s (%5.1f B/clk/SM @1.9GHz)   store %7.1f GB/s\n", chunk,
             shared ? "L2-resident 32MB region, CTAs interleaved" : "256MB region, disjoint per CTA      ",
             bytes / ms0 * 1e-6, bytes / ms0 * 1e-6 / 148 / 1.9, bytes / ms1 * 1e-6);
    }
  }
  const char* names[3] = {"red.v4 row-stride 1KB/lane", "red.v4 lanes contiguous   ", "red.f32 coalesced         "};
  for (int mode = 0; mode < 3; ++mode) {
    int iters = 100;
    float ms = 0;
    if (mode == 0) ms = time_ms([&] { k_red<0><<<sms, 128>>>(d, region, iters); });
    if (mode == 1) ms = time_ms([&] { k_red<1><<<sms, 128>>>(d, region, iters); });
    if (mode == 2) ms = time_ms([&] { k_red<2><<<sms, 128>>>(d, region, iters); });
    double bytes = (double)sms * iters * 128 * 1024;
    printf("%s: %7.1f GB/s (%5.1f B/clk/SM @1.9GHz)\n", names[mode], bytes / ms * 1e-6, bytes / ms * 1e-6 / 148 / 1.9);
  }
  // small region (L2 resident, 16 MB) for the per-thread variants
  for (int mode = 0; mode < 3; ++mode) {
    int iters = 100;
    size_t reg = 8u << 20;  // 32 MB
    float ms = 0;
    if (mode == 0) ms = time_ms([&] { k_red<0><<<sms, 128>>>(d, reg, iters); });
    if (mode == 1) ms = time_ms([&] { k_red<1><<<sms, 128>>>(d, reg, iters); });
    if (mode == 2) ms = time_ms([&] { k_red<2><<<sms, 128>>>(d, reg, iters); });
    double bytes = (double)sms * iters * 128 * 1024;
    printf("[32MB region] %s: %7.1f GB/s (%5.1f B/clk/SM)\n", names[mode], bytes / ms * 1e-6, bytes / ms * 1e-6 / 148 / 1.9);
  }
  return 0;
}
