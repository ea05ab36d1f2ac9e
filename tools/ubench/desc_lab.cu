// Descriptor lab: checks the UMMA shared-memory descriptor encodings (K-major / MN-major, 128-byte and
// 32-byte swizzle) and the TMA SWIZZLE_32B layout used by the region-word backward against a CPU matmul.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o desc_lab desc_lab.cu -lcuda && ./desc_lab
// Host code builds the byte image of every operand tile and the per-k-step descriptors; the kernel only
// copies the image into shared memory, issues the MMAs and dumps the accumulator.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../demovlp_b200/csrc/tc_ptx.cuh"
using namespace dvlp::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int kSmem = 200 * 1024;
struct Job {
  int nk, N;
  uint32_t idesc;
  uint64_t da[16], db[16];   // descriptors with the address field relative to the smem base
};

__global__ void __launch_bounds__(128, 1) k_lab(const uint8_t* image, int image_bytes, Job job, float* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar_mem;
  __shared__ uint32_t tmem_slot;
  const uint32_t sbase = smem_u32(sm);
  for (int i = threadIdx.x * 16; i < image_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(sm + i) = *reinterpret_cast<const uint4*>(image + i);
  fence_proxy_async();
  const uint32_t bar = smem_u32(&bar_mem);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_slot), 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    for (int k = 0; k < job.nk; ++k)
      mma_bf16(tmem, job.da[k] + (uint64_t)((sbase >> 4) & 0x3FFF), job.db[k] + (uint64_t)((sbase >> 4) & 0x3FFF), job.idesc, k != 0);
    tc_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = warp * 32 + lane;
  for (int c = 0; c < job.N; c += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, reinterpret_cast<uint32_t*>(v));
    tmem_ld_wait();
    for (int x = 0; x < 16; ++x) out[row * 256 + c + x] = v[x];
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

// TMA SWIZZLE_32B round trip: global [128][80] -> smem (5 boxes) -> dump image; smem -> global copy
__global__ void __launch_bounds__(128, 1) k_tma(const __grid_constant__ CUtensorMap in_map,
                                                const __grid_constant__ CUtensorMap out_map, uint8_t* image) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar_mem;
  const uint32_t sbase = smem_u32(sm), bar = smem_u32(&bar_mem);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 5 * 4096);
    for (int c = 0; c < 5; ++c) tma_load_2d(sbase + c * 4096, &in_map, c * 16, 0, bar);
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < 5 * 4096; i += blockDim.x) image[i] = sm[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int c = 0; c < 5; ++c) tma_store_2d(&out_map, sbase + c * 4096, c * 16, 0);
    bulk_commit();
    bulk_wait_all();
  }
}

// ---------------------------------------------------------------- host-side layouts
static inline uint16_t bf(float x) { __nv_bfloat16 b = __float2bfloat16(x); uint16_t u; memcpy(&u, &b, 2); return u; }
// [rows][cols] bf16 tile, 128-byte swizzle, chunks of 64 columns (chunk stride rows*128)
static size_t off_sw128(int rows, int row, int col) {
  return (size_t)(col / 64) * rows * 128 + (size_t)(row / 8) * 1024 + (row % 8) * 128 + ((((col % 64) / 8) ^ (row % 8)) * 16) + (col % 8) * 2;
}
// [rows][cols] bf16 tile, 32-byte swizzle, chunks of 16 columns (chunk stride rows*32)
static size_t off_sw32(int rows, int row, int col) {
  return (size_t)(col / 16) * rows * 32 + (size_t)row * 32 + ((((col % 16) / 8) ^ ((row >> 2) & 1)) * 16) + (col % 8) * 2;
}
static uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int type) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | ((uint64_t)type << 61);
}
enum Kind { K128, MN128, K32, MN32 };
struct Tile {            // storage: [rows][cols]; logical operand X(mn, k)
  Kind kind; int rows, cols; uint32_t base; std::vector<float> v;   // v[row*cols+col]
  float at(int mn, int k) const { return (kind == K128 || kind == K32) ? v[mn * cols + k] : v[k * cols + mn]; }
  size_t off(int row, int col) const { return base + ((kind == K128 || kind == MN128) ? off_sw128(rows, row, col) : off_sw32(rows, row, col)); }
  uint64_t kdesc(int k0) const {
    switch (kind) {
      case K128: return desc(base + (k0 / 64) * rows * 128 + (k0 % 64) * 2, 0, 1024, 2);
      case MN128: return desc(base + (k0 / 8) * 1024, rows * 128, 1024, 2);
      case K32: return desc(base + (k0 / 16) * rows * 32, 0, 256, 6);
      case MN32: return desc(base + (k0 / 8) * 256, rows * 32, 256, 6);
    }
    return 0;
  }
};

static uint32_t idesc(int M, int N, int a_mn, int b_mn) { return idesc_bf16(M, N, a_mn, b_mn); }

static int run(const char* name, Tile& A, Tile& B, int Mvalid, int N, int K) {
  std::vector<uint8_t> img(kSmem, 0);
  // garbage everywhere first (finite), so reads beyond a tile do not accidentally look right
  for (size_t i = 0; i + 1 < img.size(); i += 2) { uint16_t g = bf((float)((i / 2) % 7) - 3.f); memcpy(&img[i], &g, 2); }
  for (Tile* t : {&A, &B})
    for (int r = 0; r < t->rows; ++r)
      for (int c = 0; c < t->cols; ++c) { uint16_t u = bf(t->v[r * t->cols + c]); memcpy(&img[t->off(r, c)], &u, 2); }
  Job job{};
  job.nk = K / 16; job.N = N;
  job.idesc = idesc(128, N, A.kind == MN128 || A.kind == MN32, B.kind == MN128 || B.kind == MN32);
  for (int k = 0; k < job.nk; ++k) { job.da[k] = A.kdesc(k * 16); job.db[k] = B.kdesc(k * 16); }
  uint8_t* dimg; float* dout;
  CK(cudaMalloc(&dimg, kSmem)); CK(cudaMalloc(&dout, 128 * 256 * 4));
  CK(cudaMemcpy(dimg, img.data(), kSmem, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0, 128 * 256 * 4));
  k_lab<<<1, 128, kSmem>>>(dimg, kSmem, job, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-28s LAUNCH ERROR %s\n", name, cudaGetErrorString(e)); exit(2); }
  std::vector<float> out(128 * 256);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0; double maxerr = 0;
  for (int m = 0; m < Mvalid; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A.at(m, k) * B.at(n, k);
      double err = fabs(ref - out[m * 256 + n]);
      if (err > maxerr) maxerr = err;
      if (err > 1e-3) ++bad;
    }
  printf("%-28s %s  (mismatches %d of %d, max err %.3g)\n", name, bad ? "FAIL" : "ok", bad, Mvalid * N, maxerr);
  cudaFree(dimg); cudaFree(dout);
  return bad;
}

static Tile make(Kind kind, int rows, int cols, uint32_t base, unsigned seed) {
  Tile t{kind, rows, cols, base, std::vector<float>((size_t)rows * cols)};
  srand(seed);
  for (auto& x : t.v) x = (float)(rand() % 9 - 4);
  return t;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  CK(cudaFuncSetAttribute(k_lab, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  int bad = 0;
  const uint32_t bA = 0, bB = 64 * 1024;
  { Tile A = make(K128, 128, 256, bA, 1), B = make(K128, 80, 256, bB, 2); bad += run("S1: K128 x K128 (known)", A, B, 128, 80, 256); }
  { Tile A = make(K32, 128, 80, bA, 3), B = make(K128, 80, 80, bB, 4); bad += run("Y1: K32 x K128", A, B, 128, 80, 80); }
  { Tile A = make(K128, 128, 128, bA, 5), B = make(MN32, 128, 80, bB, 6); bad += run("Y2: K128 x MN32", A, B, 128, 80, 128); }
  { Tile A = make(MN32, 128, 80, bA, 7), B = make(MN32, 128, 80, bB, 8); bad += run("dGV: MN32 x MN32", A, B, 80, 80, 128); }
  { Tile A = make(K32, 128, 80, bA, 9), B = make(K32, 128, 80, bB, 10); bad += run("dGT: K32 x K32", A, B, 128, 128, 80); }
  { Tile A = make(MN32, 128, 80, bA, 11), B = make(MN128, 128, 256, bB, 12); bad += run("dV': MN32 x MN128", A, B, 80, 256, 128); }
  { Tile A = make(K32, 128, 80, bA, 13), B = make(MN128, 80, 256, bB, 14); bad += run("dT': K32 x MN128", A, B, 128, 256, 80); }

  // ---- TMA SWIZZLE_32B round trip
  {
    void* p = nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    Tile X = make(K32, 128, 80, 0, 21);
    std::vector<uint16_t> h(128 * 80);
    for (size_t i = 0; i < h.size(); ++i) h[i] = bf(X.v[i]);
    uint16_t *din, *dout2; uint8_t* dimg;
    CK(cudaMalloc(&din, h.size() * 2)); CK(cudaMalloc(&dout2, h.size() * 2)); CK(cudaMalloc(&dimg, 5 * 4096));
    CK(cudaMemcpy(din, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout2, 0, h.size() * 2));
    CUtensorMap mi, mo;
    cuuint64_t gdim[2] = {80, 128}; cuuint64_t gstr[1] = {160}; cuuint32_t box[2] = {16, 128}; cuuint32_t es[2] = {1, 1};
    CUresult r1 = enc(&mi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, din, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&mo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dout2, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("tensor map encode failed %d %d\n", (int)r1, (int)r2); return 3; }
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    k_tma<<<1, 128, 32768>>>(mi, mo, dimg);
    CK(cudaDeviceSynchronize());
    std::vector<uint8_t> img(5 * 4096); std::vector<uint16_t> back(h.size());
    CK(cudaMemcpy(img.data(), dimg, img.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(back.data(), dout2, back.size() * 2, cudaMemcpyDeviceToHost));
    int b1 = 0, b2 = 0;
    for (int r = 0; r < 128; ++r)
      for (int c = 0; c < 80; ++c) {
        uint16_t u; memcpy(&u, &img[X.off(r, c)], 2);
        if (u != h[r * 80 + c]) ++b1;
        if (back[r * 80 + c] != h[r * 80 + c]) ++b2;
      }
    printf("%-28s %s  (smem image mismatches %d, store round-trip mismatches %d)\n", "TMA SWIZZLE_32B", (b1 | b2) ? "FAIL" : "ok", b1, b2);
    bad += b1 + b2;
  }
  printf(bad ? "DESC LAB: FAILURES\n" : "DESC LAB: all ok\n");
  return bad ? 1 : 0;
}
