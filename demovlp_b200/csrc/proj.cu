// Projection heads that FEED the region-word path: the reference maps the 768-wide tower outputs to the
// common 256-wide embedding space with
//     text:   txt_proj = Sequential(ReLU(), Linear(768, 256))           model/model.py:40-43, applied at :88
//     video:  proj     = Linear(768, 256, bias=False)                    model/object_transformer.py:381, applied at :451
// and hands fp32 (B, L+1, 256) tensors to the loss, which the tensor-core kernels round to bf16 again.  Here the
// projection is ONE tcgen05 GEMM per tower whose epilogue adds the bias and writes the bf16 rows the loss kernels read
// directly:   Y[M tokens][256] = act(X)[M][K] W^T[K][256] + b,   act = ReLU (text) or identity (video).
// fp32 inputs (and the input ReLU) go through one fused elementwise pass that produces the bf16 A operand.
//
// CTA = 128 tokens x all 256 outputs (one TMEM accumulator of 256 columns), K streamed in 64-column stages by TMA.
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dvlp {
using namespace ptx;

int make_map_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols,
                CUtensorMapSwizzle swz);

namespace pj {
constexpr int kN = 256, kThreads = 192, kStages = 4;
constexpr int kABytes = 128 * 128, kBBytes = 256 * 128, kStage = kABytes + kBBytes;     // 64 bf16 columns per stage
constexpr int BARS = kStages * kStage, TMEMPTR = BARS + 16 * 8, kAlloc = TMEMPTR + 16;

template <typename Tin>
__global__ void __launch_bounds__(256) k_act_cast(const Tin* __restrict__ x, size_t n, int relu, __nv_bfloat16* __restrict__ o) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (i + e < n) {
      float v = to_f32(x[i + e]);
      if (relu) v = fmaxf(v, 0.f);
      o[i + e] = __float2bfloat16(v);
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
k_proj_gemm(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, int M, int K,
            const float* __restrict__ bias, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if (sbase & 1023u) __trap();
  volatile uint32_t* s_tmem = (volatile uint32_t*)(smem_raw + TMEMPTR);
  auto full = [&](int s) { return sbase + BARS + 8 * s; };
  auto empty = [&](int s) { return sbase + BARS + 8 * (kStages + s); };
  const uint32_t done = sbase + BARS + 8 * (2 * kStages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * 128;
  const int n_steps = K / 64;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapX); prefetch_tmap(&mapW);
    for (int s = 0; s < 2 * kStages + 1; ++s) mbar_init(sbase + BARS + 8 * s, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(sbase + TMEMPTR, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 0) {
    if (elect_one()) {
      for (int k = 0; k < n_steps; ++k) {
        const int s = k % kStages, it = k / kStages;
        if (it > 0) mbar_wait(empty(s), (it - 1) & 1);
        const uint32_t st = sbase + s * kStage;
        mbar_arrive_expect_tx(full(s), kStage);
        tma_load_2d(st, &mapX, k * 64, row0, full(s));                  // rows past M are zero-filled
        tma_load_2d(st + kABytes, &mapW, k * 64, 0, full(s));
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t id = idesc_bf16(128, 256);                        // X tile (K-major) x W (K-major: [256 outputs][K])
    const uint32_t hi = desc_hi(1024);
    for (int k = 0; k < n_steps; ++k) {
      const int s = k % kStages, it = k / kStages;
      mbar_wait(full(s), it & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = sbase + s * kStage;
        const uint32_t loA = desc_lo(st), loB = desc_lo(st + kABytes);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_bf16_w(tmem, loA + ((ks * 32) >> 4), hi, loB + ((ks * 32) >> 4), hi, id, (k | ks) != 0);
        tc_commit(empty(s));
        if (k == n_steps - 1) tc_commit(done);
      }
      __syncwarp();
    }
  } else {
    // epilogue: thread = token row; + bias, bf16 rows
    const int q = warp & 3, row = row0 + q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    mbar_wait(done, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < kN; c += 32) {
      float v[32];
      tmem_ld16(tmem + lane_off + c, reinterpret_cast<uint32_t*>(v));
      tmem_ld16(tmem + lane_off + c + 16, reinterpret_cast<uint32_t*>(v) + 16);
      tmem_ld_wait();
      if (row < M) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float b0 = bias ? bias[c + e] : 0.f, b1 = bias ? bias[c + e + 1] : 0.f;
          pk[e >> 1] = pack_bf16(v[e] + b0, v[e + 1] + b1);
        }
        uint4* dst = reinterpret_cast<uint4*>(out + (size_t)row * kN + c);
#pragma unroll
        for (int e = 0; e < 4; ++e) dst[e] = make_uint4(pk[4 * e], pk[4 * e + 1], pk[4 * e + 2], pk[4 * e + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}
}  // namespace pj

size_t proj_workspace_bytes(int M, int K, int in_dtype, int relu_in) {
  const bool pass = in_dtype == DVLP_F32 || relu_in;
  return align_up(pass ? (size_t)M * K * 2 : 0) + 256;
}

int run_proj_fwd(const void* x, int in_dtype, int M, int K, const void* w_bf16, const float* bias, int relu_in, void* out,
                 char* ws, cudaStream_t st) {
  const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
  if (in_dtype == DVLP_F32 || relu_in) {
    const size_t n = (size_t)M * K;
    const unsigned grid = (unsigned)((n / 4 + 255) / 256 + 1);
    if (in_dtype == DVLP_F32) pj::k_act_cast<float><<<grid, 256, 0, st>>>((const float*)x, n, relu_in, (__nv_bfloat16*)ws);
    else pj::k_act_cast<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, n, relu_in, (__nv_bfloat16*)ws);
    DVLP_LAUNCH_OK();
    xb = (const __nv_bfloat16*)ws;
  }
  CUtensorMap mX, mW;
  int rc;
  if ((rc = make_map_2d(&mX, xb, (uint64_t)M, (uint64_t)K, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map_2d(&mW, w_bf16, 256, (uint64_t)K, 256, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    DVLP_CUDA_OK(cudaFuncSetAttribute(pj::k_proj_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, pj::kAlloc));
    attr_once.done();
  }
  pj::k_proj_gemm<<<(M + 127) / 128, pj::kThreads, pj::kAlloc, st>>>(mX, mW, M, K, bias, (__nv_bfloat16*)out);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

}  // namespace dvlp
