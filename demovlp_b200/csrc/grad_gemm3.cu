// Gradient contractions of the LARGE-SHAPE region-word backward (partner of pair_bwd3.cu) as accumulator-stationary
// tcgen05 GEMMs.  pair_bwd3.cu leaves, per (text tile jt, video i) unit u = jt * Bv + i, three bf16 tiles:
//     P   [128 words][Rp regions] = dS alpha_r beta_w          (autograd of model/loss.py:235: S = Vh Th^T)
//     UD1 [128 words][Rp regions] = (u1 alpha) 2 dD1           (weights of d |c_w|^2 / dV, direct form of the cosine
//     Cb  [128 words][256]        = the weighted contexts c_w   denominator, model/loss.py:262-269, 286-291)
// and this file contracts them with the RAW bf16 embeddings (the alpha / beta factors are already inside P / UD1):
//
//   MODE V   dV'[i]^T (d x Rp)  = sum_jt  T_jt^T P[u] + Cb[u]^T UD1[u]     one CTA per (video, 128-column half of d,
//            K split over text tiles); accumulator lanes = embedding columns, TMEM columns = regions, so the epilogue
//            writes 128-byte rows of dV' per warp;  K = the 128 words of a tile
//   MODE T   dT'[jt]  (128 x d) = sum_i sum_c P[u][:, c] V_i[c]            one CTA per (text tile, video split);
//            K = the NB regions of a chunk, the video streams in chunks like in the pair kernels
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dvlp {
using namespace ptx;

int make_map_items(CUtensorMap* m, const void* base, uint64_t n0, uint64_t n1, uint64_t n2, uint32_t b1, uint32_t b2);
int make_map_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols,
                CUtensorMapSwizzle swz);

namespace gg3 {
constexpr int kRows = 128, kThreads = 192;
constexpr int kTileBytes = 16384, kCh32 = 4096;

struct Args {
  int Bv, Bt, R, W, Rp, nch, NB, WP, TPT, tiles;
  int n_groups, steps_per_split;   // MODE V: (video, half) groups x text-tile splits;  MODE T: text tiles x video splits
  int accumulate;                  // 1: red.global.add into `out` (zeroed by the caller), 0: plain stores
  float* out;                      // MODE V: dV' (Bv, R, 256);  MODE T: dT' (Bt * W, 256)
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(float* p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ uint32_t desc_hi32(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (6u << 29);
}

// ------------------------------------------------------------------------------------------------ MODE V



// dV'^T accumulators: lanes = 128 embedding columns of half h, columns = regions.  Shared memory holds ONE stage
// (X half 32 KB + a whole P tile, <= 144 KB at Rp = 576), double-buffered only when two fit.
template <int NST>
__global__ void __launch_bounds__(kThreads, 1)
k_grad_gemm3_v(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapU,
               const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapC, Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if (sbase & 1023u) __trap();
  const int ncs = a.Rp / 16;                               // 16-column sub-chunks of a P tile
  const uint32_t stage_bytes = 2 * kTileBytes + (uint32_t)ncs * kCh32;
  const uint32_t bars = NST * stage_bytes;
  volatile uint32_t* s_tmem = (volatile uint32_t*)(smem_raw + bars + 64);
  auto full = [&](int s) { return sbase + bars + 8 * s; };
  auto empty = [&](int s) { return sbase + bars + 8 * (NST + s); };
  const uint32_t done = sbase + bars + 8 * (2 * NST);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int grp = blockIdx.x % a.n_groups, z = blockIdx.x / a.n_groups;
  const int i = grp >> 1, h = grp & 1;
  const int k0 = z * a.steps_per_split;
  const int n_tiles = max(0, min(a.steps_per_split, a.tiles - k0));
  const int n_steps = 2 * n_tiles;                         // (P, T) and (UD1, Cb) per text tile

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapP); prefetch_tmap(&mapU); prefetch_tmap(&mapT); prefetch_tmap(&mapC);
    for (int s = 0; s < 2 * NST + 1; ++s) mbar_init(sbase + bars + 8 * s, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(sbase + bars + 64, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 0) {
    if (elect_one()) {
      for (int k = 0; k < n_steps; ++k) {
        const int s = k % NST, it = k / NST;
        if (it > 0) mbar_wait(empty(s), (it - 1) & 1);
        const uint32_t st = sbase + s * stage_bytes;
        const int jt = k0 + (k >> 1);
        const long long u = (long long)jt * a.Bv + i;
        mbar_arrive_expect_tx(full(s), stage_bytes);
        if ((k & 1) == 0) {
          for (int c = 0; c < 2; ++c) tma_load_3d(st + c * kTileBytes, &mapT, h * 128 + c * 64, 0, jt * a.TPT, full(s));
          for (int c = 0; c < ncs; ++c) tma_load_2d(st + 2 * kTileBytes + c * kCh32, &mapP, c * 16, (int)(u * kRows), full(s));
        } else {
          for (int c = 0; c < 2; ++c) tma_load_2d(st + c * kTileBytes, &mapC, h * 128 + c * 64, (int)(u * kRows), full(s));
          for (int c = 0; c < ncs; ++c) tma_load_2d(st + 2 * kTileBytes + c * kCh32, &mapU, c * 16, (int)(u * kRows), full(s));
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t hiK = desc_hi(1024), hi32 = desc_hi32(256);
    const int n1 = a.Rp > 256 ? 256 : a.Rp, n2 = a.Rp - n1;
    const uint32_t id1 = idesc_bf16(128, n1, 1, 1), id2 = idesc_bf16(128, n2 > 0 ? n2 : 16, 1, 1);
    for (int k = 0; k < n_steps; ++k) {
      const int s = k % NST, it = k / NST;
      mbar_wait(full(s), it & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = sbase + s * stage_bytes;
        // A = X^T: MN-major, M = 128 embedding columns (two 64-column groups 16 KB apart), K = 128 words
        const uint32_t loX = desc_lo(st, kTileBytes);
        // B = P: MN-major, N = regions (16-column sub-chunks 4 KB apart), K = 128 words
        const uint32_t loP = desc_lo(st + 2 * kTileBytes, kCh32);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          mma_bf16_w(tmem, loX + ((ks * 2048) >> 4), hiK, loP + ((ks * 512) >> 4), hi32, id1, (k | ks) != 0);
          if (n2 > 0)
            mma_bf16_w(tmem + 256, loX + ((ks * 2048) >> 4), hiK, loP + ((16 * kCh32 + ks * 512) >> 4), hi32, id2, (k | ks) != 0);
        }
        tc_commit(empty(s));
        if (k == n_steps - 1) tc_commit(done);
      }
      __syncwarp();
    }
  } else {
    // epilogue: thread = embedding column (TMEM lane), walks the regions
    const int q = warp & 3;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int dcol = h * 128 + q * 32 + lane;
    if (n_steps > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
      float* dst = a.out + (size_t)i * a.R * 256 + dcol;
      for (int r0 = 0; r0 < a.R; r0 += 16) {
        float v[16];
        tmem_ld16(tmem + lane_off + r0, reinterpret_cast<uint32_t*>(v));
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          if (r0 + e < a.R) {
            if (a.accumulate) red_add(dst + (size_t)(r0 + e) * 256, v[e]);
            else dst[(size_t)(r0 + e) * 256] = v[e];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ MODE T
template <int NB>
struct LayT {
  static constexpr int NG = NB / 16;
  static constexpr int VBOX = NB * 128, VSTAGE = 4 * VBOX, PB = NG * kCh32;
  static constexpr int STAGE = PB + VSTAGE;
  static constexpr int kStages = 4;
  static constexpr int BARS = kStages * STAGE;
  static constexpr int TMEMPTR = BARS + 16 * 8;
  static constexpr int kAlloc = TMEMPTR + 16;
  static_assert(STAGE % 1024 == 0, "alignment");
  static_assert(kAlloc <= 232448, "shared memory budget");
};

template <int NB>
__global__ void __launch_bounds__(kThreads, 1)
k_grad_gemm3_t(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapV, Args a) {
  using L = LayT<NB>;
  constexpr int kStages = L::kStages, NG = L::NG;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if (sbase & 1023u) __trap();
  volatile uint32_t* s_tmem = (volatile uint32_t*)(smem_raw + L::TMEMPTR);
  auto full = [&](int s) { return sbase + L::BARS + 8 * s; };
  auto empty = [&](int s) { return sbase + L::BARS + 8 * (kStages + s); };
  const uint32_t done = sbase + L::BARS + 8 * (2 * kStages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int jt = blockIdx.x % a.n_groups, z = blockIdx.x / a.n_groups;
  const int i0 = z * a.steps_per_split;
  const int n_vid = max(0, min(a.steps_per_split, a.Bv - i0));
  const int n_steps = n_vid * a.nch;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapP); prefetch_tmap(&mapV);
    for (int s = 0; s < 2 * kStages + 1; ++s) mbar_init(sbase + L::BARS + 8 * s, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(sbase + L::TMEMPTR, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 0) {
    if (elect_one()) {
      int i = i0, c = 0;
      for (int k = 0; k < n_steps; ++k) {
        const int s = k % kStages, it = k / kStages;
        if (it > 0) mbar_wait(empty(s), (it - 1) & 1);
        const uint32_t st = sbase + s * L::STAGE;
        const long long u = (long long)jt * a.Bv + i;
        mbar_arrive_expect_tx(full(s), L::STAGE);
        for (int sc = 0; sc < NG; ++sc) tma_load_2d(st + sc * kCh32, &mapP, c * NB + sc * 16, (int)(u * kRows), full(s));
        for (int kc = 0; kc < 4; ++kc) tma_load_3d(st + L::PB + kc * L::VBOX, &mapV, kc * 64, c * NB, i, full(s));
        if (++c == a.nch) { c = 0; ++i; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idT = idesc_bf16(128, 256, 0, 1);      // P chunk (K-major, M = words) x V chunk (MN-major, N = d)
    const uint32_t hiK = desc_hi(1024), hi32 = desc_hi32(256);
    for (int k = 0; k < n_steps; ++k) {
      const int s = k % kStages, it = k / kStages;
      mbar_wait(full(s), it & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = sbase + s * L::STAGE;
        const uint32_t loP = desc_lo(st), loV = desc_lo(st + L::PB, L::VBOX);
#pragma unroll
        for (int ks = 0; ks < NG; ++ks)
          mma_bf16_w(tmem, loP + ((ks * kCh32) >> 4), hi32, loV + ((ks * 2048) >> 4), hiK, idT, (k | ks) != 0);
        tc_commit(empty(s));
        if (k == n_steps - 1) tc_commit(done);
      }
      __syncwarp();
    }
  } else {
    // epilogue: thread = word row of the tile
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    if (n_steps > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
      const int ts = row / a.WP, w = row - ts * a.WP, j = jt * a.TPT + ts;
      const bool valid = j < a.Bt && w < a.W;
      float* dst = a.out + ((size_t)j * a.W + w) * 256;
#pragma unroll 1
      for (int c = 0; c < 256; c += 32) {
        float v[32];
        tmem_ld16(tmem + lane_off + c, reinterpret_cast<uint32_t*>(v));
        tmem_ld16(tmem + lane_off + c + 16, reinterpret_cast<uint32_t*>(v) + 16);
        tmem_ld_wait();
        if (valid) {
          if (a.accumulate) {
#pragma unroll
            for (int x = 0; x < 32; x += 4) red_add_v4(dst + c + x, v[x], v[x + 1], v[x + 2], v[x + 3]);
          } else {
#pragma unroll
            for (int x = 0; x < 32; x += 4) *reinterpret_cast<float4*>(dst + c + x) = make_float4(v[x], v[x + 1], v[x + 2], v[x + 3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// dgt (Bt, W, W) fp32 from the padded per-tile accumulators dgt3 (tiles, 128, 128) of pair_bwd3.cu
__global__ void __launch_bounds__(256) k_unpack_dgt3(const float* __restrict__ dgt3, int Bt, int W, int WP, int TPT,
                                                     float* __restrict__ dgt) {
  const size_t n = (size_t)Bt * W * W;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int w2 = (int)(e % W), w = (int)((e / W) % W), j = (int)(e / ((size_t)W * W));
    const int tile = j / TPT, ts = j - tile * TPT;
    dgt[e] = dgt3[((size_t)tile * 128 + ts * WP + w) * 128 + ts * WP + w2];
  }
}
}  // namespace gg3

// dV' (Bv, R, 256) and dT' (Bt * W, 256), both OVERWRITTEN / accumulated from zero: the caller zeroes them first;
// dgt (Bt, W, W) unpacked from dgt3.
int run_grad_gemm3(const void* vb, const void* tb, int Bv, int Bt, int R, int W, int d, char* ws, const Workspace& Lw,
                   cudaStream_t st) {
  Shape3 s3;
  if (!shape3(R, W, d, &s3)) return DVLP_ERR_BAD_SHAPE;
  const int tiles = (Bt + s3.TPT - 1) / s3.TPT;
  const long long units = (long long)tiles * Bv;
  CUtensorMap mP, mU, mT, mC, mV;
  int rc;
  if ((rc = make_map_2d(&mP, ws + Lw.pa, (uint64_t)units * 128, (uint64_t)s3.Rp, 128, 16, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  if ((rc = make_map_2d(&mU, ws + Lw.ud1, (uint64_t)units * 128, (uint64_t)s3.Rp, 128, 16, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  if ((rc = make_map_2d(&mC, ws + Lw.cb, (uint64_t)units * 128, 256, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map_items(&mT, tb, 256, (uint64_t)W, (uint64_t)Bt, (uint32_t)s3.WP, (uint32_t)s3.TPT))) return rc;
  if ((rc = make_map_items(&mV, vb, 256, (uint64_t)R, (uint64_t)Bv, (uint32_t)s3.NB, 1))) return rc;
  int dev = 0, sms = 148;
  DVLP_CUDA_OK(cudaGetDevice(&dev));
  DVLP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  gg3::Args a{};
  a.Bv = Bv; a.Bt = Bt; a.R = R; a.W = W; a.Rp = s3.Rp; a.nch = s3.nch; a.NB = s3.NB; a.WP = s3.WP; a.TPT = s3.TPT;
  a.tiles = tiles;
  // ---- MODE V
  {
    a.out = (float*)(ws + Lw.dvh);
    a.n_groups = Bv * 2;
    int splits = sms / a.n_groups;
    if (splits < 1) splits = 1;
    if (splits > tiles) splits = tiles;
    a.steps_per_split = (tiles + splits - 1) / splits;
    splits = (tiles + a.steps_per_split - 1) / a.steps_per_split;
    a.accumulate = splits > 1 ? 1 : 0;
    const uint32_t stage_bytes = 2 * gg3::kTileBytes + (uint32_t)(s3.Rp / 16) * gg3::kCh32;
    const int nst = 2 * stage_bytes + 256 <= 232448 ? 2 : 1;
    const size_t smem = (size_t)nst * stage_bytes + 256;
    static DeviceOnce attr_once;
    if (attr_once.pending()) {
      DVLP_CUDA_OK(cudaFuncSetAttribute(gg3::k_grad_gemm3_v<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      DVLP_CUDA_OK(cudaFuncSetAttribute(gg3::k_grad_gemm3_v<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      attr_once.done();
    }
    if (nst == 2) gg3::k_grad_gemm3_v<2><<<a.n_groups * splits, gg3::kThreads, smem, st>>>(mP, mU, mT, mC, a);
    else gg3::k_grad_gemm3_v<1><<<a.n_groups * splits, gg3::kThreads, smem, st>>>(mP, mU, mT, mC, a);
    DVLP_LAUNCH_OK();
  }
  // ---- MODE T
  {
    a.out = (float*)(ws + Lw.dth);
    a.n_groups = tiles;
    int splits = sms / tiles;
    if (splits < 1) splits = 1;
    if (splits > Bv) splits = Bv;
    a.steps_per_split = (Bv + splits - 1) / splits;
    splits = (Bv + a.steps_per_split - 1) / a.steps_per_split;
    a.accumulate = splits > 1 ? 1 : 0;
    static DeviceOnce attr_once;
    if (attr_once.pending()) {
      DVLP_CUDA_OK(cudaFuncSetAttribute(gg3::k_grad_gemm3_t<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, gg3::LayT<48>::kAlloc));
      DVLP_CUDA_OK(cudaFuncSetAttribute(gg3::k_grad_gemm3_t<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, gg3::LayT<32>::kAlloc));
      attr_once.done();
    }
    if (s3.NB == 48) gg3::k_grad_gemm3_t<48><<<tiles * splits, gg3::kThreads, gg3::LayT<48>::kAlloc, st>>>(mP, mV, a);
    else gg3::k_grad_gemm3_t<32><<<tiles * splits, gg3::kThreads, gg3::LayT<32>::kAlloc, st>>>(mP, mV, a);
    DVLP_LAUNCH_OK();
  }
  {
    const size_t n = (size_t)Bt * W * W;
    int grid = (int)((n + 255) / 256);
    grid = grid > 148 * 8 ? 148 * 8 : grid;
    gg3::k_unpack_dgt3<<<grid, 256, 0, st>>>((const float*)(ws + Lw.dgt3), Bt, W, s3.WP, s3.TPT, (float*)(ws + Lw.dgt));
    DVLP_LAUNCH_OK();
  }
  return DVLP_OK;
}

}  // namespace dvlp
