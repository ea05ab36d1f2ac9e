// Region-word pair scores on tcgen05 / TMEM / TMA (sm_100a): BACKWARD kernel.
//
// Recomputes the forward of pair_tc.cu per (video, 4-text tile) -- the attention tensors are never
// stored -- and back-propagates d_scores through both attention directions (DESIGN.md has the
// derivation; reference autograd path: model/loss.py:209-330).  All matrix products run on tcgen05:
//
//   S1 = T V^T, S2 = V T^T                      raw similarities, both orientations      (as forward)
//   Y1 = U1 G^V,  Y2 = U2 G^T                   Gram-form denominators                   (as forward)
//   dG^V += UD1^T U1,  dG^T += UD2^T U2         gradient of the Gram matrices  (UD = U * dD per query)
//
// The two d-wide gradient products  dV' = sum_j (P1^T + P2) T_j  and  dT' = sum_i (P1^T + P2)^T V_i
// (P = dS * alpha_r * beta_w, bf16) are NOT formed here: per-tile accumulators would have to be flushed
// with fp32 L2 atomics (250 KB per tile; measured 6 TB/s chip-wide at best, tools/ubench/red_bw.cu),
// which bounded the first version of this kernel.  Instead the two bf16 P tiles (2 x 20 KB) leave
// through TMA stores and grad_gemm.cu contracts them against T / V as dense, accumulator-stationary
// tcgen05 GEMMs (no atomics, 10x less traffic).
//
// Operand staging (all [region rows][word columns], 128-byte swizzle): U1^T / UD1^T written by group A
// (thread = word) with transposed 16-bit stores, U2 / UD2 written row-wise by group B (thread = region);
// P1^T and P2 re-use the U buffers once the dG MMAs have consumed them.  The same buffer serves as a
// K-major operand in one product and as an MN-major operand in the other (UMMA descriptors).
// The cross-row sums of the l2-norm backward (sum over the query axis = over THREADS in each group's
// orientation) are done with warp-shuffle transpose-reductions.
// The small Gram gradients dG^V / dG^T are flushed from TMEM with vector red.global adds.
//
// Warp roles (512 threads, 4 warpgroups, registers re-split with setmaxnreg):
//   WG0: warp 0 TMA producer, warp 1 MMA issuer;  WG1: group A;  WG2 / WG3: group B for texts 0,1 / 2,3 of the tile.
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace dvlp {
using namespace ptx;

void* g_trace_buffer = nullptr;   // set through dvlp_debug_set_trace (tools/trace_tiles.py)

namespace tcb {
constexpr int kD = 256, kW = 32, kTPT = 4, kRows = 128, kThreads = 512;
constexpr int kTileBytes = 16384;
constexpr int kGroup = 384;

template <int R>
struct Lay {
  static constexpr int kVChunk = R * 128;
  static constexpr int V = 0;
  static constexpr int T = V + 4 * kVChunk;
  static constexpr int GV = T + 4 * kTileBytes;
  static constexpr int GT = GV + 2 * kVChunk;
  static constexpr int A1 = GT + kTileBytes;        // U1^T -> P1^T   [R rows][2 chunks x 128 B]
  static constexpr int A2 = A1 + 2 * kVChunk;       // UD1^T
  static constexpr int B1 = A2 + 2 * kVChunk;       // U2   -> P2
  static constexpr int B2 = B1 + 2 * kVChunk;       // UD2
  static constexpr int VEC = B2 + 2 * kVChunk;
  static constexpr int ALPHA = VEC;                 // [R]
  static constexpr int M2V = ALPHA + R * 4;         // [R]
  static constexpr int BETA = M2V + R * 4;          // [128]
  static constexpr int M2T = BETA + 512;            // [128]
  static constexpr int INVNA = M2T + 512;           // [4][R]
  static constexpr int INVNB = INVNA + 4 * R * 4;   // [128]
  static constexpr int TA = INVNB + 512;            // [4][R]   sum_w A dA per (text, region)
  static constexpr int TBP = TA + 4 * R * 4;        // [4 texts][4 warps][32] per-warp partials of sum_r A dA
  static constexpr int BARS = TBP + 2048;
  static constexpr int TMEMPTR = BARS + 16 * 8;
  static constexpr int kBytes = TMEMPTR + 16;
  static constexpr int kAlloc = kBytes;             // dynamic shared memory starts 1024-B aligned (checked at run time)
  static_assert(kAlloc <= 232448, "shared memory budget");
};

enum Bar { V_FULL = 0, T_FULL, GT_FULL, T_EMPTY, S_DONE, U1_READY, U2_READY, Y1_DONE, Y2_DONE, UD1_READY, UD2_READY,
           DGV_DONE, DGT_DONE, P_READY, P_STORED, kNumBars };

// TMEM columns (everything is per-tile scratch, reused phase by phase)
constexpr uint32_t cS1 = 0, cS2 = 128, cY2 = 256, cDGT = 384;

struct Args {
  int Bv, Bt, tiles_per_video;
  long long total_units;
  const float *alpha, *beta, *m2v, *m2t, *d_scores;
  const float *nrm_a, *nrm_b;   // attention l2-norm factors written by the forward kernel (pair_tc.cu)
  float lam, lam2;
  int focal;
  float *dgv, *dgt;
  long long* trace;   // debug: per-phase clock64 stamps of CTA 0 (tools/trace_tiles.py), NULL in production
};
// trace slot layout: [role 0..4][tile 0..31][event 0..15]
#define TR(role, ev)                                                                        \
  do {                                                                                      \
    if (a.trace && blockIdx.x == 0 && lane == 0 && k < 32) a.trace[((role) * 32 + k) * 16 + (ev)] = clock64(); \
  } while (0)
#define TRA(ev) do { if (warp == 4) TR(2, ev); } while (0)
#define TRB(ev) do { if (q == 0) TR(3 + hb, ev); } while (0)

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ uint16_t bf16_bits(float x) {
  __nv_bfloat16 b = __float2bfloat16(x);
  return *reinterpret_cast<uint16_t*>(&b);
}
// Transpose-reduce: on return lane l holds sum over the 32 lanes of v[l].
__device__ __forceinline__ float lanes_to_columns32(float (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      float send = up ? v[i] : v[i + h];
      float keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0];
}
// 16 columns: lane l ends with the 32-lane sum of v[l & 15].
__device__ __forceinline__ float lanes_to_columns16(float (&v)[16], int lane) {
#pragma unroll
  for (int h = 8; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      float send = up ? v[i] : v[i + h];
      float keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

template <int R>
__global__ void __launch_bounds__(kThreads, 1)
k_pair_tc_bwd(const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapT,
              const __grid_constant__ CUtensorMap mapGV, const __grid_constant__ CUtensorMap mapGT,
              const __grid_constant__ CUtensorMap mapPa, const __grid_constant__ CUtensorMap mapPb, Args a) {
  using L = Lay<R>;
  static_assert(R % 16 == 0 && R >= 16 && R <= 128, "R must be a multiple of 16 in [16,128]");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if (sbase & 1023u) __trap();          // the 128-byte-swizzle tiles need a 1024-byte aligned base
  uint8_t* sgen = smem_raw;
  float* s_alpha = (float*)(sgen + L::ALPHA);
  float* s_m2v = (float*)(sgen + L::M2V);
  float* s_beta = (float*)(sgen + L::BETA);
  float* s_m2t = (float*)(sgen + L::M2T);
  float* s_invnA = (float*)(sgen + L::INVNA);
  float* s_invnB = (float*)(sgen + L::INVNB);
  float* s_tA = (float*)(sgen + L::TA);
  float* s_tBp = (float*)(sgen + L::TBP);
  volatile uint32_t* s_tmem = (volatile uint32_t*)(sgen + L::TMEMPTR);
  auto bar = [&](int i) { return sbase + L::BARS + 8 * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long u0 = a.total_units * blockIdx.x / gridDim.x;
  const long long u1 = a.total_units * (blockIdx.x + 1) / gridDim.x;
  const int n_tiles = (int)(u1 - u0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapV); prefetch_tmap(&mapT); prefetch_tmap(&mapGV); prefetch_tmap(&mapGT);
    prefetch_tmap(&mapPa); prefetch_tmap(&mapPb);
    for (int i = 0; i < kNumBars; ++i) {
      uint32_t cnt = 1;
      if (i == U1_READY || i == UD1_READY) cnt = 128;
      if (i == U2_READY || i == UD2_READY) cnt = 256;
      if (i == P_READY) cnt = 384;
      mbar_init(bar(i), cnt);
    }
    fence_mbar_init();
  }
  if (warp == 3) {
    tmem_alloc(sbase + L::TMEMPTR, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  auto video_of = [&](int k) { return (int)((u0 + k) / a.tiles_per_video); };
  auto tile_of = [&](int k) { return (int)((u0 + k) % a.tiles_per_video); };
  auto is_switch = [&](int k) { return k == 0 || video_of(k) != video_of(k - 1); };

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0 && lane == 0) {
      // =============================== TMA producer ===============================
      for (int k = 0; k < n_tiles; ++k) {
        const int i = video_of(k), jt = tile_of(k);
        if (k > 0) mbar_wait(bar(T_EMPTY), (k - 1) & 1);   // S GEMMs of tile k-1 have read T
        mbar_arrive_expect_tx(bar(T_FULL), 4 * kTileBytes);
        for (int c = 0; c < 4; ++c) tma_load_2d(sbase + L::T + c * kTileBytes, &mapT, c * 64, jt * kRows, bar(T_FULL));
        // everything else is still read by the groups / the Y GEMMs until tile k-1 has written its P tiles
        if (k > 0) mbar_wait(bar(P_READY), (k - 1) & 1);
        if (is_switch(k)) {
          mbar_arrive_expect_tx(bar(V_FULL), 6 * L::kVChunk + 2 * R * 4);
          for (int c = 0; c < 4; ++c) tma_load_2d(sbase + L::V + c * L::kVChunk, &mapV, c * 64, i * R, bar(V_FULL));
          for (int c = 0; c < 2; ++c) tma_load_2d(sbase + L::GV + c * L::kVChunk, &mapGV, c * 64, i * R, bar(V_FULL));
          bulk_load(sbase + L::ALPHA, a.alpha + (size_t)i * R, R * 4, bar(V_FULL));
          bulk_load(sbase + L::M2V, a.m2v + (size_t)i * R, R * 4, bar(V_FULL));
        }
        mbar_arrive_expect_tx(bar(GT_FULL), kTileBytes + 1024 + 4 * R * 4 + 512);
        bulk_load(sbase + L::INVNA, a.nrm_a + (size_t)(u0 + k) * kTPT * R, 4 * R * 4, bar(GT_FULL));
        bulk_load(sbase + L::INVNB, a.nrm_b + (size_t)(u0 + k) * kRows, 512, bar(GT_FULL));
        tma_load_2d(sbase + L::GT, &mapGT, 0, jt * kRows, bar(GT_FULL));
        bulk_load(sbase + L::BETA, a.beta + (size_t)jt * kRows, 512, bar(GT_FULL));
        bulk_load(sbase + L::M2T, a.m2t + (size_t)jt * kRows, 512, bar(GT_FULL));
      }
    } else if (warp == 1) {
      // ============== P stores: the two bf16 P tiles of each unit leave through TMA (grad_gemm.cu) ==============
      if (lane == 0) {
        for (int k = 0; k < n_tiles; ++k) {
          mbar_wait(bar(P_READY), k & 1);
          const int row = (int)((u0 + k) * R);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            tma_store_2d(&mapPa, sbase + L::A1 + c * L::kVChunk, c * 64, row);
            tma_store_2d(&mapPb, sbase + L::B1 + c * L::kVChunk, c * 64, row);
          }
          bulk_commit();
          bulk_wait_read0();                 // the staging tiles may be overwritten (next tile's U)
          mbar_arrive(bar(P_STORED));
        }
        bulk_wait_all();                     // global writes complete before the CTA retires
      }
    } else if (warp == 2) {
      // ================== MMA issuer, t2i (group B) chain: Y2 = U2 G^T, dG^T = UD2^T U2 ==================
      constexpr uint32_t idY2 = idesc_bf16(128, 32);
      constexpr uint32_t idDGT = idesc_bf16(128, 128, 1, 1);       // UD2^T U2: both MN-major
      constexpr uint32_t CH = L::kVChunk;
      const uint32_t hiK = desc_hi(1024);
      const uint32_t loGT = desc_lo(sbase + L::GT), loB1 = desc_lo(sbase + L::B1);
      const uint32_t loB1m = desc_lo(sbase + L::B1, CH), loB2m = desc_lo(sbase + L::B2, CH);
      for (int k = 0; k < n_tiles; ++k) {
        const uint32_t ph = k & 1;
        if (k > 0) mbar_wait(bar(P_READY), (k - 1) & 1);   // group A has read dG^T of tile k-1
        mbar_wait(bar(U2_READY), ph);
        mbar_wait(bar(GT_FULL), ph);
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int t = 0; t < kTPT; ++t)
#pragma unroll
            for (int s = 0; s < 2; ++s)
              mma_bf16_w(tmem + cY2 + t * 32, loB1 + (((t >> 1) * CH + (t & 1) * 64 + s * 32) >> 4), hiK,
                         loGT + ((t * 4096 + s * 32) >> 4), hiK, idY2, s != 0);
          tc_commit(bar(Y2_DONE));
        }
        __syncwarp();
        mbar_wait(bar(UD2_READY), ph);
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int ks = 0; ks < R / 16; ++ks)
            mma_bf16_w(tmem + cDGT, loB2m + ((ks * 2048) >> 4), hiK, loB1m + ((ks * 2048) >> 4), hiK, idDGT, ks != 0);
          tc_commit(bar(DGT_DONE));
        }
        __syncwarp();
      }
    } else if (warp == 3) {
      // ========= MMA issuer: S GEMMs and the i2t (group A) chain: Y1 = U1 G^V, dG^V = UD1^T U1 =========
      // Descriptors are loop invariant: their words are built once and every MMA only adds a constant to
      // the 14-bit address field (a single thread issuing dependent 64-bit arithmetic per MMA cost ~130
      // cycles per instruction in the first version of this kernel).
      constexpr uint32_t idS1 = idesc_bf16(128, R), idS2 = idesc_bf16(128, 128);
      constexpr uint32_t idY1 = idesc_bf16(128, R, 1, 0);          // A = U1 (MN-major view of U1^T)
      constexpr uint32_t CH = L::kVChunk;                          // chunk stride of R-row buffers
      const uint32_t hiK = desc_hi(1024);                          // K-major tiles: 8-row groups 1024 B apart
      const uint32_t loT = desc_lo(sbase + L::T), loV = desc_lo(sbase + L::V), loGV = desc_lo(sbase + L::GV);
      const uint32_t loA1 = desc_lo(sbase + L::A1), loA2 = desc_lo(sbase + L::A2);
      const uint32_t loA1m = desc_lo(sbase + L::A1, CH);           // MN-major view
      int vloads = 0;
      for (int k = 0; k < n_tiles; ++k) {
        const uint32_t ph = k & 1;
        if (is_switch(k)) { mbar_wait(bar(V_FULL), vloads & 1); ++vloads; }
        mbar_wait(bar(T_FULL), ph);
        if (k > 0) mbar_wait(bar(P_READY), (k - 1) & 1);   // both groups are done with the TMEM of tile k-1
        tc_fence_after();
        TR(1, 0);
        if (lane == 0) {
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {
            const uint32_t oT = ((ks >> 2) * kTileBytes + (ks & 3) * 32) >> 4, oV = ((ks >> 2) * CH + (ks & 3) * 32) >> 4;
            mma_bf16_w(tmem + cS1, loT + oT, hiK, loV + oV, hiK, idS1, ks != 0);
          }
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {
            const uint32_t oT = ((ks >> 2) * kTileBytes + (ks & 3) * 32) >> 4, oV = ((ks >> 2) * CH + (ks & 3) * 32) >> 4;
            mma_bf16_w(tmem + cS2, loV + oV, hiK, loT + oT, hiK, idS2, ks != 0);
          }
          tc_commit(bar(S_DONE));
          tc_commit(bar(T_EMPTY));
        }
        __syncwarp();
        TR(1, 1);
        // ---- Y1 = U1 G^V (A: MN-major from U1^T)
        mbar_wait(bar(U1_READY), ph);
        tc_fence_after();
        TR(1, 2);
        if (lane == 0) {
#pragma unroll
          for (int ks = 0; ks < R / 16; ++ks)
            mma_bf16_w(tmem + cS1, loA1m + ((ks * 2048) >> 4), hiK, loGV + (((ks >> 2) * CH + (ks & 3) * 32) >> 4), hiK,
                       idY1, ks != 0);
          tc_commit(bar(Y1_DONE));
        }
        __syncwarp();
        TR(1, 3);
        // ---- dG^V = UD1^T(K-major rows=regions) x U1^T(K-major rows=regions), K = 128 words
        mbar_wait(bar(UD1_READY), ph);
        tc_fence_after();
        TR(1, 4);
        if (lane == 0) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t o = ((ks >> 2) * CH + (ks & 3) * 32) >> 4;
            mma_bf16_w(tmem + cS1, loA2 + o, hiK, loA1 + o, hiK, idS1, ks != 0);
          }
          tc_commit(bar(DGV_DONE));
        }
        __syncwarp();
        TR(1, 5);
      }
    }
  } else if (warp < 8) {
    // ================== group A: i2t direction, thread = word row ==================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    // transposed element (row r, col m): chunk = m/64, 16-byte segment (m%64)/8 ^ (r%8), 2-byte slot m%8
    const uint32_t t_chunk = (uint32_t)(m >> 6) * L::kVChunk, t_seg = (uint32_t)((m & 63) >> 3), t_sub = (uint32_t)(m & 7) * 2;
    auto t_addr = [&](uint32_t buf, int r) {
      return buf + t_chunk + (uint32_t)(r >> 3) * 1024 + (uint32_t)(r & 7) * 128 + ((t_seg ^ (uint32_t)(r & 7)) << 4) + t_sub;
    };
    int vloads = 0;
    for (int k = 0; k < n_tiles; ++k) {
      const uint32_t ph = k & 1;
      const int i = video_of(k), jt = tile_of(k), j = jt * kTPT + q;
      TRA(0);
      if (is_switch(k)) { mbar_wait(bar(V_FULL), vloads & 1); ++vloads; }
      mbar_wait(bar(GT_FULL), ph);
      mbar_wait(bar(S_DONE), ph);
      tc_fence_after();
      TRA(1);
      const float beta_m = s_beta[m];
      const float g = j < a.Bt ? a.d_scores[(size_t)i * a.Bt + j] : 0.f;
      float S[R], Y[R];
      uint32_t up[R / 2];                      // u = focal-selected exp, bf16 pairs (exactly what the MMA sees)
      {
        uint32_t* raw = reinterpret_cast<uint32_t*>(S);
#pragma unroll
        for (int c = 0; c < R / 16; ++c) tmem_ld16(tmem + cS1 + lane_off + c * 16, raw + c * 16);
        tmem_ld_wait();
      }
      // the query-axis l2 norms of model/loss.py:238 come from the forward kernel (nrm_a / nrm_b, loaded with
      // the per-tile vectors): no norm pass and no exchange between the two groups in the backward
#pragma unroll
      for (int r = 0; r < R; r += 4) {
        const float4 al = *reinterpret_cast<const float4*>(s_alpha + r);
        S[r + 0] *= beta_m * al.x; S[r + 1] *= beta_m * al.y; S[r + 2] *= beta_m * al.z; S[r + 3] *= beta_m * al.w;
      }
      TRA(3);
      if (k > 0) mbar_wait(bar(P_STORED), (k - 1) & 1);   // TMA stores of tile k-1 have read A1 / B1
      float acc4[4];
      // forward pass 2 (e kept in Y[] until the focal threshold is known)
      acc4[0] = acc4[1] = acc4[2] = acc4[3] = 0.f;
#pragma unroll
      for (int r = 0; r < R; r += 4) {
        const float4 na = *reinterpret_cast<const float4*>(s_invnA + q * R + r);
        const float4 mk = *reinterpret_cast<const float4*>(s_m2v + r);
        Y[r + 0] = ex2(fmaf(leaky(S[r + 0]) * na.x, a.lam2, mk.x));
        Y[r + 1] = ex2(fmaf(leaky(S[r + 1]) * na.y, a.lam2, mk.y));
        Y[r + 2] = ex2(fmaf(leaky(S[r + 2]) * na.z, a.lam2, mk.z));
        Y[r + 3] = ex2(fmaf(leaky(S[r + 3]) * na.w, a.lam2, mk.w));
#pragma unroll
        for (int h = 0; h < 4; ++h) acc4[h] += Y[r + h];
      }
      const float thr = a.focal ? ((acc4[0] + acc4[1]) + (acc4[2] + acc4[3])) * (1.f / R) : -1.f;
      acc4[0] = acc4[1] = acc4[2] = acc4[3] = 0.f;
#pragma unroll
      for (int r = 0; r < R; r += 2) {
        const uint32_t pk = pack_bf16(Y[r] > thr ? Y[r] : 0.f, Y[r + 1] > thr ? Y[r + 1] : 0.f);
        up[r / 2] = pk;
        acc4[(r / 2) & 3] = fmaf(bf16_lo(pk), S[r], acc4[(r / 2) & 3]);
        acc4[(r / 2) & 3] = fmaf(bf16_hi(pk), S[r + 1], acc4[(r / 2) & 3]);
        st_shared_u16(t_addr(sbase + L::A1, r), (uint16_t)(pk & 0xffffu));
        st_shared_u16(t_addr(sbase + L::A1, r + 1), (uint16_t)(pk >> 16));
      }
      const float num = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
      fence_proxy_async();
      mbar_arrive(bar(U1_READY));
      TRA(4);
      // ---- denominators and per-word scalars
      mbar_wait(bar(Y1_DONE), ph);
      tc_fence_after();
      TRA(5);
      {
        uint32_t* raw = reinterpret_cast<uint32_t*>(Y);
#pragma unroll
        for (int c = 0; c < R / 16; ++c) tmem_ld16(tmem + cS1 + lane_off + c * 16, raw + c * 16);
        tmem_ld_wait();
      }
      acc4[0] = acc4[1] = acc4[2] = acc4[3] = 0.f;
#pragma unroll
      for (int r = 0; r < R; r += 2) {
        acc4[(r / 2) & 3] = fmaf(bf16_lo(up[r / 2]), Y[r], acc4[(r / 2) & 3]);
        acc4[(r / 2) & 3] = fmaf(bf16_hi(up[r / 2]), Y[r + 1], acc4[(r / 2) & 3]);
      }
      const float Dn = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
      const bool ok = Dn > 1e-30f;
      const float rden = ok ? rsqrtf(Dn) : 0.f;
      const float cosv = num * rden;
      const float gam = g * (1.f / kW);
      const float c1 = gam * rden;                     // d cos / d num
      const float kap = cosv * rden;                   // cos / den
      const float dD = -0.5f * gam * cosv * rden * rden;
      const float lc1 = a.lam * c1;
      // pass B1: UD1 = u dD (operand of dG^V);  dA = lambda c1 u (S - kappa Y)   (kept in Y[])
#pragma unroll
      for (int r = 0; r < R; r += 2) {
        const float ua = bf16_lo(up[r / 2]), ub = bf16_hi(up[r / 2]);
        st_shared_u16(t_addr(sbase + L::A2, r), bf16_bits(ua * dD));
        st_shared_u16(t_addr(sbase + L::A2, r + 1), bf16_bits(ub * dD));
        Y[r] = lc1 * ua * fmaf(-kap, Y[r], S[r]);
        Y[r + 1] = lc1 * ub * fmaf(-kap, Y[r + 1], S[r + 1]);
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(bar(UD1_READY));
      TRA(6);
      // ---- t_r = sum over this text's words of A dA  (l2-norm backward, model/loss.py:238)
      {
        float tot[(R + 31) / 32];
#pragma unroll
        for (int b = 0; b < R / 32; ++b) {
          float v[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int r = b * 32 + c;
            v[c] = leaky(S[r]) * s_invnA[q * R + r] * Y[r];
          }
          tot[b] = lanes_to_columns32(v, lane);
        }
        if (R % 32) {
          float v[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int r = (R / 32) * 32 + c;
            v[c] = leaky(S[r]) * s_invnA[q * R + r] * Y[r];
          }
          tot[R / 32] = lanes_to_columns16(v, lane);
        }
#pragma unroll
        for (int b = 0; b < R / 32; ++b) s_tA[q * R + b * 32 + lane] = tot[b];
        if ((R % 32) && lane < 16) s_tA[q * R + (R / 32) * 32 + lane] = tot[R / 32];
        __syncwarp();
      }
      TRA(7);
      mbar_wait(bar(DGV_DONE), ph);                    // U1^T / UD1^T consumed: A1 may take P1^T
      TRA(8);
      // pass B3: dS = (dA - A t) / n * leaky'(S) + c1 u ;  P = dS alpha_r beta_w
#pragma unroll
      for (int r = 0; r < R; r += 4) {
        const float4 na = *reinterpret_cast<const float4*>(s_invnA + q * R + r);
        const float4 tt = *reinterpret_cast<const float4*>(s_tA + q * R + r);
        const float4 al = *reinterpret_cast<const float4*>(s_alpha + r);
        const float nav[4] = {na.x, na.y, na.z, na.w}, ttv[4] = {tt.x, tt.y, tt.z, tt.w}, alv[4] = {al.x, al.y, al.z, al.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int rr = r + h;
          const float s = S[rr];
          const float Av = leaky(s) * nav[h];
          const float u = (rr & 1) ? bf16_hi(up[rr / 2]) : bf16_lo(up[rr / 2]);
          float dS = (Y[rr] - Av * ttv[h]) * nav[h] * leaky_grad(s) + c1 * u;
          st_shared_u16(t_addr(sbase + L::A1, rr), bf16_bits(dS * alv[h] * beta_m));
        }
      }
      // dG^T block of this text: rows = this word, 32 columns
      mbar_wait(bar(DGT_DONE), ph);
      tc_fence_after();
      {
        float gt[32];
        tmem_ld16(tmem + cDGT + lane_off + q * 32, reinterpret_cast<uint32_t*>(gt));
        tmem_ld16(tmem + cDGT + lane_off + q * 32 + 16, reinterpret_cast<uint32_t*>(gt) + 16);
        tmem_ld_wait();
        if (j < a.Bt) {
          float* dst = a.dgt + ((size_t)jt * kRows + m) * kW;
#pragma unroll
          for (int c = 0; c < 32; c += 4) red_add_v4(dst + c, gt[c], gt[c + 1], gt[c + 2], gt[c + 3]);
        }
      }
      // dG^V rows (TMEM lanes = regions: this warp's quadrant holds regions 32q..32q+31); group B is the
      // critical path of the tile, group A has slack
      if (q * 32 < R) {
        float* dst = a.dgv + ((size_t)i * R + m) * R;
#pragma unroll 1
        for (int c = 0; c < R; c += 16) {
          float v[16];
          tmem_ld16(tmem + cS1 + lane_off + c, reinterpret_cast<uint32_t*>(v));
          tmem_ld_wait();
          if (m < R) {
#pragma unroll
            for (int x4 = 0; x4 < 16; x4 += 4) red_add_v4(dst + c + x4, v[x4], v[x4 + 1], v[x4 + 2], v[x4 + 3]);
          }
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(bar(P_READY));
      TRA(9);
    }
  } else {
    // ================= group B: t2i direction, thread = region row =================
    // two halves (warps 8-11: texts 0,1; warps 12-15: texts 2,3): twice the resident warps, half the chain
    asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
    const int hb = (warp >= 12) ? 1 : 0, t_lo = hb * 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const bool warp_active = q * 32 < R;
    const bool active = r < R;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t row_off = (uint32_t)(r >> 3) * 1024 + (uint32_t)(r & 7) * 128;
    auto seg_addr = [&](uint32_t buf, int t, int g4) {     // 16-byte segment g4 (0..3) of text t in this row
      const uint32_t seg = (uint32_t)((t & 1) * 4 + g4);
      return buf + (uint32_t)(t >> 1) * L::kVChunk + row_off + ((seg ^ (uint32_t)(r & 7)) << 4);
    };
    int vloads = 0;
    for (int k = 0; k < n_tiles; ++k) {
      const uint32_t ph = k & 1;
      const int i = video_of(k), jt = tile_of(k);
      TRB(0);
      if (is_switch(k)) { mbar_wait(bar(V_FULL), vloads & 1); ++vloads; }
      mbar_wait(bar(GT_FULL), ph);
      mbar_wait(bar(S_DONE), ph);
      tc_fence_after();
      TRB(1);
      const float alpha_r = active ? s_alpha[r] : 0.f;
      float num[2], c1v[2], kapv[2];
      TRB(3);
      if (k > 0) mbar_wait(bar(P_STORED), (k - 1) & 1);   // TMA stores of tile k-1 have read A1 / B1
      if (warp_active) {
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          const int t = t_lo + tt;
          float x[32], e[32];
          tmem_ld16(tmem + cS2 + lane_off + t * 32, reinterpret_cast<uint32_t*>(x));
          tmem_ld16(tmem + cS2 + lane_off + t * 32 + 16, reinterpret_cast<uint32_t*>(x) + 16);
          tmem_ld_wait();
          float e4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int w = 0; w < 32; w += 4) {
            const float4 be = *reinterpret_cast<const float4*>(s_beta + t * 32 + w);
            const float4 nb = *reinterpret_cast<const float4*>(s_invnB + t * 32 + w);
            const float4 mk = *reinterpret_cast<const float4*>(s_m2t + t * 32 + w);
            x[w + 0] *= alpha_r * be.x; x[w + 1] *= alpha_r * be.y; x[w + 2] *= alpha_r * be.z; x[w + 3] *= alpha_r * be.w;
            e[w + 0] = ex2(fmaf(leaky(x[w + 0]) * nb.x, a.lam2, mk.x));
            e[w + 1] = ex2(fmaf(leaky(x[w + 1]) * nb.y, a.lam2, mk.y));
            e[w + 2] = ex2(fmaf(leaky(x[w + 2]) * nb.z, a.lam2, mk.z));
            e[w + 3] = ex2(fmaf(leaky(x[w + 3]) * nb.w, a.lam2, mk.w));
#pragma unroll
            for (int h = 0; h < 4; ++h) e4[h] += e[w + h];
          }
          const float thr = a.focal ? ((e4[0] + e4[1]) + (e4[2] + e4[3])) * (1.f / kW) : -1.f;
          float n4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            uint32_t pk[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int w = g4 * 8 + h * 2;
              pk[h] = pack_bf16(e[w] > thr ? e[w] : 0.f, e[w + 1] > thr ? e[w + 1] : 0.f);
              n4[h] = fmaf(bf16_lo(pk[h]), x[w], n4[h]);
              n4[h] = fmaf(bf16_hi(pk[h]), x[w + 1], n4[h]);
            }
            if (active) st_shared_v4(seg_addr(sbase + L::B1, t, g4), pk[0], pk[1], pk[2], pk[3]);
          }
          num[tt] = (n4[0] + n4[1]) + (n4[2] + n4[3]);
        }
      }
      fence_proxy_async();
      mbar_arrive(bar(U2_READY));
      TRB(4);
      // ---- loop A: denominators, per-(region, text) scalars, UD2 = u dD
      mbar_wait(bar(Y2_DONE), ph);
      tc_fence_after();
      TRB(5);
      if (warp_active) {
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          const int t = t_lo + tt;
          const int j = jt * kTPT + t;
          const float g = j < a.Bt ? a.d_scores[(size_t)i * a.Bt + j] : 0.f;
          float y[32];
          tmem_ld16(tmem + cY2 + lane_off + t * 32, reinterpret_cast<uint32_t*>(y));
          tmem_ld16(tmem + cY2 + lane_off + t * 32 + 16, reinterpret_cast<uint32_t*>(y) + 16);
          tmem_ld_wait();
          uint32_t pk[16];
          float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            ld_shared_v4(seg_addr(sbase + L::B1, t, g4), pk[g4 * 4], pk[g4 * 4 + 1], pk[g4 * 4 + 2], pk[g4 * 4 + 3]);
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              d4[h] = fmaf(bf16_lo(pk[g4 * 4 + h]), y[g4 * 8 + h * 2], d4[h]);
              d4[h] = fmaf(bf16_hi(pk[g4 * 4 + h]), y[g4 * 8 + h * 2 + 1], d4[h]);
            }
          }
          const float Dn = (d4[0] + d4[1]) + (d4[2] + d4[3]);
          const bool ok = active && Dn > 1e-30f;
          const float rden = ok ? rsqrtf(Dn) : 0.f;
          const float cosv = num[tt] * rden;
          const float gam = g * (1.f / R);
          c1v[tt] = ok ? gam * rden : 0.f;
          kapv[tt] = ok ? cosv * rden : 0.f;
          const float dD = ok ? -0.5f * gam * cosv * rden * rden : 0.f;
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            uint32_t o[4];
#pragma unroll
            for (int h = 0; h < 4; ++h)
              o[h] = pack_bf16(bf16_lo(pk[g4 * 4 + h]) * dD, bf16_hi(pk[g4 * 4 + h]) * dD);
            if (active) st_shared_v4(seg_addr(sbase + L::B2, t, g4), o[0], o[1], o[2], o[3]);
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(bar(UD2_READY));
      TRB(6);
      // ---- loop B: per text, dA, the cross-region sum t_w = sum_r A dA, dS, P2
#pragma unroll 1
      for (int tt = 0; tt < 2; ++tt) {
        const int t = t_lo + tt;
        float x[32], dA[32];
        uint32_t pk[16];
        const float c1 = tt ? c1v[1] : c1v[0], kap = tt ? kapv[1] : kapv[0];
        if (warp_active) {
          float y[32];
          tmem_ld16(tmem + cS2 + lane_off + t * 32, reinterpret_cast<uint32_t*>(x));
          tmem_ld16(tmem + cS2 + lane_off + t * 32 + 16, reinterpret_cast<uint32_t*>(x) + 16);
          tmem_ld16(tmem + cY2 + lane_off + t * 32, reinterpret_cast<uint32_t*>(y));
          tmem_ld16(tmem + cY2 + lane_off + t * 32 + 16, reinterpret_cast<uint32_t*>(y) + 16);
          tmem_ld_wait();
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4)
            ld_shared_v4(seg_addr(sbase + L::B1, t, g4), pk[g4 * 4], pk[g4 * 4 + 1], pk[g4 * 4 + 2], pk[g4 * 4 + 3]);
          const float lc1 = a.lam * c1;
          float v[32];
#pragma unroll
          for (int w = 0; w < 32; ++w) {
            x[w] *= alpha_r * s_beta[t * 32 + w];
            const float u = (w & 1) ? bf16_hi(pk[w / 2]) : bf16_lo(pk[w / 2]);
            dA[w] = lc1 * u * fmaf(-kap, y[w], x[w]);
            v[w] = active ? leaky(x[w]) * s_invnB[t * 32 + w] * dA[w] : 0.f;
          }
          const float part = lanes_to_columns32(v, lane);
          s_tBp[(t * 4 + q) * 32 + lane] = part;
        } else {
          s_tBp[(t * 4 + q) * 32 + lane] = 0.f;
        }
        named_bar_sync(3 + hb, 128);                   // the four region-quadrant warps of this half
        if (tt == 0) mbar_wait(bar(DGT_DONE), ph);     // U2 / UD2 consumed: B1 may take P2
        if (warp_active) {
          const float* tb = s_tBp + t * 128;
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            uint32_t o[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              float pv[2];
#pragma unroll
              for (int z = 0; z < 2; ++z) {
                const int w = g4 * 8 + h * 2 + z;
                const float tw = (tb[w] + tb[32 + w]) + (tb[64 + w] + tb[96 + w]);
                const float nb = s_invnB[t * 32 + w];
                const float Av = leaky(x[w]) * nb;
                const float u = z ? bf16_hi(pk[g4 * 4 + h]) : bf16_lo(pk[g4 * 4 + h]);
                const float dS = (dA[w] - Av * tw) * nb * leaky_grad(x[w]) + c1 * u;
                pv[z] = dS * alpha_r * s_beta[t * 32 + w];
              }
              o[h] = pack_bf16(pv[0], pv[1]);
            }
            if (active) st_shared_v4(seg_addr(sbase + L::B1, t, g4), o[0], o[1], o[2], o[3]);
          }
        }
      }
      TRB(7);
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(bar(P_READY));
      TRB(9);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
static int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return DVLP_ERR_CUDA;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return DVLP_ERR_CUDA; }
  return DVLP_OK;
}

template <int R>
static int launch_bwd(const void* vb, const void* tb, int Bv, int Bt, float lambda, int focal, const float* d_scores,
                      char* ws, const Workspace& Lw, cudaStream_t st) {
  CUtensorMap mV, mT, mGV, mGT, mPa, mPb;
  int rc;
  Args a{};
  a.Bv = Bv; a.Bt = Bt;
  a.tiles_per_video = (Bt + kTPT - 1) / kTPT;
  a.total_units = (long long)Bv * a.tiles_per_video;
  if ((rc = make_map(&mV, vb, (uint64_t)Bv * R, kD, R))) return rc;
  if ((rc = make_map(&mT, tb, (uint64_t)Bt * kW, kD, kRows))) return rc;
  if ((rc = make_map(&mGV, ws + Lw.gvb, (uint64_t)Bv * R, R, R))) return rc;
  if ((rc = make_map(&mGT, ws + Lw.gtb, (uint64_t)Bt * kW, kW, kRows))) return rc;
  // P spill: [unit][region][128 words] bf16, one R-row box per 64-word half
  if ((rc = make_map(&mPa, ws + Lw.pa, (uint64_t)a.total_units * R, kRows, R))) return rc;
  if ((rc = make_map(&mPb, ws + Lw.pb, (uint64_t)a.total_units * R, kRows, R))) return rc;
  a.alpha = (const float*)(ws + Lw.alpha); a.beta = (const float*)(ws + Lw.beta);
  a.m2v = (const float*)(ws + Lw.m2v); a.m2t = (const float*)(ws + Lw.m2t);
  a.d_scores = d_scores;
  a.lam = lambda; a.lam2 = lambda * kLog2e; a.focal = focal;
  a.dgv = (float*)(ws + Lw.dgv); a.dgt = (float*)(ws + Lw.dgt);
  a.nrm_a = (const float*)(ws + Lw.nrm_a); a.nrm_b = (const float*)(ws + Lw.nrm_b);
  a.trace = (long long*)g_trace_buffer;
  int dev = 0, sms = 148;
  DVLP_CUDA_OK(cudaGetDevice(&dev));
  DVLP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = (int)(a.total_units < sms ? a.total_units : sms);
  static bool attr_set = false;
  if (!attr_set) {
    DVLP_CUDA_OK(cudaFuncSetAttribute(k_pair_tc_bwd<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<R>::kAlloc));
    attr_set = true;
  }
  k_pair_tc_bwd<R><<<grid, kThreads, Lay<R>::kAlloc, st>>>(mV, mT, mGV, mGT, mPa, mPb, a);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}
}  // namespace tcb

bool tc_bwd_available() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DVLP_TC_BWD");     // DVLP_TC_BWD=0: tcgen05 forward + generic backward (debug A/B)
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

// DVLP_TC_BWD=1: first tcgen05 backward (this file + grad_gemm.cu); default 2: pair_bwd2.cu + grad_gemm2.cu
int tc_bwd_version() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DVLP_TC_BWD");
    v = (e && e[0] == '1') ? 1 : 2;
  }
  return v;
}

// DVLP_TC_FWD=1: first tcgen05 forward (pair_tc.cu); default 2: pair_fwd2.cu
int tc_fwd_version() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DVLP_TC_FWD");
    v = (e && e[0] == '1') ? 1 : 2;
  }
  return v;
}

int run_pair_tc_bwd(const void* vb, const void* tb, int Bv, int Bt, int R, int W, int d, float lambda, int focal,
                    const float* d_scores, char* ws, const Workspace& L, cudaStream_t st) {
  if (d != tcb::kD || W != tcb::kW) return DVLP_ERR_BAD_SHAPE;
  switch (R) {
    case 80: return tcb::launch_bwd<80>(vb, tb, Bv, Bt, lambda, focal, d_scores, ws, L, st);
    case 64: return tcb::launch_bwd<64>(vb, tb, Bv, Bt, lambda, focal, d_scores, ws, L, st);
    case 48: return tcb::launch_bwd<48>(vb, tb, Bv, Bt, lambda, focal, d_scores, ws, L, st);
    case 32: return tcb::launch_bwd<32>(vb, tb, Bv, Bt, lambda, focal, d_scores, ws, L, st);
  }
  return DVLP_ERR_BAD_SHAPE;
}

}  // namespace dvlp
