// Region-word pair scores on tcgen05 / TMEM / TMA (sm_100a): BACKWARD kernel, single-orientation design.
//
// Back-propagates d_scores through both attention directions of model/loss.py:209-330 for one
// (video, 4-text tile) unit at a time, recomputing the attention from the raw similarities (nothing of size
// Bv*Bt*R*W is ever stored).  Everything lives in ONE orientation: TMEM lane = word row (4 texts x 32 words),
// TMEM column = region.  The R columns are split into 16-column chunks, one compute warpgroup per chunk
// (R/16 warpgroups, all running the same code), so every element (word, region) is owned by exactly one
// thread which evaluates BOTH directions for it and emits ONE gradient tile
//     P = (dS_i2t + dS_t2i) * alpha_r * beta_w                      [128 words][R regions] bf16
// that grad_gemm.cu contracts against T / V.  Sums over regions (softmax axis of the i2t direction, l2-norm
// axis of t2i) are row sums: per-thread partials exchanged between the warpgroups through shared memory.
// Sums over the words of a text (softmax axis of t2i, l2-norm axis of i2t) are sums over the 32 lanes of
// one warp: shuffle transpose-reductions.  The per-query scalars of the forward (attention l2 norms, focal
// thresholds, 1/|c|, cosine) are not recomputed: pair_tc.cu saves them per unit (rowv / colv).
//
// Tensor-core products per unit (all operands in shared memory, accumulators in TMEM):
//   S   = T V^T                     M=128 words, N=R, K=256      raw similarities
//   Y1  = U1 G^V                    M=128, N=R, K=R              i2t: Gram-form |weighted context|^2
//   Y2  = Gbd U2                    M=128, N=R, K=128            t2i: Gbd = block-diagonal text Gram matrices
//   dGV += UD1^T U1                 M=R(128), N=R, K=128         accumulated in TMEM over the tiles of a video
//   dGT = UD2 U2^T                  M=128, N=128, K=R            diagonal 32x32 blocks used
// U / UD / P tiles are [word rows][region columns] in the 32-byte-swizzle layout (one 4 KB chunk per 16
// columns): written row-wise by their owner threads, consumed K-major or MN-major through the UMMA
// descriptors (encodings checked by tools/ubench/desc_lab.cu).
//
// Warp roles: the compute warps come first (warpgroup g = column chunk, warp q of the group = text slot = TMEM lane
// quadrant), the control warpgroup last (TMA producer, MMA issuer, P-tile TMA stores, TMEM owner).
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace dvlp {
using namespace ptx;

extern void* g_trace_buffer;   // api.cu (dvlp_debug_set_trace)

namespace tb2 {
constexpr int kD = 256, kW = 32, kTPT = 4, kRows = 128;
constexpr int kTileBytes = 16384;   // 128 rows x 128 B (one 64-column chunk of a 128-byte-swizzle tile)
constexpr int kCh32 = 4096;         // 128 rows x 32 B  (one 16-column chunk of a 32-byte-swizzle tile)

template <int R>
struct Lay {
  static constexpr int NC = R / 16;                 // column chunks = compute warpgroups
  static constexpr int NW = 4 * NC;                 // compute warps
  static constexpr int kThreads = 128 * (1 + NC);
  static constexpr int kVChunk = R * 128;
  static constexpr int V = 0;                       // 4 chunks (K = 256)
  static constexpr int GV = V + 4 * kVChunk;        // 2 chunks
  static constexpr int T = GV + 2 * kVChunk;        // 4 x 16 KB; dead after the S GEMM: UD1 / UD2 alias it
  static constexpr int UD1 = T, UD2 = T + 2 * kTileBytes;
  static constexpr int GBD = T + 4 * kTileBytes;    // 2 x 16 KB
  static constexpr int U1 = GBD + 2 * kTileBytes;   // NC x 4 KB; re-used for P
  static constexpr int U2 = U1 + NC * kCh32;
  static constexpr int ROWV = U2 + NC * kCh32;      // [128] float4 (nB, thr1, rden1, cos1)
  static constexpr int COLV = ROWV + 2048;          // [4][R] float4 (nA, thr2, rden2, cos2)
  static constexpr int BETA = COLV + 4 * R * 16;    // [128]
  static constexpr int M2T = BETA + 512;            // [128]
  static constexpr int ALPHA = M2T + 512;           // [R]
  static constexpr int M2V = ALPHA + R * 4;         // [R]
  static constexpr int TAB = M2V + R * 4;           // per compute warp 12 x [16] floats (enum Tab), 768 B
  static constexpr int TBP = TAB + NW * 768;        // [NC][128] partial row sums
  static constexpr int BARS = TBP + NC * 512;
  static constexpr int TMEMPTR = BARS + 16 * 8;
  static constexpr int kAlloc = TMEMPTR + 16;
  static_assert(NC * kCh32 <= 2 * kTileBytes, "UD tiles must fit the T alias");
  static_assert(kAlloc <= 232448, "shared memory budget");
};

// per-warp tables of the warp's 16 columns, one float array per quantity (adjacent columns load as packed pairs)
enum Tab { T_NAL = 0, T_THR2, T_ALPHA, T_M2V,   // phase 1: nA lam2, thr2, alpha, m2v
           T_LC1B, T_NKAPB, T_NA,               // phase 2: lambda c1B, -kapB, nA
           T_DDB,                               // UD2 operand: -0.5 c1B kapB
           T_NNATA,                             // phase 3: -(nA^3 tA), written after phase 2
           kNumTab };
static_assert(kNumTab <= 12, "768 bytes per warp");

enum Bar { V_FULL = 0, T_FULL, GBD_FULL, VEC_FULL, VEC_FREE, S_DONE, U_READY, Y_DONE, DG_DONE, P_READY, P_STORED, UD_READY,
           kNumBars };

// TMEM columns
constexpr uint32_t cS = 0, cY1 = 96, cY2 = 192, cDGV = 288, cDGT = 384;

struct Args {
  int Bv, Bt, tiles_per_video;
  long long total_units;
  const float *alpha, *beta, *m2v, *m2t, *d_scores;
  const float4 *rowv, *colv;     // per-unit scalars saved by the forward kernel
  float lam, lam2;
  float *dgv, *dgt;
  long long* trace;   // debug: per-phase clock64 stamps of CTA 0 (tools/trace_tiles.py), NULL in production
};
// trace slot layout: [role 0..4][tile 0..31][event 0..15]; roles: 1 = MMA warp, 2 = compute warp (g0,q0), 3 = (g4,q3)
#define TR2(role, ev)                                                                        \
  do {                                                                                      \
    if (TRACE && a.trace && blockIdx.x == 0 && lane == 0 && k < 32) a.trace[((role) * 32 + k) * 16 + (ev)] = clock64(); \
  } while (0)
#define TRC(ev) do { if (cw == 0) TR2(2, ev); else if (cw == NW - 1) TR2(3, ev); } while (0)

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint32_t desc_hi32(uint32_t sbo_bytes) {      // 32-byte swizzle
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (6u << 29);
}
__device__ __forceinline__ float2 shfl_xor2(float2 v, int m) {
  return make_float2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
__device__ __forceinline__ float2 sel2(bool c, float2 a, float2 b) { return make_float2(c ? a.x : b.x, c ? a.y : b.y); }
// the two bf16 halves of a packed word as an fp32 pair
__device__ __forceinline__ float2 bf16_pair(uint32_t w) { return make_float2(bf16_lo(w), bf16_hi(w)); }
__device__ __forceinline__ float2 focal2(float2 e, float2 thr) { return fmul2(e, gt_mask2(e, thr)); }
// x > 0 ? pos : neg from the 1.0 / 0.0 compare mask
__device__ __forceinline__ float2 pos_sel2(float2 x, float2 pos, float2 neg) {
  return ffma2(gt_mask2(x, make_float2(0.f, 0.f)), make_float2(pos.x - neg.x, pos.y - neg.y), neg);
}

template <int R, bool TRACE>
__global__ void __launch_bounds__(Lay<R>::kThreads, 1)
k_pair_bwd2(const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapT,
            const __grid_constant__ CUtensorMap mapGV, const __grid_constant__ CUtensorMap mapGBD,
            const __grid_constant__ CUtensorMap mapP, Args a) {
  using L = Lay<R>;
  constexpr int NC = L::NC, NW = L::NW;
  static_assert(R % 16 == 0 && R >= 32 && R <= 80, "R must be a multiple of 16 in [32,80]");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if (sbase & 1023u) __trap();
  uint8_t* sgen = smem_raw;
  volatile uint32_t* s_tmem = (volatile uint32_t*)(sgen + L::TMEMPTR);
  auto bar = [&](int i) { return sbase + L::BARS + 8 * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // The warp scheduler prefers the highest warp id among its ready warps, so the control warps are the LAST
  // warpgroup (see pair_fwd2.cu).  role: 0 TMA producer, 1 MMA issuer, 2 stores, 3 TMEM owner; -1 compute.
  const int role = warp >= NW ? warp - NW : -1, cw = warp;
  const long long u0 = a.total_units * blockIdx.x / gridDim.x;
  const long long u1 = a.total_units * (blockIdx.x + 1) / gridDim.x;
  const int n_tiles = (int)(u1 - u0);
  const int i0 = (int)(u0 / a.tiles_per_video), jt0 = (int)(u0 % a.tiles_per_video);
  const int tpv = a.tiles_per_video;

  if (role == 0 && lane == 0) {
    prefetch_tmap(&mapV); prefetch_tmap(&mapT); prefetch_tmap(&mapGV); prefetch_tmap(&mapGBD); prefetch_tmap(&mapP);
    for (int b = 0; b < kNumBars; ++b) {
      uint32_t cnt = 1;
      if (b == VEC_FREE || b == U_READY || b == P_READY || b == UD_READY) cnt = NW;
      mbar_init(bar(b), cnt);
    }
    fence_mbar_init();
  }
  if (role == 3) {
    tmem_alloc(sbase + L::TMEMPTR, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (role >= 0) {
    if (NC == 5) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (role == 0) {
      // =============================== TMA producer ===============================
      if (elect_one()) {
        int i = i0, jt = jt0;
        for (int k = 0; k < n_tiles; ++k) {
          const uint32_t pp = (k - 1) & 1;
          const long long unit = u0 + k;
          // per-unit / per-tile vectors: the staging area is free once every warp has built its tables
          if (k > 0) mbar_wait(bar(VEC_FREE), pp);
          mbar_arrive_expect_tx(bar(VEC_FULL), 2048 + 4 * R * 16 + 1024);
          bulk_load(sbase + L::ROWV, a.rowv + unit * kRows, 2048, bar(VEC_FULL));
          bulk_load(sbase + L::COLV, a.colv + unit * (kTPT * R), 4 * R * 16, bar(VEC_FULL));
          bulk_load(sbase + L::BETA, a.beta + (size_t)jt * kRows, 512, bar(VEC_FULL));
          bulk_load(sbase + L::M2T, a.m2t + (size_t)jt * kRows, 512, bar(VEC_FULL));
          // text tile: its buffer doubles as UD1 / UD2 until the dG GEMMs of the previous tile are done
          if (k > 0) mbar_wait(bar(DG_DONE), pp);
          mbar_arrive_expect_tx(bar(T_FULL), 4 * kTileBytes);
          for (int c = 0; c < 4; ++c) tma_load_2d(sbase + L::T + c * kTileBytes, &mapT, c * 64, jt * kRows, bar(T_FULL));
          // block-diagonal text Gram tile: the Y2 GEMM of the previous tile has read it
          if (k > 0) mbar_wait(bar(Y_DONE), pp);
          mbar_arrive_expect_tx(bar(GBD_FULL), 2 * kTileBytes);
          for (int c = 0; c < 2; ++c) tma_load_2d(sbase + L::GBD + c * kTileBytes, &mapGBD, c * 64, jt * kRows, bar(GBD_FULL));
          if (k == 0 || jt == 0) {
            // the resident video changes: everybody is done with tile k-1
            if (k > 0) mbar_wait(bar(P_READY), pp);
            mbar_arrive_expect_tx(bar(V_FULL), 6 * L::kVChunk + 2 * R * 4);
            for (int c = 0; c < 4; ++c) tma_load_2d(sbase + L::V + c * L::kVChunk, &mapV, c * 64, i * R, bar(V_FULL));
            for (int c = 0; c < 2; ++c) tma_load_2d(sbase + L::GV + c * L::kVChunk, &mapGV, c * 64, i * R, bar(V_FULL));
            bulk_load(sbase + L::ALPHA, a.alpha + (size_t)i * R, R * 4, bar(V_FULL));
            bulk_load(sbase + L::M2V, a.m2v + (size_t)i * R, R * 4, bar(V_FULL));
          }
          if (++jt == tpv) { jt = 0; ++i; }
        }
      }
    } else if (role == 1) {
      // ================================ MMA issuer ================================
      constexpr uint32_t idS = idesc_bf16(128, R), idY2 = idesc_bf16(128, R, 0, 1), idDGV = idesc_bf16(128, R, 1, 1);
      constexpr uint32_t idDGT = idesc_bf16(128, 128);
      constexpr uint32_t CH = L::kVChunk;
      const uint32_t hi128 = desc_hi(1024), hi32 = desc_hi32(256);
      const uint32_t loT = desc_lo(sbase + L::T), loV = desc_lo(sbase + L::V), loGV = desc_lo(sbase + L::GV);
      const uint32_t loGBD = desc_lo(sbase + L::GBD);
      const uint32_t loU1k = desc_lo(sbase + L::U1), loU2k = desc_lo(sbase + L::U2), loUD2k = desc_lo(sbase + L::UD2);
      const uint32_t loU1m = desc_lo(sbase + L::U1, kCh32), loU2m = desc_lo(sbase + L::U2, kCh32);
      const uint32_t loUD1m = desc_lo(sbase + L::UD1, kCh32);
      int vloads = 0, jt = jt0;
      auto issue_S = [&](int k, bool sw) {
        if (sw) { mbar_wait(bar(V_FULL), vloads & 1); ++vloads; }
        mbar_wait(bar(T_FULL), k & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {
            const uint32_t oT = ((ks >> 2) * kTileBytes + (ks & 3) * 32) >> 4, oV = ((ks >> 2) * CH + (ks & 3) * 32) >> 4;
            mma_bf16_w(tmem + cS, loT + oT, hi128, loV + oV, hi128, idS, ks != 0);
          }
          tc_commit(bar(S_DONE));
        }
        __syncwarp();
      };
      issue_S(0, true);
      for (int k = 0; k < n_tiles; ++k) {
        const uint32_t ph = k & 1;
        const bool first_of_video = (k == 0 || jt == 0);
        // ---- Y1 = U1 G^V,  Y2 = Gbd U2
        mbar_wait(bar(GBD_FULL), ph);
        mbar_wait(bar(U_READY), ph);
        tc_fence_after();
        TR2(1, 0);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < NC; ++ks)
            mma_bf16_w(tmem + cY1, loU1k + ((ks * kCh32) >> 4), hi32, loGV + (((ks >> 2) * CH + (ks & 3) * 32) >> 4), hi128,
                       idS, ks != 0);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            mma_bf16_w(tmem + cY2, loGBD + (((ks >> 2) * kTileBytes + (ks & 3) * 32) >> 4), hi128, loU2m + ((ks * 512) >> 4),
                       hi32, idY2, ks != 0);
          tc_commit(bar(Y_DONE));
        }
        __syncwarp();
        // ---- dGV += UD1^T U1 (K = 128 words),  dGT = UD2 U2^T (K = R regions): UD is written under the Y GEMMs
        mbar_wait(bar(UD_READY), ph);
        tc_fence_after();
        TR2(1, 1);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            mma_bf16_w(tmem + cDGV, loUD1m + ((ks * 512) >> 4), hi32, loU1m + ((ks * 512) >> 4), hi32, idDGV,
                       !(first_of_video && ks == 0));
#pragma unroll
          for (int ks = 0; ks < NC; ++ks)
            mma_bf16_w(tmem + cDGT, loUD2k + ((ks * kCh32) >> 4), hi32, loU2k + ((ks * kCh32) >> 4), hi32, idDGT, ks != 0);
          tc_commit(bar(DG_DONE));
        }
        __syncwarp();
        TR2(1, 3);
        if (++jt == tpv) jt = 0;
        if (k + 1 < n_tiles) issue_S(k + 1, jt == 0);
        TR2(1, 4);
      }
    } else if (role == 2) {
      // ====== stores: the bf16 gradient tile P of each unit (TMA tile store) ======
      if (elect_one()) {
        for (int k = 0; k < n_tiles; ++k) {
          mbar_wait(bar(P_READY), k & 1);
          const int row = (int)((u0 + k) * kRows);
#pragma unroll
          for (int c = 0; c < NC; ++c) tma_store_2d(&mapP, sbase + L::U1 + c * kCh32, c * 16, row);
          bulk_commit();
          bulk_wait_read0();                 // U1 may take the next tile's u
          mbar_arrive(bar(P_STORED));
        }
        bulk_wait_all();
      }
    }
  } else {
    // ============================ compute warps ============================
    if (NC == 5) asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");   // the pool holds what the control warps released: 4 x 48 = 20 x 8 + spare
    const int g = cw >> 2, q = cw & 3;          // column chunk, text slot / TMEM lane quadrant
    const int m = q * 32 + lane;                // word row of the tile
    const int cl = lane & 15;                   // table column handled by this lane
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t sw = (uint32_t)((m >> 2) & 1) << 4;
    const uint32_t row_off = (uint32_t)g * kCh32 + (uint32_t)m * 32;
    const uint32_t seg0 = row_off + sw, seg1 = row_off + (16u ^ sw);
    float* tab = reinterpret_cast<float*>(sgen + L::TAB + cw * 768);
    const float4* tab4 = reinterpret_cast<const float4*>(tab);
    const float4* s_rowv = reinterpret_cast<const float4*>(sgen + L::ROWV);
    const float4* s_colv = reinterpret_cast<const float4*>(sgen + L::COLV);
    const float* s_beta = reinterpret_cast<const float*>(sgen + L::BETA);
    const float* s_m2t = reinterpret_cast<const float*>(sgen + L::M2T);
    const float* s_alpha = reinterpret_cast<const float*>(sgen + L::ALPHA);
    const float* s_m2v = reinterpret_cast<const float*>(sgen + L::M2V);
    float* s_tbp = reinterpret_cast<float*>(sgen + L::TBP);
    int vloads = 0, i = i0, jt = jt0;
    auto load_g = [&](int vi, int tj) {         // d_scores of this warp's (video, text); 0 for a padded text
      const int jj = tj * kTPT + q;
      float v = 0.f;
      if (jj < a.Bt) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(a.d_scores + (size_t)vi * a.Bt + jj));
      return v;
    };
    float gs_next = load_g(i, jt);
    // The Gram gradients of a tile leave TMEM one tile late, in the shadow of the next tile's Y GEMMs (the warps would
    // only wait there): dG^T block of this warp's text straight from registers with red.global (16-byte chunks of a row
    // XOR-swizzled with the row index - the permutation the global dG^T buffer has always kept and epilogue.cu undoes,
    // dgt_swz), dG^V rows after the last tile of a video.  Both are read before this warp signals UD_READY, i.e.
    // before the next dG GEMMs can touch the accumulators.
    int jt_prev = 0, i_prev = 0;
    bool last_prev = false;
    auto flush_gram = [&]() {
      tc_fence_after();
      if (g < 2 && jt_prev * kTPT + q < a.Bt) {
        float v[16];
        tmem_ld16(tmem + cDGT + lane_off + q * 32 + g * 16, reinterpret_cast<uint32_t*>(v));
        tmem_ld_wait();
        float* dst = a.dgt + ((size_t)jt_prev * kRows + m) * kW;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          red_add_v4(dst + (((g * 4 + c) ^ (lane & 7)) << 2), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      }
      if (last_prev && q * 32 < R) {
        // dG^V rows (TMEM lanes = regions), this warpgroup's 16 columns
        float v[16];
        tmem_ld16(tmem + cDGV + lane_off + g * 16, reinterpret_cast<uint32_t*>(v));
        tmem_ld_wait();
        if (m < R) {
          float* dst = a.dgv + ((size_t)i_prev * R + m) * R + g * 16;
#pragma unroll
          for (int c = 0; c < 16; c += 4) red_add_v4(dst + c, v[c], v[c + 1], v[c + 2], v[c + 3]);
        }
      }
      tc_fence_before();
    };
    for (int k = 0; k < n_tiles; ++k) {
      const uint32_t ph = k & 1;
      const bool last_of_video = (k == n_tiles - 1) || (jt == tpv - 1);
      TRC(11);
      const float gs = gs_next;
      if (k + 1 < n_tiles) gs_next = (jt + 1 == tpv) ? load_g(i + 1, 0) : load_g(i, jt + 1);   // one tile ahead
      TRC(0);
      if (k == 0 || jt == 0) { mbar_wait(bar(V_FULL), vloads & 1); ++vloads; }
      mbar_wait(bar(VEC_FULL), ph);
      TRC(1);
      // ---- per-row scalars (this thread's word) and the per-column tables of this warp's 16 regions
      const float4 rv = s_rowv[m];              // nB, thr1, rden1, cos1
      const float beta_m = s_beta[m], m2t_m = s_m2t[m];
      float nA_cl;
      {
        const int r = g * 16 + cl;
        const float4 cv = s_colv[q * R + r];    // nA, thr2, rden2, cos2
        const float al = s_alpha[r], mv = s_m2v[r];
        const float c1B = gs * (1.f / R) * cv.z, kapB = cv.w * cv.z;
        nA_cl = cv.x;
        if (lane < 16) {
          tab[T_NAL * 16 + cl] = cv.x * a.lam2; tab[T_THR2 * 16 + cl] = cv.y;
          tab[T_ALPHA * 16 + cl] = al; tab[T_M2V * 16 + cl] = mv;
          tab[T_LC1B * 16 + cl] = a.lam * c1B; tab[T_NKAPB * 16 + cl] = -kapB; tab[T_NA * 16 + cl] = cv.x;
          tab[T_DDB * 16 + cl] = -0.5f * c1B * kapB;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(VEC_FREE));
      const float nB = rv.x;
      const float c1A = gs * (1.f / kW) * rv.z, kapA = rv.w * rv.z;
      const float lc1A = a.lam * c1A, inv_lam = 1.f / a.lam;

      // -------------- phase 1: recompute u1, u2; U = u and UD = u dD operand tiles --------------
      mbar_wait(bar(S_DONE), ph);
      tc_fence_after();
      TRC(2);
      float2 x[8];                              // pair p = columns 2p, 2p+1 of this warp's chunk
      tmem_ld16(tmem + cS + lane_off + g * 16, reinterpret_cast<uint32_t*>(x));
      tmem_ld_wait();
      uint32_t w1[8], w2[8];                    // u1 (i2t) / u2 (t2i) of the pairs, bf16 as the MMAs see them
      {
        const float2 nBl = f2(nB * a.lam2), mt = f2(m2t_m), th1 = f2(rv.y), bb = f2(beta_m);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 na = tab4[T_NAL * 4 + c4], th = tab4[T_THR2 * 4 + c4], al = tab4[T_ALPHA * 4 + c4],
                       mv = tab4[T_M2V * 4 + c4];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int p = c4 * 2 + e;
            const float2 na2 = e ? make_float2(na.z, na.w) : make_float2(na.x, na.y);
            const float2 th2 = e ? make_float2(th.z, th.w) : make_float2(th.x, th.y);
            const float2 al2 = e ? make_float2(al.z, al.w) : make_float2(al.x, al.y);
            const float2 mv2 = e ? make_float2(mv.z, mv.w) : make_float2(mv.x, mv.y);
            x[p] = fmul2(x[p], fmul2(al2, bb));
            const float2 l = leaky2(x[p]);
            const float2 tA2 = ffma2(l, na2, mv2), tB2 = ffma2(l, nBl, mt);
            const float2 eA = make_float2(ex2(tA2.x), ex2(tA2.y)), eB = make_float2(ex2(tB2.x), ex2(tB2.y));
            // focal select e > thr ? e : 0 as a multiply by the 1.0 / 0.0 compare mask (e is finite)
            const float2 u1 = focal2(eA, th1), u2 = focal2(eB, th2);
            w1[p] = pack_bf16(u1.x, u1.y);
            w2[p] = pack_bf16(u2.x, u2.y);
          }
        }
      }
      if (k > 0) mbar_wait(bar(P_STORED), (k - 1) & 1);     // the TMA store of tile k-1 has read the U1 buffer
      st_shared_v4(sbase + L::U1 + seg0, w1[0], w1[1], w1[2], w1[3]);
      st_shared_v4(sbase + L::U1 + seg1, w1[4], w1[5], w1[6], w1[7]);
      st_shared_v4(sbase + L::U2 + seg0, w2[0], w2[1], w2[2], w2[3]);
      st_shared_v4(sbase + L::U2 + seg1, w2[4], w2[5], w2[6], w2[7]);
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(U_READY));
      if (k > 0) flush_gram();                  // Gram gradients of tile k-1
      // the u dD operand tiles of the Gram-gradient GEMMs are formed while the Y GEMMs run
      {
        const float2 dDA = f2(-0.5f * c1A * kapA);
        uint32_t e1[8], e2[8];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 dd = tab4[T_DDB * 4 + c4];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int p = c4 * 2 + e;
            const float2 v1 = fmul2(bf16_pair(w1[p]), dDA);
            const float2 v2 = fmul2(bf16_pair(w2[p]), e ? make_float2(dd.z, dd.w) : make_float2(dd.x, dd.y));
            e1[p] = pack_bf16(v1.x, v1.y);
            e2[p] = pack_bf16(v2.x, v2.y);
          }
        }
        st_shared_v4(sbase + L::UD1 + seg0, e1[0], e1[1], e1[2], e1[3]);
        st_shared_v4(sbase + L::UD1 + seg1, e1[4], e1[5], e1[6], e1[7]);
        st_shared_v4(sbase + L::UD2 + seg0, e2[0], e2[1], e2[2], e2[3]);
        st_shared_v4(sbase + L::UD2 + seg1, e2[4], e2[5], e2[6], e2[7]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(UD_READY));
      TRC(3);

      // ---------------- phase 2: dA of both directions, l2-norm backward sums ----------------
      // (u is re-read from the operand tiles, Q = dA1 nA + dA2 nB + (c1A u1 + c1B u2) / leaky'(x) is parked in
      //  this thread's own Y1 columns: nothing but x stays in registers across the phases)
      mbar_wait(bar(Y_DONE), ph);
      tc_fence_after();
      TRC(4);
      float tB, tA;
      {
        float2 p4[4], tB2 = f2(0.f);
        const float2 lc1A2 = f2(lc1A), nkapA = f2(-kapA), nB2 = f2(nB);
        // 1 / leaky'(x) / lambda = x > 0 ? 1 / lambda : 10 / lambda, from the 1.0 / 0.0 compare mask
        const float2 il10 = f2(10.f * inv_lam), il1 = f2(inv_lam);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t ua[4], ubb[4];
          const uint32_t so = half ? seg1 : seg0;
          ld_shared_v4(sbase + L::U1 + so, ua[0], ua[1], ua[2], ua[3]);
          ld_shared_v4(sbase + L::U2 + so, ubb[0], ubb[1], ubb[2], ubb[3]);
#pragma unroll
          for (int qt = 0; qt < 2; ++qt) {      // quarters of 4 columns keep the register footprint below the cap
            // (requesting the next quarter's Y values before this one's math - measured: no gain, more spills)
            float2 y1[2], y2[2], Qh[2];
            const int c4 = half * 2 + qt;
            const uint32_t col = g * 16 + c4 * 4;
            tmem_ld4(tmem + cY1 + lane_off + col, reinterpret_cast<uint32_t*>(y1));
            tmem_ld4(tmem + cY2 + lane_off + col, reinterpret_cast<uint32_t*>(y2));
            const float4 t_lc = tab4[T_LC1B * 4 + c4], t_nk = tab4[T_NKAPB * 4 + c4], t_na = tab4[T_NA * 4 + c4];
            tmem_ld_wait();
            float2 pa[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int p = c4 * 2 + e, w = qt * 2 + e;   // pair index; bf16 word holding these two columns
              const float2 lc1B = e ? make_float2(t_lc.z, t_lc.w) : make_float2(t_lc.x, t_lc.y);
              const float2 nkapB = e ? make_float2(t_nk.z, t_nk.w) : make_float2(t_nk.x, t_nk.y);
              const float2 nA2 = e ? make_float2(t_na.z, t_na.w) : make_float2(t_na.x, t_na.y);
              const float2 xv = x[p], l = leaky2(xv);
              // dA = lambda c1 u (S - kap Y); the factors that are constant along a reduction axis (nA along the
              // words, nB along the regions) are applied after the sums
              const float2 gA0 = fmul2(bf16_pair(ua[w]), lc1A2), gB0 = fmul2(bf16_pair(ubb[w]), lc1B);
              const float2 G = fmul2(gA0, ffma2(nkapA, y1[e], xv));
              const float2 F = fmul2(gB0, ffma2(nkapB, y2[e], xv));
              pa[e] = fmul2(l, G);
              tB2 = ffma2(l, F, tB2);
              const float2 ilp = pos_sel2(xv, il1, il10);
              Qh[e] = ffma2(fadd2(gA0, gB0), ilp, ffma2(nA2, G, fmul2(nB2, F)));
            }
            // first butterfly step (lane bit 0 <-> column bit 0) of the lanes->columns reduction of A1 dA1:
            // the lane keeps the column of each pair whose low bit matches its own
            const bool up1 = (lane & 1) != 0;
            const float2 send = make_float2(up1 ? pa[0].x : pa[0].y, up1 ? pa[1].x : pa[1].y);
            const float2 keep = make_float2(up1 ? pa[0].y : pa[0].x, up1 ? pa[1].y : pa[1].x);
            p4[c4] = fadd2(keep, shfl_xor2(send, 1));
            tmem_st4(tmem + cY1 + lane_off + col, reinterpret_cast<const uint32_t*>(Qh));
          }
        }
        tB = tB2.x + tB2.y;
        // remaining steps: p4[c4] = columns 4 c4 + {0, 2} + (lane & 1); lane bit 2h <-> column bit 2h
        {
          const bool up = (lane & 8) != 0;
#pragma unroll
          for (int e = 0; e < 2; ++e) p4[e] = fadd2(sel2(up, p4[e + 2], p4[e]), shfl_xor2(sel2(up, p4[e], p4[e + 2]), 8));
        }
        {
          const bool up = (lane & 4) != 0;
          p4[0] = fadd2(sel2(up, p4[1], p4[0]), shfl_xor2(sel2(up, p4[0], p4[1]), 4));
        }
        const bool up = (lane & 2) != 0;
        const float t = (up ? p4[0].y : p4[0].x) + __shfl_xor_sync(0xffffffffu, up ? p4[0].x : p4[0].y, 2);
        tA = t + __shfl_xor_sync(0xffffffffu, t, 16);   // lane l: column l & 15, summed over 32 words
      }
      tc_fence_before();
      s_tbp[g * kRows + m] = tB;
      if (lane < 16) tab[T_NNATA * 16 + cl] = -(nA_cl * nA_cl * nA_cl * tA);
      TRC(5);
      named_bar_sync(1 + q, 32 * NC);           // the NC warps that share this text slot
      TRC(6);
      float zB = 0.f;
#pragma unroll
      for (int gg = 0; gg < NC; ++gg) zB += s_tbp[gg * kRows + m];
      zB *= nB * nB * nB;

      // ------------------------------ phase 3: dS, P ------------------------------
      uint32_t pk[8];
      tmem_st_wait();
      {
        const float2 nzB = f2(-zB), bb = f2(beta_m), one2 = f2(1.f), c01 = f2(kLeaky);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float2 Qh[4];
          tmem_ld8(tmem + cY1 + lane_off + g * 16 + half * 8, reinterpret_cast<uint32_t*>(Qh));
          tmem_ld_wait();
#pragma unroll
          for (int qt = 0; qt < 2; ++qt) {
            const int c4 = half * 2 + qt;
            const float4 t_nn = tab4[T_NNATA * 4 + c4], t_al = tab4[T_ALPHA * 4 + c4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int p = c4 * 2 + e;
              const float2 nn = e ? make_float2(t_nn.z, t_nn.w) : make_float2(t_nn.x, t_nn.y);
              const float2 al2 = e ? make_float2(t_al.z, t_al.w) : make_float2(t_al.x, t_al.y);
              const float2 xv = x[p];
              // dS = leaky'(x) (Q - leaky(x) (nA^3 tA + zB))
              const float2 lg = pos_sel2(xv, one2, c01);
              const float2 dS = fmul2(lg, ffma2(leaky2(xv), fadd2(nn, nzB), Qh[qt * 2 + e]));
              const float2 pv = fmul2(dS, fmul2(al2, bb));
              pk[p] = pack_bf16(pv.x, pv.y);
            }
          }
        }
      }
      TRC(7);
      mbar_wait(bar(DG_DONE), ph);              // the dG GEMMs have consumed U1 / U2: U1 takes P
      TRC(8);
      st_shared_v4(sbase + L::U1 + seg0, pk[0], pk[1], pk[2], pk[3]);
      st_shared_v4(sbase + L::U1 + seg1, pk[4], pk[5], pk[6], pk[7]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(P_READY));
      TRC(9);

      jt_prev = jt; i_prev = i; last_prev = last_of_video;
      TRC(10);
      if (++jt == tpv) { jt = 0; ++i; }
    }
    if (n_tiles > 0) flush_gram();              // (the last tile's dG GEMMs: DG_DONE was waited for in its phase 3)
  }

  tc_fence_before();
  __syncthreads();
  if (role == 3) tmem_dealloc(tmem, 512);
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
// bf16 matrix (rows x cols, row-major) -> boxes of box_cols x box_rows, 128- or 32-byte swizzle
static int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols,
                    CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return DVLP_ERR_CUDA;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return DVLP_ERR_CUDA; }
  return DVLP_OK;
}

// videos [v0, v0 + Bv) of the batch: every per-video array is offset, the P tiles of the range go to the start of the
// (chunk-sized) P buffer
template <int R>
static int launch(const void* vb_all, const void* tb, int v0, int Bv, int Bt, float lambda, const float* d_scores_all, char* ws,
                  const Workspace& Lw, cudaStream_t st) {
  using L = Lay<R>;
  CUtensorMap mV, mT, mGV, mGBD, mP;
  int rc;
  Args a{};
  a.Bv = Bv; a.Bt = Bt;
  const void* vb = (const char*)vb_all + (size_t)v0 * R * kD * 2;
  const float* d_scores = d_scores_all + (size_t)v0 * Bt;
  a.tiles_per_video = (Bt + kTPT - 1) / kTPT;
  a.total_units = (long long)Bv * a.tiles_per_video;
  if ((rc = make_map(&mV, vb, (uint64_t)Bv * R, kD, R, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&mT, tb, (uint64_t)Bt * kW, kD, kRows, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&mGV, ws + Lw.gvb + (size_t)v0 * R * R * 2, (uint64_t)Bv * R, R, R, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&mGBD, ws + Lw.gbd, (uint64_t)a.tiles_per_video * kRows, kRows, kRows, 64, CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  // P spill: [unit][128 words][R regions] bf16, one 16-column box per chunk
  if ((rc = make_map(&mP, ws + Lw.pa, (uint64_t)a.total_units * kRows, R, kRows, 16, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  a.alpha = (const float*)(ws + Lw.alpha) + (size_t)v0 * R; a.beta = (const float*)(ws + Lw.beta);
  a.m2v = (const float*)(ws + Lw.m2v) + (size_t)v0 * R; a.m2t = (const float*)(ws + Lw.m2t);
  a.d_scores = d_scores;
  a.rowv = (const float4*)(ws + Lw.rowv) + (size_t)v0 * a.tiles_per_video * kRows;
  a.colv = (const float4*)(ws + Lw.colv) + (size_t)v0 * a.tiles_per_video * kTPT * R;
  a.lam = lambda; a.lam2 = lambda * kLog2e;
  a.dgv = (float*)(ws + Lw.dgv) + (size_t)v0 * R * R; a.dgt = (float*)(ws + Lw.dgt);
  a.trace = (long long*)g_trace_buffer;
  int dev = 0, sms = 148;
  DVLP_CUDA_OK(cudaGetDevice(&dev));
  DVLP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = (int)(a.total_units < sms ? a.total_units : sms);
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    DVLP_CUDA_OK(cudaFuncSetAttribute(k_pair_bwd2<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc));
    DVLP_CUDA_OK(cudaFuncSetAttribute(k_pair_bwd2<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc));
    attr_once.done();
  }
  if (a.trace)      // debugging aid (tools/trace_tiles.py): the instrumented instantiation
    k_pair_bwd2<R, true><<<grid, L::kThreads, L::kAlloc, st>>>(mV, mT, mGV, mGBD, mP, a);
  else
    k_pair_bwd2<R, false><<<grid, L::kThreads, L::kAlloc, st>>>(mV, mT, mGV, mGBD, mP, a);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}
}  // namespace tb2

// backward of the videos [v0, v0 + nv) against all Bt texts (vb, d_scores: the whole batch)
int run_pair_bwd2(const void* vb, const void* tb, int v0, int nv, int Bt, int R, int W, int d, float lambda,
                  const float* d_scores, char* ws, const Workspace& L, cudaStream_t st) {
  if (d != tb2::kD || W != tb2::kW) return DVLP_ERR_BAD_SHAPE;
  switch (R) {
    case 80: return tb2::launch<80>(vb, tb, v0, nv, Bt, lambda, d_scores, ws, L, st);
    case 64: return tb2::launch<64>(vb, tb, v0, nv, Bt, lambda, d_scores, ws, L, st);
    case 48: return tb2::launch<48>(vb, tb, v0, nv, Bt, lambda, d_scores, ws, L, st);
    case 32: return tb2::launch<32>(vb, tb, v0, nv, Bt, lambda, d_scores, ws, L, st);
  }
  return DVLP_ERR_BAD_SHAPE;
}

}  // namespace dvlp
