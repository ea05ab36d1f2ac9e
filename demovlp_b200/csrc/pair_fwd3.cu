// Region-word pair scores on tcgen05 / TMEM / TMA (sm_100a): FORWARD kernel for the LARGE / REAL shapes
// (any R <= 288 regions, any W <= 128 words, d = 256): BASELINE config 5 (R = 288, W = 64), the reference's shipped
// shapes R = 30 / 240 with W = 99, the MC 5-option case.  Reference arithmetic: model/loss.py:209-330.
// pair_fwd2.cu keeps the whole video and every S column of a tile on chip; here neither fits, so the kernel STREAMS
// the video in chunks of NB regions and walks every (video, text tile) unit in TWO SWEEPS over the chunks:
//
//   tile      128 TMEM lanes = TPT texts x WP padded words (WP = 32 / 64 / 128, TPT = 128 / WP; words >= W and texts
//             >= Bt are zero rows that TMA fills in), TMEM columns = the NB regions of the current chunk
//   sweep 1   S_c = T V_c^T (tcgen05, K = 256)  ->  x = raw alpha beta, l = leaky(x): column norms nA (complete inside a
//             chunk: the sum runs over the words of a text), row sums of l^2 (-> nB) and of e1 (-> focal threshold thr1)
//   sweep 2   S_c again (the tensor pipe has slack; S is never kept: 288 columns of it would not leave room for the
//             rest)  ->  u1 (i2t direction, focal-selected exponentials), e2 with nB, its column sums (-> thr2), u2,
//             both numerators;  then on the tensor cores
//                 C   += (u1 alpha)_c V_c           the weighted context of every word, DIRECT form (N = 256, K = NB):
//                                                    |c_w|^2 is a row sum of squares at the end of the unit -- no
//                                                    R x R video Gram matrix exists at these sizes (166 KB at R = 288)
//                 Y2_c  = Gbd U2_c                  Gram form of the t2i direction (text Gram blocks are <= 128 x 128)
//   end       D1 = |C_w|^2, cos1 = num1 / sqrt(D1);  cos2 = num2 / sqrt(sum_w u2 Y2) per chunk;  score = mean + mean
//
// Units run TEXT-major (unit = text tile * Bv + video): the text tile T and its block-diagonal Gram tile stay resident,
// the video streams through a 3-stage ring (each stage serves the S product K-major and the C product MN-major).
// Real counts R, W enter only the focal thresholds and the means; padded rows / columns are exact zeros with a mask
// term of -1e30 (tests/emul_chunked.py pins these padding rules against the oracle).
//
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warp 3 TMEM owner, warps 4.. compute: warpgroup g owns the 16
// columns 16g..16g+15 of every chunk (NB / 16 warpgroups), its warp q the TMEM lane quadrant q = 32 word rows.  Sums over
// the words of a text = 32-lane shuffle transpose-reductions, combined across the WP / 32 warps of a text through shared
// memory; sums over regions = per-thread accumulators over the chunks, combined across warpgroups once per sweep.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dvlp {
using namespace ptx;

namespace tf3 {
constexpr int kD = 256, kRows = 128;
constexpr int kTileBytes = 16384, kCh32 = 4096;
constexpr int kNS = 3;            // video-chunk ring stages
constexpr int kNVec = 3;          // ring of per-chunk (alpha, m2v) vectors
constexpr int kMaxTab = 576;      // TPT * Rp floats of nA lam2 kept from sweep 1 to sweep 2

template <int NB>
struct Lay {
  static constexpr int NG = NB / 16, NW = 4 * NG, kThreads = 128 * (1 + NG);
  static constexpr int VBOX = NB * 128;                 // one 64-column box of a video chunk
  static constexpr int VSTAGE = 4 * VBOX;               // NB rows x 512 B
  static constexpr int UBYTES = NG * kCh32;             // one [128 rows][NB columns] bf16 operand tile
  static constexpr int T = 0;                           // 4 x 16 KB
  static constexpr int GBD = T + 4 * kTileBytes;        // 2 x 16 KB
  static constexpr int VR = GBD + 2 * kTileBytes;       // kNS stages
  static constexpr int U1 = VR + kNS * VSTAGE;          // 2 slots
  static constexpr int U2 = U1 + 2 * UBYTES;            // 2 slots
  static constexpr int VEC = U2 + 2 * UBYTES;           // kNVec x (alpha[NB], m2v[NB])
  static constexpr int BETA = VEC + kNVec * NB * 8;     // [128]
  static constexpr int M2T = BETA + 512;                // [128]
  static constexpr int TABNA = M2T + 512;               // [TPT][Rp]
  static constexpr int COLX = TABNA + kMaxTab * 4;      // [2][NW][16] column partial sums (text spanning several warps)
  static constexpr int ROWX = COLX + 2 * NW * 64;       // [NG][128] float2 row partial sums
  static constexpr int FIN = ROWX + NG * 1024;          // [NW] t2i sums, [4] i2t sums
  static constexpr int BARS = FIN + NW * 4 + 16;
  static constexpr int TMEMPTR = BARS + 32 * 8;
  static constexpr int kAlloc = TMEMPTR + 16;
  static_assert(VSTAGE % 1024 == 0 && UBYTES % 1024 == 0, "swizzled tiles need 1 KB alignment");
  static_assert(kAlloc <= 232448, "shared memory budget");
};

enum Bar { TJ_FULL = 0, UNIT_DONE, C_FULL, C_FREE, V_FULL /*3*/, V_EMPTY = V_FULL + kNS, VEC_FULL = V_EMPTY + kNS,
           VEC_EMPTY = VEC_FULL + kNVec, S_FULL = VEC_EMPTY + kNVec /*2*/, S_FREE = S_FULL + 2, U_READY = S_FREE + 2,
           Y2_FULL = U_READY + 2, kNumBars = Y2_FULL + 2 };
static_assert(kNumBars <= 32, "barrier area");

// TMEM columns: two S slots, two Y2 slots, C
constexpr uint32_t cS = 0, cY2 = 128, cC = 256;

struct Args {
  int Bv, Bt, R, W, Rp, nch, tiles;
  int grouped, n_t, tpg;           // grouped mode (multiple choice): video g against ITS n_t texts only; tpg tiles per group
  long long total_units;
  const float *alpha3, *m2v3;      // [Bv][Rp]   (0 / -1e30 beyond R)
  const float *beta3, *m2t3;       // [tiles][128] in tile-row order (0 / -1e30 for padded words and absent texts)
  float lam2, inv_R, inv_W;
  int focal;
  float* scores;
  float4 *rowv, *colv;             // training: per-unit query scalars for the backward (NULL: inference)
};

__device__ __forceinline__ uint32_t desc_hi32(uint32_t sbo_bytes) {      // 32-byte swizzle
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (6u << 29);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// Sum over the 32 lanes of v[c], c = 0..15: lane l ends with the total of column l & 15 (see pair_fwd2.cu).
__device__ __forceinline__ float lanes_to_columns16(float (&v)[16], int lane) {
#pragma unroll
  for (int h = 8; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h];
      const float keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

template <int NB, int NQ>
__global__ void __launch_bounds__(Lay<NB>::kThreads, 1)
k_pair_fwd3(const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapT,
            const __grid_constant__ CUtensorMap mapGBD, Args a) {
  using L = Lay<NB>;
  constexpr int NG = L::NG, NW = L::NW, TPT = 4 / NQ;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if (sbase & 1023u) __trap();
  uint8_t* sgen = smem_raw;
  volatile uint32_t* s_tmem = (volatile uint32_t*)(sgen + L::TMEMPTR);
  auto bar = [&](int i) { return sbase + L::BARS + 8 * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long u0 = a.total_units * blockIdx.x / gridDim.x;
  const long long u1 = a.total_units * (blockIdx.x + 1) / gridDim.x;
  const int n_units = (int)(u1 - u0);
  const int nch = a.nch;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapV); prefetch_tmap(&mapT); prefetch_tmap(&mapGBD);
    for (int b = 0; b < kNumBars; ++b) {
      uint32_t cnt = 1;
      if (b == UNIT_DONE || b == C_FREE || (b >= VEC_EMPTY && b < VEC_EMPTY + kNVec) || (b >= S_FREE && b < S_FREE + 2) ||
          (b >= U_READY && b < U_READY + 2))
        cnt = NW;
      mbar_init(bar(b), cnt);
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

  if (warp < 4) {
    if (warp == 0) {
      // =============================== TMA producer ===============================
      if (elect_one()) {
        uint32_t nV = 0;
        int jt_prev = -1;
        for (int k = 0; k < n_units; ++k) {
          const long long u = u0 + k;
          // dense: unit = text tile * Bv + video (text-major);  grouped: unit = tile, its video = tile / tpg
          const int jt = a.grouped ? (int)u : (int)(u / a.Bv);
          const int i = a.grouped ? jt / a.tpg : (int)(u - (long long)jt * a.Bv);
          if (jt != jt_prev) {
            // new text tile: every consumer of the old one (S / Y2 products, beta / m2t readers) finished unit k-1
            if (k > 0) mbar_wait(bar(UNIT_DONE), (k - 1) & 1);
            // first text of the tile (grouped: tile tt of group g starts at text g * n_t + tt * TPT; the rows past the
            // group's last text belong to the next group: computed and never stored)
            const int j0 = a.grouped ? (jt / a.tpg) * a.n_t + (jt % a.tpg) * TPT : jt * TPT;
            mbar_arrive_expect_tx(bar(TJ_FULL), 6 * kTileBytes + 1024);
            for (int c = 0; c < 4; ++c) tma_load_3d(sbase + L::T + c * kTileBytes, &mapT, c * 64, 0, j0, bar(TJ_FULL));
            for (int c = 0; c < 2; ++c) tma_load_2d(sbase + L::GBD + c * kTileBytes, &mapGBD, c * 64, jt * kRows, bar(TJ_FULL));
            bulk_load(sbase + L::BETA, a.beta3 + (size_t)jt * kRows, 512, bar(TJ_FULL));
            bulk_load(sbase + L::M2T, a.m2t3 + (size_t)jt * kRows, 512, bar(TJ_FULL));
            jt_prev = jt;
          }
          for (int sweep = 0; sweep < 2; ++sweep)
            for (int c = 0; c < nch; ++c, ++nV) {
              const uint32_t st = nV % kNS, vs = nV % kNVec;
              if (nV >= kNS) mbar_wait(bar(V_EMPTY + st), (nV / kNS - 1) & 1);
              mbar_arrive_expect_tx(bar(V_FULL + st), L::VSTAGE);
              for (int kc = 0; kc < 4; ++kc)
                tma_load_3d(sbase + L::VR + st * L::VSTAGE + kc * L::VBOX, &mapV, kc * 64, c * NB, i, bar(V_FULL + st));
              if (nV >= kNVec) mbar_wait(bar(VEC_EMPTY + vs), (nV / kNVec - 1) & 1);
              mbar_arrive_expect_tx(bar(VEC_FULL + vs), NB * 8);
              bulk_load(sbase + L::VEC + vs * NB * 8, a.alpha3 + (size_t)i * a.Rp + c * NB, NB * 4, bar(VEC_FULL + vs));
              bulk_load(sbase + L::VEC + vs * NB * 8 + NB * 4, a.m2v3 + (size_t)i * a.Rp + c * NB, NB * 4, bar(VEC_FULL + vs));
            }
        }
      }
    } else if (warp == 1) {
      // ================================ MMA issuer ================================
      constexpr uint32_t idS = idesc_bf16(128, NB), idY2 = idesc_bf16(128, NB, 0, 1), idC = idesc_bf16(128, 256, 0, 1);
      const uint32_t hi128 = desc_hi(1024), hi32 = desc_hi32(256);
      const uint32_t loT = desc_lo(sbase + L::T), loGBD = desc_lo(sbase + L::GBD);
      uint32_t nS = 0, nU = 0;
      int jt_prev = -1, tj_loads = 0;
      // S_c = T V_c^T into S slot n & 1; `release`: the stage is not needed by a C product (sweep 1)
      auto issue_S = [&](uint32_t n, bool release) {
        const uint32_t st = n % kNS, slot = n & 1;
        mbar_wait(bar(V_FULL + st), (n / kNS) & 1);
        if (n >= 2) mbar_wait(bar(S_FREE + slot), ((n >> 1) - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t loV = desc_lo(sbase + L::VR + st * L::VSTAGE);
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {
            const uint32_t oT = ((ks >> 2) * kTileBytes + (ks & 3) * 32) >> 4, oV = ((ks >> 2) * L::VBOX + (ks & 3) * 32) >> 4;
            mma_bf16_w(tmem + cS + slot * 64, loT + oT, hi128, loV + oV, hi128, idS, ks != 0);
          }
          tc_commit(bar(S_FULL + slot));
          if (release) tc_commit(bar(V_EMPTY + st));
        }
        __syncwarp();
      };
      // C += (u1 alpha)_c V_c  and  Y2_c = Gbd U2_c for sweep-2 chunk m (operand slot m & 1), video stage of S product n
      auto issue_CY = [&](uint32_t m, uint32_t n, bool first, bool last, int k) {
        const uint32_t us = m & 1, st = n % kNS;
        mbar_wait(bar(U_READY + us), (m >> 1) & 1);
        if (first && k > 0) mbar_wait(bar(C_FREE), (k - 1) & 1);     // the previous unit's C has been read
        tc_fence_after();
        if (elect_one()) {
          const uint32_t loVm = desc_lo(sbase + L::VR + st * L::VSTAGE, L::VBOX);       // MN-major: 64-column groups VBOX apart
          const uint32_t loU1 = desc_lo(sbase + L::U1 + us * L::UBYTES);
          const uint32_t loU2m = desc_lo(sbase + L::U2 + us * L::UBYTES, kCh32);
#pragma unroll
          for (int ks = 0; ks < NG; ++ks)
            mma_bf16_w(tmem + cC, loU1 + ((ks * kCh32) >> 4), hi32, loVm + ((ks * 2048) >> 4), hi128, idC, !(first && ks == 0));
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            mma_bf16_w(tmem + cY2 + us * 64, loGBD + (((ks >> 2) * kTileBytes + (ks & 3) * 32) >> 4), hi128,
                       loU2m + ((ks * 512) >> 4), hi32, idY2, ks != 0);
          tc_commit(bar(Y2_FULL + us));
          tc_commit(bar(V_EMPTY + st));
          if (last) tc_commit(bar(C_FULL));
        }
        __syncwarp();
      };
      for (int k = 0; k < n_units; ++k) {
        const int jt = a.grouped ? (int)(u0 + k) : (int)((u0 + k) / a.Bv);
        if (jt != jt_prev) { mbar_wait(bar(TJ_FULL), tj_loads & 1); ++tj_loads; jt_prev = jt; }
        for (int c = 0; c < nch; ++c) issue_S(nS++, true);
        for (int c = 0; c < nch; ++c) {
          issue_S(nS++, false);
          if (c > 0) { issue_CY(nU, nS - 2, c == 1, false, k); ++nU; }
        }
        issue_CY(nU, nS - 1, nch == 1, true, k); ++nU;
      }
    }
  } else {
    // ============================ compute warps ============================
    const int cw = warp - 4;
    const int g = cw >> 2, q = cw & 3;          // column sub-chunk of every chunk, TMEM lane quadrant
    const int m = q * 32 + lane;                // word row of the tile
    const int cl = lane & 15;
    const int t = q / NQ, q0 = t * NQ;          // text slot of this warp, first warp of that text
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t sw = (uint32_t)((m >> 2) & 1) << 4;
    const uint32_t row_off = (uint32_t)g * kCh32 + (uint32_t)m * 32;
    const uint32_t seg0 = row_off + sw, seg1 = row_off + (16u ^ sw);
    const float* s_beta = reinterpret_cast<const float*>(sgen + L::BETA);
    const float* s_m2t = reinterpret_cast<const float*>(sgen + L::M2T);
    float* s_tab = reinterpret_cast<float*>(sgen + L::TABNA) + t * a.Rp + g * 16;     // + c * NB: this warp's 16 columns
    float* s_colx = reinterpret_cast<float*>(sgen + L::COLX);
    float2* s_rowx = reinterpret_cast<float2*>(sgen + L::ROWX);
    float* s_fin = reinterpret_cast<float*>(sgen + L::FIN);
    uint32_t xcnt = 0;
    // total over the words of this warp's text of v[c] (column c of this warp's 16), for every lane's column lane & 15
    auto colsum = [&](float (&v)[16]) -> float {
      const float r = lanes_to_columns16(v, lane);
      if (NQ == 1) return r;
      float* base = s_colx + (xcnt & 1) * NW * 16;
      ++xcnt;
      if (lane < 16) base[cw * 16 + lane] = r;
      named_bar_sync(2 + g * TPT + t, 32 * NQ);           // the NQ warps that hold this text's rows of this sub-chunk
      float tot = 0.f;
#pragma unroll
      for (int qq = 0; qq < NQ; ++qq) tot += base[(g * 4 + q0 + qq) * 16 + cl];
      return tot;
    };
    uint32_t nS = 0, nU = 0;
    int jt_prev = -1, tj_loads = 0;
    float beta_m = 0.f, m2t_m = 0.f;

    for (int k = 0; k < n_units; ++k) {
      const long long u = u0 + k;
      const int jt = a.grouped ? (int)u : (int)(u / a.Bv);
      const int i = a.grouped ? jt / a.tpg : (int)(u - (long long)jt * a.Bv);
      if (jt != jt_prev) {
        mbar_wait(bar(TJ_FULL), tj_loads & 1); ++tj_loads; jt_prev = jt;
        beta_m = s_beta[m]; m2t_m = s_m2t[m];
      }
      // ============ sweep 1: column norms (per chunk), row sums of l^2 and e1 ============
      float ssB = 0.f, sumA = 0.f;
      for (int c = 0; c < nch; ++c, ++nS) {
        const uint32_t slot = nS & 1, vs = nS % kNVec;
        float x[16], sq[16];
        mbar_wait(bar(S_FULL + slot), (nS >> 1) & 1);
        tc_fence_after();
        tmem_ld16(tmem + cS + slot * 64 + lane_off + g * 16, reinterpret_cast<uint32_t*>(x));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(S_FREE + slot));
        mbar_wait(bar(VEC_FULL + vs), (nS / kNVec) & 1);
        const float4* s_al = reinterpret_cast<const float4*>(sgen + L::VEC + vs * NB * 8) + g * 4;
        const float4* s_mv = reinterpret_cast<const float4*>(sgen + L::VEC + vs * NB * 8 + NB * 4) + g * 4;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 al = s_al[c4];
          const float av[4] = {al.x, al.y, al.z, al.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int cc = c4 * 4 + e;
            x[cc] = leaky(x[cc] * (av[e] * beta_m));         // x now holds l
            sq[cc] = x[cc] * x[cc];
            ssB += sq[cc];
          }
        }
        const float ssA = colsum(sq);                        // sum over the words of this text, column cl
        const float nAl = a.lam2 * inv_norm(ssA);            // nA * lambda * log2(e)   (model/loss.py:238,244)
        if (lane < 16 && q == q0) s_tab[c * NB + cl] = nAl;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 mv = s_mv[c4];
          const float mvv[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int cc = c4 * 4 + e;
            sumA += ex2(fmaf(x[cc], __shfl_sync(0xffffffffu, nAl, cc), mvv[e]));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(VEC_EMPTY + vs));
      }
      // row totals over all chunks: exchange between the warpgroups
      s_rowx[g * kRows + m] = make_float2(ssB, sumA);
      named_bar_sync(1, 128 * NG);
      float totB = 0.f, totA = 0.f;
#pragma unroll
      for (int gg = 0; gg < NG; ++gg) {
        const float2 p = s_rowx[gg * kRows + m];
        totB += p.x; totA += p.y;
      }
      const float nB = inv_norm(totB);
      const float nBl = nB * a.lam2;
      const float thr1 = a.focal ? totA * a.inv_R : -1.f;     // keep e iff e * R > sum_r e  (model/loss.py:280-282), REAL R

      // ============ sweep 2: u1, u2, numerators; deferred by one chunk: the Gram-form t2i denominators ============
      float num1 = 0.f, acc_t2i = 0.f;
      float num2_p = 0.f, nA_p = 0.f, thr2_p = 0.f;     // carried to the deferred pass of the chunk
      uint32_t us_p = 0, m_p = 0;
      int c_p = 0;
      // D2 of chunk c_p = sum over the words of u2 Y2 (Y2_c = Gbd U2_c ran under the next chunk's main pass);
      // cos2 = num2 / sqrt(D2), summed over this warp's 16 regions (model/loss.py:326-327)
      auto deferred_pass = [&]() {
        uint32_t uw[8];
        float y[16];
        mbar_wait(bar(Y2_FULL + us_p), (m_p >> 1) & 1);
        tc_fence_after();
        ld_shared_v4(sbase + L::U2 + us_p * L::UBYTES + seg0, uw[0], uw[1], uw[2], uw[3]);
        ld_shared_v4(sbase + L::U2 + us_p * L::UBYTES + seg1, uw[4], uw[5], uw[6], uw[7]);
        tmem_ld16(tmem + cY2 + us_p * 64 + lane_off + g * 16, reinterpret_cast<uint32_t*>(y));
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int p2 = 0; p2 < 8; ++p2) {
          y[2 * p2] *= bf16_lo(uw[p2]);
          y[2 * p2 + 1] *= bf16_hi(uw[p2]);
        }
        const float D2 = colsum(y);
        const float rden2 = D2 > 1e-30f ? rsqrtf(D2) : 0.f;
        float cos2 = num2_p * rden2;                           // padded regions: num2 == 0 exactly
        if (a.colv && lane < 16 && q == q0)
          a.colv[((size_t)u * TPT + t) * a.Rp + c_p * NB + g * 16 + cl] = make_float4(nA_p, thr2_p, rden2, cos2);
        acc_t2i += cos2;                                       // per lane (= per column); lanes meet once per unit
      };
      for (int c = 0; c < nch; ++c, ++nS, ++nU) {
        const uint32_t slot = nS & 1, vs = nS % kNVec, us = nU & 1;
        float x[16], e2[16];
        uint32_t w1[8], w2[8];
        mbar_wait(bar(S_FULL + slot), (nS >> 1) & 1);
        tc_fence_after();
        tmem_ld16(tmem + cS + slot * 64 + lane_off + g * 16, reinterpret_cast<uint32_t*>(x));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(S_FREE + slot));
        mbar_wait(bar(VEC_FULL + vs), (nS / kNVec) & 1);
        const float4* s_al = reinterpret_cast<const float4*>(sgen + L::VEC + vs * NB * 8) + g * 4;
        const float4* s_mv = reinterpret_cast<const float4*>(sgen + L::VEC + vs * NB * 8 + NB * 4) + g * 4;
        const float4* s_nl = reinterpret_cast<const float4*>(s_tab + c * NB);
        // i2t direction: u1 = focal-selected exponential; the MMA operand is bf16(u1 alpha) (C = sum_r (u1 alpha) V_r is
        // the weighted context on the RAW rows) and the numerator uses the SAME rounded value: sum_r bf16(u1 alpha) raw beta
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 al = s_al[c4], mv = s_mv[c4], nl = s_nl[c4];
          const float av[4] = {al.x, al.y, al.z, al.w}, mvv[4] = {mv.x, mv.y, mv.z, mv.w}, nv[4] = {nl.x, nl.y, nl.z, nl.w};
          float ua[4], xb[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int cc = c4 * 4 + e;
            xb[e] = x[cc] * beta_m;                            // raw beta
            x[cc] = xb[e] * av[e];                             // S
            const float l = leaky(x[cc]);
            const float e1 = ex2(fmaf(l, nv[e], mvv[e]));
            ua[e] = e1 > thr1 ? e1 * av[e] : 0.f;
            e2[cc] = ex2(fmaf(l, nBl, m2t_m));
          }
          const uint32_t p0 = pack_bf16(ua[0], ua[1]), p1 = pack_bf16(ua[2], ua[3]);
          w1[c4 * 2] = p0; w1[c4 * 2 + 1] = p1;
          num1 = fmaf(bf16_lo(p0), xb[0], num1); num1 = fmaf(bf16_hi(p0), xb[1], num1);
          num1 = fmaf(bf16_lo(p1), xb[2], num1); num1 = fmaf(bf16_hi(p1), xb[3], num1);
        }
        const float nA_cl = reinterpret_cast<const float*>(s_nl)[cl];
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(VEC_EMPTY + vs));
        // t2i direction: focal threshold = column mean of e2 over the REAL words (model/loss.py:280-282)
        float thr2_cl;
        {
          float eb[16];
#pragma unroll
          for (int cc = 0; cc < 16; ++cc) eb[cc] = e2[cc];
          const float sumB = colsum(eb);
          thr2_cl = a.focal ? sumB * a.inv_W : -1.f;
        }
        float pn[16];
#pragma unroll
        for (int cc = 0; cc < 16; cc += 2) {
          const float t0 = __shfl_sync(0xffffffffu, thr2_cl, cc), t1 = __shfl_sync(0xffffffffu, thr2_cl, cc + 1);
          const uint32_t p = pack_bf16(e2[cc] > t0 ? e2[cc] : 0.f, e2[cc + 1] > t1 ? e2[cc + 1] : 0.f);
          w2[cc >> 1] = p;
          pn[cc] = bf16_lo(p) * x[cc];
          pn[cc + 1] = bf16_hi(p) * x[cc + 1];
        }
        // operand tiles of this chunk: the products that read slot `us` two chunks ago completed before the deferred
        // pass of that chunk (Y2_FULL), which this thread has already run
        st_shared_v4(sbase + L::U1 + us * L::UBYTES + seg0, w1[0], w1[1], w1[2], w1[3]);
        st_shared_v4(sbase + L::U1 + us * L::UBYTES + seg1, w1[4], w1[5], w1[6], w1[7]);
        st_shared_v4(sbase + L::U2 + us * L::UBYTES + seg0, w2[0], w2[1], w2[2], w2[3]);
        st_shared_v4(sbase + L::U2 + us * L::UBYTES + seg1, w2[4], w2[5], w2[6], w2[7]);
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(U_READY + us));
        const float num2_cl = colsum(pn);                      // under the products
        if (c > 0) deferred_pass();
        num2_p = num2_cl; us_p = us; m_p = nU; c_p = c; thr2_p = thr2_cl; nA_p = nA_cl;
      }
      deferred_pass();
      // ============ end of the unit: |c_w|^2 from C (this warp's share of the 256 columns), cosines, score ============
      float D1p = 0.f;
      {
        constexpr int kGroups = 16;                           // 16-column groups of C
        const int g_lo = kGroups * g / NG, g_hi = kGroups * (g + 1) / NG;
        mbar_wait(bar(C_FULL), k & 1);
        tc_fence_after();
        for (int cg = g_lo; cg < g_hi; ++cg) {
          float v[16];
          tmem_ld16(tmem + cC + lane_off + cg * 16, reinterpret_cast<uint32_t*>(v));
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) D1p = fmaf(v[e], v[e], D1p);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(C_FREE));
      }
      s_rowx[g * kRows + m] = make_float2(D1p, num1);
      named_bar_sync(1, 128 * NG);
      if (g == 0) {
        float d1 = 0.f, n1 = 0.f;
#pragma unroll
        for (int gg = 0; gg < NG; ++gg) {
          const float2 p = s_rowx[gg * kRows + m];
          d1 += p.x; n1 += p.y;
        }
        const float rden1 = d1 > 1e-30f ? rsqrtf(d1) : 0.f;
        float cs = n1 * rden1;                                // padded word rows: n1 == 0 exactly
        if (a.rowv) a.rowv[(size_t)u * kRows + m] = make_float4(nB, thr1, rden1, cs);
        cs = warp_sum(cs);
        if (lane == 0) s_fin[NW + q] = cs;
      }
      acc_t2i += __shfl_xor_sync(0xffffffffu, acc_t2i, 8);    // the 16 columns of this warp (lanes l, l ^ 16 hold the same)
      acc_t2i += __shfl_xor_sync(0xffffffffu, acc_t2i, 4);
      acc_t2i += __shfl_xor_sync(0xffffffffu, acc_t2i, 2);
      acc_t2i += __shfl_xor_sync(0xffffffffu, acc_t2i, 1);
      if (lane == 0) s_fin[cw] = acc_t2i;
      named_bar_sync(1, 128 * NG);
      if (cw == 0 && lane < TPT) {
        // dense: text j of the batch, scores (Bv, Bt);  grouped: option jj of group i, scores (groups, n_t)
        const int jj = a.grouped ? (jt % a.tpg) * TPT + lane : jt * TPT + lane;
        const bool ok = a.grouped ? jj < a.n_t : jj < a.Bt;
        const int j = jj;
        if (ok) {
          float i2t = 0.f, t2i = 0.f;
#pragma unroll
          for (int qq = 0; qq < NQ; ++qq) i2t += s_fin[NW + lane * NQ + qq];
#pragma unroll
          for (int gg = 0; gg < NG; ++gg) t2i += s_fin[gg * 4 + lane * NQ];
          a.scores[(size_t)i * (a.grouped ? a.n_t : a.Bt) + j] = i2t * a.inv_W + t2i * a.inv_R;   // means over the REAL counts (loss.py:318,327)
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(UNIT_DONE));
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem, 512);
}
}  // namespace tf3
}  // namespace dvlp

namespace dvlp {
// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn3 encode_fn3() {
  static EncodeTiledFn3 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn3)p;
  }
  return fn;
}
// bf16 tensor (n2, n1, n0) row-major, boxes of (b2, b1, b0 = 64 columns), 128-byte swizzle; elements outside the
// tensor are ZERO-filled by TMA (this is what pads regions / words / absent texts)
int make_map_items(CUtensorMap* m, const void* base, uint64_t n0, uint64_t n1, uint64_t n2, uint32_t b1, uint32_t b2) {
  EncodeTiledFn3 fn = encode_fn3();
  if (!fn) return DVLP_ERR_CUDA;
  cuuint64_t gdim[3] = {n0, n1, n2};
  cuuint64_t gstride[2] = {n0 * 2, n0 * n1 * 2};
  cuuint32_t box[3] = {64, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return DVLP_ERR_CUDA; }
  return DVLP_OK;
}
int make_map_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols,
                CUtensorMapSwizzle swz) {
  EncodeTiledFn3 fn = encode_fn3();
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

// Tiling of the large-shape kernels for (R, W): chunk width NB (48, or 32 when that pads fewer regions), padded region
// count Rp = nch * NB, padded words WP = 32 / 64 / 128 (NQ = WP / 32 warps per text, TPT = 128 / WP texts per tile).
bool shape3(int R, int W, int d, Shape3* s) {
  if (d != 256 || R < 1 || W < 1 || W > 128) return false;
  Shape3 t{};
  t.WP = W <= 32 ? 32 : (W <= 64 ? 64 : 128);
  t.NQ = t.WP / 32;
  t.TPT = 128 / t.WP;
  const int rp48 = (R + 47) / 48 * 48, rp32 = (R + 31) / 32 * 32;
  t.NB = rp32 < rp48 ? 32 : 48;
  t.Rp = t.NB == 32 ? rp32 : rp48;
  t.nch = t.Rp / t.NB;
  // table of column norms kept between the sweeps; the dV' GEMM keeps one TMEM column per region (grad_gemm3.cu)
  if (t.TPT * t.Rp > tf3::kMaxTab || t.Rp > 384) return false;
  if (s) *s = t;
  return true;
}

namespace tf3 {
template <int NB, int NQ>
static int launch(const CUtensorMap& mV, const CUtensorMap& mT, const CUtensorMap& mG, const Args& a, cudaStream_t st) {
  using L = Lay<NB>;
  int dev = 0, sms = 148;
  DVLP_CUDA_OK(cudaGetDevice(&dev));
  DVLP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = (int)(a.total_units < sms ? a.total_units : sms);
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    DVLP_CUDA_OK(cudaFuncSetAttribute(k_pair_fwd3<NB, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc));
    attr_once.done();
  }
  k_pair_fwd3<NB, NQ><<<grid, L::kThreads, L::kAlloc, st>>>(mV, mT, mG, a);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}
}  // namespace tf3

// n_t == 0: dense (every video against every text).  n_t > 0: GROUPED -- video g against texts [g n_t, (g+1) n_t) only
// (multiple-choice evaluation, trainer/trainer_dist.py:792-829); Bt = Bv * n_t, scores (Bv, n_t).
int run_pair_fwd3(const void* vb, const void* tb, int Bv, int Bt, int R, int W, int d, float lambda, int focal,
                  float* scores, char* ws, const Workspace& Lw, bool save, cudaStream_t st, int n_t) {
  Shape3 s3;
  if (!shape3(R, W, d, &s3)) return DVLP_ERR_BAD_SHAPE;
  const int tpg = n_t > 0 ? (n_t + s3.TPT - 1) / s3.TPT : 0;
  const int tiles = n_t > 0 ? Bv * tpg : (Bt + s3.TPT - 1) / s3.TPT;
  CUtensorMap mV, mT, mG;
  int rc;
  if ((rc = make_map_items(&mV, vb, 256, (uint64_t)R, (uint64_t)Bv, (uint32_t)s3.NB, 1))) return rc;
  if ((rc = make_map_items(&mT, tb, 256, (uint64_t)W, (uint64_t)Bt, (uint32_t)s3.WP, (uint32_t)s3.TPT))) return rc;
  if ((rc = make_map_2d(&mG, ws + Lw.gbd3, (uint64_t)tiles * 128, 128, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  tf3::Args a{};
  a.Bv = Bv; a.Bt = Bt; a.R = R; a.W = W; a.Rp = s3.Rp; a.nch = s3.nch; a.tiles = tiles;
  a.grouped = n_t > 0 ? 1 : 0; a.n_t = n_t; a.tpg = tpg;
  a.total_units = n_t > 0 ? (long long)tiles : (long long)tiles * Bv;
  a.alpha3 = (const float*)(ws + Lw.alpha3); a.m2v3 = (const float*)(ws + Lw.m2v3);
  a.beta3 = (const float*)(ws + Lw.beta3); a.m2t3 = (const float*)(ws + Lw.m2t3);
  a.lam2 = lambda * kLog2e; a.inv_R = 1.f / (float)R; a.inv_W = 1.f / (float)W; a.focal = focal;
  a.scores = scores;
  a.rowv = save ? (float4*)(ws + Lw.rowv) : nullptr;
  a.colv = save ? (float4*)(ws + Lw.colv) : nullptr;
  if (s3.NB == 48) {
    if (s3.NQ == 1) return tf3::launch<48, 1>(mV, mT, mG, a, st);
    if (s3.NQ == 2) return tf3::launch<48, 2>(mV, mT, mG, a, st);
    return tf3::launch<48, 4>(mV, mT, mG, a, st);
  }
  if (s3.NQ == 1) return tf3::launch<32, 1>(mV, mT, mG, a, st);
  if (s3.NQ == 2) return tf3::launch<32, 2>(mV, mT, mG, a, st);
  return tf3::launch<32, 4>(mV, mT, mG, a, st);
}

}  // namespace dvlp
