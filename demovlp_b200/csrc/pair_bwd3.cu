// Region-word pair scores on tcgen05 / TMEM / TMA (sm_100a): BACKWARD kernel for the LARGE / REAL shapes
// (partner of pair_fwd3.cu; any R, W <= 128 that shape3() tiles, d = 256).  Back-propagates d_scores through both
// attention directions of model/loss.py:209-330 for one (video, text tile) unit at a time, recomputing the attention
// from the raw similarities in chunks of NB regions (nothing of size Bv*Bt*R*W is ever stored); the per-query scalars
// of the forward (attention norms, focal thresholds, 1/|c|, cosines) come from pair_fwd3.cu (rowv / colv).
//
// Two sweeps over the region chunks of a unit (tile = 128 lanes = TPT texts x WP padded words, columns = regions):
//   sweep A   S_c = T V_c^T  ->  u1, u2 (as the forward selected them)  ->  tensor cores:  C += (u1 alpha)_c V_c  (the
//             weighted context of every word, direct form) and Y2_c = Gbd U2_c;  one chunk later: the row sums
//             sB = sum_r l F of the t2i direction's l2-norm backward (F = c1B u2 (S - kapB Y2))
//   between   C (fp32, 256 TMEM columns) -> bf16: an MMA A-operand IN TMEM (128 columns) and a spill for grad_gemm3.cu
//   sweep B   S_c again, Y1_c = C_bf16 V_c^T (the i2t direction's "Gram" term alpha_r <c_w, v_r>, A operand from TMEM),
//             Y2_c again, dG^T += (u2 dD) U2^T (text Gram gradient, flushed per unit);  one chunk later, with everything
//             of the chunk at hand:  dA of both directions, the column sums sA = sum_w l G, and
//                 P   = dS alpha_r beta_w,   dS = leaky'(S) lambda (G nA + F nB - l (nA^3 sA + nB^3 sB)) + c1A u1 + c1B u2
//                 UD1 = (u1 alpha) 2 dD1     (weights of d|c_w|^2 / dV:  dV' += UD1^T C_bf16)
//             written straight from registers to global memory (bf16, [unit][128 words][Rp regions]).
// grad_gemm3.cu contracts them:  dV'[i] = sum_jt P^T T_jt + UD1^T C_jt,   dT'[jt] = sum_i P V_i.
//
// TMEM map: S ring 3 x 48 at 0, Y2 at 144, Y1 at 192, C fp32 at 256..511 (sweep A) -> C_bf16 at 256..383 and dG^T at
// 384..511 (sweep B); the fp32 -> bf16 conversion goes through columns 0..127 (idle between the sweeps).
// Warp roles as in pair_fwd3.cu.  Padded rows / columns need no special casing: alpha = 0 or beta = 0 there, so
// S = l = 0, their cosines are exactly 0 and every term that reaches a real row is multiplied by an exact zero.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dvlp {
using namespace ptx;

int make_map_items(CUtensorMap* m, const void* base, uint64_t n0, uint64_t n1, uint64_t n2, uint32_t b1, uint32_t b2);
int make_map_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols,
                CUtensorMapSwizzle swz);

namespace tb3 {
constexpr int kRows = 128;
constexpr int kTileBytes = 16384, kCh32 = 4096;
constexpr int kNVec = 3;

template <int NB, int NQ>
struct Lay {
  static constexpr int NG = NB / 16, NW = 4 * NG, kThreads = 128 * (1 + NG), TPT = 4 / NQ;
  static constexpr int NS = NQ == 1 ? 2 : 3;             // video-chunk ring stages (shared-memory budget)
  static constexpr int VBOX = NB * 128, VSTAGE = 4 * VBOX;
  static constexpr int UBYTES = NG * kCh32;
  static constexpr int VECB = NB * 8 + TPT * NB * 16;    // alpha[NB], m2v[NB], colv[TPT][NB] float4
  static constexpr int T = 0;
  static constexpr int GBD = T + 4 * kTileBytes;
  static constexpr int VR = GBD + 2 * kTileBytes;
  static constexpr int UA = VR + NS * VSTAGE;            // sweep A: U1 slots 0/1;  sweep B: slot 0 = UD2
  static constexpr int UB = UA + 2 * UBYTES;             // U2 slots 0/1 (both sweeps)
  static constexpr int VEC = UB + 2 * UBYTES;
  static constexpr int BETA = VEC + kNVec * VECB;
  static constexpr int M2T = BETA + 512;
  static constexpr int COLX = M2T + 512;                 // [2][NW][16]
  static constexpr int ROWX = COLX + 2 * NW * 64;        // [NG][128] float
  static constexpr int BARS = ROWX + NG * 512;
  static constexpr int TMEMPTR = BARS + 40 * 8;
  static constexpr int kAlloc = TMEMPTR + 16;
  static_assert(VSTAGE % 1024 == 0 && UBYTES % 1024 == 0 && VECB % 16 == 0, "alignment");
  static_assert(kAlloc <= 232448, "shared memory budget");
};

enum Bar { TJ_FULL = 0, UNIT_DONE, C_FULL, CB_READY, TB_DONE, Y_FULL, Y_FREE, DG_DONE, V_FULL /*3*/, V_EMPTY = V_FULL + 3,
           VEC_FULL = V_EMPTY + 3, VEC_EMPTY = VEC_FULL + kNVec, S_FULL = VEC_EMPTY + kNVec /*3*/, S_FREE = S_FULL + 3,
           U_READY = S_FREE + 3 /*2*/, kNumBars = U_READY + 2 };
static_assert(kNumBars <= 40, "barrier area");

constexpr uint32_t cS = 0, cY2 = 144, cY1 = 192, cC = 256, cCb = 256, cDGT = 384, cTmp = 0;

struct Args {
  int Bv, Bt, R, W, Rp, nch, tiles;
  long long total_units;
  const float *alpha3, *m2v3, *beta3, *m2t3, *d_scores;
  const float4 *rowv, *colv;
  float lam, lam2, inv_R, inv_W;
  __nv_bfloat16 *P, *UD1, *Cb;     // [units][128][Rp], [units][128][Rp], [units][128][256]
  float* dgt3;                     // [tiles][128][128] fp32, accumulated
};

__device__ __forceinline__ uint32_t desc_hi32(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (6u << 29);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// D[tmem] (+)= A[tmem, bf16 pairs per column, K-major] * B[smem desc]
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
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
__global__ void __launch_bounds__(Lay<NB, NQ>::kThreads, 1)
k_pair_bwd3(const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapT,
            const __grid_constant__ CUtensorMap mapGBD, Args a) {
  using L = Lay<NB, NQ>;
  constexpr int NG = L::NG, NW = L::NW, TPT = L::TPT, NS = L::NS;
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
      if (b == UNIT_DONE || b == CB_READY || b == TB_DONE || b == Y_FREE || (b >= VEC_EMPTY && b < VEC_EMPTY + kNVec) ||
          (b >= S_FREE && b < S_FREE + 3) || (b >= U_READY && b < U_READY + 2))
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
    // register budget: the four control warps keep 40 registers each, the compute warps take what that frees
    // (NB = 48: 384 threads x 152; NB = 32: 256 threads x 232 > the 255 cap, left alone)
    if (NG == 3) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // =============================== TMA producer ===============================
      if (elect_one()) {
        uint32_t nV = 0;
        int jt_prev = -1;
        for (int k = 0; k < n_units; ++k) {
          const long long u = u0 + k;
          const int jt = (int)(u / a.Bv), i = (int)(u - (long long)jt * a.Bv);
          if (jt != jt_prev) {
            if (k > 0) mbar_wait(bar(UNIT_DONE), (k - 1) & 1);
            mbar_arrive_expect_tx(bar(TJ_FULL), 6 * kTileBytes + 1024);
            for (int c = 0; c < 4; ++c) tma_load_3d(sbase + L::T + c * kTileBytes, &mapT, c * 64, 0, jt * TPT, bar(TJ_FULL));
            for (int c = 0; c < 2; ++c) tma_load_2d(sbase + L::GBD + c * kTileBytes, &mapGBD, c * 64, jt * kRows, bar(TJ_FULL));
            bulk_load(sbase + L::BETA, a.beta3 + (size_t)jt * kRows, 512, bar(TJ_FULL));
            bulk_load(sbase + L::M2T, a.m2t3 + (size_t)jt * kRows, 512, bar(TJ_FULL));
            jt_prev = jt;
          }
          for (int sweep = 0; sweep < 2; ++sweep)
            for (int c = 0; c < nch; ++c, ++nV) {
              const uint32_t st = nV % NS, vs = nV % kNVec;
              if (nV >= (uint32_t)NS) mbar_wait(bar(V_EMPTY + st), (nV / NS - 1) & 1);
              mbar_arrive_expect_tx(bar(V_FULL + st), L::VSTAGE);
              for (int kc = 0; kc < 4; ++kc)
                tma_load_3d(sbase + L::VR + st * L::VSTAGE + kc * L::VBOX, &mapV, kc * 64, c * NB, i, bar(V_FULL + st));
              if (nV >= (uint32_t)kNVec) mbar_wait(bar(VEC_EMPTY + vs), (nV / kNVec - 1) & 1);
              const uint32_t vb = sbase + L::VEC + vs * L::VECB;
              mbar_arrive_expect_tx(bar(VEC_FULL + vs), L::VECB);
              bulk_load(vb, a.alpha3 + (size_t)i * a.Rp + c * NB, NB * 4, bar(VEC_FULL + vs));
              bulk_load(vb + NB * 4, a.m2v3 + (size_t)i * a.Rp + c * NB, NB * 4, bar(VEC_FULL + vs));
              for (int tt = 0; tt < TPT; ++tt)
                bulk_load(vb + NB * 8 + tt * NB * 16, a.colv + ((size_t)u * TPT + tt) * a.Rp + c * NB, NB * 16, bar(VEC_FULL + vs));
            }
        }
      }
    } else if (warp == 1) {
      // ================================ MMA issuer ================================
      constexpr uint32_t idS = idesc_bf16(128, NB), idY2 = idesc_bf16(128, NB, 0, 1), idC = idesc_bf16(128, 256, 0, 1);
      constexpr uint32_t idDGT = idesc_bf16(128, 128);
      const uint32_t hi128 = desc_hi(1024), hi32 = desc_hi32(256);
      const uint32_t loT = desc_lo(sbase + L::T), loGBD = desc_lo(sbase + L::GBD);
      uint32_t nS = 0, nU = 0, nY = 0;
      int jt_prev = -1, tj_loads = 0;
      auto issue_S = [&](uint32_t n) {
        const uint32_t st = n % NS, slot = n % 3;
        mbar_wait(bar(V_FULL + st), (n / NS) & 1);
        if (n >= 3) mbar_wait(bar(S_FREE + slot), (n / 3 - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t loV = desc_lo(sbase + L::VR + st * L::VSTAGE);
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {
            const uint32_t oT = ((ks >> 2) * kTileBytes + (ks & 3) * 32) >> 4, oV = ((ks >> 2) * L::VBOX + (ks & 3) * 32) >> 4;
            mma_bf16_w(tmem + cS + slot * 48, loT + oT, hi128, loV + oV, hi128, idS, ks != 0);
          }
          tc_commit(bar(S_FULL + slot));
        }
        __syncwarp();
      };
      for (int k = 0; k < n_units; ++k) {
        const int jt = (int)((u0 + k) / a.Bv);
        if (jt != jt_prev) { mbar_wait(bar(TJ_FULL), tj_loads & 1); ++tj_loads; jt_prev = jt; }
        for (int sweep = 0; sweep < 2; ++sweep) {
          if (sweep == 1) { mbar_wait(bar(CB_READY), k & 1); tc_fence_after(); }   // C_bf16 is in place, columns 0..143 are idle
          issue_S(nS);
          for (int c = 0; c < nch; ++c, ++nS, ++nU, ++nY) {
            if (c + 1 < nch) issue_S(nS + 1);
            const uint32_t us = nU & 1, st = nS % NS;
            mbar_wait(bar(U_READY + us), (nU >> 1) & 1);
            if (sweep == 0 && c == 0 && k > 0) mbar_wait(bar(TB_DONE), (k - 1) & 1);   // the previous unit's C_bf16 / dG^T / Y are read
            tc_fence_after();
            const uint32_t loU2m = desc_lo(sbase + L::UB + us * L::UBYTES, kCh32);
            if (elect_one()) {
              if (sweep == 0) {
                // C += (u1 alpha)_c V_c
                const uint32_t loVm = desc_lo(sbase + L::VR + st * L::VSTAGE, L::VBOX);
                const uint32_t loU1 = desc_lo(sbase + L::UA + us * L::UBYTES);
#pragma unroll
                for (int ks = 0; ks < NG; ++ks)
                  mma_bf16_w(tmem + cC, loU1 + ((ks * kCh32) >> 4), hi32, loVm + ((ks * 2048) >> 4), hi128, idC, !(c == 0 && ks == 0));
              } else {
                // dG^T += (u2 dD)_c U2_c^T   (K = the NB regions of the chunk)
                const uint32_t loUD2 = desc_lo(sbase + L::UA), loU2k = desc_lo(sbase + L::UB + us * L::UBYTES);
#pragma unroll
                for (int ks = 0; ks < NG; ++ks)
                  mma_bf16_w(tmem + cDGT, loUD2 + ((ks * kCh32) >> 4), hi32, loU2k + ((ks * kCh32) >> 4), hi32, idDGT, !(c == 0 && ks == 0));
                tc_commit(bar(DG_DONE));
              }
            }
            __syncwarp();
            if (nY >= 1) mbar_wait(bar(Y_FREE), (nY - 1) & 1);      // the single Y slots have been read
            tc_fence_after();
            if (elect_one()) {
              if (sweep == 1) {
                // Y1_c = C_bf16 V_c^T: A operand from TMEM (8 columns per K step of 16)
                const uint32_t loV = desc_lo(sbase + L::VR + st * L::VSTAGE);
#pragma unroll
                for (int ks = 0; ks < 16; ++ks)
                  mma_bf16_ts(tmem + cY1, tmem + cCb + ks * 8, loV + (((ks >> 2) * L::VBOX + (ks & 3) * 32) >> 4), hi128, idS, ks != 0);
              }
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                mma_bf16_w(tmem + cY2, loGBD + (((ks >> 2) * kTileBytes + (ks & 3) * 32) >> 4), hi128, loU2m + ((ks * 512) >> 4),
                           hi32, idY2, ks != 0);
              tc_commit(bar(Y_FULL));
              tc_commit(bar(V_EMPTY + st));
              if (sweep == 0 && c == nch - 1) tc_commit(bar(C_FULL));
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ============================ compute warps ============================
    if (NG == 3) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int cw = warp - 4;
    const int g = cw >> 2, q = cw & 3;
    const int m = q * 32 + lane;
    const int cl = lane & 15;
    const int t = q / NQ, q0 = t * NQ;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t sw = (uint32_t)((m >> 2) & 1) << 4;
    const uint32_t row_off = (uint32_t)g * kCh32 + (uint32_t)m * 32;
    const uint32_t seg0 = row_off + sw, seg1 = row_off + (16u ^ sw);
    const float* s_beta = reinterpret_cast<const float*>(sgen + L::BETA);
    const float* s_m2t = reinterpret_cast<const float*>(sgen + L::M2T);
    float* s_colx = reinterpret_cast<float*>(sgen + L::COLX);
    float* s_rowx = reinterpret_cast<float*>(sgen + L::ROWX);
    uint32_t xcnt = 0;
    auto colsum = [&](float (&v)[16]) -> float {
      const float r = lanes_to_columns16(v, lane);
      if (NQ == 1) return r;
      float* base = s_colx + (xcnt & 1) * NW * 16;
      ++xcnt;
      if (lane < 16) base[cw * 16 + lane] = r;
      named_bar_sync(2 + g * TPT + t, 32 * NQ);
      float tot = 0.f;
#pragma unroll
      for (int qq = 0; qq < NQ; ++qq) tot += base[(g * 4 + q0 + qq) * 16 + cl];
      return tot;
    };
    uint32_t nS = 0, nU = 0, nY = 0, nD = 0;
    int jt_prev = -1, tj_loads = 0;
    float beta_m = 0.f, m2t_m = 0.f;
    const int kc_lo = 16 * g / NG, kc_hi = 16 * (g + 1) / NG;      // this warp's share of the 16 column groups of C

    for (int k = 0; k < n_units; ++k) {
      const long long u = u0 + k;
      const int jt = (int)(u / a.Bv), i = (int)(u - (long long)jt * a.Bv);
      if (jt != jt_prev) {
        mbar_wait(bar(TJ_FULL), tj_loads & 1); ++tj_loads; jt_prev = jt;
        beta_m = s_beta[m]; m2t_m = s_m2t[m];
      }
      // per-row scalars of the forward and the upstream gradient of this warp's (video, text)
      const float4 rv = a.rowv[(size_t)u * kRows + m];            // nB, thr1, 1/|c1|, cos1
      const int j = jt * TPT + t;
      const float gs = j < a.Bt ? a.d_scores[(size_t)i * a.Bt + j] : 0.f;
      const float nB = rv.x, thr1 = rv.y, nBl = rv.x * a.lam2;
      const float c1A = gs * a.inv_W * rv.z, kapA = rv.w * rv.z;
      const float gR = gs * a.inv_R;
      const float inv_lam2 = 1.f / a.lam2;
      __nv_bfloat16* gP = a.P + ((size_t)u * kRows + m) * a.Rp + g * 16;
      __nv_bfloat16* gUD1 = a.UD1 + ((size_t)u * kRows + m) * a.Rp + g * 16;

      // ============================ sweep A ============================
      float sB = 0.f;
      uint32_t slot_p = 0, vs_p = 0, us_p = 0, ny_p = 0;
      auto partA2 = [&]() {
        // sB += sum_r l F of the previous chunk, F = c1B u2 (S - kapB Y2)
        uint32_t uw[8];
        float x[16], y2[16];
        mbar_wait(bar(Y_FULL), ny_p & 1);
        tc_fence_after();
        tmem_ld16(tmem + cS + slot_p * 48 + lane_off + g * 16, reinterpret_cast<uint32_t*>(x));
        tmem_ld16(tmem + cY2 + lane_off + g * 16, reinterpret_cast<uint32_t*>(y2));
        ld_shared_v4(sbase + L::UB + us_p * L::UBYTES + seg0, uw[0], uw[1], uw[2], uw[3]);
        ld_shared_v4(sbase + L::UB + us_p * L::UBYTES + seg1, uw[4], uw[5], uw[6], uw[7]);
        tmem_ld_wait();
        tc_fence_before();
        const uint8_t* vb = sgen + L::VEC + vs_p * L::VECB;
        const float4* s_al = reinterpret_cast<const float4*>(vb) + g * 4;
        const float4* s_cv = reinterpret_cast<const float4*>(vb + NB * 8) + t * NB + g * 16;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 al = s_al[c4];
          const float av[4] = {al.x, al.y, al.z, al.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int cc = c4 * 4 + e;
            const float4 cv = s_cv[cc];                      // nA lam2, thr2, 1/|c2|, cos2
            const float xs = x[cc] * (av[e] * beta_m);
            const float l = leaky(xs);
            const float u2 = (cc & 1) ? bf16_hi(uw[cc >> 1]) : bf16_lo(uw[cc >> 1]);
            const float F = (gR * cv.z) * u2 * fmaf(-cv.w * cv.z, y2[cc], xs);
            sB = fmaf(l, F, sB);
          }
        }
        __syncwarp();
        if (lane == 0) { mbar_arrive(bar(S_FREE + slot_p)); mbar_arrive(bar(Y_FREE)); mbar_arrive(bar(VEC_EMPTY + vs_p)); }
      };
      for (int c = 0; c < nch; ++c, ++nS, ++nU, ++nY) {
        const uint32_t slot = nS % 3, vs = nS % kNVec, us = nU & 1;
        float x[16];
        uint32_t w1[8], w2[8];
        mbar_wait(bar(S_FULL + slot), (nS / 3) & 1);
        tc_fence_after();
        tmem_ld16(tmem + cS + slot * 48 + lane_off + g * 16, reinterpret_cast<uint32_t*>(x));
        tmem_ld_wait();
        tc_fence_before();
        mbar_wait(bar(VEC_FULL + vs), (nS / kNVec) & 1);
        const uint8_t* vb = sgen + L::VEC + vs * L::VECB;
        const float4* s_al = reinterpret_cast<const float4*>(vb) + g * 4;
        const float4* s_mv = reinterpret_cast<const float4*>(vb + NB * 4) + g * 4;
        const float4* s_cv = reinterpret_cast<const float4*>(vb + NB * 8) + t * NB + g * 16;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 al = s_al[c4], mv = s_mv[c4];
          const float av[4] = {al.x, al.y, al.z, al.w}, mvv[4] = {mv.x, mv.y, mv.z, mv.w};
          float ua[4], ub[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int cc = c4 * 4 + e;
            const float4 cv = s_cv[cc];
            const float l = leaky(x[cc] * (av[e] * beta_m));
            const float e1 = ex2(fmaf(l, cv.x, mvv[e]));
            const float e2 = ex2(fmaf(l, nBl, m2t_m));
            ua[e] = e1 > thr1 ? e1 * av[e] : 0.f;
            ub[e] = e2 > cv.y ? e2 : 0.f;
          }
          w1[c4 * 2] = pack_bf16(ua[0], ua[1]); w1[c4 * 2 + 1] = pack_bf16(ua[2], ua[3]);
          w2[c4 * 2] = pack_bf16(ub[0], ub[1]); w2[c4 * 2 + 1] = pack_bf16(ub[2], ub[3]);
        }
        st_shared_v4(sbase + L::UA + us * L::UBYTES + seg0, w1[0], w1[1], w1[2], w1[3]);
        st_shared_v4(sbase + L::UA + us * L::UBYTES + seg1, w1[4], w1[5], w1[6], w1[7]);
        st_shared_v4(sbase + L::UB + us * L::UBYTES + seg0, w2[0], w2[1], w2[2], w2[3]);
        st_shared_v4(sbase + L::UB + us * L::UBYTES + seg1, w2[4], w2[5], w2[6], w2[7]);
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(U_READY + us));
        if (c > 0) partA2();
        slot_p = slot; vs_p = vs; us_p = us; ny_p = nY;
      }
      partA2();
      // row totals of sB over all chunks
      s_rowx[g * kRows + m] = sB;
      named_bar_sync(1, 128 * NG);
      float zB = 0.f;
#pragma unroll
      for (int gg = 0; gg < NG; ++gg) zB += s_rowx[gg * kRows + m];
      zB *= nB * nB * nB;
      // ---- C (fp32) -> bf16: spill for grad_gemm3.cu and, through the idle columns 0..127, the A operand of the Y1 products
      mbar_wait(bar(C_FULL), k & 1);
      tc_fence_after();
      {
        __nv_bfloat16* gC = a.Cb + ((size_t)u * kRows + m) * 256;
        for (int kc = kc_lo; kc < kc_hi; ++kc) {
          float v[16];
          uint32_t pk[8];
          tmem_ld16(tmem + cC + lane_off + kc * 16, reinterpret_cast<uint32_t*>(v));
          tmem_ld_wait();
#pragma unroll
          for (int p2 = 0; p2 < 8; ++p2) pk[p2] = pack_bf16(v[2 * p2], v[2 * p2 + 1]);
          tmem_st8(tmem + cTmp + lane_off + kc * 8, pk);
          st_global_v4(gC + kc * 16, pk[0], pk[1], pk[2], pk[3]);
          st_global_v4(gC + kc * 16 + 8, pk[4], pk[5], pk[6], pk[7]);
        }
        tmem_st_wait();
        tc_fence_before();
        named_bar_sync(1, 128 * NG);          // every fp32 column has been read: the bf16 copy may move in
        tc_fence_after();
        for (int kc = kc_lo; kc < kc_hi; ++kc) {
          uint32_t pk[8];
          tmem_ld8(tmem + cTmp + lane_off + kc * 8, pk);
          tmem_ld_wait();
          tmem_st8(tmem + cCb + lane_off + kc * 8, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(CB_READY));
      }

      // ============================ sweep B ============================
      auto partB2 = [&](int c_p) {
        float x[16], pa[16], aa[16], bb[16];
        uint32_t uw[8];
        mbar_wait(bar(Y_FULL), ny_p & 1);
        tc_fence_after();
        tmem_ld16(tmem + cS + slot_p * 48 + lane_off + g * 16, reinterpret_cast<uint32_t*>(x));
        ld_shared_v4(sbase + L::UB + us_p * L::UBYTES + seg0, uw[0], uw[1], uw[2], uw[3]);
        ld_shared_v4(sbase + L::UB + us_p * L::UBYTES + seg1, uw[4], uw[5], uw[6], uw[7]);
        const uint8_t* vb = sgen + L::VEC + vs_p * L::VECB;
        const float4* s_al = reinterpret_cast<const float4*>(vb) + g * 4;
        const float4* s_mv = reinterpret_cast<const float4*>(vb + NB * 4) + g * 4;
        const float4* s_cv = reinterpret_cast<const float4*>(vb + NB * 8) + t * NB + g * 16;
        uint32_t wd[8];
        const float dDA2 = -c1A * kapA;                         // 2 dD1 = d loss / d |c_w|^2 times two
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float y1[8], y2[8];
          tmem_ld8(tmem + cY1 + lane_off + g * 16 + half * 8, reinterpret_cast<uint32_t*>(y1));
          tmem_ld8(tmem + cY2 + lane_off + g * 16 + half * 8, reinterpret_cast<uint32_t*>(y2));
          tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < 2; ++c4) {
            const float4 al = s_al[half * 2 + c4], mv = s_mv[half * 2 + c4];
            const float av[4] = {al.x, al.y, al.z, al.w}, mvv[4] = {mv.x, mv.y, mv.z, mv.w};
            float uav[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int cc = half * 8 + c4 * 4 + e, ch = c4 * 4 + e;
              const float4 cv = s_cv[cc];
              const float ab = av[e] * beta_m;
              const float xs = x[cc] * ab;
              const float l = leaky(xs);
              const float e1 = ex2(fmaf(l, cv.x, mvv[e]));
              const float u1 = e1 > thr1 ? e1 : 0.f;
              const float u2 = (cc & 1) ? bf16_hi(uw[cc >> 1]) : bf16_lo(uw[cc >> 1]);
              const float c1B = gR * cv.z, kapB = cv.w * cv.z;
              const float G = c1A * u1 * fmaf(-kapA * av[e], y1[ch], xs);       // Y1 = alpha_r <c_w, v_r>
              const float F = c1B * u2 * fmaf(-kapB, y2[ch], xs);
              const float nA = cv.x * inv_lam2;
              pa[cc] = l * G;
              const float lg = xs > 0.f ? a.lam : kLeaky * a.lam;              // leaky'(S) lambda
              aa[cc] = fmaf(lg, fmaf(G, nA, F * nB), fmaf(c1A, u1, c1B * u2)) * ab;
              bb[cc] = lg * l * ab;
              uav[e] = u1 * av[e] * dDA2;
            }
            wd[half * 4 + c4 * 2] = pack_bf16(uav[0], uav[1]);
            wd[half * 4 + c4 * 2 + 1] = pack_bf16(uav[2], uav[3]);
          }
        }
        tc_fence_before();
        st_global_v4(gUD1 + c_p * NB, wd[0], wd[1], wd[2], wd[3]);
        st_global_v4(gUD1 + c_p * NB + 8, wd[4], wd[5], wd[6], wd[7]);
        __syncwarp();
        if (lane == 0) { mbar_arrive(bar(S_FREE + slot_p)); mbar_arrive(bar(Y_FREE)); }
        // column sums sA = sum_w l G over the words of the text; nA^3 sA per column
        const float sA = colsum(pa);
        float nA3;
        {
          const float nAc = s_cv[cl].x * inv_lam2;
          nA3 = nAc * nAc * nAc * sA;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(VEC_EMPTY + vs_p));
        uint32_t pk[8];
#pragma unroll
        for (int cc = 0; cc < 16; cc += 2) {
          const float t0 = __shfl_sync(0xffffffffu, nA3, cc) + zB, t1 = __shfl_sync(0xffffffffu, nA3, cc + 1) + zB;
          pk[cc >> 1] = pack_bf16(fmaf(-bb[cc], t0, aa[cc]), fmaf(-bb[cc + 1], t1, aa[cc + 1]));
        }
        st_global_v4(gP + c_p * NB, pk[0], pk[1], pk[2], pk[3]);
        st_global_v4(gP + c_p * NB + 8, pk[4], pk[5], pk[6], pk[7]);
      };
      int c_prev = 0;
      for (int c = 0; c < nch; ++c, ++nS, ++nU, ++nY, ++nD) {
        const uint32_t slot = nS % 3, vs = nS % kNVec, us = nU & 1;
        float x[16];
        uint32_t w2[8], wd[8];
        mbar_wait(bar(S_FULL + slot), (nS / 3) & 1);
        tc_fence_after();
        tmem_ld16(tmem + cS + slot * 48 + lane_off + g * 16, reinterpret_cast<uint32_t*>(x));
        tmem_ld_wait();
        tc_fence_before();
        mbar_wait(bar(VEC_FULL + vs), (nS / kNVec) & 1);
        const uint8_t* vb = sgen + L::VEC + vs * L::VECB;
        const float4* s_al = reinterpret_cast<const float4*>(vb) + g * 4;
        const float4* s_cv = reinterpret_cast<const float4*>(vb + NB * 8) + t * NB + g * 16;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 al = s_al[c4];
          const float av[4] = {al.x, al.y, al.z, al.w};
          float ub[4], dd[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int cc = c4 * 4 + e;
            const float4 cv = s_cv[cc];
            const float l = leaky(x[cc] * (av[e] * beta_m));
            const float e2 = ex2(fmaf(l, nBl, m2t_m));
            ub[e] = e2 > cv.y ? e2 : 0.f;
            dd[e] = -0.5f * (gR * cv.z) * (cv.w * cv.z);       // dD2 of this (text, region)
          }
          const uint32_t p0 = pack_bf16(ub[0], ub[1]), p1 = pack_bf16(ub[2], ub[3]);
          w2[c4 * 2] = p0; w2[c4 * 2 + 1] = p1;
          wd[c4 * 2] = pack_bf16(bf16_lo(p0) * dd[0], bf16_hi(p0) * dd[1]);
          wd[c4 * 2 + 1] = pack_bf16(bf16_lo(p1) * dd[2], bf16_hi(p1) * dd[3]);
        }
        if (nD >= 1) mbar_wait(bar(DG_DONE), (nD - 1) & 1);    // the single UD2 tile has been consumed
        st_shared_v4(sbase + L::UA + seg0, wd[0], wd[1], wd[2], wd[3]);
        st_shared_v4(sbase + L::UA + seg1, wd[4], wd[5], wd[6], wd[7]);
        st_shared_v4(sbase + L::UB + us * L::UBYTES + seg0, w2[0], w2[1], w2[2], w2[3]);
        st_shared_v4(sbase + L::UB + us * L::UBYTES + seg1, w2[4], w2[5], w2[6], w2[7]);
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(U_READY + us));
        if (c > 0) partB2(c_prev);
        slot_p = slot; vs_p = vs; us_p = us; ny_p = nY; c_prev = c;
      }
      partB2(c_prev);
      // ---- text Gram gradient of this unit leaves TMEM: the diagonal WP x WP block of this warp's text
      mbar_wait(bar(DG_DONE), (nD - 1) & 1);
      tc_fence_after();
      {
        float* dst = a.dgt3 + ((size_t)jt * kRows + m) * kRows + t * (32 * NQ);
        for (int cgp = g; cgp < 2 * NQ; cgp += NG) {
          float v[16];
          tmem_ld16(tmem + cDGT + lane_off + t * (32 * NQ) + cgp * 16, reinterpret_cast<uint32_t*>(v));
          tmem_ld_wait();
          if (gs != 0.f) {
#pragma unroll
            for (int e = 0; e < 16; e += 4) red_add_v4(dst + cgp * 16 + e, v[e], v[e + 1], v[e + 2], v[e + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(TB_DONE)); mbar_arrive(bar(UNIT_DONE)); }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem, 512);
}

template <int NB, int NQ>
static int launch(const CUtensorMap& mV, const CUtensorMap& mT, const CUtensorMap& mG, const Args& a, cudaStream_t st) {
  using L = Lay<NB, NQ>;
  int dev = 0, sms = 148;
  DVLP_CUDA_OK(cudaGetDevice(&dev));
  DVLP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = (int)(a.total_units < sms ? a.total_units : sms);
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    DVLP_CUDA_OK(cudaFuncSetAttribute(k_pair_bwd3<NB, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc));
    attr_once.done();
  }
  k_pair_bwd3<NB, NQ><<<grid, L::kThreads, L::kAlloc, st>>>(mV, mT, mG, a);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}
}  // namespace tb3

int run_pair_bwd3(const void* vb, const void* tb, int Bv, int Bt, int R, int W, int d, float lambda,
                  const float* d_scores, char* ws, const Workspace& Lw, cudaStream_t st) {
  Shape3 s3;
  if (!shape3(R, W, d, &s3)) return DVLP_ERR_BAD_SHAPE;
  const int tiles = (Bt + s3.TPT - 1) / s3.TPT;
  CUtensorMap mV, mT, mG;
  int rc;
  if ((rc = make_map_items(&mV, vb, 256, (uint64_t)R, (uint64_t)Bv, (uint32_t)s3.NB, 1))) return rc;
  if ((rc = make_map_items(&mT, tb, 256, (uint64_t)W, (uint64_t)Bt, (uint32_t)s3.WP, (uint32_t)s3.TPT))) return rc;
  if ((rc = make_map_2d(&mG, ws + Lw.gbd3, (uint64_t)tiles * 128, 128, 128, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  tb3::Args a{};
  a.Bv = Bv; a.Bt = Bt; a.R = R; a.W = W; a.Rp = s3.Rp; a.nch = s3.nch; a.tiles = tiles;
  a.total_units = (long long)tiles * Bv;
  a.alpha3 = (const float*)(ws + Lw.alpha3); a.m2v3 = (const float*)(ws + Lw.m2v3);
  a.beta3 = (const float*)(ws + Lw.beta3); a.m2t3 = (const float*)(ws + Lw.m2t3);
  a.d_scores = d_scores;
  a.rowv = (const float4*)(ws + Lw.rowv); a.colv = (const float4*)(ws + Lw.colv);
  a.lam = lambda; a.lam2 = lambda * kLog2e; a.inv_R = 1.f / (float)R; a.inv_W = 1.f / (float)W;
  a.P = (__nv_bfloat16*)(ws + Lw.pa); a.UD1 = (__nv_bfloat16*)(ws + Lw.ud1); a.Cb = (__nv_bfloat16*)(ws + Lw.cb);
  a.dgt3 = (float*)(ws + Lw.dgt3);
  if (s3.NB == 48) {
    if (s3.NQ == 1) return tb3::launch<48, 1>(mV, mT, mG, a, st);
    if (s3.NQ == 2) return tb3::launch<48, 2>(mV, mT, mG, a, st);
    return tb3::launch<48, 4>(mV, mT, mG, a, st);
  }
  if (s3.NQ == 1) return tb3::launch<32, 1>(mV, mT, mG, a, st);
  if (s3.NQ == 2) return tb3::launch<32, 2>(mV, mT, mG, a, st);
  return tb3::launch<32, 4>(mV, mT, mG, a, st);
}

}  // namespace dvlp
