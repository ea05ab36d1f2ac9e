// Region-word pair scores on tcgen05 / TMEM / TMA (sm_100a): FORWARD kernel, single-orientation design
// (partner of pair_bwd2.cu; reference arithmetic: model/loss.py:209-330, restated in DESIGN.md section 3).
//
// One persistent CTA per SM walks (video, 4-text tile) units.  TMEM lane = word row (4 texts x 32 words),
// TMEM column = region; the R columns are split into 16-column chunks, one compute warpgroup per chunk, all
// running the same code, so each element (word, region) is owned by one thread that evaluates BOTH attention
// directions for it:
//   i2t (context = regions, query = words):  softmax / focal sums over regions = row sums (per-thread partials
//        exchanged between the warpgroups through shared memory), l2 norm over the words of a text = sum over the
//        32 lanes of a warp (shuffle transpose-reduction, lane l ends with column l & 15);
//   t2i (context = words, query = regions):  the other way round.
// Tensor-core products per unit:  S = T V^T (double-buffered in TMEM, the next tile's S runs under this tile's
// softmax work),  Y1 = U1 G^V,  Y2 = Gbd U2 (Gram-form |weighted context|^2).  U1 / U2 are
// [word rows][region columns] bf16 tiles in the 32-byte-swizzle layout, written row-wise by their owners.
// Besides the scores the kernel saves the per-query scalars the backward re-uses (rowv / colv, common.cuh).
//
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warp 3 TMEM owner, warps 4.. compute
// (warpgroup g = column chunk, warp q of the group = text slot = TMEM lane quadrant).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dvlp {
using namespace ptx;

namespace tf2 {
constexpr int kD = 256, kW = 32, kTPT = 4, kRows = 128;
constexpr int kTileBytes = 16384, kCh32 = 4096;

template <int R>
struct Lay {
  static constexpr int NC = R / 16, NW = 4 * NC, kThreads = 128 * (1 + NC);
  static constexpr int kVChunk = R * 128;
  static constexpr int V = 0;                       // 4 chunks (K = 256)
  static constexpr int GV = V + 4 * kVChunk;        // 2 chunks
  static constexpr int T = GV + 2 * kVChunk;        // 4 x 16 KB
  static constexpr int GBD = T + 4 * kTileBytes;    // 2 x 16 KB
  static constexpr int U1 = GBD + 2 * kTileBytes;   // NC x 4 KB
  static constexpr int U2 = U1 + NC * kCh32;
  static constexpr int BETA = U2 + NC * kCh32;      // [128]
  static constexpr int M2T = BETA + 512;            // [128]
  static constexpr int ALPHA = M2T + 512;           // [R]
  static constexpr int M2V = ALPHA + R * 4;         // [R]
  static constexpr int TAB = M2V + R * 4;           // per compute warp [16] float4 (alpha, m2v, nA lam2, thr2)
  static constexpr int P1 = TAB + NW * 256;         // [NC][128] row partials: sum l^2
  static constexpr int P2 = P1 + NC * 512;          // [NC][128] row partials: sum e1
  static constexpr int P3 = P2 + NC * 512;          // [NC][128] float2 row partials: num1, D1
  static constexpr int T2I = P3 + NC * 1024;        // [NC][4] per-warp sums of cos2 over the warp's 16 regions
  static constexpr int BARS = T2I + NC * 16;
  static constexpr int TMEMPTR = BARS + 16 * 8;
  static constexpr int kAlloc = TMEMPTR + 16;
  static_assert(kAlloc <= 232448, "shared memory budget");
};

enum Bar { V_FULL = 0, T_FULL, GBD_FULL, VEC_FULL, VEC_FREE, S_DONE0, S_DONE1, U_READY, Y_DONE, TILE_DONE, kNumBars };

// TMEM columns: two S buffers, Y1, Y2
constexpr uint32_t cS0 = 0, cS1 = 96, cY1 = 192, cY2 = 288;

struct Args {
  int Bv, Bt, tiles_per_video;
  long long total_units;
  const float *alpha, *beta, *m2v, *m2t;
  float lam2;        // lambda * log2(e)
  int focal;
  float* scores;
  float4 *rowv, *colv;    // per-unit query scalars for the backward; may be NULL (inference)
};

__device__ __forceinline__ uint32_t desc_hi32(uint32_t sbo_bytes) {      // 32-byte swizzle
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (6u << 29);
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
// Sum over the 32 lanes of v[c], c = 0..15: lane l ends with the total of column l & 15.
// Lane bit b is paired with column bit b; the columns a lane still holds after a step are those whose
// bit matches its own, so 16 -> 8 -> 4 -> 2 -> 1 values, then lanes l and l ^ 16 add up.
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

template <int R>
__global__ void __launch_bounds__(Lay<R>::kThreads, 1)
k_pair_fwd2(const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapT,
            const __grid_constant__ CUtensorMap mapGV, const __grid_constant__ CUtensorMap mapGBD, Args a) {
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
  const long long u0 = a.total_units * blockIdx.x / gridDim.x;
  const long long u1 = a.total_units * (blockIdx.x + 1) / gridDim.x;
  const int n_tiles = (int)(u1 - u0);
  const int tpv = a.tiles_per_video;
  const int i0 = (int)(u0 / tpv), jt0 = (int)(u0 % tpv);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapV); prefetch_tmap(&mapT); prefetch_tmap(&mapGV); prefetch_tmap(&mapGBD);
    for (int b = 0; b < kNumBars; ++b) {
      uint32_t cnt = 1;
      if (b == VEC_FREE || b == U_READY || b == TILE_DONE) cnt = NW;
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
    if (NC == 5) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 0) {
      // =============================== TMA producer ===============================
      if (lane == 0) {
        int i = i0, jt = jt0;
        for (int k = 0; k < n_tiles; ++k) {
          const uint32_t pp = (k - 1) & 1;
          // per-word vectors of the text tile: free once every warp has read its scalars
          if (k > 0) mbar_wait(bar(VEC_FREE), pp);
          if (k == 0 || jt == 0) {
            // the resident video changes: every warp is done with tile k-1 (tables hold alpha / m2v copies, the
            // S GEMM of tile k-1 read V, its Y1 GEMM read G^V)
            if (k > 0) mbar_wait(bar(TILE_DONE), pp);
            mbar_arrive_expect_tx(bar(V_FULL), 6 * L::kVChunk + 2 * R * 4);
            for (int c = 0; c < 4; ++c) tma_load_2d(sbase + L::V + c * L::kVChunk, &mapV, c * 64, i * R, bar(V_FULL));
            for (int c = 0; c < 2; ++c) tma_load_2d(sbase + L::GV + c * L::kVChunk, &mapGV, c * 64, i * R, bar(V_FULL));
            bulk_load(sbase + L::ALPHA, a.alpha + (size_t)i * R, R * 4, bar(V_FULL));
            bulk_load(sbase + L::M2V, a.m2v + (size_t)i * R, R * 4, bar(V_FULL));
          }
          mbar_arrive_expect_tx(bar(VEC_FULL), 1024);
          bulk_load(sbase + L::BETA, a.beta + (size_t)jt * kRows, 512, bar(VEC_FULL));
          bulk_load(sbase + L::M2T, a.m2t + (size_t)jt * kRows, 512, bar(VEC_FULL));
          // text tile: the S GEMM of the previous tile has read it
          if (k > 0) mbar_wait(bar(S_DONE0 + pp), ((k - 1) >> 1) & 1);
          mbar_arrive_expect_tx(bar(T_FULL), 4 * kTileBytes);
          for (int c = 0; c < 4; ++c) tma_load_2d(sbase + L::T + c * kTileBytes, &mapT, c * 64, jt * kRows, bar(T_FULL));
          // block-diagonal text Gram tile: the Y2 GEMM of the previous tile has read it
          if (k > 0) mbar_wait(bar(Y_DONE), pp);
          mbar_arrive_expect_tx(bar(GBD_FULL), 2 * kTileBytes);
          for (int c = 0; c < 2; ++c) tma_load_2d(sbase + L::GBD + c * kTileBytes, &mapGBD, c * 64, jt * kRows, bar(GBD_FULL));
          if (++jt == tpv) { jt = 0; ++i; }
        }
      }
    } else if (warp == 1) {
      // ================================ MMA issuer ================================
      constexpr uint32_t idS = idesc_bf16(128, R), idY2 = idesc_bf16(128, R, 0, 1);
      constexpr uint32_t CH = L::kVChunk;
      const uint32_t hi128 = desc_hi(1024), hi32 = desc_hi32(256);
      const uint32_t loT = desc_lo(sbase + L::T), loV = desc_lo(sbase + L::V), loGV = desc_lo(sbase + L::GV);
      const uint32_t loGBD = desc_lo(sbase + L::GBD);
      const uint32_t loU1k = desc_lo(sbase + L::U1), loU2m = desc_lo(sbase + L::U2, kCh32);
      int vloads = 0, jt = jt0;
      auto issue_S = [&](int k, bool sw) {
        if (sw) { mbar_wait(bar(V_FULL), vloads & 1); ++vloads; }
        mbar_wait(bar(T_FULL), k & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t dS = tmem + ((k & 1) ? cS1 : cS0);
#pragma unroll
          for (int ks = 0; ks < 16; ++ks) {
            const uint32_t oT = ((ks >> 2) * kTileBytes + (ks & 3) * 32) >> 4, oV = ((ks >> 2) * CH + (ks & 3) * 32) >> 4;
            mma_bf16_w(dS, loT + oT, hi128, loV + oV, hi128, idS, ks != 0);
          }
          tc_commit(bar(S_DONE0 + (k & 1)));
        }
        __syncwarp();
      };
      issue_S(0, true);
      for (int k = 0; k < n_tiles; ++k) {
        const uint32_t ph = k & 1;
        if (++jt == tpv) jt = 0;                       // jt of tile k+1
        const bool more = k + 1 < n_tiles;
        // S of the next tile runs ahead (its TMEM buffer was last read in tile k-1) unless the video changes:
        // the new video's tile can only be loaded once tile k is finished
        const bool ahead = more && jt != 0;
        if (ahead) issue_S(k + 1, false);
        mbar_wait(bar(GBD_FULL), ph);
        mbar_wait(bar(U_READY), ph);
        tc_fence_after();
        if (lane == 0) {
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
        if (more && !ahead) issue_S(k + 1, true);
      }
    }
  } else {
    // ============================ compute warps ============================
    if (NC == 5) asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    const int cw = warp - 4;
    const int g = cw >> 2, q = cw & 3;          // column chunk, text slot / TMEM lane quadrant
    const int m = q * 32 + lane;                // word row of the tile
    const int cl = lane & 15;                   // column whose lane-reduced sums this lane ends up holding
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t sw = (uint32_t)((m >> 2) & 1) << 4;
    const uint32_t row_off = (uint32_t)g * kCh32 + (uint32_t)m * 32;
    const uint32_t seg0 = row_off + sw, seg1 = row_off + (16u ^ sw);
    float4* tab = reinterpret_cast<float4*>(sgen + L::TAB + cw * 256);
    const float* s_beta = reinterpret_cast<const float*>(sgen + L::BETA);
    const float* s_m2t = reinterpret_cast<const float*>(sgen + L::M2T);
    const float* s_alpha = reinterpret_cast<const float*>(sgen + L::ALPHA);
    const float* s_m2v = reinterpret_cast<const float*>(sgen + L::M2V);
    float* s_p1 = reinterpret_cast<float*>(sgen + L::P1);
    float* s_p2 = reinterpret_cast<float*>(sgen + L::P2);
    float2* s_p3 = reinterpret_cast<float2*>(sgen + L::P3);
    float* s_t2i = reinterpret_cast<float*>(sgen + L::T2I);
    int vloads = 0, i = i0, jt = jt0;
    for (int k = 0; k < n_tiles; ++k) {
      const uint32_t ph = k & 1;
      const int j = jt * kTPT + q;
      if (k == 0 || jt == 0) { mbar_wait(bar(V_FULL), vloads & 1); ++vloads; }
      mbar_wait(bar(VEC_FULL), ph);
      const float beta_m = s_beta[m], m2t_m = s_m2t[m];
      if (lane < 16) tab[cl] = make_float4(s_alpha[g * 16 + cl], s_m2v[g * 16 + cl], 0.f, 0.f);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(VEC_FREE));

      // ---------------- pass 0: S = raw alpha beta, the two query-axis l2 norms of model/loss.py:238 ----------------
      mbar_wait(bar(S_DONE0 + ph), (k >> 1) & 1);
      tc_fence_after();
      float x[16];
      tmem_ld16(tmem + (ph ? cS1 : cS0) + lane_off + g * 16, reinterpret_cast<uint32_t*>(x));
      tmem_ld_wait();
      float nA_cl, nB;
      {
        float sq[16], ssB = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          x[c] *= tab[c].x * beta_m;
          const float l = leaky(x[c]);
          sq[c] = l * l;
          ssB += sq[c];
        }
        s_p1[g * kRows + m] = ssB;
        const float ssA = lanes_to_columns16(sq, lane);      // sum over the 32 words of this text, column cl
        nA_cl = 1.f / (sqrtf(ssA) + kEps);
        if (lane < 16) reinterpret_cast<float*>(tab + cl)[2] = nA_cl * a.lam2;
        named_bar_sync(1 + q, 32 * NC);                      // row partials of the NC warps sharing this text slot
        float tot = 0.f;
#pragma unroll
        for (int gg = 0; gg < NC; ++gg) tot += s_p1[gg * kRows + m];
        nB = 1.f / (sqrtf(tot) + kEps);
      }
      // ---------------- pass 1: exponentials, focal thresholds ----------------
      float eA[16], eB[16];
      float thr1, thr2_cl;
      {
        const float nBl = nB * a.lam2;
        float sumA = 0.f, eb2[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float4 t = tab[c];              // alpha, m2v, nA lam2
          const float l = leaky(x[c]);
          eA[c] = ex2(fmaf(l, t.z, t.y));
          eB[c] = ex2(fmaf(l, nBl, m2t_m));
          sumA += eA[c];
          eb2[c] = eB[c];
        }
        s_p2[g * kRows + m] = sumA;
        const float sumB = lanes_to_columns16(eb2, lane);
        thr2_cl = a.focal ? sumB * (1.f / kW) : -1.f;        // keep e iff e * W > sum_w e  (model/loss.py:280-282)
        if (lane < 16) reinterpret_cast<float*>(tab + cl)[3] = thr2_cl;
        named_bar_sync(1 + q, 32 * NC);
        float tot = 0.f;
#pragma unroll
        for (int gg = 0; gg < NC; ++gg) tot += s_p2[gg * kRows + m];
        thr1 = a.focal ? tot * (1.f / R) : -1.f;
      }
      // ---------------- pass 2: focal-selected u (bf16), operand tiles, numerators ----------------
      uint32_t ub[16];                          // lo = u1 (i2t), hi = u2 (t2i)
      float num1 = 0.f, num2_cl;
      {
        float pn[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          ub[c] = pack_bf16(eA[c] > thr1 ? eA[c] : 0.f, eB[c] > tab[c].w ? eB[c] : 0.f);
          num1 = fmaf(bf16_lo(ub[c]), x[c], num1);
          pn[c] = bf16_hi(ub[c]) * x[c];
        }
        if (k > 0) mbar_wait(bar(Y_DONE), (k - 1) & 1);      // the Y GEMMs of tile k-1 have read U1 / U2
        st_shared_v4(sbase + L::U1 + seg0, prmt(ub[0], ub[1], 0x5410), prmt(ub[2], ub[3], 0x5410), prmt(ub[4], ub[5], 0x5410),
                     prmt(ub[6], ub[7], 0x5410));
        st_shared_v4(sbase + L::U1 + seg1, prmt(ub[8], ub[9], 0x5410), prmt(ub[10], ub[11], 0x5410),
                     prmt(ub[12], ub[13], 0x5410), prmt(ub[14], ub[15], 0x5410));
        st_shared_v4(sbase + L::U2 + seg0, prmt(ub[0], ub[1], 0x7632), prmt(ub[2], ub[3], 0x7632), prmt(ub[4], ub[5], 0x7632),
                     prmt(ub[6], ub[7], 0x7632));
        st_shared_v4(sbase + L::U2 + seg1, prmt(ub[8], ub[9], 0x7632), prmt(ub[10], ub[11], 0x7632),
                     prmt(ub[12], ub[13], 0x7632), prmt(ub[14], ub[15], 0x7632));
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(U_READY));
        num2_cl = lanes_to_columns16(pn, lane);               // under the Y GEMMs
      }
      // ---------------- pass 3: Gram-form denominators, cosines, score ----------------
      mbar_wait(bar(Y_DONE), ph);
      tc_fence_after();
      float D1 = 0.f, D2_cl;
      {
        float y[16], pd[16];
        tmem_ld16(tmem + cY1 + lane_off + g * 16, reinterpret_cast<uint32_t*>(y));
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) D1 = fmaf(bf16_lo(ub[c]), y[c], D1);
        tmem_ld16(tmem + cY2 + lane_off + g * 16, reinterpret_cast<uint32_t*>(y));
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) pd[c] = bf16_hi(ub[c]) * y[c];
        tc_fence_before();
        D2_cl = lanes_to_columns16(pd, lane);
      }
      s_p3[g * kRows + m] = make_float2(num1, D1);
      {
        // t2i: cos2 of this warp's 16 (text, region) queries, summed over the regions (model/loss.py:326-327)
        const float rden2 = D2_cl > 1e-30f ? rsqrtf(D2_cl) : 0.f;
        float cos2 = num2_cl * rden2;
        if (a.colv && lane < 16)
          a.colv[((size_t)(u0 + k) * kTPT + q) * R + g * 16 + cl] = make_float4(nA_cl, thr2_cl, rden2, cos2);
        cos2 += __shfl_xor_sync(0xffffffffu, cos2, 8);
        cos2 += __shfl_xor_sync(0xffffffffu, cos2, 4);
        cos2 += __shfl_xor_sync(0xffffffffu, cos2, 2);
        cos2 += __shfl_xor_sync(0xffffffffu, cos2, 1);
        if (lane == 0) s_t2i[g * 4 + q] = cos2;
      }
      named_bar_sync(1 + q, 32 * NC);
      if (g == 0) {
        float n = 0.f, d = 0.f;
#pragma unroll
        for (int gg = 0; gg < NC; ++gg) { const float2 p = s_p3[gg * kRows + m]; n += p.x; d += p.y; }
        const float rden1 = d > 1e-30f ? rsqrtf(d) : 0.f;
        float cs = n * rden1;
        if (a.rowv) a.rowv[(size_t)(u0 + k) * kRows + m] = make_float4(nB, thr1, rden1, cs);
        cs = warp_sum(cs);                      // mean over the 32 words of text q (model/loss.py:318)
        if (lane == 0 && j < a.Bt) {
          float t2i = 0.f;
#pragma unroll
          for (int gg = 0; gg < NC; ++gg) t2i += s_t2i[gg * 4 + q];
          a.scores[(size_t)i * a.Bt + j] = cs * (1.f / kW) + t2i * (1.f / R);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(TILE_DONE));
      if (++jt == tpv) { jt = 0; ++i; }
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
static int launch(const void* vb, const void* tb, int Bv, int Bt, float lambda, int focal, float* scores, char* ws,
                  const Workspace& Lw, bool save, cudaStream_t st) {
  using L = Lay<R>;
  CUtensorMap mV, mT, mGV, mGBD;
  int rc;
  Args a{};
  a.Bv = Bv; a.Bt = Bt;
  a.tiles_per_video = (Bt + kTPT - 1) / kTPT;
  a.total_units = (long long)Bv * a.tiles_per_video;
  if ((rc = make_map(&mV, vb, (uint64_t)Bv * R, kD, R))) return rc;
  if ((rc = make_map(&mT, tb, (uint64_t)Bt * kW, kD, kRows))) return rc;
  if ((rc = make_map(&mGV, ws + Lw.gvb, (uint64_t)Bv * R, R, R))) return rc;
  if ((rc = make_map(&mGBD, ws + Lw.gbd, (uint64_t)a.tiles_per_video * kRows, kRows, kRows))) return rc;
  a.alpha = (const float*)(ws + Lw.alpha); a.beta = (const float*)(ws + Lw.beta);
  a.m2v = (const float*)(ws + Lw.m2v); a.m2t = (const float*)(ws + Lw.m2t);
  a.lam2 = lambda * kLog2e; a.focal = focal; a.scores = scores;
  a.rowv = save ? (float4*)(ws + Lw.rowv) : nullptr;
  a.colv = save ? (float4*)(ws + Lw.colv) : nullptr;
  int dev = 0, sms = 148;
  DVLP_CUDA_OK(cudaGetDevice(&dev));
  DVLP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = (int)(a.total_units < sms ? a.total_units : sms);
  static bool attr_set = false;
  if (!attr_set) {
    DVLP_CUDA_OK(cudaFuncSetAttribute(k_pair_fwd2<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc));
    attr_set = true;
  }
  k_pair_fwd2<R><<<grid, L::kThreads, L::kAlloc, st>>>(mV, mT, mGV, mGBD, a);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}
}  // namespace tf2

int run_pair_fwd2(const void* vb, const void* tb, int Bv, int Bt, int R, int W, int d, float lambda, int focal,
                  float* scores, char* ws, const Workspace& L, bool save, cudaStream_t st) {
  if (d != tf2::kD || W != tf2::kW) return DVLP_ERR_BAD_SHAPE;
  switch (R) {
    case 80: return tf2::launch<80>(vb, tb, Bv, Bt, lambda, focal, scores, ws, L, save, st);
    case 64: return tf2::launch<64>(vb, tb, Bv, Bt, lambda, focal, scores, ws, L, save, st);
    case 48: return tf2::launch<48>(vb, tb, Bv, Bt, lambda, focal, scores, ws, L, save, st);
    case 32: return tf2::launch<32>(vb, tb, Bv, Bt, lambda, focal, scores, ws, L, save, st);
  }
  return DVLP_ERR_BAD_SHAPE;
}

}  // namespace dvlp
