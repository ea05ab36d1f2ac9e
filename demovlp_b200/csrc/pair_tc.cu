// Region-word pair scores on tcgen05 / TMEM / TMA (sm_100a): forward kernel.
//
// One persistent CTA per SM walks (video, 4-text tile) units: 128 word rows x R regions per tile
// (R <= 80, W = 32, d = 256).  Per tile the tensor cores produce BOTH orientations of the raw
// similarity block, so that each attention direction of model/loss.py:209-271 is evaluated by a
// warpgroup whose threads own one full softmax row (no cross-lane reductions):
//
//   S1[word m][region r] = T V^T   M=128, N=R,   K=256   -> warpgroup A: i2t direction (softmax over regions)
//   S2[region r][word m] = V T^T   M=128, N=128, K=256   -> warpgroup B: t2i direction (softmax over words)
//   Y1 = U1 G^V  (M=128 words,   N=R,  K=R)   |weighted context|^2 of direction A via the Gram matrix
//   Y2 = U2 G^T  (M=128 regions, N=32, K=32 per text)      ... of direction B
//
// GEMM1 consumes the RAW bf16 embeddings (exact products); the l2 normalisation of model/loss.py:225-226
// is applied afterwards in fp32 as S = raw * alpha_r * beta_w.  The only cross-group traffic is the two
// query-axis l2-norm vectors of model/loss.py:238 (each is a thread-local sum in the OTHER orientation)
// and the final per-text partial means.  Nothing of size Bv*Bt*R*W ever leaves the SM.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2-5 = group A (thread = word row), warps 6-9 = group B (thread = region row).
// TMEM: two 256-column sets (S1|Y1 at +0, S2|Y2 at +128) ping-pong between consecutive tiles so the
// S GEMMs of tile k+1 overlap the softmax work of tile k.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dvlp {

using namespace ptx;

namespace tc {
constexpr int kD = 256, kW = 32, kTPT = 4, kRows = 128, kThreads = 512;
constexpr int kTileBytes = 16384;   // 128 rows x 128 B: one K-chunk (64 bf16) of a 128-row operand tile
constexpr int kGroup = 384;         // threads of group A + both halves of group B (named barriers)

template <int R>
struct Lay {
  static constexpr int kVChunk = R * 128;                  // one K-chunk of the video tile
  static constexpr int V = 0;                              // 4 chunks
  static constexpr int T = V + 4 * kVChunk;                // 4 chunks x 16 KB
  static constexpr int GV = T + 4 * kTileBytes;            // 2 chunks (cols 0-63, 64-127) of G^V
  static constexpr int GT = GV + 2 * kVChunk;              // 2 sets x 16 KB (4 texts x 32 rows x 128 B)
  static constexpr int Q1 = GT + 2 * kTileBytes;           // U1: 128 word rows x (R cols in 2 chunks)
  static constexpr int Q2 = Q1 + 2 * kTileBytes;           // U2: 128 region rows x 128 word cols (2 chunks)
  static constexpr int VEC = Q2 + 2 * kTileBytes;
  // fp32 vectors
  static constexpr int ALPHA = VEC;                        // [128]
  static constexpr int M2V = ALPHA + 512;                  // [128]
  static constexpr int BETA = M2V + 512;                   // [2][128]
  static constexpr int M2T = BETA + 1024;                  // [2][128]
  static constexpr int INVNA = M2T + 1024;                 // [4][R]  1/(|L_r.| + eps) per (text, region)
  static constexpr int INVNB = INVNA + 4 * R * 4;          // [128]   1/(|L_.w| + eps) per word row
  static constexpr int PART = INVNB + 512;                 // [4 warps][4 texts]
  static constexpr int BARS = PART + 64;                   // 16 mbarriers
  static constexpr int TMEMPTR = BARS + 16 * 8;
  static constexpr int kBytes = TMEMPTR + 16;
  static constexpr int kAlloc = kBytes + 1024;             // slack for manual 1024-B alignment
};

enum Bar { V_FULL = 0, T_FULL, T_EMPTY, Q1_FULL, Q2_FULL, S_FULL0, S_FULL1, Y1_FULL0, Y1_FULL1, Y2_FULL0,
           Y2_FULL1, GT_FULL0, GT_FULL1, TM_EMPTY0, TM_EMPTY1, kNumBars };

struct Args {
  int Bv, Bt, tiles_per_video;
  long long total_units;
  const float *alpha, *beta, *m2v, *m2t;
  float lam2;        // lambda * log2(e)
  int focal;
  float* scores;
  float *nrm_a, *nrm_b;   // per unit: [4][R] and [128] attention l2-norm factors, kept for the backward
  float4 *rowv, *colv;    // per unit: per-word / per-(text, region) query scalars for pair_bwd2.cu (see common.cuh)
};

template <int R>
__global__ void __launch_bounds__(kThreads, 1)
k_pair_tc_fwd(const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapT,
              const __grid_constant__ CUtensorMap mapGV, const __grid_constant__ CUtensorMap mapGT, Args a) {
  using L = Lay<R>;
  static_assert(R % 16 == 0 && R >= 16 && R <= 128, "R must be a multiple of 16 in [16,128]");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  float* s_alpha = (float*)(sgen + L::ALPHA);
  float* s_m2v = (float*)(sgen + L::M2V);
  float* s_beta = (float*)(sgen + L::BETA);
  float* s_m2t = (float*)(sgen + L::M2T);
  float* s_invnA = (float*)(sgen + L::INVNA);
  float* s_invnB = (float*)(sgen + L::INVNB);
  float* s_part = (float*)(sgen + L::PART);
  volatile uint32_t* s_tmem = (volatile uint32_t*)(sgen + L::TMEMPTR);
  auto bar = [&](int i) { return sbase + L::BARS + 8 * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long u0 = a.total_units * blockIdx.x / gridDim.x;
  const long long u1 = a.total_units * (blockIdx.x + 1) / gridDim.x;
  const int n_tiles = (int)(u1 - u0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapV); prefetch_tmap(&mapT); prefetch_tmap(&mapGV); prefetch_tmap(&mapGT);
    for (int i = 0; i < kNumBars; ++i) {
      uint32_t cnt = 1;
      if (i == Q1_FULL) cnt = 128;
      if (i == Q2_FULL) cnt = 256;
      if (i == TM_EMPTY0 || i == TM_EMPTY1) cnt = 384;
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
   if (warp == 2) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      for (int k = 0; k < n_tiles; ++k) {
        const int set = k & 1, i = video_of(k), jt = tile_of(k);
        if (is_switch(k)) {
          // the resident video changes: every earlier tile must be finished with V / G^V / alpha
          if (k > 0) mbar_wait(bar(TM_EMPTY0 + ((k - 1) & 1)), ((k - 1) >> 1) & 1);
          mbar_arrive_expect_tx(bar(V_FULL), 6 * L::kVChunk + 2 * R * 4);
          for (int c = 0; c < 4; ++c) tma_load_2d(sbase + L::V + c * L::kVChunk, &mapV, c * 64, i * R, bar(V_FULL));
          for (int c = 0; c < 2; ++c) tma_load_2d(sbase + L::GV + c * L::kVChunk, &mapGV, c * 64, i * R, bar(V_FULL));
          bulk_load(sbase + L::ALPHA, a.alpha + (size_t)i * R, R * 4, bar(V_FULL));
          bulk_load(sbase + L::M2V, a.m2v + (size_t)i * R, R * 4, bar(V_FULL));
        }
        if (k > 0) mbar_wait(bar(T_EMPTY), (k - 1) & 1);          // S GEMMs of tile k-1 have read T
        mbar_arrive_expect_tx(bar(T_FULL), 4 * kTileBytes);
        for (int c = 0; c < 4; ++c) tma_load_2d(sbase + L::T + c * kTileBytes, &mapT, c * 64, jt * kRows, bar(T_FULL));
        if (k >= 2) mbar_wait(bar(TM_EMPTY0 + set), ((k - 2) >> 1) & 1);   // tile k-2 done with this set
        mbar_arrive_expect_tx(bar(GT_FULL0 + set), kTileBytes + 1024);
        tma_load_2d(sbase + L::GT + set * kTileBytes, &mapGT, 0, jt * kRows, bar(GT_FULL0 + set));
        bulk_load(sbase + L::BETA + set * 512, a.beta + (size_t)jt * kRows, 512, bar(GT_FULL0 + set));
        bulk_load(sbase + L::M2T + set * 512, a.m2t + (size_t)jt * kRows, 512, bar(GT_FULL0 + set));
      }
    }
  } else if (warp == 3) {
    // ================================ MMA issuer ================================
    // (warps 2/3: SMSPs 2 and 3 carry the least group work, so the single issuing thread is not starved;
    //  descriptor words are loop invariant, a K step adds a constant to the low word)
    constexpr uint32_t idS1 = idesc_bf16(128, R), idS2 = idesc_bf16(128, 128), idY2 = idesc_bf16(128, 32);
    const uint32_t hiK = desc_hi(1024);
    const uint32_t loT = desc_lo(sbase + L::T), loV = desc_lo(sbase + L::V), loGV = desc_lo(sbase + L::GV);
    const uint32_t loQ1 = desc_lo(sbase + L::Q1), loQ2 = desc_lo(sbase + L::Q2);
    int vloads = 0;
    auto issue_S = [&](int k) {
      const int set = k & 1;
      if (is_switch(k)) { mbar_wait(bar(V_FULL), vloads & 1); ++vloads; }
      mbar_wait(bar(T_FULL), k & 1);
      if (k >= 2) mbar_wait(bar(TM_EMPTY0 + set), ((k - 2) >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t dS1 = tmem + set * 256, dS2 = tmem + set * 256 + 128;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          const uint32_t oT = ((ks >> 2) * kTileBytes + (ks & 3) * 32) >> 4, oV = ((ks >> 2) * L::kVChunk + (ks & 3) * 32) >> 4;
          mma_bf16_w(dS1, loT + oT, hiK, loV + oV, hiK, idS1, ks != 0);
        }
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          const uint32_t oT = ((ks >> 2) * kTileBytes + (ks & 3) * 32) >> 4, oV = ((ks >> 2) * L::kVChunk + (ks & 3) * 32) >> 4;
          mma_bf16_w(dS2, loV + oV, hiK, loT + oT, hiK, idS2, ks != 0);
        }
        tc_commit(bar(S_FULL0 + set));
        tc_commit(bar(T_EMPTY));
      }
      __syncwarp();
    };
    if (n_tiles > 0) issue_S(0);
    for (int k = 0; k < n_tiles; ++k) {
      const int set = k & 1;
      const bool more = k + 1 < n_tiles;
      const bool ahead = more && !is_switch(k + 1);
      if (ahead) issue_S(k + 1);
      // Y1 = U1 G^V
      mbar_wait(bar(Q1_FULL), k & 1);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int s = 0; s < R / 16; ++s)
          mma_bf16_w(tmem + set * 256, loQ1 + (((s / 4) * kTileBytes + (s % 4) * 32) >> 4), hiK,
                     loGV + (((s / 4) * L::kVChunk + (s % 4) * 32) >> 4), hiK, idS1, s != 0);
        tc_commit(bar(Y1_FULL0 + set));
      }
      __syncwarp();
      // Y2 = U2 G^T, block-diagonal over the 4 texts
      mbar_wait(bar(Q2_FULL), k & 1);
      mbar_wait(bar(GT_FULL0 + set), (k >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t loGT = desc_lo(sbase + L::GT + set * kTileBytes);
#pragma unroll
        for (int t = 0; t < kTPT; ++t)
#pragma unroll
          for (int s = 0; s < 2; ++s)
            mma_bf16_w(tmem + set * 256 + 128 + t * 32, loQ2 + (((t / 2) * kTileBytes + (t % 2) * 64 + s * 32) >> 4), hiK,
                       loGT + ((t * 4096 + s * 32) >> 4), hiK, idY2, s != 0);
        tc_commit(bar(Y2_FULL0 + set));
      }
      __syncwarp();
      if (more && !ahead) issue_S(k + 1);
    }
   }
  } else if (warp < 8) {
    // ================== group A: i2t direction, thread = word row ==================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int q = warp & 3;                 // TMEM lane quadrant of this warp == text slot of the tile
    const int m = q * 32 + lane;            // word row of the tile
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t q1_row = sbase + L::Q1 + (m >> 3) * 1024 + (m & 7) * 128;
    int vloads = 0;
    for (int k = 0; k < n_tiles; ++k) {
      const int set = k & 1, i = video_of(k), j = tile_of(k) * kTPT + q;
      if (is_switch(k)) { mbar_wait(bar(V_FULL), vloads & 1); ++vloads; }
      mbar_wait(bar(GT_FULL0 + set), (k >> 1) & 1);
      mbar_wait(bar(S_FULL0 + set), (k >> 1) & 1);
      tc_fence_after();
      const float beta_m = s_beta[set * 128 + m];
      float S[R], e[R];
      {
        uint32_t* raw = reinterpret_cast<uint32_t*>(S);
#pragma unroll
        for (int c = 0; c < R / 16; ++c) tmem_ld16(tmem + set * 256 + lane_off + c * 16, raw + c * 16);
        tmem_ld_wait();
      }
      // pass 1: S = raw * alpha_r * beta_w ; |LeakyReLU(S)|^2 over regions = this word's dir-B norm
      // (4 independent accumulators: the chains are latency-, not throughput-bound at 2 warps/scheduler)
      float acc4[4] = {0.f, 0.f, 0.f, 0.f};
      float inb;
#pragma unroll
      for (int r = 0; r < R; r += 4) {
        const float4 al = *reinterpret_cast<const float4*>(s_alpha + r);
        S[r + 0] *= beta_m * al.x; S[r + 1] *= beta_m * al.y; S[r + 2] *= beta_m * al.z; S[r + 3] *= beta_m * al.w;
#pragma unroll
        for (int h = 0; h < 4; ++h) { float l = leaky(S[r + h]); acc4[h] = fmaf(l, l, acc4[h]); }
      }
      {
        inb = 1.f / (sqrtf((acc4[0] + acc4[1]) + (acc4[2] + acc4[3])) + kEps);
        s_invnB[m] = inb;
        a.nrm_b[(size_t)(u0 + k) * kRows + m] = inb;
      }
      named_bar_sync(1, kGroup);             // norms exchanged between the groups
      // pass 2: A = L / n_r ; e = exp(lambda (A + mask_r) - lambda) ; focal ; numerator
      acc4[0] = acc4[1] = acc4[2] = acc4[3] = 0.f;
#pragma unroll
      for (int r = 0; r < R; r += 4) {
        const float4 na = *reinterpret_cast<const float4*>(s_invnA + q * R + r);
        const float4 mk = *reinterpret_cast<const float4*>(s_m2v + r);
        e[r + 0] = ex2(fmaf(leaky(S[r + 0]) * na.x, a.lam2, mk.x));
        e[r + 1] = ex2(fmaf(leaky(S[r + 1]) * na.y, a.lam2, mk.y));
        e[r + 2] = ex2(fmaf(leaky(S[r + 2]) * na.z, a.lam2, mk.z));
        e[r + 3] = ex2(fmaf(leaky(S[r + 3]) * na.w, a.lam2, mk.w));
#pragma unroll
        for (int h = 0; h < 4; ++h) acc4[h] += e[r + h];
      }
      // focal "equal": keep e_r iff e_r * R > sum_r e_r  (model/loss.py:280-282); all kept otherwise
      const float thr = a.focal ? ((acc4[0] + acc4[1]) + (acc4[2] + acc4[3])) * (1.f / R) : -1.f;
      acc4[0] = acc4[1] = acc4[2] = acc4[3] = 0.f;
#pragma unroll
      for (int g = 0; g < R / 8; ++g) {
        uint32_t pk[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int r = g * 8 + h * 2;
          pk[h] = pack_bf16(e[r] > thr ? e[r] : 0.f, e[r + 1] > thr ? e[r + 1] : 0.f);
          e[r] = bf16_lo(pk[h]);
          e[r + 1] = bf16_hi(pk[h]);
          acc4[h] = fmaf(e[r], S[r], acc4[h]);
          acc4[h] = fmaf(e[r + 1], S[r + 1], acc4[h]);
        }
        st_shared_v4(q1_row + (g >> 3) * kTileBytes + (((g & 7) ^ (m & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
      }
      const float num = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
      fence_proxy_async();
      mbar_arrive(bar(Q1_FULL));
      // denominator: D = u^T G^V u = sum_r u_r Y1_r
      mbar_wait(bar(Y1_FULL0 + set), (k >> 1) & 1);
      tc_fence_after();
      {
        uint32_t* raw = reinterpret_cast<uint32_t*>(S);
#pragma unroll
        for (int c = 0; c < R / 16; ++c) tmem_ld16(tmem + set * 256 + lane_off + c * 16, raw + c * 16);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(bar(TM_EMPTY0 + set));      // this thread is done with TMEM set `set`
      float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int r = 0; r < R; r += 4)
#pragma unroll
        for (int h = 0; h < 4; ++h) d4[h] = fmaf(e[r + h], S[r + h], d4[h]);
      const float Dn = (d4[0] + d4[1]) + (d4[2] + d4[3]);
      const float rden = Dn > 1e-30f ? rsqrtf(Dn) : 0.f;
      float cs = num * rden;
      a.rowv[(size_t)(u0 + k) * kRows + m] = make_float4(inb, thr, rden, cs);
      cs = warp_sum(cs);                      // mean over the 32 words of text q (model/loss.py:318)
      named_bar_sync(2, kGroup);              // group B's per-text partial sums are in s_part
      if (lane == 0 && j < a.Bt) {
        float t2i = (s_part[q] + s_part[4 + q]) + (s_part[8 + q] + s_part[12 + q]);
        a.scores[(size_t)i * a.Bt + j] = cs * (1.f / kW) + t2i * (1.f / R);
      }
    }
  } else {
    // ================= group B: t2i direction, thread = region row =================
    // two halves (warps 8-11: texts 0,1 of the tile; warps 12-15: texts 2,3) double the resident warps
    asm volatile("setmaxnreg.dec.sync.aligned.u32 120;");
    const int t_lo = (warp >= 12) ? 2 : 0;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const bool warp_active = q * 32 < R;      // warps whose 32 rows are all padding only keep the protocol going
    const bool active = r < R;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t q2_row = sbase + L::Q2 + (r >> 3) * 1024 + (r & 7) * 128;
    int vloads = 0;
    for (int k = 0; k < n_tiles; ++k) {
      const int set = k & 1;
      if (is_switch(k)) { mbar_wait(bar(V_FULL), vloads & 1); ++vloads; }
      mbar_wait(bar(GT_FULL0 + set), (k >> 1) & 1);
      mbar_wait(bar(S_FULL0 + set), (k >> 1) & 1);
      tc_fence_after();
      const float alpha_r = active ? s_alpha[r] : 0.f;
      const float* beta = s_beta + set * 128;
      const float* m2t = s_m2t + set * 128;
      const uint32_t tS2 = tmem + set * 256 + 128 + lane_off;
      float num[2], inav[2], thrv[2];
      if (warp_active) {
        // pass 1: per text, |LeakyReLU(S)|^2 over its 32 words = this region's dir-A norm
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          const int t = t_lo + tt;
          float x[32];
          tmem_ld16(tS2 + t * 32, reinterpret_cast<uint32_t*>(x));
          tmem_ld16(tS2 + t * 32 + 16, reinterpret_cast<uint32_t*>(x) + 16);
          tmem_ld_wait();
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int w = 0; w < 32; w += 4) {
            const float4 be = *reinterpret_cast<const float4*>(beta + t * 32 + w);
            float l0 = leaky(x[w] * (alpha_r * be.x)), l1 = leaky(x[w + 1] * (alpha_r * be.y));
            float l2 = leaky(x[w + 2] * (alpha_r * be.z)), l3 = leaky(x[w + 3] * (alpha_r * be.w));
            s4[0] = fmaf(l0, l0, s4[0]); s4[1] = fmaf(l1, l1, s4[1]);
            s4[2] = fmaf(l2, l2, s4[2]); s4[3] = fmaf(l3, l3, s4[3]);
          }
          const float ss = (s4[0] + s4[1]) + (s4[2] + s4[3]);
          const float ina = 1.f / (sqrtf(ss) + kEps);
          inav[tt] = ina;
          if (active) {
            s_invnA[t * R + r] = ina;
            a.nrm_a[((size_t)(u0 + k) * kTPT + t) * R + r] = ina;
          }
        }
      }
      named_bar_sync(1, kGroup);
      if (warp_active) {
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          const int t = t_lo + tt;
          float x[32], e[32];
          tmem_ld16(tS2 + t * 32, reinterpret_cast<uint32_t*>(x));
          tmem_ld16(tS2 + t * 32 + 16, reinterpret_cast<uint32_t*>(x) + 16);
          tmem_ld_wait();
          float e4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int w = 0; w < 32; w += 4) {
            const float4 be = *reinterpret_cast<const float4*>(beta + t * 32 + w);
            const float4 nb = *reinterpret_cast<const float4*>(s_invnB + t * 32 + w);
            const float4 mk = *reinterpret_cast<const float4*>(m2t + t * 32 + w);
            x[w + 0] *= alpha_r * be.x; x[w + 1] *= alpha_r * be.y; x[w + 2] *= alpha_r * be.z; x[w + 3] *= alpha_r * be.w;
            e[w + 0] = ex2(fmaf(leaky(x[w + 0]) * nb.x, a.lam2, mk.x));
            e[w + 1] = ex2(fmaf(leaky(x[w + 1]) * nb.y, a.lam2, mk.y));
            e[w + 2] = ex2(fmaf(leaky(x[w + 2]) * nb.z, a.lam2, mk.z));
            e[w + 3] = ex2(fmaf(leaky(x[w + 3]) * nb.w, a.lam2, mk.w));
#pragma unroll
            for (int h = 0; h < 4; ++h) e4[h] += e[w + h];
          }
          const float thr = a.focal ? ((e4[0] + e4[1]) + (e4[2] + e4[3])) * (1.f / kW) : -1.f;
          thrv[tt] = thr;
          float n4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t pk[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int w = g * 8 + h * 2;
              pk[h] = pack_bf16(e[w] > thr ? e[w] : 0.f, e[w + 1] > thr ? e[w + 1] : 0.f);
              n4[h] = fmaf(bf16_lo(pk[h]), x[w], n4[h]);
              n4[h] = fmaf(bf16_hi(pk[h]), x[w + 1], n4[h]);
            }
            const int seg = (t & 1) * 4 + g;
            st_shared_v4(q2_row + (t >> 1) * kTileBytes + ((seg ^ (r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
          }
          const float nm = (n4[0] + n4[1]) + (n4[2] + n4[3]);
          num[tt] = nm;
        }
      }
      fence_proxy_async();
      mbar_arrive(bar(Q2_FULL));
      mbar_wait(bar(Y2_FULL0 + set), (k >> 1) & 1);
      tc_fence_after();
      float cs[2] = {0.f, 0.f};
      if (warp_active) {
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          const int t = t_lo + tt;
          float y[32];
          tmem_ld16(tS2 + t * 32, reinterpret_cast<uint32_t*>(y));
          tmem_ld16(tS2 + t * 32 + 16, reinterpret_cast<uint32_t*>(y) + 16);
          tmem_ld_wait();
          float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t pk[4];
            const int seg = (t & 1) * 4 + g;
            ld_shared_v4(q2_row + (t >> 1) * kTileBytes + ((seg ^ (r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              d4[h] = fmaf(bf16_lo(pk[h]), y[g * 8 + h * 2], d4[h]);
              d4[h] = fmaf(bf16_hi(pk[h]), y[g * 8 + h * 2 + 1], d4[h]);
            }
          }
          const float Dn = (d4[0] + d4[1]) + (d4[2] + d4[3]);
          const float rden = (active && Dn > 1e-30f) ? rsqrtf(Dn) : 0.f;
          cs[tt] = active ? num[tt] * rden : 0.f;
          if (active) a.colv[((size_t)(u0 + k) * kTPT + t) * R + r] = make_float4(inav[tt], thrv[tt], rden, cs[tt]);
        }
      }
      tc_fence_before();
      mbar_arrive(bar(TM_EMPTY0 + set));
#pragma unroll
      for (int tt = 0; tt < 2; ++tt) {
        float v = warp_sum(cs[tt]);
        if (lane == 0) s_part[q * 4 + t_lo + tt] = v;
      }
      named_bar_sync(2, kGroup);
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

// bf16 matrix (rows x cols, row-major) -> boxes of 64 columns x box_rows rows, 128-byte swizzle
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
static int launch_fwd(const void* vb, const void* tb, int Bv, int Bt, float lambda, int focal, float* scores,
                      char* ws, const Workspace& Lw, cudaStream_t st) {
  CUtensorMap mV, mT, mGV, mGT;
  int rc;
  if ((rc = make_map(&mV, vb, (uint64_t)Bv * R, kD, R))) return rc;
  if ((rc = make_map(&mT, tb, (uint64_t)Bt * kW, kD, kRows))) return rc;
  if ((rc = make_map(&mGV, ws + Lw.gvb, (uint64_t)Bv * R, R, R))) return rc;
  if ((rc = make_map(&mGT, ws + Lw.gtb, (uint64_t)Bt * kW, kW, kRows))) return rc;
  Args a{};
  a.Bv = Bv; a.Bt = Bt;
  a.tiles_per_video = (Bt + kTPT - 1) / kTPT;
  a.total_units = (long long)Bv * a.tiles_per_video;
  a.alpha = (const float*)(ws + Lw.alpha); a.beta = (const float*)(ws + Lw.beta);
  a.m2v = (const float*)(ws + Lw.m2v); a.m2t = (const float*)(ws + Lw.m2t);
  a.lam2 = lambda * kLog2e; a.focal = focal; a.scores = scores;
  a.nrm_a = (float*)(ws + Lw.nrm_a); a.nrm_b = (float*)(ws + Lw.nrm_b);
  a.rowv = (float4*)(ws + Lw.rowv); a.colv = (float4*)(ws + Lw.colv);
  int dev = 0, sms = 148;
  DVLP_CUDA_OK(cudaGetDevice(&dev));
  DVLP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = (int)(a.total_units < sms ? a.total_units : sms);
  static bool attr_set = false;
  if (!attr_set) {
    DVLP_CUDA_OK(cudaFuncSetAttribute(k_pair_tc_fwd<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<R>::kAlloc));
    attr_set = true;
  }
  k_pair_tc_fwd<R><<<grid, kThreads, Lay<R>::kAlloc, st>>>(mV, mT, mGV, mGT, a);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}
}  // namespace tc

bool tc_supported(int R, int W, int d) { return d == tc::kD && W == tc::kW && (R == 80 || R == 64 || R == 48 || R == 32); }

int run_pair_tc_fwd(const void* vb, const void* tb, int Bv, int Bt, int R, int W, int d, float lambda, int focal,
                    float* scores, char* ws, const Workspace& L, cudaStream_t st) {
  if (!tc_supported(R, W, d)) return DVLP_ERR_BAD_SHAPE;
  switch (R) {
    case 80: return tc::launch_fwd<80>(vb, tb, Bv, Bt, lambda, focal, scores, ws, L, st);
    case 64: return tc::launch_fwd<64>(vb, tb, Bv, Bt, lambda, focal, scores, ws, L, st);
    case 48: return tc::launch_fwd<48>(vb, tb, Bv, Bt, lambda, focal, scores, ws, L, st);
    case 32: return tc::launch_fwd<32>(vb, tb, Bv, Bt, lambda, focal, scores, ws, L, st);
  }
  return DVLP_ERR_BAD_SHAPE;
}
}  // namespace dvlp
