// Gradient contractions of the region-word backward as dense, accumulator-stationary tcgen05 GEMMs
// (single-tile variant, partner of pair_bwd2.cu).
//
// pair_bwd2.cu leaves, per (video i, 4-text tile jt) unit, ONE bf16 tile
//     P = (dS_i2t + dS_t2i) * alpha_r * beta_w      [128 words][R regions], 32-byte-swizzle chunks of 16 columns
// (autograd of model/loss.py:235: S = Vh Th^T  =>  dVh = dS Th, dTh = dS^T Vh).  This file contracts it:
//
//   MODE_V   dV'[i]  (R x d)   = sum_jt P[i, jt]^T T_jt      one CTA per video, K loop over text tiles
//   MODE_T   dT'[jt] (128 x d) = sum_i  P[i, jt]   V_i       one CTA per (text tile, video split)
//   MODE_VT  dV'[i]^T (d x R)  = sum_jt T_jt^T P[i, jt]      MODE_V with the roles of the operands swapped: M = d (two
//            halves of 128 lanes), N = the R regions of BOTH videos of the CTA (their P tiles are adjacent in shared
//            memory, so one MMA covers them), K = 128 words.  MODE_V pads its M = R rows to 128 lanes - 37 % of its
//            tensor work at R = 80, and ncu shows the launch 66 % tensor-active - the transposed form executes exactly
//            the algorithmic products.  The epilogue writes one dV' row element per lane (128-byte runs per warp).
//
// T / V are the RAW bf16 embeddings (the alpha/beta factors are already in P).  The accumulator
// (128 lanes x 256 fp32 columns) stays in TMEM for the whole K loop, operands arrive by TMA into a
// 2-stage ring, one elected thread issues the MMAs; MODE_V writes its rows with plain stores (the CTA
// owns them), MODE_T adds its split's partial sums with red.global (a few MB in total).
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue.
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace dvlp {
using namespace ptx;

namespace gg2 {
constexpr int kD = 256, kRows = 128, kThreads = 192;
constexpr int kTileBytes = 16384, kCh32 = 4096;
enum { MODE_V = 0, MODE_T = 1, MODE_VT = 2 };   // MODE_VT: dV' in the transposed orientation (below), same work split as MODE_V

template <int R, int MODE, int NP>
struct Lay {
  static constexpr int NC = R / 16;                                   // 16-column chunks of the P tile
  static constexpr int CH = R * 128;                                  // one 64-column chunk of an R-row tile
  static constexpr int kStages = NP == 2 ? 2 : (MODE == MODE_T ? 3 : 2);
  static constexpr int PBYTES = NC * kCh32;
  static constexpr int P = 0, X = NP * PBYTES;                        // NP P tiles share one X operand
  static constexpr int XBYTES = MODE == MODE_T ? 4 * CH : 4 * kTileBytes;
  static constexpr int STAGE = NP * PBYTES + XBYTES;
  static constexpr int BARS = kStages * STAGE;
  static constexpr int TMEMPTR = BARS + 10 * 8;
  static constexpr int kAlloc = TMEMPTR + 16;
  static_assert(STAGE % 1024 == 0 && X % 1024 == 0, "128-byte-swizzle tiles need 1 KB alignment");
  static_assert(kAlloc <= 232448, "shared memory budget");
};

struct Args {
  int Bv, Bt, tiles_per_video;
  int n_groups, steps_per_split;  // MODE_V: video groups x text-tile splits;  MODE_T: text-tile groups x video splits
  int accumulate;                 // 1: red.global.add into `out` (zeroed by the caller), 0: plain stores (one split)
  int persistent;                 // 1: gridDim.x CTAs share the (group, K step) sequence evenly (stream-K), always accumulating
  long long text_rows;            // Bt * W: valid rows of dT'
  float* out;                     // MODE_V: dV' (Bv, R, d);  MODE_T: dT' (Bt*W, d)
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// NP = 1: one accumulator per CTA.  NP = 2 (long K loops, i.e. many texts): the CTA keeps TWO accumulators (all of
// TMEM) that share the streamed operand -- per step T_jt once + P[2p,jt], P[2p+1,jt] (MODE_V) or V_i once + P[i,2p],
// P[i,2p+1] (MODE_T): 92 KB instead of 144 KB of L2 -> shared-memory loads per unit over both modes.
template <int R, int MODE, int NP>
__global__ void __launch_bounds__(kThreads, 1)
k_grad_gemm2(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapX, Args a) {
  using L = Lay<R, MODE, NP>;
  constexpr int kStages = L::kStages, NC = L::NC;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if (sbase & 1023u) __trap();
  volatile uint32_t* s_tmem = (volatile uint32_t*)(smem_raw + L::TMEMPTR);
  auto full = [&](int s) { return sbase + L::BARS + 8 * s; };
  auto empty = [&](int s) { return sbase + L::BARS + 8 * (kStages + s); };
  const uint32_t done = sbase + L::BARS + 8 * (2 * kStages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // Work of this CTA: a list of SEGMENTS, each = one group of NP videos (MODE_V) / text tiles (MODE_T) and a K range.
  //   classic:    one segment, (group, split) from blockIdx
  //   persistent: the (group, K step) sequence of the whole launch is cut into gridDim.x equal pieces (stream-K), so
  //               that every SM gets the same number of K steps whatever the group count; a piece may span groups
  constexpr bool kByVideo = MODE != MODE_T;       // groups are videos, the K loop runs over text tiles
  const int k_total = kByVideo ? a.tiles_per_video : a.Bv;
  long long pos0, pos1;
  if (a.persistent) {
    const long long total = (long long)a.n_groups * k_total;
    pos0 = total * blockIdx.x / gridDim.x;
    pos1 = total * (blockIdx.x + 1) / gridDim.x;
  } else {
    const int pr0 = blockIdx.x % a.n_groups, z0 = blockIdx.x / a.n_groups;
    const int kb = z0 * a.steps_per_split;
    const int ns = max(0, min(a.steps_per_split, k_total - kb));
    pos0 = (long long)pr0 * k_total + kb;
    pos1 = pos0 + ns;
  }
  const uint32_t accfree = sbase + L::BARS + 8 * (2 * kStages + 1);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapP); prefetch_tmap(&mapX);
    for (int s = 0; s < 2 * kStages + 1; ++s) mbar_init(sbase + L::BARS + 8 * s, 1);
    mbar_init(accfree, 4);            // the four epilogue warps have read the accumulators of a segment
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(sbase + L::TMEMPTR, 256 * NP);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 0) {
    if (elect_one()) {
      long long kk_glob = 0;                                  // ring position over all segments
      for (long long pos = pos0; pos < pos1;) {
        const int pr = (int)(pos / k_total), k0 = (int)(pos % k_total);
        const int n_steps = (int)min((long long)(k_total - k0), pos1 - pos);
        const int m0 = NP * pr;
        const int n_mem = min(NP, (kByVideo ? a.Bv : a.tiles_per_video) - m0);
        for (int k = 0; k < n_steps; ++k, ++kk_glob) {
          const int s = (int)(kk_glob % kStages);
          const long long it = kk_glob / kStages;
          if (it > 0) mbar_wait(empty(s), (uint32_t)((it - 1) & 1));
          const uint32_t st = sbase + s * L::STAGE;
          const int kk = k0 + k;                            // text tile (MODE_V) / video (MODE_T) of this step
          mbar_arrive_expect_tx(full(s), L::XBYTES + n_mem * L::PBYTES);
          for (int h = 0; h < n_mem; ++h) {
            const long long unit = kByVideo ? (long long)(m0 + h) * a.tiles_per_video + kk
                                            : (long long)kk * a.tiles_per_video + (m0 + h);
            for (int c = 0; c < NC; ++c)
              tma_load_2d(st + L::P + h * L::PBYTES + c * kCh32, &mapP, c * 16, (int)(unit * kRows), full(s));
          }
          if (kByVideo) {
            for (int c = 0; c < 4; ++c) tma_load_2d(st + L::X + c * kTileBytes, &mapX, c * 64, kk * kRows, full(s));
          } else {
            for (int c = 0; c < 4; ++c) tma_load_2d(st + L::X + c * L::CH, &mapX, c * 64, kk * R, full(s));
          }
        }
        pos += n_steps;
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idV = idesc_bf16(128, 256, 1, 1);     // P^T (MN-major view, M = regions) x T (MN-major)
    constexpr uint32_t idT = idesc_bf16(128, 256, 0, 1);     // P (K-major, M = words) x V (MN-major)
    const uint32_t hiK = desc_hi(1024);
    const uint32_t hi32 = ((256u >> 4) & 0x3FFF) | (1u << 14) | (6u << 29);   // 32-byte swizzle, 8-row groups 256 B apart
    long long kk_glob = 0;
    int seg = 0;
    for (long long pos = pos0; pos < pos1; ++seg) {
      const int pr = (int)(pos / k_total), k0 = (int)(pos % k_total);
      const int n_steps = (int)min((long long)(k_total - k0), pos1 - pos);
      const int m0 = NP * pr;
      const int n_mem = min(NP, (kByVideo ? a.Bv : a.tiles_per_video) - m0);
      if (seg > 0) { mbar_wait(accfree, (uint32_t)((seg - 1) & 1)); tc_fence_after(); }   // the epilogue has drained the accumulators
      for (int k = 0; k < n_steps; ++k, ++kk_glob) {
        const int s = (int)(kk_glob % kStages);
        const long long it = kk_glob / kStages;
        mbar_wait(full(s), (uint32_t)(it & 1));
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = sbase + s * L::STAGE;
          if (MODE == MODE_VT) {
            // A = T_jt^T: d is the contiguous (M) direction, 64-column chunks 16 KB apart, one accumulator per half of d;
            // B = the P tiles of the CTA's videos: regions contiguous (N), 16-column chunks 4 KB apart and adjacent
            // across the videos; K = 16 words per instruction
            const uint32_t loX = desc_lo(st + L::X, kTileBytes);
            const uint32_t loP = desc_lo(st + L::P, kCh32);
            const uint32_t id = n_mem == 2 ? idesc_bf16(128, 2 * R, 1, 1) : idesc_bf16(128, R, 1, 1);
#pragma unroll
            for (int half = 0; half < 2; ++half)
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                mma_bf16_w(tmem + half * 256, loX + ((half * 2 * kTileBytes + ks * 2048) >> 4), hiK, loP + ((ks * 512) >> 4),
                           hi32, id, (k | ks) != 0);
          }
          for (int h = 0; MODE != MODE_VT && h < n_mem; ++h) {
            const uint32_t acc = tmem + h * 256;
            const uint32_t pb = st + L::P + h * L::PBYTES;
            if (MODE == MODE_V) {
              // K = 128 words: P rows are the K rows of an MN-major operand whose M (regions) chunks are 4 KB apart;
              // M = 128 reads past the R valid regions (into the next tile): those accumulator lanes are never stored
              const uint32_t loX = desc_lo(st + L::X, kTileBytes);
              const uint32_t loP = desc_lo(pb, kCh32);
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                mma_bf16_w(acc, loP + ((ks * 512) >> 4), hi32, loX + ((ks * 2048) >> 4), hiK, idV, (k | ks) != 0);
            } else {
              // K = R regions: one 16-column chunk of P per K step
              const uint32_t loX = desc_lo(st + L::X, L::CH);
              const uint32_t loP = desc_lo(pb);
#pragma unroll
              for (int ks = 0; ks < NC; ++ks)
                mma_bf16_w(acc, loP + ((ks * kCh32) >> 4), hi32, loX + ((ks * 2048) >> 4), hiK, idT, (k | ks) != 0);
            }
          }
          tc_commit(empty(s));
          if (k == n_steps - 1) tc_commit(done);
        }
        __syncwarp();
      }
      pos += n_steps;
    }
  } else {
    // epilogue: thread = accumulator row (TMEM lane quadrant = warp % 4); one pass per segment
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int seg = 0;
    for (long long pos = pos0; pos < pos1; ++seg) {
      const int pr = (int)(pos / k_total), k0 = (int)(pos % k_total);
      const int n_steps = (int)min((long long)(k_total - k0), pos1 - pos);
      const int m0 = NP * pr;
      const int n_mem = min(NP, (kByVideo ? a.Bv : a.tiles_per_video) - m0);
      pos += n_steps;
      mbar_wait(done, (uint32_t)(seg & 1));
      tc_fence_after();
      if (MODE == MODE_VT) {
        // lane = column of dV' (half * 128 + row), TMEM column = (video of the CTA, region) = consecutive dV' rows
        const int ncols = n_mem * R;
        for (int half = 0; half < 2; ++half) {
          float* dst = a.out + (size_t)m0 * R * kD + half * 128 + row;
#pragma unroll 1
          for (int c = 0; c < ncols; c += 16) {
            float v[16];
            tmem_ld16(tmem + half * 256 + lane_off + c, reinterpret_cast<uint32_t*>(v));
            tmem_ld_wait();
            if (a.accumulate) {
#pragma unroll
              for (int x = 0; x < 16; ++x) atomicAdd(dst + (size_t)(c + x) * kD, v[x]);
            } else {
#pragma unroll
              for (int x = 0; x < 16; ++x) dst[(size_t)(c + x) * kD] = v[x];
            }
          }
        }
      }
      for (int h = 0; MODE != MODE_VT && h < n_mem; ++h) {
        bool valid;
        float* dst;
        if (MODE == MODE_V) {
          valid = row < R;
          dst = a.out + ((size_t)(m0 + h) * R + row) * kD;
        } else {
          const long long grow = (long long)(m0 + h) * kRows + row;
          valid = grow < a.text_rows;
          dst = a.out + (size_t)grow * kD;
        }
        const bool warp_valid = MODE == MODE_V ? (q * 32 < R) : true;
        if (warp_valid) {
#pragma unroll 1
          for (int c = 0; c < kD; c += 32) {
            float v[32];
            tmem_ld16(tmem + h * 256 + lane_off + c, reinterpret_cast<uint32_t*>(v));
            tmem_ld16(tmem + h * 256 + lane_off + c + 16, reinterpret_cast<uint32_t*>(v) + 16);
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
      // the accumulators may be overwritten by the next segment
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accfree);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256 * NP);
}

// ---------------------------------------------------------------------------------------------------------------
// Paired CTAs (two accumulators, long K loops), fine-grained ring.  With whole (text tile / video) steps the 2-stage
// ring of k_grad_gemm2 has one 80 - 104 KB stage in flight while the other is consumed: ncu shows the launches half
// tensor-active, waiting for the fill (HBM / L2 latency + the per-SM fill rate).  Here a ring slot is a K SLICE of a
// step - 32 words of a text tile (MODE_VT: 10 + 16 KB, 8 slots) or 16 regions of a video (MODE_T: 8 + 8 KB, 12
// slots) - so ~180 KB are in flight per SM instead of ~90.  Same work split, operands and results as k_grad_gemm2.
template <int R, int MODE>
struct LayF {
  static_assert(MODE == MODE_VT || MODE == MODE_T, "paired modes only");
  static constexpr int NC = R / 16, NP = 2;
  static constexpr int kSub = MODE == MODE_VT ? 4 : NC;               // slices per step
  static constexpr int PSL = MODE == MODE_VT ? 1024 : kCh32;          // one P chunk slice: 32 rows x 32 B | 128 rows x 32 B
  static constexpr int PBYTES = MODE == MODE_VT ? NC * PSL : PSL;     // per video / text tile and slice
  static constexpr int XCH = MODE == MODE_VT ? 4096 : 2048;           // 64-column chunk of the X slice: 32 | 16 rows x 128 B
  static constexpr int P = 0, X = (NP * PBYTES + 1023) / 1024 * 1024;
  static constexpr int STAGE = X + 4 * XCH;
  static constexpr int kStages = (200 * 1024) / STAGE < 12 ? (200 * 1024) / STAGE : 12;
  static constexpr int BARS = kStages * STAGE;
  static constexpr int TMEMPTR = BARS + (2 * kStages + 2) * 8;
  static constexpr int kAlloc = TMEMPTR + 16;
  static_assert(STAGE % 1024 == 0, "128-byte-swizzle tiles need 1 KB alignment");
  static_assert(kAlloc <= 232448, "shared memory budget");
};

template <int R, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
k_grad_gemm2f(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapX, Args a) {
  using L = LayF<R, MODE>;
  constexpr int kStages = L::kStages, NC = L::NC, NP = 2, kSub = L::kSub;
  constexpr bool kByVideo = MODE == MODE_VT;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if (sbase & 1023u) __trap();
  volatile uint32_t* s_tmem = (volatile uint32_t*)(smem_raw + L::TMEMPTR);
  auto full = [&](int s) { return sbase + L::BARS + 8 * s; };
  auto empty = [&](int s) { return sbase + L::BARS + 8 * (kStages + s); };
  const uint32_t done = sbase + L::BARS + 8 * (2 * kStages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int k_total = kByVideo ? a.tiles_per_video : a.Bv;
  const int pr = blockIdx.x % a.n_groups, z0 = blockIdx.x / a.n_groups;
  const int k0 = z0 * a.steps_per_split;
  const int n_steps = max(0, min(a.steps_per_split, k_total - k0));
  const int m0 = NP * pr;
  const int n_mem = min(NP, (kByVideo ? a.Bv : a.tiles_per_video) - m0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapP); prefetch_tmap(&mapX);
    for (int s = 0; s < 2 * kStages + 1; ++s) mbar_init(sbase + L::BARS + 8 * s, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(sbase + L::TMEMPTR, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 0) {
    if (elect_one()) {
      int cnt = 0;
      for (int k = 0; k < n_steps; ++k) {
        const int kk = k0 + k;                              // text tile (MODE_VT) / video (MODE_T) of this step
        for (int j = 0; j < kSub; ++j, ++cnt) {
          const int s = cnt % kStages, it = cnt / kStages;
          if (it > 0) mbar_wait(empty(s), (uint32_t)((it - 1) & 1));
          const uint32_t st = sbase + s * L::STAGE;
          mbar_arrive_expect_tx(full(s), 4 * L::XCH + n_mem * L::PBYTES);
          for (int h = 0; h < n_mem; ++h) {
            if (kByVideo) {                                   // rows 32 j .. 32 j + 31 of every region chunk
              const long long unit = (long long)(m0 + h) * a.tiles_per_video + kk;
              for (int c = 0; c < NC; ++c)
                tma_load_2d(st + L::P + (h * NC + c) * L::PSL, &mapP, c * 16, (int)(unit * kRows) + j * 32, full(s));
            } else {                                          // region chunk j, all 128 rows
              const long long unit = (long long)kk * a.tiles_per_video + (m0 + h);
              tma_load_2d(st + L::P + h * L::PSL, &mapP, j * 16, (int)(unit * kRows), full(s));
            }
          }
          const int xrow = kByVideo ? kk * kRows + j * 32 : kk * R + j * 16;
          for (int c = 0; c < 4; ++c) tma_load_2d(st + L::X + c * L::XCH, &mapX, c * 64, xrow, full(s));
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idT = idesc_bf16(128, 256, 0, 1);     // P (K-major, M = words) x V (MN-major)
    const uint32_t idVT = n_mem == 2 ? idesc_bf16(128, 2 * R, 1, 1) : idesc_bf16(128, R, 1, 1);
    const uint32_t hiK = desc_hi(1024);
    const uint32_t hi32 = ((256u >> 4) & 0x3FFF) | (1u << 14) | (6u << 29);   // 32-byte swizzle, 8-row groups 256 B apart
    int cnt = 0;
    for (int k = 0; k < n_steps; ++k) {
      for (int j = 0; j < kSub; ++j, ++cnt) {
        const int s = cnt % kStages, it = cnt / kStages;
        mbar_wait(full(s), (uint32_t)(it & 1));
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = sbase + s * L::STAGE;
          const uint32_t loX = desc_lo(st + L::X, L::XCH);
          if (MODE == MODE_VT) {
            // A = 32 words of T_jt^T (M = d: 64-column chunks 4 KB apart, one accumulator per half of d), B = the same
            // words of the P tiles (N = regions: 16-column chunks 1 KB apart, adjacent across the videos)
            const uint32_t loP = desc_lo(st + L::P, L::PSL);
#pragma unroll
            for (int half = 0; half < 2; ++half)
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                mma_bf16_w(tmem + half * 256, loX + ((half * 2 * L::XCH + ks * 2048) >> 4), hiK, loP + ((ks * 512) >> 4), hi32,
                           idVT, (k | j | ks) != 0);
          } else {
            for (int h = 0; h < n_mem; ++h)
              mma_bf16_w(tmem + h * 256, desc_lo(st + L::P + h * L::PSL), hi32, loX, hiK, idT, (k | j) != 0);
          }
          tc_commit(empty(s));
          if (k == n_steps - 1 && j == kSub - 1) tc_commit(done);
        }
        __syncwarp();
      }
    }
  } else if (n_steps > 0) {
    // epilogue (as k_grad_gemm2): thread = accumulator lane (TMEM lane quadrant = warp % 4)
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    mbar_wait(done, 0);
    tc_fence_after();
    if (MODE == MODE_VT) {
      const int ncols = n_mem * R;
      for (int half = 0; half < 2; ++half) {
        float* dst = a.out + (size_t)m0 * R * kD + half * 128 + row;
#pragma unroll 1
        for (int c = 0; c < ncols; c += 16) {
          float v[16];
          tmem_ld16(tmem + half * 256 + lane_off + c, reinterpret_cast<uint32_t*>(v));
          tmem_ld_wait();
          if (a.accumulate) {
#pragma unroll
            for (int x = 0; x < 16; ++x) atomicAdd(dst + (size_t)(c + x) * kD, v[x]);
          } else {
#pragma unroll
            for (int x = 0; x < 16; ++x) dst[(size_t)(c + x) * kD] = v[x];
          }
        }
      }
    } else {
      for (int h = 0; h < n_mem; ++h) {
        const long long grow = (long long)(m0 + h) * kRows + row;
        const bool valid = grow < a.text_rows;
        float* dst = a.out + (size_t)grow * kD;
#pragma unroll 1
        for (int c = 0; c < kD; c += 32) {
          float v[32];
          tmem_ld16(tmem + h * 256 + lane_off + c, reinterpret_cast<uint32_t*>(v));
          tmem_ld16(tmem + h * 256 + lane_off + c + 16, reinterpret_cast<uint32_t*>(v) + 16);
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

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
static int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                    uint32_t box_cols = 64, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return DVLP_ERR_CUDA;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return DVLP_ERR_CUDA; }
  return DVLP_OK;
}

template <int R, int MODE, int NP>
static int launch_mode(const CUtensorMap& mP, const CUtensorMap& mX, Args a, int n_members, int k_total, int sms,
                       bool may_split, size_t out_bytes, cudaStream_t st) {
  using L = Lay<R, MODE, NP>;
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    DVLP_CUDA_OK(cudaFuncSetAttribute(k_grad_gemm2<R, MODE, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc));
    attr_once.done();
  }
  a.n_groups = (n_members + NP - 1) / NP;
  // Stream-K variant (DVLP_GG_PERSIST=1; off by default): one CTA per SM, each takes an equal share of the (group, K step)
  // sequence, so the launch is a single full wave whatever the group count (512 video pairs on 148 SMs are 3.46 waves,
  // 128 text-tile pairs leave 20 SMs idle).  MEASURED on B200 (round 2): 3.19 ms instead of 3.08 at config 3 on one GPU,
  // 82 us instead of 73 at config 2 -- the GEMMs stream the P tiles from HBM (10.8 GB in 3.1 ms), a partial last wave
  // costs nothing because the memory system is the limiter, while stream-K adds a memset and red.global epilogues.
  static const int persist = [] { const char* e = getenv("DVLP_GG_PERSIST"); return e ? atoi(e) : 0; }();
  if (persist) {
    const long long total = (long long)a.n_groups * k_total;
    int grid = (int)(total < sms ? total : sms);
    a.persistent = 1;
    a.steps_per_split = 0;
    if (!a.accumulate) {
      DVLP_CUDA_OK(cudaMemsetAsync(a.out, 0, out_bytes, st));
      a.accumulate = 1;
    }
    k_grad_gemm2<R, MODE, NP><<<grid, kThreads, L::kAlloc, st>>>(mP, mX, a);
    DVLP_LAUNCH_OK();
    return DVLP_OK;
  }
  a.persistent = 0;
  // groups x K splits ~ one wave of CTAs
  int splits = may_split ? sms / a.n_groups : 1;
  if (splits < 1) splits = 1;
  if (splits > k_total) splits = k_total;
  a.steps_per_split = (k_total + splits - 1) / splits;
  splits = (k_total + a.steps_per_split - 1) / a.steps_per_split;
  if (!a.accumulate && splits > 1) {       // split partial sums need a zeroed output
    DVLP_CUDA_OK(cudaMemsetAsync(a.out, 0, out_bytes, st));
    a.accumulate = 1;
  }
  k_grad_gemm2<R, MODE, NP><<<a.n_groups * splits, kThreads, L::kAlloc, st>>>(mP, mX, a);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

// the fine-grained-ring kernel for the paired modes (groups x K splits ~ one wave of CTAs, like launch_mode)
template <int R, int MODE>
static int launch_fine(const CUtensorMap& mP, const CUtensorMap& mX, Args a, int n_members, int k_total, int sms,
                       size_t out_bytes, cudaStream_t st) {
  using L = LayF<R, MODE>;
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    DVLP_CUDA_OK(cudaFuncSetAttribute(k_grad_gemm2f<R, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc));
    attr_once.done();
  }
  a.n_groups = (n_members + 1) / 2;
  a.persistent = 0;
  int splits = sms / a.n_groups;
  if (splits < 1) splits = 1;
  if (splits > k_total) splits = k_total;
  a.steps_per_split = (k_total + splits - 1) / splits;
  splits = (k_total + a.steps_per_split - 1) / a.steps_per_split;
  if (!a.accumulate && splits > 1) {       // split partial sums need a zeroed output
    DVLP_CUDA_OK(cudaMemsetAsync(a.out, 0, out_bytes, st));
    a.accumulate = 1;
  }
  k_grad_gemm2f<R, MODE><<<a.n_groups * splits, kThreads, L::kAlloc, st>>>(mP, mX, a);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

// videos [v0, v0 + Bv): their P tiles sit at the start of the (chunk-sized) P buffer; dV' rows of the range are written,
// dT' accumulates over the range
template <int R>
static int launch(const void* vb_all, const void* tb, int v0, int Bv, int Bt, int W, char* ws, const Workspace& Lw, cudaStream_t st) {
  const int tiles = (Bt + 3) / 4;
  const void* vb = (const char*)vb_all + (size_t)v0 * R * kD * 2;
  const long long units = (long long)Bv * tiles;
  CUtensorMap mP, mT, mV;
  int rc;
  if ((rc = make_map(&mP, ws + Lw.pa, (uint64_t)units * kRows, R, kRows, 16, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  if ((rc = make_map(&mT, tb, (uint64_t)Bt * W, kD, kRows))) return rc;
  if ((rc = make_map(&mV, vb, (uint64_t)Bv * R, kD, R))) return rc;
  int dev = 0, sms = 148;
  DVLP_CUDA_OK(cudaGetDevice(&dev));
  DVLP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  Args a{};
  a.Bv = Bv; a.Bt = Bt; a.tiles_per_video = tiles; a.text_rows = (long long)Bt * W;
  // Two accumulators per CTA share the streamed operand.  With whole-step stages (k_grad_gemm2) that paid only for
  // long K loops (>= 128 text tiles, the sharded global batch); with the fine-grained ring it wins from the per-GPU
  // batch on (measured on B200, both GEMMs: 128 x 128 pairs 72 -> 68 us, 256 x 256 244 -> 173 us, 512 x 512 700 -> 601 us,
  // 1024 x 1024 2.72 -> 2.39 ms).  DVLP_GG_PAIR = 0 / 1 overrides.
  static const int forced = [] { const char* e = getenv("DVLP_GG_PAIR"); return e ? atoi(e) : -1; }();
  const bool pair = forced >= 0 ? forced != 0 : tiles >= 32;
  // dV' (Bv, R, d): overwritten unless split;  dT' (Bt*W, d): accumulated (the caller zeroes it)
  a.out = (float*)(ws + Lw.dvh) + (size_t)v0 * R * kD; a.accumulate = 0;
  const size_t dv_bytes = (size_t)Bv * R * kD * 4;
  // (a chunk of few videos: split the K loop over CTAs so that the launch still fills the GPU)
  // DVLP_GG_FINE = 0: the paired CTAs with whole-step stages (k_grad_gemm2) instead of the fine-grained ring
  static const int fine = [] { const char* e = getenv("DVLP_GG_FINE"); return e ? atoi(e) : 1; }();
  if (pair && fine) {
    CUtensorMap mP32, mT32, mV16;            // K-slice boxes: 32 words of a P chunk / text tile, 16 regions of a video
    if ((rc = make_map(&mP32, ws + Lw.pa, (uint64_t)units * kRows, R, 32, 16, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
    if ((rc = make_map(&mT32, tb, (uint64_t)Bt * W, kD, 32))) return rc;
    if ((rc = make_map(&mV16, vb, (uint64_t)Bv * R, kD, 16))) return rc;
    if ((rc = launch_fine<R, MODE_VT>(mP32, mT32, a, Bv, tiles, sms, dv_bytes, st))) return rc;
    a.out = (float*)(ws + Lw.dth); a.accumulate = 1;
    return launch_fine<R, MODE_T>(mP, mV16, a, tiles, Bv, sms, 0, st);
  }
  rc = pair ? launch_mode<R, MODE_VT, 2>(mP, mT, a, Bv, tiles, sms, true, dv_bytes, st)
            : launch_mode<R, MODE_V, 1>(mP, mT, a, Bv, tiles, sms, 2 * Bv <= sms, dv_bytes, st);
  if (rc) return rc;
  a.out = (float*)(ws + Lw.dth); a.accumulate = 1;
  rc = pair ? launch_mode<R, MODE_T, 2>(mP, mV, a, tiles, Bv, sms, true, 0, st)
            : launch_mode<R, MODE_T, 1>(mP, mV, a, tiles, Bv, sms, true, 0, st);
  return rc;
}
}  // namespace gg2

// dV' (overwritten) and dT' (accumulated: the caller zeroes it) from the P tiles pair_bwd2.cu stored
int run_grad_gemm2(const void* vb, const void* tb, int v0, int nv, int Bt, int R, int W, int d, char* ws, const Workspace& L,
                  cudaStream_t st) {
  if (d != gg2::kD || W != 32) return DVLP_ERR_BAD_SHAPE;
  switch (R) {
    case 80: return gg2::launch<80>(vb, tb, v0, nv, Bt, W, ws, L, st);
    case 64: return gg2::launch<64>(vb, tb, v0, nv, Bt, W, ws, L, st);
    case 48: return gg2::launch<48>(vb, tb, v0, nv, Bt, W, ws, L, st);
    case 32: return gg2::launch<32>(vb, tb, v0, nv, Bt, W, ws, L, st);
  }
  return DVLP_ERR_BAD_SHAPE;
}

}  // namespace dvlp
