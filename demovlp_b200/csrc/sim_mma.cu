// sim_matrix (model/model.py:582-590) and its backward for bf16 CLS embeddings with d == 256, on mma.sync:
//   out[i][j] = a_i . b_j / (max(|a_i|, eps) max(|b_j|, eps))
//   d_a[i]    = inv_i (g_i - an_i (an_i . g_i)),  g_i = sum_j d_out[i][j] bn_j          (and symmetrically d_b)
// The matrices are small (B x 256, B = 128 .. 1024): a handful of CTAs, raw bf16 operands (exact products,
// fp32 accumulate), the inverse norms applied in fp32 afterwards.  The fp32 d_out weights of the backward are
// split into bf16 hi + lo parts (two MMAs), so the gradient keeps fp32-level accuracy.
#include "common.cuh"
#include "mma_sync.cuh"

namespace dvlp {

constexpr int kSimD = 256;
constexpr int kSimLd = (kSimD + 8) * 2;      // bytes per staged row: conflict-free ldmatrix

// inverse norms of the rows of a (N) and b (M): one warp per row
__global__ void k_sim_inv_norms(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, int N, int M,
                                float eps, float* __restrict__ ia, float* __restrict__ ib) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= N + M) return;
  const __nv_bfloat16* x = row < N ? a + (size_t)row * kSimD : b + (size_t)(row - N) * kSimD;
  const uint4 v = reinterpret_cast<const uint4*>(x)[lane];
  const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
  float ss = 0.f;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float lo = __uint_as_float(w4[h] << 16), hi = __uint_as_float(w4[h] & 0xffff0000u);
    ss = fmaf(lo, lo, ss);
    ss = fmaf(hi, hi, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) (row < N ? ia + row : ib + (row - N))[0] = 1.f / fmaxf(sqrtf(ss), eps);
}

// 32 x 32 output tile per CTA, 4 warps (16 x 16 each); the CTA computes the norms of its own rows
__global__ void __launch_bounds__(128) k_sim_fwd_mma(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                     int N, int M, float eps, float* __restrict__ out) {
  __shared__ __align__(16) uint8_t sa[32 * kSimLd], sb[32 * kSimLd];
  __shared__ float s_ia[32], s_ib[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
    const int row = e >> 5, ch = e & 31;
    uint4 va = make_uint4(0, 0, 0, 0), vb = va;
    if (i0 + row < N) va = reinterpret_cast<const uint4*>(a + (size_t)(i0 + row) * kSimD)[ch];
    if (j0 + row < M) vb = reinterpret_cast<const uint4*>(b + (size_t)(j0 + row) * kSimD)[ch];
    *reinterpret_cast<uint4*>(sa + row * kSimLd + ch * 16) = va;
    *reinterpret_cast<uint4*>(sb + row * kSimLd + ch * 16) = vb;
  }
  __syncthreads();
  for (int rr = warp; rr < 64; rr += 4) {
    const uint8_t* src = (rr < 32 ? sa + rr * kSimLd : sb + (rr - 32) * kSimLd);
    const uint4 v = *reinterpret_cast<const uint4*>(src + lane * 16);
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
    float ss = 0.f;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float lo = __uint_as_float(w4[h] << 16), hi = __uint_as_float(w4[h] & 0xffff0000u);
      ss = fmaf(lo, lo, ss);
      ss = fmaf(hi, hi, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) (rr < 32 ? s_ia + rr : s_ib + (rr - 32))[0] = 1.f / fmaxf(sqrtf(ss), eps);
  }
  __syncthreads();
  const uint32_t ab = (uint32_t)__cvta_generic_to_shared(sa), bb = (uint32_t)__cvta_generic_to_shared(sb);
  const int r0 = (warp & 1) * 16, c0 = (warp >> 1) * 16;
  float acc[2][4] = {};
#pragma unroll 4
  for (int k0 = 0; k0 < kSimD; k0 += 16) {
    uint32_t fa[4], fb[4];
    mm::load_a(ab, kSimLd, r0, k0, lane, fa);
    mm::load_b_nk(bb, kSimLd, c0, k0, lane, fb);
    mm::mma_16816(acc[0], fa[0], fa[1], fa[2], fa[3], fb[0], fb[1]);
    mm::mma_16816(acc[1], fa[0], fa[1], fa[2], fa[3], fb[2], fb[3]);
  }
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = r0 + g + half * 8, c = c0 + h * 8 + tq * 2;
      const int gi = i0 + r, gj = j0 + c;
      if (gi < N) {
        if (gj < M) out[(size_t)gi * M + gj] = acc[h][half * 2] * s_ia[r] * s_ib[c];
        if (gj + 1 < M) out[(size_t)gi * M + gj + 1] = acc[h][half * 2 + 1] * s_ia[r] * s_ib[c + 1];
      }
    }
}

// One CTA per 16 rows of a (side 0) or of b (side 1); 8 warps, warp w owns gradient columns 32w..32w+31.
template <bool OUT_BF16>
__global__ void __launch_bounds__(256) k_sim_bwd_mma(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                     int N, int M, float eps, const float* __restrict__ ia,
                                                     const float* __restrict__ ib, const float* __restrict__ d_out,
                                                     void* __restrict__ d_a, void* __restrict__ d_b, int blocks_a) {
  constexpr int kLdW = (64 + 8) * 2;
  __shared__ __align__(16) uint8_t sy[64 * kSimLd];        // 64 rows of the other matrix
  __shared__ __align__(16) uint8_t sx[16 * kSimLd];        // this CTA's own rows
  __shared__ __align__(16) uint8_t swh[16 * kLdW], swl[16 * kLdW];
  __shared__ float s_dot[8][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool side_b = (int)blockIdx.x >= blocks_a;
  const int i0 = (side_b ? blockIdx.x - blocks_a : blockIdx.x) * 16;
  const __nv_bfloat16* X = side_b ? b : a;
  const __nv_bfloat16* Y = side_b ? a : b;
  const float* ix = side_b ? ib : ia;
  const float* iy = side_b ? ia : ib;
  const int nx = side_b ? M : N, ny = side_b ? N : M;
  void* out = side_b ? d_b : d_a;
  if (!out) return;
  for (int e = threadIdx.x; e < 16 * 32; e += blockDim.x) {
    const int row = e >> 5, ch = e & 31;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (i0 + row < nx) v = reinterpret_cast<const uint4*>(X + (size_t)(i0 + row) * kSimD)[ch];
    *reinterpret_cast<uint4*>(sx + row * kSimLd + ch * 16) = v;
  }
  const uint32_t yb = (uint32_t)__cvta_generic_to_shared(sy);
  const uint32_t whb = (uint32_t)__cvta_generic_to_shared(swh), wlb = (uint32_t)__cvta_generic_to_shared(swl);
  float acc[4][4] = {};
  // The next 64 rows of the other matrix and their weights travel global -> registers while this chunk's MMAs run
  // (the loop used to expose one L2 round trip per chunk: 16 of them at 1024 rows).  256 threads: 8 uint4 + 4 weights each.
  uint4 pre_y[8];
  float pre_w[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int e = threadIdx.x + t * 256, row = e >> 5, ch = e & 31;
      pre_y[t] = make_uint4(0, 0, 0, 0);
      if (k0 + row < ny) pre_y[t] = reinterpret_cast<const uint4*>(Y + (size_t)(k0 + row) * kSimD)[ch];
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = threadIdx.x + t * 256;
      const int r = side_b ? (e & 15) : (e >> 6), c = side_b ? (e >> 4) : (e & 63);
      const int i = i0 + r, j = k0 + c;
      pre_w[t] = 0.f;
      if (i < nx && j < ny) pre_w[t] = (side_b ? d_out[(size_t)j * M + i] : d_out[(size_t)i * M + j]) * iy[j];
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < ny; k0 += 64) {
    __syncthreads();                       // previous chunk consumed
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int e = threadIdx.x + t * 256, row = e >> 5, ch = e & 31;
      *reinterpret_cast<uint4*>(sy + row * kSimLd + ch * 16) = pre_y[t];
    }
    // weights w[r][c] = d_out(own row i0+r, other row k0+c) * inv_other, split into bf16 hi + lo
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = threadIdx.x + t * 256;
      const int r = side_b ? (e & 15) : (e >> 6), c = side_b ? (e >> 4) : (e & 63);
      const float w = pre_w[t];
      const __nv_bfloat16 hi = __float2bfloat16(w);
      const __nv_bfloat16 lo = __float2bfloat16(w - __bfloat162float(hi));
      *reinterpret_cast<__nv_bfloat16*>(swh + r * kLdW + c * 2) = hi;
      *reinterpret_cast<__nv_bfloat16*>(swl + r * kLdW + c * 2) = lo;
    }
    if (k0 + 64 < ny) fetch(k0 + 64);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 64; kk += 16) {
      uint32_t fh[4], fl[4];
      mm::load_a(whb, kLdW, 0, kk, lane, fh);
      mm::load_a(wlb, kLdW, 0, kk, lane, fl);
#pragma unroll
      for (int n2 = 0; n2 < 2; ++n2) {
        uint32_t fb[4];
        mm::load_b_kn(yb, kSimLd, kk, warp * 32 + n2 * 16, lane, fb);
        mm::mma_16816(acc[2 * n2], fh[0], fh[1], fh[2], fh[3], fb[0], fb[1]);
        mm::mma_16816(acc[2 * n2 + 1], fh[0], fh[1], fh[2], fh[3], fb[2], fb[3]);
        mm::mma_16816(acc[2 * n2], fl[0], fl[1], fl[2], fl[3], fb[0], fb[1]);
        mm::mma_16816(acc[2 * n2 + 1], fl[0], fl[1], fl[2], fl[3], fb[2], fb[3]);
      }
    }
  }
  // row dots x_r . g_r (g = acc): this thread holds rows g, g+8 and 8 columns of each
  const int g = lane >> 2, tq = lane & 3;
  float dot_lo = 0.f, dot_hi = 0.f;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const int c = warp * 32 + n * 8 + tq * 2;
    const uint32_t x_lo = *reinterpret_cast<const uint32_t*>(sx + g * kSimLd + c * 2);
    const uint32_t x_hi = *reinterpret_cast<const uint32_t*>(sx + (g + 8) * kSimLd + c * 2);
    dot_lo = fmaf(__uint_as_float(x_lo << 16), acc[n][0], dot_lo);
    dot_lo = fmaf(__uint_as_float(x_lo & 0xffff0000u), acc[n][1], dot_lo);
    dot_hi = fmaf(__uint_as_float(x_hi << 16), acc[n][2], dot_hi);
    dot_hi = fmaf(__uint_as_float(x_hi & 0xffff0000u), acc[n][3], dot_hi);
  }
  dot_lo += __shfl_xor_sync(0xffffffffu, dot_lo, 1); dot_lo += __shfl_xor_sync(0xffffffffu, dot_lo, 2);
  dot_hi += __shfl_xor_sync(0xffffffffu, dot_hi, 1); dot_hi += __shfl_xor_sync(0xffffffffu, dot_hi, 2);
  if (tq == 0) { s_dot[warp][g] = dot_lo; s_dot[warp][g + 8] = dot_hi; }
  __syncthreads();
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = g + half * 8, i = i0 + r;
    if (i >= nx) continue;
    float dot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) dot += s_dot[w][r];
    const float inv = ix[i];
    // |x| < eps: the forward scaled the row by the constant 1/eps (model/model.py:587), no projection term
    const float kq = inv >= 1.f / eps ? 0.f : inv * inv * inv * dot;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int c = warp * 32 + n * 8 + tq * 2;
      const uint32_t xr = *reinterpret_cast<const uint32_t*>(sx + r * kSimLd + c * 2);
      const float o0 = fmaf(-kq, __uint_as_float(xr << 16), inv * acc[n][half * 2]);
      const float o1 = fmaf(-kq, __uint_as_float(xr & 0xffff0000u), inv * acc[n][half * 2 + 1]);
      if (OUT_BF16)
        *reinterpret_cast<__nv_bfloat162*>((__nv_bfloat16*)out + (size_t)i * kSimD + c) = __floats2bfloat162_rn(o0, o1);
      else
        *reinterpret_cast<float2*>((float*)out + (size_t)i * kSimD + c) = make_float2(o0, o1);
    }
  }
}

int run_sim_fwd_mma(const void* a, const void* b, int N, int M, float eps, float* out, cudaStream_t st) {
  dim3 grid((M + 31) / 32, (N + 31) / 32);
  k_sim_fwd_mma<<<grid, 128, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, N, M, eps, out);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

int run_sim_bwd_mma(const void* a, const void* b, int N, int M, float eps, const float* d_out, void* d_a, void* d_b,
                    int out_dtype, float* scratch, cudaStream_t st) {
  float *ia = scratch, *ib = scratch + N;
  k_sim_inv_norms<<<(N + M + 7) / 8, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, N, M, eps, ia, ib);
  DVLP_LAUNCH_OK();
  const int ba = (N + 15) / 16, bb = (M + 15) / 16;
  if (out_dtype == DVLP_BF16)
    k_sim_bwd_mma<true><<<ba + bb, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, N, M, eps, ia, ib, d_out,
                                                  d_a, d_b, ba);
  else
    k_sim_bwd_mma<false><<<ba + bb, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, N, M, eps, ia, ib, d_out,
                                                   d_a, d_b, ba);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

}  // namespace dvlp
