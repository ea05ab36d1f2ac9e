// Gradient finalisation: from gradients w.r.t. the l2-normalised rows (plus the Gram-matrix
// gradient) to gradients w.r.t. the raw inputs -- the backward of model/loss.py:333-338
// (l2norm) composed with G = Xh Xh^T.
//
//   g_r   = dXh_r [/ inv_r when the GEMM left an extra inv_r factor] + 2 sum_c dG[r][c] Xh_c
//   d_x_r = inv_r ( g_r - Xh_r (Xh_r . g_r) )            inv_r = 1/(|x_r| + eps)
#include "common.cuh"
#include "mma_sync.cuh"

namespace dvlp {

template <typename Tin>
__global__ void __launch_bounds__(256) k_grad_finalize(const Tin* __restrict__ x, const float* __restrict__ inv,
                                                       const float* __restrict__ dxh, const float* __restrict__ dG,
                                                       int L, int d, int prescaled, float* __restrict__ d_x) {
  __shared__ float red[8];
  __shared__ float s_dot;
  const int row = blockIdx.x;            // item * L + r
  const int item = row / L, r = row - item * L;
  const Tin* X = x + (size_t)item * L * d;
  const float* iv = inv + (size_t)item * L;
  const float* g_row = dG + ((size_t)item * L + r) * L;
  const float a = iv[r];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float dot_part = 0.f;
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    float acc = dxh[(size_t)row * d + k];
    if (prescaled) acc /= a;
    float gsum = 0.f;
    for (int c = 0; c < L; ++c) gsum = fmaf(g_row[c] * iv[c], to_f32(X[(size_t)c * d + k]), gsum);
    acc = fmaf(2.f, gsum, acc);
    d_x[(size_t)row * d + k] = acc;      // staged
    dot_part = fmaf(to_f32(X[(size_t)r * d + k]) * a, acc, dot_part);
  }
  dot_part = warp_sum(dot_part);
  if (lane == 0) red[warp] = dot_part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    s_dot = t;
  }
  __syncthreads();
  const float dot = s_dot;
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    float acc = d_x[(size_t)row * d + k];
    float xh = to_f32(X[(size_t)r * d + k]) * a;
    d_x[(size_t)row * d + k] = a * (acc - xh * dot);
  }
}

// Tiled version (d <= 256): one CTA per item, thread = embedding column, 16 output rows per pass.
// The 16 x L slice of dG sits in shared memory (transposed: float4 broadcast reads); every row of X
// is streamed once per pass with coalesced loads.
template <typename Tin>
__global__ void __launch_bounds__(256) k_grad_finalize_tile(const Tin* __restrict__ x, const float* __restrict__ inv,
                                                            const float* __restrict__ dxh, const float* __restrict__ dG,
                                                            int L, int d, int prescaled, float* __restrict__ d_x) {
  constexpr int RT = 16;
  extern __shared__ float sm[];
  float* gs = sm;                    // [L][RT]   gs[c*RT + rr] = dG[r0+rr][c] * inv[c]
  float* red = sm + (size_t)L * RT;  // [RT][8]
  float* dots = red + RT * 8;        // [RT]
  const int item = blockIdx.x;
  const Tin* X = x + (size_t)item * L * d;
  const float* iv = inv + (size_t)item * L;
  const float* Gi = dG + (size_t)item * L * L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, k = threadIdx.x;
  const bool on = k < d;
  for (int r0 = 0; r0 < L; r0 += RT) {
    __syncthreads();
    for (int e = threadIdx.x; e < L * RT; e += blockDim.x) {
      const int rr = e / L, c = e - rr * L;
      gs[c * RT + rr] = (r0 + rr < L) ? Gi[(size_t)(r0 + rr) * L + c] * iv[c] : 0.f;
    }
    __syncthreads();
    float acc[RT], xh[RT];
#pragma unroll
    for (int rr = 0; rr < RT; ++rr) acc[rr] = 0.f;
    if (on) {
      for (int c = 0; c < L; ++c) {
        const float xc = to_f32(X[(size_t)c * d + k]);
        const float4* g4 = reinterpret_cast<const float4*>(gs + c * RT);
#pragma unroll
        for (int v = 0; v < RT / 4; ++v) {
          const float4 g = g4[v];
          acc[v * 4 + 0] = fmaf(g.x, xc, acc[v * 4 + 0]); acc[v * 4 + 1] = fmaf(g.y, xc, acc[v * 4 + 1]);
          acc[v * 4 + 2] = fmaf(g.z, xc, acc[v * 4 + 2]); acc[v * 4 + 3] = fmaf(g.w, xc, acc[v * 4 + 3]);
        }
      }
    }
    // g = dXh [/ inv] + 2 dG Xh ;  dot(Xh_r, g_r) over the row for RT rows at once
#pragma unroll
    for (int rr = 0; rr < RT; ++rr) {
      const int r = r0 + rr;
      float p = 0.f;
      xh[rr] = 0.f;
      if (r < L && on) {
        const float a = iv[r];
        float g = dxh[((size_t)item * L + r) * d + k];
        if (prescaled) g /= a;
        acc[rr] = fmaf(2.f, acc[rr], g);
        xh[rr] = to_f32(X[(size_t)r * d + k]) * a;
        p = xh[rr] * acc[rr];
      }
      p = warp_sum(p);
      if (lane == 0) red[rr * 8 + warp] = p;
    }
    __syncthreads();
    if (threadIdx.x < RT) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += red[threadIdx.x * 8 + w];
      dots[threadIdx.x] = s;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < RT; ++rr) {
      const int r = r0 + rr;
      if (r < L && on) d_x[((size_t)item * L + r) * d + k] = iv[r] * (acc[rr] - xh[rr] * dots[rr]);
    }
  }
}

// tcgen05 path (bf16 rows, d == 256, L % 16 == 0, L <= 128, dX' prescaled by inv_r): one CTA per video / text.
//   D_r   = dX'_r + sum_c (2 inv_r inv_c dG[r][c]) x_c         (= inv_r g_r;  mma.sync, bf16 operands, fp32 acc)
//   d_x_r = D_r - inv_r^2 (x_r . D_r) x_r                       (the l2norm backward, model/loss.py:333-338)
// Warp w owns rows 16w..16w+15 and all 256 columns (128 accumulator registers per thread).
constexpr int kFinLdX = (256 + 8) * 2;      // bytes
template <bool OUT_BF16, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT <= 320 ? 2 : 1) k_finalize_item(const __nv_bfloat16* __restrict__ v, const __nv_bfloat16* __restrict__ t,
                                                       const float* __restrict__ alpha, const float* __restrict__ beta,
                                                       const float* __restrict__ dvh, const float* __restrict__ dth,
                                                       const float* __restrict__ dgv, const float* __restrict__ dgt,
                                                       int Bv, int R, int W, int dgt_swz, void* __restrict__ d_im,
                                                       void* __restrict__ d_s) {
  extern __shared__ __align__(16) uint8_t fsm[];
  __shared__ float s_inv[128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool text = (int)blockIdx.x >= Bv;
  const int item = text ? blockIdx.x - Bv : blockIdx.x;
  const int L = text ? W : R;
  const __nv_bfloat16* X = (text ? t : v) + (size_t)item * L * 256;
  const float* inv = (text ? beta : alpha) + (size_t)item * L;
  const float* dX = (text ? dth : dvh) + (size_t)item * L * 256;
  const float* dG = (text ? dgt : dgv) + (size_t)item * L * L;
  void* out_base = text ? d_s : d_im;
  const size_t out_off = (size_t)item * L * 256;
  const int ldA = (L + 8) * 2;
  uint8_t* xs = fsm;                         // [L][264] bf16 raw rows
  uint8_t* as = fsm + (size_t)L * kFinLdX;   // [L][L+8] bf16: 2 inv_r inv_c dG[r][c]
  const uint32_t xb = (uint32_t)__cvta_generic_to_shared(xs), ab = (uint32_t)__cvta_generic_to_shared(as);
  for (int e = threadIdx.x; e < L; e += blockDim.x) s_inv[e] = inv[e];
  for (int e = threadIdx.x; e < L * 32; e += blockDim.x) {
    const int row = e >> 5, ch = e & 31;
    *reinterpret_cast<uint4*>(xs + row * kFinLdX + ch * 16) = reinterpret_cast<const uint4*>(X)[e];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < L * L / 2; e += blockDim.x) {
    const int r = (2 * e) / L, c = 2 * e - r * L;
    // pair_bwd2.cu leaves dG^T rows (W == 32) with their 16-byte chunks XOR-swizzled by the row index
    const int cs = (text && dgt_swz) ? ((((c >> 2) ^ (r & 7)) << 2) | (c & 3)) : c;
    const float2 g = *reinterpret_cast<const float2*>(dG + (size_t)r * L + cs);
    const float sr = 2.f * s_inv[r];
    *reinterpret_cast<__nv_bfloat162*>(as + r * ldA + c * 2) = __floats2bfloat162_rn(g.x * sr * s_inv[c], g.y * sr * s_inv[c + 1]);
  }
  __syncthreads();
  // warp = (16-row block, 128-column half): 64 accumulator registers per thread, so that two CTAs fit an SM and all
  // items of a config-2 step are resident at once.  The launch provides one warp per task (2 L / 16 warps).
  __shared__ float s_dot[2][128];
  const int nrb = L >> 4;
  const int g = lane >> 2, tq = lane & 3;
  float acc[16][4];
  float dot_lo = 0.f, dot_hi = 0.f;
  const bool active = warp < 2 * nrb;
  const int row0 = (warp >> 1) * 16, ch = warp & 1, col0 = ch * 128;
  if (active) {
#pragma unroll
    for (int n = 0; n < 16; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
    for (int k0 = 0; k0 < L; k0 += 16) {
      uint32_t a[4];
      mm::load_a(ab, ldA, row0, k0, lane, a);
#pragma unroll
      for (int n2 = 0; n2 < 8; ++n2) {
        uint32_t b[4];
        mm::load_b_kn(xb, kFinLdX, k0, col0 + n2 * 16, lane, b);
        mm::mma_16816(acc[2 * n2], a[0], a[1], a[2], a[3], b[0], b[1]);
        mm::mma_16816(acc[2 * n2 + 1], a[0], a[1], a[2], a[3], b[2], b[3]);
      }
    }
    // add the prescaled S-path gradient; partial row dots with the raw row
    const int r_lo = row0 + g, r_hi = row0 + g + 8;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int c = col0 + n * 8 + tq * 2;
      const float2 d_lo = *reinterpret_cast<const float2*>(dX + (size_t)r_lo * 256 + c);
      const float2 d_hi = *reinterpret_cast<const float2*>(dX + (size_t)r_hi * 256 + c);
      acc[n][0] += d_lo.x; acc[n][1] += d_lo.y; acc[n][2] += d_hi.x; acc[n][3] += d_hi.y;
      const uint32_t x_lo = *reinterpret_cast<const uint32_t*>(xs + r_lo * kFinLdX + c * 2);
      const uint32_t x_hi = *reinterpret_cast<const uint32_t*>(xs + r_hi * kFinLdX + c * 2);
      dot_lo = fmaf(__uint_as_float(x_lo << 16), acc[n][0], dot_lo);
      dot_lo = fmaf(__uint_as_float(x_lo & 0xffff0000u), acc[n][1], dot_lo);
      dot_hi = fmaf(__uint_as_float(x_hi << 16), acc[n][2], dot_hi);
      dot_hi = fmaf(__uint_as_float(x_hi & 0xffff0000u), acc[n][3], dot_hi);
    }
    dot_lo += __shfl_xor_sync(0xffffffffu, dot_lo, 1); dot_lo += __shfl_xor_sync(0xffffffffu, dot_lo, 2);
    dot_hi += __shfl_xor_sync(0xffffffffu, dot_hi, 1); dot_hi += __shfl_xor_sync(0xffffffffu, dot_hi, 2);
    if (tq == 0) { s_dot[ch][r_lo] = dot_lo; s_dot[ch][r_hi] = dot_hi; }
  }
  __syncthreads();
  if (active) {
    const int r_lo = row0 + g, r_hi = row0 + g + 8;
    const float k_lo = s_inv[r_lo] * s_inv[r_lo] * (s_dot[0][r_lo] + s_dot[1][r_lo]);
    const float k_hi = s_inv[r_hi] * s_inv[r_hi] * (s_dot[0][r_hi] + s_dot[1][r_hi]);
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int c = col0 + n * 8 + tq * 2;
      const float a0 = acc[n][0], a1 = acc[n][1], a2 = acc[n][2], a3 = acc[n][3];
      const uint32_t x_lo = *reinterpret_cast<const uint32_t*>(xs + r_lo * kFinLdX + c * 2);
      const uint32_t x_hi = *reinterpret_cast<const uint32_t*>(xs + r_hi * kFinLdX + c * 2);
      float2 o_lo, o_hi;
      o_lo.x = fmaf(-k_lo, __uint_as_float(x_lo << 16), a0);
      o_lo.y = fmaf(-k_lo, __uint_as_float(x_lo & 0xffff0000u), a1);
      o_hi.x = fmaf(-k_hi, __uint_as_float(x_hi << 16), a2);
      o_hi.y = fmaf(-k_hi, __uint_as_float(x_hi & 0xffff0000u), a3);
      if (OUT_BF16) {
        __nv_bfloat16* ob = (__nv_bfloat16*)out_base + out_off;
        *reinterpret_cast<__nv_bfloat162*>(ob + (size_t)r_lo * 256 + c) = __floats2bfloat162_rn(o_lo.x, o_lo.y);
        *reinterpret_cast<__nv_bfloat162*>(ob + (size_t)r_hi * 256 + c) = __floats2bfloat162_rn(o_hi.x, o_hi.y);
      } else {
        float* of = (float*)out_base + out_off;
        *reinterpret_cast<float2*>(of + (size_t)r_lo * 256 + c) = o_lo;
        *reinterpret_cast<float2*>(of + (size_t)r_hi * 256 + c) = o_hi;
      }
    }
  }
}

// both sides in one launch (tcgen05 path)
int run_grad_finalize_tc(const void* v, const void* t, const float* alpha, const float* beta, const float* dvh,
                         const float* dth, const float* dgv, const float* dgt, int Bv, int Bt, int R, int W, int dgt_swz,
                         void* d_im, void* d_s, int out_dtype, cudaStream_t st) {
  const int Lmax = R > W ? R : W;
  const size_t smem = (size_t)Lmax * kFinLdX + (size_t)Lmax * (Lmax + 8) * 2;
  const int threads = 64 * (Lmax / 16);        // one warp per (16-row block, 128-column half)
  auto launch = [&](auto kern) -> int {
    // two CTAs per SM need the large shared-memory carveout (the default one fits a single 57 KB block)
    DVLP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * kFinLdX + 128 * 136 * 2));
    DVLP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    kern<<<Bv + Bt, threads, smem, st>>>((const __nv_bfloat16*)v, (const __nv_bfloat16*)t, alpha, beta, dvh, dth, dgv, dgt,
                                         Bv, R, W, dgt_swz, d_im, d_s);
    return DVLP_OK;
  };
  int rc;
  if (threads <= 320)
    rc = out_dtype == DVLP_BF16 ? launch(k_finalize_item<true, 320>) : launch(k_finalize_item<false, 320>);
  else
    rc = out_dtype == DVLP_BF16 ? launch(k_finalize_item<true, 512>) : launch(k_finalize_item<false, 512>);
  if (rc) return rc;
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

// No Gram term (the direct-form video side of pair_bwd3.cu): the plain l2-norm backward, one warp per row.
template <typename Tin>
__global__ void __launch_bounds__(256) k_l2norm_bwd_rows(const Tin* __restrict__ x, const float* __restrict__ inv,
                                                         const float* __restrict__ dxh, int rows, int d, int prescaled,
                                                         float* __restrict__ d_x) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float a = inv[row];
  const float gsc = prescaled ? 1.f / a : 1.f;
  const Tin* xr = x + (size_t)row * d;
  const float* gr = dxh + (size_t)row * d;
  float dot = 0.f;
  for (int k = lane; k < d; k += 32) dot = fmaf(to_f32(xr[k]) * a, gr[k] * gsc, dot);
  dot = warp_sum(dot);
  for (int k = lane; k < d; k += 32) d_x[(size_t)row * d + k] = a * (gr[k] * gsc - to_f32(xr[k]) * a * dot);
}

int run_grad_finalize(const void* x, int in_dtype, const float* inv, const float* dxh, const float* dG, int items,
                      int L, int d, int prescaled, float* d_x, cudaStream_t st) {
  if (!dG) {
    const int rows = items * L;
    if (in_dtype == DVLP_BF16)
      k_l2norm_bwd_rows<__nv_bfloat16><<<(rows + 7) / 8, 256, 0, st>>>((const __nv_bfloat16*)x, inv, dxh, rows, d, prescaled, d_x);
    else
      k_l2norm_bwd_rows<float><<<(rows + 7) / 8, 256, 0, st>>>((const float*)x, inv, dxh, rows, d, prescaled, d_x);
    DVLP_LAUNCH_OK();
    return DVLP_OK;
  }
  const size_t smem = ((size_t)L * 16 + 16 * 8 + 16) * 4;
  if (d <= 256 && smem <= 48 * 1024) {
    if (in_dtype == DVLP_BF16)
      k_grad_finalize_tile<__nv_bfloat16><<<items, 256, smem, st>>>((const __nv_bfloat16*)x, inv, dxh, dG, L, d,
                                                                    prescaled, d_x);
    else
      k_grad_finalize_tile<float><<<items, 256, smem, st>>>((const float*)x, inv, dxh, dG, L, d, prescaled, d_x);
    DVLP_LAUNCH_OK();
    return DVLP_OK;
  }
  if (in_dtype == DVLP_BF16)
    k_grad_finalize<__nv_bfloat16><<<items * L, 256, 0, st>>>((const __nv_bfloat16*)x, inv, dxh, dG, L, d,
                                                              prescaled, d_x);
  else
    k_grad_finalize<float><<<items * L, 256, 0, st>>>((const float*)x, inv, dxh, dG, L, d, prescaled, d_x);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

}  // namespace dvlp
