// Gradient finalisation: from gradients w.r.t. the l2-normalised rows (plus the Gram-matrix
// gradient) to gradients w.r.t. the raw inputs -- the backward of model/loss.py:333-338
// (l2norm) composed with G = Xh Xh^T.
//
//   g_r   = dXh_r [/ inv_r when the GEMM left an extra inv_r factor] + 2 sum_c dG[r][c] Xh_c
//   d_x_r = inv_r ( g_r - Xh_r (Xh_r . g_r) )            inv_r = 1/(|x_r| + eps)
#include "common.cuh"

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

int run_grad_finalize(const void* x, int in_dtype, const float* inv, const float* dxh, const float* dG, int items,
                      int L, int d, int prescaled, float* d_x, cudaStream_t st) {
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
