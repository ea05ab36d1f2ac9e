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

int run_grad_finalize(const void* x, int in_dtype, const float* inv, const float* dxh, const float* dG, int items,
                      int L, int d, int prescaled, float* d_x, cudaStream_t st) {
  if (in_dtype == DVLP_BF16)
    k_grad_finalize<__nv_bfloat16><<<items * L, 256, 0, st>>>((const __nv_bfloat16*)x, inv, dxh, dG, L, d,
                                                              prescaled, d_x);
  else
    k_grad_finalize<float><<<items * L, 256, 0, st>>>((const float*)x, inv, dxh, dG, L, d, prescaled, d_x);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

}  // namespace dvlp
