// Per-item prologue: inverse row norms, mask terms, Gram matrices.
//
// Reference arithmetic restated here:
//   l2norm(X) = X / (sqrt(sum X^2) + 1e-8)            model/loss.py:333-338 (called at :225-226)
// The Gram matrices G = Xh Xh^T (Xh = l2-normalised rows of one video / one text) are what
// lets the kernels evaluate |weighted context|^2 = q^T G q without forming the (.., d)-wide
// weighted-context tensor of model/loss.py:262-269 ("Gram form").
#include "common.cuh"

namespace dvlp {

// one warp per row: inv[row] = 1/(|x|+eps); optional normalised fp32 copy / raw bf16 copy
template <typename Tin>
__global__ void k_prep_rows(const Tin* __restrict__ x, int rows, int d, float* __restrict__ inv,
                            float* __restrict__ xh_f32, __nv_bfloat16* __restrict__ xraw_bf16) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const Tin* p = x + (size_t)row * d;
  float ss = 0.f;
  for (int k = lane; k < d; k += 32) {
    float v = to_f32(p[k]);
    if (xraw_bf16) v = __bfloat162float(__float2bfloat16(v));   // norm of what the MMA will see
    ss += v * v;
  }
  ss = warp_sum(ss);
  float a = 1.f / (sqrtf(ss) + kEps);
  if (lane == 0) inv[row] = a;
  if (xh_f32) {
    float* o = xh_f32 + (size_t)row * d;
    for (int k = lane; k < d; k += 32) o[k] = to_f32(p[k]) * a;
  }
  if (xraw_bf16) {
    __nv_bfloat16* o = xraw_bf16 + (size_t)row * d;
    for (int k = lane; k < d; k += 32) o[k] = __float2bfloat16(to_f32(p[k]));
  }
}

// one warp per item: mterm[c] = lambda * (mask[c] - max_c mask[c])
// (softmax over the context is shift invariant, model/loss.py:242-245; subtracting the
//  item's largest mask value keeps a fully masked context finite exactly like the
//  reference's max-subtracting softmax does)
__global__ void k_mask_terms(const float* __restrict__ mask, int items, int L, float lambda,
                             float* __restrict__ mterm) {
  int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (item >= items) return;
  const float* m = mask + (size_t)item * L;
  float mx = -INFINITY;
  for (int c = lane; c < L; c += 32) mx = fmaxf(mx, m[c]);
  mx = warp_max(mx);
  for (int c = lane; c < L; c += 32) mterm[(size_t)item * L + c] = lambda * (m[c] - mx);
}

// G[item][a][b] = inv[a] inv[b] sum_k x[a][k] x[b][k]; one CTA per item, symmetric fill.
template <typename Tin, typename Tout>
__global__ void k_gram(const Tin* __restrict__ x, const float* __restrict__ inv, int L, int d,
                       Tout* __restrict__ G) {
  int item = blockIdx.x;
  const Tin* X = x + (size_t)item * L * d;
  const float* iv = inv + (size_t)item * L;
  Tout* g = G + (size_t)item * L * L;
  for (int idx = threadIdx.x; idx < L * L; idx += blockDim.x) {
    int a = idx / L, b = idx - a * L;
    const Tin* pa = X + (size_t)a * d;
    const Tin* pb = X + (size_t)b * d;
    float acc = 0.f;
    for (int k = 0; k < d; ++k) acc = fmaf(to_f32(pa[k]), to_f32(pb[k]), acc);
    acc *= iv[a] * iv[b];
    if constexpr (sizeof(Tout) == 2)
      g[idx] = __float2bfloat16(acc);
    else
      g[idx] = acc;
  }
}

template <typename Tin>
static int prep_side(const Tin* x, const float* mask, int items, int L, int d, float lambda,
                     float* inv, float* mterm, float* xh, __nv_bfloat16* xraw, void* G, bool g_bf16,
                     cudaStream_t st) {
  int rows = items * L;
  k_prep_rows<Tin><<<(rows + 7) / 8, 256, 0, st>>>(x, rows, d, inv, xh, xraw);
  DVLP_LAUNCH_OK();
  k_mask_terms<<<(items + 7) / 8, 256, 0, st>>>(mask, items, L, lambda, mterm);
  DVLP_LAUNCH_OK();
  if (g_bf16 && xraw)
    k_gram<__nv_bfloat16, __nv_bfloat16><<<items, 256, 0, st>>>(xraw, inv, L, d, (__nv_bfloat16*)G);
  else if (g_bf16)
    k_gram<Tin, __nv_bfloat16><<<items, 256, 0, st>>>(x, inv, L, d, (__nv_bfloat16*)G);
  else
    k_gram<Tin, float><<<items, 256, 0, st>>>(x, inv, L, d, (float*)G);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

// Runs the prologue for both sides.  tc=false: fp32 normalised rows + fp32 Gram (generic path);
// tc=true: bf16 Gram, raw bf16 copy only when the input is fp32.
int run_prep(const void* im, const void* s, const float* im_mask, const float* s_mask, int Bv, int Bt,
             int R, int W, int d, float lambda, int in_dtype, bool tc, char* ws, const Workspace& L,
             cudaStream_t st) {
  float* alpha = (float*)(ws + L.alpha);
  float* beta = (float*)(ws + L.beta);
  float* mv = (float*)(ws + L.mterm_v);
  float* mt = (float*)(ws + L.mterm_t);
  int rc;
  if (in_dtype == DVLP_BF16) {
    auto* v = (const __nv_bfloat16*)im;
    auto* t = (const __nv_bfloat16*)s;
    rc = prep_side(v, im_mask, Bv, R, d, lambda, alpha, mv, tc ? nullptr : (float*)(ws + L.vh), nullptr,
                   ws + L.gv, tc, st);
    if (rc) return rc;
    rc = prep_side(t, s_mask, Bt, W, d, lambda, beta, mt, tc ? nullptr : (float*)(ws + L.th), nullptr,
                   ws + L.gt, tc, st);
  } else {
    auto* v = (const float*)im;
    auto* t = (const float*)s;
    rc = prep_side(v, im_mask, Bv, R, d, lambda, alpha, mv, tc ? nullptr : (float*)(ws + L.vh),
                   tc ? (__nv_bfloat16*)(ws + L.vh) : nullptr, ws + L.gv, tc, st);
    if (rc) return rc;
    rc = prep_side(t, s_mask, Bt, W, d, lambda, beta, mt, tc ? nullptr : (float*)(ws + L.th),
                   tc ? (__nv_bfloat16*)(ws + L.th) : nullptr, ws + L.gt, tc, st);
  }
  return rc;
}

}  // namespace dvlp
