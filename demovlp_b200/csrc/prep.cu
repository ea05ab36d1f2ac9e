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
                             float* __restrict__ mterm, float* __restrict__ mterm2) {
  int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (item >= items) return;
  const float* m = mask + (size_t)item * L;
  float mx = -INFINITY;
  for (int c = lane; c < L; c += 32) mx = fmaxf(mx, m[c]);
  mx = warp_max(mx);
  for (int c = lane; c < L; c += 32) {
    float v = lambda * (m[c] - mx);
    mterm[(size_t)item * L + c] = v;
    // tcgen05 path: exp(z - lambda) evaluated as ex2(A * lambda*log2e + mterm2), A <= 1 keeps it <= 0
    if (mterm2) mterm2[(size_t)item * L + c] = (v - lambda) * kLog2e;
  }
}

// G[item][a][b] = inv[a] inv[b] sum_k x[a][k] x[b][k]; one CTA per item, symmetric fill.
template <typename Tin>
__global__ void k_gram(const Tin* __restrict__ x, const float* __restrict__ inv, int L, int d,
                       float* __restrict__ G, __nv_bfloat16* __restrict__ Gb) {
  int item = blockIdx.x;
  const Tin* X = x + (size_t)item * L * d;
  const float* iv = inv + (size_t)item * L;
  float* g = G + (size_t)item * L * L;
  __nv_bfloat16* gb = Gb ? Gb + (size_t)item * L * L : nullptr;
  for (int idx = threadIdx.x; idx < L * L; idx += blockDim.x) {
    int a = idx / L, b = idx - a * L;
    const Tin* pa = X + (size_t)a * d;
    const Tin* pb = X + (size_t)b * d;
    float acc = 0.f;
    for (int k = 0; k < d; ++k) acc = fmaf(to_f32(pa[k]), to_f32(pb[k]), acc);
    acc *= iv[a] * iv[b];
    g[idx] = acc;
    if (gb) gb[idx] = __float2bfloat16(acc);
  }
}

template <typename Tin>
static int prep_side(const Tin* x, const float* mask, int items, int L, int d, float lambda, float* inv,
                     float* mterm, float* mterm2, float* xh, float* G, __nv_bfloat16* Gb, cudaStream_t st) {
  int rows = items * L;
  k_prep_rows<Tin><<<(rows + 7) / 8, 256, 0, st>>>(x, rows, d, inv, xh, nullptr);
  DVLP_LAUNCH_OK();
  k_mask_terms<<<(items + 7) / 8, 256, 0, st>>>(mask, items, L, lambda, mterm, mterm2);
  DVLP_LAUNCH_OK();
  k_gram<Tin><<<items, 256, 0, st>>>(x, inv, L, d, G, Gb);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

__global__ void k_to_bf16(const float* __restrict__ x, size_t n, __nv_bfloat16* __restrict__ o) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = __float2bfloat16(x[i]);
}

// Per-item prologue for both sides.  Generic fields (fp32 normalised rows, fp32 Gram, natural mask
// terms) are always produced; with tc=true the bf16 Gram copies and log2-domain mask terms the
// tcgen05 kernels read are produced as well, and fp32 inputs are first rounded to bf16 ("bf16 in")
// so that every later stage sees exactly the operands the tensor cores see.
int run_prep(const void* im, const void* s, const float* im_mask, const float* s_mask, int Bv, int Bt,
             int R, int W, int d, float lambda, int in_dtype, bool tc, char* ws, const Workspace& L,
             cudaStream_t st) {
  float* alpha = (float*)(ws + L.alpha);
  float* beta = (float*)(ws + L.beta);
  float* mv = (float*)(ws + L.mterm_v);
  float* mt = (float*)(ws + L.mterm_t);
  float* vh = (float*)(ws + L.vh);
  float* th = (float*)(ws + L.th);
  float* gv = (float*)(ws + L.gv);
  float* gt = (float*)(ws + L.gt);
  if (!tc) {
    int rc;
    if (in_dtype == DVLP_BF16) {
      rc = prep_side((const __nv_bfloat16*)im, im_mask, Bv, R, d, lambda, alpha, mv, nullptr, vh, gv, nullptr, st);
      if (rc) return rc;
      return prep_side((const __nv_bfloat16*)s, s_mask, Bt, W, d, lambda, beta, mt, nullptr, th, gt, nullptr, st);
    }
    rc = prep_side((const float*)im, im_mask, Bv, R, d, lambda, alpha, mv, nullptr, vh, gv, nullptr, st);
    if (rc) return rc;
    return prep_side((const float*)s, s_mask, Bt, W, d, lambda, beta, mt, nullptr, th, gt, nullptr, st);
  }
  const __nv_bfloat16* vb = (const __nv_bfloat16*)im;
  const __nv_bfloat16* tb = (const __nv_bfloat16*)s;
  {
    // text-side vectors are read in whole 128-row tiles: keep the padded tail finite
    const size_t nt_pad = ((size_t)Bt * W + 127) / 128 * 128;
    DVLP_CUDA_OK(cudaMemsetAsync(ws + L.beta, 0, nt_pad * 4, st));
    DVLP_CUDA_OK(cudaMemsetAsync(ws + L.m2t, 0, nt_pad * 4, st));
  }
  if (in_dtype == DVLP_F32) {
    size_t nv = (size_t)Bv * R * d, nt = (size_t)Bt * W * d;
    k_to_bf16<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>((const float*)im, nv, (__nv_bfloat16*)(ws + L.vb));
    DVLP_LAUNCH_OK();
    k_to_bf16<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>((const float*)s, nt, (__nv_bfloat16*)(ws + L.tb));
    DVLP_LAUNCH_OK();
    vb = (const __nv_bfloat16*)(ws + L.vb);
    tb = (const __nv_bfloat16*)(ws + L.tb);
  }
  int rc = prep_side(vb, im_mask, Bv, R, d, lambda, alpha, mv, (float*)(ws + L.m2v), vh, gv,
                     (__nv_bfloat16*)(ws + L.gvb), st);
  if (rc) return rc;
  return prep_side(tb, s_mask, Bt, W, d, lambda, beta, mt, (float*)(ws + L.m2t), th, gt,
                   (__nv_bfloat16*)(ws + L.gtb), st);
}

}  // namespace dvlp
