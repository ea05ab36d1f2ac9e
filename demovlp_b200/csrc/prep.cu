// Per-item prologue: inverse row norms, mask terms, Gram matrices.
//
// Reference arithmetic restated here:
//   l2norm(X) = X / (sqrt(sum X^2) + 1e-8)            model/loss.py:333-338 (called at :225-226)
// The Gram matrices G = Xh Xh^T (Xh = l2-normalised rows of one video / one text) are what
// lets the kernels evaluate |weighted context|^2 = q^T G q without forming the (.., d)-wide
// weighted-context tensor of model/loss.py:262-269 ("Gram form").
#include "common.cuh"
#include "mma_sync.cuh"

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

// Fast Gram for the tcgen05 path (bf16 rows, d == 256, L <= 128): the item's rows are staged once in
// shared memory as fp32 (row stride d+1 words: conflict-free) and each thread owns 4x4 blocks of G.
__global__ void __launch_bounds__(256) k_gram_tile(const __nv_bfloat16* __restrict__ x, const float* __restrict__ inv,
                                                   int L, int d, float* __restrict__ G, __nv_bfloat16* __restrict__ Gb) {
  extern __shared__ float xs[];                 // [L][d + 1], rows pre-scaled by inv
  const int item = blockIdx.x;
  const __nv_bfloat16* X = x + (size_t)item * L * d;
  const float* iv = inv + (size_t)item * L;
  const int ld = d + 1;
  for (int e = threadIdx.x; e < L * d / 2; e += blockDim.x) {
    const int row = (2 * e) / d, col = (2 * e) - row * d;
    const __nv_bfloat162 v = reinterpret_cast<const __nv_bfloat162*>(X)[e];
    const float a = iv[row];
    xs[row * ld + col] = __bfloat162float(v.x) * a;
    xs[row * ld + col + 1] = __bfloat162float(v.y) * a;
  }
  __syncthreads();
  const int nb = (L + 3) / 4;
  for (int blk = threadIdx.x; blk < nb * nb; blk += blockDim.x) {
    // block (ab, bb) owns rows {ab + h*nb} x columns {bb + h*nb}: consecutive threads read consecutive
    // shared-memory rows (row stride d+1 words), i.e. distinct banks
    const int ab = blk / nb, bb = blk % nb;
    if (bb < ab) continue;                      // symmetric: compute the upper blocks, mirror on store
    float acc[4][4] = {};
    int ra[4], rb[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) { ra[h] = min(ab + h * nb, L - 1) * ld; rb[h] = min(bb + h * nb, L - 1) * ld; }
    for (int k = 0; k < d; ++k) {
      float va[4], vb[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) { va[h] = xs[ra[h] + k]; vb[h] = xs[rb[h] + k]; }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(va[p], vb[q], acc[p][q]);
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int aa = ab + p * nb, bq = bb + q * nb;
        if (aa < L && bq < L) {
          const size_t i1 = (size_t)item * L * L + (size_t)aa * L + bq, i2 = (size_t)item * L * L + (size_t)bq * L + aa;
          if (G) { G[i1] = acc[p][q]; G[i2] = acc[p][q]; }
          if (Gb) { Gb[i1] = __float2bfloat16(acc[p][q]); Gb[i2] = __float2bfloat16(acc[p][q]); }
        }
      }
  }
}


// Fused per-item prologue of the tcgen05 path (bf16 rows, d == 256, L % 16 == 0, L <= 128): one CTA per
// video / text stages the item's rows in shared memory once and produces the inverse row norms, both mask
// terms and the bf16 Gram matrix G = Xh Xh^T.  The Gram product runs on mma.sync with the RAW bf16 rows
// (exact products, fp32 accumulate) and is scaled by inv_a inv_b afterwards, like the pair kernels do for S.
// Block Bv + Bt (if launched) zeroes the padded tail of the text-side vectors.
constexpr int kPrepLd = (256 + 8) * 2;     // shared-memory row stride in bytes: conflict-free ldmatrix
__global__ void __launch_bounds__(512) k_prep_item(const __nv_bfloat16* __restrict__ v, const __nv_bfloat16* __restrict__ t,
                                                   const float* __restrict__ vmask, const float* __restrict__ tmask,
                                                   int Bv, int Bt, int R, int W, float lambda,
                                                   float* __restrict__ alpha, float* __restrict__ beta,
                                                   float* __restrict__ mterm_v, float* __restrict__ mterm_t,
                                                   float* __restrict__ m2v, float* __restrict__ m2t,
                                                   __nv_bfloat16* __restrict__ gvb, __nv_bfloat16* __restrict__ gtb,
                                                   __nv_bfloat16* __restrict__ gbd, int pad_from, int pad_to) {
  extern __shared__ __align__(16) uint8_t psm[];
  __shared__ float s_inv[128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if ((int)blockIdx.x == Bv + Bt) {            // padded tail of beta / m2t (whole 128-row tiles are bulk-copied)
    for (int e = pad_from + threadIdx.x; e < pad_to; e += blockDim.x) { beta[e] = 0.f; m2t[e] = 0.f; }
    // rows of the missing texts in the last block-diagonal Gram tile
    if (gbd)
      for (int e = pad_from * 128 + threadIdx.x; e < pad_to * 128; e += blockDim.x) gbd[e] = __float2bfloat16(0.f);
    return;
  }
  const bool text = (int)blockIdx.x >= Bv;
  const int item = text ? blockIdx.x - Bv : blockIdx.x;
  const int L = text ? W : R;
  const __nv_bfloat16* X = (text ? t : v) + (size_t)item * L * 256;
  const float* mask = (text ? tmask : vmask) + (size_t)item * L;
  float* inv_out = (text ? beta : alpha) + (size_t)item * L;
  float* mt = (text ? mterm_t : mterm_v) + (size_t)item * L;
  float* m2 = (text ? m2t : m2v) + (size_t)item * L;
  __nv_bfloat16* G = (text ? gtb : gvb) + (size_t)item * L * L;
  // block-diagonal copy of the text Gram matrix (W == 32): rows item*32.. of a (tiles*128) x 128 matrix,
  // this text's 32 x 32 block at columns (item % 4) * 32, zeros elsewhere (pair_bwd2.cu: Y2 = Gbd U2)
  __nv_bfloat16* Gd = (text && gbd) ? gbd + (size_t)item * 32 * 128 : nullptr;
  const int dcol = (item & 3) * 32;
  if (Gd)
    for (int e = threadIdx.x; e < 32 * 128; e += blockDim.x)
      if (((e & 127) >> 5) != (item & 3)) Gd[e] = __float2bfloat16(0.f);
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(psm);
  // stage rows (16-byte chunks: 32 per row)
  for (int e = threadIdx.x; e < L * 32; e += blockDim.x) {
    const int row = e >> 5, ch = e & 31;
    const uint4 val = reinterpret_cast<const uint4*>(X)[e];
    *reinterpret_cast<uint4*>(psm + row * kPrepLd + ch * 16) = val;
  }
  __syncthreads();
  // inverse norms: one warp per row, 8 elements per lane
  for (int row = warp; row < L; row += nwarp) {
    const uint4 val = *reinterpret_cast<const uint4*>(psm + row * kPrepLd + lane * 16);
    const uint32_t w4[4] = {val.x, val.y, val.z, val.w};
    float ss = 0.f;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float lo = __uint_as_float(w4[h] << 16), hi = __uint_as_float(w4[h] & 0xffff0000u);
      ss = fmaf(lo, lo, ss);
      ss = fmaf(hi, hi, ss);
    }
    ss = warp_sum(ss);
    const float a = 1.f / (sqrtf(ss) + kEps);
    if (lane == 0) { s_inv[row] = a; inv_out[row] = a; }
  }
  // mask terms (last warp; see k_mask_terms)
  if (warp == nwarp - 1) {
    float mx = -INFINITY;
    for (int c = lane; c < L; c += 32) mx = fmaxf(mx, mask[c]);
    mx = warp_max(mx);
    for (int c = lane; c < L; c += 32) {
      const float val = lambda * (mask[c] - mx);
      mt[c] = val;
      m2[c] = (val - lambda) * kLog2e;
    }
  }
  __syncthreads();
  // Gram: 16 x 16 output tiles, upper triangle computed and mirrored
  const int nt = L >> 4;
  for (int tile = warp; tile < nt * (nt + 1) / 2; tile += nwarp) {     // upper-triangular tiles only
    int ti = 0, tj = tile;
    while (tj >= nt - ti) { tj -= nt - ti; ++ti; }
    tj += ti;
    float acc[2][4] = {};
#pragma unroll 4
    for (int k0 = 0; k0 < 256; k0 += 16) {
      uint32_t a[4], b[4];
      mm::load_a(sb, kPrepLd, ti * 16, k0, lane, a);
      mm::load_b_nk(sb, kPrepLd, tj * 16, k0, lane, b);
      mm::mma_16816(acc[0], a[0], a[1], a[2], a[3], b[0], b[1]);
      mm::mma_16816(acc[1], a[0], a[1], a[2], a[3], b[2], b[3]);
    }
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = ti * 16 + g + half * 8, c = tj * 16 + h * 8 + tq * 2;
        const float x0 = acc[h][half * 2] * s_inv[r] * s_inv[c], x1 = acc[h][half * 2 + 1] * s_inv[r] * s_inv[c + 1];
        *reinterpret_cast<__nv_bfloat162*>(G + (size_t)r * L + c) = __floats2bfloat162_rn(x0, x1);
        if (Gd) *reinterpret_cast<__nv_bfloat162*>(Gd + (size_t)r * 128 + dcol + c) = __floats2bfloat162_rn(x0, x1);
        if (tj != ti) {
          G[(size_t)c * L + r] = __float2bfloat16(x0);
          G[(size_t)(c + 1) * L + r] = __float2bfloat16(x1);
          if (Gd) {
            Gd[(size_t)c * 128 + dcol + r] = __float2bfloat16(x0);
            Gd[(size_t)(c + 1) * 128 + dcol + r] = __float2bfloat16(x1);
          }
        }
      }
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
             int R, int W, int d, float lambda, int in_dtype, bool tc, bool full, char* ws, const Workspace& L,
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
  // text-side vectors are read in whole 128-row tiles: keep the padded tail finite
  const size_t nt_pad = ((size_t)Bt * W + 127) / 128 * 128;
  const bool fused = !full && d == 256 && (R % 16) == 0 && (W % 16) == 0 && R <= 128 && W <= 128;
  if (!fused) {
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
  if (fused) {
    const int Lmax = R > W ? R : W;
    const size_t smem = (size_t)Lmax * kPrepLd;
    static DeviceOnce attr_once;
    if (attr_once.pending()) {
      DVLP_CUDA_OK(cudaFuncSetAttribute(k_prep_item, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * kPrepLd));
      attr_once.done();
    }
    const bool pad = nt_pad > (size_t)Bt * W;
    k_prep_item<<<Bv + Bt + (pad ? 1 : 0), Lmax > 64 ? 512 : 256, smem, st>>>(
        vb, tb, im_mask, s_mask, Bv, Bt, R, W, lambda, alpha, beta, mv, mt, (float*)(ws + L.m2v), (float*)(ws + L.m2t),
        (__nv_bfloat16*)(ws + L.gvb), (__nv_bfloat16*)(ws + L.gtb), W == 32 ? (__nv_bfloat16*)(ws + L.gbd) : nullptr, Bt * W,
        (int)nt_pad);
    DVLP_LAUNCH_OK();
    return DVLP_OK;
  }
  // lean tcgen05 prologue: inverse norms, mask terms (natural + log2 domain), bf16 Gram.  The fp32
  // normalised rows / fp32 Gram of the generic kernels are produced only when `full` is set
  // (generic backward requested for a tcgen05 forward: DVLP_TC_BWD=0).
  struct Side { const __nv_bfloat16* x; const float* mask; int items, Lc; float *inv, *mt, *m2, *xh, *G; __nv_bfloat16* Gb; };
  Side sides[2] = {{vb, im_mask, Bv, R, alpha, mv, (float*)(ws + L.m2v), vh, gv, (__nv_bfloat16*)(ws + L.gvb)},
                   {tb, s_mask, Bt, W, beta, mt, (float*)(ws + L.m2t), th, gt, (__nv_bfloat16*)(ws + L.gtb)}};
  for (const Side& sd : sides) {
    const int rows = sd.items * sd.Lc;
    k_prep_rows<__nv_bfloat16><<<(rows + 7) / 8, 256, 0, st>>>(sd.x, rows, d, sd.inv, full ? sd.xh : nullptr, nullptr);
    DVLP_LAUNCH_OK();
    k_mask_terms<<<(sd.items + 7) / 8, 256, 0, st>>>(sd.mask, sd.items, sd.Lc, lambda, sd.mt, sd.m2);
    DVLP_LAUNCH_OK();
    const size_t smem = (size_t)sd.Lc * (d + 1) * 4;
    if (smem <= 200 * 1024) {
      static DeviceOnce attr_once;
      if (attr_once.pending()) {
        DVLP_CUDA_OK(cudaFuncSetAttribute(k_gram_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_once.done();
      }
      k_gram_tile<<<sd.items, 256, smem, st>>>(sd.x, sd.inv, sd.Lc, d, full ? sd.G : nullptr, sd.Gb);
    } else {
      k_gram<__nv_bfloat16><<<sd.items, 256, 0, st>>>(sd.x, sd.inv, sd.Lc, d, sd.G, sd.Gb);
    }
    DVLP_LAUNCH_OK();
  }
  return DVLP_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Prologue of the large-shape tcgen05 kernels (pair_fwd3.cu): the per-row / per-item quantities above (inverse norms,
// log2-domain mask terms, bf16 text Gram matrices) re-laid out for its padded tiles:
//   alpha3 / m2v3 [Bv][Rp]        regions >= R: alpha 0, mask term -1e30 (their exponentials are exact zeros)
//   beta3 / m2t3  [tiles][128]    tile row = text slot * WP + word; padded words and absent texts: 0 / -1e30
//   gbd3          [tiles][128][128] bf16: the TPT text Gram matrices of a tile on the diagonal, zeros elsewhere
__global__ void __launch_bounds__(256) k_pack3(const float* __restrict__ alpha, const float* __restrict__ m2v,
                                               const float* __restrict__ beta, const float* __restrict__ m2t,
                                               const __nv_bfloat16* __restrict__ gtb, int Bv, int Bt, int R, int W, int Rp,
                                               int WP, int TPT, int tiles, int n_t, int tpg, float* __restrict__ alpha3,
                                               float* __restrict__ m2v3, float* __restrict__ beta3,
                                               float* __restrict__ m2t3, __nv_bfloat16* __restrict__ gbd3) {
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t e = t0; e < (size_t)Bv * Rp; e += stride) {
    const int i = (int)(e / Rp), r = (int)(e - (size_t)i * Rp);
    const bool ok = r < R;
    alpha3[e] = ok ? alpha[(size_t)i * R + r] : 0.f;
    m2v3[e] = ok ? m2v[(size_t)i * R + r] : -1e30f;
  }
  for (size_t e = t0; e < (size_t)tiles * 128; e += stride) {
    const int row = (int)(e & 127), tile = (int)(e >> 7);
    // texts of a tile: tile tt of group g holds texts g n_t + tt TPT + slot (dense: one group of Bt texts)
    const int ts = row / WP, w = row - ts * WP, jg = (tile % tpg) * TPT + ts, j = (tile / tpg) * n_t + jg;
    const bool ok = jg < n_t && j < Bt && w < W;
    beta3[e] = ok ? beta[(size_t)j * W + w] : 0.f;
    m2t3[e] = ok ? m2t[(size_t)j * W + w] : -1e30f;
  }
  for (size_t e = t0; e < (size_t)tiles * 128 * 64; e += stride) {       // two bf16 columns per thread
    const int col = (int)(e & 63) * 2, row = (int)((e >> 6) & 127), tile = (int)(e >> 13);
    const int ts = row / WP, w = row - ts * WP, jg = (tile % tpg) * TPT + ts, j = (tile / tpg) * n_t + jg;
    float v[2] = {0.f, 0.f};
    if (jg < n_t && j < Bt && w < W) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = col + h, cs = c / WP, w2 = c - cs * WP;
        if (cs == ts && w2 < W) v[h] = __bfloat162float(gtb[((size_t)j * W + w) * W + w2]);
      }
    }
    reinterpret_cast<__nv_bfloat162*>(gbd3)[e] = __floats2bfloat162_rn(v[0], v[1]);
  }
}

int run_prep3(const void* im, const void* s, const float* im_mask, const float* s_mask, int Bv, int Bt, int R, int W,
              int d, float lambda, int in_dtype, char* ws, const Workspace& L, cudaStream_t st, int n_t) {
  Shape3 s3;
  if (!shape3(R, W, d, &s3)) return DVLP_ERR_BAD_SHAPE;
  const __nv_bfloat16* vb = (const __nv_bfloat16*)im;
  const __nv_bfloat16* tb = (const __nv_bfloat16*)s;
  if (in_dtype == DVLP_F32) {      // "bf16 in": every later stage sees exactly the operands the tensor cores see
    size_t nv = (size_t)Bv * R * d, nt = (size_t)Bt * W * d;
    k_to_bf16<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>((const float*)im, nv, (__nv_bfloat16*)(ws + L.vb));
    DVLP_LAUNCH_OK();
    k_to_bf16<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>((const float*)s, nt, (__nv_bfloat16*)(ws + L.tb));
    DVLP_LAUNCH_OK();
    vb = (const __nv_bfloat16*)(ws + L.vb);
    tb = (const __nv_bfloat16*)(ws + L.tb);
  }
  float* alpha = (float*)(ws + L.alpha);
  float* beta = (float*)(ws + L.beta);
  k_prep_rows<__nv_bfloat16><<<(Bv * R + 7) / 8, 256, 0, st>>>(vb, Bv * R, d, alpha, nullptr, nullptr);
  DVLP_LAUNCH_OK();
  k_prep_rows<__nv_bfloat16><<<(Bt * W + 7) / 8, 256, 0, st>>>(tb, Bt * W, d, beta, nullptr, nullptr);
  DVLP_LAUNCH_OK();
  k_mask_terms<<<(Bv + 7) / 8, 256, 0, st>>>(im_mask, Bv, R, lambda, (float*)(ws + L.mterm_v), (float*)(ws + L.m2v));
  DVLP_LAUNCH_OK();
  k_mask_terms<<<(Bt + 7) / 8, 256, 0, st>>>(s_mask, Bt, W, lambda, (float*)(ws + L.mterm_t), (float*)(ws + L.m2t));
  DVLP_LAUNCH_OK();
  {
    const size_t smem = (size_t)W * (d + 1) * 4;
    static DeviceOnce attr_once;
    if (attr_once.pending()) {
      DVLP_CUDA_OK(cudaFuncSetAttribute(k_gram_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_once.done();
    }
    k_gram_tile<<<Bt, 256, smem, st>>>(tb, beta, W, d, nullptr, (__nv_bfloat16*)(ws + L.gtb));
    DVLP_LAUNCH_OK();
  }
  const int tpg = n_t > 0 ? (n_t + s3.TPT - 1) / s3.TPT : (Bt + s3.TPT - 1) / s3.TPT;
  const int tiles = n_t > 0 ? Bv * tpg : tpg;
  const size_t work = (size_t)tiles * 128 * 64;
  int grid = (int)((work + 256 * 4 - 1) / (256 * 4));
  grid = grid < 1 ? 1 : (grid > 148 * 8 ? 148 * 8 : grid);
  k_pack3<<<grid, 256, 0, st>>>(alpha, (const float*)(ws + L.m2v), beta, (const float*)(ws + L.m2t),
                                (const __nv_bfloat16*)(ws + L.gtb), Bv, Bt, R, W, s3.Rp, s3.WP, s3.TPT, tiles, n_t > 0 ? n_t : Bt, tpg,
                                (float*)(ws + L.alpha3), (float*)(ws + L.m2v3), (float*)(ws + L.beta3),
                                (float*)(ws + L.m2t3), (__nv_bfloat16*)(ws + L.gbd3));
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

}  // namespace dvlp
