// Generic-shape region-word pair kernels (fp32, CUDA cores): any R, W, d.
//
// These serve the shapes the tcgen05 path does not cover (the reference's real
// shapes R=30/240, W=99; the 5x5 multiple-choice eval; odd sizes) and double as the
// high-precision mode for the eval matrix.  One persistent CTA walks (video, text)
// pairs; all R x W intermediates live in a per-CTA scratch slab (L1/L2 resident), so
// nothing of size Bv*Bt*R*W is ever allocated.
//
// Arithmetic (reference: model/loss.py, lines cited per step; DESIGN.md has the derivation):
//   S   = Vh Th^T                      :225-235   normalised rows -> cosine of every (region, word)
//   L   = LeakyReLU_0.1(S)             :236
//   per direction (context axis c, query axis q):
//     A   = L / (||L||_over-q + eps)   :238
//     P   = softmax_c(lambda (A + mask_c))            :242-245  (the query-side mask is a per-row
//                                                                constant of the softmax: no effect)
//     H   = [P * Lc - sum_c P > 0]  (focal "equal")   :251-254, :274-283
//     Q   = H P / sum_c H P                           :257-259
//     cos_q = (Q^T S)_q / sqrt(Q_q^T G Q_q)           :262-269 + :286-291 in Gram form, G = Xh Xh^T
//   score = mean_w cos^A_w + mean_r cos^B_r           :317-318, :326-329
#include "common.cuh"

namespace dvlp {


struct Dir {
  int Lc, Lq, sc, sq;   // element (c, q) of an R x W buffer lives at c*sc + q*sq
  const float* mterm;   // (Lc) lambda * (mask - max mask)
  const float* G;       // (Lc, Lc)
};

// forward of one attention direction; returns sum_q cos_q in red[slot] (smem)
__device__ void dir_forward(const Dir D, const float* S, float* Q, float* Y, float* nrm, float* num,
                            float* den, float* cosv, float lambda, int focal, float* red_slot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  // (1) l2 norm of LeakyReLU(S) over the query axis, one value per context element
  for (int c = warp; c < D.Lc; c += nwarp) {
    float ss = 0.f;
    for (int q = lane; q < D.Lq; q += 32) {
      float l = leaky(S[c * D.sc + q * D.sq]);
      ss = fmaf(l, l, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) nrm[c] = sqrtf(ss) + kEps;
  }
  __syncthreads();
  // (2) masked softmax over the context, focal selection, renormalise; numerator Q^T S
  for (int q = warp; q < D.Lq; q += nwarp) {
    float m = -INFINITY;
    for (int c = lane; c < D.Lc; c += 32) {
      float z = lambda * leaky(S[c * D.sc + q * D.sq]) / nrm[c] + D.mterm[c];
      m = fmaxf(m, z);
    }
    m = warp_max(m);
    float E = 0.f;
    for (int c = lane; c < D.Lc; c += 32) {
      float z = lambda * leaky(S[c * D.sc + q * D.sq]) / nrm[c] + D.mterm[c];
      E += __expf(z - m);
    }
    E = warp_sum(E);
    float U = 0.f;
    for (int c = lane; c < D.Lc; c += 32) {
      float z = lambda * leaky(S[c * D.sc + q * D.sq]) / nrm[c] + D.mterm[c];
      float e = __expf(z - m);
      bool keep = focal ? (e * (float)D.Lc > E) : true;
      float u = keep ? e : 0.f;
      Q[c * D.sc + q * D.sq] = u;
      U += u;
    }
    U = warp_sum(U);
    float nm = 0.f;
    for (int c = lane; c < D.Lc; c += 32) {
      int idx = c * D.sc + q * D.sq;
      float qv = Q[idx] / U;
      Q[idx] = qv;
      nm = fmaf(qv, S[idx], nm);
    }
    nm = warp_sum(nm);
    if (lane == 0) num[q] = nm;
  }
  __syncthreads();
  // (3) Y = G Q  (per query column)
  const int n = D.Lc * D.Lq;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    int c, q;
    if (D.sq == 1) { c = e / D.Lq; q = e - c * D.Lq; } else { q = e / D.Lc; c = e - q * D.Lc; }
    const float* g = D.G + (size_t)c * D.Lc;
    float acc = 0.f;
    for (int c2 = 0; c2 < D.Lc; ++c2) acc = fmaf(g[c2], Q[c2 * D.sc + q * D.sq], acc);
    Y[c * D.sc + q * D.sq] = acc;
  }
  __syncthreads();
  // (4) cos_q = num_q / sqrt(Q_q^T Y_q)
  for (int q = warp; q < D.Lq; q += nwarp) {
    float dd = 0.f;
    for (int c = lane; c < D.Lc; c += 32) {
      int idx = c * D.sc + q * D.sq;
      dd = fmaf(Q[idx], Y[idx], dd);
    }
    dd = warp_sum(dd);
    if (lane == 0) {
      float dn = sqrtf(dd);
      float cs = dd > 1e-30f ? num[q] / dn : 0.f;   // cosine clamp of :291: zero context -> 0
      den[q] = dn;
      cosv[q] = cs;
    }
  }
  __syncthreads();
  // fixed-order sum of the cosines: results do not depend on scheduling
  if (warp == 0) {
    float csum = 0.f;
    for (int q = lane; q < D.Lq; q += 32) csum += cosv[q];
    csum = warp_sum(csum);
    if (lane == 0) *red_slot = csum;
  }
  __syncthreads();
}

// backward of one direction: accumulates into dS (R x W) and the item's dG (global atomics)
__device__ void dir_backward(const Dir D, const float* S, const float* Q, float* Y, const float* nrm,
                             const float* den, const float* cosv, float* tvec, float* dS, float* dG,
                             float lambda, float g) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const float gam = g / (float)D.Lq;
  const int n = D.Lc * D.Lq;
  // (a) dA = lambda * (gam/den_q) * Q * (S - (cos_q/den_q) Y), stored over Y
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    int c, q;
    if (D.sq == 1) { c = e / D.Lq; q = e - c * D.Lq; } else { q = e / D.Lc; c = e - q * D.Lc; }
    int idx = c * D.sc + q * D.sq;
    float dn = den[q];
    float v = 0.f;
    if (dn * dn > 1e-30f) v = lambda * (gam / dn) * Q[idx] * (S[idx] - (cosv[q] / dn) * Y[idx]);
    Y[idx] = v;
  }
  __syncthreads();
  // (b) t_c = sum_q A dA
  for (int c = warp; c < D.Lc; c += nwarp) {
    float t = 0.f;
    for (int q = lane; q < D.Lq; q += 32) {
      int idx = c * D.sc + q * D.sq;
      t = fmaf(leaky(S[idx]) / nrm[c], Y[idx], t);
    }
    t = warp_sum(t);
    if (lane == 0) tvec[c] = t;
  }
  __syncthreads();
  // (c) dS += (dA - A t_c)/nrm_c * leaky'(S) + (gam/den_q) Q
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    int c, q;
    if (D.sq == 1) { c = e / D.Lq; q = e - c * D.Lq; } else { q = e / D.Lc; c = e - q * D.Lc; }
    int idx = c * D.sc + q * D.sq;
    float s = S[idx];
    float a = leaky(s) / nrm[c];
    float dn = den[q];
    float direct = dn * dn > 1e-30f ? (gam / dn) * Q[idx] : 0.f;
    dS[idx] += (Y[idx] - a * tvec[c]) / nrm[c] * leaky_grad(s) + direct;
  }
  // (d) dG[c][c2] += sum_q dD_q Q[c,q] Q[c2,q],  dD_q = -gam cos_q / (2 den_q^2)
  for (int e = threadIdx.x; e < D.Lc * D.Lc; e += blockDim.x) {
    int c = e / D.Lc, c2 = e - c * D.Lc;
    float acc = 0.f;
    for (int q = 0; q < D.Lq; ++q) {
      float dn = den[q];
      float dD = dn * dn > 1e-30f ? -gam * cosv[q] / (2.f * dn * dn) : 0.f;
      acc = fmaf(dD * Q[c * D.sc + q * D.sq], Q[c2 * D.sc + q * D.sq], acc);
    }
    atomicAdd(dG + e, acc);
  }
  __syncthreads();
}

template <bool kBackward>
__global__ void __launch_bounds__(256) k_pair_generic(PairArgs a) {
  __shared__ float red[2];
  const int R = a.R, W = a.W, d = a.d, RW = R * W;
  const int Lmax = R > W ? R : W;
  float* base = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;
  float* S = base;
  float* Q = S + RW;
  float* Y = Q + RW;
  float* dS = Y + RW;
  float* nrm = dS + RW;
  float* num = nrm + Lmax;
  float* den = num + Lmax;
  float* cosv = den + Lmax;
  float* tvec = cosv + Lmax;

  const long long n_pairs = (long long)a.Bv * a.Bt;
  for (long long pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
    const int i = (int)(pair / a.Bt), j = (int)(pair - (long long)i * a.Bt);
    const float* V = a.vh + (size_t)i * R * d;
    const float* T = a.th + (size_t)j * W * d;
    if (threadIdx.x < 2) red[threadIdx.x] = 0.f;
    // S = Vh Th^T
    for (int e = threadIdx.x; e < RW; e += blockDim.x) {
      int r = e / W, w = e - r * W;
      const float4* pv = reinterpret_cast<const float4*>(V + (size_t)r * d);
      const float4* pt = reinterpret_cast<const float4*>(T + (size_t)w * d);
      float acc = 0.f;
      for (int k = 0; k < d / 4; ++k) {
        float4 x = pv[k], y = pt[k];
        acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc);
        acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
      }
      S[e] = acc;
      if (kBackward) dS[e] = 0.f;
    }
    __syncthreads();
    Dir dA{R, W, W, 1, a.mterm_v + (size_t)i * R, a.gv + (size_t)i * R * R};   // context = regions
    Dir dB{W, R, 1, W, a.mterm_t + (size_t)j * W, a.gt + (size_t)j * W * W};   // context = words
    float g = kBackward ? a.d_scores[pair] : 0.f;

    dir_forward(dA, S, Q, Y, nrm, num, den, cosv, a.lambda, a.focal, &red[0]);
    if (kBackward)
      dir_backward(dA, S, Q, Y, nrm, den, cosv, tvec, dS, a.dgv + (size_t)i * R * R, a.lambda, g);
    dir_forward(dB, S, Q, Y, nrm, num, den, cosv, a.lambda, a.focal, &red[1]);
    if (kBackward)
      dir_backward(dB, S, Q, Y, nrm, den, cosv, tvec, dS, a.dgt + (size_t)j * W * W, a.lambda, g);

    if (!kBackward) {
      if (threadIdx.x == 0) a.scores[pair] = red[0] / (float)W + red[1] / (float)R;
    } else {
      // dVh[i] += dS Th[j] ;  dTh[j] += dS^T Vh[i]
      float* dV = a.dvh + (size_t)i * R * d;
      float* dT = a.dth + (size_t)j * W * d;
      for (int k = threadIdx.x; k < d; k += blockDim.x) {
        for (int r = 0; r < R; ++r) {
          float acc = 0.f;
          for (int w = 0; w < W; ++w) acc = fmaf(dS[r * W + w], T[(size_t)w * d + k], acc);
          atomicAdd(dV + (size_t)r * d + k, acc);
        }
        for (int w = 0; w < W; ++w) {
          float acc = 0.f;
          for (int r = 0; r < R; ++r) acc = fmaf(dS[r * W + w], V[(size_t)r * d + k], acc);
          atomicAdd(dT + (size_t)w * d + k, acc);
        }
      }
    }
    __syncthreads();
  }
}

int run_pair_generic(bool backward, const PairArgs& a, int n_cta, cudaStream_t st) {
  if (backward)
    k_pair_generic<true><<<n_cta, 256, 0, st>>>(a);
  else
    k_pair_generic<false><<<n_cta, 256, 0, st>>>(a);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

}  // namespace dvlp
