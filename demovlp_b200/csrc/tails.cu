// Loss tails, global cosine matrix and retrieval ranks.
//   RWALoss.forward          model/loss.py:108-114
//   NormSoftmaxLoss.forward  model/loss.py:126-138
//   sim_matrix               model/model.py:582-590
//   t2v_metrics / v2t_metrics rank search  model/metric.py:22-66, 142-189
#include <cooperative_groups.h>

#include "common.cuh"

namespace dvlp {

// ------------------------------------------------------------------ RWALoss
// One thread-block cluster of 8 CTAs (256 warps, one score row per warp at a time): rows are independent, the
// per-CTA partial sums meet in CTA 0 through distributed shared memory in a fixed order (deterministic, no
// scratch buffer, no atomics).  Rows of up to 1024 scores stay in registers between the passes.
constexpr int kRwaCluster = 8;
__global__ void __cluster_dims__(kRwaCluster, 1, 1) __launch_bounds__(1024)
k_rwa_loss(const float* __restrict__ scores, int n_rows, int n_cols, int row_offset, int n_total, float lambda,
           float* __restrict__ loss_sum, float* __restrict__ d_scores) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float part[32];
  __shared__ float cta_sum;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const float log_on = logf(1.f + 1e-6f), log_off = logf(1e-6f);   // torch.log(labels + 1e-6) in fp32
  const float sc = lambda / (float)n_total;
  float acc = 0.f;
  for (int i = blockIdx.x * nwarp + warp; i < n_rows; i += gridDim.x * nwarp) {
    const float* row = scores + (size_t)i * n_cols;
    const int tgt = i + row_offset;
    if (n_cols <= 1024) {
      float z[32];
      float m = -INFINITY;
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const int j = lane + 32 * t;
        z[t] = j < n_cols ? lambda * row[j] : -INFINITY;
        m = fmaxf(m, z[t]);
      }
      m = warp_max(m);
      float se = 0.f;
#pragma unroll
      for (int t = 0; t < 32; ++t) se += (lane + 32 * t < n_cols) ? expf(z[t] - m) : 0.f;
      se = warp_sum(se);
      const float lse = m + logf(se);
      float kl = 0.f;
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const int j = lane + 32 * t;
        if (j < n_cols) {
          const float logp = z[t] - lse;
          kl += expf(logp) * (logp - (j == tgt ? log_on : log_off));
        }
      }
      kl = warp_sum(kl);
      if (d_scores) {
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const int j = lane + 32 * t;
          if (j < n_cols) {
            const float logp = z[t] - lse;
            d_scores[(size_t)i * n_cols + j] = sc * expf(logp) * ((logp - (j == tgt ? log_on : log_off)) - kl);
          }
        }
      }
      acc += kl;   // identical in all lanes
    } else {
      float m = -INFINITY;
      for (int j = lane; j < n_cols; j += 32) m = fmaxf(m, lambda * row[j]);
      m = warp_max(m);
      float se = 0.f;
      for (int j = lane; j < n_cols; j += 32) se += expf(lambda * row[j] - m);
      se = warp_sum(se);
      const float lse = m + logf(se);
      float kl = 0.f;
      for (int j = lane; j < n_cols; j += 32) {
        const float logp = lambda * row[j] - lse;
        kl += expf(logp) * (logp - (j == tgt ? log_on : log_off));
      }
      kl = warp_sum(kl);
      if (d_scores) {
        for (int j = lane; j < n_cols; j += 32) {
          const float logp = lambda * row[j] - lse;
          d_scores[(size_t)i * n_cols + j] = sc * expf(logp) * ((logp - (j == tgt ? log_on : log_off)) - kl);
        }
      }
      acc += kl;
    }
  }
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < nwarp; ++w) t += part[w];
    cta_sum = t;
  }
  cluster.sync();
  if (cluster.block_rank() == 0 && threadIdx.x == 0) {
    float t = 0.f;
    for (unsigned r = 0; r < cluster.num_blocks(); ++r) t += *cluster.map_shared_rank(&cta_sum, r);
    *loss_sum = t / (float)n_total;
  }
  cluster.sync();      // remote shared memory must outlive the reads above
}

// ---------------------------------------------------------- NormSoftmaxLoss
__global__ void k_lse_rows(const float* __restrict__ x, int N, int M, float inv_t, float* __restrict__ lse) {
  int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= N) return;
  const float* row = x + (size_t)i * M;
  float m = -INFINITY;
  for (int j = lane; j < M; j += 32) m = fmaxf(m, row[j] * inv_t);
  m = warp_max(m);
  float se = 0.f;
  for (int j = lane; j < M; j += 32) se += expf(row[j] * inv_t - m);
  se = warp_sum(se);
  if (lane == 0) lse[i] = m + logf(se);
}
__global__ void k_lse_cols(const float* __restrict__ x, int N, int M, float inv_t, float* __restrict__ lse) {
  int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= M) return;
  float m = -INFINITY;
  for (int i = lane; i < N; i += 32) m = fmaxf(m, x[(size_t)i * M + j] * inv_t);
  m = warp_max(m);
  float se = 0.f;
  for (int i = lane; i < N; i += 32) se += expf(x[(size_t)i * M + j] * inv_t - m);
  se = warp_sum(se);
  if (lane == 0) lse[j] = m + logf(se);
}
__global__ void __launch_bounds__(1024) k_ns_loss(const float* __restrict__ x, int N, int M, float inv_t,
                                                  const float* __restrict__ lse_r, const float* __restrict__ lse_c,
                                                  float* __restrict__ loss) {
  __shared__ float part[32];
  const int n = N < M ? N : M;
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float z = x[(size_t)i * M + i] * inv_t;
    a += (z - lse_r[i]) + (z - lse_c[i]);
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    *loss = -t / (float)n;
  }
}
__global__ void k_ns_grad(const float* __restrict__ x, int N, int M, float inv_t, const float* __restrict__ lse_r,
                          const float* __restrict__ lse_c, float* __restrict__ d_x) {
  const int n = N < M ? N : M;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)N * M) return;
  int i = (int)(e / M), j = (int)(e - (size_t)i * M);
  float z = x[e] * inv_t;
  float diag = (i == j) ? 1.f : 0.f;
  float g = 0.f;
  if (i < n) g += expf(z - lse_r[i]) - diag;
  if (j < n) g += expf(z - lse_c[j]) - diag;
  d_x[e] = g * inv_t / (float)n;
}

// Small matrices (per-GPU batch: N, M <= 256): the four steps above in ONE launch.  A cluster of 8 CTAs computes
// the row / column log-sum-exps (one warp each), meets at a cluster barrier and then writes the loss (CTA 0) and
// the gradient (all CTAs); lse_r / lse_c travel through the caller's scratch buffer (L2).
constexpr int kNsCluster = 8;
__global__ void __cluster_dims__(kNsCluster, 1, 1) __launch_bounds__(1024)
k_ns_small(const float* __restrict__ x, int N, int M, float inv_t, float* lse, float* __restrict__ loss,
           float* __restrict__ d_x) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float part[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int idx = blockIdx.x * nwarp + warp; idx < N + M; idx += gridDim.x * nwarp) {
    float m = -INFINITY, se = 0.f;
    if (idx < N) {
      const float* row = x + (size_t)idx * M;
      for (int j = lane; j < M; j += 32) m = fmaxf(m, row[j] * inv_t);
      m = warp_max(m);
      for (int j = lane; j < M; j += 32) se += expf(row[j] * inv_t - m);
    } else {
      const int j = idx - N;
      for (int i = lane; i < N; i += 32) m = fmaxf(m, x[(size_t)i * M + j] * inv_t);
      m = warp_max(m);
      for (int i = lane; i < N; i += 32) se += expf(x[(size_t)i * M + j] * inv_t - m);
    }
    se = warp_sum(se);
    if (lane == 0) lse[idx] = m + logf(se);
  }
  __threadfence();
  cluster.sync();
  const float* lse_r = lse;
  const float* lse_c = lse + N;
  const int n = N < M ? N : M;
  if (blockIdx.x == 0) {
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float z = x[(size_t)i * M + i] * inv_t;
      a += (z - __ldcg(lse_r + i)) + (z - __ldcg(lse_c + i));
    }
    a = warp_sum(a);
    if (lane == 0) part[warp] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < nwarp; ++w) t += part[w];
      *loss = -t / (float)n;
    }
  }
  if (d_x) {
    const size_t total = (size_t)N * M;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
      const int i = (int)(e / M), j = (int)(e - (size_t)i * M);
      const float z = x[e] * inv_t;
      const float diag = (i == j) ? 1.f : 0.f;
      float g = 0.f;
      if (i < n) g += expf(z - __ldcg(lse_r + i)) - diag;
      if (j < n) g += expf(z - __ldcg(lse_c + j)) - diag;
      d_x[e] = g * inv_t / (float)n;
    }
  }
}

// --------------------------------------------------------------- sim_matrix
template <typename T>
__global__ void k_inv_norm_max(const T* __restrict__ x, int rows, int d, float eps, float* __restrict__ inv) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  float ss = 0.f;
  for (int k = lane; k < d; k += 32) {
    float v = to_f32(x[(size_t)row * d + k]);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) inv[row] = 1.f / fmaxf(sqrtf(ss), eps);
}

// out[i][j] = ia[i] ib[j] sum_k a[i][k] b[j][k];  32x32 output tile per CTA, k in chunks of 32
template <typename T>
__global__ void __launch_bounds__(1024) k_sim_fwd(const T* __restrict__ a, const T* __restrict__ b, int N, int M,
                                                  int d, const float* __restrict__ ia, const float* __restrict__ ib,
                                                  float* __restrict__ out) {
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  float acc = 0.f;
  for (int k0 = 0; k0 < d; k0 += 32) {
    int k = k0 + tx;
    sa[ty][tx] = (i0 + ty < N && k < d) ? to_f32(a[(size_t)(i0 + ty) * d + k]) : 0.f;
    sb[ty][tx] = (j0 + ty < M && k < d) ? to_f32(b[(size_t)(j0 + ty) * d + k]) : 0.f;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) acc = fmaf(sa[ty][kk], sb[tx][kk], acc);
    __syncthreads();
  }
  int i = i0 + ty, j = j0 + tx;
  if (i < N && j < M) out[(size_t)i * M + j] = acc * ia[i] * ib[j];
}

constexpr int kSimRB = 4;   // rows of `a` per CTA (small: the matrices are small, parallelism matters)

// d_a[i] = inv_i ( g_i - an_i (an_i . g_i) ),  g_i = sum_j w[i][j] bn_j,  w = d_out (TRANS: d_out^T)
// valid for |a_i| >= eps; a row shorter than eps is scaled by the constant 1/eps instead.
// The CTA's weights w[i][j] ib[j] are staged in shared memory so the j loop only streams rows of b
// (8 independent coalesced loads in flight per thread).
template <typename T, bool TRANS>
__global__ void __launch_bounds__(256) k_sim_bwd_side(const T* __restrict__ a, const T* __restrict__ b, int N, int M,
                                                      int d, float eps, const float* __restrict__ ia,
                                                      const float* __restrict__ ib, const float* __restrict__ d_out,
                                                      float* __restrict__ d_a) {
  // this CTA: RB rows of `a`, all d columns (thread k strides d)
  constexpr int RB = kSimRB;
  extern __shared__ float wsm[];          // [M][RB]
  __shared__ float red[RB][8];
  const int i0 = blockIdx.x * RB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < RB * M; e += blockDim.x) {
    const int r = e / M, j = e - r * M, i = i0 + r;
    float w = 0.f;
    if (i < N) w = (TRANS ? d_out[(size_t)j * N + i] : d_out[(size_t)i * M + j]) * ib[j];
    wsm[j * RB + r] = w;
  }
  __syncthreads();
  for (int kb = 0; kb < d; kb += blockDim.x) {
    int k = kb + threadIdx.x;
    float g[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) g[r] = 0.f;
    if (k < d) {
      int j = 0;
      for (; j + 8 <= M; j += 8) {
        float bv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) bv[u] = to_f32(b[(size_t)(j + u) * d + k]);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 w = *reinterpret_cast<const float4*>(wsm + (j + u) * RB);
          g[0] = fmaf(w.x, bv[u], g[0]); g[1] = fmaf(w.y, bv[u], g[1]);
          g[2] = fmaf(w.z, bv[u], g[2]); g[3] = fmaf(w.w, bv[u], g[3]);
        }
      }
      for (; j < M; ++j) {
        const float bvv = to_f32(b[(size_t)j * d + k]);
#pragma unroll
        for (int r = 0; r < RB; ++r) g[r] = fmaf(wsm[j * RB + r], bvv, g[r]);
      }
    }
    // store g temporarily in d_a, accumulate dot(an, g)
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      int i = i0 + r;
      float dot = 0.f;
      if (i < N && k < d) {
        d_a[(size_t)i * d + k] = g[r];
        dot = to_f32(a[(size_t)i * d + k]) * ia[i] * g[r];
      }
      dot = warp_sum(dot);
      if (lane == 0) {
        if (kb == 0) red[r][warp] = dot; else red[r][warp] += dot;
      }
    }
  }
  __syncthreads();
  for (int r = 0; r < RB; ++r) {
    int i = i0 + r;
    if (i >= N) continue;
    float dot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) dot += red[r][w];
    float inv = ia[i];
    bool clamped = inv >= 1.f / eps;   // |a| <= eps: a/eps is linear in a
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
      float gk = d_a[(size_t)i * d + k];
      float an = to_f32(a[(size_t)i * d + k]) * inv;
      d_a[(size_t)i * d + k] = clamped ? gk * inv : inv * (gk - an * dot);
    }
  }
}

// -------------------------------------------------------------------- ranks
// t2v: one warp per query row; count of entries strictly greater than the ground truth
__global__ void k_ranks_t2v(const double* __restrict__ sims, int Nq, int Nv, int32_t* __restrict__ ranks) {
  int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= Nq) return;
  const int per = Nq / Nv;
  const double* row = sims + (size_t)i * Nv;
  const double gt = row[i / per];
  int cnt = 0;
  for (int j = lane; j < Nv; j += 32) cnt += (row[j] > gt) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) ranks[i] = cnt;
}

// v2t: one CTA per video; each warp ranks one of the video's captions inside column v
__global__ void __launch_bounds__(256) k_ranks_v2t(const double* __restrict__ sims, int Nq, int Nv,
                                                   double* __restrict__ ranks) {
  __shared__ double best[8];
  const int v = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int per = Nq / Nv;
  double mine = INFINITY;
  for (int c = v * per + warp; c < (v + 1) * per; c += nwarp) {
    const double ref = sims[(size_t)c * Nv + v];
    int gr = 0, eq = 0;
    for (int q = lane; q < Nq; q += 32) {
      double x = sims[(size_t)q * Nv + v];
      gr += (x > ref) ? 1 : 0;
      eq += (x == ref) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      gr += __shfl_xor_sync(0xffffffffu, gr, o);
      eq += __shfl_xor_sync(0xffffffffu, eq, o);
    }
    // mean index of the tied block in the sorted row (metric.py:181-186): exact in float64
    double r = (double)gr + ((double)eq - 1.0) / 2.0;
    mine = fmin(mine, r);
  }
  if (lane == 0) best[warp] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = best[0];
    for (int w = 1; w < nwarp; ++w) b = fmin(b, best[w]);
    ranks[v] = b;
  }
}

// ------------------------------------------------------------ host wrappers
int run_rwa_loss(const float* scores, int n_rows, int n_cols, int row_offset, int n_total, float lambda,
                 float* loss_sum, float* d_scores, cudaStream_t st) {
  k_rwa_loss<<<kRwaCluster, 1024, 0, st>>>(scores, n_rows, n_cols, row_offset, n_total, lambda, loss_sum, d_scores);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

int run_normsoftmax(const float* x, int N, int M, float temperature, float* loss, float* d_x, float* scratch,
                    cudaStream_t st) {
  const float inv_t = 1.f / temperature;
  float* lr = scratch;
  float* lc = scratch + N;
  if (N <= 256 && M <= 256) {
    k_ns_small<<<kNsCluster, 1024, 0, st>>>(x, N, M, inv_t, scratch, loss, d_x);
    DVLP_LAUNCH_OK();
    return DVLP_OK;
  }
  k_lse_rows<<<(N + 7) / 8, 256, 0, st>>>(x, N, M, inv_t, lr);
  DVLP_LAUNCH_OK();
  k_lse_cols<<<(M + 7) / 8, 256, 0, st>>>(x, N, M, inv_t, lc);
  DVLP_LAUNCH_OK();
  k_ns_loss<<<1, 1024, 0, st>>>(x, N, M, inv_t, lr, lc, loss);
  DVLP_LAUNCH_OK();
  if (d_x) {
    size_t n = (size_t)N * M;
    k_ns_grad<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, N, M, inv_t, lr, lc, d_x);
    DVLP_LAUNCH_OK();
  }
  return DVLP_OK;
}

template <typename T>
static int sim_fwd_t(const T* a, const T* b, int N, int M, int d, float eps, float* out, float* scratch,
                     cudaStream_t st) {
  float* ia = scratch;
  float* ib = scratch + N;
  k_inv_norm_max<T><<<(N + 7) / 8, 256, 0, st>>>(a, N, d, eps, ia);
  DVLP_LAUNCH_OK();
  k_inv_norm_max<T><<<(M + 7) / 8, 256, 0, st>>>(b, M, d, eps, ib);
  DVLP_LAUNCH_OK();
  dim3 grid((M + 31) / 32, (N + 31) / 32);
  k_sim_fwd<T><<<grid, 1024, 0, st>>>(a, b, N, M, d, ia, ib, out);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}
template <typename T>
static int sim_bwd_t(const T* a, const T* b, int N, int M, int d, float eps, const float* d_out, float* d_a,
                     float* d_b, float* scratch, cudaStream_t st) {
  float* ia = scratch;
  float* ib = scratch + N;
  k_inv_norm_max<T><<<(N + 7) / 8, 256, 0, st>>>(a, N, d, eps, ia);
  DVLP_LAUNCH_OK();
  k_inv_norm_max<T><<<(M + 7) / 8, 256, 0, st>>>(b, M, d, eps, ib);
  DVLP_LAUNCH_OK();
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    DVLP_CUDA_OK(cudaFuncSetAttribute(k_sim_bwd_side<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DVLP_CUDA_OK(cudaFuncSetAttribute(k_sim_bwd_side<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_once.done();
  }
  if ((size_t)(N > M ? N : M) * kSimRB * 4 > 200 * 1024) return DVLP_ERR_BAD_SHAPE;
  if (d_a) {
    k_sim_bwd_side<T, false><<<(N + kSimRB - 1) / kSimRB, 256, (size_t)M * kSimRB * 4, st>>>(a, b, N, M, d, eps, ia, ib, d_out, d_a);
    DVLP_LAUNCH_OK();
  }
  if (d_b) {
    k_sim_bwd_side<T, true><<<(M + kSimRB - 1) / kSimRB, 256, (size_t)N * kSimRB * 4, st>>>(b, a, M, N, d, eps, ib, ia, d_out, d_b);
    DVLP_LAUNCH_OK();
  }
  return DVLP_OK;
}

int run_sim_fwd_mma(const void* a, const void* b, int N, int M, float eps, float* out, cudaStream_t st);
int run_sim_bwd_mma(const void* a, const void* b, int N, int M, float eps, const float* d_out, void* d_a, void* d_b,
                    int out_dtype, float* scratch, cudaStream_t st);

int run_sim_fwd(const void* a, const void* b, int N, int M, int d, float eps, int dt, float* out, float* scratch,
                cudaStream_t st) {
  if (dt == DVLP_BF16 && d == 256) return run_sim_fwd_mma(a, b, N, M, eps, out, st);
  if (dt == DVLP_BF16)
    return sim_fwd_t((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, N, M, d, eps, out, scratch, st);
  return sim_fwd_t((const float*)a, (const float*)b, N, M, d, eps, out, scratch, st);
}
int run_sim_bwd(const void* a, const void* b, int N, int M, int d, float eps, int dt, const float* d_out, void* d_a,
                void* d_b, int out_dtype, float* scratch, cudaStream_t st) {
  if (dt == DVLP_BF16 && d == 256) return run_sim_bwd_mma(a, b, N, M, eps, d_out, d_a, d_b, out_dtype, scratch, st);
  if (out_dtype != DVLP_F32) return DVLP_ERR_BAD_DTYPE;      // the fp32 CUDA-core kernels write fp32 only
  if (dt == DVLP_BF16)
    return sim_bwd_t((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, N, M, d, eps, d_out, (float*)d_a, (float*)d_b,
                     scratch, st);
  return sim_bwd_t((const float*)a, (const float*)b, N, M, d, eps, d_out, (float*)d_a, (float*)d_b, scratch, st);
}

int run_ranks_t2v(const double* sims, int Nq, int Nv, int32_t* ranks, cudaStream_t st) {
  k_ranks_t2v<<<(Nq + 7) / 8, 256, 0, st>>>(sims, Nq, Nv, ranks);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}
int run_ranks_v2t(const double* sims, int Nq, int Nv, double* ranks, cudaStream_t st) {
  k_ranks_v2t<<<Nv, 256, 0, st>>>(sims, Nq, Nv, ranks);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

}  // namespace dvlp
