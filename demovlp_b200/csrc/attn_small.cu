// Masked attention over the SHORT token sequences of the ObjectTransformer: the reference's
//     attn_mask(q, k, v, mask) = softmax(q k^T + mask, dim=-1) v            model/object_transformer.py:93-100
// is called by VarAttention.forward (:125-196) on (batch x heads) sequences of 1 + R region tokens per frame (space
// attention: 30 queries x 31 keys), 1 + F frame tokens per region (time attention: 8 x 9) and, for the CLS token, ONE query
// against all 1 + F R keys.  Each call is two batched fp32 GEMMs, an add and a softmax that round-trip (b, i, j) tensors
// through HBM; here one CTA handles one sequence end to end and touches q, k, v, out exactly once.
//
//   k_attn_mma_fwd / _bwd   up to 64 queries x 64 keys, head dimension 64: warp = 16 query rows, products on
//       mma.sync.m16n8k16 with every fp32 operand split into bf16 hi + lo parts (a b ~ ah bh + ah bl + al bh: fp32-class
//       accuracy, relative error ~ 2^-17 -- the reference trains in fp32 without AMP, so plain bf16 / tf32 operands would
//       not be a drop-in); softmax on the accumulator fragments (quad shuffles); P and dS go back into the tensor cores
//       as A fragments built from the accumulators (no shared-memory round trip) and, for dK / dV, transposed through a
//       small shared-memory tile (ldmatrix.trans), so the backward needs no atomics.
//   k_attn_cls_fwd / _bwd   one query, up to 1024 keys (the CLS token): warp per key stripe, lanes over the head
//       dimension, coalesced row reads, partial softmax states merged through shared memory.
#include "common.cuh"
#include "mma_sync.cuh"

namespace dvlp {
namespace at {
constexpr int D = 64, LD = (D + 8) * 2;          // bytes per staged bf16 row: conflict-free ldmatrix
constexpr int kMaxT = 4;                          // 16-row / 16-key blocks per sequence (64 tokens)

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - __low2float(h), b - __high2float(h));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// A fragment of A[m][k] = X[k][m] (X row-major in shared memory, rows k0.., columns m0..): transposed load
__device__ __forceinline__ void load_a_kn(uint32_t base, int ld, int k0, int m0, int lane, uint32_t (&a)[4]) {
  const int r = k0 + (lane & 7) + (lane >> 4) * 8, c = m0 + ((lane >> 3) & 1) * 8;
  mm::ldmatrix_x4_trans(base + r * ld + c * 2, a[0], a[1], a[2], a[3]);
}
// stage `rows` fp32 rows of 64 (zero beyond `valid`) as bf16 hi / lo planes [rows][LD bytes]
__device__ __forceinline__ void stage_rows(const float* __restrict__ src, int valid, int rows, uint8_t* hi, uint8_t* lo,
                                           int tid, int nthr) {
  for (int e = tid; e < rows * (D / 2); e += nthr) {
    const int r = e >> 5, c = (e & 31) * 2;
    float2 v = make_float2(0.f, 0.f);
    if (r < valid) v = *reinterpret_cast<const float2*>(src + (size_t)r * D + c);
    uint32_t h, l;
    split2(v.x, v.y, h, l);
    *reinterpret_cast<uint32_t*>(hi + r * LD + c * 2) = h;
    *reinterpret_cast<uint32_t*>(lo + r * LD + c * 2) = l;
  }
}
// acc (16 x 8 fp32) += A B with hi / lo splits of both operands (lo x lo dropped)
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                     uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mm::mma_16816(c, ah[0], ah[1], ah[2], ah[3], bh0, bh1);
  mm::mma_16816(c, ah[0], ah[1], ah[2], ah[3], bl0, bl1);
  mm::mma_16816(c, al[0], al[1], al[2], al[3], bh0, bh1);
}

struct Smem {          // byte offsets inside the dynamic shared memory of one sequence
  int qh, ql, kh, kl, vh, vl, gh, gl, ph, pl, sh, sl, mask, total, ldp;
};
__host__ __device__ inline Smem smem_layout(int IP, int JP, bool bwd) {
  Smem s{};
  int off = 0;
  auto take = [&](int bytes) { int o = off; off += (bytes + 15) & ~15; return o; };
  s.qh = take(IP * LD); s.ql = take(IP * LD);
  s.kh = take(JP * LD); s.kl = take(JP * LD);
  s.vh = take(JP * LD); s.vl = take(JP * LD);
  s.ldp = (JP + 8) * 2;
  if (bwd) {
    s.gh = take(IP * LD); s.gl = take(IP * LD);
    s.ph = take(IP * s.ldp); s.pl = take(IP * s.ldp);
    s.sh = take(IP * s.ldp); s.sl = take(IP * s.ldp);
  }
  s.mask = take(JP * 4);
  s.total = off;
  return s;
}

// S = Q K^T + mask for this warp's 16 query rows: acc[t] = n8 tile t (keys 8t .. 8t+7)
__device__ __forceinline__ void scores16(uint32_t sb, const Smem& L, int row0, int JB, int lane, const float* msk,
                                         float (&acc)[2 * kMaxT][4]) {
#pragma unroll
  for (int t = 0; t < 2 * kMaxT; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
#pragma unroll
  for (int k0 = 0; k0 < D; k0 += 16) {
    uint32_t ah[4], al[4];
    mm::load_a(sb + L.qh, LD, row0, k0, lane, ah);
    mm::load_a(sb + L.ql, LD, row0, k0, lane, al);
#pragma unroll
    for (int jb = 0; jb < kMaxT; ++jb) {
      if (jb < JB) {
        uint32_t bh[4], bl[4];
        mm::load_b_nk(sb + L.kh, LD, jb * 16, k0, lane, bh);
        mm::load_b_nk(sb + L.kl, LD, jb * 16, k0, lane, bl);
        mma3(acc[2 * jb], ah, al, bh[0], bh[1], bl[0], bl[1]);
        mma3(acc[2 * jb + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
      }
    }
  }
  const int tq = lane & 3;
#pragma unroll
  for (int t = 0; t < 2 * kMaxT; ++t) {
    if (t < 2 * JB) {
      const float m0 = msk[t * 8 + tq * 2], m1 = msk[t * 8 + tq * 2 + 1];
      acc[t][0] += m0; acc[t][1] += m1; acc[t][2] += m0; acc[t][3] += m1;
    }
  }
}

// one CTA per sequence, warp w = query rows 16w .. 16w+15
__global__ void __launch_bounds__(128) k_attn_mma_fwd(const float* __restrict__ q, const float* __restrict__ k,
                                                      const float* __restrict__ v, const float* __restrict__ mask, int I,
                                                      int J, float* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int IB = (I + 15) >> 4, JB = (J + 15) >> 4, IP = IB * 16, JP = JB * 16;
  const Smem L = smem_layout(IP, JP, false);
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nthr = blockDim.x;
  stage_rows(q + (size_t)b * I * D, I, IP, sm + L.qh, sm + L.ql, threadIdx.x, nthr);
  stage_rows(k + (size_t)b * J * D, J, JP, sm + L.kh, sm + L.kl, threadIdx.x, nthr);
  stage_rows(v + (size_t)b * J * D, J, JP, sm + L.vh, sm + L.vl, threadIdx.x, nthr);
  float* msk = reinterpret_cast<float*>(sm + L.mask);
  for (int j = threadIdx.x; j < JP; j += nthr) msk[j] = j < J ? (mask ? mask[(size_t)b * J + j] : 0.f) : -INFINITY;
  __syncthreads();
  if (warp >= IB) return;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sm);
  const int row0 = warp * 16, g = lane >> 2, tq = lane & 3;
  float acc[2 * kMaxT][4];
  scores16(sb, L, row0, JB, lane, msk, acc);
  // softmax over the keys: rows g (elements 0, 1) and g + 8 (elements 2, 3) of every tile
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int t = 0; t < 2 * kMaxT; ++t)
    if (t < 2 * JB) { mx0 = fmaxf(mx0, fmaxf(acc[t][0], acc[t][1])); mx1 = fmaxf(mx1, fmaxf(acc[t][2], acc[t][3])); }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int t = 0; t < 2 * kMaxT; ++t)
    if (t < 2 * JB) {
      acc[t][0] = __expf(acc[t][0] - mx0); acc[t][1] = __expf(acc[t][1] - mx0);
      acc[t][2] = __expf(acc[t][2] - mx1); acc[t][3] = __expf(acc[t][3] - mx1);
      s0 += acc[t][0] + acc[t][1]; s1 += acc[t][2] + acc[t][3];
    }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  // O = P V: P as A fragments straight from the accumulators
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int jb = 0; jb < kMaxT; ++jb) {
    if (jb < JB) {
      uint32_t ah[4], al[4];
      split2(acc[2 * jb][0], acc[2 * jb][1], ah[0], al[0]);
      split2(acc[2 * jb][2], acc[2 * jb][3], ah[1], al[1]);
      split2(acc[2 * jb + 1][0], acc[2 * jb + 1][1], ah[2], al[2]);
      split2(acc[2 * jb + 1][2], acc[2 * jb + 1][3], ah[3], al[3]);
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        uint32_t bh[4], bl[4];
        mm::load_b_kn(sb + L.vh, LD, jb * 16, nb * 16, lane, bh);
        mm::load_b_kn(sb + L.vl, LD, jb * 16, nb * 16, lane, bl);
        mma3(o[2 * nb], ah, al, bh[0], bh[1], bl[0], bl[1]);
        mma3(o[2 * nb + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
      }
    }
  }
  const float i0 = 1.f / s0, i1 = 1.f / s1;
  const int r0 = row0 + g, r1 = row0 + g + 8;
  float* ob = out + (size_t)b * I * D;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int c = n * 8 + tq * 2;
    if (r0 < I) *reinterpret_cast<float2*>(ob + (size_t)r0 * D + c) = make_float2(o[n][0] * i0, o[n][1] * i0);
    if (r1 < I) *reinterpret_cast<float2*>(ob + (size_t)r1 * D + c) = make_float2(o[n][2] * i1, o[n][3] * i1);
  }
  if (lse && tq == 0) {
    if (r0 < I) lse[(size_t)b * I + r0] = mx0 + __logf(s0);
    if (r1 < I) lse[(size_t)b * I + r1] = mx1 + __logf(s1);
  }
}

// backward: P = exp(S - lse), dP = dO V^T, dS = P (dP - delta), dQ = dS K;  then, warps over KEY blocks:
// dV = P^T dO, dK = dS^T Q  (P / dS transposed through shared memory)
__global__ void __launch_bounds__(128) k_attn_mma_bwd(const float* __restrict__ q, const float* __restrict__ k,
                                                      const float* __restrict__ v, const float* __restrict__ mask,
                                                      const float* __restrict__ out, const float* __restrict__ lse,
                                                      const float* __restrict__ d_out, int I, int J,
                                                      float* __restrict__ d_q, float* __restrict__ d_k,
                                                      float* __restrict__ d_v) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int IB = (I + 15) >> 4, JB = (J + 15) >> 4, IP = IB * 16, JP = JB * 16;
  const Smem L = smem_layout(IP, JP, true);
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nthr = blockDim.x, nwarp = nthr >> 5;
  stage_rows(q + (size_t)b * I * D, I, IP, sm + L.qh, sm + L.ql, threadIdx.x, nthr);
  stage_rows(k + (size_t)b * J * D, J, JP, sm + L.kh, sm + L.kl, threadIdx.x, nthr);
  stage_rows(v + (size_t)b * J * D, J, JP, sm + L.vh, sm + L.vl, threadIdx.x, nthr);
  stage_rows(d_out + (size_t)b * I * D, I, IP, sm + L.gh, sm + L.gl, threadIdx.x, nthr);
  float* msk = reinterpret_cast<float*>(sm + L.mask);
  for (int j = threadIdx.x; j < JP; j += nthr) msk[j] = j < J ? (mask ? mask[(size_t)b * J + j] : 0.f) : -INFINITY;
  __syncthreads();
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sm);
  const int g = lane >> 2, tq = lane & 3;
  if (warp < IB) {
    const int row0 = warp * 16, r0 = row0 + g, r1 = row0 + g + 8;
    float acc[2 * kMaxT][4];
    scores16(sb, L, row0, JB, lane, msk, acc);
    // per-row log-sum-exp and delta = dO . O (fp32 from global memory: 16 columns per thread of the quad)
    const float l0 = r0 < I ? lse[(size_t)b * I + r0] : 0.f, l1 = r1 < I ? lse[(size_t)b * I + r1] : 0.f;
    float dl0 = 0.f, dl1 = 0.f;
    {
      const float* ob = out + (size_t)b * I * D;
      const float* gb = d_out + (size_t)b * I * D;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const int c = tq * 16 + c4 * 4;
        if (r0 < I) {
          const float4 a = *reinterpret_cast<const float4*>(ob + (size_t)r0 * D + c), gg = *reinterpret_cast<const float4*>(gb + (size_t)r0 * D + c);
          dl0 += a.x * gg.x + a.y * gg.y + a.z * gg.z + a.w * gg.w;
        }
        if (r1 < I) {
          const float4 a = *reinterpret_cast<const float4*>(ob + (size_t)r1 * D + c), gg = *reinterpret_cast<const float4*>(gb + (size_t)r1 * D + c);
          dl1 += a.x * gg.x + a.y * gg.y + a.z * gg.z + a.w * gg.w;
        }
      }
      dl0 += __shfl_xor_sync(0xffffffffu, dl0, 1); dl0 += __shfl_xor_sync(0xffffffffu, dl0, 2);
      dl1 += __shfl_xor_sync(0xffffffffu, dl1, 1); dl1 += __shfl_xor_sync(0xffffffffu, dl1, 2);
    }
    // dP = dO V^T
    float dp[2 * kMaxT][4];
#pragma unroll
    for (int t = 0; t < 2 * kMaxT; ++t) dp[t][0] = dp[t][1] = dp[t][2] = dp[t][3] = 0.f;
#pragma unroll
    for (int k0 = 0; k0 < D; k0 += 16) {
      uint32_t ah[4], al[4];
      mm::load_a(sb + L.gh, LD, row0, k0, lane, ah);
      mm::load_a(sb + L.gl, LD, row0, k0, lane, al);
#pragma unroll
      for (int jb = 0; jb < kMaxT; ++jb) {
        if (jb < JB) {
          uint32_t bh[4], bl[4];
          mm::load_b_nk(sb + L.vh, LD, jb * 16, k0, lane, bh);
          mm::load_b_nk(sb + L.vl, LD, jb * 16, k0, lane, bl);
          mma3(dp[2 * jb], ah, al, bh[0], bh[1], bl[0], bl[1]);
          mma3(dp[2 * jb + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
        }
      }
    }
    // P and dS (in place: acc = P, dp = dS); both go to shared memory (bf16 hi / lo) for the key-side products
#pragma unroll
    for (int t = 0; t < 2 * kMaxT; ++t) {
      if (t < 2 * JB) {
        acc[t][0] = __expf(acc[t][0] - l0); acc[t][1] = __expf(acc[t][1] - l0);
        acc[t][2] = __expf(acc[t][2] - l1); acc[t][3] = __expf(acc[t][3] - l1);
        if (r0 >= I) acc[t][0] = acc[t][1] = 0.f;             // padded query rows contribute nothing to dK / dV
        if (r1 >= I) acc[t][2] = acc[t][3] = 0.f;
        dp[t][0] = acc[t][0] * (dp[t][0] - dl0); dp[t][1] = acc[t][1] * (dp[t][1] - dl0);
        dp[t][2] = acc[t][2] * (dp[t][2] - dl1); dp[t][3] = acc[t][3] * (dp[t][3] - dl1);
        const int c = t * 8 + tq * 2;
        uint32_t h, l;
        split2(acc[t][0], acc[t][1], h, l);
        *reinterpret_cast<uint32_t*>(sm + L.ph + r0 * L.ldp + c * 2) = h; *reinterpret_cast<uint32_t*>(sm + L.pl + r0 * L.ldp + c * 2) = l;
        split2(acc[t][2], acc[t][3], h, l);
        *reinterpret_cast<uint32_t*>(sm + L.ph + r1 * L.ldp + c * 2) = h; *reinterpret_cast<uint32_t*>(sm + L.pl + r1 * L.ldp + c * 2) = l;
        split2(dp[t][0], dp[t][1], h, l);
        *reinterpret_cast<uint32_t*>(sm + L.sh + r0 * L.ldp + c * 2) = h; *reinterpret_cast<uint32_t*>(sm + L.sl + r0 * L.ldp + c * 2) = l;
        split2(dp[t][2], dp[t][3], h, l);
        *reinterpret_cast<uint32_t*>(sm + L.sh + r1 * L.ldp + c * 2) = h; *reinterpret_cast<uint32_t*>(sm + L.sl + r1 * L.ldp + c * 2) = l;
      }
    }
    // dQ = dS K
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
    for (int jb = 0; jb < kMaxT; ++jb) {
      if (jb < JB) {
        uint32_t ah[4], al[4];
        split2(dp[2 * jb][0], dp[2 * jb][1], ah[0], al[0]);
        split2(dp[2 * jb][2], dp[2 * jb][3], ah[1], al[1]);
        split2(dp[2 * jb + 1][0], dp[2 * jb + 1][1], ah[2], al[2]);
        split2(dp[2 * jb + 1][2], dp[2 * jb + 1][3], ah[3], al[3]);
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
          uint32_t bh[4], bl[4];
          mm::load_b_kn(sb + L.kh, LD, jb * 16, nb * 16, lane, bh);
          mm::load_b_kn(sb + L.kl, LD, jb * 16, nb * 16, lane, bl);
          mma3(o[2 * nb], ah, al, bh[0], bh[1], bl[0], bl[1]);
          mma3(o[2 * nb + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
        }
      }
    }
    float* qb = d_q + (size_t)b * I * D;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int c = n * 8 + tq * 2;
      if (r0 < I) *reinterpret_cast<float2*>(qb + (size_t)r0 * D + c) = make_float2(o[n][0], o[n][1]);
      if (r1 < I) *reinterpret_cast<float2*>(qb + (size_t)r1 * D + c) = make_float2(o[n][2], o[n][3]);
    }
  }
  __syncthreads();
  // key side: warp = 16 keys;  dV = P^T dO,  dK = dS^T Q  (sum over the query rows = the K dimension of the products)
  for (int jb = warp; jb < JB; jb += nwarp) {
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
#pragma unroll
    for (int ib = 0; ib < kMaxT; ++ib) {
      if (ib < IB) {
        uint32_t ph[4], pl[4], sh[4], sl[4];
        load_a_kn(sb + L.ph, L.ldp, ib * 16, jb * 16, lane, ph);
        load_a_kn(sb + L.pl, L.ldp, ib * 16, jb * 16, lane, pl);
        load_a_kn(sb + L.sh, L.ldp, ib * 16, jb * 16, lane, sh);
        load_a_kn(sb + L.sl, L.ldp, ib * 16, jb * 16, lane, sl);
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
          uint32_t bh[4], bl[4];
          mm::load_b_kn(sb + L.gh, LD, ib * 16, nb * 16, lane, bh);
          mm::load_b_kn(sb + L.gl, LD, ib * 16, nb * 16, lane, bl);
          mma3(dv[2 * nb], ph, pl, bh[0], bh[1], bl[0], bl[1]);
          mma3(dv[2 * nb + 1], ph, pl, bh[2], bh[3], bl[2], bl[3]);
          mm::load_b_kn(sb + L.qh, LD, ib * 16, nb * 16, lane, bh);
          mm::load_b_kn(sb + L.ql, LD, ib * 16, nb * 16, lane, bl);
          mma3(dk[2 * nb], sh, sl, bh[0], bh[1], bl[0], bl[1]);
          mma3(dk[2 * nb + 1], sh, sl, bh[2], bh[3], bl[2], bl[3]);
        }
      }
    }
    const int j0 = jb * 16 + g, j1 = jb * 16 + g + 8;
    float* kb = d_k + (size_t)b * J * D;
    float* vb = d_v + (size_t)b * J * D;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int c = n * 8 + tq * 2;
      if (j0 < J) {
        *reinterpret_cast<float2*>(kb + (size_t)j0 * D + c) = make_float2(dk[n][0], dk[n][1]);
        *reinterpret_cast<float2*>(vb + (size_t)j0 * D + c) = make_float2(dv[n][0], dv[n][1]);
      }
      if (j1 < J) {
        *reinterpret_cast<float2*>(kb + (size_t)j1 * D + c) = make_float2(dk[n][2], dk[n][3]);
        *reinterpret_cast<float2*>(vb + (size_t)j1 * D + c) = make_float2(dv[n][2], dv[n][3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------- the CLS token: one query
constexpr int kClsWarps = 4;
// each warp takes keys w, w + 4, ...: lanes over the head dimension (2 elements), an online softmax state per warp
__global__ void __launch_bounds__(32 * kClsWarps) k_attn_cls_fwd(const float* __restrict__ q, const float* __restrict__ k,
                                                                const float* __restrict__ v, const float* __restrict__ mask,
                                                                int J, float* __restrict__ out, float* __restrict__ lse) {
  __shared__ float s_m[kClsWarps], s_l[kClsWarps], s_o[kClsWarps][D];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float q0 = q[(size_t)b * D + lane], q1 = q[(size_t)b * D + lane + 32];
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int j = warp; j < J; j += kClsWarps) {
    const float* kr = k + ((size_t)b * J + j) * D;
    const float* vr = v + ((size_t)b * J + j) * D;
    float s = warp_sum(q0 * kr[lane] + q1 * kr[lane + 32]) + (mask ? mask[(size_t)b * J + j] : 0.f);
    const float mn = fmaxf(m, s);
    const float corr = __expf(m - mn), p = __expf(s - mn);
    l = l * corr + p;
    o0 = o0 * corr + p * vr[lane];
    o1 = o1 * corr + p * vr[lane + 32];
    m = mn;
  }
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
  s_o[warp][lane] = o0; s_o[warp][lane + 32] = o1;
  __syncthreads();
  if (warp == 0) {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kClsWarps; ++w) M = fmaxf(M, s_m[w]);
    float Ls = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int w = 0; w < kClsWarps; ++w) {
      const float c = s_m[w] == -INFINITY ? 0.f : __expf(s_m[w] - M);
      Ls += s_l[w] * c; a0 += s_o[w][lane] * c; a1 += s_o[w][lane + 32] * c;
    }
    out[(size_t)b * D + lane] = a0 / Ls;
    out[(size_t)b * D + lane + 32] = a1 / Ls;
    if (lane == 0 && lse) lse[b] = M + __logf(Ls);
  }
}

__global__ void __launch_bounds__(32 * kClsWarps) k_attn_cls_bwd(const float* __restrict__ q, const float* __restrict__ k,
                                                                const float* __restrict__ v, const float* __restrict__ mask,
                                                                const float* __restrict__ out, const float* __restrict__ lse,
                                                                const float* __restrict__ d_out, int J, float* __restrict__ d_q,
                                                                float* __restrict__ d_k, float* __restrict__ d_v) {
  __shared__ float s_dq[kClsWarps][D];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float q0 = q[(size_t)b * D + lane], q1 = q[(size_t)b * D + lane + 32];
  const float g0 = d_out[(size_t)b * D + lane], g1 = d_out[(size_t)b * D + lane + 32];
  const float delta = warp_sum(g0 * out[(size_t)b * D + lane] + g1 * out[(size_t)b * D + lane + 32]);
  const float Ls = lse[b];
  float dq0 = 0.f, dq1 = 0.f;
  for (int j = warp; j < J; j += kClsWarps) {
    const size_t r = ((size_t)b * J + j) * D;
    const float k0 = k[r + lane], k1 = k[r + lane + 32], v0 = v[r + lane], v1 = v[r + lane + 32];
    const float s = warp_sum(q0 * k0 + q1 * k1) + (mask ? mask[(size_t)b * J + j] : 0.f);
    const float p = __expf(s - Ls);
    const float ds = p * (warp_sum(g0 * v0 + g1 * v1) - delta);
    d_v[r + lane] = p * g0; d_v[r + lane + 32] = p * g1;
    d_k[r + lane] = ds * q0; d_k[r + lane + 32] = ds * q1;
    dq0 = fmaf(ds, k0, dq0); dq1 = fmaf(ds, k1, dq1);
  }
  s_dq[warp][lane] = dq0; s_dq[warp][lane + 32] = dq1;
  __syncthreads();
  if (warp == 0) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int w = 0; w < kClsWarps; ++w) { a0 += s_dq[w][lane]; a1 += s_dq[w][lane + 32]; }
    d_q[(size_t)b * D + lane] = a0;
    d_q[(size_t)b * D + lane + 32] = a1;
  }
}
}  // namespace at

// 1: tensor-core kernels (I, J <= 64), 2: single-query kernels (I == 1, any J <= 1024), 0: not served
int attn_small_supported(int I, int J, int d) {
  if (d != at::D || I < 1 || J < 1) return 0;
  if (I == 1 && J <= 1024) return 2;
  if (I <= 64 && J <= 64) return 1;
  return 0;
}

int run_attn_small_fwd(const float* q, const float* k, const float* v, const float* mask, int B, int I, int J,
                       float* out, float* lse, cudaStream_t st) {
  if (attn_small_supported(I, J, at::D) == 2) {
    at::k_attn_cls_fwd<<<B, 32 * at::kClsWarps, 0, st>>>(q, k, v, mask, J, out, lse);
    DVLP_LAUNCH_OK();
    return DVLP_OK;
  }
  const int IB = (I + 15) / 16, JB = (J + 15) / 16;
  const at::Smem L = at::smem_layout(IB * 16, JB * 16, false);
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    DVLP_CUDA_OK(cudaFuncSetAttribute(at::k_attn_mma_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_once.done();
  }
  at::k_attn_mma_fwd<<<B, 32 * IB, L.total, st>>>(q, k, v, mask, I, J, out, lse);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

int run_attn_small_bwd(const float* q, const float* k, const float* v, const float* mask, const float* out,
                       const float* lse, const float* d_out, int B, int I, int J, float* d_q, float* d_k, float* d_v,
                       cudaStream_t st) {
  if (attn_small_supported(I, J, at::D) == 2) {
    at::k_attn_cls_bwd<<<B, 32 * at::kClsWarps, 0, st>>>(q, k, v, mask, out, lse, d_out, J, d_q, d_k, d_v);
    DVLP_LAUNCH_OK();
    return DVLP_OK;
  }
  const int IB = (I + 15) / 16, JB = (J + 15) / 16;
  const at::Smem L = at::smem_layout(IB * 16, JB * 16, true);
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    DVLP_CUDA_OK(cudaFuncSetAttribute(at::k_attn_mma_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    attr_once.done();
  }
  const int nw = IB > JB ? IB : JB;
  at::k_attn_mma_bwd<<<B, 32 * nw, L.total, st>>>(q, k, v, mask, out, lse, d_out, I, J, d_q, d_k, d_v);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

}  // namespace dvlp
