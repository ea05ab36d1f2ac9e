// Shared device/host helpers for the demovlp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/demovlp_b200.h"

namespace dvlp {

constexpr float kEps = 1e-8f;          // model/loss.py:333 l2norm eps
constexpr float kLeaky = 0.1f;         // model/loss.py:236 LeakyReLU slope
constexpr float kLog2e = 1.4426950408889634f;

extern thread_local int g_last_cuda_error;
void count_launch(int n = 1);

#define DVLP_CUDA_OK(expr)                                   \
  do {                                                       \
    cudaError_t e__ = (expr);                                \
    if (e__ != cudaSuccess) {                                \
      dvlp::g_last_cuda_error = (int)e__;                    \
      return DVLP_ERR_CUDA;                                  \
    }                                                        \
  } while (0)

#define DVLP_LAUNCH_OK()                                     \
  do {                                                       \
    dvlp::count_launch();                                    \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) {                                \
      dvlp::g_last_cuda_error = (int)e__;                    \
      return DVLP_ERR_CUDA;                                  \
    }                                                        \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }

__device__ __forceinline__ float leaky(float x) { return fmaxf(x, kLeaky * x); }
__device__ __forceinline__ float leaky_grad(float x) { return x > 0.f ? 1.f : kLeaky; }

// Packed fp32 pairs (sm_100: FFMA2 / FMUL2 / FADD2 do two IEEE fp32 operations per issue slot; the results are
// bit-identical to the scalar instructions).  The pair kernels are issue-bound on their fp32 elementwise work, so
// adjacent columns of a thread are processed as pairs wherever both need the same operation.
__device__ __forceinline__ unsigned long long f2_pack(float2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 f2_unpack(unsigned long long r) {
  float2 a;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
  return a;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(r);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(r);
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
  return f2_unpack(r);
}
__device__ __forceinline__ float2 f2(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 leaky2(float2 x) {
  const float2 t = fmul2(x, f2(kLeaky));
  return make_float2(fmaxf(x.x, t.x), fmaxf(x.y, t.y));
}
// 1.0f where a > b, else 0.0f (one FSET; a select against zero becomes a multiply by this mask, which
// moves half of its cost from the half-rate ALU pipe to a packed multiply)
__device__ __forceinline__ float gt_mask(float a, float b) {
  float r;
  asm("set.gt.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float2 gt_mask2(float2 a, float2 b) { return make_float2(gt_mask(a.x, b.x), gt_mask(a.y, b.y)); }

// 1 / (sqrt(x) + eps) (the l2norm of model/loss.py:333) on the two MUFU approximations (2 ulp each) instead of the IEEE
// square root and division with their Newton steps and range branches: in the pair kernels both sit on the per-tile
// dependency chain of every warp (forward at config 3: -3.4 %)
__device__ __forceinline__ float inv_norm(float x) {
  float s, r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s + kEps));
  return r;
}

// Optional per-kernel timing (bench.py's roofline): CUDA events recorded on the launching stream
// around the kernels of one slot.  Off by default; costs nothing when off.
enum ProfSlot { kProfFwd = 0, kProfBwd = 1, kProfPrep = 2, kProfFinal = 3, kProfGG = 4, kProfSlots = 5 };
struct ProfScope {
  int slot;
  cudaStream_t st;
  void* rec;
  ProfScope(int slot, cudaStream_t st);
  ~ProfScope();
};

// cudaFuncSetAttribute (and the sm_100 check) are PER DEVICE: remembers on which devices of this process a
// one-time setup has run (device ids < 64; beyond that the setup simply repeats, which is harmless).
struct DeviceOnce {
  std::atomic<unsigned long long> mask{0};
  bool pending() const {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    return !((mask.load(std::memory_order_acquire) >> dev) & 1ull);
  }
  void done() {
    int dev = -1;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) mask.fetch_or(1ull << dev, std::memory_order_release);
  }
};

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------
// Workspace layout shared by forward and backward (all offsets 256-B aligned).
// ---------------------------------------------------------------------------
struct Workspace {
  // per-item prologue
  size_t alpha, beta;        // fp32 inverse row norms 1/(|x|+eps): (Bv*R), (Bt*W)
  size_t mterm_v, mterm_t;   // fp32 lambda*(mask - max mask) per region / word
  size_t vh, th;             // fp32 normalised rows
  size_t gv, gt;             // fp32 Gram matrices (Bv,R,R) / (Bt,W,W)
  // tcgen05 extras (zero-sized otherwise)
  size_t vb, tb;             // raw rows rounded to bf16 (only when the input is fp32)
  size_t gvb, gtb;           // bf16 Gram matrices
  size_t m2v, m2t;           // (mterm - lambda) * log2e
  // backward accumulators
  size_t dvh, dth;           // fp32 (Bv,R,d), (Bt,W,d)
  size_t dgv, dgt;           // fp32 (Bv,R,R), (Bt,W,W)
  size_t pa;                 // tcgen05 backward: bf16 gradient tiles P (Bv * ceil(Bt/4), 128 words, R) spilled for grad_gemm2.cu
  size_t rowv, colv;         // tcgen05: per-unit query scalars saved by the forward for pair_bwd2.cu: float4 (units,128) =
                             // (nB, thr1, 1/|c1|, cos1) per word, float4 (units,4,R) = (nA, thr2, 1/|c2|, cos2) per (text, region)
  size_t gbd;                // tcgen05: block-diagonal bf16 text Gram tiles (ceil(Bt/4), 128, 128)
  // large-shape tcgen05 kernels (pair_fwd3.cu): padded per-region / per-tile-row vectors and Gram tiles
  size_t alpha3, m2v3;       // fp32 (Bv, Rp): inverse norms (0 beyond R), log2-domain mask terms (-1e30 beyond R)
  size_t beta3, m2t3;        // fp32 (tiles, 128) in tile-row order (text slot * WP + word)
  size_t gbd3;               // bf16 (tiles, 128, 128) block-diagonal text Gram tiles for TPT texts of WP padded words
  size_t dgt3;               // fp32 (tiles, 128, 128) text Gram gradients per tile (pair_bwd3.cu), unpacked into dgt
  size_t ud1, cb;            // bf16 spills of pair_bwd3.cu: UD1 (units, 128, Rp), C_bf16 (units, 128, 256); P uses `pa`
  size_t scratch;            // generic kernels: per-CTA pair scratch
  size_t total;
  int n_cta;                 // generic kernels: persistent grid size
  size_t scratch_per_cta;    // floats
};

// which pair kernels serve a call
enum Path { kPathGeneric = 0, kPathTc2 = 2, kPathTc3 = 3 };   // fp32 CUDA cores; pair_fwd2/bwd2 (W = 32, R <= 80); pair_fwd3

// tiling of the large-shape kernels (pair_fwd3.cu: shape3)
struct Shape3 { int NB, NQ, WP, TPT, Rp, nch; };
bool shape3(int R, int W, int d, Shape3* s);

Workspace make_workspace(int Bv, int Bt, int R, int W, int d, int in_dtype, int path, bool train = true, int n_t = 0);

// arguments of the generic pair kernels
struct PairArgs {
  const float *vh, *th, *gv, *gt, *mterm_v, *mterm_t;
  int Bv, Bt, R, W, d;
  float lambda;
  int focal;
  float* scores;
  const float* d_scores;
  float *dvh, *dth, *dgv, *dgt;
  float* scratch;
  size_t scratch_per_cta;
};

}  // namespace dvlp
