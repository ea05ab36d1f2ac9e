// extern "C" entry points declared in include/demovlp_b200.h: validation + dispatch.
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace dvlp {

thread_local int g_last_cuda_error = 0;
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_default_impl{DVLP_IMPL_AUTO};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

struct ProfRec { int slot; cudaEvent_t a, b; };
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
ProfScope::ProfScope(int slot_, cudaStream_t st_) : slot(slot_), st(st_), rec(nullptr) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  auto* r = new ProfRec{slot, nullptr, nullptr};
  cudaEventCreate(&r->a);
  cudaEventCreate(&r->b);
  cudaEventRecord(r->a, st);
  rec = r;
}
ProfScope::~ProfScope() {
  if (!rec) return;
  auto* r = (ProfRec*)rec;
  cudaEventRecord(r->b, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(*r);
  delete r;
}

int run_prep(const void*, const void*, const float*, const float*, int, int, int, int, int, float, int, bool, bool, char*,
             const Workspace&, cudaStream_t);
int run_prep3(const void*, const void*, const float*, const float*, int, int, int, int, int, float, int, char*,
              const Workspace&, cudaStream_t, int n_t = 0);
int run_pair_bwd3(const void* im_bf16, const void* s_bf16, int Bv, int Bt, int R, int W, int d, float lambda,
                  const float* d_scores, char* ws, const Workspace& L, cudaStream_t st);
int run_grad_gemm3(const void* im_bf16, const void* s_bf16, int Bv, int Bt, int R, int W, int d, char* ws,
                   const Workspace& L, cudaStream_t st);
int run_pair_fwd3(const void* im_bf16, const void* s_bf16, int Bv, int Bt, int R, int W, int d, float lambda, int focal,
                  float* scores, char* ws, const Workspace& L, bool save, cudaStream_t st, int n_t = 0);
int run_pair_generic(bool backward, const PairArgs& a, int n_cta, cudaStream_t st);
int run_grad_finalize(const void*, int, const float*, const float*, const float*, int, int, int, int, float*,
                      cudaStream_t);
int run_grad_finalize_tc(const void*, const void*, const float*, const float*, const float*, const float*, const float*,
                         const float*, int, int, int, int, int, void*, void*, int, cudaStream_t);
int run_rwa_loss(const float*, int, int, int, int, float, float*, float*, cudaStream_t);
int run_normsoftmax(const float*, int, int, float, float*, float*, float*, cudaStream_t);
int run_sim_fwd(const void*, const void*, int, int, int, float, int, float*, float*, cudaStream_t);
int run_sim_bwd(const void*, const void*, int, int, int, float, int, const float*, void*, void*, int, float*,
                cudaStream_t);
int run_ranks_t2v(const double*, int, int, int32_t*, cudaStream_t);
int run_ranks_v2t(const double*, int, int, double*, cudaStream_t);
int run_p2p_allgather(void* const*, int, int, int, const void* const*, const size_t*, const size_t*, cudaStream_t);
int run_p2p_sum_scalar(void* const*, int, int, const float*, float*, cudaStream_t);
int run_p2p_reduce_scatter(void* const*, int, int, size_t, size_t, void*, int, cudaStream_t);
int run_p2p_check(const void*, cudaStream_t);
size_t proj_workspace_bytes(int M, int K, int in_dtype, int relu_in);
int attn_small_supported(int I, int J, int d);
int run_attn_small_fwd(const float*, const float*, const float*, const float*, int, int, int, float*, float*, cudaStream_t);
int run_attn_small_bwd(const float*, const float*, const float*, const float*, const float*, const float*, const float*, int,
                       int, int, float*, float*, float*, cudaStream_t);
int run_proj_fwd(const void* x, int in_dtype, int M, int K, const void* w_bf16, const float* bias, int relu_in, void* out,
                 char* ws, cudaStream_t st);
// tcgen05 path: pair_fwd2.cu, pair_bwd2.cu, grad_gemm2.cu
int run_pair_fwd2(const void* im_bf16, const void* s_bf16, int Bv, int Bt, int R, int W, int d, float lambda, int focal,
                  float* scores, char* ws, const Workspace& L, bool save, cudaStream_t st);
int run_pair_bwd2(const void* im_bf16, const void* s_bf16, int v0, int nv, int Bt, int R, int W, int d, float lambda,
                  const float* d_scores, char* ws, const Workspace& L, cudaStream_t st);
int run_grad_gemm2(const void* im_bf16, const void* s_bf16, int v0, int nv, int Bt, int R, int W, int d, char* ws,
                   const Workspace& L, cudaStream_t st);
void* g_trace_buffer = nullptr;   // set through dvlp_debug_set_trace (tools/trace_tiles.py)

// shapes served by the tcgen05 kernels: d = 256, W = 32 words, R a multiple of 16 in [32, 80]
bool tc_supported(int R, int W, int d) { return d == 256 && W == 32 && (R == 80 || R == 64 || R == 48 || R == 32); }

// The small-shape backward hands its bf16 gradient tiles P (20 KB per (video, 4-text tile) at R = 80) to the gradient
// GEMMs through memory: 84 MB at config 2, 671 MB per rank at config 3 -- more than the 126 MB L2 keeps.  The backward
// can therefore walk the videos in CHUNKS whose P tiles stay in L2 (pair kernel of a chunk, then both GEMMs of that
// chunk, all re-using ONE chunk-sized buffer): DVLP_P_CHUNK_MB = target chunk size in MB.  MEASURED on B200 (round 2,
// config 3 on one GPU, 40 MB chunks = 8 videos): the pair kernel gets 18 % slower (148 persistent CTAs x ~14 units per
// launch: pipeline fill / drain per launch) and the GEMMs 2.6 x slower (K loops of 7 tiles, split-K red.global adds),
// 25.0 ms per step instead of 20.2 -- the re-reads cost less than the short launches, the GEMMs run at ~900 algorithmic
// TFLOP/s unchunked and are not HBM-bound.  So chunking is OFF by default (0); the path stays for experiments.
// Returns the videos per chunk (the same for workspace sizing and launch).
int tc2_chunk_videos(int Bv, int Bt, int R) {
  static const double target_mb = [] {
    const char* e = getenv("DVLP_P_CHUNK_MB");
    return e ? atof(e) : 0.0;
  }();
  if (!(target_mb > 0.0)) return Bv;
  const int tiles = (Bt + 3) / 4;
  const double per_video = (double)tiles * 128 * R * 2;
  const double nvd = target_mb * 1048576.0 / per_video;
  if (nvd >= (double)Bv) return Bv;
  int nv = (int)nvd;
  if (nv < 1) nv = 1;
  const int n_chunks = (Bv + nv - 1) / nv;
  nv = (Bv + n_chunks - 1) / n_chunks;               // even chunks
  // prefer a size whose unit count fills the 148 persistent CTAs evenly
  int best = nv;
  double best_eff = 0.0;
  for (int c = nv; c >= 1 && c >= nv - nv / 4; --c) {
    const long long units = (long long)c * tiles;
    const double eff = (double)units / (double)(((units + 147) / 148) * 148);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = c; }
  }
  return best;
}

Workspace make_workspace(int Bv, int Bt, int R, int W, int d, int in_dtype, int path, bool train, int n_t) {
  Workspace L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
  const bool tc = path != kPathGeneric, tc2 = path == kPathTc2, tc3 = path == kPathTc3;
  const size_t nv = (size_t)Bv * R, nt = (size_t)Bt * W;
  // text-side vectors are padded to whole 128-row tiles (the tcgen05 kernel bulk-copies tiles)
  const size_t nt_pad = (nt + 127) / 128 * 128;
  L.alpha = take(nv * 4);
  L.beta = take(nt_pad * 4);
  L.mterm_v = take(nv * 4);
  L.mterm_t = take(nt_pad * 4);
  // (train == false: forward-only layout -- nothing the backward alone needs is allocated)
  const bool generic_rows = (train && !tc3) || !tc;   // fp32 normalised rows / Gram matrices of the generic kernels
  L.vh = take(generic_rows ? nv * d * 4 : 0);
  L.th = take(generic_rows ? nt * d * 4 : 0);
  L.gv = take(generic_rows ? nv * R * 4 : 0);
  L.gt = take(generic_rows ? nt * W * 4 : 0);
  bool copy = tc && in_dtype == DVLP_F32;
  L.vb = take(copy ? nv * d * 2 : 0);
  L.tb = take(copy ? nt * d * 2 : 0);
  L.gvb = take(tc2 ? nv * R * 2 : 0);
  L.gtb = take(tc ? nt_pad * W * 2 : 0);
  L.m2v = take(tc ? nv * 4 : 0);
  L.m2t = take(tc ? nt_pad * 4 : 0);
  L.dvh = take(train ? nv * d * 4 : 0);
  L.dth = take(train ? nt * d * 4 : 0);
  L.dgv = take(train ? nv * R * 4 : 0);
  L.dgt = take(train ? nt * W * 4 : 0);
  if (tc3 && train) {          // (directly behind dgt: one memset clears every accumulator of the backward)
    Shape3 s3{};
    shape3(R, W, d, &s3);
    const size_t tiles = (size_t)(Bt + s3.TPT - 1) / s3.TPT, units = tiles * Bv;
    L.dgt3 = take(tiles * 128 * 128 * 4);
    L.pa = take(units * 128 * s3.Rp * 2);
    L.ud1 = take(units * 128 * s3.Rp * 2);
    L.cb = take(units * 128 * 256 * 2);
    L.rowv = take(units * 128 * 16);
    L.colv = take(units * s3.TPT * s3.Rp * 16);
  } else {
    const size_t units = (size_t)Bv * ((Bt + 3) / 4);
    const size_t p_units = tc2 && train ? (size_t)tc2_chunk_videos(Bv, Bt, R) * ((Bt + 3) / 4) : 0;    // one chunk of videos
    L.pa = take(p_units * R * 128 * 2);
    L.rowv = take(tc2 && train ? units * 128 * 16 : 0);
    L.colv = take(tc2 && train ? units * 4 * R * 16 : 0);
    L.gbd = take(tc2 ? (size_t)((Bt + 3) / 4) * 128 * 128 * 2 : 0);
  }
  if (tc3) {
    Shape3 s3{};
    shape3(R, W, d, &s3);
    // grouped (multiple-choice) calls tile every group's n_t texts separately
    const size_t tiles = n_t > 0 ? (size_t)Bv * ((n_t + s3.TPT - 1) / s3.TPT) : (size_t)(Bt + s3.TPT - 1) / s3.TPT;
    L.alpha3 = take((size_t)Bv * s3.Rp * 4);
    L.m2v3 = take((size_t)Bv * s3.Rp * 4);
    L.beta3 = take(tiles * 128 * 4);
    L.m2t3 = take(tiles * 128 * 4);
    L.gbd3 = take(tiles * 128 * 128 * 2);
  }
  // generic kernels (also the backward fallback of the tcgen05 forward): per-CTA scratch
  long long pairs = (long long)Bv * Bt;
  int sms = 148;
  L.n_cta = (int)(pairs < (long long)sms * 4 ? pairs : (long long)sms * 4);
  int Lmax = R > W ? R : W;
  L.scratch_per_cta = align_up(((size_t)4 * R * W + 5 * Lmax) * 4) / 4;
  L.scratch = take(tc3 ? 0 : (size_t)L.n_cta * L.scratch_per_cta * 4);
  L.total = off;
  return L;
}

static int check_common(const void* im, const void* s, const float* im_mask, const float* s_mask, int Bv, int Bt,
                        int R, int W, int d, int in_dtype) {
  if (!im || !s || !im_mask || !s_mask) return DVLP_ERR_NULL_POINTER;
  if (Bv <= 0 || Bt <= 0 || R <= 0 || W <= 0 || d <= 0 || (d % 8) != 0) return DVLP_ERR_BAD_SHAPE;
  if (in_dtype != DVLP_F32 && in_dtype != DVLP_BF16) return DVLP_ERR_BAD_DTYPE;
  if (((uintptr_t)im | (uintptr_t)s) & 15) return DVLP_ERR_MISALIGNED;
  return DVLP_OK;
}

// The tcgen05 kernels evaluate exp(lambda (A - 1)) with the FIXED maximum lambda (A <= 1) instead of a running row
// maximum: beyond this bound a row of small attentions could underflow entirely (the reference's max-subtracting
// softmax would not), so such calls go to the generic kernels (auto) or are refused (explicit tcgen05).
constexpr float kMaxLambdaTc = 40.f;
// the large-shape kernels (pair_fwd3.cu) serve the forward; their backward partner is selected here once it exists
constexpr bool kTc3Train = true;

// kPathTc2: W = 32 and R in {32,48,64,80} (pair_fwd2 / pair_bwd2); kPathTc3: any other d = 256 shape the streaming
// large-shape kernel tiles (R up to 288 at W <= 64, W <= 128: config 5, the shipped R = 30 / 240 x W = 99 shapes)
static int shape_path(int R, int W, int d, bool train) {
  if (tc_supported(R, W, d)) return kPathTc2;
  if ((!train || kTc3Train) && shape3(R, W, d, nullptr)) return kPathTc3;
  return kPathGeneric;
}

static int pick_path(int impl, int R, int W, int d, bool train, int* path, float lambda = 20.f) {
  if (impl == DVLP_IMPL_AUTO) impl = g_default_impl.load();
  if (impl == DVLP_IMPL_GENERIC) { *path = kPathGeneric; return DVLP_OK; }
  if (impl != DVLP_IMPL_AUTO && impl != DVLP_IMPL_TCGEN05) return DVLP_ERR_BAD_ARG;
  int p = shape_path(R, W, d, train);
  if (!(lambda > 0.f && lambda <= kMaxLambdaTc)) {
    if (impl == DVLP_IMPL_TCGEN05) return DVLP_ERR_BAD_ARG;
    p = kPathGeneric;
  }
  if (impl == DVLP_IMPL_TCGEN05 && p == kPathGeneric) return DVLP_ERR_BAD_SHAPE;
  *path = p;
  return DVLP_OK;
}

static int device_ok() {
  // per device: a process may drive several GPUs (the result of device 0 says nothing about device 1)
  static std::atomic<unsigned long long> ok_mask{0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return DVLP_ERR_NO_DEVICE;
  if (dev >= 0 && dev < 64 && ((ok_mask.load(std::memory_order_acquire) >> dev) & 1ull)) return DVLP_OK;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10)
    return DVLP_ERR_NO_DEVICE;
  if (dev >= 0 && dev < 64) ok_mask.fetch_or(1ull << dev, std::memory_order_release);
  return DVLP_OK;
}

}  // namespace dvlp

using namespace dvlp;

extern "C" {

int dvlp_version(void) { return 100; }

const char* dvlp_error_string(int code) {
  switch (code) {
    case DVLP_OK: return "ok";
    case DVLP_ERR_BAD_SHAPE: return "bad or unsupported shape";
    case DVLP_ERR_BAD_DTYPE: return "unsupported dtype";
    case DVLP_ERR_NULL_POINTER: return "null pointer";
    case DVLP_ERR_MISALIGNED: return "pointer not 16-byte aligned";
    case DVLP_ERR_WORKSPACE: return "workspace missing or too small";
    case DVLP_ERR_CUDA: return "CUDA error";
    case DVLP_ERR_NO_DEVICE: return "no sm_100 CUDA device (no CPU fallback exists)";
    case DVLP_ERR_BAD_ARG: return "bad argument";
    case DVLP_ERR_TIMEOUT: return "peer-memory exchange timed out (a rank is missing or late beyond DVLP_P2P_TIMEOUT_S)";
  }
  return "unknown error";
}

int dvlp_last_cuda_error(void) { return g_last_cuda_error; }
int dvlp_rw_tcgen05_supported(int R, int W, int d) { return shape_path(R, W, d, true) != kPathGeneric ? 1 : 0; }
int dvlp_rw_path(int R, int W, int d, int train) { return shape_path(R, W, d, train != 0); }
void dvlp_set_default_impl(int impl) { g_default_impl.store(impl); }
uint64_t dvlp_launch_count(void) { return g_launches.load(); }

/* debug only: device buffer of 5*32*16 int64 clock stamps written by pair_bwd2.cu (tools/trace_tiles.py) */
void dvlp_debug_set_trace(void* dev_buffer) { g_trace_buffer = dev_buffer; }

void dvlp_profile_enable(int on) { g_prof_on.store(on ? 1 : 0); }

int dvlp_profile_read(double* ms, int* counts) {
  if (!ms || !counts) return DVLP_ERR_NULL_POINTER;
  for (int i = 0; i < kProfSlots; ++i) { ms[i] = 0.0; counts[i] = 0; }
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) {
    float e = 0.f;
    DVLP_CUDA_OK(cudaEventSynchronize(r.b));
    DVLP_CUDA_OK(cudaEventElapsedTime(&e, r.a, r.b));
    ms[r.slot] += e;
    counts[r.slot] += 1;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  return DVLP_OK;
}

static size_t workspace_bytes(int Bv, int Bt, int R, int W, int d, int in_dtype, int impl, bool train) {
  if (Bv <= 0 || Bt <= 0 || R <= 0 || W <= 0 || d <= 0) return 0;
  int path = kPathGeneric;
  if (pick_path(impl, R, W, d, train, &path) != DVLP_OK) return 0;
  size_t n = make_workspace(Bv, Bt, R, W, d, in_dtype, path, train).total;
  // The size query does not know lambda: under DVLP_IMPL_AUTO a call with lambda beyond the tensor-core bound (40)
  // is served by the generic kernels, so the workspace must be large enough for either path.
  int eff = impl == DVLP_IMPL_AUTO ? g_default_impl.load() : impl;
  if (eff == DVLP_IMPL_AUTO && path != kPathGeneric) {
    const size_t g = make_workspace(Bv, Bt, R, W, d, in_dtype, kPathGeneric, train).total;
    if (g > n) n = g;
  }
  return n;
}
size_t dvlp_rw_workspace_bytes(int Bv, int Bt, int R, int W, int d, int in_dtype, int impl) {
  return workspace_bytes(Bv, Bt, R, W, d, in_dtype, impl, true);
}
size_t dvlp_rw_infer_workspace_bytes(int Bv, int Bt, int R, int W, int d, int in_dtype, int impl) {
  return workspace_bytes(Bv, Bt, R, W, d, in_dtype, impl, false);
}

static int scores_fwd(const void* im, const void* s, const float* im_mask, const float* s_mask, int Bv, int Bt,
                      int R, int W, int d, float lambda, int focal_equal, int in_dtype, int impl, float* scores,
                      void* workspace, size_t workspace_bytes, void* stream, bool train) {
  int rc = check_common(im, s, im_mask, s_mask, Bv, Bt, R, W, d, in_dtype);
  if (rc) return rc;
  if (!scores) return DVLP_ERR_NULL_POINTER;
  int path = kPathGeneric;
  if ((rc = pick_path(impl, R, W, d, train, &path, lambda))) return rc;
  if ((rc = device_ok())) return rc;
  const bool tc = path != kPathGeneric;
  Workspace L = make_workspace(Bv, Bt, R, W, d, in_dtype, path, train);
  if (!workspace || workspace_bytes < L.total) return DVLP_ERR_WORKSPACE;
  if ((uintptr_t)workspace & 255) return DVLP_ERR_MISALIGNED;
  char* ws = (char*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope p(kProfPrep, st);
    if (path == kPathTc3) {
      if ((rc = run_prep3(im, s, im_mask, s_mask, Bv, Bt, R, W, d, lambda, in_dtype, ws, L, st))) return rc;
    } else if ((rc = run_prep(im, s, im_mask, s_mask, Bv, Bt, R, W, d, lambda, in_dtype, tc, false, ws, L, st))) {
      return rc;
    }
  }
  ProfScope p(kProfFwd, st);
  if (tc) {
    const void* vb = in_dtype == DVLP_BF16 ? im : (const void*)(ws + L.vb);
    const void* tb = in_dtype == DVLP_BF16 ? s : (const void*)(ws + L.tb);
    if (path == kPathTc3) return run_pair_fwd3(vb, tb, Bv, Bt, R, W, d, lambda, focal_equal, scores, ws, L, train, st);
    return run_pair_fwd2(vb, tb, Bv, Bt, R, W, d, lambda, focal_equal, scores, ws, L, train, st);
  }
  PairArgs a{};
  a.vh = (const float*)(ws + L.vh); a.th = (const float*)(ws + L.th);
  a.gv = (const float*)(ws + L.gv); a.gt = (const float*)(ws + L.gt);
  a.mterm_v = (const float*)(ws + L.mterm_v); a.mterm_t = (const float*)(ws + L.mterm_t);
  a.Bv = Bv; a.Bt = Bt; a.R = R; a.W = W; a.d = d; a.lambda = lambda; a.focal = focal_equal;
  a.scores = scores;
  a.scratch = (float*)(ws + L.scratch); a.scratch_per_cta = L.scratch_per_cta;
  return run_pair_generic(false, a, L.n_cta, st);
}

int dvlp_rw_scores_fwd(const void* im, const void* s, const float* im_mask, const float* s_mask, int Bv, int Bt,
                       int R, int W, int d, float lambda, int focal_equal, int in_dtype, int impl, float* scores,
                       void* workspace, size_t workspace_bytes, void* stream) {
  return scores_fwd(im, s, im_mask, s_mask, Bv, Bt, R, W, d, lambda, focal_equal, in_dtype, impl, scores, workspace,
                    workspace_bytes, stream, true);
}

int dvlp_rw_scores_infer(const void* im, const void* s, const float* im_mask, const float* s_mask, int Bv, int Bt,
                         int R, int W, int d, float lambda, int focal_equal, int in_dtype, int impl, float* scores,
                         void* workspace, size_t workspace_bytes, void* stream) {
  return scores_fwd(im, s, im_mask, s_mask, Bv, Bt, R, W, d, lambda, focal_equal, in_dtype, impl, scores, workspace,
                    workspace_bytes, stream, false);
}

size_t dvlp_rw_grouped_workspace_bytes(int n_groups, int n_per_group, int R, int W, int d, int in_dtype) {
  if (n_groups <= 0 || n_per_group <= 0 || !shape3(R, W, d, nullptr)) return 0;
  return make_workspace(n_groups, n_groups * n_per_group, R, W, d, in_dtype, kPathTc3, false, n_per_group).total;
}

int dvlp_rw_scores_grouped_infer(const void* im, const void* s, const float* im_mask, const float* s_mask, int n_groups,
                                 int n_per_group, int R, int W, int d, float lambda, int focal_equal, int in_dtype,
                                 float* scores, void* workspace, size_t workspace_bytes, void* stream) {
  const int Bv = n_groups, Bt = n_groups * n_per_group;
  if (n_per_group <= 0) return DVLP_ERR_BAD_SHAPE;
  int rc = check_common(im, s, im_mask, s_mask, Bv, Bt, R, W, d, in_dtype);
  if (rc) return rc;
  if (!scores) return DVLP_ERR_NULL_POINTER;
  if (!shape3(R, W, d, nullptr)) return DVLP_ERR_BAD_SHAPE;
  if (!(lambda > 0.f && lambda <= kMaxLambdaTc)) return DVLP_ERR_BAD_ARG;
  if ((rc = device_ok())) return rc;
  Workspace L = make_workspace(Bv, Bt, R, W, d, in_dtype, kPathTc3, false, n_per_group);
  if (!workspace || workspace_bytes < L.total) return DVLP_ERR_WORKSPACE;
  if ((uintptr_t)workspace & 255) return DVLP_ERR_MISALIGNED;
  char* ws = (char*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope p(kProfPrep, st);
    if ((rc = run_prep3(im, s, im_mask, s_mask, Bv, Bt, R, W, d, lambda, in_dtype, ws, L, st, n_per_group))) return rc;
  }
  ProfScope p(kProfFwd, st);
  const void* vb = in_dtype == DVLP_BF16 ? im : (const void*)(ws + L.vb);
  const void* tb = in_dtype == DVLP_BF16 ? s : (const void*)(ws + L.tb);
  return run_pair_fwd3(vb, tb, Bv, Bt, R, W, d, lambda, focal_equal, scores, ws, L, false, st, n_per_group);
}

int dvlp_rw_scores_bwd(const void* im, const void* s, const float* im_mask, const float* s_mask, int Bv, int Bt,
                       int R, int W, int d, float lambda, int focal_equal, int in_dtype, int impl,
                       const float* d_scores, void* d_im, void* d_s, int out_dtype, void* workspace,
                       size_t workspace_bytes, void* stream) {
  int rc = check_common(im, s, im_mask, s_mask, Bv, Bt, R, W, d, in_dtype);
  if (rc) return rc;
  if (!d_scores || !d_im || !d_s) return DVLP_ERR_NULL_POINTER;
  if (out_dtype != DVLP_F32 && out_dtype != DVLP_BF16) return DVLP_ERR_BAD_DTYPE;
  int path = kPathGeneric;
  if ((rc = pick_path(impl, R, W, d, true, &path, lambda))) return rc;
  if ((rc = device_ok())) return rc;
  const bool tc = path != kPathGeneric;
  Workspace L = make_workspace(Bv, Bt, R, W, d, in_dtype, path, true);
  if (!workspace || workspace_bytes < L.total) return DVLP_ERR_WORKSPACE;
  char* ws = (char*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  // zero the accumulators (one contiguous range: dvh .. dgt [.. dgt3]; the tcgen05 small-shape dV' GEMM overwrites dvh,
  // so it starts at dth)
  {
    const size_t z0 = path == kPathTc2 ? L.dth : L.dvh;
    DVLP_CUDA_OK(cudaMemsetAsync(ws + z0, 0, L.pa - z0, st));
  }
  if (path == kPathTc3) {
    const void* vb = in_dtype == DVLP_BF16 ? im : (const void*)(ws + L.vb);
    const void* tb = in_dtype == DVLP_BF16 ? s : (const void*)(ws + L.tb);
    if (out_dtype != DVLP_F32) return DVLP_ERR_BAD_DTYPE;
    {
      ProfScope pb(kProfBwd, st);
      if ((rc = run_pair_bwd3(vb, tb, Bv, Bt, R, W, d, lambda, d_scores, ws, L, st))) return rc;
    }
    {
      ProfScope pg(kProfGG, st);
      if ((rc = run_grad_gemm3(vb, tb, Bv, Bt, R, W, d, ws, L, st))) return rc;
    }
    // l2-norm backward of the raw rows; the video side has no Gram term in the direct form
    ProfScope pf(kProfFinal, st);
    if ((rc = run_grad_finalize(vb, DVLP_BF16, (const float*)(ws + L.alpha), (const float*)(ws + L.dvh), nullptr, Bv, R, d, 1,
                                (float*)d_im, st)))
      return rc;
    return run_grad_finalize(tb, DVLP_BF16, (const float*)(ws + L.beta), (const float*)(ws + L.dth),
                             (const float*)(ws + L.dgt), Bt, W, d, 1, (float*)d_s, st);
  }
  int prescaled = 0, dgt_swz = 0;
  {
  if (path == kPathTc2) {
    const void* vb = in_dtype == DVLP_BF16 ? im : (const void*)(ws + L.vb);
    const void* tb = in_dtype == DVLP_BF16 ? s : (const void*)(ws + L.tb);
    // chunks of videos whose P tiles stay in L2 between the pair kernel and the gradient GEMMs (tc2_chunk_videos)
    const int chunk = tc2_chunk_videos(Bv, Bt, R);
    for (int v0 = 0; v0 < Bv; v0 += chunk) {
      const int nv = Bv - v0 < chunk ? Bv - v0 : chunk;
      {
        ProfScope pb(kProfBwd, st);
        if ((rc = run_pair_bwd2(vb, tb, v0, nv, Bt, R, W, d, lambda, d_scores, ws, L, st))) return rc;
      }
      {
        ProfScope pg(kProfGG, st);
        if ((rc = run_grad_gemm2(vb, tb, v0, nv, Bt, R, W, d, ws, L, st))) return rc;
      }
    }
    dgt_swz = 1;
    prescaled = 1;
  } else {
    PairArgs a{};
    a.vh = (const float*)(ws + L.vh); a.th = (const float*)(ws + L.th);
    a.gv = (const float*)(ws + L.gv); a.gt = (const float*)(ws + L.gt);
    a.mterm_v = (const float*)(ws + L.mterm_v); a.mterm_t = (const float*)(ws + L.mterm_t);
    a.Bv = Bv; a.Bt = Bt; a.R = R; a.W = W; a.d = d; a.lambda = lambda; a.focal = focal_equal;
    a.d_scores = d_scores;
    a.dvh = (float*)(ws + L.dvh); a.dth = (float*)(ws + L.dth);
    a.dgv = (float*)(ws + L.dgv); a.dgt = (float*)(ws + L.dgt);
    a.scratch = (float*)(ws + L.scratch); a.scratch_per_cta = L.scratch_per_cta;
    ProfScope pb(kProfBwd, st);
    if ((rc = run_pair_generic(true, a, L.n_cta, st))) return rc;
  }
  }
  ProfScope pf(kProfFinal, st);
  const void* xv = (tc && in_dtype == DVLP_F32) ? (const void*)(ws + L.vb) : im;
  const void* xt = (tc && in_dtype == DVLP_F32) ? (const void*)(ws + L.tb) : s;
  int xdt = (tc && in_dtype == DVLP_F32) ? DVLP_BF16 : in_dtype;
  if (prescaled && d == 256 && (R % 16) == 0 && (W % 16) == 0 && R <= 128 && W <= 128)
    return run_grad_finalize_tc(xv, xt, (const float*)(ws + L.alpha), (const float*)(ws + L.beta),
                                (const float*)(ws + L.dvh), (const float*)(ws + L.dth), (const float*)(ws + L.dgv),
                                (const float*)(ws + L.dgt), Bv, Bt, R, W, dgt_swz, d_im, d_s, out_dtype, st);
  if (out_dtype != DVLP_F32) return DVLP_ERR_BAD_DTYPE;       // the fp32 CUDA-core finalisation writes fp32 only
  if ((rc = run_grad_finalize(xv, xdt, (const float*)(ws + L.alpha), (const float*)(ws + L.dvh),
                              (const float*)(ws + L.dgv), Bv, R, d, prescaled, (float*)d_im, st)))
    return rc;
  return run_grad_finalize(xt, xdt, (const float*)(ws + L.beta), (const float*)(ws + L.dth),
                           (const float*)(ws + L.dgt), Bt, W, d, prescaled, (float*)d_s, st);
}

int dvlp_rwa_loss_fwd_bwd(const float* scores, int n_rows, int n_cols, int row_offset, int n_total, float lambda,
                          float* loss_sum, float* d_scores, void* stream) {
  if (!scores || !loss_sum) return DVLP_ERR_NULL_POINTER;
  if (n_rows <= 0 || n_cols <= 0 || n_total <= 0 || row_offset < 0 || row_offset + n_rows > n_cols)
    return DVLP_ERR_BAD_SHAPE;
  int rc = device_ok();
  if (rc) return rc;
  return run_rwa_loss(scores, n_rows, n_cols, row_offset, n_total, lambda, loss_sum, d_scores, (cudaStream_t)stream);
}

int dvlp_sim_matrix_fwd(const void* a, const void* b, int N, int M, int d, float eps, int in_dtype, float* out,
                        float* scratch, void* stream) {
  if (!a || !b || !out || !scratch) return DVLP_ERR_NULL_POINTER;
  if (N <= 0 || M <= 0 || d <= 0) return DVLP_ERR_BAD_SHAPE;
  if (in_dtype != DVLP_F32 && in_dtype != DVLP_BF16) return DVLP_ERR_BAD_DTYPE;
  int rc = device_ok();
  if (rc) return rc;
  return run_sim_fwd(a, b, N, M, d, eps, in_dtype, out, scratch, (cudaStream_t)stream);
}

int dvlp_sim_matrix_bwd(const void* a, const void* b, int N, int M, int d, float eps, int in_dtype,
                        const float* d_out, void* d_a, void* d_b, int out_dtype, float* scratch, void* stream) {
  if (!a || !b || !d_out || !scratch) return DVLP_ERR_NULL_POINTER;
  if (N <= 0 || M <= 0 || d <= 0) return DVLP_ERR_BAD_SHAPE;
  if (in_dtype != DVLP_F32 && in_dtype != DVLP_BF16) return DVLP_ERR_BAD_DTYPE;
  if (out_dtype != DVLP_F32 && out_dtype != DVLP_BF16) return DVLP_ERR_BAD_DTYPE;
  int rc = device_ok();
  if (rc) return rc;
  return run_sim_bwd(a, b, N, M, d, eps, in_dtype, d_out, d_a, d_b, out_dtype, scratch, (cudaStream_t)stream);
}

int dvlp_normsoftmax_fwd_bwd(const float* x, int N, int M, float temperature, float* loss, float* d_x,
                             float* scratch, void* stream) {
  if (!x || !loss || !scratch) return DVLP_ERR_NULL_POINTER;
  if (N <= 0 || M <= 0) return DVLP_ERR_BAD_SHAPE;
  if (!(temperature > 0.f)) return DVLP_ERR_BAD_ARG;
  int rc = device_ok();
  if (rc) return rc;
  return run_normsoftmax(x, N, M, temperature, loss, d_x, scratch, (cudaStream_t)stream);
}

int dvlp_ranks_t2v(const double* sims, int Nq, int Nv, int32_t* ranks, void* stream) {
  if (!sims || !ranks) return DVLP_ERR_NULL_POINTER;
  if (Nq <= 0 || Nv <= 0 || Nq < Nv) return DVLP_ERR_BAD_SHAPE;
  int rc = device_ok();
  if (rc) return rc;
  return run_ranks_t2v(sims, Nq, Nv, ranks, (cudaStream_t)stream);
}

int dvlp_ranks_v2t(const double* sims, int Nq, int Nv, double* ranks, void* stream) {
  if (!sims || !ranks) return DVLP_ERR_NULL_POINTER;
  if (Nq <= 0 || Nv <= 0 || Nq < Nv) return DVLP_ERR_BAD_SHAPE;
  int rc = device_ok();
  if (rc) return rc;
  return run_ranks_v2t(sims, Nq, Nv, ranks, (cudaStream_t)stream);
}

int dvlp_rw_bwd_out_dtype_supported(int impl, int R, int W, int d, float lambda, int out_dtype) {
  if (out_dtype == DVLP_F32) return 1;
  if (out_dtype != DVLP_BF16) return 0;
  int path = kPathGeneric;
  if (pick_path(impl, R, W, d, true, &path, lambda) != DVLP_OK || path != kPathTc2) return 0;
  return (d == 256 && (R % 16) == 0 && (W % 16) == 0 && R <= 128 && W <= 128) ? 1 : 0;   // run_grad_finalize_tc
}

size_t dvlp_proj_workspace_bytes(int M, int K, int in_dtype, int relu_in) {
  if (M <= 0 || K <= 0 || (K % 64) != 0) return 0;
  return proj_workspace_bytes(M, K, in_dtype, relu_in);
}

int dvlp_proj_fwd(const void* x, int in_dtype, int M, int K, int N, const void* w_bf16, const float* bias, int relu_in,
                  void* out_bf16, void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !w_bf16 || !out_bf16) return DVLP_ERR_NULL_POINTER;
  if (M <= 0 || K <= 0 || (K % 64) != 0 || N != 256) return DVLP_ERR_BAD_SHAPE;
  if (in_dtype != DVLP_F32 && in_dtype != DVLP_BF16) return DVLP_ERR_BAD_DTYPE;
  if (((uintptr_t)x | (uintptr_t)w_bf16 | (uintptr_t)out_bf16) & 15) return DVLP_ERR_MISALIGNED;
  if (!workspace || workspace_bytes < proj_workspace_bytes(M, K, in_dtype, relu_in)) return DVLP_ERR_WORKSPACE;
  if ((uintptr_t)workspace & 255) return DVLP_ERR_MISALIGNED;
  int rc = device_ok();
  if (rc) return rc;
  return run_proj_fwd(x, in_dtype, M, K, w_bf16, bias, relu_in, out_bf16, (char*)workspace, (cudaStream_t)stream);
}

int dvlp_attn_small_supported(int I, int J, int d) { return attn_small_supported(I, J, d); }

int dvlp_attn_small_fwd(const float* q, const float* k, const float* v, const float* mask, int B, int I, int J, int d,
                        float* out, float* lse, void* stream) {
  if (!q || !k || !v || !out) return DVLP_ERR_NULL_POINTER;
  if (B <= 0 || !attn_small_supported(I, J, d)) return DVLP_ERR_BAD_SHAPE;
  if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) return DVLP_ERR_MISALIGNED;
  int rc = device_ok();
  if (rc) return rc;
  return run_attn_small_fwd(q, k, v, mask, B, I, J, out, lse, (cudaStream_t)stream);
}

int dvlp_attn_small_bwd(const float* q, const float* k, const float* v, const float* mask, const float* out,
                        const float* lse, const float* d_out, int B, int I, int J, int d, float* d_q, float* d_k,
                        float* d_v, void* stream) {
  if (!q || !k || !v || !out || !lse || !d_out || !d_q || !d_k || !d_v) return DVLP_ERR_NULL_POINTER;
  if (B <= 0 || !attn_small_supported(I, J, d)) return DVLP_ERR_BAD_SHAPE;
  if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out | (uintptr_t)d_out | (uintptr_t)d_q | (uintptr_t)d_k | (uintptr_t)d_v) & 15)
    return DVLP_ERR_MISALIGNED;
  int rc = device_ok();
  if (rc) return rc;
  return run_attn_small_bwd(q, k, v, mask, out, lse, d_out, B, I, J, d_q, d_k, d_v, (cudaStream_t)stream);
}

size_t dvlp_p2p_control_bytes(void) { return 4096; }

int dvlp_p2p_check(const void* own_arena, void* stream) {
  if (!own_arena) return DVLP_ERR_NULL_POINTER;
  return run_p2p_check(own_arena, (cudaStream_t)stream);
}

int dvlp_p2p_allgather(void* const* peer_arenas, int world, int rank, int n_tensors, const void* const* shards_host,
                       const size_t* shard_bytes_host, const size_t* arena_offsets_host, void* stream) {
  if (!peer_arenas || !shards_host || !shard_bytes_host || !arena_offsets_host) return DVLP_ERR_NULL_POINTER;
  if (world < 1 || world > 64 || rank < 0 || rank >= world || n_tensors < 1 || n_tensors > 8) return DVLP_ERR_BAD_ARG;
  for (int t = 0; t < n_tensors; ++t) {
    if (!shards_host[t]) return DVLP_ERR_NULL_POINTER;
    if (shard_bytes_host[t] == 0 || (shard_bytes_host[t] & 15) || (arena_offsets_host[t] & 15) ||
        ((uintptr_t)shards_host[t] & 15))
      return DVLP_ERR_MISALIGNED;
    if (arena_offsets_host[t] < dvlp_p2p_control_bytes()) return DVLP_ERR_BAD_ARG;
  }
  int rc = device_ok();
  if (rc) return rc;
  return run_p2p_allgather(peer_arenas, world, rank, n_tensors, shards_host, shard_bytes_host, arena_offsets_host,
                           (cudaStream_t)stream);
}

int dvlp_p2p_sum_scalar(void* const* peer_arenas, int world, int rank, const float* value, float* out, void* stream) {
  if (!peer_arenas || !value || !out) return DVLP_ERR_NULL_POINTER;
  if (world < 1 || world > 64 || rank < 0 || rank >= world) return DVLP_ERR_BAD_ARG;
  int rc = device_ok();
  if (rc) return rc;
  return run_p2p_sum_scalar(peer_arenas, world, rank, value, out, (cudaStream_t)stream);
}

int dvlp_p2p_reduce_scatter(void* const* peer_arenas, int world, int rank, size_t arena_offset, size_t shard_floats,
                            void* out, int out_dtype, void* stream) {
  if (!peer_arenas || !out) return DVLP_ERR_NULL_POINTER;
  if (world < 1 || world > 64 || rank < 0 || rank >= world) return DVLP_ERR_BAD_ARG;
  if (out_dtype != DVLP_F32 && out_dtype != DVLP_BF16) return DVLP_ERR_BAD_DTYPE;
  if (shard_floats == 0 || (shard_floats & 3) || (arena_offset & 15) || ((uintptr_t)out & 15)) return DVLP_ERR_MISALIGNED;
  if (arena_offset < dvlp_p2p_control_bytes()) return DVLP_ERR_BAD_ARG;
  int rc = device_ok();
  if (rc) return rc;
  return run_p2p_reduce_scatter(peer_arenas, world, rank, arena_offset, shard_floats, out, out_dtype,
                                (cudaStream_t)stream);
}

}  // extern "C"
