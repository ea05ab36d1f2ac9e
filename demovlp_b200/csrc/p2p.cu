// Peer-memory exchanges of the sharded region-word loss over NVLink / NVSwitch (one process per GPU).
//
// The three exchange steps of the video-row sharding (the reference's AllGather_multi pattern,
// trainer/trainer_dist.py:13-31) are a few MiB each, i.e. latency-bound: what matters is one launch and one
// flag round trip per step, not link bandwidth.  Every rank owns an ARENA with the same layout (allocated as
// symmetric memory by the host side, demovlp_b200/p2p.py); `peer[r]` is this process's mapping of rank r's arena.
//
//   all-gather      PUSH: every rank stores its shard of each tensor straight into all arenas, then raises its
//                   flag in every arena and waits until all ranks have raised theirs in its own.
//   scalar sum      every rank stores (epoch, value) into slot [rank] of all arenas, waits for the world's slots
//                   in its own arena and adds them in rank order (identical bits on every rank).
//   reduce-scatter  PULL: after a flag round trip ("my partial sums are complete") every rank reads its rows from
//                   all arenas, adds them in rank order and writes fp32 or bf16.
//
// Flags carry a monotonically increasing epoch kept in the arena (advanced by the kernels themselves), so the
// launches can be captured in a CUDA graph and replayed.  Re-use of the data regions is ordered by the next
// exchange of the step sequence gather -> scalar -> reduce-scatter -> gather ....
// Waits are bounded by wall time (minutes by default) and end in a sticky error word, not a trap (wait_flag).
#include <cstdlib>

#include "common.cuh"

namespace dvlp {

namespace p2p {
constexpr int kMaxWorld = 64, kMaxSeg = 8;
// control block at the start of every arena (bytes)
constexpr size_t kCtrGather = 0, kCtrScalar = 64, kCtrReduce = 128;      // {epoch, block arrivals} u64 pairs
constexpr size_t kErrWord = 192;      // sticky u64: 0 = healthy, else 1 + the u64 index of the flag a wait gave up on
constexpr size_t kFlagGather = 256, kFlagReduce = 256 + 512, kSlotScalar = 256 + 1024;   // u64 [kMaxWorld] each

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Spin until *p (a flag written by a peer into THIS rank's arena) reaches `want`.  The wait is bounded by WALL TIME
// (%globaltimer), minutes by default (DVLP_P2P_TIMEOUT_S, host side): ranks of a real training job legitimately drift
// by many seconds (checkpoint writes, data-loader stalls, lazy initialisation), where NCCL would simply wait.  When
// the bound is exceeded nothing traps: the waiting rank records a STICKY error word in its own arena's control block
// and leaves; every later wait on that arena returns at once, the kernels finish (with meaningless data) and the host
// turns the word into DVLP_ERR_TIMEOUT through dvlp_p2p_check (PeerExchange.check_errors).
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long wait_flag(char* mine, unsigned long long timeout_ns,
                                                        const unsigned long long* p, unsigned long long want,
                                                        unsigned long long mask = ~0ull, int shift = 0) {
  unsigned long long* err = reinterpret_cast<unsigned long long*>(mine + kErrWord);
  unsigned long long v;
  unsigned int spins = 0;
  unsigned long long t0 = 0;
  while ((((v = ld_acquire_sys(p)) & mask) >> shift) < want) {
    __nanosleep(64);
    if ((++spins & 1023u) == 0) {                 // every ~64 us: look at the clock and at the sticky error word
      if (*reinterpret_cast<volatile unsigned long long*>(err) != 0) break;
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > timeout_ns) {
        atomicCAS(err, 0ull, 1ull + (unsigned long long)(p - reinterpret_cast<const unsigned long long*>(mine)));
        __threadfence();
        break;
      }
    }
  }
  return v;
}

struct GatherArgs {
  char* const* peer;            // device array [world]
  int world, rank, n_seg;
  unsigned long long timeout_ns;
  const uint4* src[kMaxSeg];    // this rank's shard of tensor s
  unsigned long long n16[kMaxSeg];      // shard size in 16-byte units
  unsigned long long dst_off[kMaxSeg];  // arena offset of the gathered tensor
};

__global__ void __launch_bounds__(256) k_p2p_allgather(GatherArgs a) {
  char* mine = a.peer[a.rank];
  unsigned long long* ctr = reinterpret_cast<unsigned long long*>(mine + kCtrGather);
  const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(ctr) + 1;
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int s = 0; s < a.n_seg; ++s) {
    const size_t n = a.n16[s];
    for (size_t e = t0; e < n * a.world; e += stride) {
      const int p = (int)(e / n);
      const size_t c = e - (size_t)p * n;
      reinterpret_cast<uint4*>(a.peer[p] + a.dst_off[s])[(size_t)a.rank * n + c] = a.src[s][c];
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(ctr + 1, 1ull) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  // every block of this rank has stored (and fenced) its part: raise the flag everywhere, wait for the world
  __threadfence_system();
  if ((int)threadIdx.x < a.world) {
    st_release_sys(reinterpret_cast<unsigned long long*>(a.peer[threadIdx.x] + kFlagGather) + a.rank, epoch);
    wait_flag(mine, a.timeout_ns, reinterpret_cast<const unsigned long long*>(mine + kFlagGather) + threadIdx.x, epoch);
  }
  __syncthreads();
  if (threadIdx.x == 0) { ctr[1] = 0; ctr[0] = epoch; __threadfence(); }
}

__global__ void __launch_bounds__(64) k_p2p_sum_scalar(char* const* peer, int world, int rank, unsigned long long timeout_ns,
                                                       const float* __restrict__ value, float* __restrict__ out) {
  char* mine = peer[rank];
  unsigned long long* ctr = reinterpret_cast<unsigned long long*>(mine + kCtrScalar);
  const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(ctr) + 1;
  __shared__ float vals[kMaxWorld];
  if ((int)threadIdx.x < world) {
    const unsigned long long packed = (epoch << 32) | (unsigned long long)__float_as_uint(*value);
    st_release_sys(reinterpret_cast<unsigned long long*>(peer[threadIdx.x] + kSlotScalar) + rank, packed);
    const unsigned long long got =
        wait_flag(mine, timeout_ns, reinterpret_cast<const unsigned long long*>(mine + kSlotScalar) + threadIdx.x, epoch, ~0ull, 32);
    vals[threadIdx.x] = __uint_as_float((unsigned int)(got & 0xffffffffull));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int p = 0; p < world; ++p) t += vals[p];     // rank order: the same bits on every rank
    *out = t;
    ctr[0] = epoch;
    __threadfence();
  }
}

template <bool BF16>
__global__ void __launch_bounds__(256) k_p2p_reduce_scatter(char* const* peer, int world, int rank,
                                                            unsigned long long timeout_ns,
                                                            unsigned long long src_off, unsigned long long shard4,
                                                            void* __restrict__ out) {
  char* mine = peer[rank];
  unsigned long long* ctr = reinterpret_cast<unsigned long long*>(mine + kCtrReduce);
  const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(ctr) + 1;
  // the partial sums of this rank were completed by the previous kernels of this stream
  if (blockIdx.x == 0 && (int)threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned long long*>(peer[threadIdx.x] + kFlagReduce) + rank, epoch);
  }
  if ((int)threadIdx.x < world)
    wait_flag(mine, timeout_ns, reinterpret_cast<const unsigned long long*>(mine + kFlagReduce) + threadIdx.x, epoch);
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < shard4; e += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < world; ++p) {
      const float4 v = __ldcv(reinterpret_cast<const float4*>(peer[p] + src_off) + (size_t)rank * shard4 + e);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (BF16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
      reinterpret_cast<uint2*>(out)[e] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    } else {
      reinterpret_cast<float4*>(out)[e] = acc;
    }
  }
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(ctr + 1, 1ull) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) { ctr[1] = 0; ctr[0] = epoch; __threadfence(); }
}
}  // namespace p2p

// wall-time bound of every flag wait: DVLP_P2P_TIMEOUT_S seconds (default 600; a rank may legitimately be late by minutes)
static unsigned long long p2p_timeout_ns() {
  static const unsigned long long ns = [] {
    const char* e = getenv("DVLP_P2P_TIMEOUT_S");
    double sec = e ? atof(e) : 600.0;
    if (!(sec > 0.0)) sec = 600.0;
    return (unsigned long long)(sec * 1e9);
  }();
  return ns;
}

int run_p2p_allgather(void* const* peer, int world, int rank, int n_seg, const void* const* src_host,
                      const size_t* bytes_host, const size_t* dst_off_host, cudaStream_t st) {
  p2p::GatherArgs a{};
  a.peer = (char* const*)peer; a.world = world; a.rank = rank; a.n_seg = n_seg;
  a.timeout_ns = p2p_timeout_ns();
  size_t total16 = 0;
  for (int s = 0; s < n_seg; ++s) {
    a.src[s] = (const uint4*)src_host[s];
    a.n16[s] = bytes_host[s] / 16;
    a.dst_off[s] = dst_off_host[s];
    total16 += a.n16[s] * world;
  }
  int grid = (int)((total16 + 256 * 8 - 1) / (256 * 8));
  grid = grid < 1 ? 1 : (grid > 296 ? 296 : grid);
  p2p::k_p2p_allgather<<<grid, 256, 0, st>>>(a);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

int run_p2p_sum_scalar(void* const* peer, int world, int rank, const float* value, float* out, cudaStream_t st) {
  p2p::k_p2p_sum_scalar<<<1, 64, 0, st>>>((char* const*)peer, world, rank, p2p_timeout_ns(), value, out);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

int run_p2p_reduce_scatter(void* const* peer, int world, int rank, size_t src_off, size_t shard_floats, void* out,
                           int out_dtype, cudaStream_t st) {
  const size_t shard4 = shard_floats / 4;
  int grid = (int)((shard4 + 256 * 4 - 1) / (256 * 4));
  grid = grid < 1 ? 1 : (grid > 296 ? 296 : grid);
  if (out_dtype == DVLP_BF16)
    p2p::k_p2p_reduce_scatter<true><<<grid, 256, 0, st>>>((char* const*)peer, world, rank, p2p_timeout_ns(), src_off, shard4, out);
  else
    p2p::k_p2p_reduce_scatter<false><<<grid, 256, 0, st>>>((char* const*)peer, world, rank, p2p_timeout_ns(), src_off, shard4, out);
  DVLP_LAUNCH_OK();
  return DVLP_OK;
}

// reads the sticky error word of THIS rank's arena (synchronises `st`): 0 healthy, else DVLP_ERR_TIMEOUT
int run_p2p_check(const void* own_arena, cudaStream_t st) {
  unsigned long long w = 0;
  DVLP_CUDA_OK(cudaMemcpyAsync(&w, (const char*)own_arena + p2p::kErrWord, sizeof(w), cudaMemcpyDeviceToHost, st));
  DVLP_CUDA_OK(cudaStreamSynchronize(st));
  return w == 0 ? DVLP_OK : DVLP_ERR_TIMEOUT;
}

}  // namespace dvlp
