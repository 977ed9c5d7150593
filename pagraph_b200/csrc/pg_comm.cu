// pg_comm.cu — gradient all-reduce fused with the Adam step, over NVLink peer memory.
//
// Reference: the only cross-GPU traffic of the trainer is DistributedDataParallel's all-reduce of the 23 k gradient
// floats (92 KB), followed by optimizer.step() (examples/profile/pa_gcn.py:65,96-97). At that size the collective is pure
// latency: NCCL costs ~40-50 us per step inside the captured compute graph, plus a division and the optimizer kernel.
// Here one kernel does all of it:
//   push    every CTA owns one slice of the flat gradient; it stores the slice into every peer's receive area over NVLink
//           (peer pointers from CUDA IPC), fences at system scope and raises a per-slice flag on the peer;
//   reduce  the same CTA spins (acquire, system scope) on its own flags until every peer's slice for this step has
//           landed, sums the slices in rank order (every rank computes the same bits), divides by the world size;
//   update  and applies Adam to its slice of the parameters / moments in place.
// No grid-wide or host synchronisation: a slice is independent end to end. Receive areas are double-buffered by step
// parity: a rank can run at most one step ahead of the slowest peer (it needs that peer's push to finish its own
// reduce), so a buffer is never overwritten before it has been consumed. With world == 1 the kernel is just the
// optimizer step. The flag values are the (monotonic) step number, so nothing is ever reset and the launch can be
// captured in a CUDA graph; the step number is read from device memory.
#include <algorithm>
#include <cstring>

#include "pg_common.cuh"

namespace {

constexpr int kCommThreads = 256;
constexpr int kCommMaxCtas = 64;

struct PeerArgs {
  int world, rank;
  float* recv[PG_MAX_RANKS];            // recv[p]: base of rank p's receive area [2][world][n_pad] (mapped here)
  unsigned long long* flags[PG_MAX_RANKS];  // flags[p]: base of rank p's flags [2][world][kCommMaxCtas]
  int64_t n, n_pad;
};

struct AdamArgs {
  float* param;
  float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  const float* step;       // optimizer step count (already incremented for this step), device scalar
  float lr, beta1, beta2, eps, weight_decay;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_cg(const float* p) {   // bypass L1: the line was written by a peer
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(kCommThreads) allreduce_adam_kernel(PeerArgs pa, AdamArgs ad, const int64_t* step_id_ptr) {
  const int64_t per = (pa.n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = min(pa.n, lo + per);
  const unsigned long long step_id = (unsigned long long)*step_id_ptr;
  const int parity = (int)(step_id & 1);
  const int W = pa.world, me = pa.rank;
  if (W > 1) {
    // ---- push my slice to every peer's receive area [parity][me]
    for (int p = 0; p < W; ++p) {
      if (p == me) continue;
      float* dst = pa.recv[p] + ((size_t)parity * W + me) * pa.n_pad;
      for (int64_t i = lo + threadIdx.x; i < hi; i += kCommThreads) dst[i] = ad.grad[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < W && threadIdx.x != me)
      st_release_sys(pa.flags[threadIdx.x] + ((size_t)parity * W + me) * kCommMaxCtas + blockIdx.x, step_id);
    // ---- wait for every peer's slice of this step
    if (threadIdx.x < W && threadIdx.x != me) {
      const unsigned long long* f = pa.flags[me] + ((size_t)parity * W + threadIdx.x) * kCommMaxCtas + blockIdx.x;
      while (ld_acquire_sys(f) < step_id) {
      }
    }
    __syncthreads();
  }
  // ---- reduce in rank order + Adam
  const float step = *ad.step;
  const float bc1 = 1.0f - powf(ad.beta1, step), bc2 = 1.0f - powf(ad.beta2, step);
  const float step_size = ad.lr / bc1, bc2_sqrt = sqrtf(bc2), inv_w = 1.0f / (float)W;
  const float* mine = pa.recv[me];
  for (int64_t i = lo + threadIdx.x; i < hi; i += kCommThreads) {
    float g = 0.f;
    for (int r = 0; r < W; ++r) g += (r == me) ? ad.grad[i] : ld_cg(mine + ((size_t)parity * W + r) * pa.n_pad + i);
    g *= inv_w;
    ad.grad[i] = g;                                   // the averaged gradient stays visible to the caller
    float p = ad.param[i];
    if (ad.weight_decay != 0.f) g += ad.weight_decay * p;
    const float m = ad.exp_avg[i] + (g - ad.exp_avg[i]) * (1.0f - ad.beta1);
    const float v = ad.exp_avg_sq[i] * ad.beta2 + g * g * (1.0f - ad.beta2);
    ad.exp_avg[i] = m;
    ad.exp_avg_sq[i] = v;
    ad.param[i] = p - step_size * m / (sqrtf(v) / bc2_sqrt + ad.eps);
  }
}

}  // namespace

struct pg_peer_group {
  PeerArgs args;
  int dev = 0, ctas = 0;
  void* local = nullptr;                 // my region (cudaMalloc)
  void* opened[PG_MAX_RANKS] = {nullptr};
  size_t region_bytes = 0;
};

static size_t region_layout(int world, int64_t n_pad, size_t* flags_off) {
  const size_t recv_bytes = (size_t)2 * world * n_pad * sizeof(float);
  *flags_off = (recv_bytes + 255) / 256 * 256;
  return *flags_off + (size_t)2 * world * kCommMaxCtas * sizeof(unsigned long long);
}

extern "C" {

pg_status pg_peer_group_create(int world, int rank, int64_t n, int dev, pg_peer_group** out, unsigned char* handle_out) {
  PG_REQUIRE(out && handle_out && world >= 1 && world <= PG_MAX_RANKS && rank >= 0 && rank < world && n >= 1,
             "pg_peer_group_create: bad arguments");
  pg::DeviceGuard guard(dev);
  pg_peer_group* g = new pg_peer_group;
  memset(&g->args, 0, sizeof(g->args));
  g->dev = dev;
  g->args.world = world;
  g->args.rank = rank;
  g->args.n = n;
  g->args.n_pad = (n + 63) / 64 * 64;
  g->ctas = (int)std::min<int64_t>(kCommMaxCtas, std::max<int64_t>(1, (n + 511) / 512));  // 2 elements per thread
  size_t flags_off = 0;
  g->region_bytes = region_layout(world, g->args.n_pad, &flags_off);
  if (cudaMalloc(&g->local, g->region_bytes) != cudaSuccess) {
    cudaGetLastError();
    delete g;
    pg::set_error("pg_peer_group_create: out of device memory");
    return PG_ERR_NOMEM;
  }
  PG_CUDA(cudaMemset(g->local, 0, g->region_bytes));
  g->args.recv[rank] = (float*)g->local;
  g->args.flags[rank] = (unsigned long long*)((char*)g->local + flags_off);
  cudaIpcMemHandle_t h;
  memset(&h, 0, sizeof(h));
  if (world > 1) PG_CUDA(cudaIpcGetMemHandle(&h, g->local));
  static_assert(sizeof(h) == PG_IPC_HANDLE_BYTES, "CUDA IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  *out = g;
  return PG_OK;
}

pg_status pg_peer_group_connect(pg_peer_group* g, const unsigned char* handles /* [world][PG_IPC_HANDLE_BYTES] */) {
  PG_REQUIRE(g && handles, "pg_peer_group_connect: bad arguments");
  pg::DeviceGuard guard(g->dev);
  size_t flags_off = 0;
  region_layout(g->args.world, g->args.n_pad, &flags_off);
  for (int p = 0; p < g->args.world; ++p) {
    if (p == g->args.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)p * PG_IPC_HANDLE_BYTES, sizeof(h));
    void* ptr = nullptr;
    PG_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    g->opened[p] = ptr;
    g->args.recv[p] = (float*)ptr;
    g->args.flags[p] = (unsigned long long*)((char*)ptr + flags_off);
  }
  return PG_OK;
}

void pg_peer_group_destroy(pg_peer_group* g) {
  if (!g) return;
  pg::DeviceGuard guard(g->dev);
  cudaDeviceSynchronize();
  for (int p = 0; p < PG_MAX_RANKS; ++p)
    if (g->opened[p]) cudaIpcCloseMemHandle(g->opened[p]);
  cudaFree(g->local);
  cudaGetLastError();
  delete g;
}

/* ---- plain peer-visible device buffers (the peer-GPU cache tier's tables) */
pg_status pg_peer_alloc(size_t bytes, int dev, void** d_out, unsigned char* handle_out) {
  PG_REQUIRE(d_out && handle_out && bytes > 0, "pg_peer_alloc: bad arguments");
  pg::DeviceGuard guard(dev);
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    pg::set_error("pg_peer_alloc: out of device memory (%zu bytes)", bytes);
    return PG_ERR_NOMEM;
  }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
    pg::set_error("pg_peer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(p);
    return PG_ERR_CUDA;
  }
  memcpy(handle_out, &h, sizeof(h));
  *d_out = p;
  return PG_OK;
}

pg_status pg_peer_open(const unsigned char* handle, int dev, void** d_out) {
  PG_REQUIRE(handle && d_out, "pg_peer_open: bad arguments");
  pg::DeviceGuard guard(dev);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  PG_CUDA(cudaIpcOpenMemHandle(d_out, h, cudaIpcMemLazyEnablePeerAccess));
  return PG_OK;
}

pg_status pg_peer_close(void* d_ptr, int dev) {
  if (!d_ptr) return PG_OK;
  pg::DeviceGuard guard(dev);
  PG_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return PG_OK;
}

pg_status pg_peer_free(void* d_ptr, int dev) {
  if (!d_ptr) return PG_OK;
  pg::DeviceGuard guard(dev);
  PG_CUDA(cudaFree(d_ptr));
  return PG_OK;
}

pg_status pg_allreduce_adam(pg_peer_group* g, float* d_param, float* d_grad, float* d_exp_avg, float* d_exp_avg_sq,
                            const float* d_step, const int64_t* d_step_id, float lr, float beta1, float beta2, float eps,
                            float weight_decay, void* stream) {
  PG_REQUIRE(g && d_param && d_grad && d_exp_avg && d_exp_avg_sq && d_step && d_step_id, "pg_allreduce_adam: null argument");
  for (int p = 0; p < g->args.world; ++p)
    PG_REQUIRE(g->args.recv[p] != nullptr, "pg_allreduce_adam: peer group is not connected");
  pg::DeviceGuard guard(g->dev);
  AdamArgs ad{d_param, d_grad, d_exp_avg, d_exp_avg_sq, d_step, lr, beta1, beta2, eps, weight_decay};
  pg::TimedScope timed(PG_T_OPT, (cudaStream_t)stream);
  pg::prefer_max_smem_k(allreduce_adam_kernel);
  allreduce_adam_kernel<<<g->ctas, kCommThreads, 0, (cudaStream_t)stream>>>(g->args, ad, d_step_id);
  PG_CHECK_LAUNCH();
  return PG_OK;
}

}  // extern "C"
