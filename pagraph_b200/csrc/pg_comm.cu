// pg_comm.cu — gradient all-reduce fused with the Adam step, over NVLink peer memory.
//
// Reference: the only cross-GPU traffic of the trainer is DistributedDataParallel's all-reduce of the 23 k gradient
// floats (92 KB), followed by optimizer.step() (examples/profile/pa_gcn.py:65,96-97). At that size the collective is pure
// latency: NCCL costs ~40-50 us per step inside the captured compute graph, plus a division and the optimizer kernel.
// Here one kernel does all of it, and every THREAD is independent end to end (no fence, no barrier, no separate flag):
//   push    a thread owns pairs of gradient floats; it stores each pair into every peer's receive area over NVLink
//           (peer pointers from CUDA IPC) as ONE 16-byte line {g0, step, g1, step}: the data carries its own flag, so a
//           line that reads back with both step words equal to this step's number is complete (8-byte halves are
//           written atomically) — the protocol NCCL calls LL. One NVLink store latency, no system-scope fence, no
//           second round trip for a flag;
//   reduce  the same thread polls the W-1 lines of its pairs in its own receive area until they carry this step's
//           number, sums them in rank order (every rank computes the same bits), divides by the world size;
//   update  and applies Adam to its elements of the parameters / moments in place.
// Receive areas are double-buffered by step parity: a rank can run at most one step ahead of the slowest peer (it needs
// that peer's push of step s+1 to finish its own step s+1, and the peer pushes s+1 only after its step-s kernel has
// consumed the step-s lines), so a line is never overwritten before it has been read. Step numbers are monotonic (>= 1,
// the areas start zeroed), so nothing is ever reset and the launch can be captured in a CUDA graph; the step number is
// read from device memory. With world == 1 the kernel is just the optimizer step.
#include <algorithm>
#include <cstring>

#include "pg_common.cuh"

namespace {

constexpr int kCommThreads = 128;
constexpr int kCommMaxCtas = 148;
constexpr int kCommPairs = 2;   // pairs per thread in flight between the push and the poll

struct PeerArgs {
  int world, rank;
  uint4* recv[PG_MAX_RANKS];   // recv[p]: base of rank p's receive area [2][world][n_pad / 2] lines (mapped here)
  int64_t n, n_pad;
};

struct AdamArgs {
  float* param;
  float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  float* step;             // optimizer step count, device scalar (already incremented for this step unless `advance`)
  float lr, beta1, beta2, eps, weight_decay;
};

__device__ __forceinline__ void st_line(uint4* p, uint32_t a, uint32_t b, uint32_t flag) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(flag), "r"(b), "r"(flag) : "memory");
}
__device__ __forceinline__ uint4 ld_line(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void adam_one(const AdamArgs& ad, int64_t i, float g, float step_size, float bc2_sqrt) {
  ad.grad[i] = g;                                   // the averaged gradient stays visible to the caller
  float p = ad.param[i];
  if (ad.weight_decay != 0.f) g += ad.weight_decay * p;
  const float m = ad.exp_avg[i] + (g - ad.exp_avg[i]) * (1.0f - ad.beta1);
  const float v = ad.exp_avg_sq[i] * ad.beta2 + g * g * (1.0f - ad.beta2);
  ad.exp_avg[i] = m;
  ad.exp_avg_sq[i] = v;
  ad.param[i] = p - step_size * m / (sqrtf(v) / bc2_sqrt + ad.eps);
}

// advance != 0: the two step counters hold the counts BEFORE this step; the kernel works with count + 1 and the last CTA to
// finish (ticket) writes the new counts back — every CTA has read them by then. Saves the two one-element increment
// kernels that would otherwise sit in front of this one on the critical path of the training step.
__global__ void __launch_bounds__(kCommThreads) allreduce_adam_kernel(PeerArgs pa, AdamArgs ad, int64_t* step_id_ptr, int advance,
                                                                      unsigned* ticket) {
  const unsigned long long step_id = (unsigned long long)*step_id_ptr + (advance ? 1ull : 0ull);
  const uint32_t flag = (uint32_t)step_id;
  const int parity = (int)(step_id & 1);
  const int W = pa.world, me = pa.rank;
  const int64_t lines = pa.n_pad / 2;                      // one line = two gradient floats
  const float step = *ad.step + (advance ? 1.0f : 0.0f);
  const float bc1 = 1.0f - powf(ad.beta1, step), bc2 = 1.0f - powf(ad.beta2, step);
  const float step_size = ad.lr / bc1, bc2_sqrt = sqrtf(bc2), inv_w = 1.0f / (float)W;
  const int64_t nth = (int64_t)gridDim.x * kCommThreads;
  for (int64_t j0 = (int64_t)blockIdx.x * kCommThreads + threadIdx.x; j0 < lines; j0 += nth * kCommPairs) {
    float g0[kCommPairs], g1[kCommPairs];
#pragma unroll
    for (int k = 0; k < kCommPairs; ++k) {
      const int64_t j = j0 + k * nth, i = 2 * j;
      g0[k] = (j < lines && i < pa.n) ? ad.grad[i] : 0.f;
      g1[k] = (j < lines && i + 1 < pa.n) ? ad.grad[i + 1] : 0.f;
    }
    if (W > 1) {
      // ---- push my lines to every peer's receive area [parity][me], nearest-higher rank first (spreads the links)
#pragma unroll
      for (int k = 0; k < kCommPairs; ++k) {
        const int64_t j = j0 + k * nth;
        if (j >= lines) continue;
        for (int d = 1; d < W; ++d) {
          const int p = (me + d) % W;
          st_line(pa.recv[p] + ((size_t)parity * W + me) * lines + j, __float_as_uint(g0[k]), __float_as_uint(g1[k]), flag);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kCommPairs; ++k) {
      const int64_t j = j0 + k * nth, i = 2 * j;
      if (j >= lines) continue;
      float s0 = 0.f, s1 = 0.f;
      if (W > 1) {
        // ---- poll the peers' lines of this step, sum in rank order
        const uint4* mine = pa.recv[me] + (size_t)parity * W * lines + j;
        for (int r = 0; r < W; ++r) {
          if (r == me) {
            s0 += g0[k];
            s1 += g1[k];
            continue;
          }
          uint4 v = ld_line(mine + (size_t)r * lines);
          while (v.y != flag || v.w != flag) v = ld_line(mine + (size_t)r * lines);
          s0 += __uint_as_float(v.x);
          s1 += __uint_as_float(v.z);
        }
        s0 *= inv_w;
        s1 *= inv_w;
      } else {
        s0 = g0[k];
        s1 = g1[k];
      }
      if (i < pa.n) adam_one(ad, i, s0, step_size, bc2_sqrt);
      if (i + 1 < pa.n) adam_one(ad, i + 1, s1, step_size, bc2_sqrt);
    }
  }
  if (advance) {
    __syncthreads();                                       // every thread of the CTA has read the counters
    if (threadIdx.x == 0 && atomicAdd(ticket, 1u) == gridDim.x - 1) {
      *ticket = 0;                                         // re-armed for the next launch (stream-ordered behind this one)
      *step_id_ptr = (int64_t)step_id;
      *ad.step = step;
    }
  }
}

}  // namespace

struct pg_peer_group {
  PeerArgs args;
  int dev = 0, ctas = 0;
  void* local = nullptr;                 // my region (cudaMalloc)
  void* opened[PG_MAX_RANKS] = {nullptr};
  size_t region_bytes = 0;
  unsigned* ticket = nullptr;            // CTA ticket of the counter-advancing launch (zero at rest)
};

static size_t region_bytes_for(int world, int64_t n_pad) { return (size_t)2 * world * (n_pad / 2) * sizeof(uint4); }

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
  const int64_t per_cta = (int64_t)2 * kCommPairs * kCommThreads;   // gradient floats per CTA and pass
  g->ctas = (int)std::min<int64_t>(kCommMaxCtas, std::max<int64_t>(1, (g->args.n_pad + per_cta - 1) / per_cta));
  g->region_bytes = region_bytes_for(world, g->args.n_pad);
  if (cudaMalloc(&g->local, g->region_bytes) != cudaSuccess) {
    cudaGetLastError();
    delete g;
    pg::set_error("pg_peer_group_create: out of device memory");
    return PG_ERR_NOMEM;
  }
  PG_CUDA(cudaMemset(g->local, 0, g->region_bytes));
  PG_CUDA(cudaMalloc((void**)&g->ticket, 256));
  PG_CUDA(cudaMemset(g->ticket, 0, 256));
  g->args.recv[rank] = (uint4*)g->local;
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
  for (int p = 0; p < g->args.world; ++p) {
    if (p == g->args.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)p * PG_IPC_HANDLE_BYTES, sizeof(h));
    void* ptr = nullptr;
    PG_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    g->opened[p] = ptr;
    g->args.recv[p] = (uint4*)ptr;
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
  cudaFree(g->ticket);
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

static pg_status allreduce_adam_launch(pg_peer_group* g, float* d_param, float* d_grad, float* d_exp_avg, float* d_exp_avg_sq,
                                       float* d_step, int64_t* d_step_id, float lr, float beta1, float beta2, float eps,
                                       float weight_decay, int advance, void* stream) {
  PG_REQUIRE(g && d_param && d_grad && d_exp_avg && d_exp_avg_sq && d_step && d_step_id, "pg_allreduce_adam: null argument");
  for (int p = 0; p < g->args.world; ++p)
    PG_REQUIRE(g->args.recv[p] != nullptr, "pg_allreduce_adam: peer group is not connected");
  pg::DeviceGuard guard(g->dev);
  AdamArgs ad{d_param, d_grad, d_exp_avg, d_exp_avg_sq, d_step, lr, beta1, beta2, eps, weight_decay};
  pg::TimedScope timed(PG_T_OPT, (cudaStream_t)stream);
  pg::prefer_max_smem_k(allreduce_adam_kernel);
  allreduce_adam_kernel<<<g->ctas, kCommThreads, 0, (cudaStream_t)stream>>>(g->args, ad, d_step_id, advance, g->ticket);
  PG_CHECK_LAUNCH();
  return PG_OK;
}

pg_status pg_allreduce_adam(pg_peer_group* g, float* d_param, float* d_grad, float* d_exp_avg, float* d_exp_avg_sq,
                            const float* d_step, const int64_t* d_step_id, float lr, float beta1, float beta2, float eps,
                            float weight_decay, void* stream) {
  return allreduce_adam_launch(g, d_param, d_grad, d_exp_avg, d_exp_avg_sq, (float*)d_step, (int64_t*)d_step_id, lr, beta1, beta2,
                               eps, weight_decay, 0, stream);
}

pg_status pg_allreduce_adam_next(pg_peer_group* g, float* d_param, float* d_grad, float* d_exp_avg, float* d_exp_avg_sq,
                                 float* d_step, int64_t* d_step_id, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, void* stream) {
  return allreduce_adam_launch(g, d_param, d_grad, d_exp_avg, d_exp_avg_sq, d_step, d_step_id, lr, beta1, beta2, eps, weight_decay,
                               1, stream);
}

}  // extern "C"
