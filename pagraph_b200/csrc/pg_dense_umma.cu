// pg_dense_umma.cu — the first NodeUpdate's forward product on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Reference: PaGraph/model/gcn_nssc.py:14-24 (z = Linear(x); out = cat(z, relu z)) and :66-67 (dropout ahead of the next
// block_compute). x [n, K] is the aggregated input block (K = 600: 85 MB per minibatch at config 2), W [32, K].
//
// fp32-level accuracy on a TF32 tensor core = error-compensated product (3xTF32): x = x_hi + x_lo, W = W_hi + W_lo with the
// hi parts being what `kind::tf32` keeps of a 32-bit container (it ignores the low 13 mantissa bits), and
//     z = x_lo W_hi^T + x_hi W_lo^T + x_hi W_hi^T   (fp32 accumulate in TMEM; the x_lo W_lo term is 2^-22 relative).
//
// One CTA per SM, 128-row tiles, K walked in chunks of 32 columns (one 128-byte swizzle row):
//   warp 0      TMA producer: per chunk one [128 x 32] box of x and one [32 x 32] box of W (SWIZZLE_128B, OOB -> 0) into
//               a kStages-deep shared-memory ring
//   warps 8-11  transform, one tile row per thread: reads its 32 floats of x from the ring (swizzle-aware, conflict-free),
//               splits them and writes x_hi / x_lo into TENSOR MEMORY (tcgen05.st) — the A operand of the MMAs is read
//               from TMEM, not from shared memory: an SS-mode M = 128 tf32 MMA would fetch 4 KB of A per instruction and
//               run at the shared-memory port's 128 B/clk (~32 clk) instead of the tensor core's 16 clk; it also splits
//               the chunk's W box in place next to it (W_lo plane)
//   warp 1      MMA issuer (one thread): per chunk 4 k-steps x 2 instructions of tcgen05.mma.cta_group::1.kind::tf32 (M 128,
//               K 8), A from TMEM, B from shared memory: x_hi against [W_hi ; W_lo] with N = 64 (two of the three products
//               in one instruction — an instruction with its A operand in TMEM is paced by that operand's read, not by N)
//               and x_lo against W_hi with N = 32; tcgen05.commit releases the ring slot
//   warps 4-7   epilogue, one output row per thread: tcgen05.ld of the 64 accumulator columns, the two halves summed,
//               + bias, relu, concat, dropout mask (pg_common.cuh drop_hash contract); the tile is staged as [128 x 32]
//               SWIZZLE_128B boxes and stored with one TMA tensor store per box; double-buffered accumulators
// Measured (ncu, profiles/r2_dense_kernel_timing.md): 6.9 us fixed + 8.8 us per wave of 148 tiles — 0.79 of the measured HBM
// peak while streaming; 24.4 us at config 2 (HBM floor 13.5 us).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "pg_common.cuh"

namespace {

constexpr int kBlockM = 128;       // rows per tile (UMMA M)
constexpr int kN = 32;             // outputs (UMMA N)
constexpr int kChunk = 32;         // K per stage: 32 floats = 128 bytes = one swizzle row
constexpr int kStages = 6;
constexpr int kUmmaK = 8;          // K per tcgen05.mma for 32-bit operands
constexpr int kThreads = 12 * 32;
constexpr uint32_t kXBytes = kBlockM * kChunk * 4;            // 16 KB
constexpr uint32_t kWBytes = kN * kChunk * 4;                 // 4 KB
constexpr uint32_t kStageBytes = kXBytes + 2 * kWBytes;       // x, w_hi, w_lo = 24 KB
constexpr uint32_t kEpiBoxBytes = kBlockM * kN * 4;            // one [128 rows x 32 floats] store box (SWIZZLE_128B) = 16 KB
constexpr uint32_t kEpiBytes = 4 * kEpiBoxBytes;              // z / relu z halves of out and of out_drop = 64 KB
constexpr uint32_t kAccStage = 2 * kN;                        // accumulator stage: [x_hi W_hi + x_lo W_hi | x_hi W_lo]
constexpr uint32_t kAccCols = 2 * kAccStage;                  // 2 accumulator stages
constexpr uint32_t kTmemCols = 512;                           // 128 accumulator + kStages * 64 operand columns (power of 2)
static_assert(kAccCols + kStages * 2 * kChunk <= kTmemCols, "TMEM budget");

using pg::smem_u32;
using pg::mbar_init;
using pg::mbar_expect_tx;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(src)
               : "memory");
}
// K-major SWIZZLE_128B operand tile (rows of 128 bytes, 8-row groups 1024 bytes apart, tile base 1024-byte aligned):
// cute::UMMA::SmemDescriptor with version 1, layout_type 2, LBO 1, SBO 64 (both in 16-byte units)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)64 << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// cute::UMMA::InstrDescriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), K-major both, N >> 3 at 17, M >> 4 at 24
constexpr uint32_t umma_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}
constexpr uint32_t kIdesc = umma_idesc(kN);            // N = 32
constexpr uint32_t kIdesc2 = umma_idesc(2 * kN);       // N = 64: B = [hi plane ; lo plane], two products per instruction

// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate,
                                             uint32_t idesc = kIdesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {   // arrives when all prior MMAs of this thread are done
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ uint32_t tf32_lo(uint32_t bits) {  // x - (what kind::tf32 keeps of x)
  return __float_as_uint(__uint_as_float(bits) - __uint_as_float(bits & 0xFFFFE000u));
}

struct UmmaDrop {
  uint32_t thr;        // drop when the 16-bit hash lane < thr (0 = no dropout)
  float scale;         // 1 / (1 - p)
  uint64_t seed;
  const int64_t* step; // optional device counter added to the seed
};

__global__ void __launch_bounds__(kThreads, 1)
    linear_concat_fwd_umma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                                  const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_drop,
                                  const float* __restrict__ bias, int64_t n, int K, int concat, int dropping, UmmaDrop drop) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[3 * kStages + 4];   // full[s], ready[s], empty[s], tmem_full[2], tmem_empty[2]
  __shared__ uint32_t tmem_base_sh;
  __shared__ uint64_t colkey_sh[16];
  __shared__ float bias_sh[kN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + kStages * kStageBytes;   // epilogue staging rows behind the ring
  auto full = [&](int s) { return smem_u32(&bars[s]); };
  auto ready = [&](int s) { return smem_u32(&bars[kStages + s]); };
  auto empty = [&](int s) { return smem_u32(&bars[2 * kStages + s]); };
  auto tfull = [&](int a) { return smem_u32(&bars[3 * kStages + a]); };
  auto tempty = [&](int a) { return smem_u32(&bars[3 * kStages + 2 + a]); };
  const int nchunks = (K + kChunk - 1) / kChunk;
  const int64_t ntiles = (n + kBlockM - 1) / kBlockM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full(s), 1);         // the producer's arrive.expect_tx; the TMA completes the transaction bytes
      mbar_init(ready(s), 128);      // every transform thread
      mbar_init(empty(s), 1);        // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull(a), 1);        // tcgen05.commit after the tile's last MMA
      mbar_init(tempty(a), 128);     // every epilogue thread
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 16) colkey_sh[threadIdx.x] = pg::drop_colkey((uint32_t)threadIdx.x);
  if (threadIdx.x >= 32 && threadIdx.x < 32 + kN) bias_sh[threadIdx.x - 32] = bias ? bias[threadIdx.x - 32] : 0.f;
  if (warp == 1) {                   // TMEM: allocated and later freed by the same warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_sh;
  const uint32_t tmem_opnd = tmem_base + kAccCols;    // operand stages behind the two accumulators

  if (warp == 0) {
    // ================================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int c = 0; c < nchunks; ++c, ++it) {
          const int s = it % kStages;
          mbar_wait(empty(s), ((it / kStages) & 1) ^ 1);             // a fresh barrier passes the wait for parity 1
          const uint32_t st = smem_base + s * kStageBytes;
          mbar_expect_tx(full(s), kXBytes + kWBytes);
          tma_load_2d(st, &tm_x, c * kChunk, (int)(tile * kBlockM), full(s));       // x (OOB rows / cols -> 0)
          tma_load_2d(st + kXBytes, &tm_w, c * kChunk, 0, full(s));                 // W chunk
        }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      uint32_t it = 0, tl = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
        const int a = tl & 1;
        mbar_wait(tempty(a), ((tl >> 1) & 1) ^ 1);                   // epilogue has drained this accumulator stage
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + a * kAccStage;
        for (int c = 0; c < nchunks; ++c, ++it) {
          const int s = it % kStages;
          mbar_wait(ready(s), (it / kStages) & 1);                   // x_hi / x_lo in TMEM, W_lo in shared memory
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_base + s * kStageBytes;
          const uint32_t a_hi = tmem_opnd + s * 2 * kChunk, a_lo = a_hi + kChunk;
          // An MMA with the A operand in tensor memory is paced by that operand's read (4 KB per instruction: ~64 clk in the
          // r2l profile, whatever N is), so the two products of x_hi go out as ONE instruction with N = 64 — its B tile is
          // the W_hi box and the W_lo plane behind it, 64 rows of 128 bytes — into columns [x W_hi | x W_lo] of the stage;
          // x_lo W_hi follows with N = 32 into the first 32. The epilogue adds the two halves.
          const uint64_t b_hi = umma_desc_k_sw128(st + kXBytes);
#pragma unroll
          for (int k = 0; k < kChunk / kUmmaK; ++k) {
            const uint64_t adv = (uint64_t)((k * kUmmaK * 4) >> 4);  // 32 bytes per k-step inside the 128-byte swizzle row
            umma_tf32_ts(d, a_hi + k * kUmmaK, b_hi + adv, (c | k) != 0, kIdesc2);
            umma_tf32_ts(d, a_lo + k * kUmmaK, b_hi + adv, 1, kIdesc);
          }
          umma_commit(empty(s));                                     // ring slot + TMEM operand stage reusable
        }
        umma_commit(tfull(a));                                       // accumulator complete
      }
    }
  } else if (warp >= 8) {
    // ================================================================== transform: split x into TMEM, W_lo in place
    const int q = warp & 3;                                          // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;                                   // tile row owned by this thread
    const int tt = threadIdx.x - 8 * 32;                             // 0 .. 127
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      for (int c = 0; c < nchunks; ++c, ++it) {
        const int s = it % kStages;
        mbar_wait(full(s), (it / kStages) & 1);
        const uint32_t st = smem_base + s * kStageBytes;
        // row `row` of the box: 8 x 16 bytes, logical unit u stored at unit u ^ (row & 7) (SWIZZLE_128B)
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t addr = st + (uint32_t)row * 128u + (uint32_t)((u ^ (row & 7)) << 4);
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(hi[4 * u]), "=r"(hi[4 * u + 1]), "=r"(hi[4 * u + 2]), "=r"(hi[4 * u + 3])
                       : "r"(addr));
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) lo[j] = tf32_lo(hi[j]);
        const uint32_t ta = tmem_opnd + ((uint32_t)(q * 32) << 16) + s * 2 * kChunk;
        tmem_st32(ta, hi);
        tmem_st32(ta + kChunk, lo);
        // W box: elementwise, so the swizzled placement does not matter — same offset in the lo plane
#pragma unroll
        for (int h = 0; h < (int)(kWBytes / 16 / 128); ++h) {
          const uint32_t off = (uint32_t)(h * 128 + tt) * 16;
          uint32_t w0, w1, w2, w3;
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(st + kXBytes + off));
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(st + kXBytes + kWBytes + off), "r"(tf32_lo(w0)),
                       "r"(tf32_lo(w1)), "r"(tf32_lo(w2)), "r"(tf32_lo(w3))
                       : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(ready(s));
      }
  } else if (warp >= 4) {
    // ================================================================== epilogue: one output row per thread
    const int q = warp & 3;
    const uint64_t stepkey = drop.thr ? pg::drop_stepkey(drop.seed + (drop.step ? (uint64_t)*drop.step : 0ull)) : 0ull;
    const uint32_t thr_hi = drop.thr << 16;
    uint32_t tl = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
      const int a = tl & 1;
      mbar_wait(tfull(a), (tl >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[32], r2[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * kAccStage, r);
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * kAccStage + kN, r2);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tempty(a));                                        // the MMA warp may overwrite this accumulator stage
#pragma unroll
      for (int j = 0; j < kN; ++j) r[j] = __float_as_uint(__uint_as_float(r2[j]) + __uint_as_float(r[j]));   // small term first
      // The thread owns one output row, but a row-per-lane 16-byte store is 32 sectors per warp instruction, and one bulk
      // copy per row is issued lane by lane (a uniform-datapath instruction: 64 serialised UBLKCP per warp and tile — in
      // the r2l profile the epilogue, not the main loop, bounded the kernel: 8.5 us per tile against 5 us of streaming).
      // The tile is therefore staged as [128 rows x 32 floats] boxes in the SWIZZLE_128B layout (16-byte unit u of row r at
      // unit u ^ (r & 7): conflict-free for row-per-lane stores) and leaves as ONE tensor store per box, issued by one
      // thread; rows beyond n are clipped by the tensor map.
      const int trow = q * 32 + lane;
      const uint32_t e_row = epi_base + (uint32_t)trow * 128u, sw = (uint32_t)(trow & 7);
      const int64_t grow = tile * kBlockM + trow;
      if (threadIdx.x == 4 * 32) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous tile's stores have read the boxes
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const uint64_t rk = dropping ? pg::drop_rowkey(stepkey, (uint64_t)grow) : 0ull;
      auto masked = [&](float4 v, int g) {   // dropout of the 4 columns of group g under the drop_hash contract
        const uint64_t h = pg::drop_mix(rk, colkey_sh[g]);
        const uint32_t hl = (uint32_t)h, hh = (uint32_t)(h >> 32);
        v.x = (hl << 16) >= thr_hi ? v.x * drop.scale : 0.f;
        v.y = hl >= thr_hi ? v.y * drop.scale : 0.f;
        v.z = (hh << 16) >= thr_hi ? v.z * drop.scale : 0.f;
        v.w = hh >= thr_hi ? v.w * drop.scale : 0.f;
        return v;
      };
      auto sts4 = [](uint32_t addr, float4 v) {
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      };
      // boxes: 0 = first 32 columns of out, 1 = its relu half (concat), 2 / 3 = the same of out_drop
#pragma unroll
      for (int j = 0; j < kN; j += 4) {
        const float4 z = make_float4(__uint_as_float(r[j]) + bias_sh[j], __uint_as_float(r[j + 1]) + bias_sh[j + 1],
                                     __uint_as_float(r[j + 2]) + bias_sh[j + 2], __uint_as_float(r[j + 3]) + bias_sh[j + 3]);
        const float4 p = make_float4(fmaxf(z.x, 0.f), fmaxf(z.y, 0.f), fmaxf(z.z, 0.f), fmaxf(z.w, 0.f));
        const uint32_t e = e_row + ((((uint32_t)j >> 2) ^ sw) << 4);
        if (concat) {
          sts4(e, z);
          sts4(e + kEpiBoxBytes, p);
          if (dropping) {
            sts4(e + 2 * kEpiBoxBytes, masked(z, j >> 2));
            sts4(e + 3 * kEpiBoxBytes, masked(p, (kN + j) >> 2));
          }
        } else {
          sts4(e, p);
          if (dropping) sts4(e + 2 * kEpiBoxBytes, masked(p, j >> 2));
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this thread's writes -> visible to the tensor stores
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 4 * 32) {
        const int r0 = (int)(tile * kBlockM);
        tma_store_2d(&tm_out, 0, r0, epi_base);
        if (concat) tma_store_2d(&tm_out, kN, r0, epi_base + kEpiBoxBytes);
        if (dropping) {
          tma_store_2d(&tm_drop, 0, r0, epi_base + 2 * kEpiBoxBytes);
          if (concat) tma_store_2d(&tm_drop, kN, r0, epi_base + 3 * kEpiBoxBytes);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (threadIdx.x == 4 * 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the CTA retires
  }
  // ---- teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(EncodeTiledFn enc, CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint64_t row_stride_floats,
             uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};                        // innermost first
  cuuint64_t strides[1] = {row_stride_floats * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)kChunk, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return (int)enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace

namespace pg {

// PG_ERR_INVALID = not eligible (layout / driver entry point), nothing launched: the caller uses the mma.sync kernel.
pg_status linear_concat_fwd_umma(const float* d_x, int64_t x_stride, const float* d_weight, const float* d_bias, int64_t n,
                                 int32_t K, int concat, float* d_out, int64_t out_stride, float* d_out_drop, int64_t od_stride,
                                 float dropout_p, uint64_t dropout_seed, const int64_t* d_step, int dev, cudaStream_t st) {
  const bool ok = K % 4 == 0 && x_stride % 4 == 0 && out_stride % 4 == 0 && (!d_out_drop || od_stride % 4 == 0) &&
                  (((uintptr_t)d_x | (uintptr_t)d_weight | (uintptr_t)d_out | (uintptr_t)d_out_drop) & 15) == 0 &&
                  n < (int64_t)1 << 31;
  if (!ok) return PG_ERR_INVALID;
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
      cudaGetLastError();
      return PG_ERR_INVALID;
    }
    enc = (EncodeTiledFn)fn;
  }
  // tensor maps of the most recent (pointer, shape) pairs: the training loop calls with the same persistent buffers
  // every step, so the driver encode (a few microseconds of host time each) is paid once
  struct MapEntry { const float* base; uint64_t rows, cols, stride; uint32_t box; CUtensorMap map; };
  static MapEntry cache[32];
  static int next_slot = 0;
  static const bool no_cache = getenv("PG_UMMA_NOCACHE") != nullptr;
  auto get_map = [&](const float* base, uint64_t rows, uint64_t cols, uint64_t stride, uint32_t box) -> const CUtensorMap* {
    if (!no_cache)
      for (MapEntry& e : cache)
      if (e.base == base && e.rows == rows && e.cols == cols && e.stride == stride && e.box == box) return &e.map;
    MapEntry& e = cache[next_slot];
    next_slot = (next_slot + 1) % 32;
    e.base = nullptr;
    if (make_map(enc, &e.map, base, rows, cols, stride, box)) return nullptr;
    e.base = base; e.rows = rows; e.cols = cols; e.stride = stride; e.box = box;
    return &e.map;
  };
  // each map is copied out before the next lookup: an insertion may recycle the slot a previous lookup returned
  const CUtensorMap* pm = get_map(d_x, (uint64_t)n, (uint64_t)K, (uint64_t)x_stride, kBlockM);
  if (!pm) return PG_ERR_INVALID;
  const CUtensorMap tm_x = *pm;
  pm = get_map(d_weight, kN, (uint64_t)K, (uint64_t)K, kN);
  if (!pm) return PG_ERR_INVALID;
  const CUtensorMap tm_w = *pm;
  // store maps: [128 rows x 32 floats] boxes of out / out_drop (rows beyond n are clipped)
  const uint64_t ocols = concat ? 2 * kN : kN;
  pm = get_map(d_out, (uint64_t)n, ocols, (uint64_t)out_stride, kBlockM);
  if (!pm) return PG_ERR_INVALID;
  const CUtensorMap tm_out = *pm;
  CUtensorMap tm_drop = tm_out;
  if (d_out_drop) {
    pm = get_map(d_out_drop, (uint64_t)n, ocols, (uint64_t)od_stride, kBlockM);
    if (!pm) return PG_ERR_INVALID;
    tm_drop = *pm;
  }
  const size_t smem = (size_t)kStages * kStageBytes + kEpiBytes + 1024;
  static bool attr_set[64] = {false};
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    PG_CUDA(cudaFuncSetAttribute(linear_concat_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PG_CUDA(cudaFuncSetAttribute(linear_concat_fwd_umma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int64_t ntiles = (n + kBlockM - 1) / kBlockM;
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)pg::sm_count(dev));
  UmmaDrop drop;
  drop.thr = (d_out_drop && dropout_p > 0.f) ? (uint32_t)(dropout_p * 65536.0f + 0.5f) : 0u;
  drop.scale = drop.thr ? 1.0f / (1.0f - dropout_p) : 1.0f;
  drop.seed = dropout_seed;
  drop.step = d_step;
  linear_concat_fwd_umma_kernel<<<grid, kThreads, smem, st>>>(tm_x, tm_w, tm_out, tm_drop, d_bias, n, K, concat,
                                                             d_out_drop != nullptr, drop);
  PG_CHECK_LAUNCH();
  return PG_OK;
}

}  // namespace pg

// =====================================================================================================================
// Backward of the first NodeUpdate on tcgen05:  dW [32, K] = gz^T x,  db [32] = sum_r gz,  gz [n, 32] = the gradient of z
// recovered from grad_out and out (dropout', relu' and the concat split folded in — same definition as
// linear_concat_dw2_kernel in pg_dense_mma.cu).
//
// The reduction runs over ROWS, so the tensor-core operands are the transposes of what HBM holds:
//     D_fb [128 features x 32]  +=  A_fb [128 features x 8 rows] * B [32 outputs x 8 rows]^T ,   fb = 0 .. ceil(K / 128) - 1.
//   * A = x^T lives in TENSOR MEMORY: lane = feature, column = row. The transform threads (one feature lane each) read
//     the TMA-staged [32 rows x 128 features] boxes column-wise — for a fixed row the 32 lanes of a warp touch 32
//     consecutive words, conflict-free — so the transposition costs nothing; x_hi / x_lo go to TMEM with tcgen05.st.
//   * B = gz^T is built by the same threads straight into the canonical K-major SWIZZLE_128B layout ([32 outputs] rows of
//     128 bytes = 32 rows of the minibatch), hi and lo planes.
//   * 3xTF32 as in the forward (A_lo B_hi + A_hi B_lo + A_hi B_hi), fp32 accumulators in TMEM. An instruction with its A
//     operand in tensor memory is paced by that operand's read, so A_hi meets both B planes in ONE N = 64 instruction
//     (accumulator columns [A_hi B_hi | A_hi B_lo]) and A_lo B_hi follows with N = 32: ceil(K/128) x 64 columns.
// One CTA per SM takes 32-row super-chunks from an atomic counter; at the end every thread of the epilogue warps adds the
// 32 outputs of its feature to dW with scalar reductions (a warp covers 32 consecutive floats of a dW row: one 128-byte
// packet per instruction), db likewise.
// TMEM: 320 accumulator columns + 2 operand stages x (5 blocks x 2 planes x 8 rows) = 480 of 512.
namespace {

constexpr int kDwRows = 32;                       // rows per super-chunk (one 128-byte swizzle row of gz^T)
constexpr int kDwHalf = 8;                        // rows per TMEM operand stage (one k-step)
constexpr int kDwParts = 4;                       // operand stages per super-chunk (kDwRows / kDwHalf), two in flight
constexpr int kDwFB = 5;                          // feature blocks of 128 (K <= 640)
constexpr int kDwXStages = 2;                     // x ring (super-chunks)
constexpr uint32_t kDwBoxBytes = kDwRows * 128 * 4;                    // [32 rows x 128 features] = 16 KB
constexpr uint32_t kDwGBytes = kDwRows * 2 * kN * 4;                   // staged [32 rows x 64] box of grad_out / of y = 8 KB
constexpr uint32_t kDwXStageBytes = kDwFB * kDwBoxBytes + 2 * kDwGBytes;   // x (80 KB) + grad_out + y rows = 96 KB
constexpr uint32_t kDwBBytes = kN * kDwRows * 4;                       // one gz^T plane = 4 KB
constexpr uint32_t kDwSmemBytes = kDwXStages * kDwXStageBytes + 2 * 2 * kDwBBytes;   // 192 KB + 16 KB
constexpr uint32_t kDwAccCols = kDwFB * 2 * kN;                        // 320: per feature block [x gz_hi | x_hi gz_lo]
constexpr uint32_t kDwOpStageCols = kDwFB * 2 * kDwHalf;               // 80
static_assert(kDwAccCols + 2 * kDwOpStageCols <= 512, "TMEM budget of the dW kernel");

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// Work counters of the dW kernel: {next super-chunk, CTAs that have exited}, one pair per launch slot (launches take the
// slots round-robin; the last CTA of a launch zeroes its pair again, so a captured launch can be replayed).
constexpr int kDwCounterSlots = 64;
__device__ unsigned g_dw_counters[kDwCounterSlots][2];

__global__ void __launch_bounds__(kThreads, 1)
    linear_concat_dw_umma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g,
                                 const __grid_constant__ CUtensorMap tm_y, int64_t n, int K, int concat, UmmaDrop drop,
                                 float* __restrict__ dW, float* __restrict__ db, unsigned* __restrict__ counters, int dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // xfull[2], xempty[2] (x ring) | ready[2], opfree[2] (TMEM operand stages) | bfree[2] (gz^T buffers) | done
  __shared__ __align__(8) uint64_t bars[2 * kDwXStages + 4 + 2 + 1];
  __shared__ uint32_t tmem_base_sh;
  __shared__ uint64_t colkey_sh[16];
  __shared__ float db_sh[kN];
  __shared__ int sc_ring[4];            // super-chunk index of iteration it (slot it & 3), -1 = no work left
  __shared__ volatile int stop_sh;      // the transform warps found the end marker (read by the MMA issuer)
  __shared__ volatile int iters_sh;     // super-chunks this CTA has accumulated
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base + kDwXStages * kDwXStageBytes;     // [buf][plane] gz^T tiles, 1024-byte aligned
  auto xfull = [&](int s) { return smem_u32(&bars[s]); };
  auto xempty = [&](int s) { return smem_u32(&bars[kDwXStages + s]); };
  auto ready = [&](int h) { return smem_u32(&bars[2 * kDwXStages + h]); };
  auto opfree = [&](int h) { return smem_u32(&bars[2 * kDwXStages + 2 + h]); };
  auto bfree = [&](int b) { return smem_u32(&bars[2 * kDwXStages + 4 + b]); };
  const uint32_t done = smem_u32(&bars[2 * kDwXStages + 6]);
  const int nfb = (K + 127) / 128;
  const int64_t nsc = (n + kDwRows - 1) / kDwRows;                      // super-chunks
  const int gcols = concat ? 2 * kN : kN;                               // columns of grad_out and of y
  const uint32_t g_bytes = (uint32_t)(kDwRows * gcols * 4);             // one staged [32 rows x gcols] box

  if (threadIdx.x == 0) {
    for (int s = 0; s < kDwXStages; ++s) {
      mbar_init(xfull(s), 1);
      mbar_init(xempty(s), 128);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(ready(h), 128);
      mbar_init(opfree(h), 1);
      mbar_init(bfree(h), 1);
    }
    mbar_init(done, 1);
    stop_sh = 0;
    iters_sh = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 16) colkey_sh[threadIdx.x] = pg::drop_colkey((uint32_t)threadIdx.x);
  if (threadIdx.x < kN) db_sh[threadIdx.x] = 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_sh;
  const uint32_t tmem_opnd = tmem_base + kDwAccCols;

  if (warp == 0) {
    // ================================================================== TMA producer: x, grad_out and y rows of a super-chunk
    // Super-chunks are handed out by an atomic counter, not by a fixed stride: in the training pipeline this kernel
    // starts on the few SMs the input aggregation of the next minibatch leaves free and gets the rest of the GPU when
    // that kernel ends — the CTAs that start early simply take more of the work.
    if (lane == 0) {
      unsigned sc = atomicAdd(&counters[0], 1u);                     // claimed one super-chunk ahead: the round trip of the
      for (uint32_t it = 0;; ++it) {                                 // atomic hides behind the wait for a free stage
        const int s = it % kDwXStages;
        mbar_wait(xempty(s), ((it / kDwXStages) & 1) ^ 1);
        if ((int64_t)sc >= nsc) {
          sc_ring[it & 3] = -1;
          mbar_arrive(xfull(s));                                     // end marker: a phase without bytes
          break;
        }
        sc_ring[it & 3] = (int)sc;
        const uint32_t st = smem_base + s * kDwXStageBytes;
        mbar_expect_tx(xfull(s), (uint32_t)nfb * kDwBoxBytes + 2 * g_bytes);
        for (int fb = 0; fb < nfb; ++fb) tma_load_2d(st + fb * kDwBoxBytes, &tm_x, fb * 128, (int)(sc * kDwRows), xfull(s));
        tma_load_2d(st + kDwFB * kDwBoxBytes, &tm_g, 0, (int)(sc * kDwRows), xfull(s));
        tma_load_2d(st + kDwFB * kDwBoxBytes + kDwGBytes, &tm_y, 0, (int)(sc * kDwRows), xfull(s));
        sc = atomicAdd(&counters[0], 1u);
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      for (uint32_t it = 0;; ++it) {
        const int b = it & 1;
        // B tile of the N = 64 instruction: the gz_hi^T plane and the gz_lo^T plane behind it (64 rows of 128 bytes)
        const uint64_t b_hi = umma_desc_k_sw128(b_base + b * 2 * kDwBBytes);
        bool stop = false;
        for (int part = 0; part < kDwParts; ++part) {
          const int stg = part & 1;
          const uint32_t use = 2 * it + (uint32_t)(part >> 1);       // how often this operand stage has been filled before
          mbar_wait(ready(stg), use & 1);                            // 8 rows of x^T in TMEM, gz^T tile in shared memory
          if (part == 0 && stop_sh) {                                // (or: the end marker)
            stop = true;
            break;
          }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t adv = (uint64_t)((part * kDwHalf * 4) >> 4);   // rows part*8 .. inside the 128-byte row
          const uint32_t acc = (it | (uint32_t)part) != 0;
          for (int fb = 0; fb < nfb; ++fb) {
            // as in the forward: the A operand's read paces the instruction, so x_hi meets [gz_hi ; gz_lo] in ONE N = 64
            // instruction (columns [x_hi gz_hi | x_hi gz_lo] of the block's accumulator) and x_lo gz_hi follows with N = 32
            const uint32_t d = tmem_base + fb * 2 * kN;
            const uint32_t a_hi = tmem_opnd + stg * kDwOpStageCols + fb * 2 * kDwHalf, a_lo = a_hi + kDwHalf;
            umma_tf32_ts(d, a_hi, b_hi + adv, acc, kIdesc2);
            umma_tf32_ts(d, a_lo, b_hi + adv, 1, kIdesc);
          }
          umma_commit(opfree(stg));                                  // TMEM operand stage reusable
        }
        if (stop) break;
        umma_commit(bfree(b));                                       // gz^T buffer b reusable
      }
      umma_commit(done);
    }
  } else if (warp >= 8) {
    // ================================================================== transform: x^T -> TMEM, gz^T -> shared memory
    const int q = warp & 3;
    const int tt = threadIdx.x - 8 * 32;                             // 0 .. 127 = feature lane inside a block
    const uint64_t stepkey = drop.thr ? pg::drop_stepkey(drop.seed + (drop.step ? (uint64_t)*drop.step : 0ull)) : 0ull;
    const uint32_t thr_hi = drop.thr << 16;
    const int gr = tt >> 2, gj0 = (tt & 3) * 8;                      // gz: row gr of the super-chunk, outputs gj0 .. gj0 + 7
    float dbv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const uint32_t g_off = (uint32_t)(gr * gcols + gj0) * 4u;        // this thread's 8 values inside a staged [32 x gcols] box
    auto lds4 = [](uint32_t addr) {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
      return v;
    };
    uint32_t it = 0;
    for (;; ++it) {
      const int s = it % kDwXStages, b = it & 1;
      mbar_wait(xfull(s), (it / kDwXStages) & 1);                    // x, grad_out and y rows of the super-chunk are staged
      const int sc = ((volatile int*)sc_ring)[it & 3];
      if (sc < 0) break;
      const uint32_t st = smem_base + s * kDwXStageBytes;
      // ---- gz of this thread's (row, 8 outputs); rows beyond n were zero-filled by the TMA
      const int64_t r = (int64_t)sc * kDwRows + gr;
      float gzv[8];
      {
        const uint32_t gs = st + kDwFB * kDwBoxBytes + g_off, ys = gs + kDwGBytes;
        float4 ga[2], gb[2], gy[2];
        ga[0] = lds4(gs);
        ga[1] = lds4(gs + 16);
        if (concat) {
          gb[0] = lds4(gs + kN * 4);
          gb[1] = lds4(gs + kN * 4 + 16);
          gy[0] = lds4(ys + kN * 4);
          gy[1] = lds4(ys + kN * 4 + 16);
        } else {
          gb[0] = gb[1] = make_float4(0.f, 0.f, 0.f, 0.f);
          gy[0] = lds4(ys);
          gy[1] = lds4(ys + 16);
        }
        const uint64_t rk = drop.thr ? pg::drop_rowkey(stepkey, (uint64_t)r) : 0ull;
        auto factor4 = [&](int g, float (&f)[4]) {                   // keep-scale of the 4 columns of group g
          const uint64_t h = pg::drop_mix(rk, colkey_sh[g]);
          const uint32_t hl = (uint32_t)h, hh = (uint32_t)(h >> 32);
          f[0] = (hl << 16) >= thr_hi ? drop.scale : 0.f;
          f[1] = hl >= thr_hi ? drop.scale : 0.f;
          f[2] = (hh << 16) >= thr_hi ? drop.scale : 0.f;
          f[3] = hh >= thr_hi ? drop.scale : 0.f;
        };
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const float av[4] = {ga[half].x, ga[half].y, ga[half].z, ga[half].w};
          const float bv[4] = {gb[half].x, gb[half].y, gb[half].z, gb[half].w};
          const float yv[4] = {gy[half].x, gy[half].y, gy[half].z, gy[half].w};
          float fa[4] = {1.f, 1.f, 1.f, 1.f}, fb_[4] = {1.f, 1.f, 1.f, 1.f};
          if (drop.thr) {
            factor4((gj0 >> 2) + half, fa);
            if (concat) factor4(8 + (gj0 >> 2) + half, fb_);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float v;
            if (concat) v = av[e] * fa[e] + (yv[e] > 0.f ? bv[e] * fb_[e] : 0.f);
            else v = yv[e] > 0.f ? av[e] * fa[e] : 0.f;
            gzv[half * 4 + e] = v;
            dbv[half * 4 + e] += v;
          }
        }
      }
      // ---- gz^T tile: output j is a 128-byte row, minibatch row gr the element inside it (SWIZZLE_128B K-major)
      mbar_wait(bfree(b), ((it >> 1) & 1) ^ 1);
      {
        const uint32_t bh = b_base + b * 2 * kDwBBytes, bl = bh + kDwBBytes;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = gj0 + e;
          const uint32_t off = (uint32_t)j * 128u + (uint32_t)(((gr >> 2) ^ (j & 7)) << 4) + (uint32_t)(gr & 3) * 4u;
          const uint32_t bits = __float_as_uint(gzv[e]);
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(bh + off), "r"(bits) : "memory");
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(bl + off), "r"(tf32_lo(bits)) : "memory");
        }
      }
      // ---- x^T: feature lane tt of every block, 8 rows (one k-step) per TMEM operand stage
      const uint32_t sx = st + (uint32_t)tt * 4u;
      for (int part = 0; part < kDwParts; ++part) {
        const int stg = part & 1;
        const uint32_t use = 2 * it + (uint32_t)(part >> 1);
        mbar_wait(opfree(stg), (use & 1) ^ 1);                       // the MMAs of this stage's previous fill are done
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int fb = 0; fb < nfb; ++fb) {
          uint32_t hi[kDwHalf], lo[kDwHalf];
#pragma unroll
          for (int rr = 0; rr < kDwHalf; ++rr)
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(hi[rr]) : "r"(sx + fb * kDwBoxBytes + (uint32_t)(part * kDwHalf + rr) * 512u));
#pragma unroll
          for (int rr = 0; rr < kDwHalf; ++rr) lo[rr] = tf32_lo(hi[rr]);
          const uint32_t ta = tmem_opnd + ((uint32_t)(q * 32) << 16) + stg * kDwOpStageCols + fb * 2 * kDwHalf;
          tmem_st8(ta, hi);
          tmem_st8(ta + kDwHalf, lo);
        }
        if (part == kDwParts - 1) mbar_arrive(xempty(s));            // every load of this stage has been consumed (lo[])
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the gz^T stores -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(ready(stg));
      }
    }
    // end marker: tell the MMA issuer (it waits on ready(0) of this iteration) and the epilogue how much was done. As in a
    // working iteration the arrival is ordered behind the MMAs of the previous super-chunk's half 0 — the issuer has then
    // consumed ready(0)'s previous phase, so the barrier can never run two phases ahead of the thread that polls it.
    if (it > 0) mbar_wait(opfree(0), 1);                             // use 2 * it of stage 0: its previous fill's phase is odd
    if (tt == 0) {
      stop_sh = 1;
      iters_sh = (int)it;
    }
    mbar_arrive(ready(0));
    // db: the 8 lanes of a warp that share (lane & 3) own the same 8 outputs — fold them with shuffles first (a float
    // atomicAdd on shared memory is a CAS loop: 1024 of them onto 32 words cost ~12 us of the r2a kernel)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = dbv[e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 4 && v != 0.f) atomicAdd(&db_sh[gj0 + e], v);
    }
  }
  // ---- epilogue: accumulators -> reductions into dW
  __syncthreads();                                                    // db_sh / iters_sh complete; producers / transform have issued everything
  // Every CTA waits for the issuer's last commit, also one that got no work: the commit's arrival is asynchronous, and a
  // CTA that exits before it lands leaves a stray arrive for whatever the next CTA on this SM keeps at that address.
  if (warp >= 4 && warp < 8) mbar_wait(done, 0);
  // Every CTA adds its kN x K block to dW straight from the accumulators: the thread that owns feature f (its TMEM lane)
  // adds the 32 outputs of f, so a warp's reduction covers 32 consecutive floats of one dW row — one 128-byte packet for
  // the L2 atomic units, the same number of packets as 16-byte vector reductions would make. (r2q/r2r timing experiments:
  // of the 8 us this epilogue used to take, the atomics were 2; staging the block in shared memory for vector reductions,
  // the index arithmetic and the second pass over it were 6 — and summing the blocks of a cluster through distributed
  // shared memory first, 4x fewer atomics, made the kernel slower, not faster.)
  if (warp >= 4 && warp < 8 && iters_sh > 0 && !(dbg & 2)) {          // a CTA that got no work has nothing to add
    const int q = warp & 3;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // blocks are visited from a CTA-dependent start so that the CTAs, which finish together, spread over dW
    for (int i = 0; i < nfb; ++i) {
      const int fb = (i + (int)(blockIdx.x % (unsigned)nfb)) % nfb;
      uint32_t rg[32], rg2[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + fb * 2 * kN, rg);
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + fb * 2 * kN + kN, rg2);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int f = fb * 128 + q * 32 + lane;
      if (f < K && !(dbg & 1)) {
        float* col = dW + f;
#pragma unroll
        for (int j = 0; j < kN; ++j) atomicAdd(col + (size_t)j * K, __uint_as_float(rg2[j]) + __uint_as_float(rg[j]));   // small term first
      }
    }
    const int t = threadIdx.x - 4 * 32;
    if (db && t < kN) atomicAdd(db + t, db_sh[t]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  // The last CTA to leave re-arms the counters for a replay. This thread's own claims on counters[0] have all returned
  // their values, i.e. they have been performed, before it counts itself out: no fence is needed in between.
  if (threadIdx.x == 0) {
    if (atomicAdd(&counters[1], 1u) == gridDim.x - 1) {
      counters[0] = 0;
      counters[1] = 0;
      __threadfence();
    }
  }
}

int make_map_plain(EncodeTiledFn enc, CUtensorMap* m, const float* base, uint64_t rows, uint64_t cols, uint64_t row_stride_floats,
                   uint32_t box_cols, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_floats * sizeof(float)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return (int)enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace

namespace pg {

// dW / db must be zeroed by the caller. PG_ERR_INVALID = not eligible, nothing launched.
pg_status linear_concat_dw_umma(const float* d_x, int64_t x_stride, const float* d_gout, int64_t g_stride, const float* d_y,
                                int64_t y_stride, int64_t n, int32_t K, int concat, float dropout_p, uint64_t dropout_seed,
                                const int64_t* d_step, float* d_gw, float* d_gb, int dev, cudaStream_t st) {
  const bool ok = K % 4 == 0 && K <= 128 * kDwFB && x_stride % 4 == 0 && g_stride % 4 == 0 && y_stride % 4 == 0 &&
                  (((uintptr_t)d_x | (uintptr_t)d_gout | (uintptr_t)d_y | (uintptr_t)d_gw) & 15) == 0 && n < (int64_t)1 << 31;
  if (!ok) return PG_ERR_INVALID;
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
      cudaGetLastError();
      return PG_ERR_INVALID;
    }
    enc = (EncodeTiledFn)fn;
  }
  // tensor maps of the most recent (pointer, shape, box) tuples: the training loop calls with the same persistent buffers
  struct MapEntry { const float* base; uint64_t rows, cols, stride; uint32_t bc, br; CUtensorMap map; };
  static MapEntry cache[16];
  static int next_slot = 0;
  auto get_map = [&](const float* base, uint64_t rows, uint64_t cols, uint64_t stride, uint32_t bc, uint32_t br) -> const CUtensorMap* {
    for (MapEntry& e : cache)
      if (e.base == base && e.rows == rows && e.cols == cols && e.stride == stride && e.bc == bc && e.br == br) return &e.map;
    MapEntry& e = cache[next_slot];
    next_slot = (next_slot + 1) % 16;
    e.base = nullptr;
    if (make_map_plain(enc, &e.map, base, rows, cols, stride, bc, br)) return nullptr;
    e.base = base; e.rows = rows; e.cols = cols; e.stride = stride; e.bc = bc; e.br = br;
    return &e.map;
  };
  // each map is copied out before the next lookup (an insertion may recycle the slot a previous lookup returned)
  const uint32_t gcols = concat ? 2 * kN : kN;
  const CUtensorMap* pm = get_map(d_x, (uint64_t)n, (uint64_t)K, (uint64_t)x_stride, 128, kDwRows);
  if (!pm) return PG_ERR_INVALID;
  const CUtensorMap tm_x = *pm;
  pm = get_map(d_gout, (uint64_t)n, gcols, (uint64_t)g_stride, gcols, kDwRows);
  if (!pm) return PG_ERR_INVALID;
  const CUtensorMap tm_g = *pm;
  pm = get_map(d_y, (uint64_t)n, gcols, (uint64_t)y_stride, gcols, kDwRows);
  if (!pm) return PG_ERR_INVALID;
  const CUtensorMap tm_y = *pm;
  const size_t smem = (size_t)kDwSmemBytes + 1024;
  UmmaDrop drop;
  drop.thr = dropout_p > 0.f ? (uint32_t)(dropout_p * 65536.0f + 0.5f) : 0u;
  drop.scale = drop.thr ? 1.0f / (1.0f - dropout_p) : 1.0f;
  drop.seed = dropout_seed;
  drop.step = d_step;
  const int64_t nsc = (n + kDwRows - 1) / kDwRows;
  const int grid = (int)std::min<int64_t>(nsc, (int64_t)pg::sm_count(dev));
  static unsigned* counters_of[64] = {nullptr};                      // a __device__ symbol has one instance per device
  static int launch_seq = 0;
  unsigned* counters = (dev >= 0 && dev < 64) ? counters_of[dev] : nullptr;
  if (!counters) {                                                   // `dev` is the current device (the caller asked the runtime)
    PG_CUDA(cudaGetSymbolAddress((void**)&counters, g_dw_counters));
    if (dev >= 0 && dev < 64) counters_of[dev] = counters;
  }
  unsigned* ctr = counters + 2 * (launch_seq++ % kDwCounterSlots);   // zero at rest: the previous user's last CTA re-armed it
  static bool attr_set[64] = {false};
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    PG_CUDA(cudaFuncSetAttribute(linear_concat_dw_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PG_CUDA(cudaFuncSetAttribute(linear_concat_dw_umma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  static const int dbg = getenv("PG_DW_DEBUG") ? atoi(getenv("PG_DW_DEBUG")) : 0;   // timing experiments only (wrong results)
  linear_concat_dw_umma_kernel<<<grid, kThreads, smem, st>>>(tm_x, tm_g, tm_y, n, K, concat, drop, d_gw, d_gb, ctr, dbg);
  PG_CHECK_LAUNCH();
  return PG_OK;
}

}  // namespace pg
