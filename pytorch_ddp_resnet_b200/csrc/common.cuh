// Shared device helpers for the sm_100a kernels: PTX wrappers for mbarrier, TMA,
// tcgen05 (MMA / TMEM), vectorised global access, bf16 packing and a counter-based RNG.
// Everything here is hand-written inline PTX; no CUTLASS/CuTe headers are included.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace b200 {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

// One lane of a fully converged warp (the same lane every time): the issuer of TMA / tcgen05 ops.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// round-to-nearest-even fp32 -> bf16 -> fp32 (the rounding points of the autocast reference)
__device__ __forceinline__ float round_bf16(float v) {
  return __bfloat162float(__float2bfloat16_rn(v));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void unpack_bf16x2(uint32_t v, float& lo, float& hi) {
  lo = __uint_as_float(v << 16);
  hi = __uint_as_float(v & 0xffff0000u);
}

// round 8 floats to bf16 precision in place, two at a time through one packing convert: 3 instructions per
// pair against 4 for two round_bf16 (the element-wise BN kernels are issue-bound)
__device__ __forceinline__ void round_bf16_pairs8(float* f) {
#pragma unroll
  for (int j = 0; j < 8; j += 2) unpack_bf16x2(pack_bf16x2(f[j], f[j + 1]), f[j], f[j + 1]);
}

// 8 bf16 <-> 8 floats through one 16-byte vector
struct alignas(16) Vec8 {
  uint4 raw;
  __device__ __forceinline__ void to_float(float* f) const {
    unpack_bf16x2(raw.x, f[0], f[1]);
    unpack_bf16x2(raw.y, f[2], f[3]);
    unpack_bf16x2(raw.z, f[4], f[5]);
    unpack_bf16x2(raw.w, f[6], f[7]);
  }
  __device__ __forceinline__ void from_float(const float* f) {
    raw.x = pack_bf16x2(f[0], f[1]);
    raw.y = pack_bf16x2(f[2], f[3]);
    raw.z = pack_bf16x2(f[4], f[5]);
    raw.w = pack_bf16x2(f[6], f[7]);
  }
};

// streaming 128-bit global loads / stores (no L1 allocation: every element is touched once)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// 256-bit global accesses (sm_100: LDG.256 / STG.256): one full 32-byte sector per lane. The conv epilogues
// own one output ROW per lane (the TMEM layout), so a warp access touches 32 different rows; with 16-byte
// accesses every sector was requested twice (two instructions, each using half of it).
__device__ __forceinline__ void ldg256_stream(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z),
               "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// counter-based RNG for dropout: Philox-2x32 rounds (one 32x32->64 multiply each) keyed by the seed,
// counter = index of a group of four elements. One call yields four 16-bit uniforms, ~3 integer
// operations per element. Four rounds: the keep bits at p = 0.3 show no serial correlation at lags 1 .. 5120
// elements nor across step seeds over 16 M draws (|r| < 3e-4 = noise); three rounds show up to 2.6 %.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint2 rng_draw4(uint64_t seed, uint64_t idx) {
  uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32) ^ (uint32_t)(seed >> 32);
  uint32_t key = (uint32_t)seed;
  for (int r = 0; r < 4; ++r) {
    const uint64_t p = (uint64_t)c0 * 0xD256D193u;
    c0 = (uint32_t)(p >> 32) ^ key ^ c1;
    c1 = (uint32_t)p;
    key += 0x9E3779B9u;
  }
  return make_uint2(c0, c1);
}

// seed actually used by a launch: the host seed plus a device-resident step counter, so that a
// captured CUDA graph draws fresh dropout masks at every replay
__device__ __forceinline__ uint64_t effective_seed(uint64_t seed, const uint64_t* seed_offset) {
  return seed_offset ? seed + (*seed_offset) * 0x9E3779B97F4A7C15ull : seed;
}

// ---------------------------------------------------------------------------------------------
// BN accumulator workspace layout (shared by the BN kernels and the conv epilogues)
// ---------------------------------------------------------------------------------------------
// Accumulator workspace of the reduction kernels: BN_SLOTS copies of fp64 [2][C], `slot_stride` doubles
// apart (>= 4 KB so the copies land in different L2 slices), followed by the ticket counter. Block b adds
// into copy b % BN_SLOTS: ncu showed ~600 blocks firing their atomics at the same few hundred addresses
// in one burst took as long as the streaming loop itself (same-address atomics serialise in one slice).
constexpr int BN_SLOTS = 8;
__host__ __device__ __forceinline__ size_t bn_slot_stride(int C) {
  const size_t need = 2 * (size_t)C;              // doubles
  return (need < 512 ? 512 : (need + 31) / 32 * 32) + 32;
}

// sum of accumulator `a` of channel c over the slots; clears them
__device__ __forceinline__ double drain_slots(double* accum, int C, int a, int c) {
  const size_t stride = bn_slot_stride(C);
  double v[BN_SLOTS];
#pragma unroll
  for (int k = 0; k < BN_SLOTS; ++k) v[k] = __ldcg(accum + k * stride + (size_t)a * C + c);
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < BN_SLOTS; ++k) {
    s += v[k];
    accum[k * stride + (size_t)a * C + c] = 0.0;
  }
  return s;
}

// What the LAST CTA of a conv kernel with fused BN statistics does with the accumulated sums (the finalize
// step of the batch norm that follows the conv, folded into the conv's tail): mean / invstd of `rows`
// values per channel. mean == nullptr: leave the sums for b200_bn_stats_finalize.
// STATS = 2 kernels (a dgrad whose output is the dy of a fused BN + ReLU + dropout backward) accumulate sum(g) and
// sum(g * x) instead, and the last CTA turns them into dbeta / dgamma: mean / invstd are then INPUTS.
struct EpiStatsFinal {
  unsigned int* ticket;   // zero on entry, zero again on exit
  float* mean;
  float* invstd;
  float eps;
  long long rows;
  float* dgamma;          // STATS = 2: outputs (nullptr: leave the sums in the workspace)
  float* dbeta;
};

// STATS = 2: what the epilogue needs of the BN + ReLU + dropout whose backward consumes this dgrad's output. x (the
// BN input, same shape as the output) travels in the residual slot of the kernel arguments.
struct EpiBnBwd {
  const uint8_t* mask;    // one byte per (pixel, 8 channels): the ReLU-and-keep bits bn_act_fwd wrote
  float inv_keep;         // 1 / (1 - p)
  int drop;               // p > 0: g = bf16(dy * inv_keep) where kept
};

__device__ __forceinline__ void sums_to_mean_invstd(double s, double ss, int c, long long rows, float eps,
                                                    float* mean, float* invstd) {
  const double m = s / (double)rows;
  double var = ss / (double)rows - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must abort the kernel (trap -> launch error), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) loads, tile mode, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// multicast variant: the box lands at the same shared-memory offset of every CTA in `mask`, and each
// destination CTA's mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mcast(void* dst, const CUtensorMap* tm, uint64_t* bar,
                                                  int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

// cta_group::2 loads: the data lands in THIS CTA's shared memory, the complete_tx goes to the mbarrier
// at cluster address `bar_cluster_addr` (the pair leader's barrier)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* tm,
                                                uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* tm,
                                                uint32_t bar_cluster_addr, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d_mcast(void* dst, const CUtensorMap* tm, uint64_t* bar,
                                                  int c0, int c1, int c2, int c3, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// thread-block clusters
// ---------------------------------------------------------------------------------------------
// shared::cluster address of `local` (a shared::cta address) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
  return r;
}
// Arrive on a barrier of another CTA of the cluster (the epilogue warps hand a drained TMEM accumulator back to
// the leader's MMA warp). NOT `.release.cluster`: that form compiles to MEMBAR.ALL.CTA + MEMBAR.ALL.GPU + ERRBAR
// in front of the arrive, i.e. every epilogue warp waited after every tile until all of its output stores had been
// acknowledged GPU-wide (ncu on the ImageNet-shape stem: 11 % of all stall samples on that ERRBAR, 5.6 us per tile
// for a 0.7 us main loop). Nothing the MMA warp touches depends on those stores: the accumulator reads are complete
// at tcgen05.wait::ld and ordered by tcgen05.fence::before_thread_sync. The default semantics (release at CTA
// scope, what CUTLASS's ClusterBarrier::arrive(cta_id) emits) is a bare SYNCS.ARRIVE.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM loads, fences
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued MMA of this thread has retired.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(bar))
      : "memory");
}
// Same, arriving on the barrier at this shared-memory offset in every CTA of `mask`.
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// ---- cta_group::2 (an SM pair works on one 256-row tile; the leader CTA issues) ----------------
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::tf32: fp32 operands in shared memory (the tensor core uses the upper 19 bits), fp32 accumulate, K = 8
__device__ __forceinline__ void umma_tf32_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp receives lane (base+i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 64-bit shared-memory matrix descriptor (sm_100 format, version 1).
//   bits [0,14)  start address >> 4      bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4 bits [46,48) version = 1
//   bits [61,64) swizzle: 0 none, 2 = 128B, 4 = 64B, 6 = 32B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7u) << 61;
  return d;
}

// The descriptor split in two 32-bit halves: the high word is constant for a kernel, the low word is
// (start address >> 4) | (LBO >> 4) << 16 and advances by plain integer adds.
__device__ __forceinline__ uint32_t smem_desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
  return ((sbo_bytes >> 4) & 0x3fffu) | (1u << 14) | ((layout_type & 7u) << 29);
}
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3ffffu) >> 4) | (((lbo_bytes >> 4) & 0x3fffu) << 16);
}
__device__ __forceinline__ uint64_t smem_desc_join(uint32_t lo, uint32_t hi) {
  return ((uint64_t)hi << 32) | (uint64_t)lo;
}

// 32-bit instruction descriptor for kind::f16 with bf16 operands and fp32 accumulation.
//   c_format [4,6)=1 (f32)  a_format [7,10)=1 (bf16)  b_format [10,13)=1 (bf16)
//   a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major)  n>>3 [17,23)  m>>4 [24,29)
//   operand format: 0 = f16, 1 = bf16, 2 = tf32 (kind::tf32)
__host__ __device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major,
                                                        uint32_t fmt) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= fmt << 7;
  d |= fmt << 10;
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N, int a_mn_major,
                                                             int b_mn_major) {
  return make_idesc(M, N, a_mn_major, b_mn_major, 1u);
}

}  // namespace b200
