// tcgen05 implicit-GEMM convolution kernels for sm_100a.
//
// Both kernels see a convolution as a sum over filter taps of "shifted-window" GEMMs: for tap t the
// activation operand is the NHWC tensor read through a 4-D TMA box displaced by (dh, dw) pixels
// (out-of-bounds pixels are zero-filled by the TMA unit, which implements the zero padding), so no
// im2col buffer ever exists. Stride-2 convolutions read a parity-split copy of the input
// ([4 phases][N][H/2][W/2][C]) through the same mechanism: every tap maps to one phase image
// and a unit-stride displacement inside it.
//
//   conv_tc_kernel   : D[pixel][k] = sum_t sum_c A_t[pixel][c] * B[k][t,c]       (fprop and dgrad)
//                      A, B K-major in shared memory (channels contiguous), 128-pixel M tile,
//                      N tile = BN output channels, accumulators double-buffered in TMEM,
//                      warp-specialised: TMA producer / MMA issuer / 4 epilogue warps, persistent.
//   wgrad_tc_kernel  : D[(t,c)][k] = sum_pixel X_t[pixel][c] * dY[pixel][k]      (wgrad)
//                      both operands MN-major (the reduction runs over pixels, the slow axis of NHWC),
//                      M tile = 128 rows made of 128/SL channel slabs of any (tap, channel-chunk),
//                      split over pixel ranges with fp32 reduction in global memory.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int TC_MAX_TAPS = 9;
constexpr int TC_MAX_STAGES = 20;
constexpr int TC_THREADS = 192;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..5: epilogue
// SM-pair conv kernels: 8 epilogue warps, two per TMEM lane quadrant, each draining half of the
// accumulator columns (ncu: with 4 warps the epilogue of a 160-channel tile was as long as its MMAs)
constexpr int TC2_THREADS = 64 + 256;
constexpr int TC2_EPI_THREADS = 256;

struct TapTable {
  int n;
  int dh[TC_MAX_TAPS];    // pixel displacement of the window (rows)
  int dw[TC_MAX_TAPS];    // pixel displacement of the window (cols)
  int dn[TC_MAX_TAPS];    // image-index displacement (selects the parity phase image)
  int wcol[TC_MAX_TAPS];  // first column of this tap inside the filter matrix
};

struct ConvTcArgs {
  int bw, bh, bn;                   // TMA box in pixels (w, h, images); bw*bh*bn <= 128
  int rows_valid;                   // bw*bh*bn
  int tiles_w, tiles_h, tiles_n;    // pixel tiles per dimension
  int n_ntiles;                     // output-channel tiles
  int BN;                           // output channels per tile (UMMA N)
  int nkc;                          // ceil(Cin / KC)
  int cin;                          // input channels (the last K-block of a tap may be partial)
  int P, Q, Nimg;                   // output extent
  int ldo;                          // output channel pitch (elements)
  int num_tiles;
  int stages;
  uint32_t stage_bytes;
  uint32_t a_bytes;                 // 128 * KC * 2
  uint32_t tx_bytes;                // bytes landed per K-block (per CTA)
  int gblk;                         // K-blocks per pipeline stage (SM-pair kernel)
  int cstride;                      // conv stride: window origin = output pixel * cstride + tap displacement
                                    // (stride 2 reads every other pixel through the TMA map's elementStrides)
  // Output phases (SM-pair kernel; stride-2 dgrad in ONE launch): with nphase = 4 a unit is (phase, pixel tile,
  // channel tile); phase ph = 2a + b covers the output pixels (2h' + a, 2w' + b), tiles are boxes in (w', h', n),
  // and only the taps [phase_tap0[ph], phase_tap0[ph+1]) of the table contribute to it. Units are ordered
  // phase-major, heaviest phase first, so the round-robin schedule hands every SM pair a similar mix.
  // nphase = 1: one phase holding all taps (phase_tap0 = {0, taps.n}).
  int nphase;
  int phase_tap0[5];
  int phase_id[4];                  // ph -> 2a + b (the phases are sorted by tap count)
  uint32_t block_bytes;             // shared-memory bytes of one K-block (A tile + B tile)
  TapTable taps;
  bf16* out;
  const bf16* residual;
  const float* bias;
  double* stats;                    // BN accumulator workspace (STATS kernels) or nullptr
  EpiStatsFinal fin;                // in-kernel finalize of the statistics (fin.mean == nullptr: off)
  int relu;                         // epilogue ReLU after bias / residual (eval with batch norm folded into the conv)
  EpiBnBwd bb;                      // STATS = 2 (residual then holds the BN input x and is NOT added)
  int det;                          // deterministic mode: fixed-order reduction of the epilogue statistics
};

// -------------------------------------------------------------------------------------------------
// Fused batch-norm statistics of a conv output (epilogue side, STATS = true instantiations).
// The BN that follows a conv needs sum(y) and sum(y^2) per output channel over all pixels; computing
// them here saves the separate statistics kernel (one more pass over y and ~12 us of launch / drain
// per BN layer). Per 16-column chunk each epilogue warp transposes-and-reduces its 32 rows with 32
// shuffles (lane L ends with column L >> 1: sum in even lanes, sum of squares in odd lanes) and adds
// them to the CTA's shared partials; per tile the 4 epilogue warps flush those with one fp64 atomic per
// (channel, statistic) into slot (tile % BN_SLOTS) of the accumulator workspace (common.cuh).
// -------------------------------------------------------------------------------------------------
// Residual operand of the bf16 epilogues: a thread's slice of its output row (<= 80 columns = 10 16-byte
// vectors; the host caps the N tile at 160 channels when a residual is fused) is fetched ONE TILE AHEAD into
// registers: the loads for tile i+1 are issued when tile i's accumulator is ready and complete while tile i
// drains. Loading inside the chunk loop, or all at once right before the drain, left the epilogue waiting a
// full DRAM round trip per tile (86 us against 55 us for the 160-channel 32x32 layer, gpurun_out/call14).
constexpr int EPI_RES_VECS = 10;
constexpr int EPI_MAX_CHUNKS = 8;         // 16-column chunks per epilogue thread (N tile <= 256, two halves)
constexpr int EPI_STATS_MAX_BN = 256;
constexpr int EPI_STATS_MAX_C = 1024;   // output channels a CTA can keep partial sums for (8 KB of shared memory)

// sv / qv: the two per-element terms of this thread's 16 columns (STATS = 1: y and y^2; STATS = 2: g and g * x)
// Returns this lane's share of the warp's 32-row reduction: column (lane >> 1) of the chunk, statistic (lane & 1).
// The caller keeps one such register per chunk across ALL tiles of the CTA (epi_stats_regs_flush): round 2 first
// added it to shared memory right here, a float atomicAdd = a compare-and-swap loop (ATOMS.CAST.SPIN) on which the
// four warps that drain the same columns of a tile collide - the epilogue of a 160-channel tile then outlasted its
// MMAs (6.2 us per tile at 128x32x32x160).
__device__ __forceinline__ float epi_stats_chunk2(const float* sv, const float* qv, int lane) {
  float s[16], q[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    s[j] = sv[j];
    q[j] = qv[j];
  }
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    const int off = 16 >> step;  // partner distance 16, 8, 4, 2
    const int w = 8 >> step;     // values kept 8, 4, 2, 1
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float rs = __shfl_xor_sync(0xffffffffu, up ? s[i] : s[i + w], off);
      const float rq = __shfl_xor_sync(0xffffffffu, up ? q[i] : q[i + w], off);
      s[i] = (up ? s[i + w] : s[i]) + rs;
      q[i] = (up ? q[i + w] : q[i]) + rq;
    }
  }
  const float s0 = s[0] + __shfl_xor_sync(0xffffffffu, s[0], 1);
  const float q0 = q[0] + __shfl_xor_sync(0xffffffffu, q[0], 1);
  return (lane & 1) ? q0 : s0;
}

__device__ __forceinline__ float epi_stats_chunk(const float* fr, int lane) {
  float q[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) q[j] = fr[j] * fr[j];
  return epi_stats_chunk2(fr, q, lane);
}

// the per-chunk registers of a warp -> the CTA's shared partial sums of channel tile `sp` (when the channel tile
// changes between two units of the CTA, and after its last unit)
// det_wq < 0: shared-memory float atomics (the four warps of the four TMEM lane quadrants add to the same slots in
// whatever order they arrive). det_wq = the warp's quadrant (deterministic mode): the quadrants take turns,
// separated by barriers among the epilogue warps, so every slot is summed in the fixed order 0, 1, 2, 3 (the two
// warps of one quadrant own disjoint columns). Every epilogue warp of the CTA makes this call for the same units.
template <int NCH>
__device__ __forceinline__ void epi_stats_regs_flush(float* racc, float* sp, int c_lo, int c_hi, int lane,
                                                     int det_wq) {
  if (det_wq >= 0) {
    for (int turn = 0; turn < 4; ++turn) {
      if (det_wq == turn) {
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c = c_lo + 16 * ci;
          if (c < c_hi) sp[(c + (lane >> 1)) * 2 + (lane & 1)] += racc[ci];
          racc[ci] = 0.f;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    return;
  }
#pragma unroll
  for (int ci = 0; ci < NCH; ++ci) {
    const int c = c_lo + 16 * ci;
    if (c < c_hi) atomicAdd(&sp[(c + (lane >> 1)) * 2 + (lane & 1)], racc[ci]);
    racc[ci] = 0.f;
  }
}

// STATS = 2, one 16-column chunk: f = the bf16-rounded dgrad output of this thread's row (= dy of the BN backward),
// x = the BN input, bits = the ReLU-and-keep mask of these 16 units. g = dy (* 1/(1-p), rounded to bf16) where the
// bit is set, else 0 - the arithmetic of masked_grad_from<true> (elementwise.cuh) - and the sums are over g, g * x.
__device__ __forceinline__ float epi_bnbwd_chunk(const float* f, const Vec8& x0, const Vec8& x1, uint32_t bits,
                                                 bool valid, const EpiBnBwd& bb, int lane) {
  float g[16], q[16], xf[16];
  x0.to_float(xf);
  x1.to_float(xf + 8);
#pragma unroll
  for (int j = 0; j < 16; ++j) g[j] = f[j];
  if (bb.drop) {
#pragma unroll
    for (int j = 0; j < 16; ++j) g[j] *= bb.inv_keep;
    round_bf16_pairs8(g);
    round_bf16_pairs8(g + 8);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    g[j] = (valid && ((bits >> j) & 1u)) ? g[j] : 0.f;
    q[j] = valid ? g[j] * xf[j] : 0.f;
  }
  return epi_stats_chunk2(g, q, lane);
}

// barrier among the 256 epilogue threads of a CTA (warps 2..9), id 1 (id 0 is __syncthreads)
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// after the LAST tile of a CTA: flush its partial sums of all C output channels (the partials of the ~7 tiles a
// CTA processes stay in shared memory in fp32; round 1 flushed after every tile: two more barriers and 2*BN
// global atomics inside the tile loop)
__device__ __forceinline__ void epi_stats_flush(float* s_part, double* accum, int C, int e) {
  epi_bar();
  double* dst = accum + (size_t)(blockIdx.x % BN_SLOTS) * bn_slot_stride(C);
  for (int i = e; i < 2 * C; i += TC2_EPI_THREADS)
    atomicAdd(dst + (size_t)(i & 1) * C + (i >> 1), (double)s_part[i]);
  epi_bar();
}

// after the last tile of a CTA: the CTA that finishes last turns the accumulated sums into mean / invstd
// (the finalize step of the following batch norm; saves one launch per BN layer) and clears them
template <bool BWD>
__device__ __forceinline__ void epi_stats_finalize(const EpiStatsFinal& fin, double* accum, int C, int e) {
  __shared__ int s_last;
  if (fin.mean == nullptr || (BWD && fin.dgamma == nullptr)) return;
  // every epilogue thread of this CTA has issued its atomics (epi_stats_flush ends with a barrier)
  if (e == 0) {
    __threadfence();
    s_last = (atomicAdd(fin.ticket, 1u) == gridDim.x - 1u) ? 1 : 0;
    __threadfence();
  }
  epi_bar();
  if (s_last) {
    for (int c = e; c < C; c += TC2_EPI_THREADS) {
      const double s0 = drain_slots(accum, C, 0, c), s1 = drain_slots(accum, C, 1, c);
      if (BWD) {   // dbeta = sum g, dgamma = sum g * xhat = invstd * (sum g*x - mean * sum g)
        fin.dbeta[c] = (float)s0;
        fin.dgamma[c] = (float)((double)fin.invstd[c] * (s1 - (double)fin.mean[c] * s0));
      } else {
        sums_to_mean_invstd(s0, s1, c, fin.rows, fin.eps, fin.mean, fin.invstd);
      }
    }
    if (e == 0) *fin.ticket = 0u;
  }
}

// K-major operand tile: rows of KC elements of ES bytes (128 / 64 / 32-byte rows = the three swizzle modes);
// one MMA consumes 32 bytes of every row (16 bf16 or 8 tf32).
template <int KC, int ES>
struct KMajorCfgE {
  static constexpr uint32_t ROW_BYTES = KC * ES;
  static constexpr uint32_t SBO = 8 * ROW_BYTES;
  static constexpr uint32_t LAYOUT = (ROW_BYTES == 128) ? 2u : (ROW_BYTES == 64) ? 4u : 6u;
  static constexpr int KSTEPS = ROW_BYTES / 32;
  static_assert(ROW_BYTES == 128 || ROW_BYTES == 64 || ROW_BYTES == 32, "row = one swizzle span");
};
template <int KC>
using KMajorCfg = KMajorCfgE<KC, 2>;

// CS = cluster size. With CS > 1 the CS CTAs of a cluster work on CS consecutive pixel tiles of the same
// output-channel tile in lock step; each loads 1/CS of the filter tile and multicasts it to all of them
// (filter traffic from L2 drops by CS), and every MMA issuer releases a pipeline slot in ALL CTAs.
template <int KC, int CS>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ ConvTcArgs args) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[TC_MAX_STAGES];
  __shared__ uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ uint64_t tfull_bar[2];
  __shared__ uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < args.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CS);
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], 4);
    mbar_init(&tempty_bar[1], 4);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();  // peers' barriers are initialised before any remote arrive / copy
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int tiles_hw = args.tiles_w * args.tiles_h;
  // cluster-tile schedule: cluster c takes cluster tiles c, c + #clusters, ...; cluster tile ct covers
  // pixel tiles (ct / n_ntiles) * CS + rank of output-channel tile ct % n_ntiles
  const int crank = (CS > 1) ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CS;
  const int num_clusters = gridDim.x / CS;
  const int num_ctiles = args.num_tiles / CS;
  const uint16_t cmask = (uint16_t)((1u << CS) - 1u);

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) ==============
    int s = 0;
    uint32_t ph = 0;
    const int bslice = args.BN / CS;                           // filter rows this CTA loads
    const uint32_t bslice_bytes = (uint32_t)bslice * KC * 2u;
    for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
      const int nt = ct % args.n_ntiles;
      const int mt = (ct / args.n_ntiles) * CS + crank;
      const int w0 = (mt % args.tiles_w) * args.bw * args.cstride;
      const int h0 = ((mt / args.tiles_w) % args.tiles_h) * args.bh * args.cstride;
      const int n0 = (mt / tiles_hw) * args.bn;
      for (int t = 0; t < args.taps.n; ++t) {
        const int cw = w0 + args.taps.dw[t];
        const int ch = h0 + args.taps.dh[t];
        const int cn = n0 + args.taps.dn[t];
        const int wc = args.taps.wcol[t];
        for (int kc = 0; kc < args.nkc; ++kc) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (elect_one()) {
            uint8_t* a_dst = smem + (size_t)s * args.stage_bytes;
            uint8_t* b_dst = a_dst + args.a_bytes;
            mbar_expect_tx(&full_bar[s], args.tx_bytes);
            tma_load_4d(a_dst, &tmA, &full_bar[s], kc * KC, cw, ch, cn);
            if (CS == 1) {
              tma_load_2d(b_dst, &tmB, &full_bar[s], wc + kc * KC, nt * args.BN);
            } else {
              tma_load_2d_mcast(b_dst + (size_t)crank * bslice_bytes, &tmB, &full_bar[s], wc + kc * KC,
                                nt * args.BN + crank * bslice, cmask);
            }
          }
          __syncwarp();
          if (++s == args.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp loops, one elected lane issues) ================
    const uint32_t idesc = make_idesc_bf16(128, args.BN, 0, 0);
    const uint32_t dhi = smem_desc_hi(KMajorCfg<KC>::SBO, KMajorCfg<KC>::LAYOUT);
    const uint32_t smem_base = smem_u32(smem);
    int s = 0;
    uint32_t ph = 0;
    int as = 0;
    uint32_t aph = 0;
    const int nk = args.taps.n * args.nkc;
    for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
      mbar_wait(&tempty_bar[as], aph ^ 1);
      tc_fence_after();
      const uint32_t d_addr = tmem_base + (uint32_t)as * 256u;
      for (int it = 0; it < nk; ++it) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_base + (uint32_t)s * args.stage_bytes;
          const uint32_t alo = smem_desc_lo(a_addr, 16);
          const uint32_t blo = smem_desc_lo(a_addr + args.a_bytes, 16);
          const uint32_t ahi = dhi;
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
            umma_bf16_ss(d_addr, smem_desc_join(alo + 2 * k, ahi), smem_desc_join(blo + 2 * k, dhi),
                         idesc, (it | k) != 0 ? 1u : 0u);
          }
          if (CS == 1) umma_commit(&empty_bar[s]);
          else umma_commit_mcast(&empty_bar[s], cmask);
          if (it == nk - 1) umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++s == args.stages) { s = 0; ph ^= 1; }
      }
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
  } else {
    // ===================== epilogue (4 warps, one TMEM lane quarter each) =====================
    const int wq = warp & 3;
    const int m = wq * 32 + lane;
    const int wi = m % args.bw;
    const int hi = (m / args.bw) % args.bh;
    const int ni = m / (args.bw * args.bh);
    int as = 0;
    uint32_t aph = 0;
    for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
      const int nt = ct % args.n_ntiles;
      const int mt = (ct / args.n_ntiles) * CS + crank;
      const int w = (mt % args.tiles_w) * args.bw + wi;
      const int h = ((mt / args.tiles_w) % args.tiles_h) * args.bh + hi;
      const int n = (mt / tiles_hw) * args.bn + ni;
      const bool valid = (m < args.rows_valid) && (w < args.Q) && (h < args.P) && (n < args.Nimg);
      const size_t pix = ((size_t)n * args.P + h) * args.Q + w;
      const size_t off = pix * (size_t)args.ldo + (size_t)nt * args.BN;
      bf16* orow = args.out + off;
      const bf16* rrow = args.residual ? args.residual + off : nullptr;
      const float* brow = args.bias ? args.bias + (size_t)nt * args.BN : nullptr;

      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)as * 256u;
      for (int c = 0; c < args.BN; c += 16) {
        uint32_t v[16];
        tmem_ld16(t_addr + c, v);
        tmem_ld_wait();
        if (valid) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
          if (brow) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] += round_bf16(__ldg(brow + c + j));
          }
          if (rrow) {
            Vec8 r0, r1;
            r0.raw = *reinterpret_cast<const uint4*>(rrow + c);
            r1.raw = *reinterpret_cast<const uint4*>(rrow + c + 8);
            float rf[16];
            r0.to_float(rf);
            r1.to_float(rf + 8);
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = round_bf16(f[j]) + rf[j];
          }
          if (args.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          Vec8 o0, o1;
          o0.from_float(f);
          o1.from_float(f + 8);
          *reinterpret_cast<uint4*>(orow + c) = o0.raw;
          *reinterpret_cast<uint4*>(orow + c + 8) = o1.raw;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();  // no CTA leaves while a peer may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// -------------------------------------------------------------------------------------------------
// conv_tc2_kernel: the same shifted-window GEMM on an SM PAIR (tcgen05 cta_group::2).
// The pair computes a 256-pixel x BN tile: each CTA stages its own 128-pixel A tile and HALF of the
// filter tile (BN/2 rows), the leader CTA issues M=256 MMAs that read both CTAs' shared memory, and each
// CTA ends up with its own 128 x BN accumulator in its own TMEM. Per 128x160x16 of work an SM now moves
// 4 KB (A) + 2.5 KB (B/2) in and out of shared memory instead of 4 + 5 KB: shared-memory bandwidth,
// not L2, is what limited the single-CTA kernel (ncu: tensor pipe active 53 %, 150 cycles per MMA
// against 144 cycles of shared-memory traffic).
//   barriers: full[s] lives in the leader (both CTAs' TMA loads complete_tx on it); empty[s] / tfull[a]
//   are local in each CTA and signalled by the leader's multicast tcgen05.commit; tempty[a] lives in the
//   leader and collects the 8 epilogue warps of the pair.
// -------------------------------------------------------------------------------------------------
// TF32 = true: the fp32 / TF32 precision mode (the reference's evaluation path and its non-AMP training path,
// evaluation.py:32-39, training.py:101-102): fp32 activations and filters in shared memory, kind::tf32 MMAs
// (K = 8), fp32 output, fp32 bias / residual added without intermediate rounding. KC counts fp32 elements.
template <int KC, int STATS, bool TF32 = false>
__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ ConvTcArgs args) {
  using Cfg = KMajorCfgE<KC, TF32 ? 4 : 2>;
  static_assert(!(TF32 && STATS), "fused statistics: bf16 training path only");
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[TC_MAX_STAGES];
  __shared__ uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ uint64_t tfull_bar[2];
  __shared__ uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_part[STATS ? 2 * EPI_STATS_MAX_C : 1];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  const int crank = (int)cluster_ctarank();
  const bool leader = crank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < args.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], 16);
    mbar_init(&tempty_bar[1], 16);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(&tmem_base_smem, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int tiles_hw = args.tiles_w * args.tiles_h;
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_ptiles = args.num_tiles >> 1;  // pair tiles
  const int bhalf = args.BN >> 1;

  const int units_per_phase = num_ptiles / args.nphase;
  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    // a stage holds up to gblk K-blocks (any mix of taps / channel chunks), one barrier round trip
    int s = 0;
    uint32_t ph = 0;
    for (int ct = pair_id; ct < num_ptiles; ct += num_pairs) {
      const int phs = ct / units_per_phase;
      const int cu = ct - phs * units_per_phase;
      const int tap0 = args.phase_tap0[phs];
      const int nk = (args.phase_tap0[phs + 1] - tap0) * args.nkc;   // K-blocks of this unit
      const int nst = (nk + args.gblk - 1) / args.gblk;              // pipeline stages of this unit
      const int nt = cu % args.n_ntiles;
      const int mt = (cu / args.n_ntiles) * 2 + crank;
      const int w0 = (mt % args.tiles_w) * args.bw * args.cstride;
      const int h0 = ((mt / args.tiles_w) % args.tiles_h) * args.bh * args.cstride;
      const int n0 = (mt / tiles_hw) * args.bn;
      for (int st = 0; st < nst; ++st) {
        const int blk0 = st * args.gblk;
        const int cnt = min(args.gblk, nk - blk0);
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* base = smem + (size_t)s * args.stage_bytes;
          const uint32_t full_leader = map_to_cta(smem_u32(&full_bar[s]), 0);
          if (leader) mbar_expect_tx(&full_bar[s], 2u * (uint32_t)cnt * args.tx_bytes);
          for (int g = 0; g < cnt; ++g) {
            const int blk = blk0 + g;
            const int tl = blk / args.nkc;
            const int kc = blk - tl * args.nkc;
            const int t = tap0 + tl;
            uint8_t* a_dst = base + (size_t)g * args.block_bytes;
            tma_load_4d_2sm(a_dst, &tmA, full_leader, kc * KC, w0 + args.taps.dw[t],
                            h0 + args.taps.dh[t], n0 + args.taps.dn[t]);
            tma_load_2d_2sm(a_dst + args.a_bytes, &tmB, full_leader, args.taps.wcol[t] + kc * KC,
                            nt * args.BN + crank * bhalf);
          }
        }
        __syncwarp();
        if (++s == args.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer (leader CTA only) =====================
      const uint32_t idesc = make_idesc(256, args.BN, 0, 0, TF32 ? 2u : 1u);
      const uint32_t dhi = smem_desc_hi(Cfg::SBO, Cfg::LAYOUT);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t blk_step = args.block_bytes >> 4;
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int ct = pair_id; ct < num_ptiles; ct += num_pairs) {
        const int phs = ct / units_per_phase;
        const int nk = (args.phase_tap0[phs + 1] - args.phase_tap0[phs]) * args.nkc;
        const int nst = (nk + args.gblk - 1) / args.gblk;
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)as * 256u;
        for (int st = 0; st < nst; ++st) {
          const int cnt = min(args.gblk, nk - st * args.gblk);
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_base + (uint32_t)s * args.stage_bytes;
            uint32_t alo = smem_desc_lo(a_addr, 16);
            uint32_t blo = smem_desc_lo(a_addr + args.a_bytes, 16);
            uint32_t acc = st != 0 ? 1u : 0u;
            for (int g = 0; g < cnt; ++g) {
#pragma unroll
              for (int k = 0; k < Cfg::KSTEPS; ++k) {
                if (TF32)
                  umma_tf32_ss_2sm(d_addr, smem_desc_join(alo + 2 * k, dhi),
                                   smem_desc_join(blo + 2 * k, dhi), idesc, acc);
                else
                  umma_bf16_ss_2sm(d_addr, smem_desc_join(alo + 2 * k, dhi),
                                   smem_desc_join(blo + 2 * k, dhi), idesc, acc);
                acc = 1u;
              }
              alo += blk_step;
              blo += blk_step;
            }
            umma_commit_2sm_mcast(&empty_bar[s], 3);
            if (st == nst - 1) umma_commit_2sm_mcast(&tfull_bar[as], 3);
          }
          __syncwarp();
          if (++s == args.stages) { s = 0; ph ^= 1; }
        }
        as ^= 1;
        if (as == 0) aph ^= 1;
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own TMEM, own pixel tile) =====================
    const int wq = warp & 3;                       // TMEM lane quadrant of this warp
    const int half = (warp - 2) >> 2;              // which half of the columns it drains
    const int e = (warp - 2) * 32 + lane;          // index among the epilogue threads
    const int c_lo = half * (args.BN >> 1), c_hi = c_lo + (args.BN >> 1);
    const int m = wq * 32 + lane;
    const int wi = m % args.bw;
    const int hi = (m / args.bw) % args.bh;
    const int ni = m / (args.bw * args.bh);
    int as = 0;
    uint32_t aph = 0;
    if (STATS) {
      for (int i = e; i < 2 * EPI_STATS_MAX_C; i += TC2_EPI_THREADS) s_part[i] = 0.f;
      epi_bar();
    }
    float racc[STATS ? EPI_MAX_CHUNKS : 1];   // this lane's running sums, one per 16-column chunk of its warp
#pragma unroll
    for (int i = 0; i < (STATS ? EPI_MAX_CHUNKS : 1); ++i) racc[i] = 0.f;
    int racc_nt = -1;                         // channel tile they belong to
    // output row of this thread in unit `ct`: element offset of its first column, validity, channel tile
    auto unit_row = [&](int ct, bool& valid, int& nt) -> size_t {
      const int phs = ct / units_per_phase;
      const int cu = ct - phs * units_per_phase;
      nt = cu % args.n_ntiles;
      const int mt = (cu / args.n_ntiles) * 2 + crank;
      int w = (mt % args.tiles_w) * args.bw + wi;
      int h = ((mt / args.tiles_w) % args.tiles_h) * args.bh + hi;
      const int n = (mt / tiles_hw) * args.bn + ni;
      if (args.nphase > 1) {   // output pixel (2h' + a, 2w' + b) of phase 2a + b; P, Q are the full extents
        const int pid = args.phase_id[phs];
        h = 2 * h + (pid >> 1);
        w = 2 * w + (pid & 1);
      }
      valid = (m < args.rows_valid) && (w < args.Q) && (h < args.P) && (n < args.Nimg);
      const size_t pix = ((size_t)n * args.P + h) * args.Q + w;
      return pix * (size_t)args.ldo + (size_t)nt * args.BN;
    };
    auto load_residual = [&](Vec8* dst, size_t off) {
#pragma unroll
      for (int i = 0; i < EPI_RES_VECS / 2; ++i)
        if (c_lo + 16 * i < c_hi) ldg256_stream(args.residual + off + c_lo + 16 * i, dst[2 * i].raw, dst[2 * i + 1].raw);
    };
    Vec8 rcur[EPI_RES_VECS], rnext[EPI_RES_VECS];
    // STATS = 2: the mask bytes of this thread's row slice, 16 columns per 16-bit load, prefetched like the residual
    uint32_t mcur[STATS == 2 ? EPI_RES_VECS / 2 : 1], mnext[STATS == 2 ? EPI_RES_VECS / 2 : 1];
    auto load_mask = [&](uint32_t* dst, size_t off) {
      const uint16_t* mp = reinterpret_cast<const uint16_t*>(args.bb.mask) + ((off + c_lo) >> 4);
#pragma unroll
      for (int i = 0; i < EPI_RES_VECS / 2; ++i)
        if (c_lo + 16 * i < c_hi) dst[STATS == 2 ? i : 0] = __ldg(mp + i);
    };
    const bool prefetch_res = !TF32 && args.residual != nullptr;
    if (prefetch_res && pair_id < num_ptiles) {
      bool v0;
      int nt0;
      const size_t off0 = unit_row(pair_id, v0, nt0);
      if (v0) load_residual(rcur, off0);
      if (STATS == 2 && v0) load_mask(mcur, off0);
    }
    for (int ct = pair_id; ct < num_ptiles; ct += num_pairs) {
      bool valid;
      int nt;
      const size_t off = unit_row(ct, valid, nt);
      if (STATS && nt != racc_nt) {
        if (racc_nt >= 0) epi_stats_regs_flush<EPI_MAX_CHUNKS>(racc, s_part + 2 * racc_nt * args.BN, c_lo, c_hi, lane,
                                                              args.det ? wq : -1);
        racc_nt = nt;
      }
      bf16* orow = args.out + off;
      const bf16* rrow = args.residual ? args.residual + off : nullptr;
      const float* brow = args.bias ? args.bias + (size_t)nt * args.BN : nullptr;

      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      if (prefetch_res && ct + num_pairs < num_ptiles) {   // next unit's residual row: in flight while this one drains
        bool vn;
        int ntn;
        const size_t offn = unit_row(ct + num_pairs, vn, ntn);
        if (vn) load_residual(rnext, offn);
        if (STATS == 2 && vn) load_mask(mnext, offn);
      }
      const uint32_t t_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)as * 256u;
#pragma unroll
      for (int ci = 0; ci < EPI_MAX_CHUNKS; ++ci) {
        const int c = c_lo + 16 * ci;
        if (c >= c_hi) break;
        const int ri = (2 * ci + 1 < EPI_RES_VECS) ? 2 * ci : 0;   // (a fused residual implies <= 5 chunks)
        uint32_t v[16];
        tmem_ld16(t_addr + c, v);
        tmem_ld_wait();
        if constexpr (TF32) {   // fp32 in, fp32 out: no rounding points
          if (valid) {
            float* of = reinterpret_cast<float*>(args.out) + off + c;
            const float* rf = args.residual ? reinterpret_cast<const float*>(args.residual) + off + c : nullptr;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                     __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
              if (brow) {
                o.x += __ldg(brow + c + 4 * q); o.y += __ldg(brow + c + 4 * q + 1);
                o.z += __ldg(brow + c + 4 * q + 2); o.w += __ldg(brow + c + 4 * q + 3);
              }
              if (rf) {
                const float4 r = *reinterpret_cast<const float4*>(rf + 4 * q);
                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
              }
              if (args.relu) {
                o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
              }
              *reinterpret_cast<float4*>(of + 4 * q) = o;
            }
          }
        } else {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = 0.f;
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
            if (brow) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] += round_bf16(__ldg(brow + c + j));
            }
            if (rrow && STATS != 2) {
              float rf[16];
              rcur[ri].to_float(rf);
              rcur[ri + 1].to_float(rf + 8);
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = round_bf16(f[j]) + rf[j];
            }
            if (args.relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (STATS) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = round_bf16(f[j]);  // statistics of the values stored
            }
            Vec8 o0, o1;
            o0.from_float(f);
            o1.from_float(f + 8);
            stg256(orow + c, o0.raw, o1.raw);
          }
          if (STATS == 1) racc[STATS ? ci : 0] += epi_stats_chunk(f, lane);
          if (STATS == 2)
            racc[STATS ? ci : 0] += epi_bnbwd_chunk(f, rcur[ri], rcur[ri + 1],
                                                    mcur[(STATS == 2 && ci < EPI_RES_VECS / 2) ? ci : 0], valid,
                                                    args.bb, lane);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tempty_bar[as]), 0));
      if (prefetch_res) {
#pragma unroll
        for (int i = 0; i < EPI_RES_VECS; ++i) rcur[i] = rnext[i];
        if (STATS == 2) {
#pragma unroll
          for (int i = 0; i < EPI_RES_VECS / 2; ++i) mcur[STATS == 2 ? i : 0] = mnext[STATS == 2 ? i : 0];
        }
      }
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
    if (STATS) {
      if (racc_nt >= 0) epi_stats_regs_flush<EPI_MAX_CHUNKS>(racc, s_part + 2 * racc_nt * args.BN, c_lo, c_hi, lane,
                                                              args.det ? wq : -1);
      epi_stats_flush(s_part, args.stats, args.ldo, e);
      epi_stats_finalize<STATS == 2>(args.fin, args.stats, args.ldo, e);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// -------------------------------------------------------------------------------------------------
// conv_tc2h_kernel: SM-pair kernel with HALO REUSE for 3x3 / stride 1 / pad 1 convolutions.
// The output tile is 8 pixels wide x 16 rows, so every 8-row MMA group is one image-row segment and the
// groups of a shifted window sit at a uniform stride in shared memory. Per K-block ONE activation patch
// (18 rows x 16 pixels, halo included, zero-filled by TMA outside the image) is staged and all 9 taps read
// it through descriptor start offsets of (dh+1)*16 + (dw+1) rows (measured on B200: the UMMA swizzle
// follows absolute shared-memory address bits, so a start address that is not 8-row aligned is legal with
// base_offset = 0). Activation traffic through TMA drops from 9 x 16 KB to 36 KB per K-block; the per-tap
// filter tiles stream through their own ring (several taps per stage).
// -------------------------------------------------------------------------------------------------
constexpr int HALO_PH = 18;   // patch rows (16 + 2 halo)
constexpr int HALO_BSTAGES_MAX = 8;

struct ConvHaloArgs {
  int tiles_w, tiles_h;             // 8x16 tiles per image
  int n_ntiles, BN, nkc;
  int P, Q, Nimg, ldo;
  int num_tiles;                    // pixel tiles x channel tiles (even count of pixel tiles)
  int ntaps, tpb, ntg;              // taps, taps per filter stage, filter stages per K-block
  int pw;                           // patch width in pixels (10: tile + halo; 16: padded to 2 swizzle atoms)
  int bstages;
  uint32_t patch_bytes, btile_bytes, bstage_bytes;  // slot strides (1 KB multiples)
  uint32_t patch_tx_bytes;          // bytes one patch box delivers (pw * 18 * KC * 2)
  int tap_rowoff[TC_MAX_TAPS];      // (dh + 1) * pw + (dw + 1)
  int tap_wcol[TC_MAX_TAPS];
  bf16* out;
  const bf16* residual;
  const float* bias;
  double* stats;                    // BN accumulator workspace (STATS kernels) or nullptr
  EpiStatsFinal fin;                // in-kernel finalize of the statistics (fin.mean == nullptr: off)
  int relu;                         // epilogue ReLU after bias / residual (eval with batch norm folded into the conv)
  EpiBnBwd bb;                      // STATS = 2 (residual then holds the BN input x and is NOT added)
  int det;                          // deterministic mode: fixed-order reduction of the epilogue statistics
};

// MT = pixel tiles per CTA that share every filter stage (MT accumulators of BN columns in TMEM, single-
// buffered when MT = 2, epilogue with 4 warps PER TILE). ncu on MT = 1: 330 MB per launch through the
// L2->SM path, 236 MB of it the same filter tiles fetched once per tile pair; MT = 2 halves that but was
// measured SLOWER (983 vs 1183 TFLOP/s at 160 channels): kept for experiments, default MT = 1.
// TAIL32 (with KC = 64): input-channel counts of the form 64 n + 32 (the 160-channel WRN layers). The first n
// K-blocks are 64 channels wide (128-byte rows, SWIZZLE_128B: ncu measured the tensor pipe at 98 % of its
// instruction rate while active) and ONE last block holds the remaining 32 channels on 64-byte rows (SWIZZLE_64B,
// 72 %): 80 % of the MMAs run at the fast rate instead of none (round 1 ran such layers entirely on 64-byte rows).
// The tail block has its own tensor maps (tmA32 / tmB32: 32-channel boxes) and reuses the patch / filter slots.
template <int KC, int MT, int STATS, bool TAIL32 = false>
__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_tc2h_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmA32, const __grid_constant__ CUtensorMap tmB32,
                 const __grid_constant__ ConvHaloArgs args) {
  static_assert(!TAIL32 || KC == 64, "the 32-channel tail block follows 64-channel blocks");
  constexpr int NBUF = (MT == 1) ? 2 : 1;   // accumulator sets
  const int nblk = args.nkc + (TAIL32 ? 1 : 0);   // K-blocks per unit
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t pfull_bar[2], pempty_bar[2];
  __shared__ uint64_t bfull_bar[HALO_BSTAGES_MAX], bempty_bar[HALO_BSTAGES_MAX];
  __shared__ uint64_t tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_part[STATS ? 2 * EPI_STATS_MAX_C : 1];
  static_assert(!STATS || MT == 1, "fused statistics: one pixel tile per CTA");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  const uint32_t pslot_bytes = MT * args.patch_bytes;
  uint8_t* bring = smem + 2 * (size_t)pslot_bytes;
  const int crank = (int)cluster_ctarank();
  const bool leader = crank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (TAIL32) {
      tma_prefetch_desc(&tmA32);
      tma_prefetch_desc(&tmB32);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&pfull_bar[i], 1);
      mbar_init(&pempty_bar[i], 1);
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 16);
    }
    for (int i = 0; i < args.bstages; ++i) {
      mbar_init(&bfull_bar[i], 1);
      mbar_init(&bempty_bar[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(&tmem_base_smem, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int tiles_img = args.tiles_w * args.tiles_h;
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_units = args.num_tiles / (2 * MT);   // one unit = MT pixel tiles per CTA of the pair
  const int bhalf = args.BN >> 1;
  const int pw = args.pw;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int ps = 0, bs = 0;
    uint32_t pph = 0, bph = 0;
    for (int ct = pair_id; ct < num_units; ct += num_pairs) {
      const int nt = ct % args.n_ntiles;
      const int mt0 = ((ct / args.n_ntiles) * 2 + crank) * MT;
      for (int kc = 0; kc < nblk; ++kc) {
        const bool tail = TAIL32 && kc == args.nkc;          // the 32-channel block: half the bytes
        const CUtensorMap* mapA = tail ? &tmA32 : &tmA;
        const CUtensorMap* mapB = tail ? &tmB32 : &tmB;
        mbar_wait(&pempty_bar[ps], pph ^ 1);
        if (elect_one()) {
          const uint32_t pfull_leader = map_to_cta(smem_u32(&pfull_bar[ps]), 0);
          if (leader) mbar_expect_tx(&pfull_bar[ps], (2u * MT * args.patch_tx_bytes) >> (tail ? 1 : 0));
#pragma unroll
          for (int t = 0; t < MT; ++t) {
            const int mt = mt0 + t;
            const int w0 = (mt % args.tiles_w) * 8;
            const int h0 = ((mt / args.tiles_w) % args.tiles_h) * 16;
            const int n0 = mt / tiles_img;
            tma_load_4d_2sm(smem + (size_t)ps * pslot_bytes + (size_t)t * args.patch_bytes, mapA,
                            pfull_leader, kc * KC, w0 - 1, h0 - 1, n0);
          }
        }
        __syncwarp();
        if (++ps == 2) { ps = 0; pph ^= 1; }
        for (int tg = 0; tg < args.ntg; ++tg) {
          mbar_wait(&bempty_bar[bs], bph ^ 1);
          if (elect_one()) {
            const uint32_t bfull_leader = map_to_cta(smem_u32(&bfull_bar[bs]), 0);
            const int t0 = tg * args.tpb;
            const int cnt = min(args.tpb, args.ntaps - t0);
            if (leader) mbar_expect_tx(&bfull_bar[bs], (2u * (uint32_t)cnt * args.btile_bytes) >> (tail ? 1 : 0));
            uint8_t* dst = bring + (size_t)bs * args.bstage_bytes;
            for (int j = 0; j < cnt; ++j)
              tma_load_2d_2sm(dst + (size_t)j * args.btile_bytes, mapB, bfull_leader,
                              args.tap_wcol[t0 + j] + kc * KC, nt * args.BN + crank * bhalf);
          }
          __syncwarp();
          if (++bs == args.bstages) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer (leader CTA only) =====================
      const uint32_t idesc = make_idesc_bf16(256, args.BN, 0, 0);
      const uint32_t bhi = smem_desc_hi(KMajorCfg<KC>::SBO, KMajorCfg<KC>::LAYOUT);
      const uint32_t ahi = smem_desc_hi((uint32_t)pw * KMajorCfg<KC>::ROW_BYTES, KMajorCfg<KC>::LAYOUT);
      const uint32_t bhi32 = smem_desc_hi(KMajorCfg<32>::SBO, KMajorCfg<32>::LAYOUT);   // tail block: 64-byte rows
      const uint32_t ahi32 = smem_desc_hi((uint32_t)pw * KMajorCfg<32>::ROW_BYTES, KMajorCfg<32>::LAYOUT);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t bring_base = smem_u32(bring);
      const uint32_t bstep = args.btile_bytes >> 4;
      const uint32_t pstep = args.patch_bytes >> 4;
      int ps = 0, bs = 0, as = 0;
      uint32_t pph = 0, bph = 0, aph = 0;
      for (int ct = pair_id; ct < num_units; ct += num_pairs) {
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)as * 256u;
        uint32_t acc = 0;
        for (int kc = 0; kc < nblk; ++kc) {
          const bool tail = TAIL32 && kc == args.nkc;
          mbar_wait(&pfull_bar[ps], pph);
          const uint32_t patch_lo = smem_desc_lo(smem_base + (uint32_t)ps * pslot_bytes, 16);
          for (int tg = 0; tg < args.ntg; ++tg) {
            mbar_wait(&bfull_bar[bs], bph);
            tc_fence_after();
            if (elect_one()) {
              const int t0 = tg * args.tpb;
              const int cnt = min(args.tpb, args.ntaps - t0);
              uint32_t blo = smem_desc_lo(bring_base + (uint32_t)bs * args.bstage_bytes, 16);
              if (!tail) {
                for (int j = 0; j < cnt; ++j) {
                  const uint32_t alo =
                      patch_lo + (((uint32_t)args.tap_rowoff[t0 + j] * KMajorCfg<KC>::ROW_BYTES) >> 4);
#pragma unroll
                  for (int k = 0; k < KC / 16; ++k) {
                    const uint64_t bd = smem_desc_join(blo + 2 * k, bhi);
#pragma unroll
                    for (int t = 0; t < MT; ++t)
                      umma_bf16_ss_2sm(d_addr + (uint32_t)(t * args.BN),
                                       smem_desc_join(alo + t * pstep + 2 * k, ahi), bd, idesc, acc);
                    acc = 1u;
                  }
                  blo += bstep;
                }
              } else {
                for (int j = 0; j < cnt; ++j) {
                  const uint32_t alo =
                      patch_lo + (((uint32_t)args.tap_rowoff[t0 + j] * KMajorCfg<32>::ROW_BYTES) >> 4);
#pragma unroll
                  for (int k = 0; k < 2; ++k) {
                    const uint64_t bd = smem_desc_join(blo + 2 * k, bhi32);
#pragma unroll
                    for (int t = 0; t < MT; ++t)
                      umma_bf16_ss_2sm(d_addr + (uint32_t)(t * args.BN),
                                       smem_desc_join(alo + t * pstep + 2 * k, ahi32), bd, idesc, acc);
                    acc = 1u;
                  }
                  blo += bstep;
                }
              }
              umma_commit_2sm_mcast(&bempty_bar[bs], 3);
              if (tg == args.ntg - 1) {
                umma_commit_2sm_mcast(&pempty_bar[ps], 3);
                if (kc == nblk - 1) umma_commit_2sm_mcast(&tfull_bar[as], 3);
              }
            }
            __syncwarp();
            if (++bs == args.bstages) { bs = 0; bph ^= 1; }
          }
          if (++ps == 2) { ps = 0; pph ^= 1; }
        }
        if (++as == NBUF) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own TMEM): 4 warps per pixel tile =====================
    // 8 epilogue warps: MT = 1 -> two warps per quadrant, each draining half of the columns;
    //                   MT = 2 -> four warps per pixel tile
    const int wq = warp & 3;                 // TMEM lane quadrant this warp may read
    const int grp = (warp - 2) >> 2;
    const int t = (MT == 2) ? grp : 0;       // pixel tile of this CTA handled by this warp
    const int e = (warp - 2) * 32 + lane;    // index among the epilogue threads
    const int cw = (MT == 2) ? args.BN : (args.BN >> 1);
    const int c_lo = (MT == 2) ? 0 : grp * cw, c_hi = c_lo + cw;
    const int m = wq * 32 + lane;
    const int wi = m & 7;
    const int hi = m >> 3;
    int as = 0;
    uint32_t aph = 0;
    if (STATS) {
      for (int i = e; i < 2 * EPI_STATS_MAX_C; i += TC2_EPI_THREADS) s_part[i] = 0.f;
      epi_bar();
    }
    float racc[STATS ? EPI_MAX_CHUNKS : 1];   // this lane's running sums, one per 16-column chunk of its warp
#pragma unroll
    for (int i = 0; i < (STATS ? EPI_MAX_CHUNKS : 1); ++i) racc[i] = 0.f;
    int racc_nt = -1;                         // channel tile they belong to
    auto unit_row = [&](int ct, int& nt) -> size_t {
      nt = ct % args.n_ntiles;
      const int mt = ((ct / args.n_ntiles) * 2 + crank) * MT + t;
      const int w = (mt % args.tiles_w) * 8 + wi;
      const int h = ((mt / args.tiles_w) % args.tiles_h) * 16 + hi;
      const int n = mt / tiles_img;
      const size_t pix = ((size_t)n * args.P + h) * args.Q + w;
      return pix * (size_t)args.ldo + (size_t)nt * args.BN;
    };
    auto load_residual = [&](Vec8* dst, size_t off) {
#pragma unroll
      for (int i = 0; i < EPI_RES_VECS / 2; ++i)
        if (c_lo + 16 * i < c_hi) ldg256_stream(args.residual + off + c_lo + 16 * i, dst[2 * i].raw, dst[2 * i + 1].raw);
    };
    Vec8 rcur[EPI_RES_VECS], rnext[EPI_RES_VECS];
    // STATS = 2: the mask bytes of this thread's row slice, 16 columns per 16-bit load, prefetched like the residual
    uint32_t mcur[STATS == 2 ? EPI_RES_VECS / 2 : 1], mnext[STATS == 2 ? EPI_RES_VECS / 2 : 1];
    auto load_mask = [&](uint32_t* dst, size_t off) {
      const uint16_t* mp = reinterpret_cast<const uint16_t*>(args.bb.mask) + ((off + c_lo) >> 4);
#pragma unroll
      for (int i = 0; i < EPI_RES_VECS / 2; ++i)
        if (c_lo + 16 * i < c_hi) dst[STATS == 2 ? i : 0] = __ldg(mp + i);
    };
    const bool prefetch_res = args.residual != nullptr;
    if (prefetch_res && pair_id < num_units) {
      int nt0;
      const size_t off0 = unit_row(pair_id, nt0);
      load_residual(rcur, off0);
      if (STATS == 2) load_mask(mcur, off0);
    }
    for (int ct = pair_id; ct < num_units; ct += num_pairs) {
      int nt;
      const size_t off = unit_row(ct, nt);
      if (STATS && nt != racc_nt) {
        if (racc_nt >= 0) epi_stats_regs_flush<EPI_MAX_CHUNKS>(racc, s_part + 2 * racc_nt * args.BN, c_lo, c_hi, lane,
                                                              args.det ? wq : -1);
        racc_nt = nt;
      }
      bf16* orow = args.out + off;
      const bf16* rrow = args.residual ? args.residual + off : nullptr;
      const float* brow = args.bias ? args.bias + (size_t)nt * args.BN : nullptr;

      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      if (prefetch_res && ct + num_pairs < num_units) {   // next unit's residual row: in flight while this one drains
        int ntn;
        const size_t offn = unit_row(ct + num_pairs, ntn);
        load_residual(rnext, offn);
        if (STATS == 2) load_mask(mnext, offn);
      }
      const uint32_t t_addr =
          tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)as * 256u + (uint32_t)(t * args.BN);
#pragma unroll
      for (int ci = 0; ci < EPI_MAX_CHUNKS * 2; ++ci) {
        const int c = c_lo + 16 * ci;
        if (c >= c_hi) break;
        const int ri = (2 * ci + 1 < EPI_RES_VECS) ? 2 * ci : 0;   // (a fused residual implies <= 5 chunks)
        uint32_t v[16];
        tmem_ld16(t_addr + c, v);
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
        if (brow) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] += round_bf16(__ldg(brow + c + j));
        }
        if (rrow && STATS != 2) {
          float rf[16];
          rcur[ri].to_float(rf);
          rcur[ri + 1].to_float(rf + 8);
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = round_bf16(f[j]) + rf[j];
        }
        if (args.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (STATS) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = round_bf16(f[j]);  // statistics of the values stored
        }
        Vec8 o0, o1;
        o0.from_float(f);
        o1.from_float(f + 8);
        stg256(orow + c, o0.raw, o1.raw);
        if (STATS == 1) racc[(STATS && ci < EPI_MAX_CHUNKS) ? ci : 0] += epi_stats_chunk(f, lane);
        if (STATS == 2)
          racc[(STATS && ci < EPI_MAX_CHUNKS) ? ci : 0] +=
              epi_bnbwd_chunk(f, rcur[ri], rcur[ri + 1], mcur[(STATS == 2 && ci < EPI_RES_VECS / 2) ? ci : 0], true,
                              args.bb, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tempty_bar[as]), 0));
      if (prefetch_res) {
#pragma unroll
        for (int i = 0; i < EPI_RES_VECS; ++i) rcur[i] = rnext[i];
        if (STATS == 2) {
#pragma unroll
          for (int i = 0; i < EPI_RES_VECS / 2; ++i) mcur[STATS == 2 ? i : 0] = mnext[STATS == 2 ? i : 0];
        }
      }
      if (++as == NBUF) { as = 0; aph ^= 1; }
    }
    if (STATS) {
      if (racc_nt >= 0) epi_stats_regs_flush<EPI_MAX_CHUNKS>(racc, s_part + 2 * racc_nt * args.BN, c_lo, c_hi, lane,
                                                              args.det ? wq : -1);
      epi_stats_flush(s_part, args.stats, args.ldo, e);
      epi_stats_finalize<STATS == 2>(args.fin, args.stats, args.ldo, e);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// -------------------------------------------------------------------------------------------------
// wgrad
// -------------------------------------------------------------------------------------------------
struct WgradTcArgs {
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n;
  int num_ptiles;        // pixel tiles in total
  int kmmas;             // 16-pixel MMAs per pixel tile (rows_valid / 16)
  int nslabs_total;      // ntaps * ceil(Cin / SL)
  int slabs_per_tap;     // ceil(Cin / SL); the last slab of a tap may be partial (TMA zero-fills it)
  int n_mtiles;          // ceil(nslabs_total / (128 / SL))
  int n_mgroups;         // ceil(n_mtiles / MT), padded to the cluster size
  int n_ntiles;          // Cout / BN
  int cin;               // input channels
  int nb;                // dY slabs per stage = ceil(BN / SL)
  int BN;
  int splits;            // pixel-range splits (gridDim.y)
  int ktot;              // ntaps * Cin: row pitch of dw
  int stages;
  int tmem_cols;
  int cstride;           // conv stride: x window origin = dY pixel * cstride + tap displacement
  uint32_t stage_bytes;
  uint32_t slab_bytes;   // rows per stage * SL * 2 (also the descriptor's leading byte offset)
  TapTable taps;
  float* dw;
  // deterministic mode: pixel-range split y stores its partial filter gradient at part + y * part_stride (plain
  // stores, same indexing as dw) and wgrad_reduce_splits_kernel sums the splits in a fixed order; nullptr: the
  // splits add into a zeroed dw with fp32 atomics, in whatever order they finish
  float* part;
  size_t part_stride;
};

template <int SL, int ES = 2>
struct MnMajorCfg {
  static constexpr uint32_t ROW_BYTES = SL * ES;
  static constexpr uint32_t SBO = 8 * ROW_BYTES;              // next group of 8 pixels
  static constexpr uint32_t LBO = 128 * ROW_BYTES;            // next channel slab (128-pixel slabs)
  static constexpr uint32_t LAYOUT = (ROW_BYTES == 128) ? 2u : (ROW_BYTES == 64) ? 4u : 6u;
  static constexpr int SLABS_PER_MTILE = 128 / SL;
  static constexpr int KPIX = 32 / ES;                        // pixels reduced by one MMA: 16 bf16 / 8 tf32
};

// MT = M tiles (128 rows of (tap, channel) each) per CTA: they share every dY stage and accumulate
// in MT TMEM accumulators, so per MMA the CTA pulls (MT*4 KB + 5 KB)/MT instead of 9 KB through L2/TMA.
// CS = cluster size: the CS CTAs of a cluster own consecutive M-tile groups of the same (N tile, pixel
// range); the dY slabs they all need are loaded once (slab i by CTA i % CS) and multicast.
// TF32 = true: fp32 x and dY (kind::tf32, 8 pixels per MMA); SL then counts fp32 channels per slab.
template <int SL, int CS, int MT, bool TF32 = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy,
                const __grid_constant__ WgradTcArgs args) {
  using Cfg = MnMajorCfg<SL, TF32 ? 4 : 2>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[TC_MAX_STAGES];
  __shared__ uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ uint64_t tfull_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDy);
    for (int s = 0; s < args.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CS);
    }
    mbar_init(&tfull_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, (uint32_t)args.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();
  tc_fence_after();
  const int crank = (CS > 1) ? (int)cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << CS) - 1u);
  const uint32_t tmem_base = tmem_base_smem;

  const int mg = blockIdx.x % args.n_mgroups;   // group of MT M tiles
  const int nt = blockIdx.x / args.n_mgroups;
  const int per = (args.num_ptiles + args.splits - 1) / args.splits;
  const int pt0 = blockIdx.y * per;
  const int pt1 = min(args.num_ptiles, pt0 + per);
  const int tiles_hw = args.tiles_w * args.tiles_h;
  const int nb = args.nb;                                              // dY slabs per stage
  const uint32_t a_tile_bytes = Cfg::SLABS_PER_MTILE * args.slab_bytes;
  const uint32_t b_off = MT * a_tile_bytes;
  int q0[MT], na[MT], na_total = 0;
#pragma unroll
  for (int j = 0; j < MT; ++j) {
    q0[j] = (mg * MT + j) * Cfg::SLABS_PER_MTILE;
    na[j] = max(0, min(Cfg::SLABS_PER_MTILE, args.nslabs_total - q0[j]));  // 0: padding tile
    na_total += na[j];
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    // Everything that does not depend on the pixel tile is hoisted out of the loop: the producer issues
    // 9+ TMA copies per stage and was the bottleneck when it re-derived (tap, chunk) per copy with
    // integer divisions (ncu: the MMA warp starved while the producer never waited for a free slot).
    int s = 0;
    uint32_t ph = 0;
    const uint32_t tx = (uint32_t)(na_total + nb) * (uint32_t)(args.kmmas * Cfg::KPIX) * Cfg::ROW_BYTES;
    int sc[MT][Cfg::SLABS_PER_MTILE], sw[MT][Cfg::SLABS_PER_MTILE], sh[MT][Cfg::SLABS_PER_MTILE],
        sn[MT][Cfg::SLABS_PER_MTILE];
#pragma unroll
    for (int j = 0; j < MT; ++j) {
#pragma unroll
      for (int i = 0; i < Cfg::SLABS_PER_MTILE; ++i) {
        const int q = min(q0[j] + i, args.nslabs_total - 1);
        const int t = q / args.slabs_per_tap;
        sc[j][i] = (q - t * args.slabs_per_tap) * SL;
        sw[j][i] = args.taps.dw[t];
        sh[j][i] = args.taps.dh[t];
        sn[j][i] = args.taps.dn[t];
      }
    }
    const int bc0 = nt * args.BN;
    // pixel-tile coordinates advance incrementally (w fastest, then h, then image group)
    int tw = pt0 % args.tiles_w;
    int th = (pt0 / args.tiles_w) % args.tiles_h;
    int tn = pt0 / tiles_hw;
    for (int pt = pt0; pt < pt1; ++pt) {
      const int w0 = tw * args.bw, h0 = th * args.bh, n0 = tn * args.bn;
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (elect_one()) {
        uint8_t* a_dst = smem + (size_t)s * args.stage_bytes;
        uint8_t* b_dst = a_dst + b_off;
        mbar_expect_tx(&full_bar[s], tx);
#pragma unroll
        for (int j = 0; j < MT; ++j) {
#pragma unroll
          for (int i = 0; i < Cfg::SLABS_PER_MTILE; ++i) {
            if (i < na[j])
              tma_load_4d(a_dst + (size_t)j * a_tile_bytes + (size_t)i * args.slab_bytes, &tmX,
                          &full_bar[s], sc[j][i], w0 * args.cstride + sw[j][i], h0 * args.cstride + sh[j][i],
                          n0 + sn[j][i]);
          }
        }
        if (CS == 1) {
          for (int i = 0; i < nb; ++i)
            tma_load_4d(b_dst + (size_t)i * args.slab_bytes, &tmDy, &full_bar[s], bc0 + i * SL, w0, h0,
                        n0);
        } else {
          for (int i = crank; i < nb; i += CS)
            tma_load_4d_mcast(b_dst + (size_t)i * args.slab_bytes, &tmDy, &full_bar[s], bc0 + i * SL,
                              w0, h0, n0, cmask);
        }
      }
      __syncwarp();
      if (++s == args.stages) { s = 0; ph ^= 1; }
      if (++tw == args.tiles_w) {
        tw = 0;
        if (++th == args.tiles_h) { th = 0; ++tn; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(128, args.BN, 1, 1, TF32 ? 2u : 1u);
    const uint32_t dhi = smem_desc_hi(Cfg::SBO, Cfg::LAYOUT);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t kstep = ((uint32_t)Cfg::KPIX * Cfg::ROW_BYTES) >> 4;  // descriptor-address units per MMA
    const uint32_t tile_step = a_tile_bytes >> 4;
    int s = 0;
    uint32_t ph = 0;
    for (int pt = pt0; pt < pt1; ++pt) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_base + (uint32_t)s * args.stage_bytes;
        const uint32_t alo = smem_desc_lo(a_addr, args.slab_bytes);
        const uint32_t blo = smem_desc_lo(a_addr + b_off, args.slab_bytes);
        const uint32_t acc = pt > pt0 ? 1u : 0u;
        for (int k = 0; k < args.kmmas; ++k) {
          const uint64_t bd = smem_desc_join(blo + k * kstep, dhi);
#pragma unroll
          for (int j = 0; j < MT; ++j) {
            if (TF32)
              umma_tf32_ss(tmem_base + (uint32_t)j * 256u,
                           smem_desc_join(alo + j * tile_step + k * kstep, dhi), bd, idesc,
                           (acc | (uint32_t)k) != 0 ? 1u : 0u);
            else
              umma_bf16_ss(tmem_base + (uint32_t)j * 256u,
                           smem_desc_join(alo + j * tile_step + k * kstep, dhi), bd, idesc,
                           (acc | (uint32_t)k) != 0 ? 1u : 0u);
          }
        }
        if (CS == 1) umma_commit(&empty_bar[s]);
        else umma_commit_mcast(&empty_bar[s], cmask);
        if (pt == pt1 - 1) umma_commit(&tfull_bar);
      }
      __syncwarp();
      if (++s == args.stages) { s = 0; ph ^= 1; }
    }
    if (pt1 <= pt0) {
      if (elect_one()) umma_commit(&tfull_bar);
      __syncwarp();
    }
  } else if (pt1 > pt0) {
    // epilogue: lane m of accumulator j is (slab, channel) = (m / SL, m % SL) of M tile j
    const int wq = warp & 3;
    const int m = wq * 32 + lane;
    mbar_wait(&tfull_bar, 0);
    tc_fence_after();
#pragma unroll
    for (int j = 0; j < MT; ++j) {
      const int q = q0[j] + m / SL;
      bool valid = q < args.nslabs_total;
      int col = 0;
      if (valid) {
        const int t = q / args.slabs_per_tap;
        const int ck = q % args.slabs_per_tap;
        const int c = ck * SL + (m % SL);
        valid = c < args.cin;  // rows of a partial last slab
        col = args.taps.wcol[t] + c;
      }
      float* drow = (args.part ? args.part + (size_t)blockIdx.y * args.part_stride : args.dw) +
                    (size_t)nt * args.BN * args.ktot + col;
      const uint32_t t_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)j * 256u;
      for (int c = 0; c < args.BN; c += 16) {
        uint32_t v[16];
        tmem_ld16(t_addr + c, v);
        tmem_ld_wait();
        if (valid) {
          if (args.splits == 1 || args.part) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj)
              drow[(size_t)(c + jj) * args.ktot] = __uint_as_float(v[jj]);
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj)
              atomicAdd(drow + (size_t)(c + jj) * args.ktot, __uint_as_float(v[jj]));
          }
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)args.tmem_cols);
  }
}

// -------------------------------------------------------------------------------------------------
// wgrad_tc2h_kernel: SM-pair wgrad with HALO REUSE for 3x3 / stride 1 / pad 1 filters.
//
// The plain wgrad kernel above fetches one 128-pixel x window per (tap, 32-channel chunk) slab, i.e.
// every activation byte nine times, and is bound by the L2->SM (TMA) path (ncu: ~81 B/clk/SM asked,
// ~50 delivered). Here one TMA box per (filter row r, chunk) holds the pixel tile widened by the
// horizontal halo: [bn*bh = 16 pixel rows][PW columns w0-1 ..][32 channels], and the three taps
// s = 0..2 of that filter row are the SAME box read through descriptors that start s pixel rows (64 B)
// later (UMMA swizzling follows absolute shared-memory address bits: probed on the hardware in round 1). The
// 8-pixel groups of the MMA K dimension are image rows, so their stride (SBO) is PW pixel rows.
//
// Work decomposition: a "combo" is (chunk, r); a column is 4 combos = the 4 channel slabs of a
// 128-row M tile (slab stride LBO = one box); accumulator j of a CTA is filter column s = j of its
// column, so a CTA owns 3 accumulators of 128 x BN (<= 480 TMEM columns) and issues 24 MMAs per
// 128-pixel stage from 4 boxes + its half of the dY tile. Two columns form an SM pair
// (cta_group::2, M = 256): each CTA stages its own boxes and HALF of the dY tile, which halves the dY
// traffic and the shared-memory read rate per MMA (the limit of the 1-CTA kernel).
// -------------------------------------------------------------------------------------------------
constexpr int WGH_STAGES_MAX = 4;

struct WgradHaloArgs {
  int bh, bn, pw;          // pixel tile = 8 x bh x bn (= 128 pixels); patch width in pixels (>= 10)
  int tiles_w, tiles_h;
  int num_ptiles;
  int ncombo;              // 3 * C / 32
  int ncols;               // ceil(ncombo / 4) rounded up to even
  int n_ntiles, BN, nbh;   // Cout tiles; dY slabs per CTA = ceil(BN / 2 / 32)
  int splits, ktot, stages;
  uint32_t box_bytes, slab_bytes, stage_bytes;
  int wcol[TC_MAX_TAPS];
  float* dw;
  float* part;             // deterministic mode, see WgradTcArgs
  size_t part_stride;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
wgrad_tc2h_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy,
                  const __grid_constant__ WgradHaloArgs args) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[WGH_STAGES_MAX];
  __shared__ uint64_t empty_bar[WGH_STAGES_MAX];
  __shared__ uint64_t tfull_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  const int crank = (int)cluster_ctarank();
  const bool leader = crank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDy);
    for (int s = 0; s < args.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tfull_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(&tmem_base_smem, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int col = blockIdx.x % args.ncols;   // ncols is even: a cluster is columns (2p, 2p + 1)
  const int nt = blockIdx.x / args.ncols;
  const int per = (args.num_ptiles + args.splits - 1) / args.splits;
  const int pt0 = blockIdx.y * per;
  const int pt1 = min(args.num_ptiles, pt0 + per);
  const int q0 = col * 4;
  const int nbox = max(0, min(4, args.ncombo - q0));
  const int nbox_peer = max(0, min(4, args.ncombo - (col ^ 1) * 4));
  const uint32_t b_off = 4u * args.box_bytes;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int s = 0;
    uint32_t ph = 0;
    const uint32_t tx = (uint32_t)(nbox + nbox_peer) * args.box_bytes + 2u * (uint32_t)args.nbh * args.slab_bytes;
    int bc[4], bdh[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = min(q0 + i, args.ncombo - 1);
      bc[i] = (q / 3) * 32;
      bdh[i] = (q % 3) - 1;
    }
    const int bcol0 = nt * args.BN + crank * (args.BN >> 1);
    const int tiles_hw = args.tiles_w * args.tiles_h;
    int tw = pt0 % args.tiles_w;
    int th = (pt0 / args.tiles_w) % args.tiles_h;
    int tn = pt0 / tiles_hw;
    for (int pt = pt0; pt < pt1; ++pt) {
      const int w0 = tw * 8, h0 = th * args.bh, n0 = tn * args.bn;
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (elect_one()) {
        const uint32_t full_leader = map_to_cta(smem_u32(&full_bar[s]), 0);
        if (leader) mbar_expect_tx(&full_bar[s], tx);
        uint8_t* a_dst = smem + (size_t)s * args.stage_bytes;
        uint8_t* b_dst = a_dst + b_off;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < nbox)
            tma_load_4d_2sm(a_dst + (size_t)i * args.box_bytes, &tmX, full_leader, bc[i], w0 - 1,
                            h0 + bdh[i], n0);
        }
        for (int i = 0; i < args.nbh; ++i)
          tma_load_4d_2sm(b_dst + (size_t)i * args.slab_bytes, &tmDy, full_leader, bcol0 + i * 32, w0,
                          h0, n0);
      }
      __syncwarp();
      if (++s == args.stages) { s = 0; ph ^= 1; }
      if (++tw == args.tiles_w) {
        tw = 0;
        if (++th == args.tiles_h) { th = 0; ++tn; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer (leader CTA only) =====================
      const uint32_t idesc = make_idesc_bf16(256, args.BN, 1, 1);
      const uint32_t ahi = smem_desc_hi((uint32_t)args.pw * 64u, 4u);   // next image row of the patch
      const uint32_t bhi = smem_desc_hi(8u * 64u, 4u);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t a_kstep = ((uint32_t)args.pw * 128u) >> 4;         // two image rows per K = 16
      const uint32_t b_kstep = (16u * 64u) >> 4;
      int s = 0;
      uint32_t ph = 0;
      for (int pt = pt0; pt < pt1; ++pt) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_base + (uint32_t)s * args.stage_bytes;
          const uint32_t alo = smem_desc_lo(a_addr, args.box_bytes);
          const uint32_t blo = smem_desc_lo(a_addr + b_off, args.slab_bytes);
          const uint32_t acc = pt > pt0 ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t bd = smem_desc_join(blo + k * b_kstep, bhi);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              umma_bf16_ss_2sm(tmem_base + (uint32_t)(j * args.BN),
                               smem_desc_join(alo + 4u * j + k * a_kstep, ahi), bd, idesc,
                               (acc | (uint32_t)k) != 0 ? 1u : 0u);
            }
          }
          umma_commit_2sm_mcast(&empty_bar[s], 3);
          if (pt == pt1 - 1) umma_commit_2sm_mcast(&tfull_bar, 3);
        }
        __syncwarp();
        if (++s == args.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (pt1 > pt0) {
    // ===================== epilogue (both CTAs): warp quadrant wq = slab = combo q0 + wq ============
    const int wq = warp & 3;
    if (wq < nbox) {
      const int q = q0 + wq;
      const int c = (q / 3) * 32 + lane;
      const int r = q % 3;
      mbar_wait(&tfull_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        float* drow = (args.part ? args.part + (size_t)blockIdx.y * args.part_stride : args.dw) +
                      (size_t)nt * args.BN * args.ktot + args.wcol[r * 3 + j] + c;
        const uint32_t t_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(j * args.BN);
        for (int cc = 0; cc < args.BN; cc += 16) {
          uint32_t v[16];
          tmem_ld16(t_addr + cc, v);
          tmem_ld_wait();
          if (args.splits == 1 || args.part) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj)
              drow[(size_t)(cc + jj) * args.ktot] = __uint_as_float(v[jj]);
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj)
              atomicAdd(drow + (size_t)(cc + jj) * args.ktot, __uint_as_float(v[jj]));
          }
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}
// Deterministic mode: out[i] = part[0][i] + part[1][i] + ... + part[splits-1][i], in that order (fp32), for the
// filter-gradient partials of the pixel-range splits and for the per-block partials of the bias gradient.
__global__ void wgrad_reduce_splits_kernel(const float* __restrict__ part, float* __restrict__ out, size_t n,
                                           size_t stride, int splits) {
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  if ((n & 3) == 0 && (stride & 3) == 0 && (((uintptr_t)part | (uintptr_t)out) & 15) == 0) {
    for (size_t i = i0; i < (n >> 2); i += step) {
      float4 a = __ldcs(reinterpret_cast<const float4*>(part) + i);
      for (int s = 1; s < splits; ++s) {
        const float4 b = __ldcs(reinterpret_cast<const float4*>(part + (size_t)s * stride) + i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      reinterpret_cast<float4*>(out)[i] = a;
    }
  } else {
    for (size_t i = i0; i < n; i += step) {
      float a = part[i];
      for (int s = 1; s < splits; ++s) a += part[(size_t)s * stride + i];
      out[i] = a;
    }
  }
}

}  // namespace b200
