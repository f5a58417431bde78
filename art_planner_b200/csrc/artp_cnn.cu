// art_planner_b200/csrc/artp_cnn.cu -- the learned motion-cost network on sm_100a.
//
// Reference: art_planner_motion_cost/src/art_planner_motion_cost/predictor/network_light.py
//   CNNpart :78-110  conv3x3(1->24)+BN, conv3x3(24->24)+BN+LReLU(0.3), maxpool 2/2, conv3x3(24->48)+BN+LReLU,
//                    conv3x3(48->48)+BN+LReLU, maxpool 3/1, conv3x3(48->48)+BN+LReLU, conv15x15(48->48)+BN+LReLU
//   FCpart  :113-165 and CostQuery.__call__ (cost_query.py:39-69), server centring (cost_query_server.py:160-161)
//
// Numerics: the reference evaluates in fp16; parity here is against the fp32 evaluation of the same module to 1e-4
// relative, so everything accumulates in fp32 and the 15x15 convolution -- 83.6 % of the FLOPs, implicit GEMM
// M = output pixels, N = 48, K = 225 taps x 48 channels -- runs on the 5th-gen tensor cores with an error-compensated
// fp16 split (a = a_hi + a_lo, w = w_hi + w_lo; D += a_hi*w_hi + a_hi*w_lo + a_lo*w_hi, fp32 accumulate in TMEM),
// which keeps ~22 mantissa bits per product.
//
// Tensor-core kernel (conv_tc_kernel<KS,...>, described for the 15x15 layer): one CTA per 16(y) x 8(x) output tile
// (UMMA M = 128, N = 48).
//   * The whole input halo brick [30 y][24 x][64 ch] (hi and lo, 92 KB each) is TMA-loaded ONCE into 128B-swizzled
//     shared memory; pixels are 128-byte rows, image rows are 24 pixels = 3072 B = 3 swizzle atoms apart, so every one
//     of the 225 taps is just a shifted view of the same brick: start address += (ky*24 + kx)*128 B, stride between
//     8-pixel groups (SBO) = 3072 B. Measured on B200: the 128B swizzle XOR is applied on absolute shared-memory
//     address bits, so the shifted (128 B-aligned, not 1024 B-aligned) start needs descriptor base_offset = 0
//     (setting it to (shift & 7) double-swizzles; profiles/debug_conv15.py). No per-tap activation traffic.
//   * Weights [tap][48][64] fp16 hi/lo stream through a 3-stage TMA ring (12 KB per tap).
//   * Warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one elected lane), warps 2..5 = epilogue
//     (tcgen05.ld -> +bias -> LeakyReLU -> fp32 NHWC feature map).
// Layers 2..5 (3x3) use the same kernel with an 18 x 16-pixel brick and 9 taps, writing either fp32 NHWC (before a
// max-pool) or directly the next layer's fp16 hi/lo NHWC-64 input; layer 1 (Cin = 1) stays on CUDA cores.
// A complete fp32 CUDA-core path (conv3x3_kernel, conv15_reference_kernel) is kept as the in-library cross-check.
// This translation unit is compiled WITHOUT -fmad=false (no bit-exactness requirement here).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "artp_cnn.h"

namespace artp_cnn {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
#define CNN_TRY(expr)                                                              \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(_e); return -3; } \
  } while (0)

struct LayerDef { int cout, cin, k; bool bn; };
static const LayerDef kLayers[14] = {
    {24, 1, 3, true},  {24, 24, 3, true}, {48, 24, 3, true}, {48, 48, 3, true}, {48, 48, 3, true}, {48, 48, 15, true},
    {16, 10, 1, true}, {48, 64, 1, true}, {24, 48, 1, true}, {24, 48, 1, true}, {36, 48, 1, true},
    {1, 24, 1, false}, {1, 24, 1, false}, {1, 36, 1, false}};
static const float kBnEps = 1e-5f;
__device__ constexpr float kWScale = 1024.0f;   // undone exactly in the 15x15 epilogue

size_t blob_floats() {
  size_t n = 0;
  for (const auto& l : kLayers) n += (size_t)l.cout * l.cin * l.k * l.k + (l.bn ? 4 * l.cout : l.cout);
  return n;
}

// ------------------------------------------------------------------------------------------------
// fp32 direct 3x3 convolution, NHWC, BN folded, optional LeakyReLU(0.3). One thread per output pixel, all COUT
// accumulators in registers; input tile and weight slice staged in shared memory per chunk of CK input channels.
// SRC_MAP: the input is the heightfield layer as artp_set_map stores it (H[x + z*pitch] = layer(x, nz-1-z));
// the network input E[r][c] = layer(rows-1-r, cols-1-c) = H[(rows-1-r) + c*pitch] (cost_query_server.py:74).
// ------------------------------------------------------------------------------------------------
template <int CIN, int COUT, int CK, bool ACT, bool SRC_MAP>
__global__ void __launch_bounds__(256) conv3x3_kernel(const float* __restrict__ in, int H, int W, int in_pitch,
                                                      const float* __restrict__ wf /*[9][CIN][COUT]*/,
                                                      const float* __restrict__ bias, float* __restrict__ out) {
  constexpr int T = 16;
  __shared__ float tile[(T + 2) * (T + 2) * CK];
  __shared__ __align__(16) float wsm[9 * CK * COUT];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int ox0 = blockIdx.x * T, oy0 = blockIdx.y * T;
  const int OH = H - 2, OW = W - 2;
  float acc[COUT];
#pragma unroll
  for (int i = 0; i < COUT; ++i) acc[i] = 0.0f;
  for (int c0 = 0; c0 < CIN; c0 += CK) {
    for (int i = threadIdx.x; i < (T + 2) * (T + 2) * CK; i += 256) {
      const int ci = i % CK, p = i / CK, px = p % (T + 2), py = p / (T + 2);
      const int iy = oy0 + py, ix = ox0 + px;
      float v = 0.0f;
      if (iy < H && ix < W) {
        if (SRC_MAP) v = __ldg(in + (size_t)ix * in_pitch + (H - 1 - iy));   // E[iy][ix], rows = H (x), cols = W (y)
        else v = __ldg(in + ((size_t)iy * W + ix) * CIN + c0 + ci);
      }
      tile[i] = v;
    }
    for (int i = threadIdx.x; i < 9 * CK * COUT; i += 256) {
      const int co = i % COUT, r = i / COUT, ci = r % CK, tap = r / CK;
      wsm[i] = __ldg(wf + ((size_t)tap * CIN + c0 + ci) * COUT + co);
    }
    __syncthreads();
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
      const float* tp = tile + ((ty + ky) * (T + 2) + tx + kx) * CK;
#pragma unroll
      for (int ci = 0; ci < CK; ++ci) {
        const float v = tp[ci];
        const float4* wp = reinterpret_cast<const float4*>(wsm + (tap * CK + ci) * COUT);
#pragma unroll
        for (int q = 0; q < COUT / 4; ++q) {
          const float4 w4 = wp[q];
          acc[4 * q + 0] = fmaf(v, w4.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
        }
      }
    }
    __syncthreads();
  }
  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy < OH && ox < OW) {
    float* op = out + ((size_t)oy * OW + ox) * COUT;
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      float v = acc[co] + __ldg(bias + co);
      if (ACT) v = v > 0.0f ? v : 0.3f * v;
      op[co] = v;
    }
  }
}

// max pooling, NHWC fp32: K x K window, stride S (2/2 and 3/1 in the reference).
__global__ void maxpool_kernel(const float* __restrict__ in, int H, int W, int C, int K, int S, float* __restrict__ out,
                               int OH, int OW) {
  const size_t total = (size_t)OH * OW * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t p = i / C;
    const int ox = (int)(p % OW), oy = (int)(p / OW);
    float m = -INFINITY;
    for (int dy = 0; dy < K; ++dy)
      for (int dx = 0; dx < K; ++dx) m = fmaxf(m, in[((size_t)(oy * S + dy) * W + ox * S + dx) * C + c]);
    out[i] = m;
  }
}

// Fold BN into conv weights: wf[tap][cin][cout] = w[cout][cin][tap] * gamma/sqrt(var+eps); bias = beta - mean*scale.
__global__ void fold_conv_kernel(const float* __restrict__ w, const float* __restrict__ bn /*gamma,beta,mean,var*/, int cout,
                                 int cin, int kk, float* __restrict__ wf, float* __restrict__ bias) {
  const int total = cout * cin * kk;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % cout, r = i / cout, ci = r % cin, tap = r / cin;
    const float s = bn[co] / sqrtf(bn[3 * cout + co] + kBnEps);
    wf[i] = w[((size_t)co * cin + ci) * kk + tap] * s;
  }
  for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < cout; co += gridDim.x * blockDim.x) {
    const float s = bn[co] / sqrtf(bn[3 * cout + co] + kBnEps);
    bias[co] = bn[cout + co] - bn[2 * cout + co] * s;
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMA / mbarrier primitives (inline PTX, sm_100a)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major, 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);             // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version 1 (sm_100), bits [46,48)
  d |= (uint64_t)(base_offset & 7) << 49;              // matrix base offset, bits [49,52)
  d |= (uint64_t)2 << 61;                              // layout type SWIZZLE_128B, bits [61,64)
  return d;
}

constexpr int kTileY = 16, kTileX = 8;           // output tile (UMMA M = 128 = 16 groups of 8 pixels along x)
constexpr int kWStages = 3;

template <int KS, int NOUT>
struct ConvCfg {
  static constexpr int kBrickY = kTileY + KS - 1;
  static constexpr int kBrickX = ((kTileX + KS - 1) + 7) & ~7;   // row pitch = whole swizzle atoms (8 pixels = 1024 B)
  static constexpr int kBrickBytes = kBrickY * kBrickX * 128;    // per split term
  static constexpr int kWStageBytes = 2 * NOUT * 128;            // hi + lo weight tile of one tap
  static constexpr int kSmem = 2 * kBrickBytes + kWStages * kWStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  // instruction descriptor: D = f32 (bits [4,6) = 1), A = B = f16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
  static constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(NOUT >> 3) << 17) | ((128u >> 4) << 24);
};

// Implicit-GEMM convolution on tcgen05 (see the file header). KS x KS taps, KSTEPS x 16 input channels multiplied
// per tap (channels are stored padded to 64 = one 128-byte swizzle row per pixel), NOUT = UMMA N (multiple of 16),
// NMAIN fp32 accumulators for the a_hi*w_hi products (tap t -> t % NMAIN) + 1 for the two correction terms.
// SPLIT_OUT: write the activation as fp16 hi/lo NHWC-64 (the next layer's TMA source) instead of fp32 NHWC.
template <int KS, int KSTEPS, int NOUT, int NMAIN, bool SPLIT_OUT, int CLUSTER, int ISSUERS>
__global__ void __launch_bounds__(192 + 32 * (ISSUERS - 1))
conv_tc_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
               const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo,
               const float* __restrict__ bias, float* __restrict__ out, __half* __restrict__ out_hi,
               __half* __restrict__ out_lo, int OH, int OW, int cout, int use_base_offset) {
  using Cfg = ConvCfg<KS, NOUT>;
  constexpr int kTaps = KS * KS;
  constexpr int kAcc = NMAIN + ISSUERS;   // ISSUERS correction accumulators (one per MMA-issuing warp)
  constexpr int kTmemCols = kAcc * 64 <= 64 ? 64 : (kAcc * 64 <= 128 ? 128 : (kAcc * 64 <= 256 ? 256 : 512));
  static_assert(kAcc * 64 <= 512, "too many accumulators");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_hi = smem;
  unsigned char* a_lo = smem + Cfg::kBrickBytes;
  unsigned char* w_st = smem + 2 * Cfg::kBrickBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_st + kWStages * Cfg::kWStageBytes);
  uint64_t* bar_brick = bars;                 // TMA -> MMA: activations landed
  uint64_t* bar_full = bars + 1;              // [kWStages] TMA -> MMA: weight tap landed
  uint64_t* bar_empty = bars + 1 + kWStages;  // [kWStages] MMA -> TMA: stage consumed
  uint64_t* bar_done = bars + 1 + 2 * kWStages;   // MMA -> epilogue: accumulators complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * kWStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * kTileX, y0 = blockIdx.y * kTileY;

  if (warp == 0 && lane == 0) {
    mbar_init(bar_brick, 1);
    // CLUSTER = 2: the two CTAs of a pair each fetch half of every weight tile and multicast it to both, so a stage is
    // free only when BOTH issuers have consumed it (two arrivals on each CTA's empty barrier).
    for (int s = 0; s < kWStages; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, CLUSTER); }
    mbar_init(bar_done, ISSUERS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation: kAcc accumulators x 64 columns (NOUT used of each), one warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();   // the peer's barriers are initialised before anything is multicast to it
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = CLUSTER > 1 ? cluster_ctarank() : 0u;

  if (warp == 0) {
    if (lane == 0) {
      // activations: one box [64 ch][kBrickX x][kBrickY y] per split term (out-of-range pixels are zero-filled by TMA)
      mbar_expect_tx(bar_brick, 2 * Cfg::kBrickBytes);
      tma_load_3d(a_hi, &map_ahi, bar_brick, 0, x0, y0);
      tma_load_3d(a_lo, &map_alo, bar_brick, 0, x0, y0);
      // weights: ring over the taps
      for (int t = 0; t < kTaps; ++t) {
        const int s = t % kWStages, round = t / kWStages;
        if (round > 0) mbar_wait(bar_empty + s, (round - 1) & 1);
        mbar_expect_tx(bar_full + s, Cfg::kWStageBytes);
        if (CLUSTER == 1) {
          tma_load_2d(w_st + s * Cfg::kWStageBytes, &map_whi, bar_full + s, 0, t * NOUT);
          tma_load_2d(w_st + s * Cfg::kWStageBytes + NOUT * 128, &map_wlo, bar_full + s, 0, t * NOUT);
        } else if (crank == 0) {   // rank 0 brings w_hi, rank 1 brings w_lo; each lands in both CTAs
          tma_load_2d_mc(w_st + s * Cfg::kWStageBytes, &map_whi, bar_full + s, 0, t * NOUT, (uint16_t)3);
        } else {
          tma_load_2d_mc(w_st + s * Cfg::kWStageBytes + NOUT * 128, &map_wlo, bar_full + s, 0, t * NOUT, (uint16_t)3);
        }
      }
    }
  } else if (warp == 1 || (ISSUERS == 2 && warp == 6)) {
    // MMA issuer(s). With ISSUERS == 2 the taps alternate between two issuing warps (own accumulators each), which
    // overlaps the per-instruction issue latency of the small N = 48 MMAs.
    if (lane == 0) {
      const int me = (warp == 1) ? 0 : 1;
      mbar_wait(bar_brick, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ahi = smem_u32(a_hi), alo = smem_u32(a_lo);
      constexpr int kMainPer = NMAIN / ISSUERS;            // main accumulators per issuer
      const uint32_t d_corr = tmem_base + (uint32_t)((NMAIN + me) * 64);
      int mine = 0;
      for (int t = me; t < kTaps; t += ISSUERS, ++mine) {
        const int s = t % kWStages, round = t / kWStages;
        mbar_wait(bar_full + s, round & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int ky = t / KS, kx = t % KS;
        const uint32_t shift = (uint32_t)(ky * Cfg::kBrickX + kx);           // in pixels = 128-byte rows
        const uint32_t bo = use_base_offset ? (shift & 7u) : 0u;
        const uint32_t whi = smem_u32(w_st + s * Cfg::kWStageBytes), wlo = whi + NOUT * 128;
        // The tensor core truncates when it adds into the fp32 accumulator, so a single accumulator would take
        // thousands of biased roundings in the 15x15 layer (measured 4e-5 relative). Spread them: the a_hi*w_hi
        // products go round-robin to NMAIN accumulators, both small correction terms to one more per issuer; the
        // epilogue sums them in fp32 round-to-nearest.
        const uint32_t d_main = tmem_base + (uint32_t)((me * kMainPer + (mine % kMainPer)) * 64);
#pragma unroll
        for (int j = 0; j < KSTEPS; ++j) {   // KSTEPS x UMMA_K(16) channels; pad channels beyond are never multiplied
          const uint64_t dah = make_desc(ahi + shift * 128 + j * 32, Cfg::kBrickX * 128, bo);
          const uint64_t dal = make_desc(alo + shift * 128 + j * 32, Cfg::kBrickX * 128, bo);
          const uint64_t dwh = make_desc(whi + j * 32, 1024, 0);
          const uint64_t dwl = make_desc(wlo + j * 32, 1024, 0);
          umma_f16(d_main, dah, dwh, Cfg::kIdesc, (mine >= kMainPer || j != 0));
          umma_f16(d_corr, dah, dwl, Cfg::kIdesc, (mine | j) != 0);
          umma_f16(d_corr, dal, dwh, Cfg::kIdesc, 1);
        }
        if (CLUSTER == 1) umma_commit(bar_empty + s);   // frees the weight stage once these MMAs have read it
        else umma_commit_mc(bar_empty + s, (uint16_t)3);
      }
      umma_commit(bar_done);   // bar_done counts ISSUERS arrivals
    }
  } else {
    // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 = output pixels m = lane index; m = yl*8 + xl
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int oy = y0 + (m >> 3), ox = x0 + (m & 7);
    float acc[NOUT];
#pragma unroll
    for (int n = 0; n < NOUT; ++n) acc[n] = 0.0f;
#pragma unroll
    for (int a = 0; a < kAcc; ++a) {
      if (a < NMAIN && a >= kTaps) continue;   // accumulator never written (fewer taps than accumulators)
#pragma unroll
      for (int c = 0; c < NOUT / 16; ++c) {
        uint32_t v[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 64 + c * 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[16 * c + i] += __uint_as_float(v[i]);
      }
    }
    if (oy < OH && ox < OW) {
      const size_t pix = (size_t)oy * OW + ox;
      if (SPLIT_OUT) {
        __half2* ph = reinterpret_cast<__half2*>(out_hi + pix * 64);
        __half2* pl = reinterpret_cast<__half2*>(out_lo + pix * 64);
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) {
          float f[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int n = 2 * n2 + e;
            float v = 0.0f;
            if (n < NOUT && n < cout) { v = acc[n < NOUT ? n : 0] * (1.0f / kWScale) + __ldg(bias + n); v = v > 0.0f ? v : 0.3f * v; }
            f[e] = v;
          }
          const __half h0 = __float2half_rn(f[0]), h1 = __float2half_rn(f[1]);
          ph[n2] = __halves2half2(h0, h1);
          pl[n2] = __halves2half2(__float2half_rn(f[0] - __half2float(h0)), __float2half_rn(f[1] - __half2float(h1)));
        }
      } else {
        float* op = out + pix * cout;
#pragma unroll
        for (int n = 0; n < NOUT; ++n) {
          if (n < cout) {
            float f = acc[n] * (1.0f / kWScale) + __ldg(bias + n);
            op[n] = f > 0.0f ? f : 0.3f * f;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();   // no CTA exits while its peer may still multicast into it / arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// 15x15 layer, two-phase variant. The single-phase kernel keeps both activation bricks (a_hi, a_lo: 184 KB) resident,
// which leaves room for only 3 weight stages: 36 KB in flight per SM against ~1.5 us of L2 latency = 24 GB/s per SM,
// i.e. the 2.7 MB weight stream takes ~110 us while the MMAs need ~26 us (measured: 113 us; TMA multicast across a CTA
// pair did not help -- the limit is in-flight bytes, not L2 bandwidth). Here only ONE brick is resident at a time:
//   phase 1: a_hi brick; per tap a_hi*w_hi -> main accumulators, a_hi*w_lo -> correction accumulator (12 KB/tap)
//   phase 2: the a_lo brick replaces it; per tap a_lo*w_hi -> correction accumulator (6 KB/tap, two taps per stage)
// and the freed 92 KB become 8 more weight stages (11 x 12 KB in flight).
constexpr int kW2Stages = 11;
struct Conv15Cfg {
  using B = ConvCfg<15, 48>;
  static constexpr int kSmem = B::kBrickBytes + kW2Stages * B::kWStageBytes + 1024 + 256;
};

// NI MMA-issuing warps share the stages round-robin (stage g belongs to issuer g % NI): a single thread issues one
// small (N = 48) tcgen05.mma about every 100 cycles regardless of the ring depth (measured: 113 us for 2025 MMAs with 1
// issuer, 78 us with 2), so several issuers are needed to approach the 26 us the tensor pipe itself needs.
template <int NI>
__global__ void __launch_bounds__(192 + 32 * (NI - 1), 1)
conv15_two_phase_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                        const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo,
                        const float* __restrict__ bias, float* __restrict__ out, int OH, int OW) {
  using Cfg = ConvCfg<15, 48>;
  static_assert(2 * NI <= 8, "two TMEM accumulators (main + correction) of 64 columns per issuer");
  constexpr int KS = 15, NOUT = 48, KSTEPS = 3, kTaps = 225, kAcc = 2 * NI;
  constexpr int kTmemCols = kAcc * 64 <= 128 ? 128 : (kAcc * 64 <= 256 ? 256 : 512);
  constexpr int kPairs = (kTaps + 1) / 2;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_br = smem;
  unsigned char* w_st = smem + Cfg::kBrickBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_st + kW2Stages * Cfg::kWStageBytes);
  uint64_t* bar_brick = bars;                   // TMA -> MMA: brick landed (phase 0: a_hi, phase 1: a_lo)
  uint64_t* bar_phase = bars + 1;               // MMA -> TMA: every issuer's phase-1 MMAs finished reading the a_hi brick
  uint64_t* bar_done = bars + 2;                // MMA -> epilogue (NI arrivals)
  uint64_t* bar_full = bars + 3;                // [kW2Stages]
  uint64_t* bar_empty = bars + 3 + kW2Stages;   // [kW2Stages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 + 2 * kW2Stages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * kTileX, y0 = blockIdx.y * kTileY;
  if (warp == 0 && lane == 0) {
    mbar_init(bar_brick, 1); mbar_init(bar_phase, NI); mbar_init(bar_done, NI);
    for (int s = 0; s < kW2Stages; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int issuer = (warp == 1) ? 0 : (warp >= 6 ? warp - 5 : -1);

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_brick, Cfg::kBrickBytes);
      tma_load_3d(a_br, &map_ahi, bar_brick, 0, x0, y0);
      int g = 0;
      for (int t = 0; t < kTaps; ++t, ++g) {                       // phase 1: w_hi + w_lo of one tap per stage
        const int s = g % kW2Stages, round = g / kW2Stages;
        if (round > 0) mbar_wait(bar_empty + s, (round - 1) & 1);
        mbar_expect_tx(bar_full + s, Cfg::kWStageBytes);
        tma_load_2d(w_st + s * Cfg::kWStageBytes, &map_whi, bar_full + s, 0, t * NOUT);
        tma_load_2d(w_st + s * Cfg::kWStageBytes + NOUT * 128, &map_wlo, bar_full + s, 0, t * NOUT);
      }
      mbar_wait(bar_phase, 0);                                     // a_hi brick no longer read
      mbar_expect_tx(bar_brick, Cfg::kBrickBytes);
      tma_load_3d(a_br, &map_alo, bar_brick, 0, x0, y0);
      for (int u = 0; u < kPairs; ++u, ++g) {                      // phase 2: w_hi of two taps per stage
        const int s = g % kW2Stages, round = g / kW2Stages;
        if (round > 0) mbar_wait(bar_empty + s, (round - 1) & 1);
        const int t0 = 2 * u, t1 = 2 * u + 1;
        mbar_expect_tx(bar_full + s, (t1 < kTaps ? 2 : 1) * NOUT * 128);
        tma_load_2d(w_st + s * Cfg::kWStageBytes, &map_whi, bar_full + s, 0, t0 * NOUT);
        if (t1 < kTaps) tma_load_2d(w_st + s * Cfg::kWStageBytes + NOUT * 128, &map_whi, bar_full + s, 0, t1 * NOUT);
      }
    }
  } else if (issuer >= 0) {
    if (lane == 0) {
      const uint32_t abr = smem_u32(a_br);
      const uint32_t d_main = tmem_base + (uint32_t)(issuer * 128), d_corr = d_main + 64u;
      mbar_wait(bar_brick, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      bool first = true;
      for (int g = issuer; g < kTaps; g += NI) {                   // phase 1: g == tap
        const int s = g % kW2Stages, round = g / kW2Stages, t = g;
        mbar_wait(bar_full + s, round & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t shift = (uint32_t)((t / KS) * Cfg::kBrickX + (t % KS));
        const uint32_t whi = smem_u32(w_st + s * Cfg::kWStageBytes), wlo = whi + NOUT * 128;
#pragma unroll
        for (int j = 0; j < KSTEPS; ++j) {
          const uint64_t da = make_desc(abr + shift * 128 + j * 32, Cfg::kBrickX * 128, 0);
          umma_f16(d_main, da, make_desc(whi + j * 32, 1024, 0), Cfg::kIdesc, !(first && j == 0));
          umma_f16(d_corr, da, make_desc(wlo + j * 32, 1024, 0), Cfg::kIdesc, !(first && j == 0));
        }
        first = false;
        umma_commit(bar_empty + s);
      }
      umma_commit(bar_phase);
      mbar_wait(bar_brick, 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // phase-2 stages continue the global stage count: g2 = kTaps + u; the issuer of a stage is g2 % NI
      int u0 = ((issuer - (kTaps % NI)) % NI + NI) % NI;
      for (int u = u0; u < kPairs; u += NI) {
        const int g = kTaps + u, s = g % kW2Stages, round = g / kW2Stages;
        mbar_wait(bar_full + s, round & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int t = 2 * u + e;
          if (t < kTaps) {
            const uint32_t shift = (uint32_t)((t / KS) * Cfg::kBrickX + (t % KS));
            const uint32_t whi = smem_u32(w_st + s * Cfg::kWStageBytes) + e * NOUT * 128;
#pragma unroll
            for (int j = 0; j < KSTEPS; ++j)
              umma_f16(d_corr, make_desc(abr + shift * 128 + j * 32, Cfg::kBrickX * 128, 0), make_desc(whi + j * 32, 1024, 0),
                       Cfg::kIdesc, 1);
          }
        }
        umma_commit(bar_empty + s);
      }
      umma_commit(bar_done);
    }
  } else {
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int oy = y0 + (m >> 3), ox = x0 + (m & 7);
    float acc[NOUT];
#pragma unroll
    for (int n = 0; n < NOUT; ++n) acc[n] = 0.0f;
#pragma unroll
    for (int a = 0; a < kAcc; ++a) {
#pragma unroll
      for (int c = 0; c < NOUT / 16; ++c) {
        uint32_t v[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 64 + c * 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[16 * c + i] += __uint_as_float(v[i]);
      }
    }
    if (oy < OH && ox < OW) {
      float* op = out + ((size_t)oy * OW + ox) * 48;
#pragma unroll
      for (int n = 0; n < NOUT; ++n) {
        const float f = acc[n] * (1.0f / kWScale) + __ldg(bias + n);
        op[n] = f > 0.0f ? f : 0.3f * f;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
}

// First layer (Cin = 1, no activation after its BN) on CUDA cores, reading the map layer directly and writing the
// fp16 hi/lo NHWC-64 input of the second layer. One thread per (pixel, group of 8 output channels): the eight threads of
// a pixel write its two 128-byte rows as sixteen 16-byte stores; channels 24..63 are zero.
__global__ void __launch_bounds__(256) conv1_split_kernel(const float* __restrict__ layer, int H, int W, int pitch,
                                                          const float* __restrict__ wf /*[9][1][24]*/, const float* __restrict__ bias,
                                                          __half* __restrict__ hi, __half* __restrict__ lo) {
  __shared__ float sw[9 * 24 + 24];
  for (int i = threadIdx.x; i < 9 * 24 + 24; i += blockDim.x) sw[i] = i < 216 ? wf[i] : bias[i - 216];
  __syncthreads();
  const int OH = H - 2, OW = W - 2;
  const size_t total = (size_t)OH * OW * 8;
  // consecutive pixels take consecutive image rows (oy): the map layer is contiguous along that axis
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int v4 = (int)(i & 7);
    const size_t p = i >> 3;
    const int oy = (int)(p % OH), ox = (int)(p / OH);
    const size_t pix = (size_t)oy * OW + ox;
    __half2 hh[4], ll[4];
    if (v4 < 3) {
      float in[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) in[t] = __ldg(layer + (size_t)(ox + t % 3) * pitch + (H - 1 - (oy + t / 3)));   // E[r][c]
#pragma unroll
      for (int e2 = 0; e2 < 4; ++e2) {
        float f[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int n = 8 * v4 + 2 * e2 + e;
          float a = sw[216 + n];
#pragma unroll
          for (int t = 0; t < 9; ++t) a = fmaf(in[t], sw[t * 24 + n], a);
          f[e] = a;
        }
        const __half h0 = __float2half_rn(f[0]), h1 = __float2half_rn(f[1]);
        hh[e2] = __halves2half2(h0, h1);
        ll[e2] = __halves2half2(__float2half_rn(f[0] - __half2float(h0)), __float2half_rn(f[1] - __half2float(h1)));
      }
    } else {
#pragma unroll
      for (int e2 = 0; e2 < 4; ++e2) { hh[e2] = __float2half2_rn(0.0f); ll[e2] = __float2half2_rn(0.0f); }
    }
    reinterpret_cast<uint4*>(hi + pix * 64)[v4] = *reinterpret_cast<uint4*>(hh);
    reinterpret_cast<uint4*>(lo + pix * 64)[v4] = *reinterpret_cast<uint4*>(ll);
  }
}

// max pooling (K x K, stride S) of an fp32 NHWC [H][W][C] activation (C a multiple of 8) into fp16 hi/lo NHWC-64.
// One thread per (pixel, 8 channels): two 16-byte loads per tap, one 16-byte store per output array.
__global__ void maxpool_split_kernel(const float* __restrict__ in, int H, int W, int C, int K, int S, __half* __restrict__ hi,
                                     __half* __restrict__ lo, int OH, int OW) {
  const size_t total = (size_t)OH * OW * 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int v8 = (int)(i & 7);
    const size_t p = i >> 3;
    const int ox = (int)(p % OW), oy = (int)(p / OW);
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = 0.0f;
    if (8 * v8 < C) {
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
      for (int dy = 0; dy < K; ++dy)
        for (int dx = 0; dx < K; ++dx) {
          const float4* q = reinterpret_cast<const float4*>(in + ((size_t)(oy * S + dy) * W + ox * S + dx) * C + 8 * v8);
          const float4 a = __ldg(q), b = __ldg(q + 1);
          m[0] = fmaxf(m[0], a.x); m[1] = fmaxf(m[1], a.y); m[2] = fmaxf(m[2], a.z); m[3] = fmaxf(m[3], a.w);
          m[4] = fmaxf(m[4], b.x); m[5] = fmaxf(m[5], b.y); m[6] = fmaxf(m[6], b.z); m[7] = fmaxf(m[7], b.w);
        }
    }
    __half2 hh[4], ll[4];
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      const __half h0 = __float2half_rn(m[2 * e2]), h1 = __float2half_rn(m[2 * e2 + 1]);
      hh[e2] = __halves2half2(h0, h1);
      ll[e2] = __halves2half2(__float2half_rn(m[2 * e2] - __half2float(h0)), __float2half_rn(m[2 * e2 + 1] - __half2float(h1)));
    }
    reinterpret_cast<uint4*>(hi + p * 64)[v8] = *reinterpret_cast<uint4*>(hh);
    reinterpret_cast<uint4*>(lo + p * 64)[v8] = *reinterpret_cast<uint4*>(ll);
  }
}

// Tensor-core weights of one layer: [taps][NOUT][64] fp16 hi/lo (K-major rows of 128 B; pad rows / channels zero),
// BN scale and kWScale folded; bias = beta - mean*scale.
__global__ void fold_tc_kernel(const float* __restrict__ w /*[cout][cin][taps]*/, const float* __restrict__ bn, int cout, int cin,
                               int taps, int nout, __half* __restrict__ whi, __half* __restrict__ wlo, float* __restrict__ bias) {
  const int total = taps * nout * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i & 63, r = i >> 6, n = r % nout, tap = r / nout;
    float v = 0.0f;
    if (c < cin && n < cout) {
      const float s = bn[n] / sqrtf(bn[3 * cout + n] + kBnEps);
      v = w[((size_t)n * cin + c) * taps + tap] * s * kWScale;   // power-of-two scale keeps w_lo out of fp16 subnormals
    }
    const __half h = __float2half_rn(v);
    whi[i] = h;
    wlo[i] = __float2half_rn(v - __half2float(h));
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < cout; n += gridDim.x * blockDim.x) {
    const float s = bn[n] / sqrtf(bn[3 * cout + n] + kBnEps);
    bias[n] = bn[cout + n] - bn[2 * cout + n] * s;
  }
}

// Reference implementation of the 15x15 layer on CUDA cores (fp32), used by the self-check entry point only.
__global__ void conv15_reference_kernel(const float* __restrict__ in /*[H][W][48]*/, int H, int W,
                                        const float* __restrict__ wf /*[225][48][48]*/, const float* __restrict__ bias,
                                        float* __restrict__ out) {
  const int OH = H - 14, OW = W - 14;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= OH * OW * 48) return;
  const int n = i % 48, p = i / 48, ox = p % OW, oy = p / OW;
  float acc = 0.0f;
  for (int ky = 0; ky < 15; ++ky)
    for (int kx = 0; kx < 15; ++kx) {
      const float* ip = in + ((size_t)(oy + ky) * W + ox + kx) * 48;
      const float* wp = wf + (size_t)(ky * 15 + kx) * 48 * 48 + n;
      for (int c = 0; c < 48; ++c) acc = fmaf(ip[c], wp[c * 48], acc);
    }
  const float f = acc + bias[n];
  out[i] = f > 0.0f ? f : 0.3f * f;
}

// ------------------------------------------------------------------------------------------------
// Query head: CostQuery.__call__ + FCpart, one thread per query, folded weights in shared memory.
// ------------------------------------------------------------------------------------------------
// folded head weights: tar0 [10][16]+b[16], out0 [64][48]+b[48], o11 [48][24]+b, o12 [48][24]+b, o13 [48][36]+b,
// o21 [24]+b[1], o22 [24]+b[1], o23 [36]+b[1]
constexpr int kHeadFloats = 10 * 16 + 16 + 64 * 48 + 48 + 48 * 24 + 24 + 48 * 24 + 24 + 48 * 36 + 36 + 24 + 1 + 24 + 1 + 36 + 1;

__global__ void __launch_bounds__(128) head_kernel(const float* __restrict__ feats /*[Hf][Wf][48]*/, int Hf, int Wf,
                                                    const float* __restrict__ hw, const float* __restrict__ edges, size_t n,
                                                    float* __restrict__ cost3, double res, double Lx, double Ly, double cx,
                                                    double cy) {
  extern __shared__ float sw[];
  for (int i = threadIdx.x; i < kHeadFloats; i += blockDim.x) sw[i] = hw[i];
  __syncthreads();
  const size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (q >= n) return;
  const float* e = edges + 6 * q;
  // float64 arithmetic on the float32 request values, like the numpy/torch float64 path of the server
  double tx = (double)e[0] - cx, ty = (double)e[1] - cy, tyaw = (double)e[2];
  const double sx = (double)e[3] - cx, sy = (double)e[4] - cy, syaw = (double)e[5];
  tx -= sx; ty -= sy; tyaw -= syaw;
  const double feat_res = res * 2.0;
  const int row_bias = (int)((Lx / res - 48.0) / 2.0 * 0.5), col_bias = (int)((Ly / res - 48.0) / 2.0 * 0.5);
  double rr = sx / feat_res + row_bias, cc = sy / feat_res + col_bias;
  rr = fmin(fmax(rr, 1.0), (double)(Hf - 2));
  cc = fmin(fmax(cc, 1.0), (double)(Wf - 2));
  const int row = (int)rr, col = (int)cc;   // .long() truncation
  float x[64];
  {
    const float4* fp = reinterpret_cast<const float4*>(feats + ((size_t)row * Wf + col) * 48);
#pragma unroll
    for (int i = 0; i < 12; ++i) { const float4 v = __ldg(fp + i); x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w; }
  }
  const float dx = (float)tx, dy = (float)ty;
  float ang = (float)tyaw;
  const float PI = 3.14159265358979323846f;
  if (ang > PI) ang -= 2.0f * PI;
  if (ang < -PI) ang += 2.0f * PI;
  const float sya = (float)syaw;
  const float info[10] = {dx, dy, sqrtf(dx * dx + dy * dy), atan2f(dy, dx), ang, cosf(ang), sinf(ang), sya, cosf(sya), sinf(sya)};
  const float* w = sw;
  // tar0: 10 -> 16 (BN folded, no activation)
#pragma unroll
  for (int o = 0; o < 16; ++o) {
    float a = w[160 + o];
#pragma unroll
    for (int i = 0; i < 10; ++i) a = fmaf(info[i], w[i * 16 + o], a);
    x[48 + o] = a;
  }
  w += 176;
  float h[48];
#pragma unroll
  for (int o = 0; o < 48; ++o) h[o] = w[64 * 48 + o];
  for (int i = 0; i < 64; ++i) {
    const float v = x[i];
#pragma unroll
    for (int o = 0; o < 48; ++o) h[o] = fmaf(v, w[i * 48 + o], h[o]);
  }
#pragma unroll
  for (int o = 0; o < 48; ++o) h[o] = h[o] > 0.0f ? h[o] : 0.3f * h[o];
  w += 64 * 48 + 48;
  float outv[3];
  const int widths[3] = {24, 24, 36};
  const float* w2 = w + (48 * 24 + 24) * 2 + 48 * 36 + 36;
  for (int b = 0; b < 3; ++b) {
    const int nb = widths[b];
    float acc2 = 0.0f;
    for (int o = 0; o < nb; ++o) {
      float a = w[48 * nb + o];
      for (int i = 0; i < 48; ++i) a = fmaf(h[i], w[i * nb + o], a);
      a = a > 0.0f ? a : 0.3f * a;
      acc2 = fmaf(a, w2[o], acc2);
    }
    acc2 += w2[nb];
    outv[b] = acc2;
    w += 48 * nb + nb;
    w2 += nb + 1;
  }
  cost3[3 * q + 0] = fmaxf(outv[0], 0.0f);                        // power  (ReLU)
  cost3[3 * q + 1] = fmaxf(outv[1], 0.0f);                        // time   (ReLU)
  cost3[3 * q + 2] = 1.0f - 1.0f / (1.0f + expf(-outv[2]));       // 1 - sigmoid
}

// Fold the head's 1x1 convs (+BN) into [Cin][Cout] matrices + bias, in head_kernel's order.
__global__ void fold_head_kernel(const float* __restrict__ blob, const size_t* __restrict__ offs, float* __restrict__ hw) {
  // offs[l] = offset of layer l (6..13) in the blob; single block
  const int cin[8] = {10, 64, 48, 48, 48, 24, 24, 36}, cout[8] = {16, 48, 24, 24, 36, 1, 1, 1};
  int o = 0;
  for (int l = 0; l < 8; ++l) {
    const float* w = blob + offs[l];
    const float* bn = w + (size_t)cout[l] * cin[l];
    const bool has_bn = l < 5;
    for (int i = threadIdx.x; i < cin[l] * cout[l]; i += blockDim.x) {
      const int co = i % cout[l], ci = i / cout[l];
      const float s = has_bn ? bn[co] / sqrtf(bn[3 * cout[l] + co] + kBnEps) : 1.0f;
      hw[o + i] = w[(size_t)co * cin[l] + ci] * s;
    }
    for (int co = threadIdx.x; co < cout[l]; co += blockDim.x) {
      float b;
      if (has_bn) { const float s = bn[co] / sqrtf(bn[3 * cout[l] + co] + kBnEps); b = bn[cout[l] + co] - bn[2 * cout[l] + co] * s; }
      else b = bn[co];   // conv bias
      hw[o + cin[l] * cout[l] + co] = b;
    }
    o += cin[l] * cout[l] + cout[l];
  }
}

// ------------------------------------------------------------------------------------------------
// host state
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// One tensor-core layer: weights (hi/lo), bias, and its launch geometry.
struct TcLayer { __half *whi = nullptr, *wlo = nullptr; int nout = 0; };

struct State {
  int device = 0, sm_count = 0;
  bool has_weights = false, has_features = false;
  float* d_blob = nullptr;
  size_t layer_off[14];
  float* d_wf[5] = {};     // folded fp32 3x3 weights [9][cin][cout] (layer 0 always; 1..4 for the CUDA-core check path)
  float* d_bias[6] = {};   // folded biases (layers 0..5)
  float* d_wf6 = nullptr;  // folded fp32 15x15 weights [225][48][48] (CUDA-core check path only)
  TcLayer tc[6];           // tensor-core weights of layers 1..5 (index = layer)
  float* d_head = nullptr;
  size_t* d_offs = nullptr;
  // activations (sized for the current map); h* / l* = fp16 hi / lo NHWC-64, f* = fp32 NHWC
  int rows = 0, cols = 0;
  __half *h1 = nullptr, *l1 = nullptr, *hp2 = nullptr, *lp2 = nullptr, *h3 = nullptr, *l3 = nullptr, *hp4 = nullptr,
         *lp4 = nullptr, *h5 = nullptr, *l5 = nullptr;
  float *f1 = nullptr, *f2 = nullptr, *fp2 = nullptr, *f3 = nullptr, *f4 = nullptr, *fp4 = nullptr, *f5 = nullptr, *feat = nullptr;
  int Hf = 0, Wf = 0;
  double res = 0, Lx = 0, Ly = 0, cx = 0, cy = 0;
  EncodeTiledFn encode = nullptr;
  CUtensorMap maps[6][4];      // per tensor-core layer: activation hi/lo, weight hi/lo (rebuilt when buffers change)
  bool maps_valid = false;
  bool attrs_set = false;
  int use_base_offset = 0;
  int conv15_mode = 0;         // 15x15 layer: 0 two-phase (default), 1 single phase, 2 single phase + CTA-pair multicast
  float last_ms[3] = {0, 0, 0};
  cudaEvent_t ev[4] = {};
};

State* create(int device, int sm_count) {
  State* s = new State();
  s->device = device;
  s->sm_count = sm_count;
  return s;
}

static void free_acts(State* s) {
  void* ptrs[] = {s->h1, s->l1, s->hp2, s->lp2, s->h3, s->l3, s->hp4, s->lp4, s->h5, s->l5,
                  s->f1, s->f2, s->fp2, s->f3, s->f4, s->fp4, s->f5, s->feat};
  for (void* p : ptrs) cudaFree(p);
  s->h1 = s->l1 = s->hp2 = s->lp2 = s->h3 = s->l3 = s->hp4 = s->lp4 = s->h5 = s->l5 = nullptr;
  s->f1 = s->f2 = s->fp2 = s->f3 = s->f4 = s->fp4 = s->f5 = s->feat = nullptr;
}

void destroy(State* s) {
  if (!s) return;
  free_acts(s);
  cudaFree(s->d_blob);
  for (auto p : s->d_wf) cudaFree(p);
  for (auto p : s->d_bias) cudaFree(p);
  for (auto& t : s->tc) { cudaFree(t.whi); cudaFree(t.wlo); }
  cudaFree(s->d_wf6); cudaFree(s->d_head); cudaFree(s->d_offs);
  for (auto e : s->ev) if (e) cudaEventDestroy(e);
  delete s;
}

void set_base_offset_mode(State* s, int on) { s->use_base_offset = on ? 1 : 0; }
void set_conv15_mode(State* s, int mode) { if (s->conv15_mode != mode) { s->conv15_mode = mode; s->attrs_set = false; } }
bool has_features(const State* s) { return s->has_features; }
bool has_weights(const State* s) { return s->has_weights; }
void last_times(const State* s, float* ms3) { ms3[0] = s->last_ms[0]; ms3[1] = s->last_ms[1]; ms3[2] = s->last_ms[2]; }

static const int kTcNout[6] = {0, 32, 48, 48, 48, 48};   // UMMA N per layer (Cout 24 padded to 32)

int set_weights(State* s, const float* blob, size_t n, cudaStream_t st, std::string& err) {
  if (n != blob_floats()) { err = "weight blob has the wrong number of floats"; return -1; }
  CNN_TRY(cudaSetDevice(s->device));
  if (!s->d_blob) {
    CNN_TRY(cudaMalloc(&s->d_blob, n * sizeof(float)));
    for (int l = 0; l < 5; ++l) CNN_TRY(cudaMalloc(&s->d_wf[l], (size_t)9 * kLayers[l].cin * kLayers[l].cout * sizeof(float)));
    for (int l = 0; l < 6; ++l) CNN_TRY(cudaMalloc(&s->d_bias[l], kLayers[l].cout * sizeof(float)));
    CNN_TRY(cudaMalloc(&s->d_wf6, (size_t)225 * 48 * 48 * sizeof(float)));
    for (int l = 1; l < 6; ++l) {
      const size_t cnt = (size_t)kLayers[l].k * kLayers[l].k * kTcNout[l] * 64;
      s->tc[l].nout = kTcNout[l];
      CNN_TRY(cudaMalloc(&s->tc[l].whi, cnt * sizeof(__half)));
      CNN_TRY(cudaMalloc(&s->tc[l].wlo, cnt * sizeof(__half)));
    }
    CNN_TRY(cudaMalloc(&s->d_head, kHeadFloats * sizeof(float)));
    CNN_TRY(cudaMalloc(&s->d_offs, 8 * sizeof(size_t)));
  }
  size_t off = 0;
  for (int l = 0; l < 14; ++l) {
    s->layer_off[l] = off;
    off += (size_t)kLayers[l].cout * kLayers[l].cin * kLayers[l].k * kLayers[l].k + (kLayers[l].bn ? 4 * kLayers[l].cout : kLayers[l].cout);
  }
  CNN_TRY(cudaMemcpyAsync(s->d_blob, blob, n * sizeof(float), cudaMemcpyHostToDevice, st));
  CNN_TRY(cudaMemcpyAsync(s->d_offs, s->layer_off + 6, 8 * sizeof(size_t), cudaMemcpyHostToDevice, st));
  for (int l = 0; l < 6; ++l) {
    const int kk = kLayers[l].k * kLayers[l].k;
    const float* w = s->d_blob + s->layer_off[l];
    const float* bn = w + (size_t)kLayers[l].cout * kLayers[l].cin * kk;
    // fp32 folded weights (first layer + the CUDA-core check path) and biases
    fold_conv_kernel<<<256, 256, 0, st>>>(w, bn, kLayers[l].cout, kLayers[l].cin, kk, l < 5 ? s->d_wf[l] : s->d_wf6, s->d_bias[l]);
    if (l >= 1)
      fold_tc_kernel<<<256, 256, 0, st>>>(w, bn, kLayers[l].cout, kLayers[l].cin, kk, kTcNout[l], s->tc[l].whi, s->tc[l].wlo,
                                          s->d_bias[l]);
  }
  fold_head_kernel<<<1, 256, 0, st>>>(s->d_blob, s->d_offs, s->d_head);
  CNN_TRY(cudaGetLastError());
  CNN_TRY(cudaStreamSynchronize(st));
  s->has_weights = true;
  s->has_features = false;
  return 0;
}

static int get_encoder(State* s, std::string& err) {
  if (s->encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    err = "cuTensorMapEncodeTiled not available from the driver";
    return -3;
  }
  s->encode = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}

// maps[0..1]: activations hi/lo [H][W][64] fp16, box (64, brick_x, brick_y); maps[2..3]: weights [taps*nout][64], box (64, nout)
static int encode_layer_maps(State* s, CUtensorMap* maps, __half* ahi, __half* alo, int H, int W, int brick_x, int brick_y,
                             __half* whi, __half* wlo, int taps, int nout, std::string& err) {
  int rc = get_encoder(s, err);
  if (rc) return rc;
  const cuuint64_t adim[3] = {64, (cuuint64_t)W, (cuuint64_t)H};
  const cuuint64_t astr[2] = {128, (cuuint64_t)W * 128};
  const cuuint32_t abox[3] = {64, (cuuint32_t)brick_x, (cuuint32_t)brick_y};
  const cuuint32_t one3[3] = {1, 1, 1};
  const cuuint64_t wdim[2] = {64, (cuuint64_t)taps * nout};
  const cuuint64_t wstr[1] = {128};
  const cuuint32_t wbox[2] = {64, (cuuint32_t)nout};
  void* ptrs[4] = {ahi, alo, whi, wlo};
  for (int i = 0; i < 4; ++i) {
    CUresult r;
    if (i < 2)
      r = s->encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, ptrs[i], adim, astr, abox, one3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    else
      r = s->encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptrs[i], wdim, wstr, wbox, one3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return -3; }
  }
  return 0;
}

template <int KS, int KSTEPS, int NOUT, int NMAIN, bool SPLIT, int CLUSTER, int ISSUERS = 1>
static int launch_tc(State* s, int layer, __half* ahi, __half* alo, int H, int W, float* out, __half* ohi, __half* olo,
                     cudaStream_t st, std::string& err) {
  using Cfg = ConvCfg<KS, NOUT>;
  CUtensorMap* maps = s->maps[layer];
  if (!s->maps_valid) {
    int rc = encode_layer_maps(s, maps, ahi, alo, H, W, Cfg::kBrickX, Cfg::kBrickY, s->tc[layer].whi, s->tc[layer].wlo, KS * KS, NOUT, err);
    if (rc) return rc;
  }
  auto kern = conv_tc_kernel<KS, KSTEPS, NOUT, NMAIN, SPLIT, CLUSTER, ISSUERS>;
  if (!s->attrs_set) CNN_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  const int OH = H - KS + 1, OW = W - KS + 1;
  const int gx = (OW + kTileX - 1) / kTileX, gy = (OH + kTileY - 1) / kTileY;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(((gx + CLUSTER - 1) / CLUSTER) * CLUSTER, gy);   // padded CTAs compute an out-of-range tile (stores masked)
  cfg.blockDim = dim3(192 + 32 * (ISSUERS - 1));
  cfg.dynamicSmemBytes = Cfg::kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CLUSTER > 1 ? 1 : 0;
  const int cout = kLayers[layer].cout, ubo = s->use_base_offset;
  const float* bias = s->d_bias[layer];
  CNN_TRY(cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], bias, out, ohi, olo, OH, OW, cout, ubo));
  return 0;
}

// CostPredictor.updateFeatures (predictor.py:28-36): the CNN trunk over the `elevation` layer currently uploaded.
// use_cuda_core_path: fp32 CUDA-core direct convolutions for every layer (the in-library cross-check of the tensor path).
int update_features(State* s, const float* d_layer, int rows, int cols, int pitch, double res, double cx, double cy,
                    cudaStream_t st, int use_cuda_core_path, std::string& err) {
  if (!s->has_weights) { err = "motion-cost weights not set"; return -5; }
  if (rows < 64 || cols < 64) { err = "map too small for the motion-cost network (needs >= 64 x 64 cells)"; return -1; }
  CNN_TRY(cudaSetDevice(s->device));
  const int H0 = rows, W0 = cols;
  const int H1 = H0 - 2, W1 = W0 - 2, H2 = H1 - 2, W2 = W1 - 2, HP2 = H2 / 2, WP2 = W2 / 2;
  const int H3 = HP2 - 2, W3 = WP2 - 2, H4 = H3 - 2, W4 = W3 - 2, HP4 = H4 - 2, WP4 = W4 - 2;
  const int H5 = HP4 - 2, W5 = WP4 - 2, H6 = H5 - 14, W6 = W5 - 14;
  if (rows != s->rows || cols != s->cols) {
    free_acts(s);
    auto hl = [&](__half** h, __half** l, int hh, int ww) -> cudaError_t {
      cudaError_t e = cudaMalloc(h, (size_t)hh * ww * 64 * 2);
      return e != cudaSuccess ? e : cudaMalloc(l, (size_t)hh * ww * 64 * 2);
    };
    CNN_TRY(hl(&s->h1, &s->l1, H1, W1));
    CNN_TRY(hl(&s->hp2, &s->lp2, HP2, WP2));
    CNN_TRY(hl(&s->h3, &s->l3, H3, W3));
    CNN_TRY(hl(&s->hp4, &s->lp4, HP4, WP4));
    CNN_TRY(hl(&s->h5, &s->l5, H5, W5));
    CNN_TRY(cudaMalloc(&s->f1, (size_t)H1 * W1 * 24 * 4));
    CNN_TRY(cudaMalloc(&s->f2, (size_t)H2 * W2 * 24 * 4));
    CNN_TRY(cudaMalloc(&s->fp2, (size_t)HP2 * WP2 * 24 * 4));
    CNN_TRY(cudaMalloc(&s->f3, (size_t)H3 * W3 * 48 * 4));
    CNN_TRY(cudaMalloc(&s->f4, (size_t)H4 * W4 * 48 * 4));
    CNN_TRY(cudaMalloc(&s->fp4, (size_t)HP4 * WP4 * 48 * 4));
    CNN_TRY(cudaMalloc(&s->f5, (size_t)H5 * W5 * 48 * 4));
    CNN_TRY(cudaMalloc(&s->feat, (size_t)H6 * W6 * 48 * 4));
    s->rows = rows; s->cols = cols;
    s->maps_valid = false;
  }
  if (!s->ev[0]) for (auto& e : s->ev) CNN_TRY(cudaEventCreate(&e));
  auto grid2 = [](int oh, int ow) { return dim3((ow + 15) / 16, (oh + 15) / 16); };
  const int g1 = s->sm_count * 8;
  CNN_TRY(cudaEventRecord(s->ev[0], st));
  if (use_cuda_core_path) {
    conv3x3_kernel<1, 24, 1, false, true><<<grid2(H1, W1), 256, 0, st>>>(d_layer, H0, W0, pitch, s->d_wf[0], s->d_bias[0], s->f1);
    conv3x3_kernel<24, 24, 8, true, false><<<grid2(H2, W2), 256, 0, st>>>(s->f1, H1, W1, 0, s->d_wf[1], s->d_bias[1], s->f2);
    maxpool_kernel<<<g1, 256, 0, st>>>(s->f2, H2, W2, 24, 2, 2, s->fp2, HP2, WP2);
    conv3x3_kernel<24, 48, 8, true, false><<<grid2(H3, W3), 256, 0, st>>>(s->fp2, HP2, WP2, 0, s->d_wf[2], s->d_bias[2], s->f3);
    conv3x3_kernel<48, 48, 8, true, false><<<grid2(H4, W4), 256, 0, st>>>(s->f3, H3, W3, 0, s->d_wf[3], s->d_bias[3], s->f4);
    maxpool_kernel<<<g1, 256, 0, st>>>(s->f4, H4, W4, 48, 3, 1, s->fp4, HP4, WP4);
    conv3x3_kernel<48, 48, 8, true, false><<<grid2(H5, W5), 256, 0, st>>>(s->fp4, HP4, WP4, 0, s->d_wf[4], s->d_bias[4], s->f5);
    CNN_TRY(cudaGetLastError());
    CNN_TRY(cudaEventRecord(s->ev[1], st));
    CNN_TRY(cudaEventRecord(s->ev[2], st));
    conv15_reference_kernel<<<(H6 * W6 * 48 + 255) / 256, 256, 0, st>>>(s->f5, H5, W5, s->d_wf6, s->d_bias[5], s->feat);
    CNN_TRY(cudaGetLastError());
  } else {
    int rc;
    conv1_split_kernel<<<g1, 256, 0, st>>>(d_layer, H0, W0, pitch, s->d_wf[0], s->d_bias[0], s->h1, s->l1);
    if ((rc = launch_tc<3, 2, 32, 1, false, 1>(s, 1, s->h1, s->l1, H1, W1, s->f2, nullptr, nullptr, st, err))) return rc;
    maxpool_split_kernel<<<g1, 256, 0, st>>>(s->f2, H2, W2, 24, 2, 2, s->hp2, s->lp2, HP2, WP2);
    if ((rc = launch_tc<3, 2, 48, 1, true, 1>(s, 2, s->hp2, s->lp2, HP2, WP2, nullptr, s->h3, s->l3, st, err))) return rc;
    if ((rc = launch_tc<3, 3, 48, 1, false, 1>(s, 3, s->h3, s->l3, H3, W3, s->f4, nullptr, nullptr, st, err))) return rc;
    maxpool_split_kernel<<<g1, 256, 0, st>>>(s->f4, H4, W4, 48, 3, 1, s->hp4, s->lp4, HP4, WP4);
    if ((rc = launch_tc<3, 3, 48, 1, true, 1>(s, 4, s->hp4, s->lp4, HP4, WP4, nullptr, s->h5, s->l5, st, err))) return rc;
    CNN_TRY(cudaGetLastError());
    CNN_TRY(cudaEventRecord(s->ev[1], st));
    CNN_TRY(cudaEventRecord(s->ev[2], st));
    if (s->conv15_mode == 0) {          // two-phase (default)
      using Cfg = ConvCfg<15, 48>;
      CUtensorMap* maps = s->maps[5];
      if (!s->maps_valid) {
        rc = encode_layer_maps(s, maps, s->h5, s->l5, H5, W5, Cfg::kBrickX, Cfg::kBrickY, s->tc[5].whi, s->tc[5].wlo, 225, 48, err);
        if (rc) return rc;
      }
      constexpr int kIssuers = 4;
      if (!s->attrs_set)
        CNN_TRY(cudaFuncSetAttribute(conv15_two_phase_kernel<kIssuers>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv15Cfg::kSmem));
      dim3 grid((W6 + kTileX - 1) / kTileX, (H6 + kTileY - 1) / kTileY);
      conv15_two_phase_kernel<kIssuers><<<grid, 192 + 32 * (kIssuers - 1), Conv15Cfg::kSmem, st>>>(maps[0], maps[1], maps[2], maps[3],
                                                                                                   s->d_bias[5], s->feat, H6, W6);
      CNN_TRY(cudaGetLastError());
    } else if (s->conv15_mode == 2) {   // single phase, CTA pairs with weight multicast
      rc = launch_tc<15, 3, 48, 7, false, 2>(s, 5, s->h5, s->l5, H5, W5, s->feat, nullptr, nullptr, st, err);
      if (rc) return rc;
    } else if (s->conv15_mode == 3) {   // single phase, two MMA-issuing warps
      rc = launch_tc<15, 3, 48, 6, false, 1, 2>(s, 5, s->h5, s->l5, H5, W5, s->feat, nullptr, nullptr, st, err);
      if (rc) return rc;
    } else {                            // single phase, one CTA per tile
      rc = launch_tc<15, 3, 48, 7, false, 1>(s, 5, s->h5, s->l5, H5, W5, s->feat, nullptr, nullptr, st, err);
      if (rc) return rc;
    }
    s->maps_valid = true;
    s->attrs_set = true;
  }
  CNN_TRY(cudaEventRecord(s->ev[3], st));
  CNN_TRY(cudaStreamSynchronize(st));
  cudaEventElapsedTime(&s->last_ms[0], s->ev[0], s->ev[1]);   // layers 1..5
  cudaEventElapsedTime(&s->last_ms[1], s->ev[2], s->ev[3]);   // 15x15 layer
  cudaEventElapsedTime(&s->last_ms[2], s->ev[0], s->ev[3]);   // whole trunk
  s->Hf = H6; s->Wf = W6;
  s->res = res; s->Lx = rows * res; s->Ly = cols * res; s->cx = cx; s->cy = cy;
  s->has_features = true;
  return 0;
}

int motion_cost(State* s, const float* d_edges, size_t n, float* d_cost3, cudaStream_t st, std::string& err) {
  if (!s->has_weights) { err = "motion-cost weights not set"; return -5; }
  if (!s->has_features) { err = "features not computed (call artp_update_features after artp_set_map)"; return -5; }
  if (n == 0) return 0;
  CNN_TRY(cudaSetDevice(s->device));
  // One thread per query. Small batches (config 4: 4096 queries) use one-warp CTAs so that the batch spreads over the
  // SMs (128 CTAs instead of 32: 39 -> ~10 us); big batches amortise the 29 KB weight load over 128 queries per CTA.
  const int bt = n <= (size_t)s->sm_count * 128 ? 32 : 128;
  head_kernel<<<(unsigned)((n + bt - 1) / bt), bt, kHeadFloats * sizeof(float), st>>>(s->feat, s->Hf, s->Wf, s->d_head, d_edges, n,
                                                                                       d_cost3, s->res, s->Lx, s->Ly, s->cx, s->cy);
  CNN_TRY(cudaGetLastError());
  return 0;
}

int copy_features(State* s, float* host_out, size_t n_floats, std::string& err) {
  if (!s->has_features) { err = "features not computed"; return -5; }
  if (n_floats != (size_t)s->Hf * s->Wf * 48) { err = "feature buffer size mismatch"; return -1; }
  CNN_TRY(cudaMemcpy(host_out, s->feat, n_floats * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

void feature_shape(const State* s, int* hf, int* wf) { *hf = s->Hf; *wf = s->Wf; }

}  // namespace artp_cnn
