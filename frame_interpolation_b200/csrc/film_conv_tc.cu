// tcgen05 implicit-GEMM convolution for sm_100a (the engine's hot kernel).
//
// One CTA computes a 128-pixel (tile_h x tile_w) x BN-channel output tile of one Conv2D
// call site (film_conv.h).  GEMM view: D[128 x BN] += A[128 x 64] * W[BN x 64]^T per K block,
// K blocks = (source, 64-channel chunk, tap).
//
//   warp 0    : TMA producer.  Per K block: 4-D tiled TMA loads of the hi and lo planes of the
//               activation box (64 ch, tile_w, tile_h, 1) at the tap-shifted coordinate --
//               out-of-bounds rows/cols are zero-filled by TMA, which IS the SAME padding of
//               tf.keras Conv2D -- plus 2-D loads of the W_hi / W_lo [BN x 64] blocks.
//               Everything lands in SWIZZLE_128B K-major layout (one pixel = one 128 B row).
//   warp 1    : allocates TMEM (BN fp32 columns) and issues tcgen05.mma.kind::f16 (M=128,
//               N=BN, K=16): per K block 4 k-steps x 3 passes  A_hi*W_hi + A_hi*W_lo + A_lo*W_hi
//               (split-precision product, fp32 accumulate in TMEM).  tcgen05.commit releases
//               smem stages back to the producer and finally signals the epilogue.
//   warps 2-5 : epilogue.  tcgen05.ld (32 lanes x 32 columns per warp), + bias, LeakyReLU,
//               re-split to hi/lo 16-bit planes, 128-bit stores into the destination channel
//               slice (which is how channel concats and the NN-upsample parity scatter are
//               realised without extra passes).
//
// mbarrier pipeline: full[s] (TMA -> MMA, tx-count), empty[s] (MMA -> TMA, via tcgen05.commit),
// tmem_full (MMA -> epilogue).  A watchdog turns a stuck barrier into a trap instead of a hang.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <stdint.h>

#include "film_conv.h"
#include "film_tc_ptx.cuh"

namespace film {

namespace {
using namespace tc;

constexpr int kNumThreads = 192;
template <int BN, int KC>
struct TcCfg {
  static constexpr int kABytes = kTileM * KC * 2;  // 16 KiB (KC = 64) or 8 KiB (KC = 32) per plane
  static constexpr int kWBytes = BN * KC * 2;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kWBytes;
  // Small-N tiles have short K loops and are bound by per-tile serialisation (prologue ->
  // mainloop -> epilogue), not by pipeline depth: give them 2 stages and 2 CTAs per SM so one
  // CTA's epilogue overlaps the other's mainloop (ncu, profiles/r1_ncu_conv.md).
  static constexpr int kStages = (BN == 256) ? 2 : (BN == 128) ? 3 : 2;
  static constexpr int kMinBlocks = (BN <= 64) ? 2 : 1;
  // fused-N product for BN <= 128 (see film_conv3x3_tc.cu): accumulator = 2*BN columns
  static constexpr bool kFused = BN <= 128;
  static constexpr int kTmemCols = kFused ? 2 * BN : BN;
  // stages + barriers (8 B each) + tmem ptr + bias + flow-head weights
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 + BN * 4 + BN * 8 + 16;
};

template <int BN, int KC>
__global__ void __launch_bounds__(kNumThreads, TcCfg<BN, KC>::kMinBlocks) k_conv_tc(const ConvProblem* __restrict__ prob_base) {
  using Cfg = TcCfg<BN, KC>;
  // grid.z selects one of `group` consecutive problems with identical grids (the four parity classes
  // of the NN-upsample + 2x2 conv are one launch)
  const ConvProblem* __restrict__ prob = prob_base + blockIdx.z;
  constexpr int kABytes = Cfg::kABytes;
  extern __shared__ uint8_t smem_raw[];

  // carve shared memory (1024 B alignment required by SWIZZLE_128B)
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - raw);
  const uint32_t bar_base = base + Cfg::kStages * Cfg::kStageBytes;
  // barriers: full[kStages], empty[kStages], tmem_full ; then tmem ptr ; then bias
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * Cfg::kStages);
  uint32_t* tmem_ptr_smem =
      reinterpret_cast<uint32_t*>(gen_base + Cfg::kStages * Cfg::kStageBytes + 8 * (2 * Cfg::kStages + 1));
  float* bias_smem = reinterpret_cast<float*>(gen_base + Cfg::kStages * Cfg::kStageBytes + 256);
  float* w4_smem = bias_smem + BN;  // [BN][2] + b4[2], flow-head mode only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // problem fields -> registers once (asm "memory" clobbers would otherwise reload them from global)
  const int nsrc = prob->nsrc, ntaps = prob->ntaps, cout = prob->cout, epi_mode = prob->epi_mode;
  const bool one = prob->passes == 1;   // single-pass product: hi planes only
  const int tile_h = prob->tile_h, tile_w = prob->tile_w, tiles_x = prob->tiles_x, tiles_y = prob->tiles_y;

  // tile coordinates
  int tile = blockIdx.x;
  const int tx = tile % tiles_x;
  tile /= tiles_x;
  const int ty = tile % tiles_y;
  const int b = tile / tiles_y;
  const int y0 = ty * tile_h, x0 = tx * tile_w;
  const int n0 = blockIdx.y * BN;

  int nkb = 0;
  for (int s = 0; s < nsrc; ++s) nkb += prob->src[s].nchunk * ntaps;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), (uint32_t)Cfg::kTmemCols);
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < BN; i += 128) bias_smem[i] = (n0 + i < cout) ? prob->bias[n0 + i] : 0.f;
    if (epi_mode == 1) {
      for (int i = threadIdx.x - 64; i < 2 * BN; i += 128) w4_smem[i] = (i < 2 * cout) ? prob->head_w4[i] : 0.f;
      if (threadIdx.x - 64 < 2) w4_smem[2 * BN + threadIdx.x - 64] = prob->head_b4[threadIdx.x - 64];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform, elected lane issues) =====================
    const CUtensorMap* tm_w_hi = &prob->tm_w_hi;
    const CUtensorMap* tm_w_lo = &prob->tm_w_lo;
    int kb = 0;
    for (int s = 0; s < nsrc; ++s) {
      const int nchunk = prob->src[s].nchunk, c_off = prob->src[s].c_off;
      const int bs = prob->src[s].bswap ? prob->B - 1 - b : b;
      const CUtensorMap* tm_hi = &prob->tm_a_hi[s];
      const CUtensorMap* tm_lo = &prob->tm_a_lo[s];
      for (int ch = 0; ch < nchunk; ++ch) {
        for (int t = 0; t < ntaps; ++t, ++kb) {
          const int stage = kb % Cfg::kStages;
          const uint32_t phase = (uint32_t)(kb / Cfg::kStages) & 1u;
          const int xx = x0 + prob->tap_dx[t], yy = y0 + prob->tap_dy[t];
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (elect_one()) {
            const uint32_t sa = base + stage * Cfg::kStageBytes;
            mbar_expect_tx(full_bar(stage), one ? (uint32_t)(kABytes + Cfg::kWBytes) : (uint32_t)Cfg::kStageBytes);
            const int cc = c_off + ch * KC;
            tma_load_4d(sa, tm_hi, full_bar(stage), cc, xx, yy, bs);
            tma_load_2d(sa + 2 * kABytes, tm_w_hi, full_bar(stage), kb * KC, n0);
            if (!one) {
              tma_load_4d(sa + kABytes, tm_lo, full_bar(stage), cc, xx, yy, bs);
              tma_load_2d(sa + 2 * kABytes + Cfg::kWBytes, tm_w_lo, full_bar(stage), kb * KC, n0);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform, elected lane issues) =====================
    const uint32_t idesc = make_idesc<BN>();
    const uint32_t idesc2 = make_idesc<(Cfg::kFused ? 2 * BN : BN)>();
    for (int kb = 0; kb < nkb; ++kb) {
      const int stage = kb % Cfg::kStages;
      const uint32_t phase = (uint32_t)(kb / Cfg::kStages) & 1u;
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = base + stage * Cfg::kStageBytes;
        const uint64_t a_hi = make_desc_kc<KC>(sa), a_lo = make_desc_kc<KC>(sa + kABytes);
        const uint64_t w_hi = make_desc_kc<KC>(sa + 2 * kABytes), w_lo = make_desc_kc<KC>(sa + 2 * kABytes + Cfg::kWBytes);
        const uint32_t first = kb == 0 ? 0u : 1u;
#pragma unroll
        for (int k = 0; k < KC / 16; ++k) {
          const uint64_t adv = (uint64_t)(k * 32 >> 4);  // 16 elements x 2 B = 32 B along K
          if (one) {
            umma(tmem_base, a_hi + adv, w_hi + adv, idesc, k == 0 ? first : 1u);
          } else if constexpr (Cfg::kFused) {
            umma(tmem_base, a_hi + adv, w_hi + adv, idesc2, k == 0 ? first : 1u);  // [W_hi ; W_lo]
            umma(tmem_base, a_lo + adv, w_hi + adv, idesc, 1u);
          } else {
            umma(tmem_base, a_lo + adv, w_hi + adv, idesc, k == 0 ? first : 1u);
            umma(tmem_base, a_hi + adv, w_lo + adv, idesc, 1u);
            umma(tmem_base, a_hi + adv, w_hi + adv, idesc, 1u);
          }
        }
        umma_commit(empty_bar(stage));
        if (kb == nkb - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;            // tile row == TMEM lane
    const int py = y0 + r / tile_w, px = x0 + r % tile_w;
    const bool valid = (py < prob->H) && (px < prob->W);
    const int64_t opix = ((int64_t)b * prob->out_H + ((int64_t)py * prob->out_sy + prob->out_oy)) * prob->out_W +
                         ((int64_t)px * prob->out_sx + prob->out_ox);
    const int act = prob->act;
    sp_t* oh = prob->out_hi + opix * prob->out_C + prob->out_c_off + n0;
    sp_t* ol = prob->out_lo + opix * prob->out_C + prob->out_c_off + n0;
    const bool lo_skip = prob->out_lo_skip != 0;
    const float* head_vup = prob->head_vup;
    float* head_res = prob->head_res;
    float* head_v = prob->head_v;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (epi_mode == 1) {
      // flow head: hidden = LeakyReLU(acc + b3) stays in fp32 registers; 2-wide linear layer + v_up
      float r0 = 0.f, r1 = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(t_addr + (uint32_t)(cc * 32), v);
        if (Cfg::kFused && !one) {
          uint32_t u[32];
          tmem_ld32(t_addr + (uint32_t)(BN + cc * 32), u);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        } else {
          tmem_ld_wait();
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float hdn = leaky(__uint_as_float(v[j]) + bias_smem[cc * 32 + j]);
          r0 = fmaf(hdn, w4_smem[(cc * 32 + j) * 2], r0);
          r1 = fmaf(hdn, w4_smem[(cc * 32 + j) * 2 + 1], r1);
        }
      }
      if (valid) {
        float2 res = make_float2(r0 + w4_smem[2 * BN], r1 + w4_smem[2 * BN + 1]);
        float2 tot = res;
        if (head_vup) {
          const float2 u = reinterpret_cast<const float2*>(head_vup)[opix];
          tot.x += u.x;
          tot.y += u.y;
        }
        reinterpret_cast<float2*>(head_res)[opix] = res;
        reinterpret_cast<float2*>(head_v)[opix] = tot;
      }
    } else {
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        if (n0 + cc * 32 >= cout) break;
        uint32_t v[32];
        tmem_ld32(t_addr + (uint32_t)(cc * 32), v);
        if (Cfg::kFused && !one) {
          uint32_t u[32];
          tmem_ld32(t_addr + (uint32_t)(BN + cc * 32), u);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        } else {
          tmem_ld_wait();
        }
        if (valid) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x = __uint_as_float(v[g * 16 + j]) + bias_smem[cc * 32 + g * 16 + j];
              f[j] = act ? leaky(x) : x;
            }
            if (lo_skip) pack_store16_hi(f, oh + cc * 32 + g * 16);
            else pack_store16(f, oh + cc * 32 + g * 16, ol + cc * 32 + g * 16);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)Cfg::kTmemCols);
  }
}

template <int BN, int KC>
cudaError_t launch_bn(const ConvProblem* d_prob, const ConvProblem& h, cudaStream_t st) {
  dim3 grid(h.B * h.tiles_y * h.tiles_x, (h.cout + BN - 1) / BN, h.group > 1 ? h.group : 1);
  k_conv_tc<BN, KC><<<grid, kNumThreads, TcCfg<BN, KC>::kSmemBytes, st>>>(d_prob);
  return cudaGetLastError();
}

}  // namespace

int conv_tc_block_n(int cout) { return cout >= 256 ? 256 : cout >= 128 ? 128 : cout >= 64 ? 64 : 32; }

cudaError_t conv_tc_configure() {
  cudaError_t e;
#define FILM_CFG(BN, KC)                                                                         \
  e = cudaFuncSetAttribute(k_conv_tc<BN, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                           TcCfg<BN, KC>::kSmemBytes);                                           \
  if (e != cudaSuccess) return e;
  FILM_CFG(32, 64) FILM_CFG(64, 64) FILM_CFG(128, 64) FILM_CFG(256, 64) FILM_CFG(32, 32) FILM_CFG(64, 32)
#undef FILM_CFG
  return cudaSuccess;
}

cudaError_t launch_conv_tc(const ConvProblem* d_prob, const ConvProblem& h, cudaStream_t st) {
  const int bn = h.bn;
  if (h.kchunk == 32) {
    if (bn == 64) return launch_bn<64, 32>(d_prob, h, st);
    if (bn == 32) return launch_bn<32, 32>(d_prob, h, st);
    return cudaErrorInvalidValue;  // 32-channel K blocks are only instantiated for Cout <= 64
  }
  switch (bn) {
    case 256: return launch_bn<256, 64>(d_prob, h, st);
    case 128: return launch_bn<128, 64>(d_prob, h, st);
    case 64: return launch_bn<64, 64>(d_prob, h, st);
    default: return launch_bn<32, 64>(d_prob, h, st);
  }
}

}  // namespace film
