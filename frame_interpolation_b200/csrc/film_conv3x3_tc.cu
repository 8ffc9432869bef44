// Persistent tcgen05 3x3 convolution with activation-tile reuse across taps (sm_100a).
//
// Why a second kernel: ncu on the generic kernel (profiles/r1_ncu_conv.md) shows the N <= 64
// layers spend their time re-loading the same activation pixels once per tap (9 x 32 KiB per
// 64-channel chunk and tile) and serialising prologue / mainloop / epilogue per tile.  Here:
//
//  * Tile = 16 rows x 8 columns of output pixels.  For a 64-channel chunk the producer loads
//    THREE boxes (dx = -1, 0, +1), each (64 ch, 8 px, 18 rows) of both planes: 144 pixel rows
//    of 128 B = 18 KiB per plane.  Because a tile row is exactly one 1024-byte swizzle atom
//    (8 px x 128 B), the operand of tap (dy, dx) is the SAME smem box at byte offset dy * 1024:
//    3 loads serve 9 taps (2.67x less L2 -> smem traffic) with plain, 1024-aligned UMMA descriptors.
//  * Weights: if the whole [Cout x K] hi+lo matrix fits (<= 144 KiB: 64->64, 128->32, 64->32 ...)
//    it is loaded ONCE per CTA and stays resident; otherwise it streams through its own ring,
//    one tap ([BN x 64] hi+lo) per stage.
//  * Persistent CTAs (grid = #SMs) walk a static tile list; the fp32 accumulator is double-buffered
//    in TMEM (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Roles: warp 0 = TMA producer (+ L2 prefetch of the next tile's boxes: the first touch of an
// activation tile comes from DRAM, ~2.5 us under load, and the ring holds only 2-4 stages),
// warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "film_conv.h"
#include "film_tc_ptx.cuh"

namespace film {
namespace {
using namespace tc;

constexpr int kEpiWarps = 8;                        // two warps per TMEM lane quarter, alternating 16-column chunks
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kTileH = 16, kTileW = 8;
constexpr int kBoxRows = (kTileH + 2) * kTileW;     // 144 pixel rows per dx-copy
constexpr int kAPlane = kBoxRows * 128;             // 18432 B
constexpr int kAStage = 2 * kAPlane;                // hi + lo = 36864 B
constexpr int kMaxRing = 8;
constexpr int kSmemLimit = 227 * 1024;

__host__ __device__ inline int w_tap_bytes(int bn) { return bn * 128 * 2; }  // [BN x 64] hi + lo

template <int BN>
__global__ void __launch_bounds__(kThreads, 1) k_conv3x3_tc(const ConvProblem* __restrict__ prob) {
  extern __shared__ uint8_t smem_raw[];
  const ConvProblem& P = *prob;
  const int NA = P.v2_na, NW = P.v2_nw;
  const bool resident = P.v2_resident != 0;
  constexpr int kWTap = BN * 128 * 2;
  // Fused-N product (BN <= 128): SS-mode MMAs with small N are bound by the shared-memory read of
  // the A operand (~64 B/clk -> ~64-85 cycles per M=128,K=16 instruction, measured), not by math.
  // W_hi and W_lo blocks are contiguous in smem, so  A_hi x [W_hi ; W_lo]  is ONE MMA with N = 2*BN
  // (columns [0,BN) = hi*hi, [BN,2BN) = hi*lo) and  A_lo x W_hi  accumulates into columns [0,BN):
  // 2 A-operand reads per k-step instead of 3.  The epilogue adds the two column halves.
  constexpr bool kFused = BN <= 128;
  constexpr uint32_t kAccCols = kFused ? 2 * BN : BN;
  constexpr uint32_t kTmemCols = 2 * kAccCols;

  int nkb = 0;  // K blocks = (source, chunk, dx, dy)
  for (int s = 0; s < P.nsrc; ++s) nkb += P.src[s].nchunk * 9;

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - raw);
  const uint32_t a_base = base;
  const uint32_t w_base = a_base + (uint32_t)NA * kAStage;
  const uint32_t w_bytes = resident ? (uint32_t)nkb * kWTap : (uint32_t)NW * kWTap;
  const uint32_t tail = w_base + w_bytes;  // barriers etc.
  auto a_full = [&](int s) { return tail + 8u * s; };
  auto a_empty = [&](int s) { return tail + 8u * (kMaxRing + s); };
  auto w_full = [&](int s) { return tail + 8u * (2 * kMaxRing + s); };
  auto w_empty = [&](int s) { return tail + 8u * (3 * kMaxRing + s); };
  auto t_full = [&](int s) { return tail + 8u * (4 * kMaxRing + s); };
  auto t_empty = [&](int s) { return tail + 8u * (4 * kMaxRing + 2 + s); };
  const uint32_t tail_off = (uint32_t)NA * kAStage + w_bytes;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gen_base + tail_off + 8 * (4 * kMaxRing + 4));
  float* bias_smem = reinterpret_cast<float*>(gen_base + tail_off + 8 * (4 * kMaxRing + 4) + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = P.tiles_y * P.tiles_x;
  const int n_nt = (P.cout + BN - 1) / BN;            // N tiles (Cout = 512 -> 2)
  const int ntiles = P.B * tiles_per_img * n_nt;      // work items: (spatial tile, N tile), N fastest

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kMaxRing; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), kEpiWarps);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_ptr_smem), kTmemCols);
  if (warp >= 2)
    for (int i = threadIdx.x - 64; i < n_nt * BN; i += 32 * kEpiWarps) bias_smem[i] = (i < P.cout) ? P.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      if (resident) {
        // whole weight matrix, once: nkb blocks of [BN x 64] hi then lo
        mbar_expect_tx(w_full(0), (uint32_t)nkb * kWTap);
        for (int kb = 0; kb < nkb; ++kb) {
          tma_load_2d(w_base + kb * kWTap, &P.tm_w_hi, w_full(0), kb * kChunk, 0);
          tma_load_2d(w_base + kb * kWTap + kWTap / 2, &P.tm_w_lo, w_full(0), kb * kChunk, 0);
        }
      }
      uint32_t ia = 0, iw = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int sp = tile / n_nt, n0 = (tile % n_nt) * BN;
        const int b = sp / tiles_per_img, rem = sp % tiles_per_img;
        const int y0 = (rem / P.tiles_x) * kTileH, x0 = (rem % P.tiles_x) * kTileW;
        {
          // L2 prefetch of this CTA's NEXT spatial tile (skip if it is the same spatial tile, other N half)
          const int nt = tile + gridDim.x;
          if (nt < ntiles && nt / n_nt != sp) {
            const int nsp = nt / n_nt, nb = nsp / tiles_per_img, nrem = nsp % tiles_per_img;
            const int ny0 = (nrem / P.tiles_x) * kTileH, nx0 = (nrem % P.tiles_x) * kTileW;
            for (int s = 0; s < P.nsrc; ++s)
              for (int ch = 0; ch < P.src[s].nchunk; ++ch) {
                // the three dx boxes overlap: one 10-px-wide region == boxes at dx = 0 and dx = 2
                tma_prefetch_4d(&P.tm_a_hi[s], P.src[s].c_off + ch * kChunk, nx0 - 1, ny0 - 1, nb);
                tma_prefetch_4d(&P.tm_a_hi[s], P.src[s].c_off + ch * kChunk, nx0 + 1, ny0 - 1, nb);
                tma_prefetch_4d(&P.tm_a_lo[s], P.src[s].c_off + ch * kChunk, nx0 - 1, ny0 - 1, nb);
                tma_prefetch_4d(&P.tm_a_lo[s], P.src[s].c_off + ch * kChunk, nx0 + 1, ny0 - 1, nb);
              }
          }
        }
        int kb = 0;
        for (int s = 0; s < P.nsrc; ++s) {
          const int nchunk = P.src[s].nchunk, c_off = P.src[s].c_off;
          for (int ch = 0; ch < nchunk; ++ch) {
            for (int dx = 0; dx < 3; ++dx) {
              const int st = ia % NA;
              mbar_wait(a_empty(st), ((ia / NA) & 1u) ^ 1u);
              const uint32_t sa = a_base + st * kAStage;
              if ((P.dbg_flags & 2) && ia >= (uint32_t)NA) {
                mbar_arrive(a_full(st));
              } else {
                mbar_expect_tx(a_full(st), kAStage);
                tma_load_4d(sa, &P.tm_a_hi[s], a_full(st), c_off + ch * kChunk, x0 + dx - 1, y0 - 1, b);
                tma_load_4d(sa + kAPlane, &P.tm_a_lo[s], a_full(st), c_off + ch * kChunk, x0 + dx - 1, y0 - 1, b);
              }
              ++ia;
              if (!resident) {
                for (int dy = 0; dy < 3; ++dy, ++kb) {
                  const int ws = iw % NW;
                  mbar_wait(w_empty(ws), ((iw / NW) & 1u) ^ 1u);
                  mbar_expect_tx(w_full(ws), kWTap);
                  const uint32_t sw = w_base + ws * kWTap;
                  tma_load_2d(sw, &P.tm_w_hi, w_full(ws), kb * kChunk, n0);
                  tma_load_2d(sw + kWTap / 2, &P.tm_w_lo, w_full(ws), kb * kChunk, n0);
                  ++iw;
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc<BN>();
      const uint32_t idesc2 = make_idesc<(kFused ? 2 * BN : BN)>();
      if (resident) {
        mbar_wait(w_full(0), 0);
        tc_fence_after();
      }
      uint32_t ia = 0, iw = 0, it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1u;
        mbar_wait(t_empty(acc), ((it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccCols;
        int kb = 0;
        bool first = true;
        const int nab = nkb / 3;  // activation stages per tile
        for (int ab = 0; ab < nab; ++ab) {
          const int st = ia % NA;
          mbar_wait(a_full(st), (ia / NA) & 1u);
          tc_fence_after();
          const uint32_t sa = a_base + st * kAStage;
          for (int dy = 0; dy < 3; ++dy, ++kb) {
            uint32_t sw;
            int ws = 0;
            if (resident) {
              sw = w_base + kb * kWTap;
            } else {
              ws = iw % NW;
              mbar_wait(w_full(ws), (iw / NW) & 1u);
              tc_fence_after();
              sw = w_base + ws * kWTap;
            }
            const uint64_t a_hi = make_desc(sa + dy * 1024), a_lo = make_desc(sa + kAPlane + dy * 1024);
            const uint64_t w_hi = make_desc(sw), w_lo = make_desc(sw + kWTap / 2);
            if (!(P.dbg_flags & 4)) {
#pragma unroll
              for (int k = 0; k < kChunk / 16; ++k) {
                const uint64_t adv = (uint64_t)(k * 32 >> 4);
                if constexpr (kFused) {
                  umma(d_tmem, a_hi + adv, w_hi + adv, idesc2, first ? 0u : 1u);  // N = 2*BN: [W_hi ; W_lo]
                  first = false;
                  umma(d_tmem, a_lo + adv, w_hi + adv, idesc, 1u);
                } else {
                  umma(d_tmem, a_lo + adv, w_hi + adv, idesc, first ? 0u : 1u);
                  first = false;
                  umma(d_tmem, a_hi + adv, w_lo + adv, idesc, 1u);
                  umma(d_tmem, a_hi + adv, w_hi + adv, idesc, 1u);
                }
              }
            }
            if (!resident) {
              umma_commit(w_empty(ws));
              ++iw;
            }
          }
          umma_commit(a_empty(st));
          ++ia;
        }
        umma_commit(t_full(acc));
      }
    }
  } else {
    // ============================ epilogue (warps 2..9) ============================
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;    // which 16-column chunks (even / odd) this warp drains
    const int r = q * 32 + lane;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1u;
      const int sp = tile / n_nt, n0 = (tile % n_nt) * BN;
      const int b = sp / tiles_per_img, rem = sp % tiles_per_img;
      const int py = (rem / P.tiles_x) * kTileH + r / kTileW, px = (rem % P.tiles_x) * kTileW + r % kTileW;
      const bool valid = (py < P.H) && (px < P.W);
      const int64_t opix = ((int64_t)b * P.out_H + py) * P.out_W + px;
      sp_t* oh = P.out_hi + opix * P.out_C + P.out_c_off + n0;
      sp_t* ol = P.out_lo + opix * P.out_C + P.out_c_off + n0;
      mbar_wait(t_full(acc), (it >> 1) & 1u);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * kAccCols + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int cc = half; cc < BN / 16; cc += 2) {
        if (n0 + cc * 16 >= P.cout) break;
        uint32_t v[16];
        tmem_ld16(t_addr + (uint32_t)(cc * 16), v);
        if constexpr (kFused) {
          uint32_t u[16];
          tmem_ld16(t_addr + (uint32_t)(BN + cc * 16), u);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        } else {
          tmem_ld_wait();
        }
        if (valid && !(P.dbg_flags & 1)) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float x = __uint_as_float(v[g * 8 + j]) + bias_smem[n0 + cc * 16 + g * 8 + j];
              f[j] = P.act ? leaky(x) : x;
            }
            uint4 h, l;
            pack8(f, h, l);
            *reinterpret_cast<uint4*>(oh + cc * 16 + g * 8) = h;
            *reinterpret_cast<uint4*>(ol + cc * 16 + g * 8) = l;
          }
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

int smem_bytes_for(const ConvProblem& h, int bn) {
  const int nkb = h.ktot / kChunk;
  const int w = h.v2_resident ? nkb * w_tap_bytes(bn) : h.v2_nw * w_tap_bytes(bn);
  return h.v2_na * kAStage + w + 8 * (4 * kMaxRing + 4) + 16 + 512 * 4 + 1024 + 64;
}

}  // namespace

int conv_tc_block_n(int cout);

// Chooses resident/streamed weights and the ring depths for one 3x3 problem (Cout <= 256).
void conv3x3_tc_plan(ConvProblem& h, int num_sms) {
  const int bn = conv_tc_block_n(h.cout);
  const int nkb = h.ktot / kChunk;
  const int wtap = w_tap_bytes(bn);
  const int fixed = 8 * (4 * kMaxRing + 4) + 16 + 512 * 4 + 1024 + 64;
  const int w_all = nkb * wtap;
  h.v2_resident = 0;
  if (h.cout <= bn && w_all + 2 * kAStage + fixed <= kSmemLimit) {
    h.v2_resident = 1;
    int na = (kSmemLimit - fixed - w_all) / kAStage;
    h.v2_na = na > 6 ? 6 : na;
    h.v2_nw = 1;
  } else {
    h.v2_na = bn >= 128 ? 2 : 3;
    int nw = (kSmemLimit - fixed - h.v2_na * kAStage) / wtap;
    h.v2_nw = nw > kMaxRing ? kMaxRing : nw;
  }
  const int ntiles = h.B * h.tiles_y * h.tiles_x * ((h.cout + bn - 1) / bn);
  h.v2_grid = ntiles < num_sms ? ntiles : num_sms;
}

cudaError_t conv3x3_tc_configure() {
  cudaError_t e;
#define FILM_CFG(BN)                                                                                         \
  e = cudaFuncSetAttribute(k_conv3x3_tc<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);       \
  if (e != cudaSuccess) return e;
  FILM_CFG(32) FILM_CFG(64) FILM_CFG(128) FILM_CFG(256)
#undef FILM_CFG
  return cudaSuccess;
}

cudaError_t launch_conv3x3_tc(const ConvProblem* d_prob, const ConvProblem& h, cudaStream_t st) {
  const int bn = conv_tc_block_n(h.cout);
  const int smem = smem_bytes_for(h, bn);
  switch (bn) {
    case 256: k_conv3x3_tc<256><<<h.v2_grid, kThreads, smem, st>>>(d_prob); break;
    case 128: k_conv3x3_tc<128><<<h.v2_grid, kThreads, smem, st>>>(d_prob); break;
    case 64: k_conv3x3_tc<64><<<h.v2_grid, kThreads, smem, st>>>(d_prob); break;
    default: k_conv3x3_tc<32><<<h.v2_grid, kThreads, smem, st>>>(d_prob); break;
  }
  return cudaGetLastError();
}

}  // namespace film
