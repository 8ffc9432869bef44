// Hardware question behind the "wide halo" conv design: can a K-major SWIZZLE_128B
// UMMA descriptor start at a 128-byte (one pixel) offset inside the 1024-byte swizzle atom, with a
// stride-byte-offset that is not a multiple of 1024 (rows of 10 pixels = 1280 B)?  If yes, ONE halo box
// {64 ch, 10 px, 18 rows} serves all nine taps of a 3x3 convolution instead of three dx-shifted boxes.
//
// smem is filled by hand with the TMA SWIZZLE_128B pattern (16-byte chunk j of the 128-byte row at address
// A lands in chunk j ^ ((A >> 7) & 7)); D = A(tap) x B^T is compared with a CPU product for
//   variant 0: descriptor base_offset = 0
//   variant 1: descriptor base_offset = (start_address >> 7) & 7   (PTX ISA wording)
// Values are small integers, so every result is exact in bf16 x bf16 -> fp32.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o desc_offset_test desc_offset_test.cu && ./desc_offset_test
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../frame_interpolation_b200/csrc/film_tc_ptx.cuh"
using namespace film::tc;

constexpr int kRows = 18, kN = 64;
// kc = 64: 128-byte rows, SWIZZLE_128B (chunk ^= address bits 7..9); kc = 32: 64-byte rows, SWIZZLE_64B
// (chunk ^= address bits 7..8)

__global__ void __launch_bounds__(128, 1)
k_test(const __nv_bfloat16* img, const __nv_bfloat16* wgt, int kC, int pitch_px, int dy, int dx, int variant, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const uint32_t a_off = 0, b_off = 32 * 1024;
  // image: pixel p (row-major, pitch_px wide) at byte offset p*128, 16-byte chunks XOR-swizzled on address bits 7..9
  const int rb = kC * 2, nch = rb / 16, xmask = nch - 1;   // row bytes, 16-byte chunks per row
  for (int i = threadIdx.x; i < kRows * pitch_px * nch; i += blockDim.x) {
    const int p = i / nch, j = i % nch;
    const uint32_t addr = base + a_off + p * rb;
    const int js = j ^ ((addr >> 7) & xmask);
    *reinterpret_cast<uint4*>(sm + a_off + p * rb + js * 16) = *reinterpret_cast<const uint4*>(img + (size_t)p * kC + j * 8);
  }
  for (int i = threadIdx.x; i < kN * nch; i += blockDim.x) {
    const int n = i / nch, j = i % nch;
    const uint32_t addr = base + b_off + n * rb;
    const int js = j ^ ((addr >> 7) & xmask);
    *reinterpret_cast<uint4*>(sm + b_off + n * rb + js * 16) = *reinterpret_cast<const uint4*>(wgt + (size_t)n * kC + j * 8);
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> async proxy (UMMA)
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_ptr), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_start = base + a_off + (dy * pitch_px + dx) * rb;
    const uint64_t type = kC == 64 ? 2 : 4;
    for (int k = 0; k < kC / 16; ++k) {
      const uint32_t sa = a_start + k * 32;
      uint64_t ad = 0;
      ad |= (uint64_t)((sa & 0x3FFFFu) >> 4);
      ad |= (uint64_t)1 << 16;
      ad |= (uint64_t)((pitch_px * rb) >> 4) << 32;  // SBO: next 8-row group = next image row
      ad |= (uint64_t)1 << 46;
      if (variant == 1) ad |= (uint64_t)((a_start >> 7) & 7) << 49;
      ad |= type << 61;
      uint64_t bd = 0;
      bd |= (uint64_t)(((base + b_off + k * 32) & 0x3FFFFu) >> 4);
      bd |= (uint64_t)1 << 16;
      bd |= (uint64_t)((8 * rb) >> 4) << 32;
      bd |= (uint64_t)1 << 46;
      bd |= type << 61;
      umma(tm, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
  }
  __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t r[32];
  for (int c = 0; c < kN; c += 32) {
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(size_t)(warp * 32 + lane) * kN + c + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 64); }
}

int main() {
  const int max_pitch = 16;
  const int kCmax = 64;
  std::vector<__nv_bfloat16> h_img((size_t)kRows * max_pitch * kCmax), h_w((size_t)kN * kCmax);
  std::vector<float> f_img(h_img.size()), f_w(h_w.size());
  srand(7);
  for (size_t i = 0; i < h_img.size(); ++i) { f_img[i] = (float)(rand() % 15 - 7); h_img[i] = __float2bfloat16(f_img[i]); }
  for (size_t i = 0; i < h_w.size(); ++i) { f_w[i] = (float)(rand() % 9 - 4); h_w[i] = __float2bfloat16(f_w[i]); }
  __nv_bfloat16 *d_img, *d_w;
  float* d_out;
  cudaMalloc(&d_img, h_img.size() * 2);
  cudaMalloc(&d_w, h_w.size() * 2);
  cudaMalloc(&d_out, 128 * kN * 4);
  cudaMemcpy(d_img, h_img.data(), h_img.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d_w, h_w.data(), h_w.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  struct Case { int kc, pitch, dy, dx; } cases[] = {{64, 8, 0, 0}, {64, 8, 2, 0}, {64, 10, 0, 0}, {64, 10, 0, 1}, {64, 10, 0, 2},
                                                    {64, 10, 1, 1}, {64, 10, 2, 2}, {64, 9, 1, 1},
                                                    {32, 8, 0, 0}, {32, 8, 2, 0}, {32, 10, 0, 0}, {32, 10, 0, 1}, {32, 10, 0, 2},
                                                    {32, 10, 1, 1}, {32, 10, 2, 2}, {32, 9, 1, 1}};
  std::vector<float> h_out(128 * kN);
  for (auto& c : cases) {
    for (int variant = 0; variant < 2; ++variant) {
      cudaMemset(d_out, 0, 128 * kN * 4);
      k_test<<<1, 128, 64 * 1024>>>(d_img, d_w, c.kc, c.pitch, c.dy, c.dx, variant, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kc %d pitch %d dy %d dx %d variant %d: %s\n", c.kc, c.pitch, c.dy, c.dx, variant, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h_out.data(), d_out, 128 * kN * 4, cudaMemcpyDeviceToHost);
      double max_err = 0;
      int bad = 0;
      for (int r = 0; r < 128; ++r) {
        const int p = (r / 8 + c.dy) * c.pitch + (r % 8) + c.dx;
        for (int n = 0; n < kN; ++n) {
          double acc = 0;
          for (int k = 0; k < c.kc; ++k) acc += (double)f_img[(size_t)p * c.kc + k] * f_w[(size_t)n * c.kc + k];
          const double err = fabs(acc - h_out[(size_t)r * kN + n]);
          if (err > max_err) max_err = err;
          if (err != 0) ++bad;
        }
      }
      printf("kc %d  pitch %2d px  tap (dy %d, dx %d)  base_offset %s : max |err| %.1f, %d / %d wrong  %s\n", c.kc, c.pitch, c.dy, c.dx,
             variant ? "(addr>>7)&7" : "0          ", max_err, bad, 128 * kN, bad ? "MISMATCH" : "exact");
    }
  }
  return 0;
}
