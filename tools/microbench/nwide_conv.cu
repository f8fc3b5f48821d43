// Building block for the N-widened residual blocks, checked on ONE tile:
// a 3x3 convolution over a pixel-linear, chunk-planar fp16 image (row pitch WP) computed as
//     Z[q, (dx, n)] = sum_dy sum_k A[q + dy*WP][k] * W[dy][dx][k][n]          3 MMAs (one per dy), N = 3 * 32 = 96
//     out[p, n]     = Z[p-1, (0, n)] + Z[p, (1, n)] + Z[p+1, (2, n)]            warp shuffles in the epilogue
// where the A descriptor's 8-row groups start every SIX pixels (SBO = 96 B): group g holds pixels 6g .. 6g+7, so rows 1..6 of a
// group find both neighbours inside the group and never cross a warp or a TMEM lane quadrant.  128 MMA rows -> 96 outputs.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o nwide_conv tools/microbench/nwide_conv.cu && ./nwide_conv
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../orcai_b200/csrc/tc_common.cuh"
using namespace orcai::tc;

constexpr int WP = 62, CIN = 16, COUT = 30, NPIX = 96 + 2 * WP + 8 + 64;   // pixels resident in shared memory
constexpr uint32_t LBO_A = NPIX * 16;
constexpr uint32_t OFF_A = 0, OFF_W = 2 * LBO_A;                  // 2 chunk planes of A
constexpr uint32_t W_DY = 96 * 16 * 2;                             // one dy: 96 rows (dx, n) x 16 k, canonical: SBO 256 per 8 rows, LBO 128
constexpr uint32_t OFF_BAR = OFF_W + 3 * W_DY, SMEM = OFF_BAR + 64;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred)::"memory");
  return pred != 0;
}

// a: (NPIX, 16) fp16 pixel-major; w: (3, 3, 16, 32) fp16 [dy][dx][k][n]; out: (96, 32) float for pixels p0 .. p0+95, p0 = WP + 1
__global__ void __launch_bounds__(128, 1) nwide_kernel(const __half* a, const __half* w, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < NPIX * 16; i += 128) {
    const int p = i / 16, k = i % 16;
    *reinterpret_cast<__half*>(smem + OFF_A + (k / 8) * LBO_A + p * 16 + (k % 8) * 2) = a[i];
  }
  for (int i = tid; i < 3 * 3 * 16 * 32; i += 128) {
    const int n = i % 32, k = (i / 32) % 16, dx = (i / 512) % 3, dy = i / 1536;
    const int row = dx * 32 + n;
    *reinterpret_cast<__half*>(smem + OFF_W + dy * W_DY + (row / 8) * 256 + (row % 8) * 16 + (k / 8) * 128 + (k % 8) * 2) = w[i];
  }
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncwarp();
  if (warp == 0) tmem_alloc<128>(tslot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t sbase = smem_u32(smem);
  if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(128, 96, 0);
      // MMA row r = 8g + i  <->  pixel q = q0 + 6g + i ; q0 = first output pixel - 1 (row 0 of group 0 is the left neighbour)
      const uint32_t q0 = WP + 1 - 1;
      for (int dy = 0; dy < 3; ++dy) {
        const uint64_t dA = make_smem_desc(sbase + OFF_A + (q0 + (dy - 1) * WP) * 16, LBO_A, 96);    // SBO = 96 B: groups every 6 pixels
        const uint64_t dB = make_smem_desc(sbase + OFF_W + dy * W_DY, 128, 256);
        mma_f16_ss(tmem, dA, dB, idesc, dy != 0);
      }
      mma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const uint32_t la = tmem + ((uint32_t)(warp * 32) << 16);
  const int g = warp * 4 + lane / 8, i = lane % 8;
  for (int h = 0; h < 2; ++h) {     // 16 output channels at a time
    float z0[16], z1[16], z2[16];
    tmem_ld16(la + 0 * 32 + h * 16, z0);
    tmem_ld16(la + 1 * 32 + h * 16, z1);
    tmem_ld16(la + 2 * 32 + h * 16, z2);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float left = __shfl_up_sync(0xffffffffu, z0[c], 1);      // Z[q-1, dx=0]
      const float right = __shfl_down_sync(0xffffffffu, z2[c], 1);   // Z[q+1, dx=2]
      const float v = left + z1[c] + right;
      if (i >= 1 && i <= 6) out[(6 * g + i - 1) * 32 + h * 16 + c] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem);
}

int main() {
  std::vector<__half> a(NPIX * 16), w(3 * 3 * 16 * 32);
  srand(3);
  for (auto& x : a) x = __float2half((rand() % 2001 - 1000) / 1000.f);
  for (size_t i = 0; i < w.size(); ++i) w[i] = __float2half((i % 32) < COUT ? (rand() % 2001 - 1000) / 4000.f : 0.f);
  __half *da, *dw; float* dout;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&dw, w.size() * 2); cudaMalloc(&dout, 96 * 32 * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dw, w.data(), w.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xff, 96 * 32 * 4);
  cudaFuncSetAttribute(nwide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
  nwide_kernel<<<1, 128, SMEM>>>(da, dw, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> out(96 * 32);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  double err = 0, mag = 0;
  for (int o = 0; o < 96; ++o) {
    const int p = WP + 1 + o;
    for (int n = 0; n < 32; ++n) {
      double acc = 0;
      for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx)
          for (int k = 0; k < 16; ++k)
            acc += (double)__half2float(a[(p + (dy - 1) * WP + (dx - 1)) * 16 + k]) * (double)__half2float(w[((dy * 3 + dx) * 16 + k) * 32 + n]);
      err = std::max(err, std::abs(acc - out[o * 32 + n]));
      mag = std::max(mag, std::abs(acc));
    }
  }
  printf("N-widened 3x3 convolution tile (3 MMAs of N = 96, SBO = 96 B, shuffle epilogue): max |err| %.3e (max |ref| %.2f) over 96 pixels x 32 channels -> %s\n",
         err, mag, err < 1e-4 ? "OK" : "MISMATCH");
  return err < 1e-4 ? 0 : 1;
}
