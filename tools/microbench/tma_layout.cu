// Microbenchmark: how long does the TMA unit take to land one step's X tile of a fused residual block, as a function of the
// GLOBAL layout of the activation tensor?
//   nhwc   : (n, H, W, chunks, 8) fp16 - what round 1 uses; the box's contiguous run is ONE 8-channel chunk = 16 bytes
//   planar : (n, chunks, H, W, 8) fp16 - the box's contiguous run is a whole row segment, WP x 16 bytes
// Both land the same chunk-planar shared-memory tile [chunk][row][col][8].  Every SM runs `ctas` CTAs that walk through a
// tensor much larger than L2, as the real kernels do.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_layout tools/microbench/tma_layout.cu && ./tma_layout
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../orcai_b200/csrc/tc_common.cuh"
using namespace orcai::tc;

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

__global__ void __launch_bounds__(32) tma_kernel(const __grid_constant__ CUtensorMap tm, int planar, int chunks, int WP, int rows, int H, int n,
                                                 int steps, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  const uint32_t tile = smem_u32(smem + 128);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncwarp();
  const uint32_t plane_bytes = (uint32_t)WP * rows * 16;
  long long total = 0;
  int b = blockIdx.x % n, a = (blockIdx.x * 37) % (H - rows);
  for (int s = 0; s < steps; ++s) {
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(bar, plane_bytes * chunks);
      for (int c = 0; c < chunks; ++c) {
        if (planar) tma_load_5d(tile + c * plane_bytes, &tm, bar, 0, 40, a, c, b);
        else tma_load_5d(tile + c * plane_bytes, &tm, bar, 0, c, 40, a, b);
      }
    }
    __syncwarp();
    mbar_wait(bar, (uint32_t)(s & 1));
    total += clock64() - t0;
    a += rows - 2;
    if (a + rows > H) { a = 0; b = (b + gridDim.x) % n; }
  }
  if (threadIdx.x == 0) out[blockIdx.x] = total / steps;
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  auto encode = reinterpret_cast<EncodeTiledFn>(fp);
  struct Cfg { const char* name; int H, W, chunks, WP, rows, ctas; };
  const Cfg cfgs[] = {{"block 1 (16 ch, 8 x 62 px)", 736, 171, 2, 62, 8, 2}, {"block 2 (30->32 ch, 6 x 90 px)", 368, 86, 4, 90, 6, 1},
                      {"block 3 (40 ch, 6 x 48 px)", 184, 43, 5, 48, 6, 1}, {"block 4 (50->56 ch, 6 x 26 px)", 92, 22, 7, 26, 6, 1}};
  for (const Cfg& k : cfgs) {
    const size_t per = (size_t)k.H * k.W * k.chunks * 8;
    const int n = (int)std::max<size_t>(64, (size_t)(1.2e9) / (per * 2));   // > 1 GB: far larger than L2
    __half* d = nullptr;
    cudaMalloc(&d, per * n * 2);
    cudaMemset(d, 0, per * n * 2);
    for (int planar = 0; planar < 2; ++planar) {
      CUtensorMap tm;
      cuuint64_t dims[5], strides[4];
      cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
      if (planar) {
        dims[0] = 8; dims[1] = k.W; dims[2] = k.H; dims[3] = k.chunks; dims[4] = n;
        strides[0] = 16; strides[1] = (cuuint64_t)k.W * 16; strides[2] = (cuuint64_t)k.H * k.W * 16; strides[3] = (cuuint64_t)k.chunks * k.H * k.W * 16;
        box[0] = 8; box[1] = k.WP; box[2] = k.rows; box[3] = 1; box[4] = 1;
      } else {
        dims[0] = 8; dims[1] = k.chunks; dims[2] = k.W; dims[3] = k.H; dims[4] = n;
        strides[0] = 16; strides[1] = (cuuint64_t)k.chunks * 16; strides[2] = (cuuint64_t)k.W * k.chunks * 16; strides[3] = (cuuint64_t)k.H * k.W * k.chunks * 16;
        box[0] = 8; box[1] = 1; box[2] = k.WP; box[3] = k.rows; box[4] = 1;
      }
      CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      const int grid = 148 * k.ctas;
      long long* out = nullptr;
      cudaMalloc(&out, 8 * grid);
      const int smem = 128 + k.WP * k.rows * 16 * k.chunks;
      cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      for (int rep = 0; rep < 2; ++rep) tma_kernel<<<grid, 32, smem>>>(tm, planar, k.chunks, k.WP, k.rows, k.H, n, 200, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", k.name, cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(grid);
      cudaMemcpy(h.data(), out, 8 * grid, cudaMemcpyDeviceToHost);
      std::sort(h.begin(), h.end());
      printf("%-34s %-6s %d CTA/SM: %6lld cycles per tile (median CTA; min %lld max %lld), %d bytes\n", k.name, planar ? "planar" : "nhwc", k.ctas,
             h[grid / 2], h[0], h[grid - 1], k.WP * k.rows * 16 * k.chunks);
      cudaFree(out);
    }
    cudaFree(d);
  }
  return 0;
}
