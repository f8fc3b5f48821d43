// Microbenchmark behind DESIGN.md's "what bounds the fused residual-block kernels": cycles per tcgen05.mma (M=128, K=16, fp16,
// fp32 accumulate) as a function of N and of where the A operand lives, with the kernels' own operand layout
// (K-major, no swizzle, 8x16-byte core matrices, SBO = 128 B, LBO = one chunk plane).
//   SS : A and B from shared memory (what net_fused.cuh issues)            -> operand fetch = (128*16 + N*16)*2 bytes per MMA
//   TS : A from tensor memory (staged once with tcgen05.cp), B from smem  -> operand fetch = N*16*2 bytes per MMA
//   CP : the tcgen05.cp 128x256b smem -> TMEM staging copy on its own
// Also checks that a TS-mode MMA on an A tile staged by tcgen05.cp.128x256b reproduces the SS-mode result bit for bit.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mma_cost tools/microbench/mma_cost.cu && ./mma_cost
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../../orcai_b200/csrc/tc_common.cuh"
using namespace orcai::tc;

__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred)::"memory");
  return pred != 0;
}

constexpr int kAPix = 1024;                 // pixels (rows) per chunk plane of the A buffer
constexpr uint32_t kLboA = kAPix * 16;      // chunk plane stride
constexpr uint32_t OFF_A = 0, OFF_B = 2 * kLboA, OFF_BAR = OFF_B + 256 * 16 * 2 * 2, SMEM = OFF_BAR + 64;

// mode 0: SS, 1: TS, 2: CP only, 3: TS with one CP per 3 MMAs (the "one staged copy serves the three dy taps" pattern)
template <int N, int MODE, int CTAS, int NACC = 2, int NISS = 1>
__global__ void __launch_bounds__(128, CTAS) cost_kernel(long long* out_cycles, int reps, int per_commit) {
  constexpr int COLS = 512 / CTAS;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (int)(OFF_BAR / 4); i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0 pairs
  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); fence_mbar_init(); }
  __syncwarp();
  if (warp == 0) tmem_alloc<COLS>(tslot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t sbase = smem_u32(smem);
  if (warp >= 1 && warp <= NISS) {   // NISS issuing warps, each with its own accumulators and its own completion barrier
    bar += warp - 1;
    constexpr uint32_t idesc = make_idesc_f16(128, N, 0);
    static_assert(NACC * N + 32 <= COLS || NACC == 2, "accumulators must fit");
    constexpr uint32_t D1 = (2 * N + 32 <= COLS) ? N : 0;   // second accumulator tile (or the same one when TMEM is short)
    const uint64_t dA = make_smem_desc(sbase + OFF_A, kLboA, 128);
    const uint64_t dB = make_smem_desc(sbase + OFF_B, 128, 256);
    const uint32_t a_tm = tmem + COLS - 32;   // 8 columns per staged K=16 A tile
    long long t0 = 0, t1 = 0;
    uint32_t phase = 0;
    for (int pass = 0; pass < 2; ++pass) {   // pass 0 warms up
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (elect_one()) {
          for (int i = 0; i < per_commit; ++i) {
            const uint32_t d = tmem + (uint32_t)(warp - 1) * (NISS > 1 ? 2 * N : 0) + (NACC == 2 ? (uint32_t)(i & 1) * D1 : (uint32_t)(i % NACC) * N);   // NACC independent accumulation chains
            const uint32_t aoff = (uint32_t)((i % 9) / 3 * 62 + (i % 3));   // tap shift in pixels = 16-byte units
            if (MODE == 0) mma_f16_ss(d, dA + aoff, dB, idesc, i >= NACC);
            if (MODE == 1) mma_f16_ts(d, a_tm + (uint32_t)(i % 3) * 8, dB, idesc, i >= NACC);
            if (MODE == 2) tmem_cp_128x256b(a_tm + (uint32_t)(i % 3) * 8, dA + aoff);
            if (MODE == 3) {
              if (i % 3 == 0) tmem_cp_128x256b(a_tm + (uint32_t)((i / 3) % 3) * 8, dA + aoff);
              mma_f16_ts(d, a_tm + (uint32_t)((i / 3) % 3) * 8, dB, idesc, i >= NACC);
            }
          }
          mma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, phase);
        phase ^= 1;
      }
      t1 = clock64();
    }
    if ((tid & 31) == 0 && warp == 1) out_cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<COLS>(tmem);
}

// ---- correctness: TS-mode MMA on a tcgen05.cp-staged A tile == SS-mode MMA ----
__global__ void __launch_bounds__(128, 1) check_kernel(const __half* a_in, const __half* b_in, float* d_ss, float* d_ts) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A: 128 rows x 16 k, canonical layout: byte = (row/8)*128 + (row%8)*16 + (k/8)*LBO + (k%8)*2 ; B: 32 rows(n) x 16 k, SBO 256, LBO 128
  for (int i = tid; i < 128 * 16; i += 128) {
    const int row = i / 16, k = i % 16;
    *reinterpret_cast<__half*>(smem + OFF_A + (row / 8) * 128 + (row % 8) * 16 + (k / 8) * kLboA + (k % 8) * 2) = a_in[i];
  }
  for (int i = tid; i < 32 * 16; i += 128) {
    const int n = i / 16, k = i % 16;
    *reinterpret_cast<__half*>(smem + OFF_B + (n / 8) * 256 + (n % 8) * 16 + (k / 8) * 128 + (k % 8) * 2) = b_in[i];
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
      constexpr uint32_t idesc = make_idesc_f16(128, 32, 0);
      const uint64_t dA = make_smem_desc(sbase + OFF_A, kLboA, 128);
      const uint64_t dB = make_smem_desc(sbase + OFF_B, 128, 256);
      mma_f16_ss(tmem + 0, dA, dB, idesc, 0);
      tmem_cp_128x256b(tmem + 64, dA);
      mma_f16_ts(tmem + 32, tmem + 64, dB, idesc, 0);
      mma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  float v[16];
  const uint32_t la = tmem + ((uint32_t)(warp * 32) << 16);
  for (int h = 0; h < 2; ++h) {
    tmem_ld16(la + h * 16, v);
    for (int j = 0; j < 16; ++j) d_ss[(warp * 32 + lane) * 32 + h * 16 + j] = v[j];
    tmem_ld16(la + 32 + h * 16, v);
    for (int j = 0; j < 16; ++j) d_ts[(warp * 32 + lane) * 32 + h * 16 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem);
}

template <int N, int MODE, int CTAS, int NACC = 2, int NISS = 1>
void run(const char* name, int sms) {
  const int grid = sms * CTAS;
  long long* d = nullptr;
  cudaMalloc(&d, sizeof(long long) * grid);
  cudaFuncSetAttribute(cost_kernel<N, MODE, CTAS, NACC, NISS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
  const int reps = 200, per = 54;
  cost_kernel<N, MODE, CTAS, NACC, NISS><<<grid, 128, SMEM>>>(d, reps, per);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-24s N=%3d: %s\n", name, N, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  const double ops = (double)reps * per * CTAS * NISS;   // per SM
  printf("%-24s N=%3d  %d CTA/SM x %d issuer warp(s), %d accumulators, %3d SMs: %6.1f cycles per op per SM (median CTA; min %.1f max %.1f)\n", name, N, CTAS, NISS, NACC, sms, h[grid / 2] / ops,
         h[0] / ops, h[grid - 1] / ops);
  cudaFree(d);
}

int main() {
  // correctness first
  {
    std::vector<__half> a(128 * 16), b(32 * 16);
    srand(1);
    for (auto& x : a) x = __float2half((rand() % 2001 - 1000) / 1000.f);
    for (auto& x : b) x = __float2half((rand() % 2001 - 1000) / 1000.f);
    __half *da, *db; float *dss, *dts;
    cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dss, 128 * 32 * 4); cudaMalloc(&dts, 128 * 32 * 4);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dts, 0, 128 * 32 * 4);
    cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    check_kernel<<<1, 128, SMEM>>>(da, db, dss, dts);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("check kernel: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> ss(128 * 32), ts(128 * 32);
    cudaMemcpy(ss.data(), dss, ss.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(ts.data(), dts, ts.size() * 4, cudaMemcpyDeviceToHost);
    double err_ref = 0, err_ts = 0;
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < 32; ++n) {
        double acc = 0;
        for (int k = 0; k < 16; ++k) acc += (double)__half2float(a[r * 16 + k]) * (double)__half2float(b[n * 16 + k]);
        err_ref = std::max(err_ref, std::abs(acc - ss[r * 32 + n]));
        err_ts = std::max(err_ts, (double)std::abs(ts[r * 32 + n] - ss[r * 32 + n]));
      }
    printf("check: SS vs host max err %.3e ; TS (A staged by tcgen05.cp.128x256b) vs SS max diff %.3e\n", err_ref, err_ts);
  }
  const int sms = 148;
  run<32, 0, 1>("SS (A,B from smem)", sms);
  run<64, 0, 1>("SS (A,B from smem)", sms);
  run<96, 0, 1>("SS (A,B from smem)", sms);
  run<128, 0, 1>("SS (A,B from smem)", sms);
  run<256, 0, 1>("SS (A,B from smem)", sms);
  run<32, 0, 1, 1>("SS (A,B from smem)", sms);
  run<32, 0, 1, 3>("SS (A,B from smem)", sms);
  run<32, 0, 1, 4>("SS (A,B from smem)", sms);
  run<32, 0, 1, 6>("SS (A,B from smem)", sms);
  run<48, 0, 1, 1>("SS (A,B from smem)", sms);
  run<48, 0, 1, 3>("SS (A,B from smem)", sms);
  run<48, 0, 1, 6>("SS (A,B from smem)", sms);
  run<64, 0, 1, 1>("SS (A,B from smem)", sms);
  run<64, 0, 1, 3>("SS (A,B from smem)", sms);
  run<64, 0, 1, 6>("SS (A,B from smem)", sms);
  run<32, 0, 1, 2, 2>("SS (A,B from smem)", sms);
  run<48, 0, 1, 2, 2>("SS (A,B from smem)", sms);
  run<64, 0, 1, 2, 2>("SS (A,B from smem)", sms);
  run<96, 0, 1, 2, 2>("SS (A,B from smem)", sms);
  run<32, 0, 2>("SS (A,B from smem)", sms);
  run<32, 0, 2, 1>("SS (A,B from smem)", sms);
  run<32, 0, 2, 3>("SS (A,B from smem)", sms);
  run<64, 0, 2>("SS (A,B from smem)", sms);
  run<96, 0, 2>("SS (A,B from smem)", sms);
  run<32, 1, 1>("TS (A from TMEM)", sms);
  run<64, 1, 1>("TS (A from TMEM)", sms);
  run<128, 1, 1>("TS (A from TMEM)", sms);
  run<32, 1, 2>("TS (A from TMEM)", sms);
  run<32, 2, 1>("CP 128x256b only", sms);
  run<32, 2, 2>("CP 128x256b only", sms);
  run<32, 3, 1>("TS + 1 CP per 3 MMAs", sms);
  run<64, 3, 1>("TS + 1 CP per 3 MMAs", sms);
  run<32, 3, 2>("TS + 1 CP per 3 MMAs", sms);
  return 0;
}
