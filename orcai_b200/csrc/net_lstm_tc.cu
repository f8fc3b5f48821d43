// Recurrent tail of orcai-V1 on the tensor cores: 2 x Bidirectional(LSTM(128)) + Dense(128, relu) (+ BN, Dense(7, sigmoid)).
//
// Reference graph: src/orcAI/architectures.py:210-239 (gate order i, f, c, o; sigmoid recurrent activation; the backward
// layer runs t = Tn-1 .. 0 and is written back in forward order; [forward | backward] concatenation).
//
//   gemm_tc_kernel      C[M, N] = act(A[M, K] * B[K, N] + bias) ; A fp32 or fp16 in HBM, staged as fp16 into the canonical
//                       K-major no-swizzle UMMA layout by the worker warps, B pre-packed fp16 blocks fetched with one
//                       cp.async.bulk per K-chunk, fp32 accumulation in TMEM, 2-stage mbarrier pipeline, dedicated issuer warp.
//                       Used for the LSTM input projections (M = snippets*46, N = 1024) and Dense(128).
//   lstm_rec_tc_kernel  the recurrence: per CTA R snippets x one direction; W_hh (128 x 512, fp16, 128 KB) stays in shared
//                       memory, h_{t-1} is a 128 x 128 fp16 A operand rewritten every step, the 4 x 128 gate pre-activations
//                       of a step are exactly the 512 TMEM columns; workers add the projected input (fp32, from HBM/L2), apply
//                       the gates with fp32 state in registers and hand h_t back through shared memory.
#include <algorithm>
#include <cstring>

#include <atomic>

#include "common.h"
#include "net.h"
#include "tc_common.cuh"

namespace orcai {

namespace {

using namespace tc;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred)
      :
      : "memory");
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// contiguous global -> shared copy through the TMA unit, completion counted on an mbarrier
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint4 pack8h(const float* v) {
  uint4 r;
  __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  return r;
}

// ------------------------------------------------------------------------------------------------
// GEMM
// ------------------------------------------------------------------------------------------------
constexpr int kGM = 128, kGK = 64;   // rows per CTA, K per pipeline stage

// SPLIT: both operands as (hi, lo) fp16 pairs, three MMAs per K step (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo): fp32-grade products
// on the fp16 tensor cores.  A stage = [A_hi | A_lo], B block = [B_hi | B_lo].
template <int BN, bool SPLIT = false>
struct GemmSmem {
  static constexpr uint32_t A_HALF = kGM * kGK * 2, B_HALF = BN * kGK * 2;
  static constexpr uint32_t A_STAGE = A_HALF * (SPLIT ? 2 : 1), B_STAGE = B_HALF * (SPLIT ? 2 : 1);
  static constexpr uint32_t OFF_A = 0, OFF_B = 2 * A_STAGE, OFF_BAR = OFF_B + 2 * B_STAGE;
  static constexpr uint32_t BYTES = OFF_BAR + 8 * 8 + 16;
  static constexpr int CTAS = BYTES <= 110 * 1024 ? 2 : 1;
};

// one K-chunk (64 elements) of an A row as fp16, 8 x 16 bytes; `valid` = readable elements from p (multiple of 4 / 8)
__device__ __forceinline__ void stage_row(const float* p, int valid, uint4 (&out)[kGK / 8]) {
  float4 v[kGK / 4];
#pragma unroll
  for (int q = 0; q < kGK / 4; ++q) v[q] = 4 * q + 4 <= valid ? __ldg(reinterpret_cast<const float4*>(p) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < kGK / 8; ++j) {
    const float f[8] = {v[2 * j].x, v[2 * j].y, v[2 * j].z, v[2 * j].w, v[2 * j + 1].x, v[2 * j + 1].y, v[2 * j + 1].z, v[2 * j + 1].w};
    out[j] = pack8h(f);
  }
}
// the same chunk as (hi, lo) pairs
__device__ __forceinline__ void stage_row_split(const float* p, int valid, uint4 (&hi)[kGK / 8], uint4 (&lo)[kGK / 8]) {
  float4 v[kGK / 4];
#pragma unroll
  for (int q = 0; q < kGK / 4; ++q) v[q] = 4 * q + 4 <= valid ? __ldg(reinterpret_cast<const float4*>(p) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < kGK / 8; ++j) {
    const float f[8] = {v[2 * j].x, v[2 * j].y, v[2 * j].z, v[2 * j].w, v[2 * j + 1].x, v[2 * j + 1].y, v[2 * j + 1].z, v[2 * j + 1].w};
    hi[j] = pack8h(f);
    const __half2* h = reinterpret_cast<const __half2*>(&hi[j]);
    float r[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 g = __half22float2(h[i]);
      r[2 * i] = f[2 * i] - g.x;
      r[2 * i + 1] = f[2 * i + 1] - g.y;
    }
    lo[j] = pack8h(r);
  }
}
__device__ __forceinline__ void stage_row(const __half* p, int valid, uint4 (&out)[kGK / 8]) {
#pragma unroll
  for (int j = 0; j < kGK / 8; ++j) out[j] = 8 * j + 8 <= valid ? __ldg(reinterpret_cast<const uint4*>(p) + j) : make_uint4(0, 0, 0, 0);
}

// ACT: 0 none, 1 relu.  grid = (ceil(M / 128), N / BN), block = 160 (4 worker warps + issuer warp).  Columns >= n_valid
// (a multiple of 4) of a row are not stored.
template <typename TA, int BN, int ACT, bool SPLIT = false>
__global__ void __launch_bounds__(160, (GemmSmem<BN, SPLIT>::CTAS))
gemm_tc_kernel(const TA* __restrict__ A, int lda, const __half* __restrict__ Bp, const float* __restrict__ bias,
               float* __restrict__ C, int ldc, long long M, int K, int n_valid) {
  using S = GemmSmem<BN, SPLIT>;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);   // full_a[2] full_b[2] free[2] acc
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 7);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkc = (K + kGK - 1) / kGK;
  if (tid == 0) {
    mbar_init(&bars[0], 4); mbar_init(&bars[1], 4);
    mbar_init(&bars[2], 1); mbar_init(&bars[3], 1);
    mbar_init(&bars[4], 1); mbar_init(&bars[5], 1);
    mbar_init(&bars[6], 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc<BN>(tslot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t sbase = smem_u32(smem);
  const long long row0 = (long long)blockIdx.x * kGM;   // row tiles along x: their count exceeds the 65 535 limit of grid.y
  const int nt = blockIdx.y;

  if (warp == 4) {
    constexpr uint32_t idesc = make_idesc_f16(128, BN, 0);
    for (int kc = 0; kc < nkc; ++kc) {
      const int s = kc & 1;
      const uint32_t par = (uint32_t)((kc >> 1) & 1);
      mbar_wait(&bars[0 + s], par);
      mbar_wait(&bars[2 + s], par);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = make_smem_desc(sbase + S::OFF_A + s * S::A_STAGE, kGM * 16, 128);
        const uint64_t db = make_smem_desc(sbase + S::OFF_B + s * S::B_STAGE, 128, (kGK / 8) * 128);
#pragma unroll
        for (int ks = 0; ks < kGK / 16; ++ks) {
          mma_f16_ss(tmem, da + ((2 * ks * kGM * 16) >> 4), db + ((2 * ks * 128) >> 4), idesc, (kc | ks) != 0);
          if constexpr (SPLIT) {
            mma_f16_ss(tmem, da + ((S::A_HALF + 2 * ks * kGM * 16) >> 4), db + ((2 * ks * 128) >> 4), idesc, 1);
            mma_f16_ss(tmem, da + ((2 * ks * kGM * 16) >> 4), db + ((S::B_HALF + 2 * ks * 128) >> 4), idesc, 1);
          }
        }
        mma_commit(&bars[4 + s]);
        if (kc == nkc - 1) mma_commit(&bars[6]);
      }
      __syncwarp();
    }
  } else {
    const int row = tid;                        // 0..127: A row staged by this thread and accumulator row it drains
    const bool row_ok = row0 + row < M;
    const TA* arow = A + (size_t)(row_ok ? row0 + row : 0) * lda;
    for (int kc = 0; kc < nkc; ++kc) {
      const int s = kc & 1;
      if (kc >= 2) mbar_wait(&bars[4 + s], (uint32_t)(((kc - 2) >> 1) & 1));   // the MMAs that read this stage are done
      if (tid == 0) {
        mbar_arrive_expect_tx(&bars[2 + s], S::B_STAGE);
        bulk_copy_g2s(sbase + S::OFF_B + s * S::B_STAGE, Bp + ((size_t)nt * nkc + kc) * (S::B_STAGE / 2), S::B_STAGE, &bars[2 + s]);
      }
      unsigned char* dst = smem + S::OFF_A + s * S::A_STAGE + row * 16;
      // all loads of the chunk are issued before the first conversion (one memory round trip per chunk, not eight)
      uint4 staged[kGK / 8];
      if constexpr (SPLIT) {
        uint4 staged_lo[kGK / 8];
        stage_row_split(arow + kc * kGK, row_ok ? K - kc * kGK : 0, staged, staged_lo);
#pragma unroll
        for (int j = 0; j < kGK / 8; ++j) *reinterpret_cast<uint4*>(dst + S::A_HALF + j * (kGM * 16)) = staged_lo[j];
      } else {
        stage_row(arow + kc * kGK, row_ok ? K - kc * kGK : 0, staged);
      }
#pragma unroll
      for (int j = 0; j < kGK / 8; ++j) *reinterpret_cast<uint4*>(dst + j * (kGM * 16)) = staged[j];
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[0 + s]);
    }
    mbar_wait(&bars[6], 0);
    tc_fence_after();
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    float* crow = C + (size_t)(row0 + row) * ldc + (size_t)nt * BN;
    const bool wide = (ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(C) & 31) == 0;   // rows start on 32-byte boundaries
#pragma unroll 2
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      tmem_ld16(lane_addr + c0, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] += __ldg(bias + nt * BN + c0 + i);
        if (ACT == 1) v[i] = fmaxf(v[i], 0.f);
      }
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < 4; q += 2) {
          if (wide && nt * BN + c0 + 4 * q + 8 <= n_valid) st_global_v8(crow + c0 + 4 * q, v + 4 * q);
          else {
#pragma unroll
            for (int qq = q; qq < q + 2; ++qq)
              if (nt * BN + c0 + 4 * qq + 4 <= n_valid) reinterpret_cast<float4*>(crow + c0)[qq] = make_float4(v[4 * qq], v[4 * qq + 1], v[4 * qq + 2], v[4 * qq + 3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<BN>(tmem);
}

// ------------------------------------------------------------------------------------------------
// LSTM recurrence
// ------------------------------------------------------------------------------------------------
constexpr int kU = 128, kG4 = 4 * kU;
constexpr uint32_t kRecW = kG4 * kU * 2;             // 128 KB: W_hh^T as [n = gate*128 + unit][k] canonical, SBO = 16 chunks
constexpr uint32_t kRecH = 128 * kU * 2;             // 32 KB: h as [k-chunk][row][8]
constexpr uint32_t kRecSmem = kRecW + kRecH + 4 * 8 + 16;

// ex2.approx / rcp.approx directly (2 ulp / 1 ulp): sigmoid = 4 instructions, tanh = 5
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.f, rcp_approx(1.f + ex2_approx(-2.8853900817779268f * x)), -1.f); }

__device__ __forceinline__ void tmem_ld4f(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// grid = (ceil(n / 64), 2 directions), block = 32 warps.  A warp can only read the TMEM lane quadrant (warp id % 4), and a
// CTA owns 64 snippets = quadrants 0 and 1, so the 16 warps with (id % 4) < 2 are the workers (quad = id % 4 -> rows,
// cg = id / 4 -> 16 hidden units each), warp 3 issues the MMAs, the rest idle.  Sixteen active warps instead of eight halve
// the per-step latency of the gate math, which is what bounds this kernel (10 MUFU ops per unit and step).
constexpr int kRecRows = 64;
__global__ void __launch_bounds__(1024, 1)
lstm_rec_tc_kernel(const float* __restrict__ xz, const __half* __restrict__ whh_pack, __half* __restrict__ hout, long long n, int Tn) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* s_h = smem + kRecW;   // W_hh occupies [0, kRecW)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRecW + kRecH);   // w_full, z_ready, h_ready
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;
  const long long s0 = (long long)blockIdx.x * kRecRows;
  constexpr int kActive = 16;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], kActive);
    fence_mbar_init();
  }
  for (int i = tid; i < (int)(kRecH / 16); i += 1024) reinterpret_cast<uint4*>(s_h)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  if (warp == 0) tmem_alloc<512>(tslot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t sbase = smem_u32(smem);

  if (warp == 3) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars[0], kRecW);
      for (int i = 0; i < 8; ++i)
        bulk_copy_g2s(sbase + i * (kRecW / 8), reinterpret_cast<const unsigned char*>(whh_pack) + (size_t)dir * kRecW + (size_t)i * (kRecW / 8),
                      kRecW / 8, &bars[0]);
    }
    __syncwarp();
    mbar_wait(&bars[0], 0);
    constexpr uint32_t idesc = make_idesc_f16(128, 256, 0);
    const uint64_t dh = make_smem_desc(sbase + kRecW, 128 * 16, 128);
    const uint64_t dw = make_smem_desc(sbase, 128, (kU / 8) * 128);
    for (int ti = 1; ti < Tn; ++ti) {            // step 0 has h = 0: no recurrent term
      mbar_wait(&bars[2], (uint32_t)((ti - 1) & 1));
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int nh = 0; nh < 2; ++nh)
#pragma unroll
          for (int ks = 0; ks < kU / 16; ++ks)
            mma_f16_ss(tmem + nh * 256, dh + ((2 * ks * 128 * 16) >> 4), dw + ((nh * 32 * (kU / 8) * 128 + 2 * ks * 128) >> 4), idesc, ks != 0);
        mma_commit(&bars[1]);
      }
      __syncwarp();
    }
  } else if ((warp & 3) < 2) {
    const int quad = warp & 3, cg = warp >> 2;     // cg 0..7
    const int row = quad * 32 + lane;
    const long long s = s0 + row;
    const bool ok = s < n;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const int u0 = cg * 16;
    float cst[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) cst[i] = 0.f;
    for (int ti = 0; ti < Tn; ++ti) {
      const int t = dir ? Tn - 1 - ti : ti;
      const float* xrow = xz + ((size_t)(ok ? s : 0) * Tn + t) * (2 * kG4) + (size_t)dir * kG4 + u0;
      __half* hrow = hout + ((size_t)(ok ? s : 0) * Tn + t) * (2 * kU) + (size_t)dir * kU + u0;
      // pull the projected inputs of the step after next into L2 (they stream from HBM: 345 MB per layer do not stay resident)
      if (ti + 2 < Tn) {
        const int t2 = dir ? Tn - 3 - ti : ti + 2;
        const float* xn = xz + ((size_t)(ok ? s : 0) * Tn + t2) * (2 * kG4) + (size_t)dir * kG4 + u0;
#pragma unroll
        for (int gte = 0; gte < 4; ++gte) asm volatile("prefetch.global.L2 [%0];" ::"l"(xn + gte * kU));
      }
      // the projected input of the first four units is fetched before waiting for the tensor pipe
      float4 xa[4];
#pragma unroll
      for (int gte = 0; gte < 4; ++gte) xa[gte] = __ldg(reinterpret_cast<const float4*>(xrow + gte * kU));
      if (ti > 0) {
        mbar_wait(&bars[1], (uint32_t)((ti - 1) & 1));
        tc_fence_after();
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 xb[4];
        if (j < 3) {
#pragma unroll
          for (int gte = 0; gte < 4; ++gte) xb[gte] = __ldg(reinterpret_cast<const float4*>(xrow + gte * kU + 4 * (j + 1)));
        }
        float z[4][4];
#pragma unroll
        for (int gte = 0; gte < 4; ++gte) {
          if (ti > 0) {
            tmem_ld4f(lane_addr + gte * kU + u0 + 4 * j, z[gte]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) z[gte][i] = 0.f;
          }
          z[gte][0] += xa[gte].x; z[gte][1] += xa[gte].y; z[gte][2] += xa[gte].z; z[gte][3] += xa[gte].w;
        }
        float h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float ig = sigmoid_fast(z[0][i]), fg = sigmoid_fast(z[1][i]), gg = tanh_fast(z[2][i]), og = sigmoid_fast(z[3][i]);
          const float cn = fg * cst[4 * j + i] + ig * gg;
          cst[4 * j + i] = cn;
          h[i] = og * tanh_fast(cn);
        }
        uint2 hp;
        *reinterpret_cast<__half2*>(&hp.x) = __floats2half2_rn(h[0], h[1]);
        *reinterpret_cast<__half2*>(&hp.y) = __floats2half2_rn(h[2], h[3]);
        const int u = u0 + 4 * j;
        *reinterpret_cast<uint2*>(s_h + (size_t)(u >> 3) * (128 * 16) + row * 16 + (u & 4) * 2) = hp;
        if (ok) *reinterpret_cast<uint2*>(hrow + 4 * j) = hp;
        if (j < 3) {
#pragma unroll
          for (int gte = 0; gte < 4; ++gte) xa[gte] = xb[gte];
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0 && ti + 1 < Tn) mbar_arrive(&bars[2]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}


// ------------------------------------------------------------------------------------------------
// The recurrence at fp32 grade on the tensor cores (net_path 4): every product runs as h_hi*W_hi + h_lo*W_hi + h_hi*W_lo.
// W_hh as (hi, lo) is 256 KB per direction - more than one SM's shared memory - so a CLUSTER OF TWO CTAs owns 64 snippets x one
// direction: CTA r keeps the recurrent weights of hidden units [64 r, 64 r + 64) (all four gates: 256 B-operand rows, hi + lo =
// 128 KB) and computes those 256 gate columns (its TMEM) for the 64 snippets; both CTAs hold all of h_{t-1} as the (hi, lo) A
// operand (64 KB), and the workers of CTA r write the h_t of their units into BOTH copies, the partner's through distributed
// shared memory.  Hand-offs per step, no cluster-wide barrier:
//     z_ready  (2 arrivals)   tcgen05.commit of BOTH CTAs' step (multicast): the accumulator is complete AND neither tensor pipe
//                             reads h_{t-1} any more, so either copy may be overwritten
//     h_ready  (32 arrivals)  the 16 worker warps of either CTA have written h_t (the remote ones arrive through the cluster)
// 24 MMAs (M 128, N 256, K 16) per step and CTA; the step is bound by the gate math and the hand-off latencies.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kSrW = 256 * kU * 2;              // 64 KB: [n = gate*64 + unit][k] canonical, hi; the lo copy follows
constexpr uint32_t kSrH = 128 * kU * 2;              // 32 KB: h as [k-chunk][row][8], hi; the lo copy follows
constexpr uint32_t kSrSmem = 2 * kSrW + 2 * kSrH + 4 * 8 + 16;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, uint2 v) {
  asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {   // release at cluster scope
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// commit to the same barrier of both CTAs of the pair
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(1024, 1)
lstm_rec_split_kernel(const float* __restrict__ xz, const __half* __restrict__ whh_hi, const __half* __restrict__ whh_lo, float* __restrict__ hout,
                      long long n, int Tn) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* s_h = smem + 2 * kSrW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kSrW + 2 * kSrH);   // w_full, z_ready, h_ready
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;
  const uint32_t r = cluster_ctarank();
  const long long s0 = (long long)(blockIdx.x >> 1) * kRecRows;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 2);          // both CTAs' commits
    mbar_init(&bars[2], 32);         // 16 worker warps of either CTA
    fence_mbar_init();
  }
  for (int i = tid; i < (int)(2 * kSrH / 16); i += 1024) reinterpret_cast<uint4*>(s_h)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  if (warp == 0) tmem_alloc<256>(tslot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                // the partner's barriers and h buffers exist before anything is sent there
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t sbase = smem_u32(smem);

  if (warp == 3) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars[0], 2 * kSrW);
      for (int g = 0; g < 4; ++g) {   // this CTA's units of gate g: 8 row groups of 2 KB out of the gate's 16
        const size_t src = (size_t)dir * kRecW + (size_t)(g * 16 + 8 * r) * 2048;
        bulk_copy_g2s(sbase + g * 16384, reinterpret_cast<const unsigned char*>(whh_hi) + src, 16384, &bars[0]);
        bulk_copy_g2s(sbase + kSrW + g * 16384, reinterpret_cast<const unsigned char*>(whh_lo) + src, 16384, &bars[0]);
      }
    }
    __syncwarp();
    mbar_wait(&bars[0], 0);
    constexpr uint32_t idesc = make_idesc_f16(128, 256, 0);
    const uint64_t dh = make_smem_desc(sbase + 2 * kSrW, 128 * 16, 128);
    const uint64_t dw = make_smem_desc(sbase, 128, (kU / 8) * 128);
    for (int ti = 1; ti < Tn; ++ti) {            // step 0 has h = 0: no recurrent term
      mbar_wait_cluster(&bars[2], (uint32_t)((ti - 1) & 1));
      fence_proxy_async();                       // the partner's writes into this CTA's h (generic proxy) before the MMAs read it
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kU / 16; ++ks) {
          const uint64_t a = dh + ((2 * ks * 128 * 16) >> 4), b = dw + ((2 * ks * 128) >> 4);
          mma_f16_ss(tmem, a, b, idesc, ks != 0);
          mma_f16_ss(tmem, a + (kSrH >> 4), b, idesc, 1);
          mma_f16_ss(tmem, a, b + (kSrW >> 4), idesc, 1);
        }
        mma_commit_pair(&bars[1]);
      }
      __syncwarp();
    }
  } else if ((warp & 3) < 2) {
    const int quad = warp & 3, cg = warp >> 2;     // cg 0..7: 8 of this CTA's 64 hidden units each
    const int row = quad * 32 + lane;
    const long long s = s0 + row;
    const bool ok = s < n;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const int u0 = cg * 8, ug0 = 64 * (int)r + u0;   // first unit: column inside a gate's 64 columns / hidden unit
    const uint32_t sh_local = smem_u32(s_h), sh_peer = map_to_cta(sh_local, r ^ 1u);
    const uint32_t hbar_local = smem_u32(&bars[2]), hbar_peer = map_to_cta(hbar_local, r ^ 1u);
    float cst[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cst[i] = 0.f;
    for (int ti = 0; ti < Tn; ++ti) {
      const int t = dir ? Tn - 1 - ti : ti;
      const float* xrow = xz + ((size_t)(ok ? s : 0) * Tn + t) * (2 * kG4) + (size_t)dir * kG4 + ug0;
      float* hrow = hout + ((size_t)(ok ? s : 0) * Tn + t) * (2 * kU) + (size_t)dir * kU + ug0;
      if (ti + 2 < Tn) {   // the projected inputs of the step after next: into L2 (they stream from HBM)
        const int t2 = dir ? Tn - 3 - ti : ti + 2;
        const float* xn = xz + ((size_t)(ok ? s : 0) * Tn + t2) * (2 * kG4) + (size_t)dir * kG4 + ug0;
#pragma unroll
        for (int gte = 0; gte < 4; ++gte) asm volatile("prefetch.global.L2 [%0];" ::"l"(xn + gte * kU));
      }
      float4 xa[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int gte = 0; gte < 4; ++gte) xa[j][gte] = __ldg(reinterpret_cast<const float4*>(xrow + gte * kU + 4 * j));
      if (ti > 0) {
        mbar_wait(&bars[1], (uint32_t)((ti - 1) & 1));
        tc_fence_after();
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float z[4][4];
#pragma unroll
        for (int gte = 0; gte < 4; ++gte) {
          if (ti > 0) {
            tmem_ld4f(lane_addr + gte * 64 + u0 + 4 * j, z[gte]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) z[gte][i] = 0.f;
          }
          z[gte][0] += xa[j][gte].x; z[gte][1] += xa[j][gte].y; z[gte][2] += xa[j][gte].z; z[gte][3] += xa[j][gte].w;
        }
        float h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float ig = sigmoid_fast(z[0][i]), fg = sigmoid_fast(z[1][i]), gg = tanh_fast(z[2][i]), og = sigmoid_fast(z[3][i]);
          const float cn = fg * cst[4 * j + i] + ig * gg;
          cst[4 * j + i] = cn;
          h[i] = og * tanh_fast(cn);
        }
        const __half2 h01 = __floats2half2_rn(h[0], h[1]), h23 = __floats2half2_rn(h[2], h[3]);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(h[0] - f01.x, h[1] - f01.y), l23 = __floats2half2_rn(h[2] - f23.x, h[3] - f23.y);
        uint2 hp, lp;
        hp.x = *reinterpret_cast<const uint32_t*>(&h01); hp.y = *reinterpret_cast<const uint32_t*>(&h23);
        lp.x = *reinterpret_cast<const uint32_t*>(&l01); lp.y = *reinterpret_cast<const uint32_t*>(&l23);
        const int u = ug0 + 4 * j;
        const uint32_t off = (uint32_t)(u >> 3) * (128 * 16) + (uint32_t)row * 16 + (uint32_t)(u & 4) * 2;
        *reinterpret_cast<uint2*>(s_h + off) = hp;
        *reinterpret_cast<uint2*>(s_h + kSrH + off) = lp;
        st_cluster_v2(sh_peer + off, hp);
        st_cluster_v2(sh_peer + kSrH + off, lp);
        if (ok) *reinterpret_cast<float4*>(hrow + 4 * j) = make_float4(h[0], h[1], h[2], h[3]);
      }
      if (ti + 1 < Tn) {
        fence_proxy_async();
        tc_fence_before();
        asm volatile("fence.acq_rel.cluster;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(hbar_local);
          mbar_arrive_cluster(hbar_peer);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                // nobody leaves while the partner may still send
  if (warp == 0) tmem_dealloc<256>(tmem);
}

// Dense(7) + sigmoid on the Dense(128)+ReLU activations (BatchNorm folded into the weights): one warp per row
__global__ void __launch_bounds__(256)
dense_out_kernel(const float* __restrict__ d1, const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ out,
                 long long rows, int L) {
  __shared__ float s_w[128 * 8];
  for (int i = threadIdx.x; i < 128 * L; i += blockDim.x) s_w[i] = w2[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * 8) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(d1 + (size_t)r * 128) + lane);
    for (int o = 0; o < L; ++o) {
      float a = x.x * s_w[(4 * lane) * L + o] + x.y * s_w[(4 * lane + 1) * L + o] + x.z * s_w[(4 * lane + 2) * L + o] + x.w * s_w[(4 * lane + 3) * L + o];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
      if (lane == 0) out[(size_t)r * L + o] = 1.f / (1.f + expf(-(a + b2[o])));
    }
  }
}

// ---- host side ----------------------------------------------------------------------------------
// B operand blocks of gemm_tc_kernel: [n-tile][k-chunk][BN x 64] canonical K-major (SBO = 8 chunks)
// split: every block is [hi | lo] (GemmSmem<BN, true>).  w has n_src columns (ldw apart); columns n_src .. N-1 of the operand are zero
std::vector<__half> pack_gemm_b(const float* w, int K, int N, int BN, bool split = false, int n_src = -1, int ldw = -1) {   // w: [K][N] row-major (Keras kernel layout)
  if (n_src < 0) n_src = N;
  if (ldw < 0) ldw = n_src;
  const int nkc = (K + kGK - 1) / kGK, ntiles = N / BN;
  const size_t half_blk = (size_t)BN * kGK, blk_sz = half_blk * (split ? 2 : 1);
  std::vector<__half> out((size_t)ntiles * nkc * blk_sz, __float2half_rn(0.f));
  for (int nt = 0; nt < ntiles; ++nt)
    for (int kc = 0; kc < nkc; ++kc) {
      __half* blk = out.data() + ((size_t)nt * nkc + kc) * blk_sz;
      for (int nn = 0; nn < BN; ++nn)
        for (int kk = 0; kk < kGK; ++kk) {
          const int k = kc * kGK + kk, n = nt * BN + nn;
          if (k >= K || n >= n_src) continue;
          const size_t at = ((size_t)(nn / 8) * (kGK / 8) * 128 + (size_t)(kk / 8) * 128 + (nn % 8) * 16 + (kk % 8) * 2) / 2;
          const float v = w[(size_t)k * ldw + n];
          const __half hi = __float2half_rn(v);
          blk[at] = hi;
          if (split) blk[half_blk + at] = __float2half_rn(v - __half2float(hi));
        }
    }
  return out;
}

int upload_half(Ctx* c, const std::vector<__half>& v, __half** out) {
  void* p = nullptr;
  ORCAI_CUDA(c, cudaMalloc(&p, v.size() * sizeof(__half)));
  c->net->allocs.push_back(p);
  ORCAI_CUDA(c, cudaMemcpy(p, v.data(), v.size() * sizeof(__half), cudaMemcpyHostToDevice));
  *out = static_cast<__half*>(p);
  return ORCAI_OK;
}

template <typename TA, int BN, int ACT, bool SPLIT = false>
int run_gemm_tc(Ctx* c, const TA* A, int lda, const __half* Bp, const float* bias, float* C, int ldc, long long M, int N, int K, int n_valid = -1) {
  using S = GemmSmem<BN, SPLIT>;
  // the attribute is per device: remember which devices have it (several contexts can live in one process)
  static std::atomic<unsigned long long> attr_devices{0ull};
  if (!((attr_devices.load() >> (c->device & 63)) & 1ull)) {
    ORCAI_CUDA(c, cudaFuncSetAttribute(gemm_tc_kernel<TA, BN, ACT, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::BYTES));
    attr_devices.fetch_or(1ull << (c->device & 63));
  }
  if (M <= 0) return ORCAI_OK;
  dim3 grid((unsigned)((M + kGM - 1) / kGM), (unsigned)(N / BN));
  gemm_tc_kernel<TA, BN, ACT, SPLIT><<<grid, 160, S::BYTES, c->stream>>>(A, lda, Bp, bias, C, ldc, M, K, n_valid < 0 ? N : n_valid);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

}  // namespace

int net_tail_tc_prepare(Ctx* c) {
  NetWeights* nw = c->net;
  if (nw->tail_tc_ready) return ORCAI_OK;
  const int U = nw->U, G = 4 * U;
  if (U != kU) ORCAI_FAIL(c, ORCAI_ERR_ARG, "tensor-core LSTM kernels are built for 128 units");
  const Calib& cal = nw->calib;
  auto qerr = [](float v) { return (double)__half2float(__float2half_rn(v)) - (double)v; };
  for (int l = 0; l < 2; ++l) {
    const int I = l == 0 ? nw->feat : 2 * U;
    ORCAI_CHECK(upload_half(c, pack_gemm_b(nw->h_lstm_wih[l].data(), I, 2 * G, 256), &nw->tc_wih[l]));
    // W_hh^T per direction: B operand rows n = gate*128 + unit, K = previous hidden unit ; recurrent_kernel is [U][4U]
    std::vector<__half> wp((size_t)2 * G * U, __float2half_rn(0.f));
    for (int d = 0; d < 2; ++d)
      for (int n = 0; n < G; ++n)
        for (int k = 0; k < U; ++k)
          wp[(size_t)d * G * U + ((size_t)(n / 8) * (U / 8) * 128 + (size_t)(k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) / 2] =
              __float2half_rn(nw->h_lstm_whh[l][(size_t)d * U * G + (size_t)k * G + n]);
    ORCAI_CHECK(upload_half(c, wp, &nw->tc_whh[l]));
    // projection bias minus what fp16 rounding of W_ih and W_hh adds on average (Calib, net.h)
    std::vector<float> bih(nw->h_lstm_bih[l]);
    if (cal.valid) {
      const std::vector<double>& mu_in = l == 0 ? cal.feat : cal.h1;
      const std::vector<double>& mu_h = l == 0 ? cal.h1 : cal.h2;
      for (int n = 0; n < 2 * G; ++n) {
        double dsum = 0.0;
        for (int k = 0; k < I; ++k) dsum += qerr(nw->h_lstm_wih[l][(size_t)k * 2 * G + n]) * mu_in[k];
        const int d = n / G, nn = n % G;
        for (int k = 0; k < U; ++k) dsum += qerr(nw->h_lstm_whh[l][(size_t)d * U * G + (size_t)k * G + nn]) * mu_h[(size_t)d * U + k];
        bih[n] = (float)((double)bih[n] - dsum);
      }
    }
    ORCAI_CHECK(net_upload(c, bih, &nw->tc_bih[l]));
  }
  ORCAI_CHECK(upload_half(c, pack_gemm_b(nw->h_d1_w.data(), 2 * U, 128, 128), &nw->tc_d1));
  {
    std::vector<float> b1(nw->h_d1_b);
    if (cal.valid)
      for (int n = 0; n < 128; ++n) {
        double dsum = 0.0;
        for (int k = 0; k < 2 * U; ++k) dsum += qerr(nw->h_d1_w[(size_t)k * 128 + n]) * cal.h2[k];
        b1[n] = (float)((double)b1[n] - dsum);
      }
    ORCAI_CHECK(net_upload(c, b1, &nw->tc_d1_b));
  }
  ORCAI_CUDA(c, cudaFuncSetAttribute(lstm_rec_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRecSmem));
  nw->tail_tc_ready = true;
  return ORCAI_OK;
}

// ---- fp32-grade products on the fp16 tensor cores (net_path 4) -----------------------------------
// device copy of the packed split B operand (BN = 64) of a [K][n_src] row-major weight matrix, N columns (multiple of 64)
int net_pack_split_b(Ctx* c, const float* w, int K, int n_src, int N, __half** out) {
  return upload_half(c, pack_gemm_b(w, K, N, 64, true, n_src, n_src), out);
}

// C[M, ldc] = act(A[M, lda] * W + bias), act 0 none / 1 relu; columns >= n_valid are not stored; bias holds N floats
int net_gemm_split(Ctx* c, const float* A, int lda, const __half* Bp, const float* bias, float* C, int ldc, long long M, int N, int K, int n_valid,
                   int act) {
  if (act == 1) return run_gemm_tc<float, 64, 1, true>(c, A, lda, Bp, bias, C, ldc, M, N, K, n_valid);
  return run_gemm_tc<float, 64, 0, true>(c, A, lda, Bp, bias, C, ldc, M, N, K, n_valid);
}

int net_tail_precise_prepare(Ctx* c) {
  NetWeights* nw = c->net;
  if (nw->tail_precise_ready) return ORCAI_OK;
  const int U = nw->U, G = 4 * U;
  for (int l = 0; l < 2; ++l) {
    const int I = l == 0 ? nw->feat : 2 * U;
    // column blocks of 256: every CTA stages (splits) its 128 A rows once per block, so 4 blocks instead of 16 quarter that work
    ORCAI_CHECK(upload_half(c, pack_gemm_b(nw->h_lstm_wih[l].data(), I, 2 * G, 256, true), &nw->tp_wih[l]));
  }
  ORCAI_CHECK(net_pack_split_b(c, nw->h_d1_w.data(), 2 * U, 128, 128, &nw->tp_d1));
  if (U == kU) {
    // W_hh^T per direction as (hi, lo): B operand rows n = gate*128 + unit, K = previous hidden unit; recurrent_kernel is [U][4U]
    for (int l = 0; l < 2; ++l) {
      std::vector<__half> whi((size_t)2 * G * U), wlo((size_t)2 * G * U);
      for (int d = 0; d < 2; ++d)
        for (int n = 0; n < G; ++n)
          for (int k = 0; k < U; ++k) {
            const float w = nw->h_lstm_whh[l][(size_t)d * U * G + (size_t)k * G + n];
            const __half hi = __float2half_rn(w);
            const size_t at = (size_t)d * G * U + ((size_t)(n / 8) * (U / 8) * 128 + (size_t)(k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) / 2;
            whi[at] = hi;
            wlo[at] = __float2half_rn(w - __half2float(hi));
          }
      ORCAI_CHECK(upload_half(c, whi, &nw->tp_whh_hi[l]));
      ORCAI_CHECK(upload_half(c, wlo, &nw->tp_whh_lo[l]));
    }
    ORCAI_CUDA(c, cudaFuncSetAttribute(lstm_rec_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSrSmem));
  }
  nw->tail_precise_ready = true;
  return ORCAI_OK;
}

// LSTM x2 + dense head at fp32 grade: split-fp16 tensor-core GEMMs for the input projections and Dense(128), the fp32 CUDA-core
// recurrence of the reference-grade path (W_hh as (hi, lo) would need 256 KB of shared memory per CTA).  Scratch as net_tail_fp32.
int net_tail_precise(Ctx* c, const float* feat, float* scratch, long long m, float* d_preds_out, bool mk) {
  NetWeights* nw = c->net;
  ORCAI_CHECK(net_tail_precise_prepare(c));
  const int U = nw->U, G = 4 * U, L = nw->L;
  const int Tn = nw->H >> nw->n_blocks;
  const long long rows = m * Tn;
  float* xz = scratch;                          // (rows, 2G)
  float* h1 = xz + (size_t)rows * 2 * G;        // (rows, 2U)
  float* h2 = h1 + (size_t)rows * 2 * U;        // (rows, 2U)
  float* d1 = h2 + (size_t)rows * 2 * U;        // (rows, 128)
  ORCAI_CHECK((run_gemm_tc<float, 256, 0, true>(c, feat, nw->feat, nw->tp_wih[0], nw->lstm_bih[0], xz, 2 * G, rows, 2 * G, nw->feat, 2 * G)));
  net_mark(c, mk);  // 6: lstm1 input projection
  // the recurrence: split-fp16 tensor-core kernel on CTA pairs (lstm_rec 1, default) or the fp32 CUDA-core kernel of the reference-grade path
  auto recurrence = [&](int l, float* hout) -> int {
    if (U != kU || !nw->precise_lstm_tc) return net_lstm_rec_fp32(c, xz, nw->lstm_whh[l], hout, m, Tn);
    if (m <= 0) return ORCAI_OK;
    const dim3 rgrid((unsigned)(2 * ((m + kRecRows - 1) / kRecRows)), 2);
    lstm_rec_split_kernel<<<rgrid, 1024, kSrSmem, c->stream>>>(xz, nw->tp_whh_hi[l], nw->tp_whh_lo[l], hout, m, Tn);
    c->launches++;
    ORCAI_CUDA(c, cudaGetLastError());
    return ORCAI_OK;
  };
  ORCAI_CHECK(recurrence(0, h1));
  net_mark(c, mk);  // 7: lstm1 recurrence
  ORCAI_CHECK((run_gemm_tc<float, 256, 0, true>(c, h1, 2 * U, nw->tp_wih[1], nw->lstm_bih[1], xz, 2 * G, rows, 2 * G, 2 * U, 2 * G)));
  net_mark(c, mk);  // 8: lstm2 input projection
  ORCAI_CHECK(recurrence(1, h2));
  net_mark(c, mk);  // 9: lstm2 recurrence
  ORCAI_CHECK(net_gemm_split(c, h2, 2 * U, nw->tp_d1, nw->d1_b, d1, 128, rows, 128, 2 * U, 128, 1));
  {
    const long long grid = std::min<long long>((rows + 7) / 8, (long long)c->sm_count * 8);
    dense_out_kernel<<<(unsigned)grid, 256, 0, c->stream>>>(d1, nw->d2_w, nw->d2_b, d_preds_out, rows, L);
    c->launches++;
  }
  net_mark(c, mk);  // 10: dense head
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

// scratch: m*Tn*(2*4U + 2U + 2U + 128) floats (same budget as net_tail_fp32)
int net_tail_tc(Ctx* c, const float* feat, float* scratch, long long m, float* d_preds_out, bool mk) {
  NetWeights* nw = c->net;
  ORCAI_CHECK(net_tail_tc_prepare(c));
  const int U = nw->U, G = 4 * U, L = nw->L;
  const int Tn = nw->H >> nw->n_blocks;
  const long long rows = m * Tn;
  float* xz = scratch;                                               // (rows, 2G) fp32
  __half* h1 = reinterpret_cast<__half*>(xz + (size_t)rows * 2 * G);  // (rows, 2U) fp16
  __half* h2 = h1 + (size_t)rows * 2 * U;                             // (rows, 2U) fp16
  float* d1 = reinterpret_cast<float*>(h2 + (size_t)rows * 2 * U);    // (rows, 128) fp32
  const dim3 rgrid((unsigned)((m + kRecRows - 1) / kRecRows), 2);
  ORCAI_CHECK((run_gemm_tc<float, 256, 0>(c, feat, nw->feat, nw->tc_wih[0], nw->tc_bih[0], xz, 2 * G, rows, 2 * G, nw->feat)));
  net_mark(c, mk);  // 6: lstm1 input projection
  lstm_rec_tc_kernel<<<rgrid, 1024, kRecSmem, c->stream>>>(xz, nw->tc_whh[0], h1, m, Tn);
  c->launches++;
  net_mark(c, mk);  // 7: lstm1 recurrence
  ORCAI_CHECK((run_gemm_tc<__half, 256, 0>(c, h1, 2 * U, nw->tc_wih[1], nw->tc_bih[1], xz, 2 * G, rows, 2 * G, 2 * U)));
  net_mark(c, mk);  // 8: lstm2 input projection
  lstm_rec_tc_kernel<<<rgrid, 1024, kRecSmem, c->stream>>>(xz, nw->tc_whh[1], h2, m, Tn);
  c->launches++;
  net_mark(c, mk);  // 9: lstm2 recurrence
  ORCAI_CHECK((run_gemm_tc<__half, 128, 1>(c, h2, 2 * U, nw->tc_d1, nw->tc_d1_b, d1, 128, rows, 128, 2 * U)));
  {
    const long long grid = std::min<long long>((rows + 7) / 8, (long long)c->sm_count * 8);
    dense_out_kernel<<<(unsigned)grid, 256, 0, c->stream>>>(d1, nw->d2_w, nw->d2_b, d_preds_out, rows, L);
    c->launches++;
  }
  net_mark(c, mk);  // 10: dense head
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

}  // namespace orcai
