// One residual block of orcai-V1 as ONE persistent tcgen05 kernel (included by net_tc.cu).
//
// Reference graph: src/orcAI/architectures.py:172-196 —
//     x -> ReLU -> SepConv3x3 -> BN -> ReLU -> SepConv3x3 -> BN -> MaxPool(3,2)/2 "same"  (+)  Conv1x1/2(x)  -> y
//
// Data layout (global, fp16 NHWC, channel pitch padded to 8, padding channels zero):
//     Xr   (n, H,  W,  ICP)   ReLU(x)            — A operand of the first separable convolution
//     Xsub (n, Ho, Wo, ICP)   x at even (h, w)   — input of the residual 1x1/2 convolution (x before the ReLU)
//     Yr   (n, Ho, Wo, OCP)   ReLU(y) (or y)     — next block's Xr
//     Ysub (n, Ho/2, ceil(Wo/2), OCP)  y at even positions — next block's Xsub
//
// A CTA owns a strip of CP pooled columns of one snippet and marches down the time axis in steps of S rows.
// Shared memory holds three pixel-linear, channel-chunk-planar buffers with a common row pitch WP = 2*CP + 4:
//     X  [k-chunk][(S+2) x WP pixels][8 ch]   input rows a+1 .. a+S+2
//     S1 [k-chunk][(S+2) x WP pixels][8 ch]   first sepconv output rows a .. a+S+1   (rows 0,1 carried from the last step)
//     S2 [k-chunk][(S+1) x WP pixels][8 ch]   second sepconv output rows a .. a+S    (row 0 carried)
// Because pixels are linear with a fixed pitch, an M=128 MMA tile is simply 128 CONSECUTIVE pixels (SBO = 128 B,
// LBO = one chunk plane) and the A operand of tap (dy,dx) is the same buffer with the descriptor start shifted by
// (dy*WP + dx) pixels: no im2col.  Tiles run over the halo columns too; those accumulator rows are garbage and are
// never consumed (every accumulator row depends on its own A row only).  The depthwise filter is folded into the
// GEMM weights (W'_tap = dw[tap] * pw * bn_scale), the residual 1x1 convolution is one more small MMA into its own
// TMEM columns, and the max-pool + residual add + ReLU run in the epilogue straight out of shared memory / TMEM.
// HBM sees each block's input once (plus 4 halo columns per strip) and its pooled output once.
#pragma once

namespace fused {

__host__ __device__ constexpr int imax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int round8(int a) { return (a + 7) & ~7; }
__host__ __device__ constexpr int pow2cols(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

template <int CIN_, int COUT_, int CPOOL_, int S_, bool RELU_OUT_, int CTAS_>
struct FB {
  static constexpr int CIN = CIN_, COUT = COUT_, CP = CPOOL_, S = S_, CTAS = CTAS_;
  static constexpr bool RELU_OUT = RELU_OUT_;
  static constexpr int ICP = cpad8(CIN), OCP = cpad8(COUT);
  static constexpr int KP1 = cpad16(CIN);          // K per tap of sepconv 1 and of the residual convolution
  static constexpr int NP = cpad16(COUT);          // N of every MMA; K per tap of sepconv 2
  static constexpr int XG = ICP / 8;               // X / R chunk planes that carry data
  static constexpr int XCH = KP1 / 8;              // X / R chunk planes the MMA reads (the extra one stays zero)
  static constexpr int NG = OCP / 8;               // S1 / S2 chunk planes that carry data
  static constexpr int MCH = NP / 8;               // S1 chunk planes the MMA reads
  static constexpr int WP = 2 * CP + 4;
  static constexpr int N1 = (S * WP - 2 + 127) / 128, N2 = (S * WP - 4 + 127) / 128;
  static constexpr int P1_0 = 2 * WP + 1, P2_0 = WP + 2;
  static constexpr int XPIX = round8(imax((S + 2) * WP, P1_0 + 128 * N1 + 1));
  static constexpr int S1PIX = round8(imax((S + 2) * WP, P2_0 + 128 * N2 + WP + 1));
  static constexpr int S2PIX = round8((S + 1) * WP);
  static constexpr int RQ = (S / 2) * CP, RPIX = round8(RQ);
  static constexpr uint32_t LBO_X = XPIX * 16, LBO_S1 = S1PIX * 16, LBO_S2 = S2PIX * 16, LBO_R = RPIX * 16;
  // trimmed weight storage: only k-chunks / n-groups that carry data are stored; the MMA's reads of the missing
  // chunk alias the next group (finite values times a zero A chunk), missing n-groups only feed unused columns.
  static constexpr uint32_t SBO_W1 = XG * 128, SBO_W2 = NG * 128;
  static constexpr uint32_t TAP_W1 = NG * SBO_W1, TAP_W2 = NG * SBO_W2;
  static constexpr uint32_t W1_BYTES = 9 * TAP_W1 + 128, W2_BYTES = 9 * TAP_W2 + 128, WR_BYTES = TAP_W1 + 128;
  static constexpr uint32_t OFF_W1 = 0, OFF_W2 = OFF_W1 + W1_BYTES, OFF_WR = OFF_W2 + W2_BYTES;
  static constexpr uint32_t W_BYTES = OFF_WR + WR_BYTES;
  static constexpr uint32_t OFF_R = W_BYTES;
  static constexpr uint32_t OFF_X = OFF_R + XCH * LBO_R;
  static constexpr uint32_t OFF_S1 = OFF_X + XCH * LBO_X;
  static constexpr uint32_t OFF_S2 = OFF_S1 + MCH * LBO_S1;
  static constexpr uint32_t OFF_BIAS = OFF_S2 + NG * LBO_S2;
  static constexpr uint32_t OFF_BAR = OFF_BIAS + 3 * NP * 4;
  static constexpr uint32_t SMEM = OFF_BAR + (N1 + N2 + 1) * 8 + 16;
  static constexpr int COL_R = 0, COL_1 = NP, COL_2 = NP + N1 * NP;
  static constexpr int TM_COLS = pow2cols(NP * (1 + N1 + N2));
  static_assert(S % 2 == 0 && S >= 2, "steps advance by whole pooled rows");
  static_assert(RQ <= 128, "one residual MMA tile per step");
  static_assert(NP * (1 + N1 + N2) <= 512 && TM_COLS * CTAS <= 512, "TMEM columns");
  static_assert(SMEM <= 227 * 1024, "shared memory");
  static_assert((128 - RPIX) * 16 <= XCH * LBO_X, "residual tile over-read must stay inside the CTA's shared memory");
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// TMEM -> registers: 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint4 pack8h(const float (&v)[8]) {
  uint4 r;
  __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  return r;
}
__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
  __half2* x = reinterpret_cast<__half2*>(&a);
  const __half2* y = reinterpret_cast<const __half2*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = __hmax2(x[i], y[i]);
  return a;
}

template <class G>
__global__ void __launch_bounds__(256, G::CTAS)
fused_block_kernel(const __half* __restrict__ Xr, const __half* __restrict__ Xsub, __half* __restrict__ Yr,
                   __half* __restrict__ Ysub, int H, int W, int n_strips, long long n_items,
                   const unsigned char* __restrict__ wpack, const float* __restrict__ bias_pack) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* s_bias = reinterpret_cast<float*>(smem + G::OFF_BIAS);   // [0,NP) sep1, [NP,2NP) sep2, [2NP,3NP) residual
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G::OFF_BAR);  // [0,N1) sep1 tiles, [N1,N1+N2) sep2 tiles, last: residual
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + G::N1 + G::N2 + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int Ho = H >> 1, Wo = (W + 1) >> 1;
  const int Hs = Ho >> 1, Ws = (Wo + 1) >> 1;

  for (int i = tid; i < (int)(G::W_BYTES / 16); i += 256) reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(wpack) + i);
  for (int i = tid + G::W_BYTES / 16; i < (int)(G::OFF_BIAS / 16); i += 256) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 3 * G::NP; i += 256) s_bias[i] = bias_pack[i];
  if (tid == 0) {
    for (int i = 0; i < G::N1 + G::N2 + 1; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc<G::TM_COLS>(tslot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t sbase = smem_u32(smem);
  constexpr uint32_t idesc = make_idesc_f16(128, G::NP, 0);

  const int row = tid & 127;                     // accumulator row (TMEM lane) this thread drains
  const int half = tid >> 7;                     // the two thread halves split the channel groups
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  constexpr int G_SPLIT = (G::NG + 1) / 2;
  const int g_lo = half ? G_SPLIT : 0, g_hi = half ? G::NG : G_SPLIT;
  // pooled pixel of the pool / residual epilogue
  const int q_i = row / G::CP, q_j = row - q_i * G::CP;

  uint32_t phase = 0;
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    const long long b = item / n_strips;
    const int strip = (int)(item - b * n_strips);
    const int wo0 = strip * G::CP, cb = 2 * wo0 - 2;
    const __half* xr = Xr + (size_t)b * H * W * G::ICP;
    const __half* xs = Xsub + (size_t)b * Ho * Wo * G::ICP;
    const int n_steps = (Ho + 1 + G::S / 2 - 1) / (G::S / 2);

    // rows carried into the first step lie above the image: zero (S2 row 0 is only read by the skipped pooled row -1)
    for (int i = tid; i < G::NG * 2 * G::WP; i += 256) {
      const int g = i / (2 * G::WP), px = i - g * 2 * G::WP;
      *reinterpret_cast<uint4*>(smem + G::OFF_S1 + g * G::LBO_S1 + px * 16) = make_uint4(0, 0, 0, 0);
    }

    for (int step = 0; step < n_steps; ++step) {
      const int a = step * G::S - 2;
      // ---- loads: X rows a+1 .. a+S+2 and the residual input pixels of this step's pooled rows ----
      for (int idx = tid; idx < (G::S + 2) * G::WP * G::XG; idx += 256) {
        const int px = idx / G::XG, g = idx - px * G::XG;
        const int x = px / G::WP, c = px - x * G::WP;
        const int hh = a + 1 + x, ww = cb + c;
        const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
        const __half* src = ok ? xr + ((size_t)hh * W + ww) * G::ICP + g * 8 : xr;
        cp_async16(sbase + G::OFF_X + g * G::LBO_X + px * 16, src, ok);
      }
      for (int idx = tid; idx < G::RQ * G::XG; idx += 256) {
        const int q = idx / G::XG, g = idx - q * G::XG;
        const int i = q / G::CP, j = q - i * G::CP;
        const int ho = (a >> 1) + i, wo = wo0 + j;
        const bool ok = ho >= 0 && ho < Ho && wo < Wo;
        const __half* src = ok ? xs + ((size_t)ho * Wo + wo) * G::ICP + g * 8 : xs;
        cp_async16(sbase + G::OFF_R + g * G::LBO_R + q * 16, src, ok);
      }
      cp_async_wait_all();
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();

      // ---- residual 1x1 and first separable convolution ----
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < G::KP1 / 16; ++ks)
          mma_f16_ss(tmem + G::COL_R, make_smem_desc(sbase + G::OFF_R + 2 * ks * G::LBO_R, G::LBO_R, 128),
                     make_smem_desc(sbase + G::OFF_WR + 2 * ks * 128, 128, G::SBO_W1), idesc, ks != 0);
        mma_commit(&bars[G::N1 + G::N2]);
#pragma unroll 1
        for (int t = 0; t < G::N1; ++t) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t a0 = sbase + G::OFF_X + (uint32_t)(G::P1_0 + 128 * t - 2 * G::WP - 1 + (tap / 3) * G::WP + (tap % 3)) * 16;
            const uint32_t w0 = sbase + G::OFF_W1 + tap * G::TAP_W1;
#pragma unroll
            for (int ks = 0; ks < G::KP1 / 16; ++ks)
              mma_f16_ss(tmem + G::COL_1 + t * G::NP, make_smem_desc(a0 + 2 * ks * G::LBO_X, G::LBO_X, 128),
                         make_smem_desc(w0 + 2 * ks * 128, 128, G::SBO_W1), idesc, (tap | ks) != 0);
          }
          mma_commit(&bars[t]);
        }
      }
      // epilogue 1: + bias, ReLU, zero outside the image ("same" padding of the second convolution) -> S1
#pragma unroll 1
      for (int t = 0; t < G::N1; ++t) {
        mbar_wait(&bars[t], phase);
        tc_fence_after();
        const int p1 = G::P1_0 + 128 * t + row;
        const int y = p1 / G::WP, c = p1 - y * G::WP;
        const int hh = a + y, ww = cb + c;
        const bool inimg = hh >= 0 && hh < H && ww >= 0 && ww < W;
        const bool st = p1 < (G::S + 2) * G::WP;
        for (int g = g_lo; g < g_hi; ++g) {
          float v[8];
          tmem_ld8(lane_addr + G::COL_1 + t * G::NP + g * 8, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = inimg ? fmaxf(v[i] + s_bias[g * 8 + i], 0.f) : 0.f;
          if (st) *reinterpret_cast<uint4*>(smem + G::OFF_S1 + g * G::LBO_S1 + p1 * 16) = pack8h(v);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();

      // ---- second separable convolution ----
      if (tid == 0) {
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < G::N2; ++t) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t a0 = sbase + G::OFF_S1 + (uint32_t)(G::P2_0 + 128 * t - G::WP - 1 + (tap / 3) * G::WP + (tap % 3)) * 16;
            const uint32_t w0 = sbase + G::OFF_W2 + tap * G::TAP_W2;
#pragma unroll
            for (int ks = 0; ks < G::NP / 16; ++ks)
              mma_f16_ss(tmem + G::COL_2 + t * G::NP, make_smem_desc(a0 + 2 * ks * G::LBO_S1, G::LBO_S1, 128),
                         make_smem_desc(w0 + 2 * ks * 128, 128, G::SBO_W2), idesc, (tap | ks) != 0);
          }
          mma_commit(&bars[G::N1 + t]);
        }
      }
      // epilogue 2: + bias (folded BatchNorm), -inf outside the image (TF "same" max-pool padding) -> S2
#pragma unroll 1
      for (int t = 0; t < G::N2; ++t) {
        mbar_wait(&bars[G::N1 + t], phase);
        tc_fence_after();
        const int p2 = G::P2_0 + 128 * t + row;
        const int z = p2 / G::WP, c = p2 - z * G::WP;
        const int hh = a + z, ww = cb + c;
        const bool inimg = hh >= 0 && hh < H && ww >= 0 && ww < W;
        const bool st = p2 < (G::S + 1) * G::WP;
        for (int g = g_lo; g < g_hi; ++g) {
          float v[8];
          tmem_ld8(lane_addr + G::COL_2 + t * G::NP + g * 8, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = inimg ? v[i] + s_bias[G::NP + g * 8 + i] : -INFINITY;
          if (st) *reinterpret_cast<uint4*>(smem + G::OFF_S2 + g * G::LBO_S2 + p2 * 16) = pack8h(v);
        }
      }
      mbar_wait(&bars[G::N1 + G::N2], phase);   // residual accumulator
      tc_fence_after();
      __syncthreads();

      // ---- max-pool (3,2)/2 + residual add (+ ReLU) -> global ----
      {
        const int ho = (a >> 1) + q_i, wo = wo0 + q_j;
        const bool valid = row < G::RQ && ho >= 0 && ho < Ho && wo < Wo;
        const uint32_t p00 = (uint32_t)((2 * q_i) * G::WP + 2 + 2 * q_j);
        for (int g = g_lo; g < g_hi; ++g) {
          float r[8];
          tmem_ld8(lane_addr + G::COL_R + g * 8, r);
          if (valid) {
            const unsigned char* s2 = smem + G::OFF_S2 + g * G::LBO_S2 + p00 * 16;
            uint4 m = *reinterpret_cast<const uint4*>(s2);
            m = hmax8(m, *reinterpret_cast<const uint4*>(s2 + 16));
            m = hmax8(m, *reinterpret_cast<const uint4*>(s2 + G::WP * 16));
            m = hmax8(m, *reinterpret_cast<const uint4*>(s2 + G::WP * 16 + 16));
            m = hmax8(m, *reinterpret_cast<const uint4*>(s2 + 2 * G::WP * 16));
            m = hmax8(m, *reinterpret_cast<const uint4*>(s2 + 2 * G::WP * 16 + 16));
            const __half2* mh = reinterpret_cast<const __half2*>(&m);
            float y[8], yr[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __half22float2(mh[i]);
              y[2 * i] = f.x + r[2 * i] + s_bias[2 * G::NP + g * 8 + 2 * i];
              y[2 * i + 1] = f.y + r[2 * i + 1] + s_bias[2 * G::NP + g * 8 + 2 * i + 1];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) yr[i] = G::RELU_OUT ? fmaxf(y[i], 0.f) : y[i];
            *reinterpret_cast<uint4*>(Yr + (((size_t)b * Ho + ho) * Wo + wo) * G::OCP + g * 8) = pack8h(yr);
            if (Ysub != nullptr && !(ho & 1) && !(wo & 1))
              *reinterpret_cast<uint4*>(Ysub + (((size_t)b * Hs + (ho >> 1)) * Ws + (wo >> 1)) * G::OCP + g * 8) = pack8h(y);
          }
        }
      }
      tc_fence_before();
      __syncthreads();

      // ---- carry the overlap rows into the next step ----
      if (step + 1 < n_steps) {
        for (int i = tid; i < G::NG * 2 * G::WP; i += 256) {
          const int g = i / (2 * G::WP), px = i - g * 2 * G::WP;
          unsigned char* p = smem + G::OFF_S1 + g * G::LBO_S1 + px * 16;
          *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(p + G::S * G::WP * 16);
        }
        for (int i = tid; i < G::NG * G::WP; i += 256) {
          const int g = i / G::WP, px = i - g * G::WP;
          unsigned char* p = smem + G::OFF_S2 + g * G::LBO_S2 + px * 16;
          *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(p + G::S * G::WP * 16);
        }
      }
      phase ^= 1;
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<G::TM_COLS>(tmem);
}

}  // namespace fused
