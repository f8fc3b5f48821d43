// orcai-V1 convolutional trunk on the 5th-generation tensor cores (tcgen05 + TMEM), 16-bit activations.
//
// Same graph as net.cu (reference src/orcAI/architectures.py:120-241, called from predict.py:266-268).
// Every separable convolution  [ReLU] -> depthwise 3x3 -> pointwise 1x1 (+ folded BatchNorm) -> [ReLU]
// is executed as ONE implicit GEMM per 16x8-pixel tile:
//
//     D[128 pixels, Cout] = sum over taps (dy,dx)  A_tap[128, Cin] * W'_tap[Cin, Cout],
//     W'_tap[ci][co] = dw[tap][ci] * pw[ci][co] * bn_scale[co]
//
// The depthwise filter is folded into the GEMM weights, so no CUDA-core convolution remains.  The (18x10)-pixel
// halo tile of the input is copied ONCE into shared memory as [k-chunk of 8 channels][halo pixel][16 bytes];
// in that layout the A operand of every tap is the same buffer seen through a UMMA shared-memory descriptor
// whose start address is shifted by (dy*10+dx)*16 bytes (8-row core matrices = 8 consecutive pixels of a tile
// row, SBO = one halo row = 160 B, LBO = one k-chunk plane = 2880 B).  No im2col, no data movement per tap.
// One elected thread issues 9*Kpad/16 tcgen05.mma (M=128, N=Cout padded to 16, K=16) into a TMEM accumulator and
// commits to an mbarrier; the 4 warps then read their 32 accumulator rows with tcgen05.ld, add the bias, apply
// ReLU, convert and store NHWC with 16-byte stores.  CTAs are persistent (weights stay in shared memory).
//
// Activations between kernels are NHWC 16-bit with the channel pitch padded to a multiple of 8 (padding
// channels are written as zeros), so every (pixel, k-chunk) is one aligned 16-byte vector.
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include <atomic>

#include "common.h"
#include "net.h"
#include "tc_common.cuh"

namespace orcai {

namespace {

using namespace tc;

__host__ __device__ constexpr int cpad8(int c) { return (c + 7) & ~7; }
__host__ __device__ constexpr int cpad16(int c) { return (c + 15) & ~15; }
constexpr int kTileH = 16, kTileW = 8, kHaloW = kTileW + 2, kHaloH = kTileH + 2, kHalo = kHaloH * kHaloW;  // 180
constexpr float kTopDbF = 80.0f;

template <typename H> __device__ __forceinline__ uint4 relu8(uint4 v);
template <> __device__ __forceinline__ uint4 relu8<__half>(uint4 v) {
  __half2* h = reinterpret_cast<__half2*>(&v);
  const __half2 z = __float2half2_rn(0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __hmax2(h[i], z);
  return v;
}
template <> __device__ __forceinline__ uint4 relu8<__nv_bfloat16>(uint4 v) {
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
  const __nv_bfloat162 z = __float2bfloat162_rn(0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __hmax2(h[i], z);
  return v;
}

template <typename H>
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 r;
  H* h = reinterpret_cast<H*>(&r);
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = half_traits<H>::from_float(v[i]);
  return r;
}

// ------------------------------------------------------------------------------------------------
// fused separable convolution as implicit GEMM on tcgen05
// ------------------------------------------------------------------------------------------------
template <int CIN, int COUT>
struct SepTc {
  static constexpr int CP = cpad8(CIN);       // channel pitch of the input in global memory
  static constexpr int KP = cpad16(CIN);      // K per tap seen by the MMA
  static constexpr int NP = cpad16(COUT);     // N of the MMA
  static constexpr int NCH = KP / 8, NCH_G = CP / 8;
  static constexpr uint32_t LBO_A = kHalo * 16, SBO_A = kHaloW * 16;
  static constexpr uint32_t LBO_B = 128, SBO_B = NCH * 128;
  static constexpr int TM_COLS = NP <= 32 ? 32 : 64;
  static constexpr size_t w_bytes = (size_t)9 * NP * KP * 2;
  static constexpr size_t a_bytes = (size_t)NCH * LBO_A;
  static constexpr size_t smem = w_bytes + a_bytes + NP * sizeof(float) + 16;
};

template <int CIN, int COUT, bool RELU_IN, bool RELU_OUT, typename H, bool OUT_F32>
__global__ void __launch_bounds__(128)
sep_tc_kernel(const H* __restrict__ in, void* __restrict__ out, int Himg, int Wimg, long long n_snip,
              const H* __restrict__ wgt, const float* __restrict__ bias, int tiles_w, int tiles_h) {
  using S = SepTc<CIN, COUT>;
  constexpr int OCP = OUT_F32 ? COUT : cpad8(COUT);
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* s_w = smem;
  unsigned char* s_a = smem + S::w_bytes;
  float* s_bias = reinterpret_cast<float*>(s_a + S::a_bytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_bias + S::NP);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (int)(S::w_bytes / 16); i += 128) reinterpret_cast<uint4*>(s_w)[i] = __ldg(reinterpret_cast<const uint4*>(wgt) + i);
  for (int i = tid; i < (int)(S::a_bytes / 16); i += 128) reinterpret_cast<uint4*>(s_a)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < S::NP; i += 128) s_bias[i] = bias[i];
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncwarp();
  if (warp == 0) tmem_alloc<S::TM_COLS>(tslot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t a_base = smem_u32(s_a), w_base = smem_u32(s_w);
  constexpr uint32_t idesc = make_idesc_f16(128, S::NP, half_traits<H>::fmt);

  const long long tiles_per = (long long)tiles_w * tiles_h;
  const long long total = n_snip * tiles_per;
  uint32_t phase = 0;
  const int py = tid >> 3, px = tid & 7;
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const long long b = tile / tiles_per;
    const int tr = (int)(tile - b * tiles_per);
    const int h0 = (tr / tiles_w) * kTileH, w0 = (tr % tiles_w) * kTileW;
    const H* src = in + (size_t)b * Himg * Wimg * S::CP;
    // halo tile -> [k-chunk][halo pixel][8 channels]
    for (int idx = tid; idx < S::NCH_G * kHalo; idx += 128) {
      const int ch = idx / kHalo, hp = idx - ch * kHalo;
      const int hh = h0 + hp / kHaloW - 1, ww = w0 + hp % kHaloW - 1;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (hh >= 0 && hh < Himg && ww >= 0 && ww < Wimg) {
        v = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)hh * Wimg + ww) * S::CP + ch * 8));
        if (RELU_IN) v = relu8<H>(v);
      }
      *reinterpret_cast<uint4*>(s_a + (size_t)ch * S::LBO_A + hp * 16) = v;
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const uint32_t a_tap = a_base + ((tap / 3) * kHaloW + (tap % 3)) * 16;
        const uint32_t w_tap = w_base + tap * (S::NP * S::KP * 2);
#pragma unroll
        for (int ks = 0; ks < S::KP / 16; ++ks) {
          const uint64_t ad = make_smem_desc(a_tap + 2 * ks * S::LBO_A, S::LBO_A, S::SBO_A);
          const uint64_t bd = make_smem_desc(w_tap + 2 * ks * S::LBO_B, S::LBO_B, S::SBO_B);
          mma_f16_ss(tmem, ad, bd, idesc, (tap | ks) != 0);
        }
      }
      mma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // epilogue: accumulator row tid = pixel (py, px)
    const int hh = h0 + py, ww = w0 + px;
    const bool inside = hh < Himg && ww < Wimg;
    const size_t pix = ((size_t)b * Himg + hh) * Wimg + ww;
#pragma unroll
    for (int c0 = 0; c0 < S::NP; c0 += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] += s_bias[c0 + i];
        if (RELU_OUT) v[i] = fmaxf(v[i], 0.f);
      }
      if (inside) {
        if (OUT_F32) {
          float* o = reinterpret_cast<float*>(out) + pix * OCP + c0;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (c0 + 4 * q + 4 <= OCP) reinterpret_cast<float4*>(o)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
          H* o = reinterpret_cast<H*>(out) + pix * OCP + c0;
          if (c0 + 8 <= OCP) reinterpret_cast<uint4*>(o)[0] = pack8<H>(v);
          if (c0 + 16 <= OCP) reinterpret_cast<uint4*>(o)[1] = pack8<H>(v + 8);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc<S::TM_COLS>(tmem);
}

// ------------------------------------------------------------------------------------------------
// entry convolution (1 -> 16 channels): normalise-on-load, the 9 taps are the K dimension (padded to 16)
// ------------------------------------------------------------------------------------------------
template <typename H>
__global__ void __launch_bounds__(128)
conv0_tc_kernel(const float* __restrict__ in, int mode, long long first, int shift, int in_ld, int Himg, int Wimg,
                const SelectState* __restrict__ st, const H* __restrict__ wgt, const float* __restrict__ bias,
                H* __restrict__ out, H* __restrict__ out_sub, long long n_snip, int tiles_w, int tiles_h) {
  __shared__ __align__(128) unsigned char s_a[128 * 16 * 2];   // A: 128 rows x 16 k, LBO 128, SBO 256
  __shared__ __align__(128) unsigned char s_w[16 * 16 * 2];    // B: 16 rows x 16 k
  __shared__ float s_x[kHalo];
  __shared__ float s_bias[16];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid < 32) reinterpret_cast<uint4*>(s_w)[tid] = __ldg(reinterpret_cast<const uint4*>(wgt) + tid);
  if (tid < 16) s_bias[tid] = bias[tid];
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncwarp();
  if (warp == 0) tmem_alloc<32>(&tslot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  float db_ref = 0.f, lo = 0.f, hi = 1.f, range = 1.f;
  if (mode == 0) { db_ref = st->db_ref; lo = st->lo; hi = st->hi; range = hi - lo; }
  constexpr uint32_t idesc = make_idesc_f16(128, 16, half_traits<H>::fmt);
  const long long tiles_per = (long long)tiles_w * tiles_h;
  const long long total = n_snip * tiles_per;
  uint32_t phase = 0;
  const int py = tid >> 3, px = tid & 7;
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const long long b = tile / tiles_per;
    const int tr = (int)(tile - b * tiles_per);
    const int h0 = (tr / tiles_w) * kTileH, w0 = (tr % tiles_w) * kTileW;
    const long long row0 = (mode == 0) ? (first + b) * shift : b * (long long)Himg;
    for (int hp = tid; hp < kHalo; hp += 128) {
      const int hh = h0 + hp / kHaloW - 1, ww = w0 + hp % kHaloW - 1;
      float v = 0.f;
      if (hh >= 0 && hh < Himg && ww >= 0 && ww < Wimg) {
        v = in[(size_t)(row0 + hh) * in_ld + ww];
        if (mode == 0) {
          v = fmaxf(v - db_ref, -kTopDbF);
          v = __fdiv_rn(fminf(fmaxf(v, lo), hi) - lo, range);
        }
      }
      s_x[hp] = v;
    }
    __syncthreads();
    {
      float k[16];
#pragma unroll
      for (int t = 0; t < 9; ++t) k[t] = s_x[(py + t / 3) * kHaloW + px + t % 3];
#pragma unroll
      for (int t = 9; t < 16; ++t) k[t] = 0.f;
      unsigned char* row = s_a + (tid >> 3) * 256 + (tid & 7) * 16;
      *reinterpret_cast<uint4*>(row) = pack8<H>(k);
      *reinterpret_cast<uint4*>(row + 128) = pack8<H>(k + 8);
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mma_f16_ss(tmem, make_smem_desc(smem_u32(s_a), 128, 256), make_smem_desc(smem_u32(s_w), 128, 256), idesc, 0);
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i] + s_bias[i], 0.f);
    const int hh = h0 + py, ww = w0 + px;
    if (hh < Himg && ww < Wimg) {
      H* o = out + (((size_t)b * Himg + hh) * Wimg + ww) * 16;
      reinterpret_cast<uint4*>(o)[0] = pack8<H>(v);
      reinterpret_cast<uint4*>(o)[1] = pack8<H>(v + 8);
      if (out_sub != nullptr && !(hh & 1) && !(ww & 1)) {   // input of the first residual 1x1/2 convolution
        H* os = out_sub + (((size_t)b * (Himg >> 1) + (hh >> 1)) * ((Wimg + 1) >> 1) + (ww >> 1)) * 16;
        reinterpret_cast<uint4*>(os)[0] = pack8<H>(v);
        reinterpret_cast<uint4*>(os)[1] = pack8<H>(v + 8);
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc<32>(tmem);
}

// ------------------------------------------------------------------------------------------------
// MaxPool (3,2)/2 "same" + residual Conv1x1/2 + add on 16-bit activations (fp32 math)
// ------------------------------------------------------------------------------------------------
template <int CIN, int COUT, typename H>
__global__ void __launch_bounds__(256)
pool_res_tc_kernel(const H* __restrict__ t4, const H* __restrict__ prev, H* __restrict__ out, int Himg, int Wimg,
                   int Ho, int Wo, const float* __restrict__ rw, const float* __restrict__ rb, long long n_snip) {
  constexpr int ICP = cpad8(CIN), OCP = cpad8(COUT);
  __shared__ float s_w[CIN * COUT + COUT];
  for (int i = threadIdx.x; i < CIN * COUT; i += blockDim.x) s_w[i] = rw[i];
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) s_w[CIN * COUT + i] = rb[i];
  __syncthreads();
  const long long total = n_snip * Ho * Wo * OCP;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % OCP);
    long long r = idx / OCP;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const long long b = r / Ho;
    float res = 0.f;
    if (c < COUT) {
      const H* tb = t4 + (size_t)b * Himg * Wimg * OCP;
      float m = -INFINITY;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int hh = 2 * ho + dy, ww = 2 * wo + dx;
          if (hh < Himg && ww < Wimg) m = fmaxf(m, half_traits<H>::to_float(tb[((size_t)hh * Wimg + ww) * OCP + c]));
        }
      const H* pv = prev + (((size_t)b * Himg + 2 * ho) * Wimg + 2 * wo) * ICP;
      float a = s_w[CIN * COUT + c];
#pragma unroll 4
      for (int ci = 0; ci < CIN; ++ci) a = fmaf(half_traits<H>::to_float(pv[ci]), s_w[ci * COUT + c], a);
      res = m + a;
    }
    out[idx] = half_traits<H>::from_float(res);
  }
}

// entry-convolution weights for the CUDA-core producers / kernels: every FFMA takes its weight as a constant-bank operand
__constant__ float c_conv0[9 * 16 + 16];   // [tap][16] (BatchNorm folded), then bias[16]

#include "net_fused.cuh"
#include "net_fused_w.cuh"
#include "conv0_mma.cuh"

// ------------------------------------------------------------------------------------------------
// entry convolution for the fused path: fp32 CUDA-core math (K = 9 is no tensor-core shape), normalise-on-load from the
// raw dB buffer, folded BatchNorm + ReLU, fp16 NHWC output plus the even-position copy the first residual 1x1/2 reads.
// Weights sit in constant memory: every FFMA takes its weight as a constant-bank operand, no load instructions.
// ------------------------------------------------------------------------------------------------

constexpr int kC0TW = 32;
__host__ __device__ constexpr int c0_tile_rows(int rpt) { return 8 * rpt; }

// SPLIT: the result as (hi, lo) fp16 pairs, out_lo = fp16(v - hi) (fp32-grade path, net_precise.cuh)
// Row windows (net_precise.cuh, border images): image b is rows [off, off + Himg) of snippet b % n_snip, off = 0 for b < n_snip
// and off_bot for the others; zero padding applies outside the snippet's rows [0, Hfull).  Plain use: n_snip = image count,
// off_bot = 0, Hfull = Himg.
// One CTA = 8 warps = a tile of 8 * RPT rows x 32 columns; warp y, lane x owns column x of rows [RPT * y, RPT * y + RPT) and walks
// them with a sliding 3 x 3 window (RPT = 4 for tall images: the index arithmetic, the halo's normalisation and the barrier - half
// of the instructions with one pixel per thread - are shared by four pixels; RPT = 1 for the 8-row border images).
template <bool SPLIT, int RPT>
__global__ void __launch_bounds__(256, 6)
conv0_direct_kernel(const float* __restrict__ in, int mode, long long first, int shift, int in_ld, int Himg, int Wimg,
                    const SelectState* __restrict__ st, __half* __restrict__ out, __half* __restrict__ out_sub,
                    int tiles_w, int tiles_h, __half* __restrict__ out_lo, long long n_snip, int off_bot, int Hfull, long long plane_halfs) {
  constexpr int TH = c0_tile_rows(RPT);
  __shared__ float s_x[TH + 2][kC0TW + 2];
  const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
  const int tiles_per = tiles_w * tiles_h;
  const long long b = blockIdx.x / tiles_per;
  const int tr = (int)(blockIdx.x - b * tiles_per);
  const int h0 = (tr / tiles_w) * TH, w0 = (tr % tiles_w) * kC0TW;
  float db_ref = 0.f, lo = 0.f, hi = 1.f, range = 1.f;
  if (mode == 0) { db_ref = st->db_ref; lo = st->lo; hi = st->hi; range = hi - lo; }
  const long long snip = b % n_snip;
  const int off = b < n_snip ? 0 : off_bot;
  const long long row0 = ((mode == 0) ? (first + snip) * shift : snip * (long long)Hfull) + off;
  // halo: the (TH + 2) x 34 elements are dealt to the 256 threads as one flat list (division by the constant 34 = one multiply
  // + shift).  Dealing rows to warps and the two extra columns to lanes 0 / 1 ran the whole load + normalise body twice per row for
  // 34 useful lanes of 64: the halo cost as many instructions as half of the convolution.
  constexpr int HW = kC0TW + 2;
#pragma unroll
  for (int e = tid; e < (TH + 2) * HW; e += 256) {
    const int r = e / HW, cc = e - r * HW;
    const int hh = h0 + r - 1, ww = w0 + cc - 1;
    float v = 0.f;
    if (off + hh >= 0 && off + hh < Hfull && ww >= 0 && ww < Wimg) {
      v = in[(size_t)(row0 + hh) * in_ld + ww];
      if (mode == 0) {
        v = fmaxf(v - db_ref, -kTopDbF);
        v = __fdiv_rn(fminf(fmaxf(v, lo), hi) - lo, range);
      }
    }
    s_x[r][cc] = v;
  }
  __syncthreads();
  const int ww = w0 + tx;
  if (ww >= Wimg) return;
  float x[9];
#pragma unroll
  for (int t = 3; t < 9; ++t) x[t] = s_x[RPT * ty + t / 3 - 1][tx + t % 3];
#pragma unroll
  for (int rr = 0; rr < RPT; ++rr) {
    const int hh = h0 + RPT * ty + rr;
    if (hh >= Himg) return;
#pragma unroll
    for (int t = 0; t < 6; ++t) x[t] = x[t + 3];
#pragma unroll
    for (int t = 6; t < 9; ++t) x[t] = s_x[RPT * ty + rr + 2][tx + t - 6];
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = c_conv0[144 + c];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int c = 0; c < 16; c += 2) fused::fma2(acc[c], acc[c + 1], c_conv0[t * 16 + c], c_conv0[t * 16 + c + 1], x[t], x[t]);   // FFMA2: two channels per instruction
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = fmaxf(acc[c], 0.f);
    if constexpr (SPLIT) {
      uint4 h0v, l0v, h1v, l1v;
      fused::split8h(acc, h0v, l0v);
      fused::split8h(acc + 8, h1v, l1v);
      if (plane_halfs > 0) {
        // chunk-planar: channels 0-7 and 8-15 as two (n, H, W, 8) planes plane_halfs apart - rows of a TMA box are then contiguous
        const size_t at = (((size_t)b * Himg + hh) * Wimg + ww) * 8;
        *reinterpret_cast<uint4*>(out + at) = h0v;
        *reinterpret_cast<uint4*>(out + at + plane_halfs) = h1v;
        *reinterpret_cast<uint4*>(out_lo + at) = l0v;
        *reinterpret_cast<uint4*>(out_lo + at + plane_halfs) = l1v;
      } else {
        const size_t at = (((size_t)b * Himg + hh) * Wimg + ww) * 16;
        reinterpret_cast<uint4*>(out + at)[0] = h0v;
        reinterpret_cast<uint4*>(out + at)[1] = h1v;
        reinterpret_cast<uint4*>(out_lo + at)[0] = l0v;
        reinterpret_cast<uint4*>(out_lo + at)[1] = l1v;
      }
    } else {
      const uint4 v0 = fused::pack8h(acc), v1 = fused::pack8h(acc + 8);
      __half* o = out + (((size_t)b * Himg + hh) * Wimg + ww) * 16;
      reinterpret_cast<uint4*>(o)[0] = v0;
      reinterpret_cast<uint4*>(o)[1] = v1;
      if (out_sub != nullptr && !(hh & 1) && !(ww & 1)) {
        __half* os = out_sub + (((size_t)b * (Himg >> 1) + (hh >> 1)) * ((Wimg + 1) >> 1) + (ww >> 1)) * 16;
        reinterpret_cast<uint4*>(os)[0] = v0;
        reinterpret_cast<uint4*>(os)[1] = v1;
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <typename H> H host_cvt(float x);
template <> __half host_cvt<__half>(float x) { return __float2half_rn(x); }
template <> __nv_bfloat16 host_cvt<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// canonical K-major no-swizzle B operand: element (n, k) of tap t
inline size_t b_index(int t, int n, int k, int NP, int KP) {
  const size_t SBO = (size_t)(KP / 8) * 128;
  return ((size_t)t * NP * KP * 2 + (size_t)(n / 8) * SBO + (size_t)(k / 8) * 128 + (size_t)(n % 8) * 16 + (size_t)(k % 8) * 2) / 2;
}

template <typename H>
int build_sep(Ctx* c, const NetWeights::HostSep& hs, TcSep* out) {
  const int KP = cpad16(hs.ci), NP = cpad16(hs.co);
  std::vector<H> w((size_t)9 * NP * KP, host_cvt<H>(0.f));
  for (int t = 0; t < 9; ++t)
    for (int k = 0; k < hs.ci; ++k)
      for (int n = 0; n < hs.co; ++n)
        w[b_index(t, n, k, NP, KP)] = host_cvt<H>(hs.dw[(size_t)t * hs.ci + k] * hs.pw[(size_t)k * hs.co + n]);
  std::vector<float> b(NP, 0.f);
  for (int n = 0; n < hs.co; ++n) b[n] = hs.b[n];
  void* p = nullptr;
  ORCAI_CUDA(c, cudaMalloc(&p, w.size() * sizeof(H)));
  c->net->allocs.push_back(p);
  ORCAI_CUDA(c, cudaMemcpy(p, w.data(), w.size() * sizeof(H), cudaMemcpyHostToDevice));
  out->w = p;
  ORCAI_CHECK(net_upload(c, b, &out->bias));
  return ORCAI_OK;
}

template <typename H>
int prepare(Ctx* c, int fmt) {
  NetWeights* nw = c->net;
  {  // conv0: B[n][tap], taps 9..15 zero
    std::vector<H> w(16 * 16, host_cvt<H>(0.f));
    for (int t = 0; t < 9; ++t)
      for (int n = 0; n < 16; ++n) w[b_index(0, n, t, 16, 16)] = host_cvt<H>(nw->h_conv0_w[t * 16 + n]);
    void* p = nullptr;
    ORCAI_CUDA(c, cudaMalloc(&p, w.size() * sizeof(H)));
    nw->allocs.push_back(p);
    ORCAI_CUDA(c, cudaMemcpy(p, w.data(), w.size() * sizeof(H), cudaMemcpyHostToDevice));
    nw->tc_conv0_w[fmt] = p;
  }
  for (int b = 0; b < nw->n_blocks; ++b) {
    ORCAI_CHECK(build_sep<H>(c, nw->h_sep1[b], &nw->tc_sep1[fmt][b]));
    ORCAI_CHECK(build_sep<H>(c, nw->h_sep2[b], &nw->tc_sep2[fmt][b]));
  }
  ORCAI_CHECK(build_sep<H>(c, nw->h_fin, &nw->tc_fin[fmt]));
  nw->tc_ready[fmt] = true;
  return ORCAI_OK;
}

template <int CIN, int COUT, bool RI, bool RO, typename H, bool OUT_F32>
int run_sep(Ctx* c, const H* in, void* out, long long n, int Himg, int Wimg, const TcSep& w) {
  using S = SepTc<CIN, COUT>;
  // the attribute is per device: remember which devices have it (several contexts can live in one process)
  static std::atomic<unsigned long long> attr_devices{0ull};
  if (!((attr_devices.load() >> (c->device & 63)) & 1ull)) {
    ORCAI_CUDA(c, cudaFuncSetAttribute(sep_tc_kernel<CIN, COUT, RI, RO, H, OUT_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::smem));
    attr_devices.fetch_or(1ull << (c->device & 63));
  }
  const int tiles_w = (Wimg + kTileW - 1) / kTileW, tiles_h = (Himg + kTileH - 1) / kTileH;
  const long long total = n * tiles_w * tiles_h;
  // persistent CTAs: as many as fit (shared memory / 512 TMEM columns), capped by the tile count
  int per_sm = (int)std::min<size_t>(8, (size_t)(220 * 1024) / (S::smem + 1024));
  per_sm = std::max(1, std::min(per_sm, 512 / S::TM_COLS));
  const long long grid = std::min<long long>(total, (long long)c->sm_count * per_sm);
  sep_tc_kernel<CIN, COUT, RI, RO, H, OUT_F32><<<(unsigned)grid, 128, S::smem, c->stream>>>(
      in, out, Himg, Wimg, n, static_cast<const H*>(w.w), w.bias, tiles_w, tiles_h);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

template <int CIN, int COUT, typename H>
int run_block(Ctx* c, const H* prev, H* t_a, H* t_b, H* next, long long n, int Himg, int Wimg, int blk, int fmt) {
  NetWeights* nw = c->net;
  ORCAI_CHECK((run_sep<CIN, COUT, true, true, H, false>(c, prev, t_a, n, Himg, Wimg, nw->tc_sep1[fmt][blk])));
  ORCAI_CHECK((run_sep<COUT, COUT, false, false, H, false>(c, t_a, t_b, n, Himg, Wimg, nw->tc_sep2[fmt][blk])));
  const int Ho = (Himg + 1) / 2, Wo = (Wimg + 1) / 2;
  const long long total = n * Ho * Wo * cpad8(COUT);
  const long long grid = std::min<long long>((total + 255) / 256, (long long)c->sm_count * 32);
  pool_res_tc_kernel<CIN, COUT, H><<<(unsigned)grid, 256, 0, c->stream>>>(t_b, prev, next, Himg, Wimg, Ho, Wo, nw->res_w[blk], nw->res_b[blk], n);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

inline void set_debug(NetWeights* nw, const void* p, int kind, long long n, int h, int w, int ch, int pitch) {
  nw->dbg_ptr = p; nw->dbg_kind = kind; nw->dbg_n = n; nw->dbg_h = h; nw->dbg_w = w; nw->dbg_c = ch; nw->dbg_pitch = pitch;
}

// ---- calibration: per-channel sums of activation tensors (net_calibrate) ----------------------------------
// x: (n_rows, pitch) of fp32 (kind 0) or fp16 (kind 1); rows are pixels of (.., Himg, Wimg) images when `even` is set
__global__ void __launch_bounds__(256)
channel_sum_kernel(const void* __restrict__ x, int kind, long long n_rows, int pitch, int C, int relu, int Himg, int Wimg, int even,
                   double* __restrict__ sums) {
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  for (int ch = tx; ch < C; ch += 64) {
    double acc = 0.0;
    for (long long r = (long long)blockIdx.x * 4 + ty; r < n_rows; r += (long long)gridDim.x * 4) {
      if (even) {
        const int w = (int)(r % Wimg), h = (int)((r / Wimg) % Himg);
        if ((w | h) & 1) continue;
      }
      float v = kind == 0 ? static_cast<const float*>(x)[(size_t)r * pitch + ch] : __half2float(static_cast<const __half*>(x)[(size_t)r * pitch + ch]);
      if (relu) v = fmaxf(v, 0.f);
      acc += (double)v;
    }
    atomicAdd(&sums[ch], acc);
  }
}

// mean of the normalised spectrogram over rows [row0, row0 + n_rows) of the raw dB buffer
__global__ void __launch_bounds__(256)
spec_sum_kernel(const float* __restrict__ raw, long long n_rows, int ld, int nb, const SelectState* __restrict__ st, double* __restrict__ sum) {
  const float db_ref = st->db_ref, lo = st->lo, hi = st->hi, range = hi - lo;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows * nb; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nb;
    const int b = (int)(i - r * nb);
    float v = fmaxf(raw[(size_t)r * ld + b] - db_ref, -kTopDbF);
    acc += (double)__fdiv_rn(fminf(fmaxf(v, lo), hi) - lo, range);
  }
  atomicAdd(sum, acc);
}

struct CalibSink {          // device accumulators laid out in slots of kCalSlot doubles
  static constexpr int kCalSlot = 512;
  enum { IN_RELU = 0, IN_EVEN = 4, S1 = 8, FIN = 12, FEAT = 13, H1 = 14, H2 = 15, SPEC = 16, NSLOTS = 17 };
  double* d = nullptr;
  Ctx* c = nullptr;
  void add(int slot, const void* x, int kind, long long rows, int pitch, int C, int relu, int Himg = 1, int Wimg = 1, int even = 0) const {
    const long long grid = std::min<long long>((rows + 3) / 4, 1184);
    channel_sum_kernel<<<(unsigned)std::max<long long>(grid, 1), 256, 0, c->stream>>>(x, kind, rows, pitch, C, relu, Himg, Wimg, even, d + (size_t)slot * kCalSlot);
    c->launches++;
  }
};

template <typename H>
int forward_tc(Ctx* c, const float* d_in, int input_mode, int64_t first, int64_t n, float* d_preds, int fmt, const CalibSink* cal = nullptr) {
  NetWeights* nw = c->net;
  const int Himg = nw->H, Wf = nw->Wf, U = nw->U, L = nw->L;
  const int Tn = Himg >> nw->n_blocks;
  const int kind = fmt + 1;
  // workspace (16-bit elements): two "prev" buffers, two block temporaries; then the fp32 tail
  const size_t prev_max = (size_t)Himg * Wf * 16;
  const size_t t_max = (size_t)Himg * Wf * 32;
  const size_t tail_f = (size_t)Tn * (nw->feat + 2 * 4 * U + 2 * U + 2 * U + 128);
  const size_t per = (2 * prev_max + 2 * t_max) * 2 + tail_f * 4 + 256;
  const long long chunk = std::min<long long>(std::max(nw->chunk, 1), n);
  if (chunk <= 0) return ORCAI_OK;
  ORCAI_CHECK(ensure_device_buffer(c, &nw->tc_ws, &nw->tc_ws_cap, per * (size_t)chunk));
  H* pA = static_cast<H*>(nw->tc_ws);
  H* pB = pA + prev_max * chunk;
  H* tA = pB + prev_max * chunk;
  H* tB = tA + t_max * chunk;
  float* feat = reinterpret_cast<float*>(tB + t_max * chunk);
  float* scratch = feat + (size_t)Tn * nw->feat * chunk;
  const int shift = c->p.snippet_len / 2;
  nw->mark_i = 0;
  nw->dbg_ptr = nullptr;
  const int stop = nw->debug_stop;

  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    const long long m = std::min<long long>(chunk, n - s0);
    const bool mk = (s0 == 0);
    if (mk) nw->marked_snippets = m;
    net_mark(c, mk);
    {
      const int tiles_w = (Wf + kTileW - 1) / kTileW, tiles_h = (Himg + kTileH - 1) / kTileH;
      const long long total = m * tiles_w * tiles_h;
      const long long grid = std::min<long long>(total, (long long)c->sm_count * 8);
      const float* src = (input_mode == 0) ? d_in : d_in + (size_t)s0 * Himg * Wf;
      conv0_tc_kernel<H><<<(unsigned)grid, 128, 0, c->stream>>>(src, input_mode, first + s0, shift, input_mode == 0 ? kRawLd : Wf, Himg, Wf,
                                                                 c->d_sel, static_cast<const H*>(nw->tc_conv0_w[fmt]), nw->conv0_b, pA,
                                                                 static_cast<H*>(nullptr), m, tiles_w, tiles_h);
      c->launches++;
      ORCAI_CUDA(c, cudaGetLastError());
    }
    net_mark(c, mk);  // 0: conv0
    if (stop == 0) { set_debug(nw, pA, kind, m, Himg, Wf, 16, 16); return ORCAI_OK; }
    int h = Himg, w = Wf;
    auto cal_in = [&](int blk, const H* x, int C, int pitch) {   // A operands fed by a block's input tensor
      if (!cal) return;
      cal->add(CalibSink::IN_RELU + blk, x, 1, m * h * w, pitch, C, 1);
      cal->add(CalibSink::IN_EVEN + blk, x, 1, m * h * w, pitch, C, 0, h, w, 1);
    };
    auto cal_s1 = [&](int blk, int C, int pitch) { if (cal) cal->add(CalibSink::S1 + blk, tA, 1, m * h * w, pitch, C, 0); };
    cal_in(0, pA, 16, 16);
    ORCAI_CHECK((run_block<16, 30, H>(c, pA, tA, tB, pB, m, h, w, 0, fmt)));
    cal_s1(0, 30, 32);
    if (stop == 10) { set_debug(nw, tA, kind, m, h, w, 30, 32); return ORCAI_OK; }
    if (stop == 11) { set_debug(nw, tB, kind, m, h, w, 30, 32); return ORCAI_OK; }
    h = (h + 1) / 2; w = (w + 1) / 2;
    net_mark(c, mk);  // 1: block1
    if (stop == 1) { set_debug(nw, pB, kind, m, h, w, 30, 32); return ORCAI_OK; }
    cal_in(1, pB, 30, 32);
    ORCAI_CHECK((run_block<30, 40, H>(c, pB, tA, tB, pA, m, h, w, 1, fmt)));
    cal_s1(1, 40, 40);
    h = (h + 1) / 2; w = (w + 1) / 2;
    net_mark(c, mk);  // 2: block2
    if (stop == 2) { set_debug(nw, pA, kind, m, h, w, 40, 40); return ORCAI_OK; }
    cal_in(2, pA, 40, 40);
    ORCAI_CHECK((run_block<40, 50, H>(c, pA, tA, tB, pB, m, h, w, 2, fmt)));
    cal_s1(2, 50, 56);
    h = (h + 1) / 2; w = (w + 1) / 2;
    net_mark(c, mk);  // 3: block3
    if (stop == 3) { set_debug(nw, pB, kind, m, h, w, 50, 56); return ORCAI_OK; }
    cal_in(3, pB, 50, 56);
    ORCAI_CHECK((run_block<50, 60, H>(c, pB, tA, tB, pA, m, h, w, 3, fmt)));
    cal_s1(3, 60, 64);
    h = (h + 1) / 2; w = (w + 1) / 2;
    net_mark(c, mk);  // 4: block4
    if (stop == 4) { set_debug(nw, pA, kind, m, h, w, 60, 64); return ORCAI_OK; }
    if (cal) cal->add(CalibSink::FIN, pA, 1, m * h * w, 64, 60, 0);
    ORCAI_CHECK((run_sep<60, 36, false, true, H, true>(c, pA, feat, m, h, w, nw->tc_fin[fmt])));
    net_mark(c, mk);  // 5: final sepconv (fp32 features, w*36+c)
    if (stop == 5) { set_debug(nw, feat, 0, m, h, w, 36, 36); return ORCAI_OK; }
    ORCAI_CHECK(net_tail_fp32(c, feat, scratch, m, d_preds + (size_t)s0 * Tn * L, mk));
    if (cal) {   // tail inputs: features, hidden states of both LSTM layers (net_tail_fp32's scratch layout)
      const long long rows = m * Tn;
      const float* h1 = scratch + (size_t)rows * 2 * 4 * U;
      const float* h2 = h1 + (size_t)rows * 2 * U;
      cal->add(CalibSink::FEAT, feat, 0, rows, nw->feat, nw->feat, 0);
      cal->add(CalibSink::H1, h1, 0, rows, 2 * U, 2 * U, 0);
      cal->add(CalibSink::H2, h2, 0, rows, 2 * U, 2 * U, 0);
    }
  }
  return ORCAI_OK;
}


// ------------------------------------------------------------------------------------------------
// fused path (net_path 3): conv0 -> 4 fused residual-block kernels -> final sepconv -> fp32 LSTM/dense tail
// ------------------------------------------------------------------------------------------------
using FB1 = fused::FB<16, 30, 29, 6, true, 2, 8>;   // two CTAs per SM already give the tensor pipe two streams (a second issuer warp: 1.91 -> 2.01 ms per 10 min)
// blocks 2-4 run one CTA per SM: two issuer warps, or the single MMA stream would cost ~59 cycles per MMA instead of 47-52
using FB2 = fused::FB<30, 40, 43, 4, true, 1, 16, false, 2, 2>;   // the only block whose shared memory has room for a second X buffer
using FB3 = fused::FB<40, 50, 22, 4, true, 1, 16, false, 2>;
using FB4 = fused::FB<50, 60, 11, 4, false, 1, 16, false, 2>;
using FB1W = fused::FBW<16, 30, 29, 6, true, 2, 8>;          // block 1, N-widened MMAs (block1_path 1)
using FB1C = fused::FB<16, 30, 29, 6, true, 2, 8, true>;   // block 1 with the entry convolution fused in (conv0_path 2)
static_assert(FB1C::W_BYTES == FB1::W_BYTES && FB1C::OFF_SPEC % 128 == 0, "FB1C shares FB1's weight pack; TMA destinations are 128-byte aligned");

template <class G>
int build_fused_block(Ctx* c, int blk) {
  NetWeights* nw = c->net;
  const NetWeights::HostSep& s1 = nw->h_sep1[blk];
  const NetWeights::HostSep& s2 = nw->h_sep2[blk];
  if (s1.ci != G::CIN || s1.co != G::COUT || s2.ci != G::COUT || s2.co != G::COUT)
    ORCAI_FAIL(c, ORCAI_ERR_ARG, "fused block %d: kernel geometry does not match the loaded weights", blk + 1);
  std::vector<__half> w(G::W_BYTES / 2, __float2half_rn(0.f));
  auto at = [](uint32_t off, uint32_t sbo, int n, int k) { return (off + (uint32_t)(n / 8) * sbo + (uint32_t)(k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) / 2; };
  // d?[n] = sum_k (fp16(W[k][n]) - W[k][n]) * mean(A[k]): what fp16 weight rounding adds to every output pixel (Calib, net.h)
  const Calib& cal = nw->calib;
  std::vector<double> d1(G::COUT, 0.0), d2(G::COUT, 0.0), dr(G::COUT, 0.0);
  auto put = [&](size_t idx, double v, double mu, double* acc) {
    const __half q = __float2half_rn((float)v);
    w[idx] = q;
    if (G::PREC) w[idx + G::W_LO / 2] = __float2half_rn((float)(v - (double)__half2float(q)));   // lo set: the weight is exact to 2^-22
    else *acc += ((double)__half2float(q) - v) * mu;
  };
  for (int t = 0; t < 9; ++t) {
    for (int k = 0; k < G::CIN; ++k)
      for (int n = 0; n < G::COUT; ++n)
        put(at(G::OFF_W1 + t * G::TAP_W1, G::SBO_W1, n, k), (double)(s1.dw[(size_t)t * G::CIN + k] * s1.pw[(size_t)k * G::COUT + n]),
            cal.valid ? cal.in_relu[blk][k] : 0.0, &d1[n]);
    if (!G::UF2)
      for (int k = 0; k < G::COUT; ++k)
        for (int n = 0; n < G::COUT; ++n)
          put(at(G::OFF_W2 + t * G::TAP_W2, G::SBO_W2, n, k), (double)(s2.dw[(size_t)t * G::COUT + k] * s2.pw[(size_t)k * G::COUT + n]),
              cal.valid ? cal.s1[blk][k] : 0.0, &d2[n]);
  }
  if (G::UF2) {   // un-folded second convolution: pointwise weights as the GEMM operand, depthwise taps as fp32 for the worker warps
    for (int k = 0; k < G::COUT; ++k)
      for (int n = 0; n < G::COUT; ++n) put(at(G::OFF_W2, G::SBO_W2, n, k), (double)s2.pw[(size_t)k * G::COUT + n], 0.0, &d2[n]);
    float* taps = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(w.data()) + G::OFF_DW2);
    for (int t = 0; t < 9; ++t)
      for (int k = 0; k < G::COUT; ++k) taps[t * G::NP + k] = s2.dw[(size_t)t * G::COUT + k];
  }
  const std::vector<float>& rw = nw->h_res_w[blk];
  for (int k = 0; k < G::CIN; ++k)
    for (int n = 0; n < G::COUT; ++n) put(at(G::OFF_WR, G::SBO_W1, n, k), (double)rw[(size_t)k * G::COUT + n], cal.valid ? cal.in_even[blk][k] : 0.0, &dr[n]);
  std::vector<float> b1c(G::COUT), b2c(G::COUT), brc(G::COUT);
  for (int n = 0; n < G::COUT; ++n) {
    b1c[n] = (float)((double)s1.b[n] - d1[n]);
    b2c[n] = (float)((double)s2.b[n] - d2[n]);
    brc[n] = (float)((double)nw->h_res_b[blk][n] - dr[n]);
  }
  // bias rows [hi, lo] (k = 0, 1) of the three GEMMs and the constant "ones" A operand
  auto put_bias = [&](uint32_t off, const float* bv) {
    for (int n = 0; n < G::COUT; ++n) {
      const __half hi = __float2half_rn(bv[n]);
      const __half lo = __float2half_rn(bv[n] - __half2float(hi));
      w[(off + (uint32_t)(n / 8) * 128 + (n % 8) * 16) / 2] = hi;
      w[(off + (uint32_t)(n / 8) * 128 + (n % 8) * 16) / 2 + 1] = lo;
    }
  };
  put_bias(G::OFF_WB1, b1c.data());
  put_bias(G::OFF_WB2, b2c.data());
  put_bias(G::OFF_WBR, brc.data());
  for (int r = 0; r < 8; ++r) {
    w[(G::OFF_ONES + r * 16) / 2] = __float2half_rn(1.f);
    w[(G::OFF_ONES + r * 16) / 2 + 1] = __float2half_rn(1.f);
  }
  void* p = nullptr;
  ORCAI_CUDA(c, cudaMalloc(&p, G::W_BYTES));
  nw->allocs.push_back(p);
  ORCAI_CUDA(c, cudaMemcpy(p, w.data(), G::W_BYTES, cudaMemcpyHostToDevice));
  (G::PREC ? nw->fbp_w[blk] : nw->fb_w[blk]) = p;
  ORCAI_CUDA(c, cudaFuncSetAttribute(fused::fused_block_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
  return ORCAI_OK;
}

// operands of the N-widened block kernel (net_fused_w.cuh): per dy one B matrix whose rows are (dx, n)
template <class G>
int build_fused_block_w(Ctx* c, int blk, void** out) {
  NetWeights* nw = c->net;
  const NetWeights::HostSep& s1 = nw->h_sep1[blk];
  const NetWeights::HostSep& s2 = nw->h_sep2[blk];
  if (s1.ci != G::CIN || s1.co != G::COUT || s2.ci != G::COUT || s2.co != G::COUT)
    ORCAI_FAIL(c, ORCAI_ERR_ARG, "fused block %d: kernel geometry does not match the loaded weights", blk + 1);
  std::vector<__half> w(G::W_BYTES / 2, __float2half_rn(0.f));
  auto at = [](uint32_t off, uint32_t sbo, int n, int k) { return (off + (uint32_t)(n / 8) * sbo + (uint32_t)(k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) / 2; };
  const Calib& cal = nw->calib;
  std::vector<double> d1(G::COUT, 0.0), d2(G::COUT, 0.0), dr(G::COUT, 0.0);
  auto put = [&](size_t idx, double v, double mu, double* acc) {
    const __half q = __float2half_rn((float)v);
    w[idx] = q;
    *acc += ((double)__half2float(q) - v) * mu;
  };
  for (int t = 0; t < 9; ++t) {
    const int dy = t / 3, dx = t % 3;
    for (int k = 0; k < G::CIN; ++k)
      for (int n = 0; n < G::COUT; ++n)
        put(at(G::OFF_W1 + dy * G::DY_W1, G::SBO_W1, dx * G::NP + n, k), (double)(s1.dw[(size_t)t * G::CIN + k] * s1.pw[(size_t)k * G::COUT + n]),
            cal.valid ? cal.in_relu[blk][k] : 0.0, &d1[n]);
    for (int k = 0; k < G::COUT; ++k)
      for (int n = 0; n < G::COUT; ++n)
        put(at(G::OFF_W2 + dy * G::DY_W2, G::SBO_W2, dx * G::NP + n, k), (double)(s2.dw[(size_t)t * G::COUT + k] * s2.pw[(size_t)k * G::COUT + n]),
            cal.valid ? cal.s1[blk][k] : 0.0, &d2[n]);
  }
  const std::vector<float>& rw = nw->h_res_w[blk];
  for (int k = 0; k < G::CIN; ++k)
    for (int n = 0; n < G::COUT; ++n) put(at(G::OFF_WR, G::SBO_W1, n, k), (double)rw[(size_t)k * G::COUT + n], cal.valid ? cal.in_even[blk][k] : 0.0, &dr[n]);
  // bias rows [hi, lo] (k = 0, 1): the convolutions' sit in the centre column block (dx = 1) only
  auto put_bias = [&](uint32_t off, int n0, double bv) {
    const float b = (float)bv;
    const __half hi = __float2half_rn(b);
    const __half lo = __float2half_rn(b - __half2float(hi));
    w[(off + (uint32_t)(n0 / 8) * 128 + (n0 % 8) * 16) / 2] = hi;
    w[(off + (uint32_t)(n0 / 8) * 128 + (n0 % 8) * 16) / 2 + 1] = lo;
  };
  for (int n = 0; n < G::COUT; ++n) {
    put_bias(G::OFF_WB1, G::NP + n, (double)s1.b[n] - d1[n]);
    put_bias(G::OFF_WB2, G::NP + n, (double)s2.b[n] - d2[n]);
    put_bias(G::OFF_WBR, n, (double)nw->h_res_b[blk][n] - dr[n]);
  }
  for (int r = 0; r < 8; ++r) {
    w[(G::OFF_ONES + r * 16) / 2] = __float2half_rn(1.f);
    w[(G::OFF_ONES + r * 16) / 2 + 1] = __float2half_rn(1.f);
  }
  void* p = nullptr;
  ORCAI_CUDA(c, cudaMalloc(&p, G::W_BYTES));
  nw->allocs.push_back(p);
  ORCAI_CUDA(c, cudaMemcpy(p, w.data(), G::W_BYTES, cudaMemcpyHostToDevice));
  *out = p;
  ORCAI_CUDA(c, cudaFuncSetAttribute(fused::fused_block_w_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
  return ORCAI_OK;
}

// banded (Toeplitz) B operand of the pixel-group entry convolution + bias rows + ones tile (conv0_mma.cuh)
int build_conv0_mma(Ctx* c) {
  NetWeights* nw = c->net;
  std::vector<__half> w(conv0::kWBytes / 2, __float2half_rn(0.f));
  auto at = [](int blk, int n, int k) { return (conv0::OFF_B + blk * conv0::kBBlock + (uint32_t)(n / 8) * 256 + (uint32_t)(k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) / 2; };
  for (int dy = 0; dy < 3; ++dy)
    for (int pr = 0; pr < 2; ++pr)
      for (int k = 0; k < 16; ++k) {
        if (pr == 1 && k >= 8) continue;                 // second chunk of the [g+1 | -] step carries no weight
        const int pos = pr == 0 ? k - 8 : 8 + k;          // input pixel relative to the first pixel of the output group
        for (int j = 0; j < 8; ++j) {
          const int dx = pos - j + 1;
          if (dx < 0 || dx > 2) continue;
          for (int ch = 0; ch < 16; ++ch) w[at(dy * 2 + pr, j * 16 + ch, k)] = __float2half_rn(nw->h_conv0_w[(dy * 3 + dx) * 16 + ch]);
        }
      }
  // weight-rounding correction (Calib): every tap multiplies the same mean spectrogram value
  double d0[16] = {};
  if (nw->calib.valid)
    for (int t = 0; t < 9; ++t)
      for (int ch = 0; ch < 16; ++ch) {
        const float v = nw->h_conv0_w[t * 16 + ch];
        d0[ch] += ((double)__half2float(__float2half_rn(v)) - (double)v) * nw->calib.spec;
      }
  for (int j = 0; j < 8; ++j)
    for (int ch = 0; ch < 16; ++ch) {
      const float bv = (float)((double)nw->h_conv0_b[ch] - d0[ch]);
      const __half hi = __float2half_rn(bv);
      w[at(6, j * 16 + ch, 0)] = hi;
      w[at(6, j * 16 + ch, 1)] = __float2half_rn(bv - __half2float(hi));
    }
  for (int r = 0; r < 8; ++r) {
    w[(conv0::OFF_ONES + r * 16) / 2] = __float2half_rn(1.f);
    w[(conv0::OFF_ONES + r * 16) / 2 + 1] = __float2half_rn(1.f);
  }
  void* p = nullptr;
  ORCAI_CUDA(c, cudaMalloc(&p, conv0::kWBytes));
  nw->allocs.push_back(p);
  ORCAI_CUDA(c, cudaMemcpy(p, w.data(), conv0::kWBytes, cudaMemcpyHostToDevice));
  nw->conv0_mma_w = p;
  ORCAI_CUDA(c, cudaFuncSetAttribute(conv0::conv0_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv0::kSmem));
  return ORCAI_OK;
}

unsigned int* g_trap_host = nullptr;   // mapped host memory behind tc::g_trap_info (-DORCAI_TRAP_INFO builds, ORCAI_B200_TRAPINFO=1)
#ifdef ORCAI_FUSED_TRACE
long long* g_trace_dev = nullptr;      // device memory behind fused::g_trace (ORCAI_B200_TRACE=<file>)
#endif

// bring-up aids shared by the fused paths (compiled out of the product)
int prepare_bringup_aids(Ctx* c) {
#ifdef ORCAI_FUSED_TRACE
  if (g_trace_dev == nullptr && getenv("ORCAI_B200_TRACE") != nullptr) {
    ORCAI_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&g_trace_dev), 8 * 8 * 256));
    ORCAI_CUDA(c, cudaMemset(g_trace_dev, 0, 8 * 8 * 256));
    ORCAI_CUDA(c, cudaMemcpyToSymbol(fused::g_trace, &g_trace_dev, sizeof g_trace_dev));
  }
#endif
#ifdef ORCAI_TRAP_INFO
  if (g_trap_host == nullptr && getenv("ORCAI_B200_TRAPINFO") != nullptr) {
    unsigned int* d = nullptr;
    ORCAI_CUDA(c, cudaHostAlloc(reinterpret_cast<void**>(&g_trap_host), 64, cudaHostAllocMapped));
    memset(g_trap_host, 0, 64);
    ORCAI_CUDA(c, cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), g_trap_host, 0));
    ORCAI_CUDA(c, cudaMemcpyToSymbol(tc::g_trap_info, &d, sizeof d));
  }
#endif
  (void)c;
  return ORCAI_OK;
}

int prepare_fused(Ctx* c) {
  NetWeights* nw = c->net;
  if (nw->fused_ready) return ORCAI_OK;
  ORCAI_CHECK(prepare_bringup_aids(c));
  ORCAI_CHECK(net_tc_prepare(c, 0));   // the final sepconv reuses the fp16 layer-wise operands
  ORCAI_CHECK(build_conv0_mma(c));
  ORCAI_CHECK(build_fused_block<FB1>(c, 0));
  ORCAI_CUDA(c, cudaFuncSetAttribute(fused::fused_block_kernel<FB1C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FB1C::SMEM));
  ORCAI_CHECK(build_fused_block_w<FB1W>(c, 0, &nw->fb_w1_wide));
  ORCAI_CHECK(build_fused_block<FB2>(c, 1));
  ORCAI_CHECK(build_fused_block<FB3>(c, 2));
  ORCAI_CHECK(build_fused_block<FB4>(c, 3));
  {  // final separable convolution: the layer-wise fp16 operand with a corrected bias
    const NetWeights::HostSep& hs = nw->h_fin;
    std::vector<float> b(cpad16(hs.co), 0.f);
    for (int n = 0; n < hs.co; ++n) {
      double d = 0.0;
      if (nw->calib.valid)
        for (int t = 0; t < 9; ++t)
          for (int k = 0; k < hs.ci; ++k) {
            const float v = hs.dw[(size_t)t * hs.ci + k] * hs.pw[(size_t)k * hs.co + n];
            d += ((double)__half2float(__float2half_rn(v)) - (double)v) * nw->calib.fin[k];
          }
      b[n] = (float)((double)hs.b[n] - d);
    }
    nw->fb_fin.w = nw->tc_fin[0].w;
    ORCAI_CHECK(net_upload(c, b, &nw->fb_fin.bias));
  }
  nw->fused_ready = true;
  return ORCAI_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// rank-5 map over a (n, h, w, chunks, 8) fp16 NHWC tensor; box = {8 ch, 1 chunk, box_w, box_h, 1 snippet}
// `step` = traversal stride along w and h (2: every other pixel, box_w x box_h pixels are still what lands in shared memory)
// img_rows > 0: consecutive images start img_rows rows apart (overlapping row windows of one tall image) instead of h rows
int make_act_map(Ctx* c, CUtensorMap* map, const __half* base, long long n, int h, int w, int cpitch, int box_w, int box_h, int step = 1,
                 long long img_rows = 0) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) ORCAI_FAIL(c, ORCAI_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[5] = {8, (cuuint64_t)(cpitch / 8), (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  const cuuint64_t strides[4] = {16, (cuuint64_t)cpitch * 2, (cuuint64_t)w * cpitch * 2, (cuuint64_t)(img_rows > 0 ? img_rows : h) * w * cpitch * 2};
  const cuuint32_t box[5] = {8, 1, (cuuint32_t)(box_w * step), (cuuint32_t)(box_h * step), 1};
  const cuuint32_t estr[5] = {1, 1, (cuuint32_t)step, (cuuint32_t)step, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) ORCAI_FAIL(c, ORCAI_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for tensor (%lld, %d, %d, %d)", (int)r, n, h, w, cpitch);
  return ORCAI_OK;
}

// PREC blocks take their input tensors as (hi, lo) pairs (xr_lo, xs_lo) and write fp32 tensors (yr, ys point to floats)
// rank-5 map over a chunk-PLANAR fp16 activation tensor: `planes` planes of (n, h, w, 8) halfs, plane_halfs apart;
// box = {8 ch, box_w, box_h, 1 image, 1 plane}: every row of a box is one contiguous run of box_w * 16 bytes
int make_planar_map(Ctx* c, CUtensorMap* map, const __half* base, long long n, int h, int w, int planes, long long plane_halfs, int box_w, int box_h,
                    int step = 1, long long img_rows = 0) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) ORCAI_FAIL(c, ORCAI_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[5] = {8, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n, (cuuint64_t)planes};
  const cuuint64_t strides[4] = {16, (cuuint64_t)w * 16, (cuuint64_t)(img_rows > 0 ? img_rows : h) * w * 16, (cuuint64_t)plane_halfs * 2};
  const cuuint32_t box[5] = {8, (cuuint32_t)(box_w * step), (cuuint32_t)(box_h * step), 1, 1};
  const cuuint32_t estr[5] = {1, (cuuint32_t)step, (cuuint32_t)step, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) ORCAI_FAIL(c, ORCAI_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for the planar tensor (%lld, %d, %d) x %d", (int)r, n, h, w, planes);
  return ORCAI_OK;
}

template <class G>
int run_fused_block(Ctx* c, int blk, const __half* xr, const __half* xs, __half* yr, __half* ys, long long m, int Himg, int Wimg,
                    const __half* xr_lo = nullptr, const __half* xs_lo = nullptr, long long plane_halfs = 0) {
  NetWeights* nw = c->net;
  const int Ho = Himg / 2, Wo = (Wimg + 1) / 2;
  const int n_strips = (Wo + G::CP - 1) / G::CP;
  const long long items = m * n_strips;
  const long long grid = std::min<long long>(items, (long long)c->sm_count * G::CTAS);
  CUtensorMap tmx, tmr, tmxl, tmrl;
  ORCAI_CHECK(make_act_map(c, &tmx, xr, m, Himg, Wimg, G::ICP, G::WP, G::S + 2));
  // residual input = x at even positions: its own tensor, or (xs == nullptr: x is non-negative, ReLU(x) == x) the full tensor
  // traversed with element stride 2
  if (xs) ORCAI_CHECK(make_act_map(c, &tmr, xs, m, Ho, Wo, G::ICP, G::CP, G::S / 2));
  else ORCAI_CHECK(make_act_map(c, &tmr, xr, m, Himg, Wimg, G::ICP, G::CP, G::S / 2, 2));
  tmxl = tmx; tmrl = tmr;
  if (G::PREC) {   // chunk-planar (hi, lo) inputs; the residual convolution walks the same tensors with stride 2
    (void)xs_lo;
    ORCAI_CHECK(make_planar_map(c, &tmx, xr, m, Himg, Wimg, G::XG, plane_halfs, G::WP, G::S + 2));
    ORCAI_CHECK(make_planar_map(c, &tmr, xr, m, Himg, Wimg, G::XG, plane_halfs, G::CP, G::S / 2, 2));
    ORCAI_CHECK(make_planar_map(c, &tmxl, xr_lo, m, Himg, Wimg, G::XG, plane_halfs, G::WP, G::S + 2));
    ORCAI_CHECK(make_planar_map(c, &tmrl, xr_lo, m, Himg, Wimg, G::XG, plane_halfs, G::CP, G::S / 2, 2));
  }
  fused::fused_block_kernel<G><<<(unsigned)grid, G::NTHREADS, G::SMEM, c->stream>>>(tmx, tmr, xs ? 1 : 2, yr, ys, Himg, Wimg, n_strips, items,
                                                                                   static_cast<const unsigned char*>(G::PREC ? nw->fbp_w[blk] : nw->fb_w[blk]),
                                                                                   tmxl, tmrl, fused::TallView());
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

// block 1 through the N-widened kernel (net_fused_w.cuh)
int run_fused_block1_wide(Ctx* c, const __half* xr, __half* yr, __half* ys, long long m, int Himg, int Wimg) {
  using G = FB1W;
  NetWeights* nw = c->net;
  const int Wo = (Wimg + 1) / 2;
  const int n_strips = (Wo + G::CP - 1) / G::CP;
  const long long items = m * n_strips;
  const long long grid = std::min<long long>(items, (long long)c->sm_count * G::CTAS);
  CUtensorMap tmx, tmr;
  ORCAI_CHECK(make_act_map(c, &tmx, xr, m, Himg, Wimg, G::ICP, G::WP, G::S + 2));
  ORCAI_CHECK(make_act_map(c, &tmr, xr, m, Himg, Wimg, G::ICP, G::CP, G::S / 2, 2));
  fused::fused_block_w_kernel<G><<<(unsigned)grid, G::NTHREADS, G::SMEM, c->stream>>>(tmx, tmr, 2, yr, ys, Himg, Wimg, n_strips, items,
                                                                                     static_cast<const unsigned char*>(nw->fb_w1_wide));
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

// block 1 fed by the fp16 normalised spectrogram: its producer warps run the entry convolution (fused::FB<..., CONV0 = true>)
int run_fused_block1_conv0(Ctx* c, const __half* spec16, long long snippet_stride_rows, __half* yr, __half* ys, long long m, int Himg, int Wimg) {
  using G = FB1C;
  NetWeights* nw = c->net;
  const int Wo = (Wimg + 1) / 2;
  const int n_strips = (Wo + G::CP - 1) / G::CP;
  const long long items = m * n_strips;
  const long long grid = std::min<long long>(items, (long long)c->sm_count * G::CTAS);
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) ORCAI_FAIL(c, ORCAI_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap tms;
  const cuuint64_t row_bytes = (cuuint64_t)n_strips * G::SPW * 2;
  const cuuint64_t dims[4] = {(cuuint64_t)G::SPW, (cuuint64_t)n_strips, (cuuint64_t)Himg, (cuuint64_t)m};
  const cuuint64_t strides[3] = {(cuuint64_t)G::SPW * 2, row_bytes, (cuuint64_t)snippet_stride_rows * row_bytes};
  const cuuint32_t box[4] = {(cuuint32_t)G::SPW, 1, (cuuint32_t)G::SPH, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(&tms, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(spec16), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) ORCAI_FAIL(c, ORCAI_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for the spectrogram view of block 1", (int)r);
  fused::fused_block_kernel<G><<<(unsigned)grid, G::NTHREADS, G::SMEM, c->stream>>>(tms, tms, 2, yr, ys, Himg, Wimg, n_strips, items,
                                                                                   static_cast<const unsigned char*>(nw->fb_w[0]), tms, tms, fused::TallView());
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

int forward_fused(Ctx* c, const float* d_in, int input_mode, int64_t first, int64_t n, float* d_preds) {
  NetWeights* nw = c->net;
  using H = __half;
  const int Himg = nw->H, Wf = nw->Wf, U = nw->U, L = nw->L;
  const int Tn = Himg >> nw->n_blocks;
  // geometry of the activation tensors (halfs per snippet)
  int hs[5], ws[5];
  hs[0] = Himg; ws[0] = Wf;
  for (int b = 0; b < 4; ++b) { hs[b + 1] = hs[b] / 2; ws[b + 1] = (ws[b] + 1) / 2; }
  const int cp[5] = {16, FB1::OCP, FB2::OCP, FB3::OCP, FB4::OCP};
  size_t full[5], sub[5], halfs = 0;
  for (int b = 0; b < 5; ++b) {
    full[b] = (size_t)hs[b] * ws[b] * cp[b];
    sub[b] = b < 4 ? (size_t)hs[b + 1] * ws[b + 1] * cp[b] : 0;
    halfs += full[b] + sub[b];
  }
  // fp16 normalised spectrogram rows of the chunk (upper bound: non-overlapping snippets), plain or cut into block 1's strips
  halfs += (size_t)Himg * std::max<int>(conv0::kSpecLd, (((Wf + 1) / 2 + FB1C::CP - 1) / FB1C::CP) * FB1C::SPW);
  halfs = (halfs + 7) & ~(size_t)7;
  const size_t tail_f = (size_t)Tn * (nw->feat + 2 * 4 * U + 2 * U + 2 * U + 128);
  const size_t per = halfs * 2 + tail_f * 4;
  const long long chunk = std::min<long long>(std::max(nw->chunk_fused, 1), n);
  if (chunk <= 0) return ORCAI_OK;
  ORCAI_CHECK(ensure_device_buffer(c, &nw->tc_ws, &nw->tc_ws_cap, per * (size_t)chunk + 256));
  H* act[5]; H* acts[5]; H* spec16 = nullptr;
  {
    H* p = static_cast<H*>(nw->tc_ws);
    for (int b = 0; b < 5; ++b) { act[b] = p; p += full[b] * chunk; acts[b] = p; p += sub[b] * chunk; }
    spec16 = p;
  }
  float* feat = reinterpret_cast<float*>(static_cast<H*>(nw->tc_ws) + halfs * chunk);
  float* scratch = feat + (size_t)Tn * nw->feat * chunk;
  const int shift = c->p.snippet_len / 2;
  nw->mark_i = 0;
  nw->dbg_ptr = nullptr;
  const int stop = nw->debug_stop;
  {  // constant memory is per device, not per context: refresh it stream-ordered before every forward
    nw->h_conv0_pack.resize(160);
    memcpy(nw->h_conv0_pack.data(), nw->h_conv0_w.data(), 144 * sizeof(float));
    memcpy(nw->h_conv0_pack.data() + 144, nw->h_conv0_b.data(), 16 * sizeof(float));
    ORCAI_CUDA(c, cudaMemcpyToSymbolAsync(c_conv0, nw->h_conv0_pack.data(), 160 * sizeof(float), 0, cudaMemcpyHostToDevice, c->stream));
  }

  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    const long long m = std::min<long long>(chunk, n - s0);
    const bool mk = (s0 == 0);
    if (mk) nw->marked_snippets = m;
    net_mark(c, mk);
    const bool fuse0 = nw->conv0_path == 2 && stop != 0;   // the entry convolution runs inside block 1 (debug stage 0 needs its own output)
    if (fuse0) {
      // fp16 normalised rows of this chunk, cut into block 1's overlapping column strips (image columns 2*CP*s - 3 ...)
      const long long srows = input_mode == 0 ? (m - 1) * shift + Himg : m * (long long)Himg;
      const float* src = (input_mode == 0) ? d_in + (size_t)(first + s0) * shift * kRawLd : d_in + (size_t)s0 * Himg * Wf;
      const int n_strips1 = ((Wf + 1) / 2 + FB1C::CP - 1) / FB1C::CP;
      const long long total = srows * n_strips1 * FB1C::SPW;
      conv0::spec_strips_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, (long long)c->sm_count * 16), 256, 0, c->stream>>>(
          src, input_mode, input_mode == 0 ? kRawLd : Wf, srows, Wf, c->d_sel, spec16, n_strips1, 2 * FB1C::CP, -3, FB1C::SPW);
      c->launches++;
      ORCAI_CUDA(c, cudaGetLastError());
    } else if (nw->conv0_path >= 1) {
      // fp16 normalised rows of this chunk, then the pixel-group tensor-core convolution over a strided snippet view
      const long long srows = input_mode == 0 ? (m - 1) * shift + Himg : m * (long long)Himg;
      const float* src = (input_mode == 0) ? d_in + (size_t)(first + s0) * shift * kRawLd : d_in + (size_t)s0 * Himg * Wf;
      const long long total2 = srows * (conv0::kSpecLd / 2);
      conv0::spec_half_kernel<<<(unsigned)std::min<long long>((total2 + 255) / 256, (long long)c->sm_count * 16), 256, 0, c->stream>>>(
          src, input_mode, input_mode == 0 ? kRawLd : Wf, srows, Wf, c->d_sel, spec16);
      c->launches++;
      ORCAI_CUDA(c, cudaGetLastError());
      CUtensorMap tms;
      {
        EncodeTiledFn fn = encode_tiled_fn();
        if (!fn) ORCAI_FAIL(c, ORCAI_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        const cuuint64_t dims[3] = {(cuuint64_t)conv0::kSpecLd, (cuuint64_t)Himg, (cuuint64_t)m};
        const cuuint64_t strides[2] = {(cuuint64_t)conv0::kSpecLd * 2, (cuuint64_t)(input_mode == 0 ? shift : Himg) * conv0::kSpecLd * 2};
        const cuuint32_t box[3] = {(cuuint32_t)conv0::kBoxW, (cuuint32_t)conv0::kRI, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = fn(&tms, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, spec16, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) ORCAI_FAIL(c, ORCAI_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for the spectrogram view", (int)r);
      }
      const int tiles_per = (Himg + conv0::kRT - 1) / conv0::kRT;
      const long long n_tiles = m * tiles_per;
      const long long grid = std::min<long long>(n_tiles, (long long)c->sm_count * 4);
      conv0::conv0_mma_kernel<<<(unsigned)grid, 160, conv0::kSmem, c->stream>>>(tms, act[0], static_cast<H*>(nullptr), Himg, Wf, tiles_per, n_tiles,
                                                                               static_cast<const unsigned char*>(nw->conv0_mma_w));
      c->launches++;
      ORCAI_CUDA(c, cudaGetLastError());
    } else {
      const int tiles_w = (Wf + kC0TW - 1) / kC0TW, tiles_h = (Himg + c0_tile_rows(4) - 1) / c0_tile_rows(4);
      const float* src = (input_mode == 0) ? d_in : d_in + (size_t)s0 * Himg * Wf;
      conv0_direct_kernel<false, 4><<<(unsigned)(m * tiles_w * tiles_h), 256, 0, c->stream>>>(src, input_mode, first + s0, shift, input_mode == 0 ? kRawLd : Wf,
                                                                                          Himg, Wf, c->d_sel, act[0], static_cast<H*>(nullptr), tiles_w, tiles_h,
                                                                                          static_cast<H*>(nullptr), m, 0, Himg, 0);
      c->launches++;
      ORCAI_CUDA(c, cudaGetLastError());
    }
    net_mark(c, mk);  // 0: conv0
    if (stop == 0) { set_debug(nw, act[0], 1, m, hs[0], ws[0], 16, 16); return ORCAI_OK; }
    if (fuse0) ORCAI_CHECK(run_fused_block1_conv0(c, spec16, input_mode == 0 ? shift : Himg, act[1], acts[1], m, hs[0], ws[0]));
    else if (nw->block1_path == 1) ORCAI_CHECK(run_fused_block1_wide(c, act[0], act[1], acts[1], m, hs[0], ws[0]));
    else ORCAI_CHECK((run_fused_block<FB1>(c, 0, act[0], static_cast<const H*>(nullptr), act[1], acts[1], m, hs[0], ws[0])));
    net_mark(c, mk);  // 1
    if (stop == 1) { set_debug(nw, act[1], 1, m, hs[1], ws[1], 30, cp[1]); return ORCAI_OK; }
    if (stop == 21) { set_debug(nw, acts[1], 1, m, hs[2], ws[2], 30, cp[1]); return ORCAI_OK; }
    ORCAI_CHECK((run_fused_block<FB2>(c, 1, act[1], acts[1], act[2], acts[2], m, hs[1], ws[1])));
    net_mark(c, mk);  // 2
    if (stop == 2) { set_debug(nw, act[2], 1, m, hs[2], ws[2], 40, cp[2]); return ORCAI_OK; }
    if (stop == 22) { set_debug(nw, acts[2], 1, m, hs[3], ws[3], 40, cp[2]); return ORCAI_OK; }
    ORCAI_CHECK((run_fused_block<FB3>(c, 2, act[2], acts[2], act[3], acts[3], m, hs[2], ws[2])));
    net_mark(c, mk);  // 3
    if (stop == 3) { set_debug(nw, act[3], 1, m, hs[3], ws[3], 50, cp[3]); return ORCAI_OK; }
    if (stop == 23) { set_debug(nw, acts[3], 1, m, hs[4], ws[4], 50, cp[3]); return ORCAI_OK; }
    ORCAI_CHECK((run_fused_block<FB4>(c, 3, act[3], acts[3], act[4], static_cast<H*>(nullptr), m, hs[3], ws[3])));
    net_mark(c, mk);  // 4
    if (stop == 4) { set_debug(nw, act[4], 1, m, hs[4], ws[4], 60, cp[4]); return ORCAI_OK; }
    ORCAI_CHECK((run_sep<60, 36, false, true, H, true>(c, act[4], feat, m, hs[4], ws[4], nw->fb_fin)));
    net_mark(c, mk);  // 5: final sepconv (fp32 features, w*36+c)
    if (stop == 5) { set_debug(nw, feat, 0, m, hs[4], ws[4], 36, 36); return ORCAI_OK; }
    if (nw->tail_path == 1) ORCAI_CHECK(net_tail_tc(c, feat, scratch, m, d_preds + (size_t)s0 * Tn * L, mk));
    else ORCAI_CHECK(net_tail_fp32(c, feat, scratch, m, d_preds + (size_t)s0 * Tn * L, mk));
  }
  return ORCAI_OK;
}

#include "net_precise.cuh"

}  // namespace

// Channel means of every GEMM's A operand on the first snippets of the resident recording (layer-wise fp16 path), see Calib.
int net_calibrate(Ctx* c, int64_t max_snippets) {
  NetWeights* nw = c->net;
  if (!nw->loaded) ORCAI_FAIL(c, ORCAI_ERR_STATE, "no weights loaded (orcai_load_weights)");
  if (!c->have_stats) ORCAI_FAIL(c, ORCAI_ERR_STATE, "no spectrogram resident (orcai_spectrogram_resident)");
  const int64_t N = orcai_num_snippets(c->T, c->p.snippet_len);
  const int64_t K = std::min<int64_t>(N, max_snippets > 0 ? max_snippets : 8);
  if (K <= 0) ORCAI_FAIL(c, ORCAI_ERR_TOO_SHORT, "calibration recording is shorter than one snippet");
  ORCAI_CHECK(net_tc_prepare(c, 0));
  const int Himg = nw->H, Wf = nw->Wf, U = nw->U;
  const int Tn = Himg >> nw->n_blocks, shift = c->p.snippet_len / 2;
  constexpr int SL = CalibSink::kCalSlot;
  double* d_sums = nullptr;
  ORCAI_CUDA(c, cudaMalloc(&d_sums, sizeof(double) * SL * CalibSink::NSLOTS));
  ORCAI_CUDA(c, cudaMemsetAsync(d_sums, 0, sizeof(double) * SL * CalibSink::NSLOTS, c->stream));
  const CalibSink sink{d_sums, c};
  const long long srows = (K - 1) * shift + Himg;
  spec_sum_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->d_raw, srows, kRawLd, Wf, c->d_sel, d_sums + (size_t)CalibSink::SPEC * SL);
  float* d_tmp = nullptr;
  ORCAI_CUDA(c, cudaMalloc(&d_tmp, sizeof(float) * (size_t)K * Tn * nw->L));
  const int chunk_saved = nw->chunk, stop_saved = nw->debug_stop;
  nw->chunk = (int)std::max<int64_t>(nw->chunk, K);   // one chunk: the sums must cover every calibration snippet exactly once
  nw->debug_stop = -1;
  int rc = forward_tc<__half>(c, c->d_raw, 0, 0, K, d_tmp, 0, &sink);
  nw->chunk = chunk_saved;
  nw->debug_stop = stop_saved;
  std::vector<double> hs((size_t)SL * CalibSink::NSLOTS);
  if (rc == ORCAI_OK) {
    cudaError_t e = cudaMemcpyAsync(hs.data(), d_sums, hs.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { c->err = cudaGetErrorString(e); rc = ORCAI_ERR_CUDA; }
  }
  cudaFree(d_tmp);
  cudaFree(d_sums);
  if (rc != ORCAI_OK) return rc;
  Calib& cal = nw->calib;
  auto mean = [&](int slot, int C, double count, std::vector<double>* out) {
    out->assign(C, 0.0);
    for (int i = 0; i < C; ++i) (*out)[i] = hs[(size_t)slot * SL + i] / count;
  };
  int h = Himg, w = Wf, ci = 16;
  for (int b = 0; b < nw->n_blocks; ++b) {
    mean(CalibSink::IN_RELU + b, ci, (double)K * h * w, &cal.in_relu[b]);
    mean(CalibSink::IN_EVEN + b, ci, (double)K * ((h + 1) / 2) * ((w + 1) / 2), &cal.in_even[b]);
    mean(CalibSink::S1 + b, nw->filters[b], (double)K * h * w, &cal.s1[b]);
    ci = nw->filters[b];
    h = (h + 1) / 2; w = (w + 1) / 2;
  }
  mean(CalibSink::FIN, ci, (double)K * h * w, &cal.fin);
  mean(CalibSink::FEAT, nw->feat, (double)K * Tn, &cal.feat);
  mean(CalibSink::H1, 2 * U, (double)K * Tn, &cal.h1);
  mean(CalibSink::H2, 2 * U, (double)K * Tn, &cal.h2);
  cal.spec = hs[(size_t)CalibSink::SPEC * SL] / ((double)srows * Wf);
  cal.valid = true;
  nw->fused_ready = false;     // operands are re-packed with the corrected biases on the next forward
  nw->tail_tc_ready = false;
  return ORCAI_OK;
}

const unsigned int* net_trap_info() { return g_trap_host; }

#ifdef ORCAI_FUSED_TRACE
// called after a stream synchronise: append the recorded hand-off stamps to $ORCAI_B200_TRACE and clear them
void net_trace_dump() {
  const char* path = getenv("ORCAI_B200_TRACE");
  if (!path || !g_trace_dev) return;
  std::vector<long long> h(8 * 256);
  if (cudaMemcpy(h.data(), g_trace_dev, h.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return;
  cudaMemset(g_trace_dev, 0, h.size() * 8);
  if (FILE* f = fopen(path, "a")) {
    for (size_t i = 0; i < h.size(); ++i)
      if (h[i]) fprintf(f, "%lld %lld\n", (long long)(((i / 256) * 10) << 32 | (40 + (i % 256) / 64) << 8 | (i % 64)), h[i]);
    fprintf(f, "-1 -1\n");
    fclose(f);
  }
}
#else
void net_trace_dump() {}
#endif

int net_tc_prepare(Ctx* c, int fmt) {
  NetWeights* nw = c->net;
  if (!nw->loaded) ORCAI_FAIL(c, ORCAI_ERR_STATE, "no weights loaded (orcai_load_weights)");
  if (nw->tc_ready[fmt]) return ORCAI_OK;
  return fmt == 0 ? prepare<__half>(c, 0) : prepare<__nv_bfloat16>(c, 1);
}

int net_forward_tc(Ctx* c, const float* d_in, int input_mode, int64_t first, int64_t n, float* d_preds) {
  NetWeights* nw = c->net;
  if (nw->path == 3) {
    ORCAI_CHECK(prepare_fused(c));
    return forward_fused(c, d_in, input_mode, first, n, d_preds);
  }
  if (nw->path == 4) {
    ORCAI_CHECK(prepare_precise(c));
    return forward_precise(c, d_in, input_mode, first, n, d_preds);
  }
  const int fmt = nw->path == 2 ? 1 : 0;
  ORCAI_CHECK(net_tc_prepare(c, fmt));
  return fmt == 0 ? forward_tc<__half>(c, d_in, input_mode, first, n, d_preds, 0)
                  : forward_tc<__nv_bfloat16>(c, d_in, input_mode, first, n, d_preds, 1);
}

}  // namespace orcai
