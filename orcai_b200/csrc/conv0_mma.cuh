// Entry convolution of orcai-V1 (Conv2D 3x3 "same", 1 -> 16 channels, + folded BatchNorm + ReLU; reference
// src/orcAI/architectures.py:162-168) on the tensor cores, for the fused path.  Included by net_tc.cu.
//
// A single-channel image has no channel dimension to contract over, so the GEMM's K dimension is made of PIXELS: the
// normalised spectrogram is kept as fp16 with 8 consecutive frequency bins = one 16-byte "pixel group".  One A row of the
// MMA is one pixel group (8 K-elements); rows are consecutive groups of the pixel-linear shared-memory tile (SBO = 128 B per
// 8 groups) and the second K-chunk of a K=16 step is simply the NEXT group (LBO = 16 B).  For every vertical tap dy the
// output group g needs groups g-1, g, g+1 of input row r+dy, i.e. two K=16 steps [g-1 | g] and [g+1 | (zero weights)];
// B holds the 3x3 kernel as a banded (Toeplitz) matrix  B[(j, c)][k] = k0[dy][pos(k) - j + 1][c]  that produces all
// 8 pixels x 16 channels = 128 accumulator columns of the group at once.  6 MMAs (+1 for the bias, ones-operand trick)
// per 120 x 8 pixels; the 10x redundant MACs are free on an otherwise idle tensor pipe, the kernel is bound by its
// 32 B/pixel NHWC store.  Borders: the TMA box (192 x 7) is zero-filled outside the snippet (rows) and outside the band
// (columns); snippets are addressed through a tensor map whose outer stride is the snippet shift (368 rows), so the
// overlapping windows are never materialised (reference predict.py:252-261 copies them).
#pragma once

namespace conv0 {

constexpr int kSpecLd = 176;                 // fp16 row pitch of the normalised spectrogram (22 groups, pad columns zero)
constexpr int kBoxW = 192, kGR = kBoxW / 8;  // tile row: image columns -8 .. 183 = 24 pixel groups
constexpr int kRT = 5, kRI = kRT + 2;        // output rows per tile, input rows
constexpr uint32_t kTileBytes = kRI * kBoxW * 2;            // 2688
constexpr uint32_t kBufBytes = 128 + 2688 + 384;            // zero pad in front (one group is read), tile, zero slack behind
constexpr uint32_t kBBlock = 128 * 16 * 2;                  // one K=16 step of B: 128 rows x 16 k, canonical (SBO 256, LBO 128)
constexpr uint32_t kStage = 32 * 128;                       // per worker warp: 32 pixel groups x 4 pixels x 32 B, XOR-swizzled 16-B pieces
constexpr uint32_t OFF_B = 0, OFF_ONES = 7 * kBBlock, OFF_BUF = OFF_ONES + 256, OFF_STAGE = OFF_BUF + 2 * kBufBytes, OFF_BAR = OFF_STAGE + 4 * kStage;
constexpr uint32_t kSmem = OFF_BAR + 5 * 8 + 16;            // barriers: full[2], acc, done
constexpr uint32_t kWBytes = OFF_BUF;                       // host-packed: B blocks + ones tile

// raw dB (or already normalised snippets) -> fp16 normalised spectrogram rows with pitch kSpecLd, pad columns zero
__global__ void __launch_bounds__(256)
spec_half_kernel(const float* __restrict__ in, int mode, int in_ld, long long rows, int nb, const orcai::SelectState* __restrict__ st,
                 __half* __restrict__ out) {
  float db_ref = 0.f, lo = 0.f, hi = 1.f, range = 1.f;
  if (mode == 0) { db_ref = st->db_ref; lo = st->lo; hi = st->hi; range = hi - lo; }
  const long long total = rows * (kSpecLd / 2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (kSpecLd / 2);
    const int c = (int)(i - r * (kSpecLd / 2)) * 2;
    float v[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float x = 0.f;
      if (c + q < nb) {
        x = __ldg(in + (size_t)r * in_ld + c + q);
        if (mode == 0) {
          x = fmaxf(x - db_ref, -80.0f);
          x = __fdiv_rn(fminf(fmaxf(x, lo), hi) - lo, range);
        }
      }
      v[q] = x;
    }
    reinterpret_cast<__half2*>(out)[i] = __floats2half2_rn(v[0], v[1]);
  }
}

// The same rows cut into overlapping column strips for the entry convolution fused into block 1 (net_fused.cuh, CONV0):
// out[row][strip][x] = spectrogram[row][strip * strip_step + col0 + x] (zero outside 0 .. nb-1), x < strip_w.  Every strip
// starts a 128-byte line, so the TMA boxes that fetch them have an aligned innermost coordinate (0).
__global__ void __launch_bounds__(256)
spec_strips_kernel(const float* __restrict__ in, int mode, int in_ld, long long rows, int nb, const orcai::SelectState* __restrict__ st,
                   __half* __restrict__ out, int n_strips, int strip_step, int col0, int strip_w) {
  float db_ref = 0.f, lo = 0.f, hi = 1.f, range = 1.f;
  if (mode == 0) { db_ref = st->db_ref; lo = st->lo; hi = st->hi; range = hi - lo; }
  const int per_row = n_strips * strip_w;
  const long long total = rows * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / per_row;
    const int q = (int)(i - r * per_row);
    const int s = q / strip_w, x = q - s * strip_w;
    const int c = s * strip_step + col0 + x;
    float v = 0.f;
    if (c >= 0 && c < nb) {
      v = __ldg(in + (size_t)r * in_ld + c);
      if (mode == 0) {
        v = fmaxf(v - db_ref, -80.0f);
        v = __fdiv_rn(fminf(fmaxf(v, lo), hi) - lo, range);
      }
    }
    out[i] = __float2half_rn(v);
  }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               :
               : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(orcai::tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// grid: persistent, 4 CTAs per SM; block: 4 worker warps (TMEM drain + store) + 1 control warp (TMA + MMA issue)
__global__ void __launch_bounds__(160, 4)
conv0_mma_kernel(const __grid_constant__ CUtensorMap tmS, __half* __restrict__ out, __half* __restrict__ out_sub, int Himg, int Wimg,
                 int tiles_per, long long n_tiles, const unsigned char* __restrict__ wpack) {
  using namespace orcai::tc;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);   // full[0], full[1], acc, done
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (int)(kWBytes / 16); i += 160) reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(wpack) + i);
  for (int i = tid + kWBytes / 16; i < (int)(OFF_STAGE / 16); i += 160) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], 4);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc<128>(tslot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t sbase = smem_u32(smem);
  const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 4) {
    constexpr uint32_t idesc = make_idesc_f16(128, 128, 0);
    const uint64_t dOnes = make_smem_desc(sbase + OFF_ONES, 128, 0);
    const uint64_t dB = make_smem_desc(sbase + OFF_B, 128, 256);
    auto load = [&](long long i) {   // tile i of this CTA -> buffer i & 1
      const long long tile = blockIdx.x + i * gridDim.x;
      const long long b = tile / tiles_per;
      const int h0 = (int)(tile - b * tiles_per) * kRT;
      uint64_t* bar = &bars[i & 1];
      fused::mbar_arrive_expect_tx(bar, kTileBytes);
      tma_load_3d(sbase + OFF_BUF + (uint32_t)(i & 1) * kBufBytes + 128, &tmS, bar, -8, h0 - 1, (int)b);
    };
    if (my_tiles > 0 && fused::elect_one()) load(0);
    __syncwarp();
    for (long long i = 0; i < my_tiles; ++i) {
      mbar_wait(&bars[i & 1], (uint32_t)((i >> 1) & 1));
      if (i > 0) mbar_wait(&bars[3], (uint32_t)((i - 1) & 1));   // the workers have drained the previous accumulator
      tc_fence_after();
      if (fused::elect_one()) {
        // A rows = pixel groups of the tile in linear order; tap (dy, pair) starts at group dy*24 - 1 + 2*pair
        const uint32_t tile0 = sbase + OFF_BUF + (uint32_t)(i & 1) * kBufBytes + 128 - 16;
        mma_f16_ss(tmem, dOnes, dB + ((6 * kBBlock) >> 4), idesc, 0);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int pr = 0; pr < 2; ++pr)
            mma_f16_ss(tmem, make_smem_desc(tile0 + (uint32_t)(dy * kGR + 2 * pr) * 16, 16, 128), dB + (((dy * 2 + pr) * kBBlock) >> 4), idesc, 1);
        mma_commit(&bars[2]);
        if (i + 1 < my_tiles) load(i + 1);   // the other buffer: its last readers (tile i-1's MMAs) completed before `done`
      }
      __syncwarp();
    }
  } else {
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    const int Ho = Himg >> 1, Wo = (Wimg + 1) >> 1;
    for (long long i = 0; i < my_tiles; ++i) {
      const long long tile = blockIdx.x + i * gridDim.x;
      const long long b = tile / tiles_per;
      mbar_wait(&bars[2], (uint32_t)(i & 1));
      tc_fence_after();
      // The accumulator row of a lane is 8 pixels x 32 B = 256 contiguous bytes of the NHWC output, 256 B apart from the
      // next lane's.  Stage 4 pixels at a time through shared memory so that every store instruction writes 32 consecutive
      // 16-byte pieces (full 32-byte sectors, 512 contiguous bytes per warp instruction).
      unsigned char* stg = smem + OFF_STAGE + warp * kStage;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float v[16];
          tmem_ld16(lane_addr + (hf * 4 + jj) * 16, v);
          const uint4 v0 = fused::relu8h(fused::pack8h(v)), v1 = fused::relu8h(fused::pack8h(v + 8));
          *reinterpret_cast<uint4*>(stg + lane * 128 + (((2 * jj) ^ (lane & 7)) << 4)) = v0;
          *reinterpret_cast<uint4*>(stg + lane * 128 + (((2 * jj + 1) ^ (lane & 7)) << 4)) = v1;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int e = q * 32 + lane;                 // 16-byte piece of this warp's 32 groups x 4 pixels
          const int gl = e >> 3, part = e & 7;         // source lane (group), piece inside its 128 bytes
          const int mm = warp * 32 + gl;
          const int rr = mm / kGR, gg = mm - rr * kGR;
          const int h2 = (int)(tile - b * tiles_per) * kRT + rr;
          const int ww = 8 * (gg - 1) + hf * 4 + (part >> 1);
          const uint4 val = *reinterpret_cast<const uint4*>(stg + gl * 128 + ((part ^ (gl & 7)) << 4));
          if (rr < kRT && h2 < Himg && gg >= 1 && ww < Wimg) {
            __half* o = out + (((size_t)b * Himg + h2) * Wimg + ww) * 16 + (part & 1) * 8;
            *reinterpret_cast<uint4*>(o) = val;
            if (out_sub != nullptr && !(h2 & 1) && !(ww & 1)) {   // optional even-position copy
              __half* os = out_sub + (((size_t)b * Ho + (h2 >> 1)) * Wo + (ww >> 1)) * 16 + (part & 1) * 8;
              *reinterpret_cast<uint4*>(os) = val;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) fused::mbar_arrive(&bars[3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem);
}

}  // namespace conv0
