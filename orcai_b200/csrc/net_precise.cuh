// orcai-V1 forward at fp32 grade on the fp16 tensor cores (net_path 4; included by net_tc.cu).
//
// Reference graph: src/orcAI/architectures.py:120-241, called from predict.py:266-268; parity gate: probabilities within 1e-3.
// fp16 operands alone miss that gate (2.7e-3 over an hour of audio): tools/precision_plan.py attributes 2e-4 .. 2e-3 to EVERY
// weight tensor and EVERY stored activation, the fp16 spectrogram and the entry convolution's weights most of all.  So every
// tensor-core product here is the three-term split  A_hi*W_hi + A_lo*W_hi + A_hi*W_lo  with (hi, lo) = (fp16(v), fp16(v - hi)),
// accumulated in fp32 (the dropped term is 2^-22 relative), and everything else is fp32 CUDA-core arithmetic:
//
//   entry convolution   fp32 FFMA from the fp32 normalised spectrogram (conv0_direct_kernel<true>) -> (hi, lo) fp16 NHWC
//   block 1             fused::FB<..., PREC = true>: the fused residual-block kernel with (hi, lo) operand plane sets -> fp32 NHWC
//   blocks 2 - 4, final un-folded: depthwise 3x3 in fp32 (dw3x3_kernel) + pointwise 1x1 as a split GEMM over (pixels x channels)
//                       (gemm_tc_kernel<.., SPLIT>), max-pool + residual 1x1/2 in fp32 (pool_res_f32_kernel).  The folded
//                       form would need both weight sets (hi, lo) of nine taps resident: 112 / 207 / 295 KB for blocks 2 / 3 / 4.
//   LSTM / dense tail   split GEMMs for the input projections and Dense(128), fp32 recurrence (net_tail_precise)
//
// Activations between these kernels are fp32 NHWC with the channel pitch padded to a multiple of 8 (padding channels are zero).
#pragma once

namespace precise {

// depthwise 3x3, "same" zero padding: d[p][c] = sum_taps dw[tap][c] * f(x[p + tap][c]), f = ReLU or identity.
// x, d: (n, H, W, CP) fp32; dw: [9][CP].  One thread per (pixel, 4 channels).
template <bool RELU_IN>
__global__ void __launch_bounds__(256)
dw3x3_kernel(const float* __restrict__ x, float* __restrict__ d, const float* __restrict__ dw, long long n, int H, int W, int CP) {
  const int G4 = CP >> 2;
  const long long total = n * H * W * G4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G4);
    long long r = idx / G4;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const long long b = r / H;
    const float* xb = x + (size_t)b * H * W * CP + 4 * g;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int hh = h + dy - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int ww = w + dx - 1;
        if (ww < 0 || ww >= W) continue;
        float4 v = __ldg(reinterpret_cast<const float4*>(xb + ((size_t)hh * W + ww) * CP));
        if (RELU_IN) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        const float4 k = __ldg(reinterpret_cast<const float4*>(dw + (dy * 3 + dx) * CP + 4 * g));
        acc.x = fmaf(v.x, k.x, acc.x); acc.y = fmaf(v.y, k.y, acc.y); acc.z = fmaf(v.z, k.z, acc.z); acc.w = fmaf(v.w, k.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(d + (size_t)idx * 4) = acc;
  }
}

// MaxPool (3,2)/2 "same" (-inf beyond the image) of s2 (n, H, W, COP)  +  Conv1x1/2 of the block input  ->  y (n, Ho, Wo, COP).
// The block input at even positions is addressed as xs + b * xs_img + ho * xs_row + wo * xs_px (floats): its own sub-sampled
// tensor, or the full tensor walked with stride 2.  rw: [CIP][COP] (zero padded), rb: [COP].  One thread per (kPoolPx consecutive
// pooled pixels of a row, 4 channels): every residual weight read from shared memory serves kPoolPx pixels (with one pixel per
// thread the kernel was bound by those reads: 32 LDS.128 for 128 FMAs).
constexpr int kPoolPx = 4;
__global__ void __launch_bounds__(256)
pool_res_f32_kernel(const float* __restrict__ s2, const float* __restrict__ xs, long long xs_img, long long xs_row, int xs_px,
                    float* __restrict__ y, long long n, int H, int W, int Ho, int Wo, int CIP, int COP,
                    const float* __restrict__ rw, const float* __restrict__ rb) {
  extern __shared__ float s_w[];   // [CIP][COP] then [COP]
  for (int i = threadIdx.x; i < CIP * COP; i += blockDim.x) s_w[i] = rw[i];
  for (int i = threadIdx.x; i < COP; i += blockDim.x) s_w[CIP * COP + i] = rb[i];
  __syncthreads();
  const int G4 = COP >> 2;
  const int WB = (Wo + kPoolPx - 1) / kPoolPx;            // pixel groups per pooled row
  const long long total = n * Ho * WB * G4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G4);
    long long r = idx / G4;
    const int wb = (int)(r % WB); r /= WB;
    const int ho = (int)(r % Ho);
    const long long b = r / Ho;
    const float* sb = s2 + (size_t)b * H * W * COP + 4 * g;
    const float4 bias = *reinterpret_cast<const float4*>(s_w + CIP * COP + 4 * g);
    float4 m[kPoolPx], a[kPoolPx];
    const float* xp[kPoolPx];
    // The window loads are branch-free: a row / column beyond the image is clamped onto the image's last one, which the window
    // holds anyway (the maximum does not change), so all twelve loads of a pixel pair are in flight together.  (With "continue"
    // on the borders the loads were issued two at a time and the kernel waited on memory latency 57 % of its time.)
    const int hr0 = 2 * ho, hr1 = min(2 * ho + 1, H - 1), hr2 = min(2 * ho + 2, H - 1);
    const float* const r0p = sb + (size_t)hr0 * W * COP;
    const float* const r1p = sb + (size_t)hr1 * W * COP;
    const float* const r2p = sb + (size_t)hr2 * W * COP;
#pragma unroll
    for (int pp = 0; pp < kPoolPx; pp += 2) {
      float4 v[2][6];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int wo = min(wb * kPoolPx + pp + q, Wo - 1);
        const int c0 = 2 * wo * COP, c1 = min(2 * wo + 1, W - 1) * COP;
        v[q][0] = __ldg(reinterpret_cast<const float4*>(r0p + c0));
        v[q][1] = __ldg(reinterpret_cast<const float4*>(r0p + c1));
        v[q][2] = __ldg(reinterpret_cast<const float4*>(r1p + c0));
        v[q][3] = __ldg(reinterpret_cast<const float4*>(r1p + c1));
        v[q][4] = __ldg(reinterpret_cast<const float4*>(r2p + c0));
        v[q][5] = __ldg(reinterpret_cast<const float4*>(r2p + c1));
        xp[pp + q] = xs + (size_t)b * xs_img + (size_t)ho * xs_row + (size_t)wo * xs_px;
        a[pp + q] = bias;
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float4 mm = v[q][0];
#pragma unroll
        for (int t = 1; t < 6; ++t) { mm.x = fmaxf(mm.x, v[q][t].x); mm.y = fmaxf(mm.y, v[q][t].y); mm.z = fmaxf(mm.z, v[q][t].z); mm.w = fmaxf(mm.w, v[q][t].w); }
        m[pp + q] = mm;
      }
    }
    float4 xv[kPoolPx];
#pragma unroll
    for (int p = 0; p < kPoolPx; ++p) xv[p] = __ldg(reinterpret_cast<const float4*>(xp[p]));
    for (int ci = 0; ci < CIP; ci += 4) {
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + (ci + 0) * COP + 4 * g);
      const float4 w1 = *reinterpret_cast<const float4*>(s_w + (ci + 1) * COP + 4 * g);
      const float4 w2 = *reinterpret_cast<const float4*>(s_w + (ci + 2) * COP + 4 * g);
      const float4 w3 = *reinterpret_cast<const float4*>(s_w + (ci + 3) * COP + 4 * g);
      float4 xn[kPoolPx];
      const int cn = ci + 4 < CIP ? ci + 4 : ci;           // the next quad of input channels is requested before this one is used
#pragma unroll
      for (int p = 0; p < kPoolPx; ++p) xn[p] = __ldg(reinterpret_cast<const float4*>(xp[p] + cn));
#pragma unroll
      for (int p = 0; p < kPoolPx; ++p) {
        const float4 x4 = xv[p];
        // per output channel the same FMA order as with one pixel per thread: ci ascending
        // FFMA2 (two output channels per instruction, the pixel value as the broadcast operand); each lane rounds like fmaf
        fused::fma2(a[p].x, a[p].y, w0.x, w0.y, x4.x, x4.x); fused::fma2(a[p].z, a[p].w, w0.z, w0.w, x4.x, x4.x);
        fused::fma2(a[p].x, a[p].y, w1.x, w1.y, x4.y, x4.y); fused::fma2(a[p].z, a[p].w, w1.z, w1.w, x4.y, x4.y);
        fused::fma2(a[p].x, a[p].y, w2.x, w2.y, x4.z, x4.z); fused::fma2(a[p].z, a[p].w, w2.z, w2.w, x4.z, x4.z);
        fused::fma2(a[p].x, a[p].y, w3.x, w3.y, x4.w, x4.w); fused::fma2(a[p].z, a[p].w, w3.z, w3.w, x4.w, x4.w);
        xv[p] = xn[p];
      }
    }
#pragma unroll
    for (int p = 0; p < kPoolPx; ++p) {
      const int wo = wb * kPoolPx + p;
      if (wo < Wo)
        *reinterpret_cast<float4*>(y + ((((size_t)b * Ho + ho) * Wo + wo) * COP + 4 * g)) = make_float4(m[p].x + a[p].x, m[p].y + a[p].y, m[p].z + a[p].z, m[p].w + a[p].w);
    }
  }
}

// Border images of a level (tall-image mode, see forward_precise): image b < n is rows [0, 8) of snippet b, image n + b rows
// [H - 8, H) of snippet b.  A row comes from the snippet's own border values `alt` (the previous level's border outputs: 4 rows
// per image; rows [0, nt) of a top image, the last nb rows of a bottom image) or from the tall tensor (snippet i starts at tall
// row i * stride).  tall: (rows, W, C), alt: (2n, 4, W, C), out: (2n, 8, W, C); one thread per float4.
__global__ void __launch_bounds__(256)
gather_border_kernel(const float* __restrict__ tall, const float* __restrict__ alt, float* __restrict__ out, long long n, int stride, int H, int row_f4,
                     int nt, int nb) {
  const long long total = 2 * n * 8 * row_f4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(idx % row_f4);
    long long r = idx / row_f4;
    const int j = (int)(r % 8);
    const long long img = r / 8;
    const bool top = img < n;
    const long long i = top ? img : img - n;
    const float4* src;
    if (top) src = j < nt ? reinterpret_cast<const float4*>(alt) + (img * 4 + j) * row_f4 : reinterpret_cast<const float4*>(tall) + (i * stride + j) * row_f4;
    else src = j >= 8 - nb ? reinterpret_cast<const float4*>(alt) + (img * 4 + (j - 4)) * row_f4
                           : reinterpret_cast<const float4*>(tall) + (i * stride + H - 8 + j) * row_f4;
    reinterpret_cast<float4*>(out)[idx] = __ldg(src + q);
  }
}

// features of snippet i, row r (of Tn): the tall tensor's row i * stride + r, except the nt first / nb last rows, which are the
// snippet's own (border images of 8 rows: top image rows [0, 8), bottom image rows [Tn - 8, Tn))
__global__ void __launch_bounds__(256)
assemble_feat_kernel(const float* __restrict__ tall, const float* __restrict__ border, float* __restrict__ out, long long n, int stride, int Tn, int row_f4,
                     int nt, int nb) {
  const long long total = n * Tn * row_f4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(idx % row_f4);
    long long r = idx / row_f4;
    const int j = (int)(r % Tn);
    const long long i = r / Tn;
    const float4* src;
    if (j < nt) src = reinterpret_cast<const float4*>(border) + (i * 8 + j) * row_f4;
    else if (j >= Tn - nb) src = reinterpret_cast<const float4*>(border) + ((n + i) * 8 + (j - (Tn - 8))) * row_f4;
    else src = reinterpret_cast<const float4*>(tall) + (i * stride + j) * row_f4;
    reinterpret_cast<float4*>(out)[idx] = __ldg(src + q);
  }
}


// ------------------------------------------------------------------------------------------------
// One un-folded separable convolution as ONE kernel: depthwise 3x3 on the CUDA cores straight into the A operand of the
// pointwise split GEMM.   out[p][n] = act( sum_k d[p][k] * PW[k][n] + bias[n] ),  d[p][k] = sum_taps dw[tap][k] * f(x[p + tap][k])
//
//   tile      16 x 8 pixels = the 128 rows of one tcgen05.mma; its 18 x 10 halo arrives as ONE TMA box of the fp32 NHWC tensor
//             ([18][10][CP] floats; image borders are zero-filled by the TMA unit = the "same" padding)
//   workers   4 warps.  Thread t < 8 * Q (Q = CP / 4 channel quads) owns tile column t / Q and channel quad t % Q and walks the
//             rows with a sliding 3 x 3 window in registers (eight rows per unrolled trip): 3 conflict-free LDS.128 + 36 FFMA per
//             output float4; the results go to the (hi, lo) A planes (canonical K-major, plane pitch padded by 16 B against bank
//             conflicts).  (Dealing the (column, row) pairs out evenly to all lanes at Q = 10 was measured slower: the SM's three
//             CTAs are bound by the instructions they issue together, and that mapping issues more of them.)
//   issuer    1 warp: the TMA of the NEXT tile's halo the moment the workers have consumed the current one (before this tile's
//             MMAs are issued), then 3 MMAs per K step (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo, N = 64, fp32 accumulation in TMEM)
//   epilogue  the workers drain TMEM (thread = pixel), add the bias, apply the ReLU, store fp32 NHWC rows
// Three CTAs per SM overlap each other's phases.  x: (n, H, W, CP), out: (n, H, W, ldc); CP <= 64, ldc <= 64.
constexpr int kUfTH = 16, kUfTW = 8, kUfHalo = (kUfTH + 2) * (kUfTW + 2);
constexpr uint32_t kUfLboA = 128 * 16 + 16;                           // A plane pitch (one 8-channel chunk of 128 rows) + pad
// shared memory sized by the channel count: 56 KB at 32 channels (the register file then allows three CTAs per SM), 95 KB at 64
struct UfSmem {
  static constexpr uint32_t B_HALF = 64 * 64 * 2;                      // [hi | lo] 64 x 64 fp16 blocks
  uint32_t halo_bytes, off_a, a_half, off_b, off_bar, bytes;           // halo [18][10][CP] fp32 at 0; hi planes, lo planes; B; barriers
  __host__ __device__ explicit UfSmem(int cp) {
    const uint32_t planes = 2u * (uint32_t)((cp + 15) / 16);           // 8-channel planes the MMAs read (K steps of 16)
    halo_bytes = ((uint32_t)kUfHalo * cp * 4 + 127) & ~127u;
    off_a = halo_bytes;
    a_half = planes * kUfLboA;
    off_b = (off_a + 2 * a_half + 127) & ~127u;
    off_bar = off_b + 2 * B_HALF;                                      // halo_full, a_full, acc_full, acc_free, b_full; then 64 bias floats
    bytes = off_bar + 128 + 64 * 4;
  }
};

#ifdef ORCAI_FUSED_TRACE
#define UF_TRACE(tag, g) do { if (CP == 40 && !RELU_IN && ACT == 0 && tiles_h > 1000) fused::trace_event(70, tag, g); } while (0)
#else
#define UF_TRACE(tag, g)
#endif

template <bool RELU_IN, int ACT>
__global__ void __launch_bounds__(160, 3)   // four CTAs per SM (96 registers, spills) measured 9 % slower; a second halo buffer (two CTAs per SM) 8 % slower
sep_uf_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ dw, const __half* __restrict__ Bp, const float* __restrict__ bias,
              float* __restrict__ out, long long n_img, int H, int W, int CP, int ldc, int n_valid, int tiles_w, int tiles_h) {
  const UfSmem L(CP);
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 8);
  float* s_bias = reinterpret_cast<float*>(smem + L.off_bar + 128);     // the epilogue's bias reads stay off the global-memory path
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Q = CP >> 2;
  if (tid < 64) s_bias[tid] = __ldg(bias + tid);
  if (tid == 0) {
    mbar_init(&bars[0], 1);      // halo_full: TMA complete_tx
    mbar_init(&bars[2], 4);      // a_full: one arrival per worker warp = A planes written AND halo consumed
    mbar_init(&bars[3], 1);      // acc_full: tcgen05.commit
    mbar_init(&bars[4], 4);      // acc_free
    mbar_init(&bars[5], 1);      // b_full
    fence_mbar_init();
  }
  for (int i = tid; i < (int)(2 * L.a_half / 16); i += 160) reinterpret_cast<uint4*>(smem + L.off_a)[i] = make_uint4(0, 0, 0, 0);   // unused K planes stay zero
  __syncwarp();
  if (warp == 0) tmem_alloc<64>(tslot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t tiles_per = (uint32_t)tiles_w * (uint32_t)tiles_h;
  const uint32_t total = (uint32_t)n_img * tiles_per;                      // < 2^31 (checked by the launcher)

  if (warp == 4) {
    // =============================== issuer: TMA + MMA ===============================
    if (fused::elect_one()) {
      fused::mbar_arrive_expect_tx(&bars[5], 2 * UfSmem::B_HALF);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sbase + L.off_b), "l"(Bp),
                   "r"(2 * UfSmem::B_HALF), "r"(smem_u32(&bars[5]))
                   : "memory");
    }
    __syncwarp();
    constexpr uint32_t idesc = make_idesc_f16(128, 64, 0);
    const uint64_t da = make_smem_desc(sbase + L.off_a, kUfLboA, 128);
    const uint64_t db = make_smem_desc(sbase + L.off_b, 128, 8 * 128);
    const int ksteps = (CP + 15) >> 4;
    const uint32_t halo_tx = (uint32_t)kUfHalo * CP * 4;
    auto load_halo = [&](uint32_t tl) {
      const uint32_t b2 = tl / tiles_per, tr2 = tl - b2 * tiles_per;
      const uint32_t ty = tr2 / (uint32_t)tiles_w, tx = tr2 - ty * (uint32_t)tiles_w;
      if (fused::elect_one()) {
        fused::mbar_arrive_expect_tx(&bars[0], halo_tx);
        fused::tma_load_4d(sbase, &tmX, &bars[0], 0, (int)tx * kUfTW - 1, (int)ty * kUfTH - 1, (int)b2);
      }
      __syncwarp();
    };
    uint32_t it = 0;
    if (blockIdx.x < total) { load_halo(blockIdx.x); UF_TRACE(1, 0); }
    for (uint32_t tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      if (it == 0) mbar_wait(&bars[5], 0);
      mbar_wait(&bars[2], it & 1);                                         // A planes of this tile written, its halo consumed
      UF_TRACE(2, it);
      // the next tile's halo is requested BEFORE this tile's MMAs are issued: it lands while they run and the workers drain them
      if (tile + gridDim.x < total) { load_halo(tile + gridDim.x); UF_TRACE(1, it + 1); }
      if (it > 0) mbar_wait(&bars[4], (it - 1) & 1);                       // accumulator drained
      tc_fence_after();
      UF_TRACE(4, it);
      if (fused::elect_one()) {
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t a = da + ((2 * ks * kUfLboA) >> 4), bb = db + ((2 * ks * 128) >> 4);
          mma_f16_ss(tmem, a, bb, idesc, ks != 0);
          mma_f16_ss(tmem, a + (L.a_half >> 4), bb, idesc, 1);
          mma_f16_ss(tmem, a, bb + (UfSmem::B_HALF >> 4), idesc, 1);
        }
        mma_commit(&bars[3]);
      }
      __syncwarp();
      UF_TRACE(3, it);
    }
  } else {
    // =============================== workers ===============================
    const int nseg = 128 / (8 * Q) >= 2 ? 2 : 1;                            // Q = 8: two row segments of 8 rows, else one of 16
    const int cq = tid % (8 * Q), seg = tid / (8 * Q);
    const bool active = seg < nseg;
    const int col = cq / Q, quad = cq - col * Q;
    const int rows = kUfTH / nseg, r0 = seg * rows;
    float4 k[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) k[t] = active ? __ldg(reinterpret_cast<const float4*>(dw + t * CP + 4 * quad)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const int hrow = (kUfTW + 2) * Q;                                       // float4 per halo row
    const float4* const halo0 = reinterpret_cast<const float4*>(smem) + (r0 * hrow + col * Q + quad);   // window origin of this thread
    auto ldrow = [&](const float4* p, float4 (&v)[3]) {
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        v[x] = p[x * Q];
        if (RELU_IN) { v[x].x = fmaxf(v[x].x, 0.f); v[x].y = fmaxf(v[x].y, 0.f); v[x].z = fmaxf(v[x].z, 0.f); v[x].w = fmaxf(v[x].w, 0.f); }
      }
    };
    unsigned char* const a_dst0 = smem + L.off_a + (quad >> 1) * kUfLboA + (quad & 1) * 8 + (r0 * kUfTW + col) * 16;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    const bool wide = (ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0;   // pixel rows start on 32-byte boundaries
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const uint32_t b = tile / tiles_per, tr = tile - b * tiles_per;
      const uint32_t ty = tr / (uint32_t)tiles_w, tx = tr - ty * (uint32_t)tiles_w;
      const int h0 = (int)ty * kUfTH, w0 = (int)tx * kUfTW;
      mbar_wait(&bars[0], it & 1);
      // (the MMAs that read the A planes of the previous tile are done: this thread's epilogue below waited for them)
      if (warp == 0) UF_TRACE(10, it);
      if (active) {
        float4 w[3][3];
        ldrow(halo0, w[1]);
        ldrow(halo0 + hrow, w[2]);
        const float4* hp = halo0 + 2 * hrow;
        unsigned char* dst = a_dst0;
        for (int rr = 0; rr < rows; rr += 8) {
          // eight rows per trip, fully unrolled: the window rotates by register renaming (no moves except once per trip)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int x = 0; x < 3; ++x) { w[0][x] = w[1][x]; w[1][x] = w[2][x]; }
            ldrow(hp, w[2]);
            hp += hrow;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const float4 v = w[t / 3][t % 3];
              fused::fma2(acc.x, acc.y, v.x, v.y, k[t].x, k[t].y);   // FFMA2: two fp32 FMAs per issued instruction, each rounded like fmaf
              fused::fma2(acc.z, acc.w, v.z, v.w, k[t].z, k[t].w);
            }
            const __half2 h01 = __floats2half2_rn(acc.x, acc.y), h23 = __floats2half2_rn(acc.z, acc.w);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(acc.x - f01.x, acc.y - f01.y), l23 = __floats2half2_rn(acc.z - f23.x, acc.w - f23.y);
            uint2 hv, lv;
            hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
            lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
            *reinterpret_cast<uint2*>(dst) = hv;
            *reinterpret_cast<uint2*>(dst + L.a_half) = lv;
            dst += kUfTW * 16;
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) fused::mbar_arrive(&bars[2]);
      if (warp == 0) UF_TRACE(11, it);
      // ---- epilogue: accumulator row tid = pixel (tid / 8, tid % 8) ----
      mbar_wait(&bars[3], it & 1);
      tc_fence_after();
      if (warp == 0) UF_TRACE(12, it);
      const int hh = h0 + (tid >> 3), ww = w0 + (tid & 7);
      const bool inside = hh < H && ww < W;
      float* orow = out + (((size_t)b * H + hh) * W + ww) * ldc;
#pragma unroll 2
      for (int c0 = 0; c0 < 64; c0 += 16) {
        if (c0 >= n_valid) break;
        float v[16];
        tmem_ld16(lane_addr + c0, v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 bq = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * q);
          v[4 * q] += bq.x; v[4 * q + 1] += bq.y; v[4 * q + 2] += bq.z; v[4 * q + 3] += bq.w;
        }
        if (ACT == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (inside) {
          // one whole 32-byte sector per store instruction where the row allows it (a thread owns a pixel: its stores are 4 * ldc
          // bytes apart from its neighbours', so 16-byte stores would leave every sector half written per instruction)
#pragma unroll
          for (int q = 0; q < 4; q += 2) {
            if (wide && c0 + 4 * q + 8 <= n_valid) st_global_v8(orow + c0 + 4 * q, v + 4 * q);
            else {
#pragma unroll
              for (int qq = q; qq < q + 2; ++qq)
                if (c0 + 4 * qq + 4 <= n_valid) reinterpret_cast<float4*>(orow + c0)[qq] = make_float4(v[4 * qq], v[4 * qq + 1], v[4 * qq + 2], v[4 * qq + 3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) fused::mbar_arrive(&bars[4]);
      if (warp == 0) UF_TRACE(13, it);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

}  // namespace precise

// block 1 of the fp32-grade path: one CTA per SM (the (hi, lo) plane sets fill its shared memory), two issuer warps; its output
// is the un-rectified block output in fp32 (the next block applies the ReLU on load and walks it with stride 2 for its residual)
#ifndef ORCAI_B1_XBUF
#define ORCAI_B1_XBUF 2
#endif
using FB1P = fused::FB<16, 30, 29, 4, false, 1, 8, false, 2, ORCAI_B1_XBUF, true, true>;   // measured alternatives: 16 worker warps 3.40 vs 3.25 ms per 10 min (not issue-bound); strips of 14 columns with two CTAs per SM 11.2 vs 8.9 ms per 1 024 snippets (80 registers, more halo)

inline std::vector<float> pad_matrix(const float* w, int rows, int cols, int rows_p, int cols_p) {
  std::vector<float> out((size_t)rows_p * cols_p, 0.f);
  for (int r = 0; r < rows; ++r)
    for (int q = 0; q < cols; ++q) out[(size_t)r * cols_p + q] = w[(size_t)r * cols + q];
  return out;
}

int build_precise_sep(Ctx* c, const NetWeights::HostSep& hs, NetWeights::PreciseSep* out) {
  const int cip = cpad8(hs.ci);
  ORCAI_CHECK(net_upload(c, pad_matrix(hs.dw.data(), 9, hs.ci, 9, cip), &out->dw));
  ORCAI_CHECK(net_pack_split_b(c, hs.pw.data(), hs.ci, hs.co, 64, &out->pw));
  std::vector<float> b(64, 0.f);
  for (int n = 0; n < hs.co; ++n) b[n] = hs.b[n];
  ORCAI_CHECK(net_upload(c, b, &out->bias));
  return ORCAI_OK;
}

int prepare_precise(Ctx* c) {
  NetWeights* nw = c->net;
  if (nw->precise_ready) return ORCAI_OK;
  ORCAI_CHECK(prepare_bringup_aids(c));
  ORCAI_CHECK(build_fused_block<FB1P>(c, 0));
  for (int b = 1; b < nw->n_blocks; ++b) {
    ORCAI_CHECK(build_precise_sep(c, nw->h_sep1[b], &nw->p_sep1[b]));
    ORCAI_CHECK(build_precise_sep(c, nw->h_sep2[b], &nw->p_sep2[b]));
    const int ci = nw->h_sep1[b].ci, co = nw->h_sep1[b].co;
    ORCAI_CHECK(net_upload(c, pad_matrix(nw->h_res_w[b].data(), ci, co, cpad8(ci), cpad8(co)), &nw->p_res_w[b]));
    ORCAI_CHECK(net_upload(c, pad_matrix(nw->h_res_b[b].data(), 1, co, 1, cpad8(co)), &nw->p_res_b[b]));
  }
  ORCAI_CHECK(build_precise_sep(c, nw->h_fin, &nw->p_fin));
  nw->precise_ready = true;
  return ORCAI_OK;
}

// rank-4 map over an fp32 NHWC tensor (n, h, w, cp); box = {cp, 10, 18, 1}: the halo of a 16 x 8 pixel tile
int make_uf_map(Ctx* c, CUtensorMap* map, const float* base, long long n, long long h, int w, int cp) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) ORCAI_FAIL(c, ORCAI_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[4] = {(cuuint64_t)cp, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  const cuuint64_t strides[3] = {(cuuint64_t)cp * 4, (cuuint64_t)w * cp * 4, (cuuint64_t)h * w * cp * 4};
  const cuuint32_t box[4] = {(cuuint32_t)cp, (cuuint32_t)(precise::kUfTW + 2), (cuuint32_t)(precise::kUfTH + 2), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) ORCAI_FAIL(c, ORCAI_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for tensor (%lld, %lld, %d, %d)", (int)r, n, h, w, cp);
  return ORCAI_OK;
}

template <bool RELU_IN, int ACT>
int launch_sep_uf(Ctx* c, const float* x, float* out, long long n, long long h, int w, int cip, int ldc, int n_valid, const NetWeights::PreciseSep& ps) {
  static std::atomic<unsigned long long> attr_devices{0ull};
  if (!((attr_devices.load() >> (c->device & 63)) & 1ull)) {
    ORCAI_CUDA(c, cudaFuncSetAttribute(precise::sep_uf_kernel<RELU_IN, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)precise::UfSmem(64).bytes));
    attr_devices.fetch_or(1ull << (c->device & 63));
  }
  const precise::UfSmem L(cip);
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(3, (size_t)(225 * 1024) / (L.bytes + 1024)));
  CUtensorMap tm;
  ORCAI_CHECK(make_uf_map(c, &tm, x, n, h, w, cip));
  const int tiles_w = (w + precise::kUfTW - 1) / precise::kUfTW, tiles_h = (int)((h + precise::kUfTH - 1) / precise::kUfTH);
  const long long total = n * tiles_w * tiles_h;
  if (total >= (1ll << 31)) ORCAI_FAIL(c, ORCAI_ERR_ARG, "sep_uf_kernel: %lld tiles exceed the 32-bit tile index", total);
  const unsigned grid = (unsigned)std::min<long long>(total, (long long)c->sm_count * per_sm);
  precise::sep_uf_kernel<RELU_IN, ACT><<<grid, 160, L.bytes, c->stream>>>(tm, ps.dw, ps.pw, ps.bias, out, n, (int)h, w, cip, ldc, n_valid, tiles_w,
                                                                                        tiles_h);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

// one un-folded separable convolution: out = act(dw3x3(f(x)) * pw + bias).  x (n, h, w, cip) fp32 -> out (n, h, w, ldc) fp32.
// sep_path 1 (default): one kernel (sep_uf_kernel); 0: depthwise kernel -> d, then the split GEMM (bit-identical results: the same
// fp32 FMA order per pixel and the same three MMAs per K step).
int run_precise_sep(Ctx* c, const float* x, float* d, float* out, long long n, long long h, int w, int cip, int ldc, int n_valid, bool relu_in,
                    bool relu_out, const NetWeights::PreciseSep& ps) {
  const long long total = n * h * w * (cip / 4);
  if (total <= 0) return ORCAI_OK;
  if (c->net->precise_sep_path == 1) {
    if (relu_in && relu_out) return launch_sep_uf<true, 1>(c, x, out, n, h, w, cip, ldc, n_valid, ps);
    if (relu_in) return launch_sep_uf<true, 0>(c, x, out, n, h, w, cip, ldc, n_valid, ps);
    if (relu_out) return launch_sep_uf<false, 1>(c, x, out, n, h, w, cip, ldc, n_valid, ps);
    return launch_sep_uf<false, 0>(c, x, out, n, h, w, cip, ldc, n_valid, ps);
  }
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)c->sm_count * 32);
  if (relu_in) precise::dw3x3_kernel<true><<<grid, 256, 0, c->stream>>>(x, d, ps.dw, n, (int)h, w, cip);
  else precise::dw3x3_kernel<false><<<grid, 256, 0, c->stream>>>(x, d, ps.dw, n, (int)h, w, cip);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return net_gemm_split(c, d, cip, ps.pw, ps.bias, out, ldc, n * h * w, 64, cip, n_valid, relu_out ? 1 : 0);
}

// residual block b (1-based index blk >= 1 into the weights) on fp32 images x (n, h, w, cip), un-rectified -> y (n, h/2, ceil(w/2), cop)
int run_precise_block(Ctx* c, int blk, const float* x, float* ta, float* tb, float* tc, float* y, long long n, long long h, int w, int cip, int cop) {
  NetWeights* nw = c->net;
  ORCAI_CHECK(run_precise_sep(c, x, ta, tb, n, h, w, cip, cop, cop, true, true, nw->p_sep1[blk]));
  ORCAI_CHECK(run_precise_sep(c, tb, ta, tc, n, h, w, cop, cop, cop, false, false, nw->p_sep2[blk]));
  const long long ho = h / 2;
  const int wo = (w + 1) / 2;
  const long long total = n * ho * ((wo + precise::kPoolPx - 1) / precise::kPoolPx) * (cop / 4);
  if (total <= 0) return ORCAI_OK;
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)c->sm_count * 16);
  const size_t smem = ((size_t)cip * cop + cop) * sizeof(float);
  // the residual 1x1/2 convolution walks the block input with stride 2
  precise::pool_res_f32_kernel<<<grid, 256, smem, c->stream>>>(tc, x, h * w * cip, (long long)2 * w * cip, 2 * cip, y, n, (int)h, w, (int)ho, wo, cip, cop,
                                                              nw->p_res_w[blk], nw->p_res_b[blk]);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

int run_conv0_split(Ctx* c, const float* src, int input_mode, long long first, int in_ld, long long n_img, int Himg, int Wf, __half* hi, __half* lo,
                    long long n_snip, int off_bot, int Hfull, long long plane_halfs) {
  const int rpt = Himg >= 32 ? 4 : 1;                                  // tall images / whole snippets: four rows per thread; 8-row border images: one
  const int tiles_w = (Wf + kC0TW - 1) / kC0TW, tiles_h = (Himg + c0_tile_rows(rpt) - 1) / c0_tile_rows(rpt);
  const long long blocks = n_img * tiles_w * tiles_h;
  if (blocks <= 0) return ORCAI_OK;
  if (rpt == 4)
    conv0_direct_kernel<true, 4><<<(unsigned)blocks, 256, 0, c->stream>>>(src, input_mode, first, c->p.snippet_len / 2, in_ld, Himg, Wf, c->d_sel, hi,
                                                                          static_cast<__half*>(nullptr), tiles_w, tiles_h, lo, n_snip, off_bot, Hfull, plane_halfs);
  else
    conv0_direct_kernel<true, 1><<<(unsigned)blocks, 256, 0, c->stream>>>(src, input_mode, first, c->p.snippet_len / 2, in_ld, Himg, Wf, c->d_sel, hi,
                                                                          static_cast<__half*>(nullptr), tiles_w, tiles_h, lo, n_snip, off_bot, Hfull, plane_halfs);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

// block 1 over ONE tall image (rows, W) of (hi, lo) fp16 pairs, cut into windows of `stride` rows (+ warm-up); guard rows of
// zeros lie before x (>= warm rows) and after it (>= stride + 16 rows)
int run_block1_tall(Ctx* c, const __half* x_hi, const __half* x_lo, long long plane_halfs, float* y, long long rows, int Wimg, int stride, int warm) {
  using G = FB1P;
  NetWeights* nw = c->net;
  const int Wo = (Wimg + 1) / 2;
  const int n_strips = (Wo + G::CP - 1) / G::CP;
  const long long n_win = (rows + stride - 1) / stride;
  const long long items = n_win * n_strips;
  const long long grid = std::min<long long>(items, (long long)c->sm_count * G::CTAS);
  const int Hloc = warm + stride;
  const size_t shift = (size_t)warm * Wimg * 8;      // halfs per plane
  CUtensorMap tmx, tmr, tmxl, tmrl;
  ORCAI_CHECK(make_planar_map(c, &tmx, x_hi - shift, n_win, Hloc + 8, Wimg, G::XG, plane_halfs, G::WP, G::S + 2, 1, stride));
  ORCAI_CHECK(make_planar_map(c, &tmr, x_hi - shift, n_win, Hloc + 8, Wimg, G::XG, plane_halfs, G::CP, G::S / 2, 2, stride));
  ORCAI_CHECK(make_planar_map(c, &tmxl, x_lo - shift, n_win, Hloc + 8, Wimg, G::XG, plane_halfs, G::WP, G::S + 2, 1, stride));
  ORCAI_CHECK(make_planar_map(c, &tmrl, x_lo - shift, n_win, Hloc + 8, Wimg, G::XG, plane_halfs, G::CP, G::S / 2, 2, stride));
  fused::TallView tv;
  tv.on = 1; tv.stride = stride; tv.warm = warm; tv.rows = rows;
  fused::fused_block_kernel<G><<<(unsigned)grid, G::NTHREADS, G::SMEM, c->stream>>>(tmx, tmr, 2, reinterpret_cast<__half*>(y), static_cast<__half*>(nullptr), Hloc,
                                                                                   Wimg, n_strips, items, static_cast<const unsigned char*>(nw->fbp_w[0]), tmxl,
                                                                                   tmrl, tv);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

struct PreciseGeom {
  int hs[5], ws[5], cp[5];
};

// ---- independent snippets: host-provided batches (input_mode 1) and the stage-debug reads ------------------------------------
int forward_precise_snippets(Ctx* c, const float* d_in, int input_mode, int64_t first, int64_t n, float* d_preds, const PreciseGeom& g) {
  NetWeights* nw = c->net;
  using H16 = __half;
  const int Himg = nw->H, Wf = nw->Wf, U = nw->U, L = nw->L;
  const int Tn = Himg >> nw->n_blocks;
  const int *hs = g.hs, *ws = g.ws, *cp = g.cp;
  const size_t c0_h = (size_t)hs[0] * ws[0] * 16;                       // halfs per plane set
  const size_t y1 = (size_t)hs[1] * ws[1] * cp[1], y2 = (size_t)hs[2] * ws[2] * cp[2], y3 = (size_t)hs[3] * ws[3] * cp[3], y4 = (size_t)hs[4] * ws[4] * cp[4];
  const size_t tmp = (size_t)hs[1] * ws[1] * cp[2];
  const size_t feat_f = (size_t)Tn * nw->feat;
  const size_t tail_f = (size_t)Tn * (2 * 4 * U + 2 * U + 2 * U + 128);
  const size_t per = c0_h * 2 * 2 + (y1 + y2 + y3 + y4 + 3 * tmp + feat_f + tail_f) * 4;
  const long long chunk = std::min<long long>(std::max(nw->chunk_precise, 1), n);
  if (chunk <= 0) return ORCAI_OK;
  ORCAI_CHECK(ensure_device_buffer(c, &nw->tc_ws, &nw->tc_ws_cap, per * (size_t)chunk + 256));
  H16* c0hi = static_cast<H16*>(nw->tc_ws);
  H16* c0lo = c0hi + c0_h * chunk;
  float* Y1 = reinterpret_cast<float*>(c0lo + c0_h * chunk);
  float* Y2 = Y1 + y1 * chunk;
  float* Y3 = Y2 + y2 * chunk;
  float* Y4 = Y3 + y3 * chunk;
  float* TA = Y4 + y4 * chunk;
  float* TB = TA + tmp * chunk;
  float* TC = TB + tmp * chunk;
  float* feat = TC + tmp * chunk;
  float* scratch = feat + feat_f * chunk;
  const int stop = nw->debug_stop;
  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    const long long m = std::min<long long>(chunk, n - s0);
    const bool mk = (s0 == 0);
    if (mk) nw->marked_snippets = m;
    net_mark(c, mk);
    const float* src = (input_mode == 0) ? d_in : d_in + (size_t)s0 * Himg * Wf;
    const long long plane = (long long)m * hs[0] * ws[0] * 8;   // chunk-planar (hi, lo) output: two planes of 8 channels each
    ORCAI_CHECK(run_conv0_split(c, src, input_mode, first + s0, input_mode == 0 ? kRawLd : Wf, m, Himg, Wf, c0hi, c0lo, m, 0, Himg, plane));
    net_mark(c, mk);  // 0: conv0
    if (stop == 0) { set_debug(nw, c0hi, 1, m, hs[0], ws[0], 8, 8); return ORCAI_OK; }   // channels 0-7, hi plane
    ORCAI_CHECK((run_fused_block<FB1P>(c, 0, c0hi, static_cast<const H16*>(nullptr), reinterpret_cast<H16*>(Y1), static_cast<H16*>(nullptr), m, hs[0], ws[0],
                                       c0lo, static_cast<const H16*>(nullptr), plane)));
    net_mark(c, mk);  // 1
    if (stop == 1) { set_debug(nw, Y1, 0, m, hs[1], ws[1], 30, cp[1]); return ORCAI_OK; }
    ORCAI_CHECK(run_precise_block(c, 1, Y1, TA, TB, TC, Y2, m, hs[1], ws[1], cp[1], cp[2]));
    net_mark(c, mk);  // 2
    if (stop == 2) { set_debug(nw, Y2, 0, m, hs[2], ws[2], 40, cp[2]); return ORCAI_OK; }
    ORCAI_CHECK(run_precise_block(c, 2, Y2, TA, TB, TC, Y3, m, hs[2], ws[2], cp[2], cp[3]));
    net_mark(c, mk);  // 3
    if (stop == 3) { set_debug(nw, Y3, 0, m, hs[3], ws[3], 50, cp[3]); return ORCAI_OK; }
    ORCAI_CHECK(run_precise_block(c, 3, Y3, TA, TB, TC, Y4, m, hs[3], ws[3], cp[3], cp[4]));
    net_mark(c, mk);  // 4
    if (stop == 4) { set_debug(nw, Y4, 0, m, hs[4], ws[4], 60, cp[4]); return ORCAI_OK; }
    // final separable convolution -> features (m, Tn, w*36 + c)
    ORCAI_CHECK(run_precise_sep(c, Y4, TA, feat, m, hs[4], ws[4], cp[4], 36, 36, false, true, nw->p_fin));
    net_mark(c, mk);  // 5
    if (stop == 5) { set_debug(nw, feat, 0, m, hs[4], ws[4], 36, 36); return ORCAI_OK; }
    ORCAI_CHECK(net_tail_precise(c, feat, scratch, m, d_preds + (size_t)s0 * Tn * L, mk));
  }
  return ORCAI_OK;
}

// ---- snippets of the resident recording (input_mode 0): shared interior ---------------------------------------------------------
// Consecutive snippets overlap by half their length, and away from a snippet's own zero-padded top and bottom the convolutional
// trunk computes the same numbers for both (predict.py:252-261 cuts the windows; the 16-row alignment of the shift keeps the
// pooling grids aligned).  So the trunk runs ONCE over the chunk's rows as one tall image (half the rows of the snippet batch),
// and only the rows that feel a snippet's own border are recomputed per snippet from 8-row border images: at every level the
// first nt = 2 and last nb = 3 output rows of a snippet (1 / 1 after the entry convolution, 2 / 2 after block 1, 3 / 4 of the
// features).  Every kernel computes a pixel from its own receptive field only, with one fixed instruction sequence, so the
// result is BIT-IDENTICAL to the per-snippet evaluation (tests/test_gpu_network.py holds the two to assert_array_equal).
int forward_precise_tall(Ctx* c, const float* d_raw, int64_t first, int64_t n, float* d_preds, const PreciseGeom& g) {
  NetWeights* nw = c->net;
  using H16 = __half;
  const int Himg = nw->H, Wf = nw->Wf, U = nw->U, L = nw->L;
  const int Tn = Himg >> nw->n_blocks;
  const int *hs = g.hs, *ws = g.ws, *cp = g.cp;
  const int shift = c->p.snippet_len / 2;
  constexpr int kWarm = 2 * FB1P::S, kGuardB = 8;            // block 1: warm-up rows of a window (one step), zero rows before the tall tensor
  const int kWin = shift;                          // block 1: rows per window
  const int kGuardA = kWin + 16;                   // zero rows after the tall tensor
  const int nt[6] = {1, 2, 2, 2, 2, 3}, nb[6] = {1, 2, 3, 3, 3, 4};   // rows of a snippet that feel its top / bottom border, per level (5 = features)
  const long long chunk = std::min<long long>(std::max(nw->chunk_precise, 1), n);
  if (chunk <= 0) return ORCAI_OK;
  // workspace (per chunk of m snippets: tall rows R_l = (m + 1) * (shift >> l) at level l)
  auto rows_at = [&](long long m, int l) { return (m + 1) * (long long)(shift >> l); };
  const long long M = chunk;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t c0_plane = (size_t)(kGuardB + rows_at(M, 0) + kGuardA) * ws[0] * 16 * 2;
  const size_t o_c0hi = take(c0_plane), o_c0lo = take(c0_plane);
  size_t o_y[5] = {};
  for (int l = 1; l <= 4; ++l) o_y[l] = take((size_t)rows_at(M, l) * ws[l] * cp[l] * 4);
  const size_t tmp = (size_t)rows_at(M, 1) * ws[1] * cp[2] * 4;
  const size_t o_ta = take(tmp), o_tb = take(tmp), o_tc = take(tmp);
  const size_t o_featt = take((size_t)rows_at(M, 4) * nw->feat * 4);
  const size_t b0_plane = (size_t)2 * M * 8 * ws[0] * 16 * 2;
  const size_t o_b0hi = take(b0_plane), o_b0lo = take(b0_plane);
  const size_t alt_sz = (size_t)2 * M * 4 * ws[1] * cp[1] * 4, bimg_sz = (size_t)2 * M * 8 * ws[1] * cp[1] * 4, btmp = (size_t)2 * M * 8 * ws[1] * cp[2] * 4;
  const size_t o_alt = take(alt_sz), o_bimg = take(bimg_sz), o_bta = take(btmp), o_btb = take(btmp), o_btc = take(btmp);
  const size_t o_featb = take((size_t)2 * M * 8 * nw->feat * 4);
  const size_t o_feat = take((size_t)M * Tn * nw->feat * 4);
  const size_t o_tail = take((size_t)M * Tn * (2 * 4 * U + 2 * U + 2 * U + 128) * 4);
  ORCAI_CHECK(ensure_device_buffer(c, &nw->tc_ws, &nw->tc_ws_cap, off + 256));
  unsigned char* base = static_cast<unsigned char*>(nw->tc_ws);
  auto F = [&](size_t o) { return reinterpret_cast<float*>(base + o); };
  // entry-convolution output, chunk-planar: per (hi | lo) set two planes of (guard + rows + guard, W, 8) halfs
  const size_t c0_row = (size_t)ws[0] * 8;         // halfs per row of one plane
  const long long c0_plane_halfs = (long long)(kGuardB + rows_at(M, 0) + kGuardA) * c0_row;
  H16* c0hi = reinterpret_cast<H16*>(base + o_c0hi) + kGuardB * c0_row;
  H16* c0lo = reinterpret_cast<H16*>(base + o_c0lo) + kGuardB * c0_row;
  float* Y[5] = {nullptr, F(o_y[1]), F(o_y[2]), F(o_y[3]), F(o_y[4])};
  float *TA = F(o_ta), *TB = F(o_tb), *TC = F(o_tc), *FEATT = F(o_featt);
  H16* b0hi = reinterpret_cast<H16*>(base + o_b0hi);
  H16* b0lo = reinterpret_cast<H16*>(base + o_b0lo);
  float *ALT = F(o_alt), *BIMG = F(o_bimg), *BTA = F(o_bta), *BTB = F(o_btb), *BTC = F(o_btc), *FEATB = F(o_featb);
  float *feat = F(o_feat), *scratch = F(o_tail);

  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    const long long m = std::min<long long>(chunk, n - s0);
    const bool mk = (s0 == 0);
    if (mk) nw->marked_snippets = m;
    const long long R0 = rows_at(m, 0);
    net_mark(c, mk);
    // entry convolution: the chunk's rows as one image, and the 8-row border images of every snippet (its own zero padding)
    for (H16* p0 : {c0hi, c0lo})
      for (int pl = 0; pl < 2; ++pl) {
        H16* p = p0 + (size_t)pl * c0_plane_halfs;
        ORCAI_CUDA(c, cudaMemsetAsync(p - kGuardB * c0_row, 0, kGuardB * c0_row * 2, c->stream));
        ORCAI_CUDA(c, cudaMemsetAsync(p + R0 * c0_row, 0, (size_t)kGuardA * c0_row * 2, c->stream));
      }
    const long long b0_plane_halfs = (long long)2 * m * 8 * ws[0] * 8;
    ORCAI_CHECK(run_conv0_split(c, d_raw, 0, first + s0, kRawLd, 1, (int)R0, Wf, c0hi, c0lo, 1, 0, (int)R0, c0_plane_halfs));
    ORCAI_CHECK(run_conv0_split(c, d_raw, 0, first + s0, kRawLd, 2 * m, 8, Wf, b0hi, b0lo, m, Himg - 8, Himg, b0_plane_halfs));
    net_mark(c, mk);  // 0: conv0
    // block 1: fused kernel over windows of the tall image; border images as independent 8-row images -> ALT (2m, 4, ..)
    ORCAI_CHECK(run_block1_tall(c, c0hi, c0lo, c0_plane_halfs, Y[1], R0, Wf, kWin, kWarm));
    ORCAI_CHECK((run_fused_block<FB1P>(c, 0, b0hi, static_cast<const H16*>(nullptr), reinterpret_cast<H16*>(ALT), static_cast<H16*>(nullptr), 2 * m, 8, ws[0],
                                       b0lo, static_cast<const H16*>(nullptr), b0_plane_halfs)));
    net_mark(c, mk);  // 1
    // blocks 2 - 4: tall image, then the border images gathered from the tall tensor and the previous level's border rows
    for (int l = 1; l <= 3; ++l) {
      ORCAI_CHECK(run_precise_block(c, l, Y[l], TA, TB, TC, Y[l + 1], 1, rows_at(m, l), ws[l], cp[l], cp[l + 1]));
      const int row_f4 = ws[l] * cp[l] / 4;
      const long long total = 2 * m * 8 * row_f4;
      precise::gather_border_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, (long long)c->sm_count * 16), 256, 0, c->stream>>>(
          Y[l], ALT, BIMG, m, shift >> l, hs[l], row_f4, nt[l], nb[l]);
      c->launches++;
      ORCAI_CUDA(c, cudaGetLastError());
      ORCAI_CHECK(run_precise_block(c, l, BIMG, BTA, BTB, BTC, ALT, 2 * m, 8, ws[l], cp[l], cp[l + 1]));
      net_mark(c, mk);  // 2, 3, 4
    }
    {  // final separable convolution (no pooling: 8-row border images give 8 feature rows) and the per-snippet feature rows
      ORCAI_CHECK(run_precise_sep(c, Y[4], TA, FEATT, 1, rows_at(m, 4), ws[4], cp[4], 36, 36, false, true, nw->p_fin));
      const int row_f4 = ws[4] * cp[4] / 4;
      const long long total = 2 * m * 8 * row_f4;
      precise::gather_border_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, (long long)c->sm_count * 16), 256, 0, c->stream>>>(
          Y[4], ALT, BIMG, m, shift >> 4, hs[4], row_f4, nt[4], nb[4]);
      c->launches++;
      ORCAI_CHECK(run_precise_sep(c, BIMG, BTA, FEATB, 2 * m, 8, ws[4], cp[4], 36, 36, false, true, nw->p_fin));
      const int feat_f4 = nw->feat / 4;
      const long long tot2 = m * Tn * feat_f4;
      precise::assemble_feat_kernel<<<(unsigned)std::min<long long>((tot2 + 255) / 256, (long long)c->sm_count * 16), 256, 0, c->stream>>>(
          FEATT, FEATB, feat, m, shift >> 4, Tn, feat_f4, nt[5], nb[5]);
      c->launches++;
      ORCAI_CUDA(c, cudaGetLastError());
    }
    net_mark(c, mk);  // 5
    ORCAI_CHECK(net_tail_precise(c, feat, scratch, m, d_preds + (size_t)s0 * Tn * L, mk));
  }
  return ORCAI_OK;
}

int forward_precise(Ctx* c, const float* d_in, int input_mode, int64_t first, int64_t n, float* d_preds) {
  NetWeights* nw = c->net;
  PreciseGeom g;
  g.hs[0] = nw->H; g.ws[0] = nw->Wf;
  for (int b = 0; b < 4; ++b) { g.hs[b + 1] = g.hs[b] / 2; g.ws[b + 1] = (g.ws[b] + 1) / 2; }
  g.cp[0] = 16;
  for (int b = 0; b < 4; ++b) g.cp[b + 1] = cpad8(nw->filters[b]);
  nw->mark_i = 0;
  nw->dbg_ptr = nullptr;
  {  // constant memory is per device, not per context: refresh it stream-ordered before every forward
    nw->h_conv0_pack.resize(160);
    memcpy(nw->h_conv0_pack.data(), nw->h_conv0_w.data(), 144 * sizeof(float));
    memcpy(nw->h_conv0_pack.data() + 144, nw->h_conv0_b.data(), 16 * sizeof(float));
    ORCAI_CUDA(c, cudaMemcpyToSymbolAsync(c_conv0, nw->h_conv0_pack.data(), 160 * sizeof(float), 0, cudaMemcpyHostToDevice, c->stream));
  }
  // the shared-interior evaluation needs the geometry it was derived for: 4 blocks, snippets half a length apart, 16-row aligned
  // (and at least three snippets: the tall image of m snippets has (m + 1) / 2m of their rows plus the border images)
  const bool tall_ok = input_mode == 0 && nw->debug_stop < 0 && nw->precise_tall && nw->n_blocks == 4 && (c->p.snippet_len / 2) % 16 == 0 &&
                       nw->H >= 128 && nw->H % 16 == 0 && n >= 3;
  if (tall_ok) return forward_precise_tall(c, d_in, first, n, d_preds, g);
  return forward_precise_snippets(c, d_in, input_mode, first, n, d_preds, g);
}
