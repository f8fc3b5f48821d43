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
// tensor, or the full tensor walked with stride 2.  rw: [CIP][COP] (zero padded), rb: [COP].  One thread per (pixel, 4 channels).
__global__ void __launch_bounds__(256)
pool_res_f32_kernel(const float* __restrict__ s2, const float* __restrict__ xs, long long xs_img, long long xs_row, int xs_px,
                    float* __restrict__ y, long long n, int H, int W, int Ho, int Wo, int CIP, int COP,
                    const float* __restrict__ rw, const float* __restrict__ rb) {
  extern __shared__ float s_w[];   // [CIP][COP] then [COP]
  for (int i = threadIdx.x; i < CIP * COP; i += blockDim.x) s_w[i] = rw[i];
  for (int i = threadIdx.x; i < COP; i += blockDim.x) s_w[CIP * COP + i] = rb[i];
  __syncthreads();
  const int G4 = COP >> 2;
  const long long total = n * Ho * Wo * G4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G4);
    long long r = idx / G4;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const long long b = r / Ho;
    const float* sb = s2 + (size_t)b * H * W * COP + 4 * g;
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int hh = 2 * ho + dy;
      if (hh >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int ww = 2 * wo + dx;
        if (ww >= W) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(sb + ((size_t)hh * W + ww) * COP));
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    const float* xp = xs + (size_t)b * xs_img + (size_t)ho * xs_row + (size_t)wo * xs_px;
    float4 a = *reinterpret_cast<const float4*>(s_w + CIP * COP + 4 * g);
    for (int ci = 0; ci < CIP; ci += 4) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(xp + ci));
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + (ci + 0) * COP + 4 * g);
      const float4 w1 = *reinterpret_cast<const float4*>(s_w + (ci + 1) * COP + 4 * g);
      const float4 w2 = *reinterpret_cast<const float4*>(s_w + (ci + 2) * COP + 4 * g);
      const float4 w3 = *reinterpret_cast<const float4*>(s_w + (ci + 3) * COP + 4 * g);
      a.x = fmaf(xv.x, w0.x, a.x); a.y = fmaf(xv.x, w0.y, a.y); a.z = fmaf(xv.x, w0.z, a.z); a.w = fmaf(xv.x, w0.w, a.w);
      a.x = fmaf(xv.y, w1.x, a.x); a.y = fmaf(xv.y, w1.y, a.y); a.z = fmaf(xv.y, w1.z, a.z); a.w = fmaf(xv.y, w1.w, a.w);
      a.x = fmaf(xv.z, w2.x, a.x); a.y = fmaf(xv.z, w2.y, a.y); a.z = fmaf(xv.z, w2.z, a.z); a.w = fmaf(xv.z, w2.w, a.w);
      a.x = fmaf(xv.w, w3.x, a.x); a.y = fmaf(xv.w, w3.y, a.y); a.z = fmaf(xv.w, w3.z, a.z); a.w = fmaf(xv.w, w3.w, a.w);
    }
    *reinterpret_cast<float4*>(y + (size_t)idx * 4) = make_float4(m.x + a.x, m.y + a.y, m.z + a.z, m.w + a.w);
  }
}

}  // namespace precise

// block 1 of the fp32-grade path: one CTA per SM (the (hi, lo) plane sets fill its shared memory), two issuer warps
using FB1P = fused::FB<16, 30, 29, 6, true, 1, 8, false, 2, 1, true>;

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

// one un-folded separable convolution: d = dw3x3(f(x)) ; out = act(d * pw + bias).  x (n, h, w, cip) fp32 -> out (n, h, w, ldc) fp32
int run_precise_sep(Ctx* c, const float* x, float* d, float* out, long long n, int h, int w, int cip, int ldc, int n_valid, bool relu_in,
                    bool relu_out, const NetWeights::PreciseSep& ps) {
  const long long total = n * h * w * (cip / 4);
  if (total <= 0) return ORCAI_OK;
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)c->sm_count * 32);
  if (relu_in) precise::dw3x3_kernel<true><<<grid, 256, 0, c->stream>>>(x, d, ps.dw, n, h, w, cip);
  else precise::dw3x3_kernel<false><<<grid, 256, 0, c->stream>>>(x, d, ps.dw, n, h, w, cip);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return net_gemm_split(c, d, cip, ps.pw, ps.bias, out, ldc, n * h * w, 64, cip, n_valid, relu_out ? 1 : 0);
}

int run_pool_res(Ctx* c, const float* s2, const float* xs, long long xs_img, long long xs_row, int xs_px, float* y, long long n, int h, int w,
                 int cip, int cop, int blk) {
  NetWeights* nw = c->net;
  const int ho = h / 2, wo = (w + 1) / 2;
  const long long total = n * ho * wo * (cop / 4);
  if (total <= 0) return ORCAI_OK;
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)c->sm_count * 16);
  const size_t smem = ((size_t)cip * cop + cop) * sizeof(float);
  precise::pool_res_f32_kernel<<<grid, 256, smem, c->stream>>>(s2, xs, xs_img, xs_row, xs_px, y, n, h, w, ho, wo, cip, cop, nw->p_res_w[blk], nw->p_res_b[blk]);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

int forward_precise(Ctx* c, const float* d_in, int input_mode, int64_t first, int64_t n, float* d_preds) {
  NetWeights* nw = c->net;
  using H16 = __half;
  const int Himg = nw->H, Wf = nw->Wf, U = nw->U, L = nw->L;
  const int Tn = Himg >> nw->n_blocks;
  int hs[5], ws[5];
  hs[0] = Himg; ws[0] = Wf;
  for (int b = 0; b < 4; ++b) { hs[b + 1] = hs[b] / 2; ws[b + 1] = (ws[b] + 1) / 2; }
  const int cp[5] = {16, cpad8(nw->filters[0]), cpad8(nw->filters[1]), cpad8(nw->filters[2]), cpad8(nw->filters[3])};
  // per snippet: entry-convolution output as (hi, lo) fp16; fp32 block outputs y1 (ReLU'd) + y1 at even positions, y2, y3, y4;
  // three fp32 temporaries sized for block 2's widest tensor; the tail
  const size_t c0_h = (size_t)hs[0] * ws[0] * 16;                       // halfs per plane set
  const size_t y1 = (size_t)hs[1] * ws[1] * cp[1], y1s = (size_t)hs[2] * ws[2] * cp[1];
  const size_t y2 = (size_t)hs[2] * ws[2] * cp[2], y3 = (size_t)hs[3] * ws[3] * cp[3], y4 = (size_t)hs[4] * ws[4] * cp[4];
  const size_t tmp = (size_t)hs[1] * ws[1] * cp[2];
  const size_t feat_f = (size_t)Tn * nw->feat;
  const size_t tail_f = (size_t)Tn * (2 * 4 * U + 2 * U + 2 * U + 128);
  const size_t per = c0_h * 2 * 2 + (y1 + y1s + y2 + y3 + y4 + 3 * tmp + feat_f + tail_f) * 4;
  const long long chunk = std::min<long long>(std::max(nw->chunk_precise, 1), n);
  if (chunk <= 0) return ORCAI_OK;
  ORCAI_CHECK(ensure_device_buffer(c, &nw->tc_ws, &nw->tc_ws_cap, per * (size_t)chunk + 256));
  H16* c0hi = static_cast<H16*>(nw->tc_ws);
  H16* c0lo = c0hi + c0_h * chunk;
  float* Y1 = reinterpret_cast<float*>(c0lo + c0_h * chunk);
  float* Y1s = Y1 + y1 * chunk;
  float* Y2 = Y1s + y1s * chunk;
  float* Y3 = Y2 + y2 * chunk;
  float* Y4 = Y3 + y3 * chunk;
  float* TA = Y4 + y4 * chunk;
  float* TB = TA + tmp * chunk;
  float* TC = TB + tmp * chunk;
  float* feat = TC + tmp * chunk;
  float* scratch = feat + feat_f * chunk;
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
    {
      const int tiles_w = (Wf + kC0TW - 1) / kC0TW, tiles_h = (Himg + kC0TH - 1) / kC0TH;
      const float* src = (input_mode == 0) ? d_in : d_in + (size_t)s0 * Himg * Wf;
      conv0_direct_kernel<true><<<(unsigned)(m * tiles_w * tiles_h), 256, 0, c->stream>>>(src, input_mode, first + s0, shift, input_mode == 0 ? kRawLd : Wf,
                                                                                         Himg, Wf, c->d_sel, c0hi, static_cast<H16*>(nullptr), tiles_w, tiles_h, c0lo);
      c->launches++;
      ORCAI_CUDA(c, cudaGetLastError());
    }
    net_mark(c, mk);  // 0: conv0
    if (stop == 0) { set_debug(nw, c0hi, 1, m, hs[0], ws[0], 16, 16); return ORCAI_OK; }
    ORCAI_CHECK((run_fused_block<FB1P>(c, 0, c0hi, static_cast<const H16*>(nullptr), reinterpret_cast<H16*>(Y1), reinterpret_cast<H16*>(Y1s), m, hs[0], ws[0],
                                       c0lo, static_cast<const H16*>(nullptr))));
    net_mark(c, mk);  // 1
    if (stop == 1) { set_debug(nw, Y1, 0, m, hs[1], ws[1], 30, cp[1]); return ORCAI_OK; }
    if (stop == 21) { set_debug(nw, Y1s, 0, m, hs[2], ws[2], 30, cp[1]); return ORCAI_OK; }
    // block 2: its input arrives ReLU'd (Y1) plus un-rectified at even positions (Y1s)
    ORCAI_CHECK(run_precise_sep(c, Y1, TA, TB, m, hs[1], ws[1], cp[1], cp[2], cp[2], false, true, nw->p_sep1[1]));
    ORCAI_CHECK(run_precise_sep(c, TB, TA, TC, m, hs[1], ws[1], cp[2], cp[2], cp[2], false, false, nw->p_sep2[1]));
    ORCAI_CHECK(run_pool_res(c, TC, Y1s, (long long)y1s, (long long)ws[2] * cp[1], cp[1], Y2, m, hs[1], ws[1], cp[1], cp[2], 1));
    net_mark(c, mk);  // 2
    if (stop == 2) { set_debug(nw, Y2, 0, m, hs[2], ws[2], 40, cp[2]); return ORCAI_OK; }
    // blocks 3, 4: input un-rectified (ReLU on load; the residual convolution walks it with stride 2)
    ORCAI_CHECK(run_precise_sep(c, Y2, TA, TB, m, hs[2], ws[2], cp[2], cp[3], cp[3], true, true, nw->p_sep1[2]));
    ORCAI_CHECK(run_precise_sep(c, TB, TA, TC, m, hs[2], ws[2], cp[3], cp[3], cp[3], false, false, nw->p_sep2[2]));
    ORCAI_CHECK(run_pool_res(c, TC, Y2, (long long)y2, (long long)2 * ws[2] * cp[2], 2 * cp[2], Y3, m, hs[2], ws[2], cp[2], cp[3], 2));
    net_mark(c, mk);  // 3
    if (stop == 3) { set_debug(nw, Y3, 0, m, hs[3], ws[3], 50, cp[3]); return ORCAI_OK; }
    ORCAI_CHECK(run_precise_sep(c, Y3, TA, TB, m, hs[3], ws[3], cp[3], cp[4], cp[4], true, true, nw->p_sep1[3]));
    ORCAI_CHECK(run_precise_sep(c, TB, TA, TC, m, hs[3], ws[3], cp[4], cp[4], cp[4], false, false, nw->p_sep2[3]));
    ORCAI_CHECK(run_pool_res(c, TC, Y3, (long long)y3, (long long)2 * ws[3] * cp[3], 2 * cp[3], Y4, m, hs[3], ws[3], cp[3], cp[4], 3));
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
