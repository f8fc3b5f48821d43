// orcai-V1 forward pass, fp32 CUDA-core path (snippet batcher + model.predict).
//
// Replaces the snippet copy and model.predict of the reference (src/orcAI/predict.py:252-268) on the
// graph of src/orcAI/architectures.py:120-241 (ResNetLSTM), inference mode:
//   Conv2D 3x3 1->16 + BN + ReLU
//   4 x [ReLU, SepConv, BN, ReLU, SepConv, BN, MaxPool(3,2)/2 "same", + Conv1x1/2(block input)]
//   SepConv 60->36 + BN + ReLU, reshape (w*36+c), 2 x BiLSTM(128), Dense128 ReLU, BN, Dense7 sigmoid.
// BatchNorm (eps 1e-3, moving statistics) is folded into the preceding pointwise / the following dense
// weights on the host in float64.  Activations are NHWC float32.  The snippet batcher is a strided
// view: snippet i starts at row (first+i)*shift of the device-resident spectrogram, nothing is copied,
// and the percentile clip + min-max normalisation is applied on load by the entry convolution.
//
// This is the full-precision path that carries the 1e-3 probability parity gate; the bf16 tensor-core
// path (net_tc.cu) is checked against it and against the oracle.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>

#include <atomic>

#include "common.h"
#include "net.h"

namespace orcai {

namespace {

constexpr float kTopDbF = 80.0f;
constexpr int kEntry = 16;
constexpr int kFinal = 36;
constexpr int kDense = 128;

// ------------------------------------------------------------------------------------------------
// entry convolution: normalise-on-load + Conv2D 3x3 (1->16) + folded BN + ReLU
// ------------------------------------------------------------------------------------------------
struct NormParams { const SelectState* st; };

__global__ void __launch_bounds__(256)
conv0_kernel(const float* __restrict__ in, int mode, long long first, int shift, int in_ld, int H, int Wf,
             const SelectState* __restrict__ st, const float* __restrict__ w, const float* __restrict__ bias,
             float* __restrict__ out, long long n_snip) {
  __shared__ float s_w[9 * kEntry + kEntry];
  for (int i = threadIdx.x; i < 9 * kEntry; i += blockDim.x) s_w[i] = w[i];
  for (int i = threadIdx.x; i < kEntry; i += blockDim.x) s_w[9 * kEntry + i] = bias[i];
  __syncthreads();
  float db_ref = 0.f, lo = 0.f, range = 1.f, hi = 1.f;
  if (mode == 0) { db_ref = st->db_ref; lo = st->lo; hi = st->hi; range = hi - lo; }
  const long long total = n_snip * H * Wf;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int wq = (int)(idx % Wf);
    const long long r = idx / Wf;
    const int h = (int)(r % H);
    const long long b = r / H;
    // mode 0: rows of the recording-wide raw dB buffer ; mode 1: materialised normalised snippets
    const long long row0 = (mode == 0) ? (first + b) * shift : b * (long long)H;
    float x[9];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int hh = h + dy, ww = wq + dx;
        float v = 0.f;
        if (hh >= 0 && hh < H && ww >= 0 && ww < Wf) {
          v = in[(size_t)(row0 + hh) * in_ld + ww];
          if (mode == 0) {
            v = fmaxf(v - db_ref, -kTopDbF);
            v = __fdiv_rn(fminf(fmaxf(v, lo), hi) - lo, range);
          }
        }
        x[(dy + 1) * 3 + dx + 1] = v;
      }
    float acc[kEntry];
#pragma unroll
    for (int c = 0; c < kEntry; ++c) acc[c] = s_w[9 * kEntry + c];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int c = 0; c < kEntry; ++c) acc[c] = fmaf(x[t], s_w[t * kEntry + c], acc[c]);
    float4* o = reinterpret_cast<float4*>(out + (size_t)idx * kEntry);
#pragma unroll
    for (int q = 0; q < kEntry / 4; ++q)
      o[q] = make_float4(fmaxf(acc[4 * q], 0.f), fmaxf(acc[4 * q + 1], 0.f), fmaxf(acc[4 * q + 2], 0.f), fmaxf(acc[4 * q + 3], 0.f));
  }
}

// ------------------------------------------------------------------------------------------------
// fused separable convolution: [ReLU] -> depthwise 3x3 "same" -> pointwise 1x1 + bias(+folded BN) -> [ReLU]
// one CTA = one 8x16 pixel tile of one snippet; 128 threads; thread = output pixel in the pointwise phase
// ------------------------------------------------------------------------------------------------
constexpr int kTH = 8, kTW = 16, kTPix = kTH * kTW;

template <int CIN, int COUT>
struct SepSmem {
  static constexpr int kHalo = (kTH + 2) * (kTW + 2);
  static constexpr int kDwStride = CIN | 1;               // odd stride: conflict-free per-pixel reads
  static constexpr int kCoPad = (COUT + 3) & ~3;
  static constexpr int kOutStride = COUT | 1;
  static constexpr size_t halo_f = (size_t)kHalo * CIN;
  static constexpr size_t out_f = (size_t)kTPix * kOutStride;
  static constexpr size_t region0 = halo_f > out_f ? halo_f : out_f;   // halo tile, later the output tile
  static constexpr size_t dw_f = (size_t)kTPix * kDwStride;
  static constexpr size_t w_f = (size_t)CIN * kCoPad + kCoPad + 9 * CIN;
  static constexpr size_t bytes = (region0 + dw_f + w_f) * sizeof(float);
};

template <int CIN, int COUT, bool RELU_IN, bool RELU_OUT>
__global__ void __launch_bounds__(kTPix)
sepconv_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W,
               const float* __restrict__ dwk, const float* __restrict__ pwk, const float* __restrict__ bias,
               int tiles_w, int tiles_h) {
  using S = SepSmem<CIN, COUT>;
  extern __shared__ __align__(16) float smem[];
  float* s_halo = smem;                      // [kHalo][CIN]
  float* s_out = smem;                       // [kTPix][kOutStride] (aliases the halo once it is consumed)
  float* s_dw = smem + S::region0;           // [kTPix][kDwStride]
  float* s_pw = s_dw + S::dw_f;              // [CIN][kCoPad]
  float* s_b = s_pw + (size_t)CIN * S::kCoPad;   // [kCoPad]
  float* s_dwk = s_b + S::kCoPad;            // [9][CIN]

  const int tid = threadIdx.x;
  int bid = blockIdx.x;
  const int tw = bid % tiles_w; bid /= tiles_w;
  const int th = bid % tiles_h; bid /= tiles_h;
  const long long b = bid;
  const int h0 = th * kTH, w0 = tw * kTW;
  const float* src = in + (size_t)b * H * W * CIN;

  for (int i = tid; i < CIN * S::kCoPad; i += kTPix) {
    const int ci = i / S::kCoPad, co = i % S::kCoPad;
    s_pw[i] = (co < COUT) ? pwk[ci * COUT + co] : 0.f;
  }
  for (int i = tid; i < S::kCoPad; i += kTPix) s_b[i] = (i < COUT) ? bias[i] : 0.f;
  for (int i = tid; i < 9 * CIN; i += kTPix) s_dwk[i] = dwk[i];
  for (int i = tid; i < S::kHalo * CIN; i += kTPix) {
    const int c = i % CIN, p = i / CIN;
    const int hh = h0 + p / (kTW + 2) - 1, ww = w0 + p % (kTW + 2) - 1;
    float v = 0.f;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = src[((size_t)hh * W + ww) * CIN + c];
    s_halo[i] = RELU_IN ? fmaxf(v, 0.f) : v;
  }
  __syncthreads();
  for (int i = tid; i < kTPix * CIN; i += kTPix) {
    const int c = i % CIN, p = i / CIN;
    const int py = p / kTW, px = p % kTW;
    float a = 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
        a = fmaf(s_halo[((py + dy) * (kTW + 2) + px + dx) * CIN + c], s_dwk[(dy * 3 + dx) * CIN + c], a);
    s_dw[p * S::kDwStride + c] = a;
  }
  __syncthreads();
  float acc[S::kCoPad];
#pragma unroll
  for (int co = 0; co < S::kCoPad; ++co) acc[co] = s_b[co];
  const float* xrow = s_dw + tid * S::kDwStride;
#pragma unroll 2
  for (int ci = 0; ci < CIN; ++ci) {
    const float x = xrow[ci];
    const float4* wrow = reinterpret_cast<const float4*>(s_pw + ci * S::kCoPad);
#pragma unroll
    for (int q = 0; q < S::kCoPad / 4; ++q) {
      const float4 w4 = wrow[q];
      acc[4 * q] = fmaf(x, w4.x, acc[4 * q]);
      acc[4 * q + 1] = fmaf(x, w4.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(x, w4.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(x, w4.w, acc[4 * q + 3]);
    }
  }
  // s_out aliases s_halo: every thread is past the depthwise phase (barrier above)
#pragma unroll
  for (int co = 0; co < COUT; ++co) s_out[tid * S::kOutStride + co] = RELU_OUT ? fmaxf(acc[co], 0.f) : acc[co];
  __syncthreads();
  float* dst = out + (size_t)b * H * W * COUT;
  for (int i = tid; i < kTPix * COUT; i += kTPix) {
    const int c = i % COUT, p = i / COUT;
    const int hh = h0 + p / kTW, ww = w0 + p % kTW;
    if (hh < H && ww < W) dst[((size_t)hh * W + ww) * COUT + c] = s_out[p * S::kOutStride + c];
  }
}

// ------------------------------------------------------------------------------------------------
// MaxPool (3,2) stride 2 "same" (pad at the end with -inf) + residual Conv1x1 stride 2 + add
// ------------------------------------------------------------------------------------------------
template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
pool_res_kernel(const float* __restrict__ t4, const float* __restrict__ prev, float* __restrict__ out, int H, int W,
                int Ho, int Wo, const float* __restrict__ rw, const float* __restrict__ rb, long long n_snip) {
  __shared__ float s_w[CIN * COUT + COUT];
  for (int i = threadIdx.x; i < CIN * COUT; i += blockDim.x) s_w[i] = rw[i];
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) s_w[CIN * COUT + i] = rb[i];
  __syncthreads();
  const long long total = n_snip * Ho * Wo * COUT;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % COUT);
    long long r = idx / COUT;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const long long b = r / Ho;
    const float* tb = t4 + (size_t)b * H * W * COUT;
    float m = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int hh = 2 * ho + dy, ww = 2 * wo + dx;
        if (hh < H && ww < W) m = fmaxf(m, tb[((size_t)hh * W + ww) * COUT + c]);
      }
    const float* pv = prev + ((size_t)b * H * W + (size_t)(2 * ho) * W + 2 * wo) * CIN;
    float a = s_w[CIN * COUT + c];
#pragma unroll 4
    for (int ci = 0; ci < CIN; ++ci) a = fmaf(pv[ci], s_w[ci * COUT + c], a);
    out[idx] = m + a;
  }
}

// ------------------------------------------------------------------------------------------------
// generic fp32 GEMM  C[M,N] = act(A[M,K] * B[K,N] + bias[N]) ; 64x64 tile, 256 threads, 4x4 per thread
// ------------------------------------------------------------------------------------------------
template <int ACT>  // 0 none, 1 relu, 2 sigmoid
__global__ void __launch_bounds__(256)
gemm_bias_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ bias,
                 float* __restrict__ C, int M, int N, int K) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float sA[BK][BM + 1];
  __shared__ float sB[BK][BN + 1];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int i = tid; i < BM * BK; i += 256) {
      const int kk = i % BK, mm = i / BK;
      const int gm = m0 + mm, gk = k0 + kk;
      sA[kk][mm] = (gm < M && gk < K) ? A[(size_t)gm * K + gk] : 0.f;
    }
    for (int i = tid; i < BK * BN; i += 256) {
      const int nn = i % BN, kk = i / BN;
      const int gn = n0 + nn, gk = k0 + kk;
      sB[kk][nn] = (gn < N && gk < K) ? B[(size_t)gk * N + gn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
      if (gm < M && gn < N) {
        float v = acc[i][j] + bias[gn];
        if (ACT == 1) v = fmaxf(v, 0.f);
        if (ACT == 2) v = 1.0f / (1.0f + expf(-v));
        C[(size_t)gm * N + gn] = v;
      }
    }
}

// ------------------------------------------------------------------------------------------------
// LSTM recurrence (Keras cell: gates i,f,c,o ; sigmoid / tanh), one CTA = kSN snippets x one direction
//   xz  : (n, Tn, 2*4U)  input projections + bias, forward gates then backward gates
//   whh : (2, U, 4U)
//   out : (n, Tn, 2U)    forward h in [0,U), backward h in [U,2U)
// ------------------------------------------------------------------------------------------------
constexpr int kSN = 16;   // snippets per CTA: every step re-reads W_hh (256 KB) from L2, so more snippets per CTA = less traffic (8 -> 16: 1.35 -> ? ms per layer and hour)

template <int U>
__global__ void __launch_bounds__(4 * U)
lstm_rec_kernel(const float* __restrict__ xz, const float* __restrict__ whh, float* __restrict__ out, long long n, int Tn) {
  constexpr int G = 4 * U;
  __shared__ __align__(16) float s_h[kSN][U];
  __shared__ float s_z[kSN][G];
  const int dir = blockIdx.y;
  const long long b0 = (long long)blockIdx.x * kSN;
  const int g = threadIdx.x;
  const float* w = whh + (size_t)dir * U * G;
  for (int i = g; i < kSN * U; i += G) (&s_h[0][0])[i] = 0.f;
  float cst[kSN * U / G];  // cell states owned by this thread
#pragma unroll
  for (int q = 0; q < kSN * U / G; ++q) cst[q] = 0.f;
  __syncthreads();
  for (int step = 0; step < Tn; ++step) {
    const int t = dir ? (Tn - 1 - step) : step;
    float acc[kSN];
#pragma unroll
    for (int s = 0; s < kSN; ++s) {
      const long long b = b0 + s;
      acc[s] = (b < n) ? xz[((size_t)b * Tn + t) * (2 * G) + (size_t)dir * G + g] : 0.f;
    }
    // four hidden units per iteration: one broadcast LDS.128 of h serves four FMAs (the scalar form issued one shared-memory
    // load per FMA and was bound by the load/store unit); per accumulator the additions stay in ascending j order
#pragma unroll 2
    for (int j = 0; j < U; j += 4) {
      const float w0 = __ldg(w + (size_t)(j + 0) * G + g), w1 = __ldg(w + (size_t)(j + 1) * G + g);
      const float w2 = __ldg(w + (size_t)(j + 2) * G + g), w3 = __ldg(w + (size_t)(j + 3) * G + g);
#pragma unroll
      for (int s = 0; s < kSN; ++s) {
        const float4 h4 = *reinterpret_cast<const float4*>(&s_h[s][j]);
        acc[s] = fmaf(h4.x, w0, acc[s]);
        acc[s] = fmaf(h4.y, w1, acc[s]);
        acc[s] = fmaf(h4.z, w2, acc[s]);
        acc[s] = fmaf(h4.w, w3, acc[s]);
      }
    }
#pragma unroll
    for (int s = 0; s < kSN; ++s) s_z[s][g] = acc[s];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kSN * U / G; ++q) {
      const int item = g + q * G;          // (snippet s, unit u)
      const int s = item / U, u = item % U;
      const float zi = s_z[s][u], zf = s_z[s][U + u], zc = s_z[s][2 * U + u], zo = s_z[s][3 * U + u];
      const float ig = 1.0f / (1.0f + expf(-zi));
      const float fg = 1.0f / (1.0f + expf(-zf));
      const float og = 1.0f / (1.0f + expf(-zo));
      const float cn = fg * cst[q] + ig * tanhf(zc);
      cst[q] = cn;
      const float h = og * tanhf(cn);
      s_h[s][u] = h;
      const long long b = b0 + s;
      if (b < n) out[((size_t)b * Tn + t) * (2 * U) + (size_t)dir * U + u] = h;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int CIN, int COUT, bool RI, bool RO>
int run_sepconv(Ctx* c, const float* in, float* out, long long n, int H, int W, const NetWeights::Sep& s) {
  using S = SepSmem<CIN, COUT>;
  // the attribute is per device: remember which devices have it (several contexts can live in one process)
  static std::atomic<unsigned long long> attr_devices{0ull};
  if (!((attr_devices.load() >> (c->device & 63)) & 1ull)) {
    ORCAI_CUDA(c, cudaFuncSetAttribute(sepconv_kernel<CIN, COUT, RI, RO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
    attr_devices.fetch_or(1ull << (c->device & 63));
  }
  const int tiles_w = (W + kTW - 1) / kTW, tiles_h = (H + kTH - 1) / kTH;
  const long long grid = n * tiles_w * tiles_h;
  sepconv_kernel<CIN, COUT, RI, RO><<<(unsigned)grid, kTPix, S::bytes, c->stream>>>(in, out, H, W, s.dw, s.pw, s.b, tiles_w, tiles_h);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

template <int CIN, int COUT>
int run_block(Ctx* c, const float* prev, float* t_a, float* t_b, float* next, long long n, int H, int W, int blk) {
  NetWeights* nw = c->net;
  ORCAI_CHECK((run_sepconv<CIN, COUT, true, true>(c, prev, t_a, n, H, W, nw->sep1[blk])));
  ORCAI_CHECK((run_sepconv<COUT, COUT, false, false>(c, t_a, t_b, n, H, W, nw->sep2[blk])));
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long long total = n * Ho * Wo * COUT;
  long long grid = (total + 255) / 256;
  if (grid > (long long)c->sm_count * 32) grid = (long long)c->sm_count * 32;
  pool_res_kernel<CIN, COUT><<<(unsigned)grid, 256, 0, c->stream>>>(t_b, prev, next, H, W, Ho, Wo, nw->res_w[blk], nw->res_b[blk], n);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

template <int ACT>
int run_gemm(Ctx* c, const float* A, const float* B, const float* bias, float* C, long long M, int N, int K) {
  dim3 grid((N + 63) / 64, (unsigned)((M + 63) / 64));
  gemm_bias_kernel<ACT><<<grid, 256, 0, c->stream>>>(A, B, bias, C, (int)M, N, K);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

size_t per_snippet_floats(const NetWeights* nw) {
  // two "prev" ping-pong buffers and two block temporaries, then the LSTM / dense stage buffers (which reuse them)
  const size_t H = nw->H, W = nw->Wf;
  size_t prev_max = H * W * kEntry;
  size_t t_max = 0;
  size_t h = H, w = W;
  for (int b = 0; b < nw->n_blocks; ++b) {
    t_max = std::max(t_max, h * w * (size_t)nw->filters[b]);
    h = (h + 1) / 2; w = (w + 1) / 2;
    prev_max = std::max(prev_max, h * w * (size_t)nw->filters[b]);
  }
  return 2 * prev_max + 2 * t_max;
}

}  // namespace

int net_upload(Ctx* c, const std::vector<float>& v, float** dptr) {
  void* p = nullptr;
  ORCAI_CUDA(c, cudaMalloc(&p, v.size() * sizeof(float)));
  c->net->allocs.push_back(p);
  ORCAI_CUDA(c, cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  *dptr = static_cast<float*>(p);
  return ORCAI_OK;
}

namespace {

int upload(Ctx* c, const std::vector<float>& v, float** dptr) { return net_upload(c, v, dptr); }

struct HostTensors {
  std::map<std::string, std::pair<const float*, int64_t>> m;
  Ctx* c;
  const float* get(const std::string& name, int64_t expect) {
    auto it = m.find(name);
    if (it == m.end()) { c->err = "missing weight '" + name + "'"; return nullptr; }
    if (it->second.second != expect) {
      c->err = "weight '" + name + "' has " + std::to_string(it->second.second) + " elements, expected " + std::to_string(expect);
      return nullptr;
    }
    return it->second.first;
  }
};

// scale/shift of an inference BatchNorm: y = x*s + t
bool bn_fold(HostTensors& ht, const std::string& prefix, int C, std::vector<double>* s, std::vector<double>* t) {
  const float* g = ht.get(prefix + "/gamma", C);
  const float* b = ht.get(prefix + "/beta", C);
  const float* m = ht.get(prefix + "/moving_mean", C);
  const float* v = ht.get(prefix + "/moving_variance", C);
  if (!g || !b || !m || !v) return false;
  s->resize(C); t->resize(C);
  for (int i = 0; i < C; ++i) {
    const double sc = (double)g[i] / std::sqrt((double)v[i] + 1e-3);
    (*s)[i] = sc;
    (*t)[i] = (double)b[i] - (double)m[i] * sc;
  }
  return true;
}

int load_sep(Ctx* c, HostTensors& ht, const std::string& sp, const std::string& bnp, int ci, int co, NetWeights::Sep* out,
             NetWeights::HostSep* keep) {
  const float* dw = ht.get(sp + "/depthwise_kernel", 9LL * ci);
  const float* pw = ht.get(sp + "/pointwise_kernel", (int64_t)ci * co);
  const float* b = ht.get(sp + "/bias", co);
  std::vector<double> s, t;
  if (!dw || !pw || !b || !bn_fold(ht, bnp, co, &s, &t)) return ORCAI_ERR_ARG;
  std::vector<float> dwv(dw, dw + 9 * ci);                 // (3,3,ci,1) == [tap][ci]
  std::vector<float> pwv((size_t)ci * co), bv(co);
  for (int i = 0; i < ci; ++i)
    for (int o = 0; o < co; ++o) pwv[(size_t)i * co + o] = (float)((double)pw[(size_t)i * co + o] * s[o]);
  for (int o = 0; o < co; ++o) bv[o] = (float)((double)b[o] * s[o] + t[o]);
  out->ci = ci; out->co = co;
  keep->dw = dwv; keep->pw = pwv; keep->b = bv; keep->ci = ci; keep->co = co;
  ORCAI_CHECK(upload(c, dwv, &out->dw));
  ORCAI_CHECK(upload(c, pwv, &out->pw));
  ORCAI_CHECK(upload(c, bv, &out->b));
  return ORCAI_OK;
}

}  // namespace

int net_create(Ctx* c) {
  c->net = new NetWeights();
  for (auto& e : c->net->ev) ORCAI_CUDA(c, cudaEventCreate(&e));
  return ORCAI_OK;
}

// stage times of the first chunk of the last forward; call after the stream has been synchronised
void net_collect_stage_times(Ctx* c) {
  net_trace_dump();
  NetWeights* nw = c->net;
  if (!nw || nw->marked_snippets == 0) return;
  for (int i = 0; i + 1 < NetWeights::kNumMarks; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, nw->ev[i], nw->ev[i + 1]) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
    c->tm.net_stage_ms[i] = ms;
  }
  c->tm.net_stage_ms[15] = (float)nw->marked_snippets;
}

void net_destroy(Ctx* c) {
  if (!c->net) return;
  for (void* p : c->net->allocs) cudaFree(p);
  if (c->net->ws) cudaFree(c->net->ws);
  for (auto& e : c->net->ev) if (e) cudaEventDestroy(e);
  delete c->net;
  c->net = nullptr;
}

int net_set_path(Ctx* c, int path) {
  c->net->path = path;
  return ORCAI_OK;
}

int net_set_tail_path(Ctx* c, int path) {
  if (path < 0 || path > 1) ORCAI_FAIL(c, ORCAI_ERR_ARG, "tail_path must be 0 (fp32) or 1 (tensor cores)");
  c->net->tail_path = path;
  return ORCAI_OK;
}

int net_set_block1_path(Ctx* c, int path) {
  if (path < 0 || path > 1) ORCAI_FAIL(c, ORCAI_ERR_ARG, "block1_path must be 0 (one MMA per tap) or 1 (N-widened MMAs)");
  c->net->block1_path = path;
  return ORCAI_OK;
}

int net_set_precise_tall(Ctx* c, int on) {   // 0 / 1: precise_tall; 2 / 3: precise_sep_path 0 / 1; 4 / 5: precise_lstm_tc 0 / 1
  if (on >= 4) c->net->precise_lstm_tc = on - 4;
  else if (on >= 2) c->net->precise_sep_path = on - 2;
  else c->net->precise_tall = on ? 1 : 0;
  return ORCAI_OK;
}

int net_set_conv0_path(Ctx* c, int path) {
  if (path < 0 || path > 2) ORCAI_FAIL(c, ORCAI_ERR_ARG, "conv0_path must be 0 (fp32 CUDA cores), 1 (tensor cores) or 2 (fused into block 1)");
  c->net->conv0_path = path;
  return ORCAI_OK;
}

int net_set_debug_stop(Ctx* c, int stage) {
  c->net->debug_stop = stage;
  return ORCAI_OK;
}

int net_debug_read(Ctx* c, float* out_host, int64_t capacity, int64_t* dims_out) {
  NetWeights* nw = c->net;
  if (!nw->dbg_ptr) ORCAI_FAIL(c, ORCAI_ERR_STATE, "no debug buffer recorded (set debug_stop, then run a forward)");
  ORCAI_CUDA(c, cudaStreamSynchronize(c->stream));
  const long long npix = nw->dbg_n * nw->dbg_h * nw->dbg_w;
  dims_out[0] = nw->dbg_n; dims_out[1] = nw->dbg_h; dims_out[2] = nw->dbg_w; dims_out[3] = nw->dbg_c;
  if (npix * nw->dbg_c > capacity) ORCAI_FAIL(c, ORCAI_ERR_CAPACITY, "debug buffer needs %lld floats", npix * nw->dbg_c);
  const size_t esz = nw->dbg_kind == 0 ? 4 : 2;
  std::vector<unsigned char> raw((size_t)npix * nw->dbg_pitch * esz);
  ORCAI_CUDA(c, cudaMemcpy(raw.data(), nw->dbg_ptr, raw.size(), cudaMemcpyDeviceToHost));
  for (long long p = 0; p < npix; ++p)
    for (int ch = 0; ch < nw->dbg_c; ++ch) {
      const size_t i = (size_t)p * nw->dbg_pitch + ch;
      float v;
      if (nw->dbg_kind == 0) {
        v = reinterpret_cast<const float*>(raw.data())[i];
      } else {
        const unsigned short u = reinterpret_cast<const unsigned short*>(raw.data())[i];
        if (nw->dbg_kind == 2) {  // bf16
          const unsigned int w = (unsigned int)u << 16;
          memcpy(&v, &w, 4);
        } else {                  // fp16
          const unsigned int sign = (u >> 15) & 1u, ex = (u >> 10) & 0x1fu, man = u & 0x3ffu;
          if (ex == 0) v = std::ldexp((float)man, -24);
          else if (ex == 31) v = man ? NAN : INFINITY;
          else v = std::ldexp((float)(man | 0x400u), (int)ex - 25);
          if (sign) v = -v;
        }
      }
      out_host[(size_t)p * nw->dbg_c + ch] = v;
    }
  return ORCAI_OK;
}

int net_set_chunk(Ctx* c, int chunk) {
  if (chunk < 1) ORCAI_FAIL(c, ORCAI_ERR_ARG, "chunk must be >= 1");
  c->net->chunk = chunk;
  c->net->chunk_fused = chunk;
  c->net->chunk_precise = chunk;
  return ORCAI_OK;
}

int net_load_weights(Ctx* c, const char* const* names, const float* const* data, const int64_t* sizes, int n) {
  NetWeights* nw = c->net;
  for (void* p : nw->allocs) cudaFree(p);
  nw->allocs.clear();
  nw->loaded = false;
  nw->tc_ready[0] = nw->tc_ready[1] = false;
  nw->fused_ready = false;
  nw->tail_tc_ready = false;
  nw->precise_ready = nw->tail_precise_ready = false;
  nw->calib = Calib();
  const orcai_params& P = c->p;
  if (P.n_blocks != 4 || P.filters[0] != 30 || P.filters[1] != 40 || P.filters[2] != 50 || P.filters[3] != 60 ||
      P.kernel_size != 3 || P.lstm_units != 128)
    ORCAI_FAIL(c, ORCAI_ERR_ARG, "network kernels are built for the orcai-V1 shape (filters 30/40/50/60, k=3, 128 LSTM units)");
  if (P.snippet_len % (1 << P.n_blocks) != 0)
    ORCAI_FAIL(c, ORCAI_ERR_ARG, "snippet length must be a multiple of 2^n_blocks (TF 'same' pooling pads only at the end for even heights)");
  nw->n_blocks = P.n_blocks;
  for (int i = 0; i < P.n_blocks; ++i) nw->filters[i] = P.filters[i];
  nw->H = P.snippet_len; nw->Wf = P.n_freq; nw->U = P.lstm_units; nw->L = P.n_labels;
  HostTensors ht; ht.c = c;
  for (int i = 0; i < n; ++i) ht.m[names[i]] = {data[i], sizes[i]};

  {  // entry conv + bn0
    const float* k = ht.get("conv0/kernel", 9 * kEntry);
    const float* b = ht.get("conv0/bias", kEntry);
    std::vector<double> s, t;
    if (!k || !b || !bn_fold(ht, "bn0", kEntry, &s, &t)) return ORCAI_ERR_ARG;
    std::vector<float> kv(9 * kEntry), bv(kEntry);
    for (int tap = 0; tap < 9; ++tap)
      for (int o = 0; o < kEntry; ++o) kv[tap * kEntry + o] = (float)((double)k[tap * kEntry + o] * s[o]);
    for (int o = 0; o < kEntry; ++o) bv[o] = (float)((double)b[o] * s[o] + t[o]);
    nw->h_conv0_w = kv; nw->h_conv0_b = bv;
    ORCAI_CHECK(upload(c, kv, &nw->conv0_w));
    ORCAI_CHECK(upload(c, bv, &nw->conv0_b));
  }
  int ci = kEntry, w = nw->Wf;
  for (int b = 0; b < nw->n_blocks; ++b) {
    const int co = nw->filters[b];
    const std::string p = "block" + std::to_string(b + 1);
    ORCAI_CHECK(load_sep(c, ht, p + "/sep1", p + "/bn1", ci, co, &nw->sep1[b], &nw->h_sep1[b]));
    ORCAI_CHECK(load_sep(c, ht, p + "/sep2", p + "/bn2", co, co, &nw->sep2[b], &nw->h_sep2[b]));
    const float* rk = ht.get(p + "/res/kernel", (int64_t)ci * co);
    const float* rb = ht.get(p + "/res/bias", co);
    if (!rk || !rb) return ORCAI_ERR_ARG;
    nw->h_res_w[b].assign(rk, rk + (size_t)ci * co);
    nw->h_res_b[b].assign(rb, rb + co);
    ORCAI_CHECK(upload(c, nw->h_res_w[b], &nw->res_w[b]));
    ORCAI_CHECK(upload(c, nw->h_res_b[b], &nw->res_b[b]));
    ci = co;
    w = (w + 1) / 2;
  }
  ORCAI_CHECK(load_sep(c, ht, "final/sep", "final/bn", ci, kFinal, &nw->fin, &nw->h_fin));
  nw->feat = w * kFinal;
  const int U = nw->U, G = 4 * U;
  for (int l = 0; l < 2; ++l) {
    const int I = l == 0 ? nw->feat : 2 * U;
    const std::string p = "lstm" + std::to_string(l + 1);
    std::vector<float> wih((size_t)I * 2 * G), bih(2 * G), whh((size_t)2 * U * G);
    for (int d = 0; d < 2; ++d) {
      const std::string q = p + (d ? "/backward" : "/forward");
      const float* k = ht.get(q + "/kernel", (int64_t)I * G);
      const float* r = ht.get(q + "/recurrent_kernel", (int64_t)U * G);
      const float* bb = ht.get(q + "/bias", G);
      if (!k || !r || !bb) return ORCAI_ERR_ARG;
      for (int i = 0; i < I; ++i) memcpy(&wih[(size_t)i * 2 * G + (size_t)d * G], k + (size_t)i * G, G * sizeof(float));
      memcpy(&bih[(size_t)d * G], bb, G * sizeof(float));
      memcpy(&whh[(size_t)d * U * G], r, (size_t)U * G * sizeof(float));
    }
    nw->h_lstm_wih[l] = wih;
    nw->h_lstm_whh[l] = whh;
    nw->h_lstm_bih[l] = bih;
    ORCAI_CHECK(upload(c, wih, &nw->lstm_wih[l]));
    ORCAI_CHECK(upload(c, bih, &nw->lstm_bih[l]));
    ORCAI_CHECK(upload(c, whh, &nw->lstm_whh[l]));
  }
  {  // dense head; bn_dense (after the ReLU) folds into dense2
    const float* k1 = ht.get("dense1/kernel", (int64_t)2 * U * kDense);
    const float* b1 = ht.get("dense1/bias", kDense);
    const float* k2 = ht.get("dense2/kernel", (int64_t)kDense * nw->L);
    const float* b2 = ht.get("dense2/bias", nw->L);
    std::vector<double> s, t;
    if (!k1 || !b1 || !k2 || !b2 || !bn_fold(ht, "bn_dense", kDense, &s, &t)) return ORCAI_ERR_ARG;
    nw->h_d1_w.assign(k1, k1 + (size_t)2 * U * kDense);
    ORCAI_CHECK(upload(c, nw->h_d1_w, &nw->d1_w));
    nw->h_d1_b.assign(b1, b1 + kDense);
    ORCAI_CHECK(upload(c, nw->h_d1_b, &nw->d1_b));
    std::vector<float> w2((size_t)kDense * nw->L), bb2(nw->L);
    for (int o = 0; o < nw->L; ++o) {
      double acc = b2[o];
      for (int i = 0; i < kDense; ++i) {
        w2[(size_t)i * nw->L + o] = (float)(s[i] * (double)k2[(size_t)i * nw->L + o]);
        acc += t[i] * (double)k2[(size_t)i * nw->L + o];
      }
      bb2[o] = (float)acc;
    }
    ORCAI_CHECK(upload(c, w2, &nw->d2_w));
    ORCAI_CHECK(upload(c, bb2, &nw->d2_b));
  }
  nw->loaded = true;
  return ORCAI_OK;
}

int net_lstm_rec_fp32(Ctx* c, const float* xz, const float* whh, float* out, long long m, int Tn) {
  if (m <= 0) return ORCAI_OK;
  dim3 grid((unsigned)((m + kSN - 1) / kSN), 2);
  lstm_rec_kernel<128><<<grid, 512, 0, c->stream>>>(xz, whh, out, m, Tn);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

// LSTM x2 + dense head on fp32 features; shared by both network paths
int net_tail_fp32(Ctx* c, const float* feat, float* scratch, long long m, float* d_preds_out, bool mk) {
  NetWeights* nw = c->net;
  const int U = nw->U, G = 4 * U, L = nw->L;
  const int Tn = nw->H >> nw->n_blocks;
  const long long rows = m * Tn;
  float* xz = scratch;                          // (rows, 2G)
  float* h1 = xz + (size_t)rows * 2 * G;        // (rows, 2U)
  float* h2 = h1 + (size_t)rows * 2 * U;        // (rows, 2U)
  float* d1 = h2 + (size_t)rows * 2 * U;        // (rows, 128)
  auto mark = [&](bool on) { net_mark(c, on); };
  ORCAI_CHECK((run_gemm<0>(c, feat, nw->lstm_wih[0], nw->lstm_bih[0], xz, rows, 2 * G, nw->feat)));
  mark(mk);  // 6: lstm1 input projection
  {
    dim3 grid((unsigned)((m + kSN - 1) / kSN), 2);
    lstm_rec_kernel<128><<<grid, 512, 0, c->stream>>>(xz, nw->lstm_whh[0], h1, m, Tn);
    c->launches++;
  }
  mark(mk);  // 7: lstm1 recurrence
  ORCAI_CHECK((run_gemm<0>(c, h1, nw->lstm_wih[1], nw->lstm_bih[1], xz, rows, 2 * G, 2 * U)));
  mark(mk);  // 8: lstm2 input projection
  {
    dim3 grid((unsigned)((m + kSN - 1) / kSN), 2);
    lstm_rec_kernel<128><<<grid, 512, 0, c->stream>>>(xz, nw->lstm_whh[1], h2, m, Tn);
    c->launches++;
  }
  mark(mk);  // 9: lstm2 recurrence
  ORCAI_CHECK((run_gemm<1>(c, h2, nw->d1_w, nw->d1_b, d1, rows, kDense, 2 * U)));
  ORCAI_CHECK((run_gemm<2>(c, d1, nw->d2_w, nw->d2_b, d_preds_out, rows, L, kDense)));
  mark(mk);  // 10: dense head
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

// A recording whose clipped dB range is empty (silence: lo == hi) normalises to 0/0 = NaN everywhere in the reference
// (spectrogram.py:81-83); Keras propagates the NaNs to every probability and nothing gets labelled.  CUDA's fmaxf / __hmax2
// ReLU would silently turn NaN into 0, so the all-or-nothing case is restored here: NaN probabilities for a degenerate range.
__global__ void poison_degenerate_kernel(float* __restrict__ preds, long long count, const SelectState* __restrict__ st) {
  const float range = st->hi - st->lo;
  if (range > 0.f) return;   // false for 0 and for NaN
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    preds[i] = __int_as_float(0x7fc00000);
}

static int net_forward_paths(Ctx* c, const float* d_in, int input_mode, int64_t first, int64_t n, float* d_preds);

int net_forward(Ctx* c, const float* d_in, int input_mode, int64_t first, int64_t n, float* d_preds) {
  ORCAI_CHECK(net_forward_paths(c, d_in, input_mode, first, n, d_preds));
  if (input_mode == 0 && n > 0 && c->net->debug_stop < 0) {
    const long long count = (long long)n * (c->net->H >> c->net->n_blocks) * c->net->L;
    poison_degenerate_kernel<<<(unsigned)std::min<long long>((count + 255) / 256, 1024), 256, 0, c->stream>>>(d_preds, count, c->d_sel);
    c->launches++;
    ORCAI_CUDA(c, cudaGetLastError());
  }
  return ORCAI_OK;
}

static int net_forward_paths(Ctx* c, const float* d_in, int input_mode, int64_t first, int64_t n, float* d_preds) {
  NetWeights* nw = c->net;
  if (!nw || !nw->loaded) ORCAI_FAIL(c, ORCAI_ERR_STATE, "no weights loaded (orcai_load_weights)");
  if (nw->path != 0) return net_forward_tc(c, d_in, input_mode, first, n, d_preds);
  const int H = nw->H, Wf = nw->Wf, L = nw->L;
  const int Tn = H >> nw->n_blocks;
  const size_t per = per_snippet_floats(nw);
  const long long chunk = std::min<long long>(nw->chunk, n);
  if (chunk <= 0) return ORCAI_OK;
  {
    void* p = nw->ws;
    ORCAI_CHECK(ensure_device_buffer(c, &p, &nw->ws_cap, per * (size_t)chunk * sizeof(float)));
    nw->ws = static_cast<float*>(p);
  }
  // workspace carving
  size_t prev_max = (size_t)H * Wf * kEntry, t_max = 0;
  {
    size_t h = H, w = Wf;
    for (int b = 0; b < nw->n_blocks; ++b) {
      t_max = std::max(t_max, h * w * (size_t)nw->filters[b]);
      h = (h + 1) / 2; w = (w + 1) / 2;
      prev_max = std::max(prev_max, h * w * (size_t)nw->filters[b]);
    }
  }
  float* pA = nw->ws;
  float* pB = pA + prev_max * chunk;
  float* tA = pB + prev_max * chunk;
  float* tB = tA + t_max * chunk;
  const int shift = c->p.snippet_len / 2;

  nw->mark_i = 0;
  auto mark = [&](bool on) { net_mark(c, on); };
  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    const long long m = std::min<long long>(chunk, n - s0);
    const bool mk = (s0 == 0);
    if (mk) nw->marked_snippets = m;
    mark(mk);
    {  // entry conv
      const long long total = m * H * Wf;
      long long grid = (total + 255) / 256;
      if (grid > (long long)c->sm_count * 16) grid = (long long)c->sm_count * 16;
      const float* src = (input_mode == 0) ? d_in : d_in + (size_t)s0 * H * Wf;
      conv0_kernel<<<(unsigned)grid, 256, 0, c->stream>>>(src, input_mode, first + s0, shift,
                                                           input_mode == 0 ? kRawLd : Wf, H, Wf, c->d_sel,
                                                           nw->conv0_w, nw->conv0_b, pA, m);
      c->launches++;
      ORCAI_CUDA(c, cudaGetLastError());
    }
    mark(mk);  // 0: conv0
    int h = H, w = Wf;
    ORCAI_CHECK((run_block<16, 30>(c, pA, tA, tB, pB, m, h, w, 0))); h = (h + 1) / 2; w = (w + 1) / 2;
    mark(mk);  // 1: block1
    ORCAI_CHECK((run_block<30, 40>(c, pB, tA, tB, pA, m, h, w, 1))); h = (h + 1) / 2; w = (w + 1) / 2;
    mark(mk);  // 2: block2
    ORCAI_CHECK((run_block<40, 50>(c, pA, tA, tB, pB, m, h, w, 2))); h = (h + 1) / 2; w = (w + 1) / 2;
    mark(mk);  // 3: block3
    ORCAI_CHECK((run_block<50, 60>(c, pB, tA, tB, pA, m, h, w, 3))); h = (h + 1) / 2; w = (w + 1) / 2;
    mark(mk);  // 4: block4
    // final separable conv -> features (m, Tn, w*36) ; NHWC flattening is already w*36+c
    float* feat = pB;
    ORCAI_CHECK((run_sepconv<60, 36, false, true>(c, pA, feat, m, h, w, nw->fin)));
    mark(mk);  // 5: final sepconv
    ORCAI_CHECK(net_tail_fp32(c, feat, tA, m, d_preds + (size_t)s0 * Tn * L, mk));
    ORCAI_CUDA(c, cudaGetLastError());
  }
  return ORCAI_OK;
}

}  // namespace orcai
