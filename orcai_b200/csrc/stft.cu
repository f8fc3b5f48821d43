// K1: fused window + 512-point real FFT + |.|^2 + 10*log10 + band crop + global power max.
//
// Replaces librosa.stft / amplitude_to_db at src/orcAI/spectrogram.py:34-39 and :51-53 of the
// reference (the "ref = np.max" shift and the top_db floor are applied by the consumers from the
// exact global maximum this kernel produces, so no second pass over the 257-bin array exists).
//
// Two kernels: stft_db16_kernel (float64 default: 16 threads per frame, 2 frames per warp, stft_core16.cuh; further down) and
// stft_db_kernel (float32 variant and the round-1 float64 decomposition, option stft_threads = 8), described here.
// Mapping: 8 threads per frame, 4 frames per warp, 8 warps per CTA; each warp walks groups of 4
// consecutive frames (the 50% overlap makes the second read of every sample an L1 hit).  All FFT
// butterflies are register-resident with immediate twiddles (fft_gen.cuh); the only exchange is one
// warp-private shared-memory transpose (stft_core.cuh).  HBM traffic per frame: 256 new samples in,
// band_hi-band_lo floats out.
//
// Two arithmetic variants of the same code:
//   RealT = double : the FFT runs in float64 like the reference's numpy.fft.rfft and is rounded to
//                    complex64 before |.|^2 and log10f - parity grade (default).
//   RealT = float  : float32 FFT, lg2.approx for the logarithm - the fast variant (tail cells that
//                    sit > 60 dB below their own frame's peak can be off by up to ~2e-3 dB).
#include <atomic>

#include "common.h"
#include "stft_core.cuh"
#include "stft_core16.cuh"
#include "stft_tables.h"

namespace orcai {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kFramesPerWarp = 4;
constexpr int kTableCx = 768;

template <typename RealT>
constexpr size_t stft_smem_bytes() {
  return (size_t)kTableCx * sizeof(Cx<RealT>) + (size_t)kWarpsPerCta * kFramesPerWarp * kFrameBufCx * sizeof(Cx<RealT>);
}

__device__ __forceinline__ Cx<float> ld_pair(const float* p) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return Cx<float>{v.x, v.y};
}
__device__ __forceinline__ Cx<float> ld_pair(const int16_t* p) {
  // two PCM16 samples -> exact float integers via the 1.5*2^23 magic constant (no I2F on the slow pipe)
  const unsigned int u = __ldg(reinterpret_cast<const unsigned int*>(p));
  const int lo = (int)(short)(u & 0xffffu);
  const int hi = ((int)u) >> 16;
  Cx<float> r;
  r.x = __int_as_float(0x4B400000 + lo) - 12582912.0f;
  r.y = __int_as_float(0x4B400000 + hi) - 12582912.0f;
  return r;
}
__device__ __forceinline__ float ld_one(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_one(const int16_t* p) { return (float)__ldg(p); }

template <typename SampleT, typename RealT>
__global__ void __launch_bounds__(kWarpsPerCta * 32, sizeof(RealT) == 4 ? 2 : 1)
stft_db_kernel(const SampleT* __restrict__ pcm, long long n_samples, long long T, float* __restrict__ raw,
               int ld, int band_lo, int band_hi, const Cx<RealT>* __restrict__ tables,
               unsigned int* __restrict__ pmax_bits, long long stat_row0, long long stat_row1) {
  // dB through MUFU lg2 in both variants: |error| <= 3e-5 dB, far inside the 1e-3 dB gate, and ~15 % fewer instructions than
  // log10f in a kernel that is bound by instruction issue
  constexpr bool kPrecise = false;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cx<RealT>* s_tab = reinterpret_cast<Cx<RealT>*>(smem_raw);
  Cx<RealT>* s_buf = s_tab + kTableCx;
  for (int i = threadIdx.x; i < kTableCx; i += blockDim.x) s_tab[i] = tables[i];
  __syncthreads();
  const StftTables<RealT> tb{s_tab, s_tab + 256, s_tab + 512};

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t = lane & 7;
  const int fl = lane >> 3;
  Cx<RealT>* fbuf = s_buf + (size_t)(warp * kFramesPerWarp + fl) * kFrameBufCx;

  const long long n_groups = (T + kFramesPerWarp - 1) / kFramesPerWarp;
  const long long g_stride = (long long)gridDim.x * kWarpsPerCta;
  float pmax = 0.0f;
  constexpr int BR[32] = {ORCAI_BITREV32_LIST};

  for (long long g = (long long)blockIdx.x * kWarpsPerCta + warp; g < n_groups; g += g_stride) {
    const long long j = g * kFramesPerWarp + fl;
    const long long base = (j - 1) * kHop;  // first sample of frame j (centre padding of n_fft/2)
    Cx<float> x[32];
    // warp-uniform: all 4 frames of the group fully inside the recording
    const long long j0 = g * kFramesPerWarp;
    const bool interior = (j0 >= 1) && ((j0 + kFramesPerWarp) * kHop <= n_samples);
    if (interior) {
      const SampleT* src = pcm + base + 2 * t;
#pragma unroll
      for (int p = 0; p < 32; ++p) x[p] = ld_pair(src + 16 * BR[p]);
    } else {
#pragma unroll
      for (int p = 0; p < 32; ++p) {
        const long long s0 = base + 2 * (8 * BR[p] + t);
        x[p].x = (s0 >= 0 && s0 < n_samples) ? ld_one(pcm + s0) : 0.0f;
        x[p].y = (s0 + 1 >= 0 && s0 + 1 < n_samples) ? ld_one(pcm + s0 + 1) : 0.0f;
      }
    }
    stage_a<RealT>(x, t, tb, fbuf);
    __syncwarp();
    const bool live = j < T;
    float* out = raw + (size_t)(live ? j : 0) * ld - band_lo;
    stage_b<RealT>(t, tb, fbuf, [&](int k, RealT re, RealT im) {
      const float fr = (float)re, fi = (float)im;  // complex64 rounding of the reference's stft matrix (no-op for RealT = float)
      const float pw = fmaf(fr, fr, fi * fi);
      pmax = fmaxf(pmax, (live && j >= stat_row0 && j < stat_row1) ? pw : 0.0f);   // time chunks: only the rows this chunk owns
      if (live && k >= band_lo && k < band_hi) out[k] = power_to_db(pw, kPrecise);
    });
    // pad columns of the row (band_hi - band_lo .. ld): +inf, which the radix select counts above every real value (select.cu)
    if (live)
      for (int cpad = band_hi + t; cpad < band_lo + ld; cpad += 8) out[cpad] = __int_as_float(0x7f800000);
    __syncwarp();
  }
  unsigned int m = __reduce_max_sync(0xffffffffu, __float_as_uint(pmax));  // non-negative floats order as uints
  if (lane == 0 && m != 0u) atomicMax(pmax_bits, m);
}

// The float64 variant with SIXTEEN threads per frame (stft_core16.cuh): 2 frames per warp, 8 warps per CTA, two CTAs per SM
// (<= 128 registers: a thread holds 16 complex values instead of 32).  Same tables except the stage-A twiddles (16 x 16).
constexpr int kFramesPerWarp16 = 2;
template <typename RealT>
constexpr size_t stft16_smem_bytes() {
  return (size_t)kTableCx * sizeof(Cx<RealT>) + (size_t)kWarpsPerCta * kFramesPerWarp16 * kFrameBuf16Cx * sizeof(Cx<RealT>);
}

template <typename SampleT, typename RealT>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
stft_db16_kernel(const SampleT* __restrict__ pcm, long long n_samples, long long T, float* __restrict__ raw,
                 int ld, int band_lo, int band_hi, const Cx<RealT>* __restrict__ tables,
                 unsigned int* __restrict__ pmax_bits, long long stat_row0, long long stat_row1) {
  constexpr bool kPrecise = false;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cx<RealT>* s_tab = reinterpret_cast<Cx<RealT>*>(smem_raw);
  Cx<RealT>* s_buf = s_tab + kTableCx;
  for (int i = threadIdx.x; i < kTableCx; i += blockDim.x) s_tab[i] = tables[i];
  __syncthreads();
  const StftTables<RealT> tb{s_tab, s_tab + 256, s_tab + 512};

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t = lane & 15;
  const int fl = lane >> 4;
  Cx<RealT>* fbuf = s_buf + (size_t)(warp * kFramesPerWarp16 + fl) * kFrameBuf16Cx;

  const long long n_groups = (T + kFramesPerWarp16 - 1) / kFramesPerWarp16;
  const long long g_stride = (long long)gridDim.x * kWarpsPerCta;
  float pmax = 0.0f;
  constexpr int BR[16] = {ORCAI_BITREV16_LIST};

  for (long long g = (long long)blockIdx.x * kWarpsPerCta + warp; g < n_groups; g += g_stride) {
    const long long j = g * kFramesPerWarp16 + fl;
    const long long base = (j - 1) * kHop;  // first sample of frame j (centre padding of n_fft/2)
    Cx<float> x[16];
    const long long j0 = g * kFramesPerWarp16;
    const bool interior = (j0 >= 1) && ((j0 + kFramesPerWarp16) * kHop <= n_samples);   // warp-uniform
    if (interior) {
      const SampleT* src = pcm + base + 2 * t;
#pragma unroll
      for (int p = 0; p < 16; ++p) x[p] = ld_pair(src + 32 * BR[p]);
    } else {
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const long long s0 = base + 2 * (16 * BR[p] + t);
        x[p].x = (s0 >= 0 && s0 < n_samples) ? ld_one(pcm + s0) : 0.0f;
        x[p].y = (s0 + 1 >= 0 && s0 + 1 < n_samples) ? ld_one(pcm + s0 + 1) : 0.0f;
      }
    }
    stage_a16<RealT>(x, t, tb, fbuf);
    __syncwarp();
    RealT fr[16], fi[16];
    stage_b16_fft<RealT>(t, fbuf, fr, fi);
    __syncwarp();                                   // every row has been read: the buffer takes the upper halves now
    stage_b16_pass_on<RealT>(t, fbuf, fr, fi);
    __syncwarp();
    const bool live = j < T;
    float* out = raw + (size_t)(live ? j : 0) * ld - band_lo;
    stage_b16_pairs<RealT>(t, tb, fbuf, fr, fi, [&](int k, RealT re, RealT im) {
      const float fre = (float)re, fim = (float)im;  // complex64 rounding of the reference's stft matrix
      const float pw = fmaf(fre, fre, fim * fim);
      pmax = fmaxf(pmax, (live && j >= stat_row0 && j < stat_row1) ? pw : 0.0f);
      if (live && k >= band_lo && k < band_hi) out[k] = power_to_db(pw, kPrecise);
    });
    if (live)
      for (int cpad = band_hi + t; cpad < band_lo + ld; cpad += 16) out[cpad] = __int_as_float(0x7f800000);   // pad columns: +inf (select.cu)
    __syncwarp();
  }
  unsigned int m = __reduce_max_sync(0xffffffffu, __float_as_uint(pmax));
  if (lane == 0 && m != 0u) atomicMax(pmax_bits, m);
}

template <typename SampleT, typename RealT>
int launch_variant16(Ctx* c, const void* d_pcm, int64_t n_samples, int64_t T, float* d_raw, const void* tab, int64_t stat_row0, int64_t stat_row1) {
  static std::atomic<unsigned long long> attr_devices{0ull};
  constexpr size_t smem = stft16_smem_bytes<RealT>();
  if (!((attr_devices.load() >> (c->device & 63)) & 1ull)) {
    ORCAI_CUDA(c, cudaFuncSetAttribute(stft_db16_kernel<SampleT, RealT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_devices.fetch_or(1ull << (c->device & 63));
  }
  const long long n_groups = (T + kFramesPerWarp16 - 1) / kFramesPerWarp16;
  long long ctas = (n_groups + kWarpsPerCta - 1) / kWarpsPerCta;
  const long long max_ctas = (long long)c->sm_count * 2;
  if (ctas > max_ctas) ctas = max_ctas;
  if (ctas < 1) ctas = 1;
  stft_db16_kernel<SampleT, RealT><<<(unsigned)ctas, kWarpsPerCta * 32, smem, c->stream>>>(
      static_cast<const SampleT*>(d_pcm), n_samples, T, d_raw, kRawLd, c->p.band_lo, c->p.band_hi,
      static_cast<const Cx<RealT>*>(tab), &c->d_sel->pmax_bits, (long long)stat_row0, (long long)stat_row1);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

template <typename SampleT, typename RealT>
int launch_variant(Ctx* c, const void* d_pcm, int64_t n_samples, int64_t T, float* d_raw, const void* tab, int64_t stat_row0, int64_t stat_row1) {
  // the attribute is per device: remember which devices have it (several contexts can live in one process)
  static std::atomic<unsigned long long> attr_devices{0ull};
  constexpr size_t smem = stft_smem_bytes<RealT>();
  if (!((attr_devices.load() >> (c->device & 63)) & 1ull)) {
    ORCAI_CUDA(c, cudaFuncSetAttribute(stft_db_kernel<SampleT, RealT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_devices.fetch_or(1ull << (c->device & 63));
  }
  const long long n_groups = (T + kFramesPerWarp - 1) / kFramesPerWarp;
  long long ctas = (n_groups + kWarpsPerCta - 1) / kWarpsPerCta;
  const long long max_ctas = (long long)c->sm_count * (sizeof(RealT) == 4 ? 2 : 1);
  if (ctas > max_ctas) ctas = max_ctas;
  if (ctas < 1) ctas = 1;
  stft_db_kernel<SampleT, RealT><<<(unsigned)ctas, kWarpsPerCta * 32, smem, c->stream>>>(
      static_cast<const SampleT*>(d_pcm), n_samples, T, d_raw, kRawLd, c->p.band_lo, c->p.band_hi,
      static_cast<const Cx<RealT>*>(tab), &c->d_sel->pmax_bits, (long long)stat_row0, (long long)stat_row1);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

}  // namespace

int stft_upload_tables(Ctx* c) {
  for (int which = 0; which < 2; ++which) {
    const double scale = which ? 0.5 / 32768.0 : 0.5;
    {
      const StftHostTables<float> t = make_stft_tables<float>(scale);
      std::vector<float> flat(t.win);
      flat.insert(flat.end(), t.tw.begin(), t.tw.end());
      flat.insert(flat.end(), t.ck.begin(), t.ck.end());
      ORCAI_CUDA(c, cudaMalloc(&c->d_tables[which], flat.size() * sizeof(float)));
      ORCAI_CUDA(c, cudaMemcpy(c->d_tables[which], flat.data(), flat.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    {
      const StftHostTables<double> t = make_stft_tables<double>(scale);
      std::vector<double> flat(t.win);
      flat.insert(flat.end(), t.tw.begin(), t.tw.end());
      flat.insert(flat.end(), t.ck.begin(), t.ck.end());
      ORCAI_CUDA(c, cudaMalloc(&c->d_tables64[which], flat.size() * sizeof(double)));
      ORCAI_CUDA(c, cudaMemcpy(c->d_tables64[which], flat.data(), flat.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    {
      const StftHostTables<double> t = make_stft_tables<double>(scale, 16);   // 16 threads per frame (stft_core16.cuh)
      std::vector<double> flat(t.win);
      flat.insert(flat.end(), t.tw.begin(), t.tw.end());
      flat.insert(flat.end(), t.ck.begin(), t.ck.end());
      ORCAI_CUDA(c, cudaMalloc(&c->d_tables64_16[which], flat.size() * sizeof(double)));
      ORCAI_CUDA(c, cudaMemcpy(c->d_tables64_16[which], flat.data(), flat.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
  }
  return ORCAI_OK;
}

int launch_stft(Ctx* c, const void* d_pcm, int dtype, int64_t n_samples, int64_t T, float* d_raw, int64_t stat_row0, int64_t stat_row1) {
  const int which = (dtype == ORCAI_PCM_I16) ? 1 : 0;
  // pmax and the flag telling the consumers which logarithm K1 used
  ORCAI_CUDA(c, cudaMemsetAsync(&c->d_sel->pmax_bits, 0, sizeof(unsigned int), c->stream));
  const int precise = 0;
  ORCAI_CUDA(c, cudaMemcpyAsync(&c->d_sel->precise_log, &c->h_flags[precise], sizeof(int), cudaMemcpyHostToDevice, c->stream));
  if (c->stft_f64 && c->stft_threads == 16) {
    if (which) return launch_variant16<int16_t, double>(c, d_pcm, n_samples, T, d_raw, c->d_tables64_16[1], stat_row0, stat_row1);
    return launch_variant16<float, double>(c, d_pcm, n_samples, T, d_raw, c->d_tables64_16[0], stat_row0, stat_row1);
  }
  if (c->stft_f64) {
    if (which) return launch_variant<int16_t, double>(c, d_pcm, n_samples, T, d_raw, c->d_tables64[1], stat_row0, stat_row1);
    return launch_variant<float, double>(c, d_pcm, n_samples, T, d_raw, c->d_tables64[0], stat_row0, stat_row1);
  }
  if (which) return launch_variant<int16_t, float>(c, d_pcm, n_samples, T, d_raw, c->d_tables[1], stat_row0, stat_row1);
  return launch_variant<float, float>(c, d_pcm, n_samples, T, d_raw, c->d_tables[0], stat_row0, stat_row1);
}

}  // namespace orcai
