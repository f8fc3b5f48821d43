// K1: fused window + 512-point real FFT + |.|^2 + 10*log10 + band crop + global power max.
//
// Replaces librosa.stft / amplitude_to_db at src/orcAI/spectrogram.py:34-39 and :51-53 of the
// reference (the "ref = np.max" shift and the top_db floor are applied by the consumers from the
// exact global maximum this kernel produces, so no second pass over the 257-bin array exists).
//
// Mapping: 8 threads per frame, 4 frames per warp, 8 warps per CTA, 2 CTAs per SM; each warp walks
// groups of 4 consecutive frames (the 50% overlap makes the second read of every sample an L1 hit).
// All FFT butterflies are register-resident with immediate twiddles (fft_gen.cuh); the only exchange
// is one warp-private shared-memory transpose (stft_core.cuh).  HBM traffic per frame: 256 new
// samples in, band_hi-band_lo floats out.
#include "common.h"
#include "stft_core.cuh"
#include "stft_tables.h"

namespace orcai {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kFramesPerWarp = 4;
constexpr int kTableFloat2 = 768;
constexpr size_t kStftSmem = (size_t)kTableFloat2 * sizeof(float2) +
                             (size_t)kWarpsPerCta * kFramesPerWarp * kFrameBufFloat2 * sizeof(float2);

__device__ __forceinline__ float2 ld_pair(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 ld_pair(const int16_t* p) {
  // two PCM16 samples -> exact float integers via the 1.5*2^23 magic constant (no I2F on the slow pipe)
  const unsigned int u = __ldg(reinterpret_cast<const unsigned int*>(p));
  const int lo = (int)(short)(u & 0xffffu);
  const int hi = ((int)u) >> 16;
  float2 r;
  r.x = __int_as_float(0x4B400000 + lo) - 12582912.0f;
  r.y = __int_as_float(0x4B400000 + hi) - 12582912.0f;
  return r;
}
__device__ __forceinline__ float ld_one(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_one(const int16_t* p) { return (float)__ldg(p); }

template <typename SampleT>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
stft_db_kernel(const SampleT* __restrict__ pcm, long long n_samples, long long T, float* __restrict__ raw,
               int ld, int band_lo, int band_hi, const float2* __restrict__ tables,
               unsigned int* __restrict__ pmax_bits) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* s_tab = reinterpret_cast<float2*>(smem_raw);
  float2* s_buf = s_tab + kTableFloat2;
  for (int i = threadIdx.x; i < kTableFloat2; i += blockDim.x) s_tab[i] = tables[i];
  __syncthreads();
  const StftTables tb{s_tab, s_tab + 256, s_tab + 512};

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t = lane & 7;
  const int fl = lane >> 3;
  float2* fbuf = s_buf + (size_t)(warp * kFramesPerWarp + fl) * kFrameBufFloat2;

  const long long n_groups = (T + kFramesPerWarp - 1) / kFramesPerWarp;
  const long long g_stride = (long long)gridDim.x * kWarpsPerCta;
  float pmax = 0.0f;
  constexpr int BR[32] = {ORCAI_BITREV32_LIST};

  for (long long g = (long long)blockIdx.x * kWarpsPerCta + warp; g < n_groups; g += g_stride) {
    const long long j = g * kFramesPerWarp + fl;
    const long long base = (j - 1) * kHop;  // first sample of frame j (centre padding of n_fft/2)
    float2 x[32];
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
    stage_a(x, t, tb, fbuf);
    __syncwarp();
    const bool live = j < T;
    float* out = raw + (size_t)(live ? j : 0) * ld - band_lo;
    stage_b(t, tb, fbuf, [&](int k, float pw) {
      pmax = fmaxf(pmax, live ? pw : 0.0f);
      if (live && k >= band_lo && k < band_hi) out[k] = kTenLog10Of2 * __log2f(fmaxf(pw, kAminPower));
    });
    __syncwarp();
  }
  unsigned int m = __reduce_max_sync(0xffffffffu, __float_as_uint(pmax));  // non-negative floats order as uints
  if (lane == 0 && m != 0u) atomicMax(pmax_bits, m);
}

}  // namespace

int launch_stft(Ctx* c, const void* d_pcm, int dtype, int64_t n_samples, int64_t T, float* d_raw) {
  static bool attr_set[2] = {false, false};
  const int which = (dtype == ORCAI_PCM_I16) ? 1 : 0;
  if (!attr_set[which]) {
    if (which)
      ORCAI_CUDA(c, cudaFuncSetAttribute(stft_db_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStftSmem));
    else
      ORCAI_CUDA(c, cudaFuncSetAttribute(stft_db_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStftSmem));
    attr_set[which] = true;
  }
  ORCAI_CUDA(c, cudaMemsetAsync(&c->d_sel->pmax_bits, 0, sizeof(unsigned int), c->stream));
  const long long n_groups = (T + kFramesPerWarp - 1) / kFramesPerWarp;
  long long ctas = (n_groups + kWarpsPerCta - 1) / kWarpsPerCta;
  const long long max_ctas = (long long)c->sm_count * 2;
  if (ctas > max_ctas) ctas = max_ctas;
  if (ctas < 1) ctas = 1;
  const float2* tab = reinterpret_cast<const float2*>(c->d_tables[which]);
  if (which)
    stft_db_kernel<int16_t><<<(unsigned)ctas, kWarpsPerCta * 32, kStftSmem, c->stream>>>(
        static_cast<const int16_t*>(d_pcm), n_samples, T, d_raw, kRawLd, c->p.band_lo, c->p.band_hi, tab, &c->d_sel->pmax_bits);
  else
    stft_db_kernel<float><<<(unsigned)ctas, kWarpsPerCta * 32, kStftSmem, c->stream>>>(
        static_cast<const float*>(d_pcm), n_samples, T, d_raw, kRawLd, c->p.band_lo, c->p.band_hi, tab, &c->d_sel->pmax_bits);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

}  // namespace orcai
