// Exact nearest-rank percentiles + K2 normalise.
//
// Replaces np.percentile(..., method="nearest") x2, np.clip and the min-max normalisation of
// preprocess_spectrogram (src/orcAI/spectrogram.py:70-86 of the reference), plus the
// "ref=np.max" shift and top_db floor of amplitude_to_db (spectrogram.py:51-53) which are applied
// here on the fly:  v = max(L - L_ref, -80).
//
// The two order statistics are located exactly by a 3-pass (11+11+10 bit) radix select over the
// order-preserving integer image of the float32 values v; both ranks are resolved in the same
// passes.  Passes 2 and 3 only count elements that match the already-decided prefix, so they are
// plain streaming reads.
#include <cstring>
#include "common.h"

namespace orcai {

namespace {

constexpr float kTopDb = 80.0f;

__device__ __forceinline__ unsigned int f2key(float v) {
  const unsigned int u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned int k) {
  const unsigned int u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}
__device__ __forceinline__ float shifted_db(float raw, float db_ref) { return fmaxf(raw - db_ref, -kTopDb); }

__global__ void select_init_kernel(SelectState* st, unsigned long long rank_lo, unsigned long long rank_hi) {
  // db_ref through the same expression K1 uses, so the loudest cell is exactly 0 dB
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const float pmax = __uint_as_float(st->pmax_bits);
    st->db_ref = power_to_db(pmax, st->precise_log != 0);
    st->rank[0] = rank_lo;
    st->rank[1] = rank_hi;
    st->prefix[0] = 0u;
    st->prefix[1] = 0u;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * 2048; i += gridDim.x * blockDim.x)
    (&st->hist[0][0])[i] = 0ull;
}

// PASS 0: digit = key >> 21 (11 bits), one shared histogram (stored in hist[0]).
// PASS 1: digit = (key >> 10) & 2047 for keys whose top 11 bits equal prefix[r] >> 21.
// PASS 2: digit = key & 1023 for keys whose top 22 bits equal prefix[r] >> 10.
// The raw dB buffer is streamed as float4 (row pitch kRawLd = 44 float4; the pad columns are masked), four independent
// loads in flight per thread.  Pass 0 sees very few distinct digits (the top 11 key bits are sign + exponent + 2 mantissa
// bits of a value in [-80, 0]), so its shared-memory atomics are aggregated per warp with match.any first.
template <int PASS>
__device__ __forceinline__ void select_count(unsigned int (*sh)[2048], float raw, bool valid, float db_ref, unsigned int p0, unsigned int p1) {
  const unsigned int key = f2key(shifted_db(raw, db_ref));
  if (PASS == 0) {
    const unsigned int active = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const unsigned int d = key >> 21;
      const unsigned int peers = __match_any_sync(active, d);
      if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sh[0][d], (unsigned int)__popc(peers));
    }
  } else if (PASS == 1) {
    if (valid && (key >> 21) == (p0 >> 21)) atomicAdd(&sh[0][(key >> 10) & 2047u], 1u);
    if (valid && (key >> 21) == (p1 >> 21)) atomicAdd(&sh[1][(key >> 10) & 2047u], 1u);
  } else {
    if (valid && (key >> 10) == (p0 >> 10)) atomicAdd(&sh[0][key & 1023u], 1u);
    if (valid && (key >> 10) == (p1 >> 10)) atomicAdd(&sh[1][key & 1023u], 1u);
  }
}

template <int PASS>
__global__ void __launch_bounds__(256)
select_hist_kernel(const float* __restrict__ raw, long long T, int ld, int nb, SelectState* st) {
  __shared__ unsigned int sh[2][2048];
  for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) (&sh[0][0])[i] = 0u;
  __syncthreads();
  const float db_ref = st->db_ref;
  const unsigned int p0 = st->prefix[0], p1 = st->prefix[1];
  const int ld4 = ld >> 2;
  const long long n4 = T * ld4;
  const float4* raw4 = reinterpret_cast<const float4*>(raw);
  constexpr int ILP = 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // whole warps stay in the loop together (match.any / ballot need converged warps): iterate to a warp-uniform bound
  for (long long base = (long long)blockIdx.x * blockDim.x; base < n4; base += stride * ILP) {
    float4 v[ILP];
    long long e[ILP];
#pragma unroll
    for (int q = 0; q < ILP; ++q) {
      e[q] = base + q * stride + threadIdx.x;
      v[q] = e[q] < n4 ? __ldg(raw4 + e[q]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int q = 0; q < ILP; ++q) {
      const int c = (int)(e[q] % ld4) * 4;
      const bool in = e[q] < n4;
      select_count<PASS>(sh, v[q].x, in && c < nb, db_ref, p0, p1);
      select_count<PASS>(sh, v[q].y, in && c + 1 < nb, db_ref, p0, p1);
      select_count<PASS>(sh, v[q].z, in && c + 2 < nb, db_ref, p0, p1);
      select_count<PASS>(sh, v[q].w, in && c + 3 < nb, db_ref, p0, p1);
    }
  }
  __syncthreads();
  const int nh = (PASS == 0) ? 1 : 2;
  for (int i = threadIdx.x; i < nh * 2048; i += blockDim.x) {
    const unsigned int v = (&sh[0][0])[i];
    if (v) atomicAdd(&(&st->hist[0][0])[i], (unsigned long long)v);
  }
}

// One CTA: locate the digit holding each rank, extend the prefixes, clear the histograms.
template <int PASS>
__global__ void __launch_bounds__(1024) select_scan_kernel(SelectState* st) {
  constexpr int NB = (PASS == 2) ? 1024 : 2048;
  constexpr int SHIFT = (PASS == 0) ? 21 : (PASS == 1 ? 10 : 0);
  __shared__ unsigned long long pre[2048];
  __shared__ unsigned long long wsum[32];
  for (int r = 0; r < 2; ++r) {
    const unsigned long long* h = (PASS == 0) ? st->hist[0] : st->hist[r];
    // inclusive scan of NB counters by 1024 threads (2 items per thread when NB = 2048)
    constexpr int PER = NB / 1024;
    unsigned long long v[PER], tot = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) { v[q] = h[threadIdx.x * PER + q]; tot += v[q]; }
    unsigned long long inc = tot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long n = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += n;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    unsigned long long run = inc - tot + (warp ? wsum[warp - 1] : 0ull);  // exclusive prefix of this thread
#pragma unroll
    for (int q = 0; q < PER; ++q) { pre[threadIdx.x * PER + q] = run; run += v[q]; }
    __syncthreads();
    const unsigned long long rank = st->rank[r];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int d = threadIdx.x * PER + q;
      const unsigned long long lo = pre[d];
      const unsigned long long cnt = h[d];
      if (cnt && rank >= lo && rank < lo + cnt) {
        st->rank[r] = rank - lo;
        st->prefix[r] |= ((unsigned int)d) << SHIFT;
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) (&st->hist[0][0])[i] = 0ull;
  if (PASS == 2) {
    __syncthreads();
    if (threadIdx.x == 0) {
      st->lo = key2f(st->prefix[0]);
      st->hi = key2f(st->prefix[1]);
    }
  }
}

// K2: out[j][b] = (clip(v, lo, hi) - lo) / (hi - lo), compact (T, nb) float32.
__global__ void __launch_bounds__(256)
normalise_kernel(const float* __restrict__ raw, long long T, int ld, int nb, const SelectState* __restrict__ st,
                 float* __restrict__ out, int mode) {
  const float db_ref = st->db_ref, lo = st->lo, hi = st->hi;
  const float range = hi - lo;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long j = (long long)blockIdx.x * (blockDim.x >> 5) + warp; j < T; j += wstride) {
    const float* row = raw + (size_t)j * ld;
    float* o = out + (size_t)j * nb;
    for (int b = lane; b < nb; b += 32) {
      const float v = shifted_db(row[b], db_ref);
      o[b] = (mode == 0) ? __fdiv_rn(fminf(fmaxf(v, lo), hi) - lo, range) : v;
    }
  }
}

}  // namespace

int launch_select(Ctx* c, const float* d_raw, int64_t T) {
  const int nb = c->p.band_hi - c->p.band_lo;
  const unsigned long long n = (unsigned long long)T * (unsigned long long)nb;
  // np.percentile 'nearest': index = around((n - 1) * q), half to even
  const unsigned long long r0 = (unsigned long long)nearbyint((double)(n - 1) * c->p.q_lo);
  const unsigned long long r1 = (unsigned long long)nearbyint((double)(n - 1) * c->p.q_hi);
  const int grid = c->sm_count * 8;
  select_init_kernel<<<8, 512, 0, c->stream>>>(c->d_sel, r0, r1);
  select_hist_kernel<0><<<grid, 256, 0, c->stream>>>(d_raw, T, kRawLd, nb, c->d_sel);
  select_scan_kernel<0><<<1, 1024, 0, c->stream>>>(c->d_sel);
  select_hist_kernel<1><<<grid, 256, 0, c->stream>>>(d_raw, T, kRawLd, nb, c->d_sel);
  select_scan_kernel<1><<<1, 1024, 0, c->stream>>>(c->d_sel);
  select_hist_kernel<2><<<grid, 256, 0, c->stream>>>(d_raw, T, kRawLd, nb, c->d_sel);
  select_scan_kernel<2><<<1, 1024, 0, c->stream>>>(c->d_sel);
  c->launches += 7;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

namespace {
__global__ void select_set_prefix_kernel(SelectState* st, unsigned int p0, unsigned int p1) {
  st->prefix[0] = p0;
  st->prefix[1] = p1;
}
__global__ void select_set_keys_kernel(SelectState* st, unsigned int k0, unsigned int k1) {
  st->prefix[0] = k0;
  st->prefix[1] = k1;
  st->lo = key2f(k0);
  st->hi = key2f(k1);
}
__global__ void select_set_pmax_kernel(SelectState* st, unsigned int bits) { st->pmax_bits = bits; }
}  // namespace

int launch_select_begin(Ctx* c, float max_power) {
  unsigned int bits;
  memcpy(&bits, &max_power, 4);
  select_set_pmax_kernel<<<1, 1, 0, c->stream>>>(c->d_sel, bits);
  select_init_kernel<<<8, 512, 0, c->stream>>>(c->d_sel, 0ull, 0ull);   // db_ref from the global maximum; clears the histograms
  c->launches += 2;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

int launch_select_histogram(Ctx* c, const float* d_raw_rows, int64_t n_rows, int pass, const uint32_t prefix[2], uint64_t* h_hist) {
  const int nb = c->p.band_hi - c->p.band_lo;
  const int grid = c->sm_count * 8;
  select_set_prefix_kernel<<<1, 1, 0, c->stream>>>(c->d_sel, prefix[0], prefix[1]);
  if (n_rows > 0) {
    if (pass == 0) select_hist_kernel<0><<<grid, 256, 0, c->stream>>>(d_raw_rows, n_rows, kRawLd, nb, c->d_sel);
    else if (pass == 1) select_hist_kernel<1><<<grid, 256, 0, c->stream>>>(d_raw_rows, n_rows, kRawLd, nb, c->d_sel);
    else select_hist_kernel<2><<<grid, 256, 0, c->stream>>>(d_raw_rows, n_rows, kRawLd, nb, c->d_sel);
  }
  c->launches += 2;
  ORCAI_CUDA(c, cudaGetLastError());
  ORCAI_CUDA(c, cudaMemcpyAsync(h_hist, &c->d_sel->hist[0][0], sizeof(unsigned long long) * 2 * 2048, cudaMemcpyDeviceToHost, c->stream));
  ORCAI_CUDA(c, cudaMemsetAsync(&c->d_sel->hist[0][0], 0, sizeof(unsigned long long) * 2 * 2048, c->stream));
  ORCAI_CUDA(c, cudaStreamSynchronize(c->stream));
  return ORCAI_OK;
}

int launch_select_end(Ctx* c, const uint32_t key[2]) {
  select_set_keys_kernel<<<1, 1, 0, c->stream>>>(c->d_sel, key[0], key[1]);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

int launch_normalise(Ctx* c, const float* d_raw, int64_t T, float* d_spec) {
  const int nb = c->p.band_hi - c->p.band_lo;
  normalise_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(d_raw, T, kRawLd, nb, c->d_sel, d_spec, 0);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

int launch_read_db(Ctx* c, const float* d_raw, int64_t T, float* d_out) {
  const int nb = c->p.band_hi - c->p.band_lo;
  normalise_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(d_raw, T, kRawLd, nb, c->d_sel, d_out, 1);
  c->launches++;
  ORCAI_CUDA(c, cudaGetLastError());
  return ORCAI_OK;
}

}  // namespace orcai
