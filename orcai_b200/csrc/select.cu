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
  // negative: ~u, non-negative: u | sign bit  ==  u ^ (sign-extended sign | sign bit): one shift + one three-input LOP
  const unsigned int u = __float_as_uint(v);
  return u ^ ((unsigned int)((int)u >> 31) | 0x80000000u);
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
    st->tile_counter = 0u;   // the passes' last-CTA ticket: a select always starts from zero, whatever an earlier failed call left
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * 2048; i += gridDim.x * blockDim.x)
    (&st->hist[0][0])[i] = 0ull;
}

// Locate the digit holding each rank, extend the prefixes, clear the histograms: run by the LAST CTA of a histogram pass
// to finish (ticket in SelectState::tile_counter), so that a pass is one launch.  The counters were accumulated with L2
// atomics by every CTA of the grid and are read past L1 (__ldcg).
template <int PASS, int NT>
__device__ __forceinline__ void select_scan_body(SelectState* st, unsigned long long* wsum) {
  constexpr int NB = (PASS == 2) ? 1024 : 2048;
  constexpr int SHIFT = (PASS == 0) ? 21 : (PASS == 1 ? 10 : 0);
  constexpr int PER = NB / NT;
  static_assert(NB % NT == 0 && NT % 32 == 0 && NT / 32 <= 32, "scan layout");
  // ranks that entered the pass with one prefix were counted into histogram 0 only (select_hist_kernel)
  constexpr int SHP = (PASS == 1) ? 21 : 10;
  const bool shared_hist = (PASS == 0) || ((st->prefix[1] >> SHP) == (st->prefix[0] >> SHP));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  for (int r = 0; r < 2; ++r) {
    const unsigned long long* h = shared_hist ? st->hist[0] : st->hist[r];
    unsigned long long v[PER], tot = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) { v[q] = __ldcg(h + threadIdx.x * PER + q); tot += v[q]; }
    unsigned long long inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = lane < NT / 32 ? wsum[lane] : 0ull;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long n = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += n;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    unsigned long long run = inc - tot + (warp ? wsum[warp - 1] : 0ull);  // exclusive prefix of this thread
    const unsigned long long rank = st->rank[r];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int d = threadIdx.x * PER + q;
      if (v[q] && rank >= run && rank < run + v[q]) {
        st->rank[r] = rank - run;
        st->prefix[r] |= ((unsigned int)d) << SHIFT;
      }
      run += v[q];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 2 * 2048; i += NT) (&st->hist[0][0])[i] = 0ull;
  if (PASS == 2) {
    __syncthreads();
    if (threadIdx.x == 0) {
      st->lo = key2f(st->prefix[0]);
      st->hi = key2f(st->prefix[1]);
    }
  }
}

// PASS 0: digit = key >> 21 (11 bits), one shared histogram (stored in hist[0]).
// PASS 1: digit = (key >> 10) & 2047 for keys whose top 11 bits equal prefix[r] >> 21.
// PASS 2: digit = key & 1023 for keys whose top 22 bits equal prefix[r] >> 10.
// The raw dB buffer is streamed as whole rows of float4 (row pitch kRawLd = 44 float4).  K1 fills the pad columns of every
// row with +inf, whose key sorts above every real value: the passes count them like any other element (they land in digits
// that no rank can reach, both ranks are counted from the bottom) and need no column arithmetic or masks at all.  Every
// thread keeps two batches of four 16-byte loads in flight (the next batch is requested before the current one is counted);
// when both ranks share their prefix only one histogram is filled (the scan reads it for both).  The passes were issue-bound
// before (132 warp instructions per float4 - a 64-bit modulo for the column, four masks, branchy key - 77 % of the issue
// slots, 2.1-3.3 TB/s); profiles/README.md has the before / after rows.
// Pass 0 sees very few distinct digits (the top 11 key bits are sign + exponent + 2 mantissa bits of a value in [-80, 0]),
// so plain shared-memory atomics would serialise on a handful of addresses and match.any aggregation is bound by the XU
// pipe (188 us): the digits of the value range are counted in a table with one counter per (digit, lane) instead.
constexpr int kSelThreads = 512;
constexpr unsigned int kHotLo = 0x3D5FFFFFu >> 21;   // digit of the -80 dB floor: f2key(-80.0f) = ~0xC2A00000
constexpr int kHotDigits = 64;                       // 16 binades x 4: down to |v| = 80 / 2^15 dB
constexpr int kSelIlp = 4;
#define ORCAI_PAD_POISON __int_as_float(0x7f800000)

template <int PASS>
__global__ void __launch_bounds__(kSelThreads, 2)
select_hist_kernel(const float* __restrict__ raw, long long T, int ld, SelectState* st, int scan) {
  __shared__ unsigned int sh[2][2048];
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned int s_ticket;
  __shared__ unsigned int hot[PASS == 0 ? kHotDigits : 1][32];
  for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) (&sh[0][0])[i] = 0u;
  if (PASS == 0)
    for (int i = threadIdx.x; i < kHotDigits * 32; i += blockDim.x) (&hot[0][0])[i] = 0u;
  __syncthreads();
  const float db_ref = st->db_ref;
  constexpr int SH = (PASS == 1) ? 21 : 10;             // bits below the decided prefix (passes 1, 2)
  const unsigned int p0h = st->prefix[0] >> SH;
  const unsigned int p1h = (st->prefix[1] >> SH) == p0h ? 0xffffffffu : (st->prefix[1] >> SH);  // shared prefix: histogram 0 serves both
  const long long n4 = T * (ld >> 2);
  const float4* raw4 = reinterpret_cast<const float4*>(raw);
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float4 poison = make_float4(ORCAI_PAD_POISON, ORCAI_PAD_POISON, ORCAI_PAD_POISON, ORCAI_PAD_POISON);

  auto count4 = [&](const float4& v) {
    const unsigned int k0 = f2key(shifted_db(v.x, db_ref)), k1 = f2key(shifted_db(v.y, db_ref));
    const unsigned int k2 = f2key(shifted_db(v.z, db_ref)), k3 = f2key(shifted_db(v.w, db_ref));
    if (PASS == 0) {
      // digits of [-80 dB, -0.002 dB] (kHotDigits from the floor's digit upwards): one counter per (digit, lane), i.e. every
      // lane of a warp owns a bank - no conflicts, no aggregation; anything else (the pad columns' +inf, the loudest cell's
      // 0 dB) goes to the plain histogram, at most a lane or two per warp instruction
      auto one = [&](unsigned int key) {
        const unsigned int d = key >> 21, r = d - kHotLo;
        unsigned int* bin = r < (unsigned int)kHotDigits ? &hot[r][lane] : &sh[0][d];
        atomicAdd(bin, 1u);
      };
      one(k0); one(k1); one(k2); one(k3);
    } else {
      constexpr unsigned int S2 = (PASS == 1) ? 10u : 0u;
      constexpr unsigned int M = (PASS == 1) ? 2047u : 1023u;
      // one branch per element around both tests: the bin address is only formed for elements inside a wanted prefix
      auto one = [&](unsigned int key) {
        const unsigned int h = key >> SH;
        if (h == p0h || h == p1h) {
          unsigned int* bin = &sh[h == p0h ? 0 : 1][(key >> S2) & M];
          atomicAdd(bin, 1u);
        }
      };
      one(k0); one(k1); one(k2); one(k3);
    }
  };

  float4 cur[kSelIlp], nxt[kSelIlp];
#pragma unroll
  for (int q = 0; q < kSelIlp; ++q) cur[q] = e + q * stride < n4 ? __ldg(raw4 + e + q * stride) : poison;
  // whole warps stay in the loop together (the pass-0 votes need converged warps): the bound is the warp's first element
  while (e - lane < n4) {
    const long long en = e + kSelIlp * stride;
#pragma unroll
    for (int q = 0; q < kSelIlp; ++q) nxt[q] = en + q * stride < n4 ? __ldg(raw4 + en + q * stride) : poison;
#pragma unroll
    for (int q = 0; q < kSelIlp; ++q) {
      count4(cur[q]);
      cur[q] = nxt[q];
    }
    e = en;
  }
  __syncthreads();
  if (PASS == 0 && threadIdx.x < kHotDigits) {
    unsigned int tot = 0u;
#pragma unroll
    for (int l = 0; l < 32; ++l) tot += hot[threadIdx.x][(l + threadIdx.x) & 31];
    sh[0][kHotLo + threadIdx.x] += tot;
  }
  __syncthreads();
  const int nh = (PASS == 0) ? 1 : 2;
  for (int i = threadIdx.x; i < nh * 2048; i += blockDim.x) {
    const unsigned int v = (&sh[0][0])[i];
    if (v) atomicAdd(&(&st->hist[0][0])[i], (unsigned long long)v);
  }
  if (!scan) return;                                    // time chunks: the host sums the chunks' histograms and scans
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&st->tile_counter, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  select_scan_body<PASS, kSelThreads>(st, wsum);
  if (threadIdx.x == 0) st->tile_counter = 0u;
}

// K2: out[j][b] = (clip(v, lo, hi) - lo) / (hi - lo), compact (T, nb) float32.
__global__ void __launch_bounds__(256)
normalise_kernel(const float* __restrict__ raw, long long T, int ld, int nb, const SelectState* __restrict__ st,
                 float* __restrict__ out, int mode) {
  const float db_ref = st->db_ref, lo = st->lo, hi = st->hi;
  const float range = hi - lo;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  // a row is at most kRawLd = 176 floats: six coalesced loads per lane, all in flight before the first store (the loop
  // over a run-time column count kept ONE load in flight per thread: 3.4 TB/s)
  constexpr int Q = (kRawLd + 31) / 32;
  constexpr int R = 4;   // rows per warp iteration: twenty-four loads per lane in flight
  for (long long j = (long long)blockIdx.x * (blockDim.x >> 5) + warp; j < T; j += R * wstride) {
    float x[R][Q];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long jr = j + r * wstride;
      const float* row = raw + (size_t)(jr < T ? jr : j) * ld;
#pragma unroll
      for (int q = 0; q < Q; ++q) x[r][q] = lane + 32 * q < nb ? __ldg(row + lane + 32 * q) : 0.0f;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long jr = j + r * wstride;
      if (jr >= T) break;
      float* o = out + (size_t)jr * nb;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const float v = shifted_db(x[r][q], db_ref);
        if (lane + 32 * q < nb) o[lane + 32 * q] = (mode == 0) ? __fdiv_rn(fminf(fmaxf(v, lo), hi) - lo, range) : v;
      }
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
  const int grid = c->sm_count * 2;
  select_init_kernel<<<8, 512, 0, c->stream>>>(c->d_sel, r0, r1);
  select_hist_kernel<0><<<grid, kSelThreads, 0, c->stream>>>(d_raw, T, kRawLd, c->d_sel, 1);
  select_hist_kernel<1><<<grid, kSelThreads, 0, c->stream>>>(d_raw, T, kRawLd, c->d_sel, 1);
  select_hist_kernel<2><<<grid, kSelThreads, 0, c->stream>>>(d_raw, T, kRawLd, c->d_sel, 1);
  c->launches += 4;
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
  const int grid = c->sm_count * 2;
  select_set_prefix_kernel<<<1, 1, 0, c->stream>>>(c->d_sel, prefix[0], prefix[1]);
  if (n_rows > 0) {
    if (pass == 0) select_hist_kernel<0><<<grid, kSelThreads, 0, c->stream>>>(d_raw_rows, n_rows, kRawLd, c->d_sel, 0);
    else if (pass == 1) select_hist_kernel<1><<<grid, kSelThreads, 0, c->stream>>>(d_raw_rows, n_rows, kRawLd, c->d_sel, 0);
    else select_hist_kernel<2><<<grid, kSelThreads, 0, c->stream>>>(d_raw_rows, n_rows, kRawLd, c->d_sel, 0);
  }
  c->launches += 2;
  ORCAI_CUDA(c, cudaGetLastError());
  ORCAI_CUDA(c, cudaMemcpyAsync(h_hist, &c->d_sel->hist[0][0], sizeof(unsigned long long) * 2 * 2048, cudaMemcpyDeviceToHost, c->stream));
  ORCAI_CUDA(c, cudaMemsetAsync(&c->d_sel->hist[0][0], 0, sizeof(unsigned long long) * 2 * 2048, c->stream));
  ORCAI_CUDA(c, cudaStreamSynchronize(c->stream));
  // ranks that share their prefix were counted into histogram 0 only: hand the caller the same counters for both
  const int sh = (pass == 1) ? 21 : 10;
  if (pass > 0 && (prefix[0] >> sh) == (prefix[1] >> sh)) memcpy(h_hist + 2048, h_hist, sizeof(uint64_t) * 2048);
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
