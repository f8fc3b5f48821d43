// Per-thread building blocks of the fused STFT kernel (K1).  Host+device so that the exact
// index arithmetic can be replayed on the CPU (tests/host_emul/stft_emul.cpp).  Templated on the
// real type: float (fast variant) and double (parity-grade variant: the reference's numpy.fft.rfft
// runs in float64 and is then rounded to complex64).
//
// Replaces the reference's librosa.stft + amplitude_to_db call sites
// (src/orcAI/spectrogram.py:34-39, 51-53).
//
// One 512-sample frame is handled by 8 threads (t = 0..7):
//   z[m] = y[2m] + i*y[2m+1], m = 8*n1 + t             (256-point complex FFT of the packed real frame)
//   stage A : thread t = n2 runs a 32-point FFT over n1, applies W256^(n2*k1), writes row k1 of the
//             frame's 32x8 exchange buffer in shared memory (XOR-swizzled 2-element chunks)
//   stage B : thread t reads rows k1 in {t, 32-t, 8+t, 24-t} (thread 0: {0,16,8,24}), runs four 8-point
//             FFTs over n2 -> Z[k1 + 32*k2]; both members of every (k, 256-k) pair live in one thread
//   pairs   : X[k] = E - G, X[256-k] = conj(E + G) with E = Z[k] + conj Z[256-k], G = i*c_k*(Z[k] - conj Z[256-k])
//             (the 1/2 is folded into the window table); sink(k, re, im) receives every bin once.
#pragma once
#include "fft_gen.cuh"

#ifdef __CUDACC__
#define ORCAI_DEV_INLINE __host__ __device__ __forceinline__
#else
#define ORCAI_DEV_INLINE inline
#endif

namespace orcai {

constexpr int kNfft = 512;
constexpr int kHop = 256;
constexpr int kBins = 257;
constexpr int kFrameBufCx = 256 + 8;  // 32 rows x 8 + one row of padding: adjacent frames land in opposite bank halves
constexpr float kAminPower = 1e-10f;
constexpr float kTenLog10Of2 = 3.01029995663981195f;  // 10*log10(2)

template <typename T>
struct alignas(2 * sizeof(T)) Cx {
  T x, y;
};

// tables:  win[256] (Hann pairs * scale), tw[256] (W256^(n2*k1) at k1*8+n2), ck[256] (e^{-2 pi i k/512})
template <typename T>
struct StftTables {
  const Cx<T>* win;
  const Cx<T>* tw;
  const Cx<T>* ck;
};

ORCAI_DEV_INLINE int swz_elem(int k1, int n2) {
  // element slot (0..7) of (row k1, column n2) inside the row
  return ((((n2 >> 1) ^ ((k1 >> 1) & 3)) << 1) | (n2 & 1));
}

ORCAI_DEV_INLINE int stageb_row(int t, int s) {
  // rows owned by thread t in stage B
  switch (s) {
    case 0: return t;
    case 1: return t ? 32 - t : 16;
    case 2: return 8 + t;
    default: return 24 - t;
  }
}

// Stage A.  x[p] holds the raw float32 sample pair for n1 = BITREV32[p] (already fetched by the caller).
template <typename T>
ORCAI_DEV_INLINE void stage_a(const Cx<float> (&x)[32], int t, const StftTables<T>& tb, Cx<T>* fbuf) {
  constexpr int BR[32] = {ORCAI_BITREV32_LIST};
  T zr[32], zi[32];
#pragma unroll
  for (int p = 0; p < 32; ++p) {
    const Cx<T> w = tb.win[8 * BR[p] + t];
    zr[p] = T(x[p].x) * w.x;
    zi[p] = T(x[p].y) * w.y;
  }
  orcai_fft32_dit<T>(zr, zi);
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) {
    const Cx<T> w = tb.tw[8 * k1 + t];
    Cx<T> v;
    v.x = zr[k1] * w.x - zi[k1] * w.y;
    v.y = zr[k1] * w.y + zi[k1] * w.x;
    fbuf[8 * k1 + swz_elem(k1, t)] = v;
  }
}

// Stage B + pair post-processing.  Calls sink(k, re, im) with X[k] (or its conjugate) for every bin k in 0..256.
template <typename T, class Sink>
ORCAI_DEV_INLINE void stage_b(int t, const StftTables<T>& tb, const Cx<T>* fbuf, Sink&& sink) {
  constexpr int BR8[8] = {ORCAI_BITREV8_LIST};
  T fr[4][8], fi[4][8];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int k1 = stageb_row(t, s);
    const Cx<T>* row = fbuf + 8 * k1;
    const int x = (k1 >> 1) & 3;
    Cx<T> c[8];
    if constexpr (sizeof(T) == 4) {
      // one 16-byte load per 2-element chunk (rows are 64-byte aligned)
      struct alignas(16) Quad { T a, b, c, d; };
      const Quad* row4 = reinterpret_cast<const Quad*>(row);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const Quad v = row4[q ^ x];
        c[2 * q].x = v.a; c[2 * q].y = v.b; c[2 * q + 1].x = v.c; c[2 * q + 1].y = v.d;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // logical chunk q = columns n2 = 2q, 2q+1 lives in physical chunk q ^ x
        c[2 * q] = row[2 * (q ^ x)];
        c[2 * q + 1] = row[2 * (q ^ x) + 1];
      }
    }
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      fr[s][p] = c[BR8[p]].x;
      fi[s][p] = c[BR8[p]].y;
    }
    orcai_fft8_dit<T>(fr[s], fi[s]);
  }
  const bool t0 = (t == 0);
  // slots 0..7 : rows s=0/1 ; slots 8..15 : rows s=2/3
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    T ar, ai, br, bi;
    int k;
    if (i < 8) {
      // thread 0 pairs inside rows 0 and 16, the others pair row t with row 32-t
      constexpr int K0[8] = {0, 32, 64, 96, 16, 48, 80, 112};
      const int ia0 = (i < 4) ? i : i - 4;                         // thread-0 source of a
      const int ib0 = (i == 0) ? 0 : (i < 4 ? 8 - i : 11 - i);     // thread-0 source of b
      const T a0r = (i < 4) ? fr[0][ia0] : fr[1][ia0];
      const T a0i = (i < 4) ? fi[0][ia0] : fi[1][ia0];
      const T b0r = (i < 4) ? fr[0][ib0] : fr[1][ib0];
      const T b0i = (i < 4) ? fi[0][ib0] : fi[1][ib0];
      ar = t0 ? a0r : fr[0][i];
      ai = t0 ? a0i : fi[0][i];
      br = t0 ? b0r : fr[1][7 - i];
      bi = t0 ? b0i : fi[1][7 - i];
      k = t0 ? K0[i] : t + 32 * i;
    } else {
      const int j = i - 8;
      ar = fr[2][j];
      ai = fi[2][j];
      br = fr[3][7 - j];
      bi = fi[3][7 - j];
      k = 8 + t + 32 * j;
    }
    const Cx<T> c = tb.ck[k];
    const T er = ar + br, ei = ai - bi;   // E = a + conj(b)
    const T dr = ar - br, di = ai + bi;   // D = a - conj(b)
    // G = i * c * D
    const T gr = -(c.x * di + c.y * dr);
    const T gi = c.x * dr - c.y * di;
    sink(k, er - gr, ei - gi);            // X[k]
    sink(256 - k, er + gr, ei + gi);      // conj(X[256-k])
  }
  if (t0) {
    // k = 128: X[128] = conj(Z[128]); Z is carried at half scale
    sink(128, T(2) * fr[0][4], T(2) * fi[0][4]);
  }
}

}  // namespace orcai
