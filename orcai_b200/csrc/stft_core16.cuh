// The fused STFT kernel's per-thread building blocks with SIXTEEN threads per frame (K1, float64 variant).  Host+device like
// stft_core.cuh, so that the index arithmetic is replayed on the CPU (tests/host_emul/stft_emul.cpp).
//
// Why a second decomposition: with 8 threads per frame a thread keeps a 32-point float64 FFT in registers (255 registers, 8 warps
// per SM) and the kernel waits on fixed-latency dependencies with the FP64 pipe 35 % busy.  With 16 threads per frame a thread
// keeps 16 complex values (half the registers, twice the warps).
//
// One 512-sample frame, z[m] = y[2m] + i*y[2m+1], m = 16*n1 + n2, k = k1 + 16*k2 (256-point complex FFT of the packed frame):
//   stage A : thread t = n2 runs a 16-point FFT over n1, applies W256^(n2*k1), writes column t of row k1 of the frame's 16 x 16
//             exchange buffer (row pitch 17 elements: the row reads of stage B are conflict-free)
//   stage B : thread t = k1 reads row k1, runs a 16-point FFT over n2 -> Z[k1 + 16*k2]; it keeps k2 = 0..7 (k = t + 16*k2 <= 127)
//             and passes k2 = 8..15 on through the same buffer (row k1, slot k2 - 8)
//   pairs   : thread t pairs its Z[k], k = t + 16*j, with Z[256 - k] = row (16 - t) % 16, slot 7 - j  (thread 0: slot 8 - j, and
//             k = 0 pairs with itself); X[k] = E - G, X[256-k] = conj(E + G) with E = Z[k] + conj Z[256-k],
//             G = i*c_k*(Z[k] - conj Z[256-k]) exactly as in stft_core.cuh; thread 0 adds X[128] = 2 conj Z[128].
#pragma once
#include "stft_core.cuh"

namespace orcai {

constexpr int kPitch16 = 17;
constexpr int kFrameBuf16Cx = 16 * kPitch16;

// tables: win[256] (Hann pairs * scale, index m), tw16[256] (W256^(n2*k1) at k1*16 + n2), ck[256] (e^{-2 pi i k/512})
template <typename T>
ORCAI_DEV_INLINE void stage_a16(const Cx<float> (&x)[16], int t, const StftTables<T>& tb, Cx<T>* fbuf) {
  constexpr int BR[16] = {ORCAI_BITREV16_LIST};
  T zr[16], zi[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const Cx<T> w = tb.win[16 * BR[p] + t];
    zr[p] = T(x[p].x) * w.x;
    zi[p] = T(x[p].y) * w.y;
  }
  orcai_fft16_dit<T>(zr, zi);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    const Cx<T> w = tb.tw[16 * k1 + t];
    Cx<T> v;
    v.x = zr[k1] * w.x - zi[k1] * w.y;
    v.y = zr[k1] * w.y + zi[k1] * w.x;
    fbuf[kPitch16 * k1 + t] = v;
  }
}

// Stage B, first half: row t -> Z[t + 16*k2] in (fr, fi)[k2].  The caller synchronises the frame's threads, then calls
// stage_b16_pass_on, synchronises again and calls stage_b16_pairs.
template <typename T>
ORCAI_DEV_INLINE void stage_b16_fft(int t, const Cx<T>* fbuf, T (&fr)[16], T (&fi)[16]) {
  constexpr int BR[16] = {ORCAI_BITREV16_LIST};
  const Cx<T>* row = fbuf + kPitch16 * t;
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const Cx<T> c = row[BR[p]];
    fr[p] = c.x;
    fi[p] = c.y;
  }
  orcai_fft16_dit<T>(fr, fi);
}
template <typename T>
ORCAI_DEV_INLINE void stage_b16_pass_on(int t, Cx<T>* fbuf, const T (&fr)[16], const T (&fi)[16]) {
#pragma unroll
  for (int s = 0; s < 8; ++s) fbuf[kPitch16 * t + s] = Cx<T>{fr[8 + s], fi[8 + s]};
}
template <typename T, class Sink>
ORCAI_DEV_INLINE void stage_b16_pairs(int t, const StftTables<T>& tb, const Cx<T>* fbuf, const T (&fr)[16], const T (&fi)[16], Sink&& sink) {
  const Cx<T>* prow = fbuf + kPitch16 * ((16 - t) & 15);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = t + 16 * j;
    const T ar = fr[j], ai = fi[j];
    T br, bi;
    if (j == 0) {
      const Cx<T> b = prow[7];
      br = t ? b.x : ar;                       // k = 0 pairs with itself
      bi = t ? b.y : ai;
    } else {
      const Cx<T> b = prow[t ? 7 - j : 8 - j];
      br = b.x;
      bi = b.y;
    }
    const Cx<T> c = tb.ck[k];
    const T er = ar + br, ei = ai - bi;   // E = a + conj(b)
    const T dr = ar - br, di = ai + bi;   // D = a - conj(b)
    const T gr = -(c.x * di + c.y * dr);  // G = i * c * D
    const T gi = c.x * dr - c.y * di;
    sink(k, er - gr, ei - gi);            // X[k]
    sink(256 - k, er + gr, ei + gi);      // conj(X[256-k])
  }
  if (t == 0) sink(128, T(2) * fr[8], T(2) * fi[8]);   // X[128] = conj(Z[128]); Z is carried at half scale
}

}  // namespace orcai
