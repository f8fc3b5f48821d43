// Host-side construction of the STFT kernel's lookup tables (double precision, rounded once to T).
#pragma once
#include <cmath>
#include <vector>

namespace orcai {

template <typename T>
struct StftHostTables {
  std::vector<T> win;  // 256 x complex: (h[2m], h[2m+1]) * scale
  std::vector<T> tw;   // 256 x complex: W256^(n2*k1) at index k1*8 + n2
  std::vector<T> ck;   // 256 x complex: exp(-2 pi i k / 512)
};

// scale = 0.5 for float input in [-1,1]; 0.5/32768 for raw int16 input (both exact powers of two).
// rows = 32: the 32 x 8 decomposition of stft_core.cuh (tw at k1*8 + n2); rows = 16: the 16 x 16 one of stft_core16.cuh (k1*16 + n2)
template <typename T>
inline StftHostTables<T> make_stft_tables(double scale, int rows = 32) {
  const double PI = 3.14159265358979323846;
  StftHostTables<T> t;
  t.win.resize(512);
  t.tw.resize(512);
  t.ck.resize(512);
  for (int m = 0; m < 256; ++m) {
    // periodic Hann, scipy.signal.get_window("hann", 512, fftbins=True)
    const double h0 = 0.5 - 0.5 * std::cos(2.0 * PI * (2 * m) / 512.0);
    const double h1 = 0.5 - 0.5 * std::cos(2.0 * PI * (2 * m + 1) / 512.0);
    t.win[2 * m] = static_cast<T>(h0 * scale);
    t.win[2 * m + 1] = static_cast<T>(h1 * scale);
  }
  const int cols = 256 / rows;
  for (int k1 = 0; k1 < rows; ++k1)
    for (int n2 = 0; n2 < cols; ++n2) {
      const double a = -2.0 * PI * ((n2 * k1) % 256) / 256.0;
      t.tw[2 * (cols * k1 + n2)] = static_cast<T>(std::cos(a));
      t.tw[2 * (cols * k1 + n2) + 1] = static_cast<T>(std::sin(a));
    }
  for (int k = 0; k < 256; ++k) {
    const double a = -2.0 * PI * k / 512.0;
    t.ck[2 * k] = static_cast<T>(std::cos(a));
    t.ck[2 * k + 1] = static_cast<T>(std::sin(a));
  }
  return t;
}

}  // namespace orcai
