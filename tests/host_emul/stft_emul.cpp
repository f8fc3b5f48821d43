// CPU replay of the STFT kernel's 8-threads-per-frame dataflow (stage A -> exchange buffer -> stage B).
// Checks the index arithmetic of csrc/stft_core.cuh against a naive float64 DFT, for the float and the
// double instantiation.  Build: g++ -O2 -std=c++17 -I orcai_b200/csrc tests/host_emul/stft_emul.cpp
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>
#include <random>
#include "stft_core.cuh"
#include "stft_core16.cuh"
#include "stft_tables.h"

template <typename T>
double run(double tol) {
  using namespace orcai;
  StftHostTables<T> ht = make_stft_tables<T>(0.5);
  StftTables<T> tb{reinterpret_cast<const Cx<T>*>(ht.win.data()), reinterpret_cast<const Cx<T>*>(ht.tw.data()),
                   reinterpret_cast<const Cx<T>*>(ht.ck.data())};
  std::mt19937 rng(7);
  std::normal_distribution<float> nd(0.f, 0.1f);
  const double PI = 3.14159265358979323846;
  double worst = 0;
  for (int trial = 0; trial < 20; ++trial) {
    std::vector<float> x(512);
    for (auto& v : x) v = nd(rng);
    if (trial == 1) for (int n = 0; n < 512; ++n) x[n] = 0.5f * std::sin(2 * PI * 37.0 * n / 512.0);
    if (trial == 2) { for (auto& v : x) v = 0; x[100] = 1.f; }
    if (trial == 3) for (auto& v : x) v = 0.25f;
    std::vector<double> pref(257);
    for (int k = 0; k <= 256; ++k) {
      std::complex<double> acc = 0;
      for (int n = 0; n < 512; ++n) {
        double h = 0.5 - 0.5 * std::cos(2 * PI * n / 512.0);
        acc += double(x[n]) * h * std::exp(std::complex<double>(0, -2 * PI * k * n / 512.0));
      }
      pref[k] = std::norm(acc);
    }
    std::vector<Cx<T>> fbuf(kFrameBufCx);
    static const int BR[32] = {ORCAI_BITREV32_LIST};
    for (int t = 0; t < 8; ++t) {
      Cx<float> xin[32];
      for (int p = 0; p < 32; ++p) { int m = 8 * BR[p] + t; xin[p].x = x[2 * m]; xin[p].y = x[2 * m + 1]; }
      stage_a<T>(xin, t, tb, fbuf.data());
    }
    std::vector<double> got(257, -1.0);
    std::vector<int> cnt(257, 0);
    for (int t = 0; t < 8; ++t)
      stage_b<T>(t, tb, fbuf.data(), [&](int k, T re, T im) { got[k] = double(re) * re + double(im) * im; cnt[k]++; });
    double pmax = 0;
    for (int k = 0; k <= 256; ++k) pmax = std::max(pmax, pref[k]);
    for (int k = 0; k <= 256; ++k) {
      if (cnt[k] != 1) { std::printf("FAIL bin %d written %d times\n", k, cnt[k]); std::exit(1); }
      worst = std::max(worst, std::fabs(got[k] - pref[k]) / (pref[k] + 1e-7 * pmax));
    }
  }
  std::printf("sizeof(T)=%zu worst relative power error %.3e (tol %.1e)\n", sizeof(T), worst, tol);
  if (worst > tol) { std::printf("FAIL\n"); std::exit(1); }
  return worst;
}

// the 16-threads-per-frame decomposition (stft_core16.cuh): stage A -> exchange -> 16-point FFTs -> pass-on -> pairs
template <typename T>
double run16(double tol) {
  using namespace orcai;
  StftHostTables<T> ht = make_stft_tables<T>(0.5, 16);
  StftTables<T> tb{reinterpret_cast<const Cx<T>*>(ht.win.data()), reinterpret_cast<const Cx<T>*>(ht.tw.data()),
                   reinterpret_cast<const Cx<T>*>(ht.ck.data())};
  std::mt19937 rng(11);
  std::normal_distribution<float> nd(0.f, 0.1f);
  const double PI = 3.14159265358979323846;
  double worst = 0;
  for (int trial = 0; trial < 20; ++trial) {
    std::vector<float> x(512);
    for (auto& v : x) v = nd(rng);
    if (trial == 1) for (int n = 0; n < 512; ++n) x[n] = 0.5f * std::sin(2 * PI * 37.0 * n / 512.0);
    if (trial == 2) { for (auto& v : x) v = 0; x[100] = 1.f; }
    if (trial == 3) for (auto& v : x) v = 0.25f;
    if (trial == 4) for (int n = 0; n < 512; ++n) x[n] = (n & 1) ? -0.3f : 0.3f;   // Nyquist
    std::vector<double> pref(257);
    for (int k = 0; k <= 256; ++k) {
      std::complex<double> acc = 0;
      for (int n = 0; n < 512; ++n) {
        double h = 0.5 - 0.5 * std::cos(2 * PI * n / 512.0);
        acc += double(x[n]) * h * std::exp(std::complex<double>(0, -2 * PI * k * n / 512.0));
      }
      pref[k] = std::norm(acc);
    }
    std::vector<Cx<T>> fbuf(kFrameBuf16Cx);
    static const int BR[16] = {ORCAI_BITREV16_LIST};
    for (int t = 0; t < 16; ++t) {
      Cx<float> xin[16];
      for (int p = 0; p < 16; ++p) { int m = 16 * BR[p] + t; xin[p].x = x[2 * m]; xin[p].y = x[2 * m + 1]; }
      stage_a16<T>(xin, t, tb, fbuf.data());
    }
    T fr[16][16], fi[16][16];
    for (int t = 0; t < 16; ++t) stage_b16_fft<T>(t, fbuf.data(), fr[t], fi[t]);
    for (int t = 0; t < 16; ++t) stage_b16_pass_on<T>(t, fbuf.data(), fr[t], fi[t]);
    std::vector<double> got(257, -1.0);
    std::vector<int> cnt(257, 0);
    for (int t = 0; t < 16; ++t)
      stage_b16_pairs<T>(t, tb, fbuf.data(), fr[t], fi[t], [&](int k, T re, T im) { got[k] = double(re) * re + double(im) * im; cnt[k]++; });
    double pmax = 0;
    for (int k = 0; k <= 256; ++k) pmax = std::max(pmax, pref[k]);
    for (int k = 0; k <= 256; ++k) {
      if (cnt[k] != 1) { std::printf("FAIL (16) bin %d written %d times\n", k, cnt[k]); std::exit(1); }
      worst = std::max(worst, std::fabs(got[k] - pref[k]) / (pref[k] + 1e-7 * pmax));
    }
  }
  std::printf("16 threads per frame: sizeof(T)=%zu worst relative power error %.3e (tol %.1e)\n", sizeof(T), worst, tol);
  if (worst > tol) { std::printf("FAIL\n"); std::exit(1); }
  return worst;
}

int main() {
  run<float>(2e-4);
  run<double>(1e-9);
  run16<float>(2e-4);
  run16<double>(1e-9);
  std::printf("OK\n");
  return 0;
}
