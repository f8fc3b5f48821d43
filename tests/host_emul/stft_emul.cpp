// CPU replay of the STFT kernel's 8-threads-per-frame dataflow (stage A -> exchange buffer -> stage B).
// Checks the index arithmetic of csrc/stft_core.cuh against a naive float64 DFT.  Build: g++ -O2 -I orcai_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>
#include <random>
#include "stft_core.cuh"
#include "stft_tables.h"

int main() {
  using namespace orcai;
  StftHostTables ht = make_stft_tables(0.5);
  StftTables tb{reinterpret_cast<const float2*>(ht.win.data()), reinterpret_cast<const float2*>(ht.tw.data()),
                reinterpret_cast<const float2*>(ht.ck.data())};
  std::mt19937 rng(7);
  std::normal_distribution<float> nd(0.f, 0.1f);
  const double PI = 3.14159265358979323846;
  double worst = 0;
  int seen_total = 0;
  for (int trial = 0; trial < 20; ++trial) {
    std::vector<float> x(512);
    for (auto& v : x) v = nd(rng);
    if (trial == 1) for (int n = 0; n < 512; ++n) x[n] = 0.5f * std::sin(2 * PI * 37.0 * n / 512.0);
    if (trial == 2) { for (auto& v : x) v = 0; x[100] = 1.f; }
    if (trial == 3) for (auto& v : x) v = 0.25f;
    // reference: float64 windowed DFT
    std::vector<double> pref(257);
    for (int k = 0; k <= 256; ++k) {
      std::complex<double> acc = 0;
      for (int n = 0; n < 512; ++n) {
        double h = 0.5 - 0.5 * std::cos(2 * PI * n / 512.0);
        acc += double(x[n]) * h * std::exp(std::complex<double>(0, -2 * PI * k * n / 512.0));
      }
      pref[k] = std::norm(acc);
    }
    std::vector<float2> fbuf(kFrameBufFloat2);
    static const int BR[32] = {ORCAI_BITREV32_LIST};
    for (int t = 0; t < 8; ++t) {
      float2 xin[32];
      for (int p = 0; p < 32; ++p) { int m = 8 * BR[p] + t; xin[p].x = x[2 * m]; xin[p].y = x[2 * m + 1]; }
      stage_a(xin, t, tb, fbuf.data());
    }
    std::vector<double> got(257, -1.0);
    std::vector<int> cnt(257, 0);
    for (int t = 0; t < 8; ++t)
      stage_b(t, tb, fbuf.data(), [&](int k, float p) { got[k] = p; cnt[k]++; });
    double pmax = 0;
    for (int k = 0; k <= 256; ++k) pmax = std::max(pmax, pref[k]);
    for (int k = 0; k <= 256; ++k) {
      if (cnt[k] != 1) { std::printf("FAIL bin %d written %d times\n", k, cnt[k]); return 1; }
      seen_total++;
      double err = std::fabs(got[k] - pref[k]) / (pref[k] + 1e-7 * pmax);
      worst = std::max(worst, err);
    }
  }
  std::printf("bins checked %d, worst relative power error %.3e\n", seen_total, worst);
  if (worst > 2e-4) { std::printf("FAIL\n"); return 1; }
  std::printf("OK\n");
  return 0;
}
