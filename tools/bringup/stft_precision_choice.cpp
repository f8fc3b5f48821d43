// Calibration of the per-frame precision choice of the STFT kernel: replays the kernel's float32 and
// float64 dataflow (csrc/stft_core.cuh) on the CPU for every frame of a PCM16 recording and tabulates the dB error of the
// float32 FFT against the frame's dynamic range r = min in-band power / mean power over all bins.
//   python -c "from orcai_b200.synth import synth_pcm16; synth_pcm16(600).tofile('/tmp/p.i16')"
//   g++ -O2 -std=c++17 -I orcai_b200/csrc tools/bringup/stft_precision_choice.cpp -o /tmp/spc && /tmp/spc /tmp/p.i16
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>
#include "stft_core.cuh"
#include "stft_tables.h"
using namespace orcai;

template <typename T>
static void frame_power(const float* x, const StftTables<T>& tb, std::vector<Cx<T>>& fbuf, float* pw) {
  static const int BR[32] = {ORCAI_BITREV32_LIST};
  for (int t = 0; t < 8; ++t) {
    Cx<float> xin[32];
    for (int p = 0; p < 32; ++p) { int m = 8 * BR[p] + t; xin[p].x = x[2 * m]; xin[p].y = x[2 * m + 1]; }
    stage_a<T>(xin, t, tb, fbuf.data());
  }
  for (int t = 0; t < 8; ++t)
    stage_b<T>(t, tb, fbuf.data(), [&](int k, T re, T im) { const float fr = (float)re, fi = (float)im; pw[k] = std::fmaf(fr, fr, fi * fi); });
}

int main(int argc, char** argv) {
  FILE* f = std::fopen(argv[1], "rb");
  std::vector<int16_t> pcm;
  { int16_t buf[65536]; size_t n; while ((n = std::fread(buf, 2, 65536, f)) > 0) pcm.insert(pcm.end(), buf, buf + n); }
  const long long N = (long long)pcm.size(), T = N / 256 + 1;
  auto hf = make_stft_tables<float>(0.5 / 32768.0);
  auto hd = make_stft_tables<double>(0.5 / 32768.0);
  StftTables<float> tf{(const Cx<float>*)hf.win.data(), (const Cx<float>*)hf.tw.data(), (const Cx<float>*)hf.ck.data()};
  StftTables<double> td{(const Cx<double>*)hd.win.data(), (const Cx<double>*)hd.tw.data(), (const Cx<double>*)hd.ck.data()};
  std::vector<Cx<float>> bf(kFrameBufCx);
  std::vector<Cx<double>> bd(kFrameBufCx);
  std::vector<float> PF((size_t)T * 257), PD((size_t)T * 257);
  float gmax = 0;
  for (long long j = 0; j < T; ++j) {
    float x[512];
    for (int n = 0; n < 512; ++n) { long long s = (j - 1) * 256 + n; x[n] = (s >= 0 && s < N) ? (float)pcm[s] : 0.f; }
    frame_power<float>(x, tf, bf, &PF[j * 257]);
    frame_power<double>(x, td, bd, &PD[j * 257]);
    for (int k = 0; k < 257; ++k) gmax = std::max(gmax, PD[j * 257 + k]);
  }
  const float floor_p = gmax * 1e-8f;
  const int NB = 24;  // half-decade buckets of -log10(r)
  double worst[NB] = {0}, worst_fl[NB] = {0};
  long long cnt[NB] = {0};
  auto db = [](float p) { return 10.0 * std::log10((double)std::max(p, 1e-10f)); };
  for (long long j = 0; j < T; ++j) {
    const float* pf = &PF[j * 257]; const float* pd = &PD[j * 257];
    double sum = 0; float mn = 1e30f;
    for (int k = 0; k < 257; ++k) sum += pf[k];
    for (int k = 0; k < 171; ++k) mn = std::min(mn, pf[k]);
    const double mean = sum / 257.0;
    int b = mean > 0 ? (int)std::floor(-2.0 * std::log10(std::max((double)mn, 1e-300) / mean)) : 0;
    b = std::min(std::max(b, 0), NB - 1);
    cnt[b]++;
    for (int k = 0; k < 171; ++k) {
      const double e = std::fabs(db(pf[k]) - db(pd[k]));
      worst[b] = std::max(worst[b], e);
      // what a consumer sees: max(L - Lref, -80)
      const double ef = std::fabs(std::max(db(pf[k]) - db(gmax), -80.0) - std::max(db(pd[k]) - db(gmax), -80.0));
      worst_fl[b] = std::max(worst_fl[b], ef);
    }
  }
  std::printf("frames %lld gmax %.4g\n  -log10(min_inband/mean)   frames   max dB err   max dB err after the -80 dB floor\n", T, gmax);
  for (int b = 0; b < NB; ++b) std::printf("  [%4.1f,%4.1f)  %9lld   %.3e   %.3e\n", b * 0.5, b * 0.5 + 0.5, cnt[b], worst[b], worst_fl[b]);
  return 0;
}
