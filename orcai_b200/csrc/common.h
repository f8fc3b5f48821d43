// Shared declarations of liborcai_b200: context, error handling, stage launchers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/orcai_b200.h"

namespace orcai {

constexpr int kRawLd = 176;          // row pitch (floats) of the raw dB buffer: 704 B rows, 32 B-sector aligned
constexpr int kMaxBlocks = 8;

struct SelectState {                 // device-resident scratch of the exact two-rank radix select
  unsigned long long hist[2][2048];  // per-rank digit histograms of the current pass
  unsigned long long rank[2];        // rank still to be located inside the current prefix
  unsigned int prefix[2];            // key bits decided so far (high bits)
  unsigned int pmax_bits;            // max |S|^2 over all bins and frames (float bits, non-negative)
  float db_ref;                      // 10*log10(max(1e-10, pmax))
  float lo, hi;                      // selected percentiles (shifted + floored dB)
  unsigned int tile_counter;         // scan kernel dynamic tile id
  int precise_log;                   // 1: dB through log10f (float64 STFT variant), 0: through lg2.approx (fast variant)
};

#ifdef __CUDACC__
// 10*log10(max(1e-10, p)): the one expression K1 and the select's reference level share, so that the loudest cell is exactly 0 dB
__device__ __forceinline__ float power_to_db(float p, bool precise) {
  const float q = fmaxf(p, 1e-10f);
  return precise ? 10.0f * log10f(q) : 3.01029995663981195f * __log2f(q);
}
#endif

struct NetWeights;                   // net.cu

// One begun-but-not-collected predict call (orcai_predict_resident_begin / _end): its own post-processing scratch and
// page-locked result staging, so that the next recording can be enqueued behind it before the host has read its results.
struct AsyncSlot {
  bool busy = false;
  cudaEvent_t done = nullptr;                 // blocking-sync event: the host thread sleeps, it does not spin
  void* d_post = nullptr; size_t post_cap = 0;
  void* h_pin = nullptr;  size_t pin_cap = 0;
  int64_t T = 0;
  long long cap = 0, spec = 0, n_agg = 0, n_cnt = 0;
  bool want_agg = false;
  size_t o_lab = 0, o_sta = 0, o_sto = 0, o_agg = 0, o_cnt = 0;
  const int* d_lab = nullptr; const long long* d_sta = nullptr; const long long* d_sto = nullptr;
};
constexpr int kAsyncDepth = 2;
constexpr size_t kStageTotalOff = 0, kStageStatsOff = 64, kStageSegsOff = 256;   // layout of the head of a slot's staging

struct Ctx {
  int device = 0;
  orcai_params p{};
  cudaStream_t stream = nullptr;
  std::string err;
  // STFT tables on device: [0]=float input scale, [1]=int16 input scale ; each 768 float2
  float* d_tables[2] = {nullptr, nullptr};
  double* d_tables64[2] = {nullptr, nullptr};
  double* d_tables64_16[2] = {nullptr, nullptr};   // float64 tables of the 16-threads-per-frame kernel (16 x 16 stage-A twiddles)
  int h_flags[2] = {0, 1};           // stable host source for tiny async H2D copies
  SelectState* d_sel = nullptr;
  // current recording
  void* d_pcm = nullptr;  size_t pcm_cap = 0;  int pcm_dtype = 0;  int64_t n_samples = 0;
  // next recording, uploaded on the copy stream while the current one is being annotated (orcai_prefetch_pcm)
  void* d_pcm_next = nullptr;  size_t pcm_next_cap = 0;  int pcm_next_dtype = 0;  int64_t n_samples_next = -1;
  cudaStream_t copy_stream = nullptr;  cudaEvent_t ev_prefetch = nullptr;
  float* d_raw = nullptr; size_t raw_cap = 0;  int64_t T = 0;      // raw dB (T, kRawLd)
  float* d_spec = nullptr; size_t spec_cap = 0;                    // normalised (T, n_freq) compact
  bool have_stats = false;
  // network
  NetWeights* net = nullptr;
  float* d_preds = nullptr; size_t preds_cap = 0;                  // (N, pred_len, n_labels)
  // post-processing scratch
  void* d_post = nullptr; size_t post_cap = 0;
  // pinned staging
  void* h_pin = nullptr; size_t pin_cap = 0;
  void* h_small = nullptr;           // 256 pinned bytes: statistics read-back
  // timing / accounting of the last call
  orcai_timings tm{};
  cudaEvent_t ev[16] = {};
  uint64_t launches = 0;
  int sm_count = 148;
  int stft_f64 = 1;                   // 1: float64 FFT (parity grade, default), 0: float32 FFT (fast variant)
  int stft_threads = 16;              // threads per frame of the float64 kernel: 16 (default) or 8 (the round-1 decomposition)
  // asynchronous predict calls in flight (oldest first: async_head)
  AsyncSlot slot[kAsyncDepth];
  int async_head = 0, async_pending = 0;
};

#define ORCAI_CUDA(ctx, call)                                                            \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      char b__[512];                                                                     \
      snprintf(b__, sizeof b__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      (ctx)->err = b__;                                                                  \
      return ORCAI_ERR_CUDA;                                                             \
    }                                                                                    \
  } while (0)

#define ORCAI_FAIL(ctx, code, ...)                       \
  do {                                                   \
    char b__[512];                                       \
    snprintf(b__, sizeof b__, __VA_ARGS__);              \
    (ctx)->err = b__;                                    \
    return (code);                                       \
  } while (0)

#define ORCAI_CHECK(expr)                                \
  do {                                                   \
    int rc__ = (expr);                                   \
    if (rc__ != ORCAI_OK) return rc__;                   \
  } while (0)

int ensure_device_buffer(Ctx* c, void** p, size_t* cap, size_t bytes);
int ensure_pinned(Ctx* c, size_t bytes);   // c->h_pin: pinned host staging of at least `bytes` (post.cu)

// ---- stage launchers (all asynchronous on c->stream) -------------------------------------------
// K1: fused window + rFFT512 + |.|^2 + 10log10 + crop ; also the global power max.   (stft.cu)
int stft_upload_tables(Ctx* c);
// stat rows: the frames that count towards the global power maximum (all of them, or the rows a time chunk owns)
int launch_stft(Ctx* c, const void* d_pcm, int dtype, int64_t n_samples, int64_t T, float* d_raw, int64_t stat_row0, int64_t stat_row1);
// exact percentiles (radix select on shifted/floored dB) and K2 normalise.          (select.cu)
int launch_select(Ctx* c, const float* d_raw, int64_t T);
// time chunks of ONE recording (orcai_chunk_*): the select's passes driven from the host, which sums the chunks' histograms
int launch_select_begin(Ctx* c, float max_power);                                              // global max -> db_ref
int launch_select_histogram(Ctx* c, const float* d_raw_rows, int64_t n_rows, int pass, const uint32_t prefix[2], uint64_t* h_hist);
int launch_select_end(Ctx* c, const uint32_t key[2]);                                          // decided keys -> lo, hi
int launch_normalise(Ctx* c, const float* d_raw, int64_t T, float* d_spec);
int launch_read_db(Ctx* c, const float* d_raw, int64_t T, float* d_out);             // shifted + floored dB, compact
// network forward                                                                    (net.cu)
int net_create(Ctx* c);
void net_destroy(Ctx* c);
int net_set_chunk(Ctx* c, int chunk);
void net_collect_stage_times(Ctx* c);
int net_set_path(Ctx* c, int path);
int net_set_tail_path(Ctx* c, int path);
int net_calibrate(Ctx* c, int64_t max_snippets);
int net_set_conv0_path(Ctx* c, int path);
int net_set_block1_path(Ctx* c, int path);
int net_set_precise_tall(Ctx* c, int on);   // net_path 4 on resident recordings: 1 = shared interior (tall image + border rows), 0 = snippet by snippet
void net_trace_dump();                  // bring-up aid (net_fused.cuh: ORCAI_FUSED_TRACE), no-op in normal builds
const unsigned int* net_trap_info();   // bring-up aid (tc_common.cuh: g_trap_info), nullptr unless ORCAI_B200_TRAPINFO is set
int net_set_debug_stop(Ctx* c, int stage);
int net_debug_read(Ctx* c, float* out_host, int64_t capacity, int64_t* dims_out);
int net_load_weights(Ctx* c, const char* const* names, const float* const* data, const int64_t* sizes, int n);
// input_mode 0: raw dB buffer (pitch kRawLd, normalise on load with c->d_sel stats), snippet i starts at row
//               (first + i) * shift ; input_mode 1: normalised compact snippets (pitch n_freq), snippet stride = snippet_len rows
int net_forward(Ctx* c, const float* d_in, int input_mode, int64_t first, int64_t n, float* d_preds);
// post-processing: overlap-average + threshold + run-length segments                 (post.cu)
int launch_postprocess(Ctx* c, const float* d_preds, int64_t n_snippets, int64_t T, double threshold,
                       double* h_agg, double* h_cnt, int32_t* h_label, int64_t* h_start, int64_t* h_stop,
                       int64_t cap, int64_t* n_seg);
// the same in two halves for orcai_predict_resident_begin / _end: everything (kernels and the copies into the slot's page-locked
// staging) is enqueued by _begin; _end interprets the staging after the slot's `done` event
int postprocess_begin(Ctx* c, const float* d_preds, int64_t n_snippets, int64_t T, double threshold, bool want_agg, int64_t cap, AsyncSlot* s);
int postprocess_end(Ctx* c, AsyncSlot* s, double* h_agg, double* h_cnt, int32_t* h_label, int64_t* h_start, int64_t* h_stop,
                    int64_t cap, int64_t* n_seg);
int launch_threshold_segments(Ctx* c, const double* h_agg, const double* h_cnt, int64_t S, int L, double threshold,
                              int32_t* h_label, int64_t* h_start, int64_t* h_stop, int64_t cap, int64_t* n_seg);

}  // namespace orcai
