// extern "C" surface of liborcai_b200 (declared in include/orcai_b200.h).
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <new>

#include "common.h"
#include "stft_tables.h"

using namespace orcai;

struct orcai_ctx : public orcai::Ctx {};

namespace orcai {

int ensure_device_buffer(Ctx* c, void** p, size_t* cap, size_t bytes) {
  if (bytes <= *cap && *p) return ORCAI_OK;
  if (*p) { ORCAI_CUDA(c, cudaFree(*p)); *p = nullptr; *cap = 0; }
  if (bytes == 0) bytes = 256;
  ORCAI_CUDA(c, cudaMalloc(p, bytes));
  *cap = bytes;
  return ORCAI_OK;
}

}  // namespace orcai

namespace {

thread_local std::string g_create_error;

enum { EV_START = 0, EV_H2D, EV_STFT, EV_SELECT, EV_NORM, EV_NET, EV_POST, EV_END };

int rec(Ctx* c, int which) {
  ORCAI_CUDA(c, cudaEventRecord(c->ev[which], c->stream));
  return ORCAI_OK;
}

float elapsed(Ctx* c, int a, int b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]) != cudaSuccess) { cudaGetLastError(); return 0.f; }
  return ms;
}

struct StatsTail { unsigned long long rank[2]; unsigned int prefix[2]; unsigned int pmax_bits; float db_ref, lo, hi; };

// selected statistics of the resident recording: enqueue the read-back into pinned memory, interpret it after the next
// stream synchronise (so that a predict call pays one synchronise for statistics, segments and aggregates together)
int stats_enqueue(Ctx* c) {
  if (!c->h_small) ORCAI_CUDA(c, cudaMallocHost(&c->h_small, 256));
  ORCAI_CUDA(c, cudaMemcpyAsync(c->h_small, reinterpret_cast<const unsigned char*>(c->d_sel) + offsetof(SelectState, rank), sizeof(StatsTail),
                                cudaMemcpyDeviceToHost, c->stream));
  return ORCAI_OK;
}

void stats_finish(Ctx* c, orcai_spec_stats* stats) {
  if (!stats) return;
  StatsTail t;
  memcpy(&t, c->h_small, sizeof t);
  const int nb = c->p.band_hi - c->p.band_lo;
  const unsigned long long n = (unsigned long long)c->T * (unsigned long long)nb;
  stats->n_frames = c->T;
  memcpy(&stats->ref_power, &t.pmax_bits, 4);
  stats->db_ref = t.db_ref;
  stats->lo = t.lo;
  stats->hi = t.hi;
  stats->rank_lo = (int64_t)nearbyint((double)(n - 1) * c->p.q_lo);
  stats->rank_hi = (int64_t)nearbyint((double)(n - 1) * c->p.q_hi);
}

int fill_stats(Ctx* c, orcai_spec_stats* stats) {
  if (!stats) return ORCAI_OK;
  ORCAI_CHECK(stats_enqueue(c));
  ORCAI_CUDA(c, cudaStreamSynchronize(c->stream));
  stats_finish(c, stats);
  return ORCAI_OK;
}

int spectrogram_stages(Ctx* c, int normalise) {
  if (!c->d_pcm || c->n_samples < 0) ORCAI_FAIL(c, ORCAI_ERR_STATE, "no recording uploaded (orcai_upload_pcm)");
  c->T = orcai_num_frames(c->n_samples, c->p.hop);
  {
    void* p = c->d_raw;
    ORCAI_CHECK(ensure_device_buffer(c, &p, &c->raw_cap, (size_t)c->T * kRawLd * sizeof(float)));
    c->d_raw = static_cast<float*>(p);
  }
  ORCAI_CHECK(launch_stft(c, c->d_pcm, c->pcm_dtype, c->n_samples, c->T, c->d_raw, 0, c->T));
  ORCAI_CHECK(rec(c, EV_STFT));
  ORCAI_CHECK(launch_select(c, c->d_raw, c->T));
  ORCAI_CHECK(rec(c, EV_SELECT));
  if (normalise) {
    void* p = c->d_spec;
    ORCAI_CHECK(ensure_device_buffer(c, &p, &c->spec_cap, (size_t)c->T * c->p.n_freq * sizeof(float)));
    c->d_spec = static_cast<float*>(p);
    ORCAI_CHECK(launch_normalise(c, c->d_raw, c->T, c->d_spec));
  }
  ORCAI_CHECK(rec(c, EV_NORM));
  c->have_stats = true;
  return ORCAI_OK;
}

int ensure_preds(Ctx* c, int64_t n) {
  const int Tn = c->p.snippet_len >> c->p.n_blocks;
  void* p = c->d_preds;
  ORCAI_CHECK(ensure_device_buffer(c, &p, &c->preds_cap, (size_t)n * Tn * c->p.n_labels * sizeof(float)));
  c->d_preds = static_cast<float*>(p);
  return ORCAI_OK;
}

}  // namespace

extern "C" {

int orcai_version(void) { return 100; }

void* orcai_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}

void orcai_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

const char* orcai_last_error(const orcai_ctx* ctx) {
  if (!ctx) return g_create_error.c_str();
  const unsigned int* ti = orcai::net_trap_info();
  if (ti != nullptr && ti[0] == 0x7241u) {
    static thread_local std::string s;
    char b[160];
    snprintf(b, sizeof b, " [mbarrier wait gave up: block %u thread %u barrier@0x%x parity %u]", ti[1], ti[2], ti[3], ti[4]);
    s = ctx->err + b;
    return s.c_str();
  }
  return ctx->err.c_str();
}

int64_t orcai_num_frames(int64_t n_samples, int32_t hop) { return hop > 0 && n_samples >= 0 ? 1 + n_samples / hop : 0; }

int64_t orcai_num_snippets(int64_t n_frames, int32_t snippet_len) {
  const int64_t shift = snippet_len / 2;
  if (shift <= 0 || n_frames < snippet_len) return 0;
  return (n_frames - snippet_len) / shift + 1;
}

int orcai_create(int device, const orcai_params* p, orcai_ctx** out) {
  if (!p || !out) { g_create_error = "null argument"; return ORCAI_ERR_ARG; }
  *out = nullptr;
  if (p->n_fft != 512 || p->hop != 256) { g_create_error = "K1 is built for n_fft=512, hop=256"; return ORCAI_ERR_ARG; }
  if (p->band_lo < 0 || p->band_hi > 257 || p->band_lo >= p->band_hi || p->band_hi - p->band_lo > kRawLd) {
    g_create_error = "bad frequency band"; return ORCAI_ERR_ARG;
  }
  if (p->n_freq != p->band_hi - p->band_lo) { g_create_error = "n_freq != band_hi - band_lo"; return ORCAI_ERR_ARG; }
  if (!(p->q_lo >= 0.0 && p->q_lo <= 1.0 && p->q_hi >= 0.0 && p->q_hi <= 1.0)) { g_create_error = "quantiles outside [0,1]"; return ORCAI_ERR_ARG; }
  if (p->n_blocks < 1 || p->n_blocks > kMaxBlocks || p->snippet_len < (1 << p->n_blocks) || p->n_labels < 1) {
    g_create_error = "bad network geometry"; return ORCAI_ERR_ARG;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(e);
    cudaGetLastError();
    return ORCAI_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return ORCAI_ERR_ARG; }
  orcai_ctx* c = new (std::nothrow) orcai_ctx();
  if (!c) { g_create_error = "out of host memory"; return ORCAI_ERR_ARG; }
  c->device = device;
  c->p = *p;
  auto fail = [&](const char* what, cudaError_t err) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
    orcai_destroy(c);
    return ORCAI_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
  if (prop.major != 10) {
    g_create_error = "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ", this build targets sm_100a";
    orcai_destroy(c);
    return ORCAI_ERR_CUDA;
  }
  c->sm_count = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  if ((e = cudaEventCreateWithFlags(&c->ev_prefetch, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
  for (auto& ev : c->ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return fail("cudaEventCreate", e);
  if (stft_upload_tables(c) != ORCAI_OK) { g_create_error = c->err; orcai_destroy(c); return ORCAI_ERR_CUDA; }
  if ((e = cudaMalloc(&c->d_sel, sizeof(SelectState))) != cudaSuccess) return fail("cudaMalloc", e);
  if ((e = cudaMemset(c->d_sel, 0, sizeof(SelectState))) != cudaSuccess) return fail("cudaMemset", e);
  net_create(c);
  *out = c;
  return ORCAI_OK;
}

void orcai_destroy(orcai_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  net_destroy(c);
  for (auto& t : c->d_tables) if (t) cudaFree(t);
  for (auto& t : c->d_tables64) if (t) cudaFree(t);
  for (auto& t : c->d_tables64_16) if (t) cudaFree(t);
  if (c->d_sel) cudaFree(c->d_sel);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->d_pcm) cudaFree(c->d_pcm);
  if (c->d_pcm_next) cudaFree(c->d_pcm_next);
  if (c->ev_prefetch) cudaEventDestroy(c->ev_prefetch);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->d_raw) cudaFree(c->d_raw);
  if (c->d_spec) cudaFree(c->d_spec);
  if (c->d_preds) cudaFree(c->d_preds);
  if (c->d_post) cudaFree(c->d_post);
  if (c->h_pin) cudaFreeHost(c->h_pin);
  if (c->h_small) cudaFreeHost(c->h_small);
  for (auto& sl : c->slot) {
    if (sl.d_post) cudaFree(sl.d_post);
    if (sl.h_pin) cudaFreeHost(sl.h_pin);
    if (sl.done) cudaEventDestroy(sl.done);
  }
  for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
  if (c->stream) cudaStreamDestroy(c->stream);
  cudaGetLastError();
  delete c;
}

int orcai_get_timings(const orcai_ctx* c, orcai_timings* out) {
  if (!c || !out) return ORCAI_ERR_ARG;
  *out = c->tm;
  out->kernel_launches = c->launches;
  return ORCAI_OK;
}

int orcai_calibrate(orcai_ctx* c, int64_t max_snippets) {
  if (!c) return ORCAI_ERR_ARG;
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  return net_calibrate(c, max_snippets);
}

int orcai_set_option(orcai_ctx* c, const char* key, int64_t value) {
  if (!c || !key) return ORCAI_ERR_ARG;
  if (!strcmp(key, "chunk")) return net_set_chunk(c, (int)value);
  if (!strcmp(key, "tail_path")) return net_set_tail_path(c, (int)value);
  if (!strcmp(key, "conv0_path")) return net_set_conv0_path(c, (int)value);
  if (!strcmp(key, "block1_path")) return net_set_block1_path(c, (int)value);
  if (!strcmp(key, "precise_tall")) return net_set_precise_tall(c, (int)value);
  if (!strcmp(key, "precise_sep_path")) return net_set_precise_tall(c, value ? 3 : 2);
  if (!strcmp(key, "precise_lstm_tc")) return net_set_precise_tall(c, value ? 5 : 4);   // net_path 4: split-fp16 tensor-core recurrence (1, default) or the fp32 CUDA-core one (0)
  if (!strcmp(key, "net_path")) {
    if (value < 0 || value > 4)
      ORCAI_FAIL(c, ORCAI_ERR_ARG, "net_path must be 0 (fp32), 1 (fp16 tensor cores), 2 (bf16 tensor cores), 3 (fp16 fused residual blocks) or 4 (split-fp16 tensor cores, fp32 grade)");
    return net_set_path(c, (int)value);
  }
  if (!strcmp(key, "debug_stop")) return net_set_debug_stop(c, (int)value);
  if (!strcmp(key, "stft_f64")) { c->stft_f64 = value ? 1 : 0; c->have_stats = false; return ORCAI_OK; }
  if (!strcmp(key, "stft_threads")) { c->stft_threads = value == 8 ? 8 : 16; c->have_stats = false; return ORCAI_OK; }
  ORCAI_FAIL(c, ORCAI_ERR_ARG, "unknown option '%s'", key);
}

int orcai_debug_read(orcai_ctx* c, float* out_host, int64_t capacity, int64_t* dims_out) {
  if (!c || !out_host || !dims_out) return ORCAI_ERR_ARG;
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  return net_debug_read(c, out_host, capacity, dims_out);
}

int orcai_load_weights(orcai_ctx* c, const char* const* names, const float* const* data, const int64_t* sizes, int32_t n) {
  if (!c || !names || !data || !sizes || n <= 0) return ORCAI_ERR_ARG;
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  return net_load_weights(c, names, data, sizes, n);
}

int orcai_upload_pcm(orcai_ctx* c, const void* pcm_host, int32_t dtype, int64_t n_samples) {
  if (!c) return ORCAI_ERR_ARG;
  if (dtype != ORCAI_PCM_I16 && dtype != ORCAI_PCM_F32) ORCAI_FAIL(c, ORCAI_ERR_ARG, "unknown pcm dtype %d", dtype);
  if (n_samples < 0 || (n_samples > 0 && !pcm_host)) ORCAI_FAIL(c, ORCAI_ERR_ARG, "bad pcm buffer");
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  const size_t esz = dtype == ORCAI_PCM_I16 ? 2 : 4;
  // +16 B slack keeps vector loads of the last interior frame inside the allocation
  ORCAI_CHECK(ensure_device_buffer(c, &c->d_pcm, &c->pcm_cap, (size_t)n_samples * esz + 16));
  ORCAI_CHECK(rec(c, EV_START));
  if (n_samples) ORCAI_CUDA(c, cudaMemcpyAsync(c->d_pcm, pcm_host, (size_t)n_samples * esz, cudaMemcpyHostToDevice, c->stream));
  ORCAI_CHECK(rec(c, EV_H2D));
  ORCAI_CUDA(c, cudaStreamSynchronize(c->stream));
  c->tm.h2d_ms = elapsed(c, EV_START, EV_H2D);
  c->pcm_dtype = dtype;
  c->n_samples = n_samples;
  c->have_stats = false;
  c->T = 0;
  return ORCAI_OK;
}

int orcai_prefetch_pcm(orcai_ctx* c, const void* pcm_host, int32_t dtype, int64_t n_samples) {
  if (!c) return ORCAI_ERR_ARG;
  if (dtype != ORCAI_PCM_I16 && dtype != ORCAI_PCM_F32) ORCAI_FAIL(c, ORCAI_ERR_ARG, "unknown pcm dtype %d", dtype);
  if (n_samples < 0 || (n_samples > 0 && !pcm_host)) ORCAI_FAIL(c, ORCAI_ERR_ARG, "bad pcm buffer");
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  const size_t esz = dtype == ORCAI_PCM_I16 ? 2 : 4;
  // a pending prefetch into this buffer must have landed before it is reallocated or overwritten
  ORCAI_CUDA(c, cudaStreamSynchronize(c->copy_stream));
  ORCAI_CHECK(ensure_device_buffer(c, &c->d_pcm_next, &c->pcm_next_cap, (size_t)n_samples * esz + 16));
  if (n_samples) ORCAI_CUDA(c, cudaMemcpyAsync(c->d_pcm_next, pcm_host, (size_t)n_samples * esz, cudaMemcpyHostToDevice, c->copy_stream));
  ORCAI_CUDA(c, cudaEventRecord(c->ev_prefetch, c->copy_stream));
  c->pcm_next_dtype = dtype;
  c->n_samples_next = n_samples;
  return ORCAI_OK;
}

int orcai_swap_pcm(orcai_ctx* c) {
  if (!c) return ORCAI_ERR_ARG;
  if (c->n_samples_next < 0) ORCAI_FAIL(c, ORCAI_ERR_STATE, "no prefetched recording (orcai_prefetch_pcm)");
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  // the previous recording's kernels have finished (every predict call synchronises), so its buffer may be recycled
  ORCAI_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_prefetch, 0));
  std::swap(c->d_pcm, c->d_pcm_next);
  std::swap(c->pcm_cap, c->pcm_next_cap);
  c->pcm_dtype = c->pcm_next_dtype;
  c->n_samples = c->n_samples_next;
  c->n_samples_next = -1;
  c->have_stats = false;
  c->T = 0;
  return ORCAI_OK;
}

int orcai_spectrogram_resident(orcai_ctx* c, int32_t normalise, orcai_spec_stats* stats) {
  if (!c) return ORCAI_ERR_ARG;
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  ORCAI_CHECK(rec(c, EV_H2D));
  ORCAI_CHECK(spectrogram_stages(c, normalise));
  ORCAI_CHECK(fill_stats(c, stats));
  ORCAI_CUDA(c, cudaStreamSynchronize(c->stream));
  c->tm.stft_ms = elapsed(c, EV_H2D, EV_STFT);
  c->tm.select_ms = elapsed(c, EV_STFT, EV_SELECT);
  c->tm.normalise_ms = elapsed(c, EV_SELECT, EV_NORM);
  c->tm.total_ms = elapsed(c, EV_H2D, EV_NORM);
  return ORCAI_OK;
}

int orcai_read_spectrogram(orcai_ctx* c, int64_t row0, int64_t nrows, float* out_host) {
  if (!c || !out_host) return ORCAI_ERR_ARG;
  if (!c->have_stats || !c->d_spec) ORCAI_FAIL(c, ORCAI_ERR_STATE, "no normalised spectrogram resident");
  if (row0 < 0 || nrows < 0 || row0 + nrows > c->T) ORCAI_FAIL(c, ORCAI_ERR_ARG, "row range outside the spectrogram");
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  const size_t w = (size_t)c->p.n_freq;
  ORCAI_CUDA(c, cudaMemcpyAsync(out_host, c->d_spec + (size_t)row0 * w, (size_t)nrows * w * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  ORCAI_CUDA(c, cudaStreamSynchronize(c->stream));
  return ORCAI_OK;
}

int orcai_read_db(orcai_ctx* c, int64_t row0, int64_t nrows, float* out_host) {
  if (!c || !out_host) return ORCAI_ERR_ARG;
  if (!c->have_stats) ORCAI_FAIL(c, ORCAI_ERR_STATE, "no spectrogram resident");
  if (row0 < 0 || nrows < 0 || row0 + nrows > c->T) ORCAI_FAIL(c, ORCAI_ERR_ARG, "row range outside the spectrogram");
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  const size_t w = (size_t)c->p.n_freq;
  float* tmp = nullptr;
  ORCAI_CUDA(c, cudaMalloc(&tmp, (size_t)std::max<int64_t>(nrows, 1) * w * sizeof(float)));
  int rc = launch_read_db(c, c->d_raw + (size_t)row0 * kRawLd, nrows, tmp);
  if (rc == ORCAI_OK) {
    cudaError_t e = cudaMemcpyAsync(out_host, tmp, (size_t)nrows * w * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { c->err = cudaGetErrorString(e); rc = ORCAI_ERR_CUDA; }
  }
  cudaFree(tmp);
  return rc;
}

int orcai_spectrogram(orcai_ctx* c, const void* pcm_host, int32_t dtype, int64_t n_samples, float* spec_out_host,
                      orcai_spec_stats* stats) {
  if (!c) return ORCAI_ERR_ARG;
  ORCAI_CHECK(orcai_upload_pcm(c, pcm_host, dtype, n_samples));
  ORCAI_CHECK(orcai_spectrogram_resident(c, 1, stats));
  if (spec_out_host) {
    ORCAI_CHECK(rec(c, EV_POST));
    ORCAI_CUDA(c, cudaMemcpyAsync(spec_out_host, c->d_spec, (size_t)c->T * c->p.n_freq * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    ORCAI_CHECK(rec(c, EV_END));
    ORCAI_CUDA(c, cudaStreamSynchronize(c->stream));
    c->tm.d2h_ms = elapsed(c, EV_POST, EV_END);
  }
  return ORCAI_OK;
}

int orcai_forward_host(orcai_ctx* c, const float* snippets_host, int64_t n, float* preds_out_host) {
  if (!c || (n > 0 && (!snippets_host || !preds_out_host))) return ORCAI_ERR_ARG;
  if (n <= 0) ORCAI_FAIL(c, ORCAI_ERR_TOO_SHORT, "empty snippet batch");
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  const size_t per = (size_t)c->p.snippet_len * c->p.n_freq;
  const int Tn = c->p.snippet_len >> c->p.n_blocks;
  ORCAI_CHECK(ensure_preds(c, n));
  // stage the materialised snippets through the spectrogram buffer
  {
    void* p = c->d_spec;
    ORCAI_CHECK(ensure_device_buffer(c, &p, &c->spec_cap, (size_t)n * per * sizeof(float)));
    c->d_spec = static_cast<float*>(p);
    c->have_stats = false;  // the resident normalised spectrogram (if any) is gone
  }
  ORCAI_CHECK(rec(c, EV_START));
  ORCAI_CUDA(c, cudaMemcpyAsync(c->d_spec, snippets_host, (size_t)n * per * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  ORCAI_CHECK(rec(c, EV_H2D));
  ORCAI_CHECK(net_forward(c, c->d_spec, 1, 0, n, c->d_preds));
  ORCAI_CHECK(rec(c, EV_NET));
  ORCAI_CUDA(c, cudaMemcpyAsync(preds_out_host, c->d_preds, (size_t)n * Tn * c->p.n_labels * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  ORCAI_CHECK(rec(c, EV_END));
  ORCAI_CUDA(c, cudaStreamSynchronize(c->stream));
  c->tm.h2d_ms = elapsed(c, EV_START, EV_H2D);
  c->tm.network_ms = elapsed(c, EV_H2D, EV_NET);
  net_collect_stage_times(c);
  c->tm.d2h_ms = elapsed(c, EV_NET, EV_END);
  c->tm.total_ms = elapsed(c, EV_START, EV_END);
  return ORCAI_OK;
}

int orcai_forward_resident(orcai_ctx* c, int64_t first, int64_t n, float* preds_out_host) {
  if (!c) return ORCAI_ERR_ARG;
  if (!c->have_stats) ORCAI_FAIL(c, ORCAI_ERR_STATE, "no spectrogram resident (orcai_spectrogram_resident)");
  const int64_t N = orcai_num_snippets(c->T, c->p.snippet_len);
  if (first < 0 || n <= 0 || first + n > N) ORCAI_FAIL(c, ORCAI_ERR_ARG, "snippet range [%lld, %lld) outside [0, %lld)", (long long)first, (long long)(first + n), (long long)N);
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  const int Tn = c->p.snippet_len >> c->p.n_blocks;
  ORCAI_CHECK(ensure_preds(c, n));
  ORCAI_CHECK(rec(c, EV_NORM));
  ORCAI_CHECK(net_forward(c, c->d_raw, 0, first, n, c->d_preds));
  ORCAI_CHECK(rec(c, EV_NET));
  if (preds_out_host)
    ORCAI_CUDA(c, cudaMemcpyAsync(preds_out_host, c->d_preds, (size_t)n * Tn * c->p.n_labels * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  ORCAI_CUDA(c, cudaStreamSynchronize(c->stream));
  c->tm.network_ms = elapsed(c, EV_NORM, EV_NET);
  net_collect_stage_times(c);
  return ORCAI_OK;
}

int orcai_postprocess(orcai_ctx* c, const float* preds_host, int64_t n_snippets, int64_t n_frames, double threshold,
                      double* agg_out, double* cnt_out, int32_t* seg_label, int64_t* seg_start, int64_t* seg_stop,
                      int64_t seg_capacity, int64_t* n_segments) {
  if (!c || !preds_host || !n_segments || seg_capacity < 0) return ORCAI_ERR_ARG;
  if (n_snippets <= 0) ORCAI_FAIL(c, ORCAI_ERR_TOO_SHORT, "no snippets");
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  const int Tn = c->p.snippet_len >> c->p.n_blocks;
  ORCAI_CHECK(ensure_preds(c, n_snippets));
  ORCAI_CUDA(c, cudaMemcpyAsync(c->d_preds, preds_host, (size_t)n_snippets * Tn * c->p.n_labels * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  ORCAI_CHECK(rec(c, EV_NET));
  int rc = launch_postprocess(c, c->d_preds, n_snippets, n_frames, threshold, agg_out, cnt_out, seg_label, seg_start, seg_stop,
                              seg_capacity, n_segments);
  cudaEventRecord(c->ev[EV_POST], c->stream);
  cudaStreamSynchronize(c->stream);
  c->tm.post_ms = elapsed(c, EV_NET, EV_POST);
  return rc;
}

int orcai_threshold_segments(orcai_ctx* c, const double* agg_host, const double* cnt_host, int64_t n_steps, int32_t n_labels,
                             double threshold, int32_t* seg_label, int64_t* seg_start, int64_t* seg_stop,
                             int64_t seg_capacity, int64_t* n_segments) {
  if (!c || !agg_host || !cnt_host || !n_segments || seg_capacity < 0) return ORCAI_ERR_ARG;
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  return launch_threshold_segments(c, agg_host, cnt_host, n_steps, n_labels, threshold, seg_label, seg_start, seg_stop,
                                   seg_capacity, n_segments);
}

int orcai_predict_resident(orcai_ctx* c, double threshold, orcai_spec_stats* stats, double* agg_out, double* cnt_out,
                           int32_t* seg_label, int64_t* seg_start, int64_t* seg_stop, int64_t seg_capacity,
                           int64_t* n_segments) {
  if (!c || !n_segments || seg_capacity < 0) return ORCAI_ERR_ARG;
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  ORCAI_CHECK(rec(c, EV_H2D));
  ORCAI_CHECK(spectrogram_stages(c, 0));
  const int64_t N = orcai_num_snippets(c->T, c->p.snippet_len);
  if (N <= 0) ORCAI_FAIL(c, ORCAI_ERR_TOO_SHORT, "recording has %lld frames, shorter than one snippet of %d", (long long)c->T, c->p.snippet_len);
  ORCAI_CHECK(ensure_preds(c, N));
  ORCAI_CHECK(net_forward(c, c->d_raw, 0, 0, N, c->d_preds));
  ORCAI_CHECK(rec(c, EV_NET));
  if (stats) ORCAI_CHECK(stats_enqueue(c));
  int rc = launch_postprocess(c, c->d_preds, N, c->T, threshold, agg_out, cnt_out, seg_label, seg_start, seg_stop,
                              seg_capacity, n_segments);
  cudaEventRecord(c->ev[EV_POST], c->stream);
  cudaStreamSynchronize(c->stream);
  if (rc != ORCAI_OK && rc != ORCAI_ERR_CAPACITY) return rc;
  stats_finish(c, stats);
  const int rc2 = ORCAI_OK;
  c->tm.stft_ms = elapsed(c, EV_H2D, EV_STFT);
  c->tm.select_ms = elapsed(c, EV_STFT, EV_SELECT);
  c->tm.normalise_ms = elapsed(c, EV_SELECT, EV_NORM);
  c->tm.network_ms = elapsed(c, EV_NORM, EV_NET);
  net_collect_stage_times(c);
  c->tm.post_ms = elapsed(c, EV_NET, EV_POST);
  c->tm.total_ms = elapsed(c, EV_H2D, EV_POST);
  return rc != ORCAI_OK ? rc : rc2;
}

int orcai_predict_resident_begin(orcai_ctx* c, double threshold, int32_t want_agg, int64_t seg_capacity) {
  if (!c || seg_capacity < 0) return ORCAI_ERR_ARG;
  if (c->async_pending >= kAsyncDepth) ORCAI_FAIL(c, ORCAI_ERR_STATE, "%d predict calls already in flight (orcai_predict_resident_end)", kAsyncDepth);
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  AsyncSlot* s = &c->slot[(c->async_head + c->async_pending) % kAsyncDepth];
  if (!s->done) ORCAI_CUDA(c, cudaEventCreateWithFlags(&s->done, cudaEventBlockingSync | cudaEventDisableTiming));
  ORCAI_CHECK(spectrogram_stages(c, 0));
  const int64_t N = orcai_num_snippets(c->T, c->p.snippet_len);
  if (N <= 0) ORCAI_FAIL(c, ORCAI_ERR_TOO_SHORT, "recording has %lld frames, shorter than one snippet of %d", (long long)c->T, c->p.snippet_len);
  ORCAI_CHECK(ensure_preds(c, N));
  ORCAI_CHECK(net_forward(c, c->d_raw, 0, 0, N, c->d_preds));
  ORCAI_CHECK(postprocess_begin(c, c->d_preds, N, c->T, threshold, want_agg != 0, seg_capacity, s));
  ORCAI_CUDA(c, cudaMemcpyAsync(static_cast<unsigned char*>(s->h_pin) + kStageStatsOff,
                                reinterpret_cast<const unsigned char*>(c->d_sel) + offsetof(SelectState, rank), sizeof(StatsTail),
                                cudaMemcpyDeviceToHost, c->stream));
  ORCAI_CUDA(c, cudaEventRecord(s->done, c->stream));
  s->busy = true;
  c->async_pending++;
  return ORCAI_OK;
}

int orcai_predict_resident_end(orcai_ctx* c, orcai_spec_stats* stats, double* agg_out, double* cnt_out, int32_t* seg_label,
                               int64_t* seg_start, int64_t* seg_stop, int64_t seg_capacity, int64_t* n_segments) {
  if (!c || !n_segments || seg_capacity < 0) return ORCAI_ERR_ARG;
  if (c->async_pending <= 0) ORCAI_FAIL(c, ORCAI_ERR_STATE, "no predict call in flight (orcai_predict_resident_begin)");
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  AsyncSlot* s = &c->slot[c->async_head];
  // the slot is released whatever happens below: a failed collection must not wedge the ring
  c->async_head = (c->async_head + 1) % kAsyncDepth;
  c->async_pending--;
  s->busy = false;
  ORCAI_CUDA(c, cudaEventSynchronize(s->done));
  if (stats) {
    StatsTail t;
    memcpy(&t, static_cast<const unsigned char*>(s->h_pin) + kStageStatsOff, sizeof t);
    const int nb = c->p.band_hi - c->p.band_lo;
    const unsigned long long n = (unsigned long long)s->T * (unsigned long long)nb;
    stats->n_frames = s->T;
    memcpy(&stats->ref_power, &t.pmax_bits, 4);
    stats->db_ref = t.db_ref;
    stats->lo = t.lo;
    stats->hi = t.hi;
    stats->rank_lo = (int64_t)nearbyint((double)(n - 1) * c->p.q_lo);
    stats->rank_hi = (int64_t)nearbyint((double)(n - 1) * c->p.q_hi);
  }
  return postprocess_end(c, s, agg_out, cnt_out, seg_label, seg_start, seg_stop, seg_capacity, n_segments);
}

int orcai_predict_in_flight(const orcai_ctx* c) { return c ? c->async_pending : 0; }

/* ---- time chunks of ONE recording (orcai_b200/timesplit.py) ---------------------------------------- */
int orcai_chunk_spectrogram(orcai_ctx* c, int64_t stat_row0, int64_t stat_row1, float* max_power_out) {
  if (!c || !max_power_out) return ORCAI_ERR_ARG;
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  if (!c->d_pcm || c->n_samples < 0) ORCAI_FAIL(c, ORCAI_ERR_STATE, "no recording uploaded (orcai_upload_pcm)");
  c->T = orcai_num_frames(c->n_samples, c->p.hop);
  if (stat_row0 < 0 || stat_row1 > c->T || stat_row0 > stat_row1) ORCAI_FAIL(c, ORCAI_ERR_ARG, "statistic rows [%lld, %lld) outside the chunk's %lld frames", (long long)stat_row0, (long long)stat_row1, (long long)c->T);
  {
    void* p = c->d_raw;
    ORCAI_CHECK(ensure_device_buffer(c, &p, &c->raw_cap, (size_t)c->T * kRawLd * sizeof(float)));
    c->d_raw = static_cast<float*>(p);
  }
  c->have_stats = false;
  ORCAI_CHECK(launch_stft(c, c->d_pcm, c->pcm_dtype, c->n_samples, c->T, c->d_raw, stat_row0, stat_row1));
  if (!c->h_small) ORCAI_CUDA(c, cudaMallocHost(&c->h_small, 256));
  ORCAI_CUDA(c, cudaMemcpyAsync(c->h_small, &c->d_sel->pmax_bits, 4, cudaMemcpyDeviceToHost, c->stream));
  ORCAI_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(max_power_out, c->h_small, 4);
  return ORCAI_OK;
}

int orcai_chunk_select_begin(orcai_ctx* c, float max_power) {
  if (!c || !(max_power >= 0.f)) return ORCAI_ERR_ARG;
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  if (c->T <= 0 || !c->d_raw) ORCAI_FAIL(c, ORCAI_ERR_STATE, "no chunk spectrogram resident (orcai_chunk_spectrogram)");
  return launch_select_begin(c, max_power);
}

int orcai_chunk_histogram(orcai_ctx* c, int32_t pass, int64_t row0, int64_t row1, const uint32_t* prefix, uint64_t* hist_out) {
  if (!c || !prefix || !hist_out || pass < 0 || pass > 2) return ORCAI_ERR_ARG;
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  if (c->T <= 0 || !c->d_raw) ORCAI_FAIL(c, ORCAI_ERR_STATE, "no chunk spectrogram resident (orcai_chunk_spectrogram)");
  if (row0 < 0 || row1 > c->T || row0 > row1) ORCAI_FAIL(c, ORCAI_ERR_ARG, "rows [%lld, %lld) outside the chunk's %lld frames", (long long)row0, (long long)row1, (long long)c->T);
  return launch_select_histogram(c, c->d_raw + (size_t)row0 * kRawLd, row1 - row0, pass, prefix, hist_out);
}

int orcai_chunk_select_end(orcai_ctx* c, const uint32_t* keys, orcai_spec_stats* stats) {
  if (!c || !keys) return ORCAI_ERR_ARG;
  ORCAI_CUDA(c, cudaSetDevice(c->device));
  if (c->T <= 0 || !c->d_raw) ORCAI_FAIL(c, ORCAI_ERR_STATE, "no chunk spectrogram resident (orcai_chunk_spectrogram)");
  ORCAI_CHECK(launch_select_end(c, keys));
  c->have_stats = true;
  return fill_stats(c, stats);
}

int orcai_predict_pcm(orcai_ctx* c, const void* pcm_host, int32_t dtype, int64_t n_samples, double threshold,
                      orcai_spec_stats* stats, double* agg_out, double* cnt_out, int32_t* seg_label, int64_t* seg_start,
                      int64_t* seg_stop, int64_t seg_capacity, int64_t* n_segments) {
  if (!c) return ORCAI_ERR_ARG;
  ORCAI_CHECK(orcai_upload_pcm(c, pcm_host, dtype, n_samples));
  return orcai_predict_resident(c, threshold, stats, agg_out, cnt_out, seg_label, seg_start, seg_stop, seg_capacity, n_segments);
}

}  // extern "C"
