/* liborcai_b200 - C ABI of the B200-native orcAI prediction hot path.
 *
 * The reference (ethz-tb/orcAI v1.0.3) is pure Python and has no FFI; its boundary for this
 * path is the Python function surface of src/orcAI/spectrogram.py and src/orcAI/predict.py.
 * Each entry point below is what a ctypes binding placed under one of those functions calls
 * (INTEGRATION.md shows the stubs).  Conventions: plain pointers and sizes, `int` status
 * (0 = ok, negative = error, text via orcai_last_error), caller-allocated outputs, no C++
 * exceptions cross the ABI, one context per (process, device); calls on one context are not
 * thread-safe, distinct contexts are independent.  There is no CPU fallback: every entry
 * point that computes needs a CUDA device (sm_100a build).
 */
#ifndef ORCAI_B200_H
#define ORCAI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORCAI_OK 0
#define ORCAI_ERR_ARG (-1)         /* bad argument / unsupported parameter set */
#define ORCAI_ERR_CUDA (-2)        /* CUDA runtime error, see orcai_last_error */
#define ORCAI_ERR_STATE (-3)       /* call order (no recording loaded, no weights, ...) */
#define ORCAI_ERR_CAPACITY (-4)    /* caller buffer too small; required size reported */
#define ORCAI_ERR_TOO_SHORT (-5)   /* recording shorter than one snippet (reference: Keras raises on empty batch) */

#define ORCAI_PCM_I16 0            /* raw PCM16; the kernel applies x/32768 (libsndfile/librosa.load semantics) */
#define ORCAI_PCM_F32 1            /* float32 in [-1,1], what librosa.load hands to stft (spectrogram.py:23-35) */

typedef struct orcai_ctx orcai_ctx;

/* Parameter block = the fields of orcai_parameter.json / model_shape.json the path consumes
 * (reference: models/orcai-V1/orcai_parameter.json, model_shape.json). */
typedef struct {
  int32_t sampling_rate;  /* 48000 */
  int32_t n_fft;          /* 512  (kernel K1 is specialised to 512) */
  int32_t hop;            /* 256  ("n_overlap" in the JSON, used as hop length: spectrogram.py:37) */
  int32_t band_lo;        /* first kept rFFT bin  (0)   spectrogram.py:62-68 */
  int32_t band_hi;        /* one past last kept   (171) */
  double q_lo;            /* quantiles as np.percentile sees them: true_divide(100*q, 100) */
  double q_hi;
  int32_t snippet_len;    /* 736  model_shape.json input_shape[0] */
  int32_t n_freq;         /* 171  model_shape.json input_shape[1] == band_hi - band_lo */
  int32_t n_labels;       /* 7 */
  int32_t n_blocks;       /* len(filters) = 4 */
  int32_t filters[8];     /* 30,40,50,60 */
  int32_t kernel_size;    /* 3 */
  int32_t lstm_units;     /* 128 */
  int32_t reserved[8];
} orcai_params;

typedef struct {
  int64_t n_frames;       /* T = 1 + n_samples / hop */
  float ref_power;        /* max |S|^2 over all 257 bins and all frames (amplitude_to_db ref=np.max, squared) */
  float db_ref;           /* 10*log10(max(1e-10, ref_power)) */
  float lo;               /* nearest-rank q_lo percentile of the cropped, shifted, floored dB array */
  float hi;               /* nearest-rank q_hi percentile */
  int64_t rank_lo;        /* the two ranks (indices into the sorted flattened array) */
  int64_t rank_hi;
} orcai_spec_stats;

/* Device times (CUDA events on the context's stream) of the stages of the last call, in ms,
 * and the number of kernels this library launched since the context was created. */
typedef struct {
  float h2d_ms;
  float stft_ms;
  float select_ms;
  float normalise_ms;
  float network_ms;
  float post_ms;
  float d2h_ms;
  float total_ms;
  uint64_t kernel_launches;
  float net_stage_ms[16]; /* per network stage of the last forward chunk sequence */
} orcai_timings;

/* ---- context ------------------------------------------------------------------------------ */
int orcai_create(int device, const orcai_params* params, orcai_ctx** out);
void orcai_destroy(orcai_ctx* ctx);
const char* orcai_last_error(const orcai_ctx* ctx);   /* ctx may be NULL: error of the last failed orcai_create */
int orcai_get_timings(const orcai_ctx* ctx, orcai_timings* out);
int orcai_version(void);

/* Page-locked host memory for recordings (no context needed; usable from any thread, on any device of the process).
 * A recording decoded straight into such a buffer reaches the device by DMA at PCIe speed; from ordinary (pageable) memory
 * the upload of a 1-h PCM16 recording costs 31 ms instead of 6 (reference: librosa.load returns pageable numpy memory,
 * spectrogram.py:23-27).  Returns NULL when the allocation fails. */
void* orcai_host_alloc(size_t bytes);
void orcai_host_free(void* p);

/* ---- weights: replaces keras.saving.load_model inside load_orcai_model (io.py:386-392) ----- */
/* n named float32 tensors in Keras variable layouts (names: orcai_b200/weights.py). BatchNorm
 * folding and the device layouts are produced inside. */
int orcai_load_weights(orcai_ctx* ctx, const char* const* names, const float* const* data,
                       const int64_t* sizes, int32_t n);

/* ---- spectrogram stage: replaces calculate_spectrogram + preprocess_spectrogram
 *      (spectrogram.py:15-87; librosa.stft / amplitude_to_db / np.percentile call sites) ------ */
int64_t orcai_num_frames(int64_t n_samples, int32_t hop);
/* Upload mono PCM (host memory) and keep it resident as the context's current recording. */
int orcai_upload_pcm(orcai_ctx* ctx, const void* pcm_host, int32_t dtype, int64_t n_samples);
/* Pipelining across recordings (recording tables: predict.py:733-755 loops over rows): start the host->device copy of
 * the NEXT recording into a second device buffer on a copy stream and return at once (pin `pcm_host` for a truly
 * asynchronous copy; it must stay valid until orcai_swap_pcm).  orcai_swap_pcm makes that recording the resident one
 * (the compute stream waits for the copy, the host does not).  Usage: prefetch(k+1); predict_resident(k); swap. */
int orcai_prefetch_pcm(orcai_ctx* ctx, const void* pcm_host, int32_t dtype, int64_t n_samples);
int orcai_swap_pcm(orcai_ctx* ctx);
/* STFT -> dB -> crop -> global max -> exact percentiles on the resident recording.  With
 * `normalise` != 0 also materialises the normalised (T, n_freq) float32 spectrogram on the device. */
int orcai_spectrogram_resident(orcai_ctx* ctx, int32_t normalise, orcai_spec_stats* stats);
/* One call = upload + orcai_spectrogram_resident(normalise=1) + copy-out of the (T, n_freq) result. */
int orcai_spectrogram(orcai_ctx* ctx, const void* pcm_host, int32_t dtype, int64_t n_samples,
                      float* spec_out_host, orcai_spec_stats* stats);
/* Copy rows [row0, row0+nrows) of the normalised spectrogram / of the cropped dB array to host. */
int orcai_read_spectrogram(orcai_ctx* ctx, int64_t row0, int64_t nrows, float* out_host);
int orcai_read_db(orcai_ctx* ctx, int64_t row0, int64_t nrows, float* out_host);

/* ---- snippet batcher + orcai-V1 forward: replaces the snippet copy and model.predict
 *      (predict.py:252-268; architectures.py:120-241) ----------------------------------------- */
int64_t orcai_num_snippets(int64_t n_frames, int32_t snippet_len);
/* model.predict boundary: x (n, snippet_len, n_freq) float32 host -> (n, snippet_len/2^n_blocks, n_labels). */
int orcai_forward_host(orcai_ctx* ctx, const float* snippets_host, int64_t n, float* preds_out_host);
/* Snippets [first, first+n) cut as strided windows from the resident recording (no copy). */
int orcai_forward_resident(orcai_ctx* ctx, int64_t first, int64_t n, float* preds_out_host);

/* ---- post-processing: replaces predict.py:276-340 and auxiliary.py:420-440 ----------------- */
/* preds (n_snippets, P, L) float32 host -> overlap-averaged float64 (T/2^n_blocks, L) + counts,
 * thresholded at threshold/max(count) (strict >), run-length segments in label-major order
 * (label index, first step, inclusive last step).  agg_out / cnt_out may be NULL. */
int orcai_postprocess(orcai_ctx* ctx, const float* preds_host, int64_t n_snippets, int64_t n_frames,
                      double threshold, double* agg_out, double* cnt_out, int32_t* seg_label,
                      int64_t* seg_start, int64_t* seg_stop, int64_t seg_capacity, int64_t* n_segments);
/* compute_binary_predictions boundary (predict.py:298-317) on caller-supplied aggregates. */
int orcai_threshold_segments(orcai_ctx* ctx, const double* agg_host, const double* cnt_host,
                             int64_t n_steps, int32_t n_labels, double threshold, int32_t* seg_label,
                             int64_t* seg_start, int64_t* seg_stop, int64_t seg_capacity,
                             int64_t* n_segments);

/* ---- fused, device-resident predict_wav core (predict.py:426-451) --------------------------- */
/* Runs spectrogram -> snippets -> forward -> overlap-average -> threshold -> segments on the
 * resident recording without leaving the device; only segments (and optionally the aggregates)
 * are copied back. */
int orcai_predict_resident(orcai_ctx* ctx, double threshold, orcai_spec_stats* stats,
                           double* agg_out, double* cnt_out, int32_t* seg_label, int64_t* seg_start,
                           int64_t* seg_stop, int64_t seg_capacity, int64_t* n_segments);
/* The same call in two halves, for a caller that annotates recording after recording (the table loop, predict.py:733-755):
 * _begin enqueues everything for the resident recording - kernels and the read-back of the results into page-locked staging
 * owned by the context - and returns at once; _end waits (the host thread sleeps) for the OLDEST begun call and hands out its
 * results.  Up to two calls may be in flight, so recording k+1 is queued on the device before the host has collected recording
 * k and the device never idles while the host thread builds label tables or waits for the interpreter lock.  Between _begin and
 * its _end only orcai_prefetch_pcm, orcai_swap_pcm and this pair may be called on the context; the stage timings of
 * orcai_get_timings are not updated.  `want_agg` != 0 makes the aggregates available to _end (agg_out / cnt_out may then be
 * non-null).  _end returns ORCAI_ERR_CAPACITY (n_segments = the count needed) when either capacity was too small: the
 * recording has to be annotated again with a larger one. */
int orcai_predict_resident_begin(orcai_ctx* ctx, double threshold, int32_t want_agg, int64_t seg_capacity);
int orcai_predict_resident_end(orcai_ctx* ctx, orcai_spec_stats* stats, double* agg_out, double* cnt_out,
                               int32_t* seg_label, int64_t* seg_start, int64_t* seg_stop,
                               int64_t seg_capacity, int64_t* n_segments);
int orcai_predict_in_flight(const orcai_ctx* ctx);   /* begun and not yet collected calls (0..2) */
/* upload + orcai_predict_resident. */
int orcai_predict_pcm(orcai_ctx* ctx, const void* pcm_host, int32_t dtype, int64_t n_samples,
                      double threshold, orcai_spec_stats* stats, double* agg_out, double* cnt_out,
                      int32_t* seg_label, int64_t* seg_start, int64_t* seg_stop,
                      int64_t seg_capacity, int64_t* n_segments);

/* ---- calibration of the 16-bit tensor-core paths ------------------------------------------------ */
/* fp16 weight rounding is identical at every pixel, so (to first order) it shifts each layer's output by the constant
 * sum_k dW[k][n] * mean(A[k]).  This call measures the channel means of every GEMM input on the first `max_snippets`
 * (<= 0: 8) snippets of the RESIDENT recording (after orcai_spectrogram_resident) and folds the correction into the bias
 * rows of the tensor-core operands (they are split fp16 hi+lo, i.e. exact).  Deterministic; the fp32 path is not affected;
 * orcai_load_weights clears it.  The Python layer calibrates on a built-in synthetic recording when weights are loaded. */
int orcai_calibrate(orcai_ctx* ctx, int64_t max_snippets);

/* ---- one recording split by time across several contexts / GPUs ------------------------------------
 * The reference computes ONE reference level (ref = np.max over all 257 bins and all frames, spectrogram.py:51-53) and ONE
 * pair of percentiles (spectrogram.py:70-75) per recording, so chunks of a recording must agree on them: each context
 * holds a chunk (with halo frames) and the host combines the chunks' partial statistics between these calls
 * (orcai_b200/timesplit.py): max of the maxima, sum of the radix-select histograms.  Rows are frame indices of the chunk.
 *   orcai_chunk_spectrogram  STFT -> dB of the uploaded chunk; *max_power_out = max |S|^2 over rows [stat_row0, stat_row1)
 *   orcai_chunk_select_begin the recording-wide maximum -> reference level (computed on the device, like the one-GPU path)
 *   orcai_chunk_histogram    pass 0/1/2 of the exact radix select over rows [row0, row1): hist_out[2][2048] counts of the
 *                            11/11/10-bit digit, restricted to keys that start with prefix[r] (pass 0: one histogram)
 *   orcai_chunk_select_end   the decided 32-bit keys -> lo / hi; the chunk is then ready for orcai_forward_resident
 * Results are bit-identical to the one-context path (tests/test_timesplit.py). */
int orcai_chunk_spectrogram(orcai_ctx* ctx, int64_t stat_row0, int64_t stat_row1, float* max_power_out);
int orcai_chunk_select_begin(orcai_ctx* ctx, float max_power);
int orcai_chunk_histogram(orcai_ctx* ctx, int32_t pass, int64_t row0, int64_t row1, const uint32_t* prefix, uint64_t* hist_out);
int orcai_chunk_select_end(orcai_ctx* ctx, const uint32_t* keys, orcai_spec_stats* stats);

/* ---- knobs ---------------------------------------------------------------------------------- */
/* Options: "net_path"  4 = split-fp16 tcgen05 path at fp32 grade (what orcai_b200's Python layer selects by default): every
 *                      GEMM as A_hi*W_hi + A_lo*W_hi + A_hi*W_lo with fp32 accumulation, fp32 CUDA cores elsewhere, within
 *                      2e-5 of the fp32 graph (model.predict of predict.py:266-268 is compared at 1e-3);
 *                      0 = fp32 CUDA-core path (the in-library yardstick, library default of a fresh context),
 *                      1 = fp16 / 2 = bf16 layer-wise tcgen05 path, 3 = single-fp16 fused tcgen05 path (opt-in, outside the
 *                      1e-3 gate): tensor-core entry convolution, fused residual-block kernels, tensor-core LSTM tail;
 *          "precise_tall" (net_path 4, resident recordings) 1 = trunk once over the recording as a tall image + per-snippet
 *          border rows (default; bit-identical to 0 = snippet by snippet); "precise_sep_path" (net_path 4) 1 = depthwise filter
 *          fused into the split pointwise GEMM (default), 0 = two kernels (bit-identical);
 *          "tail_path" (net_path 3) 1 = tensor-core LSTM/dense tail (default), 0 = fp32 CUDA-core tail;
 *          "conv0_path" (net_path 3) 1 = tensor-core entry convolution (default), 0 = fp32 CUDA-core entry convolution,
 *          2 = entry convolution fused into the first residual block's kernel (measured slower, kept as an option);
 *          "block1_path" (net_path 3) 0 = one MMA per tap (default), 1 = N-widened MMAs: the three dx taps as column
 *          blocks of one MMA, combined by warp shuffles in the epilogue (correct, measured slower: epilogue-bound);
 *          "stft_f64"  1 = float64 FFT (parity grade, default), 0 = float32 FFT (fast);
 *          "stft_threads" threads per frame of the float64 FFT kernel: 16 (default: 16 x 16 decomposition, 128 registers) or
 *          8 (32 x 8 decomposition, 255 registers); the two differ by float rounding only;
 *          "chunk"     snippets per network launch sequence;
 *          "debug_stop" stop the forward after a stage (see orcai_debug_read), -1 = off. */
int orcai_set_option(orcai_ctx* ctx, const char* key, int64_t value);
/* Test hook: after a forward run with "debug_stop" >= 0, copy that stage's activations of the first chunk to
 * the host as compact float32 NHWC; dims_out receives (n, h, w, c).  Returns ORCAI_ERR_CAPACITY if too small. */
int orcai_debug_read(orcai_ctx* ctx, float* out_host, int64_t capacity, int64_t* dims_out);

#ifdef __cplusplus
}
#endif
#endif /* ORCAI_B200_H */
