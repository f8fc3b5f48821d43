// K7: overlap-average + threshold + run-length segment extraction in ONE single-pass scan kernel.
//
// Replaces, of the reference:
//   src/orcAI/predict.py:276-293   float64 overlap-average of per-snippet predictions
//   src/orcAI/predict.py:298-317   threshold / max(overlap), strict ">", per-label runs
//   src/orcAI/auxiliary.py:420-440 run starts / inclusive stops
//
// The (label, step) plane is flattened label-major (f = l * S + s), which is the order in which
// compute_binary_predictions emits its lists.  Every element recomputes the mask of its two
// neighbours, derives a start flag and a stop flag, and a decoupled-look-back prefix sum over
// (starts, stops) gives every run its output slot; the k-th start and the k-th stop of the
// flattened order belong to the same run.  Integer results are bit-exact by construction; the
// float64 average adds at most `max_overlap` float32 values in ascending snippet order, exactly
// like the reference loop.
#include <algorithm>
#include <cstring>
#include "common.h"

namespace orcai {

namespace {

constexpr int kThreads = 256;
constexpr int kItems = 4;
constexpr int kTile = kThreads * kItems;

struct PostGeom {
  long long S;        // output steps = T / ds
  long long N;        // snippets
  int L;              // labels
  int P;              // prediction steps per snippet
  int shift;          // step shift between snippets
};

struct PostScratch {        // device
  unsigned long long total; // packed (starts << 32 | stops) of the whole plane
  unsigned int tile_counter;
  unsigned int pad;
  unsigned long long maxcnt_bits;  // threshold_segments: max(count) as double bits
};

// float64 overlap-average of step s, label l (predict.py:283-293); count returned through *cnt
__device__ __forceinline__ double agg_at(const float* __restrict__ preds, const PostGeom& g, long long s, int l, int* cnt) {
  // snippets i with shift*i <= s < shift*i + P
  long long i_hi = s / g.shift;
  if (i_hi > g.N - 1) i_hi = g.N - 1;
  long long i_lo = (s - g.P + g.shift) / g.shift;  // ceil((s - P + 1) / shift) for s - P + 1 > 0
  if (s - g.P + 1 <= 0) i_lo = 0;
  double acc = 0.0;
  int n = 0;
  for (long long i = i_lo; i <= i_hi; ++i) {
    acc += (double)preds[((size_t)i * g.P + (size_t)(s - i * g.shift)) * g.L + l];
    ++n;
  }
  *cnt = n;
  return n ? acc / (double)n : 0.0;
}

struct MaskFromPreds {
  const float* preds;
  PostGeom g;
  double thr;
  __device__ __forceinline__ bool operator()(long long s, int l) const {
    int n;
    const double a = agg_at(preds, g, s, l, &n);
    return a > thr;
  }
};

struct MaskFromAgg {
  const double* agg;
  int L;
  double threshold;
  const unsigned long long* maxcnt_bits;
  __device__ __forceinline__ bool operator()(long long s, int l) const {
    const double thr = threshold / __longlong_as_double((long long)*maxcnt_bits);
    return agg[(size_t)s * L + l] > thr;
  }
};

template <class Mask>
__global__ void __launch_bounds__(kThreads)
segments_scan_kernel(Mask mask, long long S, int L, unsigned long long* __restrict__ tile_state,
                     PostScratch* __restrict__ scr, int* __restrict__ seg_label,
                     long long* __restrict__ seg_start, long long* __restrict__ seg_stop, long long cap) {
  __shared__ unsigned int s_tile;
  __shared__ unsigned long long s_warp[kThreads / 32];
  __shared__ unsigned long long s_excl;
  if (threadIdx.x == 0) s_tile = atomicAdd(&scr->tile_counter, 1u);
  __syncthreads();
  const unsigned int tile = s_tile;
  const long long total = S * L;
  const long long f0 = (long long)tile * kTile + (long long)threadIdx.x * kItems;

  bool st[kItems], sp[kItems];
  unsigned long long local = 0;  // starts << 32 | stops
#pragma unroll
  for (int q = 0; q < kItems; ++q) {
    const long long f = f0 + q;
    st[q] = sp[q] = false;
    if (f < total) {
      const int l = (int)(f / S);
      const long long s = f - (long long)l * S;
      if (mask(s, l)) {
        st[q] = (s == 0) || !mask(s - 1, l);
        sp[q] = (s == S - 1) || !mask(s + 1, l);
      }
    }
    local += ((unsigned long long)st[q] << 32) + (unsigned long long)sp[q];
  }
  // block exclusive scan of `local`
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long inc = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  unsigned long long wpre = 0, block_total = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    if (w < warp) wpre += s_warp[w];
    block_total += s_warp[w];
  }
  const unsigned long long excl_in_block = wpre + inc - local;

  // decoupled look-back on packed words: [63:62] status (1 = aggregate, 2 = inclusive prefix), [61:31] starts, [30:0] stops
  if (threadIdx.x == 0) {
    const unsigned long long agg_packed = ((block_total >> 32) << 31) | (block_total & 0x7fffffffull);
    unsigned long long excl = 0;
    if (tile == 0) {
      atomicExch(&tile_state[0], (2ull << 62) | agg_packed);
    } else {
      atomicExch(&tile_state[tile], (1ull << 62) | agg_packed);
      long long p = (long long)tile - 1;
      while (true) {
        unsigned long long w;
        do { w = atomicAdd(&tile_state[p], 0ull); } while ((w >> 62) == 0ull);
        excl += w & 0x3fffffffffffffffull;  // fields cannot carry into each other (each < 2^31 in total)
        if ((w >> 62) == 2ull) break;
        --p;
      }
      atomicExch(&tile_state[tile], (2ull << 62) | (excl + agg_packed));
    }
    s_excl = excl;
    if ((long long)(tile + 1) * kTile >= total) scr->total = excl + agg_packed;
  }
  __syncthreads();
  const unsigned long long excl_tile = s_excl;
  long long start_slot = (long long)(excl_tile >> 31) + (long long)(excl_in_block >> 32);
  long long stop_slot = (long long)(excl_tile & 0x7fffffffull) + (long long)(excl_in_block & 0xffffffffull);
#pragma unroll
  for (int q = 0; q < kItems; ++q) {
    const long long f = f0 + q;
    if (f >= total) break;
    const int l = (int)(f / S);
    const long long s = f - (long long)l * S;
    if (st[q]) {
      if (start_slot < cap) { seg_start[start_slot] = s; seg_label[start_slot] = l; }
      ++start_slot;
    }
    if (sp[q]) {
      if (stop_slot < cap) seg_stop[stop_slot] = s;
      ++stop_slot;
    }
  }
}

// optional outputs of the aggregation itself (float64 average and overlap count)
__global__ void __launch_bounds__(256)
aggregate_kernel(const float* __restrict__ preds, PostGeom g, double* __restrict__ agg, double* __restrict__ cnt) {
  const long long total = g.S * g.L;
  for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < total; f += (long long)gridDim.x * blockDim.x) {
    const long long s = f / g.L;
    const int l = (int)(f - s * g.L);
    int n;
    const double a = agg_at(preds, g, s, l, &n);
    agg[f] = a;
    if (l == 0) cnt[s] = (double)n;
  }
}

__global__ void maxcount_kernel(const double* __restrict__ cnt, long long S, unsigned long long* out_bits) {
  double m = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x)
    m = fmax(m, cnt[i]);
  for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(m));  // counts are >= 0
}

struct PostBuffers {
  PostScratch* scr;
  unsigned long long* tile_state;
  int* seg_label;
  long long* seg_start;
  long long* seg_stop;
  double* agg;
  double* cnt;
  float* preds;  // staging for host-supplied predictions
};

int carve(Ctx* c, long long n_tiles, long long cap, long long S, int L, size_t preds_bytes, bool need_agg, PostBuffers* b) {
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  const size_t o_scr = off;   off += up(sizeof(PostScratch));
  const size_t o_tile = off;  off += up((size_t)n_tiles * 8);
  const size_t o_lab = off;   off += up((size_t)cap * 4);
  const size_t o_sta = off;   off += up((size_t)cap * 8);
  const size_t o_sto = off;   off += up((size_t)cap * 8);
  const size_t o_agg = off;   off += need_agg ? up((size_t)S * L * 8) : 0;
  const size_t o_cnt = off;   off += need_agg ? up((size_t)S * 8) : 0;
  const size_t o_pre = off;   off += up(preds_bytes);
  ORCAI_CHECK(ensure_device_buffer(c, &c->d_post, &c->post_cap, off));
  unsigned char* base = static_cast<unsigned char*>(c->d_post);
  b->scr = reinterpret_cast<PostScratch*>(base + o_scr);
  b->tile_state = reinterpret_cast<unsigned long long*>(base + o_tile);
  b->seg_label = reinterpret_cast<int*>(base + o_lab);
  b->seg_start = reinterpret_cast<long long*>(base + o_sta);
  b->seg_stop = reinterpret_cast<long long*>(base + o_sto);
  b->agg = need_agg ? reinterpret_cast<double*>(base + o_agg) : nullptr;
  b->cnt = need_agg ? reinterpret_cast<double*>(base + o_cnt) : nullptr;
  b->preds = reinterpret_cast<float*>(base + o_pre);
  // scratch + tile states are contiguous at the front
  ORCAI_CUDA(c, cudaMemsetAsync(base, 0, o_lab, c->stream));
  return ORCAI_OK;
}

// Results -> host with ONE stream synchronise: the segment count, the first kSpecSegs segments (a speculative copy: the count is
// not known on the host yet), and the optional aggregates all land in pinned staging; a recording with more segments pays a
// second copy for the rest.  (Three synchronises and pageable-memory copies used to cost ~0.3 ms per recording end to end.)
constexpr long long kSpecSegs = 32768;

int copy_out_results(Ctx* c, const PostBuffers& b, long long cap, int32_t* h_label, int64_t* h_start, int64_t* h_stop, int64_t* n_seg,
                     double* h_agg, double* h_cnt, long long n_agg, long long n_cnt) {
  const long long spec = std::min<long long>(cap, kSpecSegs);
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_tot = 0, o_lab = 256, o_sta = o_lab + up((size_t)spec * 4), o_sto = o_sta + up((size_t)spec * 8);
  const size_t o_agg = o_sto + up((size_t)spec * 8), o_cnt = o_agg + (h_agg ? up((size_t)n_agg * 8) : 0);
  const size_t bytes = o_cnt + (h_cnt ? up((size_t)n_cnt * 8) : 0);
  ORCAI_CHECK(ensure_pinned(c, bytes));
  unsigned char* pin = static_cast<unsigned char*>(c->h_pin);
  ORCAI_CUDA(c, cudaMemcpyAsync(pin + o_tot, &b.scr->total, 8, cudaMemcpyDeviceToHost, c->stream));
  if (spec) {
    ORCAI_CUDA(c, cudaMemcpyAsync(pin + o_lab, b.seg_label, (size_t)spec * 4, cudaMemcpyDeviceToHost, c->stream));
    ORCAI_CUDA(c, cudaMemcpyAsync(pin + o_sta, b.seg_start, (size_t)spec * 8, cudaMemcpyDeviceToHost, c->stream));
    ORCAI_CUDA(c, cudaMemcpyAsync(pin + o_sto, b.seg_stop, (size_t)spec * 8, cudaMemcpyDeviceToHost, c->stream));
  }
  if (h_agg) ORCAI_CUDA(c, cudaMemcpyAsync(pin + o_agg, b.agg, (size_t)n_agg * 8, cudaMemcpyDeviceToHost, c->stream));
  if (h_cnt) ORCAI_CUDA(c, cudaMemcpyAsync(pin + o_cnt, b.cnt, (size_t)n_cnt * 8, cudaMemcpyDeviceToHost, c->stream));
  // a spinning wait: the synchronous call returns ~0.4 ms sooner than through a blocking-sync event (measured); callers that
  // annotate recording after recording use orcai_predict_resident_begin / _end, whose wait sleeps
  ORCAI_CUDA(c, cudaStreamSynchronize(c->stream));
  unsigned long long total = 0;
  memcpy(&total, pin + o_tot, 8);
  if (h_agg) memcpy(h_agg, pin + o_agg, (size_t)n_agg * 8);
  if (h_cnt) memcpy(h_cnt, pin + o_cnt, (size_t)n_cnt * 8);
  const long long n_starts = (long long)(total >> 31), n_stops = (long long)(total & 0x7fffffffull);
  if (n_starts != n_stops) ORCAI_FAIL(c, ORCAI_ERR_STATE, "segment scan inconsistent: %lld starts vs %lld stops", n_starts, n_stops);
  *n_seg = n_starts;
  if (n_starts > cap) ORCAI_FAIL(c, ORCAI_ERR_CAPACITY, "segment capacity %lld too small, need %lld", cap, n_starts);
  const long long first = std::min(n_starts, spec);
  if (first) {
    memcpy(h_label, pin + o_lab, (size_t)first * 4);
    memcpy(h_start, pin + o_sta, (size_t)first * 8);
    memcpy(h_stop, pin + o_sto, (size_t)first * 8);
  }
  if (n_starts > spec) {
    const long long rest = n_starts - spec;
    ORCAI_CUDA(c, cudaMemcpyAsync(h_label + spec, b.seg_label + spec, (size_t)rest * 4, cudaMemcpyDeviceToHost, c->stream));
    ORCAI_CUDA(c, cudaMemcpyAsync(h_start + spec, b.seg_start + spec, (size_t)rest * 8, cudaMemcpyDeviceToHost, c->stream));
    ORCAI_CUDA(c, cudaMemcpyAsync(h_stop + spec, b.seg_stop + spec, (size_t)rest * 8, cudaMemcpyDeviceToHost, c->stream));
    ORCAI_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return ORCAI_OK;
}

int copy_out_segments(Ctx* c, const PostBuffers& b, long long cap, int32_t* h_label, int64_t* h_start, int64_t* h_stop,
                      int64_t* n_seg) {
  return copy_out_results(c, b, cap, h_label, h_start, h_stop, n_seg, nullptr, nullptr, 0, 0);
}

}  // namespace

int ensure_pinned(Ctx* c, size_t bytes) {
  if (bytes <= c->pin_cap && c->h_pin) return ORCAI_OK;
  if (c->h_pin) { ORCAI_CUDA(c, cudaFreeHost(c->h_pin)); c->h_pin = nullptr; c->pin_cap = 0; }
  bytes = std::max<size_t>(bytes + bytes / 4, (size_t)1 << 20);
  ORCAI_CUDA(c, cudaMallocHost(&c->h_pin, bytes));
  c->pin_cap = bytes;
  return ORCAI_OK;
}

// d_preds == nullptr means "predictions are at c->d_preds" is NOT assumed: callers pass the device pointer,
// or h_preds_src != nullptr to stage host predictions first.
int launch_postprocess(Ctx* c, const float* d_preds, int64_t n_snippets, int64_t T, double threshold,
                       double* h_agg, double* h_cnt, int32_t* h_label, int64_t* h_start, int64_t* h_stop,
                       int64_t cap, int64_t* n_seg) {
  const int ds = 1 << c->p.n_blocks;
  PostGeom g;
  g.S = T / ds;
  g.N = n_snippets;
  g.L = c->p.n_labels;
  g.P = c->p.snippet_len / ds;
  g.shift = (c->p.snippet_len / 2) / ds;
  if (g.N < 1) ORCAI_FAIL(c, ORCAI_ERR_TOO_SHORT, "no snippets to aggregate");
  if (g.shift < 1) ORCAI_FAIL(c, ORCAI_ERR_ARG, "snippet shift shorter than one output step");
  *n_seg = 0;
  if (g.S == 0) return ORCAI_OK;
  long long max_overlap = (g.P + g.shift - 1) / g.shift;
  if (max_overlap > g.N) max_overlap = g.N;
  const double thr = threshold / (double)max_overlap;  // threshold / np.max(overlap_count)
  const long long total = g.S * g.L;
  const long long n_tiles = (total + kTile - 1) / kTile;
  const bool need_agg = (h_agg != nullptr) || (h_cnt != nullptr);
  PostBuffers b;
  ORCAI_CHECK(carve(c, n_tiles, cap, g.S, g.L, 0, need_agg, &b));
  MaskFromPreds m{d_preds, g, thr};
  segments_scan_kernel<MaskFromPreds><<<(unsigned)n_tiles, kThreads, 0, c->stream>>>(
      m, g.S, g.L, b.tile_state, b.scr, b.seg_label, b.seg_start, b.seg_stop, cap);
  c->launches++;
  if (need_agg) {
    aggregate_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(d_preds, g, b.agg, b.cnt);
    c->launches++;
  }
  ORCAI_CUDA(c, cudaGetLastError());
  return copy_out_results(c, b, cap, h_label, h_start, h_stop, n_seg, h_agg, h_cnt, total, g.S);
}

int postprocess_begin(Ctx* c, const float* d_preds, int64_t n_snippets, int64_t T, double threshold, bool want_agg, int64_t cap, AsyncSlot* s) {
  const int ds = 1 << c->p.n_blocks;
  PostGeom g;
  g.S = T / ds;
  g.N = n_snippets;
  g.L = c->p.n_labels;
  g.P = c->p.snippet_len / ds;
  g.shift = (c->p.snippet_len / 2) / ds;
  if (g.N < 1) ORCAI_FAIL(c, ORCAI_ERR_TOO_SHORT, "no snippets to aggregate");
  if (g.shift < 1) ORCAI_FAIL(c, ORCAI_ERR_ARG, "snippet shift shorter than one output step");
  long long max_overlap = (g.P + g.shift - 1) / g.shift;
  if (max_overlap > g.N) max_overlap = g.N;
  const double thr = threshold / (double)max_overlap;
  const long long total = g.S * g.L;
  const long long n_tiles = (total + kTile - 1) / kTile;
  s->T = T; s->cap = cap; s->want_agg = want_agg; s->n_agg = total; s->n_cnt = g.S;
  s->spec = std::min<long long>(cap, kSpecSegs);
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  s->o_lab = kStageSegsOff;
  s->o_sta = s->o_lab + up((size_t)s->spec * 4);
  s->o_sto = s->o_sta + up((size_t)s->spec * 8);
  s->o_agg = s->o_sto + up((size_t)s->spec * 8);
  s->o_cnt = s->o_agg + (want_agg ? up((size_t)total * 8) : 0);
  const size_t bytes = s->o_cnt + (want_agg ? up((size_t)g.S * 8) : 0);
  if (bytes > s->pin_cap || !s->h_pin) {
    if (s->h_pin) { ORCAI_CUDA(c, cudaFreeHost(s->h_pin)); s->h_pin = nullptr; s->pin_cap = 0; }
    const size_t want = std::max<size_t>(bytes + bytes / 4, (size_t)1 << 20);
    ORCAI_CUDA(c, cudaMallocHost(&s->h_pin, want));
    s->pin_cap = want;
  }
  unsigned char* pin = static_cast<unsigned char*>(s->h_pin);
  memset(pin + kStageTotalOff, 0, 8);
  if (g.S == 0) { s->d_lab = nullptr; return ORCAI_OK; }
  PostBuffers b;
  {  // the slot's own scratch: the next recording's post-processing must not overwrite segments that are still to be read
    std::swap(c->d_post, s->d_post);
    std::swap(c->post_cap, s->post_cap);
    const int rc = carve(c, n_tiles, cap, g.S, g.L, 0, want_agg, &b);
    std::swap(c->d_post, s->d_post);
    std::swap(c->post_cap, s->post_cap);
    ORCAI_CHECK(rc);
  }
  s->d_lab = b.seg_label; s->d_sta = b.seg_start; s->d_sto = b.seg_stop;
  MaskFromPreds m{d_preds, g, thr};
  segments_scan_kernel<MaskFromPreds><<<(unsigned)n_tiles, kThreads, 0, c->stream>>>(
      m, g.S, g.L, b.tile_state, b.scr, b.seg_label, b.seg_start, b.seg_stop, cap);
  c->launches++;
  if (want_agg) {
    aggregate_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(d_preds, g, b.agg, b.cnt);
    c->launches++;
  }
  ORCAI_CUDA(c, cudaGetLastError());
  ORCAI_CUDA(c, cudaMemcpyAsync(pin + kStageTotalOff, &b.scr->total, 8, cudaMemcpyDeviceToHost, c->stream));
  if (s->spec) {
    ORCAI_CUDA(c, cudaMemcpyAsync(pin + s->o_lab, b.seg_label, (size_t)s->spec * 4, cudaMemcpyDeviceToHost, c->stream));
    ORCAI_CUDA(c, cudaMemcpyAsync(pin + s->o_sta, b.seg_start, (size_t)s->spec * 8, cudaMemcpyDeviceToHost, c->stream));
    ORCAI_CUDA(c, cudaMemcpyAsync(pin + s->o_sto, b.seg_stop, (size_t)s->spec * 8, cudaMemcpyDeviceToHost, c->stream));
  }
  if (want_agg) {
    ORCAI_CUDA(c, cudaMemcpyAsync(pin + s->o_agg, b.agg, (size_t)total * 8, cudaMemcpyDeviceToHost, c->stream));
    ORCAI_CUDA(c, cudaMemcpyAsync(pin + s->o_cnt, b.cnt, (size_t)g.S * 8, cudaMemcpyDeviceToHost, c->stream));
  }
  return ORCAI_OK;
}

int postprocess_end(Ctx* c, AsyncSlot* s, double* h_agg, double* h_cnt, int32_t* h_label, int64_t* h_start, int64_t* h_stop,
                    int64_t cap, int64_t* n_seg) {
  const unsigned char* pin = static_cast<const unsigned char*>(s->h_pin);
  *n_seg = 0;
  if ((h_agg || h_cnt) && !s->want_agg) ORCAI_FAIL(c, ORCAI_ERR_ARG, "aggregates requested from a call begun without want_agg");
  unsigned long long total = 0;
  memcpy(&total, pin + kStageTotalOff, 8);
  if (h_agg) memcpy(h_agg, pin + s->o_agg, (size_t)s->n_agg * 8);
  if (h_cnt) memcpy(h_cnt, pin + s->o_cnt, (size_t)s->n_cnt * 8);
  const long long n_starts = (long long)(total >> 31), n_stops = (long long)(total & 0x7fffffffull);
  if (n_starts != n_stops) ORCAI_FAIL(c, ORCAI_ERR_STATE, "segment scan inconsistent: %lld starts vs %lld stops", n_starts, n_stops);
  *n_seg = n_starts;
  if (n_starts > s->cap || n_starts > cap)
    ORCAI_FAIL(c, ORCAI_ERR_CAPACITY, "segment capacity %lld too small, need %lld", (long long)std::min<long long>(s->cap, cap), n_starts);
  const long long first = std::min(n_starts, s->spec);
  if (first) {
    memcpy(h_label, pin + s->o_lab, (size_t)first * 4);
    memcpy(h_start, pin + s->o_sta, (size_t)first * 8);
    memcpy(h_stop, pin + s->o_sto, (size_t)first * 8);
  }
  if (n_starts > s->spec) {
    // the rest sits in the slot's own scratch (complete: `done` has passed); fetched on the copy stream, the compute stream may
    // already hold the next recording
    const long long rest = n_starts - s->spec;
    ORCAI_CUDA(c, cudaMemcpyAsync(h_label + s->spec, s->d_lab + s->spec, (size_t)rest * 4, cudaMemcpyDeviceToHost, c->copy_stream));
    ORCAI_CUDA(c, cudaMemcpyAsync(h_start + s->spec, s->d_sta + s->spec, (size_t)rest * 8, cudaMemcpyDeviceToHost, c->copy_stream));
    ORCAI_CUDA(c, cudaMemcpyAsync(h_stop + s->spec, s->d_sto + s->spec, (size_t)rest * 8, cudaMemcpyDeviceToHost, c->copy_stream));
    ORCAI_CUDA(c, cudaStreamSynchronize(c->copy_stream));
  }
  return ORCAI_OK;
}

int launch_threshold_segments(Ctx* c, const double* h_agg, const double* h_cnt, int64_t S, int L, double threshold,
                              int32_t* h_label, int64_t* h_start, int64_t* h_stop, int64_t cap, int64_t* n_seg) {
  *n_seg = 0;
  if (S <= 0 || L <= 0) return ORCAI_OK;
  const long long total = (long long)S * L;
  const long long n_tiles = (total + kTile - 1) / kTile;
  PostBuffers b;
  ORCAI_CHECK(carve(c, n_tiles, cap, S, L, 0, true, &b));
  ORCAI_CUDA(c, cudaMemcpyAsync(b.agg, h_agg, (size_t)total * 8, cudaMemcpyHostToDevice, c->stream));
  ORCAI_CUDA(c, cudaMemcpyAsync(b.cnt, h_cnt, (size_t)S * 8, cudaMemcpyHostToDevice, c->stream));
  maxcount_kernel<<<64, 256, 0, c->stream>>>(b.cnt, S, &b.scr->maxcnt_bits);
  MaskFromAgg m{b.agg, L, threshold, &b.scr->maxcnt_bits};
  segments_scan_kernel<MaskFromAgg><<<(unsigned)n_tiles, kThreads, 0, c->stream>>>(
      m, S, L, b.tile_state, b.scr, b.seg_label, b.seg_start, b.seg_stop, cap);
  c->launches += 2;
  ORCAI_CUDA(c, cudaGetLastError());
  return copy_out_segments(c, b, cap, h_label, h_start, h_stop, n_seg);
}

}  // namespace orcai
