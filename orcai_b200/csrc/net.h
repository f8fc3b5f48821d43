// orcai-V1 network state shared by the fp32 path (net.cu) and the tensor-core path (net_tc.cu).
#pragma once
#include <cuda_fp16.h>

#include <vector>

#include "common.h"

namespace orcai {

struct TcSep {            // one fused 3x3 implicit-GEMM layer of the tensor-core path
  void* w = nullptr;      // 9 taps x [NP x KP] 16-bit, canonical K-major no-swizzle UMMA layout
  float* bias = nullptr;  // NP floats (folded BatchNorm), zero padded
};

// Channel means of every GEMM's A operand on a calibration recording (net_calibrate).  fp16 weight rounding is the same at
// every pixel, so its effect on a layer's output is, to first order, the constant  sum_k dW[k][n] * mean(A[k])  per output
// channel; the tensor-core paths subtract exactly that from the (split-fp16, exact) bias rows when they pack their operands.
struct Calib {
  bool valid = false;
  double spec = 0.0;                          // normalised spectrogram value (entry convolution input)
  std::vector<double> in_relu[kMaxBlocks];    // ReLU(block input)                 -> first separable convolution
  std::vector<double> in_even[kMaxBlocks];    // block input at even positions     -> residual 1x1/2 convolution
  std::vector<double> s1[kMaxBlocks];         // first sepconv output (post-ReLU)  -> second separable convolution
  std::vector<double> fin;                    // block-4 output                    -> final separable convolution
  std::vector<double> feat, h1, h2;           // inputs of LSTM 1, LSTM 2 (and the recurrent terms), Dense(128)
};

struct NetWeights {
  bool loaded = false;
  Calib calib;
  int n_blocks = 0;
  int filters[kMaxBlocks] = {};
  int Wf = 0, H = 0, U = 0, L = 0, feat = 0;
  // device weights (fp32)
  float* conv0_w = nullptr;  // [9][16]
  float* conv0_b = nullptr;  // [16]
  struct Sep { float* dw = nullptr; float* pw = nullptr; float* b = nullptr; int ci = 0, co = 0; };
  Sep sep1[kMaxBlocks], sep2[kMaxBlocks], fin;
  float* res_w[kMaxBlocks] = {};  // [ci][co]
  float* res_b[kMaxBlocks] = {};
  float* lstm_wih[2] = {};  // [I][2*4U]  (forward gates | backward gates)
  float* lstm_bih[2] = {};  // [2*4U]
  float* lstm_whh[2] = {};  // [2][U][4U]
  float* d1_w = nullptr; float* d1_b = nullptr;   // [2U][128], [128]
  float* d2_w = nullptr; float* d2_b = nullptr;   // [128][L] with bn_dense folded, [L]
  std::vector<void*> allocs;
  // activation workspace
  float* ws = nullptr; size_t ws_cap = 0;
  int chunk = 128;
  // per-stage CUDA events of the first chunk of the last forward (profiling aid, see orcai_timings.net_stage_ms)
  static constexpr int kNumMarks = 14;
  cudaEvent_t ev[kNumMarks] = {};
  int mark_i = 0;
  long long marked_snippets = 0;

  // ---- tensor-core path (net_tc.cu): operands per 16-bit format (0 = fp16, 1 = bf16) ----
  bool tc_ready[2] = {false, false};
  void* tc_conv0_w[2] = {};   // [16 x 16] canonical
  TcSep tc_sep1[2][kMaxBlocks], tc_sep2[2][kMaxBlocks], tc_fin[2];
  void* tc_ws = nullptr; size_t tc_ws_cap = 0;
  int path = 0;               // 0 = fp32 CUDA cores, 1 = fp16 tensor cores, 2 = bf16 tensor cores, 3 = fp16 fused residual blocks
  // fused residual-block kernels (net_fused.cuh): packed fp16 operands [sep1 | sep2 | residual] and fp32 biases per block
  bool fused_ready = false;
  void* fb_w[kMaxBlocks] = {};
  void* fbp_w[kMaxBlocks] = {};   // precise path (net_path 4): split-fp16 (hi, lo) operand sets of the fused blocks
  float* fb_bias[kMaxBlocks] = {};
  int chunk_fused = 2048;
  TcSep fb_fin;                  // final separable convolution with the calibrated bias (fused path)
  float* tc_bih[2] = {};         // projection biases with the weight-rounding correction folded in
  float* tc_d1_b = nullptr;
  void* conv0_mma_w = nullptr;   // banded B operand of the tensor-core entry convolution (conv0_mma.cuh)
  int block1_path = 0;           // fused path, block 1: 0 = one MMA per tap (net_fused.cuh), 1 = N-widened MMAs (net_fused_w.cuh)
  void* fb_w1_wide = nullptr;    // operands of the N-widened block 1
  int conv0_path = 1;            // fused path: 1 = tensor-core pixel-group convolution, 0 = fp32 CUDA-core convolution
  // tensor-core recurrent tail (net_lstm_tc.cu)
  bool tail_tc_ready = false;
  int tail_path = 1;          // fused path only: 1 = tensor-core LSTM / dense tail, 0 = fp32 CUDA-core tail
  __half* tc_wih[2] = {};     // packed B blocks of the input projections [I][2*4U]
  __half* tc_whh[2] = {};     // [2 directions][4U x U] canonical K-major
  __half* tc_d1 = nullptr;    // Dense(128) B blocks
  std::vector<float> h_lstm_wih[2], h_lstm_whh[2], h_lstm_bih[2], h_d1_w, h_d1_b;
  std::vector<float> h_res_w[kMaxBlocks], h_res_b[kMaxBlocks];
  // fp32-grade tensor-core path (net_path 4, net_precise.cuh)
  bool precise_ready = false, tail_precise_ready = false;
  __half* tp_wih[2] = {};     // split (hi, lo) B blocks of the LSTM input projections
  __half* tp_d1 = nullptr;    // ... of Dense(128)
  int precise_lstm_tc = 1;    // net_path 4: 1 = split-fp16 tensor-core recurrence on CTA pairs, 0 = fp32 CUDA-core recurrence
  __half* tp_whh_hi[2] = {}, *tp_whh_lo[2] = {};   // W_hh^T as (hi, lo) fp16, [2 directions][4U x U] canonical K-major (the split recurrence)
  struct PreciseSep { float* dw = nullptr; __half* pw = nullptr; float* bias = nullptr; };   // [9][CIP] fp32, split B blocks [CIP][64], 64 floats
  PreciseSep p_sep1[kMaxBlocks], p_sep2[kMaxBlocks], p_fin;
  float* p_res_w[kMaxBlocks] = {};   // [CIP][COP] fp32, zero padded
  float* p_res_b[kMaxBlocks] = {};   // [COP]
  int chunk_precise = 2048;
  int precise_sep_path = 1;   // un-folded separable convolutions: 1 = depthwise fused into the split GEMM (sep_uf_kernel), 0 = two kernels
  int precise_tall = 1;       // resident recordings: trunk once over the chunk's rows as one tall image + per-snippet border rows
  int debug_stop = -1;        // stop the forward after this stage (debug reads), -1 = run everything
  // last debug buffer: kind 0 f32, 1 fp16, 2 bf16 ; NHWC with channel pitch dbg_pitch
  const void* dbg_ptr = nullptr; int dbg_kind = 0; long long dbg_n = 0; int dbg_h = 0, dbg_w = 0, dbg_c = 0, dbg_pitch = 0;
  // folded fp32 weights kept on the host for building the tensor-core operands
  std::vector<float> h_conv0_w, h_conv0_b, h_conv0_pack;
  struct HostSep { std::vector<float> dw, pw, b; int ci = 0, co = 0; };
  HostSep h_sep1[kMaxBlocks], h_sep2[kMaxBlocks], h_fin;
};

inline void net_mark(Ctx* c, bool on) {
  NetWeights* nw = c->net;
  if (on && nw->mark_i < NetWeights::kNumMarks) cudaEventRecord(nw->ev[nw->mark_i++], c->stream);
}

// LSTM + dense tail on fp32 features (m, Tn, feat).  scratch: m*Tn*(2*4U + 2U + 2U + 128) floats.
int net_tail_fp32(Ctx* c, const float* feat, float* scratch, long long m, float* d_preds_out, bool mark);
int net_upload(Ctx* c, const std::vector<float>& v, float** dptr);
int net_tail_tc(Ctx* c, const float* feat, float* scratch, long long m, float* d_preds_out, bool mark);
// fp32-grade path (net_path 4): split-fp16 tensor-core GEMMs (net_lstm_tc.cu), fp32 recurrence (net.cu)
int net_tail_precise(Ctx* c, const float* feat, float* scratch, long long m, float* d_preds_out, bool mark);
int net_pack_split_b(Ctx* c, const float* w, int K, int n_src, int N, __half** out);
int net_gemm_split(Ctx* c, const float* A, int lda, const __half* Bp, const float* bias, float* C, int ldc, long long M, int N, int K, int n_valid,
                   int act);
int net_lstm_rec_fp32(Ctx* c, const float* xz, const float* whh, float* out, long long m, int Tn);
int net_tc_prepare(Ctx* c, int fmt);
int net_calibrate(Ctx* c, int64_t max_snippets);
int net_forward_tc(Ctx* c, const float* d_in, int input_mode, int64_t first, int64_t n, float* d_preds);

}  // namespace orcai
