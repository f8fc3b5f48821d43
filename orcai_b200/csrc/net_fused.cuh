// One residual block of orcai-V1 as ONE persistent, warp-specialised tcgen05 kernel (included by net_tc.cu).
//
// Reference graph: src/orcAI/architectures.py:172-196 —
//     x -> ReLU -> SepConv3x3 -> BN -> ReLU -> SepConv3x3 -> BN -> MaxPool(3,2)/2 "same"  (+)  Conv1x1/2(x)  -> y
//
// Data layout (global, fp16 NHWC, channel pitch padded to 8, padding channels zero):
//     Xr   (n, H,  W,  ICP)   ReLU(x)            — A operand of the first separable convolution
//     Xsub (n, Ho, Wo, ICP)   x at even (h, w)   — input of the residual 1x1/2 convolution (x before the ReLU)
//     Yr   (n, Ho, Wo, OCP)   ReLU(y) (or y)     — next block's Xr
//     Ysub (n, Ho/2, ceil(Wo/2), OCP)  y at even positions — next block's Xsub
//
// A CTA owns a strip of CP pooled columns of one snippet and marches down the time axis in steps of S rows.
// Shared memory holds three pixel-linear, channel-chunk-planar buffers with a common row pitch WP = 2*CP + 4:
//     X  [k-chunk][(S+2) x WP pixels][8 ch]   input rows a+1 .. a+S+2
//     S1 [k-chunk][(S+2) x WP pixels][8 ch]   first sepconv output rows a .. a+S+1   (rows 0,1 carried from the last step)
//     S2 [k-chunk][(S+1) x WP pixels][8 ch]   second sepconv output rows a .. a+S    (row 0 carried)
// Because pixels are linear with a fixed pitch, an M=128 MMA tile is simply 128 CONSECUTIVE pixels (SBO = 128 B,
// LBO = one chunk plane) and the A operand of tap (dy,dx) is the same buffer with the descriptor start shifted by
// (dy*WP + dx) pixels: no im2col.  Tiles run over the halo columns too; those accumulator rows are garbage and are
// never consumed (every accumulator row depends on its own A row only).  The depthwise filter is folded into the
// GEMM weights (W'_tap = dw[tap] * pw * bn_scale), the residual 1x1 convolution is one more small MMA into its own
// TMEM columns, and the max-pool + residual add + ReLU run in the epilogue straight out of shared memory / TMEM.
// HBM sees each block's input once (plus 4 halo columns per strip) and its pooled output once.
//
// Warp roles: NEW worker warps (TMEM drains, pooling, stores), one issuer warp whose elected lane issues every
// tcgen05.mma, and one producer warp that feeds X / R with TMA (5-D tensor maps over the NHWC activations: the box
// {8 ch, 1 chunk, WP cols, S+2 rows, 1 snippet} lands as one pixel-linear chunk plane, borders are zero-filled by the
// TMA unit).  They meet only through mbarriers, so the tensor pipe runs the NEXT step's first convolution while the
// workers pool and store the current one.  Biases ride on the tensor pipe as well: every accumulator tile starts
// with one K=16 MMA of a constant "ones" operand against [bias_hi, bias_lo] rows (fp16 split, exact to 2^-22).
//     x_full      producer -> issuer  X / R tiles of a step have landed (TMA complete_tx)
//     bar1[t]     issuer  -> workers  tcgen05.commit: accumulator tile t of the first convolution is complete
//     s1_full[t]  workers -> issuer   S1 tile t (and the carried rows) written
//     bar2[t]     issuer  -> workers  accumulator tile t of the second convolution is complete
//     barR[2]     issuer  -> workers  residual accumulator (double buffered across steps)
//     pool_done   workers -> issuer   the pooling epilogue has read its residual accumulator (it may be overwritten)
// Worker order inside step g:  drain conv 1 (g) -> pool + store step g-1 (while the tensor pipe runs conv 2 of g)
//                              -> drain conv 2 (g) -> carry rows;  conv 1 of g+1 is already in flight by then.
#pragma once

namespace fused {

// block 1 of the fp32-grade path (FB<..., PREC, UF2>): residual accumulators (with two X buffers) and whether the second issuer
// warp takes half of the first convolution's tiles
#ifndef ORCAI_B1_RB
#define ORCAI_B1_RB 3
#endif
#ifndef ORCAI_B1_SPLIT1
#define ORCAI_B1_SPLIT1 0
#endif

__host__ __device__ constexpr int imax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int imin(int a, int b) { return a < b ? a : b; }
__host__ __device__ constexpr int round8(int a) { return (a + 7) & ~7; }
__host__ __device__ constexpr int pow2cols(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

// CONV0_: the block's input is produced in-kernel from the fp16 spectrogram by the entry convolution (CUDA-core producer warps)
// ISS_: issuer warps.  One tcgen05.mma stream sustains ~59 cycles per M128 K16 MMA whatever N <= 64 is; two streams on an
//        SM reach the shared-memory bound (41 / 47 / 52 cycles at N = 32 / 48 / 64; tools/microbench/mma_cost.cu).  With
//        ISS_ = 2 the first convolution (+ residual 1x1) and the second convolution are issued by separate warps
//        (splitting the TILES of each convolution between two warps instead measured 7-10 % slower).
// XBUF_: X / R operand buffers.  With one buffer the load of step g+1 can only start when the first convolution of step g has
//        completed, and the first convolution of step g+1 only when it has landed 2-4 k cycles later
//        (profiles/r01l_fused_block_handoff_timeline.txt).  Two buffers (where the shared memory allows: block 2) take the
//        load off that cycle: 1.18 -> 1.12 ms per 10 min of audio.  (Also tried: issuing the residual MMA after the tiles so
//        that the first convolution does not wait for the pooling epilogue of step g-2 - it then competes with the second
//        convolution of step g, which is on the critical path: 1.12 -> 1.25 ms.)
// PREC_: fp32-grade arithmetic on the fp16 tensor cores.  Every operand is the pair (hi, lo) = (fp16(v), fp16(v - hi)), exact to
//        2^-22, and every product runs as  A_hi*W_hi + A_lo*W_hi + A_hi*W_lo  into the same fp32 accumulator (the dropped
//        A_lo*W_lo term is 2^-22 relative): the X / R / S1 / S2 buffers, the weight set and the activation tensors in HBM all
//        come as a hi plane set followed by a lo plane set.  Why: fp16 operands alone put the probabilities 2.7e-3 from the
//        fp32 graph (tools/precision_plan.py: every weight tensor and every stored activation contributes 2e-4 .. 2e-3).
// UF2_ (with PREC_): the SECOND separable convolution un-folded.  Folding the depthwise filter into the GEMM costs nine taps x
//        three split products = 27 MMAs per K chunk; instead the worker warps run the depthwise 3x3 on the CUDA cores (fp32 FMAs on
//        an fp32 S1 kept in quad-planar [4-channel quad][pixel][4 floats] layout: one conflict-free LDS.128 per window position,
//        sliding window down the rows) straight into a (hi, lo) A operand D2, and the tensor pipe runs the pointwise 1x1 only:
//        3 MMAs per K chunk.  S1 is then no MMA operand any more, hence fp32.
//        (Un-folding the FIRST convolution the same way was built and measured slower - 3.25 vs 2.98 ms per 10 min of audio,
//        profiles/r02x_block1_uf1_experiment.log: its depthwise pass lands on the worker warps' critical chain, while the folded
//        form's 27 MMAs per K chunk run on an otherwise idle tensor pipe; with the pass on four dedicated warps it was still
//        slower, 9.28 vs 8.90 ms per 1 024 snippets, profiles/r02zd_block1_uf1_dedicated_warps_experiment.log: the workers' chain
//        does not shorten when the tensor pipe's operand reads go away.  Git history: eff5589 and its successor.  Likewise measured
//        and dropped: a second X buffer (9.0 ms), S1 / S2 as row rings without carry copies and three of the five barriers per step
//        (9.5 ms, profiles/r02zh_block1_rings_xbuf2_experiment.log), the first convolution queued behind the second one (9.1 ms).)
template <int CIN_, int COUT_, int CPOOL_, int S_, bool RELU_OUT_, int CTAS_, int NEW_, bool CONV0_ = false, int ISS_ = 1, int XBUF_ = 1, bool PREC_ = false,
          bool UF2_ = false>
struct FB {
  static constexpr int XBUF = XBUF_;
  static constexpr bool PREC = PREC_, UF2 = UF2_;
  static_assert(!UF2_ || (PREC_ && ISS_ == 2), "the un-folded second convolution: split-fp16 configuration with two issuer warps");
  static constexpr int PL = PREC_ ? 2 : 1;           // operand plane sets: hi (, lo)
  static_assert(!(PREC_ && CONV0_), "the in-kernel entry convolution writes single fp16 operands");
  static_assert(XBUF_ == 1 || (XBUF_ == 2 && ISS_ == 2 && !CONV0_), "two X buffers: two-issuer TMA configuration only");
  static constexpr int CIN = CIN_, COUT = COUT_, CP = CPOOL_, S = S_, CTAS = CTAS_, NEW = NEW_, ISS = ISS_;
  static constexpr bool RELU_OUT = RELU_OUT_, CONV0 = CONV0_;
  static constexpr int NPROD = CONV0 ? 2 : 1;                       // producer warps: one TMA warp, or two entry-convolution warps
  static constexpr int NWORK = NEW * 32, NTHREADS = NWORK + 32 * ISS + 32 * NPROD;   // + issuer warp(s) + producer warp(s)
  static_assert(ISS == 1 || ISS == 2, "one or two issuer warps");
  static constexpr int NT = NEW / 4;                 // worker teams per TMEM lane quadrant; a team drains 16 columns
  static constexpr int ICP = cpad8(CIN), OCP = cpad8(COUT);
  static constexpr int KP1 = cpad16(CIN);          // K per tap of sepconv 1 and of the residual convolution
  static constexpr int NP = cpad16(COUT);          // N of every MMA; K per tap of sepconv 2
  static constexpr int XG = ICP / 8;               // X / R chunk planes that carry data
  static constexpr int XCH = KP1 / 8;              // X / R chunk planes the MMA reads (the extra one stays zero)
  static constexpr int NG = OCP / 8;               // S1 / S2 chunk planes that carry data
  static constexpr int MCH = NP / 8;               // S1 chunk planes the MMA reads
  static constexpr int WP = 2 * CP + 4;
  static constexpr int N1 = (S * WP - 2 + 127) / 128, N2 = UF2_ ? 2 : (S * WP - 4 + 127) / 128;
  static constexpr int P1_0 = 2 * WP + 1, P2_0 = WP + 2;
  // one buffer: padded so that the last tile's over-read stays inside it; two buffers: packed (the over-read of rows that
  // feed no consumed accumulator row runs into the next plane / buffer, which is finite shared memory of this CTA)
  static constexpr int XPIX = XBUF_ == 2 ? round8((S + 2) * WP) : round8(imax((S + 2) * WP, P1_0 + 128 * N1 + 1));
  static constexpr int S1PIX = round8(imax((S + 2) * WP, P2_0 + 128 * N2 + WP + 1));
  // S2 is stored de-interleaved: even image columns in the first half of a chunk plane, odd columns in the second half
  // (offset by 64 B modulo 128), so that the pooling epilogue's lanes (one pooled column each) read consecutive 16-byte pieces
  static constexpr int S2HALF = round8((S + 1) * (WP / 2)) + 4;
  static constexpr int S2PIX = 2 * S2HALF;
  static constexpr int RQ = (S / 2) * CP, RPIX = round8(RQ);
  static constexpr uint32_t LBO_X = XPIX * 16, LBO_S1 = S1PIX * 16, LBO_S2 = S2PIX * 16, LBO_R = RPIX * 16;
  // trimmed weight storage: only k-chunks / n-groups that carry data are stored; the MMA's reads of the missing
  // chunk alias the next group (finite values times a zero A chunk), missing n-groups only feed unused columns.
  static constexpr uint32_t SBO_W1 = XG * 128, SBO_W2 = NG * 128;
  static constexpr uint32_t TAP_W1 = NG * SBO_W1, TAP_W2 = NG * SBO_W2;
  static constexpr uint32_t W1_BYTES = 9 * TAP_W1 + 128, W2_BYTES = (UF2_ ? 1 : 9) * TAP_W2 + 128, WR_BYTES = TAP_W1 + 128;
  static constexpr uint32_t WB_BYTES = NG * 128 + 128;   // [bias_hi, bias_lo] rows of one GEMM: n-groups of one k-chunk
  static constexpr uint32_t OFF_W1 = 0, OFF_W2 = OFF_W1 + W1_BYTES, OFF_WR = OFF_W2 + W2_BYTES;
  static constexpr uint32_t WSET = OFF_WR + WR_BYTES;        // one weight set [sep1 | sep2 | residual]; PREC: the lo set follows
  static constexpr uint32_t W_LO = PREC_ ? WSET : 0;
  static constexpr uint32_t OFF_WB1 = WSET * PL, OFF_WB2 = OFF_WB1 + WB_BYTES, OFF_WBR = OFF_WB2 + WB_BYTES;
  static constexpr uint32_t OFF_ONES = OFF_WBR + WB_BYTES;   // two 8x8 core matrices: k = 0,1 are 1.0, the rest 0 (SBO = 0)
  static constexpr uint32_t OFF_DW2 = OFF_ONES + 256;          // UF2: depthwise taps of the second convolution, [9][NP] fp32
  static constexpr uint32_t W_BYTES = OFF_DW2 + (UF2_ ? 9 * NP * 4 : 0);
  static constexpr uint32_t OFF_R = W_BYTES;
  static constexpr uint32_t R_LO = XCH * LBO_R, X_LO = XCH * LBO_X, S1_LO = MCH * LBO_S1, S2_LO = NG * LBO_S2;   // hi -> lo plane set
  static constexpr int NQ = OCP / 4;                             // UF2: fp32 S1 planes (4-channel quads), LBO_S1F apart
  // UF2: an fp32 S1 plane holds the even image columns of all rows, then (S1HALF pixels further) the odd columns: [parity][row][column / 2].
  // The depthwise pass gives every thread a PAIR of adjacent columns (4 window loads per row instead of 2 x 3) and its lanes then
  // read consecutive 16-byte pieces of one half.  S1HALF = 5 (mod 8) pixels keeps the first epilogue's stores conflict free (its
  // lanes alternate odd / even columns starting with an odd one: the two 64-byte runs of a quarter warp land 64 bytes apart modulo 128);
  // the plane pitch is 64 (mod 128) bytes so that lane pairs working on quads (q, q ^ 1) of the same pixels do not collide either.
  static constexpr int S1HALF = ((S + 2) * (WP / 2) + 7) / 8 * 8 + 5;
  static constexpr uint32_t LBO_S1F = S1PIX * 16 + 64;
  static_assert(!UF2_ || (S1HALF + (S + 2) * (WP / 2) <= S1PIX && WP % 2 == 0), "fp32 S1 plane: two column-parity halves");
  static constexpr int DWP = CP;                                 // UF2: column pairs the second convolution is evaluated on (columns 2 .. 2 CP + 1)
  static_assert(!UF2_ || S * DWP <= 128, "UF2: the even and the odd columns of a step are one accumulator tile each");
  static constexpr uint32_t S1_BYTES = UF2_ ? NQ * LBO_S1F : PL * MCH * LBO_S1;
  static constexpr int D2PIX = 2 * 128;                          // UF2: A operand of the pointwise GEMM, accumulator-row order: tile 0 = even columns, tile 1 = odd columns, row = (y - 1) * DWP + column pair
  static constexpr uint32_t LBO_D2 = D2PIX * 16, D2_LO = MCH * LBO_D2, D2_BYTES = UF2_ ? 2 * MCH * LBO_D2 : 0;
  static constexpr uint32_t OFF_X = OFF_R + PL * XCH * LBO_R;
  static constexpr uint32_t XR_BYTES = PL * (XCH * LBO_R + XCH * LBO_X);        // one R + X buffer; buffer b sits b * XR_BYTES further
  static constexpr uint32_t OFF_S1 = OFF_X + PL * XCH * LBO_X + (XBUF - 1) * XR_BYTES;
  static constexpr uint32_t OFF_S2 = OFF_S1 + S1_BYTES;
  // entry-convolution input: two (S+5) x SPW fp16 tiles of the normalised spectrogram (rows a-1 .. a+S+3, columns cb-1 ..)
  static constexpr int SPW = 64, SPH = S + 5;
  static constexpr uint32_t SPEC_BYTES = CONV0 ? SPH * SPW * 2 : 0;
  static constexpr uint32_t OFF_D2 = OFF_S2 + PL * NG * LBO_S2;
  static constexpr uint32_t OFF_SPEC = OFF_D2 + D2_BYTES;
  static constexpr uint32_t OFF_BAR = OFF_SPEC + 2 * ((SPEC_BYTES + 127) / 128 * 128);
  // residual accumulators.  Two cover "the pooling epilogue of step g-1 runs inside worker step g".  The un-folded configuration with
  // two X buffers takes three: the first convolution (and residual MMA) of step g+1 is then issued as soon as the workers have
  // drained step g's first-convolution accumulators, i.e. it runs on the tensor pipe WHILE the workers' depthwise pass of step g
  // runs on the CUDA cores - at that point the residuals of steps g-1 (not pooled yet), g and g+1 are all live.
  static constexpr int RB = (UF2_ && XBUF_ == 2) ? ORCAI_B1_RB : 2;
  // UF2: the tiles of the first convolution split between the two issuer warps (the second one is idle until the depthwise pass is done)
  static constexpr int N1A = (UF2_ && ORCAI_B1_SPLIT1) ? (N1 + 1) / 2 : N1;
  // barriers: bar1[N1] bar2[N2] barR[RB] s1_full[N1] pool_done[RB] x_full[2] x_free[2] spec_full[2] d2_full carry_s1 carry_s2
  static constexpr int B_1 = 0, B_2 = N1, B_R = N1 + N2, B_S1 = N1 + N2 + RB, B_P = 2 * N1 + N2 + RB, B_X = B_P + RB, B_XF = B_X + 2, B_SP = B_XF + 2, B_D2 = B_SP + 2, B_C1 = B_D2 + 1, B_C2 = B_D2 + 2, NBAR = B_D2 + 3;
  static constexpr uint32_t SMEM = OFF_BAR + NBAR * 8 + 16;
  static constexpr uint32_t TX_BYTES = PL * XG * ((S + 2) * WP + (S / 2) * CP) * 16;   // bytes one step's TMA loads deliver
  static constexpr int COL_R = 0, COL_1 = RB * NP, COL_2 = RB * NP + N1 * NP;
  static constexpr int TM_COLS = pow2cols(NP * (RB + N1 + N2));
  static_assert(NEW == 8 || NEW == 16, "worker warps come in groups of four (one per TMEM lane quadrant)");
  static_assert(NT * 2 >= NG, "each worker team drains two channel groups (16 accumulator columns)");
  static_assert(S % 2 == 0 && S >= 2, "steps advance by whole pooled rows");
  static_assert(RQ <= 128, "one residual MMA tile per step");
  static_assert(WP <= 126, "a second-convolution tile may only depend on first-convolution tiles t-1 .. t+1");
  static_assert(!CONV0 || (CIN == 16 && WP + 2 <= SPW && 2 * CP + 3 <= SPW), "entry-convolution tile: 16 channels, WP + 2 spectrogram columns");
  static_assert(NP * (RB + N1 + N2) <= 512 && TM_COLS * CTAS <= 512, "TMEM columns");
  static_assert(SMEM <= 227 * 1024 && (SMEM + 1024) * CTAS <= 228 * 1024, "shared memory (per CTA and per SM)");
  static_assert((128 - RPIX) * 16 <= XCH * LBO_X, "residual tile over-read must stay inside the CTA's shared memory");
  static_assert(W_BYTES % 128 == 0 && LBO_X % 128 == 0 && LBO_R % 128 == 0, "TMA destinations are 128-byte aligned");
};

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA: one box of a rank-5 tensor map -> shared memory, completion counted on an mbarrier
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
// Measured and left off (-DORCAI_B1_ACARRY=1 builds it): the carried S1 / S2 rows copied by the TMA unit (shared -> shared bulk copies)
// instead of through the workers' registers.  The copies queue in the TMA unit behind the next step's X tile (~6 k cycles) and
// the workers end up waiting for them: block 1 of a 1-h recording 18.75 ms against 16.06 ms, with one X buffer or two.
#ifndef ORCAI_B1_ACARRY
#define ORCAI_B1_ACARRY 0
#endif
constexpr bool kAsyncCarry = ORCAI_B1_ACARRY != 0;   // (written for the linear S1 layout of its time; the un-folded configuration's S1 is now split by column parity)
// timing experiments only (wrong results): fewer taps in block 1's first convolution, no depthwise pass, no pooling epilogue
#ifndef ORCAI_EXP_TAPS
#define ORCAI_EXP_TAPS 9
#endif
#ifndef ORCAI_EXP_NODW
#define ORCAI_EXP_NODW 0
#endif
#ifndef ORCAI_EXP_NOPOOL
#define ORCAI_EXP_NOPOOL 0
#endif
// shared -> shared copy inside the CTA through the TMA unit (no thread touches the data), completion counted on an mbarrier
__device__ __forceinline__ void bulk_copy_s2s(uint32_t dst, uint32_t src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "r"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               :
               : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred)
      :
      : "memory");
  return pred != 0;
}
// Bring-up aid (compile with -DORCAI_FUSED_TRACE, run with ORCAI_B200_TRACE=<file>): CTA 0 stamps clock64() at the hand-offs
// of a few steps into device memory (plain stores, one slot per event); tools/bringup/trace_timeline.py prints the timeline.  Compiled out by default.
#ifdef ORCAI_FUSED_TRACE
static __device__ long long* g_trace = nullptr;   // device memory: [CIN / 10][step - 40][tag] clock64() stamps, plain stores
__device__ __forceinline__ void trace_event(int kernel, int tag, long long g) {
  long long* t = g_trace;
  if (t == nullptr || blockIdx.x != 0 || g < 40 || g >= 44 || (threadIdx.x & 31) != 0) return;
  t[(kernel / 10) * 256 + (int)(g - 40) * 64 + tag] = clock64();
}
#define FB_TRACE(tag, g) trace_event(G::CIN, tag, g)
#else
#define FB_TRACE(tag, g)
#endif

template <int N>
__device__ __forceinline__ void worker_sync() {   // named barrier 1: the worker warps only
  asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory");
}

// TMEM -> registers: 32 lanes x 16 consecutive fp32 columns (load and wait in one statement)
__device__ __forceinline__ void tmem_ld16f(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint4 relu8h(uint4 a) {
  __half2* x = reinterpret_cast<__half2*>(&a);
  const __half2 z = __float2half2_rn(0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = __hmax2(x[i], z);
  return a;
}
__device__ __forceinline__ uint4 pack8h(const float* v) {
  uint4 r;
  __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  return r;
}
// (hi, lo) = (fp16(v), fp16(v - hi)): v to 2^-22
__device__ __forceinline__ void split8h(const float* v, uint4& hi, uint4& lo) {
  hi = pack8h(v);
  const __half2* h = reinterpret_cast<const __half2*>(&hi);
  float r[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    r[2 * i] = v[2 * i] - f.x;
    r[2 * i + 1] = v[2 * i + 1] - f.y;
  }
  lo = pack8h(r);
}
// two fp32 FMAs in one instruction (FFMA2, sm_100): d = a * b + d per lane, each rounded like a scalar fmaf
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  unsigned long long d, a, b;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
// elementwise max of m[8] with the values (hi + lo)
__device__ __forceinline__ void fmax8_split(float (&m)[8], uint4 hi, uint4 lo) {
  const __half2* h = reinterpret_cast<const __half2*>(&hi);
  const __half2* l = reinterpret_cast<const __half2*>(&lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 a = __half22float2(h[i]), b = __half22float2(l[i]);
    m[2 * i] = fmaxf(m[2 * i], a.x + b.x);
    m[2 * i + 1] = fmaxf(m[2 * i + 1], a.y + b.y);
  }
}
__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
  __half2* x = reinterpret_cast<__half2*>(&a);
  const __half2* y = reinterpret_cast<const __half2*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = __hmax2(x[i], y[i]);
  return a;
}

// Tall-image mode (net_precise.cuh): the images of the batch are overlapping row windows of ONE tall image - image b covers
// tall rows [b * stride - warm, b * stride - warm + H) (the tensor maps are built with that outer stride and a base shifted by
// -warm rows).  Padding ("outside the image") is then decided on TALL rows [0, rows), the first warm / 2 output rows of every
// window are warm-up (the carried rows of a window's first step are not real) and are not stored, and output row ho of window b
// is tall output row (b * stride - warm) / 2 + ho.  on = 0: independent images, as before.
struct TallView {
  int on = 0, stride = 0, warm = 0;
  long long rows = 0;
};

template <class G>
#if ORCAI_B1_ACARRY
__global__ void __cluster_dims__(1, 1, 1) __launch_bounds__(G::NTHREADS, G::CTAS)   // a cluster of one: shared -> shared bulk copies address shared::cluster
#else
__global__ void __launch_bounds__(G::NTHREADS, G::CTAS)
#endif
fused_block_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmR, int r_step, __half* __restrict__ Yr,
                   __half* __restrict__ Ysub, int H, int W, int n_strips, long long n_items,
                   const unsigned char* __restrict__ wpack,
                   // PREC: maps of the lo tensors; Yr / Ysub are then fp32 arrays
                   const __grid_constant__ CUtensorMap tmXl, const __grid_constant__ CUtensorMap tmRl, const TallView tall) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G::OFF_BAR);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + G::NBAR);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Ho = H >> 1, Wo = (W + 1) >> 1;
  const int Hs = Ho >> 1, Ws = (Wo + 1) >> 1;
  const int n_steps = (Ho + 1 + G::S / 2 - 1) / (G::S / 2);

  for (int i = tid; i < (int)(G::W_BYTES / 16); i += G::NTHREADS) reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(wpack) + i);
  for (int i = tid + G::W_BYTES / 16; i < (int)(G::OFF_BAR / 16); i += G::NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < G::B_S1; ++i) mbar_init(&bars[i], 1);                 // tcgen05.commit arrivals
    for (int i = G::B_S1; i < G::B_X; ++i) mbar_init(&bars[i], G::NEW);        // one arrival per worker warp (s1_full, pool_done)
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[G::B_X + i], G::CONV0 ? G::NPROD : 1);                   // TMA arrive.expect_tx, or one arrival per entry-conv warp
      mbar_init(&bars[G::B_XF + i], G::N1A < G::N1 ? 2 : 1);                   // tcgen05.commit (one per issuing warp): the buffer's first convolution is done
    }
    mbar_init(&bars[G::B_SP], 1); mbar_init(&bars[G::B_SP + 1], 1);            // spectrogram tiles of the entry convolution
    mbar_init(&bars[G::B_D2], G::NEW);                                         // UF2: D2 written, one arrival per worker warp
    mbar_init(&bars[G::B_C1], 1); mbar_init(&bars[G::B_C2], 1);                // carried S1 / S2 rows copied (bulk copies, complete_tx)
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc<G::TM_COLS>(tslot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t sbase = smem_u32(smem);

  const long long my_items = blockIdx.x < n_items ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long total_steps = my_items * n_steps;

  if (warp >= G::NEW + G::ISS) {
    if constexpr (G::CONV0) {
      // =============================== entry-convolution producers ===============================
      // Conv2D 3x3 1->16 + folded BatchNorm + ReLU (architectures.py:162-168) on the CUDA cores, straight into the X planes
      // (and the residual-input planes R = X at even positions): fp32 math on the fp16 spectrogram, weights as constant-bank
      // FFMA operands.  tmX is the rank-4 map of the strip-cut spectrogram (conv0::spec_strips_kernel: SPW columns of a strip,
      // strip, rows of a snippet, snippet); its outer stride is the snippet shift, rows outside the snippet are zero-filled,
      // columns outside the image are zero in the buffer.  Lane = X column; 64 lanes walk the S+2 rows with a sliding 3x3 window.
      const int pl = (warp - G::NEW - G::ISS) * 32 + lane;     // 0..63
      const __half* s_spec = reinterpret_cast<const __half*>(smem + G::OFF_SPEC);
      constexpr uint32_t SPEC_STRIDE = (G::SPEC_BYTES + 127) / 128 * 128;
      auto load_spec = [&](long long gg, long long item2, int step2) {   // elected lane of producer warp 0
        const long long b2 = item2 / n_strips;
        const int strip2 = (int)(item2 - b2 * n_strips);
        mbar_arrive_expect_tx(&bars[G::B_SP + (int)(gg & 1)], G::SPEC_BYTES);
        tma_load_4d(sbase + G::OFF_SPEC + (uint32_t)(gg & 1) * SPEC_STRIDE, &tmX, &bars[G::B_SP + (int)(gg & 1)], 0, strip2,
                    step2 * G::S - 3, (int)b2);
      };
      if (warp == G::NEW + G::ISS && total_steps > 0 && elect_one()) load_spec(0, blockIdx.x, 0);
      __syncwarp();
      long long g = 0;
      for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long long b = item / n_strips;
        const int strip = (int)(item - b * n_strips);
        const int wo0 = strip * G::CP, cb = 2 * wo0 - 2;
        for (int step = 0; step < n_steps; ++step, ++g) {
          const int a = step * G::S - 2;
          // X and R are free once the previous step's first convolution (and residual MMA) has completed; that also
          // means BOTH producer warps have finished step g-1, i.e. the other spectrogram buffer has no readers left
          if (g > 0) mbar_wait(&bars[G::B_1 + G::N1 - 1], (uint32_t)((g - 1) & 1));
          // prefetch the next step's spectrogram tile into the other buffer
          if (warp == G::NEW + G::ISS && g + 1 < total_steps && elect_one()) {
            if (step + 1 < n_steps) load_spec(g + 1, item, step + 1);
            else load_spec(g + 1, item + gridDim.x, 0);
          }
          __syncwarp();
          mbar_wait(&bars[G::B_SP + (int)(g & 1)], (uint32_t)((g >> 1) & 1));
          const __half* sp = s_spec + (size_t)(g & 1) * (SPEC_STRIDE / 2);   // tile row y = image row a-1+y, tile col x = image col cb-1+x
          auto conv_px = [&](const float (&w)[3][3], bool inimg, unsigned char* dst0, unsigned char* dst1) {
            float acc[16];
#pragma unroll
            for (int ch = 0; ch < 16; ++ch) acc[ch] = c_conv0[144 + ch];
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
              for (int ch = 0; ch < 16; ++ch) acc[ch] = fmaf(w[t / 3][t % 3], c_conv0[t * 16 + ch], acc[ch]);
#pragma unroll
            for (int ch = 0; ch < 16; ++ch) acc[ch] = inimg ? fmaxf(acc[ch], 0.f) : 0.f;
            *reinterpret_cast<uint4*>(dst0) = pack8h(acc);
            *reinterpret_cast<uint4*>(dst1) = pack8h(acc + 8);
          };
          if (pl < G::WP) {
            const int ww = cb + pl;
            const bool col_ok = ww >= 0 && ww < W;
            float w[3][3];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
              for (int q = 0; q < 3; ++q) w[r + 1][q] = __half2float(sp[(r + 1) * G::SPW + pl + q]);   // tile rows 1, 2
#pragma unroll 1
            for (int xr = 0; xr < G::S + 2; ++xr) {     // X local row xr = image row a+1+xr = tile rows xr+1 .. xr+3
#pragma unroll
              for (int q = 0; q < 3; ++q) { w[0][q] = w[1][q]; w[1][q] = w[2][q]; w[2][q] = __half2float(sp[(xr + 3) * G::SPW + pl + q]); }
              const int hh = a + 1 + xr;
              unsigned char* dst = smem + G::OFF_X + (xr * G::WP + pl) * 16;
              conv_px(w, col_ok && hh >= 0 && hh < H, dst, dst + G::LBO_X);
            }
          }
          // residual input: X at image (a + 2i, 2(wo0 + j)) = tile rows 2i .. 2i+2, tile columns 2+2j .. 4+2j
          for (int q = pl; q < G::RQ; q += 64) {
            const int i = q / G::CP, j = q - i * G::CP;
            const int hh = a + 2 * i, wo = wo0 + j;
            float w[3][3];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int c2 = 0; c2 < 3; ++c2) w[r][c2] = __half2float(sp[(2 * i + r) * G::SPW + 2 + 2 * j + c2]);
            unsigned char* dst = smem + G::OFF_R + q * 16;
            conv_px(w, hh >= 0 && hh < H && 2 * wo < W, dst, dst + G::LBO_R);
          }
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[G::B_X]);
        }
      }
    } else {
    // =============================== TMA producer ===============================
    long long g = 0;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const long long b = item / n_strips;
      const int strip = (int)(item - b * n_strips);
      const int wo0 = strip * G::CP, cb = 2 * wo0 - 2;
      for (int step = 0; step < n_steps; ++step, ++g) {
        const int a = step * G::S - 2;
        // X and R are free once the first convolution (and residual MMA) of the step that used the buffer has completed
        const int xb = G::XBUF == 2 ? (int)(g & 1) : 0;
        if (G::XBUF == 1) {
          if (g > 0) {
            if (G::N1A < G::N1) mbar_wait(&bars[G::B_1 + G::N1A - 1], (uint32_t)((g - 1) & 1));
            mbar_wait(&bars[G::B_1 + G::N1 - 1], (uint32_t)((g - 1) & 1));
          }
        } else if (g >= 2) {
          mbar_wait(&bars[G::B_XF + xb], (uint32_t)(((g >> 1) - 1) & 1));
        }
        FB_TRACE(5, g);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars[G::B_X + xb], G::TX_BYTES);
#pragma unroll
          for (int c = 0; c < G::XG; ++c) {
            if constexpr (G::PREC) {
              // chunk-planar (hi, lo) tensors, maps {8 ch, w, h, image, chunk}: a box row is one contiguous run (an NHWC tensor would
              // make every pixel of the box its own 16-byte element: the TMA unit then needs 6 k cycles per step)
              tma_load_5d(sbase + G::OFF_X + xb * G::XR_BYTES + c * G::LBO_X, &tmX, &bars[G::B_X + xb], 0, cb, a + 1, (int)b, c);
              tma_load_5d(sbase + G::OFF_R + xb * G::XR_BYTES + c * G::LBO_R, &tmR, &bars[G::B_X + xb], 0, wo0 * r_step, (a >> 1) * r_step, (int)b, c);
              tma_load_5d(sbase + G::OFF_X + G::X_LO + xb * G::XR_BYTES + c * G::LBO_X, &tmXl, &bars[G::B_X + xb], 0, cb, a + 1, (int)b, c);
              tma_load_5d(sbase + G::OFF_R + G::R_LO + xb * G::XR_BYTES + c * G::LBO_R, &tmRl, &bars[G::B_X + xb], 0, wo0 * r_step, (a >> 1) * r_step, (int)b, c);
            } else {
            tma_load_5d(sbase + G::OFF_X + xb * G::XR_BYTES + c * G::LBO_X, &tmX, &bars[G::B_X + xb], 0, c, cb, a + 1, (int)b);
            // residual input pixels (2*ho, 2*wo): coordinates in the even-position tensor, or in the full tensor read with stride 2
            tma_load_5d(sbase + G::OFF_R + xb * G::XR_BYTES + c * G::LBO_R, &tmR, &bars[G::B_X + xb], 0, c, wo0 * r_step, (a >> 1) * r_step, (int)b);
            }
          }
        }
        __syncwarp();
      }
    }
    }
  } else if (warp >= G::NEW) {
    // =============================== MMA issuer(s) ===============================
    // The whole warp runs the control flow; one elected lane issues (elect.sync keeps the tensor-core instructions in
    // warp-uniform code, so ptxas emits them back to back instead of wrapping each one in a per-lane loop).
    if (total_steps > 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, G::NP, 0);
      // base descriptors; per-MMA descriptors differ only in the start-address field (16-byte units, no carry out of it)
      const uint64_t dX = make_smem_desc(sbase + G::OFF_X, G::LBO_X, 128);
      const uint64_t dS1 = make_smem_desc(sbase + G::OFF_S1, G::LBO_S1, 128);
      const uint64_t dR = make_smem_desc(sbase + G::OFF_R, G::LBO_R, 128);
      const uint64_t dW1 = make_smem_desc(sbase + G::OFF_W1, 128, G::SBO_W1);
      const uint64_t dW2 = make_smem_desc(sbase + G::OFF_W2, 128, G::SBO_W2);
      const uint64_t dWR = make_smem_desc(sbase + G::OFF_WR, 128, G::SBO_W1);
      const uint64_t dOnes = make_smem_desc(sbase + G::OFF_ONES, 128, 0);       // every 8-row group reads the same core matrix
      const uint64_t dB1 = make_smem_desc(sbase + G::OFF_WB1, 128, 128);
      const uint64_t dB2 = make_smem_desc(sbase + G::OFF_WB2, 128, 128);
      const uint64_t dBR = make_smem_desc(sbase + G::OFF_WBR, 128, 128);
      // one operand product; PREC: A_hi*W_hi + A_lo*W_hi + A_hi*W_lo (a_lo / G::W_LO: hi -> lo plane set, 16-byte units)
      auto mma_x = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t a_lo) {
        mma_f16_ss(d, a, b, idesc, 1);
        if constexpr (G::PREC) {
          mma_f16_ss(d, a + a_lo, b, idesc, 1);
          mma_f16_ss(d, a, b + (G::W_LO >> 4), idesc, 1);
        }
      };
      if constexpr (G::ISS == 1) {
        auto issue_first = [&](long long g) {   // residual 1x1 and first separable convolution of step g
          if (elect_one()) {
            const uint32_t colr = tmem + G::COL_R + (uint32_t)(g % G::RB) * G::NP;
            mma_f16_ss(colr, dOnes, dBR, idesc, 0);
#pragma unroll
            for (int ks = 0; ks < G::KP1 / 16; ++ks)
              mma_x(colr, dR + ((2 * ks * G::LBO_R) >> 4), dWR + ((2 * ks * 128) >> 4), G::R_LO >> 4);
            mma_commit(&bars[G::B_R + (int)(g % G::RB)]);
#pragma unroll
            for (int t = 0; t < G::N1; ++t) {
              mma_f16_ss(tmem + G::COL_1 + t * G::NP, dOnes, dB1, idesc, 0);
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const uint32_t aoff = (uint32_t)(G::P1_0 + 128 * t - 2 * G::WP - 1 + (tap / 3) * G::WP + (tap % 3));   // pixels = 16-byte units
#pragma unroll
                for (int ks = 0; ks < G::KP1 / 16; ++ks)
                  mma_x(tmem + G::COL_1 + t * G::NP, dX + aoff + ((2 * ks * G::LBO_X) >> 4),
                        dW1 + ((tap * G::TAP_W1 + 2 * ks * 128) >> 4), G::X_LO >> 4);
              }
              mma_commit(&bars[G::B_1 + t]);
            }
          }
          __syncwarp();
        };
        mbar_wait(&bars[G::B_X], 0);
        tc_fence_after();
        issue_first(0);
        for (long long g = 0; g < total_steps; ++g) {
          const uint32_t par = (uint32_t)(g & 1);
#pragma unroll
          for (int t = 0; t < G::N2; ++t) {
            // tile t of the second convolution reads S1 tiles <= t+1 (and the carried rows, published with tile 0)
            if (t == 0) mbar_wait(&bars[G::B_S1], par);
            if (t + 1 < G::N1) mbar_wait(&bars[G::B_S1 + t + 1], par);
            tc_fence_after();
            if (elect_one()) {
              mma_f16_ss(tmem + G::COL_2 + t * G::NP, dOnes, dB2, idesc, 0);
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const uint32_t aoff = (uint32_t)(G::P2_0 + 128 * t - G::WP - 1 + (tap / 3) * G::WP + (tap % 3));
#pragma unroll
                for (int ks = 0; ks < G::NP / 16; ++ks)
                  mma_x(tmem + G::COL_2 + t * G::NP, dS1 + aoff + ((2 * ks * G::LBO_S1) >> 4),
                        dW2 + ((tap * G::TAP_W2 + 2 * ks * 128) >> 4), G::S1_LO >> 4);
              }
              mma_commit(&bars[G::B_2 + t]);
            }
            __syncwarp();
          }
          if (g + 1 < total_steps) {
            if (g + 1 >= G::RB) mbar_wait(&bars[G::B_P + (int)((g + 1) % G::RB)], (uint32_t)(((g + 1) / G::RB - 1) & 1));   // the pooling epilogue that read the residual buffer step g+1 reuses is done
            mbar_wait(&bars[G::B_X], par ^ 1);
            tc_fence_after();
            issue_first(g + 1);
          }
        }
      } else if (warp == G::NEW) {
        // ---- issuer A: residual 1x1 + first separable convolution of every step ----
        // Its own stream no longer follows the second convolution of the previous step, so it waits explicitly until the
        // workers have drained the accumulator tile it is about to overwrite (s1_full[t] of step g-1).
        for (long long g = 0; g < total_steps; ++g) {
          // three residual accumulators (UF2): this step no longer waits for the pooling epilogue of step g-2 (which the workers run
          // AFTER their depthwise pass of step g-1), only for that depthwise pass - the tensor pipe's operand reads and the pass's
          // window loads slow each other down when they overlap (measured: 14.9 against 13.3 ms per hour of audio), the pooling
          // epilogue, the carries and the second epilogue are light on shared memory and run beside this step's MMAs
          if (G::UF2 && G::RB == 3 && g >= 1) mbar_wait(&bars[G::B_D2], (uint32_t)((g - 1) & 1));
          if (g >= G::RB) mbar_wait(&bars[G::B_P + (int)(g % G::RB)], (uint32_t)((g / G::RB - 1) & 1));   // pooling of step g-RB has read the residual buffer step g reuses
          const int xb = G::XBUF == 2 ? (int)(g & 1) : 0;
          mbar_wait(&bars[G::B_X + xb], (uint32_t)((G::XBUF == 2 ? (g >> 1) : g) & 1));
          tc_fence_after();
          FB_TRACE(1, g);
          const uint64_t dXb = dX + ((xb * G::XR_BYTES) >> 4), dRb = dR + ((xb * G::XR_BYTES) >> 4);
          if (elect_one()) {
            const uint32_t colr = tmem + G::COL_R + (uint32_t)(g % G::RB) * G::NP;
            mma_f16_ss(colr, dOnes, dBR, idesc, 0);
#pragma unroll
            for (int ks = 0; ks < G::KP1 / 16; ++ks)
              mma_x(colr, dRb + ((2 * ks * G::LBO_R) >> 4), dWR + ((2 * ks * 128) >> 4), G::R_LO >> 4);
            mma_commit(&bars[G::B_R + (int)(g % G::RB)]);
          }
          __syncwarp();
#pragma unroll
          for (int t = 0; t < G::N1A; ++t) {
            if (g > 0) {
              mbar_wait(&bars[G::B_S1 + t], (uint32_t)((g - 1) & 1));
              tc_fence_after();
            }
            if (elect_one()) {
              mma_f16_ss(tmem + G::COL_1 + t * G::NP, dOnes, dB1, idesc, 0);
#pragma unroll
              for (int tap = 0; tap < (G::UF2 ? ORCAI_EXP_TAPS : 9); ++tap) {
                const uint32_t aoff = (uint32_t)(G::P1_0 + 128 * t - 2 * G::WP - 1 + (tap / 3) * G::WP + (tap % 3));   // pixels = 16-byte units
#pragma unroll
                for (int ks = 0; ks < G::KP1 / 16; ++ks)
                  mma_x(tmem + G::COL_1 + t * G::NP, dXb + aoff + ((2 * ks * G::LBO_X) >> 4),
                        dW1 + ((tap * G::TAP_W1 + 2 * ks * 128) >> 4), G::X_LO >> 4);
              }
              mma_commit(&bars[G::B_1 + t]);
              if (G::XBUF == 2 && t == G::N1A - 1) mma_commit(&bars[G::B_XF + xb]);   // this step's X / R buffer may be reloaded
            }
            __syncwarp();
          }
          FB_TRACE(2, g);
        }
      } else {
        // ---- issuer B: second separable convolution of every step ----
        for (long long g = 0; g < total_steps; ++g) {
          const uint32_t par = (uint32_t)(g & 1);
          FB_TRACE(3, g);
          if constexpr (G::N1A < G::N1) {
            // this warp's share of the first convolution (it has seen the depthwise pass of step g-1 complete in its last trip)
            const int xb = G::XBUF == 2 ? (int)(g & 1) : 0;
            mbar_wait(&bars[G::B_X + xb], (uint32_t)((G::XBUF == 2 ? (g >> 1) : g) & 1));
            tc_fence_after();
            const uint64_t dXb = dX + ((xb * G::XR_BYTES) >> 4);
#pragma unroll
            for (int t = G::N1A; t < G::N1; ++t) {
              if (g > 0) {
                mbar_wait(&bars[G::B_S1 + t], (uint32_t)((g - 1) & 1));
                tc_fence_after();
              }
              if (elect_one()) {
                mma_f16_ss(tmem + G::COL_1 + t * G::NP, dOnes, dB1, idesc, 0);
#pragma unroll
                for (int tap = 0; tap < ORCAI_EXP_TAPS; ++tap) {
                  const uint32_t aoff = (uint32_t)(G::P1_0 + 128 * t - 2 * G::WP - 1 + (tap / 3) * G::WP + (tap % 3));
#pragma unroll
                  for (int ks = 0; ks < G::KP1 / 16; ++ks)
                    mma_x(tmem + G::COL_1 + t * G::NP, dXb + aoff + ((2 * ks * G::LBO_X) >> 4),
                          dW1 + ((tap * G::TAP_W1 + 2 * ks * 128) >> 4), G::X_LO >> 4);
                }
                mma_commit(&bars[G::B_1 + t]);
                if (G::XBUF == 2 && t == G::N1 - 1) mma_commit(&bars[G::B_XF + xb]);
              }
              __syncwarp();
            }
          }
          if constexpr (G::UF2) {
            // un-folded: the workers have written D2 = depthwise(S1) as (hi, lo) rows in accumulator order; pointwise 1x1 only
            const uint64_t dD2 = make_smem_desc(sbase + G::OFF_D2, G::LBO_D2, 128);
            mbar_wait(&bars[G::B_D2], par);
            tc_fence_after();
#pragma unroll
            for (int t = 0; t < G::N2; ++t) {
              if (elect_one()) {
                mma_f16_ss(tmem + G::COL_2 + t * G::NP, dOnes, dB2, idesc, 0);
#pragma unroll
                for (int ks = 0; ks < G::NP / 16; ++ks)
                  mma_x(tmem + G::COL_2 + t * G::NP, dD2 + (uint32_t)(128 * t) + ((2 * ks * G::LBO_D2) >> 4), dW2 + ((2 * ks * 128) >> 4), G::D2_LO >> 4);
                mma_commit(&bars[G::B_2 + t]);
              }
              __syncwarp();
            }
            FB_TRACE(4, g);
            continue;
          }
#pragma unroll
          for (int t = 0; t < G::N2; ++t) {
            // tile t of the second convolution reads S1 tiles <= t+1 (and the carried rows, published with tile 0)
            if (t == 0) mbar_wait(&bars[G::B_S1], par);
            if (t + 1 < G::N1) mbar_wait(&bars[G::B_S1 + t + 1], par);
            tc_fence_after();
            if (elect_one()) {
              mma_f16_ss(tmem + G::COL_2 + t * G::NP, dOnes, dB2, idesc, 0);
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const uint32_t aoff = (uint32_t)(G::P2_0 + 128 * t - G::WP - 1 + (tap / 3) * G::WP + (tap % 3));
#pragma unroll
                for (int ks = 0; ks < G::NP / 16; ++ks)
                  mma_x(tmem + G::COL_2 + t * G::NP, dS1 + aoff + ((2 * ks * G::LBO_S1) >> 4),
                        dW2 + ((tap * G::TAP_W2 + 2 * ks * 128) >> 4), G::S1_LO >> 4);
              }
              mma_commit(&bars[G::B_2 + t]);
            }
            __syncwarp();
          }
          FB_TRACE(4, g);
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== workers ===============================
    const int quad = warp & 3, team = warp >> 2;
    const int row = quad * 32 + lane;              // accumulator row (TMEM lane) this thread drains
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    // a team drains two channel groups (16 accumulator columns), or one when there are as many teams as groups
    constexpr int kGPT = G::NT >= G::NG ? 1 : 2;
    const int g0 = team * kGPT;                    // first of the (up to) two channel groups of this team
    const bool has0 = g0 < G::NG, has1 = kGPT == 2 && g0 + 1 < G::NG;
    const int q_i = row / G::CP, q_j = row - q_i * G::CP;   // pooled pixel of the pool / residual epilogue
    // tile-invariant pixel coordinates of the accumulator rows this thread drains
    int y1[G::N1], c1[G::N1], y2[G::N2], c2[G::N2];
#pragma unroll
    for (int t = 0; t < G::N1; ++t) { const int p = G::P1_0 + 128 * t + row; y1[t] = p / G::WP; c1[t] = p - y1[t] * G::WP; }
#pragma unroll
    for (int t = 0; t < G::N2; ++t) {
      if constexpr (G::UF2) {   // tile t = the columns of parity t: accumulator row = (y - 1) * DWP + (column pair - 1)
        y2[t] = row / G::DWP + 1; c2[t] = 2 * (row - (y2[t] - 1) * G::DWP + 1) + t;
      } else {
        const int p = G::P2_0 + 128 * t + row; y2[t] = p / G::WP; c2[t] = p - y2[t] * G::WP;
      }
    }

    // pooling epilogue of global step gp: max-pool (3,2)/2 + residual add (+ ReLU) -> global
    auto pool_store = [&](long long gp, long long pb, int pwo0, int pa) {
      mbar_wait(&bars[G::B_R + (int)(gp % G::RB)], (uint32_t)((gp / G::RB) & 1));
      tc_fence_after();
      if (has0) {
        float r[16];
        tmem_ld16f(lane_addr + G::COL_R + (uint32_t)(gp % G::RB) * G::NP + g0 * 8, r);
        const int ho = (pa >> 1) + q_i, wo = pwo0 + q_j;
        // output row in the output tensor: image pb's own rows, or (tall view) the window's rows past its warm-up
        const long long orow = tall.on ? (pb * tall.stride - tall.warm) / 2 + ho : pb * Ho + ho;
        const bool row_ok = tall.on ? (ho >= tall.warm / 2 && ho < Ho && orow < tall.rows / 2) : (ho >= 0 && ho < Ho);
        if (row < G::RQ && row_ok && wo < Wo && !ORCAI_EXP_NOPOOL) {
          const uint32_t p00 = (uint32_t)((2 * q_i) * (G::WP / 2) + 1 + q_j);   // even column 2 + 2j of row 2i; the odd column 3 + 2j sits S2HALF further
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (u == 1 && !has1) break;
            const int gg = g0 + u;
            const unsigned char* s2 = smem + G::OFF_S2 + gg * G::LBO_S2 + p00 * 16;
            const size_t o_full = ((size_t)orow * Wo + wo) * G::OCP + gg * 8;
            const size_t o_sub = (((size_t)pb * Hs + (ho >> 1)) * Ws + (wo >> 1)) * G::OCP + gg * 8;
            const bool sub = Ysub != nullptr && !(ho & 1) && !(wo & 1);
            if constexpr (G::PREC) {
              float y[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) y[i] = -INFINITY;
#pragma unroll
              for (int q = 0; q < 6; ++q) {
                const uint32_t off = (uint32_t)((q >> 1) * (G::WP / 2) + (q & 1) * G::S2HALF) * 16;
                if constexpr (G::UF2) {
                  // fp32 S2, one plane per channel quad: quads 2 gg, 2 gg + 1
                  const float4 a = *reinterpret_cast<const float4*>(s2 + (size_t)gg * G::LBO_S2 + off);
                  const float4 b = *reinterpret_cast<const float4*>(s2 + (size_t)(gg + 1) * G::LBO_S2 + off);
                  y[0] = fmaxf(y[0], a.x); y[1] = fmaxf(y[1], a.y); y[2] = fmaxf(y[2], a.z); y[3] = fmaxf(y[3], a.w);
                  y[4] = fmaxf(y[4], b.x); y[5] = fmaxf(y[5], b.y); y[6] = fmaxf(y[6], b.z); y[7] = fmaxf(y[7], b.w);
                } else {
                  fmax8_split(y, *reinterpret_cast<const uint4*>(s2 + off), *reinterpret_cast<const uint4*>(s2 + G::S2_LO + off));
                }
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) y[i] += r[8 * u + i];
              // PREC blocks hand fp32 tensors to the next stage (Yr / Ysub are float arrays of the same shape)
              float* yf = reinterpret_cast<float*>(Yr) + o_full;
              if (sub) {
                float* ys = reinterpret_cast<float*>(Ysub) + o_sub;
                st_global_v8(ys, y);
              }
              if (G::RELU_OUT) {
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = fmaxf(y[i], 0.f);
              }
              st_global_v8(yf, y);   // the pixel's 8-channel group = one whole 32-byte sector (OCP is a multiple of 8, the tensors 256-byte aligned)
            } else {
            uint4 m = *reinterpret_cast<const uint4*>(s2);
            m = hmax8(m, *reinterpret_cast<const uint4*>(s2 + G::S2HALF * 16));
            m = hmax8(m, *reinterpret_cast<const uint4*>(s2 + (G::WP / 2) * 16));
            m = hmax8(m, *reinterpret_cast<const uint4*>(s2 + (G::WP / 2 + G::S2HALF) * 16));
            m = hmax8(m, *reinterpret_cast<const uint4*>(s2 + G::WP * 16));
            m = hmax8(m, *reinterpret_cast<const uint4*>(s2 + (G::WP + G::S2HALF) * 16));
            const __half2* mh = reinterpret_cast<const __half2*>(&m);
            float y[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __half22float2(mh[i]);
              y[2 * i] = f.x + r[8 * u + 2 * i];
              y[2 * i + 1] = f.y + r[8 * u + 2 * i + 1];
            }
            const uint4 yp = pack8h(y);
            *reinterpret_cast<uint4*>(Yr + o_full) = G::RELU_OUT ? relu8h(yp) : yp;
            if (sub) *reinterpret_cast<uint4*>(Ysub + o_sub) = yp;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[G::B_P + (int)(gp % G::RB)]);
    };

    long long g = 0;
    long long prev_b = 0;
    int prev_wo0 = 0, prev_a = 0;
    bool prev_carry = false;
    uint32_t n_c1 = 0, w_c1 = 0, n_c2 = 0, w_c2 = 0;   // bulk carries issued / awaited (S1, S2): the same count in every worker
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const long long b = item / n_strips;
      const int strip = (int)(item - b * n_strips);
      const int wo0 = strip * G::CP, cb = 2 * wo0 - 2;
      // rows that count as "inside the image": this image's [0, H), or (tall view) the tall image's [0, rows) seen from window b
      const long long img_row0 = tall.on ? b * tall.stride - tall.warm : 0, img_rows = tall.on ? tall.rows : H;
      for (int step = 0; step < n_steps; ++step, ++g) {
        const int a = step * G::S - 2;
        const uint32_t par = (uint32_t)(g & 1);
        // ---- epilogue 1: ReLU, zero outside the image ("same" padding of the second convolution) -> S1 ----
#pragma unroll
        for (int t = 0; t < G::N1; ++t) {
          mbar_wait(&bars[G::B_1 + t], par);
          tc_fence_after();
          if (G::PREC && t == 0 && n_c1 > w_c1) { mbar_wait(&bars[G::B_C1], (uint32_t)(w_c1 & 1)); ++w_c1; }   // carried S1 rows in place
          if (warp == 0) FB_TRACE(10 + t, g);
          if (has0) {
            float v[16];
            tmem_ld16f(lane_addr + G::COL_1 + t * G::NP + g0 * 8, v);
            const int p1 = G::P1_0 + 128 * t + row;
            const long long hh = img_row0 + a + y1[t];
            const int ww = cb + c1[t];
            const bool inimg = hh >= 0 && hh < img_rows && ww >= 0 && ww < W;
            if (p1 < (G::S + 2) * G::WP) {
              const uint4 z = make_uint4(0, 0, 0, 0);
              unsigned char* dst = smem + G::OFF_S1 + g0 * G::LBO_S1 + p1 * 16;
              if constexpr (G::UF2) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = inimg ? fmaxf(v[i], 0.f) : 0.f;
                unsigned char* dq = smem + G::OFF_S1 + (2 * g0) * G::LBO_S1F + ((c1[t] & 1) * G::S1HALF + y1[t] * (G::WP / 2) + (c1[t] >> 1)) * 16;     // quads 2 g0 .. 2 g0 + 3
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (q < 2 || has1) *reinterpret_cast<float4*>(dq + q * G::LBO_S1F) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
              } else if constexpr (G::PREC) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = inimg ? fmaxf(v[i], 0.f) : 0.f;
                uint4 hi, lo;
                split8h(v, hi, lo);
                *reinterpret_cast<uint4*>(dst) = hi;
                *reinterpret_cast<uint4*>(dst + G::S1_LO) = lo;
                if (has1) {
                  split8h(v + 8, hi, lo);
                  *reinterpret_cast<uint4*>(dst + G::LBO_S1) = hi;
                  *reinterpret_cast<uint4*>(dst + G::LBO_S1 + G::S1_LO) = lo;
                }
              } else {
              *reinterpret_cast<uint4*>(dst) = inimg ? relu8h(pack8h(v)) : z;
              if (has1) *reinterpret_cast<uint4*>(dst + G::LBO_S1) = inimg ? relu8h(pack8h(v + 8)) : z;
              }
            }
          }
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[G::B_S1 + t]);
        }
        if constexpr (G::UF2) {
          // ---- depthwise 3x3 of the second separable convolution on the CUDA cores: S1 (fp32) -> D2 (hi, lo) ----
          worker_sync<G::NWORK>();   // every S1 pixel of this step and the carried rows are in place
          // thread = (column pair jj: columns 2 jj, 2 jj + 1; channel quad q); lanes 2 i, 2 i + 1 take quads (q, q ^ 1) of the same pair, so
          // that their 8-byte D2 stores fill one 16-byte row piece and a warp's stores are contiguous
          if (tid < (ORCAI_EXP_NODW ? 0 : G::NQ * G::DWP)) {
            const int pr = tid >> 1, qp = pr / G::DWP, jj = pr - qp * G::DWP + 1, q = 2 * qp + (tid & 1);
            const float4* kw = reinterpret_cast<const float4*>(smem + G::OFF_DW2) + q;
            float4 k[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) k[t] = kw[t * (G::NP / 4)];
            const unsigned char* s1e = smem + G::OFF_S1 + q * G::LBO_S1F + jj * 16;   // even half: column 2 jj of row 0
            const unsigned char* s1o = s1e + G::S1HALF * 16;                           // odd half: column 2 jj + 1 of row 0
            // window columns 2 jj - 1 .. 2 jj + 2 = odd[jj - 1], even[jj], odd[jj], even[jj + 1]
            float4 w[3][4];
            auto ld_row = [&](int y, float4 (&d)[4]) {
              const int o = y * (G::WP / 2) * 16;
              d[0] = *reinterpret_cast<const float4*>(s1o + o - 16);
              d[1] = *reinterpret_cast<const float4*>(s1e + o);
              d[2] = *reinterpret_cast<const float4*>(s1o + o);
              d[3] = *reinterpret_cast<const float4*>(s1e + o + 16);
            };
            ld_row(0, w[1]);
            ld_row(1, w[2]);
            unsigned char* d2 = smem + G::OFF_D2 + qp * G::LBO_D2 + (jj - 1) * 16 + (tid & 1) * 8;
#pragma unroll
            for (int r = 1; r <= G::S; ++r) {                        // S2 row r <- S1 rows r-1 .. r+1 (unrolled: the window rotates by renaming)
#pragma unroll
              for (int x = 0; x < 4; ++x) { w[0][x] = w[1][x]; w[1][x] = w[2][x]; }
              ld_row(r + 1, w[2]);
#pragma unroll
              for (int px = 0; px < 2; ++px) {                       // px = column parity = accumulator tile
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                  const float4 a = w[t / 3][t % 3 + px];
                  fma2(acc.x, acc.y, a.x, a.y, k[t].x, k[t].y);
                  fma2(acc.z, acc.w, a.z, a.w, k[t].z, k[t].w);
                }
                const __half2 h01 = __floats2half2_rn(acc.x, acc.y), h23 = __floats2half2_rn(acc.z, acc.w);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 l01 = __floats2half2_rn(acc.x - f01.x, acc.y - f01.y), l23 = __floats2half2_rn(acc.z - f23.x, acc.w - f23.y);
                unsigned char* dst = d2 + (px * 128 + (r - 1) * G::DWP) * 16;
                uint2 hv, lv;
                hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                *reinterpret_cast<uint2*>(dst) = hv;
                *reinterpret_cast<uint2*>(dst + G::D2_LO) = lv;
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[G::B_D2]);
          if (warp == 0) FB_TRACE(23, g);
        }
        // ---- previous step: pool + store while the tensor pipe runs this step's second convolution ----
        if (warp == 0) FB_TRACE(20, g);
        if (g > 0) {
          pool_store(g - 1, prev_b, prev_wo0, prev_a);
          if (warp == 0) FB_TRACE(21, g);
          worker_sync<G::NWORK>();   // pooling has finished reading S2
          if (G::PREC && kAsyncCarry && prev_carry) {
            // as for S1: one row per plane half (even / odd columns), copied by the TMA unit; awaited before epilogue 2 stores
            if (tid == 0) {
              fence_proxy_async();
              mbar_arrive_expect_tx(&bars[G::B_C2], (uint32_t)(G::PL * G::NG * 2 * (G::WP / 2) * 16));
#pragma unroll
              for (int gq = 0; gq < G::PL * G::NG; ++gq)
#pragma unroll
                for (int hp = 0; hp < 2; ++hp) {
                  const uint32_t d = sbase + G::OFF_S2 + gq * G::LBO_S2 + hp * G::S2HALF * 16;
                  bulk_copy_s2s(d, d + G::S * (G::WP / 2) * 16, (G::WP / 2) * 16, &bars[G::B_C2]);
                }
            }
            ++n_c2;
          } else if (prev_carry) {
            // all loads of a thread first, then its stores (one shared-memory round trip instead of one per element)
            constexpr int kN2 = G::PL * G::NG * G::WP, kIt2 = (kN2 + G::NWORK - 1) / G::NWORK;
            unsigned char* cp2[kIt2];
            uint4 cv2[kIt2];
#pragma unroll
            for (int it = 0; it < kIt2; ++it) {
              const int i = tid + it * G::NWORK;                             // the lo plane set follows the hi set
              const int gq = i / G::WP, px = i - gq * G::WP;
              const int hp = px / (G::WP / 2), cc = px - hp * (G::WP / 2);   // half-plane (column parity), column / 2
              cp2[it] = smem + G::OFF_S2 + gq * G::LBO_S2 + (hp * G::S2HALF + cc) * 16;
              if (i < kN2) cv2[it] = *reinterpret_cast<const uint4*>(cp2[it] + G::S * (G::WP / 2) * 16);
            }
#pragma unroll
            for (int it = 0; it < kIt2; ++it)
              if (tid + it * G::NWORK < kN2) *reinterpret_cast<uint4*>(cp2[it]) = cv2[it];
            worker_sync<G::NWORK>();  // carried S2 row in place before epilogue 2 overwrites its source row
          }
        }
        // ---- epilogue 2: -inf outside the image (TF "same" max-pool padding) -> S2 ----
#pragma unroll
        for (int t = 0; t < G::N2; ++t) {
          if (warp == 0 && t == 0) FB_TRACE(22, g);
          mbar_wait(&bars[G::B_2 + t], par);
          tc_fence_after();
          if (G::PREC && t == 0 && n_c2 > w_c2) { mbar_wait(&bars[G::B_C2], (uint32_t)(w_c2 & 1)); ++w_c2; }   // carried S2 row in place
          if (warp == 0) FB_TRACE(30 + t, g);
          if (has0) {
            float v[16];
            tmem_ld16f(lane_addr + G::COL_2 + t * G::NP + g0 * 8, v);
            const int p2 = G::P2_0 + 128 * t + row;
            const long long hh = img_row0 + a + y2[t];
            const int ww = cb + c2[t];
            const bool inimg = hh >= 0 && hh < img_rows && ww >= 0 && ww < W;
            if (G::UF2 ? row < G::S * G::DWP : p2 < (G::S + 1) * G::WP) {
              const uint4 ninf = make_uint4(0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u);
              unsigned char* dst = smem + G::OFF_S2 + g0 * G::LBO_S2 + ((c2[t] & 1) * G::S2HALF + y2[t] * (G::WP / 2) + (c2[t] >> 1)) * 16;
              if constexpr (G::UF2) {
                // S2 only feeds the pooling epilogue: fp32, one plane per channel quad (quads 2 g0 .. 2 g0 + 3), same pixel order
                unsigned char* dq = dst + (size_t)g0 * G::LBO_S2;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (q < 2 || has1)
                    *reinterpret_cast<float4*>(dq + q * G::LBO_S2) = inimg ? make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3])
                                                                           : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
              } else if constexpr (G::PREC) {
                const uint4 z = make_uint4(0, 0, 0, 0);
                uint4 hi, lo;
                split8h(v, hi, lo);
                *reinterpret_cast<uint4*>(dst) = inimg ? hi : ninf;
                *reinterpret_cast<uint4*>(dst + G::S2_LO) = inimg ? lo : z;
                if (has1) {
                  split8h(v + 8, hi, lo);
                  *reinterpret_cast<uint4*>(dst + G::LBO_S2) = inimg ? hi : ninf;
                  *reinterpret_cast<uint4*>(dst + G::LBO_S2 + G::S2_LO) = inimg ? lo : z;
                }
              } else {
              *reinterpret_cast<uint4*>(dst) = inimg ? pack8h(v) : ninf;
              if (has1) *reinterpret_cast<uint4*>(dst + G::LBO_S2) = inimg ? pack8h(v + 8) : ninf;
              }
            }
          }
        }
        tc_fence_before();
        if (G::PREC && kAsyncCarry) fence_proxy_async();   // this thread's S1 / S2 accesses are ordered before the bulk copies issued behind the barrier
        if (warp == 0) FB_TRACE(39, g);
        worker_sync<G::NWORK>();   // every worker has seen the second convolution complete: S1 is free, S2 is written
        if (warp == 0) FB_TRACE(40, g);
        // ---- carry the S1 overlap rows into the next step (rows above the next strip's first row are zero) ----
        const bool carry = step + 1 < n_steps;
        if (g + 1 < total_steps) {
          constexpr int kCarryPlanes = G::UF2 ? G::NQ : G::PL * G::NG;
          if (G::PREC && kAsyncCarry && carry) {
            // the TMA unit copies the two rows plane by plane while the workers go on; they wait for it (carry_s1) before the
            // next step's first epilogue overwrites the source rows.  (Copied by the workers - 16 KB through their registers, with
            // a barrier behind it - this was 12 % of their step.)
            if (tid == 0) {
              mbar_arrive_expect_tx(&bars[G::B_C1], (uint32_t)kCarryPlanes * 2 * G::WP * 16);
#pragma unroll
              for (int gs = 0; gs < kCarryPlanes; ++gs) {
                const int gq = G::UF2 ? gs : gs % G::NG + (gs / G::NG) * G::MCH;
                const uint32_t d = sbase + G::OFF_S1 + gq * G::LBO_S1;
                bulk_copy_s2s(d, d + G::S * G::WP * 16, 2 * G::WP * 16, &bars[G::B_C1]);
              }
            }
            ++n_c1;
          } else {
          constexpr int kN1 = kCarryPlanes * 2 * G::WP, kIt1 = (kN1 + G::NWORK - 1) / G::NWORK;
          unsigned char* cp1[kIt1];
          uint4 cv1[kIt1];
#pragma unroll
          for (int it = 0; it < kIt1; ++it) {
            const int i = tid + it * G::NWORK;
            const int gs = i / (2 * G::WP), px = i - gs * 2 * G::WP;
            const int gq = G::UF2 ? gs : gs % G::NG + (gs / G::NG) * G::MCH;   // lo planes sit MCH planes after the hi planes
            if constexpr (G::UF2) {   // rows S, S + 1 -> rows 0, 1 of either column-parity half: WP contiguous pixels per half
              const int hp = px / G::WP;
              cp1[it] = smem + G::OFF_S1 + gq * G::LBO_S1F + (hp * G::S1HALF + px - hp * G::WP) * 16;
              cv1[it] = (carry && i < kN1) ? *reinterpret_cast<const uint4*>(cp1[it] + G::S * (G::WP / 2) * 16) : make_uint4(0, 0, 0, 0);
              continue;
            }
            cp1[it] = smem + G::OFF_S1 + gq * G::LBO_S1 + px * 16;
            cv1[it] = (carry && i < kN1) ? *reinterpret_cast<const uint4*>(cp1[it] + G::S * G::WP * 16) : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int it = 0; it < kIt1; ++it)
            if (tid + it * G::NWORK < kN1) *reinterpret_cast<uint4*>(cp1[it]) = cv1[it];
          worker_sync<G::NWORK>();   // carried rows in place before epilogue 1 overwrites their source rows
          }
        }
        if (warp == 0) FB_TRACE(41, g);
        prev_b = b; prev_wo0 = wo0; prev_a = a; prev_carry = carry;
      }
    }
    if (g > 0) pool_store(g - 1, prev_b, prev_wo0, prev_a);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<G::TM_COLS>(tmem);
}

}  // namespace fused
