// N-widened variant of the fused residual-block kernel (net_fused.cuh), used for block 1 when "block1_path" = 1.
//
// Same graph, same shared-memory tiles (X, S1, S2, R), same producer, same pooling epilogue.  What changes is how the two 3x3
// convolutions reach the tensor pipe.  net_fused.cuh issues one M128 x N32 MMA per tap and K-chunk: 9 taps re-read the same
// 4 KB A tile, and at N = 32 an MMA costs 41 cycles of shared-memory operand traffic against a 16-cycle math floor
// (tools/microbench/mma_cost.cu).  Here the three dx taps are three COLUMN BLOCKS of one MMA:
//     Z[q, (dx, n)] = sum_dy sum_k A[q + dy * WP][k] * W'[dy][dx][k][n]              one MMA per (dy, K-chunk), N = 3 * 32 = 96
//     out[p, n]     = Z[p - 1, (0, n)] + Z[p, (1, n)] + Z[p + 1, (2, n)]              warp shuffles in the epilogue
// N = 96 costs 56 cycles where three N = 32 MMAs cost 124.  So that the shuffles never cross a warp (= a TMEM lane quadrant),
// the A descriptor's 8-row groups start every SIX pixels (SBO = 96 B): group g holds pixels 6g .. 6g + 7, rows 1 .. 6 of a group
// find both neighbours inside the group.  An M = 128 tile therefore yields 96 output pixels (tools/microbench/nwide_conv.cu
// checks this building block on its own).  Accumulator tiles are three times as wide, so they no longer fit per step: two
// 96-column tiles are recycled tile by tile (full / free barriers) next to the residual convolution's two 32-column tiles.
#pragma once

namespace fused {

template <int CIN_, int COUT_, int CPOOL_, int S_, bool RELU_OUT_, int CTAS_, int NEW_>
struct FBW {
  static constexpr int CIN = CIN_, COUT = COUT_, CP = CPOOL_, S = S_, CTAS = CTAS_, NEW = NEW_;
  static constexpr bool RELU_OUT = RELU_OUT_;
  static constexpr int NWORK = NEW * 32, NTHREADS = NWORK + 64;   // + issuer warp + producer warp
  static constexpr int ICP = cpad8(CIN), OCP = cpad8(COUT);
  static constexpr int KP1 = cpad16(CIN);          // K per dy of sepconv 1 and of the residual convolution
  static constexpr int NP = cpad16(COUT);          // output channels per column block; K per dy of sepconv 2
  static constexpr int NZ = 3 * NP;                // N of the convolution MMAs: (dx, n)
  static constexpr int XG = ICP / 8, XCH = KP1 / 8, NG = OCP / 8, MCH = NP / 8;
  static constexpr int WP = 2 * CP + 4;
  static constexpr int TPX = 96;                   // output pixels per accumulator tile
  static constexpr int N1 = (S * WP - 2 + TPX - 1) / TPX, N2 = (S * WP - 4 + TPX - 1) / TPX;
  static constexpr int P1_0 = 2 * WP + 1, P2_0 = WP + 2;
  // MMA row r = 8 g + i reads pixel (tile origin - 1) + 6 g + i + tap offset: the last row of a tile is 97 pixels past its origin
  static constexpr int XPIX = round8(imax((S + 2) * WP, P1_0 + TPX * (N1 - 1) + 98));
  static constexpr int S1PIX = round8(imax((S + 2) * WP, P2_0 + TPX * (N2 - 1) + WP + 98));
  static constexpr int S2HALF = round8((S + 1) * (WP / 2)) + 4;
  static constexpr int S2PIX = 2 * S2HALF;
  static constexpr int RQ = (S / 2) * CP, RPIX = round8(RQ);
  static constexpr uint32_t LBO_X = XPIX * 16, LBO_S1 = S1PIX * 16, LBO_S2 = S2PIX * 16, LBO_R = RPIX * 16;
  // weights: per dy one B operand of NZ rows (dx, n) x K; canonical layout, SBO = (K / 8) * 128
  static constexpr uint32_t SBO_W1 = XCH * 128, SBO_W2 = MCH * 128;
  static constexpr uint32_t DY_W1 = (NZ / 8) * SBO_W1, DY_W2 = (NZ / 8) * SBO_W2;
  static constexpr uint32_t W1_BYTES = 3 * DY_W1, W2_BYTES = 3 * DY_W2, WR_BYTES = (NP / 8) * SBO_W1;
  static constexpr uint32_t WBZ_BYTES = (NZ / 8) * 128, WBR_BYTES = (NP / 8) * 128;   // bias rows [hi, lo]: one k-chunk
  static constexpr uint32_t OFF_W1 = 0, OFF_W2 = OFF_W1 + W1_BYTES, OFF_WR = OFF_W2 + W2_BYTES;
  static constexpr uint32_t OFF_WB1 = OFF_WR + WR_BYTES, OFF_WB2 = OFF_WB1 + WBZ_BYTES, OFF_WBR = OFF_WB2 + WBZ_BYTES;
  static constexpr uint32_t OFF_ONES = OFF_WBR + WBR_BYTES;
  static constexpr uint32_t W_BYTES = OFF_ONES + 256;
  static constexpr uint32_t OFF_R = W_BYTES;
  static constexpr uint32_t OFF_X = OFF_R + XCH * LBO_R;
  static constexpr uint32_t OFF_S1 = OFF_X + XCH * LBO_X;
  static constexpr uint32_t OFF_S2 = OFF_S1 + MCH * LBO_S1;
  static constexpr uint32_t OFF_BAR = OFF_S2 + NG * LBO_S2;
  // barriers: full[2] free[2] barR[2] s1_full[N1] pool_done x_full x_free
  static constexpr int B_FULL = 0, B_FREE = 2, B_R = 4, B_S1 = 6, B_P = 6 + N1, B_X = B_P + 1, B_XF = B_X + 1, NBAR = B_XF + 1;
  static constexpr uint32_t SMEM = OFF_BAR + NBAR * 8 + 16;
  static constexpr uint32_t TX_BYTES = XG * ((S + 2) * WP + (S / 2) * CP) * 16;
  static constexpr int COL_R = 0, COL_Z = 2 * NP;
  static constexpr int TM_COLS = pow2cols(2 * NP + 2 * NZ);
  static_assert(NEW == 8, "two worker teams of four warps (one per TMEM lane quadrant), 16 channels per team");
  static_assert(NP == 32 && XG == XCH && NG == MCH, "written for 32 output channels per block and unpadded K chunks");
  static_assert((N1 + N2) % 2 == 0, "the two accumulator tiles alternate with a period of one step");
  static_assert(S % 2 == 0 && RQ <= 128 && WP <= 94, "steps advance by whole pooled rows; a second-convolution tile depends on first-convolution tiles <= t+1");
  static_assert(TM_COLS * CTAS <= 512, "TMEM columns");
  static_assert(SMEM <= 227 * 1024 && (SMEM + 1024) * CTAS <= 228 * 1024, "shared memory (per CTA and per SM)");
  static_assert((128 - RPIX) * 16 <= XCH * LBO_X, "residual tile over-read must stay inside the CTA's shared memory");
};

template <class G>
__global__ void __launch_bounds__(G::NTHREADS, G::CTAS)
fused_block_w_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmR, int r_step, __half* __restrict__ Yr,
                     __half* __restrict__ Ysub, int H, int W, int n_strips, long long n_items, const unsigned char* __restrict__ wpack) {
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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[G::B_FULL + i], 1);        // tcgen05.commit
      mbar_init(&bars[G::B_FREE + i], G::NEW);   // every worker warp has loaded its part of the tile
      mbar_init(&bars[G::B_R + i], 1);
    }
    for (int i = G::B_S1; i <= G::B_P; ++i) mbar_init(&bars[i], G::NEW);
    mbar_init(&bars[G::B_X], 1);
    mbar_init(&bars[G::B_XF], 1);
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
  constexpr int TILES = G::N1 + G::N2;            // accumulator tiles per step; tile u of a step uses ring slot u & 1

  if (warp == G::NEW + 1) {
    // =============================== TMA producer (as in net_fused.cuh) ===============================
    long long g = 0;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const long long b = item / n_strips;
      const int strip = (int)(item - b * n_strips);
      const int wo0 = strip * G::CP, cb = 2 * wo0 - 2;
      for (int step = 0; step < n_steps; ++step, ++g) {
        const int a = step * G::S - 2;
        if (g > 0) mbar_wait(&bars[G::B_XF], (uint32_t)((g - 1) & 1));   // the first convolution and the residual MMA of step g-1 are done
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars[G::B_X], G::TX_BYTES);
#pragma unroll
          for (int c = 0; c < G::XG; ++c) {
            tma_load_5d(sbase + G::OFF_X + c * G::LBO_X, &tmX, &bars[G::B_X], 0, c, cb, a + 1, (int)b);
            tma_load_5d(sbase + G::OFF_R + c * G::LBO_R, &tmR, &bars[G::B_X], 0, c, wo0 * r_step, (a >> 1) * r_step, (int)b);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == G::NEW) {
    // =============================== MMA issuer ===============================
    if (total_steps > 0) {
      constexpr uint32_t idz = make_idesc_f16(128, G::NZ, 0), idr = make_idesc_f16(128, G::NP, 0);
      const uint64_t dX = make_smem_desc(sbase + G::OFF_X, G::LBO_X, 96);      // 8-row groups every 6 pixels
      const uint64_t dS1 = make_smem_desc(sbase + G::OFF_S1, G::LBO_S1, 96);
      const uint64_t dR = make_smem_desc(sbase + G::OFF_R, G::LBO_R, 128);
      const uint64_t dW1 = make_smem_desc(sbase + G::OFF_W1, 128, G::SBO_W1);
      const uint64_t dW2 = make_smem_desc(sbase + G::OFF_W2, 128, G::SBO_W2);
      const uint64_t dWR = make_smem_desc(sbase + G::OFF_WR, 128, G::SBO_W1);
      const uint64_t dOnes = make_smem_desc(sbase + G::OFF_ONES, 128, 0);
      const uint64_t dB1 = make_smem_desc(sbase + G::OFF_WB1, 128, 128);
      const uint64_t dB2 = make_smem_desc(sbase + G::OFF_WB2, 128, 128);
      const uint64_t dBR = make_smem_desc(sbase + G::OFF_WBR, 128, 128);
      long long seq = 0;   // accumulator tiles issued so far: slot = seq & 1, this is use (seq >> 1) of the slot
      auto acquire = [&]() -> uint32_t {
        const int slot = (int)(seq & 1);
        const long long use = seq >> 1;
        if (use > 0) {
          mbar_wait(&bars[G::B_FREE + slot], (uint32_t)((use - 1) & 1));
          tc_fence_after();
        }
        return tmem + G::COL_Z + (uint32_t)slot * G::NZ;
      };
      auto issue_first = [&](long long g) {
        if (elect_one()) {
          const uint32_t colr = tmem + G::COL_R + (uint32_t)(g & 1) * G::NP;
          mma_f16_ss(colr, dOnes, dBR, idr, 0);
#pragma unroll
          for (int ks = 0; ks < G::KP1 / 16; ++ks)
            mma_f16_ss(colr, dR + ((2 * ks * G::LBO_R) >> 4), dWR + ((2 * ks * 128) >> 4), idr, 1);
          mma_commit(&bars[G::B_R + (int)(g & 1)]);
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < G::N1; ++t) {
          const uint32_t z = acquire();
          if (elect_one()) {
            mma_f16_ss(z, dOnes, dB1, idz, 0);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const uint32_t aoff = (uint32_t)(G::P1_0 + G::TPX * t - 2 * G::WP - 1 + dy * G::WP);   // row 0 = left neighbour of the first output pixel
#pragma unroll
              for (int ks = 0; ks < G::KP1 / 16; ++ks)
                mma_f16_ss(z, dX + aoff + ((2 * ks * G::LBO_X) >> 4), dW1 + ((dy * G::DY_W1 + 2 * ks * 128) >> 4), idz, 1);
            }
            mma_commit(&bars[G::B_FULL + (int)(seq & 1)]);
            if (t == G::N1 - 1) mma_commit(&bars[G::B_XF]);   // X and R of this step may be reloaded
          }
          __syncwarp();
          ++seq;
        }
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
          const uint32_t z = acquire();
          if (elect_one()) {
            mma_f16_ss(z, dOnes, dB2, idz, 0);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const uint32_t aoff = (uint32_t)(G::P2_0 + G::TPX * t - G::WP - 1 + dy * G::WP);
#pragma unroll
              for (int ks = 0; ks < G::NP / 16; ++ks)
                mma_f16_ss(z, dS1 + aoff + ((2 * ks * G::LBO_S1) >> 4), dW2 + ((dy * G::DY_W2 + 2 * ks * 128) >> 4), idz, 1);
            }
            mma_commit(&bars[G::B_FULL + (int)(seq & 1)]);
          }
          __syncwarp();
          ++seq;
        }
        if (g + 1 < total_steps) {
          if (g >= 1) mbar_wait(&bars[G::B_P], par ^ 1);   // pooling of step g-1 has read the residual buffer step g+1 reuses
          mbar_wait(&bars[G::B_X], par ^ 1);
          tc_fence_after();
          issue_first(g + 1);
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== workers ===============================
    const int quad = warp & 3, team = warp >> 2;
    const int row = quad * 32 + lane;              // accumulator row (TMEM lane) this thread drains
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const int g0 = team * 2;                       // first of the two channel groups (16 accumulator columns) of this team
    const int q_i = row / G::CP, q_j = row - q_i * G::CP;   // pooled pixel of the pool / residual epilogue
    // a row is the output of pixel (tile origin + 6 * group + i - 1) when 1 <= i <= 6; rows 0 and 7 only serve as neighbours
    const int grp = row >> 3, gi = row & 7;
    const bool out_row = gi >= 1 && gi <= 6;
    const int opix = 6 * grp + gi - 1;
    int y1[G::N1], c1[G::N1], y2[G::N2], c2[G::N2];
#pragma unroll
    for (int t = 0; t < G::N1; ++t) { const int p = G::P1_0 + G::TPX * t + opix; y1[t] = p / G::WP; c1[t] = p - y1[t] * G::WP; }
#pragma unroll
    for (int t = 0; t < G::N2; ++t) { const int p = G::P2_0 + G::TPX * t + opix; y2[t] = p / G::WP; c2[t] = p - y2[t] * G::WP; }
    long long seq = 0;   // accumulator tiles drained so far (same order as the issuer)

    // drain one accumulator tile: the team's 16 channels of the three column blocks, combined across neighbouring rows
    auto drain = [&](float (&v)[16]) {
      const int slot = (int)(seq & 1);
      mbar_wait(&bars[G::B_FULL + slot], (uint32_t)((seq >> 1) & 1));
      tc_fence_after();
      const uint32_t z = lane_addr + G::COL_Z + (uint32_t)slot * G::NZ + g0 * 8;
      float zl[16], zr[16];
      tmem_ld16f(z, zl);
      tmem_ld16f(z + G::NP, v);
      tmem_ld16f(z + 2 * G::NP, zr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[G::B_FREE + slot]);
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] += __shfl_up_sync(0xffffffffu, zl[c], 1) + __shfl_down_sync(0xffffffffu, zr[c], 1);
      ++seq;
    };

    auto pool_store = [&](long long gp, long long pb, int pwo0, int pa) {
      mbar_wait(&bars[G::B_R + (int)(gp & 1)], (uint32_t)((gp >> 1) & 1));
      tc_fence_after();
      float r[16];
      tmem_ld16f(lane_addr + G::COL_R + (uint32_t)(gp & 1) * G::NP + g0 * 8, r);
      const int ho = (pa >> 1) + q_i, wo = pwo0 + q_j;
      if (row < G::RQ && ho >= 0 && ho < Ho && wo < Wo) {
        const uint32_t p00 = (uint32_t)((2 * q_i) * (G::WP / 2) + 1 + q_j);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int gg = g0 + u;
          const unsigned char* s2 = smem + G::OFF_S2 + gg * G::LBO_S2 + p00 * 16;
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
          *reinterpret_cast<uint4*>(Yr + (((size_t)pb * Ho + ho) * Wo + wo) * G::OCP + gg * 8) = G::RELU_OUT ? relu8h(yp) : yp;
          if (Ysub != nullptr && !(ho & 1) && !(wo & 1))
            *reinterpret_cast<uint4*>(Ysub + (((size_t)pb * Hs + (ho >> 1)) * Ws + (wo >> 1)) * G::OCP + gg * 8) = yp;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[G::B_P]);
    };

    long long g = 0;
    long long prev_b = 0;
    int prev_wo0 = 0, prev_a = 0;
    bool prev_carry = false;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const long long b = item / n_strips;
      const int strip = (int)(item - b * n_strips);
      const int wo0 = strip * G::CP, cb = 2 * wo0 - 2;
      for (int step = 0; step < n_steps; ++step, ++g) {
        const int a = step * G::S - 2;
        // ---- epilogue 1: ReLU, zero outside the image ("same" padding of the second convolution) -> S1 ----
#pragma unroll
        for (int t = 0; t < G::N1; ++t) {
          float v[16];
          drain(v);
          const int p1 = G::P1_0 + G::TPX * t + opix;
          const int hh = a + y1[t], ww = cb + c1[t];
          const bool inimg = hh >= 0 && hh < H && ww >= 0 && ww < W;
          if (out_row && p1 < (G::S + 2) * G::WP) {
            const uint4 z = make_uint4(0, 0, 0, 0);
            unsigned char* dst = smem + G::OFF_S1 + g0 * G::LBO_S1 + p1 * 16;
            *reinterpret_cast<uint4*>(dst) = inimg ? relu8h(pack8h(v)) : z;
            *reinterpret_cast<uint4*>(dst + G::LBO_S1) = inimg ? relu8h(pack8h(v + 8)) : z;
          }
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[G::B_S1 + t]);
        }
        // ---- previous step: pool + store while the tensor pipe runs this step's second convolution ----
        if (g > 0) {
          pool_store(g - 1, prev_b, prev_wo0, prev_a);
          worker_sync<G::NWORK>();   // pooling has finished reading S2
          if (prev_carry) {
            for (int i = tid; i < G::NG * G::WP; i += G::NWORK) {
              const int gq = i / G::WP, px = i - gq * G::WP;
              const int hp = px / (G::WP / 2), cc = px - hp * (G::WP / 2);
              unsigned char* p = smem + G::OFF_S2 + gq * G::LBO_S2 + (hp * G::S2HALF + cc) * 16;
              *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(p + G::S * (G::WP / 2) * 16);
            }
            worker_sync<G::NWORK>();
          }
        }
        // ---- epilogue 2: -inf outside the image (TF "same" max-pool padding) -> S2 ----
#pragma unroll
        for (int t = 0; t < G::N2; ++t) {
          float v[16];
          drain(v);
          const int p2 = G::P2_0 + G::TPX * t + opix;
          const int hh = a + y2[t], ww = cb + c2[t];
          const bool inimg = hh >= 0 && hh < H && ww >= 0 && ww < W;
          if (out_row && p2 < (G::S + 1) * G::WP) {
            const uint4 ninf = make_uint4(0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u);
            unsigned char* dst = smem + G::OFF_S2 + g0 * G::LBO_S2 + ((c2[t] & 1) * G::S2HALF + y2[t] * (G::WP / 2) + (c2[t] >> 1)) * 16;
            *reinterpret_cast<uint4*>(dst) = inimg ? pack8h(v) : ninf;
            *reinterpret_cast<uint4*>(dst + G::LBO_S2) = inimg ? pack8h(v + 8) : ninf;
          }
        }
        worker_sync<G::NWORK>();   // every worker has drained the second convolution: S1 is free, S2 is written
        // ---- carry the S1 overlap rows into the next step (rows above the next strip's first row are zero) ----
        const bool carry = step + 1 < n_steps;
        if (g + 1 < total_steps) {
          for (int i = tid; i < G::NG * 2 * G::WP; i += G::NWORK) {
            const int gq = i / (2 * G::WP), px = i - gq * 2 * G::WP;
            unsigned char* p = smem + G::OFF_S1 + gq * G::LBO_S1 + px * 16;
            *reinterpret_cast<uint4*>(p) = carry ? *reinterpret_cast<const uint4*>(p + G::S * G::WP * 16) : make_uint4(0, 0, 0, 0);
          }
          worker_sync<G::NWORK>();
        }
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
