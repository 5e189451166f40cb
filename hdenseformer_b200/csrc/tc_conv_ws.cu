// Weight-stationary tcgen05 convolution for the Cout = 32 layers (3x3x3, stride 1, pad 1; models/HDenseFormer.py:151,167):
// block_1_2_left/right, block_1_1_right, up3 and the input gradients that have 32 output channels -- 53 % of the
// model's FLOPs.
//
// Why a second formulation.  tc_conv_fwd_kernel (tc_conv.cu) puts the voxels on the MMA's M side (A operand = input box in
// shared memory) and Cout (x3 kw taps) on N = 96.  Measured on B200 (profiles/r2_umma_probe2.txt): an SS-mode MMA is bound
// by the 128 B/clk the tensor core reads from shared memory -- (4 KB of A + 32 N bytes of B) / 128 clk, i.e. 56 clk for
// N = 96 against 48 clk of math -- and that shared memory is also where TMA writes the next boxes (3 per tile).  Here the
// roles are swapped and the weights never leave the tensor core's own memory:
//     D[(kw, co), v] += sum_ci  Wt[kd,kh,kw][co][ci] * X[v + (kd,kh)][ci]
//   * A operand = weights, M = 128 rows = 3 kw taps x 32 output channels (+32 zero rows), bf16 pairs resident in TMEM
//     for the whole kernel (written once with tcgen05.st); TS-mode MMAs read no A bytes from shared memory, so an MMA
//     costs N/2 clk (profiles/r2_umma_probe2.txt);
//   * B operand = the input box, N = TH x TW voxels of one plane (up to 176 accumulator columns), K-major swizzled rows
//     exactly as TMA delivers channels-last voxels; the three kh taps read the same box at line-aligned row offsets;
//   * plane ring: a persistent CTA walks a column of tiles along D; the box of input plane z is loaded ONCE and used by the
//     three output planes z-1, z, z+1 (the kd taps), i.e. (TH+2)/TH x TW/(TW-2) = ~1.5 box rows per output voxel instead
//     of the 3 x 1.33 of the per-tile kd loads -- the TMA unit was the other wall of the old kernel;
//   * epilogue: rows (kw, co) of one output channel live in the same TMEM lane quarter (lane = 8 kw + c), are fetched with
//     tcgen05.ld.16x128b (a thread holds row r and r+8 = taps kw0 / kw1, a second load gives kw2) and recombined with two
//     shuffles per 32 outputs:  Y[co][j] = D0[co][j] + D1[co][j+1] + D2[co][j+2];  bf16 results are staged in shared
//     memory and leave as full 64-byte voxel rows.
// The same kernel serves the input gradient (tap-flipped, channel-swapped packed weights), like tc_conv_fwd_kernel.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

int hdf_sm_count_cached();
unsigned* hdf_ticket_slot(void* stream);      // glue.cu

namespace {
using namespace tcptx;

constexpr int WS_THREADS = 320;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue (2 per TMEM lane quarter)
constexpr int WS_COUT = 32;

struct WsParams {
  int N, D, H, W, Cin;
  int TH, TW, TWu;               // tile: TH lines x TW box columns (TWu = TW - 2 output columns)
  int nTh, nTw, nSeg, seg_len;   // D is cut into nSeg segments of seg_len planes (work item = one (n, seg, th, tw))
  int num_items;
  int NT;                        // MMA N = TH * TW accumulator columns
  int ksteps;                    // Cin / 16
  int stages;
  uint32_t box_bytes, stage_bytes, line_bytes, stg_bytes;
  uint32_t layout, sbo;
  uint32_t w_col0, acc_col0, acc_stride, tmem_cols;
  const bf16* wp;                // packed weights [27][32][Cin]
  const float* bias;
  bf16* y;
  long long ldy;
  double* stats;                 // optional InstanceNorm partial sums [N][gridDim.x][2][32] (sum, sum of squares of the
                                 // bf16-rounded outputs per sample and channel), else null
  unsigned* tickets;             // with stats: non-null -> the last CTA turns the partials into mean / rstd itself
  float* mean; float* rstd;      // [N][32]
  float eps; long long V;
  unsigned long long* dbg;
};

template <int GROUPS>   // GROUPS = TW / 4 column groups per line (4 or 8)
__global__ void __launch_bounds__(WS_THREADS, 1) tc_conv_ws_kernel(const __grid_constant__ CUtensorMap tmx, const WsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int S = p.stages;
  const uint32_t ring_base = smem_base;
  const uint32_t stg_base = ring_base + (uint32_t)S * p.stage_bytes;
  const uint32_t bar_base = stg_base + 2u * p.stg_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmx);
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // ---- resident weights: TMEM lane m = 32 q + 8 kw + c holds Wt[(kd,kh), kw][co = 8 q + c][:], lanes 32q+24.. are zero.
  // Slice (b = kd*3+kh, ks) occupies 8 columns (16 bf16 input channels) at w_col0 + (b*ksteps + ks)*8.
  if (warp >= 2) {
    const int q = warp & 3;
    const int kw = lane >> 3, c = lane & 7;
    const bool valid = lane < 24;
    for (int b = (warp - 2) >> 2; b < 9; b += 2) {
      const bf16* row = p.wp + ((size_t)((b * 3 + (valid ? kw : 0)) * WS_COUT + q * 8 + c)) * p.Cin;
      for (int ks = 0; ks < p.ksteps; ++ks) {
        uint4 v0 = make_uint4(0, 0, 0, 0), v1 = make_uint4(0, 0, 0, 0);
        if (valid) {
          v0 = *reinterpret_cast<const uint4*>(row + ks * 16);
          v1 = *reinterpret_cast<const uint4*>(row + ks * 16 + 8);
        }
        const uint32_t r[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        tmem_st_32x32b_x8(tmem_base + ((uint32_t)(q * 32) << 16) + p.w_col0 + (uint32_t)((b * p.ksteps + ks) * 8), r);
      }
    }
    tmem_st_wait();
  }
  // InstanceNorm statistics of the outputs, fused into the epilogue (p.stats != null): the epilogue threads hold one output
  // channel each, so sum / sum-of-squares accumulate in two registers per thread and meet in shared memory once per sample
  __shared__ float st_sh[8][WS_COUT][2];
  for (int i = threadIdx.x; i < 8 * WS_COUT * 2; i += WS_THREADS) (&st_sh[0][0][0])[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int per_n = p.nSeg * p.nTh * p.nTw;
  auto decode = [&](int item, int& n, int& d0, int& d1, int& h0, int& w0) {
    n = item / per_n;
    int r = item - n * per_n;
    const int seg = r / (p.nTh * p.nTw);
    r -= seg * (p.nTh * p.nTw);
    const int th = r / p.nTw, tw = r - th * p.nTw;
    d0 = seg * p.seg_len;
    d1 = min(p.D, d0 + p.seg_len);
    h0 = th * p.TH;
    w0 = tw * p.TWu;
  };

  if (warp == 0) {
    // ===== TMA producer: one box per input plane z = d0-1 .. d1 of every work item (planes outside the volume are
    // zero-filled by TMA = the convolution's padding along D; same for the h / w halo)
    {
      const uint32_t issue = elect_one_sync() ? 1u : 0u;
      uint32_t s = 0, ph = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        int n, d0, d1, h0, w0;
        decode(item, n, d0, d1, h0, w0);
        for (int z = d0 - 1; z <= d1; ++z) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx_p(full_bar(s), p.box_bytes, issue);
          tma_load_5d_p(ring_base + s * p.stage_bytes, &tmx, full_bar(s), 0, w0 - 1, h0 - 1, z, n, issue);
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp waits, one elected lane issues)
    const uint32_t idesc = umma_idesc(128, p.NT, 0, 0);
    const uint64_t desc_hi = umma_desc(0, 16, p.sbo, p.layout);
    const int ksteps = p.ksteps;
    const uint32_t stage_bytes = p.stage_bytes, line_bytes = p.line_bytes;
    const uint32_t w_tmem = tmem_base + p.w_col0;
    const uint32_t issue = elect_one_sync() ? 1u : 0u;
    // ring bookkeeping, all incremental (no divisions on the issue path): s0 = slot of plane d-1 of the current tile;
    // (ws, wph) = next slot / phase to be observed full; ahead = planes observed full from s0 on (need 3 per tile)
    uint32_t s0 = 0, ws = 0, wph = 0;
    int ahead = 0;
    int acc = 0; uint32_t accph = 0;
    bool tempty_ok = false;    // the current tile's accumulator was already observed free (during the previous tile)
    long long w_full = 0, w_tempty = 0; const long long mt0 = p.dbg ? clock64() : 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int n, d0, d1, h0, w0;
      decode(item, n, d0, d1, h0, w0);
      const int nd = d1 - d0;
      for (int t = 0; t < nd; ++t) {
        long long t0 = p.dbg ? clock64() : 0;
        while (ahead < 3) {
          mbar_wait(full_bar(ws), wph);
          if (++ws == (uint32_t)S) { ws = 0; wph ^= 1u; }
          ++ahead;
        }
        if (p.dbg) { const long long t1 = clock64(); w_full += t1 - t0; t0 = t1; }
        if (!tempty_ok) mbar_wait(tempty_bar(acc), accph ^ 1u);
        tempty_ok = false;
        if (p.dbg) w_tempty += clock64() - t0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + p.acc_col0 + (uint32_t)acc * p.acc_stride;
        const uint32_t s1 = s0 + 1 == (uint32_t)S ? 0u : s0 + 1;
        const uint32_t s2 = s1 + 1 == (uint32_t)S ? 0u : s1 + 1;
        uint32_t accflag = 0;
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
          // one-plane volumes (the 2-D model run as flat volumes): planes z-1 and z+1 are the zero padding, their taps add nothing
          if (p.D == 1 && kd != 1) continue;
          if (kd == 2 && t + 1 < nd) {
            // The barrier round trips of the NEXT tile (its newest plane, its accumulator) are taken here, while the
            // tensor core still works through the MMAs queued above: the issuing thread runs in lockstep with the
            // tensor pipe (MMA issue blocks when the queue is full), so waits at the tile boundary were pure bubbles.
            long long t2 = p.dbg ? clock64() : 0;
            mbar_wait(full_bar(ws), wph);
            if (++ws == (uint32_t)S) { ws = 0; wph ^= 1u; }
            ++ahead;
            if (p.dbg) { const long long t3 = clock64(); w_full += t3 - t2; t2 = t3; }
            mbar_wait(tempty_bar(acc ^ 1), (acc == 1 ? accph ^ 1u : accph) ^ 1u);
            tempty_ok = true;
            if (p.dbg) w_tempty += clock64() - t2;
            tc_fence_after();
          }
          const uint32_t box = ring_base + (kd == 0 ? s0 : kd == 1 ? s1 : s2) * stage_bytes;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint64_t bdesc = desc_hi | (uint64_t)(((box + (uint32_t)kh * line_bytes) >> 4) & 0x3FFF);
            const uint32_t a0 = w_tmem + (uint32_t)((kd * 3 + kh) * ksteps * 8);
            for (int ks = 0; ks < ksteps; ++ks) {     // +32 B per K=16 step inside the swizzled row (encoded >>4)
              umma_ts_p(d_tmem, a0 + (uint32_t)(ks * 8), bdesc + (uint64_t)(2 * ks), idesc, accflag, issue);
              accflag = 1;
            }
          }
        }
        umma_commit_p(empty_bar(s0), issue);             // plane d-1 is no longer needed
        if (t == nd - 1) {                               // end of the column: release the last two planes as well
          umma_commit_p(empty_bar(s1), issue);
          umma_commit_p(empty_bar(s2), issue);
          s0 = s2 + 1 == (uint32_t)S ? 0u : s2 + 1;
          ahead -= 3;
        } else {
          s0 = s1;
          ahead -= 1;
        }
        umma_commit_p(tfull_bar(acc), issue);
        if (++acc == 2) { acc = 0; accph ^= 1u; }
      }
    }
    if (p.dbg && lane == 0) {
      p.dbg[blockIdx.x * 8 + 0] = (unsigned long long)w_full; p.dbg[blockIdx.x * 8 + 1] = (unsigned long long)w_tempty;
      p.dbg[blockIdx.x * 8 + 2] = (unsigned long long)(clock64() - mt0);
    }
  } else {
    // ===== epilogue warps 2..9: TMEM lane quarter q = warp % 4 (output channels 8q .. 8q+7), two warps per quarter that
    // take alternate lines of the tile
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int c = lane >> 2, pcol = lane & 3;
    const int src1 = (lane & ~3) | ((pcol + 1) & 3), src2 = (lane & ~3) | ((pcol + 2) & 3);
    const int etid = threadIdx.x - 64;            // 0..255
    const float bias = p.bias ? p.bias[q * 8 + c] : 0.f;
    constexpr int TW = GROUPS * 4, TWu = TW - 2;
    const int TH = p.TH;
    // staging offset of this thread's (channel, column phase): [line][column][32 ch] bf16, the voxel's four 16-byte channel
    // groups XOR-swizzled by (column >> 1) so that the 2-byte stores of a warp spread over all banks
    const uint32_t stg_thr = (uint32_t)(pcol * 64 + c * 2);
    const int xs = pcol >> 1;
    int acc = 0; uint32_t accph = 0;
    const bool do_stats = p.stats != nullptr;
    float st1 = 0.f, st2 = 0.f;
    int n_cur = -1;
    // partial sums of this thread's channel over the 4 column phases -> shared memory (two adds per address per sample:
    // the two warps of the quarter; floating-point addition of two values commutes, so the result is deterministic)
    auto flush_stats = [&](int n) {
      if (!do_stats || n < 0) return;
      float a1 = st1, a2 = st2;
      a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
      a1 += __shfl_xor_sync(0xffffffffu, a1, 2); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
      if (pcol == 0) { atomicAdd(&st_sh[n][q * 8 + c][0], a1); atomicAdd(&st_sh[n][q * 8 + c][1], a2); }
      st1 = 0.f; st2 = 0.f;
    };
    long long e_wait = 0; const long long et0 = p.dbg ? clock64() : 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int n, d0, d1, h0, w0;
      decode(item, n, d0, d1, h0, w0);
      if (n != n_cur) { flush_stats(n_cur); n_cur = n; }
      for (int d = d0; d < d1; ++d) {
        const long long t0 = p.dbg ? clock64() : 0;
        mbar_wait(tfull_bar(acc), accph);
        if (p.dbg) e_wait += clock64() - t0;
        tc_fence_after();
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + p.acc_col0 + (uint32_t)acc * p.acc_stride;
        uint8_t* stg = smem_gen + (stg_base - smem_base) + (uint32_t)acc * p.stg_bytes;
        auto load_line = [&](int lh, uint32_t* a, uint32_t* b) {
          if (GROUPS == 8) {
            tmem_ld_16x128b_x8(tcol + (uint32_t)(lh * TW), a);
            tmem_ld_16x128b_x8(tcol + (16u << 16) + (uint32_t)(lh * TW), b);
          } else {
            tmem_ld_16x128b_x4(tcol + (uint32_t)(lh * TW), a);
            tmem_ld_16x128b_x4(tcol + (16u << 16) + (uint32_t)(lh * TW), b);
          }
        };
        auto combine_line = [&](int lh, const uint32_t* a, const uint32_t* b) {
          uint8_t* row = stg + (uint32_t)(lh * TW * 64) + stg_thr;
#pragma unroll
          for (int g = 0; g < GROUPS; ++g) {
            const int gn = g + 1 < GROUPS ? g + 1 : g;
            // a[2g] = D0 (kw 0), a[2g+1] = D1 (kw 1), b[2g] = D2 (kw 2) of channel 8q+c at box column j = 4g + pcol
            const float send1 = __uint_as_float(pcol == 0 ? a[2 * gn + 1] : a[2 * g + 1]);
            const float send2 = __uint_as_float(pcol < 2 ? b[2 * gn] : b[2 * g]);
            const float r1 = __shfl_sync(0xffffffffu, send1, src1);     // D1 at column j + 1
            const float r2 = __shfl_sync(0xffffffffu, send2, src2);     // D2 at column j + 2
            const float yv = __uint_as_float(a[2 * g]) + r1 + r2 + bias;
            const bf16 yb = __float2bfloat16_rn(yv);
            *reinterpret_cast<bf16*>(row + (uint32_t)(g * 256) + (uint32_t)((q ^ ((2 * g + xs) & 3)) * 16)) = yb;
            if (do_stats && 4 * g + pcol < TWu && w0 + 4 * g + pcol < p.W && h0 + lh < p.H) {
              const float yr = __bfloat162float(yb);      // statistics of the values the next kernels will read
              st1 += yr;
              st2 = fmaf(yr, yr, st2);
            }
          }
        };
        {
          uint32_t a0[2 * GROUPS], b0[2 * GROUPS], a1[2 * GROUPS], b1[2 * GROUPS];
          int lh = half;
          if (lh < TH) load_line(lh, a0, b0);
          while (lh < TH) {
            tmem_ld_wait();
            if (lh + 2 < TH) load_line(lh + 2, a1, b1);
            combine_line(lh, a0, b0);
            lh += 2;
            if (lh >= TH) break;
            tmem_ld_wait();
            if (lh + 2 < TH) load_line(lh + 2, a0, b0);
            combine_line(lh, a1, b1);
            lh += 2;
          }
        }
        // accumulator drained: hand it back to the MMA warp before the copy-out
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        named_bar_sync(1, 256);          // all four quarters of every line of the tile are staged
        const int nchunks = TH * TW * 4;
        for (int idx = etid; idx < nchunks; idx += 256) {
          const int voxel = idx >> 2, chunk = idx & 3;
          const int lh = voxel / TW, j = voxel - lh * TW;
          const int h = h0 + lh, w = w0 + j;
          if (j < TWu && h < p.H && w < p.W) {
            const uint4 v = *reinterpret_cast<const uint4*>(stg + (uint32_t)(voxel * 64) + (uint32_t)((chunk ^ ((j >> 1) & 3)) * 16));
            *reinterpret_cast<uint4*>(p.y + ((((long long)n * p.D + d) * p.H + h) * p.W + w) * p.ldy + chunk * 8) = v;
          }
        }
        // (the other staging buffer is used next; this one is rewritten two tiles later, after another named barrier)
        if (++acc == 2) { acc = 0; accph ^= 1u; }
      }
    }
    if (do_stats) {
      flush_stats(n_cur);
      named_bar_sync(1, 256);
      // partial[((n * chunks + cta) * 2 + which) * 32 + channel], the layout of the stand-alone statistics pass
      for (int i = etid; i < p.N * 2 * WS_COUT; i += 256) {
        const int n = i / (2 * WS_COUT), which = (i / WS_COUT) & 1, ch = i % WS_COUT;
        p.stats[(((size_t)n * gridDim.x + blockIdx.x) * 2 + which) * WS_COUT + ch] = (double)st_sh[n][ch][which];
      }
      if (p.tickets) {
        // the last CTA to get here finalises (same summation order as stats_finalize_kernel in glue.cu: lanes stride over
        // the CTAs' partials, fixed-order shuffle reduction)
        __shared__ int ws_last;
        __threadfence();
        named_bar_sync(1, 256);
        if (etid == 0) ws_last = atomicAdd(p.tickets, 1u) == gridDim.x - 1;
        named_bar_sync(1, 256);
        if (ws_last) {
          __threadfence();
          const int ew = etid >> 5, chunks = gridDim.x;
          for (int i = ew; i < p.N * WS_COUT; i += 8) {
            const int n = i / WS_COUT, ch = i % WS_COUT;
            double sm = 0.0, sq = 0.0;
            for (int k = lane; k < chunks; k += 32) {
              sm += __ldcg(p.stats + (((size_t)n * chunks + k) * 2 + 0) * WS_COUT + ch);
              sq += __ldcg(p.stats + (((size_t)n * chunks + k) * 2 + 1) * WS_COUT + ch);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { sm += __shfl_xor_sync(0xffffffffu, sm, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
            if (lane == 0) {
              const double m = sm / (double)p.V;
              double var = sq / (double)p.V - m * m;
              if (var < 0.0) var = 0.0;
              p.mean[i] = (float)m;
              p.rstd[i] = (float)(1.0 / sqrt(var + (double)p.eps));
            }
          }
          if (etid == 0) *p.tickets = 0;
        }
      }
    }
    if (p.dbg && warp == 2 && lane == 0) {
      p.dbg[blockIdx.x * 8 + 3] = (unsigned long long)e_wait; p.dbg[blockIdx.x * 8 + 4] = (unsigned long long)(clock64() - et0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn ws_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

}  // namespace

extern "C" {

// 1 if the weight-stationary kernel takes this stride-1 convolution
int hdf_tc_ws_supported(int mode, int Cin, int Cout) {
  const char* off = getenv("HDF_TC_NO_WS");     // read per call: tests and benches A/B the two kernels in one process
  return !(off && off[0] == '1') && mode == 0 && Cout == WS_COUT && (Cin == 32 || Cin == 64);
}

// Plan + launch.  Same contract as hdf_tc_conv3d_fwd(mode 0): x [N,D,H,W,Cin] bf16 (channel stride ldx), packed weights
// [27][32][Cin] bf16, y [N,D,H,W,32] bf16 (channel stride ldy), optional fp32 bias.
struct WsStatsOut { float* mean; float* rstd; float eps; bool fused; };

static int ws_launch(const void* x, long long ldx, const void* w_packed_bf16, const float* bias, void* y, long long ldy, int N,
                     int D, int H, int W, int Cin, double* stats, int* grid_out, void* stream, WsStatsOut* so = nullptr) {
  HDF_REQUIRE(hdf_tc_ws_supported(0, Cin, WS_COUT), "hdf_tc_ws_conv3d_fwd: unsupported Cin=%d", Cin);
  HDF_REQUIRE(x && w_packed_bf16 && y, "hdf_tc_ws_conv3d_fwd: null pointer");
  HDF_REQUIRE((ldx % 8 == 0) && (ldy % 8 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) &&
                  ((uintptr_t)w_packed_bf16 % 16 == 0),
              "hdf_tc_ws_conv3d_fwd: operands must be 16-byte aligned with channel strides multiple of 8");
  EncodeTiledFn enc = ws_get_encode();
  if (!enc) { hdf_set_error("hdf_tc_ws_conv3d_fwd: cuTensorMapEncodeTiled unavailable"); return HDF_ERR_CUDA; }
  WsParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin;
  p.ksteps = Cin / 16;
  const int rowbytes = Cin * 2;
  p.layout = rowbytes == 128 ? 2u : 4u;
  p.sbo = 8u * rowbytes;
  const int wcols = 9 * p.ksteps * 8;
  int nmax = ((512 - wcols) / 2) & ~15;
  static const int nmax_env = getenv("HDF_TC_WS_NMAX") ? atoi(getenv("HDF_TC_WS_NMAX")) : 0;
  if (nmax_env >= 16 && nmax_env < nmax) nmax = nmax_env & ~15;
  // tile: TW in {32, 16} box columns (2 of them halo), TH lines, N = TH*TW <= nmax accumulator columns (x2 buffers);
  // cost ~ accumulator columns processed per plane + a fixed per-tile overhead
  double best = -1;
  for (int tw = 32; tw >= 16; tw /= 2)
    for (int th = 1; th * tw <= nmax; ++th) {
      if ((th * tw) % 16) continue;
      const double cost = (double)cdiv(H, th) * cdiv(W, tw - 2) * (th * tw + 24.0);
      if (best < 0 || cost < best) { best = cost; p.TH = th; p.TW = tw; }
    }
  p.TWu = p.TW - 2;
  p.NT = p.TH * p.TW;
  p.nTh = cdiv(H, p.TH); p.nTw = cdiv(W, p.TWu);
  p.line_bytes = (uint32_t)p.TW * rowbytes;
  p.box_bytes = (uint32_t)(p.TH + 2) * p.line_bytes;
  p.stage_bytes = (p.box_bytes + 1023u) & ~1023u;
  p.stg_bytes = (uint32_t)p.NT * 64u;
  static const int smem_kb = getenv("HDF_TC_SMEM_KB") ? atoi(getenv("HDF_TC_SMEM_KB")) : 192;
  static const int max_stages = getenv("HDF_TC_WS_STAGES") ? atoi(getenv("HDF_TC_WS_STAGES")) : 8;
  p.stages = (int)(((size_t)smem_kb * 1024 - 2 * p.stg_bytes - 2048) / p.stage_bytes);
  if (p.stages > max_stages) p.stages = max_stages;
  HDF_REQUIRE(p.stages >= 4, "hdf_tc_ws_conv3d_fwd: plane ring needs >= 4 stages (box %u bytes)", p.box_bytes);
  p.w_col0 = 0;
  p.acc_col0 = (uint32_t)((wcols + 31) & ~31);
  p.acc_stride = (uint32_t)p.NT;
  p.tmem_cols = 512;
  HDF_REQUIRE(p.acc_col0 + 2 * p.acc_stride <= 512, "hdf_tc_ws_conv3d_fwd: TMEM plan does not fit");
  // segments along D: fewest (longest) segments whose round-robin schedule over the persistent CTAs is >= 97 % balanced
  const int sms = hdf_sm_count_cached();
  const long long cols = (long long)N * p.nTh * p.nTw;
  {
    double best_eff = -1; int best_len = D;
    for (int k = 1; k <= D; ++k) {
      const int len = cdiv(D, k), nseg = cdiv(D, len);
      const long long items = cols * nseg;
      const long long rounds = (items + sms - 1) / sms;
      const double eff = (double)cols * D / ((double)rounds * sms * len) * (len / (len + 0.25));   // mild penalty for halo planes
      if (eff > best_eff + 1e-9) { best_eff = eff; best_len = len; }
      if (eff >= 0.97) { best_len = len; break; }
    }
    static const int seg_env = getenv("HDF_TC_WS_SEG") ? atoi(getenv("HDF_TC_WS_SEG")) : 0;
    p.seg_len = seg_env > 0 ? (seg_env < D ? seg_env : D) : best_len;
    p.nSeg = cdiv(D, p.seg_len);
  }
  p.num_items = (int)(cols * p.nSeg);
  p.wp = (const bf16*)w_packed_bf16; p.bias = bias; p.y = (bf16*)y; p.ldy = ldy;
  p.stats = stats;
  if (stats && so) {
    p.tickets = hdf_ticket_slot(stream);
    p.mean = so->mean; p.rstd = so->rstd; p.eps = so->eps; p.V = (long long)D * H * W;
    so->fused = p.tickets != nullptr;
  }
  static const char* dbg_env = getenv("HDF_TC_DEBUG");
  static unsigned long long* dbg_buf = nullptr;
  if (dbg_env) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 148 * 8 * sizeof(unsigned long long));
    p.dbg = dbg_buf;
  }
  CUtensorMap tmx;
  {
    cuuint64_t gdim[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t gstr[4] = {(cuuint64_t)ldx * 2, (cuuint64_t)W * ldx * 2, (cuuint64_t)H * W * ldx * 2, (cuuint64_t)D * H * W * ldx * 2};
    cuuint32_t box[5] = {(cuuint32_t)Cin, (cuuint32_t)p.TW, (cuuint32_t)(p.TH + 2), 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, rowbytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hdf_set_error("hdf_tc_ws_conv3d_fwd: encode(x) failed: %d", (int)r); return HDF_ERR_CUDA; }
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 2 * (size_t)p.stg_bytes + 1024 + 8 * (2 * p.stages + 6) + 64;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_ws_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024));   // + 2 KB static
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_conv_ws_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024));
    if (e == cudaSuccess && getenv("HDF_NO_MAX_CARVEOUT") == nullptr) {
      e = cudaFuncSetAttribute(tc_conv_ws_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(tc_conv_ws_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    }
    if (e != cudaSuccess) { hdf_set_error("hdf_tc_ws_conv3d_fwd: smem attribute: %s", cudaGetErrorString(e)); return HDF_ERR_CUDA; }
    configured = true;
  }
  const int grid = p.num_items < sms ? p.num_items : sms;
  if (grid_out) *grid_out = grid;
  if (p.TW == 32) tc_conv_ws_kernel<8><<<grid, WS_THREADS, smem, (cudaStream_t)stream>>>(tmx, p);
  else tc_conv_ws_kernel<4><<<grid, WS_THREADS, smem, (cudaStream_t)stream>>>(tmx, p);
  HDF_LAUNCH_CHECK("hdf_tc_ws_conv3d_fwd");
  if (p.dbg) {
    unsigned long long h[8];
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tc_conv_ws dbg] Cin=%d TH=%d TW=%d NT=%d stages=%d seg=%d items=%d | mma: wait_full=%llu wait_tempty=%llu total=%llu | "
            "epilogue: wait_tfull=%llu total=%llu\n", Cin, p.TH, p.TW, p.NT, p.stages, p.seg_len, p.num_items, h[0], h[1], h[2], h[3], h[4]);
  }
  return HDF_OK;
}

int hdf_tc_ws_conv3d_fwd(const void* x, long long ldx, const void* w_packed_bf16, const float* bias, void* y, long long ldy, int N,
                         int D, int H, int W, int Cin, void* stream) {
  return ws_launch(x, ldx, w_packed_bf16, bias, y, ldy, N, D, H, W, Cin, nullptr, nullptr, stream);
}

// Same convolution with the InstanceNorm statistics of its output (models/HDenseFormer.py:152,168: per (sample, channel)
// mean and 1/sqrt(biased var + eps) over D*H*W) produced by the epilogue: no separate pass over the 2 x 191 MB output.
// workspace: hdf_tc_ws_stats_workspace(N) bytes.
size_t hdf_tc_ws_stats_workspace(int N) { return (size_t)N * 148 * 2 * WS_COUT * sizeof(double); }
int hdf_instnorm_stats_finalize(const double* partial, int chunks, int N, int C, long long V, float eps, float* mean, float* rstd,
                                void* stream);
int hdf_tc_ws_conv3d_fwd_stats(const void* x, long long ldx, const void* w_packed_bf16, const float* bias, void* y, long long ldy,
                               int N, int D, int H, int W, int Cin, float eps, float* mean, float* rstd, void* workspace,
                               size_t ws_bytes, void* stream) {
  HDF_REQUIRE(mean && rstd && workspace && N >= 1 && N <= 8, "hdf_tc_ws_conv3d_fwd_stats: bad args (N must be <= 8)");
  HDF_REQUIRE(ws_bytes >= hdf_tc_ws_stats_workspace(N) && hdf_sm_count_cached() <= 148, "hdf_tc_ws_conv3d_fwd_stats: workspace too small");
  int grid = 0;
  WsStatsOut so{mean, rstd, eps, false};
  const int rc = ws_launch(x, ldx, w_packed_bf16, bias, y, ldy, N, D, H, W, Cin, (double*)workspace, &grid, stream, &so);
  if (rc) return rc;
  if (so.fused) return HDF_OK;
  return hdf_instnorm_stats_finalize((const double*)workspace, grid, N, WS_COUT, (long long)D * H * W, eps, mean, rstd, stream);
}

}  // extern "C"
