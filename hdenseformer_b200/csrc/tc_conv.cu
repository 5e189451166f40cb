// tcgen05 / TMEM / TMA implicit-GEMM 3x3x3 convolution for sm_100a (bf16 activations, fp32 accumulate).
//
//   Y[v, co] = sum_{tap, ci} X[v + off(tap), ci] * Wt[tap][co][ci]          (nn.Conv3d k3 s1 p1,
//                                                                            models/HDenseFormer.py:151,167)
// GEMM view per CTA tile: M = 128 output voxels (a TD x TH x TW box of one sample), N = Cout (16..256),
// K = 27 taps x Cin.  For every (tap, 16/32/64-channel chunk) the producer warp issues
//   * one 5-D TMA box load of the shifted input box (out-of-bounds rows are zero-filled by TMA, which is the
//     conv's zero padding), landing as a K-major [128 x KC] bf16 tile with hardware swizzle, and
//   * one 2-D TMA load of the [Cout x KC] weight slice of that tap,
// into a multi-stage shared-memory ring guarded by full/empty mbarriers.  One elected thread issues
// tcgen05.mma (M128 x N x K16, cta_group::1) accumulating in TMEM; the accumulator is double-buffered so the
// four epilogue warps (tcgen05.ld -> +bias -> bf16 -> 128-bit global stores) overlap the next tile's MMAs.
// The kernel is persistent: grid = min(#tiles, #SMs), static round-robin tile schedule.
// The same kernel computes the input gradient (dgrad) with tap-flipped, channel-swapped packed weights.
#include <cuda.h>
#include <string.h>

#include "common.cuh"

int hdf_sm_count_cached();
// (weight-stationary kernel for the Cout = 32 layers: hdf_tc_ws_* in tc_conv_ws.cu, declared in hdf_b200.h)

namespace {

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded spin: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (it > (1u << 26)) __trap();
  }
}
// one lane of a converged warp; the compiler keeps the guarded block on the uniform datapath (no per-lane loop
// around instructions that take uniform-register operands such as tcgen05.mma / TMA)
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// predicated forms (see umma_bf16_p): every lane of the producer warp computes the (uniform) coordinates, only the lane
// elected at kernel start executes the instruction -- no per-lane R2UR loop around the uniform-datapath TMA instruction
__device__ __forceinline__ void tma_load_5d_p(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3,
                                              int c4, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %8, 0;\n\t"
      "@q cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n\t}"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_p(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_p(uint32_t bar, uint32_t bytes, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes), "r"(issue) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulation
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Predicated forms for the MMA-issuing warp.  Measured on B200 (profiles/r2_umma_probe2.txt vs profiles/r1_umma_probe.txt):
// an SS-mode M128 x N96 x K16 MMA executes in 56 clk when the issuing code keeps its operands in uniform registers, but
// the round-1 kernels spent ~147 clk per MMA because the descriptors were computed inside the divergent
// `if (elect_one_sync())` region: per-thread registers, i.e. an R2UR round trip for every operand of every MMA.  With
// these forms ALL 32 lanes run the same (uniform) address arithmetic, control flow never diverges, and only the tensor-core
// instruction itself is predicated on the lane elected once at kernel start.
__device__ __forceinline__ void umma_bf16_p(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc,
                                            uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void umma_commit_p(uint32_t bar, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (sm_100 UMMA): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout type [61,64): 2 = 128B swizzle, 4 = 64B, 6 = 32B
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}
// instruction descriptor: c=f32 (bit 4), a=bf16 (bits 7-9 = 1), b=bf16 (bits 10-12 = 1), majors (bits 15,16),
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ inline uint32_t umma_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ----------------------------------------------------------------------------- forward / dgrad kernel
// Tap table: the same kernel runs (mode 0) conv k3 s1 p1 -- 1 class, 27 taps, input coord = j + (k-1);
// (mode 1) transposed conv k3 s2 p1 op1 -- 8 output-parity classes with 1/2/4/8 taps over the INPUT grid,
// input coord = j + {0,1}, output coord = 2j + parity; (mode 2) conv k3 s2 p1 (input gradient of
// mode 1) -- 27 taps, input coord = 2j + (k-1) through a TMA map with element stride 2.
struct TcTaps {
  signed char dd[64], dh[64], dw[64];
  signed char widx[64];
  signed char first[9];
  signed char pd[8], ph[8], pw[8];
  int ncls;
};

struct TcConvParams {
  int N, D, H, W, Cin, Cout;   // D,H,W: base grid the tiles walk over
  int Do, Ho, Wo;              // output tensor dims
  int in_scale, out_scale;
  TcTaps taps;
  int TD, TH, TW;              // TMA box (TW is the box width, including the 2 halo columns in fold mode)
  int TWstep;                  // tile origin step along W (= TW, or TW-2 in fold mode)
  int nTd, nTh, nTw;
  int num_tiles;
  int KC, kchunks, stages;
  int fold;                    // 1: the three kw taps are folded into the MMA N dimension (N = 3*Cout) and
                               //    recombined across TMEM lanes (rows) in the epilogue with warp shuffles
  int Nmma;                    // MMA N = accumulator columns (Cout, or 3*Cout when folded)
  int khfold;                  // 1 (fold mode, resident weights, TD == 1): the box carries TH+2 lines and the three kh
                               //   taps are three MMA groups reading the same box at line-aligned row offsets, so a
                               //   tile needs 3 pipeline stages (kd) instead of 9
  uint32_t line_bytes;         // TW * KC * 2
  int b_resident;              // 1: all weight tiles stay resident in shared memory for the whole kernel
  int a_cpasync;               // 1: the input box is gathered by 4 producer warps with 16-byte cp.async (zero-filled
                               //    out of bounds) instead of TMA: the TMA unit retires ~1 box row per 5.5 clk, which
                               //    caps 64/128-byte rows well below L2 bandwidth (profiles/r1_ncu_conv_fwd_tma.txt)
  int commit_group;            // pipeline stages released per batch of tcgen05.commit (1..8)
  unsigned long long* dbg;     // optional per-CTA wait-cycle counters [grid][8] (HDF_TC_DEBUG), else null
  const bf16* x;               // input tensor (cp.async path)
  long long ldx;
  int Di, Hi, Wi;              // input tensor dims
  uint32_t b_region;           // round1024(b_bytes)
  uint32_t a_bytes, b_bytes, stage_bytes;  // stage_bytes = round1024(a) (+ round1024(b) when B is streamed)
  uint32_t layout, sbo;                    // UMMA layout type / stride-byte-offset of the swizzle mode
  uint32_t tmem_cols;
  long long ldy;
  const float* bias;
  bf16* y;
};

// Dynamic shared memory budget of one persistent convolution CTA.  192 KB (not the full 227) leaves ~34 KB of the SM for
// the token-branch kernels of the side streams (dct_c_fwd needs 28.7 KB), which otherwise wait for the CTA to drain.
#ifndef HDF_TC_SMEM_KB_DEFAULT
#define HDF_TC_SMEM_KB_DEFAULT 192
#endif
constexpr int TC_THREADS = 256;     // wgrad kernel: warp 0 TMA, 1 MMA, 2 TMEM alloc, 4-7 epilogue
constexpr int FWD_THREADS = 288;    // fwd kernel: warps 0-3 producers, 4 MMA + TMEM alloc, 5-8 epilogue
constexpr int CP_LAG = 3;           // cp.async groups kept in flight per producer thread

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(FWD_THREADS, 1)
tc_conv_fwd_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmw, const TcConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t a_region = (p.a_bytes + 1023u) & ~1023u;
  const int ntaps_total = p.taps.first[p.taps.ncls];
  // smem: [resident weight tiles] [stage ring] [barriers]
  const uint32_t bres_base = smem_base;
  const uint32_t ring_base =
      smem_base + (p.b_resident ? (uint32_t)((p.khfold ? 9 : ntaps_total) * p.kchunks) * p.b_region : 0u);
  const uint32_t bar_base = ring_base + p.stages * p.stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + 2 + a); };
  const uint32_t bres_bar = bar_base + 8u * (2 * p.stages + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 5);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmw);
  }
  if (warp == 1 && lane == 0) {
    // full barrier arrivals per phase: TMA path = 1 (expect_tx); cp.async path = 128 producer threads
    // (+1 expect_tx arrival when the weight tile is streamed by TMA into the same stage)
    const uint32_t full_count = p.a_cpasync ? (128u + (p.b_resident ? 0u : 1u)) : 1u;
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), full_count); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int tiles_per_n = p.nTd * p.nTh * p.nTw;
  const int tiles_per_cls = p.N * tiles_per_n;

  if (warp < 4) {
    // ===== producers =====
    const int ptid = threadIdx.x;   // 0..127
    if (ptid == 0 && p.b_resident) {
      const int nres = p.khfold ? 9 : ntaps_total;      // resident weight tiles: one per (kd,kh) in kh-fold mode
      mbar_expect_tx(bres_bar, (uint32_t)(nres * p.kchunks) * p.b_bytes);
      for (int e = 0; e < nres; ++e)
        for (int kc = 0; kc < p.kchunks; ++kc)
          tma_load_2d(bres_base + (uint32_t)(e * p.kchunks + kc) * p.b_region, &tmw, bres_bar, kc * p.KC,
                      (p.khfold ? e * 3 : (int)p.taps.widx[e]) * p.Cout);
    }
    if (!p.a_cpasync) {
      // ---- TMA: warp 0 walks the tile schedule with all lanes (uniform coordinates), one elected lane issues
      if (warp == 0) {
        const uint32_t issue = elect_one_sync() ? 1u : 0u;
        long long dbg_acc = 0;
        int s = 0; uint32_t ph = 0;               // stage / phase of the next iteration
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
          const int cls = tile / tiles_per_cls;
          int r = tile - cls * tiles_per_cls;
          const int n = r / tiles_per_n;
          r -= n * tiles_per_n;
          const int tw = r % p.nTw; r /= p.nTw;
          const int th = r % p.nTh;
          const int td = r / p.nTh;
          const int d0 = td * p.TD * p.in_scale, h0 = th * p.TH * p.in_scale, w0 = tw * p.TWstep * p.in_scale;
          const int e0 = p.taps.first[cls];
          const int ntap = p.taps.first[cls + 1] - e0;
          for (int t = 0; t < ntap; ++t) {
            const int e = e0 + t;
            const int cw = w0 + p.taps.dw[e], ch = h0 + p.taps.dh[e], cd = d0 + p.taps.dd[e];
            const int wrow = (int)p.taps.widx[e] * p.Cout;
            for (int kc = 0; kc < p.kchunks; ++kc) {
              const long long t0 = p.dbg ? clock64() : 0;
              mbar_wait(empty_bar(s), ph ^ 1u);
              if (p.dbg) dbg_acc += clock64() - t0;
              mbar_expect_tx_p(full_bar(s), p.a_bytes + (p.b_resident ? 0u : p.b_bytes), issue);
              const uint32_t a_dst = ring_base + s * p.stage_bytes;
              tma_load_5d_p(a_dst, &tmx, full_bar(s), kc * p.KC, cw, ch, cd, n, issue);
              if (!p.b_resident) tma_load_2d_p(a_dst + a_region, &tmw, full_bar(s), kc * p.KC, wrow, issue);
              if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
          }
        }
        if (p.dbg && lane == 0) { p.dbg[blockIdx.x * 8 + 0] = dbg_acc; }
      }
    } else {
      // ---- cp.async gather: thread -> (16-byte chunk j of the row, rows rsub + i*rpp), written with the same
      // XOR swizzle the UMMA descriptor expects (address bits [4,7) ^= bits [7,10) within the swizzle span)
      const int cpr = p.KC / 8;          // 16-byte chunks per row: 2, 4 or 8
      const int rpp = 128 / cpr;         // rows covered by one pass of the 128 producer threads
      const int j = ptid % cpr, rsub = ptid / cpr;
      const uint32_t rowbytes = (uint32_t)p.KC * 2u;
      int rcol[8], rlh[8], rld[8];
      uint32_t roff[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rsub + i * rpp;
        rcol[i] = (r % p.TW) * p.in_scale;
        const int line = r / p.TW;
        rlh[i] = (line % p.TH) * p.in_scale;
        rld[i] = (line / p.TH) * p.in_scale;
        const uint32_t phys = (uint32_t)j ^ (((uint32_t)r * rowbytes >> 7) & (uint32_t)(cpr - 1));
        roff[i] = (uint32_t)r * rowbytes + phys * 16u;
      }
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int cls = tile / tiles_per_cls;
        int r = tile - cls * tiles_per_cls;
        const int n = r / tiles_per_n;
        r -= n * tiles_per_n;
        const int tw = r % p.nTw; r /= p.nTw;
        const int th = r % p.nTh;
        const int td = r / p.nTh;
        const int d0 = td * p.TD * p.in_scale, h0 = th * p.TH * p.in_scale, w0 = tw * p.TWstep * p.in_scale;
        const int e0 = p.taps.first[cls];
        const int kiters = (p.taps.first[cls + 1] - e0) * p.kchunks;
        const bf16* xn = p.x + (long long)n * p.Di * p.Hi * p.Wi * p.ldx + j * 8;
        for (int it = 0; it < kiters; ++it) {
          const int e = e0 + it / p.kchunks, kc = it % p.kchunks;
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t a_dst = ring_base + s * p.stage_bytes;
          if (ptid == 0 && !p.b_resident) {
            mbar_expect_tx(full_bar(s), p.b_bytes);
            tma_load_2d(a_dst + a_region, &tmw, full_bar(s), kc * p.KC, (int)p.taps.widx[e] * p.Cout);
          }
          const int wo = w0 + p.taps.dw[e], ho = h0 + p.taps.dh[e], dO = d0 + p.taps.dd[e];
          const bf16* xk = xn + kc * p.KC;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (i < cpr) {
              const int w = wo + rcol[i], h = ho + rlh[i], d = dO + rld[i];
              const bool ok = (unsigned)w < (unsigned)p.Wi && (unsigned)h < (unsigned)p.Hi && (unsigned)d < (unsigned)p.Di;
              const bf16* src = ok ? xk + (((long long)d * p.Hi + h) * p.Wi + w) * p.ldx : p.x;
              cp_async16(a_dst + roff[i], src, ok ? 16u : 0u);
            }
          }
          // the barrier receives this thread's arrival when all of its copies above have landed: no thread ever
          // blocks on its own loads, so up to `stages` boxes (the whole ring) are in flight per SM
          cp_async_mbar_arrive(full_bar(s));
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
      cp_async_wait<0>();
    }
  } else if (warp == 4) {
    {
      // ===== MMA issuer (whole warp waits on the barriers, one elected lane issues) =====
      // Everything loop-invariant lives in registers: the asm "memory" clobbers would otherwise force the kernel
      // parameters to be re-read from the constant bank around every instruction, and this single thread's
      // issue latency paces the whole pipeline (measured: ~600 clk per stage before this was hoisted).
      const uint32_t idesc = umma_idesc(128, p.Nmma, 0, 0);
      const int stages = p.stages, kchunks = p.kchunks, ksteps = p.KC / 16, ncls = p.taps.ncls;
      const uint32_t stage_bytes = p.stage_bytes, b_region = p.b_region, Nmma = (uint32_t)p.Nmma;
      const bool bres = p.b_resident != 0, cpa = p.a_cpasync != 0, dbg = p.dbg != nullptr, khfold = p.khfold != 0;
      const uint32_t line_bytes = p.line_bytes;
      const uint64_t desc_hi = umma_desc(0, 16, p.sbo, p.layout);
      const int num_tiles = p.num_tiles, gstride = gridDim.x;
      const int kiters_c0 = (p.taps.first[1] - p.taps.first[0]) * kchunks;
      int s = 0; uint32_t ph = 0;
      uint32_t a_addr = ring_base, fullb = full_bar(0), emptyb = empty_bar(0);
      int acc = 0; uint32_t accph = 0;
      const uint32_t issue = elect_one_sync() ? 1u : 0u;     // the one lane whose tcgen05 instructions are not predicated off
      long long w_full = 0, w_tempty = 0, w_mma = 0, w_commit = 0; const long long mt0 = dbg ? clock64() : 0;
      if (bres) { mbar_wait(bres_bar, 0); tc_fence_after(); }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gstride) {
        int kiters = kiters_c0, e0 = 0;
        if (ncls > 1) {
          const int cls = tile / tiles_per_cls;
          e0 = p.taps.first[cls];
          kiters = (p.taps.first[cls + 1] - e0) * kchunks;
        }
        long long t0 = dbg ? clock64() : 0;
        mbar_wait(tempty_bar(acc), accph ^ 1u);
        if (dbg) w_tempty += clock64() - t0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * Nmma;
        uint32_t b_addr = bres_base + (uint32_t)(e0 * kchunks) * b_region;
        uint32_t kh_b_addr = bres_base;     // kh-fold: resident weight tile (kd, kh = 0, kc) of the current stage
        int kh_kc = 0;
        uint32_t accflag = 0;
        for (int it = 0; it < kiters; ++it) {
          if (dbg) t0 = clock64();
          mbar_wait(fullb, ph);
          if (dbg) w_full += clock64() - t0;
          if (cpa) fence_proxy_async();
          tc_fence_after();
          long long t1 = dbg ? clock64() : 0;
          if (!khfold) {
            const uint64_t adesc = desc_hi | (uint64_t)((a_addr >> 4) & 0x3FFF);
            const uint64_t bdesc = desc_hi | (uint64_t)(((bres ? b_addr : a_addr + a_region) >> 4) & 0x3FFF);
#pragma unroll 4
            for (int k = 0; k < ksteps; ++k) {  // +32 B per K=16 step inside the swizzled row (encoded >>4)
              umma_bf16_p(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accflag, issue);
              accflag = 1;
            }
          } else {
            // stage = (kd, kc); kh = 0,1,2 reads box lines kh .. kh+TH-1 (row offset kh * TW rows, a multiple
            // of the 8-row swizzle atom) against the resident weight tile (kd, kh, kc)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              const uint64_t adesc = desc_hi | (uint64_t)(((a_addr + (uint32_t)kh * line_bytes) >> 4) & 0x3FFF);
              const uint32_t bt = kh_b_addr + (uint32_t)(kh * kchunks) * b_region;
              const uint64_t bdesc = desc_hi | (uint64_t)((bt >> 4) & 0x3FFF);
#pragma unroll 4
              for (int k = 0; k < ksteps; ++k) {
                umma_bf16_p(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accflag, issue);
                accflag = 1;
              }
            }
            // next stage: kc + 1, or the next kd's first chunk (weight tiles are laid out [(kd*3+kh)*kchunks + kc])
            if (++kh_kc == kchunks) { kh_kc = 0; kh_b_addr += (uint32_t)(2 * kchunks + 1) * b_region; }
            else kh_b_addr += b_region;
          }
          umma_commit_p(emptyb, issue);
          if (it == kiters - 1) umma_commit_p(tfull_bar(acc), issue);
          __syncwarp();
          accflag = 1;
          long long t2 = dbg ? clock64() : 0;
          if (dbg) { w_mma += t2 - t1; }
          a_addr += stage_bytes; fullb += 8; emptyb += 8; b_addr += b_region;
          if (++s == stages) { s = 0; ph ^= 1u; a_addr = ring_base; fullb = full_bar(0); emptyb = empty_bar(0); }
        }
        if (++acc == 2) { acc = 0; accph ^= 1u; }
      }
      if (dbg && lane == 0) { p.dbg[blockIdx.x * 8 + 2] = w_full; p.dbg[blockIdx.x * 8 + 3] = w_tempty; p.dbg[blockIdx.x * 8 + 4] = clock64() - mt0;
                 p.dbg[blockIdx.x * 8 + 7] = w_mma; p.dbg[blockIdx.x * 8 + 1] = w_commit; }
    }
  } else {
    // ===== epilogue (warps 5..8): TMEM -> registers -> (+bias) -> bf16 -> global =====
    const int q = warp % 4;  // TMEM lane quadrant this warp may access
    const int m = q * 32 + lane;
    const int mw = m % p.TW, mh = (m / p.TW) % p.TH, md = m / (p.TW * p.TH);
    int acc = 0; uint32_t accph = 0;
    long long ew_tfull = 0; const long long ept0 = p.dbg ? clock64() : 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int cls = tile / tiles_per_cls;
      int r = tile - cls * tiles_per_cls;
      const int n = r / tiles_per_n;
      r -= n * tiles_per_n;
      const int tw = r % p.nTw; r /= p.nTw;
      const int th = r % p.nTh;
      const int td = r / p.nTh;
      const int d = td * p.TD + md, h = th * p.TH + mh, w = tw * p.TWstep + mw;
      const bool valid = d < p.D && h < p.H && w < p.W && mw < p.TWstep;
      const int od = d * p.out_scale + p.taps.pd[cls], oh = h * p.out_scale + p.taps.ph[cls],
                ow = w * p.out_scale + p.taps.pw[cls];
      bf16* yrow = p.y + ((((long long)n * p.Do + od) * p.Ho + oh) * p.Wo + ow) * p.ldy;
      const long long et0 = p.dbg ? clock64() : 0;
      mbar_wait(tfull_bar(acc), accph);
      if (p.dbg) ew_tfull += clock64() - et0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Nmma);
      if (!p.fold) {
        for (int c0 = 0; c0 < p.Cout; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)c0, v);
          tmem_ld_wait();
          if (valid) {
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + (p.bias ? p.bias[c0 + j] : 0.f);
            store8<bf16>(yrow + c0, f);
            store8<bf16>(yrow + c0 + 8, f + 8);
          }
        }
      } else {
        // row m holds, for kw = 0,1,2, the partial sums P_kw[m] = sum_{kd,kh,ci} X[box row m] * W[kd,kh,kw];
        // output column mw needs P_0[m] + P_1[m+1] + P_2[m+2]  (box row m+kw = input w0-1+mw+kw): the rows are
        // neighbouring TMEM lanes of the same warp (a box line never straddles a warp), fetched with shuffles.
        for (int c0 = 0; c0 < p.Cout; c0 += 16) {
          uint32_t v0[16], v1[16], v2[16];
          tmem_ld16(taddr + (uint32_t)c0, v0);
          tmem_ld16(taddr + (uint32_t)(p.Cout + c0), v1);
          tmem_ld16(taddr + (uint32_t)(2 * p.Cout + c0), v2);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a1 = __shfl_down_sync(0xffffffffu, __uint_as_float(v1[j]), 1);
            const float a2 = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[j]), 2);
            f[j] = __uint_as_float(v0[j]) + a1 + a2 + (p.bias ? p.bias[c0 + j] : 0.f);
          }
          if (valid) {
            store8<bf16>(yrow + c0, f);
            store8<bf16>(yrow + c0 + 8, f + 8);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; accph ^= 1u; }
    }
    if (p.dbg && warp == 5 && lane == 0) { p.dbg[blockIdx.x * 8 + 5] = ew_tfull; p.dbg[blockIdx.x * 8 + 6] = clock64() - ept0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}


// ----------------------------------------------------------------------------- weight-gradient kernel
//   dW[tap][ci][co] = sum_v X[v + off(tap)][ci] * dY[v][co]
// GEMM per CTA: K = voxels of this CTA's slab (KV = 128 or 64 per pipeline chunk), N = Cout (the dY tile, MN-major
// B operand), M = 128 rows made of SPG = 128/CW shifted/offset input sub-tiles [KV x CW channels] stacked along M
// through the descriptor's leading-byte-offset (CW = min(Cin, 64)): for Cin = 32 one MMA covers 4 taps, for
// Cin = 64 two taps, for Cin >= 128 one tap's 128-channel slice.  Each such "group" owns Cout TMEM columns; a CTA
// handles one pass = up to 512/Cout groups, over one slab of voxel chunks (split-K); fp32 partials go to the
// workspace and a second kernel reduces them deterministically into the torch-layout gradient.
struct TcWgradParams {
  int N, D, H, W, Cin, Cout;
  int TD, TH, TW, nTd, nTh, nTw;
  int num_chunks, chunks_per_slab, num_slabs;
  int KV;                  // voxels per chunk (rows of every smem tile)
  int CW, SPG, sub_per_tap, total_sub, total_groups, groups_per_pass;
  int CWn, nsub_b;
  int a_stages;
  int a_scale;             // 1: A-side coords j + (k-1); 2: A-side is the 2x-resolution tensor, coords 2j + (k-1)
  int ntaps;               // 27, or 1 for the stem's im2col'ed operand (a single unshifted "tap"), or 9 with tap0 = 9 for one-plane volumes
  int tap0;                // first tap of the processed range (0; 9 = the kd = 1 plane when D == 1: the other planes only see padding)
  uint32_t a_sub_bytes, a_stage_bytes, b_sub_bytes, b_stage_bytes;
  uint32_t a_layout, a_sbo, b_layout, b_sbo;
  uint32_t tmem_cols;
  float* partial;          // [num_slabs][ntaps][Cin][Cout]
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmdy, const TcWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t b_base = smem_base + p.a_stages * p.a_stage_bytes;
  const uint32_t bar_base = b_base + 2 * p.b_stage_bytes;
  auto afull = [&](int s) { return bar_base + 8u * s; };
  auto aempty = [&](int s) { return bar_base + 8u * (p.a_stages + s); };
  auto bfull = [&](int s) { return bar_base + 8u * (2 * p.a_stages + s); };
  auto bempty = [&](int s) { return bar_base + 8u * (2 * p.a_stages + 2 + s); };
  const uint32_t accfull = bar_base + 8u * (2 * p.a_stages + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.a_stages + 5);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int slab = blockIdx.x, pass = blockIdx.y;
  const int g_begin = pass * p.groups_per_pass;
  const int g_end = min(p.total_groups, g_begin + p.groups_per_pass);
  const int c_begin = slab * p.chunks_per_slab;
  const int c_end = min(p.num_chunks, c_begin + p.chunks_per_slab);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmx); tma_prefetch_desc(&tmdy); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    mbar_init(accfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int tiles_per_n = p.nTd * p.nTh * p.nTw;

  if (warp == 0) {
    {
      // ===== TMA producer: all lanes walk the (uniform) schedule, one elected lane issues =====
      const uint32_t issue = elect_one_sync() ? 1u : 0u;
      const bool one_tap = p.ntaps == 1;
      int s = 0; uint32_t ph = 0;
      int bs = 0; uint32_t bph = 0;
      for (int c = c_begin; c < c_end; ++c) {
        const int n = c / tiles_per_n;
        int r = c - n * tiles_per_n;
        const int tw = r % p.nTw; r /= p.nTw;
        const int th = r % p.nTh;
        const int td = r / p.nTh;
        const int d0 = td * p.TD, h0 = th * p.TH, w0 = tw * p.TW;
        const int ad0 = d0 * p.a_scale - 1, ah0 = h0 * p.a_scale - 1, aw0 = w0 * p.a_scale - 1;
        mbar_wait(bempty(bs), bph ^ 1u);
        mbar_expect_tx_p(bfull(bs), p.b_sub_bytes * p.nsub_b, issue);
        for (int j = 0; j < p.nsub_b; ++j)
          tma_load_5d_p(b_base + bs * p.b_stage_bytes + j * p.b_sub_bytes, &tmdy, bfull(bs), j * p.CWn, w0, h0, d0, n, issue);
        if (++bs == 2) { bs = 0; bph ^= 1u; }
        // sub-tile u = g*SPG + j  ->  (tap, channel chunk), tracked incrementally (no per-load divisions)
        int tap = (g_begin * p.SPG) / p.sub_per_tap, sub = (g_begin * p.SPG) - tap * p.sub_per_tap;
        int kd = (tap + p.tap0) / 9, kh = (tap / 3) % 3, kw = tap % 3;
        for (int g = g_begin; g < g_end; ++g) {
          const int u0 = g * p.SPG;
          const int nsub = min(p.SPG, p.total_sub - u0);
          mbar_wait(aempty(s), ph ^ 1u);
          mbar_expect_tx_p(afull(s), p.a_sub_bytes * nsub, issue);
          for (int j = 0; j < nsub; ++j) {
            tma_load_5d_p(smem_base + s * p.a_stage_bytes + j * p.a_sub_bytes, &tmx, afull(s), sub * p.CW,
                          aw0 + (one_tap ? 1 : kw), ah0 + (one_tap ? 1 : kh), ad0 + (one_tap ? 1 : kd), n, issue);
            if (++sub == p.sub_per_tap) {
              sub = 0;
              if (++kw == 3) { kw = 0; if (++kh == 3) { kh = 0; ++kd; } }
            }
          }
          if (++s == p.a_stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ===== MMA issuer (whole warp waits, one elected lane issues) =====
      const uint32_t idesc = umma_idesc(128, p.Cout, 1, 1);   // both operands MN-major (K = voxel rows)
      int s = 0; uint32_t ph = 0;
      int bs = 0; uint32_t bph = 0;
      const int ksteps = p.KV / 16, a_stages = p.a_stages, Cout = p.Cout;
      const uint32_t a_stage_bytes = p.a_stage_bytes, b_stage_bytes = p.b_stage_bytes;
      const uint64_t adesc_hi = umma_desc(0, p.a_sub_bytes, p.a_sbo, p.a_layout);
      const uint64_t bdesc_hi = umma_desc(0, p.b_sub_bytes, p.b_sbo, p.b_layout);
      const uint64_t a_adv = (uint64_t)((2u * p.a_sbo) >> 4), b_adv = (uint64_t)((2u * p.b_sbo) >> 4);
      const uint32_t issue = elect_one_sync() ? 1u : 0u;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(bfull(bs), bph);
        tc_fence_after();
        const uint64_t bdesc = bdesc_hi | (uint64_t)(((b_base + bs * b_stage_bytes) >> 4) & 0x3FFF);
        const uint32_t accflag0 = (c != c_begin) ? 1u : 0u;
        uint32_t d_tmem = tmem_base;
        for (int g = g_begin; g < g_end; ++g) {
          mbar_wait(afull(s), ph);
          tc_fence_after();
          const uint64_t adesc = adesc_hi | (uint64_t)(((smem_base + s * a_stage_bytes) >> 4) & 0x3FFF);
          uint64_t ad = adesc, bd = bdesc;
          umma_bf16_p(d_tmem, ad, bd, idesc, accflag0, issue);
#pragma unroll 4
          for (int k = 1; k < ksteps; ++k) {   // advance 16 voxel rows = 2 swizzle-atom groups = 2*SBO bytes
            ad += a_adv; bd += b_adv;
            umma_bf16_p(d_tmem, ad, bd, idesc, 1u, issue);
          }
          umma_commit_p(aempty(s), issue);
          d_tmem += (uint32_t)Cout;
          if (++s == a_stages) { s = 0; ph ^= 1u; }
        }
        umma_commit_p(bempty(bs), issue);
        if (++bs == 2) { bs = 0; bph ^= 1u; }
      }
      umma_commit_p(accfull, issue);
    }
  } else if (warp >= 4) {
    // ===== epilogue: accumulators -> fp32 partials =====
    const int q = warp - 4;
    const int m = q * 32 + lane;
    mbar_wait(accfull, 0);
    tc_fence_after();
    for (int g = g_begin; g < g_end; ++g) {
      const int u = g * p.SPG + m / p.CW;
      const bool valid = u < p.total_sub;
      const int tap = valid ? u / p.sub_per_tap : 0;
      const int ci = valid ? (u - tap * p.sub_per_tap) * p.CW + m % p.CW : 0;
      float* dst = p.partial + (((long long)slab * p.ntaps + tap) * p.Cin + ci) * p.Cout;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((g - g_begin) * p.Cout);
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ----------------------------------------------------------------------------- weight-gradient kernel, v2
// Measured on B200 (profiles/r1_umma_probe.txt): an SS-mode tcgen05.mma with M=128 costs >= ~86-130 clk whatever N is
// (the 128-row A operand read from shared memory), and N/2 clk of math: small N wastes the tensor core.  v2 therefore
// puts the *shifted* operand S (27 taps x Cs channels) on the N side, up to 256 (tap, channel) columns per MMA, stacked
// from consecutive [KV x CWs] TMA sub-tiles through the B descriptor's leading-byte-offset; the fixed operand F
// (Cf channels, one tile per voxel chunk) is the 128-row M side (replicated sub-tiles when Cf < 128).
//   D[f, (tap, s)] = sum_v F[v][f] * S[v + off(tap)][s]        both operands MN-major, K = voxels
// A CTA owns one slab of voxel chunks (split-K) and one pass = the accumulators that fit 512 TMEM columns.
struct TcWgrad2Params {
  int N, D, H, W;
  int Cs, Cf;
  int TD, TH, TW, nTd, nTh, nTw;
  int num_chunks, chunks_per_slab, num_slabs;
  int KV;
  int CWs, spt, total_sub, SPGn, G;      // S sub-tile width, sub-tiles per tap, total sub-tiles, sub-tiles per group, groups
  int CWf, nsub_f, nrep_f, nMh;          // F sub-tile width, distinct sub-tiles, TMA replicas (Cf < 128), M halves
  int groups_per_pass, mh_per_pass, passes_g, passes_m;
  int s_stages;
  int s_scale;                           // 1, or 2 when S is the 2x-resolution tensor (transposed conv)
  uint32_t s_sub_bytes, s_stage_bytes, f_sub_bytes, f_stage_bytes;
  uint32_t s_layout, s_sbo, f_layout, f_sbo;
  uint32_t tmem_cols;
  float* partial;                        // [num_slabs][27][Cs][Cf]
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_wgrad2_kernel(const __grid_constant__ CUtensorMap tms, const __grid_constant__ CUtensorMap tmf, const TcWgrad2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t f_base = smem_base + p.s_stages * p.s_stage_bytes;
  const uint32_t bar_base = f_base + 2 * p.f_stage_bytes;
  auto sfull = [&](int s) { return bar_base + 8u * s; };
  auto sempty = [&](int s) { return bar_base + 8u * (p.s_stages + s); };
  auto ffull = [&](int s) { return bar_base + 8u * (2 * p.s_stages + s); };
  auto fempty = [&](int s) { return bar_base + 8u * (2 * p.s_stages + 2 + s); };
  const uint32_t accfull = bar_base + 8u * (2 * p.s_stages + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.s_stages + 5);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int slab = blockIdx.x;
  const int pass_g = blockIdx.y / p.passes_m, pass_m = blockIdx.y % p.passes_m;
  const int g_begin = pass_g * p.groups_per_pass, g_end = min(p.G, g_begin + p.groups_per_pass);
  const int h_begin = pass_m * p.mh_per_pass, h_end = min(p.nMh, h_begin + p.mh_per_pass);
  const int c_begin = slab * p.chunks_per_slab, c_end = min(p.num_chunks, c_begin + p.chunks_per_slab);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tms); tma_prefetch_desc(&tmf); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.s_stages; ++s) { mbar_init(sfull(s), 1); mbar_init(sempty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(ffull(s), 1); mbar_init(fempty(s), 1); }
    mbar_init(accfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int tiles_per_n = p.nTd * p.nTh * p.nTw;

  if (warp == 0) {
    // ===== TMA producer (one elected lane) =====
    int s = 0; uint32_t ph = 0;
    int fs = 0; uint32_t fph = 0;
    for (int c = c_begin; c < c_end; ++c) {
      const int n = c / tiles_per_n;
      int r = c - n * tiles_per_n;
      const int tw = r % p.nTw; r /= p.nTw;
      const int th = r % p.nTh;
      const int td = r / p.nTh;
      const int d0 = td * p.TD, h0 = th * p.TH, w0 = tw * p.TW;
      const int sd0 = d0 * p.s_scale - 1, sh0 = h0 * p.s_scale - 1, sw0 = w0 * p.s_scale - 1;
      mbar_wait(fempty(fs), fph ^ 1u);
      if (elect_one_sync()) {
        // the M side always spans 128 channels: [h_begin, h_end) halves of Cf, or nrep_f replicas when Cf < 128
        const int nld = (p.Cf < 128) ? p.nrep_f * p.nsub_f : (h_end - h_begin) * (128 / p.CWf);
        mbar_expect_tx(ffull(fs), p.f_sub_bytes * nld);
        for (int j = 0; j < nld; ++j) {
          const int sub = (p.Cf < 128) ? (j % p.nsub_f) : (h_begin * (128 / p.CWf) + j);
          tma_load_5d(f_base + fs * p.f_stage_bytes + j * p.f_sub_bytes, &tmf, ffull(fs), sub * p.CWf, w0, h0, d0, n);
        }
      }
      __syncwarp();
      if (++fs == 2) { fs = 0; fph ^= 1u; }
      for (int g = g_begin; g < g_end; ++g) {
        const int u0 = g * p.SPGn;
        const int nsub = min(p.SPGn, p.total_sub - u0);
        mbar_wait(sempty(s), ph ^ 1u);
        if (elect_one_sync()) {
          mbar_expect_tx(sfull(s), p.s_sub_bytes * nsub);
          for (int j = 0; j < nsub; ++j) {
            const int u = u0 + j;
            const int tap = u / p.spt, ch0 = (u - tap * p.spt) * p.CWs;
            const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
            tma_load_5d(smem_base + s * p.s_stage_bytes + j * p.s_sub_bytes, &tms, sfull(s), ch0, sw0 + kw, sh0 + kh, sd0 + kd, n);
          }
        }
        __syncwarp();
        if (++s == p.s_stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const int ksteps = p.KV / 16, SPGn = p.SPGn, CWs = p.CWs, total_sub = p.total_sub, s_stages = p.s_stages;
    const uint32_t s_stage_bytes = p.s_stage_bytes, f_stage_bytes = p.f_stage_bytes;
    const uint64_t sdesc_hi = umma_desc(0, p.s_sub_bytes, p.s_sbo, p.s_layout);
    const uint64_t fdesc_hi = umma_desc(0, p.f_sub_bytes, p.f_sbo, p.f_layout);
    const uint64_t s_adv = (uint64_t)((2u * p.s_sbo) >> 4), f_adv = (uint64_t)((2u * p.f_sbo) >> 4);
    const uint32_t f_half_bytes = (p.Cf < 128) ? 0u : p.f_sub_bytes * (uint32_t)(128 / p.CWf);
    const int nh = h_end - h_begin;
    int s = 0; uint32_t ph = 0;
    int fs = 0; uint32_t fph = 0;
    for (int c = c_begin; c < c_end; ++c) {
      mbar_wait(ffull(fs), fph);
      tc_fence_after();
      const uint32_t f_addr = f_base + fs * f_stage_bytes;
      uint32_t col = 0;
      for (int g = g_begin; g < g_end; ++g) {
        const int nsub = min(SPGn, total_sub - g * SPGn);
        const int Ng = nsub * CWs;
        mbar_wait(sfull(s), ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t idesc = umma_idesc(128, Ng, 1, 1);
          const uint64_t bdesc = sdesc_hi | (uint64_t)(((smem_base + s * s_stage_bytes) >> 4) & 0x3FFF);
          for (int h = 0; h < nh; ++h) {
            const uint64_t adesc = fdesc_hi | (uint64_t)(((f_addr + (uint32_t)h * f_half_bytes) >> 4) & 0x3FFF);
            const uint32_t d_tmem = tmem_base + col + (uint32_t)(h * Ng);
#pragma unroll 4
            for (int k = 0; k < ksteps; ++k)
              umma_bf16(d_tmem, adesc + f_adv * k, bdesc + s_adv * k, idesc, (c != c_begin) || (k != 0));
          }
          umma_commit(sempty(s));
        }
        __syncwarp();
        col += (uint32_t)(nh * Ng);
        if (++s == s_stages) { s = 0; ph ^= 1u; }
      }
      if (elect_one_sync()) umma_commit(fempty(fs));
      __syncwarp();
      if (++fs == 2) { fs = 0; fph ^= 1u; }
    }
    if (elect_one_sync()) umma_commit(accfull);
    __syncwarp();
  } else if (warp >= 4) {
    // ===== epilogue: accumulators -> fp32 partials [tap][s][f] (lane = f: warp-coalesced stores) =====
    const int q = warp - 4;
    const int m = q * 32 + lane;
    mbar_wait(accfull, 0);
    tc_fence_after();
    uint32_t col = 0;
    for (int g = g_begin; g < g_end; ++g) {
      const int nsub = min(p.SPGn, p.total_sub - g * p.SPGn);
      const int Ng = nsub * p.CWs;
      for (int h = h_begin; h < h_end; ++h) {
        const int f = (p.Cf < 128) ? m : h * 128 + m;
        const bool fvalid = (p.Cf < 128) ? (m < p.Cf) : true;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + col + (uint32_t)((h - h_begin) * Ng);
        for (int c0 = 0; c0 < Ng; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)c0, v);
          tmem_ld_wait();
          if (fvalid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int cc = c0 + j;
              const int u = g * p.SPGn + cc / p.CWs;
              const int tap = u / p.spt;
              const int sch = (u - tap * p.spt) * p.CWs + cc % p.CWs;
              p.partial[(((long long)slab * 27 + tap) * p.Cs + sch) * p.Cf + f] = __uint_as_float(v[j]);
            }
          }
        }
      }
      col += (uint32_t)((h_end - h_begin) * Ng);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}


// partials [S][27][Cin][Cout] -> g[ci*sci + co*sco + tap] (+= if accumulate); fixed summation order
__global__ void tc_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ g, int S, int Cin, int Cout,
                                       long long sci, long long sco, int accumulate, int flip, int ntaps = 27, int tap0 = 0) {
  const long long per = (long long)ntaps * Cin * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < S; ++z) s += part[z * per + i];
    const int co = i % Cout;
    const int ci = (i / Cout) % Cin;
    const int tap = tap0 + (int)(i / ((long long)Cin * Cout));
    float* q = g + ci * sci + co * sco + (flip ? 26 - tap : tap);
    *q = accumulate ? (*q + s) : s;
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
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

CUtensorMapSwizzle swizzle_for(int inner_bytes) {
  return inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                             : CU_TENSOR_MAP_SWIZZLE_32B;
}

// pick (TD,TH,TW), powers of two with product 128, minimising the number of (possibly partial) tiles
void pick_tile(int D, int H, int W, int& TD, int& TH, int& TW) {
  long long best = -1;
  for (int tw = 1; tw <= 128; tw *= 2)
    for (int th = 1; th * tw <= 128; th *= 2) {
      const int td = 128 / (tw * th);
      const long long tiles = (long long)cdiv(D, td) * cdiv(H, th) * cdiv(W, tw);
      // prefer wide-W boxes on ties (longer contiguous global rows)
      if (best < 0 || tiles < best || (tiles == best && tw > TW)) { best = tiles; TD = td; TH = th; TW = tw; }
    }
}

// fold mode (mode 0 only): 9 (kd,kh) entries; the box starts one column left (dw = -1) and the weight tile of
// an entry is the 3*Cout rows of taps (kd,kh,0..2), which are consecutive in the packed [27][Cout][Cin] layout
void build_taps_fold(TcTaps& t) {
  memset(&t, 0, sizeof(t));
  t.ncls = 1;
  t.first[0] = 0; t.first[1] = 9;
  for (int e = 0; e < 9; ++e) {
    t.dd[e] = (signed char)(e / 3 - 1); t.dh[e] = (signed char)(e % 3 - 1); t.dw[e] = -1;
    t.widx[e] = (signed char)(e * 3);
  }
}

// one-plane volumes (D == 1: the 2-D model run as flat volumes): the taps of the planes above / below read nothing but
// zero padding -> keep only the entries with dd == 0 (a third of the boxes and MMAs)
void filter_taps_flat(TcTaps& t) {
  TcTaps o;
  memset(&o, 0, sizeof(o));
  o.ncls = t.ncls;
  int n = 0;
  for (int c = 0; c < t.ncls; ++c) {
    o.first[c] = (signed char)n;
    o.pd[c] = t.pd[c]; o.ph[c] = t.ph[c]; o.pw[c] = t.pw[c];
    for (int e = t.first[c]; e < t.first[c + 1]; ++e)
      if (t.dd[e] == 0) { o.dd[n] = 0; o.dh[n] = t.dh[e]; o.dw[n] = t.dw[e]; o.widx[n] = t.widx[e]; ++n; }
  }
  o.first[t.ncls] = (signed char)n;
  t = o;
}

void build_taps(int mode, TcTaps& t) {
  memset(&t, 0, sizeof(t));
  if (mode == 0 || mode == 2) {
    t.ncls = 1;
    t.first[0] = 0; t.first[1] = 27;
    for (int k = 0; k < 27; ++k) {
      t.dd[k] = (signed char)(k / 9 - 1); t.dh[k] = (signed char)((k / 3) % 3 - 1); t.dw[k] = (signed char)(k % 3 - 1);
      t.widx[k] = (signed char)k;
    }
  } else {
    // per dim: parity 0 -> tap k=1 reads input j ; parity 1 -> k=0 reads j+1, k=2 reads j   (o = 2 i - 1 + k)
    t.ncls = 8;
    int e = 0;
    for (int c = 0; c < 8; ++c) {
      const int pd = (c >> 2) & 1, ph = (c >> 1) & 1, pw = c & 1;
      t.first[c] = (signed char)e;
      t.pd[c] = (signed char)pd; t.ph[c] = (signed char)ph; t.pw[c] = (signed char)pw;
      const int kd_n = pd ? 2 : 1, kh_n = ph ? 2 : 1, kw_n = pw ? 2 : 1;
      for (int a = 0; a < kd_n; ++a)
        for (int b = 0; b < kh_n; ++b)
          for (int cc = 0; cc < kw_n; ++cc) {
            const int kd = pd ? (a == 0 ? 0 : 2) : 1, kh = ph ? (b == 0 ? 0 : 2) : 1, kw = pw ? (cc == 0 ? 0 : 2) : 1;
            t.dd[e] = (signed char)(kd == 0 ? 1 : 0); t.dh[e] = (signed char)(kh == 0 ? 1 : 0);
            t.dw[e] = (signed char)(kw == 0 ? 1 : 0);
            t.widx[e] = (signed char)(kd * 9 + kh * 3 + kw);
            ++e;
          }
    }
    t.first[8] = (signed char)e;   // 27
  }
}

__global__ void tc_pack_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cin, int Cout, long long sci,
                               long long sco, int flip, int cin_valid) {
  const long long total = 27ll * Cin * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = i % Cin;
    const int co = (i / Cin) % Cout;
    const int tap = i / ((long long)Cin * Cout);
    out[i] = ci < cin_valid ? __float2bfloat16_rn(w[ci * sci + co * sco + (flip ? 26 - tap : tap)]) : __float2bfloat16_rn(0.f);
  }
}

// (A tiled variant with coalesced 32-byte runs through shared memory was measured SLOWER, 36 vs 26 us per launch: the
// weights are L2-resident, so the strided 4-byte reads cost no DRAM traffic and the element-wise form has more loads in
// flight -- profiles/r2_timeline_v5.txt.)
// ----------------------------------------------------------------------------- stem (first conv, 1..4 input channels)
// The first layer has only M = 1..4 input channels (block_1_1_left, SURVEY 8d: AI 51, bandwidth-bound).  Zero-padding it
// to 16 channels costs 9 MMAs with 32-byte rows per 128 voxels (the slowest UMMA operand shape, profiles/r1_umma_probe.txt)
// and 27 shifted box loads in the weight gradient.  Instead the 27*M taps are gathered once per step into a K-major
// matrix Xcol [N*V, Kp] (Kp = 27*M rounded up to 64, bf16): forward = one 128B-swizzled box + Kp/16 MMAs per tile,
// weight gradient = one unshifted operand pair per voxel chunk (Xcol is kept for the backward pass).
// block = one (n, d, h) output line.  Phase 1 stages the 9 (kd,kh) input lines of every channel in shared memory
// (coalesced float loads, zero halo / out-of-range lines); phase 2 assembles the W output rows from shared memory with
// consecutive threads writing consecutive 16-byte chunks of a row (fully coalesced 2*Kp-byte rows).
template <int CIN>
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, bf16* __restrict__ out, int D, int H, int W,
                                                         int Kp) {
  extern __shared__ float lines[];            // [9*CIN][W+2]
  int line = blockIdx.x;
  const int h = line % H; line /= H;
  const int d = line % D;
  const long long n = line / D;
  const long long V = (long long)D * H * W;
  const float* xn = x + n * CIN * V;
  const int Wp = W + 2;
  // one warp per staged line (no per-element index division), up to 8 independent loads in flight per lane
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int l = warp; l < 9 * CIN; l += nwarps) {
    const int t9 = l / CIN, ci = l - t9 * CIN;
    const int dd = d + t9 / 3 - 1, hh = h + t9 % 3 - 1;
    const bool okl = (unsigned)dd < (unsigned)D && (unsigned)hh < (unsigned)H;
    const float* src = xn + ci * V + ((long long)dd * H + hh) * W - 1;     // src[p] = input column p-1
    float* dst = lines + l * Wp;
    for (int p0 = lane; p0 < Wp; p0 += 256) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int pcol = p0 + u * 32;
        v[u] = (okl && pcol >= 1 && pcol <= W) ? src[pcol] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int pcol = p0 + u * 32;
        if (pcol < Wp) dst[pcol] = v[u];
      }
    }
  }
  __syncthreads();
  // 256 % (Kp/8) == 0: a thread keeps the same 16-byte chunk of the row for all its items, so the tap decode of its
  // 8 k's is done once (the kernel was issue-bound when it decoded per element)
  const int cpr = Kp / 8;
  const int k0 = (threadIdx.x % cpr) * 8;
  int off[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = k0 + j;
    const int tap = k / CIN, ci = k - tap * CIN;
    const int t9 = tap / 3, kw = tap - t9 * 3;
    off[j] = tap < 27 ? (t9 * CIN + ci) * Wp + kw : -1;
  }
  bf16* orow = out + ((long long)blockIdx.x * W) * Kp;
  for (int i = threadIdx.x; i < W * cpr; i += blockDim.x) {
    const int w = i / cpr;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = off[j] >= 0 ? lines[off[j] + w] : 0.f;
    store8<bf16>(orow + (long long)i * 8, v);
  }
}

// packed[co][k = tap*Cin + ci] (bf16, zero for k >= 27*Cin) from the torch weight [Cout][Cin][27]
__global__ void stem_pack_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cin, int Cout, int Kp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * Kp) return;
  const int co = i / Kp, k = i % Kp;
  const int tap = k / Cin, ci = k % Cin;
  out[i] = __float2bfloat16_rn(tap < 27 ? w[((long long)co * Cin + ci) * 27 + tap] : 0.f);
}

// partial [S][Kp][Cout] -> dw [Cout][Cin][27]
__global__ void stem_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ g, int S, int Cin, int Cout, int Kp,
                                         int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 27 * Cin * Cout) return;
  const int co = i % Cout, k = i / Cout;
  const int tap = k / Cin, ci = k % Cin;
  float s = 0.f;
  for (int z = 0; z < S; ++z) s += part[((long long)z * Kp + k) * Cout + co];
  float* q = g + ((long long)co * Cin + ci) * 27 + tap;
  *q = accumulate ? (*q + s) : s;
}

}  // namespace

extern "C" {

int hdf_tc_supported(int mode, int Cin, int Cout) {
  if (mode < 0 || mode > 2) return 0;
  if (Cin % 16 != 0 || Cin < 16) return 0;
  if (Cout % 16 != 0 || Cout < 16 || Cout > 256) return 0;
  return 1;
}

size_t hdf_tc_pack_bytes(int Cin, int Cout) { return (size_t)27 * Cin * Cout * sizeof(bf16); }

// packed[tap][co][ci] (bf16) = ci < cin_valid ? w[ci*stride_ci + co*stride_co + (flip ? 26-tap : tap)] : 0
// (cin_valid < Cin zero-pads the K dimension so that 2..4-channel inputs can use the 16-channel TMA/UMMA path)
int hdf_tc_pack_weights(const float* w, void* packed_bf16, int Cin, int Cout, long long stride_ci, long long stride_co,
                        int flip, int cin_valid, void* stream) {
  HDF_REQUIRE(w && packed_bf16, "hdf_tc_pack_weights: null pointer");
  const long long total = 27ll * Cin * Cout;
  tc_pack_kernel<<<min(1024, cdiv(total, 256)), 256, 0, (cudaStream_t)stream>>>(w, (bf16*)packed_bf16, Cin, Cout, stride_ci,
                                                                              stride_co, flip, cin_valid);
  HDF_LAUNCH_CHECK("hdf_tc_pack_weights");
  return HDF_OK;
}

// mode 3 (internal, stem path): 1x1x1 "conv" over an im2col'ed tensor [N, D, H, W, Cin = Kp]; weights [Cout][Kp]
static int tc_conv_fwd_launch(int mode, const void* x, long long ldx, const void* w_packed_bf16, const float* bias, void* y,
                              long long ldy, int N, int Do, int Ho, int Wo, int Cin, int Cout, void* stream);

int hdf_tc_conv3d_fwd(int mode, const void* x, long long ldx, const void* w_packed_bf16, const float* bias, void* y,
                      long long ldy, int N, int Do, int Ho, int Wo, int Cin, int Cout, void* stream) {
  if (!hdf_tc_supported(mode, Cin, Cout)) {
    hdf_set_error("hdf_tc_conv3d_fwd: unsupported channels Cin=%d Cout=%d", Cin, Cout);
    return HDF_ERR_UNSUPPORTED;
  }
  if (hdf_tc_ws_supported(mode, Cin, Cout))
    return hdf_tc_ws_conv3d_fwd(x, ldx, w_packed_bf16, bias, y, ldy, N, Do, Ho, Wo, Cin, stream);
  return tc_conv_fwd_launch(mode, x, ldx, w_packed_bf16, bias, y, ldy, N, Do, Ho, Wo, Cin, Cout, stream);
}

static int tc_conv_fwd_launch(int mode, const void* x, long long ldx, const void* w_packed_bf16, const float* bias, void* y,
                              long long ldy, int N, int Do, int Ho, int Wo, int Cin, int Cout, void* stream) {
  HDF_REQUIRE(x && w_packed_bf16 && y, "hdf_tc_conv3d_fwd: null pointer");
  HDF_REQUIRE((ldx % 8 == 0) && (ldy % 8 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0),
              "hdf_tc_conv3d_fwd: operands must be 16-byte aligned with channel strides multiple of 8");
  EncodeTiledFn enc = get_encode();
  if (!enc) { hdf_set_error("hdf_tc_conv3d_fwd: cuTensorMapEncodeTiled unavailable"); return HDF_ERR_CUDA; }

  TcConvParams p;
  if (mode == 1) HDF_REQUIRE(((Do | Ho | Wo) & 1) == 0, "hdf_tc_conv3d_fwd: transposed conv needs even output dims");
  // base grid the tiles walk over: output grid (modes 0, 2) or input grid (mode 1)
  const int D = mode == 1 ? Do / 2 : Do, H = mode == 1 ? Ho / 2 : Ho, W = mode == 1 ? Wo / 2 : Wo;
  // input tensor dims
  const int Di = mode == 2 ? 2 * Do : D, Hi = mode == 2 ? 2 * Ho : H, Wi = mode == 2 ? 2 * Wo : W;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.Do = Do; p.Ho = Ho; p.Wo = Wo;
  p.in_scale = mode == 2 ? 2 : 1;
  p.out_scale = mode == 1 ? 2 : 1;
  p.KC = (Cin % 64 == 0) ? 64 : (Cin % 32 == 0) ? 32 : 16;
  p.kchunks = Cin / p.KC;
  p.a_bytes = 128u * p.KC * 2u;
  const int inner = p.KC * 2;
  p.layout = inner == 128 ? 2u : inner == 64 ? 4u : 6u;
  p.sbo = 8u * inner;

  // ---- plain plan: one TMA box + one MMA group per tap
  if (mode == 3) {
    memset(&p.taps, 0, sizeof(p.taps));
    p.taps.ncls = 1; p.taps.first[0] = 0; p.taps.first[1] = 1;      // one tap, no shift
  } else {
    build_taps(mode, p.taps);
  }
  pick_tile(D, H, W, p.TD, p.TH, p.TW);
  p.TWstep = p.TW;
  p.nTd = cdiv(D, p.TD); p.nTh = cdiv(H, p.TH); p.nTw = cdiv(W, p.TW);
  p.fold = 0;
  p.Nmma = Cout;
  double best_cost = (double)p.nTd * p.nTh * p.nTw * 27.0 * (128.0 + Cout);   // ~ L2->SMEM rows per sample
  // ---- folded plan (stride-1 conv, 3*Cout <= 256): box width 32 or 16 incl. 2 halo columns, 9 boxes per tile
  static const char* no_fold = getenv("HDF_TC_NO_FOLD");
  if (mode == 0 && 3 * Cout <= 256 && !no_fold) {
    for (int wb = 32; wb >= 16; wb /= 2) {
      const int lines = 128 / wb, step = wb - 2;
      int bth = 1, btd = lines; long long bt = -1;
      for (int th = 1; th <= lines; th *= 2) {
        const int td = lines / th;
        const long long t = (long long)cdiv(D, td) * cdiv(H, th);
        if (bt < 0 || t < bt) { bt = t; bth = th; btd = td; }
      }
      const double cost = (double)bt * cdiv(W, step) * 9.0 * 128.0;
      if (cost < best_cost) {
        best_cost = cost;
        p.fold = 1; p.TW = wb; p.TWstep = step; p.TH = bth; p.TD = btd;
        p.nTd = cdiv(D, btd); p.nTh = cdiv(H, bth); p.nTw = cdiv(W, step);
        p.Nmma = 3 * Cout;
      }
    }
    if (p.fold) build_taps_fold(p.taps);
  }
  const bool flat = mode == 0 && D == 1;
  if (flat) filter_taps_flat(p.taps);
  p.num_tiles = p.taps.ncls * N * p.nTd * p.nTh * p.nTw;
  p.b_bytes = (uint32_t)p.Nmma * p.KC * 2u;
  p.b_region = (p.b_bytes + 1023u) & ~1023u;
  const int ntaps_total = p.taps.first[p.taps.ncls];
  const size_t bres_bytes = (size_t)ntaps_total * p.kchunks * p.b_region;
  static const char* no_res = getenv("HDF_TC_NO_RESIDENT");
  // two resident CTAs per SM (each <= 110 KB of shared memory and <= 256 TMEM columns) when the folded weights are small
  // enough to stay resident beside >= 3 ring stages: two MMA-issue streams per SM hide each other's barrier round trips
  static const int fwd_ctas_cfg = getenv("HDF_TC_FWD_CTAS") ? atoi(getenv("HDF_TC_FWD_CTAS")) : 1;
  const size_t kh_stage = ((size_t)p.TW * (128 / p.TW + 2) * p.KC * 2u + 1023u) & ~(size_t)1023u;
  const bool two_ctas = fwd_ctas_cfg >= 2 && p.fold && 2 * p.Nmma <= 256 && bres_bytes + 3 * kh_stage <= 108 * 1024 && !no_res;
  static const int smem_kb_env = getenv("HDF_TC_SMEM_KB") ? atoi(getenv("HDF_TC_SMEM_KB")) : HDF_TC_SMEM_KB_DEFAULT;
  const int smem_kb_cfg = two_ctas ? 108 : smem_kb_env;
  // with a reduced budget (experiments) resident weights must leave room for >= 4 ring stages of the largest input box
  p.b_resident = two_ctas ? 1 : (bres_bytes <= 114 * 1024 && (smem_kb_cfg >= HDF_TC_SMEM_KB_DEFAULT || bres_bytes + 4 * 24 * 1024 <= (size_t)smem_kb_cfg * 1024) && !no_res) ? 1 : 0;
  p.khfold = 0;
  p.line_bytes = (uint32_t)p.TW * p.KC * 2u;
  static const char* no_kh = getenv("HDF_TC_NO_KHFOLD");
  if (p.fold && p.b_resident && !no_kh && !flat) {
    // kh-fold: one box of TH+2 lines per kd; needs all lines of the tile in one d-plane
    const int lines = 128 / p.TW;
    p.khfold = 1;
    p.TH = lines; p.TD = 1;
    p.nTd = D; p.nTh = cdiv(H, lines);
    p.num_tiles = N * p.nTd * p.nTh * p.nTw;
    memset(&p.taps, 0, sizeof(p.taps));
    p.taps.ncls = 1; p.taps.first[0] = 0; p.taps.first[1] = 3;
    for (int e = 0; e < 3; ++e) { p.taps.dd[e] = (signed char)(e - 1); p.taps.dh[e] = -1; p.taps.dw[e] = -1; p.taps.widx[e] = (signed char)(e * 9); }
    p.a_bytes = (uint32_t)p.TW * (lines + 2) * p.KC * 2u;
  }
  const uint32_t a_region_h = (p.a_bytes + 1023u) & ~1023u;
  p.stage_bytes = a_region_h + (p.b_resident ? 0u : p.b_region);
  // total dynamic smem target per CTA (KB): smaller values leave room for the transformer branch's small kernels to
  // co-reside with the persistent conv CTAs on the same SM
  const int smem_kb = smem_kb_cfg;
  const size_t ring_budget = (size_t)smem_kb * 1024 - (p.b_resident ? bres_bytes : 0);
  p.stages = (int)(ring_budget / p.stage_bytes);
  if (p.stages > 12) p.stages = 12;
  if (p.stages < 2) { hdf_set_error("hdf_tc_conv3d_fwd: stage too large"); return HDF_ERR_UNSUPPORTED; }
  uint32_t cols = 32;
  while (cols < 2u * p.Nmma) cols *= 2;
  p.tmem_cols = cols;
  p.ldy = ldy; p.bias = bias; p.y = (bf16*)y;
  // measured on B200 (profiles/r1_microbench_conv_v2_cpasync.txt): the cp.async gather is ~2x slower than the
  // TMA box loads for these shapes, so it stays opt-in for experiments
  static const char* use_cp = getenv("HDF_TC_CPASYNC");
  p.a_cpasync = use_cp ? 1 : 0;
  p.x = (const bf16*)x; p.ldx = ldx; p.Di = Di; p.Hi = Hi; p.Wi = Wi;
  p.commit_group = 1;
  p.dbg = nullptr;
  static const char* dbg_env = getenv("HDF_TC_DEBUG");
  static unsigned long long* dbg_buf = nullptr;
  if (dbg_env) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 148 * 8 * sizeof(unsigned long long));
    p.dbg = dbg_buf;
  }

  CUtensorMap tmx, tmw;
  {
    const cuuint32_t es = (cuuint32_t)p.in_scale;
    cuuint64_t gdim[5] = {(cuuint64_t)Cin, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)Di, (cuuint64_t)N};
    cuuint64_t gstr[4] = {(cuuint64_t)ldx * 2, (cuuint64_t)Wi * ldx * 2, (cuuint64_t)Hi * Wi * ldx * 2,
                          (cuuint64_t)Di * Hi * Wi * ldx * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.KC, (cuuint32_t)p.TW * es, (cuuint32_t)(p.TH + (p.khfold ? 2 : 0)) * es,
                         (cuuint32_t)p.TD * es, 1};
    cuuint32_t estr[5] = {1, es, es, es, 1};
    CUresult r = enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(inner), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hdf_set_error("hdf_tc_conv3d_fwd: encode(x) failed: %d", (int)r); return HDF_ERR_CUDA; }
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)Cin, (cuuint64_t)(mode == 3 ? 1 : 27) * Cout};
    cuuint64_t gstr[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.KC, (cuuint32_t)p.Nmma};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed_bf16), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(inner), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hdf_set_error("hdf_tc_conv3d_fwd: encode(w) failed: %d", (int)r); return HDF_ERR_CUDA; }
  }
  // (kh-fold reads up to 2 lines past the 128 rows of the last MMA window: they are inside the stage's box)
  const size_t smem = (p.b_resident ? bres_bytes : 0) + (size_t)p.stages * p.stage_bytes + 1024 /*align slack*/ +
                      8 * (2 * p.stages + 6) + 64;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    // full 228 KB carve-out even when the CTA asks for less: the driver otherwise picks the smallest configuration that
    // fits THIS kernel (e.g. 196 KB for a 194 KB CTA), and the side streams' kernels find no shared memory left on the SM
    if (e == cudaSuccess && getenv("HDF_NO_MAX_CARVEOUT") == nullptr)
      e = cudaFuncSetAttribute(tc_conv_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) { hdf_set_error("hdf_tc_conv3d_fwd: smem attribute: %s", cudaGetErrorString(e)); return HDF_ERR_CUDA; }
    configured = 227 * 1024;
  }
  const int max_ctas = (two_ctas ? 2 : 1) * hdf_sm_count_cached();
  const int grid = p.num_tiles < max_ctas ? p.num_tiles : max_ctas;
  tc_conv_fwd_kernel<<<grid, FWD_THREADS, smem, (cudaStream_t)stream>>>(tmx, tmw, p);
  HDF_LAUNCH_CHECK("hdf_tc_conv3d_fwd");
  if (p.dbg) {   // debug only: synchronous dump of CTA 0's wait-cycle counters
    unsigned long long h[8];
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tc_conv dbg] Cin=%d Cout=%d fold=%d khfold=%d tiles=%d stages=%d | producer: wait_empty=%llu | mma: wait_full=%llu "
            "wait_tempty=%llu mma_issue=%llu commit=%llu total=%llu | epilogue: wait_tfull=%llu total=%llu\n", Cin, Cout, p.fold, p.khfold,
            p.num_tiles, p.stages, h[0], h[2], h[3], h[7], h[1], h[4], h[5], h[6]);
  }
  return HDF_OK;
}

static int wgrad_ok(int c) { return c == 16 || c == 32 || (c % 64 == 0 && c >= 64); }

// the N side of the GEMM (dy channels for mode 0, x channels for the transposed conv) must fit one MMA (<= 256)
// v2 (taps stacked along N) issues 40 % fewer, larger MMAs but re-loads the fixed operand per pass and replicates it
// to fill M; both versions turn out to be bound by L2->SMEM bytes (~2.8 clk per 64-byte box row per SM = the chip's
// ~6.3 TB/s L2 limit; profiles/r1_microbench_conv_v4_wgrad2.txt), where v1 moves fewer bytes -> v1 is the default.
static bool wgrad_use_v1() { static const char* e = getenv("HDF_TC_WGRAD_V2"); return e == nullptr; }
int hdf_tc_wgrad_supported(int mode, int Cin, int Cout) {
  if (mode != 0 && mode != 1) return 0;
  if (!wgrad_use_v1()) return wgrad_ok(Cin) && wgrad_ok(Cout);
  return wgrad_ok(Cin) && wgrad_ok(Cout) && (mode == 0 ? Cout : Cin) <= 256;
}

// Resident CTAs per SM the plan aims for.  Measured (profiles/r1_microbench_conv_v6_wgrad_2cta.txt): one CTA has a single
// TMA-issuing and a single MMA-issuing thread whose dependent waits serialise; two CTAs per SM (<= 112 KB of shared
// memory and <= 256 TMEM columns each) overlap them and make the narrow layers (Cin 16/32: 1.71 -> 0.96 ms at 144^3)
// much faster, at the price of more passes over the (27x smaller) fixed operand.
static int wgrad_target_ctas(int Cin, int Cout) {
  static const int cfg = getenv("HDF_TC_WGRAD_CTAS") ? atoi(getenv("HDF_TC_WGRAD_CTAS")) : 0;
  if (cfg > 0) return cfg;
  (void)Cin; (void)Cout;
  return 2;
}

static int tc_wgrad_plan(int N, int D, int H, int W, int Cin, int Cout, TcWgradParams& p, int ntaps = 27) {
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.ntaps = ntaps;
  p.tap0 = 0;
  p.a_scale = 1;
  static const int kv_cfg = getenv("HDF_TC_WGRAD_KV") ? atoi(getenv("HDF_TC_WGRAD_KV")) : 0;
  p.KV = (Cout > 128) ? 64 : 128;
  if (kv_cfg == 64 || kv_cfg == 128) p.KV = (Cout > 128) ? 64 : kv_cfg;
  // tile = KV voxels
  {
    long long best = -1; p.TD = p.TH = p.TW = 1;
    for (int tw = 1; tw <= p.KV; tw *= 2)
      for (int th = 1; th * tw <= p.KV; th *= 2) {
        const int td = p.KV / (tw * th);
        const long long tiles = (long long)cdiv(D, td) * cdiv(H, th) * cdiv(W, tw);
        if (best < 0 || tiles < best || (tiles == best && tw > p.TW)) { best = tiles; p.TD = td; p.TH = th; p.TW = tw; }
      }
  }
  p.nTd = cdiv(D, p.TD); p.nTh = cdiv(H, p.TH); p.nTw = cdiv(W, p.TW);
  p.num_chunks = N * p.nTd * p.nTh * p.nTw;
  p.CW = Cin < 64 ? Cin : 64;
  p.SPG = 128 / p.CW;
  p.sub_per_tap = Cin / p.CW;
  p.total_sub = ntaps * p.sub_per_tap;
  p.total_groups = cdiv(p.total_sub, p.SPG);
  p.CWn = Cout < 64 ? Cout : 64;
  p.nsub_b = Cout / p.CWn;
  p.a_sub_bytes = (uint32_t)p.KV * p.CW * 2;
  p.a_stage_bytes = p.a_sub_bytes * p.SPG;          // = KV*256 bytes, multiple of 1024
  p.b_sub_bytes = (uint32_t)p.KV * p.CWn * 2;
  p.b_stage_bytes = (p.b_sub_bytes * p.nsub_b + 1023u) & ~1023u;
  static const int smem_kb = getenv("HDF_TC_SMEM_KB") ? atoi(getenv("HDF_TC_SMEM_KB")) : HDF_TC_SMEM_KB_DEFAULT;
  static const int max_stages = getenv("HDF_TC_WGRAD_STAGES") ? atoi(getenv("HDF_TC_WGRAD_STAGES")) : 6;
  int ctas = wgrad_target_ctas(Cin, Cout);
  for (;; --ctas) {
    // per-CTA budgets for `ctas` resident CTAs per SM (227 KB and 512 TMEM columns per SM; ~1.5 KB of barriers/slack each)
    static const int pair_kb = getenv("HDF_TC_WGRAD_PAIR_KB") ? atoi(getenv("HDF_TC_WGRAD_PAIR_KB")) : 224;
    const unsigned budget = ctas <= 1 ? (unsigned)smem_kb * 1024u : (unsigned)(pair_kb * 1024 / ctas) - 2048u;
    int tcols = 512;
    for (int c = 1; c < ctas; c *= 2) tcols /= 2;                 // 512, 256, 128, 128 ...
    p.groups_per_pass = tcols / Cout;
    if (p.groups_per_pass > p.total_groups) p.groups_per_pass = p.total_groups;
    const long long room = (long long)budget - 2ll * p.b_stage_bytes;
    p.a_stages = room > 0 ? (int)(room / p.a_stage_bytes) : 0;
    if (p.a_stages > max_stages) p.a_stages = max_stages;
    if (ctas <= 1 || (p.a_stages >= 2 && p.groups_per_pass >= 1)) break;
  }
  if (p.a_stages < 2) p.a_stages = 2;
  if (p.groups_per_pass < 1) p.groups_per_pass = 1;
  const int ia = p.CW * 2, ib = p.CWn * 2;
  p.a_layout = ia == 128 ? 2u : ia == 64 ? 4u : 6u;
  p.b_layout = ib == 128 ? 2u : ib == 64 ? 4u : 6u;
  p.a_sbo = 8u * ia;
  p.b_sbo = 8u * ib;
  uint32_t cols = 32;
  while (cols < (uint32_t)(p.groups_per_pass * Cout)) cols *= 2;
  p.tmem_cols = cols;
  const int passes = cdiv(p.total_groups, p.groups_per_pass);
  // split-K: ~2 CTAs per SM in total (one wave when two are resident), at least 8 chunks per slab
  int slabs = cdiv(2 * hdf_sm_count_cached(), passes);
  const int max_slabs = p.num_chunks / 8 > 0 ? p.num_chunks / 8 : 1;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  p.chunks_per_slab = cdiv(p.num_chunks, slabs);
  p.num_slabs = cdiv(p.num_chunks, p.chunks_per_slab);
  return passes;
}

// ---- v2 planning: which operand is shifted (S, N side) and which is fixed (F, M side)
struct Wg2Roles { bool s_is_x; int Cs, Cf, s_scale, flip; };
static Wg2Roles wg2_roles(int mode, int Cin, int Cout) {
  Wg2Roles r;
  if (mode == 1) { r.s_is_x = false; r.Cs = Cout; r.Cf = Cin; r.s_scale = 2; r.flip = 0; }      // dW = sum_i x[i] dy[2i-1+k]
  else if (Cout < Cin) { r.s_is_x = false; r.Cs = Cout; r.Cf = Cin; r.s_scale = 1; r.flip = 1; } // shift dy (tap' = 26-tap)
  else { r.s_is_x = true; r.Cs = Cin; r.Cf = Cout; r.s_scale = 1; r.flip = 0; }
  return r;
}
static int wg2_supported(int mode, int Cin, int Cout) {
  if (mode != 0 && mode != 1) return 0;
  return wgrad_ok(Cin) && wgrad_ok(Cout);
}
static int tc_wgrad2_plan(int N, int D, int H, int W, int Cs, int Cf, TcWgrad2Params& p) {
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cs = Cs; p.Cf = Cf;
  p.KV = 64;
  {
    long long best = -1; p.TD = p.TH = p.TW = 1;
    for (int tw = 1; tw <= p.KV; tw *= 2)
      for (int th = 1; th * tw <= p.KV; th *= 2) {
        const int td = p.KV / (tw * th);
        const long long tiles = (long long)cdiv(D, td) * cdiv(H, th) * cdiv(W, tw);
        if (best < 0 || tiles < best || (tiles == best && tw > p.TW)) { best = tiles; p.TD = td; p.TH = th; p.TW = tw; }
      }
  }
  p.nTd = cdiv(D, p.TD); p.nTh = cdiv(H, p.TH); p.nTw = cdiv(W, p.TW);
  p.num_chunks = N * p.nTd * p.nTh * p.nTw;
  p.CWs = Cs < 64 ? Cs : 64;
  p.spt = Cs / p.CWs;
  p.total_sub = 27 * p.spt;
  p.SPGn = 256 / p.CWs;
  p.G = cdiv(p.total_sub, p.SPGn);
  p.CWf = Cf < 64 ? Cf : 64;
  p.nsub_f = Cf / p.CWf;
  p.nrep_f = Cf < 128 ? 128 / Cf : 1;
  p.nMh = Cf < 128 ? 1 : Cf / 128;
  p.mh_per_pass = p.nMh < 2 ? p.nMh : 2;
  p.groups_per_pass = p.mh_per_pass == 1 ? 2 : 1;
  if (p.groups_per_pass > p.G) p.groups_per_pass = p.G;
  p.passes_g = cdiv(p.G, p.groups_per_pass);
  p.passes_m = cdiv(p.nMh, p.mh_per_pass);
  p.s_sub_bytes = (uint32_t)p.KV * p.CWs * 2;
  p.s_stage_bytes = p.s_sub_bytes * p.SPGn;            // KV * 512 bytes = 32 KB
  p.f_sub_bytes = (uint32_t)p.KV * p.CWf * 2;
  p.f_stage_bytes = p.f_sub_bytes * (128 / p.CWf) * p.mh_per_pass;
  p.s_stages = (int)((200u * 1024u - 2u * p.f_stage_bytes) / p.s_stage_bytes);
  if (p.s_stages > 5) p.s_stages = 5;
  const int is = p.CWs * 2, jf = p.CWf * 2;
  p.s_layout = is == 128 ? 2u : is == 64 ? 4u : 6u;
  p.f_layout = jf == 128 ? 2u : jf == 64 ? 4u : 6u;
  p.s_sbo = 8u * is;
  p.f_sbo = 8u * jf;
  uint32_t need = (uint32_t)(p.groups_per_pass * p.mh_per_pass * 256), cols = 32;
  while (cols < need) cols *= 2;
  p.tmem_cols = cols;
  const int passes = p.passes_g * p.passes_m;
  int slabs = cdiv(2 * hdf_sm_count_cached(), passes);
  const int max_slabs = p.num_chunks / 16 > 0 ? p.num_chunks / 16 : 1;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  p.chunks_per_slab = cdiv(p.num_chunks, slabs);
  p.num_slabs = cdiv(p.num_chunks, p.chunks_per_slab);
  return passes;
}

static int tc_wgrad2_launch(int mode, const void* x, long long ldx, const void* dy, long long ldy, float* dw, long long stride_ci,
                            long long stride_co, int N, int Do, int Ho, int Wo, int Cin, int Cout, void* workspace,
                            size_t ws_bytes, int accumulate, cudaStream_t stream) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { hdf_set_error("hdf_tc_conv3d_wgrad: cuTensorMapEncodeTiled unavailable"); return HDF_ERR_CUDA; }
  const Wg2Roles r = wg2_roles(mode, Cin, Cout);
  const int D = mode == 0 ? Do : Do / 2, H = mode == 0 ? Ho : Ho / 2, W = mode == 0 ? Wo : Wo / 2;   // base grid
  TcWgrad2Params p;
  const int passes = tc_wgrad2_plan(N, D, H, W, r.Cs, r.Cf, p);
  p.s_scale = r.s_scale;
  HDF_REQUIRE(ws_bytes >= (size_t)p.num_slabs * 27 * Cin * Cout * sizeof(float), "hdf_tc_conv3d_wgrad: workspace too small");
  p.partial = (float*)workspace;
  const void* s_ptr = r.s_is_x ? x : dy;
  const void* f_ptr = r.s_is_x ? dy : x;
  const long long s_ld = r.s_is_x ? ldx : ldy, f_ld = r.s_is_x ? ldy : ldx;
  CUtensorMap tms, tmf;
  for (int which = 0; which < 2; ++which) {
    const void* base = which == 0 ? s_ptr : f_ptr;
    const long long ld = which == 0 ? s_ld : f_ld;
    const int C = which == 0 ? r.Cs : r.Cf;
    const int cw = which == 0 ? p.CWs : p.CWf;
    const cuuint32_t es = which == 0 ? (cuuint32_t)p.s_scale : 1u;
    const int Dt = D * (int)es, Ht = H * (int)es, Wt = W * (int)es;
    cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)Wt, (cuuint64_t)Ht, (cuuint64_t)Dt, (cuuint64_t)N};
    cuuint64_t gstr[4] = {(cuuint64_t)ld * 2, (cuuint64_t)Wt * ld * 2, (cuuint64_t)Ht * Wt * ld * 2,
                          (cuuint64_t)Dt * Ht * Wt * ld * 2};
    cuuint32_t box[5] = {(cuuint32_t)cw, (cuuint32_t)p.TW * es, (cuuint32_t)p.TH * es, (cuuint32_t)p.TD * es, 1};
    cuuint32_t estr[5] = {1, es, es, es, 1};
    CUresult rr = enc(which == 0 ? &tms : &tmf, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box,
                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(cw * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rr != CUDA_SUCCESS) { hdf_set_error("hdf_tc_conv3d_wgrad: encode failed: %d", (int)rr); return HDF_ERR_CUDA; }
  }
  const size_t smem = (size_t)p.s_stages * p.s_stage_bytes + 2 * (size_t)p.f_stage_bytes + 1024 + 8 * (2 * p.s_stages + 6) + 64;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    // full 228 KB carve-out even when the CTA asks for less: the driver otherwise picks the smallest configuration that
    // fits THIS kernel (e.g. 196 KB for a 194 KB CTA), and the side streams' kernels find no shared memory left on the SM
    if (e == cudaSuccess && getenv("HDF_NO_MAX_CARVEOUT") == nullptr)
      e = cudaFuncSetAttribute(tc_conv_wgrad2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) { hdf_set_error("hdf_tc_conv3d_wgrad: smem attribute: %s", cudaGetErrorString(e)); return HDF_ERR_CUDA; }
    configured = true;
  }
  HDF_REQUIRE(smem <= 227 * 1024 && p.s_stages >= 2, "hdf_tc_conv3d_wgrad: smem plan does not fit (%zu bytes, %d stages)", smem, p.s_stages);
  dim3 grid(p.num_slabs, passes);
  tc_conv_wgrad2_kernel<<<grid, TC_THREADS, smem, stream>>>(tms, tmf, p);
  HDF_LAUNCH_CHECK("hdf_tc_conv3d_wgrad(v2)");
  // partial[slab][tap][s][f] -> dw: (s,f) = (ci,co) when S is x, (co,ci) otherwise
  const long long ss = r.s_is_x ? stride_ci : stride_co, sf = r.s_is_x ? stride_co : stride_ci;
  const long long per = 27ll * Cin * Cout;
  tc_wgrad_reduce_kernel<<<min(2048, cdiv(per, 256)), 256, 0, stream>>>((const float*)workspace, dw, p.num_slabs, r.Cs, r.Cf, ss,
                                                                      sf, accumulate, r.flip);
  HDF_LAUNCH_CHECK("hdf_tc_conv3d_wgrad(v2)/reduce");
  return HDF_OK;
}

// mode 0 with fewer output than input channels: shift dy instead of x (dW[tap] = sum_j x[j] dy[j - off(tap)]), which
// makes the 27-times-reloaded operand the narrower one (halves L2->SMEM traffic for 64->32, 128->64, 256->128)
static bool wgrad_swap_roles(int mode, int Cin, int Cout) { return mode == 0 && Cout < Cin && Cin <= 256; }

size_t hdf_tc_wgrad_workspace(int mode, int N, int Do, int Ho, int Wo, int Cin, int Cout) {
  if (!hdf_tc_wgrad_supported(mode, Cin, Cout)) return 0;
  if (hdf_tc_wgrad_ws_supported(mode, Cin, Cout)) return hdf_tc_wgrad_ws_workspace(N, Do, Ho, Wo, Cin, Cout);
  if (!wgrad_use_v1()) {
    const Wg2Roles r = wg2_roles(mode, Cin, Cout);
    TcWgrad2Params q;
    if (mode == 0) tc_wgrad2_plan(N, Do, Ho, Wo, r.Cs, r.Cf, q);
    else tc_wgrad2_plan(N, Do / 2, Ho / 2, Wo / 2, r.Cs, r.Cf, q);
    return (size_t)q.num_slabs * 27 * Cin * Cout * sizeof(float);
  }
  TcWgradParams p;
  const int ntaps = (mode == 0 && Do == 1) ? 9 : 27;      // one-plane volumes: the kd = 1 taps only (same plan as the launch)
  if (mode == 0 && !wgrad_swap_roles(mode, Cin, Cout)) tc_wgrad_plan(N, Do, Ho, Wo, Cin, Cout, p, ntaps);
  else if (mode == 0) tc_wgrad_plan(N, Do, Ho, Wo, Cout, Cin, p, ntaps);
  else tc_wgrad_plan(N, Do / 2, Ho / 2, Wo / 2, Cout, Cin, p);
  return (size_t)p.num_slabs * 27 * Cin * Cout * sizeof(float);
}

int hdf_tc_conv3d_wgrad(int mode, const void* x, long long ldx, const void* dy, long long ldy, float* dw, long long stride_ci,
                        long long stride_co, int N, int Do, int Ho, int Wo, int Cin, int Cout, void* workspace,
                        size_t ws_bytes, int accumulate, void* stream) {
  if (!hdf_tc_wgrad_supported(mode, Cin, Cout)) {
    hdf_set_error("hdf_tc_conv3d_wgrad: unsupported mode/channels mode=%d Cin=%d Cout=%d", mode, Cin, Cout);
    return HDF_ERR_UNSUPPORTED;
  }
  HDF_REQUIRE(x && dy && dw && workspace, "hdf_tc_conv3d_wgrad: null pointer");
  HDF_REQUIRE((ldx % 8 == 0) && (ldy % 8 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)dy % 16 == 0),
              "hdf_tc_conv3d_wgrad: operands must be 16-byte aligned with channel strides multiple of 8");
  if (hdf_tc_wgrad_ws_supported(mode, Cin, Cout))
    return hdf_tc_wgrad_ws(x, ldx, dy, ldy, dw, stride_ci, stride_co, N, Do, Ho, Wo, Cin, Cout, workspace, ws_bytes, accumulate, stream);
  if (!wgrad_use_v1())
    return tc_wgrad2_launch(mode, x, ldx, dy, ldy, dw, stride_ci, stride_co, N, Do, Ho, Wo, Cin, Cout, workspace, ws_bytes,
                            accumulate, (cudaStream_t)stream);
  EncodeTiledFn enc = get_encode();
  if (!enc) { hdf_set_error("hdf_tc_conv3d_wgrad: cuTensorMapEncodeTiled unavailable"); return HDF_ERR_CUDA; }
  // A side = the shifted operand whose taps are stacked along M; B side = the fixed tile (GEMM N)
  //   mode 0: A = x (shift k-1),            B = dy ; base grid = output grid
  //   mode 1: A = dy (2x res, shift 2j+k-1), B = x  ; base grid = input grid        (dW[ci][co][k] = sum_i x[i] dy[2i-1+k])
  //   mode 0 swapped (Cout < Cin): A = dy (shift -(k-1), i.e. tap' = 26 - tap), B = x
  const bool swap = wgrad_swap_roles(mode, Cin, Cout);
  const bool a_is_x = (mode == 0 && !swap);
  const int D = mode == 0 ? Do : Do / 2, H = mode == 0 ? Ho : Ho / 2, W = mode == 0 ? Wo : Wo / 2;
  const void* a_ptr = a_is_x ? x : dy;
  const void* b_ptr = a_is_x ? dy : x;
  const long long a_ld = a_is_x ? ldx : ldy, b_ld = a_is_x ? ldy : ldx;
  const int Ca = a_is_x ? Cin : Cout, Cb = a_is_x ? Cout : Cin;
  const long long sa = a_is_x ? stride_ci : stride_co, sb = a_is_x ? stride_co : stride_ci;
  TcWgradParams p;
  // one-plane volumes (the 2-D model as flat volumes): only the kd = 1 taps see anything but zero padding
  const bool flat = mode == 0 && D == 1;
  const int passes = tc_wgrad_plan(N, D, H, W, Ca, Cb, p, flat ? 9 : 27);
  if (flat) p.tap0 = 9;
  p.a_scale = mode == 0 ? 1 : 2;
  HDF_REQUIRE(ws_bytes >= (size_t)p.num_slabs * 27 * Cin * Cout * sizeof(float), "hdf_tc_conv3d_wgrad: workspace too small");
  p.partial = (float*)workspace;
  CUtensorMap tmx, tmdy;
  for (int which = 0; which < 2; ++which) {
    const void* base = which == 0 ? a_ptr : b_ptr;
    const long long ld = which == 0 ? a_ld : b_ld;
    const int C = which == 0 ? Ca : Cb;
    const int cw = which == 0 ? p.CW : p.CWn;
    const cuuint32_t es = which == 0 ? (cuuint32_t)p.a_scale : 1u;
    const int Dt = D * (int)es, Ht = H * (int)es, Wt = W * (int)es;
    cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)Wt, (cuuint64_t)Ht, (cuuint64_t)Dt, (cuuint64_t)N};
    cuuint64_t gstr[4] = {(cuuint64_t)ld * 2, (cuuint64_t)Wt * ld * 2, (cuuint64_t)Ht * Wt * ld * 2,
                          (cuuint64_t)Dt * Ht * Wt * ld * 2};
    cuuint32_t box[5] = {(cuuint32_t)cw, (cuuint32_t)p.TW * es, (cuuint32_t)p.TH * es, (cuuint32_t)p.TD * es, 1};
    cuuint32_t estr[5] = {1, es, es, es, 1};
    CUresult r = enc(which == 0 ? &tmx : &tmdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(cw * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hdf_set_error("hdf_tc_conv3d_wgrad: encode failed: %d", (int)r); return HDF_ERR_CUDA; }
  }
  const size_t smem = (size_t)p.a_stages * p.a_stage_bytes + 2 * (size_t)p.b_stage_bytes + 1024 + 8 * (2 * p.a_stages + 6) + 16;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    // full 228 KB carve-out even when the CTA asks for less: the driver otherwise picks the smallest configuration that
    // fits THIS kernel (e.g. 196 KB for a 194 KB CTA), and the side streams' kernels find no shared memory left on the SM
    if (e == cudaSuccess && getenv("HDF_NO_MAX_CARVEOUT") == nullptr)
      e = cudaFuncSetAttribute(tc_conv_wgrad_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) { hdf_set_error("hdf_tc_conv3d_wgrad: smem attribute: %s", cudaGetErrorString(e)); return HDF_ERR_CUDA; }
    configured = true;
  }
  HDF_REQUIRE(smem <= 227 * 1024, "hdf_tc_conv3d_wgrad: smem plan too large (%zu)", smem);
  dim3 grid(p.num_slabs, passes);
  tc_conv_wgrad_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(tmx, tmdy, p);
  HDF_LAUNCH_CHECK("hdf_tc_conv3d_wgrad");
  if (flat && !accumulate) {
    // the 18 taps of the two padding planes are exactly zero (stride-1 conv weights are one dense [Cout][Cin][27] block)
    cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)27 * Cin * Cout * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) { hdf_set_error("hdf_tc_conv3d_wgrad: memset failed: %s", cudaGetErrorString(e)); return HDF_ERR_CUDA; }
  }
  const long long per = (long long)p.ntaps * Cin * Cout;
  tc_wgrad_reduce_kernel<<<min(2048, cdiv(per, 256)), 256, 0, (cudaStream_t)stream>>>((const float*)workspace, dw, p.num_slabs,
                                                                                     Ca, Cb, sa, sb, accumulate, swap ? 1 : 0,
                                                                                     p.ntaps, p.tap0);
  HDF_LAUNCH_CHECK("hdf_tc_conv3d_wgrad/reduce");
  return HDF_OK;
}


// ---- stem path (see stem_im2col_kernel)
int hdf_stem_kp(int Cin) { return (Cin >= 1 && Cin <= 4) ? (27 * Cin + 63) / 64 * 64 : 0; }

int hdf_stem_supported(int Cin, int Cout) { return hdf_stem_kp(Cin) > 0 && (Cout == 16 || Cout == 32 || Cout == 64); }

int hdf_stem_im2col(const float* x_ncdhw, void* xcol_bf16, int N, int Cin, int D, int H, int W, void* stream) {
  HDF_REQUIRE(x_ncdhw && xcol_bf16 && hdf_stem_kp(Cin) > 0, "hdf_stem_im2col: bad args (Cin must be 1..4)");
  const int Kp = hdf_stem_kp(Cin);
  const unsigned grid = (unsigned)((long long)N * D * H);
  const size_t smem = (size_t)9 * Cin * (W + 2) * sizeof(float);
  HDF_REQUIRE(smem <= 200 * 1024, "hdf_stem_im2col: W=%d too large", W);
  cudaStream_t s = (cudaStream_t)stream;
#define HDF_STEM_IM2COL(CI)                                                                                              \
  {                                                                                                                      \
    if (smem > 48 * 1024) cudaFuncSetAttribute(stem_im2col_kernel<CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    stem_im2col_kernel<CI><<<grid, 256, smem, s>>>(x_ncdhw, (bf16*)xcol_bf16, D, H, W, Kp);                               \
  }
  switch (Cin) {
    case 1: HDF_STEM_IM2COL(1) break;
    case 2: HDF_STEM_IM2COL(2) break;
    case 3: HDF_STEM_IM2COL(3) break;
    default: HDF_STEM_IM2COL(4) break;
  }
#undef HDF_STEM_IM2COL
  HDF_LAUNCH_CHECK("hdf_stem_im2col");
  return HDF_OK;
}

int hdf_stem_pack_weights(const float* w, void* packed_bf16, int Cin, int Cout, void* stream) {
  HDF_REQUIRE(w && packed_bf16 && hdf_stem_kp(Cin) > 0, "hdf_stem_pack_weights: bad args");
  const int Kp = hdf_stem_kp(Cin);
  stem_pack_kernel<<<cdiv((long long)Cout * Kp, 256), 256, 0, (cudaStream_t)stream>>>(w, (bf16*)packed_bf16, Cin, Cout, Kp);
  HDF_LAUNCH_CHECK("hdf_stem_pack_weights");
  return HDF_OK;
}

int hdf_stem_conv_fwd(const void* xcol_bf16, const void* w_packed_bf16, void* y, long long ldy, int N, int D, int H, int W,
                      int Cin, int Cout, void* stream) {
  HDF_REQUIRE(hdf_stem_supported(Cin, Cout), "hdf_stem_conv_fwd: unsupported Cin=%d Cout=%d", Cin, Cout);
  const int Kp = hdf_stem_kp(Cin);
  return tc_conv_fwd_launch(3, xcol_bf16, Kp, w_packed_bf16, nullptr, y, ldy, N, D, H, W, Kp, Cout, stream);
}

size_t hdf_stem_wgrad_workspace(int N, int D, int H, int W, int Cin, int Cout) {
  if (!hdf_stem_supported(Cin, Cout)) return 0;
  TcWgradParams p;
  tc_wgrad_plan(N, D, H, W, hdf_stem_kp(Cin), Cout, p, 1);
  return (size_t)p.num_slabs * hdf_stem_kp(Cin) * Cout * sizeof(float);
}

// dw [Cout][Cin][27] (+)= sum_v dy[v][co] * xcol[v][tap*Cin + ci]
int hdf_stem_conv_wgrad(const void* xcol_bf16, const void* dy, long long ldy, float* dw, int N, int D, int H, int W, int Cin,
                        int Cout, void* workspace, size_t ws_bytes, int accumulate, void* stream) {
  HDF_REQUIRE(hdf_stem_supported(Cin, Cout), "hdf_stem_conv_wgrad: unsupported Cin=%d Cout=%d", Cin, Cout);
  HDF_REQUIRE(xcol_bf16 && dy && dw && workspace && (ldy % 8 == 0) && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)xcol_bf16 % 16 == 0),
              "hdf_stem_conv_wgrad: bad args");
  EncodeTiledFn enc = get_encode();
  if (!enc) { hdf_set_error("hdf_stem_conv_wgrad: cuTensorMapEncodeTiled unavailable"); return HDF_ERR_CUDA; }
  const int Kp = hdf_stem_kp(Cin);
  TcWgradParams p;
  const int passes = tc_wgrad_plan(N, D, H, W, Kp, Cout, p, 1);
  HDF_REQUIRE(ws_bytes >= (size_t)p.num_slabs * Kp * Cout * sizeof(float), "hdf_stem_conv_wgrad: workspace too small");
  p.partial = (float*)workspace;
  CUtensorMap tmx, tmdy;
  for (int which = 0; which < 2; ++which) {
    const void* base = which == 0 ? xcol_bf16 : dy;
    const long long ld = which == 0 ? Kp : ldy;
    const int C = which == 0 ? Kp : Cout;
    const int cw = which == 0 ? p.CW : p.CWn;
    cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t gstr[4] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2, (cuuint64_t)D * H * W * ld * 2};
    cuuint32_t box[5] = {(cuuint32_t)cw, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TD, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(which == 0 ? &tmx : &tmdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(cw * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hdf_set_error("hdf_stem_conv_wgrad: encode failed: %d", (int)r); return HDF_ERR_CUDA; }
  }
  const size_t smem = (size_t)p.a_stages * p.a_stage_bytes + 2 * (size_t)p.b_stage_bytes + 1024 + 8 * (2 * p.a_stages + 6) + 16;
  cudaError_t e = cudaFuncSetAttribute(tc_conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
  if (e == cudaSuccess && getenv("HDF_NO_MAX_CARVEOUT") == nullptr)
    e = cudaFuncSetAttribute(tc_conv_wgrad_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) { hdf_set_error("hdf_stem_conv_wgrad: smem attribute: %s", cudaGetErrorString(e)); return HDF_ERR_CUDA; }
  HDF_REQUIRE(smem <= 227 * 1024, "hdf_stem_conv_wgrad: smem plan too large (%zu)", smem);
  dim3 grid(p.num_slabs, passes);
  tc_conv_wgrad_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(tmx, tmdy, p);
  HDF_LAUNCH_CHECK("hdf_stem_conv_wgrad");
  stem_wgrad_reduce_kernel<<<cdiv(27ll * Cin * Cout, 128), 128, 0, (cudaStream_t)stream>>>((const float*)workspace, dw, p.num_slabs,
                                                                                          Cin, Cout, Kp, accumulate);
  HDF_LAUNCH_CHECK("hdf_stem_conv_wgrad/reduce");
  return HDF_OK;
}

}  // extern "C"

// C++-linkage helper for stem_tc.cu: partial [S][Kp][Cout] -> dw [Cout][Cin][27]
int hdf_stem_wgrad_reduce(const float* part, float* dw, int S, int Cin, int Cout, int Kp, int accumulate, void* stream) {
  stem_wgrad_reduce_kernel<<<cdiv(27ll * Cin * Cout, 128), 128, 0, (cudaStream_t)stream>>>(part, dw, S, Cin, Cout, Kp, accumulate);
  HDF_LAUNCH_CHECK("hdf_stem_fused_wgrad/reduce");
  return HDF_OK;
}
