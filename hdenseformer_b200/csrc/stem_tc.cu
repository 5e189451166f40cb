// First convolution of the network, block_1_1_left.conv = Conv3d(M -> nf, k3, p1) with M = 1..4 input modalities
// (models/HDenseFormer.py:151-152 at :184), forward and weight gradient, WITHOUT the im2col matrix.
//
// Round 1 gathered the 27*M taps of every voxel into Xcol [N*V, Kp] bf16 once per step (764 MB at 2 x 144^3, 0.46 ms to
// write, read again by the GEMM and by the weight gradient, kept alive through the backward pass).  Here the A operand
// tile (128 voxels x Kp taps) is assembled in shared memory by producer threads straight from the fp32 NCDHW volume
// (48 MB, L1/L2-resident neighbourhoods): gather -> bf16 -> the 128-byte-swizzled UMMA layout, fence.proxy.async, mbarrier.
//   forward : D[128 voxels, Cout] = A[128, Kp] W^T              (A K-major, weights resident in shared memory)
//   wgrad   : D[Kp, Cout]        += A^T[Kp, 128] dY[128, Cout]   (the SAME shared-memory image read as an MN-major operand,
//                                                                dY tiles by TMA; fp32 partial per CTA, fixed-order reduce)
// k = tap * M + ci, zero-padded to Kp = 64 (M <= 2) or 128 (M = 3, 4) -- the packing of hdf_stem_pack_weights.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

int hdf_sm_count_cached();

namespace {
using namespace tcptx;

// Producer threads are organised in independent groups, each assembling whole tiles (tile t of the CTA goes to group
// t % G): a tile costs a full L2 round trip (index math -> 27*M gathers -> convert -> store -> fence -> arrive), so several
// tiles must be in flight per SM.  Kp = 64: 3 groups x 128 threads (a thread gathers all 64 taps of its voxel row);
// Kp = 128: 2 groups x 256 threads (two threads per row, 64 taps each).
template <int CIN> struct SfCfg {
  static constexpr int KP = CIN <= 2 ? 64 : 128;
  static constexpr int TPR = KP / 64;                 // threads per voxel row
  static constexpr int GROUPS = CIN <= 2 ? 3 : 2;
  static constexpr int GROUP_THREADS = 128 * TPR;
  static constexpr int PROD = GROUPS * GROUP_THREADS;  // 384 / 512
  static constexpr int STAGES = CIN <= 2 ? 3 : 4;
  static constexpr int FWD_THREADS = PROD + 32 + 128;  // + MMA warp + 4 epilogue warps
  static constexpr int WG_THREADS = PROD + 64 + 128;   // + MMA warp + TMA warp + 4 epilogue warps
};

struct SfParams {
  const float* x;                    // [N][CIN][D][H][W] fp32
  int N, D, H, W, Cout;
  long long V, Vtot;                 // D*H*W, N*V
  int num_tiles;
  int tiles_per_cta;                 // 0: persistent (CTA b takes tiles b, b + grid, ...); else CTA b takes [b*T, (b+1)*T)
  const bf16* wp;                    // fwd: packed weights [Cout][Kp]
  bf16* y; long long ldy;            // fwd: output [Vtot][ldy]
  float* partial;                    // wgrad: [gridDim.x][Kp][Cout]
};

__device__ __forceinline__ int sf_tile0(const SfParams& p) { return p.tiles_per_cta ? blockIdx.x * p.tiles_per_cta : blockIdx.x; }
__device__ __forceinline__ int sf_tstep(const SfParams& p) { return p.tiles_per_cta ? 1 : gridDim.x; }
__device__ __forceinline__ int sf_tend(const SfParams& p) {
  return p.tiles_per_cta ? min(p.num_tiles, blockIdx.x * p.tiles_per_cta + p.tiles_per_cta) : p.num_tiles;
}

// 128-voxel x KP-tap tile row r of tile `tile`, taps [K0, K0 + NK) -> swizzled shared memory at `stage`
template <int CIN, int K0, int NK>
__device__ __forceinline__ void sf_gather_store(const SfParams& p, const float* __restrict__ xv, bool valid, int d, int h, int w,
                                                uint32_t stage, int r, uint32_t empty_bar, uint32_t parity) {
  const long long HW = (long long)p.H * p.W;
  const bool vd[3] = {d >= 1, true, d <= p.D - 2}, vh[3] = {h >= 1, true, h <= p.H - 2}, vw[3] = {w >= 1, true, w <= p.W - 2};
  float f[NK];
#pragma unroll
  for (int j = 0; j < NK; ++j) {
    const int k = K0 + j;
    const int tap = k / CIN, ci = k % CIN;
    if (tap < 27) {
      const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
      const bool ok = valid && vd[kd] && vh[kh] && vw[kw];
      f[j] = ok ? __ldg(xv + ci * p.V + (kd - 1) * HW + (kh - 1) * (long long)p.W + (kw - 1)) : 0.f;
    } else {
      f[j] = 0.f;
    }
  }
  mbar_wait(empty_bar, parity);
  const uint32_t row = stage + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
#pragma unroll
  for (int c = 0; c < NK / 8; ++c) {
    const int chunk = (K0 + c * 8) / 8;                 // 16-byte chunk of the Kp-wide row
    const uint32_t dst = row + (uint32_t)(chunk >> 3) * 16384u + (uint32_t)(((chunk & 7) ^ (r & 7)) * 16);
    __nv_bfloat162 h0 = __floats2bfloat162_rn(f[c * 8 + 0], f[c * 8 + 1]), h1 = __floats2bfloat162_rn(f[c * 8 + 2], f[c * 8 + 3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[c * 8 + 4], f[c * 8 + 5]), h3 = __floats2bfloat162_rn(f[c * 8 + 6], f[c * 8 + 7]);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(*reinterpret_cast<uint32_t*>(&h0)),
                 "r"(*reinterpret_cast<uint32_t*>(&h1)), "r"(*reinterpret_cast<uint32_t*>(&h2)), "r"(*reinterpret_cast<uint32_t*>(&h3))
                 : "memory");
  }
}

// producer loop shared by both kernels: group g of the CTA assembles tiles g, g + G, ... ; GROUP_THREADS arrivals per stage
template <int CIN>
__device__ __forceinline__ void sf_producer(const SfParams& p, uint32_t ring_base, uint32_t stage_bytes, uint32_t full0, uint32_t empty0) {
  using C = SfCfg<CIN>;
  const int grp = threadIdx.x / C::GROUP_THREADS, gt = threadIdx.x % C::GROUP_THREADS;
  const int r = gt & 127, half = gt >> 7;
  const unsigned HWu = (unsigned)(p.H * p.W), Vu = (unsigned)p.V;
  const int tile0 = p.tiles_per_cta ? blockIdx.x * p.tiles_per_cta : blockIdx.x;
  const int tstep = p.tiles_per_cta ? 1 : gridDim.x;
  const int tend = p.tiles_per_cta ? min(p.num_tiles, tile0 + p.tiles_per_cta) : p.num_tiles;
  int t = grp;                                   // sequence number of the tile within this CTA
  for (int tile = tile0 + grp * tstep; tile < tend; tile += C::GROUPS * tstep, t += C::GROUPS) {
    const uint32_t s = (uint32_t)(t % C::STAGES), ph = (uint32_t)((t / C::STAGES) & 1);
    const long long v = (long long)tile * 128 + r;
    const bool valid = v < p.Vtot;
    const unsigned vv = valid ? (unsigned)v : 0u;          // Vtot < 2^31 (checked by the host): 32-bit divisions
    const unsigned n = vv / Vu, rem = vv - n * Vu;
    const int d = (int)(rem / HWu);
    const unsigned hw = rem - (unsigned)d * HWu;
    const int h = (int)(hw / (unsigned)p.W), w = (int)(hw - (unsigned)h * (unsigned)p.W);
    const float* xv = p.x + (long long)n * CIN * p.V + rem;
    const uint32_t stage = ring_base + s * stage_bytes;
    if (C::TPR == 1) sf_gather_store<CIN, 0, 64>(p, xv, valid, d, h, w, stage, r, empty0 + 8u * s, ph ^ 1u);
    else if (half == 0) sf_gather_store<CIN, 0, 64>(p, xv, valid, d, h, w, stage, r, empty0 + 8u * s, ph ^ 1u);
    else sf_gather_store<CIN, C::KP - 64, 64>(p, xv, valid, d, h, w, stage, r, empty0 + 8u * s, ph ^ 1u);
    fence_proxy_async();
    mbar_arrive(full0 + 8u * s);
  }
}

template <int CIN>
__global__ void __launch_bounds__(SfCfg<CIN>::FWD_THREADS, 1) stem_tc_fwd_kernel(const SfParams p) {
  using C = SfCfg<CIN>;
  constexpr int KP = C::KP, NSUB = KP / 64, SF_STAGES = C::STAGES, PW = C::PROD / 32;   // PW = producer warps
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t stage_bytes = NSUB * 16384u;
  const uint32_t wsub_bytes = (uint32_t)p.Cout * 128u;                 // one 64-k weight sub-tile [Cout rows x 128 B]
  const uint32_t w_base = smem_base;
  const uint32_t ring_base = w_base + ((NSUB * wsub_bytes + 1023u) & ~1023u);
  const uint32_t bar_base = ring_base + SF_STAGES * stage_bytes;
  const uint32_t full0 = bar_base, empty0 = bar_base + 8u * SF_STAGES;
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * SF_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * SF_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * SF_STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  const uint32_t acc_cols = p.Cout < 32 ? 32u : (uint32_t)p.Cout;      // accumulator stride (TMEM allocations are >= 32 columns)
  const uint32_t tmem_cols = 2 * acc_cols <= 64 ? 64u : 128u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SF_STAGES; ++s) { mbar_init(full0 + 8u * s, C::GROUP_THREADS); mbar_init(empty0 + 8u * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    fence_barrier_init();
  }
  if (warp == PW) tmem_alloc(tmem_slot, tmem_cols);
  // resident weights: [Cout][Kp] bf16 -> NSUB K-major swizzled sub-tiles
  for (int i = threadIdx.x; i < p.Cout * (KP / 8); i += C::FWD_THREADS) {
    const int row = i / (KP / 8), chunk = i % (KP / 8);
    const uint4 v = *reinterpret_cast<const uint4*>(p.wp + (size_t)row * KP + chunk * 8);
    *reinterpret_cast<uint4*>(smem_gen + (w_base - smem_base) + (uint32_t)(chunk >> 3) * wsub_bytes + (uint32_t)(row >> 3) * 1024u +
                              (uint32_t)(row & 7) * 128u + (uint32_t)(((chunk & 7) ^ (row & 7)) * 16)) = v;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < PW) {
    sf_producer<CIN>(p, ring_base, stage_bytes, full0, empty0);
  } else if (warp == PW) {
    const uint32_t issue = elect_one_sync() ? 1u : 0u;
    const uint32_t idesc = umma_idesc(128, p.Cout, 0, 0);
    const uint64_t desc_hi = umma_desc(0, 16, 1024, 2);
    uint32_t s = 0, ph = 0;
    int acc = 0; uint32_t accph = 0;
    for (int tile = sf_tile0(p); tile < sf_tend(p); tile += sf_tstep(p)) {
      mbar_wait(tempty_bar(acc), accph ^ 1u);
      mbar_wait(full0 + 8u * s, ph);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_cols;
#pragma unroll
      for (int sub = 0; sub < NSUB; ++sub) {
        const uint64_t adesc = desc_hi | (uint64_t)(((ring_base + s * stage_bytes + sub * 16384u) >> 4) & 0x3FFF);
        const uint64_t bdesc = desc_hi | (uint64_t)(((w_base + sub * wsub_bytes) >> 4) & 0x3FFF);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss_p(d_tmem, adesc + (uint64_t)(2 * ks), bdesc + (uint64_t)(2 * ks), idesc, (sub > 0 || ks > 0) ? 1u : 0u, issue);
      }
      umma_commit_p(empty0 + 8u * s, issue);
      umma_commit_p(tfull_bar(acc), issue);
      if (++s == SF_STAGES) { s = 0; ph ^= 1u; }
      if (++acc == 2) { acc = 0; accph ^= 1u; }
    }
  } else {
    // four epilogue warps: TMEM lane quarter = warp % 4, thread = voxel row
    const int q = warp & 3, m = q * 32 + lane;
    int acc = 0; uint32_t accph = 0;
    for (int tile = sf_tile0(p); tile < sf_tend(p); tile += sf_tstep(p)) {
      const long long v = (long long)tile * 128 + m;
      mbar_wait(tfull_bar(acc), accph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_cols;
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        uint32_t u[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
              "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
            : "r"(taddr + (uint32_t)c0));
        tmem_ld_wait();
        if (v < p.Vtot) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(u[j]);
          store8<bf16>(p.y + v * p.ldy + c0, f);
          store8<bf16>(p.y + v * p.ldy + c0 + 8, f + 8);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; accph ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

template <int CIN>
__global__ void __launch_bounds__(SfCfg<CIN>::WG_THREADS, 1) stem_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmdy, const SfParams p) {
  using C = SfCfg<CIN>;
  constexpr int KP = C::KP, NSUB = KP / 64, SF_STAGES = C::STAGES, PW = C::PROD / 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t a_bytes = NSUB * 16384u;
  const uint32_t row_bytes = (uint32_t)p.Cout * 2u;                    // dY row: 32 / 64 / 128 bytes
  const uint32_t b_bytes = 128u * row_bytes;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023u) & ~1023u);
  const uint32_t ring_base = smem_base;
  const uint32_t bar_base = ring_base + SF_STAGES * stage_bytes;
  const uint32_t full0 = bar_base, empty0 = bar_base + 8u * SF_STAGES;
  const uint32_t acc_bar = bar_base + 8u * (2 * SF_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * SF_STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  const uint32_t tmem_cols = p.Cout <= 32 ? 32u : 64u;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmdy);
    for (int s = 0; s < SF_STAGES; ++s) { mbar_init(full0 + 8u * s, C::GROUP_THREADS + 1); mbar_init(empty0 + 8u * s, 1); }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == PW) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < PW) {
    sf_producer<CIN>(p, ring_base, stage_bytes, full0, empty0);
  } else if (warp == PW) {
    // D[Kp (+ unused rows up to 128), Cout] += A^T dY: both operands MN-major, K = 128 voxel rows = 8 steps of 16
    const uint32_t issue = elect_one_sync() ? 1u : 0u;
    const uint32_t idesc = umma_idesc(128, p.Cout, 1, 1);
    const uint32_t b_layout = row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u, b_sbo = 8u * row_bytes;
    // LBO = distance between 64-element blocks along M: the second 64-tap sub-tile (Kp = 128); with Kp = 64 rows 64..127
    // of the accumulator alias rows 0..63 and are ignored
    const uint64_t adesc_hi = umma_desc(0, NSUB == 2 ? 16384u : 0u, 1024, 2);
    const uint64_t bdesc_hi = umma_desc(0, b_bytes, b_sbo, b_layout);
    const uint64_t a_adv = (uint64_t)((2u * 1024u) >> 4), b_adv = (uint64_t)((2u * b_sbo) >> 4);
    uint32_t s = 0, ph = 0;
    uint32_t accflag = 0;
    for (int tile = sf_tile0(p); tile < sf_tend(p); tile += sf_tstep(p)) {
      mbar_wait(full0 + 8u * s, ph);
      tc_fence_after();
      uint64_t ad = adesc_hi | (uint64_t)(((ring_base + s * stage_bytes) >> 4) & 0x3FFF);
      uint64_t bd = bdesc_hi | (uint64_t)(((ring_base + s * stage_bytes + a_bytes) >> 4) & 0x3FFF);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        umma_ss_p(tmem_base, ad, bd, idesc, accflag, issue);
        accflag = 1u;
        ad += a_adv; bd += b_adv;
      }
      umma_commit_p(empty0 + 8u * s, issue);
      if (++s == SF_STAGES) { s = 0; ph ^= 1u; }
    }
    umma_commit_p(acc_bar, issue);
  } else if (warp == PW + 1) {
    // TMA: dY rows [tile*128, +128) x Cout (rows past the end are zero-filled)
    const uint32_t issue = elect_one_sync() ? 1u : 0u;
    uint32_t s = 0, ph = 0;
    for (int tile = sf_tile0(p); tile < sf_tend(p); tile += sf_tstep(p)) {
      mbar_wait(empty0 + 8u * s, ph ^ 1u);
      mbar_expect_tx_p(full0 + 8u * s, b_bytes, issue);
      asm volatile(
          "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
          "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
          ::"r"(ring_base + s * stage_bytes + a_bytes), "l"(&tmdy), "r"(full0 + 8u * s), "r"(0), "r"(tile * 128), "r"(issue)
          : "memory");
      if (++s == SF_STAGES) { s = 0; ph ^= 1u; }
    }
  } else {
    // four epilogue warps: lane = tap row k
    const int q = warp & 3, m = q * 32 + lane;
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    float* dst = p.partial + ((long long)blockIdx.x * KP + m) * p.Cout;
    for (int c0 = 0; c0 < p.Cout; c0 += 16) {
      uint32_t u[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
            "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
          : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0));
      tmem_ld_wait();
      if (m < KP) {
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[c0 + j] = __uint_as_float(u[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn sf_get_encode() {
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

template <typename K>
int sf_set_smem(K kernel, size_t smem, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { hdf_set_error("%s: smem attribute: %s", what, cudaGetErrorString(e)); return HDF_ERR_CUDA; }
  return HDF_OK;
}

SfParams sf_params(const float* x, int N, int D, int H, int W, int Cout) {
  SfParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.N = N; p.D = D; p.H = H; p.W = W; p.Cout = Cout;
  p.V = (long long)D * H * W;
  p.Vtot = p.V * N;
  p.num_tiles = (int)((p.Vtot + 127) / 128);
  return p;
}

}  // namespace

// partial [S][Kp][Cout] -> dw [Cout][Cin][27]  (tc_conv.cu)
int hdf_stem_wgrad_reduce(const float* part, float* dw, int S, int Cin, int Cout, int Kp, int accumulate, void* stream);

extern "C" {

int hdf_stem_kp(int Cin);

// 1 if the first convolution runs without the im2col matrix (HDF_NO_STEM_FUSED=1 restores hdf_stem_im2col + hdf_stem_conv_*)
int hdf_stem_fused_supported(int Cin, int Cout) {
  const char* off = getenv("HDF_NO_STEM_FUSED");
  return !(off && off[0] == '1') && Cin >= 1 && Cin <= 4 && (Cout == 16 || Cout == 32 || Cout == 64);
}

// y [N*D*H*W, Cout] bf16 (row stride ldy) = conv3d(x, w), x NCDHW fp32, w_packed = hdf_stem_pack_weights output
int hdf_stem_fused_fwd(const float* x_ncdhw, const void* w_packed_bf16, void* y, long long ldy, int N, int Cin, int D, int H, int W,
                       int Cout, void* stream) {
  HDF_REQUIRE(hdf_stem_fused_supported(Cin, Cout), "hdf_stem_fused_fwd: unsupported Cin=%d Cout=%d", Cin, Cout);
  HDF_REQUIRE(x_ncdhw && w_packed_bf16 && y && (ldy % 8 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)w_packed_bf16 % 16 == 0),
              "hdf_stem_fused_fwd: bad args");
  SfParams p = sf_params(x_ncdhw, N, D, H, W, Cout);
  HDF_REQUIRE(p.Vtot < (1ll << 31) - 256, "hdf_stem_fused_fwd: volume too large");
  p.wp = (const bf16*)w_packed_bf16; p.y = (bf16*)y; p.ldy = ldy;
  const int Kp = hdf_stem_kp(Cin), nsub = Kp / 64;
  const int stages = Cin <= 2 ? SfCfg<1>::STAGES : SfCfg<4>::STAGES;
  const size_t smem = (((size_t)nsub * Cout * 128 + 1023) & ~(size_t)1023) + (size_t)stages * nsub * 16384 + 1024 +
                      8 * (2 * stages + 6) + 64;
  const int sms = hdf_sm_count_cached();
  // contiguous chunks of tiles per CTA, ~10 waves: the forward runs next to the token-branch and weight-packing kernels at
  // the start of the step, where a persistent one-CTA-per-SM grid waits for its slowest SM (0.42 ms vs 0.2 ms alone)
  static const int tpc_env = getenv("HDF_STEM_TILES_PER_CTA") ? atoi(getenv("HDF_STEM_TILES_PER_CTA")) : 32;
  p.tiles_per_cta = tpc_env > 0 && p.num_tiles > 4 * sms ? tpc_env : 0;
  const int grid = p.tiles_per_cta ? cdiv(p.num_tiles, p.tiles_per_cta) : (p.num_tiles < sms ? p.num_tiles : sms);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = HDF_OK;
  switch (Cin) {
    case 1: rc = sf_set_smem(stem_tc_fwd_kernel<1>, smem, "hdf_stem_fused_fwd"); if (!rc) stem_tc_fwd_kernel<1><<<grid, SfCfg<1>::FWD_THREADS, smem, s>>>(p); break;
    case 2: rc = sf_set_smem(stem_tc_fwd_kernel<2>, smem, "hdf_stem_fused_fwd"); if (!rc) stem_tc_fwd_kernel<2><<<grid, SfCfg<2>::FWD_THREADS, smem, s>>>(p); break;
    case 3: rc = sf_set_smem(stem_tc_fwd_kernel<3>, smem, "hdf_stem_fused_fwd"); if (!rc) stem_tc_fwd_kernel<3><<<grid, SfCfg<3>::FWD_THREADS, smem, s>>>(p); break;
    default: rc = sf_set_smem(stem_tc_fwd_kernel<4>, smem, "hdf_stem_fused_fwd"); if (!rc) stem_tc_fwd_kernel<4><<<grid, SfCfg<4>::FWD_THREADS, smem, s>>>(p); break;
  }
  if (rc) return rc;
  HDF_LAUNCH_CHECK("hdf_stem_fused_fwd");
  return HDF_OK;
}

size_t hdf_stem_fused_wgrad_workspace(int Cin, int Cout) { return (size_t)hdf_sm_count_cached() * hdf_stem_kp(Cin) * Cout * sizeof(float); }

// dw [Cout][Cin][27] fp32 (+)= sum over voxels of dy[v][co] * x[neighbour(v, tap)][ci];  dy [N*D*H*W, Cout] bf16 (row stride ldy)
int hdf_stem_fused_wgrad(const float* x_ncdhw, const void* dy, long long ldy, float* dw, int N, int Cin, int D, int H, int W, int Cout,
                         void* workspace, size_t ws_bytes, int accumulate, void* stream) {
  HDF_REQUIRE(hdf_stem_fused_supported(Cin, Cout), "hdf_stem_fused_wgrad: unsupported Cin=%d Cout=%d", Cin, Cout);
  HDF_REQUIRE(x_ncdhw && dy && dw && workspace && (ldy % 8 == 0) && ((uintptr_t)dy % 16 == 0), "hdf_stem_fused_wgrad: bad args");
  HDF_REQUIRE(ws_bytes >= hdf_stem_fused_wgrad_workspace(Cin, Cout), "hdf_stem_fused_wgrad: workspace too small");
  EncodeTiledFn enc = sf_get_encode();
  if (!enc) { hdf_set_error("hdf_stem_fused_wgrad: cuTensorMapEncodeTiled unavailable"); return HDF_ERR_CUDA; }
  SfParams p = sf_params(x_ncdhw, N, D, H, W, Cout);
  HDF_REQUIRE(p.Vtot < (1ll << 31) - 256, "hdf_stem_fused_wgrad: volume too large");
  p.partial = (float*)workspace;
  const int Kp = hdf_stem_kp(Cin), nsub = Kp / 64;
  CUtensorMap tmdy;
  {
    cuuint64_t gdim[2] = {(cuuint64_t)Cout, (cuuint64_t)p.Vtot};
    cuuint64_t gstr[1] = {(cuuint64_t)ldy * 2};
    cuuint32_t box[2] = {(cuuint32_t)Cout, 128};
    cuuint32_t estr[2] = {1, 1};
    const int rb = Cout * 2;
    CUresult r = enc(&tmdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(dy), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hdf_set_error("hdf_stem_fused_wgrad: encode(dy) failed: %d", (int)r); return HDF_ERR_CUDA; }
  }
  const size_t stage = (size_t)nsub * 16384 + (((size_t)128 * Cout * 2 + 1023) & ~(size_t)1023);
  const int stages = Cin <= 2 ? SfCfg<1>::STAGES : SfCfg<4>::STAGES;
  const size_t smem = (size_t)stages * stage + 1024 + 8 * (2 * stages + 3) + 64;
  const int sms = hdf_sm_count_cached();
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = HDF_OK;
  switch (Cin) {
    case 1: rc = sf_set_smem(stem_tc_wgrad_kernel<1>, smem, "hdf_stem_fused_wgrad"); if (!rc) stem_tc_wgrad_kernel<1><<<grid, SfCfg<1>::WG_THREADS, smem, s>>>(tmdy, p); break;
    case 2: rc = sf_set_smem(stem_tc_wgrad_kernel<2>, smem, "hdf_stem_fused_wgrad"); if (!rc) stem_tc_wgrad_kernel<2><<<grid, SfCfg<2>::WG_THREADS, smem, s>>>(tmdy, p); break;
    case 3: rc = sf_set_smem(stem_tc_wgrad_kernel<3>, smem, "hdf_stem_fused_wgrad"); if (!rc) stem_tc_wgrad_kernel<3><<<grid, SfCfg<3>::WG_THREADS, smem, s>>>(tmdy, p); break;
    default: rc = sf_set_smem(stem_tc_wgrad_kernel<4>, smem, "hdf_stem_fused_wgrad"); if (!rc) stem_tc_wgrad_kernel<4><<<grid, SfCfg<4>::WG_THREADS, smem, s>>>(tmdy, p); break;
  }
  if (rc) return rc;
  HDF_LAUNCH_CHECK("hdf_stem_fused_wgrad");
  return hdf_stem_wgrad_reduce((const float*)workspace, dw, grid, Cin, Cout, Kp, accumulate, stream);
}

}  // extern "C"
