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

#include "common.cuh"

int hdf_sm_count_cached();

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
struct TcConvParams {
  int N, D, H, W, Cin, Cout;
  int TD, TH, TW;
  int nTd, nTh, nTw;
  int num_tiles;
  int KC, kchunks, stages;
  uint32_t a_bytes, b_bytes, stage_bytes;  // stage_bytes = round1024(a) + round1024(b)
  uint32_t layout, sbo;                    // UMMA layout type / stride-byte-offset of the swizzle mode
  uint32_t tmem_cols;
  long long ldy;
  const float* bias;
  bf16* y;
};

constexpr int TC_THREADS = 256;

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_fwd_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmw, const TcConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t a_region = (p.a_bytes + 1023u) & ~1023u;
  // barrier block lives after the stage ring
  const uint32_t bar_base = smem_base + p.stages * p.stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmw);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int kiters = 27 * p.kchunks;
  const int tiles_per_n = p.nTd * p.nTh * p.nTw;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_n;
        int r = tile - n * tiles_per_n;
        const int tw = r % p.nTw; r /= p.nTw;
        const int th = r % p.nTh;
        const int td = r / p.nTh;
        const int d0 = td * p.TD, h0 = th * p.TH, w0 = tw * p.TW;
        for (int it = 0; it < kiters; ++it) {
          const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
          const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), p.a_bytes + p.b_bytes);
          const uint32_t a_dst = smem_base + s * p.stage_bytes;
          tma_load_5d(a_dst, &tmx, full_bar(s), kc * p.KC, w0 + kw - 1, h0 + kh - 1, d0 + kd - 1, n);
          tma_load_2d(a_dst + a_region, &tmw, full_bar(s), kc * p.KC, tap * p.Cout);
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = umma_idesc(128, p.Cout, 0, 0);
      int s = 0; uint32_t ph = 0;
      int acc = 0; uint32_t accph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), accph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.Cout);
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_addr = smem_base + s * p.stage_bytes;
          const uint64_t adesc = umma_desc(a_addr, 16, p.sbo, p.layout);
          const uint64_t bdesc = umma_desc(a_addr + a_region, 16, p.sbo, p.layout);
          const int ksteps = p.KC / 16;
          for (int k = 0; k < ksteps; ++k)  // +32 B per K=16 step inside the swizzled row (encoded >>4)
            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) != 0);
          umma_commit(empty_bar(s));
          if (it == kiters - 1) umma_commit(tfull_bar(acc));
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        if (++acc == 2) { acc = 0; accph ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> (+bias) -> bf16 -> global =====
    const int q = warp - 4;  // TMEM lane quadrant of this warp
    const int m = q * 32 + lane;
    const int mw = m % p.TW, mh = (m / p.TW) % p.TH, md = m / (p.TW * p.TH);
    int acc = 0; uint32_t accph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n = tile / tiles_per_n;
      int r = tile - n * tiles_per_n;
      const int tw = r % p.nTw; r /= p.nTw;
      const int th = r % p.nTh;
      const int td = r / p.nTh;
      const int d = td * p.TD + md, h = th * p.TH + mh, w = tw * p.TW + mw;
      const bool valid = d < p.D && h < p.H && w < p.W;
      bf16* yrow = p.y + ((((long long)n * p.D + d) * p.H + h) * p.W + w) * p.ldy;
      mbar_wait(tfull_bar(acc), accph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Cout);
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; accph ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
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

__global__ void tc_pack_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cin, int Cout, long long sci,
                               long long sco, int flip) {
  const long long total = 27ll * Cin * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = i % Cin;
    const int co = (i / Cin) % Cout;
    const int tap = i / ((long long)Cin * Cout);
    out[i] = __float2bfloat16_rn(w[ci * sci + co * sco + (flip ? 26 - tap : tap)]);
  }
}

}  // namespace

extern "C" {

int hdf_tc_supported(int mode, int Cin, int Cout) {
  if (mode != 0) return 0;
  if (Cin % 16 != 0 || Cin < 16) return 0;
  if (Cout % 16 != 0 || Cout < 16 || Cout > 256) return 0;
  return 1;
}

size_t hdf_tc_pack_bytes(int Cin, int Cout) { return (size_t)27 * Cin * Cout * sizeof(bf16); }

// packed[tap][co][ci] (bf16) = w[ci*stride_ci + co*stride_co + (flip ? 26-tap : tap)]
int hdf_tc_pack_weights(const float* w, void* packed_bf16, int Cin, int Cout, long long stride_ci, long long stride_co,
                        int flip, void* stream) {
  HDF_REQUIRE(w && packed_bf16, "hdf_tc_pack_weights: null pointer");
  const long long total = 27ll * Cin * Cout;
  tc_pack_kernel<<<min(1024, cdiv(total, 256)), 256, 0, (cudaStream_t)stream>>>(w, (bf16*)packed_bf16, Cin, Cout, stride_ci,
                                                                              stride_co, flip);
  HDF_LAUNCH_CHECK("hdf_tc_pack_weights");
  return HDF_OK;
}

int hdf_tc_conv3d_fwd(const void* x, long long ldx, const void* w_packed_bf16, const float* bias, void* y, long long ldy,
                      int N, int D, int H, int W, int Cin, int Cout, double* stats_partial, void* stream) {
  (void)stats_partial;
  if (!hdf_tc_supported(0, Cin, Cout)) {
    hdf_set_error("hdf_tc_conv3d_fwd: unsupported channels Cin=%d Cout=%d", Cin, Cout);
    return HDF_ERR_UNSUPPORTED;
  }
  HDF_REQUIRE(x && w_packed_bf16 && y, "hdf_tc_conv3d_fwd: null pointer");
  HDF_REQUIRE((ldx % 8 == 0) && (ldy % 8 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0),
              "hdf_tc_conv3d_fwd: operands must be 16-byte aligned with channel strides multiple of 8");
  EncodeTiledFn enc = get_encode();
  if (!enc) { hdf_set_error("hdf_tc_conv3d_fwd: cuTensorMapEncodeTiled unavailable"); return HDF_ERR_CUDA; }

  TcConvParams p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  pick_tile(D, H, W, p.TD, p.TH, p.TW);
  p.nTd = cdiv(D, p.TD); p.nTh = cdiv(H, p.TH); p.nTw = cdiv(W, p.TW);
  p.num_tiles = N * p.nTd * p.nTh * p.nTw;
  p.KC = (Cin % 64 == 0) ? 64 : (Cin % 32 == 0) ? 32 : 16;
  p.kchunks = Cin / p.KC;
  p.a_bytes = 128u * p.KC * 2u;
  p.b_bytes = (uint32_t)Cout * p.KC * 2u;
  p.stage_bytes = ((p.a_bytes + 1023u) & ~1023u) + ((p.b_bytes + 1023u) & ~1023u);
  const int inner = p.KC * 2;
  p.layout = inner == 128 ? 2u : inner == 64 ? 4u : 6u;
  p.sbo = 8u * inner;
  p.stages = (int)(196608u / p.stage_bytes);
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) { hdf_set_error("hdf_tc_conv3d_fwd: stage too large"); return HDF_ERR_UNSUPPORTED; }
  uint32_t cols = 32;
  while (cols < 2u * Cout) cols *= 2;
  p.tmem_cols = cols;
  p.ldy = ldy; p.bias = bias; p.y = (bf16*)y;

  CUtensorMap tmx, tmw;
  {
    cuuint64_t gdim[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t gstr[4] = {(cuuint64_t)ldx * 2, (cuuint64_t)W * ldx * 2, (cuuint64_t)H * W * ldx * 2,
                          (cuuint64_t)D * H * W * ldx * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.KC, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TD, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(inner), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hdf_set_error("hdf_tc_conv3d_fwd: encode(x) failed: %d", (int)r); return HDF_ERR_CUDA; }
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)Cin, (cuuint64_t)27 * Cout};
    cuuint64_t gstr[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.KC, (cuuint32_t)Cout};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed_bf16), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(inner), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hdf_set_error("hdf_tc_conv3d_fwd: encode(w) failed: %d", (int)r); return HDF_ERR_CUDA; }
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024 /*align slack*/ + 8 * (2 * p.stages + 4) + 16;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e != cudaSuccess) { hdf_set_error("hdf_tc_conv3d_fwd: smem attribute: %s", cudaGetErrorString(e)); return HDF_ERR_CUDA; }
    configured = 227 * 1024;
  }
  const int grid = p.num_tiles < hdf_sm_count_cached() ? p.num_tiles : hdf_sm_count_cached();
  tc_conv_fwd_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(tmx, tmw, p);
  HDF_LAUNCH_CHECK("hdf_tc_conv3d_fwd");
  return HDF_OK;
}

size_t hdf_tc_wgrad_workspace(int, int, int, int, int, int) { return 0; }
int hdf_tc_conv3d_wgrad(const void*, long long, const void*, long long, float*, long long, long long, int, int, int, int, int,
                        int, void*, size_t, int, void*) {
  hdf_set_error("hdf_tc_conv3d_wgrad: not built yet");
  return HDF_ERR_UNSUPPORTED;
}

}  // extern "C"
