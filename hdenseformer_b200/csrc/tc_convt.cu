// Shift-major tcgen05 kernel for the big transposed convolution, upconv_1 = ConvTranspose3d(64 -> 32, k3, s2, p1, op1)
// (models/HDenseFormer.py:215, called at :249) whose output is the 2 x 144^3 grid.
//
//   out[2j + p] = b + sum over taps (k, s) of parity class p :  W[:, :, k]^T x[j + s]
//   per dimension:  p = 0 -> (k = 1, s = 0);   p = 1 -> (k = 2, s = 0), (k = 0, s = 1)
//
// The generic kernel (tc_conv.cu, mode 1) walks the 27 taps and loads one 128-voxel input box per tap: every input
// voxel crosses L2 -> shared memory 27 times (2.5 GB for this layer, 0.52 ms, 160 TF/s: profiles/r2_microbench_conv_v3).
// But only 8 distinct shifts s in {0,1}^3 exist; the box of shift s serves every class p >= s at once.  So here
//   * a tile is 4 x 4 x 8 input voxels (M = 128 rows of 64 channels = one 128-byte swizzled row each);
//   * 8 TMA boxes per tile (one per shift), 3.4x fewer bytes than tap-major;
//   * the accumulator holds all 8 parity classes side by side, 8 x 32 = 256 TMEM columns, double buffered (512);
//   * classes sit in Gray-code order (000 001 011 010 110 111 101 100) so the classes served by one shift form at most
//     two runs of adjacent columns: 10 MMAs of N = 256/128/64/32 per K step instead of 27 of N = 32;
//   * the 27 weight tiles [32 co x 64 ci] stay resident in shared memory (108 KB), stored run by run so that a run is
//     one K-major B operand;
//   * epilogue: a thread owns one input voxel and writes its 8 output voxels as 64-byte rows (+ bias).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

int hdf_sm_count_cached();

namespace {
using namespace tcptx;

constexpr int CT_THREADS = 320;      // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue
constexpr int CT_CIN = 64, CT_COUT = 32;
constexpr int CT_TD = 4, CT_TH = 4, CT_TW = 8;
constexpr int CT_STAGES = 6;
constexpr uint32_t CT_BOX_BYTES = 128u * CT_CIN * 2u;          // 16 KB
constexpr uint32_t CT_WBLK_BYTES = CT_COUT * CT_CIN * 2u;      // 4 KB per (shift, class) weight tile
constexpr uint32_t CT_W_BYTES = 27u * CT_WBLK_BYTES;

// accumulator position -> parity class (bit 2 = d, bit 1 = h, bit 0 = w)
__host__ __device__ constexpr int ct_gray(int pos) { return pos ^ (pos >> 1); }
// runs of adjacent accumulator positions served by one shift; weight tiles are stored in this order
struct CtRun { int shift, pos, len, blk; };
__host__ __device__ constexpr CtRun ct_run(int r) {
  constexpr CtRun t[10] = {{0, 0, 8, 0},  {1, 1, 2, 8},  {1, 5, 2, 10}, {2, 2, 4, 12}, {3, 2, 1, 16},
                           {3, 5, 1, 17}, {4, 4, 4, 18}, {5, 5, 2, 22}, {6, 4, 2, 24}, {7, 5, 1, 26}};
  return t[r];
}

struct CtParams {
  int N, D, H, W;                 // input grid
  int nTd, nTh, nTw, num_tiles;
  const uint4* wimg;              // packed weights: the 108 KB shared-memory image (hdf_tc_convt_pack_weights)
  const float* bias;
  bf16* y;                        // [N, 2D, 2H, 2W, 32] bf16, channel stride ldy
  long long ldy;
};

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

__global__ void __launch_bounds__(CT_THREADS, 1) tc_convt_kernel(const __grid_constant__ CUtensorMap tmx, const CtParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t w_base = smem_base;
  const uint32_t ring_base = w_base + CT_W_BYTES;
  const uint32_t bar_base = ring_base + CT_STAGES * CT_BOX_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (CT_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * CT_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * CT_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * CT_STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  __shared__ float bias_sh[CT_COUT];

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmx);
    for (int s = 0; s < CT_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  // resident weights: the packed buffer already is the swizzled shared-memory image
  {
    uint4* dst = reinterpret_cast<uint4*>(smem_gen + (w_base - smem_base));
    for (int i = threadIdx.x; i < (int)(CT_W_BYTES / 16); i += CT_THREADS) dst[i] = p.wimg[i];
    if (threadIdx.x < CT_COUT) bias_sh[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
    fence_proxy_async();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int per_n = p.nTd * p.nTh * p.nTw;
  auto decode = [&](int tile, int& n, int& d0, int& h0, int& w0) {
    n = tile / per_n;
    int r = tile - n * per_n;
    const int td = r / (p.nTh * p.nTw);
    r -= td * (p.nTh * p.nTw);
    const int th = r / p.nTw, tw = r - th * p.nTw;
    d0 = td * CT_TD; h0 = th * CT_TH; w0 = tw * CT_TW;
  };

  if (warp == 0) {
    // ===== TMA producer: the 8 shifted boxes of every tile (rows past the volume are zero-filled = no contribution)
    const uint32_t issue = elect_one_sync() ? 1u : 0u;
    uint32_t s = 0, ph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int n, d0, h0, w0;
      decode(tile, n, d0, h0, w0);
#pragma unroll
      for (int sh = 0; sh < 8; ++sh) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx_p(full_bar(s), CT_BOX_BYTES, issue);
        tma_load_5d_p(ring_base + s * CT_BOX_BYTES, &tmx, full_bar(s), 0, w0 + (sh & 1), h0 + ((sh >> 1) & 1), d0 + (sh >> 2), n, issue);
        if (++s == (uint32_t)CT_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    const uint32_t issue = elect_one_sync() ? 1u : 0u;
    const uint64_t desc_hi = umma_desc(0, 16, 1024, 2);      // K-major, 128-byte swizzle, 8-row atoms 1 KB apart
    uint32_t s = 0, ph = 0;
    int acc = 0; uint32_t accph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(acc), accph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
#pragma unroll
      for (int sh = 0; sh < 8; ++sh) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint64_t adesc = desc_hi | (uint64_t)(((ring_base + s * CT_BOX_BYTES) >> 4) & 0x3FFF);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
          const CtRun run = ct_run(r);
          if (run.shift != sh) continue;
          const uint32_t idesc = umma_idesc(128, 32 * run.len, 0, 0);
          const uint64_t bdesc = desc_hi | (uint64_t)(((w_base + (uint32_t)run.blk * CT_WBLK_BYTES) >> 4) & 0x3FFF);
#pragma unroll
          for (int ks = 0; ks < CT_CIN / 16; ++ks)       // +32 B per K = 16 step inside the swizzled row (encoded >> 4)
            umma_ss_p(d_tmem + (uint32_t)(run.pos * 32), adesc + (uint64_t)(2 * ks), bdesc + (uint64_t)(2 * ks), idesc,
                      (sh > 0 || ks > 0) ? 1u : 0u, issue);
        }
        umma_commit_p(empty_bar(s), issue);
        if (++s == (uint32_t)CT_STAGES) { s = 0; ph ^= 1u; }
      }
      umma_commit_p(tfull_bar(acc), issue);
      if (++acc == 2) { acc = 0; accph ^= 1u; }
    }
  } else {
    // ===== epilogue warps 2..9: TMEM lane quarter q = warp % 4 (tile rows 32q..32q+31), two warps per quarter, each
    // takes 4 of the 8 class positions
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int m = q * 32 + lane;                     // tile row = (td * 4 + th) * 8 + tw
    const int td = m >> 5, th = (m >> 3) & 3, tw = m & 7;
    const long long Ho = 2ll * p.H, Wo = 2ll * p.W, Do = 2ll * p.D;
    int acc = 0; uint32_t accph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int n, d0, h0, w0;
      decode(tile, n, d0, h0, w0);
      const int jd = d0 + td, jh = h0 + th, jw = w0 + tw;
      const bool valid = jd < p.D && jh < p.H && jw < p.W;
      mbar_wait(tfull_bar(acc), accph);
      tc_fence_after();
      const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u + (uint32_t)(half * 128);
      uint32_t v[2][32];
      tmem_ld_32x32b_x32(tcol, v[0]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        tmem_ld_wait();
        if (i + 1 < 4) tmem_ld_32x32b_x32(tcol + (uint32_t)((i + 1) * 32), v[(i + 1) & 1]);
        const int cls = ct_gray(half * 4 + i);
        if (valid) {
          bf16* row = p.y + ((((long long)n * Do + (2 * jd + (cls >> 2))) * Ho + (2 * jh + ((cls >> 1) & 1))) * Wo +
                             (2 * jw + (cls & 1))) * p.ldy;
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 8) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[i & 1][c0 + j]) + bias_sh[c0 + j];
            store8<bf16>(row + c0, f);
          }
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
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// w [64 ci][32 co][27 taps] fp32 (torch ConvTranspose3d layout) -> the kernel's shared-memory image: 27 tiles of
// [32 co rows x 64 ci] bf16, K-major with the 128-byte swizzle (16-byte chunk index XOR row % 8), ordered run by run
__global__ void tc_convt_pack_kernel(const float* __restrict__ w, bf16* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 27 * CT_COUT * CT_CIN) return;
  const int ci = idx % CT_CIN, co = (idx / CT_CIN) % CT_COUT, blk = idx / (CT_CIN * CT_COUT);
  int shift = 0, pos = 0;
  for (int r = 0; r < 10; ++r) {
    const CtRun run = ct_run(r);
    if (blk >= run.blk && blk < run.blk + run.len) { shift = run.shift; pos = run.pos + (blk - run.blk); }
  }
  const int cls = ct_gray(pos);
  int k[3];
  for (int a = 0; a < 3; ++a) {                 // a = 0: w, 1: h, 2: d
    const int pb = (cls >> a) & 1, sb = (shift >> a) & 1;
    k[a] = pb == 0 ? 1 : (sb == 0 ? 2 : 0);
  }
  const int tap = (k[2] * 3 + k[1]) * 3 + k[0];
  const float v = w[((long long)ci * CT_COUT + co) * 27 + tap];
  const int chunk = (ci >> 3) ^ (co & 7);
  out[(size_t)blk * (CT_COUT * CT_CIN) + (size_t)(co >> 3) * 512 + (co & 7) * 64 + chunk * 8 + (ci & 7)] = __float2bfloat16_rn(v);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn ct_get_encode() {
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

// 1 if the shift-major kernel takes this transposed convolution (k3, s2, p1, op1)
int hdf_tc_convt_supported(int Cin, int Cout) {
  const char* off = getenv("HDF_TC_NO_CONVT");     // read per call: tests A/B the two kernels in one process
  return !(off && off[0] == '1') && Cin == CT_CIN && Cout == CT_COUT;
}

size_t hdf_tc_convt_packed_bytes(void) { return CT_W_BYTES; }

// w: torch ConvTranspose3d weight [64][32][3][3][3] fp32 -> packed_bf16 (hdf_tc_convt_packed_bytes() bytes)
int hdf_tc_convt_pack_weights(const float* w, void* packed_bf16, void* stream) {
  HDF_REQUIRE(w && packed_bf16, "hdf_tc_convt_pack_weights: null pointer");
  const int total = 27 * CT_COUT * CT_CIN;
  tc_convt_pack_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(w, (bf16*)packed_bf16);
  HDF_LAUNCH_CHECK("hdf_tc_convt_pack_weights");
  return HDF_OK;
}

// x [N, D, H, W, 64] bf16 (channel stride ldx) -> y [N, 2D, 2H, 2W, 32] bf16 (channel stride ldy), optional fp32 bias[32]
int hdf_tc_convt_fwd(const void* x, long long ldx, const void* w_packed_bf16, const float* bias, void* y, long long ldy, int N,
                     int D, int H, int W, void* stream) {
  HDF_REQUIRE(x && w_packed_bf16 && y, "hdf_tc_convt_fwd: null pointer");
  HDF_REQUIRE((ldx % 8 == 0) && (ldy % 8 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) &&
                  ((uintptr_t)w_packed_bf16 % 16 == 0),
              "hdf_tc_convt_fwd: operands must be 16-byte aligned with channel strides multiple of 8");
  HDF_REQUIRE(N >= 1 && D >= 1 && H >= 1 && W >= 1, "hdf_tc_convt_fwd: bad shape");
  EncodeTiledFn enc = ct_get_encode();
  if (!enc) { hdf_set_error("hdf_tc_convt_fwd: cuTensorMapEncodeTiled unavailable"); return HDF_ERR_CUDA; }
  CtParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.nTd = cdiv(D, CT_TD); p.nTh = cdiv(H, CT_TH); p.nTw = cdiv(W, CT_TW);
  const long long tiles = (long long)N * p.nTd * p.nTh * p.nTw;
  HDF_REQUIRE(tiles < (1ll << 31), "hdf_tc_convt_fwd: too many tiles");
  p.num_tiles = (int)tiles;
  p.wimg = (const uint4*)w_packed_bf16; p.bias = bias; p.y = (bf16*)y; p.ldy = ldy;
  CUtensorMap tmx;
  {
    cuuint64_t gdim[5] = {(cuuint64_t)CT_CIN, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t gstr[4] = {(cuuint64_t)ldx * 2, (cuuint64_t)W * ldx * 2, (cuuint64_t)H * W * ldx * 2, (cuuint64_t)D * H * W * ldx * 2};
    cuuint32_t box[5] = {(cuuint32_t)CT_CIN, CT_TW, CT_TH, CT_TD, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hdf_set_error("hdf_tc_convt_fwd: encode(x) failed: %d", (int)r); return HDF_ERR_CUDA; }
  }
  const size_t smem = (size_t)CT_W_BYTES + (size_t)CT_STAGES * CT_BOX_BYTES + 1024 + 8 * (2 * CT_STAGES + 6) + 64;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_convt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess && getenv("HDF_NO_MAX_CARVEOUT") == nullptr)
      e = cudaFuncSetAttribute(tc_convt_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) { hdf_set_error("hdf_tc_convt_fwd: smem attribute: %s", cudaGetErrorString(e)); return HDF_ERR_CUDA; }
    configured = true;
  }
  const int sms = hdf_sm_count_cached();
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  tc_convt_kernel<<<grid, CT_THREADS, smem, (cudaStream_t)stream>>>(tmx, p);
  HDF_LAUNCH_CHECK("hdf_tc_convt_fwd");
  return HDF_OK;
}

}  // extern "C"
