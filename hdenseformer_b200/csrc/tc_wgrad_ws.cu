// Plane-ring weight gradient for the 3x3x3 convolutions that have a 32-channel operand (block_1_2_left/right 32x32,
// block_1_1_right 64->32, block_2_1_left 32->64, up3 64->32: the full- and half-resolution layers, 55 % of the weight-
// gradient time of the first-generation kernel).
//
//   dW[kd,kh,kw][ci][co] = sum_v X[v + (kd-1, kh-1, kw-1)][ci] * dY[v][co]          (autograd of models/HDenseFormer.py:151,167)
//
// tc_conv_wgrad_kernel (tc_conv.cu) loads one shifted [128 voxel x 32 ch] tile per tap per voxel chunk: 27 + 1 tile loads and
// 7 MMAs of N = 32 per 16 voxels; measured it is bound by shared-memory bandwidth (TMA writes + operand reads,
// profiles/r1_ncu_conv_wgrad_32x32_2cta.txt: DRAM read 1.8x algorithmic, L2 53 % busy).  Here the 27 relative shifts are
// split between the two operands so that one pair of resident tiles serves many taps:
//   * P = the 32-channel operand: three w-shifted boxes (box start w0-1, w0, w0+1) of the tile's own plane, stacked along
//     the MMA M dimension through the descriptor's leading-byte-offset (M = 128 = 3 x 32 channels + 32 unused rows);
//   * Q = the other operand, 32 channels per pass: one box of TH+2 lines per plane; the three kh shifts are three
//     overlapping sub-tiles at line-aligned offsets stacked along N (N = 96, LBO = one line), the three kd shifts are the
//     three planes z-1, z, z+1 of a ring the persistent CTA fills while it walks a column of tiles along D -- every Q plane
//     is loaded once and used by three tiles.
// Per 16 voxels: 3 MMAs (128 x 96 x 16, both operands MN-major, K = voxel rows) instead of 7, and ~4.7 tile loads per
// 128 voxels instead of 28.  Accumulators (3 x 96 TMEM columns) stay resident for the CTA's whole slab of tiles (split-K
// over CTAs); fp32 partials are reduced in a fixed order by wg3_reduce_kernel (deterministic).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

int hdf_sm_count_cached();

namespace {
using namespace tcptx;

constexpr int WG3_THREADS = 192;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: final epilogue

struct Wg3Params {
  int N, D, H, W;
  int Cq;                        // channels of the plane-ring operand (multiple of 32; blockIdx.y = 32-channel chunk)
  int p_is_x;                    // 1: P = x (Cin = 32), Q = dy ; 0: P = dy (Cout = 32), Q = x
  int TH, TW, K;                 // tile: TH lines x TW columns = K voxel rows (multiple of 16)
  int nTh, nTw, nSeg, seg_len, num_items;
  int sp, sq;                    // P stages / Q ring slots
  uint32_t p_box_bytes, p_stage_bytes, q_box_bytes, q_slot_bytes, line_bytes;
  float* partial;                // [gridDim.x][27][32][Cq]
};

__global__ void __launch_bounds__(WG3_THREADS, 1)
tc_wgrad_ws_kernel(const __grid_constant__ CUtensorMap tmp, const __grid_constant__ CUtensorMap tmq, const Wg3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int SP = p.sp, SQ = p.sq;
  const uint32_t p_base = smem_base;
  const uint32_t q_base = p_base + (uint32_t)SP * p.p_stage_bytes;
  const uint32_t bar_base = q_base + (uint32_t)SQ * p.q_slot_bytes;
  auto pfull = [&](int s) { return bar_base + 8u * s; };
  auto pempty = [&](int s) { return bar_base + 8u * (SP + s); };
  auto qfull = [&](int s) { return bar_base + 8u * (2 * SP + s); };
  auto qempty = [&](int s) { return bar_base + 8u * (2 * SP + SQ + s); };
  const uint32_t accfull = bar_base + 8u * (2 * SP + 2 * SQ);
  const uint32_t tmem_slot = accfull + 8u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmp);
    tma_prefetch_desc(&tmq);
    for (int s = 0; s < SP; ++s) { mbar_init(pfull(s), 1); mbar_init(pempty(s), 1); }
    for (int s = 0; s < SQ; ++s) { mbar_init(qfull(s), 1); mbar_init(qempty(s), 1); }
    mbar_init(accfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int chunk = blockIdx.y;
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
    w0 = tw * p.TW;
  };

  if (warp == 0) {
    // ===== TMA producer (all lanes walk the uniform schedule, the elected lane issues).  Order per item: Q planes d0-1, d0,
    // then for every tile its newest Q plane (d+1) followed by its three P boxes -- the order the consumer needs them in.
    const uint32_t issue = elect_one_sync() ? 1u : 0u;
    uint32_t ps = 0, pph = 0, qs = 0, qph = 0;
    auto load_q = [&](int z, int h0, int w0, int n) {
      mbar_wait(qempty(qs), qph ^ 1u);
      mbar_expect_tx_p(qfull(qs), p.q_box_bytes, issue);
      tma_load_5d_p(q_base + qs * p.q_slot_bytes, &tmq, qfull(qs), chunk * 32, w0, h0 - 1, z, n, issue);
      if (++qs == (uint32_t)SQ) { qs = 0; qph ^= 1u; }
    };
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int n, d0, d1, h0, w0;
      decode(item, n, d0, d1, h0, w0);
      load_q(d0 - 1, h0, w0, n);
      load_q(d0, h0, w0, n);
      for (int d = d0; d < d1; ++d) {
        load_q(d + 1, h0, w0, n);
        mbar_wait(pempty(ps), pph ^ 1u);
        mbar_expect_tx_p(pfull(ps), 3u * p.p_box_bytes, issue);
#pragma unroll
        for (int i = 0; i < 3; ++i)
          tma_load_5d_p(p_base + ps * p.p_stage_bytes + (uint32_t)i * p.p_box_bytes, &tmp, pfull(ps), 0, w0 + i - 1, h0, d, n, issue);
        if (++ps == (uint32_t)SP) { ps = 0; pph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    const uint32_t issue = elect_one_sync() ? 1u : 0u;
    const uint32_t idesc = umma_idesc(128, 96, 1, 1);                 // both operands MN-major (K = voxel rows)
    // SW64 MN-major: SBO = 8 rows x 64 B between 8-row groups along K; LBO = byte stride between the 32-channel sub-tiles
    // stacked along M (P: one box) / along N (Q: one line)
    const uint64_t adesc_hi = umma_desc(0, p.p_box_bytes, 512, 4);
    const uint64_t bdesc_hi = umma_desc(0, p.line_bytes, 512, 4);
    const int ksteps = p.K / 16;
    const bool flat = p.D == 1;
    uint32_t ps = 0, pph = 0;
    uint32_t s0 = 0, ws = 0, wph = 0;
    int ahead = 0;
    uint32_t accflag = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int n, d0, d1, h0, w0;
      decode(item, n, d0, d1, h0, w0);
      const int nd = d1 - d0;
      for (int t = 0; t < nd; ++t) {
        while (ahead < 3) {
          mbar_wait(qfull(ws), wph);
          if (++ws == (uint32_t)SQ) { ws = 0; wph ^= 1u; }
          ++ahead;
        }
        mbar_wait(pfull(ps), pph);
        tc_fence_after();
        const uint32_t s1 = s0 + 1 == (uint32_t)SQ ? 0u : s0 + 1;
        const uint32_t s2 = s1 + 1 == (uint32_t)SQ ? 0u : s1 + 1;
        uint64_t ad = adesc_hi | (uint64_t)(((p_base + ps * p.p_stage_bytes) >> 4) & 0x3FFF);
        uint64_t b0 = bdesc_hi | (uint64_t)(((q_base + s0 * p.q_slot_bytes) >> 4) & 0x3FFF);
        uint64_t b1 = bdesc_hi | (uint64_t)(((q_base + s1 * p.q_slot_bytes) >> 4) & 0x3FFF);
        uint64_t b2 = bdesc_hi | (uint64_t)(((q_base + s2 * p.q_slot_bytes) >> 4) & 0x3FFF);
        for (int k = 0; k < ksteps; ++k) {       // 16 voxel rows = 1024 B per step in every sub-tile
          // one-plane volumes (the 2-D model as flat volumes): planes d-1 / d+1 are zero padding, their taps are written as
          // zeros by the epilogue instead of being accumulated
          if (!flat) umma_ss_p(tmem_base + 0u, ad, b0, idesc, accflag, issue);
          umma_ss_p(tmem_base + 96u, ad, b1, idesc, accflag, issue);
          if (!flat) umma_ss_p(tmem_base + 192u, ad, b2, idesc, accflag, issue);
          accflag = 1;
          ad += 64; b0 += 64; b1 += 64; b2 += 64;
        }
        umma_commit_p(pempty(ps), issue);
        if (++ps == (uint32_t)SP) { ps = 0; pph ^= 1u; }
        umma_commit_p(qempty(s0), issue);
        if (t == nd - 1) {
          umma_commit_p(qempty(s1), issue);
          umma_commit_p(qempty(s2), issue);
          s0 = s2 + 1 == (uint32_t)SQ ? 0u : s2 + 1;
          ahead -= 3;
        } else {
          s0 = s1;
          ahead -= 1;
        }
      }
    }
    umma_commit_p(accfull, issue);
  } else {
    // ===== epilogue (once): TMEM lane m = 32 i + pc (i = w-shift slot of P, pc = P channel), column = 96 sd + 32 j + c
    // (sd = plane slot, j = line-shift slot of Q, c = Q channel of this pass)
    const int i = warp & 3;
    mbar_wait(accfull, 0);
    tc_fence_after();
    if (i < 3 && blockIdx.x < (unsigned)p.num_items) {
      const int pc = lane;
      const int kw = p.p_is_x ? i : 2 - i;
      float* base = p.partial + (size_t)blockIdx.x * 27 * 32 * p.Cq;
#pragma unroll 1
      for (int sd = 0; sd < 3; ++sd) {
        const int kd = p.p_is_x ? 2 - sd : sd;
#pragma unroll 1
        for (int j = 0; j < 3; ++j) {
          const int kh = p.p_is_x ? 2 - j : j;
          const int tap = kd * 9 + kh * 3 + kw;
          float* dst = base + ((size_t)tap * 32 + pc) * p.Cq + chunk * 32;
          const uint32_t taddr = tmem_base + ((uint32_t)(i * 32) << 16) + (uint32_t)(sd * 96 + j * 32);
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 8) {
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(taddr + (uint32_t)c0));
            tmem_ld_wait();
            if (p.D == 1 && sd != 1) {
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = 0u;
            }
            *reinterpret_cast<float4*>(dst + c0) = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
            *reinterpret_cast<float4*>(dst + c0 + 4) = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// partial[S][27][32][Cq] -> g[ci*sci + co*sco + tap] (+= if accumulate); fixed summation order
__global__ void wg3_reduce_kernel(const float* __restrict__ part, float* __restrict__ g, int S, int Cq, int p_is_x, long long sci,
                                  long long sco, int accumulate) {
  const long long per = 27ll * 32 * Cq;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < per; idx += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < S; ++z) s += part[z * per + idx];
    const int qc = idx % Cq;
    const int pc = (idx / Cq) % 32;
    const int tap = idx / (32ll * Cq);
    const int ci = p_is_x ? pc : qc, co = p_is_x ? qc : pc;
    float* q = g + ci * sci + co * sco + tap;
    *q = accumulate ? (*q + s) : s;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn wg3_get_encode() {
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

// geometry + schedule shared by the workspace query and the launch
void wg3_plan(int N, int D, int H, int W, int Cin, int Cout, Wg3Params& p, int& grid_x, int& passes) {
  memset(&p, 0, sizeof(p));
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.p_is_x = Cin == 32 ? 1 : 0;
  p.Cq = p.p_is_x ? Cout : Cin;
  passes = p.Cq / 32;
  // tile: TW columns (multiple of 8: sub-tiles at line offsets stay aligned to the 8-row swizzle atom), TH lines,
  // K = TH*TW voxel rows per tile (multiple of 16, <= 160); cost ~ rows processed + per-tile overhead + halo lines of Q
  double best = -1;
  p.TH = 1; p.TW = 16;
  for (int tw = 16; tw <= 128; tw += 8)
    for (int th = 1; th * tw <= 160; ++th) {
      if ((th * tw) % 16) continue;
      const double tiles = (double)cdiv(H, th) * cdiv(W, tw);
      const double cost = tiles * (th * tw + 32.0 + 0.15 * (th + 2) * tw);
      if (best < 0 || cost < best) { best = cost; p.TH = th; p.TW = tw; }
    }
  p.K = p.TH * p.TW;
  p.nTh = cdiv(H, p.TH); p.nTw = cdiv(W, p.TW);
  p.line_bytes = (uint32_t)p.TW * 64u;
  p.p_box_bytes = (uint32_t)p.K * 64u;                       // multiple of 1024 (K % 16 == 0)
  p.p_stage_bytes = 3u * p.p_box_bytes;
  p.q_box_bytes = (uint32_t)(p.TH + 2) * p.line_bytes;
  p.q_slot_bytes = (p.q_box_bytes + 1023u) & ~1023u;
  static const int smem_kb = getenv("HDF_TC_SMEM_KB") ? atoi(getenv("HDF_TC_SMEM_KB")) : 192;
  // ring depths: P stages are used once, Q planes by three consecutive tiles (3 in use + prefetch)
  p.sq = 6;
  p.sp = 3;
  while (p.sp > 2 && (size_t)p.sp * p.p_stage_bytes + (size_t)p.sq * p.q_slot_bytes + 2048 > (size_t)smem_kb * 1024) --p.sp;
  while (p.sq > 4 && (size_t)p.sp * p.p_stage_bytes + (size_t)p.sq * p.q_slot_bytes + 2048 > (size_t)smem_kb * 1024) --p.sq;
  const int sms = hdf_sm_count_cached();
  grid_x = sms / passes > 0 ? sms / passes : 1;
  const long long cols = (long long)N * p.nTh * p.nTw;
  {
    double best_eff = -1; int best_len = D;
    for (int k = 1; k <= D; ++k) {
      const int len = cdiv(D, k), nseg = cdiv(D, len);
      const long long items = cols * nseg;
      const long long rounds = (items + grid_x - 1) / grid_x;
      const double eff = (double)cols * D / ((double)rounds * grid_x * len) * (len / (len + 0.7));   // 2 extra Q planes per segment
      if (eff > best_eff + 1e-9) { best_eff = eff; best_len = len; }
      if (eff >= 0.96) { best_len = len; break; }
    }
    p.seg_len = best_len;
    p.nSeg = cdiv(D, p.seg_len);
  }
  p.num_items = (int)(cols * p.nSeg);
  if (grid_x > p.num_items) grid_x = p.num_items;
}

}  // namespace

extern "C" {

int hdf_tc_wgrad_ws_supported(int mode, int Cin, int Cout) {
  const char* off = getenv("HDF_TC_NO_WGRAD_WS");
  if (off && off[0] == '1') return 0;
  if (mode != 0) return 0;
  if (Cin == 32) return Cout % 32 == 0 && Cout >= 32 && Cout <= 256;
  if (Cout == 32) return Cin % 32 == 0 && Cin >= 32 && Cin <= 256;
  return 0;
}

size_t hdf_tc_wgrad_ws_workspace(int N, int D, int H, int W, int Cin, int Cout) {
  if (!hdf_tc_wgrad_ws_supported(0, Cin, Cout)) return 0;
  Wg3Params p; int gx, passes;
  wg3_plan(N, D, H, W, Cin, Cout, p, gx, passes);
  return (size_t)gx * 27 * 32 * p.Cq * sizeof(float);
}

// dw[ci*stride_ci + co*stride_co + tap] (+)= sum_v x[v + off(tap)][ci] * dy[v][co]     (mode 0, Cin == 32 or Cout == 32)
int hdf_tc_wgrad_ws(const void* x, long long ldx, const void* dy, long long ldy, float* dw, long long stride_ci, long long stride_co,
                    int N, int D, int H, int W, int Cin, int Cout, void* workspace, size_t ws_bytes, int accumulate, void* stream) {
  HDF_REQUIRE(hdf_tc_wgrad_ws_supported(0, Cin, Cout), "hdf_tc_wgrad_ws: unsupported Cin=%d Cout=%d", Cin, Cout);
  HDF_REQUIRE(x && dy && dw && workspace, "hdf_tc_wgrad_ws: null pointer");
  HDF_REQUIRE((ldx % 8 == 0) && (ldy % 8 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)dy % 16 == 0),
              "hdf_tc_wgrad_ws: operands must be 16-byte aligned with channel strides multiple of 8");
  EncodeTiledFn enc = wg3_get_encode();
  if (!enc) { hdf_set_error("hdf_tc_wgrad_ws: cuTensorMapEncodeTiled unavailable"); return HDF_ERR_CUDA; }
  Wg3Params p; int gx, passes;
  wg3_plan(N, D, H, W, Cin, Cout, p, gx, passes);
  HDF_REQUIRE(ws_bytes >= (size_t)gx * 27 * 32 * p.Cq * sizeof(float), "hdf_tc_wgrad_ws: workspace too small");
  p.partial = (float*)workspace;
  const void* pp = p.p_is_x ? x : dy;
  const void* qp = p.p_is_x ? dy : x;
  const long long pld = p.p_is_x ? ldx : ldy, qld = p.p_is_x ? ldy : ldx;
  CUtensorMap tmp, tmq;
  for (int which = 0; which < 2; ++which) {
    const void* base = which == 0 ? pp : qp;
    const long long ld = which == 0 ? pld : qld;
    const int C = which == 0 ? 32 : p.Cq;
    cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t gstr[4] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2, (cuuint64_t)D * H * W * ld * 2};
    cuuint32_t box[5] = {32, (cuuint32_t)p.TW, (cuuint32_t)(which == 0 ? p.TH : p.TH + 2), 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(which == 0 ? &tmp : &tmq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hdf_set_error("hdf_tc_wgrad_ws: encode failed: %d", (int)r); return HDF_ERR_CUDA; }
  }
  const size_t smem = (size_t)p.sp * p.p_stage_bytes + (size_t)p.sq * p.q_slot_bytes + 1024 + 8 * (2 * p.sp + 2 * p.sq + 4) + 64;
  HDF_REQUIRE(p.sp >= 2 && p.sq >= 4 && smem <= 227 * 1024, "hdf_tc_wgrad_ws: smem plan does not fit (%zu bytes)", smem);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e == cudaSuccess && getenv("HDF_NO_MAX_CARVEOUT") == nullptr)
      e = cudaFuncSetAttribute(tc_wgrad_ws_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) { hdf_set_error("hdf_tc_wgrad_ws: smem attribute: %s", cudaGetErrorString(e)); return HDF_ERR_CUDA; }
    configured = true;
  }
  dim3 grid(gx, passes);
  tc_wgrad_ws_kernel<<<grid, WG3_THREADS, smem, (cudaStream_t)stream>>>(tmp, tmq, p);
  HDF_LAUNCH_CHECK("hdf_tc_wgrad_ws");
  const long long per = 27ll * 32 * p.Cq;
  wg3_reduce_kernel<<<min(2048, cdiv(per, 256)), 256, 0, (cudaStream_t)stream>>>((const float*)workspace, dw, gx, p.Cq, p.p_is_x,
                                                                                 stride_ci, stride_co, accumulate);
  HDF_LAUNCH_CHECK("hdf_tc_wgrad_ws/reduce");
  return HDF_OK;
}

}  // extern "C"
