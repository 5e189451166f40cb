// Patch embedding on the tensor cores (bf16 path): Conv3d(1 -> E, kernel 16, stride 16) of one modality
// (models/HDenseFormer.py:115-118,133) = the implicit GEMM  tok[m, e] = sum_k patch[m, k] W[e, k],  m = (b, pd, ph, pw) over
// B x 9^3 = 1458 tokens at 2 x 144^3, k = (kd, kh, kw) in the 16^3 patch (K = 4096), E = 128.
//   * A operand (128 tokens x 64 k per stage) is gathered straight from the fp32 NCDHW volume by 128 producer threads --
//     one token row each: four 64-byte image rows -> bf16 -> the K-major 128-byte-swizzled UMMA layout in shared memory
//     (generic-proxy stores + fence.proxy.async), no im2col buffer;
//   * B operand (E x 64 k) arrives by TMA from the bf16 copy of the weight;
//   * tcgen05.mma M = 128, N = E, fp32 accumulator in TMEM; K is split 8 ways across CTAs (the GEMM has only 12 M tiles) and the
//     fp32 partials are summed in a fixed order by patch_finish_kernel together with bias + position embedding + dropout
//     (simt_gemm.cu), exactly like the fp32 SIMT path it replaces.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {
using namespace tcptx;

constexpr int PT_THREADS = 192;      // warps 0-3: A producers, then epilogue; warp 4: MMA issuer + TMEM; warp 5: TMA (weights)
constexpr int PT_STAGES = 4;
constexpr int PT_SPLIT = 8;          // K = 4096 -> 8 x 512
constexpr int PT_KC = 64;            // k per stage = 4 image rows of 16 voxels

struct PtParams {
  const float* img;                  // offset to the modality plane of sample 0
  long long batch_stride;
  int D, H, W, M, E;
  float* part;                       // [PT_SPLIT][M][E]
};

__global__ void __launch_bounds__(PT_THREADS, 1) patch_embed_tc_kernel(const __grid_constant__ CUtensorMap tmw, const PtParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t a_bytes = 128u * PT_KC * 2u, b_bytes = (uint32_t)p.E * PT_KC * 2u;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023u) & ~1023u);
  const uint32_t bar_base = smem_base + PT_STAGES * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (PT_STAGES + s); };
  const uint32_t acc_bar = bar_base + 8u * (2 * PT_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * PT_STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  const uint32_t tmem_cols = p.E <= 32 ? 32u : p.E <= 64 ? 64u : p.E <= 128 ? 128u : 256u;

  if (warp == 5 && lane == 0) {
    tma_prefetch_desc(&tmw);
    for (int s = 0; s < PT_STAGES; ++s) { mbar_init(full_bar(s), 128 + 1); mbar_init(empty_bar(s), 1); }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int m0 = blockIdx.x * 128, split = blockIdx.y;
  const int kbase = split * (4096 / PT_SPLIT);
  constexpr int NCH = 4096 / PT_SPLIT / PT_KC;      // 8 stages of work per CTA

  if (warp < 4) {
    // ===== A producers: thread = token row
    const int r = threadIdx.x;
    int m = m0 + r;
    const bool valid = m < p.M;
    const int w16 = p.W / 16, h16 = p.H / 16, d16 = p.D / 16;
    const int pw = m % w16; m /= w16;
    const int ph = m % h16; m /= h16;
    const int pd = m % d16;
    const int b = m / d16;
    const float* base = p.img + b * p.batch_stride + ((long long)(pd * 16) * p.H + ph * 16) * p.W + pw * 16;
    const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
    uint32_t s = 0, ph_ = 0;
    for (int c = 0; c < NCH; ++c) {
      const int k0 = kbase + c * PT_KC;
      const int kd = k0 >> 8, kh0 = (k0 >> 4) & 15;
      float4 v[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4* src = reinterpret_cast<const float4*>(base + ((long long)kd * p.H + kh0 + j) * p.W);
#pragma unroll
        for (int q = 0; q < 4; ++q) v[j * 4 + q] = valid ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      mbar_wait(empty_bar(s), ph_ ^ 1u);
      const uint32_t dst = smem_base + s * stage_bytes + row_off;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {               // 16-byte chunk = 8 bf16 = two float4
        const float4 a = v[ch * 2], bq = v[ch * 2 + 1];
        __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(bq.x, bq.y), h3 = __floats2bfloat162_rn(bq.z, bq.w);
        const uint32_t phys = (uint32_t)(ch ^ (r & 7)) * 16u;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + phys), "r"(*reinterpret_cast<uint32_t*>(&h0)),
                     "r"(*reinterpret_cast<uint32_t*>(&h1)), "r"(*reinterpret_cast<uint32_t*>(&h2)),
                     "r"(*reinterpret_cast<uint32_t*>(&h3))
                     : "memory");
      }
      fence_proxy_async();                 // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(full_bar(s));
      if (++s == PT_STAGES) { s = 0; ph_ ^= 1u; }
    }
    // ===== epilogue: TMEM lane = token row -> fp32 partial row
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int mrow = m0 + r;
    float* out = p.part + ((long long)split * p.M + mrow) * p.E;
    for (int c0 = 0; c0 < p.E; c0 += 16) {
      uint32_t v16[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v16[0]), "=r"(v16[1]), "=r"(v16[2]), "=r"(v16[3]), "=r"(v16[4]), "=r"(v16[5]), "=r"(v16[6]), "=r"(v16[7]),
            "=r"(v16[8]), "=r"(v16[9]), "=r"(v16[10]), "=r"(v16[11]), "=r"(v16[12]), "=r"(v16[13]), "=r"(v16[14]), "=r"(v16[15])
          : "r"(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(out + c0 + q * 4) = make_float4(__uint_as_float(v16[q * 4]), __uint_as_float(v16[q * 4 + 1]),
                                                                   __uint_as_float(v16[q * 4 + 2]), __uint_as_float(v16[q * 4 + 3]));
      }
    }
  } else if (warp == 4) {
    // ===== MMA issuer
    const uint32_t issue = elect_one_sync() ? 1u : 0u;
    const uint32_t idesc = umma_idesc(128, p.E, 0, 0);
    const uint64_t desc_hi = umma_desc(0, 16, 1024, 2);
    uint32_t s = 0, ph_ = 0;
    for (int c = 0; c < NCH; ++c) {
      mbar_wait(full_bar(s), ph_);
      tc_fence_after();
      const uint32_t a0 = smem_base + s * stage_bytes, b0 = a0 + a_bytes;
      const uint64_t adesc = desc_hi | (uint64_t)((a0 >> 4) & 0x3FFF), bdesc = desc_hi | (uint64_t)((b0 >> 4) & 0x3FFF);
#pragma unroll
      for (int ks = 0; ks < PT_KC / 16; ++ks)
        umma_ss_p(tmem_base, adesc + (uint64_t)(2 * ks), bdesc + (uint64_t)(2 * ks), idesc, (c > 0 || ks > 0) ? 1u : 0u, issue);
      umma_commit_p(empty_bar(s), issue);
      if (++s == PT_STAGES) { s = 0; ph_ ^= 1u; }
    }
    umma_commit_p(acc_bar, issue);
  } else {
    // ===== TMA: weight tile [E rows x 64 k] per stage
    const uint32_t issue = elect_one_sync() ? 1u : 0u;
    uint32_t s = 0, ph_ = 0;
    for (int c = 0; c < NCH; ++c) {
      mbar_wait(empty_bar(s), ph_ ^ 1u);
      mbar_expect_tx_p(full_bar(s), b_bytes, issue);
      asm volatile(
          "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
          "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
          ::"r"(smem_base + s * stage_bytes + a_bytes), "l"(&tmw), "r"(full_bar(s)), "r"(kbase + c * PT_KC), "r"(0), "r"(issue)
          : "memory");
      if (++s == PT_STAGES) { s = 0; ph_ ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

__global__ void pt_cast_kernel(const float* __restrict__ w, bf16* __restrict__ out, long long n) {
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(w + i);
    *reinterpret_cast<__nv_bfloat162*>(out + i) = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(out + i + 2) = __floats2bfloat162_rn(v.z, v.w);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn pt_get_encode() {
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

// sums the split-K partials, adds bias + position embedding, applies dropout (simt_gemm.cu)
int hdf_patch_finish(const float* part, int S, float* out, long long ldo, const float* bias, const float* pos, int ntok, int E,
                     long long total, float p, const unsigned long long* seed_ptr, unsigned long long seed, unsigned call_id,
                     void* stream);

extern "C" {

int hdf_patch_embed_tc_supported(int E) { return E == 64 || E == 128 || E == 256; }

size_t hdf_patch_embed_tc_workspace(int B, int D, int H, int W, int E) {
  const long long M = (long long)B * (D / 16) * (H / 16) * (W / 16);
  return (size_t)E * 4096 * sizeof(bf16) + 1024 + (size_t)PT_SPLIT * M * E * sizeof(float);
}

// Same contract as hdf_patch_embed_fwd (tokens = patch_conv(img[:, modality]) + bias + pos, then dropout; fp32 out), bf16
// operands on the tensor cores with fp32 accumulation.
int hdf_patch_embed_tc_fwd(const float* img, int B, int Mch, int modality, int D, int H, int W, const float* weight,
                           const float* bias, const float* pos, float* out, long long ldo, int E, float p,
                           const unsigned long long* seed_ptr, unsigned long long seed, unsigned call_id, void* workspace,
                           size_t ws_bytes, void* stream) {
  HDF_REQUIRE(img && weight && out && workspace && (D % 16 == 0) && (H % 16 == 0) && (W % 16 == 0),
              "hdf_patch_embed_tc_fwd: bad args (every spatial dim must be a multiple of 16)");
  HDF_REQUIRE(hdf_patch_embed_tc_supported(E), "hdf_patch_embed_tc_fwd: unsupported E=%d", E);
  HDF_REQUIRE((uintptr_t)img % 16 == 0, "hdf_patch_embed_tc_fwd: image must be 16-byte aligned");
  HDF_REQUIRE(ws_bytes >= hdf_patch_embed_tc_workspace(B, D, H, W, E), "hdf_patch_embed_tc_fwd: workspace too small");
  EncodeTiledFn enc = pt_get_encode();
  if (!enc) { hdf_set_error("hdf_patch_embed_tc_fwd: cuTensorMapEncodeTiled unavailable"); return HDF_ERR_CUDA; }
  cudaStream_t s = (cudaStream_t)stream;
  const int ntok = (D / 16) * (H / 16) * (W / 16);
  const int M = B * ntok;
  // workspace: [bf16 weight E x 4096 | pad to 1 KB | fp32 partials]
  uint8_t* wsb = (uint8_t*)(((uintptr_t)workspace + 127) & ~(uintptr_t)127);
  bf16* wb = (bf16*)wsb;
  float* part = (float*)(wsb + (((size_t)E * 4096 * sizeof(bf16) + 127) & ~(size_t)127));
  pt_cast_kernel<<<128, 256, 0, s>>>(weight, wb, (long long)E * 4096);
  HDF_LAUNCH_CHECK("hdf_patch_embed_tc_fwd/cast");
  CUtensorMap tmw;
  {
    cuuint64_t gdim[2] = {4096, (cuuint64_t)E};
    cuuint64_t gstr[1] = {4096 * sizeof(bf16)};
    cuuint32_t box[2] = {PT_KC, (cuuint32_t)E};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wb, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { hdf_set_error("hdf_patch_embed_tc_fwd: encode(w) failed: %d", (int)r); return HDF_ERR_CUDA; }
  }
  PtParams prm{img + (long long)modality * D * H * W, (long long)Mch * D * H * W, D, H, W, M, E, part};
  const uint32_t stage_bytes = 128u * PT_KC * 2u + (((uint32_t)E * PT_KC * 2u + 1023u) & ~1023u);
  const size_t smem = (size_t)PT_STAGES * stage_bytes + 1024 + 8 * (2 * PT_STAGES + 3) + 64;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(patch_embed_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { hdf_set_error("hdf_patch_embed_tc_fwd: smem attribute: %s", cudaGetErrorString(e)); return HDF_ERR_CUDA; }
    configured = true;
  }
  patch_embed_tc_kernel<<<dim3(cdiv(M, 128), PT_SPLIT), PT_THREADS, smem, s>>>(tmw, prm);
  HDF_LAUNCH_CHECK("hdf_patch_embed_tc_fwd");
  return hdf_patch_finish(part, PT_SPLIT, out, ldo, bias, pos, ntok, E, (long long)M * E, p, seed_ptr, seed, call_id, stream);
}

}  // extern "C"
