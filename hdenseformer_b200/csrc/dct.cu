// Densely-connected-transformer (DCT) token kernels, fp32: LayerNorm fwd/bwd (warp per row,
// shuffle reductions), flash-style attention fwd/bwd for head_dim 4 (scores never materialised;
// softmax statistics in registers), GELU/dropout backward, positional-embedding gradient.
// Reference semantics: models/HDenseFormer.py:11-17 (PreNorm), :33-44 (DenseForward), :47-75
// (Dense_Attention).  Token GEMMs live in simt_gemm.cu (hdf_gemm_rowmajor / hdf_gemm_at_b).
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------
// LayerNorm over the last dim C (C <= 1024), one warp per row
// ---------------------------------------------------------------------------
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ out, long long ldo,
                                     float* __restrict__ mean, float* __restrict__ rstd, int M, int C, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= M) return;
  const float* xr = x + (long long)warp * ldx;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mu = warp_sum(s) / C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = xr[c] - mu; q += d * d; }
  const float rs = rsqrtf(warp_sum(q) / C + eps);
  for (int c = lane; c < C; c += 32) out[(long long)warp * ldo + c] = (xr[c] - mu) * rs * gamma[c] + beta[c];
  if (lane == 0) { mean[warp] = mu; rstd[warp] = rs; }
}

// dx (+)= LN backward ; partial[block][2][C] = per-block (dgamma, dbeta)
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, long long ldd,
                                                           const float* __restrict__ x, long long ldx,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma, float* __restrict__ dx,
                                                           long long ldo, int accumulate, int M, int C,
                                                           float* __restrict__ partial) {
  extern __shared__ float sm[];  // [8 warps][2][C]
  const int wid = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nw = blockDim.x / 32;
  for (int i = threadIdx.x; i < nw * 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  for (int row = blockIdx.x * nw + wid; row < M; row += gridDim.x * nw) {
    const float mu = mean[row], rs = rstd[row];
    float a = 0.f, b = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float g = dy[(long long)row * ldd + c] * gamma[c];
      const float xh = (x[(long long)row * ldx + c] - mu) * rs;
      a += g;
      b += g * xh;
    }
    a = warp_sum(a) / C;
    b = warp_sum(b) / C;
    for (int c = lane; c < C; c += 32) {
      const float d = dy[(long long)row * ldd + c];
      const float xh = (x[(long long)row * ldx + c] - mu) * rs;
      const float v = rs * (d * gamma[c] - a - xh * b);
      float* q = dx + (long long)row * ldo + c;
      *q = accumulate ? *q + v : v;
      sm[(wid * 2 + 0) * C + c] += d * xh;
      sm[(wid * 2 + 1) * C + c] += d;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += sm[w * 2 * C + i];
    partial[(long long)blockIdx.x * 2 * C + i] = s;
  }
}

// one warp per channel, lanes stride over the per-block partials (fixed order -> deterministic)
__global__ void ln_param_finalize_kernel(const float* __restrict__ partial, int blocks, int C, float* __restrict__ dgamma,
                                         float* __restrict__ dbeta, int accumulate) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int k = lane; k < blocks; k += 32) { a += partial[(long long)k * 2 * C + c]; b += partial[(long long)k * 2 * C + C + c]; }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane != 0) return;
  dgamma[c] = accumulate ? dgamma[c] + a : a;
  dbeta[c] = accumulate ? dbeta[c] + b : b;
}

// ---------------------------------------------------------------------------
// attention, heads*4 = inner dim.  qkv rows: [q(inner) | k(inner) | v(inner)]
// ---------------------------------------------------------------------------
// Attention (models/HDenseFormer.py:47-75): 8 heads of dim 4, N = 729 tokens at 144^3: tiny in FLOPs, long in dependent
// latency.  A query (key, for dk/dv) is owned by AP = 8 adjacent lanes that each walk every 8th key of the shared-memory
// tile and are merged with shuffles, so the serial loop per thread is N/8 long and the grid has 8x more threads than
// rows (B*H*N = 11.6 k rows would otherwise fill 4 % of the GPU); the forward rescales its running max once per tile
// (16 independent exponentials per thread per tile) instead of once per key.
constexpr int AT = 128;  // keys (queries for dk/dv) per shared-memory tile == threads per block
constexpr int AP = 8;    // lanes per owner row
constexpr int AQ = AT / AP;   // owner rows per block
constexpr int AU = AT / AP;   // tile rows walked per lane

__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// T = keys staged per pass (multiple of AT, <= 1024: all 729 tokens of the headline config in ONE global round trip;
// next to the persistent convolution kernels each dependent trip to global memory costs microseconds)
__global__ void __launch_bounds__(AT) attn_fwd_kernel(const float* __restrict__ qkv, long long ld, float* __restrict__ o,
                                                     long long ldo, float* __restrict__ lse, int N, int H, float scale, int T) {
  extern __shared__ float4 asm4[];
  float4* sk = asm4;
  float4* sv = asm4 + T;
  const int h = blockIdx.y, b = blockIdx.z;
  const int inner = 4 * H;
  const int part = threadIdx.x % AP;
  const int i = blockIdx.x * AQ + threadIdx.x / AP;
  const float* base = qkv + (long long)b * N * ld;
  float4 q = make_float4(0, 0, 0, 0);
  if (i < N) q = *reinterpret_cast<const float4*>(base + (long long)i * ld + 4 * h);
  q.x *= scale; q.y *= scale; q.z *= scale; q.w *= scale;
  float m = -INFINITY, l = 0.f;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int j0 = 0; j0 < N; j0 += T) {
    for (int t = threadIdx.x; t < T; t += AT) {
      const int j = j0 + t;
      if (j < N) {
        sk[t] = *reinterpret_cast<const float4*>(base + (long long)j * ld + inner + 4 * h);
        sv[t] = *reinterpret_cast<const float4*>(base + (long long)j * ld + 2 * inner + 4 * h);
      }
    }
    __syncthreads();
    const int tcnt = min(T, N - j0);
    for (int g0 = 0; g0 < tcnt; g0 += AT) {
      const int cnt = min(AT, tcnt - g0);
      float sc[AU];
      float mt = -INFINITY;
#pragma unroll
      for (int u = 0; u < AU; ++u) {
        const int t = part + u * AP;
        sc[u] = t < cnt ? dot4(q, sk[g0 + t]) : -INFINITY;
        mt = fmaxf(mt, sc[u]);
      }
      if (mt > -INFINITY) {
        const float mn = fmaxf(m, mt);
        const float r = __expf(m - mn);          // m = -inf on the first group -> 0
        l *= r; acc.x *= r; acc.y *= r; acc.z *= r; acc.w *= r;
        m = mn;
#pragma unroll
        for (int u = 0; u < AU; ++u) {
          const int t = part + u * AP;
          if (t < cnt) {
            const float pr = __expf(sc[u] - mn);
            const float4 v = sv[g0 + t];
            l += pr;
            acc.x = fmaf(pr, v.x, acc.x); acc.y = fmaf(pr, v.y, acc.y); acc.z = fmaf(pr, v.z, acc.z); acc.w = fmaf(pr, v.w, acc.w);
          }
        }
      }
    }
    __syncthreads();
  }
  // merge the 8 partial (m, l, acc) of the row
  float ma = m;
  ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1));
  ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
  ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 4));
  const float r = (m > -INFINITY) ? __expf(m - ma) : 0.f;
  l = group8_sum(l * r);
  acc.x = group8_sum(acc.x * r); acc.y = group8_sum(acc.y * r); acc.z = group8_sum(acc.z * r); acc.w = group8_sum(acc.w * r);
  if (i < N && part == 0) {
    const float inv = 1.f / l;
    *reinterpret_cast<float4*>(o + ((long long)b * N + i) * ldo + 4 * h) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    lse[((long long)b * H + h) * N + i] = ma + __logf(l);
  }
}

// dq: 8 lanes per query
__device__ __forceinline__ void attn_bwd_dq_body(float4* smem4, int T, int bx, const float* __restrict__ qkv, long long ld,
                                                 const float* __restrict__ o, long long ldo, const float* __restrict__ dout,
                                                 long long lddo, const float* __restrict__ lse, float* __restrict__ dqkv,
                                                 long long ldg, int N, int H, float scale) {
  float4* sk = smem4;
  float4* sv = smem4 + T;
  const int h = blockIdx.y, b = blockIdx.z;
  const int inner = 4 * H;
  const int part = threadIdx.x % AP;
  const int i = bx * AQ + threadIdx.x / AP;
  const float* base = qkv + (long long)b * N * ld;
  float4 q = make_float4(0, 0, 0, 0), dO = q, O = q;
  float L = 0.f;
  if (i < N) {
    q = *reinterpret_cast<const float4*>(base + (long long)i * ld + 4 * h);
    dO = *reinterpret_cast<const float4*>(dout + ((long long)b * N + i) * lddo + 4 * h);
    O = *reinterpret_cast<const float4*>(o + ((long long)b * N + i) * ldo + 4 * h);
    L = lse[((long long)b * H + h) * N + i];
  }
  const float Di = dot4(dO, O);
  float4 dq = make_float4(0, 0, 0, 0);
  for (int j0 = 0; j0 < N; j0 += T) {
    for (int t = threadIdx.x; t < T; t += AT) {
      const int j = j0 + t;
      if (j < N) {
        sk[t] = *reinterpret_cast<const float4*>(base + (long long)j * ld + inner + 4 * h);
        sv[t] = *reinterpret_cast<const float4*>(base + (long long)j * ld + 2 * inner + 4 * h);
      }
    }
    __syncthreads();
    const int tcnt = min(T, N - j0);
#pragma unroll 4
    for (int t = part; t < tcnt; t += AP) {
      const float4 k = sk[t], v = sv[t];
      const float pr = __expf(scale * dot4(q, k) - L);
      const float ds = pr * (dot4(dO, v) - Di);
      dq.x = fmaf(ds, k.x, dq.x); dq.y = fmaf(ds, k.y, dq.y); dq.z = fmaf(ds, k.z, dq.z); dq.w = fmaf(ds, k.w, dq.w);
    }
    __syncthreads();
  }
  dq.x = group8_sum(dq.x); dq.y = group8_sum(dq.y); dq.z = group8_sum(dq.z); dq.w = group8_sum(dq.w);
  if (i < N && part == 0)
    *reinterpret_cast<float4*>(dqkv + ((long long)b * N + i) * ldg + 4 * h) =
        make_float4(dq.x * scale, dq.y * scale, dq.z * scale, dq.w * scale);
}

// dk, dv: 8 lanes per key
__device__ __forceinline__ void attn_bwd_dkv_body(float4* smem4, int T, int bx, const float* __restrict__ qkv, long long ld,
                                                  const float* __restrict__ o, long long ldo, const float* __restrict__ dout,
                                                  long long lddo, const float* __restrict__ lse, float* __restrict__ dqkv,
                                                  long long ldg, int N, int H, float scale) {
  float4* sq = smem4;
  float4* sdo = smem4 + T;
  float* sL = reinterpret_cast<float*>(smem4 + 2 * T);
  float* sD = sL + T;
  const int h = blockIdx.y, b = blockIdx.z;
  const int inner = 4 * H;
  const int part = threadIdx.x % AP;
  const int j = bx * AQ + threadIdx.x / AP;
  const float* base = qkv + (long long)b * N * ld;
  float4 k = make_float4(0, 0, 0, 0), v = k;
  if (j < N) {
    k = *reinterpret_cast<const float4*>(base + (long long)j * ld + inner + 4 * h);
    v = *reinterpret_cast<const float4*>(base + (long long)j * ld + 2 * inner + 4 * h);
  }
  float4 dk = make_float4(0, 0, 0, 0), dv = dk;
  for (int i0 = 0; i0 < N; i0 += T) {
    for (int t = threadIdx.x; t < T; t += AT) {
      const int i = i0 + t;
      if (i < N) {
        const float4 q = *reinterpret_cast<const float4*>(base + (long long)i * ld + 4 * h);
        const float4 dO = *reinterpret_cast<const float4*>(dout + ((long long)b * N + i) * lddo + 4 * h);
        const float4 O = *reinterpret_cast<const float4*>(o + ((long long)b * N + i) * ldo + 4 * h);
        sq[t] = q;
        sdo[t] = dO;
        sL[t] = lse[((long long)b * H + h) * N + i];
        sD[t] = dot4(dO, O);
      }
    }
    __syncthreads();
    const int tcnt = min(T, N - i0);
#pragma unroll 4
    for (int t = part; t < tcnt; t += AP) {
      const float4 q = sq[t], dO = sdo[t];
      const float pr = __expf(scale * dot4(q, k) - sL[t]);
      dv.x = fmaf(pr, dO.x, dv.x); dv.y = fmaf(pr, dO.y, dv.y); dv.z = fmaf(pr, dO.z, dv.z); dv.w = fmaf(pr, dO.w, dv.w);
      const float ds = pr * (dot4(dO, v) - sD[t]);
      dk.x = fmaf(ds, q.x, dk.x); dk.y = fmaf(ds, q.y, dk.y); dk.z = fmaf(ds, q.z, dk.z); dk.w = fmaf(ds, q.w, dk.w);
    }
    __syncthreads();
  }
  dk.x = group8_sum(dk.x); dk.y = group8_sum(dk.y); dk.z = group8_sum(dk.z); dk.w = group8_sum(dk.w);
  dv.x = group8_sum(dv.x); dv.y = group8_sum(dv.y); dv.z = group8_sum(dv.z); dv.w = group8_sum(dv.w);
  if (j < N && part == 0) {
    float* g = dqkv + ((long long)b * N + j) * ldg;
    *reinterpret_cast<float4*>(g + inner + 4 * h) = make_float4(dk.x * scale, dk.y * scale, dk.z * scale, dk.w * scale);
    *reinterpret_cast<float4*>(g + 2 * inner + 4 * h) = dv;
  }
}

// one launch for both halves of the attention backward: blocks [0, nb) compute dq, blocks [nb, 2nb) compute dk/dv
__global__ void __launch_bounds__(AT) attn_bwd_kernel(const float* __restrict__ qkv, long long ld, const float* __restrict__ o,
                                                     long long ldo, const float* __restrict__ dout, long long lddo,
                                                     const float* __restrict__ lse, float* __restrict__ dqkv, long long ldg,
                                                     int N, int H, float scale, int T) {
  extern __shared__ float4 asm4[];
  const int nb = gridDim.x / 2;
  if ((int)blockIdx.x < nb) attn_bwd_dq_body(asm4, T, blockIdx.x, qkv, ld, o, ldo, dout, lddo, lse, dqkv, ldg, N, H, scale);
  else attn_bwd_dkv_body(asm4, T, blockIdx.x - nb, qkv, ld, o, ldo, dout, lddo, lse, dqkv, ldg, N, H, scale);
}

// dz[m,n] = dy[m,n] * dropout_scale(m*N+n) * gelu'(pre[m,n])   (act=1)   or   dy * dropout_scale (act=0)
__global__ void act_dropout_bwd_kernel(const float* __restrict__ dy, long long ldd, const float* __restrict__ pre,
                                       float* __restrict__ dz, long long ldz, int N, long long total, int act, float p,
                                       const unsigned long long* seed_ptr, unsigned long long seed, unsigned call_id) {
  if (seed_ptr) seed += *seed_ptr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / N;
    const int n = (int)(i % N);
    float g = dy[m * ldd + n] * hdf_dropout_scale(seed, call_id, (unsigned long long)i, p);
    if (act == 1) {
      const float z = pre[i];
      const float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * __expf(-0.5f * z * z);
      g *= cdf + z * pdf;
    }
    dz[m * ldz + n] = g;
  }
}

// dpos[t, e] (+)= sum_b dtok[b*ntok + t, e]
__global__ void posemb_grad_kernel(const float* __restrict__ dtok, long long ld, float* __restrict__ dpos, int B, int ntok,
                                   int E, int accumulate) {
  const long long total = (long long)ntok * E;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % E);
    const long long t = i / E;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dtok[((long long)b * ntok + t) * ld + e];
    dpos[i] = accumulate ? dpos[i] + s : s;
  }
}

// dst[m, 0:C] (ld) (+)= src[m, 0:C] (ld)    fp32 row slices
__global__ void add_rows_f32_kernel(float* __restrict__ dst, long long ldd, const float* __restrict__ src, long long lds,
                                    int C, long long total, int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / C;
    const int c = (int)(i % C);
    float* q = dst + m * ldd + c;
    const float v = src[m * lds + c];
    *q = accumulate ? *q + v : v;
  }
}

// keys staged per pass.  Measured inside the training step (profiles/r1_timeline_*.txt): staging all 729 keys at once
// (24-30 KB of shared memory per CTA) makes the kernel slower next to the persistent convolution CTAs than 128-key
// tiles (4-5 KB), although it is faster stand-alone; HDF_ATTN_TILE overrides for experiments.
int attn_tile(int N) {
  static const int cfg = getenv("HDF_ATTN_TILE") ? atoi(getenv("HDF_ATTN_TILE")) : AT;
  int t = (N + AT - 1) / AT * AT;
  if (t > cfg) t = cfg;
  t = t / AT * AT;
  return t < AT ? AT : (t > 1024 ? 1024 : t);
}

int ln_blocks(int M) {
  int b = cdiv(M, 8);
  return b > 296 ? 296 : (b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

int hdf_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float* out, long long ldo,
                      float* mean, float* rstd, int M, int C, float eps, void* stream) {
  HDF_REQUIRE(x && gamma && beta && out && mean && rstd, "hdf_layernorm_fwd: null pointer");
  layernorm_fwd_kernel<<<cdiv((long long)M * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, gamma, beta, out, ldo, mean,
                                                                                       rstd, M, C, eps);
  HDF_LAUNCH_CHECK("hdf_layernorm_fwd");
  return HDF_OK;
}

size_t hdf_layernorm_bwd_workspace(int M, int C) { return (size_t)ln_blocks(M) * 2 * C * sizeof(float); }

int hdf_layernorm_bwd(const float* dy, long long ldd, const float* x, long long ldx, const float* mean, const float* rstd,
                      const float* gamma, float* dx, long long ldo, int accumulate_dx, float* dgamma, float* dbeta,
                      int accumulate_params, int M, int C, void* workspace, size_t ws_bytes, void* stream) {
  HDF_REQUIRE(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && workspace, "hdf_layernorm_bwd: null pointer");
  HDF_REQUIRE(ws_bytes >= hdf_layernorm_bwd_workspace(M, C), "hdf_layernorm_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int blocks = ln_blocks(M);
  layernorm_bwd_kernel<<<blocks, 256, (size_t)8 * 2 * C * sizeof(float), s>>>(dy, ldd, x, ldx, mean, rstd, gamma, dx, ldo,
                                                                              accumulate_dx, M, C, (float*)workspace);
  HDF_LAUNCH_CHECK("hdf_layernorm_bwd");
  ln_param_finalize_kernel<<<cdiv((long long)C * 32, 128), 128, 0, s>>>((const float*)workspace, blocks, C, dgamma, dbeta, accumulate_params);
  HDF_LAUNCH_CHECK("hdf_layernorm_bwd/finalize");
  return HDF_OK;
}

int hdf_attention_fwd(const float* qkv, long long ld, float* o, long long ldo, float* lse, int B, int N, int H, float scale,
                      void* stream) {
  HDF_REQUIRE(qkv && o && lse && (ld % 4 == 0) && (ldo % 4 == 0), "hdf_attention_fwd: bad args");
  dim3 grid(cdiv(N, AQ), H, B);
  const int T = attn_tile(N);
  attn_fwd_kernel<<<grid, AT, (size_t)T * 32, (cudaStream_t)stream>>>(qkv, ld, o, ldo, lse, N, H, scale, T);
  HDF_LAUNCH_CHECK("hdf_attention_fwd");
  return HDF_OK;
}

int hdf_attention_bwd(const float* qkv, long long ld, const float* o, long long ldo, const float* dout, long long lddo,
                      const float* lse, float* dqkv, long long ldg, int B, int N, int H, float scale, void* stream) {
  HDF_REQUIRE(qkv && o && dout && lse && dqkv && (ld % 4 == 0) && (ldo % 4 == 0) && (lddo % 4 == 0) && (ldg % 4 == 0),
              "hdf_attention_bwd: bad args");
  dim3 grid(2 * cdiv(N, AQ), H, B);
  const int T = attn_tile(N);
  attn_bwd_kernel<<<grid, AT, (size_t)T * 40, (cudaStream_t)stream>>>(qkv, ld, o, ldo, dout, lddo, lse, dqkv, ldg, N, H, scale, T);
  HDF_LAUNCH_CHECK("hdf_attention_bwd");
  return HDF_OK;
}

int hdf_act_dropout_bwd(const float* dy, long long ldd, const float* pre, float* dz, long long ldz, int M, int N, int act,
                        float p, const unsigned long long* seed_ptr, unsigned long long seed, unsigned call_id, void* stream) {
  HDF_REQUIRE(dy && dz && (act == 0 || pre), "hdf_act_dropout_bwd: null pointer");
  const long long total = (long long)M * N;
  act_dropout_bwd_kernel<<<min(2048, cdiv(total, 256)), 256, 0, (cudaStream_t)stream>>>(dy, ldd, pre, dz, ldz, N, total, act,
                                                                                      p, seed_ptr, seed, call_id);
  HDF_LAUNCH_CHECK("hdf_act_dropout_bwd");
  return HDF_OK;
}

int hdf_posemb_grad(const float* dtok, long long ld, float* dpos, int B, int ntok, int E, int accumulate, void* stream) {
  HDF_REQUIRE(dtok && dpos, "hdf_posemb_grad: null pointer");
  posemb_grad_kernel<<<min(1024, cdiv((long long)ntok * E, 256)), 256, 0, (cudaStream_t)stream>>>(dtok, ld, dpos, B, ntok, E,
                                                                                                accumulate);
  HDF_LAUNCH_CHECK("hdf_posemb_grad");
  return HDF_OK;
}

int hdf_add_rows_f32(float* dst, long long ldd, const float* src, long long lds, long long rows, int C, int accumulate,
                     void* stream) {
  HDF_REQUIRE(dst && src, "hdf_add_rows_f32: null pointer");
  const long long total = rows * C;
  add_rows_f32_kernel<<<min(2048, cdiv(total, 256)), 256, 0, (cudaStream_t)stream>>>(dst, ldd, src, lds, C, total, accumulate);
  HDF_LAUNCH_CHECK("hdf_add_rows_f32");
  return HDF_OK;
}

}  // extern "C"

// =============================================================================================================
// Fused row-local chain of one DCT inner layer (models/HDenseFormer.py:95-98):
//     h1 = drop_a(o Wo^T + bo) + h0 ;  h2 = FF(LN2(h1)) + h1 ;  feature = FF(LN2(h2))
//     FF(x) = drop(W2 drop(gelu(W1 x + b1)) + b2)        (the SAME ff module twice, fresh masks)
// One block = 32 token rows, 4 threads per row; weights staged (transposed) in shared memory.  Replaces 7 launches
// in forward and ~35 in backward; weight-gradient contributions are reduced inside the block and written as
// per-block partials that dct_c_reduce_kernel sums in a fixed order (deterministic).
// =============================================================================================================
namespace {

constexpr int FR = 32;      // rows per block
constexpr int FG = 32;      // growth rate (token width)
constexpr int FH = 64;      // mlp hidden width
// per-block weight-gradient partial layout (floats)
constexpr int P_W2 = 0, P_B2 = P_W2 + FG * FH, P_W1 = P_B2 + FG, P_B1 = P_W1 + FH * FG, P_WO = P_B1 + FH,
              P_BO = P_WO + FG * FG, P_GM = P_BO + FG, P_BT = P_GM + FG, P_TOTAL = P_BT + FG;

__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad(float z) {
  const float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * z * z);
  return cdf + z * pdf;
}
// out[j] += sum_k xs[k] * Ws[k*N + sub*(N/4) + j]      (row vector in smem, weights [K][N] in smem)
template <int K, int N>
__device__ __forceinline__ void rowmm(const float* xs, const float* Ws, int sub, float* out) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float x = xs[k];
    const float* w = Ws + k * N + sub * (N / 4);
#pragma unroll
    for (int j = 0; j < N / 4; ++j) out[j] = fmaf(x, w[j], out[j]);
  }
}
// sum over the 4 threads that share a row (adjacent lanes)
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

struct DctCParams {
  // activations (row-major fp32)
  const float* o; const float* h0;
  float* h1; float* n2; float* z1; float* f1; float* h2; float* n3; float* z1b; float* g1;
  float* m2; float* r2; float* m3; float* r3;
  float* fout; long long ldf;
  // parameters
  const float* Wo; const float* bo; const float* gm; const float* bt; const float* W1; const float* b1; const float* W2;
  const float* b2;
  int R;
  float p; const unsigned long long* seed_ptr; unsigned long long seed; unsigned ida, idb, idc, idd, ide;
};

__global__ void __launch_bounds__(128) dct_c_fwd_kernel(const DctCParams q) {
  extern __shared__ float sm[];
  float* WoT = sm;                  // [32][32]   WoT[k][j] = Wo[j][k]
  float* W1T = WoT + FG * FG;       // [32][64]
  float* W2T = W1T + FG * FH;       // [64][32]
  float* xs = W2T + FH * FG;        // [32 rows][64]
  const int tid = threadIdx.x;
  const int rl = tid / 4, sub = tid % 4;
  const long long row = (long long)blockIdx.x * FR + rl;
  const bool ok = row < q.R;
  const long long rr = ok ? row : 0;
  // every global input of the thread is requested up front, together with the weights: one round trip to memory
  // instead of three dependent ones (this kernel sits on the forward critical path next to the convolution kernels)
  float p_o[8], p_h0[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { p_o[j] = q.o[rr * FG + sub * 8 + j]; p_h0[j] = q.h0[rr * FG + sub * 8 + j]; }
  const unsigned long long seed = q.seed + (q.seed_ptr ? *q.seed_ptr : 0ull);
  for (int i = tid; i < FG * FG; i += 128) WoT[(i % FG) * FG + i / FG] = q.Wo[i];
  for (int i = tid; i < FH * FG; i += 128) W1T[(i % FG) * FH + i / FG] = q.W1[i];   // W1 [64][32]
  for (int i = tid; i < FG * FH; i += 128) W2T[(i % FH) * FG + i / FH] = q.W2[i];   // W2 [32][64]
  float* xr = xs + rl * FH;
  // ---- h1 = drop_a(o Wo^T + bo) + h0
#pragma unroll
  for (int j = 0; j < 8; ++j) xr[sub * 8 + j] = p_o[j];
  __syncthreads();
  float h1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) h1[j] = q.bo[sub * 8 + j];
  rowmm<FG, FG>(xr, WoT, sub, h1);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = sub * 8 + j;
    h1[j] = h1[j] * hdf_dropout_scale(seed, q.ida, (unsigned long long)rr * FG + c, q.p) + p_h0[j];
  }
  float hcur[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) hcur[j] = h1[j];
  for (int pass = 0; pass < 2; ++pass) {
    // ---- n = LN2(hcur)
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += hcur[j];
    const float mu = quad_sum(s) * (1.f / FG);
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = hcur[j] - mu; v += d * d; }
    const float rs = rsqrtf(quad_sum(v) * (1.f / FG) + 1e-5f);
    float nn[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) nn[j] = (hcur[j] - mu) * rs * q.gm[sub * 8 + j] + q.bt[sub * 8 + j];
    float* hsave = pass == 0 ? q.h1 : q.h2;
    float* nsave = pass == 0 ? q.n2 : q.n3;
    if (ok) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { hsave[row * FG + sub * 8 + j] = hcur[j]; nsave[row * FG + sub * 8 + j] = nn[j]; }
      if (sub == 0) { (pass == 0 ? q.m2 : q.m3)[row] = mu; (pass == 0 ? q.r2 : q.r3)[row] = rs; }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) xr[sub * 8 + j] = nn[j];
    __syncwarp();
    // ---- z = W1 n + b1 ; f = drop(gelu(z))
    float z[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) z[j] = q.b1[sub * 16 + j];
    rowmm<FG, FH>(xr, W1T, sub, z);
    float* zsave = pass == 0 ? q.z1 : q.z1b;
    float* fsave = pass == 0 ? q.f1 : q.g1;
    const unsigned id1 = pass == 0 ? q.idb : q.idd, id2 = pass == 0 ? q.idc : q.ide;
    float f[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = sub * 16 + j;
      f[j] = gelu_f(z[j]) * hdf_dropout_scale(seed, id1, (unsigned long long)rr * FH + c, q.p);
      if (ok) { zsave[row * FH + c] = z[j]; fsave[row * FH + c] = f[j]; }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; ++j) xr[sub * 16 + j] = f[j];
    __syncwarp();
    // ---- y = drop(W2 f + b2)
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = q.b2[sub * 8 + j];
    rowmm<FH, FG>(xr, W2T, sub, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] *= hdf_dropout_scale(seed, id2, (unsigned long long)rr * FG + sub * 8 + j, q.p);
    if (pass == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) hcur[j] = y[j] + h1[j];     // h2 = FF(LN(h1)) + h1
    } else if (ok) {
#pragma unroll
      for (int j = 0; j < 8; ++j) q.fout[row * q.ldf + sub * 8 + j] = y[j];   // appended feature (no residual)
    }
    __syncwarp();
  }
}

struct DctCBwdParams {
  const float* dg2; long long ldg;           // grad of the appended feature (slice of dF)
  const float* o; const float* h1; const float* n2; const float* z1; const float* f1; const float* h2; const float* n3;
  const float* z1b; const float* g1; const float* m2; const float* r2; const float* m3; const float* r3;
  const float* Wo; const float* gm; const float* W1; const float* W2;
  float* d_o; float* dh1;                    // outputs [R,32]
  float* partial;                            // [gridDim.x][P_TOTAL]
  int R;
  float p; const unsigned long long* seed_ptr; unsigned long long seed; unsigned ida, idb, idc, idd, ide;
};

__global__ void __launch_bounds__(128) dct_c_bwd_kernel(const DctCBwdParams q) {
  extern __shared__ float sm[];
  float* Wo = sm;                    // [32][32] natural:  d_o[k] = sum_j dzo[j] Wo[j][k]
  float* W1 = Wo + FG * FG;          // [64][32] natural:  dn[k]  = sum_j dz[j]  W1[j][k]
  float* W2 = W1 + FH * FG;          // [32][64] natural:  df[k]  = sum_j dzz[j] W2[j][k]
  float* xs = W2 + FG * FH;          // [32][64] row scratch
  float* s_dzz = xs + FR * FH;       // [32][32]   grads wrt pre-dropout FF outputs (2nd use)
  float* s_dzz2 = s_dzz + FR * FG;   // [32][32]   (1st use)
  float* s_dz1b = s_dzz2 + FR * FG;  // [32][64]
  float* s_dz1 = s_dz1b + FR * FH;   // [32][64]
  float* s_dzo = s_dz1 + FR * FH;    // [32][32]
  float* s_g1 = s_dzo + FR * FG;     // [32][64]
  float* s_f1 = s_g1 + FR * FH;      // [32][64]
  float* s_n3 = s_f1 + FR * FH;      // [32][32]
  float* s_n2 = s_n3 + FR * FG;      // [32][32]
  float* s_o = s_n2 + FR * FG;       // [32][32]
  float* s_dgm = s_o + FR * FG;      // [32][32]  per-row d(gamma) contributions
  float* s_dbt = s_dgm + FR * FG;    // [32][32]
  const int tid = threadIdx.x;
  for (int i = tid; i < FG * FG; i += 128) Wo[i] = q.Wo[i];
  for (int i = tid; i < FH * FG; i += 128) W1[i] = q.W1[i];
  for (int i = tid; i < FG * FH; i += 128) W2[i] = q.W2[i];
  const int rl = tid / 4, sub = tid % 4;
  const long long row = (long long)blockIdx.x * FR + rl;
  const bool ok = row < q.R;
  const long long rr = ok ? row : 0;
  const float live = ok ? 1.f : 0.f;
  const unsigned long long seed = q.seed + (q.seed_ptr ? *q.seed_ptr : 0ull);
  // per-row inputs of both passes, requested before anything is consumed (one round trip instead of six)
  float p_dg2[8], p_zb[16], p_z[16], p_h2[8], p_h1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = sub * 8 + j;
    p_dg2[j] = q.dg2[rr * q.ldg + c];
    p_h2[j] = q.h2[rr * FG + c];
    p_h1[j] = q.h1[rr * FG + c];
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) { p_zb[j] = q.z1b[rr * FH + sub * 16 + j]; p_z[j] = q.z1[rr * FH + sub * 16 + j]; }
  const float p_m3 = q.m3[rr], p_m2 = q.m2[rr], p_r3 = q.r3[rr], p_r2 = q.r2[rr];
  // stage the saved activations needed by the weight gradients
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    s_g1[rl * FH + sub * 16 + j] = live * q.g1[rr * FH + sub * 16 + j];
    s_f1[rl * FH + sub * 16 + j] = live * q.f1[rr * FH + sub * 16 + j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_n3[rl * FG + sub * 8 + j] = live * q.n3[rr * FG + sub * 8 + j];
    s_n2[rl * FG + sub * 8 + j] = live * q.n2[rr * FG + sub * 8 + j];
    s_o[rl * FG + sub * 8 + j] = live * q.o[rr * FG + sub * 8 + j];
  }
  __syncthreads();
  float* xr = xs + rl * FH;
  float dgm_acc[8], dbt_acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) dgm_acc[j] = dbt_acc[j] = 0.f;
  float dh[8];   // running gradient of the residual stream
  // ======== second FF use (feature branch): dg2 -> dh2
  float dzz[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = sub * 8 + j;
    dzz[j] = live * p_dg2[j] * hdf_dropout_scale(seed, q.ide, (unsigned long long)rr * FG + c, q.p);
    s_dzz[rl * FG + c] = dzz[j];
    xr[c] = dzz[j];
  }
  __syncwarp();
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    // d(hidden) = dzz W2 ; dz = d(hidden) * mask * gelu'(z)
    float dhid[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) dhid[j] = 0.f;
    rowmm<FG, FH>(xr, W2, sub, dhid);
    const unsigned idh = pass == 0 ? q.idd : q.idb;
    float* s_dz = pass == 0 ? s_dz1b : s_dz1;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = sub * 16 + j;
      const float dz = dhid[j] * hdf_dropout_scale(seed, idh, (unsigned long long)rr * FH + c, q.p) * gelu_grad(pass == 0 ? p_zb[j] : p_z[j]);
      s_dz[rl * FH + c] = live * dz;
      xr[c] = dz;
    }
    __syncwarp();
    // dn = dz W1 ; LayerNorm backward
    float dn[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dn[j] = 0.f;
    rowmm<FH, FG>(xr, W1, sub, dn);
    const float mu = pass == 0 ? p_m3 : p_m2, rs = pass == 0 ? p_r3 : p_r2;
    float xh[8], a = 0.f, b = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = sub * 8 + j;
      xh[j] = ((pass == 0 ? p_h2[j] : p_h1[j]) - mu) * rs;
      const float g = dn[j] * q.gm[c];
      a += g;
      b += g * xh[j];
      dgm_acc[j] += live * dn[j] * xh[j];
      dbt_acc[j] += live * dn[j];
    }
    a = quad_sum(a) * (1.f / FG);
    b = quad_sum(b) * (1.f / FG);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = rs * (dn[j] * q.gm[sub * 8 + j] - a - xh[j] * b);
      dh[j] = pass == 0 ? v : dh[j] + v;       // pass 0: dh2 (h2 feeds only LN) ; pass 1: dh1 = dh2 + LN-branch
    }
    __syncwarp();
    if (pass == 0) {
      // first FF use: h2 = drop_c(W2 f1 + b2) + h1
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = sub * 8 + j;
        const float d2 = dh[j] * hdf_dropout_scale(seed, q.idc, (unsigned long long)rr * FG + c, q.p);
        s_dzz2[rl * FG + c] = live * d2;
        xr[c] = d2;
      }
      __syncwarp();
    }
  }
  // ======== h1 = drop_a(o Wo^T + bo) + h0
  float dzo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = sub * 8 + j;
    dzo[j] = dh[j] * hdf_dropout_scale(seed, q.ida, (unsigned long long)rr * FG + c, q.p);
    s_dzo[rl * FG + c] = live * dzo[j];
    xr[c] = dzo[j];
    s_dgm[rl * FG + c] = dgm_acc[j];
    s_dbt[rl * FG + c] = dbt_acc[j];
  }
  __syncwarp();
  float d_o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) d_o[j] = 0.f;
  rowmm<FG, FG>(xr, Wo, sub, d_o);
  if (ok) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { q.d_o[row * FG + sub * 8 + j] = d_o[j]; q.dh1[row * FG + sub * 8 + j] = dh[j]; }
  }
  __syncthreads();
  // ======== block-level weight-gradient partials
  float* part = q.partial + (long long)blockIdx.x * P_TOTAL;
  for (int i = tid; i < FG * FH; i += 128) {          // dW2[j][k] (W2 is [32][64])
    const int j = i / FH, k = i % FH;
    float s = 0.f;
    for (int r = 0; r < FR; ++r) s += s_dzz[r * FG + j] * s_g1[r * FH + k] + s_dzz2[r * FG + j] * s_f1[r * FH + k];
    part[P_W2 + i] = s;
  }
  for (int i = tid; i < FH * FG; i += 128) {          // dW1[j][k] (W1 is [64][32])
    const int j = i / FG, k = i % FG;
    float s = 0.f;
    for (int r = 0; r < FR; ++r) s += s_dz1b[r * FH + j] * s_n3[r * FG + k] + s_dz1[r * FH + j] * s_n2[r * FG + k];
    part[P_W1 + i] = s;
  }
  for (int i = tid; i < FG * FG; i += 128) {          // dWo[j][k]
    const int j = i / FG, k = i % FG;
    float s = 0.f;
    for (int r = 0; r < FR; ++r) s += s_dzo[r * FG + j] * s_o[r * FG + k];
    part[P_WO + i] = s;
  }
  if (tid < FG) {
    float sb2 = 0.f, sbo = 0.f, sg = 0.f, sb = 0.f;
    for (int r = 0; r < FR; ++r) {
      sb2 += s_dzz[r * FG + tid] + s_dzz2[r * FG + tid];
      sbo += s_dzo[r * FG + tid];
      sg += s_dgm[r * FG + tid];
      sb += s_dbt[r * FG + tid];
    }
    part[P_B2 + tid] = sb2; part[P_BO + tid] = sbo; part[P_GM + tid] = sg; part[P_BT + tid] = sb;
  }
  if (tid < FH) {
    float s = 0.f;
    for (int r = 0; r < FR; ++r) s += s_dz1b[r * FH + tid] + s_dz1[r * FH + tid];
    part[P_B1 + tid] = s;
  }
}

struct DctCGrads { float* w2; float* b2; float* w1; float* b1; float* wo; float* bo; float* gm; float* bt; };

__global__ void dct_c_reduce_kernel(const float* __restrict__ partial, int nblocks, DctCGrads g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P_TOTAL) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partial[(long long)b * P_TOTAL + i];
  float* dst;
  if (i < P_B2) dst = g.w2 + (i - P_W2);
  else if (i < P_W1) dst = g.b2 + (i - P_B2);
  else if (i < P_B1) dst = g.w1 + (i - P_W1);
  else if (i < P_WO) dst = g.b1 + (i - P_B1);
  else if (i < P_BO) dst = g.wo + (i - P_WO);
  else if (i < P_GM) dst = g.bo + (i - P_BO);
  else if (i < P_BT) dst = g.gm + (i - P_GM);
  else dst = g.bt + (i - P_BT);
  *dst += s;
}

constexpr size_t DCT_C_FWD_SMEM = (size_t)(FG * FG + FG * FH + FH * FG + FR * FH) * sizeof(float);
constexpr size_t DCT_C_BWD_SMEM =
    (size_t)(FG * FG + FH * FG + FG * FH + FR * FH + 2 * FR * FG + 2 * FR * FH + FR * FG + 2 * FR * FH + 3 * FR * FG + 2 * FR * FG) *
    sizeof(float);

}  // namespace

extern "C" {

// Fused forward of the post-attention chain of one DCT inner layer.  All row tensors are dense fp32 [R, 32] / [R, 64];
// fout is the feature slice (stride ldf).  growth_rate 32 / mlp 64 / are the reference's constants.
int hdf_dct_c_fwd(const float* o, const float* h0, float* h1, float* n2, float* z1, float* f1, float* h2, float* n3, float* z1b,
                  float* g1, float* m2, float* r2, float* m3, float* r3, float* fout, long long ldf, const float* Wo,
                  const float* bo, const float* gm, const float* bt, const float* W1, const float* b1, const float* W2,
                  const float* b2, int R, float p, const unsigned long long* seed_ptr, unsigned long long seed, unsigned ida,
                  unsigned idb, unsigned idc, unsigned idd, unsigned ide, void* stream) {
  HDF_REQUIRE(o && h0 && h1 && n2 && z1 && f1 && h2 && n3 && z1b && g1 && m2 && r2 && m3 && r3 && fout && Wo && bo && gm && bt &&
                  W1 && b1 && W2 && b2 && R > 0, "hdf_dct_c_fwd: null pointer");
  DctCParams q{o, h0, h1, n2, z1, f1, h2, n3, z1b, g1, m2, r2, m3, r3, fout, ldf, Wo, bo, gm, bt, W1, b1, W2, b2, R, p, seed_ptr,
               seed, ida, idb, idc, idd, ide};
  dct_c_fwd_kernel<<<cdiv(R, FR), 128, DCT_C_FWD_SMEM, (cudaStream_t)stream>>>(q);
  HDF_LAUNCH_CHECK("hdf_dct_c_fwd");
  return HDF_OK;
}

size_t hdf_dct_c_bwd_workspace(int R) { return (size_t)cdiv(R, FR) * P_TOTAL * sizeof(float); }

// Fused backward of the same chain: d_o (grad of the attention output) and dh1 (grad of the residual stream entering the
// layer's attention sub-block) plus += into the eight parameter gradients.
int hdf_dct_c_bwd(const float* dg2, long long ldg, const float* o, const float* h1, const float* n2, const float* z1,
                  const float* f1, const float* h2, const float* n3, const float* z1b, const float* g1, const float* m2,
                  const float* r2, const float* m3, const float* r3, const float* Wo, const float* gm, const float* W1,
                  const float* W2, float* d_o, float* dh1, float* dW2, float* db2, float* dW1, float* db1, float* dWo, float* dbo,
                  float* dgm, float* dbt, int R, float p, const unsigned long long* seed_ptr, unsigned long long seed,
                  unsigned ida, unsigned idb, unsigned idc, unsigned idd, unsigned ide, void* workspace, size_t ws_bytes,
                  void* stream) {
  HDF_REQUIRE(dg2 && o && h1 && n2 && z1 && f1 && h2 && n3 && z1b && g1 && m2 && r2 && m3 && r3 && Wo && gm && W1 && W2 && d_o &&
                  dh1 && dW2 && db2 && dW1 && db1 && dWo && dbo && dgm && dbt && workspace, "hdf_dct_c_bwd: null pointer");
  HDF_REQUIRE(ws_bytes >= hdf_dct_c_bwd_workspace(R), "hdf_dct_c_bwd: workspace too small");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dct_c_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DCT_C_BWD_SMEM);
    if (e != cudaSuccess) { hdf_set_error("hdf_dct_c_bwd: smem attribute: %s", cudaGetErrorString(e)); return HDF_ERR_CUDA; }
    configured = true;
  }
  const int nb = cdiv(R, FR);
  DctCBwdParams q{dg2, ldg, o, h1, n2, z1, f1, h2, n3, z1b, g1, m2, r2, m3, r3, Wo, gm, W1, W2, d_o, dh1, (float*)workspace, R, p,
                  seed_ptr, seed, ida, idb, idc, idd, ide};
  dct_c_bwd_kernel<<<nb, 128, DCT_C_BWD_SMEM, (cudaStream_t)stream>>>(q);
  HDF_LAUNCH_CHECK("hdf_dct_c_bwd");
  DctCGrads g{dW2, db2, dW1, db1, dWo, dbo, dgm, dbt};
  dct_c_reduce_kernel<<<cdiv(P_TOTAL, 128), 128, 0, (cudaStream_t)stream>>>((const float*)workspace, nb, g);
  HDF_LAUNCH_CHECK("hdf_dct_c_bwd/reduce");
  return HDF_OK;
}

}  // extern "C"

// =============================================================================================================
// Fused head of one DCT inner layer (models/HDenseFormer.py:94-96, :57):
//     h0 = Linear_l(cat(features)) ;  n1 = LN1(h0) ;  qkv = n1 Wqkv^T
// and its backward (dn1 = dqkv Wqkv ; dh0 = dh1 + LN1'(dn1) ; dF[:, :Cl] += dh0 Wl ; weight-gradient partials).
// Same organisation as the dct_c kernels: 32 rows per block, 4 threads per row, weights in shared memory.
// =============================================================================================================
namespace {

constexpr int FQ = 96;   // 3 * growth

// out[j] (j < NOUT/4 per thread) with runtime K:  out += xs[0..K) * Ws[k*NOUT + sub*(NOUT/4) + j]
template <int NOUT>
__device__ __forceinline__ void rowmm_k(const float* xs, const float* Ws, int K, int sub, float* out) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float x = xs[k];
    const float* w = Ws + k * NOUT + sub * (NOUT / 4);
#pragma unroll
    for (int j = 0; j < NOUT / 4; ++j) out[j] = fmaf(x, w[j], out[j]);
  }
}

struct DctAParams {
  const float* F; long long ldf; int Cl;
  const float* Wl; const float* bl; const float* gm; const float* bt; const float* Wqkv;
  float* h0; float* n1; float* m1; float* r1; float* qkv;
  int R;
};

__global__ void __launch_bounds__(128) dct_a_fwd_kernel(const DctAParams q) {
  extern __shared__ float sm[];
  const int Cl = q.Cl;
  float* WlT = sm;                    // [Cl][32]
  float* WqT = WlT + Cl * FG;         // [32][96]
  float* fs = WqT + FG * FQ;          // [32 rows][Cl]
  float* xs = fs + FR * Cl;           // [32 rows][32]
  const int tid = threadIdx.x;
  for (int i = tid; i < FG * Cl; i += 128) WlT[(i % Cl) * FG + i / Cl] = q.Wl[i];     // Wl [32][Cl]
  for (int i = tid; i < FQ * FG; i += 128) WqT[(i % FG) * FQ + i / FG] = q.Wqkv[i];   // Wqkv [96][32]
  const long long row0 = (long long)blockIdx.x * FR;
  for (int i = tid; i < FR * Cl; i += 128) {
    const int r = i / Cl, k = i % Cl;
    fs[i] = (row0 + r < q.R) ? q.F[(row0 + r) * q.ldf + k] : 0.f;
  }
  __syncthreads();
  const int rl = tid / 4, sub = tid % 4;
  const long long row = row0 + rl;
  const bool ok = row < q.R;
  float h[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = q.bl[sub * 8 + j];
  rowmm_k<FG>(fs + rl * Cl, WlT, Cl, sub, h);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += h[j];
  const float mu = quad_sum(s) * (1.f / FG);
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { const float d = h[j] - mu; v += d * d; }
  const float rs = rsqrtf(quad_sum(v) * (1.f / FG) + 1e-5f);
  float nn[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    nn[j] = (h[j] - mu) * rs * q.gm[sub * 8 + j] + q.bt[sub * 8 + j];
    xs[rl * FG + sub * 8 + j] = nn[j];
  }
  if (ok) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { q.h0[row * FG + sub * 8 + j] = h[j]; q.n1[row * FG + sub * 8 + j] = nn[j]; }
    if (sub == 0) { q.m1[row] = mu; q.r1[row] = rs; }
  }
  __syncwarp();
  float o[24];
#pragma unroll
  for (int j = 0; j < 24; ++j) o[j] = 0.f;
  rowmm<FG, FQ>(xs + rl * FG, WqT, sub, o);
  if (ok) {
#pragma unroll
    for (int j = 0; j < 24; ++j) q.qkv[row * FQ + sub * 24 + j] = o[j];
  }
}

struct DctABwdParams {
  const float* dqkv; const float* dh1; const float* h0; const float* n1; const float* m1; const float* r1;
  const float* F; long long ldf; int Cl;
  const float* Wqkv; const float* gm; const float* Wl;
  float* dF; long long lddf;
  float* partial;      // [grid][FQ*FG + 2*FG + FG*Cl + FG]
  int R;
};

__global__ void __launch_bounds__(128) dct_a_bwd_kernel(const DctABwdParams q) {
  extern __shared__ float sm[];
  const int Cl = q.Cl;
  float* Wq = sm;                      // [96][32] natural
  float* Wl = Wq + FQ * FG;            // [32][Cl] natural
  float* s_dq = Wl + FG * Cl;          // [32][96]
  float* s_n1 = s_dq + FR * FQ;        // [32][32]
  float* s_dh0 = s_n1 + FR * FG;       // [32][32]
  float* s_f = s_dh0 + FR * FG;        // [32][Cl]
  float* s_dgm = s_f + FR * Cl;        // [32][32]
  float* s_dbt = s_dgm + FR * FG;      // [32][32]
  const int tid = threadIdx.x;
  for (int i = tid; i < FQ * FG; i += 128) Wq[i] = q.Wqkv[i];
  for (int i = tid; i < FG * Cl; i += 128) Wl[i] = q.Wl[i];
  const long long row0 = (long long)blockIdx.x * FR;
  for (int i = tid; i < FR * Cl; i += 128) {
    const int r = i / Cl, k = i % Cl;
    s_f[i] = (row0 + r < q.R) ? q.F[(row0 + r) * q.ldf + k] : 0.f;
  }
  for (int i = tid; i < FR * FQ; i += 128) {
    const int r = i / FQ;
    s_dq[i] = (row0 + r < q.R) ? q.dqkv[(row0 + r) * FQ + i % FQ] : 0.f;
  }
  for (int i = tid; i < FR * FG; i += 128) {
    const int r = i / FG;
    s_n1[i] = (row0 + r < q.R) ? q.n1[(row0 + r) * FG + i % FG] : 0.f;
  }
  __syncthreads();
  const int rl = tid / 4, sub = tid % 4;
  const long long row = row0 + rl;
  const bool ok = row < q.R;
  const long long rr = ok ? row : 0;
  const float live = ok ? 1.f : 0.f;
  // dn1 = dqkv Wqkv
  float dn[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) dn[j] = 0.f;
  rowmm<FQ, FG>(s_dq + rl * FQ, Wq, sub, dn);
  // LayerNorm backward, plus the residual-stream gradient coming from the layer's tail
  const float mu = q.m1[rr], rs = q.r1[rr];
  float xh[8], a = 0.f, b = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = sub * 8 + j;
    xh[j] = (q.h0[rr * FG + c] - mu) * rs;
    const float g = dn[j] * q.gm[c];
    a += g;
    b += g * xh[j];
    s_dgm[rl * FG + c] = live * dn[j] * xh[j];
    s_dbt[rl * FG + c] = live * dn[j];
  }
  a = quad_sum(a) * (1.f / FG);
  b = quad_sum(b) * (1.f / FG);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = sub * 8 + j;
    const float dh0 = live * (q.dh1[rr * FG + c] + rs * (dn[j] * q.gm[c] - a - xh[j] * b));
    s_dh0[rl * FG + c] = dh0;
  }
  __syncwarp();
  // dF[row, :Cl] += dh0 Wl    (each of the 4 threads of the row owns Cl/4 consecutive columns, 8 at a time)
  if (ok) {
    const int per = Cl / 4;
    for (int c0 = 0; c0 < per; c0 += 8) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 4
      for (int k = 0; k < FG; ++k) {
        const float x = s_dh0[rl * FG + k];
        const float* w = Wl + k * Cl + sub * per + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(x, w[j], acc[j]);
      }
      float* dst = q.dF + row * q.lddf + sub * per + c0;
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] += acc[j];
    }
  }
  __syncthreads();
  // block-level weight-gradient partials: [dWqkv 96x32][dgm 32][dbt 32][dWl 32xCl][dbl 32]
  float* part = q.partial + (long long)blockIdx.x * (FQ * FG + 2 * FG + FG * Cl + FG);
  for (int i = tid; i < FQ * FG; i += 128) {
    const int j = i / FG, k = i % FG;
    float s = 0.f;
    for (int r = 0; r < FR; ++r) s += s_dq[r * FQ + j] * s_n1[r * FG + k];
    part[i] = s;
  }
  if (tid < FG) {
    float sg = 0.f, sb = 0.f, sl = 0.f;
    for (int r = 0; r < FR; ++r) { sg += s_dgm[r * FG + tid]; sb += s_dbt[r * FG + tid]; sl += s_dh0[r * FG + tid]; }
    part[FQ * FG + tid] = sg;
    part[FQ * FG + FG + tid] = sb;
    part[FQ * FG + 2 * FG + FG * Cl + tid] = sl;
  }
  for (int i = tid; i < FG * Cl; i += 128) {
    const int j = i / Cl, k = i % Cl;
    float s = 0.f;
    for (int r = 0; r < FR; ++r) s += s_dh0[r * FG + j] * s_f[r * Cl + k];
    part[FQ * FG + 2 * FG + i] = s;
  }
}

__global__ void dct_a_reduce_kernel(const float* __restrict__ partial, int nblocks, int Cl, float* dWqkv, float* dgm, float* dbt,
                                    float* dWl, float* dbl) {
  const int total = FQ * FG + 2 * FG + FG * Cl + FG;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partial[(long long)b * total + i];
  float* dst;
  if (i < FQ * FG) dst = dWqkv + i;
  else if (i < FQ * FG + FG) dst = dgm + (i - FQ * FG);
  else if (i < FQ * FG + 2 * FG) dst = dbt + (i - FQ * FG - FG);
  else if (i < FQ * FG + 2 * FG + FG * Cl) dst = dWl + (i - FQ * FG - 2 * FG);
  else dst = dbl + (i - FQ * FG - 2 * FG - FG * Cl);
  *dst += s;
}

}  // namespace

extern "C" {

int hdf_dct_a_fwd(const float* F, long long ldf, int Cl, const float* Wl, const float* bl, const float* gm, const float* bt,
                  const float* Wqkv, float* h0, float* n1, float* m1, float* r1, float* qkv, int R, void* stream) {
  HDF_REQUIRE(F && Wl && bl && gm && bt && Wqkv && h0 && n1 && m1 && r1 && qkv && R > 0 && Cl > 0 && Cl % 32 == 0 && Cl <= 512,
              "hdf_dct_a_fwd: bad args (Cl must be a multiple of 32, <= 512)");
  const size_t smem = (size_t)(Cl * FG + FG * FQ + FR * Cl + FR * FG) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(dct_a_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { hdf_set_error("hdf_dct_a_fwd: smem attribute: %s", cudaGetErrorString(e)); return HDF_ERR_CUDA; }
    configured = 200 * 1024;
  }
  DctAParams q{F, ldf, Cl, Wl, bl, gm, bt, Wqkv, h0, n1, m1, r1, qkv, R};
  dct_a_fwd_kernel<<<cdiv(R, FR), 128, smem, (cudaStream_t)stream>>>(q);
  HDF_LAUNCH_CHECK("hdf_dct_a_fwd");
  return HDF_OK;
}

size_t hdf_dct_a_bwd_workspace(int R, int Cl) { return (size_t)cdiv(R, FR) * (FQ * FG + 2 * FG + FG * Cl + FG) * sizeof(float); }

int hdf_dct_a_bwd(const float* dqkv, const float* dh1, const float* h0, const float* n1, const float* m1, const float* r1,
                  const float* F, long long ldf, int Cl, const float* Wqkv, const float* gm, const float* Wl, float* dF,
                  long long lddf, float* dWqkv, float* dgm, float* dbt, float* dWl, float* dbl, int R, void* workspace,
                  size_t ws_bytes, void* stream) {
  HDF_REQUIRE(dqkv && dh1 && h0 && n1 && m1 && r1 && F && Wqkv && gm && Wl && dF && dWqkv && dgm && dbt && dWl && dbl && workspace &&
                  R > 0 && Cl % 32 == 0 && Cl <= 512, "hdf_dct_a_bwd: bad args");
  HDF_REQUIRE(ws_bytes >= hdf_dct_a_bwd_workspace(R, Cl), "hdf_dct_a_bwd: workspace too small");
  const size_t smem = (size_t)(FQ * FG + FG * Cl + FR * FQ + 2 * FR * FG + FR * Cl + 2 * FR * FG) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(dct_a_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { hdf_set_error("hdf_dct_a_bwd: smem attribute: %s", cudaGetErrorString(e)); return HDF_ERR_CUDA; }
    configured = 200 * 1024;
  }
  HDF_REQUIRE(smem <= 200 * 1024, "hdf_dct_a_bwd: Cl=%d needs too much shared memory", Cl);
  const int nb = cdiv(R, FR);
  DctABwdParams q{dqkv, dh1, h0, n1, m1, r1, F, ldf, Cl, Wqkv, gm, Wl, dF, lddf, (float*)workspace, R};
  dct_a_bwd_kernel<<<nb, 128, smem, (cudaStream_t)stream>>>(q);
  HDF_LAUNCH_CHECK("hdf_dct_a_bwd");
  const int total = FQ * FG + 2 * FG + FG * Cl + FG;
  dct_a_reduce_kernel<<<cdiv(total, 128), 128, 0, (cudaStream_t)stream>>>((const float*)workspace, nb, Cl, dWqkv, dgm, dbt, dWl, dbl);
  HDF_LAUNCH_CHECK("hdf_dct_a_bwd/reduce");
  return HDF_OK;
}

}  // extern "C"
