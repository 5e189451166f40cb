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
constexpr int AT = 128;  // queries per block == keys per smem tile

__global__ void __launch_bounds__(AT) attn_fwd_kernel(const float* __restrict__ qkv, long long ld, float* __restrict__ o,
                                                     long long ldo, float* __restrict__ lse, int N, int H, float scale) {
  __shared__ float4 sk[AT], sv[AT];
  const int h = blockIdx.y, b = blockIdx.z;
  const int inner = 4 * H;
  const int i = blockIdx.x * AT + threadIdx.x;
  const float* base = qkv + (long long)b * N * ld;
  float4 q = make_float4(0, 0, 0, 0);
  if (i < N) q = *reinterpret_cast<const float4*>(base + (long long)i * ld + 4 * h);
  q.x *= scale; q.y *= scale; q.z *= scale; q.w *= scale;
  float m = -INFINITY, l = 0.f;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int j0 = 0; j0 < N; j0 += AT) {
    const int j = j0 + threadIdx.x;
    if (j < N) {
      sk[threadIdx.x] = *reinterpret_cast<const float4*>(base + (long long)j * ld + inner + 4 * h);
      sv[threadIdx.x] = *reinterpret_cast<const float4*>(base + (long long)j * ld + 2 * inner + 4 * h);
    }
    __syncthreads();
    const int cnt = min(AT, N - j0);
    for (int t = 0; t < cnt; ++t) {
      const float4 k = sk[t], v = sv[t];
      const float s = q.x * k.x + q.y * k.y + q.z * k.z + q.w * k.w;
      if (s > m) {
        const float r = __expf(m - s);
        l *= r; acc.x *= r; acc.y *= r; acc.z *= r; acc.w *= r;
        m = s;
      }
      const float p = __expf(s - m);
      l += p;
      acc.x = fmaf(p, v.x, acc.x); acc.y = fmaf(p, v.y, acc.y); acc.z = fmaf(p, v.z, acc.z); acc.w = fmaf(p, v.w, acc.w);
    }
    __syncthreads();
  }
  if (i < N) {
    const float inv = 1.f / l;
    *reinterpret_cast<float4*>(o + ((long long)b * N + i) * ldo + 4 * h) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    lse[((long long)b * H + h) * N + i] = m + __logf(l);
  }
}

// dq: thread per query
__global__ void __launch_bounds__(AT) attn_bwd_dq_kernel(const float* __restrict__ qkv, long long ld,
                                                        const float* __restrict__ o, long long ldo,
                                                        const float* __restrict__ dout, long long lddo,
                                                        const float* __restrict__ lse, float* __restrict__ dqkv,
                                                        long long ldg, int N, int H, float scale) {
  __shared__ float4 sk[AT], sv[AT];
  const int h = blockIdx.y, b = blockIdx.z;
  const int inner = 4 * H;
  const int i = blockIdx.x * AT + threadIdx.x;
  const float* base = qkv + (long long)b * N * ld;
  float4 q = make_float4(0, 0, 0, 0), dO = q, O = q;
  float L = 0.f;
  if (i < N) {
    q = *reinterpret_cast<const float4*>(base + (long long)i * ld + 4 * h);
    dO = *reinterpret_cast<const float4*>(dout + ((long long)b * N + i) * lddo + 4 * h);
    O = *reinterpret_cast<const float4*>(o + ((long long)b * N + i) * ldo + 4 * h);
    L = lse[((long long)b * H + h) * N + i];
  }
  const float Di = dO.x * O.x + dO.y * O.y + dO.z * O.z + dO.w * O.w;
  float4 dq = make_float4(0, 0, 0, 0);
  for (int j0 = 0; j0 < N; j0 += AT) {
    const int j = j0 + threadIdx.x;
    if (j < N) {
      sk[threadIdx.x] = *reinterpret_cast<const float4*>(base + (long long)j * ld + inner + 4 * h);
      sv[threadIdx.x] = *reinterpret_cast<const float4*>(base + (long long)j * ld + 2 * inner + 4 * h);
    }
    __syncthreads();
    const int cnt = min(AT, N - j0);
    for (int t = 0; t < cnt; ++t) {
      const float4 k = sk[t], v = sv[t];
      const float s = scale * (q.x * k.x + q.y * k.y + q.z * k.z + q.w * k.w);
      const float p = __expf(s - L);
      const float dp = dO.x * v.x + dO.y * v.y + dO.z * v.z + dO.w * v.w;
      const float ds = p * (dp - Di);
      dq.x = fmaf(ds, k.x, dq.x); dq.y = fmaf(ds, k.y, dq.y); dq.z = fmaf(ds, k.z, dq.z); dq.w = fmaf(ds, k.w, dq.w);
    }
    __syncthreads();
  }
  if (i < N)
    *reinterpret_cast<float4*>(dqkv + ((long long)b * N + i) * ldg + 4 * h) =
        make_float4(dq.x * scale, dq.y * scale, dq.z * scale, dq.w * scale);
}

// dk, dv: thread per key
__global__ void __launch_bounds__(AT) attn_bwd_dkv_kernel(const float* __restrict__ qkv, long long ld,
                                                         const float* __restrict__ o, long long ldo,
                                                         const float* __restrict__ dout, long long lddo,
                                                         const float* __restrict__ lse, float* __restrict__ dqkv,
                                                         long long ldg, int N, int H, float scale) {
  __shared__ float4 sq[AT], sdo[AT];
  __shared__ float sL[AT], sD[AT];
  const int h = blockIdx.y, b = blockIdx.z;
  const int inner = 4 * H;
  const int j = blockIdx.x * AT + threadIdx.x;
  const float* base = qkv + (long long)b * N * ld;
  float4 k = make_float4(0, 0, 0, 0), v = k;
  if (j < N) {
    k = *reinterpret_cast<const float4*>(base + (long long)j * ld + inner + 4 * h);
    v = *reinterpret_cast<const float4*>(base + (long long)j * ld + 2 * inner + 4 * h);
  }
  float4 dk = make_float4(0, 0, 0, 0), dv = dk;
  for (int i0 = 0; i0 < N; i0 += AT) {
    const int i = i0 + threadIdx.x;
    if (i < N) {
      const float4 q = *reinterpret_cast<const float4*>(base + (long long)i * ld + 4 * h);
      const float4 dO = *reinterpret_cast<const float4*>(dout + ((long long)b * N + i) * lddo + 4 * h);
      const float4 O = *reinterpret_cast<const float4*>(o + ((long long)b * N + i) * ldo + 4 * h);
      sq[threadIdx.x] = q;
      sdo[threadIdx.x] = dO;
      sL[threadIdx.x] = lse[((long long)b * H + h) * N + i];
      sD[threadIdx.x] = dO.x * O.x + dO.y * O.y + dO.z * O.z + dO.w * O.w;
    }
    __syncthreads();
    const int cnt = min(AT, N - i0);
    for (int t = 0; t < cnt; ++t) {
      const float4 q = sq[t], dO = sdo[t];
      const float s = scale * (q.x * k.x + q.y * k.y + q.z * k.z + q.w * k.w);
      const float p = __expf(s - sL[t]);
      dv.x = fmaf(p, dO.x, dv.x); dv.y = fmaf(p, dO.y, dv.y); dv.z = fmaf(p, dO.z, dv.z); dv.w = fmaf(p, dO.w, dv.w);
      const float dp = dO.x * v.x + dO.y * v.y + dO.z * v.z + dO.w * v.w;
      const float ds = p * (dp - sD[t]);
      dk.x = fmaf(ds, q.x, dk.x); dk.y = fmaf(ds, q.y, dk.y); dk.z = fmaf(ds, q.z, dk.z); dk.w = fmaf(ds, q.w, dk.w);
    }
    __syncthreads();
  }
  if (j < N) {
    float* g = dqkv + ((long long)b * N + j) * ldg;
    *reinterpret_cast<float4*>(g + inner + 4 * h) = make_float4(dk.x * scale, dk.y * scale, dk.z * scale, dk.w * scale);
    *reinterpret_cast<float4*>(g + 2 * inner + 4 * h) = dv;
  }
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
  dim3 grid(cdiv(N, AT), H, B);
  attn_fwd_kernel<<<grid, AT, 0, (cudaStream_t)stream>>>(qkv, ld, o, ldo, lse, N, H, scale);
  HDF_LAUNCH_CHECK("hdf_attention_fwd");
  return HDF_OK;
}

int hdf_attention_bwd(const float* qkv, long long ld, const float* o, long long ldo, const float* dout, long long lddo,
                      const float* lse, float* dqkv, long long ldg, int B, int N, int H, float scale, void* stream) {
  HDF_REQUIRE(qkv && o && dout && lse && dqkv && (ld % 4 == 0) && (ldo % 4 == 0) && (lddo % 4 == 0) && (ldg % 4 == 0),
              "hdf_attention_bwd: bad args");
  dim3 grid(cdiv(N, AT), H, B);
  attn_bwd_dq_kernel<<<grid, AT, 0, (cudaStream_t)stream>>>(qkv, ld, o, ldo, dout, lddo, lse, dqkv, ldg, N, H, scale);
  HDF_LAUNCH_CHECK("hdf_attention_bwd/dq");
  attn_bwd_dkv_kernel<<<grid, AT, 0, (cudaStream_t)stream>>>(qkv, ld, o, ldo, dout, lddo, lse, dqkv, ldg, N, H, scale);
  HDF_LAUNCH_CHECK("hdf_attention_bwd/dkv");
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
