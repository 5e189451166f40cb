// fp32-accumulate SIMT implicit-GEMM core and the operators built on it:
//   conv3d k3 (s1 / transposed s2 / strided s2) forward + weight-gradient,
//   patch-embedding GEMM (k16 s16 conv), dense token GEMMs with fused epilogues.
// This is the exact-precision path (fp32 / TF32-off parity gate) and the
// generic fallback for shapes the tcgen05 path (tc_conv.cu) does not take.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

// ---------------------------------------------------------------------------
// tile kernel.  AL::M_CONTIG: load4(m,k) returns A(m..m+3,k) else A(m,k..k+3)
//               BL::N_CONTIG: load4(k,n) returns B(k,n..n+3) else B(k..k+3,n)
// grid: x = mtiles*ntiles, y = "group" (tap / batch), z = split-K
// ---------------------------------------------------------------------------
template <class AL, class BL, class EP>
__global__ void __launch_bounds__(NT) gemm_tile_kernel(AL al, BL bl, EP ep, int M, int N, int K, int ntiles_n,
                                                      int k_per_split) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tile_m = blockIdx.x / ntiles_n, tile_n = blockIdx.x % ntiles_n;
  const int m0 = tile_m * BM, n0 = tile_n * BN;
  const int grp = blockIdx.y;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  al.init(grp, m0, tid);
  bl.init(grp, n0, tid);
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // Register prefetch, PF K-chunks deep: the token GEMMs (M ~ 1.5 k rows, K <= 224) run as a handful of CTAs next to the
  // persistent convolution kernels, where every dependent global round trip costs microseconds; with the loads of the
  // next PF chunks in flight a K = 224 product needs ~4 round trips instead of 14.
  constexpr int PF = 4;
  float av[PF][4], bv[PF][4];
  const int nchunks = (kend - kbeg + BK - 1) / BK;
  auto load_ab = [&](int k0, float* a4, float* b4) {
    if (AL::M_CONTIG) al.load4(m0 + (tid % 16) * 4, k0 + tid / 16, kend, a4);
    else al.load4(m0 + tid / 4, k0 + (tid % 4) * 4, kend, a4);
    if (BL::N_CONTIG) bl.load4(k0 + tid / 16, n0 + (tid % 16) * 4, kend, b4);
    else bl.load4(k0 + (tid % 4) * 4, n0 + tid / 4, kend, b4);
  };
#pragma unroll
  for (int p = 0; p < PF; ++p)
    if (p < nchunks) load_ab(kbeg + p * BK, av[p], bv[p]);
  for (int c0 = 0; c0 < nchunks; c0 += PF) {
#pragma unroll
    for (int p = 0; p < PF; ++p) {
      const int c = c0 + p;
      if (c < nchunks) {      // block-uniform
        if (AL::M_CONTIG) {
          *reinterpret_cast<float4*>(&As[tid / 16][(tid % 16) * 4]) = make_float4(av[p][0], av[p][1], av[p][2], av[p][3]);
        } else {
          const int m = tid / 4, k = (tid % 4) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) As[k + i][m] = av[p][i];
        }
        if (BL::N_CONTIG) {
          *reinterpret_cast<float4*>(&Bs[tid / 16][(tid % 16) * 4]) = make_float4(bv[p][0], bv[p][1], bv[p][2], bv[p][3]);
        } else {
          const int n = tid / 4, k = (tid % 4) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) Bs[k + i][n] = bv[p][i];
        }
        __syncthreads();
        if (c + PF < nchunks) load_ab(kbeg + (c + PF) * BK, av[p], bv[p]);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
          const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
          const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
          const float aa[4] = {a.x, a.y, a.z, a.w};
          const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m < M) ep.store(grp, blockIdx.z, m, n0 + tx * 4, acc[i], M, N);
  }
}

// ---------------------------------------------------------------------------
// conv geometry: maps (output voxel, tap) -> input voxel
// ---------------------------------------------------------------------------
struct ConvGeom {
  int N, Do, Ho, Wo, Di, Hi, Wi, mode;
};
// returns false if the tap does not contribute
__device__ __forceinline__ bool conv_in_coord(int mode, int o, int k, int In, int& i) {
  if (mode == 0) {
    i = o + k - 1;
  } else if (mode == 1) {  // transposed s2 p1: o = 2 i - 1 + k
    int t = o + 1 - k;
    if (t & 1) return false;
    i = t >> 1;
  } else {  // strided s2 p1: i = 2 o - 1 + k
    i = 2 * o - 1 + k;
  }
  return i >= 0 && i < In;
}

// A(m = output voxel, k = tap*Cin + ci), k-contiguous
template <typename T>
struct ConvFwdA {
  static constexpr bool M_CONTIG = false;
  const T* x;
  long long ldx;
  int Cin, K, M;
  ConvGeom g;
  int n, d, h, w;
  bool valid;
  __device__ void init(int, int m0, int tid) {
    int m = m0 + tid / 4;
    valid = m < M;
    w = m % g.Wo; m /= g.Wo;
    h = m % g.Ho; m /= g.Ho;
    d = m % g.Do; n = m / g.Do;
  }
  __device__ void load4(int, int k, int kend, float* o) const {
    o[0] = o[1] = o[2] = o[3] = 0.f;
    if (!valid || k >= kend) return;
    if (((Cin | (int)ldx) & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) & (4 * sizeof(T) - 1)) == 0)) {
      const int tap = k / Cin, ci = k - tap * Cin;
      int id, ih, iw;
      if (!conv_in_coord(g.mode, d, tap / 9, g.Di, id)) return;
      if (!conv_in_coord(g.mode, h, (tap / 3) % 3, g.Hi, ih)) return;
      if (!conv_in_coord(g.mode, w, tap % 3, g.Wi, iw)) return;
      const T* p = x + ((((long long)n * g.Di + id) * g.Hi + ih) * g.Wi + iw) * ldx + ci;
      ::load4<T>(p, o);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = k + j;
        if (kk >= kend) break;
        const int tap = kk / Cin, ci = kk - tap * Cin;
        int id, ih, iw;
        if (!conv_in_coord(g.mode, d, tap / 9, g.Di, id)) continue;
        if (!conv_in_coord(g.mode, h, (tap / 3) % 3, g.Hi, ih)) continue;
        if (!conv_in_coord(g.mode, w, tap % 3, g.Wi, iw)) continue;
        o[j] = to_f(x[((((long long)n * g.Di + id) * g.Hi + ih) * g.Wi + iw) * ldx + ci]);
      }
    }
  }
};

// B(k, n) = P[k*N + n] row-major fp32 (packed conv weights, or any [K][N] matrix)
struct RowMajorB {
  static constexpr bool N_CONTIG = true;
  const float* p;
  long long ld;
  int N;
  long long grp_stride;
  const float* base;
  __device__ void init(int grp, int, int) { base = p + grp * grp_stride; }
  __device__ void load4(int k, int n, int kend, float* o) const {
    o[0] = o[1] = o[2] = o[3] = 0.f;
    if (k >= kend) return;
    const float* q = base + (long long)k * ld + n;
    if (n + 3 < N && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) {
      ::load4<float>(q, o);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < N) o[j] = q[j];
    }
  }
};

// B(k, n) = W[n*ld + k]  (k contiguous; torch Linear weight used as x @ W^T)
struct ColMajorB {
  static constexpr bool N_CONTIG = false;
  const float* p;
  long long ld;
  int N;
  __device__ void init(int, int, int) {}
  __device__ void load4(int k, int n, int kend, float* o) const {
    o[0] = o[1] = o[2] = o[3] = 0.f;
    if (n >= N) return;
    const float* q = p + (long long)n * ld + k;
    if (k + 3 < kend && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) {
      ::load4<float>(q, o);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k + j < kend) o[j] = q[j];
    }
  }
};

// A(m, k) = X[m*ld + k] fp32 row-major (k contiguous)
struct RowMajorA {
  static constexpr bool M_CONTIG = false;
  const float* p;
  long long ld;
  int M;
  __device__ void init(int, int, int) {}
  __device__ void load4(int m, int k, int kend, float* o) const {
    o[0] = o[1] = o[2] = o[3] = 0.f;
    if (m >= M) return;
    const float* q = p + (long long)m * ld + k;
    if (k + 3 < kend && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) {
      ::load4<float>(q, o);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k + j < kend) o[j] = q[j];
    }
  }
};

// A(m, k) = X[k*ld + m] fp32 (m contiguous) : used for dW = dY^T X
struct ColMajorA {
  static constexpr bool M_CONTIG = true;
  const float* p;
  long long ld;
  int M;
  __device__ void init(int, int, int) {}
  __device__ void load4(int m, int k, int kend, float* o) const {
    o[0] = o[1] = o[2] = o[3] = 0.f;
    if (k >= kend) return;
    const float* q = p + (long long)k * ld + m;
    if (m + 3 < M && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) {
      ::load4<float>(q, o);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (m + j < M) o[j] = q[j];
    }
  }
};

// wgrad operands.  K index = output voxel.  A(m = ci, k = voxel) = x[in(voxel, tap)][ci]
template <typename T>
struct ConvWgradA {
  static constexpr bool M_CONTIG = true;
  const T* x;
  long long ldx;
  int Cin;
  ConvGeom g;
  int tap;
  __device__ void init(int grp, int, int) { tap = grp; }
  __device__ void load4(int m, int k, int kend, float* o) const {
    o[0] = o[1] = o[2] = o[3] = 0.f;
    if (k >= kend || m >= Cin) return;
    int v = k;
    const int w = v % g.Wo; v /= g.Wo;
    const int h = v % g.Ho; v /= g.Ho;
    const int d = v % g.Do;
    const int n = v / g.Do;
    int id, ih, iw;
    if (!conv_in_coord(g.mode, d, tap / 9, g.Di, id)) return;
    if (!conv_in_coord(g.mode, h, (tap / 3) % 3, g.Hi, ih)) return;
    if (!conv_in_coord(g.mode, w, tap % 3, g.Wi, iw)) return;
    const T* p = x + ((((long long)n * g.Di + id) * g.Hi + ih) * g.Wi + iw) * ldx + m;
    if (m + 3 < Cin && ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & (4 * sizeof(T) - 1)) == 0)) {
      ::load4<T>(p, o);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (m + j < Cin) o[j] = to_f(p[j]);
    }
  }
};
// B(k = voxel, n = co) = dy[voxel][co]
template <typename T>
struct ActRowsB {
  static constexpr bool N_CONTIG = true;
  const T* y;
  long long ldy;
  int C;
  __device__ void init(int, int, int) {}
  __device__ void load4(int k, int n, int kend, float* o) const {
    o[0] = o[1] = o[2] = o[3] = 0.f;
    if (k >= kend || n >= C) return;
    const T* p = y + (long long)k * ldy + n;
    if (n + 3 < C && ((ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & (4 * sizeof(T) - 1)) == 0)) {
      ::load4<T>(p, o);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < C) o[j] = to_f(p[j]);
    }
  }
};

// patch-embed A(m = token (b, pd, ph, pw), k = (kd,kh,kw) in a 16^3 patch) from an NCDHW fp32 image
struct PatchA {
  static constexpr bool M_CONTIG = false;
  const float* img;  // already offset to the modality plane of batch 0
  long long batch_stride;
  int D, H, W, M;
  const float* base;
  bool valid;
  int pdep;          // patch depth: 16 (Conv3d k16) or 1 (Conv2d k16 on a [B, M, 1, H, W] slice: K = 256)
  __device__ void init(int, int m0, int tid) {
    int m = m0 + tid / 4;
    valid = m < M;
    const int w16 = W / 16, h16 = H / 16, d16 = D / pdep;
    const int pw = m % w16; m /= w16;
    const int ph = m % h16; m /= h16;
    const int pd = m % d16;
    const int b = m / d16;
    base = img + b * batch_stride + ((long long)(pd * pdep) * H + ph * 16) * W + pw * 16;
  }
  __device__ void load4(int, int k, int kend, float* o) const {
    o[0] = o[1] = o[2] = o[3] = 0.f;
    if (!valid || k >= kend) return;
    const int kw = k & 15, kh = (k >> 4) & 15, kd = k >> 8;
    ::load4<float>(base + ((long long)kd * H + kh) * W + kw, o);
  }
};
// transposed version for the patch-embed weight gradient: A(m = k-in-patch, k = token)
struct PatchAT {
  static constexpr bool M_CONTIG = true;
  const float* img;
  long long batch_stride;
  int D, H, W;
  int pdep;
  __device__ void init(int, int, int) {}
  __device__ void load4(int m, int k, int kend, float* o) const {
    o[0] = o[1] = o[2] = o[3] = 0.f;
    if (k >= kend) return;
    const int w16 = W / 16, h16 = H / 16, d16 = D / pdep;
    int t = k;
    const int pw = t % w16; t /= w16;
    const int ph = t % h16; t /= h16;
    const int pd = t % d16;
    const int b = t / d16;
    const int kw = m & 15, kh = (m >> 4) & 15, kd = m >> 8;
    ::load4<float>(img + b * batch_stride + ((long long)(pd * pdep + kd) * H + ph * 16 + kh) * W + pw * 16 + kw, o);
  }
};

// ---------------------------------------------------------------------------
// epilogues
// ---------------------------------------------------------------------------
template <typename T>
struct ConvStoreEp {
  T* y;
  long long ldy;
  const float* bias;
  __device__ void store(int, int, int m, int n, const float* acc, int, int N) const {
    T* p = y + (long long)m * ldy + n;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[j] + ((bias && n + j < N) ? bias[n + j] : 0.f);
    if (n + 3 < N && ((ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & (4 * sizeof(T) - 1)) == 0)) {
      store4<T>(p, v);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < N) p[j] = from_f<T>(v[j]);
    }
  }
};

// split-K partials: out[((split*G + grp)*M + m)*N + n]
struct PartialEp {
  float* out;
  int G;
  __device__ void store(int grp, int split, int m, int n, const float* acc, int M, int N) const {
    float* p = out + (((long long)split * G + grp) * M + m) * N + n;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < N) p[j] = acc[j];
  }
};

// token-GEMM epilogue: C = act(acc + bias) * dropout + residual ; optional pre-activation save
struct LinearEp {
  float* c;
  long long ldc;
  const float* bias;       // [N] or null
  const float* residual;   // [M, ldr] or null
  long long ldr;
  float* pre;              // pre-activation (acc + bias) save for backward, [M, N] dense, or null
  int act;                 // 0 none, 1 exact GELU
  float p;                 // dropout prob (0 = off)
  const unsigned long long* seed_ptr;  // device-resident base seed (CUDA-graph replays advance it) or null
  unsigned long long seed;             // offset added to *seed_ptr
  unsigned call_id;
  float beta;              // c = beta*c + result   (0 or 1)
  __device__ void store(int, int, int m, int n, const float* acc, int, int N) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j >= N) break;
      float v = acc[j] + (bias ? bias[n + j] : 0.f);
      if (pre) pre[(long long)m * N + n + j] = v;
      if (act == 1) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
      if (p > 0.f) v *= hdf_dropout_scale(seed + (seed_ptr ? *seed_ptr : 0ull), call_id, (unsigned long long)m * N + n + j, p);
      if (residual) v += residual[(long long)m * ldr + n + j];
      float* q = c + (long long)m * ldc + n + j;
      *q = (beta != 0.f) ? (*q + v) : v;
    }
  }
};

template <class AL, class BL, class EP>
int launch_gemm(AL al, BL bl, EP ep, int M, int N, int K, int groups, int nsplit, cudaStream_t s, const char* name) {
  if (M <= 0 || N <= 0 || K <= 0) return HDF_OK;
  const int tm = cdiv(M, BM), tn = cdiv(N, BN);
  int kps = cdiv(K, nsplit);
  kps = cdiv(kps, BK) * BK;
  nsplit = cdiv(K, kps);
  dim3 grid(tm * tn, groups, nsplit);
  gemm_tile_kernel<AL, BL, EP><<<grid, NT, 0, s>>>(al, bl, ep, M, N, K, tn, kps);
  HDF_LAUNCH_CHECK(name);
  return HDF_OK;
}

int conv_geom(int mode, int N, int Do, int Ho, int Wo, ConvGeom& g) {
  g.N = N; g.Do = Do; g.Ho = Ho; g.Wo = Wo; g.mode = mode;
  if (mode == 0) { g.Di = Do; g.Hi = Ho; g.Wi = Wo; }
  else if (mode == 1) {
    if ((Do | Ho | Wo) & 1) return -1;
    g.Di = Do / 2; g.Hi = Ho / 2; g.Wi = Wo / 2;
  } else if (mode == 2) { g.Di = Do * 2; g.Hi = Ho * 2; g.Wi = Wo * 2; }
  else return -1;
  return 0;
}

// weight (re)packing:  out[tap][a][b] = w[a*sa + b*sb + tap_src], tap_src = flip ? 26 - tap : tap
__global__ void pack_w_kernel(const float* __restrict__ w, float* __restrict__ out, int A, int B, long long sa,
                              long long sb, int flip) {
  const long long total = 27ll * A * B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = i % B;
    const int a = (i / B) % A;
    const int tap = i / ((long long)A * B);
    out[i] = w[a * sa + b * sb + (flip ? 26 - tap : tap)];
  }
}

// reduce split-K partials [S][27][A][B] -> torch-layout gradient g[a*sa + b*sb + tap] (+= if accumulate)
__global__ void reduce_wgrad_kernel(const float* __restrict__ part, float* __restrict__ g, int S, int A, int B,
                                    long long sa, long long sb, int accumulate) {
  const long long per = 27ll * A * B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < S; ++z) s += part[z * per + i];
    const int b = i % B;
    const int a = (i / B) % A;
    const int tap = i / ((long long)A * B);
    float* q = g + a * sa + b * sb + tap;
    *q = accumulate ? (*q + s) : s;
  }
}

// reduce generic partials [S][count] -> out[count]
__global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, int S, long long count,
                                       int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < S; ++z) s += part[z * count + i];
    out[i] = accumulate ? out[i] + s : s;
  }
}

// partials [S][R][C] -> out[c*R + r]
__global__ void reduce_transpose_kernel(const float* __restrict__ part, float* __restrict__ out, int S, int R, int C,
                                        int accumulate) {
  const long long cnt = (long long)R * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cnt; i += (long long)gridDim.x * blockDim.x) {
    const int r = i % R, c = i / R;  // consecutive threads -> consecutive r (coalesced write)
    float s = 0.f;
    for (int z = 0; z < S; ++z) s += part[z * cnt + (long long)r * C + c];
    out[i] = accumulate ? out[i] + s : s;
  }
}

// tok[m, e] = (sum_s part[s][m][e] + bias[e] + pos[m % ntok, e]) * dropout
__global__ void patch_finish_kernel(const float* __restrict__ part, int S, float* __restrict__ tok, long long ld,
                                    const float* __restrict__ bias, const float* __restrict__ pos, int ntok, int E,
                                    long long total, float p, const unsigned long long* seed_ptr, unsigned long long seed,
                                    unsigned call_id) {
  if (seed_ptr) seed += *seed_ptr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = i % E;
    const long long m = i / E;
    float v = (bias ? bias[e] : 0.f) + pos[(m % ntok) * E + e];
    for (int z = 0; z < S; ++z) v += part[z * total + i];
    tok[m * ld + e] = v * hdf_dropout_scale(seed, call_id, (unsigned long long)i, p);
  }
}

// tok[m, e] = (tok[m, e] + pos[m % ntok, e]) * dropout
__global__ void posemb_dropout_kernel(float* __restrict__ tok, long long ld, const float* __restrict__ pos, int ntok, int E,
                                      long long total, float p, const unsigned long long* seed_ptr,
                                      unsigned long long seed, unsigned call_id) {
  if (seed_ptr) seed += *seed_ptr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = i % E;
    const long long m = i / E;
    float v = tok[m * ld + e] + pos[(m % ntok) * E + e];
    tok[m * ld + e] = v * hdf_dropout_scale(seed, call_id, (unsigned long long)i, p);
  }
}

int wgrad_splits(long long K, int tiles_total) {
  // aim for ~4 waves of 148 SMs, at least 2048 voxels per split
  long long want = (4 * 148 + tiles_total - 1) / tiles_total;
  long long maxs = K / 2048 > 0 ? K / 2048 : 1;
  long long s = want < maxs ? want : maxs;
  return (int)(s < 1 ? 1 : s);
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int hdf_conv_pack_weights(const float* w, float* packed, int A, int B, long long stride_a, long long stride_b,
                          int flip, void* stream) {
  HDF_REQUIRE(w && packed && A > 0 && B > 0, "hdf_conv_pack_weights: bad args");
  const long long total = 27ll * A * B;
  pack_w_kernel<<<min(1024, cdiv(total, 256)), 256, 0, (cudaStream_t)stream>>>(w, packed, A, B, stride_a, stride_b, flip);
  HDF_LAUNCH_CHECK("hdf_conv_pack_weights");
  return HDF_OK;
}

int hdf_conv3d_fwd(int dtype, int mode, const void* x, long long ldx, const float* w_packed, const float* bias, void* y,
                   long long ldy, int N, int Do, int Ho, int Wo, int Cin, int Cout, void* stream) {
  ConvGeom g;
  HDF_REQUIRE(conv_geom(mode, N, Do, Ho, Wo, g) == 0, "hdf_conv3d_fwd: bad mode/dims (mode %d, %dx%dx%d)", mode, Do, Ho, Wo);
  HDF_REQUIRE(x && w_packed && y && Cin > 0 && Cout > 0 && ldx >= Cin && ldy >= Cout, "hdf_conv3d_fwd: bad args");
  const long long M = (long long)N * Do * Ho * Wo;
  HDF_REQUIRE(M < (1ll << 31), "hdf_conv3d_fwd: too many voxels");
  HDF_DISPATCH_DTYPE(dtype, T, {
    ConvFwdA<T> al{(const T*)x, ldx, Cin, 27 * Cin, (int)M, g};
    RowMajorB bl{w_packed, Cout, Cout, 0, nullptr};
    ConvStoreEp<T> ep{(T*)y, ldy, bias};
    return launch_gemm(al, bl, ep, (int)M, Cout, 27 * Cin, 1, 1, (cudaStream_t)stream, "hdf_conv3d_fwd");
  });
}

size_t hdf_conv3d_wgrad_workspace(int N, int Do, int Ho, int Wo, int Cin, int Cout) {
  const long long K = (long long)N * Do * Ho * Wo;
  const int tiles = cdiv(Cin, BM) * cdiv(Cout, BN) * 27;
  const int S = wgrad_splits(K, tiles);
  return (size_t)S * 27 * Cin * Cout * sizeof(float);
}

// dw_torch[a*stride_a + b*stride_b + tap] with a = ci (x channels), b = co (dy channels)
int hdf_conv3d_wgrad(int dtype, int mode, const void* x, long long ldx, const void* dy, long long ldy, float* dw,
                     long long stride_ci, long long stride_co, int N, int Do, int Ho, int Wo, int Cin, int Cout,
                     void* workspace, size_t ws_bytes, int accumulate, void* stream) {
  ConvGeom g;
  HDF_REQUIRE(conv_geom(mode, N, Do, Ho, Wo, g) == 0, "hdf_conv3d_wgrad: bad mode/dims");
  HDF_REQUIRE(x && dy && dw && workspace, "hdf_conv3d_wgrad: null pointer");
  const long long K = (long long)N * Do * Ho * Wo;
  HDF_REQUIRE(K < (1ll << 31), "hdf_conv3d_wgrad: too many voxels");
  const int tiles = cdiv(Cin, BM) * cdiv(Cout, BN) * 27;
  int S = wgrad_splits(K, tiles);
  HDF_REQUIRE(ws_bytes >= (size_t)S * 27 * Cin * Cout * sizeof(float), "hdf_conv3d_wgrad: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  int kps = cdiv(cdiv(K, S), BK) * BK;
  S = cdiv(K, kps);
  HDF_DISPATCH_DTYPE(dtype, T, {
    ConvWgradA<T> al{(const T*)x, ldx, Cin, g, 0};
    ActRowsB<T> bl{(const T*)dy, ldy, Cout};
    PartialEp ep{(float*)workspace, 27};
    int rc = launch_gemm(al, bl, ep, Cin, Cout, (int)K, 27, S, s, "hdf_conv3d_wgrad");
    if (rc) return rc;
  });
  const long long per = 27ll * Cin * Cout;
  reduce_wgrad_kernel<<<min(2048, cdiv(per, 256)), 256, 0, s>>>((const float*)workspace, dw, S, Cin, Cout, stride_ci,
                                                               stride_co, accumulate);
  HDF_LAUNCH_CHECK("hdf_conv3d_wgrad/reduce");
  return HDF_OK;
}

// D == 1 selects the 2-D patch embedding (Conv2d k16 s16 of models/HDenseFormer_2D.py:112-115): patch depth 1, K = 256
static inline int patch_depth(int D) { return D == 1 ? 1 : 16; }

size_t hdf_patch_embed_fwd_workspace(int B, int D, int H, int W, int E) {
  const long long M = (long long)B * (D / patch_depth(D)) * (H / 16) * (W / 16);
  return (size_t)8 * M * E * sizeof(float);
}

// tokens[b, n, :] (ld = ldo) = patch_conv(img[b, modality]) + bias + pos[n, :], then dropout
// img: NCDHW fp32 [B, Mch, D, H, W]; weight [E, 4096] (torch [E,1,16,16,16]); out fp32.
// K = 4096 is split 8 ways (the GEMM has only ~46 output tiles and sits at the head of the transformer branch's
// dependency chain); the finishing kernel sums the partials in a fixed order and applies bias + pos-emb + dropout.
int hdf_patch_embed_fwd(const float* img, int B, int Mch, int modality, int D, int H, int W, const float* weight,
                        const float* bias, const float* pos, float* out, long long ldo, int E, float p,
                        const unsigned long long* seed_ptr, unsigned long long seed, unsigned call_id, void* workspace,
                        size_t ws_bytes, void* stream) {
  const int pdep = patch_depth(D), K = pdep * 256;
  HDF_REQUIRE(img && weight && out && workspace && (D % pdep == 0) && (H % 16 == 0) && (W % 16 == 0) && (W % 4 == 0),
              "hdf_patch_embed_fwd: bad args (every spatial dim must be a multiple of 16; depth 1 = 2-D patches)");
  const int ntok = (D / pdep) * (H / 16) * (W / 16);
  const int M = B * ntok;
  HDF_REQUIRE(ws_bytes >= (size_t)8 * M * E * sizeof(float), "hdf_patch_embed_fwd: workspace too small");
  PatchA al{img + (long long)modality * D * H * W, (long long)Mch * D * H * W, D, H, W, M, nullptr, false, pdep};
  ColMajorB bl{weight, K, E};
  PartialEp ep{(float*)workspace, 1};
  int rc = launch_gemm(al, bl, ep, M, E, K, 1, 8, (cudaStream_t)stream, "hdf_patch_embed_fwd");
  if (rc) return rc;
  const long long total = (long long)M * E;
  patch_finish_kernel<<<min(1024, cdiv(total, 256)), 256, 0, (cudaStream_t)stream>>>((const float*)workspace, 8, out, ldo, bias,
                                                                                    pos, ntok, E, total, p, seed_ptr, seed, call_id);
  HDF_LAUNCH_CHECK("hdf_patch_embed_fwd/finish");
  return HDF_OK;
}

size_t hdf_patch_embed_wgrad_workspace(int B, int D, int H, int W, int E) {
  const int pk = patch_depth(D) * 256;
  const long long K = (long long)B * (D / patch_depth(D)) * (H / 16) * (W / 16);
  const int tiles = cdiv(pk, BM) * cdiv(E, BN);
  int S = (int)((2 * 148 + tiles - 1) / tiles);
  if (S > K) S = (int)K;
  if (S < 1) S = 1;
  return (size_t)S * pk * E * sizeof(float);
}

// dweight[E,4096] (+)= dtok^T @ patches ; dtok [B*ntok, ldd] fp32 (already multiplied by the dropout mask)
int hdf_patch_embed_wgrad(const float* img, int B, int Mch, int modality, int D, int H, int W, const float* dtok,
                          long long ldd, float* dweight, int E, void* workspace, size_t ws_bytes, int accumulate,
                          void* stream) {
  HDF_REQUIRE(img && dtok && dweight && workspace, "hdf_patch_embed_wgrad: null pointer");
  const int pdep = patch_depth(D), pk = pdep * 256;
  const int ntok = (D / pdep) * (H / 16) * (W / 16);
  const int K = B * ntok;
  const int tiles = cdiv(pk, BM) * cdiv(E, BN);
  int S = (2 * 148 + tiles - 1) / tiles;
  if (S > K) S = K;
  if (S < 1) S = 1;
  HDF_REQUIRE(ws_bytes >= (size_t)S * pk * E * sizeof(float), "hdf_patch_embed_wgrad: workspace too small");
  int kps = cdiv(cdiv(K, S), BK) * BK;
  S = cdiv(K, kps);
  // computes P[m = k-in-patch][n = e]; stored transposed into dweight[e][m] by the reducer
  PatchAT al{img + (long long)modality * D * H * W, (long long)Mch * D * H * W, D, H, W, pdep};
  ActRowsB<float> bl{dtok, ldd, E};
  PartialEp ep{(float*)workspace, 1};
  int rc = launch_gemm(al, bl, ep, pk, E, K, 1, S, (cudaStream_t)stream, "hdf_patch_embed_wgrad");
  if (rc) return rc;
  const long long cnt = (long long)pk * E;
  reduce_transpose_kernel<<<min(1024, cdiv(cnt, 256)), 256, 0, (cudaStream_t)stream>>>((const float*)workspace, dweight, S,
                                                                                      pk, E, accumulate);
  HDF_LAUNCH_CHECK("hdf_patch_embed_wgrad/reduce");
  return HDF_OK;
}

// C[M,N] (ldc) = epilogue(A[M,K] (lda) @ op(B)), op(B) = B^T with B [N,K] (ldb) if b_is_nk else B [K,N] (ldb)
int hdf_gemm_rowmajor(const float* A, long long lda, const float* Bm, long long ldb, int b_is_nk, float* C, long long ldc,
                      int M, int N, int K, const float* bias, const float* residual, long long ldr, float* pre, int act,
                      float p, const unsigned long long* seed_ptr, unsigned long long seed, unsigned call_id, int accumulate,
                      void* stream) {
  HDF_REQUIRE(A && Bm && C, "hdf_gemm_rowmajor: null pointer");
  RowMajorA al{A, lda, M};
  LinearEp ep{C, ldc, bias, residual, ldr, pre, act, p, seed_ptr, seed, call_id, accumulate ? 1.f : 0.f};
  if (b_is_nk) {
    ColMajorB bl{Bm, ldb, N};
    return launch_gemm(al, bl, ep, M, N, K, 1, 1, (cudaStream_t)stream, "hdf_gemm_rowmajor(nk)");
  }
  RowMajorB bl{Bm, ldb, N, 0, nullptr};
  return launch_gemm(al, bl, ep, M, N, K, 1, 1, (cudaStream_t)stream, "hdf_gemm_rowmajor(kn)");
}

size_t hdf_gemm_at_b_workspace(int M, int N, int K) {
  const int tiles = cdiv(M, BM) * cdiv(N, BN);
  int S = (148 + tiles - 1) / tiles;
  if (S > cdiv(K, 64)) S = cdiv(K, 64);
  if (S < 1) S = 1;
  return (size_t)S * M * N * sizeof(float);
}

// C[M,N] (dense) (+)= A^T @ B with A [K,M] (lda), B [K,N] (ldb)  -- Linear weight gradient dW = dY^T X
int hdf_gemm_at_b(const float* A, long long lda, const float* Bm, long long ldb, float* C, int M, int N, int K,
                  void* workspace, size_t ws_bytes, int accumulate, void* stream) {
  HDF_REQUIRE(A && Bm && C && workspace, "hdf_gemm_at_b: null pointer");
  const int tiles = cdiv(M, BM) * cdiv(N, BN);
  int S = (148 + tiles - 1) / tiles;
  if (S > cdiv(K, 64)) S = cdiv(K, 64);
  if (S < 1) S = 1;
  HDF_REQUIRE(ws_bytes >= (size_t)S * M * N * sizeof(float), "hdf_gemm_at_b: workspace too small");
  int kps = cdiv(cdiv(K, S), BK) * BK;
  S = cdiv(K, kps);
  ColMajorA al{A, lda, M};
  ActRowsB<float> bl{Bm, ldb, N};
  PartialEp ep{(float*)workspace, 1};
  int rc = launch_gemm(al, bl, ep, M, N, K, 1, S, (cudaStream_t)stream, "hdf_gemm_at_b");
  if (rc) return rc;
  const long long cnt = (long long)M * N;
  reduce_partials_kernel<<<min(1024, cdiv(cnt, 256)), 256, 0, (cudaStream_t)stream>>>((const float*)workspace, C, S, cnt,
                                                                                     accumulate);
  HDF_LAUNCH_CHECK("hdf_gemm_at_b/reduce");
  return HDF_OK;
}

}  // extern "C"

// C++-linkage helper for patch_tc.cu: the split-K finishing pass of the patch embedding
int hdf_patch_finish(const float* part, int S, float* out, long long ldo, const float* bias, const float* pos, int ntok, int E,
                     long long total, float p, const unsigned long long* seed_ptr, unsigned long long seed, unsigned call_id,
                     void* stream) {
  patch_finish_kernel<<<min(1024, cdiv(total, 256)), 256, 0, (cudaStream_t)stream>>>(part, S, out, ldo, bias, pos, ntok, E, total, p,
                                                                                    seed_ptr, seed, call_id);
  HDF_LAUNCH_CHECK("hdf_patch_embed_tc_fwd/finish");
  return HDF_OK;
}
