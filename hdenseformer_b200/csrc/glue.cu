// Bandwidth-bound glue on channels-last activations [N, V=D*H*W, C] (channel stride "ld" so every
// operand can be a channel slice of a pre-allocated concat buffer: SURVEY 2.1 K10 "no cat kernel"):
// InstanceNorm statistics / apply(+ReLU)(+residual) / backward, 2x2x2 max-pool, trilinear x2,
// 1x1x1 heads, layout conversion, axpy, column sums.  128-bit vectorised accesses, fp32 math,
// fp64 cross-thread reduction of statistics.
#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

// Ticket counters for "the last CTA finalises" reductions: 64 zero-initialised counters per (device, stream), handed out on
// first use.  Launches on one stream are ordered, and the kernels leave every counter at zero again, so a slot is never
// shared by two reductions in flight -- the same contract as the per-stream partial-sum workspace the callers pass in.  The
// pool itself is allocated on the first call OUTSIDE stream capture (cudaMalloc is not capturable); until then, and when the
// pool is exhausted, nullptr is returned and the callers launch their separate finalising kernels.
// MEASURED SLOWER, therefore opt-in (HDF_FUSED_FINALIZE=1): 95 fewer launches per step (769 -> 674) but 21.65 vs 20.17 ms --
// the last CTA walks chunks x 2C partials with 8 warps and ~10 dependent L2 round trips per channel (~24 us), where the
// stand-alone finalising kernel spreads the same walk over N*C warps (~5 us).  profiles/r2_fused_finalize_ab.txt.
unsigned* hdf_ticket_slot(void* stream) {
  static std::mutex mu;
  static std::map<std::pair<int, void*>, unsigned*> slots;
  static std::map<int, std::pair<unsigned*, int>> pools;      // device -> (base, used)
  constexpr int kSlots = 512, kPer = 64;
  static const bool on = getenv("HDF_FUSED_FINALIZE") != nullptr;
  if (!on) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  auto it = slots.find({dev, stream});
  if (it != slots.end()) return it->second;
  auto& pool = pools[dev];
  if (!pool.first) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing((cudaStream_t)stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
      cudaGetLastError();
      return nullptr;
    }
    unsigned* base = nullptr;
    if (cudaMalloc(&base, (size_t)kSlots * kPer * sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaMemset(base, 0, (size_t)kSlots * kPer * sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    pool = {base, 0};
  }
  if (pool.second >= kSlots) return nullptr;
  unsigned* p = pool.first + (size_t)pool.second * kPer;
  ++pool.second;
  slots[{dev, stream}] = p;
  return p;
}
constexpr int HDF_TICKETS_PER_SLOT = 64;

namespace {

template <typename T> struct VecOf { static constexpr int value = 4; };
template <> struct VecOf<bf16> { static constexpr int value = 8; };

template <typename T, int VEC> __device__ __forceinline__ void loadv(const T* p, float* o) {
  if (VEC == 8) load8<T>(p, o);
  else if (VEC == 4) load4<T>(p, o);
  else o[0] = to_f(p[0]);
}
template <typename T, int VEC> __device__ __forceinline__ void storev(T* p, const float* o) {
  if (VEC == 8) store8<T>(p, o);
  else if (VEC == 4) store4<T>(p, o);
  else p[0] = from_f<T>(o[0]);
}

template <typename T>
bool can_vec(const void* p, long long ld, int C) {
  const int v = VecOf<T>::value;
  return (C % v == 0) && (ld % v == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
}

// ---------------------------------------------------------------------------
// generic per-(n,c) two-quantity reduction over V rows.
// grid (chunks, N), block 256.  partial[n][chunk][2][C] (double)
// ---------------------------------------------------------------------------
// What the LAST CTA of a sample does with the per-chunk partials (same summation order as the stand-alone finalising
// kernels below: lanes stride over the chunks, fixed-order shuffle reduction -> bit-identical results).
struct FinNone {
  static constexpr int kind = 0;
  __device__ void apply(int, int, int, double, double) const {}
};
struct FinStats {            // mean / rstd of InstanceNorm
  static constexpr int kind = 1;
  long long V; float eps; float* mean; float* rstd;
  __device__ void apply(int n, int c, int C, double s, double q) const {
    const double m = s / (double)V;
    double var = q / (double)V - m * m;
    if (var < 0.0) var = 0.0;
    mean[n * C + c] = (float)m;
    rstd[n * C + c] = (float)(1.0 / sqrt(var + (double)eps));
  }
};
struct FinSums {             // s1 / s2 of the InstanceNorm backward (+ gamma / beta gradients once every sample is done)
  static constexpr int kind = 2;
  float* s1; float* s2; float* dgamma; float* dbeta; int accumulate;
  __device__ void apply(int n, int c, int C, double s, double q) const {
    s1[n * C + c] = (float)s;
    if (s2) s2[n * C + c] = (float)q;
  }
};
struct FinColSum {           // out[c] (+)= column sum (N = 1)
  static constexpr int kind = 3;
  float* out; int accumulate;
  __device__ void apply(int, int c, int, double s, double) const { out[c] = accumulate ? out[c] + (float)s : (float)s; }
};

template <typename T, int VEC, class F, class FIN>
__global__ void __launch_bounds__(256) rowreduce_kernel(F f, int V, int C, int rows_per_chunk, double* __restrict__ partial, FIN fin,
                                                        unsigned* __restrict__ tickets) {
  extern __shared__ double smd[];  // [2][rows_per_iter][C]
  const int cpv = C / VEC;
  const int rpi = 256 / cpv;  // rows per iteration (>=1 guaranteed by host)
  const int tid = threadIdx.x;
  const int cg = tid % cpv, r = tid / cpv;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int v0 = chunk * rows_per_chunk;
  const int v1 = min(V, v0 + rows_per_chunk);
  float sa[VEC], sb[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) sa[i] = sb[i] = 0.f;
  if (r < rpi) {
    // the thread's channel group is fixed: per-channel constants live in registers, and four independent rows are
    // in flight per iteration (a single 16-byte load per thread cannot cover the HBM latency on B200)
    typename F::template Coef<VEC> k;
    f.template prep<VEC>(n, cg * VEC, k);
    int v = v0 + r;
    // (fp32 tensors keep the plain sequential order: that path is the bit-exact-argmax reference path, not the fast one)
    for (; sizeof(T) == 2 && v + 3 * rpi < v1; v += 4 * rpi) {
      float a[4][VEC], b[4][VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u) f.template eval<VEC>(n, v + u * rpi, cg * VEC, k, a[u], b[u]);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        sa[i] += (a[0][i] + a[1][i]) + (a[2][i] + a[3][i]);
        sb[i] += (b[0][i] + b[1][i]) + (b[2][i] + b[3][i]);
      }
    }
    for (; v < v1; v += rpi) {
      float a[VEC], b[VEC];
      f.template eval<VEC>(n, v, cg * VEC, k, a, b);
#pragma unroll
      for (int i = 0; i < VEC; ++i) { sa[i] += a[i]; sb[i] += b[i]; }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      smd[(0 * rpi + r) * C + cg * VEC + i] = (double)sa[i];
      smd[(1 * rpi + r) * C + cg * VEC + i] = (double)sb[i];
    }
  }
  __syncthreads();
  for (int i = tid; i < 2 * C; i += 256) {
    const int q = i / C, c = i % C;
    double s = 0.0;
    for (int rr = 0; rr < rpi; ++rr) s += smd[(q * rpi + rr) * C + c];
    partial[(((long long)n * gridDim.x + chunk) * 2 + q) * C + c] = s;
  }
  if (FIN::kind == 0) return;
  // ---- the last CTA of sample n to get here finalises it (threadFenceReduction pattern)
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(&tickets[n], 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int chunks = gridDim.x, warp = tid >> 5, lane = tid & 31;
  for (int c = warp; c < C; c += 8) {
    double s = 0.0, q = 0.0;
    for (int k = lane; k < chunks; k += 32) {
      s += __ldcg(partial + (((long long)n * chunks + k) * 2 + 0) * C + c);
      q += __ldcg(partial + (((long long)n * chunks + k) * 2 + 1) * C + c);
    }
    s = warp_sum_d(s);
    q = warp_sum_d(q);
    if (lane == 0) fin.apply(n, c, C, s, q);
  }
  if (tid == 0) tickets[n] = 0;                       // leave the counter ready for the next launch on this stream
  if (FIN::kind == 2) {
    const FinSums& fs = reinterpret_cast<const FinSums&>(fin);
    if (fs.dgamma == nullptr && fs.dbeta == nullptr) return;
    // gamma / beta gradients sum s2 / s1 over the samples: done by the CTA that finalises the last sample
    const int N = gridDim.y;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&tickets[HDF_TICKETS_PER_SLOT - 1], 1u) == (unsigned)N - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int c = tid; c < C; c += 256) {
      float a = 0.f, b = 0.f;
      for (int m = 0; m < N; ++m) { a += __ldcg(fs.s2 + m * C + c); b += __ldcg(fs.s1 + m * C + c); }
      if (fs.dgamma) fs.dgamma[c] = fs.accumulate ? fs.dgamma[c] + a : a;
      if (fs.dbeta) fs.dbeta[c] = fs.accumulate ? fs.dbeta[c] + b : b;
    }
    if (tid == 0) tickets[HDF_TICKETS_PER_SLOT - 1] = 0;
  }
}

struct NoCoef {};

template <typename T>
struct StatsF {
  const T* y; long long ld; long long V;
  template <int VEC> using Coef = NoCoef;
  template <int VEC> __device__ void prep(int, int, NoCoef&) const {}
  template <int VEC> __device__ void eval(int n, int v, int c, const NoCoef&, float* a, float* b) const {
    loadv<T, VEC>(y + ((long long)n * V + v) * ld + c, a);
#pragma unroll
    for (int i = 0; i < VEC; ++i) b[i] = a[i] * a[i];
  }
};

template <int VEC> struct InBwdCoef { float m[VEC], rs[VEC], gm[VEC], bt[VEC]; };

template <typename T>
struct InBwdF {
  const T* dout; long long ldd; const T* y; long long ldy; long long V;
  const float* mean; const float* rstd; const float* gamma; const float* beta; int C; int relu;
  template <int VEC> using Coef = InBwdCoef<VEC>;
  template <int VEC> __device__ void prep(int n, int c, InBwdCoef<VEC>& k) const {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      k.m[i] = mean[n * C + c + i];
      k.rs[i] = rstd[n * C + c + i];
      k.gm[i] = gamma ? gamma[c + i] : 1.f;
      k.bt[i] = beta ? beta[c + i] : 0.f;
    }
  }
  template <int VEC> __device__ void eval(int n, int v, int c, const InBwdCoef<VEC>& k, float* a, float* b) const {
    float g[VEC], x[VEC];
    loadv<T, VEC>(dout + ((long long)n * V + v) * ldd + c, g);
    loadv<T, VEC>(y + ((long long)n * V + v) * ldy + c, x);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float xh = (x[i] - k.m[i]) * k.rs[i];
      const float z = xh * k.gm[i] + k.bt[i];
      const float dz = (!relu || z > 0.f) ? g[i] : 0.f;
      a[i] = dz;
      b[i] = dz * xh;
    }
  }
};

template <typename T>
struct ColSumF {
  const T* x; long long ld; long long V;
  template <int VEC> using Coef = NoCoef;
  template <int VEC> __device__ void prep(int, int, NoCoef&) const {}
  template <int VEC> __device__ void eval(int n, int v, int c, const NoCoef&, float* a, float* b) const {
    loadv<T, VEC>(x + ((long long)n * V + v) * ld + c, a);
#pragma unroll
    for (int i = 0; i < VEC; ++i) b[i] = 0.f;
  }
};

// one warp per (n, c): lanes stride over the chunk partials, fixed-order shuffle reduction (deterministic)
__global__ void stats_finalize_kernel(const double* __restrict__ partial, int chunks, int C, long long V, float eps,
                                      float* __restrict__ mean, float* __restrict__ rstd, int NC) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (i >= NC) return;
  const int n = i / C, c = i % C;
  double s = 0.0, q = 0.0;
  for (int k = lane; k < chunks; k += 32) {
    s += partial[(((long long)n * chunks + k) * 2 + 0) * C + c];
    q += partial[(((long long)n * chunks + k) * 2 + 1) * C + c];
  }
  s = warp_sum_d(s);
  q = warp_sum_d(q);
  if (lane != 0) return;
  const double m = s / (double)V;
  double var = q / (double)V - m * m;
  if (var < 0.0) var = 0.0;
  mean[i] = (float)m;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void sums_finalize_kernel(const double* __restrict__ partial, int chunks, int C, float* __restrict__ s1,
                                     float* __restrict__ s2, int NC) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (i >= NC) return;
  const int n = i / C, c = i % C;
  double s = 0.0, q = 0.0;
  for (int k = lane; k < chunks; k += 32) {
    s += partial[(((long long)n * chunks + k) * 2 + 0) * C + c];
    q += partial[(((long long)n * chunks + k) * 2 + 1) * C + c];
  }
  s = warp_sum_d(s);
  q = warp_sum_d(q);
  if (lane != 0) return;
  s1[i] = (float)s;
  if (s2) s2[i] = (float)q;
}

__global__ void colsum_finalize_kernel(const double* __restrict__ partial, int chunks, int C, float* __restrict__ out,
                                       int accumulate) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (c >= C) return;
  double s = 0.0;
  for (int k = lane; k < chunks; k += 32) s += partial[((long long)k * 2 + 0) * C + c];
  s = warp_sum_d(s);
  if (lane != 0) return;
  out[c] = accumulate ? out[c] + (float)s : (float)s;
}

__global__ void in_param_grads_kernel(const float* __restrict__ s1, const float* __restrict__ s2, int N, int C,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int n = 0; n < N; ++n) { a += s2[n * C + c]; b += s1[n * C + c]; }
  if (dgamma) dgamma[c] = accumulate ? dgamma[c] + a : a;
  if (dbeta) dbeta[c] = accumulate ? dbeta[c] + b : b;
}

int reduce_chunks(long long V, int N) {
  // ~4 waves of CTAs, at least 128 rows per chunk (token tensors have only ~1.5 k rows)
  long long c = (4 * 148 + N - 1) / N;
  long long mx = V / 128 > 0 ? V / 128 : 1;
  if (c > mx) c = mx;
  return (int)(c < 1 ? 1 : c);
}

template <typename T, class F, class FIN>
int launch_rowreduce(F f, const void* p0, long long ld0, const void* p1, long long ld1, int N, long long V, int C,
                     double* partial, int& chunks, cudaStream_t s, const char* name, FIN fin, unsigned* tickets) {
  chunks = reduce_chunks(V, N);
  const int rpc = cdiv(V, chunks);
  chunks = cdiv(V, rpc);
  dim3 grid(chunks, N);
  const bool vec = can_vec<T>(p0, ld0, C) && (p1 == nullptr || can_vec<T>(p1, ld1, C)) && (C / VecOf<T>::value) <= 256;
  if (vec) {
    constexpr int VEC = VecOf<T>::value;
    const int rpi = 256 / (C / VEC);
    const size_t smem = (size_t)2 * rpi * C * sizeof(double);
    rowreduce_kernel<T, VEC, F, FIN><<<grid, 256, smem, s>>>(f, (int)V, C, rpc, partial, fin, tickets);
  } else {
    HDF_REQUIRE(C <= 256, "%s: scalar path supports C <= 256 (C=%d)", name, C);
    const int rpi = 256 / C;
    const size_t smem = (size_t)2 * rpi * C * sizeof(double);
    rowreduce_kernel<T, 1, F, FIN><<<grid, 256, smem, s>>>(f, (int)V, C, rpc, partial, fin, tickets);
  }
  HDF_LAUNCH_CHECK(name);
  return HDF_OK;
}

// ticket slot for a fused finalise, or null (N too large for the slot, pool not available yet, HDF_NO_FUSED_FINALIZE)
unsigned* fin_tickets(int N, cudaStream_t s) { return N < HDF_TICKETS_PER_SLOT ? hdf_ticket_slot((void*)s) : nullptr; }

// ---------------------------------------------------------------------------
// elementwise kernels over [N, V, C] vectors
// ---------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void in_apply_kernel(const T* __restrict__ y, long long ldy, const float* __restrict__ mean,
                                const float* __restrict__ rstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const T* __restrict__ res, long long ldr,
                                T* __restrict__ out, long long ldo, long long V, int C, int relu, long long total) {
  const int cpv = C / VEC;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cpv) * VEC;
    const long long row = i / cpv;
    const int n = (int)(row / V);
    float x[VEC], r[VEC];
    loadv<T, VEC>(y + row * ldy + c, x);
    if (res) loadv<T, VEC>(res + row * ldr + c, r);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float z = (x[k] - mean[n * C + c + k]) * rstd[n * C + c + k];
      z = z * (gamma ? gamma[c + k] : 1.f) + (beta ? beta[c + k] : 0.f);
      if (relu) z = fmaxf(z, 0.f);
      if (res) z += r[k];
      x[k] = z;
    }
    storev<T, VEC>(out + row * ldo + c, x);
  }
}

template <typename T, int VEC>
__global__ void in_bwd_apply_kernel(const T* __restrict__ dout, long long ldd, const T* __restrict__ y, long long ldy,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ s1, const float* __restrict__ s2, T* __restrict__ dy,
                                    long long ldo, long long V, int C, int relu, long long total) {
  const int cpv = C / VEC;
  const float invV = 1.f / (float)V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cpv) * VEC;
    const long long row = i / cpv;
    const int n = (int)(row / V);
    float g[VEC], x[VEC];
    loadv<T, VEC>(dout + row * ldd + c, g);
    loadv<T, VEC>(y + row * ldy + c, x);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int nc = n * C + c + k;
      const float rs = rstd[nc];
      const float xh = (x[k] - mean[nc]) * rs;
      const float gm = gamma ? gamma[c + k] : 1.f;
      const float z = xh * gm + (beta ? beta[c + k] : 0.f);
      const float dz = (!relu || z > 0.f) ? g[k] : 0.f;
      x[k] = gm * rs * (dz - s1[nc] * invV - xh * s2[nc] * invV);
    }
    storev<T, VEC>(dy + row * ldo + c, x);
  }
}

// Fast forms for the common case 256 % (C/VEC) == 0: grid (row chunks, N); a thread keeps one channel group for its whole
// life, so the per-channel statistics / affine constants sit in registers (the generic kernels above re-read 4-6
// scalars per element and pay two 64-bit divisions per vector), and four rows are in flight per iteration.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) in_apply_fast_kernel(const T* __restrict__ y, long long ldy, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const T* __restrict__ res,
                                                           long long ldr, T* __restrict__ out, long long ldo, int V, int C,
                                                           int relu, int rows_per_block) {
  const int cpv = C / VEC, rpi = 256 / cpv;
  const int cg = threadIdx.x % cpv, r = threadIdx.x / cpv, n = blockIdx.y, c = cg * VEC;
  float m[VEC], a[VEC], b[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    m[k] = mean[n * C + c + k];
    a[k] = rstd[n * C + c + k];
    b[k] = beta ? beta[c + k] : 0.f;
  }
  float gm[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) gm[k] = gamma ? gamma[c + k] : 1.f;
  const int v0 = blockIdx.x * rows_per_block, v1 = min(V, v0 + rows_per_block);
  const T* yp = y + (long long)n * V * ldy + c;
  const T* rp = res ? res + (long long)n * V * ldr + c : nullptr;
  T* op = out + (long long)n * V * ldo + c;
  const bool has_res = rp != nullptr;
  auto one = [&](float* x, const float* rr) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float z = (x[k] - m[k]) * a[k];
      z = z * gm[k] + b[k];
      if (relu) z = fmaxf(z, 0.f);
      if (has_res) z += rr[k];
      x[k] = z;
    }
  };
  int v = v0 + r;
  for (; v + 3 * rpi < v1; v += 4 * rpi) {
    float x[4][VEC], rr[4][VEC];
#pragma unroll
    for (int u = 0; u < 4; ++u) loadv<T, VEC>(yp + (long long)(v + u * rpi) * ldy, x[u]);
    if (has_res) {
#pragma unroll
      for (int u = 0; u < 4; ++u) loadv<T, VEC>(rp + (long long)(v + u * rpi) * ldr, rr[u]);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) rr[u][k] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      one(x[u], rr[u]);
      storev<T, VEC>(op + (long long)(v + u * rpi) * ldo, x[u]);
    }
  }
  for (; v < v1; v += rpi) {
    float x[VEC], rr[VEC];
    loadv<T, VEC>(yp + (long long)v * ldy, x);
    if (has_res) loadv<T, VEC>(rp + (long long)v * ldr, rr);
    else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) rr[k] = 0.f;
    }
    one(x, rr);
    storev<T, VEC>(op + (long long)v * ldo, x);
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) in_bwd_apply_fast_kernel(const T* __restrict__ dout, long long ldd, const T* __restrict__ y,
                                                               long long ldy, const float* __restrict__ mean,
                                                               const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, const float* __restrict__ s1,
                                                               const float* __restrict__ s2, T* __restrict__ dy, long long ldo,
                                                               int V, int C, int relu, int rows_per_block) {
  const int cpv = C / VEC, rpi = 256 / cpv;
  const int cg = threadIdx.x % cpv, r = threadIdx.x / cpv, n = blockIdx.y, c = cg * VEC;
  const float invV = 1.f / (float)V;
  float m[VEC], rs[VEC], gm[VEC], bt[VEC], q1[VEC], q2[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const int nc = n * C + c + k;
    m[k] = mean[nc];
    rs[k] = rstd[nc];
    gm[k] = gamma ? gamma[c + k] : 1.f;
    bt[k] = beta ? beta[c + k] : 0.f;
    q1[k] = s1[nc] * invV;
    q2[k] = s2[nc] * invV;
  }
  const int v0 = blockIdx.x * rows_per_block, v1 = min(V, v0 + rows_per_block);
  const T* gp = dout + (long long)n * V * ldd + c;
  const T* yp = y + (long long)n * V * ldy + c;
  T* op = dy + (long long)n * V * ldo + c;
  auto one = [&](const float* g, float* x) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float xh = (x[k] - m[k]) * rs[k];
      const float z = xh * gm[k] + bt[k];
      const float dz = (!relu || z > 0.f) ? g[k] : 0.f;
      x[k] = gm[k] * rs[k] * (dz - q1[k] - xh * q2[k]);
    }
  };
  int v = v0 + r;
  for (; v + 3 * rpi < v1; v += 4 * rpi) {
    float g[4][VEC], x[4][VEC];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      loadv<T, VEC>(gp + (long long)(v + u * rpi) * ldd, g[u]);
      loadv<T, VEC>(yp + (long long)(v + u * rpi) * ldy, x[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      one(g[u], x[u]);
      storev<T, VEC>(op + (long long)(v + u * rpi) * ldo, x[u]);
    }
  }
  for (; v < v1; v += rpi) {
    float g[VEC], x[VEC];
    loadv<T, VEC>(gp + (long long)v * ldd, g);
    loadv<T, VEC>(yp + (long long)v * ldy, x);
    one(g, x);
    storev<T, VEC>(op + (long long)v * ldo, x);
  }
}

// InstanceNorm apply + affine + ReLU fused with the 1x1x1 head that reads its output (models/HDenseFormer.py:253-255: the four
// deep-supervision heads read the outputs of block_k_2_right / block_4_2_left): the head's dot products are taken from the
// registers that hold the (bf16-rounded) output row, so the head does not read the 2 x 191 MB activation again.  Thread =
// (row, 8-channel group) like in_apply_fast_kernel; the C/8 threads of a row combine their partial sums with shuffles.
__device__ __forceinline__ void unpack8(const uint4& v, float* o) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { o[2 * i] = __low2float(h[i]); o[2 * i + 1] = __high2float(h[i]); }
}
template <int NC>
__global__ void __launch_bounds__(256) in_apply_head_kernel(const bf16* __restrict__ y, long long ldy, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, bf16* __restrict__ out, long long ldo,
                                                           int V, int C, int relu, int rows_per_block,
                                                           const float* __restrict__ hw, const float* __restrict__ hb,
                                                           bf16* __restrict__ logits, int ncls) {
  constexpr int U = 4;
  const int cpv = C / 8, rpi = 256 / cpv;
  const int cg = threadIdx.x % cpv, r = threadIdx.x / cpv, n = blockIdx.y, c = cg * 8;
  float m[8], a[8], gm[8], b[8], wk[NC][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    m[k] = mean[n * C + c + k];
    a[k] = rstd[n * C + c + k];
    gm[k] = gamma ? gamma[c + k] : 1.f;
    b[k] = beta ? beta[c + k] : 0.f;
  }
#pragma unroll
  for (int q = 0; q < NC; ++q) {
#pragma unroll
    for (int k = 0; k < 8; ++k) wk[q][k] = q < ncls ? hw[q * C + c + k] : 0.f;
  }
  const float bias = (cg < ncls && hb) ? hb[cg] : 0.f;      // thread cg of the row writes class cg
  const int v0 = blockIdx.x * rows_per_block, v1 = min(V, v0 + rows_per_block);
  const bf16* yp = y + (long long)n * V * ldy + c;
  bf16* op = out + (long long)n * V * ldo + c;
  bf16* lp = logits + (long long)n * ncls * V;
  auto one = [&](const uint4& ry, int v, bool valid) {
    float x[8];
    unpack8(ry, x);
    uint4 pk;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float z = (x[k] - m[k]) * a[k];           // same operation order as in_apply_fast_kernel
      z = z * gm[k] + b[k];
      if (relu) z = fmaxf(z, 0.f);
      x[k] = z;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    if (valid) *reinterpret_cast<uint4*>(op + (long long)v * ldo) = pk;
    unpack8(pk, x);                             // the head sees what a separate head kernel would read back: bf16 values
    float mine = 0.f;
#pragma unroll
    for (int q = 0; q < NC; ++q) {
      float sdot = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) sdot = fmaf(x[k], wk[q][k], sdot);
      for (int o = cpv >> 1; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
      if (q == cg) mine = sdot;
    }
    if (valid && cg < ncls) lp[(long long)cg * V + v] = __float2bfloat16_rn(mine + bias);
  };
  // every lane of a warp takes part in the shuffles: rows past the end of the block's range recompute a clamped row and
  // only their stores are masked (the trip count is uniform over the block)
  const int v = v0 + r;
  const int iters = (v1 - v0 + rpi - 1) / rpi;
  for (int it = 0; it < iters; it += U) {
    uint4 ry[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int vv = min(v + (it + u) * rpi, V - 1);
      ry[u] = *reinterpret_cast<const uint4*>(yp + (long long)vv * ldy);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int vv = v + (it + u) * rpi;
      if (it + u < iters) one(ry[u], vv, vv < v1);
    }
  }
}

// rows per block for the fast elementwise kernels: ~8 resident-CTA waves over the 148 SMs, >= 4 passes of the block
inline int fast_rows_per_block(long long V, int N, int rpi) {
  long long target = (148ll * 16 + N - 1) / N;
  long long rpb = (V + target - 1) / target;
  const long long unit = 4ll * rpi;
  rpb = (rpb + unit - 1) / unit * unit;
  return (int)(rpb < unit ? unit : rpb);
}
inline bool fast_ok(int C, int vec, long long V) { const int cpv = C / vec; return cpv >= 1 && cpv <= 256 && 256 % cpv == 0 && V < (1ll << 31); }

template <typename T, int VEC>
__global__ void add_kernel(T* __restrict__ dst, long long ldd, const T* __restrict__ src, long long lds, int C,
                           long long total) {
  const int cpv = C / VEC;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cpv) * VEC;
    const long long row = i / cpv;
    float a[VEC], b[VEC];
    loadv<T, VEC>(dst + row * ldd + c, a);
    loadv<T, VEC>(src + row * lds + c, b);
#pragma unroll
    for (int k = 0; k < VEC; ++k) a[k] += b[k];
    storev<T, VEC>(dst + row * ldd + c, a);
  }
}

template <typename T, int VEC>
__global__ void copy_kernel(T* __restrict__ dst, long long ldd, const T* __restrict__ src, long long lds, int C,
                            long long total) {
  const int cpv = C / VEC;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cpv) * VEC;
    const long long row = i / cpv;
    float a[VEC];
    loadv<T, VEC>(src + row * lds + c, a);
    storev<T, VEC>(dst + row * ldd + c, a);
  }
}

// fp32 rows [rows, C] (ld) -> T rows (ld)   (token features into a channel slice of the conv input)
template <typename T>
__global__ void cast_rows_kernel(const float* __restrict__ src, long long lds, T* __restrict__ dst, long long ldd, int C,
                                 long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long row = i / C;
    dst[row * ldd + c] = from_f<T>(src[row * lds + c]);
  }
}
template <typename T>
__global__ void uncast_rows_kernel(const T* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd, int C,
                                   long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long row = i / C;
    dst[row * ldd + c] = to_f(src[row * lds + c]);
  }
}

// The spatial kernels below run one block per (n, d, h) line: the line coordinates come from blockIdx (a few 32-bit
// divisions per block), a thread walks the (w, channel-group) items of the line.

// 2x2x2 max pool, first maximum in (kd,kh,kw) scan order wins (torch semantics: strict '>' update)
template <typename T, int VEC>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, long long ldx, T* __restrict__ out, long long ldo, int Do,
                                   int Ho, int Wo, int C, int pd) {
  // pd = pooling window along depth: 2 (MaxPool3d(2)) or 1 (MaxPool2d(2) on a [N, 1, H, W, C] slice, models/HDenseFormer_2D.py)
  const int cpv = C / VEC;
  const int Hi = 2 * Ho, Wi = 2 * Wo, Di = pd * Do;
  int line = blockIdx.x;
  const int h = line % Ho; line /= Ho;
  const int d = line % Do;
  const long long n = line / Do;
  const T* xin = x + (((n * Di + pd * d) * Hi + 2 * h) * (long long)Wi) * ldx;
  T* orow = out + ((long long)blockIdx.x * Wo) * ldo;
  for (int i = threadIdx.x; i < Wo * cpv; i += blockDim.x) {
    const int w = i / cpv, c = (i - w * cpv) * VEC;
    float m[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) m[k] = -INFINITY;
    float v[8][VEC];
#pragma unroll
    for (int t = 0; t < 8; ++t)
      if ((t >> 2) < pd) loadv<T, VEC>(xin + (((long long)(t >> 2) * Hi + ((t >> 1) & 1)) * Wi + 2 * w + (t & 1)) * ldx + c, v[t]);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if ((t >> 2) >= pd) continue;
#pragma unroll
      for (int k = 0; k < VEC; ++k) m[k] = (v[t][k] > m[k] || v[t][k] != v[t][k]) ? v[t][k] : m[k];
    }
    storev<T, VEC>(orow + (long long)w * ldo + c, m);
  }
}

// dx[argmax] (+)= dpool ; other 7 positions get 0 when !accumulate
template <typename T, int VEC>
__global__ void maxpool_bwd_kernel(const T* __restrict__ x, long long ldx, const T* __restrict__ dp, long long ldp,
                                   T* __restrict__ dx, long long lddx, int Do, int Ho, int Wo, int C, int accumulate, int pd) {
  const int cpv = C / VEC;
  const int Hi = 2 * Ho, Wi = 2 * Wo, Di = pd * Do;
  int line = blockIdx.x;
  const int h = line % Ho; line /= Ho;
  const int d = line % Do;
  const long long n = line / Do;
  const long long row0 = ((n * Di + pd * d) * Hi + 2 * h) * (long long)Wi;
  const T* prow = dp + ((long long)blockIdx.x * Wo) * ldp;
  for (int i = threadIdx.x; i < Wo * cpv; i += blockDim.x) {
    const int w = i / cpv, c = (i - w * cpv) * VEC;
    float m[VEC], g[VEC];
    int am[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) { m[k] = -INFINITY; am[k] = 0; }
    long long rows[8];
    float v[8][VEC];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      rows[t] = row0 + ((long long)(t >> 2) * Hi + ((t >> 1) & 1)) * Wi + 2 * w + (t & 1);
      if ((t >> 2) < pd) loadv<T, VEC>(x + rows[t] * ldx + c, v[t]);
    }
    loadv<T, VEC>(prow + (long long)w * ldp + c, g);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if ((t >> 2) >= pd) continue;
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        if (v[t][k] > m[k] || v[t][k] != v[t][k]) { m[k] = v[t][k]; am[k] = t; }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if ((t >> 2) >= pd) continue;
      float o[VEC];
      if (accumulate) loadv<T, VEC>(dx + rows[t] * lddx + c, o);
#pragma unroll
      for (int k = 0; k < VEC; ++k) o[k] = (accumulate ? o[k] : 0.f) + (am[k] == t ? g[k] : 0.f);
      storev<T, VEC>(dx + rows[t] * lddx + c, o);
    }
  }
}

// trilinear x2, align_corners=False: src = (dst+0.5)/2-0.5 clamped at 0
__device__ __forceinline__ void up2_src(int o, int In, int& i0, int& i1, float& w1) {
  float s = (o + 0.5f) * 0.5f - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  i1 = i0 + (i0 < In - 1 ? 1 : 0);
  w1 = s - (float)i0;
}

template <typename T, int VEC>
__global__ void upsample2_fwd_kernel(const T* __restrict__ x, long long ldx, T* __restrict__ out, long long ldo, int Di,
                                     int Hi, int Wi, int C, int sd) {
  // sd = scale along depth: 2 (trilinear) or 1 (bilinear x2 of a [N, 1, H, W, C] slice, models/HDenseFormer_2D.py:171)
  const int cpv = C / VEC;
  const int Do = sd * Di, Ho = 2 * Hi, Wo = 2 * Wi;
  int line = blockIdx.x;
  const int h = line % Ho; line /= Ho;
  const int d = line % Do;
  const long long n = line / Do;
  int d0, d1, h0, h1;
  float fd, fh;
  if (sd == 1) { d0 = d1 = d; fd = 0.f; }
  else up2_src(d, Di, d0, d1, fd);
  up2_src(h, Hi, h0, h1, fh);
  const T* l00 = x + (((n * Di + d0) * Hi + h0) * (long long)Wi) * ldx;
  const T* l01 = x + (((n * Di + d0) * Hi + h1) * (long long)Wi) * ldx;
  const T* l10 = x + (((n * Di + d1) * Hi + h0) * (long long)Wi) * ldx;
  const T* l11 = x + (((n * Di + d1) * Hi + h1) * (long long)Wi) * ldx;
  const float w00 = (1.f - fd) * (1.f - fh), w01 = (1.f - fd) * fh, w10 = fd * (1.f - fh), w11 = fd * fh;
  T* orow = out + ((long long)blockIdx.x * Wo) * ldo;
  for (int i = threadIdx.x; i < Wo * cpv; i += blockDim.x) {
    const int w = i / cpv, c = (i - w * cpv) * VEC;
    int w0, w1;
    float fw;
    up2_src(w, Wi, w0, w1, fw);
    float v[8][VEC];
    loadv<T, VEC>(l00 + (long long)w0 * ldx + c, v[0]);
    loadv<T, VEC>(l00 + (long long)w1 * ldx + c, v[1]);
    loadv<T, VEC>(l01 + (long long)w0 * ldx + c, v[2]);
    loadv<T, VEC>(l01 + (long long)w1 * ldx + c, v[3]);
    loadv<T, VEC>(l10 + (long long)w0 * ldx + c, v[4]);
    loadv<T, VEC>(l10 + (long long)w1 * ldx + c, v[5]);
    loadv<T, VEC>(l11 + (long long)w0 * ldx + c, v[6]);
    loadv<T, VEC>(l11 + (long long)w1 * ldx + c, v[7]);
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    // same accumulation order as the (d,h,w)-bit enumeration t = 0..7 of the weights
    const float wt[8] = {w00 * (1.f - fw), w00 * fw, w01 * (1.f - fw), w01 * fw, w10 * (1.f - fw), w10 * fw, w11 * (1.f - fw), w11 * fw};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = fmaf(wt[t], v[t][k], acc[k]);
    }
    storev<T, VEC>(orow + (long long)w * ldo + c, acc);
  }
}

// bf16 fast form, cell-centred and separable: the cell between input voxels (c-1, c) per dim (clamped at the borders)
// owns the outputs (2c-1, 2c); a thread loads the cell's 8 input vectors once and produces its (up to) 8 output vectors
// with three 2-tap stages (w, h, d).  The per-output kernel above re-loads and re-converts 8 inputs for every output and
// is issue-bound (84 % issue slots busy, 25 % of HBM: profiles/r1_ncu_glue.txt); this one does ~6x fewer instructions.
// Weights per dim: output 2c-1 = .75 lo + .25 hi, output 2c = .25 lo + .75 hi (lo == hi at the clamped borders, where the
// fused multiply-add returns the input exactly), i.e. align_corners=False.
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__global__ void __launch_bounds__(256) upsample2_fwd_cell_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ out,
                                                                long long ldo, int Di, int Hi, int Wi, int C) {
  const int cpv = C / 8;
  int cell = blockIdx.x;
  const int ch = cell % (Hi + 1); cell /= (Hi + 1);
  const int cd = cell % (Di + 1);
  const long long n = cell / (Di + 1);
  const int d0 = max(cd - 1, 0), d1 = min(cd, Di - 1), h0 = max(ch - 1, 0), h1 = min(ch, Hi - 1);
  const bf16* l[4] = {x + (((n * Di + d0) * Hi + h0) * (long long)Wi) * ldx, x + (((n * Di + d0) * Hi + h1) * (long long)Wi) * ldx,
                      x + (((n * Di + d1) * Hi + h0) * (long long)Wi) * ldx, x + (((n * Di + d1) * Hi + h1) * (long long)Wi) * ldx};
  const int Do = 2 * Di, Ho = 2 * Hi, Wo = 2 * Wi;
  const bool vd[2] = {cd >= 1, cd <= Di - 1}, vh[2] = {ch >= 1, ch <= Hi - 1};
  for (int i = threadIdx.x; i < (Wi + 1) * cpv; i += blockDim.x) {
    const int cw = i / cpv, c = (i - cw * cpv) * 8;
    const int w0 = max(cw - 1, 0), w1 = min(cw, Wi - 1);
    const bool vw[2] = {cw >= 1, cw <= Wi - 1};
    uint4 in[4][2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      in[q][0] = *reinterpret_cast<const uint4*>(l[q] + (long long)w0 * ldx + c);
      in[q][1] = *reinterpret_cast<const uint4*>(l[q] + (long long)w1 * ldx + c);
    }
    uint4 o[2][2][2];   // [pd][ph][pw]
#pragma unroll
    for (int r = 0; r < 4; ++r) {          // 32-bit word r = channels 2r, 2r+1
      float res[2][2][2][2];               // [pd][ph][pw][half]
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float a[2][2][2];                  // [d][h][pw]
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t u0 = (&in[q][0].x)[r], u1 = (&in[q][1].x)[r];
          const float lo = half ? bf_hi(u0) : bf_lo(u0), hi = half ? bf_hi(u1) : bf_lo(u1);
          a[q >> 1][q & 1][0] = fmaf(0.25f, hi, 0.75f * lo);
          a[q >> 1][q & 1][1] = fmaf(0.75f, hi, 0.25f * lo);
        }
        float b[2][2][2];                  // [d][ph][pw]
#pragma unroll
        for (int d = 0; d < 2; ++d) {
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            b[d][0][pw] = fmaf(0.25f, a[d][1][pw], 0.75f * a[d][0][pw]);
            b[d][1][pw] = fmaf(0.75f, a[d][1][pw], 0.25f * a[d][0][pw]);
          }
        }
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            res[0][ph][pw][half] = fmaf(0.25f, b[1][ph][pw], 0.75f * b[0][ph][pw]);
            res[1][ph][pw][half] = fmaf(0.75f, b[1][ph][pw], 0.25f * b[0][ph][pw]);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e)
        (&o[e >> 2][(e >> 1) & 1][e & 1].x)[r] = pack_bf2(res[e >> 2][(e >> 1) & 1][e & 1][0], res[e >> 2][(e >> 1) & 1][e & 1][1]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int pd = e >> 2, ph = (e >> 1) & 1, pw = e & 1;
      if (vd[pd] && vh[ph] && vw[pw]) {
        const long long orow = ((n * Do + 2 * cd - 1 + pd) * Ho + 2 * ch - 1 + ph) * (long long)Wo + 2 * cw - 1 + pw;
        *reinterpret_cast<uint4*>(out + orow * ldo + c) = o[pd][ph][pw];
      }
    }
  }
}

// gather form of the transpose: input voxel i receives from outputs 2i-1 (.25), 2i (.75 | 1 at i=0),
// 2i+1 (.75 | 1 at i=In-1), 2i+2 (.25)
__device__ __forceinline__ void up2_bwd_taps(int i, int In, int* o, float* w) {
  o[0] = 2 * i - 1; w[0] = (i >= 1) ? 0.25f : 0.f;
  o[1] = 2 * i;     w[1] = (i == 0) ? 1.0f : 0.75f;
  o[2] = 2 * i + 1; w[2] = (i == In - 1) ? 1.0f : 0.75f;
  o[3] = 2 * i + 2; w[3] = (i <= In - 2) ? 0.25f : 0.f;
}

template <typename T, int VEC>
__global__ void upsample2_bwd_kernel(const T* __restrict__ dout, long long ldd, T* __restrict__ dx, long long lddx, int Di,
                                     int Hi, int Wi, int C, int accumulate, int sd) {
  const int cpv = C / VEC;
  const int Ho = 2 * Hi, Wo = 2 * Wi, Do = sd * Di;
  int line = blockIdx.x;
  const int h = line % Hi; line /= Hi;
  const int d = line % Di;
  const long long n = line / Di;
  int od[4], oh[4];
  float wd[4], wh[4];
  if (sd == 1) { od[0] = od[1] = od[2] = od[3] = d; wd[0] = wd[2] = wd[3] = 0.f; wd[1] = 1.f; }
  else up2_bwd_taps(d, Di, od, wd);
  up2_bwd_taps(h, Hi, oh, wh);
  T* xrow = dx + ((long long)blockIdx.x * Wi) * lddx;
  for (int i = threadIdx.x; i < Wi * cpv; i += blockDim.x) {
    const int w = i / cpv, c = (i - w * cpv) * VEC;
    int ow[4];
    float ww[4];
    up2_bwd_taps(w, Wi, ow, ww);
    float acc[VEC];
    if (accumulate) loadv<T, VEC>(xrow + (long long)w * lddx + c, acc);
    else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (wd[a] == 0.f) continue;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (wh[b] == 0.f) continue;
        const T* lrow = dout + (((n * Do + od[a]) * Ho + oh[b]) * (long long)Wo) * ldd + c;
        const float wab = wd[a] * wh[b];
        float v[4][VEC];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (ww[e] != 0.f) loadv<T, VEC>(lrow + (long long)ow[e] * ldd, v[e]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (ww[e] == 0.f) continue;
          const float wt = wab * ww[e];
#pragma unroll
          for (int k = 0; k < VEC; ++k) acc[k] = fmaf(wt, v[e][k], acc[k]);
        }
      }
    }
    storev<T, VEC>(xrow + (long long)w * lddx + c, acc);
  }
}

// bf16 fast form of the transpose, separable through shared memory: one CTA per input line (n, d, h).  Phase 1 folds the
// (up to) 4 x 4 contributing output lines into one fp32 line s[ow][c] = sum_ab wd[a] wh[b] dout[od[a], oh[b], ow, c]
// (coalesced row loads); phase 2 applies the 4 w taps from shared memory.  ~1.8x fewer instructions than the 64-tap
// gather above, but measured SLOWER on B200 (0.49 vs 0.30 ms at 144^3) -> opt-in experiment (HDF_UPS_BWD_V2=1).
__global__ void __launch_bounds__(256) upsample2_bwd_line_kernel(const bf16* __restrict__ dout, long long ldd, bf16* __restrict__ dx,
                                                                long long lddx, int Di, int Hi, int Wi, int C, int accumulate) {
  extern __shared__ float srow[];            // [2*Wi][C]
  const int cpv = C / 8;
  const int Ho = 2 * Hi, Wo = 2 * Wi, Do = 2 * Di;
  int line = blockIdx.x;
  const int h = line % Hi; line /= Hi;
  const int d = line % Di;
  const long long n = line / Di;
  int od[4], oh[4];
  float wd[4], wh[4];
  up2_bwd_taps(d, Di, od, wd);
  up2_bwd_taps(h, Hi, oh, wh);
  for (int i = threadIdx.x; i < Wo * cpv; i += blockDim.x) {
    const int ow = i / cpv, c = (i - ow * cpv) * 8;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (wd[a] == 0.f) continue;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (wh[b] == 0.f) continue;
        const float wt = wd[a] * wh[b];
        float v[8];
        load8<bf16>(dout + ((((n * Do + od[a]) * Ho + oh[b]) * (long long)Wo) + ow) * ldd + c, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(wt, v[k], acc[k]);
      }
    }
    float* dst = srow + (long long)ow * C + c;
    *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  __syncthreads();
  bf16* xrow = dx + ((long long)blockIdx.x * Wi) * lddx;
  for (int i = threadIdx.x; i < Wi * cpv; i += blockDim.x) {
    const int w = i / cpv, c = (i - w * cpv) * 8;
    int ow[4];
    float ww[4];
    up2_bwd_taps(w, Wi, ow, ww);
    float acc[8];
    if (accumulate) load8<bf16>(xrow + (long long)w * lddx + c, acc);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (ww[e] == 0.f) continue;
      const float* src = srow + (long long)ow[e] * C + c;
      const float4 s0 = *reinterpret_cast<const float4*>(src), s1 = *reinterpret_cast<const float4*>(src + 4);
      acc[0] = fmaf(ww[e], s0.x, acc[0]); acc[1] = fmaf(ww[e], s0.y, acc[1]); acc[2] = fmaf(ww[e], s0.z, acc[2]); acc[3] = fmaf(ww[e], s0.w, acc[3]);
      acc[4] = fmaf(ww[e], s1.x, acc[4]); acc[5] = fmaf(ww[e], s1.y, acc[5]); acc[6] = fmaf(ww[e], s1.z, acc[6]); acc[7] = fmaf(ww[e], s1.w, acc[7]);
    }
    store8<bf16>(xrow + (long long)w * lddx + c, acc);
  }
}

// bf16 fast form of the transpose, register blocked: a CTA produces a 2 x 2 block of input lines (d, h) and a thread the
// four results of one column (8 channels each) from the 6 x 6 output lines that reach them -- 36 sixteen-byte loads per
// result instead of the 64 of the gather above, and the CTA reads 9 output lines per input line instead of 16: the
// gather is bound by L2 -> SM traffic (1.5 GB for the 2 x 144^3 x 32 level, 0.30 ms), not by HBM (0.43 GB).
// w[i][a]: weight of output offset a (output index 2 x0 - 1 + a) in input x0 + i.
__device__ __forceinline__ void up2_blk_weights(int x0, int In, float (*w)[6]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int o[4];
    float t[4];
    up2_bwd_taps(x0 + i, In, o, t);
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const int tap = a - 2 * i;
      w[i][a] = (tap >= 0 && tap <= 3 && x0 + i < In) ? t[tap < 0 ? 0 : (tap > 3 ? 3 : tap)] : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256, 2) upsample2_bwd_blk_kernel(const bf16* __restrict__ dout, long long ldd, bf16* __restrict__ dx,
                                                               long long lddx, int Di, int Hi, int Wi, int C, int accumulate) {
  const int cpv = C / 8;
  const int Ho = 2 * Hi, Wo = 2 * Wi, Do = 2 * Di;
  const int Hb = (Hi + 1) / 2, Db = (Di + 1) / 2;
  int blk = blockIdx.x;
  const int hb = blk % Hb; blk /= Hb;
  const int db = blk % Db;
  const long long n = blk / Db;
  const int d0 = 2 * db, h0 = 2 * hb;
  float wh[2][6];
  up2_blk_weights(h0, Hi, wh);
  // the d weights are looked up per plane (the plane loop is not unrolled: 36 unrolled lines spill)
  __shared__ float wd_sh[2][6];
  if (threadIdx.x == 0) {
    float wd[2][6];
    up2_blk_weights(d0, Di, wd);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int a = 0; a < 6; ++a) wd_sh[i][a] = wd[i][a];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < Wi * cpv; t += blockDim.x) {
    const int w = t / cpv, c = (t - w * cpv) * 8;
    int ow[4];
    float ww[4];
    up2_bwd_taps(w, Wi, ow, ww);
    float acc[2][2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[i][j][e] = 0.f;
#pragma unroll 1
    for (int a = 0; a < 6; ++a) {
      const int od = 2 * d0 - 1 + a;
      if (od < 0 || od >= Do) continue;
      const float wd0 = wd_sh[0][a], wd1 = wd_sh[1][a];
#pragma unroll
      for (int b = 0; b < 6; ++b) {
        const int oh = 2 * h0 - 1 + b;
        if (oh < 0 || oh >= Ho) continue;
        const bf16* lrow = dout + (((n * Do + od) * Ho + oh) * (long long)Wo) * ldd + c;
        float v[4][8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (ww[e] != 0.f) load8<bf16>(lrow + (long long)ow[e] * ldd, v[e]);
        }
        float s[8];                    // the w taps folded: this line's contribution before the (d, h) weights
#pragma unroll
        for (int q = 0; q < 8; ++q) s[q] = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (ww[e] == 0.f) continue;
#pragma unroll
          for (int q = 0; q < 8; ++q) s[q] = fmaf(ww[e], v[e][q], s[q]);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (b - 2 * j < 0 || b - 2 * j > 3) continue;
          const float w0 = wd0 * wh[j][b], w1 = wd1 * wh[j][b];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            acc[0][j][q] = fmaf(w0, s[q], acc[0][j][q]);
            acc[1][j][q] = fmaf(w1, s[q], acc[1][j][q]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (d0 + i >= Di || h0 + j >= Hi) continue;
        bf16* row = dx + ((((n * Di + d0 + i) * Hi + h0 + j) * (long long)Wi) + w) * lddx + c;
        if (accumulate) {
          float old[8];
          load8<bf16>(row, old);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[i][j][q] += old[q];
        }
        store8<bf16>(row, acc[i][j]);
      }
  }
}

// ---------------------------------------------------------------------------
// 1x1x1 heads: out[n, k, v] (NCDHW, T) = sum_c a[n, v, c] * w[k, c] + b[k]
// ---------------------------------------------------------------------------
constexpr int MAXCLS = 8;

template <typename T, int VEC>
__global__ void head_fwd_kernel(const T* __restrict__ a, long long lda, const float* __restrict__ w,
                                const float* __restrict__ b, T* __restrict__ out, long long V, int C, int ncls,
                                long long total) {
  extern __shared__ float sw[];  // [ncls][C] + [ncls]
  for (int i = threadIdx.x; i < ncls * C; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < ncls; i += blockDim.x) sw[ncls * C + i] = b ? b[i] : 0.f;
  __syncthreads();
  for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < total; row += (long long)gridDim.x * blockDim.x) {
    float acc[MAXCLS];
#pragma unroll
    for (int k = 0; k < MAXCLS; ++k) acc[k] = (k < ncls) ? sw[ncls * C + k] : 0.f;
    for (int c = 0; c < C; c += VEC) {
      float x[VEC];
      loadv<T, VEC>(a + row * lda + c, x);
#pragma unroll
      for (int k = 0; k < MAXCLS; ++k) {
        if (k < ncls) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[k] = fmaf(x[j], sw[k * C + c + j], acc[k]);
        }
      }
    }
    const long long n = row / V, v = row % V;
#pragma unroll
    for (int k = 0; k < MAXCLS; ++k)
      if (k < ncls) out[(n * ncls + k) * V + v] = from_f<T>(acc[k]);
  }
}

// da[n, v, c] (+)= sum_k g[n, k, v] * w[k, c]
template <typename T, int VEC>
__global__ void head_dgrad_kernel(const T* __restrict__ g, const float* __restrict__ w, T* __restrict__ da, long long ldd,
                                  long long V, int C, int ncls, int accumulate, long long total) {
  extern __shared__ float sw[];
  for (int i = threadIdx.x; i < ncls * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < total; row += (long long)gridDim.x * blockDim.x) {
    const long long n = row / V, v = row % V;
    float gk[MAXCLS];
#pragma unroll
    for (int k = 0; k < MAXCLS; ++k) gk[k] = (k < ncls) ? to_f(g[(n * ncls + k) * V + v]) : 0.f;
    for (int c = 0; c < C; c += VEC) {
      float o[VEC];
      if (accumulate) loadv<T, VEC>(da + row * ldd + c, o);
      else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < MAXCLS; ++k) {
        if (k < ncls) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) o[j] = fmaf(gk[k], sw[k * C + c + j], o[j]);
        }
      }
      storev<T, VEC>(da + row * ldd + c, o);
    }
  }
}

// partial[n][chunk][k][C+1]: dW[k][c] = sum_v g[k][v] a[v][c];  column C holds db[k].
// grid (chunks, N).  Thread = (row lane, VEC-channel group): 128-bit loads of the activation row, U rows in flight.
// NC = class-count bucket (2, 4 or 8) so that two-class heads do not carry 64 dead accumulators (1 CTA/SM before).
template <typename T, int VEC, int NC>
__global__ void __launch_bounds__(256) head_wgrad_kernel(const T* __restrict__ g, const T* __restrict__ a, long long lda,
                                                        int V, int C, int ncls, int rows_per_chunk,
                                                        float* __restrict__ partial) {
  extern __shared__ float sm[];  // [lanes][ncls][C] then [lanes][ncls]
  constexpr int U = NC <= 2 ? 8 : 4;
  const int tid = threadIdx.x;
  const int cpv = C / VEC;                 // channel groups per row (<= 256 guaranteed by host)
  const int lanes = 256 / cpv;
  const int lane = tid / cpv, cg = tid % cpv;
  const int n = blockIdx.y;
  const int v0 = blockIdx.x * rows_per_chunk, v1 = min(V, v0 + rows_per_chunk);
  const T* gp = g + (long long)n * ncls * V;
  const T* ap = a + (long long)n * V * lda + cg * VEC;
  float acc[NC][VEC], accb[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    accb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[k][j] = 0.f;
  }
  if (lane < lanes) {
    int v = v0 + lane;
    for (; v + (U - 1) * lanes < v1; v += U * lanes) {
      float x[U][VEC], gv[U][NC];
#pragma unroll
      for (int u = 0; u < U; ++u) loadv<T, VEC>(ap + (long long)(v + u * lanes) * lda, x[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int k = 0; k < NC; ++k) gv[u][k] = (k < ncls) ? to_f(gp[(long long)k * V + v + u * lanes]) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          accb[k] += gv[u][k];
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[k][j] = fmaf(gv[u][k], x[u][j], acc[k][j]);
        }
      }
    }
    for (; v < v1; v += lanes) {
      float x[VEC];
      loadv<T, VEC>(ap + (long long)v * lda, x);
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        if (k < ncls) {
          const float gv = to_f(gp[(long long)k * V + v]);
          accb[k] += gv;
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[k][j] = fmaf(gv, x[j], acc[k][j]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      if (k < ncls) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) sm[(lane * ncls + k) * C + cg * VEC + j] = acc[k][j];
        if (cg == 0) sm[lanes * ncls * C + lane * ncls + k] = accb[k];
      }
    }
  }
  __syncthreads();
  const long long blk = (long long)n * gridDim.x + blockIdx.x;
  for (int i = tid; i < ncls * C; i += 256) {
    float s2 = 0.f;
    for (int l = 0; l < lanes; ++l) s2 += sm[l * ncls * C + i];
    const int k = i / C, c = i % C;
    partial[(blk * ncls + k) * (C + 1) + c] = s2;
  }
  if (tid < ncls) {
    float s2 = 0.f;
    for (int l = 0; l < lanes; ++l) s2 += sm[lanes * ncls * C + l * ncls + tid];
    partial[(blk * ncls + tid) * (C + 1) + C] = s2;
  }
}

template <typename T, int VEC, int NC>
void launch_head_wgrad(dim3 grid, size_t smem, cudaStream_t s, const T* g, const T* a, long long lda, int V, int C, int ncls,
                       int rpc, float* partial) {
  if (smem > 48 * 1024) cudaFuncSetAttribute(head_wgrad_kernel<T, VEC, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  head_wgrad_kernel<T, VEC, NC><<<grid, 256, smem, s>>>(g, a, lda, V, C, ncls, rpc, partial);
}

// da[n, v, c] (+)= sum_k g[n, k, v] * w[k, c]; same thread layout as the weight gradient: a warp writes whole rows
// (the row-per-thread kernel above stores 16 bytes per lane at a 2*C-byte stride: half-filled sectors)
template <typename T, int VEC, int NC>
__global__ void __launch_bounds__(256) head_dgrad_fast_kernel(const T* __restrict__ g, const float* __restrict__ w,
                                                             T* __restrict__ da, long long ldd, int V, int C, int ncls,
                                                             int accumulate, int rows_per_block) {
  const int cpv = C / VEC, lanes = 256 / cpv;
  const int lane = threadIdx.x / cpv, cg = threadIdx.x % cpv, n = blockIdx.y;
  float wk[NC][VEC];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) wk[k][j] = (k < ncls) ? w[k * C + cg * VEC + j] : 0.f;
  }
  const int v0 = blockIdx.x * rows_per_block, v1 = min(V, v0 + rows_per_block);
  const T* gp = g + (long long)n * ncls * V;
  T* dp = da + (long long)n * V * ldd + cg * VEC;
  auto one = [&](int v) {
    float o[VEC];
    if (accumulate) loadv<T, VEC>(dp + (long long)v * ldd, o);
    else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) o[j] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      if (k < ncls) {
        const float gv = to_f(gp[(long long)k * V + v]);
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] = fmaf(gv, wk[k][j], o[j]);
      }
    }
    storev<T, VEC>(dp + (long long)v * ldd, o);
  };
  int v = v0 + lane;
  for (; v + 3 * lanes < v1; v += 4 * lanes) {
#pragma unroll
    for (int u = 0; u < 4; ++u) one(v + u * lanes);
  }
  for (; v < v1; v += lanes) one(v);
}

__global__ void head_wgrad_finalize_kernel(const float* __restrict__ partial, int chunks, int C, int ncls,
                                           float* __restrict__ dw, float* __restrict__ db, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncls * (C + 1)) return;
  const int k = i / (C + 1), c = i % (C + 1);
  float s = 0.f;
  for (int z = 0; z < chunks; ++z) s += partial[((long long)z * ncls + k) * (C + 1) + c];
  if (c < C) dw[k * C + c] = accumulate ? dw[k * C + c] + s : s;
  else if (db) db[k] = accumulate ? db[k] + s : s;
}

// x NCDHW fp32 [N, C, V] -> channels-last T [N, V, C] (ld)
template <typename T>
__global__ void ncdhw_to_cl_kernel(const float* __restrict__ x, T* __restrict__ out, long long ldo, long long V, int C,
                                   long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / V, v = i % V;
    for (int c = 0; c < C; ++c) out[i * ldo + c] = from_f<T>(x[(n * C + c) * V + v]);
  }
}

// threads per line-block: the smallest multiple of 32 (<= 256) that covers the line's items in the fewest passes
int line_block(int items) {
  const int passes = (items + 255) / 256;
  int t = ((items + passes - 1) / passes + 31) / 32 * 32;
  return t < 32 ? 32 : (t > 256 ? 256 : t);
}

int grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = 148ll * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

#define HDF_VEC_DISPATCH(vec_ok, ...) \
  if (vec_ok) { constexpr int VEC = VecOf<T>::value; __VA_ARGS__; } else { constexpr int VEC = 1; __VA_ARGS__; }

extern "C" {

size_t hdf_reduce_workspace(int N, long long V, int C) {
  return (size_t)N * reduce_chunks(V, N) * 2 * C * sizeof(double) + 256;
}

int hdf_instnorm_stats(int dtype, const void* y, long long ldy, int N, long long V, int C, float eps, float* mean,
                       float* rstd, void* workspace, size_t ws_bytes, void* stream) {
  HDF_REQUIRE(y && mean && rstd && workspace && ws_bytes >= hdf_reduce_workspace(N, V, C), "hdf_instnorm_stats: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int chunks = 0;
  unsigned* tk = fin_tickets(N, s);
  HDF_DISPATCH_DTYPE(dtype, T, {
    StatsF<T> f{(const T*)y, ldy, V};
    int rc = tk ? launch_rowreduce<T>(f, y, ldy, nullptr, 0, N, V, C, (double*)workspace, chunks, s, "hdf_instnorm_stats",
                                      FinStats{V, eps, mean, rstd}, tk)
                : launch_rowreduce<T>(f, y, ldy, nullptr, 0, N, V, C, (double*)workspace, chunks, s, "hdf_instnorm_stats", FinNone{},
                                      nullptr);
    if (rc) return rc;
  });
  if (tk) return HDF_OK;
  stats_finalize_kernel<<<cdiv((long long)N * C * 32, 128), 128, 0, s>>>((const double*)workspace, chunks, C, V, eps, mean, rstd, N * C);
  HDF_LAUNCH_CHECK("hdf_instnorm_stats/finalize");
  return HDF_OK;
}

// mean / rstd from per-chunk partial sums [N][chunks][2][C] (double) written by another kernel (fused conv epilogue)
int hdf_instnorm_stats_finalize(const double* partial, int chunks, int N, int C, long long V, float eps, float* mean, float* rstd,
                                void* stream) {
  HDF_REQUIRE(partial && mean && rstd && chunks >= 1, "hdf_instnorm_stats_finalize: bad args");
  stats_finalize_kernel<<<cdiv((long long)N * C * 32, 128), 128, 0, (cudaStream_t)stream>>>(partial, chunks, C, V, eps, mean, rstd, N * C);
  HDF_LAUNCH_CHECK("hdf_instnorm_stats_finalize");
  return HDF_OK;
}

int hdf_instnorm_apply(int dtype, const void* y, long long ldy, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, const void* residual, long long ldr, void* out, long long ldo, int N,
                       long long V, int C, int relu, void* stream) {
  HDF_REQUIRE(y && mean && rstd && out, "hdf_instnorm_apply: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  HDF_DISPATCH_DTYPE(dtype, T, {
    const bool vec = can_vec<T>(y, ldy, C) && can_vec<T>(out, ldo, C) && (!residual || can_vec<T>(residual, ldr, C));
    HDF_VEC_DISPATCH(vec, {
        if (vec && sizeof(T) == 2 && fast_ok(C, VEC, V)) {
          const int rpb = fast_rows_per_block(V, N, 256 / (C / VEC));
          in_apply_fast_kernel<T, VEC><<<dim3((unsigned)cdiv(V, rpb), N), 256, 0, s>>>((const T*)y, ldy, mean, rstd, gamma, beta, (const T*)residual, ldr, (T*)out, ldo, (int)V, C, relu, rpb);
        } else {
          long long total = (long long)N * V * (C / VEC);
          in_apply_kernel<T, VEC><<<grid_for(total), 256, 0, s>>>((const T*)y, ldy, mean, rstd, gamma, beta, (const T*)residual, ldr, (T*)out, ldo, V, C, relu, total);
        } });
  });
  HDF_LAUNCH_CHECK("hdf_instnorm_apply");
  return HDF_OK;
}

// 1 if hdf_instnorm_apply_head takes this shape (bf16, C in {16..256} multiple of 8 with C/8 a power of two <= 32, ncls <= 4)
int hdf_instnorm_apply_head_supported(int C, int ncls) {
  static const bool off = getenv("HDF_NO_APPLY_HEAD") != nullptr;
  const int cpv = C / 8;
  return !off && C % 8 == 0 && cpv >= 1 && cpv <= 32 && (cpv & (cpv - 1)) == 0 && ncls >= 1 && ncls <= 4 && ncls <= cpv;
}

// out = relu(IN(y) * gamma + beta) (bf16, channel stride ldo)  AND  logits[n, k, v] = head_b[k] + sum_c out[n, v, c] head_w[k, c]
// (bf16, NCDHW) in one pass over y.  No residual.  Same results as hdf_instnorm_apply followed by hdf_head_fwd up to the
// fp32 summation order of the 1x1x1 head.
int hdf_instnorm_apply_head(const void* y, long long ldy, const float* mean, const float* rstd, const float* gamma, const float* beta,
                            void* out, long long ldo, int N, long long V, int C, int relu, const float* head_w, const float* head_b,
                            void* logits, int ncls, void* stream) {
  HDF_REQUIRE(y && mean && rstd && out && head_w && logits, "hdf_instnorm_apply_head: null pointer");
  HDF_REQUIRE(hdf_instnorm_apply_head_supported(C, ncls), "hdf_instnorm_apply_head: unsupported C=%d ncls=%d", C, ncls);
  HDF_REQUIRE(can_vec<bf16>(y, ldy, C) && can_vec<bf16>(out, ldo, C) && fast_ok(C, 8, V),
              "hdf_instnorm_apply_head: operands must be 16-byte aligned channel-vector rows");
  cudaStream_t s = (cudaStream_t)stream;
  const int rpb = fast_rows_per_block(V, N, 256 / (C / 8));
  const dim3 grid((unsigned)cdiv(V, rpb), N);
  if (ncls <= 2)
    in_apply_head_kernel<2><<<grid, 256, 0, s>>>((const bf16*)y, ldy, mean, rstd, gamma, beta, (bf16*)out, ldo, (int)V, C, relu, rpb, head_w,
                                                 head_b, (bf16*)logits, ncls);
  else
    in_apply_head_kernel<4><<<grid, 256, 0, s>>>((const bf16*)y, ldy, mean, rstd, gamma, beta, (bf16*)out, ldo, (int)V, C, relu, rpb, head_w,
                                                 head_b, (bf16*)logits, ncls);
  HDF_LAUNCH_CHECK("hdf_instnorm_apply_head");
  return HDF_OK;
}

// dy = IN/ReLU backward; also (d)gamma/(d)beta.  s1/s2: [N*C] scratch outputs.
int hdf_instnorm_bwd(int dtype, const void* dout, long long ldd, const void* y, long long ldy, const float* mean,
                     const float* rstd, const float* gamma, const float* beta, void* dy, long long ldo, int N, long long V,
                     int C, int relu, float* s1, float* s2, float* dgamma, float* dbeta, int accumulate_params,
                     void* workspace, size_t ws_bytes, void* stream) {
  HDF_REQUIRE(dout && y && mean && rstd && dy && s1 && s2 && workspace && ws_bytes >= hdf_reduce_workspace(N, V, C),
              "hdf_instnorm_bwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int chunks = 0;
  unsigned* tk = fin_tickets(N, s);
  HDF_DISPATCH_DTYPE(dtype, T, {
    InBwdF<T> f{(const T*)dout, ldd, (const T*)y, ldy, V, mean, rstd, gamma, beta, C, relu};
    int rc = tk ? launch_rowreduce<T>(f, dout, ldd, y, ldy, N, V, C, (double*)workspace, chunks, s, "hdf_instnorm_bwd/reduce",
                                      FinSums{s1, s2, dgamma, dbeta, accumulate_params}, tk)
                : launch_rowreduce<T>(f, dout, ldd, y, ldy, N, V, C, (double*)workspace, chunks, s, "hdf_instnorm_bwd/reduce",
                                      FinNone{}, nullptr);
    if (rc) return rc;
  });
  if (!tk) {
    sums_finalize_kernel<<<cdiv((long long)N * C * 32, 128), 128, 0, s>>>((const double*)workspace, chunks, C, s1, s2, N * C);
    HDF_LAUNCH_CHECK("hdf_instnorm_bwd/finalize");
    if (dgamma || dbeta) {
      in_param_grads_kernel<<<cdiv(C, 128), 128, 0, s>>>(s1, s2, N, C, dgamma, dbeta, accumulate_params);
      HDF_LAUNCH_CHECK("hdf_instnorm_bwd/params");
    }
  }
  HDF_DISPATCH_DTYPE(dtype, T, {
    const bool vec = can_vec<T>(dout, ldd, C) && can_vec<T>(y, ldy, C) && can_vec<T>(dy, ldo, C);
    HDF_VEC_DISPATCH(vec, {
        if (vec && sizeof(T) == 2 && fast_ok(C, VEC, V)) {
          const int rpb = fast_rows_per_block(V, N, 256 / (C / VEC));
          in_bwd_apply_fast_kernel<T, VEC><<<dim3((unsigned)cdiv(V, rpb), N), 256, 0, s>>>((const T*)dout, ldd, (const T*)y, ldy, mean, rstd, gamma, beta, s1, s2, (T*)dy, ldo, (int)V, C, relu, rpb);
        } else {
          long long total = (long long)N * V * (C / VEC);
          in_bwd_apply_kernel<T, VEC><<<grid_for(total), 256, 0, s>>>((const T*)dout, ldd, (const T*)y, ldy, mean, rstd, gamma, beta, s1, s2, (T*)dy, ldo, V, C, relu, total);
        } });
  });
  HDF_LAUNCH_CHECK("hdf_instnorm_bwd/apply");
  return HDF_OK;
}

// out[C] (+)= sum over rows of x[rows, C]
int hdf_colsum(int dtype, const void* x, long long ld, long long rows, int C, float* out, int accumulate, void* workspace,
               size_t ws_bytes, void* stream) {
  HDF_REQUIRE(x && out && workspace && ws_bytes >= hdf_reduce_workspace(1, rows, C), "hdf_colsum: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int chunks = 0;
  unsigned* tk = fin_tickets(1, s);
  HDF_DISPATCH_DTYPE(dtype, T, {
    ColSumF<T> f{(const T*)x, ld, rows};
    int rc = tk ? launch_rowreduce<T>(f, x, ld, nullptr, 0, 1, rows, C, (double*)workspace, chunks, s, "hdf_colsum",
                                      FinColSum{out, accumulate}, tk)
                : launch_rowreduce<T>(f, x, ld, nullptr, 0, 1, rows, C, (double*)workspace, chunks, s, "hdf_colsum", FinNone{}, nullptr);
    if (rc) return rc;
  });
  if (tk) return HDF_OK;
  colsum_finalize_kernel<<<cdiv((long long)C * 32, 128), 128, 0, s>>>((const double*)workspace, chunks, C, out, accumulate);
  HDF_LAUNCH_CHECK("hdf_colsum/finalize");
  return HDF_OK;
}

int hdf_add_(int dtype, void* dst, long long ldd, const void* src, long long lds, long long rows, int C, void* stream) {
  HDF_REQUIRE(dst && src, "hdf_add_: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  HDF_DISPATCH_DTYPE(dtype, T, {
    const bool vec = can_vec<T>(dst, ldd, C) && can_vec<T>(src, lds, C);
    HDF_VEC_DISPATCH(vec, { long long total = rows * (C / VEC); add_kernel<T, VEC><<<grid_for(total), 256, 0, s>>>((T*)dst, ldd, (const T*)src, lds, C, total); });
  });
  HDF_LAUNCH_CHECK("hdf_add_");
  return HDF_OK;
}

int hdf_copy_rows(int dtype, void* dst, long long ldd, const void* src, long long lds, long long rows, int C, void* stream) {
  HDF_REQUIRE(dst && src, "hdf_copy_rows: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  HDF_DISPATCH_DTYPE(dtype, T, {
    const bool vec = can_vec<T>(dst, ldd, C) && can_vec<T>(src, lds, C);
    HDF_VEC_DISPATCH(vec, { long long total = rows * (C / VEC); copy_kernel<T, VEC><<<grid_for(total), 256, 0, s>>>((T*)dst, ldd, (const T*)src, lds, C, total); });
  });
  HDF_LAUNCH_CHECK("hdf_copy_rows");
  return HDF_OK;
}

int hdf_cast_rows_from_f32(int dtype, const float* src, long long lds, void* dst, long long ldd, long long rows, int C,
                           void* stream) {
  HDF_REQUIRE(dst && src, "hdf_cast_rows_from_f32: null pointer");
  const long long total = rows * C;
  HDF_DISPATCH_DTYPE(dtype, T, {
    cast_rows_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(src, lds, (T*)dst, ldd, C, total);
  });
  HDF_LAUNCH_CHECK("hdf_cast_rows_from_f32");
  return HDF_OK;
}

int hdf_cast_rows_to_f32(int dtype, const void* src, long long lds, float* dst, long long ldd, long long rows, int C,
                         void* stream) {
  HDF_REQUIRE(dst && src, "hdf_cast_rows_to_f32: null pointer");
  const long long total = rows * C;
  HDF_DISPATCH_DTYPE(dtype, T, {
    uncast_rows_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)src, lds, dst, ldd, C, total);
  });
  HDF_LAUNCH_CHECK("hdf_cast_rows_to_f32");
  return HDF_OK;
}

// pd = pooling window along depth: 2 = MaxPool3d(2); 1 = MaxPool2d(2) applied to [N, Do, 2Ho, 2Wo, C] slices (flat volumes)
int hdf_maxpool2_fwd_ex(int dtype, const void* x, long long ldx, void* out, long long ldo, int N, int Do, int Ho, int Wo,
                        int C, int pd, void* stream) {
  HDF_REQUIRE(x && out && (pd == 1 || pd == 2), "hdf_maxpool2_fwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  HDF_DISPATCH_DTYPE(dtype, T, {
    const bool vec = can_vec<T>(x, ldx, C) && can_vec<T>(out, ldo, C);
    HDF_VEC_DISPATCH(vec, { maxpool_fwd_kernel<T, VEC><<<(unsigned)((long long)N * Do * Ho), line_block(Wo * (C / VEC)), 0, s>>>((const T*)x, ldx, (T*)out, ldo, Do, Ho, Wo, C, pd); });
  });
  HDF_LAUNCH_CHECK("hdf_maxpool2_fwd");
  return HDF_OK;
}
int hdf_maxpool2_fwd(int dtype, const void* x, long long ldx, void* out, long long ldo, int N, int Do, int Ho, int Wo,
                     int C, void* stream) {
  return hdf_maxpool2_fwd_ex(dtype, x, ldx, out, ldo, N, Do, Ho, Wo, C, 2, stream);
}

int hdf_maxpool2_bwd_ex(int dtype, const void* x, long long ldx, const void* dpool, long long ldp, void* dx, long long lddx,
                        int N, int Do, int Ho, int Wo, int C, int accumulate, int pd, void* stream) {
  HDF_REQUIRE(x && dpool && dx && (pd == 1 || pd == 2), "hdf_maxpool2_bwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  HDF_DISPATCH_DTYPE(dtype, T, {
    const bool vec = can_vec<T>(x, ldx, C) && can_vec<T>(dpool, ldp, C) && can_vec<T>(dx, lddx, C);
    HDF_VEC_DISPATCH(vec, { maxpool_bwd_kernel<T, VEC><<<(unsigned)((long long)N * Do * Ho), line_block(Wo * (C / VEC)), 0, s>>>((const T*)x, ldx, (const T*)dpool, ldp, (T*)dx, lddx, Do, Ho, Wo, C, accumulate, pd); });
  });
  HDF_LAUNCH_CHECK("hdf_maxpool2_bwd");
  return HDF_OK;
}
int hdf_maxpool2_bwd(int dtype, const void* x, long long ldx, const void* dpool, long long ldp, void* dx, long long lddx,
                     int N, int Do, int Ho, int Wo, int C, int accumulate, void* stream) {
  return hdf_maxpool2_bwd_ex(dtype, x, ldx, dpool, ldp, dx, lddx, N, Do, Ho, Wo, C, accumulate, 2, stream);
}

// sd = scale along depth: 2 = trilinear x2; 1 = bilinear x2 of [N, Di, Hi, Wi, C] slices (flat volumes; generic kernel only)
int hdf_upsample2_fwd_ex(int dtype, const void* x, long long ldx, void* out, long long ldo, int N, int Di, int Hi, int Wi,
                         int C, int sd, void* stream) {
  HDF_REQUIRE(x && out && (sd == 1 || sd == 2), "hdf_upsample2_fwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  HDF_DISPATCH_DTYPE(dtype, T, {
    const bool vec = can_vec<T>(x, ldx, C) && can_vec<T>(out, ldo, C);
    if (vec && dtype == HDF_BF16 && sd == 2) {
      upsample2_fwd_cell_kernel<<<(unsigned)((long long)N * (Di + 1) * (Hi + 1)), line_block((Wi + 1) * (C / 8)), 0, s>>>((const bf16*)x, ldx, (bf16*)out, ldo, Di, Hi, Wi, C);
    } else {
      HDF_VEC_DISPATCH(vec, { upsample2_fwd_kernel<T, VEC><<<(unsigned)((long long)N * Di * sd * Hi * 2), line_block(2 * Wi * (C / VEC)), 0, s>>>((const T*)x, ldx, (T*)out, ldo, Di, Hi, Wi, C, sd); });
    }
  });
  HDF_LAUNCH_CHECK("hdf_upsample2_fwd");
  return HDF_OK;
}
int hdf_upsample2_fwd(int dtype, const void* x, long long ldx, void* out, long long ldo, int N, int Di, int Hi, int Wi,
                      int C, void* stream) {
  return hdf_upsample2_fwd_ex(dtype, x, ldx, out, ldo, N, Di, Hi, Wi, C, 2, stream);
}

int hdf_upsample2_bwd_ex(int dtype, const void* dout, long long ldd, void* dx, long long lddx, int N, int Di, int Hi, int Wi,
                         int C, int accumulate, int sd, void* stream) {
  HDF_REQUIRE(dout && dx && (sd == 1 || sd == 2), "hdf_upsample2_bwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  HDF_DISPATCH_DTYPE(dtype, T, {
    const bool vec = can_vec<T>(dout, ldd, C) && can_vec<T>(dx, lddx, C);
    const size_t line_smem = (size_t)2 * Wi * C * sizeof(float);
    // measured slower than the gather form (0.49 vs 0.30 ms at 144^3: two phases, 18 KB + 99 registers per CTA) -> opt-in
    static const bool ups_v2 = getenv("HDF_UPS_BWD_V2") != nullptr;
    static const bool ups_gather = getenv("HDF_UPS_BWD_GATHER") != nullptr;
    if (vec && dtype == HDF_BF16 && !ups_gather && !ups_v2 && sd == 2) {
      const int Db = (Di + 1) / 2, Hb = (Hi + 1) / 2;
      upsample2_bwd_blk_kernel<<<(unsigned)((long long)N * Db * Hb), line_block(Wi * (C / 8)), 0, s>>>((const bf16*)dout, ldd, (bf16*)dx, lddx, Di, Hi, Wi, C, accumulate);
    } else if (vec && dtype == HDF_BF16 && line_smem <= 48 * 1024 && ups_v2 && sd == 2) {
      upsample2_bwd_line_kernel<<<(unsigned)((long long)N * Di * Hi), line_block(2 * Wi * (C / 8)), line_smem, s>>>((const bf16*)dout, ldd, (bf16*)dx, lddx, Di, Hi, Wi, C, accumulate);
    } else {
      HDF_VEC_DISPATCH(vec, { upsample2_bwd_kernel<T, VEC><<<(unsigned)((long long)N * Di * Hi), line_block(Wi * (C / VEC)), 0, s>>>((const T*)dout, ldd, (T*)dx, lddx, Di, Hi, Wi, C, accumulate, sd); });
    }
  });
  HDF_LAUNCH_CHECK("hdf_upsample2_bwd");
  return HDF_OK;
}
int hdf_upsample2_bwd(int dtype, const void* dout, long long ldd, void* dx, long long lddx, int N, int Di, int Hi, int Wi,
                      int C, int accumulate, void* stream) {
  return hdf_upsample2_bwd_ex(dtype, dout, ldd, dx, lddx, N, Di, Hi, Wi, C, accumulate, 2, stream);
}

int hdf_head_fwd(int dtype, const void* a, long long lda, const float* w, const float* b, void* out, int N, long long V,
                 int C, int ncls, void* stream) {
  HDF_REQUIRE(a && w && out && ncls >= 1 && ncls <= MAXCLS, "hdf_head_fwd: bad args (n_cls must be 1..%d)", MAXCLS);
  cudaStream_t s = (cudaStream_t)stream;
  const long long total = (long long)N * V;
  const size_t smem = (size_t)(ncls * C + ncls) * sizeof(float);
  HDF_DISPATCH_DTYPE(dtype, T, {
    const bool vec = can_vec<T>(a, lda, C);
    HDF_VEC_DISPATCH(vec, { head_fwd_kernel<T, VEC><<<grid_for(total, 128), 128, smem, s>>>((const T*)a, lda, w, b, (T*)out, V, C, ncls, total); });
  });
  HDF_LAUNCH_CHECK("hdf_head_fwd");
  return HDF_OK;
}

static int head_chunks_per_n(int N, long long V) {
  long long c = (V + 2047) / 2048, cap = (148 * 4 + N - 1) / N;
  if (c > cap) c = cap;
  return (int)(c < 1 ? 1 : c);
}

size_t hdf_head_bwd_workspace(int N, long long V, int C, int ncls) {
  return (size_t)N * head_chunks_per_n(N, V) * ncls * (C + 1) * sizeof(float);
}

// g: NCDHW [N, ncls, V] (T).  da (+)= g . w ; dw/db (+)= reductions
int hdf_head_bwd(int dtype, const void* g, const void* a, long long lda, const float* w, void* da, long long ldd,
                 float* dw, float* db, int N, long long V, int C, int ncls, int accumulate_da, int accumulate_params,
                 void* workspace, size_t ws_bytes, void* stream) {
  HDF_REQUIRE(g && a && w && da && dw && workspace && ncls >= 1 && ncls <= MAXCLS, "hdf_head_bwd: bad args");
  HDF_REQUIRE(ws_bytes >= hdf_head_bwd_workspace(N, V, C, ncls), "hdf_head_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  HDF_REQUIRE(V < (1ll << 31), "hdf_head_bwd: V too large");
  const long long total = (long long)N * V;
  int cpn = head_chunks_per_n(N, V);
  const int rpc = cdiv(V, cpn);
  cpn = cdiv(V, rpc);
  const int chunks = cpn * N;
  HDF_DISPATCH_DTYPE(dtype, T, {
    {
      const bool vecw = can_vec<T>(a, lda, C);
      HDF_REQUIRE(vecw ? (C / VecOf<T>::value <= 256) : (C <= 256), "hdf_head_bwd: C=%d too large", C);
      HDF_VEC_DISPATCH(vecw, {
        const int lanes = 256 / (C / VEC);
        const size_t smem = (size_t)lanes * ncls * (C + 1) * sizeof(float);
        if (ncls <= 2) launch_head_wgrad<T, VEC, 2>(dim3(cpn, N), smem, s, (const T*)g, (const T*)a, lda, (int)V, C, ncls, rpc, (float*)workspace);
        else if (ncls <= 4) launch_head_wgrad<T, VEC, 4>(dim3(cpn, N), smem, s, (const T*)g, (const T*)a, lda, (int)V, C, ncls, rpc, (float*)workspace);
        else launch_head_wgrad<T, VEC, 8>(dim3(cpn, N), smem, s, (const T*)g, (const T*)a, lda, (int)V, C, ncls, rpc, (float*)workspace);
      });
    }
    HDF_LAUNCH_CHECK("hdf_head_bwd/wgrad");
    head_wgrad_finalize_kernel<<<cdiv(ncls * (C + 1), 128), 128, 0, s>>>((const float*)workspace, chunks, C, ncls, dw, db, accumulate_params);
    HDF_LAUNCH_CHECK("hdf_head_bwd/finalize");
    const bool vec = can_vec<T>(da, ldd, C);
    const size_t smem2 = (size_t)ncls * C * sizeof(float);
    HDF_VEC_DISPATCH(vec, {
      if (vec && sizeof(T) == 2 && fast_ok(C, VEC, V)) {
        const int rpb = fast_rows_per_block(V, N, 256 / (C / VEC));
        const dim3 gr((unsigned)cdiv(V, rpb), N);
        if (ncls <= 2) head_dgrad_fast_kernel<T, VEC, 2><<<gr, 256, 0, s>>>((const T*)g, w, (T*)da, ldd, (int)V, C, ncls, accumulate_da, rpb);
        else if (ncls <= 4) head_dgrad_fast_kernel<T, VEC, 4><<<gr, 256, 0, s>>>((const T*)g, w, (T*)da, ldd, (int)V, C, ncls, accumulate_da, rpb);
        else head_dgrad_fast_kernel<T, VEC, 8><<<gr, 256, 0, s>>>((const T*)g, w, (T*)da, ldd, (int)V, C, ncls, accumulate_da, rpb);
      } else {
        head_dgrad_kernel<T, VEC><<<grid_for(total, 128), 128, smem2, s>>>((const T*)g, w, (T*)da, ldd, V, C, ncls, accumulate_da, total);
      } });
  });
  HDF_LAUNCH_CHECK("hdf_head_bwd/dgrad");
  return HDF_OK;
}

int hdf_ncdhw_to_cl(int dtype, const float* x, void* out, long long ldo, int N, int C, long long V, void* stream) {
  HDF_REQUIRE(x && out, "hdf_ncdhw_to_cl: null pointer");
  const long long total = (long long)N * V;
  HDF_DISPATCH_DTYPE(dtype, T, {
    ncdhw_to_cl_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, (T*)out, ldo, V, C, total);
  });
  HDF_LAUNCH_CHECK("hdf_ncdhw_to_cl");
  return HDF_OK;
}

}  // extern "C"
