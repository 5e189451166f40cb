// Tensor-core forward of one inner layer of the densely connected transformer (DCT), bf16 path
// (models/HDenseFormer.py:91-98: Linear_l -> +Attn(LN) -> +FF(LN) -> append FF(LN(x)); :47-75 Dense_Attention; :33-44
// DenseForward).  Two kernels per layer instead of five:
//   tok_a_fwd : h0 = cat(features) Wl^T + bl ; n1 = LN1(h0) ; qkv = n1 Wqkv^T          (row-local, 16 rows per block)
//   tok_c_fwd : o = softmax(q k^T / 2) v  for the 8 heads of a 16-query tile (one warp per head), then the row-local
//               chain h1 = drop(o Wo^T + bo) + h0 ; h2 = FF(LN2(h1)) + h1 ; feature = FF(LN2(h2))
// All contractions run on the tensor cores (mma.sync.m16n8k16 bf16 for the Linears with fp32 accumulation; tf32
// m16n8k4 / m16n8k8 for Q K^T and P V, whose K = head_dim = 4 matches the tf32 K atom exactly -- no zero padding); softmax
// statistics stay in fp32 registers and are exchanged with warp shuffles; scores are never materialised.  The tensors the
// backward pass needs are written in fp32, in the layout of the first-generation kernels (csrc/dct.cu), so the same
// backward kernels consume them; dropout masks use the same counter-based generator and element indices.
// The fp32 exact path keeps the SIMT kernels (no tensor-core rounding there).
#include "common.cuh"

namespace {

constexpr int TG = 32;    // growth rate = token width inside a layer
constexpr int TH_ = 64;   // MLP hidden width
constexpr int HEADS = 8;  // heads of dim 4

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t f2tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
// D += A(16x16 bf16, row) * B(16x8 bf16, col)
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D += A(16x4 tf32) * B(4x8 tf32)
__device__ __forceinline__ void mma_tf32_k4(float* c, uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(b0));
}
// D += A(16x8 tf32) * B(8x8 tf32)
__device__ __forceinline__ void mma_tf32_k8(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }

// Fragment conventions (PTX ISA, m16n8k16): g = lane / 4, t = lane % 4.
//   A regs: a0 = (row g, k 2t..2t+1), a1 = (row g+8, same k), a2 = (row g, k 2t+8..2t+9), a3 = (row g+8, k 2t+8..)
//   B regs: b0 = (k 2t..2t+1, n g), b1 = (k 2t+8..2t+9, n g)        C regs: c0,c1 = (row g, n 2t..2t+1), c2,c3 = (row g+8, ..)
// so the C fragments of output tiles 2s and 2s+1 are exactly the A fragment of k-step s of the next GEMM.
//
// A fragments of k-step s from a row-major fp32 tile in shared memory (LD = 72 floats: conflict-free 64-bit loads)
constexpr int SLD = 72;
__device__ __forceinline__ void a_from_smem(const float (*x)[SLD], int s, int g, int t, uint32_t* a) {
  const float2 p0 = *reinterpret_cast<const float2*>(&x[g][16 * s + 2 * t]), p1 = *reinterpret_cast<const float2*>(&x[g + 8][16 * s + 2 * t]);
  const float2 p2 = *reinterpret_cast<const float2*>(&x[g][16 * s + 2 * t + 8]), p3 = *reinterpret_cast<const float2*>(&x[g + 8][16 * s + 2 * t + 8]);
  a[0] = pack_bf16x2(p0.x, p0.y); a[1] = pack_bf16x2(p1.x, p1.y); a[2] = pack_bf16x2(p2.x, p2.y); a[3] = pack_bf16x2(p3.x, p3.y);
}
// B fragments (bf16 pairs) of output tile j, k-steps 0..KS-1, from a torch Linear weight [out][ldw] in global memory
template <int KS>
__device__ __forceinline__ void load_w_tile(const float* __restrict__ W, int ldw, int j, int g, int t, uint32_t (*b)[2]) {
#pragma unroll
  for (int s = 0; s < KS; ++s) {
    const float* w = W + (size_t)(8 * j + g) * ldw + 16 * s + 2 * t;
    const float2 w0 = *reinterpret_cast<const float2*>(w), w1 = *reinterpret_cast<const float2*>(w + 8);
    b[s][0] = pack_bf16x2(w0.x, w0.y); b[s][1] = pack_bf16x2(w1.x, w1.y);
  }
}

// ---------------------------------------------------------------------------------------------------- layer head
struct TokAParams {
  const float* F; long long ldf; int Cl;
  const float* Wl; const float* bl; const float* gm; const float* bt; const float* Wqkv;
  float* h0; float* n1; float* m1; float* r1; float* qkv;
  int R;
};

// 16 rows per block, 4 warps: the K = Cl contraction of Linear_l is split over the warps (a single warp would walk 8..14
// dependent global round trips), partial sums meet in shared memory, LayerNorm runs with one row per warp pass (lane =
// column), and the 12 output tiles of to_qkv are shared out three per warp.
__global__ void __launch_bounds__(128) tok_a_fwd_kernel(const TokAParams q) {
  __shared__ float part[4][16][TG + 1];
  __shared__ __align__(16) float xn[16][SLD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long rbase = (long long)blockIdx.x * 16;
  const long long r0 = rbase + g, r1 = r0 + 8;
  const float* x0 = q.F + (r0 < q.R ? r0 : q.R - 1) * q.ldf;
  const float* x1 = q.F + (r1 < q.R ? r1 : q.R - 1) * q.ldf;
  // to_qkv weight tiles of this warp (output tiles 3w .. 3w+2), requested before anything else
  uint32_t wq[3][2][2];
#pragma unroll
  for (int i = 0; i < 3; ++i) load_w_tile<2>(q.Wqkv, TG, 3 * warp + i, g, t, wq[i]);
  // ---- partial h0 over k-steps s = warp, warp+4, ...
  float h[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { h[j][0] = h[j][1] = h[j][2] = h[j][3] = 0.f; }
  const int ks = q.Cl / 16;
  for (int s = warp; s < ks; s += 4) {
    const int k = 16 * s + 2 * t;
    const float2 p0 = *reinterpret_cast<const float2*>(x0 + k), p1 = *reinterpret_cast<const float2*>(x1 + k);
    const float2 p2 = *reinterpret_cast<const float2*>(x0 + k + 8), p3 = *reinterpret_cast<const float2*>(x1 + k + 8);
    const uint32_t a[4] = {pack_bf16x2(p0.x, p0.y), pack_bf16x2(p1.x, p1.y), pack_bf16x2(p2.x, p2.y), pack_bf16x2(p3.x, p3.y)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float* w = q.Wl + (size_t)(8 * j + g) * q.Cl + k;
      const float2 w0 = *reinterpret_cast<const float2*>(w), w1 = *reinterpret_cast<const float2*>(w + 8);
      mma_bf16(h[j], a, pack_bf16x2(w0.x, w0.y), pack_bf16x2(w1.x, w1.y));
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    part[warp][g][8 * j + 2 * t] = h[j][0]; part[warp][g][8 * j + 2 * t + 1] = h[j][1];
    part[warp][g + 8][8 * j + 2 * t] = h[j][2]; part[warp][g + 8][8 * j + 2 * t + 1] = h[j][3];
  }
  __syncthreads();
  // ---- h0 = sum of partials + bias ; n1 = LN1(h0): warp w owns rows 4w .. 4w+3, lane = column
  {
    const float bl = q.bl[lane], gmv = q.gm[lane], btv = q.bt[lane];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = 4 * warp + i;
      const float v = part[0][row][lane] + part[1][row][lane] + part[2][row][lane] + part[3][row][lane] + bl;
      const float mu = warp_sum(v) * (1.f / TG);
      const float d = v - mu;
      const float rs = rsqrtf(warp_sum(d * d) * (1.f / TG) + 1e-5f);
      const float nv = d * rs * gmv + btv;
      xn[row][lane] = nv;
      const long long r = rbase + row;
      if (r < q.R) {
        q.h0[r * TG + lane] = v;
        q.n1[r * TG + lane] = nv;
        if (lane == 0) { q.m1[r] = mu; q.r1[r] = rs; }
      }
    }
  }
  __syncthreads();
  // ---- qkv = n1 Wqkv^T (no bias): output tiles 3w .. 3w+2
  uint32_t a[2][4];
  a_from_smem(xn, 0, g, t, a[0]);
  a_from_smem(xn, 1, g, t, a[1]);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    mma_bf16(o, a[0], wq[i][0][0], wq[i][0][1]);
    mma_bf16(o, a[1], wq[i][1][0], wq[i][1][1]);
    const int c = 8 * (3 * warp + i) + 2 * t;
    if (r0 < q.R) *reinterpret_cast<float2*>(q.qkv + r0 * (3 * TG) + c) = make_float2(o[0], o[1]);
    if (r1 < q.R) *reinterpret_cast<float2*>(q.qkv + r1 * (3 * TG) + c) = make_float2(o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------------- attention + chain
struct TokCParams {
  const float* qkv; const float* h0;
  float* o; float* lse;
  float* h1; float* n2; float* z1; float* f1; float* h2; float* n3; float* z1b; float* g1;
  float* m2; float* r2; float* m3; float* r3;
  float* fout; long long ldf;
  const float* Wo; const float* bo; const float* gm; const float* bt; const float* W1; const float* b1; const float* W2; const float* b2;
  int B, N;
  float scale, p;
  const unsigned long long* seed_ptr; unsigned long long seed; unsigned ida, idb, idc, idd, ide;
};

// 16 query rows per block, 8 warps.  Phase 1: warp = head.  Phase 2: the row-local chain with the activation tile in shared
// memory; every GEMM's output tiles are shared out over the warps (4 tiles -> warps 0-3, 8 tiles -> all), LayerNorm runs with
// two rows per warp (lane = column).  Each warp requests its weight fragments of the whole chain before the attention loop,
// so that the chain never waits for global memory: a single warp walking the chain was latency-bound (52 us per launch).
__global__ void __launch_bounds__(256) tok_c_fwd_kernel(const TokCParams q) {
  __shared__ __align__(16) float xs[16][SLD];       // o -> n2 -> n3 (A source of Wo / W1)
  __shared__ __align__(16) float fs[16][SLD];       // f1 / g1 (A source of W2)
  __shared__ float hs[16][TG + 1];    // residual stream h1 -> h2
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * 16;                       // first query of the tile (within the sample)
  const int N = q.N;
  const float* base = q.qkv + (long long)b * N * (3 * TG);
  const bool ok0 = i0 + g < N, ok1 = i0 + g + 8 < N;
  const long long r0 = (long long)b * N + (ok0 ? i0 + g : N - 1), r1 = (long long)b * N + (ok1 ? i0 + g + 8 : N - 1);
  // ---- chain weights of this warp: Wo / W2 output tile (warp & 3), W1 output tile warp
  const int j4 = warp & 3;
  uint32_t wo[2][2], w1[2][2], w2[4][2];
  load_w_tile<2>(q.Wo, TG, j4, g, t, wo);
  load_w_tile<2>(q.W1, TG, warp, g, t, w1);
  load_w_tile<4>(q.W2, TH_, j4, g, t, w2);
  const int c4 = 8 * j4 + 2 * t, c8 = 8 * warp + 2 * t;
  const float2 bo2 = *reinterpret_cast<const float2*>(q.bo + c4), b12 = *reinterpret_cast<const float2*>(q.b1 + c8),
               b22 = *reinterpret_cast<const float2*>(q.b2 + c4);
  const float2 h00 = *reinterpret_cast<const float2*>(q.h0 + r0 * TG + c4), h01 = *reinterpret_cast<const float2*>(q.h0 + r1 * TG + c4);
  const float gml = q.gm[lane], btl = q.bt[lane];
  {
    // ===== attention of head `warp` for queries i0 .. i0+15: S = (q * scale) k^T with m16n8k4 (K = head_dim = 4), online
    // softmax over blocks of 32 keys, O += P V with m16n8k8 (N = 8: 4 value dims + 4 zero columns).  In the C fragment a
    // thread holds keys (2t, 2t+1) of every 8-key tile; feeding them to the A fragment positions (t, t+4) of the P V MMA
    // just renames the keys, so the B fragment takes V rows (2t, 2t+1): no shuffles between the two MMAs.  The K / V
    // values of the next 32 keys are requested before the current block is processed.
    const int h = warp;
    const int qa = min(i0 + g, N - 1), qb = min(i0 + g + 8, N - 1);
    const uint32_t a0 = f2tf32(base[(long long)qa * (3 * TG) + 4 * h + t] * q.scale);
    const uint32_t a1 = f2tf32(base[(long long)qb * (3 * TG) + 4 * h + t] * q.scale);
    const float* kp = base + TG + 4 * h + t;                          // K[key][dim t]
    const float* vp = base + 2 * TG + 4 * h + (g & 3);                // V[key][dim g] (lanes g >= 4 feed zero columns)
    const bool vlane = g < 4;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float oc[4] = {0.f, 0.f, 0.f, 0.f};
    float kn[4], vn0[4], vn1[4];
    auto fetch = [&](int kb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = kb + 8 * j + g, kv = kb + 8 * j + 2 * t;
        kn[j] = key < N ? kp[(long long)key * (3 * TG)] : 0.f;
        vn0[j] = (vlane && kv < N) ? vp[(long long)kv * (3 * TG)] : 0.f;
        vn1[j] = (vlane && kv + 1 < N) ? vp[(long long)(kv + 1) * (3 * TG)] : 0.f;
      }
    };
    fetch(0);
    for (int kb = 0; kb < N; kb += 32) {
      uint32_t bk[4], vb0[4], vb1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { bk[j] = f2tf32(kn[j]); vb0[j] = f2tf32(vn0[j]); vb1[j] = f2tf32(vn1[j]); }
      if (kb + 32 < N) fetch(kb + 32);
      float s[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kv = kb + 8 * j + 2 * t;
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        mma_tf32_k4(s[j], a0, a1, bk[j]);
        if (kv >= N) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (kv + 1 >= N) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
      float mx0 = fmaxf(fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1])), fmaxf(fmaxf(s[2][0], s[2][1]), fmaxf(s[3][0], s[3][1])));
      float mx1 = fmaxf(fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3])), fmaxf(fmaxf(s[2][2], s[2][3]), fmaxf(s[3][2], s[3][3])));
      const float mn0 = fmaxf(m0, quad_max(mx0)), mn1 = fmaxf(m1, quad_max(mx1));     // finite: every block has >= 1 valid key
      const float sc0 = __expf(m0 - mn0), sc1 = __expf(m1 - mn1);
      m0 = mn0; m1 = mn1;
      l0 *= sc0; l1 *= sc1;
      oc[0] *= sc0; oc[1] *= sc0; oc[2] *= sc1; oc[3] *= sc1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p0 = __expf(s[j][0] - mn0), p1 = __expf(s[j][1] - mn0), p2 = __expf(s[j][2] - mn1), p3 = __expf(s[j][3] - mn1);
        l0 += p0 + p1; l1 += p2 + p3;
        mma_tf32_k8(oc, f2tf32(p0), f2tf32(p2), f2tf32(p1), f2tf32(p3), vb0[j], vb1[j]);
      }
    }
    l0 = quad_sum(l0); l1 = quad_sum(l1);
    const float i0l = 1.f / l0, i1l = 1.f / l1;
    if (t < 2) {      // columns 2t, 2t+1 < 4 are the head's four value dims
      const int c = 4 * h + 2 * t;
      xs[g][c] = oc[0] * i0l; xs[g][c + 1] = oc[1] * i0l;
      xs[g + 8][c] = oc[2] * i1l; xs[g + 8][c + 1] = oc[3] * i1l;
      if (ok0) *reinterpret_cast<float2*>(q.o + r0 * TG + c) = make_float2(oc[0] * i0l, oc[1] * i0l);
      if (ok1) *reinterpret_cast<float2*>(q.o + r1 * TG + c) = make_float2(oc[2] * i1l, oc[3] * i1l);
    }
    if (t == 0) {
      if (ok0) q.lse[((long long)b * HEADS + h) * N + i0 + g] = m0 + __logf(l0);
      if (ok1) q.lse[((long long)b * HEADS + h) * N + i0 + g + 8] = m1 + __logf(l1);
    }
  }
  __syncthreads();
  // ===== row-local chain on the 16 rows of the tile
  const unsigned long long seed = q.seed + (q.seed_ptr ? *q.seed_ptr : 0ull);
  uint32_t a[4][4];
  // h1 = drop_a(o Wo^T + bo) + h0        (warps 0-3: output tile j4)
  if (warp < 4) {
    a_from_smem(xs, 0, g, t, a[0]);
    a_from_smem(xs, 1, g, t, a[1]);
    float c[4] = {bo2.x, bo2.y, bo2.x, bo2.y};
    mma_bf16(c, a[0], wo[0][0], wo[0][1]);
    mma_bf16(c, a[1], wo[1][0], wo[1][1]);
    c[0] = c[0] * hdf_dropout_scale(seed, q.ida, (unsigned long long)r0 * TG + c4, q.p) + h00.x;
    c[1] = c[1] * hdf_dropout_scale(seed, q.ida, (unsigned long long)r0 * TG + c4 + 1, q.p) + h00.y;
    c[2] = c[2] * hdf_dropout_scale(seed, q.ida, (unsigned long long)r1 * TG + c4, q.p) + h01.x;
    c[3] = c[3] * hdf_dropout_scale(seed, q.ida, (unsigned long long)r1 * TG + c4 + 1, q.p) + h01.y;
    hs[g][c4] = c[0]; hs[g][c4 + 1] = c[1]; hs[g + 8][c4] = c[2]; hs[g + 8][c4 + 1] = c[3];
    if (ok0) *reinterpret_cast<float2*>(q.h1 + r0 * TG + c4) = make_float2(c[0], c[1]);
    if (ok1) *reinterpret_cast<float2*>(q.h1 + r1 * TG + c4) = make_float2(c[2], c[3]);
  }
  __syncthreads();
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    // ---- n = LN2(h): warp w owns rows 2w, 2w+1, lane = column
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = 2 * warp + i;
      const float v = hs[row][lane];
      const float mu = warp_sum(v) * (1.f / TG);
      const float d = v - mu;
      const float rs = rsqrtf(warp_sum(d * d) * (1.f / TG) + 1e-5f);
      const float nv = d * rs * gml + btl;
      xs[row][lane] = nv;
      if (i0 + row < N) {
        const long long r = (long long)b * N + i0 + row;
        (pass == 0 ? q.n2 : q.n3)[r * TG + lane] = nv;
        if (lane == 0) { (pass == 0 ? q.m2 : q.m3)[r] = mu; (pass == 0 ? q.r2 : q.r3)[r] = rs; }
      }
    }
    __syncthreads();
    // ---- z = n W1^T + b1 (saved) ; f = drop(gelu(z)) (saved)       (all 8 warps: output tile `warp`)
    {
      a_from_smem(xs, 0, g, t, a[0]);
      a_from_smem(xs, 1, g, t, a[1]);
      float c[4] = {b12.x, b12.y, b12.x, b12.y};
      mma_bf16(c, a[0], w1[0][0], w1[0][1]);
      mma_bf16(c, a[1], w1[1][0], w1[1][1]);
      float* zsave = pass == 0 ? q.z1 : q.z1b;
      float* fsave = pass == 0 ? q.f1 : q.g1;
      const unsigned id1 = pass == 0 ? q.idb : q.idd;
      if (ok0) *reinterpret_cast<float2*>(zsave + r0 * TH_ + c8) = make_float2(c[0], c[1]);
      if (ok1) *reinterpret_cast<float2*>(zsave + r1 * TH_ + c8) = make_float2(c[2], c[3]);
      c[0] = gelu_f(c[0]) * hdf_dropout_scale(seed, id1, (unsigned long long)r0 * TH_ + c8, q.p);
      c[1] = gelu_f(c[1]) * hdf_dropout_scale(seed, id1, (unsigned long long)r0 * TH_ + c8 + 1, q.p);
      c[2] = gelu_f(c[2]) * hdf_dropout_scale(seed, id1, (unsigned long long)r1 * TH_ + c8, q.p);
      c[3] = gelu_f(c[3]) * hdf_dropout_scale(seed, id1, (unsigned long long)r1 * TH_ + c8 + 1, q.p);
      fs[g][c8] = c[0]; fs[g][c8 + 1] = c[1]; fs[g + 8][c8] = c[2]; fs[g + 8][c8 + 1] = c[3];
      if (ok0) *reinterpret_cast<float2*>(fsave + r0 * TH_ + c8) = make_float2(c[0], c[1]);
      if (ok1) *reinterpret_cast<float2*>(fsave + r1 * TH_ + c8) = make_float2(c[2], c[3]);
    }
    __syncthreads();
    // ---- y = drop(f W2^T + b2) ; pass 0: h2 = y + h1 ; pass 1: appended feature        (warps 0-3: output tile j4)
    if (warp < 4) {
#pragma unroll
      for (int s = 0; s < 4; ++s) a_from_smem(fs, s, g, t, a[s]);
      float c[4] = {b22.x, b22.y, b22.x, b22.y};
#pragma unroll
      for (int s = 0; s < 4; ++s) mma_bf16(c, a[s], w2[s][0], w2[s][1]);
      const unsigned id2 = pass == 0 ? q.idc : q.ide;
      c[0] *= hdf_dropout_scale(seed, id2, (unsigned long long)r0 * TG + c4, q.p);
      c[1] *= hdf_dropout_scale(seed, id2, (unsigned long long)r0 * TG + c4 + 1, q.p);
      c[2] *= hdf_dropout_scale(seed, id2, (unsigned long long)r1 * TG + c4, q.p);
      c[3] *= hdf_dropout_scale(seed, id2, (unsigned long long)r1 * TG + c4 + 1, q.p);
      if (pass == 0) {
        c[0] += hs[g][c4]; c[1] += hs[g][c4 + 1]; c[2] += hs[g + 8][c4]; c[3] += hs[g + 8][c4 + 1];
        hs[g][c4] = c[0]; hs[g][c4 + 1] = c[1]; hs[g + 8][c4] = c[2]; hs[g + 8][c4 + 1] = c[3];
        if (ok0) *reinterpret_cast<float2*>(q.h2 + r0 * TG + c4) = make_float2(c[0], c[1]);
        if (ok1) *reinterpret_cast<float2*>(q.h2 + r1 * TG + c4) = make_float2(c[2], c[3]);
      } else {
        if (ok0) *reinterpret_cast<float2*>(q.fout + r0 * q.ldf + c4) = make_float2(c[0], c[1]);
        if (ok1) *reinterpret_cast<float2*>(q.fout + r1 * q.ldf + c4) = make_float2(c[2], c[3]);
      }
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" {

// tensor-core version of hdf_dct_a_fwd (same operands): h0 = F[:, :Cl] Wl^T + bl ; n1 = LN1(h0) ; qkv = n1 Wqkv^T
int hdf_tok_a_fwd(const float* F, long long ldf, int Cl, const float* Wl, const float* bl, const float* gm, const float* bt,
                  const float* Wqkv, float* h0, float* n1, float* m1, float* r1, float* qkv, int R, void* stream) {
  HDF_REQUIRE(F && Wl && bl && gm && bt && Wqkv && h0 && n1 && m1 && r1 && qkv && R > 0, "hdf_tok_a_fwd: null pointer");
  HDF_REQUIRE(Cl % 16 == 0 && Cl >= 16 && ldf % 2 == 0, "hdf_tok_a_fwd: Cl=%d must be a multiple of 16", Cl);
  TokAParams q{F, ldf, Cl, Wl, bl, gm, bt, Wqkv, h0, n1, m1, r1, qkv, R};
  tok_a_fwd_kernel<<<cdiv(R, 16), 128, 0, (cudaStream_t)stream>>>(q);
  HDF_LAUNCH_CHECK("hdf_tok_a_fwd");
  return HDF_OK;
}

// attention (8 heads of dim 4) fused with the post-attention chain of hdf_dct_c_fwd; writes o / lse for the backward pass
int hdf_tok_c_fwd(const float* qkv, const float* h0, float* o, float* lse, float* h1, float* n2, float* z1, float* f1, float* h2,
                  float* n3, float* z1b, float* g1, float* m2, float* r2, float* m3, float* r3, float* fout, long long ldf,
                  const float* Wo, const float* bo, const float* gm, const float* bt, const float* W1, const float* b1, const float* W2,
                  const float* b2, int B, int N, float scale, float p, const unsigned long long* seed_ptr, unsigned long long seed,
                  unsigned ida, unsigned idb, unsigned idc, unsigned idd, unsigned ide, void* stream) {
  HDF_REQUIRE(qkv && h0 && o && lse && h1 && n2 && z1 && f1 && h2 && n3 && z1b && g1 && m2 && r2 && m3 && r3 && fout && Wo && bo && gm &&
                  bt && W1 && b1 && W2 && b2 && B > 0 && N > 0, "hdf_tok_c_fwd: null pointer");
  HDF_REQUIRE(ldf % 2 == 0, "hdf_tok_c_fwd: feature stride must be even");
  TokCParams q{qkv, h0, o, lse, h1, n2, z1, f1, h2, n3, z1b, g1, m2, r2, m3, r3, fout, ldf, Wo, bo, gm, bt, W1, b1, W2, b2, B, N,
               scale, p, seed_ptr, seed, ida, idb, idc, idd, ide};
  dim3 grid(cdiv(N, 16), B);
  tok_c_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(q);
  HDF_LAUNCH_CHECK("hdf_tok_c_fwd");
  return HDF_OK;
}

}  // extern "C"
