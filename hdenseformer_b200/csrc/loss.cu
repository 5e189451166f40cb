// Dice + cross-entropy deep-supervision loss (loss/combine_loss.py:25-35,72-79; loss/dice_loss.py:26-41,
// 70-87; loss/cross_entropy.py:10-22) as one reduction kernel + one gradient kernel per level, and the
// sliding-window softmax accumulate / count-normalise / argmax (trainer.py:560-582).
// logits are NCDHW [B, C, V] (f32 or bf16); the one-hot target is the full-resolution fp32 tensor,
// sampled with stride 2^level (nearest-neighbour resize of F.interpolate, SURVEY 2.1 K8').
#include "common.cuh"

namespace {

constexpr int MAXCLS = 8;

struct LevelGeom {
  int Dl, Hl, Wl;      // level dims
  int stride;          // 2^level
  int dstride;         // stride along depth: = stride, or 1 for flat volumes (2-D model: [B, C, 1, H, W])
  long long Vfull;     // full-res voxels per channel
  int Hf, Wf;          // full-res H, W
};

__device__ __forceinline__ long long target_index(const LevelGeom& g, long long v) {
  const int w = v % g.Wl;
  const long long r = v / g.Wl;
  const int h = r % g.Hl;
  const int d = r / g.Hl;
  return ((long long)(d * g.dstride) * g.Hf + h * g.stride) * g.Wf + w * g.stride;
}

// sums layout (double): [B][C][3] = inter, psum, tsum ; then [2] = ce_num, ce_den
template <typename T>
__global__ void __launch_bounds__(256) loss_reduce_kernel(const T* __restrict__ logits, const float* __restrict__ target,
                                                         const float* __restrict__ cw, LevelGeom g, int C, long long V,
                                                         int rows_per_chunk, double* __restrict__ sums, int B) {
  const int b = blockIdx.y;
  const long long v0 = (long long)blockIdx.x * rows_per_chunk;
  const long long v1 = min(V, v0 + rows_per_chunk);
  float inter[MAXCLS], ps[MAXCLS], ts[MAXCLS];
#pragma unroll
  for (int k = 0; k < MAXCLS; ++k) inter[k] = ps[k] = ts[k] = 0.f;
  float cen = 0.f, ced = 0.f;
  const T* lb = logits + (long long)b * C * V;
  const float* tb = target + (long long)b * C * g.Vfull;
  for (long long v = v0 + threadIdx.x; v < v1; v += 256) {
    float x[MAXCLS], t[MAXCLS];
    const long long ti = target_index(g, v);
    float m = -INFINITY, tm = -INFINITY;
    int am = 0;
#pragma unroll
    for (int k = 0; k < MAXCLS; ++k) {
      if (k < C) {
        x[k] = to_f(lb[(long long)k * V + v]);
        t[k] = tb[(long long)k * g.Vfull + ti];
        m = fmaxf(m, x[k]);
        if (t[k] > tm) { tm = t[k]; am = k; }
      }
    }
    float S = 0.f, xt = 0.f;
#pragma unroll
    for (int k = 0; k < MAXCLS; ++k)
      if (k < C) { if (k == am) xt = x[k] - m; x[k] = __expf(x[k] - m); S += x[k]; }
    const float inv = 1.f / S;
#pragma unroll
    for (int k = 0; k < MAXCLS; ++k) {
      if (k < C) {
        const float p = x[k] * inv;
        inter[k] += p * t[k];
        ps[k] += p;
        ts[k] += t[k];
      }
    }
    const float w = cw ? cw[am] : 1.f;
    cen += -w * (xt - logf(S));
    ced += w;
  }
  __shared__ float red[8][3 * MAXCLS + 2];
  const int wid = threadIdx.x / 32, lane = threadIdx.x % 32;
#pragma unroll
  for (int k = 0; k < MAXCLS; ++k) {
    const float a = warp_sum(inter[k]), bq = warp_sum(ps[k]), c = warp_sum(ts[k]);
    if (lane == 0) { red[wid][3 * k] = a; red[wid][3 * k + 1] = bq; red[wid][3 * k + 2] = c; }
  }
  cen = warp_sum(cen);
  ced = warp_sum(ced);
  if (lane == 0) { red[wid][3 * MAXCLS] = cen; red[wid][3 * MAXCLS + 1] = ced; }
  __syncthreads();
  if (threadIdx.x < 3 * MAXCLS + 2) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += (double)red[w][threadIdx.x];
    const int i = threadIdx.x;
    if (i < 3 * MAXCLS) {
      const int k = i / 3;
      if (k < C) atomicAdd(&sums[((long long)b * C + k) * 3 + i % 3], s);
    } else {
      atomicAdd(&sums[(long long)B * C * 3 + (i - 3 * MAXCLS)], s);
    }
  }
}

// level loss -> out_level[0]; total[0] += level_weight * level loss
__global__ void loss_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ cw, int B, int C,
                                     int ignore_index, int has_ignore, float smooth, float level_weight, float ce_w,
                                     float dice_w, float* __restrict__ out_level, float* __restrict__ total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double* ce = sums + (long long)B * C * 3;
  const double cel = ce[0] / ce[1];
  double dice = 0.0;
  for (int c = 0; c < C; ++c) {
    if (has_ignore && c == ignore_index) continue;
    double l = 0.0;
    for (int b = 0; b < B; ++b) {
      const double I = sums[((long long)b * C + c) * 3], U = sums[((long long)b * C + c) * 3 + 1] + sums[((long long)b * C + c) * 3 + 2];
      l += 1.0 - (2.0 * I + smooth) / (U + smooth);
    }
    l /= B;
    if (cw) l *= cw[c];
    dice += l;
  }
  dice /= has_ignore ? (C - 1) : C;
  const double loss = ce_w * cel + dice_w * dice;
  out_level[0] = (float)loss;
  out_level[1] = (float)cel;
  out_level[2] = (float)dice;
  total[0] += level_weight * (float)loss;
}

template <typename T>
__global__ void __launch_bounds__(256) loss_grad_kernel(const T* __restrict__ logits, const float* __restrict__ target,
                                                       const float* __restrict__ cw, LevelGeom g, int C, long long V,
                                                       const double* __restrict__ sums, int B, int ignore_index,
                                                       int has_ignore, float smooth, float level_weight, float ce_w,
                                                       float dice_w, const float* __restrict__ gout,
                                                       T* __restrict__ dlogits) {
  const int b = blockIdx.y;
  const float up = level_weight * (gout ? gout[0] : 1.f);
  __shared__ float sA[MAXCLS], sB[MAXCLS];  // dL/dp_c = sA[c] * t + sB[c]
  if (threadIdx.x < MAXCLS) {
    const int c = threadIdx.x;
    float a = 0.f, bb = 0.f;
    if (c < C && !(has_ignore && c == ignore_index)) {
      const double I = sums[((long long)b * C + c) * 3], U = sums[((long long)b * C + c) * 3 + 1] + sums[((long long)b * C + c) * 3 + 2];
      const double coef = (double)up * dice_w * (cw ? cw[c] : 1.f) / ((has_ignore ? (C - 1) : C) * (double)B);
      const double den = (U + smooth);
      // d/dp [1 - (2I+s)/(U+s)] with dI/dp = t, dU/dp = 1:  -(2 t (U+s) - (2I+s)) / (U+s)^2
      a = (float)(coef * (-2.0 / den));
      bb = (float)(coef * ((2.0 * I + smooth) / (den * den)));
    }
    sA[c] = a; sB[c] = bb;
  }
  __syncthreads();
  const float ce_scale = ce_w * up / (float)sums[(long long)B * C * 3 + 1];
  const T* lb = logits + (long long)b * C * V;
  T* gb = dlogits + (long long)b * C * V;
  const float* tb = target + (long long)b * C * g.Vfull;
  for (long long v = (long long)blockIdx.x * 256 + threadIdx.x; v < V; v += (long long)gridDim.x * 256) {
    float x[MAXCLS], gk[MAXCLS];
    const long long ti = target_index(g, v);
    float m = -INFINITY, tm = -INFINITY;
    int am = 0;
#pragma unroll
    for (int k = 0; k < MAXCLS; ++k) {
      if (k < C) {
        x[k] = to_f(lb[(long long)k * V + v]);
        const float t = tb[(long long)k * g.Vfull + ti];
        gk[k] = sA[k] * t + sB[k];
        m = fmaxf(m, x[k]);
        if (t > tm) { tm = t; am = k; }
      }
    }
    float S = 0.f;
#pragma unroll
    for (int k = 0; k < MAXCLS; ++k)
      if (k < C) { x[k] = __expf(x[k] - m); S += x[k]; }
    const float inv = 1.f / S;
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < MAXCLS; ++k)
      if (k < C) { x[k] *= inv; dot += gk[k] * x[k]; }
    const float wce = ce_scale * (cw ? cw[am] : 1.f);
#pragma unroll
    for (int k = 0; k < MAXCLS; ++k)
      if (k < C) gb[(long long)k * V + v] = from_f<T>(x[k] * (gk[k] - dot) + wce * (x[k] - (k == am ? 1.f : 0.f)));
  }
}

// ---- sliding window ----
template <typename T>
__global__ void sw_accumulate_kernel(const T* __restrict__ logits, float* __restrict__ agg, int C, int X, int Y, int Z,
                                     int x0, int y0, int z0, int px, int py, int pz, long long total) {
  const long long PV = (long long)px * py * pz, VV = (long long)X * Y * Z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int z = i % pz;
    const long long r = i / pz;
    const int y = r % py;
    const int x = r / py;
    float e[MAXCLS];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < MAXCLS; ++k)
      if (k < C) { e[k] = to_f(logits[(long long)k * PV + i]); m = fmaxf(m, e[k]); }
    float S = 0.f;
#pragma unroll
    for (int k = 0; k < MAXCLS; ++k)
      if (k < C) { e[k] = __expf(e[k] - m); S += e[k]; }
    const float inv = 1.f / S;
    const long long o = ((long long)(x0 + x) * Y + (y0 + y)) * Z + (z0 + z);
#pragma unroll
    for (int k = 0; k < MAXCLS; ++k)
      if (k < C) agg[(long long)k * VV + o] += e[k] * inv;
  }
}

struct SwSteps {
  int n[3];
  int s[3][24];
  int patch[3];
};

__global__ void sw_finalize_kernel(float* __restrict__ agg, long long* __restrict__ mask, int C, int X, int Y, int Z,
                                   SwSteps st, int normalise) {
  const long long VV = (long long)X * Y * Z;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < VV; i += (long long)gridDim.x * blockDim.x) {
    const int z = i % Z;
    const long long r = i / Z;
    const int y = r % Y;
    const int x = r / Y;
    const int pos[3] = {x, y, z};
    int cnt = 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      int c = 0;
      for (int k = 0; k < st.n[d]; ++k) c += (pos[d] >= st.s[d][k] && pos[d] < st.s[d][k] + st.patch[d]) ? 1 : 0;
      cnt *= c;
    }
    const float inv = 1.f / (float)cnt;
    float best = -INFINITY;
    int am = 0;
    for (int k = 0; k < C; ++k) {
      const float p = agg[(long long)k * VV + i] * inv;
      if (normalise) agg[(long long)k * VV + i] = p;
      if (p > best) { best = p; am = k; }
    }
    mask[i] = am;
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// Metric tail (trainer.py:382-398, 919-945; metrics.py:82-151): per-sample confusion matrices of the arg-max masks, one
// pass over the logits and the one-hot target, no host synchronisation.  conf[b][t][p] += #{voxels: argmax(target) = t,
// argmax(logits) = p} (first maximal index on ties, like torch.argmax).  compute_dice and RunningDice are derived from it.
namespace {
template <typename T>
__global__ void __launch_bounds__(256) confusion_kernel(const T* __restrict__ logits, const float* __restrict__ target, int C,
                                                        long long V, unsigned long long* __restrict__ conf) {
  __shared__ unsigned int s_cnt[MAXCLS * MAXCLS];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  const T* lg = logits + (long long)b * C * V;
  const float* tg = target + (long long)b * C * V;
  const int lane = threadIdx.x & 31;
  for (long long v0 = (long long)blockIdx.x * blockDim.x; v0 < V; v0 += (long long)gridDim.x * blockDim.x) {
    const long long v = v0 + threadIdx.x;
    int idx = -1;
    if (v < V) {
      float bl = to_f(lg[v]), bt = tg[v];
      int pl = 0, pt = 0;
      for (int c = 1; c < C; ++c) {
        const float l = to_f(lg[c * V + v]), t = tg[c * V + v];
        if (l > bl) { bl = l; pl = c; }
        if (t > bt) { bt = t; pt = c; }
      }
      idx = pt * C + pl;
    }
    // the background pair (0, 0) is by far the most frequent: count it with one ballot per warp instead of 32 atomics
    const unsigned m0 = __ballot_sync(0xffffffffu, idx == 0);
    if (lane == 0 && m0) atomicAdd(&s_cnt[0], (unsigned)__popc(m0));
    if (idx > 0) atomicAdd(&s_cnt[idx], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(&conf[(long long)b * C * C + i], (unsigned long long)s_cnt[i]);
}
}  // namespace

extern "C" {

size_t hdf_loss_sums_bytes(int B, int C) { return (size_t)(B * C * 3 + 2) * sizeof(double); }

// sums must be zeroed by this call (done here with a memset node on the stream)
int hdf_loss_level_fwd_ex(int dtype, const void* logits, const float* target, const float* class_weight, int B, int C, int Dl,
                          int Hl, int Wl, int level_stride, int depth_stride, int ignore_index, int has_ignore, float smooth,
                          float level_weight, float ce_weight, float dice_weight, double* sums, float* out_level, float* total,
                          void* stream);
int hdf_loss_level_fwd(int dtype, const void* logits, const float* target, const float* class_weight, int B, int C, int Dl,
                       int Hl, int Wl, int level_stride, int ignore_index, int has_ignore, float smooth, float level_weight,
                       float ce_weight, float dice_weight, double* sums, float* out_level, float* total, void* stream) {
  return hdf_loss_level_fwd_ex(dtype, logits, target, class_weight, B, C, Dl, Hl, Wl, level_stride, level_stride, ignore_index,
                               has_ignore, smooth, level_weight, ce_weight, dice_weight, sums, out_level, total, stream);
}
// depth_stride: stride of the full-resolution target along depth (= level_stride for volumes, 1 for flat [B, C, 1, H, W] inputs)
int hdf_loss_level_fwd_ex(int dtype, const void* logits, const float* target, const float* class_weight, int B, int C, int Dl,
                          int Hl, int Wl, int level_stride, int depth_stride, int ignore_index, int has_ignore, float smooth,
                          float level_weight, float ce_weight, float dice_weight, double* sums, float* out_level, float* total,
                          void* stream) {
  HDF_REQUIRE(logits && target && sums && out_level && total && C >= 1 && C <= MAXCLS,
              "hdf_loss_level_fwd: bad args (n_cls must be 1..%d)", MAXCLS);
  cudaStream_t s = (cudaStream_t)stream;
  const long long V = (long long)Dl * Hl * Wl;
  LevelGeom g{Dl, Hl, Wl, level_stride, depth_stride, V * depth_stride * level_stride * level_stride, Hl * level_stride, Wl * level_stride};
  cudaError_t e = cudaMemsetAsync(sums, 0, hdf_loss_sums_bytes(B, C), s);
  if (e != cudaSuccess) { hdf_set_error("hdf_loss_level_fwd: memset failed: %s", cudaGetErrorString(e)); return HDF_ERR_CUDA; }
  int chunks = (int)((V + 4095) / 4096);
  const int cap = (4 * 148 + B - 1) / B;
  if (chunks > cap) chunks = cap;
  const int rpc = cdiv(V, chunks);
  chunks = cdiv(V, rpc);
  dim3 grid(chunks, B);
  HDF_DISPATCH_DTYPE(dtype, T, {
    loss_reduce_kernel<T><<<grid, 256, 0, s>>>((const T*)logits, target, class_weight, g, C, V, rpc, sums, B);
  });
  HDF_LAUNCH_CHECK("hdf_loss_level_fwd/reduce");
  loss_finalize_kernel<<<1, 32, 0, s>>>(sums, class_weight, B, C, ignore_index, has_ignore, smooth, level_weight, ce_weight, dice_weight, out_level, total);
  HDF_LAUNCH_CHECK("hdf_loss_level_fwd/finalize");
  return HDF_OK;
}

int hdf_loss_level_bwd_ex(int dtype, const void* logits, const float* target, const float* class_weight, int B, int C, int Dl,
                          int Hl, int Wl, int level_stride, int depth_stride, int ignore_index, int has_ignore, float smooth,
                          float level_weight, float ce_weight, float dice_weight, const double* sums, const float* grad_out,
                          void* dlogits, void* stream);
int hdf_loss_level_bwd(int dtype, const void* logits, const float* target, const float* class_weight, int B, int C, int Dl,
                       int Hl, int Wl, int level_stride, int ignore_index, int has_ignore, float smooth, float level_weight,
                       float ce_weight, float dice_weight, const double* sums, const float* grad_out, void* dlogits,
                       void* stream) {
  return hdf_loss_level_bwd_ex(dtype, logits, target, class_weight, B, C, Dl, Hl, Wl, level_stride, level_stride, ignore_index,
                               has_ignore, smooth, level_weight, ce_weight, dice_weight, sums, grad_out, dlogits, stream);
}
int hdf_loss_level_bwd_ex(int dtype, const void* logits, const float* target, const float* class_weight, int B, int C, int Dl,
                          int Hl, int Wl, int level_stride, int depth_stride, int ignore_index, int has_ignore, float smooth,
                          float level_weight, float ce_weight, float dice_weight, const double* sums, const float* grad_out,
                          void* dlogits, void* stream) {
  HDF_REQUIRE(logits && target && sums && dlogits && C >= 1 && C <= MAXCLS, "hdf_loss_level_bwd: bad args");
  const long long V = (long long)Dl * Hl * Wl;
  LevelGeom g{Dl, Hl, Wl, level_stride, depth_stride, V * depth_stride * level_stride * level_stride, Hl * level_stride, Wl * level_stride};
  int gx = (int)((V + 255) / 256);
  const int cap = (8 * 148 + B - 1) / B;
  if (gx > cap) gx = cap;
  dim3 grid(gx, B);
  HDF_DISPATCH_DTYPE(dtype, T, {
    loss_grad_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)logits, target, class_weight, g, C, V, sums, B,
                                                              ignore_index, has_ignore, smooth, level_weight, ce_weight,
                                                              dice_weight, grad_out, (T*)dlogits);
  });
  HDF_LAUNCH_CHECK("hdf_loss_level_bwd");
  return HDF_OK;
}

int hdf_sw_accumulate(int dtype, const void* logits, float* agg, int C, int X, int Y, int Z, int x0, int y0, int z0, int px,
                      int py, int pz, void* stream) {
  HDF_REQUIRE(logits && agg && C >= 1 && C <= MAXCLS && x0 >= 0 && y0 >= 0 && z0 >= 0 && x0 + px <= X && y0 + py <= Y &&
                  z0 + pz <= Z, "hdf_sw_accumulate: window out of range");
  const long long total = (long long)px * py * pz;
  HDF_DISPATCH_DTYPE(dtype, T, {
    sw_accumulate_kernel<T><<<min(148 * 16, cdiv(total, 256)), 256, 0, (cudaStream_t)stream>>>((const T*)logits, agg, C, X, Y,
                                                                                              Z, x0, y0, z0, px, py, pz, total);
  });
  HDF_LAUNCH_CHECK("hdf_sw_accumulate");
  return HDF_OK;
}

int hdf_sw_finalize(float* agg, long long* mask, int C, int X, int Y, int Z, const int* steps_x, int nx, const int* steps_y,
                    int ny, const int* steps_z, int nz, int patch_x, int patch_y, int patch_z, int normalise, void* stream) {
  HDF_REQUIRE(agg && mask && nx >= 1 && ny >= 1 && nz >= 1 && nx <= 24 && ny <= 24 && nz <= 24,
              "hdf_sw_finalize: bad args (at most 24 window starts per axis)");
  SwSteps st;
  st.n[0] = nx; st.n[1] = ny; st.n[2] = nz;
  for (int i = 0; i < nx; ++i) st.s[0][i] = steps_x[i];
  for (int i = 0; i < ny; ++i) st.s[1][i] = steps_y[i];
  for (int i = 0; i < nz; ++i) st.s[2][i] = steps_z[i];
  st.patch[0] = patch_x; st.patch[1] = patch_y; st.patch[2] = patch_z;
  const long long VV = (long long)X * Y * Z;
  sw_finalize_kernel<<<min(148 * 16, cdiv(VV, 256)), 256, 0, (cudaStream_t)stream>>>(agg, mask, C, X, Y, Z, st, normalise);
  HDF_LAUNCH_CHECK("hdf_sw_finalize");
  return HDF_OK;
}

// conf [B][C][C] uint64 (+)= per-sample confusion counts (rows = target class, columns = predicted class); conf is NOT
// zeroed here.  logits [B, C, *] (fp32 / bf16, NCDHW), target one-hot fp32 of the same shape, V = voxels per sample.
int hdf_confusion_update(int dtype, const void* logits, const float* target, int B, int C, long long V, unsigned long long* conf,
                         void* stream) {
  HDF_REQUIRE(logits && target && conf && B >= 1 && C >= 1 && C <= MAXCLS && V >= 1, "hdf_confusion_update: bad args");
  int gx = (int)((V + 255) / 256);
  const int cap = (8 * 148 + B - 1) / B;
  if (gx > cap) gx = cap;
  dim3 grid(gx, B);
  HDF_DISPATCH_DTYPE(dtype, T, {
    confusion_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)logits, target, C, V, conf);
  });
  HDF_LAUNCH_CHECK("hdf_confusion_update");
  return HDF_OK;
}

}  // extern "C"
