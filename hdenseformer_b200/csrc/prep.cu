// GPU input pipeline: the reference's per-sample numpy / skimage transforms
//   RandomCrop3D -> PETandCTNormalize | MRNormalize | Trunc_and_Normalize -> RandomTranslationRotationZoom3D -> RandomFlip3D
//   -> To_Tensor            (data_utils/transformer_3d.py:7-169, data_utils/data_loader.py:16-68,126-159, trainer.py:128-141)
// as two kernels over a raw volume that already sits in HBM:
//   1. prep_stats_kernel : per-channel sum / sum of squares / min / max over the crop window (the PET z-score and the MR
//      max-normalisation are statistics of the CROPPED patch: normalisation comes after the crop in the reference's list);
//   2. prep_sample_kernel: one thread per output voxel.  It maps the voxel back through the flip and the affine warp to a
//      source coordinate in crop space, gathers the (up to) 8 neighbours of every channel straight from the raw volume,
//      normalises each neighbour, interpolates in double like scipy.ndimage.map_coordinates(order=1, mode='constant',
//      cval=0) -- which is what skimage.transform.warp(image, coords) runs -- and writes the network input [M,D,H,W] fp32
//      and the one-hot label [C,D,H,W] fp32.  No intermediate volume is written; the random parameters (crop origin,
//      affine matrix, flip axis) are drawn on the host with the reference's RNG calls and passed in.
// HBM-bound gather: 4 (M + 1) bytes read and 4 (M + C) bytes written per output voxel.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int PREP_MAXC = 8;         // channels / classes handled per thread
constexpr int PREP_CHUNKS = 148;      // statistics partials per channel (one CTA each)

struct PrepGeom {
  int M, Dv, Hv, Wv;                 // raw volume [M][Dv][Hv][Wv], raw label [Dv][Hv][Wv]
  int d0, h0, w0;                    // crop origin
  int D, H, W;                       // crop (= output) size
};

// partial[m][chunk][4] = sum, sum of squares, min, max of channel m over the crop window (chunk = strided part of it)
__global__ void __launch_bounds__(256) prep_stats_kernel(const float* __restrict__ vol, PrepGeom g, double* __restrict__ partial) {
  const int m = blockIdx.y, chunk = blockIdx.x;
  const long long V = (long long)g.D * g.H * g.W;
  const float* base = vol + (long long)m * g.Dv * g.Hv * g.Wv;
  double s = 0.0, q = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  for (long long i = (long long)chunk * 256 + threadIdx.x; i < V; i += (long long)gridDim.x * 256) {
    const int w = (int)(i % g.W), h = (int)((i / g.W) % g.H), d = (int)(i / ((long long)g.W * g.H));
    const float x = base[((long long)(g.d0 + d) * g.Hv + (g.h0 + h)) * g.Wv + (g.w0 + w)];
    s += (double)x;
    q += (double)x * (double)x;
    mn = fminf(mn, x);
    mx = fmaxf(mx, x);
  }
  __shared__ double sh[4][8];
  s = warp_sum_d(s);
  q = warp_sum_d(q);
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = q; sh[2][warp] = (double)mn; sh[3][warp] = (double)mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, c = sh[2][0], d = sh[3][0];
    for (int k = 0; k < 8; ++k) { a += sh[0][k]; b += sh[1][k]; c = fmin(c, sh[2][k]); d = fmax(d, sh[3][k]); }
    double* out = partial + ((long long)m * gridDim.x + chunk) * 4;
    out[0] = a; out[1] = b; out[2] = c; out[3] = d;
  }
}

struct PrepNorm {
  int mode;                          // 0 none, 1 PET/CT, 2 MR max, 3 truncate + scale
  float p0, p1;                      // PET/CT: CT window centre, half-width; truncate: lo, hi
};

struct PrepChan { float a, b, lo, hi; int kind; };   // per-channel constants derived from the statistics

// x -> normalised value, in fp32 with the reference's operation order
__device__ __forceinline__ float prep_norm(float x, const PrepChan& k) {
  switch (k.kind) {
    case 1: return __fdiv_rn(fminf(fmaxf(x, k.lo), k.hi) - k.a, k.b);       // (clip(x, m-w, m+w) - m) / w
    case 2: return __fdiv_rn(x - k.a, k.b);                                // (x - mean) / (std + 1e-3)
    case 3: { const float v = k.b != 0.f ? __fdiv_rn(x, k.b) : x; return v < 0.f ? 0.f : v; }   // x / max, negatives -> 0
    case 4: return __fdiv_rn(fminf(fmaxf(x - k.lo, 0.f), k.hi), k.hi);      // clip(x - lo, 0, range) / range
    default: return x;
  }
}

__global__ void __launch_bounds__(256) prep_sample_kernel(const float* __restrict__ vol, const float* __restrict__ lab, PrepGeom g,
                                                         PrepNorm nm, const double* __restrict__ partial, int chunks,
                                                         const double* __restrict__ affine, int flip_axis, int num_class,
                                                         float* __restrict__ img_out, float* __restrict__ lab_out) {
  __shared__ PrepChan ch[PREP_MAXC];
  __shared__ double A[12];
  if (threadIdx.x < g.M) {
    const int m = threadIdx.x;
    PrepChan k{0.f, 1.f, 0.f, 0.f, 0};
    if (nm.mode == 1 && m == 0) { k.kind = 1; k.a = nm.p0; k.b = nm.p1; k.lo = nm.p0 - nm.p1; k.hi = nm.p0 + nm.p1; }
    if ((nm.mode == 1 && m == 1) || nm.mode == 2) {
      double s = 0.0, q = 0.0, mx = -INFINITY;
      for (int c = 0; c < chunks; ++c) {                  // fixed order: deterministic
        const double* p = partial + ((long long)m * chunks + c) * 4;
        s += p[0]; q += p[1]; mx = fmax(mx, p[3]);
      }
      const double V = (double)g.D * g.H * g.W;
      if (nm.mode == 1) {
        const double mean = s / V;
        double var = q / V - mean * mean;
        if (var < 0.0) var = 0.0;
        k.kind = 2; k.a = (float)mean; k.b = (float)sqrt(var) + 1e-3f;
      } else {
        k.kind = 3; k.b = (float)mx;
      }
    }
    if (nm.mode == 3) { k.kind = 4; k.lo = nm.p0; k.hi = nm.p1 - nm.p0; }
    ch[m] = k;
  }
  if (affine && threadIdx.x >= 32 && threadIdx.x < 44) A[threadIdx.x - 32] = affine[threadIdx.x - 32];
  __syncthreads();

  const long long V = (long long)g.D * g.H * g.W;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= V) return;
  const int w = (int)(i % g.W), h = (int)((i / g.W) % g.H), d = (int)(i / ((long long)g.W * g.H));
  // RandomFlip3D is applied after the warp: output voxel (d, h, w) shows warped voxel (d, H-1-h, w) or (d, h, W-1-w)
  const int hs = flip_axis == 1 ? g.H - 1 - h : h, ws = flip_axis == 2 ? g.W - 1 - w : w;
  const long long plane = (long long)g.Hv * g.Wv, cvol = plane * g.Dv;
  const long long crop0 = ((long long)g.d0 * g.Hv + g.h0) * g.Wv + g.w0;

  if (!affine) {
    const long long src = crop0 + ((long long)d * g.Hv + hs) * g.Wv + ws;
    for (int m = 0; m < g.M; ++m) img_out[m * V + i] = prep_norm(vol[m * cvol + src], ch[m]);
    if (lab_out) {
      const float l = lab[src];
      bool any = false;
      for (int z = 1; z < num_class; ++z) {
        const bool hit = l == (float)z;
        any |= hit;
        lab_out[z * V + i] = hit ? 1.f : 0.f;
      }
      lab_out[i] = any ? 0.f : 1.f;
    }
    return;
  }

  // source coordinate in crop space: A (idx - size/2) + t + size/2, all in double like the reference's np.dot
  const double c0 = (double)d - g.D / 2.0, c1 = (double)hs - g.H / 2.0, c2 = (double)ws - g.W / 2.0;
  const double sd = A[0] * c0 + A[1] * c1 + A[2] * c2 + A[3] + g.D / 2.0;
  const double sh = A[4] * c0 + A[5] * c1 + A[6] * c2 + A[7] + g.H / 2.0;
  const double sw = A[8] * c0 + A[9] * c1 + A[10] * c2 + A[11] + g.W / 2.0;
  // map_coordinates(mode='constant'): a coordinate outside [0, n-1] in any dimension gives cval = 0 (no interpolation
  // across the border); inside, order 1 = weights (1 - t, t) on floor and floor + 1
  const bool inside = sd >= 0.0 && sd <= g.D - 1.0 && sh >= 0.0 && sh <= g.H - 1.0 && sw >= 0.0 && sw <= g.W - 1.0;
  // weights are applied one dimension after the other to the sample, ((x * wd) * wh) * ww, the order scipy's
  // NI_GeometricTransform multiplies its spline values in
  double wd[2] = {0.0, 0.0}, wh[2] = {0.0, 0.0}, ww[2] = {0.0, 0.0};
  long long off[8];
  if (inside) {
    const double fd = floor(sd), fh = floor(sh), fw = floor(sw);
    const int id = (int)fd, ih = (int)fh, iw = (int)fw;
    wd[1] = sd - fd; wd[0] = 1.0 - wd[1];
    wh[1] = sh - fh; wh[0] = 1.0 - wh[1];
    ww[1] = sw - fw; ww[0] = 1.0 - ww[1];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int bd = k >> 2, bh = (k >> 1) & 1, bw = k & 1;
      // the upper neighbour of a coordinate that sits exactly on the last sample has weight 0: clamp its index
      const int jd = min(id + bd, g.D - 1), jh = min(ih + bh, g.H - 1), jw = min(iw + bw, g.W - 1);
      off[k] = crop0 + ((long long)jd * g.Hv + jh) * g.Wv + jw;
    }
  }
  for (int m = 0; m < g.M; ++m) {
    double acc = 0.0;
    if (inside) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(__dmul_rn((double)prep_norm(vol[m * cvol + off[k]], ch[m]), wd[k >> 2]),
                                                 wh[(k >> 1) & 1]), ww[k & 1]));     // no FMA contraction: scipy rounds each step
    }
    img_out[m * V + i] = (float)acc;
  }
  if (lab_out) {
    float l[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) l[k] = inside ? lab[off[k]] : -1.f;
    // per class: warp the indicator volume, threshold at 0.5, later classes overwrite earlier ones (transformer_3d.py:111-114)
    int cls = 0;
    for (int z = 1; z < num_class; ++z) {
      double acc = 0.0;
      if (inside) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(__dmul_rn(l[k] == (float)z ? 1.0 : 0.0, wd[k >> 2]), wh[(k >> 1) & 1]), ww[k & 1]));
      }
      if ((float)acc >= 0.5f) cls = z;
    }
    for (int z = 1; z < num_class; ++z) lab_out[z * V + i] = cls == z ? 1.f : 0.f;
    lab_out[i] = cls == 0 ? 1.f : 0.f;
  }
}

}  // namespace

extern "C" {

size_t hdf_prep_workspace(int M) { return (size_t)M * PREP_CHUNKS * 4 * sizeof(double); }

// One sample of the 3-D training / validation input pipeline.
//   vol   [M][Dv][Hv][Wv] fp32 raw image channels, lab [Dv][Hv][Wv] fp32 raw label (class ids), both on the device
//   crop  origin (d0, h0, w0), size (D, H, W)                       RandomCrop3D (transformer_3d.py:7-42)
//   norm_mode 0 none | 1 PETandCTNormalize(mean = p0, w = p1) (data_loader.py:53-68) | 2 MRNormalize (:39-50)
//             | 3 Trunc_and_Normalize(scale = (p0, p1)) (:16-36); statistics are those of the crop window
//   affine 3 x 4 row-major double on the DEVICE (rows 0-2 of compose(T, R, Z), transformer_3d.py:66-105) or null = no warp
//   flip_axis 0 none | 1 H | 2 W                                    RandomFlip3D (transformer_3d.py:122-169)
//   img_out [M][D][H][W] fp32, lab_out [num_class][D][H][W] fp32 one-hot with channel 0 = background (To_Tensor,
//   data_loader.py:126-159) or null.  workspace: hdf_prep_workspace(M) bytes.
int hdf_prep_sample(const float* vol, const float* lab, int M, int Dv, int Hv, int Wv, int d0, int h0, int w0, int D, int H, int W,
                    int norm_mode, float p0, float p1, const double* affine, int flip_axis, int num_class, float* img_out,
                    float* lab_out, void* workspace, size_t ws_bytes, void* stream) {
  HDF_REQUIRE(vol && img_out && workspace, "hdf_prep_sample: null pointer");
  HDF_REQUIRE(M >= 1 && M <= PREP_MAXC && num_class >= 1 && num_class <= PREP_MAXC, "hdf_prep_sample: M and num_class must be in 1..8");
  HDF_REQUIRE(lab_out == nullptr || lab != nullptr, "hdf_prep_sample: label output without label input");
  HDF_REQUIRE(d0 >= 0 && h0 >= 0 && w0 >= 0 && D >= 1 && H >= 1 && W >= 1 && d0 + D <= Dv && h0 + H <= Hv && w0 + W <= Wv,
              "hdf_prep_sample: crop window outside the volume");
  HDF_REQUIRE(norm_mode >= 0 && norm_mode <= 3 && flip_axis >= 0 && flip_axis <= 2, "hdf_prep_sample: bad mode");
  HDF_REQUIRE(norm_mode != 1 || M >= 2, "hdf_prep_sample: PET/CT normalisation needs 2 channels");
  HDF_REQUIRE(norm_mode != 3 || p1 > p0, "hdf_prep_sample: empty truncation range");
  HDF_REQUIRE(ws_bytes >= hdf_prep_workspace(M), "hdf_prep_sample: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  PrepGeom g{M, Dv, Hv, Wv, d0, h0, w0, D, H, W};
  const long long V = (long long)D * H * W;
  int chunks = (int)((V + 2047) / 2048);
  if (chunks > PREP_CHUNKS) chunks = PREP_CHUNKS;
  if (norm_mode == 1 || norm_mode == 2) {
    prep_stats_kernel<<<dim3(chunks, M), 256, 0, s>>>(vol, g, (double*)workspace);
    HDF_LAUNCH_CHECK("hdf_prep_sample/stats");
  }
  PrepNorm nm{norm_mode, p0, p1};
  prep_sample_kernel<<<(unsigned)((V + 255) / 256), 256, 0, s>>>(vol, lab, g, nm, (const double*)workspace, chunks, affine, flip_axis,
                                                               num_class, img_out, lab_out);
  HDF_LAUNCH_CHECK("hdf_prep_sample");
  return HDF_OK;
}

}  // extern "C"
