// Fused Adam / AdamW over the flat gradient arena (reference optimizer selection and the
// no-decay grouping of 1-D parameters and biases: trainer.py:793-840).  One launch updates all ~400 parameter tensors:
// a device table maps every 4096-element chunk to (tensor, offset); gradients and both moments live in flat fp32
// buffers with the arena's offsets, parameters stay in their own torch storage (checkpoint layout untouched).
// Step count and learning rate are read from device memory so that a captured CUDA graph can be replayed with a
// scheduler changing the rate between replays.
#include "common.cuh"

namespace {

constexpr int ADAM_CHUNK = 4096;

struct AdamTensor {
  float* p;            // parameter storage (contiguous)
  long long off;       // offset of its gradient / moments in the flat buffers
  long long n;         // elements
  float wd;            // weight decay of its group
};

// hyper[0] = lr, hyper[1] = step count (already incremented for this step)
__global__ void __launch_bounds__(256) adam_kernel(const AdamTensor* __restrict__ tab, const int2* __restrict__ chunks,
                                                  const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                  const float* __restrict__ hyper, float b1, float b2, float eps, int adamw,
                                                  float grad_scale) {
  const int2 ch = chunks[blockIdx.x];
  const AdamTensor t = tab[ch.x];
  const long long i0 = (long long)ch.y * ADAM_CHUNK;
  const long long i1 = min(t.n, i0 + ADAM_CHUNK);
  const float lr = hyper[0], step = hyper[1];
  const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    float p = t.p[i];
    float gr = g[t.off + i] * grad_scale;
    if (adamw) p *= 1.f - lr * t.wd;
    else gr = fmaf(t.wd, p, gr);
    const float mm = fmaf(b1, m[t.off + i], (1.f - b1) * gr);
    const float vv = fmaf(b2, v[t.off + i], (1.f - b2) * gr * gr);
    m[t.off + i] = mm;
    v[t.off + i] = vv;
    const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
    t.p[i] = p - step_size * (mm / denom);
  }
}

__global__ void adam_advance_kernel(float* hyper) { hyper[1] += 1.f; }

}  // namespace

extern "C" {

size_t hdf_adam_table_bytes(int ntensors) { return (size_t)ntensors * sizeof(AdamTensor); }
int hdf_adam_chunk(void) { return ADAM_CHUNK; }

// host-side packing of one table entry (the caller uploads the table once; layout private to the library)
int hdf_adam_table_set(void* host_table, int index, float* param, long long offset, long long n, float weight_decay) {
  HDF_REQUIRE(host_table && param && index >= 0 && n > 0, "hdf_adam_table_set: bad args");
  AdamTensor* t = reinterpret_cast<AdamTensor*>(host_table) + index;
  t->p = param; t->off = offset; t->n = n; t->wd = weight_decay;
  return HDF_OK;
}

// One optimizer step.  table/chunks: device copies (chunks = int2 (tensor, chunk index) per CTA).  hyper: device
// float[2] = {lr, step}; the step counter is advanced here before the update (bias correction uses the new value).
int hdf_adam_step(const void* table, const void* chunks, int nchunks, const float* grad_flat, float* m_flat, float* v_flat,
                  float* hyper, float beta1, float beta2, float eps, int adamw, float grad_scale, void* stream) {
  HDF_REQUIRE(table && chunks && nchunks > 0 && grad_flat && m_flat && v_flat && hyper, "hdf_adam_step: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  adam_advance_kernel<<<1, 1, 0, s>>>(hyper);
  HDF_LAUNCH_CHECK("hdf_adam_step/advance");
  adam_kernel<<<nchunks, 256, 0, s>>>((const AdamTensor*)table, (const int2*)chunks, grad_flat, m_flat, v_flat, hyper, beta1, beta2,
                                      eps, adamw, grad_scale);
  HDF_LAUNCH_CHECK("hdf_adam_step");
  return HDF_OK;
}

}  // extern "C"
