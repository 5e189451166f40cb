// Library-level entry points: init / error string / version.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";
static int g_sm_count = 0;
static int g_device = -1;
unsigned long long g_hdf_launches = 0;

void hdf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int hdf_sm_count_cached() { return g_sm_count > 0 ? g_sm_count : 148; }

extern "C" {

const char* hdf_last_error_string(void) { return g_err; }

int hdf_version(void) { return 100; }

int hdf_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    hdf_set_error("hdf_init: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    return HDF_ERR_CUDA;
  }
  if (device < 0 || device >= count) {
    hdf_set_error("hdf_init: device %d out of range (0..%d)", device, count - 1);
    return HDF_ERR_ARG;
  }
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) {
    hdf_set_error("hdf_init: cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    return HDF_ERR_CUDA;
  }
  if (p.major != 10) {
    hdf_set_error("hdf_init: device %d is sm_%d%d; libhdf_b200 is built for sm_100a (B200) only", device, p.major, p.minor);
    return HDF_ERR_ARCH;
  }
  g_sm_count = p.multiProcessorCount;
  g_device = device;
  // Experiment knob: prefer the shared-memory-heavy L1 carve-out device-wide, like the persistent convolution CTAs.
  // Measured neutral (29.9 vs 29.4 ms/step): the token kernels that stall next to a convolution kernel do so because
  // their own shared memory (dct_c_bwd: 100 KB) does not fit beside a 200 KB CTA, not because of the carve-out.
  if (getenv("HDF_PREFER_SHARED") != nullptr) {
    int cur = -1;
    cudaGetDevice(&cur);
    cudaSetDevice(device);
    cudaDeviceSetCacheConfig(cudaFuncCachePreferShared);
    if (cur >= 0 && cur != device) cudaSetDevice(cur);
  }
  return HDF_OK;
}

int hdf_sm_count(void) { return g_sm_count; }

unsigned long long hdf_launch_count(void) { return g_hdf_launches; }

}  // extern "C"
