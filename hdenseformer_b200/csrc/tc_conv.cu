// placeholder (replaced by the tcgen05 implementation)
#include "common.cuh"
extern "C" {
int hdf_tc_supported(int, int, int) { return 0; }
size_t hdf_tc_pack_bytes(int Cin, int Cout) { return (size_t)27 * Cin * Cout * 2; }
int hdf_tc_pack_weights(const float*, void*, int, int, long long, long long, int, void*) { hdf_set_error("tc path not built"); return HDF_ERR_UNSUPPORTED; }
int hdf_tc_conv3d_fwd(const void*, long long, const void*, const float*, void*, long long, int, int, int, int, int, int, double*, void*) { hdf_set_error("tc path not built"); return HDF_ERR_UNSUPPORTED; }
size_t hdf_tc_wgrad_workspace(int, int, int, int, int, int) { return 0; }
int hdf_tc_conv3d_wgrad(const void*, long long, const void*, long long, float*, long long, long long, int, int, int, int, int, int, void*, size_t, int, void*) { hdf_set_error("tc path not built"); return HDF_ERR_UNSUPPORTED; }
}
