// Library-level entry points of the C ABI (version, status strings, device check).
#include "common.cuh"

extern "C" int cn_version(void) { return 100; }   // 0.1.0

extern "C" const char* cn_status_string(int status) {
  switch (status) {
    case CN_OK: return "ok";
    case CN_ERR_BAD_ARG: return "bad argument (null pointer, negative size or inconsistent shapes)";
    case CN_ERR_UNSUPPORTED: return "unsupported shape for the sm_100a kernels";
    case CN_ERR_LAUNCH: return "CUDA launch failed";
    case CN_ERR_WORKSPACE: return "workspace too small";
    case CN_ERR_DEVICE: return "no sm_100 (Blackwell) device";
    default: return "unknown status";
  }
}

extern "C" int cn_device_check(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return CN_ERR_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return CN_ERR_DEVICE;
  return prop.major == 10 ? CN_OK : CN_ERR_DEVICE;
}
