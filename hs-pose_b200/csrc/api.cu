// Library-level entry points of libhspose_b200.so: version, error text, device gate.
#include "common.cuh"

extern "C" int hsp_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* hsp_strerror(int code) {
  switch (code) {
    case HSP_OK: return "ok";
    case HSP_EINVAL: return "invalid argument";
    case HSP_ELAUNCH: return "CUDA kernel launch failed";
    case HSP_EDEVICE: return "device is not an sm_100 (B200) GPU";
    case HSP_EWORKSPACE: return "workspace missing or too small";
    default: return "unknown error";
  }
}

extern "C" int hsp_device_check(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return HSP_EDEVICE;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return HSP_EDEVICE;
  return major == 10 ? HSP_OK : HSP_EDEVICE;
}
