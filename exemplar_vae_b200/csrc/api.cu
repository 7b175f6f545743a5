// Library-level entry points of the exvae_b200 C ABI (version, errors, device query).
#include "common.cuh"

namespace exvae {
int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      n <= 0) {
    (void)cudaGetLastError();
    return 148;  // B200; only reached when sizing a workspace without a visible device
  }
  cached = n;
  return n;
}
}  // namespace exvae

extern "C" int exvae_abi_version(void) { return EXVAE_ABI_VERSION; }

extern "C" const char* exvae_error_string(int code) {
  switch (code) {
    case EXVAE_OK: return "ok";
    case EXVAE_ERR_INVALID_ARG: return "invalid argument (null pointer or non-positive size)";
    case EXVAE_ERR_UNSUPPORTED: return "unsupported configuration for this kernel";
    case EXVAE_ERR_WORKSPACE: return "workspace too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown exvae error";
}

extern "C" int exvae_device_info(int* sm_count_out, int* cc_major, int* cc_minor) {
  int dev = 0;
  EXVAE_CUDA(cudaGetDevice(&dev));
  int n = 0, maj = 0, min = 0;
  EXVAE_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  EXVAE_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  EXVAE_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count_out) *sm_count_out = n;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  return EXVAE_OK;
}
