// ABI version + error strings for libveonlift.
#include "common.cuh"

extern "C" int veon_abi_version(void) { return VEON_ABI_VERSION; }

extern "C" const char* veon_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case VEON_E_BADARG: return "veon: bad argument (null pointer or non-positive dimension)";
    case VEON_E_WORKSPACE: return "veon: workspace too small";
    case VEON_E_RANGE: return "veon: index space exceeds int32 rank arrays";
    case VEON_E_UNSUPPORTED: return "veon: unsupported configuration";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "veon: unknown error";
}
