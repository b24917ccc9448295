// ABI version + error strings for libveonlift.
#include <atomic>

#include "common.cuh"

static std::atomic<unsigned long long> g_launches{0};
extern "C" void veon_count_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" uint64_t veon_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

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

// Leave `n` SMs free of the path's persistent grids (0 = use them all); returns the previous
// value.  Process-wide: meant to be set once by a multi-GPU caller that overlaps collectives.
extern "C" int veon_reserve_sms(int n) {
  int& r = veon::reserved_sms();
  const int old = r;
  r = n < 0 ? 0 : n;
  return old;
}
