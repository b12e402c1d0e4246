// Library-level entry points: ABI version, error text, launch counter.
#include <atomic>
#include <stdlib.h>

#include "common.cuh"

namespace bevpool {
static std::atomic<int64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BEVPOOL_PDL");
    v = !(e && e[0] == '0');
  }
  return v != 0;
}
}  // namespace bevpool

extern "C" int bevpool_b200_abi_version(void) { return BEVPOOL_B200_ABI_VERSION; }

extern "C" int64_t bevpool_b200_launch_count(void) { return bevpool::g_launches.load(std::memory_order_relaxed); }

extern "C" const char* bevpool_b200_strerror(int code) {
  switch (code) {
    case BEVPOOL_OK: return "ok";
    case BEVPOOL_ERR_BAD_ARG: return "bad argument (null pointer, negative size, unsupported dtype or layout)";
    case BEVPOOL_ERR_BAD_CHANNELS: return "unsupported channel count";
    case BEVPOOL_ERR_WORKSPACE: return "workspace too small";
    case BEVPOOL_ERR_OVERFLOW: return "problem too large for the int32 ranks the API mandates";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown bevpool_b200 status";
}
