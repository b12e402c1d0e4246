// Library-level entry points: ABI version, error text, launch counter.
#include <atomic>
#include <map>
#include <mutex>
#include <stdlib.h>
#include <utility>
#include <vector>

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

// The opt-in to more than 48 KB of dynamic shared memory is a per-function AND per-device attribute: the cache is
// keyed by both (one process may drive several GPUs, and distinct instantiations never share an entry).
int ensure_dynamic_smem_impl(const void* kern, size_t smem) {
  if (smem <= 48 * 1024) return 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> granted;
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = granted[std::make_pair(kern, dev)];
  if (smem <= have) return 0;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  have = smem;
  return 0;
}

// Pinned 8-byte landing slots + events for the one entry point that hands counts back to the host
// (bevpool_prepare_v2_counts). Slots are recycled; events are per device.
static std::mutex g_slot_mu;
static std::vector<CountSlot*> g_free_slots;

CountSlot* acquire_count_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  {
    std::lock_guard<std::mutex> lock(g_slot_mu);
    for (size_t i = 0; i < g_free_slots.size(); ++i)
      if (g_free_slots[i]->dev == dev) {
        CountSlot* s = g_free_slots[i];
        g_free_slots.erase(g_free_slots.begin() + i);
        return s;
      }
  }
  CountSlot* s = new CountSlot;
  s->dev = dev;
  s->host = nullptr;
  if (cudaHostAlloc((void**)&s->host, 2 * sizeof(int32_t), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev, cudaEventDisableTiming) != cudaSuccess) {
    if (s->host) cudaFreeHost(s->host);
    delete s;
    return nullptr;
  }
  return s;
}

void release_count_slot(CountSlot* s) {
  std::lock_guard<std::mutex> lock(g_slot_mu);
  g_free_slots.push_back(s);
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
