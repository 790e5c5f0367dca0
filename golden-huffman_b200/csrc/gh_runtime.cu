// Process-wide bookkeeping of the library: error slot, launch counter, device properties.
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "gh_common.cuh"

namespace gh {

static thread_local int tl_cuda_error = 0;
static std::atomic<uint64_t> g_launches{0};

int cuda_fail(cudaError_t e) {
  tl_cuda_error = int(e);
  return GH_ERR_CUDA;
}

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {  // of the CURRENT device (a process may drive several)
  static std::atomic<int> cached[64];
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < 64 && (n = cached[dev].load(std::memory_order_relaxed)) > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < 64) cached[dev].store(n, std::memory_order_relaxed);
  return n;
}

// ---- per-kernel timing ------------------------------------------------------------------------------
// Off by default. When on, every GH_LAUNCH is bracketed by two events on its own stream; gh_profile_fetch
// resolves them. Used by bench.py's roofline pass (never inside the timed steps) and by nothing else.
#ifndef GH_EMUL
namespace {
struct ProfSlot {
  const char* name;
  uint64_t count;
  double ms;
};
struct ProfPending {
  int slot;
  cudaEvent_t a, b;
};
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfSlot> g_prof_slots;
std::vector<ProfPending> g_prof_pending;
std::vector<cudaEvent_t> g_prof_free;

cudaEvent_t prof_event() {
  if (!g_prof_free.empty()) {
    cudaEvent_t e = g_prof_free.back();
    g_prof_free.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void profile_begin(const char* name, void* stream) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  int slot = -1;
  for (size_t i = 0; i < g_prof_slots.size(); ++i)
    if (g_prof_slots[i].name == name || strcmp(g_prof_slots[i].name, name) == 0) slot = int(i);
  if (slot < 0) {
    g_prof_slots.push_back(ProfSlot{name, 0, 0.0});
    slot = int(g_prof_slots.size()) - 1;
  }
  ProfPending p{slot, prof_event(), prof_event()};
  cudaEventRecord(p.a, (cudaStream_t)stream);
  g_prof_pending.push_back(p);
}

void profile_end(void* stream) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (!g_prof_pending.empty()) cudaEventRecord(g_prof_pending.back().b, (cudaStream_t)stream);
}

static size_t profile_fetch(char* buf, size_t cap) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (auto& p : g_prof_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      g_prof_slots[p.slot].count += 1;
      g_prof_slots[p.slot].ms += ms;
    }
    g_prof_free.push_back(p.a);
    g_prof_free.push_back(p.b);
  }
  g_prof_pending.clear();
  size_t used = 0;
  for (auto& s : g_prof_slots) {
    if (!s.count) continue;
    int n = snprintf(buf + used, used < cap ? cap - used : 0, "%s %llu %.6f\n", s.name, (unsigned long long)s.count, s.ms);
    if (n < 0 || used + size_t(n) >= cap) break;
    used += size_t(n);
  }
  for (auto& s : g_prof_slots) s.count = 0, s.ms = 0.0;
  return used;
}
static void profile_enable(int on) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof_on = on != 0;
}
#else
void profile_begin(const char*, void*) {}
void profile_end(void*) {}
static size_t profile_fetch(char*, size_t) { return 0; }
static void profile_enable(int) {}
#endif

int check_launch() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GH_OK : cuda_fail(e);
}

}  // namespace gh

extern "C" {
int gh_last_cuda_error(void) { return gh::tl_cuda_error; }
uint64_t gh_launch_count(void) { return gh::g_launches.load(std::memory_order_relaxed); }
void gh_profile_enable(int on) { gh::profile_enable(on); }
size_t gh_profile_fetch(char* buf, size_t cap) { return gh::profile_fetch(buf, cap); }
}
