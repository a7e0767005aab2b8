// Timing / introspection extensions (include/cuvs_b200/ext.h).
#include "common.hpp"
#include "timing.hpp"

#include <cuvs_b200/ext.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace b200 {
namespace {
std::atomic<int> g_timing{0};
std::atomic<long long> g_launches{0};
thread_local int g_last_flagged = 0;
std::mutex g_mu;
std::map<std::string, std::vector<std::pair<cudaEvent_t, cudaEvent_t>>> g_events;
}  // namespace

bool timing_enabled() { return g_timing.load(std::memory_order_relaxed) != 0; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void set_last_flagged(int n) { g_last_flagged = n; }

void timing_begin(const char*, cudaStream_t stream, cudaEvent_t* ev_start)
{
  cudaEventCreate(ev_start);
  cudaEventRecord(*ev_start, stream);
}
void timing_end(const char* name, cudaStream_t stream, cudaEvent_t ev_start)
{
  cudaEvent_t stop;
  cudaEventCreate(&stop);
  cudaEventRecord(stop, stream);
  std::lock_guard<std::mutex> lk(g_mu);
  g_events[name].emplace_back(ev_start, stop);
}
}  // namespace b200

using namespace b200;

extern "C" {
void cuvsB200TimingEnable(int on) { g_timing.store(on); }
void cuvsB200TimingReset(void)
{
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& kv : g_events)
    for (auto& p : kv.second) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  g_events.clear();
}
double cuvsB200TimingTotalMs(const char* name, int* count)
{
  std::lock_guard<std::mutex> lk(g_mu);
  double total = 0;
  int n        = 0;
  auto it      = g_events.find(name ? name : "");
  if (it != g_events.end()) {
    for (auto& p : it->second) {
      float ms = 0;
      if (cudaEventSynchronize(p.second) == cudaSuccess && cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) {
        total += ms;
        ++n;
      }
    }
  }
  if (count) *count = n;
  return total;
}
int cuvsB200LastFlagged(void) { return g_last_flagged; }
long long cuvsB200KernelLaunches(void) { return g_launches.load(); }
}
