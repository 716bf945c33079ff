#include "profiler.h"

#include <mutex>
#include <vector>

namespace ls {
namespace {
struct Rec {
  cudaEvent_t a, b;
  int kind;
  double flops, bytes;
};
std::mutex g_mu;
bool g_on = false;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;

cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

bool prof_enabled() { return g_on; }

void prof_begin() {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& r : g_recs) g_pool.push_back(r.a), g_pool.push_back(r.b);
  g_recs.clear();
  g_on = true;
}

void prof_end(long long* launches, double* ms, double* flops, double* bytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_on = false;
  cudaDeviceSynchronize();
  for (int k = 0; k < PK_COUNT; ++k) launches[k] = 0, ms[k] = 0, flops[k] = 0, bytes[k] = 0;
  for (auto& r : g_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) ms[r.kind] += t;
    launches[r.kind] += 1;
    flops[r.kind] += r.flops;
    bytes[r.kind] += r.bytes;
    g_pool.push_back(r.a), g_pool.push_back(r.b);
  }
  g_recs.clear();
}

ProfScope::ProfScope(cudaStream_t s, int kind, double flops, double bytes) : s_(s), slot_(-1) {
  if (!g_on) return;
  std::lock_guard<std::mutex> lk(g_mu);
  Rec r{get_event(), get_event(), kind, flops, bytes};
  cudaEventRecord(r.a, s);
  slot_ = (int)g_recs.size();
  g_recs.push_back(r);
}

ProfScope::~ProfScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  cudaEventRecord(g_recs[slot_].b, s_);
}

}  // namespace ls
