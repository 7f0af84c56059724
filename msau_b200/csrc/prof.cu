#include "prof.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/msau_b200.h"
#include "common.cuh"

namespace msau {

struct ProfRec {
  std::string kernel;
  double flops, bytes;
  cudaEvent_t e0, e1;
};

static bool g_on = false;
static std::mutex g_mu;
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;

bool prof_enabled() { return g_on; }
bool prof_detail() {
  static int d = -1;
  if (d < 0) { const char* e = getenv("MSAU_PROF_DETAIL"); d = (e && atoi(e)) ? 1 : 0; }
  return d == 1;
}

static cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

int prof_begin(const char* kernel, double flops, double bytes, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_mu);
  ProfRec r;
  r.kernel = kernel; r.flops = flops; r.bytes = bytes;
  r.e0 = get_event(); r.e1 = get_event();
  cudaEventRecord(r.e0, st);
  g_recs.push_back(r);
  return (int)g_recs.size() - 1;
}

void prof_end(int slot, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_mu);
  cudaEventRecord(g_recs[slot].e1, st);
}

}  // namespace msau

using namespace msau;

extern "C" int msau_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_on = on != 0;
  return MSAU_OK;
}

// JSON: {"kernel": {"launches": n, "ms": total, "flops": total, "bytes": total}, ...}; clears the records.
extern "C" int msau_profile_report(char* buf, size_t cap) {
  MSAU_CHECK_ARG(buf && cap > 2, "profile_report: bad buffer");
  MSAU_CUDA_TRY(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_mu);
  struct Agg { long n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : g_recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    Agg& a = agg[r.kernel];
    a.n++; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
    g_pool.push_back(r.e0); g_pool.push_back(r.e1);
  }
  g_recs.clear();
  std::string s = "{";
  bool first = true;
  for (auto& kv : agg) {
    char tmp[512];
    snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"launches\": %ld, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}", first ? "" : ", ",
             kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops, kv.second.bytes);
    s += tmp;
    first = false;
  }
  s += "}";
  if (s.size() + 1 > cap) { set_error("profile_report: buffer too small (%zu needed)", s.size() + 1); return MSAU_ERR_ARG; }
  memcpy(buf, s.c_str(), s.size() + 1);
  return MSAU_OK;
}
