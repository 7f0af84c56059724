// Optional per-kernel timing (CUDA events on the launching stream) with algorithmic work counters.
// Disabled by default; bench.py enables it for a separate pass to derive the per-kernel roofline numbers.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

namespace msau {

bool prof_enabled();
bool prof_detail();
// returns a slot index (or -1 when disabled); records the start event on `st`
int prof_begin(const char* kernel, double flops, double bytes, cudaStream_t st);
void prof_end(int slot, cudaStream_t st);

struct ProfScope {
  int slot;
  cudaStream_t st;
  ProfScope(const char* kernel, double flops, double bytes, cudaStream_t s) : slot(-1), st(s) {
    if (prof_enabled()) slot = prof_begin(kernel, flops, bytes, s);
  }
  // detailed variant (MSAU_PROF_DETAIL=1): the key also carries the layer shape
  ProfScope(const char* kernel, int cin, int cout, int k, int dil, int W, int flags, double flops, double bytes, cudaStream_t s) : slot(-1), st(s) {
    if (!prof_enabled()) return;
    if (prof_detail()) {
      char buf[96];
      snprintf(buf, sizeof(buf), "%s[c%d->%d k%d d%d w%d f%d]", kernel, cin, cout, k, dil, W, flags);
      slot = prof_begin(buf, flops, bytes, s);
    } else {
      slot = prof_begin(kernel, flops, bytes, s);
    }
  }
  ~ProfScope() {
    if (slot >= 0) prof_end(slot, st);
  }
};

}  // namespace msau
