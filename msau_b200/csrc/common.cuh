// Shared helpers for the msau_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define MSAU_OK 0
#define MSAU_ERR_ARG (-1)
#define MSAU_ERR_CUDA (-2)
#define MSAU_ERR_UNSUPPORTED (-3)
#define MSAU_ERR_WORKSPACE (-4)

namespace msau {

void set_error(const char* fmt, ...);
const char* get_error();

#define MSAU_CUDA_TRY(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::msau::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return MSAU_ERR_CUDA;                                                                   \
    }                                                                                         \
  } while (0)

#define MSAU_TRY(expr)            \
  do {                            \
    int _r = (expr);              \
    if (_r != MSAU_OK) return _r; \
  } while (0)

#define MSAU_CHECK_ARG(cond, ...)      \
  do {                                 \
    if (!(cond)) {                     \
      ::msau::set_error(__VA_ARGS__);  \
      return MSAU_ERR_ARG;             \
    }                                  \
  } while (0)

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }
static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

int sm_count();
void count_launch(int n);   // bench.py's gpu_launches counter

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// Kernels launched through launch_pdl() may start while their predecessor in the stream is still draining: their prologue (TMEM
// allocation, barrier initialisation, weight images -> shared memory, index arithmetic) overlaps the predecessor's tail and the
// launch latency disappears.  Protocol, the same in every kernel that is launched this way:
//     prologue (touches nothing but kernel parameters, shared memory, TMEM and packed weights)
//     pdl_wait();      // predecessor complete and its memory visible -- BEFORE the first access to any activation / gradient
//     pdl_trigger();   // successors may now be scheduled (after the wait, so a successor's prologue only ever overlaps THIS
//                      // kernel's body: everything older is complete, in particular the weight-packing kernels)
// The host side decides per launch (thread-local state set by the plan): off for the first launches after the packing kernels and
// when the engine option "pdl" is 0.  Without the launch attribute both device instructions are no-ops.
void pdl_set(bool enabled, int hold);   // hold: number of upcoming launches that must stay fully serialised
bool pdl_take();                        // consumes one launch's decision

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_take() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ------------------------------------------------------------------ kernel argument blocks
// A 4-D activation in HBM: NHWC fp32, `pitch` floats per pixel (>= channels used, multiple of 4).
// nchw=1 (network input only): planar [B, c_logical, H, W].

struct ConvArgs {
  const float* src1; int c1; int p1; int src1_nchw; int c1_logical;
  const float* mask1; int pm1;     // optional: src1 is zeroed where mask1 <= 0 (same pixel, same channel)
  int relu1;                       // relu applied to src1 on load
  const float* src2; int c2; int p2;  // optional second source (channel concat [src1 | src2])
  const float* w;                  // packed [taps][c1+c2][coutp]
  const float* bias;               // [coutp] or nullptr
  float* out; int po; int coutp;
  int B, Hin, Win;                 // source extent
  int Hq, Wq;                      // iteration extent (virtual output grid)
  int kh, kw, dil, stride, pad_t, pad_l;   // in_y = q_y*stride - pad_t + ky*dil
  int Hout, Wout, osy, oy0, ox0;   // physical output: (q_y*osy + oy0, q_x*osy + ox0) inside [Hout, Wout]
  int relu;                        // relu after bias
  const float* res; int pr;        // + res (after relu)
  int relu2;                       // relu after the residual add
  const float* omask; int pom;     // * (omask > 0)
  const float* add; int pa; const float* addmask; int pam;   // + add * (addmask > 0 if addmask)
  int accumulate;                  // out += value
  // transposed-conv views (tensor-core path only; model/layers/layers.py:207-260):
  //  s2d: src1 is the space-to-depth view of a physical [B, Hs, Ws, cph] tensor: virtual channel (py, px, c) of
  //       virtual pixel (y, x) = physical pixel (2y+py, 2x+px), channel c;  c1 = 4*cph, (Hin, Win) = virtual extent
  //  d2s: the output is the depth-to-space view of [B, Hout, Wout, cph]: virtual output channel (py, px, c) of
  //       virtual pixel (y, x) is stored at physical pixel (2y+py, 2x+px), channel c;  coutp = 4*cph
  int s2d, d2s, cph, Hs, Ws;
  int d2s_col0;                    // d2s: this launch produces the virtual output columns [d2s_col0, d2s_col0 + coutp) of the 4 * cph
  // tensor-core kernels return at once when *skip_flag == 0 (the structured first-layer kernels did the work)
  const int* skip_flag;
};

struct WgradArgs {
  // dW[tap][ca][cb] += sum_q A[b, sa*qy + ty*dila - pada_t, ..][ca] * Bm[b, sb*qy + ty*dilb - padb_t, ..][cb]
  const float* A; int ca; int pa; int a_nchw; int ca_logical; int reluA;
  int Ha, Wa, sa, dila, pada_t, pada_l;
  const float* Bm; int cb; int pb; const float* maskB; int pmb;
  int Hb, Wb, sb, dilb, padb_t, padb_l;
  int B, Hq, Wq, kh, kw;
  float* dW; long s_ca, s_cb; int ca_lim, cb_lim;   // dW index = ca*s_ca + cb*s_cb + (ty*kw+tx)
  float* dbias;                                     // += sum_q Bm (tap-independent B only) or nullptr
  // b_s2d (tensor-core path): Bm is the space-to-depth view of a physical [B, Hb, Wb, cph] tensor (cb = 4*cph) and the
  // result is the weight gradient of a ConvTranspose2d(k3, s2, p1): virtual tap (ty, tx) x phase (py, px) -> (ky, kx)
  int b_s2d, cph;
  int b_col0;                                       // b_s2d: first virtual column of this launch's window (cb columns) inside the 4 * cph
  const int* skip_flag;                             // as in ConvArgs
  int no_tc4;                                       // engine option wgrad_multi_plane = 0: keep this launch off wgrad_tc4 (tests of the other kernels)
};

int launch_conv(const ConvArgs& a, cudaStream_t st);
int launch_wgrad(const WgradArgs& a, cudaStream_t st);

}  // namespace msau
