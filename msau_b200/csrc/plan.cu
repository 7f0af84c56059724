// MsauPlan: the launch plan of the MSAU network for a fixed (B, H, W).
//
// Walks the reference architecture (model/model.py:53-396) once on the host, assigning every activation
// an offset in a caller-provided workspace and every parameter tensor its offset in the flat
// state_dict-ordered parameter buffer, then replays the walk as kernel launches:
//   msau_forward        MSAUNet.forward model.py:378-396 (down tower :129-164, up tower :224-259,
//                       residual block :37-50, attention attention.py:152-162, 4x4 heads :375-376,390)
//   msau_loss_backward  MSAUWrapper.loss model.py:446-459 + the hand-derived gradient of the above
// Concats are never materialised (two-source convs), F.pad never runs (bounds-checked tiles), ReLU /
// residual add / ReLU-backward masks live in conv epilogues and operand loaders.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/msau_b200.h"
#include "attention.cuh"
#include "common.cuh"
#include "conv1x1.cuh"
#include "conv_tc.cuh"
#include "first_layer.cuh"
#include "optim.cuh"
#include "pointwise.cuh"

namespace msau {

// ------------------------------------------------------------------ error / runtime helpers
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

static thread_local bool t_pdl_on = false;
static thread_local int t_pdl_hold = 0;
void pdl_set(bool enabled, int hold) { t_pdl_on = enabled; t_pdl_hold = hold; }
bool pdl_take() {
  if (t_pdl_hold > 0) { --t_pdl_hold; return false; }
  return t_pdl_on;
}

std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct Tensor {
  long off = -1;  // floats, inside the activation arena (and, mirrored, inside the gradient arena)
  int C = 0, H = 0, W = 0;
  int id = -1;
};

struct ConvLayer {
  long w_off = 0, b_off = 0;
  int cout = 0, cin1 = 0, cin2 = 0, k = 1, dil = 1;   // logical dims; cin2 = second concat source
  int coutp = 0, c1p = 0, c2p = 0;                      // padded dims
  long pk_w = -1, pk_b = -1, pk_d1 = -1, pk_d2 = -1;    // packed: fwd weights, bias, dgrad wrt src1 / src2
  long tc_w = -1, tc_d1 = -1, tc_d2 = -1;               // bf16 tensor-core weight images (bf16 element offsets)
  long t3_w = -1, t3_d1 = -1, t3_d2 = -1;               // same for the kx-folded 3x3 kernel (conv3_tc.cu)
  // dilation >= map size in both directions (the deepest levels of the S5 / S6 models: dilation 16 / 32 on <= 16 x 16 maps):
  // every tap but the centre reads SAME padding only, so the layer IS a 1x1 conv with the centre-tap weights (tc images of that
  // slab: tc_wc / tc_dc) and only the centre tap has a weight gradient -- the other 8 are exactly zero
  bool center = false;
  long tc_wc = -1, tc_dc = -1;
  int pad() const { return ((k - 1) * dil) / 2; }       // SAME "before" padding, model/layers/utils.py:13-18
};

struct DeconvLayer {
  long w_off = 0, b_off = 0;
  int cin = 0, cout = 0, cinp = 0, coutp = 0;
  long pk_phase[4] = {-1, -1, -1, -1}, pk_b = -1, pk_d = -1;
  // tensor-core path: all four sub-pixel phases as ONE 2x2-tap conv with 4*coutp output channels (depth-to-space
  // store), and its data gradient as ONE 2x2-tap conv over the space-to-depth view of dOut
  long pk_m = -1, pk_dm = -1, tc_m = -1, tc_dm = -1;
};

struct AttnLayer {
  ConvLayer fg;   // f | g fused into one 1x1 conv with cout = 2d (its w_off/b_off are f's)
  long g_w_off = 0, g_b_off = 0;
  ConvLayer h;
  int C = 0, d = 0;
};

struct Level {
  ConvLayer conv1, coupl;
  std::vector<ConvLayer> res;
  Tensor z1, y1, rr, cc, pooled;
  std::vector<Tensor> a;
};

struct UpLevel {
  DeconvLayer deconv;
  ConvLayer conv1, coupl;
  std::vector<ConvLayer> res;
  Tensor d, u, ur, uc;
  std::vector<Tensor> a;
};

struct Block {
  std::vector<Level> down;
  std::vector<UpLevel> up;
  AttnLayer attn;
  bool has_attn = false;   // blocks 0 .. num_blocks-2; the last block's attention output is never read
  Tensor fg, hh, att, mrow, zinv, dvec;
  ConvLayer end;
  Tensor logits;
};

// Engine options live in the plan (two plans with different options can run side by side, one host thread per GPU);
// msau_set_option only edits the defaults a plan copies at msau_plan_create, msau_plan_set_option edits one plan.
struct EngineOpts {
  int use_tc = 1;        // tensor_core_conv: tcgen05 kernels where they apply (0 = fp32 CUDA-core cross-check path)
  int structured = 1;    // structured_first_layer: one-hot inputs through the id-gather first layer (first_layer.cu)
  int side_stream = 1;   // wgrad_side_stream: weight-gradient kernels on a plan-owned side stream
  int fuse_mask = 1;     // fuse_relu_mask: gradient writers apply the ReLU mask of the tensor they feed
  int use_pw = 1;        // pointwise_conv: 1x1 convs of the narrow levels on the fp32 streaming kernel (conv1x1.cu)
  int use_c3 = 1;        // conv3_fold: kx-folded 3x3 kernel (conv3_tc.cu) where it applies
  int c3_max = 32;       // conv3_max_channels: ... for at most this many output channels (32: +1.3 % per step since the TMA raw ring; 64 does not fit)
  int lrn_coop = 1;      // 0 = thread-per-pixel LRN kernels only, 1 = lane-cooperative where it wins, 2 = from 8 channels up
  int c3_tma = 1;        // conv3_tma: conv3_tc's raw halo planes by TMA tensor loads (0 = cp.async ring)
  int pdl = 1;           // pdl: programmatic dependent launch of the hot kernels (prologue overlaps the predecessor's tail)
  int use_wg4 = 1;       // wgrad_multi_plane: weight gradients of the >= 16-channel layers on wgrad_tc4 (0 = the per-plane kernels)
};
}  // namespace msau

using namespace msau;

struct MsauPlan {
  MsauConfig cfg;
  EngineOpts opt;
  int B, H, W;
  std::vector<int> Hl, Wl;
  std::vector<Block> blocks;
  std::vector<std::pair<long, long>> params;   // (offset, numel) in state_dict order
  long n_params = 0;
  long packed_floats = 0;
  long act_floats = 0;
  int n_tensors = 0;
  std::vector<PackDesc> descs;
  PackDesc* d_descs = nullptr;
  long pack_blocks = 0;
  long misc_floats = 0;
  long attn_scratch_off = 0;   // floats, inside misc: operand images of the tensor-core attention
  long ids_off = -1;           // floats, inside misc: int16 id map of a one-hot input + the is-one-hot flag (first_layer.cu)
  std::vector<TcPackDesc> tc_descs, t3_descs;
  TcPackDesc* d_tc_descs = nullptr;
  TcPackDesc* d_t3_descs = nullptr;
  long tc_blocks = 0, t3_blocks = 0;
  long tc_elems = 0;      // bf16 elements
  uint16_t* pktc = nullptr;
  // runtime state (set per call)
  float* pk = nullptr;
  float* act = nullptr;
  float* grad = nullptr;
  float* misc = nullptr;
  float* gparams = nullptr;
  std::vector<char> written;
  std::vector<Tensor> all_tensors;
  cudaStream_t st = nullptr;
  // weight gradients run on a side stream (they only feed the optimiser): fork after dY is final, join at the end of backward
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t wst = nullptr;        // stream the weight-gradient kernels of the current call go to
  const int* first_skip = nullptr;   // set by msau_forward when the structured first layer ran (device flag)
  // box-constant input as (row-id map, feature table): x_layout 3
  const float* ftable = nullptr; int ftable_rows = 0;
  float* tabP = nullptr; float* tabH = nullptr; long tab_cap = 0;   // plan-owned scratch: projected taps / row histogram
  const short* first_ids = nullptr;  // ... and the id map it used (plan workspace, or the caller's with x_layout 2)
  int lp = 8;                        // channel pitch of the logits tensors: 8 / 16 / 32 >= n_class
  int* d_flags = nullptr;            // plan-owned sticky device error flags (bit 0: a label was out of range)

  Tensor alloc(int C, int Hh, int Ww) {
    Tensor t;
    t.off = act_floats; t.C = C; t.H = Hh; t.W = Ww; t.id = n_tensors++;
    const long n = (long)B * Hh * Ww * C;
    act_floats += (n + 63) / 64 * 64;
    all_tensors.push_back(t);
    return t;
  }
  long alloc_packed(long n) {
    const long o = packed_floats;
    packed_floats += (n + 63) / 64 * 64;
    return o;
  }
  long add_param(long numel) {
    const long o = n_params;
    params.push_back({o, numel});
    n_params += numel;
    return o;
  }
  float* A(const Tensor& t) const { return act + t.off; }
  float* G(const Tensor& t) const { return grad + t.off; }
  long npix(const Tensor& t) const { return (long)B * t.H * t.W; }
  // whether the gradient of `t` already holds a contribution; marks it written
  int touch(const Tensor& t) {
    const int w = written[t.id];
    written[t.id] = 1;
    return w;
  }
};

namespace msau {

static int pad4(int c) { return round_up(c, 4); }
static int pad8(int c) { return round_up(c, 8); }

// ------------------------------------------------------------------ pack descriptors
static void add_desc(MsauPlan* p, long dst, int TH, int TW, int I, int O, int i_dst0, int o_dst0, int I_log, int O_log, long src,
                     int i_off, int o_off, long s_i, long s_o, int ky0, int kys, int kx0, int kxs, int KW, int TWd = 0, int ty_d0 = 0,
                     int tx_d0 = 0) {
  PackDesc d;
  d.TWd = TWd ? TWd : TW; d.ty_d0 = ty_d0; d.tx_d0 = tx_d0;
  d.dst_off = dst; d.src_off = src; d.TH = TH; d.TW = TW; d.I = I; d.O = O; d.i_dst0 = i_dst0; d.o_dst0 = o_dst0;
  d.I_log = I_log; d.O_log = O_log; d.i_off = i_off; d.o_off = o_off; d.s_i = s_i; d.s_o = s_o;
  d.ky0 = ky0; d.kys = kys; d.kx0 = kx0; d.kxs = kxs; d.KW = KW;
  d.blk0 = p->pack_blocks;
  p->pack_blocks += cdiv((long)TH * TW * I_log * O_log, 256);
  p->descs.push_back(d);
}

static long add_tc(MsauPlan* p, long src_off, int taps, int cin, int coutp) {
  // conv_tc takes up to 128 output columns per launch; a 256-column layer (the S6 model's deepest level) gets two images back
  // to back (tc_half_elems apart) and runs as two launches on the two halves of the output channels
  // (the merged transposed conv of the S6 model has 4 * 128 = 512 phase-major columns: four images)
  if (cin % 8 != 0 || (coutp > 128 && coutp % 128 != 0) || coutp > 512) return -1;
  const int halves = coutp > 128 ? coutp / 128 : 1, cw = coutp / halves;
  long first = -1;
  for (int h = 0; h < halves; ++h) {
    TcPackDesc d;
    d.src_off = src_off; d.dst_off = p->tc_elems; d.taps = taps; d.cin = cin; d.coutp = cw;
    d.src_pitch = coutp; d.col0 = h * cw;
    d.N = cw < 16 ? 16 : round_up(cw, 16);
    const long elems = (long)(cin / 8) * (taps + (taps + 1) / 2) * d.N * 16;
    d.blk0 = p->tc_blocks;
    p->tc_blocks += cdiv(elems / 8, 256);       // one thread per [8 ch] slot (16 bytes)
    p->tc_elems += (elems + 127) / 128 * 128;
    p->tc_descs.push_back(d);
    if (h == 0) first = d.dst_off;
  }
  return first;
}

static long add_tc3(MsauPlan* p, long src_off, int cin, int coutp, int ks = 3) {
  if (cin % 8 != 0 || !(coutp == 8 || coutp == 16 || coutp == 32 || coutp == 64) || (ks == 4 && coutp != 8)) return -1;
  TcPackDesc d;
  d.src_off = src_off; d.dst_off = p->tc_elems; d.taps = ks * ks; d.cin = cin; d.coutp = coutp;
  d.src_pitch = coutp; d.col0 = 0;
  d.N = round_up(ks * coutp, 16);
  const long elems = (long)(cin / 8) * (ks + (ks + 1) / 2) * d.N * 16;
  d.blk0 = p->t3_blocks;
  p->t3_blocks += cdiv(elems / 8, 256);
  p->tc_elems += (elems + 127) / 128 * 128;
  p->t3_descs.push_back(d);
  return d.dst_off;
}

// torch Conv2d weight [cout][cin1+cin2][k][k]:
//   fwd   [tap][c1p + c2p][coutp]                     rows = input channels of [src1 | src2]
//   dgrad [flipped tap][coutp][c_s p] per source s     rows = output channels, cols = that source's channels
static void setup_conv(MsauPlan* p, ConvLayer& L, int cout, int cin1, int cin2, int k, int dil, bool need_d1, bool need_d2,
                       int c1p_override = 0, int coutp_override = 0) {
  L.cout = cout; L.cin1 = cin1; L.cin2 = cin2; L.k = k; L.dil = dil;
  L.coutp = coutp_override ? coutp_override : pad8(cout);
  L.c1p = c1p_override ? c1p_override : pad4(cin1);
  L.c2p = cin2 ? pad4(cin2) : 0;
  const int cin = cin1 + cin2;
  const long kk = (long)k * k;
  L.w_off = p->add_param((long)cout * cin * kk);
  L.b_off = p->add_param(cout);
  const int cinp = L.c1p + L.c2p;
  L.pk_w = p->alloc_packed(kk * cinp * L.coutp);
  add_desc(p, L.pk_w, k, k, cinp, L.coutp, 0, 0, cin1, cout, L.w_off, 0, 0, kk, cin * kk, 0, 1, 0, 1, k);
  if (cin2) add_desc(p, L.pk_w, k, k, cinp, L.coutp, L.c1p, 0, cin2, cout, L.w_off, cin1, 0, kk, cin * kk, 0, 1, 0, 1, k);
  L.pk_b = p->alloc_packed(L.coutp);
  add_desc(p, L.pk_b, 1, 1, 1, L.coutp, 0, 0, 1, cout, L.b_off, 0, 0, 0, 1, 0, 1, 0, 1, 1);
  if (need_d1) {
    L.pk_d1 = p->alloc_packed(kk * L.coutp * L.c1p);
    add_desc(p, L.pk_d1, k, k, L.coutp, L.c1p, 0, 0, cout, cin1, L.w_off, 0, 0, cin * kk, kk, k - 1, -1, k - 1, -1, k);
  }
  if (need_d2 && cin2) {
    L.pk_d2 = p->alloc_packed(kk * L.coutp * L.c2p);
    add_desc(p, L.pk_d2, k, k, L.coutp, L.c2p, 0, 0, cout, cin2, L.w_off, 0, cin1, cin * kk, kk, k - 1, -1, k - 1, -1, k);
  }
  L.tc_w = add_tc(p, L.pk_w, k * k, cinp, L.coutp);
  if (L.pk_d1 >= 0) L.tc_d1 = add_tc(p, L.pk_d1, k * k, L.coutp, L.c1p);
  if (L.pk_d2 >= 0) L.tc_d2 = add_tc(p, L.pk_d2, k * k, L.coutp, L.c2p);
  if ((k == 3 || k == 4) && dil == 1) {
    L.t3_w = add_tc3(p, L.pk_w, cinp, L.coutp, k);
    if (L.pk_d1 >= 0) L.t3_d1 = add_tc3(p, L.pk_d1, L.coutp, L.c1p, k);
    if (L.pk_d2 >= 0) L.t3_d2 = add_tc3(p, L.pk_d2, L.coutp, L.c2p, k);
  }
}

// torch ConvTranspose2d(cin, cout, 3, stride 2, padding 1) weight [cin][cout][3][3]; out[2 iy - 1 + ky] += x[iy] W[ky]
static void setup_deconv(MsauPlan* p, DeconvLayer& L, int cin, int cout) {
  L.cin = cin; L.cout = cout; L.cinp = pad4(cin); L.coutp = pad8(cout);
  L.w_off = p->add_param((long)cin * cout * 9);
  L.b_off = p->add_param(cout);
  // sub-pixel phases: even output rows use ky=1 (input row q); odd rows use ky=2 (row q) and ky=0 (row q+1)
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      const int th = py ? 2 : 1, tw = px ? 2 : 1;
      L.pk_phase[py * 2 + px] = p->alloc_packed((long)th * tw * L.cinp * L.coutp);
      add_desc(p, L.pk_phase[py * 2 + px], th, tw, L.cinp, L.coutp, 0, 0, cin, cout, L.w_off, 0, 0, (long)cout * 9, 9,
               py ? 2 : 1, py ? -2 : 0, px ? 2 : 1, px ? -2 : 0, 3);
    }
  L.pk_b = p->alloc_packed(L.coutp);
  add_desc(p, L.pk_b, 1, 1, 1, L.coutp, 0, 0, 1, cout, L.b_off, 0, 0, 0, 1, 0, 1, 0, 1, 1);
  // dgrad: d_in[iy] = sum_ky dOut[2 iy - 1 + ky] W[ci][co][ky]: stride-2 conv over dOut (rows = co, cols = ci)
  L.pk_d = p->alloc_packed(9L * L.coutp * L.cinp);
  add_desc(p, L.pk_d, 3, 3, L.coutp, L.cinp, 0, 0, cout, cin, L.w_off, 0, 0, 9, (long)cout * 9, 0, 1, 0, 1, 3);
  if (L.cinp % 8 == 0 && L.coutp % 8 == 0 && (4 * L.coutp <= 128 || (4 * L.coutp) % 128 == 0) && 4 * L.coutp <= 512 && L.cinp <= 256) {
    // forward: Wm[ty][tx][ci][(py,px,co)], tap (ty,tx) reads input pixel (q+ty, q'+tx)
    L.pk_m = p->alloc_packed(4L * L.cinp * 4 * L.coutp);
    // dgrad: Wdm[ty][tx][(py,px,co)][ci], tap (ty,tx) reads virtual dOut pixel (q-1+ty, q'-1+tx):
    //   (ty=0,py=1) -> ky=0   (ty=1,py=0) -> ky=1   (ty=1,py=1) -> ky=2   (ty=0,py=0) -> no contribution
    L.pk_dm = p->alloc_packed(4L * 4 * L.coutp * L.cinp);
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const int ph = py * 2 + px, th = py ? 2 : 1, tw = px ? 2 : 1;
        add_desc(p, L.pk_m, th, tw, L.cinp, 4 * L.coutp, 0, ph * L.coutp, cin, cout, L.w_off, 0, 0, (long)cout * 9, 9,
                 py ? 2 : 1, py ? -2 : 0, px ? 2 : 1, px ? -2 : 0, 3, 2, 0, 0);
        add_desc(p, L.pk_dm, th, tw, 4 * L.coutp, L.cinp, ph * L.coutp, 0, cout, cin, L.w_off, 0, 0, 9, (long)cout * 9,
                 py ? 0 : 1, py ? 2 : 0, px ? 0 : 1, px ? 2 : 0, 3, 2, py ? 0 : 1, px ? 0 : 1);
      }
    L.tc_m = add_tc(p, L.pk_m, 4, L.cinp, 4 * L.coutp);
    L.tc_dm = add_tc(p, L.pk_dm, 4, 4 * L.coutp, L.cinp);
  }
}

static void setup_attn(MsauPlan* p, AttnLayer& at, int C) {
  const int d = C / 8;
  at.C = C; at.d = d;
  ConvLayer& L = at.fg;
  L.cout = 2 * d; L.cin1 = C; L.cin2 = 0; L.k = 1; L.dil = 1; L.coutp = pad8(2 * d); L.c1p = C; L.c2p = 0;
  L.w_off = p->add_param((long)d * C); L.b_off = p->add_param(d);          // f.conv.{weight,bias}
  at.g_w_off = p->add_param((long)d * C); at.g_b_off = p->add_param(d);    // g.conv.{weight,bias}
  setup_conv(p, at.h, C, C, 0, 1, 1, true, false);                          // h.conv.{weight,bias}
  L.pk_w = p->alloc_packed((long)C * L.coutp);
  add_desc(p, L.pk_w, 1, 1, C, L.coutp, 0, 0, C, d, L.w_off, 0, 0, 1, C, 0, 1, 0, 1, 1);
  add_desc(p, L.pk_w, 1, 1, C, L.coutp, 0, d, C, d, at.g_w_off, 0, 0, 1, C, 0, 1, 0, 1, 1);
  L.pk_b = p->alloc_packed(L.coutp);
  add_desc(p, L.pk_b, 1, 1, 1, L.coutp, 0, 0, 1, d, L.b_off, 0, 0, 0, 1, 0, 1, 0, 1, 1);
  add_desc(p, L.pk_b, 1, 1, 1, L.coutp, 0, d, 1, d, at.g_b_off, 0, 0, 0, 1, 0, 1, 0, 1, 1);
  L.pk_d1 = p->alloc_packed((long)L.coutp * C);
  add_desc(p, L.pk_d1, 1, 1, L.coutp, C, 0, 0, d, C, L.w_off, 0, 0, C, 1, 0, 1, 0, 1, 1);
  add_desc(p, L.pk_d1, 1, 1, L.coutp, C, d, 0, d, C, at.g_w_off, 0, 0, C, 1, 0, 1, 0, 1, 1);
  L.tc_w = add_tc(p, L.pk_w, 1, C, L.coutp);
  L.tc_d1 = add_tc(p, L.pk_d1, 1, L.coutp, C);
}

// ------------------------------------------------------------------ launch helpers
static EngineOpts g_defaults;
static std::mutex g_defaults_mu;

static int set_opt(EngineOpts& o, const char* name, int value) {
  if (!strcmp(name, "tensor_core_conv")) { o.use_tc = value != 0; return MSAU_OK; }
  if (!strcmp(name, "conv3_fold")) { o.use_c3 = value != 0; return MSAU_OK; }
  if (!strcmp(name, "pointwise_conv")) { o.use_pw = value != 0; return MSAU_OK; }
  if (!strcmp(name, "wgrad_side_stream")) { o.side_stream = value != 0; return MSAU_OK; }
  if (!strcmp(name, "structured_first_layer")) { o.structured = value != 0; return MSAU_OK; }
  if (!strcmp(name, "fuse_relu_mask")) { o.fuse_mask = value != 0; return MSAU_OK; }
  if (!strcmp(name, "lrn_coop")) { o.lrn_coop = value; return MSAU_OK; }
  if (!strcmp(name, "conv3_max_channels")) { o.c3_max = value; return MSAU_OK; }
  if (!strcmp(name, "pdl")) { o.pdl = value != 0; return MSAU_OK; }
  if (!strcmp(name, "conv3_tma")) { o.c3_tma = value != 0; return MSAU_OK; }
  if (!strcmp(name, "wgrad_multi_plane")) { o.use_wg4 = value != 0; return MSAU_OK; }
  set_error("set_option: unknown option '%s'", name);
  return MSAU_ERR_ARG;
}

struct ConvOpt {
  bool relu1 = false, relu = false, relu2 = false;
  int accumulate = 0;
  const float* mask1 = nullptr; int pm1 = 0;
  const float* res = nullptr; int pr = 0;
  const float* omask = nullptr; int pom = 0;
  const float* add = nullptr; int pa = 0;
  const float* addmask = nullptr; int pam = 0;
};

// same-size stride-1 convolution (forward of a layer, or a dgrad with flipped packed weights)
static int conv_same(MsauPlan* p, const float* src1, int c1, int p1, int nchw, int c1_logical, const float* src2, int c2, int p2,
                     const float* w, const float* bias, float* out, int po, int coutp, int H, int W, int k, int dil, int pad,
                     const ConvOpt& o, long tc_off = -1, long t3_off = -1, const int* skip_flag = nullptr) {
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.src1 = src1; a.c1 = c1; a.p1 = p1; a.src1_nchw = nchw; a.c1_logical = c1_logical;
  a.mask1 = o.mask1; a.pm1 = o.pm1; a.relu1 = o.relu1;
  a.src2 = src2; a.c2 = c2; a.p2 = p2;
  a.w = w; a.bias = bias; a.out = out; a.po = po; a.coutp = coutp;
  a.B = p->B; a.Hin = H; a.Win = W; a.Hq = H; a.Wq = W;
  a.kh = k; a.kw = k; a.dil = dil; a.stride = 1; a.pad_t = pad; a.pad_l = pad;
  a.Hout = H; a.Wout = W; a.osy = 1; a.oy0 = 0; a.ox0 = 0;
  a.relu = o.relu; a.res = o.res; a.pr = o.pr; a.relu2 = o.relu2;
  a.omask = o.omask; a.pom = o.pom; a.add = o.add; a.pa = o.pa; a.addmask = o.addmask; a.pam = o.pam;
  a.accumulate = o.accumulate;
  a.skip_flag = skip_flag;
  count_launch(1);
  if (p->opt.use_tc && p->opt.use_pw && k == 1 && conv1x1_supported(a)) return launch_conv1x1(a, p->st);
  if (p->opt.use_tc && p->opt.use_c3 && t3_off >= 0 && coutp <= p->opt.c3_max && conv3_tc_supported(a)) return launch_conv3_tc(a, p->pktc + t3_off, p->st, p->opt.c3_tma);
  if (p->opt.use_tc && tc_off >= 0 && conv_tc_supported(a)) return launch_conv_tc(a, p->pktc + tc_off, p->st);
  if (p->opt.use_tc && tc_off >= 0 && coutp == 256) {     // two launches over the two halves of the output channels (add_tc)
    ConvArgs h0 = a;
    h0.coutp = 128;
    if (conv_tc_supported(h0)) {
      ConvArgs h1 = h0;
      h1.out += 128;
      if (h1.bias) h1.bias += 128;
      if (h1.res) h1.res += 128;
      if (h1.omask) h1.omask += 128;
      if (h1.add) h1.add += 128;
      if (h1.addmask) h1.addmask += 128;
      count_launch(1);
      MSAU_TRY(launch_conv_tc(h0, p->pktc + tc_off, p->st));
      return launch_conv_tc(h1, p->pktc + tc_off + tc_half_elems(k * k, c1 + c2), p->st);
    }
  }
  return launch_conv(a, p->st);
}

static int layer_fwd(MsauPlan* p, const ConvLayer& L, const Tensor& s1, const Tensor* s2, const Tensor& out, const ConvOpt& o) {
  if (L.center)      // centre tap only (see ConvLayer::center): a 1x1 conv on the centre slab of the packed weights
    return conv_same(p, p->A(s1), L.c1p, s1.C, 0, L.c1p, nullptr, 0, 0, p->pk + L.pk_w + 4L * (L.c1p + L.c2p) * L.coutp, p->pk + L.pk_b,
                     p->A(out), out.C, L.coutp, out.H, out.W, 1, 1, 0, o, L.tc_wc, -1);
  return conv_same(p, p->A(s1), L.c1p, s1.C, 0, L.c1p, s2 ? p->A(*s2) : nullptr, s2 ? L.c2p : 0, s2 ? s2->C : 0, p->pk + L.pk_w,
                   p->pk + L.pk_b, p->A(out), out.C, L.coutp, out.H, out.W, L.k, L.dil, L.pad(), o, L.tc_w, L.t3_w);
}

// data gradient of a conv layer wrt source `which` (1 or 2): dY (channels coutp) -> dX
static int layer_dgrad(MsauPlan* p, const ConvLayer& L, int which, const float* dy, int pdy, const float* dymask, int pm,
                       const Tensor& dst, ConvOpt o) {
  const int cs = which == 1 ? L.c1p : L.c2p;
  const long pkd = which == 1 ? L.pk_d1 : L.pk_d2;
  if (pkd < 0) { set_error("internal: dgrad weights missing"); return MSAU_ERR_ARG; }
  o.mask1 = dymask; o.pm1 = pm;
  o.accumulate = p->touch(dst);
  if (L.center && which == 1)
    return conv_same(p, dy, L.coutp, pdy, 0, L.coutp, nullptr, 0, 0, p->pk + pkd + 4L * L.coutp * cs, nullptr, p->G(dst), dst.C, cs, dst.H, dst.W,
                     1, 1, 0, o, L.tc_dc, -1);
  const int padd = (L.k - 1) * L.dil - L.pad();
  return conv_same(p, dy, L.coutp, pdy, 0, L.coutp, nullptr, 0, 0, p->pk + pkd, nullptr, p->G(dst), dst.C, cs, dst.H, dst.W, L.k,
                   L.dil, padd, o, which == 1 ? L.tc_d1 : L.tc_d2, which == 1 ? L.t3_d1 : L.t3_d2);
}

// the weight-gradient kernel about to be launched may start once everything enqueued so far on the main stream is done
static int wgrad_fork(MsauPlan* p) {
  if (p->wst == p->st) return MSAU_OK;
  MSAU_CUDA_TRY(cudaEventRecord(p->ev_fork, p->st));
  MSAU_CUDA_TRY(cudaStreamWaitEvent(p->wst, p->ev_fork, 0));
  return MSAU_OK;
}

// weight (+bias) gradient of a conv layer wrt source `which`
static int layer_wgrad(MsauPlan* p, const ConvLayer& L, int which, const float* src, int psrc, int nchw, int c_logical, bool reluA,
                       const float* dy, int pdy, const float* dymask, int pm, int H, int W, long w_off_override = -1,
                       long b_off_override = -1, int cb_off = 0, int cb = -1, int cb_lim = -1, const int* skip_flag = nullptr) {
  WgradArgs a;
  memset(&a, 0, sizeof(a));
  const int cin = L.cin1 + L.cin2;
  const long kk = (long)L.k * L.k;
  a.A = src; a.ca = which == 1 ? L.c1p : L.c2p; a.pa = psrc; a.a_nchw = nchw; a.ca_logical = c_logical; a.reluA = reluA;
  a.Ha = H; a.Wa = W; a.sa = 1; a.dila = L.dil; a.pada_t = L.pad(); a.pada_l = L.pad();
  a.Bm = dy + cb_off; a.cb = cb < 0 ? L.coutp : cb; a.pb = pdy; a.maskB = dymask ? dymask + cb_off : nullptr; a.pmb = pm;
  a.Hb = H; a.Wb = W; a.sb = 1; a.dilb = 0; a.padb_t = 0; a.padb_l = 0;
  a.B = p->B; a.Hq = H; a.Wq = W; a.kh = L.k; a.kw = L.k;
  const long w_off = w_off_override >= 0 ? w_off_override : L.w_off;
  a.dW = p->gparams + w_off + (which == 2 ? (long)L.cin1 * kk : 0);
  a.s_ca = kk; a.s_cb = cin * kk;
  a.ca_lim = which == 1 ? L.cin1 : L.cin2;
  a.cb_lim = cb_lim < 0 ? L.cout : cb_lim;
  a.dbias = which == 1 ? p->gparams + (b_off_override >= 0 ? b_off_override : L.b_off) : nullptr;
  a.skip_flag = skip_flag;
  a.no_tc4 = !p->opt.use_wg4;
  if (L.center) {     // only the centre tap (index 4 of the 3x3) sees anything but padding
    a.kh = a.kw = 1; a.dila = 1; a.pada_t = a.pada_l = 0;
    a.dW += 4;
  }
  if (p->opt.use_tc && a.cb == 256 && cb < 0 && !wgrad_tc_supported(a)) {
    // the tensor-core kernels take up to 128 output channels: run the two halves of a 256-channel dY separately
    WgradArgs h = a;
    h.cb = 128;
    if (wgrad_tc_supported(h)) {
      const long base_w = w_off_override >= 0 ? w_off_override : L.w_off, base_b = b_off_override >= 0 ? b_off_override : L.b_off;
      for (int hf = 0; hf < 2; ++hf) {
        const int lim = L.cout - 128 * hf < 0 ? 0 : (L.cout - 128 * hf > 128 ? 128 : L.cout - 128 * hf);
        MSAU_TRY(layer_wgrad(p, L, which, src, psrc, nchw, c_logical, reluA, dy, pdy, dymask, pm, H, W, base_w + 128L * hf * cin * kk,
                             base_b + 128 * hf, cb_off + 128 * hf, 128, lim, skip_flag));
      }
      return MSAU_OK;
    }
  }
  count_launch(1);
  MSAU_TRY(wgrad_fork(p));
  if (p->opt.use_tc && wgrad_tc_supported(a)) return launch_wgrad_tc(a, p->wst);
  return launch_wgrad(a, p->wst);
}

// MultiConvResidualBlock forward, model/model.py:37-50
static int res_fwd(MsauPlan* p, const std::vector<ConvLayer>& res, const Tensor& in, const std::vector<Tensor>& a, const Tensor& out) {
  const int R = (int)res.size();
  for (int r = 0; r < R; ++r) {
    ConvOpt o;
    const Tensor& src = r == 0 ? in : a[r - 1];
    o.relu1 = (r == 0);
    if (r < R - 1) {
      o.relu = true;
      MSAU_TRY(layer_fwd(p, res[r], src, nullptr, a[r], o));
    } else {
      o.res = p->A(in); o.pr = in.C; o.relu2 = true;
      MSAU_TRY(layer_fwd(p, res[r], src, nullptr, out, o));
    }
  }
  return MSAU_OK;
}

// backward of the residual block: G(out) -> G(in) (overwritten: `in` has no other consumer) + weight grads
static int res_bwd(MsauPlan* p, const std::vector<ConvLayer>& res, const Tensor& in, const std::vector<Tensor>& a, const Tensor& out,
                   bool premasked = false) {
  const int R = (int)res.size();
  // gradient through the block's final ReLU, materialised in place: it feeds the last conv's wgrad and dgrad and
  // the identity skip path.  premasked: the single kernel that wrote G(out) already applied the mask in its epilogue
  if (!premasked) {
    count_launch(1);
    MSAU_TRY(launch_relu_mask(p->G(out), p->A(out), p->npix(out) * out.C, p->st));
  }
  for (int r = R - 1; r >= 0; --r) {
    const bool last = (r == R - 1);
    // dY of conv r: last conv -> G(out) (masked above); inner convs -> G(a[r]) (masked when produced)
    const float* dy = last ? p->G(out) : p->G(a[r]);
    const int pdy = last ? out.C : a[r].C;
    const Tensor& src = r == 0 ? in : a[r - 1];
    MSAU_TRY(layer_wgrad(p, res[r], 1, p->A(src), src.C, 0, res[r].c1p, r == 0, dy, pdy, nullptr, 0, in.H, in.W));
    ConvOpt o;
    if (r > 0) {
      o.omask = p->A(a[r - 1]); o.pom = a[r - 1].C;          // relu after conv r-1
      MSAU_TRY(layer_dgrad(p, res[r], 1, dy, pdy, nullptr, 0, a[r - 1], o));
    } else {
      o.omask = p->A(in); o.pom = in.C;                       // relu applied to the block input
      o.add = p->G(out); o.pa = out.C;                        // identity skip path (already masked)
      MSAU_TRY(layer_dgrad(p, res[r], 1, dy, pdy, nullptr, 0, in, o));
    }
  }
  return MSAU_OK;
}

// coupling 1x1 conv on cat[prev, cur] + ReLU (model/model.py:143-148, 246-252)
// `cur` is the output of a residual block whose only reader is this conv: its gradient has this one writer, which therefore
// also applies the block's final ReLU mask (omask = A(cur)) -- *mask_cur tells res_bwd to skip its relu_mask pass
// mask_prev: `prev` is a post-ReLU tensor all of whose gradient writers mask their own contribution ((a + b) m = a m + b m)
static int coupl_bwd(MsauPlan* p, const ConvLayer& L, const Tensor& prev, const Tensor& cur, const Tensor& out, bool* mask_cur,
                     bool out_premasked = false, bool mask_prev = false) {
  if (!out_premasked) {
    count_launch(1);
    MSAU_TRY(launch_relu_mask(p->G(out), p->A(out), p->npix(out) * out.C, p->st));   // 4 consumers below
  }
  const float* dy = p->G(out);
  MSAU_TRY(layer_wgrad(p, L, 1, p->A(prev), prev.C, 0, L.c1p, false, dy, out.C, nullptr, 0, out.H, out.W));
  MSAU_TRY(layer_wgrad(p, L, 2, p->A(cur), cur.C, 0, L.c2p, false, dy, out.C, nullptr, 0, out.H, out.W));
  ConvOpt o;
  if (mask_prev) { o.omask = p->A(prev); o.pom = prev.C; }
  MSAU_TRY(layer_dgrad(p, L, 1, dy, out.C, nullptr, 0, prev, o));
  o.omask = nullptr; o.pom = 0;
  *mask_cur = p->opt.fuse_mask && !p->written[cur.id];
  if (*mask_cur) { o.omask = p->A(cur); o.pom = cur.C; }
  MSAU_TRY(layer_dgrad(p, L, 2, dy, out.C, nullptr, 0, cur, o));
  return MSAU_OK;
}

static int deconv_fwd(MsauPlan* p, const DeconvLayer& L, const Tensor& in, const Tensor& out) {
  if (p->opt.use_tc && L.tc_m >= 0) {
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.src1 = p->A(in); a.c1 = L.cinp; a.p1 = in.C; a.c1_logical = L.cinp;
    a.w = p->pk + L.pk_m; a.bias = p->pk + L.pk_b;
    a.out = p->A(out); a.po = out.C; a.coutp = 4 * L.coutp;
    a.B = p->B; a.Hin = in.H; a.Win = in.W; a.Hq = in.H; a.Wq = in.W;
    a.kh = 2; a.kw = 2; a.dil = 1; a.stride = 1; a.pad_t = 0; a.pad_l = 0;
    a.Hout = out.H; a.Wout = out.W; a.osy = 1;
    a.d2s = 1; a.cph = L.coutp;
    if (conv_tc_supported(a)) {
      count_launch(1);
      return launch_conv_tc(a, p->pktc + L.tc_m, p->st);
    }
    if (a.coutp > 128) {
      // more than 128 phase-major columns (64- / 128-channel transposed convs of the S5 / S6 models): windows of 128 columns,
      // one weight image each (add_tc); the depth-to-space store maps a window's columns back to (phase, channel)
      ConvArgs w0 = a;
      w0.coutp = 128;
      if (conv_tc_supported(w0)) {
        const int nw = a.coutp / 128;
        for (int w = 0; w < nw; ++w) {
          ConvArgs aw = w0;
          aw.d2s_col0 = 128 * w;
          count_launch(1);
          MSAU_TRY(launch_conv_tc(aw, p->pktc + L.tc_m + (long)w * tc_half_elems(4, L.cinp), p->st));
        }
        return MSAU_OK;
      }
    }
  }
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      ConvArgs a;
      memset(&a, 0, sizeof(a));
      a.src1 = p->A(in); a.c1 = L.cinp; a.p1 = in.C; a.c1_logical = L.cinp;
      a.w = p->pk + L.pk_phase[py * 2 + px]; a.bias = p->pk + L.pk_b;
      a.out = p->A(out); a.po = out.C; a.coutp = L.coutp;
      a.B = p->B; a.Hin = in.H; a.Win = in.W;
      a.Hq = (out.H - py + 1) / 2; a.Wq = (out.W - px + 1) / 2;
      if (a.Hq <= 0 || a.Wq <= 0) continue;
      a.kh = py ? 2 : 1; a.kw = px ? 2 : 1; a.dil = 1; a.stride = 1; a.pad_t = 0; a.pad_l = 0;
      a.Hout = out.H; a.Wout = out.W; a.osy = 2; a.oy0 = py; a.ox0 = px;
      count_launch(1);
      MSAU_TRY(launch_conv(a, p->st));
    }
  return MSAU_OK;
}

static int deconv_bwd(MsauPlan* p, const DeconvLayer& L, const Tensor& in, const Tensor& out, bool mask_in = false) {
  // weights: dW[ci][co][ky][kx] = sum_q x[q][ci] * dOut[2q - 1 + (ky,kx)][co]
  bool w_done = false;
  if (p->opt.use_tc && L.tc_m >= 0) {
    // = weight gradient of the merged 2x2-tap forward conv: A = x (taps (ty,tx) read x[q + (ty,tx)]), B = space-to-depth
    // view of dOut; the kernel maps (tap, phase) back to (ky, kx) and folds the bias gradient into the dOut loader
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.A = p->A(in); a.ca = L.cinp; a.pa = in.C; a.ca_logical = L.cinp;
    a.Ha = in.H; a.Wa = in.W; a.sa = 1; a.dila = 1; a.pada_t = 0; a.pada_l = 0;
    a.Bm = p->G(out); a.cb = 4 * L.coutp; a.pb = out.C;
    a.Hb = out.H; a.Wb = out.W; a.sb = 1; a.dilb = 0; a.padb_t = 0; a.padb_l = 0;
    a.b_s2d = 1; a.cph = L.coutp;
    a.B = p->B; a.Hq = in.H; a.Wq = in.W; a.kh = 2; a.kw = 2;
    a.dW = p->gparams + L.w_off; a.s_ca = (long)L.cout * 9; a.s_cb = 9; a.ca_lim = L.cin; a.cb_lim = L.cout;
    a.dbias = p->gparams + L.b_off;
    a.no_tc4 = !p->opt.use_wg4;
    if (wgrad_tc_supported(a)) {
      count_launch(1);
      MSAU_TRY(wgrad_fork(p));
      MSAU_TRY(launch_wgrad_tc(a, p->wst));
      w_done = true;
    } else if (a.cb > 128 && a.cb % 128 == 0) {
      WgradArgs w0 = a;            // windows of 128 of the 4 * coutp phase-major dY columns (tensor-core kernels take <= 128)
      w0.cb = 128;
      if (wgrad_tc_supported(w0)) {
        for (int w = 0; w < a.cb / 128; ++w) {
          WgradArgs aw = w0;
          aw.b_col0 = 128 * w;
          count_launch(1);
          MSAU_TRY(wgrad_fork(p));
          MSAU_TRY(launch_wgrad_tc(aw, p->wst));
        }
        w_done = true;
      }
    }
  }
  if (!w_done) {
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.A = p->A(in); a.ca = L.cinp; a.pa = in.C; a.ca_logical = L.cinp;
    a.Ha = in.H; a.Wa = in.W; a.sa = 1; a.dila = 0; a.pada_t = 0; a.pada_l = 0;
    a.Bm = p->G(out); a.cb = L.coutp; a.pb = out.C;
    a.Hb = out.H; a.Wb = out.W; a.sb = 2; a.dilb = 1; a.padb_t = 1; a.padb_l = 1;
    a.B = p->B; a.Hq = in.H; a.Wq = in.W; a.kh = 3; a.kw = 3;
    a.dW = p->gparams + L.w_off; a.s_ca = (long)L.cout * 9; a.s_cb = 9; a.ca_lim = L.cin; a.cb_lim = L.cout;
    a.dbias = nullptr;
    count_launch(1);
    MSAU_TRY(wgrad_fork(p));
    MSAU_TRY(launch_wgrad(a, p->wst));
  }
  if (!w_done) {
    count_launch(1);
    MSAU_TRY(launch_colsum(p->G(out), p->npix(out), out.C, L.cout, p->gparams + L.b_off, p->st));
  }
  if (p->opt.use_tc && L.tc_dm >= 0) {
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.src1 = p->G(out); a.c1 = 4 * L.coutp; a.p1 = out.C; a.c1_logical = a.c1;
    a.s2d = 1; a.cph = L.coutp; a.Hs = out.H; a.Ws = out.W;
    a.w = p->pk + L.pk_dm; a.bias = nullptr;
    a.out = p->G(in); a.po = in.C; a.coutp = L.cinp;
    a.B = p->B; a.Hin = in.H; a.Win = in.W; a.Hq = in.H; a.Wq = in.W;
    a.kh = 2; a.kw = 2; a.dil = 1; a.stride = 1; a.pad_t = 1; a.pad_l = 1;
    a.Hout = in.H; a.Wout = in.W; a.osy = 1;
    if (mask_in) { a.omask = p->A(in); a.pom = in.C; }
    if (conv_tc_supported(a)) {
      a.accumulate = p->touch(in);
      count_launch(1);
      return launch_conv_tc(a, p->pktc + L.tc_dm, p->st);
    }
    if (a.coutp == 256) {          // 256 input channels of the transposed conv: two launches over the two halves (add_tc)
      ConvArgs h0 = a;
      h0.coutp = 128;
      if (conv_tc_supported(h0)) {
        h0.accumulate = p->touch(in);
        ConvArgs h1 = h0;
        h1.out += 128;
        if (h1.omask) h1.omask += 128;
        count_launch(2);
        MSAU_TRY(launch_conv_tc(h0, p->pktc + L.tc_dm, p->st));
        return launch_conv_tc(h1, p->pktc + L.tc_dm + tc_half_elems(4, 4 * L.coutp), p->st);
      }
    }
  }
  // data: stride-2 gather over dOut
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.src1 = p->G(out); a.c1 = L.coutp; a.p1 = out.C; a.c1_logical = L.coutp;
  a.w = p->pk + L.pk_d; a.bias = nullptr;
  a.out = p->G(in); a.po = in.C; a.coutp = L.cinp;
  a.B = p->B; a.Hin = out.H; a.Win = out.W; a.Hq = in.H; a.Wq = in.W;
  a.kh = 3; a.kw = 3; a.dil = 1; a.stride = 2; a.pad_t = 1; a.pad_l = 1;
  a.Hout = in.H; a.Wout = in.W; a.osy = 1;
  if (mask_in) { a.omask = p->A(in); a.pom = in.C; }
  a.accumulate = p->touch(in);
  count_launch(1);
  return launch_conv(a, p->st);
}

static size_t ws_bytes(const MsauPlan* p, int training) {
  const long floats = p->packed_floats + (p->tc_elems + 1) / 2 + 64 + p->act_floats * (training ? 2 : 1) + p->misc_floats;
  return (size_t)floats * sizeof(float);
}

static int bind(MsauPlan* p, void* ws, size_t bytes, int training, void* stream) {
  MSAU_CHECK_ARG(ws != nullptr, "workspace is null");
  if (bytes < ws_bytes(p, training)) {
    set_error("workspace too small: %zu < %zu bytes", bytes, ws_bytes(p, training));
    return MSAU_ERR_WORKSPACE;
  }
  MSAU_CHECK_ARG(((uintptr_t)ws & 255) == 0, "workspace must be 256-byte aligned");
  p->pk = reinterpret_cast<float*>(ws);
  p->pktc = reinterpret_cast<uint16_t*>(p->pk + p->packed_floats);
  p->act = p->pk + p->packed_floats + ((p->tc_elems + 1) / 2 + 63) / 64 * 64;
  p->grad = training ? p->act + p->act_floats : nullptr;
  p->misc = p->act + p->act_floats * (training ? 2 : 1);
  p->st = reinterpret_cast<cudaStream_t>(stream);
  return MSAU_OK;
}

}  // namespace msau

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" const char* msau_last_error(void) { return msau::get_error(); }
extern "C" int msau_version(void) { return 100; }
extern "C" long long msau_launch_count(void) { return msau::g_launches.load(); }
extern "C" void msau_launch_count_add(long long n) { msau::g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int msau_plan_create(const MsauConfig* cfg, int batch, int height, int width, MsauPlan** out) {
  MSAU_CHECK_ARG(cfg && out, "plan_create: null argument");
  MSAU_CHECK_ARG(batch >= 1 && height >= 1 && width >= 1, "plan_create: bad shape %dx%dx%d", batch, height, width);
  MSAU_CHECK_ARG(cfg->filter_size == 3 && cfg->pool_size == 2, "plan_create: only filter_size=3, pool_size=2 are supported");
  MSAU_CHECK_ARG(cfg->num_blocks >= 2 && cfg->num_blocks <= 8, "plan_create: num_blocks must be in [2,8]");
  MSAU_CHECK_ARG(cfg->scale_space_num >= 2 && cfg->res_depth >= 1, "plan_create: need scale_space_num >= 2, res_depth >= 1");
  MSAU_CHECK_ARG(cfg->feat_root >= 8 && cfg->feat_root % 8 == 0, "plan_create: feat_root must be a multiple of 8");
  MSAU_CHECK_ARG(cfg->n_class >= 2 && cfg->n_class <= 32, "plan_create: n_class must be in [2,32]");
  MSAU_CHECK_ARG(cfg->channels >= 1, "plan_create: channels must be >= 1");
  const int S = cfg->scale_space_num, R = cfg->res_depth, NB = cfg->num_blocks;
  const int fa = cfg->feat_root << (S - 1), da = fa / 8;
  if (!attn_supported(fa, da)) {
    set_error("plan_create: attention at %d channels (d=%d) is not supported (supported: 32, 64, 128, 256 channels)", fa, da);
    return MSAU_ERR_UNSUPPORTED;
  }
  for (int l = 0; l < S; ++l) {
    const int f = cfg->feat_root << l;
    if (!(f == 8 || f == 16 || f == 32 || f == 64 || f == 128 || f == 256)) {
      set_error("plan_create: level %d has %d channels; LRN kernels cover 8/16/32/64/128/256", l, f);
      return MSAU_ERR_UNSUPPORTED;
    }
  }
  MsauPlan* p = new MsauPlan();
  p->cfg = *cfg; p->B = batch; p->H = height; p->W = width;
  { std::lock_guard<std::mutex> lk(g_defaults_mu); p->opt = g_defaults; }
  p->lp = cfg->n_class <= 8 ? 8 : (cfg->n_class <= 16 ? 16 : 32);
  p->Hl.resize(S); p->Wl.resize(S);
  p->Hl[0] = height; p->Wl[0] = width;
  for (int l = 1; l < S; ++l) { p->Hl[l] = (p->Hl[l - 1] + 1) / 2; p->Wl[l] = (p->Wl[l - 1] + 1) / 2; }
  p->blocks.resize(NB);
  // ---- parameters in state_dict order (down tower conv_res_list, conv1s, conv1_1s,
  //      layer_attentions f,g,h; up tower conv_res_list, conv1s, conv1_1s, deconvs; then end_convs) ----
  for (int b = 0; b < NB; ++b) {
    Block& blk = p->blocks[b];
    blk.down.resize(S); blk.up.resize(S - 1);
    blk.has_attn = b < NB - 1;
    const int cin0 = b == 0 ? cfg->channels : cfg->n_class;
    for (int l = 0; l < S; ++l) {
      const int f = cfg->feat_root << l;
      blk.down[l].res.resize(R);
      for (int r = 0; r < R; ++r) setup_conv(p, blk.down[l].res[r], f, f, 0, 3, 1, true, false);
    }
    for (int l = 0; l < S; ++l) {
      const int f = cfg->feat_root << l;
      const int cin = l == 0 ? cin0 : (cfg->feat_root << (l - 1));
      // blocks >= 1 read the previous block's logits, stored with pitch lp
      setup_conv(p, blk.down[l].conv1, f, cin, 0, 3, 1 << l, !(b == 0 && l == 0), false, (b > 0 && l == 0) ? p->lp : 0);
      ConvLayer& K1 = blk.down[l].conv1;
      if (l > 0 && (1 << l) >= p->Hl[l] && (1 << l) >= p->Wl[l]) {
        K1.center = true;
        K1.tc_wc = add_tc(p, K1.pk_w + 4L * (K1.c1p + K1.c2p) * K1.coutp, 1, K1.c1p + K1.c2p, K1.coutp);
        if (K1.pk_d1 >= 0) K1.tc_dc = add_tc(p, K1.pk_d1 + 4L * K1.coutp * K1.c1p, 1, K1.coutp, K1.c1p);
      }
    }
    if (b > 0)
      for (int l = 0; l < S; ++l) {
        const int f = cfg->feat_root << l;
        setup_conv(p, blk.down[l].coupl, f, f, f, 1, 1, true, true);
      }
    setup_attn(p, blk.attn, fa);
    for (int l = 0; l < S - 1; ++l) {
      const int f = cfg->feat_root << l;
      blk.up[l].res.resize(R);
      for (int r = 0; r < R; ++r) setup_conv(p, blk.up[l].res[r], f, f, 0, 3, 1, true, false);
    }
    for (int l = 0; l < S - 1; ++l) {
      const int f = cfg->feat_root << l;
      setup_conv(p, blk.up[l].conv1, f, f, f, 3, 1, true, true);
    }
    if (b > 0)
      for (int l = 0; l < S - 1; ++l) {
        const int f = cfg->feat_root << l;
        setup_conv(p, blk.up[l].coupl, f, f, f, 1, 1, true, true);
      }
    for (int l = 0; l < S - 1; ++l) {
      const int f = cfg->feat_root << l;
      setup_deconv(p, blk.up[l].deconv, 2 * f, f);
    }
  }
  for (int b = 0; b < NB; ++b) setup_conv(p, p->blocks[b].end, cfg->n_class, cfg->feat_root, 0, 4, 1, true, false, 0, p->lp);

  // ---- activations ----
  for (int b = 0; b < NB; ++b) {
    Block& blk = p->blocks[b];
    for (int l = 0; l < S; ++l) {
      Level& L = blk.down[l];
      const int f = cfg->feat_root << l, Hh = p->Hl[l], Ww = p->Wl[l];
      L.z1 = p->alloc(f, Hh, Ww); L.y1 = p->alloc(f, Hh, Ww);
      for (int r = 0; r + 1 < R; ++r) L.a.push_back(p->alloc(f, Hh, Ww));
      L.rr = p->alloc(f, Hh, Ww);
      L.cc = b > 0 ? p->alloc(f, Hh, Ww) : L.rr;
      if (l < S - 1) L.pooled = p->alloc(f, p->Hl[l + 1], p->Wl[l + 1]);
    }
    if (blk.has_attn) {
      const int Hh = p->Hl[S - 1], Ww = p->Wl[S - 1];
      blk.fg = p->alloc(blk.attn.fg.coutp, Hh, Ww);
      blk.hh = p->alloc(fa, Hh, Ww);
      blk.att = p->alloc(fa, Hh, Ww);
      blk.mrow = p->alloc(1, Hh, Ww); blk.zinv = p->alloc(1, Hh, Ww); blk.dvec = p->alloc(1, Hh, Ww);
    }
    for (int l = S - 2; l >= 0; --l) {
      UpLevel& U = blk.up[l];
      const int f = cfg->feat_root << l, Hh = p->Hl[l], Ww = p->Wl[l];
      U.d = p->alloc(f, Hh, Ww); U.u = p->alloc(f, Hh, Ww);
      for (int r = 0; r + 1 < R; ++r) U.a.push_back(p->alloc(f, Hh, Ww));
      U.ur = p->alloc(f, Hh, Ww);
      U.uc = b > 0 ? p->alloc(f, Hh, Ww) : U.ur;
    }
    blk.logits = p->alloc(p->lp, height, width);
  }
  p->written.assign(p->n_tensors, 0);
  p->misc_floats = round_up(2 * loss_partial_count(batch, (long)height * width) + round_up((int)loss_scratch_ints(batch), 64) + 2048, 64);
  p->attn_scratch_off = p->misc_floats;
  if (attn_tc_supported(fa, da))
    p->misc_floats += (long)((attn_tc_scratch_bytes(batch, p->Hl[S - 1] * p->Wl[S - 1], fa) + 255) / 256 * 64);
  {
    const ConvLayer& c0 = p->blocks[0].down[0].conv1;
    if (c0.coutp == 8 && c0.cout <= 8 && ((long)height * width) % 4 == 0 && 9L * cfg->channels * 32 <= 200 * 1024 && cfg->channels < 32768) {
      p->ids_off = p->misc_floats;
      p->misc_floats += round_up((int)(((long)batch * height * width + 1) / 2) + 64, 64);
    }
  }
  // descriptor table: the only device memory the plan owns
  cudaError_t e = cudaMalloc(&p->d_descs, sizeof(PackDesc) * p->descs.size());
  if (e == cudaSuccess) e = cudaMalloc(&p->d_flags, 64);
  if (e == cudaSuccess) e = cudaMemset(p->d_flags, 0, 64);
  if (e == cudaSuccess) e = cudaMemcpy(p->d_descs, p->descs.data(), sizeof(PackDesc) * p->descs.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !p->tc_descs.empty()) {
    e = cudaMalloc(&p->d_tc_descs, sizeof(TcPackDesc) * p->tc_descs.size());
    if (e == cudaSuccess)
      e = cudaMemcpy(p->d_tc_descs, p->tc_descs.data(), sizeof(TcPackDesc) * p->tc_descs.size(), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess && !p->t3_descs.empty()) {
    e = cudaMalloc(&p->d_t3_descs, sizeof(TcPackDesc) * p->t3_descs.size());
    if (e == cudaSuccess)
      e = cudaMemcpy(p->d_t3_descs, p->t3_descs.data(), sizeof(TcPackDesc) * p->t3_descs.size(), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    set_error("plan_create: descriptor upload failed: %s", cudaGetErrorString(e));
    if (p->d_descs) cudaFree(p->d_descs);
    if (p->d_flags) cudaFree(p->d_flags);
    if (p->d_tc_descs) cudaFree(p->d_tc_descs);
    if (p->d_t3_descs) cudaFree(p->d_t3_descs);
    delete p;
    return MSAU_ERR_CUDA;
  }
  *out = p;
  return MSAU_OK;
}

extern "C" void msau_plan_destroy(MsauPlan* p) {
  if (!p) return;
  if (p->d_descs) cudaFree(p->d_descs);
  if (p->d_flags) cudaFree(p->d_flags);
  if (p->d_tc_descs) cudaFree(p->d_tc_descs);
  if (p->d_t3_descs) cudaFree(p->d_t3_descs);
  if (p->tabP) cudaFree(p->tabP);
  if (p->tabH) cudaFree(p->tabH);
  if (p->side) cudaStreamDestroy(p->side);
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_join) cudaEventDestroy(p->ev_join);
  delete p;
}

extern "C" int msau_plan_set_feature_table(MsauPlan* p, const float* table, int rows) {
  MSAU_CHECK_ARG(p && table && rows >= 1 && rows <= 32767, "set_feature_table: need a device table with 1..32767 rows");
  if (rows > p->tab_cap) {     // setup path, not the data path: grow the plan-owned scratch geometrically
    long cap = p->tab_cap ? p->tab_cap : 1024;
    while (cap < rows) cap *= 2;
    if (p->tabP) cudaFree(p->tabP);
    if (p->tabH) cudaFree(p->tabH);
    p->tabP = p->tabH = nullptr;
    MSAU_CUDA_TRY(cudaMalloc(&p->tabP, sizeof(float) * 72 * cap));
    MSAU_CUDA_TRY(cudaMalloc(&p->tabH, sizeof(float) * 72 * cap));
    p->tab_cap = cap;
  }
  p->ftable = table; p->ftable_rows = rows;
  return MSAU_OK;
}

extern "C" long long msau_param_count(const MsauPlan* p) { return p ? p->n_params : 0; }

extern "C" int msau_param_info(const MsauPlan* p, int idx, long long* offset, long long* numel) {
  MSAU_CHECK_ARG(p && idx >= 0 && idx < (int)p->params.size(), "param_info: index %d out of range", idx);
  if (offset) *offset = p->params[idx].first;
  if (numel) *numel = p->params[idx].second;
  return MSAU_OK;
}

extern "C" int msau_workspace_bytes(const MsauPlan* p, int training, size_t* bytes) {
  MSAU_CHECK_ARG(p && bytes, "workspace_bytes: null argument");
  *bytes = ws_bytes(p, training);
  return MSAU_OK;
}

extern "C" int msau_forward(MsauPlan* p, const float* x, int x_layout, const float* params, void* workspace, size_t workspace_bytes,
                            int training, float* logits, float* aux, float* probs, uint8_t* argmax, void* stream) {
  MSAU_CHECK_ARG(p && x && params, "forward: null argument");
  MSAU_TRY(bind(p, workspace, workspace_bytes, training, stream));
  const MsauConfig& cfg = p->cfg;
  const int S = cfg.scale_space_num, NB = cfg.num_blocks;
  // repack the (possibly just updated) parameters into kernel layouts
  MSAU_CUDA_TRY(cudaMemsetAsync(p->pk, 0, sizeof(float) * p->packed_floats, p->st));
  count_launch(1);
  MSAU_TRY(launch_pack(params, p->pk, p->d_descs, (int)p->descs.size(), p->pack_blocks, p->st));
  if (p->opt.use_tc) {
    count_launch(1);
    MSAU_TRY(launch_pack_tc(p->pk, p->pktc, p->d_tc_descs, (int)p->tc_descs.size(), p->tc_blocks, p->st));
    if (!p->t3_descs.empty()) {
      count_launch(1);
      MSAU_TRY(launch_pack_tc3(p->pk, p->pktc, p->d_t3_descs, (int)p->t3_descs.size(), p->t3_blocks, p->st));
    }
  }

  // the two launches after the packing kernels stay fully serialised (a PDL prologue reads packed weights: see common.cuh)
  pdl_set(p->opt.pdl != 0, 2);
  struct PdlOff { ~PdlOff() { pdl_set(false, 0); } } pdl_off;
  for (int b = 0; b < NB; ++b) {
    Block& blk = p->blocks[b];
    Block* prev = b > 0 ? &p->blocks[b - 1] : nullptr;
    // ---- down tower, model/model.py:129-164 ----
    for (int l = 0; l < S; ++l) {
      Level& L = blk.down[l];
      ConvOpt o;
      if (b == 0 && l == 0) {
        const int c1 = pad4(cfg.channels);
        const int* skip = nullptr;
        p->first_ids = nullptr;
        if (x_layout == 3) {
          // BERT grid as (row-id map, feature table): projected taps + id-gather instead of the 768-channel dense conv
          if (!p->ftable || L.conv1.coutp != 8) { set_error("forward: x_layout 3 needs msau_plan_set_feature_table() and an 8-channel first layer"); return MSAU_ERR_ARG; }
          count_launch(2);
          MSAU_TRY(launch_table_first_fwd(reinterpret_cast<const short*>(x), p->ftable, p->ftable_rows, cfg.channels, c1, p->pk + L.conv1.pk_w,
                                          p->pk + L.conv1.pk_b, p->tabP, p->B, p->H, p->W, p->A(L.z1), L.z1.C, p->st));
          p->first_skip = nullptr;
          p->first_ids = reinterpret_cast<const short*>(x);
        } else if (x_layout == 2) {
          // the caller already holds the id map (e.g. from the rasteriser, layout 2): no dense tensor exists at all
          if (p->ids_off < 0) { set_error("forward: x_layout 2 (id map) needs a first layer with 8 output channels"); return MSAU_ERR_UNSUPPORTED; }
          int* flag = reinterpret_cast<int*>(p->misc + p->ids_off) + (((long)p->B * p->H * p->W + 1) / 2 + 8);
          MSAU_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), p->st));
          count_launch(1);
          MSAU_TRY(launch_first_fwd(reinterpret_cast<const short*>(x), flag, p->pk + L.conv1.pk_w, p->pk + L.conv1.pk_b, cfg.channels, c1,
                                    p->B, p->H, p->W, p->A(L.z1), L.z1.C, p->st));
          p->first_skip = flag;
          p->first_ids = reinterpret_cast<const short*>(x);
        } else {
        if (p->opt.use_tc && p->opt.structured && p->ids_off >= 0) {
          // chargrid input: scan for one-hot structure, then the id-gather conv; the dense kernel below skips itself
          short* ids = reinterpret_cast<short*>(p->misc + p->ids_off);
          int* flag = reinterpret_cast<int*>(p->misc + p->ids_off) + (((long)p->B * p->H * p->W + 1) / 2 + 8);
          count_launch(2);
          MSAU_TRY(launch_onehot_scan(x, x_layout == 0, cfg.channels, c1, p->B, p->H, p->W, ids, flag, p->st));
          MSAU_TRY(launch_first_fwd(ids, flag, p->pk + L.conv1.pk_w, p->pk + L.conv1.pk_b, cfg.channels, c1, p->B, p->H, p->W, p->A(L.z1),
                                    L.z1.C, p->st));
          skip = flag;
          p->first_ids = ids;
        }
        p->first_skip = skip;
        MSAU_TRY(conv_same(p, x, c1, c1, x_layout == 0, cfg.channels, nullptr, 0, 0, p->pk + L.conv1.pk_w, p->pk + L.conv1.pk_b,
                           p->A(L.z1), L.z1.C, L.conv1.coutp, p->H, p->W, 3, 1, 1, o, L.conv1.tc_w, -1, skip));
        }
      } else {
        const Tensor& src = l == 0 ? prev->logits : blk.down[l - 1].pooled;
        MSAU_TRY(layer_fwd(p, L.conv1, src, nullptr, L.z1, o));
      }
      count_launch(1);
      MSAU_TRY(launch_lrn_fwd(p->A(L.z1), p->A(L.y1), p->npix(L.z1), L.z1.C, p->opt.lrn_coop, p->st));
      MSAU_TRY(res_fwd(p, L.res, L.y1, L.a, L.rr));
      if (b > 0) {
        const Tensor& pd = (l == S - 1) ? prev->att : prev->down[l].cc;
        ConvOpt oc; oc.relu = true;
        MSAU_TRY(layer_fwd(p, L.coupl, pd, &L.rr, L.cc, oc));
      }
      if (l == S - 1) {
        if (blk.has_attn) {
          ConvOpt oa;
          MSAU_TRY(layer_fwd(p, blk.attn.fg, L.cc, nullptr, blk.fg, oa));
          MSAU_TRY(layer_fwd(p, blk.attn.h, L.cc, nullptr, blk.hh, oa));
          if (p->opt.use_tc && attn_tc_supported(blk.attn.C, blk.attn.d)) {
            count_launch(3);
            MSAU_TRY(launch_attn_tc_fwd(p->A(blk.fg), p->A(blk.hh), p->A(L.cc), p->B, L.cc.H * L.cc.W, blk.attn.C, blk.attn.d,
                                        p->A(blk.mrow), p->A(blk.att), p->misc + p->attn_scratch_off, p->st));
          } else {
            count_launch(2);
            MSAU_TRY(launch_attn_fwd(p->A(blk.fg), p->A(blk.hh), p->A(L.cc), p->B, L.cc.H * L.cc.W, blk.attn.C, blk.attn.d,
                                     p->A(blk.mrow), p->A(blk.zinv), p->A(blk.att), p->st));
          }
        }
      } else {
        count_launch(1);
        MSAU_TRY(launch_pool_fwd(p->A(L.cc), p->A(L.pooled), p->B, L.cc.H, L.cc.W, L.cc.C, p->st));
      }
    }
    // ---- up tower, model/model.py:224-259 ----
    for (int l = S - 2; l >= 0; --l) {
      UpLevel& U = blk.up[l];
      const Tensor& xin = (l == S - 2) ? blk.down[S - 1].cc : blk.up[l + 1].uc;
      MSAU_TRY(deconv_fwd(p, U.deconv, xin, U.d));
      ConvOpt o;
      MSAU_TRY(layer_fwd(p, U.conv1, blk.down[l].cc, &U.d, U.u, o));
      MSAU_TRY(res_fwd(p, U.res, U.u, U.a, U.ur));
      if (b > 0) {
        ConvOpt oc; oc.relu = true;
        MSAU_TRY(layer_fwd(p, U.coupl, prev->up[l].uc, &U.ur, U.uc, oc));
      }
    }
    ConvOpt oe;
    MSAU_TRY(layer_fwd(p, blk.end, blk.up[0].uc, nullptr, blk.logits, oe));
  }
  const long npp = (long)p->H * p->W;
  if (aux) {
    count_launch(1);
    MSAU_TRY(launch_head(p->A(p->blocks[NB - 2].logits), p->lp, cfg.n_class, p->B, npp, aux, nullptr, nullptr, p->st));
  }
  if (logits || probs || argmax) {
    count_launch(1);
    MSAU_TRY(launch_head(p->A(p->blocks[NB - 1].logits), p->lp, cfg.n_class, p->B, npp, logits, probs, argmax, p->st));
  }
  return MSAU_OK;
}

extern "C" int msau_loss_backward_ex(MsauPlan* p, const float* x, int x_layout, const void* labels, const void* labels_aux,
                                     int label_dtype, const MsauLossSpec* spec, float loss_scale, void* workspace, size_t workspace_bytes,
                                     float* loss, float* loss_main, int32_t* accuracy, float* grads, void* stream) {
  MSAU_CHECK_ARG(p && x && labels && loss && grads && spec, "loss_backward: null argument");
  MSAU_CHECK_ARG(label_dtype == 0 || label_dtype == 1, "loss_backward: label_dtype must be 0 (uint8) or 1 (int64)");
  MSAU_TRY(bind(p, workspace, workspace_bytes, 1, stream));
  const MsauConfig& cfg = p->cfg;
  const int S = cfg.scale_space_num, NB = cfg.num_blocks;
  p->gparams = grads;
  std::fill(p->written.begin(), p->written.end(), 0);
  MSAU_CUDA_TRY(cudaMemsetAsync(grads, 0, sizeof(float) * p->n_params, p->st));
  p->wst = p->st;
  if (p->opt.side_stream) {
    if (!p->side) {
      MSAU_CUDA_TRY(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
      MSAU_CUDA_TRY(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
      MSAU_CUDA_TRY(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
    }
    p->wst = p->side;
  }
  const long npp = (long)p->H * p->W;
  pdl_set(p->opt.pdl != 0, 0);
  struct PdlOff { ~PdlOff() { pdl_set(false, 0); } } pdl_off;
  {
    Block& last = p->blocks[NB - 1];
    Block& auxb = p->blocks[NB - 2];
    int* ints = reinterpret_cast<int*>(p->misc);
    float* partial = p->misc + round_up((int)loss_scratch_ints(p->B), 64);
    LossSpec ls;
    ls.mode = spec->mode; ls.w_main = spec->weight_main; ls.w_aux = spec->weight_aux;
    ls.class_weights = spec->h_class_weights;
    count_launch(3);
    MSAU_TRY(launch_loss(p->A(last.logits), p->A(auxb.logits), p->lp, cfg.n_class, labels, labels_aux ? labels_aux : labels, label_dtype,
                         p->B, npp, loss_scale, ls, p->G(last.logits), p->G(auxb.logits), ints, p->d_flags, partial, loss, loss_main, p->st));
    if (accuracy)
      MSAU_CUDA_TRY(cudaMemcpyAsync(accuracy, ints + 2L * p->B * 32, 2 * sizeof(int), cudaMemcpyDeviceToDevice, p->st));
    p->touch(last.logits); p->touch(auxb.logits);
  }
  for (int b = NB - 1; b >= 0; --b) {
    Block& blk = p->blocks[b];
    Block* prev = b > 0 ? &p->blocks[b - 1] : nullptr;
    // ---- 4x4 head ----
    {
      const Tensor& src = blk.up[0].uc;
      MSAU_TRY(layer_wgrad(p, blk.end, 1, p->A(src), src.C, 0, blk.end.c1p, false, p->G(blk.logits), p->lp, nullptr, 0, p->H, p->W));
      ConvOpt o;
      // up-tower outputs uc (post-ReLU) have two gradient writers, this head / the deconv below and the next block's coupling:
      // each masks its own contribution, so the relu_mask pass over G(uc) is not needed
      if (p->opt.fuse_mask) { o.omask = p->A(src); o.pom = src.C; }
      MSAU_TRY(layer_dgrad(p, blk.end, 1, p->G(blk.logits), p->lp, nullptr, 0, src, o));
    }
    // ---- up tower (forward ran l = S-2..0, so backward runs l = 0..S-2) ----
    for (int l = 0; l <= S - 2; ++l) {
      UpLevel& U = blk.up[l];
      const Tensor& xin = (l == S - 2) ? blk.down[S - 1].cc : blk.up[l + 1].uc;
      bool pre = (b == 0) && p->opt.fuse_mask;       // block 0: uc IS the residual block's output ur
      if (b > 0) MSAU_TRY(coupl_bwd(p, U.coupl, prev->up[l].uc, U.ur, U.uc, &pre, p->opt.fuse_mask, p->opt.fuse_mask));
      MSAU_TRY(res_bwd(p, U.res, U.u, U.a, U.ur, pre));
      p->touch(U.u);
      // conv1s on cat[dw[l], deconv]
      const Tensor& skip = blk.down[l].cc;
      MSAU_TRY(layer_wgrad(p, U.conv1, 1, p->A(skip), skip.C, 0, U.conv1.c1p, false, p->G(U.u), U.u.C, nullptr, 0, U.u.H, U.u.W));
      MSAU_TRY(layer_wgrad(p, U.conv1, 2, p->A(U.d), U.d.C, 0, U.conv1.c2p, false, p->G(U.u), U.u.C, nullptr, 0, U.u.H, U.u.W));
      ConvOpt o;
      MSAU_TRY(layer_dgrad(p, U.conv1, 1, p->G(U.u), U.u.C, nullptr, 0, skip, o));
      MSAU_TRY(layer_dgrad(p, U.conv1, 2, p->G(U.u), U.u.C, nullptr, 0, U.d, o));
      MSAU_TRY(deconv_bwd(p, U.deconv, xin, U.d, p->opt.fuse_mask && l < S - 2));   // xin = uc of level l+1 (l = S-2: the deepest cc)
    }
    // ---- down tower, deepest level first ----
    for (int l = S - 1; l >= 0; --l) {
      Level& L = blk.down[l];
      if (l == S - 1 && blk.has_attn) {
        if (p->written[blk.att.id]) {   // only the next block's coupling reads the attention output
          const AttnLayer& at = blk.attn;
          const int N = L.cc.H * L.cc.W;
          if (p->opt.use_tc && attn_tc_supported(at.C, at.d)) {
            count_launch(3);
            MSAU_TRY(launch_attn_tc_bwd(p->A(blk.fg), p->A(blk.hh), p->G(blk.att), p->A(blk.mrow), p->B, N, at.C, at.d, p->G(blk.fg),
                                        p->G(blk.hh), p->misc + p->attn_scratch_off, p->st));
          } else {
            count_launch(3);
            MSAU_TRY(launch_attn_bwd(p->A(blk.fg), p->A(blk.hh), p->G(blk.att), p->A(blk.mrow), p->A(blk.zinv), p->B, N, at.C, at.d,
                                     p->A(blk.dvec), p->G(blk.fg), p->G(blk.hh), p->st));
          }
          // residual path out = x + o
          count_launch(1);
          MSAU_TRY(launch_add(p->G(L.cc), p->G(blk.att), p->npix(L.cc) * L.cc.C, p->touch(L.cc), p->st));
          // h conv
          MSAU_TRY(layer_wgrad(p, at.h, 1, p->A(L.cc), L.cc.C, 0, at.h.c1p, false, p->G(blk.hh), blk.hh.C, nullptr, 0, L.cc.H, L.cc.W));
          ConvOpt o;
          MSAU_TRY(layer_dgrad(p, at.h, 1, p->G(blk.hh), blk.hh.C, nullptr, 0, L.cc, o));
          // f | g conv: two parameter tensors, column windows [0,d) and [d,2d) of the fused output
          MSAU_TRY(layer_wgrad(p, at.fg, 1, p->A(L.cc), L.cc.C, 0, at.fg.c1p, false, p->G(blk.fg), blk.fg.C, nullptr, 0, L.cc.H, L.cc.W,
                               at.fg.w_off, at.fg.b_off, 0, at.d, at.d));
          MSAU_TRY(layer_wgrad(p, at.fg, 1, p->A(L.cc), L.cc.C, 0, at.fg.c1p, false, p->G(blk.fg), blk.fg.C, nullptr, 0, L.cc.H, L.cc.W,
                               at.g_w_off, at.g_b_off, at.d, at.d, at.d));
          MSAU_TRY(layer_dgrad(p, at.fg, 1, p->G(blk.fg), blk.fg.C, nullptr, 0, L.cc, o));
        }
      }
      // the pool gradient is the last contribution to G(cc) (the skip connection and the next block's coupling came earlier):
      // it also applies the mask of the ReLU that produced cc, so the relu_mask pass of coupl_bwd / res_bwd is skipped
      const bool pool_masks = p->opt.fuse_mask && l < S - 1;
      if (l < S - 1) {
        count_launch(1);
        MSAU_TRY(launch_pool_bwd(p->A(L.cc), p->G(L.pooled), p->G(L.cc), p->B, L.cc.H, L.cc.W, L.cc.C, p->touch(L.cc), pool_masks, p->st));
      }
      bool pre = (b == 0) && pool_masks;        // block 0 has no coupling conv: cc IS the residual block's output
      if (b > 0) {
        const Tensor& pd = (l == S - 1) ? prev->att : prev->down[l].cc;
        MSAU_TRY(coupl_bwd(p, L.coupl, pd, L.rr, L.cc, &pre, pool_masks));
      }
      MSAU_TRY(res_bwd(p, L.res, L.y1, L.a, L.rr, pre));
      count_launch(1);
      MSAU_TRY(launch_lrn_bwd(p->A(L.z1), p->G(L.y1), p->G(L.z1), p->npix(L.z1), L.z1.C, p->opt.lrn_coop, p->st));
      if (b == 0 && l == 0) {
        const int c1 = pad4(cfg.channels);
        if (x_layout == 3) {
          count_launch(3);
          MSAU_TRY(wgrad_fork(p));
          MSAU_TRY(launch_table_first_wgrad(p->first_ids, p->ftable, p->ftable_rows, cfg.channels, L.conv1.cout, p->G(L.z1), L.z1.C, p->tabH,
                                            p->B, p->H, p->W, p->gparams + L.conv1.w_off, p->wst));
          MSAU_TRY(launch_colsum(p->G(L.z1), p->npix(L.z1), L.z1.C, L.conv1.cout, p->gparams + L.conv1.b_off, p->wst));
        } else if (p->first_skip) {
          count_launch(1);
          MSAU_TRY(wgrad_fork(p));
          MSAU_TRY(launch_first_wgrad(p->first_ids, p->first_skip, p->G(L.z1), L.z1.C, cfg.channels, L.conv1.cout, p->B, p->H, p->W,
                                      p->gparams + L.conv1.w_off, p->gparams + L.conv1.b_off, p->wst));
        }
        if (x_layout != 2 && x_layout != 3)
        MSAU_TRY(layer_wgrad(p, L.conv1, 1, x, c1, x_layout == 0, cfg.channels, false, p->G(L.z1), L.z1.C, nullptr, 0, p->H, p->W, -1, -1,
                             0, -1, -1, p->first_skip));
      } else {
        const Tensor& src = l == 0 ? prev->logits : blk.down[l - 1].pooled;
        MSAU_TRY(layer_wgrad(p, L.conv1, 1, p->A(src), src.C, 0, L.conv1.c1p, false, p->G(L.z1), L.z1.C, nullptr, 0, L.z1.H, L.z1.W));
        ConvOpt o;
        MSAU_TRY(layer_dgrad(p, L.conv1, 1, p->G(L.z1), L.z1.C, nullptr, 0, src, o));
      }
    }
  }
  if (p->wst != p->st) {     // join: the caller's stream continues only after the last weight-gradient kernel
    MSAU_CUDA_TRY(cudaEventRecord(p->ev_join, p->wst));
    MSAU_CUDA_TRY(cudaStreamWaitEvent(p->st, p->ev_join, 0));
  }
  return MSAU_OK;
}

extern "C" int msau_loss_backward(MsauPlan* p, const float* x, int x_layout, const void* labels, int label_dtype, float loss_scale,
                                  void* workspace, size_t workspace_bytes, float* loss, float* grads, void* stream) {
  MsauLossSpec spec;
  spec.mode = 0; spec.weight_main = 1.f; spec.weight_aux = 1.f; spec.h_class_weights = nullptr;
  return msau_loss_backward_ex(p, x, x_layout, labels, nullptr, label_dtype, &spec, loss_scale, workspace, workspace_bytes, loss, nullptr, nullptr,
                               grads, stream);
}

extern "C" int msau_plan_error_flags(MsauPlan* p, int* h_flags) {
  MSAU_CHECK_ARG(p && h_flags, "plan_error_flags: null argument");
  // on the stream of the plan's last call (a non-blocking stream does not synchronise with the NULL stream)
  MSAU_CUDA_TRY(cudaMemcpyAsync(h_flags, p->d_flags, sizeof(int), cudaMemcpyDeviceToHost, p->st));
  MSAU_CUDA_TRY(cudaStreamSynchronize(p->st));
  if (*h_flags) {
    MSAU_CUDA_TRY(cudaMemsetAsync(p->d_flags, 0, sizeof(int), p->st));
    MSAU_CUDA_TRY(cudaStreamSynchronize(p->st));
  }
  return MSAU_OK;
}

extern "C" int msau_clip_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n, int step, float lr,
                                   float beta1, float beta2, float eps, float max_norm, float* scratch, float* total_norm,
                                   void* stream) {
  MSAU_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && scratch, "clip_adam_step: null argument");
  count_launch(2);
  return launch_clip_adam(params, grads, exp_avg, exp_avg_sq, n, step, lr, beta1, beta2, eps, max_norm, scratch, total_norm,
                          reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int msau_optimizer_step(int kind, float* params, float* grads, float* state1, float* state2, long long n, int step,
                                   int32_t* step_dev, float lr, float beta1, float beta2, float eps, float weight_decay, float max_norm,
                                   float* scratch, float* total_norm, void* stream) {
  MSAU_CHECK_ARG(params && grads && scratch, "optimizer_step: null argument");
  MSAU_CHECK_ARG((kind == 1 || state1) && (kind == 2 || state2), "optimizer_step: missing optimiser state buffer");
  count_launch(2);
  return launch_optim(kind, params, grads, state1, state2, n, step, step_dev, lr, beta1, beta2, eps, weight_decay, max_norm, scratch,
                      total_norm, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int msau_onehot_argmax(const void* onehot, int dtype, int batch, int channels, long long npix_per_page, int channels_last,
                                  uint8_t* out, void* stream) {
  MSAU_CHECK_ARG(onehot && out && batch >= 1 && npix_per_page >= 1, "onehot_argmax: bad argument");
  count_launch(1);
  return launch_onehot_argmax(onehot, dtype, batch, channels, npix_per_page, channels_last, out, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int msau_confusion_counts(const uint8_t* pred, const void* labels, int label_dtype, long long n, int n_class,
                                     long long* confusion, void* stream) {
  MSAU_CHECK_ARG(pred && labels && confusion && n >= 0, "confusion_counts: bad argument");
  MSAU_CHECK_ARG(label_dtype == 0 || label_dtype == 1, "confusion_counts: label_dtype must be 0 (uint8) or 1 (int64)");
  count_launch(1);
  return launch_confusion(pred, labels, label_dtype, n, n_class, confusion, reinterpret_cast<cudaStream_t>(stream));
}

// SelfAttentionBlock as a stand-alone operator (model/layers/attention.py:152-162), tensor-core path
extern "C" size_t msau_attention_scratch_bytes(int batch, int n_pos, int channels) {
  return attn_tc_scratch_bytes(batch, n_pos, channels);
}

extern "C" int msau_attention_forward(const float* fg, const float* hh, const float* x, int batch, int n_pos, int channels, float* lse,
                                      float* out, void* scratch, size_t scratch_bytes, void* stream) {
  MSAU_CHECK_ARG(fg && hh && x && lse && out && scratch, "attention_forward: null argument");
  MSAU_CHECK_ARG(batch >= 1 && n_pos >= 1, "attention_forward: bad shape");
  MSAU_CHECK_ARG(attn_tc_supported(channels, channels / 8), "attention_forward: channels must be 32 or 64");
  if (scratch_bytes < attn_tc_scratch_bytes(batch, n_pos, channels)) {
    set_error("attention_forward: scratch too small");
    return MSAU_ERR_WORKSPACE;
  }
  count_launch(3);
  return launch_attn_tc_fwd(fg, hh, x, batch, n_pos, channels, channels / 8, lse, out, scratch, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int msau_attention_backward(const float* fg, const float* hh, const float* d_out, const float* lse, int batch, int n_pos,
                                       int channels, float* d_fg, float* d_hh, void* scratch, size_t scratch_bytes, void* stream) {
  MSAU_CHECK_ARG(fg && hh && d_out && lse && d_fg && d_hh && scratch, "attention_backward: null argument");
  MSAU_CHECK_ARG(batch >= 1 && n_pos >= 1, "attention_backward: bad shape");
  MSAU_CHECK_ARG(attn_tc_supported(channels, channels / 8), "attention_backward: channels must be 32 or 64");
  if (scratch_bytes < attn_tc_scratch_bytes(batch, n_pos, channels)) {
    set_error("attention_backward: scratch too small");
    return MSAU_ERR_WORKSPACE;
  }
  count_launch(3);
  return launch_attn_tc_bwd(fg, hh, d_out, lse, batch, n_pos, channels, channels / 8, d_fg, d_hh, scratch,
                            reinterpret_cast<cudaStream_t>(stream));
}

// ---- debugging aids (tests/test_model_gpu.py compares every internal activation / activation gradient
//      with the oracle's traced forward); not part of the reference-facing surface ----
extern "C" int msau_debug_c3_prof(unsigned long long* h_counters16) {
  MSAU_CHECK_ARG(h_counters16, "debug_c3_prof: null argument");
  return debug_c3_prof(h_counters16);
}

extern "C" int msau_debug_layout(const MsauPlan* p, long long* packed_floats, long long* act_floats, int* n_tensors) {
  MSAU_CHECK_ARG(p, "debug_layout: null plan");
  if (packed_floats) *packed_floats = p->packed_floats + ((p->tc_elems + 1) / 2 + 63) / 64 * 64;
  if (act_floats) *act_floats = p->act_floats;
  if (n_tensors) *n_tensors = p->n_tensors;
  return MSAU_OK;
}

extern "C" int msau_debug_tensor(const MsauPlan* p, int id, long long* off, int* Cc, int* Hh, int* Ww) {
  MSAU_CHECK_ARG(p && id >= 0 && id < p->n_tensors, "debug_tensor: bad id");
  const Tensor& t = p->all_tensors[id];
  if (off) *off = t.off;
  if (Cc) *Cc = t.C;
  if (Hh) *Hh = t.H;
  if (Ww) *Ww = t.W;
  return MSAU_OK;
}

extern "C" int msau_set_option(const char* name, int value) {
  MSAU_CHECK_ARG(name, "set_option: null name");
  std::lock_guard<std::mutex> lk(g_defaults_mu);
  return set_opt(g_defaults, name, value);
}

extern "C" int msau_plan_set_option(MsauPlan* p, const char* name, int value) {
  MSAU_CHECK_ARG(p && name, "plan_set_option: null argument");
  return set_opt(p->opt, name, value);
}
