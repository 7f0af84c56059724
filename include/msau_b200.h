/* msau_b200 -- C ABI of the B200-native engine for the datvo06/MSAU hot path.
 *
 * The reference (pure Python / PyTorch / NumPy / SciPy) has NO native or FFI layer: its hot path sits
 * behind Python classes and functions.  This header is therefore the boundary a
 * maintainer binds with ctypes; every entry point names the reference symbol whose
 * arithmetic it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function returns 0 on success, a negative MSAU_ERR_* code otherwise; msau_last_error() gives
 *     a thread-local human-readable message;
 *   - every pointer is a DEVICE pointer owned by the caller unless the name starts with `h_`
 *     (host pointer); the library never frees or retains caller memory past the call (the only retained
 *     memory is the small descriptor table inside an MsauPlan);
 *   - every call takes the CUDA stream to enqueue on (`void*` = cudaStream_t) and is asynchronous on
 *     it: no hidden synchronisation, no hidden allocation on the data path;
 *   - activations are fp32; the network runs channels-last (NHWC) internally, the API tensors keep the
 *     reference's NCHW layout unless stated otherwise.
 */
#ifndef MSAU_B200_H
#define MSAU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSAU_OK 0
#define MSAU_ERR_ARG (-1)
#define MSAU_ERR_CUDA (-2)
#define MSAU_ERR_UNSUPPORTED (-3)
#define MSAU_ERR_WORKSPACE (-4)

const char* msau_last_error(void);
int msau_version(void);
/* number of kernels this library has launched in the calling process so far (bench.py `gpu_launches`) */
long long msau_launch_count(void);
/* a caller that replays a captured CUDA graph of these calls adds the graph's kernel count per replay (the library only sees the
 * launches it enqueues itself, i.e. the capture) */
void msau_launch_count_add(long long n);

/* ------------------------------------------------------------------------------------------------
 * Model: MSAUWrapper / MSAUNet            model/model.py:347-459 (kwargs :406-419)
 * ------------------------------------------------------------------------------------------------ */
typedef struct MsauConfig {
  int channels;         /* input channels (chargrid: n_token, BERT-grid: 768)          model.py:400 */
  int n_class;          /* logits channels                                              model.py:400 */
  int scale_space_num;  /* S: levels per U-Net block                                    model.py:406 */
  int res_depth;        /* R: convs per residual block                                  model.py:407 */
  int feat_root;        /* featRoot                                                     model.py:408 */
  int filter_size;      /* 3 (only value supported)                                     model.py:409 */
  int pool_size;        /* 2 (only value supported)                                     model.py:410 */
  int num_blocks;       /* 3, hard-coded in the reference                               model.py:354 */
} MsauConfig;

typedef struct MsauPlan MsauPlan;

/* Builds the launch plan for a fixed (batch, height, width).  Parameter order inside the flat
 * parameter / gradient buffers is the reference's state_dict() iteration order. */
int msau_plan_create(const MsauConfig* cfg, int batch, int height, int width, MsauPlan** plan);
void msau_plan_destroy(MsauPlan* plan);
long long msau_param_count(const MsauPlan* plan);
/* offset (in floats) and element count of the idx-th state_dict tensor; returns MSAU_ERR_ARG past the end */
int msau_param_info(const MsauPlan* plan, int idx, long long* offset, long long* numel);
/* Box-constant (BERT-grid) input, reference data_generator_funsd_bert.py:64-93: every pixel of a box carries that box's
 * feature vector.  `table` is a device fp32 [rows, channels] array (row r = the vector painted where the id map says r) that
 * must stay valid while msau_forward / msau_loss_backward run with x_layout 3 (x = int16 [B, H, W] row ids, -1 = background).
 * The 768 -> 8 first conv then runs as a rows x channels x 72 projection + an id gather; rows <= 32767.  Not a data-path
 * call: it may (re)allocate two small plan-owned scratch buffers. */
int msau_plan_set_feature_table(MsauPlan* plan, const float* table, int rows);
int msau_workspace_bytes(const MsauPlan* plan, int training, size_t* bytes);

/* MSAUWrapper.forward (model/model.py:435-437) = MSAUNet.forward (:378-396) + Softmax(dim=1).
 *   x            [B, channels, H, W] fp32 (x_layout 0, NCHW) or [B, H, W, round_up(channels,4)] (1, NHWC), or the int16 id map
 *                [B, H, W] of a one-hot chargrid (2: channel index per pixel, -1 = all-zero pixel; what
 *                msau_raster_features writes with layout 2), or the int16 row-id map of a box-constant grid (3, see
 *                msau_plan_set_feature_table)
 *   params       flat fp32 parameter buffer (msau_param_count floats)
 *   logits, aux  [B, n_class, H, W] fp32 NCHW, either may be NULL
 *   probs        [B, n_class, H, W] softmax over classes, or NULL
 *   argmax       [B, H, W] uint8 (first maximum wins; train...py:135, kv_model.py:162), or NULL
 * With training != 0 every activation needed by msau_loss_backward stays in `workspace`. */
int msau_forward(MsauPlan* plan, const float* x, int x_layout, const float* params, void* workspace,
                 size_t workspace_bytes, int training, float* logits, float* aux, float* probs,
                 uint8_t* argmax, void* stream);

/* MSAUWrapper.loss (model/model.py:446-459) averaged over the pages of the batch +
 * loss.backward() (train_chargrid_funsd_msau.py:56-57).  Must follow msau_forward(training=1) on the
 * same plan / workspace / x.
 *   labels       [B, H, W]; label_dtype 0 = uint8, 1 = int64; 0 = ignored pixel
 *   loss_scale   multiplies every gradient (1/world_size for data-parallel averaging); `loss` itself
 *                is reported unscaled
 *   loss         1 float (device)
 *   grads        flat fp32 gradient buffer (same layout as params); overwritten */
int msau_loss_backward(MsauPlan* plan, const float* x, int x_layout, const void* labels, int label_dtype,
                       float loss_scale, void* workspace, size_t workspace_bytes, float* loss,
                       float* grads, void* stream);

/* Generalised loss + backward: MSAUWrapper.loss (mode 0, above) or UNetLoss.forward of the alternative trainer
 * (mode 1; model/training/cost.py:35-65, called from model/training/trainer.py:124-137):
 *     loss = weight_main * CrossEntropyLoss(logits, tgt) + weight_aux * CrossEntropyLoss(aux_logits, aux_tgt)
 * over ALL pixels of the batch (class 0 included), the reference's 0.5 / 0.5 (cost.py:61), optional class weights with
 * torch's weighted-mean reduction (cost.py:27-31).  labels / labels_aux are the arg-max class maps of the reference's one-hot
 * targets (cost.py:41,52; msau_onehot_argmax); labels_aux NULL = labels.  n_class <= 32.
 *   loss_main    optional device float: the main head's own cross-entropy (UNetLoss's `final_loss`, cost.py:57-60)
 *   accuracy     optional device int32[2] = {pixels with label != 0 whose arg-max equals the label, pixels with label != 0}
 *                (cost.py:44-50; evaluate() of train...py:133-159 computes the same ratio)
 * A label outside [0, n_class) contributes nothing and raises bit 0 of the plan's sticky error flags (torch raises). */
typedef struct MsauLossSpec {
  int mode;                       /* 0 = MSAUWrapper.loss, 1 = UNetLoss */
  float weight_main, weight_aux;  /* mode 1 only (0.5 / 0.5 in the reference) */
  const float* h_class_weights;   /* mode 1 only: HOST pointer to n_class floats, or NULL */
} MsauLossSpec;
int msau_loss_backward_ex(MsauPlan* plan, const float* x, int x_layout, const void* labels, const void* labels_aux,
                          int label_dtype, const MsauLossSpec* spec, float loss_scale, void* workspace,
                          size_t workspace_bytes, float* loss, float* loss_main, int32_t* accuracy, float* grads,
                          void* stream);
/* reads and clears the plan's sticky device error flags (synchronises with the device; not a data-path call) */
int msau_plan_error_flags(MsauPlan* plan, int* h_flags);

/* nn.utils.clip_grad_norm(params, 1.0) + torch.optim.Adam.step()   train...py:24-26,58-59.
 * `scratch` >= 4 KiB; total_norm (device float, pre-clip norm) may be NULL. */
int msau_clip_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                        int step, float lr, float beta1, float beta2, float eps, float max_norm,
                        float* scratch, float* total_norm, void* stream);

/* get_optimizer (model/training/optimizer.py:4-30) on the flat buffers, one fused kernel pair:
 *   kind 0  torch.optim.Adam     state1 = exp_avg, state2 = exp_avg_sq, beta1 / beta2 / eps
 *   kind 1  torch.optim.RMSprop  state2 = square_avg, beta1 = alpha (0.99), eps (1e-8); momentum 0, not centred (:14-16, the default)
 *   kind 2  torch.optim.SGD      state1 = momentum buffer, beta1 = momentum (0.9), dampening 0 (:8-13)
 * weight_decay: g += weight_decay * p (the reference passes lr_decay_rate here, :7,12,16,21); max_norm > 0 applies
 * clip_grad_norm_ first (the alternative trainer does not clip: pass 0).  The step count (Adam's bias correction, SGD's first
 * step) comes from `step` (1-based, host) or, when step_dev is not NULL, from that device counter, which the call increments
 * first -- that form can be captured into a CUDA graph.  Trainer.adjust_lr (trainer.py:45-49) is the caller's `lr`. */
int msau_optimizer_step(int kind, float* params, float* grads, float* state1, float* state2, long long n, int step,
                        int32_t* step_dev, float lr, float beta1, float beta2, float eps, float weight_decay,
                        float max_norm, float* scratch, float* total_norm, void* stream);

/* torch.argmax(tgt, dim=1) on one-hot targets [B, C, H, W] (cost.py:41,52), or -- channels_last = 1 -- np.argmax(pred_mask,
 * axis=-1) on a [B, H, W, C] probability map (kv_model.py:162): dtype 0 = uint8, 1 = int64, 2 = float32; out uint8 [B, H, W],
 * first maximum wins. */
int msau_onehot_argmax(const void* onehot, int dtype, int batch, int channels, long long npix_per_page, int channels_last,
                       uint8_t* out, void* stream);
/* evaluate() of train_chargrid_funsd_msau.py:133-159 on the device: confusion[label][pred] += 1 over the pixels with
 * label != 0 (int64 [n_class][n_class], accumulated; accuracy = trace / sum, micro precision = micro recall = accuracy). */
int msau_confusion_counts(const uint8_t* pred, const void* labels, int label_dtype, long long n, int n_class,
                          long long* confusion, void* stream);

/* SelfAttentionBlock.forward (model/layers/attention.py:138-162) as a stand-alone operator on
 * channels-last tensors (the fused tcgen05 kernels the plan uses; the N x N map is never materialised).
 *   fg   [B, N, 2d]  f | g projections (d = channels / 8)      attention.py:152-153
 *   hh   [B, N, C]   h projection                               attention.py:154
 *   x    [B, N, C]   block input;  out = x + o                  attention.py:156-161
 *   lse  [B, N]      log2-domain log-sum-exp of every soft-max row, kept for the backward
 *   scratch          >= msau_attention_scratch_bytes, 128-byte aligned, contents need not survive
 * backward: d_out [B,N,C] -> d_fg [B,N,2d], d_hh [B,N,C] (the residual path d_x += d_out is the caller's).
 * channels must be 32 or 64. */
size_t msau_attention_scratch_bytes(int batch, int n_pos, int channels);
int msau_attention_forward(const float* fg, const float* hh, const float* x, int batch, int n_pos,
                           int channels, float* lse, float* out, void* scratch, size_t scratch_bytes,
                           void* stream);
int msau_attention_backward(const float* fg, const float* hh, const float* d_out, const float* lse,
                            int batch, int n_pos, int channels, float* d_fg, float* d_hh, void* scratch,
                            size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Rasterisers (CSR batches of pages; all coordinates fp64, arithmetic bit-exact with the reference)
 * ------------------------------------------------------------------------------------------------ */
/* Grid geometry of every page: get_min_max_x_y_w_h (data_generator_funsd_bert.py:49-61) + grid extent
 * (:72-73 / :154-155) + min_scale (:156-160).  page_ptr [n_pages+1] int32 CSR over boxes;
 * n_chars [n_boxes] int32 (characters per word; NULL for R2).  geom [n_pages][8] fp64 =
 * {min_x, min_y, min_w, min_h, min_scale, Hn, Wn, 0}. */
int msau_raster_geometry(const double* x, const double* y, const double* w, const double* h,
                         const int32_t* n_chars, const int32_t* page_ptr, int n_pages, double* geom,
                         void* stream);

/* R1 get_box_mask_box_label_word (data_generator_funsd_bert.py:149-186) and
 * R2 get_box_mask_box_label      (data_generator_funsd_bert.py:64-93).
 *   boxes          x,y,w,h fp64 [n_boxes]; page_ptr CSR [n_pages+1]
 *   char_ptr       [n_boxes+1] int32 CSR over characters, char_feat [n_chars_total] int32 = row of
 *                  feat_table per character (R1).  NULL for R2, where box i uses row feat_row[i].
 *   feat_row       [n_boxes] int32 (R2) or NULL
 *   feat_table     [n_rows, D] fp64 (charset one-hot rows / BERT embeddings); converted to fp32 by
 *                  round-to-nearest exactly like torch.Tensor(float64 ndarray)
 *   geom           from msau_raster_geometry of the WORD boxes (R1) / CELL boxes (R2)
 *   use_min_scale  1 = R1 word fill (x by min_scale), 0 = R2 box fill (x by min_w)
 *   out_h, out_w   allocated page extent; pages whose (Hn,Wn) differ are clipped / zero-filled
 *   grid           fp32, layout 0: [n_pages, D, out_h, out_w]  1: [n_pages, out_h, out_w, round_up(D,4)]
 *   owner_scratch  int32 [n_pages*out_h*out_w] */
int msau_raster_features(const double* x, const double* y, const double* w, const double* h,
                         const int32_t* page_ptr, int n_pages, int n_boxes, const int32_t* char_ptr,
                         const int32_t* char_feat, const int32_t* feat_row, const double* feat_table,
                         int feat_dim, const double* geom, int use_min_scale, int out_h, int out_w,
                         int layout, float* grid, int32_t* owner_scratch, void* stream);
/* label mask of R1/R2: label_mask[ny:ny+nh, nx:nx+nw] = label+1 (uint8), boxes scaled by min_w/min_h. */
int msau_raster_labels(const double* x, const double* y, const double* w, const double* h,
                       const int32_t* labels, const int32_t* page_ptr, int n_pages, int n_boxes,
                       const double* geom, int out_h, int out_w, uint8_t* label_mask, int32_t* owner_scratch, void* stream);

/* R3 KVModel._generate_masks_from_label (inference/kv_model.py:83-148).
 *   boxes [n_lines,4] fp64 x1,y1,x2,y2; char_ptr/char_ids CSR of token ids per line (after the host-side
 *   digit folding and tok_to_id lookup, :126,:140).  geom3 [n_pages][8] fp64 out =
 *   {min_x - bg_pad, min_y - bg_pad, scale, bg_pad, H, W, median_h, 0}; scaled_boxes [n_lines,4] int32.
 *   input/line/char masks uint16 [n_pages, out_h, out_w]. */
int msau_raster_kv_geometry(const double* boxes, const int32_t* page_ptr, int n_pages, double* geom3,
                            void* stream);
int msau_raster_kv(const double* boxes, const int32_t* page_ptr, int n_pages, int n_lines,
                   const int32_t* char_ptr, const int32_t* char_ids, const double* geom3, int out_h, int out_w,
                   uint16_t* input_mask, uint16_t* line_mask, uint16_t* char_mask, int32_t* scaled_boxes,
                   int32_t* owner_scratch, void* stream);
/* to_categorical + transposes (inference/generic_util.py:94-95, kv_model.py:274-278):
 * ids uint16 [n, H, W] -> one-hot fp32, layout 0: [n, n_token, H, W], 1: [n, H, W, round_up(n_token,4)] */
int msau_one_hot(const uint16_t* ids, int n_pages, int height, int width, int n_token, int layout,
                 float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Post-process: inference/morph_util.py:13-22,65-84 (SciPy ndimage semantics), kv_model.py:162-177
 * ------------------------------------------------------------------------------------------------ */
/* r_dilation / r_erosion: rectangular max / min filter, mode='constant' (zeros outside), SciPy origin. */
int msau_rect_filter(const uint8_t* in, uint8_t* out, int n_maps, int height, int width, int size_h,
                     int size_w, int origin_h, int origin_w, int is_max, void* stream);
/* r_closing(pred_class == cls, (1, size_w)) -- kv_model.py:175-176 -- in one pass over the uint8 class map: size_w in 1..4, width a
 * multiple of 16, 16-byte aligned maps; identical to msau_class_equals + msau_rect_filter(max) + msau_rect_filter(min) with origin 0. */
int msau_class_closing_row(const uint8_t* class_map, uint8_t* out, int n_maps, int height, int width, int cls, int size_w,
                           void* stream);
/* out = (class_map == cls) as uint8 0/1 (kv_model.py:175) */
int msau_class_equals(const uint8_t* class_map, uint8_t* out, long long n, int cls, void* stream);
/* connected_components: scipy.ndimage.label (4-connectivity, labels ordered by first raster pixel)
 * + find_objects.  labels int32 [n_maps,H,W]; n_labels int32 [n_maps]; bboxes int32
 * [n_maps, max_labels, 4] = y0,y1,x0,x1 half-open (labels beyond max_labels are counted, not boxed).
 * scratch int32 [n_maps*H*W + n_maps*H*ceil(W/32) + n_maps*(ceil(H*ceil(W/32)/32)+2)] (parent map | candidate-root bitmap |
 * chunk counts) */
int msau_ccl4(const uint8_t* binary, int n_maps, int height, int width, int32_t* labels, int32_t* n_labels,
              int32_t* bboxes, int max_labels, int32_t* scratch, void* stream);

/* Tail of KVModel._extract_value (inference/kv_model.py:178-261) on the device, after msau_ccl4 of every foreground class.
 * msau_kv_select_components: for every pixel whose component k (labels int32 [n_maps, H, W], one map per class) was picked by
 * the host (slot_of int32 [n_maps, max_labels + 1]: global slot id of component k of map m, or -1):
 *     presence[slot][line_mask[p]] = 1      -- np.unique(line_mask[labels == k])              kv_model.py:208,212
 *     new_mask[m][p] = 1 (else 0)           -- new_pred_mask[:, :, c][labels == k] = 1        kv_model.py:213,221
 * line_mask uint16 [H, W] (R3 line ids, 0 = none); presence uint8 [n_slots, n_lines + 1] (zeroed by the call);
 * new_mask uint8 [n_maps, H, W] (fully written).
 * msau_kv_char_range: queries int32 [n_queries, 5] = {map, x1, y1, x2, y2} (already clipped to the image); out int32
 * [n_queries, 2] = {min, max} of char_mask over the box pixels with new_mask[map] > 0 and char_mask > 0, {INT_MAX, 0} if there
 * are none                                                                                    kv_model.py:236-241 */
int msau_kv_select_components(const int32_t* labels, const uint16_t* line_mask, int n_maps, int height, int width,
                              const int32_t* slot_of, int max_labels, int n_slots, int n_lines, uint8_t* presence,
                              uint8_t* new_mask, void* stream);
int msau_kv_char_range(const uint16_t* char_mask, const uint8_t* new_mask, int height, int width, const int32_t* queries,
                       int n_queries, int32_t* out, void* stream);

/* Engine options.  "tensor_core_conv" (default 1): run the convolutions that fit on the tcgen05 implicit-GEMM
 * kernel; 0 = every convolution on the fp32 CUDA-core kernel (used by the parity tests to cross-check).  Options are
 * per plan: msau_set_option edits the defaults that msau_plan_create copies into a new plan (existing plans are not
 * touched), msau_plan_set_option edits one plan; there is no other mutable global state. */
int msau_set_option(const char* name, int value);
int msau_plan_set_option(MsauPlan* plan, const char* name, int value);

/* Per-kernel timing for bench.py's roofline report: when enabled every launch is bracketed by CUDA events on
 * its stream; msau_profile_report synchronises the device and writes a JSON object
 * {"<kernel>": {"launches", "ms", "flops", "bytes"}} (algorithmic work) into buf, then clears. */
int msau_profile_enable(int on);
int msau_profile_report(char* h_buf, size_t capacity);

/* Debugging aids used by the parity tests: workspace layout = [packed weights | activations |
 * activation gradients (training) | scratch], all in floats; tensor `id` (allocation order) lives at
 * activations + off as [B, H, W, C] fp32. */
/* conv3_tc role timers (env MSAU_TC_DEBUG=32): 16 cycle counters summed over the launches since the last call (conv3_tc.cu) */
int msau_debug_c3_prof(unsigned long long* h_counters16);
int msau_debug_layout(const MsauPlan* plan, long long* packed_floats, long long* act_floats, int* n_tensors);
int msau_debug_tensor(const MsauPlan* plan, int id, long long* off, int* channels, int* height, int* width);

#ifdef __cplusplus
}
#endif
#endif /* MSAU_B200_H */
