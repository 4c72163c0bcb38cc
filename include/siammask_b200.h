/* siammask_b200 — C ABI of the B200-native SiamMask per-frame inference hot path.
 *
 * The reference (foolwood/SiamMask) has no FFI for this path: the boundary is the duck-typed Python
 * object `state['net']` used by tools/test.py (siamese_init :155, siamese_track :201,203,257).  The entry
 * points below are what a binding for that object needs; each cites the reference method it replaces.
 * `siammask_b200/custom.py` is the ctypes binding that restores the Python API on top of them
 * (see INTEGRATION.md).
 *
 * Conventions: plain pointers and sizes only.  All *device* pointers are fp32 NCHW exactly as the
 * reference exchanges them (tools/test.py:61-64,205-206); the caller owns every I/O buffer, the engine
 * owns weights, workspace and per-slot caches.  Work is enqueued on `stream` (a cudaStream_t passed as
 * void*) and returns without synchronising.  Return value 0 = ok, negative = error with the message in
 * sm_last_error() (thread local).  Nothing here ever falls back to a CPU implementation: without a
 * CUDA device every compute entry point fails with an error.
 */
#ifndef SIAMMASK_B200_H
#define SIAMMASK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sm_engine sm_engine;

enum { SM_PRECISION_EXACT = 0, /* fp16 hi+lo operands, 3 tensor-core MMAs per k-step: fp32-class results */
       SM_PRECISION_FAST = 1   /* single fp16 MMA per k-step */ };
enum { SM_BACKEND_TENSOR = 0,  /* tcgen05 implicit-GEMM convolutions */
       SM_BACKEND_SIMT = 1     /* CUDA-core reference convolutions (debug / bisecting only) */ };
enum { SM_TRACK_MASK_FEATURES = 1, /* keep p0/p1/p2 + mask corr feature for sm_refine (track_mask) */
       SM_TRACK_MASK_HEAD = 2      /* also evaluate the 256->3969 mask head (dead under --refine) */ };

typedef struct sm_config {
  int32_t search_size;   /* hp.instance_size: 255 (config_davis.json) or e.g. 383 */
  int32_t max_batch;     /* largest B passed to sm_track / sm_template */
  int32_t num_slots;     /* tracker streams whose template kernels stay cached on the device */
  int32_t precision;     /* SM_PRECISION_* */
  int32_t backend;       /* SM_BACKEND_* */
  int32_t anchor_num;    /* len(ratios)*len(scales), models/siammask_sharp.py:17 (5) */
  int32_t with_mask;     /* build mask_model + refine_model (siammask_sharp) or RPN only (siamrpn_resnet) */
} sm_config;

/* One checkpoint tensor: reference state-dict name, host fp32 data, shape. */
typedef struct sm_tensor_desc {
  const char* name;
  const float* data;
  int32_t ndim;
  int64_t shape[4];
} sm_tensor_desc;

/* Custom.__init__ — experiments/siammask_sharp/custom.py:162-168 */
int sm_engine_create(const sm_config* cfg, sm_engine** out);
void sm_engine_destroy(sm_engine* e);

/* load_pretrain + model.load_state_dict — utils/load_helper.py:30-54.  Folds eval-mode BatchNorm
 * (eps 1e-5) into the convolutions, repacks to the kernels' layouts and uploads.  Copies; the caller
 * keeps ownership of `tensors`. */
int sm_engine_load_weights(sm_engine* e, const sm_tensor_desc* tensors, int32_t n);

/* The packed device weight arena (layout is a pure function of sm_config).  Multi-GPU init: rank 0
 * calls sm_engine_load_weights, every rank broadcasts [ptr, ptr+bytes) with NCCL, the other ranks
 * call sm_engine_adopt_weights. */
int sm_engine_weight_blob(sm_engine* e, void** dev_ptr, size_t* bytes);
int sm_engine_adopt_weights(sm_engine* e);

/* Static activation scales.  Activations live in HBM as two fp16 planes of value * 2^s (22 significant bits), so
 * |value * 2^s| must stay below 65504 and well above fp16's subnormals.  Without calibration s = 0 everywhere, which fits
 * BN-normalised checkpoints (activations O(1)..O(10^3)).  sm_engine_calibrate runs template + track_mask (+ mask head)
 * + refine on a representative sample batch (z f32 [B,3,127,127], x f32 [B,3,S,S], device; slots 0..B-1 are
 * overwritten), measures max |value| per tensor and re-packs every layer with per-tensor power-of-two scales chosen
 * for ~64x headroom: conv + BN is linear and ReLU / max-pool / crops commute with a positive scale, so the scales are
 * free at run time and results are unchanged.  Needs the weights to have come through sm_engine_load_weights on this
 * engine; the scales travel inside the weight arena (broadcast / packed file).
 * sm_engine_status: synchronises and returns flags; bit 0 = some activation left fp16's range since the last calibrate /
 * engine creation (results invalid: calibrate with representative data). */
int sm_engine_calibrate(sm_engine* e, int32_t B, const float* z_nchw, const float* x_nchw, void* stream);
int sm_engine_status(sm_engine* e, int32_t* flags);

/* Custom.template — custom.py:173-174.  z: device f32 [B,3,127,127].  Caches, for slots
 * slot0..slot0+B-1, the template feature and the three conv_kernel outputs (models/rpn.py:64),
 * which the reference recomputes every frame. */
int sm_template(sm_engine* e, int32_t slot0, int32_t B, const float* z_nchw, void* stream);

/* Custom.track / Custom.track_mask — custom.py:176-186.  x: device f32 [B,3,S,S] paired with slots
 * slot0..slot0+B-1.  cls: f32 [B,2A,R,R], loc: f32 [B,4A,R,R], mask: f32 [B,3969,R,R] or NULL.
 * flags: SM_TRACK_*. */
int sm_track(sm_engine* e, int32_t slot0, int32_t B, const float* x_nchw, float* cls, float* loc, float* mask,
             int32_t flags, void* stream);

/* Custom.track_refine — custom.py:188-190 -> Refine.forward(test=True) :131-154.  pos: device int32 [B,2]
 * (dy,dx) per stream; out: device f32 [B,127*127].  Uses the features cached by the preceding
 * sm_track(..., SM_TRACK_MASK_FEATURES) with the same B. */
int sm_refine(sm_engine* e, int32_t B, const int32_t* pos, float* out, void* stream);

/* get_subwindow_tracking — tools/test.py:67-110 — on the device: frames uint8 HWC (BGR as cv2.imread returns them),
 * frame b at frames + b*frame_stride (stride 0 = all streams share one frame); boxes int32 [B][8] (device) =
 * {context_xmin, context_ymin, original_sz, uint8(avg_chans[0..2]), 0, 0} in frame coordinates BEFORE padding (the
 * window may leave the frame; those pixels take the average colour, :89-100).  Resizes original_sz -> model_size
 * bit-exactly like cv2.resize(INTER_LINEAR) on 8-bit data and writes out f32 [B][3][model][model] (:61-64). */
int sm_crop_resize(const uint8_t* frames, size_t frame_stride, int32_t H, int32_t W, const int32_t* boxes, int32_t B,
                   int32_t model_size, float* out, void* stream);

/* Mask paste-back — crop_back() in siamese_track, tools/test.py:263-282: cv2.warpAffine(src f32 [B][src_h][src_w],
 * maps f64 [B][6] (forward 2x3 maps, device), (dst_w, dst_h), INTER_LINEAR, BORDER_CONSTANT, border_value), bit-exact
 * with OpenCV's fixed-point coordinate generation.  dst f32 [B][dst_h][dst_w].  All device pointers. */
int sm_warp_affine(const float* src, int32_t src_h, int32_t src_w, const double* maps, float* dst, int32_t dst_h,
                   int32_t dst_w, float border_value, int32_t B, void* stream);

/* Score / box post-processing + argmax of siamese_track — tools/test.py:205-254 — on the device, so that
 * sm_track -> sm_select -> sm_refine needs no host round trip.  All pointers are device pointers:
 * cls/loc as returned by sm_track; anchors f32 [A*R*R][4] = (cx,cy,w,h) in generate_anchor order (tools/test.py:113-129);
 * window f32 [A*R*R] (tiled hanning, :157-161); target_sz_in_crop f64 [B][2] = target_sz * scale_x (:226; float64 as
 * in the reference, whose penalty terms are evaluated in float64).
 * Outputs: best_idx int32 [B] (np.argmax of pscore, :237 — incl. its NaN rule: the first NaN wins), pos int32 [B][2] =
 * (delta_y, delta_x) (:253-254), records f32 [B][8] = decoded box cx,cy,w,h of the winner in crop units (:209-212),
 * score, penalty, pscore, best index (exact: < 2^24). */
int sm_select(sm_engine* e, int32_t B, const float* cls, const float* loc, const float* anchors, const float* window,
              const double* target_sz_in_crop, double penalty_k, double window_influence, int32_t* best_idx, int32_t* pos,
              float* records, void* stream);

/* Device-resident tracker state for B concurrent streams — the host arithmetic of siamese_track around the network
 * (tools/test.py:172-200, 239-249, 263-282, 305-315), float64 as numpy evaluates it, so that frame k+1's crop follows
 * from frame k's result without a host round trip.  state f64 [B][4] = target_pos (x, y), target_sz (w, h), device.
 *  sm_tracker_prepare: search window of the next frame -> boxes int32 [B][8] for sm_crop_resize (:180-198, :71-76),
 *    target_sz_in_crop f64 [B][2] for sm_select / sm_step (:226), aux f64 [B][4] = scale_x, round(s_x), crop_box x0, y0.
 *  sm_tracker_update: records f32 [B][8] of sm_select / sm_step + aux -> lr-smoothed, frame-clamped state (:239-249,
 *    :305-315; the penalty of the winner is re-evaluated in float64); maps f64 [B][6] (may be NULL) = forward affine map
 *    of crop_back (:263-275) for sm_warp_affine; out f64 [B][8] (may be NULL) = x, y, w, h, score, penalty, lr, best index.
 *    im_wh int32 [B][2] = frame width, height. */
typedef struct sm_tracker_hp {
  double context_amount, penalty_k, window_influence, lr;   /* utils/tracker_config.py:10-21 / config_davis.json */
  int32_t exemplar_size, instance_size, total_stride, base_size, out_size, reserved;
} sm_tracker_hp;
int sm_tracker_prepare(int32_t B, const double* state, const int32_t* avg_chans, const sm_tracker_hp* hp, int32_t* boxes,
                       double* target_sz_in_crop, double* aux, void* stream);
int sm_tracker_update(int32_t B, double* state, const float* records, const double* aux, const int32_t* im_wh,
                      const sm_tracker_hp* hp, int32_t anchor_num, int32_t score_size, double* maps, double* out,
                      void* stream);

/* One whole frame of siamese_track (tools/test.py:201-261) on the device, all pointers device pointers:
 * sm_track(flags) -> sm_select -> sm_refine at the position sm_select chose (refine_out != NULL needs
 * SM_TRACK_MASK_FEATURES) -> optionally mask_col f32 [B][3969] = mask[b, :, dy, dx] (:259-260; needs
 * SM_TRACK_MASK_HEAD and `mask`).  Unlike the three separate calls, the engine's two lanes run their halves of the
 * batch start to end without meeting in between.  refine_out / mask / mask_col may be NULL. */
int sm_step(sm_engine* e, int32_t slot0, int32_t B, const float* x_nchw, const double* target_sz_in_crop,
            const float* anchors, const float* window, double penalty_k, double window_influence, int32_t flags,
            float* cls, float* loc, float* mask, int32_t* best_idx, int32_t* pos, float* records, float* refine_out,
            float* mask_col, void* stream);

/* The same frame through HOST buffers: H2D of x and target_sz_in_crop, sm_step on staging buffers, D2H of the
 * records (always) and of whichever of refine / mask_col / cls / loc are non-NULL.  anchors / window stay device
 * pointers (per-tracker constants, tools/test.py:142-161).  Asynchronous with the ticket protocol of
 * sm_track_host_async (wait with sm_track_host_wait); with SM_TRACK_MASK_HEAD the raw 3969-channel head output stays on
 * the device (the reference reads one column of it, :259-260). */
typedef struct sm_step_io {
  const float* x_host;        /* f32 [B,3,S,S] */
  const double* tsz_host;     /* f64 [B,2] target_sz * scale_x */
  const float* anchors_dev;   /* f32 [A*R*R,4] device */
  const float* window_dev;    /* f32 [A*R*R] device */
  double penalty_k, window_influence;
  int32_t flags;              /* SM_TRACK_* */
  float* records_host;        /* f32 [B,8], required */
  float* refine_host;         /* f32 [B,127*127] or NULL */
  float* mask_col_host;       /* f32 [B,3969] or NULL */
  float* cls_host;            /* f32 [B,2A,R,R] or NULL */
  float* loc_host;            /* f32 [B,4A,R,R] or NULL */
} sm_step_io;
int sm_step_host_async(sm_engine* e, int32_t slot0, int32_t B, const sm_step_io* io, void* stream, int32_t* ticket);

/* Whole step through HOST buffers (pinned recommended): H2D of x, track(+mask features), optional refine,
 * D2H of cls / loc / refine logits, then stream synchronise.  mask_out_host may be NULL (no refine). */
int sm_track_host(sm_engine* e, int32_t slot0, int32_t B, const float* x_host, float* cls_host, float* loc_host,
                  const int32_t* pos_host, float* mask_out_host, void* stream);

/* Asynchronous form of sm_track_host: returns at once with a ticket (0/1); sm_track_host_wait(ticket) blocks
 * until that step's results are in the host buffers.  Two staging sets alternate, so submitting step k+1 before
 * waiting for step k overlaps its H2D (and step k's D2H) with compute.  Host buffers must stay valid (and pinned,
 * for real overlap) until the wait returns.  For B >= 16 the two halves of the batch run on two internal lanes
 * (own streams) that are ordered only by the ticket: the results are defined after sm_track_host_wait, not by
 * `stream` order; the next stream-ordered entry point (sm_template / sm_track / sm_refine / sm_export) joins the lanes
 * into its stream first. */
int sm_track_host_async(sm_engine* e, int32_t slot0, int32_t B, const float* x_host, float* cls_host, float* loc_host,
                        const int32_t* pos_host, float* mask_out_host, void* stream, int32_t* ticket);
int sm_track_host_wait(sm_engine* e, int32_t ticket);

/* conv2d_dw_group — models/rpn.py:32-38, standalone: x f32 [B,C,H,W], k f32 [B,C,kh,kw] ->
 * out f32 [B,C,H-kh+1,W-kw+1], all device pointers. */
int sm_xcorr_depthwise(const float* x, const float* k, float* out, int32_t B, int32_t C, int32_t H, int32_t W,
                       int32_t kh, int32_t kw, void* stream);

/* F.conv2d + folded affine (+ReLU) as a standalone operator, used by the kernel-level parity tests:
 * x f32 NCHW [B,Cin,H,W], w f32 [Cout,Cin,KH,KW], scale/shift f32 [Cout] (may be NULL), out f32 NCHW.
 * backend/precision as in sm_config. */
int sm_conv2d(const float* x, const float* w, const float* scale, const float* shift, float* out, int32_t B,
              int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t KH, int32_t KW, int32_t stride, int32_t pad,
              int32_t dil, int32_t relu, int32_t backend, int32_t precision, void* stream);

/* Copies a cached intermediate of the last sm_template / sm_track as f32 NCHW (parity checks):
 * "p0","p1","p2","p3","search","corr_cls","corr_loc","corr_mask","zf".  shape4 receives [B,C,H,W];
 * with out == NULL only the shape is returned. */
int sm_export(sm_engine* e, const char* what, float* out, int64_t* shape4, void* stream);

/* CUDA-graph replay of sm_track / sm_refine (off by default).  A call whose arguments (pointers, batch, flags,
 * stream) repeat is captured on its second occurrence and replayed afterwards: one graph launch instead of ~80
 * kernel launches — for the launch-bound small-batch / single-stream tracker loop. */
int sm_engine_set_graphs(sm_engine* e, int32_t on);

/* Per-launch CUDA-event timing on the caller's stream (bench.py's roofline leg).  While enabled every kernel
 * launch is bracketed by events; sm_profile_dump synchronises and returns tab-separated lines
 * "name\tcategory\tms\tflops\tbytes\n" (algorithmic FLOPs / bytes of that launch) and clears the log.
 * With buf == NULL or cap too small it returns the required size (call again). */
int sm_profile_enable(sm_engine* e, int32_t on);
int64_t sm_profile_dump(sm_engine* e, char* buf, size_t cap);

/* Number of kernel launches the engine has issued since creation (bench.py reports it). */
int64_t sm_launch_count(const sm_engine* e);
/* Device bytes held by the engine (weights + workspace + caches). */
size_t sm_engine_bytes(const sm_engine* e);

const char* sm_last_error(void);
const char* sm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SIAMMASK_B200_H */
