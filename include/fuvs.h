/*
 * fuvs.h — C ABI of libfuvs.so: the B200 (sm_100a) implementation of the
 * inter-frame segmentation interpolation path of
 * lenke182/flood-uav-video-segmentation.
 *
 * The reference has no FFI of its own: the seam is Python-method level
 * (SURVEY.md §8b).  Each entry point below names the reference call site(s)
 * it replaces, relative to the reference repository root.  The Python host
 * layer (flood_uav_video_segmentation_b200/) binds these with ctypes; a
 * maintainer's binding stub is shown in INTEGRATION.md.
 *
 * Conventions
 *  - Plain pointers and sizes only.  Every pointer is DEVICE memory unless the
 *    name ends in `_host`.  All float tensors are fp32, contiguous NCHW.
 *    Label maps are uint8 [n,H,W].  Count buffers are int64 [3,K] laid out as
 *    rows (area_intersection, area_union, area_target) and are ACCUMULATED
 *    in place (the caller zeroes them at epoch start).
 *  - The library allocates nothing persistent, borrows inputs for the
 *    duration of the enqueue and never synchronises the device: every call
 *    only enqueues kernels on `stream` (a cudaStream_t passed as void*).
 *  - Return value: 0 on success, a negative FUVS_E* code otherwise (never
 *    throws).  fuvs_last_error() returns a thread-local message for the last
 *    failure.  Asynchronous CUDA faults surface at the caller's next sync.
 *  - Arithmetic: every fp32 rounding step of the eager ATen op sequence the
 *    reference issues (grid_sampler_2d, upsample_bilinear2d, mul, add) is
 *    reproduced, so label maps and counts are bit-identical to the reference
 *    run on torch-CUDA (see DESIGN.md "Numerics").
 *  - No CPU fallback and no multi-backend dispatch: on a machine without an
 *    sm_100 device every compute entry returns FUVS_ENODEV.
 */
#ifndef FUVS_H_
#define FUVS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FUVS_ABI_VERSION 3   /* 3: + fuvs_dense_lowres_interval_ptrs, fuvs_comm_*, fuvs_allreduce_counts (additive) */

/* error codes */
#define FUVS_OK        0
#define FUVS_EINVAL   -1   /* bad shape / null pointer / unsupported size   */
#define FUVS_EALIGN   -2   /* pointer not aligned as documented             */
#define FUVS_ECUDA    -3   /* a CUDA runtime call failed (see last_error)   */
#define FUVS_ENODEV   -4   /* no CUDA device / not an sm_100 device         */

/* histogram conventions for fuvs_confusion (flags bit 0) */
#define FUVS_BINS_HISTC      0  /* torch.histc(bins=K,min=0,max=K-1): util/util.py:59-61 */
#define FUVS_BINS_NPHIST     1  /* np.histogram(bins=arange(K+1)), closed last bin: util/util.py:43-45 */
/* flags bit 1: write the ignore substitution back into `pred`
 * (intersectionAndUnionGPU mutates its caller's tensor, util/util.py:57) */
#define FUVS_MUTATE_PRED     2

#if defined(__GNUC__)
#define FUVS_API __attribute__((visibility("default")))
#else
#define FUVS_API
#endif

typedef void* fuvs_stream_t;   /* cudaStream_t */

FUVS_API int         fuvs_abi_version(void);
FUVS_API const char* fuvs_last_error(void);
/* Number of kernels this library has launched in the calling process. */
FUVS_API long long   fuvs_launch_count(void);
/* 0 if the current device can run the kernels (compute capability 10.x). */
FUVS_API int         fuvs_device_ok(void);

/* ---------------------------------------------------------------------------
 * Linear temporal blending (model.no_warp=True, model.feature_based=False).
 * Replaces, for one interval: flow/model.py:231-239 (fusion + cat) with
 * warp()==identity (flow/model.py:244-249), flow/base.py:276 (max(1)[1]),
 * flow/base.py:277 (uint8 cast) and the temporal-consistency metric loop
 * flow/base.py:280-295 + util/util.py:52-63.
 *
 *   prev,next : [C,H,W] key-frame logit maps
 *   n         : frames in the interval (= data.frame_delta); frame 0 is the
 *               unblended key frame, frame p = fl(fl(w0p*prev)+fl(w1p*next))
 *               with w0p=(float)((double)(n-p)/n), w1p=(float)((double)p/n)
 *   labels    : [n,H,W] uint8 out (arg-max over C; first max wins ties, NaN
 *               beats every number), or NULL
 *   logits    : [n,C,H,W] fp32 out, or NULL (FlowModel.predict()'s "pred")
 *   tc_prev   : [H,W] uint8 last label map of the previous interval
 *               (FlowBaseModel.last_output) or NULL for the first interval
 *   counts    : [3,K=C] int64 accumulate, or NULL: temporal-consistency
 *               (I,U,T) of label p against label p-1 (p=0 against tc_prev)
 *   ignore_index : as model.ignore_index (255)
 * ------------------------------------------------------------------------- */
FUVS_API int fuvs_linear_blend_argmax(const float* prev, const float* next,
                             int C, int H, int W, int n,
                             uint8_t* labels, float* logits,
                             const uint8_t* tc_prev, long long* counts,
                             int ignore_index, fuvs_stream_t stream);

/* ---------------------------------------------------------------------------
 * Linear temporal blending with the key frames given at DECODER resolution
 * (SURVEY.md §8f rank 1).  Replaces, in addition to what
 * fuvs_linear_blend_argmax replaces, the two F.interpolate(bilinear,
 * align_corners=True) calls that bring the decoder output to the frame size
 * (flow/model.py:191-193, 205-206): frame p = fl(fl(w0p*up(prev_lr)) +
 * fl(w1p*up(next_lr))), the up-sample evaluated on the fly with ATen's
 * arithmetic, so the 2*C*H*W*4 bytes of full-resolution logits are never
 * written.  prev_lr,next_lr : [C,hl,wl]; everything else as in
 * fuvs_linear_blend_argmax (labels [n,H,W], logits [n,C,H,W], ...).
 * hl==H && wl==W forwards to fuvs_linear_blend_argmax.  Supported when
 * fuvs_linear_lowres_supported(C,H,W) != 0 (2 <= C <= 5, W % 4 == 0);
 * otherwise FUVS_EINVAL: up-sample with fuvs_upsample_bilinear_ac first.
 * ------------------------------------------------------------------------- */
FUVS_API int fuvs_linear_lowres_supported(int C, int H, int W);
FUVS_API int fuvs_linear_lowres_blend_argmax(const float* prev_lr, const float* next_lr,
                             int C, int hl, int wl, int H, int W, int n,
                             uint8_t* labels, float* logits,
                             const uint8_t* tc_prev, long long* counts,
                             int ignore_index, fuvs_stream_t stream);

/* ---------------------------------------------------------------------------
 * One backward-warp step: F.grid_sample(src, grid, mode="bilinear",
 * padding_mode="border", align_corners=False|True) — flow/model.py:248
 * (align_corners=0) and flow/model.py:157 (align_corners=1).
 * Up to two independent (src,grid,dst) problems of the same shape are warped
 * in one launch (the forward and the backward chain advance in lock step).
 *
 *   src*  : [C,Hin,Win]   grid* : [Hg,Wg,2] (x,y) normalised   dst* : [C,Hg,Wg]
 *   src1/grid1/dst1 may all be NULL for a single-sided step.
 * ------------------------------------------------------------------------- */
FUVS_API int fuvs_warp_step(const float* src0, const float* grid0, float* dst0,
                   const float* src1, const float* grid1, float* dst1,
                   int C, int Hin, int Win, int Hg, int Wg, int align_corners,
                   fuvs_stream_t stream);

/* ---------------------------------------------------------------------------
 * Dense-flow interval (grids are [H,W,2]; the up-sample of flow/model.py:
 * 217-218,227-228 is skipped because shapes already match).  Replaces
 * flow/model.py:208-239 + flow/base.py:276-277 for one interval.
 *
 *   prev,next   : [C,H,W]
 *   grids_left  : [n-1,H,W,2]  mvs_left[0..n-2] stacked
 *   grids_right : [n-1,H,W,2]  mvs_right[0..n-2] stacked
 *   scratch     : fp32 workspace of fuvs_dense_scratch_floats(C,H,W,n) floats
 *                 (chain states 1..n-2 of both sides)
 *   labels      : [n,H,W] uint8 out or NULL;  logits : [n,C,H,W] out or NULL
 *   tc_prev/counts/ignore_index : as in fuvs_linear_blend_argmax
 * ------------------------------------------------------------------------- */
FUVS_API long long fuvs_dense_scratch_floats(int C, int H, int W, int n);
FUVS_API int fuvs_dense_interval(const float* prev, const float* next,
                        const float* grids_left, const float* grids_right,
                        int C, int H, int W, int n,
                        float* scratch,
                        uint8_t* labels, float* logits,
                        const uint8_t* tc_prev, long long* counts,
                        int ignore_index, fuvs_stream_t stream);
/* The same with the grids in the reference's own format: the python LISTS mvs_left / mvs_right of n-1 separate
 * [1,H,W,2] tensors (flow/dataset.py:138-146) arrive as HOST arrays of n-1 DEVICE pointers, read during the call;
 * nothing is stacked or copied.  (The layout of `scratch` is private to the library.) */
FUVS_API int fuvs_dense_interval_ptrs(const float* prev, const float* next,
                        const float* const* grids_left_host, const float* const* grids_right_host,
                        int C, int H, int W, int n,
                        float* scratch,
                        uint8_t* labels, float* logits,
                        const uint8_t* tc_prev, long long* counts,
                        int ignore_index, fuvs_stream_t stream);
/* Dense-flow interval with the key frames given at DECODER resolution: the
 * F.interpolate(decoder output, size=(h,w), mode="bilinear", align_corners=True) of flow/model.py:191-193, 205-206
 * runs inside the call and writes each up-sampled key frame straight into the layout the first warp step reads
 * (for C = 5 and odd n the strip kernel's channel-interleaved layout: its step 1 then gathers with a quarter of the
 * shared-memory loads), so the planar full-resolution key frame is never materialised.
 *   prev_lr, next_lr : [C,hl,wl] decoder outputs of the two key frames
 *   prev_up, next_up : caller-owned device buffers of C*H*W floats each (16-byte aligned), DIFFERENT buffers; they
 *                      receive the up-sampled key frames in a layout private to the library
 *   prev_up_ready    : != 0: prev_up already holds the up-sample of prev_lr, written as `next_up` by the previous
 *                      interval's call with the same C, H, W, n and the same outputs requested (key-frame reuse: the
 *                      interval's `next` is the following interval's `prev`, flow/model.py:189,202); 0: computed here
 *   everything else  : as in fuvs_dense_interval_ptrs (scratch: fuvs_dense_scratch_floats(C,H,W,n))
 * Results are bit-identical to fuvs_upsample_bilinear_ac on both key frames followed by fuvs_dense_interval. */
FUVS_API int fuvs_dense_lowres_interval_ptrs(const float* prev_lr, const float* next_lr, int hl, int wl,
                        float* prev_up, int prev_up_ready, float* next_up,
                        const float* const* grids_left_host, const float* const* grids_right_host,
                        int C, int H, int W, int n,
                        float* scratch,
                        uint8_t* labels, float* logits,
                        const uint8_t* tc_prev, long long* counts,
                        int ignore_index, fuvs_stream_t stream);

/* ---------------------------------------------------------------------------
 * Macro-block-grid interval (reference-faithful: grids are [Hg,Wg,2] with
 * Hg=H/16, Wg=W/16).  The chains run at grid resolution (flow/model.py:
 * 214-215,224-225), each state is bilinearly up-sampled to (H,W) with
 * align_corners=True (flow/model.py:217-218,227-228), blended
 * (flow/model.py:233-237) and arg-maxed (flow/base.py:276).
 *
 *   grids_left/right : [n-1,Hg,Wg,2]
 *   scratch          : fuvs_block_scratch_floats(C,Hg,Wg,n) floats
 *
 * Like every entry point it only enqueues work on `stream` and may be
 * captured into a CUDA graph; the chain is one cooperative launch when
 * enqueued eagerly and n-1 plain launches while `stream` is being captured
 * (same results; kernel-to-kernel latency inside a graph is below a grid-wide
 * barrier).
 * ------------------------------------------------------------------------- */
FUVS_API long long fuvs_block_scratch_floats(int C, int Hg, int Wg, int n);
FUVS_API int fuvs_block_interval(const float* prev, const float* next,
                        const float* grids_left, const float* grids_right,
                        int C, int H, int W, int Hg, int Wg, int n,
                        float* scratch,
                        uint8_t* labels, float* logits,
                        const uint8_t* tc_prev, long long* counts,
                        int ignore_index, fuvs_stream_t stream);
/* grids as host arrays of n-1 device pointers (the list format of flow/dataset.py:138-146), see
 * fuvs_dense_interval_ptrs */
FUVS_API int fuvs_block_interval_ptrs(const float* prev, const float* next,
                        const float* const* grids_left_host, const float* const* grids_right_host,
                        int C, int H, int W, int Hg, int Wg, int n,
                        float* scratch,
                        uint8_t* labels, float* logits,
                        const uint8_t* tc_prev, long long* counts,
                        int ignore_index, fuvs_stream_t stream);
/* The key frames at DECODER resolution [C,hl,wl]: prev = up(prev_lr), next = up(next_lr) with
 * up = F.interpolate(bilinear, align_corners=True) to (H,W) (flow/model.py:191-193, 205-206), evaluated inside the kernels
 * — the first chain step samples the up-sample at its 4 taps, frame 0 is the arg-max of the up-sample — so the two
 * full-resolution key frames (2*C*H*W*4 bytes) are never written or read.  Bit-identical to fuvs_upsample_bilinear_ac +
 * fuvs_block_interval.  Supported when fuvs_block_lowres_supported(...) != 0 (2 <= C <= 5, W % 4 == 0, at most four
 * key-frame rows under one grid row); hl==H && wl==W forwards to fuvs_block_interval_ptrs. */
FUVS_API int fuvs_block_lowres_supported(int C, int hl, int wl, int H, int W, int Hg, int Wg);
FUVS_API int fuvs_block_lowres_interval_ptrs(const float* prev_lr, const float* next_lr, int hl, int wl,
                        const float* const* grids_left_host, const float* const* grids_right_host,
                        int C, int H, int W, int Hg, int Wg, int n,
                        float* scratch,
                        uint8_t* labels, float* logits,
                        const uint8_t* tc_prev, long long* counts,
                        int ignore_index, fuvs_stream_t stream);
/* m CONSECUTIVE intervals of one clip in one call (what m successive predict_step calls compute, flow/base.py:259-295):
 * interval i lies between keys_host[i] and keys_host[i+1] (m+1 key-frame logit maps), uses the grids
 * grids_*_host[i*(n-1) .. i*(n-1)+n-2], writes labels_host[i] ([n,H,W]) and logits_host[i] (or NULL arrays), and its
 * frame 0 is compared with the last label map of interval i-1 (tc_prev for i = 0).  The chains of a clip's intervals
 * do not depend on each other, so each of the n-1 dependent chain steps is ONE launch for all m intervals: the launch
 * latencies that dominate a single block-grid interval are paid once per clip.
 *   scratch : m * fuvs_block_scratch_floats(C,Hg,Wg,n) floats;  all *_host arguments are host arrays of device pointers */
FUVS_API int fuvs_block_clip(int m, const float* const* keys_host,
                        const float* const* grids_left_host, const float* const* grids_right_host,
                        int C, int H, int W, int Hg, int Wg, int n,
                        float* scratch,
                        uint8_t* const* labels_host, float* const* logits_host,
                        const uint8_t* tc_prev, long long* counts,
                        int ignore_index, fuvs_stream_t stream);

/* ---------------------------------------------------------------------------
 * Feature-level interval (model.feature_based=True) — FlowModel.predict_feature,
 * flow/model.py:131-173: both warp chains over the encoder features, the
 * up-sample of every chain state back to the feature size (align_corners=
 * True), the temporal blend and the concatenation into the decoder batch.
 *   f_prev,f_next : [C,fh,fw] encoder features of the two key frames
 *   grids_*       : [n-1,Hg,Wg,2]
 *   default_grid  : [Hd,Wd,2] FlowModel.default_motion_vector or NULL; when
 *                   given, frame 0 = up(grid_sample(f_prev, default_grid,
 *                   align_corners=True)) exactly as flow/model.py:154-159,
 *                   else frame 0 = f_prev
 *   scratch       : fuvs_feature_scratch_floats(C,Hg,Wg,Hd,Wd,n) floats
 *   out           : [n,C,fh,fw] — the tensor the decoder is called on
 * ------------------------------------------------------------------------- */
FUVS_API long long fuvs_feature_scratch_floats(int C, int Hg, int Wg, int Hd, int Wd, int n);
FUVS_API int fuvs_feature_interval(const float* f_prev, const float* f_next,
                          const float* grids_left, const float* grids_right,
                          const float* default_grid, int Hd, int Wd,
                          int C, int fh, int fw, int Hg, int Wg, int n,
                          float* scratch, float* out, fuvs_stream_t stream);
/* grids as host arrays of n-1 device pointers, see fuvs_dense_interval_ptrs */
FUVS_API int fuvs_feature_interval_ptrs(const float* f_prev, const float* f_next,
                          const float* const* grids_left_host, const float* const* grids_right_host,
                          const float* default_grid, int Hd, int Wd,
                          int C, int fh, int fw, int Hg, int Wg, int n,
                          float* scratch, float* out, fuvs_stream_t stream);

/* ---------------------------------------------------------------------------
 * F.interpolate(src, size=(Hout,Wout), mode="bilinear", align_corners=True)
 * — flow/model.py:42,68,86,103,139,150,159,179,193,206,218,228 and
 * flow/base.py:219,232,275.   src [N*C,Hin,Win] -> dst [N*C,Hout,Wout].
 * ------------------------------------------------------------------------- */
FUVS_API int fuvs_upsample_bilinear_ac(const float* src, float* dst, long long planes,
                              int Hin, int Win, int Hout, int Wout,
                              fuvs_stream_t stream);

/* ---------------------------------------------------------------------------
 * out = fl(fl(wa*a) + fl(wb*b)) over `planes` [H*W] planes, optional arg-max
 * over the C planes of each of the `frames` images (planes = frames*C).
 * Replaces flow/model.py:104 (x * ((n-index)/n)) + flow/model.py:64,84 (sum
 * of the two sides) and flow/model.py:168,170,234-236 for one frame.
 * wa/wb are the python doubles; the library rounds them to fp32 as ATen does.
 * b may be NULL (out = fl(wa*a)).  out and labels may each be NULL.
 * ------------------------------------------------------------------------- */
FUVS_API int fuvs_blend_argmax(const float* a, const float* b, double wa, double wb,
                      int frames, int C, long long HW,
                      float* out, uint8_t* labels, fuvs_stream_t stream);

/* ---------------------------------------------------------------------------
 * labels = logits.max(1)[1]   — flow/base.py:147,167,276.
 *   logits [frames,C,HW] fp32 -> labels_u8 [frames,HW] and/or labels_i64.
 * ------------------------------------------------------------------------- */
FUVS_API int fuvs_argmax(const float* logits, int frames, int C, long long HW,
                uint8_t* labels_u8, long long* labels_i64,
                fuvs_stream_t stream);

/* ---------------------------------------------------------------------------
 * predict_step's final resize, fused with the arg-max — flow/base.py:275-277:
 *   output = F.interpolate(output, (1072, 1920), mode='bilinear', align_corners=True)
 *   output = output.data.max(1)[1]            (+ the uint8 cast of :277)
 * logits [frames,C,Hin,Win] -> labels [frames,Hout,Wout] uint8; the resized logits are only written when `resized`
 * ([frames,C,Hout,Wout]) is not NULL.  Hin==Hout && Win==Wout is the copy ATen makes, i.e. a plain arg-max.
 * ------------------------------------------------------------------------- */
FUVS_API int fuvs_upsample_argmax(const float* logits, int frames, int C, int Hin, int Win, int Hout, int Wout,
                         uint8_t* labels, float* resized, fuvs_stream_t stream);

/* ---------------------------------------------------------------------------
 * intersectionAndUnionGPU / intersectionAndUnion — util/util.py:52-63, 36-47.
 *   pred   : [N] labels, uint8 (pred_is_i64=0) or int64 (pred_is_i64=1)
 *   target : [N] labels, uint8 (target_is_i64=0) or int64 (target_is_i64=1)
 *   counts : [3,K] int64 accumulate (I,U,T)
 *   flags  : FUVS_BINS_HISTC | FUVS_BINS_NPHIST, optionally | FUVS_MUTATE_PRED
 * ------------------------------------------------------------------------- */
FUVS_API int fuvs_confusion(void* pred, int pred_is_i64,
                   const void* target, int target_is_i64,
                   long long N, int K, int ignore_index, int flags,
                   long long* counts, fuvs_stream_t stream);

/* ---------------------------------------------------------------------------
 * Temporal-consistency metric over finished label maps — flow/base.py:280-295:
 * for p in 0..n-1: metric(output=labels[p], target=labels[p-1]) with p=0
 * compared against tc_prev when it is not NULL (skipped otherwise).
 * A label counts iff it is < K (and, for the output, the target is not
 * ignore_index).  The reference's numpy metric (util/util.py:36-47) would
 * also count the value K itself as class K-1 (np.histogram closes its last
 * bin); arg-max label maps — all the reference ever passes here — never hold
 * it.  fuvs_confusion's FUVS_BINS_NPHIST reproduces that bin for label maps
 * that can.
 * ------------------------------------------------------------------------- */
FUVS_API int fuvs_temporal_counts(const uint8_t* labels, int n, long long HW,
                         const uint8_t* tc_prev, int K, int ignore_index,
                         long long* counts, fuvs_stream_t stream);

/* ---------------------------------------------------------------------------
 * The path's only collective: sum of the int64 count buffers over the ranks that shard the clips — one
 * process per GPU, one NCCL all-reduce per evaluation.  The reference has no counterpart (it averages per-rank mIoU
 * scalars, base/foundation.py:166-168); summing the integer (I,U,T) counts of util/util.py:36-63 reproduces the
 * single-process result exactly.  NCCL is resolved at run time from the libnccl.so.2 the host process has loaded.
 *   fuvs_comm_unique_id  : rank 0 fills 128 bytes (ncclUniqueId) that the host hands to every rank by its own means
 *   fuvs_comm_init       : every rank, on its own device: joins the communicator (collective call)
 *   fuvs_allreduce_counts: in-place sum of n int64 values (e.g. 3*K) on `stream`; counts is a device pointer
 *   fuvs_comm_world_size : ranks of the communicator, 0 without one;  fuvs_comm_destroy: releases it
 * ------------------------------------------------------------------------- */
FUVS_API int fuvs_comm_unique_id(void* id128);
FUVS_API int fuvs_comm_init(const void* id128, int rank, int world_size);
FUVS_API int fuvs_comm_world_size(void);
FUVS_API int fuvs_allreduce_counts(long long* counts, long long n, fuvs_stream_t stream);
FUVS_API int fuvs_comm_destroy(void);

/* ---------------------------------------------------------------------------
 * Sliding-crop inference (model.no_cropping=False) — flow/base.py:182-234.
 *
 * fuvs_crop_grid: crop_motion_vector for one grid (flow/transform.py:215-261,
 * called per crop at flow/base.py:204): the blocks covered by the crop are
 * sliced out, re-normalised to the crop's coordinate frame with numpy's
 * float32 arithmetic and resized to [crop_h/16, crop_w/16] like
 * cv2.resize(INTER_LINEAR).  The reference round-trips every grid through
 * host memory for this; here it stays on the device.
 *   grid : [Hg,Wg,2] fp32 (x,y) normalised     out : [crop_h/16, crop_w/16, 2]
 *   H,W  : size of the frame the grid belongs to; (h_off,w_off) crop origin
 *
 * fuvs_crop_accumulate: canvas[:, :, h_off:h_off+crop_h, w_off:w_off+crop_w]
 * += softmax(logits, dim=1); count[window] += 1   (flow/base.py:220,233 and
 * 206-207).  logits [n,C,crop_h,crop_w] fp32; canvas [n,C,H,W] fp64; count
 * [H,W] fp64, both zeroed by the caller before the first crop.
 *
 * fuvs_crop_finish: canvas /= count (flow/base.py:208, in place) and
 * labels[n,HW] = canvas.max(1)[1] (flow/base.py:167,276); labels may be NULL.
 * ------------------------------------------------------------------------- */
FUVS_API int fuvs_crop_grid_shape(int crop_h, int crop_w, int* out_h, int* out_w);
FUVS_API int fuvs_crop_grid(const float* grid, int Hg, int Wg, int H, int W,
                   int crop_h, int crop_w, int h_off, int w_off,
                   float* out, fuvs_stream_t stream);
FUVS_API int fuvs_crop_accumulate(const float* logits, double* canvas, double* count,
                         int n, int C, int crop_h, int crop_w, int H, int W,
                         int h_off, int w_off, fuvs_stream_t stream);
FUVS_API int fuvs_crop_finish(double* canvas, const double* count, int n, int C,
                     long long HW, uint8_t* labels, fuvs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FUVS_H_ */
