/*
 * ssd_b200.h -- C ABI of the B200-native SSD anchor pipeline (libssd_b200.so).
 *
 * Drop-in boundary for the per-image anchor pipeline of georgymironov/single_shot_detection
 * (paths below are relative to that repository).  The reference has no FFI of its own: the
 * path is plain Python calling torch CPU ops.  Each entry point here replaces one of those
 * Python functions; the Python classes in single_shot_detection_b200/ keep the reference's
 * names and signatures and forward to these symbols (see INTEGRATION.md for the binding a
 * maintainer of the reference would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises and nothing allocates: scratch comes from a caller-provided workspace whose
 *     size is returned by the matching *_workspace_bytes() query;
 *   - every function returns an ssd_status; ssd_b200_last_error() gives the message of the
 *     last failure on the calling thread;
 *   - fp32 everywhere (the reference calls .float() on its inputs), int64 class ids where the
 *     reference passes a LongTensor, uint8 masks (torch.bool storage);
 *   - arithmetic that decides an index (IoU, thresholds, NMS overlap) is done with separately
 *     rounded IEEE fp32 operations in the reference's operation order (no FMA contraction), so
 *     matched indices and keep lists are bit exact on identical inputs.
 */
#ifndef SSD_B200_H_
#define SSD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSD_B200_ABI_VERSION 3

#if defined(__GNUC__)
#define SSD_API __attribute__((visibility("default")))
#else
#define SSD_API
#endif

typedef enum {
    SSD_OK = 0,
    SSD_ERR_INVALID_ARGUMENT = 1,   /* null pointer, negative size, bad enum            */
    SSD_ERR_MISALIGNED = 2,         /* a pointer violates the alignment stated below    */
    SSD_ERR_WORKSPACE = 3,          /* workspace smaller than *_workspace_bytes()        */
    SSD_ERR_UNSUPPORTED = 4,        /* shape outside the supported envelope (see docs)  */
    SSD_ERR_CUDA = 5,               /* a CUDA runtime call failed                        */
    SSD_ERR_NO_DEVICE = 6           /* no sm_100 device / kernel image not loadable      */
} ssd_status;

/* target row layout -- detection/target_assigner.py:7-14 */
#define SSD_TARGET_COLS 6
#define SSD_CLASS_COL 4
#define SSD_SCORE_COL 5
#define SSD_NEGATIVE_CLASS 0
#define SSD_IGNORE_CLASS (-1)
/* matcher sentinels -- detection/matcher.py:4-5 */
#define SSD_NOT_MATCHED (-2)
#define SSD_IGNORE (-1)

SSD_API int ssd_b200_abi_version(void);
SSD_API const char* ssd_b200_last_error(void);
/* kernels this library has launched so far in this process (bench.py reports it as gpu_launches) */
SSD_API unsigned long long ssd_b200_launch_count(void);
/* SSD_OK when the current device is compute capability 10.x and the sm_100a image loads. */
SSD_API int ssd_b200_device_check(void);
/* Diagnostics: when enabled, every kernel launch of this library is bracketed by CUDA events on
 * its stream (do not enable while capturing a CUDA graph).  The report synchronises the device and
 * writes "label:mean_us:count,..." into buf; returns the number of characters written. */
SSD_API void ssd_b200_timing_enable(int on);
SSD_API size_t ssd_b200_timing_report(char* buf, size_t capacity);
/* Diagnostics: device-side timeline.  `device_slots` points to ssd_b200_trace_slots() pairs of
 * uint64 in device memory (initialise every pair to {UINT64_MAX, 0}); every kernel of the library
 * then records min(start) / max(end) of %globaltimer in its slot -- also inside a replayed CUDA
 * graph.  Pass NULL to switch the trace off (the default). */
/* Resident CTAs per SM of the logit-streaming kernels (process-wide, read when a launch is issued or captured):
 * 0 = automatic (fill the shared memory: fastest for a step that runs alone); 1 leaves room for the kernels of
 * other steps when several step graphs are in flight. */
SSD_API int ssd_b200_set_stream_ctas_per_sm(int ctas);
/* Threads per (image, class) CTA of the NMS kernel (process-wide, read when a launch is issued or captured):
 * 0 = default (128: best aggregate throughput with several steps in flight), or 32 / 64 / 128 / 256 (256 is
 * ~1 us faster for a step that runs alone).  Every setting produces identical detections. */
SSD_API int ssd_b200_set_nms_threads(int threads);
/* Candidate selection of the post-processor: -1 = automatic (one thread-block cluster per image when the image's
 * logits fit the cluster's shared memory -- one launch instead of pass 1 / gates / pass 2), 0 = always the streaming
 * two-pass path, 1/2/4/8 = that cluster size when it fits.  Both paths produce identical detections. */
SSD_API int ssd_b200_set_fused_select(int mode);
SSD_API int ssd_b200_trace_enable(unsigned long long* device_slots);
SSD_API int ssd_b200_trace_slots(void);

/* ------------------------------------------------------------------------------------------
 * a2  bf/utils/box_utils.py:83-101  iou(a, b) -- pairwise IoU of corner boxes.
 *     out[g*A + a] = inter / ((area_a[g] + area_b[a]) - inter), 0/0 -> NaN as in the reference.
 * ---------------------------------------------------------------------------------------- */
SSD_API int ssd_pairwise_iou(const float* a_corners, int num_a, const float* b_corners, int num_b,
                     float* out, void* stream);
/* f4  bf/utils/box_utils.py:104-143  generalized_iou(a, b, cartesian): iou - (enclosing - union) / enclosing.
 *     cartesian != 0: out [num_a, num_b]; else element-wise, num_a == num_b, out [num_a]. */
SSD_API int ssd_generalized_iou(const float* a_corners, int num_a, const float* b_corners, int num_b,
                        int cartesian, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * a3  detection/matcher.py:33-56  match_per_prediction(weights, matched, unmatched, force)
 *     weights[G, A] row-major -> box_idx[A] int64 in {-2, -1, 0..G-1}.  Thresholds are the
 *     fp32-rounded values (the reference compares in fp32).
 * ---------------------------------------------------------------------------------------- */
SSD_API int ssd_match_per_prediction(const float* weights, int num_gt, int num_anchors,
                             float matched_threshold, float unmatched_threshold, int force_match,
                             int64_t* box_idx_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * a1+a2+a3+a4  detection/target_assigner.py:22-63  TargetAssigner.encode_ground_truth
 *     One launch for the whole batch: to_corners(anchors) -> IoU -> per-anchor / per-GT argmax
 *     -> forced match -> target rows.
 *   anchors      [A,4]   (cx,cy,w,h) pixels, 16-byte aligned
 *   gt_rows      [sum G_i, gt_cols] rows (x1,y1,x2,y2,class,score,...), gt_cols >= 6
 *   gt_offsets   [B+1]   int32 row offsets into gt_rows (G_i = off[i+1]-off[i], 0 allowed)
 *   max_gt       max_i G_i (host knows it when it packs the list); <= 4096
 *   target_out   [B,A,6] fp32, 8-byte aligned
 *   match_out    [B,A]   int32 matcher output per anchor, or NULL
 *   stats_out    [B,4]   int32 {positives, ignored, positives with NaN box, G_i}, or NULL
 *   workspace    ssd_assign_workspace_bytes(batch, max_gt) bytes, 256-byte aligned (per-GT argmax
 *                table and per-image counters; zeroed inside the call)
 * ---------------------------------------------------------------------------------------- */
SSD_API size_t ssd_assign_workspace_bytes(int batch, int max_gt);
SSD_API int ssd_assign_targets(const float* anchors, const float* gt_rows, int gt_cols,
                       const int32_t* gt_offsets, int max_gt, int batch, int num_anchors,
                       float matched_threshold, float unmatched_threshold, int force_match,
                       float* target_out, int32_t* match_out, int32_t* stats_out, void* workspace,
                       size_t workspace_bytes, void* stream);
/* The same launch with the box columns written ALREADY passed through the loss route's box coding:
 * box_utils.to_centroids(inplace=True) + BoxCoder.encode_box(inplace=True), detection/losses/
 * multibox_loss.py:81-82, applied to every row as the reference does (its only consumer of the target
 * tensor does exactly that right away).  Bit-identical to ssd_assign_targets followed by
 * ssd_box_transform(SSD_BOX_CENTROIDS_ENCODE_INPLACE); match_out is required.  The NaN statistic still
 * refers to the corner boxes (target_assigner.py:60-61). */
SSD_API int ssd_assign_targets_encoded(const float* anchors, const float* gt_rows, int gt_cols,
                               const int32_t* gt_offsets, int max_gt, int batch, int num_anchors,
                               float matched_threshold, float unmatched_threshold, int force_match,
                               float xy_scale, float wh_scale, float eps, float* target_out,
                               int32_t* match_out, int32_t* stats_out, void* workspace,
                               size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * a1/a6/a7  box format + coding, bf/utils/box_utils.py:16-36 and detection/box_coder.py:13-57.
 *     The in-place and out-of-place branches of the reference round differently; both orders
 *     are provided.  src may equal dst.  Row strides are in floats (4 for packed boxes, 6 when
 *     operating on the box columns of a target tensor view).  priors[A,4] are indexed by
 *     row % A and may be NULL for the two format conversions.
 * ---------------------------------------------------------------------------------------- */
typedef enum {
    SSD_BOX_TO_CORNERS = 0,               /* box_utils.py:23                                  */
    SSD_BOX_TO_CENTROIDS = 1,             /* box_utils.py:36   ((max+min)/2, max-min)         */
    SSD_BOX_TO_CENTROIDS_INPLACE = 2,     /* box_utils.py:33-34 (min+(max-min)/2)             */
    SSD_BOX_ENCODE = 3,                   /* box_coder.py:32-34 log((wh+eps)/p_wh)            */
    SSD_BOX_ENCODE_INPLACE = 4,           /* box_coder.py:22-29 log(wh/p_wh+eps)              */
    SSD_BOX_DECODE = 5,                   /* box_coder.py:55-57                               */
    SSD_BOX_DECODE_INPLACE = 6,           /* box_coder.py:47-52                               */
    SSD_BOX_CENTROIDS_ENCODE_INPLACE = 7, /* 2 then 4 in one pass (multibox_loss.py:81-82)    */
    SSD_BOX_DECODE_TO_CORNERS = 8         /* 5 then 0 in one pass (postprocessor.py:52-53)    */
} ssd_box_op;

SSD_API int ssd_box_transform(int op, const float* src, int64_t src_row_stride, float* dst,
                      int64_t dst_row_stride, const float* priors, int64_t num_rows,
                      int num_anchors, float xy_scale, float wh_scale, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * a5  detection/sampler.py:9-10  naive_sampler: mask = class != 0 && class != -1
 * ---------------------------------------------------------------------------------------- */
SSD_API int ssd_positive_mask(const int64_t* target_classes, int64_t count, uint8_t* mask_out,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * a5  detection/sampler.py:12-25  hard_negative_mining
 *   logits          [B,A,C] fp32, 16-byte aligned
 *   target_classes  [B,A]   int64, 16-byte aligned
 *   loss_override   [B,A]   fp32 or NULL.  When given, it replaces -log_softmax(logits)[...,0]
 *                   (stage-boundary parity: selection on identical fp32 inputs) and logits may
 *                   be NULL.
 *   ratio           negative_per_positive_ratio; ratio_is_integer says whether the reference
 *                   would have computed n_pos*ratio in int64 (Python int) or fp32 (Python float)
 *   mask_out        [B,A]   uint8 (torch.bool)
 *   stats_out       [B,4]   int32 {positives, negatives, negatives selected, loss ties at cut}
 *                   or NULL.  Ties at the cut are broken towards the lower anchor index.
 * ---------------------------------------------------------------------------------------- */
SSD_API size_t ssd_hard_negative_workspace_bytes(int batch, int num_anchors);
/* First half only: the streamed mining criterion folded with the class id into one sortable
 * uint32 key per anchor (0 = ignored, 0xFFFFFFFF = positive, else ordered(-log_softmax[0])).  One launch,
 * no atomics, no memset (the selection builds its loss histogram itself); the workspace argument is kept
 * for ABI stability and not touched.  This is the HBM-bound kernel of the sampler; exported so it can be
 * timed / profiled alone. */
SSD_API int ssd_mining_keys(const float* logits, const int64_t* target_classes, int batch, int num_anchors,
                    int num_cols, uint32_t* keys_out, void* workspace, size_t workspace_bytes, void* stream);
SSD_API int ssd_hard_negative_mask(const float* logits, const int64_t* target_classes,
                           const float* loss_override, int batch, int num_anchors, int num_cols,
                           double ratio, int ratio_is_integer, double min_negatives,
                           uint8_t* mask_out, int32_t* stats_out, void* workspace,
                           size_t workspace_bytes, void* stream);

/* The selection alone, on the RAW criterion keys ssd_postprocess_pass1 emits (ordered_key(-log_softmax[0]),
 * class-agnostic) + the classes: int64 [B,A] (class_stride == 0) or the class column of fp32 target rows
 * (`classes` = &target[0][0][4], class_stride = 6 floats).  `match` (optional) is match_out of the ssd_assign_targets
 * call that wrote those rows: an unmatched / ignored anchor's class (0 / -1) is then taken from it (a coalesced read)
 * and only the few matched anchors read the strided class column.  No workspace. */
SSD_API int ssd_hard_negative_mask_from_keys(const uint32_t* loss_keys, const void* classes, int class_stride,
                                     const int32_t* match, int batch, int num_anchors, double ratio, int ratio_is_integer,
                                     double min_negatives, uint8_t* mask_out, int32_t* stats_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * a10 / f1  detection/losses/multibox_loss.py:56-92  masked multibox losses, forward + gradient.
 *     class_loss = classification_loss(scores[sampled_mask], classes[sampled_mask]) (reduction sum),
 *     loc_loss   = SmoothL1Loss(sum)(locs[positive], target_locs[positive]); both are scaled by
 *     their weight and divided by max(#positives, 1); loss = class_loss + loc_loss.
 *     kind SSD_LOSS_SOFTMAX_CE: torch CrossEntropyLoss(ignore_index = -1) over the C columns;
 *     kind SSD_LOSS_SIGMOID_FOCAL: bf/modules/losses.py:34-54 with the one-hot soft target
 *     (column class-1, value = ground-truth score) of multibox_loss.py:64-67; as in the reference
 *     this term is the MEAN over the sampled anchors (its constructor never receives
 *     reduction='sum': bf/utils/misc_utils.py:21-29 filters the keyword out), NaN for an empty mask.
 *   logits       [B,A,C]  fp32
 *   locs         [B,A,4]  fp32, 16-byte aligned
 *   target       [B,A,6]  fp32 rows (encoded box, class, score): AFTER to_centroids + encode_box
 *   sampled_mask [B,A]    uint8 (the sampler's output)
 *   grad_logits  [B,A,C]  d loss / d logits, or NULL      (dense: zero rows outside the mask)
 *   grad_locs    [B,A,4]  d loss / d locs, 16-byte aligned, or NULL
 *   loss_out     [3]      fp32 {loss, class_loss, loc_loss}
 * ---------------------------------------------------------------------------------------- */
#define SSD_LOSS_SOFTMAX_CE 0
#define SSD_LOSS_SIGMOID_FOCAL 1
SSD_API size_t ssd_multibox_loss_workspace_bytes(int batch, int num_anchors);
SSD_API int ssd_multibox_loss(const float* logits, const float* locs, const float* target,
                      const uint8_t* sampled_mask, int batch, int num_anchors, int num_cols, int kind,
                      float gamma, float alpha, float class_weight, float loc_weight,
                      float* grad_logits, float* grad_locs, float* loss_out, void* workspace,
                      size_t workspace_bytes, void* stream);
/* The same with bf/modules/losses.py:109-114 GeneralizedIoULoss as the localisation term
 * (multibox_loss.py:77-79): `target` rows hold the CORNER boxes (no coding), locs are decoded against
 * `priors` [A,4] (cx,cy,w,h) with BoxCoder.decode_box + to_corners, loc_loss = sum over the positives of
 * 1 - generalized_iou; grad_locs is d loss / d locs through GIoU, to_corners and the decoding. */
SSD_API int ssd_multibox_loss_giou(const float* logits, const float* locs, const float* target, const float* priors,
                           const uint8_t* sampled_mask, int batch, int num_anchors, int num_cols, int kind,
                           float gamma, float alpha, float class_weight, float loc_weight, float xy_scale,
                           float wh_scale, float* grad_logits, float* grad_locs, float* loss_out, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * a7+a8+a9  detection/postprocessor.py:24-78  Postprocessor.postprocess, including
 *           bf/utils/box_utils.py:165-194 nms (top-k + torchvision.ops.nms semantics).
 * ---------------------------------------------------------------------------------------- */
typedef enum { SSD_CONVERT_SOFTMAX = 0, SSD_CONVERT_SIGMOID = 1, SSD_CONVERT_IDENTITY = 2 } ssd_converter;
typedef enum { SSD_BOXES_ENCODED = 0, SSD_BOXES_CORNERS = 1 } ssd_box_input;

typedef struct {
    int32_t batch;               /* B                                                        */
    int32_t num_anchors;         /* A                                                        */
    int32_t num_cols;            /* C score columns per anchor                               */
    int32_t converter;           /* ssd_converter                                            */
    int32_t first_fg_col;        /* columns below it are skipped; SOFTMAX: 1, SIGMOID: 0     */
    int32_t box_input;           /* ssd_box_input: locs to decode, or ready corner boxes     */
    float xy_scale, wh_scale;    /* BoxCoder scales                                          */
    float score_threshold;       /* fp32-rounded; candidates need score > threshold          */
    int32_t max_per_class;       /* K, 1..512                                                */
    double overlap_threshold;    /* NMS IoU threshold, compared as (double)iou > threshold   */
    int32_t max_total;           /* T; <= 0 means no final top-k                             */
    int32_t det_capacity;        /* rows per image in dets_out; >= min(T or inf, Cf*K)       */
    int32_t soft_nms;            /* != 0: soft-NMS (box_utils.py:145-163) instead of hard NMS */
    float soft_sigma;            /* Gaussian decay exp(-iou^2 / sigma)                       */
    float soft_threshold;        /* threshold of the soft-NMS loop (Postprocessor: = score_threshold) */
    int32_t resume_after_pass1;  /* != 0: ssd_postprocess_pass1 already ran on this workspace    */
} ssd_postprocess_params;

SSD_API size_t ssd_postprocess_workspace_bytes(const ssd_postprocess_params* p);
/*   scores     [B,A,C]  fp32, 16-byte aligned
 *   boxes      [B,A,4]  fp32 locs (SSD_BOXES_ENCODED) or corner boxes (SSD_BOXES_CORNERS), 16-byte aligned
 *   priors     [A,4]    fp32 (cx,cy,w,h); may be NULL with SSD_BOXES_CORNERS
 *   dets_out   [B,det_capacity,6] fp32 rows (x1,y1,x2,y2,class,score); rows >= count are untouched
 *   count_out  [B]      int32 rows written per image
 *   anchor_out [B,det_capacity] int32 anchor index of each row, or NULL
 *   status_out [4]      int32 {errors (always 0), (image,class) lists that took the exact
 *                       column-rescan fallback, 0, 0} or NULL -- informational, results are exact
 * Row order per image: > T rows -> descending score (ties: class-major position);
 * otherwise class-major, descending score inside a class (ties: lower anchor first). */
SSD_API int ssd_postprocess(const ssd_postprocess_params* p, const float* scores, const float* boxes,
                    const float* priors, float* dets_out, int32_t* count_out, int32_t* anchor_out,
                    int32_t* status_out, void* workspace, size_t workspace_bytes, void* stream);

/* The first launch of ssd_postprocess on its own: row statistics (max, sum exp) and the gate bookkeeping
 * go into the workspace; with loss_keys_out != NULL (SOFTMAX only) the same streamed read of the logits
 * also emits the hard-negative-mining criterion of detection/sampler.py:13 per anchor as an order-
 * preserving uint32 key [B,A] (for ssd_hard_negative_mask_from_keys) -- in an eval step
 * (detection/init.py:117-122) sampler and post-processor see the same logits, so one read serves both.
 * Finish with ssd_postprocess(resume_after_pass1 = 1) on the same workspace. */
SSD_API int ssd_postprocess_pass1(const ssd_postprocess_params* p, const float* scores, uint32_t* loss_keys_out,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * a9  bf/utils/box_utils.py:165-194  nms(boxes, scores, overlap_threshold, score_threshold,
 *     max_per_class) for one box set (hard NMS).  keep_out[i] = input row of the i-th kept box
 *     (descending score).  keep_out holds min(n, max_per_class) entries (n when max_per_class<=0
 *     is not supported: pass max_per_class in 1..512).
 * ---------------------------------------------------------------------------------------- */
SSD_API size_t ssd_nms_workspace_bytes(int num_boxes, int max_per_class);
SSD_API int ssd_nms(const float* corner_boxes, const float* scores, int num_boxes, int max_per_class,
            double overlap_threshold, int64_t* keep_out, int32_t* count_out, void* workspace,
            size_t workspace_bytes, void* stream);

/* f4  bf/utils/box_utils.py:145-163  nms(..., soft=True): Gaussian soft-NMS of one box set, the
 *     reference's loop statement by statement (including its loop condition on the SUM of the
 *     remaining indices).  keep_out[i] = input row of the i-th PICKED box (pick order). */
SSD_API int ssd_soft_nms(const float* corner_boxes, const float* scores, int num_boxes, int max_per_class,
                 float score_threshold, float sigma, int64_t* keep_out, int32_t* count_out, void* workspace,
                 size_t workspace_bytes, void* stream);

/* box_utils.nms without a bound on the boxes that enter the NMS (bf/utils/box_utils.py:165-194 with
 * max_per_class=None or max_per_class > SSD_MAX_PER_CLASS): global-memory sort (score descending, index ascending),
 * the first max_keep rows (0 = all), then hard NMS with torchvision's semantics (64 x 64 tiles of the suppression
 * bit matrix + sweep; keep_out in descending score order) or, soft != 0, the reference's Gaussian soft-NMS loop
 * (box_utils.py:145-163; keep_out in pick order).  keep_out [min(max_keep or n, n)] int64 INPUT row indices,
 * count_out [1].  num_boxes <= 2^20; the bit matrix takes k * ceil(k / 64) * 8 workspace bytes.  A corner of the
 * API no sample configuration reaches: correct for any input, not tuned. */
#define SSD_MAX_PER_CLASS 512
SSD_API size_t ssd_nms_large_workspace_bytes(int num_boxes, int max_keep);
SSD_API int ssd_nms_large(const float* corner_boxes, const float* scores, int num_boxes, int max_keep,
                  double overlap_threshold, int soft, float soft_threshold, float soft_sigma, int64_t* keep_out,
                  int32_t* count_out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * f2  detection/metrics/mean_average_precision.py:10-116 + the accumulation loop of bf/eval.py:54-70
 *     (the consumer of the post-processor's output).
 *   ssd_map_append: compacts one batch of padded detections (dets [B, T, 6] + counts [B], as
 *     ssd_postprocess writes them) into rows [capacity, 7] = (image id, x1, y1, x2, y2, class, score),
 *     image id = image_base + b (eval.py:55-58), starting at *cursor_in; writes cursor_out[0] =
 *     *cursor_in + sum(counts) and cursor_out[1] = capacity (rows past the capacity are dropped: the
 *     caller compares the two).  cursor_in and cursor_out are two different int64[2] device words.
 *   ssd_mean_average_precision: preds [count, 7] in any order; ground truth as for
 *     ssd_assign_targets (a 7th "difficult" column is honoured when use_difficult);  classes are
 *     0 .. num_classes-1 (detections of any other class never enter the mean).
 *     ap_out [num_classes] fp32: AP per class, -1 for a class without (non-difficult) ground truth;
 *     map_out double[2] = (mAP, number of classes in the mean);  order_out [count] uint32 = the
 *     class-major / descending-score order the matching used;  flags_out [count] uint8 in that
 *     order: 1 = true positive, 2 = false positive, 0 = neither (matched a difficult box).
 *     voc != 0: 11-point interpolated AP, else area under the precision envelope.
 * ---------------------------------------------------------------------------------------- */
SSD_API size_t ssd_map_workspace_bytes(int64_t count, int total_gt, int num_classes);
SSD_API int ssd_map_append(const float* dets, const int32_t* counts, int batch, int max_total, int image_base,
                   float* rows, int64_t capacity, const int64_t* cursor_in, int64_t* cursor_out, void* stream);
SSD_API int ssd_mean_average_precision(const float* preds, int64_t count, const float* gt_rows, int gt_cols,
                               const int32_t* gt_offsets, int num_images, int total_gt, int num_classes,
                               float iou_threshold, int use_difficult, int voc, float* ap_out, double* map_out,
                               uint8_t* flags_out, uint32_t* order_out, void* workspace, size_t workspace_bytes,
                               void* stream);

/* ------------------------------------------------------------------------------------------
 * f3  detection/anchor_generators/ssd.py:106-151, retina_net.py:28-54, detection/detector.py:82-86
 *     the [A, 4] (cx, cy, w, h) anchor table of all levels, written on the device by one launch.
 *   levels: HOST array.  Per level the cell centres are torch.linspace(x_start, x_end, cells_x) /
 *   (y_start, y_end, cells_y) evaluated as ATen's fp32 CPU kernel does; wh holds the (w, h) of the
 *   num_boxes boxes of a cell.  Layout per level: rows, then columns, then boxes (the head layout).
 * ---------------------------------------------------------------------------------------- */
#define SSD_MAX_ANCHOR_LEVELS 8
#define SSD_MAX_BOXES_PER_CELL 16
typedef struct {
    int32_t cells_x, cells_y;
    float x_start, x_end, y_start, y_end;
    int32_t num_boxes;
    float wh[2 * SSD_MAX_BOXES_PER_CELL];
} SsdAnchorLevel;
SSD_API int ssd_generate_anchors(const SsdAnchorLevel* levels, int num_levels, float* anchors_out,
                         int64_t num_anchors, void* stream);

/* ------------------------------------------------------------------------------------------
 * e   the exchange step (no reference equivalent): pack the local images' padded
 *     detections [B,T,6], counts [B] and statistics into shard_out [capacity, ssd_shard_row_words(T)] fp32
 *     words: T*6 detection words, then int32 {count, positives, hard negatives selected, ignored,
 *     detections}, then zero padding to a 16-byte multiple; rows >= batch are padding with count = -1.  assign_stats = stats_out of
 *     ssd_assign_targets, mining_stats = stats_out of ssd_hard_negative_mask (either may be NULL).
 *     stats_out [B,4] int32 (optional) receives the four statistics.  The buffer is what ONE
 *     all-gather (NCCL over NVLink) moves.
 * ---------------------------------------------------------------------------------------- */
SSD_API int ssd_pack_shard(const float* dets, const int32_t* counts, const int32_t* assign_stats,
                   const int32_t* mining_stats, int batch, int max_total, int capacity, float* shard_out,
                   int32_t* stats_out, void* stream);

/* Words per packed row: T*6 detections + count + 4 statistics, rounded up to a multiple of four (16-byte rows). */
SSD_API int ssd_shard_row_words(int max_total);

/* The same packing FUSED with the exchange over NVLink peer memory: every CTA reads its image's row once and stores
 * it (16-byte stores) into slot `rank` of EVERY rank's gathered buffer, no NCCL call (csrc/exchange.cu has the
 * protocol).  `peer_arenas` is a HOST array of `world` device pointers: the arena of every rank as mapped into this
 * process (entry `rank` is the own arena).  An arena is ssd_exchange_arena_bytes(...) bytes of ZERO-filled device
 * memory (ssd_exchange_arena_alloc also returns its 64-byte CUDA IPC handle; a peer maps it with
 * ssd_exchange_peer_open).  Gathered slot `slot` of the own arena starts at ssd_exchange_slot_offset(...) and holds
 * [world, capacity, ssd_shard_row_words(T)] words once ssd_exchange_wait(slot) has completed on the stream.
 * One slot per step graph / stream: launches that use the same slot must be serialised, and every launch is
 *     ssd_exchange_open(slot)  ...the step...  ssd_pack_exchange(slot)  [ssd_exchange_wait(slot), readers]
 * -- the open call releases the slot's previous contents on every rank (a writer never overwrites rows a peer may
 * still be reading: it waits until every rank has opened the same launch).  A peer that stops answering turns into a
 * non-zero first header word (int64) after 10 s, never into a hang. */
#define SSD_EXCHANGE_MAX_WORLD 8
#define SSD_EXCHANGE_MAX_SLOTS 16
#define SSD_EXCHANGE_MAX_ROWS 512            /* images per rank and slot */
SSD_API int ssd_exchange_enable_peer(int device, int peer_device);   /* cudaDeviceEnablePeerAccess, idempotent */
SSD_API size_t ssd_exchange_arena_bytes(int world, int slots, int capacity, int max_total);
SSD_API size_t ssd_exchange_slot_offset(int world, int slot, int capacity, int max_total);
SSD_API int ssd_exchange_arena_alloc(size_t bytes, void** arena_out, void* ipc_handle_out64);
SSD_API int ssd_exchange_arena_free(void* arena);
SSD_API int ssd_exchange_peer_open(const void* ipc_handle64, void** mapped_out);
SSD_API int ssd_exchange_peer_close(void* mapped);
SSD_API int ssd_exchange_open(void* const* peer_arenas, int world, int rank, int slot, void* stream);
SSD_API int ssd_pack_exchange(const float* dets, const int32_t* counts, const int32_t* assign_stats,
                      const int32_t* mining_stats, int batch, int max_total, int capacity,
                      void* const* peer_arenas, int world, int rank, int slot, int32_t* stats_out, void* stream);
SSD_API int ssd_exchange_wait(void* own_arena, int world, int capacity, int slot, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSD_B200_H_ */
