/*
 * ovdet.h - C ABI of libovdet.so: the B200 (sm_100a) open-vocabulary detection head and
 * post-processing.  This is the drop-in boundary for the hot path of `yolo_clip_detector`.
 *
 * The reference is pure Python (no plugin / operator registry, no FFI); each entry point below
 * names the reference code it replaces (paths relative to /root/reference/yolo_clip_detector).
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is CALLER-OWNED DEVICE memory (e.g. torch.Tensor.data_ptr()) unless the
 *     comment says "host"; the library never allocates, frees or keeps device pointers;
 *   - every function enqueues on `stream` (a cudaStream_t passed as void*) and returns without
 *     synchronising; the caller has already selected the device;
 *   - return value: 0 on success, a negative ovdet_status otherwise; nothing throws or aborts;
 *   - there is NO CPU fallback: on a device that is not compute capability 10.x every compute
 *     entry returns OVDET_ERR_WRONG_ARCH;
 *   - strides and pitches are in ELEMENTS, sizes in elements unless named *_bytes.
 */
#ifndef OVDET_H_
#define OVDET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OVDET_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define OVDET_API __attribute__((visibility("default")))
#else
#define OVDET_API
#endif

typedef enum ovdet_status {
  OVDET_OK = 0,
  OVDET_ERR_INVALID_ARG = -1,       /* null pointer, negative size, bad enum */
  OVDET_ERR_UNSUPPORTED_SHAPE = -2, /* e.g. embed dim not a multiple of 64, > 8 levels */
  OVDET_ERR_WRONG_ARCH = -3,        /* device is not sm_100 */
  OVDET_ERR_CUDA = -4,              /* a CUDA call failed; see ovdet_last_cuda_error() */
  OVDET_ERR_WORKSPACE = -5,         /* workspace too small / misaligned */
  OVDET_ERR_DRIVER = -6             /* cuTensorMapEncodeTiled unavailable or failed */
} ovdet_status;

typedef enum ovdet_dtype { OVDET_F32 = 0, OVDET_BF16 = 1 } ovdet_dtype;
typedef enum ovdet_activation { OVDET_ACT_NONE = 0, OVDET_ACT_SIGMOID = 1 } ovdet_activation;

#define OVDET_MAX_LEVELS 8
#define OVDET_MAX_PEERS 8            /* GPUs of one NVSwitch box a vocabulary can be sharded over */

OVDET_API int ovdet_version(void);
OVDET_API const char* ovdet_strerror(int status);
/* cudaError_t of the most recent failing CUDA call on this thread (0 if none). */
OVDET_API int ovdet_last_cuda_error(void);
/* 0 when the current device can run the kernels (compute capability 10.x). */
OVDET_API int ovdet_check_device(void);

/* ------------------------------------------------------------------------------------------
 * K1a  L2 norm of the per-anchor region embeddings + tensor-core operand.
 * Replaces: model/heads/text_contrastive.py:134,137  (permute/reshape + F.normalize(p=2,
 *           eps=1e-12) of obj_embed), for one feature level.
 *
 *   x         fp32 [batch, dim, hw]   NCHW as obj_embed_conv emits it; element (b,d,a) at
 *             x[b*stride_b + d*stride_d + a]; the hw axis must be contiguous.
 *   operand   bf16 [batch, rows_per_batch, kop]; this level writes rows
 *             [row_offset, row_offset+hw) of every batch.  kop = dim (split=0) or 2*dim
 *             (split=1: columns [0,dim) hold hi = bf16(x), [dim,2dim) hold lo = bf16(x - hi)).
 *             Values are NOT pre-scaled: the 1/norm factor is applied by ovdet_similarity.
 *   inv_norm  fp32 [batch, rows_per_batch], same row window: 1 / max(||x||_2, 1e-12).
 * ---------------------------------------------------------------------------------------- */
OVDET_API int ovdet_l2norm_regions(const float* x, int64_t batch, int64_t dim, int64_t hw,
                         int64_t stride_b, int64_t stride_d,
                         void* operand, int64_t rows_per_batch, int64_t row_offset,
                         int64_t kop, int split, float* inv_norm, void* stream);

/* K1b  L2 norm of the text-prompt embeddings + tensor-core operand.
 * Replaces: model/heads/text_contrastive.py:138 (F.normalize of text_embed).
 *
 *   t         fp32 [batch, classes, dim]; element (b,c,d) at t[b*stride_b + c*stride_c + d].
 *             Pass batch=1 for a shared vocabulary (stride-0 expand, model/yolo_clip.py:123).
 *   operand   bf16 [batch, classes, kop], NORMALISED rows t/max(||t||,1e-12); split = 0: kop = dim;
 *             split = 1: [hi | lo], kop = 2*dim (ovdet_similarity, split recipe); split = 2:
 *             [hi | lo | hi], kop = 3*dim (the operand of ovdet_similarity_fused_fp32); split = 3:
 *             one FP16 segment holding 16 x the unit row, kop = dim (ovdet_similarity_fused_fp16).
 *   inv_norm  optional fp32 [batch, classes] (diagnostics), may be NULL.
 */
OVDET_API int ovdet_l2norm_text(const float* t, int64_t batch, int64_t classes, int64_t dim,
                      int64_t stride_b, int64_t stride_c,
                      void* operand, int64_t kop, int split, float* inv_norm, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2  region x text similarity on tcgen05 tensor cores (TMA-staged, TMEM accumulators).
 * Replaces: model/heads/text_contrastive.py:144,147 (matmul + alpha*s+beta) and, through the
 *           fused epilogue, model/yolo_clip.py:198-202 (max/argmax over classes).
 *
 *   regions_op   bf16 [batch, rows, kop]   from ovdet_l2norm_regions
 *   text_op      bf16 [text_batch, classes, kop] from ovdet_l2norm_text; text_batch is 1
 *                (shared) or batch (per-image text, as the neck produces, repvl_pan.py:173-182)
 *   inv_norm_r   fp32 [batch, rows] row scale; NULL = 1
 *   logits       optional [batch, rows, ldc] (fp32 or bf16 per logits_dtype), columns
 *                [0,classes) written: alpha * <a^,t^> + beta.  Memory order [B,HW,C] is what the
 *                reference's compute_similarity returns a transposed view of
 *                (text_contrastive.py:150-151).
 *   row_max      optional fp32 [batch, rows]  max over classes of the logits (before rounding
 *                to bf16 when logits_dtype is bf16)
 *   row_arg      optional int32 [batch, rows] argmax, lowest class index on ties
 *   split        0: one bf16 pass (operand error ~2^-9 relative, |dlogit| <~ 4e-3);
 *                1: three bf16 passes over hi/lo operands, fp32-class accuracy (~1e-6).
 *   dim must be a multiple of 64; kop = dim*(1+split).
 * ---------------------------------------------------------------------------------------- */
OVDET_API int ovdet_similarity(const void* regions_op, const void* text_op, const float* inv_norm_r,
                     int64_t batch, int64_t rows, int64_t classes, int64_t dim,
                     int split, int text_batched, float alpha, float beta,
                     void* logits, int logits_dtype, int64_t ldc,
                     float* row_max, int32_t* row_arg, void* stream);

/* K1+K2 fused  L2 norm of the region embeddings + similarity + class max/argmax in one kernel
 * that reads the fp32 NCHW conv outputs of ALL levels directly (no bf16 operand round trip
 * through HBM; the converted 128-anchor tile stays in tensor memory for the whole vocabulary).
 * Replaces: model/heads/text_contrastive.py:134-147 for every level + model/yolo_clip.py:198-206.
 *
 *   obj_embeds   HOST array of num_levels (<= 4) device pointers, level l is fp32
 *                [batch, dim, hw[l]] with element (b,d,a) at p[b*stride_b[l] + d*stride_d[l] + a];
 *                pointers 16-byte aligned, strides multiples of 4 elements (TMA), else
 *                OVDET_ERR_UNSUPPORTED_SHAPE (use ovdet_l2norm_regions + ovdet_similarity)
 *   text_op      bf16 [text_batch, classes, dim] from ovdet_l2norm_text (split = 0)
 *   dim          multiple of 64, <= 512;  one bf16 tensor-core pass (|dlogit| <~ 8e-3)
 *   logits       optional [batch, anchors, ldc], anchors = sum hw, levels concatenated in order
 *   row_max / row_arg   optional fp32 / int32 [batch, anchors]
 *   inv_norm     optional fp32 [batch, anchors] out: 1 / max(||x||_2, 1e-12)
 */
OVDET_API int ovdet_similarity_fused(const float* const* obj_embeds, const int64_t* hw,
                                     const int64_t* stride_b, const int64_t* stride_d,
                                     int num_levels, int64_t batch, int64_t dim,
                                     const void* text_op, int64_t classes, int text_batched,
                                     float alpha, float beta, void* logits, int logits_dtype,
                                     int64_t ldc, float* row_max, int32_t* row_arg,
                                     float* inv_norm, void* stream);

/* K1+K2 fused with a scratch buffer for SMALL launches (scores / argmax only).  When the launch has
 * fewer anchor tiles than half the CTA pairs of the GPU (batch 1 at 640^2: 67 tiles on 74 pairs), the
 * class tiles of every anchor tile are split over up to 4 work items; each writes its partial
 * (max, argmax) to the workspace and the last one to arrive - counted with an atomic per 32-row
 * group - merges them (lower class index wins ties, as in the unsplit kernel).  The workspace
 * (ovdet_similarity_split_workspace_bytes; 0 = this shape never splits) must be ZERO before its first
 * use and is left zero.  embed_dtype: OVDET_F32 or OVDET_BF16 activations. */
OVDET_API size_t ovdet_similarity_split_workspace_bytes(int64_t batch, int64_t anchors);
OVDET_API int ovdet_similarity_fused_ws(const float* const* obj_embeds, const int64_t* hw,
                                        const int64_t* stride_b, const int64_t* stride_d,
                                        int num_levels, int64_t batch, int64_t dim,
                                        const void* text_op, int64_t classes, int text_batched,
                                        float alpha, float beta, float* row_max, int32_t* row_arg,
                                        float* inv_norm, void* workspace, size_t workspace_bytes,
                                        int embed_dtype, void* stream);

/* K1+K2 fused, bf16 activations: as ovdet_similarity_fused, but obj_embeds[l] is bf16
 * [batch, dim, hw[l]] (the head convolutions ran under autocast; the reference's
 * compute_similarity accepts any float dtype).  The sum of squares is accumulated in fp32; the
 * tensor-core operand is the input itself, so no rounding is added.  Strides multiples of 8
 * elements, pointers 16-byte aligned. */
OVDET_API int ovdet_similarity_fused_bf16in(const void* const* obj_embeds, const int64_t* hw,
                                            const int64_t* stride_b, const int64_t* stride_d,
                                            int num_levels, int64_t batch, int64_t dim,
                                            const void* text_op, int64_t classes, int text_batched,
                                            float alpha, float beta, void* logits, int logits_dtype,
                                            int64_t ldc, float* row_max, int32_t* row_arg,
                                            float* inv_norm, void* stream);

/* K1+K2 fused, fp32-accurate: the same kernel with the three-pass recipe (x = hi + lo in bf16,
 * hi*hi + hi*lo + lo*hi accumulated in fp32, |dlogit| ~ 1e-5) for vocabularies of at most 128
 * classes (BASELINE configs[1]: 80 COCO prompts): with a single N tile every converted activation
 * block is consumed once, so the blocks STREAM through an 8-slot ring of tensor memory and dim = 512
 * fits.  Larger vocabularies: ovdet_l2norm_regions(split = 1) + ovdet_similarity(split = 1).
 *   text_op3   bf16 [text_batch, classes, 3 * dim]: unit-norm rows laid out [hi | lo | hi]
 *   other arguments as ovdet_similarity_fused; dim % 64 == 0, dim <= 512, classes <= 128
 *   (dim <= 128: any class count). */
OVDET_API int ovdet_similarity_fused_fp32(const float* const* obj_embeds, const int64_t* hw,
                                          const int64_t* stride_b, const int64_t* stride_d,
                                          int num_levels, int64_t batch, int64_t dim,
                                          const void* text_op3, int64_t classes, int text_batched,
                                          float alpha, float beta, void* logits, int logits_dtype,
                                          int64_t ldc, float* row_max, int32_t* row_arg,
                                          float* inv_norm, void* stream);

/* The same kernel with FP16 tensor-core operands instead of bf16 (the "fp16" precision tier): one
 * pass at the bf16 rate, 11-bit significands: |dlogit| ~ 1e-5 (max ~6e-5) at unit-norm scale against
 * ~4e-3 for bf16 and ~3e-5 for the three-pass fp32-accurate mode.  Every anchor row is scaled by a
 * power of two taken from its first 64 channels before the fp16 rounding (cosine similarity is
 * scale-invariant), so activations of any magnitude are accepted; a row whose later channels exceed
 * 8000 x the largest of its first 64 saturates at +-65504.  dim must be 512, fp32 activations.
 *   text_op16  fp16 [text_batch, classes, dim] from ovdet_l2norm_text(split = 3)
 * Other arguments as ovdet_similarity_fused. */
OVDET_API int ovdet_similarity_fused_fp16(const float* const* obj_embeds, const int64_t* hw,
                                          const int64_t* stride_b, const int64_t* stride_d,
                                          int num_levels, int64_t batch, int64_t dim,
                                          const void* text_op16, int64_t classes, int text_batched,
                                          float alpha, float beta, void* logits, int logits_dtype,
                                          int64_t ldc, float* row_max, int32_t* row_arg,
                                          float* inv_norm, void* stream);

/* K1+K2 projected ("next" row f-2)  the head's last layer - nn.Conv2d(hidden_dim, embed_dim, 1),
 * model/heads/text_contrastive.py:67 applied at :112 - folded into the similarity: the kernel
 * reads the HIDDEN features and never forms the embed_dim-wide embedding.
 * Replaces: text_contrastive.py:112 (1x1 conv) + :134-147 + model/yolo_clip.py:198-206, all levels.
 *
 *   hidden      HOST array of num_levels device pointers, fp32 [batch, hidden_dim, hw[l]] (same
 *               addressing / alignment rules as ovdet_similarity_fused); hidden_dim <= 448
 *   level_ops   HOST array of num_levels device pointers: level l's bf16 operand
 *               [text_batch, Cpad + kop, kop], kop = ceil(hidden_dim / 64) * 64 + 16,
 *               Cpad = classes rounded up to 128: rows [0, classes) = [W^T t_c | <b, t_c>] (t_c the
 *               unit-norm text row, W / b the level's 1x1 conv), rows [Cpad, Cpad + kop) =
 *               G' = [[W^T W, W^T b], [b^T W, b^T b]] (the constant of x' = [x, 1] sits at column
 *               ceil(hidden_dim / 64) * 64); built by the host layer (ops.project_vocabulary)
 *   row_max     fp32 [batch, anchors]: alpha * max_c cos(W x + b, t_c) + beta   (alpha >= 0)
 *   row_arg     optional int32 [batch, anchors]; inv_norm optional fp32: 1 / max(||W x + b||, 1e-12)
 * One bf16 tensor-core pass (|dscore| <~ 8e-3); no logits in this mode.
 */
OVDET_API int ovdet_similarity_projected(const float* const* hidden, const int64_t* hw,
                                         const int64_t* stride_b, const int64_t* stride_d,
                                         int num_levels, int64_t batch, int64_t hidden_dim,
                                         const void* const* level_ops, int64_t classes, int text_batched,
                                         float alpha, float beta, float* row_max, int32_t* row_arg,
                                         float* inv_norm, void* stream);

/* K2b  max/argmax over classes of materialised logits (any producer).
 * Replaces: model/yolo_clip.py:198-202 (similarity.max(dim=1)); ties -> lowest class index.
 *   logits [rows, ldc] fp32 or bf16, columns [0,classes) are read. */
OVDET_API int ovdet_rowmax(const void* logits, int logits_dtype, int64_t rows, int64_t classes, int64_t ldc,
                 float* row_max, int32_t* row_arg, void* stream);

/* E1  `obj_embeddings` of the forward dict: fp32 NCHW conv output of one level -> rows
 * [row_offset, row_offset + hw) of the anchor-major [batch, rows_per_batch, dim] array (levels
 * concatenated P3 | P4 | P5 by one call per level).
 * Replaces: model/yolo_clip.py:208-214 (permute(0,2,3,1).reshape(B,HW,D) per level + torch.cat).
 *   x   fp32 [batch, dim, hw], element (b,d,a) at x[b*stride_b + d*stride_d + a]
 *   out fp32 [batch, rows_per_batch, dim] contiguous */
OVDET_API int ovdet_concat_embeddings(const float* x, int64_t batch, int64_t dim, int64_t hw,
                                      int64_t stride_b, int64_t stride_d, float* out,
                                      int64_t rows_per_batch, int64_t row_offset, void* stream);

/* E2  Re-pitch the rows of a conv output (row r of `row_elems` elements at src + r*src_pitch -> dst +
 * r*dst_pitch, the tail [row_elems, dst_pitch) zero filled) so that the TMA loads of ovdet_similarity_fused*
 * can address a level whose H*W is not a multiple of 4 (13x13, 15x15, 19x19 ... at image sizes 416, 480, 608).
 * No reference counterpart: it stands in for the alignment torch's own kernels do not need
 * (model/heads/text_contrastive.py:134 permutes any H*W).  elem_size 4 (fp32) or 2 (bf16); pitches in elements. */
OVDET_API int ovdet_repitch_rows(const void* src, int64_t rows, int64_t row_elems, int64_t src_pitch,
                                 void* dst, int64_t dst_pitch, int elem_size, void* stream);

/* ------------------------------------------------------------------------------------------
 * K3  DFL box decode for all levels + score activation + confidence threshold.
 * Replaces: model/heads/box_head.py:150-218 (decode_boxes; grid of :115-148 is implicit),
 *           inference/detector.py:184 (scores > conf).
 *
 *   box_preds    HOST array of num_levels device pointers, level l is fp32
 *                [batch, 4*bins, heights[l], widths[l]] with batch stride batch_strides[l]
 *                (channel stride heights[l]*widths[l], spatial contiguous)
 *   heights/widths/strides/batch_strides  HOST arrays of num_levels entries
 *   scores       optional fp32 [batch, anchors] (anchors = sum h*w, levels concatenated in
 *                order, row-major inside a level - model/yolo_clip.py:205)
 *   boxes        fp32 [batch, anchors, 4] xyxy, centre/size decode of the reference:
 *                c = (cell + E[xy]) * stride, wh = exp(E[wh]) * stride * {width,height}_scale
 *   scores_act   optional fp32 [batch, anchors]: activation(scores) (needed for sigmoid)
 *   pass_mask    optional uint32 [batch, ceil(anchors/32)]: bit a%32 of word a/32 set iff
 *                activation(score[a]) > conf   (strict, NaN never passes)
 * ---------------------------------------------------------------------------------------- */
OVDET_API int ovdet_decode_filter(const float* const* box_preds, const int32_t* heights,
                        const int32_t* widths, const int32_t* strides,
                        const int64_t* batch_strides, int num_levels, int bins,
                        int64_t batch, float width_scale, float height_scale,
                        const float* scores, float conf, int activation,
                        float* boxes, float* scores_act, uint32_t* pass_mask, void* stream);

/* K3 for bf16 box logits (the head ran under autocast): same arguments, box_preds[l] is bf16
 * [batch, 4 * bins, h, w]; the softmax expectation and the decode run in fp32. */
OVDET_API int ovdet_decode_filter_bf16in(const void* const* box_preds, const int32_t* heights,
                        const int32_t* widths, const int32_t* strides,
                        const int64_t* batch_strides, int num_levels, int bins,
                        int64_t batch, float width_scale, float height_scale,
                        const float* scores, float conf, int activation,
                        float* boxes, float* scores_act, uint32_t* pass_mask, void* stream);

/* ------------------------------------------------------------------------------------------
 * K4  per-image candidate gather, rescale/clip, sort, greedy NMS.
 * Replaces: inference/detector.py:185-208 (mask, boxes/scale, clip, _nms :225-256,
 *           _compute_iou :258-287) for EVERY image of the batch (the reference handles image 0).
 *
 *   boxes [batch, anchors, 4], scores [batch, anchors], classes (optional int32 [batch,anchors])
 *   pass_mask  optional (NULL = every anchor is a candidate)
 *   scale      optional fp32 [batch]: boxes are divided by it (IEEE division); NULL = 1
 *   clip_wh    optional fp32 [batch, 2] = (orig_w, orig_h): x clipped to [0,w], y to [0,h]
 *   iou_thr    a candidate is dropped when NOT (iou <= iou_thr) against a kept one, with
 *              iou = inter / (area_a + area_b - inter + 1e-7f) in float32, no FMA contraction
 *   class_aware 0 = class-agnostic (reference); 1 = only same-class boxes suppress each other
 *   topk       0 = all candidates (reference); K = keep the K best (score desc, index desc)
 *              candidates before NMS
 *   max_det    capacity of the per-image output rows
 * Outputs (per image, in kept order = score desc, ties: higher anchor index first):
 *   out_boxes [batch,max_det,4] rescaled+clipped, out_scores, out_classes (optional),
 *   out_anchor int32 (anchor index), out_keep int32 (index into the thresholded array - what
 *   the reference's _nms returns), out_count int32 [batch] (kept, <= max_det),
 *   out_candidates optional int32 [batch] (number that passed the threshold)
 *   workspace: ovdet_nms_workspace_bytes(batch, anchors) bytes, 16-byte aligned.
 * ---------------------------------------------------------------------------------------- */
OVDET_API size_t ovdet_nms_workspace_bytes(int64_t batch, int64_t anchors);
OVDET_API int ovdet_nms_batched(const float* boxes, const float* scores, const int32_t* classes,
                      const uint32_t* pass_mask, int64_t batch, int64_t anchors,
                      const float* scale, const float* clip_wh,
                      float iou_thr, int class_aware, int topk, int64_t max_det,
                      float* out_boxes, float* out_scores, int32_t* out_classes,
                      int32_t* out_anchor, int32_t* out_keep, int32_t* out_count,
                      int32_t* out_candidates, void* workspace, size_t workspace_bytes,
                      void* stream);

/* K4 with the confidence threshold inside: candidates are the anchors with scores > conf (strict,
 * NaN never passes - inference/detector.py:184), evaluated by the kernel itself instead of read
 * from a pass mask.  The box decode then no longer depends on the scores and can run beside the
 * similarity kernel (HeadPipeline's captured graph forks it).  Same outputs as ovdet_nms_batched;
 * anchors <= 65536, else OVDET_ERR_UNSUPPORTED_SHAPE. */
OVDET_API int ovdet_nms_batched_conf(const float* boxes, const float* scores, const int32_t* classes,
                                     float conf, int64_t batch, int64_t anchors, const float* scale,
                                     const float* clip_wh, float iou_thr, int class_aware, int topk,
                                     int64_t max_det, float* out_boxes, float* out_scores,
                                     int32_t* out_classes, int32_t* out_anchor, int32_t* out_keep,
                                     int32_t* out_count, int32_t* out_candidates, void* workspace,
                                     size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * P1  letterbox pre-processing ("next" row f-4 of SURVEY.md section 8).
 * Replaces: inference/detector.py:139-156 - cv2.resize(image, (resized_w, resized_h)) with the
 *           default INTER_LINEAR, paste at the top-left of a zero canvas, astype(float32) / 255,
 *           HWC -> CHW.  Bit-exact with OpenCV's 8-bit fixed-point bilinear resize (including
 *           the 2x-decimation INTER_AREA fast path cv::resize switches to).
 *
 *   images      HOST array of `count` DEVICE pointers, image i is uint8 [heights[i], widths[i], 3]
 *               (RGB, interleaved) with rows row_strides[i] BYTES apart
 *   resized_h/w HOST arrays: size of the resized image, int(orig * scale) as the caller computed it
 *               (detector.py:141-142); must fit the canvas
 *   out         fp32 [count, 3, out_h, out_w]
 * ---------------------------------------------------------------------------------------- */
OVDET_API int ovdet_letterbox_u8(const uint8_t* const* images, const int32_t* heights,
                                 const int32_t* widths, const int64_t* row_strides,
                                 const int32_t* resized_h, const int32_t* resized_w, int count,
                                 int out_h, int out_w, float* out, void* stream);

/* P2  int-truncated boxes of the kept detections (detector.py:216: boxes[i].astype(int)).
 *   boxes fp32 [batch, max_det, 4] and count int32 [batch] as written by ovdet_nms_batched;
 *   out int32 [batch, max_det, 4], rows >= count[b] are zero.  Both 16-byte aligned. */
OVDET_API int ovdet_pack_boxes_i32(const float* boxes, const int32_t* count, int64_t batch,
                                   int64_t max_det, int32_t* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * N1  max-sigmoid text attention of the neck's text-guided CSP layer ("next" row f-1).
 * Replaces: model/repvl_pan.py:77-95 - permute, matmul(y, text'^T), max over classes, sigmoid,
 *           y * weight, permute back - in two launches (tcgen05 GEMM with the class max fused in
 *           its epilogue, then a streaming scale).
 *
 * ovdet_cast_text: fp32 text rows [batch, classes, dim] (element strides stride_b / stride_c,
 *   unit stride along dim) -> RAW (not normalised) bf16 operand [batch, classes, kop], zero padded
 *   to dpad = ceil(dim / 64) * 64 columns per segment; split3 = 0: kop = dpad, [hi];
 *   split3 = 1: kop = 3 * dpad, [hi | lo | hi] for the fp32-accurate three-pass product.
 * ovdet_max_sigmoid_attention:
 *   y        fp32 [batch, channels, hw], element (b, c, a) at y[b*stride_b + c*stride_c + a];
 *            hw and the strides multiples of 4, pointers 16-byte aligned, channels <= 512
 *            (<= 128 with precise), else OVDET_ERR_UNSUPPORTED_SHAPE
 *   text_op  from ovdet_cast_text(dim = channels, split3 = precise); text_batched: one table per
 *            image, else one shared table
 *   row_max  fp32 [batch, hw] out: max_c <y[:, a], t'_c>  (kept: it is also the attention logit)
 *   out      fp32, same indexing as y with out_stride_b / out_stride_c; may alias y
 * ---------------------------------------------------------------------------------------- */
OVDET_API int ovdet_cast_text(const float* t, int64_t batch, int64_t classes, int64_t dim,
                              int64_t stride_b, int64_t stride_c, void* operand, int64_t kop,
                              int split3, void* stream);
OVDET_API int ovdet_max_sigmoid_attention(const float* y, int64_t batch, int64_t channels, int64_t hw,
                                          int64_t stride_b, int64_t stride_c, const void* text_op,
                                          int64_t classes, int text_batched, int precise,
                                          float* row_max, float* out, int64_t out_stride_b,
                                          int64_t out_stride_c, void* stream);

/* ------------------------------------------------------------------------------------------
 * The whole bf16 step in ONE call: K1+K2 fused -> K3 -> K4 enqueued on `stream`.
 * Replaces: model/yolo_clip.py:173-214 followed by inference/detector.py:184-208, every image.
 * Same kernels, same argument meaning as ovdet_similarity_fused / ovdet_decode_filter /
 * ovdet_nms_batched (intermediates `scores`, `class_ids`, `boxes`, `pass_mask` are caller buffers
 * too: [batch, anchors], [batch, anchors], [batch, anchors, 4], [batch, ceil(anchors / 32)]).
 * A binding layer that pays per call (ctypes: ~20 us) launches the step for the price of one.
 * ---------------------------------------------------------------------------------------- */
typedef struct ovdet_head_step_args {
  int32_t num_levels;              /* <= 4 */
  int32_t bins;                    /* reg_max + 1 */
  int64_t batch, dim, classes;
  const void* obj_embeds[4];       /* fp32 or bf16 (embed_dtype) [batch, dim, h, w] per level */
  const void* box_preds[4];        /* fp32 or bf16 (box_dtype) [batch, 4 * bins, h, w] per level */
  int32_t heights[4], widths[4], strides[4];
  int64_t emb_stride_b[4], emb_stride_d[4], box_stride_b[4];
  const void* text_op;             /* bf16 [text_batch, classes, dim], unit-norm rows (fp16 x 16 with text_fp16) */
  int32_t text_batched;
  int32_t activation;              /* ovdet_activation */
  int32_t class_aware, topk;
  int32_t embed_dtype;             /* ovdet_dtype of obj_embeds: OVDET_F32 or OVDET_BF16 */
  int32_t box_dtype;               /* ovdet_dtype of box_preds */
  float alpha, beta, conf, iou_thr;
  int64_t max_det;
  float* scores;                   /* out, also K3/K4 input */
  int32_t* class_ids;
  float* inv_norm;                 /* optional */
  float* boxes;
  float* scores_act;               /* sigmoid activation only */
  uint32_t* pass_mask;
  const float* scale;              /* optional [batch] */
  const float* clip_wh;            /* optional [batch, 2] */
  float* out_boxes; float* out_scores; int32_t* out_classes; int32_t* out_anchor; int32_t* out_keep;
  int32_t* out_count; int32_t* out_candidates;
  void* workspace; size_t workspace_bytes;           /* K4: ovdet_nms_workspace_bytes */
  void* sim_workspace; size_t sim_workspace_bytes;   /* optional, ZEROED once by the caller:
                                                        ovdet_similarity_split_workspace_bytes */
  int32_t text_fp16;               /* 1: text_op is the fp16 operand of ovdet_l2norm_text(split = 3) and the
                                      similarity runs with fp16 tensor-core operands (ovdet_similarity_fused_fp16);
                                      fp32 activations, dim = 512 */
  int32_t reserved;
} ovdet_head_step_args;

OVDET_API int ovdet_head_step(const ovdet_head_step_args* args, void* stream);
/* sizeof(ovdet_head_step_args) as compiled into the library (binding layers check their mirror). */
OVDET_API size_t ovdet_head_step_args_size(void);

/* ------------------------------------------------------------------------------------------
 * Vocabulary-parallel similarity (SURVEY section 8 e "not built (optional later)" / f-4): the
 * C prompts are sharded over the GPUs of one NVLink box, every rank sees the whole image batch
 * and owns classes [class_offset, class_offset + classes).  The one exchange step of the path is
 * the reduction of the per-anchor (max, argmax) pairs of model/yolo_clip.py:198-206 over the
 * class shards.  It is fused into the similarity kernel: a finished row is packed into a 64-bit
 * key (score ascending, class descending: the maximum is the best score, lowest class index on
 * ties - torch.max's rule) and max-reduced with system-scope atomics straight into EVERY rank's
 * key array through NVLink peer mappings, from the epilogue warp that produced it.  No NCCL call
 * on the data path; ovdet_vp_signal / ovdet_vp_wait_unpack are the flag handshake that replaces
 * the collective's synchronisation.
 *
 * Exchange buffer (one per rank, ovdet_vp_buffer_bytes(rows, world) bytes, rows = batch * anchors):
 *   keys[2][rows] u64 (step parity p uses keys[p]; wait_unpack hands a row back as 0 after reading
 *   it, which is ordered before any peer's step + 2 atomics by the step + 1 handshake) followed by
 *   flags[world] u64 (flags[g] = last step rank g has finished contributing to).
 * Peer buffers: ovdet_peer_buffer_create allocates device memory and exports a 64-byte IPC handle;
 * the other ranks (one process per GPU) map it with ovdet_peer_buffer_open.  Within one process
 * (tests: several virtual ranks on one GPU) the pointers are used directly.
 * The step counters live in the buffers themselves (ctr[2] u64 behind the flags, moved by
 * ovdet_vp_signal), so no argument changes from step to step and the sequence
 * similarity_fused_vp -> vp_signal -> vp_wait_unpack can be captured in a CUDA graph; every rank
 * must issue the same number of steps.
 * ---------------------------------------------------------------------------------------- */
OVDET_API int ovdet_peer_buffer_create(size_t bytes, void** ptr, void* handle64);
OVDET_API int ovdet_peer_buffer_open(const void* handle64, void** ptr);
OVDET_API int ovdet_peer_buffer_close(void* ptr);
OVDET_API int ovdet_peer_buffer_destroy(void* ptr);
OVDET_API size_t ovdet_vp_buffer_bytes(int64_t rows, int world);
/* keys and flags to zero (once, before step 1; every rank, then a host-side barrier). */
OVDET_API int ovdet_vp_buffer_init(void* buffer, int64_t rows, int world, void* stream);
/* ovdet_similarity_fused_ws with the row results reduced into peer_buffers[0..world) instead of
 * row_max / row_arg.  text_op holds only this rank's `classes` rows. */
OVDET_API int ovdet_similarity_fused_vp(const void* const* obj_embeds, const int64_t* hw,
                                        const int64_t* stride_b, const int64_t* stride_d,
                                        int num_levels, int64_t batch, int64_t dim, const void* text_op,
                                        int64_t classes, int text_batched, float alpha, float beta,
                                        float* inv_norm, void* workspace, size_t workspace_bytes,
                                        int embed_dtype, int64_t class_offset,
                                        void* const* peer_buffers, int world, int rank, void* stream);
/* after the similarity kernel on the same stream: tell every rank that this rank's keys of the
 * current step have landed (system-scope release store of the step into flags[rank] of every
 * buffer) and advance this rank's counters. */
OVDET_API int ovdet_vp_signal(void* const* peer_buffers, int world, int rank, int64_t rows, void* stream);
/* wait until flags[g] >= step for every g, then keys[step & 1] -> scores fp32 / class_ids int32
 * (global class indices) and hand the rows back.  The wait is bounded (timeout_ms, 0 = 2000).  On
 * expiry NOTHING is unpacked or handed back: scores are set to -inf / class_ids to 0 (K3 / K4 then
 * yield no detections), a sticky word in the rank's own buffer makes every later call do the same
 * (a late peer's atomics may still be landing in the key arrays) until ovdet_vp_buffer_init re-arms
 * the buffer, and *status (int32, optional; device memory or mapped pinned host memory, which the
 * host can poll without synchronising) is set to 1.  Callers must check status. */
OVDET_API int ovdet_vp_wait_unpack(void* local_buffer, int world, int64_t rows,
                                   float* scores, int32_t* class_ids, int32_t* status, int timeout_ms,
                                   void* stream);
/* The vocabulary-parallel step in ONE call (ovdet_head_step for a class shard): similarity with the
 * in-kernel exchange -> signal -> wait + unpack -> K3 -> K4.  `args` describes this rank's shard
 * (args->classes rows in args->text_op); args->scores / args->class_ids receive the merged result. */
OVDET_API int ovdet_head_step_vp(const ovdet_head_step_args* args, int64_t class_offset,
                                 void* const* peer_buffers, int world, int rank, int32_t* status,
                                 int timeout_ms, void* stream);
/* The same reduction as a library collective (the baseline the fused exchange is measured against):
 * pack (score, class_offset + class) into int64 keys whose SIGNED order is the key order above, so
 * that an all-reduce(MAX) over int64 (NCCL, or gloo in the CPU tests) merges the shards; unpack. */
OVDET_API int ovdet_pack_score_keys(const float* scores, const int32_t* class_ids, int64_t n,
                                    int64_t class_offset, int64_t* keys, void* stream);
OVDET_API int ovdet_unpack_score_keys(const int64_t* keys, int64_t n, float* scores, int32_t* class_ids,
                                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OVDET_H_ */
