// One C call for the whole bf16 step: K1+K2 fused -> K3 -> K4 on one stream.  The kernels are the
// ones behind ovdet_similarity_fused / ovdet_decode_filter / ovdet_nms_batched; this entry only
// removes the per-kernel host round trips of a binding layer (from Python a ctypes call costs
// ~20 us, more than each kernel at batch 1).  Replaces model/yolo_clip.py:173-214 followed by
// inference/detector.py:184-208 for every image of the batch.
#include "common.cuh"
#include <cstdlib>

extern "C" int ovdet_head_step(const ovdet_head_step_args* a, void* stream) {
  if (!a) return OVDET_ERR_INVALID_ARG;
  if (a->num_levels <= 0 || a->num_levels > 4) return OVDET_ERR_UNSUPPORTED_SHAPE;
  int64_t hw[4];
  int64_t anchors = 0;
  for (int l = 0; l < a->num_levels; ++l) {
    hw[l] = (int64_t)a->heights[l] * a->widths[l];
    anchors += hw[l];
  }
  int rc;
  if (a->text_fp16) {
    if (a->embed_dtype != OVDET_F32) return OVDET_ERR_UNSUPPORTED_SHAPE;
    rc = ovdet_similarity_fused_fp16(reinterpret_cast<const float* const*>(a->obj_embeds), hw, a->emb_stride_b,
                                     a->emb_stride_d, a->num_levels, a->batch, a->dim, a->text_op, a->classes,
                                     a->text_batched, a->alpha, a->beta, nullptr, OVDET_F32, a->classes, a->scores,
                                     a->class_ids, a->inv_norm, stream);
  } else {
    rc = ovdet_similarity_fused_ws(reinterpret_cast<const float* const*>(a->obj_embeds), hw, a->emb_stride_b,
                                   a->emb_stride_d, a->num_levels, a->batch, a->dim, a->text_op, a->classes,
                                   a->text_batched, a->alpha, a->beta, a->scores, a->class_ids, a->inv_norm,
                                   a->sim_workspace, a->sim_workspace_bytes, a->embed_dtype, stream);
  }
  if (rc != OVDET_OK) return rc;
  // K3 and K4 are launched with programmatic stream serialization: their CTAs may be scheduled while
  // the preceding kernel drains (K3 decodes the boxes beside the similarity kernel's tail and waits
  // for it only before it reads the scores; K4 waits at its first instruction).  OVDET_PDL=0 disables.
  static const int pdl = []() { const char* e = getenv("OVDET_PDL"); return e ? atoi(e) : 1; }();
  rc = ovdet_decode_launch_internal(pdl, a->box_dtype == OVDET_BF16, a->box_preds, a->heights, a->widths,
                                    a->strides, a->box_stride_b, a->num_levels, a->bins, a->batch, 1.0f, 1.0f,
                                    a->scores, a->conf, a->activation, a->boxes, a->scores_act, a->pass_mask,
                                    stream);
  if (rc != OVDET_OK) return rc;
  const float* nms_scores = (a->activation == OVDET_ACT_SIGMOID && a->scores_act) ? a->scores_act : a->scores;
  return ovdet_nms_launch_internal(pdl, 0, 0.f, a->boxes, nms_scores, a->class_ids, a->pass_mask, a->batch,
                                   anchors, a->scale, a->clip_wh, a->iou_thr, a->class_aware, a->topk, a->max_det,
                                   a->out_boxes, a->out_scores, a->out_classes, a->out_anchor, a->out_keep,
                                   a->out_count, a->out_candidates, a->workspace, a->workspace_bytes, stream);
}

// The vocabulary-parallel step behind one C call: sharded similarity with the in-kernel key exchange
// -> signal -> wait + unpack -> K3 -> K4.  `a` describes this rank's shard (a->classes local rows in
// a->text_op); a->scores / a->class_ids receive the merged full-vocabulary result.
extern "C" int ovdet_head_step_vp(const ovdet_head_step_args* a, int64_t class_offset,
                                  void* const* peer_buffers, int world, int rank, int32_t* status,
                                  int timeout_ms, void* stream) {
  if (!a || !peer_buffers) return OVDET_ERR_INVALID_ARG;
  if (a->num_levels <= 0 || a->num_levels > 4) return OVDET_ERR_UNSUPPORTED_SHAPE;
  if (rank < 0 || rank >= world || world > OVDET_MAX_PEERS) return OVDET_ERR_INVALID_ARG;
  int64_t hw[4];
  int64_t anchors = 0;
  for (int l = 0; l < a->num_levels; ++l) {
    hw[l] = (int64_t)a->heights[l] * a->widths[l];
    anchors += hw[l];
  }
  if (a->batch == 0) return ovdet_check_device();
  const int64_t rows = a->batch * anchors;
  int rc = ovdet_similarity_fused_vp(a->obj_embeds, hw, a->emb_stride_b, a->emb_stride_d, a->num_levels,
                                     a->batch, a->dim, a->text_op, a->classes, a->text_batched, a->alpha,
                                     a->beta, a->inv_norm, a->sim_workspace, a->sim_workspace_bytes,
                                     a->embed_dtype, class_offset, peer_buffers, world, rank, stream);
  if (rc != OVDET_OK) return rc;
  rc = ovdet_vp_signal(peer_buffers, world, rank, rows, stream);
  if (rc != OVDET_OK) return rc;
  rc = ovdet_vp_wait_unpack(peer_buffers[rank], world, rows, a->scores, a->class_ids, status, timeout_ms, stream);
  if (rc != OVDET_OK) return rc;
  static const int pdl = []() { const char* e = getenv("OVDET_PDL"); return e ? atoi(e) : 1; }();
  rc = ovdet_decode_launch_internal(pdl, a->box_dtype == OVDET_BF16, a->box_preds, a->heights, a->widths,
                                    a->strides, a->box_stride_b, a->num_levels, a->bins, a->batch, 1.0f, 1.0f,
                                    a->scores, a->conf, a->activation, a->boxes, a->scores_act, a->pass_mask,
                                    stream);
  if (rc != OVDET_OK) return rc;
  const float* nms_scores = (a->activation == OVDET_ACT_SIGMOID && a->scores_act) ? a->scores_act : a->scores;
  return ovdet_nms_launch_internal(pdl, 0, 0.f, a->boxes, nms_scores, a->class_ids, a->pass_mask, a->batch,
                                   anchors, a->scale, a->clip_wh, a->iou_thr, a->class_aware, a->topk, a->max_det,
                                   a->out_boxes, a->out_scores, a->out_classes, a->out_anchor, a->out_keep,
                                   a->out_count, a->out_candidates, a->workspace, a->workspace_bytes, stream);
}

extern "C" size_t ovdet_head_step_args_size(void) { return sizeof(ovdet_head_step_args); }
