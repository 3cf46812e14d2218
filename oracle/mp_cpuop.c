/* TEST INFRASTRUCTURE ONLY.  Restatement of the reference's inference proposal operator MultiProposal:
 *   - variant CPU  (flags = 0): MultiProposalGPUOp<cpu>::Forward, SNIPER-mxnet/src/operator/multi_proposal.cc:150-374
 *     (plain greedy NMS over the 12000 best-scoring anchors).  This is the operator that can be built and run here
 *     (oracle/_ref/libref_mp.so); tests/test_oracle_cpu.py checks this file against it bit for bit (use_libm_exp=1).
 *   - variant GPU-build (flags & 1 / flags & 2): the two extras of multi_proposal.cu (host code as well):
 *       MP_SUPPRESS_TYPES  anchor types with (i+4)%7==0 || (i+2)%7==0 get score -1            (.cu:505-508)
 *       MP_FAST_NMS        "FastNMS": a kept proposal only tests the anchors listed in the precomputed
 *                          anchor-overlap map (IoU of the *undecoded* anchors >= roi_iou_thresh), with the map's
 *                          (row, column) offsets applied swapped as the reference does           (.cu:267-387, 511-575)
 *     pinned by the .cu text only (that file needs CUDA + libmxnet and cannot be built here).
 * use_libm_exp = 0 switches exp to oracle_expf, the correctly rounded sequence the CUDA kernel runs.
 * Rows past the kept count are rand() boxes in the reference (.cc:360-367); here they are the deterministic
 * (k%100, k%100, k%100+200, k%100+200) filler of the training operator and score 0 -- documented deviation. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void oracle_generate_anchors(int feat_stride, const float* ratios, int nr, const float* scales, int ns, float* out);
float oracle_expf(float x);

enum { MP_SUPPRESS_TYPES = 1, MP_FAST_NMS = 2 };

typedef struct {
  float score;
  int32_t id;
} MpKey;

static int mp_cmp(const void* a, const void* b) {
  const MpKey* x = (const MpKey*)a;
  const MpKey* y = (const MpKey*)b;
  if (x->score > y->score) return -1;
  if (x->score < y->score) return 1;
  return x->id < y->id ? -1 : (x->id > y->id ? 1 : 0);
}

/* anchor-overlap predicate of the FastNMS map (.cu:528-569): is there an entry (dx, dy, c2) in anchor_iou_dp[c1] with
 * dx == rel_w and dy == rel_h?  Entries are (+-j, +-k, c2) for every anchor c2 at grid row j, column k (image 0,
 * undecoded) whose IoU with anchor c1 at the origin is >= thresh, except (j = 0, k = 0, c2 = c1). */
static int in_overlap_map(const float* anchors, const float* anchor_area, int c1, int c2, int rel_w, int rel_h, int H,
                          int W, int stride, float thresh) {
  const int j = abs(rel_w), k = abs(rel_h); /* dx = +-j (a ROW index), dy = +-k (a COLUMN index): swapped on use */
  if (j >= H || k >= W) return 0;
  if (j == 0 && k == 0 && c2 == c1) return 0;
  const float a1x1 = anchors[4 * c1], a1y1 = anchors[4 * c1 + 1], a1x2 = anchors[4 * c1 + 2], a1y2 = anchors[4 * c1 + 3];
  const float a2x1 = anchors[4 * c2] + k * stride, a2y1 = anchors[4 * c2 + 1] + j * stride;
  const float a2x2 = anchors[4 * c2 + 2] + k * stride, a2y2 = anchors[4 * c2 + 3] + j * stride;
  float xx1 = fmaxf(a1x1, a2x1), yy1 = fmaxf(a1y1, a2y1);
  float xx2 = fminf(a1x2, a2x2), yy2 = fminf(a1y2, a2y2);
  float w = fmaxf(0.0f, xx2 - xx1 + 1.0f);
  float h = fmaxf(0.0f, yy2 - yy1 + 1.0f);
  float inter = w * h;
  float s0 = anchor_area[c1] + anchor_area[c2];
  float den = s0 - inter;
  float ovr = inter / den;
  return ovr >= thresh;
}

int oracle_multi_proposal(const float* cls_prob, const float* bbox_pred, const float* im_info, int B, int A, int H, int W,
                          int pre_nms_top_n, int post_nms_top_n, int feat_stride, const float* scales, int ns,
                          const float* ratios, int nr, int flags, float roi_iou_thresh, int use_libm_exp, float* rois,
                          float* scores_out, int32_t* keep_idx, int32_t* num_kept) {
  if (A != ns * nr) return -1;
  const int chip = A * H * W, hw = H * W;
  const size_t total = (size_t)B * chip;
  const int R = post_nms_top_n;
  float* anchors = (float*)malloc(sizeof(float) * 4 * A);
  float* anchor_area = (float*)malloc(sizeof(float) * A);
  float* P = (float*)malloc(sizeof(float) * 5 * total);
  float* area = (float*)malloc(sizeof(float) * total);
  oracle_generate_anchors(feat_stride, ratios, nr, scales, ns, anchors);
  for (int i = 0; i < A; ++i)
    anchor_area[i] = (anchors[4 * i + 2] - anchors[4 * i] + 1) * (anchors[4 * i + 3] - anchors[4 * i + 1] + 1);
  for (size_t t = 0; t < total; ++t) {
    const int b = (int)(t / chip), index = (int)(t % chip);
    const int a = index / hw, mat = index % hw, w = mat % W, h = mat / W;
    float* bx = P + 5 * t;
    bx[0] = anchors[4 * a] + w * feat_stride;
    bx[1] = anchors[4 * a + 1] + h * feat_stride;
    bx[2] = anchors[4 * a + 2] + w * feat_stride;
    bx[3] = anchors[4 * a + 3] + h * feat_stride;
    if ((flags & MP_SUPPRESS_TYPES) && ((a + 4) % 7 == 0 || (a + 2) % 7 == 0))
      bx[4] = -1;
    else
      bx[4] = cls_prob[(size_t)b * chip * 2 + ((size_t)(A + a) * H + h) * W + w];
    /* BBoxTransformInv (.cc:40-87) */
    float width = (float)(bx[2] - bx[0] + 1.0);
    float height = (float)(bx[3] - bx[1] + 1.0);
    float ctr_x = (float)(bx[0] + 0.5 * (width - 1.0));
    float ctr_y = (float)(bx[1] + 0.5 * (height - 1.0));
    const size_t dbase = (size_t)b * chip * 4;
    float dx = bbox_pred[dbase + (size_t)(a * 4 + 0) * hw + h * W + w];
    float dy = bbox_pred[dbase + (size_t)(a * 4 + 1) * hw + h * W + w];
    float dw = bbox_pred[dbase + (size_t)(a * 4 + 2) * hw + h * W + w];
    float dh = bbox_pred[dbase + (size_t)(a * 4 + 3) * hw + h * W + w];
    float t0 = dx * width;
    float pred_ctr_x = t0 + ctr_x;
    float t1 = dy * height;
    float pred_ctr_y = t1 + ctr_y;
    float pred_w = (use_libm_exp ? expf(dw) : oracle_expf(dw)) * width;
    float pred_h = (use_libm_exp ? expf(dh) : oracle_expf(dh)) * height;
    float x1 = (float)(pred_ctr_x - 0.5 * (pred_w - 1.0));
    float y1 = (float)(pred_ctr_y - 0.5 * (pred_h - 1.0));
    float x2 = (float)(pred_ctr_x + 0.5 * (pred_w - 1.0));
    float y2 = (float)(pred_ctr_y + 0.5 * (pred_h - 1.0));
    bx[0] = fmaxf(fminf(x1, im_info[3 * b + 1] - 1.0f), 0.0f);
    bx[1] = fmaxf(fminf(y1, im_info[3 * b] - 1.0f), 0.0f);
    bx[2] = fmaxf(fminf(x2, im_info[3 * b + 1] - 1.0f), 0.0f);
    bx[3] = fmaxf(fminf(y2, im_info[3 * b] - 1.0f), 0.0f);
    /* FilterBox(.., 3) (.cc:91-105) */
    const float min_size = 3;
    float iw = bx[2] - bx[0] + 1.0f;
    float ih = bx[3] - bx[1] + 1.0f;
    if (iw < min_size || ih < min_size) {
      bx[0] -= min_size / 2;
      bx[1] -= min_size / 2;
      bx[2] += min_size / 2;
      bx[3] += min_size / 2;
      bx[4] = -1.0f;
    }
    area[t] = (bx[2] - bx[0] + 1) * (bx[3] - bx[1] + 1);
  }
  const int max_nms = pre_nms_top_n < chip ? pre_nms_top_n : chip;
  MpKey* keys = (MpKey*)malloc(sizeof(MpKey) * chip);
  int32_t* keep_tmp = (int32_t*)malloc(sizeof(int32_t) * (R > 0 ? R : 1));
  for (int i = 0; i < B; ++i) {
    const size_t ci = (size_t)i * chip;
    for (int j = 0; j < chip; ++j) {
      keys[j].score = P[5 * (ci + j) + 4];
      keys[j].id = j;
    }
    qsort(keys, chip, sizeof(MpKey), mp_cmp);
    /* suppression state lives on the anchor itself: the plain NMS marks its sorted copy, the FastNMS marks dets[] of
     * ANY anchor (.cu:251-265) -- equivalent to one flag per anchor because only the top max_nms are ever visited */
    int vct = 0;
    for (int j = 0; j < max_nms && vct < R; ++j) {
      const int id1 = keys[j].id;
      float* d1 = P + 5 * (ci + id1);
      if (d1[4] == -1) continue;
      float* o = rois + 5 * ((size_t)i * R + vct);
      o[0] = i; o[1] = d1[0]; o[2] = d1[1]; o[3] = d1[2]; o[4] = d1[3];
      if (keep_idx) keep_idx[(size_t)i * R + vct] = id1;
      keep_tmp[vct] = id1;
      ++vct;
      const float ix1 = d1[0], iy1 = d1[1], ix2 = d1[2], iy2 = d1[3], iarea = area[ci + id1];
      const int c1 = id1 / hw, h1 = (id1 % hw) / W, w1 = (id1 % hw) % W;
      if (flags & MP_FAST_NMS) {
        /* every anchor of the chip that the overlap map of c1 reaches from (h1, w1) -- whatever its rank, including
         * anchors kept earlier (their output score then reads -1, .cu:592) */
        for (int id2 = 0; id2 < chip; ++id2) {
          const int c2 = id2 / hw, h2 = (id2 % hw) / W, w2 = (id2 % hw) % W;
          if (!in_overlap_map(anchors, anchor_area, c1, c2, w2 - w1, h2 - h1, H, W, feat_stride, roi_iou_thresh)) continue;
          float* d = P + 5 * (ci + id2);
          if (d[4] == -1) continue;
          float xx1 = fmaxf(ix1, d[0]), yy1 = fmaxf(iy1, d[1]);
          float xx2 = fminf(ix2, d[2]), yy2 = fminf(iy2, d[3]);
          float w = fmaxf(0.0f, xx2 - xx1 + 1.0f);
          float h = fmaxf(0.0f, yy2 - yy1 + 1.0f);
          float inter = w * h;
          float s0 = iarea + area[ci + id2];
          float den = s0 - inter;
          float ovr = inter / den;
          if (ovr > 0.7) d[4] = -1;
        }
      } else {
        for (int pind = j + 1; pind < max_nms; ++pind) {
          float* d = P + 5 * (ci + keys[pind].id);
          if (d[4] == -1) continue;
          float xx1 = fmaxf(ix1, d[0]), yy1 = fmaxf(iy1, d[1]);
          float xx2 = fminf(ix2, d[2]), yy2 = fminf(iy2, d[3]);
          float w = fmaxf(0.0f, xx2 - xx1 + 1.0f);
          float h = fmaxf(0.0f, yy2 - yy1 + 1.0f);
          float inter = w * h;
          float s0 = iarea + area[ci + keys[pind].id];
          float den = s0 - inter;
          float ovr = inter / den;
          if (ovr > 0.7) d[4] = -1;
        }
      }
    }
    /* output scores are read after the NMS (.cc:357 / .cu:592) */
    for (int j = 0; j < vct; ++j) scores_out[(size_t)i * R + j] = P[5 * (ci + keep_tmp[j]) + 4];
    if (num_kept) num_kept[i] = vct;
    for (int j = vct; j < R; ++j) {
      const int k = i * chip + j; /* global row index, as the training operator's filler (.cu:244-249) */
      float* o = rois + 5 * ((size_t)i * R + j);
      o[0] = i; o[1] = k % 100; o[2] = k % 100; o[3] = k % 100 + 200; o[4] = k % 100 + 200;
      scores_out[(size_t)i * R + j] = 0.0f;
      if (keep_idx) keep_idx[(size_t)i * R + j] = -1;
    }
  }
  free(keys);
  free(keep_tmp);
  free(anchors);
  free(anchor_area);
  free(P);
  free(area);
  return 0;
}
