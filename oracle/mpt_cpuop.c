/* TEST INFRASTRUCTURE ONLY.  Restatement of the reference's *CPU* operator MultiProposalTargetOp<cpu>::Forward
 * (SNIPER-mxnet/src/operator/multi_proposal_target.cc:40-497), written in the same style as oracle/mpt.c (which
 * restates the *GPU* operator, multi_proposal_target.cu) and sharing its anchor generator.
 *
 * Why it exists: the reference CPU operator is the one piece of the MultiProposalTarget path that can be BUILT AND RUN
 * here (oracle/_ref/libref_mpt.so, recipe in oracle/Makefile).  tests/test_oracle_cpu.py checks this restatement
 * bit for bit against that binary, and then checks oracle/mpt.c against this file wherever the .cc and the .cu agree
 * (anchor grid, delta decode + clip, GT append, IoU assignment, regression targets).  The lines where the GPU operator
 * deliberately differs are listed in DESIGN.md section 4 and stay pinned by the .cu text only.
 *
 * Differences from the GPU operator restated in mpt.c (cc line numbers):
 *   exp            libm expf (cc:68-69; the .cu uses CUDA expf -> oracle_expf)
 *   min size       (w+1 < 3 || h+1 < 3), box grown by 1.5 (float min_size / 2)              (cc:89-103)
 *   area / range   (w+1)*(h+1); score = -1 if area > hi^2 || area < lo^2                     (cc:164-168)
 *   NMS            std::sort by score, only the top 6000 enter; `continue` on -1; IoU areas with +1   (cc:171-229)
 *   fillers        (0, 0, 100, 100)                                                          (cc:343-350)
 *   targets        bbox_scale * (5, 5, 10, 10)                                               (cc:491-494)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void oracle_generate_anchors(int feat_stride, const float* ratios, int nr, const float* scales, int ns, float* out);

typedef struct {
  float score;
  int32_t id;
} SortKey;

/* descending score; ties by ascending index (std::sort leaves tie order unspecified: the tests use tie-free scores) */
static int cmp_key(const void* a, const void* b) {
  const SortKey* x = (const SortKey*)a;
  const SortKey* y = (const SortKey*)b;
  if (x->score > y->score) return -1;
  if (x->score < y->score) return 1;
  return x->id < y->id ? -1 : (x->id > y->id ? 1 : 0);
}

/* dets_out (optional): [B*A*H*W*5] decoded + filtered rows x1,y1,x2,y2,score as they enter the NMS (cc:311-313) */
int oracle_multi_proposal_target_cpuop(const float* cls_prob, const float* bbox_pred, const float* im_info,
                                       const float* gt_boxes, const float* valid_ranges, int B, int A, int H, int W,
                                       int post_nms_top_n, int feat_stride, const float* scales, int ns,
                                       const float* ratios, int nr, float bbox_scale, float* rois, float* label,
                                       float* bbox_target, float* bbox_weight, int32_t* keep_idx, int32_t* num_kept,
                                       float* dets_out) {
  const int max_nms = 6000;
  if (A != ns * nr) return -1;
  const int chip = A * H * W;
  if (chip < max_nms) return -1; /* the reference reads 6000 sorted rows unconditionally (cc:171,186) */
  const size_t total = (size_t)B * chip;
  float* anchors = (float*)malloc(sizeof(float) * 4 * A);
  float* P = (float*)malloc(sizeof(float) * 5 * total);
  float* area = (float*)malloc(sizeof(float) * total);
  oracle_generate_anchors(feat_stride, ratios, nr, scales, ns, anchors);
  const int hw = H * W;
  /* anchor grid + scores (cc:296-309), BBoxTransformInv (cc:40-87) */
  for (size_t t = 0; t < total; ++t) {
    const int b = (int)(t / chip), index = (int)(t % chip);
    const int a = index / hw, mat = index % hw, w = mat % W, h = mat / W;
    float* bx = P + 5 * t;
    bx[0] = anchors[4 * a] + w * feat_stride;
    bx[1] = anchors[4 * a + 1] + h * feat_stride;
    bx[2] = anchors[4 * a + 2] + w * feat_stride;
    bx[3] = anchors[4 * a + 3] + h * feat_stride;
    bx[4] = cls_prob[(size_t)b * chip * 2 + ((size_t)(A + a) * H + h) * W + w];
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
    float pred_w = expf(dw) * width;
    float pred_h = expf(dh) * height;
    float x1 = (float)(pred_ctr_x - 0.5 * (pred_w - 1.0));
    float y1 = (float)(pred_ctr_y - 0.5 * (pred_h - 1.0));
    float x2 = (float)(pred_ctr_x + 0.5 * (pred_w - 1.0));
    float y2 = (float)(pred_ctr_y + 0.5 * (pred_h - 1.0));
    bx[0] = fmaxf(fminf(x1, im_info[3 * b + 1] - 1.0f), 0.0f);
    bx[1] = fmaxf(fminf(y1, im_info[3 * b] - 1.0f), 0.0f);
    bx[2] = fmaxf(fminf(x2, im_info[3 * b + 1] - 1.0f), 0.0f);
    bx[3] = fmaxf(fminf(y2, im_info[3 * b] - 1.0f), 0.0f);
  }
  /* FilterBox(proposals, total, 3) (cc:89-103) */
  for (size_t t = 0; t < total; ++t) {
    float* bx = P + 5 * t;
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
  }
  /* NonMaximumSuppression (cc:148-232): areas + valid-range filter, sort, greedy over the top 6000 */
  for (size_t t = 0; t < total; ++t) {
    float* bx = P + 5 * t;
    area[t] = (bx[2] - bx[0] + 1) * (bx[3] - bx[1] + 1);
    const int b = (int)(t / chip);
    if (area[t] > valid_ranges[2 * b + 1] * valid_ranges[2 * b + 1] || area[t] < valid_ranges[2 * b] * valid_ranges[2 * b])
      bx[4] = -1;
  }
  if (dets_out) memcpy(dets_out, P, sizeof(float) * 5 * total);
  SortKey* keys = (SortKey*)malloc(sizeof(SortKey) * chip);
  float* dbuf = (float*)malloc(sizeof(float) * 6 * max_nms);
  const int R = post_nms_top_n;
  for (int i = 0; i < B; ++i) {
    const size_t ci = (size_t)i * chip;
    for (int j = 0; j < chip; ++j) {
      keys[j].score = P[5 * (ci + j) + 4];
      keys[j].id = j;
    }
    qsort(keys, chip, sizeof(SortKey), cmp_key);
    for (int j = 0; j < max_nms; ++j) {
      const size_t idx = ci + keys[j].id;
      for (int c = 0; c < 5; ++c) dbuf[6 * j + c] = P[5 * idx + c];
      dbuf[6 * j + 5] = area[idx];
    }
    int vct = 0;
    for (int j = 0; j < max_nms && vct < R; ++j) {
      if (dbuf[6 * j + 4] == -1) continue;
      const float ix1 = dbuf[6 * j], iy1 = dbuf[6 * j + 1], ix2 = dbuf[6 * j + 2], iy2 = dbuf[6 * j + 3];
      const float iarea = dbuf[6 * j + 5];
      float* o = rois + 5 * ((size_t)i * R + vct);
      o[0] = i; o[1] = ix1; o[2] = iy1; o[3] = ix2; o[4] = iy2;
      if (keep_idx) keep_idx[(size_t)i * R + vct] = keys[j].id;
      ++vct;
      for (int pind = j + 1; pind < max_nms; ++pind) {
        float* d = dbuf + 6 * pind;
        if (d[4] == -1) continue;
        float xx1 = fmaxf(ix1, d[0]), yy1 = fmaxf(iy1, d[1]);
        float xx2 = fminf(ix2, d[2]), yy2 = fminf(iy2, d[3]);
        float w = fmaxf(0.0f, xx2 - xx1 + 1.0f);
        float h = fmaxf(0.0f, yy2 - yy1 + 1.0f);
        float inter = w * h;
        float s0 = iarea + d[5];
        float den = s0 - inter;
        float ovr = inter / den;
        if (ovr > 0.7) d[4] = -1;
      }
    }
    if (num_kept) num_kept[i] = vct;
    for (int j = vct; j < R; ++j) { /* fillers (cc:343-350) */
      float* o = rois + 5 * ((size_t)i * R + j);
      o[0] = i; o[1] = 0; o[2] = 0; o[3] = 100; o[4] = 100;
      if (keep_idx) keep_idx[(size_t)i * R + j] = -1;
    }
  }
  free(keys);
  free(dbuf);
  /* GT append + assignment + targets (cc:354-497) */
  const int G5 = 100 * 5;
  for (int i = 0; i < B; ++i) {
    int numgt = 0;
    for (int j = 0; j < 100; ++j)
      if (gt_boxes[i * G5 + j * 5 + 4] != -1) numgt++;
    for (int j = 0; j < R; ++j) {
      const size_t bp = (size_t)R * i + j;
      label[bp] = 0;
      for (int c = 0; c < 4; ++c) {
        bbox_target[4 * bp + c] = 1.0;
        bbox_weight[4 * bp + c] = 0.0;
      }
    }
    for (int k = R - numgt, j = 0; k < R; ++j, ++k) {
      float w = gt_boxes[i * G5 + j * 5 + 2] - gt_boxes[i * G5 + j * 5];
      float h = gt_boxes[i * G5 + j * 5 + 3] - gt_boxes[i * G5 + j * 5 + 1];
      float ar = w * h;
      if (ar >= valid_ranges[2 * i] * valid_ranges[2 * i] && ar <= valid_ranges[2 * i + 1] * valid_ranges[2 * i + 1])
        for (int c = 0; c < 4; ++c) rois[((size_t)i * R + k) * 5 + 1 + c] = gt_boxes[i * G5 + j * 5 + c];
    }
    if (numgt == 0) continue;
    float* max_ov = (float*)calloc(R, sizeof(float));
    int* max_id = (int*)calloc(R, sizeof(int));
    char* pos = (char*)calloc(R, 1);
    for (int g = 0; g < numgt; ++g) {
      const float x1 = gt_boxes[i * G5 + g * 5], y1 = gt_boxes[i * G5 + g * 5 + 1];
      const float x2 = gt_boxes[i * G5 + g * 5 + 2], y2 = gt_boxes[i * G5 + g * 5 + 3];
      const float a1 = (x2 - x1) * (y2 - y1);
      for (int j = 0; j < R; ++j) {
        const float* r = rois + ((size_t)R * i + j) * 5;
        float xx1 = fmaxf(x1, r[1]), yy1 = fmaxf(y1, r[2]);
        float xx2 = fminf(x2, r[3]), yy2 = fminf(y2, r[4]);
        float w = fmaxf(0.0f, xx2 - xx1 + 1.0f);
        float h = fmaxf(0.0f, yy2 - yy1 + 1.0f);
        float a2 = (r[3] - r[1]) * (r[4] - r[2]);
        float inter = w * h;
        float s0 = a1 + a2;
        float den = s0 - inter;
        float ovr = inter / den;
        if (ovr > max_ov[j] && ovr > 0.5) {
          max_ov[j] = ovr;
          max_id[j] = g;
          label[(size_t)i * R + j] = gt_boxes[i * G5 + g * 5 + 4];
          pos[j] = 1;
        }
      }
    }
    for (int pid = 0; pid < R; ++pid) {
      if (!pos[pid]) continue;
      const size_t bid = (size_t)i * R + pid;
      for (int c = 0; c < 4; ++c) bbox_weight[bid * 4 + c] = 1;
      const int g = max_id[pid];
      const float gx1 = gt_boxes[i * G5 + g * 5], gy1 = gt_boxes[i * G5 + g * 5 + 1];
      const float gx2 = gt_boxes[i * G5 + g * 5 + 2], gy2 = gt_boxes[i * G5 + g * 5 + 3];
      float gw = gx2 - gx1 + 1;
      float gh = gy2 - gy1 + 1;
      float gcx = (float)(gx1 + gw * 0.5);
      float gcy = (float)(gy1 + gh * 0.5);
      const float px1 = rois[bid * 5 + 1], py1 = rois[bid * 5 + 2], px2 = rois[bid * 5 + 3], py2 = rois[bid * 5 + 4];
      float pw = px2 - px1 + 1;
      float ph = py2 - py1 + 1;
      float pcx = (float)(px1 + (pw - 1) * 0.5);
      float pcy = (float)(py1 + (ph - 1) * 0.5);
      bbox_target[4 * bid] = (float)(bbox_scale * 5 * (gcx - pcx) / (pw + 1e-7));
      bbox_target[4 * bid + 1] = (float)(bbox_scale * 5 * (gcy - pcy) / (ph + 1e-7));
      bbox_target[4 * bid + 2] = (float)(bbox_scale * 10 * log(gw / (pw + 1e-7)));
      bbox_target[4 * bid + 3] = (float)(bbox_scale * 10 * log(gh / (ph + 1e-7)));
    }
    free(max_ov);
    free(max_id);
    free(pos);
  }
  free(anchors);
  free(P);
  free(area);
  return 0;
}
