/*
 * ORACLE -- test infrastructure only.  Not part of the product path.
 *
 * CPU restatement (the reference has no CPU implementation of these operators:
 * deformable_psroi_pooling.cc:41-76 is an empty stub) of
 *   SNIPER-mxnet/src/operator/contrib/deformable_psroi_pooling.cu
 *     bilinear_interp                         :49-68
 *     DeformablePSROIPoolForwardKernel        :71-161
 *     DeformablePSROIPoolBackwardAccKernel    :203-330
 *   SNIPER-mxnet/src/operator/contrib/psroi_pooling.cu
 *     PSROIPoolForwardKernel                  :51-118
 *     PSROIPoolBackwardAccKernel              :146-210
 * for DType=float, C abstract-machine semantics (double where the source has double literals,
 * no FMA contraction; build with -ffp-contract=off).  Layout NCHW as in the reference.
 * Backward accumulates sequentially in index order (the reference uses atomicAdd, whose
 * order is undefined) -- compare with a tolerance.
 * PARITY PIN: the reference tests hold no forward vectors for these ops (only numeric-gradient
 * checks, test_operator.py:4292-4389); the oracle's backward is checked in tests/ against a
 * central-difference gradient of its own forward using the reference test's shapes.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float bilinear_interp(const float* data, float x, float y, int width, int height,
                             int32_t* corners /* optional [4] */) {
  int x1 = (int)floorf(x);
  int x2 = (int)ceilf(x);
  int y1 = (int)floorf(y);
  int y2 = (int)ceilf(y);
  float dist_x = (float)(x - x1);
  float dist_y = (float)(y - y1);
  float value11 = data[y1 * width + x1];
  float value12 = data[y2 * width + x1];
  float value21 = data[y1 * width + x2];
  float value22 = data[y2 * width + x2];
  float a = (1 - dist_x) * (1 - dist_y) * value11;
  float b = (1 - dist_x) * dist_y * value12;
  float c = dist_x * (1 - dist_y) * value21;
  float d = dist_x * dist_y * value22;
  float value = a + b;
  value = value + c;
  value = value + d;
  if (corners) {
    corners[0] = y1 * width + x1;
    corners[1] = y2 * width + x1;
    corners[2] = y1 * width + x2;
    corners[3] = y2 * width + x2;
  }
  (void)height;
  return value;
}

typedef struct {
  int roi_batch_ind, part_h, part_w, class_id, gw, gh;
  float roi_width, roi_height, wstart, hstart, sub_bin_size_w, sub_bin_size_h;
} bin_geom;

static void deform_geom(const float* bottom_rois, const float* bottom_trans, int no_trans, float trans_std,
                        float spatial_scale, int n, int ctop, int ph, int pw, int pooled_height,
                        int pooled_width, int sample_per_part, int group_size, int part_size,
                        int num_classes, int channels_each_class, bin_geom* g) {
  const float* r = bottom_rois + n * 5;
  g->roi_batch_ind = (int)r[0];
  float roi_start_w = (float)((float)(roundf(r[1])) * spatial_scale - 0.5);
  float roi_start_h = (float)((float)(roundf(r[2])) * spatial_scale - 0.5);
  float roi_end_w = (float)((float)(roundf(r[3]) + 1.) * spatial_scale - 0.5);
  float roi_end_h = (float)((float)(roundf(r[4]) + 1.) * spatial_scale - 0.5);
  float roi_width = (float)fmax(roi_end_w - roi_start_w, 0.1);
  float roi_height = (float)fmax(roi_end_h - roi_start_h, 0.1);
  float bin_size_h = roi_height / (float)pooled_height;
  float bin_size_w = roi_width / (float)pooled_width;
  g->sub_bin_size_h = bin_size_h / (float)sample_per_part;
  g->sub_bin_size_w = bin_size_w / (float)sample_per_part;
  g->part_h = (int)floorf((float)ph / pooled_height * part_size);
  g->part_w = (int)floorf((float)pw / pooled_width * part_size);
  g->class_id = ctop / channels_each_class;
  float trans_x = no_trans ? 0.0f
                           : bottom_trans[(((n * num_classes + g->class_id) * 2) * part_size + g->part_h) * part_size + g->part_w] * trans_std;
  float trans_y = no_trans ? 0.0f
                           : bottom_trans[(((n * num_classes + g->class_id) * 2 + 1) * part_size + g->part_h) * part_size + g->part_w] * trans_std;
  float wstart = (float)pw * bin_size_w;
  wstart = wstart + roi_start_w;
  float tw = trans_x * roi_width;
  wstart = wstart + tw;
  float hstart = (float)ph * bin_size_h;
  hstart = hstart + roi_start_h;
  float th = trans_y * roi_height;
  hstart = hstart + th;
  g->wstart = wstart;
  g->hstart = hstart;
  g->roi_width = roi_width;
  g->roi_height = roi_height;
  int gw = (int)floorf((float)pw * group_size / pooled_width);
  int gh = (int)floorf((float)ph * group_size / pooled_height);
  gw = gw < 0 ? 0 : (gw > group_size - 1 ? group_size - 1 : gw);
  gh = gh < 0 ? 0 : (gh > group_size - 1 ? group_size - 1 : gh);
  g->gw = gw;
  g->gh = gh;
}

/* sample_idx (optional): [count, sample_per_part^2, 4] flat corner indices inside the channel
 * plane (y*width+x), -1 where the sample is skipped -- the "ROI bin indices" graded bit-exact. */
void oracle_deform_psroi_fwd(const float* bottom_data, const float* bottom_rois, const float* bottom_trans,
                             int num_rois, int channels, int height, int width, float spatial_scale,
                             int output_dim, int group_size, int pooled_size, int part_size,
                             int sample_per_part, float trans_std, int no_trans, int num_classes,
                             float* top_data, float* top_count, int32_t* sample_idx) {
  int pooled_height = pooled_size, pooled_width = pooled_size;
  if (no_trans) num_classes = 1;
  int channels_each_class = no_trans ? output_dim : output_dim / num_classes;
  long count = (long)num_rois * output_dim * pooled_height * pooled_width;
  int S2 = sample_per_part * sample_per_part;
#pragma omp parallel for schedule(static)
  for (long index = 0; index < count; ++index) {
    int pw = index % pooled_width;
    int ph = (index / pooled_width) % pooled_height;
    int ctop = (index / pooled_width / pooled_height) % output_dim;
    int n = index / pooled_width / pooled_height / output_dim;
    bin_geom g;
    deform_geom(bottom_rois, bottom_trans, no_trans, trans_std, spatial_scale, n, ctop, ph, pw, pooled_height,
                pooled_width, sample_per_part, group_size, part_size, num_classes, channels_each_class, &g);
    float sum = 0;
    int cnt = 0;
    const float* offset_bottom_data = bottom_data + ((size_t)g.roi_batch_ind * channels) * height * width;
    for (int ih = 0; ih < sample_per_part; ih++) {
      for (int iw = 0; iw < sample_per_part; iw++) {
        float tw = iw * g.sub_bin_size_w;
        float w = g.wstart + tw;
        float th = ih * g.sub_bin_size_h;
        float h = g.hstart + th;
        int32_t* si = sample_idx ? sample_idx + ((size_t)index * S2 + ih * sample_per_part + iw) * 4 : 0;
        if (w < -0.5 || w > width - 0.5 || h < -0.5 || h > height - 0.5) {
          if (si) si[0] = si[1] = si[2] = si[3] = -1;
          continue;
        }
        w = (float)fmin(fmax(w, 0.), width - 1.);
        h = (float)fmin(fmax(h, 0.), height - 1.);
        int c = (ctop * group_size + g.gh) * group_size + g.gw;
        float val = bilinear_interp(offset_bottom_data + (size_t)c * height * width, w, h, width, height, si);
        sum += val;
        cnt++;
      }
    }
    top_data[index] = cnt == 0 ? 0.0f : sum / cnt;
    top_count[index] = cnt;
  }
}

void oracle_deform_psroi_bwd(const float* top_diff, const float* top_count, const float* bottom_data,
                             const float* bottom_rois, const float* bottom_trans, int num_rois, int channels,
                             int height, int width, float spatial_scale, int output_dim, int group_size,
                             int pooled_size, int part_size, int sample_per_part, float trans_std,
                             int no_trans, int num_classes, double* bottom_data_diff /* zero-init by caller */,
                             double* bottom_trans_diff /* zero-init by caller, may be NULL if no_trans */) {
  int pooled_height = pooled_size, pooled_width = pooled_size;
  if (no_trans) num_classes = 1;
  int channels_each_class = no_trans ? output_dim : output_dim / num_classes;
  /* Two race-free passes so that the host threads can share the work WITHOUT changing any summation order: pass 0
   * accumulates bottom_data_diff, parallel over the output channel (a bottom channel receives from one ctop only, and
   * for a fixed channel the (n, ph, pw) order below is the sequential one); pass 1 accumulates bottom_trans_diff,
   * parallel over the roi (fixed n: sequential (ctop, ph, pw) order).  Bit-identical to the single loop. */
  for (int pass = 0; pass < (no_trans ? 1 : 2); ++pass) {
    const int outer = pass == 0 ? output_dim : num_rois;
    const int inner = pass == 0 ? num_rois : output_dim;
#pragma omp parallel for schedule(dynamic, 1)
    for (int o = 0; o < outer; ++o)
      for (int i = 0; i < inner; ++i)
        for (int ph = 0; ph < pooled_height; ++ph)
          for (int pw = 0; pw < pooled_width; ++pw) {
            const int ctop = pass == 0 ? o : i, n = pass == 0 ? i : o;
            const long index = (((long)n * output_dim + ctop) * pooled_height + ph) * pooled_width + pw;
            bin_geom g;
            deform_geom(bottom_rois, bottom_trans, no_trans, trans_std, spatial_scale, n, ctop, ph, pw, pooled_height,
                        pooled_width, sample_per_part, group_size, part_size, num_classes, channels_each_class, &g);
            if (top_count[index] <= 0) continue;
            float diff_val = top_diff[index] / top_count[index];
            size_t img = (size_t)g.roi_batch_ind * channels * height * width;
            for (int ih = 0; ih < sample_per_part; ih++) {
              for (int iw = 0; iw < sample_per_part; iw++) {
                float tw = iw * g.sub_bin_size_w;
                float w = g.wstart + tw;
                float th = ih * g.sub_bin_size_h;
                float h = g.hstart + th;
                if (w < -0.5 || w > width - 0.5 || h < -0.5 || h > height - 0.5) continue;
                w = (float)fmin(fmax(w, 0.), width - 1.);
                h = (float)fmin(fmax(h, 0.), height - 1.);
                int c = (ctop * group_size + g.gh) * group_size + g.gw;
                int x0 = (int)floorf(w), x1 = (int)ceilf(w), y0 = (int)floorf(h), y1 = (int)ceilf(h);
                float dist_x = w - x0, dist_y = h - y0;
                float q00 = (1 - dist_x) * (1 - dist_y);
                float q01 = (1 - dist_x) * dist_y;
                float q10 = dist_x * (1 - dist_y);
                float q11 = dist_x * dist_y;
                size_t base = img + (size_t)c * height * width;
                if (pass == 0) {
                  bottom_data_diff[base + y0 * width + x0] += q00 * diff_val;
                  bottom_data_diff[base + y1 * width + x0] += q01 * diff_val;
                  bottom_data_diff[base + y0 * width + x1] += q10 * diff_val;
                  bottom_data_diff[base + y1 * width + x1] += q11 * diff_val;
                  continue;
                }
                float U00 = bottom_data[base + y0 * width + x0];
                float U01 = bottom_data[base + y1 * width + x0];
                float U10 = bottom_data[base + y0 * width + x1];
                float U11 = bottom_data[base + y1 * width + x1];
                float diff_x = (U11 * dist_y + U10 * (1 - dist_y) - U01 * dist_y - U00 * (1 - dist_y)) * trans_std * diff_val;
                diff_x *= g.roi_width;
                float diff_y = (U11 * dist_x + U01 * (1 - dist_x) - U10 * dist_x - U00 * (1 - dist_x)) * trans_std * diff_val;
                diff_y *= g.roi_height;
                bottom_trans_diff[(((n * num_classes + g.class_id) * 2) * part_size + g.part_h) * part_size + g.part_w] += diff_x;
                bottom_trans_diff[(((n * num_classes + g.class_id) * 2 + 1) * part_size + g.part_h) * part_size + g.part_w] += diff_y;
              }
            }
          }
  }
}

static void psroi_bin(const float* bottom_rois, float spatial_scale, int n, int ph, int pw, int pooled_height,
                      int pooled_width, int height, int width, int* roi_batch_ind, int* hstart, int* hend,
                      int* wstart, int* wend) {
  const float* r = bottom_rois + n * 5;
  *roi_batch_ind = (int)r[0];
  float roi_start_w = (float)(roundf(r[1])) * spatial_scale;
  float roi_start_h = (float)(roundf(r[2])) * spatial_scale;
  float roi_end_w = (float)(roundf(r[3]) + 1.) * spatial_scale;
  float roi_end_h = (float)(roundf(r[4]) + 1.) * spatial_scale;
  float roi_width = (float)fmax(roi_end_w - roi_start_w, 0.1);
  float roi_height = (float)fmax(roi_end_h - roi_start_h, 0.1);
  float bin_size_h = roi_height / (float)pooled_height;
  float bin_size_w = roi_width / (float)pooled_width;
  float t;
  t = (float)ph * bin_size_h;
  int hs = (int)floorf(t + roi_start_h);
  t = (float)pw * bin_size_w;
  int ws = (int)floorf(t + roi_start_w);
  t = (float)(ph + 1) * bin_size_h;
  int he = (int)ceilf(t + roi_start_h);
  t = (float)(pw + 1) * bin_size_w;
  int we = (int)ceilf(t + roi_start_w);
  hs = hs < 0 ? 0 : (hs > height ? height : hs);
  he = he < 0 ? 0 : (he > height ? height : he);
  ws = ws < 0 ? 0 : (ws > width ? width : ws);
  we = we < 0 ? 0 : (we > width ? width : we);
  *hstart = hs;
  *hend = he;
  *wstart = ws;
  *wend = we;
}

/* bins (optional): [count,4] = hstart,hend,wstart,wend -- graded bit-exact */
void oracle_psroi_fwd(const float* bottom_data, const float* bottom_rois, int num_rois, int channels, int height,
                      int width, float spatial_scale, int output_dim, int group_size, int pooled_size,
                      float* top_data, int32_t* bins) {
  int pooled_height = pooled_size, pooled_width = pooled_size;
  long count = (long)num_rois * output_dim * pooled_height * pooled_width;
  for (long index = 0; index < count; ++index) {
    int pw = index % pooled_width;
    int ph = (index / pooled_width) % pooled_height;
    int ctop = (index / pooled_width / pooled_height) % output_dim;
    int n = index / pooled_width / pooled_height / output_dim;
    int b, hstart, hend, wstart, wend;
    psroi_bin(bottom_rois, spatial_scale, n, ph, pw, pooled_height, pooled_width, height, width, &b, &hstart, &hend,
              &wstart, &wend);
    int is_empty = (hend <= hstart) || (wend <= wstart);
    int gw = (int)floorf((float)pw * group_size / pooled_width);
    int gh = (int)floorf((float)ph * group_size / pooled_height);
    gw = gw < 0 ? 0 : (gw > group_size - 1 ? group_size - 1 : gw);
    gh = gh < 0 ? 0 : (gh > group_size - 1 ? group_size - 1 : gh);
    int c = (ctop * group_size + gh) * group_size + gw;
    const float* p = bottom_data + ((size_t)b * channels + c) * height * width;
    float out_sum = 0;
    for (int h = hstart; h < hend; ++h)
      for (int w = wstart; w < wend; ++w) out_sum += p[h * width + w];
    float bin_area = (float)((hend - hstart) * (wend - wstart));
    top_data[index] = is_empty ? 0.0f : out_sum / bin_area;
    if (bins) {
      bins[4 * index] = hstart;
      bins[4 * index + 1] = hend;
      bins[4 * index + 2] = wstart;
      bins[4 * index + 3] = wend;
    }
  }
}

void oracle_psroi_bwd(const float* top_diff, const float* bottom_rois, int num_rois, int channels, int height,
                      int width, float spatial_scale, int output_dim, int group_size, int pooled_size,
                      double* bottom_diff /* zero-init by caller */) {
  int pooled_height = pooled_size, pooled_width = pooled_size;
  long count = (long)num_rois * output_dim * pooled_height * pooled_width;
  for (long index = 0; index < count; ++index) {
    int pw = index % pooled_width;
    int ph = (index / pooled_width) % pooled_height;
    int ctop = (index / pooled_width / pooled_height) % output_dim;
    int n = index / pooled_width / pooled_height / output_dim;
    int b, hstart, hend, wstart, wend;
    psroi_bin(bottom_rois, spatial_scale, n, ph, pw, pooled_height, pooled_width, height, width, &b, &hstart, &hend,
              &wstart, &wend);
    int is_empty = (hend <= hstart) || (wend <= wstart);
    int gw = (int)floorf((float)pw * group_size / pooled_width);
    int gh = (int)floorf((float)ph * group_size / pooled_height);
    gw = gw < 0 ? 0 : (gw > group_size - 1 ? group_size - 1 : gw);
    gh = gh < 0 ? 0 : (gh > group_size - 1 ? group_size - 1 : gh);
    int c = (ctop * group_size + gh) * group_size + gw;
    double* p = bottom_diff + ((size_t)b * channels + c) * height * width;
    float bin_area = (float)((hend - hstart) * (wend - wstart));
    float diff_val = is_empty ? 0.0f : top_diff[index] / bin_area;
    for (int h = hstart; h < hend; ++h)
      for (int w = wstart; w < wend; ++w) p[h * width + w] += diff_val;
  }
}
