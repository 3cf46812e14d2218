/*
 * ORACLE -- test infrastructure only.  Not part of the product path.
 *
 * C restatements of the reference's host-side Cython:
 *   lib/nms/cpu_nms.pyx   cpu_nms       :112-163  (order = scores.argsort()[::-1] is supplied by the caller,
 *                                                  because numpy's tie order is the reference's tie order)
 *                         cpu_soft_nms  :17-110   (method 2 = gaussian as used by nms.py:7-12; in-place swaps)
 *   lib/bbox/bbox.pyx     bbox_overlaps_cython :17-57, ignore_overlaps_cython :59-95 (float64, +1 convention)
 * PARITY PIN: tests/test_host_cpu.py checks this file (and the product's host_ops.cpp) against the reference's OWN
 * Cython modules, compiled from /root/reference by oracle/build_ref_cython.py into oracle/_ref: keep lists, surviving
 * rows and their order, overlaps bit for bit; soft-NMS scores bit for bit (linear, hard) / within 2e-7 (Gaussian: the
 * reference calls numpy's exp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

int oracle_cpu_nms(const float* dets, const int64_t* order, int ndets, double thresh, int32_t* keep) {
  float* areas = (float*)malloc(sizeof(float) * (ndets > 0 ? ndets : 1));
  char* suppressed = (char*)calloc(ndets > 0 ? ndets : 1, 1);
  for (int i = 0; i < ndets; ++i) {
    float w = dets[5 * i + 2] - dets[5 * i] + 1;
    float h = dets[5 * i + 3] - dets[5 * i + 1] + 1;
    areas[i] = w * h;
  }
  int nkeep = 0;
  for (int _i = 0; _i < ndets; ++_i) {
    int i = (int)order[_i];
    if (suppressed[i]) continue;
    keep[nkeep++] = i;
    float ix1 = dets[5 * i], iy1 = dets[5 * i + 1], ix2 = dets[5 * i + 2], iy2 = dets[5 * i + 3];
    float iarea = areas[i];
    for (int _j = _i + 1; _j < ndets; ++_j) {
      int j = (int)order[_j];
      if (suppressed[j]) continue;
      float xx1 = fmaxf(ix1, dets[5 * j]), yy1 = fmaxf(iy1, dets[5 * j + 1]);
      float xx2 = fminf(ix2, dets[5 * j + 2]), yy2 = fminf(iy2, dets[5 * j + 3]);
      float w = (float)fmax(0.0, xx2 - xx1 + 1);
      float h = (float)fmax(0.0, yy2 - yy1 + 1);
      float inter = w * h;
      float s0 = iarea + areas[j];
      float ovr = inter / (s0 - inter);
      if (ovr >= thresh) suppressed[j] = 1;
    }
  }
  free(areas);
  free(suppressed);
  return nkeep;
}

/* boxes [N,5] modified in place; returns the new N (rows [0,N) are the survivors, in final order) */
int oracle_cpu_soft_nms(float* boxes, int N, float sigma, float Nt, float threshold, unsigned method) {
  for (int i = 0; i < N; ++i) {
    float maxscore = boxes[5 * i + 4];
    int maxpos = i;
    float tx1 = boxes[5 * i], ty1 = boxes[5 * i + 1], tx2 = boxes[5 * i + 2], ty2 = boxes[5 * i + 3], ts = boxes[5 * i + 4];
    int pos = i + 1;
    while (pos < N) {
      if (maxscore < boxes[5 * pos + 4]) {
        maxscore = boxes[5 * pos + 4];
        maxpos = pos;
      }
      pos = pos + 1;
    }
    for (int c = 0; c < 5; ++c) boxes[5 * i + c] = boxes[5 * maxpos + c];
    boxes[5 * maxpos] = tx1; boxes[5 * maxpos + 1] = ty1; boxes[5 * maxpos + 2] = tx2; boxes[5 * maxpos + 3] = ty2;
    boxes[5 * maxpos + 4] = ts;
    tx1 = boxes[5 * i]; ty1 = boxes[5 * i + 1]; tx2 = boxes[5 * i + 2]; ty2 = boxes[5 * i + 3]; ts = boxes[5 * i + 4];
    pos = i + 1;
    while (pos < N) {
      float x1 = boxes[5 * pos], y1 = boxes[5 * pos + 1], x2 = boxes[5 * pos + 2], y2 = boxes[5 * pos + 3];
      /* Cython promotes the Python-int literal in `x2 - x1 + 1` to the C double 1.0 (generated C: ((x2 - x1) + 1.0)):
       * the sums and the products below run in double and are narrowed once, on the assignment to a float variable */
      float area = (float)(((double)(x2 - x1) + 1.0) * ((double)(y2 - y1) + 1.0));
      float iw = (float)((double)(fminf(tx2, x2) - fmaxf(tx1, x1)) + 1.0);
      if (iw > 0) {
        float ih = (float)((double)(fminf(ty2, y2) - fmaxf(ty1, y1)) + 1.0);
        if (ih > 0) {
          float t2 = iw * ih;
          float ua = (float)(((((double)(tx2 - tx1) + 1.0) * ((double)(ty2 - ty1) + 1.0)) + (double)area) - (double)t2);
          float ov = t2 / ua;
          float weight;
          if (method == 1) weight = ov > Nt ? (float)(1.0 - (double)ov) : 1;
          else if (method == 2) {
            float q = -(ov * ov) / sigma;
            weight = (float)exp((double)q);
          } else weight = ov > Nt ? 0 : 1;
          boxes[5 * pos + 4] = weight * boxes[5 * pos + 4];
          if (boxes[5 * pos + 4] < threshold) {
            for (int c = 0; c < 5; ++c) boxes[5 * pos + c] = boxes[5 * (N - 1) + c];
            N = N - 1;
            pos = pos - 1;
          }
        }
      }
      pos = pos + 1;
    }
  }
  return N;
}

void oracle_bbox_overlaps(const double* boxes, int N, const double* query, int K, double* overlaps, int ignore) {
  for (int i = 0; i < N * K; ++i) overlaps[i] = 0;
  for (int k = 0; k < K; ++k) {
    double box_area = (query[4 * k + 2] - query[4 * k] + 1) * (query[4 * k + 3] - query[4 * k + 1] + 1);
    for (int n = 0; n < N; ++n) {
      double iw = fmin(boxes[4 * n + 2], query[4 * k + 2]) - fmax(boxes[4 * n], query[4 * k]) + 1;
      if (iw > 0) {
        double ih = fmin(boxes[4 * n + 3], query[4 * k + 3]) - fmax(boxes[4 * n + 1], query[4 * k + 1]) + 1;
        if (ih > 0) {
          double ua;
          if (ignore) ua = box_area;
          else {
            double t = (boxes[4 * n + 2] - boxes[4 * n] + 1) * (boxes[4 * n + 3] - boxes[4 * n + 1] + 1);
            t = t + box_area;
            double u = iw * ih;
            ua = t - u;
          }
          double num = iw * ih;
          overlaps[(size_t)n * K + k] = num / ua;
        }
      }
    }
  }
}
