// ORACLE harness (test infrastructure only): C entry point over the *reference's own*
// lib/chips/cchips.cpp (compiled from /root/reference where it lies, see Makefile `ref`).
#include <cstdlib>
#include <vector>
#include "cchips.h"

extern "C" int ref_chips_generate(const float* boxes, int num_boxes, int width, int height, int chipsize, int stride,
                                  unsigned seed, float* out_chips, int max_out) {
  std::vector<std::vector<float> > b(num_boxes, std::vector<float>(4));
  for (int i = 0; i < num_boxes; ++i)
    for (int c = 0; c < 4; ++c) b[i][c] = boxes[4 * i + c];
  srand(seed);
  std::vector<std::vector<float> > r = chips::cgenerate(width, height, chipsize, b, num_boxes, stride);
  int n = (int)r.size();
  for (int i = 0; i < n && i < max_out; ++i)
    for (int c = 0; c < 4; ++c) out_chips[4 * i + c] = r[i][c];
  return n;
}

// Same call WITHOUT re-seeding: the reference never seeds rand(), successive cgenerate calls of one worker process
// continue one rand() stream (lib/data_utils/data_workers.py:394-450 calls it once per scale).  Tests seed once with
// srand() and then drive whole chip_extractor / box_assigner sequences through this entry point.
extern "C" int ref_chips_generate_noseed(const float* boxes, int num_boxes, int width, int height, int chipsize,
                                         int stride, float* out_chips, int max_out) {
  std::vector<std::vector<float> > b(num_boxes, std::vector<float>(4));
  for (int i = 0; i < num_boxes; ++i)
    for (int c = 0; c < 4; ++c) b[i][c] = boxes[4 * i + c];
  std::vector<std::vector<float> > r = chips::cgenerate(width, height, chipsize, b, num_boxes, stride);
  int n = (int)r.size();
  for (int i = 0; i < n && i < max_out; ++i)
    for (int c = 0; c < 4; ++c) out_chips[4 * i + c] = r[i][c];
  return n;
}
