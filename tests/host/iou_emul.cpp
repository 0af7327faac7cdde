// CPU build of the per-pair body of the 2D-keypoint based 3D IoU kernel (csrc/iou_core.cuh is plain C++ behind TD3D_HD):
//   iou_emul <pairs.f32> <n> <portrait>   ->  one line per pair: iou, then the 27 lifted coordinates of the first set
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../3d-object-detection.pytorch_b200/csrc/iou_core.cuh"

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const int n = atoi(argv[2]), portrait = atoi(argv[3]);
  std::vector<float> buf((size_t)n * 36);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(buf.data(), sizeof(float), buf.size(), f) != buf.size()) return 3;
  fclose(f);
  const double cam[4] = {2.0, 2.0, 0.0, 0.0};      // NDC form of the reference's default camera matrix (geometry.py:16-37)
  for (int i = 0; i < n; ++i) {
    const float* pred = buf.data() + (size_t)i * 36;
    const float* gt = pred + 18;
    double lifted[9][3];
    td3d::iou3d::lift_2d(pred, portrait, cam, lifted);
    printf("%.17g", td3d::iou3d::iou_from_keypoints(pred, gt, portrait, cam));
    for (int k = 0; k < 9; ++k)
      for (int c = 0; c < 3; ++c) printf(" %.17g", lifted[k][c]);
    printf("\n");
  }
  return 0;
}
