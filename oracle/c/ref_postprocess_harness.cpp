// Thin extern "C" shim over the REFERENCE's own header, compiled where it lies:
//   g++ -I/root/reference/unina_yolo_dla/ros2_ws/src/perception/include ... -> oracle/_ref/
// Nothing of the reference is copied into this repo; this file only calls
// decode_head / nms / compute_iou from postprocess.hpp.  TEST INFRASTRUCTURE.
#include <vector>

#include "postprocess.hpp"  // the reference header (Detection, compute_iou, nms, decode_head)

struct RefDet { float x1, y1, x2, y2, conf; int cls; };

extern "C" {

int ref_decode_head(const float *cls, const float *reg, int w, int h, int stride, int nc,
                    float thr, float q, RefDet *out, int cap) {
  std::vector<Detection> dets;
  decode_head(cls, reg, w, h, stride, nc, thr, q, dets);
  int n = (int)dets.size();
  for (int i = 0; i < n && i < cap; ++i)
    out[i] = RefDet{dets[i].x1, dets[i].y1, dets[i].x2, dets[i].y2, dets[i].confidence,
                    dets[i].class_id};
  return n;
}

// Runs the header's nms (sort + greedy); returns kept detections in kept order.
int ref_nms(const RefDet *in, int n, float thr, RefDet *out) {
  std::vector<Detection> dets(n);
  for (int i = 0; i < n; ++i) {
    dets[i].x1 = in[i].x1; dets[i].y1 = in[i].y1; dets[i].x2 = in[i].x2; dets[i].y2 = in[i].y2;
    dets[i].confidence = in[i].conf; dets[i].class_id = in[i].cls;
  }
  std::vector<Detection> kept = nms(dets, thr);
  for (size_t i = 0; i < kept.size(); ++i)
    out[i] = RefDet{kept[i].x1, kept[i].y1, kept[i].x2, kept[i].y2, kept[i].confidence,
                    kept[i].class_id};
  return (int)kept.size();
}

float ref_iou(const RefDet *a, const RefDet *b) {
  Detection da{a->x1, a->y1, a->x2, a->y2, a->conf, a->cls};
  Detection db{b->x1, b->y1, b->x2, b->y2, b->conf, b->cls};
  return compute_iou(da, db);
}
}
