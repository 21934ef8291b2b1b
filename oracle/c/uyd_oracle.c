/* CPU oracle, plain C.  TEST INFRASTRUCTURE -- the product never links this.
 *
 * Restates, in scalar fp32 arithmetic (compile with -ffp-contract=off):
 *   - torchvision.ops.nms semantics used by Ultralytics non_max_suppression
 *     (reference call sites train.py:396-405, eval.py:32):
 *     stable score-descending order, suppress iff IoU > thr, IoU = inter/(a_i+a_j-inter)
 *     with max(0,.) clamps and no epsilon.
 *   - the reference's own CPU post-processing header
 *     ros2_ws/src/perception/include/postprocess.hpp: compute_iou :28-39,
 *     nms :44-67, apply_conformal_prediction :77-85, decode_head :94-145.
 * Pinned by tests/test_oracle_pins.py against torchvision and against the compiled
 * reference header (oracle/_ref/libref_postprocess.so).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- torchvision-style greedy NMS on boxes already sorted by the caller ------------ */
/* boxes: [n,4] xyxy (class offset already added), in processing order.
 * keep_out: indices (into the sorted order) of kept boxes, at most max_keep of them.
 * Returns the number kept.  Work stops once max_keep boxes are kept (the caller
 * truncates to max_det anyway and greedy kept order == score order). */
int uydo_nms_sorted(const float *boxes, int n, double iou_thr, int max_keep, int *keep_out) {
  unsigned char *sup = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
  float *area = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    const float *b = boxes + 4 * (size_t)i;
    area[i] = (b[2] - b[0]) * (b[3] - b[1]);
  }
  for (int i = 0; i < n && kept < max_keep; ++i) {
    if (sup[i]) continue;
    keep_out[kept++] = i;
    const float *a = boxes + 4 * (size_t)i;
    const float ax1 = a[0], ay1 = a[1], ax2 = a[2], ay2 = a[3], aa = area[i];
    for (int j = i + 1; j < n; ++j) {
      if (sup[j]) continue;
      const float *b = boxes + 4 * (size_t)j;
      float xx1 = ax1 > b[0] ? ax1 : b[0];
      float yy1 = ay1 > b[1] ? ay1 : b[1];
      float xx2 = ax2 < b[2] ? ax2 : b[2];
      float yy2 = ay2 < b[3] ? ay2 : b[3];
      float w = xx2 - xx1; if (!(w > 0.0f)) w = 0.0f;
      float h = yy2 - yy1; if (!(h > 0.0f)) h = 0.0f;
      float inter = w * h;
      float ovr = inter / (aa + area[j] - inter);
      /* torchvision compares the fp32 IoU against the *double* threshold */
      if ((double)ovr > iou_thr) sup[j] = 1;
    }
  }
  free(sup);
  free(area);
  return kept;
}

/* ---- postprocess.hpp restated ------------------------------------------------------ */
typedef struct { float x1, y1, x2, y2, conf; int cls; } uydo_det;

static float hpp_iou(const uydo_det *a, const uydo_det *b) { /* postprocess.hpp:28-39 */
  float ix1 = a->x1 > b->x1 ? a->x1 : b->x1;
  float iy1 = a->y1 > b->y1 ? a->y1 : b->y1;
  float ix2 = a->x2 < b->x2 ? a->x2 : b->x2;
  float iy2 = a->y2 < b->y2 ? a->y2 : b->y2;
  if (ix1 >= ix2 || iy1 >= iy2) return 0.0f;
  float inter = (ix2 - ix1) * (iy2 - iy1);
  float aa = (a->x2 - a->x1) * (a->y2 - a->y1);
  float ab = (b->x2 - b->x1) * (b->y2 - b->y1);
  return inter / (aa + ab - inter);
}

/* decode_head (postprocess.hpp:94-145): CHW fp32 logits, sigmoid-argmax (strict >, first
 * max wins, max_conf starts at 0), keep iff conf > thr, TLBR * stride around the cell
 * centre, optional dilation by q*w, q*h.  Row-major cell order.  Returns count. */
int uydo_decode_tlbr_cells(const float *cls, const float *reg, int w, int h, int stride, int nc,
                           float thr, float q, uydo_det *out, int *cells, int cap) {
  int n = 0, hw = w * h;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int best = -1; float mc = 0.0f;
      for (int c = 0; c < nc; ++c) {
        float p = 1.0f / (1.0f + expf(-cls[c * hw + y * w + x]));
        if (p > mc) { mc = p; best = c; }
      }
      if (mc > thr) {
        int g = y * w + x;
        float xc = (x + 0.5f) * stride, yc = (y + 0.5f) * stride;
        uydo_det d;
        d.x1 = xc - reg[0 * hw + g] * stride;
        d.y1 = yc - reg[1 * hw + g] * stride;
        d.x2 = xc + reg[2 * hw + g] * stride;
        d.y2 = yc + reg[3 * hw + g] * stride;
        d.conf = mc; d.cls = best;
        if (q > 0.0f) { /* postprocess.hpp:77-85 */
          float dw = (d.x2 - d.x1) * q, dh = (d.y2 - d.y1) * q;
          d.x1 -= dw; d.y1 -= dh; d.x2 += dw; d.y2 += dh;
        }
        if (n < cap) { out[n] = d; if (cells) cells[n] = g; }
        ++n;
      }
    }
  return n;
}

int uydo_decode_tlbr(const float *cls, const float *reg, int w, int h, int stride, int nc,
                     float thr, float q, uydo_det *out, int cap) {
  return uydo_decode_tlbr_cells(cls, reg, w, h, stride, nc, thr, q, out, 0, cap);
}

/* nms (postprocess.hpp:44-67) on detections ALREADY sorted by confidence descending
 * (the header uses an unstable std::sort; ties are the caller's business). */
int uydo_nms_hpp_sorted(const uydo_det *d, int n, float thr, int *keep_out) {
  unsigned char *sup = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    if (sup[i]) continue;
    keep_out[kept++] = i;
    for (int j = i + 1; j < n; ++j)
      if (!sup[j] && d[i].cls == d[j].cls && hpp_iou(&d[i], &d[j]) > thr) sup[j] = 1;
  }
  free(sup);
  return kept;
}
