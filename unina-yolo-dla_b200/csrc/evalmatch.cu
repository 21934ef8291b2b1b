// Batched evaluation consumer of the [N, 6] detections, the step right after
// the hot path and the reason its NCCL gather exists:
//   * small-object TP / FP / FN  -- UninaValidator.update_metrics, trainer.py:210-265 (pixel xyxy boxes,
//     "small" = width and height below size_thr, match = same class and Ultralytics box_iou > iou_thr;
//     TP = small ground truths hit by ANY prediction, FP = small predictions hitting no small ground truth)
//   * conformal nonconformity scores 1 - IoU of greedily matched pairs -- calibrate_conformal_prediction,
//     train.py:335-470 (predictions in confidence order, best unmatched same-class ground truth with
//     IoU >= match_thr, ties -> first ground truth).
// One CTA per image; detections arrive in confidence order (the NMS emits them that way).
#include "common.cuh"

namespace uyd {
namespace {

// Ultralytics metrics.box_iou: inter / (area1 + area2 - inter + eps), eps = 1e-7, fp32
__device__ __forceinline__ float iou_ultra(const float4 &a, const float4 &b) {
  const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
  const float inter = __fmul_rn(w, h);
  const float a1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y)), a2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  return __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(a1, a2), inter), 1e-7f));
}

// train.py:335-350: early-out on empty intersection, no epsilon
__device__ __forceinline__ float iou_plain(const float4 &a, const float4 &b) {
  const float x1 = fmaxf(a.x, b.x), y1 = fmaxf(a.y, b.y), x2 = fminf(a.z, b.z), y2 = fminf(a.w, b.w);
  if (x2 <= x1 || y2 <= y1) return 0.f;
  const float inter = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
  const float uni = __fsub_rn(__fadd_rn(__fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y)), __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y))), inter);
  return uni > 0.f ? __fdiv_rn(inter, uni) : 0.f;
}

constexpr int kThreads = 128, kMaxGt = 1024, kMaxDet = 1024;

__global__ void __launch_bounds__(kThreads) eval_update_kernel(const float *__restrict__ det, const int *__restrict__ cnt, int max_det,
                                                               const float *__restrict__ gt, const int *__restrict__ gt_cnt, int gmax,
                                                               float size_thr, float small_iou_thr, float match_iou_thr,
                                                               unsigned long long *__restrict__ counters, float *__restrict__ scores) {
  __shared__ float4 gbox[kMaxGt];
  __shared__ int gcls[kMaxGt];
  __shared__ unsigned char gsmall[kMaxGt], gused[kMaxGt];
  __shared__ int s_small_gt, s_tp, s_fp;
  __shared__ float w_iou[kThreads / 32];
  __shared__ int w_idx[kThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = min(cnt[b], max_det), g = min(gt_cnt[b], min(gmax, kMaxGt));
  const float *dimg = det + (long long)b * max_det * 6;
  if (tid == 0) s_small_gt = s_tp = s_fp = 0;
  __syncthreads();
  int local_small = 0;
  for (int j = tid; j < g; j += kThreads) {
    const float *q = gt + ((long long)b * gmax + j) * 5;
    gbox[j] = make_float4(q[1], q[2], q[3], q[4]);
    gcls[j] = (int)q[0];
    const bool sm = __fsub_rn(q[3], q[1]) < size_thr && __fsub_rn(q[4], q[2]) < size_thr;
    gsmall[j] = sm;
    gused[j] = 0;
    local_small += sm;
  }
  if (local_small) atomicAdd(&s_small_gt, local_small);
  __syncthreads();
  // ---- small-object counters (skipped entirely when the image has no small ground truth, trainer.py:231) ----
  if (s_small_gt > 0) {
    int tp = 0, fp = 0;
    for (int j = tid; j < g; j += kThreads) {
      if (!gsmall[j]) continue;
      bool hit = false;
      for (int i = 0; i < n && !hit; ++i) {
        const float *d = dimg + i * 6;
        hit = (int)d[5] == gcls[j] && iou_ultra(make_float4(d[0], d[1], d[2], d[3]), gbox[j]) > small_iou_thr;
      }
      tp += hit;
    }
    for (int i = tid; i < n; i += kThreads) {
      const float *d = dimg + i * 6;
      const float4 pb = make_float4(d[0], d[1], d[2], d[3]);
      if (!(__fsub_rn(pb.z, pb.x) < size_thr && __fsub_rn(pb.w, pb.y) < size_thr)) continue;
      bool hit = false;
      for (int j = 0; j < g && !hit; ++j) hit = gsmall[j] && (int)d[5] == gcls[j] && iou_ultra(pb, gbox[j]) > small_iou_thr;
      fp += !hit;
    }
    if (tp) atomicAdd(&s_tp, tp);
    if (fp) atomicAdd(&s_fp, fp);
  }
  __syncthreads();
  if (tid == 0 && s_small_gt > 0) {
    atomicAdd(&counters[0], (unsigned long long)s_tp);
    atomicAdd(&counters[1], (unsigned long long)s_fp);
    atomicAdd(&counters[2], (unsigned long long)(s_small_gt - s_tp));
  }
  // ---- conformal scores: greedy matching in confidence order ----
  if (!scores) return;
  for (int i = 0; i < max_det; ++i) {
    float best = 0.f;
    int best_j = 0x7fffffff;
    if (i < n) {
      const float *d = dimg + i * 6;
      const float4 pb = make_float4(d[0], d[1], d[2], d[3]);
      const int pc = (int)d[5];
      for (int j = tid; j < g; j += kThreads) {
        if (gused[j] || gcls[j] != pc) continue;
        const float v = iou_plain(pb, gbox[j]);
        if (v > best && v >= match_iou_thr) { best = v; best_j = j; }  // strict >: the first index wins a tie inside a thread
      }
    }
    // block argmax, ties -> lowest ground-truth index
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
      if (ov > best || (ov == best && oj < best_j)) { best = ov; best_j = oj; }
    }
    if ((tid & 31) == 0) { w_iou[tid >> 5] = best; w_idx[tid >> 5] = best_j; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kThreads / 32; ++w)
        if (w_iou[w] > best || (w_iou[w] == best && w_idx[w] < best_j)) { best = w_iou[w]; best_j = w_idx[w]; }
      const bool ok = i < n && best_j != 0x7fffffff && best > 0.f;
      if (ok) gused[best_j] = 1;
      scores[(long long)b * max_det + i] = ok ? __fsub_rn(1.0f, best) : -1.0f;
    }
    __syncthreads();
  }
}


// ---- data_loader.SmallObjectMetric.update (data_loader.py:322-389), one CTA per image ------------------------
// Rows in the reference's own format: pred (x_c, y_c, w, h, conf, cls) and gt (cls, x_c, y_c, w, h), normalised to the
// image; predictions arrive in confidence order.  A "small" box has w * image_size and h * image_size below size_thr
// (double arithmetic, like the Python floats of _is_small).  Per prediction: the best IoU (> 0, strict: the first
// ground truth wins a tie) over the unmatched small ground truths of its class; IoU >= iou_thr -> TP and the ground
// truth is matched, else FP iff the prediction itself is small.  FN = small ground truths left unmatched.
__device__ __forceinline__ float iou_cxcywh(const float4 &a, const float4 &b) {  // data_loader.py:288-320, fp32 like the tensors
  const float ax1 = __fsub_rn(a.x, __fmul_rn(a.z, 0.5f)), ay1 = __fsub_rn(a.y, __fmul_rn(a.w, 0.5f));
  const float ax2 = __fadd_rn(a.x, __fmul_rn(a.z, 0.5f)), ay2 = __fadd_rn(a.y, __fmul_rn(a.w, 0.5f));
  const float bx1 = __fsub_rn(b.x, __fmul_rn(b.z, 0.5f)), by1 = __fsub_rn(b.y, __fmul_rn(b.w, 0.5f));
  const float bx2 = __fadd_rn(b.x, __fmul_rn(b.z, 0.5f)), by2 = __fadd_rn(b.y, __fmul_rn(b.w, 0.5f));
  const float iw = fmaxf(0.f, __fsub_rn(fminf(ax2, bx2), fmaxf(ax1, bx1))), ih = fmaxf(0.f, __fsub_rn(fminf(ay2, by2), fmaxf(ay1, by1)));
  const float inter = __fmul_rn(iw, ih);
  const float uni = __fsub_rn(__fadd_rn(__fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1)), __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1))), inter);
  return uni <= 0.f ? 0.f : __fdiv_rn(inter, uni);
}

__global__ void __launch_bounds__(kThreads) small_object_metric_kernel(const float *__restrict__ pred, const int *__restrict__ cnt,
                                                                       int max_det, const float *__restrict__ gt,
                                                                       const int *__restrict__ gt_cnt, int gmax, double size_thr,
                                                                       double iou_thr, double image_size,
                                                                       unsigned long long *__restrict__ counters) {
  __shared__ float4 gbox[kMaxGt];
  __shared__ int gcls[kMaxGt];
  __shared__ unsigned char gused[kMaxGt];
  __shared__ int s_n;
  __shared__ float w_iou[kThreads / 32];
  __shared__ int w_idx[kThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = min(cnt[b], max_det), g = min(gt_cnt[b], min(gmax, kMaxGt));
  if (tid == 0) {  // ordered compaction of the small ground truths (the order decides IoU ties)
    int k = 0;
    for (int j = 0; j < g; ++j) {
      const float *q = gt + ((long long)b * gmax + j) * 5;
      if ((double)q[3] * image_size < size_thr && (double)q[4] * image_size < size_thr) {
        gbox[k] = make_float4(q[1], q[2], q[3], q[4]);
        gcls[k] = (int)q[0];
        gused[k] = 0;
        ++k;
      }
    }
    s_n = k;
  }
  __syncthreads();
  const int ns = s_n;
  if (ns == 0) return;  // no small objects in this image (data_loader.py:337-338)
  int tp = 0, fp = 0;
  for (int i = 0; i < n; ++i) {
    const float *d = pred + ((long long)b * max_det + i) * 6;
    const float4 pb = make_float4(d[0], d[1], d[2], d[3]);
    const int pc = (int)d[5];
    float best = 0.f;
    int best_j = 0x7fffffff;
    for (int j = tid; j < ns; j += kThreads) {
      if (gused[j] || gcls[j] != pc) continue;
      const float v = iou_cxcywh(pb, gbox[j]);
      if (v > best) { best = v; best_j = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
      if (ov > best || (ov == best && oj < best_j)) { best = ov; best_j = oj; }
    }
    if ((tid & 31) == 0) { w_iou[tid >> 5] = best; w_idx[tid >> 5] = best_j; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kThreads / 32; ++w)
        if (w_iou[w] > best || (w_iou[w] == best && w_idx[w] < best_j)) { best = w_iou[w]; best_j = w_idx[w]; }
      if ((double)best >= iou_thr) {
        ++tp;
        if (best_j != 0x7fffffff) gused[best_j] = 1;   // (iou_thr <= 0 with no candidate: the reference adds index -1)
      } else if ((double)pb.z * image_size < size_thr && (double)pb.w * image_size < size_thr) {
        ++fp;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    int matched = 0;
    for (int j = 0; j < ns; ++j) matched += gused[j];
    atomicAdd(&counters[0], (unsigned long long)tp);
    atomicAdd(&counters[1], (unsigned long long)fp);
    atomicAdd(&counters[2], (unsigned long long)(ns - matched));
  }
}

}  // namespace
}  // namespace uyd

extern "C" int uyd_small_object_metric_update(uyd_ctx *ctx, const float *pred, const int *count, int batch, int max_det, const float *gt,
                                              const int *gt_count, int gt_max, double size_thr, double iou_thr, double image_size,
                                              unsigned long long *counters, uyd_stream stream) {
  UYD_REQUIRE(pred && count && gt && gt_count && counters && batch > 0 && max_det > 0 && gt_max > 0, UYD_E_ARG,
              "uyd_small_object_metric_update: bad arguments");
  uyd::DeviceGuard guard(uyd::ctx_device(ctx));
  UYD_REQUIRE(gt_max <= uyd::kMaxGt, UYD_E_UNSUPPORTED, "uyd_small_object_metric_update: at most %d ground truths per image", uyd::kMaxGt);
  uyd::small_object_metric_kernel<<<batch, uyd::kThreads, 0, (cudaStream_t)stream>>>(pred, count, max_det, gt, gt_count, gt_max, size_thr,
                                                                                      iou_thr, image_size, counters);
  return (int)cudaGetLastError();
}

extern "C" int uyd_eval_update(uyd_ctx *ctx, const float *det, const int *count, int batch, int max_det, const float *gt,
                               const int *gt_count, int gt_max, float size_thr, float small_iou_thr, float match_iou_thr,
                               unsigned long long *counters, float *scores, uyd_stream stream) {
  UYD_REQUIRE(det && count && gt && gt_count && counters && batch > 0 && max_det > 0 && gt_max > 0, UYD_E_ARG,
              "uyd_eval_update: bad arguments");
  uyd::DeviceGuard guard(uyd::ctx_device(ctx));
  UYD_REQUIRE(gt_max <= uyd::kMaxGt && max_det <= uyd::kMaxDet, UYD_E_UNSUPPORTED, "uyd_eval_update: at most %d ground truths / %d detections per image",
              uyd::kMaxGt, uyd::kMaxDet);
  uyd::eval_update_kernel<<<batch, uyd::kThreads, 0, (cudaStream_t)stream>>>(det, count, max_det, gt, gt_count, gt_max, size_thr,
                                                                              small_iou_thr, match_iou_thr, counters, scores);
  return (int)cudaGetLastError();
}
