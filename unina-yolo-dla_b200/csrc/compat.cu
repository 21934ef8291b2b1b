// libuyd_compat.so: the reference's own extern "C" entry points (include/uyd_compat.h) on top of libuyd.so.
// Host code only -- every kernel lives in libuyd.so (decode.cu, nms.cu, preprocess.cu).
#include <cstdio>

#include "../../include/uyd.h"
#include "../../include/uyd_compat.h"

static_assert(sizeof(GpuDetection) == sizeof(uyd_detection) && sizeof(GpuDetection) == 32, "GpuDetection layout");
static_assert(sizeof(NormParams) == sizeof(uyd_norm_params), "NormParams layout");

namespace {

// What gpu_postprocess.cu:43-57 keeps in its global PostprocessWorkspace, plus the per-slot grid cell that makes
// the NMS order deterministic.
struct Workspace {
  uyd_ctx *ctx = nullptr;
  int *d_count = nullptr;            // atomic slot counter of decode_yolo_head
  int *d_cell = nullptr;             // [MAX_DETECTIONS] grid cell of every slot
  int *d_kept = nullptr;             // survivors of the last run_gpu_nms / compaction
  uyd_detection *d_compact = nullptr;  // [MAX_DETECTIONS]
  const void *cell_owner = nullptr;  // detection buffer the cells in d_cell belong to
} g;

inline cudaError_t as_cuda(int code) {
  if (code == 0) return cudaSuccess;
  if (code >= 10000) {  // UYD_E_*: argument / state errors have no cudaError_t of their own
    std::fprintf(stderr, "libuyd_compat: %s\n", uyd_last_error());
    return cudaErrorInvalidValue;
  }
  return (cudaError_t)code;
}

inline uyd_norm_params as_uyd(const NormParams &p) {
  uyd_norm_params q = {p.mean_r, p.mean_g, p.mean_b, p.std_r, p.std_g, p.std_b};
  return q;
}

}  // namespace

extern "C" {

cudaError_t init_postprocess_resources() {
  if (g.ctx) return cudaSuccess;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (int c = uyd_create(dev, &g.ctx)) return as_cuda(c);
  if ((e = cudaMalloc(&g.d_count, sizeof(int))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&g.d_cell, MAX_DETECTIONS * sizeof(int))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&g.d_kept, sizeof(int))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&g.d_compact, MAX_DETECTIONS * sizeof(uyd_detection))) != cudaSuccess) return e;
  return cudaMemset(g.d_count, 0, sizeof(int));
}

cudaError_t cleanup_postprocess_resources() {
  if (g.d_count) cudaFree(g.d_count);
  if (g.d_cell) cudaFree(g.d_cell);
  if (g.d_kept) cudaFree(g.d_kept);
  if (g.d_compact) cudaFree(g.d_compact);
  if (g.ctx) uyd_destroy(g.ctx);
  g = Workspace();
  return cudaSuccess;
}

cudaError_t reset_detection_counter(cudaStream_t stream) {
  if (!g.ctx) return cudaErrorInitializationError;
  g.cell_owner = nullptr;
  return cudaMemsetAsync(g.d_count, 0, sizeof(int), stream);
}

cudaError_t get_detection_count(int *count, cudaStream_t stream) {
  if (!g.ctx) return cudaErrorInitializationError;
  return cudaMemcpyAsync(count, g.d_count, sizeof(int), cudaMemcpyDeviceToHost, stream);
}

cudaError_t decode_yolo_head(const float *d_cls, const float *d_reg, GpuDetection *d_detections, int grid_w, int grid_h,
                             int stride, int num_classes, float conf_threshold, float conformal_q, cudaStream_t stream) {
  if (!g.ctx) return cudaErrorInitializationError;
  g.cell_owner = d_detections;
  // strict = 0: keep a cell iff conf >= threshold (gpu_postprocess.cu:132); levels get disjoint cell ranges
  return as_cuda(uyd_decode_tlbr(g.ctx, d_cls, d_reg, reinterpret_cast<uyd_detection *>(d_detections), g.d_cell, g.d_count,
                                 MAX_DETECTIONS, grid_w, grid_h, stride, num_classes, conf_threshold, conformal_q, /*strict=*/0,
                                 /*cell_base=*/stride << 20, stream));
}

cudaError_t run_gpu_nms(GpuDetection *d_detections, int num_detections, float iou_threshold, cudaStream_t stream) {
  if (num_detections == 0) return cudaSuccess;
  if (!g.ctx) return cudaErrorInitializationError;
  if (num_detections < 0 || num_detections > MAX_DETECTIONS) return cudaErrorInvalidValue;
  const int *cell = g.cell_owner == d_detections ? g.d_cell : nullptr;
  g.cell_owner = nullptr;  // the sort permutes the slots: the cells are consumed
  return as_cuda(uyd_nms_detections_inplace(g.ctx, reinterpret_cast<uyd_detection *>(d_detections), cell, num_detections,
                                            iou_threshold, g.d_kept, stream));
}

cudaError_t copy_valid_detections_to_host(const GpuDetection *d_detections, GpuDetection *h_detections, int num_detections,
                                          int *out_valid_count, cudaStream_t stream) {
  if (num_detections == 0) {
    *out_valid_count = 0;
    return cudaSuccess;
  }
  if (!g.ctx) return cudaErrorInitializationError;
  if (num_detections < 0 || num_detections > MAX_DETECTIONS) return cudaErrorInvalidValue;
  if (int c = uyd_compact_valid(g.ctx, reinterpret_cast<const uyd_detection *>(d_detections), num_detections, g.d_compact, g.d_kept,
                                stream))
    return as_cuda(c);
  int valid = 0;
  cudaError_t e = cudaMemcpyAsync(&valid, g.d_kept, sizeof(int), cudaMemcpyDeviceToHost, stream);
  if (e != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
  *out_valid_count = valid;
  if (valid > 0) {
    e = cudaMemcpyAsync(h_detections, g.d_compact, (size_t)valid * sizeof(GpuDetection), cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(stream);  // the node reads h_detections right after the call (perception_node.cpp:657-660)
  }
  return e;
}

NormParams create_norm_params_imagenet(void) { return create_norm_params(0.485f, 0.456f, 0.406f, 0.229f, 0.224f, 0.225f); }

NormParams create_norm_params(float mean_r, float mean_g, float mean_b, float std_r, float std_g, float std_b) {
  NormParams p = {mean_r, mean_g, mean_b, std_r, std_g, std_b};
  return p;
}

cudaError_t preprocess_bgra_resize(const uint8_t *d_input, float *d_output, int src_width, int src_height, int src_pitch,
                                   int dst_width, int dst_height, NormParams params, cudaStream_t stream) {
  return as_cuda(uyd_preprocess_bgra_resize(d_input, d_output, src_width, src_height, src_pitch, dst_width, dst_height,
                                            as_uyd(params), stream));
}

cudaError_t preprocess_bgra(const uint8_t *d_input, float *d_output, int width, int height, int pitch, NormParams params,
                            cudaStream_t stream) {
  return as_cuda(uyd_preprocess_bgra(d_input, d_output, width, height, pitch, as_uyd(params), stream));
}

cudaError_t preprocess_nv12(const uint8_t *d_y_plane, const uint8_t *d_uv_plane, float *d_output, int width, int height,
                            int y_pitch, int uv_pitch, NormParams params, cudaStream_t stream) {
  return as_cuda(uyd_preprocess_nv12(d_y_plane, d_uv_plane, d_output, width, height, y_pitch, uv_pitch, as_uyd(params), stream));
}

float *allocate_preprocess_buffer(int width, int height) {
  float *p = nullptr;
  const cudaError_t e = cudaMalloc(&p, (size_t)3 * width * height * sizeof(float));
  if (e != cudaSuccess) {
    std::fprintf(stderr, "Failed to allocate preprocess buffer: %s\n", cudaGetErrorString(e));
    return nullptr;
  }
  return p;
}

void free_preprocess_buffer(float *d_buffer) {
  if (d_buffer) cudaFree(d_buffer);
}

cudaStream_t create_preprocess_stream(void) {
  cudaStream_t s = nullptr;
  if (cudaStreamCreate(&s) != cudaSuccess) return nullptr;  // a blocking stream, like cuda_preprocess.cu:419-428
  return s;
}

void destroy_preprocess_stream(cudaStream_t stream) {
  if (stream) cudaStreamDestroy(stream);
}

}  // extern "C"
