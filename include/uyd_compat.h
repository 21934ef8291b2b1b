/* libuyd_compat.so -- the reference's OWN extern "C" symbols for this path, implemented on libuyd.so.
 *
 * A maintainer of the reference links libuyd_compat.so instead of compiling
 * ros2_ws/src/perception/src/gpu_postprocess.cu and cuda_preprocess.cu; perception_node.cpp
 * (processGpuBuffer, :581-689; resources :470-483,696-720) builds and runs unchanged.  Every prototype below is
 * the reference's, name for name and argument for argument:
 *
 *   gpu_postprocess.h:45   init_postprocess_resources        gpu_postprocess.h:50   cleanup_postprocess_resources
 *   gpu_postprocess.h:55   reset_detection_counter           gpu_postprocess.h:60   get_detection_count
 *   gpu_postprocess.h:65   decode_yolo_head                  gpu_postprocess.h:72   run_gpu_nms
 *   gpu_postprocess.h:78   copy_valid_detections_to_host
 *   cuda_preprocess.h:50   create_norm_params_imagenet       cuda_preprocess.h:55   create_norm_params
 *   cuda_preprocess.h:63   preprocess_bgra_resize            cuda_preprocess.h:71   preprocess_bgra
 *   cuda_preprocess.h:78   preprocess_nv12
 *   cuda_preprocess.h:98   allocate_preprocess_buffer        cuda_preprocess.h:103  free_preprocess_buffer
 *   cuda_preprocess.h:108  create_preprocess_stream          cuda_preprocess.h:113  destroy_preprocess_stream
 *
 * This header may be included AFTER the reference's headers (tests/compat/node_sequence.cpp does): the struct
 * definitions are skipped and the compiler checks that every prototype agrees with the reference's declaration.
 *
 * Behaviour that differs from gpu_postprocess.cu on purpose:
 *   - run_gpu_nms is the exact, deterministic greedy NMS of the reference's CPU header (postprocess.hpp:28-67)
 *     instead of the racy nms_kernel with its +1e-6 denominator (gpu_postprocess.cu:82,213-228); equal
 *     confidences are ordered by grid cell.  The buffer is left as the reference leaves it: sorted by confidence,
 *     valid = 1 / 0.
 *   - the workspace singleton (gpu_postprocess.cu:56) belongs to the device that was current in
 *     init_postprocess_resources.
 */
#ifndef UYD_COMPAT_H
#define UYD_COMPAT_H

#include <stdint.h>

#include <cuda_runtime_api.h>

#ifndef GPU_POSTPROCESS_H /* the reference header defines the same types */
#define MAX_DETECTIONS 1024
#ifdef __cplusplus
struct alignas(32) GpuDetection {
#else
struct GpuDetection {
#endif
  float x1, y1, x2, y2;
  float confidence;
  int class_id;
  int valid;
  int _pad;
};
#endif

#ifdef __cplusplus
extern "C" {
#endif

#ifndef CUDA_PREPROCESS_H
typedef struct {
  float mean_r;
  float mean_g;
  float mean_b;
  float std_r;
  float std_g;
  float std_b;
} NormParams;
#endif

cudaError_t init_postprocess_resources();
cudaError_t cleanup_postprocess_resources();
cudaError_t reset_detection_counter(cudaStream_t stream);
cudaError_t get_detection_count(int *count, cudaStream_t stream);
cudaError_t decode_yolo_head(const float *d_cls, const float *d_reg, struct GpuDetection *d_detections, int grid_w, int grid_h,
                             int stride, int num_classes, float conf_threshold, float conformal_q, cudaStream_t stream);
cudaError_t run_gpu_nms(struct GpuDetection *d_detections, int num_detections, float iou_threshold, cudaStream_t stream);
cudaError_t copy_valid_detections_to_host(const struct GpuDetection *d_detections, struct GpuDetection *h_detections,
                                          int num_detections, int *out_valid_count, cudaStream_t stream);

NormParams create_norm_params_imagenet(void);
NormParams create_norm_params(float mean_r, float mean_g, float mean_b, float std_r, float std_g, float std_b);
cudaError_t preprocess_bgra_resize(const uint8_t *d_input, float *d_output, int src_width, int src_height, int src_pitch,
                                   int dst_width, int dst_height, NormParams params, cudaStream_t stream);
cudaError_t preprocess_bgra(const uint8_t *d_input, float *d_output, int width, int height, int pitch, NormParams params,
                            cudaStream_t stream);
cudaError_t preprocess_nv12(const uint8_t *d_y_plane, const uint8_t *d_uv_plane, float *d_output, int width, int height,
                            int y_pitch, int uv_pitch, NormParams params, cudaStream_t stream);
float *allocate_preprocess_buffer(int width, int height);
void free_preprocess_buffer(float *d_buffer);
cudaStream_t create_preprocess_stream(void);
void destroy_preprocess_stream(cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UYD_COMPAT_H */
