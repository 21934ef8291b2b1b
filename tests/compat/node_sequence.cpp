// Drop-in test of libuyd_compat.so: this file #includes the REFERENCE's own headers (compiled where they lie under
// /root/reference, nothing copied), links libuyd_compat.so, and drives the call sequence of
// perception_node.cpp:470-483 (resources) and :604-660 (processGpuBuffer):
//   preprocess_bgra_resize -> [engine] -> reset_detection_counter -> decode_yolo_head x3 -> get_detection_count ->
//   cudaStreamSynchronize -> run_gpu_nms -> copy_valid_detections_to_host.
// Checked against (a) the reference's CPU header postprocess.hpp (decode_head + nms), included here, and (b) the
// reference's own pre-processing kernels in oracle/_ref/libref_preprocess.so (dlopen, same symbol names).
// Built by oracle/build.py into oracle/_ref/compat_node_test (a prebuilt binary on the GPU box). TEST INFRASTRUCTURE.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "cuda_preprocess.h"  // reference
#include "gpu_postprocess.h"  // reference
#include "postprocess.hpp"    // reference (CPU decode_head / nms: the oracle)

#include "uyd_compat.h"       // ours: every prototype must agree with the reference declarations above

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      std::printf("FAIL %s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

typedef cudaError_t (*resize_fn)(const uint8_t *, float *, int, int, int, int, int, NormParams, cudaStream_t);
typedef cudaError_t (*bgra_fn)(const uint8_t *, float *, int, int, int, NormParams, cudaStream_t);
typedef cudaError_t (*nv12_fn)(const uint8_t *, const uint8_t *, float *, int, int, int, int, NormParams, cudaStream_t);

static int check_preprocess(const char *ref_so, cudaStream_t stream) {
  void *h = dlopen(ref_so, RTLD_NOW | RTLD_LOCAL);
  if (!h) { std::printf("SKIP preprocess (no %s: %s)\n", ref_so, dlerror()); return 0; }
  resize_fn ref_resize = (resize_fn)dlsym(h, "preprocess_bgra_resize");
  bgra_fn ref_bgra = (bgra_fn)dlsym(h, "preprocess_bgra");
  nv12_fn ref_nv12 = (nv12_fn)dlsym(h, "preprocess_nv12");
  if (!ref_resize || !ref_bgra || !ref_nv12) { std::printf("FAIL reference preprocess symbols missing\n"); return 1; }
  const int sw = 1280, sh = 720, pitch = 1280 * 4 + 256 - (1280 * 4) % 256, dw = 640, dh = 640;  // 256-byte pitch (:590)
  std::mt19937 rng(5);
  std::vector<uint8_t> frame((size_t)pitch * sh);
  for (auto &v : frame) v = (uint8_t)(rng() & 0xFF);
  uint8_t *d_in = nullptr;
  CK(cudaMalloc(&d_in, frame.size()));
  CK(cudaMemcpy(d_in, frame.data(), frame.size(), cudaMemcpyHostToDevice));
  float *d_a = allocate_preprocess_buffer(std::max(dw, sw), std::max(dh, sh));
  float *d_b = allocate_preprocess_buffer(std::max(dw, sw), std::max(dh, sh));
  if (!d_a || !d_b) { std::printf("FAIL allocate_preprocess_buffer\n"); return 1; }
  NormParams np = create_norm_params_imagenet();
  auto compare = [&](const char *what, size_t n, float tol) -> int {
    std::vector<float> a(n), b(n);
    cudaStreamSynchronize(stream);
    cudaMemcpy(a.data(), d_a, n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), d_b, n * 4, cudaMemcpyDeviceToHost);
    float worst = 0.f;
    for (size_t i = 0; i < n; ++i) worst = std::max(worst, std::fabs(a[i] - b[i]));
    std::printf("%s %s: max |ours - reference kernel| = %.3g (tol %.1g)\n", worst <= tol ? "ok  " : "FAIL", what, worst, tol);
    return worst <= tol ? 0 : 1;
  };
  int bad = 0;
  CK(preprocess_bgra_resize(d_in, d_a, sw, sh, pitch, dw, dh, np, stream));
  CK(ref_resize(d_in, d_b, sw, sh, pitch, dw, dh, np, stream));
  bad += compare("preprocess_bgra_resize 1280x720 -> 640x640", (size_t)3 * dw * dh, 2e-6f);
  CK(preprocess_bgra(d_in, d_a, sw, sh, pitch, np, stream));
  CK(ref_bgra(d_in, d_b, sw, sh, pitch, np, stream));
  bad += compare("preprocess_bgra 1280x720", (size_t)3 * sw * sh, 0.f);
  CK(preprocess_nv12(d_in, d_in + (size_t)pitch * 480, d_a, 1280, 480, pitch, pitch, np, stream));
  CK(ref_nv12(d_in, d_in + (size_t)pitch * 480, d_b, 1280, 480, pitch, pitch, np, stream));
  bad += compare("preprocess_nv12 1280x480", (size_t)3 * 1280 * 480, 2e-6f);
  free_preprocess_buffer(d_a);
  free_preprocess_buffer(d_b);
  cudaFree(d_in);
  return bad;
}

int main(int argc, char **argv) {
  const std::string ref_pre = argc > 1 ? argv[1] : "oracle/_ref/libref_preprocess.so";
  cudaStream_t stream = create_preprocess_stream();
  if (!stream) { std::printf("FAIL create_preprocess_stream\n"); return 1; }
  int bad = check_preprocess(ref_pre.c_str(), stream);

  // ---- head tensors of the three levels (perception_node.cpp:612-618: p2/p3/p4 cls [4,H,W], reg [4,H,W]) ----
  const int nc = 4, strides[3] = {4, 8, 16}, gw[3] = {160, 80, 40};
  const float conf_thr = 0.5f, iou_thr = 0.45f;
  CK(init_postprocess_resources());
  GpuDetection *d_dets = nullptr;
  CK(cudaMalloc(&d_dets, MAX_DETECTIONS * sizeof(GpuDetection)));
  std::vector<GpuDetection> h_dets(MAX_DETECTIONS);
  for (int round = 0; round < 3; ++round) {  // 0: plain, 1: conformal dilation q = 0.1, 2: empty frame
    const float q = round == 1 ? 0.1f : 0.0f;
    std::mt19937 rng(17 + round);
    std::normal_distribution<float> logit(round == 2 ? -9.0f : -3.6f, 1.3f);
    std::uniform_real_distribution<float> dist(0.5f, 6.0f);
    std::vector<float> cls[3], reg[3];
    float *d_cls[3], *d_reg[3];
    std::vector<Detection> want;
    for (int l = 0; l < 3; ++l) {
      const size_t hw = (size_t)gw[l] * gw[l];
      cls[l].resize(nc * hw);
      reg[l].resize(4 * hw);
      for (auto &v : cls[l]) v = logit(rng);
      for (auto &v : reg[l]) v = dist(rng);
      CK(cudaMalloc(&d_cls[l], cls[l].size() * 4));
      CK(cudaMalloc(&d_reg[l], reg[l].size() * 4));
      CK(cudaMemcpy(d_cls[l], cls[l].data(), cls[l].size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_reg[l], reg[l].data(), reg[l].size() * 4, cudaMemcpyHostToDevice));
      decode_head(cls[l].data(), reg[l].data(), gw[l], gw[l], strides[l], nc, conf_thr, q, want);  // reference CPU decode
    }
    const size_t n_decoded = want.size();
    std::vector<Detection> want_kept = nms(want, iou_thr);                                        // reference CPU NMS

    // ---- the node's sequence, verbatim order ----
    CK(reset_detection_counter(stream));
    for (int l = 0; l < 3; ++l)
      CK(decode_yolo_head(d_cls[l], d_reg[l], d_dets, gw[l], gw[l], strides[l], nc, conf_thr, q, stream));
    int num = 0;
    CK(get_detection_count(&num, stream));
    CK(cudaStreamSynchronize(stream));
    int valid = 0;
    if (num > 0) {
      num = std::min(num, (int)MAX_DETECTIONS);
      CK(run_gpu_nms(d_dets, num, iou_thr, stream));
      CK(copy_valid_detections_to_host(d_dets, h_dets.data(), num, &valid, stream));
    }
    // ---- compare ----
    int mism = 0;
    if ((size_t)num != n_decoded) { std::printf("FAIL round %d: decoded %d cells, reference header %zu\n", round, num, n_decoded); ++mism; }
    if ((size_t)valid != want_kept.size()) { std::printf("FAIL round %d: kept %d, reference header %zu\n", round, valid, want_kept.size()); ++mism; }
    for (int i = 0; i < valid && (size_t)i < want_kept.size() && mism < 5; ++i) {
      const GpuDetection &g = h_dets[i];
      const Detection &w = want_kept[i];
      const bool same = g.x1 == w.x1 && g.y1 == w.y1 && g.x2 == w.x2 && g.y2 == w.y2 && g.class_id == w.class_id &&
                        std::fabs(g.confidence - w.confidence) <= 5e-7f && g.valid == 1;
      if (!same) {
        std::printf("FAIL round %d row %d: got (%.9g %.9g %.9g %.9g %.9g %d) want (%.9g %.9g %.9g %.9g %.9g %d)\n", round, i, g.x1, g.y1,
                    g.x2, g.y2, g.confidence, g.class_id, w.x1, w.y1, w.x2, w.y2, w.confidence, w.class_id);
        ++mism;
      }
    }
    // the buffer run_gpu_nms leaves behind: confidence-descending, valid flags = survivors
    if (num > 0) {
      std::vector<GpuDetection> all(num);
      CK(cudaMemcpy(all.data(), d_dets, (size_t)num * sizeof(GpuDetection), cudaMemcpyDeviceToHost));
      int nvalid = 0;
      for (int i = 0; i < num; ++i) {
        nvalid += all[i].valid != 0;
        if (i && all[i - 1].confidence < all[i].confidence) { std::printf("FAIL round %d: buffer not sorted at %d\n", round, i); ++mism; break; }
      }
      if (nvalid != valid) { std::printf("FAIL round %d: %d valid flags vs %d copied\n", round, nvalid, valid); ++mism; }
    }
    std::printf("%s round %d (q = %.1f): %d decoded, %d kept; reference header %zu / %zu\n", mism ? "FAIL" : "ok  ", round, q, num, valid,
                n_decoded, want_kept.size());
    bad += mism;
    for (int l = 0; l < 3; ++l) { cudaFree(d_cls[l]); cudaFree(d_reg[l]); }
  }
  cudaFree(d_dets);
  CK(cleanup_postprocess_resources());
  destroy_preprocess_stream(stream);
  std::printf(bad ? "FAIL (%d)\n" : "PASS\n", bad);
  return bad ? 1 : 0;
}
