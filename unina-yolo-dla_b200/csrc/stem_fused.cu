// Fused stem: model.0 Conv(3,16,3,2) + model.1 Conv(16,32,3,2) (unina-yolo-dla-m.yaml:24-25, both
// Conv+BN+ReLU, BN folded) in one launch.  Unfused, the 320x320x16 tensor between them costs 6.6 MB of
// HBM traffic per frame and the 16-channel layer cannot feed tcgen05 efficiently (32-byte TMA rows).
//
// One CTA = 8 x 16 outputs of layer 1:
//   * the 35 x 68 x 3 frame patch is staged in shared memory as two bf16 planes hi + lo (hi + lo == the
//     fp32 value up to 2^-17, split once per element; uint8 frames: x / 255 first, the reference
//     predictor's pre-process), so the frame itself is not rounded (weights are bf16 as everywhere);
//   * layer 0 over the 17 x 33 region runs on tensor cores (mma.sync m16n8k16, bf16 x bf16 -> fp32):
//     K = (ci, ky) x 4 column slots kx = -1..2 (slot -1 has zero weights and makes every slot pair a
//     4-byte aligned shared-memory word);
//   * its bf16 result (zero outside the image = layer 1's padding) never leaves shared memory;
//   * layer 1 is nine k-steps of mma.sync (one filter tap = 16 channels per step).
#include <type_traits>

#include "stem_v2.cuh"

#include "common.cuh"

namespace uyd {

struct StemArgs {
  const void *in;      // NCHW frames, fp32 or uint8
  __nv_bfloat16 *out;  // NHWC slice base of image 0
  const uint32_t *wfrag;  // [L0: 3 k-steps x 2 n-tiles | L1: 9 k-steps x 4 n-tiles] x 32 lanes x 2 words
  const float *bias;      // [16 | 32]
  int n, ih, iw, oh, ow, out_pitch, u8;
  int pw;  // 1: the 1x1 Conv(32,16) follows in the same launch (second-generation kernel only); out has 16 channels
  // camera frames straight into the patch (no CHW fp32 tensor in HBM): cam = 1 packed BGRA (bilinear
  // resize when the frame extent differs from ih x iw), 2 = NV12; `in` = BGRA pixels / Y plane
  int cam, src_w, src_h, src_pitch, uv_pitch;
  long long frame_stride, uv_frame_stride;
  const uint8_t *uv;
  float mean[3], stdv[3];  // r, g, b
};

namespace {

constexpr int kTH = 8, kTW = 16;            // layer-1 output tile
constexpr int kL0H = 2 * kTH + 1, kL0W = 2 * kTW + 1;   // 17 x 33 layer-0 region
constexpr int kInH = 2 * kL0H + 1, kInW = 2 * kL0W + 2;  // 35 rows x 68 columns (frame columns 4*ox0 - 4 ...; column 0 only feeds the zero slot)
constexpr int kL0Pitch = 20;                // bf16 per layer-0 pixel in shared memory (16 + 4: conflict-free stride-2 reads)
constexpr int kL0Px = kL0H * kL0W;          // 561
constexpr int kThreads = 256;

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&h);
}
// v -> (hi, lo) bf16 pairs with hi + lo == v up to 2^-17 relative
__device__ __forceinline__ void split_bf16(float2 v, uint32_t &hi, uint32_t &lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  const float2 hf = __bfloat1622float2(h);
  hi = *reinterpret_cast<uint32_t *>(&h);
  lo = pack_bf16(v.x - hf.x, v.y - hf.y);
}

struct CamIn {};  // TIn tag of the camera instantiations of stem_v2_kernel

// One model-input pixel (iy, ix) from the camera frame, normalised like cuda_preprocess.cu:99-253:
// ((v / 255) - mean) / std with IEEE divisions; BGRA bilinear taps and BT.601 NV12 in the reference's operand order.
__device__ __forceinline__ float cam_norm(float v, float mean, float stdv) { return __fdiv_rn(__fsub_rn(__fdiv_rn(v, 255.0f), mean), stdv); }

__device__ __forceinline__ void cam_pixel(const StemArgs &a, const uint8_t *frame, const uint8_t *uvp, int iy, int ix, float (&rgb)[3]) {
  float r, g, b;
  if (a.cam == 2) {
    const float Y = frame[(long long)iy * a.src_pitch + ix];
    const uint8_t *q = uvp + (long long)(iy / 2) * a.uv_pitch + (ix / 2) * 2;
    const float U = q[0] - 128.0f, V = q[1] - 128.0f;
    r = Y + 1.402f * V;
    g = Y - 0.344136f * U - 0.714136f * V;
    b = Y + 1.772f * U;
    r = fmaxf(0.0f, fminf(255.0f, r)); g = fmaxf(0.0f, fminf(255.0f, g)); b = fmaxf(0.0f, fminf(255.0f, b));
  } else if (a.src_w == a.iw && a.src_h == a.ih) {
    const uint32_t px = *reinterpret_cast<const uint32_t *>(frame + (long long)iy * a.src_pitch + 4 * ix);
    b = (float)(px & 0xFF); g = (float)((px >> 8) & 0xFF); r = (float)((px >> 16) & 0xFF);
  } else {
    const float ratio_x = (float)a.src_w / a.iw, ratio_y = (float)a.src_h / a.ih;
    const float sx = fmaxf(0.0f, fminf((ix + 0.5f) * ratio_x - 0.5f, a.src_w - 1.0f));
    const float sy = fmaxf(0.0f, fminf((iy + 0.5f) * ratio_y - 0.5f, a.src_h - 1.0f));
    const int xa = (int)sx, ya = (int)sy, xb = min(xa + 1, a.src_w - 1), yb = min(ya + 1, a.src_h - 1);
    const float fx = sx - xa, fy = sy - ya;
    const float waa = (1.0f - fx) * (1.0f - fy), wab = fx * (1.0f - fy), wba = (1.0f - fx) * fy, wbb = fx * fy;
    const uint8_t *ra = frame + (long long)ya * a.src_pitch, *rb = frame + (long long)yb * a.src_pitch;
    const uint32_t paa = *reinterpret_cast<const uint32_t *>(ra + 4 * xa), pab = *reinterpret_cast<const uint32_t *>(ra + 4 * xb);
    const uint32_t pba = *reinterpret_cast<const uint32_t *>(rb + 4 * xa), pbb = *reinterpret_cast<const uint32_t *>(rb + 4 * xb);
    auto mix = [&](int sh) {
      return waa * (float)((paa >> sh) & 0xFF) + wab * (float)((pab >> sh) & 0xFF) + wba * (float)((pba >> sh) & 0xFF) +
             wbb * (float)((pbb >> sh) & 0xFF);
    };
    r = mix(16); g = mix(8); b = mix(0);
  }
  rgb[0] = cam_norm(r, a.mean[0], a.stdv[0]);
  rgb[1] = cam_norm(g, a.mean[1], a.stdv[1]);
  rgb[2] = cam_norm(b, a.mean[2], a.stdv[2]);
}

template <typename TIn>
__global__ void __launch_bounds__(kThreads) stem_fused_kernel(StemArgs a) {
  pdl_trigger();
  extern __shared__ __align__(16) unsigned char smem[];
  __nv_bfloat16 *hi_s = reinterpret_cast<__nv_bfloat16 *>(smem);                   // [3][35][68] high parts
  __nv_bfloat16 *lo_s = hi_s + 3 * kInH * kInW;                                     // [3][35][68] residuals
  __nv_bfloat16 *l0_s = lo_s + 3 * kInH * kInW;                                     // [561 + 19][20]
  uint32_t *w_s = reinterpret_cast<uint32_t *>(l0_s + (kL0Px + 19) * kL0Pitch);     // fragments (16-byte aligned)
  float *b_s = reinterpret_cast<float *>(w_s + (6 + 36) * 64);                      // [48]
  __nv_bfloat16 *stage_s = hi_s;                                                    // output staging aliases the frame patch

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int ox0 = blockIdx.x * kTW, oy0 = blockIdx.y * kTH, n = blockIdx.z;
  const int ix0 = 4 * ox0 - 4, iy0 = 4 * oy0 - 3;  // frame coordinates of the patch origin (16-byte aligned columns)

  // All global loads of a thread are issued before the first dependent store (the patch is 7 16-byte loads per
  // thread: issued one by one they serialise 7 DRAM latencies, measured 28 % of all stall samples).
  constexpr int kWVec = (6 + 36) * 64 / 4, kWIters = (kWVec + kThreads - 1) / kThreads;
  uint4 wv[kWIters];
#pragma unroll
  for (int it = 0; it < kWIters; ++it) {
    const int i = tid + it * kThreads;
    wv[it] = i < kWVec ? reinterpret_cast<const uint4 *>(a.wfrag)[i] : make_uint4(0u, 0u, 0u, 0u);
  }
  // ---- frame patch -> shared fp32 (zero outside the frame) ----
  const TIn *img = reinterpret_cast<const TIn *>(a.in) + (long long)n * 3 * a.ih * a.iw;
  constexpr int kVecPerRow = kInW / 4, kPatchVec = 3 * kInH * kVecPerRow, kPIters = (kPatchVec + kThreads - 1) / kThreads;
  float pv[kPIters][4];
  const bool vec_ok = (a.iw & 3) == 0;
#pragma unroll
  for (int it = 0; it < kPIters; ++it) {
    const int i = tid + it * kThreads;
    const int c = i / (kInH * kVecPerRow), r = (i / kVecPerRow) % kInH, j = i % kVecPerRow;
    const int iy = iy0 + r, ix = ix0 + 4 * j;
    pv[it][0] = pv[it][1] = pv[it][2] = pv[it][3] = 0.f;
    if (i < kPatchVec && iy >= 0 && iy < a.ih) {
      const TIn *src = img + ((long long)c * a.ih + iy) * a.iw + ix;
      if (vec_ok && ix >= 0 && ix + 3 < a.iw) {
        if (sizeof(TIn) == 4) {
          const float4 q = *reinterpret_cast<const float4 *>(src);
          pv[it][0] = q.x; pv[it][1] = q.y; pv[it][2] = q.z; pv[it][3] = q.w;
        } else {
          const uchar4 q = *reinterpret_cast<const uchar4 *>(src);
          pv[it][0] = (float)q.x; pv[it][1] = (float)q.y; pv[it][2] = (float)q.z; pv[it][3] = (float)q.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (ix + e >= 0 && ix + e < a.iw) pv[it][e] = (float)src[e];
      }
    }
  }
  if (tid < 48) b_s[tid] = a.bias[tid];
#pragma unroll
  for (int it = 0; it < kWIters; ++it) {
    const int i = tid + it * kThreads;
    if (i < kWVec) reinterpret_cast<uint4 *>(w_s)[i] = wv[it];
  }
#pragma unroll
  for (int it = 0; it < kPIters; ++it) {
    const int i = tid + it * kThreads;
    if (i >= kPatchVec) continue;
    uint32_t h[2], l[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float2 v = make_float2(pv[it][2 * e], pv[it][2 * e + 1]);
      if (sizeof(TIn) == 1) v = make_float2(__fdiv_rn(v.x, 255.f), __fdiv_rn(v.y, 255.f));
      split_bf16(v, h[e], l[e]);
    }
    reinterpret_cast<uint2 *>(hi_s)[i] = make_uint2(h[0], h[1]);  // i enumerates (c, r, 4-column group) = the plane layout
    reinterpret_cast<uint2 *>(lo_s)[i] = make_uint2(l[0], l[1]);
  }
  __syncthreads();

  // ---- layer 0: 17 x 33 region, 16 channels, tensor cores ----
  {
    uint32_t bf[3][2][2];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint2 v = reinterpret_cast<const uint2 *>(w_s)[(s * 2 + j) * 32 + lane];
        bf[s][j][0] = v.x; bf[s][j][1] = v.y;
      }
    // this thread's two (ci, ky) combinations per k-step and its column slot
    int off0[3], off2[3];
    const int kx = 2 * (t & 1);  // slot pair (-1, 0) or (1, 2) = patch columns 2x + kx, 2x + kx + 1
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int c0 = min(4 * s + (t >> 1), 8), c2 = min(4 * s + 2 + (t >> 1), 8);  // combos >= 9 carry zero weights
      off0[s] = ((c0 / 3) * kInH + c0 % 3) * kInW + kx;
      off2[s] = ((c2 / 3) * kInH + c2 % 3) * kInW + kx;
    }
    const int ly0 = 2 * oy0 - 1, lx0 = 2 * ox0 - 1;  // layer-0 coordinates of the region origin
    constexpr int kSegs = (kL0Px + 15) / 16;
    for (int seg = warp; seg < kSegs; seg += kThreads / 32) {
      const int p0 = min(seg * 16 + g, kL0Px - 1), p1 = min(seg * 16 + g + 8, kL0Px - 1);
      const int y0 = p0 / kL0W, x0 = p0 % kL0W, y1 = p1 / kL0W, x1 = p1 % kL0W;
      const int r0 = 2 * y0 * kInW + 2 * x0, r1 = 2 * y1 * kInW + 2 * x1;
      float acc[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        uint32_t hi[4], lo[4];
        hi[0] = *reinterpret_cast<const uint32_t *>(hi_s + r0 + off0[s]); lo[0] = *reinterpret_cast<const uint32_t *>(lo_s + r0 + off0[s]);
        hi[1] = *reinterpret_cast<const uint32_t *>(hi_s + r1 + off0[s]); lo[1] = *reinterpret_cast<const uint32_t *>(lo_s + r1 + off0[s]);
        hi[2] = *reinterpret_cast<const uint32_t *>(hi_s + r0 + off2[s]); lo[2] = *reinterpret_cast<const uint32_t *>(lo_s + r0 + off2[s]);
        hi[3] = *reinterpret_cast<const uint32_t *>(hi_s + r1 + off2[s]); lo[3] = *reinterpret_cast<const uint32_t *>(lo_s + r1 + off2[s]);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          mma16816(acc[j], hi, bf[s][j][0], bf[s][j][1]);
          mma16816(acc[j], lo, bf[s][j][0], bf[s][j][1]);
        }
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int p = seg * 16 + g + 8 * half;
        if (p >= kL0Px) continue;
        const int y = half ? y1 : y0, x = half ? x1 : x0;
        const bool inside = (unsigned)(ly0 + y) < (unsigned)(a.ih / 2) && (unsigned)(lx0 + x) < (unsigned)(a.iw / 2);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = 8 * j + 2 * t;
          const float v0 = fmaxf(acc[j][2 * half] + b_s[c], 0.f), v1 = fmaxf(acc[j][2 * half + 1] + b_s[c + 1], 0.f);
          *reinterpret_cast<uint32_t *>(l0_s + p * kL0Pitch + c) = inside ? pack_bf16(v0, v1) : 0u;
        }
      }
    }
  }
  __syncthreads();

  // ---- layer 1: warp w = output row oy0 + w, 16 pixels x 32 channels, one tap per k-step ----
  {
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    const uint2 *wf = reinterpret_cast<const uint2 *>(w_s + 6 * 64);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int r = tap / 3, s = tap % 3;
      const __nv_bfloat16 *p0 = l0_s + ((2 * warp + r) * kL0W + 2 * g + s) * kL0Pitch + 2 * t;
      uint32_t af[4];
      af[0] = *reinterpret_cast<const uint32_t *>(p0);
      af[1] = *reinterpret_cast<const uint32_t *>(p0 + 16 * kL0Pitch);
      af[2] = *reinterpret_cast<const uint32_t *>(p0 + 8);
      af[3] = *reinterpret_cast<const uint32_t *>(p0 + 16 * kL0Pitch + 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint2 b = wf[(tap * 4 + j) * 32 + lane];
        mma16816(acc[j], af, b.x, b.y);
      }
    }
    // bias, ReLU, bf16 -> staging (the frame patch is dead: every warp passed the barrier above) -> 16-byte stores
    __nv_bfloat16 *st = stage_s + warp * 16 * 32;
#pragma unroll
    for (int half = 0; half < 2; ++half)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = 8 * j + 2 * t;
        const float v0 = fmaxf(acc[j][2 * half] + b_s[16 + c], 0.f), v1 = fmaxf(acc[j][2 * half + 1] + b_s[16 + c + 1], 0.f);
        *reinterpret_cast<uint32_t *>(st + (g + 8 * half) * 32 + c) = pack_bf16(v0, v1);
      }
    __syncwarp();
    const int oy = oy0 + warp;
    if (oy < a.oh) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int px = (lane >> 2) + 8 * i, ch = (lane & 3) * 8;
        if (ox0 + px < a.ow)
          *reinterpret_cast<uint4 *>(a.out + (((long long)n * a.oh + oy) * a.ow + ox0 + px) * a.out_pitch + ch) =
              *reinterpret_cast<const uint4 *>(st + px * 32 + ch);
      }
    }
  }
}

// =================================================================================================
// Second generation (design and lane maps: stem_v2.cuh).  Lives in uyd::stemv2 so that the header's constants win
// over the first generation's equally named ones above.
// =================================================================================================
}  // namespace
namespace stemv2 {
__device__ __forceinline__ void mma1688_tf32(float (&d)[4], const uint4 &a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma16816v(float (&d)[4], const uint4 &a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t v) {
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

// PERSIST: two CTAs per SM walk the tiles and keep the 18 layer-1 A fragments in registers (the per-tile re-read of
// those 9 KB by each of the eight warps is 576 of the ~3100 L1/shared wavefronts a tile costs, the kernel's bound).
template <typename TIn, bool PW, bool PERSIST>
__global__ void __launch_bounds__(stemv2::kThreads, PERSIST ? 2 : 3) stem_v2_kernel(StemArgs a) {
  pdl_trigger();
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char *patch = smem;                             // [3][38 lines: even rows at 0, odd rows at 20][68] fp32 (tf32)
  unsigned char *l0s = smem + kPatchBytes;                 // [584 pixels q = 34 y + x][48 B]: word w = channels (w, w+8)
  unsigned char *w1s = smem + kPatchBytes + kL0Bytes;      // layer-1 A fragments [tap][m-tile][lane] x 16 B (!PERSIST)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  uint4 w1r[PERSIST ? 18 : 1];
  if (PERSIST) {
#pragma unroll
    for (int i = 0; i < (PERSIST ? 18 : 0); ++i) w1r[i] = __ldg(reinterpret_cast<const uint4 *>(a.wfrag + kW0Words) + i * 32 + lane);
  }
  const int tiles_x = (a.ow + kTW - 1) / kTW, tiles_y = (a.oh + kTH - 1) / kTH;
  const int total_tiles = PERSIST ? tiles_x * tiles_y * a.n : 1;
  for (int tile = PERSIST ? (int)blockIdx.x : 0; tile < total_tiles; tile += PERSIST ? (int)gridDim.x : 1) {
  int ox0, oy0, n;
  if (PERSIST) {
    const int per = tiles_x * tiles_y, r = tile % per;
    n = tile / per; oy0 = (r / tiles_x) * kTH; ox0 = (r % tiles_x) * kTW;
  } else {
    ox0 = blockIdx.x * kTW; oy0 = blockIdx.y * kTH; n = blockIdx.z;
  }
  const int ix0 = 4 * ox0 - 4, iy0 = 4 * oy0 - 3;  // frame coordinates of the patch origin (16-byte aligned columns)

  // ---- all global loads first: layer-1 fragments and this thread's 7 patch vectors (column j, lines rl + 15 i) ----
  constexpr int kW1Vec = kW1Words / 4, kWIters = (kW1Vec + kThreads - 1) / kThreads;
  const uint4 *w1g = reinterpret_cast<const uint4 *>(a.wfrag + kW0Words);
  uint4 wv[kWIters];
#pragma unroll
  for (int it = 0; it < kWIters; ++it) {
    const int i = tid + it * kThreads;
    wv[it] = (!PERSIST && i < kW1Vec) ? __ldg(w1g + i) : make_uint4(0u, 0u, 0u, 0u);
  }
  // thread (pj, prl): column vector pj of patch rows prl, prl + 15 and (prl < 5) prl + 30 of each channel
  constexpr int kVecPerRow = kInW / 4, kRowLanes = 15;  // 17 vectors per row, 255 loading threads
  static_assert(2 * kRowLanes <= kInH && 3 * kRowLanes >= kInH && kRowLanes * kVecPerRow <= kThreads, "patch load map");
  constexpr bool kCam = std::is_same<TIn, CamIn>::value;
  const TIn *img = reinterpret_cast<const TIn *>(a.in) + (kCam ? 0ll : (long long)n * 3 * a.ih * a.iw);
  asm volatile("" : "+l"(img));  // keep the per-load address arithmetic 32-bit: one IMAD.WIDE.U32 on this base
  const int pj = tid % kVecPerRow, prl = tid / kVecPerRow;
  const int ix = ix0 + 4 * pj;
  const bool col_ok = prl < kRowLanes && ix >= 0 && ix + 3 < a.iw;
  const unsigned plane = (unsigned)(a.ih * a.iw);
  bool ok[3];
  unsigned off[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int r = prl + kRowLanes * k, iy = iy0 + r;
    ok[k] = col_ok && r < kInH && (unsigned)iy < (unsigned)a.ih;
    off[k] = (unsigned)(iy * a.iw + ix);  // meaningful only when ok[k]
  }
  uint4 pv[3][3];
  if constexpr (kCam) {  // camera bytes -> normalised fp32 bit patterns of the four columns of every row this thread owns
    const uint8_t *frame = reinterpret_cast<const uint8_t *>(a.in) + (long long)n * a.frame_stride;
    const uint8_t *uvp = a.uv ? a.uv + (long long)n * a.uv_frame_stride : nullptr;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float px[4][3];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        px[j][0] = px[j][1] = px[j][2] = 0.f;
        if (ok[k]) cam_pixel(a, frame, uvp, iy0 + prl + kRowLanes * k, ix + j, px[j]);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c)
        pv[c][k] = make_uint4(__float_as_uint(px[0][c]), __float_as_uint(px[1][c]), __float_as_uint(px[2][c]), __float_as_uint(px[3][c]));
    }
  } else {
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      pv[c][k] = make_uint4(0u, 0u, 0u, 0u);
      if (ok[k]) {
        const TIn *src = img + (off[k] + c * plane);
        if (sizeof(TIn) == 4) pv[c][k] = __ldg(reinterpret_cast<const uint4 *>(src));
        else pv[c][k].x = __ldg(reinterpret_cast<const unsigned int *>(src));
      }
    }
  }
#pragma unroll
  for (int it = 0; it < kWIters; ++it) {
    const int i = tid + it * kThreads;
    if (!PERSIST && i < kW1Vec) reinterpret_cast<uint4 *>(w1s)[i] = wv[it];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int r = prl + kRowLanes * k;
    if (prl < kRowLanes && r < kInH) {
      unsigned char *dst = patch + (patch_line(0, r) * kInW + 4 * pj) * 4;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint4 v = pv[c][k];
        uint4 o;
        if (sizeof(TIn) == 4 || kCam) {  // round to nearest tf32: the mma reads the upper 19 bits
          o = ok[k] ? make_uint4(v.x + 0x1000u, v.y + 0x1000u, v.z + 0x1000u, v.w + 0x1000u) : v;
        } else {                 // x / 255 exactly as the float pre-process computes it, then the same rounding
          auto cv = [&](uint32_t b) { return ok[k] ? __float_as_uint(div255((float)b)) + 0x1000u : 0u; };
          o = make_uint4(cv(v.x & 0xffu), cv((v.x >> 8) & 0xffu), cv((v.x >> 16) & 0xffu), cv(v.x >> 24));
        }
        *reinterpret_cast<uint4 *>(dst + c * kChanLines * kInW * 4) = o;
      }
    }
  }
  __syncthreads();

  // ---- layer 0: 73 groups of 8 flat pixels, 16 channels = the 16 mma rows, mma.m16n8k8.tf32 ----
  {
    const uint4 *w0g = reinterpret_cast<const uint4 *>(a.wfrag);
    uint4 af[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) af[s] = __ldg(w0g + s * 32 + lane);
    const float b_lo = __ldg(a.bias + g), b_hi = __ldg(a.bias + g + 8);
    const unsigned char *pb[5];
#pragma unroll
    for (int s = 0; s < 5; ++s)  // l0_k_off(s, t) from two compile-time constants
      pb[s] = patch + 64 * warp + 8 * g + 8 * (t & 1) + 4 * ((t >> 1) ? l0_k_off(s, 2) : l0_k_off(s, 0));
    unsigned char *sb = l0s + (8 * warp + 2 * t) * kL0Pitch + 4 * g;
    // groups warp + 8 i: pixels q = 8 (warp + 8 i) + (B: g | D: 2t, 2t+1); three independent mma chains at a time
    auto groups = [&](int i0, auto cnt) {
      constexpr int NG = decltype(cnt)::value;
      float acc[NG][4];
#pragma unroll
      for (int j = 0; j < NG; ++j) { acc[j][0] = acc[j][1] = b_lo; acc[j][2] = acc[j][3] = b_hi; }
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        uint2 v[NG];
#pragma unroll
        for (int j = 0; j < NG; ++j) v[j] = ld64(pb[s] + 512 * (i0 + j));
#pragma unroll
        for (int j = 0; j < NG; ++j) mma1688_tf32(acc[j], af[s], v[j].x, v[j].y);
      }
#pragma unroll
      for (int j = 0; j < NG; ++j) {
        st32(sb + 64 * kL0Pitch * (i0 + j), relu_pack_bf16(acc[j][0], acc[j][2]));
        st32(sb + 64 * kL0Pitch * (i0 + j) + kL0Pitch, relu_pack_bf16(acc[j][1], acc[j][3]));
      }
    };
    static_assert(kGroups == 73, "nine groups per warp and one more for warp 0");
    groups(0, std::integral_constant<int, 3>{});
    groups(3, std::integral_constant<int, 3>{});
    groups(6, std::integral_constant<int, 3>{});
    if (warp == 0) groups(9, std::integral_constant<int, 1>{});
  }
  __syncthreads();
  // layer 1's zero padding: layer-0 row -1 / column -1 exist only in the tiles on the top / left image border
  if (oy0 == 0 || ox0 == 0) {
    if (oy0 == 0)
      for (int i = tid; i < kL0P * 8; i += kThreads) st32(l0s + (i >> 3) * kL0Pitch + 4 * (i & 7), 0u);
    if (ox0 == 0 && tid < kL0H * 8) st32(l0s + (tid >> 3) * kL0P * kL0Pitch + 4 * (tid & 7), 0u);
    __syncthreads();
  }

  // ---- layer 1: warp = output row oy0 + warp, two groups of 8 pixels x two 16-channel m-tiles, one tap per k-step ----
  {
    float acc[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const float2 bz = __ldg(reinterpret_cast<const float2 *>(a.bias + 16 + 4 * g + 2 * mt));
#pragma unroll
      for (int xg = 0; xg < 2; ++xg) { acc[xg][mt][0] = acc[xg][mt][1] = bz.x; acc[xg][mt][2] = acc[xg][mt][3] = bz.y; }
    }
    const unsigned char *bb = l0s + l1_b_off(warp, 0, g, t, 0);
    const uint4 *wf = reinterpret_cast<const uint4 *>(w1s) + lane;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const uint4 a0 = PERSIST ? w1r[PERSIST ? tap * 2 : 0] : wf[(tap * 2 + 0) * 32], a1 = PERSIST ? w1r[PERSIST ? tap * 2 + 1 : 0] : wf[(tap * 2 + 1) * 32];
#pragma unroll
      for (int xg = 0; xg < 2; ++xg) {
        const uint2 b = ld64(bb + ((tap / 3) * kL0P + 16 * xg + tap % 3) * kL0Pitch);
        mma16816v(acc[xg][0], a0, b.x, b.y);
        mma16816v(acc[xg][1], a1, b.x, b.y);
      }
    }
    const int oy = oy0 + warp;
    __nv_bfloat16 *orow = a.out + ((long long)n * a.oh + oy) * a.ow * a.out_pitch;
    if (!PW) {
#pragma unroll
      for (int xg = 0; xg < 2; ++xg)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int px = ox0 + 8 * xg + 2 * t + h;
          if (oy < a.oh && px < a.ow)
            *reinterpret_cast<uint2 *>(orow + (long long)px * a.out_pitch + 4 * g) =
                make_uint2(relu_pack_bf16(acc[xg][0][h], acc[xg][0][2 + h]), relu_pack_bf16(acc[xg][1][h], acc[xg][1][2 + h]));
        }
    } else {
      // 1x1 conv on the bf16-rounded layer-1 output: (channel x pixel) accumulator tiles -> movmatrix -> B fragments
      const uint4 *w2 = reinterpret_cast<const uint4 *>(a.wfrag + kW0Words + kW1Words) + lane;
      const uint4 p0 = __ldg(w2), p1 = __ldg(w2 + 32);
      const float2 bz = __ldg(reinterpret_cast<const float2 *>(a.bias + 48 + 2 * g));
#pragma unroll
      for (int xg = 0; xg < 2; ++xg) {
        float acc2[4] = {bz.x, bz.x, bz.y, bz.y};
        uint32_t bq[2][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) bq[mt][hh] = movmatrix_trans(relu_pack_bf16(acc[xg][mt][2 * hh], acc[xg][mt][2 * hh + 1]));
        mma16816v(acc2, p0, bq[0][0], bq[0][1]);
        mma16816v(acc2, p1, bq[1][0], bq[1][1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int px = ox0 + 8 * xg + 2 * t + h;
          if (oy < a.oh && px < a.ow)
            *reinterpret_cast<uint32_t *>(orow + (long long)px * a.out_pitch + 2 * g) = relu_pack_bf16(acc2[h], acc2[2 + h]);
        }
      }
    }
  }
  }  // tile loop
}
}  // namespace stemv2
namespace {

uint32_t pack2(float lo, float hi) {
  __nv_bfloat16 x = __float2bfloat16_rn(lo), y = __float2bfloat16_rn(hi);
  uint16_t ux, uy;
  memcpy(&ux, &x, 2);
  memcpy(&uy, &y, 2);
  return (uint32_t)ux | ((uint32_t)uy << 16);
}

// B fragments of mma.m16n8k16 for K x N weights given by wfun(k, n)
template <class F>
void pack_frags(std::vector<uint32_t> &out, int ksteps, int ntiles, F wfun) {
  for (int s = 0; s < ksteps; ++s)
    for (int j = 0; j < ntiles; ++j)
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3, n = 8 * j + g;
        out.push_back(pack2(wfun(16 * s + 2 * t, n), wfun(16 * s + 2 * t + 1, n)));
        out.push_back(pack2(wfun(16 * s + 2 * t + 8, n), wfun(16 * s + 2 * t + 9, n)));
      }
}

}  // namespace

bool stem_fused_supported(int c0, int c1, int ih, int iw, int out_pitch, int out_coff) {
  return c0 == 16 && c1 == 32 && ih % 4 == 0 && iw % 4 == 0 && out_pitch % 8 == 0 && out_coff % 8 == 0;
}

// w0 [16][3][3][3], w1 [32][16][3][3] (BN folded, PyTorch layout)
// frags = [legacy | second generation (stem_v2.cuh)], bias = [16 | 32 | 16 (1x1, zero without w2)]
void stem_fused_pack(const float *w0, const float *b0, const float *w1, const float *b1, const float *w2, const float *b2,
                     std::vector<uint32_t> &frags, std::vector<float> &bias) {
  frags.clear();
  pack_frags(frags, 3, 2, [&](int k, int n) {  // k = (ci * 3 + ky) * 4 + slot, slot <-> kx = slot - 1
    const int combo = k / 4, kx = k % 4 - 1;
    return (combo < 9 && kx >= 0) ? w0[((size_t)n * 3 + combo / 3) * 9 + (combo % 3) * 3 + kx] : 0.f;
  });
  pack_frags(frags, 9, 4, [&](int k, int n) {  // k = tap * 16 + ci
    return w1[((size_t)n * 16 + k % 16) * 9 + k / 16];
  });
  bias.assign(64, 0.f);
  for (int i = 0; i < 16; ++i) bias[i] = b0[i];
  for (int i = 0; i < 32; ++i) bias[16 + i] = b1[i];
  for (int i = 0; i < 16 && b2; ++i) bias[48 + i] = b2[i];
  stemv2::pack(w0, w1, w2, frags);
}

static int stem_v2_launch(const StemArgs &a0, cudaStream_t s) {
  StemArgs a = a0;
  a.wfrag += (6 + 36) * 64;  // skip the legacy fragments
  static const bool persist = [] { const char *v = getenv("UYD_STEM_PERSIST"); return !(v && *v == '0'); }();
  const int sms = current_sm_count();
  UYD_REQUIRE(sms > 0, UYD_E_NOGPU, "stem: no current CUDA device");
  const int tiles = ceil_div(a.ow, stemv2::kTW) * ceil_div(a.oh, stemv2::kTH) * a.n;
  const size_t smem = stemv2::kSmemBytes;
  auto go = [&](auto kern, bool pers) -> int {
    if (int e = smem_optin(kern, stemv2::kSmemBytes)) return e;
    if (pers) kern<<<tiles < 2 * sms ? tiles : 2 * sms, stemv2::kThreads, smem, s>>>(a);
    else kern<<<dim3(ceil_div(a.ow, stemv2::kTW), ceil_div(a.oh, stemv2::kTH), a.n), stemv2::kThreads, smem, s>>>(a);
    return (int)cudaGetLastError();
  };
  // persistent CTAs pay off once there are several tiles per CTA (batch-1 frames keep one tile per CTA)
  const bool pers = persist && tiles >= 8 * sms;
  using namespace stemv2;
  if (a.cam) {
    if (a.pw) return pers ? go(stem_v2_kernel<CamIn, true, true>, true) : go(stem_v2_kernel<CamIn, true, false>, false);
    return pers ? go(stem_v2_kernel<CamIn, false, true>, true) : go(stem_v2_kernel<CamIn, false, false>, false);
  }
  if (a.u8) {
    if (a.pw) return pers ? go(stem_v2_kernel<uint8_t, true, true>, true) : go(stem_v2_kernel<uint8_t, true, false>, false);
    return pers ? go(stem_v2_kernel<uint8_t, false, true>, true) : go(stem_v2_kernel<uint8_t, false, false>, false);
  }
  if (a.pw) return pers ? go(stem_v2_kernel<float, true, true>, true) : go(stem_v2_kernel<float, true, false>, false);
  return pers ? go(stem_v2_kernel<float, false, true>, true) : go(stem_v2_kernel<float, false, false>, false);
}

int stem_fused_launch(const StemArgs &a, cudaStream_t s) {
  static const bool legacy = [] { const char *v = getenv("UYD_STEM_LEGACY"); return v && *v == '1'; }();
  if (!legacy || a.pw || a.cam) return stem_v2_launch(a, s);
  const size_t smem = (size_t)2 * 3 * kInH * kInW * 2 + (size_t)(kL0Px + 19) * kL0Pitch * 2 + (6 + 36) * 64 * 4 + 48 * 4;
  if (int e = smem_optin(stem_fused_kernel<float>, smem)) return e;
  if (int e = smem_optin(stem_fused_kernel<uint8_t>, smem)) return e;
  dim3 grid(ceil_div(a.ow, kTW), ceil_div(a.oh, kTH), a.n);
  if (a.u8) stem_fused_kernel<uint8_t><<<grid, kThreads, smem, s>>>(a);
  else stem_fused_kernel<float><<<grid, kThreads, smem, s>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace uyd
