// Memory-bound helpers of the conv stack: SPPF max-pool cascade, nearest x2 upsample,
// NHWC -> NCHW export of the raw heads.
#include "common.cuh"

namespace uyd {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint4 max8(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162 *pa = reinterpret_cast<const __nv_bfloat162 *>(&a);
  const __nv_bfloat162 *pb = reinterpret_cast<const __nv_bfloat162 *>(&b);
  __nv_bfloat162 *pr = reinterpret_cast<__nv_bfloat162 *>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

// SPPF_DLA (trainer.py:119-124 / model.py:127-132): y1 = P(x), y2 = P(y1), y3 = P(y2) with
// P = MaxPool2d(5, 1, 2) and -inf padding.  Cascading clipped windows composes exactly:
// y1 = max over the clipped 5x5, y2 = 9x9, y3 = 13x13 window of x.  One CTA handles one
// image row x 8-channel group: the 13 input rows are reduced column-wise into shared memory
// (vertical maxima for radius 2/4/6), then each thread reduces horizontally.
__global__ void __launch_bounds__(kThreads) sppf_pool_kernel(__nv_bfloat16 *base, int h, int w, int pitch, int c) {
  pdl_trigger();
  extern __shared__ uint4 col[];  // [3][w][cg] vertical maxima
  const int cg = c / 8;
  const int y = blockIdx.x % h;
  const int n = blockIdx.x / h;
  const __nv_bfloat16 *img = base + (long long)n * h * w * pitch;
  const int items = w * cg;
  const uint4 ninf = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);  // bf16 -inf x8
  for (int it = threadIdx.x; it < items; it += kThreads) {
    const int x = it / cg, g = it % cg;
    uint4 m2 = ninf, m4 = ninf, m6 = ninf;
#pragma unroll
    for (int dy = -6; dy <= 6; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= h) continue;
      const uint4 v = *reinterpret_cast<const uint4 *>(img + ((long long)yy * w + x) * pitch + g * 8);
      m6 = max8(m6, v);
      if (dy >= -4 && dy <= 4) m4 = max8(m4, v);
      if (dy >= -2 && dy <= 2) m2 = max8(m2, v);
    }
    col[it] = m2;
    col[items + it] = m4;
    col[2 * items + it] = m6;
  }
  __syncthreads();
  __nv_bfloat16 *row = base + ((long long)n * h + y) * w * pitch;
  for (int it = threadIdx.x; it < items; it += kThreads) {
    const int x = it / cg, g = it % cg;
    uint4 r2 = ninf, r4 = ninf, r6 = ninf;
#pragma unroll
    for (int dx = -6; dx <= 6; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= w) continue;
      const int j = xx * cg + g;
      r6 = max8(r6, col[2 * items + j]);
      if (dx >= -4 && dx <= 4) r4 = max8(r4, col[items + j]);
      if (dx >= -2 && dx <= 2) r2 = max8(r2, col[j]);
    }
    __nv_bfloat16 *o = row + (long long)x * pitch + g * 8;
    *reinterpret_cast<uint4 *>(o + c) = r2;
    *reinterpret_cast<uint4 *>(o + 2 * c) = r4;
    *reinterpret_cast<uint4 *>(o + 3 * c) = r6;
  }
}

// Plane formulation of the same cascade (used when two planes of h x w x 32 bytes fit shared memory, i.e. the
// network's 40 x 40 SPPF): one CTA owns (image, two 8-channel groups = one 32-byte sector per pixel), loads the plane
// ONCE and runs the three cascaded 5 x 5 pools separably in shared memory (horizontal into B, vertical back into A and
// out to the concat slice).  The row kernel above re-reads every input row 13 times through L2 (3.5 TB/s of L2
// traffic, 49 us at batch 64); here every byte is read once.
constexpr int kPlaneThreads = 512;
__global__ void __launch_bounds__(kPlaneThreads) sppf_plane_kernel(__nv_bfloat16 *base, int h, int w, int pitch, int c, int gp) {
  pdl_trigger();
  extern __shared__ uint4 pl[];  // A [h*w*gp] | B [h*w*gp]
  const int groups = (c / 8) / gp;
  const int n = blockIdx.x / groups, g0 = (blockIdx.x % groups) * gp;
  const int items = h * w * gp;
  uint4 *A = pl, *B = pl + items;
  __nv_bfloat16 *img = base + (long long)n * h * w * pitch + g0 * 8;
  const uint4 ninf = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);  // bf16 -inf x8
  for (int it = threadIdx.x; it < items; it += kPlaneThreads) {
    const int px = it / gp, g = it - px * gp;
    A[it] = *reinterpret_cast<const uint4 *>(img + (long long)px * pitch + g * 8);
  }
  __syncthreads();
  const int rowi = w * gp;
  for (int stage = 1; stage <= 3; ++stage) {
    for (int it = threadIdx.x; it < items; it += kPlaneThreads) {  // horizontal 5-max, clipped = -inf padding
      const int x = (it % rowi) / gp;
      uint4 m = A[it];
      if (x >= 1) m = max8(m, A[it - gp]);
      if (x >= 2) m = max8(m, A[it - 2 * gp]);
      if (x + 1 < w) m = max8(m, A[it + gp]);
      if (x + 2 < w) m = max8(m, A[it + 2 * gp]);
      B[it] = m;
    }
    __syncthreads();
    for (int it = threadIdx.x; it < items; it += kPlaneThreads) {  // vertical 5-max -> A (input of the next pool) and out
      const int y = it / rowi;
      uint4 m = B[it];
      if (y >= 1) m = max8(m, B[it - rowi]);
      if (y >= 2) m = max8(m, B[it - 2 * rowi]);
      if (y + 1 < h) m = max8(m, B[it + rowi]);
      if (y + 2 < h) m = max8(m, B[it + 2 * rowi]);
      A[it] = m;
      const int px = it / gp, g = it - px * gp;
      *reinterpret_cast<uint4 *>(img + (long long)px * pitch + stage * c + g * 8) = m;
    }
    __syncthreads();
  }
}

// One thread per INPUT element (8 channels of one pixel): one 16-byte load, four 16-byte stores (was one load per
// output element: 4x the load instructions for the same bytes; c64 -> 160 x 160 at batch 64: 84 us).
__global__ void __launch_bounds__(kThreads) upsample2x_kernel(const __nv_bfloat16 *in, int in_pitch, __nv_bfloat16 *out,
                                                              int out_pitch, long long total_in, int oh, int ow, int cg) {
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (t >= total_in) return;
  const int ih = oh / 2, iw = ow / 2;
  const int g = (int)(t % cg);
  const long long p = t / cg;   // input pixel (n, iy, ix)
  const int ix = (int)(p % iw);
  const int iy = (int)((p / iw) % ih);
  const long long n = p / ((long long)iw * ih);
  const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in + p * in_pitch + g * 8));
  __nv_bfloat16 *o = out + ((n * oh + 2 * iy) * ow + 2 * ix) * out_pitch + g * 8;
  *reinterpret_cast<uint4 *>(o) = v;
  *reinterpret_cast<uint4 *>(o + out_pitch) = v;
  o += (long long)ow * out_pitch;
  *reinterpret_cast<uint4 *>(o) = v;
  *reinterpret_cast<uint4 *>(o + out_pitch) = v;
}

// [n,h,w,c] fp32 -> [n,c,h,w] fp32 through a 32x32 shared-memory transpose of (hw, c).
__global__ void nhwc_to_nchw_kernel(const float *in, float *out, int hw, int c) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float *src = in + (long long)n * hw * c;
  float *dst = out + (long long)n * hw * c;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, cc = c0 + threadIdx.x;
    if (p < hw && cc < c) tile[i][threadIdx.x] = src[(long long)p * c + cc];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int cc = c0 + i, p = p0 + threadIdx.x;
    if (p < hw && cc < c) dst[(long long)cc * hw + p] = tile[threadIdx.x][i];
  }
}

}  // namespace

int sppf_pool_launch(__nv_bfloat16 *base, int n, int h, int w, int pitch, int c, cudaStream_t s) {
  UYD_REQUIRE(c % 8 == 0 && pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0, UYD_E_UNSUPPORTED,
              "sppf pool needs C %% 8 == 0 and 16-byte aligned slices");
  static const bool rows_only = [] { const char *v = getenv("UYD_SPPF_ROWS"); return v && *v == '1'; }();
  const int gp = (c / 8) % 2 == 0 ? 2 : 1;
  const size_t plane = (size_t)2 * h * w * gp * sizeof(uint4);
  if (!rows_only && plane <= 110 * 1024) {  // two CTAs per SM
    if (int e = smem_optin(sppf_plane_kernel, 110 * 1024)) return e;
    sppf_plane_kernel<<<n * ((c / 8) / gp), kPlaneThreads, plane, s>>>(base, h, w, pitch, c, gp);
    return (int)cudaGetLastError();
  }
  const size_t smem = (size_t)3 * w * (c / 8) * sizeof(uint4);
  UYD_REQUIRE(smem <= 200 * 1024, UYD_E_UNSUPPORTED, "sppf row does not fit shared memory");
  if (smem > 48 * 1024)
    if (int e = smem_optin(sppf_pool_kernel, 200 * 1024)) return e;
  sppf_pool_kernel<<<n * h, kThreads, smem, s>>>(base, h, w, pitch, c);
  return (int)cudaGetLastError();
}

int upsample2x_launch(const __nv_bfloat16 *in, int in_pitch, __nv_bfloat16 *out, int out_pitch, int n, int ih, int iw,
                      int c, cudaStream_t s) {
  UYD_REQUIRE(c % 8 == 0 && in_pitch % 8 == 0 && out_pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              UYD_E_UNSUPPORTED, "upsample needs C %% 8 == 0 and 16-byte aligned slices");
  const long long total = (long long)n * ih * iw * (c / 8);   // input elements
  upsample2x_kernel<<<(unsigned)((total + kThreads - 1) / kThreads), kThreads, 0, s>>>(in, in_pitch, out, out_pitch,
                                                                                        total, ih * 2, iw * 2, c / 8);
  return (int)cudaGetLastError();
}

int nhwc_to_nchw_f32_launch(const float *in, float *out, int n, int h, int w, int c, cudaStream_t s) {
  dim3 grid(ceil_div(h * w, 32), ceil_div(c, 32), n), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, s>>>(in, out, h * w, c);
  return (int)cudaGetLastError();
}

}  // namespace uyd
