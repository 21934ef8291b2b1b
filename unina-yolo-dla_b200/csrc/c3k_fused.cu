// Host side of the fused C3k block (shape rule, weight packing, tile-height choice, launch of the flat-frame kernel
// in c3k_flat.cu) and the fused class-branch kernel of the Detect head that the unchained plans use
// (UYD_NO_CHAIN=1, shapes conv_chain.cu does not take): DWConv -> 1x1 -> DWConv -> 1x1 -> 1x1 in one launch on
// mma.sync.m16n8k16, the stage helpers below.  (The first-generation C3k kernel that lived here was removed after
// a round of soak of c3k_flat_kernel.)
#include <cstdlib>

#include "common.cuh"

namespace uyd {

// c3k_flat.cu
size_t c3k_flat_smem_bytes(int c_, int th);
void c3k_flat_pack(int c, const float *const w[7], std::vector<uint32_t> &frags);
struct C3kArgs;
int c3k_flat_launch(int c, const C3kArgs &a, cudaStream_t s);
int c3k_flat_q_launch(int c, const C3kArgs &a, cudaStream_t s);   // INT8 (fake-quant) variant
size_t c3k_flat_q_smem_bytes(int c_, int th);
void c3k_flat_pack(int c, const float *const w[7], std::vector<uint32_t> &frags);
int c3k_flat_words(int c);
// c3k_tc.cu: the tcgen05 generation (c = 16 / 32, batches that fill the GPU with full-width strips)
bool c3k_tc_supported(int c, int h, int w);
void c3k_tc_pack(int c, const float *const w[7], std::vector<uint32_t> &out);
int c3k_tc_launch(int c, const C3kArgs &a, const uint32_t *w_tc, cudaStream_t s);
int c3k_tc_strips(int c, int h, int w, int n);
bool c3k_tc_preferred(int c, int h, int w, int n);

struct C3kArgs {
  const __nv_bfloat16 *in;
  __nv_bfloat16 *out;
  const uint32_t *wfrag;  // packed B fragments of the 7 convs
  const float *bias;      // [7][32]
  int n, h, w, in_pitch, out_pitch, th, tiles_x, tiles_y;
};

struct ClsArgs {
  const __nv_bfloat16 *in;
  float *out;               // head buffer slice base (fp32), pitch in floats
  const __nv_bfloat16 *wdw; // [9][cin] then [9][32]  (bf16, tap-major)
  const uint32_t *wfrag;    // pw1, pw2, pw3 fragments
  const float *bias;        // [5][128]: dw1, pw1, dw2, pw2, pw3
  int n, h, w, in_pitch, out_pitch, nc, tiles_x, tiles_y;
};

namespace {

constexpr int kTW = 40;         // output tile width
constexpr int kPW = kTW + 8;    // region row pitch in pixels (4-pixel halo each side) = 3 x 16
constexpr int kWarps = 8;
constexpr int kThreadsC3k = kWarps * 32;

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
__device__ __forceinline__ float2 unpack_bf16(uint32_t v) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&v));
}

// K-index layout of a 3x3 conv with C channels per tap: k-step s holds taps [s*TPK, (s+1)*TPK),
// TPK = 16 / C; column kk of the step is (tap s*TPK + kk / C, channel kk % C).
template <int C>
struct K3 {
  static constexpr int TPK = 16 / C;
  static constexpr int STEPS = (9 + TPK - 1) / TPK;
};

// Geometry of one stage: output region rows [r0, r1) x cols [c0, c1) in the halo frame.
struct Region { int r0, r1, c0, c1; };

// 3x3 conv  src[frame][C] -> dst[frame][C]  (+ residual from res[frame][C]) over `reg`.
// gy0/gx0: image coordinates of frame pixel (0,0).
template <int C>
__device__ __forceinline__ void stage_conv3(const __nv_bfloat16 *src, __nv_bfloat16 *dst, const __nv_bfloat16 *res,
                                            const uint32_t *wf, const float *bias, Region reg, int gy0, int gx0, int H,
                                            int W, int warp, int lane) {
  constexpr int NT = (C + 7) / 8;
  constexpr int STEPS = K3<C>::STEPS, TPK = K3<C>::TPK;
  const int g = lane >> 2, t = lane & 3;
  uint32_t bf[STEPS][NT][2];
#pragma unroll
  for (int s = 0; s < STEPS; ++s)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const uint2 v = reinterpret_cast<const uint2 *>(wf)[(s * NT + j) * 32 + lane];
      bf[s][j][0] = v.x;
      bf[s][j][1] = v.y;
    }
  // this thread's two k-columns inside a step: 2t (a0/a1) and 2t+8 (a2/a3)
  const int j0 = (2 * t) / C, ch0 = (2 * t) % C, j2 = (2 * t + 8) / C, ch2 = (2 * t + 8) % C;
  // element offsets of this thread's two A columns relative to the segment's first pixel, per k-step
  // (hoisted: the tap -> (row, col) arithmetic would otherwise run for every segment)
  int off0[STEPS], off2[STEPS];
#pragma unroll
  for (int s = 0; s < STEPS; ++s) {
    const int tap0 = s * TPK + j0, tap2 = s * TPK + j2;
    off0[s] = tap0 < 9 ? ((tap0 / 3 - 1) * kPW + tap0 % 3 - 1) * C + ch0 : -(1 << 30);
    off2[s] = tap2 < 9 ? ((tap2 / 3 - 1) * kPW + tap2 % 3 - 1) * C + ch2 : -(1 << 30);
  }
  const int segs_per_row = (reg.c1 - reg.c0 + 15) / 16;
  const int nseg = (reg.r1 - reg.r0) * segs_per_row;
  for (int seg = warp; seg < nseg; seg += kWarps) {
    const int ry = reg.r0 + seg / segs_per_row;
    const int rx = reg.c0 + (seg % segs_per_row) * 16;
    const __nv_bfloat16 *seg_base = src + (ry * kPW + rx + g) * C;
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int s = 0; s < STEPS; ++s) {
      uint32_t a[4];
      if (off0[s] > -(1 << 29)) {
        const __nv_bfloat16 *p = seg_base + off0[s];
        a[0] = *reinterpret_cast<const uint32_t *>(p);
        a[1] = *reinterpret_cast<const uint32_t *>(p + 8 * C);
      } else {
        a[0] = a[1] = 0u;
      }
      if (off2[s] > -(1 << 29)) {
        const __nv_bfloat16 *p = seg_base + off2[s];
        a[2] = *reinterpret_cast<const uint32_t *>(p);
        a[3] = *reinterpret_cast<const uint32_t *>(p + 8 * C);
      } else {
        a[2] = a[3] = 0u;
      }
#pragma unroll
      for (int j = 0; j < NT; ++j) mma16816(acc[j], a, bf[s][j][0], bf[s][j][1]);
    }
    // epilogue: rows g and g+8 of the segment, channels 8j + 2t, 8j + 2t + 1
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int px = rx + g + 8 * half;
      if (px >= reg.c1) continue;
      const bool inside = (unsigned)(gy0 + ry) < (unsigned)H && (unsigned)(gx0 + px) < (unsigned)W;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int n = 8 * j + 2 * t;
        if (n >= C) continue;
        float v0 = fmaxf(acc[j][2 * half] + bias[n], 0.f), v1 = fmaxf(acc[j][2 * half + 1] + bias[n + 1], 0.f);
        const int o = (ry * kPW + px) * C + n;
        if (res) {
          const float2 r = unpack_bf16(*reinterpret_cast<const uint32_t *>(res + o));
          v0 += r.x;
          v1 += r.y;
        }
        *reinterpret_cast<uint32_t *>(dst + o) = inside ? pack_bf16(v0, v1) : 0u;
      }
    }
  }
}

// 1x1 conv over one or two channel sources: K = [srcA (CA ch, frame-indexed) | srcB (CB ch)].
// srcB (if CB > 0) and dst may use a compact frame (pitch dpw, origin at region (r0, c0)).
template <int CA, int CB, int COUT, int NW = kWarps>
__device__ __forceinline__ void stage_conv1(const __nv_bfloat16 *srcA, const __nv_bfloat16 *srcB, __nv_bfloat16 *dst,
                                            bool dst_compact, const uint32_t *wf, const float *bias, Region reg, int gy0,
                                            int gx0, int H, int W, int warp, int lane) {
  constexpr int K = CA + CB, STEPS = (K + 15) / 16, NT = (COUT + 7) / 8;
  const int g = lane >> 2, t = lane & 3;
  uint32_t bf[STEPS][NT][2];
#pragma unroll
  for (int s = 0; s < STEPS; ++s)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const uint2 v = reinterpret_cast<const uint2 *>(wf)[(s * NT + j) * 32 + lane];
      bf[s][j][0] = v.x;
      bf[s][j][1] = v.y;
    }
  const int cw = reg.c1 - reg.c0;  // compact pitch
  const int segs_per_row = (cw + 15) / 16;
  const int nseg = (reg.r1 - reg.r0) * segs_per_row;
  for (int seg = warp; seg < nseg; seg += NW) {
    const int ry = reg.r0 + seg / segs_per_row;
    const int rx = reg.c0 + (seg % segs_per_row) * 16;
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int s = 0; s < STEPS; ++s) {
      uint32_t a[4];
#pragma unroll
      for (int q = 0; q < 2; ++q) {  // q = 0: k = 16s + 2t ; q = 1: k = 16s + 2t + 8
        const int k = 16 * s + 2 * t + 8 * q;
        uint32_t lo = 0u, hi = 0u;
        if (k < CA) {
          const __nv_bfloat16 *p = srcA + (ry * kPW + rx + g) * CA + k;
          lo = *reinterpret_cast<const uint32_t *>(p);
          hi = *reinterpret_cast<const uint32_t *>(p + 8 * CA);
        } else if (k < K) {
          const __nv_bfloat16 *p = srcB + ((ry - reg.r0) * cw + rx - reg.c0 + g) * CB + (k - CA);
          lo = *reinterpret_cast<const uint32_t *>(p);
          hi = *reinterpret_cast<const uint32_t *>(p + 8 * CB);
        }
        a[2 * q] = lo;
        a[2 * q + 1] = hi;
      }
#pragma unroll
      for (int j = 0; j < NT; ++j) mma16816(acc[j], a, bf[s][j][0], bf[s][j][1]);
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int px = rx + g + 8 * half;
      if (px >= reg.c1) continue;
      const bool inside = (unsigned)(gy0 + ry) < (unsigned)H && (unsigned)(gx0 + px) < (unsigned)W;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int n = 8 * j + 2 * t;
        if (n >= COUT) continue;
        const float v0 = fmaxf(acc[j][2 * half] + bias[n], 0.f), v1 = fmaxf(acc[j][2 * half + 1] + bias[n + 1], 0.f);
        const int o = dst_compact ? ((ry - reg.r0) * cw + px - reg.c0) * COUT + n : (ry * kPW + px) * COUT + n;
        *reinterpret_cast<uint32_t *>(dst + o) = inside ? pack_bf16(v0, v1) : 0u;
      }
    }
  }
}

// words of packed B fragments per conv
template <int C> constexpr int frag_words_3() { return K3<C>::STEPS * ((C + 7) / 8) * 64; }
constexpr int frag_words_1(int k, int cout) { return ((k + 15) / 16) * ((cout + 7) / 8) * 64; }

constexpr int kClsTH = 8;
constexpr int kClsWarps = 16, kClsThreads = kClsWarps * 32;

// depth-wise 3x3 + bias + ReLU over `reg`:  src[frame][C] -> dst[frame][C]; thread = (2 adjacent pixels, 8 channels):
// the 3 x 4 input window and the converted weights are shared by both outputs.
template <int C, int NTHREADS>
__device__ __forceinline__ void stage_dw(const __nv_bfloat16 *src, __nv_bfloat16 *dst, const __nv_bfloat16 *w,
                                         const float *bias, Region reg, int tid) {
  constexpr int CG = C / 8;
  const int rw = reg.c1 - reg.c0, rw2 = (rw + 1) / 2;
  const int total = (reg.r1 - reg.r0) * rw2 * CG;
  for (int i = tid; i < total; i += NTHREADS) {
    const int cg = i % CG, px = ((i / CG) % rw2) * 2 + reg.c0, ry = i / (CG * rw2) + reg.r0;
    float acc[2][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = bias[cg * 8 + j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      float wv[3][8];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const uint4 wr = *reinterpret_cast<const uint4 *>(w + (ky * 3 + kx) * C + cg * 8);
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(&wr);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16(wp[j]);
          wv[kx][2 * j] = f.x;
          wv[kx][2 * j + 1] = f.y;
        }
      }
      const __nv_bfloat16 *row = src + ((ry + ky - 1) * kPW + px - 1) * C + cg * 8;
#pragma unroll
      for (int col = 0; col < 4; ++col) {
        const uint4 xr = *reinterpret_cast<const uint4 *>(row + col * C);
        const uint32_t *xp = reinterpret_cast<const uint32_t *>(&xr);
        float xv[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16(xp[j]);
          xv[2 * j] = f.x;
          xv[2 * j + 1] = f.y;
        }
        if (col < 3) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[0][j] = fmaf(xv[j], wv[col][j], acc[0][j]);
        }
        if (col > 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[1][j] = fmaf(xv[j], wv[col - 1][j], acc[1][j]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (px + q >= reg.c1) break;
      uint4 o;
      uint32_t *op = reinterpret_cast<uint32_t *>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) op[j] = pack_bf16(fmaxf(acc[q][2 * j], 0.f), fmaxf(acc[q][2 * j + 1], 0.f));
      *reinterpret_cast<uint4 *>(dst + (ry * kPW + px + q) * C + cg * 8) = o;
    }
  }
}

template <int CIN>
__global__ void __launch_bounds__(kClsThreads) cls_branch_fused_kernel(ClsArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int MID = 32, CW = CIN > MID ? CIN : MID;
  constexpr int TH = kClsTH;
  constexpr int frame_px = (TH + 4 + 1) * kPW;  // halo 2 + one slack row
  __nv_bfloat16 *XZ = reinterpret_cast<__nv_bfloat16 *>(smem);  // x [frame][CIN], later z1 / z2 [frame][32]
  __nv_bfloat16 *Y = XZ + (size_t)frame_px * CW;                // y1 [frame][CIN], later y2 [frame][32]
  __shared__ float sbias[5][128];
  __shared__ __align__(16) __nv_bfloat16 swdw[9 * (CIN + MID)];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int n = tile / (a.tiles_x * a.tiles_y);
  const int tr = tile % (a.tiles_x * a.tiles_y);
  const int ty0 = (tr / a.tiles_x) * TH, tx0 = (tr % a.tiles_x) * kTW;
  const int gy0 = ty0 - 2, gx0 = tx0 - 2;
  const int H = a.h, W = a.w;
  for (int i = tid; i < 5 * 128; i += kClsThreads) sbias[i / 128][i % 128] = a.bias[i];
  for (int i = tid; i < 9 * (CIN + MID); i += kClsThreads) swdw[i] = a.wdw[i];
  {  // x tile + halo (zero outside the image / in the slack columns and row)
    constexpr int CH16 = CIN / 8;
    const __nv_bfloat16 *img = a.in + (long long)n * H * W * a.in_pitch;
#pragma unroll 4
    for (int i = tid; i < frame_px * CH16; i += kClsThreads) {
      const int px = i / CH16, ch = i % CH16;
      const int ry = px / kPW, rx = px % kPW;
      const int gy = gy0 + ry, gx = gx0 + rx;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (ry < TH + 4 && rx < kTW + 4 && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W)
        v = *reinterpret_cast<const uint4 *>(img + ((long long)gy * W + gx) * a.in_pitch + ch * 8);
      *reinterpret_cast<uint4 *>(XZ + (size_t)px * CIN + ch * 8) = v;
    }
    // (Y is read beyond the region only by mma rows that are never stored: rows are independent, no init needed)
  }
  __syncthreads();
  const Region R1{1, TH + 3, 1, kTW + 3}, R0{2, TH + 2, 2, kTW + 2};
  constexpr int W1 = frag_words_1(CIN, MID), W2 = frag_words_1(MID, MID);
  stage_dw<CIN, kClsThreads>(XZ, Y, swdw, sbias[0], R1, tid);                                                           // y1
  __syncthreads();
  stage_conv1<CIN, 0, MID, kClsWarps>(Y, nullptr, XZ, false, a.wfrag, sbias[1], R1, gy0, gx0, H, W, warp, lane);      // z1 (0 outside)
  __syncthreads();
  stage_dw<MID, kClsThreads>(XZ, Y, swdw + 9 * CIN, sbias[2], R0, tid);                                                 // y2
  __syncthreads();
  stage_conv1<MID, 0, MID, kClsWarps>(Y, nullptr, XZ, false, a.wfrag + W1, sbias[3], R0, gy0, gx0, H, W, warp, lane); // z2
  __syncthreads();
  {  // logits = pw3(z2) + bias  ->  fp32 head slice (N padded to 8, K = 32 = 2 k-steps)
    const uint32_t *wf = a.wfrag + W1 + W2;
    const int g = lane >> 2, t = lane & 3;
    uint32_t bf[2][2];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      const uint2 v = reinterpret_cast<const uint2 *>(wf)[s2 * 32 + lane];
      bf[s2][0] = v.x; bf[s2][1] = v.y;
    }
    float *img = a.out + (long long)n * H * W * a.out_pitch;
    const int nseg = TH * 3;
    for (int seg = warp; seg < nseg; seg += kClsWarps) {
      const int ry = R0.r0 + seg / 3, rx = R0.c0 + (seg % 3) * 16;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        uint32_t af[4];
        const __nv_bfloat16 *p = XZ + (ry * kPW + rx + g) * MID + 16 * s2 + 2 * t;
        af[0] = *reinterpret_cast<const uint32_t *>(p);
        af[1] = *reinterpret_cast<const uint32_t *>(p + 8 * MID);
        af[2] = *reinterpret_cast<const uint32_t *>(p + 8);
        af[3] = *reinterpret_cast<const uint32_t *>(p + 8 * MID + 8);
        mma16816(acc, af, bf[s2][0], bf[s2][1]);
      }
      const int nch = 2 * t;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int px = rx + g + 8 * half;
        const int gy = gy0 + ry, gx = gx0 + px;
        if (px >= R0.c1 || gy >= H || gx >= W) continue;
        float *o = img + ((long long)gy * W + gx) * a.out_pitch + nch;
        if (nch < a.nc) o[0] = acc[2 * half] + sbias[4][nch];
        if (nch + 1 < a.nc) o[1] = acc[2 * half + 1] + sbias[4][nch + 1];
      }
    }
  }
}

// ---- host-side weight packing ------------------------------------------------------------------
inline uint32_t pack2(float lo, float hi) {
  __nv_bfloat16 a = __float2bfloat16_rn(lo), b = __float2bfloat16_rn(hi);
  uint16_t ua, ub;
  memcpy(&ua, &a, 2);
  memcpy(&ub, &b, 2);
  return (uint32_t)ua | ((uint32_t)ub << 16);
}

// kfun(k) -> weight of K-index k for output channel n (0 when padded)
template <class F>
void pack_frags(std::vector<uint32_t> &out, int ksteps, int ntiles, int cout, F wfun) {
  for (int s = 0; s < ksteps; ++s)
    for (int j = 0; j < ntiles; ++j)
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3, n = 8 * j + g;
        auto w = [&](int kk) { return n < cout ? wfun(s, kk, n) : 0.f; };
        out.push_back(pack2(w(2 * t), w(2 * t + 1)));
        out.push_back(pack2(w(2 * t + 8), w(2 * t + 9)));
      }
}

}  // namespace

size_t c3k_smem_bytes(int c_, int th) {
  const size_t frame_px = (size_t)(th + 8 + 1) * kPW;
  // + 16 pixels of slack behind the compact b tile: partial segments of the last row read past it
  return frame_px * (2 * c_ + c_ + c_) * 2 + ((size_t)th * kTW + 16) * c_ * 2;
}

int c3k_pick_th(int h) { return h % 32 == 0 ? 32 : (h % 20 == 0 ? 20 : (h % 16 == 0 ? 16 : 0)); }

bool c3k_supported(int c, int h, int w, int in_pitch, int in_coff, int out_pitch, int out_coff) {
  if (!(c == 8 || c == 16 || c == 32)) return false;
  if (w % kTW || c3k_pick_th(h) == 0) return false;
  if (in_pitch % 8 || in_coff % 8 || out_pitch % 8 || out_coff % 8) return false;
  return c3k_smem_bytes(c / 2, c3k_pick_th(h)) <= 220 * 1024;
}

// weights (PyTorch layout, BN folded): w[0]=cv1 [c_][c], w[1]=cv2 [c_][c], w[2..5] = m0.cv1, m0.cv2, m1.cv1,
// m1.cv2 [c_][c_][3][3], w[6] = cv3 [c][2c_].  Returns packed fragments + [7][32] biases.
void c3k_pack(int c, const float *const w[7], const float *const b[7], std::vector<uint32_t> &frags, std::vector<float> &bias) {
  const int C = c / 2;
  frags.clear();
  bias.assign(7 * 32, 0.f);
  const int couts[7] = {C, C, C, C, C, C, c};
  for (int i = 0; i < 7; ++i)
    for (int n = 0; n < couts[i]; ++n) bias[i * 32 + n] = b[i][n];
  c3k_flat_pack(c, w, frags);
  if (c == 16 || c == 32) c3k_tc_pack(c, w, frags);  // the tcgen05 kernel's weight blocks follow the flat fragments
}

bool cls_branch_supported(int cin, int mid, int nc, int h, int w, int in_pitch, int in_coff, int out_pitch, int out_coff) {
  if (!(cin == 32 || cin == 64) || mid != 32 || nc < 1 || nc > 8) return false;
  if (w % kTW || h % kClsTH) return false;
  return in_pitch % 8 == 0 && in_coff % 8 == 0;
}

// w[0] dw1 [cin][1][3][3], w[1] pw1 [32][cin], w[2] dw2 [32][1][3][3], w[3] pw2 [32][32], w[4] pw3 [nc][32]
void cls_branch_pack(int cin, int nc, const float *const w[5], const float *const b[5], std::vector<__nv_bfloat16> &wdw,
                     std::vector<uint32_t> &frags, std::vector<float> &bias) {
  const int mid = 32;
  wdw.assign((size_t)9 * (cin + mid), __float2bfloat16_rn(0.f));
  for (int t = 0; t < 9; ++t) {
    for (int c = 0; c < cin; ++c) wdw[(size_t)t * cin + c] = __float2bfloat16_rn(w[0][(size_t)c * 9 + t]);
    for (int c = 0; c < mid; ++c) wdw[(size_t)9 * cin + (size_t)t * mid + c] = __float2bfloat16_rn(w[2][(size_t)c * 9 + t]);
  }
  bias.assign(5 * 128, 0.f);
  const int widths[5] = {cin, mid, mid, mid, nc};
  for (int i = 0; i < 5; ++i)
    for (int c = 0; c < widths[i]; ++c) bias[i * 128 + c] = b[i][c];
  frags.clear();
  const float *w1 = w[1], *w3 = w[3], *w4 = w[4];
  pack_frags(frags, (cin + 15) / 16, mid / 8, mid, [&](int s, int kk, int n) { return w1[(size_t)n * cin + 16 * s + kk]; });
  pack_frags(frags, mid / 16, mid / 8, mid, [&](int s, int kk, int n) { return w3[(size_t)n * mid + 16 * s + kk]; });
  pack_frags(frags, mid / 16, 1, nc, [&](int s, int kk, int n) { return w4[(size_t)n * mid + 16 * s + kk]; });
}

int cls_branch_launch(int cin, const ClsArgs &a0, cudaStream_t s) {
  ClsArgs a = a0;
  a.tiles_x = a.w / kTW;
  a.tiles_y = a.h / kClsTH;
  const int cw = cin > 32 ? cin : 32;
  const size_t smem = (size_t)(kClsTH + 5) * kPW * cw * 2 * 2;
  const unsigned grid = (unsigned)(a.n * a.tiles_x * a.tiles_y);
  if (cin == 32) {
    if (int e = smem_optin(cls_branch_fused_kernel<32>, 200 * 1024)) return e;
    cls_branch_fused_kernel<32><<<grid, kClsThreads, smem, s>>>(a);
  } else {
    if (int e = smem_optin(cls_branch_fused_kernel<64>, 200 * 1024)) return e;
    cls_branch_fused_kernel<64><<<grid, kClsThreads, smem, s>>>(a);
  }
  return (int)cudaGetLastError();
}

// Tile height for a launch: the tallest tile (least halo recompute) unless it leaves more than half of the SMs
// without a CTA -- small batches (latency runs) get short tiles so that one frame spreads over many SMs instead
// of 2-20 CTAs.  (Measured: at batch 64 a shorter tile only adds halo work: 40x40 c32 runs 33 us with TH = 20 on
// 128 CTAs and 51 us with TH = 8 on 320.)
int c3k_launch_th(int n, int h, int w) {
  const int tallest = c3k_pick_th(h);
  if (const char *v = getenv("UYD_C3K_TH")) {  // experiment hook
    const int th = atoi(v);
    if (th > 0 && th <= tallest && h % th == 0) return th;
  }
  const int cands[5] = {32, 20, 16, 8, 4};
  int best = tallest;
  for (int i = 0; i < 5; ++i) {
    const int th = cands[i];
    if (th > tallest || h % th) continue;
    best = th;
    if ((long long)n * (w / kTW) * (h / th) >= 74) break;
  }
  return best;
}

int c3k_launch(int c, const C3kArgs &a0, cudaStream_t s) {
  C3kArgs a = a0;
  a.th = c3k_launch_th(a.n, a.h, a.w);
  a.tiles_x = a.w / kTW;
  a.tiles_y = a.h / a.th;
  // tcgen05 generation: full-width strips, so it needs a batch that fills the GPU (small batches keep the 2-D tiles)
  const char *force = getenv("UYD_C3K_TC");  // 1: always, 0: never (tests / experiments)
  if ((c == 16 || c == 32) && c3k_tc_supported(c, a.h, a.w) && (force ? *force == '1' : c3k_tc_preferred(c, a.h, a.w, a.n)))
    return c3k_tc_launch(c, a, a.wfrag + c3k_flat_words(c), s);
  return c3k_flat_launch(c, a, s);
}


// INT8 variant: the extra code frame makes a tile bigger, so the tile height is chosen per shape: the tallest of
// {40, 32, 20, 16} rows that lets three, else two CTAs share an SM, else the tallest that fits at all (then 8 / 4 rows).
// Never the tcgen05 kernel.
int c3k_launch_q(int c, const C3kArgs &a0, cudaStream_t s) {
  C3kArgs a = a0;
  const int first = c3k_launch_th(a.n, a.h, a.w);
  int th = 0;
  if (const char *v = getenv("UYD_C3K_Q_TH")) {  // experiment hook
    const int t = atoi(v);
    if (t > 0 && a.h % t == 0 && c3k_flat_q_smem_bytes(c / 2, t) <= 226 * 1024) th = t;
  }
  const size_t limits[3] = {75 * 1024, 112 * 1024, 226 * 1024};
  const int tall[4] = {40, 32, 20, 16}, shorter[2] = {8, 4};
  for (int l = 0; l < 3 && !th; ++l)
    for (int i = 0; i < 4 && !th; ++i)
      if (tall[i] <= (first > 20 ? first : 20) && a.h % tall[i] == 0 && c3k_flat_q_smem_bytes(c / 2, tall[i]) <= limits[l]) th = tall[i];
  for (int i = 0; i < 2 && !th; ++i)
    if (a.h % shorter[i] == 0 && c3k_flat_q_smem_bytes(c / 2, shorter[i]) <= limits[2]) th = shorter[i];
  UYD_REQUIRE(th > 0, UYD_E_UNSUPPORTED, "fused int8 c3k: no tile height fits shared memory (c = %d, h = %d)", c, a.h);
  a.th = th;
  a.tiles_x = a.w / kTW;
  a.tiles_y = a.h / a.th;
  return c3k_flat_q_launch(c, a, s);
}

// Host side of uyd_plan_add_c3k_s8: int8 weight codes -> the flat kernel's bf16 fragments (exact), and the table the
// kernel reads behind C3kArgs::bias: [7][32] bias | [7][32] multiplier | 7 input scales (+ 1 pad).
void c3k_pack_q(int c, const int8_t *const wq[7], const float *const mult[7], const float *const bias[7], const float in_scale[7],
                std::vector<uint32_t> &frags, std::vector<float> &table) {
  const int C = c / 2;
  const int couts[7] = {C, C, C, C, C, C, c};
  const size_t sizes[7] = {(size_t)C * c, (size_t)C * c, (size_t)C * C * 9, (size_t)C * C * 9, (size_t)C * C * 9, (size_t)C * C * 9, (size_t)c * c};
  std::vector<std::vector<float>> wf(7);
  const float *wp[7];
  for (int i = 0; i < 7; ++i) {
    wf[i].resize(sizes[i]);
    for (size_t k = 0; k < sizes[i]; ++k) wf[i][k] = (float)wq[i][k];
    wp[i] = wf[i].data();
  }
  frags.clear();
  c3k_flat_pack(c, wp, frags);
  table.assign(14 * 32 + 8, 0.f);
  for (int i = 0; i < 7; ++i) {
    for (int n = 0; n < couts[i]; ++n) {
      table[i * 32 + n] = bias[i][n];
      table[7 * 32 + i * 32 + n] = mult[i][n];
    }
    table[14 * 32 + i] = in_scale[i];
  }
}

}  // namespace uyd
