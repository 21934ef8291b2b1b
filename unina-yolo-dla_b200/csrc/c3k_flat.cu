// Fused C3k block kernel, flat-frame formulation: see c3k_flat.cuh for the design and the per-lane maps.
#include "c3k_flat.cuh"

#include "common.cuh"

namespace uyd {

struct C3kArgs {  // same struct as in c3k_fused.cu / api.cu
  const __nv_bfloat16 *in;
  __nv_bfloat16 *out;
  const uint32_t *wfrag;
  const float *bias;
  int n, h, w, in_pitch, out_pitch, th, tiles_x, tiles_y;
};

namespace {
using namespace c3kf;

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int KS, int NT>
__device__ __forceinline__ void load_b(const uint32_t *wf, int lane, uint32_t (&bf)[KS][NT][2]) {
#pragma unroll
  for (int s = 0; s < KS; ++s)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const uint2 v = __ldg(reinterpret_cast<const uint2 *>(wf) + (s * NT + j) * 32 + lane);
      bf[s][j][0] = v.x;
      bf[s][j][1] = v.y;
    }
}

template <int NT>
__device__ __forceinline__ void init_acc(float (&acc)[NT][4], const float (&bz)[NT][2]) {
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    acc[j][0] = acc[j][2] = bz[j][0];
    acc[j][1] = acc[j][3] = bz[j][1];
  }
}

template <int C, bool RES>
__device__ __forceinline__ void run_conv3(const unsigned char *S, unsigned char *D, const uint32_t *wf, const float *bias,
                                          const uint32_t *maskw, int blk_lo, int blk_hi, int warp, int lane, int nwarps) {
  using G = Geo<C>;
  using St = Stage3<C>;
  uint32_t bf[St::KS][St::NT][2];
  load_b<St::KS, St::NT>(wf, lane, bf);
  float bz[St::NT][2];
  St::bias_regs(bias, lane, bz);
  for (int blk = blk_lo + warp; blk < blk_hi; blk += nwarps) {
    const int f = blk * 32;
    const uint32_t mw = maskw[blk];
#pragma unroll
    for (int mt = 0; mt < G::MT; ++mt) {
      float acc[St::NT][4];
      init_acc<St::NT>(acc, bz);
#pragma unroll
      for (int s = 0; s < St::KS; ++s) {
        uint32_t a[4];
        St::load_a(S, f, lane, s, mt, a);
#pragma unroll
        for (int j = 0; j < St::NT; ++j) mma16816(acc[j], a, bf[s][j][0], bf[s][j][1]);
      }
      St::template store<RES>(D, f, lane, mt, acc, mw);
    }
  }
}

template <int C, int NW>
__global__ void __launch_bounds__(NW * 32) c3k_flat_kernel(C3kArgs a) {
  pdl_trigger();
  using G = Geo<C>;
  extern __shared__ __align__(16) unsigned char smem[];
  const int TH = a.th;
  const Layout L = make_layout<C>(TH);
  unsigned char *X = smem + L.x_off;   // x frame [FR][2C]; later t frame [FR][C] and the staged y tile [TH*48][2C]
  unsigned char *A = smem + L.a_off;   // a -> u -> v  [FR][C]
  unsigned char *Bv = smem + L.b_off;  // b            [TH*48][C], origin = frame row 4
  uint32_t *maskw = reinterpret_cast<uint32_t *>(smem + L.mask_off);
  float *sbias = reinterpret_cast<float *>(smem + L.bias_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int n = tile / (a.tiles_x * a.tiles_y);
  const int tr = tile % (a.tiles_x * a.tiles_y);
  const int ty0 = (tr / a.tiles_x) * TH, tx0 = (tr % a.tiles_x) * kTW;
  const int gy0 = ty0 - 4, gx0 = tx0 - 4;  // image coordinates of frame pixel 0
  const int H = a.h, W = a.w;
  constexpr int NT = NW * 32;

  // ---- stage 0: input tile + halo -> X (zero outside the image).  Eight 16-byte loads per thread are issued
  // before anything else; the bias copy and the inside-image bits are computed while they are in flight. ----
  {
    constexpr int CH16 = G::CC / 8;  // 16-byte chunks per pixel
    constexpr int UN = 8;
    const __nv_bfloat16 *img = a.in + (long long)n * H * W * a.in_pitch;
    const int total = L.FR * CH16;
    for (int base = 0; base < total; base += UN * NT) {
      uint4 v[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int i = base + u * NT + tid;
        const int px = i / CH16, ch = i % CH16;
        const int ry = px / kPW, rx = px - ry * kPW;
        const int gy = gy0 + ry, gx = gx0 + rx;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (i < total && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W)
          v[u] = __ldg(reinterpret_cast<const uint4 *>(img + (unsigned)((gy * W + gx) * a.in_pitch + ch * 8)));
      }
      if (base == 0) {
        for (int i = tid; i < 7 * 32; i += NT) sbias[i] = a.bias[i];
        // inside-image bit per frame pixel (rows and columns outside the image are every conv's zero padding)
        for (int i = tid; i < L.FR / 32; i += NT) {
          uint32_t m = 0u;
          int ry = (i * 32) / kPW, rx = (i * 32) % kPW;
          for (int b = 0; b < 32; ++b) {
            if ((unsigned)(gy0 + ry) < (unsigned)H && (unsigned)(gx0 + rx) < (unsigned)W) m |= 1u << b;
            if (++rx == kPW) { rx = 0; ++ry; }
          }
          maskw[i] = m;
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int i = base + u * NT + tid;
        if (i < total) reinterpret_cast<uint4 *>(X)[i] = v[u];
      }
    }
  }
  __syncthreads();

  const uint32_t *wf = a.wfrag;
  // ---- stage 1: a on the whole frame, b on the output rows ----
  {
    using St = Stage1<C>;
    uint32_t bf[St::KS][St::NT][2];
    load_b<St::KS, St::NT>(wf, lane, bf);
    float bz[St::NT][2];
    St::bias_regs(sbias, lane, bz);
    const int b_lo = kB0 / 32, b_hi = (TH + 4) * kPW / 32;
    for (int blk = warp; blk < L.FR / 32; blk += NW) {
      const int f = blk * 32;
      const uint32_t mw = maskw[blk];
      const bool do_b = blk >= b_lo && blk < b_hi;
#pragma unroll
      for (int mt = 0; mt < G::MT; ++mt) {
        float acc[St::NT][4];
        init_acc<St::NT>(acc, bz);
#pragma unroll
        for (int s = 0; s < St::KS; ++s) {
          uint32_t af[4];
          St::load_a(X, f, lane, s, mt, af);
#pragma unroll
          for (int j = 0; j < St::NT; ++j) mma16816(acc[j], af, bf[s][j][0], bf[s][j][1]);
        }
        St::store(A, Bv, f, lane, mt, acc, mw, do_b);
      }
    }
  }
  __syncthreads();
  // ---- stages 2..5: t1 = m0.cv1(a); u = a + m0.cv2(t1); t2 = m1.cv1(u); v = u + m1.cv2(t2)  (t aliases x) ----
  unsigned char *T = X;
  int lo, hi;
  conv3_blocks(0, TH, lo, hi);
  run_conv3<C, false>(A, T, wf + G::W1, sbias + 2 * 32, maskw, lo, hi, warp, lane, NW);
  __syncthreads();
  conv3_blocks(1, TH, lo, hi);
  run_conv3<C, true>(T, A, wf + G::W1 + G::W3, sbias + 3 * 32, maskw, lo, hi, warp, lane, NW);
  __syncthreads();
  conv3_blocks(2, TH, lo, hi);
  run_conv3<C, false>(A, T, wf + G::W1 + 2 * G::W3, sbias + 4 * 32, maskw, lo, hi, warp, lane, NW);
  __syncthreads();
  conv3_blocks(3, TH, lo, hi);
  run_conv3<C, true>(T, A, wf + G::W1 + 3 * G::W3, sbias + 5 * 32, maskw, lo, hi, warp, lane, NW);
  __syncthreads();
  // ---- stage 6: y = cv3([v | b]) on the output rows -> staged in X (t is dead) ----
  {
    using St = Stage6<C>;
    uint32_t bf[St::KS][St::NT][2];
    load_b<St::KS, St::NT>(wf + G::W1 + 4 * G::W3, lane, bf);
    float bz[St::NT][2];
    St::bias_regs(sbias + 6 * 32, lane, bz);
    const int b_lo = kB0 / 32, b_hi = (TH + 4) * kPW / 32;
    for (int blk = b_lo + warp; blk < b_hi; blk += NW) {
      const int f = blk * 32;
#pragma unroll
      for (int mt = 0; mt < G::MT; ++mt) {
        float acc[St::NT][4];
        init_acc<St::NT>(acc, bz);
#pragma unroll
        for (int s = 0; s < St::KS; ++s) {
          uint32_t af[4];
          St::load_a(A, Bv, f, lane, s, mt, af);
#pragma unroll
          for (int j = 0; j < St::NT; ++j) mma16816(acc[j], af, bf[s][j][0], bf[s][j][1]);
        }
        St::store(X, f, lane, mt, acc);
      }
    }
  }
  __syncthreads();
  {
    constexpr int CH16 = G::CC / 8;
    __nv_bfloat16 *img = a.out + (long long)n * H * W * a.out_pitch;
    const int total = TH * kTW * CH16;
#pragma unroll 4
    for (int i = tid; i < total; i += NT) {
      const int px = i / CH16, ch = i % CH16;
      const int ry = px / kTW, rx = px - ry * kTW;
      *reinterpret_cast<uint4 *>(img + ((long long)(ty0 + ry) * W + tx0 + rx) * a.out_pitch + ch * 8) =
          *reinterpret_cast<const uint4 *>(X + ((ry * kPW + rx + 4) * CH16 + ch) * 16);
    }
  }
}

// =================================================================================================
// INT8 (fake-quant) C3k block in ONE launch (c3k_flat.cuh, namespace q8): the same flat-frame machine with
//   * x quantised while it is staged -- once with cv1's input scale (stage 1a: a), then, re-read from L2, with cv2's
//     (stage 1b: b on the output rows); no second x frame in shared memory;
//   * a second a-sized frame Aq for the int8 codes the 3x3 convs read, next to the bf16 frame A that carries the
//     residuals (a -> u -> v); t1 / t2 / b only exist as codes;
//   * accumulators start at zero and the epilogues apply float(acc) * m_c + b_c (tables [7][32] bias | [7][32]
//     multiplier | 7 input scales behind a.bias).
// Replaces 7 conv_s8 launches + 4-5 quantize launches per block (uyd_plan_add_c3k_s8).
// =================================================================================================
struct LayoutQ {
  int FR, x_off, a_off, aq_off, b_off, mask_off, tab_off, total;
};
template <int C>
C3K_HD LayoutQ make_layout_q(int TH) {
  using G = Geo<C>;
  LayoutQ l;
  l.FR = (TH + 8) * kPW;
  l.x_off = 0;
  l.a_off = l.x_off + l.FR * G::PXX;
  l.aq_off = l.a_off + l.FR * G::PXA;
  l.b_off = l.aq_off + l.FR * G::PXA;
  l.mask_off = l.b_off + (TH * kPW + 32) * G::PXA;  // + 32 pixels: over-reads of the Aq frame's last block end here
  l.tab_off = l.mask_off + (l.FR / 32) * 4;
  l.total = l.tab_off + (14 * 32 + 8) * 4;
  return l;
}
constexpr int kQTab = 14 * 32 + 8;  // floats behind C3kArgs::bias for the INT8 variant

template <int C, bool RES>
__device__ __forceinline__ void run_conv3_q(const unsigned char *S, unsigned char *Dr, unsigned char *Dq, const uint32_t *wf, const float *bias,
                                            const float *mult, float s_next, const uint32_t *maskw, int blk_lo, int blk_hi, int warp,
                                            int lane, int nwarps) {
  using G = Geo<C>;
  using St = Stage3<C>;
  uint32_t bf[St::KS][St::NT][2];
  load_b<St::KS, St::NT>(wf, lane, bf);
  float bz[St::NT][2], mz[St::NT][2];
  St::bias_regs(bias, lane, bz);
  St::bias_regs(mult, lane, mz);
  for (int blk = blk_lo + warp; blk < blk_hi; blk += nwarps) {
    const int f = blk * 32;
    const uint32_t mw = maskw[blk];
#pragma unroll
    for (int mt = 0; mt < G::MT; ++mt) {
      float acc[St::NT][4];
#pragma unroll
      for (int j = 0; j < St::NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
      for (int s = 0; s < St::KS; ++s) {
        uint32_t a[4];
        St::load_a(S, f, lane, s, mt, a);
#pragma unroll
        for (int j = 0; j < St::NT; ++j) mma16816(acc[j], a, bf[s][j][0], bf[s][j][1]);
      }
      q8::store_3<C, RES>(Dr, Dq, f, lane, mt, acc, bz, mz, mw, s_next);
    }
  }
}

template <int C, int NW>
__global__ void __launch_bounds__(NW * 32) c3k_flat_q_kernel(C3kArgs a) {
  pdl_trigger();
  using G = Geo<C>;
  extern __shared__ __align__(16) unsigned char smem[];
  const int TH = a.th;
  const LayoutQ L = make_layout_q<C>(TH);
  unsigned char *X = smem + L.x_off;    // int8 codes of x [FR][2C] (cv1's scale, then cv2's); later t1 / t2 and the staged y tile
  unsigned char *A = smem + L.a_off;    // bf16 a -> u -> v  [FR][C]: the residuals
  unsigned char *Aq = smem + L.aq_off;  // their int8 codes for m0.cv1 / m1.cv1 / cv3
  unsigned char *Bv = smem + L.b_off;   // codes of b for cv3, origin = frame row 4
  uint32_t *maskw = reinterpret_cast<uint32_t *>(smem + L.mask_off);
  float *sbias = reinterpret_cast<float *>(smem + L.tab_off), *smult = sbias + 7 * 32, *sscale = sbias + 14 * 32;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int n = tile / (a.tiles_x * a.tiles_y);
  const int tr = tile % (a.tiles_x * a.tiles_y);
  const int ty0 = (tr / a.tiles_x) * TH, tx0 = (tr % a.tiles_x) * kTW;
  const int gy0 = ty0 - 4, gx0 = tx0 - 4;
  const int H = a.h, W = a.w;
  constexpr int NT = NW * 32;
  const float s_cv1 = a.bias[14 * 32 + 0], s_cv2 = a.bias[14 * 32 + 1];

  // x tile + halo -> codes for input scale s (zero outside the image: the code of the zero padding)
  auto stage_x = [&](float s, bool first) {
    constexpr int CH16 = G::CC / 8, UN = 8;
    const __nv_bfloat16 *img = a.in + (long long)n * H * W * a.in_pitch;
    // the second pass (cv2's scale) only feeds stage 1b, which reads the output rows: frame rows 4 .. TH + 3
    const int total = (first ? L.FR : (TH + 4) * kPW) * CH16;
    for (int base = first ? 0 : 4 * kPW * CH16; base < total; base += UN * NT) {
      uint4 v[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int i = base + u * NT + tid;
        const int px = i / CH16, ch = i % CH16;
        const int ry = px / kPW, rx = px - ry * kPW;
        const int gy = gy0 + ry, gx = gx0 + rx;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (i < total && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W)
          v[u] = __ldg(reinterpret_cast<const uint4 *>(img + (unsigned)((gy * W + gx) * a.in_pitch + ch * 8)));
      }
      if (first && base == 0) {
        for (int i = tid; i < kQTab; i += NT) sbias[i] = a.bias[i];
        for (int i = tid; i < L.FR / 32; i += NT) {
          uint32_t m = 0u;
          int ry = (i * 32) / kPW, rx = (i * 32) % kPW;
          for (int b = 0; b < 32; ++b) {
            if ((unsigned)(gy0 + ry) < (unsigned)H && (unsigned)(gx0 + rx) < (unsigned)W) m |= 1u << b;
            if (++rx == kPW) { rx = 0; ++ry; }
          }
          maskw[i] = m;
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int i = base + u * NT + tid;
        if (i < total)
          reinterpret_cast<uint4 *>(X)[i] = make_uint4(q8::qpack(v[u].x, s), q8::qpack(v[u].y, s), q8::qpack(v[u].z, s), q8::qpack(v[u].w, s));
      }
    }
  };
  stage_x(s_cv1, true);
  __syncthreads();

  const uint32_t *wf = a.wfrag;
  using S1 = Stage1<C>;
  constexpr int NA = S1::NT / 2;  // tiles of the a half; the b half follows
  const int b_lo = kB0 / 32, b_hi = (TH + 4) * kPW / 32;
  {
    // ---- stage 1a: a = relu(cv1 q1(x)) on the whole frame -> A (bf16) and Aq (codes for m0.cv1) ----
    uint32_t bf[S1::KS][S1::NT][2];
    load_b<S1::KS, S1::NT>(wf, lane, bf);
    float bz[S1::NT][2], mz[S1::NT][2];
    S1::bias_regs(sbias, lane, bz);
    S1::bias_regs(smult, lane, mz);
    const float s_next = sscale[2];
    for (int blk = warp; blk < L.FR / 32; blk += NW) {
      const int f = blk * 32;
      const uint32_t mw = maskw[blk];
#pragma unroll
      for (int mt = 0; mt < G::MT; ++mt) {
        float acc[S1::NT][4];
#pragma unroll
        for (int j = 0; j < S1::NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
        for (int s = 0; s < S1::KS; ++s) {
          uint32_t af[4];
          S1::load_a(X, f, lane, s, mt, af);
#pragma unroll
          for (int j = 0; j < NA; ++j) mma16816(acc[j], af, bf[s][j][0], bf[s][j][1]);
        }
        q8::store_a<C>(A, Aq, f, lane, mt, acc, bz, mz, mw, s_next);
      }
    }
    __syncthreads();
    // ---- stage 1b: x again (L2), coded with cv2's scale; b = relu(cv2 q2(x)) on the output rows -> Bv (codes for cv3) ----
    stage_x(s_cv2, false);
    __syncthreads();
    const float s_cv3 = sscale[6];
    for (int blk = b_lo + warp; blk < b_hi; blk += NW) {
      const int f = blk * 32;
#pragma unroll
      for (int mt = 0; mt < G::MT; ++mt) {
        float acc[S1::NT][4];
#pragma unroll
        for (int j = 0; j < S1::NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
        for (int s = 0; s < S1::KS; ++s) {
          uint32_t af[4];
          S1::load_a(X, f, lane, s, mt, af);
#pragma unroll
          for (int j = NA; j < S1::NT; ++j) mma16816(acc[j], af, bf[s][j][0], bf[s][j][1]);
        }
        q8::store_b<C>(Bv, f, lane, mt, acc, bz, mz, s_cv3);
      }
    }
  }
  __syncthreads();
  // ---- stages 2..5: t1 = m0.cv1(a); u = a + m0.cv2(t1); t2 = m1.cv1(u); v = u + m1.cv2(t2)  (t aliases x) ----
  unsigned char *T = X;
  int lo, hi;
  conv3_blocks(0, TH, lo, hi);
  run_conv3_q<C, false>(Aq, nullptr, T, wf + G::W1, sbias + 2 * 32, smult + 2 * 32, sscale[3], maskw, lo, hi, warp, lane, NW);
  __syncthreads();
  conv3_blocks(1, TH, lo, hi);
  run_conv3_q<C, true>(T, A, Aq, wf + G::W1 + G::W3, sbias + 3 * 32, smult + 3 * 32, sscale[4], maskw, lo, hi, warp, lane, NW);
  __syncthreads();
  conv3_blocks(2, TH, lo, hi);
  run_conv3_q<C, false>(Aq, nullptr, T, wf + G::W1 + 2 * G::W3, sbias + 4 * 32, smult + 4 * 32, sscale[5], maskw, lo, hi, warp, lane, NW);
  __syncthreads();
  conv3_blocks(3, TH, lo, hi);
  run_conv3_q<C, true>(T, A, Aq, wf + G::W1 + 3 * G::W3, sbias + 5 * 32, smult + 5 * 32, sscale[6], maskw, lo, hi, warp, lane, NW);
  __syncthreads();
  // ---- stage 6: y = cv3([v | b]) on the output rows -> staged in X (t is dead) ----
  {
    using St = Stage6<C>;
    uint32_t bf[St::KS][St::NT][2];
    load_b<St::KS, St::NT>(wf + G::W1 + 4 * G::W3, lane, bf);
    float bz[St::NT][2], mz[St::NT][2];
    St::bias_regs(sbias + 6 * 32, lane, bz);
    St::bias_regs(smult + 6 * 32, lane, mz);
    for (int blk = b_lo + warp; blk < b_hi; blk += NW) {
      const int f = blk * 32;
#pragma unroll
      for (int mt = 0; mt < G::MT; ++mt) {
        float acc[St::NT][4];
#pragma unroll
        for (int j = 0; j < St::NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
        for (int s = 0; s < St::KS; ++s) {
          uint32_t af[4];
          St::load_a(Aq, Bv, f, lane, s, mt, af);
#pragma unroll
          for (int j = 0; j < St::NT; ++j) mma16816(acc[j], af, bf[s][j][0], bf[s][j][1]);
        }
        q8::store_y<C>(X, f, lane, mt, acc, bz, mz);
      }
    }
  }
  __syncthreads();
  {
    constexpr int CH16 = G::CC / 8;
    __nv_bfloat16 *img = a.out + (long long)n * H * W * a.out_pitch;
    const int total = TH * kTW * CH16;
#pragma unroll 4
    for (int i = tid; i < total; i += NT) {
      const int px = i / CH16, ch = i % CH16;
      const int ry = px / kTW, rx = px - ry * kTW;
      *reinterpret_cast<uint4 *>(img + ((long long)(ty0 + ry) * W + tx0 + rx) * a.out_pitch + ch * 8) =
          *reinterpret_cast<const uint4 *>(X + ((ry * kPW + rx + 4) * CH16 + ch) * 16);
    }
  }
}

}  // namespace

size_t c3k_flat_smem_bytes(int c_, int th) {
  return c_ == 4 ? make_layout<4>(th).total : (c_ == 8 ? make_layout<8>(th).total : make_layout<16>(th).total);
}

int c3k_flat_words(int c) { return c == 8 ? Geo<4>::WORDS : (c == 16 ? Geo<8>::WORDS : Geo<16>::WORDS); }

void c3k_flat_pack(int c, const float *const w[7], std::vector<uint32_t> &frags) {
  if (c == 8) pack_all<4>(w, frags);
  else if (c == 16) pack_all<8>(w, frags);
  else pack_all<16>(w, frags);
}

// a.th / tiles_x / tiles_y are set by the caller (c3k_launch); a.wfrag points at the flat-layout fragments.
int c3k_flat_launch(int c, const C3kArgs &a, cudaStream_t s) {
  const size_t smem = c3k_flat_smem_bytes(c / 2, a.th);
  const unsigned grid = (unsigned)(a.n * a.tiles_x * a.tiles_y);
  auto go = [&](auto kern, int, int threads) -> int {
    if (int e = smem_optin(kern, 227 * 1024)) return e;
    kern<<<grid, threads, smem, s>>>(a);
    return (int)cudaGetLastError();
  };
  if (c == 8) return go(c3k_flat_kernel<4, 8>, 0, 256);
  if (c == 16) return go(c3k_flat_kernel<8, 8>, 1, 256);
  return go(c3k_flat_kernel<16, 16>, 2, 512);
}


size_t c3k_flat_q_smem_bytes(int c_, int th) {
  return c_ == 4 ? make_layout_q<4>(th).total : (c_ == 8 ? make_layout_q<8>(th).total : make_layout_q<16>(th).total);
}

// a.th / tiles_x / tiles_y are set by the caller (c3k_launch_q); a.bias points at the kQTab-float table.
int c3k_flat_q_launch(int c, const C3kArgs &a, cudaStream_t s) {
  const size_t smem = c3k_flat_q_smem_bytes(c / 2, a.th);
  const unsigned grid = (unsigned)(a.n * a.tiles_x * a.tiles_y);
  auto go = [&](auto kern, int threads) -> int {
    if (int e = smem_optin(kern, 227 * 1024)) return e;
    kern<<<grid, threads, smem, s>>>(a);
    return (int)cudaGetLastError();
  };
  if (c == 8) return go(c3k_flat_q_kernel<4, 8>, 256);
  if (c == 16) return go(c3k_flat_q_kernel<8, 8>, 256);   // 109 registers: two 256-thread CTAs per SM, not one of 512
  return go(c3k_flat_q_kernel<16, 16>, 512);
}

}  // namespace uyd
