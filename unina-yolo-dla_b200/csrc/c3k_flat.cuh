// Fused C3k block, "flat frame" formulation (second generation of c3k_fused.cu).
//
//     a  = relu(cv1 x)            1x1  c  -> c_          (c_ = c / 2 = 4, 8 or 16)
//     b  = relu(cv2 x)            1x1  c  -> c_
//     t1 = relu(m0.cv1 a)         3x3      u  = a + relu(m0.cv2 t1)   3x3
//     t2 = relu(m1.cv1 u)         3x3      v  = u + relu(m1.cv2 t2)   3x3
//     y  = relu(cv3 [v | b])      1x1  2c_ -> c
//
// ncu of the first-generation kernel: 2.6 % of the issued warp instructions were HMMA, 38 % integer index
// arithmetic, predicates and branches (per 16-pixel segment: row/column division, nine tap offsets, bounds
// tests).  This formulation removes the index arithmetic instead of hiding it:
//
//   * The tile + 4-pixel halo is a FLAT array of (TH + 8) x 48 pixels.  A 3x3 tap is a constant offset
//     dy * 48 + dx in that array, a stage is a loop over 32-pixel blocks of it, and every address in the loop
//     is  block base + per-thread constant + immediate.  Blocks ignore row ends: the wrapped-around columns
//     compute garbage that no valid output reads (each stage's valid region shrinks by one pixel, and a
//     mma row only depends on its own pixel's window).
//   * "pixel outside the image -> 0" (each conv zero-pads ITS input) is one bit per frame pixel, built once.
//   * Every 16-column k-step is permuted so that thread t owns the four physical columns 4t..4t+3: one
//     LDS.64 per fragment half instead of two LDS.32; output channels are permuted the same way so the
//     epilogue stores 8 or 16 contiguous bytes per thread.
//   * c_ = 4 (the 160 x 160 blocks, the most expensive ones): an mma row is a PAIR of horizontally adjacent
//     pixels.  K = 3 rows x (4 columns x 4 channels) = 3 full k-steps, N = 2 pixels x 4 channels = 8: no
//     padding in either dimension (the first generation padded N 4 -> 8 and K 36 -> 48 per pixel).
//   * cv1 and cv2 read the same pixels: one pass with N = [a | b].  Biases start the accumulators.
//
// The per-lane address and fragment maps below are shared with a host emulation of the warp
// (tests/c3k_emu.cpp) that checks them against torch on the CPU: test infrastructure, not a fallback.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#if defined(__CUDACC__)
#define C3K_HD __host__ __device__ __forceinline__
#else
#include <cuda_runtime.h>  // uint2 / uint4 for the host emulation
#define C3K_HD inline
#endif

namespace uyd {
namespace c3kf {

constexpr int kPW = 48;   // frame row pitch in pixels = 40 + 2 * 4
constexpr int kTW = 40;   // output tile width
constexpr int kB0 = 4 * kPW;  // flat index of the first output row (frame row 4): origin of the b / staging tiles

// ---- bf16 bit helpers ----------------------------------------------------------------------------
C3K_HD float bf16_lo(uint32_t v) {
  const uint32_t u = v << 16;
  float f;
#if defined(__CUDA_ARCH__)
  f = __uint_as_float(u);
#else
  memcpy(&f, &u, 4);
#endif
  return f;
}
C3K_HD float bf16_hi(uint32_t v) {
  const uint32_t u = v & 0xffff0000u;
  float f;
#if defined(__CUDA_ARCH__)
  f = __uint_as_float(u);
#else
  memcpy(&f, &u, 4);
#endif
  return f;
}
inline uint16_t host_f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
C3K_HD uint32_t pack_bf16(float lo, float hi) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
#else
  return (uint32_t)host_f2bf(lo) | ((uint32_t)host_f2bf(hi) << 16);
#endif
}
C3K_HD uint32_t relu_pack_bf16(float lo, float hi) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
#else
  return pack_bf16(lo > 0.f ? lo : 0.f, hi > 0.f ? hi : 0.f);
#endif
}
C3K_HD float relu(float v) { return v > 0.f ? v : 0.f; }

C3K_HD uint2 ld64(const unsigned char *p) { return *reinterpret_cast<const uint2 *>(p); }
C3K_HD uint32_t ld32(const unsigned char *p) { return *reinterpret_cast<const uint32_t *>(p); }
C3K_HD void st32(unsigned char *p, uint32_t v) { *reinterpret_cast<uint32_t *>(p) = v; }
C3K_HD void st64(unsigned char *p, uint32_t x, uint32_t y) { *reinterpret_cast<uint2 *>(p) = make_uint2(x, y); }
C3K_HD void st128(unsigned char *p, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  *reinterpret_cast<uint4 *>(p) = make_uint4(x, y, z, w);
}

// Logical k column of an m16n8k16 A/B fragment -> physical column: thread t holds logical columns
// {2t, 2t+1} (a0/a1, b0) and {2t+8, 2t+9} (a2/a3, b1) = physical 4t .. 4t+3 (one 8-byte load).
C3K_HD int phys_col(int kk) { return 4 * ((kk & 7) >> 1) + 2 * (kk >> 3) + (kk & 1); }

// ldmatrix.x4 A fragments (wherever the 16-byte chunks of a row are 16-byte aligned): lane L supplies the address of
// row (L & 7) + 8 ((L >> 3) & 1), chunk L >> 4 (chunk 0 = logical k 0..7, chunk 1 = k 8..15), and the four result
// registers ARE a0..a3 -- no LDS.64 pair to interleave with register moves.  Logical k = physical column with this path.
// Used by the c_ = 8 stages whose rows are 16 bytes apart (80x80 blocks: 38.6 -> 36.9 us).
C3K_HD int ldsm_row(int lane) { return (lane & 7) + 8 * ((lane >> 3) & 1); }
C3K_HD int ldsm_chunk(int lane) { return lane >> 4; }
// rowfn(L) = the address lane L supplies
template <class F>
C3K_HD void ldsm_a(F rowfn, int lane, uint32_t (&a)[4]) {
#if defined(__CUDA_ARCH__)
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(rowfn(lane));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
               : "r"(addr));
#else
  const int g = lane >> 2, t = lane & 3;  // host emulation: register i = word t of the row lane 8 i + g points at
  for (int i = 0; i < 4; ++i) a[i] = ld32(rowfn(8 * i + g) + 4 * t);
#endif
}

template <int C>
struct Geo {
  static_assert(C == 4 || C == 8 || C == 16, "c_ in {4, 8, 16}");
  static constexpr bool PAIR = C == 4;
  static constexpr int CC = 2 * C;
  static constexpr int PXA = C * 2;   // bytes per pixel of a / t / b
  static constexpr int PXX = CC * 2;  // bytes per pixel of x / staged y
  static constexpr int MT = PAIR ? 1 : 2;            // mma row tiles per 32-pixel block
  static constexpr int ROWA = PAIR ? 2 * PXA : PXA;  // bytes between mma rows (a / t / b)
  static constexpr int ROWX = PAIR ? 2 * PXX : PXX;  // bytes between mma rows (x / y)
  static constexpr int K1 = C == 16 ? 2 : 1, N1 = C == 16 ? 4 : 2;  // stage 1: [a | b] = cv1|cv2 (x)
  static constexpr int K3 = C == 4 ? 3 : (C == 8 ? 5 : 9), N3 = C == 16 ? 2 : 1;  // 3x3 stages
  static constexpr int K6 = C == 16 ? 2 : 1, N6 = C == 16 ? 4 : 2;  // stage 6: y = cv3([v | b])
  static constexpr int W1 = K1 * N1 * 64, W3 = K3 * N3 * 64, W6 = K6 * N6 * 64;  // fragment words per stage
  static constexpr int WORDS = W1 + 4 * W3 + W6;
};

struct Layout {
  int FR;  // frame pixels
  int x_off, a_off, b_off, mask_off, bias_off, total;  // bytes; t and the staged output alias x
};
template <int C>
C3K_HD Layout make_layout(int TH) {
  using G = Geo<C>;
  Layout l;
  l.FR = (TH + 8) * kPW;
  l.x_off = 0;
  l.a_off = l.x_off + l.FR * G::PXX;
  l.b_off = l.a_off + l.FR * G::PXA;
  l.mask_off = l.b_off + (TH * kPW + 32) * G::PXA;  // + 32 pixels: over-reads of the a frame's last block end here
  l.bias_off = l.mask_off + (l.FR / 32) * 4;
  l.total = l.bias_off + 7 * 32 * 4;
  return l;
}

// Block range [lo, hi) (in 32-pixel blocks) of the 3x3 stage k = 0..3 (valid rows k+1 .. TH+7-k).
C3K_HD void conv3_blocks(int k, int TH, int &lo, int &hi) {
  lo = ((k + 1) * kPW) / 32;
  hi = ((TH + 7 - k) * kPW + 31) / 32;
}

// ---- stage 1: [a | b] = relu(W [x]) ----------------------------------------------------------------
template <int C>
struct Stage1 {
  using G = Geo<C>;
  static constexpr int KS = G::K1, NT = G::N1;
  static constexpr bool LDSM = false;  // rows are 32 / 64 bytes apart: an 8-row ldmatrix phase would hit every bank twice
  // A fragment of k-step s, row tile mt: a[0]/a[2] rows g, a[1]/a[3] rows g + 8
  C3K_HD static void load_a(const unsigned char *X, int f, int lane, int s, int mt, uint32_t (&a)[4]) {
    const int g = lane >> 2, t = lane & 3;
    const unsigned char *p = X + f * G::PXX + g * G::ROWX + 8 * t + 32 * s + mt * 16 * G::PXX;
    const uint2 lo = ld64(p), hi = ld64(p + 8 * G::ROWX);
    a[0] = lo.x; a[2] = lo.y; a[1] = hi.x; a[3] = hi.y;
  }
  C3K_HD static void bias_regs(const float *bias, int lane, float (&bz)[NT][2]) {  // bias = [7][32]
    const int t = lane & 3;
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (C == 4) bz[j][e] = bias[j * 32 + 2 * (t & 1) + e];
        else if (C == 8) bz[j][e] = bias[j * 32 + 2 * t + e];
        else bz[j][e] = bias[(j >> 1) * 32 + 4 * t + 2 * (j & 1) + e];
      }
  }
  // mw: inside-image bits of the block's 32 pixels; do_b: the block lies in the output rows
  C3K_HD static void store(unsigned char *A, unsigned char *Bv, int f, int lane, int mt, const float (&acc)[NT][4], uint32_t mw,
                           bool do_b) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (C == 4) {
        const int bit = 2 * g + (t >> 1) + 16 * h;
        const int o = (f + 2 * g + 16 * h) * G::PXA + 4 * t;
        const uint32_t va = relu_pack_bf16(acc[0][2 * h], acc[0][2 * h + 1]);
        st32(A + o, (mw >> bit) & 1u ? va : 0u);
        if (do_b) st32(Bv + o - kB0 * G::PXA, relu_pack_bf16(acc[1][2 * h], acc[1][2 * h + 1]));
      } else if (C == 8) {
        const int bit = 16 * mt + 8 * h + g;
        const int o = (f + bit) * G::PXA + 4 * t;
        const uint32_t va = relu_pack_bf16(acc[0][2 * h], acc[0][2 * h + 1]);
        st32(A + o, (mw >> bit) & 1u ? va : 0u);
        if (do_b) st32(Bv + o - kB0 * G::PXA, relu_pack_bf16(acc[1][2 * h], acc[1][2 * h + 1]));
      } else {
        const int bit = 16 * mt + 8 * h + g;
        const int o = (f + bit) * G::PXA + 8 * t;
        const bool in = (mw >> bit) & 1u;
        const uint32_t v0 = relu_pack_bf16(acc[0][2 * h], acc[0][2 * h + 1]), v1 = relu_pack_bf16(acc[1][2 * h], acc[1][2 * h + 1]);
        st64(A + o, in ? v0 : 0u, in ? v1 : 0u);
        if (do_b)
          st64(Bv + o - kB0 * G::PXA, relu_pack_bf16(acc[2][2 * h], acc[2][2 * h + 1]), relu_pack_bf16(acc[3][2 * h], acc[3][2 * h + 1]));
      }
    }
  }
  // weight of physical column p of k-step s for fragment column n (= lane >> 2) of tile j; w0 = cv1, w1 = cv2, [C][2C]
  static float weight(const float *w0, const float *w1, int s, int p, int j, int n) {
    constexpr int CC = G::CC;
    if (C == 4) {
      const int qk = p >> 3, ci = p & 7, q = n >> 2, co = n & 3;
      return qk == q ? (j ? w1 : w0)[co * CC + ci] : 0.f;
    } else if (C == 8) {
      return (j ? w1 : w0)[n * CC + p];
    } else {
      const int jj = j & 1, co = 4 * (n >> 1) + 2 * jj + (n & 1);
      return (j >> 1 ? w1 : w0)[co * CC + 16 * s + p];
    }
  }
};

// ---- 3x3 stages: dst = relu(W * src) (+ res), zero outside the image --------------------------------
template <int C>
struct Stage3 {
  using G = Geo<C>;
  static constexpr int KS = G::K3, NT = G::N3;
  // pixel offset of (k-step s, slot) for c_ = 8: steps 0..2 = row s-1, columns (-1, 0); step 3 = column +1 of rows
  // (-1, 0); step 4 = (+1, +1) and a zero-weight slot that re-reads the same pixel.
  C3K_HD static int tap_px8(int s, int slot) {
    return s <= 2 ? (s - 1) * kPW - 1 + slot : (s == 3 ? -kPW + 1 + kPW * slot : kPW + 1);
  }
  // ldmatrix where consecutive rows are 16 bytes apart (c_ = 8): the eight rows of a phase cover all banks once.
  // c_ = 4: the window of a pixel pair starts at an odd pixel (8-byte aligned only); c_ = 16: 32-byte rows would
  // conflict two-way (measured: 23 -> 29 us).
  static constexpr bool LDSM = C == 8;
  C3K_HD static void load_a(const unsigned char *S, int f, int lane, int s, int mt, uint32_t (&a)[4]) {
    if (C == 8) {  // chunk = slot: the two pixels of the k-step
      ldsm_a([=](int L) { return S + (f + 16 * mt + ldsm_row(L) + tap_px8(s, ldsm_chunk(L))) * G::PXA; }, lane, a);
    } else {
      const int g = lane >> 2, t = lane & 3;
      const unsigned char *p = C == 4 ? S + (f + 2 * g - 1 + (s - 1) * kPW) * G::PXA + 8 * t
                                      : S + (f + 16 * mt + g + (s / 3 - 1) * kPW + s % 3 - 1) * G::PXA + 8 * t;
      const uint2 lo = ld64(p), hi = ld64(p + 8 * G::ROWA);
      a[0] = lo.x; a[2] = lo.y; a[1] = hi.x; a[3] = hi.y;
    }
  }
  C3K_HD static void bias_regs(const float *bias, int lane, float (&bz)[NT][2]) {  // bias = this conv's [32]
    const int t = lane & 3;
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (C == 4) bz[j][e] = bias[2 * (t & 1) + e];
        else if (C == 8) bz[j][e] = bias[2 * t + e];
        else bz[j][e] = bias[4 * t + 2 * j + e];
      }
  }
  template <bool RES>
  C3K_HD static void store(unsigned char *D, int f, int lane, int mt, const float (&acc)[NT][4], uint32_t mw) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (C == 4 || C == 8) {
        const int bit = C == 4 ? 2 * g + (t >> 1) + 16 * h : 16 * mt + 8 * h + g;
        const int o = C == 4 ? (f + 2 * g + 16 * h) * G::PXA + 4 * t : (f + bit) * G::PXA + 4 * t;
        uint32_t v;
        if (RES) {
          const uint32_t r = ld32(D + o);
          v = pack_bf16(relu(acc[0][2 * h]) + bf16_lo(r), relu(acc[0][2 * h + 1]) + bf16_hi(r));
        } else {
          v = relu_pack_bf16(acc[0][2 * h], acc[0][2 * h + 1]);
        }
        st32(D + o, (mw >> bit) & 1u ? v : 0u);
      } else {
        const int bit = 16 * mt + 8 * h + g;
        const int o = (f + bit) * G::PXA + 8 * t;
        uint32_t v0, v1;
        if (RES) {
          const uint2 r = ld64(D + o);
          v0 = pack_bf16(relu(acc[0][2 * h]) + bf16_lo(r.x), relu(acc[0][2 * h + 1]) + bf16_hi(r.x));
          v1 = pack_bf16(relu(acc[1][2 * h]) + bf16_lo(r.y), relu(acc[1][2 * h + 1]) + bf16_hi(r.y));
        } else {
          v0 = relu_pack_bf16(acc[0][2 * h], acc[0][2 * h + 1]);
          v1 = relu_pack_bf16(acc[1][2 * h], acc[1][2 * h + 1]);
        }
        const bool in = (mw >> bit) & 1u;
        st64(D + o, in ? v0 : 0u, in ? v1 : 0u);
      }
    }
  }
  // w = [C][C][3][3]
  static float weight(const float *w, int s, int p, int j, int n) {
    if (C == 4) {
      const int col = p >> 2, ci = p & 3, q = n >> 2, co = n & 3, dx = col - 1 - q;
      return dx >= -1 && dx <= 1 ? w[(co * C + ci) * 9 + s * 3 + dx + 1] : 0.f;
    } else if (C == 8) {
      const int slot = p >> 3, ci = p & 7;
      int ky, kx;
      if (s <= 2) { ky = s; kx = slot; }
      else if (s == 3) { ky = slot; kx = 2; }
      else { if (slot) return 0.f; ky = 2; kx = 2; }
      return w[(n * C + ci) * 9 + ky * 3 + kx];
    } else {
      const int co = 4 * (n >> 1) + 2 * j + (n & 1);
      return w[(co * C + p) * 9 + s];
    }
  }
};

// ---- stage 6: y = relu(cv3 [v | b]) -> staged output tile -------------------------------------------
template <int C>
struct Stage6 {
  using G = Geo<C>;
  static constexpr int KS = G::K6, NT = G::N6;
  // V = the a frame (holds v), Bv = b tile (origin kB0)
  static constexpr bool LDSM = C == 8;
  C3K_HD static void load_a(const unsigned char *V, const unsigned char *Bv, int f, int lane, int s, int mt, uint32_t (&a)[4]) {
    if (C == 8) {  // chunk 0 = the pixel's 16 bytes of v, chunk 1 = of b
      ldsm_a([=](int L) { return (ldsm_chunk(L) == 0 ? V + f * G::PXA : Bv + (f - kB0) * G::PXA) + (16 * mt + ldsm_row(L)) * G::PXA; }, lane, a);
    } else {
      const int g = lane >> 2, t = lane & 3;
      const unsigned char *p;
      if (C == 4) p = (t < 2 ? V + f * G::PXA : Bv + (f - kB0) * G::PXA) + g * G::ROWA + 8 * (t & 1);
      else p = (s == 0 ? V + f * G::PXA : Bv + (f - kB0) * G::PXA) + (16 * mt + g) * G::PXA + 8 * t;
      const uint2 lo = ld64(p), hi = ld64(p + 8 * G::ROWA);
      a[0] = lo.x; a[2] = lo.y; a[1] = hi.x; a[3] = hi.y;
    }
  }
  C3K_HD static void bias_regs(const float *bias, int lane, float (&bz)[NT][2]) {  // bias = cv3's [32]
    const int t = lane & 3;
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (C == 4) bz[j][e] = bias[4 * (t & 1) + 2 * j + e];
        else if (C == 8) bz[j][e] = bias[4 * t + 2 * j + e];
        else bz[j][e] = bias[8 * t + 2 * j + e];
      }
  }
  C3K_HD static void store(unsigned char *Y, int f, int lane, int mt, const float (&acc)[NT][4]) {  // Y origin kB0, [px][2C]
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (C == 4) {
        st64(Y + (f - kB0 + 2 * g + 16 * h) * G::PXX + 8 * t, relu_pack_bf16(acc[0][2 * h], acc[0][2 * h + 1]),
             relu_pack_bf16(acc[1][2 * h], acc[1][2 * h + 1]));
      } else if (C == 8) {
        st64(Y + (f - kB0 + 16 * mt + 8 * h + g) * G::PXX + 8 * t, relu_pack_bf16(acc[0][2 * h], acc[0][2 * h + 1]),
             relu_pack_bf16(acc[1][2 * h], acc[1][2 * h + 1]));
      } else {
        st128(Y + (f - kB0 + 16 * mt + 8 * h + g) * G::PXX + 16 * t, relu_pack_bf16(acc[0][2 * h], acc[0][2 * h + 1]),
              relu_pack_bf16(acc[1][2 * h], acc[1][2 * h + 1]), relu_pack_bf16(acc[2][2 * h], acc[2][2 * h + 1]),
              relu_pack_bf16(acc[3][2 * h], acc[3][2 * h + 1]));
      }
    }
  }
  // w = cv3 [2C][2C], input channels [v | b]
  static float weight(const float *w, int s, int p, int j, int n) {
    constexpr int CC = G::CC;
    if (C == 4) {
      const int src = p >> 3, qk = (p >> 2) & 1, ci = p & 3, q = n >> 2, co = 4 * ((n >> 1) & 1) + 2 * j + (n & 1);
      return qk == q ? w[co * CC + src * 4 + ci] : 0.f;
    } else if (C == 8) {
      const int co = 4 * (n >> 1) + 2 * j + (n & 1);
      return w[co * CC + p];
    } else {
      const int co = 8 * (n >> 1) + 2 * j + (n & 1);
      return w[co * CC + 16 * s + p];
    }
  }
};

// ---- INT8 (fake-quant, qat.py:109-124) variant of the block: exact integers in bf16 containers ---------------------------
// Every conv input of the QAT graph is q = clamp(rne(x * s), -127, 127) and every weight an int8 code.  Both are exact in
// bf16, and with K * 127^2 < 2^24 (K <= 144 here) the fp32 accumulation of mma.sync is exact too: acc IS the int32 sum of
// the integer convolution.  The epilogues then follow the integer reference step by step: y = float(acc) * m_c + b_c
// (separate round-to-nearest multiply and add), ReLU, + residual (the bf16 activation, added in fp32), ONE rounding to
// bf16 -- the activation the graph defines -- and, for the conv that consumes it, its int8 code stored as a bf16 value.
#if defined(__CUDACC__)
namespace q8 {
__device__ __forceinline__ float rq(float acc, float m, float b) { return __fadd_rn(__fmul_rn(acc, m), b); }
// clamp(rne(v)) == rne(clamp(v)) for the integer bounds +-127, and for |v| <= 127 the round-to-nearest-even is two fp32
// additions with 1.5 * 2^23 (FADD pipe, 128 / clk / SM) instead of cvt.rni (XU pipe, 16 / clk / SM)
__device__ __forceinline__ float qv(float y, float s) {
  const float v = fminf(fmaxf(__fmul_rn(y, s), -127.f), 127.f);
  return __fadd_rn(__fadd_rn(v, 12582912.f), -12582912.f);
}
// two bf16 activations (packed) -> their int8 codes for input scale s, as packed bf16 values
__device__ __forceinline__ uint32_t qpack(uint32_t y2, float s) { return pack_bf16(qv(bf16_lo(y2), s), qv(bf16_hi(y2), s)); }

// Stage 1, a half (tiles j < NT / 2 of Stage1): real a -> A (residual of m0), its code for m0.cv1 -> Aq
template <int C>
__device__ __forceinline__ void store_a(unsigned char *A, unsigned char *Aq, int f, int lane, int mt, const float (&acc)[Geo<C>::N1][4],
                                        const float (&bz)[Geo<C>::N1][2], const float (&mz)[Geo<C>::N1][2], uint32_t mw, float s_next) {
  using G = Geo<C>;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    auto y2 = [&](int j) { return relu_pack_bf16(rq(acc[j][2 * h], mz[j][0], bz[j][0]), rq(acc[j][2 * h + 1], mz[j][1], bz[j][1])); };
    if (C == 4 || C == 8) {
      const int bit = C == 4 ? 2 * g + (t >> 1) + 16 * h : 16 * mt + 8 * h + g;
      const int o = C == 4 ? (f + 2 * g + 16 * h) * G::PXA + 4 * t : (f + bit) * G::PXA + 4 * t;
      const bool in = (mw >> bit) & 1u;
      const uint32_t ya = y2(0);
      st32(A + o, in ? ya : 0u);
      st32(Aq + o, in ? qpack(ya, s_next) : 0u);
    } else {
      const int bit = 16 * mt + 8 * h + g;
      const int o = (f + bit) * G::PXA + 8 * t;
      const bool in = (mw >> bit) & 1u;
      const uint32_t v0 = y2(0), v1 = y2(1);
      st64(A + o, in ? v0 : 0u, in ? v1 : 0u);
      st64(Aq + o, in ? qpack(v0, s_next) : 0u, in ? qpack(v1, s_next) : 0u);
    }
  }
}
// Stage 1, b half (tiles j >= NT / 2): only cv3 reads b -> its code for cv3's input scale
template <int C>
__device__ __forceinline__ void store_b(unsigned char *Bv, int f, int lane, int mt, const float (&acc)[Geo<C>::N1][4],
                                        const float (&bz)[Geo<C>::N1][2], const float (&mz)[Geo<C>::N1][2], float s_cv3) {
  using G = Geo<C>;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    auto y2 = [&](int j) { return relu_pack_bf16(rq(acc[j][2 * h], mz[j][0], bz[j][0]), rq(acc[j][2 * h + 1], mz[j][1], bz[j][1])); };
    if (C == 4) {
      st32(Bv + (f + 2 * g + 16 * h - kB0) * G::PXA + 4 * t, qpack(y2(1), s_cv3));
    } else if (C == 8) {
      st32(Bv + (f + 16 * mt + 8 * h + g - kB0) * G::PXA + 4 * t, qpack(y2(1), s_cv3));
    } else {
      st64(Bv + (f + 16 * mt + 8 * h + g - kB0) * G::PXA + 8 * t, qpack(y2(2), s_cv3), qpack(y2(3), s_cv3));
    }
  }
}
// 3x3 stages.  !RES (t1, t2: read by one conv only): the code for that conv -> Dq.  RES (u, v): y = bf16(relu(.) + r) with
// r the bf16 residual in Dr (a, resp. u): y -> Dr (the next residual), its code -> Dq.
template <int C, bool RES>
__device__ __forceinline__ void store_3(unsigned char *Dr, unsigned char *Dq, int f, int lane, int mt, const float (&acc)[Geo<C>::N3][4],
                                        const float (&bz)[Geo<C>::N3][2], const float (&mz)[Geo<C>::N3][2], uint32_t mw, float s_next) {
  using G = Geo<C>;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    auto y2 = [&](int j, uint32_t r) {
      const float lo = rq(acc[j][2 * h], mz[j][0], bz[j][0]), hi = rq(acc[j][2 * h + 1], mz[j][1], bz[j][1]);
      if (RES) return pack_bf16(__fadd_rn(relu(lo), bf16_lo(r)), __fadd_rn(relu(hi), bf16_hi(r)));
      return relu_pack_bf16(lo, hi);
    };
    if (C == 4 || C == 8) {
      const int bit = C == 4 ? 2 * g + (t >> 1) + 16 * h : 16 * mt + 8 * h + g;
      const int o = C == 4 ? (f + 2 * g + 16 * h) * G::PXA + 4 * t : (f + bit) * G::PXA + 4 * t;
      const bool in = (mw >> bit) & 1u;
      const uint32_t v = y2(0, RES ? ld32(Dr + o) : 0u);
      if (RES) st32(Dr + o, in ? v : 0u);
      st32(Dq + o, in ? qpack(v, s_next) : 0u);
    } else {
      const int bit = 16 * mt + 8 * h + g;
      const int o = (f + bit) * G::PXA + 8 * t;
      const bool in = (mw >> bit) & 1u;
      uint2 r = make_uint2(0u, 0u);
      if (RES) r = ld64(Dr + o);
      const uint32_t v0 = y2(0, r.x), v1 = y2(1, r.y);
      if (RES) st64(Dr + o, in ? v0 : 0u, in ? v1 : 0u);
      st64(Dq + o, in ? qpack(v0, s_next) : 0u, in ? qpack(v1, s_next) : 0u);
    }
  }
}
// Stage 6: y = bf16(relu(float(acc) m + b)) -> staged output tile (the block's bf16 output; no consumer code here)
template <int C>
__device__ __forceinline__ void store_y(unsigned char *Y, int f, int lane, int mt, const float (&acc)[Geo<C>::N6][4],
                                        const float (&bz)[Geo<C>::N6][2], const float (&mz)[Geo<C>::N6][2]) {
  using G = Geo<C>;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    auto y2 = [&](int j) { return relu_pack_bf16(rq(acc[j][2 * h], mz[j][0], bz[j][0]), rq(acc[j][2 * h + 1], mz[j][1], bz[j][1])); };
    if (C == 4) st64(Y + (f - kB0 + 2 * g + 16 * h) * G::PXX + 8 * t, y2(0), y2(1));
    else if (C == 8) st64(Y + (f - kB0 + 16 * mt + 8 * h + g) * G::PXX + 8 * t, y2(0), y2(1));
    else st128(Y + (f - kB0 + 16 * mt + 8 * h + g) * G::PXX + 16 * t, y2(0), y2(1), y2(2), y2(3));
  }
}
}  // namespace q8
#endif

// ---- host-side packing: B fragments of every stage (m16n8k16: lane (g, t) holds B[2t, 2t+1][g] and B[2t+8, 2t+9][g])
// ldsm: the stage loads A with ldmatrix (logical k = physical column), otherwise with LDS.64 (phys_col permutation)
template <class F>
void pack_stage(std::vector<uint32_t> &out, int ksteps, int ntiles, bool ldsm, F wfun) {
  for (int s = 0; s < ksteps; ++s)
    for (int j = 0; j < ntiles; ++j)
      for (int lane = 0; lane < 32; ++lane) {
        const int n = lane >> 2, t = lane & 3;
        auto w = [&](int kk) { return host_f2bf(wfun(s, ldsm ? kk : phys_col(kk), j, n)); };
        out.push_back((uint32_t)w(2 * t) | ((uint32_t)w(2 * t + 1) << 16));
        out.push_back((uint32_t)w(2 * t + 8) | ((uint32_t)w(2 * t + 9) << 16));
      }
}

// weights (PyTorch layout, BN folded): w[0]=cv1 [c_][c], w[1]=cv2 [c_][c], w[2..5] [c_][c_][3][3], w[6]=cv3 [c][c]
template <int C>
void pack_all(const float *const w[7], std::vector<uint32_t> &out) {
  using G = Geo<C>;
  pack_stage(out, G::K1, G::N1, Stage1<C>::LDSM, [&](int s, int p, int j, int n) { return Stage1<C>::weight(w[0], w[1], s, p, j, n); });
  for (int i = 2; i < 6; ++i)
    pack_stage(out, G::K3, G::N3, Stage3<C>::LDSM, [&](int s, int p, int j, int n) { return Stage3<C>::weight(w[i], s, p, j, n); });
  pack_stage(out, G::K6, G::N6, Stage6<C>::LDSM, [&](int s, int p, int j, int n) { return Stage6<C>::weight(w[6], s, p, j, n); });
}

}  // namespace c3kf
}  // namespace uyd
