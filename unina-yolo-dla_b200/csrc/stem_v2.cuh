// Fused stem, second generation: model.0 Conv(3,16,3,2) + model.1 Conv(16,32,3,2) (+ optionally the 1x1
// Conv(32,16) that consumes model.1, i.e. model.2.cv1 of the YAML graph) in one launch.
//
// ncu of the first generation (stem_fused_kernel): 137 M warp instructions per batch of 64, 6.6 % of them HMMA;
// the rest was the hi/lo bf16 split of the frame (two shared-memory planes, two mma passes), 24 LDS.32 per
// 16 layer-0 pixels, per-segment division by 33 and a bias/ReLU/pack epilogue of ~45 instructions.  Here:
//   * the frame patch is ONE fp32 plane rounded to tf32 (cvt.rna, 2^-11 relative: eight times finer than the
//     bf16 rounding of the layer's own output) and layer 0 runs on mma.m16n8k8.tf32: K = 5 k-steps of
//     (2 (channel, row) combinations x 4 column slots), each fragment half is one LDS.64;
//   * uint8 frames: x / 255 as q = x*r, q += fma(-q, 255, x) * r, which equals the IEEE quotient for all 256 inputs;
//   * layer-0 pixel coordinates advance incrementally, biases start the accumulators, ReLU + bf16 pack is one
//     cvt.rn.relu.bf16x2;
//   * layer 1 (mma.m16n8k16 bf16) reads layer-0 pixels with LDS.64 (k permutation as in c3k_flat.cuh), and its output
//     channels are permuted so that a thread owns 8 contiguous channels: 16-byte global stores from registers;
//   * optional 1x1 (32 -> 16): the layer-1 accumulators are re-packed in registers as A fragments (no shared
//     memory round trip), so the 32-channel 160x160 tensor is never written.
//
// The per-lane maps are shared with the host emulation tests/stem_emu.cpp (test infrastructure).
#pragma once
#include "c3k_flat.cuh"  // bf16 helpers, ld64/st32, phys_col

namespace uyd {
namespace stemv2 {
using c3kf::host_f2bf;
using c3kf::ld64;
using c3kf::phys_col;
using c3kf::relu_pack_bf16;
using c3kf::st32;

constexpr int kTH = 8, kTW = 16;                            // layer-1 output tile
constexpr int kL0H = 2 * kTH + 1, kL0W = 2 * kTW + 1;       // 17 x 33 layer-0 region
constexpr int kL0Px = kL0H * kL0W, kL0Segs = (kL0Px + 15) / 16;  // 561 pixels, 36 segments
constexpr int kInH = 2 * kL0H + 1, kInW = 2 * kL0W + 2;     // 35 x 68 frame patch (column 0 only feeds the zero slot)
constexpr int kL0Pitch = 48;                                // bytes per layer-0 pixel (16 bf16 + pad: conflict-free LDS.64 at stride 2)
constexpr int kPatchBytes = 3 * kInH * kInW * 4;
constexpr int kL0Bytes = kL0Segs * 16 * kL0Pitch;
constexpr int kW0Words = 5 * 2 * 64, kW1Words = 9 * 4 * 64, kW2Words = 2 * 2 * 64;  // fragment words: L0 (tf32), L1, 1x1
constexpr int kSmemBytes = kPatchBytes + kL0Bytes + kW1Words * 4;
constexpr int kThreads = 256;

// ---- layer 0 (tf32 m16n8k8): logical k = t -> slot 2(t&1), k = t+4 -> slot 2(t&1)+1 of combination 2s + (t>>1) ----
// float offset into the patch of this lane's two columns in k-step s, relative to the pixel term 2y*68 + 2x
C3K_HD int l0_k_off(int s, int t) {
  int c = 2 * s + (t >> 1);
  if (c > 8) c = 8;  // combination 9 has zero weights; it re-reads combination 8
  return ((c / 3) * kInH + c % 3) * kInW + 2 * (t & 1);
}
C3K_HD void l0_load_a(const unsigned char *patch, int r0, int r1, int koff, uint32_t (&a)[4]) {  // r = 2y*68 + 2x of rows g, g+8
  const uint2 lo = ld64(patch + (r0 + koff) * 4), hi = ld64(patch + (r1 + koff) * 4);
  a[0] = lo.x; a[2] = lo.y; a[1] = hi.x; a[3] = hi.y;
}
// pixel p = seg*16 + g + 8h of the region: channels 8j + 2t, +1 of n-tile j
C3K_HD void l0_store(unsigned char *l0s, int p, int t, int j, float v0, float v1, bool inside) {
  st32(l0s + p * kL0Pitch + (8 * j + 2 * t) * 2, inside ? relu_pack_bf16(v0, v1) : 0u);
}
// w0 [16][3][3][3]; k = logical column 0..7 of k-step s
inline float l0_weight(const float *w0, int s, int k, int n) {
  const int t = k & 3, combo = 2 * s + (t >> 1), slot = 2 * (t & 1) + (k >> 2), kx = slot - 1;
  if (combo > 8 || kx < 0) return 0.f;
  return w0[((n * 3 + combo / 3) * 3 + combo % 3) * 3 + kx];
}

// ---- layer 1 (bf16 m16n8k16): warp = output row, rows g / g+8 = output columns g / g+8, k-step = tap ----
C3K_HD void l1_load_a(const unsigned char *l0s, int warp, int lane, int tap, uint32_t (&a)[4]) {
  const int g = lane >> 2, t = lane & 3;
  const unsigned char *p = l0s + ((2 * warp + tap / 3) * kL0W + 2 * g + tap % 3) * kL0Pitch + 8 * t;
  const uint2 lo = ld64(p), hi = ld64(p + 16 * kL0Pitch);
  a[0] = lo.x; a[2] = lo.y; a[1] = hi.x; a[3] = hi.y;
}
// output channel of accumulator (n-tile j, column 2t + e): 8t + 2j + e  (8 contiguous channels per thread)
C3K_HD int l1_chan(int j, int t, int e) { return 8 * t + 2 * j + e; }
// w1 [32][16][3][3]; p = physical column (input channel), n = fragment column (lane >> 2)
inline float l1_weight(const float *w1, int tap, int p, int j, int n) {
  return w1[(l1_chan(j, n >> 1, n & 1) * 16 + p) * 9 + tap];
}

// ---- optional 1x1 (32 -> 16): A fragments are the re-packed layer-1 accumulators ----
// k-step ks: a0/a1 = n-tile 2ks (logical k 2t+e), a2/a3 = n-tile 2ks+1 (logical k 2t+8+e)
C3K_HD int pw_chan(int j, int t, int e) { return 4 * t + 2 * j + e; }  // output channel of acc2 (n-tile j, column 2t+e)
inline float pw_weight(const float *w2, int ks, int kk, int j, int n) {  // w2 [16][32]; kk = logical column 0..15
  const int chan = l1_chan(2 * ks + (kk >> 3), (kk & 7) >> 1, kk & 1);
  return w2[pw_chan(j, n >> 1, n & 1) * 32 + chan];
}

// ---- host-side packing: [L0 tf32 | L1 bf16 | 1x1 bf16] fragments, biases [16 | 32 | 16] ----
inline uint32_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float bf16_round(float f) { const uint32_t u = (uint32_t)host_f2bf(f) << 16; float r; memcpy(&r, &u, 4); return r; }

inline void pack(const float *w0, const float *w1, const float *w2 /* may be null */, std::vector<uint32_t> &out) {
  for (int s = 0; s < 5; ++s)
    for (int j = 0; j < 2; ++j)
      for (int lane = 0; lane < 32; ++lane) {
        const int n = 8 * j + (lane >> 2), t = lane & 3;
        out.push_back(f32_bits(bf16_round(l0_weight(w0, s, t, n))));      // b0: k = t
        out.push_back(f32_bits(bf16_round(l0_weight(w0, s, t + 4, n))));  // b1: k = t + 4
      }
  for (int tap = 0; tap < 9; ++tap)
    for (int j = 0; j < 4; ++j)
      for (int lane = 0; lane < 32; ++lane) {
        const int n = lane >> 2, t = lane & 3;
        auto w = [&](int kk) { return (uint32_t)host_f2bf(l1_weight(w1, tap, phys_col(kk), j, n)); };
        out.push_back(w(2 * t) | (w(2 * t + 1) << 16));
        out.push_back(w(2 * t + 8) | (w(2 * t + 9) << 16));
      }
  for (int ks = 0; ks < 2; ++ks)
    for (int j = 0; j < 2; ++j)
      for (int lane = 0; lane < 32; ++lane) {
        const int n = lane >> 2, t = lane & 3;
        auto w = [&](int kk) { return w2 ? (uint32_t)host_f2bf(pw_weight(w2, ks, kk, j, n)) : 0u; };
        out.push_back(w(2 * t) | (w(2 * t + 1) << 16));
        out.push_back(w(2 * t + 8) | (w(2 * t + 9) << 16));
      }
}

// x / 255 for integer x in [0, 255]: equals the IEEE fp32 quotient (checked exhaustively, tests/test_stem_emu.py)
C3K_HD float div255(float x) {
  const float r = 1.0f / 255.0f;
#if defined(__CUDA_ARCH__)
  const float q = __fmul_rn(x, r);
  return __fmaf_rn(__fmaf_rn(-q, 255.0f, x), r, q);
#else
  (void)r;
  return x / 255.0f;
#endif
}

}  // namespace stemv2
}  // namespace uyd
