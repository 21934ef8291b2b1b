// Fused stem, second generation: model.0 Conv(3,16,3,2) + model.1 Conv(16,32,3,2) (+ optionally the 1x1
// Conv(32,16) that consumes model.1, i.e. model.2.cv1 of the YAML graph) in one launch.
//
// ncu of the first generation (stem_fused_kernel): 137 M warp instructions per batch of 64, 6.6 % of them HMMA;
// the rest was the hi/lo bf16 split of the frame (two shared-memory planes, two mma passes), 24 LDS.32 per
// 16 layer-0 pixels, per-segment division by 33 and a bias/ReLU/pack epilogue of ~45 instructions.  A first
// rewrite with pixels as mma ROWS still spent 30 % of its instructions on IMAD (per-k-step addresses and the
// register moves that interleave two LDS.64 into an A fragment).  Here every GEMM is TRANSPOSED, channels are
// the mma rows (A = weights) and pixels the columns (B = activations):
//   * a B fragment (k = t, t+4 | 2t.., 2t+8..; n = g) of one pixel is ONE LDS.64 straight into the register pair
//     the HMMA reads: no moves;
//   * the frame patch is one fp32 plane (tf32 by truncation after +2^-12 ulp rounding, 2^-11 relative: eight
//     times finer than the bf16 rounding of the layer's own output), stored as even rows | odd rows per channel, so
//     that the address of layer-0 pixel q = 34 y + x is  8 q + (per-thread constant of the k-step): layer 0 is a
//     fully unrolled loop over 8-pixel groups of a FLAT pixel index with immediate offsets (column 33 of each row
//     is a garbage pixel nobody reads; mma columns are independent);
//   * uint8 frames: x / 255 as q = x*r, q += fma(-q, 255, x) * r, which equals the IEEE quotient for all 256 inputs
//     (tests/test_stem_emu.py), so a uint8 frame gives bit-identical results to its float image / 255;
//   * the patch lines are placed (odd rows at line 20, 38 lines per channel, 272-byte lines) so that the two
//     (channel, row) combinations one LDS.64 of layer 0 touches fall into disjoint shared-memory banks;
//   * the layer-0 tile keeps channels (w, w + 8) in word w of a 48-byte pixel: the transposed accumulators
//     (channel g | g+8, pixels 2t | 2t+1) pack into it with two cvt.rn.relu.bf16x2, conflict-free; the top row /
//     left column outside the image (layer 1's zero padding) is cleared afterwards, by border CTAs only;
//   * layer 1 (mma.m16n8k16 bf16, two 16-channel m-tiles, A fragments = one LDS.128 each) reads those words as B
//     fragments; its output rows are assigned to channels so that a thread owns 4 contiguous channels;
//   * optional 1x1 (32 -> 16): the bf16-rounded layer-1 accumulators become B fragments through movmatrix
//     (8x8 transpose in registers), so the 32-channel 160x160 tensor is never written.
//
// The per-lane maps are shared with the host emulation tests/stem_emu.cpp (test infrastructure).
#pragma once
#include "c3k_flat.cuh"  // bf16 helpers, ld64/st32

namespace uyd {
namespace stemv2 {
using c3kf::host_f2bf;
using c3kf::ld64;
using c3kf::relu_pack_bf16;
using c3kf::st32;

constexpr int kTH = 8, kTW = 16;                       // layer-1 output tile
constexpr int kL0H = 2 * kTH + 1, kL0P = 2 * kTW + 2;  // 17 rows x 34 pixels (pixel 33 of a row is garbage)
constexpr int kGroups = (16 * kL0P + 2 * kTW + 1 + 7) / 8;  // 8-pixel groups covering q = 0 .. 16*34 + 32: 73
constexpr int kL0Px = kGroups * 8;                     // 584
constexpr int kInH = 2 * kL0H + 1, kInW = 2 * kL0P;    // 35 x 68 frame patch (column 0 only feeds the zero slot)
constexpr int kOddLine = 20, kChanLines = 38;          // 18 even patch rows at line 0, 17 odd ones at line 20 of a channel
constexpr int kL0Pitch = 48;                           // bytes per layer-0 pixel (8 words + pad: conflict-free)
constexpr int kPatchBytes = 3 * kChanLines * kInW * 4;  // 31008
constexpr int kL0Bytes = kL0Px * kL0Pitch;             // 28032
constexpr int kW0Words = 5 * 128, kW1Words = 9 * 2 * 128, kW2Words = 2 * 128;  // A fragments: 32 lanes x 4 words each
constexpr int kSmemBytes = kPatchBytes + kL0Bytes + kW1Words * 4;
constexpr int kThreads = 256;
// fragment buffer: [L0 | L1 | 1x1]
constexpr int kFragWords = kW0Words + kW1Words + kW2Words;

// patch line of frame-patch row r of channel c: even rows first
C3K_HD int patch_line(int c, int r) { return c * kChanLines + ((r & 1) ? kOddLine + (r >> 1) : (r >> 1)); }

// ---- layer 0 (tf32 m16n8k8): logical k = t -> slot 2(t&1), k = t+4 -> slot 2(t&1)+1 of combination 2s + (t>>1) ----
// float offset of this lane's slot pair in k-step s, relative to the pixel term 2 q
C3K_HD int l0_k_off(int s, int t) {
  int c = 2 * s + (t >> 1);
  if (c > 8) c = 8;  // combination 9 has zero weights; it re-reads combination 8
  const int ky = c % 3;
  return (patch_line(c / 3, ky & 1) + (ky >> 1)) * kInW + 2 * (t & 1);   // ky = 2: even row of the NEXT pixel row
}
// w0 [16][3][3][3]; k = logical column 0..7 of k-step s
inline float l0_weight(const float *w0, int s, int k, int n) {
  const int t = k & 3, combo = 2 * s + (t >> 1), slot = 2 * (t & 1) + (k >> 2), kx = slot - 1;
  if (combo > 8 || kx < 0) return 0.f;
  return w0[((n * 3 + combo / 3) * 3 + combo % 3) * 3 + kx];
}
// accumulators of lane (g, t) for group grp: (channel g | g+8) x (pixel 8 grp + 2t | + 1) -> word g of the pixel
C3K_HD void l0_store(unsigned char *l0s, int q, int g, float lo_ch, float hi_ch) { st32(l0s + q * kL0Pitch + 4 * g, relu_pack_bf16(lo_ch, hi_ch)); }

// ---- layer 1 (bf16 m16n8k16): rows = output channels (2 m-tiles), columns = 8 output pixels of one row ----
// logical k of the B fragment -> layer-0 channel (word 2t = channels (2t, 2t+8), word 2t+1 = (2t+1, 2t+9))
C3K_HD int l1_k_chan(int kk) { return 2 * ((kk & 7) >> 1) + (kk >> 3) + 8 * (kk & 1); }
// byte offset of lane (g, t)'s B fragment: output pixel (row, 8 xg + g), tap (ky, kx)
C3K_HD int l1_b_off(int row, int xg, int g, int t, int tap) {
  return ((2 * row + tap / 3) * kL0P + 2 * (8 * xg + g) + tap % 3) * kL0Pitch + 8 * t;
}
// output channel of row r of m-tile mt: lane g owns channels 4g .. 4g+3
C3K_HD int l1_chan(int mt, int r) { return 4 * (r & 7) + 2 * mt + (r >> 3); }

// ---- optional 1x1 (32 -> 16): k-step ks = m-tile ks of layer 1, logical k = its row; output row r -> channel ----
C3K_HD int pw_chan(int r) { return 2 * (r & 7) + (r >> 3); }

// ---- host-side packing ----
inline uint32_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float bits_f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline float bf16_round(float f) { return bits_f32((uint32_t)host_f2bf(f) << 16); }

// x / 255 for integer x in [0, 255]: equals the IEEE fp32 quotient (checked exhaustively, tests/test_stem_emu.py)
C3K_HD float div255(float x) {
#if defined(__CUDA_ARCH__)
  const float r = 1.0f / 255.0f;
  const float q = __fmul_rn(x, r);
  return __fmaf_rn(__fmaf_rn(-q, 255.0f, x), r, q);
#else
  return x / 255.0f;
#endif
}

// m16n8k8 A fragment: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4); m16n8k16: a0 (g, 2t..) a1 (g+8, 2t..) a2 (g, 2t+8..) a3 (g+8, 2t+8..)
inline void pack(const float *w0, const float *w1, const float *w2 /* may be null */, std::vector<uint32_t> &out) {
  for (int s = 0; s < 5; ++s)
    for (int lane = 0; lane < 32; ++lane) {
      const int g = lane >> 2, t = lane & 3;
      auto w = [&](int row, int k) { return f32_bits(bf16_round(l0_weight(w0, s, k, row))); };
      out.push_back(w(g, t)); out.push_back(w(g + 8, t)); out.push_back(w(g, t + 4)); out.push_back(w(g + 8, t + 4));
    }
  for (int tap = 0; tap < 9; ++tap)
    for (int mt = 0; mt < 2; ++mt)
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3;
        auto w = [&](int row, int kk) {
          auto one = [&](int k1) { return (uint32_t)host_f2bf(w1[(l1_chan(mt, row) * 16 + l1_k_chan(k1)) * 9 + tap]); };
          return one(kk) | (one(kk + 1) << 16);
        };
        out.push_back(w(g, 2 * t)); out.push_back(w(g + 8, 2 * t)); out.push_back(w(g, 2 * t + 8)); out.push_back(w(g + 8, 2 * t + 8));
      }
  for (int ks = 0; ks < 2; ++ks)
    for (int lane = 0; lane < 32; ++lane) {
      const int g = lane >> 2, t = lane & 3;
      auto w = [&](int row, int kk) {
        auto one = [&](int k1) { return w2 ? (uint32_t)host_f2bf(w2[pw_chan(row) * 32 + l1_chan(ks, k1)]) : 0u; };
        return one(kk) | (one(kk + 1) << 16);
      };
      out.push_back(w(g, 2 * t)); out.push_back(w(g + 8, 2 * t)); out.push_back(w(g, 2 * t + 8)); out.push_back(w(g + 8, 2 * t + 8));
    }
}

}  // namespace stemv2
}  // namespace uyd
