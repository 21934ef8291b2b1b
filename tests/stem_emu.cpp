// Host emulation of one CTA of stem_v2_kernel (csrc/stem_fused.cu): the per-lane address maps, fragment
// permutations and weight packing are the kernel's own (csrc/stem_v2.cuh); the mma.sync fragment layouts
// (m16n8k8 tf32, m16n8k16 bf16), the segment loop and the patch load map are restated here.  Test infrastructure
// only (tests/test_stem_emu.py).  Shared memory is poisoned with NaN before every tile.
#include <cmath>
#include <cstdio>

#include "../unina-yolo-dla_b200/csrc/stem_v2.cuh"

using namespace uyd::stemv2;

static float bf2f(uint32_t b) { uint32_t u = (b & 0xffffu) << 16; float f; memcpy(&f, &u, 4); return f; }
static float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
// cvt.rna.tf32.f32: round to nearest, ties away from zero, 10 mantissa bits kept
static uint32_t to_tf32(float v) {
  uint32_t u = f32_bits(v);
  if ((u & 0x7f800000u) == 0x7f800000u) return u;
  return (u + 0x1000u) & 0xffffe000u;
}

// m16n8k8 tf32: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4); b0 (k = t, n = g) b1 (k = t+4, n = g)
static void warp_mma_tf32(float acc[32][4], const uint32_t a[32][4], const uint32_t *bfrag /* [32][2] */) {
  float A[16][8], B[8][8];
  for (int lane = 0; lane < 32; ++lane) {
    const int g = lane >> 2, t = lane & 3;
    A[g][t] = bits2f(a[lane][0]); A[g + 8][t] = bits2f(a[lane][1]);
    A[g][t + 4] = bits2f(a[lane][2]); A[g + 8][t + 4] = bits2f(a[lane][3]);
    B[t][g] = bits2f(bfrag[lane * 2]); B[t + 4][g] = bits2f(bfrag[lane * 2 + 1]);
  }
  for (int lane = 0; lane < 32; ++lane) {
    const int g = lane >> 2, t = lane & 3;
    for (int i = 0; i < 4; ++i) {
      const int m = g + 8 * (i >> 1), n = 2 * t + (i & 1);
      float s = acc[lane][i];
      for (int k = 0; k < 8; ++k) s += A[m][k] * B[k][n];
      acc[lane][i] = s;
    }
  }
}

static void warp_mma_bf16(float acc[32][4], const uint32_t a[32][4], const uint32_t *bfrag /* [32][2] */) {
  float A[16][16], B[16][8];
  for (int lane = 0; lane < 32; ++lane) {
    const int g = lane >> 2, t = lane & 3;
    const uint32_t *r = a[lane];
    A[g][2 * t] = bf2f(r[0]); A[g][2 * t + 1] = bf2f(r[0] >> 16);
    A[g + 8][2 * t] = bf2f(r[1]); A[g + 8][2 * t + 1] = bf2f(r[1] >> 16);
    A[g][2 * t + 8] = bf2f(r[2]); A[g][2 * t + 9] = bf2f(r[2] >> 16);
    A[g + 8][2 * t + 8] = bf2f(r[3]); A[g + 8][2 * t + 9] = bf2f(r[3] >> 16);
    const uint32_t b0 = bfrag[lane * 2], b1 = bfrag[lane * 2 + 1];
    B[2 * t][g] = bf2f(b0); B[2 * t + 1][g] = bf2f(b0 >> 16);
    B[2 * t + 8][g] = bf2f(b1); B[2 * t + 9][g] = bf2f(b1 >> 16);
  }
  for (int lane = 0; lane < 32; ++lane) {
    const int g = lane >> 2, t = lane & 3;
    for (int i = 0; i < 4; ++i) {
      const int m = g + 8 * (i >> 1), n = 2 * t + (i & 1);
      float s = acc[lane][i];
      for (int k = 0; k < 16; ++k) s += A[m][k] * B[k][n];
      acc[lane][i] = s;
    }
  }
}

// frame: [3][ih][iw] fp32 (already divided by 255 for 8-bit frames); out: [oh][ow][pw ? 16 : 32] fp32 (bf16 values)
extern "C" int stem_emu(int ih, int iw, int pw, const float *w0, const float *b0, const float *w1, const float *b1, const float *w2,
                        const float *b2, const float *frame, float *out) {
  if (ih % 4 || iw % 4) return 1;
  const int oh = ih / 4, ow = iw / 4, LH = ih / 2, LW = iw / 2, oc = pw ? 16 : 32;
  std::vector<uint32_t> frags;
  pack(w0, w1, pw ? w2 : nullptr, frags);
  if ((int)frags.size() != kW0Words + kW1Words + kW2Words) return 2;
  float bias[64] = {0};
  for (int i = 0; i < 16; ++i) bias[i] = b0[i];
  for (int i = 0; i < 32; ++i) bias[16 + i] = b1[i];
  for (int i = 0; i < 16 && pw; ++i) bias[48 + i] = b2[i];
  std::vector<unsigned char> smem(kSmemBytes);
  unsigned char *patch = smem.data(), *l0s = smem.data() + kPatchBytes;
  const uint32_t *w1s = frags.data() + kW0Words;

  for (int oy0 = 0; oy0 < oh; oy0 += kTH)
    for (int ox0 = 0; ox0 < ow; ox0 += kTW) {
      for (size_t i = 0; i < smem.size() / 4; ++i) reinterpret_cast<uint32_t *>(smem.data())[i] = 0x7fc00000u;
      for (int i = 0; i < kL0Segs * 16 * kL0Pitch / 2; ++i) reinterpret_cast<uint16_t *>(l0s)[i] = 0x7fc0;
      const int ix0 = 4 * ox0 - 4, iy0 = 4 * oy0 - 3;
      // patch load map of the kernel: thread -> (column vector pj, row lane prl), 7 lines prl + 15 it
      for (int tid = 0; tid < kThreads; ++tid) {
        const int pj = tid % (kInW / 4), prl = tid / (kInW / 4);
        if (prl >= 15) continue;
        const int ix = ix0 + 4 * pj;
        const bool col_ok = ix >= 0 && ix + 3 < iw;
        int c = 0, r = prl;
        for (int it = 0; it < 7; ++it) {
          const int iy = iy0 + r;
          for (int e = 0; e < 4; ++e) {
            const float v = (col_ok && iy >= 0 && iy < ih) ? frame[((size_t)c * ih + iy) * iw + ix + e] : 0.f;
            st32(patch + ((prl + 15 * it) * kInW + 4 * pj + e) * 4, to_tf32(v));
          }
          r += 15;
          if (r >= kInH) { r -= kInH; ++c; }
        }
      }
      // layer 0
      const int ly0 = 2 * oy0 - 1, lx0 = 2 * ox0 - 1;
      for (int seg = 0; seg < kL0Segs; ++seg) {
        float acc[2][32][4];
        uint32_t af[32][4];
        int ys[32][2], xs[32][2];
        for (int lane = 0; lane < 32; ++lane) {
          const int g = lane >> 2, t = lane & 3;
          for (int h = 0; h < 2; ++h) {
            const int p = seg * 16 + g + 8 * h;
            ys[lane][h] = p / kL0W; xs[lane][h] = p % kL0W;
          }
          for (int j = 0; j < 2; ++j) {
            acc[j][lane][0] = acc[j][lane][2] = bias[8 * j + 2 * t];
            acc[j][lane][1] = acc[j][lane][3] = bias[8 * j + 2 * t + 1];
          }
        }
        for (int s = 0; s < 5; ++s) {
          for (int lane = 0; lane < 32; ++lane) {
            const int t = lane & 3;
            const int r0 = 2 * std::min(ys[lane][0], kL0H - 1) * kInW + 2 * xs[lane][0];
            const int r1 = 2 * std::min(ys[lane][1], kL0H - 1) * kInW + 2 * xs[lane][1];
            l0_load_a(patch, r0, r1, l0_k_off(s, t), af[lane]);
          }
          for (int j = 0; j < 2; ++j) warp_mma_tf32(acc[j], af, frags.data() + (s * 2 + j) * 64);
        }
        for (int lane = 0; lane < 32; ++lane) {
          const int g = lane >> 2, t = lane & 3;
          for (int h = 0; h < 2; ++h) {
            const int p = seg * 16 + g + 8 * h;
            if (p >= kL0Px) continue;
            const int ly = ly0 + ys[lane][h], lx = lx0 + xs[lane][h];
            const bool in = ly >= 0 && ly < LH && lx >= 0 && lx < LW;
            for (int j = 0; j < 2; ++j) l0_store(l0s, p, t, j, acc[j][lane][2 * h], acc[j][lane][2 * h + 1], in);
          }
        }
      }
      // layer 1 (+ 1x1)
      for (int warp = 0; warp < kTH; ++warp) {
        float acc[4][32][4];
        uint32_t af[32][4];
        for (int lane = 0; lane < 32; ++lane) {
          const int t = lane & 3;
          for (int j = 0; j < 4; ++j) {
            acc[j][lane][0] = acc[j][lane][2] = bias[16 + l1_chan(j, t, 0)];
            acc[j][lane][1] = acc[j][lane][3] = bias[16 + l1_chan(j, t, 1)];
          }
        }
        for (int tap = 0; tap < 9; ++tap) {
          for (int lane = 0; lane < 32; ++lane) l1_load_a(l0s, warp, lane, tap, af[lane]);
          for (int j = 0; j < 4; ++j) warp_mma_bf16(acc[j], af, w1s + (tap * 4 + j) * 64);
        }
        const int oy = oy0 + warp;
        if (!pw) {
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3;
            for (int h = 0; h < 2; ++h) {
              const int px = ox0 + g + 8 * h;
              if (oy >= oh || px >= ow) continue;
              for (int j = 0; j < 4; ++j)
                for (int e = 0; e < 2; ++e) {
                  const uint32_t v = relu_pack_bf16(acc[j][lane][2 * h + e], 0.f);
                  out[((size_t)oy * ow + px) * oc + 8 * t + 2 * j + e] = bf2f(v);
                }
            }
          }
        } else {
          const uint32_t *w2s = frags.data() + kW0Words + kW1Words;
          float acc2[2][32][4];
          for (int lane = 0; lane < 32; ++lane) {
            const int t = lane & 3;
            for (int j = 0; j < 2; ++j) {
              acc2[j][lane][0] = acc2[j][lane][2] = bias[48 + pw_chan(j, t, 0)];
              acc2[j][lane][1] = acc2[j][lane][3] = bias[48 + pw_chan(j, t, 1)];
            }
          }
          for (int ks = 0; ks < 2; ++ks) {
            for (int lane = 0; lane < 32; ++lane) {
              af[lane][0] = relu_pack_bf16(acc[2 * ks][lane][0], acc[2 * ks][lane][1]);
              af[lane][1] = relu_pack_bf16(acc[2 * ks][lane][2], acc[2 * ks][lane][3]);
              af[lane][2] = relu_pack_bf16(acc[2 * ks + 1][lane][0], acc[2 * ks + 1][lane][1]);
              af[lane][3] = relu_pack_bf16(acc[2 * ks + 1][lane][2], acc[2 * ks + 1][lane][3]);
            }
            for (int j = 0; j < 2; ++j) warp_mma_bf16(acc2[j], af, w2s + (ks * 2 + j) * 64);
          }
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3;
            for (int h = 0; h < 2; ++h) {
              const int px = ox0 + g + 8 * h;
              if (oy >= oh || px >= ow) continue;
              for (int j = 0; j < 2; ++j)
                for (int e = 0; e < 2; ++e) {
                  const uint32_t v = relu_pack_bf16(acc2[j][lane][2 * h + e], 0.f);
                  out[((size_t)oy * ow + px) * oc + 4 * t + 2 * j + e] = bf2f(v);
                }
            }
          }
        }
      }
    }
  return 0;
}

// x / 255 by the kernel's two-step correction (device branch restated with fmaf) for all 256 inputs
extern "C" int div255_mismatches() {
  int bad = 0;
  const float r = 1.0f / 255.0f;
  for (int x = 0; x < 256; ++x) {
    const float xf = (float)x, q = xf * r;
    const float got = fmaf(fmaf(-q, 255.0f, xf), r, q);
    bad += got != xf / 255.0f;
  }
  return bad;
}
