// Host emulation of one CTA of stem_v2_kernel (csrc/stem_fused.cu): the per-lane address maps, fragment
// permutations and weight packing are the kernel's own (csrc/stem_v2.cuh); the mma.sync / movmatrix fragment
// layouts (m16n8k8 tf32, m16n8k16 bf16, m8n8 transpose) and the group / tap loops are restated here.  Test infrastructure
// only (tests/test_stem_emu.py).  Shared memory is poisoned with NaN before every tile.
#include <cmath>
#include <cstdio>

#include "../unina-yolo-dla_b200/csrc/stem_v2.cuh"

using namespace uyd::stemv2;

static float bf2f(uint32_t b) { uint32_t u = (b & 0xffffu) << 16; float f; memcpy(&f, &u, 4); return f; }
static float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// m16n8k8 tf32: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4); b0 (k = t, n = g) b1 (k = t+4, n = g)
static void warp_mma_tf32(float acc[32][4], const uint32_t *afrag /* [32][4] */, const uint32_t b[32][2]) {
  float A[16][8], B[8][8];
  for (int lane = 0; lane < 32; ++lane) {
    const int g = lane >> 2, t = lane & 3;
    const uint32_t *a = afrag + lane * 4;
    A[g][t] = bits2f(a[0] & 0xffffe000u); A[g + 8][t] = bits2f(a[1] & 0xffffe000u);
    A[g][t + 4] = bits2f(a[2] & 0xffffe000u); A[g + 8][t + 4] = bits2f(a[3] & 0xffffe000u);
    B[t][g] = bits2f(b[lane][0] & 0xffffe000u); B[t + 4][g] = bits2f(b[lane][1] & 0xffffe000u);  // the mma reads 19 bits
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

static void warp_mma_bf16(float acc[32][4], const uint32_t *afrag /* [32][4] */, const uint32_t b[32][2]) {
  float A[16][16], B[16][8];
  for (int lane = 0; lane < 32; ++lane) {
    const int g = lane >> 2, t = lane & 3;
    const uint32_t *r = afrag + lane * 4;
    A[g][2 * t] = bf2f(r[0]); A[g][2 * t + 1] = bf2f(r[0] >> 16);
    A[g + 8][2 * t] = bf2f(r[1]); A[g + 8][2 * t + 1] = bf2f(r[1] >> 16);
    A[g][2 * t + 8] = bf2f(r[2]); A[g][2 * t + 9] = bf2f(r[2] >> 16);
    A[g + 8][2 * t + 8] = bf2f(r[3]); A[g + 8][2 * t + 9] = bf2f(r[3] >> 16);
    B[2 * t][g] = bf2f(b[lane][0]); B[2 * t + 1][g] = bf2f(b[lane][0] >> 16);
    B[2 * t + 8][g] = bf2f(b[lane][1]); B[2 * t + 9][g] = bf2f(b[lane][1] >> 16);
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

// movmatrix.m8n8.trans.b16: lane (g, t) holds M[g][2t], M[g][2t+1]; afterwards M[2t][g], M[2t+1][g]
static void warp_movmatrix_trans(const uint32_t in[32], uint32_t out[32]) {
  uint16_t M[8][8];
  for (int lane = 0; lane < 32; ++lane) {
    const int g = lane >> 2, t = lane & 3;
    M[g][2 * t] = (uint16_t)in[lane]; M[g][2 * t + 1] = (uint16_t)(in[lane] >> 16);
  }
  for (int lane = 0; lane < 32; ++lane) {
    const int g = lane >> 2, t = lane & 3;
    out[lane] = (uint32_t)M[2 * t][g] | ((uint32_t)M[2 * t + 1][g] << 16);
  }
}

// frame: [3][ih][iw] fp32 (u8 = 1: integer values 0..255, divided by 255 here as the kernel's uint8 path does);
// out: [oh][ow][pw ? 16 : 32] fp32 (bf16 values)
extern "C" int stem_emu(int ih, int iw, int pw, int u8, const float *w0, const float *b0, const float *w1, const float *b1,
                        const float *w2, const float *b2, const float *frame, float *out) {
  if (ih % 4 || iw % 4) return 1;
  const int oh = ih / 4, ow = iw / 4, oc = pw ? 16 : 32;
  std::vector<uint32_t> frags;
  pack(w0, w1, pw ? w2 : nullptr, frags);
  if ((int)frags.size() != kFragWords) return 2;
  float bias[64] = {0};
  for (int i = 0; i < 16; ++i) bias[i] = b0[i];
  for (int i = 0; i < 32; ++i) bias[16 + i] = b1[i];
  for (int i = 0; i < 16 && pw; ++i) bias[48 + i] = b2[i];
  std::vector<unsigned char> smem(kSmemBytes);
  unsigned char *patch = smem.data(), *l0s = smem.data() + kPatchBytes;
  const uint32_t *w0f = frags.data(), *w1f = frags.data() + kW0Words, *w2f = w1f + kW1Words;

  for (int oy0 = 0; oy0 < oh; oy0 += kTH)
    for (int ox0 = 0; ox0 < ow; ox0 += kTW) {
      for (size_t i = 0; i < smem.size() / 4; ++i) reinterpret_cast<uint32_t *>(smem.data())[i] = 0x7fc07fc0u;  // NaN as fp32 and bf16
      const int ix0 = 4 * ox0 - 4, iy0 = 4 * oy0 - 3;
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < kInH; ++r)
          for (int col = 0; col < kInW; ++col) {
            const int iy = iy0 + r, ix = ix0 + col;
            const bool in = iy >= 0 && iy < ih && ix >= 0 && ix < iw;
            const float v = in ? frame[((size_t)c * ih + iy) * iw + ix] : 0.f;
            st32(patch + (patch_line(c, r) * kInW + col) * 4, !in ? 0u : f32_bits(u8 ? div255(v) : v) + 0x1000u);
          }
      // layer 0
      for (int grp = 0; grp < kGroups; ++grp) {
        float acc[32][4];
        uint32_t bfr[32][2];
        for (int lane = 0; lane < 32; ++lane) {
          const int g = lane >> 2;
          acc[lane][0] = acc[lane][1] = bias[g];
          acc[lane][2] = acc[lane][3] = bias[g + 8];
        }
        for (int s = 0; s < 5; ++s) {
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3;
            const uint2 v = ld64(patch + 8 * (8 * grp + g) + 4 * l0_k_off(s, t));
            bfr[lane][0] = v.x; bfr[lane][1] = v.y;
          }
          warp_mma_tf32(acc, w0f + s * 128, bfr);
        }
        for (int lane = 0; lane < 32; ++lane) {
          const int g = lane >> 2, t = lane & 3;
          l0_store(l0s, 8 * grp + 2 * t, g, acc[lane][0], acc[lane][2]);
          l0_store(l0s, 8 * grp + 2 * t + 1, g, acc[lane][1], acc[lane][3]);
        }
      }
      if (oy0 == 0)
        for (int i = 0; i < kL0P * 8; ++i) st32(l0s + (i >> 3) * kL0Pitch + 4 * (i & 7), 0u);
      if (ox0 == 0)
        for (int i = 0; i < kL0H * 8; ++i) st32(l0s + (i >> 3) * kL0P * kL0Pitch + 4 * (i & 7), 0u);
      // layer 1 (+ 1x1)
      for (int warp = 0; warp < kTH; ++warp)
        for (int xg = 0; xg < 2; ++xg) {
          float acc[2][32][4];
          uint32_t bfr[32][2];
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2;
            for (int mt = 0; mt < 2; ++mt) {
              acc[mt][lane][0] = acc[mt][lane][1] = bias[16 + 4 * g + 2 * mt];
              acc[mt][lane][2] = acc[mt][lane][3] = bias[16 + 4 * g + 2 * mt + 1];
            }
          }
          for (int tap = 0; tap < 9; ++tap) {
            for (int lane = 0; lane < 32; ++lane) {
              const uint2 v = ld64(l0s + l1_b_off(warp, xg, lane >> 2, lane & 3, tap));
              bfr[lane][0] = v.x; bfr[lane][1] = v.y;
            }
            for (int mt = 0; mt < 2; ++mt) warp_mma_bf16(acc[mt], w1f + (tap * 2 + mt) * 128, bfr);
          }
          const int oy = oy0 + warp;
          if (!pw) {
            for (int lane = 0; lane < 32; ++lane) {
              const int g = lane >> 2, t = lane & 3;
              for (int h = 0; h < 2; ++h) {
                const int px = ox0 + 8 * xg + 2 * t + h;
                if (oy >= oh || px >= ow) continue;
                for (int mt = 0; mt < 2; ++mt) {
                  const uint32_t v = relu_pack_bf16(acc[mt][lane][h], acc[mt][lane][2 + h]);
                  out[((size_t)oy * ow + px) * oc + 4 * g + 2 * mt] = bf2f(v);
                  out[((size_t)oy * ow + px) * oc + 4 * g + 2 * mt + 1] = bf2f(v >> 16);
                }
              }
            }
          } else {
            float acc2[32][4];
            uint32_t bq[2][2][32], tmp[32];
            for (int mt = 0; mt < 2; ++mt)
              for (int hh = 0; hh < 2; ++hh) {
                for (int lane = 0; lane < 32; ++lane) tmp[lane] = relu_pack_bf16(acc[mt][lane][2 * hh], acc[mt][lane][2 * hh + 1]);
                warp_movmatrix_trans(tmp, bq[mt][hh]);
              }
            for (int lane = 0; lane < 32; ++lane) {
              const int g = lane >> 2;
              acc2[lane][0] = acc2[lane][1] = bias[48 + 2 * g];
              acc2[lane][2] = acc2[lane][3] = bias[48 + 2 * g + 1];
            }
            for (int ks = 0; ks < 2; ++ks) {
              for (int lane = 0; lane < 32; ++lane) { bfr[lane][0] = bq[ks][0][lane]; bfr[lane][1] = bq[ks][1][lane]; }
              warp_mma_bf16(acc2, w2f + ks * 128, bfr);
            }
            for (int lane = 0; lane < 32; ++lane) {
              const int g = lane >> 2, t = lane & 3;
              for (int h = 0; h < 2; ++h) {
                const int px = ox0 + 8 * xg + 2 * t + h;
                if (oy >= oh || px >= ow) continue;
                const uint32_t v = relu_pack_bf16(acc2[lane][h], acc2[lane][2 + h]);
                out[((size_t)oy * ow + px) * oc + 2 * g] = bf2f(v);
                out[((size_t)oy * ow + px) * oc + 2 * g + 1] = bf2f(v >> 16);
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
