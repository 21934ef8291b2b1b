// Host emulation of one warp of c3k_flat_kernel (csrc/c3k_flat.cuh): the per-lane address maps, fragment
// permutations and weight packing are the kernel's own (shared header); the mma.sync.m16n8k16 fragment layout,
// the block loops and the barriers are restated here.  Test infrastructure only (tests/test_c3k_flat_emu.py).
// Shared memory is poisoned with NaN before every tile: a wrapped-around or over-read value reaching a valid
// output shows up as NaN.
#include <cmath>
#include <cstdio>

#include "../unina-yolo-dla_b200/csrc/c3k_flat.cuh"

using namespace uyd::c3kf;

static float bf2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

// D += A * B with the PTX fragment layout of mma.m16n8k16 (row.col, bf16 inputs, fp32 accumulate)
static void warp_mma(float acc[32][4], const uint32_t a[32][4], const uint32_t *bfrag /* [32 lanes][2] */) {
  float A[16][16], B[16][8];
  for (int lane = 0; lane < 32; ++lane) {
    const int g = lane >> 2, t = lane & 3;
    const uint32_t *r = a[lane];
    A[g][2 * t] = bf2f(r[0] & 0xffff); A[g][2 * t + 1] = bf2f(r[0] >> 16);
    A[g + 8][2 * t] = bf2f(r[1] & 0xffff); A[g + 8][2 * t + 1] = bf2f(r[1] >> 16);
    A[g][2 * t + 8] = bf2f(r[2] & 0xffff); A[g][2 * t + 9] = bf2f(r[2] >> 16);
    A[g + 8][2 * t + 8] = bf2f(r[3] & 0xffff); A[g + 8][2 * t + 9] = bf2f(r[3] >> 16);
    const uint32_t b0 = bfrag[lane * 2], b1 = bfrag[lane * 2 + 1];
    B[2 * t][g] = bf2f(b0 & 0xffff); B[2 * t + 1][g] = bf2f(b0 >> 16);
    B[2 * t + 8][g] = bf2f(b1 & 0xffff); B[2 * t + 9][g] = bf2f(b1 >> 16);
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

template <int NT>
static void emu_init(float acc[NT][32][4], const float bz[32][NT][2]) {
  for (int lane = 0; lane < 32; ++lane)
    for (int j = 0; j < NT; ++j) {
      acc[j][lane][0] = acc[j][lane][2] = bz[lane][j][0];
      acc[j][lane][1] = acc[j][lane][3] = bz[lane][j][1];
    }
}

template <int C>
static void emu_image(int H, int W, int TH, const float *const w[7], const float *const b[7], const float *x, float *y) {
  using G = Geo<C>;
  constexpr int CC = G::CC;
  std::vector<uint32_t> frags;
  pack_all<C>(w, frags);
  float bias[7 * 32] = {0};
  const int couts[7] = {C, C, C, C, C, C, CC};
  for (int i = 0; i < 7; ++i)
    for (int n = 0; n < couts[i]; ++n) bias[i * 32 + n] = b[i][n];
  const Layout L = make_layout<C>(TH);
  std::vector<unsigned char> smem(L.total);
  for (int ty0 = 0; ty0 < H; ty0 += TH)
    for (int tx0 = 0; tx0 < W; tx0 += kTW) {
      memset(smem.data(), 0xff, smem.size());  // bf16 0xffff = NaN
      unsigned char *X = smem.data() + L.x_off, *A = smem.data() + L.a_off, *Bv = smem.data() + L.b_off;
      uint32_t *maskw = reinterpret_cast<uint32_t *>(smem.data() + L.mask_off);
      const int gy0 = ty0 - 4, gx0 = tx0 - 4;
      for (int i = 0; i < L.FR / 32; ++i) {
        uint32_t m = 0;
        for (int bb = 0; bb < 32; ++bb) {
          const int px = i * 32 + bb, ry = px / kPW, rx = px % kPW;
          if (gy0 + ry >= 0 && gy0 + ry < H && gx0 + rx >= 0 && gx0 + rx < W) m |= 1u << bb;
        }
        maskw[i] = m;
      }
      for (int px = 0; px < L.FR; ++px) {
        const int gy = gy0 + px / kPW, gx = gx0 + px % kPW;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        for (int ch = 0; ch < CC; ++ch)
          reinterpret_cast<uint16_t *>(X)[px * CC + ch] = in ? host_f2bf(x[((size_t)gy * W + gx) * CC + ch]) : 0;
      }
      uint32_t a[32][4];
      // ---- stage 1 ----
      {
        using St = Stage1<C>;
        const int b_lo = kB0 / 32, b_hi = (TH + 4) * kPW / 32;
        float bz[32][St::NT][2];
        for (int lane = 0; lane < 32; ++lane) St::bias_regs(bias, lane, bz[lane]);
        for (int blk = 0; blk < L.FR / 32; ++blk)
          for (int mt = 0; mt < G::MT; ++mt) {
            float acc[St::NT][32][4];
            emu_init<St::NT>(acc, bz);
            for (int s = 0; s < St::KS; ++s) {
              for (int lane = 0; lane < 32; ++lane) St::load_a(X, blk * 32, lane, s, mt, a[lane]);
              for (int j = 0; j < St::NT; ++j) warp_mma(acc[j], a, frags.data() + (s * St::NT + j) * 64);
            }
            for (int lane = 0; lane < 32; ++lane) {
              float la[St::NT][4];
              for (int j = 0; j < St::NT; ++j) memcpy(la[j], acc[j][lane], 16);
              St::store(A, Bv, blk * 32, lane, mt, la, maskw[blk], blk >= b_lo && blk < b_hi);
            }
          }
      }
      // ---- 3x3 stages (each reads the previous stage's buffer; the in-place residual is read by the storing lane) ----
      unsigned char *T = X;
      for (int k = 0; k < 4; ++k) {
        using St = Stage3<C>;
        const unsigned char *S = (k & 1) ? T : A;
        unsigned char *D = (k & 1) ? A : T;
        const uint32_t *wf = frags.data() + G::W1 + k * G::W3;
        int lo, hi;
        conv3_blocks(k, TH, lo, hi);
        float bz[32][St::NT][2];
        for (int lane = 0; lane < 32; ++lane) St::bias_regs(bias + (2 + k) * 32, lane, bz[lane]);
        for (int blk = lo; blk < hi; ++blk)
          for (int mt = 0; mt < G::MT; ++mt) {
            float acc[St::NT][32][4];
            emu_init<St::NT>(acc, bz);
            for (int s = 0; s < St::KS; ++s) {
              for (int lane = 0; lane < 32; ++lane) St::load_a(S, blk * 32, lane, s, mt, a[lane]);
              for (int j = 0; j < St::NT; ++j) warp_mma(acc[j], a, wf + (s * St::NT + j) * 64);
            }
            for (int lane = 0; lane < 32; ++lane) {
              float la[St::NT][4];
              for (int j = 0; j < St::NT; ++j) memcpy(la[j], acc[j][lane], 16);
              if (k & 1) St::template store<true>(D, blk * 32, lane, mt, la, maskw[blk]);
              else St::template store<false>(D, blk * 32, lane, mt, la, maskw[blk]);
            }
          }
      }
      // ---- stage 6 ----
      {
        using St = Stage6<C>;
        const uint32_t *wf = frags.data() + G::W1 + 4 * G::W3;
        float bz[32][St::NT][2];
        for (int lane = 0; lane < 32; ++lane) St::bias_regs(bias + 6 * 32, lane, bz[lane]);
        for (int blk = kB0 / 32; blk < (TH + 4) * kPW / 32; ++blk)
          for (int mt = 0; mt < G::MT; ++mt) {
            float acc[St::NT][32][4];
            emu_init<St::NT>(acc, bz);
            for (int s = 0; s < St::KS; ++s) {
              for (int lane = 0; lane < 32; ++lane) St::load_a(A, Bv, blk * 32, lane, s, mt, a[lane]);
              for (int j = 0; j < St::NT; ++j) warp_mma(acc[j], a, wf + (s * St::NT + j) * 64);
            }
            for (int lane = 0; lane < 32; ++lane) {
              float la[St::NT][4];
              for (int j = 0; j < St::NT; ++j) memcpy(la[j], acc[j][lane], 16);
              St::store(X, blk * 32, lane, mt, la);
            }
          }
      }
      for (int ry = 0; ry < TH; ++ry)
        for (int rx = 0; rx < kTW; ++rx)
          for (int ch = 0; ch < CC; ++ch)
            y[((size_t)(ty0 + ry) * W + tx0 + rx) * CC + ch] = bf2f(reinterpret_cast<const uint16_t *>(X)[(ry * kPW + rx + 4) * CC + ch]);
    }
}

// x, y: [H][W][c] fp32 (x is rounded to bf16 here); w/b as uyd_plan_add_c3k.  Returns 0, or 1 for a bad shape.
extern "C" int c3k_emu(int c, int H, int W, int TH, const float *const w[7], const float *const b[7], const float *x, float *y) {
  if (W % kTW || H % TH || TH % 2) return 1;
  if (c == 8) emu_image<4>(H, W, TH, w, b, x, y);
  else if (c == 16) emu_image<8>(H, W, TH, w, b, x, y);
  else if (c == 32) emu_image<16>(H, W, TH, w, b, x, y);
  else return 1;
  return 0;
}
