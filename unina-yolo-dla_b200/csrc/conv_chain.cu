// Chained tcgen05 kernel for the Detect head (Ultralytics Detect.cv2 / cv3):
//
//     GEMM1 : 3x3 stride-1 conv  Cin -> N1   (HALO mode of conv_tc.cu: one TMA halo box per channel
//             block, nine tap-shifted UMMA descriptors)            + bias + ReLU -> bf16
//     GEMM2 : 1x1 conv           N1  -> N2   A operand = the bf16 tile the epilogue warps wrote to
//             shared memory in the canonical K-major swizzled UMMA layout (never leaves the SM)
//     FINAL : one of
//               STORE  bias (+ReLU) -> bf16 / fp32 NHWC slice               (cls stage 1, raw box logits)
//               PW3    ReLU -> bf16 -> nc <= 8 dot products on CUDA cores   (Detect.cv3[l][2])
//                      -> raw logits into the head slice and / or sigmoid scores into y
//               DFL    softmax-integral over 4 x 16 bins -> (cx, cy, w, h) * stride into y
//                      (Detect._inference + DFL + dist2bbox, same arithmetic as decode.cu)
//
// Box branch  : Conv(64,64,3) -> Conv2d(64,64,1) -> DFL                    = one launch, the 64 logits
//               of an anchor stay in TMEM / registers.
// Class branch: DWConv(c,c,3) + its 1x1 conv: [dw1, pw1] -> z1 (bf16) and [dw2, pw2, pw3] -> logits / scores.
//               The depth-wise 3x3 runs on the CUDA cores of the epilogue groups (template DW): a thread owns one
//               channel quad of one tile column, keeps its 36 weights in registers, walks the TMA halo box down
//               the column (three 8-byte loads per input row, packed fp32 FFMA2) and writes the bf16 result
//               straight into the swizzled K-major tile GEMM2 consumes.  (Round 1 ran it as a dense GEMM with a
//               diagonal weight matrix: 18 tcgen05.mma per tile on a 97 %-zero operand, 1550 cycles per tile.)
//
// Warp roles (512 threads, persistent CTA per SM): warp 0 TMA producer, warp 1 TMEM allocator + GEMM1
// issuer, warp 2 GEMM2 issuer, warps 4-15 three epilogue groups taking tiles round-robin.  GEMM1 of the
// following tiles is issued while a tile's bf16 operand is being written, so the tensor pipe never
// waits for the epilogue.
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace uyd {

enum { CH_STORE = 0, CH_PW3 = 2, CH_DFL = 3 };

struct ChainParams {
  int H, W, n0, nb;
  int tiles_x, tiles_y;
  long long total_tiles;
  int ncb, cb_bytes;  // GEMM1: input channel blocks, bytes per block row (64 / 128)
  int N1, N2;         // UMMA N of the two GEMMs
  int c2;             // real output channels of GEMM2 (STORE)
  int stages;
  uint32_t blk_bytes, tx_bytes, w1_bytes, w2_bytes, a2_bytes;
  uint32_t idesc1, idesc2, layout1, layout2;
  int relu2, final_kind, nc;
  const float *bias1, *bias2, *bias3, *w3;  // w3: fp32 [nc][N2] holding bf16-rounded values
  const float *dw_w;                        // DW: depth-wise weights fp32 [9][C] holding bf16-rounded values
  void *out;                                // raw store target (slice base of image 0) or null
  int out_pitch, out_f32;
  float *y;                                 // decoded output [B, no, a_total] or null
  int a_total, a_off, y_ch0, no;
  float stride_px;
  unsigned tpi, m_tpi, m_tx;  // tiles per image and the fastdiv magics of tiles-per-image / tiles_x
  long long *dbg;  // UYD_CHAIN_TIMELINE: clock64 stamps of CTA 0, [tile][8]
};

namespace {

constexpr int kNG = 3;                              // epilogue groups (4 warps each) taking tiles round-robin
constexpr int kChainThreads = 128 + 128 * kNG;       // TMA, GEMM1 issuer, GEMM2 issuer, (idle), epilogue groups
constexpr int kHaloRows = 18, kTileH = 16, kTileW = 8, kHaloPitch = kTileW + 2;

// ex2.approx + rcp.approx (2 ulp): identical to decode.cu's, so the fused and the stand-alone decode agree bit for bit
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// packed fp32 FMA (FFMA2 on sm_100): d = a * b + c on both halves
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

// 16 consecutive floats of a shared-memory vector as four 16-byte broadcast loads
__device__ __forceinline__ void load_bias16(const float *s, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 q = reinterpret_cast<const float4 *>(s)[i];
    v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
  }
}

__device__ __forceinline__ void stamp(const ChainParams &p, int it, int slot, int lane) {
  if (p.dbg && blockIdx.x == 0 && lane == 0 && it < 64) p.dbg[(slot >> 3) * 512 + it * 8 + (slot & 7)] = clock64();
}

// KS1 / KS2 = 32-byte k-steps per channel-block row of GEMM1 / GEMM2 (cb_bytes / 32, N1 * 2 / 32)
// DW: the first conv is depth-wise (C = 16 * KS1 channels = N1) and runs on the CUDA cores of the epilogue groups
template <int KS1, int KS2, bool DW>
__global__ void __launch_bounds__(kChainThreads, 1) conv_chain_kernel(const __grid_constant__ CUtensorMap tm_in,
                                                                      const __grid_constant__ CUtensorMap tm_w1,
                                                                      const __grid_constant__ CUtensorMap tm_w2,
                                                                      const ChainParams p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t w1_s = base;
  const uint32_t w2_s = w1_s + ((p.w1_bytes + 1023u) & ~1023u);
  const uint32_t a2_s = w2_s + ((p.w2_bytes + 1023u) & ~1023u);
  const uint32_t a_s = a2_s + (uint32_t)kNG * p.a2_bytes;
  const uint32_t bar0 = a_s + (uint32_t)p.stages * p.blk_bytes;
  // barriers: full[8] empty[8] wfull tfull1[4] tempty1[4] a2full[4] tfull2[4] tempty2[4] | slot | floats
  const uint32_t full0 = bar0, empty0 = bar0 + 64, wfull = bar0 + 128;
  const uint32_t tfull1 = bar0 + 136, tempty1 = bar0 + 168, a2full = bar0 + 200, tfull2 = bar0 + 232, tempty2 = bar0 + 264;
  const uint32_t slot = bar0 + 296;
  uint32_t *slot_ptr = reinterpret_cast<uint32_t *>(smem_dyn + (slot - raw));
  float *fs = reinterpret_cast<float *>(smem_dyn + (bar0 + 320u - raw));
  float *bias1_s = fs, *bias2_s = fs + 128, *bias3_s = fs + 256, *w3_s = fs + 264;  // w3_s: [8][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)kNG * (uint32_t)(p.N1 + p.N2)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8u * s, 1);
      mbar_init(empty0 + 8u * s, DW ? 128 : 1);  // DW: released by the 128 threads that read the halo box
    }
    mbar_init(wfull, 1);
    for (int g = 0; g < kNG; ++g) {
      mbar_init(tfull1 + 8u * g, 1);
      mbar_init(tempty1 + 8u * g, 128);
      mbar_init(a2full + 8u * g, 128);
      mbar_init(tfull2 + 8u * g, 1);
      mbar_init(tempty2 + 8u * g, 128);
    }
    fence_barrier_init();
    // weights: part of the prologue that overlaps the previous kernel's tail
    mbar_expect_tx(wfull, p.w1_bytes + p.w2_bytes);
    const int nblk = DW ? 0 : p.ncb * 9;
    for (int i = 0; i < nblk; ++i) tma_load_2d(w1_s + (uint32_t)i * p.N1 * p.cb_bytes, &tm_w1, wfull, 0, i * p.N1);
    tma_load_2d(w2_s, &tm_w2, wfull, 0, 0);
  }
  griddep_trigger();
  for (int i = threadIdx.x; i < 128; i += kChainThreads) {
    bias1_s[i] = i < p.N1 ? p.bias1[i] : 0.f;
    bias2_s[i] = i < p.N2 ? p.bias2[i] : 0.f;
  }
  if (p.final_kind == CH_PW3) {
    for (int i = threadIdx.x; i < 8 * 64; i += kChainThreads) {  // shared layout [channel pair][8 classes][2]: the FFMA2 operand
      const int k = i >> 3, c = i & 7;                           // pairs (even, odd channel) of two classes per 16-byte load
      w3_s[((k >> 1) * 8 + c) * 2 + (k & 1)] = (c < p.nc && k < p.N2) ? p.w3[c * p.N2 + k] : 0.f;
    }
    if (threadIdx.x < 8) bias3_s[threadIdx.x] = threadIdx.x < p.nc ? p.bias3[threadIdx.x] : 0.f;
  }
  if (warp == 1) tmem_alloc(slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot_ptr;
  griddep_wait();  // from here on the activations written by the previous kernel are read

  const int stages = p.stages;
  const uint32_t blk_bytes = p.blk_bytes, cb_bytes = p.cb_bytes;
  const uint32_t cb2_bytes = (uint32_t)p.N1 * 2u;  // one row of the GEMM2 A / B operands

  if (warp == 0) {
    // ================= TMA producer =================
    int stage = 0;
    uint32_t phase = 0;
    const int cb_elems = (int)cb_bytes / 2;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const unsigned ut = (unsigned)tile, img = fastdiv(ut, p.tpi, p.m_tpi), t = ut - img * p.tpi;
      const unsigned ty = fastdiv(t, (unsigned)p.tiles_x, p.m_tx);
      const int n = p.n0 + (int)img;
      const int y0 = (int)ty * kTileH, x0 = (int)(t - ty * (unsigned)p.tiles_x) * kTileW;
      for (int j = 0; j < p.ncb; ++j) {
        mbar_wait(empty0 + 8u * stage, phase ^ 1u);
        if (lane == 0) {
          const uint32_t fb = full0 + 8u * stage;
          mbar_expect_tx(fb, p.tx_bytes);
          tma_load_4d(a_s + (uint32_t)stage * blk_bytes, &tm_in, fb, j * cb_elems, x0 - 1, y0 - 1, n);
          if (p.dbg && blockIdx.x == 0) { const int itp = (int)((tile - blockIdx.x) / gridDim.x); if (itp < 64) p.dbg[itp * 8 + 0] = clock64(); }
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && !DW) {
    // ================= GEMM1 issuer =================
    // tcgen05.mma issue is nearly synchronous with execution (the queue is a few instructions deep), so
    // every cycle this thread spends polling barriers is a tensor-pipe bubble: GEMM2 has its own issuing
    // warp, whose waits overlap GEMM1's execution.
    mbar_wait(wfull, 0);
    int stage = 0;
    uint32_t phase = 0;
    const uint64_t adesc0 = make_desc_base((uint32_t)kHaloPitch * cb_bytes, p.layout1);
    const uint64_t bdesc0 = make_desc_base(8u * cb_bytes, p.layout1);
    const uint32_t wblk_units = ((uint32_t)p.N1 * cb_bytes) >> 4;
    const uint32_t px_units = cb_bytes >> 4, row_units = (uint32_t)kHaloPitch * px_units;
    int it = 0, g = 0;
    uint32_t gph = 0;  // phase of this group's barriers
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      mbar_wait(tempty1 + 8u * g, gph ^ 1u);
      const uint32_t d1 = tmem_base + (uint32_t)g * p.N1;
      for (int j = 0; j < p.ncb; ++j) {
        mbar_wait(full0 + 8u * stage, phase);
        tc_fence_after();
        stamp(p, it, 1, lane);  // GEMM1 operands landed, issue starts
        const uint64_t ablk_d = adesc0 + (uint64_t)(((a_s + (uint32_t)stage * blk_bytes) & 0x3FFFFu) >> 4);
        const uint64_t wd0 = bdesc0 + (uint64_t)(((w1_s + (uint32_t)(j * 9) * p.N1 * cb_bytes) & 0x3FFFFu) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint64_t ad = ablk_d + (uint64_t)((t / 3) * row_units + (t % 3) * px_units);
            const uint64_t wd = wd0 + (uint64_t)t * wblk_units;
#pragma unroll
            for (int k = 0; k < KS1; ++k) umma_bf16(d1, ad + 2 * k, wd + 2 * k, p.idesc1, (j != 0 || t != 0 || k != 0) ? 1u : 0u);
          }
          umma_commit(empty0 + 8u * stage);
          if (j == p.ncb - 1) umma_commit(tfull1 + 8u * g);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
      if (++g == kNG) { g = 0; gph ^= 1u; }
    }
  } else if (warp == 2) {
    // ================= GEMM2 issuer =================
    mbar_wait(wfull, 0);
    const uint64_t a2desc0 = make_desc_base(8u * cb2_bytes, p.layout2);
    const uint64_t w2d = a2desc0 + (uint64_t)((w2_s & 0x3FFFFu) >> 4);
    int it = 0, g = 0;
    uint32_t gph = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      mbar_wait(tempty2 + 8u * g, gph ^ 1u);   // the previous result of this accumulator has been drained
      mbar_wait(a2full + 8u * g, gph);         // the epilogue group has written its bf16 tile
      tc_fence_after();
      stamp(p, it, 2, lane);  // GEMM2 issue
      if (elect_one()) {
        const uint32_t d2 = tmem_base + (uint32_t)kNG * (uint32_t)p.N1 + (uint32_t)g * p.N2;
        const uint64_t ad = a2desc0 + (uint64_t)(((a2_s + (uint32_t)g * p.a2_bytes) & 0x3FFFFu) >> 4);
#pragma unroll
        for (int k = 0; k < KS2; ++k) umma_bf16(d2, ad + 2 * k, w2d + 2 * k, p.idesc2, k != 0);
        umma_commit(tfull2 + 8u * g);
      }
      __syncwarp();
      if (++g == kNG) { g = 0; gph ^= 1u; }
    }
  } else if (warp >= 4) {
    // ================= epilogue (two groups of four warps, alternate tiles) =================
    const int q = warp & 3;
    const int m = q * 32 + lane;         // accumulator row = pixel of the 16 x 8 tile
    const int g = (warp - 4) >> 2;
    const int nch1 = p.N1 >> 4;
    unsigned char *a2_row = smem_dyn + (a2_s + (uint32_t)g * p.a2_bytes - raw) + (size_t)m * cb2_bytes;
    // 16-byte chunk c of row m lives at chunk (c ^ swz) (TMA / UMMA swizzle: address bits [4,7) ^= bits [7,10),
    // restricted to the row width)
    const uint32_t swz = cb2_bytes == 128 ? (uint32_t)(m & 7) : (cb2_bytes == 64 ? (uint32_t)((m >> 1) & 3) : (uint32_t)((m >> 2) & 1));
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // DW: this thread's channel quad (4 channels x 9 taps) and bias, in registers for the whole launch
    float2 dw_wlo[9], dw_whi[9], dw_blo = make_float2(0.f, 0.f), dw_bhi = dw_blo;
    uint32_t dw_in_off[3][4] = {};
    int dw_in_base = 0, dw_out_base = 0;
    int dw_stage = g % stages;                       // pipeline stage / phase of this group's current tile (tile g, g + kNG, ...)
    uint32_t dw_phase = (uint32_t)(g / stages) & 1u;
    if constexpr (DW) {
      constexpr int C = KS1 * 16;
      const int qd = m % (C / 4);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float4 w4 = *reinterpret_cast<const float4 *>(p.dw_w + t * C + 4 * qd);
        dw_wlo[t] = make_float2(w4.x, w4.y);
        dw_whi[t] = make_float2(w4.z, w4.w);
      }
      const float4 b4 = *reinterpret_cast<const float4 *>(p.bias1 + 4 * qd);
      dw_blo = make_float2(b4.x, b4.y);
      dw_bhi = make_float2(b4.z, b4.w);
      // Shared-memory addressing of the depth-wise stage, hoisted out of the tile loop.  Halo pixel index of (input row
      // i, column dx) = p0 + 10 i + dx with p0 = 10 R half + xx; the TMA / UMMA swizzle XORs the 16-byte chunk index
      // with (pixel >> 1) & 3 (64-byte rows) or pixel & 7 (128-byte rows), and 10 i = 2 i (mod 8): the phase only
      // depends on i & 3, so twelve byte offsets per thread cover every load and the row term is an immediate.
      constexpr int NQ = C / 4, R = 16 / (128 / (NQ * 8));
      const int xx = (m / NQ) & 7, half = m / (NQ * 8);
      const int p0 = half * R * kHaloPitch + xx;
      dw_in_base = p0 * (C * 2);
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
          const int pi = p0 + dx + kHaloPitch * ph;  // any row with i & 3 == ph has the same swizzle phase
          const uint32_t sw = C == 64 ? (uint32_t)(pi & 7) : (uint32_t)((pi >> 1) & 3);
          dw_in_off[dx][ph] = (uint32_t)(dx * (C * 2)) + ((((uint32_t)qd >> 1) ^ sw) << 4) + (uint32_t)(qd & 1) * 8u;
        }
      // output pixel pm = 8 (R half + r) + xx: its phase (pm & 7, resp. (pm >> 1) & 3) is xx's -> one constant + r * immediate
      const int pm0 = half * R * 8 + xx;
      const uint32_t sw2 = C == 64 ? (uint32_t)(pm0 & 7) : (uint32_t)((pm0 >> 1) & 3);
      dw_out_base = pm0 * (C * 2) + (int)(((((uint32_t)qd >> 1) ^ sw2) << 4) + (uint32_t)(qd & 1) * 8u);
    }
    // group g takes tiles g, g + kNG, ... of this CTA's sequence; tile indices fit 32 bits (checked on the host)
    int it = g;
    uint32_t ph = 0;
    for (long long tile = blockIdx.x + (long long)g * gridDim.x; tile < p.total_tiles; tile += (long long)kNG * gridDim.x, it += kNG, ph ^= 1u) {
      const unsigned ut = (unsigned)tile, img = fastdiv(ut, p.tpi, p.m_tpi), t = ut - img * p.tpi;
      const unsigned ty = fastdiv(t, (unsigned)p.tiles_x, p.m_tx), tx = t - ty * (unsigned)p.tiles_x;
      const int n = p.n0 + (int)img;
      const int oy = (int)ty * kTileH + (m >> 3), ox = (int)tx * kTileW + (m & 7);
      const bool inside = oy < p.H && ox < p.W;
      if constexpr (DW) {
        // ---- stage 1 (depth-wise): halo box -> 3x3 depth-wise conv on CUDA cores -> bias, ReLU -> bf16 -> A of GEMM2 ----
        constexpr int C = KS1 * 16, R = 16 / (128 / ((C / 4) * 8));  // rows per thread: 8 (C = 32) / 16 (C = 64)
        mbar_wait(full0 + 8u * dw_stage, dw_phase);
        const int sidx = dw_stage;
        const unsigned char *hb = smem_dyn + (a_s + (uint32_t)sidx * blk_bytes - raw) + dw_in_base;
        unsigned char *a2_out = smem_dyn + (a2_s + (uint32_t)g * p.a2_bytes - raw) + dw_out_base;
        float2 a0[3], a1[3];  // accumulators of three output rows in flight, ring-indexed by row % 3
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
          float2 v0[3], v1[3];
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {  // address = tile base + per-thread constant (swizzle phase i & 3) + immediate
            const uint2 h2 = *reinterpret_cast<const uint2 *>(hb + i * (kHaloPitch * C * 2) + dw_in_off[dx][i & 3]);
            v0[dx] = make_float2(__uint_as_float(h2.x << 16), __uint_as_float(h2.x & 0xffff0000u));
            v1[dx] = make_float2(__uint_as_float(h2.y << 16), __uint_as_float(h2.y & 0xffff0000u));
          }
#pragma unroll
          for (int tr = 0; tr < 3; ++tr) {  // input row i is tap row tr of output row i - tr
            const int r = i - tr;
            if (r < 0 || r >= R) continue;
            float2 &o0 = a0[r % 3], &o1 = a1[r % 3];
            if (tr == 0) { o0 = dw_blo; o1 = dw_bhi; }
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              o0 = ffma2(v0[dx], dw_wlo[tr * 3 + dx], o0);
              o1 = ffma2(v1[dx], dw_whi[tr * 3 + dx], o1);
            }
            if (tr == 2)
              *reinterpret_cast<uint2 *>(a2_out + r * (8 * C * 2)) = make_uint2(relu_pack_bf16x2(o0.x, o0.y), relu_pack_bf16x2(o1.x, o1.y));
          }
        }
        mbar_arrive(empty0 + 8u * sidx);  // this thread's reads of the halo box are done
#pragma unroll
        for (int w = 0, nx = dw_stage + kNG; w < kNG; ++w) {  // this group's next tile is kNG tiles on: at most kNG wraps
          if (nx >= stages) { nx -= stages; dw_phase ^= 1u; }
          dw_stage = nx;
        }
      } else {
      // ---- stage 1: acc1 -> bias, ReLU -> bf16 -> swizzled shared-memory tile (A of GEMM2) ----
        mbar_wait(tfull1 + 8u * g, ph);
        tc_fence_after();
        if (q == 0) stamp(p, it, 3, lane);  // accumulator 1 complete
        {
          const uint32_t t1 = lane_base + (uint32_t)g * p.N1;
          uint32_t cur[16], nxt[16];
          tmem_ld16_issue(t1, cur);
          tmem_ld_wait();
          for (int c = 0; c < nch1; ++c) {
            const bool more = c + 1 < nch1;
            if (more) {
              tmem_ld16_issue(t1 + 16u * (c + 1), nxt);
            } else {
              tc_fence_before();
              mbar_arrive(tempty1 + 8u * g);
            }
            uint4 o0, o1;
            uint32_t *w0 = reinterpret_cast<uint32_t *>(&o0), *w1 = reinterpret_cast<uint32_t *>(&o1);
            float bv[16];
            load_bias16(bias1_s + c * 16, bv);
#pragma unroll
            for (int i = 0; i < 4; ++i) {  // bias -> ReLU -> bf16 pair in one cvt.rn.relu.bf16x2
              w0[i] = relu_pack_bf16x2(__uint_as_float(cur[2 * i]) + bv[2 * i], __uint_as_float(cur[2 * i + 1]) + bv[2 * i + 1]);
              w1[i] = relu_pack_bf16x2(__uint_as_float(cur[8 + 2 * i]) + bv[8 + 2 * i], __uint_as_float(cur[8 + 2 * i + 1]) + bv[8 + 2 * i + 1]);
            }
            *reinterpret_cast<uint4 *>(a2_row + (((uint32_t)(2 * c) ^ swz) << 4)) = o0;
            *reinterpret_cast<uint4 *>(a2_row + (((uint32_t)(2 * c + 1) ^ swz) << 4)) = o1;
            if (more) {
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
            }
          }
        }
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
      mbar_arrive(a2full + 8u * g);
      if (q == 0) stamp(p, it, 4, lane);  // bf16 tile written
      // ---- stage 2: acc2 -> final ----
      mbar_wait(tfull2 + 8u * g, ph);
      tc_fence_after();
      if (q == 0) stamp(p, it, 5, lane);  // accumulator 2 complete
      const uint32_t t2 = lane_base + (uint32_t)kNG * (uint32_t)p.N1 + (uint32_t)g * p.N2;
      const long long pix = ((long long)n * p.H + oy) * p.W + ox;
      const int nch2 = p.N2 >> 4;
      if (p.final_kind == CH_DFL) {
        float d[4];
        uint32_t cur[16], nxt[16];
        tmem_ld16_issue(t2, cur);
        tmem_ld_wait();
#pragma unroll
        for (int side = 0; side < 4; ++side) {
          if (side < 3) {
            tmem_ld16_issue(t2 + 16u * (side + 1), nxt);  // in flight while this side is reduced
          } else {
            tc_fence_before();
            mbar_arrive(tempty2 + 8u * g);
          }
          float v[16];
          load_bias16(bias2_s + side * 16, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(cur[i]);
          if (p.out && inside) {
            float *op = reinterpret_cast<float *>(p.out) + pix * p.out_pitch + side * 16;
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4 *>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
          float mx = v[0];
#pragma unroll
          for (int i = 1; i < 16; ++i) mx = fmaxf(mx, v[i]);
          float s = 0.f, ws = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float e = __expf(v[i] - mx);  // ex2.approx: the 16 softmax weights need no more (same in decode.cu)
            s += e;
            ws = fmaf((float)i, e, ws);
          }
          d[side] = ws / s;
          if (side < 3) {
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
          }
        }
        if (p.y && inside) {
          const float ax = (float)ox + 0.5f, ay = (float)oy + 0.5f;
          const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
          float *yo = p.y + ((long long)n * p.no + p.y_ch0) * p.a_total + p.a_off + oy * p.W + ox;
          yo[0] = (x1 + x2) * 0.5f * p.stride_px;
          yo[(long long)p.a_total] = (y1 + y2) * 0.5f * p.stride_px;
          yo[2ll * p.a_total] = (x2 - x1) * p.stride_px;
          yo[3ll * p.a_total] = (y2 - y1) * p.stride_px;
        }
      } else if (p.final_kind == CH_PW3) {
        // logits: acc[c] = (sum over even channels, sum over odd channels) of z * w3[c]; a bf16 pair of z IS an FFMA2
        // operand, so one packed FMA covers two channels of one class (64 instead of 128 FMAs for nc <= 4)
        float2 acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = make_float2(bias3_s[c], 0.f);
        uint32_t cur[16], nxt[16];
        tmem_ld16_issue(t2, cur);
        tmem_ld_wait();
        for (int ch = 0; ch < nch2; ++ch) {
          const bool more = ch + 1 < nch2;
          if (more) {
            tmem_ld16_issue(t2 + 16u * (ch + 1), nxt);
          } else {
            tc_fence_before();
            mbar_arrive(tempty2 + 8u * g);
          }
          float bv[16];
          load_bias16(bias2_s + ch * 16, bv);
          // z2 is a bf16 activation in the unfused graph: ReLU + round it the same way before the last conv
          // (one cvt.rn.relu.bf16x2 per pair, unpacked with a shift / mask)
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const uint32_t pk = relu_pack_bf16x2(__uint_as_float(cur[i]) + bv[i], __uint_as_float(cur[i + 1]) + bv[i + 1]);
            const float2 zz = make_float2(__uint_as_float(pk << 16), __uint_as_float(pk & 0xffff0000u));
            const float4 *wp = reinterpret_cast<const float4 *>(w3_s + (ch * 8 + (i >> 1)) * 16);
            const float4 w01 = wp[0], w23 = wp[1];
            acc[0] = ffma2(zz, make_float2(w01.x, w01.y), acc[0]);
            acc[1] = ffma2(zz, make_float2(w01.z, w01.w), acc[1]);
            acc[2] = ffma2(zz, make_float2(w23.x, w23.y), acc[2]);
            acc[3] = ffma2(zz, make_float2(w23.z, w23.w), acc[3]);
            if (p.nc > 4) {
              const float4 w45 = wp[2], w67 = wp[3];
              acc[4] = ffma2(zz, make_float2(w45.x, w45.y), acc[4]);
              acc[5] = ffma2(zz, make_float2(w45.z, w45.w), acc[5]);
              acc[6] = ffma2(zz, make_float2(w67.x, w67.y), acc[6]);
              acc[7] = ffma2(zz, make_float2(w67.z, w67.w), acc[7]);
            }
          }
          if (more) {
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
          }
        }
        float lg[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) lg[c] = acc[c].x + acc[c].y;
        if (inside) {
          if (p.out) {
            float *op = reinterpret_cast<float *>(p.out) + pix * p.out_pitch;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c < p.nc) op[c] = lg[c];
          }
          if (p.y) {
            float *yo = p.y + ((long long)n * p.no + p.y_ch0) * p.a_total + p.a_off + oy * p.W + ox;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c < p.nc) yo[(long long)c * p.a_total] = sigmoidf_(lg[c]);
          }
        }
      } else {  // CH_STORE: bias (+ReLU) -> bf16 / fp32 row of the NHWC slice
        for (int ch = 0; ch < nch2; ++ch) {
          uint32_t r[16];
          tmem_ld16_issue(t2 + 16u * ch, r);
          tmem_ld_wait();
          if (ch == nch2 - 1) {
            tc_fence_before();
            mbar_arrive(tempty2 + 8u * g);
          }
          if (!inside || ch * 16 >= p.c2) continue;
          float v[16];
          load_bias16(bias2_s + ch * 16, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x = __uint_as_float(r[i]) + v[i];
            v[i] = p.relu2 ? fmaxf(x, 0.f) : x;
          }
          if (p.out_f32) {
            float *op = reinterpret_cast<float *>(p.out) + pix * p.out_pitch + ch * 16;
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4 *>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
            __nv_bfloat16 *op = reinterpret_cast<__nv_bfloat16 *>(p.out) + pix * p.out_pitch + ch * 16;
            uint4 o0, o1;
            uint32_t *w0 = reinterpret_cast<uint32_t *>(&o0), *w1 = reinterpret_cast<uint32_t *>(&o1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __nv_bfloat162 ha = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
              __nv_bfloat162 hb = __floats2bfloat162_rn(v[8 + 2 * i], v[8 + 2 * i + 1]);
              w0[i] = *reinterpret_cast<uint32_t *>(&ha);
              w1[i] = *reinterpret_cast<uint32_t *>(&hb);
            }
            *reinterpret_cast<uint4 *>(op) = o0;
            *reinterpret_cast<uint4 *>(op + 8) = o1;
          }
        }
      }
      if (q == 0) stamp(p, it, 6, lane);  // final stage done
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn chain_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int chain_encode(CUtensorMap *tm, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides_bytes,
                 const cuuint32_t *box, int row_bytes) {
  EncodeTiledFn fn = chain_encode_fn();
  UYD_REQUIRE(fn, UYD_E_NOGPU, "cuTensorMapEncodeTiled is not available (no CUDA driver)");
  const cuuint32_t one4[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                 : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides_bytes, box, one4,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UYD_REQUIRE(r == CUDA_SUCCESS, UYD_E_ARG, "conv_chain: cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
  return UYD_OK;
}

uint32_t layout_code(int row_bytes) { return row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u); }

}  // namespace

struct ChainConv {
  CUtensorMap tm_in, tm_w1, tm_w2;
  ChainParams p;
  size_t smem;
  bool dw = false;
};

ChainConv *chain_new() { return new ChainConv(); }
void chain_delete(ChainConv *c) { delete c; }

// Shape rule: cin in {32, 64} (one channel block), n1 in {32, 64}, n2 <= 64.
bool chain_supported(int cin, int n1, int n2, int in_pitch, int in_coff) {
  if (!(cin == 32 || cin == 64) || !(n1 == 32 || n1 == 64) || n2 < 1 || n2 > 64) return false;
  return in_pitch % 8 == 0 && in_coff % 8 == 0;
}

size_t chain_w1_bytes(int cin, int n1, bool depthwise) { return depthwise ? (size_t)9 * cin * 4 : (size_t)9 * cin * n1 * 2; }
size_t chain_w2_bytes(int n1, int n2) { return (size_t)n1 * ((n2 + 15) / 16 * 16) * 2; }

// w1: [n1][cin][3][3] fp32 (dense) or, when depthwise, [n1][1][3][3] expanded to a diagonal dense matrix.
// Layout: bf16 [tap][n (N1)][cin]   (one channel block)
//         depth-wise: fp32 [tap][c] holding the bf16-rounded weights (CUDA-core stage of the DW kernel)
void chain_pack_w1(int cin, int n1, bool depthwise, const float *w, void *dst_host) {
  if (depthwise) {
    float *f = reinterpret_cast<float *>(dst_host);
    for (int t = 0; t < 9; ++t)
      for (int c = 0; c < cin; ++c) f[t * cin + c] = __bfloat162float(__float2bfloat16_rn(w[(size_t)c * 9 + t]));
    return;
  }
  __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(dst_host);
  for (int t = 0; t < 9; ++t)
    for (int n = 0; n < n1; ++n)
      for (int c = 0; c < cin; ++c) {
        o[((size_t)t * n1 + n) * cin + c] = __float2bfloat16_rn(w[((size_t)n * cin + c) * 9 + t]);
      }
}

// w2: [n2][n1] fp32 -> bf16 [N2 padded][n1]
void chain_pack_w2(int n1, int n2, const float *w, void *dst_host) {
  __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(dst_host);
  const int N2 = (n2 + 15) / 16 * 16;
  for (int n = 0; n < N2; ++n)
    for (int k = 0; k < n1; ++k) o[(size_t)n * n1 + k] = __float2bfloat16_rn(n < n2 ? w[(size_t)n * n1 + k] : 0.f);
}

int chain_prepare(ChainConv *cc, int cin, int n1, int n2, void *in_base, int in_pitch, int h, int w, int max_batch,
                  void *w1_dev, void *w2_dev, const float *bias1, const float *bias2, int relu2, int final_kind, int nc,
                  const float *w3_dev, const float *bias3_dev, void *out_base, int out_pitch, int out_f32, int a_total,
                  int a_off, int y_ch0, int no, float stride_px, int dw1) {
  ChainParams &p = cc->p;
  memset(&p, 0, sizeof(p));
  cc->dw = dw1 != 0;
  p.dw_w = dw1 ? reinterpret_cast<const float *>(w1_dev) : nullptr;
  p.H = h; p.W = w;
  p.ncb = 1;
  p.cb_bytes = cin * 2;
  p.N1 = n1;
  p.N2 = (n2 + 15) / 16 * 16;
  p.c2 = n2;
  UYD_REQUIRE(final_kind != CH_DFL || p.N2 == 64, UYD_E_UNSUPPORTED, "conv_chain: the DFL final needs 4 x 16 logits");
  UYD_REQUIRE(final_kind != CH_PW3 || (nc >= 1 && nc <= 8 && w3_dev && bias3_dev), UYD_E_UNSUPPORTED, "conv_chain: PW3 needs nc <= 8");
  p.tiles_x = ceil_div(w, kTileW);
  p.tiles_y = ceil_div(h, kTileH);
  p.tpi = (unsigned)(p.tiles_x * p.tiles_y);
  p.m_tpi = fastdiv_magic(p.tpi);
  p.m_tx = fastdiv_magic((unsigned)p.tiles_x);
  p.blk_bytes = ((uint32_t)kHaloRows * kHaloPitch * p.cb_bytes + 1023u) & ~1023u;
  p.tx_bytes = (uint32_t)kHaloRows * kHaloPitch * p.cb_bytes;
  p.w1_bytes = dw1 ? 0u : (uint32_t)chain_w1_bytes(cin, n1, false);
  p.w2_bytes = (uint32_t)chain_w2_bytes(n1, n2);
  p.a2_bytes = 128u * (uint32_t)n1 * 2u;
  p.layout1 = layout_code(p.cb_bytes);
  p.layout2 = layout_code(n1 * 2);
  p.idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N2 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.relu2 = relu2;
  p.final_kind = final_kind;
  p.nc = nc;
  p.bias1 = bias1; p.bias2 = bias2; p.bias3 = bias3_dev; p.w3 = w3_dev;
  p.out = out_base; p.out_pitch = out_pitch; p.out_f32 = out_f32;
  p.a_total = a_total; p.a_off = a_off; p.y_ch0 = y_ch0; p.no = no; p.stride_px = stride_px;
  const size_t fixed = 1024 + ((p.w1_bytes + 1023u) & ~1023u) + ((p.w2_bytes + 1023u) & ~1023u) + (size_t)kNG * p.a2_bytes + 320 + 4096;
  int stages = (int)((227 * 1024 - fixed) / p.blk_bytes);
  UYD_REQUIRE(stages >= 2, UYD_E_UNSUPPORTED, "conv_chain: %d -> %d -> %d leaves no room for two halo stages", cin, n1, n2);
  if (stages > 6) stages = 6;
  p.stages = stages;
  cc->smem = fixed + (size_t)stages * p.blk_bytes;
  if (!dw1) {
    const cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)9 * n1};
    const cuuint64_t str[1] = {(cuuint64_t)p.cb_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)cin, (cuuint32_t)n1};
    int e = chain_encode(&cc->tm_w1, w1_dev, 2, dims, str, box, p.cb_bytes);
    if (e) return e;
  } else {
    cc->tm_w1 = cc->tm_w2;  // unused by the DW kernel
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)n1, (cuuint64_t)p.N2};
    const cuuint64_t str[1] = {(cuuint64_t)n1 * 2};
    const cuuint32_t box[2] = {(cuuint32_t)n1, (cuuint32_t)p.N2};
    int e = chain_encode(&cc->tm_w2, w2_dev, 2, dims, str, box, n1 * 2);
    if (e) return e;
  }
  {
    const cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)max_batch};
    const cuuint64_t str[3] = {(cuuint64_t)in_pitch * 2, (cuuint64_t)w * in_pitch * 2, (cuuint64_t)h * w * in_pitch * 2};
    const cuuint32_t box[4] = {(cuuint32_t)cin, (cuuint32_t)kHaloPitch, (cuuint32_t)kHaloRows, 1};
    int e = chain_encode(&cc->tm_in, in_base, 4, dims, str, box, p.cb_bytes);
    if (e) return e;
  }
  if (int e = smem_optin(conv_chain_kernel<2, 2, false>, 227 * 1024)) return e;
  if (int e = smem_optin(conv_chain_kernel<2, 4, false>, 227 * 1024)) return e;
  if (int e = smem_optin(conv_chain_kernel<4, 2, false>, 227 * 1024)) return e;
  if (int e = smem_optin(conv_chain_kernel<4, 4, false>, 227 * 1024)) return e;
  if (int e = smem_optin(conv_chain_kernel<2, 2, true>, 227 * 1024)) return e;
  if (int e = smem_optin(conv_chain_kernel<4, 4, true>, 227 * 1024)) return e;
  return UYD_OK;
}

int chain_launch(const ChainConv *cc, int nb, float *y, int sm_count, cudaStream_t s) {
  ChainParams p = cc->p;
  p.n0 = 0;
  p.nb = nb;
  p.y = y;
  p.total_tiles = (long long)nb * p.tiles_x * p.tiles_y;
  if (p.total_tiles == 0) return UYD_OK;
  UYD_REQUIRE(p.total_tiles < (1ll << 31), UYD_E_UNSUPPORTED, "conv_chain: %lld tiles exceed the kernel's 32-bit tile index", p.total_tiles);
  const unsigned grid = (unsigned)(p.total_tiles < sm_count ? p.total_tiles : sm_count);
  const int ks1 = p.cb_bytes / 32, ks2 = p.N1 * 2 / 32;
#ifdef UYD_CHAIN_TIMELINE_BUILD  // debug build only (tools/chain_timeline.py): never in the shipped library path
  static long long *dbg_dev = nullptr;
  const bool timeline = getenv("UYD_CHAIN_TIMELINE") && p.total_tiles >= 64ll * grid;
  if (timeline) {
    if (!dbg_dev) cudaMalloc(&dbg_dev, 2 * 64 * 8 * sizeof(long long));
    cudaMemsetAsync(dbg_dev, 0, 2 * 64 * 8 * sizeof(long long), s);
    p.dbg = dbg_dev;
  }
#endif
#define UYD_CHAIN_LAUNCH(A, B, D) UYD_CUDA(launch_pdl(conv_chain_kernel<A, B, D>, dim3(grid), dim3(kChainThreads), cc->smem, s, cc->tm_in, cc->tm_w1, cc->tm_w2, p))
  if (cc->dw) {  // cin == n1: ks1 == ks2
    if (ks1 == 2) UYD_CHAIN_LAUNCH(2, 2, true);
    else UYD_CHAIN_LAUNCH(4, 4, true);
  } else if (ks1 == 2 && ks2 == 2) UYD_CHAIN_LAUNCH(2, 2, false);
  else if (ks1 == 2 && ks2 == 4) UYD_CHAIN_LAUNCH(2, 4, false);
  else if (ks1 == 4 && ks2 == 2) UYD_CHAIN_LAUNCH(4, 2, false);
  else UYD_CHAIN_LAUNCH(4, 4, false);
#undef UYD_CHAIN_LAUNCH
#ifdef UYD_CHAIN_TIMELINE_BUILD
  if (timeline) {  // debug only: dump the stamps of CTA 0 (cycles relative to its first TMA issue)
    long long h[2 * 64 * 8];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "chain timeline cin=%d N1=%d N2=%d final=%d (tma g1issue g2issue acc1 a2 acc2 done)\n", p.cb_bytes / 2, p.N1, p.N2, p.final_kind);
    for (int t = 0; t < 40; ++t) {
      fprintf(stderr, "  tile %2d:", t);
      for (int k = 0; k < 7; ++k) fprintf(stderr, " %7lld", h[t * 8 + k] ? h[t * 8 + k] - h[0] : -1);
      fprintf(stderr, "  | mma:");
      for (int k = 0; k < 6; ++k) fprintf(stderr, " %7lld", h[512 + t * 8 + k] ? h[512 + t * 8 + k] - h[0] : -1);
      fprintf(stderr, "\n");
    }
  }
#endif
  return (int)cudaGetLastError();
}

}  // namespace uyd
