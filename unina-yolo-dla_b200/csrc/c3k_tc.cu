// Fused C3k block on tcgen05 (third generation; c3k_flat.cu is the mma.sync one).
//
//     a  = relu(cv1 x)   b2 = relu(cv2 x)                       1x1   c  -> c_ (two convs, one GEMM with N = [a | b2])
//     t1 = relu(m0.cv1 a)          u = a + relu(m0.cv2 t1)      3x3   c_ -> c_
//     t2 = relu(m1.cv1 u)          v = u + relu(m1.cv2 t2)      3x3
//     y  = relu(cv3 [v | b2])                                   1x1   2c_ -> c
//
// How a C <= 16 convolution feeds the 5th-generation tensor core (measured first: tools/probes/umma_noswz.cu):
//   * Every activation lives in shared memory as PLANES of 16-byte units: one unit = 8 bf16 channels of one pixel
//     (c_ = 8: one plane; c_ = 16: two planes; c = 8 blocks stay on the mma.sync kernel).  A CTA owns a
//     strip of TH output rows x the full image width; its frame ((TH + 8) rows x (W + 2) units: 4 halo rows above and
//     below, one zero column left and right) is a FLAT array of units.
//   * 128 consecutive flat units are the 128 rows of an MMA (K-major, NO swizzle: a "core matrix" is 8 consecutive units
//     = 128 contiguous bytes, SBO = 128).  The two 16-byte K chunks of a K = 16 MMA are two different 16-byte units, LBO
//     bytes apart -- so a 3x3 tap is just a flat offset: for c_ = 8 one MMA covers TWO taps (LBO = the distance between
//     the two taps' pixels, 16 bytes for horizontal neighbours: the core matrices overlap), 5 MMAs per conv and M-tile;
//     for c_ = 16 the two chunks are the two channel planes of one tap (LBO = plane size), 9 MMAs.
//   * Units that wrap around a row end compute garbage that is never stored; "outside the image -> 0" (each conv
//     zero-pads ITS input) is applied when an epilogue stores a unit.
//   * Six stages run as six software pipelines side by side: warp s (s = 0..5) issues the MMAs of stage s, M-tile after
//     M-tile, into kNB TMEM accumulator buffers of its own, and epilogue group s (four warps) drains them (tcgen05.ld ->
//     bias, ReLU, residual, mask -> one 16-byte store per unit) and completes the tile's "done" mbarrier, which is what the
//     issuer of stage s + 1 waits for.  What was measured on the way (tools/c3k_timeline.py, tools/probes/mma_rate.cu):
//       - one thread issuing all stages in wavefront order spends ~100 instructions (~600 cycles) per item: 2.5x slower
//         than the mma.sync kernel; six issuing warps saturate the tensor pipe (39 cycles per N = 16 MMA);
//       - a warp that POLLS several barriers must use mbarrier.test_wait: try_wait parks it for the hardware's time limit
//         (~16 k cycles) on the first barrier that is not ready;
//       - issuing under `if (lane == 0)` costs twice the cycles of elect.sync; tcgen05.commit per item is free;
//       - an N = 16, K = 16 MMA reads 4.5 KB of shared memory for 32 k MACs: at 39 cycles it runs at the shared-memory
//         bandwidth (115 of 128 B/clk), which is the bound of this kernel -- epilogue loads / stores slow down 2x while
//         the MMAs of five other stages stream, and accumulator depth (kNB = 2 or 4) makes no difference.
//   * x arrives by TMA in boxes of kXRows frame rows per 8-channel plane ([channel block][row][column][8 channels],
//     zero-filled outside the image), one mbarrier per kXGroup rows, so stage 0 starts when the first rows have landed; u / v
//     overwrite a in place, t2 overwrites t1, and t1 / b2 reuse the x planes (see the dependency rule at issue_stage).
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace uyd {

struct C3kArgs {  // same struct as in c3k_fused.cu / c3k_flat.cu / api.cu
  const __nv_bfloat16 *in;
  __nv_bfloat16 *out;
  const uint32_t *wfrag;
  const float *bias;
  int n, h, w, in_pitch, out_pitch, th, tiles_x, tiles_y;
};

struct C3kTcParams {
  int H, W, n, TH, strips, pitch, frpx;
  int nt[6], lo[6], hi[6];
  uint32_t plane_bytes;   // one activation plane incl. slack
  uint32_t w_bytes;
  const __nv_bfloat16 *w;
  const float *bias;      // [7][32]
  __nv_bfloat16 *out;
  int out_pitch;
  unsigned m_pitch, m_strips;
  long long *dbg;         // UYD_C3K_TIMELINE_BUILD only: clock64 stamps of CTA 0
};
#ifdef UYD_C3K_TIMELINE_BUILD
#define C3K_STAMP(i) do { if (p.dbg && blockIdx.x == 0) p.dbg[(i)] = clock64(); } while (0)
#else
#define C3K_STAMP(i) do { } while (0)
#endif

namespace {

constexpr int kIssuers = 6;    // one MMA-issuing warp per stage
constexpr int kGroups = 6;     // epilogue groups of four warps, one per stage (the epilogues, not the MMAs, set the pace: measured)
constexpr int kEpiWarp0 = 8;   // first epilogue warp (a multiple of 4: warp % 4 selects the TMEM lane quarter)
constexpr int kThreads = 32 * kEpiWarp0 + 128 * kGroups;
constexpr int kMaxTiles = 32;  // M-tiles per stage (done barriers)
constexpr int kNB = 2;         // accumulator buffers per stage: M-tiles of one stage in flight
constexpr int kTmemCols = 512; // one CTA per SM
constexpr int kXRows = 4;      // frame rows per TMA box of x (strip heights are multiples of 4)
constexpr int kXGroup = 8;     // frame rows per "x has landed" barrier
constexpr int kMaxXBars = 8;

__device__ __forceinline__ uint64_t desc_ns(uint32_t addr, uint32_t lbo, uint32_t sbo) {  // K-major, no swizzle
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// 8 fp32 accumulators + bias -> ReLU -> (+ residual unit) -> bf16 unit (16 bytes); zero when !keep
__device__ __forceinline__ uint4 finish_unit(const uint32_t (&acc)[8], const float *bias, const uint4 *res, bool keep) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float lo = __uint_as_float(acc[2 * i]) + bias[2 * i], hi = __uint_as_float(acc[2 * i + 1]) + bias[2 * i + 1];
    if (res) {
      const uint32_t r = reinterpret_cast<const uint32_t *>(res)[i];
      __nv_bfloat162 h = __floats2bfloat162_rn(fmaxf(lo, 0.f) + __uint_as_float(r << 16), fmaxf(hi, 0.f) + __uint_as_float(r & 0xffff0000u));
      w[i] = *reinterpret_cast<uint32_t *>(&h);
    } else {
      w[i] = relu_pack_bf16x2(lo, hi);
    }
  }
  return keep ? make_uint4(w[0], w[1], w[2], w[3]) : make_uint4(0u, 0u, 0u, 0u);
}

// TMEM columns of (stage, buffer): kNB accumulator buffers per stage.  c_ = 8: six stages x kNB x 16 columns; c_ = 16:
// stages 0 / 5 (N = 32) take kNB x 32 columns each, the 3x3 stages kNB x 16 (512 columns in all).
template <int PA>
__device__ __forceinline__ uint32_t tmem_col(int s, int b) {
  return PA == 1 ? (uint32_t)(16 * (kNB * s + b))
                 : (s == 0 ? (uint32_t)(32 * b) : s == 5 ? (uint32_t)(32 * kNB + 32 * b) : (uint32_t)(64 * kNB + 16 * (kNB * (s - 1) + b)));
}

struct C3kSmem {  // shared-memory addresses (shared window) every role needs
  uint32_t x_s, a_s, t_s, b_s, w_s, xp_bytes;
  uint32_t xfull, acc_full, acc_empty, done0;  // barriers: acc_full / acc_empty [6][kNB], done [6][kMaxTiles]
};

// ---- MMA issuer of stage S: one warp, one M-tile after the other ----------------------------------------------------
// Dependency rule: tile j of stage S reads units up to its last unit + pitch + 1 of stage S - 1's output, so it waits for
// the "done" barrier of the stage S - 1 tile holding that unit (one epilogue group drains a stage front to back, so the
// done barriers of a stage complete in order).  The same wait makes the in-place updates safe: whatever tile j's epilogue
// overwrites (u / v over a, t2 over t1, t1 and b2 over x) was last read by MMAs of exactly those earlier tiles.
template <int PA, int S>
__device__ __forceinline__ void issue_stage(const C3kTcParams &p, const C3kSmem &sm, uint32_t tmem_base, int lane) {
  constexpr int N0 = 16 * PA, M3 = PA == 1 ? 5 : 9;
  constexpr int NM = (S == 0 || S == 5) ? PA : M3;
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(((S == 0 || S == 5) ? N0 : 16) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint32_t wb0 = 2u * N0 * 16u, wb3 = 2u * 16u * 16u;
  // weight blocks in shared memory: [S0: PA MMAs x (2 x N0 x 16 B)] [4 convs x M3 MMAs x (2 x 16 x 16 B)] [S5: PA MMAs x (2 x N0 x 16 B)]
  const uint32_t w3_s = sm.w_s + (uint32_t)PA * wb0, w5_s = w3_s + 4u * M3 * wb3;
  // descriptors of this stage for unit 0 of its source (an M-tile adds its first unit to the 16-byte-granular address field)
  uint64_t ad[NM], bd[NM];
  const int tap_off[10] = {-p.pitch - 1, -p.pitch, -p.pitch + 1, -1, 0, 1, p.pitch - 1, p.pitch, p.pitch + 1, p.pitch + 2};
#pragma unroll
  for (int m = 0; m < NM; ++m) {
    if (S == 0) {
      ad[m] = desc_ns(sm.x_s + (uint32_t)(2 * m) * sm.xp_bytes, sm.xp_bytes, 128);
      bd[m] = desc_ns(sm.w_s + (uint32_t)m * wb0, N0 * 16u, 128);
    } else if (S == 5) {
      // c_ = 8: K chunk 0 = b2 (lower address), chunk 1 = v; c_ = 16: MMA 0 = the v planes, MMA 1 = the b2 planes
      if (PA == 1) ad[m] = desc_ns(sm.b_s, sm.a_s - sm.b_s, 128);
      else ad[m] = m == 0 ? desc_ns(sm.a_s, p.plane_bytes, 128) : desc_ns(sm.b_s, sm.xp_bytes, 128);
      bd[m] = desc_ns(w5_s + (uint32_t)m * wb0, N0 * 16u, 128);
    } else {
      constexpr bool from_a = S == 1 || S == 3;
      const uint32_t src = from_a ? sm.a_s : sm.t_s, pstride = from_a ? p.plane_bytes : sm.xp_bytes;
      if (PA == 1) ad[m] = desc_ns((uint32_t)((int)src + tap_off[2 * m] * 16), (uint32_t)(tap_off[2 * m + 1] - tap_off[2 * m]) * 16u, 128);
      else ad[m] = desc_ns((uint32_t)((int)src + tap_off[m] * 16), pstride, 128);
      bd[m] = desc_ns(w3_s + (uint32_t)((S - 1) * M3 + m) * wb3, 256, 128);
    }
  }
  constexpr int SP = S > 0 ? S - 1 : 0;
  const int reach = (S == 5) ? 127 : 127 + p.pitch + 1;  // last unit read, relative to the tile start
  const int nt = p.nt[S], lo = p.lo[S], ntp = p.nt[SP], lop = p.lo[SP];
  int waited = -1;
  for (int j = 0; j < nt; ++j) {
    if (S == 0) {  // x arrives in groups of kXGroup frame rows
      int last = lo + 128 * j + 127;
      if (last > p.frpx - 1) last = p.frpx - 1;
      const int g = (int)fastdiv((unsigned)last, (unsigned)p.pitch, p.m_pitch) / kXGroup;
      for (; waited < g; ++waited) mbar_wait(sm.xfull + 8u * (uint32_t)(waited + 1), 0);
    }
    if (S > 0) {
      int jdep = (lo + 128 * j + reach - lop) >> 7;
      if (jdep > ntp - 1) jdep = ntp - 1;
      if (jdep > waited) {
        mbar_wait(sm.done0 + 8u * (uint32_t)(SP * kMaxTiles + jdep), 0);
        waited = jdep;
      }
    }
    const int b = j % kNB;
    mbar_wait(sm.acc_empty + 8u * (uint32_t)(kNB * S + b), (((uint32_t)(j / kNB)) & 1u) ^ 1u);
    tc_fence_after();
    if (elect_one()) {  // (issuing under `lane == 0` costs twice the cycles per tcgen05.mma: tools/probes/mma_rate.cu)
      C3K_STAMP(8 + 8 * (S * kMaxTiles + j) + 1);
      const uint32_t d = tmem_base + tmem_col<PA>(S, b);
      const uint64_t f = (uint64_t)(uint32_t)(lo + 128 * j);  // units of 16 bytes = the descriptor's address granularity
#pragma unroll
      for (int m = 0; m < NM; ++m) umma_bf16(d, ad[m] + f, bd[m], idesc, m != 0);
      umma_commit(sm.acc_full + 8u * (uint32_t)(kNB * S + b));
    }
    __syncwarp();
  }
}

// ---- epilogue of one (stage, M-tile): this warp's 32 rows -------------------------------------------------------------
template <int PA, int S>
__device__ __forceinline__ void drain_tile(const C3kTcParams &p, const C3kSmem &sm, unsigned char *base, uint32_t s0, const float *bias_s,
                                           uint32_t lane_base, int j, int m, int lane, int r0, unsigned img) {
  constexpr int NLD = (S == 0 || S == 5) ? 2 * PA : PA;  // 8-column loads
  const int b = j % kNB;
  const int f = p.lo[S] + 128 * j + m;
  const bool valid = f < p.hi[S];
  const unsigned fr = fastdiv((unsigned)f, (unsigned)p.pitch, p.m_pitch), fc = (unsigned)f - fr * (unsigned)p.pitch;
  const int iy = r0 - 4 + (int)fr, ix = (int)fc - 1;
  const bool inside = valid && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
  const uint32_t ta = lane_base + tmem_col<PA>(S, b);
  uint32_t acc[NLD][8];
#pragma unroll
  for (int i = 0; i < NLD; ++i) tmem_ld8(ta + 8u * i, acc[i]);
  tmem_ld_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(sm.acc_empty + 8u * (uint32_t)(kNB * S + b));
  unsigned char *ap = base + (sm.a_s - s0) + (size_t)f * 16, *tp = base + (sm.t_s - s0) + (size_t)f * 16, *bp = base + (sm.b_s - s0) + (size_t)f * 16;
  if (S == 0) {
    if (valid) {
#pragma unroll
      for (int pl = 0; pl < PA; ++pl) {
        *reinterpret_cast<uint4 *>(ap + pl * p.plane_bytes) = finish_unit(acc[pl], bias_s + 8 * pl, nullptr, inside);
        *reinterpret_cast<uint4 *>(bp + pl * sm.xp_bytes) = finish_unit(acc[PA + pl], bias_s + 32 + 8 * pl, nullptr, true);
      }
    }
  } else if (S == 5) {
    if (inside && fr >= 4u && (int)fr < p.TH + 4) {
      __nv_bfloat16 *op = p.out + (((long long)img * p.H + iy) * p.W + ix) * p.out_pitch;
#pragma unroll
      for (int i = 0; i < NLD; ++i) *reinterpret_cast<uint4 *>(op + 8 * i) = finish_unit(acc[i], bias_s + 6 * 32 + 8 * i, nullptr, true);
    }
  } else if (valid) {
    const float *bs = bias_s + (S + 1) * 32;
    if (S == 1 || S == 3) {
#pragma unroll
      for (int pl = 0; pl < PA; ++pl) *reinterpret_cast<uint4 *>(tp + pl * sm.xp_bytes) = finish_unit(acc[pl], bs + 8 * pl, nullptr, inside);
    } else {
#pragma unroll
      for (int pl = 0; pl < PA; ++pl) {
        uint4 *dst = reinterpret_cast<uint4 *>(ap + pl * p.plane_bytes);
        const uint4 rv = *dst;
        *dst = finish_unit(acc[pl], bs + 8 * pl, &rv, inside);
      }
    }
  }
  if (S < 5) {
    fence_proxy_async_smem();  // the next stage's MMAs read these units through the async proxy
    __syncwarp();
    if (lane == 0) mbar_arrive(sm.done0 + 8u * (uint32_t)(S * kMaxTiles + j));
  }
}

// Epilogue warp of stage S: its M-tiles front to back.
template <int PA, int S>
__device__ __forceinline__ void drain_stage(const C3kTcParams &p, const C3kSmem &sm, unsigned char *base, uint32_t s0, const float *bias_s,
                                            uint32_t lane_base, int m, int lane, int r0, unsigned img) {
  const int nt = p.nt[S];
  for (int j = 0; j < nt; ++j) {
    mbar_wait(sm.acc_full + 8u * (uint32_t)(kNB * S + j % kNB), ((uint32_t)(j / kNB)) & 1u);
    tc_fence_after();
    if (m == 0) C3K_STAMP(8 + 8 * (S * kMaxTiles + j) + 2);
    drain_tile<PA, S>(p, sm, base, s0, bias_s, lane_base, j, m, lane, r0, img);
    if (m == 0) C3K_STAMP(8 + 8 * (S * kMaxTiles + j) + 3);
  }
}

// PA = planes per activation: 1 (c_ = 8) or 2 (c_ = 16).  x has 2 PA planes, stages 0 and 5 have N = 16 PA.
template <int PA>
__global__ void __launch_bounds__(kThreads, 1) c3k_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const C3kTcParams p) {
  constexpr int PX = 2 * PA;
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t s0 = (smem_u32(smem) + 127u) & ~127u;
  unsigned char *base = smem + (s0 - smem_u32(smem));
  // x: PX dense planes, as the TMA box writes them.  t1 / t2 ALIAS x planes [0, PA) and b2 aliases x planes [PA, 2 PA):
  // stage 0 reads exactly the units [128 j, 128 j + 128) of every x plane for its M-tile j, its own epilogue writes b2
  // there afterwards, and stage 1 only stores t1 units whose stage-0 tiles have completed (dependency rule above).
  // a / u / v live in their own planes behind x (which also gives every over-read finite data).
  C3kSmem sm;
  sm.xp_bytes = (uint32_t)p.frpx * 16u;  // one x plane
  const uint32_t x_bytes = (uint32_t)PX * sm.xp_bytes;
  sm.x_s = s0;
  sm.t_s = sm.x_s;
  sm.b_s = sm.x_s + (uint32_t)PA * sm.xp_bytes;
  sm.a_s = sm.x_s + ((x_bytes + 127u) & ~127u);
  sm.w_s = sm.a_s + (uint32_t)PA * p.plane_bytes;
  const uint32_t bar0 = sm.w_s + ((p.w_bytes + 127u) & ~127u);
  sm.xfull = bar0;
  sm.acc_full = bar0 + 8 * kMaxXBars;
  sm.acc_empty = sm.acc_full + 8 * 6 * kNB;
  sm.done0 = sm.acc_empty + 8 * 6 * kNB;
  const uint32_t slot_a = sm.done0 + 8 * 6 * kMaxTiles;
  uint32_t *slot_ptr = reinterpret_cast<uint32_t *>(base + (slot_a - s0));
  float *bias_s = reinterpret_cast<float *>(base + (slot_a + 16 - s0));  // [7][32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned cta = blockIdx.x, img = fastdiv(cta, (unsigned)p.strips, p.m_strips), strip = cta - img * (unsigned)p.strips;
  const int r0 = (int)strip * p.TH;  // first output row of this strip

  if (threadIdx.x < kMaxXBars + 12 * kNB + 6 * kMaxTiles) {  // one barrier per thread
    const int i = threadIdx.x - kMaxXBars;
    if (i < 0) mbar_init(sm.xfull + 8u * threadIdx.x, 1);
    else if (i < 6 * kNB) mbar_init(sm.acc_full + 8u * i, 1);
    else if (i < 12 * kNB) mbar_init(sm.acc_empty + 8u * (i - 6 * kNB), 4);
    else mbar_init(sm.done0 + 8u * (i - 12 * kNB), 4);
    fence_barrier_init();
  }
  if (threadIdx.x == 0) C3K_STAMP(0);
  griddep_trigger();
  if (warp == 0) {
    // x first, in boxes of kXRows frame rows per plane, so that stage 0 starts when the first rows have landed
    __syncwarp();  // the x barriers were initialised by lanes of this warp
    griddep_wait();
    if (lane == 0) {
      const int rows = p.TH + 8;
      for (int g = 0; g * kXGroup < rows; ++g) {
        const int gr = rows - g * kXGroup < kXGroup ? rows - g * kXGroup : kXGroup;
        mbar_expect_tx(sm.xfull + 8u * (uint32_t)g, (uint32_t)(PX * gr * p.pitch * 16));
        for (int r = 0; r < gr; r += kXRows)
#pragma unroll
          for (int pl = 0; pl < PX; ++pl)
            tma_load_5d(sm.x_s + (uint32_t)pl * sm.xp_bytes + (uint32_t)((g * kXGroup + r) * p.pitch * 16), &tm_x, sm.xfull + 8u * (uint32_t)g, 0, -1,
                        r0 - 4 + g * kXGroup + r, pl, (int)img);
      }
    }
  } else {
    // zero the a planes incl. their slack: units no epilogue stores must read as finite values (x is finite input data)
    const int tid = threadIdx.x - 32, nth = kThreads - 32;
    uint4 *z = reinterpret_cast<uint4 *>(base + (sm.a_s - s0));
    const int nz = (int)((PA * p.plane_bytes) >> 4);
    for (int i = tid; i < nz; i += nth) z[i] = make_uint4(0u, 0u, 0u, 0u);
    const uint4 *wg = reinterpret_cast<const uint4 *>(p.w);
    uint4 *wsm = reinterpret_cast<uint4 *>(base + (sm.w_s - s0));
    for (int i = tid; i < (int)(p.w_bytes >> 4); i += nth) wsm[i] = __ldg(wg + i);
    for (int i = tid; i < 7 * 32; i += nth) bias_s[i] = __ldg(p.bias + i);
  }
  if (warp == 7) tmem_alloc(slot_a, kTmemCols);
  fence_proxy_async_smem();   // the zero fill and the weights are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot_ptr;
  griddep_wait();
  if (threadIdx.x == 0) C3K_STAMP(1);

  if (warp < kIssuers) {
    switch (warp) {
      case 0: issue_stage<PA, 0>(p, sm, tmem_base, lane); break;
      case 1: issue_stage<PA, 1>(p, sm, tmem_base, lane); break;
      case 2: issue_stage<PA, 2>(p, sm, tmem_base, lane); break;
      case 3: issue_stage<PA, 3>(p, sm, tmem_base, lane); break;
      case 4: issue_stage<PA, 4>(p, sm, tmem_base, lane); break;
      default: issue_stage<PA, 5>(p, sm, tmem_base, lane); break;
    }
  } else if (warp >= kEpiWarp0) {
    const int q = warp & 3, g = (warp - kEpiWarp0) >> 2;
    const int m = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    switch (g) {
      case 0: drain_stage<PA, 0>(p, sm, base, s0, bias_s, lane_base, m, lane, r0, img); break;
      case 1: drain_stage<PA, 1>(p, sm, base, s0, bias_s, lane_base, m, lane, r0, img); break;
      case 2: drain_stage<PA, 2>(p, sm, base, s0, bias_s, lane_base, m, lane, r0, img); break;
      case 3: drain_stage<PA, 3>(p, sm, base, s0, bias_s, lane_base, m, lane, r0, img); break;
      case 4: drain_stage<PA, 4>(p, sm, base, s0, bias_s, lane_base, m, lane, r0, img); break;
      default: drain_stage<PA, 5>(p, sm, base, s0, bias_s, lane_base, m, lane, r0, img); break;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) C3K_STAMP(3);
  if (warp == 7) tmem_dealloc(tmem_base, kTmemCols);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn c3k_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(q);
  }
  return fn;
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------
size_t c3k_tc_weight_words(int c);

// Strip height: frames must fit shared memory next to the weights.  One CTA per SM, so the launch costs
// (waves of CTAs) x (frame rows per CTA + the ~8 rows' worth of pipeline fill and drain): taller strips recompute less
// halo and pay fewer fills, shorter ones make more CTAs.
static int c3k_tc_pick_th(int c, int h, int w, int n, int sms) {
  const int PA = c == 32 ? 2 : 1;
  const int cands[8] = {40, 32, 28, 24, 20, 16, 12, 8};  // multiples of kXRows
  int best = 0;
  long long best_cost = 0;
  for (int i = 0; i < 8; ++i) {
    const int th = cands[i];
    const int pitch = w + 2, frpx = (th + 8) * pitch;
    const size_t plane = (((size_t)(frpx + 128 + pitch + 8) * 16) + 127) & ~(size_t)127;
    const size_t need = 256 + (size_t)2 * PA * frpx * 16 + 128 + PA * plane + ((c3k_tc_weight_words(c) * 4 + 127) & ~(size_t)127) + 4096;
    if (need > 227 * 1024 || (frpx + 127) / 128 > kMaxTiles) continue;
    const long long ctas = (long long)n * ((h + th - 1) / th), waves = (ctas + sms - 1) / sms;
    const long long cost = waves * (th + 16);
    if (!best || cost < best_cost) { best = th; best_cost = cost; }
  }
  return best;
}

bool c3k_tc_supported(int c, int h, int w) {
  static const bool off = [] { const char *v = getenv("UYD_C3K_NO_TC"); return v && *v == '1'; }();
  if (off) return false;
  if (!(c == 16 || c == 32)) return false;
  if (h < 1 || w < 1) return false;
  return c3k_tc_pick_th(c, h, w, 1 << 20, 148) > 0;
}

int c3k_tc_strips(int c, int h, int w, int n) {  // CTAs a launch would have
  const int th = c3k_tc_pick_th(c, h, w, n, current_sm_count());
  return th > 0 ? n * ceil_div(h, th) : 0;
}

// Used only where it was measured faster than the 2-D-tiled mma.sync kernel (tools/c3k_sweep.sh: batch 48 ... 256 at 640^2, +2
// to +4 % on the whole step; batch 32 was a wash): batches whose tall strips (>= 24 rows or half the image: little halo
// recompute, one pipeline fill per many M-tiles) still make about one CTA per SM.
bool c3k_tc_preferred(int c, int h, int w, int n) {
  if (!c3k_tc_supported(c, h, w)) return false;
  const int sms = current_sm_count(), th = c3k_tc_pick_th(c, h, w, n, sms);
  return (th >= 24 || 2 * th >= h) && 4ll * n * ceil_div(h, th) >= 3ll * sms;
}

size_t c3k_tc_weight_words(int c) {
  const int PA = c == 32 ? 2 : 1, N0 = 16 * PA, M3 = PA == 1 ? 5 : 9;
  return ((size_t)PA * 2 * N0 * 8 + (size_t)4 * M3 * 2 * 16 * 8 + (size_t)PA * 2 * N0 * 8) / 2;  // bf16 pairs per word
}

// Packed bf16 weight blocks, each [K chunk 2][n][8 input channels] (K-major, no swizzle: 8 rows x 16 B core matrices).
void c3k_tc_pack(int c, const float *const w[7], std::vector<uint32_t> &out) {
  const int PA = c == 32 ? 2 : 1, CH = 8 * PA, N0 = 16 * PA, M3 = PA == 1 ? 5 : 9;
  std::vector<__nv_bfloat16> o;
  auto put = [&](float v) { o.push_back(__float2bfloat16_rn(v)); };
  // stage 0: N = [cv1 | cv2], K = the c input channels in blocks of 16 (two 8-channel planes per MMA)
  for (int m = 0; m < PA; ++m)
    for (int kc = 0; kc < 2; ++kc)
      for (int n = 0; n < N0; ++n)
        for (int ci = 0; ci < 8; ++ci) {
          const int cin = 16 * m + 8 * kc + ci;
          put(n < CH ? w[0][(size_t)n * c + cin] : w[1][(size_t)(n - CH) * c + cin]);
        }
  // 3x3 convs: c_ = 8: MMA m = taps (2m, 2m + 1), tap 9 = zero; c_ = 16: MMA m = tap m, chunks = input channel halves
  for (int i = 2; i < 6; ++i)
    for (int m = 0; m < M3; ++m)
      for (int kc = 0; kc < 2; ++kc)
        for (int n = 0; n < 16; ++n)
          for (int ci = 0; ci < 8; ++ci) {
            float v = 0.f;
            if (PA == 1) {
              const int tap = 2 * m + kc;
              if (tap < 9 && n < 8) v = w[i][((size_t)n * 8 + ci) * 9 + tap];
            } else {
              v = w[i][((size_t)n * 16 + 8 * kc + ci) * 9 + m];
            }
            put(v);
          }
  // stage 5: input channels [v | b2]; c_ = 8: one MMA (chunk 0 = b2, chunk 1 = v: b2 sits at the lower address); c_ = 16: MMA 0 = v planes, MMA 1 = b2 planes
  for (int m = 0; m < PA; ++m)
    for (int kc = 0; kc < 2; ++kc)
      for (int n = 0; n < N0; ++n)
        for (int ci = 0; ci < 8; ++ci) {
          const int cin = PA == 1 ? 8 * (1 - kc) + ci : 16 * m + 8 * kc + ci;  // c_ = 8: chunk 0 = b2 (channels 8..15), chunk 1 = v
          put(w[6][(size_t)n * c + cin]);
        }
  const size_t words = o.size() / 2;
  const uint32_t *src = reinterpret_cast<const uint32_t *>(o.data());
  out.insert(out.end(), src, src + words);
}

int c3k_tc_launch(int c, const C3kArgs &a, const uint32_t *w_tc, cudaStream_t s) {
  const int PA = c == 32 ? 2 : 1, PX = 2 * PA;
  C3kTcParams p;
  memset(&p, 0, sizeof(p));
  const int sms = current_sm_count();
  p.H = a.h; p.W = a.w; p.n = a.n;
  p.TH = c3k_tc_pick_th(c, a.h, a.w, a.n, sms);
  UYD_REQUIRE(p.TH > 0, UYD_E_UNSUPPORTED, "c3k_tc: no strip height fits shared memory for %dx%d", a.h, a.w);
  p.strips = ceil_div(a.h, p.TH);
  p.pitch = a.w + 2;
  p.frpx = (p.TH + 8) * p.pitch;
  p.plane_bytes = (uint32_t)((((size_t)(p.frpx + 128 + p.pitch + 8) * 16) + 127) & ~(size_t)127);
  // stage s works on frame rows [r_s, TH + 8 - r_s): 0, 1, 2, 3, 4, 4
  const int rs[6] = {0, 1, 2, 3, 4, 4};
  for (int st = 0; st < 6; ++st) {
    p.lo[st] = rs[st] * p.pitch;
    p.hi[st] = (p.TH + 8 - rs[st]) * p.pitch;
    p.nt[st] = ceil_div(p.hi[st] - p.lo[st], 128);
    UYD_REQUIRE(p.nt[st] <= kMaxTiles, UYD_E_UNSUPPORTED, "c3k_tc: %d M-tiles exceed the barrier table", p.nt[st]);
  }
  p.w = reinterpret_cast<const __nv_bfloat16 *>(w_tc);
  p.w_bytes = (uint32_t)(c3k_tc_weight_words(c) * 4);
  p.bias = a.bias;
  p.out = a.out; p.out_pitch = a.out_pitch;
  p.m_pitch = fastdiv_magic((unsigned)p.pitch);
  p.m_strips = fastdiv_magic((unsigned)p.strips);
  UYD_REQUIRE(a.in_pitch % 8 == 0 && a.out_pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(a.in) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_tc) & 15) == 0 && p.w_bytes % 16 == 0,
              UYD_E_UNSUPPORTED, "c3k_tc: 16-byte aligned slices and weights");
  // x as [channel block][row][column][8 channels]: one 5-D box, zero-filled outside the image
  static std::mutex mu;
  static std::map<std::tuple<const void *, int, int, int, int, int, int>, CUtensorMap> cache;
  CUtensorMap tm;
  {
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_tuple((const void *)a.in, a.in_pitch, a.h, a.w, a.n, c, p.TH);
    auto it = cache.find(key);
    if (it == cache.end()) {
      EncodeTiledFn fn = c3k_encode_fn();
      UYD_REQUIRE(fn, UYD_E_NOGPU, "cuTensorMapEncodeTiled is not available (no CUDA driver)");
      const cuuint64_t dims[5] = {8, (cuuint64_t)a.w, (cuuint64_t)a.h, (cuuint64_t)PX, (cuuint64_t)a.n};
      const cuuint64_t str[4] = {(cuuint64_t)a.in_pitch * 2, (cuuint64_t)a.w * a.in_pitch * 2, 16, (cuuint64_t)a.h * a.w * a.in_pitch * 2};
      const cuuint32_t box[5] = {8, (cuuint32_t)p.pitch, (cuuint32_t)kXRows, 1, 1};
      const cuuint32_t one[5] = {1, 1, 1, 1, 1};
      CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<__nv_bfloat16 *>(a.in), dims, str, box, one,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      UYD_REQUIRE(r == CUDA_SUCCESS, UYD_E_ARG, "c3k_tc: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
      if (cache.size() > 4096) cache.clear();
      cache[key] = tm;
    } else {
      tm = it->second;
    }
  }
  const size_t smem = 256 + (((size_t)PX * p.frpx * 16 + 127) & ~(size_t)127) + (size_t)PA * p.plane_bytes + ((p.w_bytes + 127) & ~127u) +
                      8 * kMaxXBars + 8 * 12 * kNB + 8 * 6 * kMaxTiles + 32 + 7 * 32 * 4 + 128;
  UYD_REQUIRE(smem <= 227 * 1024, UYD_E_UNSUPPORTED, "c3k_tc: %zu bytes of shared memory", smem);
  const unsigned grid = (unsigned)(a.n * p.strips);
  p.dbg = nullptr;
#ifdef UYD_C3K_TIMELINE_BUILD  // debug build only (tools/c3k_timeline.py): never in the shipped library path
  static long long *dbg_dev = nullptr;
  const bool timeline = getenv("UYD_C3K_TIMELINE") && grid >= 64;
  if (timeline) {
    if (!dbg_dev) cudaMalloc(&dbg_dev, (8 + 8 * 6 * kMaxTiles) * sizeof(long long));
    cudaMemsetAsync(dbg_dev, 0, (8 + 8 * 6 * kMaxTiles) * sizeof(long long), s);
    p.dbg = dbg_dev;
  }
#endif
  if (PA == 1) {
    if (int e = smem_optin(c3k_tc_kernel<1>, 227 * 1024)) return e;
    UYD_CUDA(launch_pdl(c3k_tc_kernel<1>, dim3(grid), dim3(kThreads), smem, s, tm, p));
  } else {
    if (int e = smem_optin(c3k_tc_kernel<2>, 227 * 1024)) return e;
    UYD_CUDA(launch_pdl(c3k_tc_kernel<2>, dim3(grid), dim3(kThreads), smem, s, tm, p));
  }
#ifdef UYD_C3K_TIMELINE_BUILD
  if (timeline) {
    static long long h[8 + 8 * 6 * kMaxTiles];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "c3k_tc timeline c=%d %dx%d TH=%d: sync %lld end %lld\n", c, a.h, a.w, p.TH, h[1] - h[0], h[3] - h[0]);
    for (int st = 0; st < 6; ++st)
      for (int j = 0; j < p.nt[st]; ++j) {
        const long long *e = h + 8 + 8 * (st * kMaxTiles + j);
        fprintf(stderr, "  s %d j %2d: issue %6lld | acc_full %6lld done %6lld\n", st, j, e[1] - h[0], e[2] - h[0], e[3] - h[0]);
      }
  }
#endif
  return (int)cudaGetLastError();
}

}  // namespace uyd
