// Class-aware NMS, bit-exact with the CPU statement.
//
// Pipeline (Ultralytics non_max_suppression + torchvision.ops.nms, SURVEY.md A.3/A.4;
// reference call sites train.py:396-405, eval.py:32):
//   1. nms_key_kernel     : per anchor best class / score, conf filter, 64-bit sort key
//                           (image | descending-score code | anchor) -> ties resolve to the lower anchor
//   2. cub radix sort     : one sort over the whole batch; image segments come out contiguous,
//                           each in stable score-descending order.  Keys are produced in anchor order
//                           and the LSD radix sort is stable, so only the (image, score) bits are sorted:
//                           5 passes instead of 8 when conf_thr >= 0 (31-bit code of a positive score)
//   3. nms_gather_kernel  : top max_nms candidates per image -> xyxy, class-offset boxes
//   4. nms_greedy_kernel  : one CTA per image.  Candidates are consumed in chunks of 512:
//                           (a) every candidate is tested against the boxes kept so far,
//                           (b) survivors are compacted with warp ballots,
//                           (c) a 512x512 IoU bitmask among survivors is built in shared memory,
//                           (d) one warp resolves the chunk serially (suppression words are
//                               OR-ed across lanes), appending to the kept list.
//                           Stops after max_det boxes -- greedy kept order is score order, so the
//                           first max_det kept boxes never depend on later candidates.
//   5. emit kernel        : rows (x1,y1,x2,y2,conf,cls) / detection records in kept order.
// All IoU arithmetic uses explicit round-to-nearest fp32 intrinsics in the operand order of
// the CPU code it mirrors.  Two IoU policies:
//   TV  : torchvision nms_kernel_impl on class-offset boxes; the fp32 IoU is compared against
//         the largest float <= the double threshold (== torchvision's float-vs-double compare).
//   HPP : the reference's own postprocess.hpp:28-67 (same class only, early-out on empty
//         intersection, float threshold).
#include <cub/cub.cuh>

#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace uyd {
namespace {

constexpr int kChunk = 128;
constexpr int kGreedyThreads = 512;
constexpr int kMaxDetCap = 1024;
constexpr int kAnchorBits = 22;
constexpr int kScoreShift = kAnchorBits;
constexpr int kImageShift = kAnchorBits + 32;
constexpr uint64_t kInvalidKey = ~0ull;

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct Carver {
  char *base;
  size_t off = 0;
  explicit Carver(void *ws) : base((char *)ws) {}
  template <class T>
  T *take(size_t count) {
    T *p = base ? (T *)(base + off) : nullptr;
    off += align_up(count * sizeof(T));
    return p;
  }
};

struct Layout {  // carve-up of the caller's workspace for uyd_nms
  uint64_t *keys_in, *keys_out;
  int *count, *offset;  // [batch]
  float4 *box, *boxoff; // [batch][cap]
  float *conf;
  int *cls, *anchor;
  int *kept_rank;       // [batch][kMaxDetCap]
  void *cub_tmp;
  size_t cub_bytes, total;
};

Layout carve(void *ws, int batch, int anchors) {
  Layout L;
  Carver c(ws);
  const size_t n = (size_t)batch * anchors;
  L.keys_in = c.take<uint64_t>(n);
  L.keys_out = c.take<uint64_t>(n);
  L.count = c.take<int>(batch);
  L.offset = c.take<int>(batch);
  L.box = c.take<float4>(n);
  L.boxoff = c.take<float4>(n);
  L.conf = c.take<float>(n);
  L.cls = c.take<int>(n);
  L.anchor = c.take<int>(n);
  L.kept_rank = c.take<int>((size_t)batch * kMaxDetCap);
  L.cub_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, L.cub_bytes, (uint64_t *)nullptr, (uint64_t *)nullptr, (int)n, 0, 64);
  L.cub_tmp = c.take<char>(L.cub_bytes);
  L.total = c.off;
  return L;
}

// y: [batch, 4+nc, A].  One thread per (image, anchor).
// POS: every candidate score is a positive float (conf_thr >= 0): code = 0x7FFFFFFF - bits (31 bits, image at
// bit 53); otherwise code = order-preserving map of the signed float, complemented (32 bits, image at bit 54).
template <bool POS>
__global__ void __launch_bounds__(256) nms_key_kernel(const float *__restrict__ y, int batch, int nc, int A, float thr,
                                                      uint64_t *__restrict__ keys, int *__restrict__ count) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)batch * A;
  bool ok = false;
  int b = 0;
  if (t < total) {
    b = (int)(t / A);
    const int a = (int)(t % A);
    const float *p = y + ((long long)b * (4 + nc) + 4) * A + a;
    float best = p[0];
    for (int c = 1; c < nc; ++c) {
      const float v = p[(long long)c * A];
      if (v > best) best = v;  // first maximum wins
    }
    ok = best > thr;
    uint64_t key = kInvalidKey;
    if (ok) {
      const uint32_t bits = __float_as_uint(best);
      if (POS) {
        key = ((uint64_t)b << (kImageShift - 1)) | ((uint64_t)(0x7FFFFFFFu - bits) << kScoreShift) | (uint64_t)a;
      } else {  // ascending-order code of a signed float, complemented for descending order
        const uint32_t asc = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
        key = ((uint64_t)b << kImageShift) | ((uint64_t)(~asc) << kScoreShift) | (uint64_t)a;
      }
    }
    keys[t] = key;
  }
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if (m) {  // per-image counts: one atomic per warp (plus stragglers across an image boundary)
    const int lane = threadIdx.x & 31;
    const int b0 = __shfl_sync(0xffffffffu, b, __ffs(m) - 1);
    const unsigned same = __ballot_sync(0xffffffffu, ok && b == b0);
    if (lane == __ffs(same) - 1) atomicAdd(&count[b0], __popc(same));
    if (ok && b != b0) atomicAdd(&count[b], 1);
  }
}

__global__ void nms_scan_kernel(const int *count, int *offset, int batch) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < batch; ++b) {
      offset[b] = acc;
      acc += count[b];
    }
  }
}

__global__ void __launch_bounds__(256) nms_gather_kernel(const float *__restrict__ y, int nc, int A, int max_nms,
                                                         float max_wh, const uint64_t *__restrict__ keys, Layout L) {
  const int b = blockIdx.y;
  const int r = blockIdx.x * 256 + threadIdx.x;
  const int n = min(L.count[b], max_nms);
  if (r >= n) return;
  const uint64_t key = keys[(long long)L.offset[b] + r];
  const int a = (int)(key & ((1ull << kAnchorBits) - 1));
  const float *p = y + (long long)b * (4 + nc) * A + a;
  const float cx = p[0], cy = p[A], w = p[2ll * A], h = p[3ll * A];
  const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);  // == w / 2 exactly
  float4 bx;
  bx.x = __fsub_rn(cx, hw); bx.y = __fsub_rn(cy, hh); bx.z = __fadd_rn(cx, hw); bx.w = __fadd_rn(cy, hh);
  float best = p[4ll * A];
  int j = 0;
  for (int c = 1; c < nc; ++c) {
    const float v = p[(long long)(4 + c) * A];
    if (v > best) { best = v; j = c; }
  }
  const float off = __fmul_rn((float)j, max_wh);
  float4 bo;
  bo.x = __fadd_rn(bx.x, off); bo.y = __fadd_rn(bx.y, off); bo.z = __fadd_rn(bx.z, off); bo.w = __fadd_rn(bx.w, off);
  const long long o = (long long)b * A + r;
  L.box[o] = bx; L.boxoff[o] = bo; L.conf[o] = best; L.cls[o] = j; L.anchor[o] = a;
}

__device__ __forceinline__ float box_area(const float4 &b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

template <bool HPP>
__device__ __forceinline__ bool suppresses(const float4 &a, float aa, int ca, const float4 &b, float ab, int cb, float thr) {
  if (HPP) {  // postprocess.hpp:28-39,58-62
    if (ca != cb) return false;
    const float ix1 = fmaxf(a.x, b.x), iy1 = fmaxf(a.y, b.y), ix2 = fminf(a.z, b.z), iy2 = fminf(a.w, b.w);
    if (ix1 >= ix2 || iy1 >= iy2) return 0.0f > thr;
    const float inter = __fmul_rn(__fsub_rn(ix2, ix1), __fsub_rn(iy2, iy1));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter)) > thr;
  }
  // torchvision nms_kernel_impl operand order
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y), xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  // inter == 0 gives 0/x = 0 or 0/0 = NaN: never > thr for thr >= 0, so the division is skipped
  if (inter == 0.f && thr >= 0.f) return false;
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter)) > thr;
}

// boxes/cls: candidates of image b at [b * cand_stride, ...), in processing order.
// n_ptr[b] candidates (clamped to n_cap).  Writes kept ranks (<= max_det) and their number.
//
// The kernel is instruction-bound (ncu: IPC ~4 on the 64 SMs that hold a CTA), so the work per
// candidate is kept minimal: chunks of 128 candidates (the within-chunk bitmask costs chunk^2/2
// IoUs, so small chunks are cheap), 4 threads per candidate splitting the kept list in phase (a).
template <bool HPP>
__global__ void __launch_bounds__(kGreedyThreads) nms_greedy_kernel(const float4 *__restrict__ boxes,
                                                                    const int *__restrict__ cls, long long cand_stride,
                                                                    const int *__restrict__ n_ptr, int n_cap, int max_det,
                                                                    float thr, int *__restrict__ kept_rank,
                                                                    int *__restrict__ out_count) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  float4 *kbox = reinterpret_cast<float4 *>(nms_smem);      // [kMaxDetCap] kept boxes
  float4 *abox = kbox + kMaxDetCap;                         // [kChunk] survivors of this chunk
  float *karea = reinterpret_cast<float *>(abox + kChunk);  // [kMaxDetCap]
  float *aarea = karea + kMaxDetCap;                        // [kChunk]
  int *kcls = reinterpret_cast<int *>(aarea + kChunk);      // [kMaxDetCap]
  int *acls = kcls + kMaxDetCap;                            // [kChunk]
  int *ksrc = acls + kChunk;                                // [kMaxDetCap] rank of every kept box
  int *asrc = ksrc + kMaxDetCap;                            // [kChunk] rank of each survivor
  unsigned(*mask)[kChunk / 32] = reinterpret_cast<unsigned(*)[kChunk / 32]>(asrc + kChunk);
  __shared__ int warp_cnt[kGreedyThreads / 32];
  __shared__ int s_kept, s_alive;

  const int b = blockIdx.x;
  const int n = min(n_ptr[b], n_cap);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float4 *cbox = boxes + (long long)b * cand_stride;
  const int *ccls = cls + (long long)b * cand_stride;
  if (tid == 0) s_kept = 0;
  __syncthreads();

  constexpr int kSplit = kGreedyThreads / kChunk;  // threads per candidate in phase (a)
  for (int c0 = 0; c0 < n; c0 += kChunk) {
    const int kept = s_kept;
    if (kept >= max_det) break;
    // (a) test against the kept list: kSplit consecutive lanes share a candidate
    const int i = c0 + tid / kSplit, slice = tid % kSplit;
    const bool valid = i < n;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    float ar = 0.f;
    int cl = 0;
    bool sup = false;
    if (valid) {
      bx = cbox[i];
      cl = ccls[i];
      ar = box_area(bx);
      for (int k = slice; k < kept; k += kSplit)
        if (suppresses<HPP>(kbox[k], karea[k], kcls[k], bx, ar, cl, thr)) { sup = true; break; }
    }
#pragma unroll
    for (int o = 1; o < kSplit; o <<= 1) sup |= __shfl_xor_sync(0xffffffffu, sup ? 1 : 0, o) != 0;
    const bool alive = valid && !sup && slice == 0;  // one representative lane per candidate
    // (b) ordered compaction of survivors
    const unsigned bal = __ballot_sync(0xffffffffu, alive);
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    int base = 0;
    for (int w2 = 0; w2 < wid; ++w2) base += warp_cnt[w2];
    if (tid == kGreedyThreads - 1) s_alive = base + __popc(bal);
    if (alive) {
      const int slot = base + __popc(bal & ((1u << lane) - 1));
      abox[slot] = bx; aarea[slot] = ar; acls[slot] = cl; asrc[slot] = i;
    }
    __syncthreads();
    const int m = s_alive;
    const int words = (m + 31) >> 5;
    // (c) suppression bitmask among survivors (upper triangle)
    for (int item = tid; item < m * words; item += kGreedyThreads) {
      const int r = item / words, cw = item % words;
      unsigned bits = 0;
      if (cw * 32 + 31 > r) {
        const float4 rb = abox[r];
        const float ra = aarea[r];
        const int rc = acls[r];
        const int j0 = cw * 32;
#pragma unroll 4
        for (int jj = 0; jj < 32; ++jj) {
          const int j = j0 + jj;
          if (j > r && j < m && suppresses<HPP>(rb, ra, rc, abox[j], aarea[j], acls[j], thr)) bits |= 1u << jj;
        }
      }
      mask[r][cw] = bits;
    }
    __syncthreads();
    // (d) resolution by warp 0: lane l owns suppression word l.  The warp jumps straight to the next
    // unsuppressed survivor (ballot over the lanes' words + ffs): one serial link per KEPT box.
    if (wid == 0) {
      unsigned remv = 0;
      int k = kept;
      int r = 0;
      while (k < max_det) {
        const int base_bit = lane * 32;
        unsigned alive_bits = ~remv;
        if (base_bit + 32 <= r) alive_bits = 0u;
        else if (base_bit < r) alive_bits &= ~0u << (r - base_bit);
        if (base_bit >= m) alive_bits = 0u;
        else if (base_bit + 32 > m) alive_bits &= (1u << (m - base_bit)) - 1u;
        const unsigned lanes = __ballot_sync(0xffffffffu, alive_bits != 0u);
        if (!lanes) break;
        const int src_lane = __ffs(lanes) - 1;
        const int bit = __ffs(__shfl_sync(0xffffffffu, alive_bits, src_lane)) - 1;
        r = src_lane * 32 + bit;
        if (lane == 0) {
          kbox[k] = abox[r]; karea[k] = aarea[r]; kcls[k] = acls[r]; ksrc[k] = asrc[r];
        }
        ++k;
        if (lane < words) remv |= mask[r][lane];
        ++r;
      }
      if (lane == 0) s_kept = k;
    }
    __syncthreads();
  }
  const int kept = s_kept;
  for (int k = tid; k < kept; k += kGreedyThreads) kept_rank[(long long)b * kMaxDetCap + k] = ksrc[k];
  if (tid == 0) out_count[b] = kept;
}

constexpr size_t kGreedySmem = (size_t)(kMaxDetCap + kChunk) * (16 + 4 + 4 + 4) + (size_t)kChunk * (kChunk / 32) * 4;

template <bool HPP>
int launch_greedy(const float4 *boxes, const int *cls, long long cand_stride, const int *n_ptr, int n_cap, int max_det,
                  float thr, int *kept_rank, int *out_count, int batch, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    UYD_CUDA(cudaFuncSetAttribute(nms_greedy_kernel<HPP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGreedySmem));
    attr_set = true;
  }
  nms_greedy_kernel<HPP><<<batch, kGreedyThreads, kGreedySmem, s>>>(boxes, cls, cand_stride, n_ptr, n_cap, max_det, thr,
                                                                   kept_rank, out_count);
  return (int)cudaGetLastError();
}

__global__ void nms_emit_rows_kernel(Layout L, int A, int max_det, const int *__restrict__ out_count,
                                     float *__restrict__ out_det, int *__restrict__ out_idx) {
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < out_count[b]; k += blockDim.x) {
    const long long src = (long long)b * A + L.kept_rank[(long long)b * kMaxDetCap + k];
    const float4 ob = L.box[src];
    float *o = out_det + ((long long)b * max_det + k) * 6;
    o[0] = ob.x; o[1] = ob.y; o[2] = ob.z; o[3] = ob.w; o[4] = L.conf[src]; o[5] = (float)L.cls[src];
    if (out_idx) out_idx[(long long)b * max_det + k] = L.anchor[src];
  }
}

// =================================================================================================
// Fused per-image NMS (the default path of uyd_nms): ONE launch, one 1024-thread CTA per image.
//   1. scan   : best class / score per anchor, conf filter, 64-bit key (score code | anchor | class); the keys
//               of the current slab are captured in shared memory.  A slab is a key range holding at most
//               kFCap candidates: normally the whole image (<= kFCap candidates: one pass over the scores);
//               dense scenes are cut into consecutive ranges with a 2048-bin histogram of the key space
//               (three passes per slab; the ranges are processed best-first, so the loop usually ends after
//               the first slab because max_det boxes are kept).
//   2. sort   : bitonic sort of the slab in shared memory (keys are unique: ties in score resolve to the lower
//               anchor, exactly the stable descending sort of the reference).
//   3. greedy : as nms_greedy_kernel (chunks of 128, kept-list test by 8 threads per candidate, ballot
//               compaction, IoU bitmask among survivors, one warp resolves), boxes formed on the fly from y.
//   4. emit   : rows / anchor indices / count written by the same CTA.
// Replaces key kernel + 5 radix passes over batch x anchors + gather + greedy + emit (365 -> ~150 us at
// batch 64 x 33 600 anchors) and is what batch-1 latency needs (one dependent launch instead of eleven).
// =================================================================================================
constexpr int kFThreads = 1024, kFCap = 4096, kFChunk = 128, kFSplit = kFThreads / kFChunk, kFBins = 2048;
constexpr int kKeyAnchorShift = 8, kKeyCodeShift = 30;  // key = code << 30 | anchor << 8 | class

__device__ __forceinline__ uint64_t fused_key(float best, int j, int a, bool pos) {
  const uint32_t bits = __float_as_uint(best);
  uint32_t code;
  if (pos) code = 0x7FFFFFFFu - bits;
  else code = ~((bits & 0x80000000u) ? ~bits : (bits | 0x80000000u));
  return ((uint64_t)code << kKeyCodeShift) | ((uint64_t)a << kKeyAnchorShift) | (uint64_t)j;
}
__device__ __forceinline__ float fused_key_score(uint64_t key, bool pos) {
  const uint32_t code = (uint32_t)(key >> kKeyCodeShift);
  if (pos) return __uint_as_float(0x7FFFFFFFu - code);
  const uint32_t asc = ~code;
  return __uint_as_float((asc & 0x80000000u) ? (asc & 0x7FFFFFFFu) : ~asc);
}

struct FusedSmem {
  unsigned long long keys[kFCap];
  unsigned int hist[kFBins];
  float4 kbox[kMaxDetCap], kraw[kMaxDetCap];
  float karea[kMaxDetCap], kconf[kMaxDetCap];
  int kcls[kMaxDetCap], kanchor[kMaxDetCap];
  float4 abox[kFChunk], araw[kFChunk];
  float aarea[kFChunk], aconf[kFChunk];
  int acls[kFChunk], aanchor[kFChunk];
  __align__(16) unsigned mask[kFChunk][kFChunk / 32];
  int klist[kFChunk];
  int warp_cnt[kFThreads / 32];
  unsigned long long kmin, kmax, cut;
  int count, kept, alive;
};

__global__ void __launch_bounds__(kFThreads, 1) nms_image_kernel(const float *__restrict__ y, int nc, int A, float conf_thr,
                                                                 float iou_thr, int max_nms, int max_det, float max_wh, int pos,
                                                                 float *__restrict__ out_det, int *__restrict__ out_idx,
                                                                 int *__restrict__ out_count) {
  extern __shared__ __align__(16) unsigned char fused_raw[];
  FusedSmem &S = *reinterpret_cast<FusedSmem *>(fused_raw);
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float *yb = y + (long long)b * (4 + nc) * A;
  if (tid == 0) S.kept = 0;
  unsigned long long lo = 0ull;   // keys below lo have been consumed
  int processed = 0;
  __syncthreads();

  // mode 0: count keys in [lo, hi), track min / max, capture the first kFCap;  mode 1: histogram over (key - kmin) >> shift
  auto scan = [&](unsigned long long hi, int mode, int shift) {
    constexpr int U = 4;  // anchors per thread and round: U independent loads in flight per class plane
    unsigned long long mn = ~0ull, mx = 0ull;  // per-thread key range, reduced once after the loop
    for (int a0 = 0; a0 < A; a0 += U * kFThreads) {
      float best[U];
      int bj[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int a = a0 + u * kFThreads + tid;
        best[u] = a < A ? yb[4ll * A + a] : -INFINITY;
        bj[u] = 0;
      }
      for (int c = 1; c < nc; ++c) {
        const float *pc = yb + (4ll + c) * A;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int a = a0 + u * kFThreads + tid;
          const float v = a < A ? pc[a] : -INFINITY;
          if (v > best[u]) { best[u] = v; bj[u] = c; }   // first maximum wins
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int a = a0 + u * kFThreads + tid;
        bool in = false;
        unsigned long long key = 0ull;
        if (a < A && best[u] > conf_thr) {
          key = fused_key(best[u], bj[u], a, pos != 0);
          in = key >= lo && key < hi;
        }
        if (mode == 0) {
          const unsigned bal = __ballot_sync(0xffffffffu, in);
          if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&S.count, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (in) {
              const int slot = base + __popc(bal & ((1u << lane) - 1));
              if (slot < kFCap) S.keys[slot] = key;
              mn = key < mn ? key : mn;
              mx = key > mx ? key : mx;
            }
          }
        } else if (in) {
          atomicAdd(&S.hist[(unsigned)((key - S.kmin) >> shift)], 1u);
        }
      }
    }
    if (mode == 0) {  // one pair of atomics per warp
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long omn = __shfl_xor_sync(0xffffffffu, mn, o), omx = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = omn < mn ? omn : mn;
        mx = omx > mx ? omx : mx;
      }
      if (lane == 0 && mn <= mx) { atomicMin(&S.kmin, mn); atomicMax(&S.kmax, mx); }
    }
  };

  while (true) {
    // ---- 1. the next slab: keys in [lo, hi) with at most kFCap members ----
    unsigned long long hi = ~0ull;
    int n_slab = 0;
    while (true) {
      if (tid == 0) { S.count = 0; S.kmin = ~0ull; S.kmax = 0ull; }
      __syncthreads();
      scan(hi, 0, 0);
      __syncthreads();
      n_slab = S.count;
      if (n_slab <= kFCap) break;
      // too many: histogram the occupied key range and cut it where the cumulative count would exceed the capacity
      const unsigned long long kmin = S.kmin, span = S.kmax - S.kmin;
      int shift = 0;
      while ((span >> shift) >= (unsigned long long)kFBins) ++shift;
      for (int i = tid; i < kFBins; i += kFThreads) S.hist[i] = 0u;
      __syncthreads();
      scan(hi, 1, shift);
      __syncthreads();
      if (tid == 0) {
        unsigned cum = 0;
        int nb = 0;
        while (nb < kFBins && cum + S.hist[nb] <= (unsigned)kFCap) cum += S.hist[nb++];
        if (nb == 0) nb = 1;  // a single bin overflows: narrow the range to it (keys are unique, so this terminates)
        S.cut = kmin + ((unsigned long long)nb << shift);
      }
      __syncthreads();
      hi = S.cut;
      __syncthreads();
    }
    if (n_slab == 0) break;
    // ---- 2. bitonic sort of the slab (ascending key = descending score, ties -> lower anchor) ----
    int P = 32;
    while (P < n_slab) P <<= 1;
    for (int i = n_slab + tid; i < P; i += kFThreads) S.keys[i] = ~0ull;
    __syncthreads();
    // One compare-exchange per thread and step: pair p -> elements i = p with a zero inserted at bit log2(j), i | j.
    // For j <= 32 the 32 pairs of a warp live in its own 64 consecutive elements: those steps only need __syncwarp.
    for (int k = 2; k <= P; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int pr = tid; pr < (P >> 1); pr += kFThreads) {
          const int i = ((pr & ~(j - 1)) << 1) | (pr & (j - 1)), q = i | j;
          const unsigned long long x = S.keys[i], z = S.keys[q];
          if ((x > z) == ((i & k) == 0)) { S.keys[i] = z; S.keys[q] = x; }
        }
        if (j > 32 || j == 1 || P > 2 * kFThreads) __syncthreads();
        else __syncwarp();
      }
    }
    // ---- 3. greedy NMS over the slab, in order, at most max_nms candidates per image in total ----
    const int take = min(n_slab, max_nms - processed);
    for (int c0 = 0; c0 < take; c0 += kFChunk) {
      const int kept = S.kept;
      if (kept >= max_det) break;
      const int i = c0 + tid / kFSplit, slice = tid % kFSplit;
      const bool valid = i < take;
      float4 bo = make_float4(0.f, 0.f, 0.f, 0.f), bx = bo;
      float ar = 0.f, sc = 0.f;
      int cl = 0, an = 0;
      bool sup = false;
      if (valid) {
        const unsigned long long key = S.keys[i];
        an = (int)((key >> kKeyAnchorShift) & ((1u << kAnchorBits) - 1));
        cl = (int)(key & 0xFF);
        sc = fused_key_score(key, pos != 0);
        const float *p = yb + an;
        const float cx = p[0], cy = p[A], w = p[2ll * A], h = p[3ll * A];
        const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
        bx.x = __fsub_rn(cx, hw); bx.y = __fsub_rn(cy, hh); bx.z = __fadd_rn(cx, hw); bx.w = __fadd_rn(cy, hh);
        const float off = __fmul_rn((float)cl, max_wh);
        bo.x = __fadd_rn(bx.x, off); bo.y = __fadd_rn(bx.y, off); bo.z = __fadd_rn(bx.z, off); bo.w = __fadd_rn(bx.w, off);
        ar = box_area(bo);
        for (int k = slice; k < kept; k += kFSplit)
          if (suppresses<false>(S.kbox[k], S.karea[k], S.kcls[k], bo, ar, cl, iou_thr)) { sup = true; break; }
      }
#pragma unroll
      for (int o = 1; o < kFSplit; o <<= 1) sup |= __shfl_xor_sync(0xffffffffu, sup ? 1 : 0, o) != 0;
      const bool alive = valid && !sup && slice == 0;
      const unsigned bal = __ballot_sync(0xffffffffu, alive);
      if (lane == 0) S.warp_cnt[wid] = __popc(bal);
      __syncthreads();
      int base = 0;
      for (int w2 = 0; w2 < wid; ++w2) base += S.warp_cnt[w2];
      if (tid == kFThreads - 1) S.alive = base + __popc(bal);
      if (alive) {
        const int slot = base + __popc(bal & ((1u << lane) - 1));
        S.abox[slot] = bo; S.araw[slot] = bx; S.aarea[slot] = ar; S.acls[slot] = cl; S.aconf[slot] = sc; S.aanchor[slot] = an;
      }
      __syncthreads();
      const int m = S.alive;
      const int words = (m + 31) >> 5;
      // (c) suppression bitmask among survivors (upper triangle), one item = 16 pairs so that all 1024 threads work
      for (int item = tid; item < m * words * 2; item += kFThreads) {
        const int r = item / (words * 2), cw = (item >> 1) % words, half = item & 1;
        unsigned bits = 0;
        const int j0 = cw * 32 + half * 16;
        if (j0 + 15 > r) {
          const float4 rb = S.abox[r];
          const float ra = S.aarea[r];
          const int rc = S.acls[r];
#pragma unroll 4
          for (int jj = 0; jj < 16; ++jj) {
            const int j = j0 + jj;
            if (j > r && j < m && suppresses<false>(rb, ra, rc, S.abox[j], S.aarea[j], S.acls[j], iou_thr)) bits |= 1u << jj;
          }
        }
        reinterpret_cast<unsigned short *>(&S.mask[r][cw])[half] = (unsigned short)bits;
      }
      __syncthreads();
      // (d) serial resolution by ONE thread with the removal mask in registers (the chain per kept box is
      // ffs -> one 16-byte row load -> or); the kept-list entries are copied by all threads afterwards
      if (tid == 0) {
        unsigned rem[kFChunk / 32] = {0u, 0u, 0u, 0u};
        int k = kept;
#pragma unroll
        for (int w = 0; w < kFChunk / 32; ++w) {
          if (w >= words) break;
          const unsigned in_range = (w * 32 + 32 <= m) ? ~0u : ((1u << (m - w * 32)) - 1u);
          unsigned alive_bits = ~rem[w] & in_range;
          while (alive_bits && k < max_det) {
            const int bit = __ffs(alive_bits) - 1, r = w * 32 + bit;
            S.klist[k - kept] = r;
            ++k;
            const uint4 row = *reinterpret_cast<const uint4 *>(S.mask[r]);
            rem[0] |= row.x; rem[1] |= row.y; rem[2] |= row.z; rem[3] |= row.w;
            alive_bits = ~rem[w] & in_range & ~((2u << bit) - 1u);
          }
        }
        S.kept = k;
      }
      __syncthreads();
      const int k_new = S.kept - kept;
      if (tid < k_new) {
        const int r = S.klist[tid], k = kept + tid;
        S.kbox[k] = S.abox[r]; S.kraw[k] = S.araw[r]; S.karea[k] = S.aarea[r]; S.kcls[k] = S.acls[r];
        S.kconf[k] = S.aconf[r]; S.kanchor[k] = S.aanchor[r];
      }
      __syncthreads();
    }
    processed += take;
    if (hi == ~0ull || S.kept >= max_det || processed >= max_nms) break;
    lo = hi;
    __syncthreads();
  }
  // ---- 4. emit ----
  __syncthreads();
  const int kept = S.kept;
  for (int k = tid; k < max_det; k += kFThreads) {  // every row is defined: zeros / index -1 past the count
    float *o = out_det + ((long long)b * max_det + k) * 6;
    if (k < kept) {
      const float4 r = S.kraw[k];
      o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; o[4] = S.kconf[k]; o[5] = (float)S.kcls[k];
    } else {
      o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0.f;
    }
    if (out_idx) out_idx[(long long)b * max_det + k] = k < kept ? S.kanchor[k] : -1;
  }
  if (tid == 0) out_count[b] = kept;
}

// ---- detection-record variant (custom head, postprocess.hpp semantics) ----------------------
struct DetLayout {
  uint64_t *keys_in, *keys_out;
  int *slot_in, *slot_out;
  float4 *box;
  int *cls;
  int *kept_rank;
  void *cub_tmp;
  size_t cub_bytes, total;
};

DetLayout carve_det(void *ws, int cap) {
  DetLayout L;
  Carver c(ws);
  L.keys_in = c.take<uint64_t>(cap);
  L.keys_out = c.take<uint64_t>(cap);
  L.slot_in = c.take<int>(cap);
  L.slot_out = c.take<int>(cap);
  L.box = c.take<float4>(cap);
  L.cls = c.take<int>(cap);
  L.kept_rank = c.take<int>(kMaxDetCap);
  L.cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, L.cub_bytes, (uint64_t *)nullptr, (uint64_t *)nullptr, (int *)nullptr,
                                  (int *)nullptr, cap, 0, 64);
  L.cub_tmp = c.take<char>(L.cub_bytes);
  L.total = c.off;
  return L;
}

__global__ void det_key_kernel(const uyd_detection *dets, const int *cell_idx, const int *d_count, int cap, DetLayout L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  const int n = min(*d_count, cap);
  L.slot_in[i] = i;
  if (i < n) {
    const uint32_t tie = cell_idx ? (uint32_t)cell_idx[i] : (uint32_t)i;
    L.keys_in[i] = ((uint64_t)(~__float_as_uint(dets[i].confidence)) << 32) | tie;
  } else {
    L.keys_in[i] = kInvalidKey;
  }
}

__global__ void det_gather_kernel(const uyd_detection *dets, const int *d_count, int cap, DetLayout L) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= min(*d_count, cap)) return;
  const uyd_detection d = dets[L.slot_out[r]];
  L.box[r] = make_float4(d.x1, d.y1, d.x2, d.y2);
  L.cls[r] = d.class_id;
}

__global__ void det_emit_kernel(const uyd_detection *dets, DetLayout L, const int *out_count, uyd_detection *out) {
  for (int k = threadIdx.x; k < *out_count; k += blockDim.x) {
    uyd_detection d = dets[L.slot_out[L.kept_rank[k]]];
    d.valid = 1;
    out[k] = d;
  }
}

}  // namespace
}  // namespace uyd

extern "C" size_t uyd_nms_workspace_bytes(int batch, int anchors) {
  if (batch <= 0 || anchors <= 0) return 0;
  return uyd::carve(nullptr, batch, anchors).total;
}

extern "C" int uyd_nms(uyd_ctx *ctx, const float *y, int batch, int nc, int anchors, float conf_thr, double iou_thr,
                       int max_nms, int max_det, float max_wh, void *workspace, size_t workspace_bytes, float *out_det,
                       int *out_idx, int *out_count, uyd_stream stream) {
  using namespace uyd;
  (void)ctx;
  UYD_REQUIRE(y && workspace && out_det && out_count && batch > 0 && nc > 0 && anchors > 0, UYD_E_ARG, "uyd_nms: bad arguments");
  UYD_REQUIRE(batch < (1 << 10) && anchors < (1 << kAnchorBits), UYD_E_UNSUPPORTED,
              "uyd_nms: batch < 1024 and anchors < 4M per call");
  UYD_REQUIRE(max_det > 0 && max_det <= kMaxDetCap, UYD_E_UNSUPPORTED, "uyd_nms: max_det <= %d", kMaxDetCap);
  if (max_nms > anchors) max_nms = anchors;
  UYD_REQUIRE(max_nms > 0, UYD_E_ARG, "uyd_nms: max_nms must be positive");
  cudaStream_t s = (cudaStream_t)stream;
  const long long total = (long long)batch * anchors;
  float thr_f = (float)iou_thr;  // largest float <= the double threshold
  if ((double)thr_f > iou_thr) thr_f = nextafterf(thr_f, -INFINITY);
  static const bool legacy = [] { const char *v = getenv("UYD_NMS_LEGACY"); return v && *v == '1'; }();
  if (!legacy && nc <= 256) {  // fused per-image kernel (needs no workspace)
    static bool attr_set = false;
    if (!attr_set) {
      UYD_CUDA(cudaFuncSetAttribute(nms_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem)));
      attr_set = true;
    }
    nms_image_kernel<<<batch, kFThreads, sizeof(FusedSmem), s>>>(y, nc, anchors, conf_thr, thr_f, max_nms, max_det, max_wh,
                                                                 conf_thr >= 0.f ? 1 : 0, out_det, out_idx, out_count);
    UYD_CUDA(cudaGetLastError());
    return UYD_OK;
  }
  Layout L = carve(workspace, batch, anchors);
  UYD_REQUIRE(L.total <= workspace_bytes, UYD_E_ARG, "uyd_nms: workspace too small (%zu < %zu)", workspace_bytes, L.total);
  // rows past the count: zeros / index -1 (the emit kernel only writes kept rows)
  UYD_CUDA(cudaMemsetAsync(out_det, 0, (size_t)batch * max_det * 6 * sizeof(float), s));
  if (out_idx) UYD_CUDA(cudaMemsetAsync(out_idx, 0xFF, (size_t)batch * max_det * sizeof(int), s));
  UYD_CUDA(cudaMemsetAsync(out_count, 0, (size_t)batch * sizeof(int), s));

  UYD_CUDA(cudaMemsetAsync(L.count, 0, (size_t)batch * 4, s));
  const bool pos = conf_thr >= 0.f;  // NaN-safe: a NaN threshold selects nothing either way
  if (pos) nms_key_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(y, batch, nc, anchors, conf_thr, L.keys_in, L.count);
  else nms_key_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(y, batch, nc, anchors, conf_thr, L.keys_in, L.count);
  UYD_CUDA(cudaGetLastError());
  nms_scan_kernel<<<1, 32, 0, s>>>(L.count, L.offset, batch);
  size_t tmp = L.cub_bytes;
  // keys arrive in anchor order and the sort is stable: the anchor bits need no pass
  int image_bits = 1;
  while ((1 << image_bits) < batch) ++image_bits;
  const int end_bit = (pos ? kImageShift - 1 : kImageShift) + image_bits;
  UYD_CUDA(cub::DeviceRadixSort::SortKeys(L.cub_tmp, tmp, L.keys_in, L.keys_out, (int)total, kScoreShift, end_bit, s));
  dim3 ggrid((unsigned)ceil_div(max_nms, 256), (unsigned)batch);
  nms_gather_kernel<<<ggrid, 256, 0, s>>>(y, nc, anchors, max_nms, max_wh, L.keys_out, L);
  UYD_CUDA(cudaGetLastError());
  int e = launch_greedy<false>(L.boxoff, L.cls, anchors, L.count, max_nms, max_det, thr_f, L.kept_rank, out_count, batch, s);
  if (e) return e;
  nms_emit_rows_kernel<<<batch, 128, 0, s>>>(L, anchors, max_det, out_count, out_det, out_idx);
  UYD_CUDA(cudaGetLastError());
  return UYD_OK;
}

extern "C" size_t uyd_nms_detections_workspace_bytes(int cap) { return cap > 0 ? uyd::carve_det(nullptr, cap).total : 0; }

extern "C" int uyd_nms_detections(uyd_ctx *ctx, const uyd_detection *dets, const int *cell_idx, const int *d_count, int cap,
                                  float iou_thr, void *workspace, size_t workspace_bytes, uyd_detection *out,
                                  int *d_out_count, uyd_stream stream) {
  using namespace uyd;
  (void)ctx;
  UYD_REQUIRE(dets && d_count && workspace && out && d_out_count && cap > 0, UYD_E_ARG, "uyd_nms_detections: bad arguments");
  DetLayout L = carve_det(workspace, cap);
  UYD_REQUIRE(L.total <= workspace_bytes, UYD_E_ARG, "uyd_nms_detections: workspace too small (%zu < %zu)", workspace_bytes,
              L.total);
  cudaStream_t s = (cudaStream_t)stream;
  det_key_kernel<<<ceil_div(cap, 256), 256, 0, s>>>(dets, cell_idx, d_count, cap, L);
  UYD_CUDA(cudaGetLastError());
  size_t tmp = L.cub_bytes;
  UYD_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_tmp, tmp, L.keys_in, L.keys_out, L.slot_in, L.slot_out, cap, 0, 64, s));
  det_gather_kernel<<<ceil_div(cap, 256), 256, 0, s>>>(dets, d_count, cap, L);
  UYD_CUDA(cudaGetLastError());
  int e = launch_greedy<true>(L.box, L.cls, cap, d_count, cap, kMaxDetCap, iou_thr, L.kept_rank, d_out_count, 1, s);
  if (e) return e;
  det_emit_kernel<<<1, 256, 0, s>>>(dets, L, d_out_count, out);
  UYD_CUDA(cudaGetLastError());
  return UYD_OK;
}
