// Probe: issue rate / execution time of back-to-back tcgen05.mma (kind::f16, M = 128, K = 16) for several N,
// operands in shared memory (SWIZZLE_128B K-major, contents irrelevant), one CTA per SM.
// Variants: issue under `if (lane == 0)` (divergent: ptxas wraps every UTCHMMA in an ELECT loop) vs
// under elect.sync on a converged warp; same accumulator vs alternating accumulators.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
template <bool I8>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (I8)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// I8: kind::i8 (int8 x int8 -> int32, K = 32 per instruction at the same 32 operand bytes per row)
template <int MODE, bool I8 = false>  // 0: if (lane == 0), 1: elect.sync, 2: elect + alternate accumulators every 4 MMAs
__global__ void __launch_bounds__(128, 1) probe(int N, int iters, long long *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const uint32_t idesc = ((I8 ? 2u : 1u) << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t ad = make_desc(base, 1024), bd = make_desc(base + 16384, 1024);
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 1) {
    t0 = clock64();
    if (MODE == 0) {
      if (lane == 0) {
#pragma unroll 1
        for (int i = 0; i < iters; i += 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma<I8>(tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
        }
      }
      __syncwarp();
    } else {
#pragma unroll 1
      for (int i = 0; i < iters; i += 4) {
        const uint32_t d = (MODE == 2 && (i & 4)) ? tmem + 256 : tmem;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma<I8>(d, ad + 2 * k, bd + 2 * k, idesc, 1u);
        }
        __syncwarp();
      }
    }
    t1 = clock64();
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// no-swizzle K-major descriptors (layout type 0): the layout of csrc/c3k_tc.cu.  lbo / sbo in bytes.
__device__ __forceinline__ uint64_t make_desc_ns(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__global__ void __launch_bounds__(128, 1) probe_ns(int N, int iters, uint32_t lbo, long long *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  long long t0 = 0, t2 = 0;
  if (warp == 1) {
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i += 4) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma<false>(tmem + 32 * (k & 3), make_desc_ns(base + 2048 * k, lbo, 128), make_desc_ns(base + 32768 + 512 * k, N * 16, 128), idesc, 1u);
      }
      __syncwarp();
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    t2 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) { out[0] = t2 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// Items of G MMAs + one tcgen05.commit each (the shape of csrc/c3k_tc.cu), issued by NW warps at once: how long one item
// takes the issuing warp (issue only) and the whole CTA (until the last commit has arrived).
template <int G, int CE, bool ELECT>  // G MMAs per item (unrolled, descriptors in registers), a commit every CE items
__global__ void __launch_bounds__(256, 1) probe_commit(int NW, int iters, long long *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) unsigned long long bars[8][4], last[8];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 256) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int w = 0; w < 8; ++w) {
      for (int k = 0; k < 4; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[w][k])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&last[w])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 7) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (warp < NW) {
    uint64_t ad[G], bd[G];
#pragma unroll
    for (int k = 0; k < G; ++k) {
      ad[k] = make_desc_ns(base + 16 * k, 16, 128);
      bd[k] = make_desc_ns(base + 32768 + 512 * k, 256, 128);
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
      if (ELECT ? elect_one() : lane == 0) {
        const uint32_t d = tmem + 64 * warp + 16 * (i & 3);
        const uint64_t f = (uint64_t)(128u * (uint32_t)(i & 7));
#pragma unroll
        for (int k = 0; k < G; ++k) umma<false>(d, ad[k] + f, bd[k], idesc, k != 0);
        if (i % CE == CE - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[warp][i & 3])) : "memory");
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&last[warp])) : "memory");
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&last[warp])) : "memory");
    }
    const long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0 && warp == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 7) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// Layout / stride / start-offset variant (conv_tc.cu HALO): A descriptor with layout code `lay` (2 = SWIZZLE_128B,
// 4 = SWIZZLE_64B, 6 = SWIZZLE_32B), stride `sbo` bytes between 8-row groups, nine "taps" that shift the start by
// row_b * (t / 3) + px_b * (t % 3) bytes and `ks` k-steps of 32 bytes each; B rows of px_b bytes (same layout code).
__device__ __forceinline__ uint64_t make_desc_l(uint32_t addr, uint32_t sbo, uint32_t lay) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)lay << 61;
  return d;
}
__global__ void __launch_bounds__(128, 1) probe_sw(int N, int iters, uint32_t lay, uint32_t sbo, uint32_t row_b, uint32_t px_b, int ks, long long *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 100 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t ad = make_desc_l(base, sbo, lay), bd = make_desc_l(base + 65536, 8 * px_b, lay);
  if (warp == 1) {
    const long long t0 = clock64();
    int n = 0;
#pragma unroll 1
    for (int i = 0; n < iters; ++i) {
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint64_t a = ad + ((row_b * (t / 3) + px_b * (t % 3)) >> 4), b = bd + ((uint32_t)(t % 3) * N * px_b >> 4);
          for (int k = 0; k < ks; ++k) umma<false>(tmem, a + 2 * k, b + 2 * k, idesc, 1u);
        }
      }
      __syncwarp();
      n += 9 * ks;
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    const long long t2 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) { out[0] = n; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long *d, h[2];
  cudaMalloc(&d, 16);
  const int iters = 4000;
  auto run = [&](int mode, int N) {
    const size_t smem = 49 * 1024;
    void (*k)(int, int, long long *) = mode == 0 ? probe<0> : (mode == 1 ? probe<1> : probe<2>);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) k<<<148, 128, smem>>>(N, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("mode %d N %3d: issue %.1f cyc/mma, issue+drain %.1f cyc/mma  (%s)\n", mode, N, (double)h[0] / iters, (double)h[1] / iters,
           cudaGetErrorString(e));
  };
  for (int mode = 0; mode < 3; ++mode)
    for (int N : {16, 32, 64, 128, 256}) run(mode, N);
  // no-swizzle descriptors (c3k_tc.cu): N = 16 / 32, LBO = 16 B (two horizontal taps) / 1312 B / 32 KB (planes)
  for (int N : {16, 32})
    for (uint32_t lbo : {16u, 1312u, 16384u}) {
      const size_t smem = 49 * 1024;
      cudaFuncSetAttribute(probe_ns, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      for (int rep = 0; rep < 2; ++rep) probe_ns<<<148, 128, smem>>>(N, iters, lbo, d);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("no-swizzle N %3d LBO %5u: %.1f cyc/mma (%s)\n", N, lbo, (double)h[1] / iters, cudaGetErrorString(e));
    }
  // conv_tc HALO geometry: swizzle code, 8-row-group stride, tap shifts
  {
    struct C { const char *name; uint32_t lay, sbo, row_b, px_b; int ks; };
    const C cases[] = {
        {"128B rows, aligned groups (sbo 1024), no tap shift", 2, 1024, 0, 128, 4},
        {"128B rows, HALO dense box (sbo 1280, taps)        ", 2, 1280, 1280, 128, 4},
        {"128B rows, HALO padded rows (sbo 2048, taps)      ", 2, 2048, 2048, 128, 4},
        {" 64B rows, aligned groups (sbo 512), no tap shift ", 4, 512, 0, 64, 2},
        {" 64B rows, HALO dense box (sbo 640, taps)         ", 4, 640, 640, 64, 2},
        {" 64B rows, HALO padded rows (sbo 1024, taps)      ", 4, 1024, 1024, 64, 2},
        {" 32B rows, aligned groups (sbo 256), no tap shift ", 6, 256, 0, 32, 1},
        {" 32B rows, HALO dense box (sbo 320, taps)         ", 6, 320, 320, 32, 1},
    };
    const size_t smem = 200 * 1024;
    cudaFuncSetAttribute(probe_sw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (const C &c : cases)
      for (int N : {32, 64, 128}) {
        for (int rep = 0; rep < 2; ++rep) probe_sw<<<148, 128, smem>>>(N, iters, c.lay, c.sbo, c.row_b, c.px_b, c.ks, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("geometry %s N %3d: %.1f cyc/mma (%s)\n", c.name, N, (double)h[1] / (double)h[0], cudaGetErrorString(e));
      }
  }
  // items of G MMAs + commit from NW warps
  auto run_commit = [&](void (*k)(int, int, long long *), const char *name, int G) {
    for (int NW : {1, 2, 6}) {
      const size_t smem = 49 * 1024;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      for (int rep = 0; rep < 2; ++rep) k<<<148, 256, smem>>>(NW, 1000, d);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("commit %s NW %d: issue %.0f cyc/item, issue+drain %.0f cyc/item per warp = %.1f cyc/mma overall (%s)\n", name, NW, h[0] / 1000.0, h[1] / 1000.0,
             (double)h[1] / (1000.0 * G * NW), cudaGetErrorString(e));
    }
  };
  run_commit(probe_commit<1, 1, true>, "G 1 commit/1 elect", 1);
  run_commit(probe_commit<1, 8, true>, "G 1 commit/8 elect", 1);
  run_commit(probe_commit<5, 1, true>, "G 5 commit/1 elect", 5);
  run_commit(probe_commit<5, 8, true>, "G 5 commit/8 elect", 5);
  run_commit(probe_commit<5, 1, false>, "G 5 commit/1 lane0", 5);
  // kind::i8 under elect.sync: cycles per MMA and the implied dense INT8 rate of the whole GPU
  int dev = 0, clk_khz = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (int N : {16, 32, 64, 128, 256}) {
    const size_t smem = 49 * 1024;
    cudaFuncSetAttribute(probe<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) probe<1, true><<<148, 128, smem>>>(N, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const double cyc = (double)h[1] / iters;
    printf("i8   N %3d: %.1f cyc/mma -> %.0f dense INT8 TOP/s at %d SMs x %.3f GHz (max clock; %s)\n", N, cyc,
           2.0 * 128 * N * 32 / cyc * sms * clk_khz * 1e3 / 1e12, sms, clk_khz / 1e6, cudaGetErrorString(e));
  }
  return 0;
}
