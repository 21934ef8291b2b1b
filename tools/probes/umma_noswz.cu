// Probe: tcgen05.mma with K-major NO-SWIZZLE shared-memory descriptors whose two 16-byte K chunks are two 3x3 TAPS of a
// flat 16-byte-per-pixel activation frame (8 bf16 channels per pixel): LBO = byte distance between the two taps'
// pixels (16 B for horizontally adjacent taps: the two "core matrices" overlap), SBO = 128 B (8 consecutive pixels).
// A 3x3 conv 8 -> 8 over a flat frame is then 5 MMAs (K = 16 = 2 taps x 8 channels) per 128-pixel M-tile, N = 16
// (8 real output channels).  Checks the result against a host loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_noswz umma_noswz.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {  // layout type 0 = no swizzle
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

constexpr int PITCH = 42, ROWS = 12, FR = PITCH * ROWS;  // flat frame of 504 pixels, 16 B each
constexpr int F0 = PITCH + 1;                            // first output pixel (row 1, col 1): 3 M-tiles of 128 from there

__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16 *frame, const __nv_bfloat16 *wpk, float *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) unsigned long long bar;
  unsigned char *fs = smem;                 // frame [FR + slack][8] bf16
  unsigned char *ws = smem + 16384;         // weights: 5 pairs x [kchunk 2][n 16][8]
  for (int i = threadIdx.x; i < (FR + 64) * 4; i += 128) reinterpret_cast<uint32_t *>(fs)[i] = i < FR * 4 ? reinterpret_cast<const uint32_t *>(frame)[i] : 0u;
  for (int i = threadIdx.x; i < 5 * 2 * 16 * 4; i += 128) reinterpret_cast<uint32_t *>(ws)[i] = reinterpret_cast<const uint32_t *>(wpk)[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int tap_off[10] = {-PITCH - 1, -PITCH, -PITCH + 1, -1, 0, 1, PITCH - 1, PITCH, PITCH + 1, PITCH + 2};
  if (warp == 1) {
    if (elect_one()) {
      for (int mt = 0; mt < 3; ++mt)
        for (int pr = 0; pr < 5; ++pr) {
          const int o0 = tap_off[2 * pr], o1 = tap_off[2 * pr + 1];
          const uint64_t ad = desc(smem_u32(fs) + (uint32_t)((F0 + 128 * mt + o0) * 16), (uint32_t)((o1 - o0) * 16), 128);
          const uint64_t bd = desc(smem_u32(ws) + (uint32_t)(pr * 512), 256, 128);
          umma(tmem + 16 * mt, ad, bd, idesc, pr != 0);
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
  }
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int mt = 0; mt < 3; ++mt) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + 16 * mt) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[((mt * 128) + warp * 32 + lane) * 16 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
  std::vector<__nv_bfloat16> frame(FR * 8), wpk(5 * 2 * 16 * 8);
  std::vector<float> ff(FR * 8), w(8 * 8 * 9);
  srand(1);
  for (int i = 0; i < FR * 8; ++i) { float v = (rand() % 2001 - 1000) / 500.0f; frame[i] = __float2bfloat16(v); ff[i] = __bfloat162float(frame[i]); }
  for (int i = 0; i < 8 * 8 * 9; ++i) { float v = (rand() % 2001 - 1000) / 4000.0f; w[i] = __bfloat162float(__float2bfloat16(v)); }
  // wpk[pair][kchunk][n][ci]: tap = 2 pair + kchunk (tap 9 = zero), n >= 8 zero
  for (int pr = 0; pr < 5; ++pr)
    for (int kc = 0; kc < 2; ++kc)
      for (int n = 0; n < 16; ++n)
        for (int ci = 0; ci < 8; ++ci) {
          const int tap = 2 * pr + kc;
          const float v = (tap < 9 && n < 8) ? w[(n * 8 + ci) * 9 + tap] : 0.f;
          wpk[((pr * 2 + kc) * 16 + n) * 8 + ci] = __float2bfloat16(v);
        }
  __nv_bfloat16 *dF, *dW; float *dO;
  cudaMalloc(&dF, frame.size() * 2); cudaMalloc(&dW, wpk.size() * 2); cudaMalloc(&dO, 3 * 128 * 16 * 4);
  cudaMemcpy(dF, frame.data(), frame.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, wpk.data(), wpk.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<<<1, 128, 32768>>>(dF, dW, dO);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> o(3 * 128 * 16);
  cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost);
  const int tap_off[9] = {-PITCH - 1, -PITCH, -PITCH + 1, -1, 0, 1, PITCH - 1, PITCH, PITCH + 1};
  double worst = 0;
  for (int m = 0; m < 384; ++m) {
    const int f = F0 + m;
    if (f + PITCH + 1 >= FR) continue;
    for (int n = 0; n < 16; ++n) {
      double ref = 0;
      if (n < 8)
        for (int t = 0; t < 9; ++t)
          for (int ci = 0; ci < 8; ++ci) ref += (double)ff[(f + tap_off[t]) * 8 + ci] * w[(n * 8 + ci) * 9 + t];
      worst = fmax(worst, fabs(ref - o[m * 16 + n]));
    }
  }
  printf("no-swizzle two-tap MMA: %s, max |gpu - host| = %.3g (%s)\n", worst < 1e-3 ? "MATCH" : "MISMATCH", worst, cudaGetErrorString(e));
  return worst < 1e-3 ? 0 : 1;
}
