// Developer micro-benchmark: issue cost of tcgen05.mma (kind::f16, M=128, cta_group::1) under the
// accumulator / descriptor patterns conv_tc.cu could use.  One CTA per SM, operands are whatever
// is in shared memory (timing only).  Prints cycles per MMA instruction for each pattern.
//   mma_bench [iters]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_));   \
      exit(3);                                                                                  \
    }                                                                                           \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (long long spin = 0; spin < (1ll << 28); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// SWIZZLE_NONE K-major descriptor
__device__ __forceinline__ uint64_t desc_none(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// SWIZZLE_128B K-major descriptor (layout type 2), SBO = 1024 B
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc_f16(int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24); }

enum Pattern { P_SAME128, P_SAME256, P_SHIFT128, P_PAIR_OVERLAP, P_TRIPLE_V1, P_ALT2, P_PAIR_DISJOINT, P_SAME64, P_SAME32,
               P_SW128_128, P_SW128_256, P_PAIR_SMALL64, P_PAIR_SMALL32, P_GROUPED_PAIR,
               P_SW128_SHIFT128, P_SW64_SHIFT128, P_SW64_SHIFT64, P_SW64_SHIFT32, P_SW64_PAIR128, P_SW64_PAIR32, P_COUNT };
static const char* kNames[] = {"same D, N=128", "same D, N=256", "same D, N=128, A row-shifted per MMA", "v2 pair: N=256->D0, N=128->D0+128 (overlap)",
                               "v1 triple: D1,D1,D0 (N=128)", "alternate D0,D1 (N=128)", "pair: N=256->D0, N=128->D256 (disjoint)",
                               "same D, N=64", "same D, N=32", "SW128 same D, N=128", "SW128 same D, N=256",
                               "v2 pair at N=64: N=128->D0, N=64->D0+64", "v2 pair at N=32: N=64->D0, N=32->D0+32",
                               "v2 pair grouped: 8x(N=256->D0) then 8x(N=128->D0+128)",
                               "SW128 A row-shifted per MMA, N=128", "SW64 A row-shifted per MMA, N=128", "SW64 A row-shifted, N=64",
                               "SW64 A row-shifted, N=32", "SW64 A row-shifted, v2 pair N=256/128", "SW64 A row-shifted, v2 pair N=64/32"};

// One MMA of pattern P at iteration `it` (all descriptor values are loop-invariant registers).
template <int P>
__device__ __forceinline__ void issue(int it, uint32_t tmem, uint64_t dA, uint64_t dA2, uint64_t dB, uint64_t dB2, uint64_t sA, uint64_t sB) {
  const uint32_t acc = it > 0;
  if constexpr (P == P_SAME128) umma_f16(tmem, dA, dB, idesc_f16(128), acc);
  if constexpr (P == P_SAME256) umma_f16(tmem, dA, dB, idesc_f16(256), acc);
  if constexpr (P == P_SAME64) umma_f16(tmem, dA, dB, idesc_f16(64), acc);
  if constexpr (P == P_SAME32) umma_f16(tmem, dA, dB, idesc_f16(32), acc);
  if constexpr (P == P_SHIFT128) umma_f16(tmem, dA + (uint64_t)(it % 11), dB, idesc_f16(128), acc);
  if constexpr (P == P_PAIR_OVERLAP) {
    if ((it & 1) == 0) umma_f16(tmem, dA, dB, idesc_f16(256), acc);
    else umma_f16(tmem + 128, dA2, dB, idesc_f16(128), 1);
  }
  if constexpr (P == P_PAIR_SMALL64) {
    if ((it & 1) == 0) umma_f16(tmem, dA, dB, idesc_f16(128), acc);
    else umma_f16(tmem + 64, dA2, dB, idesc_f16(64), 1);
  }
  if constexpr (P == P_PAIR_SMALL32) {
    if ((it & 1) == 0) umma_f16(tmem, dA, dB, idesc_f16(64), acc);
    else umma_f16(tmem + 32, dA2, dB, idesc_f16(32), 1);
  }
  if constexpr (P == P_GROUPED_PAIR) {
    if (((it >> 3) & 1) == 0) umma_f16(tmem, dA, dB, idesc_f16(256), acc);
    else umma_f16(tmem + 128, dA2, dB, idesc_f16(128), 1);
  }
  if constexpr (P == P_PAIR_DISJOINT) {
    if ((it & 1) == 0) umma_f16(tmem, dA, dB, idesc_f16(256), acc);
    else umma_f16(tmem + 256, dA2, dB, idesc_f16(128), it > 1);
  }
  if constexpr (P == P_TRIPLE_V1) {
    const int r = it % 3;
    umma_f16(tmem + (r == 2 ? 0 : 128), r == 0 ? dA2 : dA, r == 1 ? dB2 : dB, idesc_f16(128), it > 2);
  }
  if constexpr (P == P_ALT2) umma_f16(tmem + (it & 1) * 128, dA, dB, idesc_f16(128), it > 1);
  if constexpr (P == P_SW128_128) umma_f16(tmem, sA + (uint64_t)((it & 3) * 2), sB + (uint64_t)((it & 3) * 2), idesc_f16(128), acc);
  if constexpr (P == P_SW128_256) umma_f16(tmem, sA + (uint64_t)((it & 3) * 2), sB + (uint64_t)((it & 3) * 2), idesc_f16(256), acc);
  // swizzled A (row shift = (it % 11) rows), B stays in the no-swizzle image like conv_tc.cu
  if constexpr (P == P_SW128_SHIFT128) umma_f16(tmem, sA + (uint64_t)((it % 11) * 8 + (it & 3) * 2), dB, idesc_f16(128), acc);
  const uint64_t s64 = (sA & ~(7ull << 61) & ~(0x3FFFull << 32)) | (4ull << 61) | ((uint64_t)(512 >> 4) << 32);
  if constexpr (P == P_SW64_SHIFT128) umma_f16(tmem, s64 + (uint64_t)((it % 11) * 4 + (it & 1) * 2), dB, idesc_f16(128), acc);
  if constexpr (P == P_SW64_SHIFT64) umma_f16(tmem, s64 + (uint64_t)((it % 11) * 4 + (it & 1) * 2), dB, idesc_f16(64), acc);
  if constexpr (P == P_SW64_SHIFT32) umma_f16(tmem, s64 + (uint64_t)((it % 11) * 4 + (it & 1) * 2), dB, idesc_f16(32), acc);
  if constexpr (P == P_SW64_PAIR128) {
    if ((it & 1) == 0) umma_f16(tmem, s64 + (uint64_t)((it % 11) * 4), dB, idesc_f16(256), acc);
    else umma_f16(tmem + 128, s64 + (uint64_t)(2048 + (it % 11) * 4), dB, idesc_f16(128), 1);
  }
  if constexpr (P == P_SW64_PAIR32) {
    if ((it & 1) == 0) umma_f16(tmem, s64 + (uint64_t)((it % 11) * 4), dB, idesc_f16(64), acc);
    else umma_f16(tmem + 32, s64 + (uint64_t)(2048 + (it % 11) * 4), dB, idesc_f16(32), 1);
  }
}

template <int P>
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int iters, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // fill operands with small finite fp16 values
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 32) {
    const uint32_t a0 = smem_u32(smem);                // A region: 64 KB
    const uint32_t b0 = smem_u32(smem + 64 * 1024);    // B region: 64 KB
    const uint32_t a_plane = 184 * 16, b_plane = 256 * 16;
    const uint64_t dA = desc_none(a0, a_plane, 128), dA2 = desc_none(a0 + 32768, a_plane, 128);
    const uint64_t dB = desc_none(b0, b_plane, 128), dB2 = desc_none(b0 + 32768, b_plane, 128);
    const uint64_t sA = desc_sw128(a0), sB = desc_sw128(b0);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it += 24) {
#pragma unroll
      for (int u = 0; u < 24; ++u) issue<P>(it + u, tmem, dA, dA2, dB, dB2, sA, sB);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out_cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int P>
void run_pattern(int iters, int sms, int smem, long long* d, long long* h) {
  CK(cudaFuncSetAttribute(mma_bench_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rep = 0; rep < 2; ++rep) {
    mma_bench_kernel<P><<<sms, 128, smem>>>(iters, d);
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(h, d, sms * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mn = h[0], mx = h[0];
  double sum = 0;
  for (int i = 0; i < sms; ++i) {
    if (h[i] < mn) mn = h[i];
    if (h[i] > mx) mx = h[i];
    sum += (double)h[i];
  }
  printf("%-62s cycles/MMA  min %7.1f  avg %7.1f  max %7.1f\n", kNames[P], (double)mn / iters, sum / sms / iters, (double)mx / iters);
}

template <int P>
void run_all(int iters, int sms, int smem, long long* d, long long* h) {
  if constexpr (P < P_COUNT) {
    run_pattern<P>(iters, sms, smem, d, h);
    run_all<P + 1>(iters, sms, smem, d, h);
  }
}

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 2400;
  iters = (iters + 23) / 24 * 24;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* d;
  CK(cudaMalloc(&d, sms * sizeof(long long)));
  const int smem = 160 * 1024;
  long long* h = (long long*)malloc(sms * sizeof(long long));
  run_all<0>(iters, sms, smem, d, h);
  return 0;
}
