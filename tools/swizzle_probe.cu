// Developer probe: does a tcgen05 K-major SWIZZLE_128B / SWIZZLE_64B shared-memory descriptor accept a
// start address shifted by a number of ROWS that is not a multiple of 8 (what conv_tc.cu's "a conv tap
// is a row shift" needs), and which matrix-base-offset value makes it correct?
//   For each row shift j = 0..10 and each base-offset rule (0 = always 0, 1 = (start >> 7) & 7) it runs
//   D_j[r][n] = sum_k A[r + j][k] * B[n][k] (r < 128, n < 16, k < K) on small-integer fp16 data and counts
//   mismatches against the CPU.  One CTA, timing-free.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                 \
  do {                                                                                        \
    cudaError_t e_ = (x);                                                                     \
    if (e_ != cudaSuccess) {                                                                  \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(3);                                                                                \
    }                                                                                         \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int ROWS = 160, NB = 16;

__host__ __device__ inline int a_val(int r, int k) { return (r * 7 + k * 3) % 13 - 6; }
__host__ __device__ inline int b_val(int n, int k) { return (n * 5 + k * 11) % 9 - 4; }

// swz = 128: rows of 64 halves (128 B), 16 B chunk c of row r stored at chunk c ^ (r & 7)
// swz = 64 : rows of 32 halves ( 64 B), chunk c of row r stored at chunk c ^ ((r >> 1) & 3)
__global__ void __launch_bounds__(128, 1) swizzle_probe_kernel(int swz, int shift, int rule, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int K = swz == 128 ? 64 : 32, pitch = K * 2;
  __half* As = reinterpret_cast<__half*>(smem);                // ROWS x pitch bytes, swizzled
  __half* Bs = reinterpret_cast<__half*>(smem + 32 * 1024);    // no swizzle: [k-group][n][8]
  for (int i = threadIdx.x; i < ROWS * K; i += 128) {
    const int r = i / K, k = i % K, c = k >> 3;
    const int cs = swz == 128 ? (c ^ (r & 7)) : (c ^ ((r >> 1) & 3));
    As[(r * pitch + cs * 16 + (k & 7) * 2) / 2] = __float2half((float)a_val(r, k));
  }
  for (int i = threadIdx.x; i < NB * K; i += 128) {
    const int n = i / K, k = i % K;
    Bs[((k >> 3) * NB + n) * 8 + (k & 7)] = __float2half((float)b_val(n, k));
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 32) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(NB >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a0 = smem_u32(As) + (uint32_t)shift * pitch, b0 = smem_u32(Bs);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t sa = a0 + ks * 32;
      const uint64_t layout = swz == 128 ? 2ull : 4ull;
      const uint64_t sbo = swz == 128 ? 1024 : 512;
      const uint64_t boff = rule ? ((sa >> 7) & 7u) : 0u;
      const uint64_t da = (uint64_t)((sa & 0x3FFFFu) >> 4) | (1ull << 16) | ((sbo >> 4) << 32) | (1ull << 46) | (boff << 49) | (layout << 61);
      const uint32_t sb = b0 + ks * 2 * (NB * 16);
      const uint64_t db = (uint64_t)((sb & 0x3FFFFu) >> 4) | ((uint64_t)((NB * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
      const uint32_t acc = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everyone waits for the MMAs
  {
    uint32_t done = 0;
    for (long long spin = 0; spin < (1ll << 26) && !done; ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done)
                   : "r"(smem_u32(&bar)), "r"(0)
                   : "memory");
    if (!done) __trap();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t v[16];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < 16; ++n) out[(warp * 32 + lane) * 16 + n] = __uint_as_float(v[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}

int main(int argc, char** argv) {
  const int only_swz = argc > 1 ? atoi(argv[1]) : 0, only_rule = argc > 2 ? atoi(argv[2]) : -1;
  float* d;
  CK(cudaMalloc(&d, 128 * 16 * 4));
  float h[128 * 16];
  CK(cudaFuncSetAttribute(swizzle_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  for (int swz : {128, 64}) {
    if (only_swz && swz != only_swz) continue;
    const int K = swz == 128 ? 64 : 32;
    for (int rule = 0; rule < 2; ++rule) {
      if (only_rule >= 0 && rule != only_rule) continue;
      printf("swizzle %3dB  base-offset rule %d (%s): mismatches per row shift 0..10:", swz, rule, rule ? "(start>>7)&7" : "0");
      for (int shift = 0; shift <= 10; ++shift) {
        swizzle_probe_kernel<<<1, 128, 64 * 1024>>>(swz, shift, rule, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf(" [launch failed: %s]", cudaGetErrorString(e));
          break;
        }
        CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int r = 0; r < 128; ++r)
          for (int n = 0; n < 16; ++n) {
            int ref = 0;
            for (int k = 0; k < K; ++k) ref += a_val(r + shift, k) * b_val(n, k);
            if (h[r * 16 + n] != (float)ref) ++bad;
          }
        printf(" %d", bad);
      }
      printf("\n");
    }
  }
  return 0;
}
