// Developer micro-benchmark #2 (round 2): what a tcgen05.mma costs when the SHIFTED operand is A or B, per shift residue,
// with M = 64, with A in tensor memory, with 32 B-swizzled rows -- and how fast tcgen05.ld drains TMEM.
// The conv kernels read one staged activation tile at a row offset per tap; profiles/r1_mma_bench.log showed such a
// row-shifted A operand costing 74 cycles per MMA whatever N is, which pins the C <= 64 stages.  This probe asks which
// mapping (time on M or on N, weights in smem or TMEM) avoids that.  Patterns are runtime tables in constant memory.
//   mma_bench2 [iters]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_));   \
      exit(3);                                                                                  \
    }                                                                                           \
  } while (0)

constexpr int PERIOD = 24;
struct Step {
  uint32_t a_lo, a_hi, b_lo, b_hi, idesc, dcol, a_tmem_col, a_in_tmem;
};
__constant__ Step c_steps[PERIOD];

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
__device__ __forceinline__ void umma_ss(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                        uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <bool ts>
__global__ void __launch_bounds__(128, 1) mma_kernel(int iters, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
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
    const uint32_t base = smem_u32(smem) >> 4;  // descriptors in the table are relative to the dynamic smem base
    // the whole table in registers (compile-time indices after unrolling): nothing but the MMA issue is in the timed loop
    uint32_t a_lo[PERIOD], b_lo[PERIOD], a_hi[PERIOD], b_hi[PERIOD], id[PERIOD], dc[PERIOD];
#pragma unroll
    for (int u = 0; u < PERIOD; ++u) {
      a_lo[u] = ts ? tmem + c_steps[u].a_tmem_col : c_steps[u].a_lo + base, a_hi[u] = c_steps[u].a_hi;
      b_lo[u] = c_steps[u].b_lo + base, b_hi[u] = c_steps[u].b_hi, id[u] = c_steps[u].idesc, dc[u] = tmem + c_steps[u].dcol;
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it += PERIOD) {
#pragma unroll
      for (int u = 0; u < PERIOD; ++u) {
        if (ts) umma_ts(dc[u], a_lo[u], b_lo[u], b_hi[u], id[u], 1);
        else umma_ss(dc[u], a_lo[u], a_hi[u], b_lo[u], b_hi[u], id[u], 1);
      }
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

// TMEM drain rate: `nwarps` warps (warp w reads lane quadrant w % 4) each issue `iters` tcgen05.ld.32x32b.x32 (4 KB per warp).
__global__ void __launch_bounds__(512, 1) ldtm_kernel(int iters, int x16, long long* out_cycles) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    uint32_t v[32];
    const uint32_t col = (uint32_t)((it * 32 + (warp >> 2) * 64) & 511 & ~31);
    if (x16) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr + col)
          : "memory");
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr + col + 16)
          : "memory");
    } else {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
          "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr + col)
          : "memory");
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= v[i];
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out_cycles[blockIdx.x * 16 + warp] = t1 - t0 + (acc == 0x12345678u);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------ host: pattern tables
enum Layout { SW64 = 4, SW32 = 6, SW128 = 2 };
static uint64_t desc(uint32_t off, Layout l) {  // K-major swizzled descriptor, address relative to the smem base
  const uint32_t sbo = l == SW64 ? 512 : l == SW32 ? 256 : 1024;
  return (uint64_t)(off >> 4) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | ((uint64_t)l << 61);
}
static uint32_t idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
static const uint32_t A0 = 0, B0 = 96 * 1024;  // A region 96 KB, B region 96 KB

struct Pat {
  std::string name;
  Step s[PERIOD];
};
static Step mk(uint64_t a, uint64_t b, int M, int N, int dcol) {
  Step s{};
  s.a_lo = (uint32_t)a, s.a_hi = (uint32_t)(a >> 32), s.b_lo = (uint32_t)b, s.b_hi = (uint32_t)(b >> 32);
  s.idesc = idesc(M, N), s.dcol = dcol;
  return s;
}
// shift of `rows` rows of `pitch` bytes + k-group kg (32 B) inside the row
static uint32_t off(uint32_t base, int rows, int pitch, int kg) { return base + rows * pitch + kg * 32; }

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 2400;
  iters = (iters + PERIOD - 1) / PERIOD * PERIOD;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* d;
  CK(cudaMalloc(&d, sms * 16 * sizeof(long long)));
  std::vector<long long> h(sms * 16);
  const int smem = 192 * 1024;
  CK(cudaFuncSetAttribute(mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));

  std::vector<Pat> pats;
  auto add = [&](const std::string& name, auto fn) {
    Pat p;
    p.name = name;
    for (int u = 0; u < PERIOD; ++u) p.s[u] = fn(u);
    pats.push_back(p);
  };
  char nm[160];
  // 1. time on M (today's mapping): A = activations SW64, shifted by a FIXED residue r; B = weights unshifted
  for (int N : {32, 64, 128})
    for (int r = 0; r < 8; ++r) {
      snprintf(nm, sizeof nm, "A(sw64) shift r=%d rows, B fixed, M=128 N=%d", r, N);
      add(nm, [&](int u) { return mk(desc(off(A0, r + 8 * (u % 3), 64, u & 1), SW64), desc(B0, SW64), 128, N, 0); });
    }
  // 2. time on N: B = activations SW64 shifted by residue r, A = weights unshifted (M=128)
  for (int N : {128, 256})
    for (int r = 0; r < 8; ++r) {
      snprintf(nm, sizeof nm, "B(sw64) shift r=%d rows, A fixed, M=128 N=%d", r, N);
      add(nm, [&](int u) { return mk(desc(A0, SW64), desc(off(B0, r + 8 * (u % 3), 64, u & 1), SW64), 128, N, 0); });
    }
  // 3. cycling shifts 0..10 (k=11, dil=1) like the conv loop
  for (int N : {32, 64, 128}) {
    snprintf(nm, sizeof nm, "A(sw64) shift cycling 0..10, M=128 N=%d", N);
    add(nm, [&](int u) { return mk(desc(off(A0, u % 11, 64, u & 1), SW64), desc(B0, SW64), 128, N, 0); });
  }
  for (int N : {64, 128, 256}) {
    snprintf(nm, sizeof nm, "B(sw64) shift cycling 0..10, M=128 N=%d", N);
    add(nm, [&](int u) { return mk(desc(A0, SW64), desc(off(B0, u % 11, 64, u & 1), SW64), 128, N, 0); });
  }
  // 4. M = 64
  for (int N : {64, 128, 256}) {
    snprintf(nm, sizeof nm, "M=64 N=%d, nothing shifted", N);
    add(nm, [&](int u) { return mk(desc(A0, SW64), desc(B0, SW64), 64, N, 0); });
    snprintf(nm, sizeof nm, "M=64 N=%d, B shift cycling", N);
    add(nm, [&](int u) { return mk(desc(A0, SW64), desc(off(B0, u % 11, 64, u & 1), SW64), 64, N, 0); });
  }
  // 5. A in tensor memory (weights resident in TMEM columns 256..), B shifted
  for (int N : {64, 128, 256}) {
    snprintf(nm, sizeof nm, "A in TMEM, B(sw64) shift cycling, M=128 N=%d", N);
    add(nm, [&](int u) {
      Step s = mk(0, desc(off(B0, u % 11, 64, u & 1), SW64), 128, N, 0);
      s.a_in_tmem = 1, s.a_tmem_col = 256 + 8 * (u % 22);
      return s;
    });
    snprintf(nm, sizeof nm, "A in TMEM, B unshifted, M=128 N=%d", N);
    add(nm, [&](int u) {
      Step s = mk(0, desc(B0, SW64), 128, N, 0);
      s.a_in_tmem = 1, s.a_tmem_col = 256 + 8 * (u % 22);
      return s;
    });
  }
  // 6. 32 B rows (one k-group per row, SW32): shifted A / shifted B
  for (int N : {32, 64}) {
    snprintf(nm, sizeof nm, "A(sw32) shift cycling 0..10, M=128 N=%d", N);
    add(nm, [&](int u) { return mk(desc(A0 + (u % 11) * 32 + (u & 1) * 8192, SW32), desc(B0, SW64), 128, N, 0); });
  }
  for (int r : {0, 1, 4}) {
    snprintf(nm, sizeof nm, "A(sw32) shift r=%d, M=128 N=32", r);
    add(nm, [&](int u) { return mk(desc(A0 + (r + 8 * (u % 3)) * 32 + (u & 1) * 8192, SW32), desc(B0, SW64), 128, 32, 0); });
  }
  for (int N : {128, 256}) {
    snprintf(nm, sizeof nm, "B(sw32) shift cycling 0..10, M=128 N=%d", N);
    add(nm, [&](int u) { return mk(desc(A0, SW64), desc(B0 + (u % 11) * 32 + (u & 1) * 16384, SW32), 128, N, 0); });
  }
  // 7. 128 B rows (SW128, C = 64 per row)
  for (int N : {64, 128}) {
    snprintf(nm, sizeof nm, "A(sw128) shift cycling 0..10, M=128 N=%d", N);
    add(nm, [&](int u) { return mk(desc(off(A0, u % 11, 128, u & 3), SW128), desc(B0, SW64), 128, N, 0); });
  }
  snprintf(nm, sizeof nm, "B(sw128) shift cycling 0..10, M=128 N=256");
  add(nm, [&](int u) { return mk(desc(A0, SW64), desc(off(B0, u % 11, 128, u & 3), SW128), 128, 256, 0); });
  // 8. candidate schedules per tap (time on N): [wh;wl] x xh (M=128) then wh x xl (M=64), N = 256 / 128
  for (int N : {128, 256}) {
    snprintf(nm, sizeof nm, "swapped C=64: M=128 (B=xh shifted) + M=64 (B=xl shifted), N=%d", N);
    add(nm, [&](int u) {
      const int tap = (u >> 1) % 11;
      return (u & 1) == 0 ? mk(desc(A0, SW64), desc(off(B0, tap, 64, 0), SW64), 128, N, 0)
                          : mk(desc(A0 + 16384, SW64), desc(off(B0 + 32768, tap, 64, 0), SW64), 64, N, N == 256 ? 256 : 128);
    });
    snprintf(nm, sizeof nm, "swapped, A in TMEM: M=128 (B=xh shifted) + M=128 (B=xl shifted), N=%d", N);
    add(nm, [&](int u) {
      const int tap = (u >> 1) % 11;
      Step s = (u & 1) == 0 ? mk(0, desc(off(B0, tap, 64, 0), SW64), 128, N, 0)
                            : mk(0, desc(off(B0 + 32768, tap, 64, 0), SW64), 128, N, N == 256 ? 256 : 128);
      s.a_in_tmem = 1, s.a_tmem_col = (N == 256 ? 0 : 256) + 8 * (u % 22);
      if (N == 256) s.dcol = 0, s.a_tmem_col = 256 + 8 * (u % 22);  // one accumulator only: timing probe
      return s;
    });
  }
  // 9. today's pairs for reference
  for (int N : {32, 64, 128}) {
    snprintf(nm, sizeof nm, "today: A=xh shifted N'=%d, A=xl shifted N=%d", 2 * N, N);
    add(nm, [&](int u) {
      const int tap = (u >> 1) % 11;
      return (u & 1) == 0 ? mk(desc(off(A0, tap, 64, 0), SW64), desc(B0, SW64), 128, 2 * N, 0)
                          : mk(desc(off(A0 + 32768, tap, 64, 0), SW64), desc(B0, SW64), 128, N, N);
    });
  }

  for (const Pat& p : pats) {
    CK(cudaMemcpyToSymbol(c_steps, p.s, sizeof(p.s)));
    for (int rep = 0; rep < 2; ++rep) {
      if (p.s[0].a_in_tmem) mma_kernel<true><<<sms, 128, smem>>>(iters, d);
      else mma_kernel<false><<<sms, 128, smem>>>(iters, d);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h.data(), d, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mn = h[0], mx = h[0];
    double sum = 0;
    for (int i = 0; i < sms; ++i) mn = h[i] < mn ? h[i] : mn, mx = h[i] > mx ? h[i] : mx, sum += (double)h[i];
    printf("%-78s cycles/MMA  min %7.1f  avg %7.1f  max %7.1f\n", p.name.c_str(), (double)mn / iters, sum / sms / iters,
           (double)mx / iters);
  }
  for (int x16 : {0, 1})
    for (int nw : {4, 8, 16}) {
      const int it2 = 4096;
      for (int rep = 0; rep < 2; ++rep) {
        ldtm_kernel<<<sms, nw * 32, 0>>>(it2, x16, d);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h.data(), d, sms * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
      printf("tcgen05.ld 32x32b.%s + wait::ld, %2d warps: %.1f B/cycle per SM (%.1f cycles per 4 KB warp-load)\n",
             x16 ? "x16 x2" : "x32", nw, (double)nw * it2 * 4096.0 / (double)mx, (double)mx / it2);
    }
  return 0;
}
