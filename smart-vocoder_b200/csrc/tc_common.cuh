// Device-side building blocks shared by the tcgen05 kernels of libsvk (conv_tc.cu, conv_tc_pair.cu):
// mbarrier / TMA / tcgen05 PTX wrappers, the fp16 hi-lo split, operand-tile constants.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "svk_kernels.cuh"

extern "C" int svk_g_pdl_enabled;  // defined in svk_model.cu

namespace svk {
namespace {

constexpr int KC = TC_KC;   // input channels per A chunk
constexpr int KG = KC / 8;  // 16-byte k-groups per chunk

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// mbarrier.try_wait suspends the warp until the phase completes or a time limit expires.  With the default (system)
// limit a waiting warp comes back every ~100 cycles and re-issues the loop -- eight epilogue warps of a pair kernel spent
// 6 M such retries per launch.  SVK_MBAR_HINT_NS > 0 passes a suspend-time hint instead (the warp still wakes as soon as
// the phase completes): fewer wasted issue slots next to the warps that have work.
#ifndef SVK_MBAR_HINT_NS
#define SVK_MBAR_HINT_NS 0
#endif
// Spin on the barrier's phase parity.  A bounded spin (~2 s) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
#if SVK_MBAR_HINT_NS > 0
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"((uint32_t)SVK_MBAR_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
#endif
    if (done) return;
    if ((spin & 0xFFFF) == 0xFFFF) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Programmatic dependent launch: the next kernel of the stream may be scheduled onto SMs this grid has left
// (its prologue -- barrier init, TMEM allocation, weight loads -- overlaps this grid's tail); a dependent grid
// blocks in griddep_wait() until the grid before it has completed and its writes are visible.  Both are no-ops
// for a launch without the programmatic-stream-serialization attribute.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Launch with programmatic stream serialization unless $SVK_PDL=0 (A/B measurements) or svk__set_pdl(0) (graph capture
// falls back to plain edges when the driver refuses programmatic ones; svk_pipeline.cu).
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int threads, size_t smem, cudaStream_t stream, Args... args) {
  static const bool pdl = [] {
    const char* e = getenv("SVK_PDL");
    return !(e && e[0] == '0');
  }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3((unsigned)threads), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = (pdl && svk_g_pdl_enabled) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 4-D tiled TMA load (cp.async.bulk.tensor): box -> smem, completion counted on the mbarrier.
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16 inputs, fp32 accumulate), M=128, K=16.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// Same MMA with descriptors given as (low word, shared high word): no 64-bit arithmetic per issue.
__device__ __forceinline__ void umma_f16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_u32(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// Lean wait for the MMA warp: one try_wait on the fast path; bounded spin, then trap.
__device__ __forceinline__ void mbar_wait_u32(uint32_t addr, uint32_t parity) {
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
#if SVK_MBAR_HINT_NS > 0
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"((uint32_t)SVK_MBAR_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
#endif
    if (done) return;
    if (spin > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor, sm_100 version 1):
// [0,14) start>>4 | [16,30) LBO>>4 (stride between the two 8-element k-groups of one MMA)
// | [32,46) SBO>>4 (stride between 8-row groups) | [46,48) version=1 | [61,64) layout=0.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46);
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

// WN gate of the tensor-core engines: tanh(a) * sigmoid(b) (commons.py:100-107) = (1 - 2 / (1 + e^{2a})) / (1 + e^{-b}) on the
// SFU (ex2.approx / rcp.approx: 13 instructions instead of ~60 for tanhf + expf + an IEEE division).  ncu on the fused WN
// layer showed the gate epilogue (1700 warp instructions per N-tile at 0.14 IPC per warp) taking longer than the MMAs it
// should hide under.  Absolute error <= 2e-7 on a value in (-1, 1) (1 - 2/(1 + e^{2a}) rounds twice near 1), limits exact:
// e^{2a} = inf -> 1, = 0 -> -1; NaN propagates.  The fp32 FFMA engine keeps tanhf / expf.
__device__ __forceinline__ float gate_tanh_sigmoid(float a, float b) {
  const float ea = __expf(2.0f * a), eb = __expf(-b);
  const float th = 1.0f - __fdividef(2.0f, 1.0f + ea);
  return __fdividef(th, 1.0f + eb);
}

// One 32 B global store (STG.256, sm_100): a whole sector in one request.  Two 16 B stores to the halves of a sector
// reach L2 as two partial-sector writes; on the operand-image rows (32 B per thread and plane) that costs 1.5-2x the
// time of the same bytes written as full sectors.  `p` must be 32 B-aligned.
__device__ __forceinline__ void st_global_v8(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

// The lo half of every fp16 hi/lo pair -- activations and weights alike -- is stored scaled by 2^11: lo = fp16((x - hi) * 2048).
// Unscaled, x - hi (<= 2^-11 |x|) falls into fp16's subnormals for |x| < 0.125 and the pair keeps fewer than 22 bits;
// scaled, the pair is exact to 22 bits for every |x| in [6.1e-5, 65504], the range of hi itself.  Both cross products
// (xh.wl' and xl'.wh) then carry the same factor, their accumulator is folded in as main + cross * 2^-11 (one FFMA in
// place of the FADD), and nothing else changes: power-of-two scaling is exact.
constexpr float LO_SCALE = 2048.0f, LO_INV = 1.0f / 2048.0f;

// fp32 -> (hi, lo) fp16 pair for two values; returns packed half2 words.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((a - hf.x) * LO_SCALE, (b - hf.y) * LO_SCALE);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// 8 fp32 -> 8 bf16 (round to nearest even), one 16 B row of a single-plane operand tile
__device__ __forceinline__ uint4 pack_bf16x8(const float* v) {
  uint4 r;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(v[0], v[1]);
  r.x = *reinterpret_cast<const uint32_t*>(&t);
  t = __floats2bfloat162_rn(v[2], v[3]);
  r.y = *reinterpret_cast<const uint32_t*>(&t);
  t = __floats2bfloat162_rn(v[4], v[5]);
  r.z = *reinterpret_cast<const uint32_t*>(&t);
  t = __floats2bfloat162_rn(v[6], v[7]);
  r.w = *reinterpret_cast<const uint32_t*>(&t);
  return r;
}

__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv& f) {
  const uint32_t t = __umulhi(f.m, n);
  return (t + ((n - t) >> f.sh1)) >> f.sh2;
}

inline FastDiv make_fast_div(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  f.m = (uint32_t)((((1ull << l) - d) << 32) / d + 1);
  f.sh1 = l < 1 ? l : 1;
  f.sh2 = l > 0 ? l - 1 : 0;
  return f;
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (libsvk does not link libcuda).
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

// Tensor map of an operand image [planes][B*C/32][L][32]; box = (32 ch, rows, 1, planes), 64 B swizzle (conv_tc.cu).
cudaError_t tc_make_image_map(const uint16_t* img, int B, int C, int L, int rows, int planes, CUtensorMap* map);

}  // namespace svk
