// Stride-1 Conv1d (dilated, any tap count) as an implicit GEMM on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM), at fp32-class accuracy through a three-product fp16 split:
//
//     x = xh + xl,  w*s = wh + wl   (xh = fp16(x), xl = fp16(x - xh); s = per-layer power of two)
//     y = (xh.wh + xh.wl + xl.wh) / s          products exact in the fp32 accumulator; the dropped
//                                               xl.wl term is <= 2^-22 relative
//
// GEMM mapping (one CTA = one [128*NSUB time] x [N output channels] tile of one utterance):
//     M = time (128 rows per MMA), N = output channels, K = input channels x taps
//     A = activations, K-major, NO swizzle:  A_s[kgroup(8 ch)][row = time][16 B]
//         -> a conv tap is a ROW shift, i.e. a 16 B-granular change of the descriptor start
//            address (SBO = 128 B makes 8-row groups contiguous, so any row offset is legal);
//            one staged tile with a (K-1)*dil halo serves every tap -- no im2col, no re-staging.
//     B = weights per (chunk, tap), K-major, no swizzle: B_s[kgroup][n][16 B], pre-split/pre-packed
//         at load time in exactly this image, so one cp.async.bulk per pipeline stage fills it.
//     D = fp32 in TMEM: per time sub-tile a main accumulator (xh.wh) and (optionally separate) a
//         cross accumulator (xh.wl + xl.wh), summed in the epilogue.
//
// Warp roles (192 threads): warp 0 = TMEM allocator + weight bulk-copy producer (1 lane),
// warp 1 = barrier init + MMA issuer (1 lane), warps 2..5 = activation producers (global fp32 ->
// leaky_relu/mask -> fp16 hi/lo -> smem) and, once the accumulators are committed, the epilogue
// (tcgen05.ld -> bias / residual / running sum / mask / tanh -> coalesced global stores).
// Pipelines: A ring (2 stages, mbarrier full/empty), weight ring (NW stages, expect_tx / commit),
// accumulator-full barrier.  Overlap of epilogue and mainloop comes from 2 co-resident CTAs per SM
// when the tile's smem/TMEM footprint allows it.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "svk_kernels.cuh"

namespace svk {

namespace {

constexpr int KC = TC_KC;  // input channels per A chunk
constexpr int KG = KC / 8;  // 16-byte k-groups per chunk
constexpr int NA = 2;       // A ring depth (max)
constexpr int MAXNW = 6;    // weight ring depth limit
constexpr int THREADS = 192;
constexpr int PRODUCERS = 128;

struct __align__(8) SmemHeader {
  uint64_t a_full[NA], a_empty[NA];
  uint64_t w_full[MAXNW], w_empty[MAXNW];
  uint64_t acc_full;
  uint32_t tmem_base;
  uint32_t pad;
};
constexpr int HEADER_BYTES = 256;
static_assert(sizeof(SmemHeader) <= HEADER_BYTES, "header");

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
// Spin on the barrier's phase parity.  A bounded spin (~2 s) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
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

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
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

// fp32 -> (hi, lo) fp16 pair for two values; returns packed half2 words.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ------------------------------------------------------------------------------------- kernel
__global__ void __launch_bounds__(THREADS, 2) conv_tc_kernel(const ConvTcArgs ta) {
  extern __shared__ __align__(128) uint8_t smem[];
  const ConvArgs& a = ta.c;
  SmemHeader* hdr = reinterpret_cast<SmemHeader*>(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = ta.N, nsub = ta.nsub, nw = ta.nw, na = ta.na;
  const int nacc = ta.sep_cross ? 2 : 1;
  const int K = a.K, dil = a.dil;
  const int rows = ta.rows;                      // staged time rows per chunk (multiple of 8)
  const uint32_t a_plane = (uint32_t)rows * 16u;  // one k-group plane of A
  const uint32_t a_stage = a_plane * KG * 2u;     // hi + lo
  const uint32_t w_plane = (uint32_t)N * 16u;
  const uint32_t w_stage = w_plane * KG * 2u;
  uint8_t* a_smem = smem + HEADER_BYTES;
  uint8_t* w_smem = a_smem + na * a_stage;
  const int t0 = blockIdx.x * (128 * nsub), ntile = blockIdx.y, b = blockIdx.z;
  const int nchunks = a.Cin / KC;
  const uint32_t tmem_cols = ta.tmem_cols;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < na; ++i) mbar_init(&hdr->a_full[i], PRODUCERS), mbar_init(&hdr->a_empty[i], 1);
    for (int i = 0; i < nw; ++i) mbar_init(&hdr->w_full[i], 1), mbar_init(&hdr->w_empty[i], 1);
    mbar_init(&hdr->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&hdr->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = hdr->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------ weight producer: one bulk copy per (chunk, tap)
    if (lane == 0) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(ta.wtc) + (size_t)ntile * nchunks * K * w_stage;
      const int total = nchunks * K;
      int st = 0;
      uint32_t ph = 0;
      for (int it = 0; it < total; ++it) {
        mbar_wait(&hdr->w_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&hdr->w_full[st], w_stage);
        bulk_g2s(w_smem + (size_t)st * w_stage, src + (size_t)it * w_stage, w_stage, &hdr->w_full[st]);
        if (++st == nw) st = 0, ph ^= 1;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);  // f16 x f16 -> f32, K-major
      uint32_t started = 0;  // bit (sub * nacc + acc): accumulator already holds a partial sum
      int wst = 0;
      uint32_t wph = 0;
      for (int ch = 0; ch < nchunks; ++ch) {
        const int as = ch % na;
        mbar_wait(&hdr->a_full[as], (ch / na) & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(a_smem + (size_t)as * a_stage), a_lo = a_hi + a_plane * KG;
        for (int j = 0; j < K; ++j) {
          mbar_wait(&hdr->w_full[wst], wph);
          tc_fence_after();
          const uint32_t b_hi = smem_u32(w_smem + (size_t)wst * w_stage), b_lo = b_hi + w_plane * KG;
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks) {
            const uint64_t dbh = make_desc(b_hi + 2 * ks * w_plane, w_plane, 128);
            const uint64_t dbl = make_desc(b_lo + 2 * ks * w_plane, w_plane, 128);
            for (int sub = 0; sub < nsub; ++sub) {
              const uint32_t roff = (uint32_t)(sub * 128 + j * dil) * 16u + 2 * ks * a_plane;
              const uint64_t dah = make_desc(a_hi + roff, a_plane, 128);
              const uint64_t dal = make_desc(a_lo + roff, a_plane, 128);
              const int im = sub * nacc, ic = sub * nacc + nacc - 1;
              const uint32_t dm = tmem + (uint32_t)(im * N), dc = tmem + (uint32_t)(ic * N);
              umma_f16(dc, dal, dbh, idesc, (started >> ic) & 1u);
              started |= 1u << ic;
              umma_f16(dc, dah, dbl, idesc, 1u);
              umma_f16(dm, dah, dbh, idesc, (started >> im) & 1u);
              started |= 1u << im;
            }
          }
          umma_commit(&hdr->w_empty[wst]);
          if (++wst == nw) wst = 0, wph ^= 1;
        }
        umma_commit(&hdr->a_empty[as]);
      }
      umma_commit(&hdr->acc_full);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ activation producers (128 threads)
    // One task = one time row x all KC channels of the chunk: 32 independent coalesced loads in
    // flight per thread, then leaky_relu / mask / fp16 hi-lo split and 2 x 4 conflict-free 16 B stores.
    const int rp = tid - 64;
    const float slope = a.pre_slope;
    const float* mrow = a.in_mask ? a.in_mask + (size_t)b * a.mask_stride : nullptr;
    for (int ch = 0; ch < nchunks; ++ch) {
      const int as = ch % na;
      mbar_wait(&hdr->a_empty[as], ((ch / na) & 1) ^ 1);
      uint4* Ahi = reinterpret_cast<uint4*>(a_smem + (size_t)as * a_stage);
      uint4* Alo = Ahi + KG * rows;
      const float* xb = a.x + ((size_t)b * a.x_C + a.x_ch_off + ch * KC) * a.x_stride;
      for (int r = rp; r < rows; r += PRODUCERS) {
        const int t = t0 - a.pad + r;
        const bool ok = t >= 0 && t < a.Lin;
        float v[KC];
#pragma unroll
        for (int c = 0; c < KC; ++c) v[c] = ok ? __ldg(xb + (size_t)c * a.x_stride + t) : 0.f;
        const float mk = (ok && mrow) ? __ldg(mrow + t) : 1.0f;
#pragma unroll
        for (int c = 0; c < KC; ++c) {
          float q = v[c];
          q = q > 0.f ? q : q * slope;
          v[c] = mrow ? q * mk : q;
        }
#pragma unroll
        for (int kg = 0; kg < KG; ++kg) {
          uint4 h, l;
          split2(v[kg * 8 + 0], v[kg * 8 + 1], h.x, l.x);
          split2(v[kg * 8 + 2], v[kg * 8 + 3], h.y, l.y);
          split2(v[kg * 8 + 4], v[kg * 8 + 5], h.z, l.z);
          split2(v[kg * 8 + 6], v[kg * 8 + 7], h.w, l.w);
          Ahi[kg * rows + r] = h;
          Alo[kg * rows + r] = l;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&hdr->a_full[as]);
    }

    // ------------------------------------------------ epilogue: TMEM -> registers -> global
    mbar_wait(&hdr->acc_full, 0);
    tc_fence_after();
    const int q4 = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q4 * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(q4 * 32) << 16);
    const float unscale = ta.unscale;
    const float* omask = a.out_mask ? a.out_mask + (size_t)b * a.mask_stride : nullptr;
    const int o_tile = ntile * N;
    for (int sub = 0; sub < nsub; ++sub) {
      const int t = t0 + sub * 128 + row;
      const uint32_t tsub = tlane + (uint32_t)(sub * nacc * N);

      if (a.mode == MODE_STORE) {
        // y = ((acc/s + bias) + (res + acc_in)) / post_div * mask, tanh; res / acc_in of the NEXT 16
        // channels are in flight while the current 16 are finished (same element is read then
        // written by the same thread only, so in-place operation is safe).
        const bool tin = t < a.Lout;
        const float mv = (omask && tin) ? omask[t] : 1.0f;
        const float* res_c = nullptr;
        const float* acc_c = nullptr;
        float* y_c = nullptr;
        ptrdiff_t step_c = 0;
        int um_c = 0, nv_c = 0;
        float r1[16];
        auto chunk_params = [&](int n0, const float*& res, const float*& accin, float*& y, ptrdiff_t& step, int& um,
                                int& nvalid) {
          const int o0 = o_tile + n0;
          const bool s1 = o0 >= a.split;
          const int rel0 = s1 ? o0 - a.split : o0;
          const int dC = s1 ? a.e[1].C : a.e[0].C, dch = s1 ? a.e[1].ch_off : a.e[0].ch_off;
          const int dsg = s1 ? a.e[1].ch_sign : a.e[0].ch_sign;
          const size_t off = ((size_t)b * dC + dch + dsg * rel0) * a.y_stride + t;
          const float* rs = s1 ? a.e[1].res : a.e[0].res;
          const float* ai = s1 ? a.e[1].acc_in : a.e[0].acc_in;
          res = rs ? rs + off : nullptr;
          accin = ai ? ai + off : nullptr;
          y = (s1 ? a.e[1].y : a.e[0].y) + off;
          step = (ptrdiff_t)dsg * a.y_stride;
          um = s1 ? a.e[1].use_mask : a.e[0].use_mask;
          nvalid = tin ? min(16, a.Cout - o0) : 0;
        };
        auto load16 = [&](const float* res, const float* accin, ptrdiff_t step, int nvalid, float (&q1)[16]) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float u1 = (res && e < nvalid) ? res[e * step] : 0.f;
            const float u2 = (accin && e < nvalid) ? accin[e * step] : 0.f;
            q1[e] = u1 + u2;
          }
        };
        chunk_params(0, res_c, acc_c, y_c, step_c, um_c, nv_c);
        load16(res_c, acc_c, step_c, nv_c, r1);
        for (int n0 = 0; n0 < N; n0 += 16) {
          const float* res_n = nullptr;
          const float* acc_n = nullptr;
          float* y_n = nullptr;
          ptrdiff_t step_n = 0;
          int um_n = 0, nv_n = 0;
          float p1[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) p1[e] = 0.f;
          if (n0 + 16 < N) {
            chunk_params(n0 + 16, res_n, acc_n, y_n, step_n, um_n, nv_n);
            load16(res_n, acc_n, step_n, nv_n, p1);
          }
          uint32_t m[16], c[16];
          tmem_ld16(tsub + (uint32_t)n0, m);
          if (nacc == 2) tmem_ld16(tsub + (uint32_t)(N + n0), c);
          tmem_wait_ld();
          const float* bptr = a.bias + o_tile + n0;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            if (e < nv_c) {
              float v = __uint_as_float(m[e]);
              if (nacc == 2) v += __uint_as_float(c[e]);
              v = fmaf(v, unscale, __ldg(bptr + e));
              v += r1[e];
              if (a.post_div != 1.0f) v = v / a.post_div;
              if (um_c && omask) v *= mv;
              if (a.act_tanh) v = tanhf(v);
              y_c[e * step_c] = v;
            }
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) r1[e] = p1[e];
          res_c = res_n, acc_c = acc_n, y_c = y_n, step_c = step_n, um_c = um_n, nv_c = nv_n;
        }
      } else if (a.mode == MODE_GATE) {
        // columns [0, N/2) hold the tanh half, [N/2, N) the sigmoid half of channels
        // ntile*N/2 + [0, N/2) (commons.py:100-107); bias is in the same virtual order.
        const bool tin = t < a.Lout;
        const int half = N >> 1;
        const float* bptr = a.bias + o_tile;
        float* ybase = a.e[0].y + ((size_t)b * a.e[0].C + a.e[0].ch_off + ntile * half) * a.y_stride + t;
        for (int n0 = 0; n0 < half; n0 += 16) {
          uint32_t mt[16], ms[16], ct[16], cs[16];
          tmem_ld16(tsub + (uint32_t)n0, mt);
          tmem_ld16(tsub + (uint32_t)(half + n0), ms);
          if (nacc == 2) {
            tmem_ld16(tsub + (uint32_t)(N + n0), ct);
            tmem_ld16(tsub + (uint32_t)(N + half + n0), cs);
          }
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int cidx = ntile * half + n0 + e;
            if (tin && cidx < (a.Cout >> 1)) {
              float vt = __uint_as_float(mt[e]), vs = __uint_as_float(ms[e]);
              if (nacc == 2) vt += __uint_as_float(ct[e]), vs += __uint_as_float(cs[e]);
              vt = fmaf(vt, unscale, __ldg(bptr + n0 + e));
              vs = fmaf(vs, unscale, __ldg(bptr + half + n0 + e));
              ybase[(size_t)(n0 + e) * a.y_stride] = tanhf(vt) * sigmoidf_(vs);
            }
          }
        }
      } else {
        // MODE_SHUFFLE: virtual channel o' = co*s + r of time row q lands at y[co, s*q + r - p]
        // (ConvTranspose1d polyphase form).
        const int s = a.shuf_s;
        const bool qin = t < a.Lout;
        float* ybase = a.e[0].y + ((size_t)b * a.e[0].C + a.e[0].ch_off) * a.y_stride;
        for (int n0 = 0; n0 < N; n0 += 16) {
          uint32_t m[16], c[16];
          tmem_ld16(tsub + (uint32_t)n0, m);
          if (nacc == 2) tmem_ld16(tsub + (uint32_t)(N + n0), c);
          tmem_wait_ld();
          const int o0 = o_tile + n0;
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float q = __uint_as_float(m[e]);
            if (nacc == 2) q += __uint_as_float(c[e]);
            v[e] = fmaf(q, unscale, (o0 + e < a.Cout) ? __ldg(a.bias + o0 + e) : 0.f);
          }
          if (!qin) continue;
          if (s == 8 && (a.shuf_p & 3) == 0) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int co = (o0 >> 3) + g;
              if (o0 + 8 * g >= a.Cout) break;
              float* yrow = ybase + (size_t)co * a.y_stride;
              const int tt = 8 * t - a.shuf_p;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int th = tt + 4 * h;
                if (th >= 0 && th + 3 < a.shuf_Lout) {
                  *reinterpret_cast<float4*>(yrow + th) =
                      make_float4(v[8 * g + 4 * h], v[8 * g + 4 * h + 1], v[8 * g + 4 * h + 2], v[8 * g + 4 * h + 3]);
                } else {
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    if (th + i >= 0 && th + i < a.shuf_Lout) yrow[th + i] = v[8 * g + 4 * h + i];
                }
              }
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int o = o0 + e;
              const int co = o / s, r = o - co * s;
              const int tt = s * t + r - a.shuf_p;
              if (o < a.Cout && tt >= 0 && tt < a.shuf_Lout) ybase[(size_t)co * a.y_stride + tt] = v[e];
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace

// ------------------------------------------------------------------------------- host helpers
int conv_tc_rows(int K, int dil, int nsub) { return (128 * nsub + (K - 1) * dil + 7) & ~7; }

size_t conv_tc_smem_bytes(int N, int K, int dil, int nsub, int nw, int na) {
  return HEADER_BYTES + (size_t)na * conv_tc_rows(K, dil, nsub) * 16 * KG * 2 + (size_t)nw * N * 16 * KG * 2;
}

size_t conv_tc_packed_halves(int Cin, int Cout, int K, int N) {
  const int ntiles = (Cout + N - 1) / N, nchunks = Cin / KC;
  return (size_t)ntiles * nchunks * K * N * KC * 2;
}

// Power-of-two scale that brings max|w| into [512, 1024): fp16 hi keeps 11 bits, lo another 11.
float conv_tc_weight_scale(const float* w, size_t n) {
  float mx = 0.f;
  for (size_t i = 0; i < n; ++i) {
    const float v = w[i] < 0 ? -w[i] : w[i];
    if (v > mx) mx = v;
  }
  if (!(mx > 0.f) || mx != mx) return 1.0f;
  int e = 0;
  frexpf(mx, &e);  // mx = f * 2^e, f in [0.5, 1)
  return ldexpf(1.0f, 10 - e);
}

// wv(o, c, j) is the logical fp32 weight; image = [ntile][chunk][tap][hi|lo][kgroup][n][8].
void conv_tc_pack(const float* w_ock, int Cout, int Cin, int K, int N, float scale, uint16_t* out) {
  const int ntiles = (Cout + N - 1) / N, nchunks = Cin / KC;
  size_t idx = 0;
  for (int nt = 0; nt < ntiles; ++nt)
    for (int ch = 0; ch < nchunks; ++ch)
      for (int j = 0; j < K; ++j)
        for (int part = 0; part < 2; ++part)
          for (int kg = 0; kg < KG; ++kg)
            for (int n = 0; n < N; ++n)
              for (int e = 0; e < 8; ++e) {
                const int o = nt * N + n, c = ch * KC + kg * 8 + e;
                float v = 0.f;
                if (o < Cout) v = w_ock[((size_t)o * Cin + c) * K + j] * scale;
                const __half h = __float2half_rn(v);
                const __half l = __float2half_rn(v - __half2float(h));
                const __half pick = part ? l : h;
                out[idx++] = *reinterpret_cast<const uint16_t*>(&pick);
              }
}

cudaError_t launch_conv_tc(const ConvTcArgs& ta_in, cudaStream_t stream) {
  ConvTcArgs ta = ta_in;
  const ConvArgs& a = ta.c;
  if (a.Cin % KC != 0 || ta.N % 16 != 0 || ta.N < 16 || ta.N > 256) return cudaErrorInvalidValue;
  if (ta.nsub < 1 || ta.nsub > 2 || ta.nw < 2 || ta.nw > MAXNW) return cudaErrorInvalidValue;
  if (a.split != (1 << 30) && (a.split % 16 != 0 || a.mode != MODE_STORE)) return cudaErrorInvalidValue;
  if (a.mode == MODE_GATE && (ta.N % 32 != 0 || a.Cout % 2)) return cudaErrorInvalidValue;
  ta.na = (a.Cin / KC) > 1 ? NA : 1;
  const int nacc = ta.sep_cross ? 2 : 1;
  const int need = ta.nsub * nacc * ta.N;
  if (need > 512) return cudaErrorInvalidValue;
  int cols = 32;
  while (cols < need) cols <<= 1;
  ta.tmem_cols = cols;
  ta.rows = conv_tc_rows(a.K, a.dil, ta.nsub);
  const size_t smem = conv_tc_smem_bytes(ta.N, a.K, a.dil, ta.nsub, ta.nw, ta.na);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  static size_t configured[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[dev & 63] = smem;
  }
  dim3 grid((a.Lout + 128 * ta.nsub - 1) / (128 * ta.nsub), (a.Cout + ta.N - 1) / ta.N, a.B);
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return cudaSuccess;
  conv_tc_kernel<<<grid, THREADS, smem, stream>>>(ta);
  return cudaGetLastError();
}

}  // namespace svk
