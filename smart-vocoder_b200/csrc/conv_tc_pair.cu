// One launch = one conv pair of a HiFi-GAN ResBlock1 (modules.py:211-220) on the narrow decoder stages:
//
//     xt = conv1(leaky_relu(x), dilation d)      xt never leaves the SM
//     y  = conv2(leaky_relu(xt)) + x             (+ running sum, / 3, fp32 and/or operand-image output)
//
// The narrow stages (C <= 64; 537 MB tensors at 16 x 1024 frames) are HBM-bound layer by layer: unfused,
// a pair moves x-image in, xt-image out, xt-image in, residual in, y out.  Fused, it moves x-image in,
// residual in (an L2 hit: the same lines the TMA just fetched) and y out.
//
// Same machinery as conv_tc.cu (fp16 hi/lo three-product scheme, 64 B-swizzled A tiles where a tap is a
// row shift, TMA-fed operand images, persistent CTAs, warp roles) with two accumulator sets in TMEM:
//   item = 128 - (K-1) output steps of one utterance.  conv1 computes the 128 xt rows conv2 needs (halo
//   recomputed per item), its epilogue turns them into conv2's A tile IN SHARED MEMORY (bias, zero outside
//   [0, L) = conv2's padding, leaky_relu, hi/lo split, swizzled stores, fence.proxy.async), conv2 runs
//   on that tile, its epilogue is the usual residual / running-sum / image store.
//   The MMA warp issues conv1(i+1) before conv2(i); two role-specialised groups of four epilogue warps run epi1
//   (conv1's accumulators -> xt tile) and epi2 (conv2's accumulators -> output) over all items independently, so
//   two items are in flight: epi1(i+1) under conv2(i)'s MMAs, epi2(i) under conv1(i+2)'s.
#include <string.h>

#include "svk_kernels.cuh"
#include "tc_common.cuh"

namespace svk {

namespace {

constexpr int P_NA_MAX = 8, P_MAXNW = 32;
// Items in flight: accumulator sets of both convs and xt tiles come in `ns` stages -- 2 at C = 64 (8 N = 512 TMEM columns),
// 4 at C = 32, where two stages left half of TMEM unused and the conv1 -> epi1 -> conv2 -> epi2 chain (~4000 cycles per item
// against 1336 cycles of MMAs for a 3-tap pair) bounded the kernel: with two items in flight it ran at neither its MMA
// nor its HBM floor.
constexpr int P_MAX_STAGES = 4;
// warps 4..11 run epi1 (conv1 accumulators -> xt tile in shared memory), the next eight epi2 (conv2 accumulators +
// residual -> fp32 / image in HBM -- with four warps it reached 3.7 TB/s where conv_tc's eight-warp epilogue reaches 5);
// in both roles the two warps of a TMEM lane quarter take alternate 16-column jobs.
// epi1 had four warps (-DSVK_PAIR_EPI1_WARPS=4, 512 threads) until ncu (profiles/r2_ncu_pair_c32_k7.txt) showed them busy
// 75 % of the kernel at one instruction per ~6 cycles -- a single latency-bound warp per scheduler, ~3000 cycles per item
// at C = 32, more than both convs' MMAs of a 3-tap pair.  Eight: C = 32 pairs -5..-9 %, C = 64 k = 3 pairs +3 %
// (profiles/r2_pair_epi1_warps_ab.txt); the pairs stay bound by their HBM path (same capture: epi2 stalled on the
// registers of stores the memory pipe has not accepted yet).
#ifndef SVK_PAIR_EPI1_WARPS
#define SVK_PAIR_EPI1_WARPS 8
#endif
constexpr int P_EPI1_WARPS = SVK_PAIR_EPI1_WARPS, P_EPI2_WARPS = 8;
static_assert(P_EPI1_WARPS == 4 || P_EPI1_WARPS == 8, "epi1: one or two warps per TMEM lane quarter");
// Register budget (P_EPI1_WARPS == 8: 640 threads, 96 registers each at launch): the four producer / issuer warps and the
// epi1 warps give registers back (setmaxnreg.dec) and the epi2 warps, whose operand prefetch buffers need them, take
// more (setmaxnreg.inc); each role's code sits in a branch dominated by its setmaxnreg, which is what ptxas allocates by.
constexpr int P_REGS_PRODUCER = 56, P_REGS_EPI1 = 80, P_REGS_EPI2 = 128;  // no spills in any role
constexpr int P_REGS_LAUNCH = 96;  // 65536 / 640 rounded down to the allocation unit of 8: what ptxas gives the kernel
static_assert(P_EPI1_WARPS != 8 || 128 * P_REGS_PRODUCER + 256 * P_REGS_EPI1 + 256 * P_REGS_EPI2 <= 640 * P_REGS_LAUNCH,
              "setmaxnreg.inc only draws on registers the CTA's own warps have released (a larger sum deadlocks)");
constexpr int P_THREADS = 128 + 32 * (P_EPI1_WARPS + P_EPI2_WARPS);
constexpr int P_EPI1_THREADS = 32 * P_EPI1_WARPS, P_EPI2_THREADS = 32 * P_EPI2_WARPS;

struct __align__(8) PairHeader {
  uint64_t a_full[P_NA_MAX], a_empty[P_NA_MAX];
  uint64_t w_full[P_MAXNW], w_empty[P_MAXNW];
  uint64_t acc1_full[P_MAX_STAGES], acc1_empty[P_MAX_STAGES], a2_full[P_MAX_STAGES], a2_empty[P_MAX_STAGES], acc2_full[P_MAX_STAGES], acc2_empty[P_MAX_STAGES];
  uint32_t tmem_base;
  uint32_t pad;
};
constexpr int P_HEADER_BYTES = 1024;
static_assert(sizeof(PairHeader) <= P_HEADER_BYTES, "header");

__global__ void __launch_bounds__(P_THREADS, 1) conv_tc_pair_kernel(const ConvPairArgs pa, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) uint8_t smem[];
  PairHeader* hdr = reinterpret_cast<PairHeader*>(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = pa.C, K = pa.K, nchunks = pa.C / KC;
  const int per = nchunks * K;  // weight stages per conv
  const int na = pa.na, nw = pa.nw, resident = pa.resident;
  const uint32_t a1_stage = (uint32_t)pa.rows1 * 128u;          // hi + lo planes of rows1 x 64 B
  const uint32_t a2_chunk = (uint32_t)pa.rows2 * 128u;          // one 32-channel chunk of the xt tile
  const uint32_t a2_buf = a2_chunk * (uint32_t)nchunks;         // one xt tile (two of them)
  const uint32_t w_plane2 = (uint32_t)N * 32u, w_stage = w_plane2 * KG;
  float* bias_s = reinterpret_cast<float*>(smem + P_HEADER_BYTES);  // bias1[N] then bias2[N]
  uint8_t* a1_smem = smem + pa.a_off;
  uint8_t* a2_smem = smem + pa.a2_off;
  uint8_t* w_smem = smem + pa.w_off;
  const int items = pa.items;
  const int n_my = (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int TO = pa.TO, h = pa.h;
  const int ns = pa.ns, ns_shift = ns == 4 ? 2 : 1;  // stages of acc1 / xt tile / acc2 (items in flight)

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < na; ++i) mbar_init(&hdr->a_full[i], 1), mbar_init(&hdr->a_empty[i], 1);
    for (int i = 0; i < P_MAXNW; ++i) mbar_init(&hdr->w_full[i], 1), mbar_init(&hdr->w_empty[i], 1);
    for (int i = 0; i < P_MAX_STAGES; ++i) {
      mbar_init(&hdr->acc1_full[i], 1), mbar_init(&hdr->acc1_empty[i], P_EPI1_THREADS);
      mbar_init(&hdr->a2_full[i], P_EPI1_THREADS), mbar_init(&hdr->a2_empty[i], 1);
      mbar_init(&hdr->acc2_full[i], 1), mbar_init(&hdr->acc2_empty[i], P_EPI2_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&hdr->tmem_base, (uint32_t)(4 * ns * N));
  for (int i = tid; i < 2 * N; i += P_THREADS) bias_s[i] = __ldg(i < N ? pa.bias1 + i : pa.bias2 + (i - N));
  // rows 128 .. rows2-1 of the xt tiles are read by conv2's last taps for output rows that are never
  // stored; they only have to be finite: zero the tiles once
  for (uint32_t i = tid; i < (uint32_t)ns * a2_buf / 16; i += P_THREADS) reinterpret_cast<uint4*>(a2_smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = hdr->tmem_base;
  griddep_launch_dependents();  // PDL, as in conv_tc.cu: only weights / biases are touched before griddep_wait()
  if constexpr (P_EPI1_WARPS == 8) {
    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(P_REGS_PRODUCER));
  }

  if (warp == 0) {
    // ------------------------------------------------ weight producer
    if (lane == 0) {
      const uint8_t* src[2] = {reinterpret_cast<const uint8_t*>(pa.w1), reinterpret_cast<const uint8_t*>(pa.w2)};
      if (resident) {  // both convs' images stay in shared memory for every item of this CTA
        for (int c = 0; c < 2; ++c)
          for (int it = 0; it < per; ++it) {
            const int st = c * per + it;
            mbar_arrive_expect_tx(&hdr->w_full[st], w_stage);
            bulk_g2s(w_smem + (size_t)st * w_stage, src[c] + (size_t)it * w_stage, w_stage, &hdr->w_full[st]);
          }
      } else {  // ring, in the order the MMA warp consumes: conv1(0), conv1(1), conv2(0), conv1(2), conv2(1), ...
        int st = 0;
        uint32_t ph = 0;
        auto seq = [&](int c) {
          for (int it = 0; it < per; ++it) {
            mbar_wait(&hdr->w_empty[st], ph ^ 1);
            mbar_arrive_expect_tx(&hdr->w_full[st], w_stage);
            bulk_g2s(w_smem + (size_t)st * w_stage, src[c] + (size_t)it * w_stage, w_stage, &hdr->w_full[st]);
            if (++st == nw) st = 0, ph ^= 1;
          }
        };
        for (int i = 0; i < n_my; ++i) {
          seq(0);
          if (i > 0) seq(1);
        }
        if (n_my > 0) seq(1);
      }
    }
    __syncwarp();
  } else if (warp == 1 || (warp == 3 && pa.dual_issue)) {
    // ------------------------------------------------ MMA issuer(s) (whole warp walks, elected lane issues)
    // With resident weights (`dual_issue`) warp 1 issues every conv1 and the otherwise idle warp 3 every conv2: the two
    // streams use disjoint barriers, accumulators and A tiles, so they need no mutual order, their MMAs interleave in the
    // tensor pipe and one issuer's per-item waits hide under the other's MMAs (a narrow MMA lasts about as long as it
    // takes to issue, so a single lane never runs ahead of the pipe: conv_tc.cu, tools/mma_bench3.cu).
    const uint32_t idesc1 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * N) >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t b_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_hi = (512u >> 4) | (1u << 14) | (4u << 29);  // SWIZZLE_64B, SBO = 512 B
    const uint32_t a1_lo0 = ((smem_u32(a1_smem) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t a2_lo0 = ((smem_u32(a2_smem) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t w_lo0 = ((smem_u32(w_smem) & 0x3FFFFu) >> 4) | ((w_plane2 >> 4) << 16);
    const uint32_t a1_stage16 = a1_stage >> 4, a2_chunk16 = a2_chunk >> 4, a2_buf16 = a2_buf >> 4, w_stage16 = w_stage >> 4;
    const uint32_t lo1_16 = ((uint32_t)pa.rows1 * 64u) >> 4, lo2_16 = ((uint32_t)pa.rows2 * 64u) >> 4;
    const uint32_t ks_a16 = 2u, ks_b16 = (2 * w_plane2) >> 4;
    const uint32_t bar_a_full = smem_u32(&hdr->a_full[0]), bar_a_empty = smem_u32(&hdr->a_empty[0]);
    const uint32_t bar_w_full = smem_u32(&hdr->w_full[0]), bar_w_empty = smem_u32(&hdr->w_empty[0]);
    const bool leader = elect_one();
    int ast = 0, wst = 0;
    uint32_t aph = 0, wph = 0;

    // one (chunk, tap): 2 k-steps x [main|cross] (+)= xh.[wh|wl] ; cross += xl.wh
    auto tap = [&](uint32_t dmain, uint32_t ah, uint32_t lo16, int wslot, uint32_t acc) {
      const uint32_t bw = w_lo0 + (uint32_t)wslot * w_stage16;
      umma_f16_lo(dmain, ah, bw, a_hi, b_hi, idesc2, acc);
      umma_f16_lo(dmain + (uint32_t)N, ah + lo16, bw, a_hi, b_hi, idesc1, 1u);
      umma_f16_lo(dmain, ah + ks_a16, bw + ks_b16, a_hi, b_hi, idesc2, 1u);
      umma_f16_lo(dmain + (uint32_t)N, ah + ks_a16 + lo16, bw + ks_b16, a_hi, b_hi, idesc1, 1u);
    };
    // weight slot of stage `it` of conv c: resident -> fixed slot (wait only the first time), else the ring
    auto w_acquire = [&](int c, int it, bool first) -> int {
      if (resident) {
        const int slot = c * per + it;
        if (first) {
          mbar_wait_u32(bar_w_full + 8u * slot, 0u);
          tc_fence_after();
        }
        return slot;
      }
      mbar_wait_u32(bar_w_full + 8u * wst, wph);
      tc_fence_after();
      return wst;
    };
    auto w_release = [&]() {
      if (!resident) {
        if (leader) umma_commit_u32(bar_w_empty + 8u * wst);
        if (++wst == nw) wst = 0, wph ^= 1;
      }
    };
    auto conv1 = [&](int i) {
      const int s = i & (ns - 1);
      mbar_wait(&hdr->acc1_empty[s], ((uint32_t)(i >> ns_shift) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t dmain = tmem + (uint32_t)(s * 2 * N);
      uint32_t acc = 0;
      for (int ch = 0; ch < nchunks; ++ch) {
        mbar_wait_u32(bar_a_full + 8u * ast, aph);
        tc_fence_after();
        uint32_t ah = a1_lo0 + (uint32_t)ast * a1_stage16;
        if (resident && i > 0) {
          // resident weights, already seen: nothing to wait for or release per tap.  These narrow MMAs take ~46 cycles,
          // so the issuing lane is the bottleneck: bare loop, 4 MMAs + 2 adds per tap (as in conv_tc.cu)
          if (leader) {
            uint32_t bw = w_lo0 + (uint32_t)(ch * K) * w_stage16;
            const uint32_t dstep = (uint32_t)pa.dil1 * 4u;
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
              umma_f16_lo(dmain, ah, bw, a_hi, b_hi, idesc2, acc);
              umma_f16_lo(dmain + (uint32_t)N, ah + lo1_16, bw, a_hi, b_hi, idesc1, 1u);
              umma_f16_lo(dmain, ah + ks_a16, bw + ks_b16, a_hi, b_hi, idesc2, 1u);
              umma_f16_lo(dmain + (uint32_t)N, ah + ks_a16 + lo1_16, bw + ks_b16, a_hi, b_hi, idesc1, 1u);
              acc = 1u;
              ah += dstep, bw += w_stage16;
            }
            umma_commit_u32(bar_a_empty + 8u * ast);
          }
          acc = 1u;
          if (++ast == na) ast = 0, aph ^= 1;
          continue;
        }
        for (int j = 0; j < K; ++j) {
          const int slot = w_acquire(0, ch * K + j, i == 0);
          if (leader) tap(dmain, ah, lo1_16, slot, acc);
          acc = 1u;
          ah += (uint32_t)pa.dil1 * 4u;
          w_release();
        }
        if (leader) umma_commit_u32(bar_a_empty + 8u * ast);
        if (++ast == na) ast = 0, aph ^= 1;
      }
      if (leader) umma_commit(&hdr->acc1_full[s]);
    };
    auto conv2 = [&](int i) {
      const int s = i & (ns - 1);
      mbar_wait(&hdr->a2_full[s], (uint32_t)(i >> ns_shift) & 1u);   // xt tile written by the epilogue warps
      mbar_wait(&hdr->acc2_empty[s], ((uint32_t)(i >> ns_shift) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t dmain = tmem + (uint32_t)(2 * ns * N + s * 2 * N);
      uint32_t acc = 0;
      for (int ch = 0; ch < nchunks; ++ch) {
        uint32_t ah = a2_lo0 + (uint32_t)s * a2_buf16 + (uint32_t)ch * a2_chunk16;
        if (resident && i > 0) {
          if (leader) {
            uint32_t bw = w_lo0 + (uint32_t)(per + ch * K) * w_stage16;
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
              umma_f16_lo(dmain, ah, bw, a_hi, b_hi, idesc2, acc);
              umma_f16_lo(dmain + (uint32_t)N, ah + lo2_16, bw, a_hi, b_hi, idesc1, 1u);
              umma_f16_lo(dmain, ah + ks_a16, bw + ks_b16, a_hi, b_hi, idesc2, 1u);
              umma_f16_lo(dmain + (uint32_t)N, ah + ks_a16 + lo2_16, bw + ks_b16, a_hi, b_hi, idesc1, 1u);
              acc = 1u;
              ah += 4u, bw += w_stage16;
            }
          }
          acc = 1u;
          continue;
        }
        for (int j = 0; j < K; ++j) {
          const int slot = w_acquire(1, ch * K + j, i == 0);
          if (leader) tap(dmain, ah, lo2_16, slot, acc);
          acc = 1u;
          ah += 4u;  // conv2 has dilation 1: next tap = next 64 B row
          w_release();
        }
      }
      if (leader) {
        umma_commit(&hdr->a2_empty[s]);
        umma_commit(&hdr->acc2_full[s]);
      }
    };
    if (pa.dual_issue) {
      if (warp == 1) {
        for (int i = 0; i < n_my; ++i) conv1(i);
      } else {
        for (int i = 0; i < n_my; ++i) conv2(i);
      }
    } else {
      for (int i = 0; i < n_my; ++i) {
        conv1(i);
        if (i > 0) conv2(i - 1);
      }
      if (n_my > 0) conv2(n_my - 1);
    }
    __syncwarp();
  } else if (warp < 4) {
    // ------------------------------------------------ x-image loader: one TMA box per 32-channel chunk
    if (warp == 2 && lane == 0) {
      griddep_wait();
      int as = 0;
      uint32_t ph = 0;
      const int cgs = pa.C >> 5;
      for (int i = 0; i < n_my; ++i) {
        const int item = (int)blockIdx.x + i * (int)gridDim.x;
        const int b = (int)fast_div((uint32_t)item, pa.div_t), tt = item - b * pa.ntiles_t;
        const int t_origin = tt * TO - h - h * pa.dil1;  // first staged row of x: what xt[t0 - h]'s first tap reads
        for (int ch = 0; ch < nchunks; ++ch) {
          mbar_wait(&hdr->a_empty[as], ph ^ 1);
          mbar_arrive_expect_tx(&hdr->a_full[as], a1_stage);
          tma_load_4d(a1_smem + (size_t)as * a1_stage, &tmap, 0, t_origin, b * cgs + ch, 0, &hdr->a_full[as]);
          if (++as == na) as = 0, ph ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ epilogue warps
    griddep_wait();
    // Two role-specialised groups of four warps (one per TMEM lane quarter), each walking ALL items: group 0 turns
    // conv1's accumulators into conv2's A tile (epi1), group 1 stores conv2's result (epi2).  Alternating both jobs on
    // the same warps kept only one item in flight (conv1 -> epi1 -> conv2 -> epi2 is a serial chain of ~4000 cycles per
    // item); split, epi1(i+1) runs under conv2(i)'s MMAs and epi2(i) under conv1(i+2)'s.
    const int q4 = warp & 3, role2 = warp >= 4 + P_EPI1_WARPS ? 1 : 0;
    const int part = role2 ? (warp - 4 - P_EPI1_WARPS) >> 2 : (warp - 4) >> 2;  // which of the warps of this lane quarter
    const int row = q4 * 32 + lane;
    const int hc = N >> 4;  // 16-column jobs per item and warp: 2 (N = 32) or 4 (N = 64)
    const int n_lo = 0, n_hi = N;
    const size_t plane = (size_t)pa.B * pa.C * pa.L;  // halves between the hi and lo planes of an image
    const float r_inv = pa.res_img ? 1.0f / pa.res_slope : 1.0f;

    auto item_bt = [&](int i, int& b, int& tt) {
      const int item = (int)blockIdx.x + i * (int)gridDim.x;
      b = (int)fast_div((uint32_t)item, pa.div_t);
      tt = item - b * pa.ntiles_t;
    };
    if (role2 == 0) {
    if constexpr (P_EPI1_WARPS == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(P_REGS_EPI1));
    // epi1: conv1 accumulators -> conv2's A tile in shared memory
    auto epi1 = [&](int i) {
      const int s = i & (ns - 1);
      int b, tt;
      item_bt(i, b, tt);
      const int u = tt * TO - h + row;  // time step of xt held by this thread's row
      const bool inside = u >= 0 && u < pa.L;
      mbar_wait(&hdr->acc1_full[s], (uint32_t)(i >> ns_shift) & 1u);
      mbar_wait(&hdr->a2_empty[s], ((uint32_t)(i >> ns_shift) & 1u) ^ 1u);  // conv2(i-2) has finished reading this tile
      tc_fence_after();
      const uint32_t tsub = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(s * 2 * N);
      for (int n0 = n_lo + (P_EPI1_WARPS == 8 ? 16 * part : 0); n0 < n_hi; n0 += 16 * (P_EPI1_WARPS / 4)) {
        uint32_t m[16], c[16];
        tmem_ld16(tsub + (uint32_t)n0, m);
        tmem_ld16(tsub + (uint32_t)(N + n0), c);
        tmem_wait_ld();
        float v[16];
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) {
          const float4 bq = reinterpret_cast<const float4*>(bias_s + n0)[e4];
          const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float q = fmaf(fmaf(__uint_as_float(c[4 * e4 + e]), LO_INV, __uint_as_float(m[4 * e4 + e])), pa.unscale1, bb[e]);
            q = inside ? q : 0.f;                              // conv2 zero-pads xt outside [0, L)
            v[4 * e4 + e] = q > 0.f ? q : q * pa.xt_slope;      // leaky_relu between the convs (modules.py:216)
          }
        }
        uint8_t* tile = a2_smem + (size_t)s * a2_buf + (size_t)(n0 >> 5) * a2_chunk;
        uint4* hi = reinterpret_cast<uint4*>(tile) + row * KG;
        uint4* lo = reinterpret_cast<uint4*>(tile + (size_t)pa.rows2 * 64) + row * KG;
        const int swz = (int)((smem_u32(hi) >> 7) & 3u);  // 64 B swizzle on absolute address bits 7-8
        const int kg0 = (n0 & 31) >> 3;
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          uint4 hq, lq;
          split2(v[8 * g8 + 0], v[8 * g8 + 1], hq.x, lq.x);
          split2(v[8 * g8 + 2], v[8 * g8 + 3], hq.y, lq.y);
          split2(v[8 * g8 + 4], v[8 * g8 + 5], hq.z, lq.z);
          split2(v[8 * g8 + 6], v[8 * g8 + 7], hq.w, lq.w);
          hi[(kg0 + g8) ^ swz] = hq;
          lo[(kg0 + g8) ^ swz] = lq;
        }
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
      tc_fence_before();
      mbar_arrive(&hdr->a2_full[s]);
      mbar_arrive(&hdr->acc1_empty[s]);
    };
    for (int i = 0; i < n_my; ++i) epi1(i);
    } else {
    // the epi2 code below is dominated by this point, so ptxas lets it use the larger register file
    if constexpr (P_EPI1_WARPS == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(P_REGS_EPI2));

    // epi2 operands (residual, running sum) are requested two of this warp's jobs ahead of use into two buffers that its
    // jobs use alternately.  The two epi2 warps of a lane quarter take alternate 16-column chunks of an item, so a warp has
    // hc2 = hc / 2 jobs per item (1 at N = 32, 2 at N = 64); warp job w = item * hc2 + kk covers columns 16 * (2 kk + part).
    float rA[16], rB[16];
    const int hc2 = hc >> 1;
    const bool res_raw = pa.res_img && !pa.acc_in;  // the residual buffers hold raw image words until they are used
    // raw image words (see load_ops) -> residual values: x = hi + lo * 2^-11, leaky_relu inverted
    auto decode_res = [&](float (&q)[16]) {
#pragma unroll
      for (int g8 = 0; g8 < 2; ++g8) {
        float o[8];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const uint32_t hw = __float_as_uint(q[8 * g8 + e2]), lw = __float_as_uint(q[8 * g8 + 4 + e2]);
          const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw));
          const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw));
          const float v0 = fmaf(lf.x, LO_INV, hf.x), v1 = fmaf(lf.y, LO_INV, hf.y);
          o[2 * e2] = v0 >= 0.f ? v0 : v0 * r_inv;
          o[2 * e2 + 1] = v1 >= 0.f ? v1 : v1 * r_inv;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) q[8 * g8 + e] = o[e];
      }
    };
    auto load_ops = [&](float (&q)[16], int wjob) {
      const int i = hc2 == 1 ? wjob : wjob >> 1, n0 = n_lo + ((2 * (hc2 == 1 ? 0 : (wjob & 1)) + part) << 4);
      if (i >= n_my) return;
      int b, tt;
      item_bt(i, b, tt);
      const int t = min(tt * TO + row, pa.L - 1);  // clamped: always a valid address
      const size_t off = ((size_t)b * pa.C + n0) * pa.L + t;
      if (pa.res_img) {
        // the raw fp16 words travel in the buffer (q[8 g8 + 0..3] = hi, + 4..7 = lo) and are converted where the residual
        // is USED, two jobs later: converting here would park the warp on the loads it has just issued
        const uint16_t* rp = pa.res_img + (((size_t)b * (pa.C >> 5) + (n0 >> 5)) * pa.L + t) * 32 + (n0 & 31);
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          const uint4 hq = *reinterpret_cast<const uint4*>(rp + g8 * 8);
          const uint4 lq = *reinterpret_cast<const uint4*>(rp + g8 * 8 + plane);
          q[8 * g8 + 0] = __uint_as_float(hq.x), q[8 * g8 + 1] = __uint_as_float(hq.y);
          q[8 * g8 + 2] = __uint_as_float(hq.z), q[8 * g8 + 3] = __uint_as_float(hq.w);
          q[8 * g8 + 4] = __uint_as_float(lq.x), q[8 * g8 + 5] = __uint_as_float(lq.y);
          q[8 * g8 + 6] = __uint_as_float(lq.z), q[8 * g8 + 7] = __uint_as_float(lq.w);
        }
        if (pa.acc_in) decode_res(q);  // the running sum is added to VALUES below (only the last pair of a block has one)
      } else if (pa.res) {
        const float* rp = pa.res + off;
#pragma unroll
        for (int e = 0; e < 16; ++e) q[e] = rp[(size_t)e * pa.L];
      }
      if (pa.acc_in) {
        const float* ap = pa.acc_in + off;
#pragma unroll
        for (int e = 0; e < 16; ++e) q[e] += ap[(size_t)e * pa.L];
      }
    };
#pragma unroll
    for (int e = 0; e < 16; ++e) rA[e] = 0.f, rB[e] = 0.f;  // stays zero when the pair has no residual operand at all
    load_ops(rA, 0);
    load_ops(rB, 1);

    // epi2: conv2 accumulators -> y = conv2 + bias + x (+ running sum) (/ post_div) -> fp32 and/or image
    const float4* bias2_4 = reinterpret_cast<const float4*>(bias_s + N);
    auto epi2_job = [&](float (&r)[16], int job, int b, int t, bool valid, uint32_t tsub, int n0, int release_stage) {
      uint32_t m[16], c[16];
      tmem_ld16(tsub + (uint32_t)n0, m);
      tmem_ld16(tsub + (uint32_t)(N + n0), c);
      tmem_wait_ld();
      if (release_stage >= 0) {  // this warp's last tcgen05.ld of the item has completed: hand the accumulator stage back
        tc_fence_before();       // before the residual / store work, so that conv2(i + 2) does not wait for it
        mbar_arrive(&hdr->acc2_empty[release_stage]);
      }
      if (res_raw) decode_res(r);
      float v[16];
#pragma unroll
      for (int e4 = 0; e4 < 4; ++e4) {
        const float4 q = bias2_4[(n0 >> 2) + e4];
        v[4 * e4 + 0] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 0]), LO_INV, __uint_as_float(m[4 * e4 + 0])), pa.unscale2, q.x) + r[4 * e4 + 0];
        v[4 * e4 + 1] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 1]), LO_INV, __uint_as_float(m[4 * e4 + 1])), pa.unscale2, q.y) + r[4 * e4 + 1];
        v[4 * e4 + 2] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 2]), LO_INV, __uint_as_float(m[4 * e4 + 2])), pa.unscale2, q.z) + r[4 * e4 + 2];
        v[4 * e4 + 3] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 3]), LO_INV, __uint_as_float(m[4 * e4 + 3])), pa.unscale2, q.w) + r[4 * e4 + 3];
      }
      load_ops(r, job + 2);  // refill this buffer for the job after next
      if (pa.post_div != 1.0f) {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = v[e] / pa.post_div;
      }
      if (valid) {
        if (pa.y_img) {
          uint16_t* sp = pa.y_img + (((size_t)b * (pa.C >> 5) + (n0 >> 5)) * pa.L + t) * 32 + (n0 & 31);
          uint4 hq[2], lq[2];
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            float w8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) w8[e] = v[8 * g8 + e] > 0.f ? v[8 * g8 + e] : v[8 * g8 + e] * pa.y_slope;
            split2(w8[0], w8[1], hq[g8].x, lq[g8].x);
            split2(w8[2], w8[3], hq[g8].y, lq[g8].y);
            split2(w8[4], w8[5], hq[g8].z, lq[g8].z);
            split2(w8[6], w8[7], hq[g8].w, lq[g8].w);
          }
          st_global_v8(sp, hq[0], hq[1]);  // one full 32 B sector per plane and row
          st_global_v8(sp + plane, lq[0], lq[1]);
        }
        if (pa.y) {
          float* yp = pa.y + ((size_t)b * pa.C + n0) * pa.L + t;
#pragma unroll
          for (int e = 0; e < 16; ++e) yp[(size_t)e * pa.L] = v[e];
        }
      }
    };
    // one item of this warp; bufs: the operand buffer(s) its hc2 jobs use (hc2 == 1: one, alternating between items)
    auto epi2 = [&](int i, float (&r0)[16], float (&r1)[16]) {
      const int s = i & (ns - 1);
      int b, tt;
      item_bt(i, b, tt);
      const int t = tt * TO + row;
      const bool valid = row < TO && t < pa.L;
      mbar_wait(&hdr->acc2_full[s], (uint32_t)(i >> ns_shift) & 1u);
      tc_fence_after();
      const uint32_t tsub = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(2 * ns * N + s * 2 * N);
      if (hc2 == 1) {
        epi2_job(r0, i, b, t, valid, tsub, 16 * part, s);
      } else {
        epi2_job(r0, 2 * i, b, t, valid, tsub, 16 * part, -1);
        epi2_job(r1, 2 * i + 1, b, t, valid, tsub, 16 * (2 + part), s);
      }
    };

    if (hc2 == 1) {
      for (int i = 0; i < n_my; i += 2) {  // one job per item: the two buffers alternate between items
        epi2(i, rA, rA);
        if (i + 1 < n_my) epi2(i + 1, rB, rB);
      }
    } else {
      for (int i = 0; i < n_my; ++i) epi2(i, rA, rB);
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)(4 * ns * N));
}

}  // namespace

bool conv_tc_pair_supported(int C, int K, int dil1) {
  return (C == 32 || C == 64) && K >= 1 && K <= 11 && (K & 1) && dil1 >= 1 && 128 + (K - 1) * dil1 <= 248;
}

cudaError_t launch_conv_tc_pair(const ConvPairArgs& in, cudaStream_t stream) {
  ConvPairArgs pa = in;
  if (!conv_tc_pair_supported(pa.C, pa.K, pa.dil1) || !pa.x_img || !pa.w1 || !pa.w2 || (!pa.y && !pa.y_img)) return cudaErrorInvalidValue;
  if (pa.res && pa.res_img) return cudaErrorInvalidValue;
  if (pa.B <= 0 || pa.L <= 0) return cudaSuccess;
  const int N = pa.C, nchunks = pa.C / KC, per = nchunks * pa.K;
  pa.h = (pa.K - 1) / 2;
  pa.TO = 128 - (pa.K - 1);
  pa.rows1 = (128 + (pa.K - 1) * pa.dil1 + 7) & ~7;
  pa.rows2 = (128 + pa.K - 1 + 7) & ~7;
  pa.ntiles_t = (pa.L + pa.TO - 1) / pa.TO;
  const long long items = (long long)pa.B * pa.ntiles_t;
  if (items > 0x7FFFFFFFLL / 8) return cudaErrorInvalidValue;
  pa.items = (int)items;
  pa.div_t = make_fast_div((uint32_t)pa.ntiles_t);
  // shared memory plan: header | biases | A1 ring | two xt tiles | weights
  static const int want_ns = [] {
    const char* e = getenv("SVK_PAIR_STAGES");
    return e ? atoi(e) : 4;
  }();
  pa.ns = (want_ns >= 4 && 16 * N <= 512) ? 4 : 2;  // 4 ns N TMEM columns
  size_t a1_stage = (size_t)pa.rows1 * 128, a2_bytes = (size_t)pa.ns * nchunks * pa.rows2 * 128, w_stage = (size_t)N * 128;
  if (pa.ns == 4 && (P_HEADER_BYTES + 2 * N * 4 + 1023) / 1024 * 1024 + a2_bytes + 2 * (size_t)per * w_stage + 4 * a1_stage > 227 * 1024)
    pa.ns = 2, a2_bytes = (size_t)2 * nchunks * pa.rows2 * 128;  // four xt tiles must leave room for resident weights and an A ring of four
  const size_t fixed = (P_HEADER_BYTES + (size_t)2 * N * 4 + 1023) & ~(size_t)1023;
  const size_t budget = 227 * 1024 - fixed - a2_bytes;
  pa.resident = (2 * per <= P_MAXNW && 2 * per * w_stage + 2 * a1_stage <= budget) ? 1 : 0;
  if (pa.resident) {
    pa.nw = 2 * per;
    pa.na = (int)((budget - pa.nw * w_stage) / a1_stage);
  } else {
    pa.na = 4;
    if (pa.na * a1_stage + 6 * w_stage > budget) pa.na = 2;
    pa.nw = (int)((budget - pa.na * a1_stage) / w_stage);
    if (pa.nw > P_MAXNW) pa.nw = P_MAXNW;
    if (pa.nw < 4) return cudaErrorInvalidValue;
    const int na2 = (int)((budget - pa.nw * w_stage) / a1_stage);  // left-over goes back to the A ring
    if (na2 > pa.na) pa.na = na2;
  }
  if (pa.na > P_NA_MAX) pa.na = P_NA_MAX;
  if (pa.na < 2) return cudaErrorInvalidValue;
  static const bool dual = [] {
    const char* e = getenv("SVK_DUAL_ISSUE");
    return !(e && e[0] == '0');
  }();
  pa.dual_issue = dual && pa.resident;  // the weight RING is consumed in one interleaved order: one issuer only
  pa.a_off = (int)fixed;
  pa.a2_off = (int)(fixed + pa.na * a1_stage);
  pa.w_off = (int)(pa.a2_off + a2_bytes);
  const size_t smem = pa.w_off + (size_t)pa.nw * w_stage;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;

  static int sm_count[64] = {0};
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  cudaError_t e = tc_make_image_map(pa.x_img, pa.B, pa.C, pa.L, pa.rows1, 2, &map);
  if (e != cudaSuccess) return e;
  const int grid = pa.items < sm_count[dev] ? pa.items : sm_count[dev];
  return launch_pdl(conv_tc_pair_kernel, grid, P_THREADS, smem, stream, pa, map);
}

}  // namespace svk
