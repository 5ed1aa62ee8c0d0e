// Stride-1 Conv1d (dilated, any tap count) as an implicit GEMM on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM), at fp32-class accuracy through a three-product fp16 split:
//
//     x = xh + xl,  w*s = wh + wl   (xh = fp16(x), xl = fp16(x - xh); s = per-layer power of two)
//     y = (xh.wh + xh.wl + xl.wh) / s          products exact in the fp32 accumulator; the dropped
//                                               xl.wl term is <= 2^-22 relative
//
// PERSISTENT, warp-specialised kernel: one CTA per SM walks a static round-robin list of work items
// (item = one [128 time steps] x [N output channels] tile of one utterance).
//     M = time (128 rows per MMA), N = output channels, K = input channels x taps
//     A = activations, K-major, 64 B-swizzled:  A_s[hi|lo][row = time][32 ch x 2 B], the four 16 B
//         chunks of a row XOR-ed with address bits 7-8 (CU_TENSOR_MAP_SWIZZLE_64B / UMMA layout 4)
//         -> a conv tap is a ROW shift = +64 B per row on the descriptor's start address.  The tensor
//            core de-swizzles on absolute address bits (tools/swizzle_probe.cu: any row shift is exact
//            with base offset 0), and 8 consecutive rows always hit 8 distinct bank groups, so a shifted
//            tile reads at full rate (tools/mma_bench2.cu: no shift residue costs a cycle, on either
//            operand side).  One staged tile with a (K-1)*dil halo serves every tap: no im2col.
//     B = weights per (32-channel chunk, tap), K-major, no swizzle: B_s[kgroup][hi n | lo n][16 B],
//         pre-split/pre-packed at load time in exactly this image (one cp.async.bulk per stage).
//         Because hi and lo rows are adjacent, ONE MMA with N' = 2N computes xh.wh into the main
//         accumulator and xh.wl into the cross accumulator (columns [N, 2N)); a second MMA with
//         N' = N adds xl.wh to the cross accumulator: 2 instructions and 20 KB of operand reads
//         per K=16 step instead of 3 and 24 KB.  The small cross terms never touch the main
//         accumulator (3x lower truncation error than one accumulator).
//     D = fp32 in TMEM, double buffered: 2 stages x (main + cross) x N columns <= 512.
//
// Warp roles (576 threads): warp 0 = TMEM allocator + weight bulk-copy producer (1 lane),
// warp 1 = barrier init + MMA issuer (1 lane), warps 2..9 = activation producers in two groups
// that alternate 32-channel chunks (global fp32 -> leaky_relu/mask -> fp16 hi/lo -> smem), running
// ahead of the MMAs by the depth of the A ring (across tile boundaries), warps 10..17 = epilogue
// (tcgen05.ld -> bias / residual / running sum / mask / tanh / gate / polyphase store) of tile i while
// the MMAs of tile i+1 fill the other accumulator stage.
// Pipelines: A ring (mbarrier full/empty), weight ring (expect_tx / tcgen05.commit) -- or, when the
// layer's whole weight image fits the ring, weights are loaded once and stay resident for every tile
// of the CTA -- and the accumulator full/empty pair.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <type_traits>

#include "svk_kernels.cuh"
// Barrier waits of THIS kernel suspend with a 20 us hint (tc_common.cuh): its eight epilogue warps wait for accumulators
// most of the time on the wide layers, and every retry of a waiting warp takes issue slots from the warps that work.
// A/B of the whole library with the hint (three alternations on one box): conv1 + conv2 launches -2 %, step 32.38 ->
// 32.06 ms -- but the fused pairs +1 % and the WN stack +3.6 % (their hand-overs between roles are latency-critical and
// a hinted wait wakes a little later), so the pair and WN kernels keep the default limit.
#ifndef SVK_MBAR_HINT_NS
#define SVK_MBAR_HINT_NS 20000
#endif
#include "tc_common.cuh"

namespace svk {

namespace {

constexpr int NA_MAX = 8;   // A ring depth limit (producer warps: 2 or 4, the groups own alternate slots; TMA: up to 8)
constexpr int MAXNW = 24;   // weight ring depth limit
constexpr int MAXACC = 8;   // accumulator ring depth limit (512 TMEM columns / 2N, power of two)
#ifndef SVK_TC_PROD_WARPS
#define SVK_TC_PROD_WARPS 8
#endif
#ifndef SVK_TC_EPI_WARPS
#define SVK_TC_EPI_WARPS 8
#endif
constexpr int PROD_WARPS = SVK_TC_PROD_WARPS, EPI_WARPS = SVK_TC_EPI_WARPS;
constexpr int THREADS = 64 + 32 * (PROD_WARPS + EPI_WARPS);  // 576
constexpr int PROD_GROUP = 128;                              // threads per producer group (one time row each)
constexpr int PROD_GROUPS = PROD_WARPS / 4;                  // groups take alternate chunks
constexpr int EPI_SPLIT = EPI_WARPS / 4;                     // warps sharing one TMEM lane quarter split the columns
constexpr int FIRST_EPI_WARP = 2 + PROD_WARPS;
// TMA-input variant (activations already in HBM as fp16 hi/lo operand images): warps 0..3 are the
// weight producer, the MMA issuer, the activation TMA issuer and a spare; the rest is epilogue.
#ifndef SVK_TC_EPI_WARPS_TMA
#define SVK_TC_EPI_WARPS_TMA 8
#endif
constexpr int EPI_WARPS_TMA = SVK_TC_EPI_WARPS_TMA;
constexpr int THREADS_TMA = 128 + 32 * EPI_WARPS_TMA;
static_assert(EPI_WARPS_TMA % 4 == 0, "a warp may only read TMEM lane quarter (warp id & 3)");
static_assert(PROD_WARPS % 4 == 0 && (PROD_GROUPS == 1 || PROD_GROUPS == 2), "producer groups");
static_assert(EPI_WARPS % 4 == 0 && EPI_SPLIT >= 1, "a warp may only read TMEM lane quarter (warp id & 3)");

struct __align__(8) SmemHeader {
  uint64_t a_full[NA_MAX], a_empty[NA_MAX];
  uint64_t w_full[MAXNW], w_empty[MAXNW];
  uint64_t acc_full[MAXACC], acc_empty[MAXACC];
  uint32_t tmem_base;
  uint32_t pad;
};
constexpr int HEADER_BYTES = 768;
static_assert(sizeof(SmemHeader) <= HEADER_BYTES, "header");

// ------------------------------------------------------------------------------------- kernel
// item -> (output-channel tile, utterance, time tile); consecutive items are adjacent time tiles of
// one utterance and one channel tile, so a wave of CTAs shares weights and halos in L2.
__device__ __forceinline__ void decode_item(int item, const FastDiv& dt, const FastDiv& db, int& nt, int& b, int& tt) {
  const uint32_t r = fast_div((uint32_t)item, dt);
  tt = item - (int)(r * dt.d);
  const uint32_t q = fast_div(r, db);
  b = (int)(r - q * db.d);
  nt = (int)q;
}

template <bool kTma>
__global__ void __launch_bounds__(kTma ? THREADS_TMA : THREADS, 1)
    conv_tc_kernel(const ConvTcArgs ta, const __grid_constant__ CUtensorMap tmap) {
  constexpr int kThreads = kTma ? THREADS_TMA : THREADS;
  constexpr int kFirstEpi = kTma ? 4 : FIRST_EPI_WARP;
  constexpr int kEpiSplit = (kTma ? EPI_WARPS_TMA : EPI_WARPS) / 4;
  constexpr int kEpiThreads = 32 * (kTma ? EPI_WARPS_TMA : EPI_WARPS);
  extern __shared__ __align__(128) uint8_t smem[];
  const ConvArgs& a = ta.c;
  SmemHeader* hdr = reinterpret_cast<SmemHeader*>(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = ta.N, nw = ta.nw, na = ta.na;
  const int K = a.K, dil = a.dil;
  const int rows = ta.rows;                       // staged time rows per chunk (multiple of 8)
  const int planes = ta.planes;                   // 2: fp16 hi + lo (fp32-class), 1: single bf16 pass
  const uint32_t a_plane = (uint32_t)rows * 64u;  // one plane of A: rows x (32 ch x 2 B), 64 B-swizzled
  const uint32_t a_stage = a_plane * planes;       // hi (+ lo)
  const uint32_t w_plane2 = (uint32_t)N * 16u * planes;  // one k-group plane of B: N hi rows (then N lo rows)
  const uint32_t w_stage = w_plane2 * KG;
  const int acc_cols = planes * N;                // TMEM columns of one accumulator stage: main (+ cross)
  float* bias_s = reinterpret_cast<float*>(smem + HEADER_BYTES);  // [ntiles_n * N], staged once per CTA
  uint8_t* a_smem = smem + ta.a_off;  // 1024 B-aligned: the swizzle is a function of absolute address bits
  uint8_t* w_smem = a_smem + (size_t)na * a_stage;
  const int nchunks = a.Cin / KC;
  const int per_tile = nchunks * K;  // weight stages per item
  const int items = ta.items;
  const int resident = ta.resident;
  const int n_my = (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nacc = ta.nacc, nacc_shift = 31 - __clz(nacc);  // accumulator stages (power of two)

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < na; ++i) mbar_init(&hdr->a_full[i], kTma ? 1 : PROD_GROUP), mbar_init(&hdr->a_empty[i], 1);
    for (int i = 0; i < MAXNW; ++i) mbar_init(&hdr->w_full[i], 1), mbar_init(&hdr->w_empty[i], 1);
    for (int i = 0; i < MAXACC; ++i) mbar_init(&hdr->acc_full[i], 1), mbar_init(&hdr->acc_empty[i], kEpiThreads / ta.epi_groups);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&hdr->tmem_base, ta.tmem_cols);
  for (int i = tid; i < ta.bias_count; i += kThreads) bias_s[i] = __ldg(a.bias + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = hdr->tmem_base;
  // PDL: everything above (and the weight producer below) touches only weights / biases, which no kernel of the
  // stream writes; every role that reads activations or writes results first waits for the previous grid.
  griddep_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------ weight producer: one bulk copy per (chunk, tap)
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n_my; ++i) {
        if (resident && i > 0) break;  // the whole image already sits in the ring
        int nt, b, tt;
        decode_item((int)blockIdx.x + i * (int)gridDim.x, ta.div_t, ta.div_b, nt, b, tt);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(ta.wtc) + (size_t)nt * per_tile * w_stage;
        for (int it = 0; it < per_tile; ++it) {
          if (!resident) mbar_wait(&hdr->w_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&hdr->w_full[st], w_stage);
          bulk_g2s(w_smem + (size_t)st * w_stage, src + (size_t)it * w_stage, w_stage, &hdr->w_full[st]);
          if (++st == nw) st = 0, ph ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || (kTma && warp == 3 && ta.dual_issue)) {
    // ------------------------------------------------ MMA issuer(s)
    // The whole warp walks the loop (uniform control flow, barrier waits by all lanes); one elected
    // lane issues.  The tensor pipe only stays busy if issuing an MMA costs far less than the MMA
    // itself (64-128 cycles), so descriptors are never rebuilt: their high words are constants and
    // their low words (start address >> 4 | LBO) advance by 32-bit adds.
    // Narrow layers (resident weights, N' <= 128): an MMA lasts ~46-58 cycles whatever N is (tools/mma_bench3.cu), about
    // what the lane needs to issue it, so the lane never runs ahead of the pipe and the ~400 cycles it spends per tile on
    // barrier waits / commits / bookkeeping leave the pipe idle (ncu, C=32 k=7 conv1: 1811 cycles per tile against 1372
    // of MMA time).  With `dual_issue` the spare warp 3 is a second issuer: warp 1 takes the even tiles of the CTA's
    // list, warp 3 the odd ones -- different accumulator stages and A stages, so their MMAs need no mutual order,
    // and one warp's per-tile overhead hides under the other's MMAs.  tcgen05.commit tracks the MMAs of the
    // issuing thread, which is exactly the tile the barrier belongs to.
    {
      const int first = warp == 3 ? 1 : 0, step = ta.dual_issue ? 2 : 1;
      // f16 x f16 -> f32, both operands K-major, M = 128
      // (planes == 1: bf16 x bf16, a/b format fields = 1)
      const uint32_t fmt = planes == 1 ? ((1u << 7) | (1u << 10)) : 0u;
      const uint32_t idesc1 = (1u << 4) | fmt | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * N) >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t b_hi = (128u >> 4) | (1u << 14);                // B: no swizzle, SBO = 128 B, version 1
      const uint32_t a_hi = (512u >> 4) | (1u << 14) | (4u << 29);   // A: SWIZZLE_64B, SBO = 8 rows x 64 B, base offset 0
      const uint32_t a_lo0 = ((smem_u32(a_smem) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t w_lo0 = ((smem_u32(w_smem) & 0x3FFFFu) >> 4) | ((w_plane2 >> 4) << 16);
      const uint32_t a_stage16 = a_stage >> 4, w_stage16 = w_stage >> 4;
      const uint32_t lo_plane16 = a_plane >> 4;  // hi plane -> lo plane of one A stage
      const uint32_t ks_a16 = 32u >> 4, ks_b16 = (2 * w_plane2) >> 4;  // next 16 channels: +32 B inside the A row
      const uint32_t dil16 = (uint32_t)dil * 4u;        // one tap = dil rows = dil * 64 B
      const uint32_t bar_a_full = smem_u32(&hdr->a_full[0]), bar_a_empty = smem_u32(&hdr->a_empty[0]);
      const uint32_t bar_w_full = smem_u32(&hdr->w_full[0]), bar_w_empty = smem_u32(&hdr->w_empty[0]);
      const uint32_t bar_acc_full = smem_u32(&hdr->acc_full[0]), bar_acc_empty = smem_u32(&hdr->acc_empty[0]);
      const int wlim = resident ? per_tile : nw;
      const bool leader = elect_one();
      int ast = 0, wst = 0;
      uint32_t aph = 0, wph = 0;
      for (int q = first * nchunks; q > 0; --q)  // A stage of this issuer's first chunk
        if (++ast == na) ast = 0, aph ^= 1;
      for (int i = first; i < n_my; i += step) {
        const int s = i & (nacc - 1);
        mbar_wait_u32(bar_acc_empty + 8u * s, ((uint32_t)(i >> nacc_shift) & 1u) ^ 1u);  // epilogue drained this stage
        tc_fence_after();
        const uint32_t dmain = tmem + (uint32_t)(s * acc_cols), dcross = dmain + (uint32_t)N;
        const bool wait_w = !resident || i == first;
        uint32_t acc = 0;
        for (int ch = 0; ch < nchunks; ++ch) {
          mbar_wait_u32(bar_a_full + 8u * ast, aph);
          tc_fence_after();
          uint32_t ah = a_lo0 + (uint32_t)ast * a_stage16;  // tap 0, ks 0, hi planes
          if (!wait_w) {
            // resident weights, already seen: nothing to wait for or release per tap.  Narrow layers live here
            // and are bound by how fast this lane can issue (their MMAs take ~45 cycles), so the loop is bare:
            // 4 MMAs + 2 adds per tap.
            if (leader) {
              uint32_t bw = w_lo0 + (uint32_t)(ch * K) * w_stage16;
              if (planes == 2) {
#pragma unroll 1
                for (int j = 0; j < K; ++j) {
                  umma_f16_lo(dmain, ah, bw, a_hi, b_hi, idesc2, acc);
                  umma_f16_lo(dcross, ah + lo_plane16, bw, a_hi, b_hi, idesc1, 1u);
                  umma_f16_lo(dmain, ah + ks_a16, bw + ks_b16, a_hi, b_hi, idesc2, 1u);
                  umma_f16_lo(dcross, ah + ks_a16 + lo_plane16, bw + ks_b16, a_hi, b_hi, idesc1, 1u);
                  acc = 1u;
                  ah += dil16, bw += w_stage16;
                }
              } else {
#pragma unroll 1
                for (int j = 0; j < K; ++j) {
                  umma_f16_lo(dmain, ah, bw, a_hi, b_hi, idesc1, acc);
                  umma_f16_lo(dmain, ah + ks_a16, bw + ks_b16, a_hi, b_hi, idesc1, 1u);
                  acc = 1u;
                  ah += dil16, bw += w_stage16;
                }
              }
              umma_commit_u32(bar_a_empty + 8u * ast);
            }
            acc = 1u;
            if (++ast == na) ast = 0, aph ^= 1;
            continue;
          }
          for (int j = 0; j < K; ++j) {
            if (wait_w) {
              mbar_wait_u32(bar_w_full + 8u * wst, wph);
              tc_fence_after();
            }
            if (leader) {
              const uint32_t bw = w_lo0 + (uint32_t)wst * w_stage16;
              if (planes == 2) {
                umma_f16_lo(dmain, ah, bw, a_hi, b_hi, idesc2, acc);  // [main | cross] (+)= xh . [wh | wl]
                umma_f16_lo(dcross, ah + lo_plane16, bw, a_hi, b_hi, idesc1, 1u);  // cross += xl . wh
                umma_f16_lo(dmain, ah + ks_a16, bw + ks_b16, a_hi, b_hi, idesc2, 1u);
                umma_f16_lo(dcross, ah + ks_a16 + lo_plane16, bw + ks_b16, a_hi, b_hi, idesc1, 1u);
              } else {
                umma_f16_lo(dmain, ah, bw, a_hi, b_hi, idesc1, acc);  // main (+)= bf16(x) . bf16(w)
                umma_f16_lo(dmain, ah + ks_a16, bw + ks_b16, a_hi, b_hi, idesc1, 1u);
              }
              if (!resident) umma_commit_u32(bar_w_empty + 8u * wst);
            }
            acc = 1u;
            ah += dil16;
            if (++wst == wlim) wst = 0, wph ^= 1;
          }
          if (leader) umma_commit_u32(bar_a_empty + 8u * ast);
          if (++ast == na) ast = 0, aph ^= 1;
        }
        if (leader) umma_commit_u32(bar_acc_full + 8u * s);
        for (int q = (step - 1) * nchunks; q > 0; --q)  // skip the other issuer's tile
          if (++ast == na) ast = 0, aph ^= 1;
      }
    }
    __syncwarp();
  } else if (warp < kFirstEpi) {
    if constexpr (kTma) {
      // ---------------------------------------------- activation loader: one TMA box per 32-channel chunk
      // The source tensor is the fp16 operand image [hi|lo][B*C/32][L][32] written by the producing
      // layer's epilogue (or split_image_kernel); box (32 ch, rows, 1, planes) with the 64 B swizzle
      // lands exactly in the A stage layout.  Rows outside [0, L) are zero-filled by the TMA unit = the conv's zero padding
      // (leaky_relu and mask were applied when the image was written).
      if (warp == 2 && lane == 0) {
        griddep_wait();
        const int total_q = n_my * nchunks;
        const int cg0 = a.x_ch_off >> 5, cgs = a.x_C >> 5;
        int as = 0;
        uint32_t ph = 0;
        int q = 0;
        for (int i = 0; i < n_my && q < total_q; ++i) {
          int nt, b, tt;
          decode_item((int)blockIdx.x + i * (int)gridDim.x, ta.div_t, ta.div_b, nt, b, tt);
          const int t0 = tt * 128 - a.pad;
          for (int ch = 0; ch < nchunks; ++ch, ++q) {
            mbar_wait(&hdr->a_empty[as], ph ^ 1);
            mbar_arrive_expect_tx(&hdr->a_full[as], a_stage);
            tma_load_4d(a_smem + (size_t)as * a_stage, &tmap, 0, t0, b * cgs + cg0 + ch, 0, &hdr->a_full[as]);
            if (++as == na) as = 0, ph ^= 1;
          }
        }
      }
      __syncwarp();
    } else {
      // ------------------------------------------------ activation producers (2 groups x 128 threads)
      // Group g stages chunks q = g, g+2, ... of this CTA's chunk sequence (q = tile * nchunks + chunk).
      // One task = one time row x all KC channels of the chunk: 32 independent coalesced loads in
      // flight per thread, then leaky_relu / mask / fp16 hi-lo split and 2 x 4 conflict-free 16 B stores.
      griddep_wait();
      const int g = (warp - 2) >> 2;
      const int rp = tid - 64 - PROD_GROUP * g;
      const float slope = a.pre_slope;
      const int na_shift = na == 4 ? 2 : 1;
      const int total_q = n_my * nchunks;
      for (int q = g; q < total_q; q += PROD_GROUPS) {
        const int i = q / nchunks, ch = q - i * nchunks;
        int nt, b, tt;
        decode_item((int)blockIdx.x + i * (int)gridDim.x, ta.div_t, ta.div_b, nt, b, tt);
        const int t0 = tt * 128;
        const int as = q & (na - 1);
        mbar_wait(&hdr->a_empty[as], ((uint32_t)(q >> na_shift) & 1u) ^ 1u);
        uint4* Ahi = reinterpret_cast<uint4*>(a_smem + (size_t)as * a_stage);  // row r = 4 x 16 B, chunk kg at kg ^ ((r >> 1) & 3)
        uint4* Alo = Ahi + KG * rows;
        const float* mrow = a.in_mask ? a.in_mask + (size_t)b * a.mask_stride : nullptr;
        const float* xb = a.x + ((size_t)b * a.x_C + a.x_ch_off + ch * KC) * a.x_stride;
        for (int r = rp; r < rows; r += PROD_GROUP) {
          const int t = t0 - a.pad + r;
          const bool ok = t >= 0 && t < a.Lin;
          float v[KC];
  #pragma unroll
          for (int c = 0; c < KC; ++c) v[c] = ok ? __ldg(xb + (size_t)c * a.x_stride + t) : 0.f;
          const float mk = (ok && mrow) ? __ldg(mrow + t) : 1.0f;
  #pragma unroll
          for (int c = 0; c < KC; ++c) {
            float qv = v[c];
            qv = qv > 0.f ? qv : qv * slope;
            v[c] = mrow ? qv * mk : qv;
          }
            // 64 B swizzle: chunk position = chunk ^ (address bits 7-8 of the row); both planes start 512 B-aligned
          const int swz = (int)(((smem_u32(Ahi) + (uint32_t)r * 64u) >> 7) & 3u);
          if (planes == 2) {
#pragma unroll
            for (int kg = 0; kg < KG; ++kg) {
              uint4 h, l;
              split2(v[kg * 8 + 0], v[kg * 8 + 1], h.x, l.x);
              split2(v[kg * 8 + 2], v[kg * 8 + 3], h.y, l.y);
              split2(v[kg * 8 + 4], v[kg * 8 + 5], h.z, l.z);
              split2(v[kg * 8 + 6], v[kg * 8 + 7], h.w, l.w);
              Ahi[r * KG + (kg ^ swz)] = h;
              Alo[r * KG + (kg ^ swz)] = l;
            }
          } else {
#pragma unroll
            for (int kg = 0; kg < KG; ++kg) Ahi[r * KG + (kg ^ swz)] = pack_bf16x8(&v[kg * 8]);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&hdr->a_full[as]);
      }
    }
  } else {
    // ------------------------------------------------ epilogue: TMEM -> registers -> global
    // warp w owns TMEM lane quarter w & 3 (= 32 time rows); the warps of a quarter split the tile's
    // 16-column chunks between them.  Everything that varies per element is branch-free: operand loads
    // use clamped (always valid) addresses so 16-32 of them are in flight per thread, the operands of
    // the NEXT chunk -- and of the next tile's first chunk -- are requested before the current chunk
    // is finished, and only the final store is predicated.
    // Narrow layers (several accumulator stages available) instead give whole tiles to alternating warp
    // groups, so the per-tile fixed latency (barrier wait, TMEM load, address set-up) of one group
    // overlaps the other group's tile.
    griddep_wait();
    const int q4 = warp & 3;
    const int egroups = ta.epi_groups;                 // warp groups that take alternate tiles
    const int esplit = kEpiSplit / egroups;            // warps of one lane quarter sharing a tile's columns
    const int eq = (warp - kFirstEpi) >> 2;            // 0 .. kEpiSplit-1
    const int egroup = eq % egroups, part = eq / egroups;
    const int row = q4 * 32 + lane;
    const float unscale = ta.unscale;
    const int mode = a.mode;

    // STORE-mode addressing of the 16-channel chunk starting at virtual channel o0 (never straddles a.split)
    struct ChunkIO {
      const float* res;
      const float* acc;
      float* y;
      uint16_t* sp;  // operand-image position of (hi plane, first channel group of the chunk, row t), or null
      float sp_slope;
      const uint16_t* rimg;  // residual image position (hi plane), or null
      float r_inv;           // 1 / res_slope
      size_t plane;          // halves between the hi and lo planes of the destination's images
      ptrdiff_t step;
      int um, nvalid;
    };
    const size_t sp_plane = (size_t)a.B * a.y_stride;  // x channels: halves between the hi and lo planes
    auto chunk_io = [&](int b, int o0, int t, int tl) {
      ChunkIO io;
      const int s1 = o0 >= a.split ? 1 : 0;
      const EpiDesc& d = a.e[s1];
      const int rel0 = s1 ? o0 - a.split : o0;
      const size_t off = ((size_t)b * d.C + d.ch_off + d.ch_sign * rel0) * a.y_stride;
      io.res = d.res ? d.res + off + tl : nullptr;
      io.acc = d.acc_in ? d.acc_in + off + tl : nullptr;
      io.y = d.y ? d.y + off + t : nullptr;
      const int dch0 = d.ch_off + rel0;  // first destination channel of the chunk (multiple of 16 when images are used)
      const size_t cell = (((size_t)b * (d.C >> 5) + (dch0 >> 5)) * a.y_stride) * 32 + (dch0 & 31);  // + t * 32
      io.sp = d.split ? d.split + cell + (size_t)t * 32 : nullptr;
      io.sp_slope = d.split_slope;
      io.rimg = d.res_img ? d.res_img + cell + (size_t)tl * 32 : nullptr;
      io.r_inv = 1.0f / d.res_slope;
      io.plane = sp_plane * (size_t)d.C;
      io.step = (ptrdiff_t)d.ch_sign * a.y_stride;
      io.um = d.use_mask;
      io.nvalid = max(0, min(16, a.Cout - o0));
      return io;
    };
    auto load16 = [&](const ChunkIO& io, float (&q)[16]) {
      if (io.nvalid == 16) {
        if (io.rimg) {
          // residual from its operand image: 4 x 16 B loads, r = hi + lo with the leaky_relu inverted
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            const uint4 hq = *reinterpret_cast<const uint4*>(io.rimg + g8 * 8);
            const uint4 lq = *reinterpret_cast<const uint4*>(io.rimg + g8 * 8 + io.plane);
            const uint32_t hw[4] = {hq.x, hq.y, hq.z, hq.w}, lw[4] = {lq.x, lq.y, lq.z, lq.w};
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
              const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[e2]));
              const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[e2]));
              const float v0 = fmaf(lf.x, LO_INV, hf.x), v1 = fmaf(lf.y, LO_INV, hf.y);
              q[8 * g8 + 2 * e2] = v0 >= 0.f ? v0 : v0 * io.r_inv;
              q[8 * g8 + 2 * e2 + 1] = v1 >= 0.f ? v1 : v1 * io.r_inv;
            }
          }
        } else if (io.res) {
#pragma unroll
          for (int e = 0; e < 16; ++e) q[e] = io.res[e * io.step];
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) q[e] = 0.f;
        }
        if (io.acc) {
#pragma unroll
          for (int e = 0; e < 16; ++e) q[e] += io.acc[e * io.step];
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float u1 = (io.res && e < io.nvalid) ? io.res[e * io.step] : 0.f;
          const float u2 = (io.acc && e < io.nvalid) ? io.acc[e * io.step] : 0.f;
          q[e] = u1 + u2;
        }
      }
    };

    const int nch = N >> 4, hc = (nch + esplit - 1) / esplit;
    const int n_lo = part * hc * 16, n_hi = min(nch, (part + 1) * hc) * 16;  // STORE / SHUFFLE column range
    const bool have_cols = n_lo < n_hi;

    // Residual / running-sum operands are requested TWO chunk-jobs ahead of their use (a job = one
    // 16-column chunk of one tile of this warp): on narrow layers one job is much shorter than the
    // HBM latency.  (la_i, la_n0) is the look-ahead cursor over this warp's job sequence.
    float r1[16], r2[16];
    int la_i = egroup, la_n0 = n_lo;
    auto la_load = [&](float (&q)[16]) {
      if (la_i < n_my) {
        int nt2, b2, tt2;
        decode_item((int)blockIdx.x + la_i * (int)gridDim.x, ta.div_t, ta.div_b, nt2, b2, tt2);
        const int t2 = tt2 * 128 + row;
        load16(chunk_io(b2, nt2 * N + la_n0, t2, min(t2, a.Lout - 1)), q);
        la_n0 += 16;
        if (la_n0 >= n_hi) la_n0 = n_lo, la_i += egroups;
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) q[e] = 0.f;
      }
    };
    if (mode == MODE_STORE && have_cols && !(kTma && ta.epi_fast)) {
      la_load(r1);
      la_load(r2);
    }

    // ---- lean STORE epilogue (ta.epi_fast, chosen by launch_conv_tc): the decoder ResBlock convs -- one destination
    // side, fp16 hi/lo planes, operand image out, optional fp32 residual in / fp32 tensor out, nothing else.  On the
    // narrow stages the kernel is bound by how many instructions the epilogue warps issue per 16-column job (ncu:
    // 440-620 in the generic code below, 76 % of all instructions), so this path keeps one address per tile (32-bit
    // element offsets from the tensor bases), decodes one item per tile, ping-pongs two residual buffers instead of
    // rotating three, and has no per-element predicates except the row bound.  Same arithmetic, same order.
    auto store_fast = [&](auto res_tag, auto y_tag, auto acc_tag, auto spl_tag) {
      // RES: 0 = no residual, 1 = fp32 tensor, 2 = operand image (hi + lo * 2^-11, leaky_relu inverted)
      constexpr int RES = decltype(res_tag)::value;
      constexpr bool YOUT = decltype(y_tag)::value, ACC = decltype(acc_tag)::value, SPL = decltype(spl_tag)::value;
      const EpiDesc& d = a.e[0];
      const float* resb = d.res;  // may alias yb (in-place residual update: each element is read, then written, by one thread)
      const uint16_t* rimg = d.res_img;
      const float r_inv = RES == 2 ? 1.0f / d.res_slope : 1.0f;
      const float* accb = d.acc_in;  // running sum over the stage's ResBlocks (models.py:150-155), may alias yb too
      const float post_div = a.post_div;
      float* yb = d.y;
      uint16_t* spb = d.split;
      const uint32_t ystride = (uint32_t)a.y_stride;
      const uint32_t lo_plane = (uint32_t)a.B * (uint32_t)d.C * ystride;  // halves between the hi and lo planes
      const float slope = d.split_slope;
      const int J = (n_hi - n_lo) >> 4;                                    // jobs per tile of this warp (even)
      const uint32_t cgroups = (uint32_t)d.C >> 5;
      float rA[16], rB[16];
      // ACC: the running-sum operand of the job after next is requested together with its residual into `ab` and
      // folded into that residual buffer one job later (r = res + acc, as the generic path adds them)
      float ab[16];
      bool ab_pending = false;
      // element offset of (first channel of this warp's columns, row) for item i; false when i is past the end
      auto tile_off = [&](int i, uint32_t& off_clamped, uint32_t& off_row, int& t_out, int& b_out, int& ch_out) -> bool {
        if (i >= n_my) return false;
        int nt2, b2, tt2;
        decode_item((int)blockIdx.x + i * (int)gridDim.x, ta.div_t, ta.div_b, nt2, b2, tt2);
        const int t2 = tt2 * 128 + row;
        const int ch = d.ch_off + nt2 * N + n_lo;
        const uint32_t base = ((uint32_t)b2 * (uint32_t)d.C + (uint32_t)ch) * ystride;
        off_clamped = base + (uint32_t)min(t2, a.Lout - 1);
        off_row = base + (uint32_t)t2;
        t_out = t2, b_out = b2, ch_out = ch;
        return true;
      };
      // residual of the 16 channels starting at channel `ch` of utterance `bb`, row min(tt, Lout - 1); `off` = the same
      // position as an fp32 element offset
      // RES == 2: the raw fp16 words of the residual image travel in the buffer (r[8 g8 + 0..3] = hi, + 4..7 = lo) and
      // become values (x = hi + lo * 2^-11, leaky_relu inverted) in decode_res, called where the buffer is next touched --
      // one or two jobs later.  Converting at load time parked the warp on loads it had just issued: ncu on a C = 64 conv2
      // (profiles/r2d_ncu_c64_k7_conv2.txt) had 30 % of all samples on the first conversion instructions.
      const bool lazy = ta.lazy_res != 0;  // $SVK_LAZY_RES=0: convert at load time (A/B)
      auto decode_res = [&](float (&r)[16]) {
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          float o[8];
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const uint32_t hw = __float_as_uint(r[8 * g8 + e2]), lw = __float_as_uint(r[8 * g8 + 4 + e2]);
            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw));
            const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw));
            const float v0 = fmaf(lf.x, LO_INV, hf.x), v1 = fmaf(lf.y, LO_INV, hf.y);
            o[2 * e2] = v0 >= 0.f ? v0 : v0 * r_inv;
            o[2 * e2 + 1] = v1 >= 0.f ? v1 : v1 * r_inv;
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) r[8 * g8 + e] = o[e];
        }
      };
      auto load_res = [&](float (&r)[16], uint32_t off, int bb, int ch, int tt) {
        if (RES == 2) {
          const uint32_t tl = (uint32_t)min(tt, a.Lout - 1);
          const uint16_t* rp = rimg + (((uint32_t)bb * cgroups + ((uint32_t)ch >> 5)) * ystride + tl) * 32u + ((uint32_t)ch & 31u);
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            const uint4 hq = *reinterpret_cast<const uint4*>(rp + g8 * 8);
            const uint4 lq = *reinterpret_cast<const uint4*>(rp + g8 * 8 + lo_plane);
            r[8 * g8 + 0] = __uint_as_float(hq.x), r[8 * g8 + 1] = __uint_as_float(hq.y);
            r[8 * g8 + 2] = __uint_as_float(hq.z), r[8 * g8 + 3] = __uint_as_float(hq.w);
            r[8 * g8 + 4] = __uint_as_float(lq.x), r[8 * g8 + 5] = __uint_as_float(lq.y);
            r[8 * g8 + 6] = __uint_as_float(lq.z), r[8 * g8 + 7] = __uint_as_float(lq.w);
          }
          if (!lazy) decode_res(r);
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) r[e] = resb[off + (uint32_t)e * ystride];
        }
      };
      uint32_t offc = 0, offr = 0, n_offc = 0, n_offr = 0;
      int t = 0, b = 0, ch0 = 0, n_t = 0, n_b = 0, n_ch0 = 0;
      bool have = tile_off(egroup, offc, offr, t, b, ch0);
      if (RES && have) {
        load_res(rA, offc, b, ch0, t);
        load_res(rB, offc + 16u * ystride, b, ch0 + 16, t);
        if (ACC) {
          if (RES == 2 && lazy) decode_res(rA), decode_res(rB);  // with a running sum the buffers hold VALUES from here on
#pragma unroll
          for (int e = 0; e < 16; ++e) rA[e] += accb[offc + (uint32_t)e * ystride], rB[e] += accb[offc + (uint32_t)(16 + e) * ystride];
        }
      }
      for (int i = egroup; have; i += egroups) {
        const int s = i & (nacc - 1);
        const bool have_next = tile_off(i + egroups, n_offc, n_offr, n_t, n_b, n_ch0);
        const bool tin = t < a.Lout;
        mbar_wait(&hdr->acc_full[s], (uint32_t)(i >> nacc_shift) & 1u);
        tc_fence_after();
        const uint32_t tsub = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(s * acc_cols) + (uint32_t)n_lo;
        const float* bias_t = bias_s + (ch0 - d.ch_off);
        auto job = [&](int j, float (&r)[16], float (&r_other)[16]) {
          if (ACC && ab_pending) {  // the other buffer was refilled during the previous job: complete r = res + acc
            if (RES == 2 && lazy) decode_res(r_other);
#pragma unroll
            for (int e = 0; e < 16; ++e) r_other[e] += ab[e];
            ab_pending = false;
          }
          uint32_t m[16], c[16];
          tmem_ld16(tsub + (uint32_t)(16 * j), m);
          if (planes == 2) {
            tmem_ld16(tsub + (uint32_t)(N + 16 * j), c);
          } else {  // single bf16 plane: no cross accumulator
#pragma unroll
            for (int e = 0; e < 16; ++e) c[e] = 0u;
          }
          const float4* b4 = reinterpret_cast<const float4*>(bias_t + 16 * j);
          tmem_wait_ld();
          if (j == J - 1) {  // every tcgen05.ld of this stage has completed: hand it back before the stores
            tc_fence_before();
            mbar_arrive(&hdr->acc_empty[s]);
          }
          float v[16];
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 q = b4[e4];
            v[4 * e4 + 0] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 0]), LO_INV, __uint_as_float(m[4 * e4 + 0])), unscale, q.x);
            v[4 * e4 + 1] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 1]), LO_INV, __uint_as_float(m[4 * e4 + 1])), unscale, q.y);
            v[4 * e4 + 2] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 2]), LO_INV, __uint_as_float(m[4 * e4 + 2])), unscale, q.z);
            v[4 * e4 + 3] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 3]), LO_INV, __uint_as_float(m[4 * e4 + 3])), unscale, q.w);
          }
          if (RES) {
            if (RES == 2 && !ACC && lazy) decode_res(r);  // (with ACC it was decoded when the running sum was folded in)
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] += r[e];
            // refill this buffer with the residual of the job after next (same tile, or the group's next tile)
            const int jn = j + 2;
            if (jn < J || have_next) {
              const bool same = jn < J;
              const uint32_t off = same ? offc + (uint32_t)(16 * jn) * ystride : n_offc + (uint32_t)(16 * (jn - J)) * ystride;
              load_res(r, off, same ? b : n_b, same ? ch0 + 16 * jn : n_ch0 + 16 * (jn - J), same ? t : n_t);
              if (ACC) {
#pragma unroll
                for (int e = 0; e < 16; ++e) ab[e] = accb[off + (uint32_t)e * ystride];
                ab_pending = true;
              }
            }
          }
          if (post_div != 1.0f) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = v[e] / post_div;
          }
          if (tin) {
            if (YOUT) {
              float* yp = yb + offr + (uint32_t)(16 * j) * ystride;
#pragma unroll
              for (int e = 0; e < 16; ++e) yp[(uint32_t)e * ystride] = v[e];
            }
            if (SPL) {
              // leaky_relu(y) as fp16 hi/lo: 16 channels = two 16 B pieces of the 64 B image row, per plane
              const uint32_t dch = (uint32_t)(ch0 + 16 * j);
              uint16_t* sp = spb + (((uint32_t)b * cgroups + (dch >> 5)) * ystride + (uint32_t)t) * 32u + (dch & 31u);
              uint4 h[2], l[2];
#pragma unroll
              for (int g8 = 0; g8 < 2; ++g8) {
                float w8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) w8[e] = v[8 * g8 + e] > 0.f ? v[8 * g8 + e] : v[8 * g8 + e] * slope;
                if (planes == 2) {
                  split2(w8[0], w8[1], h[g8].x, l[g8].x);
                  split2(w8[2], w8[3], h[g8].y, l[g8].y);
                  split2(w8[4], w8[5], h[g8].z, l[g8].z);
                  split2(w8[6], w8[7], h[g8].w, l[g8].w);
                } else {
                  h[g8] = pack_bf16x8(w8);
                }
              }
              st_global_v8(sp, h[0], h[1]);  // 16 channels = one 32 B sector per plane, one request each
              if (planes == 2) st_global_v8(sp + lo_plane, l[0], l[1]);
            }
          }
        };
        for (int j = 0; j < J; j += 2) {
          job(j, rA, rB);
          job(j + 1, rB, rA);
        }
        have = have_next;
        offc = n_offc, offr = n_offr, t = n_t, b = n_b, ch0 = n_ch0;
      }
    };
    const bool epi_fast = kTma && ta.epi_fast;  // only image-fed launches qualify: keep the converting variant lean
    if constexpr (kTma) if (epi_fast) {
      using I0 = std::integral_constant<int, 0>;
      using I1 = std::integral_constant<int, 1>;
      using I2 = std::integral_constant<int, 2>;
      const EpiDesc& d0 = a.e[0];
      const bool y_ = d0.y != nullptr, acc_ = d0.acc_in != nullptr, spl_ = d0.split != nullptr;
      if (!d0.res && !d0.res_img) {
        store_fast(I0{}, std::false_type{}, std::false_type{}, std::true_type{});  // conv1: image -> image
      } else if (!y_) {  // residual stream kept as images only, or conv2 of a narrow pair without fp32 out
        if (d0.res_img && acc_) store_fast(I2{}, std::false_type{}, std::true_type{}, std::true_type{});  // a stage's last conv: + running sum, image only
        else if (d0.res_img) store_fast(I2{}, std::false_type{}, std::false_type{}, std::true_type{});
        else store_fast(I1{}, std::false_type{}, std::false_type{}, std::true_type{});
      } else if (d0.res_img) {
        if (acc_) {
          if (spl_) store_fast(I2{}, std::true_type{}, std::true_type{}, std::true_type{});
          else store_fast(I2{}, std::true_type{}, std::true_type{}, std::false_type{});
        } else {
          if (spl_) store_fast(I2{}, std::true_type{}, std::false_type{}, std::true_type{});
          else store_fast(I2{}, std::true_type{}, std::false_type{}, std::false_type{});
        }
      } else {
        if (acc_) {
          if (spl_) store_fast(I1{}, std::true_type{}, std::true_type{}, std::true_type{});
          else store_fast(I1{}, std::true_type{}, std::true_type{}, std::false_type{});
        } else {
          if (spl_) store_fast(I1{}, std::true_type{}, std::false_type{}, std::true_type{});
          else store_fast(I1{}, std::true_type{}, std::false_type{}, std::false_type{});
        }
      }
    }

    for (int i = egroup; i < n_my && !epi_fast; i += egroups) {
      const int s = i & (nacc - 1);
      int ntile, b, tt;
      decode_item((int)blockIdx.x + i * (int)gridDim.x, ta.div_t, ta.div_b, ntile, b, tt);
      const int t = tt * 128 + row;
      const bool tin = t < a.Lout;
      const int tl = min(t, a.Lout - 1);
      const float* omask = a.out_mask ? a.out_mask + (size_t)b * a.mask_stride : nullptr;
      const float mv = omask ? omask[tl] : 1.0f;
      const int o_tile = ntile * N;
      mbar_wait(&hdr->acc_full[s], (uint32_t)(i >> nacc_shift) & 1u);
      tc_fence_after();
      const uint32_t tsub = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(s * acc_cols);

      if (mode == MODE_STORE) {
        // y = ((acc/s + bias) + (res + acc_in)) / post_div * mask, tanh  (same element is read then
        // written by the same thread only, so in-place operation is safe)
        for (int n0 = n_lo; n0 < n_hi; n0 += 16) {
          // request the operands of the job after next
          float p1[16];
          la_load(p1);

          uint32_t m[16], c[16];
          tmem_ld16(tsub + (uint32_t)n0, m);
          if (planes == 2) {
            tmem_ld16(tsub + (uint32_t)(N + n0), c);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) c[e] = 0u;
          }
          const float4* b4 = reinterpret_cast<const float4*>(bias_s + o_tile + n0);
          const ChunkIO io_c = chunk_io(b, o_tile + n0, t, tl);
          tmem_wait_ld();
          float v[16];
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 q = b4[e4];
            v[4 * e4 + 0] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 0]), LO_INV, __uint_as_float(m[4 * e4 + 0])), unscale, q.x) + r1[4 * e4 + 0];
            v[4 * e4 + 1] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 1]), LO_INV, __uint_as_float(m[4 * e4 + 1])), unscale, q.y) + r1[4 * e4 + 1];
            v[4 * e4 + 2] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 2]), LO_INV, __uint_as_float(m[4 * e4 + 2])), unscale, q.z) + r1[4 * e4 + 2];
            v[4 * e4 + 3] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 3]), LO_INV, __uint_as_float(m[4 * e4 + 3])), unscale, q.w) + r1[4 * e4 + 3];
          }
          if (a.post_div != 1.0f) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = v[e] / a.post_div;
          }
          if (io_c.um && omask) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] *= mv;
          }
          if (a.act_tanh) {
            bool bad = false;
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = tanhf(v[e]), bad = bad || !(fabsf(v[e]) <= 1.0f);
            if (bad && tin && a.range_flag) *a.range_flag = 1;
          }
          if (io_c.sp && tin) {
            // second output: leaky_relu(y) as fp16 hi/lo, 8 channels = one 16 B row of the operand image
            uint16_t* sp = io_c.sp;
            uint4 h[2], l[2];
#pragma unroll
            for (int g8 = 0; g8 < 2; ++g8) {
              float w8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) w8[e] = v[8 * g8 + e] > 0.f ? v[8 * g8 + e] : v[8 * g8 + e] * io_c.sp_slope;
              if (planes == 2) {
                split2(w8[0], w8[1], h[g8].x, l[g8].x);
                split2(w8[2], w8[3], h[g8].y, l[g8].y);
                split2(w8[4], w8[5], h[g8].z, l[g8].z);
                split2(w8[6], w8[7], h[g8].w, l[g8].w);
              } else {
                h[g8] = pack_bf16x8(w8);
              }
            }
            // 16 channels = one 32 B sector of the 64 B image row, written as one request per plane
            st_global_v8(sp, h[0], h[1]);
            if (planes == 2) st_global_v8(sp + sp_plane * (size_t)a.e[o_tile + n0 >= a.split ? 1 : 0].C, l[0], l[1]);
          }
          if (!io_c.y) {
            // operand image only
          } else if (io_c.nvalid == 16) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (tin) io_c.y[e * io_c.step] = v[e];
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (tin && e < io_c.nvalid) io_c.y[e * io_c.step] = v[e];
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) r1[e] = r2[e], r2[e] = p1[e];
        }
      } else if (mode == MODE_GATE) {
        // columns [0, N/2) hold the tanh half, [N/2, N) the sigmoid half of channels
        // ntile*N/2 + [0, N/2) (commons.py:100-107); bias is in the same virtual order.
        const int hN = N >> 1;
        const int npair = hN >> 4, hp = (npair + esplit - 1) / esplit;
        const int g_lo = part * hp * 16, g_hi = min(npair, (part + 1) * hp) * 16;
        const float* bptr = bias_s + o_tile;
        float* ybase = a.e[0].y ? a.e[0].y + ((size_t)b * a.e[0].C + a.e[0].ch_off + ntile * hN) * a.y_stride + t : nullptr;
        for (int n0 = g_lo; n0 < g_hi; n0 += 16) {
          uint32_t m[16], c[16];
          float g[16];
          tmem_ld16(tsub + (uint32_t)n0, m);
          if (planes == 2) {
            tmem_ld16(tsub + (uint32_t)(N + n0), c);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) c[e] = 0u;
          }
          tmem_wait_ld();
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 q = reinterpret_cast<const float4*>(bptr + n0)[e4];
            g[4 * e4 + 0] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 0]), LO_INV, __uint_as_float(m[4 * e4 + 0])), unscale, q.x);  // tanh side, pre-activation
            g[4 * e4 + 1] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 1]), LO_INV, __uint_as_float(m[4 * e4 + 1])), unscale, q.y);  // tanh side, pre-activation
            g[4 * e4 + 2] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 2]), LO_INV, __uint_as_float(m[4 * e4 + 2])), unscale, q.z);  // tanh side, pre-activation
            g[4 * e4 + 3] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 3]), LO_INV, __uint_as_float(m[4 * e4 + 3])), unscale, q.w);  // tanh side, pre-activation
          }
          tmem_ld16(tsub + (uint32_t)(hN + n0), m);
          if (planes == 2) tmem_ld16(tsub + (uint32_t)(N + hN + n0), c);
          tmem_wait_ld();
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 q = reinterpret_cast<const float4*>(bptr + hN + n0)[e4];
            g[4 * e4 + 0] = gate_tanh_sigmoid(g[4 * e4 + 0], fmaf(fmaf(__uint_as_float(c[4 * e4 + 0]), LO_INV, __uint_as_float(m[4 * e4 + 0])), unscale, q.x));
            g[4 * e4 + 1] = gate_tanh_sigmoid(g[4 * e4 + 1], fmaf(fmaf(__uint_as_float(c[4 * e4 + 1]), LO_INV, __uint_as_float(m[4 * e4 + 1])), unscale, q.y));
            g[4 * e4 + 2] = gate_tanh_sigmoid(g[4 * e4 + 2], fmaf(fmaf(__uint_as_float(c[4 * e4 + 2]), LO_INV, __uint_as_float(m[4 * e4 + 2])), unscale, q.z));
            g[4 * e4 + 3] = gate_tanh_sigmoid(g[4 * e4 + 3], fmaf(fmaf(__uint_as_float(c[4 * e4 + 3]), LO_INV, __uint_as_float(m[4 * e4 + 3])), unscale, q.w));
          }
          const int nval = max(0, min(16, (a.Cout >> 1) - (ntile * hN + n0)));
          if (a.e[0].split && tin && nval == 16) {
            // acts as the operand image of the res_skip conv (modules.py:169): no fp32 copy needed
            const int gch0 = a.e[0].ch_off + ntile * hN + n0;
            uint16_t* sp = a.e[0].split + (((size_t)b * (a.e[0].C >> 5) + (gch0 >> 5)) * a.y_stride + t) * 32 + (gch0 & 31);
            uint4 h[2], l[2];
#pragma unroll
            for (int g8 = 0; g8 < 2; ++g8) {
              if (planes == 2) {
                split2(g[8 * g8 + 0], g[8 * g8 + 1], h[g8].x, l[g8].x);
                split2(g[8 * g8 + 2], g[8 * g8 + 3], h[g8].y, l[g8].y);
                split2(g[8 * g8 + 4], g[8 * g8 + 5], h[g8].z, l[g8].z);
                split2(g[8 * g8 + 6], g[8 * g8 + 7], h[g8].w, l[g8].w);
              } else {
                h[g8] = pack_bf16x8(&g[8 * g8]);
              }
            }
            st_global_v8(sp, h[0], h[1]);
            if (planes == 2) st_global_v8(sp + sp_plane * (size_t)a.e[0].C, l[0], l[1]);
          }
          if (ybase) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (tin && e < nval) ybase[(size_t)(n0 + e) * a.y_stride] = g[e];
          }
        }
      } else {
        // MODE_SHUFFLE: virtual channel o' = co*s + r of time row q lands at y[co, s*q + r - p]
        // (ConvTranspose1d polyphase form).
        const int sh = a.shuf_s;
        const bool store_y = a.e[0].y != nullptr;  // the lean stride-2 path may write the operand image only
        float* ybase = a.e[0].y + ((size_t)b * a.e[0].C + a.e[0].ch_off) * a.y_stride;
        // Stride-2 upsamplers (k = 4, p = 1) with an operand image out -- the two launches that feed the narrow stages
        // : row predicates and addresses are computed once per tile, image rows are stored as full 32 B sectors.
        // Virtual channels (2 co, 2 co + 1) of row t are y[co, 2t - 1] and y[co, 2t]; y[co, 2t] is paired with the NEXT
        // row's y[co, 2t + 1] (one lane shuffle) -> one aligned 8 B store per lane and channel, 256 contiguous bytes
        // per warp instruction; the two ends of the warp's span are single floats.
        const bool lean2 = kTma && sh == 2 && a.shuf_p == 1 && a.Cout % 16 == 0 && (a.y_stride & 1) == 0 && planes == 2 && a.e[0].split;
        if constexpr (kTma) if (lean2) {
          const int tq = 2 * t;
          const bool p_first = tin && lane == 0 && tq >= 1;
          const bool p_pair = tin && lane < 31 && tq + 1 < a.shuf_Lout;
          const bool p_single = tin && !p_pair && tq < a.shuf_Lout;
          const bool i0 = tin && tq >= 1 && tq - 1 < a.shuf_Lout, i1 = tin && tq < a.shuf_Lout;
          const float sl = a.e[0].split_slope;
          const size_t lo_plane = sp_plane * (size_t)a.e[0].C;
          // Jobs are taken in pairs: a job holds 8 real channels (16 B of an image row) at two output steps; writing
          // them job by job leaves every 32 B sector half-written until the next job (measured: the image stores then
          // cost 0.46 of the launch's 0.68 ms).  The first job's packed rows are kept and stored together with the
          // second's: two adjacent 16 B stores = one full sector per row and plane.
          uint4 keep_h[2], keep_l[2];
          for (int n0 = n_lo; n0 < n_hi; n0 += 16) {
            const bool second = ((n0 - n_lo) & 16) != 0;
            uint32_t m[16], c[16];
            tmem_ld16(tsub + (uint32_t)n0, m);
            tmem_ld16(tsub + (uint32_t)(N + n0), c);
            const int o0 = o_tile + n0;
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + o0);
            tmem_wait_ld();
            float v[16];
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 q = b4[e4];
              v[4 * e4 + 0] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 0]), LO_INV, __uint_as_float(m[4 * e4 + 0])), unscale, q.x);
              v[4 * e4 + 1] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 1]), LO_INV, __uint_as_float(m[4 * e4 + 1])), unscale, q.y);
              v[4 * e4 + 2] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 2]), LO_INV, __uint_as_float(m[4 * e4 + 2])), unscale, q.z);
              v[4 * e4 + 3] = fmaf(fmaf(__uint_as_float(c[4 * e4 + 3]), LO_INV, __uint_as_float(m[4 * e4 + 3])), unscale, q.w);
            }
            if (store_y) {
              float* yr = ybase + (size_t)(o0 >> 1) * a.y_stride + tq;
#pragma unroll
              for (int kc = 0; kc < 8; ++kc) {
                const float nxt = __shfl_down_sync(0xffffffffu, v[2 * kc], 1);
                if (p_pair) *reinterpret_cast<float2*>(yr) = make_float2(v[2 * kc + 1], nxt);
                if (p_single) yr[0] = v[2 * kc + 1];
                if (p_first) yr[-1] = v[2 * kc];
                yr += a.y_stride;
              }
            }
            // leaky_relu(y) as the stage's operand image: 8 real channels at two output steps (16 B per step and plane)
            uint4 hq[2], lq[2];
#pragma unroll
            for (int r2 = 0; r2 < 2; ++r2) {
              float w8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) w8[e] = v[2 * e + r2] > 0.f ? v[2 * e + r2] : v[2 * e + r2] * sl;
              split2(w8[0], w8[1], hq[r2].x, lq[r2].x);
              split2(w8[2], w8[3], hq[r2].y, lq[r2].y);
              split2(w8[4], w8[5], hq[r2].z, lq[r2].z);
              split2(w8[6], w8[7], hq[r2].w, lq[r2].w);
            }
            if (!second && n0 + 16 < n_hi) {
              keep_h[0] = hq[0], keep_h[1] = hq[1], keep_l[0] = lq[0], keep_l[1] = lq[1];
            } else {
              const bool paired = second;  // false only for a trailing single job
              const int co0 = a.e[0].ch_off + ((paired ? o0 - 16 : o0) >> 1);
              uint16_t* sp = a.e[0].split + (((size_t)b * (a.e[0].C >> 5) + (co0 >> 5)) * a.y_stride + (tq - 1)) * 32 + (co0 & 31);
#pragma unroll
              for (int r2 = 0; r2 < 2; ++r2) {
                if (r2 ? i1 : i0) {
                  if (paired) {
                    st_global_v8(sp + 32 * r2, keep_h[r2], hq[r2]);
                    st_global_v8(sp + 32 * r2 + lo_plane, keep_l[r2], lq[r2]);
                  } else {
                    *reinterpret_cast<uint4*>(sp + 32 * r2) = hq[r2];
                    *reinterpret_cast<uint4*>(sp + 32 * r2 + lo_plane) = lq[r2];
                  }
                }
              }
            }
          }
        }
        for (int n0 = n_lo; n0 < n_hi && !lean2; n0 += 16) {
          uint32_t m[16], c[16];
          tmem_ld16(tsub + (uint32_t)n0, m);
          if (planes == 2) {
            tmem_ld16(tsub + (uint32_t)(N + n0), c);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) c[e] = 0u;
          }
          const int o0 = o_tile + n0;
          float bv[16];
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 q = reinterpret_cast<const float4*>(bias_s + o0)[e4];
            bv[4 * e4] = q.x, bv[4 * e4 + 1] = q.y, bv[4 * e4 + 2] = q.z, bv[4 * e4 + 3] = q.w;
          }
          tmem_wait_ld();
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = fmaf(fmaf(__uint_as_float(c[e]), LO_INV, __uint_as_float(m[e])), unscale, bv[e]);
          // Stride-8 upsamplers (k = 16, p = 4): row t holds y[co, 8t-4 .. 8t+3] for two real channels.  Its upper four
          // floats are the lower half of the 32 B sector y[co, 8t .. 8t+7]; the upper half comes from the NEXT row's lower
          // four (lane shuffles), so a lane stores one whole sector with one 32 B store.  Lane 0 also stores its own lower
          // four (the previous warp's lane 31 could not), lane 31 stores only its half.
          if (a.shuf_rmajor) {
            // r-major virtual channels (stride-8 upsamplers feeding image-only ResBlocks): this job = 16 real channels
            // co0 .. co0+15 of output step sh*t + r - p -> one 32 B sector per plane of the stage's operand image,
            // written from here instead of by split_image_kernel from an fp32 tensor nobody else reads
            const int Cr = a.Cout / sh, r = o0 / Cr, co0 = o0 - r * Cr;
            const int tq = sh * t + r - a.shuf_p;
            if (tin && o0 < a.Cout && tq >= 0 && tq < a.shuf_Lout) {
              if (a.e[0].y) {
                float* yp = ybase + (size_t)co0 * a.y_stride + tq;
#pragma unroll
                for (int e = 0; e < 16; ++e) yp[(size_t)e * a.y_stride] = v[e];
              }
              if (a.e[0].split) {
                const int dch = a.e[0].ch_off + co0;
                uint16_t* sp = a.e[0].split + (((size_t)b * (a.e[0].C >> 5) + (dch >> 5)) * a.y_stride + tq) * 32 + (dch & 31);
                const float sl = a.e[0].split_slope;
                uint4 hq[2], lq[2];
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) {
                  float w8[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) w8[e] = v[8 * g8 + e] > 0.f ? v[8 * g8 + e] : v[8 * g8 + e] * sl;
                  if (planes == 2) {
                    split2(w8[0], w8[1], hq[g8].x, lq[g8].x);
                    split2(w8[2], w8[3], hq[g8].y, lq[g8].y);
                    split2(w8[4], w8[5], hq[g8].z, lq[g8].z);
                    split2(w8[6], w8[7], hq[g8].w, lq[g8].w);
                  } else {
                    hq[g8] = pack_bf16x8(w8);
                  }
                }
                st_global_v8(sp, hq[0], hq[1]);
                if (planes == 2) st_global_v8(sp + sp_plane * (size_t)a.e[0].C, lq[0], lq[1]);
              }
            }
            continue;
          }
          const bool sect8 = sh == 8 && a.shuf_p == 4 && o0 + 16 <= a.Cout && (a.y_stride & 7) == 0;
          if (sect8) {
#pragma unroll
            for (int gq = 0; gq < 2; ++gq) {
              float nl[4];
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2) nl[k2] = __shfl_down_sync(0xffffffffu, v[8 * gq + k2], 1);
              if (tin) {
                float* yrow = ybase + (size_t)((o0 >> 3) + gq) * a.y_stride;
                const int t8 = 8 * t;
                if (lane == 0 && t8 >= 4)
                  *reinterpret_cast<float4*>(yrow + t8 - 4) = make_float4(v[8 * gq], v[8 * gq + 1], v[8 * gq + 2], v[8 * gq + 3]);
                if (lane < 31 && t8 + 7 < a.shuf_Lout) {
                  const uint4 lo4 = make_uint4(__float_as_uint(v[8 * gq + 4]), __float_as_uint(v[8 * gq + 5]),
                                               __float_as_uint(v[8 * gq + 6]), __float_as_uint(v[8 * gq + 7]));
                  const uint4 hi4 = make_uint4(__float_as_uint(nl[0]), __float_as_uint(nl[1]), __float_as_uint(nl[2]), __float_as_uint(nl[3]));
                  st_global_v8(yrow + t8, lo4, hi4);
                } else if (t8 + 3 < a.shuf_Lout) {
                  *reinterpret_cast<float4*>(yrow + t8) = make_float4(v[8 * gq + 4], v[8 * gq + 5], v[8 * gq + 6], v[8 * gq + 7]);
                }
              }
            }
          }
          if (!tin) continue;
          if (sect8) {
            // stored above
          } else if (sh == 8 && (a.shuf_p & 3) == 0) {
#pragma unroll
            for (int gq = 0; gq < 2; ++gq) {
              const int co = (o0 >> 3) + gq;
              if (o0 + 8 * gq >= a.Cout) break;
              float* yrow = ybase + (size_t)co * a.y_stride;
              const int tq = 8 * t - a.shuf_p;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int th = tq + 4 * h;
                if (th >= 0 && th + 3 < a.shuf_Lout) {
                  *reinterpret_cast<float4*>(yrow + th) =
                      make_float4(v[8 * gq + 4 * h], v[8 * gq + 4 * h + 1], v[8 * gq + 4 * h + 2], v[8 * gq + 4 * h + 3]);
                } else {
#pragma unroll
                  for (int k2 = 0; k2 < 4; ++k2)
                    if (th + k2 >= 0 && th + k2 < a.shuf_Lout) yrow[th + k2] = v[8 * gq + 4 * h + k2];
                }
              }
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int o = o0 + e;
              const int co = o / sh, r = o - co * sh;
              const int tq = sh * t + r - a.shuf_p;
              if (o < a.Cout && tq >= 0 && tq < a.shuf_Lout) ybase[(size_t)co * a.y_stride + tq] = v[e];
            }
          }
          if (a.e[0].split && sh == 2 && o0 + 16 <= a.Cout) {
            // stride-2 upsamplers: this thread holds 8 consecutive real channels at two output steps -> it also
            // writes the operand image of leaky_relu(y) that the stage's ResBlocks read (16 B per step and plane)
            const int co0 = a.e[0].ch_off + (o0 >> 1);
#pragma unroll
            for (int r2 = 0; r2 < 2; ++r2) {
              const int tq = 2 * t + r2 - a.shuf_p;
              if (tq < 0 || tq >= a.shuf_Lout) continue;
              float w8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float q = v[2 * e + r2];
                w8[e] = q > 0.f ? q : q * a.e[0].split_slope;
              }
              uint16_t* sp = a.e[0].split + (((size_t)b * (a.e[0].C >> 5) + (co0 >> 5)) * a.y_stride + tq) * 32 + (co0 & 31);
              if (planes == 2) {
                uint4 hq, lq;
                split2(w8[0], w8[1], hq.x, lq.x);
                split2(w8[2], w8[3], hq.y, lq.y);
                split2(w8[4], w8[5], hq.z, lq.z);
                split2(w8[6], w8[7], hq.w, lq.w);
                *reinterpret_cast<uint4*>(sp) = hq;
                *reinterpret_cast<uint4*>(sp + sp_plane * (size_t)a.e[0].C) = lq;
              } else {
                *reinterpret_cast<uint4*>(sp) = pack_bf16x8(w8);
              }
            }
          }
        }
      }
      // every tcgen05.ld of this stage has completed (wait::ld above): hand the stage back
      tc_fence_before();
      mbar_arrive(&hdr->acc_empty[s]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ta.tmem_cols);
}

// fp32 [B, C, L] -> operand image [hi|lo][B][C/32][L][32] of leaky_relu(x, slope): one thread = one
// (32-channel group, time) cell: 32 coalesced channel-row reads, one 64 B row per plane written.
__global__ void __launch_bounds__(256) split_image_kernel(const float* __restrict__ x, int B, int C, int L, float slope,
                                                          uint16_t* __restrict__ img, int planes) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int cg = blockIdx.y, b = blockIdx.z;
  if (t >= L) return;
  const float* xr = x + ((size_t)b * C + cg * 32) * L + t;
  float v[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    const float q = __ldg(xr + (size_t)e * L);
    v[e] = q > 0.f ? q : q * slope;
  }
  uint16_t* cell = img + (((size_t)b * (C >> 5) + cg) * L + t) * 32;
  uint4 h[4], l[4];
#pragma unroll
  for (int g8 = 0; g8 < 4; ++g8) {
    if (planes == 1) {
      h[g8] = pack_bf16x8(&v[g8 * 8]);
    } else {
      split2(v[g8 * 8 + 0], v[g8 * 8 + 1], h[g8].x, l[g8].x);
      split2(v[g8 * 8 + 2], v[g8 * 8 + 3], h[g8].y, l[g8].y);
      split2(v[g8 * 8 + 4], v[g8 * 8 + 5], h[g8].z, l[g8].z);
      split2(v[g8 * 8 + 6], v[g8 * 8 + 7], h[g8].w, l[g8].w);
    }
  }
  st_global_v8(cell, h[0], h[1]);  // the 64 B row as two full sectors
  st_global_v8(cell + 16, h[2], h[3]);
  if (planes == 2) {
    uint16_t* lo = cell + (size_t)B * C * L;
    st_global_v8(lo, l[0], l[1]);
    st_global_v8(lo + 16, l[2], l[3]);
  }
}

}  // namespace

// ------------------------------------------------------------------------------- host helpers
int conv_tc_rows(int K, int dil) { return (128 + (K - 1) * dil + 7) & ~7; }

size_t conv_tc_packed_halves(int Cin, int Cout, int K, int N, int planes) {
  const int ntiles = (Cout + N - 1) / N, nchunks = Cin / KC;
  return (size_t)ntiles * nchunks * K * N * KC * planes;
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

// wv(o, c, j) is the logical fp32 weight; image = [ntile][chunk][tap][kgroup][hi n | lo n][8].
// planes == 1: the same image with a single bf16 plane (scale must be 1).
void conv_tc_pack(const float* w_ock, int Cout, int Cin, int K, int N, float scale, uint16_t* out, int planes) {
  const int ntiles = (Cout + N - 1) / N, nchunks = Cin / KC;
  size_t idx = 0;
  for (int nt = 0; nt < ntiles; ++nt)
    for (int ch = 0; ch < nchunks; ++ch)
      for (int j = 0; j < K; ++j)
        for (int kg = 0; kg < KG; ++kg)
          for (int part = 0; part < planes; ++part)
            for (int n = 0; n < N; ++n)
              for (int e = 0; e < 8; ++e) {
                const int o = nt * N + n, c = ch * KC + kg * 8 + e;
                float v = 0.f;
                if (o < Cout) v = w_ock[((size_t)o * Cin + c) * K + j] * scale;
                if (planes == 1) {
                  const __nv_bfloat16 bv = __float2bfloat16_rn(v);
                  out[idx++] = *reinterpret_cast<const uint16_t*>(&bv);
                } else {
                  const __half h = __float2half_rn(v);
                  const __half l = __float2half_rn((v - __half2float(h)) * LO_SCALE);
                  const __half pick = part ? l : h;
                  out[idx++] = *reinterpret_cast<const uint16_t*>(&pick);
                }
              }
}

// Ring depths for one layer shape: A ring of 4 (2 when the halo makes stages large), the rest of the
// 227 KB goes to the weight ring; `resident` when every (chunk, tap) stage of the layer fits at once.
size_t conv_tc_bias_bytes(int Cout, int N) { return (((size_t)((Cout + N - 1) / N) * N * 4) + 127) & ~(size_t)127; }

void conv_tc_plan(int Cin, int Cout, int K, int dil, int N, bool tma, int planes, int* na, int* nw, int* resident,
                  size_t* smem_bytes) {
  const size_t fixed = (HEADER_BYTES + conv_tc_bias_bytes(Cout, N) + 1023) & ~(size_t)1023;  // A ring starts 1024 B-aligned
  const size_t budget = 227 * 1024 - fixed;
  const size_t a_stage = (size_t)conv_tc_rows(K, dil) * 16 * KG * planes, w_stage = (size_t)N * 16 * planes * KG;
  const int per_tile = (Cin / KC) * K, ntiles_n = (Cout + N - 1) / N;
  int A = 4;
  if (A * a_stage + 2 * w_stage > budget) A = 2;
  int W = (int)((budget - A * a_stage) / w_stage);
  if (W > MAXNW) W = MAXNW;
  int res = 0;
  if (ntiles_n == 1 && per_tile <= W) res = 1, W = per_tile;
  if (!res && tma && ntiles_n == 1 && per_tile <= MAXNW && (size_t)per_tile * w_stage + 2 * a_stage <= budget) {
    // TMA-fed layers prefer resident weights even with a 2-deep A ring: the MMA lane then runs the bare issue
    // loop (no per-tap wait / release), which is what bounds narrow layers (C = 64, k = 11: 22 stages of 8 KB)
    res = 1, W = per_tile, A = 2;
  }
  if (tma) {
    // TMA-fed A ring: the ring depth is the prefetch distance that hides HBM latency (no registers
    // involved), so it takes whatever the weights leave: all stages when resident, else >= 6 weight stages
    const int minW = res ? per_tile : (W < 6 ? W : 6);
    int A2 = (int)((budget - (size_t)minW * w_stage) / a_stage);
    if (A2 > NA_MAX) A2 = NA_MAX;
    if (A2 > A) {
      A = A2;
      if (!res) {
        W = (int)((budget - A * a_stage) / w_stage);
        if (W > MAXNW) W = MAXNW;
      }
    }
  }
  *na = A, *nw = W, *resident = res;
  *smem_bytes = fixed + A * a_stage + (size_t)W * w_stage;
}

cudaError_t launch_split_image(const float* x, int B, int C, int L, float slope, uint16_t* img, int planes,
                               cudaStream_t stream) {
  if (C % 32 != 0 || (planes != 1 && planes != 2)) return cudaErrorInvalidValue;
  if (B <= 0 || C <= 0 || L <= 0) return cudaSuccess;
  split_image_kernel<<<dim3((L + 255) / 256, C / 32, B), 256, 0, stream>>>(x, B, C, L, slope, img, planes);
  return cudaGetLastError();
}

// Operand image [hi|lo][B*C/32][L][32] as a 4-D tensor; box = (32 ch, rows, 1, planes) = one A stage, 64 B swizzle.
cudaError_t tc_make_image_map(const uint16_t* img, int B, int C, int L, int rows, int planes, CUtensorMap* map) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return cudaErrorNotSupported;
  const cuuint64_t dims[4] = {32, (cuuint64_t)L, (cuuint64_t)B * (C / 32), (cuuint64_t)planes};
  const cuuint64_t strides[3] = {64, (cuuint64_t)L * 64, (cuuint64_t)B * (C / 32) * L * 64};
  const cuuint32_t box[4] = {32, (cuuint32_t)rows, 1, (cuuint32_t)planes};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(map, planes == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<uint16_t*>(img), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}


cudaError_t launch_conv_tc(const ConvTcArgs& ta_in, cudaStream_t stream) {
  ConvTcArgs ta = ta_in;
  const ConvArgs& a = ta.c;
  if (a.Cin % KC != 0 || ta.N % 16 != 0 || ta.N < 16 || ta.N > 128) return cudaErrorInvalidValue;
  if (a.split != (1 << 30) && (a.split % 16 != 0 || a.mode != MODE_STORE)) return cudaErrorInvalidValue;
  if (a.mode == MODE_GATE && (ta.N % 32 != 0 || a.Cout % 2)) return cudaErrorInvalidValue;
  if (a.B <= 0 || a.Lout <= 0 || a.Cout <= 0) return cudaSuccess;
  size_t smem = 0;
  if (ta.planes != 1) ta.planes = 2;
  // MODE_SHUFFLE without an fp32 destination: only the lean stride-2 epilogue (operand image out) supports it
  if (a.mode == MODE_SHUFFLE && !a.e[0].y && !(a.shuf_rmajor && a.e[0].split) &&
      !(ta.x_split && a.shuf_s == 2 && a.shuf_p == 1 && a.Cout % 16 == 0 && (a.y_stride & 1) == 0 && ta.planes == 2 && a.e[0].split))
    return cudaErrorInvalidValue;
  if (a.mode == MODE_SHUFFLE && a.shuf_rmajor && (a.shuf_s < 1 || a.Cout % a.shuf_s || (a.Cout / a.shuf_s) % 16)) return cudaErrorInvalidValue;
  conv_tc_plan(a.Cin, a.Cout, a.K, a.dil, ta.N, ta.x_split != nullptr, ta.planes, &ta.na, &ta.nw, &ta.resident, &smem);
  if (ta.nw < 2 && !ta.resident) return cudaErrorInvalidValue;
  // accumulator ring: as many (main + cross) stages as fit the 512 TMEM columns, at least 2, power of two
  ta.nacc = 2;
  while (ta.nacc < MAXACC && 2 * ta.nacc * ta.planes * ta.N <= 512) ta.nacc *= 2;
  // epilogue warp groups that alternate tiles: the largest divisor of the warps-per-quarter count that
  // still leaves every group two accumulator stages (wide layers: one group, columns split instead)
  {
    const int quarters = (ta.x_split ? EPI_WARPS_TMA : EPI_WARPS) / 4;
    ta.epi_groups = 1;
    for (int g = 2; g <= quarters; ++g)
      if (quarters % g == 0 && ta.nacc >= 2 * g && ta.nacc >= 4) ta.epi_groups = g;
  }
  {
    // lean STORE epilogue: one destination side with an operand image out, optional fp32 residual in / fp32 out,
    // full 16-column jobs in pairs, 32-bit element offsets
    const EpiDesc& d = a.e[0];
    const int quarters = (ta.x_split ? EPI_WARPS_TMA : EPI_WARPS) / 4, esplit = quarters / ta.epi_groups;
    const int nch = ta.N / 16, hc = (nch + esplit - 1) / esplit;
    static const bool allow = [] {
      const char* e = getenv("SVK_EPI_FAST");
      return !(e && e[0] == '0');
    }();
    // combinations instantiated in the kernel: no residual -> image only; residual (fp32 tensor or operand image) ->
    // image only, or fp32 out with optional running sum and optional image
    const bool has_res = d.res != nullptr || d.res_img != nullptr;
    const bool combo = (!has_res && !d.y && !d.acc_in && d.split) || (has_res && !d.y && d.split && (!d.acc_in || d.res_img)) || (has_res && d.y);
    ta.epi_fast = allow && a.mode == MODE_STORE && (ta.planes == 2 || !d.res_img) && a.split == (1 << 30) && combo && !(d.res && d.res_img) &&
                  (!d.res_img || d.res_slope > 0.f) && d.ch_sign == 1 && d.ch_off % 16 == 0 && d.C % 32 == 0 && !a.act_tanh &&
                  !(d.use_mask && a.out_mask) && a.Cout % ta.N == 0 && nch % esplit == 0 && hc % 2 == 0 &&
                  2ull * a.B * d.C * (unsigned long long)a.y_stride < (1ull << 32) && ta.x_split != nullptr;
  }
  {
    // second MMA issuer (spare warp of the TMA variant) for resident-weight layers whose MMAs are short (N' = 2N <= 128)
    static const int dual = [] {
      const char* e = getenv("SVK_DUAL_ISSUE");
      return e ? atoi(e) : 1;
    }();
    ta.dual_issue = dual && ta.x_split != nullptr && ta.resident && ta.nacc >= 4 && (dual >= 2 || ta.planes * ta.N <= 128);
    // An mbarrier parity wait is only meaningful one phase away from the barrier's current phase, so every A stage must
    // always be consumed by the same issuer: ring depth = a multiple of two tiles' chunks (accumulator stages: nacc is even).
    const int two_tiles = 2 * (a.Cin / KC);
    if (ta.dual_issue && ta.na >= two_tiles) ta.na = ta.na / two_tiles * two_tiles;
    else ta.dual_issue = 0;
  }
  {
    static const int lazy = [] {
      const char* e = getenv("SVK_LAZY_RES");
      return e ? atoi(e) : 1;
    }();
    ta.lazy_res = lazy;
  }
  int cols = 32;
  while (cols < ta.nacc * ta.planes * ta.N) cols <<= 1;
  ta.tmem_cols = cols;
  ta.rows = conv_tc_rows(a.K, a.dil);
  ta.bias_bytes = (int)conv_tc_bias_bytes(a.Cout, ta.N);
  ta.a_off = (HEADER_BYTES + ta.bias_bytes + 1023) & ~1023;
  ta.bias_count = ((a.Cout + ta.N - 1) / ta.N) * ta.N;
  ta.ntiles_t = (a.Lout + 127) / 128;
  const int ntiles_n = (a.Cout + ta.N - 1) / ta.N;
  const long long items = (long long)ntiles_n * a.B * ta.ntiles_t;
  if (items > 0x7FFFFFFFLL / 8) return cudaErrorInvalidValue;
  ta.items = (int)items;
  ta.div_t = make_fast_div((uint32_t)ta.ntiles_t), ta.div_b = make_fast_div((uint32_t)a.B);
  static int sm_count[64] = {0};
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  const int grid = ta.items < sm_count[dev] ? ta.items : sm_count[dev];
  for (int sd = 0; sd < 2; ++sd) {
    if (a.e[sd].res_img && (a.mode != MODE_STORE || ta.planes == 1 || a.e[sd].ch_sign != 1 || a.e[sd].ch_off % 16 || a.e[sd].C % 32 ||
                            a.Cout % 16 || a.e[sd].res || !(a.e[sd].res_slope > 0.f)))
      return cudaErrorInvalidValue;
    if (a.e[sd].split && (a.e[sd].ch_sign != 1 || a.e[sd].ch_off % 16 || a.e[sd].C % 32 ||
                          (a.mode == MODE_STORE ? a.Cout % 16 : a.Cout % 32) ||
                          (a.mode == MODE_SHUFFLE && ((a.shuf_s != 2 && !a.shuf_rmajor) || a.shuf_Lout != a.y_stride))))
      return cudaErrorInvalidValue;
  }
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (ta.x_split) {
    if (a.x_C % 32 || a.x_ch_off % 32 || a.x_stride != a.Lin) return cudaErrorInvalidValue;
    cudaError_t e = tc_make_image_map(ta.x_split, a.B, a.x_C, a.Lin, ta.rows, ta.planes, &map);
    if (e != cudaSuccess) return e;
    return launch_pdl(conv_tc_kernel<true>, grid, THREADS_TMA, smem, stream, ta, map);
  }
  return launch_pdl(conv_tc_kernel<false>, grid, THREADS, smem, stream, ta, map);
}

}  // namespace svk
