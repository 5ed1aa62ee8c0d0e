// One launch = one layer of modules.WN.forward (reference modules.py:156-175, g = None):
//
//     x_in = in_layer(x)                      Conv1d(H -> 2H, k, padding (k-1)/2)
//     acts = tanh(x_in[:H]) * sigmoid(x_in[H:])        commons.fused_add_tanh_sigmoid_multiply, g_l = 0
//     rs   = res_skip(acts)                   Conv1d(H -> 2H, 1)   (last layer: H -> H)
//     x    = (x + rs[:H]) * mask ; out += rs[H:]        (last layer: out = (out + rs) * mask)
//
// `acts` never leaves the SM: the gate epilogue writes it as the A operand of the 1x1 conv straight into shared
// memory (fp16 hi/lo, 64 B-swizzled rows, fence.proxy.async), and the second GEMM runs on that tile.  Unfused this was
// two launches per layer (wn_in + wn_res_skip, 49 + 35 us at 16 x 1024 frames) with `acts` written to and read back
// from HBM as an operand image; fused, HBM sees x (image in, fp32 + image out), out and the weights.
//
// Same machinery as conv_tc.cu (three-product fp16 split, taps as row shifts of one 64 B-swizzled tile, TMA-fed
// operand images, weight stages by cp.async.bulk, accumulators in TMEM, warp roles, PDL).  One CTA owns one tile of
// 128 frames through BOTH GEMMs:
//     in_layer   : N-tiles nt = 0..2 of 128 virtual channels (64 tanh + the matching 64 sigmoid channels, packed that
//                  way by pack_tc(gate_half)), K = H x taps; two accumulator stages (main + cross, 2 x 256 columns =
//                  all of TMEM), so the gate epilogue of N-tile nt runs under the MMAs of nt + 1
//     res_skip   : N-tiles mt of 128 (96 on the last layer), K = H, A = the acts tile; its first chunks are issued as
//                  soon as the gate epilogues that produce them have finished (one mbarrier per N-tile of in_layer)
// The x image is ping-ponged between two buffers by the caller: neighbouring CTAs read halo rows of the OLD x while
// this CTA writes the new one.
#include <string.h>

#include "svk_kernels.cuh"
#include "tc_common.cuh"

namespace svk {

namespace {

constexpr int WN_NA_MAX = 4, WN_NW_MAX = 8, WN_MAX_NT = 8;
constexpr int WN_EPI_WARPS = 8;
constexpr int WN_THREADS = 128 + 32 * WN_EPI_WARPS;
constexpr int WN_EPI_THREADS = 32 * WN_EPI_WARPS;

struct __align__(8) WnHeader {
  uint64_t a_full[WN_NA_MAX], a_empty[WN_NA_MAX];
  uint64_t w_full[WN_NW_MAX], w_empty[WN_NW_MAX];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t acts_full[WN_MAX_NT];
  uint32_t tmem_base;
  uint32_t pad;
};
constexpr int WN_HEADER_BYTES = 512;
static_assert(sizeof(WnHeader) <= WN_HEADER_BYTES, "header");

__global__ void __launch_bounds__(WN_THREADS, 1) wn_layer_kernel(const WnLayerArgs wa, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) uint8_t smem[];
  WnHeader* hdr = reinterpret_cast<WnHeader*>(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = wa.H, K = wa.K, planes = wa.planes;
  const int nchunks = H / KC;
  const int N1 = wa.N_in, N2 = wa.N_rs, nt_in = wa.nt_in, nt_rs = wa.nt_rs;
  const int na = wa.na, nw = wa.nw;
  const uint32_t a_plane = (uint32_t)wa.rows * 64u, a_stage = a_plane * planes;
  const uint32_t acts_plane = 128u * 64u, acts_chunk = acts_plane * planes;
  const uint32_t w1_plane2 = (uint32_t)N1 * 16u * planes, w1_stage = w1_plane2 * KG;
  const uint32_t w2_plane2 = (uint32_t)N2 * 16u * planes, w2_stage = w2_plane2 * KG;
  const uint32_t w_slot = wa.w_slot;
  const int acc_stride = wa.acc_stride;  // TMEM columns between the two accumulator stages
  float* bias_s = reinterpret_cast<float*>(smem + WN_HEADER_BYTES);  // in_layer bias (virtual order), then res_skip bias
  float* bias2_s = bias_s + wa.bias_count_in;
  uint8_t* a_smem = smem + wa.a_off;
  uint8_t* acts_smem = smem + wa.acts_off;
  uint8_t* w_smem = smem + wa.w_off;
  const int items = wa.items;
  const int n_my = (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < na; ++i) mbar_init(&hdr->a_full[i], 1), mbar_init(&hdr->a_empty[i], 1);
    for (int i = 0; i < WN_NW_MAX; ++i) mbar_init(&hdr->w_full[i], 1), mbar_init(&hdr->w_empty[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&hdr->acc_full[i], 1), mbar_init(&hdr->acc_empty[i], WN_EPI_THREADS);
    for (int i = 0; i < WN_MAX_NT; ++i) mbar_init(&hdr->acts_full[i], WN_EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&hdr->tmem_base, (uint32_t)wa.tmem_cols);
  for (int i = tid; i < wa.bias_count_in; i += WN_THREADS) bias_s[i] = __ldg(wa.bias_in + i);
  for (int i = tid; i < wa.bias_count_rs; i += WN_THREADS) bias2_s[i] = __ldg(wa.bias_rs + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = hdr->tmem_base;
  griddep_launch_dependents();  // PDL: only weights / biases are touched before griddep_wait()

  if (warp == 0) {
    // ------------------------------------------------ weight producer: one bulk copy per (N-tile, chunk, tap)
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      const uint8_t* s1 = reinterpret_cast<const uint8_t*>(wa.w_in);
      const uint8_t* s2 = reinterpret_cast<const uint8_t*>(wa.w_rs);
      const int n1 = nt_in * nchunks * K, n2 = nt_rs * nchunks;
      for (int i = 0; i < n_my; ++i) {
        for (int it = 0; it < n1 + n2; ++it) {
          const bool first = it < n1;
          const uint32_t bytes = first ? w1_stage : w2_stage;
          const uint8_t* src = first ? s1 + (size_t)it * w1_stage : s2 + (size_t)(it - n1) * w2_stage;
          mbar_wait(&hdr->w_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&hdr->w_full[st], bytes);
          bulk_g2s(w_smem + (size_t)st * w_slot, src, bytes, &hdr->w_full[st]);
          if (++st == nw) st = 0, ph ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (whole warp walks the loop, one elected lane issues)
    const uint32_t fmt = planes == 1 ? ((1u << 7) | (1u << 10)) : 0u;  // bf16 x bf16 for the single-plane engine
    auto idesc = [&](int n) { return (1u << 4) | fmt | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); };
    const uint32_t id1_main = idesc(N1), id1_wide = idesc(2 * N1), id2_main = idesc(N2), id2_wide = idesc(2 * N2);
    const uint32_t b_hi = (128u >> 4) | (1u << 14);               // B: no swizzle, SBO = 128 B
    const uint32_t a_hi = (512u >> 4) | (1u << 14) | (4u << 29);  // A: SWIZZLE_64B, SBO = 8 rows x 64 B
    const uint32_t a_lo0 = ((smem_u32(a_smem) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t acts_lo0 = ((smem_u32(acts_smem) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t w_addr16 = (smem_u32(w_smem) & 0x3FFFFu) >> 4;
    const uint32_t a_stage16 = a_stage >> 4, acts_chunk16 = acts_chunk >> 4, w_slot16 = w_slot >> 4;
    const uint32_t lo1_16 = a_plane >> 4, lo2_16 = acts_plane >> 4;
    const uint32_t bar_a_full = smem_u32(&hdr->a_full[0]), bar_a_empty = smem_u32(&hdr->a_empty[0]);
    const uint32_t bar_w_full = smem_u32(&hdr->w_full[0]), bar_w_empty = smem_u32(&hdr->w_empty[0]);
    const uint32_t bar_acc_full = smem_u32(&hdr->acc_full[0]), bar_acc_empty = smem_u32(&hdr->acc_empty[0]);
    const uint32_t bar_acts_full = smem_u32(&hdr->acts_full[0]);
    const bool leader = elect_one();
    int ast = 0, wst = 0;
    uint32_t aph = 0, wph = 0;
    uint32_t q = 0;  // accumulation counter: stage q & 1, phase (q >> 1) & 1

    // one (chunk, tap): two K = 16 steps of [main | cross] (+)= xh . [wh | wl] ; cross += xl . wh
    auto tap = [&](uint32_t dmain, int N, uint32_t ah, uint32_t lo16, uint32_t bw, uint32_t ks_b16, uint32_t id_main,
                   uint32_t id_wide, uint32_t acc) {
      if (planes == 2) {
        umma_f16_lo(dmain, ah, bw, a_hi, b_hi, id_wide, acc);
        umma_f16_lo(dmain + (uint32_t)N, ah + lo16, bw, a_hi, b_hi, id_main, 1u);
        umma_f16_lo(dmain, ah + 2u, bw + ks_b16, a_hi, b_hi, id_wide, 1u);
        umma_f16_lo(dmain + (uint32_t)N, ah + 2u + lo16, bw + ks_b16, a_hi, b_hi, id_main, 1u);
      } else {
        umma_f16_lo(dmain, ah, bw, a_hi, b_hi, id_main, acc);
        umma_f16_lo(dmain, ah + 2u, bw + ks_b16, a_hi, b_hi, id_main, 1u);
      }
    };
    for (int i = 0; i < n_my; ++i) {
      // ---- in_layer: N-tiles of N1 virtual channels, K = H x taps
      for (int nt = 0; nt < nt_in; ++nt, ++q) {
        const uint32_t s = q & 1u;
        mbar_wait_u32(bar_acc_empty + 8u * s, ((q >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t dmain = tmem + s * (uint32_t)acc_stride;
        uint32_t acc = 0;
        for (int ch = 0; ch < nchunks; ++ch) {
          mbar_wait_u32(bar_a_full + 8u * ast, aph);
          tc_fence_after();
          uint32_t ah = a_lo0 + (uint32_t)ast * a_stage16;
          for (int j = 0; j < K; ++j) {
            mbar_wait_u32(bar_w_full + 8u * wst, wph);
            tc_fence_after();
            if (leader) {
              const uint32_t bw = (w_addr16 + (uint32_t)wst * w_slot16) | ((w1_plane2 >> 4) << 16);
              tap(dmain, N1, ah, lo1_16, bw, (2 * w1_plane2) >> 4, id1_main, id1_wide, acc);
              umma_commit_u32(bar_w_empty + 8u * wst);
            }
            acc = 1u;
            ah += 4u;  // dilation 1: next tap = next 64 B row
            if (++wst == nw) wst = 0, wph ^= 1;
          }
          if (leader) umma_commit_u32(bar_a_empty + 8u * ast);
          if (++ast == na) ast = 0, aph ^= 1;
        }
        if (leader) umma_commit_u32(bar_acc_full + 8u * s);
      }
      // ---- res_skip: N-tiles of N2 channels, K = H, A = the acts tile the gate epilogues wrote
      for (int mt = 0; mt < nt_rs; ++mt, ++q) {
        const uint32_t s = q & 1u;
        mbar_wait_u32(bar_acc_empty + 8u * s, ((q >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t dmain = tmem + s * (uint32_t)acc_stride;
        uint32_t acc = 0;
        for (int ch = 0; ch < nchunks; ++ch) {
          // acts chunk ch comes from in_layer N-tile (ch * 32) / (N1 / 2); completes once per item
          mbar_wait_u32(bar_acts_full + 8u * (uint32_t)((ch * KC) / (N1 >> 1)), (uint32_t)i & 1u);
          mbar_wait_u32(bar_w_full + 8u * wst, wph);
          tc_fence_after();
          if (leader) {
            const uint32_t ah = acts_lo0 + (uint32_t)ch * acts_chunk16;
            const uint32_t bw = (w_addr16 + (uint32_t)wst * w_slot16) | ((w2_plane2 >> 4) << 16);
            tap(dmain, N2, ah, lo2_16, bw, (2 * w2_plane2) >> 4, id2_main, id2_wide, acc);
            umma_commit_u32(bar_w_empty + 8u * wst);
          }
          acc = 1u;
          if (++wst == nw) wst = 0, wph ^= 1;
        }
        if (leader) umma_commit_u32(bar_acc_full + 8u * s);
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------ x-image loader: one TMA box per 32-channel chunk and N-tile
    if (lane == 0) {
      griddep_wait();
      int as = 0;
      uint32_t ph = 0;
      const int cgs = H >> 5, pad = (K - 1) / 2;
      for (int i = 0; i < n_my; ++i) {
        const int item = (int)blockIdx.x + i * (int)gridDim.x;
        const int b = (int)fast_div((uint32_t)item, wa.div_t), tt = item - b * wa.ntiles_t;
        for (int nt = 0; nt < nt_in; ++nt)
          for (int ch = 0; ch < nchunks; ++ch) {
            mbar_wait(&hdr->a_empty[as], ph ^ 1);
            mbar_arrive_expect_tx(&hdr->a_full[as], a_stage);
            tma_load_4d(a_smem + (size_t)as * a_stage, &tmap, 0, tt * 128 - pad, b * cgs + ch, 0, &hdr->a_full[as]);
            if (++as == na) as = 0, ph ^= 1;
          }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue warps: gate -> acts tile ; res_skip -> x, out
    griddep_wait();
    const int q4 = warp & 3, part = (warp - 4) >> 2;  // TMEM lane quarter; which half of a tile's columns
    const int row = q4 * 32 + lane;
    const int T = wa.T;
    const size_t img_plane = (size_t)wa.B * H * T;  // halves between the hi and lo planes of an x image
    const int hN = N1 >> 1;                          // gate channels per in_layer N-tile
    uint32_t q = 0;
    for (int i = 0; i < n_my; ++i) {
      const int item = (int)blockIdx.x + i * (int)gridDim.x;
      const int b = (int)fast_div((uint32_t)item, wa.div_t), tt = item - b * wa.ntiles_t;
      const int t = tt * 128 + row;
      const bool tin = t < T;
      const int tl = tin ? t : T - 1;
      // ---- gate epilogues: acts[c] = tanh(a[c]) * sigmoid(a[c + H]) -> fp16 hi/lo rows of the acts tile
      for (int nt = 0; nt < nt_in; ++nt, ++q) {
        const uint32_t s = q & 1u;
        mbar_wait(&hdr->acc_full[s], (q >> 1) & 1u);
        tc_fence_after();
        const uint32_t tsub = tmem + ((uint32_t)(q4 * 32) << 16) + s * (uint32_t)acc_stride;
        const float* bptr = bias_s + nt * N1;
        const int npair = hN >> 4, hp = (npair + 1) >> 1;  // 16-channel jobs of this tile, split between the two parts
        for (int jb = part * hp; jb < min(npair, (part + 1) * hp); ++jb) {
          const int n0 = jb * 16;
          uint32_t m[16], c[16];
          float g[16];
          tmem_ld16(tsub + (uint32_t)n0, m);
          if (planes == 2) {
            tmem_ld16(tsub + (uint32_t)(N1 + n0), c);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) c[e] = 0u;
          }
          tmem_wait_ld();
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 bq = reinterpret_cast<const float4*>(bptr + n0)[e4];
            const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              g[4 * e4 + e] = fmaf(fmaf(__uint_as_float(c[4 * e4 + e]), LO_INV, __uint_as_float(m[4 * e4 + e])), wa.unscale_in, bb[e]);  // tanh side, pre-activation
          }
          tmem_ld16(tsub + (uint32_t)(hN + n0), m);
          if (planes == 2) tmem_ld16(tsub + (uint32_t)(N1 + hN + n0), c);
          tmem_wait_ld();
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 bq = reinterpret_cast<const float4*>(bptr + hN + n0)[e4];
            const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              g[4 * e4 + e] = gate_tanh_sigmoid(g[4 * e4 + e], fmaf(fmaf(__uint_as_float(c[4 * e4 + e]), LO_INV, __uint_as_float(m[4 * e4 + e])), wa.unscale_in, bb[e]));
          }
          // 16 channels = two 16 B pieces of this row of chunk (channel / 32), per plane
          const int ch0 = nt * hN + n0;
          uint8_t* tile = acts_smem + (size_t)(ch0 >> 5) * acts_chunk;
          uint4* hi = reinterpret_cast<uint4*>(tile) + row * KG;
          uint4* lo = reinterpret_cast<uint4*>(tile + acts_plane) + row * KG;
          const int swz = (int)((smem_u32(hi) >> 7) & 3u);  // 64 B swizzle on absolute address bits 7-8
          const int kg0 = (ch0 & 31) >> 3;
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            if (planes == 2) {
              uint4 hq, lq;
              split2(g[8 * g8 + 0], g[8 * g8 + 1], hq.x, lq.x);
              split2(g[8 * g8 + 2], g[8 * g8 + 3], hq.y, lq.y);
              split2(g[8 * g8 + 4], g[8 * g8 + 5], hq.z, lq.z);
              split2(g[8 * g8 + 6], g[8 * g8 + 7], hq.w, lq.w);
              hi[(kg0 + g8) ^ swz] = hq;
              lo[(kg0 + g8) ^ swz] = lq;
            } else {
              hi[(kg0 + g8) ^ swz] = pack_bf16x8(&g[8 * g8]);
            }
          }
        }
        fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        tc_fence_before();
        mbar_arrive(&hdr->acts_full[nt]);
        mbar_arrive(&hdr->acc_empty[s]);
      }
      // ---- res_skip epilogues: x = (x + rs[:H]) * mask (+ its operand image), out += rs[H:]
      const float mv = __ldg(wa.mask + (size_t)b * T + tl);
      for (int mt = 0; mt < nt_rs; ++mt, ++q) {
        const uint32_t s = q & 1u;
        const int nch = N2 >> 4, hc = (nch + 1) >> 1;
        const int j_lo = part * hc, j_hi = min(nch, (part + 1) * hc);
        // the residual / running-sum operands of every job of this tile are requested BEFORE the accumulator is waited
        // for (they do not depend on it): their latency hides under the res_skip MMAs instead of stalling each job
        constexpr int MAXJ = 4;  // N2 <= 128: at most 4 jobs of 16 columns per part
        float r[MAXJ][16];
        auto load_ops = [&](int jb, float (&o)[16]) {
          const int o0 = mt * N2 + jb * 16;
          const bool res_side = !wa.last && o0 < H;
          const float* src = nullptr;
          if (o0 < wa.Cout_rs) {
            if (res_side) src = wa.x + ((size_t)b * H + o0) * T + tl;
            else if (!wa.first) src = wa.out + ((size_t)b * H + (wa.last ? o0 : o0 - H)) * T + tl;
          }
          if (src) {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = src[(size_t)e * T];
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = 0.f;
          }
        };
#pragma unroll
        for (int k = 0; k < MAXJ; ++k)
          if (j_lo + k < j_hi) load_ops(j_lo + k, r[k]);
        mbar_wait(&hdr->acc_full[s], (q >> 1) & 1u);
        tc_fence_after();
        const uint32_t tsub = tmem + ((uint32_t)(q4 * 32) << 16) + s * (uint32_t)acc_stride;
#pragma unroll
        for (int k = 0; k < MAXJ; ++k) {
          const int jb = j_lo + k;
          if (jb >= j_hi) break;
          const int n0 = jb * 16, o0 = mt * N2 + n0;
          uint32_t m[16], c[16];
          tmem_ld16(tsub + (uint32_t)n0, m);
          if (planes == 2) {
            tmem_ld16(tsub + (uint32_t)(N2 + n0), c);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) c[e] = 0u;
          }
          tmem_wait_ld();
          if (o0 < wa.Cout_rs) {
            float v[16];
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 bq = reinterpret_cast<const float4*>(bias2_s + o0)[e4];
              const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
              for (int e = 0; e < 4; ++e)
                v[4 * e4 + e] = fmaf(fmaf(__uint_as_float(c[4 * e4 + e]), LO_INV, __uint_as_float(m[4 * e4 + e])), wa.unscale_rs, bb[e]) + r[k][4 * e4 + e];
            }
            const bool res_side = !wa.last && o0 < H;
            if (res_side || wa.last) {
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] *= mv;
            }
            if (tin) {
              float* dst = res_side ? wa.x + ((size_t)b * H + o0) * T + t : wa.out + ((size_t)b * H + (wa.last ? o0 : o0 - H)) * T + t;
#pragma unroll
              for (int e = 0; e < 16; ++e) dst[(size_t)e * T] = v[e];
              if (res_side && wa.x_img_out) {
                uint16_t* sp = wa.x_img_out + (((size_t)b * (H >> 5) + (o0 >> 5)) * T + t) * 32 + (o0 & 31);
                uint4 h2[2], l2[2];
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) {
                  if (planes == 2) {
                    split2(v[8 * g8 + 0], v[8 * g8 + 1], h2[g8].x, l2[g8].x);
                    split2(v[8 * g8 + 2], v[8 * g8 + 3], h2[g8].y, l2[g8].y);
                    split2(v[8 * g8 + 4], v[8 * g8 + 5], h2[g8].z, l2[g8].z);
                    split2(v[8 * g8 + 6], v[8 * g8 + 7], h2[g8].w, l2[g8].w);
                  } else {
                    h2[g8] = pack_bf16x8(&v[8 * g8]);
                  }
                }
                st_global_v8(sp, h2[0], h2[1]);  // 16 channels = one 32 B sector per plane
                if (planes == 2) st_global_v8(sp + img_plane, l2[0], l2[1]);
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&hdr->acc_empty[s]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)wa.tmem_cols);
}

}  // namespace

// Shared-memory plan: header | biases | A ring (x tile chunks) | acts tile | weight ring.  false when the layer does
// not fit (the caller then runs the two-launch form).
static bool wn_plan(WnLayerArgs& wa, size_t* smem_bytes) {
  const int H = wa.H, planes = wa.planes, nchunks = H / KC;
  if (H % KC || wa.K < 1 || !(wa.K & 1) || wa.N_in % 32 || wa.N_in > 128 || wa.N_rs % 16 || wa.N_rs > 128 || wa.N_in < 32) return false;
  wa.nt_in = (2 * H + wa.N_in - 1) / wa.N_in;
  wa.nt_rs = (wa.Cout_rs + wa.N_rs - 1) / wa.N_rs;
  // every 32-channel chunk of the acts tile must come from ONE in_layer N-tile (one acts_full barrier per N-tile)
  if (wa.nt_in > WN_MAX_NT || (H % (wa.N_in / 2)) != 0 || (wa.N_in / 2) % KC != 0) return false;
  wa.rows = conv_tc_rows(wa.K, 1);
  wa.bias_count_in = wa.nt_in * wa.N_in;
  wa.bias_count_rs = wa.nt_rs * wa.N_rs;
  const int nmax = wa.N_in > wa.N_rs ? wa.N_in : wa.N_rs;
  wa.acc_stride = planes * nmax;
  int cols = 32;
  while (cols < 2 * wa.acc_stride) cols <<= 1;
  if (cols > 512) return false;
  wa.tmem_cols = cols;
  const size_t fixed = (WN_HEADER_BYTES + (size_t)(wa.bias_count_in + wa.bias_count_rs) * 4 + 1023) & ~(size_t)1023;
  const size_t a_stage = (size_t)wa.rows * 64 * planes, acts = (size_t)nchunks * 128 * 64 * planes;
  const size_t w1 = (size_t)wa.N_in * 16 * planes * KG, w2 = (size_t)wa.N_rs * 16 * planes * KG;
  wa.w_slot = (int)(w1 > w2 ? w1 : w2);
  const size_t budget = 227 * 1024;
  wa.na = 2;
  if (fixed + wa.na * a_stage + acts + 3 * (size_t)wa.w_slot > budget) return false;
  // (a third x-tile stage at the price of a weight stage measured slower: 0.764 vs 0.733 ms for the 16-layer encoder)
  wa.nw = (int)((budget - fixed - wa.na * a_stage - acts) / wa.w_slot);
  if (wa.nw > WN_NW_MAX) wa.nw = WN_NW_MAX;
  // left-over room goes back to the A ring
  while (wa.na < WN_NA_MAX && fixed + (wa.na + 1) * a_stage + acts + (size_t)wa.nw * wa.w_slot <= budget) wa.na++;
  wa.a_off = (int)fixed;
  wa.acts_off = (int)(fixed + wa.na * a_stage);
  wa.w_off = (int)(wa.acts_off + acts);
  *smem_bytes = wa.w_off + (size_t)wa.nw * wa.w_slot;
  return *smem_bytes <= budget;
}

bool wn_layer_supported(int H, int K, int N_in, int N_rs, int Cout_rs, int planes) {
  WnLayerArgs wa;
  memset(&wa, 0, sizeof(wa));
  wa.H = H, wa.K = K, wa.N_in = N_in, wa.N_rs = N_rs, wa.Cout_rs = Cout_rs, wa.planes = planes;
  size_t smem = 0;
  return (planes == 1 || planes == 2) && wn_plan(wa, &smem);
}

cudaError_t launch_wn_layer(const WnLayerArgs& in, cudaStream_t stream) {
  WnLayerArgs wa = in;
  if (wa.planes != 1) wa.planes = 2;
  size_t smem = 0;
  if (!wn_plan(wa, &smem) || !wa.x_img_in || !wa.out || !wa.mask || !wa.w_in || !wa.w_rs) return cudaErrorInvalidValue;
  if (!wa.last && (!wa.x || wa.Cout_rs != 2 * wa.H)) return cudaErrorInvalidValue;
  if (wa.last && wa.Cout_rs != wa.H) return cudaErrorInvalidValue;
  if (wa.B <= 0 || wa.T <= 0) return cudaSuccess;
  wa.ntiles_t = (wa.T + 127) / 128;
  const long long items = (long long)wa.B * wa.ntiles_t;
  if (items > 0x7FFFFFFFLL / 8) return cudaErrorInvalidValue;
  wa.items = (int)items;
  wa.div_t = make_fast_div((uint32_t)wa.ntiles_t);
  static int sm_count[64] = {0};
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(wn_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  cudaError_t e = tc_make_image_map(wa.x_img_in, wa.B, wa.H, wa.T, wa.rows, wa.planes, &map);
  if (e != cudaSuccess) return e;
  const int grid = wa.items < sm_count[dev] ? wa.items : sm_count[dev];
  return launch_pdl(wn_layer_kernel, grid, WN_THREADS, smem, stream, wa, map);
}

}  // namespace svk
