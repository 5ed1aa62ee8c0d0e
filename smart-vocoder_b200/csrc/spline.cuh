// Piecewise rational-quadratic spline with linear tails, one element (reference transforms.py:12-193): shared by the
// standalone operator (elementwise.cu, svk_rq_spline) and the ConvFlow operator (convflow.cu, svk_convflow).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace svk {

constexpr int SPLINE_MAX_BINS = 32;

// uw_at(i), uh_at(i): unnormalised widths / heights, i < nb; ud_at(i): unnormalised derivatives, i < nb - 1.
// Outside [-tail_bound, tail_bound]: identity, logabsdet 0, bin -1 (transforms.py:65-78).
template <class FW, class FH, class FD>
__device__ __forceinline__ void rq_spline_element(float xin, FW uw_at, FH uh_at, FD ud_at, int nb, int inverse, float tail_bound,
                                                  float min_bw, float min_bh, float min_d, float& y, float& lad, int& bin_out) {
  const float left = -tail_bound, right = tail_bound;
  if (!(xin >= left && xin <= right)) {
    y = xin, lad = 0.f, bin_out = -1;
    return;
  }
  // transforms.py:72-75 (computed in double by numpy, then stored into an fp32 tensor)
  const float cst = (float)log(exp(1.0 - (double)min_d) - 1.0);
  float cw[SPLINE_MAX_BINS + 1], chh[SPLINE_MAX_BINS + 1];
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    float* cum = pass == 0 ? cw : chh;
    const float minb = pass == 0 ? min_bw : min_bh;
    auto at = [&](int i) { return pass == 0 ? uw_at(i) : uh_at(i); };
    float mx = at(0);
    for (int i = 1; i < nb; ++i) mx = fmaxf(mx, at(i));
    float sum = 0.f;
    for (int i = 0; i < nb; ++i) {
      cum[i + 1] = expf(at(i) - mx);
      sum += cum[i + 1];
    }
    const float scale = (float)(1.0 - (double)minb * nb);
    float run = 0.f;
    for (int i = 0; i < nb; ++i) {
      const float sm = cum[i + 1] / sum;
      run = __fadd_rn(run, __fadd_rn(minb, __fmul_rn(scale, sm)));
      cum[i + 1] = __fadd_rn(__fmul_rn(right - left, run), left);
    }
    cum[0] = left;
    cum[nb] = right;
  }
  const float* knots = inverse ? chh : cw;
  int bin = -1;
  for (int i = 0; i <= nb; ++i) {
    float kn = knots[i];
    if (i == nb) kn = __fadd_rn(kn, 1e-6f);
    bin += (xin >= kn) ? 1 : 0;
  }
  bin_out = bin;
  const int bi = bin < 0 ? 0 : (bin > nb - 1 ? nb - 1 : bin);
  const float in_cw = cw[bi], in_w = cw[bi + 1] - cw[bi];
  const float in_ch = chh[bi], in_h = chh[bi + 1] - chh[bi];
  const float delta = in_h / in_w;
  const float u0 = bi == 0 ? cst : ud_at(bi - 1);
  const float u1 = bi + 1 == nb ? cst : ud_at(bi);
  const float d0 = min_d + (u0 > 20.f ? u0 : log1pf(expf(u0)));
  const float d1 = min_d + (u1 > 20.f ? u1 : log1pf(expf(u1)));
  const float s2 = d0 + d1 - 2.f * delta;
  if (inverse) {  // transforms.py:152-177
    const float dy = xin - in_ch;
    const float a = dy * s2 + in_h * (delta - d0);
    const float b = in_h * d0 - dy * s2;
    const float c = -delta * dy;
    const float disc = b * b - 4.f * a * c;
    const float root = (2.f * c) / (-b - sqrtf(disc));
    y = root * in_w + in_cw;
    const float tomt = root * (1.f - root);
    const float den = delta + s2 * tomt;
    const float num = delta * delta * (d1 * root * root + 2.f * delta * tomt + d0 * (1.f - root) * (1.f - root));
    lad = -(logf(num) - 2.f * logf(den));
  } else {  // transforms.py:178-193
    const float theta = (xin - in_cw) / in_w;
    const float tomt = theta * (1.f - theta);
    const float numr = in_h * (delta * theta * theta + d0 * tomt);
    const float den = delta + s2 * tomt;
    y = in_ch + numr / den;
    const float num = delta * delta * (d1 * theta * theta + 2.f * delta * tomt + d0 * (1.f - theta) * (1.f - theta));
    lad = logf(num) - 2.f * logf(den);
  }
}

}  // namespace svk
