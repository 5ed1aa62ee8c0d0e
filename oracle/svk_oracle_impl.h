/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the arithmetic on the reference's
 * SynthesizerTrn.infer path.  Included twice by svk_oracle.c with REAL=float / REAL=double.
 * Nothing under smart-vocoder_b200/ may link or call this; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg do.
 *
 * The reference performs these operations through PyTorch (third-party; reference pins
 * torch==1.6.0 in requirements.txt:6, this image carries 2.11.0).  Each function cites the
 * reference call site whose semantics it restates; operator math is SURVEY.md App. A.
 *
 * Layout everywhere: contiguous [batch, channels, time] ("NCT"), like the reference.
 */

#define FN2(p, n) svko_##p##_##n
#define FN1(p, n) FN2(p, n)
#define FN(n) FN1(SUFFIX, n)

/* torch.nn.utils.weight_norm, legacy hook (reference modules.py:128,135,145,191-206; models.py:125):
 * w[i,...] = g[i] * v[i,...] / ||v[i,...]||_2, norm over every dim except 0.  For ConvTranspose1d
 * dim 0 is the INPUT channel. */
void FN(weight_norm)(const REAL *v, const REAL *g, int64_t dim0, int64_t inner, REAL *w) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < dim0; ++i) {
    ACC s = 0;
    for (int64_t j = 0; j < inner; ++j) s += (ACC)v[i * inner + j] * (ACC)v[i * inner + j];
    REAL nrm = (REAL)SQRT(s);
    REAL scale = g[i] / nrm;
    for (int64_t j = 0; j < inner; ++j) w[i * inner + j] = v[i * inner + j] * scale;
  }
}

/* nn.Conv1d forward, stride 1 (reference models.py:32-33,120,135; modules.py:133,144,191-206,318,320).
 * y[b,o,t] = bias[o] + sum_c sum_j w[o,c,j] * x[b,c,t - pad + j*dil], zero outside [0,L).
 * L_out = L + 2*pad - dil*(k-1). */
void FN(conv1d)(const REAL *x, int64_t B, int64_t Cin, int64_t L, const REAL *w, const REAL *bias,
                int64_t Cout, int64_t k, int64_t dil, int64_t pad, REAL *y) {
  int64_t Lout = L + 2 * pad - dil * (k - 1);
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t b = 0; b < B; ++b)
    for (int64_t o = 0; o < Cout; ++o) {
      ACC *acc = (ACC *)malloc(sizeof(ACC) * (size_t)Lout);
      ACC b0 = bias ? (ACC)bias[o] : (ACC)0;
      for (int64_t t = 0; t < Lout; ++t) acc[t] = b0;
      for (int64_t c = 0; c < Cin; ++c) {
        const REAL *xr = x + (b * Cin + c) * L;
        for (int64_t j = 0; j < k; ++j) {
          ACC wv = (ACC)w[(o * Cin + c) * k + j];
          int64_t off = j * dil - pad; /* input index = t + off */
          int64_t t0 = off < 0 ? -off : 0;
          int64_t t1 = Lout;
          if (t1 + off > L) t1 = L - off;
          for (int64_t t = t0; t < t1; ++t) acc[t] += wv * (ACC)xr[t + off];
        }
      }
      REAL *yr = y + (b * Cout + o) * Lout;
      for (int64_t t = 0; t < Lout; ++t) yr[t] = (REAL)acc[t];
      free(acc);
    }
}

/* nn.ConvTranspose1d forward (reference models.py:123-127,149), weight [Cin, Cout, k]:
 * y[b,o,i*s - p + j] += x[b,c,i] * w[c,o,j];  L_out = (L-1)*s - 2p + k  (SURVEY App. A.5). */
void FN(conv_transpose1d)(const REAL *x, int64_t B, int64_t Cin, int64_t L, const REAL *w,
                          const REAL *bias, int64_t Cout, int64_t k, int64_t s, int64_t p, REAL *y) {
  int64_t Lout = (L - 1) * s - 2 * p + k;
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t b = 0; b < B; ++b)
    for (int64_t o = 0; o < Cout; ++o) {
      ACC *acc = (ACC *)malloc(sizeof(ACC) * (size_t)Lout);
      ACC b0 = bias ? (ACC)bias[o] : (ACC)0;
      for (int64_t t = 0; t < Lout; ++t) acc[t] = b0;
      for (int64_t c = 0; c < Cin; ++c) {
        const REAL *xr = x + (b * Cin + c) * L;
        const REAL *wr = w + (c * Cout + o) * k;
        for (int64_t i = 0; i < L; ++i) {
          ACC xv = (ACC)xr[i];
          for (int64_t j = 0; j < k; ++j) {
            int64_t t = i * s - p + j;
            if (t >= 0 && t < Lout) acc[t] += xv * (ACC)wr[j];
          }
        }
      }
      REAL *yr = y + (b * Cout + o) * Lout;
      for (int64_t t = 0; t < Lout; ++t) yr[t] = (REAL)acc[t];
      free(acc);
    }
}

/* commons.fused_add_tanh_sigmoid_multiply with input_b == 0 (reference commons.py:100-107,
 * called modules.py:163 with g_l = zeros): out[b,c,t] = tanh(a[b,c,t]) * sigmoid(a[b,c+H,t]). */
void FN(gate)(const REAL *a, int64_t B, int64_t H, int64_t L, REAL *out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t b = 0; b < B; ++b)
    for (int64_t c = 0; c < H; ++c) {
      const REAL *ta = a + (b * 2 * H + c) * L;
      const REAL *sa = a + (b * 2 * H + c + H) * L;
      REAL *o = out + (b * H + c) * L;
      for (int64_t t = 0; t < L; ++t) {
        REAL th = (REAL)TANH(ta[t]);
        REAL sg = (REAL)1 / ((REAL)1 + (REAL)EXP(-sa[t]));
        o[t] = th * sg;
      }
    }
}

/* F.leaky_relu (reference models.py:147 slope 0.1, models.py:156 default slope 0.01 -- SURVEY F9;
 * modules.py:212,216 slope 0.1). */
void FN(leaky_relu)(const REAL *x, int64_t n, REAL slope, REAL *y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) y[i] = x[i] > 0 ? x[i] : x[i] * slope;
}

/* transforms.piecewise_rational_quadratic_transform(..., tails='linear') -- reference
 * transforms.py:12-44 (dispatch), :47-52 (searchsorted), :55-94 (linear tails), :96-193 (spline).
 * One element = one (b,c,t) position with nb widths, nb heights, nb-1 interior derivatives
 * (layout [..., nb] / [..., nb-1], i.e. parameters contiguous per element, as ConvFlow builds
 * them at modules.py:378-383).  bin_out receives the searchsorted bin (-1 outside the tails):
 * integer work, must match bit for bit. */
void FN(rq_spline)(const REAL *inputs, const REAL *uw, const REAL *uh, const REAL *ud, int64_t n,
                   int nb, int inverse, REAL tail_bound, REAL min_bw, REAL min_bh, REAL min_d,
                   REAL *outputs, REAL *logabsdet, int32_t *bin_out) {
  const REAL left = -tail_bound, right = tail_bound, bottom = -tail_bound, top = tail_bound;
  /* transforms.py:72-75: boundary derivatives forced so that min_d + softplus(const) == 1 */
  const REAL cst = (REAL)log(exp(1.0 - (double)min_d) - 1.0);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n; ++e) {
    REAL x = inputs[e];
    if (!(x >= left && x <= right)) { /* transforms.py:65-78 */
      outputs[e] = x;
      logabsdet[e] = 0;
      if (bin_out) bin_out[e] = -1;
      continue;
    }
    REAL cw[SVKO_MAX_BINS + 1], chh[SVKO_MAX_BINS + 1], w[SVKO_MAX_BINS], h[SVKO_MAX_BINS],
        d[SVKO_MAX_BINS + 1];
    const REAL *pw = uw + e * nb, *ph = uh + e * nb, *pd = ud + e * (nb - 1);
    /* softmax -> min width -> cumsum -> affine to [left,right], ends forced (transforms.py:115-122) */
    for (int pass = 0; pass < 2; ++pass) {
      const REAL *pu = pass == 0 ? pw : ph;
      REAL *cum = pass == 0 ? cw : chh;
      REAL *seg = pass == 0 ? w : h;
      REAL lo = pass == 0 ? left : bottom, hi = pass == 0 ? right : top;
      REAL minb = pass == 0 ? min_bw : min_bh;
      REAL mx = pu[0];
      for (int i = 1; i < nb; ++i) mx = pu[i] > mx ? pu[i] : mx;
      REAL ex[SVKO_MAX_BINS];
      REAL sum = 0;
      for (int i = 0; i < nb; ++i) { ex[i] = (REAL)EXP(pu[i] - mx); sum += ex[i]; }
      REAL scale = (REAL)(1.0 - (double)minb * nb);
      REAL run = 0;
      cum[0] = lo;
      for (int i = 0; i < nb; ++i) {
        REAL sm = ex[i] / sum;
        run += minb + scale * sm;
        cum[i + 1] = (hi - lo) * run + lo;
      }
      cum[0] = lo;
      cum[nb] = hi;
      for (int i = 0; i < nb; ++i) seg[i] = cum[i + 1] - cum[i];
    }
    /* derivatives = min_d + softplus(padded ud) (transforms.py:124; F.softplus threshold 20) */
    for (int i = 0; i <= nb; ++i) {
      REAL u = (i == 0 || i == nb) ? cst : pd[i - 1];
      REAL sp = u > (REAL)20 ? u : (REAL)LOG1P(EXP(u));
      d[i] = min_d + sp;
    }
    /* searchsorted (transforms.py:47-52): last knot gets +1e-6, bin = #(x >= knot) - 1 */
    const REAL *knots = inverse ? chh : cw;
    int bin = -1;
    for (int i = 0; i <= nb; ++i) {
      REAL kn = knots[i];
      if (i == nb) kn = kn + (REAL)1e-6;
      bin += (x >= kn) ? 1 : 0;
    }
    if (bin_out) bin_out[e] = bin;
    int bi = bin < 0 ? 0 : (bin > nb - 1 ? nb - 1 : bin);
    REAL in_cw = cw[bi], in_w = w[bi], in_ch = chh[bi], in_h = h[bi];
    REAL delta = h[bi] / w[bi];
    REAL d0 = d[bi], d1 = d[bi + 1];
    if (inverse) { /* transforms.py:152-177 */
      REAL dy = x - in_ch;
      REAL s2 = d0 + d1 - 2 * delta;
      REAL a = dy * s2 + in_h * (delta - d0);
      REAL b = in_h * d0 - dy * s2;
      REAL c = -delta * dy;
      REAL disc = b * b - 4 * a * c;
      REAL root = (2 * c) / (-b - (REAL)SQRT(disc));
      outputs[e] = root * in_w + in_cw;
      REAL tomt = root * (1 - root);
      REAL den = delta + s2 * tomt;
      REAL num = delta * delta * (d1 * root * root + 2 * delta * tomt + d0 * (1 - root) * (1 - root));
      logabsdet[e] = -((REAL)LOG(num) - 2 * (REAL)LOG(den));
    } else { /* transforms.py:178-193 */
      REAL theta = (x - in_cw) / in_w;
      REAL tomt = theta * (1 - theta);
      REAL numr = in_h * (delta * theta * theta + d0 * tomt);
      REAL s2 = d0 + d1 - 2 * delta;
      REAL den = delta + s2 * tomt;
      outputs[e] = in_ch + numr / den;
      REAL num = delta * delta * (d1 * theta * theta + 2 * delta * tomt + d0 * (1 - theta) * (1 - theta));
      logabsdet[e] = (REAL)LOG(num) - 2 * (REAL)LOG(den);
    }
  }
}

#undef FN
#undef FN1
#undef FN2
