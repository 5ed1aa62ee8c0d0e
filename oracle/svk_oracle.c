/*
 * svk_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the operators on SMART-Vocoder's SynthesizerTrn.infer path
 * (reference models.py:331-339 and everything it calls), in fp32 and fp64.  It exists so the
 * CUDA path can be checked on hardware where /root/reference is not mounted.
 *
 * Pinning: the reference ships no tests or golden vectors.  This oracle is pinned
 * instead against outputs of the reference itself, produced in the build container by
 * tests/golden/make_golden.py (which imports /root/reference/models.py) and committed under
 * tests/golden/*.npz; tests/test_oracle.py re-checks the oracle against those files.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define SVKO_MAX_BINS 32

#define REAL float
#define ACC float
#define SUFFIX f32
#define SQRT sqrtf
#define TANH tanhf
#define EXP expf
#define LOG logf
#define LOG1P log1pf
#include "svk_oracle_impl.h"
#undef REAL
#undef ACC
#undef SUFFIX
#undef SQRT
#undef TANH
#undef EXP
#undef LOG
#undef LOG1P

#define REAL double
#define ACC double
#define SUFFIX f64
#define SQRT sqrt
#define TANH tanh
#define EXP exp
#define LOG log
#define LOG1P log1p
#include "svk_oracle_impl.h"

/* commons.sequence_mask (reference commons.py:121-125) cast to float as models.py:40 does:
 * mask[b,t] = (t < length[b]).  Integer compare -- bit-exact by construction. */
void svko_sequence_mask(const int64_t *lengths, int64_t B, int64_t T, float *mask) {
  for (int64_t b = 0; b < B; ++b)
    for (int64_t t = 0; t < T; ++t) mask[b * T + t] = (t < lengths[b]) ? 1.0f : 0.0f;
}

int svko_abi_version(void) { return 1; }
