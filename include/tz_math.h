/*
 * tz_math.h -- deterministic fp32 transcendental functions for the MCTS path.
 *
 * The reference leaves exp / log / pow to XLA (jax.nn.softmax and `**` in
 * core/evaluators/mcts/weighted_mcts.py:115,135 and core/evaluators/alphazero.py:74,76;
 * jnp.log in core/evaluators/mcts/action_selection.py:172).  XLA does not pin their
 * bit patterns, so this path DEFINES them: Cephes-style single precision kernels
 * written as individually rounded IEEE-754 binary32 multiplies and adds (no FMA
 * contraction), which therefore give the same bits in gcc (-ffp-contract=off),
 * numpy float32 and nvcc (-fmad=false / __fmul_rn,__fadd_rn).  Accuracy is ~2 ulp.
 *
 * Plain C, header-only; includable from C, C++ and CUDA.
 */
#ifndef TZ_MATH_H_
#define TZ_MATH_H_

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define TZ_HD __host__ __device__ __forceinline__
#else
#define TZ_HD static inline
#endif

/* Individually rounded mul/add. On the device use the _rn intrinsics, which ptxas
 * never contracts into FFMA; on the host rely on -ffp-contract=off. */
#if defined(__CUDA_ARCH__)
#define TZ_MUL(a, b) __fmul_rn((a), (b))
#define TZ_ADD(a, b) __fadd_rn((a), (b))
#define TZ_SUB(a, b) __fsub_rn((a), (b))
#define TZ_RINT(a) rintf(a)
#else
#include <math.h>
#define TZ_MUL(a, b) ((float)((float)(a) * (float)(b)))
#define TZ_ADD(a, b) ((float)((float)(a) + (float)(b)))
#define TZ_SUB(a, b) ((float)((float)(a) - (float)(b)))
#define TZ_RINT(a) nearbyintf(a)
#endif

#define TZ_FLT_MAX 3.402823466e+38f
#define TZ_FLT_MIN 1.175494351e-38f /* smallest normal == finfo(float32).tiny */
#define TZ_FLT_EPS 1.1920928955078125e-7f /* finfo(float32).eps */

TZ_HD float tz_bits_to_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

TZ_HD uint32_t tz_float_to_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}

/* exp(x) for x <= ~88.  x < -86 flushes to +0 (keeps 2^k normal).  Written without an early return: straight-line code,
 * so that independent calls interleave on the device instead of running one after the other behind a branch each. */
TZ_HD float tz_expf(float x) {
  const int flush = x < -86.0f;
  if (x > 88.0f) x = 88.0f;
  if (flush) x = -86.0f; /* (any in-range argument: the result is discarded) */
  float k = TZ_RINT(TZ_MUL(x, 1.44269504088896341f));
  float r = TZ_SUB(x, TZ_MUL(k, 0.693359375f));
  r = TZ_SUB(r, TZ_MUL(k, -2.12194440e-4f));
  float p = 1.9875691500e-4f;
  p = TZ_ADD(TZ_MUL(p, r), 1.3981999507e-3f);
  p = TZ_ADD(TZ_MUL(p, r), 8.3334519073e-3f);
  p = TZ_ADD(TZ_MUL(p, r), 4.1665795894e-2f);
  p = TZ_ADD(TZ_MUL(p, r), 1.6666665459e-1f);
  p = TZ_ADD(TZ_MUL(p, r), 5.0000001201e-1f);
  float r2 = TZ_MUL(r, r);
  float y = TZ_ADD(TZ_ADD(TZ_MUL(p, r2), r), 1.0f);
  int ki = (int)k;
  /* two-step scaling keeps both factors normal for ki in [-125, 128] */
  int k1 = ki / 2, k2 = ki - k1;
  float s1 = tz_bits_to_float((uint32_t)(k1 + 127) << 23);
  float s2 = tz_bits_to_float((uint32_t)(k2 + 127) << 23);
  const float res = TZ_MUL(TZ_MUL(y, s1), s2);
  return flush ? 0.0f : res;
}

/* log(x) for normal x > 0.  x < FLT_MIN is treated as FLT_MIN. */
TZ_HD float tz_logf(float x) {
  if (x < TZ_FLT_MIN) x = TZ_FLT_MIN;
  uint32_t u = tz_float_to_bits(x);
  int e = (int)((u >> 23) & 0xffu) - 126; /* x = m * 2^e, m in [0.5, 1) */
  float m = tz_bits_to_float((u & 0x007fffffu) | 0x3f000000u);
  if (m < 0.707106781186547524f) {
    e -= 1;
    m = TZ_SUB(TZ_ADD(m, m), 1.0f);
  } else {
    m = TZ_SUB(m, 1.0f);
  }
  float z = TZ_MUL(m, m);
  float y = 7.0376836292e-2f;
  y = TZ_ADD(TZ_MUL(y, m), -1.1514610310e-1f);
  y = TZ_ADD(TZ_MUL(y, m), 1.1676998740e-1f);
  y = TZ_ADD(TZ_MUL(y, m), -1.2420140846e-1f);
  y = TZ_ADD(TZ_MUL(y, m), 1.4249322787e-1f);
  y = TZ_ADD(TZ_MUL(y, m), -1.6668057665e-1f);
  y = TZ_ADD(TZ_MUL(y, m), 2.0000714765e-1f);
  y = TZ_ADD(TZ_MUL(y, m), -2.4999993993e-1f);
  y = TZ_ADD(TZ_MUL(y, m), 3.3333331174e-1f);
  y = TZ_MUL(TZ_MUL(y, m), z);
  float fe = (float)e;
  y = TZ_ADD(y, TZ_MUL(-2.12194440e-4f, fe));
  y = TZ_ADD(y, TZ_MUL(-0.5f, z));
  float r = TZ_ADD(m, y);
  r = TZ_ADD(r, TZ_MUL(0.693359375f, fe));
  return r;
}

/* x ** y for x >= 0, y > 0 (weighted_mcts.py:115, mcts.py:290).  y == 1 is the identity. */
TZ_HD float tz_powf(float x, float y) {
  if (y == 1.0f) return x;
  if (x < TZ_FLT_MIN) return 0.0f;
  return tz_expf(TZ_MUL(y, tz_logf(x)));
}

/* murmur3 fmix32: the synthetic game's only source of pseudo-randomness. */
TZ_HD uint32_t tz_mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85ebca6bu;
  x ^= x >> 13;
  x *= 0xc2b2ae35u;
  x ^= x >> 16;
  return x;
}

#endif /* TZ_MATH_H_ */
