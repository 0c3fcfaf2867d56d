// fm_math.h — scalar float math with a FIXED, documented operation order that
// evaluates bit-identically on the host (gcc, -ffp-contract=off) and on the
// device (nvcc; every product/sum that must not fuse goes through *_rn
// intrinsics, every fused multiply-add is an explicit fmaf).
//
// Why it exists: the engine reproduces serial nonlinear loops (19 kHz pilot PLL,
// RDS Costas loop, AGCs) whose hard decisions (stereo lock block, RDS bits) must
// equal the CPU oracle's. libm and libdevice transcendentals differ in the last
// ulp, so the engine uses these kernels; the oracle is built twice — once with
// libm (faithful to the reference's std::sin/std::cos/cargf/expf/logf call
// sites) and once with these — and tests/ bounds the distance between the two.
//
// Accuracy (measured in tests/test_fm_math.py against libm in float64):
//   fm_sincosf  |x| <= 16     : <= 2 ulp      (Cody-Waite pi/2 3-term + Cephes minimax)
//   fm_atan2f                 : <= 4 ulp      (Cephes atanf on min/max quotient)
//   fm_expf     |x| <= 87     : <= 2 ulp
//   fm_logf     x normal > 0  : <= 2 ulp
#ifndef FMGPU_FM_MATH_H_
#define FMGPU_FM_MATH_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FM_HD __host__ __device__ __forceinline__
#else
#define FM_HD static inline
#endif

// ---- unfused primitives ---------------------------------------------------
#if defined(__CUDA_ARCH__)
#define FM_MUL(a, b) __fmul_rn((a), (b))
#define FM_ADD(a, b) __fadd_rn((a), (b))
#define FM_SUB(a, b) __fsub_rn((a), (b))
#define FM_DIV(a, b) __fdiv_rn((a), (b))
#define FM_SQRT(a) __fsqrt_rn((a))
#define FM_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
// Host translation units that include this header MUST be compiled with
// -ffp-contract=off (oracle/Makefile and the engine build do so).
#define FM_MUL(a, b) ((a) * (b))
#define FM_ADD(a, b) ((a) + (b))
#define FM_SUB(a, b) ((a) - (b))
#define FM_DIV(a, b) ((a) / (b))
#define FM_SQRT(a) sqrtf((a))
#define FM_FMA(a, b, c) fmaf((a), (b), (c))
#endif

FM_HD int32_t fm_f2i(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(f);
#else
  int32_t i;
  memcpy(&i, &f, sizeof(i));
  return i;
#endif
}

FM_HD float fm_i2f(int32_t i) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  float f;
  memcpy(&f, &i, sizeof(f));
  return f;
#endif
}

FM_HD float fm_clampf(float v, float lo, float hi) {
  // same selection order as std::clamp
  return (v < lo) ? lo : ((hi < v) ? hi : v);
}

// round half away from zero, exact (roundf semantics)
FM_HD float fm_roundf(float x) {
  float r = truncf(x);
  float d = FM_SUB(x, r);  // exact
  if (fabsf(d) >= 0.5f) {
    r = FM_ADD(r, copysignf(1.0f, x));
  }
  return r;
}

// ---- sin / cos ------------------------------------------------------------
// Valid for |x| up to a few tens of radians (the engine only passes phases in
// [-2pi, 2pi]).
FM_HD void fm_sincosf(float x, float *s, float *c) {
  const float kf = rintf(FM_MUL(x, 0.636619772367581343f));  // x * 2/pi
  const int k = (int)kf;
  float r = FM_FMA(kf, -1.5703125f, x);
  r = FM_FMA(kf, -4.837512969970703125e-4f, r);
  r = FM_FMA(kf, -7.54978995489188216e-8f, r);
  const float z = FM_MUL(r, r);
  float ps = FM_FMA(z, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = FM_FMA(ps, z, -1.6666654611e-1f);
  const float sr = FM_FMA(FM_MUL(ps, z), r, r);
  float pc = FM_FMA(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = FM_FMA(pc, z, 4.166664568298827e-2f);
  const float cr = FM_FMA(FM_MUL(pc, z), z, FM_FMA(z, -0.5f, 1.0f));
  // quadrant selection without branches (lanes of a warp sit in different quadrants)
  const bool swap = (k & 1) != 0;
  const float s0 = swap ? cr : sr;
  const float c0 = swap ? sr : cr;
  *s = (k & 2) ? -s0 : s0;
  *c = ((k + 1) & 2) ? -c0 : c0;
}

FM_HD float fm_sinf(float x) {
  float s, c;
  fm_sincosf(x, &s, &c);
  return s;
}

FM_HD float fm_cosf(float x) {
  float s, c;
  fm_sincosf(x, &s, &c);
  return c;
}

// ---- atan2 ----------------------------------------------------------------
FM_HD float fm_atan2f(float y, float x) {
  const float kPi = 3.14159265358979323846f;
  const float kPi2 = 1.57079632679489661923f;
  const float kPi4 = 0.78539816339744830962f;
  const float ax = fabsf(x);
  const float ay = fabsf(y);
  if (ay == 0.0f) {
    const float r0 = (fm_f2i(x) < 0) ? kPi : 0.0f;  // sign bit of x, -0 included
    return copysignf(r0, y);
  }
  if (ax == 0.0f) {
    return copysignf(kPi2, y);
  }
  const float mx = (ax > ay) ? ax : ay;
  const float mn = (ax > ay) ? ay : ax;
  const float a = FM_DIV(mn, mx);  // in [0, 1]
  float y0, t;
  if (a > 0.4142135623730950f) {
    y0 = kPi4;
    t = FM_DIV(FM_SUB(a, 1.0f), FM_ADD(a, 1.0f));
  } else {
    y0 = 0.0f;
    t = a;
  }
  const float z = FM_MUL(t, t);
  float p = FM_FMA(z, 8.05374449538e-2f, -1.38776856032e-1f);
  p = FM_FMA(p, z, 1.99777106478e-1f);
  p = FM_FMA(p, z, -3.33329491539e-1f);
  float r = FM_ADD(y0, FM_FMA(FM_MUL(p, z), t, t));
  if (ay > ax) {
    r = FM_SUB(kPi2, r);
  }
  if (fm_f2i(x) < 0) {
    r = FM_SUB(kPi, r);
  }
  return copysignf(r, y);
}

// ---- exp / log ------------------------------------------------------------
FM_HD float fm_expf(float x) {
  if (fabsf(x) < 0.34f) {
    // k = rint(x * log2(e)) = 0: the reduction leaves r = x and both scale factors are 1.0f, so
    // this shortcut returns the bits of the general path below (the AGCs only ever call exp with
    // |x| = 0.5 * alpha * |log(energy)| << 0.34)
    const float z0 = FM_MUL(x, x);
    float p0 = FM_FMA(x, 1.9875691500e-4f, 1.3981999507e-3f);
    p0 = FM_FMA(p0, x, 8.3334519073e-3f);
    p0 = FM_FMA(p0, x, 4.1665795894e-2f);
    p0 = FM_FMA(p0, x, 1.6666665459e-1f);
    p0 = FM_FMA(p0, x, 5.0000001201e-1f);
    return FM_ADD(FM_FMA(p0, z0, x), 1.0f);
  }
  x = fm_clampf(x, -87.0f, 88.0f);
  const float kf = rintf(FM_MUL(x, 1.44269504088896341f));
  float r = FM_FMA(kf, -0.693359375f, x);
  r = FM_FMA(kf, 2.12194440e-4f, r);
  const float z = FM_MUL(r, r);
  float p = FM_FMA(r, 1.9875691500e-4f, 1.3981999507e-3f);
  p = FM_FMA(p, r, 8.3334519073e-3f);
  p = FM_FMA(p, r, 4.1665795894e-2f);
  p = FM_FMA(p, r, 1.6666665459e-1f);
  p = FM_FMA(p, r, 5.0000001201e-1f);
  const float y = FM_ADD(FM_FMA(p, z, r), 1.0f);
  const int k = (int)kf;
  // scale by 2^k in two exact steps so k = 128 / -126.. stay in range
  const int k1 = k / 2;
  const int k2 = k - k1;
  const float s1 = fm_i2f((k1 + 127) << 23);
  const float s2 = fm_i2f((k2 + 127) << 23);
  return FM_MUL(FM_MUL(y, s1), s2);
}

// x must be a positive normal float
FM_HD float fm_logf(float x) {
  int32_t ix = fm_f2i(x);
  int e = ((ix >> 23) & 0xff) - 126;                     // frexp exponent
  float m = fm_i2f((ix & 0x007fffff) | 0x3f000000);      // mantissa in [0.5, 1)
  if (m < 0.707106781186547524f) {
    e -= 1;
    m = FM_SUB(FM_ADD(m, m), 1.0f);
  } else {
    m = FM_SUB(m, 1.0f);
  }
  const float z = FM_MUL(m, m);
  float p = FM_FMA(m, 7.0376836292e-2f, -1.1514610310e-1f);
  p = FM_FMA(p, m, 1.1676998740e-1f);
  p = FM_FMA(p, m, -1.2420140846e-1f);
  p = FM_FMA(p, m, 1.4249322787e-1f);
  p = FM_FMA(p, m, -1.6668057665e-1f);
  p = FM_FMA(p, m, 2.0000714765e-1f);
  p = FM_FMA(p, m, -2.4999993993e-1f);
  p = FM_FMA(p, m, 3.3333331174e-1f);
  float y = FM_MUL(FM_MUL(p, m), z);
  const float fe = (float)e;
  y = FM_FMA(fe, -2.12194440e-4f, y);
  y = FM_FMA(z, -0.5f, y);
  float r = FM_ADD(m, y);
  r = FM_FMA(fe, 0.693359375f, r);
  return r;
}

#endif  // FMGPU_FM_MATH_H_
