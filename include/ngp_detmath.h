/* ngp_detmath.h — deterministic elementary functions shared by the CUDA hot path and the CPU oracle.
 *
 * The reference evaluates its ray-march stepping with logf/expf compiled under --use_fast_math
 * (CMakeLists.txt:88; include/neural-graphics-primitives/nerf_device.cuh:379-441), which is neither
 * reproducible on a CPU nor across GPU generations.  To make "per-ray sample counts" an integer,
 * bit-exact quantity that a CPU oracle can check, every transcendental on the march / colour path
 * goes through the functions below.  They use only IEEE-754 basic operations (add, mul, fma with
 * a single rounding, floor, frexp/ldexp), which round identically on sm_100a and x86-64, so the
 * CUDA kernels and the oracle agree bit for bit.  Accuracy: <= 2 ulp against a correctly rounded
 * result over the ranges used (tested in tests/test_detmath.py), i.e. tighter than the fast-math
 * intrinsics the reference uses.
 *
 * Polynomials: classic Cephes single-precision minimax fits (public domain, S. Moshier).
 *
 * Build rules that keep the two sides identical:
 *   - device code including this header is compiled with  -fmad=false
 *   - host   code including this header is compiled with  -ffp-contract=off
 *   - fused multiply-adds appear only as explicit NGP_FMA().
 */
#ifndef NGP_DETMATH_H
#define NGP_DETMATH_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NGP_DM_HD __host__ __device__ __forceinline__
#else
#define NGP_DM_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define NGP_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define NGP_MUL(a, b) __fmul_rn((a), (b))
#define NGP_ADD(a, b) __fadd_rn((a), (b))
#define NGP_SUB(a, b) __fsub_rn((a), (b))
#define NGP_DIV(a, b) __fdiv_rn((a), (b))
#define NGP_SQRT(a) __fsqrt_rn((a))
#else
#define NGP_FMA(a, b, c) fmaf((a), (b), (c))
#define NGP_MUL(a, b) ((a) * (b))
#define NGP_ADD(a, b) ((a) + (b))
#define NGP_SUB(a, b) ((a) - (b))
#define NGP_DIV(a, b) ((a) / (b))
#define NGP_SQRT(a) sqrtf((a))
#endif

/* natural logarithm, x > 0 and finite (callers guarantee it); x <= 0 returns -inf like logf(0). */
NGP_DM_HD float ngp_logf(float x) {
	if (!(x > 0.0f)) return -INFINITY;
	int e;
	float m = frexpf(x, &e); /* m in [0.5, 1) */
	if (m < 0.707106781186547524f) {
		e -= 1;
		m = NGP_SUB(NGP_ADD(m, m), 1.0f);
	} else {
		m = NGP_SUB(m, 1.0f);
	}
	const float z = NGP_MUL(m, m);
	float y = 7.0376836292E-2f;
	y = NGP_FMA(y, m, -1.1514610310E-1f);
	y = NGP_FMA(y, m, 1.1676998740E-1f);
	y = NGP_FMA(y, m, -1.2420140846E-1f);
	y = NGP_FMA(y, m, 1.4249322787E-1f);
	y = NGP_FMA(y, m, -1.6668057665E-1f);
	y = NGP_FMA(y, m, 2.0000714765E-1f);
	y = NGP_FMA(y, m, -2.4999993993E-1f);
	y = NGP_FMA(y, m, 3.3333331174E-1f);
	y = NGP_MUL(NGP_MUL(y, m), z);
	const float fe = (float)e;
	y = NGP_FMA(-2.12194440e-4f, fe, y);
	y = NGP_FMA(-0.5f, z, y);
	float r = NGP_ADD(m, y);
	r = NGP_FMA(0.693359375f, fe, r);
	return r;
}

/* e^x, clamped to [0, FLT_MAX] outside |x| < ~88. */
NGP_DM_HD float ngp_expf(float x) {
	if (x > 88.72283905206835f) return INFINITY;
	if (x < -103.0f) return 0.0f;
	float z = floorf(NGP_FMA(1.44269504088896341f, x, 0.5f));
	x = NGP_FMA(z, -0.693359375f, x);
	x = NGP_FMA(z, 2.12194440e-4f, x);
	const int n = (int)z;
	const float x2 = NGP_MUL(x, x);
	float p = 1.9875691500E-4f;
	p = NGP_FMA(p, x, 1.3981999507E-3f);
	p = NGP_FMA(p, x, 8.3334519073E-3f);
	p = NGP_FMA(p, x, 4.1665795894E-2f);
	p = NGP_FMA(p, x, 1.6666665459E-1f);
	p = NGP_FMA(p, x, 5.0000001201E-1f);
	p = NGP_FMA(p, x2, x);
	p = NGP_ADD(p, 1.0f);
	return ldexpf(p, n);
}

/* x^y for x > 0 (used for the sRGB transfer curves, common_device.cuh:61-103). */
NGP_DM_HD float ngp_powf(float x, float y) {
	if (!(x > 0.0f)) return 0.0f;
	return ngp_expf(NGP_MUL(y, ngp_logf(x)));
}

#endif /* NGP_DETMATH_H */
