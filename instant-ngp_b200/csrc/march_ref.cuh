// march_ref.cuh — the ray-generation / stepping arithmetic of the sample generator in the REFERENCE BUILD's floating-point form.
//
// The reference compiles src/testbed_nerf.cu with --use_fast_math (CMakeLists.txt:88): every `/` is div.approx.ftz, sqrtf is
// sqrt.approx.ftz, logf / expf / sinf are the MUFU-based __logf / __expf / __sinf, and a*b+c contracts to one fma wherever the
// expression tree offers it.  Per-ray sample counts are decided at voxel faces by the last bits of t, so "bit-exact sample counts
// against the reference build" (BASELINE.json north_star) needs the same operations in the same order, not a more accurate march.
// This header restates the functions generate_training_samples_nerf (src/testbed_nerf.cu:691-849) runs through — with the SAME
// expression trees after the reference's vec / mat templates are expanded (tiny-cuda-nn vec.h: cwise ops :223-300, reductions
// :417-434 start from 0 and add term by term, mat * vec :597-607 likewise) — and the translation unit that includes it
// (march_ref.cu) is compiled with the reference's flags, so nvcc makes the same contraction decisions for both.
//
// Everything else in the library (loss, render, the deterministic twin of this file: march.cuh, which the CPU oracle can follow
// bit for bit) is unaffected: the two flavours share one kernel body (gen_kernel.cuh) and differ only in this arithmetic.
//
// Each function cites what it restates.  Device only.
#pragma once

#include <cfloat>

#include "common.cuh"
#include "march.cuh"

namespace ngpb {
namespace refm {

// nerf_device.cuh:32-36 (constexpr in the reference: folded by the front end, in float)
__device__ __forceinline__ constexpr float SQRT3() { return 1.73205080757f; }
__device__ __forceinline__ constexpr float STEPSIZE() { return SQRT3() / 1024u; }
__device__ __forceinline__ constexpr float MIN_CONE_STEPSIZE() { return STEPSIZE(); }
__device__ __forceinline__ constexpr float MAX_CONE_STEPSIZE() { return STEPSIZE() * (1 << (8 - 1)) * 1024u / 128u; }

// ---- tiny-cuda-nn vec.h shapes ------------------------------------------------------------------------------------------
// REDUCTION_OP(dot / length2): result = 0; result += a[i] * b[i]
__device__ __forceinline__ float dot(V3 a, V3 b) {
	float result = 0.0f;
	result += a.x * b.x;
	result += a.y * b.y;
	result += a.z * b.z;
	return result;
}
__device__ __forceinline__ float length2(V3 a) {
	float result = 0.0f;
	result += a.x * a.x;
	result += a.y * a.y;
	result += a.z * a.z;
	return result;
}
// normalize (vec.h:467-476): v / sqrt(length2(v))
__device__ __forceinline__ V3 normalize(V3 v) {
	const float len = sqrtf(refm::length2(v));
	if (len <= 0.0f) return V3{1.0f, 0.0f, 0.0f};
	return V3{v.x / len, v.y / len, v.z / len};
}
// tmat<3,3> * tvec<3> (vec.h:597-607): result = 0; for column i, row j: result[j] += m[i][j] * v[i].  m: column major [c * 3 + r]
__device__ __forceinline__ V3 mat3_mul(const float* m, V3 v) {
	float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f;
	r0 += m[0] * v.x; r1 += m[1] * v.x; r2 += m[2] * v.x;
	r0 += m[3] * v.y; r1 += m[4] * v.y; r2 += m[5] * v.y;
	r0 += m[6] * v.z; r1 += m[7] * v.z; r2 += m[8] * v.z;
	return V3{r0, r1, r2};
}

// ---- quaternions (vec.h:1075-1190) ---------------------------------------------------------------------------------------
struct Quat {
	float w, x, y, z;
};
// tquat(const tmat<3,3>&) (vec.h:1079-1110); m[c][r] = m[c * 3 + r]
__device__ inline Quat quat_from_mat3(const float* m) {
	const float m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
	Quat q;
	const float tr = m00 + m11 + m22;
	if (tr > 0.0f) {
		const float S = sqrtf(tr + 1.0f) * 2.0f;
		q.w = 0.25f * S;
		q.x = (m12 - m21) / S;
		q.y = (m20 - m02) / S;
		q.z = (m01 - m10) / S;
	} else if (m00 > m11 && m00 > m22) {
		const float S = sqrtf(1.0f + m00 - m11 - m22) * 2.0f;
		q.w = (m12 - m21) / S;
		q.x = 0.25f * S;
		q.y = (m10 + m01) / S;
		q.z = (m20 + m02) / S;
	} else if (m11 > m22) {
		const float S = sqrtf(1.0f + m11 - m00 - m22) * 2.0f;
		q.w = (m20 - m02) / S;
		q.x = (m10 + m01) / S;
		q.y = 0.25f * S;
		q.z = (m21 + m12) / S;
	} else {
		const float S = sqrtf(1.0f + m22 - m00 - m11) * 2.0f;
		q.w = (m01 - m10) / S;
		q.x = (m20 + m02) / S;
		q.y = (m21 + m12) / S;
		q.z = 0.25f * S;
	}
	return q;
}
__device__ __forceinline__ float qdot(Quat a, Quat b) { return (a.w * b.w + a.x * b.x) + (a.y * b.y + a.z * b.z); }
__device__ __forceinline__ Quat qnormalize(Quat a) {
	const float len = sqrtf(qdot(a, a));
	if (len <= 0.0f) return Quat{1.0f, 0.0f, 0.0f, 0.0f};
	return Quat{a.w / len, a.x / len, a.y / len, a.z / len};
}
// slerp (vec.h:1147-1168)
__device__ inline Quat qslerp(Quat x, Quat y, float t) {
	Quat z = y;
	float cos_theta = qdot(x, y);
	if (cos_theta < 0.0f) {
		z = Quat{-y.w, -y.x, -y.y, -y.z};
		cos_theta = -cos_theta;
	}
	if (cos_theta > 1.0f - FLT_EPSILON) {
		// mix(x, z, t) = x * (1 - t) + z * t
		const float s = 1.0f - t;
		return Quat{x.w * s + z.w * t, x.x * s + z.x * t, x.y * s + z.y * t, x.z * s + z.z * t};
	}
	const float angle = acosf(cos_theta);
	const float sa = sinf((1.0f - t) * angle), sb = sinf(t * angle), sd = sinf(angle);
	return Quat{(sa * x.w + sb * z.w) / sd, (sa * x.x + sb * z.x) / sd, (sa * x.y + sb * z.y) / sd, (sa * x.z + sb * z.z) / sd};
}
// to_mat3 (vec.h:1186-1196); out: column major [c * 3 + r]
__device__ inline void quat_to_mat3(Quat q, float* out) {
	const float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z;
	const float qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z;
	const float qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
	out[0] = 1.0f - 2.0f * (qyy + qzz); out[1] = 2.0f * (qxy + qwz); out[2] = 2.0f * (qxz - qwy);
	out[3] = 2.0f * (qxy - qwz); out[4] = 1.0f - 2.0f * (qxx + qzz); out[5] = 2.0f * (qyz + qwx);
	out[6] = 2.0f * (qxz + qwy); out[7] = 2.0f * (qyz - qwx); out[8] = 1.0f - 2.0f * (qxx + qyy);
}
// get_xform_given_rolling_shutter -> camera_slerp(start, end, pixel_t) (common_device.cuh:665-674) with start == end and a zero
// rolling shutter: pixel_t = 0, but the rotation still goes matrix -> quaternion -> matrix (slerp(mat3, mat3, t), vec.h:1199), which
// re-orthonormalises it and moves its entries by ~1e-7; the origin comes out of mix(a[3], b[3], 0) unchanged.  The result depends
// on the view only, so it is evaluated once per view (k_ref_view_rotations) instead of once per ray.
__device__ inline void camera_slerp_rotation(const float* xform, float pixel_t, float* rot_out) {
	const Quat a = qnormalize(quat_from_mat3(xform));
	const Quat b = qnormalize(quat_from_mat3(xform));
	quat_to_mat3(qnormalize(qslerp(a, b, pixel_t)), rot_out);
}

// ---- lens (common_device.cuh:268-282, 307-353) ---------------------------------------------------------------------------
__device__ __forceinline__ void opencv_lens_distortion_delta(const float* extra_params, const float u, const float v, float* du, float* dv) {
	const float k1 = extra_params[0];
	const float k2 = extra_params[1];
	const float p1 = extra_params[2];
	const float p2 = extra_params[3];
	const float u2 = u * u;
	const float uv = u * v;
	const float v2 = v * v;
	const float r2 = u2 + v2;
	const float radial = k1 * r2 + k2 * r2 * r2;
	*du = u * radial + 2.0f * p1 * uv + p2 * (r2 + 2.0f * u2);
	*dv = v * radial + 2.0f * p2 * uv + p1 * (r2 + 2.0f * v2);
}
__device__ inline void iterative_opencv_lens_undistortion(const float* params, float* u, float* v) {
	const uint32_t kNumIterations = 100;
	const float kMaxStepNorm = 1e-10f;
	const float kRelStepSize = 1e-6f;
	float J00, J10, J01, J11;   // J[c][r]
	const float x0_0 = *u, x0_1 = *v;
	float x_0 = *u, x_1 = *v;
	float dx_0, dx_1, dx_0b_0, dx_0b_1, dx_0f_0, dx_0f_1, dx_1b_0, dx_1b_1, dx_1f_0, dx_1f_1;
	for (uint32_t i = 0; i < kNumIterations; ++i) {
		const float step0 = fmaxf(FLT_EPSILON, fabsf(kRelStepSize * x_0));
		const float step1 = fmaxf(FLT_EPSILON, fabsf(kRelStepSize * x_1));
		opencv_lens_distortion_delta(params, x_0, x_1, &dx_0, &dx_1);
		opencv_lens_distortion_delta(params, x_0 - step0, x_1, &dx_0b_0, &dx_0b_1);
		opencv_lens_distortion_delta(params, x_0 + step0, x_1, &dx_0f_0, &dx_0f_1);
		opencv_lens_distortion_delta(params, x_0, x_1 - step1, &dx_1b_0, &dx_1b_1);
		opencv_lens_distortion_delta(params, x_0, x_1 + step1, &dx_1f_0, &dx_1f_1);
		J00 = 1 + (dx_0f_0 - dx_0b_0) / (2 * step0);
		J10 = (dx_1f_0 - dx_1b_0) / (2 * step1);
		J01 = (dx_0f_1 - dx_0b_1) / (2 * step0);
		J11 = 1 + (dx_1f_1 - dx_1b_1) / (2 * step1);
		// inverse(J) = adjoint(J) / determinant(J) (vec.h:764-766, 798-803, 868-870); adjoint = {J11, -J01, -J10, J00} column major
		const float det = J00 * J11 - J01 * J10;
		const float i00 = J11 / det, i01 = -J01 / det, i10 = -J10 / det, i11 = J00 / det;   // inv[c][r]
		const float r_0 = x_0 + dx_0 - x0_0, r_1 = x_1 + dx_1 - x0_1;
		// mat2 * vec2: result = 0; result[j] += m[i][j] * v[i]
		float s_0 = 0.0f, s_1 = 0.0f;
		s_0 += i00 * r_0; s_1 += i01 * r_0;
		s_0 += i10 * r_1; s_1 += i11 * r_1;
		x_0 -= s_0;
		x_1 -= s_1;
		float l2 = 0.0f;
		l2 += s_0 * s_0;
		l2 += s_1 * s_1;
		if (l2 < kMaxStepNorm) break;
	}
	*u = x_0;
	*v = x_1;
}

// ---- bounding box (bounding_box.cuh:163-207) -----------------------------------------------------------------------------
__device__ inline void ray_intersect(const Aabb& b, V3 pos, V3 dir, float& tmin_out, float& tmax_out) {
	float tmin = (b.mn.x - pos.x) / dir.x;
	float tmax = (b.mx.x - pos.x) / dir.x;
	if (tmin > tmax) { const float s = tmin; tmin = tmax; tmax = s; }
	float tymin = (b.mn.y - pos.y) / dir.y;
	float tymax = (b.mx.y - pos.y) / dir.y;
	if (tymin > tymax) { const float s = tymin; tymin = tymax; tymax = s; }
	if (tmin > tymax || tymin > tmax) { tmin_out = FLT_MAX; tmax_out = FLT_MAX; return; }
	if (tymin > tmin) tmin = tymin;
	if (tymax < tmax) tmax = tymax;
	float tzmin = (b.mn.z - pos.z) / dir.z;
	float tzmax = (b.mx.z - pos.z) / dir.z;
	if (tzmin > tzmax) { const float s = tzmin; tzmin = tzmax; tzmax = s; }
	if (tmin > tzmax || tzmin > tmax) { tmin_out = FLT_MAX; tmax_out = FLT_MAX; return; }
	if (tzmin > tmin) tmin = tzmin;
	if (tzmax < tmax) tmax = tzmax;
	tmin_out = tmin;
	tmax_out = tmax;
}

// ---- exponential stepping (nerf_device.cuh:379-441): constants re-derived from cone_angle in place, as the reference writes it
// (they are loop invariant; the compiler hoists them in both builds)
__device__ __forceinline__ float to_stepping_space(float t, float cone_angle) {
	if (cone_angle <= 1e-5f) return t / MIN_CONE_STEPSIZE();
	const float log1p_c = logf(1.0f + cone_angle);
	const float a = (logf(MIN_CONE_STEPSIZE()) - logf(log1p_c)) / log1p_c;
	const float b = (logf(MAX_CONE_STEPSIZE()) - logf(log1p_c)) / log1p_c;
	const float at = expf(a * log1p_c);
	const float bt = expf(b * log1p_c);
	if (t <= at) {
		return (t - at) / MIN_CONE_STEPSIZE() + a;
	} else if (t <= bt) {
		return logf(t) / log1p_c;
	} else {
		return (t - bt) / MAX_CONE_STEPSIZE() + b;
	}
}
__device__ __forceinline__ float from_stepping_space(float n, float cone_angle) {
	if (cone_angle <= 1e-5f) return n * MIN_CONE_STEPSIZE();
	const float log1p_c = logf(1.0f + cone_angle);
	const float a = (logf(MIN_CONE_STEPSIZE()) - logf(log1p_c)) / log1p_c;
	const float b = (logf(MAX_CONE_STEPSIZE()) - logf(log1p_c)) / log1p_c;
	const float at = expf(a * log1p_c);
	const float bt = expf(b * log1p_c);
	if (n <= a) {
		return (n - a) * MIN_CONE_STEPSIZE() + at;
	} else if (n <= b) {
		return expf(n * log1p_c);
	} else {
		return (n - b) * MAX_CONE_STEPSIZE() + bt;
	}
}
__device__ __forceinline__ float advance_n_steps(float t, float cone_angle, float n) { return refm::from_stepping_space(refm::to_stepping_space(t, cone_angle) + n, cone_angle); }
__device__ __forceinline__ float calc_dt(float t, float cone_angle) { return refm::advance_n_steps(t, cone_angle, 1.0f) - t; }

__device__ __forceinline__ float sign1(float v) { return copysignf(1.0f, v); }   // vec.h:182

// nerf_device.cuh:360-368
__device__ __forceinline__ float distance_to_next_voxel(V3 pos, V3 dir, V3 idir, float res) {
	const V3 p = V3{res * (pos.x - 0.5f), res * (pos.y - 0.5f), res * (pos.z - 0.5f)};
	const float tx = (floorf(p.x + 0.5f + 0.5f * sign1(dir.x)) - p.x) * idir.x;
	const float ty = (floorf(p.y + 0.5f + 0.5f * sign1(dir.y)) - p.y) * idir.y;
	const float tz = (floorf(p.z + 0.5f + 0.5f * sign1(dir.z)) - p.z) * idir.z;
	const float t = fminf(fminf(tx, ty), tz);
	return fmaxf(t / res, 0.0f);
}
// nerf_device.cuh:431-441
__device__ __forceinline__ float advance_to_next_voxel(float t, float cone_angle, V3 pos, V3 dir, V3 idir, uint32_t mip) {
	const float res = scalbnf(128.0f, -(int)mip);
	float t_target = t + refm::distance_to_next_voxel(pos, dir, idir, res);
	t = refm::to_stepping_space(t, cone_angle);
	t_target = refm::to_stepping_space(t_target, cone_angle);
	return refm::from_stepping_space(t + ceilf(fmaxf(t_target - t, 0.5f)), cone_angle);
}

// nerf_device.cuh:266-272, 291-293, 307-310
__device__ __forceinline__ V3 warp_position(V3 pos, const Aabb& b) {
	return V3{(pos.x - b.mn.x) / (b.mx.x - b.mn.x), (pos.y - b.mn.y) / (b.mx.y - b.mn.y), (pos.z - b.mn.z) / (b.mx.z - b.mn.z)};
}
__device__ __forceinline__ V3 warp_direction(V3 dir) { return V3{(dir.x + 1.0f) * 0.5f, (dir.y + 1.0f) * 0.5f, (dir.z + 1.0f) * 0.5f}; }
__device__ __forceinline__ float warp_dt(float dt) {
	const float max_stepsize = MIN_CONE_STEPSIZE() * (1 << (8 - 1));
	return (dt - MIN_CONE_STEPSIZE()) / (max_stepsize - MIN_CONE_STEPSIZE());
}

}  // namespace refm
}  // namespace ngpb
