// march.cuh — scalar device helpers of the NeRF ray marcher.  Restates include/neural-graphics-primitives/nerf_device.cuh
// (stepping :360-495, occupancy lookup :317-358, activations :204-264, losses :75-143), common_device.cuh (sRGB :61-103, ray
// construction :413-490, image reads :776-872) and bounding_box.cuh (:163-221).
//
// This translation unit family is compiled with -fmad=false and takes all transcendentals from ngp_detmath.h, so the CPU
// oracle (oracle/ngp_oracle.c), which follows the same expression order, reproduces every value bit for bit.
#pragma once

#include "../../include/ngp_detmath.h"
#include "common.cuh"

namespace ngpb {

struct V3 {
	float x, y, z;
};
__host__ __device__ inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__host__ __device__ inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
__host__ __device__ inline V3 operator*(float s, V3 a) { return V3{a.x * s, a.y * s, a.z * s}; }
__host__ __device__ inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
__host__ __device__ inline float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__host__ __device__ inline float length3(V3 a) { return NGP_SQRT(dot3(a, a)); }
__host__ __device__ inline V3 normalize3(V3 a) {
	const float len = length3(a);
	if (len <= 0.0f) return V3{1.0f, 0.0f, 0.0f};
	return V3{a.x / len, a.y / len, a.z / len};
}

// constants (nerf_device.cuh:25-43)
__host__ __device__ inline constexpr float sqrt3() { return 1.73205080757f; }
__host__ __device__ inline constexpr float min_cone_stepsize() { return 1.73205080757f / 1024.0f; }
__host__ __device__ inline constexpr float max_cone_stepsize() { return 1.73205080757f; }  // STEPSIZE * 2^7 * 1024 / 128
__host__ __device__ inline constexpr float max_depth() { return 16384.0f; }                 // common.h MAX_DEPTH()
__host__ __device__ inline constexpr float min_optical_thickness() { return 0.01f; }
constexpr uint32_t N_MAX_RANDOM_SAMPLES_PER_RAY = 16;
constexpr uint32_t GRID_N_CELLS = 128u * 128u * 128u;

struct Aabb {
	V3 mn, mx;
	__host__ __device__ bool contains(V3 p) const { return p.x >= mn.x && p.x <= mx.x && p.y >= mn.y && p.y <= mx.y && p.z >= mn.z && p.z <= mx.z; }
};

// bounding_box.cuh:163-221 — slab test, returns (tmin, tmax) or (FLT_MAX, FLT_MAX)
__host__ __device__ inline void aabb_ray_intersect(const Aabb& b, V3 o, V3 d, float& tmin_out, float& tmax_out) {
	const float big = 3.402823466e+38f;
	float tmin = (b.mn.x - o.x) / d.x, tmax = (b.mx.x - o.x) / d.x;
	if (tmin > tmax) { const float t = tmin; tmin = tmax; tmax = t; }
	float tymin = (b.mn.y - o.y) / d.y, tymax = (b.mx.y - o.y) / d.y;
	if (tymin > tymax) { const float t = tymin; tymin = tymax; tymax = t; }
	if (tmin > tymax || tymin > tmax) { tmin_out = big; tmax_out = big; return; }
	if (tymin > tmin) tmin = tymin;
	if (tymax < tmax) tmax = tymax;
	float tzmin = (b.mn.z - o.z) / d.z, tzmax = (b.mx.z - o.z) / d.z;
	if (tzmin > tzmax) { const float t = tzmin; tzmin = tzmax; tzmax = t; }
	if (tmin > tzmax || tzmin > tmax) { tmin_out = big; tmax_out = big; return; }
	if (tzmin > tmin) tmin = tzmin;
	if (tzmax < tmax) tmax = tzmax;
	tmin_out = tmin;
	tmax_out = tmax;
}

// ---- exponential stepping (nerf_device.cuh:379-441) with the per-scene constants hoisted into ngp_march_consts -------
__host__ __device__ inline float to_stepping_space(float t, const ngp_march_consts& m) {
	if (m.cone_angle <= 1e-5f) return t / min_cone_stepsize();
	if (t <= m.at) return (t - m.at) / min_cone_stepsize() + m.a;
	if (t <= m.bt) return ngp_logf(t) / m.log1p_c;
	return (t - m.bt) / max_cone_stepsize() + m.b;
}
__host__ __device__ inline float from_stepping_space(float n, const ngp_march_consts& m) {
	if (m.cone_angle <= 1e-5f) return n * min_cone_stepsize();
	if (n <= m.a) return (n - m.a) * min_cone_stepsize() + m.at;
	if (n <= m.b) return ngp_expf(n * m.log1p_c);
	return (n - m.b) * max_cone_stepsize() + m.bt;
}
__host__ __device__ inline float advance_n_steps(float t, const ngp_march_consts& m, float n) { return from_stepping_space(to_stepping_space(t, m) + n, m); }
__host__ __device__ inline float calc_dt(float t, const ngp_march_consts& m) { return advance_n_steps(t, m, 1.0f) - t; }

__host__ __device__ inline float signf1(float v) { return copysignf(1.0f, v); }

// nerf_device.cuh:360-368
__host__ __device__ inline float distance_to_next_voxel(V3 pos, V3 dir, V3 idir, float res) {
	const V3 p = V3{res * (pos.x - 0.5f), res * (pos.y - 0.5f), res * (pos.z - 0.5f)};
	const float tx = (floorf(p.x + 0.5f + 0.5f * signf1(dir.x)) - p.x) * idir.x;
	const float ty = (floorf(p.y + 0.5f + 0.5f * signf1(dir.y)) - p.y) * idir.y;
	const float tz = (floorf(p.z + 0.5f + 0.5f * signf1(dir.z)) - p.z) * idir.z;
	const float t = fminf(fminf(tx, ty), tz);
	return fmaxf(t / res, 0.0f);
}
// nerf_device.cuh:431-441
__host__ __device__ inline float advance_to_next_voxel(float t, const ngp_march_consts& m, V3 pos, V3 dir, V3 idir, uint32_t mip) {
	const float res = scalbnf(128.0f, -(int)mip);
	float t_target = t + distance_to_next_voxel(pos, dir, idir, res);
	t = to_stepping_space(t, m);
	t_target = to_stepping_space(t_target, m);
	return from_stepping_space(t + ceilf(fmaxf(t_target - t, 0.5f)), m);
}
__host__ __device__ inline int clampi(int a, int lo, int hi) { return a < lo ? lo : (hi < a ? hi : a); }
// nerf_device.cuh:443-460
__host__ __device__ inline uint32_t mip_from_pos(V3 pos, uint32_t max_cascade) {
	if (max_cascade == 0) return 0;  // clamp(exponent + 1, 0, 0)
	int exponent;
	const float maxval = fmaxf(fmaxf(fabsf(pos.x - 0.5f), fabsf(pos.y - 0.5f)), fabsf(pos.z - 0.5f));
	frexpf(maxval, &exponent);
	return (uint32_t)clampi(exponent + 1, 0, (int)max_cascade);
}
__host__ __device__ inline uint32_t mip_from_dt(float dt, V3 pos, uint32_t max_cascade) {
	const uint32_t mip = mip_from_pos(pos, max_cascade);
	dt *= 2.0f * 128.0f;
	if (dt < 1.0f) return mip;
	int exponent;
	frexpf(dt, &exponent);
	return (uint32_t)clampi((int)mip, exponent, (int)max_cascade);
}
// nerf_device.cuh:317-341
__host__ __device__ inline uint32_t cascaded_grid_idx_at(V3 pos, uint32_t mip) {
	const float mip_scale = scalbnf(1.0f, -(int)mip);
	const float px = (pos.x - 0.5f) * mip_scale + 0.5f, py = (pos.y - 0.5f) * mip_scale + 0.5f, pz = (pos.z - 0.5f) * mip_scale + 0.5f;
	const int ix = (int)(px * 128.0f), iy = (int)(py * 128.0f), iz = (int)(pz * 128.0f);
	if (ix < 0 || ix >= 128 || iy < 0 || iy >= 128 || iz < 0 || iz >= 128) return 0xFFFFFFFFu;
	return morton3d((uint32_t)ix, (uint32_t)iy, (uint32_t)iz);
}
__host__ __device__ inline bool density_grid_occupied_at(V3 pos, const uint8_t* __restrict__ bitfield, uint32_t mip) {
	const uint32_t idx = cascaded_grid_idx_at(pos, mip);
	if (idx == 0xFFFFFFFFu) return false;
	return (bitfield[idx / 8 + (GRID_N_CELLS * mip) / 8] & (1u << (idx % 8))) != 0;
}

// One-entry cache of a march's last occupancy test.  Consecutive samples of a ray mostly fall into the same cell (the minimum
// step is a 4.6th of a mip-0 cell), and the test is a pure function of (cell, mip): re-using its result is exact.  Saves the
// Morton interleave and the bitfield load (ncu r1c: that load's latency was the largest single stall of the generator).
struct OccCache {
	uint32_t key = 0xFFFFFFFFu;
	bool occ = false;
};
__host__ __device__ inline bool density_grid_occupied_cached(V3 pos, const uint8_t* __restrict__ bitfield, uint32_t mip, OccCache& c) {
	const float mip_scale = scalbnf(1.0f, -(int)mip);
	const float px = (pos.x - 0.5f) * mip_scale + 0.5f, py = (pos.y - 0.5f) * mip_scale + 0.5f, pz = (pos.z - 0.5f) * mip_scale + 0.5f;
	const int ix = (int)(px * 128.0f), iy = (int)(py * 128.0f), iz = (int)(pz * 128.0f);
	if (ix < 0 || ix >= 128 || iy < 0 || iy >= 128 || iz < 0 || iz >= 128) return false;
	const uint32_t key = (uint32_t)ix | ((uint32_t)iy << 7) | ((uint32_t)iz << 14) | (mip << 21);
	if (key == c.key) return c.occ;
	const uint32_t idx = morton3d((uint32_t)ix, (uint32_t)iy, (uint32_t)iz);
	c.key = key;
	c.occ = (bitfield[idx / 8 + (GRID_N_CELLS * mip) / 8] & (1u << (idx % 8))) != 0;
	return c.occ;
}

// ---- activations (nerf_device.cuh:204-264) ---------------------------------------------------------------------------
__host__ __device__ inline float clampf(float a, float lo, float hi) { return a < lo ? lo : (hi < a ? hi : a); }
__host__ __device__ inline float logisticf(float x) { return 1.0f / (1.0f + ngp_expf(-x)); }
__host__ __device__ inline float network_to_rgb(float v, uint32_t act) {
	switch (act) {
		case NGP_ACT_NONE: return v;
		case NGP_ACT_RELU: return v > 0.0f ? v : 0.0f;
		case NGP_ACT_LOGISTIC: return logisticf(v);
		default: return ngp_expf(clampf(v, -10.0f, 10.0f));
	}
}
__host__ __device__ inline float network_to_rgb_derivative(float v, uint32_t act) {
	switch (act) {
		case NGP_ACT_NONE: return 1.0f;
		case NGP_ACT_RELU: return v > 0.0f ? 1.0f : 0.0f;
		case NGP_ACT_LOGISTIC: { const float d = logisticf(v); return d * (1.0f - d); }
		default: return ngp_expf(clampf(v, -10.0f, 10.0f));
	}
}
__host__ __device__ inline float network_to_density(float v, uint32_t act) {
	switch (act) {
		case NGP_ACT_NONE: return v;
		case NGP_ACT_RELU: return v > 0.0f ? v : 0.0f;
		case NGP_ACT_LOGISTIC: return logisticf(v);
		default: return ngp_expf(v);
	}
}
__host__ __device__ inline float network_to_density_derivative(float v, uint32_t act) {
	switch (act) {
		case NGP_ACT_NONE: return 1.0f;
		case NGP_ACT_RELU: return v > 0.0f ? 1.0f : 0.0f;
		case NGP_ACT_LOGISTIC: { const float d = logisticf(v); return d * (1.0f - d); }
		default: return ngp_expf(clampf(v, -15.0f, 15.0f));
	}
}

// ---- colour (common_device.cuh:61-103) -------------------------------------------------------------------------------
__host__ __device__ inline float srgb_to_linear(float s) { return s <= 0.04045f ? s / 12.92f : ngp_powf((s + 0.055f) / 1.055f, 2.4f); }
// common_device.cuh:71-77
__host__ __device__ inline float srgb_to_linear_derivative(float s) { return s <= 0.04045f ? 1.0f / 12.92f : 2.4f / 1.055f * ngp_powf((s + 0.055f) / 1.055f, 1.4f); }
__host__ __device__ inline float linear_to_srgb(float l) { return l < 0.0031308f ? 12.92f * l : 1.055f * ngp_powf(l, 0.41666f) - 0.055f; }

// warp helpers (nerf_device.cuh:266-315)
__host__ __device__ inline float warp_dt(float dt) {
	const float max_stepsize = min_cone_stepsize() * 128.0f;
	return (dt - min_cone_stepsize()) / (max_stepsize - min_cone_stepsize());
}
__host__ __device__ inline float unwarp_dt(float dt) {
	const float max_stepsize = min_cone_stepsize() * 128.0f;
	return dt * (max_stepsize - min_cone_stepsize()) + min_cone_stepsize();
}
__host__ __device__ inline V3 warp_position(V3 p, const Aabb& b) {
	return V3{(p.x - b.mn.x) / (b.mx.x - b.mn.x), (p.y - b.mn.y) / (b.mx.y - b.mn.y), (p.z - b.mn.z) / (b.mx.z - b.mn.z)};
}
__host__ __device__ inline V3 unwarp_position(V3 p, const Aabb& b) {
	return V3{b.mn.x + p.x * (b.mx.x - b.mn.x), b.mn.y + p.y * (b.mx.y - b.mn.y), b.mn.z + p.z * (b.mx.z - b.mn.z)};
}
__host__ __device__ inline V3 warp_direction(V3 d) { return V3{(d.x + 1.0f) * 0.5f, (d.y + 1.0f) * 0.5f, (d.z + 1.0f) * 0.5f}; }

// ---- camera (common_device.cuh:268-282 opencv distortion, :307-353 Newton undistortion, :413-490 uv_to_ray) ------------
__host__ __device__ inline void opencv_lens_distortion_delta(const float* p, float u, float v, float* du, float* dv) {
	const float k1 = p[0], k2 = p[1], p1 = p[2], p2 = p[3];
	const float u2 = u * u, uv = u * v, v2 = v * v, r2 = u2 + v2;
	const float radial = k1 * r2 + k2 * r2 * r2;
	*du = u * radial + 2.0f * p1 * uv + p2 * (r2 + 2.0f * u2);
	*dv = v * radial + 2.0f * p2 * uv + p1 * (r2 + 2.0f * v2);
}
__host__ __device__ inline void opencv_lens_undistort(const float* params, float* u, float* v) {
	const float x0 = *u, y0 = *v;
	float x = x0, y = y0;
	for (uint32_t i = 0; i < 100; ++i) {
		const float step0 = fmaxf(1.1920929e-07f, fabsf(1e-6f * x));
		const float step1 = fmaxf(1.1920929e-07f, fabsf(1e-6f * y));
		float dx, dy, dx0b, dy0b, dx0f, dy0f, dx1b, dy1b, dx1f, dy1f;
		opencv_lens_distortion_delta(params, x, y, &dx, &dy);
		opencv_lens_distortion_delta(params, x - step0, y, &dx0b, &dy0b);
		opencv_lens_distortion_delta(params, x + step0, y, &dx0f, &dy0f);
		opencv_lens_distortion_delta(params, x, y - step1, &dx1b, &dy1b);
		opencv_lens_distortion_delta(params, x, y + step1, &dx1f, &dy1f);
		// J is column major: J[c][r]
		const float j00 = 1.0f + (dx0f - dx0b) / (2.0f * step0);
		const float j10 = (dx1f - dx1b) / (2.0f * step1);
		const float j01 = (dy0f - dy0b) / (2.0f * step0);
		const float j11 = 1.0f + (dy1f - dy1b) / (2.0f * step1);
		const float rx = x + dx - x0, ry = y + dy - y0;
		const float det = j00 * j11 - j10 * j01;
		// inverse(J) * r
		const float sx = (j11 * rx - j10 * ry) / det;
		const float sy = (-j01 * rx + j00 * ry) / det;
		x -= sx;
		y -= sy;
		if (sx * sx + sy * sy < 1e-10f) break;
	}
	*u = x;
	*v = y;
}

// xform: 4x3 column major [c0 | c1 | c2 | origin]
__host__ __device__ inline V3 xform_col(const float* m, int c) { return V3{m[3 * c + 0], m[3 * c + 1], m[3 * c + 2]}; }
__host__ __device__ inline V3 xform_rotate(const float* m, V3 v) {
	return V3{(m[0] * v.x + m[3] * v.y) + m[6] * v.z, (m[1] * v.x + m[4] * v.y) + m[7] * v.z, (m[2] * v.x + m[5] * v.y) + m[8] * v.z};
}
// pinhole / OpenCV ray through uv in [0,1]^2; returns the UNNORMALISED direction like uv_to_ray.
__host__ __device__ inline void uv_to_ray(float u, float v, int w, int h, float fx, float fy, float cx, float cy, uint32_t lens_mode,
	const float* lens_params, const float* xform, V3& o, V3& d) {
	float dx = (u - cx) * (float)w / fx;
	float dy = (v - cy) * (float)h / fy;
	if (lens_mode == NGP_LENS_OPENCV) opencv_lens_undistort(lens_params, &dx, &dy);
	d = xform_rotate(xform, V3{dx, dy, 1.0f});
	o = xform_col(xform, 3);
}

__host__ __device__ inline int imin(int a, int b) { return a < b ? a : b; }
__host__ __device__ inline int imax(int a, int b) { return a > b ? a : b; }


// ------------------------------------------------------------------------------------------------------------------
// image access (common_device.cuh:776-872), shared by the sample generator (mask test) and the loss kernel (target colour)
// ------------------------------------------------------------------------------------------------------------------
struct Rgba {
	float r, g, b, a;
};
__device__ inline Rgba read_rgba_px(int px, int py, int w, const void* pixels, uint32_t type) {
	const size_t idx = (size_t)px + (size_t)py * (size_t)w;
	switch (type) {
		case NGP_IMAGE_BYTE: {
			const uint32_t val = reinterpret_cast<const uint32_t*>(pixels)[idx];
			if (val == 0x00FF00FFu) return Rgba{-1.0f, -1.0f, -1.0f, -1.0f};
			const float a = (float)((val >> 24) & 0xFFu) * (1.0f / 255.0f);
			Rgba o;
			o.r = srgb_to_linear((float)(val & 0xFFu) * (1.0f / 255.0f)) * a;
			o.g = srgb_to_linear((float)((val >> 8) & 0xFFu) * (1.0f / 255.0f)) * a;
			o.b = srgb_to_linear((float)((val >> 16) & 0xFFu) * (1.0f / 255.0f)) * a;
			o.a = a;
			return o;
		}
		case NGP_IMAGE_HALF: {
			const uint2 v = reinterpret_cast<const uint2*>(pixels)[idx];
			const __half2 lo = *reinterpret_cast<const __half2*>(&v.x), hi = *reinterpret_cast<const __half2*>(&v.y);
			return Rgba{__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi)};
		}
		case NGP_IMAGE_FLOAT: {
			const float4 v = reinterpret_cast<const float4*>(pixels)[idx];
			return Rgba{v.x, v.y, v.z, v.w};
		}
		default: return Rgba{5.0f, 0.0f, 0.0f, 1.0f};
	}
}
__device__ inline Rgba read_rgba_uv(float u, float v, int w, int h, const void* pixels, uint32_t type) {
	const int px = imin(imax((int)(u * (float)w), 0), w - 1);
	const int py = imin(imax((int)(v * (float)h), 0), h - 1);
	return read_rgba_px(px, py, w, pixels, type);
}
// is the pixel under (u, v) masked away (read_rgba(...).x < 0, testbed_nerf.cu:732-736)?  Only the sign of red matters, so the
// colour conversion of read_rgba is skipped: a Byte pixel is negative iff it is MASK_COLOR, half / float pixels carry their sign.
__device__ inline bool pixel_is_masked(float u, float v, int w, int h, const void* pixels, uint32_t type) {
	const int px = imin(imax((int)(u * (float)w), 0), w - 1);
	const int py = imin(imax((int)(v * (float)h), 0), h - 1);
	const size_t idx = (size_t)px + (size_t)py * (size_t)w;
	switch (type) {
		case NGP_IMAGE_BYTE: return reinterpret_cast<const uint32_t*>(pixels)[idx] == 0x00FF00FFu;
		case NGP_IMAGE_HALF: return __half2float(reinterpret_cast<const __half*>(pixels)[idx * 4]) < 0.0f;
		case NGP_IMAGE_FLOAT: return reinterpret_cast<const float*>(pixels)[idx * 4] < 0.0f;
		default: return false;
	}
}

// nerf_device.cuh:578-599 (uniform branch): neighbouring rays of a batch look at the same image
__host__ __device__ inline uint32_t image_idx(uint32_t base_idx, uint32_t n_rays, uint32_t n_images) {
	return ((base_idx * n_images) / n_rays) % n_images;  // uint32 arithmetic, as the reference
}

// nerf_device.cuh:553-576 (no error-map CDF)
__device__ inline void random_image_pos_training(Pcg32& rng, int w, int h, bool snap, float& u, float& v) {
	u = rng.next_float();
	v = rng.next_float();
	if (snap) {
		u = ((float)imin(imax((int)(u * (float)w), 0), w - 1) + 0.5f) / (float)w;
		v = ((float)imin(imax((int)(v * (float)h), 0), h - 1) + 0.5f) / (float)h;
	}
}

}  // namespace ngpb
