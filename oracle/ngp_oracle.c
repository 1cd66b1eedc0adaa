/* ngp_oracle.c — CPU oracle for the ray-march / compositing / loss / occupancy-grid part of instant-ngp's NeRF path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library; the product (instant-ngp_b200/) never does.  It is a plain-C restatement of the reference algorithm,
 * written from the reference sources cited at each function, independent of the CUDA implementation; the only shared
 * pieces are the boundary header (struct layouts) and include/ngp_detmath.h (elementary functions), so that integer
 * results (ray indices, per-ray sample counts, occupancy bits) and the marched coordinates can be compared BIT-EXACTLY.
 * Build: gcc -O3 -march=native -ffp-contract=off -fopenmp (FMA contraction must stay off).
 *
 * Pinning status: the stepping / occupancy / camera helpers are validated against the reference's own NGP_HOST_DEVICE
 * functions compiled from /root/reference (oracle/ref/ref_host_harness.cu -> tests/golden/ref_host_*.bin); sample
 * generation, loss / compaction and the occupancy-grid upkeep against the reference's own kernels run on a B200
 * (oracle/ref/ref_nerf_harness.cu includes src/testbed_nerf.cu -> tests/golden/ref_nerf_*.npz,
 * tests/test_oracle_vs_reference_nerf.py); the render march / composite against the reference's NerfTracer kernels driven with an
 * analytic field (same harness, render cases).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ngp_b200.h"
#include "ngp_detmath.h"

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------------------------
 * pcg32 (tiny-cuda-nn/dependencies/pcg32/pcg32.h:56-165)
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct { uint64_t state, inc; } pcg32_t;
#define PCG32_MULT 0x5851f42d4c957f2dULL
static uint32_t pcg_next_uint(pcg32_t* r) {
	uint64_t old = r->state;
	r->state = old * PCG32_MULT + r->inc;
	uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
	uint32_t rot = (uint32_t)(old >> 59u);
	return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
}
static float pcg_next_float(pcg32_t* r) {
	union { uint32_t u; float f; } x;
	x.u = (pcg_next_uint(r) >> 9) | 0x3f800000u;
	return x.f - 1.0f;
}
static void pcg_advance(pcg32_t* r, uint64_t delta) {
	uint64_t cur_mult = PCG32_MULT, cur_plus = r->inc, acc_mult = 1u, acc_plus = 0u;
	while (delta > 0) {
		if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
		cur_plus = (cur_mult + 1) * cur_plus;
		cur_mult *= cur_mult;
		delta /= 2;
	}
	r->state = acc_mult * r->state + acc_plus;
}
EXPORT void orc_pcg32_seed(uint64_t initstate, uint64_t initseq, uint64_t* state, uint64_t* inc) {
	pcg32_t r; r.state = 0; r.inc = (initseq << 1u) | 1u;
	pcg_next_uint(&r); r.state += initstate; pcg_next_uint(&r);
	*state = r.state; *inc = r.inc;
}
EXPORT uint32_t orc_pcg32_next_uint(uint64_t* state, uint64_t inc) { pcg32_t r = {*state, inc}; uint32_t v = pcg_next_uint(&r); *state = r.state; return v; }
EXPORT void orc_pcg32_advance(uint64_t* state, uint64_t inc, uint64_t delta) { pcg32_t r = {*state, inc}; pcg_advance(&r, delta); *state = r.state; }

/* ------------------------------------------------------------------------------------------------------------------
 * small vector helpers
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct { float x, y, z; } v3;
static v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static float vdot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static float vlen(v3 a) { return sqrtf(vdot(a, a)); }
static v3 vnormalize(v3 a) { float l = vlen(a); if (l <= 0.0f) return V(1, 0, 0); return V(a.x / l, a.y / l, a.z / l); }

/* constants: nerf_device.cuh:25-43 */
#define GRIDSIZE 128
#define GRID_N_CELLS (128u * 128u * 128u)
#define N_CASCADES 8u
#define NERF_STEPS 1024u
static const float SQRT3 = 1.73205080757f;
static float MIN_CONE_STEPSIZE(void) { return SQRT3 / 1024.0f; }
static float MAX_CONE_STEPSIZE(void) { return (SQRT3 / 1024.0f) * 128.0f * 1024.0f / 128.0f; }
#define MAX_DEPTH 16384.0f
#define MIN_OPTICAL_THICKNESS 0.01f
#define N_MAX_RANDOM_SAMPLES_PER_RAY 16u

/* morton (tiny-cuda-nn/common_device.h:936-960) */
static uint32_t expand_bits(uint32_t v) {
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}
static uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) { return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2); }
static uint32_t morton3D_invert(uint32_t x) {
	x = x & 0x49249249;
	x = (x | (x >> 2)) & 0xc30c30c3;
	x = (x | (x >> 4)) & 0x0f00f00f;
	x = (x | (x >> 8)) & 0xff0000ff;
	x = (x | (x >> 16)) & 0x0000ffff;
	return x;
}
EXPORT uint32_t orc_morton3d(uint32_t x, uint32_t y, uint32_t z) { return morton3D(x, y, z); }

/* ------------------------------------------------------------------------------------------------------------------
 * stepping (nerf_device.cuh:360-495)
 * ---------------------------------------------------------------------------------------------------------------- */
EXPORT void orc_march_consts(float cone_angle, ngp_march_consts* m) {
	memset(m, 0, sizeof(*m));
	m->cone_angle = cone_angle;
	if (cone_angle <= 1e-5f) return;
	float log1p_c = ngp_logf(1.0f + cone_angle);
	m->log1p_c = log1p_c;
	m->a = (ngp_logf(MIN_CONE_STEPSIZE()) - ngp_logf(log1p_c)) / log1p_c;
	m->b = (ngp_logf(MAX_CONE_STEPSIZE()) - ngp_logf(log1p_c)) / log1p_c;
	m->at = ngp_expf(m->a * log1p_c);
	m->bt = ngp_expf(m->b * log1p_c);
}
static float to_stepping_space(float t, const ngp_march_consts* m) {
	if (m->cone_angle <= 1e-5f) return t / MIN_CONE_STEPSIZE();
	if (t <= m->at) return (t - m->at) / MIN_CONE_STEPSIZE() + m->a;
	else if (t <= m->bt) return ngp_logf(t) / m->log1p_c;
	else return (t - m->bt) / MAX_CONE_STEPSIZE() + m->b;
}
static float from_stepping_space(float n, const ngp_march_consts* m) {
	if (m->cone_angle <= 1e-5f) return n * MIN_CONE_STEPSIZE();
	if (n <= m->a) return (n - m->a) * MIN_CONE_STEPSIZE() + m->at;
	else if (n <= m->b) return ngp_expf(n * m->log1p_c);
	else return (n - m->b) * MAX_CONE_STEPSIZE() + m->bt;
}
static float advance_n_steps(float t, const ngp_march_consts* m, float n) { return from_stepping_space(to_stepping_space(t, m) + n, m); }
static float calc_dt(float t, const ngp_march_consts* m) { return advance_n_steps(t, m, 1.0f) - t; }
EXPORT float orc_to_stepping_space(float t, const ngp_march_consts* m) { return to_stepping_space(t, m); }
EXPORT float orc_from_stepping_space(float n, const ngp_march_consts* m) { return from_stepping_space(n, m); }
EXPORT float orc_calc_dt(float t, const ngp_march_consts* m) { return calc_dt(t, m); }

static float sign1(float x) { return copysignf(1.0f, x); }
static float distance_to_next_voxel(v3 pos, v3 dir, v3 idir, float res) {
	v3 p = V(res * (pos.x - 0.5f), res * (pos.y - 0.5f), res * (pos.z - 0.5f));
	float tx = (floorf(p.x + 0.5f + 0.5f * sign1(dir.x)) - p.x) * idir.x;
	float ty = (floorf(p.y + 0.5f + 0.5f * sign1(dir.y)) - p.y) * idir.y;
	float tz = (floorf(p.z + 0.5f + 0.5f * sign1(dir.z)) - p.z) * idir.z;
	float t = fminf(fminf(tx, ty), tz);
	return fmaxf(t / res, 0.0f);
}
static float advance_to_next_voxel(float t, const ngp_march_consts* m, v3 pos, v3 dir, v3 idir, uint32_t mip) {
	float res = scalbnf((float)GRIDSIZE, -(int)mip);
	float t_target = t + distance_to_next_voxel(pos, dir, idir, res);
	t = to_stepping_space(t, m);
	t_target = to_stepping_space(t_target, m);
	return from_stepping_space(t + ceilf(fmaxf(t_target - t, 0.5f)), m);
}
static int clampi(int a, int lo, int hi) { return a < lo ? lo : (hi < a ? hi : a); }
static uint32_t mip_from_pos(v3 pos, uint32_t max_cascade) {
	int exponent;
	float maxval = fmaxf(fmaxf(fabsf(pos.x - 0.5f), fabsf(pos.y - 0.5f)), fabsf(pos.z - 0.5f));
	frexpf(maxval, &exponent);
	return (uint32_t)clampi(exponent + 1, 0, (int)max_cascade);
}
static uint32_t mip_from_dt(float dt, v3 pos, uint32_t max_cascade) {
	uint32_t mip = mip_from_pos(pos, max_cascade);
	dt *= 2 * GRIDSIZE;
	if (dt < 1.0f) return mip;
	int exponent;
	frexpf(dt, &exponent);
	return (uint32_t)clampi((int)mip, exponent, (int)max_cascade);
}
static uint32_t cascaded_grid_idx_at(v3 pos, uint32_t mip) {
	float mip_scale = scalbnf(1.0f, -(int)mip);
	pos = V((pos.x - 0.5f) * mip_scale + 0.5f, (pos.y - 0.5f) * mip_scale + 0.5f, (pos.z - 0.5f) * mip_scale + 0.5f);
	int ix = (int)(pos.x * (float)GRIDSIZE), iy = (int)(pos.y * (float)GRIDSIZE), iz = (int)(pos.z * (float)GRIDSIZE);
	if (ix < 0 || ix >= GRIDSIZE || iy < 0 || iy >= GRIDSIZE || iz < 0 || iz >= GRIDSIZE) return 0xFFFFFFFFu;
	return morton3D((uint32_t)ix, (uint32_t)iy, (uint32_t)iz);
}
static int density_grid_occupied_at(v3 pos, const uint8_t* bitfield, uint32_t mip) {
	uint32_t idx = cascaded_grid_idx_at(pos, mip);
	if (idx == 0xFFFFFFFFu) return 0;
	return (bitfield[idx / 8 + (GRID_N_CELLS * mip) / 8] & (1u << (idx % 8))) != 0;
}
EXPORT uint32_t orc_mip_from_dt(float dt, const float* pos, uint32_t max_cascade) { return mip_from_dt(dt, V(pos[0], pos[1], pos[2]), max_cascade); }
EXPORT uint32_t orc_cascaded_grid_idx_at(const float* pos, uint32_t mip) { return cascaded_grid_idx_at(V(pos[0], pos[1], pos[2]), mip); }
EXPORT float orc_advance_to_next_voxel(float t, const ngp_march_consts* m, const float* pos, const float* dir, uint32_t mip) {
	v3 d = V(dir[0], dir[1], dir[2]);
	return advance_to_next_voxel(t, m, V(pos[0], pos[1], pos[2]), d, V(1.0f / d.x, 1.0f / d.y, 1.0f / d.z), mip);
}

/* ------------------------------------------------------------------------------------------------------------------
 * box, camera, images
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct { v3 mn, mx; } aabb_t;
static int aabb_contains(const aabb_t* b, v3 p) { return p.x >= b->mn.x && p.x <= b->mx.x && p.y >= b->mn.y && p.y <= b->mx.y && p.z >= b->mn.z && p.z <= b->mx.z; }
/* bounding_box.cuh:163-200 */
static void aabb_ray_intersect(const aabb_t* b, v3 pos, v3 dir, float* tmin_o, float* tmax_o) {
	const float big = 3.402823466e+38f;
	float tmin = (b->mn.x - pos.x) / dir.x, tmax = (b->mx.x - pos.x) / dir.x;
	if (tmin > tmax) { float t = tmin; tmin = tmax; tmax = t; }
	float tymin = (b->mn.y - pos.y) / dir.y, tymax = (b->mx.y - pos.y) / dir.y;
	if (tymin > tymax) { float t = tymin; tymin = tymax; tymax = t; }
	if (tmin > tymax || tymin > tmax) { *tmin_o = big; *tmax_o = big; return; }
	if (tymin > tmin) tmin = tymin;
	if (tymax < tmax) tmax = tymax;
	float tzmin = (b->mn.z - pos.z) / dir.z, tzmax = (b->mx.z - pos.z) / dir.z;
	if (tzmin > tzmax) { float t = tzmin; tzmin = tzmax; tzmax = t; }
	if (tmin > tzmax || tzmin > tmax) { *tmin_o = big; *tmax_o = big; return; }
	if (tzmin > tmin) tmin = tzmin;
	if (tzmax < tmax) tmax = tzmax;
	*tmin_o = tmin; *tmax_o = tmax;
}
static aabb_t cfg_aabb(const ngp_nerf_train_cfg* c) {
	aabb_t b = {V(c->aabb_min[0], c->aabb_min[1], c->aabb_min[2]), V(c->aabb_max[0], c->aabb_max[1], c->aabb_max[2])};
	return b;
}
static v3 warp_position(v3 p, const aabb_t* b) { return V((p.x - b->mn.x) / (b->mx.x - b->mn.x), (p.y - b->mn.y) / (b->mx.y - b->mn.y), (p.z - b->mn.z) / (b->mx.z - b->mn.z)); }
static v3 unwarp_position(v3 p, const aabb_t* b) { return V(b->mn.x + p.x * (b->mx.x - b->mn.x), b->mn.y + p.y * (b->mx.y - b->mn.y), b->mn.z + p.z * (b->mx.z - b->mn.z)); }
static v3 warp_direction(v3 d) { return V((d.x + 1.0f) * 0.5f, (d.y + 1.0f) * 0.5f, (d.z + 1.0f) * 0.5f); }
static float warp_dt(float dt) { float mx = MIN_CONE_STEPSIZE() * 128.0f; return (dt - MIN_CONE_STEPSIZE()) / (mx - MIN_CONE_STEPSIZE()); }
static float unwarp_dt(float dt) { float mx = MIN_CONE_STEPSIZE() * 128.0f; return dt * (mx - MIN_CONE_STEPSIZE()) + MIN_CONE_STEPSIZE(); }

/* common_device.cuh:268-282, 307-353 */
static void opencv_lens_distortion_delta(const float* p, float u, float v, float* du, float* dv) {
	float k1 = p[0], k2 = p[1], p1 = p[2], p2 = p[3];
	float u2 = u * u, uv = u * v, v2 = v * v, r2 = u2 + v2;
	float radial = k1 * r2 + k2 * r2 * r2;
	*du = u * radial + 2.0f * p1 * uv + p2 * (r2 + 2.0f * u2);
	*dv = v * radial + 2.0f * p2 * uv + p1 * (r2 + 2.0f * v2);
}
static void opencv_lens_undistort(const float* params, float* u, float* v) {
	float x0 = *u, y0 = *v, x = x0, y = y0;
	for (uint32_t i = 0; i < 100; ++i) {
		float step0 = fmaxf(1.1920929e-07f, fabsf(1e-6f * x));
		float step1 = fmaxf(1.1920929e-07f, fabsf(1e-6f * y));
		float dx, dy, dx0b, dy0b, dx0f, dy0f, dx1b, dy1b, dx1f, dy1f;
		opencv_lens_distortion_delta(params, x, y, &dx, &dy);
		opencv_lens_distortion_delta(params, x - step0, y, &dx0b, &dy0b);
		opencv_lens_distortion_delta(params, x + step0, y, &dx0f, &dy0f);
		opencv_lens_distortion_delta(params, x, y - step1, &dx1b, &dy1b);
		opencv_lens_distortion_delta(params, x, y + step1, &dx1f, &dy1f);
		float j00 = 1.0f + (dx0f - dx0b) / (2.0f * step0);
		float j10 = (dx1f - dx1b) / (2.0f * step1);
		float j01 = (dy0f - dy0b) / (2.0f * step0);
		float j11 = 1.0f + (dy1f - dy1b) / (2.0f * step1);
		float rx = x + dx - x0, ry = y + dy - y0;
		float det = j00 * j11 - j10 * j01;
		float sx = (j11 * rx - j10 * ry) / det;
		float sy = (-j01 * rx + j00 * ry) / det;
		x -= sx; y -= sy;
		if (sx * sx + sy * sy < 1e-10f) break;
	}
	*u = x; *v = y;
}
static v3 xform_col(const float* m, int c) { return V(m[3 * c], m[3 * c + 1], m[3 * c + 2]); }
static v3 xform_rotate(const float* m, v3 v) {
	return V((m[0] * v.x + m[3] * v.y) + m[6] * v.z, (m[1] * v.x + m[4] * v.y) + m[7] * v.z, (m[2] * v.x + m[5] * v.y) + m[8] * v.z);
}
/* uv_to_ray (common_device.cuh:413-490), perspective / OpenCV lenses, no parallax, no aperture */
static void uv_to_ray(float u, float v, int w, int h, float fx, float fy, float cx, float cy, uint32_t lens_mode, const float* lens_params,
	const float* xform, v3* o, v3* d) {
	float dx = (u - cx) * (float)w / fx;
	float dy = (v - cy) * (float)h / fy;
	if (lens_mode == NGP_LENS_OPENCV) opencv_lens_undistort(lens_params, &dx, &dy);
	*d = xform_rotate(xform, V(dx, dy, 1.0f));
	*o = xform_col(xform, 3);
}
EXPORT void orc_uv_to_ray(float u, float v, const ngp_train_view* vw, float* o6) {
	v3 o, d;
	uv_to_ray(u, v, vw->width, vw->height, vw->focal_x, vw->focal_y, vw->principal_x, vw->principal_y, vw->lens_mode, vw->lens_params, vw->xform, &o, &d);
	o6[0] = o.x; o6[1] = o.y; o6[2] = o.z; o6[3] = d.x; o6[4] = d.y; o6[5] = d.z;
}

/* colour (common_device.cuh:61-103) */
static float srgb_to_linear(float s) { return s <= 0.04045f ? s / 12.92f : ngp_powf((s + 0.055f) / 1.055f, 2.4f); }
static float srgb_to_linear_derivative(float s) { return s <= 0.04045f ? 1.0f / 12.92f : 2.4f / 1.055f * ngp_powf((s + 0.055f) / 1.055f, 1.4f); }   /* common_device.cuh:71-77 */
static float linear_to_srgb(float l) { return l < 0.0031308f ? 12.92f * l : 1.055f * ngp_powf(l, 0.41666f) - 0.055f; }
/* ngp_detmath.h itself, element-wise: the header is shared by this oracle and the CUDA product, so it is checked against libm on its own
 * (tests/test_detmath.py) — a bug in it would be invisible to every CUDA-vs-oracle comparison.  which: 0 log, 1 exp, 2 pow(x, y) */
EXPORT void orc_detmath_n(int which, const float* x, const float* y, float* out, uint32_t n) {
	for (uint32_t i = 0; i < n; ++i) out[i] = which == 0 ? ngp_logf(x[i]) : (which == 1 ? ngp_expf(x[i]) : ngp_powf(x[i], y[i]));
}
EXPORT float orc_srgb_to_linear(float s) { return srgb_to_linear(s); }
EXPORT float orc_linear_to_srgb(float l) { return linear_to_srgb(l); }
EXPORT void orc_linear_to_srgb_n(const float* in, float* out, uint32_t n) {
	for (uint32_t i = 0; i < n; ++i) out[i] = linear_to_srgb(in[i]);
}
EXPORT void orc_srgb_to_linear_n(const float* in, float* out, uint32_t n) {
	for (uint32_t i = 0; i < n; ++i) out[i] = srgb_to_linear(in[i]);
}

/* fp16 <-> fp32 (IEEE binary16, round-to-nearest-even) */
static float half_to_float(uint16_t h) {
	uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1F, man = h & 0x3FF;
	union { uint32_t u; float f; } o;
	if (exp == 0) {
		if (man == 0) { o.u = sign; return o.f; }
		float f = ldexpf((float)man, -24);
		return sign ? -f : f;
	}
	if (exp == 31) { o.u = sign | 0x7F800000u | (man << 13); return o.f; }
	o.u = sign | ((exp + 112) << 23) | (man << 13);
	return o.f;
}
static uint16_t float_to_half(float f) {
	union { float f; uint32_t u; } in; in.f = f;
	uint32_t sign = (in.u >> 16) & 0x8000u;
	uint32_t x = in.u & 0x7FFFFFFFu;
	if (x >= 0x7F800000u) return (uint16_t)(sign | (x > 0x7F800000u ? 0x7E00u : 0x7C00u));
	if (x >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u); /* >= 65520 rounds to inf */
	if (x < 0x33000001u) return (uint16_t)sign;              /* < 2^-25 (or exactly 2^-25, ties to even 0) */
	int e = (int)(x >> 23) - 127;
	uint32_t man = (x & 0x7FFFFFu) | 0x800000u;
	int shift = (e < -14) ? (13 + (-14 - e)) : 13;
	uint32_t half_man = man >> shift;
	uint32_t rem = man & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
	if (rem > halfway || (rem == halfway && (half_man & 1u))) ++half_man;
	uint32_t result = (e < -14) ? half_man : (((uint32_t)(e + 15) << 10) + (half_man - 0x400u));
	return (uint16_t)(sign | result);
}
EXPORT float orc_half_to_float(uint16_t h) { return half_to_float(h); }
EXPORT uint16_t orc_float_to_half(float f) { return float_to_half(f); }

typedef struct { float r, g, b, a; } rgba_t;
/* read_rgba (common_device.cuh:846-872) */
static rgba_t read_rgba_px(int px, int py, int w, const void* pixels, uint32_t type) {
	size_t idx = (size_t)px + (size_t)py * (size_t)w;
	rgba_t o;
	switch (type) {
		case NGP_IMAGE_BYTE: {
			uint32_t val = ((const uint32_t*)pixels)[idx];
			if (val == 0x00FF00FFu) { o.r = o.g = o.b = o.a = -1.0f; return o; }
			float a = (float)((val >> 24) & 0xFFu) * (1.0f / 255.0f);
			o.r = srgb_to_linear((float)(val & 0xFFu) * (1.0f / 255.0f)) * a;
			o.g = srgb_to_linear((float)((val >> 8) & 0xFFu) * (1.0f / 255.0f)) * a;
			o.b = srgb_to_linear((float)((val >> 16) & 0xFFu) * (1.0f / 255.0f)) * a;
			o.a = a;
			return o;
		}
		case NGP_IMAGE_HALF: {
			const uint16_t* p = (const uint16_t*)pixels + idx * 4;
			o.r = half_to_float(p[0]); o.g = half_to_float(p[1]); o.b = half_to_float(p[2]); o.a = half_to_float(p[3]);
			return o;
		}
		case NGP_IMAGE_FLOAT: {
			const float* p = (const float*)pixels + idx * 4;
			o.r = p[0]; o.g = p[1]; o.b = p[2]; o.a = p[3];
			return o;
		}
		default: o.r = 5.0f; o.g = 0.0f; o.b = 0.0f; o.a = 1.0f; return o;
	}
}
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }
static rgba_t read_rgba_uv(float u, float v, int w, int h, const void* pixels, uint32_t type) {
	int px = imin(imax((int)(u * (float)w), 0), w - 1), py = imin(imax((int)(v * (float)h), 0), h - 1);
	return read_rgba_px(px, py, w, pixels, type);
}
/* nerf_device.cuh:578-599 uniform branch; :553-576 */
static uint32_t image_idx(uint32_t base_idx, uint32_t n_rays, uint32_t n_images) { return ((base_idx * n_images) / n_rays) % n_images; }
static void random_image_pos_training(pcg32_t* rng, int w, int h, int snap, float* u, float* v) {
	*u = pcg_next_float(rng);
	*v = pcg_next_float(rng);
	if (snap) {
		*u = ((float)imin(imax((int)(*u * (float)w), 0), w - 1) + 0.5f) / (float)w;
		*v = ((float)imin(imax((int)(*v * (float)h), 0), h - 1) + 0.5f) / (float)h;
	}
}

/* ------------------------------------------------------------------------------------------------------------------
 * generate_training_samples_nerf (src/testbed_nerf.cu:691-849).
 * Rays are processed in index order; slots are handed out in that order (the reference's order depends on atomics).
 * per_ray_numsteps[n_rays] receives the step count of EVERY ray (0 when it contributes nothing) so tests can compare the
 * map ray id -> count.  Returns the number of rays kept.
 * ---------------------------------------------------------------------------------------------------------------- */
EXPORT uint32_t orc_generate_training_samples(uint32_t n_rays_local, uint32_t ray_offset, uint32_t n_rays_global, uint64_t rng_state, uint64_t rng_inc,
	const ngp_nerf_train_cfg* cfg, const ngp_train_view* views, uint32_t n_views, const uint8_t* bitfield, uint32_t max_samples,
	uint32_t* n_samples_out, uint32_t* per_ray_numsteps, uint32_t* ray_indices_out, float* rays_out, uint32_t* numsteps_out, float* coords_out) {
	const aabb_t aabb = cfg_aabb(cfg);
	uint32_t ray_counter = 0, numsteps_counter = 0;
	const uint32_t ray_stride = cfg->ray_stride ? cfg->ray_stride : 1u;   /* data parallel: rank r of W marches ids r, r + W, ... */
	for (uint32_t li = 0; li < n_rays_local; ++li) {
		const uint32_t i = ray_offset + li * ray_stride;
		if (per_ray_numsteps) per_ray_numsteps[li] = 0;
		uint32_t img = image_idx(i, n_rays_global, n_views);
		const ngp_train_view* vw = &views[img];
		pcg32_t rng = {rng_state, rng_inc};
		pcg_advance(&rng, (uint64_t)i * N_MAX_RANDOM_SAMPLES_PER_RAY);
		float u, v;
		random_image_pos_training(&rng, vw->width, vw->height, cfg->snap_to_pixel_centers != 0, &u, &v);
		if (read_rgba_uv(u, v, vw->width, vw->height, vw->pixels, vw->image_type).r < 0.0f) continue;
		(void)pcg_next_float(&rng); /* motionblur_time */
		v3 ro, rd;
		uv_to_ray(u, v, vw->width, vw->height, vw->focal_x, vw->focal_y, vw->principal_x, vw->principal_y, vw->lens_mode, vw->lens_params, vw->xform, &ro, &rd);
		v3 rdn = vnormalize(rd);
		float tmin, tmax;
		aabb_ray_intersect(&aabb, ro, rdn, &tmin, &tmax);
		tmin = fmaxf(tmin, 0.0f);
		float startt = advance_n_steps(tmin, &cfg->march, pcg_next_float(&rng));
		v3 idir = V(1.0f / rdn.x, 1.0f / rdn.y, 1.0f / rdn.z);

		uint32_t j = 0;
		float t = startt;
		v3 pos;
		while (aabb_contains(&aabb, pos = vadd(ro, vscale(rdn, t))) && j < NERF_STEPS) {
			float dt = calc_dt(t, &cfg->march);
			uint32_t mip = mip_from_dt(dt, pos, cfg->max_cascade);
			if (density_grid_occupied_at(pos, bitfield, mip)) { ++j; t += dt; }
			else t = advance_to_next_voxel(t, &cfg->march, pos, rdn, idir, mip);
		}
		if (j == 0) continue;
		uint32_t numsteps = j;
		if (per_ray_numsteps) per_ray_numsteps[li] = numsteps;
		uint32_t base = numsteps_counter;
		numsteps_counter += numsteps;
		if (base + numsteps > max_samples) continue;
		uint32_t ray_idx = ray_counter++;
		if (ray_indices_out) ray_indices_out[ray_idx] = i;
		if (rays_out) { float* r = rays_out + (size_t)ray_idx * 6; r[0] = ro.x; r[1] = ro.y; r[2] = ro.z; r[3] = rd.x; r[4] = rd.y; r[5] = rd.z; }
		if (numsteps_out) { numsteps_out[ray_idx * 2] = numsteps; numsteps_out[ray_idx * 2 + 1] = base; }
		if (coords_out) {
			v3 wdir = warp_direction(rdn);
			float* co = coords_out + (size_t)base * 7;
			t = startt; j = 0;
			while (aabb_contains(&aabb, pos = vadd(ro, vscale(rdn, t))) && j < numsteps) {
				float dt = calc_dt(t, &cfg->march);
				uint32_t mip = mip_from_dt(dt, pos, cfg->max_cascade);
				if (density_grid_occupied_at(pos, bitfield, mip)) {
					v3 wp = warp_position(pos, &aabb);
					float* c = co + (size_t)j * 7;
					c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = warp_dt(dt); c[4] = wdir.x; c[5] = wdir.y; c[6] = wdir.z;
					++j; t += dt;
				} else t = advance_to_next_voxel(t, &cfg->march, pos, rdn, idir, mip);
			}
		}
	}
	if (n_samples_out) *n_samples_out = numsteps_counter;
	return ray_counter;
}

/* ------------------------------------------------------------------------------------------------------------------
 * activations and losses (nerf_device.cuh:75-143, 204-264, 601-616)
 * ---------------------------------------------------------------------------------------------------------------- */
static float clampf(float a, float lo, float hi) { return a < lo ? lo : (hi < a ? hi : a); }
static float logistic(float x) { return 1.0f / (1.0f + ngp_expf(-x)); }
static float network_to_rgb(float v, uint32_t act) {
	switch (act) { case NGP_ACT_NONE: return v; case NGP_ACT_RELU: return v > 0.0f ? v : 0.0f; case NGP_ACT_LOGISTIC: return logistic(v);
		default: return ngp_expf(clampf(v, -10.0f, 10.0f)); }
}
static float network_to_rgb_derivative(float v, uint32_t act) {
	switch (act) { case NGP_ACT_NONE: return 1.0f; case NGP_ACT_RELU: return v > 0.0f ? 1.0f : 0.0f;
		case NGP_ACT_LOGISTIC: { float d = logistic(v); return d * (1.0f - d); } default: return ngp_expf(clampf(v, -10.0f, 10.0f)); }
}
static float network_to_density(float v, uint32_t act) {
	switch (act) { case NGP_ACT_NONE: return v; case NGP_ACT_RELU: return v > 0.0f ? v : 0.0f; case NGP_ACT_LOGISTIC: return logistic(v); default: return ngp_expf(v); }
}
static float network_to_density_derivative(float v, uint32_t act) {
	switch (act) { case NGP_ACT_NONE: return 1.0f; case NGP_ACT_RELU: return v > 0.0f ? 1.0f : 0.0f;
		case NGP_ACT_LOGISTIC: { float d = logistic(v); return d * (1.0f - d); } default: return ngp_expf(clampf(v, -15.0f, 15.0f)); }
}
static void loss_and_gradient1(float target, float pred, uint32_t type, float* loss, float* grad) {
	float d = pred - target;
	switch (type) {
		case NGP_LOSS_RELATIVE_L2: { float den = pred * pred + 1e-2f; *loss = d * d / den; *grad = 2.0f * d / den; break; }
		case NGP_LOSS_L1: *loss = fabsf(d); *grad = copysignf(1.0f, d); break;
		case NGP_LOSS_MAPE: { float den = fabsf(pred) + 1e-2f; *loss = fabsf(d) / den; *grad = copysignf(1.0f / den, d); break; }
		case NGP_LOSS_SMAPE: { float den = 0.5f * (fabsf(pred) + fabsf(target)) + 1e-2f; *loss = fabsf(d) / den; *grad = copysignf(1.0f / den, d); break; }
		case NGP_LOSS_HUBER: {
			float alpha = 0.1f, ad = fabsf(d), sq = 0.5f / alpha * d * d;
			*loss = (ad > alpha ? (ad - 0.5f * alpha) : sq) / 5.0f;
			*grad = (ad > alpha ? (d > 0.0f ? 1.0f : -1.0f) : (d / alpha)) / 5.0f;
			break;
		}
		case NGP_LOSS_LOGL1: { float div = fabsf(d) + 1.0f; *loss = ngp_logf(div); *grad = copysignf(1.0f / div, d); break; }
		default: *loss = d * d; *grad = 2.0f * d; break;
	}
}

/* ------------------------------------------------------------------------------------------------------------------
 * compute_loss_kernel_train_nerf (src/testbed_nerf.cu:852-1180) — no envmap / depth / exposure / error map; train modes Rfl / RflRelax with the
 * gradient forms and the compositing arithmetic of the fused kernel they run through in the reference (fused_kernels/train_nerf.cuh:176-420).
 * network_output: n_samples x 4 binary16 values (rgb raw x3, density raw).  Rays are visited in slot order, compacted
 * slots are handed out in that order.  Returns the (unclamped) compacted sample count.
 * ---------------------------------------------------------------------------------------------------------------- */
EXPORT uint32_t orc_compute_loss(uint32_t n_rays_kept, uint32_t n_rays_global, uint64_t rng_state, uint64_t rng_inc, const ngp_nerf_train_cfg* cfg,
	const ngp_train_view* views, uint32_t n_views, const uint16_t* network_output, uint32_t max_compacted, const uint32_t* ray_indices_in,
	const float* rays_in, uint32_t* numsteps_inout, const float* coords_in, float* coords_out, uint16_t* dloss_out, float* loss_output,
	float mean_density) {
	const aabb_t aabb = cfg_aabb(cfg);
	uint32_t compacted_counter = 0;
	for (uint32_t i = 0; i < n_rays_kept; ++i) {
		uint32_t numsteps = numsteps_inout[i * 2], base = numsteps_inout[i * 2 + 1];
		const float* ci = coords_in + (size_t)base * 7;
		const uint16_t* no = network_output + (size_t)base * 4;
		float T = 1.f;
		const float EPSILON = 1e-4f;
		/* Rfl / RflRelax: the fused train_nerf kernel's compositing (fused_kernels/train_nerf.cuh:228-238, 251, 305, 363-367):
		 * transmittance = 1 - accumulated weight, background unless opaque, no L1 term on the density */
		const int rtc = cfg->train_mode != NGP_TRAIN_NERF;
		float acc = 0.f;
		v3 rgb_ray = V(0, 0, 0);
		uint32_t compacted_numsteps = 0;
		v3 ray_o = V(rays_in[(size_t)i * 6], rays_in[(size_t)i * 6 + 1], rays_in[(size_t)i * 6 + 2]);
		for (; compacted_numsteps < numsteps; ++compacted_numsteps) {
			if (T < EPSILON) break;
			const uint16_t* o = no + (size_t)compacted_numsteps * 4;
			v3 rgb = V(network_to_rgb(half_to_float(o[0]), cfg->rgb_activation), network_to_rgb(half_to_float(o[1]), cfg->rgb_activation),
				network_to_rgb(half_to_float(o[2]), cfg->rgb_activation));
			float dt = unwarp_dt(ci[(size_t)compacted_numsteps * 7 + 3]);
			float density = network_to_density(half_to_float(o[3]), cfg->density_activation);
			float alpha = 1.f - ngp_expf(-density * dt);
			float weight = alpha * T;
			rgb_ray = vadd(rgb_ray, vscale(rgb, weight));
			if (rtc) { acc += weight; T = 1.f - acc; } else T *= (1.f - alpha);
		}
		uint32_t ray_idx = ray_indices_in[i];
		pcg32_t rng = {rng_state, rng_inc};
		pcg_advance(&rng, (uint64_t)ray_idx * N_MAX_RANDOM_SAMPLES_PER_RAY);
		uint32_t img = image_idx(ray_idx, n_rays_global, n_views);
		const ngp_train_view* vw = &views[img];
		float u, v;
		random_image_pos_training(&rng, vw->width, vw->height, cfg->snap_to_pixel_centers != 0, &u, &v);
		pcg_advance(&rng, 1); /* motionblur_time */
		v3 bg = V(cfg->background_color[0], cfg->background_color[1], cfg->background_color[2]);
		if (cfg->random_bg_color) { bg.x = pcg_next_float(&rng); bg.y = pcg_next_float(&rng); bg.z = pcg_next_float(&rng); }
		bg = V(srgb_to_linear(bg.x), srgb_to_linear(bg.y), srgb_to_linear(bg.z));
		rgba_t tex = read_rgba_uv(u, v, vw->width, vw->height, vw->pixels, vw->image_type);
		/* per-image exposure (testbed_nerf.cu:979): the view's colour times 2^exposure, channel by channel */
		v3 exposure_scale = V(1.0f, 1.0f, 1.0f);
		if (cfg->cam_exposure) {
			exposure_scale = V(ngp_expf(0.6931471805599453f * cfg->cam_exposure[img * 3 + 0]), ngp_expf(0.6931471805599453f * cfg->cam_exposure[img * 3 + 1]),
				ngp_expf(0.6931471805599453f * cfg->cam_exposure[img * 3 + 2]));
			tex.r = exposure_scale.x * tex.r;
			tex.g = exposure_scale.y * tex.g;
			tex.b = exposure_scale.z * tex.b;
		}
		v3 target;
		if (cfg->linear_colors || cfg->color_space == NGP_COLOR_LINEAR) {
			target = V(tex.r + (1.0f - tex.a) * bg.x, tex.g + (1.0f - tex.a) * bg.y, tex.b + (1.0f - tex.a) * bg.z);
			if (!cfg->linear_colors) {
				target = V(linear_to_srgb(target.x), linear_to_srgb(target.y), linear_to_srgb(target.z));
				bg = V(linear_to_srgb(bg.x), linear_to_srgb(bg.y), linear_to_srgb(bg.z));
			}
		} else {
			bg = V(linear_to_srgb(bg.x), linear_to_srgb(bg.y), linear_to_srgb(bg.z));
			if (tex.a > 0) {
				target = V(linear_to_srgb(tex.r / tex.a) * tex.a + (1.0f - tex.a) * bg.x, linear_to_srgb(tex.g / tex.a) * tex.a + (1.0f - tex.a) * bg.y,
					linear_to_srgb(tex.b / tex.a) * tex.a + (1.0f - tex.a) * bg.z);
			} else target = bg;
		}
		const float T_end = T;
		const int background_shows = rtc ? !(T < EPSILON) : (compacted_numsteps == numsteps);
		if (background_shows) rgb_ray = vadd(rgb_ray, vscale(bg, T));
		/* Rfl (train_nerf.cuh:231, 251-254): the ray's accumulated per-sample loss, background term included */
		v3 loss_bg = V(0, 0, 0);
		if (cfg->train_mode == NGP_TRAIN_RFL) {
			float Tq = 1.f, accq = 0.f;
			for (uint32_t q = 0; q < compacted_numsteps; ++q) {
				const uint16_t* o = no + (size_t)q * 4;
				v3 rgb = V(network_to_rgb(half_to_float(o[0]), cfg->rgb_activation), network_to_rgb(half_to_float(o[1]), cfg->rgb_activation),
					network_to_rgb(half_to_float(o[2]), cfg->rgb_activation));
				float dt = unwarp_dt(ci[(size_t)q * 7 + 3]);
				float alpha = 1.f - ngp_expf(-network_to_density(half_to_float(o[3]), cfg->density_activation) * dt);
				float weight = alpha * Tq;
				accq += weight;
				Tq = 1.f - accq;   /* Rfl is an rtc mode */
				float l0, l1, l2, gdummy;
				loss_and_gradient1(target.x, rgb.x, cfg->loss_type, &l0, &gdummy);
				loss_and_gradient1(target.y, rgb.y, cfg->loss_type, &l1, &gdummy);
				loss_and_gradient1(target.z, rgb.z, cfg->loss_type, &l2, &gdummy);
				loss_bg = vadd(loss_bg, vscale(V(l0, l1, l2), weight));
			}
			if (background_shows) {
				float l0, l1, l2, gdummy;
				loss_and_gradient1(target.x, bg.x, cfg->loss_type, &l0, &gdummy);
				loss_and_gradient1(target.y, bg.y, cfg->loss_type, &l1, &gdummy);
				loss_and_gradient1(target.z, bg.z, cfg->loss_type, &l2, &gdummy);
				loss_bg = vadd(loss_bg, vscale(V(l0, l1, l2), T_end));
			}
		}

		uint32_t compacted_base = compacted_counter;
		compacted_counter += compacted_numsteps;
		uint32_t cb = compacted_base < max_compacted ? compacted_base : max_compacted;
		uint32_t room = max_compacted - cb;
		compacted_numsteps = room < compacted_numsteps ? room : compacted_numsteps;
		numsteps_inout[i * 2] = compacted_numsteps;
		numsteps_inout[i * 2 + 1] = compacted_base;
		if (compacted_numsteps == 0) continue;

		float lx, ly, lz; v3 g;
		loss_and_gradient1(target.x, rgb_ray.x, cfg->loss_type, &lx, &g.x);
		loss_and_gradient1(target.y, rgb_ray.y, cfg->loss_type, &ly, &g.y);
		loss_and_gradient1(target.z, rgb_ray.z, cfg->loss_type, &lz, &g.z);
		float mean_loss = ((lx + ly) + lz) / 3.0f;
		if (loss_output) loss_output[i] = mean_loss / (float)n_rays_global;

		float loss_scale = cfg->loss_scale / (float)n_rays_global;
		if (cfg->cam_exposure_gradient) {   /* testbed_nerf.cu:1142-1155 */
			v3 dgt = V(-g.x, -g.y, -g.z);
			if (!cfg->linear_colors) dgt = V(dgt.x / srgb_to_linear_derivative(target.x), dgt.y / srgb_to_linear_derivative(target.y), dgt.z / srgb_to_linear_derivative(target.z));
			float* eg = cfg->cam_exposure_gradient + (size_t)img * 3;
			eg[0] += ((loss_scale * dgt.x) * exposure_scale.x) * 0.6931471805599453f;
			eg[1] += ((loss_scale * dgt.y) * exposure_scale.y) * 0.6931471805599453f;
			eg[2] += ((loss_scale * dgt.z) * exposure_scale.z) * 0.6931471805599453f;
		}
		const float output_l2_reg = cfg->rgb_activation == NGP_ACT_EXPONENTIAL ? 1e-4f : 0.0f;
		const float output_l1_reg_density = (!rtc && mean_density < MIN_OPTICAL_THICKNESS) ? 1e-4f : 0.0f;
		float* co = coords_out + (size_t)compacted_base * 7;
		uint16_t* dl = dloss_out + (size_t)compacted_base * 4;
		v3 rgb_ray2 = V(0, 0, 0);
		v3 loss_bg2 = V(0, 0, 0);
		T = 1.f;
		acc = 0.f;
		for (uint32_t j = 0; j < compacted_numsteps; ++j) {
			const float* c = ci + (size_t)j * 7;
			for (int k = 0; k < 7; ++k) co[(size_t)j * 7 + k] = c[k];
			v3 pos = unwarp_position(V(c[0], c[1], c[2]), &aabb);
			float depth = vlen(vsub(pos, ray_o));
			float dt = unwarp_dt(c[3]);
			const uint16_t* o = no + (size_t)j * 4;
			float o0 = half_to_float(o[0]), o1 = half_to_float(o[1]), o2 = half_to_float(o[2]), o3 = half_to_float(o[3]);
			v3 rgb = V(network_to_rgb(o0, cfg->rgb_activation), network_to_rgb(o1, cfg->rgb_activation), network_to_rgb(o2, cfg->rgb_activation));
			float density = network_to_density(o3, cfg->density_activation);
			float alpha = 1.f - ngp_expf(-density * dt);
			float weight = alpha * T;
			rgb_ray2 = vadd(rgb_ray2, vscale(rgb, weight));
			if (rtc) { acc += weight; T = 1.f - acc; } else T *= (1.f - alpha);
			v3 suffix = vsub(rgb_ray, rgb_ray2);
			v3 dloss_by_drgb = vscale(g, weight);
			float dmlp_inner = 0.0f;   /* the bracket that multiplies density_derivative * dt (train_nerf.cuh:391-410) */
			if (cfg->train_mode == NGP_TRAIN_RFL) {
				v3 ll, lgr;
				loss_and_gradient1(target.x, rgb.x, cfg->loss_type, &ll.x, &lgr.x);
				loss_and_gradient1(target.y, rgb.y, cfg->loss_type, &ll.y, &lgr.y);
				loss_and_gradient1(target.z, rgb.z, cfg->loss_type, &ll.z, &lgr.z);
				loss_bg2 = vadd(loss_bg2, vscale(ll, weight));
				dloss_by_drgb = vscale(lgr, weight);
				v3 e = vsub(vscale(ll, T), vsub(loss_bg, loss_bg2));
				dmlp_inner = (e.x + e.y) + e.z;
			} else if (cfg->train_mode == NGP_TRAIN_RFL_RELAX) {
				const float tden = fmaxf(1e-6f, T);
				v3 rgb_bg = V(suffix.x / tden, suffix.y / tden, suffix.z / tden);
				v3 rgb_lerp = vadd(vscale(rgb_bg, 1.0f - alpha), vscale(rgb, alpha));
				v3 ll, lgr;
				loss_and_gradient1(target.x, rgb_lerp.x, cfg->loss_type, &ll.x, &lgr.x);
				loss_and_gradient1(target.y, rgb_lerp.y, cfg->loss_type, &ll.y, &lgr.y);
				loss_and_gradient1(target.z, rgb_lerp.z, cfg->loss_type, &ll.z, &lgr.z);
				dloss_by_drgb = vscale(lgr, weight);
				dmlp_inner = vdot(lgr, vsub(vscale(rgb, T), suffix));
			}
			float d0 = loss_scale * (dloss_by_drgb.x * network_to_rgb_derivative(o0, cfg->rgb_activation) + fmaxf(0.0f, output_l2_reg * o0));
			float d1 = loss_scale * (dloss_by_drgb.y * network_to_rgb_derivative(o1, cfg->rgb_activation) + fmaxf(0.0f, output_l2_reg * o1));
			float d2 = loss_scale * (dloss_by_drgb.z * network_to_rgb_derivative(o2, cfg->rgb_activation) + fmaxf(0.0f, output_l2_reg * o2));
			float density_derivative = network_to_density_derivative(o3, cfg->density_activation);
			v3 tr = vsub(vscale(rgb, T), suffix);
			if (cfg->train_mode == NGP_TRAIN_NERF) dmlp_inner = vdot(g, tr);
			float dloss_by_dmlp = density_derivative * (dt * dmlp_inner);
			float d3 = loss_scale * dloss_by_dmlp + (o3 < 0.0f ? -output_l1_reg_density : 0.0f) + (o3 > -10.0f && depth < cfg->near_distance ? 1e-4f : 0.0f);
			dl[(size_t)j * 4 + 0] = float_to_half(d0);
			dl[(size_t)j * 4 + 1] = float_to_half(d1);
			dl[(size_t)j * 4 + 2] = float_to_half(d2);
			dl[(size_t)j * 4 + 3] = float_to_half(d3);
		}
	}
	return compacted_counter;
}

/* fill_rollover_and_rescale / fill_rollover (tiny-cuda-nn/common_device.h:1114-1135; testbed_nerf.cu:3298-3303) */
EXPORT void orc_fill_rollover(uint32_t n_elements, uint32_t n_compacted, float* coords, uint16_t* dloss) {
	uint32_t n_input = n_compacted < n_elements ? n_compacted : n_elements;
	if (n_input == 0) return;
	for (uint32_t i = n_input; i < n_elements; ++i) {
		uint32_t src = i % n_input;
		for (int k = 0; k < 7; ++k) coords[(size_t)i * 7 + k] = coords[(size_t)src * 7 + k];
		for (int k = 0; k < 4; ++k) dloss[(size_t)i * 4 + k] = float_to_half(half_to_float(dloss[(size_t)src * 4 + k]) * (float)n_input / (float)n_elements);
	}
}

/* ------------------------------------------------------------------------------------------------------------------
 * occupancy grid maintenance (src/testbed_nerf.cu:87-162, 216-284, 316-396, 2594-2633)
 * ---------------------------------------------------------------------------------------------------------------- */
static void pos_to_uv(const float* xform, v3 pos, const ngp_train_view* vw, float* u, float* v) {
	v3 origin = xform_col(xform, 3);
	v3 dir = vsub(pos, origin);
	float a = xform[0], b = xform[3], c = xform[6], d = xform[1], e = xform[4], f = xform[7], g = xform[2], h = xform[5], k = xform[8];
	float A = e * k - f * h, B = -(d * k - f * g), C = d * h - e * g;
	float det = a * A + b * B + c * C;
	float inv00 = A / det, inv01 = -(b * k - c * h) / det, inv02 = (b * f - c * e) / det;
	float inv10 = B / det, inv11 = (a * k - c * g) / det, inv12 = -(a * f - c * d) / det;
	float inv20 = C / det, inv21 = -(a * h - b * g) / det, inv22 = (a * e - b * d) / det;
	v3 l = V((inv00 * dir.x + inv01 * dir.y) + inv02 * dir.z, (inv10 * dir.x + inv11 * dir.y) + inv12 * dir.z, (inv20 * dir.x + inv21 * dir.y) + inv22 * dir.z);
	l = V(l.x / l.z, l.y / l.z, 1.0f);
	float du = 0.0f, dv = 0.0f;
	if (vw->lens_mode == NGP_LENS_OPENCV) opencv_lens_distortion_delta(vw->lens_params, l.x, l.y, &du, &dv);
	l.x += du; l.y += dv;
	*u = l.x * vw->focal_x / (float)vw->width + vw->principal_x;
	*v = l.y * vw->focal_y / (float)vw->height + vw->principal_y;
}
EXPORT void orc_mark_untrained_density_grid(uint32_t n_elements, float* grid_out, uint32_t n_views, const ngp_train_view* views, int clear_visible_voxels) {
#pragma omp parallel for schedule(dynamic, 4096)
	for (uint32_t i = 0; i < n_elements; ++i) {
		uint32_t level = i / GRID_N_CELLS, pos_idx = i % GRID_N_CELLS;
		uint32_t x = morton3D_invert(pos_idx >> 0), y = morton3D_invert(pos_idx >> 1), z = morton3D_invert(pos_idx >> 2);
		float voxel_size = scalbnf(1.0f / GRIDSIZE, (int)level);
		float s = scalbnf(1.0f, (int)level);
		v3 pos = V(((float)x / (float)GRIDSIZE - 0.5f) * s + 0.5f, ((float)y / (float)GRIDSIZE - 0.5f) * s + 0.5f, ((float)z / (float)GRIDSIZE - 0.5f) * s + 0.5f);
		uint32_t count = 0;
		for (uint32_t j = 0; j < n_views && count < 1; ++j) {
			const ngp_train_view* vw = &views[j];
			v3 cam_o = xform_col(vw->xform, 3), cam_fwd = xform_col(vw->xform, 2);
			for (uint32_t k = 0; k < 8; ++k) {
				v3 corner = V(pos.x + ((k & 1u) ? voxel_size : 0.0f), pos.y + ((k & 2u) ? voxel_size : 0.0f), pos.z + ((k & 4u) ? voxel_size : 0.0f));
				v3 dir = vnormalize(vsub(corner, cam_o));
				if (vdot(dir, cam_fwd) < 1e-4f) continue;
				float u, v;
				pos_to_uv(vw->xform, corner, vw, &u, &v);
				v3 ro, rd;
				uv_to_ray(u, v, vw->width, vw->height, vw->focal_x, vw->focal_y, vw->principal_x, vw->principal_y, vw->lens_mode, vw->lens_params, vw->xform, &ro, &rd);
				if (vlen(vsub(vnormalize(rd), dir)) < 1e-3f && u > 0.0f && v > 0.0f && u < 1.0f && v < 1.0f) { ++count; break; }
			}
		}
		if (clear_visible_voxels || (grid_out[i] < 0) != (count < 1)) grid_out[i] = (count >= 1) ? 0.f : -1.f;
	}
}
EXPORT void orc_generate_grid_samples(uint32_t n_elements, uint64_t rng_state, uint64_t rng_inc, uint32_t step, const ngp_nerf_train_cfg* cfg,
	const float* grid_in, float* out_pos4, uint32_t* indices, uint32_t n_cascades, float thresh) {
	const aabb_t aabb = cfg_aabb(cfg);
#pragma omp parallel for
	for (uint32_t i = 0; i < n_elements; ++i) {
		pcg32_t rng = {rng_state, rng_inc};
		pcg_advance(&rng, (uint64_t)i * 4);
		uint32_t level = (uint32_t)(pcg_next_float(&rng) * (float)n_cascades) % n_cascades;
		uint32_t idx = 0;
		for (uint32_t j = 0; j < 10; ++j) {
			idx = ((i + step * n_elements) * 56924617u + j * 19349663u + 96925573u) % GRID_N_CELLS;
			idx += level * GRID_N_CELLS;
			if (grid_in[idx] > thresh) break;
		}
		uint32_t pos_idx = idx % GRID_N_CELLS;
		uint32_t x = morton3D_invert(pos_idx >> 0), y = morton3D_invert(pos_idx >> 1), z = morton3D_invert(pos_idx >> 2);
		float rx = pcg_next_float(&rng), ry = pcg_next_float(&rng), rz = pcg_next_float(&rng);
		float s = scalbnf(1.0f, (int)level);
		v3 pos = V((((float)x + rx) / (float)GRIDSIZE - 0.5f) * s + 0.5f, (((float)y + ry) / (float)GRIDSIZE - 0.5f) * s + 0.5f, (((float)z + rz) / (float)GRIDSIZE - 0.5f) * s + 0.5f);
		v3 wp = warp_position(pos, &aabb);
		out_pos4[(size_t)i * 4] = wp.x; out_pos4[(size_t)i * 4 + 1] = wp.y; out_pos4[(size_t)i * 4 + 2] = wp.z; out_pos4[(size_t)i * 4 + 3] = warp_dt(MIN_CONE_STEPSIZE());
		indices[i] = idx;
	}
}
EXPORT void orc_splat_and_ema(uint32_t n_samples, const uint32_t* indices, const uint16_t* density_raw, uint32_t density_activation, uint32_t n_elements,
	float decay, float* grid_tmp, float* grid_inout) {
	memset(grid_tmp, 0, sizeof(float) * n_elements);
	for (uint32_t i = 0; i < n_samples; ++i) {
		float mlp = network_to_density(half_to_float(density_raw[i]), density_activation);
		float optical_thickness = mlp * MIN_CONE_STEPSIZE();
		union { float f; uint32_t u; } a, b;
		a.f = optical_thickness; b.f = grid_tmp[indices[i]];
		if (a.u > b.u) grid_tmp[indices[i]] = optical_thickness; /* atomicMax on the uint bit pattern */
	}
	for (uint32_t i = 0; i < n_elements; ++i) {
		float prev = grid_inout[i];
		grid_inout[i] = (prev < 0.f) ? prev : fmaxf(prev * decay, grid_tmp[i]);
	}
}
EXPORT float orc_density_mean(const float* grid) {
	/* reduce_sum of max(v,0)/n over the first cascade; same fixed summation order as the device path (1024 strided partials) */
	float partial[1024];
	for (uint32_t t = 0; t < 1024; ++t) {
		float s = 0.0f;
		for (uint32_t i = t; i < GRID_N_CELLS; i += 1024) s += fmaxf(grid[i], 0.0f) / (float)GRID_N_CELLS;
		partial[t] = s;
	}
	float s = 0.0f;
	for (uint32_t i = 0; i < 1024; ++i) s += partial[i];
	return s;
}
EXPORT void orc_update_bitfield(uint32_t max_cascade, const float* grid, float mean_density, uint8_t* bitfield) {
	uint32_t n_elements = GRID_N_CELLS / 8 * N_CASCADES, n_nonzero = GRID_N_CELLS / 8 * (max_cascade + 1);
	float thresh = fminf(MIN_OPTICAL_THICKNESS, mean_density);
	for (uint32_t i = 0; i < n_elements; ++i) {
		if (i >= n_nonzero) { bitfield[i] = 0; continue; }
		uint8_t bits = 0;
		for (uint8_t j = 0; j < 8; ++j) bits |= grid[(size_t)i * 8 + j] > thresh ? ((uint8_t)1 << j) : 0;
		bitfield[i] = bits;
	}
	for (uint32_t level = 1; level < N_CASCADES; ++level) {
		const uint8_t* prev = bitfield + (size_t)(level - 1) * GRID_N_CELLS / 8;
		uint8_t* next = bitfield + (size_t)level * GRID_N_CELLS / 8;
		for (uint32_t i = 0; i < GRID_N_CELLS / 64; ++i) {
			uint8_t bits = 0;
			for (uint8_t j = 0; j < 8; ++j) bits |= prev[(size_t)i * 8 + j] > 0 ? ((uint8_t)1 << j) : 0;
			uint32_t x = morton3D_invert(i >> 0) + GRIDSIZE / 8, y = morton3D_invert(i >> 1) + GRIDSIZE / 8, z = morton3D_invert(i >> 2) + GRIDSIZE / 8;
			next[morton3D(x, y, z)] |= bits;
		}
	}
}

/* ------------------------------------------------------------------------------------------------------------------
 * rendering: ray generation + occupancy marching (init_rays_with_payload_kernel_nerf :1414-1528, advance_pos_nerf :398-429,
 * generate_next_nerf_network_inputs :523-577) and compositing (composite_kernel_nerf :579-689, shade_kernel_nerf :1333-1378).
 * The march is independent of the network, so the oracle first lists every candidate sample of every ray (up to
 * max_steps), the test evaluates the network on them, and orc_render_composite applies early termination.
 * ---------------------------------------------------------------------------------------------------------------- */
static uint32_t reverse_bits(uint32_t x) {
	x = (((x & 0xaaaaaaaa) >> 1) | ((x & 0x55555555) << 1));
	x = (((x & 0xcccccccc) >> 2) | ((x & 0x33333333) << 2));
	x = (((x & 0xf0f0f0f0) >> 4) | ((x & 0x0f0f0f0f) << 4));
	x = (((x & 0xff00ff00) >> 8) | ((x & 0x00ff00ff) << 8));
	return ((x >> 16) | (x << 16));
}
static uint32_t laine_karras_permutation(uint32_t x, uint32_t seed) {
	x += seed; x ^= x * 0x6c50b47cu; x ^= x * 0xb82f1e52u; x ^= x * 0xc7afe638u; x ^= x * 0x8d22f6e6u;
	return x;
}
static uint32_t nested_uniform_scramble_base2(uint32_t x, uint32_t seed) { return reverse_bits(laine_karras_permutation(reverse_bits(x), seed)); }
static uint32_t hash_combine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }
static uint32_t sobol_dim0(uint32_t index) {
	/* random_val.cuh:162-176: direction numbers of dimension 0 are 2^(31-bit) */
	uint32_t X = 0;
	for (uint32_t bit = 0; bit < 32; bit++) { uint32_t mask = (index >> bit) & 1; X ^= mask * (0x80000000u >> bit); }
	return X;
}
static float ld_random_val(uint32_t index, uint32_t seed) {
	const float S = (float)(1.0 / (double)(1ull << 32));
	index = nested_uniform_scramble_base2(index, seed);
	return (float)nested_uniform_scramble_base2(sobol_dim0(index), hash_combine(seed, 0)) * S;
}
EXPORT float orc_ld_random_val(uint32_t index, uint32_t seed) { return ld_random_val(index, seed); }

/* counts[n_pixels], coords[n_pixels x max_steps x 7] (warped pos, warped dt, warped dir), ts/dts for compositing */
EXPORT void orc_render_march(const ngp_render_cfg* cfg, int32_t y0, int32_t y1, const uint8_t* bitfield, uint32_t max_steps, uint32_t* counts, float* coords) {
	aabb_t train_aabb = {V(cfg->aabb_min[0], cfg->aabb_min[1], cfg->aabb_min[2]), V(cfg->aabb_max[0], cfg->aabb_max[1], cfg->aabb_max[2])};
	aabb_t render_aabb = {V(cfg->render_aabb_min[0], cfg->render_aabb_min[1], cfg->render_aabb_min[2]), V(cfg->render_aabb_max[0], cfg->render_aabb_max[1], cfg->render_aabb_max[2])};
	const int W = cfg->width;
#pragma omp parallel for schedule(dynamic, 16)
	for (int32_t q = 0; q < (y1 - y0) * W; ++q) {
		uint32_t x = (uint32_t)(q % W), y = (uint32_t)(y0 + q / W);
		uint32_t idx = x + (uint32_t)W * y;
		counts[q] = 0;
		float u = ((float)x + 0.5f) / (float)cfg->width, v = ((float)y + 0.5f) / (float)cfg->height;
		v3 o, d;
		uv_to_ray(u, v, cfg->width, cfg->height, cfg->focal_x, cfg->focal_y, cfg->screen_x, cfg->screen_y, NGP_LENS_PERSPECTIVE, 0, cfg->camera, &o, &d);
		o = vadd(o, vscale(d, cfg->near_distance));
		d = vnormalize(d);
		v3 idir = V(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
		float tmin, tmax;
		aabb_ray_intersect(&render_aabb, o, d, &tmin, &tmax);
		float t = fmaxf(tmin, 0.0f) + 1e-6f;
		if (!aabb_contains(&render_aabb, vadd(o, vscale(d, t)))) continue;
		t = advance_n_steps(t, &cfg->march, ld_random_val(cfg->spp_index, idx * 786433u));
		v3 wdir = warp_direction(d);
		uint32_t n = 0;
		while (n < max_steps) {
			/* if_unoccupied_advance_to_next_occupied_voxel<false> (nerf_device.cuh:462-495) */
			int found = 0;
			v3 pos;
			for (;;) {
				pos = vadd(o, vscale(d, t));
				if (t >= MAX_DEPTH || !aabb_contains(&render_aabb, pos)) break;
				uint32_t mip = mip_from_pos(pos, N_CASCADES - 1);
				if (mip > cfg->max_cascade) mip = cfg->max_cascade;
				if (density_grid_occupied_at(pos, bitfield, mip)) { found = 1; break; }
				while (mip < cfg->max_cascade && !density_grid_occupied_at(pos, bitfield, mip + 1)) ++mip;
				t = advance_to_next_voxel(t, &cfg->march, pos, d, idir, mip);
			}
			if (!found) break;
			float dt = calc_dt(t, &cfg->march);
			v3 wp = warp_position(pos, &train_aabb);
			float* c = coords + ((size_t)q * max_steps + n) * 7;
			c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = warp_dt(dt); c[4] = wdir.x; c[5] = wdir.y; c[6] = wdir.z;
			t += dt;
			++n;
		}
		counts[q] = n;
	}
}
/* network_output: [n_pixels x max_steps x 4] binary16.  rgba_out [n_pixels x 4], depth_out [n_pixels], steps_used [n_pixels] */
EXPORT void orc_render_composite(const ngp_render_cfg* cfg, int32_t y0, int32_t y1, uint32_t max_steps, const uint32_t* counts, const float* coords,
	const uint16_t* network_output, float* rgba_out, float* depth_out, uint32_t* steps_used) {
	aabb_t train_aabb = {V(cfg->aabb_min[0], cfg->aabb_min[1], cfg->aabb_min[2]), V(cfg->aabb_max[0], cfg->aabb_max[1], cfg->aabb_max[2])};
	v3 cam_fwd = xform_col(cfg->camera, 2), cam_o = xform_col(cfg->camera, 3);
	for (int32_t q = 0; q < (y1 - y0) * cfg->width; ++q) {
		float r = 0, g = 0, b = 0, a = 0, max_weight = 0.0f, depth = MAX_DEPTH;
		uint32_t j = 0;
		/* the ray origin (NerfPayload::origin: after the near-plane offset), for the Depth mode */
		v3 ray_o, ray_d;
		{
			const uint32_t x = (uint32_t)(q % cfg->width), y = (uint32_t)(y0 + q / cfg->width);
			const float u = ((float)x + 0.5f) / (float)cfg->width, v = ((float)y + 0.5f) / (float)cfg->height;
			uv_to_ray(u, v, cfg->width, cfg->height, cfg->focal_x, cfg->focal_y, cfg->screen_x, cfg->screen_y, NGP_LENS_PERSPECTIVE, 0, cfg->camera, &ray_o, &ray_d);
			ray_o = vadd(ray_o, vscale(ray_d, cfg->near_distance));
		}
		for (; j < counts[q]; ++j) {
			const float* c = coords + ((size_t)q * max_steps + j) * 7;
			const uint16_t* o = network_output + ((size_t)q * max_steps + j) * 4;
			float T = 1.f - a;
			float dt = unwarp_dt(c[3]);
			float alpha = 1.f - ngp_expf(-network_to_density(half_to_float(o[3]), cfg->density_activation) * dt);
			float weight = alpha * T;
			/* composite_kernel_nerf :641-655 */
			v3 pos = unwarp_position(V(c[0], c[1], c[2]), &train_aabb);
			float cr, cg, cb;
			if (cfg->render_mode == NGP_RENDER_POSITIONS) { cr = (pos.x - 0.5f) / 2.0f + 0.5f; cg = (pos.y - 0.5f) / 2.0f + 0.5f; cb = (pos.z - 0.5f) / 2.0f + 0.5f; }
			else if (cfg->render_mode == NGP_RENDER_DEPTH) { cr = cg = cb = vdot(cam_fwd, vsub(pos, ray_o)) * cfg->depth_scale; }
			else if (cfg->render_mode == NGP_RENDER_AO) { cr = cg = cb = alpha; }
			else { cr = network_to_rgb(half_to_float(o[0]), cfg->rgb_activation); cg = network_to_rgb(half_to_float(o[1]), cfg->rgb_activation); cb = network_to_rgb(half_to_float(o[2]), cfg->rgb_activation); }
			r += cr * weight;
			g += cg * weight;
			b += cb * weight;
			a += weight;
			if (weight > max_weight) {
				max_weight = weight;
				depth = vdot(cam_fwd, vsub(pos, cam_o));
			}
			if (a > (1.0f - cfg->min_transmittance)) { r /= a; g /= a; b /= a; a /= a; ++j; break; }
		}
		if (steps_used) steps_used[q] = j;
		/* shade_kernel_nerf (:1333-1378): Cost = steps / 128 with alpha 1; Shade: predicted colours are sRGB (linear_colors == false) ->
		 * accumulate in linear; the other modes pass the composited value through */
		if (cfg->render_mode == NGP_RENDER_COST) { r = g = b = (float)j / 128.0f; a = 1.0f; }   /* every pixel, also rays without a sample */
		else if (cfg->render_mode == NGP_RENDER_SHADE) { r = srgb_to_linear(r); g = srgb_to_linear(g); b = srgb_to_linear(b); }
		rgba_out[(size_t)q * 4 + 0] = r;
		rgba_out[(size_t)q * 4 + 1] = g;
		rgba_out[(size_t)q * 4 + 2] = b;
		rgba_out[(size_t)q * 4 + 3] = a;
		depth_out[q] = a > 0.2f ? depth : MAX_DEPTH;
	}
}
/* ---- render epilogue: accumulate_kernel / tonemap_kernel (src/render_buffer.cu:228-262, 264-342, 511-545) ---- */
EXPORT void orc_accumulate(uint32_t n_px, const float* frame, float* acc, float sample_count, uint32_t color_space) {
	for (uint32_t i = 0; i < n_px; ++i) {
		float r = frame[4 * i], g = frame[4 * i + 1], b = frame[4 * i + 2];
		if (color_space == NGP_COLOR_SRGB) { r = linear_to_srgb(r); g = linear_to_srgb(g); b = linear_to_srgb(b); }
		acc[4 * i + 0] = (acc[4 * i + 0] * sample_count + r) / (sample_count + 1.0f);
		acc[4 * i + 1] = (acc[4 * i + 1] * sample_count + g) / (sample_count + 1.0f);
		acc[4 * i + 2] = (acc[4 * i + 2] * sample_count + b) / (sample_count + 1.0f);
		acc[4 * i + 3] = (acc[4 * i + 3] * sample_count + frame[4 * i + 3]) / (sample_count + 1.0f);
	}
}
static void tonemap_curve3(float* x, uint32_t curve) {
	if (curve == NGP_TONEMAP_IDENTITY) return;
	for (int k = 0; k < 3; ++k) x[k] = fmaxf(x[k], 0.0f);
	float k0, k1, k2, k3, k4, k5;
	if (curve == NGP_TONEMAP_ACES) {
		k0 = 0.6f * 0.6f * 2.51f; k1 = 0.6f * 0.03f; k2 = 0.0f; k3 = 0.6f * 0.6f * 2.43f; k4 = 0.6f * 0.59f; k5 = 0.14f;
	} else if (curve == NGP_TONEMAP_HABLE) {
		const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
		k0 = A * F - A * E; k1 = C * B * F - B * E; k2 = 0.0f; k3 = A * F; k4 = B * F; k5 = D * F * F;
		const float W = 11.2f;
		const float nom = k0 * (W * W) + k1 * W + k2, denom = k3 * (W * W) + k4 * W + k5;
		const float white_scale = denom / nom;
		k0 = 4.0f * k0 * white_scale; k1 = 2.0f * k1 * white_scale; k2 = k2 * white_scale; k3 = 4.0f * k3; k4 = 2.0f * k4;
	} else {
		const float Y = (0.2126f * x[0] + 0.7152f * x[1]) + 0.0722f * x[2];
		const float s = 1.0f / (Y + 1.0f);
		for (int k = 0; k < 3; ++k) x[k] = x[k] * s;
		return;
	}
	for (int k = 0; k < 3; ++k) {
		const float sq = x[k] * x[k];
		const float nom = (sq * k0 + k1 * x[k]) + k2, denom = (k3 * sq + k4 * x[k]) + k5;
		x[k] = nom / denom;
	}
}
EXPORT void orc_tonemap(uint32_t n_px, const ngp_tonemap_cfg* cfg, const float* acc, float* out) {
	float bg[4] = {cfg->background_color[0], cfg->background_color[1], cfg->background_color[2], cfg->background_color[3]};
	if (cfg->color_space != NGP_COLOR_SRGB)
		for (int k = 0; k < 3; ++k) bg[k] = srgb_to_linear(bg[k]);
	const float e = ngp_powf(2.0f, cfg->exposure);
	for (uint32_t i = 0; i < n_px; ++i) {
		const float* c = acc + 4 * (size_t)i;
		const float weight = (1.0f - c[3]) * bg[3];
		float col[3] = {c[0] + bg[0] * weight, c[1] + bg[1] * weight, c[2] + bg[2] * weight};
		float a = c[3] + weight;
		if (cfg->color_space == NGP_COLOR_SRGB)
			for (int k = 0; k < 3; ++k) col[k] = srgb_to_linear(col[k]);
		for (int k = 0; k < 3; ++k) col[k] = col[k] * e;
		tonemap_curve3(col, cfg->tonemap_curve);
		if (cfg->output_color_space == NGP_COLOR_SRGB)
			for (int k = 0; k < 3; ++k) col[k] = linear_to_srgb(col[k]);
		if (cfg->unmultiply_alpha && a > 0.0f)
			for (int k = 0; k < 3; ++k) col[k] = col[k] / a;
		if (cfg->clamp_output_color) {
			for (int k = 0; k < 3; ++k) col[k] = col[k] < 0.0f ? 0.0f : (col[k] > 1.0f ? 1.0f : col[k]);
			a = a < 0.0f ? 0.0f : (a > 1.0f ? 1.0f : a);
		}
		out[4 * (size_t)i + 0] = col[0]; out[4 * (size_t)i + 1] = col[1]; out[4 * (size_t)i + 2] = col[2]; out[4 * (size_t)i + 3] = a;
	}
}
EXPORT int orc_version(void) { return 1; }
