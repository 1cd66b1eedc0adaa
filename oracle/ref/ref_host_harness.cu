// ref_host_harness.cu — golden-vector generator that calls the REFERENCE's own NGP_HOST_DEVICE helpers on the CPU.
//
// Test infrastructure.  Compiled by oracle/ref/Makefile straight from the headers under /root/reference (nothing is copied
// into this repo) into oracle/_ref/ref_host; run here (no GPU needed) by tests/golden/make_ref_host_goldens.py, which commits
// the resulting vectors under tests/golden/.  The oracle (oracle/ngp_oracle.c) is then pinned against these vectors:
// integer results exactly, transcendental results to a few ulp (the reference uses libm logf/expf on the host, the oracle
// uses include/ngp_detmath.h).
//
// Functions exercised (all from /root/reference/include/neural-graphics-primitives/):
//   nerf_device.cuh: to_stepping_space, from_stepping_space, advance_n_steps, calc_dt, distance_to_next_voxel,
//                    advance_to_next_voxel, mip_from_pos, mip_from_dt, cascaded_grid_idx_at, density_grid_occupied_at,
//                    warp_dt/unwarp_dt, network_to_*, loss_and_gradient, image_idx
//   common_device.cuh: srgb_to_linear, linear_to_srgb, uv_to_ray (perspective + OpenCV lens), pos_to_uv
//   bounding_box.cuh: ray_intersect, contains;   random_val.cuh: ld_random_val;   pcg32.h;   tcnn morton3D
#include <neural-graphics-primitives/bounding_box.cuh>
#include <neural-graphics-primitives/common.h>
#include <neural-graphics-primitives/common_device.cuh>
#include <neural-graphics-primitives/nerf_device.cuh>
#include <neural-graphics-primitives/random_val.cuh>

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace ngp;

static FILE* g_out;
static void put_f(float v) { fwrite(&v, 4, 1, g_out); }
static void put_u(uint32_t v) { fwrite(&v, 4, 1, g_out); }

int main(int argc, char** argv) {
	if (argc < 2) {
		fprintf(stderr, "usage: ref_host <out.bin>\n");
		return 2;
	}
	g_out = fopen(argv[1], "wb");
	if (!g_out) return 3;
	pcg32 rng{20260922};

	// ---- section 1: stepping functions, 2 cone angles x 4096 t values -------------------------------------------------
	const float cones[2] = {0.0f, 1.0f / 256.0f};
	put_u(4096);
	for (int c = 0; c < 2; ++c) {
		for (int i = 0; i < 4096; ++i) {
			// t spans [1e-4, 300) log-uniformly
			float t = expf(logf(1e-4f) + rng.next_float() * (logf(300.0f) - logf(1e-4f)));
			put_f(t);
			put_f(to_stepping_space(t, cones[c]));
			put_f(from_stepping_space(to_stepping_space(t, cones[c]), cones[c]));
			put_f(calc_dt(t, cones[c]));
			put_f(advance_n_steps(t, cones[c], 0.37f));
		}
	}

	// ---- section 2: occupancy addressing, 8192 positions --------------------------------------------------------------
	put_u(8192);
	for (int i = 0; i < 8192; ++i) {
		vec3 pos = {rng.next_float() * 9.0f - 4.0f, rng.next_float() * 9.0f - 4.0f, rng.next_float() * 9.0f - 4.0f};
		if (i % 4 == 0) pos = {rng.next_float(), rng.next_float(), rng.next_float()};
		float dt = expf(logf(1e-3f) + rng.next_float() * (logf(2.0f) - logf(1e-3f)));
		uint32_t max_cascade = (uint32_t)(i % 8);
		put_f(pos.x); put_f(pos.y); put_f(pos.z); put_f(dt); put_u(max_cascade);
		put_u(mip_from_pos(pos, max_cascade));
		uint32_t mip = mip_from_dt(dt, pos, max_cascade);
		put_u(mip);
		put_u(cascaded_grid_idx_at(pos, mip));
		vec3 dir = normalize(vec3{rng.next_float() - 0.5f, rng.next_float() - 0.5f, rng.next_float() - 0.5f});
		vec3 idir = vec3(1.0f) / dir;
		put_f(dir.x); put_f(dir.y); put_f(dir.z);
		put_f(distance_to_next_voxel(pos, dir, idir, scalbnf((float)NERF_GRIDSIZE(), -(int)mip)));
		float t = rng.next_float() * 5.0f + 0.01f;
		put_f(t);
		put_f(advance_to_next_voxel(t, cones[1], pos, dir, idir, mip));
		put_f(advance_to_next_voxel(t, cones[0], pos, dir, idir, mip));
	}

	// ---- section 3: camera rays, 2048 uv samples, perspective and OpenCV lens (fox-like intrinsics) ---------------------
	put_u(2048);
	for (int i = 0; i < 2048; ++i) {
		vec2 uv = {rng.next_float(), rng.next_float()};
		ivec2 res = {1080, 1920};
		vec2 focal = {1375.52f, 1374.49f};
		vec2 pp = {554.558f / 1080.0f, 965.268f / 1920.0f};
		mat4x3 cam = {
			vec3{0.8f, 0.1f, -0.59f}, vec3{-0.2f, 0.97f, -0.1f}, vec3{0.56f, 0.2f, 0.8f}, vec3{0.3f + 0.01f * (i % 7), 0.6f, -0.4f},
		};
		Lens lens = {};
		lens.mode = (i & 1) ? ELensMode::OpenCV : ELensMode::Perspective;
		lens.params[0] = 0.0578421f; lens.params[1] = -0.0805099f; lens.params[2] = -0.000980296f; lens.params[3] = 0.00015575f;
		Ray ray = uv_to_ray(0, uv, res, focal, cam, pp, vec3(0.0f), 0.0f, 1.0f, 0.0f, {}, {}, lens);
		put_f(uv.x); put_f(uv.y); put_u((uint32_t)lens.mode);
		put_f(ray.o.x); put_f(ray.o.y); put_f(ray.o.z); put_f(ray.d.x); put_f(ray.d.y); put_f(ray.d.z);
		// pos_to_uv of a point on the ray returns the uv (used by mark_untrained_density_grid)
		vec2 uv2 = pos_to_uv(ray(2.0f), res, focal, cam, pp, vec3(0.0f), {}, lens);
		put_f(uv2.x); put_f(uv2.y);
		BoundingBox aabb{vec3(-1.5f), vec3(2.5f)};
		vec3 dn = normalize(ray.d);
		vec2 tmm = aabb.ray_intersect(ray.o, dn);
		put_f(tmm.x); put_f(tmm.y);
	}

	// ---- section 4: colour transfer, activations, losses --------------------------------------------------------------
	put_u(1024);
	for (int i = 0; i < 1024; ++i) {
		float v = (float)i / 1023.0f;
		put_f(v); put_f(srgb_to_linear(v)); put_f(linear_to_srgb(v));
		float x = (rng.next_float() - 0.5f) * 24.0f;
		put_f(x);
		put_f(network_to_rgb(x, ENerfActivation::Logistic)); put_f(network_to_rgb_derivative(x, ENerfActivation::Logistic));
		put_f(network_to_density(x, ENerfActivation::Exponential)); put_f(network_to_density_derivative(x, ENerfActivation::Exponential));
		vec3 target = {rng.next_float(), rng.next_float(), rng.next_float()}, pred = {rng.next_float() * 1.2f, rng.next_float(), rng.next_float()};
		for (int lt = 0; lt < 7; ++lt) {
			// ELossType order: L2, L1, Mape, Smape, Huber, LogL1, RelativeL2
			LossAndGradient lg = loss_and_gradient(target, pred, (ELossType)lt);
			put_f(target.x); put_f(pred.x); put_f(lg.loss.x); put_f(lg.gradient.x);
		}
		put_f(warp_dt(0.01f * (i + 1) / 64.0f)); put_f(unwarp_dt(v));
	}

	// ---- section 5: integer generators ---------------------------------------------------------------------------------
	put_u(1024);
	for (uint32_t i = 0; i < 1024; ++i) {
		uint32_t x = rng.next_uint() % 128, y = rng.next_uint() % 128, z = rng.next_uint() % 128;
		put_u(x); put_u(y); put_u(z); put_u(tcnn::morton3D(x, y, z)); put_u(tcnn::morton3D_invert(tcnn::morton3D(x, y, z) >> 1));
		put_f(ld_random_val(i % 4, i * 786433u));
		put_u(image_idx(i * 257u, 262144u, 0, 50u));
		pcg32 r{1337};
		r.advance((uint64_t)i * 16);
		put_u(r.next_uint());
		put_f(r.next_float());
	}

	// ---- section 6: a full training-ray march built from the reference helpers over a synthetic occupancy sphere --------
	// (the loop body is generate_training_samples_nerf's first pass, testbed_nerf.cu:793-807, with the reference's functions)
	{
		const uint32_t max_cascade = 2;
		std::vector<uint8_t> bitfield(NERF_GRID_N_CELLS() * NERF_CASCADES() / 8, 0);
		for (uint32_t mip = 0; mip <= max_cascade; ++mip) {
			for (uint32_t idx = 0; idx < NERF_GRID_N_CELLS(); ++idx) {
				uint32_t x = tcnn::morton3D_invert(idx >> 0), y = tcnn::morton3D_invert(idx >> 1), z = tcnn::morton3D_invert(idx >> 2);
				vec3 p = (vec3{(float)x + 0.5f, (float)y + 0.5f, (float)z + 0.5f} / (float)NERF_GRIDSIZE() - 0.5f) * scalbnf(1.0f, mip) + 0.5f;
				if (length(p - vec3(0.5f)) < 0.45f) bitfield[idx / 8 + grid_mip_offset(mip) / 8] |= (1 << (idx % 8));
			}
		}
		BoundingBox aabb{vec3(-1.5f), vec3(2.5f)};
		const float cone = 1.0f / 256.0f;
		put_u(4096);
		for (int i = 0; i < 4096; ++i) {
			vec3 o = {rng.next_float() * 3.0f - 1.0f, rng.next_float() * 3.0f - 1.0f, -1.4f};
			vec3 d = normalize(vec3{0.5f, 0.5f, 0.5f} + (vec3{rng.next_float(), rng.next_float(), rng.next_float()} - 0.5f) * 0.6f - o);
			vec2 tminmax = aabb.ray_intersect(o, d);
			tminmax.x = fmaxf(tminmax.x, 0.0f);
			float startt = advance_n_steps(tminmax.x, cone, rng.next_float());
			vec3 idir = vec3(1.0f) / d;
			uint32_t j = 0;
			float t = startt;
			vec3 pos;
			while (aabb.contains(pos = o + t * d) && j < NERF_STEPS()) {
				float dt = calc_dt(t, cone);
				uint32_t mip = mip_from_dt(dt, pos, max_cascade);
				if (density_grid_occupied_at(pos, bitfield.data(), mip)) {
					++j;
					t += dt;
				} else {
					t = advance_to_next_voxel(t, cone, pos, d, idir, mip);
				}
			}
			put_f(o.x); put_f(o.y); put_f(o.z); put_f(d.x); put_f(d.y); put_f(d.z); put_f(startt); put_u(j); put_f(t);
		}
	}
	fclose(g_out);
	return 0;
}
