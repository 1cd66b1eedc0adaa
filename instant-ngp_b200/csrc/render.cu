// render.cu — NeRF rendering as ONE persistent kernel: ray generation, occupancy-grid marching, fused network
// evaluation on tensor cores and front-to-back compositing without any HBM round trip for network inputs/outputs and
// without host synchronisation.
//
// ≙ render_nerf (src/testbed_nerf.cu:1894-2150): NerfTracer::init_rays_from_camera (init_rays_with_payload_kernel_nerf
// :1414-1528 + advance_pos_nerf :398-452), NerfTracer::trace (:1677-1860: compact_kernel_nerf / host sync per round,
// generate_next_nerf_network_inputs :523-577, inference, composite_kernel_nerf :579-689) and shade_kernel_nerf :1333-1378.
//
// Design: a CTA owns 128 ray slots (thread = slot = UMMA row).  Every iteration each live ray advances to its next
// occupied sample, the 128 samples go through the fused hash-grid + MLP evaluation (tcgen05), and each thread composites
// its own sample.  A ray that terminates writes its pixel and its slot immediately pulls the next pixel from a global
// queue, so tiles stay full until the image is exhausted (the reference either re-compacts with a host sync per round or,
// in its JIT megakernel, lets finished threads idle until the whole warp is done: fused_kernels/render_nerf.cuh:106-167).
// Per-ray arithmetic is the reference's; the order in which rays are processed does not influence any ray's result.
#include "march.cuh"
#include "nerf_net.cuh"

namespace ngpb {

// ---- low-discrepancy jitter (random_val.cuh:162-325): Owen-scrambled Sobol, dimension 0 ---------------------------------
__host__ __device__ inline uint32_t reverse_bits32(uint32_t x) {
	x = (((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1));
	x = (((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2));
	x = (((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4));
	x = (((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8));
	return (x >> 16) | (x << 16);
}
__host__ __device__ inline uint32_t laine_karras_permutation(uint32_t x, uint32_t seed) {
	x += seed;
	x ^= x * 0x6c50b47cu;
	x ^= x * 0xb82f1e52u;
	x ^= x * 0xc7afe638u;
	x ^= x * 0x8d22f6e6u;
	return x;
}
__host__ __device__ inline uint32_t nested_uniform_scramble_base2(uint32_t x, uint32_t seed) {
	return reverse_bits32(laine_karras_permutation(reverse_bits32(x), seed));
}
__host__ __device__ inline uint32_t hash_combine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }
// Sobol dimension 0 has the identity direction matrix in bit-reversed order: sobol(index, 0) == reverse_bits(index).
__host__ __device__ inline float ld_random_val_dim0(uint32_t index, uint32_t seed) {
	index = nested_uniform_scramble_base2(index, seed);
	return (float)nested_uniform_scramble_base2(reverse_bits32(index), hash_combine(seed, 0u)) * 2.3283064365386963e-10f;
}

// sobol(index, 1): the dimension-1 direction numbers are Pascal's triangle mod 2, v_k = v_{k-1} ^ (v_{k-1} >> 1) (random_val.cuh:172-179)
__host__ __device__ inline uint32_t sobol_dim1(uint32_t index) {
	uint32_t X = 0, v = 0x80000000u;
	for (uint32_t bit = 0; bit < 32; ++bit) {
		if ((index >> bit) & 1u) X ^= v;
		v ^= v >> 1;
	}
	return X;
}
// ld_random_val_2d (random_val.cuh:289-293, shuffled_scrambled_sobol2d :270-277)
__host__ __device__ inline void ld_random_val_2d(uint32_t index, uint32_t seed, float& x, float& y) {
	index = nested_uniform_scramble_base2(index, seed);
	x = (float)nested_uniform_scramble_base2(reverse_bits32(index), hash_combine(seed, 0u)) * 2.3283064365386963e-10f;
	y = (float)nested_uniform_scramble_base2(sobol_dim1(index), hash_combine(seed, 1u)) * 2.3283064365386963e-10f;
}
// ld_random_pixel_offset (random_val.cuh:320-325)
void render_pixel_offset(uint32_t spp, float* out) {
	float ax, ay, bx, by;
	ld_random_val_2d(0u, 0xdeadbeefu, ax, ay);
	ld_random_val_2d(spp, 0xdeadbeefu, bx, by);
	const float ox = (0.5f - ax) + bx, oy = (0.5f - ay) + by;
	out[0] = ox - floorf(ox);
	out[1] = oy - floorf(oy);
}

// ---- arithmetic of the render march -----------------------------------------------------------------------------------------
// RenderDet: march.cuh as is (ngp_detmath.h polynomials, IEEE division) — what the CPU oracle's render march reproduces bit for bit.
// RenderFast: the same stepping functions on the SFU (lg2 / ex2 / rcp approximations), i.e. the instructions the reference build's
//   --use_fast_math turns logf / expf / `/` into (nerf_device.cuh:379-441).  A voxel skip is two logs and one exp: 3 MUFU + ~20
//   instructions instead of ~300 of polynomial arithmetic, and on an aabb_scale-4 scene a 1920x1080 frame makes ~4e8 of them.
//   Positions agree with RenderDet to ~1e-6; it is the Testbed's default (`render_math`), the oracle tests pin RenderDet.
struct RenderDet {
	static __device__ __forceinline__ float calc_dt(float t, const ngp_march_consts& m) { return ngpb::calc_dt(t, m); }
	static __device__ __forceinline__ float advance_n_steps(float t, const ngp_march_consts& m, float n) { return ngpb::advance_n_steps(t, m, n); }
	static __device__ __forceinline__ float advance_to_next_voxel(float t, const ngp_march_consts& m, V3 pos, V3 dir, V3 idir, uint32_t mip) {
		return ngpb::advance_to_next_voxel(t, m, pos, dir, idir, mip);
	}
	static __device__ __forceinline__ float expf_(float x) { return ngp_expf(x); }
};
struct RenderFast {
	static __device__ __forceinline__ float to_step(float t, const ngp_march_consts& m) {
		if (m.cone_angle <= 1e-5f) return t * (1.0f / min_cone_stepsize());
		if (t <= m.at) return (t - m.at) * (1.0f / min_cone_stepsize()) + m.a;
		if (t <= m.bt) return __fdividef(__logf(t), m.log1p_c);
		return (t - m.bt) * (1.0f / max_cone_stepsize()) + m.b;
	}
	static __device__ __forceinline__ float from_step(float n, const ngp_march_consts& m) {
		if (m.cone_angle <= 1e-5f) return n * min_cone_stepsize();
		if (n <= m.a) return (n - m.a) * min_cone_stepsize() + m.at;
		if (n <= m.b) return __expf(n * m.log1p_c);
		return (n - m.b) * max_cone_stepsize() + m.bt;
	}
	static __device__ __forceinline__ float advance_n_steps(float t, const ngp_march_consts& m, float n) { return from_step(to_step(t, m) + n, m); }
	static __device__ __forceinline__ float calc_dt(float t, const ngp_march_consts& m) { return advance_n_steps(t, m, 1.0f) - t; }
	static __device__ __forceinline__ float advance_to_next_voxel(float t, const ngp_march_consts& m, V3 pos, V3 dir, V3 idir, uint32_t mip) {
		const float res = scalbnf(128.0f, -(int)mip);
		const float t_target = t + distance_to_next_voxel(pos, dir, idir, res);
		const float n = to_step(t, m), n_target = to_step(t_target, m);
		return from_step(n + ceilf(fmaxf(n_target - n, 0.5f)), m);
	}
	static __device__ __forceinline__ float expf_(float x) { return __expf(x); }
};

// init_rays_with_payload_kernel_nerf (:1414-1528) for pixel q of the tile: pixel offset = ld_random_pixel_offset(snap ? 0 : sample_index),
// the render camera's lens in uv_to_ray, near-plane offset, box entry and the jittered first step of advance_pos_nerf (:398-452).
// Returns false for a ray that misses the render box.
template <class M>
__device__ __forceinline__ bool render_init_ray(const ngp_render_cfg& cfg, const Aabb& render_aabb, const int32_t y0, const uint32_t n_rows, const uint32_t q, uint32_t& pix,
	V3& ro, V3& rd, V3& idir, float& t) {
	// queue order: 8 x 4 pixel blocks, row-major inside a block, blocks row-major over the tile — the 32 rays a warp pulls together are
	// neighbours in x AND y, so their samples share hash-grid cells (and 32-byte sectors) for longer than 32 pixels of one scanline do.
	// (Widths that are not a multiple of 8 / row counts that are not a multiple of 4 fall back to scanline order for the remainder.)
	uint32_t x, y;
	{
		const uint32_t w = (uint32_t)cfg.width, rows = n_rows;
		const uint32_t wb = w / 8u, hb = rows / 4u, n_blocked = wb * hb * 32u;
		if (q < n_blocked) {
			const uint32_t b = q / 32u, in = q % 32u;
			x = (b % wb) * 8u + (in % 8u);
			y = (b / wb) * 4u + (in / 8u);
		} else {
			// the strip right of the last full block column, then the rows below the last full block row
			uint32_t r = q - n_blocked;
			const uint32_t right_w = w - wb * 8u, right_n = right_w * hb * 4u;
			if (r < right_n) {
				x = wb * 8u + r % right_w;
				y = r / right_w;
			} else {
				r -= right_n;
				x = r % w;
				y = hb * 4u + r / w;
			}
		}
		y += (uint32_t)y0;
	}
	pix = x + (uint32_t)cfg.width * y;
	const float u = ((float)x + cfg.pixel_offset[0]) / (float)cfg.width, v = ((float)y + cfg.pixel_offset[1]) / (float)cfg.height;
	V3 o, d;
	uv_to_ray(u, v, cfg.width, cfg.height, cfg.focal_x, cfg.focal_y, cfg.screen_x, cfg.screen_y, cfg.lens_mode, cfg.lens_params, cfg.camera, o, d);
	o = o + d * cfg.near_distance;
	d = normalize3(d);
	ro = o;
	rd = d;
	idir = V3{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
	float tmin, tmax;
	aabb_ray_intersect(render_aabb, o, d, tmin, tmax);
	const float t0 = fmaxf(tmin, 0.0f) + 1e-6f;
	if (!render_aabb.contains(o + t0 * d)) return false;
	// advance_pos_nerf: jittered first step (ld_random_val(sample_index, idx * 786433))
	t = M::advance_n_steps(t0, cfg.march, ld_random_val_dim0(cfg.spp_index, pix * 786433u));
	return true;
}

// if_unoccupied_advance_to_next_occupied_voxel<MIP_FROM_DT = false> (nerf_device.cuh:462-495), at most `budget` voxel skips.
// Returns 1 when t sits on an occupied sample, 2 when the ray has left the box (t = MAX_DEPTH), 0 when the budget ran out first.
template <class M>
__device__ __forceinline__ uint32_t render_march(const ngp_render_cfg& cfg, const Aabb& render_aabb, const uint8_t* __restrict__ bitfield, const V3 ro, const V3 rd,
	const V3 idir, float& t, V3& pos, uint32_t budget) {
	for (;;) {
		pos = ro + t * rd;
		if (t >= max_depth() || !render_aabb.contains(pos)) {
			t = max_depth();
			return 2u;
		}
		uint32_t mip = mip_from_pos(pos, NGP_NERF_CASCADES - 1);
		mip = mip > cfg.max_cascade ? cfg.max_cascade : mip;  // clamp(mip, min_mip = 0, max_mip)
		if (density_grid_occupied_at(pos, bitfield, mip)) return 1u;
		if (budget == 0u) return 0u;
		--budget;
		while (mip < cfg.max_cascade && !density_grid_occupied_at(pos, bitfield, mip + 1)) ++mip;
		t = M::advance_to_next_voxel(t, cfg.march, pos, rd, idir, mip);
	}
}

// The way from the camera to the first occupied cell is the longest empty stretch of a ray (on nerf/fox ~100 voxel skips before a
// ray reaches the scene).  Inside the tile kernel every one of them would hold a 128-row tensor-core tile hostage (ncu r2a: 5.9 of
// 32 lanes active, 6e10 warp instructions, 186 ms per frame), so they are walked here, one thread per pixel, nothing else in the
// way: t_first[q] = the t of the ray's first sample, or MAX_DEPTH for a ray that never meets an occupied cell (its pixel is
// finished here).  Same functions, same values: a ray's samples do not depend on which kernel walked up to them.
template <class M>
__global__ void __launch_bounds__(128) k_render_first_hit(const __grid_constant__ ngp_render_cfg cfg, const int32_t y0, const uint32_t n_pixels,
	const uint8_t* __restrict__ bitfield, float* __restrict__ rgba_out, float* __restrict__ depth_out, float* __restrict__ t_first) {
	const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= n_pixels) return;
	const Aabb render_aabb{V3{cfg.render_aabb_min[0], cfg.render_aabb_min[1], cfg.render_aabb_min[2]},
		V3{cfg.render_aabb_max[0], cfg.render_aabb_max[1], cfg.render_aabb_max[2]}};
	uint32_t pix;
	V3 ro, rd, idir, pos;
	float t = max_depth(), t_end = 0.0f;
	if (render_init_ray<M>(cfg, render_aabb, y0, n_pixels / (uint32_t)cfg.width, q, pix, ro, rd, idir, t)) {
		render_march<M>(cfg, render_aabb, bitfield, ro, rd, idir, t, pos, 0xFFFFFFFFu);
		if (t < max_depth()) {
			// The other long empty stretch of a ray is the way OUT: a ray that does not saturate walks from the last surface to the far side of
			// the box (31 % of nerf/fox's pixels), again ~100 voxel skips with a couple of lanes of a tile's warp active.  Walk in from the far
			// side instead, here, with the same skip logic on the reversed ray: everything beyond its first occupied cell is empty, so the
			// tile kernel may end the ray at t_end.  Conservative (margin below); a ray's samples are unchanged.
			float tmin, tmax;
			aabb_ray_intersect(render_aabb, ro, rd, tmin, tmax);
			const V3 far = ro + tmax * rd, back = V3{-rd.x, -rd.y, -rd.z}, iback = V3{-idir.x, -idir.y, -idir.z};
			float tb = 1e-6f;
			V3 pb;
			render_march<M>(cfg, render_aabb, bitfield, far, back, iback, tb, pb, 0xFFFFFFFFu);
			// margin: the occupied cell found from behind extends at most one cell of the coarsest cascade further along the ray, and a
			// forward skip lands up to one step past a cell face — two cells + two steps
			const float t_hit = tmax - tb;
			t_end = tb >= max_depth() ? max_depth() : t_hit + scalbnf(1.0f / 64.0f, (int)cfg.max_cascade) + 2.0f * M::calc_dt(fmaxf(t_hit, 1e-3f), cfg.march);
			if (!(t_end > t)) t_end = max_depth();
		}
	}
	if (t >= max_depth()) {
		// shade_kernel_nerf on an empty payload: transparent pixel (Cost mode: 0 steps, alpha 1), depth MAX_DEPTH
		reinterpret_cast<float4*>(rgba_out)[pix] = make_float4(0.f, 0.f, 0.f, cfg.render_mode == NGP_RENDER_COST ? 1.f : 0.f);
		depth_out[pix] = max_depth();
	}
	t_first[q] = t;
	t_first[n_pixels + q] = t_end;
}

// Voxel skips a slot may spend per tile iteration looking for its next sample before the tile goes ahead without it.  The skips of a gap
// are serial per ray and a warp executes them with whatever lanes are in a gap at that moment (2 of 32 on nerf/fox): the smaller the budget,
// the more iterations a gap is spread over and the more lanes of a warp share each round of skips — at the price of the gap rays' rows
// sitting out more tiles.  Measured on nerf/fox 1920x1080 (profiles/r2/r2k_render_fox_skips.json), min. transmittance 0.01 / 1e-4:
// 32: 29.5 / 47.8 ms, 16: 25.4 / 41.1, 8: 22.3 / 35.8, 4: 20.7 / 33.2, 2: 20.7 / 33.1.
constexpr uint32_t RENDER_SKIPS_PER_TILE = 4;

constexpr uint32_t RENDER_CTAS_PER_SM = 4;       // 45 KB of shared memory and 64 TMEM columns each; <= 128 registers per thread

template <uint32_t F, class M>
__global__ void __launch_bounds__(TILE, RENDER_CTAS_PER_SM) k_render_nerf(
	const __grid_constant__ NetDev net, const __grid_constant__ ngp_render_cfg cfg, const int32_t y0, const int32_t y1,
	const __half* __restrict__ params, const uint8_t* __restrict__ bitfield, float* __restrict__ rgba_out, float* __restrict__ depth_out,
	uint32_t* __restrict__ queue /* [0] next pixel, [1] total steps */, const float* __restrict__ t_first
) {
	extern __shared__ __align__(128) uint8_t smem[];
	const FwdSmem L = fwd_smem_layout(net.n_hidden_density, net.n_hidden_rgb);
	const uint32_t tid = threadIdx.x, lane = tid & 31u;
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8);
	const uint32_t wr_off = mlp_n_params(net.n_hidden_density) * 2u;

	for (uint32_t l = 0; l <= net.n_hidden_density; ++l)
		stage_weights(params + net.density_off + mlp_layer_off(net.n_hidden_density, l), mlp_layer_out(net.n_hidden_density, l),
			mlp_layer_in(net.n_hidden_density, l), smem + mlp_layer_off(net.n_hidden_density, l) * 2u, tid, TILE);
	for (uint32_t l = 0; l <= net.n_hidden_rgb; ++l)
		stage_weights(params + net.rgb_off + mlp_layer_off(net.n_hidden_rgb, l), mlp_layer_out(net.n_hidden_rgb, l), mlp_layer_in(net.n_hidden_rgb, l),
			smem + wr_off + mlp_layer_off(net.n_hidden_rgb, l) * 2u, tid, TILE);
	if (tid < 32) umma::tmem_alloc<64>(tmem_slot);
	if (tid == 0) {
		umma::mbar_init(bar, 1);
		umma::mbar_fence_init();
	}
	umma::fence_before_sync();
	__syncthreads();
	umma::fence_after_sync();
	const uint32_t tmem_base = *tmem_slot;
	uint32_t phase = 0;
	const __half* grid = params + net.grid_off;

	const Aabb train_aabb{V3{cfg.aabb_min[0], cfg.aabb_min[1], cfg.aabb_min[2]}, V3{cfg.aabb_max[0], cfg.aabb_max[1], cfg.aabb_max[2]}};
	const Aabb render_aabb{V3{cfg.render_aabb_min[0], cfg.render_aabb_min[1], cfg.render_aabb_min[2]},
		V3{cfg.render_aabb_max[0], cfg.render_aabb_max[1], cfg.render_aabb_max[2]}};
	const uint32_t n_pixels = (uint32_t)(y1 - y0) * (uint32_t)cfg.width;
	const V3 cam_fwd = xform_col(cfg.camera, 2), cam_o = xform_col(cfg.camera, 3);

	// per-slot ray state (NerfPayload + accumulators)
	bool alive = false;
	uint32_t pix = 0, n_steps = 0;
	V3 ro{0, 0, 0}, rd{0, 0, 1}, idir{0, 0, 0};
	float t = 0.0f, t_end = 0.0f, max_weight = 0.0f, depth = 0.0f;
	float acc_r = 0.0f, acc_g = 0.0f, acc_b = 0.0f, acc_a = 0.0f;
	bool queue_empty = false;
	uint32_t local_steps = 0;

	auto finish_ray = [&]() {
		// shade_kernel_nerf (:1333-1378), frame buffer starts at zero
		float r = acc_r, g = acc_g, b = acc_b, a = acc_a;
		if (cfg.render_mode == NGP_RENDER_COST) {
			r = g = b = (float)n_steps / 128.0f;
			a = 1.0f;
		} else if (cfg.render_mode == NGP_RENDER_SHADE) {
			// rgb is predicted in sRGB (linear_colors == false): accumulate in linear colours
			r = srgb_to_linear(r);
			g = srgb_to_linear(g);
			b = srgb_to_linear(b);
		}
		float4 o = make_float4(r, g, b, a);
		reinterpret_cast<float4*>(rgba_out)[pix] = o;
		depth_out[pix] = acc_a > 0.2f ? depth : max_depth();
		alive = false;
	};

	for (;;) {
		// ---- refill empty slots from the pixel queue (one atomic per warp); rays without a first sample were finished by k_render_first_hit
		const bool want = !alive && !queue_empty;
		const uint32_t want_mask = __ballot_sync(0xFFFFFFFFu, want);
		if (want_mask) {
			uint32_t base = 0;
			if (lane == 0) base = atomicAdd(&queue[0], __popc(want_mask));
			base = __shfl_sync(0xFFFFFFFFu, base, 0);
			if (want) {
				const uint32_t q = base + __popc(want_mask & ((1u << lane) - 1u));
				if (q >= n_pixels) {
					queue_empty = true;
				} else {
					const float tf = t_first[q];
					if (tf < max_depth()) {
						float t_unused;
						render_init_ray<M>(cfg, render_aabb, y0, (uint32_t)(y1 - y0), q, pix, ro, rd, idir, t_unused);
						t = tf;
						t_end = t_first[n_pixels + q];
						acc_r = acc_g = acc_b = acc_a = 0.0f;
						max_weight = 0.0f;
						depth = max_depth();
						n_steps = 0;
						alive = true;
					}
				}
			}
		}

		// ---- march to the next occupied sample; a slot still inside an empty stretch after RENDER_SKIPS_PER_TILE skips sits this tile out
		bool has_sample = false;
		V3 pos{0.5f, 0.5f, 0.5f};
		float dt = 0.0f;
		if (alive) {
			// beyond t_end every cell is empty (k_render_first_hit walked in from the far side): the ray is over without walking out
			const uint32_t r = t > t_end ? 2u : render_march<M>(cfg, render_aabb, bitfield, ro, rd, idir, t, pos, cfg.skips_per_tile ? cfg.skips_per_tile : RENDER_SKIPS_PER_TILE);
			has_sample = r == 1u;
			if (r == 2u) finish_ray();
		}
		const uint32_t any_sample = __syncthreads_or(has_sample ? 1 : 0);
		const uint32_t any_pending = __syncthreads_or((alive || !queue_empty) ? 1 : 0);
		if (!any_sample) {
			if (!any_pending) break;
			continue;
		}

		// ---- network: encode + MLPs for the tile
		float wx = 0.5f, wy = 0.5f, wz = 0.5f;
		if (has_sample) {
			dt = M::calc_dt(t, cfg.march);
			const V3 wp = warp_position(pos, train_aabb);
			wx = wp.x; wy = wp.y; wz = wp.z;
		}
		{
			__half2 enc[16];
			grid_gather<F>(net, grid, wx, wy, wz, enc);
#pragma unroll
			for (uint32_t kc = 0; kc < 4; ++kc) {
				const __half2 h[4] = {enc[kc * 4 + 0], enc[kc * 4 + 1], enc[kc * 4 + 2], enc[kc * 4 + 3]};
				store_chunk(smem + L.a0_off, tid, kc, h);
			}
			const V3 wd = warp_direction(rd);
			__half2 sh[8];
			sh4_encode(wd.x, wd.y, wd.z, sh);
			const __half2 h0[4] = {sh[0], sh[1], sh[2], sh[3]};
			const __half2 h1[4] = {sh[4], sh[5], sh[6], sh[7]};
			store_chunk(smem + L.a2_off, tid, 2, h0);
			store_chunk(smem + L.a2_off, tid, 3, h1);
		}
		__half2 dens[8], rgbh[8];
		run_mlp_fwd(smem, L.a0_off, L.h_off, 0, net.n_hidden_density, tmem_base, bar, phase, tid, dens);
		{
			const __half2 h0[4] = {dens[0], dens[1], dens[2], dens[3]};
			const __half2 h1[4] = {dens[4], dens[5], dens[6], dens[7]};
			store_chunk(smem + L.a2_off, tid, 0, h0);
			store_chunk(smem + L.a2_off, tid, 1, h1);
		}
		run_mlp_fwd(smem, L.a2_off, L.h_off, wr_off, net.n_hidden_rgb, tmem_base, bar, phase, tid, rgbh);

		// ---- composite (composite_kernel_nerf :621-680); inputs round-trip through fp32 warp/unwarp like the reference
		if (has_sample) {
			++n_steps;
			++local_steps;
			const float o0 = __low2float(rgbh[0]), o1 = __high2float(rgbh[0]), o2 = __low2float(rgbh[1]), o3 = __low2float(dens[0]);
			const float T = 1.0f - acc_a;
			const float dtu = unwarp_dt(warp_dt(dt));
			const float alpha = 1.0f - M::expf_(-network_to_density(o3, cfg.density_activation) * dtu);
			const float weight = alpha * T;
			// composite_kernel_nerf :641-655: what the colour channels carry depends on the render mode
			const V3 p = unwarp_position(V3{wx, wy, wz}, train_aabb);
			float cr, cg, cb;
			if (cfg.render_mode == NGP_RENDER_POSITIONS) {
				cr = (p.x - 0.5f) / 2.0f + 0.5f; cg = (p.y - 0.5f) / 2.0f + 0.5f; cb = (p.z - 0.5f) / 2.0f + 0.5f;
			} else if (cfg.render_mode == NGP_RENDER_DEPTH) {
				cr = cg = cb = dot3(cam_fwd, p - ro) * cfg.depth_scale;     // payload.origin = the ray origin (after the near-plane offset)
			} else if (cfg.render_mode == NGP_RENDER_AO) {
				cr = cg = cb = alpha;
			} else {
				cr = network_to_rgb(o0, cfg.rgb_activation); cg = network_to_rgb(o1, cfg.rgb_activation); cb = network_to_rgb(o2, cfg.rgb_activation);
			}
			acc_r += cr * weight;
			acc_g += cg * weight;
			acc_b += cb * weight;
			acc_a += weight;
			if (weight > max_weight) {
				max_weight = weight;
				depth = dot3(cam_fwd, p - cam_o);
			}
			t += dt;
			if (acc_a > (1.0f - cfg.min_transmittance)) {
				acc_r /= acc_a;
				acc_g /= acc_a;
				acc_b /= acc_a;
				acc_a /= acc_a;
				finish_ray();
			}
		}
	}

	// total network evaluations, for Mrays/s and steps/ray reporting
#pragma unroll
	for (uint32_t o = 16; o > 0; o >>= 1) local_steps += __shfl_xor_sync(0xFFFFFFFFu, local_steps, o);
	if (lane == 0 && local_steps) atomicAdd(&queue[1], local_steps);

	umma::fence_before_sync();
	__syncthreads();
	if (tid < 32) umma::tmem_dealloc<64>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------
// k_nerf_forward_rays — the training-time inference pass, evaluated ray by ray with early termination.
//
// The reference evaluates the network on EVERY generated sample (m_network->inference_mixed_precision over max_inference
// samples, testbed_nerf.cu:3233-3235) and only then composites; compute_loss_kernel_train_nerf stops reading a ray's samples
// at the first one where the transmittance has dropped below 1e-4 (:926-929), so on a trained scene ~90 % of the evaluated
// samples are never looked at (measured here: 3.0 M evaluated for 0.26 M used per step).  This kernel produces exactly the
// network outputs the loss kernel will read and nothing else: a CTA holds 16 ray slots x 8 consecutive samples = one
// 128-row tensor-core tile; after each tile the 8 lanes of a slot advance the ray's transmittance with the loss kernel's own
// arithmetic (same expression order, ngp_expf, no FMA contraction) and a slot whose ray is exhausted or has T < 1e-4 pulls the
// next ray from a device-side queue.  Outputs for the samples the loss kernel reads are bit-identical to the all-samples pass.
// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t FWD_RAYS_CTAS_PER_SM = 3;   // measured: 3 -> 0.249 ms, 4 -> 0.280 ms, 5 (spills) -> 0.341 ms
template <uint32_t F, uint32_t RAY_CHUNK>
__global__ void __launch_bounds__(TILE, FWD_RAYS_CTAS_PER_SM) k_nerf_forward_rays(
	const __grid_constant__ NetDev net, const ngp_nerf_counters* __restrict__ counters, uint32_t* __restrict__ queue, const uint32_t* __restrict__ numsteps,
	const float* __restrict__ coords, const __half* __restrict__ params, const uint32_t density_activation, __half* __restrict__ out, const bool weight_sum_form
) {
	extern __shared__ __align__(128) uint8_t smem[];
	const FwdSmem L = fwd_smem_layout(net.n_hidden_density, net.n_hidden_rgb);
	const uint32_t tid = threadIdx.x, lane = tid & 31u;
	const uint32_t sub = lane & (RAY_CHUNK - 1), group_lane0 = lane & ~(RAY_CHUNK - 1);
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8);
	const uint32_t wr_off = mlp_n_params(net.n_hidden_density) * 2u;
	const uint32_t n_rays = counters->n_rays;
	if (blockIdx.x * (TILE / RAY_CHUNK) >= n_rays) return;

	for (uint32_t l = 0; l <= net.n_hidden_density; ++l)
		stage_weights(params + net.density_off + mlp_layer_off(net.n_hidden_density, l), mlp_layer_out(net.n_hidden_density, l),
			mlp_layer_in(net.n_hidden_density, l), smem + mlp_layer_off(net.n_hidden_density, l) * 2u, tid, TILE);
	for (uint32_t l = 0; l <= net.n_hidden_rgb; ++l)
		stage_weights(params + net.rgb_off + mlp_layer_off(net.n_hidden_rgb, l), mlp_layer_out(net.n_hidden_rgb, l), mlp_layer_in(net.n_hidden_rgb, l),
			smem + wr_off + mlp_layer_off(net.n_hidden_rgb, l) * 2u, tid, TILE);
	if (tid < 32) umma::tmem_alloc<64>(tmem_slot);
	if (tid == 0) {
		umma::mbar_init(bar, 1);
		umma::mbar_fence_init();
	}
	umma::fence_before_sync();
	__syncthreads();
	umma::fence_after_sync();
	const uint32_t tmem_base = *tmem_slot;
	uint32_t phase = 0;
	const __half* grid = params + net.grid_off;

	// slot state, replicated in the 8 lanes of the slot
	bool have_ray = false, queue_empty = false;
	uint32_t n = 0, base = 0, k0 = 0;
	float T = 1.0f, acc = 0.0f;   // `acc`: the accumulated weight of the fused train kernel's form of T (composite_step, march.cu)
	const float EPSILON = 1e-4f;

	for (;;) {
		// ---- refill: one queue pop per slot (lane `sub == 0` of each slot asks), one atomic per warp
		const bool want = !have_ray && !queue_empty && sub == 0;
		const uint32_t want_mask = __ballot_sync(0xFFFFFFFFu, want);
		if (want_mask) {
			uint32_t qbase = 0;
			if (lane == 0) qbase = atomicAdd(queue, __popc(want_mask));
			qbase = __shfl_sync(0xFFFFFFFFu, qbase, 0);
			uint32_t r = 0xFFFFFFFFu;
			if (want) r = qbase + __popc(want_mask & ((1u << lane) - 1u));
			r = __shfl_sync(0xFFFFFFFFu, r, group_lane0);
			if (!have_ray && !queue_empty) {
				if (r >= n_rays) {
					queue_empty = true;
				} else {
					n = numsteps[r * 2 + 0];
					base = numsteps[r * 2 + 1];
					k0 = 0;
					T = 1.0f;
					acc = 0.0f;
					have_ray = n > 0;
				}
			}
		}
		const uint32_t any_ray = __syncthreads_or(have_ray ? 1 : 0);
		if (!any_ray) {
			const uint32_t any_pending = __syncthreads_or(queue_empty ? 0 : 1);
			if (!any_pending) break;
			continue;
		}

		// ---- this row's sample
		const uint32_t k = k0 + sub;
		const bool valid = have_ray && k < n;
		float c[7] = {0.5f, 0.5f, 0.5f, 0.0f, 0.5f, 0.5f, 0.5f};
		const bool from_buffer = valid;
		if (from_buffer) {
			const float* cp = coords + (size_t)(base + k) * 7;
#pragma unroll
			for (int q = 0; q < 7; ++q) c[q] = cp[q];
		}
		{
			__half2 enc[16];
			grid_gather<F>(net, grid, c[0], c[1], c[2], enc);
#pragma unroll
			for (uint32_t kc = 0; kc < 4; ++kc) {
				const __half2 h[4] = {enc[kc * 4 + 0], enc[kc * 4 + 1], enc[kc * 4 + 2], enc[kc * 4 + 3]};
				store_chunk(smem + L.a0_off, tid, kc, h);
			}
			__half2 sh[8];
			sh4_encode(c[4], c[5], c[6], sh);
			const __half2 h0[4] = {sh[0], sh[1], sh[2], sh[3]};
			const __half2 h1[4] = {sh[4], sh[5], sh[6], sh[7]};
			store_chunk(smem + L.a2_off, tid, 2, h0);
			store_chunk(smem + L.a2_off, tid, 3, h1);
		}
		__half2 dens[8], rgbh[8];
		run_mlp_fwd(smem, L.a0_off, L.h_off, 0, net.n_hidden_density, tmem_base, bar, phase, tid, dens);
		{
			const __half2 h0[4] = {dens[0], dens[1], dens[2], dens[3]};
			const __half2 h1[4] = {dens[4], dens[5], dens[6], dens[7]};
			store_chunk(smem + L.a2_off, tid, 0, h0);
			store_chunk(smem + L.a2_off, tid, 1, h1);
		}
		run_mlp_fwd(smem, L.a2_off, L.h_off, wr_off, net.n_hidden_rgb, tmem_base, bar, phase, tid, rgbh);
		float alpha = 0.0f;
		if (valid) {
			uint2 o;
			o.x = *reinterpret_cast<const uint32_t*>(&rgbh[0]);
			const __half2 t = __halves2half2(__low2half(rgbh[1]), __low2half(dens[0]));
			o.y = *reinterpret_cast<const uint32_t*>(&t);
			*reinterpret_cast<uint2*>(out + (size_t)(base + k) * 4) = o;
			// compute_loss_kernel_train_nerf :931-944 — the density read back from the fp16 output, as the loss kernel will
			const float density = network_to_density(__half2float(__low2half(dens[0])), density_activation);
			alpha = 1.0f - ngp_expf(-density * unwarp_dt(c[3]));
		}
		// ---- advance the slot's transmittance in sample order (all 8 lanes of the slot compute the same sequence)
		if (have_ray) {
			bool done = false;
			for (uint32_t j = 0; j < RAY_CHUNK; ++j) {
				const float a = __shfl_sync(0xFFFFFFFFu, alpha, group_lane0 + j);
				if (done) continue;
				if (k0 + j >= n) { done = true; continue; }
				// the loss kernel reads sample k0+j only if T >= EPSILON before it; it was evaluated above either way
				if (T < EPSILON) { done = true; continue; }
				if (weight_sum_form) {
					acc += a * T;
					T = 1.0f - acc;
				} else {
					T *= (1.0f - a);
				}
			}
			k0 += RAY_CHUNK;
			if (done || k0 >= n || T < EPSILON) have_ray = false;
		} else {
#pragma unroll
			for (uint32_t j = 0; j < RAY_CHUNK; ++j) (void)__shfl_sync(0xFFFFFFFFu, alpha, group_lane0 + j);
		}
	}
	umma::fence_before_sync();
	__syncthreads();
	if (tid < 32) umma::tmem_dealloc<64>(tmem_base);
}

// queue: a zeroed u32 (the `pad` word of the step's counter block).  Grid sized for the worst case (n_rays_max rays).
template <uint32_t F, uint32_t CHUNK>
static void launch_forward_rays(const NetDev& net, cudaStream_t stream, uint32_t n_rays_max, const ngp_nerf_counters* counters, uint32_t* queue,
	const uint32_t* numsteps, const float* coords, const __half* params, uint32_t density_activation, __half* out, bool weight_sum_form) {
	const FwdSmem L = fwd_smem_layout(net.n_hidden_density, net.n_hidden_rgb);
	const uint32_t n_tiles = div_round_up(n_rays_max, TILE / CHUNK);
	const uint32_t max_ctas = (uint32_t)device_sm_count() * FWD_RAYS_CTAS_PER_SM;
	const uint32_t grid = n_tiles < max_ctas ? n_tiles : max_ctas;
	auto kern = k_nerf_forward_rays<F, CHUNK>;
	static bool attr = false;
	if (!attr) { NGPB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr = true; }
	kern<<<grid, TILE, L.total, stream>>>(net, counters, queue, numsteps, coords, params, density_activation, out, weight_sum_form);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

// chunk: consecutive samples of a ray evaluated per tensor-core tile (4, 8, 16 or 32; 128 / chunk ray slots per CTA)
// train_mode: whose form of the transmittance decides where a ray stops (composite_step, march.cu): the loss kernel's or the fused train kernel's
void nerf_inference_rays(const ngp_nerf_desc& d, cudaStream_t stream, uint32_t n_rays_max, const ngp_nerf_counters* counters, uint32_t* queue,
	const uint32_t* numsteps, const float* coords, const __half* params, uint32_t density_activation, __half* out, uint32_t chunk, uint32_t train_mode) {
	if (n_rays_max == 0) return;
	NGPB_CHECK(chunk == 4 || chunk == 8 || chunk == 16 || chunk == 32, "inference chunk must be 4, 8, 16 or 32");
	const NetDev net = make_netdev(d);
#define NGPB_FWD_RAYS(FF, CC) launch_forward_rays<FF, CC>(net, stream, n_rays_max, counters, queue, numsteps, coords, params, density_activation, out, train_mode != NGP_TRAIN_NERF)
	if (net.n_features == 2) {
		if (chunk == 4) NGPB_FWD_RAYS(2, 4); else if (chunk == 8) NGPB_FWD_RAYS(2, 8); else if (chunk == 16) NGPB_FWD_RAYS(2, 16); else NGPB_FWD_RAYS(2, 32);
	} else {
		if (chunk == 4) NGPB_FWD_RAYS(4, 4); else if (chunk == 8) NGPB_FWD_RAYS(4, 8); else if (chunk == 16) NGPB_FWD_RAYS(4, 16); else NGPB_FWD_RAYS(4, 32);
	}
#undef NGPB_FWD_RAYS
}

size_t render_scratch_bytes(int32_t width, int32_t rows) { return 256 + 2 * sizeof(float) * (size_t)(width > 0 ? width : 0) * (size_t)(rows > 0 ? rows : 0); }

template <uint32_t F, class M>
static void launch_render(const NetDev& net, cudaStream_t stream, const ngp_render_cfg& cfg, int32_t y0, int32_t y1, const __half* params, const uint8_t* bitfield,
	float* rgba, float* depth, uint32_t* queue, float* t_first) {
	const FwdSmem L = fwd_smem_layout(net.n_hidden_density, net.n_hidden_rgb);
	const uint32_t n_pixels = (uint32_t)(y1 - y0) * (uint32_t)cfg.width;
	k_render_first_hit<M><<<div_round_up(n_pixels, 128), 128, 0, stream>>>(cfg, y0, n_pixels, bitfield, rgba, depth, t_first);
	NGPB_LAUNCHED();
	const uint32_t n_tiles = div_round_up(n_pixels, TILE);
	const uint32_t max_ctas = (uint32_t)device_sm_count() * RENDER_CTAS_PER_SM;
	const uint32_t grid = n_tiles < max_ctas ? n_tiles : max_ctas;
	auto kern = k_render_nerf<F, M>;
	static bool attr = false;
	if (!attr) { NGPB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr = true; }
	kern<<<grid, TILE, L.total, stream>>>(net, cfg, y0, y1, params, bitfield, rgba, depth, queue, t_first);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

void render_nerf(const ngp_nerf_desc& d, cudaStream_t stream, const ngp_render_cfg& cfg, int32_t y0, int32_t y1, const __half* params,
	const uint8_t* bitfield, float* rgba, float* depth, void* scratch, uint32_t* n_steps_total) {
	NGPB_CHECK(cfg.width > 0 && cfg.height > 0 && y0 >= 0 && y1 <= cfg.height && y0 < y1, "render: bad tile");
	NGPB_CHECK(cfg.math_mode <= NGP_MATH_REFERENCE, "ngp_render_cfg.math_mode: unknown arithmetic flavour");
	NGPB_CHECK(cfg.render_mode == NGP_RENDER_SHADE || cfg.render_mode == NGP_RENDER_AO || cfg.render_mode == NGP_RENDER_POSITIONS || cfg.render_mode == NGP_RENDER_DEPTH ||
		cfg.render_mode == NGP_RENDER_COST, "render: this build renders the modes Shade, AO, Positions, Depth and Cost (Normals needs network input gradients; Distortion / Slice are 2-D debug views)");
	const NetDev net = make_netdev(d);
	uint32_t* queue = reinterpret_cast<uint32_t*>(scratch);
	float* t_first = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(scratch) + 256);
	NGPB_CUDA_CHECK(cudaMemsetAsync(queue, 0, 16, stream));
	const bool fast = cfg.math_mode == NGP_MATH_REFERENCE;
	if (net.n_features == 2) {
		if (fast) launch_render<2, RenderFast>(net, stream, cfg, y0, y1, params, bitfield, rgba, depth, queue, t_first);
		else launch_render<2, RenderDet>(net, stream, cfg, y0, y1, params, bitfield, rgba, depth, queue, t_first);
	} else {
		if (fast) launch_render<4, RenderFast>(net, stream, cfg, y0, y1, params, bitfield, rgba, depth, queue, t_first);
		else launch_render<4, RenderDet>(net, stream, cfg, y0, y1, params, bitfield, rgba, depth, queue, t_first);
	}
	if (n_steps_total) NGPB_CUDA_CHECK(cudaMemcpyAsync(n_steps_total, queue + 1, 4, cudaMemcpyDeviceToDevice, stream));
}

// ------------------------------------------------------------------------------------------------------------------
// render epilogue (src/render_buffer.cu:228-262, 264-342, 511-545)
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_accumulate(const uint32_t n_px, const float* __restrict__ frame, float* __restrict__ acc, const float sample_count, const uint32_t color_space) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_px) return;
	const float4 c = reinterpret_cast<const float4*>(frame)[i];
	float4 t = reinterpret_cast<float4*>(acc)[i];
	float r = c.x, g = c.y, b = c.z;
	if (color_space == NGP_COLOR_SRGB) {
		r = linear_to_srgb(r);
		g = linear_to_srgb(g);
		b = linear_to_srgb(b);
	}
	t.x = (t.x * sample_count + r) / (sample_count + 1.0f);
	t.y = (t.y * sample_count + g) / (sample_count + 1.0f);
	t.z = (t.z * sample_count + b) / (sample_count + 1.0f);
	t.w = (t.w * sample_count + c.w) / (sample_count + 1.0f);
	reinterpret_cast<float4*>(acc)[i] = t;
}

__host__ __device__ inline void tonemap_curve3(float (&x)[3], uint32_t curve) {
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

__global__ void k_tonemap(const uint32_t n_px, const ngp_tonemap_cfg cfg, const float* __restrict__ acc, float* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_px) return;
	float bg[4] = {cfg.background_color[0], cfg.background_color[1], cfg.background_color[2], cfg.background_color[3]};
	if (cfg.color_space != NGP_COLOR_SRGB)
		for (int k = 0; k < 3; ++k) bg[k] = srgb_to_linear(bg[k]);
	const float4 c = reinterpret_cast<const float4*>(acc)[i];
	const float weight = (1.0f - c.w) * bg[3];
	float col[3] = {c.x + bg[0] * weight, c.y + bg[1] * weight, c.z + bg[2] * weight};
	float a = c.w + weight;
	if (cfg.color_space == NGP_COLOR_SRGB)
		for (int k = 0; k < 3; ++k) col[k] = srgb_to_linear(col[k]);
	const float e = ngp_powf(2.0f, cfg.exposure);
	for (int k = 0; k < 3; ++k) col[k] = col[k] * e;
	tonemap_curve3(col, cfg.tonemap_curve);
	if (cfg.output_color_space == NGP_COLOR_SRGB)
		for (int k = 0; k < 3; ++k) col[k] = linear_to_srgb(col[k]);
	if (cfg.unmultiply_alpha && a > 0.0f)
		for (int k = 0; k < 3; ++k) col[k] = col[k] / a;
	if (cfg.clamp_output_color) {
		for (int k = 0; k < 3; ++k) col[k] = clampf(col[k], 0.0f, 1.0f);
		a = clampf(a, 0.0f, 1.0f);
	}
	reinterpret_cast<float4*>(out)[i] = make_float4(col[0], col[1], col[2], a);
}

void render_accumulate(cudaStream_t stream, int32_t w, int32_t h, const float* frame, float* acc, float sample_count, uint32_t color_space) {
	NGPB_CHECK(w > 0 && h > 0, "render_accumulate: bad frame size");
	const uint32_t n = (uint32_t)w * (uint32_t)h;
	k_accumulate<<<div_round_up(n, 256), 256, 0, stream>>>(n, frame, acc, sample_count, color_space);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}
void render_tonemap(cudaStream_t stream, int32_t w, int32_t h, const ngp_tonemap_cfg& cfg, const float* acc, float* out) {
	NGPB_CHECK(w > 0 && h > 0, "render_tonemap: bad frame size");
	NGPB_CHECK(cfg.tonemap_curve <= NGP_TONEMAP_REINHARD, "render_tonemap: unknown tonemap curve");
	const uint32_t n = (uint32_t)w * (uint32_t)h;
	k_tonemap<<<div_round_up(n, 256), 256, 0, stream>>>(n, cfg, acc, out);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

// sum of n floats in a fixed order (1024 strided partials, then sequential) — NerfCounters::update_after_training's
// reduce_sum(loss) (testbed_nerf.cu:2693-2696), made order-deterministic.
__global__ void k_sum_partial(const float* __restrict__ data, uint32_t n, float* __restrict__ partial) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	float s = 0.0f;
	for (uint32_t i = t; i < n; i += 1024) s += data[i];
	partial[t] = s;
}
float reduce_sum_f32(cudaStream_t stream, const float* data, uint32_t n, float* scratch_dev) {
	k_sum_partial<<<4, 256, 0, stream>>>(data, n, scratch_dev);
	NGPB_LAUNCHED();
	float host[1024];
	NGPB_CUDA_CHECK(cudaMemcpyAsync(host, scratch_dev, sizeof(host), cudaMemcpyDeviceToHost, stream));
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
	float s = 0.0f;
	for (int i = 0; i < 1024; ++i) s += host[i];
	return s;
}

}  // namespace ngpb
