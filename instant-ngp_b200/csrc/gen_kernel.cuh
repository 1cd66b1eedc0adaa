// gen_kernel.cuh — the training sample generator (≙ generate_training_samples_nerf, src/testbed_nerf.cu:691-849), one kernel body
// for two arithmetic flavours.
//
// The body is a template over a march policy P that supplies the floating-point steps of the march (ray construction, calc_dt,
// advance_to_next_voxel, mip_from_dt, the warps).  It is instantiated twice, in two translation units with different compiler flags:
//   march.cu      DetMarch  -fmad=false, transcendentals from ngp_detmath.h: bit-identical to the CPU oracle (oracle/ngp_oracle.c)
//   march_ref.cu  RefMarch  --use_fast_math like the reference build (CMakeLists.txt:88), the reference's own expression trees
//                           (march_ref.cuh): the per-ray sample counts of the reference's kernel
//
// Structure.  A march step is t -> occupied(t) ? t + dt(t) : advance_to_next_voxel(t): a serial recurrence through log / exp (cone
// stepping), and per step position, mip level, Morton code and a bitfield load that depend on it.  With one thread per ray (the
// reference, and round 1 here) a step costs ~1000 cycles of dependent instructions and the kernel lasts as long as its longest ray:
// 1.46 ms of a 2.76 ms step on nerf/fox, whose batch is only 3.6 K rays of up to 1024 samples (profiles/r2a).  Here G lanes of a warp
// work on ONE ray:
//   * all G lanes run the recurrence alone for S <= G steps, as if every cell were occupied (identical instructions, nothing but the
//     dependent arithmetic); lane g keeps step g's (t, dt);
//   * every lane tests ITS sample: position, box, mip, occupancy bit — G tests, G loads in flight at once;
//   * a ballot gives the length of the occupied prefix: those samples are accepted (lane g stores sample j + g), the first empty cell
//     is skipped from by the lane that found it (advance_to_next_voxel, broadcast), a position outside the box ends the ray.
//   S restarts at 1 after a skip and doubles with every fully accepted round.  Every value is a pure function of t, so samples, their
//   t / dt and the skips are exactly those of the sequential loop.  G is chosen per launch from the batch size (32 K rays and more:
//   one thread per ray as before; nerf/fox: 16 lanes per ray), keeping the launch within one wave of warps.
//   * pass 1 leaves every sample's t in shared memory (G >= 16: all 1024; G < 16: the first 64 plus a checkpoint every 32 samples
//     beyond); the warp then writes the 28-byte coordinate records cooperatively (lanes = consecutive samples of one ray, transposed
//     through shared memory so every store covers 128 contiguous bytes) and, for G < 16, re-marches the tails of long rays segment by
//     segment from the checkpoints — the reference marches every ray twice (count, then write).
//   * slots are reserved once per warp (prefix sum + one atomic) instead of two atomics per ray (testbed_nerf.cu:812,819).
// Ray ids are GLOBAL: ray id = ray_offset + ray_stride * local index, so that W ranks reproduce the single-GPU batch.
#pragma once

#include "march.cuh"

namespace ngpb {

constexpr uint32_t GEN_SEG = 32;
__host__ __device__ constexpr uint32_t gen_t_slots(uint32_t G) { return G >= 16 ? NGP_NERF_STEPS : 64u; }
__host__ __device__ constexpr uint32_t gen_n_ckpt(uint32_t G) { return G >= 16 ? 0u : (NGP_NERF_STEPS - 64u + GEN_SEG - 1) / GEN_SEG; }
__host__ __device__ constexpr uint32_t gen_smem_bytes(uint32_t G) { return (gen_t_slots(G) + gen_n_ckpt(G)) * (32u / G) * (uint32_t)sizeof(float); }

// The march of one ray by the G lanes of its group, from loop state (t, j) until the ray leaves the box or has j_end samples.
// emit(jj, t, dt, pos) is called by the lane that holds sample jj.  t, j, j_end and the ray are uniform across the group; all lanes of
// the warp must call this together (groups whose ray is finished pass j >= j_end).  Equivalent, value for value, to
//     while (aabb.contains(pos = o + t d) && j < j_end) { dt = calc_dt(t); mip = mip_from_dt(dt, pos);
//         if (occupied(pos, mip)) { emit(j, t, dt, pos); ++j; t += dt; } else t = advance_to_next_voxel(t, pos, d, idir, mip); }
// schedule knobs of march_group (ngp_nerf_train_cfg.gen_walk_empty / gen_speculation override them; every ray's samples are the same for any value)
// Measured on nerf/fox, 16 lanes per ray (profiles/r2/r2o_generator_knobs.jsonl): walk 1 / speculation 1: 0.361 ms, 8 / 4: 0.332, 64 / 4: 0.382,
// 1 / 16: 0.681 (the recurrence's latency, not the round's instruction count, is what a deeper speculation pays); with one lane per ray a
// walk only makes the warp's other rays wait (synthetic scene 0.56 -> 1.09 ms at 64).
__host__ __device__ constexpr uint32_t gen_walk_empty_default(uint32_t G) { return G >= 8 ? 8u : 1u; }    // cells the lane that found an empty cell crosses before the group's next round
__host__ __device__ constexpr uint32_t gen_speculation_default(uint32_t G) { return G >= 8 ? 4u : 1u; }   // samples speculated in the first round after a skip (doubles after every fully occupied round)

template <class P, uint32_t G, class Emit>
__device__ __forceinline__ uint32_t march_group(const typename P::Ctx& pc, const Aabb& aabb, const uint32_t max_cascade, const uint8_t* __restrict__ bitfield,
	const V3 ro, const V3 rd, const V3 idir, float t, uint32_t j, const uint32_t j_end, const uint32_t walk_empty, const uint32_t s_after_skip, Emit emit) {
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t g = lane & (G - 1u), lane0 = lane & ~(G - 1u);
	const uint32_t gmask = G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u);
	bool running = j < j_end;
	uint32_t S = s_after_skip;
	while (__any_sync(0xFFFFFFFFu, running)) {
		// ---- the recurrence alone: S steps as if every cell were occupied; lane g keeps step g
		float my_t = t, my_dt = 0.0f, tc = t;
		if (running) {
			for (uint32_t s = 0; s < S; ++s) {
				const float dt = P::calc_dt(tc, pc);
				if (s == g) {
					my_t = tc;
					my_dt = dt;
				}
				tc += dt;
			}
		}
		// ---- every lane tests its own sample
		const bool mine = running && g < S;
		const V3 pos = P::ray_pos(ro, my_t, rd);
		bool inside = false, occ = false;
		uint32_t mip = 0;
		if (mine) {
			inside = aabb.contains(pos);
			mip = P::mip_from_dt(my_dt, pos, max_cascade);
			occ = inside && density_grid_occupied_at(pos, bitfield, mip);
		}
		const uint32_t okmask = (__ballot_sync(0xFFFFFFFFu, occ) >> lane0) & gmask;
		// length of the occupied prefix, 0..S (lanes g >= S report false; ~okmask is zero only when G == 32 and every lane is occupied)
		const uint32_t n_ok = okmask == 0xFFFFFFFFu ? 32u : (uint32_t)(__ffs((int)~okmask) - 1);
		const uint32_t room = running ? j_end - j : 0u;
		const uint32_t n_take = n_ok < room ? n_ok : room;
		if (g < n_take) emit(j + g, my_t, my_dt, pos);
		j += n_take;
		const bool budget_hit = running && j >= j_end;             // `j < NERF_STEPS` fails in the sequential loop
		const bool clean = n_ok == S;                              // the whole round was occupied
		// lane n_ok of the group holds the first sample that is not occupied: outside the box -> the ray ends; empty -> skip from there
		float t_new = 0.0f;
		if (running && !budget_hit && !clean && g == n_ok && inside) {
			t_new = P::advance_to_next_voxel(my_t, pc, pos, rd, idir, mip);
			// this lane keeps walking while the cells it lands in are empty (the sequential loop's next iterations, one lane, none of the
			// round's fixed cost): a ray crosses ~100 empty cells between the camera and the scene
			for (uint32_t w = 0; w < walk_empty; ++w) {
				const V3 p2 = P::ray_pos(ro, t_new, rd);
				if (!aabb.contains(p2)) {
					inside = false;
					break;
				}
				const uint32_t mip2 = P::mip_from_dt(P::calc_dt(t_new, pc), p2, max_cascade);
				if (density_grid_occupied_at(p2, bitfield, mip2)) break;
				t_new = P::advance_to_next_voxel(t_new, pc, p2, rd, idir, mip2);
			}
		}
		const uint32_t fail_lane = lane0 + (n_ok < G ? n_ok : G - 1u);
		const bool fail_inside = __shfl_sync(0xFFFFFFFFu, inside ? 1 : 0, fail_lane) != 0;
		const float t_skip = __shfl_sync(0xFFFFFFFFu, t_new, fail_lane);
		if (running) {
			if (budget_hit) {
				running = false;
			} else if (clean) {
				t = tc;                                                // carry on, twice as deep
				S = S * 2u < G ? S * 2u : G;
			} else {
				if (!fail_inside) running = false;
				t = t_skip;
				S = s_after_skip;
			}
		}
	}
	return j;
}

template <class P, uint32_t G>
__global__ void __launch_bounds__(32) k_generate_training_samples(
	const uint32_t n_rays_local, const uint32_t ray_offset, const uint32_t ray_stride, const uint32_t n_rays_global, Pcg32 rng_in, const ngp_nerf_train_cfg cfg,
	const ngp_train_view* __restrict__ views, const uint32_t n_views, const uint8_t* __restrict__ bitfield, const uint32_t max_samples,
	ngp_nerf_counters* __restrict__ counters, uint32_t* __restrict__ ray_indices_out, float* __restrict__ rays_out,
	uint32_t* __restrict__ numsteps_out, float* __restrict__ coords_out
) {
	constexpr uint32_t NG = 32u / G;                 // rays per warp
	constexpr uint32_t T_SLOTS = gen_t_slots(G);
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t g = lane & (G - 1u), grp = lane / G, lane0 = lane & ~(G - 1u);
	const uint32_t li = blockIdx.x * NG + grp;
	const bool in_range = li < n_rays_local;
	const uint32_t i = ray_offset + li * ray_stride;

	const Aabb aabb{V3{cfg.aabb_min[0], cfg.aabb_min[1], cfg.aabb_min[2]}, V3{cfg.aabb_max[0], cfg.aabb_max[1], cfg.aabb_max[2]}};
	const typename P::Ctx pc = P::make_ctx(cfg);
	extern __shared__ float t_list[];   // [T_SLOTS][NG], then checkpoints [N_CKPT][NG]
	float* ckpt = t_list + T_SLOTS * NG;
	__shared__ float coord_tile[32 * 7];
	V3 ro{0, 0, 0}, rd{0, 0, 0}, rdn{0, 0, 1}, idir{0, 0, 0};
	float startt = 0.0f;
	bool live = false;

	if (in_range) {
		// the G lanes of a group build the same ray (identical instructions on identical values)
		const uint32_t img = image_idx(i, n_rays_global, n_views);
		const ngp_train_view vw = views[img];
		Pcg32 rng = rng_in;
		rng.advance((uint64_t)i * N_MAX_RANDOM_SAMPLES_PER_RAY);
		float u, v;
		P::random_image_pos(rng, vw.width, vw.height, cfg.snap_to_pixel_centers != 0, u, v);
		const bool masked = !vw.no_mask && pixel_is_masked(u, v, vw.width, vw.height, vw.pixels, vw.image_type);
		if (!masked) {
			(void)rng.next_float();  // motion-blur time (testbed_nerf.cu:740) — consumed, unused without rolling shutter
			P::make_ray(vw, u, v, aabb, pc, rng, ro, rd, rdn, idir, startt);
			live = true;
		}
	}
	// pass 1: count the occupied steps, keeping every sample's t (or the first 64 and a checkpoint every GEN_SEG samples beyond)
	const uint32_t walk_empty = (cfg.gen_walk_empty ? cfg.gen_walk_empty : gen_walk_empty_default(G)) - 1u;
	const uint32_t spec0 = cfg.gen_speculation ? (cfg.gen_speculation < G ? cfg.gen_speculation : G) : gen_speculation_default(G);
	const uint32_t numsteps = march_group<P, G>(pc, aabb, cfg.max_cascade, bitfield, ro, rdn, idir, startt, 0u, live ? NGP_NERF_STEPS : 0u, walk_empty, spec0,
		[&](uint32_t jj, float t, float dt, V3) {
			if (jj < T_SLOTS) t_list[jj * NG + grp] = t;
			if constexpr (G < 16) {
				// resuming the loop at this t finds sample jn first
				const uint32_t jn = jj + 1u;
				if (jn >= T_SLOTS && jn < NGP_NERF_STEPS && ((jn - T_SLOTS) & (GEN_SEG - 1u)) == 0u) ckpt[((jn - T_SLOTS) / GEN_SEG) * NG + grp] = t + dt;
			}
		});

	// ---- warp-level reservation of sample slots and ray slots (lane 0 of each group speaks for its ray)
	const uint32_t mine = g == 0 ? numsteps : 0u;
	uint32_t incl = mine;
#pragma unroll
	for (uint32_t o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
		if (lane >= o) incl += t;
	}
	const uint32_t warp_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
	uint32_t warp_base = 0;
	if (lane == 0 && warp_total > 0) warp_base = atomicAdd(&counters->n_samples, warp_total);
	warp_base = __shfl_sync(0xFFFFFFFFu, warp_base, 0);
	const uint32_t base = warp_base + __shfl_sync(0xFFFFFFFFu, incl - mine, lane0);
	const bool keep = numsteps > 0 && (base + numsteps <= max_samples);
	const uint32_t keep_mask = __ballot_sync(0xFFFFFFFFu, keep && g == 0);
	uint32_t ray_base = 0;
	if (lane == 0 && keep_mask) ray_base = atomicAdd(&counters->n_rays, __popc(keep_mask));
	ray_base = __shfl_sync(0xFFFFFFFFu, ray_base, 0);
	const uint32_t ray_idx = ray_base + __popc(keep_mask & ((1u << lane0) - 1u));
	const uint32_t n_write = keep ? numsteps : 0u;
	if (keep && g == 0) {
		ray_indices_out[ray_idx] = i;
		float* r = rays_out + (size_t)ray_idx * 6;
		r[0] = ro.x; r[1] = ro.y; r[2] = ro.z; r[3] = rd.x; r[4] = rd.y; r[5] = rd.z;
		numsteps_out[ray_idx * 2 + 0] = numsteps;
		numsteps_out[ray_idx * 2 + 1] = base;
	}

	// pass 2a: the warp writes the coordinates of the samples whose t is in shared memory, ray after ray
	__syncwarp();
	const uint32_t n_listed = n_write < T_SLOTS ? n_write : T_SLOTS;
#pragma unroll 1
	for (uint32_t sg = 0; sg < NG; ++sg) {
		const uint32_t src = sg * G;
		const uint32_t n_s = __shfl_sync(0xFFFFFFFFu, n_listed, src);
		if (n_s == 0) continue;  // warp-uniform
		const uint32_t base_s = __shfl_sync(0xFFFFFFFFu, base, src);
		const V3 ro_s{__shfl_sync(0xFFFFFFFFu, ro.x, src), __shfl_sync(0xFFFFFFFFu, ro.y, src), __shfl_sync(0xFFFFFFFFu, ro.z, src)};
		const V3 rdn_s{__shfl_sync(0xFFFFFFFFu, rdn.x, src), __shfl_sync(0xFFFFFFFFu, rdn.y, src), __shfl_sync(0xFFFFFFFFu, rdn.z, src)};
		const V3 wdir = P::warp_direction(rdn_s);
		// 32 records of 28 bytes = 224 consecutive floats: transposed through shared memory so that every store instruction of the
		// warp covers 128 contiguous bytes (per-lane 4-byte stores at a 28-byte stride hit 32 sectors each)
		for (uint32_t k0 = 0; k0 < n_s; k0 += 32) {
			const uint32_t k = k0 + lane;
			__syncwarp();
			if (k < n_s) {
				const float t = t_list[k * NG + sg];
				const V3 pos = P::ray_pos(ro_s, t, rdn_s);
				const float dt = P::calc_dt(t, pc);
				const V3 wp = P::warp_position(pos, aabb);
				float* c = coord_tile + lane * 7;
				c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = P::warp_dt(dt); c[4] = wdir.x; c[5] = wdir.y; c[6] = wdir.z;
			}
			__syncwarp();
			const uint32_t cnt = ((n_s - k0) < 32u ? (n_s - k0) : 32u) * 7u;
			float* dst = coords_out + (size_t)(base_s + k0) * 7;
#pragma unroll
			for (uint32_t q = 0; q < 7; ++q) {
				const uint32_t e = q * 32u + lane;
				if (e < cnt) dst[e] = coord_tile[e];
			}
		}
	}

	// pass 2b (G < 16): the tails (samples beyond the listed 64) of the warp's long rays, cut into GEN_SEG-sample segments; the groups
	// take the warp's segments round-robin and re-march each from its checkpoint
	if constexpr (G < 16) {
		const uint32_t n_tail = n_write > T_SLOTS ? n_write - T_SLOTS : 0u;
		const uint32_t nseg = g == 0 ? (n_tail + GEN_SEG - 1u) / GEN_SEG : 0u;
		uint32_t seg_incl = nseg;
#pragma unroll
		for (uint32_t o = 1; o < 32; o <<= 1) {
			const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, seg_incl, o);
			if (lane >= o) seg_incl += v;
		}
		const uint32_t total_seg = __shfl_sync(0xFFFFFFFFu, seg_incl, 31);
		for (uint32_t s0 = 0; s0 < total_seg; s0 += NG) {
			const uint32_t sidx = s0 + grp;
			const bool active = sidx < total_seg;
			const uint32_t sq = active ? sidx : total_seg - 1u;
			// owner = first lane whose inclusive segment count exceeds sq (only lanes g == 0 carry segments)
			uint32_t owner = 0;
#pragma unroll
			for (uint32_t step = 16; step >= 1; step >>= 1) {
				const uint32_t cand = owner + step;
				const uint32_t v = __shfl_sync(0xFFFFFFFFu, seg_incl, (cand - 1u) & 31u);
				if (cand <= 31u && v <= sq) owner = cand;
			}
			const uint32_t o_incl = __shfl_sync(0xFFFFFFFFu, seg_incl, owner), o_nseg = __shfl_sync(0xFFFFFFFFu, nseg, owner);
			const uint32_t m = sq - (o_incl - o_nseg);
			const uint32_t o_base = __shfl_sync(0xFFFFFFFFu, base, owner), o_n = __shfl_sync(0xFFFFFFFFu, n_write, owner);
			const V3 o_ro{__shfl_sync(0xFFFFFFFFu, ro.x, owner), __shfl_sync(0xFFFFFFFFu, ro.y, owner), __shfl_sync(0xFFFFFFFFu, ro.z, owner)};
			const V3 o_rdn{__shfl_sync(0xFFFFFFFFu, rdn.x, owner), __shfl_sync(0xFFFFFFFFu, rdn.y, owner), __shfl_sync(0xFFFFFFFFu, rdn.z, owner)};
			const V3 o_idir{__shfl_sync(0xFFFFFFFFu, idir.x, owner), __shfl_sync(0xFFFFFFFFu, idir.y, owner), __shfl_sync(0xFFFFFFFFu, idir.z, owner)};
			const V3 wdir = P::warp_direction(o_rdn);
			const float t0 = active ? ckpt[m * NG + owner / G] : 0.0f;
			const uint32_t j0 = T_SLOTS + m * GEN_SEG;
			const uint32_t j_end = (j0 + GEN_SEG < o_n) ? j0 + GEN_SEG : o_n;
			float* co = coords_out + (size_t)o_base * 7;
			march_group<P, G>(pc, aabb, cfg.max_cascade, bitfield, o_ro, o_rdn, o_idir, t0, active ? j0 : 1u, active ? j_end : 0u, walk_empty, spec0, [&](uint32_t jj, float t, float dt, V3 pos) {
				const V3 wp = P::warp_position(pos, aabb);
				float* c = co + (size_t)jj * 7;
				c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = P::warp_dt(dt); c[4] = wdir.x; c[5] = wdir.y; c[6] = wdir.z;
			});
		}
	}
}

// lanes per ray for a batch of n rays: as many as keep the launch within one wave of warps (16 resident warps per SM leave each
// warp a scheduler slot to itself most of the time; the kernel is latency bound, see the header)
inline uint32_t gen_lanes_per_ray(uint32_t n_rays, uint32_t sm_count) {
	const uint64_t lanes = (uint64_t)sm_count * 16u * 32u;
	uint32_t G = 32;
	while (G > 1 && (uint64_t)n_rays * G > lanes) G >>= 1;
	return G;
}

template <class P, uint32_t G>
static void launch_generate_training_samples_g(cudaStream_t stream, uint32_t n_rays_local, uint32_t ray_offset, uint32_t stride, uint32_t n_rays_global,
	uint64_t rng_state, uint64_t rng_inc, const ngp_nerf_train_cfg& cfg, const ngp_train_view* views, uint32_t n_views, const uint8_t* bitfield, uint32_t max_samples,
	ngp_nerf_counters* counters, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords) {
	k_generate_training_samples<P, G><<<div_round_up(n_rays_local, 32u / G), 32, gen_smem_bytes(G), stream>>>(n_rays_local, ray_offset, stride, n_rays_global,
		Pcg32(rng_state, rng_inc, true), cfg, views, n_views, bitfield, max_samples, counters, ray_indices, rays, numsteps, coords);
}

template <class P>
void launch_generate_training_samples(cudaStream_t stream, uint32_t n_rays_local, uint32_t ray_offset, uint32_t n_rays_global, uint64_t rng_state,
	uint64_t rng_inc, const ngp_nerf_train_cfg& cfg, const ngp_train_view* views, uint32_t n_views, const uint8_t* bitfield, uint32_t max_samples,
	ngp_nerf_counters* counters, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords) {
	if (n_rays_local == 0) return;
	NGPB_CHECK(n_views > 0, "generate_training_samples: no training views");
	NGPB_CHECK(coords != nullptr, "generate_training_samples: no coordinate buffer");
	const uint32_t stride = cfg.ray_stride ? cfg.ray_stride : 1u;
	static int sm_count = 0;
	if (!sm_count) {
		int dev = 0;
		NGPB_CUDA_CHECK(cudaGetDevice(&dev));
		NGPB_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
	}
	const uint32_t G = cfg.gen_lanes_per_ray ? cfg.gen_lanes_per_ray : gen_lanes_per_ray(n_rays_local, (uint32_t)sm_count);
#define NGPB_GEN_CASE(GG) \
	case GG: launch_generate_training_samples_g<P, GG>(stream, n_rays_local, ray_offset, stride, n_rays_global, rng_state, rng_inc, cfg, views, n_views, bitfield, \
		max_samples, counters, ray_indices, rays, numsteps, coords); break;
	switch (G) {
		NGPB_GEN_CASE(1) NGPB_GEN_CASE(2) NGPB_GEN_CASE(4) NGPB_GEN_CASE(8) NGPB_GEN_CASE(16) NGPB_GEN_CASE(32)
		default: NGPB_CHECK(false, "ngp_nerf_train_cfg.gen_lanes_per_ray must be 0 (automatic) or a power of two up to 32");
	}
#undef NGPB_GEN_CASE
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ngpb
