// march.cu — NeRF training-ray generation, occupancy-grid marching, compositing/loss/compaction and density-grid maintenance.
// Compiled with -fmad=false (see march.cuh).  Restates src/testbed_nerf.cu kernels, cited per kernel.
#include "march.cuh"

namespace ngpb {

// ------------------------------------------------------------------------------------------------------------------
// image access (common_device.cuh:776-872)
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

// ------------------------------------------------------------------------------------------------------------------
// generate_training_samples_nerf (testbed_nerf.cu:691-849)
// One thread per ray.  Slots are reserved once per warp (prefix sum + one atomic) instead of two atomics per ray.
// ray ids are GLOBAL: ray_id = ray_offset + local index, so that W ranks reproduce the single-GPU batch.
// ------------------------------------------------------------------------------------------------------------------
// WRITE_ALL = false: the kernel counts every ray in full (numsteps and the slot layout are those of the reference) but writes
// the coordinates of only the first `prefix` samples of each ray and the t at which the march would continue; the consumer
// (k_nerf_forward_rays, render.cu) marches on from there for the few rays whose transmittance is still above 1e-4 after the
// prefix — the loss kernel reads ~6 % of the 4 M samples a step reserves on a trained scene (profiles/r1b).
//
// Pass 1 leaves the t of each of a ray's first GEN_T_SLOTS samples in shared memory; the coordinate pass then does not march
// at all for those: the warp walks its 32 rays one after the other and its lanes turn consecutive samples of the same ray
// into coordinates (pos = o + t d, dt = calc_dt(t): the values pass 2 of the reference recomputes, testbed_nerf.cu:822-848),
// so the work is balanced across lanes whatever the rays' lengths and the 28-byte records leave the warp contiguously.
// Beyond GEN_T_SLOTS, pass 1 keeps one checkpoint of its loop state every GEN_SEG samples; the tails of the long rays of a
// warp are cut into GEN_SEG-sample segments that the lanes re-march in parallel, each from its checkpoint (a 300-sample ray
// costs 32 sequential steps instead of 220; ncu r1c: the one-thread tail re-march was half of the kernel's instructions).
constexpr uint32_t GEN_THREADS = 128;
constexpr uint32_t GEN_T_SLOTS = 64;    // 64 x 128 x 4 B = 32 KB of shared memory per CTA
constexpr uint32_t GEN_SEG = 32;
constexpr uint32_t GEN_N_CKPT = (NGP_NERF_STEPS - GEN_T_SLOTS + GEN_SEG - 1) / GEN_SEG;   // 30 x 128 x 4 B = 15 KB
// 47 KB dynamic + 3.5 KB static (coordinate tile) + 1 KB reserved = 4 CTAs per SM = 592 resident CTAs = 75 K rays in one wave.
// (At 3 CTAs per SM a steady-state batch of 57-66 K rays spilled into a second wave: 0.53 -> 0.80 ms, bimodal from run to run.)
constexpr uint32_t GEN_SMEM_BYTES = (GEN_T_SLOTS + GEN_N_CKPT) * GEN_THREADS * sizeof(float);
template <bool WRITE_ALL>
__global__ void __launch_bounds__(GEN_THREADS) k_generate_training_samples(
	const uint32_t n_rays_local, const uint32_t ray_offset, const uint32_t n_rays_global, Pcg32 rng_in, const ngp_nerf_train_cfg cfg,
	const ngp_train_view* __restrict__ views, const uint32_t n_views, const uint8_t* __restrict__ bitfield, const uint32_t max_samples,
	ngp_nerf_counters* __restrict__ counters, uint32_t* __restrict__ ray_indices_out, float* __restrict__ rays_out,
	uint32_t* __restrict__ numsteps_out, float* __restrict__ coords_out, float* __restrict__ t_resume_out, const uint32_t prefix,
	const uint32_t* __restrict__ perm
) {
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31u;
	const bool in_range = slot < n_rays_local;
	// which ray of the batch this thread marches: identity, or the batch ordered by expected march length (k_ray_sort_*)
	const uint32_t li = (perm && in_range) ? perm[slot] : slot;
	const uint32_t i = ray_offset + li;

	const Aabb aabb{V3{cfg.aabb_min[0], cfg.aabb_min[1], cfg.aabb_min[2]}, V3{cfg.aabb_max[0], cfg.aabb_max[1], cfg.aabb_max[2]}};
	extern __shared__ float t_list[];   // [GEN_T_SLOTS][GEN_THREADS], then checkpoints [GEN_N_CKPT][GEN_THREADS]
	float* ckpt = t_list + GEN_T_SLOTS * GEN_THREADS;
	__shared__ float coord_tile[GEN_THREADS / 32][32 * 7];
	uint32_t numsteps = 0;
	V3 ro{0, 0, 0}, rd{0, 0, 0}, rdn{0, 0, 1}, idir{0, 0, 0};
	float startt = 0.0f;

	if (in_range) {
		const uint32_t img = image_idx(i, n_rays_global, n_views);
		const ngp_train_view vw = views[img];
		Pcg32 rng = rng_in;
		rng.advance((uint64_t)i * N_MAX_RANDOM_SAMPLES_PER_RAY);
		float u, v;
		random_image_pos_training(rng, vw.width, vw.height, cfg.snap_to_pixel_centers != 0, u, v);
		const bool masked = !vw.no_mask && read_rgba_uv(u, v, vw.width, vw.height, vw.pixels, vw.image_type).r < 0.0f;
		if (!masked) {
			(void)rng.next_float();  // motion-blur time (testbed_nerf.cu:740) — consumed, unused without rolling shutter
			uv_to_ray(u, v, vw.width, vw.height, vw.focal_x, vw.focal_y, vw.principal_x, vw.principal_y, vw.lens_mode, vw.lens_params, vw.xform, ro, rd);
			rdn = normalize3(rd);
			float tmin, tmax;
			aabb_ray_intersect(aabb, ro, rdn, tmin, tmax);
			tmin = fmaxf(tmin, 0.0f);
			startt = advance_n_steps(tmin, cfg.march, rng.next_float());
			idir = V3{1.0f / rdn.x, 1.0f / rdn.y, 1.0f / rdn.z};

			// pass 1: count the occupied steps, keeping the t of the first GEN_T_SLOTS samples and the march state right after them
			uint32_t j = 0;
			float t = startt;
			V3 pos;
			OccCache occ;
			while (aabb.contains(pos = ro + t * rdn) && j < NGP_NERF_STEPS) {
				const float dt = calc_dt(t, cfg.march);
				const uint32_t mip = mip_from_dt(dt, pos, cfg.max_cascade);
				if (density_grid_occupied_cached(pos, bitfield, mip, occ)) {
					if (j < GEN_T_SLOTS) t_list[j * GEN_THREADS + threadIdx.x] = t;
					++j;
					t += dt;
					if (j >= GEN_T_SLOTS && ((j - GEN_T_SLOTS) & (GEN_SEG - 1u)) == 0u) ckpt[((j - GEN_T_SLOTS) / GEN_SEG) * GEN_THREADS + threadIdx.x] = t;
				} else {
					t = advance_to_next_voxel(t, cfg.march, pos, rdn, idir, mip);
				}
			}
			numsteps = j;
		}
	}

	// ---- warp-level reservation of sample slots and ray slots
	uint32_t incl = numsteps;
#pragma unroll
	for (uint32_t o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
		if (lane >= o) incl += t;
	}
	const uint32_t warp_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
	uint32_t warp_base = 0;
	if (lane == 0 && warp_total > 0) warp_base = atomicAdd(&counters->n_samples, warp_total);
	warp_base = __shfl_sync(0xFFFFFFFFu, warp_base, 0);
	const uint32_t base = warp_base + incl - numsteps;
	const bool keep = numsteps > 0 && (base + numsteps <= max_samples);
	const uint32_t keep_mask = __ballot_sync(0xFFFFFFFFu, keep);
	uint32_t ray_base = 0;
	if (lane == 0 && keep_mask) ray_base = atomicAdd(&counters->n_rays, __popc(keep_mask));
	ray_base = __shfl_sync(0xFFFFFFFFu, ray_base, 0);
	const uint32_t ray_idx = ray_base + __popc(keep_mask & ((1u << lane) - 1u));
	const uint32_t n_write = !keep ? 0u : (WRITE_ALL ? numsteps : (numsteps < prefix ? numsteps : prefix));
	if (keep) {
		ray_indices_out[ray_idx] = i;
		float* r = rays_out + (size_t)ray_idx * 6;
		r[0] = ro.x; r[1] = ro.y; r[2] = ro.z; r[3] = rd.x; r[4] = rd.y; r[5] = rd.z;
		numsteps_out[ray_idx * 2 + 0] = numsteps;
		numsteps_out[ray_idx * 2 + 1] = base;
	}

	// pass 2a: the warp writes the coordinates of the samples whose t is in shared memory, ray after ray
	__syncwarp();
	const uint32_t n_listed = n_write < GEN_T_SLOTS ? n_write : GEN_T_SLOTS;
	const uint32_t warp_col0 = threadIdx.x & ~31u;
	for (uint32_t src = 0; src < 32; ++src) {
		const uint32_t n_s = __shfl_sync(0xFFFFFFFFu, n_listed, src);
		if (n_s == 0) continue;  // warp-uniform
		const uint32_t base_s = __shfl_sync(0xFFFFFFFFu, base, src);
		const V3 ro_s{__shfl_sync(0xFFFFFFFFu, ro.x, src), __shfl_sync(0xFFFFFFFFu, ro.y, src), __shfl_sync(0xFFFFFFFFu, ro.z, src)};
		const V3 rdn_s{__shfl_sync(0xFFFFFFFFu, rdn.x, src), __shfl_sync(0xFFFFFFFFu, rdn.y, src), __shfl_sync(0xFFFFFFFFu, rdn.z, src)};
		const V3 wdir = warp_direction(rdn_s);
		// 32 records of 28 bytes = 224 consecutive floats: transposed through shared memory so that every store instruction of the
		// warp covers 128 contiguous bytes (per-lane 4-byte stores at a 28-byte stride hit 32 sectors each: the L2 write path,
		// not the march, was then what this pass waited for)
		float* tile = coord_tile[threadIdx.x >> 5];
		for (uint32_t k0 = 0; k0 < n_s; k0 += 32) {
			const uint32_t k = k0 + lane;
			__syncwarp();
			if (k < n_s) {
				const float t = t_list[k * GEN_THREADS + warp_col0 + src];
				const V3 pos = ro_s + t * rdn_s;
				const float dt = calc_dt(t, cfg.march);
				const V3 wp = warp_position(pos, aabb);
				float* c = tile + lane * 7;
				c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = warp_dt(dt); c[4] = wdir.x; c[5] = wdir.y; c[6] = wdir.z;
			}
			__syncwarp();
			const uint32_t cnt = ((n_s - k0) < 32u ? (n_s - k0) : 32u) * 7u;
			float* dst = coords_out + (size_t)(base_s + k0) * 7;
#pragma unroll
			for (uint32_t q = 0; q < 7; ++q) {
				const uint32_t e = q * 32u + lane;
				if (e < cnt) dst[e] = tile[e];
			}
		}
	}

	// pass 2b: the tails (samples beyond GEN_T_SLOTS) of the warp's long rays, cut into GEN_SEG-sample segments; lane l of round
	// r re-marches segment 32 r + l from its checkpoint
	{
		const uint32_t n_tail = n_write > GEN_T_SLOTS ? n_write - GEN_T_SLOTS : 0u;
		const uint32_t nseg = (n_tail + GEN_SEG - 1u) / GEN_SEG;
		uint32_t seg_incl = nseg;
#pragma unroll
		for (uint32_t o = 1; o < 32; o <<= 1) {
			const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, seg_incl, o);
			if (lane >= o) seg_incl += v;
		}
		const uint32_t total_seg = __shfl_sync(0xFFFFFFFFu, seg_incl, 31);
		for (uint32_t s0 = 0; s0 < total_seg; s0 += 32) {
			const uint32_t sidx = s0 + lane;
			const bool active = sidx < total_seg;
			const uint32_t sq = active ? sidx : total_seg - 1u;
			// owner = first lane whose inclusive segment count exceeds sq
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
			if (!active) continue;
			const V3 o_idir{1.0f / o_rdn.x, 1.0f / o_rdn.y, 1.0f / o_rdn.z};
			const V3 wdir = warp_direction(o_rdn);
			float t = ckpt[m * GEN_THREADS + warp_col0 + owner];
			uint32_t j = GEN_T_SLOTS + m * GEN_SEG;
			const uint32_t j_end = (j + GEN_SEG < o_n) ? j + GEN_SEG : o_n;
			float* co = coords_out + (size_t)o_base * 7;
			V3 pos;
			OccCache occ;
			while (aabb.contains(pos = o_ro + t * o_rdn) && j < j_end) {
				const float dt = calc_dt(t, cfg.march);
				const uint32_t mip = mip_from_dt(dt, pos, cfg.max_cascade);
				if (density_grid_occupied_cached(pos, bitfield, mip, occ)) {
					const V3 wp = warp_position(pos, aabb);
					float* c = co + (size_t)j * 7;
					c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = warp_dt(dt); c[4] = wdir.x; c[5] = wdir.y; c[6] = wdir.z;
					++j;
					t += dt;
				} else {
					t = advance_to_next_voxel(t, cfg.march, pos, o_rdn, o_idir, mip);
				}
			}
		}
	}
	if (!WRITE_ALL && keep) {
		// where the consumer resumes the march: the loop state right after the last written sample (its t plus its dt), or the
		// first sample itself when nothing was written.  prefix <= GEN_T_SLOTS (checked by the launcher), so it is in the list.
		float t = 0.0f;
		if (n_write < numsteps) {
			if (n_write == 0) {
				t = t_list[threadIdx.x];
			} else {
				const float tl = t_list[(n_write - 1) * GEN_THREADS + threadIdx.x];
				t = tl + calc_dt(tl, cfg.march);
			}
		}
		t_resume_out[ray_idx] = t;
	}
}

// ------------------------------------------------------------------------------------------------------------------
// Ordering the rays of a batch by expected march length.
// One thread marches one ray and the rays of a batch differ 4x in length (background rays skip through an empty cube, rays
// through the object take hundreds of samples): in batch order a warp keeps 14.7 of its 32 lanes busy (ncu r1c) and, with only
// 57 K rays for 300 K lanes, nothing hides that.  Which thread marches which ray is free (slot order is atomics-dependent in the
// reference as well), so the batch is bucketed by an estimate of each ray's loop trips: 24 probes along the ray against the
// occupancy bitfield coarsened to 16^3 (a coarse cell = 64 consecutive Morton-ordered bytes), occupied stretches counted in
// steps, empty ones in voxel skips.  Three small kernels (keys + histogram, scan, scatter) produce the permutation; longest first.
// The rays, their sample counts and their coordinates are unchanged — only their slots move.
// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t SORT_BUCKETS = 64;
constexpr float SORT_TRIPS_PER_BUCKET = 16.0f;
constexpr uint32_t SORT_PROBES = 24;

__device__ inline bool coarse_cell_occupied(V3 pos, const uint8_t* __restrict__ bitfield, uint32_t mip) {
	const float mip_scale = scalbnf(1.0f, -(int)mip);
	const float px = (pos.x - 0.5f) * mip_scale + 0.5f, py = (pos.y - 0.5f) * mip_scale + 0.5f, pz = (pos.z - 0.5f) * mip_scale + 0.5f;
	const int ix = (int)(px * 16.0f), iy = (int)(py * 16.0f), iz = (int)(pz * 16.0f);
	if (ix < 0 || ix >= 16 || iy < 0 || iy >= 16 || iz < 0 || iz >= 16) return false;
	const uint4* p = reinterpret_cast<const uint4*>(bitfield + (size_t)(GRID_N_CELLS / 8) * mip + (size_t)morton3d((uint32_t)ix, (uint32_t)iy, (uint32_t)iz) * 64u);
	uint32_t any = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const uint4 v = __ldg(p + k);
		any |= v.x | v.y | v.z | v.w;
	}
	return any != 0;
}

__global__ void __launch_bounds__(GEN_THREADS) k_ray_sort_keys(
	const uint32_t n_rays_local, const uint32_t ray_offset, const uint32_t n_rays_global, Pcg32 rng_in, const ngp_nerf_train_cfg cfg,
	const ngp_train_view* __restrict__ views, const uint32_t n_views, const uint8_t* __restrict__ bitfield, uint8_t* __restrict__ keys, uint32_t* __restrict__ hist
) {
	__shared__ uint32_t local_hist[SORT_BUCKETS];
	if (threadIdx.x < SORT_BUCKETS) local_hist[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
	if (li < n_rays_local) {
		const uint32_t i = ray_offset + li;
		const Aabb aabb{V3{cfg.aabb_min[0], cfg.aabb_min[1], cfg.aabb_min[2]}, V3{cfg.aabb_max[0], cfg.aabb_max[1], cfg.aabb_max[2]}};
		const uint32_t img = image_idx(i, n_rays_global, n_views);
		const ngp_train_view vw = views[img];
		Pcg32 rng = rng_in;
		rng.advance((uint64_t)i * N_MAX_RANDOM_SAMPLES_PER_RAY);
		float u, v;
		random_image_pos_training(rng, vw.width, vw.height, cfg.snap_to_pixel_centers != 0, u, v);
		V3 ro, rd;
		uv_to_ray(u, v, vw.width, vw.height, vw.focal_x, vw.focal_y, vw.principal_x, vw.principal_y, vw.lens_mode, vw.lens_params, vw.xform, ro, rd);
		const V3 rdn = normalize3(rd);
		float tmin, tmax;
		aabb_ray_intersect(aabb, ro, rdn, tmin, tmax);
		tmin = fmaxf(tmin, 0.0f);
		float trips = 0.0f;
		if (tmax > tmin) {
			const float seg = (tmax - tmin) / (float)SORT_PROBES;
			for (uint32_t k = 0; k < SORT_PROBES; ++k) {
				const float t = tmin + ((float)k + 0.5f) * seg;
				const V3 pos = ro + t * rdn;
				const uint32_t mip = mip_from_pos(pos, cfg.max_cascade);
				if (coarse_cell_occupied(pos, bitfield, mip)) trips += seg / calc_dt(t, cfg.march);
				else trips += seg * 128.0f * scalbnf(1.0f, -(int)mip);   // one trip per cell of that cascade
			}
		}
		const uint32_t key = (uint32_t)fminf(trips / SORT_TRIPS_PER_BUCKET, (float)(SORT_BUCKETS - 1));
		keys[li] = (uint8_t)key;
		atomicAdd(&local_hist[key], 1u);
	}
	__syncthreads();
	if (threadIdx.x < SORT_BUCKETS && local_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], local_hist[threadIdx.x]);
}

// state: hist[64] | base[64] | cursor[64]; longest bucket first
__global__ void k_ray_sort_scan(uint32_t* __restrict__ state) {
	if (threadIdx.x != 0) return;
	uint32_t acc = 0;
	for (int b = (int)SORT_BUCKETS - 1; b >= 0; --b) {
		state[SORT_BUCKETS + b] = acc;
		acc += state[b];
		state[2 * SORT_BUCKETS + b] = 0;
	}
}

__global__ void __launch_bounds__(GEN_THREADS) k_ray_sort_scatter(const uint32_t n_rays_local, const uint8_t* __restrict__ keys, uint32_t* __restrict__ state,
	uint32_t* __restrict__ perm) {
	const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
	if (li >= n_rays_local) return;
	const uint32_t key = keys[li];
	const uint32_t slot = state[SORT_BUCKETS + key] + atomicAdd(&state[2 * SORT_BUCKETS + key], 1u);
	perm[slot] = li;
}

size_t ray_sort_scratch_bytes(uint32_t max_rays) { return (size_t)max_rays * 4 + next_multiple(max_rays, 16u) + 3 * SORT_BUCKETS * 4; }

// perm_out: [n_rays_local] u32 (also the start of `scratch`): scratch = perm | keys | state
const uint32_t* sort_training_rays(cudaStream_t stream, uint32_t n_rays_local, uint32_t ray_offset, uint32_t n_rays_global, uint64_t rng_state, uint64_t rng_inc,
	const ngp_nerf_train_cfg& cfg, const ngp_train_view* views, uint32_t n_views, const uint8_t* bitfield, void* scratch, uint32_t max_rays) {
	if (n_rays_local == 0) return nullptr;
	uint32_t* perm = reinterpret_cast<uint32_t*>(scratch);
	uint8_t* keys = reinterpret_cast<uint8_t*>(scratch) + (size_t)max_rays * 4;
	uint32_t* state = reinterpret_cast<uint32_t*>(keys + next_multiple(max_rays, 16u));
	NGPB_CUDA_CHECK(cudaMemsetAsync(state, 0, SORT_BUCKETS * 4, stream));
	const uint32_t blocks = div_round_up(n_rays_local, GEN_THREADS);
	k_ray_sort_keys<<<blocks, GEN_THREADS, 0, stream>>>(n_rays_local, ray_offset, n_rays_global, Pcg32(rng_state, rng_inc, true), cfg, views, n_views, bitfield, keys, state);
	k_ray_sort_scan<<<1, 32, 0, stream>>>(state);
	k_ray_sort_scatter<<<blocks, GEN_THREADS, 0, stream>>>(n_rays_local, keys, state, perm);
	NGPB_LAUNCHED(); NGPB_LAUNCHED(); NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
	return perm;
}

// ------------------------------------------------------------------------------------------------------------------
// The generator as two kernels, for the training pipeline (testbed.cu):
//   k_count_training_samples   ray generation + the counting march + slot reservation.  One thread per ray; its time is the
//                              serial latency of the batch's longest ray (~0.4 ms), not throughput — it needs no shared memory
//                              and 64 registers, so it can run beside the previous step's backward kernel and optimizer.  Every
//                              GEN_SEG-th sample's t goes to a global checkpoint table.
//   k_write_training_samples   one WARP per ray, lane m re-marches samples [32 m, 32 m + 32) from checkpoint m with the same
//                              arithmetic: all of a step's coordinates in ~32 sequential march steps.
// Outputs are those of k_generate_training_samples (same per-ray counts, records and coordinates; slot order differs).
// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t GEN_MAX_SEG = NGP_NERF_STEPS / GEN_SEG;   // 32 checkpoints per ray

__global__ void __launch_bounds__(GEN_THREADS) k_count_training_samples(
	const uint32_t n_rays_local, const uint32_t ray_offset, const uint32_t n_rays_global, Pcg32 rng_in, const ngp_nerf_train_cfg cfg,
	const ngp_train_view* __restrict__ views, const uint32_t n_views, const uint8_t* __restrict__ bitfield, const uint32_t max_samples,
	ngp_nerf_counters* __restrict__ counters, uint32_t* __restrict__ ray_indices_out, float* __restrict__ rays_out,
	uint32_t* __restrict__ numsteps_out, float* __restrict__ ckpt_out, uint32_t* __restrict__ seg_info_out
) {
	const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31u;
	const bool in_range = li < n_rays_local;
	const uint32_t i = ray_offset + li;
	const Aabb aabb{V3{cfg.aabb_min[0], cfg.aabb_min[1], cfg.aabb_min[2]}, V3{cfg.aabb_max[0], cfg.aabb_max[1], cfg.aabb_max[2]}};
	uint32_t numsteps = 0;
	V3 ro{0, 0, 0}, rd{0, 0, 0};

	if (in_range) {
		const uint32_t img = image_idx(i, n_rays_global, n_views);
		const ngp_train_view vw = views[img];
		Pcg32 rng = rng_in;
		rng.advance((uint64_t)i * N_MAX_RANDOM_SAMPLES_PER_RAY);
		float u, v;
		random_image_pos_training(rng, vw.width, vw.height, cfg.snap_to_pixel_centers != 0, u, v);
		const bool masked = !vw.no_mask && read_rgba_uv(u, v, vw.width, vw.height, vw.pixels, vw.image_type).r < 0.0f;
		if (!masked) {
			(void)rng.next_float();  // motion-blur time (testbed_nerf.cu:740)
			uv_to_ray(u, v, vw.width, vw.height, vw.focal_x, vw.focal_y, vw.principal_x, vw.principal_y, vw.lens_mode, vw.lens_params, vw.xform, ro, rd);
			const V3 rdn = normalize3(rd);
			float tmin, tmax;
			aabb_ray_intersect(aabb, ro, rdn, tmin, tmax);
			tmin = fmaxf(tmin, 0.0f);
			const float startt = advance_n_steps(tmin, cfg.march, rng.next_float());
			const V3 idir{1.0f / rdn.x, 1.0f / rdn.y, 1.0f / rdn.z};
			float* ck = ckpt_out + (size_t)li * GEN_MAX_SEG;
			uint32_t j = 0;
			float t = startt;
			V3 pos;
			OccCache occ;
			while (aabb.contains(pos = ro + t * rdn) && j < NGP_NERF_STEPS) {
				const float dt = calc_dt(t, cfg.march);
				const uint32_t mip = mip_from_dt(dt, pos, cfg.max_cascade);
				if (density_grid_occupied_cached(pos, bitfield, mip, occ)) {
					if ((j & (GEN_SEG - 1u)) == 0u) ck[j / GEN_SEG] = t;   // resuming the loop at this t finds sample j first
					++j;
					t += dt;
				} else {
					t = advance_to_next_voxel(t, cfg.march, pos, rdn, idir, mip);
				}
			}
			numsteps = j;
		}
	}

	// ---- warp-level reservation of sample slots and ray slots (as k_generate_training_samples)
	uint32_t incl = numsteps;
#pragma unroll
	for (uint32_t o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
		if (lane >= o) incl += t;
	}
	const uint32_t warp_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
	uint32_t warp_base = 0;
	if (lane == 0 && warp_total > 0) warp_base = atomicAdd(&counters->n_samples, warp_total);
	warp_base = __shfl_sync(0xFFFFFFFFu, warp_base, 0);
	const uint32_t base = warp_base + incl - numsteps;
	const bool keep = numsteps > 0 && (base + numsteps <= max_samples);
	const uint32_t keep_mask = __ballot_sync(0xFFFFFFFFu, keep);
	uint32_t ray_base = 0;
	if (lane == 0 && keep_mask) ray_base = atomicAdd(&counters->n_rays, __popc(keep_mask));
	ray_base = __shfl_sync(0xFFFFFFFFu, ray_base, 0);
	const uint32_t ray_idx = ray_base + __popc(keep_mask & ((1u << lane) - 1u));
	if (in_range) {
		seg_info_out[(size_t)li * 3 + 0] = keep ? numsteps : 0u;
		seg_info_out[(size_t)li * 3 + 1] = base;
		seg_info_out[(size_t)li * 3 + 2] = ray_idx;
	}
	if (!keep) return;
	ray_indices_out[ray_idx] = i;
	float* r = rays_out + (size_t)ray_idx * 6;
	r[0] = ro.x; r[1] = ro.y; r[2] = ro.z; r[3] = rd.x; r[4] = rd.y; r[5] = rd.z;
	numsteps_out[ray_idx * 2 + 0] = numsteps;
	numsteps_out[ray_idx * 2 + 1] = base;
}

__global__ void __launch_bounds__(GEN_THREADS) k_write_training_samples(
	const uint32_t n_rays_local, const ngp_nerf_train_cfg cfg, const uint8_t* __restrict__ bitfield, const float* __restrict__ rays_in,
	const float* __restrict__ ckpt_in, const uint32_t* __restrict__ seg_info_in, float* __restrict__ coords_out
) {
	const uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per ray
	const uint32_t m = threadIdx.x & 31u;                               // one lane per 32-sample segment
	if (li >= n_rays_local) return;
	const uint32_t n = seg_info_in[(size_t)li * 3 + 0];
	if (m * GEN_SEG >= n) return;
	const uint32_t base = seg_info_in[(size_t)li * 3 + 1], ray_idx = seg_info_in[(size_t)li * 3 + 2];
	const Aabb aabb{V3{cfg.aabb_min[0], cfg.aabb_min[1], cfg.aabb_min[2]}, V3{cfg.aabb_max[0], cfg.aabb_max[1], cfg.aabb_max[2]}};
	const float* rp = rays_in + (size_t)ray_idx * 6;
	const V3 ro{rp[0], rp[1], rp[2]};
	const V3 rdn = normalize3(V3{rp[3], rp[4], rp[5]});
	const V3 idir{1.0f / rdn.x, 1.0f / rdn.y, 1.0f / rdn.z};
	const V3 wdir = warp_direction(rdn);
	float t = ckpt_in[(size_t)li * GEN_MAX_SEG + m];
	uint32_t j = m * GEN_SEG;
	const uint32_t j_end = (j + GEN_SEG < n) ? j + GEN_SEG : n;
	float* co = coords_out + (size_t)base * 7;
	V3 pos;
	OccCache occ;
	while (aabb.contains(pos = ro + t * rdn) && j < j_end) {
		const float dt = calc_dt(t, cfg.march);
		const uint32_t mip = mip_from_dt(dt, pos, cfg.max_cascade);
		if (density_grid_occupied_cached(pos, bitfield, mip, occ)) {
			const V3 wp = warp_position(pos, aabb);
			float* c = co + (size_t)j * 7;
			c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = warp_dt(dt); c[4] = wdir.x; c[5] = wdir.y; c[6] = wdir.z;
			++j;
			t += dt;
		} else {
			t = advance_to_next_voxel(t, cfg.march, pos, rdn, idir, mip);
		}
	}
}

// ------------------------------------------------------------------------------------------------------------------
// losses (nerf_device.cuh:75-143, 601-616)
// ------------------------------------------------------------------------------------------------------------------
__device__ inline void loss_and_gradient1(float target, float pred, uint32_t type, float& loss, float& grad) {
	const float d = pred - target;
	switch (type) {
		case NGP_LOSS_RELATIVE_L2: { const float den = pred * pred + 1e-2f; loss = d * d / den; grad = 2.0f * d / den; break; }
		case NGP_LOSS_L1: loss = fabsf(d); grad = copysignf(1.0f, d); break;
		case NGP_LOSS_MAPE: { const float den = fabsf(pred) + 1e-2f; loss = fabsf(d) / den; grad = copysignf(1.0f / den, d); break; }
		case NGP_LOSS_SMAPE: { const float den = 0.5f * (fabsf(pred) + fabsf(target)) + 1e-2f; loss = fabsf(d) / den; grad = copysignf(1.0f / den, d); break; }
		case NGP_LOSS_HUBER: {
			const float alpha = 0.1f;
			const float ad = fabsf(d);
			const float sq = 0.5f / alpha * d * d;
			loss = (ad > alpha ? (ad - 0.5f * alpha) : sq) / 5.0f;
			grad = (ad > alpha ? (d > 0.0f ? 1.0f : -1.0f) : (d / alpha)) / 5.0f;
			break;
		}
		case NGP_LOSS_LOGL1: { const float div = fabsf(d) + 1.0f; loss = ngp_logf(div); grad = copysignf(1.0f / div, d); break; }
		default: loss = d * d; grad = 2.0f * d; break;
	}
}

// ------------------------------------------------------------------------------------------------------------------
// compute_loss_kernel_train_nerf (testbed_nerf.cu:852-1180), Nerf train mode, no envmap / depth / exposure / error map.
// network_output: 4 halves per sample (rgb raw x3, density raw).
// ------------------------------------------------------------------------------------------------------------------
// One WARP per ray: the 32 lanes evaluate the activations / transmittance factors of 32 consecutive samples in parallel
// (all transcendentals live there) and the ray's running sums are then advanced in the reference's sequential order by
// every lane redundantly from warp shuffles, so every value is produced by exactly the arithmetic of the one-thread-per-ray
// reference kernel — and of the oracle — while 32x more threads are in flight and long rays no longer stall a warp.
struct SampleTerms {
	float r, g, b;   // network_to_rgb of the three colour channels
	float alpha;     // 1 - exp(-density * dt)
};
__device__ __forceinline__ SampleTerms sample_terms(const __half* __restrict__ no, const float* __restrict__ ci, uint32_t k, const ngp_nerf_train_cfg& cfg,
	float& o0, float& o1, float& o2, float& o3, float& dt) {
	const uint2 raw = *reinterpret_cast<const uint2*>(no + (size_t)k * 4);
	const __half2 h01 = *reinterpret_cast<const __half2*>(&raw.x), h23 = *reinterpret_cast<const __half2*>(&raw.y);
	o0 = __low2float(h01); o1 = __high2float(h01); o2 = __low2float(h23); o3 = __high2float(h23);
	dt = unwarp_dt(ci[(size_t)k * 7 + 3]);
	SampleTerms s;
	s.r = network_to_rgb(o0, cfg.rgb_activation);
	s.g = network_to_rgb(o1, cfg.rgb_activation);
	s.b = network_to_rgb(o2, cfg.rgb_activation);
	const float density = network_to_density(o3, cfg.density_activation);
	s.alpha = 1.0f - ngp_expf(-density * dt);
	return s;
}

__global__ void __launch_bounds__(128) k_compute_loss(
	const uint32_t n_rays_global, Pcg32 rng_in, const ngp_nerf_train_cfg cfg, const ngp_train_view* __restrict__ views, const uint32_t n_views,
	const __half* __restrict__ network_output, const uint32_t max_compacted, ngp_nerf_counters* __restrict__ counters,
	const uint32_t* __restrict__ ray_indices_in, const float* __restrict__ rays_in, uint32_t* __restrict__ numsteps_in,
	const float* __restrict__ coords_in, float* __restrict__ coords_out, __half* __restrict__ dloss_out, float* __restrict__ loss_output,
	const float* __restrict__ mean_density_ptr
) {
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // ray slot
	const uint32_t lane = threadIdx.x & 31u;
	if (i >= counters->n_rays) return;  // warp-uniform
	const Aabb aabb{V3{cfg.aabb_min[0], cfg.aabb_min[1], cfg.aabb_min[2]}, V3{cfg.aabb_max[0], cfg.aabb_max[1], cfg.aabb_max[2]}};
	const float EPSILON = 1e-4f;

	const uint32_t numsteps = numsteps_in[i * 2 + 0];
	const uint32_t base = numsteps_in[i * 2 + 1];
	const V3 ray_o = V3{rays_in[(size_t)i * 6 + 0], rays_in[(size_t)i * 6 + 1], rays_in[(size_t)i * 6 + 2]};
	const __half* no = network_output + (size_t)base * 4;
	const float* ci = coords_in + (size_t)base * 7;

	// ---- target colour: same draws as the generator (testbed_nerf.cu:951-1004); computed by every lane (identical values)
	const uint32_t ray_idx = ray_indices_in[i];
	Pcg32 rng = rng_in;
	rng.advance((uint64_t)ray_idx * N_MAX_RANDOM_SAMPLES_PER_RAY);
	const uint32_t img = image_idx(ray_idx, n_rays_global, n_views);
	const ngp_train_view* vw = &views[img];
	float u, v;
	random_image_pos_training(rng, vw->width, vw->height, cfg.snap_to_pixel_centers != 0, u, v);
	rng.advance(1);  // motion-blur time
	V3 bg{cfg.background_color[0], cfg.background_color[1], cfg.background_color[2]};
	if (cfg.random_bg_color) {
		bg.x = rng.next_float();
		bg.y = rng.next_float();
		bg.z = rng.next_float();
	}
	bg = V3{srgb_to_linear(bg.x), srgb_to_linear(bg.y), srgb_to_linear(bg.z)};
	const Rgba tex = read_rgba_uv(u, v, vw->width, vw->height, vw->pixels, vw->image_type);
	V3 target;
	if (cfg.linear_colors || cfg.color_space == NGP_COLOR_LINEAR) {
		target = V3{tex.r + (1.0f - tex.a) * bg.x, tex.g + (1.0f - tex.a) * bg.y, tex.b + (1.0f - tex.a) * bg.z};
		if (!cfg.linear_colors) {
			target = V3{linear_to_srgb(target.x), linear_to_srgb(target.y), linear_to_srgb(target.z)};
			bg = V3{linear_to_srgb(bg.x), linear_to_srgb(bg.y), linear_to_srgb(bg.z)};
		}
	} else {
		bg = V3{linear_to_srgb(bg.x), linear_to_srgb(bg.y), linear_to_srgb(bg.z)};
		if (tex.a > 0.0f) {
			target = V3{linear_to_srgb(tex.r / tex.a) * tex.a + (1.0f - tex.a) * bg.x, linear_to_srgb(tex.g / tex.a) * tex.a + (1.0f - tex.a) * bg.y,
				linear_to_srgb(tex.b / tex.a) * tex.a + (1.0f - tex.a) * bg.z};
		} else {
			target = bg;
		}
	}
	// ---- pass 1: composite front to back until T < EPSILON (testbed_nerf.cu:926-948)
	float T = 1.0f;
	V3 rgb_ray{0, 0, 0}, loss_bg{0, 0, 0};
	uint32_t compacted_numsteps = 0;
	bool stopped = false;
	for (uint32_t c0 = 0; c0 < numsteps && !stopped; c0 += 32) {
		const uint32_t k = c0 + lane;
		SampleTerms st{0, 0, 0, 0};
		float o0, o1, o2, o3, dt;
		if (k < numsteps) st = sample_terms(no, ci, k, cfg, o0, o1, o2, o3, dt);
		const uint32_t n_here = (numsteps - c0) < 32u ? (numsteps - c0) : 32u;
		for (uint32_t j = 0; j < n_here; ++j) {
			if (T < EPSILON) {
				stopped = true;
				break;
			}
			const float a = __shfl_sync(0xFFFFFFFFu, st.alpha, j);
			const float r = __shfl_sync(0xFFFFFFFFu, st.r, j), g = __shfl_sync(0xFFFFFFFFu, st.g, j), b = __shfl_sync(0xFFFFFFFFu, st.b, j);
			const float weight = a * T;
			rgb_ray = rgb_ray + weight * V3{r, g, b};
			T *= (1.0f - a);
			++compacted_numsteps;
			if (cfg.train_mode == NGP_TRAIN_RFL) {
				// train_nerf.cuh:231: the ray's accumulated per-sample radiance-field loss
				float l0, l1, l2, gd;
				loss_and_gradient1(target.x, r, cfg.loss_type, l0, gd);
				loss_and_gradient1(target.y, g, cfg.loss_type, l1, gd);
				loss_and_gradient1(target.z, b, cfg.loss_type, l2, gd);
				loss_bg = loss_bg + weight * V3{l0, l1, l2};
			}
		}
	}

	if (compacted_numsteps == numsteps) {
		rgb_ray = rgb_ray + T * bg;
		if (cfg.train_mode == NGP_TRAIN_RFL) {   // train_nerf.cuh:251-254
			float l0, l1, l2, gd;
			loss_and_gradient1(target.x, bg.x, cfg.loss_type, l0, gd);
			loss_and_gradient1(target.y, bg.y, cfg.loss_type, l1, gd);
			loss_and_gradient1(target.z, bg.z, cfg.loss_type, l2, gd);
			loss_bg = loss_bg + T * V3{l0, l1, l2};
		}
	}
	V3 lg_grad;
	float lx, ly, lz;
	loss_and_gradient1(target.x, rgb_ray.x, cfg.loss_type, lx, lg_grad.x);
	loss_and_gradient1(target.y, rgb_ray.y, cfg.loss_type, ly, lg_grad.y);
	loss_and_gradient1(target.z, rgb_ray.z, cfg.loss_type, lz, lg_grad.z);
	const float mean_loss = ((lx + ly) + lz) / 3.0f;

	// ---- compaction: one reservation per ray (testbed_nerf.cu:1010-1016)
	uint32_t compacted_base = 0;
	if (lane == 0) compacted_base = atomicAdd(&counters->n_samples_compacted, compacted_numsteps);
	compacted_base = __shfl_sync(0xFFFFFFFFu, compacted_base, 0);
	const uint32_t cb_clamped = compacted_base < max_compacted ? compacted_base : max_compacted;
	const uint32_t room = max_compacted - cb_clamped;
	compacted_numsteps = room < compacted_numsteps ? room : compacted_numsteps;
	if (lane == 0) {
		numsteps_in[i * 2 + 0] = compacted_numsteps;
		numsteps_in[i * 2 + 1] = compacted_base;
	}
	if (compacted_numsteps == 0) return;
	if (loss_output && lane == 0) loss_output[i] = mean_loss / (float)n_rays_global;

	const float loss_scale = cfg.loss_scale / (float)n_rays_global;
	const float output_l2_reg = cfg.rgb_activation == NGP_ACT_EXPONENTIAL ? 1e-4f : 0.0f;
	const float output_l1_reg_density = *mean_density_ptr < min_optical_thickness() ? 1e-4f : 0.0f;

	// ---- pass 2: gradients and compaction (testbed_nerf.cu:1078-1140)
	float* co = coords_out + (size_t)compacted_base * 7;
	__half* dl = dloss_out + (size_t)compacted_base * 4;
	V3 rgb_ray2{0, 0, 0}, loss_bg2{0, 0, 0};
	T = 1.0f;
	for (uint32_t c0 = 0; c0 < compacted_numsteps; c0 += 32) {
		const uint32_t k = c0 + lane;
		const bool mine = k < compacted_numsteps;
		SampleTerms st{0, 0, 0, 0};
		float o0 = 0, o1 = 0, o2 = 0, o3 = 0, dt = 0;
		float c[7] = {0, 0, 0, 0, 0, 0, 0};
		if (mine) {
			st = sample_terms(no, ci, k, cfg, o0, o1, o2, o3, dt);
#pragma unroll
			for (int q = 0; q < 7; ++q) c[q] = ci[(size_t)k * 7 + q];
		}
		const uint32_t n_here = (compacted_numsteps - c0) < 32u ? (compacted_numsteps - c0) : 32u;
		// the compacted copy of these n_here coordinate records is one contiguous block: 128 contiguous bytes per store instruction
		// instead of 4-byte stores at a 28-byte stride
		for (uint32_t e = lane; e < n_here * 7u; e += 32) co[(size_t)c0 * 7 + e] = ci[(size_t)c0 * 7 + e];
		float my_weight = 0.0f, my_T = 0.0f;
		V3 my_rgb_ray2{0, 0, 0}, my_loss_bg2{0, 0, 0};
		// Rfl: this lane's per-channel loss and gradient at its own sample colour (train_nerf.cuh:393-397)
		V3 my_ll{0, 0, 0}, my_lgr{0, 0, 0};
		if (cfg.train_mode == NGP_TRAIN_RFL && mine) {
			loss_and_gradient1(target.x, st.r, cfg.loss_type, my_ll.x, my_lgr.x);
			loss_and_gradient1(target.y, st.g, cfg.loss_type, my_ll.y, my_lgr.y);
			loss_and_gradient1(target.z, st.b, cfg.loss_type, my_ll.z, my_lgr.z);
		}
		for (uint32_t j = 0; j < n_here; ++j) {
			const float a = __shfl_sync(0xFFFFFFFFu, st.alpha, j);
			const float r = __shfl_sync(0xFFFFFFFFu, st.r, j), g = __shfl_sync(0xFFFFFFFFu, st.g, j), b = __shfl_sync(0xFFFFFFFFu, st.b, j);
			const float weight = a * T;
			rgb_ray2 = rgb_ray2 + weight * V3{r, g, b};
			T *= (1.0f - a);
			if (cfg.train_mode == NGP_TRAIN_RFL) {
				const V3 ll{__shfl_sync(0xFFFFFFFFu, my_ll.x, j), __shfl_sync(0xFFFFFFFFu, my_ll.y, j), __shfl_sync(0xFFFFFFFFu, my_ll.z, j)};
				loss_bg2 = loss_bg2 + weight * ll;
			}
			if (lane == j) {
				my_weight = weight;
				my_T = T;
				my_rgb_ray2 = rgb_ray2;
				my_loss_bg2 = loss_bg2;
			}
		}
		if (mine) {
			const V3 pos = unwarp_position(V3{c[0], c[1], c[2]}, aabb);
			const float depth = length3(pos - ray_o);
			const V3 rgb = V3{st.r, st.g, st.b};
			const V3 suffix = rgb_ray - my_rgb_ray2;
			V3 dloss_by_drgb = my_weight * lg_grad;
			float dmlp_inner;   // the bracket that multiplies density_derivative * dt (train_nerf.cuh:391-410)
			if (cfg.train_mode == NGP_TRAIN_RFL) {
				dloss_by_drgb = my_weight * my_lgr;
				const V3 e = my_T * my_ll - (loss_bg - my_loss_bg2);
				dmlp_inner = (e.x + e.y) + e.z;
			} else if (cfg.train_mode == NGP_TRAIN_RFL_RELAX) {
				const float tden = fmaxf(1e-6f, my_T);
				const V3 rgb_bg{suffix.x / tden, suffix.y / tden, suffix.z / tden};
				const V3 rgb_lerp = (1.0f - st.alpha) * rgb_bg + st.alpha * rgb;
				V3 ll, lgr;
				loss_and_gradient1(target.x, rgb_lerp.x, cfg.loss_type, ll.x, lgr.x);
				loss_and_gradient1(target.y, rgb_lerp.y, cfg.loss_type, ll.y, lgr.y);
				loss_and_gradient1(target.z, rgb_lerp.z, cfg.loss_type, ll.z, lgr.z);
				dloss_by_drgb = my_weight * lgr;
				dmlp_inner = dot3(lgr, my_T * rgb - suffix);
			} else {
				dmlp_inner = dot3(lg_grad, my_T * rgb - suffix);
			}
			const float d0 = loss_scale * (dloss_by_drgb.x * network_to_rgb_derivative(o0, cfg.rgb_activation) + fmaxf(0.0f, output_l2_reg * o0));
			const float d1 = loss_scale * (dloss_by_drgb.y * network_to_rgb_derivative(o1, cfg.rgb_activation) + fmaxf(0.0f, output_l2_reg * o1));
			const float d2 = loss_scale * (dloss_by_drgb.z * network_to_rgb_derivative(o2, cfg.rgb_activation) + fmaxf(0.0f, output_l2_reg * o2));
			const float density_derivative = network_to_density_derivative(o3, cfg.density_activation);
			const float dloss_by_dmlp = density_derivative * (dt * dmlp_inner);
			const float d3 = loss_scale * dloss_by_dmlp + (o3 < 0.0f ? -output_l1_reg_density : 0.0f) + (o3 > -10.0f && depth < cfg.near_distance ? 1e-4f : 0.0f);
			const __half2 w01 = __floats2half2_rn(d0, d1), w23 = __floats2half2_rn(d2, d3);
			uint2 outv;
			outv.x = *reinterpret_cast<const uint32_t*>(&w01);
			outv.y = *reinterpret_cast<const uint32_t*>(&w23);
			*reinterpret_cast<uint2*>(dl + (size_t)k * 4) = outv;
		}
	}
}

// fill_rollover_and_rescale<half>(dloss, stride 4) + fill_rollover<float>(coords, stride 7)
// (common_device.h:1114-1135 as called at testbed_nerf.cu:3298-3303)
__global__ void k_fill_rollover(const uint32_t n_elements, const ngp_nerf_counters* __restrict__ counters, float* __restrict__ coords, __half* __restrict__ dloss) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per padded sample
	const uint32_t n_input = counters->n_samples_compacted < n_elements ? counters->n_samples_compacted : n_elements;
	if (i < n_input || i >= n_elements || n_input == 0) return;
	const uint32_t src = i % n_input;
#pragma unroll
	for (int k = 0; k < 7; ++k) coords[(size_t)i * 7 + k] = coords[(size_t)src * 7 + k];
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const float v = __half2float(dloss[(size_t)src * 4 + k]);
		dloss[(size_t)i * 4 + k] = __float2half_rn(v * (float)n_input / (float)n_elements);
	}
}

// ------------------------------------------------------------------------------------------------------------------
// density grid maintenance (testbed_nerf.cu:87-162, 216-284, 316-396, 2476-2633)
// ------------------------------------------------------------------------------------------------------------------
__device__ inline V3 pos_to_uv_dir(const float* xform, V3 pos, const ngp_train_view& vw, float& u, float& v) {
	// pos_to_uv (common_device.cuh:527-577) for perspective/OpenCV lenses, parallax 0
	const V3 origin = xform_col(xform, 3);
	V3 dir = pos - origin;
	// inverse(mat3(camera)) * dir  — general 3x3 inverse via adjugate
	const float a = xform[0], b = xform[3], c = xform[6], d = xform[1], e = xform[4], f = xform[7], g = xform[2], h = xform[5], k = xform[8];
	const float A = e * k - f * h, B = -(d * k - f * g), C = d * h - e * g;
	const float det = a * A + b * B + c * C;
	const float inv00 = A / det, inv01 = -(b * k - c * h) / det, inv02 = (b * f - c * e) / det;
	const float inv10 = B / det, inv11 = (a * k - c * g) / det, inv12 = -(a * f - c * d) / det;
	const float inv20 = C / det, inv21 = -(a * h - b * g) / det, inv22 = (a * e - b * d) / det;
	V3 l{(inv00 * dir.x + inv01 * dir.y) + inv02 * dir.z, (inv10 * dir.x + inv11 * dir.y) + inv12 * dir.z, (inv20 * dir.x + inv21 * dir.y) + inv22 * dir.z};
	l = V3{l.x / l.z, l.y / l.z, 1.0f};
	float du = 0.0f, dv = 0.0f;
	if (vw.lens_mode == NGP_LENS_OPENCV) opencv_lens_distortion_delta(vw.lens_params, l.x, l.y, &du, &dv);
	l.x += du;
	l.y += dv;
	u = l.x * vw.focal_x / (float)vw.width + vw.principal_x;
	v = l.y * vw.focal_y / (float)vw.height + vw.principal_y;
	return dir;
}

__global__ void k_mark_untrained_density_grid(const uint32_t n_elements, float* __restrict__ grid_out, const uint32_t n_views,
	const ngp_train_view* __restrict__ views, const bool clear_visible_voxels) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	const uint32_t level = i / GRID_N_CELLS, pos_idx = i % GRID_N_CELLS;
	const uint32_t x = morton_compact3(pos_idx >> 0), y = morton_compact3(pos_idx >> 1), z = morton_compact3(pos_idx >> 2);
	const float voxel_size = scalbnf(1.0f / 128.0f, (int)level);
	const float s = scalbnf(1.0f, (int)level);
	const V3 pos{((float)x / 128.0f - 0.5f) * s + 0.5f, ((float)y / 128.0f - 0.5f) * s + 0.5f, ((float)z / 128.0f - 0.5f) * s + 0.5f};
	uint32_t count = 0;
	for (uint32_t j = 0; j < n_views && count < 1; ++j) {
		const ngp_train_view vw = views[j];
		const V3 cam_o = xform_col(vw.xform, 3), cam_fwd = xform_col(vw.xform, 2);
		for (uint32_t k = 0; k < 8; ++k) {
			const V3 corner{pos.x + ((k & 1u) ? voxel_size : 0.0f), pos.y + ((k & 2u) ? voxel_size : 0.0f), pos.z + ((k & 4u) ? voxel_size : 0.0f)};
			const V3 dir = normalize3(corner - cam_o);
			if (dot3(dir, cam_fwd) < 1e-4f) continue;
			float u, v;
			pos_to_uv_dir(vw.xform, corner, vw, u, v);
			V3 ro, rd;
			uv_to_ray(u, v, vw.width, vw.height, vw.focal_x, vw.focal_y, vw.principal_x, vw.principal_y, vw.lens_mode, vw.lens_params, vw.xform, ro, rd);
			const V3 diff = normalize3(rd) - dir;
			if (length3(diff) < 1e-3f && u > 0.0f && v > 0.0f && u < 1.0f && v < 1.0f) {
				++count;
				break;
			}
		}
	}
	if (clear_visible_voxels || (grid_out[i] < 0.0f) != (count < 1)) grid_out[i] = (count >= 1) ? 0.0f : -1.0f;
}

__global__ void k_generate_grid_samples(const uint32_t n_elements, Pcg32 rng, const uint32_t step, const Aabb aabb, const float* __restrict__ grid_in,
	float* __restrict__ out_pos /* 4 floats per sample: xyz + dt */, uint32_t* __restrict__ indices, const uint32_t n_cascades, const float thresh) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	rng.advance((uint64_t)i * 4);
	const uint32_t level = (uint32_t)(rng.next_float() * (float)n_cascades) % n_cascades;
	uint32_t idx = 0;
	for (uint32_t j = 0; j < 10; ++j) {
		idx = ((i + step * n_elements) * 56924617u + j * 19349663u + 96925573u) % GRID_N_CELLS;
		idx += level * GRID_N_CELLS;
		if (grid_in[idx] > thresh) break;
	}
	const uint32_t pos_idx = idx % GRID_N_CELLS;
	const uint32_t x = morton_compact3(pos_idx >> 0), y = morton_compact3(pos_idx >> 1), z = morton_compact3(pos_idx >> 2);
	const float rx = rng.next_float(), ry = rng.next_float(), rz = rng.next_float();
	const float s = scalbnf(1.0f, (int)level);
	const V3 pos{(((float)x + rx) / 128.0f - 0.5f) * s + 0.5f, (((float)y + ry) / 128.0f - 0.5f) * s + 0.5f, (((float)z + rz) / 128.0f - 0.5f) * s + 0.5f};
	const V3 wp = warp_position(pos, aabb);
	out_pos[(size_t)i * 4 + 0] = wp.x;
	out_pos[(size_t)i * 4 + 1] = wp.y;
	out_pos[(size_t)i * 4 + 2] = wp.z;
	out_pos[(size_t)i * 4 + 3] = warp_dt(min_cone_stepsize());
	indices[i] = idx;
}

__global__ void k_splat_grid_samples(const uint32_t n_elements, const uint32_t* __restrict__ indices, const __half* __restrict__ density_raw,
	float* __restrict__ grid_out, const uint32_t density_activation) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	const float mlp = network_to_density(__half2float(density_raw[i]), density_activation);
	const float optical_thickness = mlp * min_cone_stepsize();
	atomicMax(reinterpret_cast<uint32_t*>(grid_out) + indices[i], __float_as_uint(optical_thickness));
}

__global__ void k_ema_grid_samples(const uint32_t n_elements, const float decay, float* __restrict__ grid_out, const float* __restrict__ grid_in) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	const float importance = grid_in[i];
	const float prev = grid_out[i];
	grid_out[i] = (prev < 0.0f) ? prev : fmaxf(prev * decay, importance);
}

// mean of max(v, 0) over the first cascade, in a fixed order so the CPU oracle can reproduce it:
// 1024 partial sums (strided) followed by a sequential sum of the partials.
__global__ void k_density_mean_partial(const float* __restrict__ grid, float* __restrict__ partial /*1024*/) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;  // 1024 threads
	float s = 0.0f;
	for (uint32_t i = t; i < GRID_N_CELLS; i += 1024) s += fmaxf(grid[i], 0.0f) / (float)GRID_N_CELLS;
	partial[t] = s;
}
__global__ void k_density_mean_final(const float* __restrict__ partial, float* __restrict__ mean) {
	float s = 0.0f;
	for (uint32_t i = 0; i < 1024; ++i) s += partial[i];
	*mean = s;
}

__global__ void k_grid_to_bitfield(const uint32_t n_elements, const uint32_t n_nonzero, const float* __restrict__ grid, uint8_t* __restrict__ bitfield,
	const float* __restrict__ mean_density) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	if (i >= n_nonzero) {
		bitfield[i] = 0;
		return;
	}
	const float thresh = fminf(min_optical_thickness(), *mean_density);
	uint8_t bits = 0;
#pragma unroll
	for (uint8_t j = 0; j < 8; ++j) bits |= grid[(size_t)i * 8 + j] > thresh ? ((uint8_t)1 << j) : 0;
	bitfield[i] = bits;
}

__global__ void k_bitfield_max_pool(const uint32_t n_elements, const uint8_t* __restrict__ prev_level, uint8_t* __restrict__ next_level) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	uint8_t bits = 0;
#pragma unroll
	for (uint8_t j = 0; j < 8; ++j) bits |= prev_level[(size_t)i * 8 + j] > 0 ? ((uint8_t)1 << j) : 0;
	const uint32_t x = morton_compact3(i >> 0) + 128 / 8, y = morton_compact3(i >> 1) + 128 / 8, z = morton_compact3(i >> 2) + 128 / 8;
	next_level[morton3d(x, y, z)] |= bits;
}

// ------------------------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------------------------
void nerf_density(const ngp_nerf_desc& d, cudaStream_t stream, uint32_t n, const float* positions, uint32_t pos_stride, const __half* params, __half* out);

static Aabb cfg_aabb(const ngp_nerf_train_cfg& cfg) {
	return Aabb{V3{cfg.aabb_min[0], cfg.aabb_min[1], cfg.aabb_min[2]}, V3{cfg.aabb_max[0], cfg.aabb_max[1], cfg.aabb_max[2]}};
}

void generate_training_samples(cudaStream_t stream, uint32_t n_rays_local, uint32_t ray_offset, uint32_t n_rays_global, uint64_t rng_state,
	uint64_t rng_inc, const ngp_nerf_train_cfg& cfg, const ngp_train_view* views, uint32_t n_views, const uint8_t* bitfield, uint32_t max_samples,
	ngp_nerf_counters* counters, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords, float* t_resume, uint32_t prefix, const uint32_t* perm) {
	if (n_rays_local == 0) return;
	NGPB_CHECK(n_views > 0, "generate_training_samples: no training views");
	NGPB_CHECK(coords != nullptr, "generate_training_samples: no coordinate buffer");
	static bool attr_set = false;
	if (!attr_set) {
		NGPB_CUDA_CHECK(cudaFuncSetAttribute(k_generate_training_samples<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEN_SMEM_BYTES));
		NGPB_CUDA_CHECK(cudaFuncSetAttribute(k_generate_training_samples<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEN_SMEM_BYTES));
		attr_set = true;
	}
	if (!t_resume) {
		k_generate_training_samples<true><<<div_round_up(n_rays_local, GEN_THREADS), GEN_THREADS, GEN_SMEM_BYTES, stream>>>(n_rays_local, ray_offset, n_rays_global,
			Pcg32(rng_state, rng_inc, true), cfg, views, n_views, bitfield, max_samples, counters, ray_indices, rays, numsteps, coords, nullptr, 0u, perm);
	} else {
		NGPB_CHECK(prefix % 8u == 0u && prefix <= GEN_T_SLOTS, "generate_training_samples: the eager prefix must be a multiple of 8, at most 64");
		k_generate_training_samples<false><<<div_round_up(n_rays_local, GEN_THREADS), GEN_THREADS, GEN_SMEM_BYTES, stream>>>(n_rays_local, ray_offset, n_rays_global,
			Pcg32(rng_state, rng_inc, true), cfg, views, n_views, bitfield, max_samples, counters, ray_indices, rays, numsteps, coords, t_resume, prefix, perm);
	}
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

size_t generator_scratch_floats(uint32_t max_rays) { return (size_t)max_rays * GEN_MAX_SEG; }   // checkpoint table
size_t generator_scratch_u32(uint32_t max_rays) { return (size_t)max_rays * 3; }                 // (count, base, ray slot) per ray

void count_training_samples(cudaStream_t stream, uint32_t n_rays_local, uint32_t ray_offset, uint32_t n_rays_global, uint64_t rng_state, uint64_t rng_inc,
	const ngp_nerf_train_cfg& cfg, const ngp_train_view* views, uint32_t n_views, const uint8_t* bitfield, uint32_t max_samples, ngp_nerf_counters* counters,
	uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* ckpt, uint32_t* seg_info) {
	if (n_rays_local == 0) return;
	NGPB_CHECK(n_views > 0, "generate_training_samples: no training views");
	k_count_training_samples<<<div_round_up(n_rays_local, GEN_THREADS), GEN_THREADS, 0, stream>>>(n_rays_local, ray_offset, n_rays_global,
		Pcg32(rng_state, rng_inc, true), cfg, views, n_views, bitfield, max_samples, counters, ray_indices, rays, numsteps, ckpt, seg_info);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

void write_training_samples(cudaStream_t stream, uint32_t n_rays_local, const ngp_nerf_train_cfg& cfg, const uint8_t* bitfield, const float* rays, const float* ckpt,
	const uint32_t* seg_info, float* coords) {
	if (n_rays_local == 0) return;
	k_write_training_samples<<<div_round_up(n_rays_local * 32u, GEN_THREADS), GEN_THREADS, 0, stream>>>(n_rays_local, cfg, bitfield, rays, ckpt, seg_info, coords);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

void compute_loss(cudaStream_t stream, uint32_t n_rays_local, uint32_t n_rays_global, uint64_t rng_state, uint64_t rng_inc,
	const ngp_nerf_train_cfg& cfg, const ngp_train_view* views, uint32_t n_views, const __half* network_output, uint32_t max_compacted,
	ngp_nerf_counters* counters, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords, float* coords_compacted,
	__half* dloss, float* loss_per_ray, const float* mean_density) {
	if (n_rays_local == 0) return;
	// one warp per ray, 4 rays per CTA
	k_compute_loss<<<div_round_up(n_rays_local, 4), 128, 0, stream>>>(n_rays_global, Pcg32(rng_state, rng_inc, true), cfg, views, n_views,
		network_output, max_compacted, counters, ray_indices, rays, numsteps, coords, coords_compacted, dloss, loss_per_ray, mean_density);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

void fill_rollover(cudaStream_t stream, uint32_t target_batch, const ngp_nerf_counters* counters, float* coords_compacted, __half* dloss) {
	k_fill_rollover<<<div_round_up(target_batch, 256), 256, 0, stream>>>(target_batch, counters, coords_compacted, dloss);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

void update_bitfield(cudaStream_t stream, uint32_t max_cascade, const float* density_grid, uint8_t* bitfield, float* mean_density, float* partial1024) {
	k_density_mean_partial<<<4, 256, 0, stream>>>(density_grid, partial1024);
	k_density_mean_final<<<1, 1, 0, stream>>>(partial1024, mean_density);
	const uint32_t n_bytes = GRID_N_CELLS / 8 * NGP_NERF_CASCADES;
	k_grid_to_bitfield<<<div_round_up(n_bytes, 256), 256, 0, stream>>>(n_bytes, GRID_N_CELLS / 8 * (max_cascade + 1), density_grid, bitfield, mean_density);
	NGPB_LAUNCHED(); NGPB_LAUNCHED(); NGPB_LAUNCHED();
	for (uint32_t level = 1; level < NGP_NERF_CASCADES; ++level) {
		k_bitfield_max_pool<<<div_round_up(GRID_N_CELLS / 64, 256), 256, 0, stream>>>(GRID_N_CELLS / 64, bitfield + (size_t)(level - 1) * GRID_N_CELLS / 8,
			bitfield + (size_t)level * GRID_N_CELLS / 8);
		NGPB_LAUNCHED();
	}
	NGPB_CUDA_CHECK(cudaGetLastError());
}

// scratch layout: positions [n_total x 4 f32] | indices [n_total u32] | density_tmp [n_elements f32] | mlp_out [n_total f16] | partial [1024 f32]
size_t density_grid_scratch_bytes(uint32_t max_cascade) {
	const size_t n_elements = (size_t)GRID_N_CELLS * (max_cascade + 1);
	const size_t n_total = n_elements;  // worst case: the first 256 steps sample every cell once
	return n_total * 16 + n_total * 4 + n_elements * 4 + next_multiple((uint32_t)(n_total * 2), 16) + 1024 * 4 + 256;
}

void update_density_grid(const ngp_nerf_desc& d, cudaStream_t stream, const ngp_nerf_train_cfg& cfg, const __half* params, uint64_t* grid_rng_state,
	uint64_t grid_rng_inc, uint32_t training_step, uint32_t ema_step, float decay, const ngp_train_view* views, uint32_t n_views, float* density_grid,
	uint8_t* bitfield, float* mean_density, void* scratch) {
	const uint32_t n_cascades = cfg.max_cascade + 1;
	const uint32_t n_elements = GRID_N_CELLS * n_cascades;
	uint32_t n_uniform, n_nonuniform;
	if (training_step < 256) {
		n_uniform = GRID_N_CELLS * n_cascades;
		n_nonuniform = 0;
	} else {
		n_uniform = GRID_N_CELLS / 4 * n_cascades;
		n_nonuniform = GRID_N_CELLS / 4 * n_cascades;
	}
	const uint32_t n_total = n_uniform + n_nonuniform;
	uint8_t* p = reinterpret_cast<uint8_t*>(scratch);
	float* positions = reinterpret_cast<float*>(p);
	p += (size_t)n_elements * 16;
	uint32_t* indices = reinterpret_cast<uint32_t*>(p);
	p += (size_t)n_elements * 4;
	float* density_tmp = reinterpret_cast<float*>(p);
	p += (size_t)n_elements * 4;
	__half* mlp_out = reinterpret_cast<__half*>(p);
	p += next_multiple(n_elements * 2, 16);
	float* partial = reinterpret_cast<float*>(p);

	if (training_step == 0) {
		k_mark_untrained_density_grid<<<div_round_up(n_elements, 128), 128, 0, stream>>>(n_elements, density_grid, n_views, views, true);
		NGPB_LAUNCHED();
	}
	NGPB_CUDA_CHECK(cudaMemsetAsync(density_tmp, 0, sizeof(float) * n_elements, stream));
	const Aabb aabb = cfg_aabb(cfg);
	Pcg32 rng(*grid_rng_state, grid_rng_inc, true);
	k_generate_grid_samples<<<div_round_up(n_uniform, 128), 128, 0, stream>>>(n_uniform, rng, ema_step, aabb, density_grid, positions, indices, n_cascades, -0.01f);
	NGPB_LAUNCHED();
	rng.advance();
	if (n_nonuniform > 0) {
		k_generate_grid_samples<<<div_round_up(n_nonuniform, 128), 128, 0, stream>>>(n_nonuniform, rng, ema_step, aabb, density_grid,
			positions + (size_t)n_uniform * 4, indices + n_uniform, n_cascades, min_optical_thickness());
		NGPB_LAUNCHED();
	}
	rng.advance();
	*grid_rng_state = rng.state;

	nerf_density(d, stream, n_total, positions, 4, params, mlp_out);
	k_splat_grid_samples<<<div_round_up(n_total, 256), 256, 0, stream>>>(n_total, indices, mlp_out, density_tmp, cfg.density_activation);
	k_ema_grid_samples<<<div_round_up(n_elements, 256), 256, 0, stream>>>(n_elements, decay, density_grid, density_tmp);
	NGPB_LAUNCHED(); NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
	update_bitfield(stream, cfg.max_cascade, density_grid, bitfield, mean_density, partial);
}

}  // namespace ngpb
