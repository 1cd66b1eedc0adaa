// march.cu — NeRF training-ray generation, occupancy-grid marching, compositing/loss/compaction and density-grid maintenance.
// Compiled with -fmad=false (see march.cuh).  Restates src/testbed_nerf.cu kernels, cited per kernel.
#include "gen_kernel.cuh"

namespace ngpb {

// ------------------------------------------------------------------------------------------------------------------
// generate_training_samples_nerf (testbed_nerf.cu:691-849): kernel body in gen_kernel.cuh, here with the deterministic arithmetic
// of march.cuh / ngp_detmath.h (this translation unit is compiled with -fmad=false), which the CPU oracle follows bit for bit.
// The reference build's own fast-math arithmetic is the second instantiation, march_ref.cu.
// ------------------------------------------------------------------------------------------------------------------
struct DetMarch {
	struct Ctx {
		ngp_march_consts m;
	};
	static __device__ __forceinline__ Ctx make_ctx(const ngp_nerf_train_cfg& cfg) { return Ctx{cfg.march}; }
	static __device__ __forceinline__ float calc_dt(float t, const Ctx& c) { return ngpb::calc_dt(t, c.m); }
	static __device__ __forceinline__ V3 ray_pos(V3 o, float t, V3 d) { return o + t * d; }
	static __device__ __forceinline__ uint32_t mip_from_dt(float dt, V3 pos, uint32_t max_cascade) { return ngpb::mip_from_dt(dt, pos, max_cascade); }
	static __device__ __forceinline__ float advance_to_next_voxel(float t, const Ctx& c, V3 pos, V3 dir, V3 idir, uint32_t mip) {
		return ngpb::advance_to_next_voxel(t, c.m, pos, dir, idir, mip);
	}
	static __device__ __forceinline__ V3 warp_position(V3 p, const Aabb& b) { return ngpb::warp_position(p, b); }
	static __device__ __forceinline__ V3 warp_direction(V3 d) { return ngpb::warp_direction(d); }
	static __device__ __forceinline__ float warp_dt(float dt) { return ngpb::warp_dt(dt); }
	static __device__ __forceinline__ void random_image_pos(Pcg32& rng, int w, int h, bool snap, float& u, float& v) { random_image_pos_training(rng, w, h, snap, u, v); }
	static __device__ __forceinline__ void make_ray(const ngp_train_view& vw, float u, float v, const Aabb& aabb, const Ctx& c, Pcg32& rng, V3& ro, V3& rd, V3& rdn,
		V3& idir, float& startt) {
		uv_to_ray(u, v, vw.width, vw.height, vw.focal_x, vw.focal_y, vw.principal_x, vw.principal_y, vw.lens_mode, vw.lens_params, vw.xform, ro, rd);
		rdn = normalize3(rd);
		float tmin, tmax;
		aabb_ray_intersect(aabb, ro, rdn, tmin, tmax);
		tmin = fmaxf(tmin, 0.0f);
		startt = advance_n_steps(tmin, c.m, rng.next_float());
		idir = V3{1.0f / rdn.x, 1.0f / rdn.y, 1.0f / rdn.z};
	}
};

void generate_training_samples_ref(cudaStream_t stream, uint32_t n_rays_local, uint32_t ray_offset, uint32_t n_rays_global, uint64_t rng_state,
	uint64_t rng_inc, const ngp_nerf_train_cfg& cfg, const ngp_train_view* views, uint32_t n_views, const uint8_t* bitfield, uint32_t max_samples,
	ngp_nerf_counters* counters, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords);   // march_ref.cu

// cfg.math_mode selects the flavour: NGP_MATH_DETERMINISTIC (0) or NGP_MATH_REFERENCE (1)
void generate_training_samples(cudaStream_t stream, uint32_t n_rays_local, uint32_t ray_offset, uint32_t n_rays_global, uint64_t rng_state,
	uint64_t rng_inc, const ngp_nerf_train_cfg& cfg, const ngp_train_view* views, uint32_t n_views, const uint8_t* bitfield, uint32_t max_samples,
	ngp_nerf_counters* counters, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords) {
	if (cfg.math_mode == NGP_MATH_REFERENCE)
		generate_training_samples_ref(stream, n_rays_local, ray_offset, n_rays_global, rng_state, rng_inc, cfg, views, n_views, bitfield, max_samples, counters,
			ray_indices, rays, numsteps, coords);
	else
		launch_generate_training_samples<DetMarch>(stream, n_rays_local, ray_offset, n_rays_global, rng_state, rng_inc, cfg, views, n_views, bitfield, max_samples,
			counters, ray_indices, rays, numsteps, coords);
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

// Which rays lose their gradients when a step's compacted samples exceed the batch (`compacted_base >= max_compacted`), and which
// are repeated when they fall short (fill_rollover), follows from the order in which rays take their compacted slots.  The reference
// reserves with one atomic per ray after its one-thread-per-ray loop (testbed_nerf.cu:1010): the 32 rays of a warp — consecutive ray
// ids, i.e. pixels of one training view — take consecutive slots, warps in the order they finish (early in training, when every ray
// runs to the step budget, that is launch order).  The order matters more than it looks: the controller rounds rays_per_batch UP to a
// multiple of 256 (NerfCounters::update_after_training), so most steps overshoot the batch by a few percent, and the first steps
// overshoot it tenfold.  One atomic per ray in a warp-per-ray kernel cuts rays one by one in the order they finish; on nerf/fox that
// leaves floaters in two held-out views in most runs (-0.2 dB in the mean, profiles/r2/psnr_ab.md) whatever the sample generator's
// schedule, while the reference's grouping does not.  So the order is computed instead of raced for: pass 1
// (k_compute_loss<LOSS_PASS1>) composites every ray and records its sample count under its local ray index, k_compaction_order hands
// out the slots along cfg.compaction_order, pass 2 (k_compute_loss<LOSS_PASS2>) writes gradients and compacted coordinates.
enum LossMode : uint32_t { LOSS_FUSED = 0, LOSS_PASS1 = 1, LOSS_PASS2 = 2 };
// One compositing step: returns the sample's weight and advances the transmittance.  The reference has two forms of the same quantity:
// its loss kernel keeps T as a running product (testbed_nerf.cu:946, 1095), its fused train_nerf kernel — the one the Rfl / RflRelax
// train modes run through — as one minus the accumulated weight (fused_kernels/train_nerf.cuh:228-230, 363-367).  They differ by rounding,
// which decides where `T < 1e-4` falls, so each train mode gets its kernel's form (`accumulated` is the weight sum of the second form).
__device__ __forceinline__ float composite_step(const bool weight_sum_form, const float alpha, float& T, float& accumulated) {
	const float weight = alpha * T;
	if (weight_sum_form) {
		accumulated += weight;
		T = 1.0f - accumulated;
	} else {
		T *= (1.0f - alpha);
	}
	return weight;
}
struct LossRayScratch {   // pass 1 -> pass 2, one per ray slot
	float rgb_ray[3];
	float loss_bg[3];
	uint32_t count;        // samples the ray is read for (before truncation)
	uint32_t pad;
};
constexpr uint32_t COMPACTION_GROUP = 32;   // rays that stay together, in ray order: a warp of the reference's kernel

// local ray index of the q-th ray in the step's compaction order
__device__ __forceinline__ uint32_t compaction_ray(uint32_t q, uint32_t n_groups, uint32_t order, uint32_t offset) {
	if (order == NGP_COMPACTION_RAY_ORDER) return q;
	// groups of 32 consecutive rays in shuffled order: an affine walk through [0, n_groups) with a prime stride larger than any count
	const uint32_t grp = (uint32_t)(((uint64_t)(q / COMPACTION_GROUP) * 2654435761ull + offset) % n_groups);
	return grp * COMPACTION_GROUP + (q % COMPACTION_GROUP);
}

// exclusive prefix sum of the rays' sample counts along the compaction order, in place (count -> first slot); one CTA of 1024 threads,
// a contiguous chunk of the order each
__global__ void __launch_bounds__(1024) k_compaction_order(const uint32_t n_rays_local, ngp_nerf_counters* __restrict__ counters, uint32_t* __restrict__ by_ray,
	const uint32_t order, const uint32_t offset) {
	const uint32_t n_groups = (n_rays_local + COMPACTION_GROUP - 1u) / COMPACTION_GROUP;
	const uint32_t n = n_groups * COMPACTION_GROUP;
	const uint32_t chunk = (n + 1023u) / 1024u;
	const uint32_t q0 = threadIdx.x * chunk < n ? threadIdx.x * chunk : n, q1 = (q0 + chunk < n) ? q0 + chunk : n;
	uint32_t sum = 0;
	for (uint32_t q = q0; q < q1; ++q) {
		const uint32_t li = compaction_ray(q, n_groups, order, offset);
		if (li < n_rays_local) sum += by_ray[li];
	}
	__shared__ uint32_t warp_sums[32];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint32_t incl = sum;
#pragma unroll
	for (uint32_t o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
		if (lane >= o) incl += t;
	}
	if (lane == 31) warp_sums[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		const uint32_t v = warp_sums[lane];
		uint32_t w = v;
#pragma unroll
		for (uint32_t o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, o);
			if (lane >= o) w += t;
		}
		warp_sums[lane] = w - v;
		if (lane == 31) counters->n_samples_compacted = w;   // what the reference's atomic counter ends at: not clamped
	}
	__syncthreads();
	uint32_t run = warp_sums[warp] + incl - sum;
	for (uint32_t q = q0; q < q1; ++q) {
		const uint32_t li = compaction_ray(q, n_groups, order, offset);
		if (li < n_rays_local) {
			const uint32_t c = by_ray[li];
			by_ray[li] = run;
			run += c;
		}
	}
}

template <uint32_t MODE>
__global__ void __launch_bounds__(128) k_compute_loss(
	const uint32_t n_rays_global, Pcg32 rng_in, const ngp_nerf_train_cfg cfg, const ngp_train_view* __restrict__ views, const uint32_t n_views,
	const __half* __restrict__ network_output, const uint32_t max_compacted, ngp_nerf_counters* __restrict__ counters,
	const uint32_t* __restrict__ ray_indices_in, const float* __restrict__ rays_in, uint32_t* __restrict__ numsteps_in,
	const float* __restrict__ coords_in, float* __restrict__ coords_out, __half* __restrict__ dloss_out, float* __restrict__ loss_output,
	const float* __restrict__ mean_density_ptr, LossRayScratch* __restrict__ scratch, uint32_t* __restrict__ by_ray, const uint32_t n_rays_local
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
	Rgba tex = read_rgba_uv(u, v, vw->width, vw->height, vw->pixels, vw->image_type);
	// per-image exposure (testbed_nerf.cu:979): the view's colour times 2^exposure, channel by channel
	V3 exposure_scale{1.0f, 1.0f, 1.0f};
	if (cfg.cam_exposure) {
		exposure_scale = V3{ngp_expf(0.6931471805599453f * cfg.cam_exposure[img * 3 + 0]), ngp_expf(0.6931471805599453f * cfg.cam_exposure[img * 3 + 1]),
			ngp_expf(0.6931471805599453f * cfg.cam_exposure[img * 3 + 2])};
		tex.r = exposure_scale.x * tex.r;
		tex.g = exposure_scale.y * tex.g;
		tex.b = exposure_scale.z * tex.b;
	}
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
	// ---- pass 1: composite front to back until T < EPSILON (testbed_nerf.cu:926-948; train_nerf.cuh:176-238 for the Rfl modes)
	const bool rtc = cfg.train_mode != NGP_TRAIN_NERF;   // the arithmetic of the reference's fused train_nerf kernel
	float T = 1.0f, acc = 0.0f;
	V3 rgb_ray{0, 0, 0}, loss_bg{0, 0, 0};
	uint32_t compacted_numsteps = 0;
	bool stopped = false;
	if constexpr (MODE == LOSS_PASS2) {
		const LossRayScratch r = scratch[i];
		rgb_ray = V3{r.rgb_ray[0], r.rgb_ray[1], r.rgb_ray[2]};
		loss_bg = V3{r.loss_bg[0], r.loss_bg[1], r.loss_bg[2]};
		compacted_numsteps = r.count;
	}
	for (uint32_t c0 = 0; MODE != LOSS_PASS2 && c0 < numsteps && !stopped; c0 += 32) {
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
			const float weight = composite_step(rtc, a, T, acc);
			rgb_ray = rgb_ray + weight * V3{r, g, b};
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

	// the background shows through: when every sample was read (loss kernel, :950) / when the ray is not opaque yet (train_nerf.cuh:251)
	if (MODE != LOSS_PASS2 && (rtc ? !(T < EPSILON) : compacted_numsteps == numsteps)) {
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

	if constexpr (MODE == LOSS_PASS1) {
		if (lane == 0) {
			LossRayScratch r;
			r.rgb_ray[0] = rgb_ray.x; r.rgb_ray[1] = rgb_ray.y; r.rgb_ray[2] = rgb_ray.z;
			r.loss_bg[0] = loss_bg.x; r.loss_bg[1] = loss_bg.y; r.loss_bg[2] = loss_bg.z;
			r.count = compacted_numsteps;
			r.pad = 0;
			scratch[i] = r;
			by_ray[(ray_idx / (cfg.ray_stride ? cfg.ray_stride : 1u)) % n_rays_local] = compacted_numsteps;
		}
		return;
	}
	// ---- compaction: one reservation per ray (testbed_nerf.cu:1010-1016)
	uint32_t compacted_base = 0;
	if constexpr (MODE == LOSS_PASS2) {
		compacted_base = by_ray[(ray_idx / (cfg.ray_stride ? cfg.ray_stride : 1u)) % n_rays_local];
	} else {
		if (lane == 0) compacted_base = atomicAdd(&counters->n_samples_compacted, compacted_numsteps);
		compacted_base = __shfl_sync(0xFFFFFFFFu, compacted_base, 0);
	}
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
	const float output_l1_reg_density = (!rtc && *mean_density_ptr < min_optical_thickness()) ? 1e-4f : 0.0f;   // train_nerf.cuh:305 has it switched off

	// ---- the gradient with respect to the view's exposure (testbed_nerf.cu:1142-1155), once per ray that keeps samples
	if (cfg.cam_exposure_gradient && lane == 0) {
		V3 dloss_by_dgt{-lg_grad.x, -lg_grad.y, -lg_grad.z};   // "assume symmetric loss"; uv_pdf == 1
		if (!cfg.linear_colors) {
			dloss_by_dgt = V3{dloss_by_dgt.x / srgb_to_linear_derivative(target.x), dloss_by_dgt.y / srgb_to_linear_derivative(target.y),
				dloss_by_dgt.z / srgb_to_linear_derivative(target.z)};
		}
		float* eg = cfg.cam_exposure_gradient + (size_t)img * 3;
		atomicAdd(eg + 0, ((loss_scale * dloss_by_dgt.x) * exposure_scale.x) * 0.6931471805599453f);
		atomicAdd(eg + 1, ((loss_scale * dloss_by_dgt.y) * exposure_scale.y) * 0.6931471805599453f);
		atomicAdd(eg + 2, ((loss_scale * dloss_by_dgt.z) * exposure_scale.z) * 0.6931471805599453f);
	}

	// ---- pass 2: gradients and compaction (testbed_nerf.cu:1078-1140)
	float* co = coords_out + (size_t)compacted_base * 7;
	__half* dl = dloss_out + (size_t)compacted_base * 4;
	V3 rgb_ray2{0, 0, 0}, loss_bg2{0, 0, 0};
	T = 1.0f;
	acc = 0.0f;
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
			const float weight = composite_step(rtc, a, T, acc);
			rgb_ray2 = rgb_ray2 + weight * V3{r, g, b};
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

// a memory pool of the current device whose freed blocks stay allocated (release threshold: never)
static cudaMemPool_t scratch_pool() {
	static cudaMemPool_t pools[64] = {};
	int dev = 0;
	NGPB_CUDA_CHECK(cudaGetDevice(&dev));
	NGPB_CHECK(dev >= 0 && dev < 64, "device index out of range");
	if (!pools[dev]) {
		cudaMemPoolProps props{};
		props.allocType = cudaMemAllocationTypePinned;
		props.handleTypes = cudaMemHandleTypeNone;
		props.location.type = cudaMemLocationTypeDevice;
		props.location.id = dev;
		NGPB_CUDA_CHECK(cudaMemPoolCreate(&pools[dev], &props));
		uint64_t keep = UINT64_MAX;
		NGPB_CUDA_CHECK(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
	}
	return pools[dev];
}

void compute_loss(cudaStream_t stream, uint32_t n_rays_local, uint32_t n_rays_global, uint64_t rng_state, uint64_t rng_inc,
	const ngp_nerf_train_cfg& cfg, const ngp_train_view* views, uint32_t n_views, const __half* network_output, uint32_t max_compacted,
	ngp_nerf_counters* counters, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords, float* coords_compacted,
	__half* dloss, float* loss_per_ray, const float* mean_density) {
	if (n_rays_local == 0) return;
	// one warp per ray, 4 rays per CTA
	const Pcg32 rng(rng_state, rng_inc, true);
	NGPB_CHECK(cfg.compaction_order <= NGP_COMPACTION_RAY_ORDER, "ngp_nerf_train_cfg.compaction_order: 0 (groups of 32 rays, shuffled), 1 (one atomic per ray), 2 (ray order)");
	if (cfg.compaction_order == NGP_COMPACTION_ATOMIC) {
		k_compute_loss<LOSS_FUSED><<<div_round_up(n_rays_local, 4), 128, 0, stream>>>(n_rays_global, rng, cfg, views, n_views,
			network_output, max_compacted, counters, ray_indices, rays, numsteps, coords, coords_compacted, dloss, loss_per_ray, mean_density, nullptr, nullptr, n_rays_local);
		NGPB_LAUNCHED();
	} else {
		// stream-ordered scratch from a pool that keeps its memory (the default pool hands it back at every synchronisation)
		const size_t slot_bytes = sizeof(LossRayScratch) * (size_t)n_rays_local;
		uint8_t* mem = nullptr;
		NGPB_CUDA_CHECK(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&mem), slot_bytes + sizeof(uint32_t) * (size_t)n_rays_local, scratch_pool(), stream));
		LossRayScratch* scratch = reinterpret_cast<LossRayScratch*>(mem);
		uint32_t* by_ray = reinterpret_cast<uint32_t*>(mem + slot_bytes);
		NGPB_CUDA_CHECK(cudaMemsetAsync(by_ray, 0, sizeof(uint32_t) * (size_t)n_rays_local, stream));
		k_compute_loss<LOSS_PASS1><<<div_round_up(n_rays_local, 4), 128, 0, stream>>>(n_rays_global, rng, cfg, views, n_views,
			network_output, max_compacted, counters, ray_indices, rays, numsteps, coords, coords_compacted, dloss, loss_per_ray, mean_density, scratch, by_ray, n_rays_local);
		// the shuffle's offset changes with the step's random stream (the multiplier is fixed)
		k_compaction_order<<<1, 1024, 0, stream>>>(n_rays_local, counters, by_ray, cfg.compaction_order, (uint32_t)((rng_state >> 33) ^ rng_state));
		k_compute_loss<LOSS_PASS2><<<div_round_up(n_rays_local, 4), 128, 0, stream>>>(n_rays_global, rng, cfg, views, n_views,
			network_output, max_compacted, counters, ray_indices, rays, numsteps, coords, coords_compacted, dloss, loss_per_ray, mean_density, scratch, by_ray, n_rays_local);
		NGPB_LAUNCHED(); NGPB_LAUNCHED(); NGPB_LAUNCHED();
		NGPB_CUDA_CHECK(cudaFreeAsync(mem, stream));
	}
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
