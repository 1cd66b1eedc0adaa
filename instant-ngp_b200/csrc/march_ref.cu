// march_ref.cu — the training sample generator (gen_kernel.cuh) in the REFERENCE BUILD's arithmetic.
//
// This translation unit is compiled with --use_fast_math, the flag the reference gives every one of its .cu files
// (CMakeLists.txt:88): division and square root become the approximate MUFU forms, logf / expf / sinf become __logf / __expf /
// __sinf, denormals flush to zero and a*b+c contracts to one fma.  The floating-point steps come from march_ref.cuh, which restates
// the reference's functions with the expression trees its vec / mat templates expand to, so that nvcc is handed the same trees under
// the same flags as in the reference build.  Per-ray sample counts are decided at voxel faces by the last bits of t: this flavour is
// what makes "sample counts equal to the reference's kernel" an equality instead of a 99.x % statement (tests/test_gpu_vs_reference_nerf.py;
// the deterministic flavour in march.cu is the one the CPU oracle can follow).
//
// Nothing else of the library is compiled with these flags.
#include "gen_kernel.cuh"
#include "march_ref.cuh"

namespace ngpb {

struct RefMarch {
	struct Ctx {
		float cone_angle;
	};
	static __device__ __forceinline__ Ctx make_ctx(const ngp_nerf_train_cfg& cfg) { return Ctx{cfg.march.cone_angle}; }
	static __device__ __forceinline__ float calc_dt(float t, const Ctx& c) { return refm::calc_dt(t, c.cone_angle); }
	// ray(t) = o + t * d (common.h Ray::operator()): component-wise multiply, then add
	static __device__ __forceinline__ V3 ray_pos(V3 o, float t, V3 d) { return V3{o.x + t * d.x, o.y + t * d.y, o.z + t * d.z}; }
	// integer / exponent arithmetic only (frexpf, comparisons, one exact multiply by 256): flag independent
	static __device__ __forceinline__ uint32_t mip_from_dt(float dt, V3 pos, uint32_t max_cascade) { return ngpb::mip_from_dt(dt, pos, max_cascade); }
	static __device__ __forceinline__ float advance_to_next_voxel(float t, const Ctx& c, V3 pos, V3 dir, V3 idir, uint32_t mip) {
		return refm::advance_to_next_voxel(t, c.cone_angle, pos, dir, idir, mip);
	}
	static __device__ __forceinline__ V3 warp_position(V3 p, const Aabb& b) { return refm::warp_position(p, b); }
	static __device__ __forceinline__ V3 warp_direction(V3 d) { return refm::warp_direction(d); }
	static __device__ __forceinline__ float warp_dt(float dt) { return refm::warp_dt(dt); }
	// nerf_random_image_pos_training (nerf_device.cuh:553-576): (vec2(clamp(ivec2(uv * vec2(res)), 0, res - 1)) + 0.5f) / vec2(res)
	static __device__ __forceinline__ void random_image_pos(Pcg32& rng, int w, int h, bool snap, float& u, float& v) {
		u = rng.next_float();
		v = rng.next_float();
		if (snap) {
			u = ((float)imin(imax((int)(u * (float)w), 0), w - 1) + 0.5f) / (float)w;
			v = ((float)imin(imax((int)(v * (float)h), 0), h - 1) + 0.5f) / (float)h;
		}
	}
	// testbed_nerf.cu:745-798 for a perspective / OpenCV lens without rolling shutter, per-pixel rays, distortion map or parallax
	static __device__ __forceinline__ void make_ray(const ngp_train_view& vw, float u, float v, const Aabb& aabb, const Ctx& c, Pcg32& rng, V3& ro, V3& rd, V3& rdn,
		V3& idir, float& startt) {
		// get_xform_given_rolling_shutter (common_device.cuh:670-674): pixel_t = 0 for a zero rolling shutter, but the rotation still goes
		// matrix -> quaternion -> slerp -> matrix; the origin is mix(a[3], b[3], 0) = a[3]
		float rot[9];
		refm::camera_slerp_rotation(vw.xform, 0.0f, rot);
		// uv_to_ray (common_device.cuh:413-490): dir = {(uv - screen_center) * resolution / focal_length, 1}
		float dx = (u - vw.principal_x) * (float)vw.width / vw.focal_x;
		float dy = (v - vw.principal_y) * (float)vw.height / vw.focal_y;
		if (vw.lens_mode == NGP_LENS_OPENCV) refm::iterative_opencv_lens_undistortion(vw.lens_params, &dx, &dy);
		rd = refm::mat3_mul(rot, V3{dx, dy, 1.0f});
		ro = V3{vw.xform[9], vw.xform[10], vw.xform[11]};
		rdn = refm::normalize(rd);
		float tmin, tmax;
		refm::ray_intersect(aabb, ro, rdn, tmin, tmax);
		tmin = fmaxf(tmin, 0.0f);
		startt = refm::advance_n_steps(tmin, c.cone_angle, rng.next_float());
		idir = V3{1.0f / rdn.x, 1.0f / rdn.y, 1.0f / rdn.z};
	}
};

void generate_training_samples_ref(cudaStream_t stream, uint32_t n_rays_local, uint32_t ray_offset, uint32_t n_rays_global, uint64_t rng_state,
	uint64_t rng_inc, const ngp_nerf_train_cfg& cfg, const ngp_train_view* views, uint32_t n_views, const uint8_t* bitfield, uint32_t max_samples,
	ngp_nerf_counters* counters, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords) {
	launch_generate_training_samples<RefMarch>(stream, n_rays_local, ray_offset, n_rays_global, rng_state, rng_inc, cfg, views, n_views, bitfield, max_samples,
		counters, ray_indices, rays, numsteps, coords);
}

}  // namespace ngpb
