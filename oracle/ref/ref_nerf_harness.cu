// ref_nerf_harness.cu — launches the REFERENCE's own NeRF training kernels on inputs read from disk and writes what they produced.
//
// Test infrastructure.  This translation unit textually includes /root/reference/src/testbed_nerf.cu (nothing is copied into this
// repo) so that its __global__ kernels are visible here, and calls them exactly as Testbed::train_nerf_step and
// Testbed::update_density_grid_nerf do (src/testbed_nerf.cu:3195-3295, 2476-2633):
//     generate_training_samples_nerf, compute_loss_kernel_train_nerf                                    (:691-849, :852-1180)
//     mark_untrained_density_grid, generate_grid_samples_nerf_nonuniform, splat_grid_samples_nerf_max_nearest_neighbor,
//     ema_grid_samples_nerf, grid_to_bitfield, bitfield_max_pool, the mean reduction                     (:87-396, :2594-2633)
//     init_rays_with_payload_kernel_nerf, advance_pos_nerf_kernel, compact_kernel_nerf, generate_next_nerf_network_inputs,
//     composite_kernel_nerf, shade_kernel_nerf in the loop of NerfTracer::init_rays_from_camera / ::trace / render_nerf
//     (:1380-1528, :398-452, :523-689, :1333-1378, :1588-1800, :2040-2135) with an analytic field in place of the network
// The Testbed class itself is not constructed and nothing else of the application is linked: oracle/ref/Makefile compiles with the
// reference's own flags (--use_fast_math as in its CMakeLists.txt:88), supplies a two-line stand-in for the CMake-generated
// <cmrc/cmrc.hpp> (oracle/ref/stubs) and links with --unresolved-symbols=ignore-all, because the member functions that come along
// with the included file refer to the rest of the application and are never called here.
//
//   ref_nerf <case dir>        reads <case dir>/case.json + *.bin (tools/ref_nerf_cases.py), writes <case dir>/out_*.bin
//
// Needs a GPU: tools/make_ref_nerf_golden.sh runs it on the GPU box and packs the outputs into tests/golden/ref_nerf_*.npz.
#include <testbed_nerf.cu>

#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

using namespace ngp;
using json = nlohmann::json;

static std::string g_dir;

template <typename T>
static std::vector<T> read_bin(const std::string& name) {
	std::ifstream f{g_dir + "/" + name, std::ios::binary | std::ios::ate};
	if (!f.good()) throw std::runtime_error("cannot open " + name);
	size_t n = (size_t)f.tellg();
	f.seekg(0);
	std::vector<T> v(n / sizeof(T));
	f.read((char*)v.data(), (std::streamsize)(v.size() * sizeof(T)));
	return v;
}

template <typename T>
static void write_dev(const std::string& name, const T* dev, size_t count) {
	std::vector<T> h(count);
	if (count) CUDA_CHECK_THROW(cudaMemcpy(h.data(), dev, count * sizeof(T), cudaMemcpyDeviceToHost));
	std::ofstream f{g_dir + "/" + name, std::ios::binary};
	f.write((const char*)h.data(), (std::streamsize)(count * sizeof(T)));
}

struct Views {
	std::vector<GPUMemory<float>> pixels;
	GPUMemory<TrainingImageMetadata> metadata;
	GPUMemory<TrainingXForm> xforms;
	uint32_t n = 0;
};

static void load_views(const json& j, Views& v) {
	v.n = (uint32_t)j.size();
	std::vector<TrainingImageMetadata> md(v.n);
	std::vector<TrainingXForm> xf(v.n);
	v.pixels.resize(v.n);
	for (uint32_t i = 0; i < v.n; ++i) {
		const json& jv = j[i];
		auto px = read_bin<float>("pixels_" + std::to_string(i) + ".bin");
		v.pixels[i].resize_and_copy_from_host(px);
		TrainingImageMetadata& m = md[i];
		m.pixels = v.pixels[i].data();
		m.image_data_type = EImageDataType::Float;
		m.resolution = ivec2((int)jv["w"], (int)jv["h"]);
		m.focal_length = vec2((float)jv["fx"], (float)jv["fy"]);
		m.principal_point = vec2((float)jv["px"], (float)jv["py"]);
		m.lens = Lens{};
		if ((int)jv["lens_mode"] == 1) {
			m.lens.mode = ELensMode::OpenCV;
			for (int k = 0; k < 4; ++k) m.lens.params[k] = (float)jv["lens_params"][k];
		}
		for (int c = 0; c < 4; ++c)
			for (int r = 0; r < 3; ++r) xf[i].start[c][r] = (float)jv["xform"][c * 3 + r];
		xf[i].end = xf[i].start;
	}
	v.metadata.resize_and_copy_from_host(md);
	v.xforms.resize_and_copy_from_host(xf);
}

static default_rng_t make_rng(const json& j) {
	default_rng_t rng;
	rng.state = std::stoull(j["rng_state"].get<std::string>());
	rng.inc = std::stoull(j["rng_inc"].get<std::string>());
	return rng;
}

static BoundingBox make_aabb(const json& j) {
	return BoundingBox{
		vec3((float)j["aabb_min"][0], (float)j["aabb_min"][1], (float)j["aabb_min"][2]),
		vec3((float)j["aabb_max"][0], (float)j["aabb_max"][1], (float)j["aabb_max"][2])};
}

static void run_train_case(const json& j) {
	cudaStream_t stream = nullptr;
	Views views;
	load_views(j["views"], views);
	const uint32_t n_rays = j["n_rays"], n_rays_total = j["n_rays_total"], max_samples = j["max_samples"], batch = j["batch"];
	const uint32_t max_cascade = j["max_cascade"];
	const BoundingBox aabb = make_aabb(j);
	const default_rng_t rng = make_rng(j);
	const bool snap = (int)j["snap_to_pixel_centers"] != 0;
	const float cone_angle_constant = j["cone_angle_constant"];

	auto bitfield_h = read_bin<uint8_t>("bitfield.bin");
	GPUMemory<uint8_t> bitfield;
	bitfield.resize_and_copy_from_host(bitfield_h);

	GPUMemory<uint32_t> counters(4);
	counters.memset(0);
	GPUMemory<uint32_t> ray_indices(n_rays), numsteps(n_rays * 2);
	GPUMemory<Ray> rays(n_rays);
	GPUMemory<float> coords((size_t)max_samples * 7);
	ray_indices.memset(0);
	numsteps.memset(0);
	rays.memset(0);
	coords.memset(0);
	uint32_t* ray_counter = counters.data();
	uint32_t* numsteps_counter = counters.data() + 1;

	linear_kernel(
		generate_training_samples_nerf, 0, stream, n_rays, aabb, max_samples, n_rays_total, rng, ray_counter, numsteps_counter, ray_indices.data(),
		rays.data(), numsteps.data(), PitchedPtr<NerfCoordinate>((NerfCoordinate*)coords.data(), 1, 0, 0), views.n, views.metadata.data(),
		views.xforms.data(), bitfield.data(), max_cascade, false, (float*)nullptr, snap, false, cone_angle_constant, Buffer2DView<const vec2>{},
		(const float*)nullptr, (const float*)nullptr, (const float*)nullptr, ivec2(0), (const float*)nullptr, 0u
	);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	uint32_t cnt[2];
	CUDA_CHECK_THROW(cudaMemcpy(cnt, counters.data(), 8, cudaMemcpyDeviceToHost));
	const uint32_t n_kept = cnt[0], n_samples = std::min(cnt[1], max_samples);
	write_dev("out_gen_counters.bin", counters.data(), 2);
	write_dev("out_ray_indices.bin", ray_indices.data(), n_kept);
	write_dev("out_rays.bin", (const float*)rays.data(), (size_t)n_kept * 6);
	write_dev("out_numsteps.bin", numsteps.data(), (size_t)n_kept * 2);
	write_dev("out_coords.bin", coords.data(), (size_t)n_samples * 7);
	printf("generate: %u rays kept, %u samples (counter %u)\n", n_kept, n_samples, cnt[1]);

	// ---- loss + compaction on synthetic network outputs (4 halves per sample: padded_output_width = 4)
	auto net_h = read_bin<__half>("net_out.bin");
	GPUMemory<__half> net_out;
	net_out.resize_and_copy_from_host(net_h);
	std::vector<uint32_t> numsteps_h((size_t)n_rays * 2);
	CUDA_CHECK_THROW(cudaMemcpy(numsteps_h.data(), numsteps.data(), numsteps_h.size() * 4, cudaMemcpyDeviceToHost));
	GPUMemory<float> coords_compacted((size_t)batch * 7), loss(n_rays), mean_density(1);
	GPUMemory<__half> dloss((size_t)batch * 4);
	GPUMemory<uint32_t> compacted_counter(1);
	GPUMemory<vec3> exposure(views.n);
	exposure.memset(0);
	const float md = j["mean_density"];
	CUDA_CHECK_THROW(cudaMemcpy(mean_density.data(), &md, 4, cudaMemcpyHostToDevice));
	const vec3 background{(float)j["background_color"][0], (float)j["background_color"][1], (float)j["background_color"][2]};
	int v = 0;
	for (const json& lv : j["loss_variants"]) {
		CUDA_CHECK_THROW(cudaMemcpy(numsteps.data(), numsteps_h.data(), numsteps_h.size() * 4, cudaMemcpyHostToDevice));
		coords_compacted.memset(0);
		dloss.memset(0);
		loss.memset(0);
		compacted_counter.memset(0);
		linear_kernel(
			compute_loss_kernel_train_nerf, 0, stream, n_rays, aabb, n_rays_total, rng, batch, ray_counter, (float)j["loss_scale"], 4, Buffer2DView<const vec4>{},
			(float*)nullptr, ivec2(0), ELossType::L2, background, (EColorSpace)(int)j["color_space"], (int)lv["random_bg_color"] != 0,
			(int)j["linear_colors"] != 0, views.n, views.metadata.data(), net_out.data(), compacted_counter.data(), ray_indices.data(), rays.data(),
			numsteps.data(), PitchedPtr<const NerfCoordinate>((NerfCoordinate*)coords.data(), 1, 0, 0),
			PitchedPtr<NerfCoordinate>((NerfCoordinate*)coords_compacted.data(), 1, 0, 0), dloss.data(), (ELossType)(int)lv["loss_type"], ELossType::L1,
			loss.data(), false, (float*)nullptr, (ENerfActivation)(int)j["rgb_activation"], (ENerfActivation)(int)j["density_activation"], snap, (float*)nullptr,
			(const float*)nullptr, (const float*)nullptr, (const float*)nullptr, ivec2(0), ivec2(0), (const float*)nullptr, ivec2(0), (float*)nullptr,
			(float*)nullptr, mean_density.data(), max_cascade, exposure.data(), (vec3*)nullptr, 0.0f, (float)j["near_distance"]
		);
		CUDA_CHECK_THROW(cudaDeviceSynchronize());
		uint32_t total = 0;
		CUDA_CHECK_THROW(cudaMemcpy(&total, compacted_counter.data(), 4, cudaMemcpyDeviceToHost));
		const uint32_t n_comp = std::min(total, batch);
		const std::string p = "out_loss" + std::to_string(v) + "_";
		write_dev(p + "counter.bin", compacted_counter.data(), 1);
		write_dev(p + "numsteps.bin", numsteps.data(), (size_t)n_kept * 2);
		if (v == 0) write_dev(p + "coords.bin", coords_compacted.data(), (size_t)n_comp * 7);   // plain copies of the inputs: one variant is enough
		write_dev(p + "dloss.bin", dloss.data(), (size_t)n_comp * 4);
		write_dev(p + "loss.bin", loss.data(), n_kept);
		printf("loss variant %d: %u compacted samples\n", v, total);
		++v;
	}
}

static void run_grid_case(const json& j) {
	cudaStream_t stream = nullptr;
	Views views;
	load_views(j["views"], views);
	const uint32_t max_cascade = j["max_cascade"];
	const uint32_t n_elements = NERF_GRID_N_CELLS() * (max_cascade + 1);
	const BoundingBox aabb = make_aabb(j);
	default_rng_t rng = make_rng(j);
	const float decay = j["decay"];
	const ENerfActivation rgb_act = (ENerfActivation)(int)j["rgb_activation"], density_act = (ENerfActivation)(int)j["density_activation"];

	GPUMemory<float> grid(n_elements), grid_tmp(n_elements);
	grid.memset(0);
	GPUMemory<uint8_t> bitfield(grid_mip_offset(NERF_CASCADES()) / 8);
	GPUMemory<float> mean(reduce_sum_workspace_size(NERF_GRID_N_CELLS()));
	uint32_t ema_step = 0;
	int k = 0;
	for (const json& st : j["steps"]) {
		const uint32_t n_uni = st["n_uniform"], n_non = st["n_nonuniform"], n_tot = n_uni + n_non;
		if ((int)st["mark_untrained"] != 0) {
			linear_kernel(mark_untrained_density_grid, 0, stream, n_elements, grid.data(), views.n, views.metadata.data(), views.xforms.data(), (int)st["clear_visible"] != 0);
			CUDA_CHECK_THROW(cudaDeviceSynchronize());
			write_dev("out_grid" + std::to_string(k) + "_marked.bin", grid.data(), n_elements);
		}
		GPUMemory<NerfPosition> positions(n_tot);
		GPUMemory<uint32_t> indices(n_tot);
		grid_tmp.memset(0);
		linear_kernel(generate_grid_samples_nerf_nonuniform, 0, stream, n_uni, rng, ema_step, aabb, grid.data(), positions.data(), indices.data(), max_cascade + 1, -0.01f);
		rng.advance();
		linear_kernel(generate_grid_samples_nerf_nonuniform, 0, stream, n_non, rng, ema_step, aabb, grid.data(), positions.data() + n_uni, indices.data() + n_uni, max_cascade + 1, NERF_MIN_OPTICAL_THICKNESS());
		rng.advance();
		CUDA_CHECK_THROW(cudaDeviceSynchronize());
		const std::string p = "out_grid" + std::to_string(k) + "_";
		write_dev(p + "positions.bin", (const float*)positions.data(), (size_t)n_tot * (sizeof(NerfPosition) / sizeof(float)));   // 3 floats per sample here
		write_dev(p + "indices.bin", indices.data(), n_tot);
		auto net_h = read_bin<__half>("grid_net_" + std::to_string(k) + ".bin");
		if (net_h.size() != n_tot) throw std::runtime_error("grid_net size mismatch");
		GPUMemory<__half> mlp_out;
		mlp_out.resize_and_copy_from_host(net_h);
		linear_kernel(splat_grid_samples_nerf_max_nearest_neighbor, 0, stream, n_tot, indices.data(), mlp_out.data(), grid_tmp.data(), rgb_act, density_act);
		linear_kernel(ema_grid_samples_nerf, 0, stream, n_elements, decay, ema_step, grid.data(), grid_tmp.data());
		++ema_step;
		// update_density_grid_mean_and_bitfield
		const uint32_t n_cells = NERF_GRID_N_CELLS();
		CUDA_CHECK_THROW(cudaMemsetAsync(mean.data(), 0, sizeof(float), stream));
		reduce_sum(grid.data(), [n_cells] __device__(float val) { return fmaxf(val, 0.f) / (n_cells); }, mean.data(), n_cells, stream);
		linear_kernel(grid_to_bitfield, 0, stream, n_cells / 8 * NERF_CASCADES(), n_cells / 8 * (max_cascade + 1), grid.data(), bitfield.data(), mean.data());
		for (uint32_t level = 1; level < NERF_CASCADES(); ++level) {
			linear_kernel(bitfield_max_pool, 0, stream, n_cells / 64, bitfield.data() + grid_mip_offset(level - 1) / 8, bitfield.data() + grid_mip_offset(level) / 8);
		}
		CUDA_CHECK_THROW(cudaDeviceSynchronize());
		write_dev(p + "grid.bin", grid.data(), n_elements);
		write_dev(p + "mean.bin", mean.data(), 1);
		write_dev(p + "bitfield.bin", bitfield.data(), bitfield.size());
		printf("grid step %d: %u samples\n", k, n_tot);
		++k;
	}
	std::ofstream f{g_dir + "/out_rng.txt"};
	f << rng.state << " " << rng.inc << "\n";
}

// The "network" of the render case: raw outputs as a closed form of the warped position, in round-to-nearest float arithmetic that
// numpy reproduces exactly (tools/ref_nerf_cases.py: render_field), written where NerfNetwork::inference_mixed_precision would put
// them — row-major [4 x n_elements] halves (GPUMatrix<network_precision_t, RM>, testbed_nerf.cu:1766).
__global__ void analytic_field_kernel(const uint32_t n, const uint32_t stride, const float* __restrict__ coords, __half* __restrict__ out, float a, float b, float c) {
	const uint32_t i = threadIdx.x + blockIdx.x * blockDim.x;
	if (i >= n) return;
	const float dx = __fsub_rn(coords[i * 7 + 0], 0.5f), dy = __fsub_rn(coords[i * 7 + 1], 0.5f), dz = __fsub_rn(coords[i * 7 + 2], 0.5f);
	const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
	out[i + 0 * stride] = __float2half_rn(__fmul_rn(dx, c));
	out[i + 1 * stride] = __float2half_rn(__fmul_rn(dy, c));
	out[i + 2 * stride] = __float2half_rn(__fmul_rn(dz, c));
	out[i + 3 * stride] = __float2half_rn(__fsub_rn(a, __fmul_rn(r2, b)));
}

struct RaysSoa {
	GPUMemory<vec4> rgba;
	GPUMemory<float> depth;
	GPUMemory<NerfPayload> payload;
	void resize(size_t n) {
		rgba.resize(n);
		depth.resize(n);
		payload.resize(n);
		rgba.memset(0);
		depth.memset(0);
		payload.memset(0);
	}
};

static void run_render_case(const json& j) {
	cudaStream_t stream = nullptr;
	const ivec2 resolution((int)j["width"], (int)j["height"]);
	const uint32_t n_pixels = (uint32_t)resolution.x * (uint32_t)resolution.y;
	const size_t n_padded = next_multiple((size_t)n_pixels, size_t(BATCH_SIZE_GRANULARITY));
	const vec2 focal_length((float)j["focal_x"], (float)j["focal_y"]);
	const vec2 screen_center((float)j["screen_x"], (float)j["screen_y"]);
	mat4x3 camera;
	for (int c = 0; c < 4; ++c)
		for (int r = 0; r < 3; ++r) camera[c][r] = (float)j["camera"][c * 3 + r];
	const BoundingBox train_aabb = make_aabb(j);
	const BoundingBox render_aabb{
		vec3((float)j["render_aabb_min"][0], (float)j["render_aabb_min"][1], (float)j["render_aabb_min"][2]),
		vec3((float)j["render_aabb_max"][0], (float)j["render_aabb_max"][1], (float)j["render_aabb_max"][2])};
	const mat3 render_aabb_to_local = mat3::identity();
	const uint32_t max_cascade = j["max_cascade"], sample_index = j["spp_index"];
	const float cone_angle_constant = j["cone_angle_constant"], min_transmittance = j["min_transmittance"], near_distance = j["near_distance"];
	const ENerfActivation rgb_act = (ENerfActivation)(int)j["rgb_activation"], density_act = (ENerfActivation)(int)j["density_activation"];
	const float fa = j["field_a"], fb = j["field_b"], fc = j["field_c"];

	auto bitfield_h = read_bin<uint8_t>("bitfield.bin");
	GPUMemory<uint8_t> bitfield;
	bitfield.resize_and_copy_from_host(bitfield_h);

	GPUMemory<vec4> frame(n_pixels);
	GPUMemory<float> depth_buffer(n_pixels);
	frame.memset(0);
	depth_buffer.memset(0);
	RaysSoa rays[2], rays_hit;
	rays[0].resize(n_padded);
	rays[1].resize(n_padded);
	rays_hit.resize(n_padded);
	GPUMemory<__half> network_output(n_padded * MAX_STEPS_INBETWEEN_COMPACTION * 4);
	GPUMemory<float> network_input(n_padded * MAX_STEPS_INBETWEEN_COMPACTION * 7);
	GPUMemory<uint32_t> counters(64);
	counters.memset(0);
	uint32_t* hit_counter = counters.data();
	uint32_t* alive_counter = counters.data() + 32;

	// ---- NerfTracer::init_rays_from_camera
	const dim3 threads = {16, 8, 1};
	const dim3 blocks = {div_round_up((uint32_t)resolution.x, threads.x), div_round_up((uint32_t)resolution.y, threads.y), 1};
	init_rays_with_payload_kernel_nerf<<<blocks, threads, 0, stream>>>(
		sample_index, rays[0].payload.data(), resolution, focal_length, camera, camera, vec4(0.0f), screen_center, vec3(0.0f), true, render_aabb,
		render_aabb_to_local, near_distance, 1.0f, 0.0f, Foveation{}, Lens{}, Buffer2DView<const vec4>{}, frame.data(), depth_buffer.data(),
		Buffer2DView<const uint8_t>{}, Buffer2DView<const vec2>{}, ERenderMode::Shade
	);
	linear_kernel(
		advance_pos_nerf_kernel, 0, stream, n_pixels, render_aabb, render_aabb_to_local, camera[2], focal_length, sample_index, rays[0].payload.data(),
		bitfield.data(), 0u, max_cascade, cone_angle_constant
	);

	// ---- NerfTracer::trace
	uint32_t n_alive = n_pixels;
	uint32_t i = 1;
	uint32_t double_buffer_index = 0;
	uint64_t n_queries = 0;
	while (i < MARCH_ITER) {
		RaysSoa& rays_current = rays[(double_buffer_index + 1) % 2];
		RaysSoa& rays_tmp = rays[double_buffer_index % 2];
		++double_buffer_index;
		CUDA_CHECK_THROW(cudaMemsetAsync(alive_counter, 0, sizeof(uint32_t), stream));
		linear_kernel(
			compact_kernel_nerf, 0, stream, n_alive, rays_tmp.rgba.data(), rays_tmp.depth.data(), rays_tmp.payload.data(), rays_current.rgba.data(),
			rays_current.depth.data(), rays_current.payload.data(), rays_hit.rgba.data(), rays_hit.depth.data(), rays_hit.payload.data(), alive_counter,
			hit_counter
		);
		CUDA_CHECK_THROW(cudaMemcpyAsync(&n_alive, alive_counter, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
		CUDA_CHECK_THROW(cudaStreamSynchronize(stream));
		if (n_alive == 0) break;
		const uint32_t target_n_queries = 2 * 1024 * 1024;
		const uint32_t n_steps_between_compaction =
			clamp(target_n_queries / n_alive, (uint32_t)MIN_STEPS_INBETWEEN_COMPACTION, (uint32_t)MAX_STEPS_INBETWEEN_COMPACTION);
		PitchedPtr<NerfCoordinate> input_data((NerfCoordinate*)network_input.data(), 1, 0, 0);
		linear_kernel(
			generate_next_nerf_network_inputs, 0, stream, n_alive, render_aabb, render_aabb_to_local, train_aabb, focal_length, camera[2],
			rays_current.payload.data(), input_data, n_steps_between_compaction, bitfield.data(), 0u, max_cascade, cone_angle_constant, (const float*)nullptr
		);
		const uint32_t n_elements = next_multiple(n_alive * n_steps_between_compaction, BATCH_SIZE_GRANULARITY);
		// rows a ray did not fill (it left the volume mid-batch) hold stale coordinates; composite_kernel_nerf never reads their outputs
		linear_kernel(analytic_field_kernel, 0, stream, n_alive * n_steps_between_compaction, n_elements, network_input.data(), network_output.data(), fa, fb, fc);
		n_queries += (uint64_t)n_alive * n_steps_between_compaction;
		linear_kernel(
			composite_kernel_nerf, 0, stream, n_alive, n_elements, i, train_aabb, camera, focal_length, 1.0f, false, rays_current.rgba.data(),
			rays_current.depth.data(), rays_current.payload.data(), input_data, network_output.data(), 4u, n_steps_between_compaction, ERenderMode::Shade,
			bitfield.data(), rgb_act, density_act, -1, min_transmittance
		);
		i += n_steps_between_compaction;
	}
	uint32_t n_hit = 0;
	CUDA_CHECK_THROW(cudaMemcpy(&n_hit, hit_counter, sizeof(uint32_t), cudaMemcpyDeviceToHost));
	// ---- render_nerf: shade the rays that hit something
	linear_kernel(
		shade_kernel_nerf, 0, stream, n_hit, false, camera, false, 1.0f, rays_hit.rgba.data(), rays_hit.depth.data(), rays_hit.payload.data(),
		ERenderMode::Shade, false, frame.data(), depth_buffer.data()
	);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	write_dev("out_frame.bin", (const float*)frame.data(), (size_t)n_pixels * 4);
	write_dev("out_depth.bin", depth_buffer.data(), n_pixels);
	// per pixel: how many steps the ray took (payload.n_steps of the rays that hit), 0 elsewhere
	std::vector<NerfPayload> ph(n_hit);
	if (n_hit) CUDA_CHECK_THROW(cudaMemcpy(ph.data(), rays_hit.payload.data(), n_hit * sizeof(NerfPayload), cudaMemcpyDeviceToHost));
	std::vector<uint32_t> steps(n_pixels, 0);
	for (const auto& p : ph) steps[p.idx] = p.n_steps;
	{
		std::ofstream f{g_dir + "/out_steps.bin", std::ios::binary};
		f.write((const char*)steps.data(), (std::streamsize)(steps.size() * 4));
	}
	printf("render: %u of %u rays hit, %llu field queries, %u march iterations\n", n_hit, n_pixels, (unsigned long long)n_queries, i);
}

int main(int argc, char** argv) {
	if (argc != 2) {
		fprintf(stderr, "usage: ref_nerf <case dir>\n");
		return 64;
	}
	g_dir = argv[1];
	try {
		std::ifstream f{g_dir + "/case.json"};
		json j = json::parse(f);
		const std::string type = j["type"];
		if (type == "train") run_train_case(j);
		else if (type == "grid") run_grid_case(j);
		else if (type == "render") run_render_case(j);
		else throw std::runtime_error("unknown case type " + type);
	} catch (const std::exception& e) {
		fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
	return 0;
}
