// testbed.cu — host-side Testbed for the NeRF path and the extern "C" boundary (include/ngp_b200.h).
// Mirrors the parts of Testbed the hot path needs: reset_network (src/testbed.cu:4160-4412), train (:4561-4647),
// train_nerf / train_nerf_step / training_prep_nerf / NerfCounters (src/testbed_nerf.cu:2669-3398).
#include <cmath>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <vector>

#include "common.cuh"
#include "dp_nccl.h"
#include "host_util.h"
#include "json_mini.h"
#include "march.cuh"
#include "msgpack_mini.h"

namespace ngpb {

unsigned long long g_launch_count = 0;
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

// kernels' host launchers (nerf_net.cu, march.cu, optimizer.cu, render.cu)
void nerf_inference(const ngp_nerf_desc& d, cudaStream_t stream, uint32_t n, const float* coords, const __half* params, __half* out, uint32_t out_stride);
void nerf_inference_counted(const ngp_nerf_desc& d, cudaStream_t stream, uint32_t n_max, const uint32_t* n_dev, const float* coords, const __half* params, __half* out);
void nerf_inference_rays(const ngp_nerf_desc& d, cudaStream_t stream, uint32_t n_rays_max, const ngp_nerf_counters* counters, uint32_t* queue,
	const uint32_t* numsteps, const float* coords, const __half* params, uint32_t density_activation, __half* out, uint32_t chunk, uint32_t train_mode);
void nerf_density(const ngp_nerf_desc& d, cudaStream_t stream, uint32_t n, const float* positions, uint32_t pos_stride, const __half* params, __half* out);
void grid_encode(const ngp_grid_desc& g, cudaStream_t stream, uint32_t n, const float* positions, uint32_t pos_stride, const __half* grid, __half* out);
void nerf_forward_backward(const ngp_nerf_desc& d, cudaStream_t stream, uint32_t n, const float* coords, const __half* params, const __half* dL_dout,
	__half* grads, float* mlp_grads_f32, __half* out);
void set_scatter_aggregation(int mode);
void profile_mlp_phase(const ngp_nerf_desc& d, cudaStream_t stream, uint32_t n, const float* coords, const __half* params, const __half* dL_dout, __half* grads,
	float* mlp_grads_f32);
void optimizer_step(const ngp_nerf_desc& d, cudaStream_t stream, const ngp_adam_cfg& cfg, float* params_fp32, __half* params_fp16, __half* params_ema,
	__half* grads, float* m1, float* m2, uint32_t* steps);
void generate_training_samples(cudaStream_t stream, uint32_t n_rays_local, uint32_t ray_offset, uint32_t n_rays_global, uint64_t rng_state, uint64_t rng_inc,
	const ngp_nerf_train_cfg& cfg, const ngp_train_view* views, uint32_t n_views, const uint8_t* bitfield, uint32_t max_samples, ngp_nerf_counters* counters,
	uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords);
void compute_loss(cudaStream_t stream, uint32_t n_rays_local, uint32_t n_rays_global, uint64_t rng_state, uint64_t rng_inc, const ngp_nerf_train_cfg& cfg,
	const ngp_train_view* views, uint32_t n_views, const __half* network_output, uint32_t max_compacted, ngp_nerf_counters* counters,
	const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords, float* coords_compacted, __half* dloss, float* loss_per_ray,
	const float* mean_density);
void fill_rollover(cudaStream_t stream, uint32_t target_batch, const ngp_nerf_counters* counters, float* coords_compacted, __half* dloss);
void update_bitfield(cudaStream_t stream, uint32_t max_cascade, const float* density_grid, uint8_t* bitfield, float* mean_density, float* partial1024);
size_t density_grid_scratch_bytes(uint32_t max_cascade);
void update_density_grid(const ngp_nerf_desc& d, cudaStream_t stream, const ngp_nerf_train_cfg& cfg, const __half* params, uint64_t* grid_rng_state,
	uint64_t grid_rng_inc, uint32_t training_step, uint32_t ema_step, float decay, const ngp_train_view* views, uint32_t n_views, float* density_grid,
	uint8_t* bitfield, float* mean_density, void* scratch);
size_t render_scratch_bytes(int32_t width, int32_t rows);
void render_nerf(const ngp_nerf_desc& d, cudaStream_t stream, const ngp_render_cfg& cfg, int32_t y0, int32_t y1, const __half* params, const uint8_t* bitfield,
	float* rgba, float* depth, void* scratch, uint32_t* n_steps_total);
float reduce_sum_f32(cudaStream_t stream, const float* data, uint32_t n, float* scratch_dev);
void render_accumulate(cudaStream_t stream, int32_t w, int32_t h, const float* frame, float* acc, float sample_count, uint32_t color_space);
void render_tonemap(cudaStream_t stream, int32_t w, int32_t h, const ngp_tonemap_cfg& cfg, const float* acc, float* out);
void render_pixel_offset(uint32_t spp, float* out);

// ------------------------------------------------------------------------------------------------------------------
// descriptors
// ------------------------------------------------------------------------------------------------------------------
static uint32_t powi_u32(uint32_t b, uint32_t e) {
	uint32_t r = 1;
	for (uint32_t i = 0; i < e; ++i) r *= b;
	return r;
}

void grid_desc_init(ngp_grid_desc* g, uint32_t n_levels, uint32_t F, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale,
	uint32_t aabb_scale) {
	NGPB_CHECK(n_levels >= 1 && n_levels <= NGP_MAX_LEVELS, "HashGrid: n_levels out of range");
	NGPB_CHECK(F == 1 || F == 2 || F == 4 || F == 8, "HashGrid: n_features_per_level must be 1, 2, 4 or 8");
	memset(g, 0, sizeof(*g));
	g->n_levels = n_levels;
	g->n_features_per_level = F;
	g->log2_hashmap_size = log2_hashmap_size;
	g->base_resolution = base_resolution;
	if (per_level_scale <= 0.0f && n_levels > 1) {
		// src/testbed.cu:4241-4255: finest level resolves 2048 * aabb_scale cells across the cube
		per_level_scale = std::exp(std::log(2048.0f * (float)aabb_scale / (float)base_resolution) / (float)(n_levels - 1));
	}
	if (n_levels == 1 && per_level_scale <= 0.0f) per_level_scale = 1.0f;
	g->per_level_scale = per_level_scale;
	const float log2_scale = std::log2(per_level_scale);
	uint32_t offset = 0;
	for (uint32_t l = 0; l < n_levels; ++l) {
		// grid.h:699-722 (GridType::Hash, 3 position dims)
		const float scale = exp2f((float)l * log2_scale) * (float)base_resolution - 1.0f;
		const uint32_t res = (uint32_t)ceilf(scale) + 1;
		const uint32_t max_params = 0xFFFFFFFFu / 2;
		uint32_t params_in_level = std::pow((float)res, 3.0f) > (float)max_params ? max_params : powi_u32(res, 3);
		params_in_level = next_multiple(params_in_level, 8u);
		params_in_level = std::min(params_in_level, 1u << log2_hashmap_size);
		g->offsets[l] = offset;
		g->resolutions[l] = res;
		g->scales[l] = scale;
		offset += params_in_level;
	}
	g->offsets[n_levels] = offset;
	g->n_params = offset * F;
}

static uint32_t mlp_params(uint32_t n_hidden) { return 64 * 32 + (n_hidden - 1) * 64 * 64 + 16 * 64; }

void nerf_desc_init(ngp_nerf_desc* d, const ngp_grid_desc* g, uint32_t n_hidden_density, uint32_t n_hidden_rgb) {
	NGPB_CHECK(n_hidden_density >= 1 && n_hidden_rgb >= 1, "FullyFusedMLP requires at least 1 hidden layer");
	NGPB_CHECK(g->n_levels * g->n_features_per_level == 32, "this build fuses a 32-wide encoding (n_levels * n_features_per_level == 32)");
	memset(d, 0, sizeof(*d));
	d->grid = *g;
	d->n_hidden_density = n_hidden_density;
	d->n_hidden_rgb = n_hidden_rgb;
	d->density_mlp_offset = 0;
	d->rgb_mlp_offset = mlp_params(n_hidden_density);
	d->grid_offset = d->rgb_mlp_offset + mlp_params(n_hidden_rgb);
	d->n_mlp_params = d->grid_offset;
	d->n_params = d->grid_offset + g->n_params;
}

void march_consts_init(ngp_march_consts* m, float cone_angle) {
	memset(m, 0, sizeof(*m));
	m->cone_angle = cone_angle;
	if (cone_angle <= 1e-5f) return;
	// nerf_device.cuh:384-390, evaluated with the deterministic functions of ngp_detmath.h
	const float log1p_c = ngp_logf(1.0f + cone_angle);
	m->log1p_c = log1p_c;
	m->a = (ngp_logf(min_cone_stepsize()) - ngp_logf(log1p_c)) / log1p_c;
	m->b = (ngp_logf(max_cone_stepsize()) - ngp_logf(log1p_c)) / log1p_c;
	m->at = ngp_expf(m->a * log1p_c);
	m->bt = ngp_expf(m->b * log1p_c);
}

// Trainer::initialize_params: MLPs Xavier-uniform on the host pcg32 (gpu_matrix.h:292-307), hash grid with the GPU fill
// pattern of generate_random_kernel (random.h:40-67: thread i draws 4 values for indices i + n_threads * j).
void nerf_init_params_host(const ngp_nerf_desc* d, uint64_t seed, float* out) {
	std::seed_seq seq{(uint32_t)seed};
	std::vector<uint32_t> seeds(2);
	seq.generate(seeds.begin(), seeds.end());
	Pcg32 rng((uint64_t)seeds.front());
	auto xavier = [&](float* w, uint32_t rows, uint32_t cols) {
		const float scale = std::sqrt(6.0f / (float)(rows + cols));
		for (uint32_t i = 0; i < rows * cols; ++i) w[i] = rng.next_float() * 2.0f * scale - scale;
	};
	float* p = out;
	for (uint32_t net = 0; net < 2; ++net) {
		const uint32_t nh = net ? d->n_hidden_rgb : d->n_hidden_density;
		xavier(p, 64, 32);
		p += 64 * 32;
		for (uint32_t l = 1; l < nh; ++l) {
			xavier(p, 64, 64);
			p += 64 * 64;
		}
		xavier(p, 16, 64);
		p += 16 * 64;
	}
	const size_t n = d->grid.n_params;
	const size_t n_threads_req = (n + 3) / 4;
	const size_t n_threads = ((n_threads_req + 127) / 128) * 128;
	for (size_t i = 0; i < n_threads; ++i) {
		if (i >= n) break;
		Pcg32 r = rng;
		r.advance((uint64_t)i * 4);
		for (size_t j = 0; j < 4; ++j) {
			const size_t idx = i + n_threads * j;
			if (idx >= n) break;
			p[idx] = r.next_float() * (1e-4f - (-1e-4f)) + (-1e-4f);
		}
	}
}

}  // namespace ngpb

using namespace ngpb;

struct ngp_testbed {
	int device = 0;
	cudaStream_t stream = nullptr;

	// dataset (NerfDataset: nerf_loader.h)
	uint32_t n_images = 0, aabb_scale = 1, n_images_for_training = 0;
	float scene_scale = 0.33f;  // NERF_SCALE
	float scene_offset[3] = {0.5f, 0.5f, 0.5f};
	std::vector<ngp_train_view> views;
	std::vector<void*> pixel_bufs;
	DevBuf<ngp_train_view> views_dev;
	bool views_dirty = true;

	// network
	float background_alpha = 1.0f;    // m_background_color.a (testbed.h:1031): used by the render epilogue only
	float exposure = 0.0f;            // m_exposure (testbed.h:1030): stops, applied by the render epilogue's tonemap
	uint32_t render_spp_index = 0;    // sample index of the next render (jitters each ray's first step, advance_pos_nerf)
	bool has_network = false;
	Json network_config;              // as given to reload_network_from_json/file (m_network_config)
	ngp_nerf_desc desc{};
	OptimizerConfig opt;
	float lr_factor = 1.0f;
	uint32_t optimizer_step = 0;
	uint64_t seed = 1337;
	bool train_network = true, train_encoding = true;
	uint32_t drop_overflowing_rays = 0;   // sample generator capacity: 0 = the sample buffer (16 x batch), 1 = the previous step's sample count like the reference (tb_generator_capacity)
	uint32_t full_inference = 2;    // training-time inference schedule: 0 ray-ordered, 1 every generated sample (the reference's), 2 chosen per step (tb_full_inference)
	uint32_t inference_chunk = 0;   // consecutive samples of a ray per tile of the ray-ordered inference pass; 0 = from the last step's samples per ray

	DevBuf<float> params_fp32, m1, m2, mlp_grads_f32;
	DevBuf<__half> params, params_ema, grads;
	DevBuf<uint32_t> param_steps;

	// nerf state
	ngp_nerf_train_cfg cfg{};
	float density_grid_decay = 0.95f;
	DevBuf<float> density_grid, mean_density;
	DevBuf<uint8_t> bitfield, grid_scratch;
	Pcg32 rng{1337}, density_grid_rng{};
	uint32_t training_step = 0, density_grid_ema_step = 0;
	uint32_t rays_per_batch = 1u << 12, measured_batch_size = 0, measured_batch_size_before_compaction = 0, n_rays_total = 0;
	float loss_scalar = 0.0f;
	bool shall_train = true;
	// per-image exposure optimisation (Nerf::Training::optimize_exposure, testbed_nerf.cu:2962-3000): one host-side Adam per image
	// (AdamOptimizer<vec3>, adam_optimizer.h:129-152: lr 1e-3 replaced by the network optimizer's, eps 1e-8, betas 0.9 / 0.99)
	struct ExposureAdam { float variable[3] = {0, 0, 0}, first_moment[3] = {0, 0, 0}, second_moment[3] = {0, 0, 0}; uint32_t iter = 0; };
	bool optimize_exposure = false;
	float exposure_l2_reg = 0.0f;
	uint32_t n_steps_between_cam_updates = 16, n_steps_since_cam_update = 0;
	std::vector<ExposureAdam> cam_exposure;
	DevBuf<float> cam_exposure_dev, cam_exposure_gradient_dev;
	bool cam_exposure_nonzero = false;
	// update_image_async: the host-to-device copy of a replacement frame runs on its own stream into one of two staging buffers, beside
	// whatever the training stream is doing; the training stream picks it up (a device-to-device copy) right before the loss kernel, the
	// first reader of the pixels of a view without masked pixels (tb_apply_pending_images)
	struct ImageStaging { void* buf = nullptr; size_t cap = 0; cudaEvent_t copied = nullptr, consumed = nullptr; bool in_flight = false; };
	struct PendingImage { uint32_t idx, slot; size_t bytes; };
	cudaStream_t copy_stream = nullptr;
	ImageStaging staging[2];
	uint32_t staging_next = 0;
	std::vector<PendingImage> pending_images;

	// per-step scratch.  Everything the sample generator writes exists twice: while step k trains, the generator of step
	// k+1 already runs on a side stream into the other set (it depends only on the occupancy bitfield, the RNG and
	// rays_per_batch, all known once step k's counters are back).
	struct RaySet {
		DevBuf<ngp_nerf_counters> counters;
		DevBuf<uint32_t> ray_indices, numsteps;
		DevBuf<float> rays, coords;
	} set[2];
	uint32_t cur = 0;                     // set used by the step in flight
	DevBuf<float> coords_compacted, loss_per_ray, reduce_scratch;
	DevBuf<__half> mlp_out, dloss;
	uint32_t last_batch = 0;
	bool grads_pending = false;
	bool get_loss_pending = false;
	uint32_t overlap_sample_generation = 2;   // next step's generator on the side stream, gated on this step's forward/backward kernel (tb_prefetch): 0 never, 1 always, 2 data parallel only
	bool front_done = false, controller_done = false;
	cudaStream_t side_stream = nullptr;
	cudaEvent_t ev_front_done = nullptr, ev_prefetch_done = nullptr, ev_main_ready = nullptr, ev_back_done = nullptr;
	struct Readback {
		ngp_nerf_counters counters;
		float loss_partial[1024];
	}* readback = nullptr;                // pinned host memory
	// a sample-generator launch that has been issued ahead of time
	bool prefetch_valid = false;
	uint32_t prefetch_step = 0, prefetch_batch = 0, prefetch_rays = 0, prefetch_max_inference = 0;
	uint64_t prefetch_rng_state = 0;
	uint32_t step_rays_local = 0, step_max_inference = 0, step_batch = 0;
	DevBuf<ngp_nerf_counters> dp_counters;  // stable address for the caller's all-reduce of the counter block

	// data parallel
	uint32_t dp_rank = 0, dp_world = 1;
	// the exchange inside the library (ngp_testbed_init_dp): gradients on the training stream, counters / loss partials on comm_stream
	ngpb::NcclApi::comm_t comm_grads = nullptr, comm_small = nullptr;
	cudaStream_t comm_stream = nullptr;
	cudaEvent_t ev_counters_ready = nullptr;

	// profiling (CUDA events per phase)
	bool profiling = false;
	cudaEvent_t ev[NGP_N_PHASES][2] = {};
	bool ev_used[NGP_N_PHASES] = {};
	float phase_ms[NGP_N_PHASES] = {};
	uint32_t phase_steps = 0;

	// render
	DevBuf<uint8_t> render_scratch;
	DevBuf<float> render_rgba, render_depth;
	DevBuf<uint32_t> render_counter;
	float render_min_transmittance = 0.01f;
	bool render_snap_to_pixel_centers = false;   // m_snap_to_pixel_centers (testbed.h): the reference's default jitters pixels per sample index
	bool render_with_lens_distortion = false;    // m_render_with_lens_distortion / m_render_lens (set by set_camera_to_training_view)
	uint32_t render_lens_mode = NGP_LENS_PERSPECTIVE;
	float render_lens_params[4] = {0, 0, 0, 0};
	uint32_t render_mode = NGP_RENDER_SHADE;     // m_render_mode (ERenderMode)
	uint32_t render_skips_per_tile = 0;          // ngp_render_cfg.skips_per_tile (0 = default)
	uint32_t render_math = NGP_MATH_REFERENCE;   // arithmetic of the render march (ngp_render_cfg.math_mode); `render_math` option

	~ngp_testbed() {
		if (side_stream) {
			cudaStreamSynchronize(side_stream);
			cudaStreamDestroy(side_stream);
		}
		if (comm_stream) cudaStreamSynchronize(comm_stream);
		if (comm_grads) ngpb::NcclApi::get().CommDestroy(comm_grads);
		if (comm_small) ngpb::NcclApi::get().CommDestroy(comm_small);
		if (comm_stream) cudaStreamDestroy(comm_stream);
		for (cudaEvent_t e : {ev_front_done, ev_prefetch_done, ev_main_ready, ev_back_done, ev_counters_ready})
			if (e) cudaEventDestroy(e);
		if (readback) cudaFreeHost(readback);
		if (copy_stream) {
			cudaStreamSynchronize(copy_stream);
			cudaStreamDestroy(copy_stream);
		}
		for (ImageStaging& st : staging) {
			if (st.buf) cudaFree(st.buf);
			if (st.copied) cudaEventDestroy(st.copied);
			if (st.consumed) cudaEventDestroy(st.consumed);
		}
		for (void* p : pixel_bufs)
			if (p) cudaFree(p);
	}
};

namespace ngpb {

static void tb_set_defaults(ngp_testbed* t) {
	ngp_nerf_train_cfg& c = t->cfg;
	memset(&c, 0, sizeof(c));
	c.snap_to_pixel_centers = 1;
	c.random_bg_color = 1;
	c.linear_colors = 0;
	c.color_space = NGP_COLOR_LINEAR;
	c.background_color[0] = c.background_color[1] = c.background_color[2] = 0.0f;
	c.loss_type = NGP_LOSS_L2;
	c.rgb_activation = NGP_ACT_LOGISTIC;
	c.density_activation = NGP_ACT_EXPONENTIAL;
	c.near_distance = 0.1f;
	c.loss_scale = NGP_LOSS_SCALE;
	c.math_mode = NGP_MATH_REFERENCE;   // the sample generator marches with the reference build's arithmetic (march_ref.cu); the oracle tests select NGP_MATH_DETERMINISTIC
}

static void tb_update_scene(ngp_testbed* t) {
	// load_nerf_post (testbed_nerf.cu:2425-2440)
	const float half = 0.5f * (float)std::min<uint32_t>(1u << (NGP_NERF_CASCADES - 1), t->aabb_scale);
	for (int k = 0; k < 3; ++k) {
		t->cfg.aabb_min[k] = 0.5f - half;
		t->cfg.aabb_max[k] = 0.5f + half;
	}
	t->cfg.max_cascade = 0;
	while ((1u << t->cfg.max_cascade) < t->aabb_scale) ++t->cfg.max_cascade;
	march_consts_init(&t->cfg.march, t->aabb_scale <= 1 ? 0.0f : (1.0f / 256.0f));
}

static void tb_upload_views(ngp_testbed* t) {
	if (!t->views_dirty) return;
	t->views_dev.ensure(std::max<size_t>(t->views.size(), 1));
	if (!t->views.empty())
		NGPB_CUDA_CHECK(cudaMemcpyAsync(t->views_dev.p, t->views.data(), t->views.size() * sizeof(ngp_train_view), cudaMemcpyHostToDevice, t->stream));
	t->views_dirty = false;
}

static void tb_alloc_network(ngp_testbed* t) {
	const uint32_t n = t->desc.n_params;
	t->params_fp32.ensure(n);
	t->params.ensure(n);
	t->params_ema.ensure(n);
	t->grads.ensure(n);
	t->m1.ensure(n);
	t->m2.ensure(n);
	t->param_steps.ensure(n);
	t->mlp_grads_f32.ensure(t->desc.n_mlp_params);
	NGPB_CUDA_CHECK(cudaMemsetAsync(t->grads.p, 0, n * sizeof(__half), t->stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(t->m1.p, 0, n * sizeof(float), t->stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(t->m2.p, 0, n * sizeof(float), t->stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(t->param_steps.p, 0, n * sizeof(uint32_t), t->stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(t->mlp_grads_f32.p, 0, t->desc.n_mlp_params * sizeof(float), t->stream));
}

__global__ void k_cast_params(const uint32_t n, const float* __restrict__ src, __half* a, __half* b) {  // a may equal b
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const __half h = __float2half_rn(src[i]);
	a[i] = h;
	b[i] = h;
}

static void tb_set_params_fp32(ngp_testbed* t, const float* host, uint32_t n) {
	NGPB_CHECK(t->has_network, "no network configured");
	NGPB_CHECK(n == t->desc.n_params, "set_params: wrong parameter count");
	NGPB_CUDA_CHECK(cudaMemcpyAsync(t->params_fp32.p, host, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, t->stream));
	k_cast_params<<<div_round_up(n, 256), 256, 0, t->stream>>>(n, t->params_fp32.p, t->params.p, t->params_ema.p);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
}

// Testbed::reset_network (src/testbed.cu:4160-4412), NeRF mode
static void tb_reset_network(ngp_testbed* t, const Json& config_in) {
	t->network_config = config_in;
	t->network_config.obj.erase("snapshot");
	const Json& config = t->network_config;
	const Json& enc = config.sub("encoding");
	const Json& net = config.sub("network");
	const Json& rgb = config.sub("rgb_network");
	const Json& dir = config.sub("dir_encoding");
	const Json& loss = config.sub("loss");

	const std::string enc_type = to_lower(enc.value("otype", std::string("HashGrid")));
	NGPB_CHECK(enc_type == "hashgrid" || enc_type == "grid", "encoding.otype '" + enc_type + "' is not supported by ngp_b200 (HashGrid only)");
	if (enc.contains("type")) NGPB_CHECK(to_lower(enc.value("type", std::string("hash"))) == "hash", "encoding.type must be Hash");
	if (enc.contains("interpolation")) NGPB_CHECK(to_lower(enc.value("interpolation", std::string("linear"))) == "linear", "encoding.interpolation must be Linear");
	auto check_mlp = [&](const Json& j, const char* name) {
		const std::string ot = to_lower(j.value("otype", std::string("FullyFusedMLP")));
		NGPB_CHECK(ot == "fullyfusedmlp" || ot == "cutlassmlp" || ot == "megakernelmlp", std::string(name) + ".otype not supported");
		NGPB_CHECK(to_lower(j.value("activation", std::string("ReLU"))) == "relu", std::string(name) + ".activation must be ReLU");
		NGPB_CHECK(to_lower(j.value("output_activation", std::string("None"))) == "none", std::string(name) + ".output_activation must be None");
		NGPB_CHECK((uint32_t)j.value("n_neurons", 64.0) == 64, std::string(name) + ".n_neurons must be 64");
	};
	check_mlp(net, "network");
	check_mlp(rgb, "rgb_network");
	if (dir.contains("otype")) {
		const std::string dt = to_lower(dir.value("otype", std::string("Composite")));
		if (dt == "composite") {
			NGPB_CHECK(dir.sub("nested").type == Json::Array && !dir.sub("nested").arr.empty(), "dir_encoding.nested missing");
			const Json& sh = dir.sub("nested").arr[0];
			NGPB_CHECK(to_lower(sh.value("otype", std::string(""))) == "sphericalharmonics" && (uint32_t)sh.value("degree", 4.0) == 4,
				"dir_encoding must be SphericalHarmonics degree 4");
		} else {
			NGPB_CHECK(dt == "sphericalharmonics" && (uint32_t)dir.value("degree", 4.0) == 4, "dir_encoding must be SphericalHarmonics degree 4");
		}
	}

	const uint32_t F = (uint32_t)enc.value("n_features_per_level", 2.0);
	uint32_t L = (uint32_t)enc.value("n_levels", 16.0);
	if (enc.contains("n_features") && enc.value("n_features", 0.0) > 0) L = (uint32_t)enc.value("n_features", 0.0) / F;
	const uint32_t log2_T = (uint32_t)enc.value("log2_hashmap_size", 15.0);
	uint32_t base_res = (uint32_t)enc.value("base_resolution", 0.0);
	if (!base_res) base_res = 1u << (log2_T / 3);
	const float pls = (float)enc.value("per_level_scale", 0.0);
	ngp_grid_desc g;
	grid_desc_init(&g, L, F, log2_T, base_res, pls, t->aabb_scale);
	NGPB_CHECK(F == 2 || F == 4, "HashGrid: n_features_per_level must be 2 or 4 in this build");
	nerf_desc_init(&t->desc, &g, (uint32_t)net.value("n_hidden_layers", 1.0), (uint32_t)rgb.value("n_hidden_layers", 2.0));

	// loss (string_to_loss_type) — NeRF bypasses tcnn's Loss object (src/testbed.cu:4208-4215)
	t->cfg.loss_type = parse_loss_type(loss);

	t->opt = parse_optimizer_chain(config.sub("optimizer"));

	t->has_network = true;
	tb_alloc_network(t);

	// state reset (src/testbed.cu:4163-4178)
	t->rng = Pcg32(t->seed);
	t->rays_per_batch = 1u << 12;
	t->measured_batch_size = 0;
	t->measured_batch_size_before_compaction = 0;
	t->density_grid_rng = Pcg32((uint64_t)t->rng.next_uint());
	t->training_step = 0;
	t->optimizer_step = 0;
	t->lr_factor = 1.0f;
	t->density_grid_ema_step = 0;
	t->n_rays_total = 0;
	t->loss_scalar = 0.0f;
	t->shall_train = true;
	// reset_camera_extrinsics (src/testbed.cu:4180, testbed_nerf.cu:2215-2227): the per-image exposures start over as well
	t->cam_exposure.clear();
	t->cam_exposure_nonzero = false;
	t->cam_exposure_dev.release();
	t->cam_exposure_gradient_dev.release();
	t->n_steps_since_cam_update = 0;

	const uint32_t n_grid = GRID_N_CELLS * (t->cfg.max_cascade + 1);
	t->density_grid.ensure(GRID_N_CELLS * NGP_NERF_CASCADES);
	t->bitfield.ensure(GRID_N_CELLS / 8 * NGP_NERF_CASCADES);
	t->mean_density.ensure(4);
	NGPB_CUDA_CHECK(cudaMemsetAsync(t->density_grid.p, 0, sizeof(float) * GRID_N_CELLS * NGP_NERF_CASCADES, t->stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(t->bitfield.p, 0, GRID_N_CELLS / 8 * NGP_NERF_CASCADES, t->stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(t->mean_density.p, 0, 16, t->stream));
	(void)n_grid;

	std::vector<float> init(t->desc.n_params);
	nerf_init_params_host(&t->desc, t->seed, init.data());
	tb_set_params_fp32(t, init.data(), t->desc.n_params);
}

static void tb_ensure_step_scratch(ngp_testbed* t, uint32_t batch) {
	const uint32_t max_samples = batch * 16;
	const uint32_t max_rays = 1u << 18;
	t->dp_counters.ensure(1);
	for (auto& rs : t->set) {
		rs.counters.ensure(1);
		rs.ray_indices.ensure(max_rays);
		rs.numsteps.ensure((size_t)max_rays * 2);
		rs.rays.ensure((size_t)max_rays * 6);
		rs.coords.ensure((size_t)max_samples * 7);
	}
	t->loss_per_ray.ensure(max_rays);
	t->reduce_scratch.ensure(1024);
	t->mlp_out.ensure((size_t)max_samples * 4);
	if (!t->side_stream) {
		NGPB_CUDA_CHECK(cudaStreamCreateWithFlags(&t->side_stream, cudaStreamNonBlocking));
		NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&t->ev_front_done, cudaEventDisableTiming));
		NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&t->ev_prefetch_done, cudaEventDisableTiming));
		NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&t->ev_back_done, cudaEventDisableTiming));
		NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&t->ev_main_ready, cudaEventDisableTiming));
		NGPB_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&t->readback), sizeof(*t->readback), cudaHostAllocDefault));
	}
	t->coords_compacted.ensure((size_t)batch * 7);
	t->dloss.ensure((size_t)batch * 4);
	t->grid_scratch.ensure(density_grid_scratch_bytes(t->cfg.max_cascade));
	t->last_batch = batch;
}

struct PhaseTimer {
	ngp_testbed* t;
	int phase;
	PhaseTimer(ngp_testbed* tb, int ph) : t(tb), phase(ph) {
		if (!t->profiling) return;
		if (!t->ev[phase][0]) {
			cudaEventCreate(&t->ev[phase][0]);
			cudaEventCreate(&t->ev[phase][1]);
		}
		cudaEventRecord(t->ev[phase][0], t->stream);
	}
	~PhaseTimer() {
		if (!t->profiling) return;
		cudaEventRecord(t->ev[phase][1], t->stream);
		t->ev_used[phase] = true;
	}
};
static void tb_collect_phases(ngp_testbed* t) {
	if (!t->profiling) return;
	for (int p = 0; p < NGP_N_PHASES; ++p) {
		if (!t->ev_used[p]) continue;
		float ms = 0.0f;
		cudaEventSynchronize(t->ev[p][1]);
		if (cudaEventElapsedTime(&ms, t->ev[p][0], t->ev[p][1]) == cudaSuccess) t->phase_ms[p] += ms;
		t->ev_used[p] = false;
	}
	++t->phase_steps;
}

static uint32_t tb_n_views(const ngp_testbed* t) { return std::min(t->n_images_for_training, t->n_images); }

// training_prep_nerf (testbed_nerf.cu:3385-3398)
static void tb_training_prep(ngp_testbed* t) {
	tb_upload_views(t);
	uint64_t state = t->density_grid_rng.state;
	update_density_grid(t->desc, t->stream, t->cfg, t->params.p, &state, t->density_grid_rng.inc, t->training_step, t->density_grid_ema_step,
		t->density_grid_decay, t->views_dev.p, tb_n_views(t), t->density_grid.p, t->bitfield.p, t->mean_density.p, t->grid_scratch.p);
	t->density_grid_rng.state = state;
	++t->density_grid_ema_step;
}

__global__ void k_sum_partial_1024(const float* __restrict__ data, uint32_t n, float* __restrict__ partial) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	float s = 0.0f;
	for (uint32_t i = t; i < n; i += 1024) s += data[i];
	partial[t] = s;
}

static uint32_t tb_max_inference(const ngp_testbed* t, uint32_t batch) {
	const uint32_t max_samples = batch * 16;
	if (t->measured_batch_size_before_compaction == 0) return max_samples;
	return next_multiple(std::min(t->measured_batch_size_before_compaction, max_samples), NGP_BATCH_GRANULARITY);
}
// What the sample generator may fill.  The reference sizes its inference launch on the host from the PREVIOUS step's sample count
// (max_inference, testbed_nerf.cu:3055-3061) and so drops every ray whose samples end beyond it: whenever a step generates more than
// the last one did, the rays that reserve their slots last — the longest — are lost.  Here the inference pass reads the count from the
// device counter and the only limit is the buffer (16 x batch, the reference's max_samples); `nerf.training.drop_overflowing_rays` = 1
// restores the reference's rule.
static uint32_t tb_generator_capacity(const ngp_testbed* t, uint32_t batch) { return t->drop_overflowing_rays ? tb_max_inference(t, batch) : batch * 16; }
// The training-time inference pass has two schedules.  Ray-ordered (k_nerf_forward_rays): a tile is 128 / chunk rays x chunk
// consecutive samples and a ray is walked chunk by chunk until the loss kernel would stop reading it — up to chunk - 1 evaluations
// per ray are wasted and a long ray takes many sequential tiles, but nothing behind the stopping point is evaluated.  Flat
// (k_nerf_forward over every generated sample, the reference's schedule, testbed_nerf.cu:3233-3235): full tiles, no dependence
// between them, but it evaluates everything.  Which one wins depends on how much of what the generator produced the loss kernel
// reads: the synthetic unit-cube scene reads 7 % (ray-ordered: 0.25 vs 1.3 ms), nerf/fox reads 55 % (flat: 0.12 ms, ray-ordered with
// 8 / 16 / 32-sample chunks: 0.33 / 0.21 / 0.15 ms; profiles/r2).  Both choices follow the previous step's counters; the outputs the
// loss kernel reads are bit-identical either way (tests/test_gpu_march.py).
static bool tb_full_inference(const ngp_testbed* t) {
	if (t->full_inference != 2) return t->full_inference == 1;
	if (t->measured_batch_size == 0) return false;
	return (float)t->measured_batch_size_before_compaction < 2.5f * (float)t->measured_batch_size;
}
static uint32_t tb_inference_chunk(const ngp_testbed* t) {
	if (t->inference_chunk) return t->inference_chunk;
	const float per_ray = t->rays_per_batch ? (float)t->measured_batch_size / (float)t->rays_per_batch : 0.0f;
	return per_ray >= 64.0f ? 32u : (per_ray >= 24.0f ? 16u : 8u);
}
static bool tb_prep_due(uint32_t training_step) {
	// Testbed::train (src/testbed.cu:4596-4614): density-grid prep every clamp(step/16, 1, 16) steps
	const uint32_t n_prep_to_skip = std::min(std::max(training_step / 16u, 1u), 16u);
	return training_step % n_prep_to_skip == 0;
}
// Any state change the prefetched generator launch could not have seen makes it void.
static void tb_invalidate_prefetch(ngp_testbed* t) {
	if (t->prefetch_valid && t->side_stream) cudaStreamSynchronize(t->side_stream);
	t->prefetch_valid = false;
}
// Rank r of W marches the global ray ids r, r + W, r + 2W, ... of a batch of W x rays_local rays: image_idx (nerf_device.cuh:598) maps
// consecutive ids to the same view, so a contiguous block per rank would give every rank a disjoint 1/W slice of the views (and a
// rank whose views look at empty space idles while the others march); interleaved, every rank draws from every view.
static void tb_launch_generator(ngp_testbed* t, cudaStream_t stream, uint32_t set, uint32_t rays_local, uint32_t max_inference) {
	ngp_testbed::RaySet& rs = t->set[set];
	NGPB_CUDA_CHECK(cudaMemsetAsync(rs.counters.p, 0, sizeof(ngp_nerf_counters), stream));
	ngp_nerf_train_cfg cfg = t->cfg;
	cfg.ray_stride = t->dp_world;
	generate_training_samples(stream, rays_local, t->dp_rank, rays_local * t->dp_world, t->rng.state, t->rng.inc, cfg, t->views_dev.p, tb_n_views(t),
		t->bitfield.p, max_inference, rs.counters.p, rs.ray_indices.p, rs.rays.p, rs.numsteps.p, rs.coords.p);
}

// Replacement frames handed over by ngp_testbed_update_image_async become visible to `stream` here: wait for each frame's staging copy,
// move it into the view's pixel buffer, release the staging buffer.
static void tb_apply_pending_images(ngp_testbed* t, cudaStream_t stream) {
	for (const ngp_testbed::PendingImage& pi : t->pending_images) {
		ngp_testbed::ImageStaging& st = t->staging[pi.slot];
		NGPB_CUDA_CHECK(cudaStreamWaitEvent(stream, st.copied, 0));
		NGPB_CUDA_CHECK(cudaMemcpyAsync(t->pixel_bufs[pi.idx], st.buf, pi.bytes, cudaMemcpyDeviceToDevice, stream));
		NGPB_CUDA_CHECK(cudaEventRecord(st.consumed, stream));
		st.in_flight = false;
	}
	t->pending_images.clear();
}

// Per-image exposure (testbed_nerf.cu:979, 1142-1155, 2962-3000).  The loss kernel multiplies a view's colour by 2^exposure and, while
// `optimize_exposure` is on, accumulates the reference's gradient expression per view; every n_steps_between_cam_updates steps the host
// takes one Adam step per view with the network optimizer's current learning rate and re-centres the exposures on a zero mean.
static void tb_exposure_begin_step(ngp_testbed* t) {
	const uint32_t n = tb_n_views(t);
	const bool active = t->optimize_exposure || t->cam_exposure_nonzero;
	if (!active) {
		t->cfg.cam_exposure = nullptr;
		t->cfg.cam_exposure_gradient = nullptr;
		return;
	}
	NGPB_CHECK(t->dp_world == 1, "per-image exposure is not exchanged between ranks: optimize_exposure needs a single process");
	if (t->cam_exposure.size() != t->n_images) {
		t->cam_exposure.assign(t->n_images, ngp_testbed::ExposureAdam{});
		t->cam_exposure_nonzero = false;
		t->cam_exposure_dev.release();
	}
	if (t->cam_exposure_dev.n < (size_t)t->n_images * 3) {
		t->cam_exposure_dev.release();
		t->cam_exposure_gradient_dev.release();
		t->cam_exposure_dev.ensure_zeroed((size_t)t->n_images * 3);
		t->cam_exposure_gradient_dev.ensure_zeroed((size_t)t->n_images * 3);
		std::vector<float> host((size_t)t->n_images * 3);
		for (uint32_t i = 0; i < t->n_images; ++i)
			for (int c = 0; c < 3; ++c) host[(size_t)i * 3 + c] = t->cam_exposure[i].variable[c];
		NGPB_CUDA_CHECK(cudaMemcpyAsync(t->cam_exposure_dev.p, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, t->stream));
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	}
	t->cfg.cam_exposure = t->cam_exposure_dev.p;
	t->cfg.cam_exposure_gradient = t->optimize_exposure ? t->cam_exposure_gradient_dev.p : nullptr;
	// train_nerf (:2731-2735): the accumulators start from zero after every camera update
	if (t->optimize_exposure && t->n_steps_since_cam_update == 0)
		NGPB_CUDA_CHECK(cudaMemsetAsync(t->cam_exposure_gradient_dev.p, 0, sizeof(float) * 3 * (size_t)n, t->stream));
}

// after the optimizer step (train_nerf, :2878-3003)
static void tb_exposure_end_step(ngp_testbed* t) {
	t->n_steps_since_cam_update += 1;
	if (!t->optimize_exposure || t->n_steps_since_cam_update < t->n_steps_between_cam_updates) return;
	const uint32_t n = tb_n_views(t);
	std::vector<float> grad((size_t)n * 3);
	NGPB_CUDA_CHECK(cudaMemcpyAsync(grad.data(), t->cam_exposure_gradient_dev.p, grad.size() * sizeof(float), cudaMemcpyDeviceToHost, t->stream));
	NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	const float per_camera_loss_scale = (float)n / t->cfg.loss_scale / (float)t->n_steps_between_cam_updates;
	const float learning_rate = t->opt.learning_rate * t->lr_factor;   // m_optimizer->learning_rate(): the nested Adam's, after the decay
	const float beta1 = 0.9f, beta2 = 0.99f, epsilon = 1e-8f;
	float mean[3] = {0, 0, 0};
	for (uint32_t i = 0; i < n; ++i) {
		ngp_testbed::ExposureAdam& a = t->cam_exposure[i];
		++a.iter;
		const float actual_lr = learning_rate * std::sqrt(1.0f - std::pow(beta2, (float)a.iter)) / (1.0f - std::pow(beta1, (float)a.iter));
		for (int c = 0; c < 3; ++c) {
			const float g = grad[(size_t)i * 3 + c] * per_camera_loss_scale + a.variable[c] * t->exposure_l2_reg;
			a.first_moment[c] = beta1 * a.first_moment[c] + (1.0f - beta1) * g;
			a.second_moment[c] = beta2 * a.second_moment[c] + (1.0f - beta2) * g * g;
			a.variable[c] -= actual_lr * a.first_moment[c] / (std::sqrt(a.second_moment[c]) + epsilon);
			mean[c] += a.variable[c];
		}
	}
	for (int c = 0; c < 3; ++c) mean[c] /= (float)n;
	std::vector<float> host((size_t)t->n_images * 3, 0.0f);
	for (uint32_t i = 0; i < t->n_images; ++i)
		for (int c = 0; c < 3; ++c) {
			if (i < n) t->cam_exposure[i].variable[c] -= mean[c];   // renormalise: zero mean over the training views
			host[(size_t)i * 3 + c] = t->cam_exposure[i].variable[c];
		}
	NGPB_CUDA_CHECK(cudaMemcpyAsync(t->cam_exposure_dev.p, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, t->stream));
	NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));   // `host` goes out of scope
	t->cam_exposure_nonzero = true;
	t->n_steps_since_cam_update = 0;
	tb_invalidate_prefetch(t);
}

// A training step (train_nerf_step, testbed_nerf.cu:3007-3382 + optimizer_step :2770) is issued in three parts so that a
// data-parallel caller can put its collectives where they belong and so that the next step's sample generation overlaps them:
//   front  occupancy-grid upkeep, ray generation (or the prefetched one), ray-ordered inference, loss + compaction.
//          Data parallel: the caller now sums the 16-byte counter block over ranks.
//   back   forward/backward kernel queued; while it runs the host reads the counters, updates the rays_per_batch controller
//          (NerfCounters::update_after_training, :2678-2702) and queues the NEXT step's generator on the side stream, gated on
//          the forward/backward kernel.  Data parallel: the caller now all-reduces the gradients — the generator (0.5 ms) runs
//          beside the all-reduce and the optimizer instead of after them.
//   apply  optimizer.
// Everything is asynchronous; the only host wait is for the counters (ready before the forward/backward kernel starts).
static void tb_front(ngp_testbed* t, uint32_t batch) {
	NGPB_CHECK(t->has_network, "Testbed::train: no network (call reload_network_from_json/file first)");
	NGPB_CHECK(tb_n_views(t) > 0, "Testbed::train: no training images (set nerf.training.n_images_for_training)");
	NGPB_CHECK(batch % NGP_BATCH_GRANULARITY == 0 && batch > 0, "Testbed::train: batch size must be a positive multiple of 256");
	NGPB_CHECK(!t->grads_pending && !t->front_done, "training step parts out of order: front called twice");
	for (uint32_t i = 0; i < tb_n_views(t); ++i) NGPB_CHECK(t->views[i].pixels != nullptr, "Testbed::train: training image " + std::to_string(i) + " was never set");
	if (batch != t->last_batch) tb_invalidate_prefetch(t);  // buffers may be re-allocated
	tb_ensure_step_scratch(t, batch);
	tb_upload_views(t);

	tb_exposure_begin_step(t);
	const bool prep = tb_prep_due(t->training_step);
	if (prep) {
		tb_invalidate_prefetch(t);  // the bitfield is about to change (a prefetch is never issued for such a step anyway)
		PhaseTimer pt(t, 0);
		tb_training_prep(t);
	}
	if (t->training_step == 0) t->n_rays_total = 0;
	const uint32_t rays_local = t->rays_per_batch;
	const uint32_t rays_global = rays_local * t->dp_world;
	if (t->measured_batch_size_before_compaction == 0) t->measured_batch_size_before_compaction = tb_max_inference(t, batch);
	const uint32_t max_inference = tb_generator_capacity(t, batch);
	t->n_rays_total += rays_global;

	const bool use_prefetch = t->prefetch_valid && t->prefetch_step == t->training_step && t->prefetch_batch == batch && t->prefetch_rays == rays_local &&
		t->prefetch_max_inference == max_inference && t->prefetch_rng_state == t->rng.state;
	if (use_prefetch) {
		// the generator of this step already ran (or is finishing) on the side stream into the other buffer set
		t->cur ^= 1u;
		NGPB_CUDA_CHECK(cudaStreamWaitEvent(t->stream, t->ev_prefetch_done, 0));
	} else {
		tb_invalidate_prefetch(t);
		PhaseTimer pt(t, 1);
		tb_launch_generator(t, t->stream, t->cur, rays_local, max_inference);
	}
	t->prefetch_valid = false;
	ngp_testbed::RaySet& rs = t->set[t->cur];
	NGPB_CUDA_CHECK(cudaMemsetAsync(t->loss_per_ray.p, 0, sizeof(float) * rays_local, t->stream));
	{
		PhaseTimer pt(t, 2);
		if (tb_full_inference(t)) {
			// the reference's schedule: evaluate every generated sample (testbed_nerf.cu:3233-3235)
			nerf_inference_counted(t->desc, t->stream, max_inference, &rs.counters.p->n_samples, rs.coords.p, t->params.p, t->mlp_out.p);
		} else {
			// evaluate, ray by ray, only the samples the loss kernel will read (bit-identical outputs for those)
			nerf_inference_rays(t->desc, t->stream, rays_local, rs.counters.p, &rs.counters.p->pad, rs.numsteps.p, rs.coords.p, t->params.p,
				t->cfg.density_activation, t->mlp_out.p, tb_inference_chunk(t), t->cfg.train_mode);
		}
	}
	tb_apply_pending_images(t, t->stream);   // frames streamed in by update_image_async: first read by the loss kernel
	{
		PhaseTimer pt(t, 3);
		compute_loss(t->stream, rays_local, rays_global, t->rng.state, t->rng.inc, t->cfg, t->views_dev.p, tb_n_views(t), t->mlp_out.p, batch, rs.counters.p,
			rs.ray_indices.p, rs.rays.p, rs.numsteps.p, rs.coords.p, t->coords_compacted.p, t->dloss.p, t->loss_per_ray.p, t->mean_density.p);
		fill_rollover(t->stream, batch, rs.counters.p, t->coords_compacted.p, t->dloss.p);
	}
	t->get_loss_pending = (t->training_step % 16 == 0);
	if (t->get_loss_pending) {
		k_sum_partial_1024<<<4, 256, 0, t->stream>>>(t->loss_per_ray.p, rays_local, t->reduce_scratch.p);
		NGPB_LAUNCHED();
		NGPB_CUDA_CHECK(cudaMemcpyAsync(t->readback->loss_partial, t->reduce_scratch.p, sizeof(float) * 1024, cudaMemcpyDeviceToHost, t->stream));
	}
	if (t->dp_world > 1) NGPB_CUDA_CHECK(cudaMemcpyAsync(t->dp_counters.p, rs.counters.p, sizeof(ngp_nerf_counters), cudaMemcpyDeviceToDevice, t->stream));
	t->step_rays_local = rays_local;
	t->step_max_inference = max_inference;
	t->step_batch = batch;
	t->front_done = true;
	t->controller_done = false;
}

static void tb_update_controller(ngp_testbed* t);
static void tb_prefetch(ngp_testbed* t);

// The counter block (and, every 16th step, the loss partials) summed over the ranks on the communication stream, beside the
// forward/backward kernel: the rays_per_batch controller (NerfCounters::update_after_training) sees the same numbers on every rank,
// and nothing of it is on the training stream's critical path.
static void tb_exchange_counters(ngp_testbed* t) {
	const ngpb::NcclApi& nccl = ngpb::NcclApi::get();
	NGPB_CUDA_CHECK(cudaEventRecord(t->ev_counters_ready, t->stream));
	NGPB_CUDA_CHECK(cudaStreamWaitEvent(t->comm_stream, t->ev_counters_ready, 0));
	nccl.check(nccl.AllReduce(t->dp_counters.p, t->dp_counters.p, 4, ngpb::NcclApi::ncclUint32, ngpb::NcclApi::ncclSum, t->comm_small, t->comm_stream), "ncclAllReduce(counters)");
	NGPB_CUDA_CHECK(cudaMemcpyAsync(&t->readback->counters, t->dp_counters.p, sizeof(ngp_nerf_counters), cudaMemcpyDeviceToHost, t->comm_stream));
	if (t->get_loss_pending) {
		nccl.check(nccl.AllReduce(t->reduce_scratch.p, t->reduce_scratch.p, 1024, ngpb::NcclApi::ncclFloat32, ngpb::NcclApi::ncclSum, t->comm_small, t->comm_stream), "ncclAllReduce(loss)");
		NGPB_CUDA_CHECK(cudaMemcpyAsync(t->readback->loss_partial, t->reduce_scratch.p, sizeof(float) * 1024, cudaMemcpyDeviceToHost, t->comm_stream));
	}
	NGPB_CUDA_CHECK(cudaEventRecord(t->ev_front_done, t->comm_stream));
}

// `early_controller`: read the counters and prefetch now (requires that, data parallel, the counter block has been summed — by the
// caller, or by tb_exchange_counters when the library owns the communicators)
static void tb_back(ngp_testbed* t, bool early_controller) {
	NGPB_CHECK(t->front_done, "training step parts out of order: back without front");
	ngp_testbed::RaySet& rs = t->set[t->cur];
	if (t->dp_world > 1 && t->comm_small && early_controller) {
		tb_exchange_counters(t);
	} else if (t->dp_world == 1 || early_controller) {
		NGPB_CUDA_CHECK(cudaMemcpyAsync(&t->readback->counters, t->dp_world == 1 ? rs.counters.p : t->dp_counters.p, sizeof(ngp_nerf_counters),
			cudaMemcpyDeviceToHost, t->stream));
		NGPB_CUDA_CHECK(cudaEventRecord(t->ev_front_done, t->stream));
	}
	{
		PhaseTimer pt(t, 4);
		nerf_forward_backward(t->desc, t->stream, t->step_batch, t->coords_compacted.p, t->params.p, t->dloss.p, t->grads.p, t->mlp_grads_f32.p, nullptr);
	}
	NGPB_CUDA_CHECK(cudaEventRecord(t->ev_back_done, t->stream));
	t->rng.advance();
	t->front_done = false;
	t->grads_pending = true;
	if (t->dp_world == 1 || early_controller) {
		tb_update_controller(t);
		tb_prefetch(t);
	}
}

// NerfCounters::update_after_training (testbed_nerf.cu:2678-2702).  The counters were copied out before the forward/backward
// kernel was queued: the wait ends while the GPU still works.
static void tb_update_controller(ngp_testbed* t) {
	const uint32_t batch = t->step_batch;
	NGPB_CUDA_CHECK(cudaEventSynchronize(t->ev_front_done));
	const ngp_nerf_counters c = t->readback->counters;
	t->controller_done = true;
	const uint32_t n_samples = c.n_samples / t->dp_world, n_compacted = c.n_samples_compacted / t->dp_world;
	t->measured_batch_size = 0;
	t->measured_batch_size_before_compaction = 0;
	if (n_samples == 0 || n_compacted == 0) {
		t->loss_scalar = 0.0f;
		t->shall_train = false;  // "Nerf training generated 0 samples. Aborting training."
		return;
	}
	t->measured_batch_size_before_compaction = n_samples;
	t->measured_batch_size = n_compacted;
	if (t->get_loss_pending) {
		float sum = 0.0f;
		for (int i = 0; i < 1024; ++i) sum += t->readback->loss_partial[i];
		t->loss_scalar = sum * (float)t->measured_batch_size / (float)batch;
	}
	t->rays_per_batch = (uint32_t)((float)t->rays_per_batch * (float)batch / (float)t->measured_batch_size);
	t->rays_per_batch = std::min(next_multiple(t->rays_per_batch, NGP_BATCH_GRANULARITY), 1u << 18);
}

// The generator of the NEXT step on the side stream.  It depends on nothing the forward/backward kernel or the optimizer write
// (rays, occupancy bits, views), so it may run beside the optimizer and — data parallel — beside the gradient all-reduce.  It is
// gated on the forward/backward kernel: run beside THAT kernel it only slowed the step down (k_nerf_train owns the SMs'
// registers and shared memory: 1.90 vs 1.68 ms, profiles/r1a).
static void tb_prefetch(ngp_testbed* t) {
	const uint32_t next_step = t->training_step + 1;
	// Single GPU: since the generator marches with G lanes per ray it takes 0.34 ms on nerf/fox, and running it beside the optimizer
	// (both memory-latency bound, the optimizer streaming 340 MB) costs more than it hides: 1.50 vs 1.41 ms/step (profiles/r2).  Data
	// parallel it runs beside the gradient all-reduce, where the SMs are idle.
	const bool worth_it = t->overlap_sample_generation == 1 || (t->overlap_sample_generation == 2 && t->dp_world > 1);
	if (!(worth_it && t->shall_train && !tb_prep_due(next_step) && !t->views_dirty && !t->profiling)) return;
	const uint32_t batch = t->step_batch;
	const uint32_t next = t->cur ^ 1u;
	const uint32_t max_inference = tb_generator_capacity(t, batch);
	// the other buffer set was last read by the loss kernel of the previous step, which precedes ev_back_done on the main stream;
	// the bitfield and the views are not written by anything in flight
	NGPB_CUDA_CHECK(cudaStreamWaitEvent(t->side_stream, t->ev_back_done, 0));
	tb_launch_generator(t, t->side_stream, next, t->rays_per_batch, max_inference);
	NGPB_CUDA_CHECK(cudaEventRecord(t->ev_prefetch_done, t->side_stream));
	t->prefetch_valid = true;
	t->prefetch_step = next_step;
	t->prefetch_batch = batch;
	t->prefetch_rays = t->rays_per_batch;
	t->prefetch_max_inference = max_inference;
	t->prefetch_rng_state = t->rng.state;
}

// optimizer_step (testbed_nerf.cu:2770-2788).  Data parallel without the early controller: the caller has summed the counter
// block over ranks by now, so the controller update happens here.
static void tb_apply_grads(ngp_testbed* t) {
	NGPB_CHECK(t->grads_pending, "train_apply_grads without train_compute_grads / train_back");
	if (!t->controller_done) {
		NGPB_CUDA_CHECK(cudaMemcpyAsync(&t->readback->counters, t->dp_counters.p, sizeof(ngp_nerf_counters), cudaMemcpyDeviceToHost, t->stream));
		NGPB_CUDA_CHECK(cudaEventRecord(t->ev_front_done, t->stream));
	}
	const ngp_adam_cfg a = next_adam_cfg(t->opt, t->optimizer_step, t->lr_factor, t->cfg.loss_scale, t->train_network, t->train_encoding);
	{
		PhaseTimer pt(t, 5);
		optimizer_step(t->desc, t->stream, a, t->params_fp32.p, t->params.p, t->params_ema.p, t->grads.p, t->m1.p, t->m2.p, t->param_steps.p);
	}
	if (!t->controller_done) tb_update_controller(t);
	++t->training_step;
	t->grads_pending = false;
	tb_exposure_end_step(t);
	if (t->profiling) {
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));  // phase timing wants completed events; only paid while profiling
		tb_collect_phases(t);
	}
}

// front + back without the early controller: the (data-parallel) caller sums counters and gradients, then calls apply
static void tb_compute_grads(ngp_testbed* t, uint32_t batch) {
	tb_front(t, batch);
	tb_back(t, t->dp_world == 1);
}

}  // namespace ngpb

// ------------------------------------------------------------------------------------------------------------------
// extern "C"
// ------------------------------------------------------------------------------------------------------------------
extern "C" {

const char* ngp_last_error(void) { return g_last_error.c_str(); }
int ngp_version(void) { return 1; }
int ngp_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}
uint64_t ngp_launch_count(void) { return g_launch_count; }

int ngp_grid_desc_init(ngp_grid_desc* g, uint32_t n_levels, uint32_t F, uint32_t log2_T, uint32_t base_res, float pls, uint32_t aabb_scale) {
	NGPB_TRY(grid_desc_init(g, n_levels, F, log2_T, base_res, pls, aabb_scale));
}
int ngp_nerf_desc_init(ngp_nerf_desc* d, const ngp_grid_desc* g, uint32_t nhd, uint32_t nhr) { NGPB_TRY(nerf_desc_init(d, g, nhd, nhr)); }
int ngp_march_consts_init(ngp_march_consts* m, float cone_angle) { NGPB_TRY(march_consts_init(m, cone_angle)); }
int ngp_nerf_init_params_host(const ngp_nerf_desc* d, uint64_t seed, float* out) { NGPB_TRY(nerf_init_params_host(d, seed, out)); }


int ngp_nerf_inference(const ngp_nerf_desc* d, void* stream, uint32_t n, const float* coords, const void* params, void* out, uint32_t out_stride) {
	NGPB_TRY(require_device(); nerf_inference(*d, (cudaStream_t)stream, n, coords, (const __half*)params, (__half*)out, out_stride));
}
int ngp_nerf_inference_rays(const ngp_nerf_desc* d, void* stream, uint32_t n_rays_max, const ngp_nerf_counters* counters, uint32_t* queue, const uint32_t* numsteps,
	const float* coords, const void* params, uint32_t density_activation, void* out, uint32_t train_mode) {
	NGPB_TRY(require_device(); nerf_inference_rays(*d, (cudaStream_t)stream, n_rays_max, counters, queue, numsteps, coords, (const __half*)params, density_activation,
		(__half*)out, 8, train_mode));
}
int ngp_nerf_density(const ngp_nerf_desc* d, void* stream, uint32_t n, const float* positions, uint32_t pos_stride, const void* params, void* out) {
	NGPB_TRY(require_device(); nerf_density(*d, (cudaStream_t)stream, n, positions, pos_stride, (const __half*)params, (__half*)out));
}
int ngp_grid_encode(const ngp_grid_desc* g, void* stream, uint32_t n, const float* positions, uint32_t pos_stride, const void* grid, void* out) {
	NGPB_TRY(require_device(); grid_encode(*g, (cudaStream_t)stream, n, positions, pos_stride, (const __half*)grid, (__half*)out));
}
int ngp_nerf_forward_backward(const ngp_nerf_desc* d, void* stream, uint32_t n, const float* coords, const void* params, const void* dL_dout, void* grads,
	void* out) {
	NGPB_TRY({
		require_device();
		// transient fp32 accumulator for the MLP weight gradients
		float* tmp = nullptr;
		NGPB_CUDA_CHECK(cudaMallocAsync(&tmp, d->n_mlp_params * sizeof(float), (cudaStream_t)stream));
		NGPB_CUDA_CHECK(cudaMemsetAsync(tmp, 0, d->n_mlp_params * sizeof(float), (cudaStream_t)stream));
		nerf_forward_backward(*d, (cudaStream_t)stream, n, coords, (const __half*)params, (const __half*)dL_dout, (__half*)grads, tmp, (__half*)out);
		NGPB_CUDA_CHECK(cudaFreeAsync(tmp, (cudaStream_t)stream));
	});
}
void ngp_set_scatter_aggregation(int mode) { set_scatter_aggregation(mode); }
int ngp_profile_mlp_phase(const ngp_nerf_desc* d, void* stream, uint32_t n, const float* coords, const void* params, const void* dL_dout, void* grads, float* mlp_scratch_f32) {
	NGPB_TRY(require_device(); profile_mlp_phase(*d, (cudaStream_t)stream, n, coords, (const __half*)params, (const __half*)dL_dout, (__half*)grads, mlp_scratch_f32));
}
int ngp_optimizer_step(const ngp_nerf_desc* d, void* stream, const ngp_adam_cfg* cfg, float* p32, void* p16, void* ema, void* grads, float* m1, float* m2,
	uint32_t* steps) {
	NGPB_TRY(require_device(); optimizer_step(*d, (cudaStream_t)stream, *cfg, p32, (__half*)p16, (__half*)ema, (__half*)grads, m1, m2, steps));
}
int ngp_nerf_generate_training_samples(void* stream, uint32_t n_rays, uint32_t ray_offset, uint32_t n_rays_global, uint64_t rng_state, uint64_t rng_inc,
	const ngp_nerf_train_cfg* cfg, const ngp_train_view* views, uint32_t n_views, const uint8_t* bitfield, uint32_t max_samples, ngp_nerf_counters* counters,
	uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords) {
	NGPB_TRY(require_device(); const uint32_t stride = cfg->ray_stride ? cfg->ray_stride : 1u;
		NGPB_CHECK(n_rays_global >= n_rays && (n_rays == 0 || (uint64_t)ray_offset + (uint64_t)(n_rays - 1) * stride < n_rays_global), "ray shard outside the global batch");
		NGPB_CHECK(cfg->math_mode <= NGP_MATH_REFERENCE, "ngp_nerf_train_cfg.math_mode: unknown arithmetic flavour");
		generate_training_samples((cudaStream_t)stream, n_rays, ray_offset, n_rays_global, rng_state, rng_inc, *cfg, views, n_views, bitfield,
		max_samples, counters, ray_indices, rays, numsteps, coords));
}
int ngp_nerf_compute_loss(void* stream, uint32_t n_rays, uint32_t n_rays_global, uint64_t rng_state, uint64_t rng_inc, const ngp_nerf_train_cfg* cfg,
	const ngp_train_view* views, uint32_t n_views, const void* network_output, uint32_t max_compacted, ngp_nerf_counters* counters,
	const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords, float* coords_compacted, void* dloss, float* loss_per_ray,
	const float* mean_density) {
	NGPB_TRY(require_device(); compute_loss((cudaStream_t)stream, n_rays, n_rays_global, rng_state, rng_inc, *cfg, views, n_views, (const __half*)network_output,
		max_compacted, counters, ray_indices, rays, numsteps, coords, coords_compacted, (__half*)dloss, loss_per_ray, mean_density));
}
int ngp_nerf_fill_rollover(void* stream, uint32_t target_batch, const ngp_nerf_counters* counters, float* coords_compacted, void* dloss) {
	NGPB_TRY(require_device(); fill_rollover((cudaStream_t)stream, target_batch, counters, coords_compacted, (__half*)dloss));
}
size_t ngp_nerf_density_grid_scratch_bytes(uint32_t max_cascade) { return density_grid_scratch_bytes(max_cascade); }
int ngp_nerf_update_density_grid(const ngp_nerf_desc* d, void* stream, const ngp_nerf_train_cfg* cfg, const void* params, uint64_t* rng_state,
	uint64_t rng_inc, uint32_t training_step, uint32_t ema_step, float decay, const ngp_train_view* views, uint32_t n_views, float* density_grid,
	uint8_t* bitfield, float* mean_density, void* scratch) {
	NGPB_TRY(require_device(); update_density_grid(*d, (cudaStream_t)stream, *cfg, (const __half*)params, rng_state, rng_inc, training_step, ema_step, decay,
		views, n_views, density_grid, bitfield, mean_density, scratch));
}
int ngp_nerf_update_bitfield(void* stream, uint32_t max_cascade, const float* density_grid, uint8_t* bitfield, float* mean_density) {
	NGPB_TRY({
		require_device();
		float* partial = nullptr;
		NGPB_CUDA_CHECK(cudaMallocAsync(&partial, 1024 * sizeof(float), (cudaStream_t)stream));
		update_bitfield((cudaStream_t)stream, max_cascade, density_grid, bitfield, mean_density, partial);
		NGPB_CUDA_CHECK(cudaFreeAsync(partial, (cudaStream_t)stream));
	});
}
size_t ngp_nerf_render_scratch_bytes(int32_t width, int32_t rows) { return render_scratch_bytes(width, rows); }
int ngp_nerf_render(const ngp_nerf_desc* d, void* stream, const ngp_render_cfg* cfg, int32_t y0, int32_t y1, const void* params, const uint8_t* bitfield,
	float* rgba, float* depth, void* scratch, uint32_t* n_steps_total) {
	NGPB_TRY(require_device(); render_nerf(*d, (cudaStream_t)stream, *cfg, y0, y1, (const __half*)params, bitfield, rgba, depth, scratch, n_steps_total));
}

// ---- B2 ----------------------------------------------------------------------------------------------------------
ngp_testbed* ngp_testbed_create(int device, void* stream) {
	try {
		require_device();
		NGPB_CUDA_CHECK(cudaSetDevice(device));
		auto* t = new ngp_testbed();
		t->device = device;
		t->stream = (cudaStream_t)stream;
		tb_set_defaults(t);
		tb_update_scene(t);
		return t;
	} catch (const std::exception& e) {
		set_last_error(e.what());
		return nullptr;
	}
}
void ngp_testbed_destroy(ngp_testbed* t) {
	if (!t) return;
	cudaStreamSynchronize(t->stream);
	delete t;
}
int ngp_testbed_create_empty_nerf_dataset(ngp_testbed* t, uint32_t n_images, uint32_t aabb_scale) {
	NGPB_TRY({
		tb_invalidate_prefetch(t);
		tb_apply_pending_images(t, t->stream);
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));   // the pixel buffers are freed below
		NGPB_CHECK(n_images > 0, "create_empty_nerf_dataset: n_images must be > 0");
		NGPB_CHECK(aabb_scale >= 1 && aabb_scale <= 128 && (aabb_scale & (aabb_scale - 1)) == 0, "aabb_scale must be a power of two in [1, 128]");
		for (void* p : t->pixel_bufs)
			if (p) cudaFree(p);
		t->n_images = n_images;
		t->aabb_scale = aabb_scale;
		t->n_images_for_training = 0;  // Testbed::create_empty_nerf_dataset, testbed_nerf.cu:2349
		t->views.assign(n_images, ngp_train_view{});
		t->pixel_bufs.assign(n_images, nullptr);
		for (auto& v : t->views) {
			v.focal_x = v.focal_y = 1000.0f;
			v.principal_x = v.principal_y = 0.5f;
			const float ident[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
			memcpy(v.xform, ident, sizeof(ident));
		}
		t->views_dirty = true;
		tb_update_scene(t);
	});
}
int ngp_testbed_set_image(ngp_testbed* t, uint32_t idx, const float* rgba_host, int32_t w, int32_t h) {
	NGPB_TRY({
		tb_invalidate_prefetch(t);
		if (!t->pending_images.empty()) {   // a streamed-in frame still on its way: land it before the buffer it targets may be freed
			tb_apply_pending_images(t, t->stream);
			NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
		}
		NGPB_CHECK(idx < t->n_images, "Invalid frame index");
		NGPB_CHECK(w > 0 && h > 0, "image must have positive size");
		if (t->pixel_bufs[idx]) cudaFree(t->pixel_bufs[idx]);
		void* p = nullptr;
		const size_t bytes = (size_t)w * h * 4 * sizeof(float);
		NGPB_CUDA_CHECK(cudaMalloc(&p, bytes));
		NGPB_CUDA_CHECK(cudaMemcpy(p, rgba_host, bytes, cudaMemcpyHostToDevice));
		t->pixel_bufs[idx] = p;
		t->views[idx].pixels = p;
		t->views[idx].image_type = NGP_IMAGE_FLOAT;  // python_api.cu:45-72 keeps float images as they are
		{
			bool any_masked = false;  // read_rgba(...).x < 0 marks a masked-away pixel (testbed_nerf.cu:732-736)
			const size_t n_px = (size_t)w * h;
			for (size_t k = 0; k < n_px && !any_masked; ++k) any_masked = rgba_host[k * 4] < 0.0f;
			t->views[idx].no_mask = any_masked ? 0u : 1u;
		}
		t->views[idx].width = w;
		t->views[idx].height = h;
		t->views_dirty = true;
	});
}
int ngp_testbed_set_image_bytes(ngp_testbed* t, uint32_t idx, const uint8_t* rgba8_host, int32_t w, int32_t h) {
	NGPB_TRY({
		tb_invalidate_prefetch(t);
		if (!t->pending_images.empty()) {   // a streamed-in frame still on its way: land it before the buffer it targets may be freed
			tb_apply_pending_images(t, t->stream);
			NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
		}
		NGPB_CHECK(idx < t->n_images, "Invalid frame index");
		NGPB_CHECK(w > 0 && h > 0, "image must have positive size");
		if (t->pixel_bufs[idx]) cudaFree(t->pixel_bufs[idx]);
		void* p = nullptr;
		const size_t bytes = (size_t)w * h * 4;
		NGPB_CUDA_CHECK(cudaMalloc(&p, bytes));
		NGPB_CUDA_CHECK(cudaMemcpy(p, rgba8_host, bytes, cudaMemcpyHostToDevice));
		t->pixel_bufs[idx] = p;
		t->views[idx].pixels = p;
		t->views[idx].image_type = NGP_IMAGE_BYTE;
		{
			bool any_masked = false;  // MASK_COLOR 0x00FF00FF marks a masked-away pixel (common_device.cuh:698-707)
			const uint32_t* px = reinterpret_cast<const uint32_t*>(rgba8_host);
			const size_t n_px = (size_t)w * h;
			for (size_t k = 0; k < n_px && !any_masked; ++k) any_masked = px[k] == 0x00FF00FFu;
			t->views[idx].no_mask = any_masked ? 0u : 1u;
		}
		t->views[idx].width = w;
		t->views[idx].height = h;
		t->views_dirty = true;
	});
}
void ngp_render_pixel_offset(uint32_t sample_index, float* offset_xy) { render_pixel_offset(sample_index, offset_xy); }
int ngp_testbed_set_camera_extrinsics(ngp_testbed* t, uint32_t idx, const float* m, int convert_to_ngp) {
	NGPB_TRY({
		tb_invalidate_prefetch(t);
		NGPB_CHECK(idx < t->n_images, "Invalid frame index");
		// input: row-major 3x4 camera-to-world; columns c0,c1,c2,origin
		float col[4][3];
		for (int c = 0; c < 4; ++c)
			for (int r = 0; r < 3; ++r) col[c][r] = m[r * 4 + c];
		if (convert_to_ngp) {
			// nerf_matrix_to_ngp (nerf_loader.h:101-120)
			for (int r = 0; r < 3; ++r) {
				col[1][r] *= -1.0f;
				col[2][r] *= -1.0f;
				col[3][r] = col[3][r] * t->scene_scale + t->scene_offset[r];
			}
			// cycle axes xyz <- yzx (row 0 <- row 1, row 1 <- row 2, row 2 <- old row 0)
			for (int c = 0; c < 4; ++c) {
				const float tmp = col[c][0];
				col[c][0] = col[c][1];
				col[c][1] = col[c][2];
				col[c][2] = tmp;
			}
		}
		for (int c = 0; c < 4; ++c)
			for (int r = 0; r < 3; ++r) t->views[idx].xform[c * 3 + r] = col[c][r];
		t->views_dirty = true;
	});
}
int ngp_testbed_set_camera_intrinsics(ngp_testbed* t, uint32_t idx, float fx, float fy, float cx, float cy, float k1, float k2, float p1, float p2) {
	NGPB_TRY({
		tb_invalidate_prefetch(t);
		NGPB_CHECK(idx < t->n_images, "Invalid frame index");
		ngp_train_view& v = t->views[idx];
		NGPB_CHECK(v.width > 0, "set_camera_intrinsics: set the image first (principal point is relative to the resolution)");
		// testbed_nerf.cu:2151-2186
		if (fx <= 0.0f) fx = fy;
		if (fy <= 0.0f) fy = fx;
		cx = cx < 0.0f ? -cx : cx / (float)v.width;
		cy = cy < 0.0f ? -cy : cy / (float)v.height;
		v.lens_mode = NGP_LENS_PERSPECTIVE;
		memset(v.lens_params, 0, sizeof(v.lens_params));
		if (k1 != 0.0f || k2 != 0.0f || p1 != 0.0f || p2 != 0.0f) {
			v.lens_mode = NGP_LENS_OPENCV;
			v.lens_params[0] = k1; v.lens_params[1] = k2; v.lens_params[2] = p1; v.lens_params[3] = p2;
		}
		v.principal_x = cx;
		v.principal_y = cy;
		v.focal_x = fx;
		v.focal_y = fy;
		t->views_dirty = true;
	});
}
int ngp_testbed_reload_network_from_json(ngp_testbed* t, const char* json_text) {
	NGPB_TRY({
		tb_invalidate_prefetch(t);
		const std::string text(json_text);
		Json cfg = JsonParser(text).parse();
		tb_reset_network(t, cfg);
	});
}
int ngp_testbed_reload_network_from_file(ngp_testbed* t, const char* path) {
	NGPB_TRY({
		tb_invalidate_prefetch(t);
		const Json cfg = load_network_config_file(path);   // "parent" chains resolved (merge_parent_network_config, src/testbed.cu:86-97)
		tb_reset_network(t, cfg);
	});
}
int ngp_testbed_set_seed(ngp_testbed* t, uint64_t seed) {
	tb_invalidate_prefetch(t);
	t->seed = seed;
	return 0;
}
// Testbed::reset_network(clear_density_grid) (src/testbed.cu:4163-4178, python_api.cu:534 "reset"): the network of the current
// config re-initialised from the seed, optimizer and counters cleared; the occupancy grid is kept on request
int ngp_testbed_reset(ngp_testbed* t, int reset_density_grid) {
	NGPB_TRY({
		NGPB_CHECK(t->has_network, "reset: no network");
		tb_invalidate_prefetch(t);
		std::vector<float> kept;
		const size_t n_grid = (size_t)GRID_N_CELLS * (t->cfg.max_cascade + 1);
		if (!reset_density_grid) {
			kept.resize(n_grid);
			NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
			NGPB_CUDA_CHECK(cudaMemcpy(kept.data(), t->density_grid.p, n_grid * sizeof(float), cudaMemcpyDeviceToHost));
		}
		const Json cfg = t->network_config;
		tb_reset_network(t, cfg);
		if (!reset_density_grid && ngp_testbed_set_density_grid(t, kept.data(), (uint32_t)n_grid)) throw std::runtime_error(g_last_error);
	});
}
// one training view as the Testbed holds it (TrainingImageMetadata + TrainingXForm.start); `pixels` is a device pointer
int ngp_testbed_get_view(ngp_testbed* t, uint32_t idx, ngp_train_view* out) {
	NGPB_TRY({
		NGPB_CHECK(out != nullptr, "get_view: null output");
		NGPB_CHECK(idx < t->n_images, "get_view: image index out of range");
		*out = t->views[idx];
	});
}
int ngp_testbed_get_camera_exposure(ngp_testbed* t, uint32_t idx, float* rgb_out) {
	NGPB_TRY({
		NGPB_CHECK(idx < t->n_images, "image index out of range");
		for (int c = 0; c < 3; ++c) rgb_out[c] = idx < t->cam_exposure.size() ? t->cam_exposure[idx].variable[c] : 0.0f;
	});
}
int ngp_testbed_set_camera_exposure(ngp_testbed* t, uint32_t idx, const float* rgb) {
	NGPB_TRY({
		NGPB_CHECK(idx < t->n_images, "image index out of range");
		tb_invalidate_prefetch(t);
		if (t->cam_exposure.size() != t->n_images) t->cam_exposure.assign(t->n_images, ngp_testbed::ExposureAdam{});
		ngp_testbed::ExposureAdam a{};   // reset_camera_extrinsics-style: a set value starts a fresh optimizer state (testbed_nerf.cu:2207)
		for (int c = 0; c < 3; ++c) a.variable[c] = rgb[c];
		t->cam_exposure[idx] = a;
		t->cam_exposure_nonzero = true;
		t->cam_exposure_dev.release();   // re-uploaded by the next training step
		t->cam_exposure_gradient_dev.release();
	});
}
int ngp_testbed_set_option(ngp_testbed* t, const char* name_c, double value) {
	NGPB_TRY({
		tb_invalidate_prefetch(t);
		const std::string n(name_c);
		ngp_nerf_train_cfg& c = t->cfg;
		if (n == "nerf.training.n_images_for_training") { NGPB_CHECK(value >= 0 && value <= t->n_images, "n_images_for_training out of range"); t->n_images_for_training = (uint32_t)value; }
		else if (n == "nerf.training.random_bg_color") c.random_bg_color = value != 0;
		else if (n == "nerf.training.linear_colors") c.linear_colors = value != 0;
		else if (n == "nerf.training.snap_to_pixel_centers") c.snap_to_pixel_centers = value != 0;
		else if (n == "nerf.training.near_distance") c.near_distance = (float)value;
		else if (n == "nerf.training.loss_type") c.loss_type = (uint32_t)value;
		else if (n == "nerf.training.density_grid_decay") t->density_grid_decay = (float)value;
		else if (n == "nerf.rgb_activation") c.rgb_activation = (uint32_t)value;
		else if (n == "nerf.density_activation") c.density_activation = (uint32_t)value;
		else if (n == "nerf.cone_angle_constant") march_consts_init(&c.march, (float)value);
		else if (n == "nerf.render_min_transmittance") t->render_min_transmittance = (float)value;
		else if (n == "snap_to_pixel_centers") t->render_snap_to_pixel_centers = value != 0;
		else if (n == "render_with_lens_distortion") t->render_with_lens_distortion = value != 0;
		else if (n == "render_lens.mode") { NGPB_CHECK(value == NGP_LENS_PERSPECTIVE || value == NGP_LENS_OPENCV, "render_lens.mode: perspective or OpenCV"); t->render_lens_mode = (uint32_t)value; }
		else if (n.rfind("render_lens.params.", 0) == 0 && n.size() == 20 && n[19] >= '0' && n[19] <= '3') t->render_lens_params[n[19] - '0'] = (float)value;
		else if (n == "color_space") c.color_space = (uint32_t)value;
		else if (n == "background_color.r") c.background_color[0] = (float)value;
		else if (n == "background_color.g") c.background_color[1] = (float)value;
		else if (n == "background_color.b") c.background_color[2] = (float)value;
		else if (n == "background_color.a") t->background_alpha = (float)value;
		else if (n == "exposure") t->exposure = (float)value;
		else if (n == "nerf.training.math_mode") { NGPB_CHECK(value == 0 || value == 1, "math_mode must be 0 (deterministic) or 1 (reference)"); tb_invalidate_prefetch(t); t->cfg.math_mode = (uint32_t)value; }
		else if (n == "nerf.training.gen_walk_empty") { NGPB_CHECK(value >= 0 && value <= 1024, "gen_walk_empty: 0 (default) .. 1024"); tb_invalidate_prefetch(t); t->cfg.gen_walk_empty = (uint32_t)value; }
		else if (n == "nerf.training.gen_speculation") { NGPB_CHECK(value >= 0 && value <= 32, "gen_speculation: 0 (default) .. 32"); tb_invalidate_prefetch(t); t->cfg.gen_speculation = (uint32_t)value; }
		else if (n == "nerf.training.optimize_exposure") { t->optimize_exposure = value != 0; }
		else if (n == "nerf.training.exposure_l2_reg") { t->exposure_l2_reg = (float)value; }
		else if (n == "nerf.training.n_steps_between_cam_updates") { NGPB_CHECK(value >= 1 && value <= 65536, "n_steps_between_cam_updates: 1 .. 65536"); t->n_steps_between_cam_updates = (uint32_t)value; }
		else if (n == "nerf.training.compaction_order") { NGPB_CHECK(value == 0 || value == 1 || value == 2, "compaction_order: 0 (groups of 32 rays, shuffled), 1 (one atomic per ray), 2 (ray order)"); t->cfg.compaction_order = (uint32_t)value; }
		else if (n == "nerf.training.drop_overflowing_rays") { NGPB_CHECK(value == 0 || value == 1, "drop_overflowing_rays: 0 or 1"); tb_invalidate_prefetch(t); t->drop_overflowing_rays = (uint32_t)value; }
		else if (n == "nerf.training.gen_lanes_per_ray") { const uint32_t g = (uint32_t)value; NGPB_CHECK(g <= 32 && (g & (g - 1)) == 0, "gen_lanes_per_ray must be 0 or a power of two up to 32"); tb_invalidate_prefetch(t); t->cfg.gen_lanes_per_ray = g; }
		else if (n == "render_mode") { NGPB_CHECK(value == NGP_RENDER_SHADE || value == NGP_RENDER_AO || value == NGP_RENDER_POSITIONS || value == NGP_RENDER_DEPTH || value == NGP_RENDER_COST,
			"render_mode: this build renders Shade, AO, Positions, Depth and Cost"); t->render_mode = (uint32_t)value; }
		else if (n == "render_skips_per_tile") { NGPB_CHECK(value >= 0 && value <= 1024, "render_skips_per_tile out of range"); t->render_skips_per_tile = (uint32_t)value; }
		else if (n == "render_math") { NGPB_CHECK(value == 0 || value == 1, "render_math must be 0 (deterministic) or 1 (reference)"); t->render_math = (uint32_t)value; }
		else if (n == "nerf.training.full_inference") { NGPB_CHECK(value == 0 || value == 1 || value == 2, "full_inference: 0 ray-ordered, 1 every sample, 2 automatic"); tb_invalidate_prefetch(t); t->full_inference = (uint32_t)value; }
		else if (n == "nerf.training.train_mode") { NGPB_CHECK(value == 0 || value == 1 || value == 2, "train_mode must be Nerf (0), Rfl (1) or RflRelax (2)"); c.train_mode = (uint32_t)value; }
		else if (n == "nerf.training.inference_chunk") { NGPB_CHECK(value == 0 || value == 4 || value == 8 || value == 16 || value == 32, "inference_chunk must be 0 (automatic), 4, 8, 16 or 32"); t->inference_chunk = (uint32_t)value; }
		else if (n == "nerf.training.overlap_sample_generation") { NGPB_CHECK(value == 0 || value == 1 || value == 2, "overlap_sample_generation: 0 never, 1 always, 2 data parallel only"); tb_invalidate_prefetch(t); t->overlap_sample_generation = (uint32_t)value; }
		else if (n == "train_network") t->train_network = value != 0;
		else if (n == "train_encoding") t->train_encoding = value != 0;
		else if (n == "shall_train") t->shall_train = value != 0;
		else if (n == "nerf.training.dataset.scale") t->scene_scale = (float)value;
		else if (n == "nerf.training.dataset.offset.x") t->scene_offset[0] = (float)value;
		else if (n == "nerf.training.dataset.offset.y") t->scene_offset[1] = (float)value;
		else if (n == "nerf.training.dataset.offset.z") t->scene_offset[2] = (float)value;
		else NGPB_CHECK(false, "unknown option '" + n + "'");
	});
}
double ngp_testbed_get_option(ngp_testbed* t, const char* name_c) {
	const std::string n(name_c);
	const ngp_nerf_train_cfg& c = t->cfg;
	if (n == "nerf.training.n_images_for_training") return t->n_images_for_training;
	if (n == "nerf.training.random_bg_color") return c.random_bg_color;
	if (n == "nerf.training.linear_colors") return c.linear_colors;
	if (n == "nerf.training.snap_to_pixel_centers") return c.snap_to_pixel_centers;
	if (n == "nerf.training.near_distance") return c.near_distance;
	if (n == "nerf.training.loss_type") return c.loss_type;
	if (n == "nerf.training.train_mode") return c.train_mode;
	if (n == "nerf.training.math_mode") return c.math_mode;
	if (n == "nerf.training.gen_lanes_per_ray") return c.gen_lanes_per_ray;
	if (n == "nerf.training.gen_walk_empty") return c.gen_walk_empty;
	if (n == "nerf.training.gen_speculation") return c.gen_speculation;
	if (n == "nerf.training.optimize_exposure") return t->optimize_exposure ? 1.0 : 0.0;
	if (n == "nerf.training.exposure_l2_reg") return t->exposure_l2_reg;
	if (n == "nerf.training.n_steps_between_cam_updates") return t->n_steps_between_cam_updates;
	if (n == "nerf.training.compaction_order") return c.compaction_order;
	if (n == "nerf.training.drop_overflowing_rays") return t->drop_overflowing_rays;
	if (n == "render_math") return t->render_math;
	if (n == "render_mode") return t->render_mode;
	if (n == "nerf.training.density_grid_decay") return t->density_grid_decay;
	if (n == "nerf.rgb_activation") return c.rgb_activation;
	if (n == "nerf.density_activation") return c.density_activation;
	if (n == "nerf.cone_angle_constant") return c.march.cone_angle;
	if (n == "nerf.render_min_transmittance") return t->render_min_transmittance;
	if (n == "nerf.max_cascade") return c.max_cascade;
	if (n == "snap_to_pixel_centers") return t->render_snap_to_pixel_centers;
	if (n == "render_with_lens_distortion") return t->render_with_lens_distortion;
	if (n == "render_lens.mode") return t->render_lens_mode;
	if (n.rfind("render_lens.params.", 0) == 0 && n.size() == 20 && n[19] >= '0' && n[19] <= '3') return t->render_lens_params[n[19] - '0'];
	if (n == "color_space") return c.color_space;
	if (n == "shall_train") return t->shall_train;
	if (n == "aabb_scale") return t->aabb_scale;
	if (n == "learning_rate") return t->opt.learning_rate * t->lr_factor;
	if (n == "exposure") return t->exposure;
	if (n == "background_color.r") return c.background_color[0];
	if (n == "background_color.g") return c.background_color[1];
	if (n == "background_color.b") return c.background_color[2];
	if (n == "background_color.a") return t->background_alpha;
	if (n == "nerf.training.dataset.scale") return t->scene_scale;
	if (n == "nerf.training.dataset.offset.x") return t->scene_offset[0];
	if (n == "nerf.training.dataset.offset.y") return t->scene_offset[1];
	if (n == "nerf.training.dataset.offset.z") return t->scene_offset[2];
	if (n == "nerf.training.dataset.n_images") return t->n_images;
	set_last_error("unknown option '" + n + "'");
	return NAN;
}
int ngp_testbed_set_dp(ngp_testbed* t, uint32_t rank, uint32_t world) {
	NGPB_TRY({
		tb_invalidate_prefetch(t);
		NGPB_CHECK(world >= 1 && rank < world, "set_dp: rank must be < world");
		t->dp_rank = rank;
		t->dp_world = world;
	});
}
int ngp_testbed_train_compute_grads(ngp_testbed* t, uint32_t batch) { NGPB_TRY(tb_compute_grads(t, batch)); }
int ngp_testbed_train_front(ngp_testbed* t, uint32_t batch) { NGPB_TRY(tb_front(t, batch)); }
int ngp_testbed_train_back(ngp_testbed* t) { NGPB_TRY(tb_back(t, true)); }
int ngp_testbed_train_apply_grads(ngp_testbed* t) { NGPB_TRY(tb_apply_grads(t)); }
// One NCCL all-reduce (sum) of the flat fp16 gradient buffer — hash grid + MLPs, the buffer the optimizer reads (trainer.h:492-494
// is the reference's layout of it) — on the training stream, behind the forward/backward kernel.  The next step's sample generator
// already runs beside it on the side stream (tb_prefetch).
static void tb_allreduce_grads(ngp_testbed* t) {
	const ngpb::NcclApi& nccl = ngpb::NcclApi::get();
	PhaseTimer pt(t, 6);
	nccl.check(nccl.AllReduce(t->grads.p, t->grads.p, t->desc.n_params, ngpb::NcclApi::ncclFloat16, ngpb::NcclApi::ncclSum, t->comm_grads, t->stream), "ncclAllReduce(gradients)");
}

int ngp_testbed_train(ngp_testbed* t, uint32_t batch) {
	NGPB_TRY({
		if (!t->shall_train) return 0;
		if (t->dp_world > 1 && t->comm_grads) {
			tb_front(t, batch);
			tb_back(t, true);
			tb_allreduce_grads(t);
		} else {
			tb_compute_grads(t, batch);
		}
		tb_apply_grads(t);
	});
}

// ---- data parallel set-up: one process per GPU -------------------------------------------------------------------------------
size_t ngp_dp_unique_id_bytes(void) { return 2 * sizeof(ngpb::NcclApi::unique_id); }
int ngp_dp_unique_id(uint8_t* out, size_t capacity) {
	NGPB_TRY({
		NGPB_CHECK(out && capacity >= ngp_dp_unique_id_bytes(), "ngp_dp_unique_id: buffer too small (ngp_dp_unique_id_bytes())");
		const ngpb::NcclApi& nccl = ngpb::NcclApi::get();
		ngpb::NcclApi::unique_id ids[2];
		nccl.check(nccl.GetUniqueId(&ids[0]), "ncclGetUniqueId");
		nccl.check(nccl.GetUniqueId(&ids[1]), "ncclGetUniqueId");
		memcpy(out, ids, sizeof(ids));
	});
}
int ngp_testbed_init_dp(ngp_testbed* t, uint32_t rank, uint32_t world, const uint8_t* unique_id, size_t n_bytes) {
	NGPB_TRY({
		require_device();
		NGPB_CHECK(world >= 1 && rank < world, "init_dp: rank must be < world");
		NGPB_CHECK(unique_id && n_bytes >= ngp_dp_unique_id_bytes(), "init_dp: the unique id of ngp_dp_unique_id (rank 0) is required on every rank");
		NGPB_CHECK(!t->comm_grads, "init_dp: already initialised");
		tb_invalidate_prefetch(t);
		NGPB_CUDA_CHECK(cudaSetDevice(t->device));
		const ngpb::NcclApi& nccl = ngpb::NcclApi::get();
		ngpb::NcclApi::unique_id ids[2];
		memcpy(ids, unique_id, sizeof(ids));
		if (world > 1) {
			nccl.check(nccl.CommInitRank(&t->comm_grads, (int)world, ids[0], (int)rank), "ncclCommInitRank");
			nccl.check(nccl.CommInitRank(&t->comm_small, (int)world, ids[1], (int)rank), "ncclCommInitRank");
			NGPB_CUDA_CHECK(cudaStreamCreateWithFlags(&t->comm_stream, cudaStreamNonBlocking));
			NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&t->ev_counters_ready, cudaEventDisableTiming));
		}
		t->dp_rank = rank;
		t->dp_world = world;
	});
}
// Row-tile render sharding: after every rank has rendered its rows [y0, y1) into its own full-size frame buffers
// (ngp_testbed_render_device), rank `root` receives the tiles of the others: one ncclBroadcast per rank of its row block, in place.
// rgba / depth: device pointers to [height x width x 4] / [height x width] floats on every rank; rows are split evenly, ngp_dp_rows.
void ngp_dp_rows(uint32_t rank, uint32_t world, int32_t height, int32_t* y0, int32_t* y1) {
	const int32_t per = (height + (int32_t)world - 1) / (int32_t)world;
	*y0 = std::min(height, (int32_t)rank * per);
	*y1 = std::min(height, *y0 + per);
}
int ngp_testbed_gather_rows(ngp_testbed* t, int32_t width, int32_t height, float* rgba, float* depth) {
	NGPB_TRY({
		NGPB_CHECK(t->dp_world > 1 && t->comm_grads, "gather_rows: data parallel mode is not initialised (ngp_testbed_init_dp)");
		const ngpb::NcclApi& nccl = ngpb::NcclApi::get();
		for (uint32_t r = 0; r < t->dp_world; ++r) {
			int32_t y0, y1;
			ngp_dp_rows(r, t->dp_world, height, &y0, &y1);
			if (y1 <= y0) continue;
			float* c = rgba + (size_t)y0 * width * 4;
			float* d = depth + (size_t)y0 * width;
			nccl.check(nccl.Broadcast(c, c, (size_t)(y1 - y0) * width * 4, ngpb::NcclApi::ncclFloat32, (int)r, t->comm_grads, t->stream), "ncclBroadcast(rgba rows)");
			if (depth) nccl.check(nccl.Broadcast(d, d, (size_t)(y1 - y0) * width, ngpb::NcclApi::ncclFloat32, (int)r, t->comm_grads, t->stream), "ncclBroadcast(depth rows)");
		}
	});
}
void* ngp_testbed_grads(ngp_testbed* t) { return t->grads.p; }
void* ngp_testbed_params(ngp_testbed* t) { return t->params.p; }
void* ngp_testbed_params_inference(ngp_testbed* t) { return t->params_ema.p; }
float* ngp_testbed_params_fp32(ngp_testbed* t) { return t->params_fp32.p; }
uint32_t* ngp_testbed_dp_counters(ngp_testbed* t) {
	t->dp_counters.ensure(1);
	return reinterpret_cast<uint32_t*>(t->dp_counters.p);
}
uint32_t ngp_testbed_n_params(ngp_testbed* t) { return t->has_network ? t->desc.n_params : 0; }
uint32_t ngp_testbed_training_step(ngp_testbed* t) { return t->training_step; }
float ngp_testbed_loss(ngp_testbed* t) { return t->loss_scalar; }
int ngp_testbed_get_counters(ngp_testbed* t, uint32_t* rpb, uint32_t* mbs, uint32_t* mbsb) {
	if (rpb) *rpb = t->rays_per_batch;
	if (mbs) *mbs = t->measured_batch_size;
	if (mbsb) *mbsb = t->measured_batch_size_before_compaction;
	return 0;
}
int ngp_testbed_get_desc(ngp_testbed* t, ngp_nerf_desc* out) {
	NGPB_TRY({
		NGPB_CHECK(t->has_network, "no network configured");
		*out = t->desc;
	});
}
int ngp_testbed_set_params_fp32(ngp_testbed* t, const float* host, uint32_t n) { NGPB_TRY(tb_set_params_fp32(t, host, n)); }
int ngp_testbed_get_params_fp16(ngp_testbed* t, void* host, uint32_t n, int inference) {
	NGPB_TRY({
		NGPB_CHECK(t->has_network && n == t->desc.n_params, "get_params: wrong parameter count");
		NGPB_CUDA_CHECK(cudaMemcpyAsync(host, inference ? t->params_ema.p : t->params.p, (size_t)n * 2, cudaMemcpyDeviceToHost, t->stream));
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	});
}
int ngp_testbed_get_density_grid(ngp_testbed* t, float* grid_host, uint32_t n, uint8_t* bitfield_host, uint32_t n_bytes) {
	NGPB_TRY({
		NGPB_CHECK(t->density_grid.p, "no density grid yet");
		if (grid_host) {
			NGPB_CHECK(n <= GRID_N_CELLS * NGP_NERF_CASCADES, "density grid read too large");
			NGPB_CUDA_CHECK(cudaMemcpyAsync(grid_host, t->density_grid.p, (size_t)n * 4, cudaMemcpyDeviceToHost, t->stream));
		}
		if (bitfield_host) {
			NGPB_CHECK(n_bytes <= GRID_N_CELLS / 8 * NGP_NERF_CASCADES, "bitfield read too large");
			NGPB_CUDA_CHECK(cudaMemcpyAsync(bitfield_host, t->bitfield.p, n_bytes, cudaMemcpyDeviceToHost, t->stream));
		}
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	});
}
int ngp_testbed_set_density_grid(ngp_testbed* t, const float* grid_host, uint32_t n) {
	NGPB_TRY({
		tb_invalidate_prefetch(t);
		NGPB_CHECK(t->density_grid.p, "configure a network first");
		NGPB_CHECK(n <= GRID_N_CELLS * NGP_NERF_CASCADES, "density grid too large");
		NGPB_CUDA_CHECK(cudaMemcpyAsync(t->density_grid.p, grid_host, (size_t)n * 4, cudaMemcpyHostToDevice, t->stream));
		t->reduce_scratch.ensure(1024);
		update_bitfield(t->stream, t->cfg.max_cascade, t->density_grid.p, t->bitfield.p, t->mean_density.p, t->reduce_scratch.p);
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	});
}

static void tb_fill_render_cfg(ngp_testbed* t, ngp_render_cfg& rc, int32_t width, int32_t height, const float* cam, float fx, float fy, float cx, float cy) {
	memset(&rc, 0, sizeof(rc));
	rc.width = width;
	rc.height = height;
	rc.focal_x = fx;
	rc.focal_y = fy;
	rc.screen_x = cx;
	rc.screen_y = cy;
	for (int c = 0; c < 4; ++c)
		for (int r = 0; r < 3; ++r) rc.camera[c * 3 + r] = cam[r * 4 + c];
	for (int k = 0; k < 3; ++k) {
		rc.aabb_min[k] = rc.render_aabb_min[k] = t->cfg.aabb_min[k];
		rc.aabb_max[k] = rc.render_aabb_max[k] = t->cfg.aabb_max[k];
	}
	rc.max_cascade = t->cfg.max_cascade;
	rc.march = t->cfg.march;
	rc.rgb_activation = t->cfg.rgb_activation;
	rc.density_activation = t->cfg.density_activation;
	rc.min_transmittance = t->render_min_transmittance;
	rc.spp_index = t->render_spp_index;
	rc.near_distance = 0.0f;
	render_pixel_offset(t->render_snap_to_pixel_centers ? 0u : t->render_spp_index, rc.pixel_offset);
	rc.lens_mode = t->render_with_lens_distortion ? t->render_lens_mode : (uint32_t)NGP_LENS_PERSPECTIVE;
	for (int k = 0; k < 4; ++k) rc.lens_params[k] = t->render_with_lens_distortion ? t->render_lens_params[k] : 0.0f;
	rc.math_mode = t->render_math;
	rc.render_mode = t->render_mode;
	rc.skips_per_tile = t->render_skips_per_tile;
	rc.depth_scale = 1.0f / t->scene_scale;   // testbed_nerf.cu:2037
}

int ngp_testbed_render_device(ngp_testbed* t, int32_t width, int32_t height, const float* cam, float fx, float fy, float cx, float cy, int32_t y0,
	int32_t y1, float* rgba_dev, float* depth_dev) {
	NGPB_TRY({
		NGPB_CHECK(t->has_network, "render: no network");
		NGPB_CHECK(width > 0 && height > 0 && y0 >= 0 && y1 <= height && y0 < y1, "render: bad frame / tile bounds");
		ngp_render_cfg rc;
		tb_fill_render_cfg(t, rc, width, height, cam, fx, fy, cx, cy);
		t->render_scratch.ensure(render_scratch_bytes(width, y1 - y0));
		t->render_counter.ensure(4);
		render_nerf(t->desc, t->stream, rc, y0, y1, t->params_ema.p, t->bitfield.p, rgba_dev, depth_dev, t->render_scratch.p, t->render_counter.p);
	});
}
int ngp_testbed_render(ngp_testbed* t, int32_t width, int32_t height, const float* cam, float fx, float fy, float cx, float cy, int32_t y0, int32_t y1,
	float* rgba_host, float* depth_host, uint32_t* n_steps_total) {
	NGPB_TRY({
		NGPB_CHECK(width > 0 && height > 0, "render: bad frame size");
		const size_t n_px = (size_t)width * height;
		t->render_rgba.ensure(n_px * 4);
		t->render_depth.ensure(n_px);
		if (ngp_testbed_render_device(t, width, height, cam, fx, fy, cx, cy, y0, y1, t->render_rgba.p, t->render_depth.p)) throw std::runtime_error(g_last_error);
		const size_t off = (size_t)y0 * width, cnt = (size_t)(y1 - y0) * width;
		NGPB_CUDA_CHECK(cudaMemcpyAsync(rgba_host + off * 4, t->render_rgba.p + off * 4, cnt * 16, cudaMemcpyDeviceToHost, t->stream));
		if (depth_host) NGPB_CUDA_CHECK(cudaMemcpyAsync(depth_host + off, t->render_depth.p + off, cnt * 4, cudaMemcpyDeviceToHost, t->stream));
		uint32_t steps = 0;
		NGPB_CUDA_CHECK(cudaMemcpyAsync(&steps, t->render_counter.p, 4, cudaMemcpyDeviceToHost, t->stream));
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
		if (n_steps_total) *n_steps_total = steps;
	});
}

int ngp_render_accumulate(void* stream, int32_t w, int32_t h, const float* frame, float* acc, float sample_count, uint32_t color_space) {
	NGPB_TRY(require_device(); render_accumulate((cudaStream_t)stream, w, h, frame, acc, sample_count, color_space));
}
int ngp_render_tonemap(void* stream, int32_t w, int32_t h, const ngp_tonemap_cfg* cfg, const float* acc, float* out) {
	NGPB_TRY(require_device(); render_tonemap((cudaStream_t)stream, w, h, *cfg, acc, out));
}
int ngp_testbed_render_ex(ngp_testbed* t, int32_t width, int32_t height, const float* cam, float fx, float fy, float cx, float cy, uint32_t spp, int linear,
	float* rgba_host, float* depth_host) {
	NGPB_TRY({
		NGPB_CHECK(width > 0 && height > 0 && spp >= 1, "render: bad frame size / spp");
		const size_t n_px = (size_t)width * height;
		t->render_rgba.ensure(n_px * 4);
		t->render_depth.ensure(n_px);
		DevBuf<float> acc, out;
		acc.ensure(n_px * 4);
		out.ensure(n_px * 4);
		NGPB_CUDA_CHECK(cudaMemsetAsync(acc.p, 0, n_px * 16, t->stream));
		const uint32_t saved = t->render_spp_index;
		for (uint32_t sidx = 0; sidx < spp; ++sidx) {
			t->render_spp_index = sidx;
			if (ngp_testbed_render_device(t, width, height, cam, fx, fy, cx, cy, 0, height, t->render_rgba.p, t->render_depth.p)) {
				t->render_spp_index = saved;
				throw std::runtime_error(g_last_error);
			}
			render_accumulate(t->stream, width, height, t->render_rgba.p, acc.p, (float)sidx, NGP_COLOR_LINEAR);
		}
		t->render_spp_index = saved;
		ngp_tonemap_cfg tm{};
		tm.exposure = t->exposure;
		for (int k = 0; k < 3; ++k) tm.background_color[k] = t->cfg.background_color[k];
		tm.background_color[3] = t->background_alpha;
		tm.color_space = NGP_COLOR_LINEAR;
		tm.output_color_space = linear ? NGP_COLOR_LINEAR : NGP_COLOR_SRGB;
		tm.tonemap_curve = NGP_TONEMAP_IDENTITY;
		tm.clamp_output_color = 0;
		tm.unmultiply_alpha = 0;
		render_tonemap(t->stream, width, height, tm, acc.p, out.p);
		NGPB_CUDA_CHECK(cudaMemcpyAsync(rgba_host, out.p, n_px * 16, cudaMemcpyDeviceToHost, t->stream));
		if (depth_host) NGPB_CUDA_CHECK(cudaMemcpyAsync(depth_host, t->render_depth.p, n_px * 4, cudaMemcpyDeviceToHost, t->stream));
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	});
}

// ".ngpb": a flat little-endian dump of the full training state (fp32 masters, EMA params, optimizer moments, RNG streams, density
// grid) for exact resume.  The reference's own container (.ingp / .msgpack) follows below.
struct SnapshotHeader {
	char magic[8];
	uint32_t version, n_params, n_grid, training_step, optimizer_step, rays_per_batch, measured_batch_size, measured_before, ema_step, aabb_scale;
	float lr_factor;
	uint64_t rng_state, rng_inc, grng_state, grng_inc;
	ngp_nerf_desc desc;
};
}  // extern "C"

// ---- the reference's container: msgpack of {network config..., "snapshot": {...}}, gzip-wrapped for ".ingp"
// (Testbed::save_snapshot / load_snapshot, src/testbed.cu:5288-5485; Trainer::serialize trainer.h:442-482; Adam / Ema /
// ExponentialDecay serialize adam.h:304-326, ema.h:190-206, exponential_decay.h:136-148; NerfDataset json_binding.h:112-190).
static bool has_ext(const std::string& path, const char* ext) {
	const size_t n = strlen(ext);
	if (path.size() < n) return false;
	return to_lower(path.substr(path.size() - n)) == ext;
}
template <typename T>
static Json dev_to_bin(const T* dev, size_t count) {
	std::vector<T> h(count);
	NGPB_CUDA_CHECK(cudaMemcpy(h.data(), dev, count * sizeof(T), cudaMemcpyDeviceToHost));
	return jbin(h.data(), count * sizeof(T));
}
static Json mat4x3_to_json(const float* m) {  // tmat<float,4,3>: 3 rows of 4 (vec_json.h:38-46); m is column-major [c*3 + r]
	Json rows = jarr();
	for (int r = 0; r < 3; ++r) {
		Json row = jarr();
		for (int c = 0; c < 4; ++c) row.arr.push_back(jnum(m[c * 3 + r]));
		rows.arr.push_back(row);
	}
	return rows;
}
static void mat4x3_from_json(const Json& j, float* m) {
	NGPB_CHECK(j.type == Json::Array && j.arr.size() == 3, "snapshot: bad mat4x3");
	for (int r = 0; r < 3; ++r) {
		NGPB_CHECK(j.arr[r].type == Json::Array && j.arr[r].arr.size() == 4, "snapshot: bad mat4x3 row");
		for (int c = 0; c < 4; ++c) m[c * 3 + r] = (float)j.arr[r].arr[c].num;
	}
}
static Json aabb_to_json(const float* mn, const float* mx) {
	Json b = jobj();
	b.obj["min"] = jvec(mn, 3);
	b.obj["max"] = jvec(mx, 3);
	return b;
}

static Json tb_snapshot_json(ngp_testbed* t, bool include_optimizer_state) {
	NGPB_CHECK(t->has_network, "save_snapshot: no network");
	NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	if (t->side_stream) NGPB_CUDA_CHECK(cudaStreamSynchronize(t->side_stream));
	const size_t n = t->desc.n_params;
	Json cfg = t->network_config;
	Json snap = jobj();
	// Trainer::serialize: the INFERENCE (EMA) weights
	snap.obj["n_params"] = jint((int64_t)n);
	snap.obj["params_type"] = jstr("__half");
	snap.obj["params_binary"] = dev_to_bin(t->params_ema.p, n);
	if (include_optimizer_state) {
		Json adam = jobj();
		adam.obj["current_step"] = jint(t->optimizer_step);
		adam.obj["base_learning_rate"] = jnum(t->opt.learning_rate);
		adam.obj["first_moments_binary"] = dev_to_bin(t->m1.p, n);
		adam.obj["second_moments_binary"] = dev_to_bin(t->m2.p, n);
		adam.obj["param_steps_binary"] = dev_to_bin(t->param_steps.p, n);
		Json o = adam;
		if (t->opt.has_decay) {
			Json d = jobj();
			d.obj["nested"] = o;
			d.obj["learning_rate"] = jnum(t->opt.learning_rate);
			d.obj["learning_rate_factor"] = jnum(t->lr_factor);
			o = d;
		}
		if (t->opt.has_ema) {
			Json e = jobj();
			e.obj["nested"] = o;
			e.obj["weights_ema_binary"] = dev_to_bin(t->params_ema.p, n);
			o = e;
		}
		snap.obj["optimizer"] = o;
	}
	snap.obj["version"] = jint(1);
	snap.obj["mode"] = jstr("nerf");
	snap.obj["density_grid_size"] = jint(128);
	{
		const size_t n_grid = (size_t)GRID_N_CELLS * (t->cfg.max_cascade + 1);
		std::vector<float> g(n_grid);
		NGPB_CUDA_CHECK(cudaMemcpy(g.data(), t->density_grid.p, n_grid * 4, cudaMemcpyDeviceToHost));
		std::vector<__half> gh(n_grid);
		for (size_t i = 0; i < n_grid; ++i) gh[i] = __float2half_rn(g[i]);
		snap.obj["density_grid_binary"] = jbin(gh.data(), n_grid * 2);
	}
	Json nerf = jobj();
	nerf.obj["aabb_scale"] = jint(t->aabb_scale);
	Json rgb = jobj();
	rgb.obj["rays_per_batch"] = jint(t->rays_per_batch);
	rgb.obj["measured_batch_size"] = jint(t->measured_batch_size);
	rgb.obj["measured_batch_size_before_compaction"] = jint(t->measured_batch_size_before_compaction);
	nerf.obj["rgb"] = rgb;
	{
		Json ds = jobj();
		ds.obj["n_images"] = jint(t->n_images);
		Json paths = jarr(), meta = jarr(), xf = jarr();
		for (uint32_t i = 0; i < t->n_images; ++i) {
			const ngp_train_view& v = t->views[i];
			paths.arr.push_back(jstr(""));
			Json m = jobj();
			const float fl[2] = {v.focal_x, v.focal_y}, pp[2] = {v.principal_x, v.principal_y}, rs[4] = {0, 0, 0, 0};
			m.obj["focal_length"] = jvec(fl, 2);
			Json lens = jobj();
			if (v.lens_mode == NGP_LENS_OPENCV) {
				lens.obj["is_fisheye"] = jbool(false);
				lens.obj["k1"] = jnum(v.lens_params[0]);
				lens.obj["k2"] = jnum(v.lens_params[1]);
				lens.obj["p1"] = jnum(v.lens_params[2]);
				lens.obj["p2"] = jnum(v.lens_params[3]);
			}
			m.obj["lens"] = lens;
			m.obj["principal_point"] = jvec(pp, 2);
			m.obj["rolling_shutter"] = jvec(rs, 4);
			Json res = jarr();
			res.arr.push_back(jint(v.width));
			res.arr.push_back(jint(v.height));
			m.obj["resolution"] = res;
			meta.arr.push_back(m);
			Json x = jobj();
			x.obj["start"] = mat4x3_to_json(v.xform);
			x.obj["end"] = mat4x3_to_json(v.xform);
			xf.arr.push_back(x);
		}
		ds.obj["paths"] = paths;
		ds.obj["metadata"] = meta;
		ds.obj["xforms"] = xf;
		ds.obj["render_aabb"] = aabb_to_json(t->cfg.aabb_min, t->cfg.aabb_max);
		const float idm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
		Json rl = jarr();
		for (int r = 0; r < 3; ++r) rl.arr.push_back(jvec(idm + 3 * r, 3));
		ds.obj["render_aabb_to_local"] = rl;
		const float up[3] = {0, 1, 0};
		ds.obj["up"] = jvec(up, 3);
		ds.obj["offset"] = jvec(t->scene_offset, 3);
		Json er = jarr();
		er.arr.push_back(jint(0));
		er.arr.push_back(jint(0));
		ds.obj["envmap_resolution"] = er;
		ds.obj["scale"] = jnum(t->scene_scale);
		ds.obj["aabb_scale"] = jint(t->aabb_scale);
		ds.obj["from_mitsuba"] = jbool(false);
		ds.obj["is_hdr"] = jbool(false);
		ds.obj["wants_importance_sampling"] = jbool(true);
		ds.obj["n_extra_learnable_dims"] = jint(0);
		nerf.obj["dataset"] = ds;
	}
	snap.obj["nerf"] = nerf;
	snap.obj["training_step"] = jint(t->training_step);
	snap.obj["loss"] = jnum(t->loss_scalar);
	snap.obj["aabb"] = aabb_to_json(t->cfg.aabb_min, t->cfg.aabb_max);
	snap.obj["render_aabb"] = aabb_to_json(t->cfg.aabb_min, t->cfg.aabb_max);
	snap.obj["bounding_radius"] = jnum(1.0);
	{
		const float bg[4] = {t->cfg.background_color[0], t->cfg.background_color[1], t->cfg.background_color[2], t->background_alpha};
		snap.obj["background_color"] = jvec(bg, 4);
	}
	snap.obj["exposure"] = jnum(t->exposure);
	cfg.obj["snapshot"] = snap;
	return cfg;
}

static void bin_to_dev(const Json& j, void* dev, size_t bytes, const char* what) {
	NGPB_CHECK(j.type == Json::Binary && j.bin.size() == bytes, std::string("snapshot: '") + what + "' has the wrong size");
	NGPB_CUDA_CHECK(cudaMemcpy(dev, j.bin.data(), bytes, cudaMemcpyHostToDevice));
}
__global__ void k_half_to_float(const uint32_t n, const __half* __restrict__ in, float* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = __half2float(in[i]);
}
static void tb_load_snapshot_json(ngp_testbed* t, const Json& config) {
	NGPB_CHECK(config.contains("snapshot"), "file does not contain a snapshot");
	const Json& snap = config.at("snapshot");
	NGPB_CHECK((uint32_t)snap.value("version", 0.0) >= 1, "Snapshot uses an old format and can not be loaded.");
	NGPB_CHECK(to_lower(snap.value("mode", std::string("nerf"))) == "nerf", "snapshot mode is not nerf");
	NGPB_CHECK((uint32_t)snap.value("density_grid_size", 0.0) == 128, "Incompatible grid size.");
	tb_invalidate_prefetch(t);
	const Json& nerf = snap.sub("nerf");
	// dataset: keep a loaded one, else take the metadata from the snapshot (render-only use; src/testbed.cu:5387-5392)
	if (t->n_images == 0 && nerf.contains("dataset")) {
		const Json& ds = nerf.at("dataset");
		const uint32_t n_img = (uint32_t)ds.value("n_images", 0.0);
		t->n_images = n_img;
		t->aabb_scale = (uint32_t)ds.value("aabb_scale", 1.0);
		// NerfDataset::scale / offset (json_binding.h:160-166): what set_nerf_camera_matrix converts poses with
		t->scene_scale = (float)ds.value("scale", (double)t->scene_scale);
		if (ds.contains("offset") && ds.at("offset").type == Json::Array && ds.at("offset").arr.size() == 3)
			for (int k = 0; k < 3; ++k) t->scene_offset[k] = (float)ds.at("offset").arr[k].num;
		t->views.assign(n_img, ngp_train_view{});
		t->pixel_bufs.assign(n_img, nullptr);
		for (uint32_t i = 0; i < n_img; ++i) {
			ngp_train_view& v = t->views[i];
			const Json& m = ds.at("metadata").arr.at(i);
			v.width = (int32_t)m.at("resolution").arr.at(0).num;
			v.height = (int32_t)m.at("resolution").arr.at(1).num;
			v.focal_x = (float)m.at("focal_length").arr.at(0).num;
			v.focal_y = (float)m.at("focal_length").arr.at(1).num;
			v.principal_x = (float)m.at("principal_point").arr.at(0).num;
			v.principal_y = (float)m.at("principal_point").arr.at(1).num;
			const Json& lens = m.sub("lens");
			// from_json(Lens) (json_binding.h:67-99) tells the mode from the keys present; only perspective and OpenCV are built here
			NGPB_CHECK(!(lens.contains("k1") && lens.contains("is_fisheye") && lens.at("is_fisheye").b) && !lens.contains("ftheta_p0") && !lens.contains("latlong") &&
				!lens.contains("equirectangular") && !lens.contains("orthographic"),
				"snapshot: view " + std::to_string(i) + " uses a lens this build does not implement (perspective and OpenCV only)");
			if (lens.contains("k1")) {
				v.lens_mode = NGP_LENS_OPENCV;
				v.lens_params[0] = (float)lens.value("k1", 0.0);
				v.lens_params[1] = (float)lens.value("k2", 0.0);
				v.lens_params[2] = (float)lens.value("p1", 0.0);
				v.lens_params[3] = (float)lens.value("p2", 0.0);
			}
			mat4x3_from_json(ds.at("xforms").arr.at(i).at("start"), v.xform);
		}
		t->n_images_for_training = 0;  // no pixels: not trainable until images are set
		t->views_dirty = true;
	} else if (nerf.contains("aabb_scale")) {
		t->aabb_scale = (uint32_t)nerf.value("aabb_scale", (double)t->aabb_scale);
	}
	tb_update_scene(t);
	tb_reset_network(t, config);   // reset_network(false) from the snapshot's own config
	const size_t n = t->desc.n_params;
	NGPB_CHECK((size_t)snap.value("n_params", 0.0) == n, "snapshot: n_params does not match the network config");

	// Trainer::deserialize: all three parameter buffers from the stored (inference) weights
	const std::string ptype = snap.value("params_type", std::string("__half"));
	const Json& pb = snap.at("params_binary");
	if (ptype == "float") {
		bin_to_dev(pb, t->params_fp32.p, n * 4, "params_binary");
	} else {
		NGPB_CHECK(ptype == "__half", "Trainer: snapshot parameters must be of type float of __half");
		bin_to_dev(pb, t->params_ema.p, n * 2, "params_binary");
		k_half_to_float<<<div_round_up((uint32_t)n, 256), 256, 0, t->stream>>>((uint32_t)n, t->params_ema.p, t->params_fp32.p);
		NGPB_LAUNCHED();
	}
	k_cast_params<<<div_round_up((uint32_t)n, 256), 256, 0, t->stream>>>((uint32_t)n, t->params_fp32.p, t->params.p, t->params_ema.p);
	NGPB_LAUNCHED();

	if (snap.contains("optimizer")) {
		const Json* o = &snap.at("optimizer");
		if (t->opt.has_ema) {
			bin_to_dev(o->at("weights_ema_binary"), t->params_ema.p, n * 2, "weights_ema_binary");
			o = &o->at("nested");
		}
		if (t->opt.has_decay) {
			t->lr_factor = (float)o->value("learning_rate_factor", 1.0);
			o = &o->at("nested");
		}
		bin_to_dev(o->at("first_moments_binary"), t->m1.p, n * 4, "first_moments_binary");
		bin_to_dev(o->at("second_moments_binary"), t->m2.p, n * 4, "second_moments_binary");
		if (o->contains("param_steps_binary")) bin_to_dev(o->at("param_steps_binary"), t->param_steps.p, n * 4, "param_steps_binary");
		t->optimizer_step = (uint32_t)o->value("current_step", 0.0);
	}

	// density grid: fp16 -> float, then mean + bitfield (update_density_grid_mean_and_bitfield)
	const Json& dg = snap.at("density_grid_binary");
	NGPB_CHECK(dg.type == Json::Binary, "snapshot: density_grid_binary missing");
	const size_t n_grid = dg.bin.size() / 2;
	NGPB_CHECK(n_grid == 0 || n_grid == (size_t)GRID_N_CELLS * (t->cfg.max_cascade + 1), "Incompatible number of grid cascades.");
	if (n_grid) {
		DevBuf<__half> tmp;
		tmp.ensure(n_grid);
		NGPB_CUDA_CHECK(cudaMemcpy(tmp.p, dg.bin.data(), n_grid * 2, cudaMemcpyHostToDevice));
		k_half_to_float<<<div_round_up((uint32_t)n_grid, 256), 256, 0, t->stream>>>((uint32_t)n_grid, tmp.p, t->density_grid.p);
		NGPB_LAUNCHED();
		t->reduce_scratch.ensure(1024);
		update_bitfield(t->stream, t->cfg.max_cascade, t->density_grid.p, t->bitfield.p, t->mean_density.p, t->reduce_scratch.p);
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	}
	const Json& rgb = nerf.sub("rgb");
	// per-step buffers are sized for at most 2^18 rays (tb_ensure_step_scratch); the controller keeps multiples of the batch granularity
	t->rays_per_batch = std::min(std::max(next_multiple((uint32_t)rgb.value("rays_per_batch", (double)t->rays_per_batch), NGP_BATCH_GRANULARITY), NGP_BATCH_GRANULARITY), 1u << 18);
	t->measured_batch_size = (uint32_t)rgb.value("measured_batch_size", 0.0);
	t->measured_batch_size_before_compaction = (uint32_t)rgb.value("measured_batch_size_before_compaction", 0.0);
	t->training_step = (uint32_t)snap.value("training_step", 0.0);
	if (!snap.contains("optimizer")) t->optimizer_step = 0;   // no optimizer state in the file: Adam restarts (moments zero, step 0), like Trainer::deserialize without "optimizer"
	t->loss_scalar = (float)snap.value("loss", 0.0);
	t->exposure = (float)snap.value("exposure", (double)t->exposure);
	if (snap.contains("background_color") && snap.at("background_color").type == Json::Array && snap.at("background_color").arr.size() == 4) {
		const Json& bg = snap.at("background_color");
		for (int k = 0; k < 3; ++k) t->cfg.background_color[k] = (float)bg.arr[k].num;
		t->background_alpha = (float)bg.arr[3].num;
	}
	// the occupancy grid of a trained model is past its warm-up phase
	t->density_grid_ema_step = t->training_step / 16 + (t->training_step < 256 ? t->training_step : 0);
	NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
}

extern "C" {

int ngp_testbed_save_snapshot_ex(ngp_testbed* t, const char* path, int include_optimizer_state, int compress) {
	NGPB_TRY({
		const std::string p(path);
		NGPB_CHECK(has_ext(p, ".ingp") || has_ext(p, ".msgpack"), "save_snapshot_ex: path must end in .ingp or .msgpack");
		MsgPackWriter w;
		w.write(tb_snapshot_json(t, include_optimizer_state != 0));
		std::vector<uint8_t> bytes = std::move(w.out);
		if (has_ext(p, ".ingp")) bytes = gzip_compress(bytes, compress ? Z_DEFAULT_COMPRESSION : Z_NO_COMPRESSION);
		std::ofstream f(path, std::ios::binary);
		NGPB_CHECK(f.good(), std::string("cannot open ") + path);
		f.write((const char*)bytes.data(), (std::streamsize)bytes.size());
		NGPB_CHECK(f.good(), "snapshot write failed");
	});
}

int ngp_testbed_save_snapshot(ngp_testbed* t, const char* path) {
	if (has_ext(path, ".ingp") || has_ext(path, ".msgpack")) return ngp_testbed_save_snapshot_ex(t, path, 0, 1);
	NGPB_TRY({
		NGPB_CHECK(t->has_network, "save_snapshot: no network");
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
		SnapshotHeader h{};
		memcpy(h.magic, "NGPB200\0", 8);
		h.version = 2;
		h.n_params = t->desc.n_params;
		h.n_grid = GRID_N_CELLS * (t->cfg.max_cascade + 1);
		h.training_step = t->training_step;
		h.optimizer_step = t->optimizer_step;
		h.rays_per_batch = t->rays_per_batch;
		h.measured_batch_size = t->measured_batch_size;
		h.measured_before = t->measured_batch_size_before_compaction;
		h.ema_step = t->density_grid_ema_step;
		h.aabb_scale = t->aabb_scale;
		h.lr_factor = t->lr_factor;
		h.rng_state = t->rng.state; h.rng_inc = t->rng.inc; h.grng_state = t->density_grid_rng.state; h.grng_inc = t->density_grid_rng.inc;
		h.desc = t->desc;
		std::ofstream f(path, std::ios::binary);
		NGPB_CHECK(f.good(), std::string("cannot open ") + path);
		f.write(reinterpret_cast<const char*>(&h), sizeof(h));
		{
			// version 2: the network config (encoding, MLPs, loss, optimizer chain) as MessagePack, so that a fresh Testbed resumes with the
			// optimizer and loss it was trained with instead of the defaults
			MsgPackWriter w;
			w.write(t->network_config);
			const uint64_t n_cfg = w.out.size();
			f.write(reinterpret_cast<const char*>(&n_cfg), sizeof(n_cfg));
			f.write(reinterpret_cast<const char*>(w.out.data()), (std::streamsize)n_cfg);
		}
		auto dump = [&](const void* dev, size_t bytes) {
			std::vector<char> buf(bytes);
			NGPB_CUDA_CHECK(cudaMemcpy(buf.data(), dev, bytes, cudaMemcpyDeviceToHost));
			f.write(buf.data(), (std::streamsize)bytes);
		};
		const size_t n = h.n_params;
		dump(t->params_fp32.p, n * 4);
		dump(t->params_ema.p, n * 2);
		dump(t->m1.p, n * 4);
		dump(t->m2.p, n * 4);
		dump(t->param_steps.p, n * 4);
		dump(t->density_grid.p, (size_t)h.n_grid * 4);
		NGPB_CHECK(f.good(), "snapshot write failed");
	});
}
int ngp_testbed_load_snapshot(ngp_testbed* t, const char* path) {
	if (has_ext(path, ".ingp") || has_ext(path, ".msgpack")) {
		NGPB_TRY({
			std::ifstream f(path, std::ios::binary);
			NGPB_CHECK(f.good(), std::string("Network snapshot '") + path + "' does not exist.");
			std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
			if (has_ext(path, ".ingp")) bytes = gzip_decompress(bytes);
			MsgPackReader r(bytes.data(), bytes.size());
			tb_load_snapshot_json(t, r.read());
		});
	}
	NGPB_TRY({
		tb_invalidate_prefetch(t);
		std::ifstream f(path, std::ios::binary);
		NGPB_CHECK(f.good(), std::string("snapshot not found: ") + path);
		SnapshotHeader h{};
		f.read(reinterpret_cast<char*>(&h), sizeof(h));
		NGPB_CHECK(f.good() && memcmp(h.magic, "NGPB200\0", 8) == 0 && (h.version == 1 || h.version == 2), "not an ngp_b200 snapshot");
		NGPB_CHECK(h.aabb_scale >= 1 && h.aabb_scale <= (1u << (NGP_NERF_CASCADES - 1)) && (h.aabb_scale & (h.aabb_scale - 1)) == 0, "snapshot: bad aabb_scale");
		if (h.version >= 2) {
			// rebuild the network from the stored config (tb_reset_network: descriptor, loss, optimizer chain), then check it against the header
			uint64_t n_cfg = 0;
			f.read(reinterpret_cast<char*>(&n_cfg), sizeof(n_cfg));
			NGPB_CHECK(f.good() && n_cfg > 0 && n_cfg < (64u << 20), "snapshot: bad network config block");
			std::vector<uint8_t> cfg_bytes(n_cfg);
			f.read(reinterpret_cast<char*>(cfg_bytes.data()), (std::streamsize)n_cfg);
			NGPB_CHECK(f.good(), "snapshot truncated");
			t->aabb_scale = h.aabb_scale;
			tb_update_scene(t);
			MsgPackReader r(cfg_bytes.data(), cfg_bytes.size());
			tb_reset_network(t, r.read());
		} else {
			// version 1 carries no config: only into a Testbed whose network is already configured the same way
			NGPB_CHECK(t->has_network && t->aabb_scale == h.aabb_scale, "this .ngpb (version 1) carries no network config: configure the same network and dataset first");
		}
		// the buffers below are sized from t->desc: the header must describe exactly that network
		NGPB_CHECK(memcmp(&h.desc, &t->desc, sizeof(ngp_nerf_desc)) == 0, "snapshot: the network descriptor does not match the network config");
		NGPB_CHECK(h.n_params == t->desc.n_params, "snapshot: parameter count does not match the network");
		NGPB_CHECK(h.n_grid == GRID_N_CELLS * (t->cfg.max_cascade + 1), "snapshot: density grid size does not match aabb_scale");
		NGPB_CHECK(h.rays_per_batch >= 1 && h.rays_per_batch <= (1u << 18), "snapshot: rays_per_batch out of range");
		t->density_grid.ensure(GRID_N_CELLS * NGP_NERF_CASCADES);
		t->bitfield.ensure(GRID_N_CELLS / 8 * NGP_NERF_CASCADES);
		t->mean_density.ensure(4);
		NGPB_CUDA_CHECK(cudaMemset(t->density_grid.p, 0, sizeof(float) * GRID_N_CELLS * NGP_NERF_CASCADES));
		auto load = [&](void* dev, size_t bytes) {
			std::vector<char> buf(bytes);
			f.read(buf.data(), (std::streamsize)bytes);
			NGPB_CHECK(f.good(), "snapshot truncated");
			NGPB_CUDA_CHECK(cudaMemcpy(dev, buf.data(), bytes, cudaMemcpyHostToDevice));
		};
		const size_t n = h.n_params;
		load(t->params_fp32.p, n * 4);
		load(t->params_ema.p, n * 2);
		load(t->m1.p, n * 4);
		load(t->m2.p, n * 4);
		load(t->param_steps.p, n * 4);
		load(t->density_grid.p, (size_t)h.n_grid * 4);
		// fp16 working copy = cast of the master weights (Trainer::deserialize rewrites all three buffers, trainer.h:457-482)
		std::vector<__half> tmp_unused;
		k_cast_params<<<div_round_up((uint32_t)n, 256), 256, 0, t->stream>>>((uint32_t)n, t->params_fp32.p, t->params.p, t->params.p);
		NGPB_LAUNCHED();
		t->training_step = h.training_step;
		t->optimizer_step = h.optimizer_step;
		t->rays_per_batch = h.rays_per_batch;
		t->measured_batch_size = h.measured_batch_size;
		t->measured_batch_size_before_compaction = h.measured_before;
		t->density_grid_ema_step = h.ema_step;
		t->lr_factor = h.lr_factor;
		t->rng = Pcg32(h.rng_state, h.rng_inc, true);
		t->density_grid_rng = Pcg32(h.grng_state, h.grng_inc, true);
		t->reduce_scratch.ensure(1024);
		update_bitfield(t->stream, t->cfg.max_cascade, t->density_grid.p, t->bitfield.p, t->mean_density.p, t->reduce_scratch.p);
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	});
}
// Testbed::load_network_config for a .json file, "parent" inheritance applied; JSON text out (host only, no device needed)
int ngp_load_network_config(const char* path, char* out, size_t capacity, size_t* n_out) {
	NGPB_TRY({
		std::string text;
		json_dump(load_network_config_file(path), text);
		*n_out = text.size();
		NGPB_CHECK(text.size() + 1 <= capacity, "ngp_load_network_config: output buffer too small");
		memcpy(out, text.c_str(), text.size() + 1);
	});
}
// codec hooks for the tests: JSON text -> msgpack (optionally gzip) and back (binary values print as {"bytes": n})
int ngp_json_to_msgpack(const char* json_text, int gzip, uint8_t* out, size_t capacity, size_t* n_out) {
	NGPB_TRY({
		const std::string text(json_text);
		MsgPackWriter w;
		w.write(JsonParser(text).parse());
		std::vector<uint8_t> bytes = std::move(w.out);
		if (gzip) bytes = gzip_compress(bytes);
		*n_out = bytes.size();
		NGPB_CHECK(bytes.size() <= capacity, "ngp_json_to_msgpack: output buffer too small");
		memcpy(out, bytes.data(), bytes.size());
	});
}
int ngp_msgpack_to_json(const uint8_t* data, size_t n, int gzip, char* out, size_t capacity, size_t* n_out) {
	NGPB_TRY({
		std::vector<uint8_t> bytes(data, data + n);
		if (gzip) bytes = gzip_decompress(bytes);
		MsgPackReader r(bytes.data(), bytes.size());
		const Json j = r.read();
		NGPB_CHECK(r.at_end(), "msgpack: trailing bytes");
		std::string text;
		json_dump(j, text);
		*n_out = text.size();
		NGPB_CHECK(text.size() + 1 <= capacity, "ngp_msgpack_to_json: output buffer too small");
		memcpy(out, text.c_str(), text.size() + 1);
	});
}
int ngp_testbed_set_profiling(ngp_testbed* t, int enable) {
	t->profiling = enable != 0;
	for (int p = 0; p < NGP_N_PHASES; ++p) {
		t->phase_ms[p] = 0.0f;
		t->ev_used[p] = false;
	}
	t->phase_steps = 0;
	return 0;
}
int ngp_testbed_get_phase_ms(ngp_testbed* t, float* ms_out, uint32_t* n_steps) {
	for (int p = 0; p < NGP_N_PHASES; ++p) {
		ms_out[p] = t->phase_ms[p];
		t->phase_ms[p] = 0.0f;
	}
	if (n_steps) *n_steps = t->phase_steps;
	t->phase_steps = 0;
	return 0;
}
int ngp_testbed_update_image_async(ngp_testbed* t, uint32_t idx, const void* rgba_host) {
	NGPB_TRY({
		NGPB_CHECK(idx < t->n_images && t->pixel_bufs[idx], "update_image_async: image slot was never set");
		const ngp_train_view& v = t->views[idx];
		const size_t px_bytes = v.image_type == NGP_IMAGE_BYTE ? 4 : (v.image_type == NGP_IMAGE_HALF ? 8 : 16);
		const size_t bytes = (size_t)v.width * v.height * px_bytes;
		if (!v.no_mask) {
			// The sample generator looks at pixels only to skip masked ones.  A view that may contain masked pixels is replaced in stream order,
			// ahead of the generator, and a generator launch already in flight for the next step is discarded.
			tb_invalidate_prefetch(t);
			tb_apply_pending_images(t, t->stream);
			NGPB_CUDA_CHECK(cudaMemcpyAsync(t->pixel_bufs[idx], rgba_host, bytes, cudaMemcpyHostToDevice, t->stream));
			return 0;
		}
		// A view registered without masked pixels (no_mask, established by set_image) promises that replacement frames have none either: its
		// pixels are first read by the loss kernel.  The upload runs on the copy stream into a staging buffer — beside the optimizer of the
		// step before and the generator / inference of this one — and tb_front moves it into place right before the loss kernel.
		if (!t->copy_stream) NGPB_CUDA_CHECK(cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking));
		const uint32_t slot = t->staging_next++ & 1u;
		ngp_testbed::ImageStaging& st = t->staging[slot];
		if (st.in_flight) tb_apply_pending_images(t, t->stream);   // more than two frames between training steps: land the earlier ones now
		if (!st.copied) {
			NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&st.copied, cudaEventDisableTiming));
			NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&st.consumed, cudaEventDisableTiming));
			NGPB_CUDA_CHECK(cudaEventRecord(st.consumed, t->stream));
		}
		if (st.cap < bytes) {
			NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));   // the staging buffer's last reader
			if (st.buf) NGPB_CUDA_CHECK(cudaFree(st.buf));
			st.buf = nullptr;
			NGPB_CUDA_CHECK(cudaMalloc(&st.buf, bytes));
			st.cap = bytes;
		}
		NGPB_CUDA_CHECK(cudaStreamWaitEvent(t->copy_stream, st.consumed, 0));   // the previous frame in this buffer has been moved out
		NGPB_CUDA_CHECK(cudaMemcpyAsync(st.buf, rgba_host, bytes, cudaMemcpyHostToDevice, t->copy_stream));
		NGPB_CUDA_CHECK(cudaEventRecord(st.copied, t->copy_stream));
		st.in_flight = true;
		t->pending_images.push_back(ngp_testbed::PendingImage{idx, slot, bytes});
	});
}
int ngp_testbed_sync(ngp_testbed* t) {
	NGPB_TRY({
		tb_apply_pending_images(t, t->stream);
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	});
}

}  // extern "C"
