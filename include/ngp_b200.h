/* ngp_b200.h — C-ABI of libngp_b200.so, the B200-native (sm_100a) drop-in for instant-ngp's NeRF hot path.
 *
 * Two nested boundaries:
 *
 *   B1  op level    — replaces what tiny-cuda-nn's `tcnn::cpp::Module` vtable offers
 *                     (dependencies/tiny-cuda-nn/include/tiny-cuda-nn/cpp_api.h:92-125) plus the NeRF kernels
 *                     Testbed launches directly (src/testbed_nerf.cu).  All pointers are DEVICE pointers owned
 *                     by the caller, matrices are sample-contiguous ("column major [dims x n]", cpp_api.cu:82-83),
 *                     every call is asynchronous on the given stream.
 *   B2  app level   — a `Testbed` handle mirroring the subset of `pyngp.Testbed` (src/python_api.cu:439-853) that
 *                     the NeRF path touches: create_empty_nerf_dataset / set_image / set_camera_*, reload_network_*,
 *                     train, render, density grid, snapshot of params.
 *
 * Conventions
 *   - plain C types only; `stream` is a cudaStream_t passed as void*.
 *   - every function returns 0 on success, non-zero on failure; ngp_last_error() returns the message
 *     (the reference throws std::runtime_error: tiny-cuda-nn/common_host.h:71-111).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with an error.
 */
#ifndef NGP_B200_H
#define NGP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGP_MAX_LEVELS 32u
#define NGP_BATCH_GRANULARITY 256u /* tiny-cuda-nn common.h:246 BATCH_SIZE_GRANULARITY */
#define NGP_NERF_GRIDSIZE 128u     /* nerf_device.cuh:25 */
#define NGP_NERF_CASCADES 8u       /* nerf_device.cuh:30 */
#define NGP_NERF_STEPS 1024u       /* nerf_device.cuh:29 */
#define NGP_LOSS_SCALE 128.0f      /* neural-graphics-primitives/common.h LOSS_SCALE() */

/* ---------------------------------------------------------------------------------------------------------------
 * Descriptors (plain data, filled by ngp_*_desc_init)
 * ------------------------------------------------------------------------------------------------------------- */

/* Multiresolution hash grid — mirrors GridEncodingTemplated's constructor arithmetic
 * (tiny-cuda-nn/encodings/grid.h:673-737) and grid_scale/grid_resolution (common_device.h:886-895). */
typedef struct ngp_grid_desc {
	uint32_t n_levels;
	uint32_t n_features_per_level; /* 2 or 4 */
	uint32_t log2_hashmap_size;
	uint32_t base_resolution;
	float per_level_scale;
	uint32_t n_params;                     /* = offsets[n_levels] * n_features_per_level */
	uint32_t offsets[NGP_MAX_LEVELS + 1];  /* in grid entries, as ParamsOffsetTable */
	uint32_t resolutions[NGP_MAX_LEVELS];
	float scales[NGP_MAX_LEVELS];
} ngp_grid_desc;

/* NerfNetwork (include/neural-graphics-primitives/nerf_network.h:81-101):
 *   pos[3] -HashGrid-> 32 -[64 x n_hidden_density, ReLU]-> 16 (row 0 = raw density)
 *   [density out(16) | SH4(dir)(16)] -[64 x n_hidden_rgb, ReLU]-> 16 (rows 0..2 = raw rgb)
 * Flat parameter order: density MLP | rgb MLP | hash grid (nerf_network.h:357-372); weights row-major [out x in]. */
typedef struct ngp_nerf_desc {
	ngp_grid_desc grid;
	uint32_t n_hidden_density; /* configs/nerf/base.json: 1 */
	uint32_t n_hidden_rgb;     /* configs/nerf/base.json: 2 */
	uint32_t density_mlp_offset;
	uint32_t rgb_mlp_offset;
	uint32_t grid_offset;
	uint32_t n_mlp_params;     /* = grid_offset: "matrix params" for the optimizer (adam.h:150-155) */
	uint32_t n_params;
} ngp_nerf_desc;

typedef enum ngp_activation { NGP_ACT_NONE = 0, NGP_ACT_RELU = 1, NGP_ACT_LOGISTIC = 2, NGP_ACT_EXPONENTIAL = 3 } ngp_activation;
typedef enum ngp_loss_type { NGP_LOSS_L2 = 0, NGP_LOSS_L1 = 1, NGP_LOSS_MAPE = 2, NGP_LOSS_SMAPE = 3, NGP_LOSS_HUBER = 4, NGP_LOSS_LOGL1 = 5, NGP_LOSS_RELATIVE_L2 = 6 } ngp_loss_type;
/* ETrainMode (common.h:47-51).  Rfl / RflRelax are what the reference runs through its fused train_nerf kernel: they change how the
 * per-sample gradients are formed from the composited ray (fused_kernels/train_nerf.cuh:391-410) — Rfl supervises every sample's colour
 * with the radiance-field loss, RflRelax evaluates the loss gradient at the colour the ray would have if the medium behind the sample
 * were opaque — and they composite with that kernel's arithmetic: transmittance = 1 - accumulated weight (:228-230, 363-367), the
 * background shows through unless the ray is opaque (:251), no L1 term on the density (:305).  Nerf is the non-fused kernels' arithmetic
 * (testbed_nerf.cu:926-1140). */
typedef enum ngp_train_mode { NGP_TRAIN_NERF = 0, NGP_TRAIN_RFL = 1, NGP_TRAIN_RFL_RELAX = 2 } ngp_train_mode;
/* Arithmetic flavour of the training-ray march (ngp_nerf_generate_training_samples):
 *   NGP_MATH_DETERMINISTIC  IEEE add / mul / fma only, transcendentals from ngp_detmath.h, no contraction: what the CPU oracle reproduces
 *                           bit for bit (ray ids, per-ray sample counts, coordinates);
 *   NGP_MATH_REFERENCE      the reference BUILD's arithmetic — its expression trees compiled with --use_fast_math (approximate division /
 *                           square root, MUFU log / exp, FMA contraction, CMakeLists.txt:88): the per-ray sample counts of the reference's
 *                           own generate_training_samples_nerf kernel.  What Testbed::train uses. */
typedef enum ngp_math_mode { NGP_MATH_DETERMINISTIC = 0, NGP_MATH_REFERENCE = 1 } ngp_math_mode;
typedef enum ngp_lens_mode { NGP_LENS_PERSPECTIVE = 0, NGP_LENS_OPENCV = 1 } ngp_lens_mode;
typedef enum ngp_image_type { NGP_IMAGE_NONE = 0, NGP_IMAGE_BYTE = 1, NGP_IMAGE_HALF = 2, NGP_IMAGE_FLOAT = 3 } ngp_image_type;
typedef enum ngp_color_space { NGP_COLOR_LINEAR = 0, NGP_COLOR_SRGB = 1 } ngp_color_space;

/* One training view: TrainingImageMetadata + TrainingXForm.start (nerf_device.cuh:45-60, common.h:189-194).
 * xform is the 4x3 camera-to-world matrix in ngp convention, column major: [right | up | forward | origin]. */
typedef struct ngp_train_view {
	const void* pixels; /* device pointer, RGBA, layout per image_type */
	uint32_t image_type;
	int32_t width, height;
	float focal_x, focal_y;
	float principal_x, principal_y; /* in [0,1] */
	uint32_t lens_mode;
	float lens_params[4]; /* k1 k2 p1 p2 */
	float xform[12];
	uint32_t no_mask; /* 1: the image is known to contain no masked-away pixels (negative red / 0x00FF00FF), so the sample
	                     generator need not read it; 0: unknown, pixels are inspected like the reference does */
} ngp_train_view;

/* Exponential-stepping constants, evaluated once on the host with ngp_detmath.h so the device march and the oracle
 * see the same values (nerf_device.cuh:379-421 recomputes them per call). */
typedef struct ngp_march_consts {
	float cone_angle;
	float log1p_c, a, b, at, bt; /* valid when cone_angle > 1e-5 */
} ngp_march_consts;

/* The reference's order is that of its atomics (testbed_nerf.cu:1010): the 32 rays of a warp (consecutive ray ids = pixels of one view)
 * together, warps as they finish. */
typedef enum ngp_compaction_order {
	NGP_COMPACTION_GROUPS = 0,    /* groups of 32 consecutive rays, the groups in a per-step shuffle (two passes + a prefix sum) */
	NGP_COMPACTION_ATOMIC = 1,    /* one atomic per ray as the rays finish: single rays, shortest first (one pass) */
	NGP_COMPACTION_RAY_ORDER = 2  /* ascending ray id */
} ngp_compaction_order;

typedef struct ngp_nerf_train_cfg {
	float aabb_min[3], aabb_max[3];
	uint32_t max_cascade; /* testbed_nerf.cu:2433-2436 */
	ngp_march_consts march;
	uint32_t snap_to_pixel_centers;
	uint32_t random_bg_color;
	uint32_t linear_colors;
	uint32_t color_space;
	float background_color[3];
	uint32_t loss_type;
	uint32_t rgb_activation;     /* default Logistic (testbed.h) */
	uint32_t density_activation; /* default Exponential */
	float near_distance;
	float loss_scale;
	uint32_t train_mode; /* ngp_train_mode; NGP_TRAIN_NERF is what the reference runs without JIT fusion (testbed_nerf.cu:3091-3093) */
	uint32_t ray_stride; /* sample generator: global ray id = ray_offset + ray_stride * local index (0 reads as 1).  Data parallel, rank r of W
	                        takes ids r, r + W, r + 2W, ... so that every rank draws from every training view */
	uint32_t math_mode;  /* ngp_math_mode of the sample generator's march */
	uint32_t gen_lanes_per_ray; /* sample generator: lanes of a warp that march one ray together (1, 2, 4, 8, 16 or 32); 0 = chosen from the
	                               batch size.  Changes the schedule only: every ray's samples are the same for every value */
	uint32_t gen_walk_empty;    /* sample generator: empty cells the lane that met one crosses on its own before the group's next round (>= 1);
	                               0 = library default.  Schedule only */
	uint32_t gen_speculation;   /* sample generator: samples a group speculates on in the first round after a skip (1 .. lanes per ray; doubles
	                               after every fully occupied round); 0 = library default.  Schedule only */
	uint32_t compaction_order;  /* ngp_compaction_order: the order in which rays take their slots in the compacted batch, i.e. which rays are cut when
	                               a step's samples exceed the batch and which are repeated when they fall short.  A ray's gradients do not depend on it */
	const float* cam_exposure;      /* per-image exposure, 3 floats per training view: the view's colour is multiplied by 2^exposure per channel before it
	                                   becomes the ray's target (testbed_nerf.cu:979-995).  Device memory (host memory for the oracle); NULL = none */
	float* cam_exposure_gradient;   /* 3 floats per view, accumulated by the loss kernel with atomics (testbed_nerf.cu:1142-1155: the reference's
	                                   expression, which leaves out the pixel colour); NULL = exposure is not being optimised */
} ngp_nerf_train_cfg;

/* Counters written by the training sample generator / loss kernel (NerfCounters, testbed.h). */
typedef struct ngp_nerf_counters {
	uint32_t n_rays;              /* rays that produced >= 1 sample */
	uint32_t n_samples;           /* before compaction */
	uint32_t n_samples_compacted; /* after compaction, not clamped */
	uint32_t pad;
} ngp_nerf_counters;

typedef struct ngp_adam_cfg {
	float learning_rate, beta1, beta2, epsilon, l2_reg; /* l2_reg on matrix params only (adam.h:91-95) */
	float loss_scale;
	float ema_decay;    /* ema.h */
	uint32_t ema_step;  /* 1-based optimizer step, for EMA debiasing */
	uint32_t optimize_matrix_params, optimize_non_matrix_params;
} ngp_adam_cfg;

const char* ngp_last_error(void);
int ngp_version(void);
int ngp_device_count(void);

/* ---------------------------------------------------------------------------------------------------------------
 * B1 — descriptors and parameter initialisation (host only)
 * ------------------------------------------------------------------------------------------------------------- */

/* per_level_scale <= 0: derive as Testbed::reset_network does (src/testbed.cu:4241-4255) from desired_resolution
 * (2048) * aabb_scale. */
int ngp_grid_desc_init(ngp_grid_desc* g, uint32_t n_levels, uint32_t n_features_per_level, uint32_t log2_hashmap_size,
	uint32_t base_resolution, float per_level_scale, uint32_t aabb_scale);
int ngp_nerf_desc_init(ngp_nerf_desc* d, const ngp_grid_desc* g, uint32_t n_hidden_density, uint32_t n_hidden_rgb);
/* cone_angle_constant = aabb_scale <= 1 ? 0 : 1/256 (testbed_nerf.cu:2440). */
int ngp_march_consts_init(ngp_march_consts* m, float cone_angle);
/* Trainer::initialize_params (trainer.h:69-87): Xavier-uniform MLPs from a host pcg32, hash grid U(-1e-4, 1e-4) with the
 * GPU fill pattern of random.h:40-67.  Writes fp32 master params (host pointer). */
int ngp_nerf_init_params_host(const ngp_nerf_desc* d, uint64_t seed, float* params_fp32_host);

/* ---------------------------------------------------------------------------------------------------------------
 * B1 — network ops (device pointers)
 * ------------------------------------------------------------------------------------------------------------- */

/* ≙ NerfNetwork::inference_mixed_precision (nerf_network.h:105-139) fused into one kernel.
 * coords: n x 7 floats (NerfCoordinate: pos.xyz, dt, dir.xyz; nerf_device.cuh:176-202), params: fp16 flat buffer,
 * out: n x out_stride halves, columns 0..2 raw rgb, 3 raw density (out_stride = 16 reproduces padded_output_width). */
int ngp_nerf_inference(const ngp_nerf_desc* d, void* stream, uint32_t n, const float* coords, const void* params_fp16,
	void* out_fp16, uint32_t out_stride);

/* The same network evaluation restricted to what compute_loss_kernel_train_nerf will read: rays are taken from
 * numsteps[2*r] = (count, base) for r < counters->n_rays, their samples evaluated in order, 8 at a time, until the ray is
 * exhausted or its transmittance drops below 1e-4 (testbed_nerf.cu:926-929).  out rows that are evaluated are bit-identical
 * to ngp_nerf_inference(out_stride = 4); rows that the loss kernel never reads are left untouched.
 * queue: a zeroed device uint32 (work-queue head).  n_rays_max bounds the launch (>= counters->n_rays).  train_mode (ngp_train_mode): whose
 * transmittance decides where a ray stops — the running product of the loss kernel (Nerf) or one minus the accumulated weight of the fused
 * train kernel (Rfl, RflRelax; fused_kernels/train_nerf.cuh:228-238); ngp_nerf_compute_loss follows the same rule. */
int ngp_nerf_inference_rays(const ngp_nerf_desc* d, void* stream, uint32_t n_rays_max, const ngp_nerf_counters* counters_dev, uint32_t* queue_dev,
	const uint32_t* numsteps, const float* coords, const void* params_fp16, uint32_t density_activation, void* out_fp16, uint32_t train_mode);

/* ≙ NerfNetwork::density (nerf_network.h:270-280): hash grid + density MLP only. positions: n x pos_stride floats
 * (NerfPosition, pos_stride >= 3), out: n halves (raw density). */
int ngp_nerf_density(const ngp_nerf_desc* d, void* stream, uint32_t n, const float* positions, uint32_t pos_stride,
	const void* params_fp16, void* out_fp16);

/* ≙ Trainer::training_step with external dL/dy (trainer.h:254-357 → NerfNetwork::forward_impl/backward_impl):
 * forward + backward over n samples (n % 128 == 0), accumulating dL/dparams into grads_fp16 (same flat layout as
 * params; fp16 like the reference's grad_t for F >= 2, grid.h:660-671).  The caller zeroes grads (ngp_optimizer_step
 * does it as it consumes them).  dL_dout: n x 4 halves (d rgb raw x3, d density raw), already loss-scaled.
 * out_fp16 (optional, may be NULL): n x 4 halves forward output. */
int ngp_nerf_forward_backward(const ngp_nerf_desc* d, void* stream, uint32_t n, const float* coords, const void* params_fp16,
	const void* dL_dout_fp16, void* grads_fp16, void* out_fp16);

/* Profiling only (tools/prof_mlp_phase.py, SURVEY §8d "tensor-pipe % for the MLP phase separately"): ngp_nerf_forward_backward's kernel with the
 * hash-grid gather replaced by a register pattern and the scatter dropped — the 15 tcgen05 MMA groups of every 128-sample tile and their
 * epilogues alone.  mlp_scratch_f32: n_mlp_params floats (accumulated into).  Base network only (F = 2, 1 + 2 hidden layers). */
int ngp_profile_mlp_phase(const ngp_nerf_desc* d, void* stream, uint32_t n, const float* coords, const void* params_fp16, const void* dL_dout_fp16,
	void* grads_fp16, float* mlp_scratch_f32);

/* A/B switch (tests, profiling) of ngp_nerf_forward_backward's run aggregation: consecutive samples of a ray that fall into the same cell of a coarse
 * level are summed inside the warp and issue one set of reductions.  mode 0 = off, 1 = the coarse half of the levels (default), 2 = every level.
 * Process-wide; the result differs only by fp16 summation order. */
void ngp_set_scatter_aggregation(int mode);

/* Hash-grid encoding alone (tests, image/SDF style use): out n x (L*F) halves, sample-contiguous. ≙ kernel_grid (grid.h:48-212). */
int ngp_grid_encode(const ngp_grid_desc* g, void* stream, uint32_t n, const float* positions, uint32_t pos_stride,
	const void* grid_fp16, void* out_fp16);

/* ≙ Ema{ExponentialDecay{Adam}}::step (adam.h:48-127, ema.h:44-77): one fused pass; reads and ZEROES grads_fp16. */
int ngp_optimizer_step(const ngp_nerf_desc* d, void* stream, const ngp_adam_cfg* cfg, float* params_fp32, void* params_fp16,
	void* params_ema_fp16, void* grads_fp16, float* first_moments, float* second_moments, uint32_t* param_steps);

/* ---------------------------------------------------------------------------------------------------------------
 * B1 — NeRF ray march / compaction / accumulation (device pointers)
 * ------------------------------------------------------------------------------------------------------------- */

/* ≙ generate_training_samples_nerf (testbed_nerf.cu:691-849).  rng_state/rng_inc: the Testbed pcg32 by value.
 * This call handles rays [ray_offset, ray_offset + n_rays) of a logical batch of n_rays_global rays (single GPU:
 * ray_offset = 0, n_rays_global = n_rays; data parallel: one shard per rank — image selection and the per-ray random
 * stream are keyed on the global ray id, so the union of shards is the batch one GPU would draw).
 * Outputs: ray_indices[n_rays] (global ids), rays (6 floats o,d per ray), numsteps[2*n_rays] (count, base),
 * coords[max_samples x 7].  counters must be zeroed by the caller.  Slot order is warp-ordered here, atomics-ordered in
 * the reference; compare as a map ray_id -> (count, coordinates). */
int ngp_nerf_generate_training_samples(void* stream, uint32_t n_rays, uint32_t ray_offset, uint32_t n_rays_global, uint64_t rng_state, uint64_t rng_inc,
	const ngp_nerf_train_cfg* cfg, const ngp_train_view* views_dev, uint32_t n_views, const uint8_t* density_grid_bitfield,
	uint32_t max_samples, ngp_nerf_counters* counters_dev, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords);

/* ≙ compute_loss_kernel_train_nerf (testbed_nerf.cu:852-1180): composite, loss, compaction, dL/doutput.
 * network_output: n_samples x 4 halves. Writes coords_compacted [max_compacted x 7], dloss [max_compacted x 4 halves],
 * loss_per_ray[n_rays] (may be NULL).  Loss and gradients are normalised by n_rays_global (testbed_nerf.cu:1039,1073). */
int ngp_nerf_compute_loss(void* stream, uint32_t n_rays, uint32_t n_rays_global, uint64_t rng_state, uint64_t rng_inc,
	const ngp_nerf_train_cfg* cfg, const ngp_train_view* views_dev, uint32_t n_views, const void* network_output_fp16,
	uint32_t max_compacted, ngp_nerf_counters* counters_dev, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps,
	const float* coords, float* coords_compacted, void* dloss_fp16, float* loss_per_ray, const float* mean_density_dev);

/* ≙ fill_rollover_and_rescale + fill_rollover (common_device.h:1114-1135, testbed_nerf.cu:3298-3306). */
int ngp_nerf_fill_rollover(void* stream, uint32_t target_batch, const ngp_nerf_counters* counters_dev, float* coords_compacted,
	void* dloss_fp16);

/* ≙ update_density_grid_nerf + update_density_grid_mean_and_bitfield (testbed_nerf.cu:2476-2633).
 * density_grid: 128^3 * (max_cascade+1) floats, bitfield: 128^3 * 8 / 8 bytes, mean: 1 float, scratch sized by
 * ngp_nerf_density_grid_scratch_bytes().  step 0 additionally requires the views (mark_untrained_density_grid). */
size_t ngp_nerf_density_grid_scratch_bytes(uint32_t max_cascade);
int ngp_nerf_update_density_grid(const ngp_nerf_desc* d, void* stream, const ngp_nerf_train_cfg* cfg, const void* params_fp16,
	uint64_t* grid_rng_state_inout, uint64_t grid_rng_inc, uint32_t training_step, uint32_t ema_step, float decay,
	const ngp_train_view* views_dev, uint32_t n_views, float* density_grid, uint8_t* bitfield, float* mean_density, void* scratch);
/* threshold + 7-level max-pool only (after loading a grid). */
int ngp_nerf_update_bitfield(void* stream, uint32_t max_cascade, const float* density_grid, uint8_t* bitfield, float* mean_density);

/* ≙ render_nerf (testbed_nerf.cu:1894-2150), Shade mode, pinhole or OpenCV camera, one sample per pixel.
 * camera: 4x3 column-major camera-to-world (ngp convention); rows [y0, y1) are rendered (tile sharding).
 * rgba: H x W x 4 floats (linear, premultiplied), depth: H x W floats; both full-frame pointers. */
/* ERenderMode (common.h:68-79), values as the reference's enum.  What composite_kernel_nerf / shade_kernel_nerf put into the colour channels
 * (testbed_nerf.cu:641-655, 1355-1372): AO the sample's alpha, Positions (pos - 0.5) / 2 + 0.5, Depth the sample's distance along the view
 * direction x depth_scale, Cost n_steps / 128 with alpha 1, Shade the radiance (sRGB predictions accumulated in linear colour).  Normals
 * (network input gradients), Distortion and Slice (2-D debug views) are not built: ngp_nerf_render refuses them. */
typedef enum ngp_render_mode { NGP_RENDER_AO = 0, NGP_RENDER_SHADE = 1, NGP_RENDER_NORMALS = 2, NGP_RENDER_POSITIONS = 3, NGP_RENDER_DEPTH = 4, NGP_RENDER_DISTORTION = 5,
	NGP_RENDER_COST = 6, NGP_RENDER_SLICE = 7 } ngp_render_mode;
typedef struct ngp_render_cfg {
	int32_t width, height;
	float focal_x, focal_y;   /* pixels */
	float screen_x, screen_y; /* principal point in [0,1] */
	float camera[12];
	float aabb_min[3], aabb_max[3];         /* training aabb */
	float render_aabb_min[3], render_aabb_max[3];
	uint32_t max_cascade;
	ngp_march_consts march;
	uint32_t rgb_activation, density_activation;
	float min_transmittance;  /* 0.01 */
	uint32_t spp_index;       /* sample index: keys the low-discrepancy jitter of each ray's first step (advance_pos_nerf) */
	float near_distance;
	float pixel_offset[2];    /* ld_random_pixel_offset(snap_to_pixel_centers ? 0 : sample_index), random_val.cuh:320-325; the
	                             host evaluates it (ngp_render_pixel_offset): it depends on the sample index only */
	uint32_t lens_mode;       /* ngp_lens_mode of the render camera (Testbed::m_render_lens when m_render_with_lens_distortion) */
	float lens_params[4];
	uint32_t skips_per_tile;  /* tile kernel: voxel skips a ray may spend per tile iteration looking for its next sample before the tile goes ahead
	                             without it (0 = the default, 4).  Schedule only: a ray's samples do not depend on it */
	uint32_t render_mode;     /* ngp_render_mode */
	float depth_scale;        /* Depth mode: 1 / dataset.scale (testbed_nerf.cu:2037) */
	uint32_t math_mode;       /* ngp_math_mode of the march: NGP_MATH_DETERMINISTIC = ngp_detmath.h (bit-exact against the CPU oracle),
	                             NGP_MATH_REFERENCE = the SFU log / exp / reciprocal the reference build's --use_fast_math compiles its
	                             stepping functions to (what Testbed::render uses; ~10x cheaper per empty-voxel skip) */
} ngp_render_cfg;
/* ld_random_pixel_offset (random_val.cuh:320-325): Owen-scrambled Sobol (0, 1) point of `sample_index`, shifted so that index 0 is the
 * pixel centre.  Host arithmetic. */
void ngp_render_pixel_offset(uint32_t sample_index, float* offset_xy);
size_t ngp_nerf_render_scratch_bytes(int32_t width, int32_t rows);
int ngp_nerf_render(const ngp_nerf_desc* d, void* stream, const ngp_render_cfg* cfg, int32_t y0, int32_t y1, const void* params_fp16,
	const uint8_t* density_grid_bitfield, float* rgba, float* depth, void* scratch, uint32_t* n_steps_total_dev);

/* ≙ the render epilogue: CudaRenderBuffer::accumulate / ::tonemap (src/render_buffer.cu:228-262 accumulate_kernel,
 * :264-342 tonemap, :511-545 tonemap_kernel).  accumulate: acc = (acc * sample_count + frame) / (sample_count + 1), the frame first
 * converted to sRGB when color_space is NGP_COLOR_SRGB.  tonemap: blend the (sRGB-specified) background behind the premultiplied
 * colour, back to linear, exposure 2^e, tonemap curve, to the output colour space, optional un-premultiply and clamp. */
typedef enum ngp_tonemap_curve { NGP_TONEMAP_IDENTITY = 0, NGP_TONEMAP_ACES = 1, NGP_TONEMAP_HABLE = 2, NGP_TONEMAP_REINHARD = 3 } ngp_tonemap_curve;
typedef struct ngp_tonemap_cfg {
	float exposure;
	float background_color[4];
	uint32_t color_space;        /* space the accumulation buffer is in (ngp_color_space) */
	uint32_t output_color_space; /* NGP_COLOR_SRGB for render(linear = false) */
	uint32_t tonemap_curve;      /* ngp_tonemap_curve; Identity by default */
	uint32_t clamp_output_color;
	uint32_t unmultiply_alpha;
} ngp_tonemap_cfg;
int ngp_render_accumulate(void* stream, int32_t width, int32_t height, const float* frame_rgba, float* accumulate_rgba, float sample_count, uint32_t color_space);
int ngp_render_tonemap(void* stream, int32_t width, int32_t height, const ngp_tonemap_cfg* cfg, const float* accumulate_rgba, float* out_rgba);

/* ---------------------------------------------------------------------------------------------------------------
 * B1 (fields) — NetworkWithInputEncoding: HashGrid over 2-D or 3-D positions + one FullyFusedMLP
 * (tiny-cuda-nn/include/tiny-cuda-nn/network_with_input_encoding.h:38-170), the model of the image and SDF primitives
 * (src/testbed_image.cu:231-302 train_image, src/testbed_sdf.cu:1578-1619 train_sdf).
 *   pos[D] -HashGrid-> 32 -[64 x n_hidden, ReLU]-> 16 padded (n_output_dims used)
 * Flat parameter order: MLP | hash grid (network_with_input_encoding.h:121-128); weights row-major [out x in].
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct ngp_field_desc {
	ngp_grid_desc grid;
	uint32_t n_pos_dims;    /* 2 (image) or 3 (sdf) */
	uint32_t n_hidden;      /* hidden layers of 64 neurons, 1..4 */
	uint32_t n_output_dims; /* 1..16; output rows are padded to 16 like the reference (m_padded_output_width) */
	uint32_t mlp_offset;    /* 0 */
	uint32_t grid_offset;   /* = n_mlp_params */
	uint32_t n_mlp_params;
	uint32_t n_params;
} ngp_field_desc;

/* GridEncodingTemplated constructor (grid.h:673-737) for n_pos_dims in {2, 3}; per_level_scale must be > 0. */
int ngp_grid_desc_init_nd(ngp_grid_desc* g, uint32_t n_pos_dims, uint32_t n_levels, uint32_t n_features_per_level, uint32_t log2_hashmap_size,
	uint32_t base_resolution, float per_level_scale);
int ngp_field_desc_init(ngp_field_desc* d, const ngp_grid_desc* g, uint32_t n_pos_dims, uint32_t n_hidden, uint32_t n_output_dims);
/* Trainer::initialize_params for this model: MLP Xavier-uniform then grid U(-1e-4, 1e-4), one pcg32 stream. */
int ngp_field_init_params_host(const ngp_field_desc* d, uint64_t seed, float* params_fp32_out);

/* ≙ Module::inference (cpp_api.h:100): positions [n x n_pos_dims] float32 sample-contiguous -> out [n x out_stride] fp16
 * (out_stride >= 16: the padded row; otherwise the first min(out_stride, n_output_dims) columns).  Any n. */
int ngp_field_inference(const ngp_field_desc* d, void* stream, uint32_t n, const float* positions, const void* params_fp16, void* out, uint32_t out_stride);

/* One fused forward + loss + backward kernel (≙ Trainer::training_step, trainer.h:163-357: forward, Loss::evaluate, backward).
 * targets != NULL: [n x n_output_dims] float32, dL/dout from `loss_type` (tcnn losses/{l2,l1,mape,smape,relative_l2}.h) with
 *   n_total = n * n_output_dims; loss_values (optional) [n x n_output_dims] float32 per-element loss terms (their sum is the loss).
 * targets == NULL: dL_dout_ext [n x 16] fp16 is used instead (≙ the external dL/dy form, Module::backward cpp_api.h:108).
 * grads: flat fp16 gradient buffer; hash-grid entries are ACCUMULATED (caller zeroes, or ngp_optimizer_step_flat does), MLP
 * weights overwritten.  out (optional): [n x 16] fp16 network output.  n % 128 == 0 (the reference requires 256, common.h:246). */
int ngp_field_train_step(const ngp_field_desc* d, void* stream, uint32_t n, const float* positions, const float* targets, uint32_t loss_type,
	float loss_scale, const void* dL_dout_ext, const void* params_fp16, void* grads_fp16, float* loss_values, void* out);

/* tcnn::cpp::Module as a C handle (cpp_api.h:92-125; the object tiny-cuda-nn's PyTorch bindings drive): same argument meaning
 * and layouts — input [n x n_input_dims] float32 sample-contiguous, output / dL_doutput [n x 16] fp16 (the padded width),
 * caller-owned device params and gradients.  create: cpp_api.h:121 (HashGrid encoding over 2 or 3 dims + FullyFusedMLP, 64
 * neurons); initialize_params: pcg32{seed}, cpp_api.cu:141-144.  backward overwrites dL_dparams (GradientMode::Overwrite) and
 * needs no forward context (the fused kernel recomputes the tile's forward); dL_dinput must be NULL; n % 128 == 0. */
typedef struct ngp_module ngp_module;
ngp_module* ngp_module_create_network_with_input_encoding(uint32_t n_input_dims, uint32_t n_output_dims, const char* encoding_json, const char* network_json);
void ngp_module_free(ngp_module* m);
uint32_t ngp_module_n_input_dims(const ngp_module* m);
uint32_t ngp_module_n_output_dims(const ngp_module* m); /* padded: 16 */
size_t ngp_module_n_params(const ngp_module* m);
int ngp_module_get_desc(const ngp_module* m, ngp_field_desc* out);
int ngp_module_initialize_params(const ngp_module* m, size_t seed, float* params_fp32, float scale);
int ngp_module_inference(const ngp_module* m, void* stream, uint32_t n, const float* input, void* output, const void* params);
int ngp_module_forward(const ngp_module* m, void* stream, uint32_t n, const float* input, void* output, const void* params);
int ngp_module_backward(const ngp_module* m, void* stream, uint32_t n, float* dL_dinput, const void* dL_doutput, void* dL_dparams, const float* input,
	const void* output, const void* params);

/* ≙ Loss<T>::evaluate (loss.h:44-60) for L2 / L1 / MAPE / SMAPE / RelativeL2: predictions [n x stride] fp16, targets
 * [n x dims] float32 -> values [n x stride] float32, gradients [n x stride] fp16 (zero in the padding). */
int ngp_loss_evaluate(void* stream, uint32_t loss_type, uint32_t n, uint32_t stride, uint32_t dims, float loss_scale, const void* predictions,
	const float* targets, float* values, void* gradients);

/* ≙ Trainer::optimizer_step (trainer.h:155) for any flat parameter vector: the first n_matrix_params entries are MLP weights. */
int ngp_optimizer_step_flat(uint32_t n_matrix_params, uint32_t n_params, void* stream, const ngp_adam_cfg* cfg, float* params_fp32, void* params_fp16,
	void* params_ema_fp16, void* grads_fp16, float* adam_m1, float* adam_m2, uint32_t* param_steps);

/* ≙ train_image's data generation (src/testbed_image.cu:231-283): generate_random_uniform (random.h:40-67) of 2n floats from the
 * pcg32 (rng_state, rng_inc), optional stratify2_kernel (:66-82, n must be 4^k), eval_image_kernel_and_snap (:176-229) on an
 * RGBA image (image_type NGP_IMAGE_FLOAT or NGP_IMAGE_HALF, 4 channels, row-major).  positions [n x 2], targets [n x 3]. */
int ngp_image_generate_training_data(void* stream, uint32_t n, uint64_t rng_state, uint64_t rng_inc, int stratify, const void* image, uint32_t image_type,
	int32_t width, int32_t height, int snap_to_pixel_centers, int linear_colors, float* positions, float* targets);
/* ≙ shuffle<T> (common_device.h:1097-1111): out[i] = in[permute(i / stride + seed, n_elements)] per member. */
int ngp_shuffle(void* stream, uint32_t n_elements, uint32_t stride, uint32_t seed, const float* in, float* out);

/* ---------------------------------------------------------------------------------------------------------------
 * B2 — Testbed handle (≙ pyngp.Testbed, src/python_api.cu:439-853; Testbed::train src/testbed.cu:4561-4647)
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct ngp_testbed ngp_testbed;

ngp_testbed* ngp_testbed_create(int device, void* stream);
void ngp_testbed_destroy(ngp_testbed* t);
/* python_api.cu:444 create_empty_nerf_dataset(n_images, aabb_scale, is_hdr) */
int ngp_testbed_create_empty_nerf_dataset(ngp_testbed* t, uint32_t n_images, uint32_t aabb_scale);
/* python_api.cu:809-853 nerf.training.set_image / set_camera_extrinsics / set_camera_intrinsics.
 * rgba_host: w*h*4 float32, straight alpha, linear colour; stored premultiplied fp16 like the loader
 * (common_device.cuh:698-735). convert_to_ngp applies nerf_matrix_to_ngp (nerf_loader.h:97-116). */
int ngp_testbed_set_image(ngp_testbed* t, uint32_t idx, const float* rgba_host, int32_t w, int32_t h);
/* NerfDataset::set_training_image with EImageDataType::Byte (src/nerf_loader.cu:749-850), what load_nerf stores for 8-bit files:
 * rgba8_host: w*h*4 bytes, sRGB colour + straight alpha, after convert_rgba32 (white / black -> transparent, mask colour ->
 * 0x00FF00FF).  Pixels are converted to linear premultiplied colour on every read (read_rgba, common_device.cuh:846-872). */
int ngp_testbed_set_image_bytes(ngp_testbed* t, uint32_t idx, const uint8_t* rgba8_host, int32_t w, int32_t h);
int ngp_testbed_set_camera_extrinsics(ngp_testbed* t, uint32_t idx, const float* cam_to_world_3x4_rowmajor, int convert_to_ngp);
int ngp_testbed_set_camera_intrinsics(ngp_testbed* t, uint32_t idx, float fx, float fy, float cx, float cy, float k1, float k2,
	float p1, float p2);
/* reload_network_from_json / _from_file (python_api.cu:543-550): the tiny-cuda-nn style config
 * (configs/nerf/base.json); unsupported otypes are rejected. */
int ngp_testbed_reload_network_from_json(ngp_testbed* t, const char* json_text);
int ngp_testbed_reload_network_from_file(ngp_testbed* t, const char* path);
int ngp_testbed_set_seed(ngp_testbed* t, uint64_t seed);
/* Testbed::reset_network(clear_density_grid) (src/testbed.cu:4163-4178; pyngp "reset", python_api.cu:534) */
int ngp_testbed_reset(ngp_testbed* t, int reset_density_grid);
/* one training view as held by the Testbed (what set_camera_to_training_view reads, src/testbed.cu:486-505) */
int ngp_testbed_get_view(ngp_testbed* t, uint32_t idx, ngp_train_view* out);
int ngp_testbed_set_option(ngp_testbed* t, const char* name, double value); /* nerf.training.* knobs by name */
/* Per-image exposure in stops per colour channel (Nerf::Training::cam_exposure, testbed.h:774): optimised while the option
 * nerf.training.optimize_exposure is on (testbed_nerf.cu:2962-3000), applied to the view's colour by the loss kernel (:979).  Setting a value
 * starts that image's optimizer state afresh. */
int ngp_testbed_get_camera_exposure(ngp_testbed* t, uint32_t idx, float* rgb_out);
int ngp_testbed_set_camera_exposure(ngp_testbed* t, uint32_t idx, const float* rgb);
double ngp_testbed_get_option(ngp_testbed* t, const char* name);
/* Testbed::train(batch_size): density-grid prep on the reference's schedule, one training step, optimizer step. */
int ngp_testbed_train(ngp_testbed* t, uint32_t batch_size);
/* Data parallel training, one process per GPU (the reference has none).  Rank 0 calls ngp_dp_unique_id and hands the
 * ngp_dp_unique_id_bytes() bytes to every rank by any means (a torch.distributed / MPI broadcast, a file); every rank then calls
 * ngp_testbed_init_dp.  From then on ngp_testbed_train is the whole data-parallel step: rank r of W marches the global ray ids
 * r, r + W, ... of a batch of W x rays_per_batch rays (every rank draws from every view; image selection and the per-ray random stream are
 * keyed on the global id), the 16-byte counter block is summed over the ranks on a communication stream beside the forward/backward
 * kernel (the rays_per_batch controller stays identical on all ranks), ONE ncclAllReduce sums the flat fp16 gradient buffer (hash grid +
 * MLPs) on the training stream, and every rank runs the identical optimizer step.  NCCL is loaded at run time (libnccl.so.2). */
size_t ngp_dp_unique_id_bytes(void);
int ngp_dp_unique_id(uint8_t* out, size_t capacity);
int ngp_testbed_init_dp(ngp_testbed* t, uint32_t rank, uint32_t world, const uint8_t* unique_id, size_t n_bytes);
/* Render sharded by row tiles: rank r renders rows ngp_dp_rows(r, W, height) of the frame into its own full-size device buffers
 * (ngp_testbed_render_device with y0, y1), then ngp_testbed_gather_rows exchanges the tiles (one ncclBroadcast per rank, in place,
 * on the Testbed's stream): every rank ends up with the whole frame. */
void ngp_dp_rows(uint32_t rank, uint32_t world, int32_t height, int32_t* y0, int32_t* y1);
int ngp_testbed_gather_rows(ngp_testbed* t, int32_t width, int32_t height, float* rgba_dev, float* depth_dev);
/* Split form for data-parallel training with the CALLER's collectives: grads of this rank's ray shard, then (after the caller's all-reduce over
 * ngp_testbed_grads()) the optimizer step.  rank/world partition the global ray batch. */
int ngp_testbed_set_dp(ngp_testbed* t, uint32_t rank, uint32_t world);
int ngp_testbed_train_compute_grads(ngp_testbed* t, uint32_t batch_size);
int ngp_testbed_train_apply_grads(ngp_testbed* t);
/* Finer split that lets the next step's ray generation run beside the gradient all-reduce (the counters are needed early for that):
 *   train_front (generation .. loss)  ->  caller sums ngp_testbed_dp_counters() over ranks  ->  train_back (forward/backward queued,
 *   rays_per_batch controller updated, next generator queued on a side stream behind the backward kernel)  ->  caller all-reduces
 *   ngp_testbed_grads()  ->  train_apply_grads (optimizer).  train_compute_grads = train_front + train_back with the controller
 *   update deferred to train_apply_grads (counters may then be summed any time before it). */
int ngp_testbed_train_front(ngp_testbed* t, uint32_t batch_size);
int ngp_testbed_train_back(ngp_testbed* t);
void* ngp_testbed_grads(ngp_testbed* t);          /* device fp16 [n_params] */
void* ngp_testbed_params(ngp_testbed* t);         /* device fp16 [n_params] (training params) */
void* ngp_testbed_params_inference(ngp_testbed* t);/* device fp16 EMA params */
float* ngp_testbed_params_fp32(ngp_testbed* t);
uint32_t* ngp_testbed_dp_counters(ngp_testbed* t); /* device u32[4]: rays, samples, compacted, pad — summed across ranks by the caller */
uint32_t ngp_testbed_n_params(ngp_testbed* t);
uint32_t ngp_testbed_training_step(ngp_testbed* t);
float ngp_testbed_loss(ngp_testbed* t);
int ngp_testbed_get_counters(ngp_testbed* t, uint32_t* rays_per_batch, uint32_t* measured_batch_size,
	uint32_t* measured_batch_size_before_compaction);
int ngp_testbed_get_desc(ngp_testbed* t, ngp_nerf_desc* out);
/* params exchange (host pointers), the payload of Trainer::serialize (trainer.h:442-455). */
int ngp_testbed_set_params_fp32(ngp_testbed* t, const float* params_host, uint32_t n);
int ngp_testbed_get_params_fp16(ngp_testbed* t, void* params_host_fp16, uint32_t n, int inference_params);
int ngp_testbed_get_density_grid(ngp_testbed* t, float* grid_host, uint32_t n, uint8_t* bitfield_host, uint32_t n_bitfield_bytes);
int ngp_testbed_set_density_grid(ngp_testbed* t, const float* grid_host, uint32_t n);
/* render(width, height, spp=1, linear=true) into host buffers: rgba H*W*4 floats, depth H*W floats (may be NULL).
 * camera_3x4: ngp-convention camera-to-world, row major [3][4]. rows [y0,y1) only (pass 0,height for all). */
int ngp_testbed_render(ngp_testbed* t, int32_t width, int32_t height, const float* camera_3x4_rowmajor, float focal_x, float focal_y,
	float cx, float cy, int32_t y0, int32_t y1, float* rgba_host, float* depth_host, uint32_t* n_steps_total);
/* ≙ Testbed::render(width, height, spp, linear) (python_api.cu:507-519, render_to_cpu :138-210): spp frames with sample indices
 * 0..spp-1 accumulated, then tonemapped (exposure 0, Identity curve, the Testbed's background colour); linear = 0 returns sRGB. */
int ngp_testbed_render_ex(ngp_testbed* t, int32_t width, int32_t height, const float* camera_3x4_rowmajor, float focal_x, float focal_y, float cx, float cy,
	uint32_t spp, int linear, float* rgba_host, float* depth_host);
/* device-resident variant: rgba_dev / depth_dev are device buffers of the full frame. */
int ngp_testbed_render_device(ngp_testbed* t, int32_t width, int32_t height, const float* camera_3x4_rowmajor, float focal_x,
	float focal_y, float cx, float cy, int32_t y0, int32_t y1, float* rgba_dev, float* depth_dev);
/* Snapshots (python_api.cu:563-571; Testbed::save_snapshot / load_snapshot src/testbed.cu:5288-5485).  By extension:
 *   .ingp     the reference's container: gzip(msgpack({network config..., "snapshot": {...}})), same schema (params_binary = fp16
 *             inference weights, density_grid_binary fp16, nerf.dataset metadata, counters, optional optimizer state in the
 *             Ema{ExponentialDecay{Adam}} nesting of the config) — files written by either side load on the other;
 *   .msgpack  the same without gzip;
 *   other     ".ngpb": flat dump of the full training state (fp32 masters, RNG streams) for exact resume. */
int ngp_testbed_save_snapshot(ngp_testbed* t, const char* path);  /* include_optimizer_state = false, compress = true */
int ngp_testbed_save_snapshot_ex(ngp_testbed* t, const char* path, int include_optimizer_state, int compress);
int ngp_testbed_load_snapshot(ngp_testbed* t, const char* path);
/* Testbed::load_network_config for a .json network config (src/testbed.cu:280-309) with "parent" inheritance resolved
 * (merge_parent_network_config :86-97: the parent, loaded relative to the child's directory, patched by the child per RFC 7386).
 * Host only.  ngp_testbed_reload_network_from_file applies the same resolution. */
int ngp_load_network_config(const char* path, char* json_text_out, size_t capacity, size_t* n_out);
/* codec hooks (tests): JSON text <-> msgpack bytes as the snapshot writer / reader encode them, optionally gzip-wrapped */
int ngp_json_to_msgpack(const char* json_text, int gzip, uint8_t* out, size_t capacity, size_t* n_out);
int ngp_msgpack_to_json(const uint8_t* data, size_t n, int gzip, char* out, size_t capacity, size_t* n_out);
int ngp_testbed_sync(ngp_testbed* t);
/* Per-phase device time of Testbed::train, measured with CUDA events on the Testbed stream (the reference only has host
 * wall-clock EMAs m_training_prep_ms / m_training_ms, testbed.h:1023-1027).  Phases: 0 occupancy-grid update, 1 training
 * sample generation, 2 inference over the generated samples, 3 loss + compaction + roll-over, 4 fused forward/backward,
 * 5 optimizer, 6 gradient all-reduce (data parallel through ngp_testbed_init_dp).  get_phase_ms returns the accumulated
 * milliseconds and the number of steps since the last call. */
#define NGP_N_PHASES 7
int ngp_testbed_set_profiling(ngp_testbed* t, int enable);
int ngp_testbed_get_phase_ms(ngp_testbed* t, float* ms_out, uint32_t* n_steps);
/* Streaming data: overwrite training image `idx` (already set once with ngp_testbed_set_image / _set_image_bytes, same size and
 * pixel type: 16 or 4 bytes per pixel) from a host buffer, asynchronously (pinned memory makes the copy truly asynchronous; the
 * buffer must stay valid until the next training step has finished on the device or ngp_testbed_sync has returned).  A view without masked
 * pixels is uploaded on a copy stream of the Testbed's own into a staging buffer, beside whatever the Testbed stream is doing, and
 * becomes visible right before the next training step's loss kernel — the first reader of its pixels — or at ngp_testbed_sync;
 * frames land in call order.  A view that may contain masked pixels (the sample generator reads those) is replaced in stream order. */
int ngp_testbed_update_image_async(ngp_testbed* t, uint32_t idx, const void* rgba_host);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches). */
/* ---------------------------------------------------------------------------------------------------------------
 * B2 (fields) — Testbed(ETestbedMode::Image) and Testbed(ETestbedMode::Sdf) (common.h:149-155; python_api.cu:440-442):
 * train_image (src/testbed_image.cu:231-302) and train_sdf on caller-provided records (override_sdf_training_data,
 * python_api.cu:74-113,551; src/testbed_sdf.cu:1578-1619).  Mesh loading / BVH sampling of the SDF mode are out of scope.
 * ------------------------------------------------------------------------------------------------------------- */
typedef enum ngp_testbed_mode { NGP_MODE_NERF = 0, NGP_MODE_SDF = 1, NGP_MODE_IMAGE = 2 } ngp_testbed_mode;
typedef struct ngp_field_testbed ngp_field_testbed;
ngp_field_testbed* ngp_field_testbed_create(uint32_t mode, int device, void* stream);
void ngp_field_testbed_destroy(ngp_field_testbed* t);
/* image: w*h*4 float32 RGBA, linear colour (≙ load_image's EDataType::Float path, testbed_image.cu:420-470) */
int ngp_field_testbed_set_image(ngp_field_testbed* t, const float* rgba_host, int32_t w, int32_t h);
/* positions [n x 3] already in the unit cube, distances [n] (what override_sdf_training_data stores after its normalisation) */
int ngp_field_testbed_set_sdf_training_data(ngp_field_testbed* t, const float* positions_host, const float* distances_host, uint32_t n);
int ngp_field_testbed_reload_network_from_json(ngp_field_testbed* t, const char* json_text);
/* Testbed::save_snapshot / load_snapshot (src/testbed.cu:5288-5485) for the image / SDF modes: the reference's container (msgpack, gzip-wrapped
 * for ".ingp") with the network config and "snapshot": {n_params, params_type "__half", params_binary = inference weights in the order
 * MLP | encoding, version, mode "image" / "sdf", training_step, loss, aabb}.  Loading rebuilds the network from the stored config. */
int ngp_field_testbed_save_snapshot(ngp_field_testbed* t, const char* path, int include_optimizer_state, int compress);
int ngp_field_testbed_load_snapshot(ngp_field_testbed* t, const char* path);
int ngp_field_testbed_set_seed(ngp_field_testbed* t, uint64_t seed);
/* options: image.training.snap_to_pixel_centers, image.training.linear_colors, image.random_mode_stratified, train_network,
 * train_encoding, loss_scale */
int ngp_field_testbed_set_option(ngp_field_testbed* t, const char* name, double value);
int ngp_field_testbed_train(ngp_field_testbed* t, uint32_t batch_size);
float ngp_field_testbed_loss(const ngp_field_testbed* t);
uint32_t ngp_field_testbed_training_step(const ngp_field_testbed* t);
size_t ngp_field_testbed_n_params(const ngp_field_testbed* t);
int ngp_field_testbed_get_desc(const ngp_field_testbed* t, ngp_field_desc* out);
void* ngp_field_testbed_params(ngp_field_testbed* t);           /* device fp16 */
void* ngp_field_testbed_params_inference(ngp_field_testbed* t); /* device fp16 (EMA weights) */
void* ngp_field_testbed_grads(ngp_field_testbed* t);            /* device fp16 */
int ngp_field_testbed_set_params_fp32(ngp_field_testbed* t, const float* params_host, size_t n);
int ngp_field_testbed_get_params_fp16(ngp_field_testbed* t, void* params_host, size_t n, int inference);
int ngp_field_testbed_evaluate(ngp_field_testbed* t, const float* positions_host, uint32_t n, float* out_host);
int ngp_field_testbed_render_image(ngp_field_testbed* t, int32_t w, int32_t h, float* rgba_host);
int ngp_field_testbed_sync(ngp_field_testbed* t);

uint64_t ngp_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NGP_B200_H */
