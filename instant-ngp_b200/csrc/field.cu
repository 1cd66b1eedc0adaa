// field.cu — the image and SDF primitives: NetworkWithInputEncoding (HashGrid over 2-D / 3-D positions + one FullyFusedMLP)
// with its training step, the image-primitive data generation and the host-side Testbed for these two modes.
//
//   k_field_forward   gather + MLP, 256 threads per 128-sample tile (two threads per sample, like k_nerf_train).
//                     ≙ NetworkWithInputEncoding::inference_mixed_precision_impl (network_with_input_encoding.h:58-68)
//                     = kernel_grid + kernel_mlp_fused in the reference.
//   k_field_train     forward + loss + backward + weight gradients + hash-grid scatter in ONE kernel.
//                     ≙ Trainer::training_step (trainer.h:254-357): forward (activations to HBM), Loss::evaluate kernel,
//                     kernel_mlp_fused_backward, split-K CUTLASS weight-gradient GEMMs, kernel_grid_backward.
//   k_image_*         generate_random_uniform / stratify2_kernel / eval_image_kernel_and_snap (src/testbed_image.cu:66-229).
//
// Compiled with -fmad=false: the loss and the image lookup use only IEEE add/mul/div, so they agree with the numpy oracle
// bit for bit; the hash-grid arithmetic uses explicit fmaf / __hfma2 as everywhere else.
#include <cmath>
#include <cstring>
#include <random>
#include <vector>

#include <fstream>

#include "host_util.h"
#include "march.cuh"
#include "mlp_train.cuh"
#include "msgpack_mini.h"

namespace ngpb {

struct FieldDev {
	LevelMeta levels[MAX_DEV_LEVELS];
	uint32_t n_levels, n_features, n_hidden, n_out;
	uint32_t mlp_off, grid_off, n_mlp_params;
};

struct FieldSmem {
	uint32_t w_bytes;
	uint32_t a0_off;      // [128 x 32] encoding
	uint32_t hid_off;     // hidden activations, `hid_stride` bytes apart (0: one buffer, inference)
	uint32_t hid_stride;
	uint32_t g64_off, g16_off;
	uint32_t bar_off, total;
};
__host__ __device__ inline FieldSmem field_smem_layout(uint32_t n_hidden, bool train) {
	FieldSmem s;
	s.w_bytes = mlp_n_params(n_hidden) * 2u;
	if (train) {
		s.a0_off = s.w_bytes;
		s.hid_off = s.a0_off + TILE * ENC_WIDTH * 2u;
		s.hid_stride = TILE * MLP_WIDTH * 2u;
		s.g64_off = s.hid_off + n_hidden * s.hid_stride;
		s.g16_off = s.g64_off + TILE * MLP_WIDTH * 2u;
		s.bar_off = s.g16_off + TILE * MLP_OUT * 2u;
	} else {
		s.hid_off = s.w_bytes;
		s.a0_off = s.hid_off;  // aliases the hidden buffer (see fwd_smem_layout)
		s.hid_stride = 0;
		s.g64_off = s.g16_off = 0;
		s.bar_off = s.hid_off + TILE * MLP_WIDTH * 2u;
	}
	s.total = s.bar_off + 16u;
	return s;
}
__host__ __device__ inline uint32_t field_tmem_cols(uint32_t n_hidden) {
	uint32_t c = 64;
	for (uint32_t l = 0; l <= n_hidden; ++l) c += wgrad_cols(n_hidden, l);
	return c;
}

// common prologue: weights -> smem, TMEM, barrier
template <uint32_t TMEM_COLS>
__device__ __forceinline__ uint32_t field_setup(const FieldDev& net, const __half* __restrict__ params, uint8_t* smem, const FieldSmem& L, uint32_t tid) {
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8);
	for (uint32_t l = 0; l <= net.n_hidden; ++l)
		stage_weights(params + net.mlp_off + mlp_layer_off(net.n_hidden, l), mlp_layer_out(net.n_hidden, l), mlp_layer_in(net.n_hidden, l),
			smem + mlp_layer_off(net.n_hidden, l) * 2u, tid, TRAIN_THREADS);
	if (tid < 32) umma::tmem_alloc<TMEM_COLS>(tmem_slot);
	if (tid == 0) {
		umma::mbar_init(bar, 1);
		umma::mbar_fence_init();
	}
	umma::fence_before_sync();
	__syncthreads();
	umma::fence_after_sync();
	return *tmem_slot;
}

template <uint32_t F, uint32_t D>
__global__ void __launch_bounds__(TRAIN_THREADS, 3) k_field_forward(
	const __grid_constant__ FieldDev net, const uint32_t n, const float* __restrict__ positions, const __half* __restrict__ params, __half* __restrict__ out,
	const uint32_t out_stride
) {
	extern __shared__ __align__(128) uint8_t smem[];
	const FieldSmem L = field_smem_layout(net.n_hidden, false);
	const uint32_t tid = threadIdx.x, row = tid & (TILE - 1), half = tid >> 7;
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
	const uint32_t tmem_base = field_setup<64>(net, params, smem, L, tid);
	uint32_t phase = 0;
	const __half* grid = params + net.grid_off;
	const uint32_t n_tiles = (n + TILE - 1) / TILE;
	const uint32_t out_cols = out_stride >= MLP_OUT ? MLP_OUT : (out_stride < net.n_out ? out_stride : net.n_out);
	for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		const uint32_t i = tile * TILE + row;
		const bool valid = i < n;
		float x[D];
#pragma unroll
		for (uint32_t d = 0; d < D; ++d) x[d] = valid ? positions[(size_t)i * D + d] : 0.5f;
		{
			__half2 enc[8];
			grid_gather_half_nd<F, D>(net, grid, half, x, enc);
			const __half2 h0[4] = {enc[0], enc[1], enc[2], enc[3]};
			const __half2 h1[4] = {enc[4], enc[5], enc[6], enc[7]};
			store_chunk(smem + L.a0_off, row, 2 * half + 0, h0);
			store_chunk(smem + L.a0_off, row, 2 * half + 1, h1);
		}
		__half2 o[8];
		run_mlp_fwd_keep(smem, L.a0_off, L.hid_off, 0, net.n_hidden, tmem_base, bar, phase, tid, o, 0u);
		if (valid) {
			__half* dst = out + (size_t)i * out_stride;
			if (out_cols == MLP_OUT && (out_stride & 7u) == 0u) {
				uint4 v;
				v.x = *reinterpret_cast<const uint32_t*>(&o[4 * half + 0]);
				v.y = *reinterpret_cast<const uint32_t*>(&o[4 * half + 1]);
				v.z = *reinterpret_cast<const uint32_t*>(&o[4 * half + 2]);
				v.w = *reinterpret_cast<const uint32_t*>(&o[4 * half + 3]);
				*reinterpret_cast<uint4*>(dst + 8 * half) = v;
			} else {
#pragma unroll
				for (uint32_t j = 0; j < 8; ++j) {
					const uint32_t col = 8 * half + j;
					const __half2 p = o[col >> 1];
					if (col < out_cols) dst[col] = (col & 1u) ? __high2half(p) : __low2half(p);
				}
			}
		}
	}
	umma::fence_before_sync();
	__syncthreads();
	if (tid < 32) umma::tmem_dealloc<64>(tmem_base);
}

// tcnn losses (losses/l2.h:36-71, l1.h, mape.h:40-77, smape.h, relative_l2.h) with pdf = 1: returns the loss-scaled gradient
// (before its cast to fp16) and the per-element loss term.
__host__ __device__ inline float field_loss(uint32_t type, float prediction, float target, float loss_scale, float n_total, float& value) {
	const float difference = prediction - target;
	float gradient;
	switch (type) {
		case NGP_LOSS_L1:
			value = fabsf(difference) / n_total;
			gradient = copysignf(1.0f, difference);
			break;
		case NGP_LOSS_MAPE: {
			const float scale = 1.0f / (fabsf(target) + 1e-2f);
			value = fabsf(difference) * scale / n_total;
			gradient = copysignf(scale, difference);
		} break;
		case NGP_LOSS_SMAPE: {
			const float scale = 1.0f / (0.5f * (fabsf(target) + fabsf(prediction)) + 1e-2f);
			value = fabsf(difference) * scale / n_total;
			gradient = copysignf(scale, difference);
		} break;
		case NGP_LOSS_RELATIVE_L2: {
			const float psq = prediction * prediction + 0.01f;
			value = difference * difference / psq / n_total;
			gradient = 2.0f * difference / psq;
		} break;
		default:  // NGP_LOSS_L2
			value = difference * difference / n_total;
			gradient = 2.0f * difference;
			break;
	}
	return loss_scale * gradient / n_total;
}

template <uint32_t F, uint32_t D, uint32_t TMEM_COLS>
__global__ void __launch_bounds__(TRAIN_THREADS, 2) k_field_train(
	const __grid_constant__ FieldDev net, const uint32_t n, const float* __restrict__ positions, const float* __restrict__ targets, const uint32_t loss_type,
	const float loss_scale, const __half* __restrict__ dL_dout_ext, const __half* __restrict__ params, __half* __restrict__ grads,
	float* __restrict__ mlp_grads_f32, float* __restrict__ loss_values, __half* __restrict__ out
) {
	extern __shared__ __align__(128) uint8_t smem[];
	const uint32_t nh = net.n_hidden;
	const FieldSmem L = field_smem_layout(nh, true);
	const uint32_t tid = threadIdx.x, row = tid & (TILE - 1), half = tid >> 7;
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
	const uint32_t tmem_base = field_setup<TMEM_COLS>(net, params, smem, L, tid);
	uint32_t phase = 0;

	const __half* grid = params + net.grid_off;
	__half* grid_grad = grads + net.grid_off;
	const uint32_t n_tiles = n / TILE;
	const float n_total = (float)(n * net.n_out);
	uint32_t iter = 0;
	for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
		const uint32_t i = tile * TILE + row;
		float x[D];
#pragma unroll
		for (uint32_t d = 0; d < D; ++d) x[d] = positions[(size_t)i * D + d];
		{
			__half2 enc[8];
			grid_gather_half_nd<F, D>(net, grid, half, x, enc);
			const __half2 h0[4] = {enc[0], enc[1], enc[2], enc[3]};
			const __half2 h1[4] = {enc[4], enc[5], enc[6], enc[7]};
			store_chunk(smem + L.a0_off, row, 2 * half + 0, h0);
			store_chunk(smem + L.a0_off, row, 2 * half + 1, h1);
		}
		__half2 o[8];
		run_mlp_fwd_keep(smem, L.a0_off, L.hid_off, 0, nh, tmem_base, bar, phase, tid, o);
		if (out) {
			uint4 v;
			v.x = *reinterpret_cast<const uint32_t*>(&o[4 * half + 0]);
			v.y = *reinterpret_cast<const uint32_t*>(&o[4 * half + 1]);
			v.z = *reinterpret_cast<const uint32_t*>(&o[4 * half + 2]);
			v.w = *reinterpret_cast<const uint32_t*>(&o[4 * half + 3]);
			*reinterpret_cast<uint4*>(out + (size_t)i * MLP_OUT + 8 * half) = v;
		}

		// ---------------- dL/d(output): this thread's 8 columns [8*half, 8*half+8)
		{
			__half2 gl[4];
			if (targets) {
				__half gh[8];
#pragma unroll
				for (uint32_t j = 0; j < 8; ++j) {
					const uint32_t col = 8 * half + j;
					gh[j] = __float2half_rn(0.0f);
					if (col < net.n_out) {
						const __half2 p = o[col >> 1];
						const float prediction = __half2float((col & 1u) ? __high2half(p) : __low2half(p));
						float value;
						const float g = field_loss(loss_type, prediction, targets[(size_t)i * net.n_out + col], loss_scale, n_total, value);
						gh[j] = __float2half_rn(g);
						if (loss_values) loss_values[(size_t)i * net.n_out + col] = value;
					}
				}
#pragma unroll
				for (uint32_t j = 0; j < 4; ++j) gl[j] = __halves2half2(gh[2 * j], gh[2 * j + 1]);
			} else {
				const uint4 v = __ldg(reinterpret_cast<const uint4*>(dL_dout_ext + (size_t)i * MLP_OUT + 8 * half));
				gl[0] = *reinterpret_cast<const __half2*>(&v.x);
				gl[1] = *reinterpret_cast<const __half2*>(&v.y);
				gl[2] = *reinterpret_cast<const __half2*>(&v.z);
				gl[3] = *reinterpret_cast<const __half2*>(&v.w);
			}
			store_chunk(smem + L.g16_off, row, half, gl);
		}

		// ---------------- backward
		float dx[16];
		run_mlp_bwd(smem, L.a0_off, L.hid_off, L.g64_off, L.g16_off, 0, nh, tmem_base, 64u, iter > 0 ? 1u : 0u, bar, phase, tid, dx);
		{
			__half2 g[8];
#pragma unroll
			for (uint32_t j = 0; j < 8; ++j) g[j] = __floats2half2_rn(dx[2 * j], dx[2 * j + 1]);
			grid_scatter_half_nd<F, D>(net, grid_grad, half, x, g);
		}
	}

	// ---------------- flush the weight-gradient accumulators (see k_nerf_train)
	umma::fence_before_sync();
	__syncthreads();
	umma::fence_after_sync();
	if (iter > 0) {
		const uint32_t warp4 = (tid >> 5) & 3u, lane = tid & 31u;
		const uint32_t wrow = warp4 * 16u + lane;  // valid for lane < 16
		const uint32_t lane_taddr = tmem_base + ((row & ~31u) << 16);
		uint32_t group = 0, cbase = 64;
		for (uint32_t l = 0; l <= nh; ++l) {
			const uint32_t cols = wgrad_cols(nh, l);
			const uint32_t K = mlp_layer_in(nh, l);
			float* dst = mlp_grads_f32 + net.mlp_off + mlp_layer_off(nh, l);
			for (uint32_t c0 = 0; c0 < cols; c0 += 16, ++group) {
				if ((group & 1u) != half) continue;  // warp-uniform
				uint32_t v[16];
				umma::tmem_ld16(lane_taddr + cbase + c0, v);
				umma::tmem_ld_wait();
				if (lane < 16) {
#pragma unroll
					for (uint32_t j = 0; j < 16; ++j) {
						const uint32_t col = c0 + j;
						const uint32_t idx = (l == nh) ? (col * MLP_WIDTH + wrow) : (wrow * K + col);
						atomicAdd(dst + idx, __uint_as_float(v[j]));
					}
				}
			}
			cbase += cols;
		}
	}
	umma::fence_before_sync();
	__syncthreads();
	if (tid < 32) umma::tmem_dealloc<TMEM_COLS>(tmem_base);
}

__global__ void k_field_grads_finalize(const uint32_t n, float* __restrict__ src, __half* __restrict__ dst) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	dst[i] = __float2half_rn(src[i]);
	src[i] = 0.0f;
}

// ≙ l2_loss & friends as a standalone kernel (loss.h:44-60 Loss::evaluate)
__global__ void k_loss_evaluate(const uint32_t n_elements, const uint32_t stride, const uint32_t dims, const uint32_t loss_type, const float loss_scale,
	const __half* __restrict__ predictions, const float* __restrict__ targets, float* __restrict__ values, __half* __restrict__ gradients) {
	const uint32_t i = threadIdx.x + blockIdx.x * blockDim.x;
	if (i >= n_elements) return;
	const uint32_t intra = i % stride, inter = i / stride;
	if (intra >= dims) {
		values[i] = 0.0f;
		gradients[i] = __float2half_rn(0.0f);
		return;
	}
	const float n_total = (float)(n_elements / stride * dims);
	float value;
	const float g = field_loss(loss_type, __half2float(predictions[i]), targets[inter * dims + intra], loss_scale, n_total, value);
	values[i] = value;
	gradients[i] = __float2half_rn(g);
}

// ---- image primitive data -------------------------------------------------------------------------------------------
// generate_random_kernel (random.h:40-54): thread i advances the stream by 4 i and writes indices i + n_threads * j
__global__ void k_random_uniform(const uint32_t n_elements, Pcg32 rng, float* __restrict__ out) {
	const uint32_t i = threadIdx.x + blockIdx.x * blockDim.x;
	const uint32_t n_threads = blockDim.x * gridDim.x;
	rng.advance((uint64_t)i * 4);
#pragma unroll
	for (uint32_t j = 0; j < 4; ++j) {
		const uint64_t idx = (uint64_t)i + (uint64_t)n_threads * j;
		if (idx >= n_elements) return;
		out[idx] = rng.next_float() * (1.0f - 0.0f) + 0.0f;
	}
}

__global__ void k_stratify2(const uint32_t n_elements, const uint32_t log2_batch_size, float* __restrict__ inout) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	const uint32_t log2_size = log2_batch_size / 2;
	const uint32_t size = 1u << log2_size;
	const uint32_t in_batch = i & ((1u << log2_batch_size) - 1u);
	const uint32_t x = in_batch & ((1u << log2_size) - 1u);
	const uint32_t y = in_batch >> log2_size;
	const float vx = inout[2 * i + 0], vy = inout[2 * i + 1];
	inout[2 * i + 0] = vx / (float)size + ((float)x / (float)size);
	inout[2 * i + 1] = vy / (float)size + ((float)y / (float)size);
}

template <typename T>
__device__ __forceinline__ void image_read(const T* __restrict__ tex, int w, int x, int y, bool linear_colors, float (&c)[3]) {
	const T* p = tex + ((size_t)y * w + x) * 4;
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		c[k] = (float)p[k];
		if (!linear_colors) c[k] = linear_to_srgb(c[k]);
	}
}

template <typename T>
__global__ void k_eval_image_and_snap(const uint32_t n_elements, const T* __restrict__ tex, float* __restrict__ positions, const int w, const int h,
	float* __restrict__ result, const bool snap_to_pixel_centers, const bool linear_colors) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	float px = positions[2 * i + 0], py = positions[2 * i + 1];
	float val[3];
	if (snap_to_pixel_centers) {
		int ix = (int)floorf(px * (float)w), iy = (int)floorf(py * (float)h);
		positions[2 * i + 0] = ((float)ix + 0.5f) / (float)w;
		positions[2 * i + 1] = ((float)iy + 0.5f) / (float)h;
		ix = clampi(ix, 0, w - 1);
		iy = clampi(iy, 0, h - 1);
		image_read(tex, w, ix, iy, linear_colors, val);
	} else {
		px = clampf(px * (float)w - 0.5f, 0.0f, (float)w - (1.0f + 1e-4f));
		py = clampf(py * (float)h - 0.5f, 0.0f, (float)h - (1.0f + 1e-4f));
		const int ix = (int)px, iy = (int)py;
		const float wx = px - (float)ix, wy = py - (float)iy;
		const int jx = clampi(ix, 0, w - 2), jy = clampi(iy, 0, h - 2);
		float c00[3], c10[3], c01[3], c11[3];
		image_read(tex, w, jx, jy, linear_colors, c00);
		image_read(tex, w, jx + 1, jy, linear_colors, c10);
		image_read(tex, w, jx, jy + 1, linear_colors, c01);
		image_read(tex, w, jx + 1, jy + 1, linear_colors, c11);
		const float w00 = (1.0f - wx) * (1.0f - wy), w10 = wx * (1.0f - wy), w01 = (1.0f - wx) * wy, w11 = wx * wy;
#pragma unroll
		for (int k = 0; k < 3; ++k) val[k] = ((w00 * c00[k] + w10 * c10[k]) + w01 * c01[k]) + w11 * c11[k];
	}
	result[3 * i + 0] = val[0];
	result[3 * i + 1] = val[1];
	result[3 * i + 2] = val[2];
}

__host__ __device__ inline uint32_t permute_u32(uint32_t num, uint32_t size) {
	const uint32_t A = 1434869437u, B = 2097192037u;
	return (uint32_t)(((uint64_t)num * A + B) % size);
}
__global__ void k_shuffle(const uint32_t n_elements, const uint32_t stride, const uint32_t seed, const float* __restrict__ in, float* __restrict__ out) {
	const uint32_t i = threadIdx.x + blockIdx.x * blockDim.x;
	if (i >= n_elements * stride) return;
	const uint32_t elem = i / stride, member = i % stride;
	out[i] = in[permute_u32(elem + seed, n_elements) * stride + member];
}

// pixel-centre query positions of a w x h frame and the shade pass of render_image (src/testbed_image.cu:140-174)
__global__ void k_image_pixel_coords(const int w, const int h, float* __restrict__ positions) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (uint32_t)w * h) return;
	const uint32_t x = i % w, y = i / w;
	positions[2 * i + 0] = ((float)x + 0.5f) / (float)w;
	positions[2 * i + 1] = ((float)y + 0.5f) / (float)h;
}
__global__ void k_image_shade(const uint32_t n, const __half* __restrict__ colors, const uint32_t stride, const bool linear_colors, float* __restrict__ rgba) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		float c = __half2float(colors[(size_t)i * stride + k]);
		if (!linear_colors) c = srgb_to_linear(c);
		rgba[4 * (size_t)i + k] = c;
	}
	rgba[4 * (size_t)i + 3] = 1.0f;
}

// ----------------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------------
static uint32_t powi_u32(uint32_t b, uint32_t e) {
	uint32_t r = 1;
	for (uint32_t i = 0; i < e; ++i) r *= b;
	return r;
}

void grid_desc_init_nd(ngp_grid_desc* g, uint32_t n_pos_dims, uint32_t n_levels, uint32_t F, uint32_t log2_hashmap_size, uint32_t base_resolution,
	float per_level_scale) {
	NGPB_CHECK(n_pos_dims == 2 || n_pos_dims == 3, "HashGrid: n_pos_dims must be 2 or 3");
	NGPB_CHECK(n_levels >= 1 && n_levels <= NGP_MAX_LEVELS, "HashGrid: n_levels out of range");
	NGPB_CHECK(F == 1 || F == 2 || F == 4 || F == 8, "HashGrid: n_features_per_level must be 1, 2, 4 or 8");
	NGPB_CHECK(per_level_scale > 0.0f, "HashGrid: per_level_scale must be positive");
	NGPB_CHECK(log2_hashmap_size < 32, "HashGrid: log2_hashmap_size out of range");
	memset(g, 0, sizeof(*g));
	g->n_levels = n_levels;
	g->n_features_per_level = F;
	g->log2_hashmap_size = log2_hashmap_size;
	g->base_resolution = base_resolution;
	g->per_level_scale = per_level_scale;
	const float log2_scale = std::log2(per_level_scale);
	uint32_t offset = 0;
	for (uint32_t l = 0; l < n_levels; ++l) {
		// grid.h:699-722 (GridType::Hash)
		const float scale = exp2f((float)l * log2_scale) * (float)base_resolution - 1.0f;
		const uint32_t res = (uint32_t)ceilf(scale) + 1;
		const uint32_t max_params = 0xFFFFFFFFu / 2;
		uint32_t params_in_level = std::pow((float)res, (float)n_pos_dims) > (float)max_params ? max_params : powi_u32(res, n_pos_dims);
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

void field_desc_init(ngp_field_desc* d, const ngp_grid_desc* g, uint32_t n_pos_dims, uint32_t n_hidden, uint32_t n_output_dims) {
	NGPB_CHECK(n_pos_dims == 2 || n_pos_dims == 3, "field: n_pos_dims must be 2 or 3");
	NGPB_CHECK(n_hidden >= 1 && n_hidden <= MAX_HIDDEN, "FullyFusedMLP: 1..4 hidden layers supported");
	NGPB_CHECK(n_output_dims >= 1 && n_output_dims <= MLP_OUT, "FullyFusedMLP: 1..16 output dims supported");
	NGPB_CHECK(g->n_levels * g->n_features_per_level == ENC_WIDTH, "this build fuses a 32-wide encoding (n_levels * n_features_per_level == 32)");
	memset(d, 0, sizeof(*d));
	d->grid = *g;
	d->n_pos_dims = n_pos_dims;
	d->n_hidden = n_hidden;
	d->n_output_dims = n_output_dims;
	d->mlp_offset = 0;
	d->n_mlp_params = mlp_n_params(n_hidden);
	d->grid_offset = d->n_mlp_params;
	d->n_params = d->n_mlp_params + g->n_params;
}

static bool level_is_dense_nd(uint32_t n_pos_dims, uint32_t resolution, uint32_t size) {
	if (n_pos_dims == 3) return level_is_dense_3d(resolution, size);
	if (resolution > 0xFFFFu) return false;
	return !((uint64_t)size < (uint64_t)resolution * resolution);
}

static FieldDev make_fielddev(const ngp_field_desc& d) {
	FieldDev n{};
	const ngp_grid_desc& g = d.grid;
	NGPB_CHECK(g.n_features_per_level == 2 || g.n_features_per_level == 4, "HashGrid: n_features_per_level must be 2 or 4 in this build");
	NGPB_CHECK(g.n_levels * g.n_features_per_level == ENC_WIDTH, "HashGrid: n_levels * n_features_per_level must be 32");
	NGPB_CHECK(d.n_pos_dims == 2 || d.n_pos_dims == 3, "field: n_pos_dims must be 2 or 3");
	NGPB_CHECK(d.n_hidden >= 1 && d.n_hidden <= MAX_HIDDEN, "FullyFusedMLP: 1..4 hidden layers supported");
	NGPB_CHECK(d.n_output_dims >= 1 && d.n_output_dims <= MLP_OUT, "FullyFusedMLP: 1..16 output dims supported");
	n.n_levels = g.n_levels;
	n.n_features = g.n_features_per_level;
	for (uint32_t l = 0; l < g.n_levels; ++l) {
		LevelMeta& m = n.levels[l];
		m.offset = g.offsets[l];
		m.size = g.offsets[l + 1] - g.offsets[l];
		m.resolution = g.resolutions[l];
		m.scale = g.scales[l];
		m.dense = level_is_dense_nd(d.n_pos_dims, m.resolution, m.size) ? 1u : 0u;
		NGPB_CHECK(m.dense || (m.size & (m.size - 1u)) == 0u, "HashGrid: a hashed level must have a power-of-two size");
	}
	n.n_hidden = d.n_hidden;
	n.n_out = d.n_output_dims;
	n.mlp_off = d.mlp_offset;
	n.grid_off = d.grid_offset;
	n.n_mlp_params = d.n_mlp_params;
	return n;
}

// NetworkWithInputEncoding::initialize_params: the MLP matrices Xavier-uniform on the host (gpu_matrix.h:292-307), then the hash
// grid with generate_random_uniform's fill pattern, all from one pcg32; `scale` multiplies both ranges (cpp_api.cu:141-144).
static void field_init_params_rng(const ngp_field_desc* d, Pcg32 rng, float scale_all, float* out) {
	float* p = out;
	for (uint32_t l = 0; l <= d->n_hidden; ++l) {
		const uint32_t rows = mlp_layer_out(d->n_hidden, l), cols = mlp_layer_in(d->n_hidden, l);
		const float scale = std::sqrt(6.0f / (float)(rows + cols)) * scale_all;
		for (uint32_t i = 0; i < rows * cols; ++i) p[i] = rng.next_float() * 2.0f * scale - scale;
		p += rows * cols;
	}
	const size_t n = d->grid.n_params;
	const size_t n_threads_req = (n + 3) / 4;
	const size_t n_threads = ((n_threads_req + 127) / 128) * 128;
	const float lo = -1e-4f * scale_all, hi = 1e-4f * scale_all;
	for (size_t i = 0; i < n_threads; ++i) {
		if (i >= n) break;
		Pcg32 r = rng;
		r.advance((uint64_t)i * 4);
		for (size_t j = 0; j < 4; ++j) {
			const size_t idx = i + n_threads * j;
			if (idx >= n) break;
			p[idx] = r.next_float() * (hi - lo) + lo;
		}
	}
}
// Trainer::initialize_params (trainer.h:69-87): the trainer's rng comes from std::seed_seq{seed}
void field_init_params_host(const ngp_field_desc* d, uint64_t seed, float* out) {
	std::seed_seq seq{(uint32_t)seed};
	std::vector<uint32_t> seeds(2);
	seq.generate(seeds.begin(), seeds.end());
	field_init_params_rng(d, Pcg32((uint64_t)seeds.front()), 1.0f, out);
}

template <uint32_t F, uint32_t D>
static void launch_field_forward(const FieldDev& net, cudaStream_t stream, uint32_t n, const float* positions, const __half* params, __half* out, uint32_t out_stride) {
	const FieldSmem L = field_smem_layout(net.n_hidden, false);
	auto kern = k_field_forward<F, D>;
	static bool attr_set = false;
	if (!attr_set) {
		NGPB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		attr_set = true;
	}
	const uint32_t n_tiles = div_round_up(n, TILE);
	const uint32_t max_ctas = (uint32_t)device_sm_count() * 3u;
	const uint32_t grid = n_tiles < max_ctas ? n_tiles : max_ctas;
	kern<<<grid, TRAIN_THREADS, L.total, stream>>>(net, n, positions, params, out, out_stride);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

void field_inference(const ngp_field_desc& d, cudaStream_t stream, uint32_t n, const float* positions, const __half* params, __half* out, uint32_t out_stride) {
	NGPB_CHECK(out_stride >= 1, "ngp_field_inference: out_stride must be >= 1");
	if (n == 0) return;
	const FieldDev net = make_fielddev(d);
	if (net.n_features == 2) {
		if (d.n_pos_dims == 2) launch_field_forward<2, 2>(net, stream, n, positions, params, out, out_stride);
		else launch_field_forward<2, 3>(net, stream, n, positions, params, out, out_stride);
	} else {
		if (d.n_pos_dims == 2) launch_field_forward<4, 2>(net, stream, n, positions, params, out, out_stride);
		else launch_field_forward<4, 3>(net, stream, n, positions, params, out, out_stride);
	}
}

template <uint32_t F, uint32_t D, uint32_t TMEM_COLS>
static void launch_field_train(const FieldDev& net, cudaStream_t stream, uint32_t n, const float* positions, const float* targets, uint32_t loss_type,
	float loss_scale, const __half* dL_dout_ext, const __half* params, __half* grads, float* mlp_grads_f32, float* loss_values, __half* out) {
	const FieldSmem L = field_smem_layout(net.n_hidden, true);
	auto kern = k_field_train<F, D, TMEM_COLS>;
	static bool attr_set = false;
	if (!attr_set) {
		NGPB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		attr_set = true;
	}
	const uint32_t n_tiles = n / TILE;
	uint32_t per_sm = (227u * 1024u) / (L.total + 1024u);
	if (per_sm > 512u / TMEM_COLS) per_sm = 512u / TMEM_COLS;
	if (per_sm > 2) per_sm = 2;
	if (per_sm < 1) per_sm = 1;
	const uint32_t max_ctas = (uint32_t)device_sm_count() * per_sm;
	const uint32_t grid = n_tiles < max_ctas ? n_tiles : max_ctas;
	kern<<<grid, TRAIN_THREADS, L.total, stream>>>(net, n, positions, targets, loss_type, loss_scale, dL_dout_ext, params, grads, mlp_grads_f32, loss_values, out);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

template <uint32_t F, uint32_t D>
static void dispatch_field_train(const FieldDev& net, uint32_t cols, cudaStream_t stream, uint32_t n, const float* positions, const float* targets,
	uint32_t loss_type, float loss_scale, const __half* dL_dout_ext, const __half* params, __half* grads, float* mlp_grads_f32, float* loss_values, __half* out) {
	if (cols <= 128) launch_field_train<F, D, 128>(net, stream, n, positions, targets, loss_type, loss_scale, dL_dout_ext, params, grads, mlp_grads_f32, loss_values, out);
	else if (cols <= 256) launch_field_train<F, D, 256>(net, stream, n, positions, targets, loss_type, loss_scale, dL_dout_ext, params, grads, mlp_grads_f32, loss_values, out);
	else launch_field_train<F, D, 512>(net, stream, n, positions, targets, loss_type, loss_scale, dL_dout_ext, params, grads, mlp_grads_f32, loss_values, out);
}

static void check_field_loss(uint32_t loss_type) {
	NGPB_CHECK(loss_type == NGP_LOSS_L2 || loss_type == NGP_LOSS_L1 || loss_type == NGP_LOSS_MAPE || loss_type == NGP_LOSS_SMAPE ||
		loss_type == NGP_LOSS_RELATIVE_L2, "field loss must be one of L2, L1, MAPE, SMAPE, RelativeL2 (tcnn Loss objects)");
}

// mlp_grads_f32: device scratch of n_mlp_params floats, all zero on entry (left zeroed on exit).
void field_train_step(const ngp_field_desc& d, cudaStream_t stream, uint32_t n, const float* positions, const float* targets, uint32_t loss_type, float loss_scale,
	const __half* dL_dout_ext, const __half* params, __half* grads, float* mlp_grads_f32, float* loss_values, __half* out) {
	NGPB_CHECK(n % TILE == 0, "ngp_field_train_step: batch size must be a multiple of 128");
	NGPB_CHECK(targets || dL_dout_ext, "ngp_field_train_step: targets or an external dL/dout is required");
	if (targets) check_field_loss(loss_type);
	if (n == 0) return;
	const FieldDev net = make_fielddev(d);
	const uint32_t cols = field_tmem_cols(net.n_hidden);
	NGPB_CHECK(cols <= 512, "MLP too deep for the training kernel's TMEM budget");
	NGPB_CHECK(field_smem_layout(net.n_hidden, true).total <= 227 * 1024, "MLP too deep for the training kernel's shared memory budget");
	if (net.n_features == 2) {
		if (d.n_pos_dims == 2) dispatch_field_train<2, 2>(net, cols, stream, n, positions, targets, loss_type, loss_scale, dL_dout_ext, params, grads, mlp_grads_f32, loss_values, out);
		else dispatch_field_train<2, 3>(net, cols, stream, n, positions, targets, loss_type, loss_scale, dL_dout_ext, params, grads, mlp_grads_f32, loss_values, out);
	} else {
		if (d.n_pos_dims == 2) dispatch_field_train<4, 2>(net, cols, stream, n, positions, targets, loss_type, loss_scale, dL_dout_ext, params, grads, mlp_grads_f32, loss_values, out);
		else dispatch_field_train<4, 3>(net, cols, stream, n, positions, targets, loss_type, loss_scale, dL_dout_ext, params, grads, mlp_grads_f32, loss_values, out);
	}
	k_field_grads_finalize<<<div_round_up(d.n_mlp_params, 256), 256, 0, stream>>>(d.n_mlp_params, mlp_grads_f32, grads + d.mlp_offset);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

void loss_evaluate(cudaStream_t stream, uint32_t loss_type, uint32_t n, uint32_t stride, uint32_t dims, float loss_scale, const __half* predictions,
	const float* targets, float* values, __half* gradients) {
	check_field_loss(loss_type);
	NGPB_CHECK(dims >= 1 && dims <= stride, "ngp_loss_evaluate: need 1 <= dims <= stride");
	if (n == 0) return;
	const uint32_t n_elements = n * stride;
	k_loss_evaluate<<<div_round_up(n_elements, 128), 128, 0, stream>>>(n_elements, stride, dims, loss_type, loss_scale, predictions, targets, values, gradients);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

void image_generate_training_data(cudaStream_t stream, uint32_t n, uint64_t rng_state, uint64_t rng_inc, bool stratify, const void* image, uint32_t image_type,
	int32_t w, int32_t h, bool snap, bool linear_colors, float* positions, float* targets) {
	NGPB_CHECK(image_type == NGP_IMAGE_FLOAT || image_type == NGP_IMAGE_HALF, "image primitive: image_type must be Float or Half");
	NGPB_CHECK(w >= 2 && h >= 2, "image primitive: image must be at least 2x2");
	if (n == 0) return;
	const uint32_t n_elements = n * 2;
	const uint32_t n_threads = div_round_up(n_elements, 4u);
	k_random_uniform<<<div_round_up(n_threads, 128), 128, 0, stream>>>(n_elements, Pcg32(rng_state, rng_inc, true), positions);
	NGPB_LAUNCHED();
	if (stratify) {
		uint32_t log2_n = 0;
		while ((1u << log2_n) < n) ++log2_n;
		NGPB_CHECK((1u << log2_n) == n && (log2_n % 2) == 0, "stratified sampling needs a square power-of-two batch size");
		k_stratify2<<<div_round_up(n, 128), 128, 0, stream>>>(n, log2_n, positions);
		NGPB_LAUNCHED();
	}
	if (image_type == NGP_IMAGE_FLOAT) {
		k_eval_image_and_snap<float><<<div_round_up(n, 128), 128, 0, stream>>>(n, (const float*)image, positions, w, h, targets, snap, linear_colors);
	} else {
		k_eval_image_and_snap<__half><<<div_round_up(n, 128), 128, 0, stream>>>(n, (const __half*)image, positions, w, h, targets, snap, linear_colors);
	}
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

void shuffle_f32(cudaStream_t stream, uint32_t n_elements, uint32_t stride, uint32_t seed, const float* in, float* out) {
	if (n_elements == 0) return;
	k_shuffle<<<div_round_up(n_elements * stride, 128), 128, 0, stream>>>(n_elements, stride, seed, in, out);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ngpb

// ------------------------------------------------------------------------------------------------------------------
// Testbed for ETestbedMode::Image / ::Sdf (src/testbed.cu:4561-4647 train; testbed_image.cu:231-302; testbed_sdf.cu:1578-1619)
// ------------------------------------------------------------------------------------------------------------------
using namespace ngpb;

struct ngp_field_testbed {
	uint32_t mode = NGP_MODE_IMAGE;
	int device = 0;
	cudaStream_t stream = nullptr;
	uint64_t seed = 1337;
	Pcg32 rng{1337};

	bool has_network = false;
	Json network_config;          // as given to reload_network_from_json (m_network_config)
	ngp_field_desc desc{};
	OptimizerConfig opt;
	uint32_t loss_type = NGP_LOSS_L2;
	float loss_scale = 128.0f;  // default_loss_scale<__half>() (common.h:243)
	float lr_factor = 1.0f;
	uint32_t optimizer_step = 0, training_step = 0;
	float loss_scalar = 0.0f;
	bool train_network = true, train_encoding = true;

	DevBuf<float> params_fp32, m1, m2, mlp_grads_f32;
	DevBuf<__half> params, params_ema, grads;
	DevBuf<uint32_t> param_steps;

	// image primitive (m_image)
	DevBuf<float> image;
	int32_t img_w = 0, img_h = 0;
	bool snap_to_pixel_centers = false, linear_colors = false, stratified = true;  // ERandomMode::Stratified is the default (testbed.h)

	// sdf primitive (m_sdf.training)
	DevBuf<float> sdf_positions, sdf_distances;
	uint32_t sdf_size = 0;

	// step scratch
	DevBuf<float> positions, targets, loss_values, reduce_scratch;
	DevBuf<__half> net_out;
};

static void ftb_alloc_network(ngp_field_testbed* t) {
	const size_t n = t->desc.n_params;
	t->params_fp32.ensure(n);
	t->m1.ensure(n);
	t->m2.ensure(n);
	t->params.ensure(n);
	t->params_ema.ensure(n);
	t->grads.ensure(n);
	t->param_steps.ensure(n);
	t->mlp_grads_f32.ensure(t->desc.n_mlp_params);
	NGPB_CUDA_CHECK(cudaMemsetAsync(t->m1.p, 0, n * 4, t->stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(t->m2.p, 0, n * 4, t->stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(t->param_steps.p, 0, n * 4, t->stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(t->grads.p, 0, n * 2, t->stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(t->mlp_grads_f32.p, 0, t->desc.n_mlp_params * 4, t->stream));
}

__global__ void k_cast_f32_to_f16_pair(const uint32_t n, const float* __restrict__ src, __half* __restrict__ a, __half* __restrict__ b) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const __half h = __float2half_rn(src[i]);
	a[i] = h;
	b[i] = h;
}

static void ftb_set_params_fp32(ngp_field_testbed* t, const float* host, size_t n) {
	NGPB_CHECK(t->has_network && n == t->desc.n_params, "set_params: size mismatch");
	NGPB_CUDA_CHECK(cudaMemcpyAsync(t->params_fp32.p, host, n * 4, cudaMemcpyHostToDevice, t->stream));
	k_cast_f32_to_f16_pair<<<div_round_up((uint32_t)n, 256), 256, 0, t->stream>>>((uint32_t)n, t->params_fp32.p, t->params.p, t->params_ema.p);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
}

static void ftb_reset_network_json(ngp_field_testbed* t, const Json& config_in) {
	Json config = config_in;
	config.obj.erase("snapshot");
	t->network_config = config;
	const Json& enc = config.sub("encoding");
	const Json& net = config.sub("network");
	const std::string enc_type = to_lower(enc.value("otype", std::string("HashGrid")));
	NGPB_CHECK(enc_type == "hashgrid" || enc_type == "grid", "encoding.otype must be HashGrid / Grid");
	if (enc.contains("type")) NGPB_CHECK(to_lower(enc.value("type", std::string("hash"))) == "hash", "encoding.type must be Hash");
	{
		const std::string ot = to_lower(net.value("otype", std::string("FullyFusedMLP")));
		NGPB_CHECK(ot == "fullyfusedmlp" || ot == "megakernelmlp", "network.otype must be FullyFusedMLP");
		NGPB_CHECK(to_lower(net.value("activation", std::string("ReLU"))) == "relu", "network.activation must be ReLU");
		NGPB_CHECK(to_lower(net.value("output_activation", std::string("None"))) == "none", "network.output_activation must be None");
		NGPB_CHECK((uint32_t)net.value("n_neurons", 64.0) == 64, "network.n_neurons must be 64");
	}
	const uint32_t n_pos = t->mode == NGP_MODE_IMAGE ? 2u : 3u;
	const uint32_t n_out = t->mode == NGP_MODE_IMAGE ? 3u : 1u;  // network_dims_image / network_dims_sdf
	const uint32_t F = (uint32_t)enc.value("n_features_per_level", 2.0);
	uint32_t L = (uint32_t)enc.value("n_levels", 16.0);
	if (enc.contains("n_features") && enc.value("n_features", 0.0) > 0) L = (uint32_t)enc.value("n_features", 0.0) / F;
	const uint32_t log2_T = (uint32_t)enc.value("log2_hashmap_size", 15.0);
	uint32_t base_res = (uint32_t)enc.value("base_resolution", 0.0);
	if (!base_res) base_res = 1u << (log2_T / n_pos);
	float pls = (float)enc.value("per_level_scale", 0.0);
	if (pls <= 0.0f && L > 1) {
		// src/testbed.cu:4236-4255: finest level at max(image resolution) / 2 for images, 2048 otherwise (aabb_scale = 1)
		float desired = 2048.0f;
		if (t->mode == NGP_MODE_IMAGE) {
			NGPB_CHECK(t->img_w > 0, "image mode: set the image before (re)loading the network (per_level_scale derives from its resolution)");
			desired = (float)std::max(t->img_w, t->img_h) / 2.0f;
		}
		pls = std::exp(std::log(desired * 1.0f / (float)base_res) / (float)(L - 1));
	}
	if (pls <= 0.0f) pls = 1.0f;
	ngp_grid_desc g;
	grid_desc_init_nd(&g, n_pos, L, F, log2_T, base_res, pls);
	NGPB_CHECK(F == 2 || F == 4, "HashGrid: n_features_per_level must be 2 or 4 in this build");
	field_desc_init(&t->desc, &g, n_pos, (uint32_t)net.value("n_hidden_layers", 2.0), n_out);
	t->loss_type = parse_loss_type(config.sub("loss"));
	check_field_loss(t->loss_type);
	t->opt = parse_optimizer_chain(config.sub("optimizer"));
	t->has_network = true;
	ftb_alloc_network(t);
	t->rng = Pcg32(t->seed);
	t->training_step = 0;
	t->optimizer_step = 0;
	t->lr_factor = 1.0f;
	t->loss_scalar = 0.0f;
	std::vector<float> init(t->desc.n_params);
	field_init_params_host(&t->desc, t->seed, init.data());
	ftb_set_params_fp32(t, init.data(), init.size());
}
static void ftb_reset_network(ngp_field_testbed* t, const std::string& json_text) { ftb_reset_network_json(t, JsonParser(json_text).parse()); }

// ---- snapshots of the image / SDF modes in the reference's container (Testbed::save_snapshot / load_snapshot, src/testbed.cu:5288-5485):
// msgpack({network config..., "snapshot": {Trainer::serialize (n_params, params_type, params_binary = the INFERENCE weights, trainer.h:442-452),
// version, mode, training_step, loss, aabb, ...}}), gzip-wrapped for ".ingp".  Parameter order MLP | encoding
// (network_with_input_encoding.h:121-128).  No optimizer state (include_optimizer_state is refused).
static bool ftb_has_ext(const std::string& path, const char* ext) {
	const size_t n = strlen(ext);
	return path.size() >= n && to_lower(path.substr(path.size() - n)) == ext;
}
static void ftb_save_snapshot(ngp_field_testbed* t, const std::string& path, bool include_optimizer_state, bool compress) {
	NGPB_CHECK(t->has_network, "save_snapshot: no network");
	NGPB_CHECK(!include_optimizer_state, "save_snapshot: optimizer state is not written for the image / SDF modes");
	NGPB_CHECK(ftb_has_ext(path, ".ingp") || ftb_has_ext(path, ".msgpack"), "save_snapshot: the image / SDF modes write .ingp / .msgpack");
	NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	const size_t n = t->desc.n_params;
	std::vector<uint8_t> host(n * sizeof(__half));
	NGPB_CUDA_CHECK(cudaMemcpy(host.data(), t->params_ema.p, host.size(), cudaMemcpyDeviceToHost));
	Json config = t->network_config;
	Json snap = jobj();
	snap.obj["n_params"] = jint((int64_t)n);
	snap.obj["params_type"] = jstr("__half");
	snap.obj["params_binary"] = jbin(host.data(), host.size());
	snap.obj["version"] = jint(1);
	snap.obj["mode"] = jstr(t->mode == NGP_MODE_IMAGE ? "image" : "sdf");
	snap.obj["training_step"] = jint(t->training_step);
	snap.obj["loss"] = jnum(t->loss_scalar);
	Json aabb = jobj(), mn = jarr(), mx = jarr();
	for (int k = 0; k < 3; ++k) {
		mn.arr.push_back(jnum(0.0));
		mx.arr.push_back(jnum(1.0));
	}
	aabb.obj["min"] = mn;
	aabb.obj["max"] = mx;
	snap.obj["aabb"] = aabb;
	if (t->mode == NGP_MODE_IMAGE) {
		Json res = jarr();
		res.arr.push_back(jint(t->img_w));
		res.arr.push_back(jint(t->img_h));
		snap.obj["image_resolution"] = res;   // not a reference key: lets a render-only load size its frame (the reference re-reads the image file)
	}
	config.obj["snapshot"] = snap;
	MsgPackWriter w;
	w.write(config);
	std::vector<uint8_t> bytes = ftb_has_ext(path, ".ingp") ? gzip_compress(w.out, compress ? Z_DEFAULT_COMPRESSION : Z_NO_COMPRESSION) : w.out;
	std::ofstream f(path, std::ios::binary);
	NGPB_CHECK(f.good(), "cannot open " + path);
	f.write(reinterpret_cast<const char*>(bytes.data()), (std::streamsize)bytes.size());
	NGPB_CHECK(f.good(), "snapshot write failed");
}
static void ftb_load_snapshot(ngp_field_testbed* t, const std::string& path) {
	std::ifstream f(path, std::ios::binary);
	NGPB_CHECK(f.good(), "Network snapshot '" + path + "' does not exist.");
	std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	if (ftb_has_ext(path, ".ingp")) bytes = gzip_decompress(bytes);
	MsgPackReader r(bytes.data(), bytes.size());
	const Json config = r.read();
	NGPB_CHECK(config.contains("snapshot"), "file does not contain a snapshot");
	const Json& snap = config.at("snapshot");
	NGPB_CHECK((uint32_t)snap.value("version", 0.0) >= 1, "Snapshot uses an old format and can not be loaded.");
	const std::string mode = to_lower(snap.value("mode", std::string("")));
	NGPB_CHECK(mode == (t->mode == NGP_MODE_IMAGE ? "image" : "sdf"), "snapshot mode '" + mode + "' does not match this Testbed's mode");
	if (t->mode == NGP_MODE_IMAGE && t->img_w == 0 && snap.contains("image_resolution")) {
		// render-only load without the training image: the per-level scale derives from the resolution (src/testbed.cu:4236-4255)
		t->img_w = (int32_t)snap.at("image_resolution").arr.at(0).num;
		t->img_h = (int32_t)snap.at("image_resolution").arr.at(1).num;
	}
	ftb_reset_network_json(t, config);
	const size_t n = t->desc.n_params;
	NGPB_CHECK((size_t)snap.value("n_params", 0.0) == n, "snapshot: parameter count does not match the network config");
	const Json& pb = snap.at("params_binary");
	NGPB_CHECK(pb.type == Json::Binary, "snapshot: params_binary is not binary");
	const std::string ptype = snap.value("params_type", std::string("__half"));
	std::vector<float> p32(n);
	if (ptype == "float") {
		NGPB_CHECK(pb.bin.size() == n * 4, "snapshot: params_binary has the wrong size");
		memcpy(p32.data(), pb.bin.data(), n * 4);
	} else {
		NGPB_CHECK(ptype == "__half" && pb.bin.size() == n * 2, "snapshot: params_binary has the wrong size / type");
		const __half* h = reinterpret_cast<const __half*>(pb.bin.data());
		for (size_t i = 0; i < n; ++i) p32[i] = __half2float(h[i]);
	}
	ftb_set_params_fp32(t, p32.data(), n);   // Trainer::deserialize: all three parameter buffers from the file (trainer.h:454-478)
	t->training_step = (uint32_t)snap.value("training_step", 0.0);
	t->optimizer_step = 0;
	t->loss_scalar = (float)snap.value("loss", 0.0);
}

static void ftb_train(ngp_field_testbed* t, uint32_t batch) {
	NGPB_CHECK(t->has_network, "train: no network (reload_network_from_json first)");
	NGPB_CHECK(batch >= TILE && batch % TILE == 0, "train: batch size must be a positive multiple of 128");
	const uint32_t n_out = t->desc.n_output_dims, D = t->desc.n_pos_dims;
	t->positions.ensure((size_t)batch * D);
	t->targets.ensure((size_t)batch * n_out);
	t->loss_values.ensure((size_t)batch * n_out);
	t->reduce_scratch.ensure(1024);
	const float* pos = nullptr;
	const float* tgt = nullptr;
	if (t->mode == NGP_MODE_IMAGE) {
		NGPB_CHECK(t->img_w > 0, "train: no image");
		// stratification is skipped (with a warning in the reference) for batch sizes that are not 4^k
		uint32_t log2_n = 0;
		while ((1u << log2_n) < batch) ++log2_n;
		const bool can_stratify = (1u << log2_n) == batch && (log2_n % 2) == 0;
		image_generate_training_data(t->stream, batch, t->rng.state, t->rng.inc, t->stratified && can_stratify, t->image.p, NGP_IMAGE_FLOAT, t->img_w, t->img_h,
			t->snap_to_pixel_centers, t->linear_colors, t->positions.p, t->targets.p);
		t->rng.advance((uint64_t)batch * 2);  // generate_random: rng.advance(n_elements) (random.h:63)
		pos = t->positions.p;
		tgt = t->targets.p;
	} else {
		NGPB_CHECK(t->sdf_size >= batch, "train: fewer SDF training records than the batch size (testbed_sdf.cu:1582)");
		t->positions.ensure((size_t)t->sdf_size * 3);
		t->targets.ensure(t->sdf_size);
		shuffle_f32(t->stream, t->sdf_size, 3, t->training_step, t->sdf_positions.p, t->positions.p);
		shuffle_f32(t->stream, t->sdf_size, 1, t->training_step, t->sdf_distances.p, t->targets.p);
		pos = t->positions.p;
		tgt = t->targets.p;
	}
	field_train_step(t->desc, t->stream, batch, pos, tgt, t->loss_type, t->loss_scale, nullptr, t->params.p, t->grads.p, t->mlp_grads_f32.p, t->loss_values.p, nullptr);
	const ngp_adam_cfg a = next_adam_cfg(t->opt, t->optimizer_step, t->lr_factor, t->loss_scale, t->train_network, t->train_encoding);
	optimizer_step_flat(t->desc.n_mlp_params, t->desc.n_params, t->stream, a, t->params_fp32.p, t->params.p, t->params_ema.p, t->grads.p, t->m1.p, t->m2.p,
		t->param_steps.p);
	++t->training_step;
	t->loss_scalar = reduce_sum_f32(t->stream, t->loss_values.p, batch * n_out, t->reduce_scratch.p);  // Trainer::loss (trainer.h:372)
}

extern "C" {

int ngp_grid_desc_init_nd(ngp_grid_desc* g, uint32_t n_pos_dims, uint32_t n_levels, uint32_t F, uint32_t log2_T, uint32_t base_res, float pls) {
	NGPB_TRY(grid_desc_init_nd(g, n_pos_dims, n_levels, F, log2_T, base_res, pls));
}
int ngp_field_desc_init(ngp_field_desc* d, const ngp_grid_desc* g, uint32_t n_pos_dims, uint32_t n_hidden, uint32_t n_out) {
	NGPB_TRY(field_desc_init(d, g, n_pos_dims, n_hidden, n_out));
}
int ngp_field_init_params_host(const ngp_field_desc* d, uint64_t seed, float* out) { NGPB_TRY(field_init_params_host(d, seed, out)); }
int ngp_field_inference(const ngp_field_desc* d, void* stream, uint32_t n, const float* positions, const void* params, void* out, uint32_t out_stride) {
	NGPB_TRY(require_device(); field_inference(*d, (cudaStream_t)stream, n, positions, (const __half*)params, (__half*)out, out_stride));
}
int ngp_field_train_step(const ngp_field_desc* d, void* stream, uint32_t n, const float* positions, const float* targets, uint32_t loss_type, float loss_scale,
	const void* dL_dout_ext, const void* params, void* grads, float* loss_values, void* out) {
	NGPB_TRY({
		require_device();
		float* tmp = nullptr;
		NGPB_CUDA_CHECK(cudaMallocAsync(&tmp, d->n_mlp_params * sizeof(float), (cudaStream_t)stream));
		NGPB_CUDA_CHECK(cudaMemsetAsync(tmp, 0, d->n_mlp_params * sizeof(float), (cudaStream_t)stream));
		field_train_step(*d, (cudaStream_t)stream, n, positions, targets, loss_type, loss_scale, (const __half*)dL_dout_ext, (const __half*)params, (__half*)grads,
			tmp, loss_values, (__half*)out);
		NGPB_CUDA_CHECK(cudaFreeAsync(tmp, (cudaStream_t)stream));
	});
}
int ngp_loss_evaluate(void* stream, uint32_t loss_type, uint32_t n, uint32_t stride, uint32_t dims, float loss_scale, const void* predictions,
	const float* targets, float* values, void* gradients) {
	NGPB_TRY(require_device(); loss_evaluate((cudaStream_t)stream, loss_type, n, stride, dims, loss_scale, (const __half*)predictions, targets, values,
		(__half*)gradients));
}
int ngp_optimizer_step_flat(uint32_t n_matrix, uint32_t n_params, void* stream, const ngp_adam_cfg* cfg, float* p32, void* p16, void* ema, void* grads, float* m1,
	float* m2, uint32_t* steps) {
	NGPB_TRY(require_device(); NGPB_CHECK(n_matrix <= n_params, "ngp_optimizer_step_flat: n_matrix_params > n_params");
		optimizer_step_flat(n_matrix, n_params, (cudaStream_t)stream, *cfg, p32, (__half*)p16, (__half*)ema, (__half*)grads, m1, m2, steps));
}
int ngp_image_generate_training_data(void* stream, uint32_t n, uint64_t rng_state, uint64_t rng_inc, int stratify, const void* image, uint32_t image_type,
	int32_t w, int32_t h, int snap, int linear_colors, float* positions, float* targets) {
	NGPB_TRY(require_device(); image_generate_training_data((cudaStream_t)stream, n, rng_state, rng_inc, stratify != 0, image, image_type, w, h, snap != 0,
		linear_colors != 0, positions, targets));
}
int ngp_shuffle(void* stream, uint32_t n_elements, uint32_t stride, uint32_t seed, const float* in, float* out) {
	NGPB_TRY(require_device(); shuffle_f32((cudaStream_t)stream, n_elements, stride, seed, in, out));
}

// ---- tcnn::cpp::Module as a C handle (cpp_api.h:92-125, cpp_api.cu:72-153) --------------------------------------------------
struct ngp_module {
	ngp_field_desc desc;
};
ngp_module* ngp_module_create_network_with_input_encoding(uint32_t n_input_dims, uint32_t n_output_dims, const char* encoding_json, const char* network_json) {
	try {
		const std::string et(encoding_json), nt(network_json);
		const Json enc = JsonParser(et).parse(), net = JsonParser(nt).parse();
		const std::string enc_type = to_lower(enc.value("otype", std::string("HashGrid")));
		NGPB_CHECK(enc_type == "hashgrid" || enc_type == "grid", "encoding.otype must be HashGrid / Grid");
		if (enc.contains("type")) NGPB_CHECK(to_lower(enc.value("type", std::string("hash"))) == "hash", "encoding.type must be Hash");
		const std::string ot = to_lower(net.value("otype", std::string("FullyFusedMLP")));
		NGPB_CHECK(ot == "fullyfusedmlp" || ot == "megakernelmlp", "network.otype must be FullyFusedMLP");
		NGPB_CHECK(to_lower(net.value("activation", std::string("ReLU"))) == "relu", "network.activation must be ReLU");
		NGPB_CHECK(to_lower(net.value("output_activation", std::string("None"))) == "none", "network.output_activation must be None");
		NGPB_CHECK((uint32_t)net.value("n_neurons", 128.0) == 64, "network.n_neurons must be 64");
		// GridEncodingTemplated defaults (grid.h:1713-1760 create_grid_encoding): L 16, F 2, T 2^19, base 16, scale 2
		const uint32_t F = (uint32_t)enc.value("n_features_per_level", 2.0);
		uint32_t L = (uint32_t)enc.value("n_levels", 16.0);
		if (enc.contains("n_features") && enc.value("n_features", 0.0) > 0) L = (uint32_t)enc.value("n_features", 0.0) / F;
		ngp_grid_desc g;
		grid_desc_init_nd(&g, n_input_dims, L, F, (uint32_t)enc.value("log2_hashmap_size", 19.0), (uint32_t)enc.value("base_resolution", 16.0),
			(float)enc.value("per_level_scale", 2.0));
		NGPB_CHECK(F == 2 || F == 4, "HashGrid: n_features_per_level must be 2 or 4 in this build");
		auto* m = new ngp_module();
		field_desc_init(&m->desc, &g, n_input_dims, (uint32_t)net.value("n_hidden_layers", 5.0), n_output_dims);
		return m;
	} catch (const std::exception& e) {
		set_last_error(e.what());
		return nullptr;
	}
}
void ngp_module_free(ngp_module* m) { delete m; }
uint32_t ngp_module_n_input_dims(const ngp_module* m) { return m->desc.n_pos_dims; }
uint32_t ngp_module_n_output_dims(const ngp_module* m) { return MLP_OUT; }  // the padded width, as Module::n_output_dims
size_t ngp_module_n_params(const ngp_module* m) { return m->desc.n_params; }
int ngp_module_get_desc(const ngp_module* m, ngp_field_desc* out) { NGPB_TRY(*out = m->desc); }
int ngp_module_initialize_params(const ngp_module* m, size_t seed, float* params_fp32, float scale) {
	NGPB_TRY(field_init_params_rng(&m->desc, Pcg32((uint64_t)seed), scale, params_fp32));
}
int ngp_module_inference(const ngp_module* m, void* stream, uint32_t n, const float* input, void* output, const void* params) {
	NGPB_TRY(require_device(); field_inference(m->desc, (cudaStream_t)stream, n, input, (const __half*)params, (__half*)output, MLP_OUT));
}
// forward needs no context: the fused backward kernel recomputes the tile's forward from `input`
int ngp_module_forward(const ngp_module* m, void* stream, uint32_t n, const float* input, void* output, const void* params) {
	return ngp_module_inference(m, stream, n, input, output, params);
}
int ngp_module_backward(const ngp_module* m, void* stream, uint32_t n, float* dL_dinput, const void* dL_doutput, void* dL_dparams, const float* input,
	const void* output, const void* params) {
	NGPB_TRY({
		require_device();
		(void)output;
		NGPB_CHECK(dL_dinput == nullptr, "ngp_module_backward: gradients w.r.t. the input positions are not implemented");
		NGPB_CHECK(dL_dparams != nullptr, "ngp_module_backward: dL_dparams is required");
		// GradientMode::Overwrite (cpp_api.cu:118)
		NGPB_CUDA_CHECK(cudaMemsetAsync((__half*)dL_dparams + m->desc.grid_offset, 0, (size_t)m->desc.grid.n_params * 2, (cudaStream_t)stream));
		float* tmp = nullptr;
		NGPB_CUDA_CHECK(cudaMallocAsync(&tmp, m->desc.n_mlp_params * sizeof(float), (cudaStream_t)stream));
		NGPB_CUDA_CHECK(cudaMemsetAsync(tmp, 0, m->desc.n_mlp_params * sizeof(float), (cudaStream_t)stream));
		field_train_step(m->desc, (cudaStream_t)stream, n, input, nullptr, 0, 0.0f, (const __half*)dL_doutput, (const __half*)params, (__half*)dL_dparams, tmp, nullptr, nullptr);
		NGPB_CUDA_CHECK(cudaFreeAsync(tmp, (cudaStream_t)stream));
	});
}

// ---- B2: Testbed(ETestbedMode::Image / ::Sdf) ----------------------------------------------------------------------
ngp_field_testbed* ngp_field_testbed_create(uint32_t mode, int device, void* stream) {
	try {
		require_device();
		NGPB_CHECK(mode == NGP_MODE_IMAGE || mode == NGP_MODE_SDF, "ngp_field_testbed_create: mode must be NGP_MODE_IMAGE or NGP_MODE_SDF");
		NGPB_CUDA_CHECK(cudaSetDevice(device));
		auto* t = new ngp_field_testbed();
		t->mode = mode;
		t->device = device;
		t->stream = (cudaStream_t)stream;
		return t;
	} catch (const std::exception& e) {
		set_last_error(e.what());
		return nullptr;
	}
}
void ngp_field_testbed_destroy(ngp_field_testbed* t) {
	if (!t) return;
	cudaStreamSynchronize(t->stream);
	delete t;
}
int ngp_field_testbed_set_image(ngp_field_testbed* t, const float* rgba_host, int32_t w, int32_t h) {
	NGPB_TRY({
		NGPB_CHECK(t->mode == NGP_MODE_IMAGE, "set_image: not an image testbed");
		NGPB_CHECK(w >= 2 && h >= 2, "set_image: image must be at least 2x2");
		t->image.ensure((size_t)w * h * 4);
		NGPB_CUDA_CHECK(cudaMemcpyAsync(t->image.p, rgba_host, (size_t)w * h * 16, cudaMemcpyHostToDevice, t->stream));
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
		t->img_w = w;
		t->img_h = h;
	});
}
int ngp_field_testbed_set_sdf_training_data(ngp_field_testbed* t, const float* positions_host, const float* distances_host, uint32_t n) {
	NGPB_TRY({
		NGPB_CHECK(t->mode == NGP_MODE_SDF, "set_sdf_training_data: not an SDF testbed");
		t->sdf_positions.ensure((size_t)n * 3);
		t->sdf_distances.ensure(n);
		NGPB_CUDA_CHECK(cudaMemcpyAsync(t->sdf_positions.p, positions_host, (size_t)n * 12, cudaMemcpyHostToDevice, t->stream));
		NGPB_CUDA_CHECK(cudaMemcpyAsync(t->sdf_distances.p, distances_host, (size_t)n * 4, cudaMemcpyHostToDevice, t->stream));
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
		t->sdf_size = n;
	});
}
int ngp_field_testbed_reload_network_from_json(ngp_field_testbed* t, const char* json_text) { NGPB_TRY(ftb_reset_network(t, json_text)); }
int ngp_field_testbed_save_snapshot(ngp_field_testbed* t, const char* path, int include_optimizer_state, int compress) {
	NGPB_TRY(ftb_save_snapshot(t, path, include_optimizer_state != 0, compress != 0));
}
int ngp_field_testbed_load_snapshot(ngp_field_testbed* t, const char* path) { NGPB_TRY(ftb_load_snapshot(t, path)); }
int ngp_field_testbed_set_seed(ngp_field_testbed* t, uint64_t seed) { NGPB_TRY(t->seed = seed; t->rng = Pcg32(seed)); }
int ngp_field_testbed_set_option(ngp_field_testbed* t, const char* name, double v) {
	NGPB_TRY({
		const std::string k = name;
		if (k == "image.training.snap_to_pixel_centers") t->snap_to_pixel_centers = v != 0;
		else if (k == "image.training.linear_colors") t->linear_colors = v != 0;
		else if (k == "image.random_mode_stratified") t->stratified = v != 0;
		else if (k == "train_network") t->train_network = v != 0;
		else if (k == "train_encoding") t->train_encoding = v != 0;
		else if (k == "loss_scale") t->loss_scale = (float)v;
		else NGPB_CHECK(false, "unknown option '" + k + "'");
	});
}
int ngp_field_testbed_train(ngp_field_testbed* t, uint32_t batch_size) { NGPB_TRY(ftb_train(t, batch_size)); }
float ngp_field_testbed_loss(const ngp_field_testbed* t) { return t->loss_scalar; }
uint32_t ngp_field_testbed_training_step(const ngp_field_testbed* t) { return t->training_step; }
size_t ngp_field_testbed_n_params(const ngp_field_testbed* t) { return t->has_network ? t->desc.n_params : 0; }
int ngp_field_testbed_get_desc(const ngp_field_testbed* t, ngp_field_desc* out) { NGPB_TRY(NGPB_CHECK(t->has_network, "no network"); *out = t->desc); }
void* ngp_field_testbed_params(ngp_field_testbed* t) { return t->params.p; }
void* ngp_field_testbed_params_inference(ngp_field_testbed* t) { return t->params_ema.p; }
void* ngp_field_testbed_grads(ngp_field_testbed* t) { return t->grads.p; }
int ngp_field_testbed_set_params_fp32(ngp_field_testbed* t, const float* host, size_t n) { NGPB_TRY(ftb_set_params_fp32(t, host, n)); }
int ngp_field_testbed_get_params_fp16(ngp_field_testbed* t, void* host, size_t n, int inference) {
	NGPB_TRY({
		NGPB_CHECK(t->has_network && n == t->desc.n_params, "get_params: size mismatch");
		NGPB_CUDA_CHECK(cudaMemcpyAsync(host, inference ? t->params_ema.p : t->params.p, n * 2, cudaMemcpyDeviceToHost, t->stream));
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	});
}
// network output (inference weights) at host positions [n x D] -> host [n x n_output_dims] float32
int ngp_field_testbed_evaluate(ngp_field_testbed* t, const float* positions_host, uint32_t n, float* out_host) {
	NGPB_TRY({
		NGPB_CHECK(t->has_network, "evaluate: no network");
		const uint32_t D = t->desc.n_pos_dims, n_out = t->desc.n_output_dims;
		t->positions.ensure((size_t)n * D);
		t->net_out.ensure((size_t)n * MLP_OUT);
		NGPB_CUDA_CHECK(cudaMemcpyAsync(t->positions.p, positions_host, (size_t)n * D * 4, cudaMemcpyHostToDevice, t->stream));
		field_inference(t->desc, t->stream, n, t->positions.p, t->params_ema.p, t->net_out.p, MLP_OUT);
		std::vector<__half> h((size_t)n * MLP_OUT);
		NGPB_CUDA_CHECK(cudaMemcpyAsync(h.data(), t->net_out.p, h.size() * 2, cudaMemcpyDeviceToHost, t->stream));
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
		for (uint32_t i = 0; i < n; ++i)
			for (uint32_t k = 0; k < n_out; ++k) out_host[(size_t)i * n_out + k] = __half2float(h[(size_t)i * MLP_OUT + k]);
	});
}
// ≙ render_image (src/testbed_image.cu:304-400) for the default full-frame view: the network at every pixel centre of a
// w x h frame, shade_kernel_image's colour handling; rgba_host: h x w x 4 float32.
int ngp_field_testbed_render_image(ngp_field_testbed* t, int32_t w, int32_t h, float* rgba_host) {
	NGPB_TRY({
		NGPB_CHECK(t->mode == NGP_MODE_IMAGE && t->has_network, "render_image: image testbed with a network required");
		NGPB_CHECK(w > 0 && h > 0, "render_image: bad resolution");
		const uint32_t n = (uint32_t)w * (uint32_t)h;
		t->positions.ensure((size_t)n * 2);
		t->net_out.ensure((size_t)n * MLP_OUT);
		DevBuf<float> rgba;
		rgba.ensure((size_t)n * 4);
		k_image_pixel_coords<<<div_round_up(n, 256), 256, 0, t->stream>>>(w, h, t->positions.p);
		NGPB_LAUNCHED();
		field_inference(t->desc, t->stream, n, t->positions.p, t->params_ema.p, t->net_out.p, MLP_OUT);
		k_image_shade<<<div_round_up(n, 256), 256, 0, t->stream>>>(n, t->net_out.p, MLP_OUT, t->linear_colors, rgba.p);
		NGPB_LAUNCHED();
		NGPB_CUDA_CHECK(cudaMemcpyAsync(rgba_host, rgba.p, (size_t)n * 16, cudaMemcpyDeviceToHost, t->stream));
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	});
}
int ngp_field_testbed_sync(ngp_field_testbed* t) { NGPB_TRY(NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream))); }

}  // extern "C"
