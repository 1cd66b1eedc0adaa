// nerf_net.cu — fused NeRF network kernels for sm_100a and their C-ABI entry points.
//
//   k_nerf_forward   hash-grid gather + density MLP + SH + rgb MLP in one kernel; activations never leave the SM.
//                    ≙ NerfNetwork::inference_mixed_precision_impl (nerf_network.h:105-139) = kernel_grid + transpose +
//                    kernel_mlp_fused x2 + kernel_sh + extract_density in the reference.
//   k_nerf_density   hash-grid gather + density MLP (≙ NerfNetwork::density, nerf_network.h:270-280).
//   k_grid_encode    the encoding alone (≙ kernel_grid, grid.h:48-212), sample-contiguous output.
//
// One CTA = 128 threads = one 128-sample tile = one UMMA M=128 accumulator; thread t owns sample t and TMEM lane t.
// All MLP weights stay resident in shared memory in the chunk-major operand layout for the kernel's lifetime.
#include <atomic>
#include "nerf_net.cuh"
#include "mlp_train.cuh"

#include <vector>

namespace ngpb {

template <uint32_t F, bool DENSITY_ONLY>
__global__ void __launch_bounds__(TILE, 4) k_nerf_forward(
	const __grid_constant__ NetDev net, const uint32_t n_max, const uint32_t* __restrict__ n_dev, const float* __restrict__ in,
	const uint32_t in_stride, const uint32_t dir_offset, const __half* __restrict__ params, __half* __restrict__ out, const uint32_t out_stride
) {
	extern __shared__ __align__(128) uint8_t smem[];
	// optional device-side element count (the sample generator's counter): no host round trip to size the launch
	uint32_t n = n_max;
	if (n_dev) {
		const uint32_t nd = *n_dev;
		n = nd < n_max ? nd : n_max;
	}
	if (blockIdx.x * TILE >= n) return;
	const FwdSmem L = fwd_smem_layout(net.n_hidden_density, net.n_hidden_rgb);
	const uint32_t tid = threadIdx.x;
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8);

	// ---- one-time setup: weights -> smem, TMEM allocation, barrier
	{
		const uint32_t wd_off = 0, wr_off = mlp_n_params(net.n_hidden_density) * 2u;
		for (uint32_t l = 0; l < mlp_n_layers(net.n_hidden_density); ++l) {
			stage_weights(params + net.density_off + mlp_layer_off(net.n_hidden_density, l), mlp_layer_out(net.n_hidden_density, l),
				mlp_layer_in(net.n_hidden_density, l), smem + wd_off + mlp_layer_off(net.n_hidden_density, l) * 2u, tid, TILE);
		}
		if (!DENSITY_ONLY) {
			for (uint32_t l = 0; l < mlp_n_layers(net.n_hidden_rgb); ++l) {
				stage_weights(params + net.rgb_off + mlp_layer_off(net.n_hidden_rgb, l), mlp_layer_out(net.n_hidden_rgb, l),
					mlp_layer_in(net.n_hidden_rgb, l), smem + wr_off + mlp_layer_off(net.n_hidden_rgb, l) * 2u, tid, TILE);
			}
		}
	}
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
	const uint32_t n_tiles = (n + TILE - 1) / TILE;
	for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		const uint32_t i = tile * TILE + tid;
		const bool valid = i < n;
		const float* c = in + (size_t)(valid ? i : 0) * in_stride;
		const float x = c[0], y = c[1], z = c[2];

		// ---- encoding -> A0
		{
			__half2 enc[16];
			grid_gather<F>(net, grid, x, y, z, enc);
#pragma unroll
			for (uint32_t kc = 0; kc < 4; ++kc) {
				const __half2 h[4] = {enc[kc * 4 + 0], enc[kc * 4 + 1], enc[kc * 4 + 2], enc[kc * 4 + 3]};
				store_chunk(smem + L.a0_off, tid, kc, h);
			}
		}
		if (!DENSITY_ONLY) {
			__half2 sh[8];
			sh4_encode(c[dir_offset + 0], c[dir_offset + 1], c[dir_offset + 2], sh);
			const __half2 h0[4] = {sh[0], sh[1], sh[2], sh[3]};
			const __half2 h1[4] = {sh[4], sh[5], sh[6], sh[7]};
			store_chunk(smem + L.a2_off, tid, 2, h0);
			store_chunk(smem + L.a2_off, tid, 3, h1);
		}

		// ---- density MLP
		__half2 dens[8];
		run_mlp_fwd(smem, L.a0_off, L.h_off, 0, net.n_hidden_density, tmem_base, bar, phase, tid, dens);

		if (DENSITY_ONLY) {
			if (valid) out[(size_t)i * out_stride] = __low2half(dens[0]);
		} else {
			const __half2 h0[4] = {dens[0], dens[1], dens[2], dens[3]};
			const __half2 h1[4] = {dens[4], dens[5], dens[6], dens[7]};
			store_chunk(smem + L.a2_off, tid, 0, h0);
			store_chunk(smem + L.a2_off, tid, 1, h1);
			__half2 rgb[8];
			run_mlp_fwd(smem, L.a2_off, L.h_off, mlp_n_params(net.n_hidden_density) * 2u, net.n_hidden_rgb, tmem_base, bar, phase, tid, rgb);
			if (valid) {
				// (rgb raw x3, density raw) — columns 0..3 of the reference's padded 16-wide output row
				uint2 o;
				o.x = *reinterpret_cast<const uint32_t*>(&rgb[0]);
				const __half2 t = __halves2half2(__low2half(rgb[1]), __low2half(dens[0]));
				o.y = *reinterpret_cast<const uint32_t*>(&t);
				*reinterpret_cast<uint2*>(out + (size_t)i * out_stride) = o;
			}
		}
	}

	umma::fence_before_sync();
	__syncthreads();
	if (tid < 32) umma::tmem_dealloc<64>(tmem_base);
}

template <uint32_t F>
__global__ void __launch_bounds__(256) k_grid_encode(
	const __grid_constant__ NetDev net, const uint32_t n, const float* __restrict__ pos, const uint32_t pos_stride,
	const __half* __restrict__ grid, __half* __restrict__ out
) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float* c = pos + (size_t)i * pos_stride;
	__half2 enc[16];
	grid_gather<F>(net, grid, c[0], c[1], c[2], enc);
	uint4* o = reinterpret_cast<uint4*>(out + (size_t)i * ENC_WIDTH);
#pragma unroll
	for (uint32_t k = 0; k < 4; ++k) {
		uint4 v;
		v.x = *reinterpret_cast<const uint32_t*>(&enc[k * 4 + 0]);
		v.y = *reinterpret_cast<const uint32_t*>(&enc[k * 4 + 1]);
		v.z = *reinterpret_cast<const uint32_t*>(&enc[k * 4 + 2]);
		v.w = *reinterpret_cast<const uint32_t*>(&enc[k * 4 + 3]);
		o[k] = v;
	}
}

// ----------------------------------------------------------------------------------------------------------------
// k_nerf_train — forward + backward of the whole NeRF network for one 128-sample tile per CTA iteration.
//   ≙ Trainer::training_step(external dL/dy) → NerfNetwork::forward_impl + backward_impl (nerf_network.h:145-268):
//     kernel_grid, kernel_mlp_fused (x2, writing activations to HBM), kernel_sh, extract_rgb, kernel_mlp_fused_backward (x2),
//     5 CUTLASS split-K weight-gradient GEMMs, add_density_gradient, transpose, kernel_grid_backward in the reference.
//   Here: activations live in shared memory for the tile's lifetime, every contraction (forward, data gradient, weight
//   gradient) is a tcgen05.mma with fp32 accumulation in TMEM, weight gradients accumulate in TMEM across all tiles of
//   the CTA and leave the SM once, hash-grid gradients go out as fp16x2 reductions exactly like grid.h:252-255.
// ----------------------------------------------------------------------------------------------------------------
struct TrainSmem {
	uint32_t w_bytes;
	uint32_t a0_off;               // [128 x 32] encoding
	uint32_t hd_off;               // n_hidden_density x [128 x 64]
	uint32_t a2_off;               // [128 x 32]
	uint32_t hr_off;               // n_hidden_rgb x [128 x 64]
	uint32_t g64_off;              // [128 x 64] dL/d(hidden pre-activation), reused layer by layer
	uint32_t g16_off;              // [128 x 16] dL/d(output layer output)
	uint32_t bar_off;
	uint32_t total;
};
__host__ __device__ inline TrainSmem train_smem_layout(uint32_t nhd, uint32_t nhr) {
	TrainSmem s;
	s.w_bytes = (mlp_n_params(nhd) + mlp_n_params(nhr)) * 2u;
	s.a0_off = s.w_bytes;
	s.hd_off = s.a0_off + TILE * ENC_WIDTH * 2u;
	s.a2_off = s.hd_off + nhd * TILE * MLP_WIDTH * 2u;
	s.hr_off = s.a2_off + TILE * ENC_WIDTH * 2u;
	s.g64_off = s.hr_off + nhr * TILE * MLP_WIDTH * 2u;
	s.g16_off = s.g64_off + TILE * MLP_WIDTH * 2u;
	s.bar_off = s.g16_off + TILE * MLP_OUT * 2u;
	s.total = s.bar_off + 16u;
	return s;
}

// TMEM column map of the training kernel: [0,64) working accumulator, then one weight-gradient accumulator per layer.
// Hidden layers hold dW [64(out) x K(in)] (K columns); 16-wide output layers hold dW^T [64(in) x 16(out)] (16 columns).
__host__ __device__ inline uint32_t wgrad_col_base(uint32_t nhd, uint32_t nhr, bool rgb, uint32_t l) {
	uint32_t c = 64;
	if (rgb) {
		for (uint32_t i = 0; i <= nhd; ++i) c += wgrad_cols(nhd, i);
		for (uint32_t i = 0; i < l; ++i) c += wgrad_cols(nhr, i);
	} else {
		for (uint32_t i = 0; i < l; ++i) c += wgrad_cols(nhd, i);
	}
	return c;
}
__host__ __device__ inline uint32_t train_tmem_cols(uint32_t nhd, uint32_t nhr) { return wgrad_col_base(nhd, nhr, true, nhr + 1); }

// MLP_ONLY (profiling, ngp_profile_mlp_phase): the same tile loop with the hash-grid gather replaced by a register pattern and the
// scatter dropped, i.e. the 15 tensor-core groups of a tile and their epilogues alone — the "MLP phase" SURVEY §8d asks to see separately.
template <uint32_t F, uint32_t TMEM_COLS, bool MLP_ONLY = false, uint32_t AGG = 1>
__global__ void __launch_bounds__(TRAIN_THREADS, 2) k_nerf_train(
	const __grid_constant__ NetDev net, const uint32_t n, const float* __restrict__ coords, const __half* __restrict__ params,
	const __half* __restrict__ dL_dout, __half* __restrict__ grads, float* __restrict__ mlp_grads_f32, __half* __restrict__ out
) {
	extern __shared__ __align__(128) uint8_t smem[];
	const uint32_t nhd = net.n_hidden_density, nhr = net.n_hidden_rgb;
	const TrainSmem L = train_smem_layout(nhd, nhr);
	const uint32_t tid = threadIdx.x;
	const uint32_t row = tid & (TILE - 1), half = tid >> 7;
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8);
	const uint32_t wr_off = mlp_n_params(nhd) * 2u;

	for (uint32_t l = 0; l <= nhd; ++l)
		stage_weights(params + net.density_off + mlp_layer_off(nhd, l), mlp_layer_out(nhd, l), mlp_layer_in(nhd, l), smem + mlp_layer_off(nhd, l) * 2u, tid, TRAIN_THREADS);
	for (uint32_t l = 0; l <= nhr; ++l)
		stage_weights(params + net.rgb_off + mlp_layer_off(nhr, l), mlp_layer_out(nhr, l), mlp_layer_in(nhr, l), smem + wr_off + mlp_layer_off(nhr, l) * 2u, tid, TRAIN_THREADS);
	if (tid < 32) umma::tmem_alloc<TMEM_COLS>(tmem_slot);
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
	__half* grid_grad = grads + net.grid_off;
	const uint32_t n_tiles = n / TILE;
	uint32_t iter = 0;
	for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
		const uint32_t i = tile * TILE + row;
		const float* c = coords + (size_t)i * 7;
		const float x = c[0], y = c[1], z = c[2];
		{
			__half2 enc[8];
			if constexpr (MLP_ONLY) {
#pragma unroll
				for (uint32_t j = 0; j < 8; ++j) enc[j] = __floats2half2_rn(x * (float)(j + 1), y - z * (float)j);
			} else {
				grid_gather_half<F>(net, grid, half, x, y, z, enc);
			}
			const __half2 h0[4] = {enc[0], enc[1], enc[2], enc[3]};
			const __half2 h1[4] = {enc[4], enc[5], enc[6], enc[7]};
			store_chunk(smem + L.a0_off, row, 2 * half + 0, h0);
			store_chunk(smem + L.a0_off, row, 2 * half + 1, h1);
			if (half == 1) {
				__half2 sh[8];
				sh4_encode(c[4], c[5], c[6], sh);
				const __half2 s0[4] = {sh[0], sh[1], sh[2], sh[3]};
				const __half2 s1[4] = {sh[4], sh[5], sh[6], sh[7]};
				store_chunk(smem + L.a2_off, row, 2, s0);
				store_chunk(smem + L.a2_off, row, 3, s1);
			}
		}
		// ---------------- forward
		__half2 dens[8], rgb[8];
		run_mlp_fwd_keep(smem, L.a0_off, L.hd_off, 0, nhd, tmem_base, bar, phase, tid, dens);
		{
			const __half2 h0[4] = {dens[4 * half + 0], dens[4 * half + 1], dens[4 * half + 2], dens[4 * half + 3]};
			store_chunk(smem + L.a2_off, row, half, h0);
		}
		run_mlp_fwd_keep(smem, L.a2_off, L.hr_off, wr_off, nhr, tmem_base, bar, phase, tid, rgb);
		if (out && half == 0) {
			uint2 o;
			o.x = *reinterpret_cast<const uint32_t*>(&rgb[0]);
			const __half2 t = __halves2half2(__low2half(rgb[1]), __low2half(dens[0]));
			o.y = *reinterpret_cast<const uint32_t*>(&t);
			*reinterpret_cast<uint2*>(out + (size_t)i * 4) = o;
		}

		// ---------------- backward
		const uint2 dl = __ldg(reinterpret_cast<const uint2*>(dL_dout + (size_t)i * 4));
		const __half2 dl01 = *reinterpret_cast<const __half2*>(&dl.x);  // d rgb0, d rgb1
		const __half2 dl23 = *reinterpret_cast<const __half2*>(&dl.y);  // d rgb2, d density
		{
			const __half2 zero = __float2half2_rn(0.0f);
			if (half == 0) {
				const __half2 h0[4] = {dl01, __halves2half2(__low2half(dl23), __float2half_rn(0.0f)), zero, zero};
				store_chunk(smem + L.g16_off, row, 0, h0);
			} else {
				const __half2 h1[4] = {zero, zero, zero, zero};
				store_chunk(smem + L.g16_off, row, 1, h1);
			}
		}
		float dx[16];
		run_mlp_bwd(smem, L.a2_off, L.hr_off, L.g64_off, L.g16_off, wr_off, nhr, tmem_base, wgrad_col_base(nhd, nhr, true, 0), iter > 0 ? 1u : 0u, bar,
			phase, tid, dx);
		if (half == 0) {
			// dL/d(density-net output) = first 16 columns of dL/d(rgb-net input) (held by half 0); the density gradient joins
			// column 0 in fp16 (add_density_gradient, nerf_network.h:62-74).  Columns 16..31 (SH inputs) carry no parameters.
			__half2 h[8];
#pragma unroll
			for (uint32_t j = 0; j < 8; ++j) h[j] = __floats2half2_rn(dx[2 * j], dx[2 * j + 1]);
			h[0] = __halves2half2(__hadd(__low2half(h[0]), __high2half(dl23)), __high2half(h[0]));
			const __half2 h0[4] = {h[0], h[1], h[2], h[3]};
			const __half2 h1[4] = {h[4], h[5], h[6], h[7]};
			store_chunk(smem + L.g16_off, row, 0, h0);
			store_chunk(smem + L.g16_off, row, 1, h1);
		}
		run_mlp_bwd(smem, L.a0_off, L.hd_off, L.g64_off, L.g16_off, 0, nhd, tmem_base, wgrad_col_base(nhd, nhr, false, 0), iter > 0 ? 1u : 0u, bar, phase,
			tid, dx);
		{
			// this thread's 16 columns of dL/d(encoding) are exactly the features of its own levels
			__half2 g[8];
#pragma unroll
			for (uint32_t j = 0; j < 8; ++j) g[j] = __floats2half2_rn(dx[2 * j], dx[2 * j + 1]);
			if constexpr (MLP_ONLY) {
				// keep the gradient live without touching the table
				if (__hlt(__low2half(g[0]), __float2half_rn(-60000.0f))) grid_grad[0] = __low2half(g[1]);
			} else {
				grid_scatter_half<F, AGG>(net, grid_grad, half, x, y, z, g);
			}
		}
	}

	// ---------------- flush the weight-gradient accumulators (M = 64 accumulator layout: row m sits in TMEM lane
	// (m % 16) + 32 * (m / 16), i.e. the low 16 lanes of each warp's 32-lane window); the two halves take alternate
	// 16-column groups
	umma::fence_before_sync();
	__syncthreads();
	umma::fence_after_sync();
	if (iter > 0) {
		const uint32_t warp4 = (tid >> 5) & 3u, lane = tid & 31u;
		const uint32_t wrow = warp4 * 16u + lane;  // valid for lane < 16
		const uint32_t lane_taddr = tmem_base + ((row & ~31u) << 16);
		uint32_t group = 0;
		for (uint32_t net_i = 0; net_i < 2; ++net_i) {
			const uint32_t nh = net_i ? nhr : nhd;
			const uint32_t goff = net_i ? net.rgb_off : net.density_off;
			for (uint32_t l = 0; l <= nh; ++l) {
				const uint32_t cols = wgrad_cols(nh, l);
				const uint32_t cbase = wgrad_col_base(nhd, nhr, net_i != 0, l);
				const uint32_t K = mlp_layer_in(nh, l);
				float* dst = mlp_grads_f32 + goff + mlp_layer_off(nh, l);
				for (uint32_t c0 = 0; c0 < cols; c0 += 16, ++group) {
					if ((group & 1u) != half) continue;  // warp-uniform
					uint32_t v[16];
					umma::tmem_ld16(lane_taddr + cbase + c0, v);
					umma::tmem_ld_wait();
					if (lane < 16) {
#pragma unroll
						for (uint32_t j = 0; j < 16; ++j) {
							const uint32_t col = c0 + j;
							// hidden layer: accumulator = dW[out=row][in=col]; output layer: accumulator = dW^T[in=row][out=col]
							const uint32_t idx = (l == nh) ? (col * MLP_WIDTH + wrow) : (wrow * K + col);
							atomicAdd(dst + idx, __uint_as_float(v[j]));
						}
					}
				}
			}
		}
	}
	umma::fence_before_sync();
	__syncthreads();
	if (tid < 32) umma::tmem_dealloc<TMEM_COLS>(tmem_base);
}

__global__ void k_mlp_grads_finalize(const uint32_t n, float* __restrict__ src, __half* __restrict__ dst) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	dst[i] = __float2half_rn(src[i]);
	src[i] = 0.0f;
}

// ----------------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------------
NetDev make_netdev(const ngp_nerf_desc& d) {
	NetDev n{};
	const ngp_grid_desc& g = d.grid;
	NGPB_CHECK(g.n_features_per_level == 2 || g.n_features_per_level == 4, "HashGrid: n_features_per_level must be 2 or 4 in this build");
	NGPB_CHECK(g.n_levels * g.n_features_per_level == ENC_WIDTH, "HashGrid: n_levels * n_features_per_level must be 32");
	NGPB_CHECK(g.n_levels <= MAX_DEV_LEVELS, "HashGrid: too many levels");
	NGPB_CHECK(d.n_hidden_density >= 1 && d.n_hidden_density <= MAX_HIDDEN, "density network: 1..4 hidden layers supported");
	NGPB_CHECK(d.n_hidden_rgb >= 1 && d.n_hidden_rgb <= MAX_HIDDEN, "rgb network: 1..4 hidden layers supported");
	n.n_levels = g.n_levels;
	n.n_features = g.n_features_per_level;
	for (uint32_t l = 0; l < g.n_levels; ++l) {
		LevelMeta& m = n.levels[l];
		m.offset = g.offsets[l];
		m.size = g.offsets[l + 1] - g.offsets[l];
		m.resolution = g.resolutions[l];
		m.scale = g.scales[l];
		m.dense = level_is_dense_3d(m.resolution, m.size) ? 1u : 0u;
		NGPB_CHECK(m.dense || (m.size & (m.size - 1u)) == 0u, "HashGrid: a hashed level must have a power-of-two size");
	}
	n.n_hidden_density = d.n_hidden_density;
	n.n_hidden_rgb = d.n_hidden_rgb;
	n.density_off = d.density_mlp_offset;
	n.rgb_off = d.rgb_mlp_offset;
	n.grid_off = d.grid_offset;
	n.n_mlp_params = d.n_mlp_params;
	return n;
}

int device_sm_count() {
	static int sms = 0;
	if (!sms) {
		int dev = 0;
		NGPB_CUDA_CHECK(cudaGetDevice(&dev));
		NGPB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	}
	return sms;
}

template <uint32_t F, bool DENSITY_ONLY>
static void launch_forward(const NetDev& net, cudaStream_t stream, uint32_t n, const uint32_t* n_dev, const float* in, uint32_t in_stride,
	uint32_t dir_offset, const __half* params, __half* out, uint32_t out_stride) {
	if (n == 0) return;
	const FwdSmem L = fwd_smem_layout(net.n_hidden_density, net.n_hidden_rgb);
	auto kern = k_nerf_forward<F, DENSITY_ONLY>;
	static bool attr_set = false;
	if (!attr_set) {
		NGPB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
		attr_set = true;
	}
	NGPB_CHECK(L.total <= 100 * 1024, "MLP too large for the forward kernel's shared memory budget");
	const uint32_t n_tiles = div_round_up(n, TILE);
	const uint32_t max_ctas = (uint32_t)device_sm_count() * 4u;
	const uint32_t grid = n_tiles < max_ctas ? n_tiles : max_ctas;
	kern<<<grid, TILE, L.total, stream>>>(net, n, n_dev, in, in_stride, dir_offset, params, out, out_stride);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

void nerf_inference(const ngp_nerf_desc& d, cudaStream_t stream, uint32_t n, const float* coords, const __half* params, __half* out,
	uint32_t out_stride) {
	NGPB_CHECK(out_stride >= 4 && (out_stride % 4) == 0, "ngp_nerf_inference: out_stride must be a multiple of 4 (>= 4)");
	const NetDev net = make_netdev(d);
	if (net.n_features == 2) {
		launch_forward<2, false>(net, stream, n, nullptr, coords, 7, 4, params, out, out_stride);
	} else {
		launch_forward<4, false>(net, stream, n, nullptr, coords, 7, 4, params, out, out_stride);
	}
}

// inference over min(*n_dev, n_max) samples, 4 halves per output row
void nerf_inference_counted(const ngp_nerf_desc& d, cudaStream_t stream, uint32_t n_max, const uint32_t* n_dev, const float* coords,
	const __half* params, __half* out) {
	const NetDev net = make_netdev(d);
	if (net.n_features == 2) {
		launch_forward<2, false>(net, stream, n_max, n_dev, coords, 7, 4, params, out, 4);
	} else {
		launch_forward<4, false>(net, stream, n_max, n_dev, coords, 7, 4, params, out, 4);
	}
}

void nerf_density(const ngp_nerf_desc& d, cudaStream_t stream, uint32_t n, const float* positions, uint32_t pos_stride, const __half* params,
	__half* out) {
	NGPB_CHECK(pos_stride >= 3, "ngp_nerf_density: pos_stride must be >= 3");
	const NetDev net = make_netdev(d);
	if (net.n_features == 2) {
		launch_forward<2, true>(net, stream, n, nullptr, positions, pos_stride, 0, params, out, 1);
	} else {
		launch_forward<4, true>(net, stream, n, nullptr, positions, pos_stride, 0, params, out, 1);
	}
}

void grid_encode(const ngp_grid_desc& g, cudaStream_t stream, uint32_t n, const float* positions, uint32_t pos_stride, const __half* grid,
	__half* out) {
	if (n == 0) return;
	ngp_nerf_desc d{};
	d.grid = g;
	d.n_hidden_density = 1;
	d.n_hidden_rgb = 1;
	const NetDev net = make_netdev(d);
	const uint32_t blocks = div_round_up(n, 256);
	if (net.n_features == 2) {
		k_grid_encode<2><<<blocks, 256, 0, stream>>>(net, n, positions, pos_stride, grid, out);
	} else {
		k_grid_encode<4><<<blocks, 256, 0, stream>>>(net, n, positions, pos_stride, grid, out);
	}
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}


// A/B switch of the training kernel's run aggregation (mlp_train.cuh, grid_scatter_half_impl<AGG>); on by default
static std::atomic<int> g_scatter_aggregation{1};
void set_scatter_aggregation(int mode) { g_scatter_aggregation.store(mode < 0 ? 0 : mode > 2 ? 2 : mode, std::memory_order_relaxed); }

template <uint32_t F, uint32_t TMEM_COLS, bool MLP_ONLY = false>
static void launch_train(const NetDev& net, cudaStream_t stream, uint32_t n, const float* coords, const __half* params, const __half* dL_dout,
	__half* grads, float* mlp_grads_f32, __half* out) {
	const TrainSmem L = train_smem_layout(net.n_hidden_density, net.n_hidden_rgb);
	const int agg = MLP_ONLY ? 1 : g_scatter_aggregation.load(std::memory_order_relaxed);
	auto kern = agg == 1 ? k_nerf_train<F, TMEM_COLS, MLP_ONLY, 1> : agg == 2 ? k_nerf_train<F, TMEM_COLS, MLP_ONLY, 2> : k_nerf_train<F, TMEM_COLS, MLP_ONLY, 0>;
	static bool attr_set[3] = {false, false, false};
	if (!attr_set[agg]) {
		NGPB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		attr_set[agg] = true;
	}
	NGPB_CHECK(L.total <= 227 * 1024, "MLP too deep for the training kernel's shared memory budget");
	const uint32_t n_tiles = n / TILE;
	// resident CTAs per SM: limited by shared memory and by TMEM columns
	uint32_t per_sm = (227u * 1024u) / (L.total + 1024u);
	if (per_sm > 512u / TMEM_COLS) per_sm = 512u / TMEM_COLS;
	if (per_sm > 2) per_sm = 2;
	if (per_sm < 1) per_sm = 1;
	const uint32_t max_ctas = (uint32_t)device_sm_count() * per_sm;
	const uint32_t grid = n_tiles < max_ctas ? n_tiles : max_ctas;
	kern<<<grid, TRAIN_THREADS, L.total, stream>>>(net, n, coords, params, dL_dout, grads, mlp_grads_f32, out);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

// mlp_grads_f32: device scratch of n_mlp_params floats, all zero on entry (left zeroed on exit).
void nerf_forward_backward(const ngp_nerf_desc& d, cudaStream_t stream, uint32_t n, const float* coords, const __half* params,
	const __half* dL_dout, __half* grads, float* mlp_grads_f32, __half* out) {
	NGPB_CHECK(n % TILE == 0, "ngp_nerf_forward_backward: batch size must be a multiple of 128");
	if (n == 0) return;
	const NetDev net = make_netdev(d);
	const uint32_t cols = train_tmem_cols(net.n_hidden_density, net.n_hidden_rgb);
	NGPB_CHECK(cols <= 512, "MLP too deep for the training kernel's TMEM budget");
	if (net.n_features == 2) {
		if (cols <= 256) launch_train<2, 256>(net, stream, n, coords, params, dL_dout, grads, mlp_grads_f32, out);
		else launch_train<2, 512>(net, stream, n, coords, params, dL_dout, grads, mlp_grads_f32, out);
	} else {
		if (cols <= 256) launch_train<4, 256>(net, stream, n, coords, params, dL_dout, grads, mlp_grads_f32, out);
		else launch_train<4, 512>(net, stream, n, coords, params, dL_dout, grads, mlp_grads_f32, out);
	}
	k_mlp_grads_finalize<<<div_round_up(d.n_mlp_params, 256), 256, 0, stream>>>(d.n_mlp_params, mlp_grads_f32, grads);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

// profiling entry: the MLP phase of the fused training kernel alone (base network: 1 + 2 hidden layers, F = 2)
void profile_mlp_phase(const ngp_nerf_desc& d, cudaStream_t stream, uint32_t n, const float* coords, const __half* params, const __half* dL_dout, __half* grads,
	float* mlp_grads_f32) {
	NGPB_CHECK(n % TILE == 0 && n > 0, "ngp_profile_mlp_phase: batch size must be a positive multiple of 128");
	const NetDev net = make_netdev(d);
	NGPB_CHECK(net.n_features == 2 && train_tmem_cols(net.n_hidden_density, net.n_hidden_rgb) <= 256, "ngp_profile_mlp_phase: base network (F = 2, 1 + 2 hidden layers) only");
	launch_train<2, 256, true>(net, stream, n, coords, params, dL_dout, grads, mlp_grads_f32, nullptr);
}

}  // namespace ngpb
