// mlp_train.cuh — device building blocks shared by the training kernels (k_nerf_train in nerf_net.cu, k_field_train in
// field.cu): the two-threads-per-sample epilogues, forward-with-kept-activations, the backward chain with tcgen05 data- and
// weight-gradient MMAs, and the half-of-the-levels hash-grid gather / scatter.
#pragma once

#include "nerf_net.cuh"

namespace ngpb {

// weight-gradient accumulator width (TMEM columns) of layer l: hidden layers hold dW [64(out) x K(in)] (K columns);
// 16-wide output layers hold dW^T [64(in) x 16(out)] (16 columns).
__host__ __device__ inline uint32_t wgrad_cols(uint32_t n_hidden, uint32_t l) { return l == n_hidden ? MLP_OUT : mlp_layer_in(n_hidden, l); }

// ---- two threads per sample -------------------------------------------------------------------------------------------
// The training CTA has 256 threads for its 128-sample tile: thread t and thread t+128 own the same sample (row t & 127,
// TMEM lane t & 127 — warps w and w+4 address the same 32-lane TMEM window).  "half" h = t >> 7 selects which half of the
// per-sample work a thread does: levels [h*L/2, (h+1)*L/2) of the gather / scatter, columns [32h, 32h+32) of every
// 64-wide epilogue.  Same shared memory and TMEM per CTA as one thread per sample, twice the warps in flight per SM
// (ncu of the one-thread version: 12 % warps active, latency bound; profiles/r1_kernels.md).
constexpr uint32_t TRAIN_THREADS = 2 * TILE;

// epilogue of a hidden layer for this thread's 32 columns: TMEM row -> ReLU -> fp16 -> operand buffer
__device__ __forceinline__ void tmem_row_to_smem32_relu(uint32_t taddr_row, uint32_t half, uint8_t* buf, uint32_t row) {
	uint32_t v0[16], v1[16];
	umma::tmem_ld16(taddr_row + half * 32u + 0, v0);
	umma::tmem_ld16(taddr_row + half * 32u + 16, v1);
	umma::tmem_ld_wait();
	auto emit = [&](const uint32_t(&v)[16], uint32_t kc0) {
#pragma unroll
		for (uint32_t c = 0; c < 2; ++c) {
			__half2 h[4];
#pragma unroll
			for (uint32_t j = 0; j < 4; ++j) h[j] = relu2(__floats2half2_rn(__uint_as_float(v[c * 8 + 2 * j]), __uint_as_float(v[c * 8 + 2 * j + 1])));
			store_chunk(buf, row, kc0 + c, h);
		}
	};
	emit(v0, half * 4u + 0);
	emit(v1, half * 4u + 2);
}

// forward of one MLP keeping every hidden activation: layer l writes hidden buffer l (buffers `hid_stride` bytes apart;
// 0 = inference, every layer overwrites the same buffer once its MMA has read it).  Returns the 16 outputs of the row.
__device__ __forceinline__ void run_mlp_fwd_keep(
	uint8_t* smem, uint32_t in_off, uint32_t hid_off, uint32_t w_off, uint32_t n_hidden, uint32_t tmem_base, uint64_t* bar, uint32_t& phase,
	uint32_t tid, __half2 (&out)[8], uint32_t hid_stride = TILE * MLP_WIDTH * 2u
) {
	const uint32_t smem_base = umma::smem_u32(smem);
	const uint32_t row = tid & (TILE - 1), half = tid >> 7;
	const uint32_t lane_taddr = tmem_base + ((row & ~31u) << 16);
	for (uint32_t l = 0; l <= n_hidden; ++l) {
		const uint32_t K = mlp_layer_in(n_hidden, l), N = mlp_layer_out(n_hidden, l);
		const uint32_t a_off = (l == 0) ? in_off : hid_off + (l - 1) * hid_stride;
		umma::fence_smem_to_async();
		umma::fence_before_sync();
		__syncthreads();
		if (tid == 0) {
			umma::fence_after_sync();
			issue_layer_fwd(smem_base + a_off, K, smem_base + w_off + mlp_layer_off(n_hidden, l) * 2u, N, tmem_base, bar);
		}
		umma::mbar_wait(bar, phase);
		phase ^= 1u;
		umma::fence_after_sync();
		if (l < n_hidden) {
			tmem_row_to_smem32_relu(lane_taddr, half, smem + hid_off + l * hid_stride, row);
		} else {
			tmem_row_to_regs16(lane_taddr, out);
		}
	}
}

// backward of one MLP.  On entry g16 holds dL/d(output) [128 x 16].  For every layer, weight gradient and data
// gradient are issued back to back and waited for together.  Returns this thread's 16 columns [16*half, 16*half+16) of
// dL/d(mlp input) (the MLP input is 32 wide).
__device__ __forceinline__ void run_mlp_bwd(
	uint8_t* smem, uint32_t in_off, uint32_t hid_off, uint32_t g64_off, uint32_t g16_off, uint32_t w_off, uint32_t n_hidden, uint32_t tmem_base,
	uint32_t wg_col0, uint32_t wg_accumulate, uint64_t* bar, uint32_t& phase, uint32_t tid, float (&dx)[16]
) {
	const uint32_t smem_base = umma::smem_u32(smem);
	const uint32_t row = tid & (TILE - 1), half = tid >> 7;
	const uint32_t lane_taddr = tmem_base + ((row & ~31u) << 16);
	for (int32_t l = (int32_t)n_hidden; l >= 0; --l) {
		const uint32_t K = mlp_layer_in(n_hidden, l), N = mlp_layer_out(n_hidden, l);
		const uint32_t x_off = (l == 0) ? in_off : hid_off + (l - 1) * TILE * MLP_WIDTH * 2u;
		const uint32_t dy_off = ((uint32_t)l == n_hidden) ? g16_off : g64_off;
		uint32_t wg_col = wg_col0;
		for (int32_t i = 0; i < l; ++i) wg_col += wgrad_cols(n_hidden, (uint32_t)i);
		umma::fence_smem_to_async();
		umma::fence_before_sync();
		__syncthreads();
		if (tid == 0) {
			umma::fence_after_sync();
			if ((uint32_t)l == n_hidden) {
				issue_wgrad(smem_base + x_off, MLP_WIDTH, smem_base + dy_off, MLP_OUT, tmem_base + wg_col, wg_accumulate);  // dW^T [in x out]
			} else {
				issue_wgrad(smem_base + dy_off, MLP_WIDTH, smem_base + x_off, K, tmem_base + wg_col, wg_accumulate);  // dW [out x in]
			}
			issue_layer_dgrad(smem_base + dy_off, N, smem_base + w_off + mlp_layer_off(n_hidden, l) * 2u, K, tmem_base, bar);
		}
		umma::mbar_wait(bar, phase);
		phase ^= 1u;
		umma::fence_after_sync();
		if (l > 0) {
			// dL/d(hidden l-1 pre-activation) = dX * (hidden_{l-1} > 0)  -> g64, this thread's 32 columns
			uint32_t v0[16], v1[16];
			umma::tmem_ld16(lane_taddr + half * 32u + 0, v0);
			umma::tmem_ld16(lane_taddr + half * 32u + 16, v1);
			umma::tmem_ld_wait();
			const uint8_t* act = smem + x_off;
			uint8_t* g = smem + g64_off;
			auto emit = [&](const uint32_t(&v)[16], uint32_t kc0) {
#pragma unroll
				for (uint32_t c = 0; c < 2; ++c) {
					const uint4 a = *reinterpret_cast<const uint4*>(act + (kc0 + c) * (TILE * 16u) + row * 16u);
					const __half2 ah[4] = {*reinterpret_cast<const __half2*>(&a.x), *reinterpret_cast<const __half2*>(&a.y),
						*reinterpret_cast<const __half2*>(&a.z), *reinterpret_cast<const __half2*>(&a.w)};
					__half2 h[4];
#pragma unroll
					for (uint32_t j = 0; j < 4; ++j) {
						const __half2 t = __floats2half2_rn(__uint_as_float(v[c * 8 + 2 * j]), __uint_as_float(v[c * 8 + 2 * j + 1]));
						const __half2 m = __hgt2(ah[j], __float2half2_rn(0.0f));  // 1.0 where act > 0
						h[j] = __hmul2(t, m);
					}
					store_chunk(g, row, kc0 + c, h);
				}
			};
			emit(v0, half * 4u + 0);
			emit(v1, half * 4u + 2);
		} else {
			uint32_t v0[16];
			umma::tmem_ld16(lane_taddr + half * 16u, v0);
			umma::tmem_ld_wait();
#pragma unroll
			for (uint32_t j = 0; j < 16; ++j) dx[j] = __uint_as_float(v0[j]);
		}
	}
}

// 2-D variant of level_corner_indices (common.cuh): corner c = x + 2y.  grid_index<2> (common_device.h:847-884): dense
// iff res <= 0xFFFF and size >= res^2; hash = x ^ y * 2654435761.
__device__ __forceinline__ void level_corner_indices_2d(const LevelMeta& lv, uint32_t gx, uint32_t gy, uint32_t (&idx)[4]) {
	if (lv.dense) {
		const uint32_t res = lv.resolution;
		if (gx < res && gy < res) {
			const uint32_t b = gx + gy * res;
#pragma unroll
			for (uint32_t c = 0; c < 4; ++c) {
				uint32_t i = b + (c & 1u) + ((c & 2u) ? res : 0u);
				if (i >= lv.size) i -= lv.size;
				idx[c] = i;
			}
		} else {
#pragma unroll
			for (uint32_t c = 0; c < 4; ++c) idx[c] = ((gx + (c & 1u)) + (gy + (c >> 1)) * res) % lv.size;
		}
	} else {
		const uint32_t mask = lv.size - 1u;
		const uint32_t hy0 = gy * 2654435761u;
		const uint32_t hy[2] = {hy0, hy0 + 2654435761u};
#pragma unroll
		for (uint32_t c = 0; c < 4; ++c) idx[c] = ((gx + (c & 1u)) ^ hy[c >> 1]) & mask;
	}
}

// per-level cell + interpolation weights of a D-dimensional position (kernel_grid, grid.h:95-110: pos_fract)
template <uint32_t D>
struct LevelCell {
	uint32_t g[D];
	float w0[D], w1[D];
	__device__ __forceinline__ LevelCell(const LevelMeta& lv, const float (&x)[D]) {
#pragma unroll
		for (uint32_t d = 0; d < D; ++d) {
			const float p = fmaf(lv.scale, x[d], 0.5f);
			const float f = floorf(p);
			g[d] = (uint32_t)(int)f;
			w1[d] = p - f;
			w0[d] = 1.0f - w1[d];
		}
	}
	__device__ __forceinline__ float weight(uint32_t c) const {
		float w = (c & 1u) ? w1[0] : w0[0];
#pragma unroll
		for (uint32_t d = 1; d < D; ++d) w *= ((c >> d) & 1u) ? w1[d] : w0[d];
		return w;
	}
	__device__ __forceinline__ void corners(const LevelMeta& lv, uint32_t (&idx)[1u << D]) const {
		if constexpr (D == 3) level_corner_indices(lv, g[0], g[1], g[2], idx);
		else level_corner_indices_2d(lv, g[0], g[1], idx);
	}
};

// gather of HALF of the levels of one sample: 16 encoded features (8 half2) = chunks {2*half, 2*half+1} of the A0 row.
// NET is any struct with a `levels` table (NetDev, FieldDev).
// HALF is a template argument so that the level table is read with constant indices: with a run-time half the 5 words of every level's
// metadata come through LDC with a register offset, 13 % of k_nerf_train's stall samples (ncu r2f); the dispatcher below is warp uniform.
template <uint32_t F, uint32_t D, uint32_t half, typename NET>
__device__ __forceinline__ void grid_gather_half_impl(const NET& net, const __half* __restrict__ grid, const float (&x)[D], __half2 (&enc)[8]) {
	constexpr uint32_t H2_PER_LEVEL = F / 2;
	constexpr uint32_t LEVELS_PER_HALF = (ENC_WIDTH / F) / 2;
	constexpr uint32_t NC = 1u << D;
#pragma unroll
	for (uint32_t ll = 0; ll < LEVELS_PER_HALF; ++ll) {
		const LevelMeta lv = net.levels[half * LEVELS_PER_HALF + ll];
		const LevelCell<D> cell(lv, x);
		const __half* lgrid = grid + (size_t)lv.offset * F;
		uint32_t idx[NC];
		cell.corners(lv, idx);
		if constexpr (F == 2) {
			__half2 v[NC];
			// The SM's load/store unit takes scattered accesses at 1.01 cycles per lane for gathers and 1.50 for reductions (measured on the
			// B200, tools/microbench/lsu_scatter.cu): 128 + 128 of them per sample would be 0.29 ms per 2^18 samples.  The two x-neighbours of a corner
			// pair are adjacent entries whenever the first index is even (always on dense levels up to the wrap, and on hashed levels
			// because the x prime is 1): one 8-byte access then serves both.
#pragma unroll
			for (uint32_t c = 0; c < NC; c += 2) {
				const __half2* p0 = reinterpret_cast<const __half2*>(lgrid) + idx[c];
				if (((idx[c] & 1u) == 0u) && idx[c + 1] == idx[c] + 1u) {
					const uint2 t = __ldg(reinterpret_cast<const uint2*>(p0));
					v[c] = *reinterpret_cast<const __half2*>(&t.x);
					v[c + 1] = *reinterpret_cast<const __half2*>(&t.y);
				} else {
					v[c] = __ldg(p0);
					v[c + 1] = __ldg(reinterpret_cast<const __half2*>(lgrid) + idx[c + 1]);
				}
			}
			__half2 acc = __float2half2_rn(0.0f);
#pragma unroll
			for (uint32_t c = 0; c < NC; ++c) acc = __hfma2(__float2half2_rn(cell.weight(c)), v[c], acc);
			enc[ll] = acc;
		} else {
			uint2 v[NC];
#pragma unroll
			for (uint32_t c = 0; c < NC; c += 2) {   // F = 4: 8-byte entries, adjacent aligned x-neighbours as one 16-byte load
				const uint2* p0 = reinterpret_cast<const uint2*>(lgrid) + idx[c];
				if (((idx[c] & 1u) == 0u) && idx[c + 1] == idx[c] + 1u) {
					const uint4 t = __ldg(reinterpret_cast<const uint4*>(p0));
					v[c] = make_uint2(t.x, t.y);
					v[c + 1] = make_uint2(t.z, t.w);
				} else {
					v[c] = __ldg(p0);
					v[c + 1] = __ldg(reinterpret_cast<const uint2*>(lgrid) + idx[c + 1]);
				}
			}
			__half2 a0 = __float2half2_rn(0.0f), a1 = a0;
#pragma unroll
			for (uint32_t c = 0; c < NC; ++c) {
				const __half2 wh = __float2half2_rn(cell.weight(c));
				a0 = __hfma2(wh, *reinterpret_cast<const __half2*>(&v[c].x), a0);
				a1 = __hfma2(wh, *reinterpret_cast<const __half2*>(&v[c].y), a1);
			}
			enc[ll * H2_PER_LEVEL + 0] = a0;
			enc[ll * H2_PER_LEVEL + 1] = a1;
		}
	}
}

template <uint32_t F, uint32_t D, typename NET>
__device__ __forceinline__ void grid_gather_half_nd(const NET& net, const __half* __restrict__ grid, uint32_t half, const float (&x)[D], __half2 (&enc)[8]) {
	if (half == 0) grid_gather_half_impl<F, D, 0>(net, grid, x, enc);
	else grid_gather_half_impl<F, D, 1>(net, grid, x, enc);
}

// scatter dL/d(encoding) of HALF of the levels of one sample into the fp16 gradient table
// (≙ kernel_grid_backward, grid.h:214-320: fp16 weight x fp16 gradient, red.global.add.f16x2 per corner).
// AGG (NeRF training only): the compacted samples of one ray sit in consecutive rows, i.e. consecutive lanes, and at the coarse levels many of
// them fall into the same cell.  Each run of equal cells in the warp is summed with a segmented shuffle scan and its first lane issues the
// reductions: a reduction costs 1.5 load/store-unit cycles per lane (profiles/r2/r2l_lsu_scatter.jsonl), a shuffle + HADD2 a fraction of that.
template <uint32_t F, uint32_t D, uint32_t half, bool AGG, typename NET>
__device__ __forceinline__ void grid_scatter_half_impl(const NET& net, __half* __restrict__ grid_grad, const float (&x)[D], const __half2 (&g)[8]) {
	static_assert(F == 2 || F == 4, "2 or 4 features per level");
	constexpr uint32_t LEVELS_PER_HALF = (ENC_WIDTH / F) / 2;
	constexpr uint32_t NC = 1u << D;
	constexpr uint32_t H = F / 2;  // half2 words per table entry
	// fully unrolled: with a partial unroll g[ll] is indexed dynamically and the gradient array lands in local memory — the LDL
	// behind every corner's HMUL2 was 9 % of the training kernel's stall samples (ncu r1c)
#pragma unroll
	for (uint32_t ll = 0; ll < LEVELS_PER_HALF; ++ll) {
		const LevelMeta lv = net.levels[half * LEVELS_PER_HALF + ll];
		const LevelCell<D> cell(lv, x);
		__half2* lgrad = reinterpret_cast<__half2*>(grid_grad + (size_t)lv.offset * F);
		uint32_t cidx[NC];
		cell.corners(lv, cidx);
		__half2 v[NC][H];
#pragma unroll
		for (uint32_t c = 0; c < NC; ++c) {
			const __half2 wh = __float2half2_rn(cell.weight(c));
#pragma unroll
			for (uint32_t h = 0; h < H; ++h) v[c][h] = __hmul2(wh, g[ll * H + h]);
		}
		bool issue = true;
		if constexpr (AGG) {
			static_assert(D == 3, "run aggregation is written for 3-D cells");
			const uint32_t lane = threadIdx.x & 31u;
			const unsigned long long key = (unsigned long long)cell.g[0] | ((unsigned long long)cell.g[1] << 21) | ((unsigned long long)cell.g[2] << 42);
			const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
			const uint32_t lo = (uint32_t)__ffs((int)peers) - 1u, hi = 31u - (uint32_t)__clz((int)peers);
			// only contiguous runs are merged (lanes lo..hi all in `peers`); anything else keeps its own reductions
			const bool run = hi > lo && peers == ((0xFFFFFFFFu >> (31u - hi)) & (0xFFFFFFFFu << lo));
			if (__any_sync(0xFFFFFFFFu, run)) {
				// segmented suffix sum: after step d, lane l holds the sum over lanes l .. min(l + 2d - 1, hi)
#pragma unroll
				for (uint32_t d = 1; d < 32u; d <<= 1) {
					const bool take = run && lane + d <= hi;
#pragma unroll
					for (uint32_t c = 0; c < NC; ++c) {
#pragma unroll
						for (uint32_t h = 0; h < H; ++h) {
							const __half2 o = __shfl_down_sync(0xFFFFFFFFu, v[c][h], d);
							if (take) v[c][h] = __hadd2(v[c][h], o);
						}
					}
				}
				issue = !run || lane == lo;
			}
		}
		if (!issue) continue;
		// fire-and-forget reductions (NOT atomicAdd(__half2*): on a generic pointer that compiles to ATOM + predicate + retry branch, a full L2
		// round trip per corner — 59 % of all stall samples in profiles/r1_kernels.md); pairs of x-neighbours as one vector reduction where
		// they are adjacent and aligned (see grid_gather_half_impl)
		if constexpr (D == 3 || D == 2) {
#pragma unroll
			for (uint32_t c = 0; c < NC; c += 2) {
				const bool paired = ((cidx[c] & 1u) == 0u) && cidx[c + 1] == cidx[c] + 1u;
				if constexpr (F == 2) {
					const uint32_t v0 = *reinterpret_cast<const uint32_t*>(&v[c][0]), v1 = *reinterpret_cast<const uint32_t*>(&v[c + 1][0]);
					if (paired) {
						asm volatile("red.relaxed.gpu.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(lgrad + cidx[c]), "r"(v0), "r"(v1) : "memory");
					} else {
						asm volatile("red.relaxed.gpu.global.add.noftz.f16x2 [%0], %1;" ::"l"(lgrad + cidx[c]), "r"(v0) : "memory");
						asm volatile("red.relaxed.gpu.global.add.noftz.f16x2 [%0], %1;" ::"l"(lgrad + cidx[c + 1]), "r"(v1) : "memory");
					}
				} else {
					const uint32_t a0 = *reinterpret_cast<const uint32_t*>(&v[c][0]), a1 = *reinterpret_cast<const uint32_t*>(&v[c][1]);
					const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&v[c + 1][0]), b1 = *reinterpret_cast<const uint32_t*>(&v[c + 1][1]);
					if (paired) {
						asm volatile("red.relaxed.gpu.global.add.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(lgrad + (size_t)cidx[c] * 2), "r"(a0), "r"(a1), "r"(b0), "r"(b1) : "memory");
					} else {
						asm volatile("red.relaxed.gpu.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(lgrad + (size_t)cidx[c] * 2), "r"(a0), "r"(a1) : "memory");
						asm volatile("red.relaxed.gpu.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(lgrad + (size_t)cidx[c + 1] * 2), "r"(b0), "r"(b1) : "memory");
					}
				}
			}
		}
	}
}

template <uint32_t F, uint32_t D, uint32_t AGG = 0, typename NET>
__device__ __forceinline__ void grid_scatter_half_nd(const NET& net, __half* __restrict__ grid_grad, uint32_t half, const float (&x)[D], const __half2 (&g)[8]) {
	// `half` is warp-uniform at every call site (threads 0..127 / 128..255 of the CTA), as the warp collectives of AGG require.
	// AGG: 0 = off, 1 = the coarse half of the levels, 2 = every level
	if (half == 0) grid_scatter_half_impl<F, D, 0, (AGG >= 1)>(net, grid_grad, x, g);
	else grid_scatter_half_impl<F, D, 1, (AGG >= 2)>(net, grid_grad, x, g);
}

template <uint32_t F>
__device__ __forceinline__ void grid_gather_half(const NetDev& net, const __half* __restrict__ grid, uint32_t half, float x, float y, float z, __half2 (&enc)[8]) {
	const float p[3] = {x, y, z};
	grid_gather_half_nd<F, 3>(net, grid, half, p, enc);
}
template <uint32_t F, uint32_t AGG = 0>
__device__ __forceinline__ void grid_scatter_half(const NetDev& net, __half* __restrict__ grid_grad, uint32_t half, float x, float y, float z, const __half2 (&g)[8]) {
	const float p[3] = {x, y, z};
	grid_scatter_half_nd<F, 3, AGG>(net, grid_grad, half, p, g);
}

}  // namespace ngpb
