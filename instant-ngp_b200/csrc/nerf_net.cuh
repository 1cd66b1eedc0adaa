// nerf_net.cuh — device building blocks of the fused NeRF network kernels:
//   multiresolution hash-grid gather (restates tiny-cuda-nn/encodings/grid.h:48-212 kernel_grid),
//   degree-4 spherical harmonics (tiny-cuda-nn/common_device.h:475-503 sh_enc, spherical_harmonics.h:66-72),
//   tcgen05 layer issue + TMEM epilogues for the 64-wide MLPs (semantics of fully_fused_mlp.cu:499-557).
#pragma once

#include "common.cuh"
#include "umma.cuh"

namespace ngpb {

constexpr uint32_t TILE = 128;          // samples per CTA tile = UMMA M = TMEM lanes
constexpr uint32_t ENC_WIDTH = 32;      // L * F, the density-MLP input width of every supported config
constexpr uint32_t MLP_WIDTH = 64;
constexpr uint32_t MLP_OUT = 16;        // padded output width (fully_fused_mlp.cu: m_padded_output_width)
constexpr uint32_t MAX_DEV_LEVELS = 16;
constexpr uint32_t MAX_HIDDEN = 4;

// Everything a kernel needs to know about the network; passed by value (__grid_constant__).
struct NetDev {
	LevelMeta levels[MAX_DEV_LEVELS];
	uint32_t n_levels;
	uint32_t n_features;  // per level: 2 or 4
	uint32_t n_hidden_density;
	uint32_t n_hidden_rgb;
	uint32_t density_off;  // element offsets into the flat param buffer
	uint32_t rgb_off;
	uint32_t grid_off;
	uint32_t n_mlp_params;
};

// layer table helpers ------------------------------------------------------------------------------------------------
// density MLP: [64x32] , (n_hidden_density-1) x [64x64] , [16x64]
// rgb MLP    : [64x32] , (n_hidden_rgb-1)     x [64x64] , [16x64]
__host__ __device__ inline uint32_t mlp_n_layers(uint32_t n_hidden) { return n_hidden + 1; }
__host__ __device__ inline uint32_t mlp_layer_in(uint32_t n_hidden, uint32_t l) { return l == 0 ? ENC_WIDTH : MLP_WIDTH; }
__host__ __device__ inline uint32_t mlp_layer_out(uint32_t n_hidden, uint32_t l) { return l == n_hidden ? MLP_OUT : MLP_WIDTH; }
__host__ __device__ inline uint32_t mlp_layer_off(uint32_t n_hidden, uint32_t l) {
	// offset (elements) of layer l inside its MLP
	if (l == 0) return 0;
	return MLP_WIDTH * ENC_WIDTH + (l - 1) * MLP_WIDTH * MLP_WIDTH;
}
__host__ __device__ inline uint32_t mlp_n_params(uint32_t n_hidden) {
	return MLP_WIDTH * ENC_WIDTH + (n_hidden - 1) * MLP_WIDTH * MLP_WIDTH + MLP_OUT * MLP_WIDTH;
}

// ------------------------------------------------------------------------------------------------------------------
// hash-grid gather for ONE sample: 32 encoded features as 16 half2 (sample-contiguous order level*F + f).
// Arithmetic follows kernel_grid exactly: pos = fma(scale, x, 0.5); floor; 8 corners in index order 0..7 with
// weight = ((wx)*wy)*wz in fp32, cast to fp16, result = fma(weight_h, value_h, result) in fp16 (grid.h:144-163).
// ------------------------------------------------------------------------------------------------------------------
template <uint32_t F>
__device__ __forceinline__ void grid_gather(const NetDev& net, const __half* __restrict__ grid, float x, float y, float z, __half2 (&enc)[16]) {
	static_assert(F == 2 || F == 4, "features per level");
	constexpr uint32_t H2_PER_LEVEL = F / 2;
	const uint32_t n_levels = ENC_WIDTH / F;
#pragma unroll 4
	for (uint32_t l = 0; l < n_levels; ++l) {
		const LevelMeta lv = net.levels[l];
		const float px = fmaf(lv.scale, x, 0.5f), py = fmaf(lv.scale, y, 0.5f), pz = fmaf(lv.scale, z, 0.5f);
		const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
		const uint32_t gx = (uint32_t)(int)fx, gy = (uint32_t)(int)fy, gz = (uint32_t)(int)fz;
		const float wx1 = px - fx, wy1 = py - fy, wz1 = pz - fz;
		const float wx0 = 1.0f - wx1, wy0 = 1.0f - wy1, wz0 = 1.0f - wz1;
		const __half* lgrid = grid + (size_t)lv.offset * F;

		__half2 acc[H2_PER_LEVEL];
#pragma unroll
		for (uint32_t h = 0; h < H2_PER_LEVEL; ++h) acc[h] = __float2half2_rn(0.0f);

		// issue all 8 corner loads before using them
		uint32_t idx[8];
		level_corner_indices(lv, gx, gy, gz, idx);
		if constexpr (F == 2) {
			__half2 v[8];
			// x-neighbours that are adjacent, aligned entries share one 8-byte load (the LSU takes scattered accesses at about one
			// lane per clock; see grid_gather_half_nd in mlp_train.cuh)
#pragma unroll
			for (uint32_t c = 0; c < 8; c += 2) {
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
#pragma unroll
			for (uint32_t c = 0; c < 8; ++c) {
				const float w = ((c & 1u) ? wx1 : wx0) * ((c & 2u) ? wy1 : wy0) * ((c & 4u) ? wz1 : wz0);
				acc[0] = __hfma2(__float2half2_rn(w), v[c], acc[0]);
			}
		} else {
			uint2 v[8];
#pragma unroll
			for (uint32_t c = 0; c < 8; c += 2) {   // adjacent aligned x-neighbours as one 16-byte load
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
#pragma unroll
			for (uint32_t c = 0; c < 8; ++c) {
				const float w = ((c & 1u) ? wx1 : wx0) * ((c & 2u) ? wy1 : wy0) * ((c & 4u) ? wz1 : wz0);
				const __half2 wh = __float2half2_rn(w);
				acc[0] = __hfma2(wh, *reinterpret_cast<const __half2*>(&v[c].x), acc[0]);
				acc[1] = __hfma2(wh, *reinterpret_cast<const __half2*>(&v[c].y), acc[1]);
			}
		}
#pragma unroll
		for (uint32_t h = 0; h < H2_PER_LEVEL; ++h) enc[l * H2_PER_LEVEL + h] = acc[h];
	}
}

// ------------------------------------------------------------------------------------------------------------------
// SH degree 4 (16 coefficients) of d = 2*dir - 1, rounded to fp16 (sh_enc, common_device.h:475-503).
// ------------------------------------------------------------------------------------------------------------------
// Written with explicit single-rounding multiplies / adds: the same values whether or not the including translation unit
// allows FMA contraction (render.cu is built with -fmad=false, nerf_net.cu is not), and the same as the oracle.
__device__ __forceinline__ void sh4_encode(float dx, float dy, float dz, __half2 (&o)[8]) {
#define M(a, b) __fmul_rn((a), (b))
#define A(a, b) __fadd_rn((a), (b))
#define S(a, b) __fsub_rn((a), (b))
	const float x = S(M(dx, 2.f), 1.f), y = S(M(dy, 2.f), 1.f), z = S(M(dz, 2.f), 1.f);
	const float xy = M(x, y), xz = M(x, z), yz = M(y, z), x2 = M(x, x), y2 = M(y, y), z2 = M(z, z);
	float s[16];
	s[0] = 0.28209479177387814f;
	s[1] = M(-0.48860251190291987f, y);
	s[2] = M(0.48860251190291987f, z);
	s[3] = M(-0.48860251190291987f, x);
	s[4] = M(1.0925484305920792f, xy);
	s[5] = M(-1.0925484305920792f, yz);
	s[6] = S(M(0.94617469575755997f, z2), 0.31539156525251999f);
	s[7] = M(-1.0925484305920792f, xz);
	s[8] = S(M(0.54627421529603959f, x2), M(0.54627421529603959f, y2));
	s[9] = M(M(0.59004358992664352f, y), A(M(-3.0f, x2), y2));
	s[10] = M(M(2.8906114426405538f, xy), z);
	s[11] = M(M(0.45704579946446572f, y), S(1.0f, M(5.0f, z2)));
	s[12] = M(M(0.3731763325901154f, z), S(M(5.0f, z2), 3.0f));
	s[13] = M(M(0.45704579946446572f, x), S(1.0f, M(5.0f, z2)));
	s[14] = M(M(1.4453057213202769f, z), S(x2, y2));
	s[15] = M(M(0.59004358992664352f, x), A(-x2, M(3.0f, y2)));
#undef M
#undef A
#undef S
#pragma unroll
	for (int i = 0; i < 8; ++i) o[i] = __floats2half2_rn(s[2 * i], s[2 * i + 1]);
}

// ------------------------------------------------------------------------------------------------------------------
// shared-memory operand plumbing
// ------------------------------------------------------------------------------------------------------------------
// Copy one row-major [N x K] fp16 weight matrix from global into the chunk-major operand layout (umma.cuh).
__device__ __forceinline__ void stage_weights(const __half* __restrict__ w, uint32_t N, uint32_t K, uint8_t* dst, uint32_t tid, uint32_t nthreads) {
	const uint32_t kchunks = K >> 3;
	const uint32_t n_chunks = N * kchunks;
	for (uint32_t c = tid; c < n_chunks; c += nthreads) {
		const uint32_t n = c / kchunks, kc = c - n * kchunks;
		const uint4 v = __ldg(reinterpret_cast<const uint4*>(w + (size_t)n * K + kc * 8));
		*reinterpret_cast<uint4*>(dst + kc * (N * 16u) + n * 16u) = v;
	}
}

// store 8 halves (one chunk) of this thread's row
__device__ __forceinline__ void store_chunk(uint8_t* buf, uint32_t row, uint32_t kchunk, const __half2 (&h)[4]) {
	uint4 v;
	v.x = *reinterpret_cast<const uint32_t*>(&h[0]);
	v.y = *reinterpret_cast<const uint32_t*>(&h[1]);
	v.z = *reinterpret_cast<const uint32_t*>(&h[2]);
	v.w = *reinterpret_cast<const uint32_t*>(&h[3]);
	*reinterpret_cast<uint4*>(buf + kchunk * (TILE * 16u) + row * 16u) = v;
}

// D[128 x N] = A[128 x K] * W[N x K]^T ; one thread issues K/16 tcgen05.mma and commits to `bar`.
__device__ __forceinline__ void issue_layer_fwd(uint32_t a_smem, uint32_t K, uint32_t w_smem, uint32_t N, uint32_t tmem_d, uint64_t* bar) {
	const uint32_t idesc = umma::make_idesc(TILE, N, 0, 0);
	for (uint32_t k = 0; k < K; k += 16) {
		const uint64_t da = umma::make_desc(a_smem + (k >> 3) * (TILE * 16u), TILE * 16u, 128u);
		const uint64_t db = umma::make_desc(w_smem + (k >> 3) * (N * 16u), N * 16u, 128u);
		umma::mma_f16_ss(tmem_d, da, db, idesc, k > 0 ? 1u : 0u);
	}
	umma::commit(bar);
}

// dX[128 x K] = dY[128 x N] * W[N x K] : the same weight buffer read MN-major (its "K" dimension is now N).
__device__ __forceinline__ void issue_layer_dgrad(uint32_t dy_smem, uint32_t N, uint32_t w_smem, uint32_t K, uint32_t tmem_d, uint64_t* bar) {
	const uint32_t idesc = umma::make_idesc(TILE, K, 0, 1);
	for (uint32_t n = 0; n < N; n += 16) {
		const uint64_t da = umma::make_desc(dy_smem + (n >> 3) * (TILE * 16u), TILE * 16u, 128u);
		// B[K_out_dim = K (rows of B, "N" of the MMA)][reduction = n]: element (k, n) lives at (k/8)*(N*16) + n*16 + (k%8)*2
		// MN-major canonical form: (mn/8)*SBO + (kk/8)*LBO + (kk%8)*16 + (mn%8)*2  with mn = k, kk = n  => SBO = N*16, LBO = 128
		const uint64_t db = umma::make_desc(w_smem + n * 16u, 128u, N * 16u);
		umma::mma_f16_ss(tmem_d, da, db, idesc, n > 0 ? 1u : 0u);
	}
	umma::commit(bar);
}

// dW[M x N] += P[128 x M]^T * Q[128 x N] over the 128 samples of the tile (both operands read MN-major, K = samples).
// No commit here: the caller commits once per tile.
__device__ __forceinline__ void issue_wgrad(uint32_t p_smem, uint32_t M, uint32_t q_smem, uint32_t N, uint32_t tmem_d, uint32_t accumulate) {
	const uint32_t idesc = umma::make_idesc(M, N, 1, 1);
	for (uint32_t s = 0; s < TILE; s += 16) {
		// element (sample s, col c) at (c/8)*(TILE*16) + s*16 + (c%8)*2 ; mn = c, kk = s  => SBO = TILE*16, LBO = 128
		const uint64_t da = umma::make_desc(p_smem + s * 16u, 128u, TILE * 16u);
		const uint64_t db = umma::make_desc(q_smem + s * 16u, 128u, TILE * 16u);
		umma::mma_f16_ss(tmem_d, da, db, idesc, (accumulate || s > 0) ? 1u : 0u);
	}
}

__device__ __forceinline__ __half2 relu2(__half2 v) { return __hmax2(v, __float2half2_rn(0.0f)); }

// Epilogue of a hidden layer: this thread's TMEM row (64 fp32) -> ReLU -> fp16 -> its row of a [128 x 64] operand buffer.
template <bool RELU>
__device__ __forceinline__ void tmem_row_to_smem64(uint32_t taddr_row, uint8_t* buf, uint32_t row) {
	uint32_t v0[16], v1[16], v2[16], v3[16];
	umma::tmem_ld16(taddr_row + 0, v0);
	umma::tmem_ld16(taddr_row + 16, v1);
	umma::tmem_ld16(taddr_row + 32, v2);
	umma::tmem_ld16(taddr_row + 48, v3);
	umma::tmem_ld_wait();
	auto emit = [&](const uint32_t(&v)[16], uint32_t kc0) {
#pragma unroll
		for (uint32_t c = 0; c < 2; ++c) {
			__half2 h[4];
#pragma unroll
			for (uint32_t j = 0; j < 4; ++j) {
				__half2 t = __floats2half2_rn(__uint_as_float(v[c * 8 + 2 * j]), __uint_as_float(v[c * 8 + 2 * j + 1]));
				h[j] = RELU ? relu2(t) : t;
			}
			store_chunk(buf, row, kc0 + c, h);
		}
	};
	emit(v0, 0);
	emit(v1, 2);
	emit(v2, 4);
	emit(v3, 6);
}

// Epilogue of a 16-wide output layer: returns the 16 fp16 outputs of this thread's row (no activation).
__device__ __forceinline__ void tmem_row_to_regs16(uint32_t taddr_row, __half2 (&o)[8]) {
	uint32_t v[16];
	umma::tmem_ld16(taddr_row, v);
	umma::tmem_ld_wait();
#pragma unroll
	for (uint32_t j = 0; j < 8; ++j) o[j] = __floats2half2_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
}

// ----------------------------------------------------------------------------------------------------------------
// shared-memory carve-up for the forward kernels
// ----------------------------------------------------------------------------------------------------------------
struct FwdSmem {
	uint32_t w_bytes;   // all MLP weights
	uint32_t a0_off;    // [128 x 32] encoded positions (aliases h)
	uint32_t h_off;     // [128 x 64] hidden activations
	uint32_t a2_off;    // [128 x 32] rgb-net input: density-net output | SH
	uint32_t bar_off;   // mbarrier (8 B) + tmem base (4 B)
	uint32_t total;
};
__host__ __device__ inline FwdSmem fwd_smem_layout(uint32_t n_hidden_density, uint32_t n_hidden_rgb) {
	FwdSmem s;
	s.w_bytes = (mlp_n_params(n_hidden_density) + mlp_n_params(n_hidden_rgb)) * 2u;
	// A0 aliases the first half of H: a layer's epilogue writes H only after its MMA (the last reader of A0) has committed,
	// and the next tile's gather writes A0 only after the last layer's commit.  44 KB per CTA -> 5 CTAs per SM.
	s.h_off = s.w_bytes;
	s.a0_off = s.h_off;
	s.a2_off = s.h_off + TILE * MLP_WIDTH * 2u;
	s.bar_off = s.a2_off + TILE * ENC_WIDTH * 2u;
	s.total = s.bar_off + 16u;
	return s;
}

// Runs one MLP (first layer reads `in_buf` [128x32], hidden layers ping through `h_buf`) and returns the 16 outputs
// of this thread's row.  `w_smem` points at the MLP's first weight matrix in shared memory.
__device__ __forceinline__ void run_mlp_fwd(
	uint8_t* smem, uint32_t in_off, uint32_t h_off, uint32_t w_off, uint32_t n_hidden, uint32_t tmem_base, uint64_t* bar, uint32_t& phase,
	uint32_t tid, __half2 (&out)[8]
) {
	const uint32_t smem_base = umma::smem_u32(smem);
	const uint32_t lane_taddr = tmem_base + ((tid & ~31u) << 16);
	const uint32_t n_layers = mlp_n_layers(n_hidden);
	for (uint32_t l = 0; l < n_layers; ++l) {
		const uint32_t K = mlp_layer_in(n_hidden, l), N = mlp_layer_out(n_hidden, l);
		const uint32_t a_off = (l == 0) ? in_off : h_off;
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
		if (l + 1 < n_layers) {
			tmem_row_to_smem64<true>(lane_taddr, smem + h_off, tid);
		} else {
			tmem_row_to_regs16(lane_taddr, out);
		}
	}
}


NetDev make_netdev(const ngp_nerf_desc& d);
int device_sm_count();

}  // namespace ngpb
