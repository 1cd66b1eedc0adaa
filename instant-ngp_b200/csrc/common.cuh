// common.cuh — small device/host utilities shared by all kernels of libngp_b200.
// Everything here restates behaviour of the reference (file:line cited per item); none of it is copied.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <stdexcept>
#include <string>

#include "../../include/ngp_b200.h"

namespace ngpb {

// ------------------------------------------------------------------------------------------------------------------
// error handling at the C boundary (reference: CUDA_CHECK_THROW, tiny-cuda-nn/common_host.h:97-111)
// ------------------------------------------------------------------------------------------------------------------
void set_last_error(const std::string& msg);
extern unsigned long long g_launch_count;

#define NGPB_STR2(x) #x
#define NGPB_STR(x) NGPB_STR2(x)
#define NGPB_CUDA_CHECK(x)                                                                                             \
	do {                                                                                                                 \
		cudaError_t _e = (x);                                                                                              \
		if (_e != cudaSuccess)                                                                                             \
			throw std::runtime_error(std::string(__FILE__ ":" NGPB_STR(__LINE__) " " #x " failed: ") + cudaGetErrorString(_e)); \
	} while (0)
#define NGPB_CHECK(cond, msg)                                                     \
	do {                                                                            \
		if (!(cond)) throw std::runtime_error(std::string("ngp_b200: ") + (msg));     \
	} while (0)
#define NGPB_LAUNCHED() (++::ngpb::g_launch_count)

inline uint32_t div_round_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
inline uint32_t next_multiple(uint32_t a, uint32_t b) { return div_round_up(a, b) * b; }

// ------------------------------------------------------------------------------------------------------------------
// pcg32 — the published PCG-XSH-RR 64/32 generator (M. O'Neill), as used by the reference through
// tiny-cuda-nn/dependencies/pcg32/pcg32.h (seed :56-62, next_uint :65-71, next_float :104-113, advance :143-165).
// ------------------------------------------------------------------------------------------------------------------
struct Pcg32 {
	uint64_t state;
	uint64_t inc;

	static constexpr uint64_t MULT = 0x5851f42d4c957f2dULL;

	__host__ __device__ Pcg32() : state(0x853c49e6748fea9bULL), inc(0xda3e39cb94b95bdbULL) {}
	__host__ __device__ Pcg32(uint64_t s, uint64_t i, bool /*raw*/) : state(s), inc(i) {}
	__host__ __device__ explicit Pcg32(uint64_t initstate, uint64_t initseq = 1u) {
		state = 0u;
		inc = (initseq << 1u) | 1u;
		next_uint();
		state += initstate;
		next_uint();
	}
	__host__ __device__ uint32_t next_uint() {
		const uint64_t old = state;
		state = old * MULT + inc;
		const uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
		const uint32_t rot = (uint32_t)(old >> 59u);
		return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
	}
	__host__ __device__ float next_float() {
		const uint32_t u = (next_uint() >> 9) | 0x3f800000u;
#if defined(__CUDA_ARCH__)
		return __uint_as_float(u) - 1.0f;
#else
		float f;
		memcpy(&f, &u, 4);
		return f - 1.0f;
#endif
	}
	// jump ahead by delta draws in O(log delta) (Brown, "Random Number Generation with Arbitrary Stride")
	__host__ __device__ void advance(uint64_t delta = (1ull << 32)) {
		uint64_t cur_mult = MULT, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
		while (delta > 0) {
			if (delta & 1) {
				acc_mult *= cur_mult;
				acc_plus = acc_plus * cur_mult + cur_plus;
			}
			cur_plus = (cur_mult + 1) * cur_plus;
			cur_mult *= cur_mult;
			delta >>= 1;
		}
		state = acc_mult * state + acc_plus;
	}
};

// ------------------------------------------------------------------------------------------------------------------
// Morton codes for the 128^3 occupancy grid (tiny-cuda-nn/common_device.h:936-960 semantics).
// ------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline uint32_t morton_spread3(uint32_t v) {
	v = (v | (v << 16)) & 0xFF0000FFu;  // identical result to the multiply form for v < 1024
	v = (v | (v << 8)) & 0x0F00F00Fu;
	v = (v | (v << 4)) & 0xC30C30C3u;
	v = (v | (v << 2)) & 0x49249249u;
	return v;
}
__host__ __device__ inline uint32_t morton3d(uint32_t x, uint32_t y, uint32_t z) {
	return morton_spread3(x) | (morton_spread3(y) << 1) | (morton_spread3(z) << 2);
}
__host__ __device__ inline uint32_t morton_compact3(uint32_t x) {
	x &= 0x49249249u;
	x = (x | (x >> 2)) & 0xc30c30c3u;
	x = (x | (x >> 4)) & 0x0f00f00fu;
	x = (x | (x >> 8)) & 0xff0000ffu;
	x = (x | (x >> 16)) & 0x0000ffffu;
	return x;
}

// ------------------------------------------------------------------------------------------------------------------
// Hash-grid addressing (tiny-cuda-nn/common_device.h:787-791 coherent prime hash, :847-884 grid_index).
// ------------------------------------------------------------------------------------------------------------------
struct LevelMeta {
	uint32_t offset;       // first entry of the level
	uint32_t size;         // entries in the level ("hashmap_size")
	uint32_t resolution;   // grid vertices per axis
	uint32_t dense;        // 1: x + y*res + z*res^2 ; 0: xor-prime hash
	float scale;
};

__host__ __device__ inline uint32_t grid_index_3d(uint32_t x, uint32_t y, uint32_t z, uint32_t resolution, uint32_t size, bool dense) {
	uint32_t index;
	if (dense) {
		index = x + y * resolution + z * resolution * resolution;
	} else {
		index = (x * 1u) ^ (y * 2654435761u) ^ (z * 805459861u);
	}
	return index % size;
}

// The 8 corner indices of one level (corner c = x + 2y + 4z offsets, the order kernel_grid visits them) without an integer
// division: `index % size` of grid_index (common_device.h:880) costs ~20 instructions per corner and was a third of all
// instructions of the gather.  Hashed levels have a power-of-two size (checked in make_netdev), so the modulo is a mask and
// the per-axis products are shared by the corners; dense levels have index <= res + res^2 + res^3 < 2*size whenever the
// cell is inside the grid, so one conditional subtract is exact.  Cells outside [0, res) (positions outside the unit cube)
// take the generic path, which keeps the wrap-around semantics of the reference bit for bit.
__device__ __forceinline__ void level_corner_indices(const LevelMeta& lv, uint32_t gx, uint32_t gy, uint32_t gz, uint32_t (&idx)[8]) {
	if (lv.dense) {
		const uint32_t res = lv.resolution;
		if (gx < res && gy < res && gz < res) {
			const uint32_t r2 = res * res;
			const uint32_t b = gx + gy * res + gz * r2;
#pragma unroll
			for (uint32_t c = 0; c < 8; ++c) {
				uint32_t i = b + (c & 1u) + ((c & 2u) ? res : 0u) + ((c & 4u) ? r2 : 0u);
				if (i >= lv.size) i -= lv.size;
				idx[c] = i;
			}
		} else {
#pragma unroll
			for (uint32_t c = 0; c < 8; ++c) idx[c] = grid_index_3d(gx + (c & 1u), gy + ((c >> 1) & 1u), gz + ((c >> 2) & 1u), res, lv.size, true);
		}
	} else {
		const uint32_t mask = lv.size - 1u;
		const uint32_t hx[2] = {gx, gx + 1u};
		const uint32_t hy0 = gy * 2654435761u, hz0 = gz * 805459861u;
		const uint32_t hy[2] = {hy0, hy0 + 2654435761u};
		const uint32_t hz[2] = {hz0, hz0 + 805459861u};
#pragma unroll
		for (uint32_t c = 0; c < 8; ++c) idx[c] = (hx[c & 1u] ^ hy[(c >> 1) & 1u] ^ hz[(c >> 2) & 1u]) & mask;
	}
}

// Is a 3-D level stored densely?  grid_index (common_device.h:866-881): stride = res^3 when res <= 0x659, else
// 0xFFFFFFFF; hashed iff size < stride.
inline bool level_is_dense_3d(uint32_t resolution, uint32_t size) {
	if (resolution > 0x659u) return false;
	const uint64_t stride = (uint64_t)resolution * resolution * resolution;
	return !((uint64_t)size < stride);
}

}  // namespace ngpb
