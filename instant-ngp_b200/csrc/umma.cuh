// umma.cuh — thin inline-PTX layer over Blackwell's 5th-generation tensor cores (tcgen05) for sm_100a.
//
// Shared-memory operand convention used by every kernel in this library ("chunk-major, no swizzle"):
// a [rows x K] fp16 operand is stored as 16-byte chunks of 8 consecutive K elements,
//     byte_offset(row, k) = (k / 8) * (rows * 16) + row * 16 + (k % 8) * 2.
// That is the canonical SWIZZLE_NONE layout of the UMMA shared-memory descriptor
//   * read K-major  (rows = M or N):  SBO = 128 B (next group of 8 rows), LBO = rows*16 B (next 8 K elements)
//   * read MN-major (rows = K, the chunk index runs along M/N): SBO = rows*16 B, LBO = 128 B
// so one buffer serves forward (K-major), data-gradient (weights read MN-major) and weight-gradient
// (activations read MN-major, K = samples) without any transposition.
#pragma once

#include <cuda_fp16.h>
#include <stdint.h>

namespace ngpb {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (sm_100).
//   bits [0,14)  start address >> 4      bits [16,30) leading-dimension byte offset >> 4
//   bits [32,46) stride-dimension byte offset >> 4      bits [46,48) = 1 (descriptor version)      bits [61,64) = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
	uint64_t d = 0;
	d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
	d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
	d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
	d |= (uint64_t)1 << 46;
	return d;
}

// 32-bit instruction descriptor for kind::f16: fp16 A/B, fp32 accumulate.
//   [4,6) D format = 1 (f32)   [7,10) A format = 0 (f16)   [10,13) B format = 0 (f16)
//   bit 15 A major (0 = K, 1 = MN)   bit 16 B major   [17,23) N >> 3   [24,29) M >> 4
__device__ __forceinline__ uint32_t make_idesc(uint32_t m, uint32_t n, uint32_t a_mn_major, uint32_t b_mn_major) {
	uint32_t d = 0;
	d |= 1u << 4;
	d |= (a_mn_major & 1u) << 15;
	d |= (b_mn_major & 1u) << 16;
	d |= ((n >> 3) & 0x3Fu) << 17;
	d |= ((m >> 4) & 0x1Fu) << 24;
	return d;
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
		"}\n" ::"r"(tmem_d),
		"l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
		: "memory");
}

// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void commit(uint64_t* bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand fetch)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// TMEM allocation: executed by one full warp; the base address lands in *smem_out.
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_out) {
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_out)), "n"(NCOLS) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// mbarrier helpers (count-1 barrier used as "MMA done" signal)
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	const uint32_t addr = smem_u32(bar);
	uint32_t done;
	do {
		asm volatile(
			"{\n\t"
			".reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t"
			"}\n"
			: "=r"(done)
			: "r"(addr), "r"(parity)
			: "memory");
	} while (!done);
}

// TMEM -> registers: each thread of a warp reads its own lane (32 lanes per warp), N consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
		: "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
		  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
		: "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// byte offset of element (row, k) inside a chunk-major operand buffer with `rows` rows
__device__ __forceinline__ uint32_t chunk_off(uint32_t rows, uint32_t row, uint32_t k) { return (k >> 3) * (rows * 16u) + row * 16u + (k & 7u) * 2u; }

}  // namespace umma
}  // namespace ngpb
