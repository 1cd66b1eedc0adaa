// optimizer.cu — fused Ema{ExponentialDecay{Adam}} step.
//   ≙ adam_step (tiny-cuda-nn/optimizers/adam.h:48-127) + ema_step_half_precision (ema.h:63-77), one pass over the
//   parameters instead of two kernels plus a 26 MB gradient memset: the gradient is read and ZEROED here, so the next
//   step's backward accumulates into a clean buffer without a separate cudaMemsetAsync (grid.h:865-867).
//   The learning-rate decay of exponential_decay.h:60-72 is folded into cfg.learning_rate by the host.
// Semantics kept from the reference: per-parameter step counter for bias correction, hash-grid ("non-matrix") entries
// with an exactly-zero gradient are skipped entirely (adam.h:79-83), L2 regularisation on matrix params only, fp32 master
// weights with an fp16 working copy, EMA over the fp16 working copy with debiasing.
#include "common.cuh"

namespace ngpb {

struct AdamDev {
	float lr, beta1, beta2, epsilon, l2_reg, inv_loss_scale;
	float log2_beta1, log2_beta2;
	float ema_decay, ema_debias_old, ema_debias_new;
	uint32_t n_matrix, n_total;
	uint32_t optimize_matrix, optimize_non_matrix;
};

// Adam's bias correction uses powf(beta, step) with a per-parameter step (adam.h:111-113).  beta^step is evaluated as
// exp2(step * log2(beta)) with the hardware ex2 unit (2 ulp): the libm powf costs ~100 instructions per parameter and made
// this HBM-streaming kernel instruction-bound (ncu: 17 % DRAM throughput, profiles/r1_kernels.md).
__device__ __forceinline__ float adam_one(const AdamDev& a, bool matrix, float g_scaled, float& w, float& m, float& v, uint32_t& step) {
	float gradient = g_scaled * a.inv_loss_scale;
	if (matrix) gradient += a.l2_reg * w;
	m = a.beta1 * m + (1.0f - a.beta1) * gradient;
	v = a.beta2 * v + (1.0f - a.beta2) * gradient * gradient;
	++step;
	const float fs = (float)step;
	const float lr = a.lr * sqrtf(1.0f - exp2f(fs * a.log2_beta2)) / (1.0f - exp2f(fs * a.log2_beta1));
	const float eff = fminf(fmaxf(lr / (sqrtf(v) + a.epsilon), 0.0f), 3.402823466e+38f);
	w = w - eff * m;
	return w;
}

// One warp owns 256 consecutive parameters and walks them as 4 rows of 64: lane l takes the pair (2l, 2l+1) of each row, so
// every access of the warp is one fully used 128-byte (fp16 arrays) or 256-byte (fp32 arrays) segment.  The kernel is a
// latency problem, not a bandwidth one (ncu r1b: 241 cycles of long-scoreboard stall per issue at 83 % occupancy): all
// loads of a phase are issued for the 4 rows before anything depends on them — first the three fp16 streams every entry
// needs, then, only for pairs with a touched entry, the four fp32/u32 state streams (a pair's untouched twin is loaded
// and written back unchanged; it shares the sector).
constexpr uint32_t ADAM_ROWS = 4;
__global__ void __launch_bounds__(256) k_adam_ema(const AdamDev a, float* __restrict__ w32, __half* __restrict__ w16, __half* __restrict__ ema,
	__half* __restrict__ grads, float* __restrict__ m1, float* __restrict__ m2, uint32_t* __restrict__ steps) {
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
	const uint32_t warp_base = warp * (64u * ADAM_ROWS);
	if (warp_base >= a.n_total) return;
	const float ema_w = 1.0f - a.ema_decay;
	const uint32_t n_even = a.n_total & ~1u;

	__half2 g2[ADAM_ROWS], w2[ADAM_ROWS], e2[ADAM_ROWS];
	bool in[ADAM_ROWS], t0[ADAM_ROWS], t1[ADAM_ROWS];
#pragma unroll
	for (uint32_t r = 0; r < ADAM_ROWS; ++r) {
		const uint32_t i = warp_base + r * 64u + lane * 2u;
		in[r] = i < n_even;
		g2[r] = w2[r] = e2[r] = __float2half2_rn(0.0f);
		if (in[r]) {
			g2[r] = *reinterpret_cast<const __half2*>(grads + i);
			w2[r] = *reinterpret_cast<const __half2*>(w16 + i);
			e2[r] = *reinterpret_cast<const __half2*>(ema + i);
		}
	}
	float2 W[ADAM_ROWS], M[ADAM_ROWS], V[ADAM_ROWS];
	uint2 S[ADAM_ROWS];
#pragma unroll
	for (uint32_t r = 0; r < ADAM_ROWS; ++r) {
		const uint32_t i = warp_base + r * 64u + lane * 2u;
		t0[r] = in[r] && ((i < a.n_matrix) ? (a.optimize_matrix != 0) : (a.optimize_non_matrix != 0 && __low2float(g2[r]) != 0.0f));
		t1[r] = in[r] && ((i + 1 < a.n_matrix) ? (a.optimize_matrix != 0) : (a.optimize_non_matrix != 0 && __high2float(g2[r]) != 0.0f));
		if (t0[r] || t1[r]) {
			W[r] = *reinterpret_cast<const float2*>(w32 + i);
			M[r] = *reinterpret_cast<const float2*>(m1 + i);
			V[r] = *reinterpret_cast<const float2*>(m2 + i);
			S[r] = *reinterpret_cast<const uint2*>(steps + i);
		}
	}
#pragma unroll
	for (uint32_t r = 0; r < ADAM_ROWS; ++r) {
		const uint32_t i = warp_base + r * 64u + lane * 2u;
		if (!in[r]) continue;
		*reinterpret_cast<__half2*>(grads + i) = __float2half2_rn(0.0f);
		__half2 wn = w2[r];
		if (t0[r] || t1[r]) {
			if (t0[r]) wn = __halves2half2(__float2half_rn(adam_one(a, i < a.n_matrix, __low2float(g2[r]), W[r].x, M[r].x, V[r].x, S[r].x)), __high2half(wn));
			if (t1[r]) wn = __halves2half2(__low2half(wn), __float2half_rn(adam_one(a, i + 1 < a.n_matrix, __high2float(g2[r]), W[r].y, M[r].y, V[r].y, S[r].y)));
			*reinterpret_cast<float2*>(w32 + i) = W[r];
			*reinterpret_cast<float2*>(m1 + i) = M[r];
			*reinterpret_cast<float2*>(m2 + i) = V[r];
			*reinterpret_cast<uint2*>(steps + i) = S[r];
			*reinterpret_cast<__half2*>(w16 + i) = wn;
		}
		// EMA of the (updated) working weights -> inference weights (ema.h:63-77); every entry, every step
		const float f0 = (__low2float(e2[r]) * a.ema_decay * a.ema_debias_old + __low2float(wn) * ema_w) * a.ema_debias_new;
		const float f1 = (__high2float(e2[r]) * a.ema_decay * a.ema_debias_old + __high2float(wn) * ema_w) * a.ema_debias_new;
		*reinterpret_cast<__half2*>(ema + i) = __floats2half2_rn(f0, f1);
	}
	// odd tail element (n_total odd): the last lane that would have covered it
	if ((a.n_total & 1u) && warp_base + 64u * ADAM_ROWS >= a.n_total && lane == 0) {
		const uint32_t i = a.n_total - 1u;
		if (i >= warp_base) {
			const float g0 = __half2float(grads[i]);
			grads[i] = __float2half_rn(0.0f);
			const bool t = (i < a.n_matrix) ? (a.optimize_matrix != 0) : (a.optimize_non_matrix != 0 && g0 != 0.0f);
			if (t) {
				float w = w32[i], m = m1[i], v = m2[i];
				uint32_t st = steps[i];
				adam_one(a, i < a.n_matrix, g0, w, m, v, st);
				w32[i] = w; m1[i] = m; m2[i] = v; steps[i] = st;
				w16[i] = __float2half_rn(w);
			}
			const float f0 = (__half2float(ema[i]) * a.ema_decay * a.ema_debias_old + __half2float(w16[i]) * ema_w) * a.ema_debias_new;
			ema[i] = __float2half_rn(f0);
		}
	}
}

// n_matrix_params leading "matrix" parameters (MLP weights: L2-regularised, always stepped), the rest are hash-grid entries
void optimizer_step_flat(uint32_t n_matrix_params, uint32_t n_params, cudaStream_t stream, const ngp_adam_cfg& cfg, float* params_fp32, __half* params_fp16,
	__half* params_ema, __half* grads, float* m1, float* m2, uint32_t* steps) {
	NGPB_CHECK(cfg.ema_step >= 1, "ngp_optimizer_step: ema_step is 1-based");
	AdamDev a;
	a.lr = cfg.learning_rate;
	a.beta1 = cfg.beta1;
	a.beta2 = cfg.beta2;
	a.epsilon = cfg.epsilon;
	a.l2_reg = cfg.l2_reg;
	a.inv_loss_scale = 1.0f / cfg.loss_scale;
	a.log2_beta1 = std::log2(cfg.beta1);
	a.log2_beta2 = std::log2(cfg.beta2);
	a.ema_decay = cfg.ema_decay;
	a.ema_debias_old = 1.0f - (float)std::pow(cfg.ema_decay, (float)(cfg.ema_step - 1));
	a.ema_debias_new = 1.0f / (1.0f - (float)std::pow(cfg.ema_decay, (float)cfg.ema_step));
	a.n_matrix = n_matrix_params;
	a.n_total = n_params;
	a.optimize_matrix = cfg.optimize_matrix_params;
	a.optimize_non_matrix = cfg.optimize_non_matrix_params;
	const uint32_t n_warps = div_round_up(n_params, 64u * ADAM_ROWS);
	k_adam_ema<<<div_round_up(n_warps * 32, 256), 256, 0, stream>>>(a, params_fp32, params_fp16, params_ema, grads, m1, m2, steps);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

void optimizer_step(const ngp_nerf_desc& d, cudaStream_t stream, const ngp_adam_cfg& cfg, float* params_fp32, __half* params_fp16, __half* params_ema,
	__half* grads, float* m1, float* m2, uint32_t* steps) {
	optimizer_step_flat(d.n_mlp_params, d.n_params, stream, cfg, params_fp32, params_fp16, params_ema, grads, m1, m2, steps);
}

}  // namespace ngpb
