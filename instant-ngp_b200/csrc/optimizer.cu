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
__device__ __forceinline__ void adam_one(const AdamDev& a, uint32_t i, float g_scaled, float* __restrict__ w32, __half* __restrict__ w16,
	float* __restrict__ m1, float* __restrict__ m2, uint32_t* __restrict__ steps) {
	float gradient = g_scaled * a.inv_loss_scale;
	const float w = w32[i];
	if (i < a.n_matrix) gradient += a.l2_reg * w;
	const float m = a.beta1 * m1[i] + (1.0f - a.beta1) * gradient;
	const float v = a.beta2 * m2[i] + (1.0f - a.beta2) * gradient * gradient;
	m1[i] = m;
	m2[i] = v;
	const uint32_t step = ++steps[i];
	const float fs = (float)step;
	const float lr = a.lr * sqrtf(1.0f - exp2f(fs * a.log2_beta2)) / (1.0f - exp2f(fs * a.log2_beta1));
	const float eff = fminf(fmaxf(lr / (sqrtf(v) + a.epsilon), 0.0f), 3.402823466e+38f);
	const float nw = w - eff * m;
	w32[i] = nw;
	w16[i] = __float2half_rn(nw);
}

// One warp owns 256 consecutive parameters and walks them as 4 rows of 64: lane l takes the pair (2l, 2l+1) of each row, so
// every access of the warp is one fully used 128-byte (fp16 arrays) or 256-byte (fp32 arrays) segment.  The first version
// gave each thread 8 consecutive parameters: every scalar access of a warp then straddled 32 sectors and the kernel ran at
// 18 % of DRAM throughput (profiles/r1a_first_correct_path.md).  The heavy fp32 state is touched only where needed.
__global__ void __launch_bounds__(256) k_adam_ema(const AdamDev a, float* __restrict__ w32, __half* __restrict__ w16, __half* __restrict__ ema,
	__half* __restrict__ grads, float* __restrict__ m1, float* __restrict__ m2, uint32_t* __restrict__ steps) {
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
	const uint32_t warp_base = warp * 256u;
	if (warp_base >= a.n_total) return;
	const float ema_w = 1.0f - a.ema_decay;
#pragma unroll
	for (uint32_t row = 0; row < 4; ++row) {
		const uint32_t i = warp_base + row * 64u + lane * 2u;
		if (i >= a.n_total) break;
		if (i + 1 < a.n_total) {
			const __half2 g2 = *reinterpret_cast<const __half2*>(grads + i);
			*reinterpret_cast<__half2*>(grads + i) = __float2half2_rn(0.0f);
			const float g0 = __low2float(g2), g1 = __high2float(g2);
			const bool t0 = (i < a.n_matrix) ? (a.optimize_matrix != 0) : (a.optimize_non_matrix != 0 && g0 != 0.0f);
			const bool t1 = (i + 1 < a.n_matrix) ? (a.optimize_matrix != 0) : (a.optimize_non_matrix != 0 && g1 != 0.0f);
			if (t0) adam_one(a, i, g0, w32, w16, m1, m2, steps);
			if (t1) adam_one(a, i + 1, g1, w32, w16, m1, m2, steps);
			// EMA of the working weights -> inference weights (ema.h:63-77); every entry, every step
			const __half2 w2 = *reinterpret_cast<const __half2*>(w16 + i);
			const __half2 e2 = *reinterpret_cast<const __half2*>(ema + i);
			const float f0 = (__low2float(e2) * a.ema_decay * a.ema_debias_old + __low2float(w2) * ema_w) * a.ema_debias_new;
			const float f1 = (__high2float(e2) * a.ema_decay * a.ema_debias_old + __high2float(w2) * ema_w) * a.ema_debias_new;
			*reinterpret_cast<__half2*>(ema + i) = __floats2half2_rn(f0, f1);
		} else {
			const float g0 = __half2float(grads[i]);
			grads[i] = __float2half_rn(0.0f);
			const bool t0 = (i < a.n_matrix) ? (a.optimize_matrix != 0) : (a.optimize_non_matrix != 0 && g0 != 0.0f);
			if (t0) adam_one(a, i, g0, w32, w16, m1, m2, steps);
			const float f0 = (__half2float(ema[i]) * a.ema_decay * a.ema_debias_old + __half2float(w16[i]) * ema_w) * a.ema_debias_new;
			ema[i] = __float2half_rn(f0);
		}
	}
}

void optimizer_step(const ngp_nerf_desc& d, cudaStream_t stream, const ngp_adam_cfg& cfg, float* params_fp32, __half* params_fp16, __half* params_ema,
	__half* grads, float* m1, float* m2, uint32_t* steps) {
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
	a.n_matrix = d.n_mlp_params;
	a.n_total = d.n_params;
	a.optimize_matrix = cfg.optimize_matrix_params;
	a.optimize_non_matrix = cfg.optimize_non_matrix_params;
	const uint32_t n_warps = div_round_up(d.n_params, 256);
	k_adam_ema<<<div_round_up(n_warps * 32, 256), 256, 0, stream>>>(a, params_fp32, params_fp16, params_ema, grads, m1, m2, steps);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ngpb
