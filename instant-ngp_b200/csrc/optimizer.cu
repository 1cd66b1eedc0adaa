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

// 8 parameters per thread: one 128-bit gradient load; the heavy fp32 state is touched only where the gradient is non-zero.
__global__ void __launch_bounds__(256) k_adam_ema(const AdamDev a, float* __restrict__ w32, __half* __restrict__ w16, __half* __restrict__ ema,
	__half* __restrict__ grads, float* __restrict__ m1, float* __restrict__ m2, uint32_t* __restrict__ steps) {
	const uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) * 8u;
	if (base >= a.n_total) return;
	const uint32_t n = (a.n_total - base) < 8u ? (a.n_total - base) : 8u;
	__half g[8];
	if (n == 8) {
		*reinterpret_cast<uint4*>(g) = *reinterpret_cast<const uint4*>(grads + base);
		*reinterpret_cast<uint4*>(grads + base) = make_uint4(0, 0, 0, 0);
	} else {
		for (uint32_t k = 0; k < n; ++k) {
			g[k] = grads[base + k];
			grads[base + k] = __float2half_rn(0.0f);
		}
	}
	bool touched[8];
	bool any = false;
#pragma unroll
	for (uint32_t k = 0; k < 8; ++k) {
		touched[k] = false;
		if (k < n) {
			const uint32_t i = base + k;
			const float gv = __half2float(g[k]);
			if (i < a.n_matrix) {
				touched[k] = a.optimize_matrix != 0;
			} else {
				touched[k] = a.optimize_non_matrix != 0 && gv != 0.0f;
			}
			if (touched[k]) {
				adam_one(a, i, gv, w32, w16, m1, m2, steps);
				any = true;
			}
		}
	}
	// EMA of the working weights -> inference weights (ema.h:63-77); every entry, every step
	(void)any;
	if (n == 8) {
		__half wv[8], ev[8];
		*reinterpret_cast<uint4*>(wv) = *reinterpret_cast<const uint4*>(w16 + base);
		*reinterpret_cast<uint4*>(ev) = *reinterpret_cast<const uint4*>(ema + base);
#pragma unroll
		for (uint32_t k = 0; k < 8; ++k) {
			const float f = (__half2float(ev[k]) * a.ema_decay * a.ema_debias_old + __half2float(wv[k]) * (1.0f - a.ema_decay)) * a.ema_debias_new;
			ev[k] = __float2half_rn(f);
		}
		*reinterpret_cast<uint4*>(ema + base) = *reinterpret_cast<const uint4*>(ev);
	} else {
		for (uint32_t k = 0; k < n; ++k) {
			const float f = (__half2float(ema[base + k]) * a.ema_decay * a.ema_debias_old + __half2float(w16[base + k]) * (1.0f - a.ema_decay)) * a.ema_debias_new;
			ema[base + k] = __float2half_rn(f);
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
	const uint32_t n_threads = div_round_up(d.n_params, 8);
	k_adam_ema<<<div_round_up(n_threads, 256), 256, 0, stream>>>(a, params_fp32, params_fp16, params_ema, grads, m1, m2, steps);
	NGPB_LAUNCHED();
	NGPB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ngpb
