/* ngp_net_cpu.c — the NeRF network of the hot path (hash-grid encoding + density MLP + SH-4 + rgb MLP, forward and backward) as a plain
 * multi-threaded CPU program: the CPU BASELINE of bench.py (`cpu_baseline`, `--impl reference`).  TEST / MEASUREMENT INFRASTRUCTURE ONLY:
 * nothing under instant-ngp_b200/ links or calls it.
 *
 * tiny-cuda-nn has no CPU implementation, so this is the "port" leg of the baseline: the same algorithm as
 * oracle/net_oracle.py (which restates grid.h:48-320, fully_fused_mlp.cu:499-557, nerf_network.h:105-268 and is pinned against the
 * reference's own objects), in fp32, OpenMP over samples, the way a CPU user would write it.  It is checked against net_oracle.py to fp16
 * rounding in tests/test_cpu_baseline.py; it is NOT a parity oracle (fp32 arithmetic, not the reference's fp16).
 *
 * Parameter layout (nerf_network.h:357-372): density MLP | rgb MLP | hash grid; MLP matrices row-major [out x in]. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))
#define W 64
#define NIN 32
#define NOUT 16
#define MAXH 4

typedef struct {
	uint32_t n_levels, F;
	const uint32_t* offsets;      /* n_levels + 1, in entries */
	const uint32_t* resolutions;
	const float* scales;
	uint32_t n_hidden_density, n_hidden_rgb;
} net_t;

static uint32_t grid_index(uint32_t x, uint32_t y, uint32_t z, uint32_t res, uint32_t size) {
	uint32_t index;
	const uint64_t stride = res <= 0x659u ? (uint64_t)res * res * res : 0xFFFFFFFFull;
	if ((uint64_t)size < stride) index = (x * 1u) ^ (y * 2654435761u) ^ (z * 805459861u);
	else index = x + y * res + z * res * res;
	return index % size;
}

static void sh4(const float* d01, float* o) {
	const float x = d01[0] * 2.0f - 1.0f, y = d01[1] * 2.0f - 1.0f, z = d01[2] * 2.0f - 1.0f;
	const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
	o[0] = 0.28209479177387814f; o[1] = -0.48860251190291987f * y; o[2] = 0.48860251190291987f * z; o[3] = -0.48860251190291987f * x;
	o[4] = 1.0925484305920792f * xy; o[5] = -1.0925484305920792f * yz; o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
	o[7] = -1.0925484305920792f * xz; o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
	o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2); o[10] = 2.8906114426405538f * xy * z; o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
	o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f); o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
	o[14] = 1.4453057213202769f * z * (x2 - y2); o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

static uint32_t mlp_n_params(uint32_t n_hidden) { return W * NIN + (n_hidden - 1) * W * W + NOUT * W; }

/* acts[0] = input (32), acts[l+1] = layer l output; returns the 16 outputs in acts[n_hidden + 1] */
static void mlp_forward(const float* w, uint32_t n_hidden, float acts[MAXH + 2][W]) {
	uint32_t in_dim = NIN;
	for (uint32_t l = 0; l <= n_hidden; ++l) {
		const uint32_t out_dim = l == n_hidden ? NOUT : W;
		for (uint32_t o = 0; o < out_dim; ++o) {
			float s = 0.0f;
			const float* row = w + (size_t)o * in_dim;
			for (uint32_t k = 0; k < in_dim; ++k) s += row[k] * acts[l][k];
			acts[l + 1][o] = (l == n_hidden) ? s : (s > 0.0f ? s : 0.0f);
		}
		w += (size_t)out_dim * in_dim;
		in_dim = out_dim;
	}
}
/* dy: gradient wrt the 16 outputs; accumulates weight gradients into gw (same layout as w); dx: gradient wrt the 32 inputs */
static void mlp_backward(const float* w, uint32_t n_hidden, float acts[MAXH + 2][W], const float* dy_out, float* gw, float* dx_out) {
	size_t off[MAXH + 2];
	uint32_t dims[MAXH + 2];
	dims[0] = NIN;
	size_t o = 0;
	for (uint32_t l = 0; l <= n_hidden; ++l) {
		off[l] = o;
		dims[l + 1] = l == n_hidden ? NOUT : W;
		o += (size_t)dims[l + 1] * dims[l];
	}
	float dy[W], dx[W];
	memcpy(dy, dy_out, sizeof(float) * NOUT);
	for (int l = (int)n_hidden; l >= 0; --l) {
		const uint32_t in_dim = dims[l], out_dim = dims[l + 1];
		const float* wl = w + off[l];
		float* gl = gw + off[l];
		for (uint32_t k = 0; k < in_dim; ++k) dx[k] = 0.0f;
		for (uint32_t q = 0; q < out_dim; ++q) {
			const float g = dy[q];
			if (g == 0.0f) continue;
			for (uint32_t k = 0; k < in_dim; ++k) {
				gl[(size_t)q * in_dim + k] += g * acts[l][k];
				dx[k] += g * wl[(size_t)q * in_dim + k];
			}
		}
		if (l > 0) for (uint32_t k = 0; k < in_dim; ++k) dy[k] = acts[l][k] > 0.0f ? dx[k] : 0.0f;
	}
	memcpy(dx_out, dx, sizeof(float) * NIN);
}

/* coords: n x 7 (pos, dt, dir in [0,1]); params: fp32 copy of the flat parameter buffer; out: n x 4 (rgb raw x3, density raw).
 * dL_dout (n x 4) == NULL: forward only.  grads: n_params floats, accumulated into (the caller zeroes them). */
EXPORT void orc_net_cpu(uint32_t n, const float* coords, const float* params, const float* dL_dout, float* grads, float* out, uint32_t n_levels, uint32_t F,
	const uint32_t* offsets, const uint32_t* resolutions, const float* scales, uint32_t n_hidden_density, uint32_t n_hidden_rgb) {
	const uint32_t n_dens = mlp_n_params(n_hidden_density), n_rgb = mlp_n_params(n_hidden_rgb), n_mlp = n_dens + n_rgb;
	const float* wd = params;
	const float* wr = params + n_dens;
	const float* table = params + n_mlp;
	float* gtable = grads ? grads + n_mlp : NULL;
#pragma omp parallel
	{
		float* gw = dL_dout ? (float*)calloc(n_mlp, sizeof(float)) : NULL;   /* per-thread MLP weight gradients */
#pragma omp for schedule(static)
		for (int64_t s = 0; s < (int64_t)n; ++s) {
			const float* c = coords + (size_t)s * 7;
			float acts_d[MAXH + 2][W], acts_r[MAXH + 2][W];
			uint32_t idx[16][8];
			float wgt[16][8];
			/* hash grid */
			for (uint32_t l = 0; l < n_levels; ++l) {
				const float scale = scales[l];
				const uint32_t res = resolutions[l], size = offsets[l + 1] - offsets[l];
				float p[3], fl[3], w1[3];
				uint32_t gi[3];
				for (int d = 0; d < 3; ++d) {
					p[d] = fmaf(scale, c[d], 0.5f);
					fl[d] = floorf(p[d]);
					gi[d] = (uint32_t)(int32_t)fl[d];
					w1[d] = p[d] - fl[d];
				}
				float acc[4] = {0, 0, 0, 0};
				for (uint32_t k = 0; k < 8; ++k) {
					const uint32_t bx = k & 1u, by = (k >> 1) & 1u, bz = (k >> 2) & 1u;
					const float wk = (bx ? w1[0] : 1.0f - w1[0]) * (by ? w1[1] : 1.0f - w1[1]) * (bz ? w1[2] : 1.0f - w1[2]);
					const uint32_t i = offsets[l] + grid_index(gi[0] + bx, gi[1] + by, gi[2] + bz, res, size);
					idx[l][k] = i;
					wgt[l][k] = wk;
					for (uint32_t f = 0; f < F; ++f) acc[f] += wk * table[(size_t)i * F + f];
				}
				for (uint32_t f = 0; f < F; ++f) acts_d[0][l * F + f] = acc[f];
			}
			mlp_forward(wd, n_hidden_density, acts_d);
			const float* dens_out = acts_d[n_hidden_density + 1];
			for (int k = 0; k < 16; ++k) acts_r[0][k] = dens_out[k];
			sh4(c + 4, acts_r[0] + 16);
			mlp_forward(wr, n_hidden_rgb, acts_r);
			const float* rgb_out = acts_r[n_hidden_rgb + 1];
			if (out) {
				out[(size_t)s * 4 + 0] = rgb_out[0]; out[(size_t)s * 4 + 1] = rgb_out[1]; out[(size_t)s * 4 + 2] = rgb_out[2]; out[(size_t)s * 4 + 3] = dens_out[0];
			}
			if (!dL_dout) continue;
			/* backward */
			float dy[NOUT] = {0}, dx_r[NIN], dx_d[NIN];
			dy[0] = dL_dout[(size_t)s * 4 + 0]; dy[1] = dL_dout[(size_t)s * 4 + 1]; dy[2] = dL_dout[(size_t)s * 4 + 2];
			mlp_backward(wr, n_hidden_rgb, acts_r, dy, gw + n_dens, dx_r);
			float dyd[NOUT];
			for (int k = 0; k < 16; ++k) dyd[k] = dx_r[k];
			dyd[0] += dL_dout[(size_t)s * 4 + 3];                      /* add_density_gradient (nerf_network.h:62-74) */
			mlp_backward(wd, n_hidden_density, acts_d, dyd, gw, dx_d);
			for (uint32_t l = 0; l < n_levels; ++l)
				for (uint32_t k = 0; k < 8; ++k)
					for (uint32_t f = 0; f < F; ++f) {
						const float g = wgt[l][k] * dx_d[l * F + f];
#pragma omp atomic
						gtable[(size_t)idx[l][k] * F + f] += g;
					}
		}
		if (gw) {
#pragma omp critical
			for (uint32_t k = 0; k < n_mlp; ++k) grads[k] += gw[k];
			free(gw);
		}
	}
}

EXPORT int orc_net_cpu_threads(void) {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/* Adam (adam.h:48-127, dense form): grads carry the loss scale; parameters untouched by the step (zero gradient) are skipped like the
 * reference's per-parameter step does. */
EXPORT void orc_adam_cpu(uint64_t n, float* params, const float* grads, float* m1, float* m2, float lr, float beta1, float beta2, float eps, float loss_scale, uint32_t step) {
	const float c1 = 1.0f - powf(beta1, (float)step), c2 = 1.0f - powf(beta2, (float)step);
#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t)n; ++i) {
		const float g = grads[i] / loss_scale;
		if (g == 0.0f) continue;
		const float a = m1[i] = beta1 * m1[i] + (1.0f - beta1) * g;
		const float b = m2[i] = beta2 * m2[i] + (1.0f - beta2) * g * g;
		params[i] -= lr * (a / c1) / (sqrtf(b / c2) + eps);
	}
}
