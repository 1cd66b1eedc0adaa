// ref_tcnn_harness.cu — golden-vector generator that runs the REFERENCE network code on a GPU.
//
// Test infrastructure.  Links the reference's tiny-cuda-nn (compiled by oracle/ref/Makefile from the sources under
// /root/reference, offline kernels, no NVRTC) and instantiates the reference's own ngp::NerfNetwork<__half>
// (include/neural-graphics-primitives/nerf_network.h) and its Ema{ExponentialDecay{Adam}} optimizer.  Built here, run on the
// GPU box through gpurun (tests/golden/make_ref_tcnn_goldens.sh); the vectors it writes are committed under tests/golden/
// and pin the CPU oracle (oracle/net_oracle.py) and, through it, the CUDA kernels.
//
// File layout (little endian): u32 magic 'NGPR', u32 n_levels, u32 F, u32 log2_T, f32 per_level_scale, u32 n_params,
//   u32 n_samples, then fp16 params[n_params], f32 coords[n*7], fp16 inference_out[n*16], fp16 forward_out[n*16],
//   fp16 dL_dout[n*16], fp16 grads[n_params], then 2 optimizer steps: f32 w32[n_params], fp16 w16[n_params], fp16 ema[n_params], fp16 grads[n_params] (the gradient consumed by that step).
#include <neural-graphics-primitives/common.h>
#include <neural-graphics-primitives/nerf_network.h>

#include <tiny-cuda-nn/gpu_matrix.h>
#include <tiny-cuda-nn/loss.h>
#include <tiny-cuda-nn/network_with_input_encoding.h>
#include <tiny-cuda-nn/optimizer.h>

#include <cstdio>
#include <vector>

// tiny-cuda-nn's non-RTC stub (src/rtc_kernel.cu:63-80) omits this member; the reference's CMake build always compiles the
// RTC variant.  Never called here (JIT fusion is off), it only has to link.
namespace tcnn { void CudaRtcKernel::set(CUfunction_attribute, int) {} }

using namespace tcnn;
using namespace ngp;
using precision_t = network_precision_t;

template <typename T> static void put(FILE* f, const std::vector<T>& v) { fwrite(v.data(), sizeof(T), v.size(), f); }
template <typename T> static std::vector<T> download(const T* dev, size_t n) {
	std::vector<T> h(n);
	CUDA_CHECK_THROW(cudaMemcpy(h.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost));
	return h;
}

static int run(const char* path, uint32_t n_levels, uint32_t F, uint32_t log2_T, float per_level_scale, uint32_t n_samples) {
	json enc = {{"otype", "HashGrid"}, {"n_levels", n_levels}, {"n_features_per_level", F}, {"log2_hashmap_size", log2_T}, {"base_resolution", 16},
		{"per_level_scale", per_level_scale}};
	json dir_enc = {{"otype", "Composite"}, {"nested", {{{"n_dims_to_encode", 3}, {"otype", "SphericalHarmonics"}, {"degree", 4}}, {{"otype", "Identity"}}}}};
	json net = {{"otype", "FullyFusedMLP"}, {"activation", "ReLU"}, {"output_activation", "None"}, {"n_neurons", 64}, {"n_hidden_layers", 1}};
	json rgb = {{"otype", "FullyFusedMLP"}, {"activation", "ReLU"}, {"output_activation", "None"}, {"n_neurons", 64}, {"n_hidden_layers", 2}};
	auto network = std::make_shared<NerfNetwork<precision_t>>(3, 3, 0, 4, enc, dir_enc, net, rgb);
	network->set_jit_fusion(false);
	const size_t n_params = network->n_params();

	pcg32 rng{424242};
	std::vector<float> p32(n_params);
	const size_t n_mlp = 3072 + 7168;
	// MLPs: Xavier-like magnitudes; hash grid: "trained-like" values so that activations are O(0.1..1)
	for (size_t i = 0; i < n_params; ++i) p32[i] = (rng.next_float() * 2.0f - 1.0f) * (i < n_mlp ? 0.25f : 0.3f);
	std::vector<precision_t> p16(n_params);
	for (size_t i = 0; i < n_params; ++i) p16[i] = (precision_t)p32[i];
	for (size_t i = 0; i < n_params; ++i) p32[i] = (float)p16[i];

	GPUMemory<float> params_fp32(n_params);
	GPUMemory<precision_t> params(n_params), grads(n_params);
	params_fp32.copy_from_host(p32);
	params.copy_from_host(p16);
	grads.memset(0);
	network->set_params(params.data(), params.data(), grads.data());

	std::vector<float> coords(n_samples * 7);
	for (uint32_t i = 0; i < n_samples; ++i) {
		for (int k = 0; k < 3; ++k) coords[i * 7 + k] = rng.next_float();
		coords[i * 7 + 3] = rng.next_float();
		float d[3] = {rng.next_float() - 0.5f, rng.next_float() - 0.5f, rng.next_float() - 0.5f};
		float l = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) + 1e-6f;
		for (int k = 0; k < 3; ++k) coords[i * 7 + 4 + k] = (d[k] / l + 1.0f) * 0.5f;
	}
	if (n_samples > 4) { coords[0] = 0.0f; coords[1] = 1.0f; coords[2] = 0.5f; }  // exact boundary positions
	cudaStream_t stream = nullptr;
	GPUMatrix<float> input(7, n_samples);
	CUDA_CHECK_THROW(cudaMemcpy(input.data(), coords.data(), coords.size() * 4, cudaMemcpyHostToDevice));
	GPUMatrix<precision_t> out_inf(16, n_samples), out_fwd(16, n_samples), dL(16, n_samples);
	out_inf.memset(0);
	out_fwd.memset(0);

	network->inference_mixed_precision(stream, input, out_inf, false);
	auto ctx = network->forward(stream, input, &out_fwd, false, false);
	std::vector<precision_t> dl_host(n_samples * 16, (precision_t)0.0f);
	for (uint32_t i = 0; i < n_samples; ++i)
		for (int k = 0; k < 4; ++k) dl_host[i * 16 + k] = (precision_t)((rng.next_float() - 0.5f) * 0.5f);
	CUDA_CHECK_THROW(cudaMemcpy(dL.data(), dl_host.data(), dl_host.size() * sizeof(precision_t), cudaMemcpyHostToDevice));
	network->backward(stream, *ctx, input, out_fwd, dL, nullptr, false, GradientMode::Overwrite);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());

	FILE* f = fopen(path, "wb");
	if (!f) return 3;
	uint32_t hdr[4] = {0x5250474Eu, n_levels, F, log2_T};
	fwrite(hdr, 4, 4, f);
	fwrite(&per_level_scale, 4, 1, f);
	uint32_t np32 = (uint32_t)n_params;
	fwrite(&np32, 4, 1, f);
	fwrite(&n_samples, 4, 1, f);
	put(f, p16);
	put(f, coords);
	put(f, download(out_inf.data(), (size_t)n_samples * 16));
	put(f, download(out_fwd.data(), (size_t)n_samples * 16));
	put(f, dl_host);
	put(f, download(grads.data(), n_params));

	// ---- optimizer: configs/nerf/base.json's Ema{ExponentialDecay{Adam}} ------------------------------------------------
	json opt_cfg = {{"otype", "Ema"}, {"decay", 0.95}, {"nested", {{"otype", "ExponentialDecay"}, {"decay_start", 20000}, {"decay_interval", 10000},
		{"decay_base", 0.33}, {"nested", {{"otype", "Adam"}, {"learning_rate", 1e-2}, {"beta1", 0.9}, {"beta2", 0.99}, {"epsilon", 1e-15}, {"l2_reg", 1e-6}}}}}};
	std::shared_ptr<Optimizer<precision_t>> optimizer{create_optimizer<precision_t>(opt_cfg)};
	optimizer->allocate(network);
	// the trainer points inference params at the EMA weights and seeds them with the cast master weights (trainer.h:409-421)
	precision_t* ema = optimizer->custom_weights();
	CUDA_CHECK_THROW(cudaMemcpy(ema, params.data(), n_params * sizeof(precision_t), cudaMemcpyDeviceToDevice));
	for (int step = 0; step < 2; ++step) {
		if (step > 0) {
			// new gradients for steps 2 and 3: a fresh backward pass with the updated weights
			auto c2 = network->forward(stream, input, &out_fwd, false, false);
			network->backward(stream, *c2, input, out_fwd, dL, nullptr, false, GradientMode::Overwrite);
		}
		optimizer->step(stream, 128.0f, params_fp32.data(), params.data(), grads.data());
		CUDA_CHECK_THROW(cudaDeviceSynchronize());
		put(f, download(params_fp32.data(), n_params));
		put(f, download(params.data(), n_params));
		put(f, download(ema, n_params));
		put(f, download(grads.data(), n_params));
	}
	fclose(f);
	printf("wrote %s: L=%u F=%u T=2^%u n_params=%zu n=%u\n", path, n_levels, F, log2_T, n_params, n_samples);
	return 0;
}

// ---- image / SDF primitive: the reference's NetworkWithInputEncoding<__half> + Loss objects, as Trainer::forward / backward use
// them without JIT fusion (trainer.h:95-150).
// File layout: u32 magic 'NGPF', u32 n_pos_dims, u32 n_levels, u32 F, u32 log2_T, f32 per_level_scale, u32 n_hidden, u32 n_out,
//   u32 loss_type (0 L2, 2 MAPE), u32 n_params, u32 n_samples, then fp16 params[n_params], f32 positions[n*D], f32 targets[n*n_out],
//   fp16 inference_out[n*16], fp16 forward_out[n*16], f32 loss_values[n*16], fp16 dL_dout[n*16], fp16 grads[n_params].
static int run_field(const char* path, uint32_t n_pos, uint32_t n_levels, uint32_t F, uint32_t log2_T, float per_level_scale, uint32_t n_hidden, uint32_t n_out,
	uint32_t loss_type, uint32_t n_samples) {
	json enc = {{"otype", "HashGrid"}, {"n_levels", n_levels}, {"n_features_per_level", F}, {"log2_hashmap_size", log2_T}, {"base_resolution", 16},
		{"per_level_scale", per_level_scale}};
	json net = {{"otype", "FullyFusedMLP"}, {"activation", "ReLU"}, {"output_activation", "None"}, {"n_neurons", 64}, {"n_hidden_layers", n_hidden}};
	auto network = std::make_shared<NetworkWithInputEncoding<precision_t>>(n_pos, n_out, enc, net);
	network->set_jit_fusion(false);
	std::shared_ptr<Loss<precision_t>> loss{create_loss<precision_t>(json{{"otype", loss_type == 2 ? "MAPE" : "L2"}})};
	const size_t n_params = network->n_params();
	const size_t n_mlp = 64 * 32 + (n_hidden - 1) * 64 * 64 + 16 * 64;

	pcg32 rng{20240922};
	std::vector<float> p32(n_params);
	for (size_t i = 0; i < n_params; ++i) p32[i] = (rng.next_float() * 2.0f - 1.0f) * (i < n_mlp ? 0.25f : 0.3f);
	std::vector<precision_t> p16(n_params);
	for (size_t i = 0; i < n_params; ++i) p16[i] = (precision_t)p32[i];
	GPUMemory<precision_t> params(n_params), grads(n_params);
	params.copy_from_host(p16);
	grads.memset(0);
	network->set_params(params.data(), params.data(), grads.data());

	std::vector<float> pos((size_t)n_samples * n_pos), tgt((size_t)n_samples * n_out);
	for (auto& v : pos) v = rng.next_float();
	for (auto& v : tgt) v = rng.next_float() * 0.9f + 0.05f;
	for (uint32_t k = 0; k < n_pos; ++k) { pos[k] = 0.0f; pos[n_pos + k] = 1.0f; }  // exact boundary positions
	cudaStream_t stream = nullptr;
	GPUMatrix<float> input(n_pos, n_samples), target(n_out, n_samples);
	CUDA_CHECK_THROW(cudaMemcpy(input.data(), pos.data(), pos.size() * 4, cudaMemcpyHostToDevice));
	CUDA_CHECK_THROW(cudaMemcpy(target.data(), tgt.data(), tgt.size() * 4, cudaMemcpyHostToDevice));
	const uint32_t padded = network->padded_output_width();
	if (padded != 16) return 4;
	GPUMatrix<precision_t> out_inf(padded, n_samples), out_fwd(padded, n_samples), dL(padded, n_samples);
	GPUMatrix<float> values(padded, n_samples);
	out_inf.memset(0);
	out_fwd.memset(0);
	network->inference_mixed_precision(stream, input, out_inf, false);
	auto ctx = network->forward(stream, input, &out_fwd, false, false);
	loss->evaluate(stream, 128.0f, out_fwd, target, values, dL);
	network->backward(stream, *ctx, input, out_fwd, dL, nullptr, false, GradientMode::Overwrite);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());

	FILE* f = fopen(path, "wb");
	if (!f) return 3;
	uint32_t hdr[5] = {0x4650474Eu, n_pos, n_levels, F, log2_T};
	fwrite(hdr, 4, 5, f);
	fwrite(&per_level_scale, 4, 1, f);
	uint32_t tail[5] = {n_hidden, n_out, loss_type, (uint32_t)n_params, n_samples};
	fwrite(tail, 4, 5, f);
	put(f, p16);
	put(f, pos);
	put(f, tgt);
	put(f, download(out_inf.data(), (size_t)n_samples * 16));
	put(f, download(out_fwd.data(), (size_t)n_samples * 16));
	put(f, download(values.data(), (size_t)n_samples * 16));
	put(f, download(dL.data(), (size_t)n_samples * 16));
	put(f, download(grads.data(), n_params));
	fclose(f);
	printf("wrote %s: D=%u L=%u F=%u T=2^%u hidden=%u out=%u loss=%u n_params=%zu n=%u\n", path, n_pos, n_levels, F, log2_T, n_hidden, n_out, loss_type, n_params, n_samples);
	return 0;
}

int main(int argc, char** argv) {
	if (argc < 2) {
		fprintf(stderr, "usage: ref_tcnn <out_dir>\n");
		return 2;
	}
	try {
		std::string dir = argv[1];
		int rc = run((dir + "/ref_tcnn_L16F2.bin").c_str(), 16, 2, 10, 1.5157166f, 512);
		if (rc) return rc;
		rc = run((dir + "/ref_tcnn_L8F4.bin").c_str(), 8, 4, 10, 2.4380093f, 512);
		if (rc) return rc;
		// image primitive (2-D grid, 3 outputs, L2) and SDF primitive (3-D grid, 1 output, MAPE), configs/{image,sdf}/base.json shapes
		rc = run_field((dir + "/ref_field_image.bin").c_str(), 2, 16, 2, 10, 1.3819129f, 2, 3, 0, 512);
		if (rc) return rc;
		rc = run_field((dir + "/ref_field_sdf.bin").c_str(), 3, 16, 2, 10, 1.3819129f, 2, 1, 2, 512);
		return rc;
	} catch (const std::exception& e) {
		fprintf(stderr, "ref_tcnn failed: %s\n", e.what());
		return 1;
	}
}
