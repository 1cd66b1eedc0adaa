// host_util.h — host-side pieces shared by the NeRF Testbed (testbed.cu) and the image / SDF field Testbed (field.cu):
// device buffers, the optimizer-chain / loss parsers of Testbed::reset_network (src/testbed.cu:4160-4412) and the
// extern "C" error convention.
#pragma once

#include <fstream>
#include <sstream>
#include <string>

#include "common.cuh"
#include "json_mini.h"

namespace ngpb {

void set_last_error(const std::string& msg);
float reduce_sum_f32(cudaStream_t stream, const float* data, uint32_t n, float* scratch_dev);
void optimizer_step_flat(uint32_t n_matrix_params, uint32_t n_params, cudaStream_t stream, const ngp_adam_cfg& cfg, float* params_fp32, __half* params_fp16,
	__half* params_ema, __half* grads, float* m1, float* m2, uint32_t* steps);

inline void require_device() {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		n = 0;
	}
	NGPB_CHECK(n > 0, "no CUDA device: libngp_b200 has no CPU fallback");
}

template <typename T>
struct DevBuf {
	T* p = nullptr;
	size_t n = 0;
	~DevBuf() { release(); }
	void release() {
		if (p) cudaFree(p);
		p = nullptr;
		n = 0;
	}
	void ensure(size_t count) {
		if (count <= n) return;
		release();
		NGPB_CUDA_CHECK(cudaMalloc(&p, count * sizeof(T)));
		n = count;
	}
	void ensure_zeroed(size_t count) {
		const bool fresh = count > n;
		ensure(count);
		if (fresh) NGPB_CUDA_CHECK(cudaMemset(p, 0, n * sizeof(T)));
	}
};

struct OptimizerConfig {
	float learning_rate = 1e-2f, beta1 = 0.9f, beta2 = 0.99f, epsilon = 1e-15f, l2_reg = 1e-6f;
	float ema_decay = 0.95f;
	bool has_ema = true;
	bool has_decay = true;
	uint32_t decay_start = 20000, decay_interval = 10000, decay_end = 10000000;
	float decay_base = 0.33f;
};

// optimizer: Ema{ExponentialDecay{Adam}} or any suffix of that chain (create_optimizer, tiny-cuda-nn/src/optimizer.cu)
inline OptimizerConfig parse_optimizer_chain(const Json& optimizer) {
	OptimizerConfig opt;
	opt.has_ema = false;
	opt.has_decay = false;
	opt.learning_rate = 1e-3f; opt.beta2 = 0.999f; opt.epsilon = 1e-8f; opt.l2_reg = 1e-8f;  // adam.h defaults
	const Json* o = &optimizer;
	for (;;) {
		const std::string ot = to_lower(o->value("otype", std::string("Adam")));
		if (ot == "ema") {
			opt.has_ema = true;
			opt.ema_decay = (float)o->value("decay", 0.99);
			NGPB_CHECK(!(o->contains("full_precision") && o->sub("full_precision").b), "Ema.full_precision is not supported");
		} else if (ot == "exponentialdecay") {
			opt.has_decay = true;
			opt.decay_base = (float)o->value("decay_base", 0.1);
			opt.decay_interval = (uint32_t)o->value("decay_interval", 10000.0);
			opt.decay_start = (uint32_t)o->value("decay_start", 10000.0);
			opt.decay_end = (uint32_t)o->value("decay_end", 10000000.0);
		} else if (ot == "adam") {
			opt.learning_rate = (float)o->value("learning_rate", 1e-3);
			opt.beta1 = (float)o->value("beta1", 0.9);
			opt.beta2 = (float)o->value("beta2", 0.999);
			opt.epsilon = (float)o->value("epsilon", 1e-8);
			opt.l2_reg = (float)o->value("l2_reg", 1e-8);
			// options of the reference's Adam (adam.h:133-170, 258-300) that k_adam_ema does not implement: refuse them instead of training differently
			auto refuse = [&](const char* key, double dflt) {
				NGPB_CHECK(!o->contains(key) || o->value(key, dflt) == dflt, std::string("optimizer: Adam option '") + key + "' is not implemented by ngp_b200 (only its default is accepted)");
			};
			refuse("relative_decay", 0.0);
			refuse("absolute_decay", 0.0);
			refuse("adabound", 0.0);
			refuse("clipping_magnitude", 0.0);
			refuse("non_matrix_learning_rate_factor", 1.0);
			break;
		} else {
			NGPB_CHECK(false, "optimizer.otype '" + ot + "' is not supported (Ema / ExponentialDecay / Adam)");
		}
		NGPB_CHECK(o->contains("nested"), "optimizer: missing nested");
		o = &o->sub("nested");
	}
	return opt;
}

inline uint32_t parse_loss_type(const Json& loss) {
	const std::string lt = to_lower(loss.value("otype", std::string("L2")));
	if (lt == "l2") return NGP_LOSS_L2;
	if (lt == "l1") return NGP_LOSS_L1;
	if (lt == "mape") return NGP_LOSS_MAPE;
	if (lt == "smape") return NGP_LOSS_SMAPE;
	if (lt == "huber") return NGP_LOSS_HUBER;
	if (lt == "logl1") return NGP_LOSS_LOGL1;
	if (lt == "relativel2") return NGP_LOSS_RELATIVE_L2;
	NGPB_CHECK(false, "loss.otype '" + lt + "' is not supported");
	return 0;
}

// ExponentialDecayOptimizer::step (exponential_decay.h:60-72) + the Adam/EMA hyper-parameters of one optimizer step.
// `optimizer_step` is the 0-based count of steps taken so far; it is incremented.
inline ngp_adam_cfg next_adam_cfg(const OptimizerConfig& opt, uint32_t& optimizer_step, float& lr_factor, float loss_scale, bool train_matrix, bool train_non_matrix) {
	if (optimizer_step == 0) lr_factor = 1.0f;
	if (opt.has_decay && optimizer_step >= opt.decay_start && (optimizer_step - opt.decay_start) % opt.decay_interval == 0 && optimizer_step <= opt.decay_end) {
		lr_factor *= opt.decay_base;
	}
	++optimizer_step;
	ngp_adam_cfg a{};
	a.learning_rate = opt.learning_rate * lr_factor;
	a.beta1 = opt.beta1;
	a.beta2 = opt.beta2;
	a.epsilon = opt.epsilon;
	a.l2_reg = opt.l2_reg;
	a.loss_scale = loss_scale;
	a.ema_decay = opt.has_ema ? opt.ema_decay : 0.0f;
	a.ema_step = optimizer_step;
	a.optimize_matrix_params = train_matrix;
	a.optimize_non_matrix_params = train_non_matrix;
	return a;
}

// ---- network config files: Testbed::load_network_config for ".json" + merge_parent_network_config (src/testbed.cu:86-97, 280-309).
// A config may name a "parent" file (relative to its own directory); the parent is loaded first, recursively, and the child is applied
// on top as an RFC 7386 merge patch (nlohmann::json::merge_patch): objects merge key by key, null deletes, everything else replaces.
inline void json_merge_patch(Json& target, const Json& patch) {
	if (patch.type != Json::Object) {
		target = patch;
		return;
	}
	if (target.type != Json::Object) target = jobj();
	for (const auto& kv : patch.obj) {
		if (kv.second.type == Json::Null) {
			target.obj.erase(kv.first);
		} else {
			json_merge_patch(target.obj[kv.first], kv.second);
		}
	}
}
inline Json load_network_config_file(const std::string& path, int depth = 0) {
	NGPB_CHECK(depth < 16, "network config: 'parent' chain too deep (cycle?)");
	std::ifstream f(path);
	NGPB_CHECK(f.good(), std::string("Network config '") + path + "' does not exist.");
	std::stringstream ss;
	ss << f.rdbuf();
	const std::string text = ss.str();
	Json child = JsonParser(text).parse();
	if (child.type != Json::Object || !child.contains("parent")) return child;
	NGPB_CHECK(child.at("parent").type == Json::String, "network config: 'parent' must be a file name");
	const size_t slash = path.find_last_of('/');
	const std::string dir = slash == std::string::npos ? std::string() : path.substr(0, slash + 1);
	Json parent = load_network_config_file(dir + child.at("parent").str, depth + 1);
	json_merge_patch(parent, child);
	return parent;
}

}  // namespace ngpb

#define NGPB_TRY(...)                      \
	try {                                    \
		__VA_ARGS__;                           \
		return 0;                              \
	} catch (const std::exception& e) {      \
		ngpb::set_last_error(e.what());        \
		return 1;                              \
	}
