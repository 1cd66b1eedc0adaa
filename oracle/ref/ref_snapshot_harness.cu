// ref_snapshot_harness.cu — writes and reads snapshot containers with the REFERENCE's own serializer stack, on the CPU.
//
// Test infrastructure.  Compiled by oracle/ref/Makefile straight from the headers under /root/reference (nothing is copied
// into this repo) into oracle/_ref/ref_snapshot.  The application itself (Testbed::save_snapshot, src/testbed.cu:5288-5355)
// cannot be built under the task rules, but everything a snapshot's bytes depend on can: nlohmann::json's to_msgpack /
// from_msgpack (dependencies/tiny-cuda-nn/dependencies/json/json.hpp), zstr's gzip streams (dependencies/zstr/src/zstr.hpp),
// tcnn's vector / matrix converters (tiny-cuda-nn/vec_json.h) and the application's own to_json / from_json for BoundingBox,
// Lens, TrainingXForm and NerfDataset (neural-graphics-primitives/json_binding.h).  This program assembles the same keys in
// the same order of assignments as save_snapshot and Trainer::serialize (trainer.h:442-455) through those converters:
//
//   ref_snapshot write <out.ingp|.msgpack> <network_config.json> <n_params> <n_cascades> <compress 0|1>
//       parameters and density grid are closed-form patterns (param_pattern / grid_pattern below) so a test can predict them
//   ref_snapshot dump <in.ingp|.msgpack>
//       reads like Testbed::load_network_config (src/testbed.cu:280-309), passes nerf.dataset, aabb and render_aabb through
//       the reference's from_json -> to_json (so a file the reference's parser rejects fails here) and prints JSON text with
//       every binary value replaced by {"bytes": n, "fnv1a64": "<hex>", "wsum64": "<decimal>"}
//   ref_snapshot pack <in.json> <out.msgpack>
//       nlohmann::json::parse -> to_msgpack: the byte-level known answers for the product's MessagePack writer
#include <neural-graphics-primitives/json_binding.h>

#include <zstr.hpp>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

using namespace ngp;
using json = nlohmann::json;

// tcnn's logging hook lives in common_host.cu, which needs the CUDA driver library; this CPU-only program provides the one
// symbol the headers reference (used by GPUMemory's destructor diagnostics only).
namespace tcnn {
static std::function<void(LogSeverity, const std::string&)> g_cb = [](LogSeverity, const std::string& msg) { fprintf(stderr, "%s\n", msg.c_str()); };
const std::function<void(LogSeverity, const std::string&)>& log_callback() { return g_cb; }
}

static float param_pattern(size_t i) { return ((float)((i * 37u) % 1001u) - 500.0f) / 4000.0f; }
static float grid_pattern(size_t i) { return (i % 7u == 0u) ? 1.5f : 0.001f * (float)(i % 5u); }

static uint64_t fnv1a64(const uint8_t* p, size_t n) {
	uint64_t h = 1469598103934665603ull;
	for (size_t i = 0; i < n; ++i) {
		h ^= p[i];
		h *= 1099511628211ull;
	}
	return h;
}

static bool has_ext(const std::string& p, const char* ext) {
	const size_t n = strlen(ext);
	return p.size() >= n && p.compare(p.size() - n, n, ext) == 0;
}

static json::binary_t half_binary(size_t n, float (*f)(size_t)) {
	json::binary_t b;
	b.resize(n * sizeof(__half));
	__half* h = reinterpret_cast<__half*>(b.data());
	for (size_t i = 0; i < n; ++i) h[i] = __float2half_rn(f(i));
	return b;
}

static int do_write(const std::string& path, const std::string& config_path, size_t n_params, uint32_t n_cascades, bool compress) {
	std::ifstream cf{config_path};
	json config = json::parse(cf, nullptr, true, true);

	// Trainer::serialize (trainer.h:442-455), PARAMS_T = __half
	json data;
	data["n_params"] = n_params;
	data["params_type"] = "__half";
	data["params_binary"] = half_binary(n_params, param_pattern);
	config["snapshot"] = data;

	auto& snapshot = config["snapshot"];
	snapshot["version"] = (size_t)1;
	snapshot["mode"] = "nerf";
	snapshot["density_grid_size"] = (uint32_t)128;
	snapshot["density_grid_binary"] = half_binary((size_t)128 * 128 * 128 * n_cascades, grid_pattern);

	NerfDataset dataset;
	dataset.n_images = 3;
	dataset.paths = {"images/0001.png", "images/0002.png", ""};
	dataset.metadata.resize(3);
	dataset.xforms.resize(3);
	for (uint32_t i = 0; i < 3; ++i) {
		auto& m = dataset.metadata[i];
		m.resolution = ivec2(640 + 16 * (int)i, 480);
		m.focal_length = vec2(500.25f + (float)i, 501.5f);
		m.principal_point = vec2(0.5f + 0.01f * (float)i, 0.49f);
		m.rolling_shutter = vec4(0.0f);
		mat4x3 x = mat4x3::identity();
		x[0] = vec3(0.0f, 0.6f, 0.8f);
		x[1] = vec3(1.0f, 0.0f, 0.0f);
		x[2] = vec3(0.0f, 0.8f, -0.6f);
		x[3] = vec3(0.5f + 0.25f * (float)i, 0.5f, 0.125f * (float)i);
		dataset.xforms[i].start = x;
		dataset.xforms[i].end = x;
	}
	dataset.metadata[1].lens.mode = ELensMode::OpenCV;
	dataset.metadata[1].lens.params[0] = 0.0625f;
	dataset.metadata[1].lens.params[1] = -0.03125f;
	dataset.metadata[1].lens.params[2] = 0.001953125f;
	dataset.metadata[1].lens.params[3] = -0.0009765625f;
	dataset.aabb_scale = 1 << (n_cascades - 1);
	dataset.render_aabb = BoundingBox{vec3(0.5f - 0.5f * (float)dataset.aabb_scale), vec3(0.5f + 0.5f * (float)dataset.aabb_scale)};
	dataset.scale = 0.33f;
	dataset.offset = vec3(0.5f, 0.5f, 0.5f);

	snapshot["nerf"]["aabb_scale"] = dataset.aabb_scale;
	snapshot["training_step"] = (uint32_t)1234;
	snapshot["loss"] = 0.00123f;
	snapshot["aabb"] = dataset.render_aabb;
	snapshot["bounding_radius"] = 1.0f;
	snapshot["render_aabb_to_local"] = mat3::identity();
	snapshot["render_aabb"] = dataset.render_aabb;
	snapshot["up_dir"] = vec3(0.0f, 1.0f, 0.0f);
	snapshot["sun_dir"] = normalize(vec3(1.0f));
	snapshot["exposure"] = 0.0f;
	snapshot["background_color"] = vec4(0.0f, 0.0f, 0.0f, 1.0f);
	snapshot["camera"]["matrix"] = dataset.xforms[0].start;
	snapshot["camera"]["fov_axis"] = 1;
	snapshot["camera"]["relative_focal_length"] = vec2(1.0f);
	snapshot["camera"]["screen_center"] = vec2(0.5f);
	snapshot["camera"]["zoom"] = 1.0f;
	snapshot["camera"]["scale"] = 1.5f;
	snapshot["camera"]["aperture_size"] = 0.0f;
	snapshot["camera"]["autofocus"] = false;
	snapshot["camera"]["autofocus_target"] = vec3(0.5f);
	snapshot["camera"]["autofocus_depth"] = 1.0f;
	snapshot["nerf"]["rgb"]["rays_per_batch"] = (uint32_t)4096;
	snapshot["nerf"]["rgb"]["measured_batch_size"] = (uint32_t)250000;
	snapshot["nerf"]["rgb"]["measured_batch_size_before_compaction"] = (uint32_t)300000;
	snapshot["nerf"]["dataset"] = dataset;

	std::ofstream f{path, std::ios::out | std::ios::binary};
	if (has_ext(path, ".ingp")) {
		zstr::ostream zf{f, zstr::default_buff_size, compress ? Z_DEFAULT_COMPRESSION : Z_NO_COMPRESSION};
		json::to_msgpack(config, zf);
	} else {
		json::to_msgpack(config, f);
	}
	return 0;
}

static void strip_binaries(json& j) {
	if (j.is_binary()) {
		const auto& b = j.get_binary();
		char hex[32];
		snprintf(hex, sizeof(hex), "%016llx", (unsigned long long)fnv1a64(b.data(), b.size()));
		uint64_t wsum = 0;  // sum of byte[i] * (i + 1) mod 2^64: a second checksum a numpy test can compute over megabytes
		for (size_t i = 0; i < b.size(); ++i) wsum += (uint64_t)b[i] * (uint64_t)(i + 1);
		char dec[32];
		snprintf(dec, sizeof(dec), "%llu", (unsigned long long)wsum);
		j = json{{"bytes", b.size()}, {"fnv1a64", hex}, {"wsum64", dec}};
	} else if (j.is_object() || j.is_array()) {
		for (auto& e : j) strip_binaries(e);
	}
}

static int do_dump(const std::string& path) {
	std::ifstream f{path, std::ios::in | std::ios::binary};
	if (!f.good()) {
		fprintf(stderr, "cannot open %s\n", path.c_str());
		return 2;
	}
	json config;
	if (has_ext(path, ".ingp")) {
		zstr::istream zf{f};
		config = json::from_msgpack(zf);
	} else {
		config = json::from_msgpack(f);
	}
	if (!config.contains("snapshot")) {
		fprintf(stderr, "File '%s' does not contain a snapshot.\n", path.c_str());
		return 3;
	}
	json& snapshot = config["snapshot"];
	// the typed reads Testbed::load_snapshot performs (src/testbed.cu:5357-5460) — any of them throws on a malformed file
	if (snapshot.value("version", 0) < 1) throw std::runtime_error{"Snapshot uses an old format and can not be loaded."};
	if (snapshot["density_grid_size"] != 128u) throw std::runtime_error{"Incompatible grid size."};
	BoundingBox aabb = snapshot.value("aabb", BoundingBox{});
	BoundingBox render_aabb = snapshot.value("render_aabb", BoundingBox{});
	vec4 background = snapshot.value("background_color", vec4(0.0f));
	(void)background;
	uint32_t rays_per_batch = snapshot["nerf"]["rgb"]["rays_per_batch"];
	uint32_t measured = snapshot["nerf"]["rgb"]["measured_batch_size"];
	uint32_t measured_before = snapshot["nerf"]["rgb"]["measured_batch_size_before_compaction"];
	(void)rays_per_batch; (void)measured; (void)measured_before;
	uint32_t training_step = snapshot["training_step"];
	float loss = snapshot["loss"];
	(void)training_step; (void)loss;
	if (!snapshot["params_binary"].is_binary() || !snapshot["density_grid_binary"].is_binary()) throw std::runtime_error{"binary values expected"};
	snapshot["aabb"] = aabb;
	snapshot["render_aabb"] = render_aabb;
	if (snapshot["nerf"].contains("dataset")) {
		NerfDataset dataset = snapshot["nerf"]["dataset"];
		snapshot["nerf"]["dataset"] = json{};
		snapshot["nerf"]["dataset"] = dataset;
	}
	strip_binaries(config);
	std::cout << config.dump(1) << std::endl;
	return 0;
}

int main(int argc, char** argv) {
	try {
		if (argc == 7 && std::string(argv[1]) == "write") {
			return do_write(argv[2], argv[3], (size_t)atoll(argv[4]), (uint32_t)atoi(argv[5]), atoi(argv[6]) != 0);
		}
		if (argc == 3 && std::string(argv[1]) == "dump") return do_dump(argv[2]);
		if (argc == 4 && std::string(argv[1]) == "pack") {
			std::ifstream in{argv[2]};
			json j = json::parse(in, nullptr, true, true);
			std::ofstream out{argv[3], std::ios::out | std::ios::binary};
			json::to_msgpack(j, out);
			return 0;
		}
	} catch (const std::exception& e) {
		fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
	fprintf(stderr, "usage: ref_snapshot write <out> <config.json> <n_params> <n_cascades> <compress> | dump <in> | pack <in.json> <out.msgpack>\n");
	return 64;
}
