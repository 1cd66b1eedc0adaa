// ref_loader_harness.cu — runs the REFERENCE's own dataset loader (ngp::load_nerf, src/nerf_loader.cu:271-735) on scenes on disk
// and prints what it produced.
//
// Test infrastructure.  oracle/ref/Makefile compiles src/nerf_loader.cu, src/common_host.cu, src/tinyexr_wrapper.cu and
// src/thread_pool.cpp where they lie under /root/reference (nothing is copied into this repo) and links them with this file
// into oracle/_ref/ref_loader.  load_nerf uploads every image to the device (NerfDataset::set_training_image), so the program
// needs a GPU: tools/make_ref_loader_golden.sh runs it on the GPU box over the scenes tests/loader_scenes.py writes, and the
// JSON it prints is committed as tests/golden/ref_loader.json.  tests/test_nerf_loader.py then checks the product's loader
// (instant-ngp_b200/nerf_loader.py) against it on the same scenes, regenerated on the CPU.
//
//   ref_loader <out.json> <scene dir or transforms.json> [more scenes...]
//
// Per scene: to_json(NerfDataset) (json_binding.h:112-136: paths, per-image focal length / lens / principal point / rolling
// shutter / resolution, xforms, render_aabb, up, offset, scale, aabb_scale, from_mitsuba, is_hdr, ...) plus, per image, the
// stored pixel type and a checksum of the stored pixel bytes (sum of byte[i] * (i + 1) mod 2^64).
#include <neural-graphics-primitives/json_binding.h>
#include <neural-graphics-primitives/nerf_loader.h>

#include <filesystem/directory.h>

#include <algorithm>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

using namespace ngp;
using json = nlohmann::json;

static std::string wsum64(const std::vector<uint8_t>& b) {
	uint64_t s = 0;
	for (size_t i = 0; i < b.size(); ++i) s += (uint64_t)b[i] * (uint64_t)(i + 1);
	return std::to_string((unsigned long long)s);
}

static const char* type_name(EImageDataType t) {
	switch (t) {
		case EImageDataType::None: return "None";
		case EImageDataType::Byte: return "Byte";
		case EImageDataType::Half: return "Half";
		case EImageDataType::Float: return "Float";
	}
	return "?";
}

int main(int argc, char** argv) {
	if (argc < 3) {
		fprintf(stderr, "usage: ref_loader <out.json> <scene dir or transforms.json> [more scenes...]\n");
		return 64;
	}
	json out = json::object();
	for (int a = 2; a < argc; ++a) {
		const std::string scene = argv[a];
		json entry;
		try {
			fs::path p{scene};
			std::vector<fs::path> jsonpaths;
			if (p.is_directory()) {
				// Testbed::load_nerf(data_path): every *.json in the directory (src/testbed_nerf.cu:2445-2462)
				for (const auto& q : fs::directory{p}) {
					if (q.is_file() && equals_case_insensitive(q.extension(), "json")) jsonpaths.emplace_back(q);
				}
				// readdir order is arbitrary; the files are read in name order here so that the golden is reproducible
				std::sort(jsonpaths.begin(), jsonpaths.end(), [](const fs::path& x, const fs::path& y) { return x.str() < y.str(); });
			} else {
				jsonpaths.emplace_back(p);
			}
			NerfDataset ds = load_nerf(jsonpaths, 0.0f);
			entry = ds;
			entry["has_rays"] = ds.has_rays;
			entry["has_light_dirs"] = ds.has_light_dirs;
			entry["sharpness_resolution"] = ds.sharpness_resolution;
			json px = json::array();
			for (size_t i = 0; i < ds.n_images; ++i) {
				std::vector<uint8_t> host(ds.pixelmemory[i].size());
				if (!host.empty()) ds.pixelmemory[i].copy_to_host(host.data());
				json e;
				e["type"] = type_name(ds.metadata[i].image_data_type);
				e["bytes"] = host.size();
				e["wsum64"] = wsum64(host);
				json head = json::array();
				for (size_t k = 0; k < host.size() && k < 16; ++k) head.push_back((int)host[k]);
				e["head"] = head;
				e["has_depth"] = ds.metadata[i].depth != nullptr;
				px.push_back(e);
			}
			entry["pixels"] = px;
			// file names only: the scenes are regenerated under another root on the CPU side
			json names = json::array();
			for (const auto& s : ds.paths) names.push_back(fs::path{s}.filename());
			entry["paths"] = names;
		} catch (const std::exception& e) {
			entry = json{{"error", e.what()}};
		}
		std::string key = fs::path{scene}.filename();
		out[key] = entry;
	}
	std::ofstream f{argv[1]};
	f << out.dump(1) << std::endl;
	return 0;
}
