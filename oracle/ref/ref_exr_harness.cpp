// ref_exr_harness.cpp — writes and decodes OpenEXR files with the REFERENCE's own codec (tinyexr, dependencies/tinyexr/tinyexr.h),
// on the CPU.
//
// Test infrastructure.  Compiled by oracle/ref/Makefile straight from the header under /root/reference (nothing is copied into
// this repo) into oracle/_ref/ref_exr.  The reference reads EXR through `load_exr` -> `LoadEXRFromMemory` and writes it through
// `save_exr` -> `SaveEXRImageToMemory` with channels stored in (A)BGR order (src/tinyexr_wrapper.cu:45-143); the two modes here
// make the same library calls:
//
//   ref_exr encode <out.exr> <width> <height> <n_channels 1|3|4> <half 0|1> <compression 0 none|1 rle|2 zips|3 zip>
//       pixels are the closed-form pattern() below
//   ref_exr decode <in.exr> <out.bin>
//       out.bin = int32 width, int32 height, float RGBA as LoadEXRFromMemory returns it
#define TINYEXR_IMPLEMENTATION
#include <tinyexr/tinyexr.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

static float pattern(int x, int y, int c) { return (float)((x * 7 + y * 13 + c * 29) % 97) / 32.0f - 0.75f + (c == 3 ? 1.0f : 0.0f); }

static int encode(const char* path, int width, int height, int n_channels, bool half, int compression) {
	EXRHeader header;
	InitEXRHeader(&header);
	EXRImage image;
	InitEXRImage(&image);
	image.num_channels = n_channels;
	std::vector<std::vector<float>> images(n_channels);
	std::vector<float*> image_ptr(n_channels);
	for (int c = 0; c < n_channels; ++c) {
		images[c].resize((size_t)width * height);
		for (int y = 0; y < height; ++y)
			for (int x = 0; x < width; ++x) images[c][(size_t)y * width + x] = pattern(x, y, c);
	}
	for (int i = 0; i < n_channels; ++i) image_ptr[i] = images[n_channels - i - 1].data();
	image.images = (unsigned char**)image_ptr.data();
	image.width = width;
	image.height = height;
	header.num_channels = n_channels;
	header.channels = (EXRChannelInfo*)malloc(sizeof(EXRChannelInfo) * header.num_channels);
	const char* channel_names[] = {"R", "G", "B", "A"};
	for (int i = 0; i < n_channels; ++i) {
		memset(header.channels[i].name, 0, sizeof(header.channels[i].name));
		strncpy(header.channels[i].name, n_channels == 1 ? "Y" : channel_names[n_channels - i - 1], 255);
	}
	header.pixel_types = (int*)malloc(sizeof(int) * header.num_channels);
	header.requested_pixel_types = (int*)malloc(sizeof(int) * header.num_channels);
	for (int i = 0; i < header.num_channels; i++) {
		header.pixel_types[i] = TINYEXR_PIXELTYPE_FLOAT;
		header.requested_pixel_types[i] = half ? TINYEXR_PIXELTYPE_HALF : TINYEXR_PIXELTYPE_FLOAT;
	}
	header.compression_type = compression;
	const char* err = nullptr;
	unsigned char* buffer = nullptr;
	size_t n_bytes = SaveEXRImageToMemory(&image, &header, &buffer, &err);
	if (n_bytes == 0) {
		fprintf(stderr, "Failed to save EXR image: %s\n", err ? err : "?");
		return 1;
	}
	std::ofstream f{path, std::ios::out | std::ios::binary};
	f.write((char*)buffer, (std::streamsize)n_bytes);
	free(header.channels);
	free(header.pixel_types);
	free(header.requested_pixel_types);
	free(buffer);
	return 0;
}

static int decode(const char* in, const char* out) {
	std::vector<unsigned char> buffer;
	{
		std::ifstream f{in, std::ios::in | std::ios::binary | std::ios::ate};
		if (!f.good()) {
			fprintf(stderr, "Failed to open EXR file\n");
			return 2;
		}
		size_t size = (size_t)f.tellg();
		f.seekg(0, std::ios::beg);
		buffer.resize(size);
		f.read((char*)buffer.data(), (std::streamsize)size);
	}
	float* data = nullptr;
	int width = 0, height = 0;
	const char* err = nullptr;
	int ret = LoadEXRFromMemory(&data, &width, &height, buffer.data(), buffer.size(), &err);
	if (ret != TINYEXR_SUCCESS) {
		fprintf(stderr, "Failed to load EXR image: %s\n", err ? err : "?");
		return 1;
	}
	std::ofstream f{out, std::ios::out | std::ios::binary};
	f.write((const char*)&width, 4);
	f.write((const char*)&height, 4);
	f.write((const char*)data, (std::streamsize)((size_t)width * height * 4 * sizeof(float)));
	free(data);
	return 0;
}

int main(int argc, char** argv) {
	if (argc == 8 && std::string(argv[1]) == "encode") return encode(argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]) != 0, atoi(argv[7]));
	if (argc == 4 && std::string(argv[1]) == "decode") return decode(argv[2], argv[3]);
	fprintf(stderr, "usage: ref_exr encode <out.exr> <w> <h> <n_channels> <half> <compression> | decode <in.exr> <out.bin>\n");
	return 64;
}
