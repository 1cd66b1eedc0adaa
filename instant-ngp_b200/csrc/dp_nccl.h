// dp_nccl.h — the data-parallel exchange of Testbed::train, inside the library: NCCL all-reduce of the flat fp16 gradient buffer on the
// training stream and of the 16-byte counter block (+ the loss partials every 16th step) on a communication stream beside the
// forward/backward kernel.  The reference has no data parallelism; SURVEY §8e specifies this exchange.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 — inside a PyTorch process that is the copy torch already loaded, otherwise the
// system library): a single-GPU user of libngp_b200.so needs no NCCL at all.  Only the handful of entry points below is used; their
// prototypes and enum values are those of the public nccl.h (2.x ABI).
#pragma once

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace ngpb {

struct NcclApi {
	typedef struct ncclComm* comm_t;
	struct unique_id {
		char internal[128];
	};
	enum { ncclSuccess = 0 };
	enum { ncclUint32 = 3, ncclFloat16 = 6, ncclFloat32 = 7 };   // ncclDataType_t
	enum { ncclSum = 0 };                                          // ncclRedOp_t

	int (*GetUniqueId)(unique_id*) = nullptr;
	int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
	int (*CommDestroy)(comm_t) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
	int (*Broadcast)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
	int (*GetVersion)(int*) = nullptr;
	void* handle = nullptr;

	static NcclApi& get() {
		static NcclApi api;
		if (!api.handle) {
			const char* names[] = {"libnccl.so.2", "libnccl.so"};
			for (const char* n : names) {
				api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
				if (api.handle) break;
			}
			if (!api.handle) throw std::runtime_error(std::string("ngp_b200: data parallel training needs NCCL, and libnccl.so.2 could not be loaded: ") + dlerror());
			auto sym = [&](const char* s) {
				void* p = dlsym(api.handle, s);
				if (!p) throw std::runtime_error(std::string("ngp_b200: libnccl lacks ") + s);
				return p;
			};
			api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
			api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
			api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
			api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
			api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
			api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
			api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
		}
		return api;
	}
	void check(int rc, const char* what) const {
		if (rc != ncclSuccess) throw std::runtime_error(std::string("ngp_b200: ") + what + " failed: " + (GetErrorString ? GetErrorString(rc) : "NCCL error"));
	}
};

}  // namespace ngpb
