// NCCL, loaded at run time.  libxxmotif_gpu.so does not link libnccl: a host that already carries a copy (PyTorch does)
// must not end up with two, so the library first asks the loader for an already loaded libnccl.so.2 and only then opens
// the system one.  Only the handful of entry points the refinement step needs are bound.
#pragma once
#include <dlfcn.h>
#include <nccl.h>  // types and enums only

#include <string>

namespace xxm {

struct NcclApi {
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	void* handle = nullptr;

	bool load(std::string* err) {
		if (handle) return true;
		handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
		if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!handle) { *err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
		GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(handle, "ncclGetUniqueId"));
		CommInitRank = reinterpret_cast<decltype(CommInitRank)>(dlsym(handle, "ncclCommInitRank"));
		CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(handle, "ncclCommDestroy"));
		CommAbort = reinterpret_cast<decltype(CommAbort)>(dlsym(handle, "ncclCommAbort"));
		AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(handle, "ncclAllReduce"));
		AllGather = reinterpret_cast<decltype(AllGather)>(dlsym(handle, "ncclAllGather"));
		Broadcast = reinterpret_cast<decltype(Broadcast)>(dlsym(handle, "ncclBroadcast"));
		Send = reinterpret_cast<decltype(Send)>(dlsym(handle, "ncclSend"));
		Recv = reinterpret_cast<decltype(Recv)>(dlsym(handle, "ncclRecv"));
		GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(handle, "ncclGetErrorString"));
		if (!GetUniqueId || !CommInitRank || !CommDestroy || !CommAbort || !AllReduce || !AllGather || !Broadcast || !Send || !Recv || !GetErrorString) { *err = "libnccl.so.2 lacks an expected symbol"; handle = nullptr; return false; }
		return true;
	}
};

inline NcclApi& nccl_api() {
	static NcclApi api;
	return api;
}

}  // namespace xxm
