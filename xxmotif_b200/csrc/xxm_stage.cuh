// The scatter half of the batched table staging of xxm_reinitialize (xxm_gpu.cu): one pinned buffer, one copy, one kernel that
// distributes the segments to their device arrays.
// (In a header of its own so that tests/simt can run it from source; the library includes this file where the kernel used to stand.)
#pragma once
#include <cstdint>

namespace xxm {
constexpr int kMaxUpSegs = 16;
struct UpSegs { void* dst[kMaxUpSegs]; uint32_t off[kMaxUpSegs], bytes[kMaxUpSegs]; };
// blockIdx.y = segment; bytes are copied one by one (the segments are a few KB)
__global__ void k_scatter_stage(const uint8_t* __restrict__ stage, UpSegs segs) {
	const int sgm = blockIdx.y;
	uint8_t* dst = static_cast<uint8_t*>(segs.dst[sgm]);
	const uint8_t* src = stage + segs.off[sgm];
	const uint32_t n = segs.bytes[sgm];
	if (((reinterpret_cast<uintptr_t>(dst) | segs.off[sgm]) & 3u) == 0) {
		const uint32_t n4 = n >> 2;
		for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x)
			reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
		for (uint32_t i = (n4 << 2) + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
	} else {
		for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
	}
}
}  // namespace xxm
