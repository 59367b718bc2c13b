// a15 on the device: the start-position histogram behind Motif::updateEnrichedRegion (xxm_gpu_part4.inc).
// (In a header of its own so that tests/simt can run it from source; the library includes this file where the kernel used to stand.)
#pragma once
#include "xxm_device.cuh"

namespace xxm {
// start-position histogram of the candidates with float(pValScore) < 1e-3 inside the search window
// (Motif::updateEnrichedRegion, Motif.cpp:989-999); positions on the reverse strand are folded back
// (bins in shared memory when they fit -- dynamic, nbins * 4 bytes -- and one count per block: 1.3e7 candidates adding to a few
// hundred global addresses and to ONE global counter took 680 us)
__global__ void __launch_bounds__(256) k_region_hist(const uint32_t* __restrict__ pos, const float* __restrict__ score, uint32_t n, int revcomp, int max_leng,
                                                     int motifLength, int lo, int hi, unsigned int* __restrict__ counts, unsigned int* __restrict__ size,
                                                     uint32_t nbinsShared) {
	extern __shared__ unsigned int s_rh[];
	__shared__ unsigned int s_n;
	for (uint32_t k = threadIdx.x; k < nbinsShared; k += blockDim.x) s_rh[k] = 0;
	if (threadIdx.x == 0) s_n = 0;
	__syncthreads();
	unsigned int mine = 0;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		if (!(static_cast<double>(score[i]) < 1e-3)) continue;
		int p = static_cast<int>(pos[i]);
		if (revcomp && p > max_leng / 2) p = max_leng - (p + motifLength - 1);
		if (p >= lo && p <= hi && p >= 0 && p <= max_leng) {
			if (static_cast<uint32_t>(p) < nbinsShared) atomicAdd(&s_rh[p], 1u);
			else atomicAdd(&counts[p], 1u);
			++mine;
		}
	}
	mine = __reduce_add_sync(0xFFFFFFFFu, mine);
	if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_n, mine);
	__syncthreads();
	for (uint32_t k = threadIdx.x; k < nbinsShared; k += blockDim.x)
		if (s_rh[k]) atomicAdd(&counts[k], s_rh[k]);
	if (threadIdx.x == 0 && s_n) atomicAdd(size, s_n);
}
}  // namespace xxm
