// The exchange vector of a sharded optimizeMotif step (xxm_gpu_part5.inc): header + value-index histogram, laid out on the device.
// (In a header of its own so that tests/simt can run it from source; the library includes this file where the kernel used to stand.)
#pragma once
#include <cstdint>

namespace xxm {
constexpr uint32_t kXchgHist = 2048;  // value indices a sharded step can exchange (the reference tables have 20-470 score classes per 8-column part); 16 KB per collective
constexpr uint32_t kXchgHead = 10;    // poison, sum nval, sum nval^2, sum hash, sum hash^2, nbSites, posSetSize, totalPositions, overflowed ranks, spare
struct XchgHead { unsigned long long v[kXchgHead]; };
// The exchange vector has the same length on every rank no matter what (ranks that disagreed about the length of a
// collective would hang): header + kXchgHist histogram slots.  Slot 0 (poison) and slot 8 (ranks whose site list outgrew its
// arrays) carry what only one rank can know, so that after the sum every rank takes the same branch.
// keptDev != nullptr: the number of kept sites (header slot 5) is still on the device; totDev != nullptr: so is the number of
// candidates the scan found, to be compared with the capacity of the arrays
__global__ void k_pack_hist(const uint32_t* __restrict__ hist, uint32_t nval, XchgHead head, const uint32_t* __restrict__ keptDev,
                            const unsigned long long* __restrict__ totDev, unsigned long long cap, unsigned long long* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < kXchgHead) {
		unsigned long long v = head.v[i];
		if (i == 5 && keptDev) v = static_cast<unsigned long long>(*keptDev);
		if (i == 8 && totDev) v = (*totDev > cap) ? 1ull : 0ull;
		out[i] = v;
	}
	if (i < kXchgHist) out[kXchgHead + i] = (i < nval) ? hist[i] : 0ull;
}
}  // namespace xxm
