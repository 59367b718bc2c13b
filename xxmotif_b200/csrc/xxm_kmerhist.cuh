// the counting loop of NullModel::init (NullModel.cpp:122-132) and the mono-nucleotide counts of
// checkSequenceSet (backgroundDistribution.cpp:50-95) over the packed sequence stream that is already resident.
//
// The reference increments, for every position l and every k = 0..order while the (k+1)-mer starting at l stays inside the
// sequence and holds no N, the counter of that (k+1)-mer: order + 1 increments per position.  Here every position adds ONE
// to a histogram keyed by its LONGEST valid k-mer (length m = min(order + 1, distance to the next N or to the end of the
// sequence -- the >= 32 padding positions after every sequence are flagged N, so the end needs no extra test));
// the counts of the shorter k-mers are prefix marginals of that histogram, summed on the host (4^(order+1) additions).
// Almost all positions have m = order + 1: their bins live in shared memory (lane-replicated and conflict free up to
// 5-mers, plain up to 48 k bins; longer k-mers take several passes over the stream, each pass owning a bin range).  The rare
// shorter ones (sequence ends, N runs) go to global memory with atomics, in pass 0 only.
// k-mer value: base i of the k-mer in bits [2i+1 : 2i] (A=0 .. T=3), i.e. the FIRST base is least significant.
#pragma once
#include "xxm_device.cuh"

namespace xxm {

struct KmerHistIO {
	const uint64_t* codes;   // [nwords + 2]
	const uint32_t* nmask;   // [nwords + 2]
	uint32_t nwords;
	int K;                   // order + 1 (1..10)
	uint32_t binLo, binCount;  // K-mers [binLo, binLo + binCount) are counted by this pass
	int laneRep;             // 1: bins replicated per lane (binCount * 32 words of shared memory, shared-memory atomics between warps)
	                         // 2: ... and per warp (binCount * 32 * warps words): plain read-modify-write, no atomics at all
	int countShort;          // this pass also counts the positions with m < K
	unsigned long long* full;    // [4^K]
	unsigned long long* shortH;  // H[m] at offset (4^m - 4) / 3, m = 1..K-1
};

constexpr int kKmerHistThreads = 512;

__global__ void __launch_bounds__(kKmerHistThreads) k_kmer_hist(KmerHistIO io) {
	extern __shared__ uint32_t s_bins[];
	const uint32_t nb = io.laneRep == 2 ? io.binCount * 32u * (kKmerHistThreads / 32) : io.laneRep ? io.binCount * 32u : io.binCount;
	for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) s_bins[i] = 0;
	__syncthreads();
	const int K = io.K;
	const uint32_t kmask = (K >= 16) ? 0xFFFFFFFFu : ((1u << (2 * K)) - 1u);
	const uint32_t lane = threadIdx.x & 31u;
	uint32_t* const s_mine = s_bins + (io.laneRep == 2 ? (threadIdx.x >> 5) * io.binCount * 32u : 0u) + lane;  // this lane's column of bins
	const uint64_t needClear = (K > 1) ? ((1ull << (31 + K)) - 1ull) : 0xFFFFFFFFull;  // N flags that matter to the word's 32 positions
	for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < io.nwords; w += gridDim.x * blockDim.x) {
		const uint64_t lo = io.codes[w], hi = io.codes[w + 1];
		const uint64_t nm = (static_cast<uint64_t>(io.nmask[w + 1]) << 32) | io.nmask[w];
		const bool clean = (nm & needClear) == 0;
#pragma unroll 4
		for (int j = 0; j < 32; ++j) {
			const uint64_t c = (j == 0) ? lo : ((lo >> (2 * j)) | (hi << (64 - 2 * j)));
			uint32_t kmer = static_cast<uint32_t>(c) & kmask;
			int m = K;
			if (!clean) {
				const uint64_t r = nm >> j;
				if (r & 1ull) continue;  // N or padding: the reference counts nothing that starts here
				m = min(K, __ffsll(static_cast<long long>(r | (1ull << 40))) - 1);
				kmer &= (1u << (2 * m)) - 1u;
			}
			if (m == K) {
				const uint32_t b = kmer - io.binLo;
				if (b < io.binCount) {
					if (io.laneRep == 2) s_mine[b * 32u] += 1u;  // nobody else touches this word
					else atomicAdd(io.laneRep ? s_mine + b * 32u : s_bins + b, 1u);
				}
			} else if (io.countShort) {
				atomicAdd(io.shortH + ((1u << (2 * m)) - 4u) / 3u + kmer, 1ull);
			}
		}
	}
	__syncthreads();
	if (io.laneRep) {
		// bin b = sum over its 32 lane copies: one warp per bin, bins round robin over the warps
		const uint32_t wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
		for (uint32_t b = wid; b < io.binCount; b += nw) {
			uint32_t v = 0;
			if (io.laneRep == 2) for (uint32_t k = 0; k < nw; ++k) v += s_bins[(k * io.binCount + b) * 32u + lane];
			else v = s_bins[b * 32u + lane];
			v = __reduce_add_sync(0xFFFFFFFFu, v);
			if (lane == 0 && v) atomicAdd(io.full + io.binLo + b, static_cast<unsigned long long>(v));
		}
	} else {
		for (uint32_t b = threadIdx.x; b < io.binCount; b += blockDim.x)
			if (s_bins[b]) atomicAdd(io.full + io.binLo + b, static_cast<unsigned long long>(s_bins[b]));
	}
}

}  // namespace xxm
