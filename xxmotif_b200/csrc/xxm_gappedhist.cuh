// SURVEY.md 8(f) row 4, first half: the counting loop of ElongCandidates::countAllOccurrences (elongationCandidates.cpp:234-262)
// over the packed sequence stream that is already resident.
//
// The reference walks every start position j of every sequence and, if none of the `coding` characters of the gapped k-mer
// (character l at j + l + gap[l-1]) is an N, increments the counter of that k-mer; gap positions may hold anything.  With a
// palindromic seed type under --revcomp only the first half of the start positions of the doubled sequence is walked.
// Here: one thread per word of 32 start positions, the coding characters gathered as GROUPS of adjacent characters (at most
// six shift-mask-or steps), one shared-memory histogram per CTA (4^5 or 4^6 bins), flushed with one atomic per non-empty bin.
// Windows running off the end of a sequence fall on the >= 32 padding positions, which are flagged N, and the last position
// of the span is always a coding character -- so the reference's per-sequence loop bound needs no test of its own.
// k-mer value on the device: character l in bits [2l+1 : 2l] (first character least significant); the host reorders.
#pragma once
#include "xxm_device.cuh"

namespace xxm {

struct GappedHistIO {
	uint32_t scanWords;     // words of the sequences that take part (sequences 1..nbSequences)
	int k;                  // coding characters (1..7)
	int span;               // offset of the last coding character + 1 (<= 32)
	int ngroups;
	int goff[8];            // base offset of the first character of group g
	int gshift[8];          // 2 * (index of that character in the k-mer)
	uint32_t gmask[8];      // (1 << 2 * characters of the group) - 1
	uint32_t codingMask;    // bit o set: a coding character sits at base offset o
	int firstHalfOnly;      // walk start positions 1 .. (n - span + 1) / 2 only (n = stored length)
	uint32_t* bins;         // [4^k]
};

constexpr int kGappedHistThreads = 256;

__global__ void __launch_bounds__(kGappedHistThreads) k_gapped_hist(SeqStream S, GappedHistIO io) {
	extern __shared__ uint32_t s_gbins[];
	const uint32_t nb = 1u << (2 * io.k);
	for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) s_gbins[i] = 0;
	__syncthreads();
	for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < io.scanWords; w += gridDim.x * blockDim.x) {
		const uint64_t nm = (static_cast<uint64_t>(S.nmask[w + 1]) << 32) | S.nmask[w];
		if (static_cast<uint32_t>(nm) == 0xFFFFFFFFu) continue;  // padding
		int jlim = 32;
		if (io.firstHalfOnly) {
			const uint32_t sq = find_seq(S, w);
			const int seqLen = (S.len[sq] - io.span + 1) / 2;
			jlim = max(0, min(32, seqLen - static_cast<int>(w - S.wordOff[sq]) * 32));
		}
		const uint64_t lo = S.codes[w], hi = S.codes[w + 1];
#pragma unroll 4
		for (int j = 0; j < 32; ++j) {
			if (j >= jlim) break;
			if (static_cast<uint32_t>(nm >> j) & io.codingMask) continue;
			const uint64_t c = (j == 0) ? lo : ((lo >> (2 * j)) | (hi << (64 - 2 * j)));
			uint32_t id = 0;
#pragma unroll
			for (int g = 0; g < 8; ++g)  // compile-time g: the descriptors stay in the constant bank
				if (g < io.ngroups) id |= (static_cast<uint32_t>(c >> (2 * io.goff[g])) & io.gmask[g]) << io.gshift[g];
			atomicAdd(s_gbins + id, 1u);
		}
	}
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x)
		if (s_gbins[i]) atomicAdd(io.bins + i, s_gbins[i]);
}

}  // namespace xxm
