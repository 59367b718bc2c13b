// K3 / K5, second generation.
//
// k_select_mops_fwd: updateMultOcc (StartPosUtils.h:12-28) for sets WITHOUT --revcomp.  Only the forward-overlap rule
//   applies there, and it can only ever replace the LAST kept site, so the greedy decomposes into independent
//   "clusters": maximal runs of list entries of one sequence whose consecutive start positions are at most `length`
//   apart (if pos[i-1] + length < pos[i] no kept site at or before pos[i-1] can overlap entry i).  The thread of the
//   first entry of a cluster walks it with the state (last kept entry, its float score) in registers.  Kept sites are
//   final when the next site is accepted or the cluster ends; at that moment their value index is added to a
//   shared-memory histogram, which replaces the sort by score (the float score is a function of the value index).
// k_count_pwm2: Motif::updateCounts (Motif.cpp:653-672) with a persistent grid so that the 31x5 counters are flushed
//   once per CTA.
#pragma once
#include "xxm_kernels.cuh"

namespace xxm {

struct SelectFwdIO {
	const uint32_t* seq;
	const uint32_t* pos;
	const uint32_t* vidx;
	uint32_t n;
	int length;
	int shift;           // startPos = pos + shift (only matters for the positional prior)
	float* kscore;       // [n]
	uint8_t* keep;       // [n] zeroed
	uint32_t* counters;  // [0] += kept sites
	uint32_t* hist;      // [nval] zeroed, or null
	uint32_t nval;
	int smemHist;        // histogram privatised in shared memory (nval * 4 bytes of dynamic smem)
};

__global__ void __launch_bounds__(256) k_select_mops_fwd(Values V, SelectFwdIO io) {
	extern __shared__ uint32_t s_hist[];
	__shared__ uint32_t s_cnt;
	if (io.smemHist)
		for (uint32_t k = threadIdx.x; k < io.nval; k += blockDim.x) s_hist[k] = 0;
	if (threadIdx.x == 0) s_cnt = 0;
	__syncthreads();
	uint32_t* hist = io.smemHist ? s_hist : io.hist;
	uint32_t kept = 0;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < io.n; i += gridDim.x * blockDim.x) {
		const uint32_t sq = io.seq[i];
		if (i > 0 && io.seq[i - 1] == sq && io.pos[i - 1] + io.length >= io.pos[i]) continue;  // not a cluster head
		uint32_t last = kNone, lastV = 0;
		float lastScore = 0.f;
		uint32_t lastPos = 0, prevPos = 0;
		for (uint32_t j = i; j < io.n; ++j) {
			const uint32_t pj = io.pos[j];
			if (j > i && (io.seq[j] != sq || prevPos + io.length < pj)) break;
			prevPos = pj;
			const uint32_t v = io.vidx[j];
			if (v == kNone) continue;
			const double pVal = final_of(V, v, j, static_cast<int>(pj) + io.shift);
			if (!(pVal < 1.0)) continue;
			if (last != kNone && lastPos + io.length >= pj) {
				if (static_cast<double>(lastScore) <= pVal) continue;  // the kept site is at least as good: drop the new one
				io.keep[last] = 0;                                        // the new site replaces the last kept one
				last = kNone;
			}
			if (last != kNone) { ++kept; if (hist) atomicAdd(&hist[lastV], 1u); }
			last = j; lastV = v; lastPos = pj; lastScore = static_cast<float>(pVal);
			io.keep[j] = 1;
			io.kscore[j] = lastScore;
		}
		if (last != kNone) { ++kept; if (hist) atomicAdd(&hist[lastV], 1u); }
	}
	if (kept) atomicAdd(&s_cnt, kept);
	__syncthreads();
	if (threadIdx.x == 0 && s_cnt) atomicAdd(&io.counters[0], s_cnt);
	if (io.smemHist)
		for (uint32_t k = threadIdx.x; k < io.nval; k += blockDim.x)
			if (s_hist[k]) atomicAdd(&io.hist[k], s_hist[k]);
}

// Warp-cooperative: every warp walks 32 entries at a time, finds the chosen ones by ballot and then spends one lane per
// PWM column on each chosen site (30 lanes busy, two shared-memory atomics per lane on distinct counters).
__global__ void __launch_bounds__(256) k_count_pwm2(SeqStream S, CountIO io) {
	__shared__ unsigned int s_cnt[31 * 5];
	for (int k = threadIdx.x; k < 31 * 5; k += blockDim.x) s_cnt[k] = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t t0 = warp * 32; t0 < io.n; t0 += nwarps * 32) {
		const uint32_t t = t0 + lane;
		uint32_t i = t;
		bool use = t < io.n;
		if (use) {
			if (io.idx) i = io.idx[t];
			else if (io.slotIdx) { i = io.slotIdx[t]; use = (i != kNone) && io.slotScore[t] <= io.thr; }
			else use = io.keep[i] && io.kscore[i] <= io.thr;
		}
		uint32_t sq = 0, ps = 0;
		if (use) { sq = io.seq[i]; ps = io.pos[i]; }
		uint32_t chosen = __ballot_sync(0xFFFFFFFFu, use);
		while (chosen) {
			const int src = __ffs(chosen) - 1;
			chosen &= chosen - 1;
			const uint32_t s = __shfl_sync(0xFFFFFFFFu, sq, src) - 1;
			const int p0 = static_cast<int>(__shfl_sync(0xFFFFFFFFu, ps, src)) + io.shift - io.firstColumn + 1;
			const int j = lane + 1;  // PWM column of this lane
			const int p = p0 + lane;
			if (j <= 30 && p >= 1 && p <= S.len[s]) {
				const uint32_t wo = S.uniformWps ? s * S.uniformWps : S.wordOff[s];
				const uint32_t g = static_cast<uint32_t>(p - 1);
				const uint32_t w = wo + (g >> 5);
				if (!((S.nmask[w] >> (g & 31)) & 1u)) {
					const uint32_t b = static_cast<uint32_t>(S.codes[w] >> (2 * (g & 31))) & 3u;
					atomicAdd(&s_cnt[j * 5 + b + 1], 1u);
					atomicAdd(&s_cnt[j * 5], 1u);
				}
			}
		}
	}
	__syncthreads();
	for (int k = threadIdx.x; k < 31 * 5; k += blockDim.x)
		if (s_cnt[k]) atomicAdd(&io.counts[k], static_cast<unsigned long long>(s_cnt[k]));
}

}  // namespace xxm
