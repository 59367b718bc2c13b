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
			const double pVal = final_of(V, v, j, static_cast<int>(pj));
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

__global__ void __launch_bounds__(256) k_count_pwm2(SeqStream S, CountIO io) {
	__shared__ unsigned int s_cnt[31 * 5];
	for (int k = threadIdx.x; k < 31 * 5; k += blockDim.x) s_cnt[k] = 0;
	__syncthreads();
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < io.n; t += gridDim.x * blockDim.x) {
		uint32_t i = t;
		bool use = true;
		if (io.idx) i = io.idx[t];
		else if (io.slotIdx) { i = io.slotIdx[t]; use = (i != kNone) && io.slotScore[t] <= io.thr; }
		else use = io.keep[i] && io.kscore[i] <= io.thr;
		if (!use) continue;
		const uint32_t s = io.seq[i] - 1;
		const int n = S.len[s];
		const int p0 = static_cast<int>(io.pos[i]) + io.shift - io.firstColumn + 1;
		const uint32_t wo = S.uniformWps ? s * S.uniformWps : S.wordOff[s];
		// the 30-column context spans at most two 32-base words: load them once
		const int pa = max(p0, 1), pb = min(p0 + 29, n);
		if (pa > pb) continue;
		const uint32_t w0 = wo + (static_cast<uint32_t>(pa - 1) >> 5);
		const uint64_t c0 = S.codes[w0], c1 = S.codes[w0 + 1];
		const uint64_t mm = (static_cast<uint64_t>(S.nmask[w0 + 1]) << 32) | S.nmask[w0];
		const int o0 = (pa - 1) & 31;
		for (int p = pa; p <= pb; ++p) {
			const int o = o0 + (p - pa);
			if ((mm >> o) & 1ull) continue;
			const int j = p - p0 + 1;
			atomicAdd(&s_cnt[j * 5 + base_at(c0, c1, o) + 1], 1u);
			atomicAdd(&s_cnt[j * 5], 1u);
		}
	}
	__syncthreads();
	for (int k = threadIdx.x; k < 31 * 5; k += blockDim.x)
		if (s_cnt[k]) atomicAdd(&io.counts[k], static_cast<unsigned long long>(s_cnt[k]));
}

}  // namespace xxm
