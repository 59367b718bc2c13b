// K3 / K5, second generation.
//
// k_select_mops_fwd: updateMultOcc (StartPosUtils.h:12-28) for sets WITHOUT --revcomp.  Only the forward-overlap rule
//   applies there, and it can only ever replace the LAST kept site, so the greedy decomposes into independent
//   "clusters": maximal runs of list entries of one sequence whose consecutive start positions are at most `length`
//   apart (if pos[i-1] + length < pos[i] no kept site at or before pos[i-1] can overlap entry i).  Every thread owns 8
//   consecutive entries and runs the greedy over them in registers, after replaying the entries of the same cluster that
//   precede its chunk.  Kept sites are added to a shared-memory histogram over value indices, which replaces the sort by
//   score (the float score is a function of the value index).
// k_count_pwm2: Motif::updateCounts (Motif.cpp:653-672), persistent grid, one lane per chosen site, ballot counting.
#pragma once
#include "xxm_kernels.cuh"

namespace xxm {

struct SelectFwdIO {
	const uint32_t* seq;
	const uint32_t* pos;
	const uint32_t* vidx;
	uint32_t n;          // entries, or -- with nDev -- the capacity of the arrays
	const unsigned long long* nDev;  // non-null: the number of entries is read from the device (written by the expand kernel before this launch)
	int length;
	int shift;           // startPos = pos + shift (only matters for the positional prior)
	float* kscore;       // [n]
	uint8_t* keep;       // [n] (every entry is written)
	uint16_t* kcls;      // [n] or null: value index of a kept entry, 0xFFFF for a dropped one (what k_count_pwm3 scans)
	uint32_t* counters;  // [0] += kept sites
	uint32_t* hist;      // [nval] zeroed, or null
	uint32_t nval;
	int smemHist;        // histogram privatised in shared memory (nval * 4 bytes of dynamic smem)
};

constexpr int kSelChunk = 8;

// state of the greedy: the kept site that a later entry can still replace
struct SelState {
	int64_t last = -1;
	float lastScore = 0.f;
	uint32_t lastPos = 0;
};
// one entry of updateMultOcc's forward rule (StartPosUtils.h:14-27).  Returns true when the entry becomes the kept site;
// *replaced says whether it pushed out the previous kept site (false: that one is final).
__device__ __forceinline__ bool sel_step(SelState& st, int64_t j, uint32_t pj, double finj, uint32_t length, bool* replaced) {
	*replaced = false;
	if (!(finj < 1.0)) return false;  // not a site (entries without a value carry p-value 1)
	if (st.last >= 0 && st.lastPos + length >= pj) {
		if (static_cast<double>(st.lastScore) <= finj) return false;  // the kept site is at least as good: drop the new one
		*replaced = true;
	}
	st.last = j; st.lastScore = static_cast<float>(finj); st.lastPos = pj;
	return true;
}

__global__ void __launch_bounds__(256, 4) k_select_mops_fwd(Values V, SelectFwdIO io) {
	extern __shared__ uint32_t s_hist[];
	__shared__ uint32_t s_cnt;
	if (io.smemHist)
		for (uint32_t k = threadIdx.x; k < io.nval; k += blockDim.x) s_hist[k] = 0;
	if (threadIdx.x == 0) s_cnt = 0;
	__syncthreads();
	uint32_t* hist = io.smemHist ? s_hist : io.hist;
	const uint32_t L = static_cast<uint32_t>(io.length);
	const int lane = threadIdx.x & 31;
	const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
	const uint64_t nwarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
	const uint32_t n = io.nDev ? static_cast<uint32_t>(min(*io.nDev, static_cast<unsigned long long>(io.n))) : io.n;
	const uint64_t nchunks = (static_cast<uint64_t>(n) + kSelChunk - 1) / kSelChunk;
	uint32_t kept = 0;
	auto fin_at = [&](int64_t j, uint32_t pj) -> double {
		const uint32_t v = io.vidx[j];
		return (v == kNone) ? 1.0 : final_of(V, v, static_cast<uint32_t>(j), static_cast<int>(pj) + io.shift);
	};
	// Every thread owns kSelChunk consecutive entries (vector loads, all in registers).  It replays the entries of its first
	// entry's cluster that lie before the chunk (read only), decides its own entries, and then looks ahead while a later
	// entry could still replace its last kept site.  Every keep flag is written by the owner of the entry.
	for (uint64_t c0 = warp * 32; c0 < nchunks; c0 += nwarps * 32) {
		const uint64_t c = c0 + lane;
		const int64_t j0 = static_cast<int64_t>(c) * kSelChunk;
		const int m = (c < nchunks) ? static_cast<int>(min(static_cast<uint64_t>(kSelChunk), static_cast<uint64_t>(n) - j0)) : 0;
		uint32_t sq[kSelChunk], ps[kSelChunk], vi[kSelChunk];
		if (m == kSelChunk) {
			const uint4 a0 = *reinterpret_cast<const uint4*>(io.seq + j0), a1 = *reinterpret_cast<const uint4*>(io.seq + j0 + 4);
			const uint4 b0 = *reinterpret_cast<const uint4*>(io.pos + j0), b1 = *reinterpret_cast<const uint4*>(io.pos + j0 + 4);
			const uint4 v0 = *reinterpret_cast<const uint4*>(io.vidx + j0), v1 = *reinterpret_cast<const uint4*>(io.vidx + j0 + 4);
			sq[0] = a0.x; sq[1] = a0.y; sq[2] = a0.z; sq[3] = a0.w; sq[4] = a1.x; sq[5] = a1.y; sq[6] = a1.z; sq[7] = a1.w;
			ps[0] = b0.x; ps[1] = b0.y; ps[2] = b0.z; ps[3] = b0.w; ps[4] = b1.x; ps[5] = b1.y; ps[6] = b1.z; ps[7] = b1.w;
			vi[0] = v0.x; vi[1] = v0.y; vi[2] = v0.z; vi[3] = v0.w; vi[4] = v1.x; vi[5] = v1.y; vi[6] = v1.z; vi[7] = v1.w;
		} else {
#pragma unroll
			for (int k = 0; k < kSelChunk; ++k) {
				const bool in = k < m;
				sq[k] = in ? io.seq[j0 + k] : 0u;
				ps[k] = in ? io.pos[j0 + k] : 0u;
				vi[k] = in ? io.vidx[j0 + k] : kNone;
			}
		}
		double fin[kSelChunk];
#pragma unroll
		for (int k = 0; k < kSelChunk; ++k)
			fin[k] = (k < m && vi[k] != kNone) ? final_of(V, vi[k], static_cast<uint32_t>(j0 + k), static_cast<int>(ps[k]) + io.shift) : 1.0;
		SelState st;
		bool rep;
		uint32_t keepBits = 0, lastSeq = 0;
		if (m > 0) {
			// ---- the part of the cluster before the chunk: back to its head, then forward again
			if (j0 > 0) {
				uint32_t pp = io.pos[j0 - 1];
				if (io.seq[j0 - 1] == sq[0] && pp + L >= ps[0]) {
					int64_t h = j0 - 1;
					while (h > 0) {
						const uint32_t p2 = io.pos[h - 1];
						if (io.seq[h - 1] != sq[0] || p2 + L < pp) break;
						pp = p2;
						--h;
					}
					for (int64_t j = h; j < j0; ++j) {
						const uint32_t pj = io.pos[j];
						sel_step(st, j, pj, fin_at(j, pj), L, &rep);
					}
				}
			}
			// ---- own entries
#pragma unroll
			for (int k = 0; k < kSelChunk; ++k) {
				if (k < m) {
					if (k > 0 && sq[k] != sq[k - 1]) st.last = -1;  // a new sequence starts inside the chunk
					const int64_t old = st.last;
					if (sel_step(st, j0 + k, ps[k], fin[k], L, &rep)) {
						keepBits |= 1u << k;
						if (rep && old >= j0) keepBits &= ~(1u << static_cast<int>(old - j0));
						lastSeq = sq[k];
					}
				}
			}
			// ---- can a later entry still replace my last kept site?
			if (st.last >= j0) {
				const uint32_t mySeq = lastSeq;
				for (int64_t j = j0 + m; j < static_cast<int64_t>(n); ++j) {
					const uint32_t pj = io.pos[j];
					if (io.seq[j] != mySeq || pj > st.lastPos + L) break;
					const double fj = fin_at(j, pj);
					if (!(fj < 1.0) || static_cast<double>(st.lastScore) <= fj) continue;
					keepBits &= ~(1u << static_cast<int>(st.last - j0));
					break;
				}
			}
			// ---- flags and float scores of the chunk
			if (m == kSelChunk) {
				uint2 kb;
				kb.x = (keepBits & 1u) | ((keepBits & 2u) << 7) | ((keepBits & 4u) << 14) | ((keepBits & 8u) << 21);
				kb.y = ((keepBits >> 4) & 1u) | ((keepBits & 32u) << 3) | ((keepBits & 64u) << 10) | ((keepBits & 128u) << 17);
				*reinterpret_cast<uint2*>(io.keep + j0) = kb;
				*reinterpret_cast<float4*>(io.kscore + j0) = make_float4(static_cast<float>(fin[0]), static_cast<float>(fin[1]), static_cast<float>(fin[2]), static_cast<float>(fin[3]));
				*reinterpret_cast<float4*>(io.kscore + j0 + 4) = make_float4(static_cast<float>(fin[4]), static_cast<float>(fin[5]), static_cast<float>(fin[6]), static_cast<float>(fin[7]));
				if (io.kcls) {
					uint32_t cw[4];
#pragma unroll
					for (int k = 0; k < 4; ++k) {
						const uint32_t lo = ((keepBits >> (2 * k)) & 1u) ? (vi[2 * k] & 0xFFFFu) : 0xFFFFu;
						const uint32_t hi = ((keepBits >> (2 * k + 1)) & 1u) ? (vi[2 * k + 1] & 0xFFFFu) : 0xFFFFu;
						cw[k] = lo | (hi << 16);
					}
					*reinterpret_cast<uint4*>(io.kcls + j0) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
				}
			} else {
#pragma unroll
				for (int k = 0; k < kSelChunk; ++k)
					if (k < m) {
						io.keep[j0 + k] = (keepBits >> k) & 1u;
						io.kscore[j0 + k] = static_cast<float>(fin[k]);
						if (io.kcls) io.kcls[j0 + k] = ((keepBits >> k) & 1u) ? static_cast<uint16_t>(vi[k]) : static_cast<uint16_t>(0xFFFFu);
					}
			}
			kept += __popc(keepBits);
		}
		if (hist) {
#pragma unroll
			for (int k = 0; k < kSelChunk; ++k) {
				if ((keepBits >> k) & 1u) atomicAdd(&hist[vi[k]], 1u);
			}
		}
	}
	if (kept) atomicAdd(&s_cnt, kept);
	__syncthreads();
	if (threadIdx.x == 0 && s_cnt) atomicAdd(&io.counters[0], s_cnt);
	if (io.smemHist)
		for (uint32_t k = threadIdx.x; k < io.nval; k += blockDim.x)
			if (s_hist[k]) atomicAdd(&io.hist[k], s_hist[k]);
}

// Motif::updateCounts over the chosen sites.  A warp scans 32 entries at a time and parks the chosen ones in a small
// shared-memory buffer; whenever 32 are pending, every lane takes ONE site: it loads the two code words and N masks
// that cover the 30-column context (all 32 sites' loads in flight together), aligns them to PWM column 1, and the warp
// then counts column by column with three ballots (valid, base bit 0, base bit 1).  Lane j keeps the five counters of
// PWM column j in registers, so there are no atomics until the warp is done.
// the 30-column context of the site in buf[lane]: bases = 2 bits per PWM column (column 1 in bits [1:0]); bit j of bad: PWM
// column j + 1 is outside the sequence or holds an N
__device__ __forceinline__ void site_context(const SeqStream& S, const CountIO& io, const uint32_t* buf, int m, int lane, uint64_t* basesOut, uint32_t* badOut) {
	uint64_t bases = 0;
	uint32_t bad = 0xFFFFFFFFu;
	if (lane < m) {
		const uint32_t i = buf[lane];
		const uint32_t s = io.seq[i] - 1;
		const int g0 = static_cast<int>(io.pos[i]) + io.shift - io.firstColumn;  // 0-based stream offset of PWM column 1 inside the sequence
		const uint32_t wo = S.uniformWps ? s * S.uniformWps : S.wordOff[s];
		const int len = S.uniformWps ? S.uniformLen : S.len[s];
		const int skip = g0 < 0 ? -g0 : 0;  // columns before the start of the sequence
		if (skip < 30 && g0 + skip < len) {
			const int gs = g0 + skip;
			const uint32_t w = wo + (static_cast<uint32_t>(gs) >> 5);
			const int off = gs & 31;
			const uint64_t ca = S.codes[w], cb = S.codes[w + 1];
			const uint64_t na = S.nmask[w], nb = S.nmask[w + 1];
			uint64_t b = ca >> (2 * off);
			if (off) b |= cb << (64 - 2 * off);
			uint64_t nmv = ((nb << 32) | na) >> off;  // the >= 32 padding positions after every sequence are flagged: no length test needed
			b <<= 2 * skip;
			nmv = (nmv << skip) | ((1ull << skip) - 1ull);
			const int over = g0 + 30 - len;  // columns past the end of the sequence (they are flagged N anyway while over <= 32)
			if (over > 0) nmv |= ~0ull << (30 - over);
			bases = b;
			bad = static_cast<uint32_t>(nmv);
		}
	}
	*basesOut = bases;
	*badOut = bad;
}

__device__ __forceinline__ void count_batch(const SeqStream& S, const CountIO& io, const uint32_t* buf, int m, int lane, uint32_t (&c)[5]) {
	uint64_t bases;
	uint32_t bad;
	site_context(S, io, buf, m, lane, &bases, &bad);
#pragma unroll
	for (int j = 0; j < 30; ++j) {
		const uint32_t mv = __ballot_sync(0xFFFFFFFFu, !((bad >> j) & 1u));
		const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, (bases >> (2 * j)) & 1ull);
		const uint32_t m1 = __ballot_sync(0xFFFFFFFFu, (bases >> (2 * j + 1)) & 1ull);
		if (lane == j + 1) {
			c[0] += __popc(mv);
			c[1] += __popc(mv & ~m1 & ~m0);
			c[2] += __popc(mv & ~m1 & m0);
			c[3] += __popc(mv & m1 & ~m0);
			c[4] += __popc(mv & m1 & m0);
		}
	}
}

__global__ void __launch_bounds__(256) k_count_pwm2(SeqStream S, CountIO io) {
	__shared__ unsigned int s_cnt[31 * 5];
	__shared__ uint32_t s_buf[8][64];
	for (int k = threadIdx.x; k < 31 * 5; k += blockDim.x) s_cnt[k] = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
	uint32_t* buf = s_buf[wid];
	uint32_t c[5] = {0, 0, 0, 0, 0};
	int npend = 0;
	uint32_t nsel = 0;  // chosen sites of this warp: row 0 of the count matrix is unused by the PWM and carries their number
	for (uint32_t t0 = warp * 32; t0 < io.n; t0 += nwarps * 32) {
		const uint32_t t = t0 + lane;
		uint32_t i = t;
		bool use = t < io.n;
		if (use) {
			if (io.idx) i = io.idx[t];
			else if (io.slotIdx) { i = io.slotIdx[t]; use = (i != kNone) && io.slotScore[t] <= io.thr; }
			else use = io.keep[i] && io.kscore[i] <= io.thr;
		}
		const uint32_t chosen = __ballot_sync(0xFFFFFFFFu, use);
		if (use) buf[npend + __popc(chosen & ((1u << lane) - 1u))] = i;
		npend += __popc(chosen);
		nsel += __popc(chosen);
		__syncwarp();
		if (npend >= 32) {
			count_batch(S, io, buf, 32, lane, c);
			__syncwarp();
			const uint32_t rest = (lane < npend - 32) ? buf[32 + lane] : 0u;
			__syncwarp();
			buf[lane] = rest;
			npend -= 32;
			__syncwarp();
		}
	}
	if (npend) count_batch(S, io, buf, npend, lane, c);
	if (lane == 0 && nsel) atomicAdd(&s_cnt[0], nsel);
	if (lane >= 1 && lane <= 30) {
#pragma unroll
		for (int k = 0; k < 5; ++k)
			if (c[k]) atomicAdd(&s_cnt[lane * 5 + k], c[k]);
	}
	__syncthreads();
	for (int k = threadIdx.x; k < 31 * 5; k += blockDim.x)
		if (s_cnt[k]) atomicAdd(&io.counts[k], static_cast<unsigned long long>(s_cnt[k]));
}

// One site per lane, counted with shared-memory atomics: at step t lane l handles PWM column (t + l) mod 30, so the 32 lanes
// of a step hit different columns.  Only the four base counters are kept; the column total ([.][0]) is their sum.
__device__ __forceinline__ void count_batch3(const SeqStream& S, const CountIO& io, const uint32_t* buf, int m, int lane, unsigned int* s_cnt) {
	uint64_t bases;
	uint32_t bad;
	site_context(S, io, buf, m, lane, &bases, &bad);
	if (bad == 0xFFFFFFFFu) return;
	const uint32_t lo = static_cast<uint32_t>(bases), hi = static_cast<uint32_t>(bases >> 32);
	int j = lane % 30;
#pragma unroll 6
	for (int t = 0; t < 30; ++t) {
		const uint32_t b = (j < 16 ? lo >> (2 * j) : hi >> (2 * j - 32)) & 3u;
		if (!((bad >> j) & 1u)) atomicAdd(&s_cnt[(j + 1) * 5 + 1 + b], 1u);
		j = (j == 29) ? 0 : j + 1;
	}
}

// The same counts from the compact class array the select kernel leaves behind: a lane reads eight entries at a time
// (16 bytes), the few chosen ones (value index <= maxClass; dropped entries carry 0xFFFF) are parked per warp, and
// count_batch takes them 32 at a time.
__global__ void __launch_bounds__(256) k_count_pwm3(SeqStream S, CountIO io) {
	__shared__ unsigned int s_cnt[31 * 5];
	__shared__ uint32_t s_buf[8][32 + 256];
	for (int k = threadIdx.x; k < 31 * 5; k += blockDim.x) s_cnt[k] = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
	uint32_t* buf = s_buf[wid];
	uint32_t npend = 0;
	// two 16-bit entries per compare: with maxClass < 0x8000, (maxClass + 0x8000) - (entry & 0x7FFF) keeps bit 15 set iff
	// (entry & 0x7FFF) <= maxClass, and no borrow crosses into the upper half; entries >= 0x8000 (0xFFFF = dropped) are masked out
	const uint32_t thr2 = (io.maxClass | (io.maxClass << 16)) | 0x80008000u;
	for (uint64_t t0 = static_cast<uint64_t>(warp) * 256; t0 < io.n; t0 += static_cast<uint64_t>(nwarps) * 256) {
		const uint64_t i0 = t0 + 8ull * lane;
		uint32_t m = 0;  // bit k: entry i0 + k is chosen
		if (i0 + 8 <= io.n) {
			const uint4 q = *reinterpret_cast<const uint4*>(io.kcls + i0);
			const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const uint32_t le = (thr2 - (w4[k] & 0x7FFF7FFFu)) & ~w4[k] & 0x80008000u;
				m |= ((le >> 15) & 1u) << (2 * k);
				m |= (le >> 31) << (2 * k + 1);
			}
		} else {
			for (int k = 0; k < 8; ++k)
				if (i0 + k < io.n && io.kcls[i0 + k] <= io.maxClass) m |= 1u << k;
		}
		const uint32_t cnt = __popc(m);
		uint32_t incl = cnt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
			if (lane >= d) incl += v;
		}
		const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
		if (total == 0) continue;
		uint32_t o = npend + incl - cnt;
		while (m) {
			const int k = __ffs(m) - 1;
			m &= m - 1;
			buf[o++] = static_cast<uint32_t>(i0) + k;
		}
		npend += total;
		__syncwarp();
		uint32_t head = 0;
		while (npend >= 32) {
			count_batch3(S, io, buf + head, 32, lane, s_cnt);
			head += 32;
			npend -= 32;
		}
		if (head) {
			__syncwarp();
			const uint32_t rest = (static_cast<uint32_t>(lane) < npend) ? buf[head + lane] : 0u;
			__syncwarp();
			buf[lane] = rest;
			__syncwarp();
		}
	}
	if (npend) count_batch3(S, io, buf, static_cast<int>(npend), lane, s_cnt);
	__syncthreads();
	for (int k = threadIdx.x; k < 31 * 5; k += blockDim.x) {
		const unsigned int v = (k % 5 == 0) ? (s_cnt[k + 1] + s_cnt[k + 2] + s_cnt[k + 3] + s_cnt[k + 4]) : s_cnt[k];
		if (v) atomicAdd(&io.counts[k], static_cast<unsigned long long>(v));
	}
}

}  // namespace xxm
