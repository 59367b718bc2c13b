// Hand-written sm_100a kernels of the PWM-refinement hot path.  No tensor cores: the path is an HBM-bound
// integer gather/compaction.  All kernels run on the context's stream.
#pragma once
#include "xxm_device.cuh"

namespace xxm {

constexpr int kScanThreads = 256;
constexpr uint64_t kFlagAgg = 1ull << 62;   // decoupled look-back: tile aggregate available
constexpr uint64_t kFlagIncl = 2ull << 62;  // inclusive prefix available
constexpr uint64_t kValMask = (1ull << 62) - 1;

// ---------------------------------------------------------------------------------------------------------------
// K1  encode: raw codes (0=N,1..4) -> packed stream, generating the reverse strand on the fly
//     (createRevcomp, seqset.cpp:128-178: S[n+1] = N, S[n+1+k] = complement(S[n+1-k])).
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_encode(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ rawOff, SeqStream S, uint32_t w0, uint32_t w1,
                         uint64_t* __restrict__ codes, uint32_t* __restrict__ nmask, int revcomp) {
	const uint32_t w = w0 + blockIdx.x * blockDim.x + threadIdx.x;  // words [w0, w1) of the stream
	if (w >= w1 || w >= S.nwords + 2) return;
	if (w >= S.nwords) { codes[w] = 0; nmask[w] = 0xFFFFFFFFu; return; }
	const uint32_t s = find_seq(S, w);
	const int n = S.len[s];
	const int n0 = revcomp ? (n - 1) / 2 : n;
	const uint8_t* r = raw + rawOff[s];
	uint64_t c = 0;
	uint32_t m = 0;
	const int q0 = static_cast<int>(w - S.wordOff[s]) * 32 + 1;
	for (int j = 0; j < 32; ++j) {
		const int q = q0 + j;
		int code = 0;
		if (q <= n0) code = r[q - 1];
		else if (revcomp && q > n0 + 1 && q <= n) { const int f = r[2 * n0 + 2 - q - 1]; code = f ? 5 - f : 0; }
		if (code == 0) m |= 1u << j;
		else c |= static_cast<uint64_t>(code - 1) << (2 * j);
	}
	codes[w] = c;
	nmask[w] = m;
}

// scatter compact (id, class) lists into the dense rank table and the "can be a site" bitmap
__global__ void k_fill_rank(const uint32_t* __restrict__ ids, const uint32_t* __restrict__ cls, const uint8_t* __restrict__ clsOk, int n,
                            uint32_t* __restrict__ rank, uint32_t* __restrict__ bitmap) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t id = ids[i], c = cls[i];
	if (!clsOk[c]) return;  // pValueHash value 0 ("not in table", branch_and_bound-inl.h:392) or exactly 1
	rank[id] = c + 1;
	if (bitmap) atomicOr(&bitmap[id >> 5], 1u << (id & 31));
}

// ---------------------------------------------------------------------------------------------------------------
// window evaluation shared by the scan and the re-score kernels: get_kmer_probability_all_positions +
// get_kmer_probability_matched_only (branch_and_bound-inl.h:361-457).  Returns the value index or kNone.
// For > maxk match columns the value is a memo slot keyed by the integer score s (the convolution at :425-449
// depends on the window only through s); slots are evaluated afterwards by k_memo_eval.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t memo_slot(const Model& M, int32_t s) {
	uint32_t h = (static_cast<uint32_t>(s) * 2654435761u) & (M.memoCap - 1);
	for (uint32_t probe = 0; probe < M.memoCap; ++probe) {
		const int32_t old = atomicCAS(&M.memoKey[h], kMemoEmpty, s);
		if (old == kMemoEmpty || old == s) return h;
		h = (h + 1) & (M.memoCap - 1);
	}
	atomicExch(M.memoOverflow, 1u);
	return kNone;
}

__device__ __forceinline__ uint32_t eval_window(const Model& M, uint64_t c0, uint64_t c1, uint32_t m0, uint32_t m1, int j) {
	// N in a match column -> 1.0 (:388,:397,:407); gap columns may hold N
	uint32_t id0 = 0, id1 = 0;
	if (M.contiguous) {
		const uint64_t mm = (static_cast<uint64_t>(m1) << 32 | m0) >> j;
		if (mm & ((M.span >= 32) ? 0xFFFFFFFFull : ((1ull << M.span) - 1))) return kNone;
		id0 = kmer_at(c0, c1, j, M.len0);
		if (M.nparts == 2) {
			for (int k = 0; k < M.len1; ++k) id1 |= base_at(c0, c1, j + M.len0 + k) << (2 * k);
		}
	} else {
		const uint64_t mm = (static_cast<uint64_t>(m1) << 32 | m0);
		for (int k = 0; k < M.L; ++k) {
			const int o = j + M.coloff[k];
			if ((mm >> o) & 1ull) return kNone;
			const uint32_t b = base_at(c0, c1, o);
			if (k < M.len0) id0 |= b << (2 * k);
			else id1 |= b << (2 * (k - M.len0));
		}
	}
	const uint32_t r0 = M.rank0[id0];
	if (r0 == 0) return kNone;
	if (M.nparts == 1) return r0 - 1;
	const uint32_t r1 = M.rank1[id1];
	if (r1 == 0) return kNone;
	if (__dmul_rn(M.cpval0[r0 - 1], M.cpval1[r1 - 1]) > M.pruneThr) return kNone;  // :417
	double score = 0;  // get_Kmer_score :42-50, natural column order, no contraction
	for (int k = 0; k < M.L; ++k) score = __dadd_rn(score, M.logodds[k][base_at(c0, c1, j + M.coloff[k])]);
	const int32_t s = static_cast<int32_t>(__dsub_rn(__dmul_rn(score, 1e6), 1.0));
	return memo_slot(M, s);
}

// ---------------------------------------------------------------------------------------------------------------
// memo evaluation for > maxk columns: the convolution of branch_and_bound-inl.h:419-449 for one integer score s.
// One thread per occupied, not yet evaluated slot; strictly the reference's loop and summation order.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int class_search(const int32_t* score, int n, int S) {  // binarySearch :339-357
	int low = 0, high = n;
	while (low < high) {
		const int mid = low + ((high - low) / 2);
		if (score[mid] > S) low = mid + 1;
		else high = mid;
	}
	if (low < n && score[low] == S) return low;
	return low - 1;
}

// One thread per slot of a compacted list of occupied, not yet evaluated slots (the table itself is sparse: walking it
// slot by slot left one or two busy lanes per warp).  The four class arrays are staged in shared memory when they fit
// (smemTables), so the dependent loads of the merge walk stay inside the SM.  state bit 0: evaluated.
__global__ void __launch_bounds__(128) k_memo_eval_list(Model M, const uint32_t* __restrict__ slots, uint32_t n, int smemTables, double* __restrict__ pS,
                                                        double* __restrict__ finalMops) {
	extern __shared__ __align__(16) uint8_t smemM[];
	const double* leftProb = M.leftProb;
	const double* cpval1 = M.cpval1;
	const int32_t* cscore0 = M.cscore0;
	const int32_t* cscore1 = M.cscore1;
	if (smemTables) {
		double* s_lp = reinterpret_cast<double*>(smemM);
		double* s_cp = s_lp + M.nclass0;
		int32_t* s_c0 = reinterpret_cast<int32_t*>(s_cp + M.nclass1);
		int32_t* s_c1 = s_c0 + M.nclass0;
		for (int k = threadIdx.x; k < M.nclass0; k += blockDim.x) { s_lp[k] = M.leftProb[k]; s_c0[k] = M.cscore0[k]; }
		for (int k = threadIdx.x; k < M.nclass1; k += blockDim.x) { s_cp[k] = M.cpval1[k]; s_c1[k] = M.cscore1[k]; }
		__syncthreads();
		leftProb = s_lp; cpval1 = s_cp; cscore0 = s_c0; cscore1 = s_c1;
	}
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	const uint32_t h = slots[t];
	const int32_t s = M.memoKey[h];
	const int s1_max = cscore0[0], s2_max = cscore1[0];
	int it = class_search(cscore1, M.nclass1, s - s1_max);
	double pValue = 0;
	for (int i = 0; i < M.nclass0; ++i) {
		const int s1 = cscore0[i];
		if (s - s2_max > s1) break;
		// the reference walks `it` down one class at a time (while (it != 0 && score[it] < s - s1) --it): the scores descend, so
		// the walk ends on the last class at or below `it` whose score is >= s - s1, or on 0.  A bisection finds the same
		// class; in a warp the walks of 32 scores take turns being the long one, the bisections all take ~log2 steps (a
		// 17-column model has 10^3 .. 10^5 right classes: 118 ms -> a few ms per scan).
		const int x = s - s1;
		if (it > 0 && cscore1[it] < x) {
			int lo = 0, hi = it;  // first index in [0, it] whose score is < x (score[it] is)
			while (lo < hi) {
				const int mid = (lo + hi) >> 1;
				if (cscore1[mid] >= x) lo = mid + 1;
				else hi = mid;
			}
			it = max(lo - 1, 0);
		} else {
			while (it != 0 && cscore1[it] < x) --it;  // (it < 0 cannot happen for a score that passed both tables; kept literal)
		}
		pValue = __dadd_rn(pValue, __dmul_rn(leftProb[i], cpval1[it]));
	}
	M.memoVal[h] = pValue;
	M.memoState[h] |= 1u;
	// pS / finalPval(mops): pValScore == 1 stays 1 (pValCalculation.h:70), otherwise * 1.05^L (:72)
	const double v = (pValue == 1.0) ? 1.0 : __dmul_rn(pValue, M.corr);
	pS[h] = v;
	finalMops[h] = v;
}

// flag[h] = slot h is occupied and lacks state bit `bit` (0: not evaluated yet, 1: no ZOOPS/OOPS finalPval yet)
__global__ void k_memo_flag(const int32_t* __restrict__ key, const uint32_t* __restrict__ state, uint32_t cap, uint32_t bit, uint8_t* __restrict__ flag) {
	const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
	if (h >= cap) return;
	flag[h] = (key[h] != kMemoEmpty) && !((state[h] >> bit) & 1u);
}
__global__ void k_gather_f64(const double* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t n, double* __restrict__ dst) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_scatter_mode(const double* __restrict__ vals, const uint32_t* __restrict__ idx, uint32_t n, double* __restrict__ dst, uint32_t* __restrict__ state) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	dst[idx[i]] = vals[i];
	state[idx[i]] |= 2u;
}

// ---------------------------------------------------------------------------------------------------------------
// evaluation of explicit windows (re-score of a candidate list, a motif's own sites, per-site p-values):
// the calculatePval part of fill_elongSites_with_possible_startPositions (StartPosUpdater-inl.h:167-190).
// ---------------------------------------------------------------------------------------------------------------
// value index of the window that starts at position pos1 + shift (1-based) of sequence seq1 (1-based), or kNone
__device__ __forceinline__ uint32_t eval_entry(const SeqStream& S, const Model& M, uint32_t seq1, uint32_t pos1, int shift) {
	const uint32_t s = seq1 - 1;
	const int p = static_cast<int>(pos1) + shift;
	// (sets of one length: no gathers from the per-sequence arrays)
	const int len = S.uniformWps ? S.uniformLen : S.len[s];
	if (p < 1 || p + M.span - 1 > len) return kNone;  // :181
	const uint32_t g = static_cast<uint32_t>(p - 1);
	const uint32_t w = (S.uniformWps ? s * S.uniformWps : S.wordOff[s]) + (g >> 5);
	return eval_window(M, S.codes[w], S.codes[w + 1], S.nmask[w], S.nmask[w + 1], g & 31);
}

__global__ void k_eval_list(SeqStream S, Model M, const uint32_t* __restrict__ seq, const uint32_t* __restrict__ pos, uint32_t n, int shift,
                            uint32_t* __restrict__ vidx) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	vidx[i] = eval_entry(S, M, seq[i], pos[i], shift);
}

// ---------------------------------------------------------------------------------------------------------------
// K3  updateMultOcc (StartPosUtils.h:12-65): greedy non-overlap in scan order incl. the reverse-strand look-back.
//     Sequential inside a sequence, independent across sequences: the thread of the first entry of every run of
//     equal sequence numbers walks its run.  `stk` holds the kept entries of the run in place ([a, a+cnt)).
// ---------------------------------------------------------------------------------------------------------------
struct SelectIO {
	const uint32_t* seq;
	const uint32_t* pos;
	const uint32_t* vidx;
	uint32_t n;
	int shift;           // added to pos (column-variant offset of the re-score, StartPosUpdater-inl.h:170)
	int length;          // motif span
	int revcomp;
	uint32_t* stk;       // [n] scratch
	float* kscore;       // [n] float(finalPval) of kept entries
	uint8_t* keep;       // [n] out (zeroed before launch)
	uint32_t* counters;  // [0] nbSites, [1] max kept per sequence
};

__device__ __forceinline__ void mops_walk(const SeqStream& S, const Values& V, const SelectIO& io, uint32_t a, uint32_t b) {
	uint32_t cnt = 0;
	uint32_t* stk = io.stk + a;
	for (uint32_t i = a; i < b; ++i) {
		const uint32_t v = io.vidx[i];
		if (v == kNone) continue;
		const int seq = static_cast<int>(io.seq[i]);
		const int pos = static_cast<int>(io.pos[i]) + io.shift;
		const double pVal = final_of(V, v, i, pos);
		if (!(pVal < 1.0)) continue;
		bool reject = false;
		if (cnt > 0) {
			const uint32_t t = stk[cnt - 1];
			if (static_cast<int>(io.seq[t]) == seq && static_cast<int>(io.pos[t]) + io.shift + io.length >= pos) {
				if (static_cast<double>(io.kscore[t]) <= pVal) reject = true;
				else --cnt;
			}
		}
		if (!reject && io.revcomp && pos > io.length / 2) {
			int back = 1;
			const int revPos = S.len[seq - 1] - io.length + 2 - pos;
			while (static_cast<int>(cnt) - back >= 0 && static_cast<int>(io.seq[stk[cnt - back]]) == seq &&
			       static_cast<int>(io.pos[stk[cnt - back]]) + io.shift > revPos + io.length - 1)
				++back;
			if (static_cast<int>(cnt) - back >= 0 && static_cast<int>(io.seq[stk[cnt - back]]) == seq &&
			    static_cast<int>(io.pos[stk[cnt - back]]) + io.shift >= revPos - io.length + 1) {
				if (static_cast<double>(io.kscore[stk[cnt - back]]) <= pVal) reject = true;
				else {
					while (back > 1) { stk[cnt - back] = stk[cnt - back + 1]; --back; }
					--cnt;
				}
			}
		}
		if (reject) continue;
		stk[cnt++] = i;
		io.kscore[i] = static_cast<float>(pVal);
	}
	for (uint32_t k = 0; k < cnt; ++k) io.keep[stk[k]] = 1;
	if (cnt) { atomicAdd(&io.counters[0], cnt); atomicMax(&io.counters[1], cnt); }
}

__global__ void k_select_mops(SeqStream S, Values V, SelectIO io) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= io.n) return;
	if (i > 0 && io.seq[i - 1] == io.seq[i]) return;  // not a run head
	uint32_t b = i + 1;
	while (b < io.n && io.seq[b] == io.seq[i]) ++b;
	mops_walk(S, V, io, i, b);
}
__global__ void k_select_mops_serial(SeqStream S, Values V, SelectIO io) {
	if (blockIdx.x == 0 && threadIdx.x == 0) mops_walk(S, V, io, 0, io.n);
}

// ---------------------------------------------------------------------------------------------------------------
// ZOOPS / OOPS: best site per sequence with the reference's comparison `finalPval(double) < slot.score(float)`
// in list order (StartPosUpdater-inl.h:196-203, :253-257, :47-53).  slotIdx[seq-1] = entry index or kNone.
// ---------------------------------------------------------------------------------------------------------------
struct SlotIO {
	const uint32_t* seq;
	const uint32_t* pos;
	const uint32_t* vidx;
	uint32_t n;
	int shift;
	uint32_t* slotIdx;   // [nslots] (filled with kNone before launch)
	float* slotScore;    // [nslots] (filled with 1.0f before launch)
	uint32_t* counters;  // [0] number of sequences that received a site
};

__device__ __forceinline__ void zoops_walk(const Values& V, const SlotIO& io, uint32_t a, uint32_t b) {
	for (uint32_t i = a; i < b; ++i) {
		const uint32_t v = io.vidx[i];
		if (v == kNone) continue;
		const int pos = static_cast<int>(io.pos[i]) + io.shift;
		const double f = final_of(V, v, i, pos);
		const uint32_t s = io.seq[i] - 1;
		if (f < static_cast<double>(io.slotScore[s])) {
			if (io.slotIdx[s] == kNone) atomicAdd(&io.counters[0], 1u);
			io.slotIdx[s] = i;
			io.slotScore[s] = static_cast<float>(f);
		}
	}
}
// Ordered lists (one contiguous run per sequence).  A block stages the finalPval and the sequence of 1024 entries in shared
// memory (coalesced loads, all table gathers in flight together); the heads of the runs that start in the tile are
// compacted, and one thread per head folds its run in list order out of shared memory -- past the end of the tile out of
// global memory, which happens once per tile -- with the slot in registers, and writes the slot once.  (The first version
// walked every run out of global memory with the slot in global memory: a chain of ~20 dependent round trips per sequence,
// 847 us for 2.05e7 candidates.)
constexpr int kZoopsTile = 1024;
__global__ void __launch_bounds__(256) k_select_zoops(Values V, SlotIO io) {
	__shared__ double s_f[kZoopsTile];
	__shared__ uint32_t s_seq[kZoopsTile];
	const uint32_t base = blockIdx.x * kZoopsTile;
	const bool needPos = V.kind != kFinalTable && V.kind != kFinalPerCand;
#pragma unroll
	for (int k = 0; k < kZoopsTile / 256; ++k) {
		const uint32_t r = k * 256 + threadIdx.x, i = base + r;
		uint32_t sq = 0;      // (sequence numbers start at 1)
		double f = 2.0;       // never below a slot score
		if (i < io.n) {
			sq = io.seq[i];
			const uint32_t v = io.vidx[i];
			if (v != kNone) f = final_of(V, v, i, needPos ? static_cast<int>(io.pos[i]) + io.shift : 0);
		}
		s_seq[r] = sq;
		s_f[r] = f;
	}
	__syncthreads();
	// the heads of the runs that start in this tile, compacted: every walking warp has all its lanes busy
	__shared__ uint16_t s_head[kZoopsTile];
	__shared__ uint32_t s_wsum[8];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint32_t hm = 0;  // bit k: entry 4 * tid + k is the head of a run
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const uint32_t r = 4 * threadIdx.x + k, i = base + r;
		if (i >= io.n) break;
		const uint32_t prev = r ? s_seq[r - 1] : (i ? io.seq[i - 1] : 0u);
		if (s_seq[r] != prev) hm |= 1u << k;
	}
	const uint32_t cnt = __popc(hm);
	uint32_t incl = cnt;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
		if (lane >= d) incl += t;
	}
	if (lane == 31) s_wsum[wid] = incl;
	__syncthreads();
	uint32_t off = incl - cnt, nheads = 0;
#pragma unroll
	for (int k = 0; k < 8; ++k) { if (k < wid) off += s_wsum[k]; nheads += s_wsum[k]; }
#pragma unroll
	for (int k = 0; k < 4; ++k)
		if ((hm >> k) & 1u) s_head[off++] = static_cast<uint16_t>(4 * threadIdx.x + k);
	__syncthreads();
	uint32_t filled = 0;
	for (uint32_t h = threadIdx.x; h < nheads; h += 256) {
		const uint32_t r = s_head[h];
		const uint32_t sq = s_seq[r];
		float slot = 1.0f;
		uint32_t idx = kNone;
		uint32_t j = r;
		for (; j < static_cast<uint32_t>(kZoopsTile) && s_seq[j] == sq; ++j) {
			const double f = s_f[j];
			if (f < static_cast<double>(slot)) { slot = static_cast<float>(f); idx = base + j; }
		}
		if (j == static_cast<uint32_t>(kZoopsTile)) {
			for (uint32_t g = base + kZoopsTile; g < io.n && io.seq[g] == sq; ++g) {
				const uint32_t v = io.vidx[g];
				if (v == kNone) continue;
				const double f = final_of(V, v, g, needPos ? static_cast<int>(io.pos[g]) + io.shift : 0);
				if (f < static_cast<double>(slot)) { slot = static_cast<float>(f); idx = g; }
			}
		}
		if (idx != kNone) {
			io.slotIdx[sq - 1] = idx;
			io.slotScore[sq - 1] = slot;
			++filled;
		}
	}
	filled = __reduce_add_sync(0xFFFFFFFFu, filled);
	if (lane == 0 && filled) atomicAdd(&io.counters[0], filled);
}
__global__ void k_select_zoops_serial(Values V, SlotIO io) {
	if (blockIdx.x == 0 && threadIdx.x == 0) zoops_walk(V, io, 0, io.n);
}

// ---------------------------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_fill_u32(uint32_t* p, uint32_t v, uint32_t n) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = v;
}
__global__ void k_fill_f32(float* p, float v, uint32_t n) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = v;
}

// flags of entries that are sites of their list (finalPval < 1 under the values of the evaluation that made it)
__global__ void k_flag_hits(Values V, const uint32_t* __restrict__ vidx, const uint32_t* __restrict__ pos, uint32_t n, uint8_t* __restrict__ flag,
                            uint32_t* __restrict__ count) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t v = vidx[i];
	const bool ok = (v != kNone) && (final_of(V, v, i, static_cast<int>(pos[i])) < 1.0);
	flag[i] = ok;
	// (the number of hits comes out of the compaction that follows; one atomic per hit on a single counter cost 0.5 ms for
	// 2 x 10^7 candidates)
	(void) count;
}

// The same for table-valued scores through one bit per value index: with a memo of up to 2^24 slots (two-part models) the final
// values are a 128 MB array and 2 x 10^7 random 8-byte gathers cost 0.53 ms; the bits fit the caches (0.2-2 MB).
__global__ void k_site_bits(const double* __restrict__ fin, uint32_t nval, uint32_t* __restrict__ bits) {
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t b = __ballot_sync(0xFFFFFFFFu, v < nval && fin[v] < 1.0);
	if ((threadIdx.x & 31u) == 0 && v < nval) bits[v >> 5] = b;
}
__global__ void k_flag_hits_bits(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ vidx, uint32_t n, uint8_t* __restrict__ flag) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t v = vidx[i];
	flag[i] = (v != kNone) && ((__ldg(bits + (v >> 5)) >> (v & 31u)) & 1u);
}

// ordered compaction of a (seq,pos,vidx) list, without any inter-CTA dependency: per-tile counts (k_cmp_count), exclusive
// prefix over the tiles (k_tile_scan, xxm_scan6.cuh), ordered scatter (k_cmp_scatter).  An element is selected by its flag
// byte or -- keys != nullptr -- when its key differs from its predecessor's (the heads of runs of equal keys in a sorted
// array).  (The first version chained the tiles through a serial decoupled look-back in one thread per tile: 34 000 dependent
// L2 round trips for 8.7e6 elements.)
constexpr int kCmpTile = 2048;
struct CompactIO {
	const uint32_t *seq, *pos, *vidx;
	const uint8_t* flag;
	const uint32_t* keys;                    // run heads of a sorted key array instead of flag bytes
	uint32_t n;
	uint32_t *oseq, *opos, *ovidx, *oidx;  // any may be null
	uint32_t* tileCount;                    // [ntiles]
	const uint64_t* tileOff;                // [ntiles + 1]
};
__device__ __forceinline__ bool cmp_selected(const CompactIO& io, uint32_t i) {
	if (io.keys) return i == 0 || io.keys[i] != io.keys[i - 1];
	return io.flag[i] != 0;
}
__global__ void __launch_bounds__(256) k_cmp_count(CompactIO io) {
	__shared__ uint32_t s_w[8];
	const uint32_t base = blockIdx.x * kCmpTile;
	uint32_t c = 0;
#pragma unroll
	for (int k = 0; k < kCmpTile / 256; ++k) {
		const uint32_t i = base + k * 256 + threadIdx.x;
		c += (i < io.n && cmp_selected(io, i)) ? 1u : 0u;
	}
	c = __reduce_add_sync(0xFFFFFFFFu, c);
	if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) io.tileCount[blockIdx.x] = s_w[0] + s_w[1] + s_w[2] + s_w[3] + s_w[4] + s_w[5] + s_w[6] + s_w[7];
}
__global__ void __launch_bounds__(256) k_cmp_scatter(CompactIO io) {
	__shared__ uint32_t s_w[8];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t base = blockIdx.x * kCmpTile + wid * (kCmpTile / 8);  // a warp owns 256 consecutive elements
	uint32_t balls[kCmpTile / 256];
	uint32_t c = 0;
#pragma unroll
	for (int k = 0; k < kCmpTile / 256; ++k) {
		const uint32_t i = base + k * 32 + lane;
		balls[k] = __ballot_sync(0xFFFFFFFFu, i < io.n && cmp_selected(io, i));
		c += __popc(balls[k]);
	}
	if (lane == 0) s_w[wid] = c;
	__syncthreads();
	uint64_t o = io.tileOff[blockIdx.x];
#pragma unroll
	for (int k = 0; k < 8; ++k)
		if (k < wid) o += s_w[k];
#pragma unroll
	for (int k = 0; k < kCmpTile / 256; ++k) {
		const uint32_t i = base + k * 32 + lane;
		if ((balls[k] >> lane) & 1u) {
			const uint64_t d = o + __popc(balls[k] & ((1u << lane) - 1u));
			if (io.oseq) io.oseq[d] = io.seq[i];
			if (io.opos) io.opos[d] = io.pos[i];
			if (io.ovidx) io.ovidx[d] = io.vidx[i];
			if (io.oidx) io.oidx[d] = i;
		}
		o += __popc(balls[k]);
	}
}

// histogram of kept entries over value indices (table-valued finalPval: the float score is a function of vidx)
__global__ void k_hist_vidx(const uint32_t* __restrict__ vidx, const uint8_t* __restrict__ keep, uint32_t n, uint32_t* __restrict__ hist) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || !keep[i]) return;
	atomicAdd(&hist[vidx[i]], 1u);
}
// nvalShared != 0: the bins fit shared memory (dynamic, nvalShared * 4 bytes) -- per-block histogram, flushed once; the
// slots of ~10^6 sequences fall into a few hundred value indices, which serialises plain global atomics (168 us per call)
__global__ void __launch_bounds__(256) k_hist_slots(const uint32_t* __restrict__ slotIdx, const uint32_t* __restrict__ vidx, uint32_t nslots,
                                                    uint32_t* __restrict__ hist, uint32_t nvalShared) {
	extern __shared__ uint32_t s_hs[];
	for (uint32_t k = threadIdx.x; k < nvalShared; k += blockDim.x) s_hs[k] = 0;
	if (nvalShared) __syncthreads();
	for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += gridDim.x * blockDim.x) {
		const uint32_t i = slotIdx[s];
		if (i == kNone) continue;
		const uint32_t v = vidx[i];
		if (v < nvalShared) atomicAdd(&s_hs[v], 1u);
		else atomicAdd(&hist[v], 1u);
	}
	if (!nvalShared) return;
	__syncthreads();
	for (uint32_t k = threadIdx.x; k < nvalShared; k += blockDim.x)
		if (s_hs[k]) atomicAdd(&hist[k], s_hs[k]);
}

// ---------------------------------------------------------------------------------------------------------------
// K5  PWM recount (Motif::updateCounts, Motif.cpp:653-672): 30-column context of every chosen site.
//     chosen = kept entry whose float score is <= thr (the optimal subset always ends at the end of a run of equal
//     scores, StartPosUpdater.h:184), or an explicit index list.
// ---------------------------------------------------------------------------------------------------------------
struct CountIO {
	const uint32_t *seq, *pos;
	int shift;
	const uint8_t* keep;        // null with idx
	const float* kscore;        // float score per entry (null with idx)
	const uint32_t* idx;        // explicit entries (null with keep)
	const uint32_t* slotIdx;    // per-sequence slots: entry of slot t (kNone: empty), chosen if slotScore[t] <= thr
	const float* slotScore;
	uint32_t n;
	float thr;
	const uint16_t* kcls;       // k_count_pwm3: value index of every kept entry, 0xFFFF for dropped ones; chosen if <= maxClass
	uint32_t maxClass;
	int firstColumn;            // motifColumns.front(): pos - firstColumn + 1 is the position of PWM column 1
	unsigned long long* counts; // [31*5] zeroed
};
// ---------------------------------------------------------------------------------------------------------------
// stable LSD radix sort on 32-bit keys with an index payload (sortSitesByScore, StartPosUtils.h:68-70: positive
// float scores order like their bit patterns; stability == glibc's merge-sort qsort).
// One warp sorts one chunk per pass: digit histogram per chunk -> global scan -> stable scatter with match_any.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSortChunk = 2048;
__global__ void k_rs_hist(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t* __restrict__ hist /*[256][nchunks]*/, uint32_t nchunks) {
	__shared__ uint32_t s_h[256];
	for (int k = threadIdx.x; k < 256; k += blockDim.x) s_h[k] = 0;
	__syncthreads();
	const uint32_t base = blockIdx.x * kSortChunk;
	for (uint32_t k = threadIdx.x; k < kSortChunk; k += blockDim.x) {
		const uint32_t i = base + k;
		if (i < n) atomicAdd(&s_h[(keys[i] >> shift) & 255u], 1u);
	}
	__syncthreads();
	for (int k = threadIdx.x; k < 256; k += blockDim.x) hist[static_cast<size_t>(k) * nchunks + blockIdx.x] = s_h[k];
}
// exclusive scan of hist in (digit, chunk) order over 32-bit counters, in place, three launches: block sums (1024 entries per
// block), exclusive scan of the block sums by one block, then every block scans its entries on top of its offset.  (One CTA
// walking the whole array took 556 us per pass for the 1.1e6 counters of 8.7e6 keys.)
__global__ void __launch_bounds__(256) k_scan_block_sums(const uint32_t* __restrict__ a, uint32_t n, uint32_t* __restrict__ sums) {
	__shared__ uint32_t s_w[8];
	const uint32_t base = blockIdx.x * 1024u + 4u * threadIdx.x;
	uint32_t v = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) v += (base + k < n) ? a[base + k] : 0u;
	v = __reduce_add_sync(0xFFFFFFFFu, v);
	if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x == 0) sums[blockIdx.x] = s_w[0] + s_w[1] + s_w[2] + s_w[3] + s_w[4] + s_w[5] + s_w[6] + s_w[7];
}
// exclusive scan of up to 1024 * 1024 block sums by ONE block of 1024 threads (every thread owns a contiguous slice)
__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* __restrict__ sums, uint32_t nb) {
	__shared__ uint32_t s_w[32];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t per = (nb + 1023u) / 1024u;
	const uint32_t lo = threadIdx.x * per, hi = min(nb, lo + per);
	uint32_t tot = 0;
	for (uint32_t i = lo; i < hi; ++i) tot += sums[i];
	uint32_t incl = tot;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
	if (lane == 31) s_w[wid] = incl;
	__syncthreads();
	uint32_t woff = 0;
	for (int k = 0; k < 32; ++k) if (k < wid) woff += s_w[k];
	uint32_t run = woff + incl - tot;
	for (uint32_t i = lo; i < hi; ++i) { const uint32_t v = sums[i]; sums[i] = run; run += v; }
}
__global__ void __launch_bounds__(256) k_scan_apply(uint32_t* __restrict__ a, uint32_t n, const uint32_t* __restrict__ sums) {
	__shared__ uint32_t s_w[8];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t base = blockIdx.x * 1024u + 4u * threadIdx.x;
	uint32_t v[4], tot = 0;
#pragma unroll
	for (int k = 0; k < 4; ++k) { v[k] = (base + k < n) ? a[base + k] : 0u; tot += v[k]; }
	uint32_t incl = tot;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
	if (lane == 31) s_w[wid] = incl;
	__syncthreads();
	uint32_t run = sums[blockIdx.x] + incl - tot;
#pragma unroll
	for (int k = 0; k < 8; ++k) if (k < wid) run += s_w[k];
#pragma unroll
	for (int k = 0; k < 4; ++k) { if (base + k < n) a[base + k] = run; run += v[k]; }
}
// Stable scatter of one chunk by one CTA of eight warps: warp w owns keys [256 w, 256 w + 256) of the chunk.  Pass 1: every warp
// ranks its keys among the warp's earlier keys of the same digit (match_any inside a row of 32, a per-warp running count across
// rows) and leaves its digit counts in shared memory; then thread d turns the eight counts of digit d into the warps' first
// output positions (chunk offset from the global scan + the counts of the warps before); pass 2: the keys go out.  (The first
// version gave the whole chunk to a single warp: 64 dependent rows, 67 us per pass for 2.5e6 keys.)
constexpr int kSortWarps = 8;
__global__ void __launch_bounds__(kSortWarps * 32) k_rs_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t n, int shift,
                                                                const uint32_t* __restrict__ hist, uint32_t nchunks, uint32_t* __restrict__ okeys,
                                                                uint32_t* __restrict__ ovals) {
	__shared__ uint32_t s_cnt[kSortWarps][256];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	for (int k = threadIdx.x; k < kSortWarps * 256; k += kSortWarps * 32) (&s_cnt[0][0])[k] = 0;
	__syncthreads();
	constexpr int kRows = kSortChunk / (kSortWarps * 32);  // 8
	const uint32_t base = blockIdx.x * kSortChunk + w * (kRows * 32);
	uint32_t key[kRows], dig[kRows], rnk[kRows];
#pragma unroll
	for (int r = 0; r < kRows; ++r) {
		const uint32_t i = base + r * 32 + lane;
		const bool ok = i < n;
		key[r] = ok ? keys[i] : 0u;
		const uint32_t d = ok ? ((key[r] >> shift) & 255u) : 256u + lane;  // inactive lanes get unique pseudo digits
		const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
		const uint32_t rank = __popc(peers & ((1u << lane) - 1));
		uint32_t prev = 0;
		if (ok && rank == 0) {  // the lowest lane of the group keeps the warp's running count of the digit
			prev = s_cnt[w][d];
			s_cnt[w][d] = prev + __popc(peers);
		}
		prev = __shfl_sync(0xFFFFFFFFu, prev, __ffs(peers) - 1);
		dig[r] = d;
		rnk[r] = prev + rank;
		__syncwarp();
	}
	__syncthreads();
	{
		const int d = threadIdx.x;  // 256 threads, 256 digits
		uint32_t run = hist[static_cast<size_t>(d) * nchunks + blockIdx.x];
#pragma unroll
		for (int k = 0; k < kSortWarps; ++k) { const uint32_t c = s_cnt[k][d]; s_cnt[k][d] = run; run += c; }
	}
	__syncthreads();
#pragma unroll
	for (int r = 0; r < kRows; ++r) {
		const uint32_t i = base + r * 32 + lane;
		if (i < n) {
			const uint32_t dst = s_cnt[w][dig[r]] + rnk[r];
			okeys[dst] = key[r];
			ovals[dst] = vals[i];
		}
	}
}

// gather float-score keys (bit patterns) of the members of a list
__global__ void k_gather_keys(const float* __restrict__ score, const uint32_t* __restrict__ idx, uint32_t n, uint32_t* __restrict__ keys,
                              uint32_t* __restrict__ vals) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	keys[i] = __float_as_uint(score[idx[i]]);
	vals[i] = idx[i];
}
// materialise xxm_site records: (score from kscore/float table, pos + shift, seq) for entries idx[0..n)
struct SiteRec { float score; int32_t startPos; int32_t sequence; };
__global__ void k_emit_sites(const uint32_t* __restrict__ seq, const uint32_t* __restrict__ pos, const float* __restrict__ score, const uint32_t* __restrict__ idx,
                             uint32_t n, int shift, SiteRec* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t e = idx ? idx[i] : i;
	SiteRec r;
	r.score = score[e];
	r.startPos = static_cast<int32_t>(pos[e]) + shift;
	r.sequence = static_cast<int32_t>(seq[e]);
	out[i] = r;
}
// float(pS) of list entries (bestStartPos[].score, StartPosUpdater-inl.h:105) or float(finalPval)
__global__ void k_entry_scores(Values V, const uint32_t* __restrict__ vidx, const uint32_t* __restrict__ pos, uint32_t n, int shift, int wantFinal,
                               float* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t v = vidx[i];
	float r = 1.0f;
	if (v != kNone) r = static_cast<float>(wantFinal ? final_of(V, v, i, static_cast<int>(pos[i]) + shift) : V.pS[v]);
	out[i] = r;
}
// doubles for xxm_window_pvals
__global__ void k_entry_doubles(Values V, const uint32_t* __restrict__ vidx, const uint32_t* __restrict__ pos, uint32_t n, double* __restrict__ fin,
                                double* __restrict__ ps) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t v = vidx[i];
	double f = 1.0, p = 1.0;
	if (v != kNone) {
		p = V.pS[v];
		f = (p == 1.0) ? 1.0 : final_of(V, v, i, static_cast<int>(pos[i]));
	}
	fin[i] = f;
	ps[i] = p;
}

// ---------------------------------------------------------------------------------------------------------------
// K6  forward / reverse-complement log-odds histogram (Output::printScoreDistribution, output.cpp:1050-1064;
//     calculateMatrixScore_logOdds, Motif.h:584-599).  One thread per 32-base word; double sums in column order.
// ---------------------------------------------------------------------------------------------------------------
struct HistModel {
	int ncols, span;
	int coloff[32];
	double lodF[32][5];  // [k][0] = N penalty, [k][1+code]
	double lodB[32][5];
	double thrF, thrB;
};
__global__ void k_logodds_hist(SeqStream S, HistModel H, uint32_t scanWords, unsigned int* __restrict__ fwd, unsigned int* __restrict__ bwd, int npos) {
	const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= scanWords) return;
	const uint64_t c0 = S.codes[w], c1 = S.codes[w + 1];
	const uint64_t mm = (static_cast<uint64_t>(S.nmask[w + 1]) << 32) | S.nmask[w];
	const uint32_t s = find_seq(S, w);
	const int n = S.len[s];
	const int p0 = static_cast<int>(w - S.wordOff[s]) * 32 + 1;
	for (int j = 0; j < 32; ++j) {
		const int pos = p0 + j;
		if (pos > n - H.span + 1 || pos > npos) break;
		double f = 0, b = 0;
		for (int k = 0; k < H.ncols; ++k) {
			const int o = j + H.coloff[k];
			const int a = ((mm >> o) & 1ull) ? 0 : 1 + static_cast<int>(base_at(c0, c1, o));
			f = __dadd_rn(f, H.lodF[k][a]);
			b = __dadd_rn(b, H.lodB[k][a]);
		}
		if (f > H.thrF) atomicAdd(&fwd[pos - 1], 1u);
		if (b > H.thrB) atomicAdd(&bwd[pos - 1], 1u);
	}
}

}  // namespace xxm
