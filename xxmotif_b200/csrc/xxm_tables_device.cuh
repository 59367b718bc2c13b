// the k-mer -> p-value table of a one-part model (at most 8 match columns) built on the DEVICE,
// bit-identical to the host build (pvalue_tables.cpp) and therefore to Kmer_Generator_Reloaded::set_kmer_to_Pvalue_hash
// (branch_and_bound-inl.h:464-577), create_similar_kmers_rec (:586-603) and calculate_kmer_probability (:55-100).
//
// What makes the reference's construction order-dependent, and how each dependency is kept:
//   * the branch-and-bound recursion visits the bases of every level in descending log-odds order and gives up on a level as
//     soon as the bound fails; because floating-point addition is monotone, a leaf is enumerated iff the bound holds at every
//     level of ITS OWN path, so all 4^len paths are tested independently (k_tab_leaves), and the recursion's visiting order is
//     the numeric order of the path index (level 1 = most significant digit);
//   * leaves enter std::map<int, list> in visiting order: classes by descending integer score, visiting order inside a class
//     = a stable sort by score (k_tab_rank: rank by counting -- a few thousand leaves);
//   * the class p-value is the running sum of the background probabilities in that order: one lane adds them up one by one
//     (k_tab_accumulate), the probabilities themselves having been computed in parallel with the reference's multiplication
//     order per leaf.
#pragma once
#include "xxm_kernels.cuh"

namespace xxm {

struct TabBuild {
	int len;              // match columns of the part (<= 8)
	int T;                // enumeration threshold (0 for the left part)
	double sw[9][5];      // [level 1..len][rank 1..4] log-odds of the level's column, descending
	double smax[10];      // [level] best score still reachable from this level on
	int shift[9];         // [level] bit position of the level's column in the k-mer id (2 * (column - 1))
	int code[9][5];       // [level][rank] base code 0..3 of the rank-th best base
	int gaps[9];          // gap columns after match column m (0-based), for the background probability
	int order;            // NullModel order
	const float* rprobs;
	const float* rconds;
};

struct LeafOut {
	int32_t* score;
	uint32_t* id;
	double* prob;
};

// the path with index t: alive? integer score, k-mer id, background probability
__device__ __forceinline__ bool tab_leaf(const TabBuild& B, uint32_t t, int32_t* scoreOut, uint32_t* idOut) {
	double s = 0;
	uint32_t id = 0;
	for (int j = 1; j <= B.len; ++j) {
		const int a = static_cast<int>((t >> (2 * (B.len - j))) & 3u) + 1;
		const double s1 = __dadd_rn(s, B.sw[j][a]);
		if (__double2int_rz(__dmul_rn(__dadd_rn(s1, B.smax[j + 1]), 1e6)) < B.T) return false;
		id |= static_cast<uint32_t>(B.code[j][a]) << B.shift[j];
		s = s1;
	}
	*scoreOut = __double2int_rz(__dmul_rn(s, 1e6));
	*idOut = id;
	return true;
}

// calculate_kmer_probability for the left part (branch_and_bound-inl.h:55-100): the k-mer expanded with its gap columns,
// 3 bits per character, first character most significant
__device__ __forceinline__ double tab_background(const TabBuild& B, uint32_t kmerId) {
	uint8_t ex[40];
	int n = 0;
	uint32_t id = kmerId;
	for (int m = 0; m < B.len; ++m) {
		ex[n++] = static_cast<uint8_t>((id & 3u) + 1u);
		id >>= 2;
		for (int g = 0; g < B.gaps[m] && n < 39; ++g) ex[n++] = 0;
	}
	const int o = B.order;
	auto packed = [&](int from, int len) {
		unsigned int r = ex[from];
		for (int i = 1; i < len; ++i) r = (r << 3) + ex[from + i];
		return r;
	};
	if (n <= o + 1) return static_cast<double>(B.rprobs[packed(0, n)]);
	double p = static_cast<double>(B.rprobs[packed(0, o + 1)]);
	for (int stop = o + 1; stop < n; ++stop) p = __dmul_rn(p, static_cast<double>(B.rconds[packed(stop - o, o + 1)]));
	return p;
}

__global__ void __launch_bounds__(256) k_tab_count(TabBuild B, uint32_t npaths, uint32_t* __restrict__ tileCount) {
	__shared__ uint32_t s_w[8];
	const uint32_t t = blockIdx.x * 256u + threadIdx.x;
	int32_t sc;
	uint32_t id;
	const bool alive = t < npaths && tab_leaf(B, t, &sc, &id);
	const uint32_t c = __popc(__ballot_sync(0xFFFFFFFFu, alive));
	if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) tileCount[blockIdx.x] = s_w[0] + s_w[1] + s_w[2] + s_w[3] + s_w[4] + s_w[5] + s_w[6] + s_w[7];
}

// leaves in visiting order (compacted by path index) with their background probability
__global__ void __launch_bounds__(256) k_tab_leaves(TabBuild B, uint32_t npaths, const uint64_t* __restrict__ tileOff, LeafOut out) {
	__shared__ uint32_t s_w[8];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t t = blockIdx.x * 256u + threadIdx.x;
	int32_t sc = 0;
	uint32_t id = 0;
	const bool alive = t < npaths && tab_leaf(B, t, &sc, &id);
	const uint32_t ball = __ballot_sync(0xFFFFFFFFu, alive);
	if (lane == 0) s_w[wid] = __popc(ball);
	__syncthreads();
	uint64_t o = tileOff[blockIdx.x];
	for (int k = 0; k < wid; ++k) o += s_w[k];
	if (alive) {
		o += __popc(ball & ((1u << lane) - 1u));
		out.score[o] = sc;
		out.id[o] = id;
		out.prob[o] = tab_background(B, id);
	}
}

// stable sort by descending score through rank by counting: position = #{higher score} + #{equal score, earlier leaf}
__global__ void __launch_bounds__(256) k_tab_rank(const int32_t* __restrict__ score, const uint32_t* __restrict__ id, const double* __restrict__ prob,
                                                  const uint64_t* __restrict__ nDev, int32_t* __restrict__ oscore, uint32_t* __restrict__ oid,
                                                  double* __restrict__ oprob) {
	__shared__ int32_t s_sc[1024];
	const uint32_t n = static_cast<uint32_t>(*nDev);
	for (uint32_t base = blockIdx.x * 256u; base < n; base += gridDim.x * 256u) {
		const uint32_t i = base + threadIdx.x;
		const int32_t mine = i < n ? score[i] : 0;
		uint32_t posn = 0;
		for (uint32_t t0 = 0; t0 < n; t0 += 1024u) {
			__syncthreads();
			for (uint32_t k = threadIdx.x; k < 1024u; k += 256u) s_sc[k] = (t0 + k < n) ? score[t0 + k] : INT32_MIN;
			__syncthreads();
			const uint32_t m = min(1024u, n - t0);
			for (uint32_t k = 0; k < m; ++k) {
				const int32_t v = s_sc[k];
				posn += (v > mine) || (v == mine && t0 + k < i);
			}
		}
		if (i < n) { oscore[posn] = mine; oid[posn] = id[i]; oprob[posn] = prob[i]; }
	}
}

// classes of equal score in sorted order; cumulative p-value accumulated leaf by leaf (lane 0), class of every leaf
struct TabClasses {
	uint32_t* cls;          // [n]
	int32_t* classScore;    // [<= n]
	int32_t* classCount;
	double* classPval;
	uint32_t* nclasses;     // out
};
__global__ void __launch_bounds__(32) k_tab_accumulate(const int32_t* __restrict__ score, const double* __restrict__ prob, const uint64_t* __restrict__ nDev,
                                                       TabClasses C) {
	__shared__ double s_p[32];
	__shared__ int32_t s_s[32];
	const int lane = threadIdx.x;
	const uint32_t n = static_cast<uint32_t>(*nDev);
	double cumulative = 0;
	uint32_t c = 0, count = 0;
	int32_t cur = 0;
	for (uint32_t base = 0; base < n; base += 32u) {
		const uint32_t i = base + lane;
		s_p[lane] = i < n ? prob[i] : 0.0;
		s_s[lane] = i < n ? score[i] : 0;
		__syncwarp();
		if (lane == 0) {
			const uint32_t m = min(32u, n - base);
			for (uint32_t k = 0; k < m; ++k) {
				if (base + k == 0) cur = s_s[0];
				if (s_s[k] != cur) {
					C.classScore[c] = cur; C.classCount[c] = static_cast<int32_t>(count); C.classPval[c] = cumulative;
					++c; count = 0; cur = s_s[k];
				}
				cumulative = __dadd_rn(cumulative, s_p[k]);
				C.cls[base + k] = c;
				++count;
			}
		}
		__syncwarp();
	}
	if (lane == 0) {
		if (n) { C.classScore[c] = cur; C.classCount[c] = static_cast<int32_t>(count); C.classPval[c] = cumulative; ++c; }
		*C.nclasses = c;
	}
}

}  // namespace xxm
