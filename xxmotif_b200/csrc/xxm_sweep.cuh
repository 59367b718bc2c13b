// a8 on the device: the run packing and the prefilter of getBestSubset's sweep for lists with very many distinct scores
// (k_pack_runs, k_sweep_eval, k_sweep_long, k_sweep_flag, k_sweep_gather), and the small list helpers of xxm_gpu_part2.inc
// (k_iota, k_flag_score, k_gather_u32, k_emit_slot_sites).
// (In a header of their own so that tests/simt can run them from source; the library includes this file where the kernels used to
// stand.)
#pragma once
#include "xxm_kernels.cuh"

namespace xxm {
// (key of the run, index of its first element) of every run of a sorted key array
__global__ void k_pack_runs(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ start, uint32_t nruns, unsigned long long* __restrict__ out) {
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nruns) return;
	const uint32_t s0 = start[r];
	out[r] = (static_cast<unsigned long long>(keys[s0]) << 32) | s0;
}
// ---- getBestSubset over very many runs: a device prefilter -------------------------------------------------------------
// The sweep (StartPosUpdater.h:135-238) evaluates calculateOrderStatisticsPvalue at the end of every run of equal scores and
// keeps the first strict minimum.  With scores that are not table valued there are 10^5 .. 10^6 runs, and the exact
// evaluation -- glibc log/exp, LOG_sum_i accumulated like the reference's -- is host work.  The device evaluates ALL runs
// approximately (same formula, lgamma for LOG_sum_i, CUDA's log), finds the minimum m, and only the runs within a relative
// 3e-4 of m (the formula's own tail loop stops at a relative 1e-4, so two evaluations of one run can differ by that much and
// no more) are evaluated exactly on the host, in order -- plus run 0, which the sweep always takes first.  A run the guard
// "N p > K" sends to exactly 0 can only win when nothing is negative, in which case run 0 already won.
__device__ __forceinline__ unsigned long long sweep_key(double v) {  // order-preserving map double -> u64
	const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
	return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double sweep_unkey(unsigned long long k) {
	const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
	return __longlong_as_double(static_cast<long long>(b));
}
// packed run r = (score bits << 32 | first index); its last element is the first index of run r + 1 (n for the last run).
// Stage 1, one thread per run: the head of the formula and up to kSweepShort terms of its tail loop -- almost every run ends there.
// A run that needs more (log p close to 0: hundreds of terms) is put on a list for stage 2 instead of holding its warp up.
constexpr int kSweepShort = 24;
constexpr int kSweepTerms = 1024;  // beyond: undecided, left to the exact host evaluation (the loop can run for 10^5 terms)
__device__ __forceinline__ bool sweep_head(const unsigned long long* __restrict__ runPack, uint32_t r, uint32_t nruns, uint32_t n, double N, double lim, double* p,
                                           double* K, double* logN) {
	const unsigned long long pk = runPack[r];
	*p = static_cast<double>(__uint_as_float(static_cast<uint32_t>(pk >> 32)));
	*K = static_cast<double>((r + 1 < nruns) ? static_cast<uint32_t>(runPack[r + 1]) : n);
	if (!(*p < 1.0 && *K <= lim && !(N * *p > *K))) return false;  // score 1, beyond the limit, or the guard's exact zero
	*logN = lgamma(N + 1.0) - lgamma(*K + 1.0) - lgamma(N - *K + 1.0) + *K * log(*p) + (N - *K) * log(1.0 - *p);
	return true;
}
__global__ void k_sweep_eval(const unsigned long long* __restrict__ runPack, uint32_t nruns, uint32_t n, double N, double lim, double* __restrict__ approx,
                             unsigned long long* __restrict__ minKey, uint32_t* __restrict__ longList, uint32_t* __restrict__ nLong) {
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	double v = 1e300;  // not evaluated by the sweep / guard zero
	if (r < nruns) {
		double p, K, logN;
		if (sweep_head(runPack, r, nruns, n, N, lim, &p, &K, &logN)) {
			double sumK = 0, Kp = 0;
			bool done = false;
			for (int it = 0; it < kSweepShort && Kp < N - K; ++it) {
				const double alpha = (p * (N - (Kp + K))) / ((1.0 - p) * ((Kp + K) + 1.0));
				sumK = (1.0 / alpha) * (sumK + 1.0);
				const double delta = log(1.0 + (1.0 / sumK));
				if (fabs(logN * 1e-4) > delta || logN > 0) { done = true; break; }
				logN += delta;
				Kp += 1.0;
			}
			if (!(Kp < N - K)) done = true;
			v = done ? logN : -1e300;  // -1e300: not decided (yet): always a candidate, not part of the minimum
			if (!done) longList[atomicAdd(nLong, 1u)] = r;
		}
		approx[r] = v;
	}
	if (v == -1e300) v = 1e300;
	unsigned long long k = sweep_key(v);
#pragma unroll
	for (int d = 16; d; d >>= 1) k = min(k, __shfl_xor_sync(0xFFFFFFFFu, k, d));
	if ((threadIdx.x & 31) == 0 && k != sweep_key(1e300)) atomicMin(minKey, k);
}
// Stage 2, one WARP per listed run: 32 terms of the tail loop at a time.  Term t turns the running sum into
// sumK' = (sumK + 1) / alpha_t -- an affine map, so the 32 maps compose by a warp scan -- and the logarithms of the 32 terms, the
// expensive part, are taken in parallel; the first term at which the sequential loop would have stopped ends the run.
__global__ void __launch_bounds__(256) k_sweep_long(const unsigned long long* __restrict__ runPack, uint32_t nruns, uint32_t n, double N, double lim,
                                                    double* __restrict__ approx, unsigned long long* __restrict__ minKey, const uint32_t* __restrict__ longList,
                                                    const uint32_t* __restrict__ nLong) {
	const int lane = threadIdx.x & 31;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	const uint32_t cnt = *nLong;
	for (uint32_t q = warp; q < cnt; q += nwarps) {
		const uint32_t r = longList[q];
		double p, K, logN;
		sweep_head(runPack, r, nruns, n, N, lim, &p, &K, &logN);  // true for every listed run
		double sumK = 0;
		bool done = false;
		for (int t0 = 0; t0 < kSweepTerms && !done; t0 += 32) {
			const double Kp = static_cast<double>(t0 + lane);
			const bool valid = Kp < N - K;
			double A = 1.0, B = 0.0;  // x -> A x + B; identity beyond the end of the loop
			if (valid) { A = ((1.0 - p) * ((Kp + K) + 1.0)) / (p * (N - (Kp + K))); B = A; }
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {  // inclusive scan of compositions: (A,B) after (A1,B1) = (A A1, A B1 + B)
				const double A1 = __shfl_up_sync(0xFFFFFFFFu, A, d), B1 = __shfl_up_sync(0xFFFFFFFFu, B, d);
				if (lane >= d) { B = A * B1 + B; A = A * A1; }
			}
			const double s = A * sumK + B;  // the running sum after term t0 + lane
			const double delta = valid ? log(1.0 + (1.0 / s)) : 0.0;
			double before = delta;          // logN in front of term t0 + lane = logN + the deltas of the lanes below
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const double x = __shfl_up_sync(0xFFFFFFFFu, before, d); if (lane >= d) before += x; }
			const double total = __shfl_sync(0xFFFFFFFFu, before, 31);
			before = logN + (before - delta);
			const bool stop = !valid || fabs(before * 1e-4) > delta || before > 0;
			const uint32_t m = __ballot_sync(0xFFFFFFFFu, stop);
			if (m) {
				logN = __shfl_sync(0xFFFFFFFFu, before, __ffs(m) - 1);
				done = true;
			} else {
				logN += total;
				sumK = __shfl_sync(0xFFFFFFFFu, s, 31);
			}
		}
		if (lane == 0) {
			if (done && logN == logN) { approx[r] = logN; atomicMin(minKey, sweep_key(logN)); }
			// else: stays -1e300, a candidate for the exact evaluation
		}
	}
}
__global__ void k_sweep_flag(const double* __restrict__ approx, uint32_t nruns, double thr, uint8_t* __restrict__ flag) {
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r < nruns) flag[r] = (r == 0) || (approx[r] <= thr);
}
// candidate c = run idx[c]: (score bits << 32 | low 32 bits of its last element's 1-based rank) -- ranks fit 32 bits (n < 2^32)
__global__ void k_sweep_gather(const unsigned long long* __restrict__ runPack, const uint32_t* __restrict__ idx, uint32_t ncand, uint32_t nruns, uint32_t n,
                               unsigned long long* __restrict__ out) {
	const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= ncand) return;
	const uint32_t r = idx[c];
	const uint32_t K = (r + 1 < nruns) ? static_cast<uint32_t>(runPack[r + 1]) : n;
	out[c] = (runPack[r] & 0xFFFFFFFF00000000ull) | K;
}
__global__ void k_iota(uint32_t* p, uint32_t n) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = i;
}
// flag = member && score (<|<=) thr ; keep may be null (all entries are members)
__global__ void k_flag_score(const float* __restrict__ score, const uint8_t* __restrict__ keep, uint32_t n, float thr, int strict, uint8_t* __restrict__ flag) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const bool member = keep ? keep[i] != 0 : true;
	const float s = score[i];
	flag[i] = member && (strict ? (s < thr) : (s <= thr));
}
__global__ void k_gather_u32(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t n, uint32_t* __restrict__ dst) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) dst[i] = src[idx[i]];
}
// slot element s -> site record (empty slot: score 1, pos 1, seq s+1 as fill_sites_with_possible_startPositions_oops initialises)
__global__ void k_emit_slot_sites(const uint32_t* __restrict__ slotIdx, const float* __restrict__ slotScore, const uint32_t* __restrict__ pos, int shift,
                                  const uint32_t* __restrict__ idx, uint32_t n, SiteRec* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t s = idx ? idx[i] : i;
	const uint32_t e = slotIdx[s];
	SiteRec r;
	r.score = slotScore[s];
	r.startPos = (e == kNone) ? 1 : static_cast<int32_t>(pos[e]) + shift;
	r.sequence = static_cast<int32_t>(s) + 1;
	out[i] = r;
}
}  // namespace xxm
