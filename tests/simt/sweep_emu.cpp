// TEST INFRASTRUCTURE ONLY (see simt_emu.h): the device prefilter of getBestSubset's sweep (xxm_sweep.cuh: k_pack_runs, k_sweep_eval,
// k_sweep_long, k_sweep_flag, k_sweep_gather, with the ordered compaction in between) from the CUDA SOURCES under the SIMT emulator,
// staged and launched as sorted_score_runs does (xxm_gpu_part2.inc).  Its contract: the runs it hands to the exact host sweep
// (best_subset_from_runs, host_stats.cpp = StartPosUpdater.h:135-238 with calculateOrderStatisticsPvalue, utils.h:80-102) must lead to
// the SAME subset -- size, log p-value bit for bit, last score -- as the sweep over all runs.  Also: every decided approximation is
// within 1e-9 relative of the formula's sequential evaluation (the warp-scan composition of k_sweep_long against the loop it
// replaces), the minimum key is the minimum of the decided values, and the winning run is among the candidates.
// (libm here is glibc's, not CUDA's: this checks the algorithm and its memory accesses, not the device's log / lgamma.)
//
// usage: sweep_emu [case ...]; exit code 0 = every case holds
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "cuda_runtime.h"  // tests/simt/shim -> simt_emu.h

#include "xxm_kernels.cuh"
#include "xxm_scan4.cuh"
#include "xxm_scan5.cuh"
#include "xxm_scan6.cuh"  // k_tile_scan
#include "xxm_sweep.cuh"

#include "host_stats.h"

namespace xxm {
alignas(16) uint8_t smem4[65536 + 65536];
alignas(16) uint8_t smem5[kPairTabWords * 4];
alignas(16) uint8_t smemM[1 << 16];
uint32_t s_hs[1 << 14];
}  // namespace xxm

using namespace xxm;

namespace {

struct Case {
	std::string name;
	uint32_t n;          // sites
	double runLen;       // mean length of a run of equal scores
	double scale;        // scores ~ scale * u^power
	double power;
	int64_t maxSites;    // N of the order statistics
	int64_t limit;
	uint32_t seed;
};

// the formula of k_sweep_eval / k_sweep_long evaluated term by term, no cap on the terms
double sequential(double p, double K, double N, bool* undecided) {
	*undecided = false;
	double logN = std::lgamma(N + 1.0) - std::lgamma(K + 1.0) - std::lgamma(N - K + 1.0) + K * std::log(p) + (N - K) * std::log(1.0 - p);
	double sumK = 0, Kp = 0;
	for (int it = 0; Kp < N - K; ++it) {
		if (it >= kSweepTerms) { *undecided = true; return logN; }
		const double alpha = (p * (N - (Kp + K))) / ((1.0 - p) * ((Kp + K) + 1.0));
		sumK = (1.0 / alpha) * (sumK + 1.0);
		const double delta = std::log(1.0 + (1.0 / sumK));
		if (std::fabs(logN * 1e-4) > delta || logN > 0) break;
		logN += delta;
		Kp += 1.0;
	}
	return logN;
}

int run(const Case& c) {
	std::mt19937 rng(c.seed);
	std::uniform_real_distribution<double> U(0.0, 1.0);
	// sorted float scores with runs of equal values
	std::vector<float> sc;
	while (sc.size() < c.n) {
		const float v = static_cast<float>(std::min(0.999, c.scale * std::pow(U(rng), c.power) + 1e-30));
		const uint32_t len = 1 + static_cast<uint32_t>(-std::log(1.0 - U(rng)) * (c.runLen - 1.0));
		for (uint32_t k = 0; k < len && sc.size() < c.n; ++k) sc.push_back(v);
	}
	std::sort(sc.begin(), sc.end());
	std::vector<uint32_t> keys(c.n);
	for (uint32_t i = 0; i < c.n; ++i) std::memcpy(&keys[i], &sc[i], 4);
	// ---- run heads by the ordered compaction, then k_pack_runs (sorted_score_runs)
	const uint32_t ntiles = (c.n + kCmpTile - 1) / kCmpTile;
	std::vector<uint32_t> tileCount(ntiles), runStart(c.n);
	std::vector<uint64_t> tileOff(static_cast<size_t>(ntiles) + 2);
	uint64_t total = 0;
	CompactIO io{};
	io.keys = keys.data(); io.n = c.n; io.tileCount = tileCount.data(); io.tileOff = tileOff.data(); io.oidx = runStart.data();
	emu::launch(ntiles, 256, [&] { k_cmp_count(io); });
	emu::launch((ntiles + 1023) / 1024, 1024, [&] { k_tile_scan(tileCount.data(), ntiles, tileOff.data(), &total); });
	emu::launch(ntiles, 256, [&] { k_cmp_scatter(io); });
	const uint32_t nruns = static_cast<uint32_t>(total);
	std::vector<unsigned long long> runPack(static_cast<size_t>(nruns) + 1, 0ull);
	emu::launch((nruns + 255) / 256, 256, [&] { k_pack_runs(keys.data(), runStart.data(), nruns, runPack.data()); });
	// ---- the exact sweep over ALL runs
	std::vector<ScoreRun> all(nruns);
	for (uint32_t r = 0; r < nruns; ++r) {
		const uint32_t s0 = static_cast<uint32_t>(runPack[r]), s1 = (r + 1 < nruns) ? static_cast<uint32_t>(runPack[r + 1]) : c.n;
		all[r] = ScoreRun{sc[s0], static_cast<int64_t>(s1) - static_cast<int64_t>(s0)};
	}
	LogSum ls;
	ls.reserve(std::max<int64_t>(1, c.maxSites));
	const SubsetChoice want = best_subset_from_runs(ls, all, c.n, c.limit, c.maxSites, false, 0.0);
	// ---- the prefilter
	const double N = static_cast<double>(c.maxSites), lim = static_cast<double>(std::min<int64_t>(c.n, c.limit));
	std::vector<double> approx(nruns, 7.0);
	unsigned long long minKey = ~0ull;
	std::vector<uint32_t> sweepLong(static_cast<size_t>(nruns) + 1, 0u);
	emu::launch((nruns + 127) / 128, 128, [&] { k_sweep_eval(runPack.data(), nruns, c.n, N, lim, approx.data(), &minKey, sweepLong.data() + 1, sweepLong.data()); });
	const uint32_t nLong = sweepLong[0];
	emu::launch(4, 256, [&] { k_sweep_long(runPack.data(), nruns, c.n, N, lim, approx.data(), &minKey, sweepLong.data() + 1, sweepLong.data()); });
	int bad = 0;
	// every decided value against the sequential formula; the minimum
	unsigned long long wantMin = ~0ull;
	uint32_t undecidedRuns = 0, worst = 0;
	double worstRel = 0;
	for (uint32_t r = 0; r < nruns; ++r) {
		const double p = static_cast<double>(sc[static_cast<uint32_t>(runPack[r])]);
		const double K = static_cast<double>((r + 1 < nruns) ? static_cast<uint32_t>(runPack[r + 1]) : c.n);
		if (!(p < 1.0 && K <= lim && !(N * p > K))) { bad += approx[r] != 1e300; continue; }
		bool und;
		const double e = sequential(p, K, N, &und);
		if (approx[r] == -1e300) { ++undecidedRuns; bad += und ? 0 : 1; continue; }  // undecided only where the loop needs more than kSweepTerms terms
		if (und) { ++bad; continue; }
		const double rel = std::fabs(approx[r] - e) / std::max(1e-300, std::fabs(e));
		if (rel > worstRel) { worstRel = rel; worst = r; }
		wantMin = std::min(wantMin, sweep_key(approx[r]));
	}
	if (worstRel > 1e-9) { std::fprintf(stderr, "FAIL %s: run %u differs from the sequential formula by %.3g relative\n", c.name.c_str(), worst, worstRel); ++bad; }
	if (minKey != wantMin) { std::fprintf(stderr, "FAIL %s: minimum key differs\n", c.name.c_str()); ++bad; }
	// ---- candidates and the exact sweep over them
	SubsetChoice got = want;
	uint32_t ncand = nruns;
	if (minKey != ~0ull) {
		const double m = sweep_unkey(minKey);
		const double thr = m + 3e-4 * std::fabs(m) + 1e-6;
		std::vector<uint8_t> flag(nruns, 0xEE);
		emu::launch((nruns + 255) / 256, 256, [&] { k_sweep_flag(approx.data(), nruns, thr, flag.data()); });
		const uint32_t nt2 = (nruns + kCmpTile - 1) / kCmpTile;
		std::vector<uint32_t> tc2(nt2), candIdx(nruns);
		std::vector<uint64_t> to2(static_cast<size_t>(nt2) + 2);
		uint64_t tot2 = 0;
		CompactIO f{};
		f.flag = flag.data(); f.n = nruns; f.tileCount = tc2.data(); f.tileOff = to2.data(); f.oidx = candIdx.data();
		emu::launch(nt2, 256, [&] { k_cmp_count(f); });
		emu::launch((nt2 + 1023) / 1024, 1024, [&] { k_tile_scan(tc2.data(), nt2, to2.data(), &tot2); });
		emu::launch(nt2, 256, [&] { k_cmp_scatter(f); });
		ncand = static_cast<uint32_t>(tot2);
		std::vector<unsigned long long> cand(std::max(1u, ncand));
		emu::launch((ncand + 255) / 256, 256, [&] { k_sweep_gather(runPack.data(), candIdx.data(), ncand, nruns, c.n, cand.data()); });
		std::vector<ScoreRun> few(ncand);
		uint32_t prev = 0;
		for (uint32_t k = 0; k < ncand; ++k) {  // (the host side of sorted_score_runs)
			const uint32_t bits = static_cast<uint32_t>(cand[k] >> 32), K = static_cast<uint32_t>(cand[k]);
			float fsc;
			std::memcpy(&fsc, &bits, 4);
			few[k] = ScoreRun{fsc, static_cast<int64_t>(K) - static_cast<int64_t>(prev)};
			prev = K;
		}
		got = best_subset_from_runs(ls, few, c.n, c.limit, c.maxSites, false, 0.0);
	}
	const bool same = got.bestMotifNb == want.bestMotifNb && !std::memcmp(&got.bestPval, &want.bestPval, 8) && got.lastScore == want.lastScore;
	if (!same) { std::fprintf(stderr, "FAIL %s: subset %d / %.17g from the candidates, %d / %.17g from all runs\n", c.name.c_str(), got.bestMotifNb, got.bestPval, want.bestMotifNb, want.bestPval); ++bad; }
	std::printf("%-22s %u sites in %u runs, N = %lld: %u runs to the warp kernel, %u undecided, %u candidates; subset %d sites, log p %.6f  %s\n", c.name.c_str(), c.n, nruns,
	            static_cast<long long>(c.maxSites), nLong, undecidedRuns, ncand, want.bestMotifNb, want.bestPval, bad ? "DIFFERENT" : "identical");
	return bad ? 1 : 0;
}

}  // namespace

int main(int argc, char** argv) {
	const std::vector<Case> cases = {
		{"many_short_runs", 60000, 1.3, 1.0, 3.0, 3000000, 1000000, 91},
		{"strong_motif", 40000, 1.1, 1e-2, 6.0, 5000000, 1000000, 92},
		{"weak_scores", 30000, 2.0, 1.0, 1.2, 400000, 1000000, 93},        // log p close to 0: long tails of the formula
		{"limit_binds", 50000, 1.5, 0.5, 2.0, 2000000, 20000, 94},
		{"few_runs", 3000, 40.0, 0.3, 2.0, 100000, 1000000, 95},
		{"tiny", 5, 1.0, 0.1, 1.0, 1000, 1000, 96},
	};
	int runs = 0, failed = 0;
	for (const Case& c : cases) {
		bool wanted = argc <= 1;
		for (int i = 1; i < argc; ++i) wanted = wanted || c.name == argv[i];
		if (!wanted) continue;
		failed += run(c);
		++runs;
	}
	std::printf("%d runs, %d failed\n", runs, failed);
	return failed ? 1 : 0;
}
