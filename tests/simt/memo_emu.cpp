// TEST INFRASTRUCTURE ONLY (see simt_emu.h): k_memo_eval_list -- the slow path of get_kmer_probability_matched_only for motifs with
// more than 8 match columns, one thread per distinct integer score s (xxm_kernels.cuh) -- from its CUDA SOURCE under the SIMT
// emulator, against the reference's loop written out literally (branch_and_bound-inl.h:425-449 with binarySearch :337-357: the
// class pointer walks down one class at a time; the kernel bisects instead), bit for bit in FP64, with the class arrays staged
// in shared memory and read from global memory.  Under build.py's sanitizers a memcheck and a racecheck of the kernel.
//
// usage: memo_emu [case ...]; exit code 0 = every case identical
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

namespace xxm {
alignas(16) uint8_t smemM[200 * 1024];  // the kernel's dynamic shared memory (opt-in maximum of the launch, xxm_create)
uint32_t s_hs[1 << 14];                 // ... and that of another kernel of the header, not launched here
}  // namespace xxm

using namespace xxm;

namespace {

struct Case {
	std::string name;
	int nclass0, nclass1, nkeys;
	uint32_t seed;
};

// binarySearch, branch_and_bound-inl.h:337-357 (scores descending; returns the class with that score, else the next better one)
int reference_search(const std::vector<int32_t>& score, int S) {
	int low = 0, high = static_cast<int>(score.size());
	while (low < high) {
		const int mid = low + ((high - low) / 2);
		if (score[mid] > S) low = mid + 1;
		else high = mid;
	}
	if (low < static_cast<int>(score.size()) && score[low] == S) return low;
	return low - 1;
}

int run(const Case& c) {
	std::mt19937 rng(c.seed);
	std::uniform_real_distribution<double> U(1e-9, 1.0);
	auto descending = [&](int n, int top, int maxGap) {
		std::vector<int32_t> v(static_cast<size_t>(n));
		int s = top;
		for (int i = 0; i < n; ++i) { v[static_cast<size_t>(i)] = s; s -= 1 + static_cast<int>(rng() % static_cast<uint32_t>(maxGap)); }
		return v;
	};
	const std::vector<int32_t> sc0 = descending(c.nclass0, 7500000, 90000), sc1 = descending(c.nclass1, 3100000, c.nclass1 > 5000 ? 300 : 40000);
	std::vector<double> leftProb(sc0.size()), cpval1(sc1.size());
	for (double& p : leftProb) p = U(rng) * 1e-4;
	double acc = 0;
	for (double& p : cpval1) { acc += U(rng) * 1e-5; p = acc; }
	// scores of windows that passed both tables: s = s1 + s2 of some pair of classes (and their neighbours)
	uint32_t cap = 1;
	while (cap < 4u * static_cast<uint32_t>(c.nkeys)) cap <<= 1;
	std::vector<int32_t> key(cap, kMemoEmpty);
	std::vector<uint32_t> slots;
	for (int k = 0; k < c.nkeys; ++k) {
		int32_t s = sc0[rng() % sc0.size()] + sc1[rng() % sc1.size()] + static_cast<int32_t>(rng() % 3) - 1;
		if (k == 0) s = sc0.front() + sc1.front();  // the best window there is
		if (k == 1) s = sc0.back() + sc1.back();    // ... and the worst
		s = std::min(s, sc0.front() + sc1.front());
		uint32_t h = (static_cast<uint32_t>(s) * 2654435761u) & (cap - 1);
		while (key[h] != kMemoEmpty && key[h] != s) h = (h + 1) & (cap - 1);
		if (key[h] == s) continue;
		key[h] = s;
		slots.push_back(h);
	}
	const uint32_t n = static_cast<uint32_t>(slots.size());
	const double corr = std::pow(1.05, 12);
	// ---- the reference, literally
	std::vector<double> want(cap, -1.0);
	unsigned long long visits = 0;
	for (uint32_t h : slots) {
		const int s = key[h];
		const int s1_max = sc0[0], s2_max = sc1[0];
		int it = reference_search(sc1, s - s1_max);
		double pValue = 0;
		for (size_t i = 0; i < sc0.size(); ++i) {
			const int s1 = sc0[i];
			if (s - s2_max > s1) break;
			while (it != 0 && sc1[static_cast<size_t>(it)] < s - s1) --it;
			pValue += leftProb[i] * cpval1[static_cast<size_t>(it)];
			++visits;
		}
		want[h] = pValue;
	}
	// ---- the kernel, class arrays in shared memory when they fit (finalize_memo, xxm_gpu.cu) and in global memory
	const size_t tabBytes = sc0.size() * 12 + sc1.size() * 12;
	int failed = 0;
	for (int smemTables = (tabBytes <= sizeof(smemM)) ? 1 : 0; smemTables >= 0; --smemTables) {
		std::vector<double> val(cap, -1.0), pS(cap, -1.0), fin(cap, -1.0);
		std::vector<uint32_t> state(cap, 0u);
		Model M{};
		M.nclass0 = static_cast<int>(sc0.size()); M.nclass1 = static_cast<int>(sc1.size());
		M.cscore0 = sc0.data(); M.cscore1 = sc1.data(); M.leftProb = leftProb.data(); M.cpval1 = cpval1.data();
		M.memoKey = key.data(); M.memoVal = val.data(); M.memoState = state.data(); M.memoCap = cap; M.corr = corr;
		emu::launch((n + 127) / 128, 128, [&] { k_memo_eval_list(M, slots.data(), n, smemTables, pS.data(), fin.data()); });
		bool same = true;
		for (uint32_t h = 0; h < cap && same; ++h) {
			const bool used = key[h] != kMemoEmpty;
			const double v = used ? ((want[h] == 1.0) ? 1.0 : want[h] * corr) : -1.0;
			same = !std::memcmp(&val[h], &want[h], 8) && state[h] == (used ? 1u : 0u) && !std::memcmp(&pS[h], &v, 8) && !std::memcmp(&fin[h], &v, 8);
		}
		std::printf("%-20s %d x %d classes, %u scores, %llu class visits, tables in %s memory  %s\n", c.name.c_str(), c.nclass0, c.nclass1, n, visits,
		            smemTables ? "shared" : "global", same ? "identical" : "DIFFERENT");
		failed += same ? 0 : 1;
	}
	return failed;
}

}  // namespace

int main(int argc, char** argv) {
	const std::vector<Case> cases = {
		{"one_class_each", 1, 1, 1, 41},
		{"twelve_columns", 37, 50, 400, 42},
		{"many_left", 400, 300, 600, 43},
		{"seventeen_columns", 180, 12000, 700, 44},
		{"too_big_for_smem", 90, 40000, 300, 45},
		{"one_thread_short", 33, 64, 129, 46},
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
