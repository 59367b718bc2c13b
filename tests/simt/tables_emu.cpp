// TEST INFRASTRUCTURE ONLY (see simt_emu.h): the device build of the k-mer -> p-value table of a one-part model (SURVEY 8(f) row 1;
// k_tab_count, k_tile_scan, k_tab_leaves, k_tab_rank, k_tab_accumulate, launched as device_build_part0 launches them,
// xxm_gpu_part8.inc) from the CUDA SOURCES under the SIMT emulator, against the product's host build (build_pvalue_tables,
// pvalue_tables.cpp -- itself `==` the reference's tables in every golden trace): ids, classes, integer scores, class sizes and the
// cumulative FP64 class p-values, bit for bit.  Under build.py's sanitizers the same runs are a memcheck and a racecheck.
//
// usage: tables_emu [case ...]; exit code 0 = every case identical
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
#include "xxm_tables_device.cuh"

#include "pvalue_tables.h"

namespace xxm {
// dynamic shared memory of kernels in the included headers that are not launched here (block-scope `extern` after the rewrite)
alignas(16) uint8_t smem5[kPairTabWords * 4];
alignas(16) uint8_t smem4[65536 + 65536];
alignas(16) uint8_t smemM[1 << 16];
uint32_t s_hs[1 << 14];
}  // namespace xxm

using namespace xxm;

namespace {

struct Case {
	std::string name;
	std::vector<int32_t> cols;
	int order;
	double sharp;   // exponent on the column probabilities: sharper PWMs enumerate fewer k-mers
	uint32_t seed;
	bool ties = false;   // columns from a handful of values (IUPAC-like): many k-mers share an integer score, so the stable order inside a class matters
};

int run(const Case& c) {
	std::mt19937 rng(c.seed);
	std::uniform_real_distribution<double> U(0.02, 1.0);
	const size_t nn = size_t(1) << (3 * (c.order + 1));
	std::vector<float> rp(nn), rc(nn);
	for (size_t i = 0; i < nn; ++i) { rp[i] = static_cast<float>(U(rng) / 64); rc[i] = static_cast<float>(U(rng) / 2); }
	NullModelView nm;
	nm.order = c.order; nm.rprobs = rp.data(); nm.rconds = rc.data();
	double bgLog[5] = {0, std::log(.3), std::log(.2), std::log(.2), std::log(.3)};
	double pwm[31 * 5];
	for (int i = 0; i < 31; ++i) {
		double col[4], tot = 0;
		for (int a = 0; a < 4; ++a) { col[a] = c.ties ? ((rng() % 3) ? 0.4 : 0.1) : std::pow(U(rng), c.sharp); tot += col[a]; }
		pwm[i * 5] = 1;
		for (int a = 0; a < 4; ++a) pwm[i * 5 + a + 1] = std::log(col[a] / tot);
	}
	const int ncols = static_cast<int>(c.cols.size());
	PvalueTables H;  // the host build
	build_pvalue_tables(pwm, c.cols.data(), ncols, 100000, 100000000, bgLog, nm, H);
	if (H.error || H.nparts != 1) { std::printf("%-18s host build: error %d, %d parts\n", c.name.c_str(), H.error, H.nparts); return 1; }
	// ---- the device build, as device_build_part0 stages and launches it
	PvalueTables T;
	model_header(pwm, c.cols.data(), ncols, 100000, 100000000, bgLog, T);
	const int len = T.L;
	PartPlan plan;
	part_plan(T, len, plan);
	TabBuild B{};
	B.len = len; B.T = 0;
	std::memcpy(B.sw, plan.sw, sizeof B.sw);
	std::memcpy(B.smax, plan.smax, sizeof B.smax);
	std::memcpy(B.shift, plan.shift, sizeof B.shift);
	std::memcpy(B.code, plan.code, sizeof B.code);
	for (int m = 0; m < len; ++m) B.gaps[m] = T.gaps[m];
	B.order = nm.order; B.rprobs = rp.data(); B.rconds = rc.data();
	const uint32_t npaths = 1u << (2 * len);
	const uint32_t ntiles = (npaths + 255) / 256;
	std::vector<uint32_t> tileCount(ntiles, 0xDEADu);
	std::vector<uint64_t> tileOff(static_cast<size_t>(ntiles) + 2, ~0ull);
	uint64_t total = ~0ull;
	emu::launch(ntiles, 256, [&] { k_tab_count(B, npaths, tileCount.data()); });
	emu::launch((ntiles + 1023) / 1024, 1024, [&] { k_tile_scan(tileCount.data(), ntiles, tileOff.data(), &total); });
	const size_t n = static_cast<size_t>(total);
	// (the library sizes these arrays for all 4^len paths; here they hold exactly the leaves, so that a write past them is seen)
	std::vector<int32_t> sc0(std::max<size_t>(1, n)), sc1(std::max<size_t>(1, n));
	std::vector<uint32_t> id0(std::max<size_t>(1, n)), id1(std::max<size_t>(1, n)), cls(std::max<size_t>(1, n));
	std::vector<double> pr0(std::max<size_t>(1, n)), pr1(std::max<size_t>(1, n));
	std::vector<int32_t> cScore(std::max<size_t>(1, n)), cCount(std::max<size_t>(1, n));
	std::vector<double> cPval(std::max<size_t>(1, n));
	uint32_t nclasses = ~0u;
	LeafOut lo{sc0.data(), id0.data(), pr0.data()};
	emu::launch(ntiles, 256, [&] { k_tab_leaves(B, npaths, tileOff.data(), lo); });
	emu::launch(std::min<uint32_t>(ntiles, 64), 256, [&] { k_tab_rank(sc0.data(), id0.data(), pr0.data(), &total, sc1.data(), id1.data(), pr1.data()); });
	TabClasses C{cls.data(), cScore.data(), cCount.data(), cPval.data(), &nclasses};
	emu::launch(1, 32, [&] { k_tab_accumulate(sc1.data(), pr1.data(), &total, C); });
	// ---- compare
	const TablePart& P = H.part[0];
	bool same = n == P.ids.size() && nclasses == P.classScore.size();
	for (size_t i = 0; same && i < n; ++i) same = id1[i] == P.ids[i] && cls[i] == P.cls[i];
	for (size_t k = 0; same && k < nclasses; ++k)
		same = cScore[k] == P.classScore[k] && cCount[k] == P.classCount[k] && !std::memcmp(&cPval[k], &P.classPval[k], 8);
	std::printf("%-18s %d columns (span %d), order %d: %zu of %u paths enumerated, %u classes  %s\n", c.name.c_str(), ncols, T.span, c.order, n, npaths, nclasses,
	            same ? "identical" : "DIFFERENT");
	return same ? 0 : 1;
}

}  // namespace

int main(int argc, char** argv) {
	const std::vector<Case> cases = {
		{"five_columns", {13, 14, 15, 16, 17}, 2, 1.0, 21},
		{"six_sharp", {10, 11, 12, 13, 14, 15}, 2, 3.0, 22},
		{"seven_order0", {9, 10, 11, 12, 13, 14, 15}, 0, 2.0, 23},
		{"eight_sharp", {12, 13, 14, 15, 16, 17, 18, 19}, 2, 4.0, 24},
		{"eight_flat", {12, 13, 14, 15, 16, 17, 18, 19}, 1, 1.5, 25},
		{"gapped_six", {8, 9, 11, 12, 15, 16}, 2, 2.5, 26},
		{"gapped_order4", {5, 7, 8, 9, 10, 12, 13}, 4, 3.0, 27},
		{"one_column", {15}, 2, 1.0, 28},
		{"ties_seven", {9, 10, 11, 12, 13, 14, 15}, 2, 1.0, 29, true},
		{"ties_gapped_eight", {6, 7, 8, 10, 11, 12, 13, 15}, 1, 1.0, 30, true},
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
