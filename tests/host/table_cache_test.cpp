// TEST (CPU): build_pvalue_tables with the background-probability cache (BgProbCache, pvalue_tables.h) against fresh builds --
// every array of every table bit-identical -- over a sequence of PWMs and column sets like a refinement run's (same null model,
// PWMs drifting, 5..8 columns, gapped sets and two-part models in between, a change of the null model).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "pvalue_tables.h"
using namespace xxm;

static bool same(const PvalueTables& a, const PvalueTables& b) {
	if (a.error != b.error || a.L != b.L || a.nparts != b.nparts || a.span != b.span || a.pruneThr != b.pruneThr) return false;
	for (int p = 0; p < a.nparts; ++p) {
		const TablePart &x = a.part[p], &y = b.part[p];
		if (x.len != y.len || x.ids != y.ids || x.cls != y.cls || x.classScore != y.classScore || x.classCount != y.classCount) return false;
		if (x.classPval.size() != y.classPval.size() || std::memcmp(x.classPval.data(), y.classPval.data(), x.classPval.size() * 8)) return false;
		if (x.prob.size() != y.prob.size() || (x.prob.size() && std::memcmp(x.prob.data(), y.prob.data(), x.prob.size() * 8))) return false;
	}
	return a.leftClassProb.size() == b.leftClassProb.size() && (a.leftClassProb.empty() || !std::memcmp(a.leftClassProb.data(), b.leftClassProb.data(), a.leftClassProb.size() * 8));
}

int main() {
	std::mt19937 rng(7);
	std::uniform_real_distribution<double> U(0.02, 1.0);
	const int order = 2;
	const size_t nn = size_t(1) << (3 * (order + 1));
	std::vector<float> rp(nn), rc(nn);
	auto new_model = [&] { for (size_t i = 0; i < nn; ++i) { rp[i] = static_cast<float>(U(rng) / 64); rc[i] = static_cast<float>(U(rng) / 2); } };
	new_model();
	NullModelView nm; nm.order = order; nm.rprobs = rp.data(); nm.rconds = rc.data();
	double bgLog[5] = {0, std::log(.25), std::log(.25), std::log(.25), std::log(.25)};
	BgProbCache cache;
	int bad = 0, builds = 0;
	double base[31][4];
	for (int i = 0; i < 31; ++i) for (int a = 0; a < 4; ++a) base[i][a] = U(rng);
	for (int step = 0; step < 120; ++step) {
		if (step == 70) { new_model(); cache.clear(); }  // what xxm_set_background does
		double pwm[31 * 5];
		for (int i = 0; i < 31; ++i) {
			double tot = 0, col[4];
			for (int a = 0; a < 4; ++a) { base[i][a] = std::max(0.01, base[i][a] + 0.05 * (U(rng) - 0.5)); col[a] = (i % 3 == 0) ? std::pow(base[i][a], 4) : base[i][a]; tot += col[a]; }
			pwm[i * 5] = 1;
			for (int a = 0; a < 4; ++a) pwm[i * 5 + a + 1] = std::log(col[a] / tot);
		}
		std::vector<int32_t> cols;
		const int first = 8 + static_cast<int>(rng() % 4), ncols = (step % 11 == 5) ? 12 : 5 + static_cast<int>(rng() % 4);
		for (int k = 0, c = first; k < ncols; ++k) { cols.push_back(c); c += (step % 7 == 3 && k == 2) ? 2 : 1; }  // sometimes a gap column
		PvalueTables A, B;
		build_pvalue_tables(pwm, cols.data(), ncols, 100000, 100000000, bgLog, nm, A);
		build_pvalue_tables(pwm, cols.data(), ncols, 100000, 100000000, bgLog, nm, B, &cache);
		++builds;
		if (!same(A, B)) { std::fprintf(stderr, "step %d (%d columns): cached build differs\n", step, ncols); ++bad; }
	}
	size_t filled = 0;
	for (const auto& v : cache.p) for (double x : v) if (x == x) ++filled;
	std::printf("%d builds, %d differ, %zu cached probabilities\n", builds, bad, filled);
	return bad ? 1 : (filled ? 0 : 2);
}
