// Host statistics of the localization mode: the enriched start-position region
// (MotifRegion::findMaximum / getRegion, reference src/nucleotides/motifRegion.cpp:13-188, driven by
// Motif::updateEnrichedRegion, refinementPhase/Motif.cpp:968-1019) and the positional p-value of a site list
// (StartPosUpdater::setPositionalPval, refinementPhase/StartPosUpdater.cpp:207-274).  Sequential, tiny, needs log().
#include "host_region.h"

#include <algorithm>
#include <cmath>

namespace xxm {
namespace {

// sliding window of 5 occupied loop positions, weights 1/(distance+1) in INTEGER division like the reference
int find_maximum(const int* cnt, int startRange, int endRange) {
	constexpr int W = 5, H = W / 2;
	int pos[W + 1];
	std::fill(pos, pos + W + 1, -1);
	int best = 0;
	double bestOcc = 0;
	for (int j = startRange; j < endRange; ++j) {
		if (pos[W - 1] == -1) {
			int k = 0;
			while (pos[k] != -1) ++k;
			pos[k] = j;
			continue;
		}
		pos[W] = j;
		const int cur = pos[H];
		double occ = 0.0;
		for (int i = 0; i < H; ++i) occ += cnt[pos[i]] / (cur - pos[i] + 1);
		occ += cnt[cur];
		for (int i = H + 1; i <= W; ++i) occ += cnt[pos[i]] / (pos[i] - cur + 1);
		if (occ > bestOcc) { best = cur; bestOcc = occ; }
		for (int i = 0; i < W; ++i) pos[i] = pos[i + 1];
	}
	if (pos[W] == -1) {
		int most = 0;
		for (int i = 0; i < W && pos[i] != -1; ++i)
			if (cnt[pos[i]] > most) { best = pos[i]; most = cnt[pos[i]]; }
	}
	return best;
}

}  // namespace

Region motif_region(LogSum& ls, const int* cnt, int counts, int startRange, int endRange) {
	const double enrichment = 0.1;
	const double significance = -std::log(1000);
	const double bonferroni = std::log(endRange) + std::log(endRange - 1) - std::log(2);
	const int peak = find_maximum(cnt, startRange, endRange);
	const int L = endRange - startRange + 1;
	auto pval = [&](int a, int b, int x) { return order_statistics_pvalue(ls, x, counts, (b - a + 1.0) / L); };
	int x1 = peak, x2 = peak;
	int inRound = 0;
	for (int i = x1; i <= x2; ++i) inRound += cnt[i];
	int occ = inRound;
	double bestProb = pval(x1, x2, occ);
	double accepted = (occ * 1.0 / counts > enrichment) ? bestProb : 100;
	int x1Prime = x1;
	bool again;
	do {
		again = false;
		bestProb = accepted;
		for (int t = x1 - 1; t >= startRange; --t) {
			if (cnt[t] == 0) continue;
			occ += cnt[t];
			const double p = pval(t, x2, occ);
			if (p < bestProb) { bestProb = p; x1Prime = t; }
		}
		const double left = bestProb;
		int x2Prime = x2;
		occ = inRound;
		bestProb = accepted;
		for (int t = x2 + 1; t <= endRange; ++t) {
			if (cnt[t] == 0) continue;
			occ += cnt[t];
			const double p = pval(x1, t, occ);
			if (p < bestProb) { bestProb = p; x2Prime = t; }
		}
		occ = inRound;
		const double right = bestProb;
		if (std::min(left, right) < accepted) {
			if (accepted == 100) {
				if (left < right) x1 = x1Prime;
				else x2 = x2Prime;
			} else {
				x1 = x1Prime;
				x2 = x2Prime;
			}
			inRound = 0;
			for (int i = x1; i <= x2; ++i) inRound += cnt[i];
			occ = inRound;
			bestProb = pval(x1, x2, occ);
			accepted = (occ * 1.0 / counts > enrichment) ? bestProb : 100;
			again = true;
		}
	} while (again);
	bestProb += bonferroni;
	Region r;
	if (occ * 1.0 / counts > enrichment && bestProb < significance) {
		r.set = 1; r.max = peak; r.start = x1; r.end = x2;
	} else {
		r.set = 0; r.max = startRange; r.start = startRange; r.end = endRange;
	}
	return r;
}

// getCombinedPval_log, pValCalculation.h:122-148
static double combined_pval_log(double logProd, int n) {
	double result = logProd, quotient = 1, sum = 1;
	for (int i = 1; i < n; ++i) {
		quotient *= ((-logProd) / i);
		sum += quotient;
		if (sum > 1e100) { sum /= 1e100; quotient /= 1e100; result += 230.258509299; }
	}
	return result + std::log(sum);
}

bool positional_pval(int max_leng, bool revcomp, int enrMax, int firstMotifColumn, int motifLength, const int32_t* sitePos, int nSites, double* out) {
	double logProd = 0;
	int trivial = 0;
	*out = 0.0;
	if (!positional_logprod(max_leng, revcomp, enrMax, firstMotifColumn, motifLength, sitePos, nSites, &logProd, &trivial)) return false;
	if (!trivial) *out = positional_combine(logProd, nSites);
	return true;
}

double positional_combine(double logProd, int64_t nSites) { return combined_pval_log(logProd, static_cast<int>(nSites)); }

// the sum over the sites of setPositionalPval (StartPosUpdater.cpp:228-262); *trivial = 1 where the reference returns 0 before it
bool positional_logprod(int max_leng, bool revcomp, int enrMax, int firstMotifColumn, int motifLength, const int32_t* sitePos, int nSites, double* out,
                        int* trivial) {
	*out = 0.0;
	*trivial = 0;
	int peak = enrMax + firstMotifColumn - 1;
	if (peak <= 1) { *trivial = 1; return true; }
	int L = max_leng;
	if (revcomp) L = (L - 1) / 2;
	const int L1 = L - motifLength + 1;
	if (revcomp) {
		if (peak > L + 1) peak = 1 + max_leng - peak - motifLength + 1;
		else if (peak > L1) return false;  // "In no man's land" -> the reference exits
	}
	double logProd = 0.0;
	for (int k = 0; k < nSites; ++k) {
		int pos = sitePos[k] + firstMotifColumn - 1;
		if (revcomp) {
			if (pos > L + 1) pos = 1 + max_leng - pos - motifLength + 1;
			else if (pos > L1) return false;
		}
		int d = peak - pos;
		if (d < 0) {
			d = std::abs(d);
			d += std::min(peak - 1, d);
		} else if (d > 0) {
			d += std::min(d, L1 - peak);
		}
		logProd += std::log(static_cast<double>(d + 1) / static_cast<double>(L1));
	}
	*out = logProd;
	return true;
}

}  // namespace xxm
