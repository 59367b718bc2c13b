// See host_stats.h.  Compiled with -ffp-contract=off.
#include "host_stats.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

namespace xxm {

void LogSum::extend(int64_t upTo) {
	// LOG_sum_i[i] = LOG_i[i] + LOG_sum_i[i-1], LOG_i[i] = log(i) (LogTable.cpp:26-31).  log() is evaluated in parallel
	// blocks, the accumulation stays strictly sequential so that every partial sum rounds like the reference's.
	const int64_t have = static_cast<int64_t>(ckpt_.size()) - 1;  // checkpoints filled up to index have*kStride
	const int64_t need = (upTo + kStride - 1) / kStride;
	if (need <= have) return;
	const int64_t kBlock = 1 << 22;
	std::vector<double> logs;
	double acc = ckpt_.back();
	int64_t i = have * kStride + 1;
	const int64_t end = need * kStride;
	unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
	while (i <= end) {
		const int64_t cnt = std::min<int64_t>(kBlock, end - i + 1);
		logs.resize(static_cast<size_t>(cnt));
		if (cnt < (1 << 16) || nt == 1) {
			for (int64_t k = 0; k < cnt; ++k) logs[k] = std::log(static_cast<double>(i + k));
		} else {
			std::vector<std::thread> th;
			const int64_t per = (cnt + nt - 1) / nt;
			for (unsigned t = 0; t < nt; ++t) {
				const int64_t a = t * per, b = std::min<int64_t>(cnt, a + per);
				if (a >= b) break;
				th.emplace_back([&logs, a, b, i]() {
					for (int64_t k = a; k < b; ++k) logs[k] = std::log(static_cast<double>(i + k));
				});
			}
			for (auto& x : th) x.join();
		}
		for (int64_t k = 0; k < cnt; ++k) {
			acc = logs[k] + acc;
			if ((i + k) % kStride == 0) ckpt_.push_back(acc);
		}
		i += cnt;
	}
}

void LogSum::reserve(int64_t upTo) { extend(upTo); }

double LogSum::at(int64_t i) {
	if (i <= 0) return 0.0;  // LOG_sum_i[0] = 0
	for (int k = 0; k < 4; ++k) if (lastIdx_[k] == i) return lastVal_[k];
	const int64_t c = i / kStride;
	if (c >= static_cast<int64_t>(ckpt_.size())) extend(i);
	double acc = ckpt_[static_cast<size_t>(c)];
	for (int64_t k = c * kStride + 1; k <= i; ++k) acc = std::log(static_cast<double>(k)) + acc;
	lastIdx_[lastSlot_] = i;
	lastVal_[lastSlot_] = acc;
	lastSlot_ = (lastSlot_ + 1) & 3;
	return acc;
}

double order_statistics_pvalue(LogSum& ls, int64_t motifNb, int64_t N, double p) {
	if (N * p > motifNb || p == 1) return 0;
	double sum_K_1, delta_K_1;
	int64_t K = 0;
	// log_binomial(N, motifNb, p): LOG_sum_i[N] - LOG_sum_i[k] - LOG_sum_i[N-k] + k*log(prob) + (N-k)*log(1-prob)
	double logN = (ls.at(N) - ls.at(motifNb) - ls.at(N - motifNb) + motifNb * std::log(p) + (N - motifNb) * std::log(1 - p));
	double alpha_K_1, sum_K = 0;
	while (K < N - motifNb) {
		alpha_K_1 = (p * (N - (K + motifNb))) / ((1 - p) * ((K + motifNb) + 1));
		sum_K_1 = (1 / alpha_K_1) * (sum_K + 1);
		sum_K = sum_K_1;
		delta_K_1 = std::log(1 + (1 / sum_K_1));
		if (std::fabs(logN * 1e-4) > delta_K_1 || logN > 0) break;
		logN += delta_K_1;
		K++;
	}
	return logN;
}

float fast_log(float x) {
	if (x != x) return x;
	static float lg2[1025];
	static float diff[1025];
	static bool initialized = false;
	if (!initialized) {
		float prev = 0.0f;
		lg2[0] = 0.0f;
		for (int i = 1; i <= 1024; ++i) {
			lg2[i] = std::log(static_cast<float>(1024 + i)) * 1.442695041f - 10.0f;  // float overload, as in the reference
			diff[i - 1] = (lg2[i] - prev) * 1.2352E-4f;
			prev = lg2[i];
		}
		initialized = true;
	}
	int32_t bits;
	std::memcpy(&bits, &x, 4);
	const int32_t a = ((bits & 0x7F800000) >> 23) - 0x7f;
	const int32_t b = (bits & 0x007FE000) >> 13;
	const int32_t c = (bits & 0x00001FFF);
	return (static_cast<float>(a) + lg2[b] + diff[b] * static_cast<float>(c)) * 0.69314718f;
}

void build_pwm_from_counts(const double* counts, double pseudo, const double* bg, double* pwm) {
	for (int a = 0; a < 5; ++a) pwm[a] = 0;
	for (int i = 1; i <= 30; ++i) {
		pwm[i * 5] = counts[i * 5];
		for (int j = 1; j <= 4; ++j)
			pwm[i * 5 + j] = fast_log(static_cast<float>((counts[i * 5 + j] + pseudo * bg[j]) / (counts[i * 5] + pseudo)));
	}
}

void init_positional_probs(std::vector<double>& pp, int max_leng, bool revcomp, int motifLength, int firstMotifColumn, int enrStart, int enrEnd,
                           int offset) {
	if (static_cast<int>(pp.size()) < max_leng + 1) pp.assign(static_cast<size_t>(max_leng) + 1, 0.0);
	int length = max_leng - motifLength + 1;
	if (revcomp) length = (max_leng / 2) - motifLength + 1;
	const int motifStart = enrStart + firstMotifColumn - 1 + offset;
	const int motifEnd = enrEnd + firstMotifColumn - 1 + offset;
	int startRegion = std::max(1, motifStart);
	int endRegion = std::min(motifEnd, length);
	for (int i = 1; i <= length; i++) {
		int dist = 0;
		if (i < startRegion) dist = startRegion - i;
		else if (i > endRegion) dist = i - endRegion;
		const int mn = std::max(1, startRegion - dist);
		const int mx = std::min(length, endRegion + dist);
		pp[i] = (mx - mn + 1.0) / length;
	}
	if (revcomp) {
		const int start = max_leng / 2 + 2;
		const int oldLength = length;
		const int length2 = max_leng - motifLength + 1;
		const int spaceToMiddleStart = max_leng / 2 - (motifStart + motifLength - 1);
		const int spaceToMiddleEnd = max_leng / 2 - (motifEnd + motifLength - 1);
		startRegion = std::max(start, start + spaceToMiddleEnd);
		endRegion = std::min(length2, start + spaceToMiddleStart);
		for (int i = start; i <= length2; i++) {
			int dist = 0;
			if (i < startRegion) dist = startRegion - i;
			else if (i > endRegion) dist = i - endRegion;
			const int mn = std::max(start, startRegion - dist);
			const int mx = std::min(length2, endRegion + dist);
			pp[i] = (mx - mn + 1.0) / (oldLength);
		}
	}
}

void init_sum_i(std::vector<double>& sum_i, int max_leng) {
	sum_i.assign(static_cast<size_t>(max_leng) + 1, 0.0);
	for (int i = 0; i <= max_leng; i++) {
		double s = 0;
		for (int j = i + 1; j <= max_leng; j++) s += 1.0 / j;
		sum_i[i] = s;
	}
}

double final_pval_localized(const std::vector<double>& sum_i, int len, int delta_0, double p) {
	const double plateau = std::pow(1 - p * len / delta_0, delta_0);
	const double hill = std::exp(-p * len * sum_i[delta_0]);
	return 1 - plateau * hill;
}

namespace {
struct Sweep {
	LogSum& ls;
	int64_t maxSites;
	bool oneOccurrence;
	double plusFrac;  // 1 + Global::plusFrac
	SubsetChoice c;
	int bestCounterPlus = 1;
	bool any = false;
	// one evaluation point at the end of a run of equal scores (StartPosUpdater.h:186-225)
	void evaluate(int motifCounter, int64_t iPlus1, float score) {
		const double pValue = order_statistics_pvalue(ls, motifCounter, maxSites, static_cast<double>(score));
		if (pValue < c.bestPval || oneOccurrence) {
			c.bestPval = pValue;
			c.bestMotifNbPlus = c.bestMotifNb = motifCounter;
			bestCounterPlus = static_cast<int>(iPlus1);
			c.bestCounter = iPlus1;
			c.lastScore = score;
		} else if (plusFrac != 1) {
			if (motifCounter > plusFrac * c.bestMotifNb) {
				c.bestMotifNbPlus = std::max(c.bestMotifNb + 1, static_cast<int>(c.bestMotifNb * plusFrac + 0.5));
				bestCounterPlus = std::max(static_cast<int>(c.bestCounter) + 1, static_cast<int>(c.bestCounter * plusFrac + 0.5));
			} else {
				c.bestMotifNbPlus = motifCounter;
				bestCounterPlus = static_cast<int>(iPlus1);
			}
		}
	}
};
}  // namespace

SubsetChoice best_subset_from_runs(LogSum& ls, const std::vector<ScoreRun>& runs, int64_t nbSites, int64_t limit, int64_t maxSites, bool oneOccurrence,
                                   double plusFracGlobal) {
	Sweep sw{ls, maxSites, oneOccurrence, 1 + plusFracGlobal, SubsetChoice(), 1, false};
	if (nbSites == 0) return sw.c;
	sw.c.bestMotifNb = 1;
	sw.c.bestMotifNbPlus = 1;
	sw.c.bestCounter = 1;
	sw.c.bestPval = 1;
	if (!runs.empty()) sw.c.lastScore = runs[0].score;
	const int64_t lim = std::min(nbSites, limit);
	int64_t cum = 0;
	for (const ScoreRun& r : runs) {
		if (r.score == 1.0f) break;  // :163
		cum += r.count;
		if (cum > lim) break;        // the loop ends inside this run: its end is never evaluated (:162, :184)
		sw.evaluate(static_cast<int>(cum), cum, r.score);
	}
	return sw.c;
}

SubsetChoice best_subset_from_sorted(LogSum& ls, const float* score, const int32_t* seq, int64_t nbSites, int64_t limit, int64_t maxSites, int nent,
                                     int maxMotifsPerSequence, bool oneOccurrence, double plusFracGlobal) {
	Sweep sw{ls, maxSites, oneOccurrence, 1 + plusFracGlobal, SubsetChoice(), 1, false};
	if (nbSites == 0) return sw.c;
	sw.c.bestMotifNb = 1;
	sw.c.bestMotifNbPlus = 1;
	sw.c.bestCounter = 1;
	sw.c.bestPval = 1;
	sw.c.lastScore = score[0];
	std::vector<int32_t> perSeq(static_cast<size_t>(nent) + 1, 0);
	int motifCounter = 0;
	const int64_t lim = std::min(nbSites, limit);
	for (int64_t i = 0; i < lim; i++) {
		if (score[i] == 1.0f) break;
		perSeq[seq[i]]++;
		if (perSeq[seq[i]] > maxMotifsPerSequence) continue;
		motifCounter++;
		if (i + 1 < nbSites && score[i + 1] == score[i]) continue;
		sw.evaluate(motifCounter, i + 1, score[i]);
	}
	return sw.c;
}

}  // namespace xxm
