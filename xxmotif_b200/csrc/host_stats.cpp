// See host_stats.h.  Compiled with -ffp-contract=off.
#include "host_stats.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

namespace xxm {

static void parallel_logs(std::vector<double>& logs, int64_t first, int64_t cnt);

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

// every entry of [0, upTo] (at most kPrefixMax) materialised once: the site counts a sweep reads (LOG_sum_i[K]) are small
// next to N, and a sweep over thousands of runs should not pay up to 127 logarithms for each
void LogSum::dense_prefix(int64_t upTo) {
	upTo = std::min<int64_t>(upTo, kPrefixMax);
	const int64_t have = static_cast<int64_t>(prefix_.size()) - 1;
	if (upTo <= have) return;
	const int64_t want = std::min<int64_t>(kPrefixMax, std::max<int64_t>(upTo, 2 * std::max<int64_t>(have, 1 << 16)));
	std::vector<double> logs;
	const int64_t first = std::max<int64_t>(have, 0);
	parallel_logs(logs, first + 1, want - first);
	if (prefix_.empty()) prefix_.push_back(0.0);
	double acc = prefix_.back();
	prefix_.reserve(static_cast<size_t>(want) + 1);
	for (int64_t k = first + 1; k <= want; ++k) {
		acc = logs[static_cast<size_t>(k - first - 1)] + acc;
		prefix_.push_back(acc);
	}
}

double LogSum::at(int64_t i) {
	if (i <= 0) return 0.0;  // LOG_sum_i[0] = 0
	if (i < static_cast<int64_t>(prefix_.size())) return prefix_[static_cast<size_t>(i)];
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

static void parallel_logs(std::vector<double>& logs, int64_t first, int64_t cnt) {
	logs.resize(static_cast<size_t>(cnt));
	const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
	if (cnt < (1 << 16) || nt == 1) {
		for (int64_t k = 0; k < cnt; ++k) logs[k] = std::log(static_cast<double>(first + k));
		return;
	}
	std::vector<std::thread> th;
	const int64_t per = (cnt + nt - 1) / nt;
	for (unsigned t = 0; t < nt; ++t) {
		const int64_t a = t * per, b = std::min<int64_t>(cnt, a + per);
		if (a >= b) break;
		th.emplace_back([&logs, a, b, first]() {
			for (int64_t k = a; k < b; ++k) logs[k] = std::log(static_cast<double>(first + k));
		});
	}
	for (auto& x : th) x.join();
}

void LogSum::dense(int which, int64_t lo, int64_t hi) {
	lo = std::max<int64_t>(0, lo);
	if (hi < lo) return;
	Window& w = win_[which & 1];
	if (!w.v.empty() && w.lo <= lo && hi < w.lo + static_cast<int64_t>(w.v.size())) return;  // already covered
	extend(hi);
	const int64_t c = lo / kStride;           // accumulate from the checkpoint at or below lo
	const int64_t first = c * kStride;        // value known: ckpt_[c]
	std::vector<double> logs;
	parallel_logs(logs, first + 1, hi - first);
	w.lo = lo;
	w.v.assign(static_cast<size_t>(hi - lo + 1), 0.0);
	double acc = ckpt_[static_cast<size_t>(c)];
	if (first >= lo) w.v[static_cast<size_t>(first - lo)] = acc;
	for (int64_t k = first + 1; k <= hi; ++k) {
		acc = logs[static_cast<size_t>(k - first - 1)] + acc;
		if (k >= lo) w.v[static_cast<size_t>(k - lo)] = acc;
	}
}

double LogSum::at_dense(int64_t i) const {
	if (i <= 0) return 0.0;
	if (i < static_cast<int64_t>(prefix_.size())) return prefix_[static_cast<size_t>(i)];
	for (const Window& w : win_)
		if (!w.v.empty() && i >= w.lo && i < w.lo + static_cast<int64_t>(w.v.size())) return w.v[static_cast<size_t>(i - w.lo)];
	const int64_t c = i / kStride;  // (the caller has reserved the checkpoints)
	double acc = ckpt_[static_cast<size_t>(c)];
	for (int64_t k = c * kStride + 1; k <= i; ++k) acc = std::log(static_cast<double>(k)) + acc;
	return acc;
}

// calculateOrderStatisticsPvalue (utils.h:80-102); AT(i) = LogTable::LOG_sum_i[i]
template <class AT>
static inline double order_statistics_pvalue_impl(AT&& at, int64_t motifNb, int64_t N, double p) {
	if (N * p > motifNb || p == 1) return 0;
	double sum_K_1, delta_K_1;
	int64_t K = 0;
	// log_binomial(N, motifNb, p): LOG_sum_i[N] - LOG_sum_i[k] - LOG_sum_i[N-k] + k*log(prob) + (N-k)*log(1-prob)
	double logN = (at(N) - at(motifNb) - at(N - motifNb) + motifNb * std::log(p) + (N - motifNb) * std::log(1 - p));
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

double order_statistics_pvalue(LogSum& ls, int64_t motifNb, int64_t N, double p) {
	return order_statistics_pvalue_impl([&ls](int64_t i) { return ls.at(i); }, motifNb, N, p);
}

float fast_log(float x) {
	if (x != x) return x;
	// (a function-local static object: initialised once, thread-safely -- contexts may be driven from several host threads)
	struct Tables {
		float lg2[1025];
		float diff[1025];
		Tables() {
			float prev = 0.0f;
			lg2[0] = 0.0f;
			diff[1024] = 0.0f;
			for (int i = 1; i <= 1024; ++i) {
				lg2[i] = std::log(static_cast<float>(1024 + i)) * 1.442695041f - 10.0f;  // float overload, as in the reference
				diff[i - 1] = (lg2[i] - prev) * 1.2352E-4f;
				prev = lg2[i];
			}
		}
	};
	static const Tables tb;
	const float* lg2 = tb.lg2;
	const float* diff = tb.diff;
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
		evaluate_with(motifCounter, iPlus1, score, order_statistics_pvalue(ls, motifCounter, maxSites, static_cast<double>(score)));
	}
	void evaluate_with(int motifCounter, int64_t iPlus1, float score, double pValue) {
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
	{
		int64_t tot = 0;
		for (const ScoreRun& r : runs) tot += r.count;
		if (runs.size() > 64) ls.dense_prefix(std::min(tot, lim));
	}
	if (runs.size() < 16384) {
		for (const ScoreRun& r : runs) {
			if (r.score == 1.0f) break;  // :163
			cum += r.count;
			if (cum > lim) break;        // the loop ends inside this run: its end is never evaluated (:162, :184)
			sw.evaluate(static_cast<int>(cum), cum, r.score);
		}
		return sw.c;
	}
	// Very many runs (scores that are not table valued): the evaluation points are independent of each other, only the
	// comparison against the running best is sequential.  All p-values are computed first -- the two LOG_sum_i ranges the
	// points read materialised once (LogSum::dense), the points spread over the host threads -- and the reference's sweep
	// then runs over the stored values.
	std::vector<int64_t> cums;
	cums.reserve(runs.size());
	for (const ScoreRun& r : runs) {
		if (r.score == 1.0f) break;
		cum += r.count;
		if (cum > lim) break;
		cums.push_back(cum);
	}
	const size_t m = cums.size();
	if (!m) return sw.c;
	ls.reserve(maxSites);
	ls.dense(0, cums.front(), cums.back());
	ls.dense(1, maxSites - cums.back(), maxSites - cums.front());
	std::vector<double> pv(m);
	{
		const LogSum& cls = ls;
		const double lsN = ls.at(maxSites);
		auto at = [&cls, lsN, maxSites](int64_t i) { return i == maxSites ? lsN : cls.at_dense(i); };
		const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
		std::vector<std::thread> th;
		const size_t per = (m + nt - 1) / nt;
		for (unsigned t = 0; t < nt; ++t) {
			const size_t a = t * per, b = std::min(m, a + per);
			if (a >= b) break;
			th.emplace_back([&, a, b]() {
				for (size_t k = a; k < b; ++k) pv[k] = order_statistics_pvalue_impl(at, cums[k], maxSites, static_cast<double>(runs[k].score));
			});
		}
		for (auto& x : th) x.join();
	}
	for (size_t k = 0; k < m; ++k) sw.evaluate_with(static_cast<int>(cums[k]), cums[k], runs[k].score, pv[k]);
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
