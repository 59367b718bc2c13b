// Host-side statistics of the refinement path that are sequential or need the C library's log/exp/pow and therefore
// stay on the host (Appendix A.6).  Each function names the reference
// code whose arithmetic it reproduces operation by operation.
#pragma once
#include <cstdint>
#include <vector>

namespace xxm {

// LogTable::LOG_sum_i (LogTable.cpp:22-31): sum_{k<=i} log(k), accumulated sequentially in double.  The reference
// stores total+1 doubles (8 GB for 10^9 positions); only three entries are read per order-statistics evaluation, so
// checkpoints every kStride entries are kept and the rest is re-accumulated from the checkpoint (bit-identical).
class LogSum {
public:
	double at(int64_t i);
	void reserve(int64_t upTo);  // extend the checkpoints (multi-threaded log(), sequential accumulation)
	// Dense windows for sweeps over very many score runs (site scores that are not table valued: localization, more than 8
	// match columns): every entry of [lo, hi] materialised -- log() in parallel, accumulation sequential from the checkpoint
	// below lo, so every value is the reference's -- after which at_dense() is a thread-safe O(1) read.  Two windows are
	// kept (the sweep reads LOG_sum_i[K] and LOG_sum_i[N - K]); a window that already covers [lo, hi] is reused.
	void dense(int which, int64_t lo, int64_t hi);
	void dense_prefix(int64_t upTo);  // [0, min(upTo, kPrefixMax)] materialised for good (grown geometrically); at() reads it first
	double at_dense(int64_t i) const;  // window hit, else the checkpoint walk without the cache (thread safe)

private:
	struct Window { int64_t lo = 0; std::vector<double> v; };
	Window win_[2];
	static constexpr int64_t kPrefixMax = 1 << 24;  // 128 MB of doubles at most
	std::vector<double> prefix_;
	static constexpr int64_t kStride = 128;
	std::vector<double> ckpt_{0.0};  // ckpt_[c] = LOG_sum_i[c * kStride]
	int64_t lastIdx_[4] = {-1, -1, -1, -1};
	double lastVal_[4] = {0, 0, 0, 0};
	int lastSlot_ = 0;
	void extend(int64_t upTo);
};

// calculateOrderStatisticsPvalue (utils.h:80-102) with log_binomial (utils.h:61-64)
double order_statistics_pvalue(LogSum& ls, int64_t motifNb, int64_t N, double p);

// fast_log (utils.h:138-162)
float fast_log(float x);

// Motif::buildPWM (Motif.cpp:584-591) applied to counts[31][5] (column 0 = total) with pseudo = pseudo * K
// (getPseudocounts, utils.h:214-217)
void build_pwm_from_counts(const double* counts31x5, double pseudoTimesK, const double* bg5, double* pwm31x5);

// PVal_Calculator::initPositionalProbs (pValCalculation.cpp:195-248)
void init_positional_probs(std::vector<double>& positionalProb, int max_leng, bool revcomp, int motifLength, int firstMotifColumn, int enrStart,
                           int enrEnd, int offset);

// PVal_Calculator::calculateFinalPval (pValCalculation.h:116-120); sum_i as in pValCalculation.cpp:5-11
double final_pval_localized(const std::vector<double>& sum_i, int len, int delta_0, double p);
void init_sum_i(std::vector<double>& sum_i, int max_leng);

// one run of equal float scores among the selected sites, ascending
struct ScoreRun {
	float score;
	int64_t count;
};
// getBestSubset (StartPosUpdater.h:135-238) over runs of equal scores when no per-sequence cap binds:
// returns bestMotifNb (0 if there is no site) and the score of its last site.
struct SubsetChoice {
	int bestMotifNb = 0;
	int64_t bestCounter = 0;
	double bestPval = 1.0;
	float lastScore = 0.f;
	int bestMotifNbPlus = 0;
};
SubsetChoice best_subset_from_runs(LogSum& ls, const std::vector<ScoreRun>& runs, int64_t nbSites, int64_t limit, int64_t maxSites, bool oneOccurrence,
                                   double plusFracGlobal);
// the literal loop over a sorted site array (scores + 1-based sequence numbers) with per-sequence caps
SubsetChoice best_subset_from_sorted(LogSum& ls, const float* score, const int32_t* seq, int64_t nbSites, int64_t limit, int64_t maxSites, int nent,
                                     int maxMotifsPerSequence, bool oneOccurrence, double plusFracGlobal);

}  // namespace xxm
