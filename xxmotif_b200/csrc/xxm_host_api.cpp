// Device-free entry points of the C ABI (include/xxmotif_gpu.h, "host-side pieces"): the same host code the GPU path
// uses, exposed so that the host logic can be unit-tested on a machine without a GPU.
#include "../../include/xxmotif_gpu.h"

#include <algorithm>
#include <thread>
#include <vector>

#include "host_pack.h"
#include "host_stats.h"
#include "pair_table.h"
#include "pvalue_tables.h"

using namespace xxm;

extern "C" int xxm_host_tables(const double* pwm, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t totalPositions, const double* bgLog,
                               int32_t order, const float* rprobs, const float* rconds, int32_t part, int64_t* info, double* pvalues, double* probs,
                               int32_t* classScore, int32_t* classCount, int64_t* classIds) {
	NullModelView nm;
	nm.order = order;
	nm.rprobs = rprobs;
	nm.rconds = rconds;
	PvalueTables T;
	build_pvalue_tables(pwm, cols, ncols, posSetSize, totalPositions, bgLog, nm, T);
	if (T.error) return T.error;
	info[0] = T.L; info[1] = T.maxk; info[2] = T.nparts;
	for (int p = 0; p < 2; ++p) {
		const bool have = p < T.nparts;
		info[3 + p] = have ? (static_cast<int64_t>(1) << (2 * T.part[p].len)) : 0;
		info[5 + p] = have ? static_cast<int64_t>(T.part[p].classScore.size()) : 0;
		info[7 + p] = have ? static_cast<int64_t>(T.part[p].ids.size()) : 0;
	}
	if (part < 0 || part >= T.nparts) return 0;
	const TablePart& P = T.part[part];
	const size_t size = static_cast<size_t>(1) << (2 * P.len);
	if (pvalues) std::fill(pvalues, pvalues + size, 0.0);
	if (probs) std::fill(probs, probs + size, 0.0);
	for (size_t i = 0; i < P.ids.size(); ++i) {
		if (pvalues) pvalues[P.ids[i]] = P.classPval[P.cls[i]];
		if (probs && !P.prob.empty()) probs[P.ids[i]] = P.prob[i];
		if (classIds) classIds[i] = P.ids[i];
	}
	for (size_t c = 0; c < P.classScore.size(); ++c) {
		if (classScore) classScore[c] = P.classScore[c];
		if (classCount) classCount[c] = P.classCount[c];
	}
	return 0;
}

extern "C" int xxm_host_best_subset(const float* score, const int32_t* seq, int32_t nbSites, int32_t posSetSize, int32_t maxSeqCount, int32_t maxSites,
                                    int32_t nent, int32_t maxMotifsPerSequence, int32_t oneOccurrence, double plusFrac, int32_t* bestMotifNb,
                                    double* bestPval) {
	LogSum ls;
	const int64_t limit = static_cast<int64_t>(posSetSize) * maxSeqCount;
	SubsetChoice c = best_subset_from_sorted(ls, score, seq, nbSites, limit, maxSites, nent, maxMotifsPerSequence, oneOccurrence != 0, plusFrac);
	// cross-check with the run-length formulation used by the histogram fast path
	std::vector<ScoreRun> runs;
	for (int32_t i = 0; i < nbSites; ++i) {
		if (!runs.empty() && runs.back().score == score[i]) runs.back().count++;
		else runs.push_back(ScoreRun{score[i], 1});
	}
	SubsetChoice r = best_subset_from_runs(ls, runs, nbSites, limit, maxSites, oneOccurrence != 0, plusFrac);
	*bestMotifNb = c.bestMotifNb;
	*bestPval = c.bestPval;
	std::vector<int32_t> per(static_cast<size_t>(nent) + 1, 0);
	bool caps = false;
	for (int32_t i = 0; i < nbSites; ++i) if (++per[seq[i]] > maxMotifsPerSequence) caps = true;
	if (!caps && plusFrac == 0.0 && (r.bestMotifNb != c.bestMotifNb || r.bestPval != c.bestPval)) return 3;
	return 0;
}

// ---- split-phase subset choice for sharded sequence sets (see xxm_gpu_part3.inc) ------------------------------------
static LogSum g_logsum;

extern "C" int xxm_host_choose_subset(const float* scores, const int64_t* counts, int32_t nruns, int64_t nbSites, int64_t limit, int64_t maxSites,
                                      int32_t oneOccurrence, double plusFrac, int64_t* bestMotifNb, double* bestPval, float* lastScore) {
	*bestMotifNb = 0;
	*bestPval = 1.0;
	*lastScore = 0.f;
	std::vector<ScoreRun> runs(static_cast<size_t>(std::max(0, nruns)));
	for (int i = 0; i < nruns; ++i) runs[i] = ScoreRun{scores[i], counts[i]};
	if (runs.empty() || runs[0].score == 1.0f) return 0;
	g_logsum.reserve(std::max<int64_t>(1, maxSites));
	SubsetChoice c = best_subset_from_runs(g_logsum, runs, nbSites, limit, maxSites, oneOccurrence != 0, plusFrac);
	*bestMotifNb = c.bestMotifNb;
	*bestPval = c.bestPval;
	*lastScore = c.lastScore;
	return 0;
}

extern "C" int xxm_host_build_pwm(const double* counts31x5, double pseudoTimesK, const double* bg5, double* pwm31x5) {
	build_pwm_from_counts(counts31x5, pseudoTimesK, bg5, pwm31x5);
	return 0;
}

extern "C" int xxm_host_pair_table(const uint32_t* ids, int64_t nids, const int32_t* coloff, int32_t len0, uint32_t* table1024, int32_t* levels) {
	int off[32];
	if (len0 < 1 || len0 > 13) return 2;
	for (int i = 0; i < len0; ++i) off[i] = coloff[i];
	int lv = 0;
	const bool ok = build_pair_table(ids, static_cast<size_t>(nids), off, len0, table1024, &lv);
	*levels = lv;
	return ok ? 0 : 1;
}

extern "C" int xxm_host_pack_seqset(int32_t nent, const int32_t* n, const uint8_t* codes, int32_t revcomp, int32_t threads, uint64_t* codes64,
                                    uint32_t* nmask32, int32_t* notX) {
	uint64_t words = 0, rawTotal = 0;
	for (int32_t i = 0; i < nent; ++i) { words += stored_words(static_cast<uint32_t>(n[i]), revcomp != 0); rawTotal += static_cast<uint32_t>(n[i]); }
	uint32_t bad = 0;
	if (threads <= 1 || nent < 2 * threads) {
		bad = pack_range(n, 0, static_cast<uint32_t>(nent), codes, rawTotal, revcomp != 0, codes64, nmask32, notX);
	} else {
		// contiguous blocks of sequences, one per thread, packed concurrently into disjoint word ranges (what xxm_load_seqset
		// does chunk by chunk)
		std::vector<uint32_t> s0(static_cast<size_t>(threads) + 1), badOf(threads, 0);
		std::vector<uint64_t> b0(static_cast<size_t>(threads) + 1), w0(static_cast<size_t>(threads) + 1);
		uint64_t b = 0, w = 0;
		int t = 0;
		for (int32_t i = 0; i < nent; ++i) {
			while (t <= threads - 1 && static_cast<int64_t>(i) == static_cast<int64_t>(nent) * t / threads) { s0[t] = static_cast<uint32_t>(i); b0[t] = b; w0[t] = w; ++t; }
			b += static_cast<uint32_t>(n[i]);
			w += stored_words(static_cast<uint32_t>(n[i]), revcomp != 0);
		}
		s0[threads] = static_cast<uint32_t>(nent); b0[threads] = b; w0[threads] = w;
		std::vector<std::thread> pool;
		for (int k = 0; k < threads; ++k)
			pool.emplace_back([&, k] {
				badOf[k] = pack_range(n, s0[k], s0[k + 1], codes + b0[k], rawTotal - b0[k], revcomp != 0, codes64 + w0[k], nmask32 + w0[k], notX + s0[k]);
			});
		for (auto& th : pool) th.join();
		for (int k = 0; k < threads && !bad; ++k) bad = badOf[k];
	}
	for (int k = 0; k < 2; ++k) { codes64[words + k] = 0; nmask32[words + k] = 0xFFFFFFFFu; }
	return static_cast<int>(bad);
}
