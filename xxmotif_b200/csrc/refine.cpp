// The callers of the hot path, re-stated above the C ABI ("next" rows of section 8):
//   IterateMotif::optimizeMotif        refinementPhase/Iterate_Motif.cpp:23-37
//   IterateMotif::maximizeMotifLength  refinementPhase/Iterate_Motif.cpp:40-149  (+ fillTmpColumn_noGaps, Iterate_Motif.h:46-99)
//   the per-motif loop of MotifContainer::iterate   refinementPhase/MotifContainer.cpp:216-270
//   Motif::offset                      refinementPhase/Motif.h:648-673
// Everything data-sized happens behind xxm_* calls; this file only sequences them and keeps the motif state.
#include "../../include/xxmotif_gpu.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Cols {
	std::vector<int32_t> c;
	int front() const { return c.front(); }
	int back() const { return c.back(); }
	int span() const { return c.back() - c.front() + 1; }
};

// fillTmpColumn_noGaps (Iterate_Motif.h:46-99): false = skip this (front, back) combination
bool variant_columns(const Cols& motif, int front, int back, int minMatch, int maxMatch, Cols& tmp) {
	tmp = motif;
	const int first = motif.front(), last = motif.back();
	if (front == 0 && back == 0) {
	} else if (front != 0 && back == 0) {
		if (front < 0) { for (int i = first - 1; i > first - 1 + front; --i) tmp.c.insert(tmp.c.begin(), i); }
		else { for (int i = 0; i < front && !tmp.c.empty(); ++i) tmp.c.erase(tmp.c.begin()); }
	} else if (front == 0 && back != 0) {
		if (back < 0) { for (int i = 0; i > back && !tmp.c.empty(); --i) tmp.c.pop_back(); }
		else { for (int i = last + 1; i < last + 1 + back; ++i) tmp.c.push_back(i); }
	} else if (front == -1 && back == -1) {
		tmp.c.insert(tmp.c.begin(), first - 1);
		tmp.c.pop_back();
	} else if (front == 1 && back == 1) {
		tmp.c.erase(tmp.c.begin());
		tmp.c.push_back(last + 1);
	} else {
		return false;
	}
	if (static_cast<int>(tmp.c.size()) < minMatch) return false;
	if (static_cast<int>(tmp.c.size()) > maxMatch) return false;
	if (tmp.c.empty() || tmp.front() < 1 || tmp.back() > XXM_PWM_LENGTH) return false;  // would leave the 30-column PWM (undefined in the reference)
	return true;
}

Cols cols_of(const xxm_motif* m) {
	Cols c;
	c.c.assign(m->cols, m->cols + m->ncols);
	return c;
}
void set_cols(xxm_motif* m, const Cols& c) {
	m->ncols = static_cast<int32_t>(c.c.size());
	std::copy(c.c.begin(), c.c.end(), m->cols);
}

int set_positional_pval(xxm_ctx* ctx, xxm_motif* m, const xxm_refine_params* P) {
	const Cols c = cols_of(m);
	double p = 0;
	(void) P;
	const int rc = xxm_positional_pval(ctx, m->enr[0], c.front(), c.span(), m->sitePos, m->nsites, &p);
	if (rc) return rc;
	m->posPval = p;
	return 0;
}

// the end of maximizeMotifLength (Iterate_Motif.cpp:129-148): the winning variant becomes the motif, re-centred in the 30 columns
int finish_maximize(xxm_ctx* ctx, xxm_motif* m, const xxm_refine_params* P, const Cols& bestElongation, double bestPval, const std::vector<double>& bestElong,
                    const std::vector<int32_t>& realSeq, const std::vector<int32_t>& realPos) {
	m->pval = bestPval;
	set_cols(m, bestElongation);
	const int newFirst = (XXM_PWM_LENGTH - bestElongation.span()) / 2 + 1;
	const int offset = newFirst - bestElongation.front();
	if (static_cast<int32_t>(realSeq.size()) > m->siteCap) return 2;
	m->nsites = static_cast<int32_t>(realSeq.size());
	for (int32_t i = 0; i < m->nsites; ++i) { m->siteSeq[i] = realSeq[i]; m->sitePos[i] = realPos[i] - offset; }
	for (int32_t col : bestElongation.c)
		for (int a = 1; a <= 4; ++a) m->pwm[col * 5 + a] = bestElong[col * 5 + a];
	if (offset != 0) {  // Motif::offset, Motif.h:648-673
		m->enr[0] -= offset; m->enr[1] -= offset; m->enr[2] -= offset;
		if (offset > 0) {
			for (auto it = bestElongation.c.rbegin(); it != bestElongation.c.rend(); ++it)
				for (int a = 1; a <= 4; ++a) m->pwm[(*it + offset) * 5 + a] = m->pwm[*it * 5 + a];
		} else {
			for (int32_t col : bestElongation.c)
				for (int a = 1; a <= 4; ++a) m->pwm[(col + offset) * 5 + a] = m->pwm[col * 5 + a];
		}
		for (int32_t i = 0; i < m->ncols; ++i) m->cols[i] += offset;
	}
	return set_positional_pval(ctx, m, P);
}

}  // namespace

extern "C" int xxm_optimize_motif(xxm_ctx* ctx, xxm_motif* m, const xxm_refine_params* P) {
	for (int iter = 0; iter < 2; ++iter) {
		if (xxm_reinitialize(ctx, m->pwm, m->cols, m->ncols, m->posSetSize, P->totalPositions)) return 1;
		int32_t nb = 0, nbest = 0, mx = 0;
		if (xxm_scan_full(ctx, m->posSetSize, m->enr[1], m->enr[2], &nb, &nbest, &mx)) return 1;
		int32_t upd = 0, bestNb = 0;
		double bestPval = 0;
		if (xxm_update_pwm_with_best_subset(ctx, XXM_LIST_SORTED_SITES, m->pwm, m->cols[0], m->posSetSize, 0, P->pseudo, &upd, &bestNb, &bestPval)) return 1;
	}
	return set_positional_pval(ctx, m, P);
}

extern "C" int xxm_maximize_motif_length(xxm_ctx* ctx, xxm_motif* m, const xxm_refine_params* P, int32_t lengthCorrection, int32_t* noChangeOut) {
	const Cols motifCols = cols_of(m);
	int bestMotifNb = 0;
	bool noChange = false;
	double bestPval = DBL_MAX;
	Cols bestElongation = motifCols;
	std::vector<double> elong(31 * 5, 0.0), bestElong(31 * 5, 0.0);
	if (P->elongPWM) std::memcpy(elong.data(), P->elongPWM, sizeof(double) * 31 * 5);
	std::vector<int32_t> realSeq, realPos;
	const int lo = lengthCorrection ? -2 : 0, hi = lengthCorrection ? 2 : 0;
	for (int front = lo; front <= hi; ++front) {
		for (int back = lo; back <= hi; ++back) {
			Cols tmp;
			if (!variant_columns(motifCols, front, back, P->minMatchPositions, P->maxMatchPositions, tmp)) continue;
			int32_t motifNb = 0;
			double overrepPval = 1e100;
			for (int iter = 0; iter < 3; ++iter) {
				const double* pwm = (iter == 0) ? m->pwm : elong.data();
				if (xxm_reinitialize(ctx, pwm, tmp.c.data(), static_cast<int32_t>(tmp.c.size()), m->posSetSize, P->totalPositions)) return 1;
				int32_t nb = 0, mx = 0;
				if (xxm_rescore(ctx, motifCols.front(), m->posSetSize, m->enr[1], m->enr[2], 0, &nb, &mx)) return 1;
				int32_t upd = 0, nbOut = 0;
				double pvOut = 0;
				std::vector<double> work(elong);
				if (xxm_update_pwm_with_best_subset(ctx, XXM_LIST_SORTED_SITES_ELONG, work.data(), tmp.front(), m->posSetSize, iter == 0, P->pseudo, &upd, &nbOut,
				                                    &pvOut))
					return 1;
				if (upd) { elong = work; motifNb = nbOut; overrepPval = pvOut; }
				if (motifNb == 0) break;
			}
			if (motifNb == 0) continue;
			const double newPval = overrepPval + tmp.span() * std::log(P->neff_pwm);  // calculate_log_bonferonni, utils.h:39-56
			if (newPval < bestPval) {
				bestPval = newPval;
				bestElongation = tmp;
				bestMotifNb = motifNb;
				noChange = (front == 0 && back == 0);
				for (int32_t col : tmp.c)
					for (int a = 1; a <= 4; ++a) bestElong[col * 5 + a] = elong[col * 5 + a];
				// fill_startPosList_with_elongSites(realStartPosList, bestMotifNb, tmpColumns) (Iterate_Motif.cpp:117-125): the list of
				// an earlier best variant is overwritten by the next one, so the device list is only set aside here and
				// materialised once, below, for the variant that won
				if (xxm_elong_keep(ctx)) return 1;
			}
		}
	}
	if (P->elongPWM) std::memcpy(P->elongPWM, elong.data(), sizeof(double) * 31 * 5);
	*noChangeOut = noChange;
	if (bestMotifNb == 0) return 0;
	{
		realSeq.assign(static_cast<size_t>(bestMotifNb), 0);
		realPos.assign(static_cast<size_t>(bestMotifNb), 0);
		int32_t n = 0;
		double worst = 1;
		if (xxm_startposlist_from_kept(ctx, bestMotifNb, bestElongation.front(), realSeq.data(), realPos.data(), bestMotifNb, &n, &worst)) return 1;
		realSeq.resize(static_cast<size_t>(n));
		realPos.resize(static_cast<size_t>(n));
		m->worstScore = worst;
	}
	return finish_maximize(ctx, m, P, bestElongation, bestPval, bestElong, realSeq, realPos);
}

// ---- the column variants of one motif several at a time (variant half) ---------------------------------------
// One variant = the three inner iterations of Iterate_Motif.cpp:73-90 under its column set.  Iteration 0 reads the motif's PWM,
// iterations 1 and 2 read elongPWM -- which iteration 0 has just rewritten in all 30 columns (updatePWM_OOL recounts every column,
// Motif.cpp:617-651), or the variant is dropped (motifNb stays 0, :88).  So a variant depends on the motif's PWM, its columns and
// the candidate list of the last full scan, and on no other variant; only the CHOICE among them (strict <, in loop order, :103)
// is sequential, and it is made after all of them are known.
namespace {

struct VariantJob {
	int front = 0, back = 0;
	Cols cols;
	// results
	int32_t motifNb = 0;
	double overrepPval = 1e100;
	bool updated = false;          // some iteration rewrote elongPWM
	std::vector<double> elong;     // elongPWM as this variant leaves it
	int ctxIndex = -1;
	int rc = 0;
};

int evaluate_variant(xxm_ctx* ctx, const xxm_motif* m, int firstMotifColumn, const xxm_refine_params* P, const std::vector<double>& elongIn, VariantJob& J) {
	J.elong = elongIn;
	for (int iter = 0; iter < 3; ++iter) {
		const double* pwm = (iter == 0) ? m->pwm : J.elong.data();
		if (xxm_reinitialize(ctx, pwm, J.cols.c.data(), static_cast<int32_t>(J.cols.c.size()), m->posSetSize, P->totalPositions)) return 1;
		int32_t nb = 0, mx = 0;
		if (xxm_rescore(ctx, firstMotifColumn, m->posSetSize, m->enr[1], m->enr[2], 0, &nb, &mx)) return 1;
		int32_t upd = 0, nbOut = 0;
		double pvOut = 0;
		std::vector<double> work(J.elong);
		if (xxm_update_pwm_with_best_subset(ctx, XXM_LIST_SORTED_SITES_ELONG, work.data(), J.cols.front(), m->posSetSize, iter == 0, P->pseudo, &upd, &nbOut, &pvOut))
			return 1;
		if (upd) { J.elong = work; J.motifNb = nbOut; J.overrepPval = pvOut; J.updated = true; }
		if (J.motifNb == 0) break;
	}
	return 0;
}

}  // namespace

extern "C" int xxm_maximize_motif_length_on(xxm_ctx* const* ctxs, int32_t nctx, xxm_motif* m, const xxm_refine_params* P, int32_t lengthCorrection,
                                            int32_t* noChangeOut) {
	if (nctx < 1 || !ctxs || !ctxs[0]) return 1;
	if (nctx == 1) return xxm_maximize_motif_length(ctxs[0], m, P, lengthCorrection, noChangeOut);
	xxm_ctx* owner = ctxs[0];
	const Cols motifCols = cols_of(m);
	std::vector<double> elong0(31 * 5, 0.0);
	if (P->elongPWM) std::memcpy(elong0.data(), P->elongPWM, sizeof(double) * 31 * 5);
	// the variants in the order of the reference's loops
	std::vector<VariantJob> jobs;
	const int lo = lengthCorrection ? -2 : 0, hi = lengthCorrection ? 2 : 0;
	for (int front = lo; front <= hi; ++front)
		for (int back = lo; back <= hi; ++back) {
			VariantJob J;
			if (!variant_columns(motifCols, front, back, P->minMatchPositions, P->maxMatchPositions, J.cols)) continue;
			J.front = front; J.back = back;
			jobs.push_back(std::move(J));
		}
	const int32_t nwork = std::min<int32_t>(nctx, std::max<int32_t>(1, static_cast<int32_t>(jobs.size())));
	for (int32_t c = 1; c < nwork; ++c)
		if (xxm_share_candidates(ctxs[c], owner)) return 1;
	// every context takes the next variant nobody has taken (so its own variants come in loop order) and sets the list of its best
	// one so far aside, by the reference's rule
	std::atomic<int> next{0};
	std::atomic<bool> stop{false};
	std::vector<int> localBest(static_cast<size_t>(nwork), -1);
	auto work = [&](int32_t c) {
		double best = DBL_MAX;
		for (;;) {
			if (stop.load(std::memory_order_relaxed)) return;
			const int i = next.fetch_add(1);
			if (i >= static_cast<int>(jobs.size())) return;
			VariantJob& J = jobs[static_cast<size_t>(i)];
			J.ctxIndex = c;
			J.rc = evaluate_variant(ctxs[c], m, motifCols.front(), P, elong0, J);
			if (J.rc) { stop.store(true); return; }
			if (J.motifNb == 0) continue;
			const double newPval = J.overrepPval + J.cols.span() * std::log(P->neff_pwm);
			if (newPval < best) {
				best = newPval;
				localBest[static_cast<size_t>(c)] = i;
				if (xxm_elong_keep(ctxs[c])) { J.rc = 1; stop.store(true); return; }
			}
		}
	};
	{
		std::vector<std::thread> pool;
		pool.reserve(static_cast<size_t>(nwork));
		try {
			for (int32_t c = 1; c < nwork; ++c) pool.emplace_back(work, c);
		} catch (...) {}  // no more threads to be had: the contexts that did start share the variants
		work(0);
		for (std::thread& t : pool) t.join();
	}
	for (const VariantJob& J : jobs)
		if (J.rc) return J.rc;
	// the reference's choice, in loop order
	int bestMotifNb = 0, winner = -1;
	bool noChange = false;
	double bestPval = DBL_MAX;
	std::vector<double> elong(elong0), bestElong(31 * 5, 0.0);
	for (size_t i = 0; i < jobs.size(); ++i) {
		const VariantJob& J = jobs[i];
		if (J.updated) elong = J.elong;  // elongPWM as the serial loop would hold it after this variant
		if (J.motifNb == 0) continue;
		const double newPval = J.overrepPval + J.cols.span() * std::log(P->neff_pwm);  // calculate_log_bonferonni, utils.h:39-56
		if (newPval < bestPval) {
			bestPval = newPval;
			bestMotifNb = J.motifNb;
			winner = static_cast<int>(i);
			noChange = (J.front == 0 && J.back == 0);
			for (int32_t col : J.cols.c)
				for (int a = 1; a <= 4; ++a) bestElong[col * 5 + a] = J.elong[col * 5 + a];
		}
	}
	if (P->elongPWM) std::memcpy(P->elongPWM, elong.data(), sizeof(double) * 31 * 5);
	*noChangeOut = noChange;
	if (bestMotifNb == 0) return 0;
	const VariantJob& W = jobs[static_cast<size_t>(winner)];
	xxm_ctx* wctx = ctxs[W.ctxIndex];
	// the overall winner is also the best of the context that evaluated it (smallest p-value, earliest in loop order among equals),
	// so its list is the one that context has kept
	if (localBest[static_cast<size_t>(W.ctxIndex)] != winner) return 3;
	const Cols& bestElongation = W.cols;
	std::vector<int32_t> realSeq(static_cast<size_t>(bestMotifNb), 0), realPos(static_cast<size_t>(bestMotifNb), 0);
	{
		int32_t n = 0;
		double worst = 1;
		if (xxm_startposlist_from_kept(wctx, bestMotifNb, bestElongation.front(), realSeq.data(), realPos.data(), bestMotifNb, &n, &worst)) return 1;
		realSeq.resize(static_cast<size_t>(n));
		realPos.resize(static_cast<size_t>(n));
		m->worstScore = worst;
	}
	return finish_maximize(owner, m, P, bestElongation, bestPval, bestElong, realSeq, realPos);
}

extern "C" int xxm_iterate_motif(xxm_ctx* ctx, xxm_motif* m, const xxm_refine_params* P, int32_t* rounds) { return xxm_iterate_motif_on(&ctx, 1, m, P, rounds); }

// the per-motif loop on ctxs[0]; with nctx > 1 the column variants of maximizeMotifLength are spread over ctxs[0] and its views
extern "C" int xxm_iterate_motif_on(xxm_ctx* const* ctxs, int32_t nctx, xxm_motif* m, const xxm_refine_params* P, int32_t* rounds) {
	if (nctx < 1 || !ctxs || !ctxs[0]) return 1;
	xxm_ctx* ctx = ctxs[0];
	double oldPval = 1000, bestPval = 1000;
	int loop = 1;
	*rounds = 0;
	std::vector<double> elong(31 * 5, 0.0);  // IterateMotif::elongPWM lives for the whole iterate() call
	xxm_refine_params Q = *P;
	Q.elongPWM = elong.data();
	do {
		oldPval = bestPval;
		int rc = xxm_optimize_motif(ctx, m, &Q);
		if (rc) return rc;
		if (P->localization) {
			const Cols c = cols_of(m);
			if (xxm_update_enriched_region(ctx, c.front(), c.span(), m->enr)) return 1;
		}
		int32_t noChange = 0;
		rc = xxm_maximize_motif_length_on(ctxs, nctx, m, &Q, P->maximizeMotifLength, &noChange);
		if (rc) return rc;
		bestPval = m->pval;
		++*rounds;
		if (loop > 0 && bestPval >= oldPval) loop = 0;
	} while (loop > 0);
	// m->getStartPosList().sort(StartPos::cmpStartPos), MotifContainer.cpp:262: by (sequence, position).  A stable LSD radix
	// sort on the 64-bit key (positions may be negative: biased), 16 bits per pass, only the passes whose digit varies
	{
		const size_t n = static_cast<size_t>(m->nsites);
		std::vector<uint64_t> a(n), b(n);
		uint64_t orAll = 0, andAll = ~0ull;
		for (size_t i = 0; i < n; ++i) {
			a[i] = (static_cast<uint64_t>(static_cast<uint32_t>(m->siteSeq[i])) << 32) | (static_cast<uint32_t>(m->sitePos[i]) ^ 0x80000000u);
			orAll |= a[i]; andAll &= a[i];
		}
		std::vector<uint32_t> cnt(65537);
		for (int pass = 0; pass < 4; ++pass) {
			const int sh = 16 * pass;
			if ((((orAll ^ andAll) >> sh) & 0xFFFFull) == 0) continue;  // the same digit everywhere
			std::fill(cnt.begin(), cnt.end(), 0u);
			for (size_t i = 0; i < n; ++i) cnt[((a[i] >> sh) & 0xFFFFull) + 1]++;
			for (int d = 0; d < 65536; ++d) cnt[d + 1] += cnt[d];
			for (size_t i = 0; i < n; ++i) b[cnt[(a[i] >> sh) & 0xFFFFull]++] = a[i];
			a.swap(b);
		}
		for (size_t i = 0; i < n; ++i) { m->siteSeq[i] = static_cast<int32_t>(a[i] >> 32); m->sitePos[i] = static_cast<int32_t>(static_cast<uint32_t>(a[i]) ^ 0x80000000u); }
	}
	return 0;
}

// The motifs of MotifContainer::iterate (MotifContainer.cpp:216-270), several at a time: one host thread per context, every
// thread takes the next motif that nobody has taken (the list is in the caller's order, so with one context this is the
// reference's loop).  A motif runs from start to end on one context: its result does not depend on the others.
extern "C" int xxm_iterate_motifs(xxm_ctx* const* ctxs, int32_t nctx, xxm_motif* motifs, int32_t nmotifs, const xxm_refine_params* params, int32_t nparams,
                                  int32_t* rounds, int32_t* failed, int32_t* failedCtx) {
	if (failed) *failed = -1;
	if (failedCtx) *failedCtx = -1;
	if (nctx < 1 || nmotifs < 0 || !ctxs || (nmotifs && (!motifs || !params || !rounds))) return 1;
	if (nparams != 1 && nparams != nmotifs) return 1;
	for (int32_t i = 0; i < nctx; ++i) {
		if (!ctxs[i]) return 1;
		for (int32_t j = 0; j < i; ++j)
			if (ctxs[j] == ctxs[i]) return 1;  // two threads on one context
	}
	std::atomic<int32_t> next{0};
	std::atomic<int32_t> firstBad{-1};
	std::vector<int32_t> code(static_cast<size_t>(nctx), 0), badMotif(static_cast<size_t>(nctx), -1);
	auto work = [&](int32_t c) {
		for (;;) {
			if (firstBad.load(std::memory_order_relaxed) >= 0) return;  // a motif failed: nobody starts another one
			const int32_t i = next.fetch_add(1);
			if (i >= nmotifs) return;
			const int rc = xxm_iterate_motif(ctxs[c], &motifs[i], &params[nparams == 1 ? 0 : i], &rounds[i]);
			if (rc) {
				code[static_cast<size_t>(c)] = rc;
				badMotif[static_cast<size_t>(c)] = i;
				int32_t none = -1;
				firstBad.compare_exchange_strong(none, c);
				return;
			}
		}
	};
	const int32_t nthreads = std::min(nctx, std::max<int32_t>(1, nmotifs));
	std::vector<std::thread> pool;
	pool.reserve(static_cast<size_t>(nthreads));
	try {
		for (int32_t c = 1; c < nthreads; ++c) pool.emplace_back(work, c);
	} catch (...) {}  // no more threads to be had: the contexts that did start share the motifs
	work(0);  // the caller's thread drives the first context
	for (std::thread& t : pool) t.join();
	// report the failure with the smallest motif index (several contexts may have failed before they saw the flag)
	int32_t bad = -1;
	for (int32_t c = 0; c < nctx; ++c)
		if (code[static_cast<size_t>(c)] && (bad < 0 || badMotif[static_cast<size_t>(c)] < badMotif[static_cast<size_t>(bad)])) bad = c;
	if (bad < 0) return 0;
	if (failed) *failed = badMotif[static_cast<size_t>(bad)];
	if (failedCtx) *failedCtx = bad;
	return code[static_cast<size_t>(bad)];
}
