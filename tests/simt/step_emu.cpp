// TEST INFRASTRUCTURE ONLY (see simt_emu.h): the two kernels that follow the scan in a refinement step, from their CUDA sources
// under the SIMT emulator:
//   k_select_cls_fwd (xxm_select3.cuh)  updateMultOcc on class indices        vs the reference's greedy, StartPosUtils.h:12-28
//   k_count_pwm3     (xxm_select2.cuh)  30-column context counts of the sites vs Motif::updateCounts, refinementPhase/Motif.cpp:617-672
// on seeded candidate lists (clustered positions, ties of float-rounded p-values, N residues, sites at both sequence ends).
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
#include "xxm_select2.cuh"
#include "xxm_select3.cuh"
#include "xxm_select4.cuh"

#include "host_pack.h"

namespace xxm {
alignas(16) uint8_t smem3[1 << 16];
alignas(16) uint8_t smem4[65536 + 65536];
alignas(16) uint8_t smem5[kPairTabWords * 4];
alignas(16) uint8_t smemM[1 << 16];
uint32_t s_hist[1 << 14];
uint32_t s_hs[1 << 14];
}  // namespace xxm

using namespace xxm;

namespace {

struct Case {
	std::string name;
	int nseq, lenLo, lenHi;
	double hitFrac;   // candidates per window
	int span;         // motif length (overlap distance)
	uint32_t nval;    // score classes
	uint32_t nSite;   // classes [0, nSite) are sites
	double nFrac;
	uint32_t seed;
};

int run_case(const Case& c) {
	std::mt19937 rng(c.seed);
	// ---- sequences
	std::vector<int32_t> n(c.nseq);
	for (auto& v : n) v = c.lenLo + static_cast<int>(rng() % static_cast<uint32_t>(c.lenHi - c.lenLo + 1));
	uint64_t tot = 0;
	for (int v : n) tot += static_cast<uint64_t>(v);
	std::vector<uint8_t> raw(tot + 64);
	for (uint64_t i = 0; i < tot; ++i) raw[i] = static_cast<uint8_t>((rng() % 1000 < c.nFrac * 1000) ? 0 : 1 + rng() % 4);
	std::vector<uint32_t> wordOff(c.nseq + 1);
	uint64_t w = 0;
	bool uniform = true;
	for (int i = 0; i < c.nseq; ++i) { wordOff[i] = static_cast<uint32_t>(w); w += stored_words(static_cast<uint32_t>(n[i]), false); uniform = uniform && n[i] == n[0]; }
	wordOff[c.nseq] = static_cast<uint32_t>(w);
	std::vector<uint64_t> codes(w + 2, 0ull);
	std::vector<uint32_t> nmask(w + 2, 0xFFFFFFFFu);
	std::vector<int32_t> notX(c.nseq);
	if (pack_range(n.data(), 0, static_cast<uint32_t>(c.nseq), raw.data(), tot, false, codes.data(), nmask.data(), notX.data())) return 2;
	SeqStream S{};
	S.codes = codes.data(); S.nmask = nmask.data(); S.wordOff = wordOff.data(); S.len = n.data(); S.nwords = static_cast<uint32_t>(w); S.nent = c.nseq;
	S.uniformWps = uniform ? wordOff[1] - wordOff[0] : 0; S.uniformLen = uniform ? n[0] : 0;
	// ---- candidates in scan order: clustered (a hit raises the chance of a hit two positions later, like TATATA)
	std::vector<uint32_t> seq, pos, vidx;
	for (int q = 0; q < c.nseq; ++q) {
		int boost = 0;
		for (int p = 1; p + c.span - 1 <= n[q]; ++p) {
			const double pr = c.hitFrac * (boost > 0 ? 12.0 : 1.0);
			if (boost > 0) --boost;
			if ((rng() % 100000) < pr * 100000) {
				seq.push_back(static_cast<uint32_t>(q) + 1);
				pos.push_back(static_cast<uint32_t>(p));
				// mostly weak classes, sometimes a class that is not a site, sometimes "none"
				uint32_t v = static_cast<uint32_t>(std::min<double>(c.nval - 1, std::abs(std::normal_distribution<double>(0.6 * c.nval, 0.3 * c.nval)(rng))));
				if (rng() % 50 == 0) v = kNone;
				vidx.push_back(v);
				boost = 6;
			}
		}
	}
	const uint32_t ncand = static_cast<uint32_t>(seq.size());
	// ---- a finalPval table that is non-decreasing in the class index, with neighbours that round to the same float
	std::vector<double> fin(c.nval);
	double v = 1e-9;
	for (uint32_t k = 0; k < c.nval; ++k) {
		v *= (rng() % 4 == 0) ? (1.0 + 1e-9) : (1.0 + (rng() % 1000) / 900.0);  // some steps are below float resolution
		fin[k] = (k < c.nSite) ? std::min(v, 0.999999) : 1.0;
	}
	for (uint32_t k = 1; k < c.nval; ++k) fin[k] = std::max(fin[k], fin[k - 1]);
	std::vector<uint16_t> thr(c.nval);
	for (uint32_t k = 0; k < c.nval; ++k) {
		const double f = static_cast<double>(static_cast<float>(fin[k]));
		thr[k] = static_cast<uint16_t>(std::lower_bound(fin.begin(), fin.end(), f) - fin.begin());  // stage_class_order, xxm_gpu.cu
	}
	const uint32_t nSite = static_cast<uint32_t>(std::lower_bound(fin.begin(), fin.end(), 1.0) - fin.begin());
	// ---- the reference's greedy (forward rule only): StartPosUtils.h:12-28, called for windows with finalPval < 1
	std::vector<uint8_t> wantKeep(ncand, 0);
	{
		std::vector<uint32_t> kept;  // indices of kept entries
		for (uint32_t i = 0; i < ncand; ++i) {
			if (vidx[i] == kNone || vidx[i] >= c.nval) continue;
			const double pv = fin[vidx[i]];
			if (!(pv < 1.0)) continue;
			if (!kept.empty()) {
				const uint32_t l = kept.back();
				if (seq[l] == seq[i] && pos[l] + static_cast<uint32_t>(c.span) >= pos[i]) {
					if (static_cast<double>(static_cast<float>(fin[vidx[l]])) <= pv) continue;
					kept.pop_back();
				}
			}
			kept.push_back(i);
		}
		for (uint32_t i : kept) wantKeep[i] = 1;
	}
	// ---- the kernel
	std::vector<uint8_t> keep(ncand + 16, 0xEE);
	std::vector<uint16_t> kcls(ncand + 16, 0xEEEE);
	std::vector<uint32_t> hist(c.nval, 0u), counters(8, 0u);
	SelectClsIO f{};
	f.seq = seq.data(); f.pos = pos.data(); f.vidx = vidx.data(); f.n = ncand; f.nDev = nullptr; f.length = c.span;
	f.clsThr = thr.data(); f.nSite = nSite; f.nval = c.nval; f.keep = keep.data(); f.kcls = kcls.data(); f.counters = counters.data(); f.hist = hist.data();
	const unsigned grid = std::max(1u, std::min(3u, (ncand + 256 * kSelChunk - 1) / (256 * kSelChunk)));
	emu::launch(grid, 32 * kSel3Warps, [&] { k_select_cls_fwd(f); });
	int bad = 0;
	uint32_t nkept = 0;
	std::vector<uint32_t> wantHist(c.nval, 0u);
	for (uint32_t i = 0; i < ncand; ++i) {
		if (wantKeep[i]) { ++nkept; ++wantHist[vidx[i]]; }
		const uint16_t wc = wantKeep[i] ? static_cast<uint16_t>(vidx[i]) : static_cast<uint16_t>(0xFFFF);
		if (keep[i] != wantKeep[i] || kcls[i] != wc) {
			if (!bad) std::fprintf(stderr, "FAIL %s: entry %u (seq %u pos %u cls %u): keep %u kcls %u, expected %u %u\n", c.name.c_str(), i, seq[i], pos[i], vidx[i], keep[i], kcls[i], wantKeep[i], wc);
			++bad;
		}
	}
	if (counters[0] != nkept) { std::fprintf(stderr, "FAIL %s: %u kept sites counted, expected %u\n", c.name.c_str(), counters[0], nkept); ++bad; }
	if (hist != wantHist) { std::fprintf(stderr, "FAIL %s: class histogram differs\n", c.name.c_str()); ++bad; }
	// ---- recount: sites with class <= maxClass
	const uint32_t maxClass = nSite / 2;
	const int firstColumn = 12, shift = 0;
	std::vector<unsigned long long> counts(31 * 5, 0ull), wantCounts(31 * 5, 0ull);
	for (uint32_t i = 0; i < ncand; ++i) {
		if (!wantKeep[i] || vidx[i] > maxClass) continue;
		const int q = static_cast<int>(seq[i]) - 1;
		const int g0 = static_cast<int>(pos[i]) + shift - firstColumn;  // 0-based offset of PWM column 1 inside the sequence
		for (int j = 1; j <= 30; ++j) {
			const int p = g0 + j - 1;
			if (p < 0 || p >= n[q]) continue;
			const uint32_t ww = wordOff[q] + static_cast<uint32_t>(p >> 5);
			if ((nmask[ww] >> (p & 31)) & 1u) continue;
			const uint32_t b = static_cast<uint32_t>(codes[ww] >> (2 * (p & 31))) & 3u;
			++wantCounts[j * 5 + 1 + b];
			++wantCounts[j * 5];
		}
	}
	CountIO io{};
	io.seq = seq.data(); io.pos = pos.data(); io.shift = shift; io.keep = keep.data(); io.n = ncand; io.kcls = kcls.data(); io.maxClass = maxClass;
	io.firstColumn = firstColumn; io.counts = counts.data();
	emu::launch(2, 256, [&] { k_count_pwm3(S, io); });
	if (counts != wantCounts) {
		std::fprintf(stderr, "FAIL %s: PWM counts differ\n", c.name.c_str());
		++bad;
	}
	// ---- k_count_pwm2 (the count kernel of float-valued lists, ZOOPS slots and explicit index lists; the chosen entries are parked in a
	// per-warp buffer and counted 32 at a time with ballots) on the same sites, in its three ways of choosing
	{
		auto expected = [&](const std::vector<uint32_t>& chosen) {
			std::vector<unsigned long long> wc(31 * 5, 0ull);
			for (uint32_t i : chosen) {
				const int q = static_cast<int>(seq[i]) - 1;
				const int g0 = static_cast<int>(pos[i]) + shift - firstColumn;
				for (int j = 1; j <= 30; ++j) {
					const int p = g0 + j - 1;
					if (p < 0 || p >= n[q]) continue;
					const uint32_t ww = wordOff[q] + static_cast<uint32_t>(p >> 5);
					if ((nmask[ww] >> (p & 31)) & 1u) continue;
					++wc[j * 5 + 1 + (static_cast<uint32_t>(codes[ww] >> (2 * (p & 31))) & 3u)];
					++wc[j * 5];
				}
			}
			wc[0] = chosen.size();  // row 0 is unused by the PWM and carries the number of chosen sites
			return wc;
		};
		std::vector<float> kscore(ncand + 16, 2.f);
		for (uint32_t i = 0; i < ncand; ++i)
			if (vidx[i] != kNone && vidx[i] < c.nval) kscore[i] = static_cast<float>(fin[vidx[i]]);
		const float thr = static_cast<float>(fin[maxClass]);
		std::vector<uint32_t> byScore, slotIdx, slotChosen;
		for (uint32_t i = 0; i < ncand; ++i)
			if (wantKeep[i] && kscore[i] <= thr) byScore.push_back(i);
		auto run2 = [&](const char* how, CountIO io2, const std::vector<uint32_t>& chosen) {
			std::vector<unsigned long long> got(31 * 5, 0ull);
			io2.seq = seq.data(); io2.pos = pos.data(); io2.shift = shift; io2.firstColumn = firstColumn; io2.counts = got.data();
			emu::launch(2, 256, [&] { k_count_pwm2(S, io2); });
			if (got != expected(chosen)) { std::fprintf(stderr, "FAIL %s: k_count_pwm2 (%s) differs\n", c.name.c_str(), how); ++bad; }
		};
		CountIO a{};  // kept entries with float score <= thr
		a.keep = keep.data(); a.kscore = kscore.data(); a.n = ncand; a.thr = thr;
		run2("keep flags + float scores", a, byScore);
		std::vector<uint32_t> idx(byScore.rbegin(), byScore.rend());  // an explicit list, in another order
		CountIO b{};
		b.idx = idx.data(); b.n = static_cast<uint32_t>(idx.size());
		if (b.n) run2("index list", b, idx);
		std::vector<float> slotScore;  // ZOOPS: one slot per sequence, the first kept entry of the sequence if any
		for (int q = 0, i = 0; q < c.nseq; ++q) {
			uint32_t pick = kNone;
			for (; i < static_cast<int>(ncand) && static_cast<int>(seq[i]) - 1 == q; ++i)
				if (pick == kNone && wantKeep[i]) pick = static_cast<uint32_t>(i);
			slotIdx.push_back(pick);
			slotScore.push_back(pick == kNone ? 1.f : kscore[pick]);
			if (pick != kNone && kscore[pick] <= thr) slotChosen.push_back(pick);
		}
		CountIO z{};
		z.slotIdx = slotIdx.data(); z.slotScore = slotScore.data(); z.n = static_cast<uint32_t>(slotIdx.size()); z.thr = thr;
		run2("ZOOPS slots", z, slotChosen);
	}
	std::printf("%-20s %7u candidates, %6u kept, %3u classes (%u sites): %s\n", c.name.c_str(), ncand, nkept, c.nval, nSite, bad ? "DIFFERENT" : "identical");
	return bad ? 1 : 0;
}

// k_eval_list + k_select_zoops (the re-score of a candidate list, xxm_kernels.cuh) against fill_elongSites_with_possible_startPositions'
// ZOOPS branch (StartPosUpdater-inl.h:167-203): window evaluation per candidate, then per sequence the first candidate whose
// finalPval is below the float-stored best so far
int run_zoops_case(const std::string& name, int nseq, int lenLo, int lenHi, double candFrac, int shift, double nFrac, uint32_t seed) {
	std::mt19937 rng(seed);
	std::vector<int32_t> n(nseq);
	for (auto& v : n) v = lenLo + static_cast<int>(rng() % static_cast<uint32_t>(lenHi - lenLo + 1));
	uint64_t tot = 0;
	for (int v : n) tot += static_cast<uint64_t>(v);
	std::vector<uint8_t> raw(tot + 64);
	for (uint64_t i = 0; i < tot; ++i) raw[i] = static_cast<uint8_t>((rng() % 1000 < nFrac * 1000) ? 0 : 1 + rng() % 4);
	std::vector<uint32_t> wordOff(nseq + 1);
	uint64_t w = 0;
	bool uniform = true;
	for (int i = 0; i < nseq; ++i) { wordOff[i] = static_cast<uint32_t>(w); w += stored_words(static_cast<uint32_t>(n[i]), false); uniform = uniform && n[i] == n[0]; }
	wordOff[nseq] = static_cast<uint32_t>(w);
	std::vector<uint64_t> codes(w + 2, 0ull);
	std::vector<uint32_t> nmask(w + 2, 0xFFFFFFFFu);
	std::vector<int32_t> notX(nseq);
	if (pack_range(n.data(), 0, static_cast<uint32_t>(nseq), raw.data(), tot, false, codes.data(), nmask.data(), notX.data())) return 2;
	SeqStream S{};
	S.codes = codes.data(); S.nmask = nmask.data(); S.wordOff = wordOff.data(); S.len = n.data(); S.nwords = static_cast<uint32_t>(w); S.nent = nseq;
	S.uniformWps = uniform ? wordOff[1] - wordOff[0] : 0; S.uniformLen = uniform ? n[0] : 0;
	// a 7-column table: half of the k-mers, classes by a random additive score
	const int L = 7;
	int sc[8][4];
	for (int k = 0; k < L; ++k) for (int a = 0; a < 4; ++a) sc[k][a] = static_cast<int>(rng() % 1000);
	const uint32_t N = 1u << (2 * L);
	std::vector<std::pair<int, uint32_t>> all(N);
	for (uint32_t id = 0; id < N; ++id) { int v = 0; for (int k = 0; k < L; ++k) v += sc[k][(id >> (2 * k)) & 3u]; all[id] = {-(v / 40), id}; }
	std::sort(all.begin(), all.end());
	std::vector<uint32_t> rank(N, 0u);
	uint32_t cls = 0;
	for (size_t i = 0; i < N / 2; ++i) { if (i && all[i].first != all[i - 1].first) ++cls; rank[all[i].second] = cls + 1; }
	const uint32_t nval = cls + 1;
	std::vector<double> fin(nval), pS(nval, 0.5);
	double v = 1e-7;
	for (uint32_t k = 0; k < nval; ++k) { v *= (rng() % 3 == 0) ? (1.0 + 2e-9) : 1.4; fin[k] = (k + 3 < nval) ? std::min(v, 0.99) : 1.0; }
	Model M{};
	M.L = L; M.maxk = 8; M.nparts = 1; M.span = L; M.contiguous = 1; M.len0 = L;
	for (int k = 0; k < L; ++k) M.coloff[k] = k;
	M.rank0 = rank.data();
	Values V{};
	V.pS = pS.data(); V.finalTab = fin.data(); V.kind = kFinalTable;
	// candidates in (sequence, position) order
	std::vector<uint32_t> seq, pos;
	for (int q = 0; q < nseq; ++q)
		for (int p = 1; p <= n[q]; ++p)
			if ((rng() % 100000) < candFrac * 100000) { seq.push_back(static_cast<uint32_t>(q) + 1); pos.push_back(static_cast<uint32_t>(p)); }
	const uint32_t nc = static_cast<uint32_t>(seq.size());
	// expectation
	std::vector<uint32_t> wantV(nc, kNone), wantIdx(nseq, kNone);
	std::vector<float> wantScore(nseq, 1.0f);
	uint32_t wantFilled = 0;
	for (uint32_t i = 0; i < nc; ++i) {
		const int q = static_cast<int>(seq[i]) - 1, p = static_cast<int>(pos[i]) + shift;
		if (p >= 1 && p + L - 1 <= n[q]) {
			uint32_t id = 0;
			bool dead = false;
			for (int k = 0; k < L; ++k) {
				const int o = p - 1 + k;
				const uint32_t ww = wordOff[q] + static_cast<uint32_t>(o >> 5);
				if ((nmask[ww] >> (o & 31)) & 1u) dead = true;
				id |= (static_cast<uint32_t>(codes[ww] >> (2 * (o & 31))) & 3u) << (2 * k);
			}
			if (!dead && rank[id]) wantV[i] = rank[id] - 1;
		}
		if (wantV[i] == kNone) continue;
		const double f = fin[wantV[i]];
		if (f < static_cast<double>(wantScore[q])) {
			if (wantIdx[q] == kNone) ++wantFilled;
			wantIdx[q] = i;
			wantScore[q] = static_cast<float>(f);
		}
	}
	int bad = 0;
	// k_eval_list, then k_select_zoops on its value indices
	std::vector<uint32_t> vidx2(nc + 1, 0u), slotIdx2(nseq, kNone), counters2(8, 0u);
	std::vector<float> slotScore2(nseq, 1.0f);
	emu::launch((nc + 255) / 256, 256, [&] { k_eval_list(S, M, seq.data(), pos.data(), nc, shift, vidx2.data()); });
	SlotIO io2{};
	io2.seq = seq.data(); io2.pos = pos.data(); io2.n = nc; io2.shift = shift;
	io2.vidx = vidx2.data(); io2.slotIdx = slotIdx2.data(); io2.slotScore = slotScore2.data(); io2.counters = counters2.data();
	emu::launch((nc + kZoopsTile - 1) / kZoopsTile, 256, [&] { k_select_zoops(V, io2); });
	for (uint32_t i = 0; i < nc && !bad; ++i)
		if (vidx2[i] != wantV[i]) { std::fprintf(stderr, "FAIL %s: candidate %u value index %u, expected %u\n", name.c_str(), i, vidx2[i], wantV[i]); ++bad; }
	if (slotIdx2 != wantIdx || slotScore2 != wantScore) { std::fprintf(stderr, "FAIL %s: slots differ\n", name.c_str()); ++bad; }
	if (counters2[0] != wantFilled) { std::fprintf(stderr, "FAIL %s: %u slots filled, expected %u\n", name.c_str(), counters2[0], wantFilled); ++bad; }
	// the serial fallback of the same selection (lists that are not ordered by sequence take it)
	{
		std::vector<uint32_t> slotIdx3(nseq, kNone), counters3(8, 0u);
		std::vector<float> slotScore3(nseq, 1.0f);
		SlotIO io3 = io2;
		io3.slotIdx = slotIdx3.data(); io3.slotScore = slotScore3.data(); io3.counters = counters3.data();
		emu::launch(1, 32, [&] { k_select_zoops_serial(V, io3); });
		if (slotIdx3 != wantIdx || slotScore3 != wantScore || counters3[0] != wantFilled) { std::fprintf(stderr, "FAIL %s: k_select_zoops_serial differs\n", name.c_str()); ++bad; }
	}
	// k_hist_slots: the value-index histogram of the filled slots (what the subset choice of a ZOOPS re-score reads), bins in shared
	// memory, in global memory, and split between the two
	{
		uint32_t nval = 1;
		for (uint32_t i = 0; i < nc; ++i)
			if (wantV[i] != kNone) nval = std::max(nval, wantV[i] + 1);
		std::vector<uint32_t> wantHist(nval, 0u);
		for (int q = 0; q < nseq; ++q)
			if (wantIdx[q] != kNone) ++wantHist[wantV[wantIdx[q]]];
		for (uint32_t shared : {nval, 0u, nval / 2}) {
			std::vector<uint32_t> hist(nval, 0u);
			emu::launch(2, 256, [&] { k_hist_slots(slotIdx2.data(), vidx2.data(), static_cast<uint32_t>(nseq), hist.data(), shared); });
			if (hist != wantHist) { std::fprintf(stderr, "FAIL %s: k_hist_slots (%u bins in shared memory) differs\n", name.c_str(), shared); ++bad; }
		}
	}
	std::printf("%-20s %7u candidates, %6u of %d slots filled, shift %+d: %s\n", name.c_str(), nc, wantFilled, nseq, shift, bad ? "DIFFERENT" : "identical");
	return bad ? 1 : 0;
}

// k_select_mops_fwd (xxm_select2.cuh): updateMultOcc's forward rule on FLOAT-valued lists -- what two-part models, localization and
// re-scored lists take (the finalPval is not a monotone function of a class index there) -- against the reference's greedy,
// StartPosUtils.h:12-28, literally.  kind 0: finalPval from a table over the value index (any order); 1: the localized MOPS formula
// from pS and the positional prior (pValCalculation.h:79); 2: one finalPval per candidate.
int run_fwd_generic_case(const std::string& name, int nseq, int lenLo, int lenHi, double hitFrac, int span, uint32_t nval, int kind, int shift, uint32_t seed) {
	std::mt19937 rng(seed);
	std::vector<int32_t> n(nseq);
	int maxLen = 0;
	for (auto& v : n) { v = lenLo + static_cast<int>(rng() % static_cast<uint32_t>(lenHi - lenLo + 1)); maxLen = std::max(maxLen, v); }
	std::vector<uint32_t> seq, pos, vidx;
	for (int q = 0; q < nseq; ++q) {
		int boost = 0;
		for (int p = std::max(1, 1 - shift); p + span - 1 <= n[q]; ++p) {
			const double pr = hitFrac * (boost > 0 ? 12.0 : 1.0);
			if (boost > 0) --boost;
			if ((rng() % 100000) < pr * 100000) {
				seq.push_back(static_cast<uint32_t>(q) + 1);
				pos.push_back(static_cast<uint32_t>(p));
				vidx.push_back((rng() % 40 == 0) ? kNone : static_cast<uint32_t>(rng() % nval));
				boost = 6;
			}
		}
	}
	const uint32_t ncand = static_cast<uint32_t>(seq.size());
	std::uniform_real_distribution<double> U(0.0, 1.0);
	std::vector<double> pS(nval), finTab(nval), posProb(static_cast<size_t>(maxLen) + 8 + static_cast<size_t>(std::abs(shift))), perCand(std::max(1u, ncand));
	for (uint32_t k = 0; k < nval; ++k) {
		pS[k] = (rng() % 9 == 0) ? 1.0 : std::pow(U(rng), 6);         // 1.0: not a site
		finTab[k] = (rng() % 7 == 0) ? 1.0 : std::pow(U(rng), 5);      // in no particular order; a few values repeat (ties keep the earlier site)
		if (k && rng() % 5 == 0) finTab[k] = finTab[k - 1];
	}
	for (double& v : posProb) v = 0.05 + 0.95 * U(rng);
	for (double& v : perCand) v = (rng() % 6 == 0) ? 1.0 : std::pow(U(rng), 4);
	Values V{};
	V.pS = pS.data(); V.finalTab = finTab.data(); V.posProb = posProb.data(); V.perCand = perCand.data();
	V.kind = kind == 0 ? kFinalTable : kind == 1 ? kFinalLocMops : kFinalPerCand;
	// ---- the reference's greedy: the incoming double finalPval against the float score stored with the kept site
	std::vector<double> fin(std::max(1u, ncand), 1.0);
	for (uint32_t i = 0; i < ncand; ++i)
		if (vidx[i] != kNone) fin[i] = final_of(V, vidx[i], i, static_cast<int>(pos[i]) + shift);
	std::vector<uint8_t> wantKeep(std::max(1u, ncand), 0);
	{
		std::vector<uint32_t> kept;
		for (uint32_t i = 0; i < ncand; ++i) {
			if (!(fin[i] < 1.0)) continue;
			if (!kept.empty()) {
				const uint32_t l = kept.back();
				if (seq[l] == seq[i] && pos[l] + static_cast<uint32_t>(span) >= pos[i]) {
					if (static_cast<double>(static_cast<float>(fin[l])) <= fin[i]) continue;
					kept.pop_back();
				}
			}
			kept.push_back(i);
		}
		for (uint32_t i : kept) wantKeep[i] = 1;
	}
	// ---- the kernel: arrays of exactly n entries, histogram in shared memory and (a second launch) in global memory
	int bad = 0;
	for (int smemHist = 1; smemHist >= 0; --smemHist) {
		std::vector<uint8_t> keep(std::max(1u, ncand), 0xEE);
		std::vector<float> kscore(std::max(1u, ncand), -5.f);
		std::vector<uint16_t> kcls(std::max(1u, ncand), 0xEEEE);
		std::vector<uint32_t> hist(nval, 0u), counters(8, 0u);
		SelectFwdIO f{};
		f.seq = seq.data(); f.pos = pos.data(); f.vidx = vidx.data(); f.n = ncand; f.nDev = nullptr; f.length = span; f.shift = shift;
		f.kscore = kscore.data(); f.keep = keep.data(); f.counters = counters.data();
		const bool wantHist = kind == 0;
		if (wantHist) { f.hist = hist.data(); f.nval = nval; f.smemHist = smemHist; f.kcls = kcls.data(); }
		const unsigned grid = std::max(1u, std::min(3u, (ncand + 256 * kSelChunk - 1) / (256 * kSelChunk)));
		emu::launch(grid, 256, [&] { k_select_mops_fwd(V, f); });
		uint32_t nkept = 0, wrong = 0;
		std::vector<uint32_t> wantHistV(nval, 0u);
		for (uint32_t i = 0; i < ncand; ++i) {
			if (wantKeep[i]) { ++nkept; if (wantHist) ++wantHistV[vidx[i]]; }
			const float ws = static_cast<float>(fin[i]);
			bool ok = keep[i] == wantKeep[i] && !std::memcmp(&kscore[i], &ws, 4);
			if (wantHist) ok = ok && kcls[i] == (wantKeep[i] ? static_cast<uint16_t>(vidx[i]) : static_cast<uint16_t>(0xFFFF));
			wrong += ok ? 0 : 1;
		}
		const bool same = wrong == 0 && counters[0] == nkept && (!wantHist || hist == wantHistV);
		std::printf("%-22s k_select_mops_fwd kind %d%s: %u candidates, %u kept, span %d, shift %d  %s\n", name.c_str(), kind,
		            wantHist ? (smemHist ? ", histogram in shared memory" : ", histogram in global memory") : "", ncand, nkept, span, shift, same ? "identical" : "DIFFERENT");
		bad += same ? 0 : 1;
		if (!wantHist) break;
	}
	return bad;
}

// ordered compaction (k_cmp_count + k_tile_scan + k_cmp_scatter, xxm_kernels.cuh): the selected elements of a (sequence, position,
// value index) list, in order -- selected by a flag byte, or as the heads of the runs of equal keys of a sorted key array --
// against std::copy_if; output arrays of exactly the number selected
int run_compact_case(const std::string& name, uint32_t n, double frac, bool runHeads, uint32_t seed) {
	std::mt19937 rng(seed);
	std::vector<uint32_t> seq(n), pos(n), vidx(n), keys(n);
	std::vector<uint8_t> flag(n);
	uint32_t key = 5;
	for (uint32_t i = 0; i < n; ++i) {
		seq[i] = 1 + i / 7; pos[i] = rng(); vidx[i] = rng() % 1000;
		flag[i] = ((rng() % 100000) < frac * 100000) ? static_cast<uint8_t>(1 + rng() % 3) : 0;  // any non-zero byte selects
		if ((rng() % 100000) < frac * 100000) key += 1 + rng() % 9;
		keys[i] = key;
	}
	std::vector<uint32_t> wSeq, wPos, wVidx, wIdx;
	for (uint32_t i = 0; i < n; ++i) {
		const bool sel = runHeads ? (i == 0 || keys[i] != keys[i - 1]) : flag[i] != 0;
		if (sel) { wSeq.push_back(seq[i]); wPos.push_back(pos[i]); wVidx.push_back(vidx[i]); wIdx.push_back(i); }
	}
	const uint32_t ntiles = (n + kCmpTile - 1) / kCmpTile;
	std::vector<uint32_t> tileCount(ntiles, 0xDEADu);
	std::vector<uint64_t> tileOff(static_cast<size_t>(ntiles) + 2, ~0ull);
	uint64_t total = ~0ull;
	CompactIO io{};
	io.seq = seq.data(); io.pos = pos.data(); io.vidx = vidx.data(); io.flag = runHeads ? nullptr : flag.data(); io.keys = runHeads ? keys.data() : nullptr; io.n = n;
	io.tileCount = tileCount.data(); io.tileOff = tileOff.data();
	emu::launch(ntiles, 256, [&] { k_cmp_count(io); });
	emu::launch((ntiles + 1023) / 1024, 1024, [&] { k_tile_scan(tileCount.data(), ntiles, tileOff.data(), &total); });
	const size_t m = wIdx.size();
	std::vector<uint32_t> oseq(std::max<size_t>(1, m), 0xABu), opos(oseq), ovidx(oseq), oidx(oseq);
	io.oseq = oseq.data(); io.opos = opos.data(); io.ovidx = ovidx.data(); io.oidx = oidx.data();
	bool same = total == m;
	if (same) {
		emu::launch(ntiles, 256, [&] { k_cmp_scatter(io); });
		oseq.resize(m); opos.resize(m); ovidx.resize(m); oidx.resize(m);
		same = oseq == wSeq && opos == wPos && ovidx == wVidx && oidx == wIdx;
		// only the index array (what the run-length encoding of the subset choice asks for)
		std::vector<uint32_t> only(std::max<size_t>(1, m), 0xCDu);
		io.oseq = io.opos = io.ovidx = nullptr; io.oidx = only.data();
		emu::launch(ntiles, 256, [&] { k_cmp_scatter(io); });
		only.resize(m);
		same = same && only == wIdx;
	}
	std::printf("%-22s k_cmp_count + k_tile_scan + k_cmp_scatter (%s): %zu of %u elements over %u tiles  %s\n", name.c_str(), runHeads ? "run heads" : "flags", m, n, ntiles,
	            same ? "identical" : "DIFFERENT");
	return same ? 0 : 1;
}

// k_select_cls_rc (xxm_select4.cuh: forward rule + one monotone cursor into the frozen forward list) against the reference's
// updateMultOcc with Global::revcomp, literally (StartPosUtils.h:12-65: the look-back walk over the kept array and the removal of
// the loser from its middle), on doubled sequences (stored length 2 n0 + 1) with candidates on both strands
int run_rc_case(const std::string& name, int nseq, int n0Lo, int n0Hi, double hitFrac, int span, uint32_t nval, uint32_t nSiteWant, uint32_t seed) {
	std::mt19937 rng(seed);
	std::vector<int32_t> len(nseq);
	for (auto& v : len) v = 2 * (n0Lo + static_cast<int>(rng() % static_cast<uint32_t>(n0Hi - n0Lo + 1))) + 1;
	bool uniform = true;
	for (int v : len) uniform = uniform && v == len[0];
	SeqStream S{};
	S.len = len.data(); S.nent = nseq; S.uniformLen = uniform ? len[0] : 0; S.uniformWps = uniform ? 1u : 0u;
	std::vector<uint32_t> seq, pos, vidx;
	for (int q = 0; q < nseq; ++q) {
		const int n0 = (len[q] - 1) / 2;
		int boost = 0;
		for (int p = 1; p + span - 1 <= len[q]; ++p) {
			if (p + span - 1 > n0 && p <= n0 + 1) continue;  // windows over the separator never are candidates (N under a column)
			const double pr = hitFrac * (boost > 0 ? 10.0 : 1.0);
			if (boost > 0) --boost;
			if ((rng() % 100000) < pr * 100000) {
				seq.push_back(static_cast<uint32_t>(q) + 1);
				pos.push_back(static_cast<uint32_t>(p));
				uint32_t v = rng() % nval;
				if (rng() % 40 == 0) v = kNone;
				vidx.push_back(v);
				boost = 5;
			}
		}
	}
	const uint32_t nc = static_cast<uint32_t>(seq.size());
	std::vector<double> fin(nval);
	double v = 1e-9;
	for (uint32_t k = 0; k < nval; ++k) { v *= (rng() % 4 == 0) ? (1.0 + 1e-9) : 1.3; fin[k] = (k < nSiteWant) ? std::min(v, 0.999999) : 1.0; }
	for (uint32_t k = 1; k < nval; ++k) fin[k] = std::max(fin[k], fin[k - 1]);
	std::vector<uint16_t> thr(nval);
	for (uint32_t k = 0; k < nval; ++k) thr[k] = static_cast<uint16_t>(std::lower_bound(fin.begin(), fin.end(), static_cast<double>(static_cast<float>(fin[k]))) - fin.begin());
	const uint32_t nSite = static_cast<uint32_t>(std::lower_bound(fin.begin(), fin.end(), 1.0) - fin.begin());
	// ---- the reference, literally
	struct Site { int32_t sequence, startPos; float score; uint32_t idx; };
	std::vector<uint8_t> wantKeep(nc, 0);
	{
		std::vector<Site> sites(nc + 1);
		int counter = 0;
		for (uint32_t i = 0; i < nc; ++i) {
			if (vidx[i] == kNone) continue;
			const double pVal = fin[vidx[i]];
			if (!(pVal < 1.0)) continue;
			const int32_t sq = static_cast<int32_t>(seq[i]), ps = static_cast<int32_t>(pos[i]);
			const int length = span;
			bool take = true;
			if (counter > 0 && sites[counter - 1].sequence == sq && sites[counter - 1].startPos + length >= ps) {
				if (static_cast<double>(sites[counter - 1].score) <= pVal) take = false;
				else counter--;
			}
			if (take && ps > length / 2) {
				int back = 1;
				const int revPos = len[sq - 1] - length + 2 - ps;
				while (counter - back >= 0 && sites[counter - back].sequence == sq && sites[counter - back].startPos > revPos + length - 1) back++;
				if (counter - back >= 0 && sites[counter - back].sequence == sq && sites[counter - back].startPos >= revPos - length + 1) {
					if (static_cast<double>(sites[counter - back].score) <= pVal) take = false;
					else {
						while (back > 1) { sites[counter - back] = sites[counter - back + 1]; back--; }
						counter--;
					}
				}
			}
			if (!take) continue;
			sites[counter++] = Site{sq, ps, static_cast<float>(pVal), i};
		}
		for (int k = 0; k < counter; ++k) wantKeep[sites[k].idx] = 1;
	}
	std::vector<uint8_t> keep(nc + 16, 0xEE);
	std::vector<uint16_t> kcls(nc + 16, 0xEEEE);
	std::vector<uint32_t> hist(nval, 0u), counters(8, 0u), stk(nc + 16, 0u);
	SelectRcIO io{};
	io.seq = seq.data(); io.pos = pos.data(); io.vidx = vidx.data(); io.n = nc; io.nDev = nullptr; io.length = span; io.clsThr = thr.data(); io.nSite = nSite; io.nval = nval;
	io.stk = stk.data(); io.keep = keep.data(); io.kcls = kcls.data(); io.counters = counters.data(); io.hist = hist.data();
	emu::launch(std::max(1u, std::min(3u, (nc + kSel4Tile - 1) / kSel4Tile)), kSel4Threads, [&] { k_select_cls_rc(S, io); });
	int bad = 0;
	uint32_t nkept = 0;
	std::vector<uint32_t> wantHist(nval, 0u);
	for (uint32_t i = 0; i < nc; ++i) {
		if (wantKeep[i]) { ++nkept; ++wantHist[vidx[i]]; }
		const uint16_t wc = wantKeep[i] ? static_cast<uint16_t>(vidx[i]) : static_cast<uint16_t>(0xFFFF);
		if (keep[i] != wantKeep[i] || kcls[i] != wc) {
			if (!bad) std::fprintf(stderr, "FAIL %s: entry %u (seq %u pos %u cls %u): keep %u kcls %u, expected %u %u\n", name.c_str(), i, seq[i], pos[i], vidx[i], keep[i], kcls[i], wantKeep[i], wc);
			++bad;
		}
	}
	if (counters[0] != nkept) { std::fprintf(stderr, "FAIL %s: %u kept sites counted, expected %u\n", name.c_str(), counters[0], nkept); ++bad; }
	if (hist != wantHist) { std::fprintf(stderr, "FAIL %s: class histogram differs\n", name.c_str()); ++bad; }
	// ---- k_select_mops / k_select_mops_serial (xxm_kernels.cuh): the walk itself, one thread per sequence, for float-valued lists with
	// --revcomp -- the same expectation (the final table here happens to be ordered; the kernels do not use that)
	{
		Values V{};
		V.finalTab = fin.data(); V.pS = fin.data(); V.kind = kFinalTable;
		uint32_t maxPer = 0;
		{
			uint32_t run = 0;
			for (uint32_t i = 0; i < nc; ++i) {
				if (i && seq[i] != seq[i - 1]) run = 0;
				if (wantKeep[i]) maxPer = std::max(maxPer, ++run);
			}
		}
		for (int serial = 0; serial < 2; ++serial) {
			std::vector<uint8_t> keep2(std::max(1u, nc), 0);
			std::vector<float> kscore2(std::max(1u, nc), -1.f);
			std::vector<uint32_t> stk2(std::max(1u, nc), 0u), counters2(8, 0u);
			SelectIO g{};
			g.seq = seq.data(); g.pos = pos.data(); g.vidx = vidx.data(); g.n = nc; g.shift = 0; g.length = span; g.revcomp = 1;
			g.stk = stk2.data(); g.kscore = kscore2.data(); g.keep = keep2.data(); g.counters = counters2.data();
			if (serial) emu::launch(1, 32, [&] { k_select_mops_serial(S, V, g); });
			else emu::launch((nc + 255) / 256, 256, [&] { k_select_mops(S, V, g); });
			uint32_t wrong = 0;
			for (uint32_t i = 0; i < nc; ++i) {
				wrong += keep2[i] != wantKeep[i];
				if (wantKeep[i]) { const float f = static_cast<float>(fin[vidx[i]]); wrong += std::memcmp(&kscore2[i], &f, 4) != 0; }
			}
			// (the serial walk treats the whole list as one run: its per-sequence maximum is the total)
			if (wrong || counters2[0] != nkept || counters2[1] != (serial ? nkept : maxPer)) {
				std::fprintf(stderr, "FAIL %s: %s: %u entries differ, %u kept (expected %u), max per sequence %u (expected %u)\n", name.c_str(),
				             serial ? "k_select_mops_serial" : "k_select_mops", wrong, counters2[0], nkept, counters2[1], serial ? nkept : maxPer);
				++bad;
			}
		}
	}
	std::printf("%-20s %7u candidates on both strands, %6u kept, span %d: %s\n", name.c_str(), nc, nkept, span, bad ? "DIFFERENT" : "identical");
	return bad ? 1 : 0;
}

// the stable LSD radix sort of (float-score bits, index) pairs (k_rs_hist, k_scan_block_sums / k_scan_sums / k_scan_apply, k_rs_scatter;
// sortSitesByScore, StartPosUtils.h:68-70 -- glibc's qsort is a merge sort, i.e. stable) against std::stable_sort
int run_sort_case(const std::string& name, uint32_t n, uint32_t distinct, uint32_t seed) {
	std::mt19937 rng(seed);
	std::vector<uint32_t> keys(n), vals(n);
	std::vector<float> pool(distinct);
	for (auto& f : pool) f = std::ldexp(static_cast<float>((rng() % 100000) + 1) / 100000.0f, -static_cast<int>(rng() % 40));
	for (uint32_t i = 0; i < n; ++i) { const float f = pool[rng() % distinct]; std::memcpy(&keys[i], &f, 4); vals[i] = i; }
	std::vector<uint32_t> order(n);
	for (uint32_t i = 0; i < n; ++i) order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
	const uint32_t nchunks = (n + kSortChunk - 1) / kSortChunk;
	std::vector<uint32_t> ka = keys, va = vals, kb(n), vb(n), hist(256u * nchunks), sums((256u * nchunks + 1023) / 1024 + 1);
	uint32_t *pa = ka.data(), *qa = va.data(), *pb = kb.data(), *qb = vb.data();
	for (int pass = 0; pass < 4; ++pass) {
		const int shift = 8 * pass;
		emu::launch(nchunks, 256, [&] { k_rs_hist(pa, n, shift, hist.data(), nchunks); });
		const uint32_t nh = 256u * nchunks, nbk = (nh + 1023) / 1024;
		emu::launch(nbk, 256, [&] { k_scan_block_sums(hist.data(), nh, sums.data()); });
		emu::launch(1, 1024, [&] { k_scan_sums(sums.data(), nbk); });
		emu::launch(nbk, 256, [&] { k_scan_apply(hist.data(), nh, sums.data()); });
		emu::launch(nchunks, kSortWarps * 32, [&] { k_rs_scatter(pa, qa, n, shift, hist.data(), nchunks, pb, qb); });
		std::swap(pa, pb);
		std::swap(qa, qb);
	}
	int bad = 0;
	for (uint32_t i = 0; i < n && !bad; ++i)
		if (qa[i] != order[i] || pa[i] != keys[order[i]]) { std::fprintf(stderr, "FAIL %s: rank %u holds entry %u, expected %u\n", name.c_str(), i, qa[i], order[i]); ++bad; }
	std::printf("%-20s %7u keys, %5u distinct scores, %u chunks: %s\n", name.c_str(), n, distinct, nchunks, bad ? "DIFFERENT" : "identical");
	return bad;
}

}  // namespace

int main(int argc, char** argv) {
	std::vector<Case> cases = {
		{"clustered_uniform", 60, 1000, 1000, 0.02, 8, 300, 280, 0.0, 1},
		{"ragged_short", 400, 8, 120, 0.06, 8, 50, 40, 0.02, 2},
		{"long_motif", 40, 600, 900, 0.03, 26, 523, 500, 0.0, 3},
		{"all_sites_dense", 20, 400, 400, 0.3, 8, 20, 20, 0.0, 4},
		{"few_sites", 30, 500, 500, 0.05, 12, 200, 3, 0.01, 5},
		{"tiny", 1, 9, 9, 0.5, 8, 4, 4, 0.0, 6},
	};
	int failed = 0, ran = 0;
	for (const Case& c : cases) {
		if (argc > 1) {
			bool sel = false;
			for (int i = 1; i < argc; ++i) sel = sel || c.name == argv[i];
			if (!sel) continue;
		}
		failed += run_case(c) ? 1 : 0;
		++ran;
	}
	struct Z { const char* name; int nseq, lenLo, lenHi; double frac; int shift; double nFrac; uint32_t seed; };
	const Z zs[] = {{"zoops_uniform", 300, 200, 200, 0.08, 0, 0.0, 11}, {"zoops_ragged_shift", 500, 10, 300, 0.1, -2, 0.02, 12}, {"zoops_long_runs", 6, 3000, 5000, 0.5, 1, 0.0, 13},
	                {"zoops_sparse", 2000, 50, 50, 0.01, 2, 0.0, 14}};
	for (const Z& z : zs) {
		if (argc > 1) {
			bool sel = false;
			for (int i = 1; i < argc; ++i) sel = sel || std::string(z.name) == argv[i];
			if (!sel) continue;
		}
		failed += run_zoops_case(z.name, z.nseq, z.lenLo, z.lenHi, z.frac, z.shift, z.nFrac, z.seed) ? 1 : 0;
		++ran;
	}
	struct Ge { const char* name; int nseq, lenLo, lenHi; double frac; int span; uint32_t nval; int kind, shift; uint32_t seed; };
	const Ge ges[] = {{"gen_table_clustered", 120, 300, 300, 0.04, 8, 300, 0, 0, 41}, {"gen_table_ragged_long", 200, 12, 500, 0.06, 21, 4000, 0, 0, 42},
	                  {"gen_localized", 150, 100, 400, 0.05, 12, 700, 1, 0, 43}, {"gen_localized_shift", 90, 200, 350, 0.08, 9, 120, 1, -2, 44},
	                  {"gen_per_candidate", 100, 50, 600, 0.05, 17, 50, 2, 1, 45}, {"gen_dense", 10, 700, 700, 0.4, 6, 25, 0, 0, 46}, {"gen_tiny", 1, 10, 10, 0.6, 8, 5, 2, 0, 47}};
	for (const Ge& z : ges) {
		if (argc > 1) {
			bool sel = false;
			for (int i = 1; i < argc; ++i) sel = sel || std::string(z.name) == argv[i];
			if (!sel) continue;
		}
		failed += run_fwd_generic_case(z.name, z.nseq, z.lenLo, z.lenHi, z.frac, z.span, z.nval, z.kind, z.shift, z.seed) ? 1 : 0;
		++ran;
	}
	struct Cm { const char* name; uint32_t n; double frac; bool heads; uint32_t seed; };
	const Cm cms[] = {{"cmp_flags_sparse", 9000, 0.02, false, 51}, {"cmp_flags_dense", 6151, 0.7, false, 52}, {"cmp_run_heads", 12345, 0.1, true, 53}, {"cmp_one", 1, 1.0, false, 54},
	                  {"cmp_none_selected", 3000, 0.0, false, 55}, {"cmp_all_one_run", 2049, 0.0, true, 56}};
	for (const Cm& z : cms) {
		if (argc > 1) {
			bool sel = false;
			for (int i = 1; i < argc; ++i) sel = sel || std::string(z.name) == argv[i];
			if (!sel) continue;
		}
		failed += run_compact_case(z.name, z.n, z.frac, z.heads, z.seed) ? 1 : 0;
		++ran;
	}
	struct Rc { const char* name; int nseq, n0Lo, n0Hi; double frac; int span; uint32_t nval, nSite, seed; };
	const Rc rcs[] = {{"rc_uniform", 200, 250, 250, 0.03, 8, 120, 110, 21}, {"rc_ragged", 300, 20, 400, 0.05, 8, 60, 50, 22}, {"rc_long_motif", 80, 300, 600, 0.04, 21, 300, 290, 23},
	                  {"rc_long_sequences", 5, 3000, 6000, 0.05, 10, 40, 40, 24}, {"rc_palindromic_dense", 60, 100, 100, 0.4, 6, 12, 12, 25}};
	for (const Rc& z : rcs) {
		if (argc > 1) {
			bool sel = false;
			for (int i = 1; i < argc; ++i) sel = sel || std::string(z.name) == argv[i];
			if (!sel) continue;
		}
		failed += run_rc_case(z.name, z.nseq, z.n0Lo, z.n0Hi, z.frac, z.span, z.nval, z.nSite, z.seed) ? 1 : 0;
		++ran;
	}
	struct So { const char* name; uint32_t n, distinct, seed; };
	const So sos[] = {{"sort_two", 2, 2, 31}, {"sort_one_chunk", 1500, 40, 32}, {"sort_ties", 9000, 3, 33}, {"sort_many_chunks", 70000, 5000, 34}};
	for (const So& z : sos) {
		if (argc > 1) {
			bool sel = false;
			for (int i = 1; i < argc; ++i) sel = sel || std::string(z.name) == argv[i];
			if (!sel) continue;
		}
		failed += run_sort_case(z.name, z.n, z.distinct, z.seed) ? 1 : 0;
		++ran;
	}
	std::printf("%d runs, %d failed\n", ran, failed);
	return failed ? 1 : (ran ? 0 : 3);
}
