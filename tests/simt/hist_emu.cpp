// TEST INFRASTRUCTURE ONLY (see simt_emu.h): the two histogram kernels of the "next" rows -- k_kmer_hist (the
// counting loop of NullModel::init, NullModel.cpp:122-132, all three bin layouts: warp- and lane-private bins without atomics,
// lane-replicated bins with shared-memory atomics, plain bins in several passes) and k_gapped_hist (row 4: the walk of
// ElongCandidates::countAllOccurrences, elongationCandidates.cpp:234-262) -- from their CUDA SOURCES under the SIMT emulator, against
// counts taken position by position from the raw sequences.  Under build.py's sanitizers the same runs are a memcheck and a
// racecheck of these kernels (the bins of k_kmer_hist's private layout are read-modify-written WITHOUT atomics by design).
//
// Also k_encode (a1 on the device: the packing kernel of the load path for hosts with few cores, against the host packing) and
// k_logodds_hist (a14: the forward / reverse-complement log-odds histogram of Output::printScoreDistribution), against the same sums
// taken window by window.
//
// usage: hist_emu [case ...]; exit code 0 = every case identical
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "cuda_runtime.h"  // tests/simt/shim -> simt_emu.h

#include "xxm_kernels.cuh"   // k_logodds_hist
#include "xxm_kmerhist.cuh"
#include "xxm_gappedhist.cuh"
#include "xxm_region.cuh"    // k_region_hist

#include "host_pack.h"

namespace xxm {
// dynamic shared memory of the two kernels (block-scope `extern` after the rewrite of `extern __shared__`)
uint32_t s_bins[56 * 1024];
uint32_t s_gbins[1 << 14];
unsigned int s_rh[8192 + 8];
alignas(16) uint8_t smemM[1 << 16];  // (kernels of xxm_kernels.cuh that are not launched here)
uint32_t s_hs[1 << 14];
}  // namespace xxm

using namespace xxm;

namespace {

struct Case {
	std::string name;
	int nseq, lenLo, lenHi;
	double nFrac;
	bool revcomp;
	uint32_t seed;
};

struct Set {
	std::vector<int32_t> n, len;
	std::vector<uint8_t> raw;               // as loaded (pre-revcomp)
	std::vector<std::vector<uint8_t>> seq;  // stored sequences (after revcomp doubling), 0 = N, 1..4
	std::vector<uint32_t> wordOff, nmask, blk2seq;
	std::vector<uint64_t> codes;
	uint32_t nwords = 0;
};

Set make_set(const Case& c) {
	std::mt19937 rng(c.seed);
	Set s;
	s.n.resize(c.nseq);
	for (int i = 0; i < c.nseq; ++i) s.n[i] = c.lenLo + static_cast<int>(rng() % static_cast<uint32_t>(c.lenHi - c.lenLo + 1));
	uint64_t tot = 0;
	for (int v : s.n) tot += static_cast<uint64_t>(v);
	s.raw.resize(tot + 64);
	for (uint64_t i = 0; i < tot; ++i) s.raw[i] = static_cast<uint8_t>(1 + rng() % 4);
	if (c.nFrac > 0) {
		const uint64_t runs = static_cast<uint64_t>(tot * c.nFrac / 4) + 1;
		for (uint64_t r = 0; r < runs; ++r) {
			const uint64_t at = rng() % tot;
			const int run = 1 + static_cast<int>(rng() % 7);
			for (int k = 0; k < run && at + k < tot; ++k) s.raw[at + k] = 0;
		}
	}
	s.wordOff.resize(c.nseq + 1);
	s.len.resize(c.nseq);
	s.seq.resize(c.nseq);
	uint64_t w = 0, off = 0;
	for (int i = 0; i < c.nseq; ++i) {
		s.wordOff[i] = static_cast<uint32_t>(w);
		w += stored_words(static_cast<uint32_t>(s.n[i]), c.revcomp);
		s.len[i] = c.revcomp ? 2 * s.n[i] + 1 : s.n[i];
		std::vector<uint8_t>& q = s.seq[i];
		q.assign(s.raw.begin() + off, s.raw.begin() + off + s.n[i]);
		if (c.revcomp) {  // createRevcomp, seqset.cpp:128-178: S[n+1] = N, S[n+1+k] = complement(S[n+1-k])
			q.push_back(0);
			for (int k = 1; k <= s.n[i]; ++k) { const uint8_t b = q[s.n[i] - k]; q.push_back(b ? static_cast<uint8_t>(5 - b) : 0); }
		}
		off += s.n[i];
	}
	s.wordOff[c.nseq] = static_cast<uint32_t>(w);
	s.nwords = static_cast<uint32_t>(w);
	s.codes.assign(w + 2, 0ull);
	s.nmask.assign(w + 2, 0xFFFFFFFFu);
	std::vector<int32_t> notX(c.nseq);
	if (pack_range(s.n.data(), 0, static_cast<uint32_t>(c.nseq), s.raw.data(), tot, c.revcomp, s.codes.data(), s.nmask.data(), notX.data())) std::exit(2);
	const uint32_t nblk = s.nwords / kBlkWords + 1;
	s.blk2seq.assign(nblk, 0);
	for (int q = 0; q < c.nseq; ++q) {
		const uint32_t w0 = s.wordOff[q], w1 = s.wordOff[q + 1];
		for (uint32_t b = (w0 + kBlkWords - 1) / kBlkWords; b < nblk && (b * kBlkWords < w1 || q + 1 == c.nseq); ++b) s.blk2seq[b] = static_cast<uint32_t>(q);
	}
	return s;
}

// ---- k_kmer_hist: every position adds one to the histogram of its LONGEST valid k-mer (length m <= K)
int run_kmer(const Case& c, const Set& s, int K, unsigned grid) {
	const uint32_t nfull = 1u << (2 * K);
	const uint32_t nshort = (nfull - 4u) / 3u;
	std::vector<unsigned long long> want(nfull + std::max(1u, nshort), 0ull), got(want.size(), 0ull);
	for (const auto& q : s.seq)
		for (size_t l = 0; l < q.size(); ++l) {
			if (!q[l]) continue;
			int m = 0;
			uint32_t v = 0;
			while (m < K && l + m < q.size() && q[l + m]) { v |= static_cast<uint32_t>(q[l + m] - 1) << (2 * m); ++m; }
			if (m == K) want[v]++;
			else want[nfull + ((1u << (2 * m)) - 4u) / 3u + v]++;
		}
	// the bin layout the host picks (xxm_kmer_counts, xxm_gpu_part6.inc)
	const bool laneRep = nfull <= 1024, warpPriv = nfull <= 64;
	const uint32_t perPass = laneRep ? nfull : 49152u;
	KmerHistIO io{};
	io.codes = s.codes.data(); io.nmask = s.nmask.data(); io.nwords = s.nwords; io.K = K;
	io.full = got.data(); io.shortH = got.data() + nfull;
	uint32_t passes = 0;
	for (uint32_t lo = 0; lo < nfull; lo += perPass, ++passes) {
		io.binLo = lo; io.binCount = std::min(perPass, nfull - lo); io.laneRep = warpPriv ? 2 : laneRep ? 1 : 0; io.countShort = passes == 0;
		emu::launch(grid, kKmerHistThreads, [&] { k_kmer_hist(io); });
	}
	const bool same = want == got;
	unsigned long long total = 0;
	for (uint32_t i = 0; i < nfull; ++i) total += got[i];
	std::printf("%-22s k_kmer_hist K=%d layout %d grid %u: %llu full-length positions, %u pass(es)  %s\n", c.name.c_str(), K, warpPriv ? 2 : laneRep ? 1 : 0, grid, total,
	            passes, same ? "identical" : "DIFFERENT");
	return same ? 0 : 1;
}

// ---- k_gapped_hist: every start position whose coding characters hold no N adds one to the counter of its gapped k-mer
int run_gapped(const Case& c, const Set& s, const std::vector<int>& offset, bool firstHalfOnly, int nseqScan, unsigned grid) {
	const int k = static_cast<int>(offset.size());
	const int span = offset.back() + 1;
	const uint32_t nb = 1u << (2 * k);
	std::vector<uint32_t> want(nb, 0u), got(nb, 0u);
	for (int i = 0; i < nseqScan; ++i) {
		const auto& q = s.seq[i];
		int last = static_cast<int>(q.size()) - span + 1;   // start positions 1..last (elongationCandidates.cpp:243-247)
		if (firstHalfOnly) last /= 2;
		for (int j = 0; j < last; ++j) {
			uint32_t v = 0;
			bool ok = true;
			for (int l = 0; l < k && ok; ++l) {
				const uint8_t b = q[j + offset[l]];
				if (!b) ok = false;
				else v |= static_cast<uint32_t>(b - 1) << (2 * l);
			}
			if (ok) want[v]++;
		}
	}
	SeqStream S{};
	S.codes = s.codes.data(); S.nmask = s.nmask.data(); S.wordOff = s.wordOff.data(); S.len = s.len.data(); S.blk2seq = s.blk2seq.data();
	S.nwords = s.nwords; S.nent = static_cast<int32_t>(s.n.size());
	GappedHistIO io{};  // the descriptors of xxm_count_gapped_kmers (xxm_gpu_part9.inc)
	io.scanWords = s.wordOff[nseqScan]; io.k = k; io.span = span; io.firstHalfOnly = firstHalfOnly ? 1 : 0;
	for (int l = 0; l < k;) {
		int e = l + 1;
		while (e < k && offset[e] == offset[e - 1] + 1) ++e;
		const int g = io.ngroups++;
		io.goff[g] = offset[l]; io.gshift[g] = 2 * l; io.gmask[g] = (1u << (2 * (e - l))) - 1u;
		l = e;
	}
	for (int l = 0; l < k; ++l) io.codingMask |= 1u << offset[l];
	io.bins = got.data();
	emu::launch(grid, kGappedHistThreads, [&] { k_gapped_hist(S, io); });
	const bool same = want == got;
	unsigned long long total = 0;
	for (uint32_t v : got) total += v;
	std::printf("%-22s k_gapped_hist k=%d span %d%s, %d of %zu sequences, grid %u: %llu k-mers  %s\n", c.name.c_str(), k, span, firstHalfOnly ? " first half" : "", nseqScan,
	            s.n.size(), grid, total, same ? "identical" : "DIFFERENT");
	return same ? 0 : 1;
}

// ---- k_logodds_hist (a14: Output::printScoreDistribution, output.cpp:1050-1064 with calculateMatrixScore_logOdds, Motif.h:584-599):
// per start position, the number of sequences whose window scores above the threshold under the PWM and under its reverse complement
int run_logodds(const Case& c, const Set& s, const std::vector<int>& coloff, int nseqScan, uint32_t seed) {
	std::mt19937 rng(seed);
	std::uniform_real_distribution<double> U(-2.5, 1.5);
	HistModel H{};
	H.ncols = static_cast<int>(coloff.size());
	H.span = coloff.back() + 1;
	for (int k = 0; k < H.ncols; ++k) {
		H.coloff[k] = coloff[static_cast<size_t>(k)];
		H.lodF[k][0] = H.lodB[k][0] = std::log(1.0 / 1000) + std::log(2.0);  // the N penalty (Motif.h:592)
		for (int a = 1; a <= 4; ++a) { H.lodF[k][a] = U(rng); H.lodB[k][a] = U(rng); }
	}
	H.thrF = 0.3 * H.ncols * 0.5; H.thrB = -0.2;
	int maxLen = 0;
	for (int i = 0; i < nseqScan; ++i) maxLen = std::max(maxLen, s.len[static_cast<size_t>(i)]);
	const int npos = std::max(1, maxLen - H.span + 1 - 3);  // a few positions short of the longest sequence: the kernel must stop at npos
	std::vector<unsigned int> wantF(static_cast<size_t>(npos), 0u), wantB(wantF), gotF(wantF), gotB(wantF);
	for (int i = 0; i < nseqScan; ++i) {
		const auto& q = s.seq[static_cast<size_t>(i)];
		const int n = static_cast<int>(q.size());
		for (int pos = 1; pos <= n - H.span + 1 && pos <= npos; ++pos) {
			double f = 0, b = 0;
			for (int k = 0; k < H.ncols; ++k) {
				const int a = q[static_cast<size_t>(pos - 1 + H.coloff[k])];
				f = f + H.lodF[k][a];
				b = b + H.lodB[k][a];
			}
			if (f > H.thrF) wantF[static_cast<size_t>(pos - 1)]++;
			if (b > H.thrB) wantB[static_cast<size_t>(pos - 1)]++;
		}
	}
	SeqStream S{};
	S.codes = s.codes.data(); S.nmask = s.nmask.data(); S.wordOff = s.wordOff.data(); S.len = s.len.data(); S.blk2seq = s.blk2seq.data();
	S.nwords = s.nwords; S.nent = static_cast<int32_t>(s.n.size());
	const uint32_t scanWords = s.wordOff[static_cast<size_t>(nseqScan)];
	emu::launch((scanWords + 255) / 256, 256, [&] { k_logodds_hist(S, H, scanWords, gotF.data(), gotB.data(), npos); });
	const bool same = wantF == gotF && wantB == gotB;
	unsigned long long tf = 0, tb = 0;
	for (unsigned v : gotF) tf += v;
	for (unsigned v : gotB) tb += v;
	std::printf("%-22s k_logodds_hist %d columns span %d, %d sequences, %d positions: %llu / %llu windows above the thresholds  %s\n", c.name.c_str(), H.ncols, H.span,
	            nseqScan, npos, tf, tb, same ? "identical" : "DIFFERENT");
	return same ? 0 : 1;
}

// ---- k_encode (a1 on the device: raw codes -> packed stream with the reverse strand generated on the fly, createRevcomp,
// seqset.cpp:128-178; the load path of hosts with few cores per GPU) against the host packing (host_pack.cpp), word for word
int run_encode(const Case& c, const Set& s) {
	uint64_t tot = 0;
	std::vector<uint64_t> rawOff(s.n.size() + 1, 0);
	for (size_t i = 0; i < s.n.size(); ++i) { rawOff[i] = tot; tot += static_cast<uint64_t>(s.n[i]); }
	rawOff[s.n.size()] = tot;
	const std::vector<uint8_t> raw(s.raw.begin(), s.raw.begin() + static_cast<std::ptrdiff_t>(tot));  // exactly the residues: no slack to over-read into
	std::vector<uint64_t> codes(static_cast<size_t>(s.nwords) + 2, 0x5555555555555555ull);
	std::vector<uint32_t> nmask(static_cast<size_t>(s.nwords) + 2, 0x12345678u);
	SeqStream S{};
	S.wordOff = s.wordOff.data(); S.len = s.len.data(); S.blk2seq = s.blk2seq.data(); S.nwords = s.nwords; S.nent = static_cast<int32_t>(s.n.size());
	// three chunks of whole words, the last one with the two guard words (load_pack_on_device, xxm_gpu.cu)
	const uint32_t cut1 = s.nwords / 3, cut2 = 2 * (s.nwords / 3) + 1;
	const uint32_t bounds[4] = {0, cut1, std::min(cut2, s.nwords), s.nwords + 2};
	for (int k = 0; k < 3; ++k) {
		const uint32_t w0 = bounds[k], w1 = bounds[k + 1];
		if (w1 <= w0) continue;
		emu::launch((w1 - w0 + 255) / 256, 256, [&] { k_encode(raw.data(), rawOff.data(), S, w0, w1, codes.data(), nmask.data(), c.revcomp ? 1 : 0); });
	}
	const bool same = codes == s.codes && nmask == s.nmask;
	std::printf("%-22s k_encode%s: %u words + 2 guard words in 3 chunks vs the host packing  %s\n", c.name.c_str(), c.revcomp ? " --revcomp" : "", s.nwords,
	            same ? "identical" : "DIFFERENT");
	return same ? 0 : 1;
}

// ---- k_region_hist (a15: the start-position histogram of Motif::updateEnrichedRegion, Motif.cpp:989-999: candidates with
// float(pValScore) < 1e-3 inside the search window, reverse-strand positions folded back)
int run_region(const std::string& name, uint32_t n, int max_leng, bool revcomp, int motifLength, int lo, int hi, uint32_t seed) {
	std::mt19937 rng(seed);
	std::vector<uint32_t> pos(n);
	std::vector<float> score(n);
	for (uint32_t i = 0; i < n; ++i) {
		pos[i] = 1 + rng() % static_cast<uint32_t>(max_leng);
		const int k = static_cast<int>(rng() % 4);
		score[i] = k == 0 ? 1e-3f : k == 1 ? 9.99e-4f : k == 2 ? 1.0f : 1e-7f;  // 1e-3f as a double is just above 1e-3: not counted
	}
	const int nc = max_leng + 2;
	std::vector<unsigned int> want(static_cast<size_t>(nc) + 1, 0u);
	for (uint32_t i = 0; i < n; ++i) {
		if (!(static_cast<double>(score[i]) < 1e-3)) continue;
		int p = static_cast<int>(pos[i]);
		if (revcomp && p > max_leng / 2) p = max_leng - (p + motifLength - 1);
		if (p >= lo && p <= hi && p >= 0 && p <= max_leng) { ++want[static_cast<size_t>(p)]; ++want[static_cast<size_t>(nc)]; }
	}
	int bad = 0;
	for (uint32_t nbs : {static_cast<uint32_t>(nc), 0u, static_cast<uint32_t>(nc) / 2}) {  // bins in shared memory, in global memory, split
		std::vector<unsigned int> got(static_cast<size_t>(nc) + 1, 0u);
		emu::launch(3, 256, [&] { k_region_hist(pos.data(), score.data(), n, revcomp ? 1 : 0, max_leng, motifLength, lo, hi, got.data(), got.data() + nc, nbs); });
		bad += got == want ? 0 : 1;
	}
	std::printf("%-22s k_region_hist %u candidates, max_leng %d%s, window [%d, %d]: %u counted  %s\n", name.c_str(), n, max_leng, revcomp ? " --revcomp" : "", lo, hi,
	            want[static_cast<size_t>(nc)], bad ? "DIFFERENT" : "identical");
	return bad ? 1 : 0;
}

}  // namespace

int main(int argc, char** argv) {
	const std::vector<Case> cases = {
		{"uniform", 24, 300, 300, 0.0, false, 11},
		{"ragged_N", 40, 5, 260, 0.06, false, 12},
		{"revcomp_ragged_N", 30, 9, 180, 0.04, true, 13},
		{"short_sequences", 60, 1, 40, 0.10, false, 14},
		{"revcomp_short", 50, 1, 70, 0.08, true, 15},
	};
	int runs = 0, failed = 0;
	for (const Case& c : cases) {
		bool wanted = argc <= 1;
		for (int i = 1; i < argc; ++i) wanted = wanted || c.name == argv[i];
		if (!wanted) continue;
		const Set s = make_set(c);
		for (int K : {1, 3, 4, 5, 7}) { failed += run_kmer(c, s, K, K == 7 ? 1u : 2u); ++runs; }   // layouts 2, 2, 1, 1, 0
		if (c.name == "ragged_N") { failed += run_kmer(c, s, 8, 1u); ++runs; }                     // 65536 bins: two passes of plain bins
		failed += run_gapped(c, s, {0, 1, 2, 3, 4}, false, c.nseq, 2); ++runs;
		failed += run_gapped(c, s, {0, 1, 4, 5, 9}, false, c.nseq, 3); ++runs;
		failed += run_gapped(c, s, {0, 2, 3, 11, 12, 20}, false, std::max(1, c.nseq / 2), 2); ++runs;
		if (c.revcomp) { failed += run_gapped(c, s, {0, 1, 2, 8, 9, 10}, true, c.nseq, 2); ++runs; }  // palindromic seeds: first half of the doubled sequence
		failed += run_encode(c, s); ++runs;
		failed += run_logodds(c, s, {0, 1, 2, 3, 4, 5, 6, 7}, c.nseq, c.seed + 100); ++runs;
		failed += run_logodds(c, s, {0, 1, 3, 4, 5, 9, 10, 12, 13, 14, 15, 20}, std::max(1, c.nseq - 3), c.seed + 200); ++runs;
	}
	struct Rg { const char* name; uint32_t n; int max_leng; bool rc; int motifLength, lo, hi; uint32_t seed; };
	const Rg rgs[] = {{"region_forward", 5000, 300, false, 8, 120, 220, 81}, {"region_revcomp", 7000, 601, true, 10, -20, 150, 82}, {"region_whole", 900, 50, false, 6, -100, 1000, 83}};
	for (const Rg& g : rgs) {
		bool wanted = argc <= 1;
		for (int i = 1; i < argc; ++i) wanted = wanted || std::string(g.name) == argv[i];
		if (!wanted) continue;
		failed += run_region(g.name, g.n, g.max_leng, g.rc, g.motifLength, g.lo, g.hi, g.seed);
		++runs;
	}
	std::printf("%d runs, %d failed\n", runs, failed);
	return failed ? 1 : 0;
}
