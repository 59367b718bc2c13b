// TEST INFRASTRUCTURE ONLY (see simt_emu.h): runs the full-scan kernels of libxxmotif_gpu.so -- k_dead_mask, k_scan5 (find),
// k_tile_scan, k_expand6 / k_expand7 -- from their CUDA SOURCES under the SIMT emulator on seeded sequence sets, and compares the
// site records with a brute-force enumeration of fill_sites_with_possible_startPositions' window loop
// (refinementPhase/StartPosUpdater-inl.h:70-140; table membership branch_and_bound-inl.h:386-392): bit-exact
// (sequence, position, class) in scan order, plus the record count.
//
// usage: scan_emu [case ...]   (no argument: all cases); exit code 0 = every case identical
#include <algorithm>
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
#include "xxm_scan6.cuh"
#include "xxm_scan7.cuh"

#include "host_pack.h"
#include "pair_table.h"

namespace xxm {
// the kernels' dynamic shared memory (block-scope `extern` declarations after the emulator's rewrite of `extern __shared__`)
alignas(16) uint8_t smem5[kPairTabWords * 4];
alignas(16) uint8_t smem4[65536 + 65536];
alignas(16) uint8_t smemM[1 << 16];
uint32_t s_hs[1 << 14];
}  // namespace xxm

using namespace xxm;

namespace {

struct Case {
	std::string name;
	int nseq;
	int lenLo, lenHi;     // lengths drawn from [lenLo, lenHi]
	double nFrac;         // fraction of N residues (in runs)
	bool revcomp;
	std::vector<int> coloff;  // match columns (offsets from the window start, coloff[0] = 0, last <= 7)
	double keepFrac;      // fraction of all k-mers in the table (hit density)
	bool lowComplexity;   // TATATA... stretches: dense hits, exercises the dense-warp path of the expand kernel
	uint32_t seed;
	int scanSeqs = 0;     // > 0: only the first scanSeqs sequences are scanned (posSetSize < nent: a negative set follows)
	std::vector<int> rightOff;  // non-empty: a two-part model -- offsets of the match columns beyond the table's (more than 8 match columns)
};

// the second part of a two-part model (branch_and_bound-inl.h:399-423): right table, class p-values, product threshold, log-odds
struct Right {
	std::vector<uint32_t> rank1;
	std::vector<double> cpval0, cpval1;
	double pruneThr = 0;
	double logodds[32][4];
};
Right make_right(const Case& c, uint32_t nclass0) {
	std::mt19937 rng(c.seed * 104729u + 5u);
	std::uniform_real_distribution<double> U(1e-6, 1e-2), Lo(-1.5, 1.0);
	Right r;
	r.rank1.assign(static_cast<size_t>(1) << (2 * c.rightOff.size()), 0u);
	const uint32_t nclass1 = 40;
	for (auto& v : r.rank1)
		if (rng() % 3) v = static_cast<uint32_t>(rng() % nclass1) + 1;
	r.cpval0.resize(std::max(1u, nclass0)); r.cpval1.resize(nclass1);
	for (double& v : r.cpval0) v = U(rng);
	for (double& v : r.cpval1) v = U(rng);
	r.pruneThr = 2.5e-5;
	for (auto& row : r.logodds)
		for (double& v : row) v = Lo(rng);
	return r;
}

struct Rec { uint32_t seq, pos, v; bool none = false; int32_t s = 0; bool twoPart = false; };  // two-part models: v is a memo slot, checked through its key s

struct Set {
	std::vector<int32_t> n;      // raw lengths
	std::vector<uint8_t> raw;    // 0 = N, 1..4
	std::vector<int32_t> len;    // stored lengths
	std::vector<uint32_t> wordOff;
	std::vector<uint64_t> codes;
	std::vector<uint32_t> nmask;
	std::vector<uint32_t> blk2seq;
	uint32_t nwords = 0;
	uint32_t uniformWps = 0;
	int32_t uniformLen = 0;
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
	if (c.lowComplexity) {
		for (uint64_t i = 0; i + 64 < tot; i += 97 + rng() % 200) {
			const int run = 20 + static_cast<int>(rng() % 120);
			for (int k = 0; k < run && i + k < tot; ++k) s.raw[i + k] = (k & 1) ? 1 : 4;  // TATATA
		}
	}
	if (c.nFrac > 0) {
		const uint64_t runs = static_cast<uint64_t>(tot * c.nFrac / 6) + 1;
		for (uint64_t r = 0; r < runs; ++r) {
			const uint64_t at = rng() % tot;
			const int run = 1 + static_cast<int>(rng() % 11);
			for (int k = 0; k < run && at + k < tot; ++k) s.raw[at + k] = 0;
		}
	}
	s.wordOff.resize(c.nseq + 1);
	s.len.resize(c.nseq);
	uint64_t w = 0;
	bool uniform = true;
	for (int i = 0; i < c.nseq; ++i) {
		s.wordOff[i] = static_cast<uint32_t>(w);
		w += stored_words(static_cast<uint32_t>(s.n[i]), c.revcomp);
		s.len[i] = c.revcomp ? 2 * s.n[i] + 1 : s.n[i];
		if (s.n[i] != s.n[0]) uniform = false;
	}
	s.wordOff[c.nseq] = static_cast<uint32_t>(w);
	s.nwords = static_cast<uint32_t>(w);
	s.codes.assign(w + 2, 0ull);
	s.nmask.assign(w + 2, 0xFFFFFFFFu);
	std::vector<int32_t> notX(c.nseq);
	const uint32_t bad = pack_range(s.n.data(), 0, static_cast<uint32_t>(c.nseq), s.raw.data(), tot, c.revcomp, s.codes.data(), s.nmask.data(), notX.data());
	if (bad) { std::fprintf(stderr, "pack_range rejected sequence %u\n", bad); std::exit(2); }
	if (uniform) { s.uniformWps = s.wordOff[1] - s.wordOff[0]; s.uniformLen = s.len[0]; }
	const uint32_t nblk = s.nwords / kBlkWords + 1;
	s.blk2seq.assign(nblk, 0);
	for (int q = 0; q < c.nseq; ++q) {
		const uint32_t w0 = s.wordOff[q], w1 = s.wordOff[q + 1];
		for (uint32_t b = (w0 + kBlkWords - 1) / kBlkWords; b < nblk && (b * kBlkWords < w1 || q + 1 == c.nseq); ++b) s.blk2seq[b] = static_cast<uint32_t>(q);
	}
	return s;
}

// a table "enumerated by a score threshold" (what set_kmer_to_Pvalue_hash builds): additive integer column scores, the best
// keepFrac of all k-mers; classes = distinct scores in descending order.  rank[id] = class + 1, 0 = not in the table.
struct Table {
	int len;
	std::vector<uint32_t> ids;
	std::vector<uint32_t> rank;
	uint32_t nclass = 0;
};

Table make_table(const Case& c) {
	std::mt19937 rng(c.seed * 7919u + 13u);
	Table t;
	t.len = static_cast<int>(c.coloff.size());
	int sc[8][4];
	for (int k = 0; k < t.len; ++k)
		for (int a = 0; a < 4; ++a) sc[k][a] = static_cast<int>(rng() % 1000);
	const uint32_t N = 1u << (2 * t.len);
	std::vector<std::pair<int, uint32_t>> all(N);
	for (uint32_t id = 0; id < N; ++id) {
		int v = 0;
		for (int k = 0; k < t.len; ++k) v += sc[k][(id >> (2 * k)) & 3u];
		all[id] = {-v, id};
	}
	std::sort(all.begin(), all.end());
	size_t keep = std::max<size_t>(1, static_cast<size_t>(N * c.keepFrac));
	while (keep < N && all[keep].first == all[keep - 1].first) ++keep;  // a threshold never splits a score class
	t.rank.assign(N, 0u);
	uint32_t cls = 0;
	for (size_t i = 0; i < keep; ++i) {
		if (i && all[i].first != all[i - 1].first) ++cls;
		t.ids.push_back(all[i].second);
		t.rank[all[i].second] = cls + 1;
	}
	t.nclass = cls + 1;
	return t;
}

std::vector<Rec> brute_force(const Case& c, const Set& s, const Table& t, const Right* R = nullptr) {
	std::vector<Rec> out;
	const int span = (c.rightOff.empty() ? c.coloff.back() : c.rightOff.back()) + 1;
	const int nscan = c.scanSeqs > 0 ? c.scanSeqs : c.nseq;
	for (int q = 0; q < nscan; ++q) {
		const uint32_t w0 = s.wordOff[q];
		const int n = s.len[q];
		auto base = [&](int p, int& isN) {  // stored position p (0-based) of sequence q, from the packed stream
			const uint32_t w = w0 + static_cast<uint32_t>(p >> 5);
			isN = (s.nmask[w] >> (p & 31)) & 1u;
			return static_cast<uint32_t>(s.codes[w] >> (2 * (p & 31))) & 3u;
		};
		for (int p = 0; p + span <= n; ++p) {
			uint32_t id = 0;
			bool dead = false;
			for (size_t k = 0; k < c.coloff.size(); ++k) {
				int isN = 0;
				const uint32_t b = base(p + c.coloff[k], isN);
				if (isN) { dead = true; break; }
				id |= b << (2 * k);
			}
			uint32_t id1 = 0;
			for (size_t k = 0; k < c.rightOff.size() && !dead; ++k) {  // an N in ANY match column kills the window (:388)
				int isN = 0;
				const uint32_t b = base(p + c.rightOff[k], isN);
				if (isN) dead = true;
				id1 |= b << (2 * k);
			}
			if (dead || !t.rank[id]) continue;
			Rec rec{static_cast<uint32_t>(q) + 1, static_cast<uint32_t>(p) + 1, t.rank[id] - 1};
			if (R) {
				rec.twoPart = true;
				const uint32_t r0 = t.rank[id], r1 = R->rank1[id1];
				if (r1 == 0 || R->cpval0[r0 - 1] * R->cpval1[r1 - 1] > R->pruneThr) rec.none = true;
				else {
					double score = 0;
					int isN = 0;
					for (size_t k = 0; k < c.coloff.size(); ++k) score = score + R->logodds[k][base(p + c.coloff[k], isN)];
					for (size_t k = 0; k < c.rightOff.size(); ++k) score = score + R->logodds[c.coloff.size() + k][base(p + c.rightOff[k], isN)];
					rec.s = static_cast<int32_t>(score * 1e6 - 1.0);
				}
			}
			out.push_back(rec);
		}
	}
	return out;
}

struct Result {
	std::vector<int32_t> memoKey;
	std::vector<Rec> recs;
	uint64_t total = 0;
	uint64_t collectives = 0;
};

// generation 6 = k_scan5 (even/odd masks) + k_expand6; generation 7 = k_scan5<natural> + k_expand7
Result run_scan(const Case& c, const Set& s, const Table& t, int generation, unsigned findGrid, size_t cap, const Right* RT = nullptr) {
	SeqStream S{};
	S.codes = s.codes.data(); S.nmask = s.nmask.data(); S.wordOff = s.wordOff.data(); S.len = s.len.data(); S.blk2seq = s.blk2seq.data();
	S.nwords = s.nwords; S.nent = c.nseq; S.uniformWps = s.uniformWps; S.uniformLen = s.uniformLen;
	Model M{};
	M.L = static_cast<int>(c.coloff.size()); M.maxk = 8; M.nparts = 1; M.span = c.coloff.back() + 1; M.contiguous = (M.span == M.L);
	M.len0 = M.L; M.len1 = 0;
	for (int k = 0; k < M.L; ++k) M.coloff[k] = c.coloff[k];
	M.rank0 = t.rank.data();
	const int leftCols = M.L;
	Result R;
	uint32_t memoOverflow = 0;
	if (RT) {  // the table covers the left part; the window has L = left + right match columns
		M.nparts = 2; M.len1 = static_cast<int>(c.rightOff.size()); M.L = leftCols + M.len1;
		for (int k = 0; k < M.len1; ++k) M.coloff[leftCols + k] = c.rightOff[static_cast<size_t>(k)];
		M.span = c.rightOff.back() + 1; M.contiguous = (M.span == M.L);
		M.rank1 = RT->rank1.data(); M.cpval0 = RT->cpval0.data(); M.cpval1 = RT->cpval1.data(); M.pruneThr = RT->pruneThr;
		M.nclass0 = static_cast<int>(RT->cpval0.size()); M.nclass1 = static_cast<int>(RT->cpval1.size());
		for (int k = 0; k < M.L; ++k)
			for (int a = 0; a < 4; ++a) M.logodds[k][a] = RT->logodds[k][a];
		R.memoKey.assign(1u << 14, kMemoEmpty);
		M.memoKey = R.memoKey.data(); M.memoCap = 1u << 14; M.memoOverflow = &memoOverflow;
	}
	std::vector<uint32_t> pairTab(kPairEntries);
	int levels = 0;
	std::vector<int> co(c.coloff);
	if (!build_pair_table(t.ids.data(), t.ids.size(), co.data(), leftCols, pairTab.data(), &levels)) {
		std::fprintf(stderr, "%s: the pair table cannot represent this table\n", c.name.c_str());
		std::exit(2);
	}
	const uint32_t scanWords = s.wordOff[c.scanSeqs > 0 ? c.scanSeqs : c.nseq];
	const uint32_t ntiles = std::max<uint32_t>(1, (scanWords + kTileWords - 1) / kTileWords);
	std::vector<uint32_t> dead(static_cast<size_t>(s.nwords) + 8, 0xDEADBEEFu), hitmask(static_cast<size_t>(scanWords) + 8, 0xDEADBEEFu);
	std::vector<uint32_t> counts(static_cast<size_t>(ntiles) * (kChunksPerTile + 1), 0u);
	std::vector<uint64_t> tileOff(static_cast<size_t>(ntiles) + 2, 0ull);
	uint64_t total = ~0ull;
	DeadCols C{};
	C.L = M.L; C.contiguous = M.contiguous;
	for (int k = 0; k < M.L; ++k) C.coloff[k] = M.coloff[k];
	const bool natural = generation >= 7;
	emu::launch((s.nwords + 8 + 255) / 256, 256, [&] {
		if (natural) k_dead_mask<true>(s.nmask.data(), s.nwords, C, dead.data());
		else k_dead_mask<false>(s.nmask.data(), s.nwords, C, dead.data());
	});
	R.collectives += emu::lastLaunch.collectives;
	uint32_t* chunkCount = counts.data();
	uint32_t* tileCount = counts.data() + static_cast<size_t>(ntiles) * kChunksPerTile;
	emu::launch(findGrid, kScan5Threads, [&] {
		if (natural) k_scan5<true>(s.codes.data(), pairTab.data(), dead.data(), hitmask.data(), chunkCount, tileCount, scanWords);
		else k_scan5<false>(s.codes.data(), pairTab.data(), dead.data(), hitmask.data(), chunkCount, tileCount, scanWords);
	});
	R.collectives += emu::lastLaunch.collectives;
	emu::launch((ntiles + 1023) / 1024, 1024, [&] { k_tile_scan(tileCount, ntiles, tileOff.data(), &total); });
	R.collectives += emu::lastLaunch.collectives;
	std::vector<uint32_t> oseq(cap + 1, 0xABABABABu), opos(cap + 1, 0xABABABABu), ovidx(cap + 1, 0xABABABABu);
	Expand6IO io{};
	io.hitmask = hitmask.data(); io.chunkCount = chunkCount; io.tileOff = tileOff.data(); io.scanWords = scanWords;
	io.oseq = oseq.data(); io.opos = opos.data(); io.ovidx = ovidx.data();
	io.cap = static_cast<uint32_t>(cap);
	io.wpsMagic = S.uniformWps ? static_cast<uint32_t>((1ull << 32) / S.uniformWps) : 0u;
	if (S.uniformWps == 1) io.wpsMagic = 0xFFFFFFFFu;
	const bool fast = M.nparts == 1 && M.contiguous;
	if (generation == 6) {
		emu::launch(ntiles, kExpand6Threads, [&] {
			if (fast) k_expand6<true>(S, M, io);
			else k_expand6<false>(S, M, io);
		});
	} else {
		emu::launch(ntiles, kExpand7Threads, [&] {
			if (fast && S.uniformWps) k_expand7<true, true>(S, M, io);
			else if (fast) k_expand7<true, false>(S, M, io);
			else if (S.uniformWps) k_expand7<false, true>(S, M, io);
			else k_expand7<false, false>(S, M, io);
		});
	}
	R.collectives += emu::lastLaunch.collectives;
	R.total = total;
	if (memoOverflow) { std::fprintf(stderr, "%s: the memo overflowed\n", c.name.c_str()); std::exit(1); }
	// nothing may be written behind the capacity
	if (oseq[cap] != 0xABABABABu || opos[cap] != 0xABABABABu || ovidx[cap] != 0xABABABABu) {
		std::fprintf(stderr, "%s: generation %d wrote behind the capacity of the record arrays\n", c.name.c_str(), generation);
		std::exit(1);
	}
	const size_t nrec = std::min<uint64_t>(total, cap);
	R.recs.resize(nrec);
	for (size_t i = 0; i < nrec; ++i) R.recs[i] = Rec{oseq[i], opos[i], ovidx[i]};
	return R;
}

bool check(const Case& c, int generation, const std::vector<Rec>& want, const Result& got, size_t cap) {
	if (got.total != want.size()) {
		std::fprintf(stderr, "FAIL %s gen %d: %llu records, expected %zu\n", c.name.c_str(), generation, static_cast<unsigned long long>(got.total), want.size());
		return false;
	}
	const size_t n = std::min(want.size(), cap);
	for (size_t i = 0; i < n; ++i) {
		bool vOk = got.recs[i].v == want[i].v;
		if (want[i].twoPart)
			vOk = want[i].none ? got.recs[i].v == kNone : (got.recs[i].v < got.memoKey.size() && got.memoKey[got.recs[i].v] == want[i].s);
		if (got.recs[i].seq != want[i].seq || got.recs[i].pos != want[i].pos || !vOk) {
			std::fprintf(stderr, "FAIL %s gen %d: record %zu is (%u, %u, %u), expected (%u, %u, %u)\n", c.name.c_str(), generation, i, got.recs[i].seq,
			             got.recs[i].pos, got.recs[i].v, want[i].seq, want[i].pos, want[i].v);
			return false;
		}
	}
	return true;
}

}  // namespace

int main(int argc, char** argv) {
	const std::vector<int> c8 = {0, 1, 2, 3, 4, 5, 6, 7};
	std::vector<Case> cases = {
		{"uniform_1kb", 40, 1000, 1000, 0.0, false, c8, 0.02, false, 1},
		{"ragged_N", 150, 20, 400, 0.02, false, c8, 0.03, false, 2},
		{"revcomp_ragged", 60, 30, 300, 0.01, true, c8, 0.03, false, 3},
		{"dense_lowcomplexity", 30, 800, 800, 0.0, false, c8, 0.25, true, 4},
		{"gapped_6cols", 80, 100, 500, 0.01, false, {0, 1, 2, 4, 5, 7}, 0.05, false, 5},
		{"five_columns", 50, 64, 64, 0.0, false, {0, 1, 2, 3, 4}, 0.10, false, 6},
		{"tiny", 1, 9, 9, 0.0, false, c8, 0.05, false, 7},
		{"every_warp_dense", 12, 700, 900, 0.001, false, {0, 1, 2, 3}, 0.5, false, 10},
		{"ten_tiles_uniform", 320, 1000, 1000, 0.0, false, c8, 0.02, false, 11},
		{"one_word_sequences", 300, 1, 31, 0.05, false, {0, 1, 2, 3}, 0.3, false, 8},
		{"small_capacity", 40, 500, 500, 0.0, false, c8, 0.05, false, 9},
		{"positive_set_prefix_ragged", 200, 20, 400, 0.01, false, c8, 0.04, false, 12, 77},
		{"positive_set_prefix_uniform", 90, 300, 300, 0.0, false, c8, 0.04, false, 13, 31},
		// two-part models (more than 8 match columns): the pair table finds the left 8, k_expand6/7<false> evaluate both tables, the product
		// test and the memo slot of the window's integer score
		{"two_part_twelve", 70, 100, 500, 0.01, false, c8, 0.03, false, 14, 0, {8, 9, 10, 11}},
		{"two_part_seventeen_rc", 40, 60, 300, 0.01, true, c8, 0.04, false, 15, 0, {8, 9, 10, 11, 12, 13, 14, 15, 16}},
		{"two_part_gapped_uniform", 50, 400, 400, 0.005, false, {0, 1, 2, 3, 5, 6, 7}, 0.05, false, 16, 0, {9, 10, 12}},
	};
	int failed = 0, ran = 0;
	if (argc <= 1 || std::string(argv[1]) == "tile_scan") {
		// k_tile_scan on its own with more than one block (a set that large is beyond the emulator): 2500 tiles
		std::mt19937 rng(99);
		const uint32_t ntiles = 2500;
		std::vector<uint32_t> cnt(ntiles);
		for (auto& v : cnt) v = rng() % 70000u;
		std::vector<uint64_t> off(ntiles + 2, ~0ull);
		uint64_t total = ~0ull;
		emu::launch((ntiles + 1023) / 1024, 1024, [&] { k_tile_scan(cnt.data(), ntiles, off.data(), &total); });
		uint64_t acc = 0;
		bool ok = true;
		for (uint32_t i = 0; i < ntiles; ++i) { ok = ok && off[i] == acc; acc += cnt[i]; }
		ok = ok && off[ntiles] == acc && total == acc;
		std::printf("%-22s %u tiles in 3 blocks  %s\n", "tile_scan", ntiles, ok ? "identical" : "DIFFERENT");
		if (!ok) ++failed;
		++ran;
	}
	for (const Case& c : cases) {
		if (argc > 1) {
			bool sel = false;
			for (int i = 1; i < argc; ++i) sel = sel || c.name == argv[i];
			if (!sel) continue;
		}
		const Set s = make_set(c);
		const Table t = make_table(c);
		Right right;
		const Right* RT = nullptr;
		if (!c.rightOff.empty()) { right = make_right(c, t.nclass); RT = &right; }
		const std::vector<Rec> want = brute_force(c, s, t, RT);
		// "small_capacity": the record arrays hold fewer records than the scan finds (the caller then grows them and repeats)
		const size_t cap = (c.name == "small_capacity") ? want.size() / 2 + 3 : want.size() + 100;
		for (int gen : {6, 7}) {
			for (unsigned grid : {1u, 3u}) {
				const Result r = run_scan(c, s, t, gen, grid, cap, RT);
				const bool ok = check(c, gen, want, r, cap);
				std::printf("%-22s gen %d find grid %u: %6zu sites over %7u words, %8llu collectives  %s\n", c.name.c_str(), gen, grid, want.size(), s.nwords,
				            static_cast<unsigned long long>(r.collectives), ok ? "identical" : "DIFFERENT");
				if (!ok) ++failed;
				++ran;
			}
		}
	}
	std::printf("%d runs, %d failed\n", ran, failed);
	return failed ? 1 : (ran ? 0 : 3);
}
