// TEST INFRASTRUCTURE ONLY (see simt_emu.h): the GENERIC find kernel -- k_build_t9 + k_scan4<true/false> (xxm_scan4.cuh), the one that
// runs whenever the pair table cannot represent the left table (stage_pair, xxm_gpu_scan5.inc: gapped left parts beyond 8 bases whose scan
// must find sites only, tables that are not threshold-shaped, XXM_FORCE_GENERIC_SCAN) --
// from the CUDA SOURCES under the SIMT emulator, staged as stage_t9 / launch_scan4 stage and launch it (xxm_gpu_scan4.inc), against
// the definition of a scan candidate taken window by window from the raw sequences: no N in any match column
// (branch_and_bound-inl.h:388) and the left k-mer in the table (:390-392).  Exact and projected table entries, bitmap in shared and
// in global memory, --revcomp, N runs, one-word sequences.  Then the generic EXPAND kernel k_expand5<true/false> (xxm_scan5.cuh: hit masks
// -> ordered (sequence, position, value index) records through a decoupled look-back over the tiles; the value index is the class of
// the left k-mer or, for two-part models, the memo slot of the window's integer score, kNone where the right table or the product
// test rejects the window, branch_and_bound-inl.h:399-423) against the same records taken window by window; record arrays of exactly
// the number of records.  Under build.py's sanitizers a memcheck and a racecheck.
//
// usage: scan4_emu [case ...]; exit code 0 = every case identical
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
#include "xxm_scan5.cuh"  // k_expand5

#include "host_pack.h"

namespace xxm {
alignas(16) uint8_t smem4[65536 + 65536];
alignas(16) uint8_t smem5[kPairTabWords * 4];
alignas(16) uint8_t smemM[1 << 16];
uint32_t s_hs[1 << 14];
}  // namespace xxm

using namespace xxm;

namespace {

struct Case {
	std::string name;
	int nseq, lenLo, lenHi;
	double nFrac;
	bool revcomp;
	std::vector<int> coloff;  // all L match columns (offsets from the window start)
	int len0;                 // columns of the left part (the table)
	double keepFrac;          // fraction of the 4^len0 k-mers in the table
	uint32_t seed;
};

struct Set {
	std::vector<int32_t> n, len;
	std::vector<uint8_t> raw;
	std::vector<std::vector<uint8_t>> seq;
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
		if (c.revcomp) {
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
	s.blk2seq.assign(s.nwords / kBlkWords + 1, 0);
	return s;
}

int run(const Case& c) {
	const Set s = make_set(c);
	std::mt19937 rng(c.seed * 31u + 7u);
	const int L = static_cast<int>(c.coloff.size());
	const size_t size = static_cast<size_t>(1) << (2 * c.len0);
	// the table: a random subset of the 4^len0 left k-mers
	std::vector<uint32_t> ids;
	std::vector<uint32_t> bm(std::max<size_t>(1, size / 32), 0u);
	for (size_t id = 0; id < size; ++id)
		if ((rng() % 100000) < c.keepFrac * 100000) { ids.push_back(static_cast<uint32_t>(id)); bm[id >> 5] |= 1u << (id & 31); }
	// ---- stage_t9 (xxm_gpu_scan4.inc): projections of the table on the columns visible from offsets 0 and 1 of a 9-base group
	T9Build B{};
	std::vector<uint32_t> pbits;
	uint32_t exact = 0;
	for (int k = 0; k < 2; ++k) {
		int cols[9], nv = 0;
		for (int cc = 0; cc < c.len0; ++cc) {
			const int o = c.coloff[cc] + k;
			if (o <= 8 && nv < 9) { B.visoff[k][nv] = o; cols[nv] = cc; ++nv; }
		}
		B.nvis[k] = nv;
		if (nv == c.len0) exact |= 1u << k;
		B.poff[k] = static_cast<uint32_t>(pbits.size());
		const size_t words = std::max<size_t>(1, (static_cast<size_t>(1) << (2 * nv)) / 32);
		const size_t o = pbits.size();
		pbits.resize(o + words, 0u);
		for (uint32_t id : ids) {
			uint32_t pid = 0;
			for (int r = 0; r < nv; ++r) pid |= ((id >> (2 * cols[r])) & 3u) << (2 * r);
			pbits[o + (pid >> 5)] |= 1u << (pid & 31);
		}
	}
	B.pbits = pbits.data();
	std::vector<uint8_t> t9(65536, 0xEE);
	emu::launch(65536 / 256, 256, [&] { k_build_t9(B, t9.data()); });
	Model M{};
	M.L = L; M.len0 = c.len0; M.nparts = (L > c.len0) ? 2 : 1;
	for (int k = 0; k < L; ++k) M.coloff[k] = c.coloff[k];
	M.span = c.coloff.back() + 1;
	M.t8 = t9.data(); M.hit0 = bm.data(); M.exactBits = exact;
	SeqStream S{};
	S.codes = s.codes.data(); S.nmask = s.nmask.data(); S.wordOff = s.wordOff.data(); S.len = s.len.data(); S.blk2seq = s.blk2seq.data();
	S.nwords = s.nwords; S.nent = c.nseq;
	// ---- launch_scan4 with a device of two SMs
	const uint32_t scanWords = s.nwords;
	const size_t bmBytes = size / 8;
	const bool smemBm = bmBytes <= 32 * 1024;
	uint32_t grid = 2 * 3;
	const uint32_t nwarps = grid * kScan4Warps;
	const uint32_t wordsPerWarp = std::max<uint32_t>(32, (((scanWords + nwarps - 1) / nwarps + 31) / 32) * 32);
	grid = std::max<uint32_t>(1, ((scanWords + wordsPerWarp - 1) / wordsPerWarp + kScan4Warps - 1) / kScan4Warps);
	std::vector<uint32_t> hitmask(scanWords, 0xDEADBEEFu);  // exactly the scanned words
	emu::launch(grid, kScan4Threads, [&] {
		if (smemBm) k_scan4<true>(S, M, hitmask.data(), scanWords, wordsPerWarp);
		else k_scan4<false>(S, M, hitmask.data(), scanWords, wordsPerWarp);
	});
	// ---- window by window
	std::vector<uint32_t> want(scanWords, 0u);
	unsigned long long cand = 0;
	for (int i = 0; i < c.nseq; ++i) {
		const auto& q = s.seq[i];
		const int n = static_cast<int>(q.size());
		for (int p = 0; p < n; ++p) {  // 0-based start; windows that run off the end die on the padding's N flags
			bool ok = p + M.span <= n;
			for (int k = 0; k < L && ok; ++k) ok = q[p + c.coloff[k]] != 0;
			if (!ok) continue;
			uint32_t id = 0;
			for (int k = 0; k < c.len0; ++k) id |= static_cast<uint32_t>(q[p + c.coloff[k]] - 1) << (2 * k);
			if (!((bm[id >> 5] >> (id & 31)) & 1u)) continue;
			want[s.wordOff[i] + static_cast<uint32_t>(p / 32)] |= 1u << (p % 32);
			++cand;
		}
	}
	const bool same = want == hitmask;
	std::printf("%-24s %d columns (left %d, span %d), table %zu of %zu, bitmap in %s memory, exact offsets %u, grid %u: %llu candidates over %u words  %s\n",
	            c.name.c_str(), L, c.len0, M.span, ids.size(), size, smemBm ? "shared" : "global", exact, grid, cand, scanWords, same ? "identical" : "DIFFERENT");
	if (!same) return 1;

	// ---- the generic expand on those masks
	const int len1 = L - c.len0;
	const uint32_t nclass0 = 50, nclass1 = 70;
	std::vector<uint32_t> rank0(size, 0u);
	for (size_t i = 0; i < ids.size(); ++i) rank0[ids[i]] = static_cast<uint32_t>(i % nclass0) + 1;
	std::vector<uint32_t> rank1;
	std::vector<double> cpval0(nclass0), cpval1(nclass1);
	std::uniform_real_distribution<double> U(1e-6, 1e-2);
	for (double& v : cpval0) v = U(rng);
	for (double& v : cpval1) v = U(rng);
	const uint32_t memoCap = 1u << 14;
	std::vector<int32_t> memoKey(memoCap, kMemoEmpty);
	uint32_t memoOverflow = 0;
	M.rank0 = rank0.data(); M.cpval0 = cpval0.data(); M.nclass0 = static_cast<int>(nclass0);
	M.len1 = len1;
	M.contiguous = (M.span == L) ? 1 : 0;
	if (len1 > 0) {
		rank1.assign(static_cast<size_t>(1) << (2 * len1), 0u);
		for (size_t id = 0; id < rank1.size(); ++id)
			if (rng() % 3) rank1[id] = static_cast<uint32_t>(rng() % nclass1) + 1;  // two thirds of the right k-mers are in the right table
		M.rank1 = rank1.data(); M.cpval1 = cpval1.data(); M.nclass1 = static_cast<int>(nclass1);
		M.pruneThr = 2.5e-5;  // about half of the products pass
		std::uniform_real_distribution<double> Lo(-1.5, 1.0);
		for (int k = 0; k < L; ++k)
			for (int a = 0; a < 4; ++a) M.logodds[k][a] = Lo(rng);
		M.memoKey = memoKey.data(); M.memoCap = memoCap; M.memoOverflow = &memoOverflow;
	}
	// the records the reference's loop would make of the candidates, in scan order
	struct Want { uint32_t seq, pos; bool none; uint32_t v; int32_t s; };
	std::vector<Want> recs;
	for (int i = 0; i < c.nseq; ++i) {
		const auto& q = s.seq[i];
		const int n = static_cast<int>(q.size());
		for (int p0 = 0; p0 < n; ++p0) {
			if (!((want[s.wordOff[i] + static_cast<uint32_t>(p0 / 32)] >> (p0 % 32)) & 1u)) continue;
			uint32_t id0 = 0, id1 = 0;
			for (int k = 0; k < c.len0; ++k) id0 |= static_cast<uint32_t>(q[p0 + c.coloff[k]] - 1) << (2 * k);
			for (int k = c.len0; k < L; ++k) id1 |= static_cast<uint32_t>(q[p0 + c.coloff[k]] - 1) << (2 * (k - c.len0));
			Want w{static_cast<uint32_t>(i + 1), static_cast<uint32_t>(p0 + 1), false, rank0[id0] - 1, 0};
			if (len1 > 0) {
				const uint32_t r0 = rank0[id0], r1 = rank1[id1];
				if (r1 == 0 || cpval0[r0 - 1] * cpval1[r1 - 1] > M.pruneThr) w.none = true;
				else {
					double score = 0;
					for (int k = 0; k < L; ++k) score = score + M.logodds[k][q[p0 + c.coloff[k]] - 1];
					w.s = static_cast<int32_t>(score * 1e6 - 1.0);
				}
			}
			recs.push_back(w);
		}
	}
	const uint32_t nrec = static_cast<uint32_t>(recs.size());
	const uint32_t ntiles = (scanWords + kTileWords - 1) / kTileWords;
	std::vector<uint64_t> status(static_cast<size_t>(ntiles) + 1, 0ull), gstatus(static_cast<size_t>(ntiles) / 32 + 2, 0ull);
	uint32_t ticket = 0;
	uint64_t total = ~0ull;
	const uint32_t nval = len1 > 0 ? memoCap : nclass0;
	std::vector<double> pS(nval);
	for (double& v : pS) v = U(rng);
	Values V{};
	V.pS = pS.data(); V.kind = kFinalTable;
	int bad = 0;
	// run 0, 1: tile = block index, exactly enough room / five records too few (they must be dropped, not written); run 2, 3: tile = ticket
	// (kTicket, XXM_EXPAND5_TICKET), blocks in order and in REVERSE order -- the ticket variant must not care in which order blocks start
	for (int variant = 0; variant < 4; ++variant) {
		const uint32_t cap = (variant == 1 && nrec > 8) ? nrec - 5 : nrec;
		const bool ticketed = variant >= 2;
		std::vector<uint32_t> oseq(std::max(1u, cap), 0xABABABABu), opos(oseq), ovidx(oseq);
		std::vector<float> oscore(std::max(1u, cap), -1.f);
		std::fill(status.begin(), status.end(), 0ull);
		std::fill(gstatus.begin(), gstatus.end(), 0ull);
		std::fill(memoKey.begin(), memoKey.end(), kMemoEmpty);
		Expand5IO io{};
		io.hitmask = hitmask.data(); io.scanWords = scanWords; io.eoFormat = 0; io.wantScore = 1;
		io.oseq = oseq.data(); io.opos = opos.data(); io.ovidx = ovidx.data(); io.oscore = oscore.data();
		io.cap = cap; io.status = status.data(); io.gstatus = gstatus.data(); io.ticket = &ticket; io.total = &total; io.ntiles = ntiles;
		io.wpsMagic = 0;
		const bool fast = M.nparts == 1 && M.contiguous;
		ticket = 0;
		emu::reverseBlocks = variant == 3;
		emu::launch(ntiles, kExpand5Threads, [&] {
			if (ticketed) {
				if (fast) k_expand5<true, true>(S, M, V, io);
				else k_expand5<false, true>(S, M, V, io);
			} else if (fast) k_expand5<true>(S, M, V, io);
			else k_expand5<false>(S, M, V, io);
		});
		emu::reverseBlocks = false;
		bool ok = total == nrec && memoOverflow == 0;
		for (uint32_t k = 0; ok && k < std::min(cap, nrec); ++k) {
			const Want& w = recs[k];
			ok = oseq[k] == w.seq && opos[k] == w.pos;
			if (!ok) break;
			if (w.none) ok = ovidx[k] == kNone && oscore[k] == 1.0f;
			else if (len1 > 0) ok = ovidx[k] < memoCap && memoKey[ovidx[k]] == w.s && oscore[k] == static_cast<float>(pS[ovidx[k]]);
			else ok = ovidx[k] == w.v && oscore[k] == static_cast<float>(pS[w.v]);
		}
		unsigned none = 0;
		for (const Want& w : recs) none += w.none;
		std::printf("%-24s k_expand5<%s%s> %u tiles%s, capacity %u: %u records (%u without a value)  %s\n", c.name.c_str(), fast ? "true" : "false",
		            ticketed ? ", ticket" : "", ntiles, variant == 3 ? " in reverse order" : "", cap, nrec, none, ok ? "identical" : "DIFFERENT");
		bad += ok ? 0 : 1;
	}
	return bad;
}

}  // namespace

int main(int argc, char** argv) {
	const std::vector<Case> cases = {
		{"contiguous_eight", 30, 40, 400, 0.03, false, {0, 1, 2, 3, 4, 5, 6, 7}, 8, 0.02, 51},
		{"gapped_span_ten", 30, 30, 300, 0.03, false, {0, 1, 2, 4, 5, 6, 8, 9}, 8, 0.03, 52},
		{"gapped_span_twelve_rc", 20, 20, 200, 0.05, true, {0, 2, 3, 5, 6, 8, 10, 11}, 8, 0.05, 53},
		{"two_part_twelve", 25, 50, 350, 0.02, false, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11}, 8, 0.02, 54},
		{"left_nine", 20, 60, 300, 0.02, false, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17}, 9, 0.01, 55},
		{"left_ten_global_bitmap", 16, 60, 260, 0.02, false, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19}, 10, 0.004, 56},
		{"five_columns_dense", 40, 1, 90, 0.08, false, {0, 1, 2, 3, 4}, 5, 0.30, 57},
		{"one_word_sequences", 70, 1, 31, 0.05, true, {0, 1, 3, 4, 5, 6}, 6, 0.20, 58},
		{"several_blocks_gapped", 150, 100, 700, 0.02, false, {0, 1, 2, 3, 5, 6, 7, 9}, 8, 0.02, 59},
		{"dense_tiles", 90, 300, 500, 0.01, false, {0, 1, 2, 3, 4}, 5, 0.50, 60},              // more hits per tile than the queue holds
		{"dense_tiles_gapped", 70, 300, 500, 0.01, false, {0, 1, 3, 4, 6}, 5, 0.45, 61},
		{"forty_tiles", 1300, 900, 1100, 0.002, false, {0, 1, 2, 3, 4, 5, 7, 8}, 8, 0.004, 62},  // the look-back crosses a group of 32 tiles
	};
	int runs = 0, failed = 0;
	for (const Case& c : cases) {
		bool wanted = argc <= 1;
		for (int i = 1; i < argc; ++i) wanted = wanted || c.name == argv[i];
		if (!wanted) continue;
		failed += run(c);
		++runs;
	}
	std::printf("%d runs, %d failed\n", runs, failed);
	return failed ? 1 : 0;
}
