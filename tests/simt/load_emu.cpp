// TEST INFRASTRUCTURE ONLY (see simt_emu.h): the load-time kernels of xxm_load_seqset (xxm_load.cuh) from their CUDA SOURCES under the
// SIMT emulator -- k_seq_layout_sums + k_seq_layout (rawOff / wordOff / stored length of every sequence from the raw lengths alone: a
// two-launch scan over blocks of 1024 sequences; calculateSequenceFeatures' totals and createRevcomp's n <- 2n + 1, seqset.cpp:21-36,
// :128-178), k_blk2seq (coarse word -> sequence index) and k_seq_stats (residues that are not N per sequence, codes outside 0..4
// flagged; 16-byte loads once the pointer is aligned) -- against the same quantities computed sequence by sequence.
//
// usage: load_emu [case ...]; exit code 0 = every case identical
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "cuda_runtime.h"  // tests/simt/shim -> simt_emu.h

#include "xxm_load.cuh"

#include "host_pack.h"

using namespace xxm;

namespace {

struct Case {
	std::string name;
	uint32_t nent;
	int lenLo, lenHi;
	bool revcomp;
	double nFrac;
	int badAt;   // >= 0: this sequence carries a code outside 0..4
	uint32_t seed;
};

int run(const Case& c) {
	std::mt19937 rng(c.seed);
	std::vector<int32_t> n(c.nent);
	for (auto& v : n) v = c.lenLo + static_cast<int>(rng() % static_cast<uint32_t>(c.lenHi - c.lenLo + 1));
	// ---- expectation, sequence by sequence
	std::vector<uint64_t> wantRaw(c.nent + 1);
	std::vector<uint32_t> wantWord(c.nent + 1);
	std::vector<int32_t> wantLen(c.nent);
	uint64_t r = 0, w = 0;
	for (uint32_t i = 0; i < c.nent; ++i) {
		wantRaw[i] = r; wantWord[i] = static_cast<uint32_t>(w);
		wantLen[i] = c.revcomp ? 2 * n[i] + 1 : n[i];
		r += static_cast<uint64_t>(n[i]);
		w += stored_words(static_cast<uint32_t>(n[i]), c.revcomp);
	}
	wantRaw[c.nent] = r; wantWord[c.nent] = static_cast<uint32_t>(w);
	// ---- the two layout kernels (xxm_load_seqset)
	const uint32_t nb = (c.nent + 1023) / 1024;
	std::vector<uint64_t> sums(2ull * nb, ~0ull), rawOff(c.nent + 1, ~0ull);
	std::vector<uint32_t> wordOff(c.nent + 1, ~0u);
	std::vector<int32_t> len(c.nent, -1);
	emu::launch(nb, 1024, [&] { k_seq_layout_sums(n.data(), c.nent, c.revcomp ? 1 : 0, sums.data()); });
	emu::launch(nb, 1024, [&] { k_seq_layout(n.data(), c.nent, c.revcomp ? 1 : 0, sums.data(), rawOff.data(), wordOff.data(), len.data()); });
	bool same = rawOff == wantRaw && wordOff == wantWord && len == wantLen;
	// ---- k_blk2seq
	const uint32_t nwords = wantWord[c.nent];
	const uint32_t nblk = nwords / kBlkWords + 1;
	std::vector<uint32_t> blk(nblk, ~0u), wantBlk(nblk, 0u);
	for (uint32_t b = 0, s = 0; b < nblk; ++b) {
		while (s + 1 < c.nent && wantWord[s + 1] <= b * kBlkWords) ++s;
		wantBlk[b] = s;
	}
	emu::launch((c.nent + 255) / 256, 256, [&] { k_blk2seq(wordOff.data(), c.nent, nblk, blk.data()); });
	const bool sameBlk = blk == wantBlk;
	// ---- k_seq_stats over two chunks of sequences, raw array of exactly the residues
	std::vector<uint8_t> raw(std::max<uint64_t>(1, r));
	for (auto& v : raw) v = static_cast<uint8_t>((rng() % 1000 < c.nFrac * 1000) ? 0 : 1 + rng() % 4);
	if (c.badAt >= 0 && n[static_cast<size_t>(c.badAt)] > 0) raw[wantRaw[static_cast<size_t>(c.badAt)] + static_cast<uint64_t>(n[static_cast<size_t>(c.badAt)]) / 2] = 7;
	std::vector<int32_t> notX(c.nent, -1), wantNotX(c.nent);
	for (uint32_t i = 0; i < c.nent; ++i) {
		int cnt = 0;
		for (uint64_t k = wantRaw[i]; k < wantRaw[i + 1]; ++k) cnt += raw[k] != 0;
		wantNotX[i] = cnt;
	}
	uint32_t bad = 0;
	const uint32_t mid = c.nent / 2;
	if (mid) emu::launch(std::min<uint32_t>((mid + 7) / 8, 4), 256, [&] { k_seq_stats(raw.data(), rawOff.data(), 0, mid, notX.data(), &bad); });
	emu::launch(std::min<uint32_t>((c.nent - mid + 7) / 8, 4), 256, [&] { k_seq_stats(raw.data(), rawOff.data(), mid, c.nent, notX.data(), &bad); });
	const bool sameStats = notX == wantNotX && bad == (c.badAt >= 0 ? static_cast<uint32_t>(c.badAt) + 1 : 0u);
	std::printf("%-22s %u sequences%s, %llu residues, %u words: layout %s, blk2seq (%u blocks) %s, residue statistics %s\n", c.name.c_str(), c.nent,
	            c.revcomp ? " --revcomp" : "", static_cast<unsigned long long>(r), nwords, same ? "identical" : "DIFFERENT", nblk, sameBlk ? "identical" : "DIFFERENT",
	            sameStats ? "identical" : "DIFFERENT");
	return (same && sameBlk && sameStats) ? 0 : 1;
}

}  // namespace

int main(int argc, char** argv) {
	const std::vector<Case> cases = {
		{"one_sequence", 1, 77, 77, false, 0.05, -1, 71},
		{"one_block_less_one", 1023, 1, 300, false, 0.03, -1, 72},
		{"exactly_one_block", 1024, 20, 20, true, 0.0, -1, 73},
		{"one_more", 1025, 1, 90, true, 0.05, -1, 74},
		{"five_blocks_ragged", 5000, 1, 400, false, 0.02, -1, 75},
		{"long_sequences", 40, 3000, 9000, true, 0.01, -1, 76},
		{"bad_code", 300, 10, 200, false, 0.02, 137, 77},
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
