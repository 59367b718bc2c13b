#include "pair_table.h"

#include <algorithm>
#include <cstring>

namespace xxm {

namespace {
struct Row256 {
	uint64_t w[4] = {0, 0, 0, 0};
	void set(uint32_t j) { w[j >> 6] |= 1ull << (j & 63); }
	int count() const { return __builtin_popcountll(w[0]) + __builtin_popcountll(w[1]) + __builtin_popcountll(w[2]) + __builtin_popcountll(w[3]); }
	bool operator==(const Row256& o) const { return w[0] == o.w[0] && w[1] == o.w[1] && w[2] == o.w[2] && w[3] == o.w[3]; }
	void operator|=(const Row256& o) { for (int k = 0; k < 4; ++k) w[k] |= o.w[k]; }
};
}  // namespace

bool build_pair_table(const uint32_t* ids, size_t nids, const int* coloff, int len0, uint32_t* table, int* levels) {
	*levels = 0;
	if (len0 < 1 || coloff[0] != 0 || coloff[len0 - 1] > 7) return false;
	int nL = 0;
	while (nL < len0 && coloff[nL] < 4) ++nL;
	const int nR = len0 - nL;
	const uint32_t rowsN = 1u << (2 * nL), colsN = 1u << (2 * nR);
	// membership matrix over (left match columns) x (right match columns)
	Row256 R[256];  // (on the stack: this runs once per refinement step)
	for (size_t i = 0; i < nids; ++i) R[ids[i] & (rowsN - 1)].set(ids[i] >> (2 * nL));
	int pc[256];
	int theta[256];
	bool present[257];
	std::memset(present, 0, sizeof present);
	for (uint32_t y = 0; y < colsN; ++y) theta[y] = 1 << 30;
	for (uint32_t x = 0; x < rowsN; ++x) {
		pc[x] = R[x].count();
		if (!pc[x]) continue;
		present[pc[x]] = true;
		for (int k = 0; k < 4; ++k) {
			uint64_t m = R[x].w[k];
			while (m) {
				const int y = 64 * k + __builtin_ctzll(m);
				m &= m - 1;
				theta[y] = std::min(theta[y], pc[x]);
			}
		}
	}
	int levelOf[257];  // 1..m for the distinct non-zero row sizes, ascending
	int m = 0;
	for (int c = 1; c <= 256; ++c) levelOf[c] = present[c] ? ++m : 0;
	if (m > 127) return false;
	// nestedness: the row of size-level k must be exactly {y : theta[y] <= that size}
	Row256 S[129];
	for (uint32_t y = 0; y < colsN; ++y)
		if (theta[y] <= 256) S[levelOf[theta[y]]].set(y);
	for (int k = 2; k <= m; ++k) S[k] |= S[k - 1];
	for (uint32_t x = 0; x < rowsN; ++x)
		if (pc[x] && !(R[x] == S[levelOf[pc[x]]])) return false;
	uint8_t aRed[256], bRed[256];
	for (uint32_t x = 0; x < rowsN; ++x) aRed[x] = pc[x] ? static_cast<uint8_t>(128 - (m - levelOf[pc[x]] + 1)) : 0;
	for (uint32_t y = 0; y < colsN; ++y) bRed[y] = (theta[y] <= 256) ? static_cast<uint8_t>(m - levelOf[theta[y]] + 1) : 0;
	// back to plain 4-mers (gap positions are "don't care")
	uint8_t a[256], b[256];
	for (uint32_t v = 0; v < 256; ++v) {
		uint32_t xr = 0, yr = 0;
		for (int c = 0; c < nL; ++c) xr |= ((v >> (2 * coloff[c])) & 3u) << (2 * c);
		for (int c = 0; c < nR; ++c) yr |= ((v >> (2 * (coloff[nL + c] - 4))) & 3u) << (2 * c);
		a[v] = aRed[xr];
		b[v] = bRed[yr];
	}
	for (uint32_t q = 0; q < kPairEntries; ++q) {
		const uint32_t x0 = q & 255u, x1 = (q >> 2) & 255u;
		table[q] = (static_cast<uint32_t>(a[x0]) << 24) | (static_cast<uint32_t>(b[x0]) << 16) | (static_cast<uint32_t>(a[x1]) << 8) | b[x1];
	}
	*levels = m;
	return true;
}

}  // namespace xxm
