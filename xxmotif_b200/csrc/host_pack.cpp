// Host-side packing of the sequence set into the device layout (host_pack.h).
#include "host_pack.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace xxm {
namespace {

// bit i of x -> bit 2 i
inline uint64_t spread32(uint32_t v) {
	uint64_t x = v;
	x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
	x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
	x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
	x = (x | (x << 2)) & 0x3333333333333333ull;
	x = (x | (x << 1)) & 0x5555555555555555ull;
	return x;
}

// Running state of one sequence: N count (bytes of nacc count down by one per N, flushed before they can wrap) and the
// OR of everything that marks a code outside 0..4.
struct PackAcc {
#if defined(__SSE2__)
	__m128i nacc = _mm_setzero_si128(), badv = _mm_setzero_si128();
	uint32_t pending = 0;
#endif
	uint64_t nN = 0;
	uint32_t bad = 0;
	void flush() {
#if defined(__SSE2__)
		const __m128i s = _mm_sad_epu8(_mm_sub_epi8(_mm_setzero_si128(), nacc), _mm_setzero_si128());
		nN += static_cast<uint64_t>(_mm_cvtsi128_si64(s)) + static_cast<uint64_t>(_mm_cvtsi128_si64(_mm_srli_si128(s, 8)));
		nacc = _mm_setzero_si128();
		pending = 0;
		bad |= static_cast<uint32_t>(_mm_movemask_epi8(badv));
#endif
	}
};

// 32 raw codes -> code word (code - 1 at bits 2j, 0 for N) and N mask (bit j: code 0)
inline void pack32(const uint8_t* r, uint64_t* c, uint32_t* m, PackAcc& A) {
#if defined(__SSE2__)
	const __m128i zero = _mm_setzero_si128(), one = _mm_set1_epi8(1), three = _mm_set1_epi8(3), four = _mm_set1_epi8(4);
	const __m128i v0 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(r));
	const __m128i v1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(r + 16));
	const __m128i z0 = _mm_cmpeq_epi8(v0, zero), z1 = _mm_cmpeq_epi8(v1, zero);
	*m = static_cast<uint32_t>(_mm_movemask_epi8(z0)) | (static_cast<uint32_t>(_mm_movemask_epi8(z1)) << 16);
	A.nacc = _mm_add_epi8(A.nacc, _mm_add_epi8(z0, z1));  // 0xFF = -1 per N
	// codes > 4 (signed compare) or >= 128 (sign bit)
	A.badv = _mm_or_si128(A.badv, _mm_or_si128(_mm_or_si128(_mm_cmpgt_epi8(v0, four), v0), _mm_or_si128(_mm_cmpgt_epi8(v1, four), v1)));
	if (++A.pending == 120) A.flush();
	// code - 1 with N -> 0 (saturating), then four 2-bit values per byte: pairs inside 16-bit lanes, nibbles inside 32-bit
	// lanes, and the low bytes of the eight 32-bit lanes packed into 64 bits
	__m128i x0 = _mm_and_si128(_mm_subs_epu8(v0, one), three), x1 = _mm_and_si128(_mm_subs_epu8(v1, one), three);
	x0 = _mm_or_si128(x0, _mm_srli_epi16(x0, 6));
	x1 = _mm_or_si128(x1, _mm_srli_epi16(x1, 6));
	const __m128i lowNib = _mm_set1_epi32(0x000F000F);
	x0 = _mm_and_si128(x0, lowNib);
	x1 = _mm_and_si128(x1, lowNib);
	x0 = _mm_or_si128(x0, _mm_srli_epi32(x0, 12));
	x1 = _mm_or_si128(x1, _mm_srli_epi32(x1, 12));
	const __m128i lowByte = _mm_set1_epi32(0x000000FF);
	const __m128i p = _mm_packs_epi32(_mm_and_si128(x0, lowByte), _mm_and_si128(x1, lowByte));
	*c = static_cast<uint64_t>(_mm_cvtsi128_si64(_mm_packus_epi16(p, zero)));
#else
	uint64_t cc = 0;
	uint32_t nm = 0;
	for (int j = 0; j < 32; ++j) {
		const uint32_t code = r[j];
		if (code == 0) { nm |= 1u << j; ++A.nN; }
		else cc |= static_cast<uint64_t>((code - 1) & 3u) << (2 * j);
		A.bad |= (code > 4);
	}
	*m = nm;
	*c = cc;
#endif
}

// order of the 32 two-bit groups of x reversed
inline uint64_t reverse_groups(uint64_t x) {
	x = __builtin_bswap64(x);
	x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
	x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
	return x;
}
inline uint32_t reverse_bits(uint32_t x) {
	x = __builtin_bswap32(x);
	x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
	x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
	x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
	return x;
}

}  // namespace

uint32_t pack_sequence(const uint8_t* raw, uint32_t n0, bool revcomp, uint64_t* codes, uint32_t* nmask, bool* bad) {
	const uint64_t words = stored_words(n0, revcomp);
	const uint32_t fullBlocks = n0 / 32, tail = n0 % 32;
	const uint32_t fwdWords = fullBlocks + (tail ? 1u : 0u);
	PackAcc A;
	for (uint32_t k = 0; k < fullBlocks; ++k) pack32(raw + 32ull * k, codes + k, nmask + k, A);
	if (tail) {
		uint8_t buf[32];
		std::memset(buf, 0, sizeof buf);
		std::memcpy(buf, raw + 32ull * fullBlocks, tail);
		pack32(buf, codes + fullBlocks, nmask + fullBlocks, A);
	}
	A.flush();
	const uint32_t nN = static_cast<uint32_t>(A.nN) - (tail ? 32u - tail : 0u);  // the zero padding of the tail block counted as N
	if (A.bad) *bad = true;
	if (!revcomp) {
		for (uint64_t k = fwdWords; k < words; ++k) { codes[k] = 0; nmask[k] = 0xFFFFFFFFu; }
		return n0 - nN;
	}
	// The doubled sequence, 0-based position p: p < n0 forward, p == n0 the separating N, n0 < p <= 2 n0 the complement
	// of forward position f = 2 n0 - p.  The 32 positions of output word k therefore read the forward positions
	// [fLo, fLo + 32), fLo = 2 n0 - 32 k - 31, in reverse order; everything outside [0, n0) reads as N, which makes the
	// separator, the padding and the words of the pure forward part fall out of the same formula.
	auto fwdCode = [&](int64_t k) -> uint64_t { return (k >= 0 && k < static_cast<int64_t>(fwdWords)) ? codes[k] : 0ull; };
	auto fwdMask = [&](int64_t k) -> uint32_t { return (k >= 0 && k < static_cast<int64_t>(fwdWords)) ? nmask[k] : 0xFFFFFFFFu; };
	// In place: the reverse part reads forward words (indices < fwdWords) and writes words >= firstRc.  The only word that
	// is both is the boundary word firstRc (when n0 is not a multiple of 32); walking k downwards writes it last, after
	// its own reads.
	const int64_t firstRc = static_cast<int64_t>(n0 / 32);  // first output word that holds a position >= n0
	for (int64_t k = static_cast<int64_t>(words) - 1; k >= firstRc; --k) {
		const int64_t fLo = 2 * static_cast<int64_t>(n0) - 32 * k - 31;
		const int64_t a = (fLo >= 0) ? fLo / 32 : -((-fLo + 31) / 32);
		const int b = static_cast<int>(fLo - 32 * a);
		uint64_t cw = fwdCode(a) >> (2 * b);
		uint32_t mw = fwdMask(a) >> b;
		if (b) {
			cw |= fwdCode(a + 1) << (64 - 2 * b);
			mw |= fwdMask(a + 1) << (32 - b);
		}
		const uint32_t rm = reverse_bits(mw);
		const uint64_t keep = spread32(~rm) * 3ull;
		const uint64_t rc = reverse_groups(~cw) & keep;
		if (k < static_cast<int64_t>(fwdWords)) {  // boundary word: forward positions and reverse positions share it
			codes[k] |= rc;
			nmask[k] &= rm;
		} else {
			codes[k] = rc;
			nmask[k] = rm;
		}
	}
	return n0 - nN;
}

uint32_t pack_range(const int32_t* n, uint32_t s0, uint32_t s1, const uint8_t* raw, bool revcomp, uint64_t* codes, uint32_t* nmask, int32_t* notX) {
	uint32_t badSeq = 0;
	for (uint32_t s = s0; s < s1; ++s) {
		const uint32_t n0 = static_cast<uint32_t>(n[s]);
		bool bad = false;
		notX[s - s0] = static_cast<int32_t>(pack_sequence(raw, n0, revcomp, codes, nmask, &bad));
		if (bad && !badSeq) badSeq = s + 1;
		const uint64_t w = stored_words(n0, revcomp);
		raw += n0;
		codes += w;
		nmask += w;
	}
	return badSeq;
}

int pack_threads_default() {
	if (const char* e = std::getenv("XXM_PACK_THREADS")) {
		const int t = std::atoi(e);
		if (t >= 1) return std::min(t, 256);
	}
	const unsigned hc = std::thread::hardware_concurrency();
	return static_cast<int>(std::max(1u, std::min(16u, hc ? hc : 1u)));
}

}  // namespace xxm
