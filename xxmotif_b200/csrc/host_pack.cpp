// Host-side packing of the sequence set into the device layout (host_pack.h).
#include "host_pack.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#if defined(__x86_64__)
#include <immintrin.h>
#define XXM_HAVE_AVX2_DISPATCH 1
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

#if defined(XXM_HAVE_AVX2_DISPATCH)
// The same packing with 256-bit registers (one register per 32 bases), chosen at run time when the CPU has AVX2.
alignas(32) const uint8_t kTailMask[64] = {255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255,
                                          255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,
                                          0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0};
// nblocks full blocks, then -- if tail != 0 -- one block of which only the first `tail` bytes belong to the sequence (the
// caller guarantees that 32 bytes can be read there); the bytes beyond them read as N
// kStream: the words are written past the caches (nobody reads them before the DMA engine does; the staging buffer is
// larger than the last-level cache, and a normal store would first read every line it is about to overwrite)
template <bool kStream>
__attribute__((target("avx2"))) void pack_blocks_avx2(const uint8_t* r, uint32_t nblocks, uint32_t tail, uint64_t* c, uint32_t* m, uint64_t* nN, uint32_t* bad) {
	const __m256i zero = _mm256_setzero_si256(), one = _mm256_set1_epi8(1), three = _mm256_set1_epi8(3), four = _mm256_set1_epi8(4);
	const __m256i lowNib = _mm256_set1_epi32(0x000F000F), lowByte = _mm256_set1_epi32(0x000000FF);
	// bytes 0, 4, 8, 12 of each 128-bit lane -> its low dword
	const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
	__m256i nacc = zero, badv = zero, total = zero;
	uint32_t pending = 0;
	const uint32_t nall = nblocks + (tail ? 1u : 0u);
	for (uint32_t k = 0; k < nall; ++k) {
		if ((k & 1u) == 0) _mm_prefetch(reinterpret_cast<const char*>(r + 32ull * k + 1024), _MM_HINT_T0);
		__m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(r + 32ull * k));
		if (k == nblocks) v = _mm256_and_si256(v, _mm256_loadu_si256(reinterpret_cast<const __m256i*>(kTailMask + 32 - tail)));
		const __m256i z = _mm256_cmpeq_epi8(v, zero);
		const uint32_t mk = static_cast<uint32_t>(_mm256_movemask_epi8(z));
		if (kStream) _mm_stream_si32(reinterpret_cast<int*>(m + k), static_cast<int>(mk));
		else m[k] = mk;
		nacc = _mm256_sub_epi8(nacc, z);  // + 1 per N
		badv = _mm256_or_si256(badv, _mm256_or_si256(_mm256_cmpgt_epi8(v, four), v));
		if (++pending == 255) { total = _mm256_add_epi64(total, _mm256_sad_epu8(nacc, zero)); nacc = zero; pending = 0; }
		__m256i x = _mm256_and_si256(_mm256_subs_epu8(v, one), three);
		x = _mm256_and_si256(_mm256_or_si256(x, _mm256_srli_epi16(x, 6)), lowNib);
		x = _mm256_and_si256(_mm256_or_si256(x, _mm256_srli_epi32(x, 12)), lowByte);
		x = _mm256_shuffle_epi8(x, gather);
		const uint64_t ck = static_cast<uint32_t>(_mm256_extract_epi32(x, 0)) | (static_cast<uint64_t>(static_cast<uint32_t>(_mm256_extract_epi32(x, 4))) << 32);
		if (kStream) _mm_stream_si64(reinterpret_cast<long long*>(c + k), static_cast<long long>(ck));
		else c[k] = ck;
	}
	total = _mm256_add_epi64(total, _mm256_sad_epu8(nacc, zero));
	*nN += static_cast<uint64_t>(_mm256_extract_epi64(total, 0)) + static_cast<uint64_t>(_mm256_extract_epi64(total, 1)) +
	       static_cast<uint64_t>(_mm256_extract_epi64(total, 2)) + static_cast<uint64_t>(_mm256_extract_epi64(total, 3));
	*bad |= static_cast<uint32_t>(_mm256_movemask_epi8(badv));
}
const bool g_haveAvx2 = __builtin_cpu_supports("avx2") && std::getenv("XXM_PACK_NO_AVX2") == nullptr;

// 512-bit registers, 64 bases (two words) per step, chosen at run time when the CPU has AVX-512 BW: the N mask of 64 bases IS the
// compare mask (no movemask), the four 2-bit codes of a byte group are folded by two multiply-adds (b0 + 4 b1, then + 16 (b2 + 4 b3))
// and a truncating move; the tail of a sequence is a masked load, so nothing is read beyond the sequence.
template <bool kStream>
__attribute__((target("avx512f,avx512bw,avx512vl,popcnt,sse4.1"))) void pack_blocks_avx512(const uint8_t* r, uint32_t n0, uint64_t* c, uint32_t* m, uint64_t* nN, uint32_t* bad) {
	const __m512i zero = _mm512_setzero_si512(), one = _mm512_set1_epi8(1), three = _mm512_set1_epi8(3), four = _mm512_set1_epi8(4);
	const __m512i mul4 = _mm512_set1_epi16(0x0401), mul16 = _mm512_set1_epi32(0x00100001);
	const uint32_t nwords = (n0 + 31) / 32;
	uint64_t ncount = 0;
	__mmask64 badm = 0;
	uint32_t w = 0;
	for (uint64_t off = 0; off < n0; off += 64, w += 2) {
		const uint64_t rem = n0 - off;
		__m512i v;
		if (rem >= 64) {
			_mm_prefetch(reinterpret_cast<const char*>(r + off + 1024), _MM_HINT_T0);
			v = _mm512_loadu_si512(reinterpret_cast<const void*>(r + off));
		} else {
			v = _mm512_maskz_loadu_epi8((1ull << rem) - 1ull, reinterpret_cast<const void*>(r + off));  // beyond the sequence: reads as N
		}
		const __mmask64 z = _mm512_cmpeq_epi8_mask(v, zero);
		badm |= _mm512_cmpgt_epu8_mask(v, four);
		ncount += static_cast<uint64_t>(__builtin_popcountll(static_cast<unsigned long long>(z)));
		const __m512i x = _mm512_and_si512(_mm512_subs_epu8(v, one), three);
		const __m512i y = _mm512_madd_epi16(_mm512_maddubs_epi16(x, mul4), mul16);
		const __m128i packed = _mm512_cvtepi32_epi8(y);  // 16 bytes: code words w and w + 1
		const uint64_t zm = static_cast<uint64_t>(z);
		const long long c0 = _mm_cvtsi128_si64(packed), c1 = _mm_extract_epi64(packed, 1);
		const bool two = w + 2 <= nwords;  // else: an odd last word
		if (kStream) {
			_mm_stream_si64(reinterpret_cast<long long*>(c + w), c0);
			_mm_stream_si32(reinterpret_cast<int*>(m + w), static_cast<int>(static_cast<uint32_t>(zm)));
			if (two) {
				_mm_stream_si64(reinterpret_cast<long long*>(c + w + 1), c1);
				_mm_stream_si32(reinterpret_cast<int*>(m + w + 1), static_cast<int>(static_cast<uint32_t>(zm >> 32)));
			}
		} else {
			c[w] = static_cast<uint64_t>(c0);
			m[w] = static_cast<uint32_t>(zm);
			if (two) { c[w + 1] = static_cast<uint64_t>(c1); m[w + 1] = static_cast<uint32_t>(zm >> 32); }
		}
	}
	const uint64_t padded = static_cast<uint64_t>((n0 + 63) / 64) * 64 - n0;  // positions beyond the sequence counted as N above
	*nN += ncount - padded;
	if (badm) *bad |= 1u;
}
const bool g_haveAvx512 = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") && std::getenv("XXM_PACK_NO_AVX512") == nullptr &&
                          std::getenv("XXM_PACK_NO_AVX2") == nullptr;
#endif

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

uint32_t pack_sequence(const uint8_t* raw, uint32_t n0, bool revcomp, uint64_t* codes, uint32_t* nmask, bool* bad, bool canOverread) {
	const uint64_t words = stored_words(n0, revcomp);
	const uint32_t fullBlocks = n0 / 32, tail = n0 % 32;
	const uint32_t fwdWords = fullBlocks + (tail ? 1u : 0u);
	PackAcc A;
#if defined(XXM_HAVE_AVX2_DISPATCH)
	if (g_haveAvx512) {
		uint64_t nN = 0;
		uint32_t b = 0;
		if (revcomp) {
			pack_blocks_avx512<false>(raw, n0, codes, nmask, &nN, &b);  // the reverse strand reads these words back
			A.nN = nN + (tail ? 32u - tail : 0u);  // the common code below takes the padding of the tail word off again
			A.bad = b;
		} else {
			pack_blocks_avx512<true>(raw, n0, codes, nmask, &nN, &b);
			for (uint64_t k = fwdWords; k < words; ++k) {
				_mm_stream_si64(reinterpret_cast<long long*>(codes + k), 0);
				_mm_stream_si32(reinterpret_cast<int*>(nmask + k), -1);
			}
			if (b) *bad = true;
			return n0 - static_cast<uint32_t>(nN);
		}
	} else if (g_haveAvx2 && (tail == 0 || canOverread)) {
		if (revcomp) {
			pack_blocks_avx2<false>(raw, fullBlocks, tail, codes, nmask, &A.nN, &A.bad);  // the reverse strand reads these words back
		} else {
			pack_blocks_avx2<true>(raw, fullBlocks, tail, codes, nmask, &A.nN, &A.bad);
			for (uint64_t k = fwdWords; k < words; ++k) {
				_mm_stream_si64(reinterpret_cast<long long*>(codes + k), 0);
				_mm_stream_si32(reinterpret_cast<int*>(nmask + k), -1);
			}
			A.flush();
			if (A.bad) *bad = true;
			return n0 - (static_cast<uint32_t>(A.nN) - (tail ? 32u - tail : 0u));
		}
	} else
#endif
	{
		for (uint32_t k = 0; k < fullBlocks; ++k) pack32(raw + 32ull * k, codes + k, nmask + k, A);
		if (tail) {
			uint8_t buf[32];
			std::memset(buf, 0, sizeof buf);
			std::memcpy(buf, raw + 32ull * fullBlocks, tail);
			pack32(buf, codes + fullBlocks, nmask + fullBlocks, A);
		}
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

uint32_t pack_range(const int32_t* n, uint32_t s0, uint32_t s1, const uint8_t* raw, uint64_t rawAvail, bool revcomp, uint64_t* codes, uint32_t* nmask, int32_t* notX) {
	uint32_t badSeq = 0;
	for (uint32_t s = s0; s < s1; ++s) {
		const uint32_t n0 = static_cast<uint32_t>(n[s]);
		bool bad = false;
		// the tail block of a sequence may be read as a full 32 bytes when the array still holds that many
		const bool canOverread = (n0 / 32) * 32ull + 32 <= rawAvail;
		notX[s - s0] = static_cast<int32_t>(pack_sequence(raw, n0, revcomp, codes, nmask, &bad, canOverread));
		if (bad && !badSeq) badSeq = s + 1;
		const uint64_t w = stored_words(n0, revcomp);
		raw += n0;
		rawAvail -= n0;
		codes += w;
		nmask += w;
	}
#if defined(__SSE2__)
	_mm_sfence();  // streaming stores visible before the chunk is reported as packed
#endif
	return badSeq;
}

int pack_threads_default() {
	if (const char* e = std::getenv("XXM_PACK_THREADS")) {
		const int t = std::atoi(e);
		if (t >= 1) return std::min(t, 256);
	}
	const unsigned hc = std::thread::hardware_concurrency();
	// the calling thread and the driver's threads need cores too: two of them on a host with 12 or more, a quarter below that
	// (16 cores: 14 packers load 1 GB in 14.0-14.5 ms, 12 in 14.6-14.8 ms, 16 in 15.1 ms)
	if (hc >= 12) return static_cast<int>(std::min(14u, hc - 2));
	return static_cast<int>(std::max(1u, hc ? (3 * hc + 3) / 4 : 1u));
}

}  // namespace xxm
