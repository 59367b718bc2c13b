// a1 on the host side of the bus: the reference's byte-per-residue sequences (Alignment.cpp:149-195: 0 = N, 1..4 = ACGT)
// are packed into the device layout (xxm_device.cuh: one 64-bit word of 2-bit codes and one 32-bit N mask per 32 stored
// positions, every sequence on a word boundary and followed by >= 32 positions flagged N) by the host cores, so that
// 0.375 bytes per base cross PCIe instead of 1.  createRevcomp (seqset.cpp:128-178: S[n+1] = N, S[n+1+k] =
// complement(S[n+1-k])) is applied on the packed words.  Pure integer code; the result is bit-identical to packing on
// the device from the raw bytes.
#pragma once
#include <cstdint>

namespace xxm {

constexpr int kHostPadBases = 32;  // == kPadBases (xxm_device.cuh)

inline uint64_t stored_words(uint32_t n0, bool revcomp) {
	const uint64_t ns = revcomp ? 2ull * n0 + 1 : n0;
	return (ns + kHostPadBases + 31) / 32;
}

// Packs one sequence of n0 raw codes into `words` = stored_words(n0, revcomp) words at codes[] / nmask[].
// Returns the number of residues that are not N; *bad is set if a code lies outside 0..4.  canOverread: the 32 bytes of
// the last, partial block of the sequence may be read (they belong to the following sequences of the same array).
uint32_t pack_sequence(const uint8_t* raw, uint32_t n0, bool revcomp, uint64_t* codes, uint32_t* nmask, bool* bad, bool canOverread = false);

// Packs sequences [s0, s1): raw codes of sequence s0 start at raw (rawAvail bytes remain in the array from there), its
// words at codes / nmask; notX[s - s0] receives the residue counts.  Returns the 1-based index of a sequence with a code outside 0..4, or 0.
uint32_t pack_range(const int32_t* n, uint32_t s0, uint32_t s1, const uint8_t* raw, uint64_t rawAvail, bool revcomp, uint64_t* codes, uint32_t* nmask, int32_t* notX);

int pack_threads_default();

}  // namespace xxm
