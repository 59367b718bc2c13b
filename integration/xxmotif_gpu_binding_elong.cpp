/*
 * xxmotif_gpu_binding_elong.cpp -- SURVEY.md 8(f) row 4: ElongCandidates::countAllOccurrences (elongationPhase/
 * elongationCandidates.cpp:234-276) with its walk over every start position of every sequence replaced by one call into
 * libxxmotif_gpu.so (xxm_count_gapped_kmers: the 4^coding counts of the k-mers without IUPAC characters); the spreading of
 * those counts over the IUPAC degenerations (:264-273) stays the reference's own mergeIUPACS_rec.
 *
 * ElongCandidates is not a template, so the replacement happens at link time: oracle/Makefile compiles the reference's
 * elongationCandidates.cpp unchanged (position independent, so that the compiler keeps the call through the symbol), weakens
 * the symbol of the reference's definition with objcopy and links this strong one.  Compiled with -fno-access-control.
 */
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "xxmotif_gpu_binding.h"
#include "../include/xxmotif_gpu.h"
#include "elongationPhase/elongationCandidates.h"

void ElongCandidates::countAllOccurrences() {
	const uint8_t* gap = (type == PALINDROME || type == TANDEM) ? gaps_palindrome[_gapIndex] : gaps[_gapIndex];
	std::vector<int32_t> offset(coding_characters, 0);
	for (int l = 1; l < coding_characters; l++) offset[l] = l + gap[l - 1];
	std::vector<uint32_t> counts(POW_4_coding);
	if (xxm_multi_count_gapped_kmers(xxm_binding_context(), nbSequences, coding_characters, offset.data(), (type == PALINDROME && Global::revcomp) ? 1 : 0,
	                                 counts.data()))
		xxm_binding_die("xxm_multi_count_gapped_kmers");
	/* the library numbers k-mers with 2 bits per character, countIUPAC with 4 (first character most significant in both) */
	std::vector<uint32_t> wide(POW_4_coding);
	for (uint32_t id = 0; id < POW_4_coding; id++) {
		uint32_t w = 0;
		for (int l = 0; l < coding_characters; l++) w = (w << 4) | ((id >> (2 * (coding_characters - 1 - l))) & 3u);
		wide[id] = w;
		countIUPAC[w] += counts[id];
	}
	/* every plain k-mer hands its count to its IUPAC degenerations, in the order the reference visits them */
	for (uint32_t id = 0; id < POW_4_coding; id++) {
		uint32_t w = 0, t = id;
		for (int l = 0; l < coding_characters; l++) { w = (w << 4) + (t & 3u); t >>= 2; }
		mergeIUPACS_rec(countIUPAC, w, countIUPAC[w], 0, 0);
	}
}
