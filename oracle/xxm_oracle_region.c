/*
 * xxm_oracle_region.c -- TEST INFRASTRUCTURE ONLY (part of liboracle.so, see xxm_oracle.h).
 *
 * Restatement of the localization statistics that consume the score-sorted candidate list:
 *   Motif::updateEnrichedRegion(sorted_sites*)   refinementPhase/Motif.cpp:968-1019
 *   MotifRegion::findMaximum / getRegion          nucleotides/motifRegion.cpp:13-188
 *   MotifRegion::calcRegionPval                   nucleotides/motifRegion.h:39-42
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "xxm_oracle.h"

/* findMaximum, motifRegion.cpp:13-65 (window size 5) */
static int region_find_maximum(const int* startPosCounts, int startRange, int endRange) {
	enum { WS = 5 };
	int positions[WS + 1];
	int max = 0;
	double maxOccurrence = 0;
	for (int i = 0; i <= WS; i++) positions[i] = -1;
	const int halfWS = WS / 2;
	for (int j = startRange; j < endRange; j++) {
		if (positions[WS - 1] == -1) {
			int pos = 0;
			while (positions[pos] != -1) pos++;
			positions[pos] = j;
		} else {
			positions[WS] = j;
			double occ = 0.0;
			const int current = positions[halfWS];
			for (int i = 0; i < halfWS; i++) occ += startPosCounts[positions[i]] / (current - positions[i] + 1);
			occ += startPosCounts[current];
			for (int i = halfWS + 1; i <= WS; i++) occ += startPosCounts[positions[i]] / (positions[i] - current + 1);
			if (occ > maxOccurrence) { max = current; maxOccurrence = occ; }
			for (int i = 0; i < WS; i++) positions[i] = positions[i + 1];
		}
	}
	if (positions[WS] == -1) {
		int maxNb = 0;
		for (int i = 0; i < WS; i++) {
			if (positions[i] == -1) break;
			else if (startPosCounts[positions[i]] > maxNb) { max = positions[i]; maxNb = startPosCounts[positions[i]]; }
		}
	}
	return max;
}

/* getRegion, motifRegion.cpp:67-188 ; out = {set, max, startRegion, endRegion} */
void xo_motif_region(const xo_ctx* c, const int32_t* startPosCounts, int32_t counts, int32_t startRange, int32_t endRange, int32_t* out) {
	const double enrichmentThreshold = 0.1;
	const double significanceThreshold = -log(1000);
	const double LOG_Bonferonni = log(endRange) + log(endRange - 1) - log(2);
	const int max = region_find_maximum(startPosCounts, startRange, endRange);
	const int L = endRange - startRange + 1;
	double newProb, oldBestProb = 0;
	int x1 = max, x2 = max;
	int nRound = 0;
	for (int i = x1; i <= x2; i++) nRound += startPosCounts[i];
	int occ = nRound;
#define REGION_PVAL(a, b, x) xo_order_statistics_pvalue(c, (x), counts, ((b) - (a) + 1.0) / L)
	double bestProb = REGION_PVAL(x1, x2, occ);
	if (occ * 1.0 / counts > enrichmentThreshold) oldBestProb = bestProb;
	else oldBestProb = 100;
	int x1Prime = x1;
	int loop, x1Test, x2Test;
	do {
		loop = 0;
		bestProb = oldBestProb;
		for (x1Test = x1 - 1; x1Test >= startRange; x1Test--) {
			if (startPosCounts[x1Test] == 0) continue;
			occ += startPosCounts[x1Test];
			newProb = REGION_PVAL(x1Test, x2, occ);
			if (newProb < bestProb) { bestProb = newProb; x1Prime = x1Test; }
		}
		const double bestProbLeft = bestProb;
		int x2Prime = x2;
		occ = nRound;
		bestProb = oldBestProb;
		for (x2Test = x2 + 1; x2Test <= endRange; x2Test++) {
			if (startPosCounts[x2Test] == 0) continue;
			occ += startPosCounts[x2Test];
			newProb = REGION_PVAL(x1, x2Test, occ);
			if (newProb < bestProb) { bestProb = newProb; x2Prime = x2Test; }
		}
		occ = nRound;
		const double bestProbRight = bestProb;
		if ((bestProbLeft < bestProbRight ? bestProbLeft : bestProbRight) < oldBestProb) {
			if (oldBestProb == 100) {
				if (bestProbLeft < bestProbRight) x1 = x1Prime;
				else x2 = x2Prime;
			} else {
				x1 = x1Prime;
				x2 = x2Prime;
			}
			nRound = 0;
			for (int i = x1; i <= x2; i++) nRound += startPosCounts[i];
			occ = nRound;
			bestProb = REGION_PVAL(x1, x2, occ);
			if (occ * 1.0 / counts > enrichmentThreshold) oldBestProb = bestProb;
			else oldBestProb = 100;
			loop = 1;
		}
	} while (loop);
#undef REGION_PVAL
	bestProb += LOG_Bonferonni;
	if (occ * 1.0 / counts > enrichmentThreshold && bestProb < significanceThreshold) {
		out[0] = 1; out[1] = max; out[2] = x1; out[3] = x2;
	} else {
		out[0] = 0; out[1] = startRange; out[2] = startRange; out[3] = endRange;
	}
}

/* Motif::updateEnrichedRegion(bestStartPos), Motif.cpp:968-1019, for --localization; `best` is the score-sorted candidate
 * list terminated by an entry with score 1; enr = {max, startRegion, endRegion} relative to PWM column 1 (in/out) */
void xo_update_enriched_region(const xo_ctx* c, const xo_site* best, int32_t revcomp, int32_t max_leng, int32_t firstMotifColumn, int32_t motifLength,
                               int32_t* enr) {
	int* countStartPos = (int*) calloc((size_t) max_leng + 2, sizeof(int));
	int size = 0;
	const int startRegion = enr[1] + firstMotifColumn - 1;
	const int endRegion = enr[2] + firstMotifColumn - 1;
	const int region_ext = (int) round(0.5 * (endRegion - startRegion + 1));
	int i = 0;
	while (best[i].score < 1e-3 || (size < 10 && best[i].score != 1)) {
		int startPos = best[i].startPos;
		if (revcomp && startPos > max_leng / 2) startPos = max_leng - (startPos + motifLength - 1);
		if (startPos >= startRegion - region_ext && startPos <= endRegion + region_ext) {
			countStartPos[startPos]++;
			size++;
		}
		i++;
	}
	int maxRegion = max_leng - motifLength + 1;
	if (revcomp) maxRegion = (max_leng - 1) / 2 - motifLength + 1;
	int32_t r[4];
	xo_motif_region(c, countStartPos, size, 1, maxRegion, r);
	enr[0] = r[1] - (firstMotifColumn - 1);
	enr[1] = r[2] - (firstMotifColumn - 1);
	enr[2] = r[3] - (firstMotifColumn - 1);
	free(countStartPos);
}
