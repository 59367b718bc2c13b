/*
 * xxmotif_gpu_binding.cpp -- definitions of the explicit specialisations declared in xxmotif_gpu_binding.h: the reference's
 * four window loops (StartPosUpdater-inl.h) replaced by calls into libxxmotif_gpu.so.  Compiled with -fno-access-control: the
 * results go where the reference's own code expects them, the private members of the StartPosUpdater singleton
 * (sortedSites, sortedSites_elong, bestStartPos, nbSites, nbBestSites, maxSites, log_posSetSize).
 *
 * One multi-device context (xxm_multi: this single host thread drives all of them): device 0 by default, XXM_DEVICE=d picks
 * another one, XXM_DEVICES=N shards the set over devices 0..N-1.  The sequence set, the options and the background model are
 * handed over at the first call, from the reference's own globals.
 */
#include "xxmotif_gpu_binding.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../include/xxmotif_gpu.h"
#include "LogTable.h"
#include "NullModel.h"

namespace {

xxm_multi* g_ctx = NULL;

void die(const char* what) { xxm_binding_die(what); }
#define XXM_OK(call) do { if (call) die(#call); } while (0)

xxm_multi* context() { return xxm_binding_context(); }

}  // namespace

void xxm_binding_die(const char* what) {
	fprintf(stderr, "libxxmotif_gpu: %s: %s\n", what, xxm_multi_last_error(g_ctx));
	exit(-1);  /* the reference's own error convention */
}

xxm_multi* xxm_binding_context() {
	if (g_ctx) return g_ctx;
	const char* dev = getenv("XXM_DEVICE");
	const char* ndev = getenv("XXM_DEVICES");
	const int32_t one = dev ? atoi(dev) : 0;
	if (ndev && atoi(ndev) > 1 ? xxm_multi_create(atoi(ndev), &g_ctx) : xxm_multi_create_on(&one, 1, &g_ctx)) die("xxm_multi_create");
	xxm_options o;
	o.mops = Global::multipleOccurrence; o.oops = Global::oneOccurrence; o.localization = Global::usePositionalProbs;
	o.maxMotifsPerSequence = Global::maxMotifsPerSequence; o.maxSeqCount = Global::maxSeqCount;
	o.plusFrac = Global::plusFrac; o.overrepCorrection = Global::overrepCorrection;
	const double* bg = Global::negSet != NULL ? Global::negBg : Global::posBg;
	const double* bgl = Global::negSet != NULL ? Global::negBg_log : Global::posBg_log;
	for (int a = 0; a < 5; a++) { o.bg[a] = bg[a]; o.bgLog[a] = bgl[a]; }
	XXM_OK(xxm_multi_set_options(g_ctx, &o));
	XXM_OK(xxm_multi_set_background(g_ctx, NullModel::_order, NullModel::_rprobs, NullModel::_rconds));
	/* Global::posSet as Alignment::getEntity left it; with --revcomp createRevcomp has doubled every sequence already
	   (seqset.cpp:128-178): the library wants the forward strands and derives the rest itself */
	ss_type set = Global::posSet;
	std::vector<int32_t> n(set->nent);
	std::vector<uint8_t> codes;
	for (int i = 1; i <= set->nent; i++) {
		const int stored = set->entity[i]->n;
		n[i - 1] = Global::revcomp ? (stored - 1) / 2 : stored;
		for (int j = 1; j <= n[i - 1]; j++) codes.push_back(set->entity[i]->S[0][j]);
	}
	XXM_OK(xxm_multi_load_seqset(g_ctx, set->nent, n.data(), codes.data(), Global::revcomp ? 1 : 0));
	return g_ctx;
}

namespace {

void reinitialize(xxm_multi* ctx, double** pwm, const motif_columns_type& columns, int posSetSize) {
	double flat[31 * 5];
	for (int a = 0; a < 5; a++) flat[a] = 0;
	for (int i = 1; i <= PWM_LENGTH; i++)
		for (int a = 0; a < 5; a++) flat[i * 5 + a] = pwm[i][a];
	std::vector<int32_t> cols(columns.begin(), columns.end());
	XXM_OK(xxm_multi_reinitialize(ctx, flat, cols.data(), (int32_t) cols.size(), posSetSize, Global::posSet->total[posSetSize]));
}

/* a device-resident list into one of the reference's arrays of site pointers */
int fetch(xxm_multi* ctx, int which, int order, sorted_sites* dst, int cap) {
	static std::vector<xxm_site> buf;
	if ((int) buf.size() < cap) buf.resize(cap);
	int32_t n = 0;
	XXM_OK(xxm_multi_get_sites(ctx, which, order, 0, buf.data(), cap, &n));
	for (int i = 0; i < n; i++) { dst[i]->score = buf[i].score; dst[i]->startPos = buf[i].startPos; dst[i]->sequence = buf[i].sequence; }
	return n;
}

}  // namespace

/* StartPosUpdater-inl.h:70-140 */
template <>
void StartPosUpdater::fill_sites_with_possible_startPositions<XxmDnaKmerGenerator>(XxmDnaKmerGenerator&, bool keepPosSetSize, Motif* motif) {
	xxm_multi* ctx = context();
	const int posSetSize = motif->getPosSetSize();
	(void) keepPosSetSize;  /* sequences beyond posSetSize are scanned and then dropped again by the reference (:114-117) */
	reinitialize(ctx, motif->getPWM(), motif->getMotifColumns(), posSetSize);
	int32_t nb = 0, nbest = 0, mx = 0;
	XXM_OK(xxm_multi_scan_full(ctx, posSetSize, motif->getEnrichment().startRegion, motif->getEnrichment().endRegion, &nb, &nbest, &mx));
	if (nbest + 1 > sitesAlloc2 || nb + 1 > sitesAlloc1) { fprintf(stderr, "libxxmotif_gpu: more sites than the reference allocates\n"); exit(-1); }
	nbBestSites = fetch(ctx, XXM_LIST_BEST_STARTPOS, XXM_ORDER_BY_SCORE, bestStartPos, nbest + 1);  /* sortSitesByScore(bestStartPos, nbBestSites), :128 */
	bestStartPos[nbBestSites]->score = 1;
	nbSites = fetch(ctx, XXM_LIST_SORTED_SITES, XXM_ORDER_AS_IS, sortedSites, nb + 1);
	sortedSites[nbSites]->score = 1;
	maxSites = mx;
	log_posSetSize = LogTable::LOG_i[Global::posSet->total[posSetSize]];
}

/* StartPosUpdater-inl.h:143-221 */
template <>
void StartPosUpdater::fill_elongSites_with_possible_startPositions<XxmDnaKmerGenerator>(XxmDnaKmerGenerator&, Motif* motif, double** pwm,
                                                                                        motif_columns_type& tmpColumns, bool palin) {
	xxm_multi* ctx = context();
	const int posSetSize = motif->getPosSetSize();
	sortSitesByStartPos(bestStartPos, nbBestSites, palin);  /* the reference re-orders its candidate array here (:164); keep the host copy in step */
	reinitialize(ctx, pwm, tmpColumns, posSetSize);
	int32_t nb = 0, mx = 0;
	XXM_OK(xxm_multi_rescore(ctx, motif->getFirstMotifColumn(), posSetSize, motif->getEnrichment().startRegion, motif->getEnrichment().endRegion, palin ? 1 : 0, &nb,
	                   &mx));
	if (Global::multipleOccurrence) {
		nbSites = fetch(ctx, XXM_LIST_SORTED_SITES_ELONG, XXM_ORDER_AS_IS, sortedSites_elong, nb + 1);
		sortedSites_elong[nbSites]->score = 1;
	} else {
		fetch(ctx, XXM_LIST_SORTED_SITES_ELONG, XXM_ORDER_AS_IS, sortedSites_elong, posSetSize);
		nbSites = posSetSize;
	}
	maxSites = mx;
	log_posSetSize = LogTable::LOG_i[Global::posSet->total[posSetSize]];
}

/* StartPosUpdater-inl.h:224-266 */
template <>
void StartPosUpdater::fill_sites_with_possible_startPositions_oops<XxmDnaKmerGenerator>(XxmDnaKmerGenerator&, Motif* motif) {
	xxm_multi* ctx = context();
	const int posSetSize = motif->getPosSetSize();
	reinitialize(ctx, motif->getPWM(), motif->getMotifColumns(), posSetSize);
	int32_t nb = 0, mx = 0;
	XXM_OK(xxm_multi_scan_oops(ctx, posSetSize, motif->getEnrichment().startRegion, motif->getEnrichment().endRegion, &nb, &mx));
	fetch(ctx, XXM_LIST_SORTED_SITES, XXM_ORDER_AS_IS, sortedSites, Global::posSet->nent);
	nbSites = posSetSize;
	maxSites = mx;
}

/* StartPosUpdater-inl.h:6-68 */
template <>
void StartPosUpdater::fill_sites_with_startPosList<XxmDnaKmerGenerator>(XxmDnaKmerGenerator&, Motif* motif, bool instanceThreshold) {
	xxm_multi* ctx = context();
	const int posSetSize = motif->getPosSetSize();
	reinitialize(ctx, motif->getPWM(), motif->getMotifColumns(), posSetSize);
	std::vector<int32_t> seq, pos;
	for (StartPosContainer::const_iterator it = motif->getStartPosList().begin(); it != motif->getStartPosList().end(); it++) {
		seq.push_back(it->seq);
		pos.push_back(it->pos);
	}
	int32_t nb = 0, mx = 0;
	XXM_OK(xxm_multi_rescore_startposlist(ctx, posSetSize, motif->getEnrichment().startRegion, motif->getEnrichment().endRegion, (int32_t) seq.size(), seq.data(),
	                                pos.data(), instanceThreshold ? 1 : 0, &nb, &mx));
	if (Global::multipleOccurrence) {
		nbSites = fetch(ctx, XXM_LIST_SORTED_SITES, XXM_ORDER_AS_IS, sortedSites, nb + 1);
	} else {
		fetch(ctx, XXM_LIST_SORTED_SITES, XXM_ORDER_AS_IS, sortedSites, posSetSize);
		nbSites = posSetSize;
	}
	maxSites = mx;
	log_posSetSize = LogTable::LOG_i[Global::posSet->total[posSetSize]];
}
