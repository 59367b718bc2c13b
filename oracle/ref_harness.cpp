/*
 * ref_harness -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Drives the UNMODIFIED reference (soedinglab/xxmotif, compiled in place from
 * /root/reference/src by oracle/Makefile) through its own public call surface
 * for the PWM-refinement hot path and dumps every boundary array into a small
 * tagged binary "trace" file.  The traces are the golden vectors that pin
 * (a) the C restatement in oracle/xxm_oracle.c and (b) the CUDA path behind
 * include/xxmotif_gpu.h.  The reference has no tests or fixtures of its own
 * (SURVEY.md section 4), so these dumps of the real reference functions are
 * the only possible pin.
 *
 * No reference source is copied: this file only *calls* reference functions
 *   Global::Global                       src/Globals.cpp:433
 *   StartModels::initStartMotif          src/StartModels.cpp:125
 *   Kmer_Generator_Reloaded::reinitialize src/branch_and_bound-inl.h:183
 *   StartPosUpdater::fill_sites_with_possible_startPositions      src/refinementPhase/StartPosUpdater-inl.h:70
 *   StartPosUpdater::fill_elongSites_with_possible_startPositions src/refinementPhase/StartPosUpdater-inl.h:143
 *   StartPosUpdater::fill_sites_with_possible_startPositions_oops src/refinementPhase/StartPosUpdater-inl.h:224
 *   StartPosUpdater::fill_sites_with_startPosList                 src/refinementPhase/StartPosUpdater-inl.h:6
 *   StartPosUpdater::update_PWM_with_bestSubset / update_elongPWM_with_bestSubset  src/refinementPhase/StartPosUpdater.cpp:78,96
 *   PVal_Calculator::calculatePval       src/pValCalculation.h:106
 *   IterateMotif::maximizeMotifLength    src/refinementPhase/Iterate_Motif.cpp:40
 *   MotifContainer::iterate              src/refinementPhase/MotifContainer.cpp:216
 * and reads their (private) result buffers; it is compiled with
 * -fno-access-control for that purpose.
 *
 *   ElongCandidates::countAllOccurrences src/elongationPhase/elongationCandidates.cpp:234 (mode "gapped")
 * usage:
 *   ref_harness gapped five|pal|tandem OUTFILE OUTDIR SEQFILE [reference flags]   -> counts of every gap pattern (binary)
 *   ref_harness trace TRACEFILE OUTDIR SEQFILE [reference flags, must include -m or -p]
 *   ref_harness bench STEPS WARMUP OUTDIR SEQFILE [reference flags]   -> one JSON line on stdout
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>
#include <unistd.h>

#include "backgroundDistribution.h"
#include "Globals.h"
#include "LogTable.h"
#include "NullModel.h"
#include "StartModels.h"
#include "utils.h"
#include "refinementPhase/MotifContainer.h"
#include "refinementPhase/Refine_Motifs.h"
#include "refinementPhase/Iterate_Motif.h"
#include "elongationPhase/elongationCandidates.h"

typedef Kmer_Generator_Reloaded<dna_hash_t> KGen;

/* ---------------------------------------------------------------- trace io */
static FILE* g_tr = NULL;

static void tr_raw(const std::string& name, char dtype, size_t esize, size_t n, const void* data) {
	uint32_t len = (uint32_t) name.size();
	uint64_t cnt = (uint64_t) n;
	fwrite(&len, 4, 1, g_tr);
	fwrite(name.data(), 1, len, g_tr);
	fwrite(&dtype, 1, 1, g_tr);
	fwrite(&cnt, 8, 1, g_tr);
	if (n) fwrite(data, esize, n, g_tr);
}
static void tr_i32(const std::string& n, const std::vector<int32_t>& v) { tr_raw(n, 'i', 4, v.size(), v.data()); }
static void tr_i64(const std::string& n, const std::vector<int64_t>& v) { tr_raw(n, 'l', 8, v.size(), v.data()); }
static void tr_f32(const std::string& n, const std::vector<float>& v) { tr_raw(n, 'f', 4, v.size(), v.data()); }
static void tr_f64(const std::string& n, const std::vector<double>& v) { tr_raw(n, 'd', 8, v.size(), v.data()); }
static void tr_u8(const std::string& n, const std::vector<uint8_t>& v) { tr_raw(n, 'b', 1, v.size(), v.data()); }
static void tr_int(const std::string& n, int64_t x) { std::vector<int64_t> v(1, x); tr_i64(n, v); }
static void tr_dbl(const std::string& n, double x) { std::vector<double> v(1, x); tr_f64(n, v); }

static void tr_sites(const std::string& p, sorted_sites* s, int n) {
	std::vector<int32_t> seq(n), pos(n);
	std::vector<float> sc(n);
	for (int i = 0; i < n; i++) { seq[i] = s[i]->sequence; pos[i] = s[i]->startPos; sc[i] = s[i]->score; }
	tr_i32(p + ".seq", seq); tr_i32(p + ".pos", pos); tr_f32(p + ".score", sc);
}
static void tr_pwm(const std::string& n, double** pwm) {
	std::vector<double> v;
	for (int i = 0; i <= PWM_LENGTH; i++) for (int a = 0; a <= 4; a++) v.push_back(i == 0 ? 0.0 : pwm[i][a]);
	tr_f64(n, v);
}
static void tr_cols(const std::string& n, const motif_columns_type& c) {
	std::vector<int32_t> v(c.begin(), c.end());
	tr_i32(n, v);
}
static void tr_motif(const std::string& p, Motif* m) {
	tr_pwm(p + ".pwm", m->getPWM());
	tr_cols(p + ".cols", m->getMotifColumns());
	tr_int(p + ".posSetSize", m->getPosSetSize());
	tr_dbl(p + ".pval", m->getPval());
	tr_dbl(p + ".posPval", m->getPosPval());
	region r = m->getEnrichment();
	std::vector<int32_t> e; e.push_back(r.max); e.push_back(r.startRegion); e.push_back(r.endRegion);
	tr_i32(p + ".enrich", e);
	std::vector<int32_t> sseq, spos;
	for (StartPosContainer::const_iterator it = m->getStartPosList().begin(); it != m->getStartPosList().end(); ++it) {
		sseq.push_back(it->seq); spos.push_back(it->pos);
	}
	tr_i32(p + ".sites.seq", sseq); tr_i32(p + ".sites.pos", spos);
}

static void tr_seqset(const std::string& p, ss_type set) {
	std::vector<int32_t> n(set->nent), total(set->nent + 1);
	std::vector<double> avg(set->nent + 1);
	std::vector<uint8_t> S;
	for (int i = 1; i <= set->nent; i++) {
		n[i - 1] = set->entity[i]->n;
		for (int j = 1; j <= set->entity[i]->n; j++) S.push_back(set->entity[i]->S[0][j]);
	}
	for (int i = 0; i <= set->nent; i++) { total[i] = set->total[i]; avg[i] = (i == 0) ? 0.0 : set->avgLength[i]; }
	tr_int(p + ".nent", set->nent);
	tr_int(p + ".max_leng", set->max_leng);
	tr_int(p + ".min_leng", set->min_leng);
	tr_i32(p + ".n", n); tr_u8(p + ".S", S); tr_i32(p + ".total", total); tr_f64(p + ".avgLength", avg);
}

/* dump the p-value tables built by the last reinitialize() (private state of the singleton) */
static void tr_tables(const std::string& p, KGen& kg) {
	tr_int(p + ".L", kg.totalLength);
	tr_int(p + ".maxk", kg.MAX_KMER_SIZE);
	tr_int(p + ".hashNumber", kg.hashNumber);
	std::vector<int32_t> gaps(kg.gaps.begin(), kg.gaps.end());
	tr_i32(p + ".gaps", gaps);
	const int nparts = kg.hashNumber;
	for (int part = 0; part < nparts; part++) {
		char b[64]; snprintf(b, sizeof b, "%s.p%d", p.c_str(), part);
		std::vector<int64_t> ids; std::vector<double> vals;
		const size_t sz = kg.pValueHash[part].size();
		for (size_t id = 0; id < sz; id++) {
			if (kg.pValueHash[part][id] != 0) { ids.push_back((int64_t) id); vals.push_back(kg.pValueHash[part][id]); }
		}
		tr_i64(std::string(b) + ".ids", ids); tr_f64(std::string(b) + ".vals", vals);
		std::vector<int32_t> cscore, ccount; std::vector<int64_t> cids;
		for (size_t c = 0; c < kg.scoreIdVector[part].size(); c++) {
			cscore.push_back(kg.scoreIdVector[part][c].score);
			ccount.push_back((int32_t) kg.scoreIdVector[part][c].id_list.size());
			for (KGen::id_list_type::const_iterator it = kg.scoreIdVector[part][c].id_list.begin();
			     it != kg.scoreIdVector[part][c].id_list.end(); ++it) cids.push_back((int64_t) *it);
		}
		snprintf(b, sizeof b, "%s.c%d", p.c_str(), part);
		tr_i32(std::string(b) + ".score", cscore); tr_i32(std::string(b) + ".count", ccount); tr_i64(std::string(b) + ".ids", cids);
	}
	if (nparts == 2) {
		std::vector<int64_t> ids; std::vector<double> vals;
		const size_t sz = kg.probabilityHash.size();
		for (size_t id = 0; id < sz; id++) {
			if (kg.probabilityHash[id] != 0) { ids.push_back((int64_t) id); vals.push_back(kg.probabilityHash[id]); }
		}
		tr_i64(p + ".prob.ids", ids); tr_f64(p + ".prob.vals", vals);
	}
}

/* per-window doubles for all windows of the first `nseq` sequences (pins the >8-column slow path in full precision) */
static void tr_lookups(const std::string& p, KGen& kg, const motif_columns_type& cols, Motif* m, int nseq, bool mops) {
	PVal_Calculator& pc = PVal_Calculator::getInstance();
	static unsigned char kmer[64];
	const int length = cols.back() - cols.front() + 1;
	const int region = m->getEnrichment().endRegion - m->getEnrichment().startRegion + 1;
	/* same positional prior as the scan that follows would set up (StartPosUpdater-inl.h:84-86, :155-157) */
	if (Global::usePositionalProbs)
		pc.initPositionalProbs(length, m->getFirstMotifColumn(), m->getEnrichment(), cols.front() - m->getFirstMotifColumn());
	std::vector<int32_t> sq, ps; std::vector<double> pv, fp;
	for (int i = 1; i <= std::min(nseq, Global::posSet->nent); i++) {
		const int len = Global::posSet->entity[i]->n - length + 1;
		const int delta_0 = std::min(len, region);
		for (int j = 1; j <= len; j++) {
			double pValScore = 0;
			double f = pc.calculatePval(kg, kmer, cols, i, j, len, delta_0, length, length,
			                            Global::posSet->avgLength[m->getPosSetSize()], pValScore, mops);
			sq.push_back(i); ps.push_back(j); pv.push_back(pValScore); fp.push_back(f);
		}
	}
	tr_i32(p + ".seq", sq); tr_i32(p + ".pos", ps); tr_f64(p + ".pValScore", pv); tr_f64(p + ".finalPval", fp);
}

static double now_s() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec; }

static int run_bench(int steps, int warmup, Motif* m) {
	KGen& kg = KGen::getInstance();
	StartPosUpdater& sp = StartPosUpdater::getInstance();
	const int length0 = m->getMotifLength();
	long long windows = 0;
	for (int i = 1; i <= m->getPosSetSize(); i++) windows += std::max(0, Global::posSet->entity[i]->n - length0 + 1);
	int bestMotifNb = 0; double bestPval = 0;
	double t0 = 0;
	std::vector<double> pwm0;  /* the start motif as initStartMotif left it: what a parity test feeds the other implementation */
	for (int i = 0; i <= PWM_LENGTH; i++)
		for (int a = 0; a <= 4; a++) pwm0.push_back(i == 0 ? 0.0 : m->getPWM()[i][a]);
	const region enr0 = m->getEnrichment();
	for (int it = 0; it < warmup + steps; it++) {
		if (it == warmup) t0 = now_s();
		kg.reinitialize(m->getPWM(), m->getMotifColumns(), m->getPosSetSize(), Global::posSet->total[m->getPosSetSize()]);
		sp.fill_sites_with_possible_startPositions(kg, true, m);
		sp.update_PWM_with_bestSubset(m, bestMotifNb, bestPval, false, Global::pseudo);
		/* with --localization the per-motif loop re-estimates the enriched region from the candidate list after every
		   optimizeMotif (MotifContainer.cpp:241); doing it after every step gives the next step a real positional prior */
		if (Global::usePositionalProbs) m->updateEnrichedRegion(sp.getBestStartPos());
	}
	double dt = now_s() - t0;
	printf("{\"impl\": \"reference\", \"steps\": %d, \"warmup\": %d, \"windows_per_step\": %lld, \"seconds\": %.6f, "
	       "\"windows_per_s\": %.6e, \"nbSites\": %d, \"nbBestSites\": %d, \"maxSites\": %d, \"bestMotifNb\": %d, \"bestPval\": %.17g",
	       steps, warmup, windows, dt, (double) windows * steps / dt, sp.nbSites, sp.nbBestSites, sp.maxSites, bestMotifNb, bestPval);
	/* the state after the last step, so that a parity test can compare more than scalars: the PWM the steps ended on */
	printf(", \"nent\": %d, \"posSetSize\": %d, \"total\": %d, \"revcomp\": %d, \"mops\": %d, \"localization\": %d, \"order\": %d, \"enr\": [%d, %d, %d], \"cols\": [",
	       Global::posSet->nent, m->getPosSetSize(), Global::posSet->total[m->getPosSetSize()], (int) Global::revcomp, (int) Global::multipleOccurrence,
	       (int) Global::usePositionalProbs, Global::order, m->getEnrichment().max, m->getEnrichment().startRegion, m->getEnrichment().endRegion);
	{
		const motif_columns_type& c = m->getMotifColumns();
		bool first = true;
		for (motif_columns_type::const_iterator it = c.begin(); it != c.end(); ++it) { printf("%s%d", first ? "" : ", ", *it); first = false; }
	}
	printf("], \"pwm\": [");
	for (int i = 0; i <= PWM_LENGTH; i++)
		for (int a = 0; a <= 4; a++) printf("%s%.17g", (i || a) ? ", " : "", i == 0 ? 0.0 : m->getPWM()[i][a]);
	printf("], \"enr0\": [%d, %d, %d], \"pwm0\": [", enr0.max, enr0.startRegion, enr0.endRegion);
	for (size_t i = 0; i < pwm0.size(); i++) printf("%s%.17g", i ? ", " : "", pwm0[i]);
	printf("]}\n");
	fflush(stdout);
	return 0;
}

/* ElongCandidates::countAllOccurrences for every gap pattern of one seed type.  File: int32 ncoding,
 * int32 npatterns, int32 firstHalfOnly, double seconds; per pattern int32 offset[ncoding] (l + gap[l-1]) and uint32 counts[4^ncoding]
 * (first character most significant), read from countIUPAC at the ids without IUPAC characters. */
static int run_gapped(const char* typeName, const char* outFile) {
	motif_type type = strcmp(typeName, "pal") == 0 ? PALINDROME : (strcmp(typeName, "tandem") == 0 ? TANDEM : FIVEMERS);
	FILE* f = fopen(outFile, "wb");
	if (!f) { perror("gapped"); return 1; }
	int npat = 0;
	double secs = 0;
	std::vector<int32_t> body;
	int ncoding = 0;
	const int have = type == FIVEMERS ? (int) (sizeof(ElongCandidates::gaps) / sizeof(ElongCandidates::gaps[0])) : 16;
	for (int k = 0; k < have; k++) {
		{
			ElongCandidates ec(k, type, 1e-3);
			ncoding = ElongCandidates::coding_characters;
			const uint8_t* gap = type == FIVEMERS ? ElongCandidates::gaps[k] : ElongCandidates::gaps_palindrome[k];
			const double t0 = now_s();
			ec.countAllOccurrences();
			secs += now_s() - t0;
			body.push_back(0);
			for (int l = 1; l < ncoding; l++) body.push_back(l + gap[l - 1]);
			for (uint32_t id = 0; id < ElongCandidates::POW_4_coding; id++) {
				uint32_t id16 = 0;
				for (int l = 0; l < ncoding; l++) id16 = (id16 << 4) + ((id >> (2 * (ncoding - 1 - l))) & 3u);
				body.push_back((int32_t) ElongCandidates::countIUPAC[id16]);
			}
			npat++;
		}
	}
	int32_t head[3] = {ncoding, npat, (type == PALINDROME && Global::revcomp) ? 1 : 0};
	fwrite(head, 4, 3, f);
	fwrite(&secs, 8, 1, f);
	fwrite(body.data(), 4, body.size(), f);
	fclose(f);
	return 0;
}

int main(int argc, char** argv) {
	if (argc >= 6 && strcmp(argv[1], "gapped") == 0) {
		int saved = dup(1);
		fflush(stdout);
		dup2(2, 1);
		std::vector<char*> av; av.push_back(argv[0]);
		for (int i = 4; i < argc; i++) av.push_back(argv[i]);
		Global G((int) av.size(), av.data());
		LogTable L;
		const int rc = run_gapped(argv[2], argv[3]);
		fflush(stdout); dup2(saved, 1);
		return rc;
	}
	if (argc < 5) { fprintf(stderr, "usage: ref_harness trace TRACE OUTDIR SEQ [flags] | bench STEPS WARMUP OUTDIR SEQ [flags] | refine OUTDIR SEQ [flags]\n"); return 2; }
	const bool bench = strcmp(argv[1], "bench") == 0;
	const bool refine = strcmp(argv[1], "refine") == 0;  /* refine OUTDIR SEQ [flags]: MotifContainer::iterate of the start motif, timed */
	int shift = bench ? 3 : (refine ? 1 : 2);
	int steps = 0, warmup = 0;
	std::string trace_path;
	if (bench) { steps = atoi(argv[2]); warmup = atoi(argv[3]); } else if (!refine) trace_path = argv[2];

	/* the reference prints its banner/progress on stdout: keep stdout clean for the JSON line */
	int saved_stdout = dup(1);
	fflush(stdout);
	dup2(2, 1);

	std::vector<char*> av; av.push_back(argv[0]);
	for (int i = shift + 1; i < argc; i++) av.push_back(argv[i]);
	Global G((int) av.size(), av.data());
	LogTable L;
	MotifContainer startModels;
	StartModels seeds;
	if (Global::profFile == NULL && Global::startMotif == NULL) { fprintf(stderr, "ref_harness needs -m or -p\n"); return 2; }
	seeds.initStartMotif(startModels);
	Motif* m = startModels.getMotifs().front();

	if (bench) {
		fflush(stdout); dup2(saved_stdout, 1);
		return run_bench(steps, warmup, m);
	}
	if (refine) {
		const double t0 = now_s();
		startModels.iterate(Global::maximizeMotifLength, false);
		const double dt = now_s() - t0;
		fflush(stdout); dup2(saved_stdout, 1);
		const motif_columns_type& c = m->getMotifColumns();
		printf("{\"impl\": \"reference\", \"mode\": \"refine\", \"seconds\": %.6f, \"nent\": %d, \"firstCol\": %d, \"lastCol\": %d, \"ncols\": %d, "
		       "\"nsites\": %d, \"pval\": %.17g}\n",
		       dt, Global::posSet->nent, c.front(), c.back(), (int) c.size(), (int) m->getStartPosList().size(), m->getPval());
		fflush(stdout);
		return 0;
	}

	g_tr = fopen(trace_path.c_str(), "wb");
	if (!g_tr) { perror("trace"); return 1; }
	fwrite("XXMTRC1\n", 1, 8, g_tr);

	/* ------------------------------------------------------------ inputs */
	tr_seqset("in.pos", Global::posSet);
	tr_int("in.hasNeg", Global::negSet != NULL);
	if (Global::negSet != NULL) tr_seqset("in.neg", Global::negSet);
	tr_int("cfg.revcomp", Global::revcomp);
	tr_int("cfg.mops", Global::multipleOccurrence);
	tr_int("cfg.oops", Global::oneOccurrence);
	tr_int("cfg.zoops", Global::zeroOrOneOccurrence);
	tr_int("cfg.localization", Global::usePositionalProbs);
	tr_int("cfg.order", Global::order);
	tr_dbl("cfg.pseudo", Global::pseudo);
	tr_dbl("cfg.overrepCorrection", Global::overrepCorrection);
	tr_int("cfg.maxMotifsPerSequence", Global::maxMotifsPerSequence);
	tr_int("cfg.maxSeqCount", Global::maxSeqCount);
	tr_dbl("cfg.plusFrac", Global::plusFrac);
	tr_dbl("cfg.neff_pwm", Global::neff_pwm);
	tr_int("cfg.minMatchPositions", Global::minMatchPositions);
	tr_int("cfg.maxMatchPositions", Global::maxMatchPositions);
	tr_dbl("cfg.pseudocountsFactor", Global::pseudocountsFactor);
	tr_dbl("cfg.countsOffset", Global::countsOffset);
	tr_int("cfg.startRegion", Global::startRegion);
	tr_int("cfg.endRegion", Global::endRegion);
	tr_int("cfg.maximizeMotifLength", Global::maximizeMotifLength);
	tr_int("cfg.fixedPosition", Global::fixedPosition);
	{
		std::vector<double> a(Global::posBg, Global::posBg + 5), b(Global::posBg_log, Global::posBg_log + 5);
		std::vector<double> c(Global::negBg, Global::negBg + 5), d(Global::negBg_log, Global::negBg_log + 5);
		tr_f64("bg.posBg", a); tr_f64("bg.posBg_log", b); tr_f64("bg.negBg", c); tr_f64("bg.negBg_log", d);
	}
	{
		tr_int("nm.order", NullModel::_order);
		tr_int("nm.rfields", NullModel::_rfields);
		const int rf = NullModel::_rfields;
		if (rf <= (1 << 18)) {
			std::vector<float> a(NullModel::_rprobs, NullModel::_rprobs + rf), b(NullModel::_rconds, NullModel::_rconds + rf);
			tr_f32("nm.rprobs", a); tr_f32("nm.rconds", b);
		} else { /* too big for a fixture: FNV-1a over the raw bytes + the non-gapped full-order slice is checked instead */
			uint64_t h1 = 1469598103934665603ULL, h2 = h1;
			const unsigned char* p1 = (const unsigned char*) NullModel::_rprobs; const unsigned char* p2 = (const unsigned char*) NullModel::_rconds;
			for (size_t i = 0; i < (size_t) rf * 4; i++) { h1 = (h1 ^ p1[i]) * 1099511628211ULL; h2 = (h2 ^ p2[i]) * 1099511628211ULL; }
			std::vector<int64_t> h; h.push_back((int64_t) h1); h.push_back((int64_t) h2);
			tr_i64("nm.hash", h);
		}
	}
	{
		std::vector<double> o;
		for (int i = 0; i < PWM_LENGTH; i++) o.push_back(PVal_Calculator::getInstance().getOverrepCorrection(i));
		tr_f64("pc.overrep", o);
	}
	tr_motif("m0", m);

	KGen& kg = KGen::getInstance();
	StartPosUpdater& sp = StartPosUpdater::getInstance();
	int bestMotifNb = 0; double bestPval = 0;

	/* ---------------------------------- A. the optimizeMotif body (Iterate_Motif.cpp:23-37), stage by stage */
	for (int iter = 0; iter < 2; iter++) {
		char pb[32]; snprintf(pb, sizeof pb, "opt%d", iter);
		std::string p(pb);
		tr_motif(p + ".in", m);
		kg.reinitialize(m->getPWM(), m->getMotifColumns(), m->getPosSetSize(), Global::posSet->total[m->getPosSetSize()]);
		tr_tables(p + ".tab", kg);
		tr_lookups(p + ".lk_mops", kg, m->getMotifColumns(), m, 4, true);
		tr_lookups(p + ".lk_mode", kg, m->getMotifColumns(), m, 4, Global::multipleOccurrence);
		sp.fill_sites_with_possible_startPositions(kg, true, m);
		tr_int(p + ".nbSites", sp.nbSites); tr_int(p + ".nbBestSites", sp.nbBestSites); tr_int(p + ".maxSites", sp.maxSites);
		tr_sites(p + ".best", sp.bestStartPos, sp.nbBestSites);
		tr_sites(p + ".sites", sp.sortedSites, sp.nbSites);
		bestMotifNb = 0; bestPval = 0;
		sp.update_PWM_with_bestSubset(m, bestMotifNb, bestPval, false, Global::pseudo);
		tr_sites(p + ".sorted", sp.sortedSites, sp.nbSites);
		tr_int(p + ".bestMotifNb", bestMotifNb); tr_dbl(p + ".bestPval", bestPval);
		tr_pwm(p + ".pwm_out", m->getPWM());
	}
	sp.setPositionalPval(m);
	tr_dbl("optA.posPval", m->getPosPval());
	if (Global::usePositionalProbs) {
		m->updateEnrichedRegion(sp.getBestStartPos());
		tr_motif("optA.enriched", m);
	}

	/* ---------------------------------- D. candidate re-scoring (StartPosUpdater-inl.h:143) for every accepted column variant */
	{
		IterateMotif im;
		double** elongPWM = im.elongPWM;
		int v = 0;
		for (int front = -2; front <= 2; front++) {
			for (int back = -2; back <= 2; back++) {
				motif_columns_type tmpColumns;
				if (im.fillTmpColumn_noGaps(front, back, tmpColumns, m)) continue;
				int motifNb = 0; double overrepPval = 1e100;
				for (int iter = 0; iter < 3; iter++) {
					char pb[48]; snprintf(pb, sizeof pb, "el%d.%d", v, iter);
					std::string p(pb);
					double** pwm = (iter == 0) ? m->getPWM() : elongPWM;
					tr_int(p + ".front", front); tr_int(p + ".back", back);
					tr_cols(p + ".cols", tmpColumns);
					tr_pwm(p + ".pwm_in", pwm);
					kg.reinitialize(pwm, tmpColumns, m->getPosSetSize(), Global::posSet->total[m->getPosSetSize()]);
					if (iter == 0) tr_tables(p + ".tab", kg);
					if (iter == 0) tr_lookups(p + ".lk_mode", kg, tmpColumns, m, 3, Global::multipleOccurrence);
					sp.fill_elongSites_with_possible_startPositions(kg, m, pwm, tmpColumns, false);
					tr_int(p + ".nbSites", sp.nbSites); tr_int(p + ".maxSites", sp.maxSites);
					tr_sites(p + ".elong", sp.sortedSites_elong, sp.nbSites);
					sp.update_elongPWM_with_bestSubset(m, elongPWM, tmpColumns, motifNb, overrepPval, iter == 0, Global::pseudo);
					tr_sites(p + ".sorted", sp.sortedSites_elong, sp.nbSites);
					tr_int(p + ".motifNb", motifNb); tr_dbl(p + ".overrepPval", overrepPval);
					tr_pwm(p + ".pwm_out", elongPWM);
					if (motifNb == 0) break;
				}
				v++;
			}
		}
		tr_int("el.count", v);
		/* ------------------------------ E. the reference's own maximizeMotifLength as a black box */
		im.maximizeMotifLength(m, Global::maximizeMotifLength, false);
		tr_motif("maxlen", m);
	}

	/* ---------------------------------- F. MotifContainer::iterate to convergence (black box) */
	startModels.iterate(Global::maximizeMotifLength, false);
	tr_motif("converged", m);

	/* ---------------------------------- J. Motif::updatePval body (Motif.cpp:509-531): re-score the motif's own site list */
	{
		kg.reinitialize(m->getPWM(), m->getMotifColumns(), m->getPosSetSize(), Global::posSet->total[m->getPosSetSize()]);
		tr_tables("upd.tab", kg);
		sp.fill_sites_with_startPosList(kg, m);
		tr_int("upd.nbSites", sp.nbSites); tr_int("upd.maxSites", sp.maxSites);
		tr_sites("upd.sites", sp.sortedSites, sp.nbSites);
		double bp = 0; int bn = 0;
		sp.fill_startPosList_with_bestSubset(m, bp, bn);
		tr_int("upd.bestMotifNb", bn); tr_dbl("upd.bestPval", bp);
		tr_motif("upd.after", m);
	}

	/* ---------------------------------- G. final OOPS scan (Motif.cpp:533-538) */
	{
		kg.reinitialize(m->getPWM(), m->getMotifColumns(), m->getPosSetSize(), Global::posSet->total[m->getPosSetSize()]);
		sp.fill_sites_with_possible_startPositions_oops(kg, m);
		tr_int("oops.nbSites", sp.nbSites); tr_int("oops.maxSites", sp.maxSites);
		tr_sites("oops.sites", sp.sortedSites, Global::posSet->nent);
		sp.set_BindingSites(m, Global::instanceThreshold);
		std::vector<int32_t> bs, bp; std::vector<double> bsc;
		for (MatchContainer::const_iterator it = m->getBindingSites().begin(); it != m->getBindingSites().end(); ++it) {
			bs.push_back(it->seq); bp.push_back(it->pos); bsc.push_back(it->score);
		}
		tr_i32("oops.bind.seq", bs); tr_i32("oops.bind.pos", bp); tr_f64("oops.bind.score", bsc);
	}

	/* ---------------------------------- H. per-site p-values for _Pvals.txt (StartPosUpdater.h:91-132) */
	{
		list<double> siteProb, siteCons; double comb = 0;
		sp.getProbabilitiesPerSite(siteProb, siteCons, comb, m);
		std::vector<double> v(siteProb.begin(), siteProb.end());
		tr_f64("persite.prob", v);
	}

	/* ---------------------------------- I. forward / reverse-complement log-odds histogram (output.cpp:1033-1078) */
	{
		double maxScore = m->getMaxMatrixScore();
		Motif* m2 = m->getPalindrome();
		double maxScore2 = m2->getMaxMatrixScore();
		std::vector<double> fw, bw;
		for (int pos = 1; pos <= Global::posSet->max_leng - m->getUsedLength() + 1; pos++) {
			double cf = 0, cb = 0;
			for (int seq = 1; seq <= m->getPosSetSize(); seq++) {
				if (Global::posSet->entity[seq]->n - m->getMotifLength() + 1 < pos) continue;
				if (m->calculateMatrixScore_logOdds(seq, pos) > maxScore * 0.5) cf++;
				if (m2->calculateMatrixScore_logOdds(seq, pos) > maxScore2 * 0.5) cb++;
			}
			fw.push_back(cf); bw.push_back(cb);
		}
		delete m2;
		tr_f64("hist.fwd", fw); tr_f64("hist.bwd", bw);
		tr_dbl("hist.maxScore", maxScore); tr_dbl("hist.maxScore2", maxScore2);
	}
	tr_int("end", 1);
	fclose(g_tr);
	fflush(stdout);
	return 0;
}
