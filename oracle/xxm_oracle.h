/*
 * xxm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU restatement (plain C) of the reference's PWM-refinement hot path
 * (soedinglab/xxmotif; citations are file:line under /root/reference/src).
 * It exists to CHECK the CUDA path; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load it.  The product (xxmotif_b200/) never
 * links, loads or calls anything in oracle/.
 *
 * Pinned against dumps of the unmodified reference (oracle/ref_harness.cpp ->
 * tests/golden/NAME.trace.gz) by tests/test_oracle_vs_reference.py.
 */
#ifndef XXM_ORACLE_H
#define XXM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XO_PWM_LENGTH 30 /* Globals.h:27 */

/* same layout as sorted_sites_type, refinementPhase/Sorted_Sites.h:6-10 */
typedef struct {
	float score;
	int32_t startPos;
	int32_t sequence;
} xo_site;

typedef struct xo_seqset xo_seqset;
typedef struct xo_nullmodel xo_nullmodel;
typedef struct xo_tables xo_tables;
typedef struct xo_ctx xo_ctx;

/* run-time switches that the reference keeps in Global:: statics (Globals.h:113-537) */
typedef struct {
	int32_t revcomp;              /* --revcomp */
	int32_t mops;                 /* Global::multipleOccurrence */
	int32_t oops;                 /* Global::oneOccurrence */
	int32_t localization;         /* Global::usePositionalProbs */
	int32_t maxMotifsPerSequence; /* default 1000 */
	int32_t maxSeqCount;          /* default 1000 */
	double plusFrac;              /* default 0 */
	double overrepCorrection;     /* default 1.05 */
	double bg[5];                 /* negBg if a negSet exists else posBg */
	double bgLog[5];
} xo_config;

/* ---- sequences (seqset.cpp, seqFormat/Alignment.h:23-41) ---- */
xo_seqset* xo_seqset_new(int32_t nent, const int32_t* n, const uint8_t* flat_codes);
void xo_seqset_free(xo_seqset*);
void xo_seqset_features(xo_seqset*); /* calculateSequenceFeatures, seqset.cpp:21-36 */
void xo_seqset_revcomp(xo_seqset*);  /* createRevcomp, seqset.cpp:128-178 */
int32_t xo_seqset_nent(const xo_seqset*);
int32_t xo_seqset_max_leng(const xo_seqset*);
int32_t xo_seqset_min_leng(const xo_seqset*);
const int32_t* xo_seqset_n(const xo_seqset*);        /* [nent], 0-based array of lengths */
const int32_t* xo_seqset_total(const xo_seqset*);    /* [nent+1] */
const double* xo_seqset_avgLength(const xo_seqset*); /* [nent+1] */
int64_t xo_seqset_flat(const xo_seqset*, uint8_t* out); /* concatenated S[0][1..n]; returns byte count (out may be NULL) */
/* ElongCandidates::countAllOccurrences' counting loop (elongationCandidates.cpp:234-262); counts[4^ncoding] */
void xo_count_gapped_kmers(const xo_seqset*, int32_t nseq, int32_t ncoding, const int32_t* offset, int32_t firstHalfOnly, uint32_t* counts);

/* ---- background (backgroundDistribution.cpp:30-95, NullModel.cpp:88-326) ---- */
void xo_background(const xo_seqset*, double* bg5, double* bgLog5);
xo_nullmodel* xo_nullmodel_new(const xo_seqset*, float alpha, float beta, int32_t order);
xo_nullmodel* xo_nullmodel_from_tables(int32_t order, const float* rprobs, const float* rconds);
void xo_nullmodel_free(xo_nullmodel*);
int32_t xo_nullmodel_rfields(const xo_nullmodel*);
const float* xo_nullmodel_rprobs(const xo_nullmodel*);
const float* xo_nullmodel_rconds(const xo_nullmodel*);

/* ---- p-value tables (branch_and_bound-inl.h:183-603) ---- */
xo_tables* xo_tables_new(const double* pwm31x5, const int32_t* cols, int32_t ncols, int32_t posSetSize,
                         int32_t totalPositions, const double* bgLog5, const xo_nullmodel*);
void xo_tables_free(xo_tables*);
int32_t xo_tables_error(const xo_tables*); /* non-zero: reference behaviour undefined for this column set */
int32_t xo_tables_L(const xo_tables*);
int32_t xo_tables_maxk(const xo_tables*);
int32_t xo_tables_nparts(const xo_tables*);
int64_t xo_tables_size(const xo_tables*, int32_t part);           /* 4^len(part) */
const double* xo_tables_pvalues(const xo_tables*, int32_t part);  /* dense pValueHash[part] */
const double* xo_tables_probs(const xo_tables*);                  /* dense probabilityHash (part 0) */
int32_t xo_tables_nclasses(const xo_tables*, int32_t part);
const int32_t* xo_tables_class_score(const xo_tables*, int32_t part);
const int32_t* xo_tables_class_count(const xo_tables*, int32_t part);
const int64_t* xo_tables_class_ids(const xo_tables*, int32_t part); /* flattened in class order */
double xo_tables_lookup_window(const xo_tables*, const uint8_t* window); /* get_kmer_probability_all_positions :362 */

/* ---- scan / select layer (StartPosUpdater*, pValCalculation.*) ---- */
xo_ctx* xo_ctx_new(const xo_seqset*, const xo_config*);
void xo_ctx_free(xo_ctx*);
/* test plumbing (no counterpart in the reference): several contexts on one set, see xxm_oracle.c */
int64_t xo_ctx_cap(const xo_ctx*);
void xo_ctx_copy_candidates(xo_ctx* dst, const xo_ctx* src);
void xo_ctx_set_elong(xo_ctx*, const xo_site* sites, int64_t n);
const xo_site* xo_ctx_best(const xo_ctx*);         /* bestStartPos */
const xo_site* xo_ctx_sites(const xo_ctx*);        /* sortedSites */
const xo_site* xo_ctx_elong(const xo_ctx*);        /* sortedSites_elong */
int32_t xo_ctx_nbSites(const xo_ctx*);
int32_t xo_ctx_nbBestSites(const xo_ctx*);
int32_t xo_ctx_maxSites(const xo_ctx*);
const double* xo_ctx_positionalProb(const xo_ctx*);

void xo_init_positional_probs(xo_ctx*, int32_t motifLength, int32_t firstMotifColumn, int32_t enrStart,
                              int32_t enrEnd, int32_t offset); /* pValCalculation.cpp:195-248 */
double xo_calculate_pval(xo_ctx*, const xo_tables*, int32_t seq, int32_t pos, int32_t len, int32_t delta_0,
                         int32_t length, double avgLength, int32_t mops, double* pValScore); /* pValCalculation.h:106 */

/* fill_sites_with_possible_startPositions, StartPosUpdater-inl.h:70-140 */
void xo_scan_full(xo_ctx*, const xo_tables*, const int32_t* cols, int32_t ncols, int32_t posSetSize,
                  int32_t enrStart, int32_t enrEnd, int32_t keepPosSetSize);
/* update_PWM_with_bestSubset (which==0, sortedSites) / update_elongPWM_with_bestSubset (which==1), StartPosUpdater.cpp:78-112 */
int32_t xo_update_pwm_with_best_subset(xo_ctx*, int32_t which, double* pwm31x5, const int32_t* cols, int32_t ncols,
                                       int32_t posSetSize, int32_t oldTotalSites, int32_t plus, double pseudo,
                                       int32_t* bestMotifNb, double* bestPval);
/* fill_elongSites_with_possible_startPositions, StartPosUpdater-inl.h:143-221 */
void xo_rescore(xo_ctx*, const xo_tables*, int32_t firstMotifColumn, const int32_t* tmpCols, int32_t ntmp,
                int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t palin);
/* fill_sites_with_possible_startPositions_oops, StartPosUpdater-inl.h:224-266 */
void xo_scan_oops(xo_ctx*, const xo_tables*, const int32_t* cols, int32_t ncols, int32_t posSetSize,
                  int32_t enrStart, int32_t enrEnd);
/* set_BindingSites, StartPosUpdater.cpp:65-76 with threshold "<thr" ; returns count, sites sorted by start position */
int32_t xo_set_binding_sites(xo_ctx*, double thr);
/* fill_sites_with_startPosList, StartPosUpdater-inl.h:6-68 ; site positions are relative to PWM column 1 */
void xo_rescore_startposlist(xo_ctx*, const xo_tables*, const int32_t* cols, int32_t ncols, int32_t posSetSize,
                             int32_t enrStart, int32_t enrEnd, const int32_t* siteSeq, const int32_t* sitePos,
                             int32_t nSites, int32_t instanceThreshold);
/* fill_startPosList_with_bestSubset, StartPosUpdater.cpp:52-63 ; outputs the new site list (<= cap entries) */
int32_t xo_startposlist_best_subset(xo_ctx*, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t oldTotalSites,
                                    int32_t* outSeq, int32_t* outPos, int32_t cap, int32_t* bestMotifNb, double* bestPval);
/* fill_startPosList_with_elongSites, StartPosUpdater.cpp:114-134 ; returns worst score via *worst */
int32_t xo_startposlist_from_elong(xo_ctx*, int32_t bestMotifNb, int32_t firstTmpColumn, int32_t* outSeq, int32_t* outPos,
                                   int32_t cap, double* worst);
/* qsort restatements (Sorted_Sites.h:15-56; glibc qsort = stable merge sort) on an external array */
void xo_sort_by_score(xo_site* sites, int32_t n);
void xo_sort_by_startpos(xo_site* sites, int32_t n, int32_t palin);
/* Motif::updatePWM_OOL, Motif.cpp:617-651 */
void xo_update_pwm_ool(xo_ctx*, const xo_site* sites, double* pwm31x5, int32_t firstColumn, int32_t K, int32_t posSetSize,
                       double pseudo);
/* utils.h:80-102 */
double xo_order_statistics_pvalue(const xo_ctx*, int32_t motifNb, int32_t N, double p);
/* utils.h:138-162 */
float xo_fast_log(float x);
/* Output::printScoreDistribution counts, output.cpp:1033-1078 (+ Motif.h:584-612, Motif.cpp:1068-1079) */
void xo_logodds_hist(const xo_ctx*, const double* pwm31x5, const int32_t* cols, int32_t ncols, int32_t posSetSize,
                     double* fwd, double* bwd, int32_t npos, double* maxScore, double* maxScore2);
/* setPositionalPval, StartPosUpdater.cpp:207-274 ; site positions relative to PWM column 1 */
double xo_positional_pval(const xo_ctx*, int32_t enrMax, int32_t firstMotifColumn, int32_t motifLength,
                          const int32_t* sitePos, int32_t nSites);

/* MotifRegion::getRegion, nucleotides/motifRegion.cpp:67-188 ; out = {set, max, startRegion, endRegion} */
void xo_motif_region(const xo_ctx*, const int32_t* startPosCounts, int32_t counts, int32_t startRange, int32_t endRange, int32_t* out);
/* Motif::updateEnrichedRegion(bestStartPos), Motif.cpp:968-1019 ; best = score-sorted candidates terminated by score 1;
 * enr = {max, startRegion, endRegion} relative to PWM column 1 (in/out) */
void xo_update_enriched_region(const xo_ctx*, const xo_site* best, int32_t revcomp, int32_t max_leng, int32_t firstMotifColumn, int32_t motifLength,
                               int32_t* enr);

#ifdef __cplusplus
}
#endif
#endif
