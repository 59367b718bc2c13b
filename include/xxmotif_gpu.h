/*
 * xxmotif_gpu.h -- C ABI of libxxmotif_gpu.so: the B200 (sm_100a) implementation of XXmotif's PWM-refinement
 * hot path (window scan -> site p-value -> site selection -> PWM recount), one context per GPU.
 *
 * Every entry point replaces one reference call site (paths relative to soedinglab/xxmotif src/).  The reference
 * has no FFI of its own: the seam is the set of member functions of the process-wide singletons StartPosUpdater,
 * PVal_Calculator and Kmer_Generator_Reloaded<dna_hash_t>; INTEGRATION.md shows the
 * binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are HOST pointers; the library owns every device buffer.
 *   - every function returns 0 on success, non-zero on error; xxm_last_error() gives the message.  There is no
 *     CPU fallback: without a CUDA device xxm_create() fails.
 *   - one host thread per context (the reference is single-threaded and non-reentrant).
 *   - sequences are numbered 1..nent and positions 1..n exactly like the reference (Alignment.h:23-41); with
 *     revcomp the positions run over the doubled sequence fwd + N + revcomp (seqset.cpp:128-178).
 *   - xxm_site is bit-identical to sorted_sites_type (refinementPhase/Sorted_Sites.h:6-10).
 */
#ifndef XXMOTIF_GPU_H
#define XXMOTIF_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XXM_PWM_LENGTH 30 /* Globals.h:27; PWM arrays are double[31][5], row 0 and column 0 unused/counts */

typedef struct xxm_ctx xxm_ctx;

typedef struct {
	float score;
	int32_t startPos;
	int32_t sequence;
} xxm_site;

/* The Global:: switches the path reads (Globals.h:113-537, Globals.cpp:36-124 defaults). */
typedef struct {
	int32_t mops;                 /* Global::multipleOccurrence   (--mops)          */
	int32_t oops;                 /* Global::oneOccurrence        (--oops)          */
	int32_t localization;         /* Global::usePositionalProbs   (--localization)  */
	int32_t maxMotifsPerSequence; /* Global::maxMotifsPerSequence, default 1000     */
	int32_t maxSeqCount;          /* Global::maxSeqCount, default 1000              */
	double plusFrac;              /* Global::plusFrac, default 0                    */
	double overrepCorrection;     /* Global::overrepCorrection, default 1.05        */
	double bg[5];                 /* Global::negBg if a negSet exists else posBg    */
	double bgLog[5];              /* matching *_log array                           */
} xxm_options;

/* which site list a call refers to (members of StartPosUpdater, StartPosUpdater.h:72-74) */
enum { XXM_LIST_SORTED_SITES = 0, XXM_LIST_SORTED_SITES_ELONG = 1, XXM_LIST_BEST_STARTPOS = 2 };
/* order in which xxm_get_sites() materialises a list */
enum { XXM_ORDER_AS_IS = 0, XXM_ORDER_BY_SCORE = 1, XXM_ORDER_BY_STARTPOS = 2 };

/* ---- life cycle ----------------------------------------------------------------------------------------------- */
int xxm_create(int device, xxm_ctx** out);
void xxm_destroy(xxm_ctx* ctx);
const char* xxm_last_error(const xxm_ctx* ctx); /* ctx may be NULL: message of the last failed xxm_create */
int xxm_set_options(xxm_ctx* ctx, const xxm_options* opt);

/* ---- a1: the encoded sequence set ----------------------------------------------------------------------------------
 * replaces: Global::posSet as filled by Alignment::getEntity (seqFormat/Alignment.cpp:149-195) + createRevcomp
 * (seqset.cpp:128-178) + calculateSequenceFeatures (seqset.cpp:21-36).
 * n[i], codes: the sequences BEFORE createRevcomp, codes 0=N,1=A,2=C,3=G,4=T concatenated without separators.
 * The set is packed once (2-bit codes + 1-bit N mask, reverse strand included when revcomp != 0) and then stays resident in
 * HBM.  With at least 6 packing threads (XXM_PACK_THREADS, default: cores - 2 up to 14 on hosts with 12 or more cores, else 3/4 of them) the host cores pack it
 * chunk by chunk while earlier chunks are on the bus, so 0.375 bytes per stored base cross PCIe; with fewer, the raw bytes
 * are copied and packed on the device.  total[]/avgLength[] follow the
 * reference's arithmetic, quirks included. */
int xxm_load_seqset(xxm_ctx* ctx, int32_t nent, const int32_t* n, const uint8_t* codes, int32_t revcomp);
int xxm_seqset_info(const xxm_ctx* ctx, int32_t* total /*[nent+1] or NULL*/, double* avgLength /*[nent+1] or NULL*/,
                    int32_t* max_leng, int32_t* min_leng);
/* how the last xxm_load_seqset moved the set when it used both packing paths at once (2-5 host threads per GPU: raw chunks from
 * the front of the set over the bus, packed on the device, while the host threads pack chunks from the back): rawChunks of chunks
 * went raw; 0 of 0 when one path did everything.  XXM_LOAD_PATH=host|device|hybrid overrides the choice. */
int xxm_last_load_info(const xxm_ctx* ctx, int32_t* rawChunks, int32_t* chunks);

/* ---- a5: background + p-value tables -------------------------------------------------------------------------------
 * xxm_set_background replaces the NullModel statics read by calculate_kmer_probability (NullModel.cpp:42-54,
 * NullModel.h:13): order and the two 2^(3(order+1)) float tables _rprobs/_rconds.
 * xxm_reinitialize replaces Kmer_Generator_Reloaded<dna_hash_t>::reinitialize(pwm, cols, posSetSize, totalPositions)
 * (branch_and_bound-inl.h:183-258); pwm is double[31][5] row-major (log probabilities), cols the match columns. */
int xxm_set_background(xxm_ctx* ctx, int32_t order, const float* rprobs, const float* rconds);
int xxm_reinitialize(xxm_ctx* ctx, const double* pwm31x5, const int32_t* cols, int32_t ncols, int32_t posSetSize,
                     int32_t totalPositions);
/* introspection of the tables of the last xxm_reinitialize (tests): nparts, match columns, MAX_KMER_SIZE, and per part
 * the dense pValueHash (size 4^len), the score classes (descending) with their id lists; probabilityHash for part 0 */
int xxm_tables_info(const xxm_ctx* ctx, int32_t* L, int32_t* maxk, int32_t* nparts, int64_t size[2], int32_t nclasses[2], int64_t nids[2]);
int xxm_tables_get(const xxm_ctx* ctx, int32_t part, double* pvalues /*size*/, double* probs /*size, part 0 only, may be NULL*/,
                   int32_t* classScore, int32_t* classCount, int64_t* classIds);

/* ---- a3/a4/a13: site p-value of given windows -----------------------------------------------------------------------
 * replaces PVal_Calculator::calculatePval(kmerHash, kmer, cols, seq, pos, len, delta_0, size, length, avgLength,
 * pValScore, mops) (pValCalculation.h:106-114) for a batch of windows of the current model; used by
 * StartPosUpdater::getProbabilitiesPerSite (StartPosUpdater.h:91-132).  firstMotifColumn/enr*: arguments of
 * initPositionalProbs (pValCalculation.cpp:195) when localization is on. */
int xxm_window_pvals(xxm_ctx* ctx, int32_t nwin, const int32_t* seq, const int32_t* pos, int32_t posSetSize, int32_t firstMotifColumn,
                     int32_t enrStart, int32_t enrEnd, int32_t mops, double* finalPval, double* pValScore);

/* ---- a2: full scan -------------------------------------------------------------------------------------------------
 * replaces StartPosUpdater::fill_sites_with_possible_startPositions(kmerHash, keepPosSetSize=true, motif)
 * (StartPosUpdater-inl.h:70-140) incl. updateMultOcc (StartPosUtils.h:12-65).  Results stay on the device:
 * list BEST_STARTPOS (all windows with finalPval < 1) and list SORTED_SITES (after overlap removal). */
int xxm_scan_full(xxm_ctx* ctx, int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t* nbSites, int32_t* nbBestSites,
                  int32_t* maxSites);

/* ---- a10: candidate re-score under the current (variant) model ------------------------------------------------------
 * replaces fill_elongSites_with_possible_startPositions(kmerHash, motif, pwm, tmpColumns, palin)
 * (StartPosUpdater-inl.h:143-221); re-scores list BEST_STARTPOS into list SORTED_SITES_ELONG. */
int xxm_rescore(xxm_ctx* ctx, int32_t firstMotifColumn, int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t palin,
                int32_t* nbSites, int32_t* maxSites);

/* ---- a11: final OOPS scan ------------------------------------------------------------------------------------------
 * replaces fill_sites_with_possible_startPositions_oops (StartPosUpdater-inl.h:224-266): best site of EVERY sequence
 * with mops=false into list SORTED_SITES (one slot per sequence, score 1 = none). */
int xxm_scan_oops(xxm_ctx* ctx, int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t* nbSites, int32_t* maxSites);

/* ---- a12: re-score a motif's own site list -------------------------------------------------------------------------
 * replaces fill_sites_with_startPosList(kmerHash, motif, instanceThreshold) (StartPosUpdater-inl.h:6-68);
 * sitePos is relative to PWM column 1 like Motif::_sites. */
int xxm_rescore_startposlist(xxm_ctx* ctx, int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t nSites, const int32_t* siteSeq,
                             const int32_t* sitePos, int32_t instanceThreshold, int32_t* nbSites, int32_t* maxSites);

/* ---- a7/a8/a9: sort, optimal subset, PWM recount -------------------------------------------------------------------
 * xxm_update_pwm_with_best_subset replaces update_PWM_with_bestSubset (which = SORTED_SITES) and
 * update_elongPWM_with_bestSubset (which = SORTED_SITES_ELONG) (StartPosUpdater.cpp:78-112): stable sort by float
 * score, getBestSubset (StartPosUpdater.h:135-238), Motif::updatePWM_OOL (Motif.cpp:617-651).  pwm31x5 is overwritten
 * like the reference's pwm argument.  *updated = 0 when the reference would return early (no site / best score 1). */
int xxm_update_pwm_with_best_subset(xxm_ctx* ctx, int32_t which, double* pwm31x5, int32_t firstColumn, int32_t posSetSize, int32_t plus,
                                    double pseudo, int32_t* updated, int32_t* bestMotifNb, double* bestPval);
/* replaces fill_startPosList_with_bestSubset (StartPosUpdater.cpp:52-63): outputs the chosen sites (relative to PWM
 * column 1) sorted by start position; *nOut = -1 when the reference would return early */
int xxm_startposlist_best_subset(xxm_ctx* ctx, int32_t firstColumn, int32_t posSetSize, int32_t* outSeq, int32_t* outPos, int32_t cap,
                                 int32_t* nOut, int32_t* bestMotifNb, double* bestPval);
/* replaces fill_startPosList_with_elongSites (StartPosUpdater.cpp:114-134) */
int xxm_startposlist_from_elong(xxm_ctx* ctx, int32_t bestMotifNb, int32_t firstTmpColumn, int32_t* outSeq, int32_t* outPos, int32_t cap,
                                int32_t* nOut, double* worstScore);
/* maximizeMotifLength only uses the site list of the best column variant (Iterate_Motif.cpp:117-125 overwrites the list at every
 * improvement).  xxm_elong_keep sets the current sortedSites_elong aside on the device (no copy: the following re-scores write a
 * second set of buffers); xxm_startposlist_from_kept = xxm_startposlist_from_elong on the list set aside, called once at the end */
int xxm_elong_keep(xxm_ctx* ctx);
int xxm_startposlist_from_kept(xxm_ctx* ctx, int32_t bestMotifNb, int32_t firstTmpColumn, int32_t* outSeq, int32_t* outPos, int32_t cap,
                               int32_t* nOut, double* worstScore);
/* replaces set_BindingSites (StartPosUpdater.cpp:65-76) with ThresholdChecker "<thr": sites sorted by start position */
int xxm_set_binding_sites(xxm_ctx* ctx, double thr, xxm_site* out, int32_t cap, int32_t* nOut);
/* replaces getBestStartPos()/getSortedSites() (StartPosUpdater.h:46-47) + sortSitesByScore/sortSitesByStartPos
 * (StartPosUtils.h:68-78): copies a device-resident list to the host in the requested order */
int xxm_get_sites(xxm_ctx* ctx, int32_t which, int32_t order, int32_t palin, xxm_site* out, int32_t cap, int32_t* nOut);
/* the number of elements xxm_get_sites would write for this list and order (no device work) */
int xxm_sites_count(xxm_ctx* ctx, int32_t which, int32_t order, int32_t* n);

/* ---- a14: forward / reverse-complement log-odds histogram ----------------------------------------------------------
 * replaces the counting loops of Output::printScoreDistribution (output.cpp:1033-1078) with
 * Motif::calculateMatrixScore_logOdds / getMaxMatrixScore / getPalindrome (Motif.h:584-612, Motif.cpp:1068-1079).
 * fwd/bwd receive npos = max_leng - ncols + 1 counts. */
int xxm_logodds_hist(xxm_ctx* ctx, const double* pwm31x5, const int32_t* cols, int32_t ncols, int32_t posSetSize, double* fwd, double* bwd,
                     int32_t npos, double* maxScore, double* maxScore2);

/* ---- a15: localization statistics ------------------------------------------------------------------------------------
 * xxm_update_enriched_region replaces Motif::updateEnrichedRegion(StartPosUpdater::getBestStartPos()) (Motif.cpp:968-1019,
 * called from MotifContainer::iterate, MotifContainer.cpp:241) incl. MotifRegion::getRegion (motifRegion.cpp:67-188):
 * histogram of the start positions of the candidates with score < 1e-3 on the device, region search on the host.
 * enr = {max, startRegion, endRegion} of Motif::_enrichment (relative to PWM column 1), updated in place.
 * xxm_host_positional_pval replaces StartPosUpdater::setPositionalPval (StartPosUpdater.cpp:207-274) for a site list
 * (positions relative to PWM column 1); returns 2 where the reference prints "In no man's land" and exits. */
int xxm_update_enriched_region(xxm_ctx* ctx, int32_t firstMotifColumn, int32_t motifLength, int32_t* enr);
int xxm_host_positional_pval(int32_t max_leng, int32_t revcomp, int32_t enrMax, int32_t firstMotifColumn, int32_t motifLength, const int32_t* sitePos,
                             int32_t nSites, double* out);
/* the same for the set loaded in ctx; on a sharded set every rank passes its own sites and the sum runs over all ranks */
int xxm_positional_pval(xxm_ctx* ctx, int32_t enrMax, int32_t firstMotifColumn, int32_t motifLength, const int32_t* sitePos, int32_t nSites, double* out);

/* ---- the callers of the path, re-stated above this ABI -----------------------------------
 * xxm_motif mirrors the part of class Motif (refinementPhase/Motif.h) that the refinement loop reads and writes.
 * xxm_optimize_motif        = IterateMotif::optimizeMotif(motif, true, false)      (Iterate_Motif.cpp:23-37)
 * xxm_maximize_motif_length = IterateMotif::maximizeMotifLength(motif, lengthCorrection, false)  (Iterate_Motif.cpp:40-149)
 * xxm_iterate_motif         = the per-motif loop of MotifContainer::iterate        (MotifContainer.cpp:231-264)      */
typedef struct {
	double pwm[31 * 5];   /* Motif::_pwm, log probabilities, row-major [column 0..30][0..4] */
	int32_t ncols;
	int32_t cols[30];     /* Motif::_motifColumns */
	int32_t posSetSize;   /* Motif::_posSetSize */
	int32_t enr[3];       /* Motif::_enrichment: max, startRegion, endRegion */
	double pval;          /* Motif::_pVal (log) */
	double posPval;       /* Motif::_posPVal (log) */
	double worstScore;    /* Motif::_worstScore */
	int32_t nsites;       /* Motif::_sites: sequence and position relative to PWM column 1 */
	int32_t siteCap;
	int32_t* siteSeq;
	int32_t* sitePos;
} xxm_motif;
typedef struct {
	int32_t totalPositions;      /* Global::posSet->total[posSetSize] */
	int32_t revcomp;             /* Global::revcomp */
	int32_t localization;        /* Global::usePositionalProbs */
	int32_t maximizeMotifLength; /* Global::maximizeMotifLength (false with --no-pwm-length-optimization) */
	int32_t minMatchPositions;   /* Global::minMatchPositions, default 5 */
	int32_t maxMatchPositions;   /* Global::maxMatchPositions, default 17 */
	double pseudo;               /* Global::pseudo, default 0.1 */
	double neff_pwm;             /* Global::neff_pwm, default 10 */
	double* elongPWM;            /* optional double[31*5]: IterateMotif::elongPWM carried between calls (NULL: zeros) */
} xxm_refine_params;
int xxm_optimize_motif(xxm_ctx* ctx, xxm_motif* m, const xxm_refine_params* p);
int xxm_maximize_motif_length(xxm_ctx* ctx, xxm_motif* m, const xxm_refine_params* p, int32_t lengthCorrection, int32_t* noChange);
int xxm_iterate_motif(xxm_ctx* ctx, xxm_motif* m, const xxm_refine_params* p, int32_t* rounds);

/* ---- several motifs at once on one GPU ("batch motifs ... so small inputs still fill the GPU") ----
 * MotifContainer::iterate (MotifContainer.cpp:216-270) refines its motifs one after the other, and nothing the loop body
 * touches is shared between them (the singletons' buffers are scratch, IterateMotif::elongPWM is rewritten by iteration 0
 * of every column variant before iteration 1 reads it, Iterate_Motif.cpp:78-86).  On the sets of configs 1-3 and 5 one
 * motif's kernels occupy a fraction of a B200 and its host thread spends as long building tables as the device scanning.
 * xxm_share_seqset makes ctx a VIEW of owner: it reads owner's packed stream, per-sequence arrays, options and NullModel
 * tables (nothing is copied on the device) and keeps its own stream, model, tables and site lists.  Both contexts must be
 * on one device; owner must hold a loaded, unsharded set and may not reload it, change its background or be destroyed
 * while views exist (xxm_load_seqset / xxm_set_background fail, xxm_view_count says how many).  A view becomes an
 * ordinary context again by loading its own set.  xxm_view_create = xxm_create on owner's device + xxm_share_seqset.
 * xxm_iterate_motifs runs xxm_iterate_motif for nmotifs motifs on nctx contexts of one set (ctxs[0] and its views, or any
 * contexts holding the same set), one host thread per context, motifs handed out in order as contexts fall free; every
 * motif's result is bit-identical to a serial xxm_iterate_motif (each runs start to end on one context).  params holds
 * nparams = 1 (all motifs) or nmotifs entries; rounds[nmotifs] as in xxm_iterate_motif.  Returns 0, or the first non-zero
 * code of a motif (*failed = its index, text in xxm_last_error of the context that ran it = ctxs[*failedCtx]). */
/* The column variants of ONE motif several at a time (the other half of that row: "batch the <= 11 column variants x 3 inner
 * iterations of maximizeMotifLength").  A variant depends on the motif's PWM, its columns and the candidate list of the motif's
 * last full scan, and on no other variant (iteration 0 rewrites elongPWM in all 30 columns before iteration 1 reads it, or the
 * variant is dropped, Iterate_Motif.cpp:73-90); only the choice among them is sequential (:103) and is made once all are known.
 * xxm_share_candidates lends a view the (sequence, position) arrays of its owner's bestStartPos until the owner scans again;
 * xxm_maximize_motif_length_on / xxm_iterate_motif_on take ctxs[0] = the context that holds the motif's scan and nctx - 1 views
 * of it, one host thread per context, variants handed out in loop order; nctx = 1 is xxm_maximize_motif_length / _iterate_motif.
 * Results are bit-identical to the one-context calls (return 3 = internal inconsistency, never seen). */
int xxm_share_candidates(xxm_ctx* ctx, xxm_ctx* owner);
int xxm_maximize_motif_length_on(xxm_ctx* const* ctxs, int32_t nctx, xxm_motif* m, const xxm_refine_params* p, int32_t lengthCorrection, int32_t* noChange);
int xxm_iterate_motif_on(xxm_ctx* const* ctxs, int32_t nctx, xxm_motif* m, const xxm_refine_params* p, int32_t* rounds);
int xxm_share_seqset(xxm_ctx* ctx, xxm_ctx* owner);
int xxm_view_create(xxm_ctx* owner, xxm_ctx** out);
int xxm_view_count(const xxm_ctx* ctx);
int xxm_iterate_motifs(xxm_ctx* const* ctxs, int32_t nctx, xxm_motif* motifs, int32_t nmotifs, const xxm_refine_params* params, int32_t nparams,
                       int32_t* rounds, int32_t* failed, int32_t* failedCtx);

/* ---- a8/a9 split in phases, for a sequence set sharded over several GPUs (one context per GPU) ---------
 * The reference chooses the optimal subset over ALL sites (getBestSubset) and recounts the PWM from them.  With shards,
 * each context reports the run-length histogram of its float site scores (xxm_score_runs, ascending, score 1 excluded);
 * the caller merges the histograms of all ranks (NCCL all-gather), chooses the subset once (xxm_host_choose_subset),
 * every context counts its sites with score <= the subset's last score (xxm_count_sites), the 31x5 count matrices are
 * summed (NCCL all-reduce) and turned into the new PWM (xxm_host_build_pwm = Motif::buildPWM, Motif.cpp:584-591). */
int xxm_score_runs(xxm_ctx* ctx, int32_t which, float* scores, int64_t* counts, int32_t cap, int32_t* nruns);
int xxm_count_sites(xxm_ctx* ctx, int32_t which, float thr, int32_t firstColumn, double* counts31x5, int64_t* nsites);
int xxm_host_choose_subset(const float* scores, const int64_t* counts, int32_t nruns, int64_t nbSites, int64_t limit, int64_t maxSites,
                           int32_t oneOccurrence, double plusFrac, int64_t* bestMotifNb, double* bestPval, float* lastScore);
int xxm_host_build_pwm(const double* counts31x5, double pseudoTimesK, const double* bg5, double* pwm31x5);

/* ---- one optimizeMotif iteration as one call, over a possibly sharded set --------------------------------------------
 * xxm_optimize_step = one pass of the loop body of IterateMotif::optimizeMotif (Iterate_Motif.cpp:26-32): reinitialize,
 * full scan + updateMultOcc, getBestSubset over ALL sites, Motif::updatePWM_OOL; pwm31x5 is overwritten with the new PWM.
 * With a communicator attached (xxm_comm_init; one process and one context per GPU, every rank holding its own block of
 * sequences) the value-index histogram of the kept sites and the 31x5 counts are summed over NCCL on the device before the
 * host sees them; posSetSize / totalPositions are the rank's own, the global ones are their sums.
 * xxm_comm_unique_id: ncclGetUniqueId on rank 0; the caller distributes the 128 bytes (MPI, torch.distributed, a file). */
typedef struct {
	int64_t nbSites;       /* kept sites of this rank after updateMultOcc */
	int64_t nbBestSites;   /* windows with finalPval < 1 of this rank */
	int64_t nbSitesGlobal; /* kept sites of all ranks */
	int64_t maxSites;      /* StartPosUpdater::maxSites over all ranks */
	int64_t bestMotifNb;
	double bestPval;
	int32_t updated;       /* 0: the reference returns early (no site) and the PWM is unchanged */
} xxm_step_info;
/* for callers that run the exchange themselves (the split-phase entry points above, any transport): the counts of the WHOLE
 * sharded set -- sums of nent, total[nent] and xxm_seqset_logsum over the ranks, maximum of max_leng -- which then replace the
 * rank's own in table thresholds, maxSites and the ZOOPS exponent (sumLogLength = NaN keeps the rank's own avgLength).
 * Valid until the next xxm_load_seqset; contexts with a communicator derive all of this themselves. */
int xxm_set_global_counts(xxm_ctx* ctx, int64_t nentGlobal, int64_t totalGlobal, double sumLogLength, int32_t maxLengGlobal);
int xxm_seqset_logsum(const xxm_ctx* ctx, double* sumLogLength);
/* the running log sum continued over this rank's sequences, one addition per sequence: chained through the ranks in order it is
 * the reference's sequential sum (seqset.cpp:31-34), bit for bit */
int xxm_seqset_logsum_from(const xxm_ctx* ctx, double start, double* end);
int xxm_comm_unique_id(uint8_t* id128);
int xxm_comm_init(xxm_ctx* ctx, int32_t rank, int32_t nranks, const uint8_t* id128);
int xxm_comm_destroy(xxm_ctx* ctx);
int xxm_optimize_step(xxm_ctx* ctx, double* pwm31x5, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t totalPositions, int32_t enrStart,
                      int32_t enrEnd, double pseudo, xxm_step_info* info);

/* ---- the counting loop of the elongation phase's seed search ---------------------------------------
 * ElongCandidates::countAllOccurrences (src/elongationPhase/elongationCandidates.cpp:234-262): for every start position of
 * sequences 1..nseq whose `ncoding` coding characters (character l at base offset offset[l] = l + gap[l-1]; gap positions may
 * hold N) are all non-N, count the gapped k-mer.  counts[id], id = sum_l (code_l - 1) * 4^(ncoding-1-l): the reference's id
 * with 2 instead of 4 bits per character (its IUPAC merge, :264-273, starts from exactly these 4^ncoding entries).
 * firstHalfOnly = 1: seed types PALINDROME / TANDEM under --revcomp walk only the first (n - span + 1) / 2 start positions of
 * the doubled sequence (:245).  With a communicator attached the counts are those of the whole sharded set. */
int xxm_count_gapped_kmers(xxm_ctx* ctx, int32_t nseq, int32_t ncoding, const int32_t* offset, int32_t firstHalfOnly, uint32_t* counts);

/* ---- multi-device context: ONE host thread, several GPUs ---------------------------------------------------------------
 * The reference is single-threaded.  xxm_multi owns one context per device (devices 0..ngpus-1), cuts the set
 * into contiguous blocks of sequences balanced by residues, runs every call on all devices concurrently (one worker thread per
 * device inside the library) and presents the union: counts summed, site lists concatenated in sequence order or merged stably by
 * score.  The shards are given the counts of the whole set, so tables, maxSites and the ZOOPS exponent are the whole set's.
 * With ngpus = 1 every call is a pass-through.  With more than one device every call refers to the whole set (posSetSize = nent).
 * This is the surface integration/xxmotif_gpu_binding.cpp uses (XXM_DEVICES=N). */
typedef struct xxm_multi xxm_multi;
int xxm_multi_create(int32_t ngpus, xxm_multi** out);
int xxm_multi_create_on(const int32_t* devices /*[ngpus] or NULL = 0..ngpus-1*/, int32_t ngpus, xxm_multi** out);
void xxm_multi_destroy(xxm_multi* m);
const char* xxm_multi_last_error(const xxm_multi* m);
int32_t xxm_multi_devices(const xxm_multi* m);
int xxm_multi_set_options(xxm_multi* m, const xxm_options* opt);
int xxm_multi_set_background(xxm_multi* m, int32_t order, const float* rprobs, const float* rconds);
int xxm_multi_load_seqset(xxm_multi* m, int32_t nent, const int32_t* n, const uint8_t* codes, int32_t revcomp);
int xxm_multi_seqset_info(const xxm_multi* m, int32_t* total, double* avgLength /*unused, may be NULL*/, int32_t* max_leng, int32_t* min_leng);
int xxm_multi_reinitialize(xxm_multi* m, const double* pwm31x5, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t totalPositions);
int xxm_multi_scan_full(xxm_multi* m, int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t* nbSites, int32_t* nbBestSites, int32_t* maxSites);
int xxm_multi_rescore(xxm_multi* m, int32_t firstMotifColumn, int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t palin, int32_t* nbSites,
                      int32_t* maxSites);
int xxm_multi_scan_oops(xxm_multi* m, int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t* nbSites, int32_t* maxSites);
int xxm_multi_rescore_startposlist(xxm_multi* m, int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t nSites, const int32_t* siteSeq,
                                   const int32_t* sitePos, int32_t instanceThreshold, int32_t* nbSites, int32_t* maxSites);
int xxm_multi_get_sites(xxm_multi* m, int32_t which, int32_t order, int32_t palin, xxm_site* out, int32_t cap, int32_t* nOut);
int xxm_multi_count_gapped_kmers(xxm_multi* m, int32_t nseq, int32_t ncoding, const int32_t* offset, int32_t firstHalfOnly, uint32_t* counts);

/* ---- host-side pieces of the path, callable without a device (unit tests of the host logic) -------------------------
 * xxm_host_tables: the table construction xxm_reinitialize performs (branch_and_bound-inl.h:183-603), without staging.
 *   info[9] = L, maxk, nparts, size0, size1, nclasses0, nclasses1, nids0, nids1; the arrays of `part` may be NULL.
 * xxm_host_best_subset: getBestSubset (StartPosUpdater.h:135-238) over a score-sorted site array; logTotal is the
 *   largest index of LogTable::LOG_sum_i that may be read (>= maxSites). */
int xxm_host_tables(const double* pwm31x5, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t totalPositions, const double* bgLog5,
                    int32_t order, const float* rprobs, const float* rconds, int32_t part, int64_t* info, double* pvalues, double* probs,
                    int32_t* classScore, int32_t* classCount, int64_t* classIds);
/* xxm_device_tables: SURVEY 8(f) row 1 -- the same construction ON THE DEVICE for one-part models (at most 8 match columns):
 *   all 4^len branch-and-bound paths tested in parallel (the recursion's visiting order is the numeric order of the path
 *   index), stable sort by integer score through rank by counting, class p-values accumulated leaf by leaf by one lane --
 *   bit-identical to xxm_host_tables.  Needs xxm_set_options (bgLog) and xxm_set_background; outputs as xxm_host_tables
 *   (part 0); last_ms (may be NULL): device time of the build kernels.  xxm_reinitialize takes this path when the environment
 *   variable XXM_DEVICE_TABLES is set. */
int xxm_device_tables(xxm_ctx* ctx, const double* pwm31x5, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t totalPositions, int64_t* info,
                      double* pvalues, int32_t* classScore, int32_t* classCount, int64_t* classIds, float* last_ms);
/* xxm_host_null_model: NullModel::init(sequences, alpha = pseudocountsFactor, beta = countsOffset, order, gaps = true,
 *   freqs = NULL) (NullModel.cpp:64-201) -> the two 2^(3(order+1)) float tables xxm_set_background takes.  n/codes are the
 *   sequences as the reference holds them when the model is trained (AFTER createRevcomp if --revcomp).
 * xxm_host_background: posBg/negBg and their logs (checkSequenceSet, backgroundDistribution.cpp:50-95). */
int xxm_host_null_model(int32_t nent, const int32_t* n, const uint8_t* codes, int32_t order, float alpha, float beta, float* rprobs,
                        float* rconds);
int xxm_host_background(int32_t nent, const int32_t* n, const uint8_t* codes, double* bg5, double* bgLog5);
int xxm_host_best_subset(const float* score, const int32_t* seq, int32_t nbSites, int32_t posSetSize, int32_t maxSeqCount, int32_t maxSites,
                         int32_t nent, int32_t maxMotifsPerSequence, int32_t oneOccurrence, double plusFrac, int32_t* bestMotifNb,
                         double* bestPval);

/* xxm_host_pack_seqset: the packing step of xxm_load_seqset alone (csrc/host_pack.h), for tests: words = sum over the
 *   sequences of ceil((stored length + 32) / 32) + 2 entries of codes64 / nmask32 (the last two are the end-of-stream
 *   guard), notX[nent] = residues that are not N; threads > 1 packs contiguous blocks of sequences concurrently, as the load
 *   does.  Returns the 1-based index of a sequence with a code outside 0..4, or 0. */
int xxm_host_pack_seqset(int32_t nent, const int32_t* n, const uint8_t* codes, int32_t revcomp, int32_t threads, uint64_t* codes64, uint32_t* nmask32,
                         int32_t* notX);

/* xxm_host_pair_table: the table of the full-scan find kernel (csrc/pair_table.h).  ids = the enumerated left k-mers that
 *   count as scan hits (column c in bits [2c+1:2c]), coloff = base offset of each of the len0 left match columns.
 *   table1024[q] describes the two windows starting at bases 0 and 1 of the 5-base group q: a window whose bases 0..3 are
 *   the 4-mer x and bases 4..7 the 4-mer y is a hit iff a[x] + b[y] >= 128.  Returns 1 when the k-mer set is not
 *   representable that way (the scan then uses its generic kernel). */
int xxm_host_pair_table(const uint32_t* ids, int64_t nids, const int32_t* coloff, int32_t len0, uint32_t* table1024, int32_t* levels);

/* ---- measurement ----------------------------------------------------------------------------------------------------
 * device-side timing of the most recent call of each kernel (CUDA events on the context's stream), launch counts */
typedef struct {
	int64_t launches;        /* kernels launched by this context so far */
	float last_scan_ms;      /* device time of the kernels of the last full scan: find + expand */
	int64_t last_scan_bases; /* stored bases the last full scan covered */
	int64_t last_scan_hits;  /* windows with finalPval < 1 it found */
	float last_find_ms;      /* device time of the find kernel alone (the rest of last_scan_ms is the expand kernel) */
} xxm_stats;
int xxm_get_stats(const xxm_ctx* ctx, xxm_stats* out);
/* host-side phase times of xxm_optimize_step summed over the steps since the last reset, in microseconds: table build and
 * staging | launching scan + overlap removal + exchange | waiting for them | subset choice | launching the recount | waiting
 * for it | PWM from the counts.  (Where does a step spend the time the kernels do not account for?) */
int xxm_step_phases(xxm_ctx* ctx, double* us7, int64_t* nsteps, int32_t reset);
/* with XXM_DEBUG_GUARDS set in the environment every device buffer carries a 256-byte canary behind its last element;
 * xxm_check_guards counts the buffers of the process whose canary a kernel overwrote (damaged must be 0) */
int xxm_check_guards(xxm_ctx* ctx, int32_t* checked, int32_t* damaged);
/* with XXM_TRACE set in the environment the library sums the wall time of every C-ABI call (and of named sections inside
 * some of them) per name; xxm_trace_dump prints the table to stderr and clears it (xxm_destroy does the same) */
int xxm_trace_dump(void);
/* XXM_SCAN_GEN=6 in the environment (read by xxm_create) selects the previous generation of the pair-table scan kernels
 * (even/odd hit masks + k_expand6) instead of the default (window-order masks + k_expand7); both give the same site lists --
 * it exists for A/B timing and for the parity tests of the older kernels */
/* XXM_EXPAND5_TICKET=1: the expand kernel of the GENERIC scan (k_expand5; only models whose left table the pair table cannot
 * represent take it) hands its tiles out by ticket instead of by block index, which removes its one assumption about the order in
 * which blocks become resident.  Same site lists; verified under the SIMT emulator with blocks in reverse order, not yet on a GPU,
 * hence opt-in. */
/* the CUDA stream the context launches on, as a cudaStream_t cast to void* (for external event timing) */
void* xxm_stream(const xxm_ctx* ctx);

/* ---- 8(f) row 3: k-mer counting for the background models, on the device -------------------------------------------------
 * xxm_kmer_counts replaces the counting loop of NullModel::init (NullModel.cpp:122-132) over the sequence set loaded by
 * xxm_load_seqset (as stored: with revcomp the doubled sequences, like the reference after createRevcomp): counts[] is
 * indexed like NullModel::_counts (sub2ind, base 5, digits 1..4 = ACGT, first character most significant) and needs
 * nfields >= 1 + 5 + ... + 5^(order+1) entries; exact 64-bit counts (the reference's float counters saturate at 2^24, which
 * xxm_null_model_from_counts reproduces).  last_ms (may be NULL): device time of the histogram passes.
 * xxm_null_model_from_counts: the rest of NullModel::init (NullModel.cpp:134-201) on the host, float arithmetic in the
 *   reference's order -> the two 2^(3(order+1)) float tables xxm_set_background takes.
 * xxm_null_model_from_seqset = the two together; xxm_background_from_seqset: posBg/negBg and their logs (checkSequenceSet,
 *   backgroundDistribution.cpp:50-95) from the same device counts. */
int xxm_kmer_counts(xxm_ctx* ctx, int32_t order, uint64_t* counts, int64_t nfields);
int xxm_kmer_counts_ms(const xxm_ctx* ctx, float* last_ms);
int xxm_null_model_from_counts(const uint64_t* counts, int32_t order, float alpha, float beta, float* rprobs, float* rconds);
int xxm_null_model_from_seqset(xxm_ctx* ctx, int32_t order, float alpha, float beta, float* rprobs, float* rconds);
int xxm_background_from_seqset(xxm_ctx* ctx, double* bg5, double* bgLog5);

#ifdef __cplusplus
}
#endif
#endif
