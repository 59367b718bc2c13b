/*
 * xxm_oracle.c -- TEST INFRASTRUCTURE ONLY (see xxm_oracle.h).
 *
 * Scalar restatement of the reference's PWM-refinement hot path.  Every
 * function cites the reference file:line it follows (paths relative to
 * /root/reference/src).  Deliberately simple: dense arrays, explicit loops,
 * one stable merge sort.  Built with -ffp-contract=off so that every double
 * operation rounds exactly like the reference's -O3 x86-64 SSE2 build.
 *
 * Known, documented limits (reference behaviour is undefined there):
 *   - gapped motifs with more than 8 match columns whose left part spans more
 *     than 10 positions overflow the `int` id in UngappedKmer.h:21-31;
 *     xo_tables_error() reports 1 for them.
 *   - conservation p-values (multi-species alignments), --ranks and
 *     Global::fixedPosition are not restated (unreachable with FASTA input /
 *     default flags).
 */
#include "xxm_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ===================================================================== */
/* sequences                                                             */
/* ===================================================================== */
struct xo_seqset {
	int32_t nent;
	int32_t max_leng, min_leng;
	int32_t* n;       /* [nent]   (0-based)                        */
	uint8_t** S;      /* [nent]   S[i][1..n] codes 0=N,1=A,2=C,3=G,4=T (Alignment.cpp:164) */
	int32_t* total;   /* [nent+1] */
	double* avgLength;/* [nent+1] */
};

xo_seqset* xo_seqset_new(int32_t nent, const int32_t* n, const uint8_t* flat) {
	xo_seqset* s = (xo_seqset*) calloc(1, sizeof *s);
	s->nent = nent;
	s->n = (int32_t*) malloc(sizeof(int32_t) * (size_t) (nent > 0 ? nent : 1));
	s->S = (uint8_t**) malloc(sizeof(uint8_t*) * (size_t) (nent > 0 ? nent : 1));
	s->total = (int32_t*) calloc((size_t) nent + 1, sizeof(int32_t));
	s->avgLength = (double*) calloc((size_t) nent + 1, sizeof(double));
	s->max_leng = 0;
	s->min_leng = 2147483647;
	size_t off = 0;
	for (int32_t i = 0; i < nent; i++) {
		s->n[i] = n[i];
		s->S[i] = (uint8_t*) calloc((size_t) n[i] + 2, 1);
		memcpy(s->S[i] + 1, flat + off, (size_t) n[i]);
		off += (size_t) n[i];
		if (n[i] > s->max_leng) s->max_leng = n[i];
		if (n[i] < s->min_leng) s->min_leng = n[i];
	}
	xo_seqset_features(s);
	return s;
}

void xo_seqset_free(xo_seqset* s) {
	if (!s) return;
	for (int32_t i = 0; i < s->nent; i++) free(s->S[i]);
	free(s->S); free(s->n); free(s->total); free(s->avgLength); free(s);
}

/* calculateSequenceFeatures, seqset.cpp:21-36 */
void xo_seqset_features(xo_seqset* s) {
	s->total[0] = 0;
	double avgLength = 0;
	for (int32_t i = 1; i <= s->nent; i++) {
		int not_X = 0;
		for (int32_t j = 1; j <= s->n[i - 1]; j++)
			if (s->S[i - 1][j] != 0) not_X++;
		if (not_X != 0) avgLength += log(not_X);
		s->total[i] = s->total[i - 1] + not_X;
		s->avgLength[i] = exp(avgLength / i);
	}
}

/* createRevcomp, seqset.cpp:128-178.  NB the feature recount at :165-177 compares codes against the
 * CHARACTER '0' (48), so every position, N included, is counted. */
void xo_seqset_revcomp(xo_seqset* s) {
	for (int32_t i = 0; i < s->nent; i++) {
		const int32_t n = s->n[i];
		s->n[i] = 2 * n + 1;
		s->S[i] = (uint8_t*) realloc(s->S[i], (size_t) n * 2 + 3);
		s->S[i][n + 1] = 0;
		for (int32_t j = n + 2, k = n; j <= 2 * n + 1; j++, k--) {
			const uint8_t c = s->S[i][k];
			s->S[i][j] = (uint8_t) (c == 0 ? 0 : 5 - c);
		}
		s->S[i][2 * n + 2] = 0;
	}
	s->max_leng = s->max_leng * 2 + 1;
	s->min_leng = s->min_leng * 2 + 1;
	double avgLength = 0;
	s->total[0] = 0;
	for (int32_t i = 1; i <= s->nent; i++) {
		int not_X = 0;
		for (int32_t j = 1; j <= s->n[i - 1]; j++)
			if (s->S[i - 1][j] != '0') not_X++;
		if (not_X != 0) avgLength += log(not_X);
		s->total[i] = s->total[i - 1] + not_X;
		s->avgLength[i] = exp(avgLength / i);
	}
}

int32_t xo_seqset_nent(const xo_seqset* s) { return s->nent; }
int32_t xo_seqset_max_leng(const xo_seqset* s) { return s->max_leng; }
int32_t xo_seqset_min_leng(const xo_seqset* s) { return s->min_leng; }
const int32_t* xo_seqset_n(const xo_seqset* s) { return s->n; }
const int32_t* xo_seqset_total(const xo_seqset* s) { return s->total; }
const double* xo_seqset_avgLength(const xo_seqset* s) { return s->avgLength; }
int64_t xo_seqset_flat(const xo_seqset* s, uint8_t* out) {
	int64_t off = 0;
	for (int32_t i = 0; i < s->nent; i++) {
		if (out) memcpy(out + off, s->S[i] + 1, (size_t) s->n[i]);
		off += s->n[i];
	}
	return off;
}

/* ===================================================================== */
/* elongation phase: gapped k-mer counts          */
/* ===================================================================== */
/* the counting loop of ElongCandidates::countAllOccurrences, elongationPhase/elongationCandidates.cpp:234-262: character l of
 * the k-mer sits at j + offset[l] (offset[l] = l + gap[l-1], :251); a start position counts if none of its coding characters
 * is an N (:241-258); seqLen = n - coding - gap[coding-2] + 1 (:244), halved for PALINDROME under --revcomp (:245).
 * counts[id]: first character most significant, 2 bits per character (the reference keeps 4 bits per character). */
void xo_count_gapped_kmers(const xo_seqset* s, int32_t nseq, int32_t ncoding, const int32_t* offset, int32_t firstHalfOnly, uint32_t* counts) {
	memset(counts, 0, sizeof(uint32_t) << (2 * ncoding));
	for (int32_t i = 0; i < nseq; i++) {
		const uint8_t* seq = s->S[i];
		int32_t seqLen = s->n[i] - offset[ncoding - 1];
		if (firstHalfOnly) seqLen /= 2;
		for (int32_t j = 1; j <= seqLen; j++) {
			uint32_t id = 0;
			int ok = 1;
			for (int32_t l = 0; l < ncoding; l++) {
				const int base = (int) seq[j + offset[l]] - 1;
				if (base < 0) { ok = 0; break; }
				id = (id << 2) + (uint32_t) base;
			}
			if (ok) counts[id]++;
		}
	}
}

/* ===================================================================== */
/* background                                                            */
/* ===================================================================== */
/* checkSequenceSet, backgroundDistribution.cpp:30-95 (mono-nucleotide part only) */
void xo_background(const xo_seqset* s, double* bg, double* bgLog) {
	double tot = 0;
	for (int a = 0; a < 5; a++) { bg[a] = 0; bgLog[a] = 0; }
	for (int32_t i = 0; i < s->nent; i++) {
		const int32_t len = s->n[i];
		tot += len;
		bg[s->S[i][1]]++;
		bg[s->S[i][2]]++;
		for (int32_t j = 3; j <= len; j++) bg[s->S[i][j]]++;
	}
	for (int a = 1; a < 5; a++) {
		const double v = bg[a] / (tot - bg[0]);
		bg[a] = v > 1e-100 ? v : 1e-100;
		bgLog[a] = log(bg[a]);
	}
}

struct xo_nullmodel {
	int32_t order;
	int32_t bits;    /* 3 */
	int32_t rfields; /* 2^(3*(order+1)) */
	float* rprobs;
	float* rconds;
};

typedef struct {
	int order, asize, fields;
	float factor;
	int* coeffs;
	float *counts, *conds, *probs;
	float *rprobs, *rconds;
	int bits;
} nm_build;

/* NullModel::sub2ind(kmer, order, byrow=true), NullModel.h:136-153 */
static int nm_sub2ind(const nm_build* b, const uint8_t* seq, int order) {
	int i = 0;
	for (int k = 0, l = order; k <= order; ++k, --l) i += b->coeffs[l] * seq[k];
	return i;
}
/* NullModel::sub2ind2, NullModel.h:155-166 */
static int nm_sub2ind2(const nm_build* b, const uint8_t* seq, int order) {
	int res = seq[0];
	for (int i = 1; i <= order; ++i) { res <<= b->bits; res += seq[i]; }
	return res;
}
/* NullModel::calculateKmerProbs(kmer, order_minus_1), NullModel.cpp:258-295 */
static void nm_kmer_probs_leaf(nm_build* b, uint8_t* kmer, int order_minus_1) {
	int index[5], iindex[5];
	float S = 0;
	const int order = order_minus_1 + 1;
	for (uint8_t k = 1; k <= b->asize; k++) {
		kmer[order] = k;
		const int ii = nm_sub2ind(b, kmer, order);
		iindex[k] = ii;
		const int i = nm_sub2ind(b, kmer, order - 1);
		index[k] = i;
		const int p = nm_sub2ind(b, kmer + 1, order - 1);
		const float conds_ii = (b->counts[ii] + b->factor * b->conds[p]) / (b->counts[i] + b->factor);
		S += conds_ii;
		b->conds[ii] = conds_ii;
	}
	for (uint8_t k = 1; k <= b->asize; k++) {
		const int i = index[k], ii = iindex[k];
		b->conds[ii] /= S;
		b->probs[ii] = b->conds[ii] * b->probs[i];
	}
}
/* NullModel::calculateKmerProbs(kmer, pos, order_minus_1), NullModel.cpp:245-256 */
static void nm_kmer_probs(nm_build* b, uint8_t* kmer, int pos, int order_minus_1) {
	for (uint8_t k = 1; k <= b->asize; k++) {
		kmer[pos] = k;
		if (pos == order_minus_1) nm_kmer_probs_leaf(b, kmer, order_minus_1);
		else nm_kmer_probs(b, kmer, pos + 1, order_minus_1);
	}
}
/* NullModel::calculateGappedKmerProbs(kmer, order, lastGap), NullModel.cpp:222-242 */
static void nm_gapped_leaf(nm_build* b, uint8_t* kmer, int order, int lastGap) {
	const uint8_t N = (uint8_t) (b->asize + 1);
	const int i = nm_sub2ind(b, kmer, order);
	if (lastGap > -1) b->conds[i] = b->probs[i] / b->probs[nm_sub2ind(b, kmer, order - 1)];
	for (int k = lastGap + 1; k <= order; k++) {
		const uint8_t X = kmer[k];
		kmer[k] = N;
		const int ii = nm_sub2ind(b, kmer, order);
		b->probs[ii] += b->probs[i];
		kmer[k] = X;
	}
}
/* NullModel::calculateGappedKmerProbs(kmer, pos, order, lastGap), NullModel.cpp:204-220 */
static void nm_gapped(nm_build* b, uint8_t* kmer, int pos, int order, int lastGap) {
	if (pos > order) { nm_gapped_leaf(b, kmer, order, lastGap); return; }
	uint8_t k;
	for (k = 1; k <= b->asize; k++) {
		kmer[pos] = k;
		nm_gapped(b, kmer, pos + 1, order, lastGap);
	}
	kmer[pos] = k;
	nm_gapped(b, kmer, pos + 1, order, pos);
}
/* NullModel::copyProbs, NullModel.cpp:297-326 */
static void nm_copy(nm_build* b, uint8_t* kmer, uint8_t* new_kmer, int pos, int order) {
	if (pos > order) {
		const int i = nm_sub2ind(b, kmer, order);
		const int new_i = nm_sub2ind2(b, new_kmer, order);
		b->rprobs[new_i] = b->probs[i];
		b->rconds[new_i] = b->conds[i];
		return;
	}
	uint8_t k;
	for (k = 1; k <= b->asize; k++) {
		kmer[pos] = k;
		new_kmer[pos] = k;
		nm_copy(b, kmer, new_kmer, pos + 1, order);
	}
	kmer[pos] = k;
	new_kmer[pos] = 0;
	nm_copy(b, kmer, new_kmer, pos + 1, order);
}

/* NullModel::init, NullModel.cpp:64-201 (gaps = true, freqs = NULL as passed at backgroundDistribution.cpp:12-14) */
xo_nullmodel* xo_nullmodel_new(const xo_seqset* s, float alpha, float beta, int32_t order) {
	nm_build b;
	b.order = order;
	b.asize = 4;
	b.factor = (float) b.asize * alpha;
	b.bits = (int) ceil(log(b.asize + 1) / log(2));
	const int base = b.asize + 1;
	b.fields = 1 + base;
	for (int i = 2, powBase = base; i < order + 2; i++) { powBase *= base; b.fields += powBase; }
	b.counts = (float*) calloc((size_t) b.fields, sizeof(float));
	b.conds = (float*) calloc((size_t) b.fields, sizeof(float));
	b.probs = (float*) calloc((size_t) b.fields, sizeof(float));
	b.coeffs = (int*) calloc((size_t) order + 2, sizeof(int));
	for (int k = 0, powBase = 1; k < order + 2; k++) { b.coeffs[k] = powBase; powBase *= base; }
	const int rfields = (int) pow(2, (order + 1) * b.bits);
	b.rconds = (float*) calloc((size_t) rfields, sizeof(float));
	b.rprobs = (float*) calloc((size_t) rfields, sizeof(float));
	uint8_t* kmer = (uint8_t*) calloc((size_t) order + 2, 1);
	uint8_t* new_kmer = (uint8_t*) calloc((size_t) order + 2, 1);

	/* counts, :122-132 (float increments, exactly like the reference) */
	for (int32_t n = 0; n < s->nent; n++) {
		const int L = s->n[n];
		const uint8_t* sq = s->S[n];
		for (int l = 1; l <= L; l++)
			for (int k = 0; k <= order && l + k <= L && sq[l + k]; k++) b.counts[nm_sub2ind(&b, sq + l, k)]++;
	}
	if (beta > 0)
		for (int i = 1; i < b.fields; i++) { float v = (float) (b.counts[i] - beta); b.counts[i] = v > 0.0f ? v : 0.0f; }
	float ncounts = 0;
	for (int i = 1; i <= b.asize; i++) ncounts += b.counts[i];
	for (int n = 1; n <= b.asize; n++) b.conds[n] = b.probs[n] = b.counts[n] / ncounts; /* :161-166 */
	for (int o = 1; o <= order; o++) nm_kmer_probs(&b, kmer, 0, o - 1);                 /* :170-172 */
	b.conds[b.asize + 1] = b.probs[b.asize + 1] = 1;                                    /* :182 */
	for (int o = 1; o <= order; o++) nm_gapped(&b, kmer, 0, o, -1);                     /* :185-186 */
	for (int o = order; o >= 0; o--) nm_copy(&b, kmer, new_kmer, 0, o);                 /* :194-196 */

	xo_nullmodel* nm = (xo_nullmodel*) calloc(1, sizeof *nm);
	nm->order = order; nm->bits = b.bits; nm->rfields = rfields; nm->rprobs = b.rprobs; nm->rconds = b.rconds;
	free(b.counts); free(b.conds); free(b.probs); free(b.coeffs); free(kmer); free(new_kmer);
	return nm;
}

xo_nullmodel* xo_nullmodel_from_tables(int32_t order, const float* rprobs, const float* rconds) {
	xo_nullmodel* nm = (xo_nullmodel*) calloc(1, sizeof *nm);
	nm->order = order; nm->bits = 3; nm->rfields = 1 << (3 * (order + 1));
	nm->rprobs = (float*) malloc(sizeof(float) * (size_t) nm->rfields);
	nm->rconds = (float*) malloc(sizeof(float) * (size_t) nm->rfields);
	memcpy(nm->rprobs, rprobs, sizeof(float) * (size_t) nm->rfields);
	memcpy(nm->rconds, rconds, sizeof(float) * (size_t) nm->rfields);
	return nm;
}
void xo_nullmodel_free(xo_nullmodel* nm) { if (nm) { free(nm->rprobs); free(nm->rconds); free(nm); } }
int32_t xo_nullmodel_rfields(const xo_nullmodel* nm) { return nm->rfields; }
const float* xo_nullmodel_rprobs(const xo_nullmodel* nm) { return nm->rprobs; }
const float* xo_nullmodel_rconds(const xo_nullmodel* nm) { return nm->rconds; }

/* UngappedKmer(const unsigned char*, len), UngappedKmer.h:21-31 : 3 bits per char, first char most significant */
static int ukmer(const uint8_t* s, int len) {
	unsigned int res = s[0];
	for (int i = 1; i < len; ++i) { res <<= 3; res += s[i]; }
	return (int) res;
}
/* NullModel::getProbability(kmer, length), NullModel.cpp:42-54 */
static double nm_probability(const xo_nullmodel* nm, const uint8_t* kmer, int length) {
	double p;
	const int o = nm->order;
	if (length <= o + 1) {
		p = nm->rprobs[ukmer(kmer, length)];
	} else {
		p = nm->rprobs[ukmer(kmer, o + 1)];
		for (int stop = o + 1; stop < length; ++stop) p *= nm->rconds[ukmer(kmer + stop - o, o + 1)];
	}
	return p;
}

/* ===================================================================== */
/* p-value tables: Kmer_Generator_Reloaded<DnaHash>                       */
/* ===================================================================== */
typedef struct { int score; int64_t id; } leaf_t;

struct xo_tables {
	int error;
	int L;          /* totalLength = number of match columns */
	int maxk;       /* MAX_KMER_SIZE */
	int nparts;     /* hashNumber */
	int posSetSize, totalPositions;
	int gaps[XO_PWM_LENGTH + 2];
	double PWM[XO_PWM_LENGTH + 2][5];     /* log-odds, 1-based match column index */
	double PWM_lin[XO_PWM_LENGTH + 2][5];
	int64_t size[2];
	double* pvalue[2];  /* pValueHash[part] (dense) */
	double* prob;       /* probabilityHash */
	int nclasses[2];
	int32_t* cscore[2];
	int32_t* ccount[2];
	int64_t* cids[2];
	int64_t nids[2];
	double* cleftprob;  /* per left class: sum of probabilityHash over its ids, in list order (:442-445) */
	double* cp2;        /* per right class: pValueHash[1][first id of the class] (:437) */
};

typedef struct { int id; uint32_t klen; } pk_t; /* prefix kmer: UngappedKmer id + length */

typedef struct {
	xo_tables* t;
	const xo_nullmodel* nm;
	/* state of one set_kmer_to_Pvalue_hash call */
	int start, stop, length, part, gapOffset, T, fill;
	int rowSwap[XO_PWM_LENGTH + 2][5];
	int colSwap[XO_PWM_LENGTH + 2];
	double swapped[XO_PWM_LENGTH + 2][5];
	double S_max[XO_PWM_LENGTH + 3];
	uint8_t idx[XO_PWM_LENGTH + 2];
	leaf_t* leaves; size_t nleaves, capleaves;
	/* prefix list (branch_and_bound.h:31, filled at :77-79) */
	pk_t* prefixes; float* prefprob; size_t npref, cappref;
	/* rightKmerCond cache, keyed by the integer id only (:305, :321) */
	int* ck; float* cv; size_t ccap, cn;
} tb_state;

static void tb_push_leaf(tb_state* st, int score, int64_t id) {
	if (st->nleaves == st->capleaves) {
		st->capleaves = st->capleaves ? st->capleaves * 2 : 1024;
		st->leaves = (leaf_t*) realloc(st->leaves, st->capleaves * sizeof(leaf_t));
	}
	st->leaves[st->nleaves].score = score;
	st->leaves[st->nleaves].id = id;
	st->nleaves++;
}

/* create_similar_kmers_rec, branch_and_bound-inl.h:586-603 */
static void tb_rec(tb_state* st, int j, double S_j) {
	for (int a = 1; a <= 4; a++) {
		const double S_j_1 = S_j + st->swapped[j][a];
		if ((int) ((S_j_1 + st->S_max[j + 1]) * 1e6) < st->T) continue;
		st->idx[st->colSwap[j] - 1] = (uint8_t) st->rowSwap[st->colSwap[j]][a];
		if (j + 1 <= st->length) {
			tb_rec(st, j + 1, S_j_1);
		} else {
			const int intScore = (int) (S_j_1 * 1e6);
			int64_t id = 0;
			for (int pos = st->length - 1; pos >= 0; pos--) id = (id << 2) + st->idx[pos] - 1;
			tb_push_leaf(st, intScore, id);
		}
	}
}

static void cache_clear(tb_state* st) {
	st->ccap = 4096; st->cn = 0;
	free(st->ck); free(st->cv);
	st->ck = (int*) malloc(sizeof(int) * st->ccap);
	st->cv = (float*) malloc(sizeof(float) * st->ccap);
	for (size_t i = 0; i < st->ccap; i++) st->ck[i] = -1;
}
static int cache_find(tb_state* st, int key, float* v) {
	size_t h = ((size_t) (unsigned) key * 2654435761u) & (st->ccap - 1);
	while (st->ck[h] != -1) {
		if (st->ck[h] == key) { *v = st->cv[h]; return 1; }
		h = (h + 1) & (st->ccap - 1);
	}
	return 0;
}
static void cache_put(tb_state* st, int key, float v) {
	if ((st->cn + 1) * 2 > st->ccap) { /* grow */
		size_t oc = st->ccap; int* ok = st->ck; float* ov = st->cv;
		st->ccap *= 2;
		st->ck = (int*) malloc(sizeof(int) * st->ccap);
		st->cv = (float*) malloc(sizeof(float) * st->ccap);
		for (size_t i = 0; i < st->ccap; i++) st->ck[i] = -1;
		st->cn = 0;
		for (size_t i = 0; i < oc; i++) if (ok[i] != -1) cache_put(st, ok[i], ov[i]);
		free(ok); free(ov);
	}
	size_t h = ((size_t) (unsigned) key * 2654435761u) & (st->ccap - 1);
	while (st->ck[h] != -1) h = (h + 1) & (st->ccap - 1);
	st->ck[h] = key; st->cv[h] = v; st->cn++;
}

/* getRightKmerConditional, branch_and_bound-inl.h:303-323 ; klen = k.length() */
static float tb_right_cond(tb_state* st, int kid, int klen) {
	float cached;
	if (cache_find(st, kid, &cached)) return cached;
	double condProb = 0;
	const int o = st->nm->order;
	const int start = st->t->maxk - o + klen - 1;
	const int stop = st->t->maxk - 1;
	const int lastPos = stop - start + 1;
	const int bits = (int) ((1u << (lastPos * 3)) - 1u); /* UngappedKmer::getLastPosCombined, UngappedKmer.h:163-166 */
	for (size_t i = 0; i < st->npref; i++) {
		const uint64_t idx = (((uint64_t) (int64_t) st->prefixes[i].id & (uint64_t) (int64_t) bits) << (klen * 3)) + (uint64_t) (int64_t) kid;
		const double c = st->nm->rconds[idx];
		if (c > condProb) condProb = c;
	}
	cache_put(st, kid, (float) condProb);
	return (float) condProb;
}

/* calculate_kmer_probability, branch_and_bound-inl.h:54-100 */
static double tb_kmer_probability(tb_state* st, int64_t kmer_id) {
	uint8_t ex[128] = {0};
	int64_t id = kmer_id;
	int totalPos = 0;
	for (int matchPos = 0; matchPos < st->length; ++matchPos) {
		ex[totalPos++] = (uint8_t) ((id & 3) + 1);
		id >>= 2;
		for (int i = 0; i < st->t->gaps[matchPos + st->gapOffset]; ++i) ex[totalPos++] = 0;
	}
	double prob;
	if (st->part == 0) {
		prob = nm_probability(st->nm, ex, totalPos);
		if (st->fill) {
			if (totalPos > 10) st->t->error = 1; /* int overflow in the reference's UngappedKmer id */
			if (st->npref == st->cappref) {
				st->cappref = st->cappref ? st->cappref * 2 : 1024;
				st->prefixes = (pk_t*) realloc(st->prefixes, st->cappref * sizeof(pk_t));
				st->prefprob = (float*) realloc(st->prefprob, st->cappref * sizeof(float));
			}
			st->prefixes[st->npref].id = ukmer(ex, totalPos);
			st->prefixes[st->npref].klen = (uint32_t) totalPos;
			st->prefprob[st->npref] = 1.0f;
			st->npref++;
		}
	} else {
		const int o = st->nm->order;
		prob = 1;
		for (int stop = 0; stop < totalPos; ++stop) {
			const int start = stop - o > 0 ? stop - o : 0;
			const int len = stop - start + 1;
			const int k = ukmer(ex + start, len);
			if (len <= o) {
				const double rightProb = tb_right_cond(st, k, len);
				prob *= rightProb;
			} else {
				prob *= st->nm->rconds[k];
			}
		}
	}
	return prob;
}

/* recalculatePrefixProbabilities, branch_and_bound-inl.h:260-301 */
static void tb_recalc_prefixes(tb_state* st) {
	float max = 0;
	for (size_t i = 0; i < st->npref; i++) {
		int col = 1;
		float probability = 1;
		const int klen = (int) st->prefixes[i].klen;
		for (int c = 0; c < klen; ++c) {
			const int shift = (klen - 1 - c) * 3; /* UngappedKmer::charAt, UngappedKmer.h:55-59 */
			const int chr = (st->prefixes[i].id >> shift) & 7;
			if (chr != 0) {
				probability = (float) (probability * st->t->PWM_lin[col][chr]);
				++col;
			}
		}
		st->prefprob[i] = probability;
		if (probability > max) max = probability;
	}
	size_t w = 0;
	for (size_t i = 0; i < st->npref; i++) {
		if (st->prefprob[i] < 0.8 * max) continue; /* erased */
		st->prefixes[w] = st->prefixes[i];
		st->prefprob[w] = st->prefprob[i];
		w++;
	}
	st->npref = w;
	cache_clear(st);
}

/* stable merge sort of leaves by score DESCENDING (std::map reverse iteration keeps, inside one key, push order) */
static void leaf_sort(leaf_t* a, size_t n, leaf_t* tmp) {
	if (n < 2) return;
	size_t h = n / 2;
	leaf_sort(a, h, tmp);
	leaf_sort(a + h, n - h, tmp);
	size_t i = 0, j = h, k = 0;
	while (i < h && j < n) tmp[k++] = (a[j].score > a[i].score) ? a[j++] : a[i++];
	while (i < h) tmp[k++] = a[i++];
	while (j < n) tmp[k++] = a[j++];
	memcpy(a, tmp, n * sizeof(leaf_t));
}

/* set_kmer_to_Pvalue_hash, branch_and_bound-inl.h:462-577 */
static void tb_set_kmer_to_pvalue(tb_state* st, int start, int stop) {
	xo_tables* t = st->t;
	st->start = start; st->stop = stop;
	st->length = stop - start + 1;
	const int length = st->length;
	memset(st->rowSwap, 0, sizeof st->rowSwap);
	memset(st->colSwap, 0, sizeof st->colSwap);
	memset(st->swapped, 0, sizeof st->swapped);
	memset(st->S_max, 0, sizeof st->S_max);
	memset(st->idx, 0, sizeof st->idx);
	/* rowSwap, :477-490 */
	for (int i = start; i <= stop; i++) {
		int* r = st->rowSwap[i - start + 1];
		for (int j = 1; j <= 4; j++) {
			if (j == 1) r[j] = 1;
			else {
				int swapIndex = j;
				while (swapIndex > 1 && t->PWM[i][r[swapIndex - 1]] < t->PWM[i][j]) {
					r[swapIndex] = r[swapIndex - 1];
					swapIndex--;
				}
				r[swapIndex] = j;
			}
		}
	}
	/* colSwap, :496-509.  NB the left operand indexes PWM[] with the RELATIVE column number (not +start-1),
	 * exactly as the reference does; for the right part this compares against left-part rows. */
	for (int i = start; i <= stop; i++) {
		if (i == start) st->colSwap[i - start + 1] = 1;
		else {
			int swapIndex = i - start + 1;
			while (swapIndex > 1 &&
			       t->PWM[st->colSwap[swapIndex - 1]][st->rowSwap[st->colSwap[swapIndex - 1]][1]] <
			       t->PWM[i][st->rowSwap[i - start + 1][1]]) {
				st->colSwap[swapIndex] = st->colSwap[swapIndex - 1];
				swapIndex--;
			}
			st->colSwap[swapIndex] = i - start + 1;
		}
	}
	/* swappedPWM, :520-527 */
	for (int i = start; i <= stop; i++)
		for (int j = 1; j <= 4; j++)
			st->swapped[i - start + 1][j] = t->PWM[st->colSwap[i - start + 1] + start - 1][st->rowSwap[st->colSwap[i - start + 1]][j]];
	/* S_max, :532-536 */
	for (int i = length; i > 0; i--) {
		if (i < length) st->S_max[i] = st->S_max[i + 1] + st->swapped[i][1];
		else st->S_max[i] = st->swapped[i][1];
	}
	st->nleaves = 0;
	tb_rec(st, 1, 0);

	/* classes in descending score, ids in push order, :545-565 */
	leaf_t* tmp = (leaf_t*) malloc((st->nleaves ? st->nleaves : 1) * sizeof(leaf_t));
	leaf_sort(st->leaves, st->nleaves, tmp);
	free(tmp);
	const int part = st->part;
	t->size[part] = (int64_t) 1 << (2 * length);
	t->pvalue[part] = (double*) calloc((size_t) t->size[part], sizeof(double));
	if (st->fill) t->prob = (double*) calloc((size_t) t->size[part], sizeof(double));
	t->cids[part] = (int64_t*) malloc((st->nleaves ? st->nleaves : 1) * sizeof(int64_t));
	t->cscore[part] = (int32_t*) malloc((st->nleaves ? st->nleaves : 1) * sizeof(int32_t));
	t->ccount[part] = (int32_t*) malloc((st->nleaves ? st->nleaves : 1) * sizeof(int32_t));
	t->nids[part] = (int64_t) st->nleaves;
	int nc = 0;
	double pValue = 0;
	size_t i = 0;
	while (i < st->nleaves) {
		size_t j = i;
		while (j < st->nleaves && st->leaves[j].score == st->leaves[i].score) j++;
		for (size_t k = i; k < j; k++) {
			const double single = tb_kmer_probability(st, st->leaves[k].id);
			if (st->fill) t->prob[st->leaves[k].id] = single;
			pValue += single;
		}
		for (size_t k = i; k < j; k++) {
			t->pvalue[part][st->leaves[k].id] = pValue;
			t->cids[part][k] = st->leaves[k].id;
		}
		t->cscore[part][nc] = st->leaves[i].score;
		t->ccount[part][nc] = (int32_t) (j - i);
		nc++;
		i = j;
	}
	t->nclasses[part] = nc;
}

/* Kmer_Generator_Reloaded::reinitialize(pwm, cols, posSetSize, totalPositions), branch_and_bound-inl.h:183-258 */
xo_tables* xo_tables_new(const double* pwm, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t totalPositions,
                         const double* bgLog, const xo_nullmodel* nm) {
	xo_tables* t = (xo_tables*) calloc(1, sizeof *t);
	t->posSetSize = posSetSize;
	t->totalPositions = totalPositions;
	t->L = ncols;
	t->maxk = 8;
	if (t->L > 17) t->maxk = t->L / 2;
	t->nparts = (t->L <= t->maxk) ? 1 : 2;
	for (int i = 1; i < ncols; i++) t->gaps[i - 1] = cols[i] - cols[i - 1] - 1;
	t->gaps[ncols - 1] = 0;
	for (int i = 1; i <= ncols; ++i) {
		double lin_sum = 0;
		for (int j = 1; j <= 4; ++j) {
			t->PWM[i][j] = pwm[cols[i - 1] * 5 + j] - bgLog[j];
			t->PWM_lin[i][j] = exp(pwm[cols[i - 1] * 5 + j]);
			lin_sum += t->PWM_lin[i][j];
		}
		if (fabs(lin_sum - 1) >= 1e-3) t->error = 2; /* the reference asserts here (:223) */
	}
	tb_state* st = (tb_state*) calloc(1, sizeof *st);
	st->t = t; st->nm = nm;
	cache_clear(st);
	if (t->nparts == 1) {
		st->part = 0; st->fill = 0; st->gapOffset = 0; st->T = 0;
		tb_set_kmer_to_pvalue(st, 1, t->L);
	} else {
		st->part = 0; st->fill = 1; st->gapOffset = 0; st->T = 0;
		st->npref = 0;
		tb_set_kmer_to_pvalue(st, 1, t->maxk);
		st->part = 1; st->fill = 0;
		tb_recalc_prefixes(st);
		const int start = t->maxk + 1, stop = t->L;
		const int Tr = (int) ((-t->maxk + stop - start + 1) * 1e6);
		st->T = Tr < 0 ? Tr : 0;
		st->gapOffset = t->maxk;
		tb_set_kmer_to_pvalue(st, start, stop);
		/* per-class left probabilities, summed in list order like :442-445 */
		t->cleftprob = (double*) calloc((size_t) (t->nclasses[0] ? t->nclasses[0] : 1), sizeof(double));
		int64_t off = 0;
		for (int c = 0; c < t->nclasses[0]; c++) {
			double lp = 0;
			for (int k = 0; k < t->ccount[0][c]; k++) lp += t->prob[t->cids[0][off + k]];
			t->cleftprob[c] = lp;
			off += t->ccount[0][c];
		}
		t->cp2 = (double*) calloc((size_t) (t->nclasses[1] ? t->nclasses[1] : 1), sizeof(double));
		off = 0;
		for (int c = 0; c < t->nclasses[1]; c++) {
			t->cp2[c] = t->pvalue[1][t->cids[1][off]];
			off += t->ccount[1][c];
		}
	}
	free(st->leaves); free(st->prefixes); free(st->prefprob); free(st->ck); free(st->cv); free(st);
	return t;
}

void xo_tables_free(xo_tables* t) {
	if (!t) return;
	for (int p = 0; p < 2; p++) { free(t->pvalue[p]); free(t->cscore[p]); free(t->ccount[p]); free(t->cids[p]); }
	free(t->prob); free(t->cleftprob); free(t->cp2); free(t);
}
int32_t xo_tables_error(const xo_tables* t) { return t->error; }
int32_t xo_tables_L(const xo_tables* t) { return t->L; }
int32_t xo_tables_maxk(const xo_tables* t) { return t->maxk; }
int32_t xo_tables_nparts(const xo_tables* t) { return t->nparts; }
int64_t xo_tables_size(const xo_tables* t, int32_t p) { return t->size[p]; }
const double* xo_tables_pvalues(const xo_tables* t, int32_t p) { return t->pvalue[p]; }
const double* xo_tables_probs(const xo_tables* t) { return t->prob; }
int32_t xo_tables_nclasses(const xo_tables* t, int32_t p) { return t->nclasses[p]; }
const int32_t* xo_tables_class_score(const xo_tables* t, int32_t p) { return t->cscore[p]; }
const int32_t* xo_tables_class_count(const xo_tables* t, int32_t p) { return t->ccount[p]; }
const int64_t* xo_tables_class_ids(const xo_tables* t, int32_t p) { return t->cids[p]; }

/* binarySearch, branch_and_bound-inl.h:339-357 (array sorted by descending score) */
static int tb_binary_search(const int32_t* score, int n, int S) {
	int low = 0, high = n;
	while (low < high) {
		const int mid = low + ((high - low) / 2);
		if (score[mid] > S) low = mid + 1;
		else high = mid;
	}
	if (low < n && score[low] == S) return low;
	return low - 1;
}

/* get_kmer_probability_matched_only, branch_and_bound-inl.h:382-457 */
static double tb_lookup_matched(const xo_tables* t, const uint8_t* kmer) {
	double pValue = 0;
	int64_t id = 0;
	if (t->L <= t->maxk) {
		for (int i = t->L - 1; i >= 0; i--) {
			if (kmer[i] == 0) return 1.0;
			id = (id << 2) + kmer[i] - 1;
		}
		if (t->pvalue[0][id] == 0) return 1.0;
		return t->pvalue[0][id];
	}
	for (int i = t->maxk - 1; i >= 0; i--) {
		if (kmer[i] == 0) return 1.0;
		id = (id << 2) + kmer[i] - 1;
	}
	if (t->pvalue[0][id] == 0) return 1.0;
	const double pVal_0 = t->pvalue[0][id];
	id = 0;
	for (int i = t->L - 1; i >= t->maxk; i--) {
		if (kmer[i] == 0) return 1.0;
		id = (id << 2) + kmer[i] - 1;
	}
	if (t->pvalue[1][id] == 0) return 1.0;
	const double pVal_1 = t->pvalue[1][id];
	if (pVal_0 * pVal_1 > t->posSetSize / (t->totalPositions * 2.0)) return 1.0;

	/* get_Kmer_score, :42-50 */
	double score = 0;
	for (int i = 1; i <= t->L; i++) score += t->PWM[i][kmer[i - 1]];
	const int s = (int) (score * 1e6 - 1);
	const int s1_max = t->cscore[0][0];
	const int s2_max = t->cscore[1][0];
	const int vector0_size = t->nclasses[0];
	int it_s_s1 = tb_binary_search(t->cscore[1], t->nclasses[1], s - s1_max);
	for (int it_s1 = 0; it_s1 < vector0_size; it_s1++) {
		const int s1 = t->cscore[0][it_s1];
		if (s - s2_max > s1) break;
		while (it_s_s1 != 0 && t->cscore[1][it_s_s1] < s - s1) it_s_s1--;
		const double pVal_s_s1_max = t->cp2[it_s_s1];
		pValue += t->cleftprob[it_s1] * pVal_s_s1_max;
	}
	return pValue;
}

/* get_kmer_probability_all_positions, branch_and_bound-inl.h:361-379 */
double xo_tables_lookup_window(const xo_tables* t, const uint8_t* window) {
	uint8_t stripped[64];
	int matchNum = 0, pos = 0;
	while (matchNum < t->L) {
		stripped[matchNum] = window[pos];
		pos += 1 + t->gaps[matchNum];
		++matchNum;
	}
	return tb_lookup_matched(t, stripped);
}

/* ===================================================================== */
/* scan / select                                                          */
/* ===================================================================== */
struct xo_ctx {
	const xo_seqset* set;
	xo_config cfg;
	/* LogTable, LogTable.cpp:14-32 */
	int logsize;
	double* LOG_i;
	double* LOG_sum_i;
	/* PVal_Calculator, pValCalculation.cpp:3-27 */
	double* positionalProb;
	double* sum_i;
	double overrep[XO_PWM_LENGTH];
	/* StartPosUpdater buffers, StartPosUpdater.cpp:5-32 */
	xo_site* sortedSites;
	xo_site* sortedSites_elong;
	xo_site* bestStartPos;
	int32_t* motifsPerSequenceCount;
	int nbSites, nbBestSites, maxSites;
	double log_posSetSize;
	xo_site* tmp; /* merge sort scratch */
	int64_t cap;
};

xo_ctx* xo_ctx_new(const xo_seqset* set, const xo_config* cfg) {
	xo_ctx* c = (xo_ctx*) calloc(1, sizeof *c);
	c->set = set;
	c->cfg = *cfg;
	const int total = set->total[set->nent];
	c->logsize = total + 1 > 10000 ? total + 1 : 10000;
	c->LOG_i = (double*) malloc(sizeof(double) * (size_t) c->logsize);
	c->LOG_sum_i = (double*) malloc(sizeof(double) * (size_t) c->logsize);
	c->LOG_i[0] = log(1e-100);
	c->LOG_sum_i[0] = 0;
	for (int i = 1; i < c->logsize; i++) {
		c->LOG_i[i] = log(i);
		c->LOG_sum_i[i] = c->LOG_i[i] + c->LOG_sum_i[i - 1];
	}
	c->positionalProb = (double*) calloc((size_t) set->max_leng + 1, sizeof(double));
	c->sum_i = (double*) calloc((size_t) set->max_leng + 1, sizeof(double));
	for (int i = 0; i <= set->max_leng; i++) {
		c->sum_i[i] = 0;
		for (int j = i + 1; j <= set->max_leng; j++) c->sum_i[i] += 1.0 / j;
	}
	for (int i = 0; i < XO_PWM_LENGTH; i++) c->overrep[i] = pow(cfg->overrepCorrection, i);
	int64_t flat = 0;
	for (int i = 0; i < set->nent; i++) flat += set->n[i];
	c->cap = flat + set->nent + 2;
	c->sortedSites = (xo_site*) calloc((size_t) c->cap, sizeof(xo_site));
	c->sortedSites_elong = (xo_site*) calloc((size_t) c->cap, sizeof(xo_site));
	c->bestStartPos = (xo_site*) calloc((size_t) c->cap, sizeof(xo_site));
	c->tmp = (xo_site*) calloc((size_t) c->cap, sizeof(xo_site));
	c->motifsPerSequenceCount = (int32_t*) calloc((size_t) set->nent + 1, sizeof(int32_t));
	return c;
}
void xo_ctx_free(xo_ctx* c) {
	if (!c) return;
	free(c->LOG_i); free(c->LOG_sum_i); free(c->positionalProb); free(c->sum_i);
	free(c->sortedSites); free(c->sortedSites_elong); free(c->bestStartPos); free(c->tmp);
	free(c->motifsPerSequenceCount); free(c);
}
/* test plumbing without a counterpart in the reference (its singletons exist once): a second context on the SAME set takes over the
 * candidate list of another one -- bestStartPos with its count, and sortedSites, whose sequences fill_startPosList_with_elongSites
 * counts (StartPosUpdater.cpp:128) -- and a list saved earlier is put back as sortedSites_elong */
int64_t xo_ctx_cap(const xo_ctx* c) { return c->cap; }
void xo_ctx_copy_candidates(xo_ctx* dst, const xo_ctx* src) {
	const int64_t n = dst->cap < src->cap ? dst->cap : src->cap;
	memcpy(dst->bestStartPos, src->bestStartPos, (size_t) n * sizeof(xo_site));
	memcpy(dst->sortedSites, src->sortedSites, (size_t) n * sizeof(xo_site));
	dst->nbBestSites = src->nbBestSites;
}
void xo_ctx_set_elong(xo_ctx* c, const xo_site* sites, int64_t n) {
	if (n > c->cap) n = c->cap;
	memcpy(c->sortedSites_elong, sites, (size_t) n * sizeof(xo_site));
}
const xo_site* xo_ctx_best(const xo_ctx* c) { return c->bestStartPos; }
const xo_site* xo_ctx_sites(const xo_ctx* c) { return c->sortedSites; }
const xo_site* xo_ctx_elong(const xo_ctx* c) { return c->sortedSites_elong; }
int32_t xo_ctx_nbSites(const xo_ctx* c) { return c->nbSites; }
int32_t xo_ctx_nbBestSites(const xo_ctx* c) { return c->nbBestSites; }
int32_t xo_ctx_maxSites(const xo_ctx* c) { return c->maxSites; }
const double* xo_ctx_positionalProb(const xo_ctx* c) { return c->positionalProb; }

/* PVal_Calculator::initPositionalProbs, pValCalculation.cpp:195-248 */
void xo_init_positional_probs(xo_ctx* c, int32_t motifLength, int32_t firstMotifColumn, int32_t enrStart, int32_t enrEnd,
                              int32_t offset) {
	const int max_leng = c->set->max_leng;
	int length = max_leng - motifLength + 1;
	if (c->cfg.revcomp) length = (max_leng / 2) - motifLength + 1;
	const int motifStart = enrStart + firstMotifColumn - 1 + offset;
	const int motifEnd = enrEnd + firstMotifColumn - 1 + offset;
	int startRegion = 1 > motifStart ? 1 : motifStart;
	int endRegion = motifEnd < length ? motifEnd : length;
	for (int i = 1; i <= length; i++) {
		int dist = 0;
		if (i < startRegion) dist = startRegion - i;
		else if (i > endRegion) dist = i - endRegion;
		const int mn = 1 > startRegion - dist ? 1 : startRegion - dist;
		const int mx = length < endRegion + dist ? length : endRegion + dist;
		c->positionalProb[i] = (mx - mn + 1.0) / length;
	}
	if (c->cfg.revcomp) {
		const int start = max_leng / 2 + 2;
		const int oldLength = length;
		const int length2 = max_leng - motifLength + 1;
		const int spaceToMiddleStart = max_leng / 2 - (motifStart + motifLength - 1);
		const int spaceToMiddleEnd = max_leng / 2 - (motifEnd + motifLength - 1);
		startRegion = start > start + spaceToMiddleEnd ? start : start + spaceToMiddleEnd;
		endRegion = length2 < start + spaceToMiddleStart ? length2 : start + spaceToMiddleStart;
		for (int i = start; i <= length2; i++) {
			int dist = 0;
			if (i < startRegion) dist = startRegion - i;
			else if (i > endRegion) dist = i - endRegion;
			const int mn = start > startRegion - dist ? start : startRegion - dist;
			const int mx = length2 < endRegion + dist ? length2 : endRegion + dist;
			c->positionalProb[i] = (mx - mn + 1.0) / (oldLength);
		}
	}
}

/* calculateFinalPval, pValCalculation.h:116-120 */
static double final_pval(const xo_ctx* c, int len, int delta_0, double p) {
	const double plateau = pow(1 - p * len / delta_0, delta_0);
	const double hill = exp(-p * len * c->sum_i[delta_0]);
	return 1 - plateau * hill;
}

/* PVal_Calculator::calculatePval / __calculatePval, pValCalculation.h:56-114 */
double xo_calculate_pval(xo_ctx* c, const xo_tables* t, int32_t seq, int32_t pos, int32_t len, int32_t delta_0, int32_t length,
                         double avgLength, int32_t mops, double* pValScoreOut) {
	const uint8_t* window = c->set->S[seq - 1] + pos;
	double finalPval;
	double pValScore = xo_tables_lookup_window(t, window);
	*pValScoreOut = pValScore;
	if (pValScore == 1 || pValScore == -1) return pValScore;
	pValScore *= c->overrep[t->L];
	*pValScoreOut = pValScore;
	if (c->cfg.localization) {
		const double pValPos = c->positionalProb[pos];
		if (mops) finalPval = (pValScore * sqrt(pValPos) - pValScore * pValScore * pValPos * 0.5) * 2;
		else finalPval = final_pval(c, len, delta_0, pValScore * pValPos);
	} else {
		if (mops) finalPval = pValScore;
		else finalPval = 1 - pow(1 - pValScore, avgLength - length + 1);
	}
	return finalPval;
}

/* updateMultOcc, StartPosUtils.h:12-65 */
static int update_mult_occ(const xo_ctx* c, xo_site* sites, int* sites_counter, int32_t seq, int32_t pos, int length, double pVal) {
	if (*sites_counter > 0 && sites[*sites_counter - 1].sequence == seq && sites[*sites_counter - 1].startPos + length >= pos) {
		if (sites[*sites_counter - 1].score <= pVal) return 0;
		else (*sites_counter)--;
	}
	if (c->cfg.revcomp && pos > length / 2) {
		int back = 1;
		const int revPos = c->set->n[seq - 1] - length + 2 - pos;
		while (*sites_counter - back >= 0 && sites[*sites_counter - back].sequence == seq &&
		       sites[*sites_counter - back].startPos > revPos + length - 1) {
			back++;
		}
		if (*sites_counter - back >= 0 && sites[*sites_counter - back].sequence == seq &&
		    sites[*sites_counter - back].startPos >= revPos - length + 1) {
			if (sites[*sites_counter - back].score <= pVal) return 0;
			while (back > 1) {
				sites[*sites_counter - back] = sites[*sites_counter - back + 1];
				back--;
			}
			(*sites_counter)--;
		}
	}
	sites[*sites_counter].sequence = seq;
	sites[*sites_counter].startPos = pos;
	sites[*sites_counter].score = (float) pVal;
	return 1;
}

/* stable merge sort == glibc qsort for these sizes */
typedef int (*site_cmp)(const xo_site*, const xo_site*);
/* SortedSites::compare_scores, Sorted_Sites.h:15-23 */
static int cmp_scores(const xo_site* a, const xo_site* b) {
	const double diff = a->score - b->score;
	if (diff > 0.0) return 1;
	if (diff < 0.0) return -1;
	return 0;
}
/* compare_startPos, Sorted_Sites.h:26-39 */
static int cmp_startpos(const xo_site* a, const xo_site* b) {
	double diff = a->sequence - b->sequence;
	if (diff > 0.0) return 1;
	if (diff < 0.0) return -1;
	diff = a->startPos - b->startPos;
	if (diff > 0.0) return 1;
	if (diff < 0.0) return -1;
	return 0;
}
/* compare_startPos_palin, Sorted_Sites.h:42-55 */
static int cmp_startpos_palin(const xo_site* a, const xo_site* b) {
	double diff = a->sequence - b->sequence;
	if (diff > 0.0) return 1;
	if (diff < 0.0) return -1;
	diff = a->startPos - b->startPos;
	if (diff < 0.0) return 1;
	if (diff > 0.0) return -1;
	return 0;
}
static void msort(xo_site* a, int64_t n, xo_site* tmp, site_cmp cmp) {
	if (n < 2) return;
	const int64_t h = n / 2;
	msort(a, h, tmp, cmp);
	msort(a + h, n - h, tmp, cmp);
	int64_t i = 0, j = h, k = 0;
	while (i < h && j < n) tmp[k++] = (cmp(&a[j], &a[i]) < 0) ? a[j++] : a[i++];
	while (i < h) tmp[k++] = a[i++];
	while (j < n) tmp[k++] = a[j++];
	memcpy(a, tmp, (size_t) n * sizeof(xo_site));
}
void xo_sort_by_score(xo_site* sites, int32_t n) {
	xo_site* tmp = (xo_site*) malloc(sizeof(xo_site) * (size_t) (n > 0 ? n : 1));
	msort(sites, n, tmp, cmp_scores);
	free(tmp);
}
void xo_sort_by_startpos(xo_site* sites, int32_t n, int32_t palin) {
	xo_site* tmp = (xo_site*) malloc(sizeof(xo_site) * (size_t) (n > 0 ? n : 1));
	msort(sites, n, tmp, palin ? cmp_startpos_palin : cmp_startpos);
	free(tmp);
}

static void set_max_sites_mops(xo_ctx* c, int posSetSize, int length) {
	c->maxSites = c->set->total[posSetSize] - posSetSize * (length - 1);
	if (c->cfg.revcomp) c->maxSites = c->set->total[posSetSize] - 2 * posSetSize * (length - 1);
}

/* fill_sites_with_possible_startPositions, StartPosUpdater-inl.h:70-140 */
void xo_scan_full(xo_ctx* c, const xo_tables* t, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t enrStart,
                  int32_t enrEnd, int32_t keepPosSetSize) {
	const int region = enrEnd - enrStart + 1;
	const int length = cols[ncols - 1] - cols[0] + 1;
	if (c->cfg.localization) xo_init_positional_probs(c, length, cols[0], enrStart, enrEnd, 0);
	int sites_counter = 0, bestSites_counter = 0;
	const int sequences = keepPosSetSize ? posSetSize : c->set->nent;
	for (int32_t i = 1; i <= sequences; i++) {
		const int len = c->set->n[i - 1] - length + 1;
		const int delta_0 = len < region ? len : region;
		for (int32_t j = 1; j <= len; j++) {
			double pValScore;
			const double finalPval = xo_calculate_pval(c, t, i, j, len, delta_0, length, c->set->avgLength[posSetSize], 1, &pValScore);
			if (finalPval >= 1) continue;
			c->bestStartPos[bestSites_counter].sequence = i;
			c->bestStartPos[bestSites_counter].startPos = j;
			c->bestStartPos[bestSites_counter].score = (float) pValScore;
			bestSites_counter++;
			if (!update_mult_occ(c, c->sortedSites, &sites_counter, i, j, length, finalPval)) continue;
			sites_counter++;
		}
	}
	c->bestStartPos[bestSites_counter].score = 1;
	while (bestSites_counter > 0 && c->bestStartPos[bestSites_counter - 1].sequence > posSetSize) c->bestStartPos[--bestSites_counter].score = 1;
	c->sortedSites[sites_counter].score = 1;
	while (sites_counter > 0 && c->sortedSites[sites_counter - 1].sequence > posSetSize) c->sortedSites[--sites_counter].score = 1;
	c->nbSites = sites_counter > 0 ? sites_counter : 0;
	c->nbBestSites = bestSites_counter > 0 ? bestSites_counter : 0;
	msort(c->bestStartPos, c->nbBestSites, c->tmp, cmp_scores);
	set_max_sites_mops(c, posSetSize, length);
	c->log_posSetSize = c->LOG_i[c->set->total[posSetSize]];
}

/* utils.h:61-64 */
static double log_binomial(const xo_ctx* c, int N, int k, double prob) {
	return (c->LOG_sum_i[N] - c->LOG_sum_i[k] - c->LOG_sum_i[N - k] + k * log(prob) + (N - k) * log(1 - prob));
}
/* calculateOrderStatisticsPvalue, utils.h:80-102 */
double xo_order_statistics_pvalue(const xo_ctx* c, int32_t motifNb, int32_t N, double p) {
	if (N * p > motifNb || p == 1) return 0;
	double sum_K_1, delta_K_1;
	int K = 0;
	double logN = log_binomial(c, N, motifNb, p);
	double alpha_K_1, sum_K = 0;
	while (K < N - motifNb) {
		alpha_K_1 = (p * (N - (K + motifNb))) / ((1 - p) * ((K + motifNb) + 1));
		sum_K_1 = (1 / alpha_K_1) * (sum_K + 1);
		sum_K = sum_K_1;
		delta_K_1 = log(1 + (1 / sum_K_1));
		if (fabs(logN * 1e-4) > delta_K_1 || logN > 0) break;
		logN += delta_K_1;
		K++;
	}
	return logN;
}

/* getBestSubset, StartPosUpdater.h:135-238 (max_MultSeq == 1, useRankPvalues == false) */
static void best_subset(xo_ctx* c, const xo_site* sites, int nbSites, int posSetSize, double* bestPval, int* bestMotifNb,
                        int* bestCounter, int* bestMotifNbPlus, int* bestCounterPlus) {
	if (nbSites == 0) { *bestMotifNb = 0; return; }
	*bestMotifNb = 1;
	*bestCounter = 1;
	memset(c->motifsPerSequenceCount, 0, ((size_t) c->set->nent + 1) * sizeof(int32_t));
	int motifCounter = 0;
	*bestPval = 1;
	const int64_t lim64 = (int64_t) posSetSize * c->cfg.maxSeqCount;
	const int lim = nbSites < lim64 ? nbSites : (int) lim64;
	for (int i = 0; i < lim; i++) {
		if (sites[i].score == 1) break;
		c->motifsPerSequenceCount[sites[i].sequence]++;
		if (c->motifsPerSequenceCount[sites[i].sequence] > c->cfg.maxMotifsPerSequence) continue;
		motifCounter++;
		if (i + 1 < nbSites && sites[i + 1].score == sites[i].score) continue;
		const double pValue = xo_order_statistics_pvalue(c, motifCounter, c->maxSites, sites[i].score);
		const double plusFrac = 1 + c->cfg.plusFrac;
		if (pValue < *bestPval || c->cfg.oops) {
			*bestPval = pValue;
			*bestMotifNbPlus = *bestMotifNb = motifCounter;
			*bestCounterPlus = *bestCounter = i + 1;
		} else if (plusFrac != 1) {
			if (motifCounter > plusFrac * *bestMotifNb) {
				const int a = *bestMotifNb + 1, b = (int) (*bestMotifNb * plusFrac + 0.5);
				*bestMotifNbPlus = a > b ? a : b;
				const int a2 = *bestCounter + 1, b2 = (int) (*bestCounter * plusFrac + 0.5);
				*bestCounterPlus = a2 > b2 ? a2 : b2;
			} else {
				*bestMotifNbPlus = motifCounter;
				*bestCounterPlus = i + 1;
			}
		}
	}
}

/* fast_log, utils.h:138-162 */
float xo_fast_log(float x) {
	if (x != x) return x;
	static float lg2[1025];
	static float diff[1025];
	static int initialized;
	if (!initialized) {
		float prev = 0.0f;
		lg2[0] = 0.0f;
		for (int i = 1; i <= 1024; ++i) {
			lg2[i] = logf((float) (1024 + i)) * 1.442695041f - 10.0f; /* C++ overload resolution picks log(float) */
			diff[i - 1] = (lg2[i] - prev) * 1.2352E-4f;
			prev = lg2[i];
		}
		initialized = 1;
	}
	int32_t bits;
	memcpy(&bits, &x, 4);
	const int32_t a = ((bits & 0x7F800000) >> 23) - 0x7f;
	const int32_t b = (bits & 0x007FE000) >> 13;
	const int32_t cc = (bits & 0x00001FFF);
	return ((float) a + lg2[b] + diff[b] * (float) (cc)) * 0.69314718f;
}

/* Motif::updatePWM_OOL + updateCounts + buildPWM, Motif.cpp:584-591, :617-672 */
void xo_update_pwm_ool(xo_ctx* c, const xo_site* sites, double* pwm, int32_t firstColumn, int32_t K, int32_t posSetSize, double pseudo) {
	pseudo = pseudo * K; /* getPseudocounts, utils.h:214-217 */
	for (int i = 1; i <= XO_PWM_LENGTH; i++)
		for (int a = 0; a <= 4; a++) pwm[i * 5 + a] = 0;
	memset(c->motifsPerSequenceCount, 0, ((size_t) posSetSize + 1) * sizeof(int32_t));
	for (int i = 0, motifCounter = 0;; i++) {
		const int seq = sites[i].sequence;
		const int pos = sites[i].startPos - firstColumn + 1;
		c->motifsPerSequenceCount[seq]++;
		if (c->motifsPerSequenceCount[seq] > c->cfg.maxMotifsPerSequence) continue;
		const int seqLen = c->set->n[seq - 1];
		const uint8_t* S = c->set->S[seq - 1];
		int tmpPos = pos;
		for (int j = 1; j <= XO_PWM_LENGTH; j++, tmpPos++) {
			if (tmpPos < 1 || tmpPos > seqLen) continue;
			const int base = S[tmpPos];
			if (base == 0) continue;
			pwm[j * 5 + base]++;
			pwm[j * 5 + 0]++;
		}
		if (++motifCounter == K) break;
	}
	for (int i = 1; i <= XO_PWM_LENGTH; i++)
		for (int j = 1; j <= 4; j++)
			pwm[i * 5 + j] = xo_fast_log((float) ((pwm[i * 5 + j] + pseudo * c->cfg.bg[j]) / (pwm[i * 5 + 0] + pseudo)));
}

/* update_PWM_with_bestSubset / update_elongPWM_with_bestSubset, StartPosUpdater.cpp:78-112 */
int32_t xo_update_pwm_with_best_subset(xo_ctx* c, int32_t which, double* pwm, const int32_t* cols, int32_t ncols, int32_t posSetSize,
                                       int32_t oldTotalSites, int32_t plus, double pseudo, int32_t* bestMotifNb, double* bestPval) {
	(void) oldTotalSites; (void) ncols;
	xo_site* sites = which ? c->sortedSites_elong : c->sortedSites;
	if (c->nbSites == 0) return 0;
	msort(sites, c->nbSites, c->tmp, cmp_scores);
	if (sites[0].score == 1) return 0;
	int bestCounter = 1, bestMotifNbPlus = 1, bestCounterPlus = 1;
	best_subset(c, sites, c->nbSites, posSetSize, bestPval, bestMotifNb, &bestCounter, &bestMotifNbPlus, &bestCounterPlus);
	const int motifNb = plus ? bestMotifNbPlus : *bestMotifNb;
	xo_update_pwm_ool(c, sites, pwm, cols[0], motifNb, posSetSize, pseudo);
	return 1;
}

/* fill_elongSites_with_possible_startPositions, StartPosUpdater-inl.h:143-221 */
void xo_rescore(xo_ctx* c, const xo_tables* t, int32_t firstMotifColumn, const int32_t* tmpCols, int32_t ntmp, int32_t posSetSize,
                int32_t enrStart, int32_t enrEnd, int32_t palin) {
	const int region = enrEnd - enrStart + 1;
	const int length = tmpCols[ntmp - 1] - tmpCols[0] + 1;
	if (c->cfg.localization) xo_init_positional_probs(c, length, firstMotifColumn, enrStart, enrEnd, tmpCols[0] - firstMotifColumn);
	if (!c->cfg.mops)
		for (int32_t i = 0; i < posSetSize; i++) c->sortedSites_elong[i].score = 1;
	int sites_counter = 0;
	msort(c->bestStartPos, c->nbBestSites, c->tmp, palin ? cmp_startpos_palin : cmp_startpos);
	for (int i = 0; i < c->nbBestSites; i++) {
		const int32_t seq = c->bestStartPos[i].sequence;
		int32_t pos = c->bestStartPos[i].startPos + tmpCols[0] - firstMotifColumn;
		if (palin) pos = c->set->n[seq - 1] - pos + 1 - length + 1;
		if (pos < 1 || pos + tmpCols[ntmp - 1] - tmpCols[0] > c->set->n[seq - 1]) continue;
		const int len = c->set->n[seq - 1] - length + 1;
		const int delta_0 = len < region ? len : region;
		double pValScore;
		const double finalPval = xo_calculate_pval(c, t, seq, pos, len, delta_0, length, c->set->avgLength[posSetSize], c->cfg.mops, &pValScore);
		if (finalPval >= 1) continue;
		if (c->cfg.mops) {
			if (!update_mult_occ(c, c->sortedSites_elong, &sites_counter, seq, pos, length, finalPval)) continue;
		} else if (finalPval < c->sortedSites_elong[seq - 1].score) {
			c->sortedSites_elong[seq - 1].sequence = seq;
			c->sortedSites_elong[seq - 1].startPos = pos;
			c->sortedSites_elong[seq - 1].score = (float) finalPval;
		}
		sites_counter++;
	}
	if (c->cfg.mops) {
		c->sortedSites_elong[sites_counter].score = 1;
		c->nbSites = sites_counter;
		set_max_sites_mops(c, posSetSize, length);
	} else {
		c->nbSites = posSetSize;
		c->maxSites = posSetSize;
	}
	c->log_posSetSize = c->LOG_i[c->set->total[posSetSize]];
}

/* fill_sites_with_possible_startPositions_oops, StartPosUpdater-inl.h:224-266 */
void xo_scan_oops(xo_ctx* c, const xo_tables* t, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t enrStart, int32_t enrEnd) {
	const int region = enrEnd - enrStart + 1;
	const int length = cols[ncols - 1] - cols[0] + 1;
	if (c->cfg.localization) xo_init_positional_probs(c, length, cols[0], enrStart, enrEnd, 0);
	for (int i = 1; i <= c->set->nent; i++) {
		c->sortedSites[i - 1].score = 1;
		c->sortedSites[i - 1].sequence = i;
		c->sortedSites[i - 1].startPos = 1;
		const int len = c->set->n[i - 1] - length + 1;
		const int delta_0 = len < region ? len : region;
		for (int32_t j = 1; j <= len; j++) {
			double pValScore;
			const double finalPval = xo_calculate_pval(c, t, i, j, len, delta_0, length, c->set->avgLength[posSetSize], 0, &pValScore);
			if (finalPval < c->sortedSites[i - 1].score) {
				c->sortedSites[i - 1].sequence = i;
				c->sortedSites[i - 1].startPos = j;
				c->sortedSites[i - 1].score = (float) finalPval;
			}
		}
	}
	c->nbSites = posSetSize;
	set_max_sites_mops(c, posSetSize, length);
}

/* set_BindingSites, StartPosUpdater.cpp:65-76 (ThresholdChecker mode L: value < thr) */
int32_t xo_set_binding_sites(xo_ctx* c, double thr) {
	if (c->nbSites == 0) return 0;
	msort(c->sortedSites, c->nbSites, c->tmp, cmp_scores);
	if (c->sortedSites[0].score == 1) return 0;
	int bestMotifNb = 0;
	while (bestMotifNb < c->nbSites && (double) c->sortedSites[bestMotifNb].score < thr) bestMotifNb++;
	msort(c->sortedSites, bestMotifNb, c->tmp, cmp_startpos);
	return bestMotifNb;
}

/* fill_sites_with_startPosList, StartPosUpdater-inl.h:6-68 */
void xo_rescore_startposlist(xo_ctx* c, const xo_tables* t, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t enrStart,
                             int32_t enrEnd, const int32_t* siteSeq, const int32_t* sitePos, int32_t nSites, int32_t instanceThreshold) {
	const int region = enrEnd - enrStart + 1;
	const int length = cols[ncols - 1] - cols[0] + 1;
	if (c->cfg.localization) xo_init_positional_probs(c, length, cols[0], enrStart, enrEnd, 0);
	if (!c->cfg.mops)
		for (int i = 0; i < posSetSize; i++) c->sortedSites[i].score = 1;
	int sites_counter = 0;
	for (int k = 0; k < nSites; k++) {
		const int32_t seq = siteSeq[k];
		const int32_t startPos = sitePos[k] + cols[0] - 1;
		const int len = c->set->n[seq - 1] - length + 1;
		const int delta_0 = len < region ? len : region;
		double pValScore;
		const double finalPval = xo_calculate_pval(c, t, seq, startPos, len, delta_0, length, c->set->avgLength[posSetSize],
		                                           c->cfg.mops || instanceThreshold, &pValScore);
		if (finalPval >= 1) continue;
		if (c->cfg.mops) {
			if (!update_mult_occ(c, c->sortedSites, &sites_counter, seq, startPos, length, finalPval)) continue;
			sites_counter++;
		} else if (finalPval < c->sortedSites[seq - 1].score) {
			if (c->sortedSites[seq - 1].score == 1) sites_counter++;
			c->sortedSites[seq - 1].sequence = seq;
			c->sortedSites[seq - 1].startPos = startPos;
			c->sortedSites[seq - 1].score = (float) finalPval;
		}
	}
	if (c->cfg.mops) {
		c->nbSites = sites_counter;
		set_max_sites_mops(c, posSetSize, length);
	} else {
		c->nbSites = posSetSize;
		c->maxSites = posSetSize;
	}
	c->log_posSetSize = c->LOG_i[c->set->total[posSetSize]];
}

/* fill_startPosList_with_bestSubset + fill_startPosList_with_sites, StartPosUpdater.cpp:52-63, StartPosUpdater.h:241-255 */
int32_t xo_startposlist_best_subset(xo_ctx* c, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t oldTotalSites,
                                    int32_t* outSeq, int32_t* outPos, int32_t cap, int32_t* bestMotifNb, double* bestPval) {
	(void) ncols; (void) oldTotalSites;
	if (c->nbSites == 0) return -1;
	msort(c->sortedSites, c->nbSites, c->tmp, cmp_scores);
	if (c->sortedSites[0].score == 1) return -1;
	int bestCounter = 1, bestMotifNbPlus = 1, bestCounterPlus = 1;
	best_subset(c, c->sortedSites, c->nbSites, posSetSize, bestPval, bestMotifNb, &bestCounter, &bestMotifNbPlus, &bestCounterPlus);
	msort(c->sortedSites, bestCounter, c->tmp, cmp_startpos);
	memset(c->motifsPerSequenceCount, 0, ((size_t) c->set->nent + 1) * sizeof(int32_t));
	int n = 0;
	for (int i = 0, motifCounter = 0; motifCounter < *bestMotifNb; i++) {
		c->motifsPerSequenceCount[c->sortedSites[i].sequence]++;
		if (c->motifsPerSequenceCount[c->sortedSites[i].sequence] > c->cfg.maxMotifsPerSequence) continue;
		++motifCounter;
		if (n < cap) { outSeq[n] = c->sortedSites[i].sequence; outPos[n] = c->sortedSites[i].startPos - cols[0] + 1; }
		n++;
	}
	return n;
}

/* fill_startPosList_with_elongSites, StartPosUpdater.cpp:114-134 (NB: the per-sequence counter reads sortedSites[], :127) */
int32_t xo_startposlist_from_elong(xo_ctx* c, int32_t bestMotifNb, int32_t firstTmpColumn, int32_t* outSeq, int32_t* outPos, int32_t cap,
                                   double* worst) {
	if (bestMotifNb == 0) { *worst = 1; return 0; }
	memset(c->motifsPerSequenceCount, 0, ((size_t) c->set->nent + 1) * sizeof(int32_t));
	int n = 0;
	for (int i = 0, motifCounter = 0;; i++) {
		if (++motifCounter > bestMotifNb) break;
		const int32_t sq = c->sortedSites[i].sequence;
		if (sq >= 0 && sq <= c->set->nent) {
			c->motifsPerSequenceCount[sq]++;
			if (c->motifsPerSequenceCount[sq] > c->cfg.maxMotifsPerSequence) continue;
		}
		if (n < cap) { outSeq[n] = c->sortedSites_elong[i].sequence; outPos[n] = c->sortedSites_elong[i].startPos - firstTmpColumn + 1; }
		n++;
	}
	*worst = c->sortedSites_elong[bestMotifNb - 1].score;
	return n;
}

/* Output::printScoreDistribution counts, output.cpp:1033-1078 */
void xo_logodds_hist(const xo_ctx* c, const double* pwm, const int32_t* cols, int32_t ncols, int32_t posSetSize, double* fwd, double* bwd,
                     int32_t npos, double* maxScoreOut, double* maxScore2Out) {
	const int first = cols[0], last = cols[ncols - 1];
	const int motifLength = last - first + 1;
	double pal[(XO_PWM_LENGTH + 1) * 5];
	memcpy(pal, pwm, sizeof pal);
	for (int j = first; j <= last; j++) /* Motif::getPalindrome, Motif.cpp:1068-1079 */
		for (int a = 1; a <= 4; a++) pal[(last - j + first) * 5 + (4 - a + 1)] = pwm[j * 5 + a];
	const double* P[2] = { pwm, pal };
	double maxS[2];
	for (int w = 0; w < 2; w++) { /* getMaxMatrixScore, Motif.h:601-612 */
		double m = 0;
		for (int k = 0; k < ncols; k++) {
			double mx = 0;
			for (int base = 1; base <= 4; base++) { const double v = P[w][cols[k] * 5 + base] - c->cfg.bgLog[base]; if (v > mx) mx = v; }
			m += mx;
		}
		maxS[w] = m;
	}
	*maxScoreOut = maxS[0]; *maxScore2Out = maxS[1];
	const double LOG1_1000 = log(1.0 / 1000);
	for (int pos = 1; pos <= c->set->max_leng - ncols + 1 && pos <= npos; pos++) {
		double cf = 0, cb = 0;
		for (int seq = 1; seq <= posSetSize; seq++) {
			if (c->set->n[seq - 1] - motifLength + 1 < pos) continue;
			const uint8_t* S = c->set->S[seq - 1];
			double sc[2];
			for (int w = 0; w < 2; w++) { /* calculateMatrixScore_logOdds, Motif.h:584-599 */
				double smin = 0;
				for (int k = 0; k < ncols; k++) {
					const int base = S[pos + cols[k] - first];
					if (base == 0) { smin += (LOG1_1000 + c->LOG_i[2]); continue; }
					smin += P[w][cols[k] * 5 + base] - c->cfg.bgLog[base];
				}
				sc[w] = smin;
			}
			if (sc[0] > maxS[0] * 0.5) cf++;
			if (sc[1] > maxS[1] * 0.5) cb++;
		}
		fwd[pos - 1] = cf; bwd[pos - 1] = cb;
	}
}

/* getCombinedPval_log, pValCalculation.h:122-148 */
static double combined_pval_log(double log_prod_pcons, int motifNb) {
	double result = log_prod_pcons;
	double quotient = 1, sum = 1;
	const double overflowConst = 1e100;
	const double log_overflowConst = 230.258509299;
	for (int i = 1; i < motifNb; i++) {
		quotient *= ((-log_prod_pcons) / i);
		sum += quotient;
		if (sum > overflowConst) { sum /= overflowConst; quotient /= overflowConst; result += log_overflowConst; }
	}
	return result + log(sum);
}

/* setPositionalPval, StartPosUpdater.cpp:207-274 ; returns the motif's posPval (or NAN on "no man's land") */
double xo_positional_pval(const xo_ctx* c, int32_t enrMax, int32_t firstMotifColumn, int32_t motifLength, const int32_t* sitePos,
                          int32_t nSites) {
	double log_prod_pPos = 0.0;
	int max = enrMax + firstMotifColumn - 1;
	if (max > 1) {
		int L = c->set->max_leng;
		if (c->cfg.revcomp) L = (L - 1) / 2;
		const int L1 = L - motifLength + 1;
		if (c->cfg.revcomp) {
			if (max > L + 1) max = 1 + c->set->max_leng - max - motifLength + 1;
			else if (max > L1) return NAN;
		}
		for (int k = 0; k < nSites; k++) {
			int pos = sitePos[k] + firstMotifColumn - 1;
			if (c->cfg.revcomp) {
				if (pos > L + 1) pos = 1 + c->set->max_leng - pos - motifLength + 1;
				else if (pos > L1) return NAN;
			}
			int distance = max - pos;
			if (distance < 0) {
				distance = abs(distance);
				distance += (max - 1 < distance ? max - 1 : distance);
			} else if (distance > 0) {
				distance += (distance < L1 - max ? distance : L1 - max);
			}
			const double pValPos = (double) (distance + 1) / (double) (L1);
			log_prod_pPos += log(pValPos);
		}
		return combined_pval_log(log_prod_pPos, nSites);
	}
	return 0.0;
}
