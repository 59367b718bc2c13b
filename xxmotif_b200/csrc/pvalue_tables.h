// Host-side construction of the k-mer -> p-value tables for one (PWM, column set).
//
// Follows Kmer_Generator_Reloaded<DnaHash>::reinitialize / set_kmer_to_Pvalue_hash / create_similar_kmers_rec /
// calculate_kmer_probability / recalculatePrefixProbabilities / getRightKmerConditional
// (reference src/branch_and_bound-inl.h:54-100, :183-323, :462-603) -- SURVEY.md section 8 row a5 ("stays on the host in
// the first slice; rows are staged H2D per iteration").  The result is kept in compact form (enumerated ids with their
// class, classes with their cumulative p-value); the dense 4^k arrays of the reference are only materialised on the
// device as rank tables.
#pragma once
#include <cstdint>
#include <vector>

namespace xxm {

struct NullModelView {
	int order = 0;
	const float* rprobs = nullptr;  // 2^(3(order+1)) entries, index = 3 bits per char, first char most significant
	const float* rconds = nullptr;
};

struct TablePart {
	int len = 0;                      // number of match columns covered by this part
	std::vector<uint32_t> ids;        // enumerated k-mer ids in class order (descending score, then enumeration order)
	std::vector<uint32_t> cls;        // class index of ids[i]
	std::vector<double> prob;         // background probability of ids[i] (the reference's probabilityHash, left part only)
	std::vector<int32_t> classScore;  // integer score of every class, descending
	std::vector<int32_t> classCount;  // ids per class
	std::vector<double> classPval;    // cumulative p-value of the class == pValueHash[part][id] for all its ids
};

struct PvalueTables {
	int error = 0;            // 1: reference behaviour undefined (int overflow for long gapped left parts), 2: PWM column not normalised
	int L = 0;                // match columns
	int maxk = 8;             // MAX_KMER_SIZE
	int nparts = 1;
	int span = 0;             // lastCol - firstCol + 1
	int coloff[32] = {0};     // offset of match column k from the first match column
	int gaps[32] = {0};
	double logodds[32][5];    // PWM[i][a] = pwm[col_i][a] - bgLog[a], i = 1..L
	double pruneThr = 0;      // posSetSize / (totalPositions * 2.0)
	TablePart part[2];
	std::vector<double> leftClassProb;  // per left class: sum of prob over its ids (list order)
};

// pieces of the construction the device build (xxm_tables_device.cuh) shares with the host build: the model header
// (log-odds, gaps, thresholds) and the visiting order of create_similar_kmers_rec for the left part of `len` columns
struct PartPlan {
	int len;
	double sw[9][5];   // [level 1..len][rank 1..4] log-odds of the level's column, descending
	double smax[10];   // [level] best score still reachable from this level on
	int shift[9];      // [level] bit position of the level's column in the k-mer id
	int code[9][5];    // [level][rank] base code 0..3
};
void model_header(const double* pwm31x5, const int32_t* cols, int ncols, int64_t posSetSize, int64_t totalPositions, const double* bgLog5, PvalueTables& out);
void part_plan(const PvalueTables& t, int len, PartPlan& out);

// Background probabilities of ungapped left k-mers under ONE null model, kept across the table builds of a refinement run:
// calculate_kmer_probability (branch_and_bound-inl.h:54-100) depends on the k-mer and the model only, and successive PWMs
// enumerate largely the same k-mers, so after the first few builds a leaf costs one table read instead of len dependent
// multiplications.  A value is stored exactly as computed (no re-association), so cached and fresh builds are bit-identical.
// The owner must call clear() whenever the null model changes (xxm_set_background); the key below guards against the rest.
struct BgProbCache {
	int order = -1;
	const float *rprobs = nullptr, *rconds = nullptr;
	std::vector<double> p[9];  // p[len]: [4^len], NaN = not computed yet (the column variants of a refinement run alternate between lengths)
	void clear() { order = -1; rprobs = rconds = nullptr; for (auto& v : p) v.clear(); }
	// scratch of the enumeration (leaves and their sorted copy), kept so that a build does not start from empty vectors
	struct LeafRec { int32_t score; uint32_t id; };
	std::vector<LeafRec> leaves, sorted;
};

void build_pvalue_tables(const double* pwm31x5, const int32_t* cols, int ncols, int64_t posSetSize, int64_t totalPositions,
                         const double* bgLog5, const NullModelView& nm, PvalueTables& out, BgProbCache* cache = nullptr);

}  // namespace xxm
