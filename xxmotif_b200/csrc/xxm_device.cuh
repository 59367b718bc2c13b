// Device-side data layout and small device helpers shared by the kernels of libxxmotif_gpu.so (sm_100a only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace xxm {

// ---------------------------------------------------------------------------------------------------------------
// Sequence stream in HBM.  Every sequence starts at a 32-base word boundary and is followed by >= 32 positions
// flagged N (padding), so that (a) no window that starts inside a sequence can reach the next sequence with a match
// column (span <= 30) and (b) windows that would run past the end of the sequence die on an N exactly like the
// reference never enumerates them.  Per 32 bases: one 64-bit word of 2-bit codes (A=0,C=1,G=2,T=3; base j in bits
// [2j+1:2j]) and one 32-bit N mask  ->  12 bytes / 32 bases = 0.375 B per stored base.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPadBases = 32;
constexpr int kBlkWords = 32;  // granularity of the word -> sequence coarse index

struct SeqStream {
	const uint64_t* codes;     // [nwords + 2]
	const uint32_t* nmask;     // [nwords + 2]
	const uint32_t* wordOff;   // [nent + 1] first word of sequence i (0-based sequence index)
	const int32_t* len;        // [nent] stored length n_i (after revcomp doubling)
	const uint32_t* blk2seq;   // [nwords / kBlkWords + 1] sequence that owns word (b * kBlkWords)
	uint32_t nwords;
	int32_t nent;
	uint32_t uniformWps;       // words per sequence when all sequences have the same stored length, else 0
	int32_t uniformLen;        // that stored length (0 for ragged sets)
};

// ---------------------------------------------------------------------------------------------------------------
// The model of the current reinitialize(): rank tables of the k-mer -> p-value maps + what the >8-column path needs.
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr int32_t kMemoEmpty = INT32_MIN;

struct Model {
	int L;                 // match columns
	int maxk;              // columns of the left part
	int nparts;
	int span;
	int contiguous;        // 1: no gap columns
	int len0, len1;        // columns in part 0 / 1
	int coloff[32];        // offset of match column k from the window start
	const uint32_t* rank0; // [4^len0] class index + 1 of the left k-mer (0: not enumerated)
	const uint32_t* rank1; // [4^len1]
	const uint32_t* hit0;  // bitmap over 4^len0: left k-mer is in the table
	const uint8_t* t8;     // [65536] stage-1 byte table of the scan (xxm_scan2.cuh)
	uint32_t exactBits;    // windows (bit j of a word) whose stage-1 bit already is the exact left-k-mer test
	const double* cpval0;  // [nclass0] cumulative p-value of left classes  (pValueHash[0])
	const double* cpval1;  // [nclass1]
	const int32_t* cscore0;// [nclass0] descending
	const int32_t* cscore1;// [nclass1] descending
	const double* leftProb;// [nclass0]
	int nclass0, nclass1;
	double pruneThr;       // posSetSize / (2 totalPositions)
	double corr;           // overrepCorrection ^ L
	double logodds[32][4]; // [column 0..L-1][base code 0..3]
	// memo of the >8-column path: distinct integer scores s -> slot
	int32_t* memoKey;      // [memoCap]
	double* memoVal;       // [memoCap] pValue of the slot (0 until evaluated)
	uint32_t* memoState;   // [memoCap] 0: fresh, 1: evaluated
	uint32_t memoCap;      // power of two
	uint32_t* memoOverflow;
};

// ---------------------------------------------------------------------------------------------------------------
// Value tables: everything a consumer needs to know about "value index" v (class index for <= 8 columns, memo slot
// otherwise).  finalKind selects how finalPval is obtained.
// ---------------------------------------------------------------------------------------------------------------
enum FinalKind { kFinalTable = 0, kFinalLocMops = 1, kFinalPerCand = 2 };

struct Values {
	const double* pS;       // [nval] pValScore after the 1.05^L correction (1.0: not a site)
	const double* finalTab; // [nval] finalPval for kFinalTable
	const double* posProb;  // [max_leng + 1] positional prior for kFinalLocMops
	const double* perCand;  // [ncand] finalPval per candidate for kFinalPerCand
	int kind;
};

__device__ __forceinline__ double final_of(const Values& V, uint32_t v, uint32_t cand, int pos) {
	if (V.kind == kFinalTable) return V.finalTab[v];
	if (V.kind == kFinalPerCand) return V.perCand[cand];
	// (pValScore*sqrt(pValPos) - pValScore*pValScore*pValPos*0.5)*2, pValCalculation.h:79, evaluated left to right
	const double pS = V.pS[v];
	if (pS == 1.0) return 1.0;
	const double pP = V.posProb[pos];
	const double a = __dmul_rn(pS, __dsqrt_rn(pP));
	const double b = __dmul_rn(__dmul_rn(__dmul_rn(pS, pS), pP), 0.5);
	return __dmul_rn(__dsub_rn(a, b), 2.0);
}

// base code (0..3) at offset o (0..63) of the 64-base register pair
__device__ __forceinline__ uint32_t base_at(uint64_t c0, uint64_t c1, int o) {
	return (o < 32) ? static_cast<uint32_t>(c0 >> (2 * o)) & 3u : static_cast<uint32_t>(c1 >> (2 * (o - 32))) & 3u;
}

// 2*len bits starting at base offset j (0..31) of the pair, len <= 16
__device__ __forceinline__ uint32_t kmer_at(uint64_t c0, uint64_t c1, int j, int len) {
	const uint64_t lo = c0 >> (2 * j);
	const uint64_t hi = (j == 0) ? 0ull : (c1 << (64 - 2 * j));
	return static_cast<uint32_t>(lo | hi) & ((len >= 16) ? 0xFFFFFFFFu : ((1u << (2 * len)) - 1u));
}

__device__ __forceinline__ uint32_t find_seq(const SeqStream& S, uint32_t w) {
	uint32_t s = S.blk2seq[w / kBlkWords];
	while (S.wordOff[s + 1] <= w) ++s;
	return s;
}

}  // namespace xxm
