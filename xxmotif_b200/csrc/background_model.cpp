// Host-side background statistics that feed the p-value tables: the order-k interpolated Markov model with gap
// symbol (NullModel::init, reference src/NullModel.cpp:88-326, called from setBackgroundDistribution,
// backgroundDistribution.cpp:10-16) and the mono-nucleotide frequencies of checkSequenceSet
// (backgroundDistribution.cpp:50-95).  SURVEY.md section 8(f) row 3 ("next": GPU k-mer counting); host code for now.
//
// All probabilities are single precision, exactly like the reference (float _counts/_conds/_probs), and every k-mer is
// visited in the reference's enumeration order because the gapped marginals are float accumulations.
#include "../../include/xxmotif_gpu.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct Tables {
	int order, base = 5, asize = 4;
	std::vector<int> pow5;          // _coeffs
	std::vector<float> counts, conds, probs;
	// index of a (k+1)-mer given as digits d[0..k] (1..4 = ACGT, 5 = gap), first digit most significant (sub2ind, byrow)
	int index(const uint8_t* d, int k) const {
		int i = 0;
		for (int j = 0; j <= k; ++j) i += pow5[k - j] * d[j];
		return i;
	}
};

// next digit string in lexicographic order over digits lo..hi; returns false after the last one
bool next_digits(uint8_t* d, int len, uint8_t lo, uint8_t hi) {
	for (int p = len - 1; p >= 0; --p) {
		if (d[p] < hi) { ++d[p]; return true; }
		d[p] = lo;
	}
	return false;
}

}  // namespace

// the model from the k-mer counts (indexed like NullModel::_counts); shared by the host counting loop below and the device
// histogram (xxm_kmer_counts, xxm_gpu_part6.inc).  The reference increments FLOAT counters, which stop growing at 2^24;
// integer counting + the same saturation is bit-identical.
extern "C" int xxm_null_model_from_counts(const uint64_t* counts64, int32_t order, float alpha, float beta, float* rprobs, float* rconds) {
	if (order < 0 || order > 9) return 1;
	Tables T;
	T.order = order;
	T.pow5.resize(order + 2);
	for (int k = 0, p = 1; k < order + 2; ++k) { T.pow5[k] = p; p *= 5; }
	int fields = 1 + 5;
	for (int i = 2, p = 5; i < order + 2; ++i) { p *= 5; fields += p; }
	T.counts.assign(fields, 0.0f);
	T.conds.assign(fields, 0.0f);
	T.probs.assign(fields, 0.0f);
	const float factor = 4.0f * alpha;
	for (int i = 0; i < fields; ++i) T.counts[i] = static_cast<float>(std::min<uint64_t>(counts64[i], 16777216ull));
	if (beta > 0)
		for (int i = 1; i < fields; ++i) T.counts[i] = std::max(0.0f, static_cast<float>(T.counts[i] - beta));
	float ncounts = 0;
	for (int i = 1; i <= 4; ++i) ncounts += T.counts[i];
	for (int a = 1; a <= 4; ++a) T.conds[a] = T.probs[a] = T.counts[a] / ncounts;  // no pseudocounts at order 0 (:161-166)

	std::vector<uint8_t> d(order + 2, 1);
	// orders 1..k: interpolated conditionals, renormalised over the 4 extensions (calculateKmerProbs, :245-295)
	for (int o = 1; o <= order; ++o) {
		std::fill(d.begin(), d.end(), 1);
		do {  // all prefixes d[0..o-1] over ACGT in lexicographic order
			int ii[5], ip[5];
			float S = 0;
			for (uint8_t a = 1; a <= 4; ++a) {
				d[o] = a;
				ii[a] = T.index(d.data(), o);
				ip[a] = T.index(d.data(), o - 1);
				const int suffix = T.index(d.data() + 1, o - 1);
				const float c = (T.counts[ii[a]] + factor * T.conds[suffix]) / (T.counts[ip[a]] + factor);
				S += c;
				T.conds[ii[a]] = c;
			}
			for (uint8_t a = 1; a <= 4; ++a) {
				T.conds[ii[a]] /= S;
				T.probs[ii[a]] = T.conds[ii[a]] * T.probs[ip[a]];
			}
		} while (next_digits(d.data(), o, 1, 4));
	}
	// gapped k-mers by marginalisation (calculateGappedKmerProbs, :204-242): digit 5 = gap
	T.conds[5] = T.probs[5] = 1;
	for (int o = 1; o <= order; ++o) {
		std::fill(d.begin(), d.end(), 1);
		do {  // all (o+1)-mers over {A,C,G,T,gap} in lexicographic order
			int lastGap = -1;
			for (int p = 0; p <= o; ++p) if (d[p] == 5) lastGap = p;
			const int i = T.index(d.data(), o);
			if (lastGap > -1) T.conds[i] = T.probs[i] / T.probs[T.index(d.data(), o - 1)];
			for (int k = lastGap + 1; k <= o; ++k) {
				const uint8_t x = d[k];
				d[k] = 5;
				T.probs[T.index(d.data(), o)] += T.probs[i];
				d[k] = x;
			}
		} while (next_digits(d.data(), o + 1, 1, 5));
	}
	// 3-bit-per-character index structure, longest k-mers first so that shorter ones overwrite (copyProbs, :193-196, :297-326)
	const size_t rfields = static_cast<size_t>(1) << (3 * (order + 1));
	std::memset(rprobs, 0, rfields * sizeof(float));
	std::memset(rconds, 0, rfields * sizeof(float));
	for (int o = order; o >= 0; --o) {
		std::fill(d.begin(), d.end(), 1);
		do {
			int r = 0;
			for (int p = 0; p <= o; ++p) r = (r << 3) + (d[p] == 5 ? 0 : d[p]);
			const int i = T.index(d.data(), o);
			rprobs[r] = T.probs[i];
			rconds[r] = T.conds[i];
		} while (next_digits(d.data(), o + 1, 1, 5));
	}
	return 0;
}

// counts of all (k+1)-mers, k = 0..order, that contain no N (NullModel.cpp:122-132), on the host
extern "C" int xxm_host_null_model(int32_t nent, const int32_t* n, const uint8_t* codes, int32_t order, float alpha, float beta, float* rprobs,
                                   float* rconds) {
	if (order < 0 || order > 9 || nent < 1) return 1;
	int fields = 1 + 5;
	for (int i = 2, p = 5; i < order + 2; ++i) { p *= 5; fields += p; }
	std::vector<uint64_t> c(fields, 0);
	size_t off = 0;
	for (int s = 0; s < nent; ++s) {
		const uint8_t* q = codes + off;
		const int L = n[s];
		for (int l = 0; l < L; ++l) {
			int idx = 0;
			for (int k = 0; k <= order && l + k < L && q[l + k]; ++k) {
				idx = idx * 5 + q[l + k];  // sub2ind of the (k+1)-mer: appending a digit multiplies by 5
				c[idx]++;
			}
		}
		off += static_cast<size_t>(L);
	}
	return xxm_null_model_from_counts(c.data(), order, alpha, beta, rprobs, rconds);
}

// mono-nucleotide background of checkSequenceSet (backgroundDistribution.cpp:50-95): counts include code 0, the
// frequencies are normalised by (total - count[0]) and floored at 1e-100
extern "C" int xxm_host_background(int32_t nent, const int32_t* n, const uint8_t* codes, double* bg5, double* bgLog5) {
	double tot = 0;
	for (int a = 0; a < 5; ++a) { bg5[a] = 0; bgLog5[a] = 0; }
	size_t off = 0;
	for (int s = 0; s < nent; ++s) {
		tot += n[s];
		for (int j = 0; j < n[s]; ++j) bg5[codes[off + j]]++;
		off += static_cast<size_t>(n[s]);
	}
	for (int a = 1; a < 5; ++a) {
		bg5[a] = std::max(1e-100, bg5[a] / (tot - bg5[0]));
		bgLog5[a] = std::log(bg5[a]);
	}
	return 0;
}
