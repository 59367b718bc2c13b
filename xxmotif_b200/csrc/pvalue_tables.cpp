// See pvalue_tables.h.  All double arithmetic is written operation by operation in the reference's order (the
// translation unit is compiled with -ffp-contract=off) because these values define every site p-value.
#include "pvalue_tables.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>

namespace xxm {
namespace {

using Leaf = BgProbCache::LeafRec;  // { int32_t score; uint32_t id; }

// 3 bits per character, first character most significant (UngappedKmer.h:21-31)
inline int packed_kmer(const uint8_t* s, int len) {
	unsigned int r = s[0];
	for (int i = 1; i < len; ++i) r = (r << 3) + s[i];
	return static_cast<int>(r);
}

class Builder {
public:
	Builder(PvalueTables& t, const NullModelView& nm, BgProbCache* cache = nullptr)
		: t_(t), nm_(nm), cache_(cache), leaves_(cache ? cache->leaves : ownLeaves_), tmp_(cache ? cache->sorted : ownTmp_) {}

	// one call of set_kmer_to_Pvalue_hash (branch_and_bound-inl.h:462-577) for match columns [start, stop] (1-based)
	void build_part(int part, int start, int stop, int threshold, bool collectPrefixes, int gapOffset) {
		part_ = part; start_ = start; len_ = stop - start + 1; T_ = threshold; collect_ = collectPrefixes; gapOffset_ = gapOffset;
		ungapped_ = true;
		for (int m = 0; m + 1 < len_; ++m) ungapped_ = ungapped_ && t_.gaps[m + gapOffset_] == 0;
		order_columns();
		leaves_.clear();
		enumerate(1, 0.0, 0u);
		// classes by descending integer score; inside a class the enumeration order is kept (std::map + list::push_back)
		sort_leaves_by_score_desc();
		TablePart& P = t_.part[part];
		P.len = len_;
		const size_t n = leaves_.size();
		// the cache serves the rolled-index case of background_probability (left part, ungapped, longer than the model's context)
		double* memo = nullptr;
		if (cache_ && part_ == 0 && !collect_ && ungapped_ && len_ > nm_.order + 1 && len_ <= 8) {
			BgProbCache& C = *cache_;
			if (C.order != nm_.order || C.rprobs != nm_.rprobs || C.rconds != nm_.rconds) {
				C.clear();
				C.order = nm_.order; C.rprobs = nm_.rprobs; C.rconds = nm_.rconds;
			}
			std::vector<double>& v = C.p[len_];
			if (v.size() != (static_cast<size_t>(1) << (2 * len_))) v.assign(static_cast<size_t>(1) << (2 * len_), std::nan(""));
			memo = v.data();
		}
		P.ids.resize(n); P.cls.resize(n);
		if (collect_) P.prob.resize(n);
		P.classScore.clear(); P.classCount.clear(); P.classPval.clear();
		double cumulative = 0;
		size_t i = 0;
		while (i < n) {
			size_t j = i;
			while (j < n && leaves_[j].score == leaves_[i].score) ++j;
			const uint32_t c = static_cast<uint32_t>(P.classScore.size());
			for (size_t k = i; k < j; ++k) {
				double p;
				if (memo) {
					p = memo[leaves_[k].id];
					if (p != p) { p = background_probability(leaves_[k].id); memo[leaves_[k].id] = p; }
				} else {
					p = background_probability(leaves_[k].id);
				}
				if (collect_) P.prob[k] = p;
				cumulative += p;
				P.ids[k] = leaves_[k].id;
				P.cls[k] = c;
			}
			P.classScore.push_back(leaves_[i].score);
			P.classCount.push_back(static_cast<int32_t>(j - i));
			P.classPval.push_back(cumulative);
			i = j;
		}
	}

	// recalculatePrefixProbabilities (branch_and_bound-inl.h:260-301): keep the left k-mers within 0.8 of the most
	// probable one under the PWM; only the survivor set is used afterwards.
	void filter_prefixes() {
		float best = 0;
		std::vector<float> w(prefixes_.size());
		for (size_t i = 0; i < prefixes_.size(); ++i) {
			int col = 1;
			float pr = 1;
			const int klen = prefixLen_[i];
			for (int c = 0; c < klen; ++c) {
				const int chr = (prefixes_[i] >> ((klen - 1 - c) * 3)) & 7;
				if (chr != 0) {
					pr = static_cast<float>(pr * t_lin_[col][chr]);
					++col;
				}
			}
			w[i] = pr;
			if (pr > best) best = pr;
		}
		size_t keep = 0;
		for (size_t i = 0; i < prefixes_.size(); ++i) {
			if (w[i] < 0.8 * best) continue;
			prefixes_[keep] = prefixes_[i];
			prefixLen_[keep] = prefixLen_[i];
			++keep;
		}
		prefixes_.resize(keep);
		prefixLen_.resize(keep);
		rightCond_.clear();
	}

	void clear_prefixes() { prefixes_.clear(); prefixLen_.clear(); }
	double t_lin_[32][5];

private:
	// stable LSD radix sort (3 x 11 bits) on the key "larger score first": same order as a stable sort by descending score,
	// which is what the reference's std::map<int, list> + push_back produce, at a fraction of std::stable_sort's cost
	void sort_leaves_by_score_desc() {
		const size_t n = leaves_.size();
		if (n < 2) return;
		tmp_.resize(n);
		if (n <= 16384) {
			// small tables (the usual case: a few thousand leaves): one stable counting pass into 2048 score-range buckets,
			// then a stable insertion sort that only ever moves elements inside their bucket
			int32_t mn = leaves_[0].score, mx = mn;
			for (size_t i = 1; i < n; ++i) { mn = std::min(mn, leaves_[i].score); mx = std::max(mx, leaves_[i].score); }
			const uint64_t range = static_cast<uint64_t>(static_cast<int64_t>(mx) - mn) + 1;
			const uint64_t scale = (2048ull << 32) / range;
			uint32_t cnt[2049];
			std::memset(cnt, 0, sizeof cnt);
			auto bucket = [&](int32_t sc) { return static_cast<uint32_t>((static_cast<uint64_t>(static_cast<int64_t>(mx) - sc) * scale) >> 32); };
			for (size_t i = 0; i < n; ++i) cnt[bucket(leaves_[i].score) + 1]++;
			for (int b = 0; b < 2048; ++b) cnt[b + 1] += cnt[b];
			for (size_t i = 0; i < n; ++i) tmp_[cnt[bucket(leaves_[i].score)]++] = leaves_[i];
			size_t lo = 0;
			for (int b = 0; b < 2048; ++b) {  // after the scatter cnt[b] is the end of bucket b
				const size_t hi = cnt[b];
				if (hi - lo > 24) {
					std::stable_sort(tmp_.begin() + lo, tmp_.begin() + hi, [](const Leaf& x, const Leaf& y) { return x.score > y.score; });
				} else {
					for (size_t i = lo + 1; i < hi; ++i) {
						const Leaf x = tmp_[i];
						size_t j = i;
						while (j > lo && tmp_[j - 1].score < x.score) { tmp_[j] = tmp_[j - 1]; --j; }
						tmp_[j] = x;
					}
				}
				lo = hi;
			}
			leaves_.swap(tmp_);
			return;
		}
		Leaf* src = leaves_.data();
		Leaf* dst = tmp_.data();
		for (int pass = 0; pass < 3; ++pass) {
			uint32_t cnt[2049];
			std::memset(cnt, 0, sizeof cnt);
			const int sh = 11 * pass;
			for (size_t i = 0; i < n; ++i) cnt[((~(static_cast<uint32_t>(src[i].score) ^ 0x80000000u) >> sh) & 2047u) + 1]++;
			for (int b = 0; b < 2048; ++b) cnt[b + 1] += cnt[b];
			for (size_t i = 0; i < n; ++i) dst[cnt[(~(static_cast<uint32_t>(src[i].score) ^ 0x80000000u) >> sh) & 2047u]++] = src[i];
			std::swap(src, dst);
		}
		if (src != leaves_.data()) std::memcpy(leaves_.data(), src, n * sizeof(Leaf));
	}

	// rowSwap / colSwap / swappedPWM / S_max (branch_and_bound-inl.h:470-536).  The column comparison indexes the
	// log-odds matrix with the part-relative column number on its left side, as the reference does.
	void order_columns() {
		for (int c = 1; c <= len_; ++c) {
			const int abs = start_ + c - 1;
			int* r = row_[c];
			r[1] = 1;
			for (int j = 2; j <= 4; ++j) {
				int s = j;
				while (s > 1 && t_.logodds[abs][r[s - 1]] < t_.logodds[abs][j]) { r[s] = r[s - 1]; --s; }
				r[s] = j;
			}
		}
		col_[1] = 1;
		for (int c = 2; c <= len_; ++c) {
			const int abs = start_ + c - 1;
			int s = c;
			while (s > 1 && t_.logodds[col_[s - 1]][row_[col_[s - 1]][1]] < t_.logodds[abs][row_[c][1]]) { col_[s] = col_[s - 1]; --s; }
			col_[s] = c;
		}
		for (int c = 1; c <= len_; ++c)
			for (int j = 1; j <= 4; ++j) sw_[c][j] = t_.logodds[col_[c] + start_ - 1][row_[col_[c]][j]];
		smax_[len_ + 1] = 0;
		for (int c = len_; c > 0; --c) smax_[c] = (c < len_) ? smax_[c + 1] + sw_[c][1] : sw_[c][1];
	}

	// create_similar_kmers_rec (branch_and_bound-inl.h:586-603)
	// (the k-mer id -- 2 bits per match column, first column least significant -- is accumulated on the way down)
	void enumerate(int j, double sj, uint32_t idj) {
		const int sh = 2 * (col_[j] - 1);
		const int* rw = row_[col_[j]];
		for (int a = 1; a <= 4; ++a) {
			const double s1 = sj + sw_[j][a];
			if (static_cast<int>((s1 + smax_[j + 1]) * 1e6) < T_) break;  // the reference says `continue`; sw_[j][] is sorted descending, so every later base fails too
			const uint32_t id = idj | (static_cast<uint32_t>(rw[a] - 1) << sh);
			if (j + 1 <= len_) enumerate(j + 1, s1, id);
			else leaves_.push_back(Leaf{static_cast<int32_t>(s1 * 1e6), id});
		}
	}

	// calculate_kmer_probability (branch_and_bound-inl.h:54-100)
	double background_probability(uint32_t kmerId) {
		if (part_ == 0 && !collect_ && ungapped_ && len_ > nm_.order + 1) {
			// same table entries multiplied in the same order as below; the (order+1)-character index is rolled along
			const int o1 = nm_.order + 1;
			const uint32_t mask = (1u << (3 * o1)) - 1u;
			uint32_t id = kmerId, k = 0;
			for (int m = 0; m < o1; ++m) { k = (k << 3) + (id & 3u) + 1u; id >>= 2; }
			double p = nm_.rprobs[k];
			for (int m = o1; m < len_; ++m) {
				k = ((k << 3) + (id & 3u) + 1u) & mask;
				id >>= 2;
				p *= nm_.rconds[k];
			}
			return p;
		}
		uint8_t ex[128];
		ex[0] = 0;
		uint32_t id = kmerId;
		int n = 0;
		for (int m = 0; m < len_; ++m) {
			ex[n++] = static_cast<uint8_t>((id & 3) + 1);
			id >>= 2;
			for (int g = 0; g < t_.gaps[m + gapOffset_]; ++g) ex[n++] = 0;
		}
		const int o = nm_.order;
		if (part_ == 0) {
			double p;
			if (n <= o + 1) {
				p = nm_.rprobs[packed_kmer(ex, n)];
			} else {
				p = nm_.rprobs[packed_kmer(ex, o + 1)];
				for (int stop = o + 1; stop < n; ++stop) p *= nm_.rconds[packed_kmer(ex + stop - o, o + 1)];
			}
			if (collect_) {
				if (n > 10) t_.error = 1;  // the reference's int id overflows (UngappedKmer.h:21-31): undefined there
				prefixes_.push_back(packed_kmer(ex, n));
				prefixLen_.push_back(n);
			}
			return p;
		}
		double p = 1;
		for (int stop = 0; stop < n; ++stop) {
			const int start = std::max(0, stop - o);
			const int len = stop - start + 1;
			const int k = packed_kmer(ex + start, len);
			if (len <= o) {
				const double rp = right_conditional(k, len);
				p *= rp;
			} else {
				p *= nm_.rconds[k];
			}
		}
		return p;
	}

	// getRightKmerConditional (branch_and_bound-inl.h:303-323): max over the surviving left k-mers of the conditional
	// of (tail of the left k-mer . k); cached by the integer id only, like the reference's unordered_map<int,float>
	float right_conditional(int kid, int klen) {
		auto it = rightCond_.find(kid);
		if (it != rightCond_.end()) return it->second;
		const int tail = nm_.order - klen + 1;  // stop - start + 1 with start = maxk - order + klen - 1, stop = maxk - 1
		const uint64_t mask = static_cast<uint64_t>(static_cast<int64_t>(static_cast<int>((1u << (tail * 3)) - 1u)));
		double best = 0;
		for (size_t i = 0; i < prefixes_.size(); ++i) {
			const uint64_t idx = ((static_cast<uint64_t>(static_cast<int64_t>(prefixes_[i])) & mask) << (klen * 3)) + static_cast<uint64_t>(kid);
			best = std::max(best, static_cast<double>(nm_.rconds[idx]));
		}
		const float f = static_cast<float>(best);
		rightCond_[kid] = f;
		return f;
	}

	PvalueTables& t_;
	const NullModelView& nm_;
	BgProbCache* cache_ = nullptr;
	int part_ = 0, start_ = 1, len_ = 0, T_ = 0, gapOffset_ = 0;
	bool collect_ = false, ungapped_ = false;
	int row_[34][5];
	int col_[34];
	double sw_[34][5];
	double smax_[36];
	std::vector<Leaf> ownLeaves_, ownTmp_;
	std::vector<Leaf>& leaves_;  // the cache's scratch when there is one
	std::vector<Leaf>& tmp_;
	std::vector<int> prefixes_;
	std::vector<int> prefixLen_;
	std::unordered_map<int, float> rightCond_;
};

}  // namespace

// the part of reinitialize (branch_and_bound-inl.h:183-258) that does not enumerate anything
static void model_header_impl(const double* pwm, const int32_t* cols, int ncols, int64_t posSetSize, int64_t totalPositions, const double* bgLog,
                              PvalueTables& t, double (*tlin)[5]) {
	t.error = 0;
	t.L = ncols;
	t.maxk = 8;
	if (t.L > 17) t.maxk = t.L / 2;
	t.nparts = (t.L <= t.maxk) ? 1 : 2;
	t.span = cols[ncols - 1] - cols[0] + 1;
	t.pruneThr = posSetSize / (totalPositions * 2.0);
	for (int i = 0; i < ncols; ++i) t.coloff[i] = cols[i] - cols[0];
	for (int i = 1; i < ncols; ++i) t.gaps[i - 1] = cols[i] - cols[i - 1] - 1;
	t.gaps[ncols - 1] = 0;
	for (int i = 1; i <= ncols; ++i) {
		double sum = 0;
		for (int a = 1; a <= 4; ++a) {
			t.logodds[i][a] = pwm[cols[i - 1] * 5 + a] - bgLog[a];
			const double lin = std::exp(pwm[cols[i - 1] * 5 + a]);
			if (tlin) tlin[i][a] = lin;
			sum += lin;
		}
		t.logodds[i][0] = 0;
		if (std::fabs(sum - 1) >= 1e-3) t.error = 2;  // the reference asserts (branch_and_bound-inl.h:223)
	}
	t.part[1] = TablePart();
	t.leftClassProb.clear();
}

void model_header(const double* pwm, const int32_t* cols, int ncols, int64_t posSetSize, int64_t totalPositions, const double* bgLog, PvalueTables& t) {
	model_header_impl(pwm, cols, ncols, posSetSize, totalPositions, bgLog, t, nullptr);
}

// rowSwap / colSwap / swappedPWM / S_max of the left part (the same statements as Builder::order_columns), as data for the
// device build
void part_plan(const PvalueTables& t, int len, PartPlan& P) {
	int row[34][5], col[34];
	for (int c = 1; c <= len; ++c) {
		int* r = row[c];
		r[1] = 1;
		for (int j = 2; j <= 4; ++j) {
			int s = j;
			while (s > 1 && t.logodds[c][r[s - 1]] < t.logodds[c][j]) { r[s] = r[s - 1]; --s; }
			r[s] = j;
		}
	}
	col[1] = 1;
	for (int c = 2; c <= len; ++c) {
		int s = c;
		while (s > 1 && t.logodds[col[s - 1]][row[col[s - 1]][1]] < t.logodds[c][row[c][1]]) { col[s] = col[s - 1]; --s; }
		col[s] = c;
	}
	P.len = len;
	for (int c = 1; c <= len; ++c) {
		for (int j = 1; j <= 4; ++j) {
			P.sw[c][j] = t.logodds[col[c]][row[col[c]][j]];
			P.code[c][j] = row[col[c]][j] - 1;
		}
		P.shift[c] = 2 * (col[c] - 1);
	}
	P.smax[len + 1] = 0;
	for (int c = len; c > 0; --c) P.smax[c] = (c < len) ? P.smax[c + 1] + P.sw[c][1] : P.sw[c][1];
}

void build_pvalue_tables(const double* pwm, const int32_t* cols, int ncols, int64_t posSetSize, int64_t totalPositions, const double* bgLog,
                         const NullModelView& nm, PvalueTables& t, BgProbCache* cache) {
	Builder b(t, nm, cache);
	model_header_impl(pwm, cols, ncols, posSetSize, totalPositions, bgLog, t, b.t_lin_);
	if (t.nparts == 1) {
		b.build_part(0, 1, t.L, 0, false, 0);
		return;
	}
	b.clear_prefixes();
	b.build_part(0, 1, t.maxk, 0, true, 0);
	b.filter_prefixes();
	const int start = t.maxk + 1, stop = t.L;
	const int thr = std::min(0, static_cast<int>((-t.maxk + stop - start + 1) * 1e6));
	b.build_part(1, start, stop, thr, false, t.maxk);
	// per left class, the background probability mass in list order (branch_and_bound-inl.h:442-445)
	const TablePart& P0 = t.part[0];
	t.leftClassProb.assign(P0.classScore.size(), 0.0);
	size_t k = 0;
	for (size_t c = 0; c < P0.classScore.size(); ++c) {
		double lp = 0;
		for (int e = 0; e < P0.classCount[c]; ++e) lp += P0.prob[k++];
		t.leftClassProb[c] = lp;
	}
}

}  // namespace xxm
