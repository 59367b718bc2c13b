// xxm_multi: ONE host thread driving SEVERAL GPUs (include/xxmotif_gpu.h, "multi-device context").
//
// The reference is a single-threaded program; the C ABI proper is one context per GPU.  This layer closes the
// gap for the calls the reference-side binding makes (integration/xxmotif_gpu_binding.cpp): it owns one context per device,
// cuts the sequence set into contiguous blocks balanced by residues, runs every call on all contexts concurrently (one worker
// thread per device), and presents the union: counts summed, site lists concatenated in sequence order (which is the
// reference's scan order) or merged stably by score.  The shards are told the counts of the whole set
// (xxm_set_global_counts; the running log sum of avgLength is chained through them in order), so tables, the ZOOPS exponent
// and maxSites are the whole set's.  No collective is needed: whatever the reference computes over all sites it computes
// itself, from the lists this layer hands back.  Plain C++ over the public ABI -- no CUDA in this file.
#include "../../include/xxmotif_gpu.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {
// One worker thread per device beyond the first, alive as long as the context: the caller's thread runs shard 0 itself.  A worker
// spins for a short while after a job (the calls of a refinement step come in bursts) and then sleeps on a condition variable.
class Workers {
public:
	explicit Workers(int n) : rc_(static_cast<size_t>(n), 0) {
		for (int r = 1; r < n; ++r) th_.emplace_back([this, r]() { loop(r); });
	}
	~Workers() {
		{
			std::lock_guard<std::mutex> g(mu_);
			stop_ = true;
			gen_.fetch_add(1);
		}
		start_.notify_all();
		for (auto& t : th_) t.join();
	}
	// f(r) for every r concurrently; returns the index of the first shard that failed, or -1
	int run(const std::function<int(int)>& f) {
		const int n = static_cast<int>(rc_.size());
		job_ = &f;
		pending_.store(n - 1, std::memory_order_release);
		if (n > 1) {
			{
				std::lock_guard<std::mutex> g(mu_);
				gen_.fetch_add(1, std::memory_order_release);
			}
			start_.notify_all();
		}
		rc_[0] = f(0);
		for (int spin = 0; spin < kSpin && pending_.load(std::memory_order_acquire); ++spin) std::this_thread::yield();
		if (pending_.load(std::memory_order_acquire)) {
			std::unique_lock<std::mutex> g(mu_);
			done_.wait(g, [&]() { return pending_.load(std::memory_order_acquire) == 0; });
		}
		for (int r = 0; r < n; ++r)
			if (rc_[r]) return r;
		return -1;
	}

private:
	static constexpr int kSpin = 2000;
	void loop(int r) {
		uint64_t seen = 0;
		for (;;) {
			for (int spin = 0; spin < kSpin && gen_.load(std::memory_order_acquire) == seen; ++spin) std::this_thread::yield();
			if (gen_.load(std::memory_order_acquire) == seen) {
				std::unique_lock<std::mutex> g(mu_);
				start_.wait(g, [&]() { return gen_.load(std::memory_order_acquire) != seen; });
			}
			seen = gen_.load(std::memory_order_acquire);
			if (stop_) return;
			rc_[static_cast<size_t>(r)] = (*job_)(r);
			if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
				std::lock_guard<std::mutex> g(mu_);
				done_.notify_one();
			}
		}
	}
	std::vector<std::thread> th_;
	std::vector<int> rc_;
	std::mutex mu_;
	std::condition_variable start_, done_;
	std::atomic<uint64_t> gen_{0};
	std::atomic<int> pending_{0};
	const std::function<int(int)>* job_ = nullptr;
	bool stop_ = false;
};
}  // namespace

struct xxm_multi {
	std::vector<xxm_ctx*> ctx;
	std::vector<int32_t> first;  // first[r] = 0-based index of the first sequence of shard r; first[n] = nent
	int32_t nent = 0;
	int revcomp = 0;
	std::string err;
	int fail(const char* fmt, ...) {
		char buf[512];
		va_list ap;
		va_start(ap, fmt);
		vsnprintf(buf, sizeof buf, fmt, ap);
		va_end(ap);
		err = buf;
		return 1;
	}
	std::unique_ptr<Workers> workers;
	std::vector<std::unique_ptr<xxm_site[]>> tmp;  // per-shard staging for lists merged by score (uninitialised, grown on demand)
	std::vector<size_t> tmpCap;
	// run f(r) for every shard concurrently; the first failure is reported
	int each(const std::function<int(int)>& f) {
		const int bad = workers->run(f);
		if (bad >= 0) { err = std::string("device ") + std::to_string(bad) + ": " + xxm_last_error(ctx[bad]); return 1; }
		return 0;
	}
	int shards() const { return static_cast<int>(ctx.size()); }
	int32_t local_n(int r) const { return first[r + 1] - first[r]; }
};

static std::string g_multi_error;

extern "C" int xxm_multi_create_on(const int32_t* devices, int32_t ngpus, xxm_multi** out) {
	*out = nullptr;
	if (ngpus < 1 || ngpus > 64) { g_multi_error = "xxm_multi_create: 1..64 devices"; return 1; }
	xxm_multi* m = new xxm_multi();
	for (int d = 0; d < ngpus; ++d) {
		xxm_ctx* c = nullptr;
		if (xxm_create(devices ? devices[d] : d, &c)) {
			g_multi_error = std::string("device ") + std::to_string(devices ? devices[d] : d) + ": " + xxm_last_error(nullptr);
			for (xxm_ctx* x : m->ctx) xxm_destroy(x);
			delete m;
			return 1;
		}
		m->ctx.push_back(c);
	}
	m->workers.reset(new Workers(ngpus));
	m->tmp.resize(static_cast<size_t>(ngpus));
	m->tmpCap.assign(static_cast<size_t>(ngpus), 0);
	*out = m;
	return 0;
}
extern "C" int xxm_multi_create(int32_t ngpus, xxm_multi** out) { return xxm_multi_create_on(nullptr, ngpus, out); }

extern "C" void xxm_multi_destroy(xxm_multi* m) {
	if (!m) return;
	m->workers.reset();
	for (xxm_ctx* c : m->ctx) xxm_destroy(c);
	delete m;
}

extern "C" const char* xxm_multi_last_error(const xxm_multi* m) { return m ? m->err.c_str() : g_multi_error.c_str(); }
extern "C" int32_t xxm_multi_devices(const xxm_multi* m) { return m->shards(); }

extern "C" int xxm_multi_set_options(xxm_multi* m, const xxm_options* opt) {
	return m->each([&](int r) { return xxm_set_options(m->ctx[r], opt); });
}
extern "C" int xxm_multi_set_background(xxm_multi* m, int32_t order, const float* rprobs, const float* rconds) {
	return m->each([&](int r) { return xxm_set_background(m->ctx[r], order, rprobs, rconds); });
}

extern "C" int xxm_multi_load_seqset(xxm_multi* m, int32_t nent, const int32_t* n, const uint8_t* codes, int32_t revcomp) {
	const int S = m->shards();
	if (nent < S) return m->fail("fewer sequences than devices");
	// contiguous blocks of sequences balanced by residues (contiguity keeps the global (sequence, position) order)
	std::vector<int64_t> off(static_cast<size_t>(nent) + 1, 0);
	for (int i = 0; i < nent; ++i) off[i + 1] = off[i] + n[i];
	m->first.assign(static_cast<size_t>(S) + 1, 0);
	for (int r = 1; r < S; ++r) {
		const int64_t want = off[nent] * r / S;
		int32_t i = static_cast<int32_t>(std::lower_bound(off.begin(), off.end(), want) - off.begin());
		i = std::max<int32_t>(i, m->first[r - 1] + 1);
		i = std::min<int32_t>(i, nent - (S - r));
		m->first[r] = i;
	}
	m->first[S] = nent;
	m->nent = nent;
	m->revcomp = revcomp;
	if (m->each([&](int r) { return xxm_load_seqset(m->ctx[r], m->local_n(r), n + m->first[r], codes + off[m->first[r]], revcomp); })) return 1;
	if (S == 1) return 0;
	// the counts of the whole set: totals and max_leng from the shards, the log sum of avgLength chained through them in order
	int64_t total = 0;
	int32_t mxl = 0;
	double acc = 0;
	for (int r = 0; r < S; ++r) {
		std::vector<int32_t> tot(static_cast<size_t>(m->local_n(r)) + 1);
		int32_t mx = 0, mn = 0;
		if (xxm_seqset_info(m->ctx[r], tot.data(), nullptr, &mx, &mn)) return m->fail("%s", xxm_last_error(m->ctx[r]));
		total += tot.back();
		mxl = std::max(mxl, mx);
		if (xxm_seqset_logsum_from(m->ctx[r], acc, &acc)) return m->fail("%s", xxm_last_error(m->ctx[r]));
	}
	if (total > 2147483647LL) return m->fail("more than 2^31-1 residues: the reference's int totals overflow");
	return m->each([&](int r) { return xxm_set_global_counts(m->ctx[r], nent, total, acc, mxl); });
}

extern "C" int xxm_multi_seqset_info(const xxm_multi* m, int32_t* total, double* avgLength, int32_t* max_leng, int32_t* min_leng) {
	// totals concatenated with running offsets; avgLength only for the whole set (index nent) -- the reference reads nothing else
	int64_t base = 0;
	int32_t mxl = 0, mnl = std::numeric_limits<int32_t>::max();
	if (total) total[0] = 0;
	for (int r = 0; r < m->shards(); ++r) {
		std::vector<int32_t> tot(static_cast<size_t>(m->local_n(r)) + 1);
		int32_t mx = 0, mn = 0;
		if (xxm_seqset_info(m->ctx[r], tot.data(), nullptr, &mx, &mn)) return 1;
		if (total)
			for (int32_t i = 1; i <= m->local_n(r); ++i) total[m->first[r] + i] = static_cast<int32_t>(base + tot[i]);
		base += tot.back();
		mxl = std::max(mxl, mx);
		mnl = std::min(mnl, mn);
	}
	(void) avgLength;
	if (max_leng) *max_leng = mxl;
	if (min_leng) *min_leng = mnl;
	return 0;
}

static int whole_set_only(xxm_multi* m, int32_t posSetSize) {
	if (m->shards() > 1 && posSetSize != m->nent) return m->fail("with several devices every call refers to the whole set (posSetSize = number of sequences)");
	return 0;
}

extern "C" int xxm_multi_reinitialize(xxm_multi* m, const double* pwm, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t totalPositions) {
	if (whole_set_only(m, posSetSize)) return 1;
	return m->each([&](int r) { return xxm_reinitialize(m->ctx[r], pwm, cols, ncols, m->shards() == 1 ? posSetSize : m->local_n(r), totalPositions); });
}

extern "C" int xxm_multi_scan_full(xxm_multi* m, int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t* nbSites, int32_t* nbBestSites, int32_t* maxSites) {
	if (whole_set_only(m, posSetSize)) return 1;
	const int S = m->shards();
	std::vector<int32_t> a(S), b(S), c(S);
	if (m->each([&](int r) { return xxm_scan_full(m->ctx[r], S == 1 ? posSetSize : m->local_n(r), enrStart, enrEnd, &a[r], &b[r], &c[r]); })) return 1;
	int64_t sa = 0, sb = 0;
	for (int r = 0; r < S; ++r) { sa += a[r]; sb += b[r]; }
	*nbSites = static_cast<int32_t>(sa);
	*nbBestSites = static_cast<int32_t>(sb);
	*maxSites = c[0];  // every shard reports the whole set's
	return 0;
}

extern "C" int xxm_multi_rescore(xxm_multi* m, int32_t firstMotifColumn, int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t palin, int32_t* nbSites,
                                 int32_t* maxSites) {
	if (whole_set_only(m, posSetSize)) return 1;
	const int S = m->shards();
	std::vector<int32_t> a(S), c(S);
	if (m->each([&](int r) { return xxm_rescore(m->ctx[r], firstMotifColumn, S == 1 ? posSetSize : m->local_n(r), enrStart, enrEnd, palin, &a[r], &c[r]); })) return 1;
	int64_t sa = 0;
	for (int r = 0; r < S; ++r) sa += a[r];
	*nbSites = static_cast<int32_t>(sa);
	*maxSites = c[0];
	return 0;
}

extern "C" int xxm_multi_scan_oops(xxm_multi* m, int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t* nbSites, int32_t* maxSites) {
	if (whole_set_only(m, posSetSize)) return 1;
	const int S = m->shards();
	std::vector<int32_t> a(S), c(S);
	if (m->each([&](int r) { return xxm_scan_oops(m->ctx[r], S == 1 ? posSetSize : m->local_n(r), enrStart, enrEnd, &a[r], &c[r]); })) return 1;
	int64_t sa = 0;
	for (int r = 0; r < S; ++r) sa += a[r];
	*nbSites = static_cast<int32_t>(sa);
	*maxSites = c[0];
	return 0;
}

extern "C" int xxm_multi_rescore_startposlist(xxm_multi* m, int32_t posSetSize, int32_t enrStart, int32_t enrEnd, int32_t nSites, const int32_t* siteSeq,
                                              const int32_t* sitePos, int32_t instanceThreshold, int32_t* nbSites, int32_t* maxSites) {
	if (whole_set_only(m, posSetSize)) return 1;
	const int S = m->shards();
	// every site goes to the shard that owns its sequence (list order kept inside a shard)
	std::vector<std::vector<int32_t>> sq(S), ps(S);
	for (int32_t i = 0; i < nSites; ++i) {
		if (siteSeq[i] < 1 || siteSeq[i] > m->nent) return m->fail("site %d: sequence %d out of range", i, siteSeq[i]);
		const int r = static_cast<int>(std::upper_bound(m->first.begin(), m->first.end(), siteSeq[i] - 1) - m->first.begin()) - 1;
		sq[r].push_back(siteSeq[i] - m->first[r]);
		ps[r].push_back(sitePos[i]);
	}
	std::vector<int32_t> a(S), c(S);
	if (m->each([&](int r) {
		    return xxm_rescore_startposlist(m->ctx[r], S == 1 ? posSetSize : m->local_n(r), enrStart, enrEnd, static_cast<int32_t>(sq[r].size()), sq[r].data(),
		                                    ps[r].data(), instanceThreshold, &a[r], &c[r]);
	    }))
		return 1;
	int64_t sa = 0;
	for (int r = 0; r < S; ++r) sa += a[r];
	*nbSites = static_cast<int32_t>(sa);
	*maxSites = c[0];
	return 0;
}

// the union of a device-resident list: shard lists concatenated in sequence order (as-is / by start position) or merged stably
// by score (ties keep sequence order, like the reference's stable sort over its scan-ordered array)
extern "C" int xxm_multi_get_sites(xxm_multi* m, int32_t which, int32_t order, int32_t palin, xxm_site* out, int32_t cap, int32_t* nOut) {
	const int S = m->shards();
	if (S == 1) {
		if (xxm_get_sites(m->ctx[0], which, order, palin, out, cap, nOut)) return m->fail("%s", xxm_last_error(m->ctx[0]));
		return 0;
	}
	std::vector<int32_t> cnt(static_cast<size_t>(S), 0);
	std::vector<int64_t> off(static_cast<size_t>(S) + 1, 0);
	for (int r = 0; r < S; ++r) {
		if (xxm_sites_count(m->ctx[r], which, order, &cnt[r])) return m->fail("%s", xxm_last_error(m->ctx[r]));
		off[r + 1] = off[r] + cnt[r];
	}
	const int64_t tot = off[S];
	*nOut = static_cast<int32_t>(tot);
	if (tot > cap) return m->fail("xxm_multi_get_sites: output capacity %d < %lld sites", cap, static_cast<long long>(tot));
	const bool merge = (order == XXM_ORDER_BY_SCORE);
	if (merge)
		for (int r = 0; r < S; ++r)
			if (m->tmpCap[r] < static_cast<size_t>(cnt[r])) {
				m->tmpCap[r] = static_cast<size_t>(cnt[r]) + static_cast<size_t>(cnt[r]) / 4 + 1;
				m->tmp[r].reset(new xxm_site[m->tmpCap[r]]);
			}
	// every shard writes its block (straight into the output unless the blocks have to be merged) and shifts its sequence numbers
	if (m->each([&](int r) {
		    xxm_site* dst = merge ? m->tmp[r].get() : out + off[r];
		    int32_t k = 0;
		    if (xxm_get_sites(m->ctx[r], which, order, palin, dst, cnt[r], &k)) return 1;
		    if (k != cnt[r]) return 1;
		    const int32_t add = m->first[r];
		    for (int32_t i = 0; i < k; ++i) dst[i].sequence += add;
		    return 0;
	    }))
		return 1;
	if (!merge) return 0;
	// P pieces of the score axis merged concurrently: thresholds from a sample, the bounds of a piece by bisection in every block
	const int P = static_cast<int>(std::min<int64_t>(std::max<int64_t>(1, tot / 65536), std::max(S, std::min(16, static_cast<int>(std::thread::hardware_concurrency()) / 2))));
	std::vector<float> sample;
	for (int r = 0; r < S; ++r)
		for (int k = 0; k < 256 && cnt[r]; ++k) sample.push_back(m->tmp[r][static_cast<size_t>((static_cast<int64_t>(cnt[r]) * k) / 256)].score);
	std::sort(sample.begin(), sample.end());
	std::vector<std::vector<int32_t>> lo(static_cast<size_t>(P) + 1, std::vector<int32_t>(static_cast<size_t>(S), 0));
	for (int p = 1; p < P; ++p) {
		if (sample.empty()) break;
		const float t = sample[sample.size() * static_cast<size_t>(p) / static_cast<size_t>(P)];
		for (int r = 0; r < S; ++r) {
			const xxm_site* b = m->tmp[r].get();
			lo[p][r] = static_cast<int32_t>(std::lower_bound(b, b + cnt[r], t, [](const xxm_site& a, float v) { return a.score < v; }) - b);
		}
	}
	for (int r = 0; r < S; ++r) lo[P][r] = cnt[r];
	auto piece = [&](int p) {
		int64_t o = 0;
		for (int r = 0; r < S; ++r) o += lo[p][r];
		std::vector<int32_t> at(lo[p]);
		const std::vector<int32_t>& end = lo[p + 1];
		for (;;) {
			int best = -1;
			float bs = 0;
			for (int r = 0; r < S; ++r)
				if (at[r] < end[r]) {
					const float sc = m->tmp[r][at[r]].score;
					if (best < 0 || sc < bs) { best = r; bs = sc; }
				}
			if (best < 0) break;
			out[o++] = m->tmp[best][at[best]++];
		}
	};
	if (P == 1) piece(0);
	else {
		std::vector<std::thread> th;
		for (int p = 1; p < P; ++p) th.emplace_back(piece, p);
		piece(0);
		for (auto& t : th) t.join();
	}
	return 0;
}

// every shard counts its sequences, the counts are added up
extern "C" int xxm_multi_count_gapped_kmers(xxm_multi* m, int32_t nseq, int32_t ncoding, const int32_t* offset, int32_t firstHalfOnly, uint32_t* counts) {
	const int S = m->shards();
	if (S == 1) {
		if (xxm_count_gapped_kmers(m->ctx[0], nseq, ncoding, offset, firstHalfOnly, counts)) return m->fail("%s", xxm_last_error(m->ctx[0]));
		return 0;
	}
	if (whole_set_only(m, nseq)) return 1;
	if (ncoding < 1 || ncoding > 7) return m->fail("xxm_multi_count_gapped_kmers: 1..7 coding characters");
	const size_t nb = static_cast<size_t>(1) << (2 * ncoding);
	std::vector<std::vector<uint32_t>> part(static_cast<size_t>(S), std::vector<uint32_t>(nb));
	if (m->each([&](int r) { return xxm_count_gapped_kmers(m->ctx[r], m->local_n(r), ncoding, offset, firstHalfOnly, part[r].data()); })) return 1;
	for (size_t i = 0; i < nb; ++i) {
		uint32_t s = 0;
		for (int r = 0; r < S; ++r) s += part[r][i];
		counts[i] = s;
	}
	return 0;
}
