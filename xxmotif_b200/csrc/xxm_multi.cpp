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
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <string>
#include <thread>
#include <vector>

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
	// run f(r) for every shard concurrently; the first failure is reported
	int each(const std::function<int(int)>& f) {
		const int n = static_cast<int>(ctx.size());
		std::vector<int> rc(n, 0);
		if (n == 1) rc[0] = f(0);
		else {
			std::vector<std::thread> th;
			for (int r = 0; r < n; ++r) th.emplace_back([&, r]() { rc[r] = f(r); });
			for (auto& t : th) t.join();
		}
		for (int r = 0; r < n; ++r)
			if (rc[r]) { err = std::string("device ") + std::to_string(r) + ": " + xxm_last_error(ctx[r]); return 1; }
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
	*out = m;
	return 0;
}
extern "C" int xxm_multi_create(int32_t ngpus, xxm_multi** out) { return xxm_multi_create_on(nullptr, ngpus, out); }

extern "C" void xxm_multi_destroy(xxm_multi* m) {
	if (!m) return;
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
// by score (ties keep sequence order, like the reference's stable qsort over its scan-ordered array)
extern "C" int xxm_multi_get_sites(xxm_multi* m, int32_t which, int32_t order, int32_t palin, xxm_site* out, int32_t cap, int32_t* nOut) {
	const int S = m->shards();
	if (S == 1) {
		if (xxm_get_sites(m->ctx[0], which, order, palin, out, cap, nOut)) return m->fail("%s", xxm_last_error(m->ctx[0]));
		return 0;
	}
	std::vector<std::vector<xxm_site>> part(S);
	std::vector<int32_t> cnt(S, 0);
	if (m->each([&](int r) {
		    part[r].resize(static_cast<size_t>(std::max(cap, 1)));
		    const int rc = xxm_get_sites(m->ctx[r], which, order, palin, part[r].data(), cap, &cnt[r]);
		    return (rc && cnt[r] > cap) ? 0 : rc;  // too small an output is reported below, with the count of the whole list
	    }))
		return 1;
	int64_t tot = 0;
	for (int r = 0; r < S; ++r) {
		for (int32_t i = 0; i < cnt[r]; ++i) part[r][i].sequence += m->first[r];
		tot += cnt[r];
	}
	*nOut = static_cast<int32_t>(tot);
	if (tot > cap) return m->fail("xxm_multi_get_sites: output capacity %d < %lld sites", cap, static_cast<long long>(tot));
	if (order != XXM_ORDER_BY_SCORE) {
		int64_t o = 0;
		for (int r = 0; r < S; ++r) { std::copy(part[r].begin(), part[r].begin() + cnt[r], out + o); o += cnt[r]; }
		return 0;
	}
	std::vector<int32_t> at(S, 0);
	for (int64_t o = 0; o < tot; ++o) {
		int best = -1;
		for (int r = 0; r < S; ++r)
			if (at[r] < cnt[r] && (best < 0 || part[r][at[r]].score < part[best][at[best]].score)) best = r;
		out[o] = part[best][at[best]++];
	}
	return 0;
}
