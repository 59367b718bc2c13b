// TEST INFRASTRUCTURE ONLY: csrc/refine.cpp (the callers of the hot path) linked against a stand-in for the data-sized xxm_* calls,
// so that the sequencing and the threading of xxm_iterate_motifs can be exercised in a container without a GPU.
// The stand-in is deterministic in its INPUTS (PWM, columns, list state of the context) and sleeps a little inside every call, so
// that the contexts' threads really interleave; a motif refined by xxm_iterate_motifs on 1..6 contexts must end exactly where
// the serial loop leaves it, failures must be reported for the right motif, and no context may ever be entered by two threads.
#include "../../include/xxmotif_gpu.h"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

struct xxm_ctx {
	std::atomic<int> inside{0};
	std::atomic<int> violations{0};
	uint64_t model = 0;      // hash of the last reinitialize
	uint64_t list = 0;       // hash of the last re-score / scan
	uint64_t cand = 0;       // candidate list of the last full scan (bestStartPos), or the one borrowed from the owner
	uint64_t kept = 0;       // list set aside by xxm_elong_keep
	int calls = 0;
	int failOnModel = -1;    // xxm_reinitialize fails when the PWM carries this marker
	std::string err;
};

namespace {
struct Enter {
	xxm_ctx* c;
	explicit Enter(xxm_ctx* ctx) : c(ctx) {
		if (c->inside.fetch_add(1) != 0) c->violations.fetch_add(1);
		c->calls++;
		std::this_thread::sleep_for(std::chrono::microseconds(20 + (c->calls * 7) % 60));
	}
	~Enter() { c->inside.fetch_sub(1); }
};
uint64_t mix(uint64_t h, uint64_t v) {
	h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
	return h * 0xBF58476D1CE4E5B9ull;
}
uint64_t hash_doubles(const double* p, int n, uint64_t h) {
	for (int i = 0; i < n; ++i) { uint64_t b; std::memcpy(&b, &p[i], 8); h = mix(h, b); }
	return h;
}
}  // namespace

extern "C" {
const char* xxm_last_error(const xxm_ctx* ctx) { return ctx ? ctx->err.c_str() : ""; }
int xxm_reinitialize(xxm_ctx* ctx, const double* pwm, const int32_t* cols, int32_t ncols, int32_t posSetSize, int32_t totalPositions) {
	Enter e(ctx);
	if (ctx->failOnModel >= 0 && pwm[0] == static_cast<double>(ctx->failOnModel)) { ctx->err = "stand-in failure"; return 1; }
	uint64_t h = mix(posSetSize, totalPositions);
	for (int i = 0; i < ncols; ++i) { h = mix(h, static_cast<uint64_t>(cols[i])); h = hash_doubles(pwm + cols[i] * 5 + 1, 4, h); }
	ctx->model = h;
	return 0;
}
int xxm_scan_full(xxm_ctx* ctx, int32_t posSetSize, int32_t, int32_t, int32_t* nb, int32_t* nbest, int32_t* mx) {
	Enter e(ctx);
	ctx->list = mix(ctx->model, 1);
	ctx->cand = mix(ctx->model, 77);
	*nb = static_cast<int32_t>(ctx->list % 1000) + 1; *nbest = *nb * 2; *mx = posSetSize;
	return 0;
}
int xxm_rescore(xxm_ctx* ctx, int32_t firstMotifColumn, int32_t posSetSize, int32_t, int32_t, int32_t, int32_t* nb, int32_t* mx) {
	Enter e(ctx);
	if (!ctx->cand) { ctx->err = "no candidate list"; return 1; }
	ctx->list = mix(mix(ctx->model, 1000 + firstMotifColumn), ctx->cand);
	*nb = static_cast<int32_t>(ctx->list % 1000) + 1; *mx = posSetSize;
	return 0;
}
int xxm_update_pwm_with_best_subset(xxm_ctx* ctx, int32_t which, double* pwm, int32_t firstColumn, int32_t, int32_t plus, double pseudo, int32_t* updated,
                                    int32_t* bestMotifNb, double* bestPval) {
	Enter e(ctx);
	const uint64_t h = mix(mix(ctx->list, which * 31 + plus), static_cast<uint64_t>(firstColumn));
	*updated = (h % 11) != 0;  // now and then the reference returns early
	if (!*updated) return 0;
	*bestMotifNb = static_cast<int32_t>(h % 500) + (h % 13 == 0 ? 0 : 1);
	*bestPval = -static_cast<double>(h % 100000) / 7.0 - pseudo;
	for (int c = 1; c <= 30; ++c) {  // a normalised column derived from the list
		double w[4], s = 0;
		for (int a = 0; a < 4; ++a) { w[a] = 1.0 + static_cast<double>(mix(h, c * 4 + a) % 97); s += w[a]; }
		for (int a = 0; a < 4; ++a) pwm[c * 5 + 1 + a] = std::log(w[a] / s);
	}
	return 0;
}
int xxm_share_candidates(xxm_ctx* ctx, xxm_ctx* owner) {
	Enter e(ctx);
	if (owner->inside.load() != 0) ctx->violations.fetch_add(1);  // the owner must be idle while its list is lent
	ctx->cand = owner->cand;
	ctx->kept = 0;
	return 0;
}
int xxm_elong_keep(xxm_ctx* ctx) { Enter e(ctx); ctx->kept = ctx->list; return 0; }
int xxm_startposlist_from_kept(xxm_ctx* ctx, int32_t bestMotifNb, int32_t firstTmpColumn, int32_t* outSeq, int32_t* outPos, int32_t cap, int32_t* nOut,
                               double* worstScore) {
	Enter e(ctx);
	const int n = std::min(cap, bestMotifNb);
	for (int i = 0; i < n; ++i) { outSeq[i] = static_cast<int32_t>(mix(ctx->kept, i) % 900) + 1; outPos[i] = static_cast<int32_t>(mix(ctx->kept, 7 * i + 3) % 200) - firstTmpColumn + 1; }
	*nOut = n;
	*worstScore = 1.0 / static_cast<double>(1 + ctx->kept % 1000);
	return 0;
}
int xxm_update_enriched_region(xxm_ctx* ctx, int32_t first, int32_t len, int32_t* enr) { Enter e(ctx); enr[0] += 0 * (first + len); return 0; }
int xxm_positional_pval(xxm_ctx* ctx, int32_t enrMax, int32_t first, int32_t len, const int32_t* sitePos, int32_t nSites, double* out) {
	Enter e(ctx);
	uint64_t h = mix(enrMax, first * 64 + len);
	for (int i = 0; i < nSites; ++i) h = mix(h, static_cast<uint64_t>(static_cast<uint32_t>(sitePos[i])));
	*out = -static_cast<double>(h % 1000003) / 13.0;
	return 0;
}
}  // extern "C"

namespace {
struct Holder {
	xxm_motif m{};
	std::vector<int32_t> seq, pos;
	void init(int k) {
		std::memset(&m, 0, sizeof m);
		for (int c = 1; c <= 30; ++c)
			for (int a = 1; a <= 4; ++a) m.pwm[c * 5 + a] = std::log(0.25) + 1e-3 * ((k * 131 + c * 7 + a) % 17);
		m.pwm[0] = k;  // marker the stand-in can fail on
		m.ncols = 6 + k % 5;
		const int first = 8 + k % 7;
		for (int i = 0; i < m.ncols; ++i) m.cols[i] = first + i;
		m.posSetSize = 900;
		m.enr[0] = m.enr[1] = 1 - first; m.enr[2] = 200 - first;
		seq.assign(4096, 0); pos.assign(4096, 0);
		m.siteCap = 4096; m.siteSeq = seq.data(); m.sitePos = pos.data();
	}
	bool same(const Holder& o) const {
		return m.ncols == o.m.ncols && !std::memcmp(m.cols, o.m.cols, sizeof(int32_t) * m.ncols) && !std::memcmp(m.pwm, o.m.pwm, sizeof m.pwm) && m.pval == o.m.pval &&
		       m.posPval == o.m.posPval && m.worstScore == o.m.worstScore && m.nsites == o.m.nsites && !std::memcmp(m.enr, o.m.enr, sizeof m.enr) &&
		       std::equal(seq.begin(), seq.begin() + m.nsites, o.seq.begin()) && std::equal(pos.begin(), pos.begin() + m.nsites, o.pos.begin());
	}
};
}  // namespace

int main() {
	const int nmotifs = 23;
	xxm_refine_params P{};
	P.totalPositions = 180000; P.maximizeMotifLength = 1; P.minMatchPositions = 5; P.maxMatchPositions = 17; P.pseudo = 0.1; P.neff_pwm = 10.0;
	int bad = 0;
	// the serial loop
	std::vector<Holder> serial(nmotifs);
	std::vector<int32_t> roundsSerial(nmotifs, 0);
	{
		xxm_ctx c;
		for (int i = 0; i < nmotifs; ++i) {
			serial[i].init(i);
			if (xxm_iterate_motif(&c, &serial[i].m, &P, &roundsSerial[i])) { std::printf("serial motif %d failed\n", i); return 1; }
		}
	}
	int distinct = 0;
	for (int i = 1; i < nmotifs; ++i) distinct += !serial[i].same(serial[0]);
	if (distinct < nmotifs / 2) { std::printf("stand-in too uniform\n"); return 1; }
	for (int nctx = 1; nctx <= 6; ++nctx) {
		std::vector<xxm_ctx> ctxs(nctx);
		std::vector<xxm_ctx*> ptr(nctx);
		for (int i = 0; i < nctx; ++i) ptr[i] = &ctxs[i];
		std::vector<Holder> h(nmotifs);
		std::vector<xxm_motif> arr(nmotifs);
		std::vector<xxm_refine_params> Ps(nmotifs, P);
		for (int i = 0; i < nmotifs; ++i) { h[i].init(i); arr[i] = h[i].m; }
		std::vector<int32_t> rounds(nmotifs, -1);
		int32_t failed = 5, failedCtx = 5;
		const bool perMotif = nctx % 2 == 0;  // one parameter block for all, or one per motif
		const int rc = xxm_iterate_motifs(ptr.data(), nctx, arr.data(), nmotifs, perMotif ? Ps.data() : &P, perMotif ? nmotifs : 1, rounds.data(), &failed, &failedCtx);
		if (rc || failed != -1 || failedCtx != -1) { std::printf("nctx %d: rc %d failed %d\n", nctx, rc, failed); ++bad; continue; }
		int used = 0;
		for (int i = 0; i < nctx; ++i) {
			if (ctxs[i].violations.load()) { std::printf("nctx %d: context %d entered by two threads\n", nctx, i); ++bad; }
			used += ctxs[i].calls > 0;
		}
		if (used != std::min(nctx, nmotifs)) { std::printf("nctx %d: only %d contexts worked\n", nctx, used); ++bad; }
		for (int i = 0; i < nmotifs; ++i) {
			h[i].m = arr[i];
			if (!h[i].same(serial[i]) || rounds[i] != roundsSerial[i]) { std::printf("nctx %d: motif %d differs from the serial loop\n", nctx, i); ++bad; }
		}
	}
	// a failing motif: reported with its index and the context that ran it; motifs before it are complete
	{
		const int nctx = 3, victim = 9;
		std::vector<xxm_ctx> ctxs(nctx);
		std::vector<xxm_ctx*> ptr(nctx);
		for (int i = 0; i < nctx; ++i) { ptr[i] = &ctxs[i]; ctxs[i].failOnModel = victim; }
		std::vector<Holder> h(nmotifs);
		std::vector<xxm_motif> arr(nmotifs);
		for (int i = 0; i < nmotifs; ++i) { h[i].init(i); arr[i] = h[i].m; }
		std::vector<int32_t> rounds(nmotifs, -1);
		int32_t failed = -1, failedCtx = -1;
		const int rc = xxm_iterate_motifs(ptr.data(), nctx, arr.data(), nmotifs, &P, 1, rounds.data(), &failed, &failedCtx);
		if (rc != 1 || failed != victim || failedCtx < 0 || failedCtx >= nctx || std::string(xxm_last_error(ptr[failedCtx])) != "stand-in failure") {
			std::printf("failure case: rc %d failed %d ctx %d\n", rc, failed, failedCtx);
			++bad;
		}
		for (int i = 0; i < victim - nctx; ++i) {  // started well before the failure: finished
			h[i].m = arr[i];
			if (!h[i].same(serial[i])) { std::printf("failure case: motif %d incomplete\n", i); ++bad; }
		}
		int untouched = 0;
		for (int i = victim + 1; i < nmotifs; ++i) untouched += rounds[i] == -1;
		if (untouched < nmotifs - victim - 1 - nctx) { std::printf("failure case: %d motifs started after the failure\n", nmotifs - victim - 1 - untouched); ++bad; }
	}
	// the column variants of one motif spread over 2..5 contexts: the same motif as the serial loop, every time
	for (int nctx = 2; nctx <= 5; ++nctx) {
		std::vector<xxm_ctx> ctxs(nctx);
		std::vector<xxm_ctx*> ptr(nctx);
		for (int i = 0; i < nctx; ++i) ptr[i] = &ctxs[i];
		for (int i = 0; i < nmotifs; ++i) {
			Holder h;
			h.init(i);
			int32_t r = -1;
			const int rc = xxm_iterate_motif_on(ptr.data(), nctx, &h.m, &P, &r);
			if (rc || r != roundsSerial[i] || !h.same(serial[i])) { std::printf("variants on %d contexts: motif %d differs (rc %d)\n", nctx, i, rc); ++bad; }
		}
		int used = 0;
		for (int i = 0; i < nctx; ++i) {
			if (ctxs[i].violations.load()) { std::printf("variants on %d contexts: context %d entered by two threads\n", nctx, i); ++bad; }
			used += ctxs[i].calls > 0;
		}
		if (used != nctx) { std::printf("variants on %d contexts: only %d worked\n", nctx, used); ++bad; }
	}
	// ... also without length correction (one variant) and with a failing table build
	{
		xxm_ctx a, b;
		xxm_ctx* ptr[2] = {&a, &b};
		xxm_refine_params Q = P;
		Q.maximizeMotifLength = 0;
		xxm_ctx c;
		for (int i = 0; i < 5; ++i) {
			Holder h1, h2;
			h1.init(i); h2.init(i);
			int32_t r1 = -1, r2 = -1;
			if (xxm_iterate_motif(&c, &h1.m, &Q, &r1) || xxm_iterate_motif_on(ptr, 2, &h2.m, &Q, &r2) || r1 != r2 || !h1.same(h2)) { std::printf("one-variant case: motif %d differs\n", i); ++bad; }
		}
		a.failOnModel = b.failOnModel = 3;
		Holder h;
		h.init(3);
		int32_t r = -1;
		if (xxm_iterate_motif_on(ptr, 2, &h.m, &P, &r) != 1) { std::printf("failing table build not reported\n"); ++bad; }
		// elongPWM carries 0 in its unused entry [0][0]: with this marker the build of iteration 1 fails INSIDE a variant, on some context
		a.failOnModel = b.failOnModel = 0;
		Holder g;
		g.init(4);
		if (xxm_iterate_motif_on(ptr, 2, &g.m, &P, &r) != 1 || (std::string(xxm_last_error(&a)) != "stand-in failure" && std::string(xxm_last_error(&b)) != "stand-in failure")) {
			std::printf("failing variant not reported\n");
			++bad;
		}
	}
	std::printf("refine batch: %d motifs on 1..6 contexts, %d problems\n", nmotifs, bad);
	return bad ? 1 : 0;
}
